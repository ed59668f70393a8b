// NVTX ranges carrying the reference's labels (dorado/torch_utils/include/torch_utils/gpu_profiling.h:32-99 and its call
// sites: "nn_forward" CRFModel.cpp:119, "conv" ConvStack.cpp:193, "lstm_stack"/"lstm_layer" LSTMStack.cpp:55,146,
// "linear" CRFModules.cpp:63, "Conv"/"TransEnc"/"TransDec"/"CRF" TxModel.cpp:23-36, "TxLayerKoiTiled", "QKV+ROTE", "MEA",
// "OUTP", "LNORM1", "FC1+SILU", "FC2", "LNORM2" TxModules.cpp:472-703, "gpu_decode", "back_guides", "beam_search", "decode"
// CUDADecoder.cpp:22-99), so an Nsight Systems capture of this engine lines up with one of the reference.  Header-only
// NVTX v3: a no-op costing one indirect call unless a tool is attached.
#pragma once

#include <nvtx3/nvToolsExt.h>

namespace b200 {

struct NvtxRange {
    explicit NvtxRange(const char* label) { nvtxRangePushA(label); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace b200
