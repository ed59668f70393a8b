// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// C-ABI shim around the *unmodified* reference CPU implementation (nanoporetech/dorado),
// compiled in place from /root/reference by oracle/Makefile into oracle/_ref/libdorado_ref.so.
// It exposes the reference's own functions so the C restatement (oracle/crf_oracle.c) and the
// CUDA engine can be checked against the real thing:
//   * model forward      dorado/basecall/model/{CRFModel,TxModel}.cpp  (libtorch CPU, fp32)
//   * CRF scans          dorado/basecall/decode/CPUDecoder.cpp:43-92   (inner::forward/backward_scores)
//   * beam search        dorado/basecall/decode/beam_search.cpp:522    (beam_search_decode)
//   * whole decode       dorado/basecall/decode/CPUDecoder.cpp:100     (CPUDecoder::beam_search_part_2)
//   * the CPU runner     dorado/basecall/ModelRunner.cpp:10-65       (ModelRunner: accept_chunk / call_chunks / sample_stats)
//   * front end          dorado/read_pipeline/base/chunk.cpp:11-47 (generate_chunks), stitch.cpp:12-96 (stitch_chunks),
//                        dorado/torch_utils/tensor_utils.cpp:254 (shift_scale_tensor_i16_to_f16_inplace)
// Everything here is glue written for this repo; no reference source is copied.
#include "basecall/ModelRunner.h"
#include "basecall/decode/CPUDecoder.h"
#include "basecall/decode/beam_search.h"
#include "basecall/model/CRFModel.h"
#include "basecall/model/TxModel.h"
#include "config/BasecallModelConfig.h"
#include "read_pipeline/base/chunk.h"
#include "read_pipeline/base/messages.h"
#include "read_pipeline/base/stitch.h"
#include "torch_utils/tensor_utils.h"

#include <ATen/ATen.h>
#include <torch/torch.h>

#include <torch/serialize.h>

#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct RefModel {
    dorado::config::BasecallModelConfig config;
    torch::nn::ModuleHolder<torch::nn::AnyModule> module{nullptr};
};

// "B2W1" flat weight container: see dorado_b200/weights.py
std::vector<at::Tensor> read_b2w(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        throw std::runtime_error("cannot open weights file " + path);
    }
    char magic[4];
    f.read(magic, 4);
    if (std::memcmp(magic, "B2W1", 4) != 0) {
        throw std::runtime_error("bad magic in " + path);
    }
    uint32_t n = 0;
    f.read(reinterpret_cast<char*>(&n), 4);
    std::vector<at::Tensor> out;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t name_len = 0, ndim = 0;
        f.read(reinterpret_cast<char*>(&name_len), 4);
        std::string name(name_len, '\0');
        f.read(name.data(), name_len);
        f.read(reinterpret_cast<char*>(&ndim), 4);
        std::vector<int64_t> dims(ndim);
        int64_t numel = 1;
        for (uint32_t d = 0; d < ndim; ++d) {
            uint32_t v = 0;
            f.read(reinterpret_cast<char*>(&v), 4);
            dims[d] = v;
            numel *= v;
        }
        at::Tensor t = at::empty(dims, at::kFloat);
        f.read(reinterpret_cast<char*>(t.data_ptr<float>()), numel * 4);
        if (!f) {
            throw std::runtime_error("truncated weights file " + path);
        }
        out.push_back(t);
    }
    return out;
}

template <typename F>
int guarded(F&& fn) {
    try {
        at::InferenceMode guard;
        fn();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_num_threads(int n) { at::set_num_threads(n); }

// config_dir: directory holding config.toml (directory name must be a non-deprecated model name).
// weights:    B2W1 file with tensors in the order of dorado/basecall/crf_utils.cpp:26-150.
void* ref_model_create(const char* config_dir, const char* weights) {
    RefModel* m = nullptr;
    int rc = guarded([&] {
        auto owned = std::make_unique<RefModel>();
        owned->config = dorado::config::load_model_config(config_dir);
        auto state = read_b2w(weights);
        const auto opts = at::TensorOptions().dtype(at::kFloat).device(at::kCPU);
        if (owned->config.is_tx_model()) {
            auto model = dorado::basecall::model::TxModel(owned->config, opts);
            dorado::utils::load_state_dict(*model, state);
            model->eval();
            owned->module = torch::nn::ModuleHolder<torch::nn::AnyModule>(torch::nn::AnyModule(model));
        } else {
            auto model = dorado::basecall::model::CRFModel(owned->config);
            model->load_state_dict(state);
            model->eval();
            owned->module = torch::nn::ModuleHolder<torch::nn::AnyModule>(torch::nn::AnyModule(model));
        }
        m = owned.release();
    });
    return rc == 0 ? m : nullptr;
}

void ref_model_destroy(void* h) { delete static_cast<RefModel*>(h); }

// info[0..5] = stride, outsize, state_len, is_tx, clamp, num_features ; qinfo = qscale, qbias
int ref_model_info(void* h, int* info, float* qinfo) {
    auto* m = static_cast<RefModel*>(h);
    info[0] = m->config.stride;
    info[1] = m->config.outsize;
    info[2] = m->config.state_len;
    info[3] = m->config.is_tx_model() ? 1 : 0;
    info[4] = m->config.clamp ? 1 : 0;
    info[5] = m->config.num_features;
    qinfo[0] = m->config.qscale;
    qinfo[1] = m->config.qbias;
    return 0;
}

// signal [N,1,T] fp32 -> scores [N,T_out,C] fp32 ; what ModelRunner::call_chunks runs first
// (dorado/basecall/ModelRunner.cpp:35).
int ref_forward(void* h, const float* signal, int N, int T, float* scores, int* t_out, int* c_out) {
    auto* m = static_cast<RefModel*>(h);
    return guarded([&] {
        at::Tensor x = at::from_blob(const_cast<float*>(signal), {N, 1, T}, at::kFloat).clone();
        at::Tensor y = m->module->forward(x).contiguous();
        *t_out = int(y.size(1));
        *c_out = int(y.size(2));
        if (scores) {
            std::memcpy(scores, y.data_ptr<float>(), size_t(y.numel()) * 4);
        }
    });
}

// One chunk: scores [T,C] fp32 -> fwd, bwd, posts [T+1, C/4] fp32 using the reference's scans
// (CPUDecoder.cpp:43-92) and softmax (CPUDecoder.cpp:130).
int ref_scans(const float* scores, int T, int C, float blank, float* fwd, float* bwd, float* posts) {
    return guarded([&] {
        namespace inner = dorado::basecall::decode::inner;
        at::Tensor s = at::from_blob(const_cast<float*>(scores), {T, 1, C}, at::kFloat).clone();
        at::Tensor f = inner::forward_scores(s, blank);
        at::Tensor b = inner::backward_scores(s, blank);
        at::Tensor p = at::softmax(f + b, -1);
        const size_t bytes = size_t(T + 1) * size_t(C / 4) * 4;
        std::memcpy(fwd, f.contiguous().data_ptr<float>(), bytes);
        std::memcpy(bwd, b.contiguous().data_ptr<float>(), bytes);
        std::memcpy(posts, p.contiguous().data_ptr<float>(), bytes);
    });
}

// Reference beam search on caller-provided guides/posts (beam_search.cpp:522-606).
// scores: [T,C] fp32 (is_half=0) or IEEE half bits (is_half=1). seq/qstr/moves sized T.
int ref_beam_search_decode(const void* scores,
                           int is_half,
                           int T,
                           int C,
                           const float* bwd,
                           const float* posts,
                           int beam_width,
                           float beam_cut,
                           float blank,
                           float q_shift,
                           float q_scale,
                           char* seq,
                           char* qstr,
                           uint8_t* moves,
                           int* n_bases) {
    return guarded([&] {
        at::Tensor s = at::from_blob(const_cast<void*>(scores), {T, C}, is_half ? at::kHalf : at::kFloat);
        at::Tensor b = at::from_blob(const_cast<float*>(bwd), {T + 1, C / 4}, at::kFloat);
        at::Tensor p = at::from_blob(const_cast<float*>(posts), {T + 1, C / 4}, at::kFloat);
        auto [sequence, qstring, mv] = dorado::basecall::decode::beam_search_decode(
                s, b, p, size_t(beam_width), beam_cut, blank, q_shift, q_scale, 1.0f);
        *n_bases = int(sequence.size());
        std::memcpy(seq, sequence.data(), sequence.size());
        std::memcpy(qstr, qstring.data(), qstring.size());
        std::memcpy(moves, mv.data(), mv.size());
    });
}

// Whole reference decode on a batch: scores [N,T,C] fp32 -> per chunk seq/qstr/moves (row pitch T).
// Follows ModelRunner::call_chunks (ModelRunner.cpp:35-39): transpose to TNC then
// CPUDecoder::beam_search_part_2.
int ref_decode_chunks(const float* scores,
                      int N,
                      int T,
                      int C,
                      int beam_width,
                      float beam_cut,
                      float blank,
                      float q_shift,
                      float q_scale,
                      char* seq,
                      char* qstr,
                      uint8_t* moves,
                      int* n_bases) {
    return guarded([&] {
        using namespace dorado::basecall::decode;
        at::Tensor s = at::from_blob(const_cast<float*>(scores), {N, T, C}, at::kFloat);
        at::Tensor tnc = s.transpose(0, 1).contiguous();
        DecoderOptions opts;
        opts.beam_width = size_t(beam_width);
        opts.beam_cut = beam_cut;
        opts.blank_score = blank;
        opts.q_shift = q_shift;
        opts.q_scale = q_scale;
        CPUDecoder dec;
        auto res = dec.beam_search_part_2(dec.beam_search_part_1({tnc, N, opts}));
        for (int i = 0; i < N; ++i) {
            n_bases[i] = int(res[i].sequence.size());
            std::memcpy(seq + size_t(i) * T, res[i].sequence.data(), res[i].sequence.size());
            std::memcpy(qstr + size_t(i) * T, res[i].qstring.data(), res[i].qstring.size());
            std::memcpy(moves + size_t(i) * T, res[i].moves.data(), res[i].moves.size());
        }
    });
}

// utils::generate_chunks (read_pipeline/base/chunk.cpp:11-47).  Returns the number of offsets (written up to cap),
// or -1 when the reference throws.
long ref_generate_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap, uint64_t* out,
                         uint64_t cap) {
    long n = -1;
    guarded([&] {
        const auto v = dorado::utils::generate_chunks(num_samples, chunk_size, stride, overlap);
        for (size_t i = 0; i < v.size() && i < cap; ++i) {
            out[i] = v[i];
        }
        n = long(v.size());
    });
    return n;
}

// utils::generate_variable_chunks (read_pipeline/base/chunk.cpp:49-113): [first, second) intervals, flattened.
long ref_generate_variable_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap, uint64_t* out,
                                  uint64_t cap) {
    long n = -1;
    guarded([&] {
        const auto v = dorado::utils::generate_variable_chunks(num_samples, chunk_size, stride, overlap);
        for (size_t i = 0; i < v.size() && i < cap; ++i) {
            out[2 * i] = v[i].first;
            out[2 * i + 1] = v[i].second;
        }
        n = long(v.size());
    });
    return n;
}

// What the model input row of one chunk holds in the reference: ScalerNode scales the whole read in place
// (shift_scale_tensor_i16_to_f16_inplace, ScalerNode.cpp:226-229), BasecallerNode slices
// raw_data[offset : offset + chunk_size] (clamped at the read end) and repeat-pads a short slice with
// at::concat({slice.repeat({1, n}), slice[:overhang]}) (BasecallerNode.cpp:395-440).  The torch calls below are the
// same calls the node makes.
int ref_make_chunk_input(const int16_t* raw, uint64_t num_samples, uint64_t offset, uint64_t chunk_size, float shift,
                         float scale, uint16_t* out) {
    return guarded([&] {
        using at::indexing::Ellipsis;
        using at::indexing::Slice;
        at::Tensor t = at::from_blob(const_cast<int16_t*>(raw), {int64_t(num_samples)}, at::kShort).clone();
        dorado::utils::shift_scale_tensor_i16_to_f16_inplace(t, shift, scale);
        at::Tensor input_slice = t.index({Ellipsis, Slice(int64_t(offset), int64_t(offset + chunk_size))});
        if (input_slice.ndimension() == 1) {
            input_slice = input_slice.unsqueeze(0);
        }
        const size_t slice_size = input_slice.size(1);
        if (slice_size != chunk_size) {
            auto [n, overhang] = std::div((int)chunk_size, (int)slice_size);
            input_slice = at::concat({input_slice.repeat({1, n}), input_slice.index({Ellipsis, Slice(0, overhang)})}, 1);
        }
        input_slice = input_slice.contiguous();
        std::memcpy(out, input_slice.data_ptr(), chunk_size * 2);
    });
}

// utils::stitch_chunks (read_pipeline/base/stitch.cpp:12-96) on n chunks given as flat arrays.
// moves/seq/qstr are concatenations; *_len give the per-chunk lengths.  Outputs are sized by the caller
// (total moves / total bases are upper bounds).
int ref_stitch_chunks(uint64_t n,
                      const uint64_t* input_offsets,
                      const uint64_t* raw_chunk_sizes,
                      const uint8_t* moves,
                      const uint64_t* moves_len,
                      const char* seq,
                      const char* qstr,
                      const uint64_t* seq_len,
                      uint64_t raw_samples,
                      int stride,
                      uint8_t* moves_out,
                      char* seq_out,
                      char* qstr_out,
                      uint64_t* n_moves_out,
                      uint64_t* n_bases_out) {
    return guarded([&] {
        std::vector<std::unique_ptr<dorado::utils::Chunk>> owned;
        std::vector<const dorado::utils::Chunk*> ptrs;
        size_t mo = 0, so = 0;
        for (uint64_t i = 0; i < n; ++i) {
            auto c = std::make_unique<dorado::utils::Chunk>(input_offsets[i], raw_chunk_sizes[i]);
            c->moves.assign(moves + mo, moves + mo + moves_len[i]);
            c->seq.assign(seq + so, seq_len[i]);
            c->qstring.assign(qstr + so, seq_len[i]);
            mo += moves_len[i];
            so += seq_len[i];
            ptrs.push_back(c.get());
            owned.push_back(std::move(c));
        }
        dorado::ReadCommon rc;
        rc.attributes.model_stride = stride;
        rc.raw_data = at::empty({int64_t(raw_samples)}, at::kShort);
        dorado::utils::stitch_chunks(rc, ptrs);
        *n_moves_out = rc.moves.size();
        *n_bases_out = rc.seq.size();
        std::memcpy(moves_out, rc.moves.data(), rc.moves.size());
        std::memcpy(seq_out, rc.seq.data(), rc.seq.size());
        std::memcpy(qstr_out, rc.qstring.data(), rc.qstring.size());
    });
}

// ---- the reference's CPU runner itself (dorado/basecall/ModelRunner.cpp) ------------------------------------------
// ModelRunner loads its weights from <model dir>/*.tensor with torch::load (crf_utils.cpp:26-150,
// tensor_utils.cpp:154-163), so the B2W1 tensors are first written out in that format next to a copy of config.toml,
// under a directory that carries the model's own (non-deprecated) name.  An empty `weights` path reuses the directory a
// previous call wrote under the same work_dir (many runners, one copy of the files).
struct RefRunner {
    std::unique_ptr<dorado::basecall::ModelRunner> runner;
    std::filesystem::path dir;
};

void* ref_runner_create(const char* config_dir, const char* weights, const char* work_dir, int batch_size, int chunk_size) {
    RefRunner* out = nullptr;
    guarded([&] {
        namespace fs = std::filesystem;
        const fs::path src(config_dir);
        const fs::path dst = fs::path(work_dir) / src.filename();
        if (weights && weights[0]) {
            fs::create_directories(dst);
            fs::copy_file(src / "config.toml", dst / "config.toml", fs::copy_options::overwrite_existing);
            {
                std::ifstream f(weights, std::ios::binary);
                if (!f) throw std::runtime_error(std::string("cannot open weights file ") + weights);
                char magic[4];
                uint32_t n = 0;
                f.read(magic, 4);
                f.read(reinterpret_cast<char*>(&n), 4);
                for (uint32_t i = 0; i < n; ++i) {
                    uint32_t name_len = 0, ndim = 0;
                    f.read(reinterpret_cast<char*>(&name_len), 4);
                    std::string name(name_len, '\0');
                    f.read(name.data(), name_len);
                    f.read(reinterpret_cast<char*>(&ndim), 4);
                    std::vector<int64_t> dims(ndim);
                    int64_t numel = 1;
                    for (uint32_t d = 0; d < ndim; ++d) {
                        uint32_t v = 0;
                        f.read(reinterpret_cast<char*>(&v), 4);
                        dims[d] = v;
                        numel *= v;
                    }
                    at::Tensor t = at::empty(dims, at::kFloat);
                    f.read(reinterpret_cast<char*>(t.data_ptr<float>()), numel * 4);
                    if (!f) throw std::runtime_error("truncated weights file");
                    torch::save(std::vector<at::Tensor>{t}, (dst / name).string());
                }
            }
        }  // else: work_dir already holds the model directory written by an earlier runner
        auto config = dorado::config::load_model_config(dst);
        config.basecaller.set_batch_size(batch_size);
        config.basecaller.set_chunk_size(chunk_size);
        config.normalise_basecaller_params();
        auto owned = std::make_unique<RefRunner>();
        owned->dir = dst;
        owned->runner = std::make_unique<dorado::basecall::ModelRunner>(config, "cpu");
        out = owned.release();
    });
    return out;
}

void ref_runner_destroy(void* h) {
    auto* r = static_cast<RefRunner*>(h);
    delete r;  // the caller owns (and removes) work_dir
}

// dims[0..2] = batch_size, chunk_size, stride
int ref_runner_dims(void* h, int* dims) {
    auto* r = static_cast<RefRunner*>(h);
    dims[0] = int(r->runner->batch_size());
    dims[1] = int(r->runner->chunk_size());
    dims[2] = r->runner->config().stride;
    return 0;
}

// ModelRunner::accept_chunk (ModelRunner.cpp:47-49): chunk [1, chunk_size] fp32 (the CPU decoder's dtype)
int ref_runner_accept_chunk(void* h, int idx, const float* samples, int len) {
    auto* r = static_cast<RefRunner*>(h);
    return guarded([&] {
        at::Tensor c = at::from_blob(const_cast<float*>(samples), {1, len}, at::kFloat);
        r->runner->accept_chunk(idx, c);
    });
}

// ModelRunner::call_chunks (ModelRunner.cpp:32-45); outputs have row pitch T_out = chunk_size / stride (may be NULL)
int ref_runner_call_chunks(void* h, int num_chunks, char* seq, char* qstr, uint8_t* moves, int* n_bases) {
    auto* r = static_cast<RefRunner*>(h);
    return guarded([&] {
        auto res = r->runner->call_chunks(num_chunks);
        const size_t T = r->runner->chunk_size() / size_t(r->runner->config().stride);
        for (size_t i = 0; i < res.size(); ++i) {
            if (n_bases) n_bases[i] = int(res[i].sequence.size());
            if (seq) std::memcpy(seq + i * T, res[i].sequence.data(), res[i].sequence.size());
            if (qstr) std::memcpy(qstr + i * T, res[i].qstring.data(), res[i].qstring.size());
            if (moves) std::memcpy(moves + i * T, res[i].moves.data(), res[i].moves.size());
        }
    });
}

// ModelRunner::sample_stats (ModelRunner.cpp:51-57): batches_called, model_ms, decode_ms
int ref_runner_stats(void* h, double* out3) {
    auto* r = static_cast<RefRunner*>(h);
    return guarded([&] {
        const auto st = r->runner->sample_stats();
        out3[0] = st.at("batches_called");
        out3[1] = st.at("model_ms");
        out3[2] = st.at("decode_ms");
    });
}

}  // extern "C"
