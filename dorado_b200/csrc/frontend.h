// Front end of the hot path: see frontend.cu.
#pragma once

#include "b200call.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// per batch slot: how the gather/scale kernel fills the slot's input row
struct RawSlot {
    int32_t slice_len;  // un-padded samples staged for this slot; 0 = slot was given as fp16
    float shift, scale;
};

uint64_t generate_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap,
                         uint64_t* offsets, uint64_t capacity);
uint64_t generate_variable_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap,
                                  uint64_t* intervals, uint64_t capacity);
void stitch_chunks(const b200_called_chunk* chunks, uint64_t n_chunks, uint64_t raw_samples, int stride,
                   uint8_t* moves_out, char* seq_out, char* qstr_out, uint64_t* n_moves_out, uint64_t* n_bases_out);
void launch_raw_chunk_gather(const int16_t* staged, const RawSlot* slots, __half* input, int num_chunks, int T_in,
                             cudaStream_t stream);

}  // namespace b200
