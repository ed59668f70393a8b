// Front end of the hot path: chunk offsets, raw int16 -> scaled fp16 batch input on the
// device, stitching of called chunks.  The integer/byte host logic mirrors what the reference does on the host; the
// scaling + slicing + repeat-padding that the reference spreads over ScalerNode (an AVX pass over every read) and
// BasecallerNode (slice, at::concat, index_put_ per chunk) is one kernel here.
#include "frontend.h"

#include "common.cuh"

#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <string>

namespace b200 {

// ---- utils::generate_chunks (dorado/read_pipeline/base/chunk.cpp:11-47) -------------------------------------------
// First chunk at 0, then steps of (chunk_size - overlap); the last chunk is pulled back so that it ends at the read
// end, rounded UP to the next stride boundary (its tail is then short and gets repeat-padded).
uint64_t generate_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap,
                         uint64_t* offsets, uint64_t capacity) {
    if (num_samples == 0) throw std::invalid_argument("generate_chunks: empty read");
    if (stride == 0) throw std::invalid_argument("generate_chunks: invalid stride 0");
    if (chunk_size == 0 || chunk_size % stride != 0 || chunk_size <= overlap) {
        throw std::invalid_argument("generate_chunks: invalid chunk size " + std::to_string(chunk_size) + " with overlap " +
                                    std::to_string(overlap) + " and stride " + std::to_string(stride));
    }
    if (overlap % stride != 0) {
        throw std::invalid_argument("generate_chunks: invalid overlap " + std::to_string(overlap) + " with stride " +
                                    std::to_string(stride));
    }
    uint64_t last = num_samples > chunk_size ? num_samples - chunk_size : 0;
    last = (last + stride - 1) / stride * stride;
    const uint64_t step = chunk_size - overlap;
    uint64_t count = 0, off = 0;
    for (;;) {
        if (count < capacity && offsets) offsets[count] = off;
        ++count;
        if (off + chunk_size >= num_samples) break;
        off = std::min(off + step, last);
    }
    return count;
}

// ---- utils::generate_variable_chunks (dorado/read_pipeline/base/chunk.cpp:49-113) --------------------------------
// As few chunks as the fixed chunking would need, but of near-equal length: the read plus the (count - 1) overlaps is
// split evenly (the first `total % count` chunks get one extra sample), then every interior start is rounded up and
// every interior end rounded down to a stride multiple.
uint64_t generate_variable_chunks(uint64_t num_samples, uint64_t chunk_size, uint64_t stride, uint64_t overlap,
                                  uint64_t* intervals, uint64_t capacity) {
    if (num_samples == 0) throw std::invalid_argument("generate_variable_chunks: empty read");
    if (stride == 0) throw std::invalid_argument("generate_variable_chunks: invalid stride 0");
    if (chunk_size == 0 || chunk_size % stride != 0 || chunk_size == stride || chunk_size <= overlap) {
        throw std::invalid_argument("generate_variable_chunks: invalid chunk size " + std::to_string(chunk_size) +
                                    " with overlap " + std::to_string(overlap) + " and stride " + std::to_string(stride));
    }
    if (overlap % stride != 0 || (stride != 1 && overlap == 0)) {
        throw std::invalid_argument("generate_variable_chunks: invalid overlap " + std::to_string(overlap) + " with stride " +
                                    std::to_string(stride));
    }
    const uint64_t step = chunk_size - overlap;
    const uint64_t count = 1 + (num_samples > chunk_size ? (num_samples - chunk_size + step - 1) / step : 0);
    const uint64_t total = num_samples + (count - 1) * overlap;
    const uint64_t base = total / count, longer = total % count;
    uint64_t start = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t end = start + base + (i < longer ? 1 : 0);
        uint64_t first = start, second = end;
        if (i > 0) first = (first + stride - 1) / stride * stride;
        if (i + 1 < count) second -= second % stride;
        if (i < capacity && intervals) {
            intervals[2 * i] = first;
            intervals[2 * i + 1] = second;
        }
        start = end - overlap;
    }
    return count;
}

// ---- utils::stitch_chunks (dorado/read_pipeline/base/stitch.cpp:12-96) --------------------------------------------
// Each overlap (in stride units) is cut at its midpoint: the earlier chunk loses the last floor(ovl / 2) blocks, the
// later one the first ovl - floor(ovl / 2).  Bases follow the moves: a chunk contributes the bases whose move lies in
// the kept block window.
void stitch_chunks(const b200_called_chunk* c, uint64_t n, uint64_t raw_samples, int stride, uint8_t* moves_out,
                   char* seq_out, char* qstr_out, uint64_t* n_moves_out, uint64_t* n_bases_out) {
    if (!c || n == 0) throw std::invalid_argument("stitch_chunks: no chunks");
    if (stride <= 0) throw std::invalid_argument("stitch_chunks: invalid stride");
    uint64_t nm = 0, nb = 0;
    uint64_t front = 0;  // blocks dropped at the front of the current chunk
    const uint64_t keep_blocks = raw_samples / (uint64_t)stride;
    for (uint64_t i = 0; i < n; ++i) {
        const b200_called_chunk& ch = c[i];
        if ((!ch.moves && ch.n_moves) || ((!ch.sequence || !ch.qstring) && ch.n_bases)) {
            throw std::invalid_argument("stitch_chunks: null chunk buffer");
        }
        const bool last = i + 1 == n;
        uint64_t rear = 0, next_front = 0;
        if (!last) {
            const int64_t ovl = (int64_t)(ch.raw_chunk_size + ch.input_offset) - (int64_t)c[i + 1].input_offset;
            if (ovl < 0 || ovl % stride != 0) {
                throw std::invalid_argument("stitch_chunks: chunks do not overlap on a stride boundary");
            }
            rear = (uint64_t)(ovl / stride) / 2;
            next_front = (uint64_t)(ovl / stride) - rear;
        }
        if (front + rear > ch.n_moves) throw std::invalid_argument("stitch_chunks: overlap longer than the chunk");
        uint64_t m_end = ch.n_moves - rear;
        if (n == 1) m_end = std::min(m_end, keep_blocks);  // a read shorter than one chunk keeps floor(samples / stride) blocks
        const uint64_t b_begin = std::accumulate(ch.moves, ch.moves + front, (uint64_t)0);
        uint64_t b_end;
        if (n == 1) {
            b_end = std::accumulate(ch.moves, ch.moves + m_end, (uint64_t)0);
        } else if (last) {
            b_end = ch.n_bases;  // every remaining base (the reference's substr(start_pos))
        } else {
            const uint64_t trimmed = std::accumulate(ch.moves + m_end, ch.moves + ch.n_moves, (uint64_t)0);
            if (trimmed > ch.n_bases) throw std::invalid_argument("stitch_chunks: moves and sequence length disagree");
            b_end = ch.n_bases - trimmed;
        }
        if (b_begin > b_end || b_end > ch.n_bases) throw std::invalid_argument("stitch_chunks: moves and sequence length disagree");
        std::copy(ch.moves + front, ch.moves + m_end, moves_out + nm);
        nm += m_end - front;
        std::copy(ch.sequence + b_begin, ch.sequence + b_end, seq_out + nb);
        std::copy(ch.qstring + b_begin, ch.qstring + b_end, qstr_out + nb);
        nb += b_end - b_begin;
        front = next_front;
    }
    // partial-stride overhang: at most one block (and its base) beyond floor(raw_samples / stride)
    if (nm > keep_blocks) {
        if (moves_out[nm - 1] == 1 && nb > 0) --nb;
        --nm;
    }
    *n_moves_out = nm;
    *n_bases_out = nb;
}

// ---- raw int16 slices -> scaled, repeat-padded fp16 batch input ---------------------------------------------------
// One block row per chunk slot; slots whose slice_len is 0 were supplied as fp16 and are left alone.  IEEE fp32
// subtract and divide, then round-to-nearest-even to fp16: bit-identical to the reference's AVX2 / scalar pass.
__global__ void __launch_bounds__(256) raw_chunk_gather_kernel(const int16_t* __restrict__ staged,
                                                               const RawSlot* __restrict__ slots,
                                                               __half* __restrict__ input, int T_in) {
    const int n = blockIdx.y;
    const RawSlot s = slots[n];
    if (s.slice_len <= 0) return;
    const int16_t* src = staged + (size_t)n * T_in;
    __half* dst = input + (size_t)n * T_in;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T_in; j += gridDim.x * blockDim.x) {
        const int k = s.slice_len == T_in ? j : j % s.slice_len;
        const float v = __fdiv_rn(__fsub_rn((float)src[k], s.shift), s.scale);
        dst[j] = __float2half_rn(v);
    }
}

void launch_raw_chunk_gather(const int16_t* staged, const RawSlot* slots, __half* input, int num_chunks, int T_in,
                             cudaStream_t stream) {
    if (num_chunks <= 0) return;
    const int bx = std::max(1, std::min(8, (T_in + 255) / 256));
    raw_chunk_gather_kernel<<<dim3(bx, num_chunks, 1), 256, 0, stream>>>(staged, slots, input, T_in);
    B200_CUDA(cudaGetLastError());
}

}  // namespace b200
