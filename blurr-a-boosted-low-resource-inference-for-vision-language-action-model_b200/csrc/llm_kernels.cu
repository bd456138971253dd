// Small kernels of the Llama-shaped decoder (OpenVLA-7B-shaped path): RoPE + KV-cache append for
// multi-head attention, token embedding gather, last-row gather and greedy argmax.  All bandwidth-trivial (KB per call);
// written for few dependent launches: one CTA per token row, vectorised 8-byte accesses, warp-shuffle reductions.
#include "bodies.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace blurr {

// HF Llama (modeling_llama.py apply_rotary_pos_emb / rotate_half): with d = head_dim, for j < d/2
//   y[j]       = bf16(bf16(x[j] * cos[j])       + bf16(-x[j + d/2] * sin[j]))
//   y[j + d/2] = bf16(bf16(x[j + d/2] * cos[j]) + bf16( x[j]       * sin[j]))
// (cos/sin are cat(freqs, freqs), cast to the activation dtype).  One thread per 4 rotation pairs.
__global__ void __launch_bounds__(256) rope_mha_kernel(const RopeMhaArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const int seq = t / a.tokens_per_seq, i = t - seq * a.tokens_per_seq;
    int pos = a.pos0 + i;
    if (pos >= a.n_pos) pos = a.n_pos - 1;               // the host validates the range
    const int half = a.head_dim >> 1;
    const int gpr = half >> 2;                            // 4-wide groups per rotated head
    const int n_rot_heads = a.n_heads + a.n_kv_heads;
    const int n_rot = n_rot_heads * gpr;
    const int n_v = (a.n_kv_heads * a.head_dim) >> 2;
    const size_t sstride = static_cast<size_t>(a.T) * a.ldp;
    const size_t cache_row = (static_cast<size_t>(seq) * a.n_slots + pos) * (a.n_kv_heads * a.head_dim);
    for (int it = threadIdx.x; it < n_rot + n_v; it += blockDim.x) {
        if (it < n_rot) {
            const int h = it / gpr, j = (it - h * gpr) << 2;
            const int col = h * a.head_dim + j;           // q heads then k heads are contiguous in the projection
            float4 x1, x2;
            if (a.lin != nullptr) {
                x1 = load_bf16x4(a.lin + static_cast<size_t>(t) * a.ldl + col);
                x2 = load_bf16x4(a.lin + static_cast<size_t>(t) * a.ldl + col + half);
            } else {
                x1 = sum_slices(a.partial + static_cast<size_t>(t) * a.ldp + col, sstride, a.splitk);
                x2 = sum_slices(a.partial + static_cast<size_t>(t) * a.ldp + col + half, sstride, a.splitk);
            }
            const float4 cs = *reinterpret_cast<const float4*>(a.cos_table + static_cast<size_t>(pos) * half + j);
            const float4 sn = *reinterpret_cast<const float4*>(a.sin_table + static_cast<size_t>(pos) * half + j);
            const float u1[4] = {bf16_round(x1.x), bf16_round(x1.y), bf16_round(x1.z), bf16_round(x1.w)};
            const float u2[4] = {bf16_round(x2.x), bf16_round(x2.y), bf16_round(x2.z), bf16_round(x2.w)};
            const float c[4] = {cs.x, cs.y, cs.z, cs.w}, s[4] = {sn.x, sn.y, sn.z, sn.w};
            float y1[4], y2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                y1[e] = bf16_round(bf16_round(u1[e] * c[e]) + bf16_round(-u2[e] * s[e]));
                y2[e] = bf16_round(bf16_round(u2[e] * c[e]) + bf16_round(u1[e] * s[e]));
            }
            bf16* dst = (h < a.n_heads) ? a.q_out + static_cast<size_t>(t) * (a.n_heads * a.head_dim) + h * a.head_dim
                                        : a.k_cache + cache_row + (h - a.n_heads) * a.head_dim;
            store_bf16x4(dst + j, make_float4(y1[0], y1[1], y1[2], y1[3]));
            store_bf16x4(dst + half + j, make_float4(y2[0], y2[1], y2[2], y2[3]));
        } else {
            const int j = (it - n_rot) << 2;
            const int col = n_rot_heads * a.head_dim + j;
            const float4 v = a.lin != nullptr ? load_bf16x4(a.lin + static_cast<size_t>(t) * a.ldl + col)
                                              : sum_slices(a.partial + static_cast<size_t>(t) * a.ldp + col, sstride, a.splitk);
            store_bf16x4(a.v_cache + cache_row + j,
                         make_float4(bf16_round(v.x), bf16_round(v.y), bf16_round(v.z), bf16_round(v.w)));
        }
    }
}

cudaError_t launch_rope_mha(cudaStream_t stream, const RopeMhaArgs& a) {
    if ((a.head_dim & 7) || a.T <= 0) return cudaErrorInvalidValue;
    return launch_kernel(rope_mha_kernel, dim3(a.T), dim3(256), 0, stream, a);
}

__global__ void __launch_bounds__(256) embed_rows_kernel(const int64_t* ids, int n, const bf16* table, long long vocab,
                                                         int width, bf16* out, int* err_flag) {
    pdl_wait();
    pdl_trigger();
    const int r = blockIdx.x;
    long long id = ids[r];
    if (id < 0 || id >= vocab) { if (threadIdx.x == 0 && err_flag) atomicExch(err_flag, 1); id = 0; }
    const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * width);
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * width);
    for (int i = threadIdx.x; i < width / 8; i += blockDim.x) dst[i] = src[i];
}

cudaError_t launch_embed_rows(cudaStream_t stream, const int64_t* ids, int n, const bf16* table, long long vocab, int width,
                              bf16* out, int* err_flag) {
    if (width & 7) return cudaErrorInvalidValue;
    return launch_kernel(embed_rows_kernel, dim3(n), dim3(256), 0, stream, ids, n, table, vocab, width, out, err_flag);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const bf16* src, int rows_per_seq, int row, int width, bf16* dst) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x;
    const uint4* s = reinterpret_cast<const uint4*>(src + (static_cast<size_t>(b) * rows_per_seq + row) * width);
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<size_t>(b) * width);
    for (int i = threadIdx.x; i < width / 8; i += blockDim.x) d[i] = s[i];
}

cudaError_t launch_gather_rows(cudaStream_t stream, const bf16* src, int batch, int rows_per_seq, int row, int width, bf16* dst) {
    if (width & 7) return cudaErrorInvalidValue;
    return launch_kernel(gather_rows_kernel, dim3(batch), dim3(256), 0, stream, src, rows_per_seq, row, width, dst);
}

// One CTA per sequence.  Greedy decoding compares the bf16 logits as fp32 (HF casts the last row to fp32 first); the
// lowest index wins among equal maxima.
__global__ void __launch_bounds__(1024) argmax_rows_kernel(const bf16* logits, int ld, int vocab, int64_t* ids, int64_t* ids_copy,
                                                           int copy_stride) {
    pdl_wait();
    pdl_trigger();
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    const int b = blockIdx.x;
    const bf16* row = logits + static_cast<size_t>(b) * ld;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = bf2f(row[i]);
        if (v != v) continue;                                   // NaN never wins
        if (v > best || (v == best && i < best_i)) { best = v; best_i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        best = lane < nw ? s_val[lane] : -INFINITY;
        best_i = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) {
            const int64_t id = best_i == 0x7fffffff ? 0 : best_i;
            ids[b] = id;
            if (ids_copy != nullptr) ids_copy[static_cast<size_t>(b) * copy_stride] = id;
        }
    }
}

cudaError_t launch_argmax_rows(cudaStream_t stream, const bf16* logits, int batch, int ld, int vocab, int64_t* ids,
                               int64_t* ids_copy, int copy_stride) {
    return launch_kernel(argmax_rows_kernel, dim3(batch), dim3(1024), 0, stream, logits, ld, vocab, ids, ids_copy, copy_stride);
}

// GLU over split-K partials of a few-token gate/up projection whose weight rows alternate gate_j, up_j (the layout of
// the EPI_GEGLU epilogue): out[t][j] = bf16(bf16(act(bf16(gate_j))) * bf16(up_j)).  8 partial columns -> 4 outputs per thread.
__global__ void __launch_bounds__(256) glu_partial_kernel(const float* partial, int splitk, int T, int Nw, int act, bf16* out,
                                                          int ldo) {
    pdl_wait();
    pdl_trigger();
    const int per_row = Nw >> 3;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= T * per_row) return;
    const int t = idx / per_row, c = (idx - t * per_row) << 3;
    const size_t sstride = static_cast<size_t>(T) * Nw;
    const float4 a = sum_slices(partial + static_cast<size_t>(t) * Nw + c, sstride, splitk);
    const float4 b = sum_slices(partial + static_cast<size_t>(t) * Nw + c + 4, sstride, splitk);
    const float g[4] = {bf16_round(a.x), bf16_round(a.z), bf16_round(b.x), bf16_round(b.z)};
    const float u[4] = {bf16_round(a.y), bf16_round(a.w), bf16_round(b.y), bf16_round(b.w)};
    store_bf16x4(out + static_cast<size_t>(t) * ldo + (c >> 1),
                 make_float4(bf16_round(glu_act_f32(g[0], act)) * u[0], bf16_round(glu_act_f32(g[1], act)) * u[1],
                             bf16_round(glu_act_f32(g[2], act)) * u[2], bf16_round(glu_act_f32(g[3], act)) * u[3]));
}

cudaError_t launch_glu_partial(cudaStream_t stream, const float* partial, int splitk, int T, int Nw, int act, bf16* out, int ldo) {
    if (Nw & 7) return cudaErrorInvalidValue;
    const int total = T * (Nw >> 3);
    return launch_kernel(glu_partial_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, partial, splitk, T, Nw, act, out, ldo);
}

}  // namespace blurr
