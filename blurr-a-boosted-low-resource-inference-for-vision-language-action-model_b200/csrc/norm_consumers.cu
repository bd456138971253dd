// GEMM consumers: deterministic split-K reduction + bias + residual/position add + norm, and
// RoPE + KV-cache append.  HBM/L2-bandwidth-bound row kernels: one CTA per token row,
// coalesced loads, block reduction through warp shuffles + shared memory.
//
// Rounding points follow the reference exactly:
//   linear output  -> bf16(acc + bias)                              (nn.Linear)
//   residual       -> bf16(res + x)                                 (siglip.py:228,236; joint_model.py:79-81,126-128)
//   GemmaRMSNorm   -> bf16((x * rsqrt(mean(x^2) + eps)) * (1 + w))  (paligemma/modules.py:13-21)
//   LayerNorm      -> bf16((x - mean) * rstd * w + b)               (nn.LayerNorm, eps 1e-6)
//   RoPE           -> bf16(bf16(x*cos) + bf16(rot(x)*sin))          (utils.py:11-16), K cached post-RoPE
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace blurr {

static constexpr int kRowThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();               // protect `red` from the previous use
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (kRowThreads / 32)) ? red[l] : 0.f;
    t = warp_sum(t);
    return t;                      // every thread holds the total
}

// sum of the split-K slices of 4 consecutive columns, 4 independent 16-byte loads in flight
__device__ __forceinline__ float4 sum_slices(const float* __restrict__ base, size_t slice_stride, int splitk) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 4 <= splitk; z += 4) {
        const float4 p0 = *reinterpret_cast<const float4*>(base + (z + 0) * slice_stride);
        const float4 p1 = *reinterpret_cast<const float4*>(base + (z + 1) * slice_stride);
        const float4 p2 = *reinterpret_cast<const float4*>(base + (z + 2) * slice_stride);
        const float4 p3 = *reinterpret_cast<const float4*>(base + (z + 3) * slice_stride);
        acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;     // fixed order z = 0, 1, 2, ...
        acc.x += p1.x; acc.y += p1.y; acc.z += p1.z; acc.w += p1.w;
        acc.x += p2.x; acc.y += p2.y; acc.z += p2.z; acc.w += p2.w;
        acc.x += p3.x; acc.y += p3.y; acc.z += p3.z; acc.w += p3.w;
    }
    for (; z < splitk; ++z) {
        const float4 p = *reinterpret_cast<const float4*>(base + z * slice_stride);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    return acc;
}

__device__ __forceinline__ float4 load_bf16x4(const bf16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_bf16x4(bf16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

// One CTA per token row; each thread owns VPT groups of 4 consecutive columns in registers.
template <int VPT>
__global__ void __launch_bounds__(kRowThreads) consumer_kernel(const ConsumerArgs a) {
    __shared__ float red[kRowThreads / 32];
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const int nvec = a.N >> 2;
    float4 x[VPT];
    float lsum = 0.f, lsq = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = threadIdx.x + i * kRowThreads;
        x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v >= nvec) continue;
        const int n = v << 2;
        float4 val;
        if (a.partial != nullptr) {
            float4 acc = sum_slices(a.partial + static_cast<size_t>(t) * a.ldp + n,
                                    static_cast<size_t>(a.T) * a.ldp, a.splitk);
            if (a.bias != nullptr) {
                const float4 b = load_bf16x4(a.bias + n);
                acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            }
            val = make_float4(bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w));
            if (a.out_scale != 1.0f)
                val = make_float4(bf16_round(val.x * a.out_scale), bf16_round(val.y * a.out_scale),
                                  bf16_round(val.z * a.out_scale), bf16_round(val.w * a.out_scale));
            if (a.add_mode == ADD_RESIDUAL) {
                const float4 r = load_bf16x4(a.res + static_cast<size_t>(t) * a.ldr + n);
                val = make_float4(bf16_round(r.x + val.x), bf16_round(r.y + val.y), bf16_round(r.z + val.z),
                                  bf16_round(r.w + val.w));
            } else if (a.add_mode == ADD_POSEMB) {
                const float4 r = load_bf16x4(a.pos + static_cast<size_t>(t % a.pos_rows) * a.N + n);
                val = make_float4(bf16_round(val.x + r.x), bf16_round(val.y + r.y), bf16_round(val.z + r.z),
                                  bf16_round(val.w + r.w));
            }
        } else {
            val = load_bf16x4(a.res + static_cast<size_t>(t) * a.ldr + n);
        }
        if (a.x_out != nullptr) store_bf16x4(a.x_out + static_cast<size_t>(t) * a.ldx + n, val);
        x[i] = val;
        lsum += (val.x + val.y) + (val.z + val.w);
        lsq += (val.x * val.x + val.y * val.y) + (val.z * val.z + val.w * val.w);
    }
    if (a.norm_mode == NORM_NONE || a.xn_out == nullptr) return;

    if (a.norm_mode == NORM_RMS_GEMMA) {
        const float ms = block_sum(lsq, red) / static_cast<float>(a.N);
        const float r = rsqrtf(ms + a.eps);
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = threadIdx.x + i * kRowThreads;
            if (v >= nvec) continue;
            const int n = v << 2;
            const float4 w = load_bf16x4(a.norm_w + n);
            const float4 y = make_float4((x[i].x * r) * (1.0f + w.x), (x[i].y * r) * (1.0f + w.y),
                                         (x[i].z * r) * (1.0f + w.z), (x[i].w * r) * (1.0f + w.w));
            store_bf16x4(a.xn_out + static_cast<size_t>(t) * a.ldn + n, y);
        }
    } else {
        const float mean = block_sum(lsum, red) / static_cast<float>(a.N);
        float lvar = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = threadIdx.x + i * kRowThreads;
            if (v >= nvec) continue;
            const float dx = x[i].x - mean, dy = x[i].y - mean, dz = x[i].z - mean, dw = x[i].w - mean;
            lvar += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        const float var = block_sum(lvar, red) / static_cast<float>(a.N);
        const float rstd = rsqrtf(var + a.eps);
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = threadIdx.x + i * kRowThreads;
            if (v >= nvec) continue;
            const int n = v << 2;
            const float4 w = load_bf16x4(a.norm_w + n), b = load_bf16x4(a.norm_b + n);
            const float4 y = make_float4((x[i].x - mean) * rstd * w.x + b.x, (x[i].y - mean) * rstd * w.y + b.y,
                                         (x[i].z - mean) * rstd * w.z + b.z, (x[i].w - mean) * rstd * w.w + b.w);
            store_bf16x4(a.xn_out + static_cast<size_t>(t) * a.ldn + n, y);
        }
    }
}

cudaError_t launch_consumer(cudaStream_t stream, const ConsumerArgs& a) {
    if ((a.N & 3) || a.N > 2 * 4 * kRowThreads || (a.ldp & 3)) return cudaErrorInvalidValue;
    if (a.N <= 4 * kRowThreads)
        return launch_kernel(consumer_kernel<1>, dim3(a.T), dim3(kRowThreads), 0, stream, a);
    return launch_kernel(consumer_kernel<2>, dim3(a.T), dim3(kRowThreads), 0, stream, a);
}

__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ partial, int splitk, int T,
                                                       int N, int ldp, const bf16* __restrict__ bias, int act,
                                                       float scale, bf16* __restrict__ out, int ldo) {
    pdl_wait();
    pdl_trigger();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nvec = N >> 2;
    if (idx >= T * nvec) return;
    const int t = idx / nvec, n = (idx - t * nvec) << 2;
    float4 acc = sum_slices(partial + static_cast<size_t>(t) * ldp + n, static_cast<size_t>(T) * ldp, splitk);
    if (bias != nullptr) {
        const float4 b = load_bf16x4(bias + n);
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    float v[4] = {bf16_round(acc.x), bf16_round(acc.y), bf16_round(acc.z), bf16_round(acc.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (act == ACT_SILU) v[i] = bf16_round(silu_f32(v[i]));
        if (scale != 1.0f) v[i] = bf16_round(v[i] * scale);
    }
    store_bf16x4(out + static_cast<size_t>(t) * ldo + n, make_float4(v[0], v[1], v[2], v[3]));
}

cudaError_t launch_bias_act(cudaStream_t stream, const float* partial, int splitk, int T, int N, int ldp,
                            const bf16* bias, int act, float scale, bf16* out, int ldo) {
    if ((N & 3) || (ldp & 3)) return cudaErrorInvalidValue;
    const int total = T * (N >> 2);
    return launch_kernel(bias_act_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, partial, splitk, T, N,
                         ldp, bias, act, scale, out, ldo);
}

// One CTA per token.  Work items: for every rotated head (queries + the key head) 32 pairs of
// float4 column groups (dims [4j,4j+4) and [128+4j,128+4j+4): the rotate_half partners), plus 64
// plain float4 groups of the value head.
__global__ void __launch_bounds__(256) rope_kv_kernel(const RopeKvArgs a) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.x;
    const int b = t / a.tokens_per_sample, i = t - b * a.tokens_per_sample;
    long long pos = a.position_ids[static_cast<size_t>(b) * a.tokens_per_sample + i];
    if (pos < 0) pos = 0;
    if (pos >= a.n_pos) pos = a.n_pos - 1;   // host validates the range; never read out of bounds
    const int slot = a.slot_base + i;
    const size_t cache_row = (static_cast<size_t>(b) * a.n_slots + slot) * 256;
    const float* prow = a.partial + static_cast<size_t>(t) * a.ldp;
    const size_t sstride = static_cast<size_t>(a.T) * a.ldp;
    const int n_rot = (a.n_heads + 1) * 32;
    const int n_items = n_rot + 64;
    for (int it = threadIdx.x; it < n_items; it += 256) {
        if (it < n_rot) {
            const int h = it >> 5, j = (it & 31) << 2;             // head, first dim of the group
            if (h < a.n_heads && a.q_out == nullptr) continue;
            float4 x1 = sum_slices(prow + h * 256 + j, sstride, a.splitk);
            float4 x2 = sum_slices(prow + h * 256 + 128 + j, sstride, a.splitk);
            const float4 cs = *reinterpret_cast<const float4*>(a.cos_table + pos * 128 + j);
            const float4 sn = *reinterpret_cast<const float4*>(a.sin_table + pos * 128 + j);
            float u1[4] = {bf16_round(x1.x), bf16_round(x1.y), bf16_round(x1.z), bf16_round(x1.w)};
            float u2[4] = {bf16_round(x2.x), bf16_round(x2.y), bf16_round(x2.z), bf16_round(x2.w)};
            const float c[4] = {cs.x, cs.y, cs.z, cs.w}, s[4] = {sn.x, sn.y, sn.z, sn.w};
            float y1[4], y2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // x*cos + rotate_half(x)*sin, every op rounded to bf16 (utils.py:11-16)
                y1[e] = bf16_round(bf16_round(u1[e] * c[e]) + bf16_round(-u2[e] * s[e]));
                y2[e] = bf16_round(bf16_round(u2[e] * c[e]) + bf16_round(u1[e] * s[e]));
            }
            bf16* dst = (h < a.n_heads) ? a.q_out + static_cast<size_t>(t) * (a.n_heads * 256) + h * 256
                                        : a.k_cache + cache_row;
            store_bf16x4(dst + j, make_float4(y1[0], y1[1], y1[2], y1[3]));
            store_bf16x4(dst + 128 + j, make_float4(y2[0], y2[1], y2[2], y2[3]));
        } else {
            const int j = (it - n_rot) << 2;
            const float4 v = sum_slices(prow + (a.n_heads + 1) * 256 + j, sstride, a.splitk);
            store_bf16x4(a.v_cache + cache_row + j,
                         make_float4(bf16_round(v.x), bf16_round(v.y), bf16_round(v.z), bf16_round(v.w)));
        }
    }
}

cudaError_t launch_rope_kv(cudaStream_t stream, const RopeKvArgs& a) {
    if (a.ldp & 3) return cudaErrorInvalidValue;
    return launch_kernel(rope_kv_kernel, dim3(a.T), dim3(256), 0, stream, a);
}

__global__ void rope_table_kernel(const float* __restrict__ inv_freq, int n_pos, float* __restrict__ cos_t,
                                  float* __restrict__ sin_t) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pos * 128) return;
    const int pos = idx >> 7, j = idx & 127;
    // GemmaRotaryEmbedding.forward (paligemma/modules.py:47-67): fp32 angle, cos/sin -> model dtype
    const float angle = __fmul_rn(inv_freq[j], static_cast<float>(pos));
    cos_t[idx] = bf16_round(cosf(angle));
    sin_t[idx] = bf16_round(sinf(angle));
}

cudaError_t launch_rope_table(cudaStream_t stream, const float* inv_freq, int n_pos, float* cos_t,
                              float* sin_t) {
    const int total = n_pos * 128;
    rope_table_kernel<<<(total + 255) / 256, 256, 0, stream>>>(inv_freq, n_pos, cos_t, sin_t);
    return cudaGetLastError();
}

}  // namespace blurr
