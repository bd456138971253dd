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

__global__ void __launch_bounds__(kRowThreads) consumer_kernel(const ConsumerArgs a) {
    extern __shared__ float row[];          // [N]
    __shared__ float red[kRowThreads / 32];
    const int t = blockIdx.x;
    if (t >= a.T) return;

    float lsum = 0.f, lsq = 0.f;
    for (int n = threadIdx.x; n < a.N; n += kRowThreads) {
        float x;
        if (a.partial != nullptr) {
            float acc = 0.f;
            for (int z = 0; z < a.splitk; ++z)
                acc += a.partial[(static_cast<size_t>(z) * a.T + t) * a.ldp + n];
            if (a.bias != nullptr) acc += bf2f(a.bias[n]);
            x = bf16_round(acc);
            if (a.out_scale != 1.0f) x = bf16_round(x * a.out_scale);
            if (a.add_mode == ADD_RESIDUAL)
                x = bf16_round(bf2f(a.res[static_cast<size_t>(t) * a.ldr + n]) + x);
            else if (a.add_mode == ADD_POSEMB)
                x = bf16_round(x + bf2f(a.pos[static_cast<size_t>(t % a.pos_rows) * a.N + n]));
        } else {
            x = bf2f(a.res[static_cast<size_t>(t) * a.ldr + n]);
        }
        if (a.x_out != nullptr) a.x_out[static_cast<size_t>(t) * a.ldx + n] = f2bf(x);
        row[n] = x;
        lsum += x;
        lsq += x * x;
    }
    if (a.norm_mode == NORM_NONE || a.xn_out == nullptr) return;

    if (a.norm_mode == NORM_RMS_GEMMA) {
        const float ms = block_sum(lsq, red) / static_cast<float>(a.N);
        const float r = rsqrtf(ms + a.eps);
        for (int n = threadIdx.x; n < a.N; n += kRowThreads) {
            const float y = (row[n] * r) * (1.0f + bf2f(a.norm_w[n]));
            a.xn_out[static_cast<size_t>(t) * a.ldn + n] = f2bf(y);
        }
    } else {
        const float mean = block_sum(lsum, red) / static_cast<float>(a.N);
        float lvar = 0.f;
        for (int n = threadIdx.x; n < a.N; n += kRowThreads) {
            const float d = row[n] - mean;
            lvar += d * d;
        }
        const float var = block_sum(lvar, red) / static_cast<float>(a.N);
        const float rstd = rsqrtf(var + a.eps);
        for (int n = threadIdx.x; n < a.N; n += kRowThreads) {
            const float y = (row[n] - mean) * rstd * bf2f(a.norm_w[n]) + bf2f(a.norm_b[n]);
            a.xn_out[static_cast<size_t>(t) * a.ldn + n] = f2bf(y);
        }
    }
}

cudaError_t launch_consumer(cudaStream_t stream, const ConsumerArgs& a) {
    consumer_kernel<<<a.T, kRowThreads, a.N * sizeof(float), stream>>>(a);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ partial, int splitk, int T,
                                                       int N, int ldp, const bf16* __restrict__ bias, int act,
                                                       float scale, bf16* __restrict__ out, int ldo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * N) return;
    const int t = idx / N, n = idx - t * N;
    float acc = 0.f;
    for (int z = 0; z < splitk; ++z) acc += partial[(static_cast<size_t>(z) * T + t) * ldp + n];
    if (bias != nullptr) acc += bf2f(bias[n]);
    float x = bf16_round(acc);
    if (act == ACT_SILU) x = bf16_round(silu_f32(x));
    if (scale != 1.0f) x = bf16_round(x * scale);
    out[static_cast<size_t>(t) * ldo + n] = f2bf(x);
}

cudaError_t launch_bias_act(cudaStream_t stream, const float* partial, int splitk, int T, int N, int ldp,
                            const bf16* bias, int act, float scale, bf16* out, int ldo) {
    const int total = T * N;
    bias_act_kernel<<<(total + 255) / 256, 256, 0, stream>>>(partial, splitk, T, N, ldp, bias, act, scale,
                                                             out, ldo);
    return cudaGetLastError();
}

// One CTA per token; thread d handles head-dim element d of every head.
__global__ void __launch_bounds__(256) rope_kv_kernel(const RopeKvArgs a) {
    extern __shared__ float qkv[];          // [(n_heads + 2) * 256], bf16-rounded values
    const int t = blockIdx.x;
    if (t >= a.T) return;
    const int d = threadIdx.x;              // 0..255
    const int ncol = (a.n_heads + 2) * 256;
    for (int c = d; c < ncol; c += 256) {
        float acc = 0.f;
        for (int z = 0; z < a.splitk; ++z) acc += a.partial[(static_cast<size_t>(z) * a.T + t) * a.ldp + c];
        qkv[c] = bf16_round(acc);
    }
    __syncthreads();
    const int b = t / a.tokens_per_sample, i = t - b * a.tokens_per_sample;
    long long pos = a.position_ids[static_cast<size_t>(b) * a.tokens_per_sample + i];
    if (pos < 0) pos = 0;
    if (pos >= a.n_pos) pos = a.n_pos - 1;   // host validates the range; never read out of bounds
    const int j = d & 127;
    const float cs = a.cos_table[pos * 128 + j];
    const float sn = a.sin_table[pos * 128 + j];
    const int slot = a.slot_base + i;
    const size_t cache_off = (static_cast<size_t>(b) * a.n_slots + slot) * 256 + d;
    // rotate_half: first half pairs with -x[d+128], second half with +x[d-128]
    for (int h = 0; h <= a.n_heads; ++h) {       // h == n_heads is the key head
        if (h < a.n_heads && a.q_out == nullptr) continue;
        const float x = qkv[h * 256 + d];
        const float partner = (d < 128) ? -qkv[h * 256 + d + 128] : qkv[h * 256 + d - 128];
        const float y = bf16_round(bf16_round(x * cs) + bf16_round(partner * sn));
        if (h < a.n_heads)
            a.q_out[static_cast<size_t>(t) * (a.n_heads * 256) + h * 256 + d] = f2bf(y);
        else
            a.k_cache[cache_off] = f2bf(y);
    }
    a.v_cache[cache_off] = f2bf(qkv[(a.n_heads + 1) * 256 + d]);
}

cudaError_t launch_rope_kv(cudaStream_t stream, const RopeKvArgs& a) {
    const size_t smem = static_cast<size_t>(a.n_heads + 2) * 256 * sizeof(float);
    rope_kv_kernel<<<a.T, 256, smem, stream>>>(a);
    return cudaGetLastError();
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
