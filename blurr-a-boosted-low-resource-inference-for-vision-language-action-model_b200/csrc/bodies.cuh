// Device bodies of every non-GEMM kernel of the control step, shared by the stand-alone
// `__global__` wrappers (norm_consumers.cu, attention.cu, misc_kernels.cu) and by the persistent
// step kernel (step_kernel.cu), which runs them as work items between grid barriers.
//
// Loads of data produced earlier in the step go through L2 (`__ldcg`, `cp.async.cg`): inside the
// persistent kernel other SMs wrote it during the same launch and L1 is not coherent.  Weights,
// masks and position ids are constant for the launch and use ordinary loads.
#pragma once

#include "common.cuh"
#include "kernels.h"

namespace blurr {

__device__ __forceinline__ bf16 ldcg_bf16(const bf16* p) {
    const unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
    return *reinterpret_cast<const bf16*>(&u);
}
__device__ __forceinline__ bf16x8 ldcg_bf16x8(const bf16* p) {
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
    bf16x8 r;
    r.u[0] = u.x; r.u[1] = u.y; r.u[2] = u.z; r.u[3] = u.w;
    return r;
}

// ===========================================================================
// GEMM consumers (split-K sum, bias, residual / position add, norms), RoPE + KV append
// ===========================================================================
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
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(base + (z + 0) * slice_stride));
        const float4 p1 = __ldcg(reinterpret_cast<const float4*>(base + (z + 1) * slice_stride));
        const float4 p2 = __ldcg(reinterpret_cast<const float4*>(base + (z + 2) * slice_stride));
        const float4 p3 = __ldcg(reinterpret_cast<const float4*>(base + (z + 3) * slice_stride));
        acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;     // fixed order z = 0, 1, 2, ...
        acc.x += p1.x; acc.y += p1.y; acc.z += p1.z; acc.w += p1.w;
        acc.x += p2.x; acc.y += p2.y; acc.z += p2.z; acc.w += p2.w;
        acc.x += p3.x; acc.y += p3.y; acc.z += p3.z; acc.w += p3.w;
    }
    for (; z < splitk; ++z) {
        const float4 p = __ldcg(reinterpret_cast<const float4*>(base + z * slice_stride));
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    return acc;
}

__device__ __forceinline__ float4 load_bf16x4(const bf16* p) {
    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p));
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
__device__ __forceinline__ void consumer_body(const ConsumerArgs& a, const int t) {
    __shared__ float red[kRowThreads / 32];
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

__device__ __forceinline__ void bias_act_body(const float* partial, int splitk, int T, int N, int ldp,
                                              const bf16* bias, int act, float scale, bf16* out, int ldo,
                                              const int bx) {
    const int idx = bx * 256 + threadIdx.x;
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

// One CTA per token.  Work items: for every rotated head (queries + the key head) 32 pairs of
// float4 column groups (dims [4j,4j+4) and [128+4j,128+4j+4): the rotate_half partners), plus 64
// plain float4 groups of the value head.
__device__ __forceinline__ void rope_kv_body(const RopeKvArgs& a, const int t) {
    const int b = t / a.tokens_per_sample, i = t - b * a.tokens_per_sample;
    long long pos = __ldcg(a.position_ids + static_cast<size_t>(b) * a.tokens_per_sample + i);
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


// ===========================================================================
// attention
// ===========================================================================
static constexpr int kAttnThreads = 256;
static constexpr int kBK = 64;   // keys per streamed block


// dst: [nrows][LDS]; 16-byte chunks; rows >= nrows_valid and columns >= hd are zero-filled
template <int LDS, int CH>
__device__ __forceinline__ void load_rows_async(bf16* dst, const bf16* src, int ld, int row0, int nrows,
                                                int nrows_valid, int hd) {
    for (int idx = threadIdx.x; idx < nrows * CH; idx += kAttnThreads) {
        const int r = idx / CH, c = idx - r * CH;
        const bool valid = (row0 + r < nrows_valid) && (c * 8 < hd);
        const bf16* g = valid ? (src + static_cast<size_t>(row0 + r) * ld + c * 8) : src;
        cp_async_16(dst + r * LDS + c * 8, g, valid);
    }
}

// BM query rows per CTA; 8 warps = (BM/16) row groups x WC column groups.
template <int HD_PAD, int BM, bool GEMMA>
__device__ __forceinline__ void attn_mma_body(const AttnMmaArgs& a, uint8_t* smem_attn, const int qt, const int h,
                                              const int b) {
    constexpr int WR = BM / 16;                 // row groups
    constexpr int WC = 8 / WR;                  // column groups
    constexpr int KPW = kBK / WC;               // keys per column group per streamed block (16 or 32)
    constexpr int NT_S = KPW / 8;               // logit n-tiles per warp per block
    constexpr int NT_ALL = HD_PAD / 8;          // output n-tiles over the head dim
    constexpr int NT_PV = (NT_ALL + WC - 1) / WC;
    constexpr int NP_PV = (NT_PV + 1) / 2;
    constexpr int CH = HD_PAD / 8;              // 16-byte chunks per row that carry data
    constexpr int LDS_MIN = WC * NP_PV * 16 > HD_PAD ? WC * NP_PV * 16 : HD_PAD;
    constexpr int LDS = LDS_MIN + 8;            // smem row stride (elements), conflict-free for ldmatrix
    const int nkb = (a.n_keys + kBK - 1) / kBK;
    const int ldl = nkb * kBK + 8;              // logit row stride (elements)
    bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
    bf16* KVs = Qs + BM * LDS;                  // 2 buffers
    bf16* Ls = KVs + 2 * kBK * LDS;             // [BM][ldl]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp % WR, wc = warp / WR;
    const int q_row0 = qt * BM;

    const bf16* qbase = a.q + static_cast<size_t>(b) * a.q_per_sample * a.ldq + a.q_col0 + h * a.head_stride_q;
    const bf16* kbase = a.k + static_cast<size_t>(b) * a.kv_per_sample * a.ldk + a.k_col0 + h * a.head_stride_kv;
    const bf16* vbase = a.v + static_cast<size_t>(b) * a.kv_per_sample * a.ldv + a.v_col0 + h * a.head_stride_kv;

    load_rows_async<LDS, CH>(Qs, qbase, a.ldq, q_row0, BM, a.q_per_sample, a.hd);
    load_rows_async<LDS, CH>(KVs, kbase, a.ldk, 0, kBK, a.n_keys, a.hd);
    cp_async_commit();

    // ---------------- phase S: logits = chain(Q K^T) -> Ls (bf16) ----------------
    for (int kb = 0; kb < nkb; ++kb) {
        bf16* Kcur = KVs + (kb & 1) * kBK * LDS;
        if (kb + 1 < nkb) {
            load_rows_async<LDS, CH>(KVs + ((kb + 1) & 1) * kBK * LDS, kbase, a.ldk, (kb + 1) * kBK, kBK, a.n_keys, a.hd);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        float acc[NT_S][4];
#pragma unroll
        for (int i = 0; i < NT_S; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
        for (int kk = 0; kk < HD_PAD / 16; ++kk) {
            uint32_t af[4];
            ldmatrix_x4(af, smem_u32(Qs + (wr * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
#pragma unroll
            for (int np = 0; np < NT_S / 2; ++np) {
                uint32_t bfr[4];
                const int mi = lane >> 3;
                const int key = wc * KPW + np * 16 + (mi >> 1) * 8 + (lane & 7);
                ldmatrix_x4(bfr, smem_u32(Kcur + key * LDS + kk * 16 + (mi & 1) * 8));
                mma_bf16_16816(acc[np * 2 + 0], af, bfr[0], bfr[1]);
                mma_bf16_16816(acc[np * 2 + 1], af, bfr[2], bfr[3]);
            }
        }
        // epilogue of this key block: rounding chain, write bf16 logits
#pragma unroll
        for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = wr * 16 + (lane >> 2) + half * 8;
                const int kcol = kb * kBK + wc * KPW + nt * 8 + (lane & 3) * 2;
                float s0 = bf16_round(acc[nt][half * 2 + 0]);
                float s1 = bf16_round(acc[nt][half * 2 + 1]);
                if (GEMMA) {
                    s0 = bf16_round(s0 * 0.0625f);              // / sqrt(256)
                    s1 = bf16_round(s1 * 0.0625f);
                    const float inv50 = 1.0f / 50.0f;            // ATen: a * (1 / b) for a scalar divisor
                    s0 = bf16_round(s0 * inv50);
                    s1 = bf16_round(s1 * inv50);
                    s0 = bf16_round(tanhf(s0));
                    s1 = bf16_round(tanhf(s1));
                    s0 = bf16_round(s0 * 50.0f);
                    s1 = bf16_round(s1 * 50.0f);
                    const int qr = q_row0 + r;
                    if (qr < a.q_per_sample) {
                        const bf16* mrow = a.mask + static_cast<size_t>(b) * a.mask_bstride +
                                           static_cast<size_t>(a.q_row_offset + qr) * a.mask_rstride;
                        if (kcol < a.n_keys) s0 = bf16_round(s0 + bf2f(mrow[kcol]));
                        if (kcol + 1 < a.n_keys) s1 = bf16_round(s1 + bf2f(mrow[kcol + 1]));
                    }
                } else {
                    s0 = bf16_round(s0 * a.scale);
                    s1 = bf16_round(s1 * a.scale);
                }
                *reinterpret_cast<uint32_t*>(Ls + r * ldl + kcol) = pack_bf16x2(s0, s1);
            }
        }
        __syncthreads();   // all warps done with Kcur before it is overwritten; Ls visible
    }

    // prefetch V block 0 while the softmax runs
    load_rows_async<LDS, CH>(KVs, vbase, a.ldv, 0, kBK, a.n_keys, a.hd);
    cp_async_commit();

    // ---------------- softmax: fp32 over bf16 logits, result bf16 in place ----------------
    for (int rr = 0; rr < BM / 8; ++rr) {
        bf16* lrow = Ls + (warp * (BM / 8) + rr) * ldl;
        float m = -INFINITY;
        for (int c = lane; c < a.n_keys; c += 32) m = fmaxf(m, bf2f(lrow[c]));
        m = warp_max(m);
        float sum = 0.f;
        for (int c = lane; c < a.n_keys; c += 32) sum += expf(bf2f(lrow[c]) - m);
        sum = warp_sum(sum);
        for (int c = lane; c < nkb * kBK; c += 32) {
            float p = 0.f;
            if (c < a.n_keys) p = expf(bf2f(lrow[c]) - m) / sum;
            lrow[c] = f2bf(p);
        }
    }
    __syncthreads();

    // ---------------- phase PV ----------------
    float oacc[NP_PV * 2][4];
#pragma unroll
    for (int i = 0; i < NP_PV * 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
    const int nt0 = wc * NT_PV;                 // first output n-tile of this column group

    for (int kb = 0; kb < nkb; ++kb) {
        bf16* Vcur = KVs + (kb & 1) * kBK * LDS;
        if (kb + 1 < nkb) {
            load_rows_async<LDS, CH>(KVs + ((kb + 1) & 1) * kBK * LDS, vbase, a.ldv, (kb + 1) * kBK, kBK, a.n_keys, a.hd);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk) {
            uint32_t af[4];
            ldmatrix_x4(af, smem_u32(Ls + (wr * 16 + (lane & 15)) * ldl + kb * kBK + kk * 16 + (lane >> 4) * 8));
#pragma unroll
            for (int np = 0; np < NP_PV; ++np) {
                if ((nt0 + np * 2) >= NT_ALL) continue;          // column group past the head dim
                uint32_t bfr[4];
                const int mi = lane >> 3;
                const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
                const int dim = (nt0 + np * 2) * 8 + (mi >> 1) * 8;
                ldmatrix_x4_trans(bfr, smem_u32(Vcur + key * LDS + dim));
                mma_bf16_16816(oacc[np * 2 + 0], af, bfr[0], bfr[1]);
                if (np * 2 + 1 < NT_PV && nt0 + np * 2 + 1 < NT_ALL)
                    mma_bf16_16816(oacc[np * 2 + 1], af, bfr[2], bfr[3]);
            }
        }
        __syncthreads();
    }

    // ---------------- store ----------------
    bf16* obase = a.out + static_cast<size_t>(b) * a.q_per_sample * a.ldo + a.o_col0 + h * a.head_stride_q;
#pragma unroll
    for (int nt = 0; nt < NT_PV; ++nt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = q_row0 + wr * 16 + (lane >> 2) + half * 8;
            const int dim = (nt0 + nt) * 8 + (lane & 3) * 2;
            if (r < a.q_per_sample && dim < a.hd)
                *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(r) * a.ldo + dim) =
                    pack_bf16x2(oacc[nt][half * 2 + 0], oacc[nt][half * 2 + 1]);
        }
    }
}

template <int HD_PAD, int BM>
inline size_t attn_smem_bytes(int n_keys) {
    constexpr int WC = 8 / (BM / 16);
    constexpr int NT_PV = (HD_PAD / 8 + WC - 1) / WC;
    constexpr int NP_PV = (NT_PV + 1) / 2;
    constexpr int LDS_MIN = WC * NP_PV * 16 > HD_PAD ? WC * NP_PV * 16 : HD_PAD;
    constexpr int LDS = LDS_MIN + 8;
    const int nkb = (n_keys + kBK - 1) / kBK;
    return static_cast<size_t>(BM + 2 * kBK) * LDS * 2 + static_cast<size_t>(BM) * (nkb * kBK + 8) * 2;
}

// ---------------------------------------------------------------------------
// few-query attention over the KV cache: one CTA per (head, query, sample); the 8 warps split
// the keys, 4 keys in flight per warp (16-byte K/V loads per lane, shuffle reductions).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void attn_fewq_body(const JointAttnArgs& a, float* fq_smem, const int h, const int qi,
                                               const int b) {
    // fq_smem: logits[n_keys_pad] | partial_out[8][256]
    __shared__ float red[8];
    const int n_pad = (a.n_keys + 3) & ~3;
    float* lg = fq_smem;
    float* po = fq_smem + n_pad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ldq = a.n_heads * 256;
    const size_t qrow = static_cast<size_t>(b) * a.q_per_sample + qi;
    const bf16* kc = a.k_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* vc = a.v_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* mrow = a.mask + static_cast<size_t>(b) * a.mask_bstride +
                       static_cast<size_t>(a.q_row_offset + qi) * a.mask_rstride;
    float qreg[8];
    {
        const bf16x8 qv = ldcg_bf16x8(a.q + qrow * ldq + h * 256 + lane * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = unpack_bf16x2(qv.u[i]);
            qreg[2 * i] = f.x; qreg[2 * i + 1] = f.y;
        }
    }
    // ---- logits: warp w owns keys [w*kpw, (w+1)*kpw), 4 at a time ----
    const int kpw = ((a.n_keys + 7) / 8 + 3) & ~3;
    const int k_begin = warp * kpw, k_end = min(k_begin + kpw, a.n_keys);
    for (int k0 = k_begin; k0 < k_end; k0 += 4) {
        bf16x8 kv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = min(k0 + j, a.n_keys - 1);
            kv[j] = ldcg_bf16x8(kc + static_cast<size_t>(k) * 256 + lane * 8);
        }
        float dot[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16x2(kv[j].u[i]);
                d += qreg[2 * i] * f.x + qreg[2 * i + 1] * f.y;
            }
            dot[j] = d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
        }
        if (lane < 4 && k0 + lane < k_end) {
            const int k = k0 + lane;
            float s = bf16_round(lane == 0 ? dot[0] : lane == 1 ? dot[1] : lane == 2 ? dot[2] : dot[3]);
            s = bf16_round(s * 0.0625f);
            s = bf16_round(s * (1.0f / 50.0f));
            s = bf16_round(tanhf(s));
            s = bf16_round(s * 50.0f);
            s = bf16_round(s + bf2f(mrow[k]));
            lg[k] = s;
        }
    }
    __syncthreads();
    // ---- softmax (fp32) -> bf16 probabilities ----
    float m = -INFINITY;
    for (int k = tid; k < a.n_keys; k += 256) m = fmaxf(m, lg[k]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int k = tid; k < a.n_keys; k += 256) sum += expf(lg[k] - m);
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i];
    __syncthreads();
    for (int k = tid; k < a.n_keys; k += 256) lg[k] = bf16_round(expf(lg[k] - m) / sum);
    __syncthreads();
    // ---- out = P V: warp w accumulates its keys for all 256 dims (8 per lane) ----
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k0 = k_begin; k0 < k_end; k0 += 4) {
        bf16x8 vv[4];
        float pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = min(k0 + j, a.n_keys - 1);
            vv[j] = ldcg_bf16x8(vc + static_cast<size_t>(k) * 256 + lane * 8);
            pk[j] = (k0 + j < k_end) ? lg[k] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_bf16x2(vv[j].u[i]);
                acc[2 * i] += pk[j] * f.x;
                acc[2 * i + 1] += pk[j] * f.y;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) po[warp * 256 + lane * 8 + i] = acc[i];
    __syncthreads();
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) o += po[w * 256 + tid];
    a.out[qrow * ldq + h * 256 + tid] = f2bf(o);
}


// ===========================================================================
// small kernels at the edges of the step
// ===========================================================================
__device__ __forceinline__ void im2col_body(const bf16* px, long long sb, long long sc, long long sh, long long sw,
                                            bf16* patches, int ldp, const int p, const int b) {
    // p: patch index within the image, row-major 16x16
    const int ph = p >> 4, pw = p & 15;
    bf16* dst = patches + (static_cast<size_t>(b) * 256 + p) * ldp;
    for (int idx = threadIdx.x; idx < 588; idx += blockDim.x) {
        const int c = idx / 196, rem = idx - c * 196;
        const int kh = rem / 14, kw = rem - kh * 14;
        dst[idx] = px[b * sb + c * sc + static_cast<long long>(ph * 14 + kh) * sh +
                      static_cast<long long>(pw * 14 + kw) * sw];
    }
}

__device__ __forceinline__ void embed_merge_body(const int64_t* ids, int seq, const bf16* table, long long vocab,
                                                 const bf16* img, int n_img, int hidden, long long image_token,
                                                 long long pad_token, float inv_div, float normalizer, bf16* out,
                                                 int* err_flag, const int pos, const int b) {
    __shared__ int s_rank;
    __syncthreads();                     // s_rank may still be in use by the previous work item
    const int64_t* row = ids + static_cast<size_t>(b) * seq;
    const long long id = __ldcg(row + pos);
    bf16* dst = out + (static_cast<size_t>(b) * seq + pos) * hidden;
    if (id == image_token) {
        if (threadIdx.x == 0) s_rank = 0;
        __syncthreads();
        int cnt = 0;
        for (int i = threadIdx.x; i < pos; i += blockDim.x) cnt += (__ldcg(row + i) == image_token) ? 1 : 0;
        cnt = static_cast<int>(warp_sum(static_cast<float>(cnt)));
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_rank, cnt);
        __syncthreads();
        const int rank = s_rank;
        if (rank >= n_img) {             // the reference would raise a shape error here
            if (threadIdx.x == 0) *err_flag = 1;
            return;
        }
        const bf16* src = img + (static_cast<size_t>(b) * n_img + rank) * hidden;
        for (int n = threadIdx.x; n < hidden; n += blockDim.x) {
            float x = bf16_round(bf2f(ldcg_bf16(src + n)) * inv_div);     // image_features / sqrt(hidden)
            x = bf16_round(x * normalizer);                   // embeds *= bf16(sqrt(hidden))
            dst[n] = f2bf(x);
        }
    } else if (id != pad_token) {
        if (id < 0 || id >= vocab) {
            if (threadIdx.x == 0) *err_flag = 2;
            return;
        }
        const bf16* src = table + static_cast<size_t>(id) * hidden;
        for (int n = threadIdx.x; n < hidden; n += blockDim.x)
            dst[n] = f2bf(bf16_round(bf2f(src[n]) * normalizer));
    } else {
        // torch.full(..., pad_token_id) rows, then *= normalizer
        const float x = bf16_round(bf16_round(static_cast<float>(pad_token)) * normalizer);
        for (int n = threadIdx.x; n < hidden; n += blockDim.x) dst[n] = f2bf(x);
    }
}

__device__ __forceinline__ void small_k_linear_body(const bf16* x, int T, int K, const bf16* W, const bf16* bias,
                                                    int N, float scale, bf16* y, int ldy, int col_off,
                                                    const bf16* time_row, int time_cols, const int bx, const int t) {
    const int n = bx * 256 + threadIdx.x;
    if (n < N) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += bf2f(ldcg_bf16(x + t * K + k)) * bf2f(W[n * K + k]);
        acc += bf2f(bias[n]);
        float v = bf16_round(acc);
        if (scale != 1.0f) v = bf16_round(v * scale);
        y[static_cast<size_t>(t) * ldy + col_off + n] = f2bf(v);
    }
    if (time_row != nullptr && n < time_cols) y[static_cast<size_t>(t) * ldy + n] = time_row[n];
}

__device__ __forceinline__ void action_tail_body(const bf16* xn, int T, int hidden, const bf16* W, const bf16* bias,
                                                 int action_dim, float dt, bf16* action, bf16* vel_tap,
                                                 const int bx) {
    const int gw = (bx * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= T * action_dim) return;
    const int t = gw / action_dim, d = gw - t * action_dim;
    float acc = 0.f;
    for (int k = lane; k < hidden; k += 32)
        acc += bf2f(ldcg_bf16(xn + static_cast<size_t>(t) * hidden + k)) * bf2f(W[static_cast<size_t>(d) * hidden + k]);
    acc = warp_sum(acc);
    if (lane == 0) {
        const float vel = bf16_round(acc + bf2f(bias[d]));
        if (vel_tap != nullptr) vel_tap[gw] = f2bf(vel);
        const float step = bf16_round(dt * vel);                       // delta_t * action_vel
        action[gw] = f2bf(bf16_round(bf2f(ldcg_bf16(action + gw)) + step));        // action += ...
    }
}

__device__ __forceinline__ void clamp_copy_body(const bf16* src, bf16* dst, int n, int do_clamp, float clip,
                                                const int bx) {
    const int i = bx * 256 + threadIdx.x;
    if (i >= n) return;
    float v = bf2f(ldcg_bf16(src + i));
    if (do_clamp) v = (v < -clip) ? -clip : ((v > clip) ? clip : v);   // NaN propagates like torch.clamp
    dst[i] = f2bf(v);
}


}  // namespace blurr
