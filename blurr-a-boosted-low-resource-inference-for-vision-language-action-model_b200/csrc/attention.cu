// Attention kernels of the Pi-0 path.
//
//  * attn_mma_kernel<HD_PAD, GEMMA>: "full-row" attention for the two many-query cases
//      - SigLIP MHA, 16 heads x 256 tokens x head_dim 72 (siglip.py:133-152), no mask;
//      - Gemma joint-attention prefill, 8 query heads sharing one KV head (MQA), 276 query rows
//        x 277 keys x head_dim 256 with tanh soft-clamp and additive block mask
//        (joint_model.py:273-288).
//    The tiles are tiny (<= 64 x 320 logits), so they run on warp-level mma.sync (bf16, fp32
//    accumulate) with ldmatrix-fed fragments; the whole logit row is kept in shared memory so
//    the softmax is the reference's exact two-pass fp32 softmax over bf16-rounded logits.
//  * attn_fewq_kernel: 1 (proprio) or 4 (action) query rows per sample over the KV cache
//    (joint_model.py:164-170 "append_non_active"): a bandwidth kernel, coalesced 16-byte K
//    loads + warp-shuffle dot products, coalesced V accumulation.
//
// Rounding points: QK^T -> bf16; every scale / tanh / mask op
// -> bf16; softmax in fp32 -> bf16; PV -> bf16.  Divisions by Python scalars are done as the
// ATen CUDA kernels do them (multiplication by the fp32 reciprocal).
#include "common.cuh"
#include "kernels.h"

namespace blurr {

static constexpr int kAttnThreads = 256;
static constexpr int kBM = 64;   // query rows per CTA
static constexpr int kBK = 64;   // keys per streamed block

struct AttnMmaArgs {
    const bf16* q; int ldq; int q_col0; int q_per_sample;
    const bf16* k; int ldk; int k_col0; int kv_per_sample;
    const bf16* v; int ldv; int v_col0;
    bf16* out; int ldo; int o_col0;
    int hd;            // real head dim (72 / 256)
    int head_stride_q; // column step between query heads (hd)
    int head_stride_kv;// column step between kv heads (0 for MQA)
    int n_keys;
    float scale;       // SigLIP: head_dim^-0.5
    const bf16* mask; long long mask_bstride, mask_rstride; int q_row_offset;
};

template <int HD_PAD>
__device__ __forceinline__ void load_rows_async(bf16* dst, const bf16* src, int ld, int row0, int nrows_valid,
                                                int hd) {
    // dst: [64][HD_PAD + 8]; 16-byte chunks; rows >= nrows_valid and columns >= hd are zero-filled
    constexpr int CH = HD_PAD / 8;
    for (int idx = threadIdx.x; idx < kBK * CH; idx += kAttnThreads) {
        const int r = idx / CH, c = idx - r * CH;
        const bool valid = (row0 + r < nrows_valid) && (c * 8 < hd);
        const bf16* g = valid ? (src + static_cast<size_t>(row0 + r) * ld + c * 8) : src;
        cp_async_16(dst + r * (HD_PAD + 8) + c * 8, g, valid);
    }
}

template <int HD_PAD, bool GEMMA>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_mma_kernel(const AttnMmaArgs a) {
    constexpr int LDS = HD_PAD + 8;            // smem row stride of Q/K/V tiles (elements)
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int nkb = (a.n_keys + kBK - 1) / kBK;
    const int ldl = nkb * kBK + 8;             // logit row stride (elements)
    bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
    bf16* KVs = Qs + kBM * LDS;                // 2 buffers
    bf16* Ls = KVs + 2 * kBK * LDS;            // [64][ldl]

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp & 3, wc = warp >> 2;
    const int q_row0 = qt * kBM;

    const bf16* qbase = a.q + static_cast<size_t>(b) * a.q_per_sample * a.ldq + a.q_col0 + h * a.head_stride_q;
    const bf16* kbase = a.k + static_cast<size_t>(b) * a.kv_per_sample * a.ldk + a.k_col0 + h * a.head_stride_kv;
    const bf16* vbase = a.v + static_cast<size_t>(b) * a.kv_per_sample * a.ldv + a.v_col0 + h * a.head_stride_kv;

    load_rows_async<HD_PAD>(Qs, qbase, a.ldq, q_row0, a.q_per_sample, a.hd);
    load_rows_async<HD_PAD>(KVs, kbase, a.ldk, 0, a.n_keys, a.hd);
    cp_async_commit();

    // ---------------- phase S: logits = chain(Q K^T) -> Ls (bf16) ----------------
    for (int kb = 0; kb < nkb; ++kb) {
        bf16* Kcur = KVs + (kb & 1) * kBK * LDS;
        if (kb + 1 < nkb) {
            load_rows_async<HD_PAD>(KVs + ((kb + 1) & 1) * kBK * LDS, kbase, a.ldk, (kb + 1) * kBK, a.n_keys, a.hd);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
        for (int kk = 0; kk < HD_PAD / 16; ++kk) {
            uint32_t af[4];
            ldmatrix_x4(af, smem_u32(Qs + (wr * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t bfr[4];
                const int mi = lane >> 3;
                const int key = wc * 32 + np * 16 + (mi >> 1) * 8 + (lane & 7);
                ldmatrix_x4(bfr, smem_u32(Kcur + key * LDS + kk * 16 + (mi & 1) * 8));
                mma_bf16_16816(acc[np * 2 + 0], af, bfr[0], bfr[1]);
                mma_bf16_16816(acc[np * 2 + 1], af, bfr[2], bfr[3]);
            }
        }
        // epilogue of this key block: rounding chain, write bf16 logits
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = wr * 16 + (lane >> 2) + half * 8;
                const int kcol = kb * kBK + wc * 32 + nt * 8 + (lane & 3) * 2;
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
    load_rows_async<HD_PAD>(KVs, vbase, a.ldv, 0, a.n_keys, a.hd);
    cp_async_commit();

    // ---------------- softmax: fp32 over bf16 logits, result bf16 in place ----------------
    for (int rr = 0; rr < kBM / 8; ++rr) {
        bf16* lrow = Ls + (warp * (kBM / 8) + rr) * ldl;
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
    constexpr int DHALF = HD_PAD / 2;          // dims per column-warp
    constexpr int NTILES = DHALF / 8;          // 16 (HD 256) or 5 (HD 80)
    constexpr int NPAIRS = (NTILES + 1) / 2;
    float oacc[NPAIRS * 2][4];
#pragma unroll
    for (int i = 0; i < NPAIRS * 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;

    for (int kb = 0; kb < nkb; ++kb) {
        bf16* Vcur = KVs + (kb & 1) * kBK * LDS;
        if (kb + 1 < nkb) {
            load_rows_async<HD_PAD>(KVs + ((kb + 1) & 1) * kBK * LDS, vbase, a.ldv, (kb + 1) * kBK, a.n_keys, a.hd);
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
            for (int np = 0; np < NPAIRS; ++np) {
                uint32_t bfr[4];
                const int mi = lane >> 3;
                const int key = kk * 16 + (mi & 1) * 8 + (lane & 7);
                const int dim = wc * DHALF + np * 16 + (mi >> 1) * 8;
                ldmatrix_x4_trans(bfr, smem_u32(Vcur + key * LDS + dim));
                mma_bf16_16816(oacc[np * 2 + 0], af, bfr[0], bfr[1]);
                if (np * 2 + 1 < NTILES) mma_bf16_16816(oacc[np * 2 + 1], af, bfr[2], bfr[3]);
            }
        }
        __syncthreads();
    }

    // ---------------- store ----------------
    bf16* obase = a.out + static_cast<size_t>(b) * a.q_per_sample * a.ldo + a.o_col0 + h * a.head_stride_q;
#pragma unroll
    for (int nt = 0; nt < NTILES; ++nt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = q_row0 + wr * 16 + (lane >> 2) + half * 8;
            const int dim = wc * DHALF + nt * 8 + (lane & 3) * 2;
            if (r < a.q_per_sample && dim < a.hd)
                *reinterpret_cast<uint32_t*>(obase + static_cast<size_t>(r) * a.ldo + dim) =
                    pack_bf16x2(oacc[nt][half * 2 + 0], oacc[nt][half * 2 + 1]);
        }
    }
}

template <int HD_PAD>
static size_t attn_smem_bytes(int n_keys) {
    const int nkb = (n_keys + kBK - 1) / kBK;
    return static_cast<size_t>(kBM + 2 * kBK) * (HD_PAD + 8) * 2 + static_cast<size_t>(kBM) * (nkb * kBK + 8) * 2;
}

cudaError_t launch_siglip_attention(cudaStream_t stream, const bf16* qkv, int ld_qkv, int batch, int seq,
                                    int n_heads, int hidden, bf16* out, int ld_out) {
    AttnMmaArgs a{};
    const int hd = hidden / n_heads;           // 72
    if (hd > 80) return cudaErrorInvalidValue;
    a.q = qkv; a.ldq = ld_qkv; a.q_col0 = 0; a.q_per_sample = seq;
    a.k = qkv; a.ldk = ld_qkv; a.k_col0 = hidden; a.kv_per_sample = seq;
    a.v = qkv; a.ldv = ld_qkv; a.v_col0 = 2 * hidden;
    a.out = out; a.ldo = ld_out; a.o_col0 = 0;
    a.hd = hd; a.head_stride_q = hd; a.head_stride_kv = hd; a.n_keys = seq;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    {   // Python: head_dim ** -0.5 evaluated in double, then used as an fp32 scalar operand
        const double s = pow(static_cast<double>(hd), -0.5);
        a.scale = static_cast<float>(s);
    }
    a.mask = nullptr;
    const size_t smem = attn_smem_bytes<80>(seq);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<80, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((seq + kBM - 1) / kBM, n_heads, batch);
    attn_mma_kernel<80, false><<<grid, kAttnThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_joint_attention_prefill(cudaStream_t stream, const JointAttnArgs& j) {
    AttnMmaArgs a{};
    a.q = j.q; a.ldq = j.n_heads * 256; a.q_col0 = 0; a.q_per_sample = j.q_per_sample;
    a.k = j.k_cache; a.ldk = 256; a.k_col0 = 0; a.kv_per_sample = j.n_slots;
    a.v = j.v_cache; a.ldv = 256; a.v_col0 = 0;
    a.out = j.out; a.ldo = j.n_heads * 256; a.o_col0 = 0;
    a.hd = 256; a.head_stride_q = 256; a.head_stride_kv = 0; a.n_keys = j.n_keys;
    a.scale = 0.f;
    a.mask = j.mask; a.mask_bstride = j.mask_bstride; a.mask_rstride = j.mask_rstride;
    a.q_row_offset = j.q_row_offset;
    const size_t smem = attn_smem_bytes<256>(j.n_keys);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<256, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((j.q_per_sample + kBM - 1) / kBM, j.n_heads, j.batch);
    attn_mma_kernel<256, true><<<grid, kAttnThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// few-query attention over the KV cache: one CTA per (head, query, sample)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_fewq_kernel(const JointAttnArgs a) {
    extern __shared__ float fq_smem[];       // q[256] | logits[n_keys]
    __shared__ float red[8];
    float* qf = fq_smem;
    float* lg = fq_smem + 256;
    const int h = blockIdx.x, qi = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ldq = a.n_heads * 256;
    const size_t qrow = static_cast<size_t>(b) * a.q_per_sample + qi;
    qf[tid] = bf2f(a.q[qrow * ldq + h * 256 + tid]);
    __syncthreads();

    const bf16* kc = a.k_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* vc = a.v_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* mrow = a.mask + static_cast<size_t>(b) * a.mask_bstride +
                       static_cast<size_t>(a.q_row_offset + qi) * a.mask_rstride;
    float qreg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qreg[i] = qf[lane * 8 + i];
    for (int k = warp; k < a.n_keys; k += 8) {
        const bf16x8 kv = *reinterpret_cast<const bf16x8*>(kc + static_cast<size_t>(k) * 256 + lane * 8);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = unpack_bf16x2(kv.u[i]);
            dot += qreg[2 * i] * f.x + qreg[2 * i + 1] * f.y;
        }
        dot = warp_sum(dot);
        if (lane == 0) {
            float s = bf16_round(dot);
            s = bf16_round(s * 0.0625f);
            s = bf16_round(s * (1.0f / 50.0f));
            s = bf16_round(tanhf(s));
            s = bf16_round(s * 50.0f);
            s = bf16_round(s + bf2f(mrow[k]));
            lg[k] = s;
        }
    }
    __syncthreads();
    // softmax (fp32) -> bf16 probabilities
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
    // out[d] = sum_k P[k] V[k][d]
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int k = 0;
    for (; k + 3 < a.n_keys; k += 4) {
        acc0 += lg[k] * bf2f(vc[static_cast<size_t>(k) * 256 + tid]);
        acc1 += lg[k + 1] * bf2f(vc[static_cast<size_t>(k + 1) * 256 + tid]);
        acc2 += lg[k + 2] * bf2f(vc[static_cast<size_t>(k + 2) * 256 + tid]);
        acc3 += lg[k + 3] * bf2f(vc[static_cast<size_t>(k + 3) * 256 + tid]);
    }
    for (; k < a.n_keys; ++k) acc0 += lg[k] * bf2f(vc[static_cast<size_t>(k) * 256 + tid]);
    a.out[qrow * ldq + h * 256 + tid] = f2bf((acc0 + acc1) + (acc2 + acc3));
}

cudaError_t launch_joint_attention_fewq(cudaStream_t stream, const JointAttnArgs& a) {
    dim3 grid(a.n_heads, a.q_per_sample, a.batch);
    const size_t smem = (256 + a.n_keys) * sizeof(float);
    attn_fewq_kernel<<<grid, 256, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace blurr
