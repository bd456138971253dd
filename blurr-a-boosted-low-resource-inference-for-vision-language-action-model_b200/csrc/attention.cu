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
#include "launch.cuh"

namespace blurr {

static constexpr int kAttnThreads = 256;
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
__global__ void __launch_bounds__(kAttnThreads) attn_mma_kernel(const AttnMmaArgs a) {
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
    extern __shared__ __align__(16) uint8_t smem_attn[];
    const int nkb = (a.n_keys + kBK - 1) / kBK;
    const int ldl = nkb * kBK + 8;              // logit row stride (elements)
    bf16* Qs = reinterpret_cast<bf16*>(smem_attn);
    bf16* KVs = Qs + BM * LDS;                  // 2 buffers
    bf16* Ls = KVs + 2 * kBK * LDS;             // [BM][ldl]

    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp % WR, wc = warp / WR;
    const int q_row0 = qt * BM;

    const bf16* qbase = a.q + static_cast<size_t>(b) * a.q_per_sample * a.ldq + a.q_col0 + h * a.head_stride_q;
    const bf16* kbase = a.k + static_cast<size_t>(b) * a.kv_per_sample * a.ldk + a.k_col0 + h * a.head_stride_kv;
    const bf16* vbase = a.v + static_cast<size_t>(b) * a.kv_per_sample * a.ldv + a.v_col0 + h * a.head_stride_kv;

    pdl_wait();
    pdl_trigger();
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
static size_t attn_smem_bytes(int n_keys) {
    constexpr int WC = 8 / (BM / 16);
    constexpr int NT_PV = (HD_PAD / 8 + WC - 1) / WC;
    constexpr int NP_PV = (NT_PV + 1) / 2;
    constexpr int LDS_MIN = WC * NP_PV * 16 > HD_PAD ? WC * NP_PV * 16 : HD_PAD;
    constexpr int LDS = LDS_MIN + 8;
    const int nkb = (n_keys + kBK - 1) / kBK;
    return static_cast<size_t>(BM + 2 * kBK) * LDS * 2 + static_cast<size_t>(BM) * (nkb * kBK + 8) * 2;
}

static constexpr int kAttnBM = 32;

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
    // Python: head_dim ** -0.5 evaluated in double, then used as an fp32 scalar operand
    a.scale = static_cast<float>(pow(static_cast<double>(hd), -0.5));
    a.mask = nullptr;
    const size_t smem = attn_smem_bytes<80, kAttnBM>(seq);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<80, kAttnBM, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((seq + kAttnBM - 1) / kAttnBM, n_heads, batch);
    return launch_kernel(attn_mma_kernel<80, kAttnBM, false>, grid, dim3(kAttnThreads), smem, stream, a);
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
    const size_t smem = attn_smem_bytes<256, kAttnBM>(j.n_keys);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<256, kAttnBM, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    dim3 grid((j.q_per_sample + kAttnBM - 1) / kAttnBM, j.n_heads, j.batch);
    return launch_kernel(attn_mma_kernel<256, kAttnBM, true>, grid, dim3(kAttnThreads), smem, stream, a);
}

// ---------------------------------------------------------------------------
// few-query attention over the KV cache: one CTA per (head, query, sample); the 8 warps split
// the keys, 4 keys in flight per warp (16-byte K/V loads per lane, shuffle reductions).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_fewq_kernel(const JointAttnArgs a) {
    extern __shared__ float fq_smem[];       // logits[n_keys_pad] | partial_out[8][256]
    __shared__ float red[8];
    const int n_pad = (a.n_keys + 3) & ~3;
    float* lg = fq_smem;
    float* po = fq_smem + n_pad;
    const int h = blockIdx.x, qi = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ldq = a.n_heads * 256;
    const size_t qrow = static_cast<size_t>(b) * a.q_per_sample + qi;
    pdl_wait();
    pdl_trigger();

    const bf16* kc = a.k_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* vc = a.v_cache + static_cast<size_t>(b) * a.n_slots * 256;
    const bf16* mrow = a.mask + static_cast<size_t>(b) * a.mask_bstride +
                       static_cast<size_t>(a.q_row_offset + qi) * a.mask_rstride;
    float qreg[8];
    {
        const bf16x8 qv = *reinterpret_cast<const bf16x8*>(a.q + qrow * ldq + h * 256 + lane * 8);
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
            kv[j] = *reinterpret_cast<const bf16x8*>(kc + static_cast<size_t>(k) * 256 + lane * 8);
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
            vv[j] = *reinterpret_cast<const bf16x8*>(vc + static_cast<size_t>(k) * 256 + lane * 8);
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

cudaError_t launch_joint_attention_fewq(cudaStream_t stream, const JointAttnArgs& a) {
    dim3 grid(a.n_heads, a.q_per_sample, a.batch);
    const size_t smem = (((a.n_keys + 3) & ~3) + 8 * 256) * sizeof(float);
    return launch_kernel(attn_fewq_kernel, grid, dim3(256), smem, stream, a);
}

}  // namespace blurr
