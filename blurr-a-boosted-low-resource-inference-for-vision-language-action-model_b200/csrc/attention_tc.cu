// Gemma joint-attention prefill and SigLIP self-attention on the tcgen05 tensor cores (joint_model.py:246-288,
// siglip.py:133-152), at every batch size since round 2 (the mma.sync kernel of attention.cu keeps the few-query
// launches of the experts and stays selectable with option attn_tc = 0).  At 64 episodes the mma.sync kernel is
// HMMA-bound (41 TFLOP/s, 15 % of the step).
//
// One CTA = 128 (head, query) pairs of one sample.  All 8 query heads share the sample's single K/V head
// (MQA), so the 2208 pairs of a sample form 18 row tiles that each read K and V exactly once.
//   S = Q K^T   : UMMA M=128 x N=288 keys (two N=144 chunks) x K=256, Q and K as K-major SW128 tiles
//                 (K by TMA straight from the cache, Q copied + swizzled by the CTA because a tile
//                 straddles two heads), fp32 accumulators in 288 TMEM columns;
//   softmax     : thread = row (TMEM lane), two warps per lane quarter split the columns; the reference's
//                 rounding chain per logit (bf16 -> /16 -> *(1/50) -> tanh -> *50 -> +mask, bf16 after every
//                 op), exact two-pass fp32 softmax over the bf16 logits, probabilities rounded to bf16 and
//                 written as the K-major SW128 A operand of the second GEMM (over the dead Q / K tiles);
//   O = P V     : UMMA M=128 x N=256 dims x K=288 keys; V is used where it lies in the cache,
//                 [key][dim] = MN-major B operand (instruction-descriptor bit 16, canonical SW128 layout
//                 ((8,n),(8,k)):((1,LBO),(8,SBO)) with LBO = one 64-dim box, SBO = 8 keys x 128 B);
//                 its TMA loads overlap the softmax.  O reuses the TMEM columns of S.
#include "bodies.cuh"
#include "gemm_tc.h"
#include "launch.cuh"

namespace blurr {

// Set (to the stage number) when a bounded barrier wait expired; read by attn_take_timeout_flag().
static __device__ int g_attn_timeout_flag = 0;

// Per-CTA timeline for tuning (global option "attn_cta_trace" = device pointer to [n_cta][8] u64, 0 = off): %globaltimer at
// 0 entry, 1 softmax pass 2 (sum of exp) done, 2 Q staged, 3 S complete, 4 probabilities written, 5 softmax pass 1 (rounding
// chain, row max) done, 6 O complete, 7 stored.
static __device__ unsigned long long* g_attn_cta_trace = nullptr;
__device__ __forceinline__ void attn_stamp(int slot) {
    unsigned long long* t = g_attn_cta_trace;
    if (t != nullptr && threadIdx.x == 0) {
        const size_t cta = blockIdx.x + static_cast<size_t>(gridDim.x) * (blockIdx.y + static_cast<size_t>(gridDim.y) * blockIdx.z);
        t[cta * 8 + slot] = globaltimer_ns();
    }
}
int attn_set_cta_trace(void* dev_ptr) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
    return cudaMemcpyToSymbol(g_attn_cta_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1;
}

static constexpr int kTcRows = 128;                  // (head, query) pairs per CTA
static constexpr int kTcKeys = 288;                  // key columns of S (two UMMA N = 144 chunks)
static constexpr int kTcKeyBlocks = 5;               // 64-key blocks of P (320 columns, zero past n_keys)
static constexpr int kTcHd = 256;
static constexpr int kTcThreads = 384;                // Gemma kernel: 3 softmax warps per TMEM lane quarter (96 key columns each)
static constexpr int kTcParts = kTcThreads / 128;
static constexpr int kSgThreads = 512;                // SigLIP kernel: 4 per quarter (64 key columns each)
static constexpr int kSgParts = kSgThreads / 128;
static constexpr int kTcQBytes = kTcRows * kTcHd * 2;                  // 64 KB: 4 k-blocks of [128][64]
static constexpr int kTcKBytes = kTcKeys * kTcHd * 2;                  // 144 KB: 4 k-blocks of 2 x [144][64]
static constexpr int kTcPBytes = kTcRows * kTcKeyBlocks * 64 * 2;      // 80 KB: 5 key blocks of [128][64]
static constexpr int kTcVBytes = kTcKeys * kTcHd * 2;                  // 144 KB: 4.5 key blocks x 4 dim groups
static constexpr int kTcSmem = kTcPBytes + kTcVBytes + 1024 /*alignment*/ + kTcParts * 512 /*row stats*/ + 64 /*barriers*/;
static_assert(kTcSmem <= 227 * 1024, "shared memory budget");
static_assert(kTcQBytes + kTcKBytes <= kTcPBytes + kTcVBytes, "phase 1 operands must fit the phase 2 footprint");

struct AttnTcArgs {
    const bf16* q;                // [B*q_per_sample][n_heads*256]
    bf16* out;                    // same shape
    const bf16* mask; long long mask_bstride, mask_rstride; int q_row_offset;
    int q_per_sample, n_heads, n_keys, n_slots;
    unsigned long long* trace;
};

// 16-byte chunk `cidx` (8 consecutive keys) of row `row` in the P operand: 64-key blocks of [128 rows][128 B],
// 128B-swizzled.  A warp's 32 rows x one chunk is conflict-free (8 rows cover all 32 banks).
__device__ __forceinline__ uint4* p_chunk(uint8_t* p_s, int row, int cidx) {
    return reinterpret_cast<uint4*>(p_s + (cidx >> 3) * (kTcRows * 128) + row * 128 + (((cidx & 7) ^ (row & 7)) << 4));
}
__device__ __forceinline__ float chunk_exp_sum(const uint4 c, float m) {
    const float2 a0 = unpack_bf16x2(c.x), a1 = unpack_bf16x2(c.y), a2 = unpack_bf16x2(c.z), a3 = unpack_bf16x2(c.w);
    return ((exp_fast_f32(a0.x - m) + exp_fast_f32(a0.y - m)) + (exp_fast_f32(a1.x - m) + exp_fast_f32(a1.y - m))) +
           ((exp_fast_f32(a2.x - m) + exp_fast_f32(a2.y - m)) + (exp_fast_f32(a3.x - m) + exp_fast_f32(a3.y - m)));
}
__device__ __forceinline__ uint4 chunk_probs(const uint4 c, float m, float inv_sum) {
    const float2 a0 = unpack_bf16x2(c.x), a1 = unpack_bf16x2(c.y), a2 = unpack_bf16x2(c.z), a3 = unpack_bf16x2(c.w);
    uint4 o;
    o.x = pack_bf16x2(exp_fast_f32(a0.x - m) * inv_sum, exp_fast_f32(a0.y - m) * inv_sum);
    o.y = pack_bf16x2(exp_fast_f32(a1.x - m) * inv_sum, exp_fast_f32(a1.y - m) * inv_sum);
    o.z = pack_bf16x2(exp_fast_f32(a2.x - m) * inv_sum, exp_fast_f32(a2.y - m) * inv_sum);
    o.w = pack_bf16x2(exp_fast_f32(a3.x - m) * inv_sum, exp_fast_f32(a3.y - m) * inv_sum);
    return o;
}

// MN-major SW128 operand: 64 elements (128 B) contiguous along N per row, rows = K index
__device__ __forceinline__ uint64_t make_smem_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;      // next 64-element group along N
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // next 8 rows along K
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(kTcThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v64,
               const __grid_constant__ CUtensorMap tmap_v32, const AttnTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(a.trace, 0);
    attn_stamp(0);
    uint8_t* base = smem_align_1024(smem_raw);
    uint8_t* q_s = base;                              // phase 1: Q | K
    uint8_t* k_s = base + kTcQBytes;
    uint8_t* p_s = base;                              // phase 2: P | V
    uint8_t* v_s = base + kTcPBytes;
    float* stat = reinterpret_cast<float*>(base + kTcPBytes + kTcVBytes);         // [kTcParts][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kTcPBytes + kTcVBytes + kTcParts * 512);
    uint64_t *bar_k = bars, *bar_s = bars + 1, *bar_v = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int n_pairs = a.n_heads * a.q_per_sample;
    const int ldq = a.n_heads * kTcHd;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v64);
        tma_prefetch_desc(&tmap_v32);
        mbar_init(bar_k, 1); mbar_init(bar_s, 1); mbar_init(bar_v, 1); mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    trace_stamp(a.trace, 1);

    // ---- K tiles by TMA: k-block kb, chunk c -> [144 keys][64 dims] ----
    const int key_row0 = b * a.n_slots;
    if (warp == 0 && elect_one_sync()) {       // issue from a converged warp under elect.sync (see gemm_body.cuh)
        mbar_arrive_expect_tx(bar_k, static_cast<uint32_t>(kTcKBytes));
        for (int kb = 0; kb < 4; ++kb)
            for (int c = 0; c < 2; ++c)
                tma_load_2d(k_s + (kb * 2 + c) * (144 * 128), &tmap_k, bar_k, kb * 64, key_row0 + c * 144);
    }
    // ---- Q tile: rows are (head, query) pairs; copy + 128B swizzle, zero rows past the last pair ----
    // cp.async (zero-fill for the padding rows): all of a thread's ~11 chunks are in flight at once.  Through registers
    // (ld.global -> st.shared in a loop) every store waited for its own load: 4.1 us for this tile (per-CTA timeline).
    for (int idx = threadIdx.x; idx < kTcRows * 32; idx += kTcThreads) {
        const int r = idx >> 5, ch = idx & 31;               // 32 chunks of 8 dims per row
        const int p = tile * kTcRows + r;
        const bool valid = p < n_pairs;
        const bf16* src = a.q;
        if (valid) {
            const int head = p / a.q_per_sample, qi = p - head * a.q_per_sample;
            src = a.q + (static_cast<size_t>(b) * a.q_per_sample + qi) * ldq + head * kTcHd + ch * 8;
        }
        const int kb = ch >> 3, c8 = ch & 7;
        cp_async_16(q_s + kb * (kTcRows * 128) + r * 128 + ((c8 ^ (r & 7)) << 4), src, valid);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();          // generic-proxy writes of Q before the tensor core reads them
    __syncthreads();
    attn_stamp(2);

    // ---- S = Q K^T ----
    if (warp == 1) {
        if (!mbar_wait_warp(bar_k, 0) && lane == 0) atomicExch(&g_attn_timeout_flag, 1);
        tcgen05_fence_after();
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_bf16(kTcRows, 144);
            for (int kb = 0; kb < 4; ++kb) {
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(q_s + kb * (kTcRows * 128)));
                for (int c = 0; c < 2; ++c) {
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(k_s + (kb * 2 + c) * (144 * 128)));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem + c * 144, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
            }
            umma_commit(bar_s);
        }
        __syncwarp();
    }

    // ---- softmax: kTcParts warps per TMEM lane quarter, each a third of the key columns ----
    constexpr int kCols = kTcKeys / kTcParts;            // 96
    const int quarter = warp & 3, half = warp >> 2;      // `half` = column part 0..kTcParts-1
    const int row = quarter * 32 + lane;
    const int pair = tile * kTcRows + row;
    const bool row_valid = pair < n_pairs;
    const int head = row_valid ? pair / a.q_per_sample : 0;
    const int qi = row_valid ? pair - head * a.q_per_sample : 0;
    const bf16* mrow = a.mask + static_cast<size_t>(b) * a.mask_bstride + static_cast<size_t>(a.q_row_offset + qi) * a.mask_rstride;
    if (!mbar_wait(bar_s, 0)) { if (lane == 0) atomicExch(&g_attn_timeout_flag, 2); }
    tcgen05_fence_after();
    __syncthreads();                   // every thread has seen S complete: Q and K tiles are dead
    attn_stamp(3);
    if (warp == 0 && elect_one_sync()) {
        // V where it lies: key block kb (64 keys; the last one 32), dim group j -> [keys][64 dims]
        mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(kTcVBytes));
        for (int kb = 0; kb < 4; ++kb)
            for (int j = 0; j < 4; ++j)
                tma_load_2d(v_s + (kb * 4 + j) * 8192, &tmap_v64, bar_v, j * 64, key_row0 + kb * 64);
        for (int j = 0; j < 4; ++j)
            tma_load_2d(v_s + 16 * 8192 + j * 4096, &tmap_v32, bar_v, j * 64, key_row0 + 256);
    }
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + half * kCols;
    const int col0 = half * kCols;
    float m = -INFINITY;
    // few-query tiles (1 or 4 queries per sample: 8 or 32 valid rows of 128): warps whose 32 rows are all padding skip
    // the softmax arithmetic (their P rows stay whatever the dead Q / K tiles left there; the O rows they feed are never stored)
    const bool warp_live = tile * kTcRows + quarter * 32 < n_pairs;
    // 16-byte mask loads when the mask rows allow it (the engine stages them with a stride of 280): one thread
    // owns one mask row, so scalar loads touch 32 sectors per warp request, 16 requests per 16 columns
    const bool mask_vec = ((a.mask_rstride | a.mask_bstride) & 7) == 0 && (reinterpret_cast<uintptr_t>(a.mask) & 15) == 0;
    // pass 1: rounding chain + mask -> bf16 logits into the P tile, running max
    for (int g = 0; warp_live && g < kCols / 16; ++g) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_addr + g * 16, r);
        uint32_t mk[8];
        if (mask_vec && row_valid && col0 + g * 16 + 16 <= ((a.n_keys + 7) & ~7)) {
            const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(mrow + col0 + g * 16));
            const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(mrow + col0 + g * 16 + 8));
            mk[0] = m0.x; mk[1] = m0.y; mk[2] = m0.z; mk[3] = m0.w; mk[4] = m1.x; mk[5] = m1.y; mk[6] = m1.z; mk[7] = m1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int col = col0 + g * 16 + 2 * i;
                const float lo = (row_valid && col < a.n_keys) ? bf2f(mrow[col]) : 0.f;
                const float hi = (row_valid && col + 1 < a.n_keys) ? bf2f(mrow[col + 1]) : 0.f;
                mk[i] = pack_bf16x2(lo, hi);
            }
        }
        tmem_ld_wait();
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            // two logits at a time: every rounding point of the chain is one packed conversion for the pair
            const int col = col0 + g * 16 + i;
            float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]);
            bf16_round2(v0, v1);
            v0 *= 0.0625f; v1 *= 0.0625f;                          // / sqrt(256): exact
            v0 *= (1.0f / 50.0f); v1 *= (1.0f / 50.0f);
            bf16_round2(v0, v1);
            v0 = tanh_fast_f32(v0); v1 = tanh_fast_f32(v1);
            bf16_round2(v0, v1);
            v0 *= 50.0f; v1 *= 50.0f;
            bf16_round2(v0, v1);
            const float2 mv = unpack_bf16x2(mk[i >> 1]);
            v0 += mv.x; v1 += mv.y;
            bf16_round2(v0, v1);
            if (!(col < a.n_keys && row_valid)) v0 = -INFINITY;
            if (!(col + 1 < a.n_keys && row_valid)) v1 = -INFINITY;
            m = fmaxf(m, fmaxf(v0, v1));
            packed[i >> 1] = pack_bf16x2(v0, v1);
        }
        const int cidx = (col0 + g * 16) >> 3;
        *p_chunk(p_s, row, cidx) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *p_chunk(p_s, row, cidx + 1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    }
    stat[half * 128 + row] = m;
    __syncthreads();
    attn_stamp(5);      // pass 1 (rounding chain, logits, row max) done
    m = stat[row];
#pragma unroll
    for (int pt = 1; pt < kTcParts; ++pt) m = fmaxf(m, stat[pt * 128 + row]);
    if (!row_valid) m = 0.f;
    __syncthreads();
    // pass 2: sum of exp over this thread's columns
    float sum = 0.f;
    for (int c = 0; warp_live && c < kCols / 8; ++c) sum += chunk_exp_sum(*p_chunk(p_s, row, (col0 >> 3) + c), m);
    stat[half * 128 + row] = sum;
    __syncthreads();
    attn_stamp(1);      // pass 2 (sum of exp) done
    sum = stat[row];
#pragma unroll
    for (int pt = 1; pt < kTcParts; ++pt) sum += stat[pt * 128 + row];
    const float inv_sum = row_valid ? 1.f / sum : 0.f;
    // pass 3: probabilities, bf16, in place
    for (int c = 0; warp_live && c < kCols / 8; ++c) {
        uint4* ptr = p_chunk(p_s, row, (col0 >> 3) + c);
        *ptr = row_valid ? chunk_probs(*ptr, m, inv_sum) : make_uint4(0u, 0u, 0u, 0u);
    }
    // key columns 288..319 of the last P block are never multiplied (the K loop stops at 288)
    tcgen05_fence_before();
    fence_proxy_async_smem();          // P (generic proxy) before the tensor core reads it
    __syncthreads();
    attn_stamp(4);

    // ---- O = P V ----
    if (warp == 1) {
        if (!mbar_wait_warp(bar_v, 0) && lane == 0) atomicExch(&g_attn_timeout_flag, 3);
        tcgen05_fence_after();
        if (elect_one_sync()) {
            // M = 128, N = 256, B operand MN-major (bit 16)
            const uint32_t idesc = make_idesc_bf16(kTcRows, kTcHd) | (1u << 16);
            for (int kb = 0; kb < kTcKeyBlocks; ++kb) {
                const int ksteps = (kb < 4) ? 4 : 2;                              // keys 256..287 only
                const uint32_t vbase = smem_u32(v_s + kb * 4 * 8192);
                const uint32_t lbo = (kb < 4) ? 8192u : 4096u;                   // one [keys][64 dims] box
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(p_s + kb * (kTcRows * 128)));
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t b_desc = make_smem_desc_sw128_mn(vbase + k * 2048, lbo);    // 16 keys = 2 x 8 rows
                    umma_bf16_ss(tmem, a_desc + 2 * k, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
            }
            umma_commit(bar_o);
        }
        __syncwarp();
    }

    // ---- epilogue: O -> bf16 -> [b*q + query][head*256 + dim] ----
    if (!mbar_wait(bar_o, 0)) { if (lane == 0) atomicExch(&g_attn_timeout_flag, 4); }
    tcgen05_fence_after();
    attn_stamp(6);
    if (half < 2) {
        const uint32_t oaddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + half * 128;
        bf16* orow = a.out + (static_cast<size_t>(b) * a.q_per_sample + qi) * ldq + head * kTcHd + half * 128;
        for (int g = 0; g < 8; ++g) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(oaddr + g * 16, r);
            tmem_ld_wait();
            if (row_valid) {
                uint4 lo, hi;
                lo.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
                lo.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
                lo.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
                lo.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
                hi.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));
                hi.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
                hi.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13]));
                hi.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
                *reinterpret_cast<uint4*>(orow + g * 16) = lo;
                *reinterpret_cast<uint4*>(orow + g * 16 + 8) = hi;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    trace_stamp(a.trace, 2);
    attn_stamp(7);
    if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// SigLIP self-attention (siglip.py:133-152) for batched episodes: one CTA = 128 queries of one head.
// head_dim 72 is padded to 128 for the tensor cores without touching memory layouts: the Q tile is copied
// with zeros in dims 72..127 (so whatever the K tile's second 64-dim TMA box drags in from the next head
// multiplies by zero), V's second 64-dim box only feeds output columns 72..127, which are never stored.
//   S = Q K^T : M=128 x N=256 keys x K=128 -> TMEM columns 0..255;  O = P V : M=128 x N=128 x K=256 -> 256..383.
struct AttnTcSiglipArgs {
    const bf16* qkv; int ld_qkv; int seq, n_heads, hidden; bf16* out; int ld_out; float scale;
    unsigned long long* trace;
};
static constexpr int kSgQBytes = 128 * 128 * 2;           // 32 KB: 2 k-blocks of [128][64]
static constexpr int kSgKBytes = 256 * 128 * 2;           // 64 KB: 2 k-blocks of [256 keys][64]
static constexpr int kSgPBytes = 128 * 256 * 2;           // 64 KB: 4 key blocks of [128][64], over Q | K
static constexpr int kSgVBytes = 256 * 128 * 2;           // 64 KB: 4 key blocks x 2 dim boxes of [64 keys][64]
static constexpr int kSgSmem = kSgQBytes + kSgKBytes + kSgVBytes + 1024 + kSgParts * 512 + 64;

__global__ void __launch_bounds__(kSgThreads, 1)
attn_tc_siglip_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const AttnTcSiglipArgs a) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(a.trace, 0);
    uint8_t* base = smem_align_1024(smem_raw);
    uint8_t* q_s = base;
    uint8_t* k_s = base + kSgQBytes;
    uint8_t* p_s = base;                                   // P overwrites Q and the first half of K
    uint8_t* v_s = base + kSgQBytes + kSgKBytes;
    float* stat = reinterpret_cast<float*>(v_s + kSgVBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + kSgVBytes + kSgParts * 512);
    uint64_t *bar_k = bars, *bar_s = bars + 1, *bar_v = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int hd = a.hidden / a.n_heads;                   // 72

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        mbar_init(bar_k, 1); mbar_init(bar_s, 1); mbar_init(bar_v, 1); mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();
    pdl_trigger();
    trace_stamp(a.trace, 1);

    const int row0 = b * a.seq;
    if (warp == 0 && elect_one_sync()) {
        mbar_arrive_expect_tx(bar_k, static_cast<uint32_t>(kSgKBytes));
        for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(k_s + kb * (256 * 128), &tmap_k, bar_k, a.hidden + h * hd + kb * 64, row0);
        mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(kSgVBytes));
        for (int kb = 0; kb < 4; ++kb)
            for (int j = 0; j < 2; ++j)
                tma_load_2d(v_s + (kb * 2 + j) * 8192, &tmap_v, bar_v, 2 * a.hidden + h * hd + j * 64, row0 + kb * 64);
    }
    // Q tile: 128 queries x 128 dims (dims >= 72 zero), swizzled K-major
    for (int idx = threadIdx.x; idx < kTcRows * 16; idx += kSgThreads) {
        const int r = idx >> 4, ch = idx & 15;
        const int qi = tile * kTcRows + r;
        const bool valid = qi < a.seq && ch * 8 < hd;
        const bf16* src = valid ? a.qkv + static_cast<size_t>(row0 + qi) * a.ld_qkv + h * hd + ch * 8 : a.qkv;
        const int kb = ch >> 3, c8 = ch & 7;
        cp_async_16(q_s + kb * (kTcRows * 128) + r * 128 + ((c8 ^ (r & 7)) << 4), src, valid);      // zero-fill past dim 72 / past the last query
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncthreads();

    if (warp == 1) {
        if (!mbar_wait_warp(bar_k, 0) && lane == 0) atomicExch(&g_attn_timeout_flag, 5);
        tcgen05_fence_after();
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_bf16(kTcRows, 256);
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(q_s + kb * (kTcRows * 128)));
                const uint64_t b_desc = make_smem_desc_sw128(smem_u32(k_s + kb * (256 * 128)));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_s);
        }
        __syncwarp();
    }

    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int qi = tile * kTcRows + row;
    const bool row_valid = qi < a.seq;
    if (!mbar_wait(bar_s, 0)) { if (lane == 0) atomicExch(&g_attn_timeout_flag, 6); }
    tcgen05_fence_after();
    __syncthreads();                   // S complete for everyone: Q / K tiles are dead
    constexpr int kCols = 256 / kSgParts;                 // 64
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + half * kCols;
    const int col0 = half * kCols;
    float m = -INFINITY;
    for (int g = 0; g < kCols / 16; ++g) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_addr + g * 16, r);
        tmem_ld_wait();
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float s[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = col0 + g * 16 + i + e;
                float v = bf16_round(bf16_round(__uint_as_float(r[i + e])) * a.scale);
                if (col >= a.seq || !row_valid) v = -INFINITY;
                s[e] = v;
                m = fmaxf(m, v);
            }
            packed[i >> 1] = pack_bf16x2(s[0], s[1]);
        }
        const int cidx = (col0 + g * 16) >> 3;
        *p_chunk(p_s, row, cidx) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *p_chunk(p_s, row, cidx + 1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    }
    stat[half * 128 + row] = m;
    __syncthreads();
    m = stat[row];
#pragma unroll
    for (int pt = 1; pt < kSgParts; ++pt) m = fmaxf(m, stat[pt * 128 + row]);
    if (!row_valid) m = 0.f;
    __syncthreads();
    float sum = 0.f;
    for (int c = 0; c < kCols / 8; ++c) sum += chunk_exp_sum(*p_chunk(p_s, row, (col0 >> 3) + c), m);
    stat[half * 128 + row] = sum;
    __syncthreads();
    sum = stat[row];
#pragma unroll
    for (int pt = 1; pt < kSgParts; ++pt) sum += stat[pt * 128 + row];
    const float inv_sum = row_valid ? 1.f / sum : 0.f;
    for (int c = 0; c < kCols / 8; ++c) {
        uint4* ptr = p_chunk(p_s, row, (col0 >> 3) + c);
        *ptr = row_valid ? chunk_probs(*ptr, m, inv_sum) : make_uint4(0u, 0u, 0u, 0u);
    }
    tcgen05_fence_before();
    fence_proxy_async_smem();
    __syncthreads();

    if (warp == 1) {
        if (!mbar_wait_warp(bar_v, 0) && lane == 0) atomicExch(&g_attn_timeout_flag, 7);
        tcgen05_fence_after();
        if (elect_one_sync()) {
            const uint32_t idesc = make_idesc_bf16(kTcRows, 128) | (1u << 16);          // B (= V) MN-major
            for (int kb = 0; kb < 4; ++kb) {
                const uint32_t vbase = smem_u32(v_s + kb * 2 * 8192);
                const uint64_t a_desc = make_smem_desc_sw128(smem_u32(p_s + kb * (kTcRows * 128)));
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem + 256, a_desc + 2 * k, make_smem_desc_sw128_mn(vbase + k * 2048, 8192u), idesc,
                                 (kb > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(bar_o);
        }
        __syncwarp();
    }

    if (!mbar_wait(bar_o, 0)) { if (lane == 0) atomicExch(&g_attn_timeout_flag, 8); }
    tcgen05_fence_after();
    if (half == 0) {
        // 72 real output dims: four groups of 16 and the first 8 of a fifth
        const uint32_t oaddr = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + 256;
        bf16* orow = a.out + static_cast<size_t>(row0 + qi) * a.ld_out + h * hd;
        for (int g = 0; g < 5; ++g) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(oaddr + g * 16, r);
            tmem_ld_wait();
            if (row_valid) {
                uint4 lo, hi;
                lo.x = pack_bf16x2(__uint_as_float(r[0]), __uint_as_float(r[1]));
                lo.y = pack_bf16x2(__uint_as_float(r[2]), __uint_as_float(r[3]));
                lo.z = pack_bf16x2(__uint_as_float(r[4]), __uint_as_float(r[5]));
                lo.w = pack_bf16x2(__uint_as_float(r[6]), __uint_as_float(r[7]));
                hi.x = pack_bf16x2(__uint_as_float(r[8]), __uint_as_float(r[9]));
                hi.y = pack_bf16x2(__uint_as_float(r[10]), __uint_as_float(r[11]));
                hi.z = pack_bf16x2(__uint_as_float(r[12]), __uint_as_float(r[13]));
                hi.w = pack_bf16x2(__uint_as_float(r[14]), __uint_as_float(r[15]));
                if (g * 16 < hd) *reinterpret_cast<uint4*>(orow + g * 16) = lo;
                if (g * 16 + 8 < hd) *reinterpret_cast<uint4*>(orow + g * 16 + 8) = hi;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    trace_stamp(a.trace, 2);
    if (warp == 2) tmem_dealloc(tmem, 512);
}

const int* attn_timeout_flag_ptr() {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_attn_timeout_flag) == cudaSuccess ? static_cast<const int*>(p) : nullptr;
}

int attn_take_timeout_flag() {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_attn_timeout_flag, sizeof(int)) != cudaSuccess) return -1;
    if (v != 0) {
        const int zero = 0;
        cudaMemcpyToSymbol(g_attn_timeout_flag, &zero, sizeof(int));
    }
    return v;
}

// -1 automatic = whenever the shape allows, at any batch size (round 2: with cp.async Q staging, shared-space smem accesses
// and the lean MUFU math the tcgen05 kernels are level with the mma.sync tile kernel at batch 1 - prefill stage 1.83 vs
// 1.89 ms, SigLIP stage 1.35 vs 1.40 ms - and 3-4x faster for batched episodes); 0 = mma.sync kernels; 1 = same as -1
static int g_attn_tc = -1;
void attn_set_tc(int mode) { g_attn_tc = mode; }
// Few-query attention (proprio: 1 query per sample, action: 4) on the tcgen05 kernel: the (head, query) pairs of a sample
// are one 128-row tile.  Measured slower than the mma.sync tile kernel at batch 1 (per-CTA timeline, tools/attn_timeline.py:
// 21 us against 18 us): with 8 or 32 live rows the softmax runs on three warps (a warp can only read its own TMEM lane
// quarter), one row per thread, 96 serial columns each - 15 us of dependent ALU / MUFU latency.  Option only: 1 = on.
static int g_attn_tc_fewq = 0;
void attn_set_tc_fewq(int mode) { g_attn_tc_fewq = mode; }
bool attn_tc_fewq_applies(const JointAttnArgs& j) {
    if (g_attn_tc_fewq <= 0) return false;
    return j.n_keys <= kTcKeys && j.n_slots >= 1 && j.q_per_sample >= 1 && j.n_heads * j.q_per_sample <= kTcRows;
}

bool attn_tc_applies(const JointAttnArgs& j) {
    if (g_attn_tc == 0) return false;
    const bool shape_ok = j.n_keys <= kTcKeys && j.n_slots >= 1 && j.q_per_sample >= 1 && j.n_heads >= 1;
    if (!shape_ok) return false;
    return true;
}

bool attn_tc_siglip_applies(int batch, int seq, int n_heads, int hidden) {
    if (g_attn_tc == 0) return false;
    const int hd = hidden / n_heads;
    if (seq > 256 || hd > 128 || (hd & 7) != 0) return false;
    return true;
}

cudaError_t launch_siglip_attention_tc(cudaStream_t stream, const bf16* qkv, int ld_qkv, int batch, int seq, int n_heads,
                                       int hidden, bf16* out, int ld_out, unsigned long long* trace) {
    CUtensorMap tk, tv;
    std::string err;
    if (gemm_get_tensor_map(qkv, batch * seq, 3 * hidden, ld_qkv, 256, &tk, &err)) return cudaErrorInvalidValue;
    if (gemm_get_tensor_map(qkv, batch * seq, 3 * hidden, ld_qkv, 64, &tv, &err)) return cudaErrorInvalidValue;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_siglip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    AttnTcSiglipArgs a{};
    a.qkv = qkv; a.ld_qkv = ld_qkv; a.seq = seq; a.n_heads = n_heads; a.hidden = hidden; a.out = out; a.ld_out = ld_out;
    a.scale = static_cast<float>(pow(static_cast<double>(hidden / n_heads), -0.5));
    a.trace = trace;
    return launch_kernel(attn_tc_siglip_kernel, dim3((seq + kTcRows - 1) / kTcRows, n_heads, batch), dim3(kSgThreads),
                         static_cast<size_t>(kSgSmem), stream, tk, tv, a);
}

cudaError_t launch_joint_attention_prefill_tc(cudaStream_t stream, const JointAttnArgs& j, std::string* err) {
    CUtensorMap tk, tv64, tv32;
    const int rows = j.batch * j.n_slots;
    if (gemm_get_tensor_map(j.k_cache, rows, kTcHd, kTcHd, 144, &tk, err)) return cudaErrorInvalidValue;
    if (gemm_get_tensor_map(j.v_cache, rows, kTcHd, kTcHd, 64, &tv64, err)) return cudaErrorInvalidValue;
    if (gemm_get_tensor_map(j.v_cache, rows, kTcHd, kTcHd, 32, &tv32, err)) return cudaErrorInvalidValue;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    AttnTcArgs a{};
    a.q = j.q; a.out = j.out; a.mask = j.mask; a.mask_bstride = j.mask_bstride; a.mask_rstride = j.mask_rstride;
    a.q_row_offset = j.q_row_offset; a.q_per_sample = j.q_per_sample; a.n_heads = j.n_heads; a.n_keys = j.n_keys;
    a.n_slots = j.n_slots; a.trace = j.trace;
    const int tiles = (j.n_heads * j.q_per_sample + kTcRows - 1) / kTcRows;
    return launch_kernel(attn_tc_kernel, dim3(tiles, j.batch), dim3(kTcThreads), static_cast<size_t>(kTcSmem), stream, tk, tv64,
                         tv32, a);
}

}  // namespace blurr
