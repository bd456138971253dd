// Shared device helpers for the blurr_b200 kernels (sm_100a only).
//
// PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (MMA, TMEM alloc, TMEM
// load) plus the bf16 rounding helpers every kernel uses to reproduce the reference's
// rounding points.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "blurr_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace blurr {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------
// bf16 rounding helpers
// ---------------------------------------------------------------------------
// Round to bf16 and back.  Through the PACKED converter (cvt.rn.bf16x2.f32 = F2FP.BF16.F32.PACK_AB, FMA pipe, full
// rate) and a shift: the scalar cvt.rn.bf16.f32 is an F2F on the XU / MIO path (16 per clock per SM), and the rounding
// chains of the attention softmax and the GEMM epilogues execute several per element (ncu, round 2: F2F.BF16.F32 with
// `mio` stalls all over the tcgen05 attention kernel).  Same result bit for bit (round to nearest even, NaN / Inf kept).
__device__ __forceinline__ float bf16_round(float x) {
    __nv_bfloat162 v = __floats2bfloat162_rn(x, 0.0f);
    return __uint_as_float(*reinterpret_cast<uint32_t*>(&v) << 16);
}
// two at once: one conversion instruction for both
__device__ __forceinline__ void bf16_round2(float& a, float& b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    const uint32_t u = *reinterpret_cast<uint32_t*>(&v);
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
}
__device__ __forceinline__ float bf2f(bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 f2bf(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

// MUFU primitives without the denormal handling nvcc wraps around __expf / division (a compare, a predicated rescale and
// a reconvergence point per call: the softmax of the tcgen05 attention kernel spent ~65 instructions per logit, half of
// them there).  .ftz: results below 2^-126 flush to zero - irrelevant for exp(x - max) and for exp(2x) with |x| <= 15.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp(x) as __expf computes it (ex2.approx of x * log2(e)), minus the denormal path
__device__ __forceinline__ float exp_fast_f32(float x) { return ex2_ftz(x * 1.4426950408889634f); }

// tanh through one ex2 and one rcp (relative error ~2^-21 against tanhf's ~2^-23).  Every use rounds the
// result to bf16 right away, so against tanhf it flips about one rounding in 2^12 - far inside the parity
// tolerance - at a third of the instructions (the GELU / GeGLU epilogues and the batched softmax are ALU-bound).
// Same bits as 1 - __fdividef(2, __expf(2x) + 1): the factors of two are exact.
__device__ __forceinline__ float tanh_fast_f32(float x) {
    x = fminf(fmaxf(x, -15.f), 15.f);
    const float e = ex2_ftz(x * 2.8853900817779268f);      // exp(2x)
    return fmaf(-2.f, rcp_ftz(e + 1.f), 1.f);
}

// torch.nn.functional.gelu(x, approximate="tanh") on a bf16 tensor: computed in fp32 from the
// bf16 input, result rounded to bf16 by the caller (ATen GeluCUDAKernelImpl).
__device__ __forceinline__ float gelu_tanh_f32(float x) {
    const float kBeta = 0.7978845608028654f;   // sqrt(2/pi)  (M_SQRT2 * M_2_SQRTPI * 0.5)
    const float kKappa = 0.044715f;
    float x_cube = x * x * x;
    float inner = kBeta * (x + kKappa * x_cube);
    return 0.5f * x * (1.0f + tanh_fast_f32(inner));
}

__device__ __forceinline__ float silu_f32(float x) { return x / (1.0f + expf(-x)); }
// exact GELU (nn.GELU() default); the tanh flavour is gelu_tanh_f32.  Chosen at compile time (EPI_GELU / EPI_GELU_ERF).
__device__ __forceinline__ float gelu_erf_f32(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// gate activation of the GLU epilogue: 0 = tanh GELU (Gemma GeGLU), 1 = SiLU (Llama SwiGLU)
__device__ __forceinline__ float glu_act_f32(float x, int act) { return act == 1 ? silu_f32(x) : gelu_tanh_f32(x); }

// ---------------------------------------------------------------------------
// shared-memory address / misc
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// First 1024-byte aligned byte of a dynamic shared-memory array, as pointer arithmetic ON the array: the compiler keeps
// the shared address space (STS / LDS).  Rounding the address through uintptr_t loses it and every access to the carved
// buffers becomes a generic ST.E / LD.E (found in the SASS of the GEMM staging stores and the attention softmax, round 2).
__device__ __forceinline__ uint8_t* smem_align_1024(uint8_t* smem_raw) {
    const uint32_t a = smem_u32(smem_raw);
    return smem_raw + ((1024u - (a & 1023u)) & 1023u);
}

// ---------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU.  On expiry (~0.2 s of SM clocks) the
// wait returns false; the caller records the failure in a device flag and unwinds, so the host
// sees an error code instead of a hung or faulted device.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 400000000LL) return false;
    }
    return true;
}

// Same wait executed by a whole converged warp; the verdict is agreed across the lanes so that the code
// after it stays warp-uniform (the issue loops keep all 32 lanes converged, see gemm_body.cuh).
__device__ __forceinline__ bool mbar_wait_warp(uint64_t* bar, uint32_t parity) {
    const bool ok = mbar_wait(bar, parity);
    return __all_sync(0xffffffffu, ok);
}

// ---------------------------------------------------------------------------
// TMA: 2D tiled load global -> shared (SWIZZLE_128B tensor maps), completes on an mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int32_t c_inner, int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// Same, with an L2 cache-policy hint (weights are streamed once: evict_first).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* map,
                                                 uint64_t* bar, int32_t c_inner, int32_t c_outer,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer),
        "l"(policy)
        : "memory");
}
// Multicast variant: the box lands at the same shared-memory offset of every CTA in `cta_mask`
// and completes bytes on the mbarrier at the same offset in each of them.
__device__ __forceinline__ void tma_load_2d_multicast_hint(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                           int32_t c_inner, int32_t c_outer, uint16_t cta_mask,
                                                           uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c_inner), "r"(c_outer),
        "l"(policy)
        : "memory");
}
// --- CTA pairs (cta_group::2): two SMs of a TPC work on one 256-row UMMA tile ----------------
// TMA load issued by either CTA of the pair; the bytes complete on the mbarrier `bar_cluster_addr`,
// a shared::cluster address that may live in the peer (leader) CTA.
__device__ __forceinline__ void tma_load_2d_2sm_hint(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                     int32_t c_inner, int32_t c_outer, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer), "l"(policy)
        : "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) as seen in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]; issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all previously issued MMAs retire) on the mbarrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint64_t make_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t make_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

// ---------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM load
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// Same arrive, delivered to the mbarrier at this offset in every CTA of `cta_mask` (a stage of a
// multicast pipeline is free only when all CTAs of the cluster have consumed it).
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (lane == TMEM lane).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// Same wait, tied to the 16 destination registers of an earlier tcgen05.ld so that no use of them can be scheduled
// above it (software-pipelined drains keep a second load in flight while the first one's values are consumed).
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored as [rows][64] (128 B per
// row) with the 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are
// 1024 B apart (SBO), LBO is 1 (unused for swizzled K-major), descriptor version 1
// (Blackwell), layout type 2 (SWIZZLE_128B).  Field layout: cute::UMMA::SmemDescriptor.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (>>4)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset (>>4)
    d |= static_cast<uint64_t>(1) << 46;                      // version = 1
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major.
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
    uint32_t d = 0;
    d |= 1u << 4;          // c_format = F32
    d |= 1u << 7;          // a_format = BF16
    d |= 1u << 10;         // b_format = BF16
    d |= (n >> 3) << 17;   // n_dim
    d |= (m >> 4) << 24;   // m_dim
    return d;
}

// ---------------------------------------------------------------------------
// legacy warp-level tensor ops used by the small attention tiles (mma.sync + ldmatrix)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
    uint32_t sz = valid ? 16u : 0u;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(sz)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte vector of 8 bf16
struct __align__(16) bf16x8 {
    uint32_t u[4];
};

}  // namespace blurr
