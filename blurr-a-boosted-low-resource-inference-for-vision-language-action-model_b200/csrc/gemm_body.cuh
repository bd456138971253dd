// Device body of the tcgen05 / TMEM / TMA GEMM (see gemm_tc.cu for the design notes), written so
// that the same code serves a stand-alone launch (one tile per CTA) and the persistent step kernel
// (many tiles and many GEMMs per CTA): the shared-memory ring, its mbarriers and the TMEM allocation
// are set up once by the caller, and the barrier phase parities live in a per-thread `GemmPipe`
// that persists across tiles.
#pragma once

#include "common.cuh"
#include "gemm_tc.h"
#include "launch.cuh"

namespace blurr {

static constexpr int kBlockM = 128;     // weight rows per CTA (UMMA M)
static constexpr int kBlockK = 64;      // bf16 elements per k-block (one 128-byte swizzle row)
static constexpr int kTileABytes = kBlockM * kBlockK * 2;
static constexpr int kGemmThreads = 256;
static constexpr int kMaxStages = 12;


struct GemmShared {
    uint8_t* ring;            // 1024-byte aligned pipeline buffers (also the epilogue tile)
    uint64_t* full_bar;       // [kMaxStages]
    uint64_t* empty_bar;      // [kMaxStages]
    uint64_t* tmem_full_bar;  // [1]
    uint32_t tmem_base;
};
// Phase parity to wait for next, one bit per barrier; only the thread that waits on a barrier
// consults its bit, so every thread keeps its own copy.
struct GemmPipe {
    uint32_t full_bits = 0, empty_bits = 0, tmem_bit = 0;
};

// Set (to 1 + role) when a pipeline wait expired; read by gemm_take_timeout_flag().
static __device__ int g_gemm_timeout_flag = 0;

// Per-CTA timeline for kernel tuning (global option "gemm_cta_trace" = device pointer to [n_cta][8] u64, 0 = off):
// %globaltimer at 0 kernel entry, 1 setup done, 2 producer past the dependency wait, 3 first stage landed,
// 4 last MMA issued, 5 accumulators complete (epilogue warps), 6 TMEM drained, 7 tile stored.
static __device__ unsigned long long* g_cta_trace = nullptr;
__device__ __forceinline__ void cta_stamp(int slot) {
    unsigned long long* t = g_cta_trace;
    if (t != nullptr && (threadIdx.x & 31) == 0) {
        const size_t cta = blockIdx.x + static_cast<size_t>(gridDim.x) * (blockIdx.y + static_cast<size_t>(gridDim.y) * blockIdx.z);
        t[cta * 8 + slot] = globaltimer_ns();
    }
}

// Epilogue phase 2: the [token][128] tile staged in shared memory -> global memory with 16-byte
// row-wise stores (all 256 threads).
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_store(const GemmDev& p, uint8_t* smem, const int bx, const int bz,
                                                    const int n0, const int t0, const int ntok_override = 0,
                                                    const int nthreads = kGemmThreads) {
    {
        const int ntok = ntok_override > 0 ? ntok_override : p.nt * p.bn;
        if (EPI == EPI_PARTIAL) {
            const float4* tile = reinterpret_cast<const float4*>(smem);
            float* dst = p.partial + static_cast<size_t>(bz) * p.T * p.Nw;
            for (int idx = threadIdx.x; idx < ntok * 32; idx += nthreads) {
                const int t = idx >> 5, ch = idx & 31;
                if (t0 + t < p.T)
                    *reinterpret_cast<float4*>(dst + static_cast<size_t>(t0 + t) * p.Nw + n0 + ch * 4) =
                        tile[t * 32 + ch];
            }
        } else if (EPI == EPI_GEGLU) {
            // weight rows alternate gate_j, up_j: 16 consecutive tile columns give 8 outputs
            const bf16x8* tile = reinterpret_cast<const bf16x8*>(smem);
            for (int idx = threadIdx.x; idx < ntok * 8; idx += nthreads) {
                const int t = idx >> 3, ch = idx & 7;
                if (t0 + t >= p.T) continue;
                const bf16x8 lo = tile[t * 16 + 2 * ch];
                const bf16x8 hi = tile[t * 16 + 2 * ch + 1];
                bf16x8 o;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 a0 = unpack_bf16x2(lo.u[2 * j]), a1 = unpack_bf16x2(lo.u[2 * j + 1]);
                    const float2 b0 = unpack_bf16x2(hi.u[2 * j]), b1 = unpack_bf16x2(hi.u[2 * j + 1]);
                    o.u[j] = pack_bf16x2(bf16_round(glu_act_f32(a0.x, p.glu_act)) * a0.y, bf16_round(glu_act_f32(a1.x, p.glu_act)) * a1.y);
                    o.u[2 + j] = pack_bf16x2(bf16_round(glu_act_f32(b0.x, p.glu_act)) * b0.y, bf16_round(glu_act_f32(b1.x, p.glu_act)) * b1.y);
                }
                *reinterpret_cast<bf16x8*>(p.out + static_cast<size_t>(t0 + t) * p.ldo + bx * (kBlockM / 2) +
                                           ch * 8) = o;
            }
        } else {
            const bf16x8* tile = reinterpret_cast<const bf16x8*>(smem);
            for (int idx = threadIdx.x; idx < ntok * 16; idx += nthreads) {
                const int t = idx >> 4, ch = idx & 15;
                if (t0 + t < p.T)
                    *reinterpret_cast<bf16x8*>(p.out + static_cast<size_t>(t0 + t) * p.ldo + n0 + ch * 8) =
                        tile[t * 16 + ch];
            }
        }
    }
}

// One 128-row weight tile x (nt x bn tokens) x one split-K slice.  Called by all 256 threads.
template <int EPI>
__device__ __forceinline__ void gemm_tile(const GemmDev& p, const CUtensorMap* tmap_w, const CUtensorMap* tmap_x,
                                          const CUtensorMap* tmap_xs, const GemmShared& sh, GemmPipe& st,
                                          const int bx, const int by, const int bz, const uint32_t crank,
                                          const bool pdl = false) {
    uint8_t* smem = sh.ring;
    const int stage_bytes = kTileABytes + p.nt * p.bn * (kBlockK * 2);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_base = sh.tmem_base;
    const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1u);

    const int n0 = bx * kBlockM;                  // first weight row of this tile
    const int t0 = by * p.nt * p.bn;              // first token row of this tile
    const int kb0 = bz * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nkb = kb1 - kb0;

    // Issue discipline (measured with tools/mma_probe.cu, round 2): the producer and the MMA warp run with all 32
    // lanes converged and warp-uniform operands; only the TMA / tcgen05 instructions sit under elect.sync.  Issued
    // from an `if (lane == 0)` region ptxas wraps EVERY UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY
    // loop (~110-160 cycles per instruction): 908 instead of 724 cycles per k-block at 2 x 144 tokens, 632
    // instead of 448 below 128 tokens, and the TMA side could not issue a k-block in less than ~0.22 us.
    if (warp == 0) {
        const uint64_t pol_w = make_policy_evict_first();   // weights: streamed once
        const uint64_t pol_x = make_policy_evict_last();    // activations: re-read by every CTA
        // Launched with PDL: the weights do not depend on the previous kernel, so the first ring of
        // weight blocks is requested before waiting for it; only the token operand waits.
        int pre = 0;
        if (pdl) {
            pre = (p.cluster == 1 && p.w_static) ? min(p.stages, nkb) : 0;
            if (elect_one_sync()) {
                for (int i = 0; i < pre; ++i) {      // fresh ring: every stage is empty
                    mbar_arrive_expect_tx(&sh.full_bar[i], static_cast<uint32_t>(stage_bytes));
                    if (p.w_packed)
                        tma_load_2d_hint(smem + i * stage_bytes, tmap_w, &sh.full_bar[i], 0,
                                         (bx * p.kb_total + kb0 + i) * kBlockM, pol_w);
                    else
                        tma_load_2d_hint(smem + i * stage_bytes, tmap_w, &sh.full_bar[i], (kb0 + i) * kBlockK, n0, pol_w);
                }
            }
            __syncwarp();
            for (int i = 0; i < pre; ++i) st.empty_bits ^= (1u << i);
            pdl_wait();
            pdl_trigger();
            trace_stamp(p.trace, 1);
            cta_stamp(2);
            if (elect_one_sync()) {
                for (int i = 0; i < pre; ++i)
                    for (int c = 0; c < p.nt; ++c)
                        tma_load_2d_hint(smem + i * stage_bytes + kTileABytes + c * p.bn * (kBlockK * 2), tmap_x,
                                         &sh.full_bar[i], (kb0 + i) * kBlockK, t0 + c * p.bn, pol_x);
            }
            __syncwarp();
        }
        int s = pre % p.stages;
        for (int i = pre; i < nkb; ++i) {
            if (!mbar_wait_warp(&sh.empty_bar[s], ((st.empty_bits >> s) & 1u) ^ 1u)) {
                if (lane == 0) atomicExch(&g_gemm_timeout_flag, 1);
                break;
            }
            st.empty_bits ^= (1u << s);
            if (elect_one_sync()) {
                uint8_t* stg = smem + s * stage_bytes;
                mbar_arrive_expect_tx(&sh.full_bar[s], static_cast<uint32_t>(stage_bytes));
                const int kcoord = (kb0 + i) * kBlockK;
                if (p.w_packed)   // tile (bx, k-block) is a contiguous 128 x 64 block
                    tma_load_2d_hint(stg, tmap_w, &sh.full_bar[s], 0, (bx * p.kb_total + kb0 + i) * kBlockM, pol_w);
                else
                    tma_load_2d_hint(stg, tmap_w, &sh.full_bar[s], kcoord, n0, pol_w);
                if (p.cluster > 1) {
                    // this CTA's slice of the shared activation tile, delivered to every CTA of the cluster
                    const int r0 = static_cast<int>(crank) * p.slice_rows;
                    tma_load_2d_multicast_hint(stg + kTileABytes + r0 * (kBlockK * 2), tmap_xs, &sh.full_bar[s],
                                               kcoord, t0 + r0, cmask, pol_x);
                } else {
                    for (int c = 0; c < p.nt; ++c)
                        tma_load_2d_hint(stg + kTileABytes + c * p.bn * (kBlockK * 2), tmap_x, &sh.full_bar[s],
                                         kcoord, t0 + c * p.bn, pol_x);
                }
            }
            __syncwarp();
            s = (s + 1 == p.stages) ? 0 : s + 1;
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(kBlockM, static_cast<uint32_t>(p.bn));
        bool ok = true;
        int s = 0;
        for (int i = 0; i < nkb; ++i) {
            if (!mbar_wait_warp(&sh.full_bar[s], (st.full_bits >> s) & 1u)) {
                if (lane == 0) atomicExch(&g_gemm_timeout_flag, 2);
                ok = false;
                break;
            }
            st.full_bits ^= (1u << s);
            tcgen05_fence_after();
            if (i == 0) cta_stamp(3);
            if (elect_one_sync()) {
                const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                for (int c = 0; c < p.nt; ++c) {
                    const uint64_t b_desc = make_smem_desc_sw128(a_addr + kTileABytes + c * p.bn * (kBlockK * 2));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the
                        // 16-byte-granular start-address field
                        umma_bf16_ss(tmem_base + c * p.bn, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                     (i > 0 || k > 0) ? 1u : 0u);
                    }
                }
                // frees the smem stage (in every CTA of the cluster) once these MMAs retire
                if (p.cluster > 1) umma_commit_multicast(&sh.empty_bar[s], cmask);
                else umma_commit(&sh.empty_bar[s]);
            }
            __syncwarp();
            s = (s + 1 == p.stages) ? 0 : s + 1;
        }
        if (ok && elect_one_sync()) umma_commit(sh.tmem_full_bar);   // accumulators complete
        __syncwarp();
        cta_stamp(4);
    } else if (warp >= 4) {
        // ---- epilogue phase 1: TMEM -> registers -> smem tile [token][128 n] ----
        const int w4 = warp - 4;               // TMEM lane quarter this warp may access
        const int nl = w4 * 32 + lane;         // local weight row == TMEM lane
        if (pdl) pdl_wait();                   // the output buffers may still be read by the previous kernel
        const bool acc_ready = mbar_wait(sh.tmem_full_bar, st.tmem_bit);
        st.tmem_bit ^= 1u;
        if (!acc_ready && lane == 0) atomicExch(&g_gemm_timeout_flag, 3);
        tcgen05_fence_after();
        if (warp == 4) cta_stamp(5);
        float bias = 0.f;
        if (EPI != EPI_PARTIAL && p.bias != nullptr) bias = bf2f(p.bias[n0 + nl]);
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16);
        const int ntok = p.nt * p.bn;
        for (int g = 0; acc_ready && g < ntok / 16; ++g) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(lane_addr + g * 16, r);
            tmem_ld_wait();
            if (EPI == EPI_PARTIAL) {
                float* tile = reinterpret_cast<float*>(smem);
#pragma unroll
                for (int i = 0; i < 16; ++i) tile[(g * 16 + i) * kBlockM + nl] = __uint_as_float(r[i]);
            } else {
                bf16* tile = reinterpret_cast<bf16*>(smem);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v = bf16_round(__uint_as_float(r[i]) + bias);
                    if (EPI == EPI_GELU) v = gelu_tanh_f32(v);
                    if (EPI == EPI_GELU_ERF) v = gelu_erf_f32(v);
                    tile[(g * 16 + i) * kBlockM + nl] = f2bf(v);
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 0) cta_stamp(6);

    gemm_epilogue_store<EPI>(p, smem, bx, bz, n0, t0);
    __syncthreads();     // the tile aliases the ring: it must be drained before the next tile's TMA writes
    if (warp == 0) cta_stamp(7);
}

// ---------------------------------------------------------------------------------------------
// Persistent variant: the CTA walks tiles `first, first + stride, ...` of the (gx, gy, gz) tile grid.
// The three roles run decoupled across tile boundaries: the TMA producer keeps filling free stages
// with the NEXT tile's operands while the epilogue of the current tile drains TMEM, so HBM stays busy
// through the prologue / epilogue bubbles that a one-tile-per-CTA launch pays once per wave (measured:
// 3.4 TB/s instead of ~6 TB/s on the gate/up shape at any token count).  The epilogue goes TMEM ->
// registers -> global memory directly (no shared-memory tile), so nothing ever aliases the ring.  With
// <= 256 accumulator columns per tile the accumulator is double-buffered in TMEM (p.acc_bufs = 2): the
// MMAs of tile i+1 run while the epilogue warps drain tile i.
template <int EPI>
__device__ __forceinline__ void gemm_persistent(const GemmDev& p, const CUtensorMap* tmap_w, const CUtensorMap* tmap_x,
                                                const GemmShared& sh, uint64_t* tmem_empty_bar, const int gx,
                                                const int gy, const int gz, const int first, const int stride,
                                                const bool pdl = false) {
    uint8_t* smem = sh.ring;
    const int stage_bytes = kTileABytes + p.nt * p.bn * (kBlockK * 2);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_base = sh.tmem_base;
    const int n_tiles = gx * gy * gz;

    if (warp == 0) {
        // all 32 lanes converged, operands warp-uniform, issue under elect.sync (see gemm_tile)
        // few tokens: weights are streamed once, activations re-read by every CTA.  Above 1024 tokens
        // (l2_policy 1) both operands are re-read by later tiles and plain LRU does better.
        const uint64_t pol_w = p.l2_policy == 1 ? make_policy_evict_normal() : make_policy_evict_first();
        const uint64_t pol_x = p.l2_policy == 1 ? make_policy_evict_normal() : make_policy_evict_last();
        uint32_t empty_bits = 0;
        int s = 0;
        int pre = 0;
        if (pdl) {
            // weights of the first ring before the PDL wait (see gemm_tile)
            if (first < n_tiles) {
                const int bx = first % gx, by = (first / gx) % gy, bz = first / (gx * gy);
                const int n0 = bx * kBlockM, t0 = by * p.nt * p.bn;
                const int kb0 = bz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                pre = p.w_static ? min(p.stages, kb1 - kb0) : 0;
                if (elect_one_sync()) {
                    for (int i = 0; i < pre; ++i) {
                        mbar_arrive_expect_tx(&sh.full_bar[i], static_cast<uint32_t>(stage_bytes));
                        if (p.w_packed)
                            tma_load_2d_hint(smem + i * stage_bytes, tmap_w, &sh.full_bar[i], 0,
                                             (bx * p.kb_total + kb0 + i) * kBlockM, pol_w);
                        else
                            tma_load_2d_hint(smem + i * stage_bytes, tmap_w, &sh.full_bar[i], (kb0 + i) * kBlockK, n0, pol_w);
                    }
                }
                __syncwarp();
                for (int i = 0; i < pre; ++i) empty_bits ^= (1u << i);
                pdl_wait();
                pdl_trigger();
                trace_stamp(p.trace, 1);
                if (elect_one_sync()) {
                    for (int i = 0; i < pre; ++i)
                        for (int c = 0; c < p.nt; ++c)
                            tma_load_2d_hint(smem + i * stage_bytes + kTileABytes + c * p.bn * (kBlockK * 2), tmap_x,
                                             &sh.full_bar[i], (kb0 + i) * kBlockK, t0 + c * p.bn, pol_x);
                }
                __syncwarp();
                s = (pre == p.stages) ? 0 : pre;
            } else {
                pdl_wait();
                pdl_trigger();
            }
        }
        for (int tile = first; tile < n_tiles; tile += stride) {
            const int bx = tile % gx, by = (tile / gx) % gy, bz = tile / (gx * gy);
            const int n0 = bx * kBlockM, t0 = by * p.nt * p.bn;
            const int kb0 = bz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            for (int kb = (tile == first ? kb0 + pre : kb0); kb < kb1; ++kb) {
                if (!mbar_wait_warp(&sh.empty_bar[s], ((empty_bits >> s) & 1u) ^ 1u)) {
                    if (lane == 0) atomicExch(&g_gemm_timeout_flag, 1);
                    return;
                }
                empty_bits ^= (1u << s);
                if (elect_one_sync()) {
                    uint8_t* stg = smem + s * stage_bytes;
                    mbar_arrive_expect_tx(&sh.full_bar[s], static_cast<uint32_t>(stage_bytes));
                    if (p.w_packed)
                        tma_load_2d_hint(stg, tmap_w, &sh.full_bar[s], 0, (bx * p.kb_total + kb) * kBlockM, pol_w);
                    else
                        tma_load_2d_hint(stg, tmap_w, &sh.full_bar[s], kb * kBlockK, n0, pol_w);
                    for (int c = 0; c < p.nt; ++c)
                        tma_load_2d_hint(stg + kTileABytes + c * p.bn * (kBlockK * 2), tmap_x, &sh.full_bar[s],
                                         kb * kBlockK, t0 + c * p.bn, pol_x);
                }
                __syncwarp();
                s = (s + 1 == p.stages) ? 0 : s + 1;
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_bf16(kBlockM, static_cast<uint32_t>(p.bn));
        uint32_t full_bits = 0, tmem_empty_bits = 0;
        int s = 0, buf = 0;
        for (int tile = first; tile < n_tiles; tile += stride) {
            const int bz = tile / (gx * gy);
            const int kb0 = bz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            // the epilogue must have drained the accumulator buffer this tile is going to use
            if (!mbar_wait_warp(&tmem_empty_bar[buf], ((tmem_empty_bits >> buf) & 1u) ^ 1u)) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 4); return; }
            tmem_empty_bits ^= (1u << buf);
            tcgen05_fence_after();
            const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * p.acc_stride);
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!mbar_wait_warp(&sh.full_bar[s], (full_bits >> s) & 1u)) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 2); return; }
                full_bits ^= (1u << s);
                tcgen05_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                    for (int c = 0; c < p.nt; ++c) {
                        const uint64_t b_desc = make_smem_desc_sw128(a_addr + kTileABytes + c * p.bn * (kBlockK * 2));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_ss(acc + c * p.bn, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                         (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&sh.empty_bar[s]);
                    if (kb + 1 == kb1) umma_commit(buf == 0 ? sh.tmem_full_bar : tmem_empty_bar + 2);      // tmem_full[buf]
                }
                __syncwarp();
                s = (s + 1 == p.stages) ? 0 : s + 1;
            }
            buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
        }
    } else if (warp >= 4) {
        // 4 epilogue warps (one per TMEM lane quarter) or 8 (two per quarter, each half of the columns)
        const int w4 = (warp - 4) & 3;         // TMEM lane quarter this warp may access
        const int nl = w4 * 32 + lane;         // local weight row == TMEM lane
        const int ntok = p.nt * p.bn;
        const int n_groups = ntok / 16;
        const int g_begin = (blockDim.x > kGemmThreads) ? ((warp - 4) >> 2) * (n_groups / 2) : 0;
        const int g_end = (blockDim.x > kGemmThreads) ? (((warp - 4) >> 2) == 0 ? n_groups / 2 : n_groups) : n_groups;
        uint32_t tmem_bits = 0;
        int buf = 0;
        if (pdl) pdl_wait();
        for (int tile = first; tile < n_tiles; tile += stride) {
            const int bx = tile % gx, by = (tile / gx) % gy, bz = tile / (gx * gy);
            const int n0 = bx * kBlockM, t0 = by * p.nt * p.bn;
            const bool ready = mbar_wait(buf == 0 ? sh.tmem_full_bar : tmem_empty_bar + 2, (tmem_bits >> buf) & 1u);
            tmem_bits ^= (1u << buf);
            if (!ready) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 3); return; }
            tcgen05_fence_after();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16) + static_cast<uint32_t>(buf * p.acc_stride);
            float bias = 0.f;
            if (EPI != EPI_PARTIAL && p.bias != nullptr) bias = bf2f(p.bias[n0 + nl]);
            if (EPI != EPI_PARTIAL && p.staging_bytes > 0) {
                // Staged epilogue (batched episodes, 8 epilogue warps): the bf16 output tile is assembled in a
                // shared-memory buffer of its own (it does not alias the ring, the producer keeps loading) and
                // leaves as 16-byte row stores - the direct path's 2-byte stores (one per lane and token) were
                // what kept it behind the MMAs.
                constexpr int OUTW = (EPI == EPI_GEGLU) ? kBlockM / 2 : kBlockM;
                bf16* stg = reinterpret_cast<bf16*>(smem + p.stages * stage_bytes);
                for (int g = g_begin; g < g_end; ++g) {
                    uint32_t r[16];
                    tmem_ld_32x32b_x16(lane_addr + g * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int t = g * 16 + i;
                        if (EPI == EPI_GEGLU) {
                            const float v = bf16_round(__uint_as_float(r[i]));
                            const float up = __shfl_down_sync(0xffffffffu, v, 1);
                            if ((lane & 1) == 0) stg[t * OUTW + (nl >> 1)] = f2bf(bf16_round(glu_act_f32(v, p.glu_act)) * up);
                        } else {
                            float v = bf16_round(__uint_as_float(r[i]) + bias);
                            if (EPI == EPI_GELU) v = gelu_tanh_f32(v);
                    if (EPI == EPI_GELU_ERF) v = gelu_erf_f32(v);
                            stg[t * OUTW + nl] = f2bf(v);
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);          // TMEM drained: the MMAs may reuse it
                asm volatile("bar.sync 1, 256;\n" ::: "memory");          // the 8 epilogue warps
                const int et = static_cast<int>(threadIdx.x) - 128;
                const int col_base = (EPI == EPI_GEGLU) ? bx * (kBlockM / 2) : n0;
                for (int idx = et; idx < ntok * (OUTW / 8); idx += 256) {
                    const int t = idx / (OUTW / 8), ch = idx - t * (OUTW / 8);
                    if (t0 + t < p.T)
                        *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(t0 + t) * p.ldo + col_base + ch * 8) =
                            *reinterpret_cast<const uint4*>(stg + t * OUTW + ch * 8);
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");          // staging free for the next tile
                buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
                continue;
            }
            for (int g = g_begin; g < g_end; ++g) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(lane_addr + g * 16, r);
                tmem_ld_wait();
                const int tb = t0 + g * 16;
                if (EPI == EPI_PARTIAL) {
                    float* dst = p.partial + (static_cast<size_t>(bz) * p.T + tb) * p.Nw + n0 + nl;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (tb + i < p.T) dst[static_cast<size_t>(i) * p.Nw] = __uint_as_float(r[i]);
                } else if (EPI == EPI_GEGLU) {
                    // rows alternate gate_j (even lane) / up_j (odd lane): pair them with one shuffle
                    bf16* dst = p.out + static_cast<size_t>(tb) * p.ldo + bx * (kBlockM / 2) + (nl >> 1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float v = bf16_round(__uint_as_float(r[i]));
                        const float up = __shfl_down_sync(0xffffffffu, v, 1);
                        if ((lane & 1) == 0 && tb + i < p.T)
                            dst[static_cast<size_t>(i) * p.ldo] = f2bf(bf16_round(glu_act_f32(v, p.glu_act)) * up);
                    }
                } else {
                    bf16* dst = p.out + static_cast<size_t>(tb) * p.ldo + n0 + nl;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v = bf16_round(__uint_as_float(r[i]) + bias);
                        if (EPI == EPI_GELU) v = gelu_tanh_f32(v);
                    if (EPI == EPI_GELU_ERF) v = gelu_erf_f32(v);
                        if (tb + i < p.T) dst[static_cast<size_t>(i) * p.ldo] = f2bf(v);
                    }
                }
            }
            // accumulators are in registers / memory: hand TMEM back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2).  The pair owns 256 weight rows (UMMA M = 256: this CTA's
// 128 rows land in its own TMEM) and shares the token operand: each CTA loads only half of every
// token chunk and the tensor cores read the other half from the peer SM's shared memory, so the
// per-SM operand ingest per k-block drops from (128 + tokens) to (128 + tokens/2) rows — the quantity
// that bounds these GEMMs at T = 256..276 (measured ~48 B/clk/SM).  TMA loads of both CTAs complete
// on the leader's (even CTA) full barrier; the leader issues the MMAs and its commits arrive on the
// barriers of both CTAs.
template <int EPI>
__device__ __forceinline__ void gemm_tile_2cta(const GemmDev& p, const CUtensorMap* tmap_w, const CUtensorMap* tmap_xh,
                                               const GemmShared& sh, GemmPipe& st, const int bx, const int by,
                                               const int bz, const uint32_t crank) {
    uint8_t* smem = sh.ring;
    const int half = p.bn / 2;                                       // token rows of a chunk held by one CTA
    const int stage_bytes = kTileABytes + p.nt * half * (kBlockK * 2);   // per CTA
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_base = sh.tmem_base;
    const bool leader = (crank == 0);

    const int n0 = bx * kBlockM;
    const int t0 = by * p.nt * p.bn;
    const int kb0 = bz * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nkb = kb1 - kb0;

    if (warp == 0) {
        // all 32 lanes converged, operands warp-uniform, issue under elect.sync (see gemm_tile)
        const uint64_t pol_w = make_policy_evict_first();
        const uint64_t pol_x = make_policy_evict_last();
        int s = 0;
        for (int i = 0; i < nkb; ++i) {
            if (!mbar_wait_warp(&sh.empty_bar[s], ((st.empty_bits >> s) & 1u) ^ 1u)) {
                if (lane == 0) atomicExch(&g_gemm_timeout_flag, 1);
                break;
            }
            st.empty_bits ^= (1u << s);
            if (elect_one_sync()) {
                uint8_t* stg = smem + s * stage_bytes;
                if (leader) mbar_arrive_expect_tx(&sh.full_bar[s], static_cast<uint32_t>(2 * stage_bytes));
                const uint32_t bar = map_to_cta(&sh.full_bar[s], 0u);       // the leader's barrier
                const int kcoord = (kb0 + i) * kBlockK;
                if (p.w_packed)
                    tma_load_2d_2sm_hint(stg, tmap_w, bar, 0, (bx * p.kb_total + kb0 + i) * kBlockM, pol_w);
                else
                    tma_load_2d_2sm_hint(stg, tmap_w, bar, kcoord, n0, pol_w);
                for (int c = 0; c < p.nt; ++c)
                    tma_load_2d_2sm_hint(stg + kTileABytes + c * half * (kBlockK * 2), tmap_xh, bar, kcoord,
                                         t0 + c * p.bn + static_cast<int>(crank) * half, pol_x);
            }
            __syncwarp();
            s = (s + 1 == p.stages) ? 0 : s + 1;
        }
    } else if (warp == 1) {
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, static_cast<uint32_t>(p.bn));
            bool ok = true;
            int s = 0;
            for (int i = 0; i < nkb; ++i) {
                if (!mbar_wait_warp(&sh.full_bar[s], (st.full_bits >> s) & 1u)) {
                    if (lane == 0) atomicExch(&g_gemm_timeout_flag, 2);
                    ok = false;
                    break;
                }
                st.full_bits ^= (1u << s);
                tcgen05_fence_after();
                if (i == 0) cta_stamp(3);
                if (elect_one_sync()) {
                    const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                    const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                    for (int c = 0; c < p.nt; ++c) {
                        const uint64_t b_desc = make_smem_desc_sw128(a_addr + kTileABytes + c * half * (kBlockK * 2));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_ss_2sm(tmem_base + c * p.bn, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                             (i > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2sm(&sh.empty_bar[s]);      // frees stage s in both CTAs
                }
                __syncwarp();
                s = (s + 1 == p.stages) ? 0 : s + 1;
            }
            if (ok && elect_one_sync()) umma_commit_2sm(sh.tmem_full_bar);  // both epilogues may start
            __syncwarp();
            cta_stamp(4);
        }
    } else if (warp >= 4) {
        const int w4 = warp - 4;
        const int nl = w4 * 32 + lane;
        const bool acc_ready = mbar_wait(sh.tmem_full_bar, st.tmem_bit);
        st.tmem_bit ^= 1u;
        if (!acc_ready && lane == 0) atomicExch(&g_gemm_timeout_flag, 3);
        tcgen05_fence_after();
        if (warp == 4) cta_stamp(5);
        float bias = 0.f;
        if (EPI != EPI_PARTIAL && p.bias != nullptr) bias = bf2f(p.bias[n0 + nl]);
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16);
        const int ntok = p.nt * p.bn;
        for (int g = 0; acc_ready && g < ntok / 16; ++g) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(lane_addr + g * 16, r);
            tmem_ld_wait();
            if (EPI == EPI_PARTIAL) {
                float* tile = reinterpret_cast<float*>(smem);
#pragma unroll
                for (int i = 0; i < 16; ++i) tile[(g * 16 + i) * kBlockM + nl] = __uint_as_float(r[i]);
            } else {
                bf16* tile = reinterpret_cast<bf16*>(smem);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v = bf16_round(__uint_as_float(r[i]) + bias);
                    if (EPI == EPI_GELU) v = gelu_tanh_f32(v);
                    if (EPI == EPI_GELU_ERF) v = gelu_erf_f32(v);
                    tile[(g * 16 + i) * kBlockM + nl] = f2bf(v);
                }
            }
        }
        tcgen05_fence_before();
    }
    // The epilogue tile aliases the ring: the peer's tensor cores may still be reading THIS CTA's B
    // half until the pair's accumulators are complete.  Warps 4..7 only write after tmem_full, which
    // the leader commits after every MMA of the tile — so the ring is free by then in both CTAs.
    __syncthreads();
    if (warp == 0) cta_stamp(6);
    gemm_epilogue_store<EPI>(p, smem, bx, bz, n0, t0);
    __syncthreads();
    if (warp == 0) cta_stamp(7);
}

// Ring + barrier carve-up of a dynamic shared-memory block and one-time initialisation.
// Returns the shared view; `smem_raw` needs ring_bytes + 1024 (alignment) + 256 (barriers).
// ---------------------------------------------------------------------------------------------
// Persistent CTA pairs (batched episodes, bf16 epilogues): the pair walks 256-weight-row x 256-token tiles;
// each CTA loads its 128 weight rows and 128 of the tokens (tcgen05 cta_group::2 reads the other half from the
// peer), the accumulator is double-buffered in TMEM (2 x 256 columns per CTA) and the epilogue - 8 warps,
// staged in a buffer of its own, 16-byte row stores - runs under the MMAs of the next tile.
//   leader (rank 0): full[s] collects the bytes of both CTAs, issues the MMAs, its commits arrive on empty[s]
//   and tmem_full[buf] of both CTAs; tmem_empty[buf] lives in the leader and counts the 16 epilogue warps of
//   the pair (the peer's arrive remotely).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}

// Raster order of the persistent pairs: bands of `band` weight tile pairs; inside a band the weight index
// runs fastest, so the pairs in flight share `band` weight tiles and (pairs / band) token tiles and a band's
// weights stay in L2 while every token tile streams past them once.  Weight-fastest over all of gxp (the
// old order) cycles the whole weight matrix through L2 once per token tile: for gate/up at 64 episodes ncu
// counted 8.8 GB of DRAM reads against 0.2 GB of operands.
__host__ __device__ __forceinline__ void pair_tile_coords(const int tile, const int gxp, const int gy, const int band, int& xp, int& ty) {
    const int per_band = band * gy;
    const int b = tile / per_band, r = tile - b * per_band;
    const int w = (gxp - b * band < band) ? gxp - b * band : band;      // the last band may be narrower
    ty = r / w;
    xp = b * band + (r - ty * w);
}

// Generalised in round 2 for the batch-1 Gemma prefill (T = 276 = two UMMA-N chunks of 144 tokens, one 288-column
// accumulator, split-K slices as a third tile dimension, fp32-partial epilogue straight from TMEM): what bounds the
// one-CTA kernel there is shared-memory bandwidth (per k-block 69 KB of operand reads by the tensor core + 52 KB of
// TMA writes = 945 clk at 128 B/clk, measured 945); a pair halves the token bytes each SM stages and reads.
template <int EPI>
__device__ __forceinline__ void gemm_pair_persistent(const GemmDev& p, const CUtensorMap* tmap_w, const CUtensorMap* tmap_xh,
                                                     const GemmShared& sh, uint64_t* xbar, const int gxp, const int gy,
                                                     const int gz, const uint32_t crank, const int first, const int stride) {
    uint8_t* smem = sh.ring;
    const int half = p.bn / 2;                                            // token rows of a chunk staged by one CTA
    const int stage_bytes = kTileABytes + p.nt * half * (kBlockK * 2);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_base = sh.tmem_base;
    const bool leader = (crank == 0);
    const int tiles_xy = gxp * gy;
    const int n_tiles = tiles_xy * gz;
    const int ntok = p.nt * p.bn;                    // tokens (TMEM columns) of one tile
    uint64_t* tmem_empty = xbar;                     // [2], used in the leader
    uint64_t* tmem_full1 = xbar + 2;                 // tmem_full of buffer 1 (buffer 0: sh.tmem_full_bar)

    if (warp == 0) {
        // all 32 lanes converged, operands warp-uniform, issue under elect.sync (see gemm_tile)
        const uint64_t pol_w = p.l2_policy == 0 ? make_policy_evict_first() : p.l2_policy == 1 ? make_policy_evict_normal() : make_policy_evict_last();
        const uint64_t pol_x = p.l2_policy == 0 ? make_policy_evict_last() : p.l2_policy == 1 ? make_policy_evict_normal() : make_policy_evict_first();
        uint32_t empty_bits = 0;
        int s = 0;
        const uint32_t bar0 = map_to_cta(&sh.full_bar[0], 0u);      // the leader's full[0]; full[s] is 8 bytes further per stage
        // Engine weights do not depend on the previous kernel: the weight blocks of the first ring are requested before
        // the programmatic-dependency wait, only the token operand waits.
        int pre = 0;
        if (p.w_static && first < n_tiles) {
            const int tz = first / tiles_xy;
            int xp, ty;
            pair_tile_coords(first - tz * tiles_xy, gxp, gy, p.band, xp, ty);
            const int bx = 2 * xp + static_cast<int>(crank);
            const int kb0 = tz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            pre = min(p.stages, kb1 - kb0);
            if (elect_one_sync()) {
                for (int i = 0; i < pre; ++i) {
                    if (leader) mbar_arrive_expect_tx(&sh.full_bar[i], static_cast<uint32_t>(2 * stage_bytes));
                    tma_load_2d_2sm_hint(smem + i * stage_bytes, tmap_w, bar0 + static_cast<uint32_t>(i) * 8u, 0,
                                         (bx * p.kb_total + kb0 + i) * kBlockM, pol_w);
                }
            }
            __syncwarp();
        }
        pdl_wait();
        pdl_trigger();
        trace_stamp(p.trace, 1);
        cta_stamp(2);
        for (int tile = first; tile < n_tiles; tile += stride) {
            const int tz = tile / tiles_xy;
            int xp, ty;
            pair_tile_coords(tile - tz * tiles_xy, gxp, gy, p.band, xp, ty);
            const int bx = 2 * xp + static_cast<int>(crank), t0 = ty * ntok;
            const int kb0 = tz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                const bool w_requested = (tile == first && kb - kb0 < pre);     // its weight block is already in flight
                if (!w_requested) {
                    if (!mbar_wait_warp(&sh.empty_bar[s], ((empty_bits >> s) & 1u) ^ 1u)) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 1); return; }
                }
                empty_bits ^= (1u << s);
                if (elect_one_sync()) {
                    uint8_t* stg = smem + s * stage_bytes;
                    const uint32_t bar = bar0 + static_cast<uint32_t>(s) * 8u;
                    if (!w_requested) {
                        if (leader) mbar_arrive_expect_tx(&sh.full_bar[s], static_cast<uint32_t>(2 * stage_bytes));
                        tma_load_2d_2sm_hint(stg, tmap_w, bar, 0, (bx * p.kb_total + kb) * kBlockM, pol_w);
                    }
                    for (int c = 0; c < p.nt; ++c)
                        tma_load_2d_2sm_hint(stg + kTileABytes + c * half * (kBlockK * 2), tmap_xh, bar, kb * kBlockK,
                                             t0 + c * p.bn + static_cast<int>(crank) * half, pol_x);
                }
                __syncwarp();
                s = (s + 1 == p.stages) ? 0 : s + 1;
            }
        }
    } else if (warp == 1) {
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, static_cast<uint32_t>(p.bn));
            uint32_t full_bits = 0, tmem_empty_bits = 0;
            int s = 0, buf = 0;
            for (int tile = first; tile < n_tiles; tile += stride) {
                const int tz = tile / tiles_xy;
                const int kb0 = tz * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                if (!mbar_wait_warp(&tmem_empty[buf], ((tmem_empty_bits >> buf) & 1u) ^ 1u)) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 4); return; }
                tmem_empty_bits ^= (1u << buf);
                tcgen05_fence_after();
                const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * p.acc_stride);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!mbar_wait_warp(&sh.full_bar[s], (full_bits >> s) & 1u)) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 2); return; }
                    full_bits ^= (1u << s);
                    tcgen05_fence_after();
                    if (tile == first && kb == kb0) cta_stamp(3);
                    if (elect_one_sync()) {
                        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
                        const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                        for (int c = 0; c < p.nt; ++c) {
                            const uint64_t b_desc = make_smem_desc_sw128(a_addr + kTileABytes + c * half * (kBlockK * 2));
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_bf16_ss_2sm(acc + c * p.bn, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_2sm(&sh.empty_bar[s]);
                        if (kb + 1 == kb1) umma_commit_2sm(buf == 0 ? sh.tmem_full_bar : tmem_full1);
                    }
                    __syncwarp();
                    s = (s + 1 == p.stages) ? 0 : s + 1;
                }
                buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
            }
            cta_stamp(4);
        }
    } else if (warp >= 4) {
        // 8 epilogue warps (batched episodes) or 12 (batch 1, where the drain and the last tile's activation math are
        // on the critical path): `parts` warps per TMEM lane quarter, each a share of the columns
        const int w4 = (warp - 4) & 3, hf = (warp - 4) >> 2;
        const int nl = w4 * 32 + lane;
        const int n_groups = ntok / 16;
        const int epi_threads = static_cast<int>(blockDim.x) - 128, parts = epi_threads >> 7;
        const int g_begin = hf * n_groups / parts, g_end = (hf + 1) * n_groups / parts;
        constexpr int OUTW = (EPI == EPI_GEGLU) ? kBlockM / 2 : kBlockM;
        bf16* stg = reinterpret_cast<bf16*>(smem + p.stages * stage_bytes);
        const uint32_t empty_remote0 = map_to_cta(&tmem_empty[0], 0u), empty_remote1 = map_to_cta(&tmem_empty[1], 0u);
        uint32_t tmem_bits = 0;
        int buf = 0;
        pdl_wait();
        for (int tile = first; tile < n_tiles; tile += stride) {
            const int tz = tile / tiles_xy;
            int xp, ty;
            pair_tile_coords(tile - tz * tiles_xy, gxp, gy, p.band, xp, ty);
            const int bx = 2 * xp + static_cast<int>(crank), t0 = ty * ntok;
            const int n0 = bx * kBlockM;
            const bool ready = mbar_wait(buf == 0 ? sh.tmem_full_bar : tmem_full1, (tmem_bits >> buf) & 1u);
            tmem_bits ^= (1u << buf);
            if (!ready) { if (lane == 0) atomicExch(&g_gemm_timeout_flag, 3); return; }
            tcgen05_fence_after();
            if (warp == 4 && tile == first) cta_stamp(5);
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(w4 * 32) << 16) + static_cast<uint32_t>(buf * p.acc_stride);
            const bool real_tile = n0 < p.Nw;               // false: the padding CTA of an odd tile count
            // The drains below keep one tcgen05.ld in flight while the previous 16 columns are consumed.
            if (EPI == EPI_PARTIAL) {
                // fp32 partial sums of split-K slice tz, straight from TMEM: a warp's 32 lanes are 32 consecutive
                // features, so every store instruction writes one full 128-byte line of a token row
                float* dst = p.partial + static_cast<size_t>(tz) * p.T * p.Nw + n0 + nl;
                uint32_t ra[16], rb[16];
                tmem_ld_32x32b_x16(lane_addr + g_begin * 16, ra);
                for (int g = g_begin; g < g_end; g += 2) {
                    tmem_ld_wait_regs(ra);
                    if (g + 1 < g_end) tmem_ld_32x32b_x16(lane_addr + (g + 1) * 16, rb);
                    {
                        const int tb = t0 + g * 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (real_tile && tb + i < p.T) dst[static_cast<size_t>(tb + i) * p.Nw] = __uint_as_float(ra[i]);
                    }
                    if (g + 1 < g_end) {
                        tmem_ld_wait_regs(rb);
                        if (g + 2 < g_end) tmem_ld_32x32b_x16(lane_addr + (g + 2) * 16, ra);
                        const int tb = t0 + (g + 1) * 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (real_tile && tb + i < p.T) dst[static_cast<size_t>(tb + i) * p.Nw] = __uint_as_float(rb[i]);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(&tmem_empty[buf]);
                    else mbar_arrive_cluster(buf == 0 ? empty_remote0 : empty_remote1);
                }
                if (warp == 4 && tile == first) { cta_stamp(6); cta_stamp(7); }
                buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
                continue;
            }
            float bias = 0.f;
            if (p.bias != nullptr && real_tile) bias = bf2f(p.bias[n0 + nl]);
            // One accumulator buffer (batch 1): the next tile's MMAs wait for this drain, so GeGLU leaves TMEM as raw
            // bf16 gate / up values ([token][128]) and the activation math runs from the staging tile, under the next
            // tile's main loop.  Two buffers (batched): the drain is off the critical path, GeGLU is applied in
            // registers and the staging tile is half as large.
            const bool raw_geglu = (EPI == EPI_GEGLU) && p.acc_bufs == 1;
            auto consume = [&](const uint32_t (&r)[16], const int g) {
                if (EPI == EPI_GEGLU && !raw_geglu) {
                    // lanes 2j / 2j+1 hold gate_j / up_j.  Two tokens per step so that both lanes of a pair do a
                    // GELU: the even lane finishes token i, the odd lane token i + 1, after one exchange.
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float v0 = bf16_round(__uint_as_float(r[i])), v1 = bf16_round(__uint_as_float(r[i + 1]));
                        const float recv = __shfl_xor_sync(0xffffffffu, (lane & 1) ? v0 : v1, 1);
                        const float gate = (lane & 1) ? recv : v0, up = (lane & 1) ? v1 : recv;
                        stg[(g * 16 + i + (lane & 1)) * OUTW + (nl >> 1)] = f2bf(bf16_round(glu_act_f32(gate, p.glu_act)) * up);
                    }
                } else if (EPI == EPI_GELU || EPI == EPI_GELU_ERF) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float x0 = bf16_round(__uint_as_float(r[i]) + bias);
                        const float v = EPI == EPI_GELU_ERF ? gelu_erf_f32(x0) : gelu_tanh_f32(x0);
                        stg[(g * 16 + i) * kBlockM + nl] = f2bf(v);
                    }
                } else {
                    // one packed conversion per two tokens (cvt throughput bounds this drain: 16 per clock per SM)
                    unsigned short* stg16 = reinterpret_cast<unsigned short*>(stg);
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const uint32_t pk = pack_bf16x2(__uint_as_float(r[i]) + bias, __uint_as_float(r[i + 1]) + bias);
                        stg16[(g * 16 + i) * kBlockM + nl] = static_cast<unsigned short>(pk & 0xffffu);
                        stg16[(g * 16 + i + 1) * kBlockM + nl] = static_cast<unsigned short>(pk >> 16);
                    }
                }
            };
            {
                uint32_t ra[16], rb[16];
                tmem_ld_32x32b_x16(lane_addr + g_begin * 16, ra);
                for (int g = g_begin; g < g_end; g += 2) {
                    tmem_ld_wait_regs(ra);
                    if (g + 1 < g_end) tmem_ld_32x32b_x16(lane_addr + (g + 1) * 16, rb);
                    consume(ra, g);
                    if (g + 1 < g_end) {
                        tmem_ld_wait_regs(rb);
                        if (g + 2 < g_end) tmem_ld_32x32b_x16(lane_addr + (g + 2) * 16, ra);
                        consume(rb, g + 1);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tmem_empty[buf]);
                else mbar_arrive_cluster(buf == 0 ? empty_remote0 : empty_remote1);
            }
            asm volatile("bar.sync 1, %0;\n" ::"r"(epi_threads) : "memory");
            if (warp == 4 && tile == first) cta_stamp(6);
            const int et = static_cast<int>(threadIdx.x) - 128;
            const int col_base = (EPI == EPI_GEGLU) ? bx * (kBlockM / 2) : n0;
            if (raw_geglu) {
                // weight rows alternate gate_j, up_j: 16 consecutive staged columns give 8 outputs
                const bf16x8* tile_s = reinterpret_cast<const bf16x8*>(stg);
                for (int idx = et; real_tile && idx < ntok * 8; idx += epi_threads) {
                    const int t = idx >> 3, ch = idx & 7;
                    if (t0 + t >= p.T) continue;
                    const bf16x8 lo = tile_s[t * 16 + 2 * ch];
                    const bf16x8 hi = tile_s[t * 16 + 2 * ch + 1];
                    bf16x8 o;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float2 a0 = unpack_bf16x2(lo.u[2 * j]), a1 = unpack_bf16x2(lo.u[2 * j + 1]);
                        const float2 b0 = unpack_bf16x2(hi.u[2 * j]), b1 = unpack_bf16x2(hi.u[2 * j + 1]);
                        o.u[j] = pack_bf16x2(bf16_round(glu_act_f32(a0.x, p.glu_act)) * a0.y, bf16_round(glu_act_f32(a1.x, p.glu_act)) * a1.y);
                        o.u[2 + j] = pack_bf16x2(bf16_round(glu_act_f32(b0.x, p.glu_act)) * b0.y, bf16_round(glu_act_f32(b1.x, p.glu_act)) * b1.y);
                    }
                    *reinterpret_cast<bf16x8*>(p.out + static_cast<size_t>(t0 + t) * p.ldo + col_base + ch * 8) = o;
                }
            } else {
                for (int idx = et; real_tile && idx < ntok * (OUTW / 8); idx += epi_threads) {
                    const int t = idx / (OUTW / 8), ch = idx - t * (OUTW / 8);
                    if (t0 + t < p.T)
                        *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(t0 + t) * p.ldo + col_base + ch * 8) =
                            *reinterpret_cast<const uint4*>(stg + t * OUTW + ch * 8);
                }
            }
            asm volatile("bar.sync 1, %0;\n" ::"r"(epi_threads) : "memory");
            if (warp == 4 && tile == first) cta_stamp(7);
            buf = (buf + 1 == p.acc_bufs) ? 0 : buf + 1;
        }
    }
}

__device__ __forceinline__ GemmShared gemm_setup_shared(uint8_t* smem_raw, int ring_bytes, int empty_count,
                                                        uint32_t tmem_cols, uint32_t* tmem_slot_out,
                                                        bool two_sm = false) {
    GemmShared sh;
    sh.ring = smem_align_1024(smem_raw);
    sh.full_bar = reinterpret_cast<uint64_t*>(sh.ring + ring_bytes);
    sh.empty_bar = sh.full_bar + kMaxStages;
    sh.tmem_full_bar = sh.empty_bar + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sh.tmem_full_bar + 1);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 1 && elect_one_sync()) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&sh.full_bar[i], 1);
            mbar_init(&sh.empty_bar[i], static_cast<uint32_t>(empty_count));
        }
        mbar_init(sh.tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (two_sm) cluster_sync_all();       // the pair allocates TMEM collectively: both CTAs are here
    if (warp == 2) {
        if (two_sm) { tmem_alloc_2sm(tmem_slot, tmem_cols); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    sh.tmem_base = *tmem_slot;
    *tmem_slot_out = sh.tmem_base;
    return sh;
}

}  // namespace blurr
