// tcgen05 / TMEM / TMA bf16 GEMM for every Linear on the Pi-0 control-step path.
//
//   Y[t, n] = sum_k X[t, k] * W[n, k]        X: activations [T, K] (K contiguous)
//                                            W: nn.Linear weight [N, K] (K contiguous)
//
// "Swap-AB" orientation: a CTA owns 128 *weight rows* (UMMA M = 128, TMEM lane = output
// feature n) and up to NT chunks of BN <= 256 *tokens* (UMMA N = BN, TMEM column = token).
// At batch 1 (T = 256 / 276 / 1 / 4) every weight byte is therefore streamed from HBM exactly
// once by exactly one CTA, and the token dimension pads to a multiple of 16 instead of 128.
// Both operands are K-major bf16 tiles of 64 elements (128 B) per row, written by TMA with
// the 128-byte swizzle and consumed through UMMA shared-memory descriptors.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue phase 1 (TMEM -> registers -> smem tile
// [token][n]); then all 8 warps run phase 2 (row-wise 16-byte vector epilogue + stores).
//
// Epilogues reproduce the reference's bf16 rounding points:
//   EPI_STORE   : bf16(acc + bias)                                   (nn.Linear output)
//   EPI_GELU    : bf16(gelu_tanh(bf16(acc + bias)))                  (SigLIP fc1, siglip.py:188-190)
//   EPI_GEGLU   : bf16(bf16(gelu_tanh(bf16(gate))) * bf16(up))       (GemmaMLP, modules.py:93-95)
//   EPI_PARTIAL : fp32 partial sums of one split-K slice -> workspace [z][t][n]; the consumer
//                 kernels in norm_consumers.cu finish bias/residual/norm in a fixed order.
#include "gemm_body.cuh"
#include "launch.cuh"

#include <mutex>
#include <string>
#include <unordered_map>

namespace blurr {


template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
               const __grid_constant__ CUtensorMap tmap_xs, const GemmDev p) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(p.trace, 0);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0) cta_stamp(0);
    if (warp == 0 && elect_one_sync()) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(p.cluster > 1 ? &tmap_xs : &tmap_x);
    }
    const int stage_bytes = kTileABytes + p.nt * p.bn * (kBlockK * 2);
    uint32_t tmem_base;
    GemmShared sh = gemm_setup_shared(smem_raw, p.stages * stage_bytes, p.cluster,
                                      static_cast<uint32_t>(p.tmem_cols), &tmem_base);
    if (p.cluster > 1) cluster_sync_all();      // peers' barriers are initialised before any remote arrive
    const uint32_t crank = (p.cluster > 1) ? cluster_ctarank() : 0u;
    if (warp == 0) cta_stamp(1);
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the tail of the
    // previous kernel, and so does the first ring of weight blocks: the producer and the epilogue warps
    // wait for the previous kernel (griddepcontrol.wait) inside gemm_tile, just before they first touch
    // the token operand / the output; the producer triggers the successor right after its wait (triggering
    // before it made the successor resident one kernel earlier and cost 0.1 ms per step: a CTA parked in
    // griddepcontrol.wait holds its shared memory and TMEM)
    GemmPipe st;
    gemm_tile<EPI>(p, &tmap_w, &tmap_x, &tmap_xs, sh, st, blockIdx.x, blockIdx.y, blockIdx.z, crank, true);
    trace_stamp(p.trace, 2);
    if (p.cluster > 1) cluster_sync_all();      // no CTA exits while peers may still signal its barriers
    if (warp == 2) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// Persistent 1-CTA kernel: at most one CTA per SM, each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads + 128, 1)
gemm_tcp_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                const GemmDev p, const int gx, const int gy, const int gz) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(p.trace, 0);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0 && elect_one_sync()) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
    }
    const int stage_bytes = kTileABytes + p.nt * p.bn * (kBlockK * 2);
    uint32_t tmem_base;
    GemmShared sh = gemm_setup_shared(smem_raw, p.stages * stage_bytes + p.staging_bytes, 1, static_cast<uint32_t>(p.tmem_cols),
                                      &tmem_base);
    uint64_t* tmem_empty_bar = sh.tmem_full_bar + 2;      // [0], [1]: after tmem_full_bar and the TMEM slot word
    if (threadIdx.x == 0) {
        const uint32_t epi_warps = blockDim.x / 32 - 4;     // 4, or 8 in the activation-heavy batched launches
        mbar_init(&tmem_empty_bar[0], epi_warps);           // one arrive per epilogue warp
        mbar_init(&tmem_empty_bar[1], epi_warps);
        mbar_init(&tmem_empty_bar[2], 1);                  // tmem_full of the second accumulator buffer
        fence_barrier_init();
    }
    __syncthreads();
    gemm_persistent<EPI>(p, &tmap_w, &tmap_x, sh, tmem_empty_bar, gx, gy, gz, blockIdx.x, gridDim.x, true);
    __syncthreads();
    trace_stamp(p.trace, 2);
    if (warp == 2) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// CTA-pair (cta_group::2) variant: launched as clusters of 2 along the weight-tile dimension.
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_xh,
                const GemmDev p) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(p.trace, 0);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0) cta_stamp(0);
    if (warp == 0 && elect_one_sync()) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_xh);
    }
    const int stage_bytes = kTileABytes + p.nt * (p.bn / 2) * (kBlockK * 2);
    uint32_t tmem_base;
    GemmShared sh = gemm_setup_shared(smem_raw, p.stages * stage_bytes, 1, static_cast<uint32_t>(p.tmem_cols),
                                      &tmem_base, true);
    cluster_sync_all();                         // both CTAs' barriers exist before any remote signal
    const uint32_t crank = cluster_ctarank();
    if (warp == 0) cta_stamp(1);
    pdl_wait();
    pdl_trigger();
    trace_stamp(p.trace, 1);
    if (warp == 0) cta_stamp(2);
    GemmPipe st;
    gemm_tile_2cta<EPI>(p, &tmap_w, &tmap_xh, sh, st, blockIdx.x, blockIdx.y, blockIdx.z, crank);
    tcgen05_fence_before();
    cluster_sync_all();                         // the peer is done with this CTA's smem / TMEM / barriers
    trace_stamp(p.trace, 2);
    if (warp == 2) tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// Persistent CTA pairs (gemm_pair_persistent in gemm_body.cuh): batched episodes with double-buffered accumulators,
// and the batch-1 Gemma prefill GEMMs (two 144-token chunks, split-K slices as tiles).
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads + 256, 1)
gemm_tcp2_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_xh, const GemmDev p,
                 const int gxp, const int gy, const int gz) {
    extern __shared__ uint8_t smem_raw[];
    trace_stamp(p.trace, 0);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    if (warp == 0) cta_stamp(0);
    if (warp == 0 && elect_one_sync()) {
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_xh);
    }
    const int stage_bytes = kTileABytes + p.nt * (p.bn / 2) * (kBlockK * 2);
    uint32_t tmem_base;
    GemmShared sh = gemm_setup_shared(smem_raw, p.stages * stage_bytes + p.staging_bytes, 1, static_cast<uint32_t>(p.tmem_cols),
                                      &tmem_base, true);
    uint64_t* xbar = sh.tmem_full_bar + 2;
    if (threadIdx.x == 0) {
        const uint32_t epi_warps = blockDim.x / 32 - 4;     // 8, or 12 in the batch-1 launches
        mbar_init(&xbar[0], 2 * epi_warps);                 // tmem_empty[buf]: the epilogue warps of both CTAs of the pair
        mbar_init(&xbar[1], 2 * epi_warps);
        mbar_init(&xbar[2], 1);                             // tmem_full of the second accumulator buffer
        fence_barrier_init();
    }
    __syncthreads();
    cluster_sync_all();                                     // both CTAs' barriers exist before any remote signal
    const uint32_t crank = cluster_ctarank();
    if (warp == 0) cta_stamp(1);
    gemm_pair_persistent<EPI>(p, &tmap_w, &tmap_xh, sh, xbar, gxp, gy, gz, crank, static_cast<int>(blockIdx.x >> 1),
                              static_cast<int>(gridDim.x >> 1));
    tcgen05_fence_before();
    cluster_sync_all();                                     // the peer is done with this CTA's smem / TMEM / barriers
    trace_stamp(p.trace, 2);
    if (warp == 2) tmem_dealloc_2sm(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// ---------------------------------------------------------------------------
// host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

struct TmapKey {
    const void* ptr;
    int rows, cols, ld, box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        h = h * 1000003u ^ static_cast<size_t>(k.rows);
        h = h * 1000003u ^ static_cast<size_t>(k.cols);
        h = h * 1000003u ^ static_cast<size_t>(k.ld);
        h = h * 1000003u ^ static_cast<size_t>(k.box_rows);
        return h;
    }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

// 2D bf16 tensor [rows][cols] with row stride `ld` elements; box = {64 cols, box_rows}.
static int get_tmap(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out,
                    std::string* err) {
    TmapKey key{ptr, rows, cols, ld, box_rows};
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *out = it->second; return 0; }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { *err = "cuTensorMapEncodeTiled entry point not available"; return -1; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((static_cast<size_t>(ld) * 2) & 15)) {
        *err = "TMA operand must be 16-byte aligned with a 16-byte-multiple row stride";
        return -1;
    }
    CUtensorMap m;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
        return -1;
    }
    g_tmaps.emplace(key, m);
    *out = m;
    return 0;
}

int gemm_get_tensor_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out, std::string* err) {
    return get_tmap(ptr, rows, cols, ld, box_rows, out, err);
}

static bool g_pdl_enabled = true;
bool pdl_enabled() { return g_pdl_enabled; }
void pdl_set_enabled(bool on) { g_pdl_enabled = on; }

int gemm_take_timeout_flag() {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, g_gemm_timeout_flag, sizeof(int)) != cudaSuccess) return -1;
    if (v != 0) {
        const int zero = 0;
        cudaMemcpyToSymbol(g_gemm_timeout_flag, &zero, sizeof(int));
    }
    return v;
}

const int* gemm_timeout_flag_ptr() {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_gemm_timeout_flag) == cudaSuccess ? static_cast<const int*>(p) : nullptr;
}

int gemm_set_cta_trace(void* dev_ptr) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
    return cudaMemcpyToSymbol(g_cta_trace, &p, sizeof(p)) == cudaSuccess ? 0 : -1;
}

void gemm_forget_tensor_maps() {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    g_tmaps.clear();
}

static int next_pow2_cols(int c) {
    int v = 32;
    while (v < c) v <<= 1;
    return v;
}

static constexpr int kSmemBudget = 227 * 1024;
static constexpr int kRingBytes = 225 * 1024;   // pipeline ring (the rest of the 227 KB: barriers, alignment slack)

// Measured on B200 (tools/time_gemm.py, round 1): multicasting the activation tile across a cluster of
// 2/4 CTAs is *slower* than unicast at every Pi-0 shape (the CTAs of a cluster advance in lock-step and
// each SM still ingests the full tile), so it is off by default; the knob stays for experiments.
static int g_cluster_max = 1;
static constexpr int kTargetCtas = 148;   // one CTA per SM of a B200
static constexpr int kCoResidentSmem = 112 * 1024;   // dynamic smem per CTA that still lets two CTAs share an SM
void gemm_set_cluster_max(int c) { g_cluster_max = c < 1 ? 1 : (c > 8 ? 8 : c); }

// Measured on B200 (round 1): CTA pairs are correct but not faster at batch 1 (61 vs 59 us on the gate/up
// shape) — the loss there is per-tile fixed cost, which the persistent kernel removes; pairs help at
// batch 64 (gate/up 1059 vs 952 TFLOP/s with two CTAs per SM).  -1 = automatic: pairs for the GeGLU GEMM above
// 1024 tokens (at 276 tokens they measured 61.8 vs 53.6 us under ncu).
static int g_use_2cta = -1;
void gemm_set_use_2cta(int on) { g_use_2cta = on < 0 ? -1 : (on ? 1 : 0); }
// Persistent kernel: measured on B200 (round 1) it wins only while the epilogue is trivial (T = 16:
// 33.5 vs 39.9 us on the gate/up shape); from T >= 64 its direct TMEM -> global epilogue (2-byte stores,
// 4 warps) stalls the next tile's MMAs longer than the bubbles it removes (GeGLU, T = 276: 95 vs 59 us),
// so it is used for few-token GEMMs only.
// Large token counts (batched episodes), measured on B200 at 64 episodes (TFLOP/s; one tile per CTA with two
// CTAs per SM | CTA pairs, two per SM | persistent 128 x 256 tiles with the accumulator double-buffered in
// TMEM and a direct TMEM -> global epilogue):
//   gate/up (GeGLU)   952 | 1059 |  728        down (fp32 partial)  1106 | 1109 | 1211
//   SigLIP fc1 (GELU) 722 |  718 |  475        VLM qkv (partial)     825 |  725 | 1091
//   SigLIP qkv (store) 958 | 857 | 1028
// The double-buffered persistent kernel wins while the epilogue is a plain store (it then hides entirely under
// the next tile's MMAs) and loses when the epilogue carries the activation math; with two epilogue warps per
// TMEM lane quarter it reaches 777 (fc1) and 974 (gate/up).  -1 = automatic: persistent CTA pairs for GeGLU
// (gemm_launch_pairp), persistent single CTAs for everything else; 0 = never persistent; 1 = single-CTA
// persistent for every epilogue; 2 = same as automatic.  (These kernels run against the 1 kW power cap: numbers
// taken back to back differ by +-10 % with the order of the launches.)
static int g_large_t_mode = -1;
void gemm_set_large_t_mode(int mode) { g_large_t_mode = mode; }
static int g_pair_band = 0;
static int g_pair_policy = -1;
void gemm_set_pair_policy(int policy) { g_pair_policy = policy; }
void gemm_set_pair_band(int band) { g_pair_band = band; }
static int g_persistent = 1;
static constexpr int kPersistentMaxTokens = 32;
void gemm_set_persistent(int on) { g_persistent = on ? 1 : 0; }

// Cap of the TMA ring depth.  4 measured best for the whole bs=1 step (4.363 ms vs 4.394 with 8, 4.566 with 3;
// same-box A/B): deeper rings buy nothing per GEMM (see the planner notes below) and a smaller footprint lets
// the kernels of the three streams share SMs.
static int g_max_stages = 4;
void gemm_set_max_stages(int n) { g_max_stages = n < 1 ? 1 : (n > 12 ? 12 : n); }

static bool plan_fits(int bn, int nt, int kb_per_split, int epi, int* stages_out, int* smem_out, int b_div = 1) {
    const int stage_bytes = kTileABytes + nt * (bn / b_div) * kBlockK * 2;
    const int tile_bytes = nt * bn * kBlockM * (epi == EPI_PARTIAL ? 4 : 2);
    const int avail = kRingBytes;
    int stages = avail / stage_bytes;
    if (stages < 1) return false;
    if (stages > g_max_stages) stages = g_max_stages;
    const int want = kb_per_split > 0 ? kb_per_split : 1;
    if (stages > want) stages = want;
    // the epilogue tile aliases the pipeline buffers: keep enough bytes for it
    while (stages * stage_bytes < tile_bytes) ++stages;
    if (stages * stage_bytes > avail) return false;
    *stages_out = stages;
    *smem_out = stages * stage_bytes + 1024 + 256;
    return true;
}

GemmPlan gemm_make_plan(int T, int Nw, int K, int splitk, int epi, int bn_override) {
    GemmPlan pl{};
    pl.valid = false;
    if (Nw % kBlockM != 0 || T <= 0 || K <= 0) return pl;
    pl.kb_total = (K + kBlockK - 1) / kBlockK;
    if (splitk < 1) splitk = 1;
    if (splitk > pl.kb_total) splitk = pl.kb_total;
    pl.kb_per_split = (pl.kb_total + splitk - 1) / splitk;
    pl.splitk = (pl.kb_total + pl.kb_per_split - 1) / pl.kb_per_split;   // no empty slices
    // Token chunking.  A chunk is one UMMA N (<= 256 tokens, multiple of 16).  Two chunks share a
    // CTA (one pass over the weight tile for up to ~448 tokens) when a >= 3-stage pipeline and the
    // epilogue tile still fit in shared memory; otherwise chunks go to blockIdx.y.
    const int t16 = (T + 15) / 16 * 16;
    int bn, nt, gy, stages = 0, smem = 0;
    if (bn_override > 0) {
        bn = bn_override; nt = 1; gy = (T + bn - 1) / bn;
        if (!plan_fits(bn, nt, pl.kb_per_split, epi, &stages, &smem)) return pl;
    } else if (t16 <= 256) {
        bn = t16; nt = 1; gy = 1;
        // A GEMM whose epilogue needs complete sums cannot split K; when its weight tiles alone
        // leave most SMs idle, split the *tokens* across CTAs instead: every CTA then ingests
        // (128 + bn) rows per k-block instead of (128 + T), and ~all SMs pull operands in parallel
        // (measured: the per-SM operand ingest rate, not HBM, bounds these small-N GEMMs).
        const int gx = Nw / kBlockM;
        if (epi != EPI_PARTIAL && gx * 2 <= kTargetCtas && t16 >= 64) {
            int chunks = kTargetCtas / gx;
            if (chunks > t16 / 32) chunks = t16 / 32;            // keep at least 32 tokens per CTA
            if (chunks > 1) {
                bn = ((T + chunks - 1) / chunks + 15) / 16 * 16;
                gy = (T + bn - 1) / bn;
            }
        }
        if (!plan_fits(bn, nt, pl.kb_per_split, epi, &stages, &smem)) return pl;
    } else {
        const int chunks = (t16 + 255) / 256;
        bn = ((T + chunks - 1) / chunks + 15) / 16 * 16;
        nt = 1; gy = chunks;
        int st2 = 0, sm2 = 0;
        if (chunks == 2 && plan_fits(bn, 2, pl.kb_per_split, epi, &st2, &sm2) &&
            (st2 >= 3 || st2 >= pl.kb_per_split)) {
            nt = 2; gy = 1; stages = st2; smem = sm2;
        } else if (!plan_fits(bn, 1, pl.kb_per_split, epi, &stages, &smem)) {
            return pl;
        }
    }
    pl.bn = bn; pl.nt = nt; pl.stages = stages; pl.smem_bytes = smem;
    pl.tmem_cols = next_pow2_cols(nt * bn);
    if (pl.tmem_cols > 512) return pl;
    pl.grid_x = Nw / kBlockM;
    pl.grid_y = gy;
    // More CTAs than SMs: a ring shallow enough for two CTAs per SM (smem and TMEM halves) lets one
    // CTA's prologue / epilogue overlap the other's main loop and removes the second wave.  Measured
    // on the gate/up shape: 16 tokens 28.8 us (4 stages) vs 42.0 (8); 144 tokens 33.5 (3) vs 47.2 (4).
    if (pl.grid_x * gy * pl.splitk > kTargetCtas && pl.tmem_cols <= 256) {
        const int stage_bytes = kTileABytes + nt * bn * kBlockK * 2;
        const int tile_bytes = nt * bn * kBlockM * (epi == EPI_PARTIAL ? 4 : 2);
        int s2 = (kCoResidentSmem - 1280) / stage_bytes;
        if (s2 > 4) s2 = 4;
        if (s2 >= 2 && s2 < pl.stages && s2 * stage_bytes >= tile_bytes) {
            pl.stages = s2;
            pl.smem_bytes = s2 * stage_bytes + 1024 + 256;
        }
    }
    // activation multicast: the CTAs of a cluster own consecutive weight tiles and share the
    // token tile; each loads 1/C of it (a multiple of 8 rows keeps the 128B swizzle phase)
    pl.cluster = 1;
    if (g_cluster_max > 1 && nt * bn >= 64) {
        for (int c = g_cluster_max; c > 1; c >>= 1) {
            if (pl.grid_x % c == 0 && (nt * bn) % (8 * c) == 0) { pl.cluster = c; break; }
        }
    }
    pl.slice_rows = nt * bn / pl.cluster;
    // CTA pairs (cta_group::2): consecutive weight tiles share the token operand, each CTA loads half
    pl.two_cta = 0;
    const bool want_pairs = g_use_2cta == 1 || (g_use_2cta < 0 && T > 1024 && epi == EPI_GEGLU);
    if (want_pairs && pl.cluster == 1 && pl.grid_x % 2 == 0 && nt * bn >= 64 && bn % 16 == 0) {
        int st2 = 0, sm2 = 0;
        if (plan_fits(bn, nt, pl.kb_per_split, epi, &st2, &sm2, 2)) {
            pl.two_cta = 1;
            pl.stages = st2;
            pl.smem_bytes = sm2;
            // same co-residency rule as above: two CTAs (of different pairs) per SM overlap one tile's
            // epilogue with the other's main loop
            if (pl.grid_x * gy * pl.splitk > kTargetCtas && pl.tmem_cols <= 256) {
                const int stage_bytes = kTileABytes + nt * (bn / 2) * kBlockK * 2;
                const int tile_bytes = nt * bn * kBlockM * (epi == EPI_PARTIAL ? 4 : 2);
                int s2 = (kCoResidentSmem - 1280) / stage_bytes;
                if (s2 > 4) s2 = 4;
                if (s2 >= 2 && s2 < pl.stages && s2 * stage_bytes >= tile_bytes) {
                    pl.stages = s2;
                    pl.smem_bytes = s2 * stage_bytes + 1024 + 256;
                }
            }
        }
    }
    pl.valid = true;
    return pl;
}

template <int EPI>
static cudaError_t launch_epip(cudaStream_t stream, const GemmPlan& pl, const CUtensorMap& tw, const CUtensorMap& tx,
                               const GemmDev& d) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tcp_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int tiles = pl.grid_x * pl.grid_y * pl.splitk;
    const int ctas = tiles < kTargetCtas ? tiles : kTargetCtas;
    // epilogues with activation math get two warps per TMEM lane quarter so that they stay hidden under the
    // next tile's MMAs (double-buffered accumulators)
    const int threads = (d.staging_bytes > 0) ? kGemmThreads + 128 : kGemmThreads;
    return launch_kernel(gemm_tcp_kernel<EPI>, dim3(ctas), dim3(threads), static_cast<size_t>(pl.smem_bytes), stream,
                         tw, tx, d, pl.grid_x, pl.grid_y, pl.splitk);
}

template <int EPI>
static cudaError_t launch_epi2(cudaStream_t stream, const GemmPlan& pl, const CUtensorMap& tw, const CUtensorMap& txh,
                               const GemmDev& d) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(pl.grid_x, pl.grid_y, pl.splitk);
    return launch_kernel_cluster(gemm_tc2_kernel<EPI>, grid, dim3(kGemmThreads), static_cast<size_t>(pl.smem_bytes),
                                 stream, 2, tw, txh, d);
}

template <int EPI>
static cudaError_t launch_epi(cudaStream_t stream, const GemmPlan& pl, const CUtensorMap& tw,
                              const CUtensorMap& tx, const CUtensorMap& txs, const GemmDev& d) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid(pl.grid_x, pl.grid_y, pl.splitk);
    return launch_kernel_cluster(gemm_tc_kernel<EPI>, grid, dim3(kGemmThreads), static_cast<size_t>(pl.smem_bytes),
                                 stream, pl.cluster, tw, tx, txs, d);
}

template <int EPI>
static cudaError_t launch_pairp(cudaStream_t stream, int n_pairs, int smem, const CUtensorMap& tw, const CUtensorMap& txh,
                                const GemmDev& d, int gxp, int gy, int gz = 1, int threads = kGemmThreads + 128) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tcp2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    return launch_kernel_cluster(gemm_tcp2_kernel<EPI>, dim3(2 * n_pairs), dim3(threads), static_cast<size_t>(smem),
                                 stream, 2, tw, txh, d, gxp, gy, gz);
}

// Batched GEMMs with a bf16 epilogue (above 1024 tokens; automatic, or modes 2 / 3 of "gemm_large_t_mode"):
// persistent CTA pairs, tiles visited in raster bands.  At 64 episodes: gate/up 1429 TFLOP/s (1050 for one tile
// per CTA pair, 1086 for persistent pairs before the bands), SigLIP fc1 1112 vs 910, the plain stores (bf16
// hand-off of down / qkv / o) 1450 and 1335 against 1115 and 1072 for the single-CTA persistent kernel.
static bool gemm_pairp_applies(const GemmCall& c) {
    const bool epi_ok = c.epi == EPI_GEGLU || c.epi == EPI_GELU || c.epi == EPI_STORE;
    return (g_large_t_mode >= 2 || g_large_t_mode < 0) && c.w_packed && epi_ok && c.splitk <= 1 && c.bn_override == 0 && c.T > 1024 &&
           c.Nw % kBlockM == 0 && c.K % kBlockK == 0 && (c.epi != EPI_GEGLU || c.Nw % (2 * kBlockM) == 0);
}

static int pair_band_for(int gxp, int n_pairs, int K) {
    if (g_pair_band > 0) return g_pair_band < gxp ? g_pair_band : gxp;
    if (gxp <= n_pairs) return gxp;                 // one wave of pairs already holds every weight tile pair
    const long long pair_bytes = 2LL * kBlockM * K * 2;
    const long long band = (32LL << 20) / pair_bytes;
    return static_cast<int>(band < 4 ? 4 : (band > gxp ? gxp : band));
}

// Host-side view of the raster (C ABI `blurr_op_pair_raster`): the order in which the persistent pairs visit the
// tiles of a [T][N] output, from the same pair_tile_coords() the kernel runs.
int gemm_pair_raster(int N, int K, int T, int* band_out, int* n_pairs_out, int32_t* order, int capacity) {
    if (N <= 0 || K <= 0 || T <= 0 || N % kBlockM != 0) return -1;
    const int bn = 256;
    const int gxp = (N / kBlockM + 1) / 2, gy = (T + bn - 1) / bn;
    const int tiles = gxp * gy;
    const int n_pairs = tiles < kTargetCtas / 2 ? tiles : kTargetCtas / 2;
    const int band = pair_band_for(gxp, n_pairs, K);
    if (band_out) *band_out = band;
    if (n_pairs_out) *n_pairs_out = n_pairs;
    for (int t = 0; order != nullptr && t < tiles && t < capacity; ++t) {
        int xp, ty;
        pair_tile_coords(t, gxp, gy, band, xp, ty);
        order[2 * t] = xp; order[2 * t + 1] = ty;
    }
    return tiles;
}

static int gemm_launch_pairp(cudaStream_t stream, const GemmCall& c, std::string* err) {
    const int kb_total = c.K / kBlockK;
    const int bn = 256, half = bn / 2;
    CUtensorMap tw, txh;
    if (get_tmap(c.W, c.Nw * kb_total, kBlockK, kBlockK, kBlockM, &tw, err)) return -1;
    if (get_tmap(c.X, c.T, c.K, c.ldx, half, &txh, err)) return -1;
    GemmDev d{};
    d.T = c.T; d.bn = bn; d.nt = 1; d.kb_total = kb_total; d.kb_per_split = kb_total;
    d.tmem_cols = 512; d.acc_bufs = 2; d.acc_stride = 256; d.Nw = c.Nw;
    d.bias = c.bias; d.out = c.out; d.ldo = c.ldo; d.partial = nullptr;
    d.cluster = 1; d.slice_rows = 0; d.w_packed = 1; d.trace = c.trace; d.w_static = c.w_static; d.glu_act = c.glu_act;
    d.staging_bytes = bn * (c.epi == EPI_GEGLU ? kBlockM / 2 : kBlockM) * 2;
    const int stage_bytes = kTileABytes + half * kBlockK * 2;
    d.stages = (kRingBytes - d.staging_bytes) / stage_bytes;
    if (d.stages > kMaxStages) d.stages = kMaxStages;
    const int smem = d.stages * stage_bytes + d.staging_bytes + 1024 + 256;
    // an odd number of weight tiles: the last pair's second CTA multiplies TMA's out-of-bounds zeros and stores nothing
    const int gxp = (c.Nw / kBlockM + 1) / 2, gy = (c.T + bn - 1) / bn;
    const int tiles = gxp * gy;
    const int n_pairs = tiles < kTargetCtas / 2 ? tiles : kTargetCtas / 2;
    // Raster bands (pair_tile_coords).  When one wave of pairs cannot hold every weight tile pair, weight-fastest
    // order streams the whole weight matrix from DRAM once per token tile; bands of ~32 MB of weights keep a band
    // L2-resident while the tokens stream past it once per band.
    d.l2_policy = g_pair_policy >= 0 ? g_pair_policy : 1;       // both operands are re-read by later tiles: plain LRU (measured best)
    d.band = pair_band_for(gxp, n_pairs, c.K);
    cudaError_t e;
    switch (c.epi) {
        case EPI_GEGLU: e = launch_pairp<EPI_GEGLU>(stream, n_pairs, smem, tw, txh, d, gxp, gy); break;
        case EPI_GELU:    e = c.glu_act == 2 ? launch_pairp<EPI_GELU_ERF>(stream, n_pairs, smem, tw, txh, d, gxp, gy) : launch_pairp<EPI_GELU>(stream, n_pairs, smem, tw, txh, d, gxp, gy); break;
        default:        e = launch_pairp<EPI_STORE>(stream, n_pairs, smem, tw, txh, d, gxp, gy); break;
    }
    if (e != cudaSuccess) { *err = std::string("gemm (persistent pairs) launch failed: ") + cudaGetErrorString(e); return -1; }
    return 1;
}

// Batch-1 Gemma prefill (256 < T <= 288 tokens: the accumulator of a 128-row weight tile is 288 TMEM columns, so one
// tile per SM at a time): persistent CTA pairs, two chunks of T/2 tokens, split-K slices as extra tiles.  Measured on
// B200 on the gate/up shape (tools/cta_timeline.py): main loop per tile 11.5 us against 15.4 us for one CTA per tile.
// In the step (in-graph trace, same box): gate/up 46.3 -> 38.6 us, down 19.8 -> 19.6, o 11.3 -> 14.6, qkv 8.3 -> 8.9 —
// the split-K GEMMs run one short tile per CTA, where the pair's longer set-up (cluster barrier, 2-SM TMEM
// allocation) costs more than its main loop saves; once they were given token chunks (Run::plan_partial, bn_override != 0:
// not served here) only down is left, where pairs win 2 us per layer.  1 = GeGLU only, 2 (default) = every epilogue, 0 = never.
static int g_pair_small = 2;
void gemm_set_pair_small(int mode) { g_pair_small = mode < 0 ? 0 : mode; }
static bool gemm_pair_small_applies(const GemmCall& c) {
    if (g_pair_small == 0 || (g_pair_small == 1 && c.epi != EPI_GEGLU)) return false;
    return c.w_packed && c.bn_override == 0 && c.T > 256 && c.T <= 288 && c.Nw % kBlockM == 0 &&
           c.K % kBlockK == 0 && (c.epi != EPI_GEGLU || c.Nw % (2 * kBlockM) == 0) && (c.epi == EPI_PARTIAL || c.splitk <= 1);
}

static int gemm_launch_pair_small(cudaStream_t stream, const GemmCall& c, std::string* err) {
    const int kb_total = c.K / kBlockK;
    int splitk = c.splitk < 1 ? 1 : c.splitk;
    if (splitk > kb_total) splitk = kb_total;
    const int kb_per_split = (kb_total + splitk - 1) / splitk;
    splitk = (kb_total + kb_per_split - 1) / kb_per_split;            // no empty slices
    const int bn = ((c.T + 1) / 2 + 15) / 16 * 16, half = bn / 2, ntok = 2 * bn;
    CUtensorMap tw, txh;
    if (get_tmap(c.W, c.Nw * kb_total, kBlockK, kBlockK, kBlockM, &tw, err)) return -1;
    if (get_tmap(c.X, c.T, c.K, c.ldx, half, &txh, err)) return -1;
    GemmDev d{};
    d.T = c.T; d.bn = bn; d.nt = 2; d.kb_total = kb_total; d.kb_per_split = kb_per_split;
    d.tmem_cols = 512; d.acc_bufs = 1; d.acc_stride = 0; d.Nw = c.Nw;
    d.bias = c.bias; d.out = c.out; d.ldo = c.ldo; d.partial = c.partial;
    d.cluster = 1; d.slice_rows = 0; d.w_packed = 1; d.trace = c.trace; d.w_static = c.w_static; d.glu_act = c.glu_act;
    d.staging_bytes = c.epi == EPI_PARTIAL ? 0 : ntok * kBlockM * 2;      // GeGLU is staged as raw gate / up values (one accumulator buffer)
    const int stage_bytes = kTileABytes + 2 * half * kBlockK * 2;
    d.stages = (kRingBytes - d.staging_bytes) / stage_bytes;
    if (d.stages > kMaxStages) d.stages = kMaxStages;
    if (d.stages > kb_per_split) d.stages = kb_per_split < 2 ? 2 : kb_per_split;
    const int smem = d.stages * stage_bytes + d.staging_bytes + 1024 + 256;
    const int gxp = (c.Nw / kBlockM + 1) / 2;
    const int tiles = gxp * splitk;
    const int n_pairs = tiles < kTargetCtas / 2 ? tiles : kTargetCtas / 2;
    d.l2_policy = 0;                 // weights are streamed once, the token tile is re-read by every pair
    d.band = gxp;
    cudaError_t e;
    switch (c.epi) {
        case EPI_GEGLU:   e = launch_pairp<EPI_GEGLU>(stream, n_pairs, smem, tw, txh, d, gxp, 1, splitk, kGemmThreads + 256); break;
        case EPI_GELU:    e = c.glu_act == 2 ? launch_pairp<EPI_GELU_ERF>(stream, n_pairs, smem, tw, txh, d, gxp, 1, splitk, kGemmThreads + 256) : launch_pairp<EPI_GELU>(stream, n_pairs, smem, tw, txh, d, gxp, 1, splitk, kGemmThreads + 256); break;
        case EPI_PARTIAL: e = launch_pairp<EPI_PARTIAL>(stream, n_pairs, smem, tw, txh, d, gxp, 1, splitk, kGemmThreads + 256); break;
        default:          e = launch_pairp<EPI_STORE>(stream, n_pairs, smem, tw, txh, d, gxp, 1, splitk, kGemmThreads + 256); break;
    }
    if (e != cudaSuccess) { *err = std::string("gemm (persistent pairs, batch 1) launch failed: ") + cudaGetErrorString(e); return -1; }
    return splitk;
}

// 33..288 token rows (a single-sequence prefill, one image): tokens split into `chunks` CTAs per weight tile x K slices, by the cost
// model measured for the Pi-0 prefill (engine.cu Run::plan_partial, tools/sweep_splitk.py): >= 0.25 us per k-block per
// CTA (0.40 when one CTA holds two 144-token chunks), the fp32 tile store, and the consumer's re-read of the slices.
int gemm_plan_chunked_splitk(int T, int Nw, int K, size_t ws_floats, int* bn_override) {
    *bn_override = 0;
    const int tiles = Nw / 128, kb = (K + 63) / 64;
    double best_cost = 1e30;
    int best_s = 1;
    for (int chunks = 1; chunks <= 4; ++chunks) {
        const int bn = ((T + chunks - 1) / chunks + 15) / 16 * 16;
        if (chunks > 1 && bn < 64) break;
        int sl = kTargetCtas / (tiles * chunks);
        if (sl < 1) break;
        if (sl > kb / 2) sl = kb / 2 > 0 ? kb / 2 : 1;
        if (sl > 16) sl = 16;
        while (sl > 1 && static_cast<size_t>(sl) * T * Nw > ws_floats) --sl;
        const int kb_per = (kb + sl - 1) / sl;
        sl = (kb + kb_per - 1) / kb_per;
        const double us_kb = (chunks == 1 && T > 256) ? 0.40 : 0.25;
        const double tile_tokens = chunks == 1 ? (T > 256 ? 288 : (T + 15) / 16 * 16) : bn;
        const double cost = kb_per * us_kb + 128.0 * tile_tokens * 4.0 / 55e3 + static_cast<double>(sl) * T * Nw * 4.0 / 6e6;
        if (cost < best_cost) { best_cost = cost; best_s = sl; *bn_override = chunks == 1 ? 0 : bn; }
    }
    return best_s;
}

int gemm_launch(cudaStream_t stream, const GemmCall& c, std::string* err) {
    if (gemm_pair_small_applies(c)) return gemm_launch_pair_small(stream, c, err);
    if (gemm_pairp_applies(c)) return gemm_launch_pairp(stream, c, err);
    GemmPlan pl = gemm_make_plan(c.T, c.Nw, c.K, c.splitk, c.epi, c.bn_override);
    if (!pl.valid) {
        *err = "gemm_launch: unsupported shape T=" + std::to_string(c.T) + " Nw=" + std::to_string(c.Nw) +
               " K=" + std::to_string(c.K);
        return -1;
    }
    if (c.epi != EPI_PARTIAL && pl.splitk != 1) { *err = "gemm_launch: split-K needs EPI_PARTIAL"; return -1; }
    CUtensorMap tw, tx, txs;
    if (c.w_packed) {
        if (get_tmap(c.W, c.Nw * pl.kb_total, kBlockK, kBlockK, kBlockM, &tw, err)) return -1;
    } else {
        if (get_tmap(c.W, c.Nw, c.K, c.ldw, kBlockM, &tw, err)) return -1;
    }
    if (get_tmap(c.X, c.T, c.K, c.ldx, pl.bn, &tx, err)) return -1;
    if (pl.cluster > 1) {
        if (get_tmap(c.X, c.T, c.K, c.ldx, pl.slice_rows, &txs, err)) return -1;
    } else {
        txs = tx;
    }
    GemmDev d{};
    d.T = c.T; d.bn = pl.bn; d.nt = pl.nt; d.stages = pl.stages; d.kb_total = pl.kb_total;
    d.kb_per_split = pl.kb_per_split; d.tmem_cols = pl.tmem_cols; d.Nw = c.Nw;
    d.bias = c.bias; d.out = c.out; d.ldo = c.ldo; d.partial = c.partial;
    d.cluster = pl.cluster; d.slice_rows = pl.slice_rows; d.w_packed = c.w_packed; d.trace = c.trace; d.w_static = c.w_static; d.glu_act = c.glu_act;
    cudaError_t e;
    if (pl.two_cta) {
        CUtensorMap txh;
        if (get_tmap(c.X, c.T, c.K, c.ldx, pl.bn / 2, &txh, err)) return -1;
        switch (c.epi) {
            case EPI_STORE:   e = launch_epi2<EPI_STORE>(stream, pl, tw, txh, d); break;
            case EPI_GELU:    e = c.glu_act == 2 ? launch_epi2<EPI_GELU_ERF>(stream, pl, tw, txh, d) : launch_epi2<EPI_GELU>(stream, pl, tw, txh, d); break;
            case EPI_GEGLU:   e = launch_epi2<EPI_GEGLU>(stream, pl, tw, txh, d); break;
            case EPI_PARTIAL: e = launch_epi2<EPI_PARTIAL>(stream, pl, tw, txh, d); break;
            default: *err = "gemm_launch: bad epilogue"; return -1;
        }
        if (e != cudaSuccess) { *err = std::string("gemm (2-CTA) launch failed: ") + cudaGetErrorString(e); return -1; }
        return pl.splitk;
    }
    d.acc_bufs = 1; d.acc_stride = 0; d.staging_bytes = 0;
    const bool large_dbuf = (g_large_t_mode == 1 || ((g_large_t_mode < 0 || g_large_t_mode >= 2) && c.epi != EPI_GEGLU && c.epi != EPI_GELU)) &&
                            c.T > 1024 && pl.nt == 1 && pl.tmem_cols <= 256 && pl.cluster == 1 && !pl.two_cta;
    if (large_dbuf) {
        d.acc_bufs = 2; d.acc_stride = pl.tmem_cols; d.tmem_cols = 2 * pl.tmem_cols;
        d.l2_policy = g_pair_policy >= 0 ? (g_pair_policy == 1 ? 1 : 0) : 1;
        // the direct epilogue stages nothing in the ring: use all of it
        const int stage_bytes = kTileABytes + pl.nt * pl.bn * kBlockK * 2;
        if (c.epi != EPI_PARTIAL && (pl.nt * pl.bn) % 32 == 0)       // staged bf16 epilogue (8 epilogue warps)
            d.staging_bytes = pl.nt * pl.bn * (c.epi == EPI_GEGLU ? kBlockM / 2 : kBlockM) * 2;
        int st = (kRingBytes - d.staging_bytes) / stage_bytes;
        if (st > kMaxStages) st = kMaxStages;
        d.stages = st;
        pl.smem_bytes = st * stage_bytes + d.staging_bytes + 1024 + 256;
    }
    if (large_dbuf || (g_persistent && pl.cluster == 1 && pl.nt * pl.bn <= kPersistentMaxTokens && c.epi != EPI_GEGLU)) {
        switch (c.epi) {
            case EPI_STORE:   e = launch_epip<EPI_STORE>(stream, pl, tw, tx, d); break;
            case EPI_GELU:    e = c.glu_act == 2 ? launch_epip<EPI_GELU_ERF>(stream, pl, tw, tx, d) : launch_epip<EPI_GELU>(stream, pl, tw, tx, d); break;
            case EPI_GEGLU:   e = launch_epip<EPI_GEGLU>(stream, pl, tw, tx, d); break;
            case EPI_PARTIAL: e = launch_epip<EPI_PARTIAL>(stream, pl, tw, tx, d); break;
            default: *err = "gemm_launch: bad epilogue"; return -1;
        }
        if (e != cudaSuccess) { *err = std::string("gemm (persistent) launch failed: ") + cudaGetErrorString(e); return -1; }
        return pl.splitk;
    }
    switch (c.epi) {
        case EPI_STORE:   e = launch_epi<EPI_STORE>(stream, pl, tw, tx, txs, d); break;
        case EPI_GELU:    e = c.glu_act == 2 ? launch_epi<EPI_GELU_ERF>(stream, pl, tw, tx, txs, d) : launch_epi<EPI_GELU>(stream, pl, tw, tx, txs, d); break;
        case EPI_GEGLU:   e = launch_epi<EPI_GEGLU>(stream, pl, tw, tx, txs, d); break;
        case EPI_PARTIAL: e = launch_epi<EPI_PARTIAL>(stream, pl, tw, tx, txs, d); break;
        default: *err = "gemm_launch: bad epilogue"; return -1;
    }
    if (e != cudaSuccess) { *err = std::string("gemm launch failed: ") + cudaGetErrorString(e); return -1; }
    return pl.splitk;
}

}  // namespace blurr
