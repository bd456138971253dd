// Operand-ingest microbenchmark for the batch-1 GEMMs (no math): what bounds a CTA that streams
// weight tiles from HBM while re-reading a token tile that every other CTA also reads?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/ingest_probe tools/ingest_probe.cu -lcuda
//   tools/bin/ingest_probe            (runs the built-in sweep, prints one line per case)
//
// Every CTA runs a TMA producer thread and a consumer thread joined by a full/empty mbarrier ring;
// the consumer releases a stage as soon as it has landed.  Per k-block a CTA loads
//   * `wbytes` of weights: contiguous 16 KB boxes from a region of its own inside a buffer larger
//     than L2 (=> DRAM), or inside a small buffer (=> L2 hits on distinct addresses);
//   * `xrows` x 128 B of a token tile that all CTAs share (=> L2 hits on the same addresses), either
//     unicast, or multicast across a cluster (each CTA loads 1/cluster of the rows for everyone).
// Reported: time per launch, aggregate GB/s delivered into shared memory (what the SMs ingest) and
// the DRAM-side GB/s (weights only).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(ra) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) if (clock64() - t0 > 200000000LL) return false;
    return true;
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
                 " [%0], [%1, {%4, %5}], [%2], %3, %6;\n"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

struct Params {
    int kblocks;        // k-blocks per CTA
    int wboxes;         // 16 KB weight boxes per k-block (0 = none)
    int xrows;          // shared token-tile rows per k-block (0 = none), 128 B each
    int stages;
    int cluster;        // 1 = unicast; >1 = the x tile is multicast, each CTA loads xrows/cluster rows
    int w_rows_per_cta; // rows (of 128 B) of the weight tensor owned by one CTA
    int w_wrap_rows;    // the weight tensor has this many rows in total (CTA regions wrap around it)
    int x_kcols;        // number of 64-element column blocks of the x tensor (k-blocks cycle through them)
    int pol_w, pol_x;   // 0 evict_first 1 normal 2 evict_last
    int stagger;        // 1: CTA b starts its walk over the x column blocks at block b (no two CTAs on the same lines at once)
    int* err;
};

__device__ __forceinline__ uint64_t policy(int which) {
    uint64_t p;
    if (which == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    else if (which == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = p.wboxes * 16384 + p.xrows * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + p.stages * stage_bytes);
    uint64_t* empty = full + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = p.cluster > 1 ? cluster_rank() : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], p.cluster); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (p.cluster > 1) cluster_sync();
    if (warp == 0 && lane == 0) {
        const uint64_t pw = policy(p.pol_w), px = policy(p.pol_x);
        const uint16_t mask = static_cast<uint16_t>((1u << p.cluster) - 1u);
        const int slice = p.cluster > 1 ? p.xrows / p.cluster : p.xrows;
        // no integer divisions in the loops: a single thread pays ~150 cycles for each
        int wrow = static_cast<int>((static_cast<long long>(blockIdx.x) * p.w_rows_per_cta) % p.w_wrap_rows);
        int kcb = (p.stagger ? static_cast<int>(blockIdx.x / p.cluster) : 0) % p.x_kcols;
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < p.kblocks; ++i) {
            if (!mbar_wait(&empty[s], ph ^ 1u)) { atomicExch(p.err, 1); break; }
            uint8_t* stg = ring + s * stage_bytes;
            mbar_expect_tx(&full[s], stage_bytes);
            for (int b = 0; b < p.wboxes; ++b) {
                tma_2d(stg + b * 16384, &tmap_w, &full[s], 0, wrow, pw);
                wrow += 128;
                if (wrow >= p.w_wrap_rows) wrow -= p.w_wrap_rows;
            }
            if (p.xrows > 0) {
                const int kc = kcb * 64;
                if (p.cluster > 1) tma_2d_mc(stg + p.wboxes * 16384 + crank * slice * 128, &tmap_x, &full[s], kc, crank * slice, mask, px);
                else if (p.xrows > 256) {      // a TMA box holds at most 256 rows: two halves
                    tma_2d(stg + p.wboxes * 16384, &tmap_x, &full[s], kc, 0, px);
                    tma_2d(stg + p.wboxes * 16384 + (p.xrows / 2) * 128, &tmap_x, &full[s], kc, p.xrows / 2, px);
                } else tma_2d(stg + p.wboxes * 16384, &tmap_x, &full[s], kc, 0, px);
                if (++kcb == p.x_kcols) kcb = 0;
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 && lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < p.kblocks; ++i) {
            if (!mbar_wait(&full[s], ph)) { atomicExch(p.err, 2); break; }
            if (p.cluster > 1) { for (int r = 0; r < p.cluster; ++r) mbar_arrive_remote(&empty[s], r); }
            else mbar_arrive(&empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    if (p.cluster > 1) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static CUtensorMap make_map(void* ptr, long long rows, int cols, long long ld_elems, int box_rows) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld_elems) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(1); }
    return m;
}

struct Case { const char* name; int grid, kblocks, wboxes, xrows, stages, cluster; bool w_l2; int pol_w, pol_x; int stagger = 0; };

int main(int argc, char** argv) {
    CK(cudaSetDevice(0));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    g_encode = reinterpret_cast<EncodeTiledFn>(fp);
    const size_t wbytes = 2ull << 30;                       // 2 GiB of "weights" (>> L2)
    uint8_t* W; CK(cudaMalloc(&W, wbytes)); CK(cudaMemset(W, 1, wbytes));
    const int XR = 288, XK = 2048;                          // token tile [288][2048] bf16 = 1.13 MB
    uint8_t* X; CK(cudaMalloc(&X, size_t(XR) * XK * 2)); CK(cudaMemset(X, 2, size_t(XR) * XK * 2));
    int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const long long w_rows_total = wbytes / 128;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    std::vector<Case> cases = {
        // name                          grid  kb  wb  xrows st cl  wl2  polw polx
        {"W only 16K x4st",              148,  64, 1,   0,   4, 1, false, 0, 2},
        {"W only 16K x8st",              148,  64, 1,   0,   8, 1, false, 0, 2},
        {"W only 16K x12st",             148,  64, 1,   0,  12, 1, false, 0, 2},
        {"W only 32K x6st",              148,  32, 2,   0,   6, 1, false, 0, 2},
        {"W only 16K x8st normal-pol",   148,  64, 1,   0,   8, 1, false, 1, 1},
        {"W only 16K x4st 296 CTAs",     296,  32, 1,   0,   4, 1, false, 0, 2},
        {"W only 16K x4st grid 32",       32,  64, 1,   0,   4, 1, false, 0, 2},
        {"W only 16K x12st grid 32",      32,  64, 1,   0,  12, 1, false, 0, 2},
        {"W(L2) only 16K x4st",          148,  64, 1,   0,   4, 1, true,  1, 1},
        {"W(L2) only 16K x12st",         148,  64, 1,   0,  12, 1, true,  1, 1},
        {"W(L2) only 16K x12st grid 32",  32,  64, 1,   0,  12, 1, true,  1, 1},
        {"X only 288r x4st",             148,  64, 0, 288,   4, 1, false, 0, 2},
        {"X only 288r x6st",             148,  64, 0, 288,   6, 1, false, 0, 2},
        {"X only 288r x6st grid 32",      32,  64, 0, 288,   6, 1, false, 0, 2},
        {"X only 144r x8st",             148,  64, 0, 144,   8, 1, false, 0, 2},
        {"X only 64r x12st",             148,  64, 0,  64,  12, 1, false, 0, 2},
        {"W+X 288r x4st (gate/up now)",  148,  64, 1, 288,   4, 1, false, 0, 2},
        {"W+X 288r x4st 256 CTAs",       256,  32, 1, 288,   4, 1, false, 0, 2},
        {"W+X 144r x6st (pair share)",   148,  64, 1, 144,   6, 1, false, 0, 2},
        {"W+X 72r x8st",                 148,  64, 1,  72,   8, 1, false, 0, 2},
        {"W+X 64r x8st (siglip qkv)",    108,  18, 1,  64,   8, 1, false, 0, 2},
        {"2W+X 288r x3st (wide)",        128,  32, 2, 288,   3, 1, false, 0, 2},
        {"W+X mc2 288r x4st",            148,  64, 1, 288,   4, 2, false, 0, 2},
        {"W+X mc4 288r x4st",            148,  64, 1, 288,   4, 4, false, 0, 2},
        {"W+X mc8 288r x4st",            144,  64, 1, 288,   4, 8, false, 0, 2},
        {"X only mc2 288r x6st",         148,  64, 0, 288,   6, 2, false, 0, 2},
        {"X only mc4 288r x6st",         148,  64, 0, 288,   6, 4, false, 0, 2},
        {"X only mc8 288r x6st",         144,  64, 0, 288,   6, 8, false, 0, 2},
        {"X only 288r x6st stagger",     148,  64, 0, 288,   6, 1, false, 0, 2, 1},
        {"W+X 288r x4st stagger",        148,  64, 1, 288,   4, 1, false, 0, 2, 1},
        {"W+X mc4 288r x4st stagger",    148,  64, 1, 288,   4, 4, false, 0, 2, 1},
    };
    const char* only = argc > 1 ? argv[1] : nullptr;
    printf("%-34s %5s %4s %3s %5s %3s %3s | %8s %8s %10s %10s %8s\n", "case", "grid", "kb", "wb", "xrows", "st", "cl", "us@kb", "us@4kb", "ingestGB/s", "dram GB/s", "B/clk/SM");
    for (const Case& c : cases) {
        if (only && !strstr(c.name, only)) continue;
        double us_at[2] = {0, 0};
        const int stage_bytes = c.wboxes * 16384 + c.xrows * 128;
        const size_t smem = size_t(c.stages) * stage_bytes + 1024 + 512;
        if (smem > 227 * 1024) { printf("%-34s skipped (smem %zu)\n", c.name, smem); continue; }
        if (c.cluster > 1 && (c.xrows % (8 * c.cluster)) != 0) { printf("%-34s skipped (rows %% 8*cluster)\n", c.name); continue; }
        for (int pass = 0; pass < 2; ++pass) {
            const int kblocks = c.kblocks * (pass == 0 ? 1 : 4);
            Params p{};
            p.kblocks = kblocks; p.wboxes = c.wboxes; p.xrows = c.xrows; p.stages = c.stages; p.cluster = c.cluster;
            p.w_rows_per_cta = kblocks * c.wboxes * 128;
            p.x_kcols = XK / 64; p.pol_w = c.pol_w; p.pol_x = c.pol_x; p.err = err; p.stagger = c.stagger;
            const int xbox = c.cluster > 1 ? c.xrows / c.cluster : (c.xrows > 256 ? c.xrows / 2 : (c.xrows > 0 ? c.xrows : 8));
            CUtensorMap tx = make_map(X, XR, XK, XK, xbox);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
            cudaLaunchAttribute at[1]; int na = 0;
            if (c.cluster > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = c.cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
            cfg.attrs = at; cfg.numAttrs = na;
            const int reps = 15;
            std::vector<float> ts;
            const long long span = static_cast<long long>(c.grid) * p.w_rows_per_cta;     // rows touched per launch
            for (int r = 0; r < reps + 3; ++r) {
                Params pr = p;
                CUtensorMap twr;
                if (c.w_l2) {
                    // everything inside 48 MB: L2 hits on distinct addresses
                    twr = make_map(W, (48ll << 20) / 128, 64, 64, 128);
                    pr.w_wrap_rows = static_cast<int>((48ll << 20) / 128);
                } else {
                    // a fresh window of the 2 GiB buffer every launch: DRAM
                    const long long windows = std::max(1ll, w_rows_total / (span + 128));
                    const long long shift_rows = (r % windows) * span;
                    twr = make_map(W + shift_rows * 128, w_rows_total - shift_rows, 64, 64, 128);
                    pr.w_wrap_rows = static_cast<int>(w_rows_total - shift_rows);
                }
                CK(cudaEventRecord(e0));
                CK(cudaLaunchKernelEx(&cfg, probe_kernel, twr, tx, pr));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 3) ts.push_back(ms);
            }
            std::sort(ts.begin(), ts.end());
            us_at[pass] = ts[ts.size() / 2] * 1e3;
        }
        // slope between kb and 4 kb: the steady-state rate without launch / fill / drain
        const double dus = us_at[1] - us_at[0];
        const double ingest = double(c.grid) * (3.0 * c.kblocks) * stage_bytes;
        const double dram = c.w_l2 ? 0.0 : double(c.grid) * (3.0 * c.kblocks) * c.wboxes * 16384;
        const int active = std::min(c.grid, 148);
        printf("%-34s %5d %4d %3d %5d %3d %3d | %8.2f %8.2f %10.0f %10.0f %8.1f\n", c.name, c.grid, c.kblocks, c.wboxes, c.xrows, c.stages, c.cluster,
               us_at[0], us_at[1], ingest / dus * 1e-3, dram / dus * 1e-3, ingest / active / (dus * 1e-6) / 1.9e9);
        int herr = 0; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
        if (herr) { printf("  !! pipeline wait timed out (code %d)\n", herr); CK(cudaMemset(err, 0, 4)); }
    }
    return 0;
}
