// tcgen05.mma issue-rate microbenchmark for the batch-1 GEMM main loop (no epilogue).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_probe tools/mma_probe.cu -lcuda
//
// Each CTA: warp 0 = TMA producer (one 16 KB weight box + the token rows per k-block, as in
// gemm_tile), warp 1 = MMA issuer (M = 128, K = 64 per k-block as 4 x K16, the token rows split
// into UMMA-N chunks given on the case line).  Modes:
//   mma   : operands static in shared memory, no TMA at all  -> pure tensor-pipe rate per k-block
//   both  : the real pipeline (TMA ring feeding the MMAs)
//   tma   : TMA ring only, the MMA warp releases stages without issuing math
// The slope between `kb` and 4 x `kb` k-blocks gives the steady-state time per k-block.
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
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n) {
    uint32_t d = 0;
    d |= 1u << 4; d |= 1u << 7; d |= 1u << 10; d |= (n >> 3) << 17; d |= (m >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

struct Params {
    int kblocks, stages, mode;      // mode 0 mma, 1 both, 2 tma
    int order;                      // 0: chunk-major (c: k0..k3); 1: k-major (k: c0, c1, ..); 2: one chunk, even/odd k16 steps on two accumulators
    int nchunks, chunk[4];          // UMMA N of each chunk (token rows)
    int xrows;                      // sum of chunks
    int w_wrap_rows, w_rows_per_cta;
    int* err;
    long long* cycles;              // per CTA: MMA-issuer cycles from first wait to final commit completion
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_x2, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = 16384 + p.xrows * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + p.stages * stage_bytes);
    uint64_t* empty = full + 16;
    uint64_t* done = empty + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp == 0 && p.mode != 0) {
        uint64_t pw, px;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pw));
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(px));
        int wrow = static_cast<int>((static_cast<long long>(blockIdx.x) * p.w_rows_per_cta) % p.w_wrap_rows);
        int s = 0, kcb = 0; uint32_t ph = 0;
        for (int i = 0; i < p.kblocks; ++i) {
            if (!mbar_wait(&empty[s], ph ^ 1u)) { atomicExch(p.err, 1); break; }
            uint8_t* stg = ring + s * stage_bytes;
            if (elect_one()) {
                mbar_expect_tx(&full[s], stage_bytes);
                tma_2d(stg, &tmap_w, &full[s], 0, wrow, pw);
                int r0 = 0;
                for (int c = 0; c < p.nchunks; ++c) {
                    tma_2d(stg + 16384 + r0 * 128, c == 0 ? &tmap_x : &tmap_x2, &full[s], kcb * 64, r0, px);
                    r0 += p.chunk[c];
                }
            }
            __syncwarp();
            wrow += 128; if (wrow >= p.w_wrap_rows) wrow -= p.w_wrap_rows;
            if (++kcb == 32) kcb = 0;
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // Whole warp converged, every operand warp-uniform; only the tcgen05 instructions sit under elect.sync.
        // (Issuing from an `if (lane == 0)` region makes ptxas wrap EVERY UTCHMMA / UTMALDG in an
        // ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY loop: ~110-160 cycles per instruction.)
        uint32_t idesc[4];
        for (int c = 0; c < 4; ++c) idesc[c] = make_idesc(128, p.chunk[c] > 0 ? p.chunk[c] : 16);
        int s = 0; uint32_t ph = 0;
        const long long t0 = clock64();
        for (int i = 0; i < p.kblocks; ++i) {
            if (p.mode != 0) {
                if (!mbar_wait(&full[s], ph)) { atomicExch(p.err, 2); break; }
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            if (elect_one()) {
                if (p.mode != 2) {
                    const uint32_t a_addr = smem_u32(ring + s * stage_bytes);
                    const uint64_t da = make_desc(a_addr);
                    if (p.order == 0) {
                        int r0 = 0, col = 0;
                        for (int c = 0; c < p.nchunks; ++c) {
                            const uint64_t db = make_desc(a_addr + 16384 + r0 * 128);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma(tmem + col, da + 2 * k, db + 2 * k, idesc[c], (i > 0 || k > 0) ? 1u : 0u);
                            r0 += p.chunk[c]; col += p.chunk[c];
                        }
                    } else if (p.order == 1) {
                        uint64_t db[4]; int col[4]; int r0 = 0, cc = 0;
                        for (int c = 0; c < 4; ++c) { db[c] = make_desc(a_addr + 16384 + r0 * 128); col[c] = cc; r0 += p.chunk[c]; cc += p.chunk[c]; }
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            for (int c = 0; c < p.nchunks; ++c) umma(tmem + col[c], da + 2 * k, db[c] + 2 * k, idesc[c], (i > 0 || k > 0) ? 1u : 0u);
                    } else {
                        const uint64_t db = make_desc(a_addr + 16384);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma(tmem + (k & 1) * 256, da + 2 * k, db + 2 * k, idesc[0], (i > 0 || k > 1) ? 1u : 0u);
                    }
                }
                if (p.mode != 0) {
                    if (p.mode == 2) mbar_arrive(&empty[s]); else umma_commit(&empty[s]);
                }
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (p.mode != 2) {
            if (elect_one()) umma_commit(done);
            __syncwarp();
            if (!mbar_wait(done, 0)) atomicExch(p.err, 3);
        }
        if (lane == 0) p.cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
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

struct Case { const char* name; int grid, kblocks, stages, mode; std::vector<int> chunks; int order = 0; };

int main(int argc, char** argv) {
    CK(cudaSetDevice(0));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    g_encode = reinterpret_cast<EncodeTiledFn>(fp);
    const size_t wbytes = 2ull << 30;
    uint8_t* W; CK(cudaMalloc(&W, wbytes)); CK(cudaMemset(W, 0, wbytes));
    const int XR = 512, XK = 2048;
    uint8_t* X; CK(cudaMalloc(&X, size_t(XR) * XK * 2)); CK(cudaMemset(X, 0, size_t(XR) * XK * 2));
    int* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
    long long* cyc; CK(cudaMalloc(&cyc, 8 * 1024)); CK(cudaMemset(cyc, 0, 8 * 1024));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const long long w_rows_total = wbytes / 128;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<Case> cases;
    const char* mname[3] = {"mma ", "both", "tma "};
    // pure MMA rate for every candidate chunking of 276..288 tokens and the small-N shapes
    std::vector<std::vector<int>> chunkings = {{144, 144}, {256, 32}, {256}, {128, 128, 32}, {192, 96}, {96, 96, 96}, {128}, {64}, {32}, {16},
                                               {208, 80}, {224, 64}, {240, 48}, {160, 128}, {256, 16}};
    for (auto& ch : chunkings) cases.push_back({"", 148, 32, 4, 0, ch});
    for (auto& ch : std::vector<std::vector<int>>{{144, 144}, {256, 32}, {256}, {128}, {64}}) cases.push_back({"", 148, 32, 4, 1, ch});
    for (auto& ch : std::vector<std::vector<int>>{{144, 144}, {64}}) cases.push_back({"", 148, 32, 4, 2, ch});
    cases.push_back({"", 148, 32, 3, 1, {144, 144}});
    cases.push_back({"", 1, 32, 4, 0, {144, 144}});
    cases.push_back({"", 1, 32, 4, 0, {256}});
    for (auto& ch : std::vector<std::vector<int>>{{144, 144}, {256, 32}, {96, 96, 96}, {128, 128}, {64, 64}, {32, 32}}) cases.push_back({"k-major", 148, 32, 4, 0, ch, 1});
    for (auto& ch : std::vector<std::vector<int>>{{128}, {64}, {32}, {16}, {256}}) cases.push_back({"2acc", 148, 32, 4, 0, ch, 2});
    cases.push_back({"k-major", 148, 32, 4, 1, {144, 144}, 1});
    cases.push_back({"k-major", 148, 32, 3, 1, {144, 144}, 1});
    cases.push_back({"2acc", 148, 32, 8, 1, {64}, 2});
    cases.push_back({"2acc", 148, 32, 6, 1, {128}, 2});
    printf("%-5s %-16s %4s %3s | %8s %8s %10s %10s %9s | %s\n", "mode", "chunks", "grid", "st", "us@kb", "us@4kb", "ns/kblock", "cyc/kblock", "dramGB/s", "issuer cycles/kblock (clock64, CTA 0)");
    for (const Case& c : cases) {
        int xrows = 0; for (int v : c.chunks) xrows += v;
        const int stage_bytes = 16384 + xrows * 128;
        const size_t smem = size_t(c.stages) * stage_bytes + 1024 + 512;
        if (smem > 227 * 1024) { printf("skipped (smem)\n"); continue; }
        bool equal = true; for (int v : c.chunks) equal &= (v == c.chunks[0]);
        if (c.mode != 0 && !equal && c.chunks.size() > 1) {
            // one tensor map box per launch: unequal chunks load with a box of the gcd rows... keep it simple: box = chunk[0], extra rows land past the chunk (harmless: same stage)
        }
        double us_at[2]; long long cyc_at[2] = {0, 0};
        for (int pass = 0; pass < 2; ++pass) {
            Params p{};
            p.kblocks = c.kblocks * (pass ? 4 : 1); p.stages = c.stages; p.mode = c.mode; p.order = c.order;
            p.nchunks = (int)c.chunks.size(); for (int i = 0; i < 4; ++i) p.chunk[i] = i < p.nchunks ? c.chunks[i] : 0;
            p.xrows = xrows; p.err = err; p.cycles = cyc;
            p.w_rows_per_cta = p.kblocks * 128;
            // TMA of the x rows: use a box of `chunk` rows when all chunks are equal, else 16-row boxes would be too many ops;
            // for unequal chunkings in `both` mode load the x rows as nchunks boxes of chunk[0] rows clipped by the stage size
            CUtensorMap tx = make_map(X, XR, XK, XK, c.chunks[0]);
            CUtensorMap tx2 = make_map(X, XR, XK, XK, c.chunks.size() > 1 ? c.chunks[1] : c.chunks[0]);
            if (c.mode != 0 && !equal) {
                // load as equal halves of the total instead (same bytes, same op count as the chunk list)
                p.nchunks = (int)c.chunks.size();
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
            std::vector<float> ts;
            const long long span = static_cast<long long>(c.grid) * p.w_rows_per_cta;
            for (int r = 0; r < 13; ++r) {
                Params pr = p;
                const long long windows = std::max(1ll, w_rows_total / (span + 128));
                const long long shift_rows = (r % windows) * span;
                CUtensorMap twr = make_map(W + shift_rows * 128, w_rows_total - shift_rows, 64, 64, 128);
                pr.w_wrap_rows = static_cast<int>(w_rows_total - shift_rows);
                CK(cudaEventRecord(e0));
                CK(cudaLaunchKernelEx(&cfg, probe_kernel, twr, tx, tx2, pr));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r >= 3) ts.push_back(ms);
            }
            std::sort(ts.begin(), ts.end());
            us_at[pass] = ts[ts.size() / 2] * 1e3;
            CK(cudaMemcpy(&cyc_at[pass], cyc, 8, cudaMemcpyDeviceToHost));
        }
        char chs[64] = ""; for (int v : c.chunks) { char t[16]; snprintf(t, sizeof t, "%d ", v); strcat(chs, t); }
        const double dus = us_at[1] - us_at[0];
        const double nkb = 3.0 * c.kblocks;
        printf("%-5s %-8s %-16s %4d %3d | %8.2f %8.2f %10.1f %10.1f %9.0f | %.1f\n", mname[c.mode], c.name, chs, c.grid, c.stages, us_at[0], us_at[1],
               dus / nkb * 1e3, dus / nkb * 1e-6 * 1.965e9, c.mode == 0 ? 0.0 : double(c.grid) * nkb * 16384 / dus * 1e-3,
               double(cyc_at[1] - cyc_at[0]) / nkb);
        int herr = 0; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
        if (herr) { printf("  !! wait timed out (code %d)\n", herr); CK(cudaMemset(err, 0, 4)); }
    }
    return 0;
}
