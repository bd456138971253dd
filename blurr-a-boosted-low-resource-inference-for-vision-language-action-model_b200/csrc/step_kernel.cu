// The persistent cooperative step kernel (see step_kernel.h).
#include "step_kernel.h"

#include "bodies.cuh"
#include "gemm_body.cuh"

namespace blurr {

static constexpr int kStepThreads = 256;
static constexpr int kStepRingBytes = 225 * 1024;                    // = kRingBytes of gemm_tc.cu
static constexpr int kStepSmemBytes = kStepRingBytes + 1024 + 256;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// All CTAs of the (co-resident) grid meet here.  Generic-proxy writes made before the barrier are
// visible afterwards to generic loads (L2) and to TMA (async proxy) reads of every CTA.
__device__ __forceinline__ bool grid_barrier(unsigned* sync, unsigned target) {
    __shared__ int s_ok;
    fence_proxy_async_all();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&sync[0], 1u);
        int ok = 1;
        const long long t0 = clock64();
        while (ld_acquire_gpu(&sync[0]) < target) {
            if (ld_acquire_gpu(&sync[1]) != 0u) { ok = 0; break; }            // another CTA gave up
            if (clock64() - t0 > 2000000000LL) { atomicExch(&sync[1], 1u); ok = 0; break; }
        }
        __threadfence();
        s_ok = ok;
    }
    __syncthreads();
    fence_proxy_async_all();
    return s_ok != 0;
}

template <int HD_PAD, bool GEMMA>
__device__ __forceinline__ void attn_step_tile(const AttnMmaArgs& a, uint8_t* smem, int bx, int by, int bz) {
    attn_mma_body<HD_PAD, kAttnTileRows, GEMMA>(a, smem, bx, by, bz);
}

__global__ void __launch_bounds__(kStepThreads, 1) step_kernel(const StepOp* __restrict__ ops, const int n_ops,
                                                               unsigned* sync) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ StepOpHot hot;
    uint32_t tmem_base;
    GemmShared sh = gemm_setup_shared(smem_raw, kStepRingBytes, 1, 512u, &tmem_base);
    GemmPipe pipe;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    unsigned target = 0;
    bool alive = true;

    for (int oi = 0; oi < n_ops && alive; ++oi) {
        // stage the op's arguments in shared memory (one coalesced copy instead of repeated L2 reads)
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&ops[oi].hot);
            uint32_t* dst = reinterpret_cast<uint32_t*>(&hot);
            for (int i = threadIdx.x; i < static_cast<int>(sizeof(StepOpHot) / 4); i += kStepThreads) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        const int gx = hot.gx, gy = hot.gy, items = hot.gx * hot.gy * hot.gz;
        const int type = hot.type;
        // inside a group of independent ops the first CTA rotates so the groups' items spread evenly
        const int first = hot.pad0 % static_cast<int>(gridDim.x);
        for (int item = (static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - first) % static_cast<int>(gridDim.x);
             item < items; item += gridDim.x) {
            const int bx = item % gx, by = (item / gx) % gy, bz = item / (gx * gy);
            switch (type) {
                case OP_GEMM: {
                    const CUtensorMap* tw = &ops[oi].tmap_w;
                    const CUtensorMap* tx = &ops[oi].tmap_x;
                    switch (hot.epi) {
                        case EPI_STORE: gemm_tile<EPI_STORE>(hot.u.gemm, tw, tx, tx, sh, pipe, bx, by, bz, 0u); break;
                        case EPI_GELU: gemm_tile<EPI_GELU>(hot.u.gemm, tw, tx, tx, sh, pipe, bx, by, bz, 0u); break;
                        case EPI_GEGLU: gemm_tile<EPI_GEGLU>(hot.u.gemm, tw, tx, tx, sh, pipe, bx, by, bz, 0u); break;
                        default: gemm_tile<EPI_PARTIAL>(hot.u.gemm, tw, tx, tx, sh, pipe, bx, by, bz, 0u); break;
                    }
                    break;
                }
                case OP_CONSUMER:
                    if (hot.u.consumer.N <= 4 * kRowThreads) consumer_body<1>(hot.u.consumer, bx);
                    else consumer_body<2>(hot.u.consumer, bx);
                    break;
                case OP_BIAS_ACT: {
                    const BiasActArgs& a = hot.u.bias_act;
                    bias_act_body(a.partial, a.splitk, a.T, a.N, a.ldp, a.bias, a.act, a.scale, a.out, a.ldo, bx);
                    break;
                }
                case OP_ROPE_KV: rope_kv_body<false>(hot.u.rope, bx); break;
                case OP_ATTN_SIGLIP: attn_step_tile<80, false>(hot.u.attn, sh.ring, bx, by, bz); break;
                case OP_ATTN_PREFILL: attn_step_tile<256, true>(hot.u.attn, sh.ring, bx, by, bz); break;
                case OP_ATTN_FEWQ: attn_step_tile<256, true>(hot.u.attn, sh.ring, bx, by, bz); break;
                case OP_EMBED_MERGE: {
                    const EmbedMergeArgs& a = hot.u.embed;
                    embed_merge_body(a.ids, a.seq, a.table, a.vocab, a.img, a.n_img, a.hidden, a.image_token, a.pad_token,
                                     a.inv_div, a.normalizer, a.out, a.err_flag, bx, by);
                    break;
                }
                case OP_SMALL_K: {
                    const SmallKArgs& a = hot.u.small_k;
                    small_k_linear_body(a.x, a.T, a.K, a.W, a.bias, a.N, a.scale, a.y, a.ldy, a.col_off, a.time_row,
                                        a.time_cols, bx, by);
                    break;
                }
                case OP_ACTION_TAIL: {
                    const ActionTailArgs& a = hot.u.tail;
                    action_tail_body(a.xn, a.T, a.hidden, a.W, a.bias, a.action_dim, a.dt, a.action, a.vel_tap, bx);
                    break;
                }
                case OP_CLAMP: {
                    const ClampArgs& a = hot.u.clamp;
                    clamp_copy_body(a.src, a.dst, a.n, a.do_clamp, a.clip, bx);
                    break;
                }
                default: break;
            }
            __syncthreads();      // shared scratch is reused by the next work item
        }
        if (hot.barrier_after) {
            target += gridDim.x;
            alive = grid_barrier(sync, target);
        } else {
            __syncthreads();      // `hot` is rewritten by the next op
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512u);
}

int step_program_upload(StepProgram& prog, std::string* err) {
    const size_t bytes = prog.ops.size() * sizeof(StepOp);
    if (prog.d_capacity < bytes) {
        if (prog.d_ops) cudaFree(prog.d_ops);
        if (cudaMalloc(&prog.d_ops, bytes) != cudaSuccess) { *err = "step program: cudaMalloc failed"; return -1; }
        prog.d_capacity = bytes;
    }
    if (!prog.d_sync) {
        if (cudaMalloc(&prog.d_sync, 64) != cudaSuccess) { *err = "step program: cudaMalloc failed"; return -1; }
        cudaMemset(prog.d_sync, 0, 64);
    }
    if (cudaMemcpy(prog.d_ops, prog.ops.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        *err = "step program: upload failed";
        return -1;
    }
    prog.n_barriers = 0;
    for (const auto& op : prog.ops) prog.n_barriers += op.hot.barrier_after ? 1 : 0;
    return 0;
}

void step_program_free(StepProgram& prog) {
    if (prog.d_ops) cudaFree(prog.d_ops);
    if (prog.d_sync) cudaFree(prog.d_sync);
    prog.d_ops = nullptr; prog.d_sync = nullptr; prog.d_capacity = 0;
    prog.ops.clear();
}

int step_program_launch(const StepProgram& prog, cudaStream_t stream, std::string* err) {
    static int grid = 0;
    if (grid == 0) {
        cudaError_t e = cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStepSmemBytes);
        if (e != cudaSuccess) { *err = std::string("step kernel smem attribute: ") + cudaGetErrorString(e); return -1; }
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_kernel, kStepThreads, kStepSmemBytes);
        if (e != cudaSuccess || per_sm < 1) { *err = "step kernel does not fit on an SM"; return -1; }
        grid = sms;                      // one persistent CTA per SM
    }
    if (cudaMemsetAsync(prog.d_sync, 0, 4, stream) != cudaSuccess) { *err = "step kernel: memset failed"; return -1; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kStepThreads);
    cfg.dynamicSmemBytes = kStepSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const StepOp* d_ops = prog.d_ops;
    const int n_ops = static_cast<int>(prog.ops.size());
    unsigned* d_sync = prog.d_sync;
    cudaError_t e = cudaLaunchKernelEx(&cfg, step_kernel, d_ops, n_ops, d_sync);
    if (e != cudaSuccess) { *err = std::string("step kernel launch: ") + cudaGetErrorString(e); return -1; }
    return 0;
}

int step_program_take_error(const StepProgram& prog) {
    if (!prog.d_sync) return 0;
    unsigned v[2] = {0, 0};
    if (cudaMemcpy(v, prog.d_sync, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    int gemm_flag = 0;
    cudaMemcpyFromSymbol(&gemm_flag, g_gemm_timeout_flag, sizeof(int));
    if (v[1] != 0 || gemm_flag != 0) {
        cudaMemset(prog.d_sync, 0, 8);
        const int zero = 0;
        cudaMemcpyToSymbol(g_gemm_timeout_flag, &zero, sizeof(int));
        return v[1] != 0 ? 1 : 10 + gemm_flag;
    }
    return 0;
}

}  // namespace blurr
