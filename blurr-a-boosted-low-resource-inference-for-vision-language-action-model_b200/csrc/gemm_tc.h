// Host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <string>

namespace blurr {

enum GemmEpilogue {
    EPI_STORE = 0,    // out[t][n] = bf16(acc + bias[n])
    EPI_GELU = 1,     // out[t][n] = bf16(gelu_tanh(bf16(acc + bias[n])))
    EPI_GEGLU = 2,    // weight rows interleaved 64 gate / 64 up per 128-row tile; out width Nw/2
    EPI_PARTIAL = 3,  // partial[z][t][n] = fp32 partial sum of split-K slice z
};

struct GemmCall {
    const __nv_bfloat16* W;   // [Nw][K], row stride ldw
    int Nw, K, ldw;
    const __nv_bfloat16* X;   // [T][K], row stride ldx
    int T, ldx;
    int epi;
    int splitk;               // >= 1 (EPI_PARTIAL only when > 1)
    const __nv_bfloat16* bias;
    __nv_bfloat16* out;
    int ldo;
    float* partial;
    int bn_override;          // 0 = automatic token chunking
};

struct GemmPlan {
    bool valid;
    int bn, nt, stages, kb_total, kb_per_split, splitk, tmem_cols, smem_bytes, grid_x, grid_y;
};

GemmPlan gemm_make_plan(int T, int Nw, int K, int splitk, int epi, int bn_override);

// Returns the number of split-K slices actually used (>= 1), or -1 with *err set.
int gemm_launch(cudaStream_t stream, const GemmCall& call, std::string* err);

// Non-zero if a pipeline wait of a GEMM kernel expired since the last call (synchronises the
// device through cudaMemcpyFromSymbol); clears the flag.
int gemm_take_timeout_flag();

// Tensor maps are cached by (pointer, shape); call when buffers are freed.
void gemm_forget_tensor_maps();

}  // namespace blurr
