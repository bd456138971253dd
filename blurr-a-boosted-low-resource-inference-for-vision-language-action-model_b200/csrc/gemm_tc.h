// Host interface of the tcgen05 GEMM (gemm_tc.cu).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <string>

namespace blurr {

enum GemmEpilogue {
    EPI_STORE = 0,    // out[t][n] = bf16(acc + bias[n])
    EPI_GELU = 1,     // out[t][n] = bf16(gelu_tanh(bf16(acc + bias[n])))
    EPI_GEGLU = 2,    // weight rows interleaved 64 gate / 64 up per 128-row tile; out width Nw/2
    EPI_PARTIAL = 3,  // partial[z][t][n] = fp32 partial sum of split-K slice z
    EPI_GELU_ERF = 4, // EPI_GELU with the exact erf GELU (nn.GELU()); callers pass EPI_GELU + glu_act = 2, gemm_launch picks this
                      // instantiation - a runtime switch inside the epilogue cost the tanh flavour 3.4 us per SigLIP fc1 launch
};

struct GemmCall {
    const __nv_bfloat16* W;   // row-major [Nw][K] (row stride ldw), or tile-packed (w_packed)
    int Nw, K, ldw;
    // Tile-packed weights: [Nw/128][ceil(K/64)][128 rows][64 cols] — every 128x64 operand tile is one
    // contiguous 16 KB block and a CTA's whole K-stream is contiguous, so HBM sees long sequential
    // bursts instead of 128-byte reads at a K*2-byte stride (gemm_pack_weight_index()).
    int w_packed;
    const __nv_bfloat16* X;   // [T][K], row stride ldx
    int T, ldx;
    int epi;
    int splitk;               // >= 1 (EPI_PARTIAL only when > 1)
    const __nv_bfloat16* bias;
    __nv_bfloat16* out;
    int ldo;
    float* partial;
    int bn_override;          // 0 = automatic token chunking
    unsigned long long* trace;   // in-graph timeline slot (launch.cuh) or nullptr
    int w_static;             // 1: W was written before any kernel still in flight (engine weights), so the
                              // kernel may fetch it ahead of the programmatic-dependency wait
    int glu_act;              // EPI_GEGLU: gate activation, 0 = tanh GELU (Gemma), 1 = SiLU (Llama SwiGLU)
};

// Device-side parameters of one GEMM (filled from a GemmPlan by gemm_launch / gemm_make_step_op).
struct GemmDev {
    int T;            // valid token rows
    int bn;           // tokens per UMMA chunk (multiple of 16, <= 256)
    int nt;           // chunks per CTA (nt * bn <= 512 TMEM columns)
    int stages;       // smem pipeline depth
    int kb_total;     // ceil(K / 64)
    int kb_per_split; // k-blocks per blockIdx.z
    int tmem_cols;    // power of two >= nt * bn
    int Nw;           // padded weight rows (multiple of 128)
    const __nv_bfloat16* bias; // [Nw] or nullptr
    __nv_bfloat16* out;        // bf16 output
    int ldo;          // output row stride (elements)
    float* partial;   // EPI_PARTIAL: [splitk][T][Nw] fp32
    int w_packed;     // weights are tile-packed (see gemm_tc.h)
    int cluster;      // CTAs (consecutive weight tiles) sharing one multicast activation tile
    int slice_rows;   // activation rows each CTA of the cluster loads and multicasts
    unsigned long long* trace;
    int w_static;     // weights may be fetched before griddepcontrol.wait
    int acc_bufs;     // persistent kernel: accumulator buffers in TMEM (1 or 2)
    int acc_stride;   // TMEM columns between them
    int staging_bytes; // persistent kernel: bf16 output staging tile behind the ring (0 = direct epilogue)
    int l2_policy;    // persistent pairs: 0 = weights evict_first / tokens evict_last, 1 = both evict_normal, 2 = weights evict_last / tokens evict_first
    int glu_act;      // EPI_GEGLU gate activation (GemmCall::glu_act)
    int band;         // persistent pairs: weight tile pairs per raster band (the band sweeps every token tile before the next one starts)
};

struct GemmPlan {
    bool valid;
    int bn, nt, stages, kb_total, kb_per_split, splitk, tmem_cols, smem_bytes, grid_x, grid_y;
    int cluster, slice_rows;
    int two_cta;       // launched as CTA pairs (tcgen05 cta_group::2)
};

GemmPlan gemm_make_plan(int T, int Nw, int K, int splitk, int epi, int bn_override);
// K slices (return value) and tokens per chunk (*bn_override, 0 = one CTA holds all rows) for an EPI_PARTIAL GEMM of
// 33..288 rows, from the measured cost model (tools/sweep_splitk.py)
int gemm_plan_chunked_splitk(int T, int Nw, int K, size_t ws_floats, int* bn_override);

// Returns the number of split-K slices actually used (>= 1), or -1 with *err set.
int gemm_launch(cudaStream_t stream, const GemmCall& call, std::string* err);

// Non-zero if a pipeline wait of a GEMM kernel expired since the last call (synchronises the
// device through cudaMemcpyFromSymbol); clears the flag.
int gemm_take_timeout_flag();

// Element offset of W[row][col] inside the tile-packed layout (K padded to a multiple of 64).
inline size_t gemm_pack_weight_index(int row, int col, int K) {
    const int kb_total = (K + 63) / 64;
    return ((static_cast<size_t>(row / 128) * kb_total + col / 64) * 128 + row % 128) * 64 + col % 64;
}

// CTA pairs (tcgen05 cta_group::2) for GEMMs with an even number of weight tiles and >= 64 tokens.
void gemm_set_use_2cta(int on);
// Persistent one-CTA-per-SM kernel with a direct TMEM -> global epilogue (default) vs one tile per CTA.
void gemm_set_persistent(int on);
void gemm_set_max_stages(int n);
// Cached SWIZZLE_128B tensor map over a row-major bf16 matrix [rows][cols] (row stride ld elements) with boxes of
// box_rows x 64 columns; returns nonzero and sets *err on failure.
int gemm_get_tensor_map(const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out, std::string* err);
void gemm_set_large_t_mode(int mode);
void gemm_set_pair_band(int band);          // 0 = automatic
void gemm_set_pair_small(int mode);         // persistent CTA pairs for 257..288 tokens: 0 never, 1 GeGLU only (default), 2 every epilogue
void gemm_set_pair_policy(int policy);      // -1 = automatic
int gemm_pair_raster(int N, int K, int T, int* band_out, int* n_pairs_out, int32_t* order, int capacity);

// Largest cluster (1, 2, 4, 8) used for activation multicast; 1 disables it.
void gemm_set_cluster_max(int c);
int gemm_set_cta_trace(void* dev_ptr);      // per-CTA timeline buffer [n_cta][8] u64 (0 = off); see gemm_body.cuh

// Tensor maps are cached by (pointer, shape); call when buffers are freed.
void gemm_forget_tensor_maps();

}  // namespace blurr
