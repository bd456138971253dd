/*
 * blurr_pi0.h — C ABI of the B200-native Pi-0 control-step engine (libblurr_pi0.so).
 *
 * The reference (JijiKing-Sam/BLURR, vendored open-pi-zero) has no plugin / FFI interface:
 * its seam for this path is the Python class
 *     src.model.vla.pizero.PiZeroInference      third_party/open_pi_zero/src/model/vla/pizero.py:721-742
 * whose `forward` is `PiZero.infer_action` (pizero.py:473-547).  This header is the boundary a
 * replacement binds underneath that class: plain pointers and sizes, no torch types.  The
 * Python mirror of the class (blurr_b200/pizero.py) calls it through ctypes; INTEGRATION.md
 * shows the reference-side binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative blurr_status on error; the message is
 *     available from blurr_last_error() (thread-local);
 *   - all `dev` pointers are device pointers on the handle's device, owned by the caller;
 *     the handle owns repacked weights, workspace, KV cache and CUDA graphs;
 *   - arithmetic dtype is bf16 storage / fp32 accumulate (the `--preset blurr` path);
 *   - calls are asynchronous on the given stream unless stated; a handle is not thread-safe;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef BLURR_PI0_H_
#define BLURR_PI0_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLURR_ABI_VERSION 1

typedef enum blurr_status {
    BLURR_OK = 0,
    BLURR_ERR_INVALID = -1,   /* bad argument / unsupported shape                         */
    BLURR_ERR_CUDA = -2,      /* CUDA runtime / driver error                              */
    BLURR_ERR_STATE = -3,     /* call order (e.g. infer before finalize, missing weight)  */
    BLURR_ERR_INPUT = -4      /* device-side input validation failed (token id range ...) */
} blurr_status;

typedef enum blurr_dtype { BLURR_BF16 = 1, BLURR_F32 = 0, BLURR_I64 = 2 } blurr_dtype;

/* Model hyper-parameters: the values of third_party/open_pi_zero/config/eval/bridge.yaml:35-136
 * that shape the computation (cfg.* names of pizero.py:35-120 in comments). */
typedef struct blurr_pi0_config {
    int32_t abi_version;            /* BLURR_ABI_VERSION                                  */
    /* SigLIP (cfg.vision.config) */
    int32_t vision_layers;          /* num_hidden_layers      27                          */
    int32_t vision_hidden;          /* hidden_size            1152                        */
    int32_t vision_intermediate;    /* intermediate_size      4304                        */
    int32_t vision_heads;           /* num_attention_heads    16                          */
    int32_t image_size;             /* 224                                                */
    int32_t patch_size;             /* 14                                                 */
    int32_t num_image_tokens;       /* 256                                                */
    float   layer_norm_eps;         /* 1e-6                                               */
    /* joint model (cfg.joint.config, cfg.mixture.*) */
    int32_t joint_layers;           /* num_hidden_layers      18                          */
    int32_t num_heads;              /* num_attention_heads    8                           */
    int32_t num_kv_heads;           /* num_key_value_heads    1                           */
    int32_t head_dim;               /* 256                                                */
    int32_t vlm_hidden;             /* mixture.vlm.hidden_size        2048                */
    int32_t vlm_intermediate;       /* mixture.vlm.intermediate_size  16384               */
    int32_t expert_hidden;          /* mixture.{proprio,action}.hidden_size        1024   */
    int32_t expert_intermediate;    /* mixture.{proprio,action}.intermediate_size  4096   */
    float   rms_norm_eps;           /* 1e-6                                               */
    /* sequence layout (pizero.py:44-51) */
    int32_t max_image_text_tokens;  /* 276                                                */
    int32_t num_proprio_tokens;     /* cond_steps      1                                  */
    int32_t num_action_tokens;      /* horizon_steps   4                                  */
    int32_t action_dim;             /* 7                                                  */
    int32_t proprio_dim;            /* 7 (Bridge) / 8 (Fractal)                           */
    /* tokens (bridge.yaml:94-96) */
    int64_t vocab_size;             /* 257216                                             */
    int64_t image_token_index;      /* 257152                                             */
    int64_t pad_token_id;           /* 0                                                  */
    /* flow matching (pizero.py:58,62) */
    int32_t num_inference_steps;    /* 1 for --preset blurr, 10 baseline                  */
    int32_t has_clip;               /* final_action_clip_value is not None                */
    float   final_action_clip_value;
} blurr_pi0_config;

typedef struct blurr_pi0 blurr_pi0_t;   /* opaque: one per (device, max_batch) */

/* Lifetime.  Replaces `PiZeroInference(cfg)` + `.to(device)` (benchmark_pi0.py:127-142). */
int blurr_pi0_create(const blurr_pi0_config* cfg, int device, int max_batch, blurr_pi0_t** out);
void blurr_pi0_destroy(blurr_pi0_t* h);

/* Weights.  Replaces `load_state_dict` (benchmark_pi0.py:139): one call per state_dict entry,
 * key spelled exactly as in the reference state_dict.  The tensor is
 * copied and repacked (fused QKV, interleaved gate/up, K padding); dtype must be BLURR_BF16. */
int blurr_pi0_set_weight(blurr_pi0_t* h, const char* state_dict_key, const void* dev_ptr,
                         const int64_t* shape, int ndim, int dtype);
/* RoPE inverse frequencies of one mixture ("vlm" | "proprio" | "action"), 128 fp32 values on the
 * HOST, already rounded the way `model.to(dtype)` rounds the `inv_freq` buffer
 * (paligemma/modules.py:41-45). */
int blurr_pi0_set_rope_inv_freq(blurr_pi0_t* h, const char* mixture, const float* host_inv_freq, int n);
/* time_cond rows for the flow steps, bf16 [num_steps][expert_hidden] on the device, computed by
 * the caller with the reference's own ops (vla/modules.py:15-22; `t` accumulates in bf16). */
int blurr_pi0_set_time_table(blurr_pi0_t* h, const void* dev_table, int num_steps);
/* Checks every weight is present, builds RoPE tables.  Synchronous. */
int blurr_pi0_finalize_weights(blurr_pi0_t* h);

/* The control step.  Replaces `PiZeroInference.forward` = `PiZero.infer_action`
 * (pizero.py:473-547, 721-742).  Shapes for batch B (Bridge sizes in brackets):
 *   input_ids                int64 [B][max_image_text_tokens]            [B][276]
 *   pixel_values             bf16  [B][3][H][W], element strides given    [B][3][224][224]
 *   image_text_proprio_mask  bf16  additive, element (b,r,c) at mask[b*bstride + r*rstride + c],
 *                                  r,c < max_image_text_tokens + num_proprio_tokens   [277x277]
 *   action_mask              bf16  (b,r,c) likewise, r < num_action_tokens, c < total [4x281]
 *   *_position_ids           int64 [B][276], [B][1], [B][4]
 *   proprios                 bf16  [B][num_proprio_tokens][proprio_dim]
 *   noise                    bf16  [B][num_action_tokens][action_dim] — the `torch.randn` draw of
 *                                  pizero.py:511-513, made by the caller
 *   actions_out              bf16  [B][num_action_tokens][action_dim]
 * Fresh KV state per call, like the reference (pizero.py:487). */
typedef struct blurr_pi0_inputs {
    const int64_t* input_ids;
    const void* pixel_values;
    int64_t pixel_strides[4];            /* elements: batch, channel, row, column */
    const void* image_text_proprio_mask;
    int64_t itp_mask_bstride, itp_mask_rstride;
    const void* action_mask;
    int64_t action_mask_bstride, action_mask_rstride;
    const int64_t* vlm_position_ids;
    const int64_t* proprio_position_ids;
    const int64_t* action_position_ids;
    const void* proprios;
    const void* noise;
} blurr_pi0_inputs;

int blurr_pi0_infer_action(blurr_pi0_t* h, void* cuda_stream, int batch, const blurr_pi0_inputs* in,
                           void* actions_out);

/* Options: "use_cuda_graph" (default 1; replay of the step's kernels as a CUDA graph), "debug_taps" (default 0; implies eager launches),
 * "chunked_splitk" (default 1: split-K GEMMs of 128..288 tokens also split the tokens across CTAs, chosen by a measured cost model),
 * "activation_clip_bits" / "activation_clip_mask" (default 0 / 0: the reference's int8 fake-quant mode, int8_linear.py:72-83 - the float32
 * bit pattern of the clamp applied to the inputs of the swapped Linears, and which modules: bit 1 proprio mixture (tied weights), bit 2
 * action mixture, bit 3 action encoder; the de-quantised weights arrive through blurr_pi0_set_weight like any others),
 * "use_pdl" (default 1: programmatic dependent launch between the step's kernels; process-wide),
 * "num_inference_steps". */
int blurr_pi0_set_option(blurr_pi0_t* h, const char* name, int64_t value);
/* Synchronises the stream and reports (and clears) the device-side sticky error words: input validation (token id outside the
 * embedding table, too many image tokens) and bounded pipeline waits that expired inside a GEMM / attention kernel.  A
 * step that tripped one of them has already written NaN actions (the step's last kernel poisons them). */
int blurr_pi0_check(blurr_pi0_t* h, void* cuda_stream);

/* Debug / parity taps: copy a named internal buffer (bf16 unless noted) to `dst_dev`.
 * Names: "k_cache" / "v_cache" [layers][B][slots][head_dim]; with option debug_taps=1 also
 * "siglip.embeddings", "siglip.layer<l>", "siglip.post_layernorm", "projector",
 * "merged_embeds", "prefill.L<l>.vlm", "prefill.L<l>.proprio", "flow<s>.L<l>.action",
 * "flow<s>.velocity".  Returns the byte size of the buffer through *bytes_out when dst_dev is NULL. */
int blurr_pi0_debug_tap(blurr_pi0_t* h, const char* name, void* dst_dev, size_t dst_bytes, size_t* bytes_out);

/* Kernel launches issued (or replayed) by the last blurr_pi0_infer_action call. */
int64_t blurr_pi0_last_launch_count(const blurr_pi0_t* h);
/* Option "profile" = 1: every kernel is launched eagerly and bracketed by CUDA events; the per-label
 * totals accumulated since "profile" = 2 (reset) are written as text into `buf`. */
int blurr_pi0_profile_report(blurr_pi0_t* h, char* buf, size_t buf_bytes);
/* Option "trace" = 1: the GEMM, consumer, RoPE and attention kernels stamp %globaltimer (first CTA
 * start, first CTA past the programmatic-dependency wait, last CTA end) while the step runs in its
 * normal regime (CUDA graph, PDL, three streams).  Writes one text line per kernel of the last call:
 * "idx stream start_us waited_us end_us label". */
int blurr_pi0_trace_report(blurr_pi0_t* h, char* buf, size_t buf_bytes);
/* Bytes of repacked weights the step reads (the algorithmic-bytes numerator of the roofline). */
int64_t blurr_pi0_weight_bytes(const blurr_pi0_t* h);

/* ---------------------------------------------------------------------------------------------
 * Observation preprocessing on the device.  Replaces, bit for bit, the host
 * work the reference does before every model call:
 *   cv2.resize(image, image_size, interpolation=cv2.INTER_LANCZOS4)     agent/env_adapter/simpler.py:59-64
 *   VLAProcessor: uint8 * (1/255.0), (x - 0.5) / 0.5, fp32                model/vla/processing.py:27-58,112-117
 *   pixel_values.to(bfloat16)                                            agent/eval.py:187
 *   normalize_bound / normalize_gaussian (float64) -> float32 -> bf16    agent/env_adapter/base.py:8-18,33-40;
 *                                                                        simpler.py:73-95; eval.py:194
 * --------------------------------------------------------------------------------------------- */
typedef struct blurr_preproc blurr_preproc_t;   /* opaque: one per (device, frame geometry) */
/* Builds OpenCV's 8-tap fixed-point Lanczos tables for src -> dst and uploads them. */
int blurr_preproc_create(int device, int src_h, int src_w, int dst_h, int dst_w, blurr_preproc_t** out);
void blurr_preproc_destroy(blurr_preproc_t* p);
/* Host only (no device needed): the table of one axis, ofs[dst] (first tap = ofs - 3) and alpha[dst*8]. */
int blurr_preproc_build_tables(int src, int dst, int32_t* ofs, int16_t* alpha);
/* Host copies of the tables: x_ofs[dst_w], x_alpha[dst_w*8], y_ofs[dst_h], y_alpha[dst_h*8]. */
int blurr_preproc_tables(const blurr_preproc_t* p, int32_t* x_ofs, int16_t* x_alpha, int32_t* y_ofs, int16_t* y_alpha);
/* frame_u8: device uint8 [batch][src_h][src_w][3] (HWC; strides in bytes).  pixel_values_bf16: device
 * bf16 [batch][3][dst_h][dst_w].  resized_u8 (nullable): device uint8 [batch][dst_h][dst_w][3], the
 * cv2.resize result itself.  Asynchronous on `cuda_stream`. */
int blurr_preproc_frame(blurr_preproc_t* p, void* cuda_stream, const void* frame_u8, int64_t row_stride_bytes,
                        int batch, int64_t frame_stride_bytes, void* pixel_values_bf16, void* resized_u8);
/* raw: device float64 [n][dim]; lo / hi: device float64 [dim] = (p01, p99) for kind 0 "bound", (mean, std)
 * for kind 1 "gaussian"; out_bf16: device bf16 [n][dim]. */
int blurr_op_normalize_proprio(void* cuda_stream, const double* raw, const double* lo, const double* hi, int kind,
                               int n, int dim, void* out_bf16);

/* Process-wide tuning knobs (no handle): "gemm_cluster_max" (1/2/4/8, activation-multicast cluster
 * size cap of the GEMM kernel), "gemm_use_2cta" (-1 automatic = CTA pairs for the GeGLU GEMM above 1024 tokens, 0, 1),
 * "gemm_large_t_mode" (above 1024 tokens: -1 automatic = persistent CTA pairs for every bf16 epilogue (GeGLU, GELU,
 * plain store), persistent single CTAs for fp32 partial / residual epilogues; 0 never persistent; 1 single-CTA
 * persistent for every epilogue; 2 and 3 = same as automatic), "attn_tc" (-1 automatic = tcgen05 prefill / SigLIP attention whenever the shape allows, 0 never (mma.sync tile kernel), 1 = -1),
 * "attn_tc_fewq" (1: the experts' few-query attention on the tcgen05 kernel too; default 0, measured slower at batch 1),
 * "attn_prefill_stream", "attn_fewq_stream", "attn_siglip_stream" (default 1: at one or two episodes the prefill, the experts'
 * few-query and the SigLIP attention run as the streaming mma.sync kernels of csrc/attention.cu - K straight into permuted
 * fragments, V in shared memory, every load requested up front; 0: the tile / tcgen05 kernels),
 * "gemm_pair_small" (257..288 tokens: 0 one CTA per weight tile, 1 persistent CTA pairs for GeGLU only, 2 (default) for every epilogue),
 * "gemm_cta_trace" / "attn_cta_trace" (device pointer to [n_cta][8] u64, 0 = off: per-CTA %globaltimer timeline), "op_gemm_bn" (token chunk
 * width used by blurr_op_gemm*, 0 = automatic),
 * "gemm_pair_band" (persistent pairs: weight tile pairs per raster band, 0 = automatic), "gemm_pair_policy" (L2
 * hints above 1024 tokens: -1 automatic = 1; 0 weights evict_first / tokens evict_last; 1 both evict_normal; 2 weights
 * evict_last / tokens evict_first),
 * "gemm_persistent" (0/1: persistent kernel for GEMMs of <= 32 tokens), "gemm_max_stages" (TMA ring depth cap),
 * "use_pdl" (0/1). */
int blurr_set_global_option(const char* name, int64_t value);

const char* blurr_last_error(void);
int blurr_abi_version(void);

/* ---- single-operator entry points (same kernels, used by the per-kernel parity tests and
 *      micro-benchmarks; synchronous on `cuda_stream` only with respect to errors) ---- */
/* Pack a row-major bf16 weight W[N][K] (N % 128 == 0, K % 64 == 0) into the tile-packed layout the
 * engine streams from HBM: [N/128][K/64][128][64] (every 128x64 operand tile contiguous). */
int blurr_op_pack_weight(void* cuda_stream, const void* W, int N, int K, int ldw, void* packed);
/* W is row-major with row stride ldw > 0, or tile-packed when ldw == 0.
 * Y[T][N] = epilogue(X[T][K] @ W[N][K]^T); epi: 0 store(+bias) 1 gelu 2 geglu 3 partial(fp32 out,
 * [splitk_used][T][N]); returns the number of split-K slices used (>=1) or a negative status. */
int blurr_op_gemm(void* cuda_stream, const void* W, int N, int K, int ldw, const void* X, int T, int ldx,
                  int epi, int splitk, const void* bias, void* out, int ldo, float* partial);
/* Host-only (no GPU needed): the order in which the persistent CTA pairs of the batched GEMM (> 1024 tokens, bf16
 * epilogue) visit the tiles of a [T][N] output under the current global options.  Writes the band width (weight
 * tile pairs per raster band), the number of CTA pairs, and up to `capacity` (weight_tile_pair, token_tile)
 * int32 pairs in visiting order; returns the number of tiles or a negative status. */
int blurr_op_pair_raster(int N, int K, int T, int* band, int* n_pairs, int32_t* order, int capacity);
/* Same launch without the trailing stream synchronisation / pipeline-timeout check (for timing
 * loops); returns the split-K slice count or a negative status. */
int blurr_op_gemm_async(void* cuda_stream, const void* W, int N, int K, int ldw, const void* X, int T, int ldx,
                        int epi, int splitk, const void* bias, void* out, int ldo, float* partial);
int blurr_op_siglip_attention(void* cuda_stream, const void* qkv, int ld_qkv, int batch, int seq, int heads,
                              int hidden, void* out, int ld_out);
int blurr_op_joint_attention(void* cuda_stream, int few_query, const void* q, int q_per_sample,
                             int q_row_offset, const void* k_cache, const void* v_cache, int n_slots,
                             int n_keys, const void* mask, int64_t mask_bstride, int64_t mask_rstride,
                             int batch, int n_heads, void* out);

#ifdef __cplusplus
}
#endif
#endif /* BLURR_PI0_H_ */
