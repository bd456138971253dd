/*
 * blurr_vit.h - C ABI of the generic pre-LN ViT encoder and of the MLP projector behind the OpenVLA-7B-shaped path's
 * vision side (BASELINE.json configs[4]: "SigLIP+DINOv2 fused encoder").
 *
 * What it replaces.  Inside `model.predict_action` (`scripts/benchmark_hf_vla.py:152`) OpenVLA's remote code runs two
 * ViTs on the same 224x224 image - DINOv2 ViT-L/14 with 4 register tokens and SigLIP-so400m/14 - takes the patch tokens of
 * each tower's second-to-last block, concatenates them along the feature dim (1024 + 1152) and feeds a 3-layer GELU MLP
 * projector (2176 -> 8704 -> 4096 -> 4096).  None of that code is in the reference tree (parity against it: unpinned);
 * the arithmetic here follows transformers 5.5 `Dinov2WithRegistersModel` / `SiglipVisionModel` with eager attention
 * (a library of this image), against which tests/test_gpu_vit.py pins it.
 *
 * One encoder handle = one tower: patch embedding (14x14 conv as im2col + GEMM), optional cls / register prefix tokens,
 * `num_layers` blocks of LN -> fused QKV -> attention -> out-proj (+ LayerScale) -> residual -> LN -> fc1 + GELU -> fc2
 * (+ LayerScale) -> residual; output = the patch-token rows after the last block that was built.  All on the kernels of
 * the Pi-0 SigLIP tower (tcgen05 GEMMs, LayerNorm consumers, the attention kernels of csrc/attention.cu).
 * Device pointers unless stated; errors as in blurr_pi0.h.  bf16 weights.
 */
#ifndef BLURR_VIT_H_
#define BLURR_VIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blurr_vit blurr_vit_t;
typedef struct blurr_mlp blurr_mlp_t;

typedef struct blurr_vit_config {
    int32_t abi_version;          /* BLURR_VIT_ABI_VERSION */
    int32_t num_layers;           /* blocks to run (OpenVLA: depth - 1 of each tower) */
    int32_t hidden;               /* 1024 (DINOv2-L) / 1152 (SigLIP-so400m); multiple of 128 */
    int32_t num_heads;            /* 16; head_dim = hidden / num_heads must be a multiple of 8 and <= 80 */
    int32_t mlp_dim;              /* 4096 / 4304 */
    int32_t image_size;           /* 224 */
    int32_t patch_size;           /* 14 */
    int32_t num_prefix_tokens;    /* 5 = cls + 4 registers (DINOv2), 0 (SigLIP) */
    int32_t use_layerscale;       /* 1 (DINOv2) / 0 */
    int32_t gelu_erf;             /* 1: exact GELU (DINOv2), 0: tanh approximation (SigLIP) */
    float ln_eps;                 /* 1e-6 */
} blurr_vit_config;
#define BLURR_VIT_ABI_VERSION 1

int blurr_vit_create(const blurr_vit_config* cfg, int device, int max_batch, blurr_vit_t** out);
void blurr_vit_destroy(blurr_vit_t* h);
/* Keys: "patch.weight" [hidden][3*p*p] (conv weight flattened c,kh,kw), "patch.bias" [hidden], "pos" [n_patches][hidden]
 * (position embeddings of the patch tokens), "prefix" [num_prefix_tokens][hidden] (the prefix rows as they enter block 0:
 * cls + its position embedding, then the register tokens), and per block N: "layers.N.ln1.weight|bias",
 * "layers.N.q|k|v|o.weight|bias", "layers.N.ls1" [hidden], "layers.N.ln2.weight|bias", "layers.N.fc1|fc2.weight|bias",
 * "layers.N.ls2" [hidden].  Tensors are copied and repacked. */
int blurr_vit_set_weight(blurr_vit_t* h, const char* key, const void* dev_ptr, const int64_t* shape, int ndim);
int blurr_vit_finalize(blurr_vit_t* h);
/* pixel_values: bf16 [batch][3][image][image] with element strides (b, c, h, w).  Writes the patch-token features,
 * bf16 [batch * n_patches][out_ld], columns 0..hidden-1 of the rows starting at `out` (so two towers can fill the two
 * halves of one concatenated feature matrix). */
int blurr_vit_forward(blurr_vit_t* h, void* cuda_stream, int batch, const void* pixel_values, const int64_t strides[4],
                      void* out, int out_ld);
/* Options: "use_cuda_graph" (default 1: the blocks replay as one CUDA graph per batch size). */
int blurr_vit_set_option(blurr_vit_t* h, const char* name, int64_t value);
int64_t blurr_vit_last_launch_count(const blurr_vit_t* h);

/* Projector: y = W_n(... GELU(W_1 x + b_1) ...) + b_n, exact GELU between the layers (nn.GELU()). dims = n_layers + 1
 * widths, e.g. {2176, 8704, 4096, 4096}. */
int blurr_mlp_create(const int32_t* dims, int n_layers, int device, int max_rows, blurr_mlp_t** out);
void blurr_mlp_destroy(blurr_mlp_t* h);
int blurr_mlp_set_weight(blurr_mlp_t* h, int layer, const void* weight_dev, const void* bias_dev);   /* [out][in], [out] */
int blurr_mlp_forward(blurr_mlp_t* h, void* cuda_stream, int rows, const void* x, int ldx, void* y, int ldy);

#ifdef __cplusplus
}
#endif
#endif /* BLURR_VIT_H_ */
