// Pi-0 control-step engine: weight repacking, workspace, the kernel schedule of one
// `infer_action` call (reference: third_party/open_pi_zero/src/model/vla/pizero.py:473-547)
// and its CUDA-graph replay, behind the C ABI of include/blurr_pi0.h.
#include "../../include/blurr_pi0.h"

#include "common.cuh"
#include "gemm_tc.h"
#include "kernels.h"
#include "launch.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <set>
#include <string>
#include <vector>

using namespace blurr;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
namespace blurr {
int record_error(int code, const std::string& msg) { return ::fail(code, msg); }      // for the other translation units
}
#define CUDA_TRY(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fail(BLURR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    } while (0)

// ---------------------------------------------------------------------------
// repack kernels
// ---------------------------------------------------------------------------
enum RowMap { MAP_OFFSET = 0, MAP_GATE = 1, MAP_UP = 2 };
// dst_ld > 0: row-major destination; dst_ld == 0: tile-packed GEMM weight layout with `kb_total`
// 64-column blocks per row ([row/128][col/64][row%128][col%64], see gemm_tc.h).
__global__ void repack_rows_kernel(const bf16* __restrict__ src, int rows, int cols, int src_ld,
                                   bf16* __restrict__ dst, int dst_ld, int mode, int row_off, int kb_total) {
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<size_t>(rows) * cols) return;
    const int r = static_cast<int>(idx / cols), c = static_cast<int>(idx - static_cast<size_t>(r) * cols);
    int dr;
    if (mode == MAP_OFFSET) dr = r + row_off;
    else dr = 2 * r + (mode == MAP_UP ? 1 : 0);       // gate_j, up_j alternate: the GeGLU epilogue pairs adjacent rows
    size_t di;
    if (dst_ld > 0) di = static_cast<size_t>(dr) * dst_ld + c;
    else di = ((static_cast<size_t>(dr / 128) * kb_total + c / 64) * 128 + dr % 128) * 64 + c % 64;
    dst[di] = src[static_cast<size_t>(r) * src_ld + c];
}

struct StageArgs {
    const int64_t* ids; const int64_t* vpos; const int64_t* ppos; const int64_t* apos;
    const bf16* proprios; const bf16* noise;
    const bf16* mask_itp; long long itp_bs, itp_rs;
    const bf16* mask_act; long long act_bs, act_rs;
    int64_t* d_ids; int64_t* d_vpos; int64_t* d_ppos; int64_t* d_apos;
    bf16* d_proprios; bf16* d_action; bf16* d_mask_itp; bf16* d_mask_act;
    int itp_ld;      // row stride of the staged image/text/proprio mask (n_itp rounded up to 8)
    int n_ids, n_ppos, n_apos, n_prop, n_noise, itp_dim, act_rows, act_cols, batch;
};
// Copies the per-call inputs into the engine's static buffers (the CUDA graph reads those).
__global__ void stage_inputs_kernel(const StageArgs a) {
    long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < a.n_ids) { a.d_ids[i] = a.ids[i]; a.d_vpos[i] = a.vpos[i]; return; }
    i -= a.n_ids;
    if (i < a.n_ppos) { a.d_ppos[i] = a.ppos[i]; return; }
    i -= a.n_ppos;
    if (i < a.n_apos) { a.d_apos[i] = a.apos[i]; return; }
    i -= a.n_apos;
    if (i < a.n_prop) { a.d_proprios[i] = a.proprios[i]; return; }
    i -= a.n_prop;
    if (i < a.n_noise) { a.d_action[i] = a.noise[i]; return; }
    i -= a.n_noise;
    const long long itp_per = static_cast<long long>(a.itp_dim) * a.itp_dim;
    if (i < itp_per * a.batch) {
        const long long b = i / itp_per, rem = i - b * itp_per;
        const long long r = rem / a.itp_dim, c = rem - r * a.itp_dim;
        // staged with rows padded to a multiple of 8 elements: 16-byte mask loads in the batched attention
        a.d_mask_itp[(b * a.itp_dim + r) * a.itp_ld + c] = a.mask_itp[b * a.itp_bs + r * a.itp_rs + c];
        return;
    }
    i -= itp_per * a.batch;
    const long long act_per = static_cast<long long>(a.act_rows) * a.act_cols;
    if (i < act_per * a.batch) {
        const long long b = i / act_per, rem = i - b * act_per;
        const long long r = rem / a.act_cols, c = rem - r * a.act_cols;
        a.d_mask_act[i] = a.mask_act[b * a.act_bs + r * a.act_rs + c];
    }
}

// ---------------------------------------------------------------------------
// engine state
// ---------------------------------------------------------------------------
struct Lin {
    bf16* w = nullptr;             // tile-packed [Nw/128][K/64][128][64] (gemm_tc.h)
    int Nw = 0, K = 0, ld = 0;     // padded rows, padded cols
    bf16* bias = nullptr;          // [Nw] zero padded, or nullptr
};
struct VisionLayer {
    Lin qkv, out, fc1, fc2;
    bf16 *ln1w, *ln1b, *ln2w, *ln2b;
};
struct MixLayer {
    Lin qkv, o, gu, down;
    bf16 *in_ln, *post_ln;
};
struct MixtureW {
    std::string name;
    int hidden = 0, inter = 0;
    std::vector<MixLayer> layers;
    bf16* final_norm = nullptr;
    float inv_freq[128];
    bool have_inv_freq = false;
    float *cos_t = nullptr, *sin_t = nullptr;
};
struct TapBuf { void* ptr; size_t bytes; };

static constexpr int kNumPos = 1024;     // RoPE table rows (position ids 0..1023)
static constexpr int kNumSMs = 148;

struct blurr_pi0 {
    blurr_pi0_config cfg;
    int device = 0, max_batch = 0;
    std::vector<void*> allocs;
    size_t weight_bytes = 0;
    // weights
    bf16* embed = nullptr;
    Lin patch; bf16* pos_emb = nullptr;
    std::vector<VisionLayer> vlayers;
    bf16 *post_ln_w = nullptr, *post_ln_b = nullptr;
    Lin proj;
    MixtureW mix[3];               // 0 vlm, 1 proprio, 2 action
    bf16 *ae1_w = nullptr, *ae1_b = nullptr; Lin ae2, ae3;
    bf16 *pe_w = nullptr, *pe_b = nullptr;
    bf16 *dec_w = nullptr, *dec_b = nullptr;
    std::set<std::string> expected, seen;
    bool finalized = false;
    // time table
    bf16* time_table = nullptr; int time_steps = 0, time_capacity = 0;
    // static inputs
    int64_t *d_ids, *d_vpos, *d_ppos, *d_apos;
    bf16 *d_proprios, *d_action, *d_mask_itp, *d_mask_act, *d_out;
    // activations
    bf16 *patches, *xs, *xn, *sqkv, *sattn, *shmid, *imgfeat;
    bf16 *E, *En, *Qv, *AOv, *H;
    bf16 *Ep, *Epn, *Qp, *AOp, *Hp;
    bf16 *Ea, *Ean, *Qa, *AOa, *Ha, *X2, *A1;
    bf16 *kcache, *vcache;
    float* ws = nullptr; size_t ws_floats = 0;
    float* ws2 = nullptr; size_t ws2_floats = 0;     // second split-K workspace: the proprio stream runs beside the VLM
    float* ws3 = nullptr;                            // third: the action stream (same size as ws2)
    // The proprio expert and the first flow step of the action expert run on their own streams,
    // concurrently with the VLM prefill: layer l of either only needs layer l's K/V of the streams
    // before it, so they pipeline one layer behind instead of adding ~290 kernels to the critical path.
    bool use_streams = true;
    bool chunked_splitk = true;    // split-K GEMMs of 128..288 tokens may also split the tokens across CTAs (Run::plan_partial)
    cudaStream_t s_prop = nullptr, s_act = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_done_p = nullptr, ev_done_a = nullptr;
    std::vector<cudaEvent_t> ev_v, ev_p;
    int* d_err = nullptr;
    // int8 fake-quant mode of the reference (QuantizedLinear.forward, int8_linear.py:72-83): the input of every quantised
    // Linear is clamped to +-act_clip first.  Bits of act_clip_mask: 1 = proprio mixture (tied weights), 2 = action mixture,
    // 3 = action encoder (the reference's swap skips the bare-Linear decoder and proprio encoder).  0 = off (every shipped config).
    float act_clip = 0.f; int act_clip_mask = 0;
    bf16* clip_action = nullptr;             // clamped copy of the action encoder's input (the flow state itself stays unclamped)
    const int *flag_gemm = nullptr, *flag_attn = nullptr;    // sticky pipeline time-out words of the kernels (NaN-poison the actions)
    // options / bookkeeping
    bool use_graph = true, debug = false;
    int stage_mask = 7;            // bit 0 vision, bit 1 prefill, bit 2 action flow (timing experiments)
    bool profile = false;          // eager launches bracketed by CUDA events, per-kernel-label totals
    // in-graph timeline: kernels stamp %globaltimer into trace_buf (launch.cuh); slots are handed out
    // while the step is issued / captured, so every graph replay rewrites the same slots
    bool lin_mode = true;          // single-slice GEMMs above 1024 tokens hand bf16 (not fp32 partials) to their consumer
    bool trace = false;
    unsigned long long* trace_buf = nullptr;       // [kTraceMax][4]
    std::vector<std::string> trace_labels;
    std::vector<int> trace_streams;
    struct ProfEntry { int count = 0; double ms = 0.0; };
    std::map<std::string, ProfEntry> prof;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> prof_pending;
    int64_t launches = 0;
    std::map<std::string, TapBuf> taps;
    struct GraphEntry { cudaGraph_t graph; cudaGraphExec_t exec; int64_t launches; };
    std::map<long long, GraphEntry> graphs;
    // derived sizes
    int T_img, n_itp, n_total;    // 256, 277, 281
};

static void* dalloc(blurr_pi0* h, size_t bytes, bool is_weight = false) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, bytes);
    h->allocs.push_back(p);
    if (is_weight) h->weight_bytes += bytes;
    return p;
}
static int pad_to(int v, int m) { return (v + m - 1) / m * m; }

static bool alloc_lin(blurr_pi0* h, Lin& L, int N, int K, bool bias) {
    L.Nw = pad_to(N, 128);
    L.K = pad_to(K, 64);
    L.ld = L.K;
    L.w = static_cast<bf16*>(dalloc(h, static_cast<size_t>(L.Nw) * L.ld * 2, true));
    if (!L.w) return false;
    if (bias) {
        L.bias = static_cast<bf16*>(dalloc(h, static_cast<size_t>(L.Nw) * 2, true));
        if (!L.bias) return false;
    }
    return true;
}
static bf16* alloc_vec(blurr_pi0* h, int n, bool weight = true) {
    return static_cast<bf16*>(dalloc(h, static_cast<size_t>(n) * 2, weight));
}

static const char* kMixNames[3] = {"vlm", "proprio", "action"};
static const char* kVT = "vision_tower.vision_model.";

static void build_expected(blurr_pi0* h) {
    auto& e = h->expected;
    e.insert("embed_tokens.weight");
    e.insert(std::string(kVT) + "embeddings.patch_embedding.weight");
    e.insert(std::string(kVT) + "embeddings.patch_embedding.bias");
    e.insert(std::string(kVT) + "embeddings.position_embedding.weight");
    for (int l = 0; l < h->cfg.vision_layers; ++l) {
        const std::string p = std::string(kVT) + "encoder.layers." + std::to_string(l) + ".";
        for (const char* n : {"q_proj", "k_proj", "v_proj", "out_proj"})
            for (const char* s : {"weight", "bias"}) e.insert(p + "self_attn." + n + "." + s);
        for (const char* n : {"layer_norm1", "layer_norm2", "mlp.fc1", "mlp.fc2"})
            for (const char* s : {"weight", "bias"}) e.insert(p + n + "." + s);
    }
    e.insert(std::string(kVT) + "post_layernorm.weight");
    e.insert(std::string(kVT) + "post_layernorm.bias");
    e.insert("multi_modal_projector.linear.weight");
    e.insert("multi_modal_projector.linear.bias");
    for (int m = 0; m < 3; ++m) {
        for (int l = 0; l < h->cfg.joint_layers; ++l) {
            const std::string p = std::string("joint_model.mixtures.") + kMixNames[m] + ".layers." +
                                  std::to_string(l) + ".";
            for (const char* n : {"self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
                                  "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj", "input_layernorm",
                                  "post_attention_layernorm"})
                e.insert(p + n + ".weight");
        }
        if (m > 0) e.insert(std::string("joint_model.mixtures.") + kMixNames[m] + ".norm.weight");
    }
    for (const char* n : {"action_encoder.linear_1", "action_encoder.linear_2", "action_encoder.linear_3",
                          "proprio_encoder", "action_decoder"})
        for (const char* s : {"weight", "bias"}) e.insert(std::string(n) + "." + s);
}

static int validate_cfg(const blurr_pi0_config& c) {
    if (c.abi_version != BLURR_ABI_VERSION) return fail(BLURR_ERR_INVALID, "config abi_version mismatch");
    if (c.head_dim != 256 || c.num_kv_heads != 1 || c.num_heads != 8)
        return fail(BLURR_ERR_INVALID, "joint attention kernels are built for head_dim 256, 8 query heads, 1 KV head (MQA)");
    if (c.vision_hidden % c.vision_heads != 0 || c.vision_hidden / c.vision_heads > 80 ||
        (c.vision_hidden / c.vision_heads) % 8 != 0)
        return fail(BLURR_ERR_INVALID, "SigLIP head_dim must be a multiple of 8 and <= 80");
    if (c.patch_size != 14 || c.image_size != 224 || c.num_image_tokens != 256)
        return fail(BLURR_ERR_INVALID, "patch-embed kernel is built for 224x224 images, 14x14 patches");
    if (c.vision_hidden % 128 || c.vlm_hidden % 128 || c.expert_hidden % 128 || c.vlm_intermediate % 64 ||
        c.expert_intermediate % 64 || (c.num_heads * c.head_dim) % 128)
        return fail(BLURR_ERR_INVALID, "hidden sizes must be multiples of 128 (intermediate: 64)");
    if (c.proprio_dim > 8 || c.action_dim > 8 || c.proprio_dim < 1 || c.action_dim < 1)
        return fail(BLURR_ERR_INVALID, "proprio_dim/action_dim must be in 1..8");
    if (c.vision_layers < 1 || c.joint_layers < 1 || c.num_inference_steps < 1)
        return fail(BLURR_ERR_INVALID, "layer counts / num_inference_steps must be >= 1");
    if (c.max_image_text_tokens < c.num_image_tokens || c.num_action_tokens < 1 || c.num_proprio_tokens < 1)
        return fail(BLURR_ERR_INVALID, "bad sequence layout");
    if (c.max_image_text_tokens + c.num_proprio_tokens + c.num_action_tokens > 512)
        return fail(BLURR_ERR_INVALID, "sequence longer than 512 tokens is not supported");
    return 0;
}

extern "C" int blurr_abi_version(void) { return BLURR_ABI_VERSION; }
extern "C" const char* blurr_last_error(void) { return g_err.c_str(); }

extern "C" void blurr_pi0_destroy(blurr_pi0_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->graphs) {
        cudaGraphExecDestroy(kv.second.exec);
        cudaGraphDestroy(kv.second.graph);
    }
    for (cudaEvent_t e : h->ev_v) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_p) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_fork, h->ev_done_p, h->ev_done_a}) if (e) cudaEventDestroy(e);
    if (h->s_prop) cudaStreamDestroy(h->s_prop);
    if (h->s_act) cudaStreamDestroy(h->s_act);
    for (auto& kv : h->taps) cudaFree(kv.second.ptr);
    for (void* p : h->allocs) cudaFree(p);
    if (h->trace_buf) cudaFree(h->trace_buf);
    gemm_forget_tensor_maps();
    delete h;
}

extern "C" int blurr_pi0_create(const blurr_pi0_config* cfg, int device, int max_batch, blurr_pi0_t** out) {
    if (!cfg || !out || max_batch < 1) return fail(BLURR_ERR_INVALID, "blurr_pi0_create: bad arguments");
    if (int r = validate_cfg(*cfg)) return r;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(BLURR_ERR_CUDA, "no CUDA device: the B200 path has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(BLURR_ERR_INVALID, "bad device index");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BLURR_ERR_CUDA, std::string("device is sm_") + std::to_string(prop.major) +
                                        std::to_string(prop.minor) + ", kernels are built for sm_100a only");
    blurr_pi0* h = new blurr_pi0();
    h->cfg = *cfg;
    h->device = device;
    h->max_batch = max_batch;
    const auto& c = h->cfg;
    h->T_img = c.num_image_tokens;
    h->n_itp = c.max_image_text_tokens + c.num_proprio_tokens;
    h->n_total = h->n_itp + c.num_action_tokens;
    bool ok = true;
    // ---- weights ----
    h->embed = static_cast<bf16*>(dalloc(h, static_cast<size_t>(c.vocab_size) * c.vlm_hidden * 2));
    ok &= h->embed != nullptr;
    ok &= alloc_lin(h, h->patch, c.vision_hidden, 3 * c.patch_size * c.patch_size, true);
    h->pos_emb = alloc_vec(h, c.num_image_tokens * c.vision_hidden);
    h->vlayers.resize(c.vision_layers);
    for (auto& L : h->vlayers) {
        ok &= alloc_lin(h, L.qkv, 3 * c.vision_hidden, c.vision_hidden, true);
        ok &= alloc_lin(h, L.out, c.vision_hidden, c.vision_hidden, true);
        ok &= alloc_lin(h, L.fc1, c.vision_intermediate, c.vision_hidden, true);
        ok &= alloc_lin(h, L.fc2, c.vision_hidden, pad_to(c.vision_intermediate, 128), true);
        L.ln1w = alloc_vec(h, c.vision_hidden); L.ln1b = alloc_vec(h, c.vision_hidden);
        L.ln2w = alloc_vec(h, c.vision_hidden); L.ln2b = alloc_vec(h, c.vision_hidden);
        ok &= L.ln1w && L.ln1b && L.ln2w && L.ln2b;
    }
    h->post_ln_w = alloc_vec(h, c.vision_hidden);
    h->post_ln_b = alloc_vec(h, c.vision_hidden);
    ok &= alloc_lin(h, h->proj, c.vlm_hidden, c.vision_hidden, true);
    const int qkv_rows = (c.num_heads + 2) * c.head_dim;
    for (int m = 0; m < 3; ++m) {
        MixtureW& M = h->mix[m];
        M.name = kMixNames[m];
        M.hidden = (m == 0) ? c.vlm_hidden : c.expert_hidden;
        M.inter = (m == 0) ? c.vlm_intermediate : c.expert_intermediate;
        M.layers.resize(c.joint_layers);
        for (auto& L : M.layers) {
            ok &= alloc_lin(h, L.qkv, qkv_rows, M.hidden, false);
            ok &= alloc_lin(h, L.o, M.hidden, c.num_heads * c.head_dim, false);
            ok &= alloc_lin(h, L.gu, 2 * M.inter, M.hidden, false);
            ok &= alloc_lin(h, L.down, M.hidden, M.inter, false);
            L.in_ln = alloc_vec(h, M.hidden);
            L.post_ln = alloc_vec(h, M.hidden);
            ok &= L.in_ln && L.post_ln;
        }
        if (m > 0) { M.final_norm = alloc_vec(h, M.hidden); ok &= M.final_norm != nullptr; }
        M.cos_t = static_cast<float*>(dalloc(h, static_cast<size_t>(kNumPos) * 128 * 4));
        M.sin_t = static_cast<float*>(dalloc(h, static_cast<size_t>(kNumPos) * 128 * 4));
        ok &= M.cos_t && M.sin_t;
    }
    h->ae1_w = alloc_vec(h, c.expert_hidden * c.action_dim); h->ae1_b = alloc_vec(h, c.expert_hidden);
    ok &= alloc_lin(h, h->ae2, c.expert_hidden, 2 * c.expert_hidden, true);
    ok &= alloc_lin(h, h->ae3, c.expert_hidden, c.expert_hidden, true);
    h->pe_w = alloc_vec(h, c.expert_hidden * c.proprio_dim); h->pe_b = alloc_vec(h, c.expert_hidden);
    h->dec_w = alloc_vec(h, c.action_dim * c.expert_hidden); h->dec_b = alloc_vec(h, c.action_dim);
    // ---- static inputs / activations ----
    const size_t B = max_batch;
    const size_t Tv = B * c.num_image_tokens, Tt = B * c.max_image_text_tokens;
    const size_t Tp = B * c.num_proprio_tokens, Ta = B * c.num_action_tokens;
    auto bufb = [&](size_t elems) { return static_cast<bf16*>(dalloc(h, elems * 2)); };
    h->d_ids = static_cast<int64_t*>(dalloc(h, Tt * 8));
    h->d_vpos = static_cast<int64_t*>(dalloc(h, Tt * 8));
    h->d_ppos = static_cast<int64_t*>(dalloc(h, Tp * 8));
    h->d_apos = static_cast<int64_t*>(dalloc(h, Ta * 8));
    h->d_proprios = bufb(Tp * c.proprio_dim);
    h->clip_action = bufb(static_cast<size_t>(h->max_batch) * c.num_action_tokens * c.action_dim);
    h->d_action = bufb(Ta * c.action_dim);
    h->d_out = bufb(Ta * c.action_dim);
    h->d_mask_itp = bufb(static_cast<size_t>(B) * h->n_itp * ((h->n_itp + 7) / 8 * 8));
    h->d_mask_act = bufb(B * c.num_action_tokens * h->n_total);
    h->patches = bufb(Tv * h->patch.K);
    h->xs = bufb(Tv * c.vision_hidden); h->xn = bufb(Tv * c.vision_hidden);
    h->sqkv = bufb(Tv * 3 * c.vision_hidden); h->sattn = bufb(Tv * c.vision_hidden);
    h->shmid = bufb(Tv * h->vlayers[0].fc1.Nw); h->imgfeat = bufb(Tv * c.vlm_hidden);
    const int qw = c.num_heads * c.head_dim;
    h->E = bufb(Tt * c.vlm_hidden); h->En = bufb(Tt * c.vlm_hidden);
    h->Qv = bufb(Tt * qw); h->AOv = bufb(Tt * qw); h->H = bufb(Tt * c.vlm_intermediate);
    h->Ep = bufb(Tp * c.expert_hidden); h->Epn = bufb(Tp * c.expert_hidden);
    h->Qp = bufb(Tp * qw); h->AOp = bufb(Tp * qw); h->Hp = bufb(Tp * c.expert_intermediate);
    h->Ea = bufb(Ta * c.expert_hidden); h->Ean = bufb(Ta * c.expert_hidden);
    h->Qa = bufb(Ta * qw); h->AOa = bufb(Ta * qw); h->Ha = bufb(Ta * c.expert_intermediate);
    h->X2 = bufb(Ta * 2 * c.expert_hidden); h->A1 = bufb(Ta * c.expert_hidden);
    const size_t cache_elems = static_cast<size_t>(c.joint_layers) * B * h->n_total * c.head_dim;
    h->kcache = bufb(cache_elems); h->vcache = bufb(cache_elems);
    // split-K workspace: at most ~kNumSMs CTAs of 128 x (tokens per CTA) fp32 for small T, or one
    // full [T][Nw] slab for large T (see pick_splitk)
    size_t ws = 0;
    {
        const size_t maxNw = std::max<size_t>(qkv_rows, std::max<size_t>(c.vlm_hidden, 3 * c.vision_hidden));
        const size_t maxT = std::max(Tt, Tv);
        ws = std::max<size_t>(maxT * maxNw, static_cast<size_t>(kNumSMs + 20) * 128 * 512);
        ws = std::max<size_t>(ws, 16 * maxT * c.expert_hidden);
    }
    h->ws_floats = ws;
    h->ws = static_cast<float*>(dalloc(h, ws * 4));
    h->ws2_floats = std::max<size_t>(static_cast<size_t>(16) * Ta * (qkv_rows + 0), static_cast<size_t>(kNumSMs + 20) * 128 * 16);
    h->ws2 = static_cast<float*>(dalloc(h, h->ws2_floats * 4));
    h->ws3 = static_cast<float*>(dalloc(h, h->ws2_floats * 4));
    {
        // the expert streams run at the lowest priority: when SMs free up, pending CTAs of the main
        // (SigLIP / VLM) stream are placed first, so a 148-CTA VLM GEMM is not split into two waves by
        // an expert kernel that happened to be launched a moment earlier
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        bool ev_ok = cudaStreamCreateWithPriority(&h->s_prop, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                     cudaStreamCreateWithPriority(&h->s_act, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                     cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&h->ev_done_p, cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&h->ev_done_a, cudaEventDisableTiming) == cudaSuccess;
        h->ev_v.resize(c.joint_layers); h->ev_p.resize(c.joint_layers);
        for (int l = 0; l < c.joint_layers && ev_ok; ++l)
            ev_ok = cudaEventCreateWithFlags(&h->ev_v[l], cudaEventDisableTiming) == cudaSuccess &&
                    cudaEventCreateWithFlags(&h->ev_p[l], cudaEventDisableTiming) == cudaSuccess;
        ok &= ev_ok;
    }
    h->d_err = static_cast<int*>(dalloc(h, 16));
    h->flag_gemm = gemm_timeout_flag_ptr();
    h->flag_attn = attn_timeout_flag_ptr();
    ok &= h->ws && h->ws2 && h->ws3 && h->d_err && h->vcache && h->kcache && h->A1 && h->H && h->shmid && h->patches;
    if (!ok) {
        blurr_pi0_destroy(h);
        return fail(BLURR_ERR_CUDA, "blurr_pi0_create: device allocation failed");
    }
    build_expected(h);
    *out = h;
    return 0;
}

// ---------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------
static int repack(const bf16* src, int rows, int cols, int src_ld, bf16* dst, int dst_ld, int mode, int row_off,
                  int kb_total = 0) {
    const size_t total = static_cast<size_t>(rows) * cols;
    repack_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256>>>(src, rows, cols, src_ld, dst, dst_ld,
                                                                            mode, row_off, kb_total);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(BLURR_ERR_CUDA, std::string("repack: ") + cudaGetErrorString(e));
    return 0;
}
static bool shape_is(const int64_t* shape, int ndim, std::initializer_list<int64_t> want) {
    if (ndim != static_cast<int>(want.size())) return false;
    int i = 0;
    for (int64_t w : want) if (shape[i++] != w) return false;
    return true;
}
static int bad_shape(const std::string& key) { return fail(BLURR_ERR_INVALID, "unexpected shape for " + key); }

static bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }

extern "C" int blurr_pi0_set_weight(blurr_pi0_t* h, const char* key_c, const void* dev_ptr,
                                    const int64_t* shape, int ndim, int dtype) {
    if (!h || !key_c || !dev_ptr || !shape) return fail(BLURR_ERR_INVALID, "set_weight: null argument");
    if (dtype != BLURR_BF16)
        return fail(BLURR_ERR_INVALID, "set_weight: only bf16 weights are supported (cast the model with .to(torch.bfloat16))");
    CUDA_TRY(cudaSetDevice(h->device));
    const std::string key(key_c);
    if (!h->expected.count(key)) return fail(BLURR_ERR_INVALID, "set_weight: unknown state_dict key " + key);
    const auto& c = h->cfg;
    const bf16* src = static_cast<const bf16*>(dev_ptr);
    const int VH = c.vision_hidden, VI = c.vision_intermediate;
    int rc = 0;
    auto vec = [&](bf16* dst, int64_t n) -> int {
        if (!shape_is(shape, ndim, {n})) return bad_shape(key);
        return repack(src, 1, static_cast<int>(n), static_cast<int>(n), dst, static_cast<int>(n), MAP_OFFSET, 0);
    };
    auto mat = [&](Lin& L, int64_t rows, int64_t cols, int mode, int row_off) -> int {
        if (!shape_is(shape, ndim, {rows, cols})) return bad_shape(key);
        return repack(src, static_cast<int>(rows), static_cast<int>(cols), static_cast<int>(cols), L.w, 0, mode, row_off,
                      L.K / 64);
    };
    auto bias_rows = [&](Lin& L, int64_t n, int off) -> int {
        if (!shape_is(shape, ndim, {n})) return bad_shape(key);
        return repack(src, 1, static_cast<int>(n), static_cast<int>(n), L.bias + off, L.Nw, MAP_OFFSET, 0);
    };

    if (key == "embed_tokens.weight") {
        if (!shape_is(shape, ndim, {c.vocab_size, c.vlm_hidden})) return bad_shape(key);
        CUDA_TRY(cudaMemcpyAsync(h->embed, src, static_cast<size_t>(c.vocab_size) * c.vlm_hidden * 2,
                                 cudaMemcpyDeviceToDevice, 0));
    } else if (starts_with(key, kVT)) {
        const std::string k = key.substr(strlen(kVT));
        if (k == "embeddings.patch_embedding.weight") {
            if (!shape_is(shape, ndim, {VH, 3, c.patch_size, c.patch_size})) return bad_shape(key);
            const int kc = 3 * c.patch_size * c.patch_size;
            rc = repack(src, VH, kc, kc, h->patch.w, 0, MAP_OFFSET, 0, h->patch.K / 64);
        } else if (k == "embeddings.patch_embedding.bias") rc = bias_rows(h->patch, VH, 0);
        else if (k == "embeddings.position_embedding.weight") {
            if (!shape_is(shape, ndim, {c.num_image_tokens, VH})) return bad_shape(key);
            rc = repack(src, c.num_image_tokens, VH, VH, h->pos_emb, VH, MAP_OFFSET, 0);
        } else if (k == "post_layernorm.weight") rc = vec(h->post_ln_w, VH);
        else if (k == "post_layernorm.bias") rc = vec(h->post_ln_b, VH);
        else {
            int l = -1; char rest[128] = {0};
            if (sscanf(k.c_str(), "encoder.layers.%d.%127s", &l, rest) != 2 || l < 0 || l >= c.vision_layers)
                return fail(BLURR_ERR_INVALID, "set_weight: cannot parse " + key);
            VisionLayer& L = h->vlayers[l];
            const std::string r(rest);
            if (r == "self_attn.q_proj.weight") rc = mat(L.qkv, VH, VH, MAP_OFFSET, 0);
            else if (r == "self_attn.k_proj.weight") rc = mat(L.qkv, VH, VH, MAP_OFFSET, VH);
            else if (r == "self_attn.v_proj.weight") rc = mat(L.qkv, VH, VH, MAP_OFFSET, 2 * VH);
            else if (r == "self_attn.q_proj.bias") rc = bias_rows(L.qkv, VH, 0);
            else if (r == "self_attn.k_proj.bias") rc = bias_rows(L.qkv, VH, VH);
            else if (r == "self_attn.v_proj.bias") rc = bias_rows(L.qkv, VH, 2 * VH);
            else if (r == "self_attn.out_proj.weight") rc = mat(L.out, VH, VH, MAP_OFFSET, 0);
            else if (r == "self_attn.out_proj.bias") rc = bias_rows(L.out, VH, 0);
            else if (r == "layer_norm1.weight") rc = vec(L.ln1w, VH);
            else if (r == "layer_norm1.bias") rc = vec(L.ln1b, VH);
            else if (r == "layer_norm2.weight") rc = vec(L.ln2w, VH);
            else if (r == "layer_norm2.bias") rc = vec(L.ln2b, VH);
            else if (r == "mlp.fc1.weight") rc = mat(L.fc1, VI, VH, MAP_OFFSET, 0);
            else if (r == "mlp.fc1.bias") rc = bias_rows(L.fc1, VI, 0);
            else if (r == "mlp.fc2.weight") rc = mat(L.fc2, VH, VI, MAP_OFFSET, 0);
            else if (r == "mlp.fc2.bias") rc = bias_rows(L.fc2, VH, 0);
            else return fail(BLURR_ERR_INVALID, "set_weight: cannot parse " + key);
        }
    } else if (key == "multi_modal_projector.linear.weight") rc = mat(h->proj, c.vlm_hidden, VH, MAP_OFFSET, 0);
    else if (key == "multi_modal_projector.linear.bias") rc = bias_rows(h->proj, c.vlm_hidden, 0);
    else if (starts_with(key, "joint_model.mixtures.")) {
        char mname[16] = {0}; int l = -1; char rest[128] = {0};
        const char* s = key.c_str() + strlen("joint_model.mixtures.");
        int m = -1;
        for (int i = 0; i < 3; ++i)
            if (starts_with(s, std::string(kMixNames[i]) + ".")) m = i;
        if (m < 0) return fail(BLURR_ERR_INVALID, "set_weight: cannot parse " + key);
        (void)mname;
        MixtureW& M = h->mix[m];
        const std::string tail(s + strlen(kMixNames[m]) + 1);
        const int Hd = M.hidden, I = M.inter, QW = c.num_heads * c.head_dim, D = c.head_dim;
        if (tail == "norm.weight") rc = vec(M.final_norm, Hd);
        else {
            if (sscanf(tail.c_str(), "layers.%d.%127s", &l, rest) != 2 || l < 0 || l >= c.joint_layers)
                return fail(BLURR_ERR_INVALID, "set_weight: cannot parse " + key);
            MixLayer& L = M.layers[l];
            const std::string r(rest);
            if (r == "self_attn.q_proj.weight") rc = mat(L.qkv, QW, Hd, MAP_OFFSET, 0);
            else if (r == "self_attn.k_proj.weight") rc = mat(L.qkv, D, Hd, MAP_OFFSET, QW);
            else if (r == "self_attn.v_proj.weight") rc = mat(L.qkv, D, Hd, MAP_OFFSET, QW + D);
            else if (r == "self_attn.o_proj.weight") rc = mat(L.o, Hd, QW, MAP_OFFSET, 0);
            else if (r == "mlp.gate_proj.weight") rc = mat(L.gu, I, Hd, MAP_GATE, 0);
            else if (r == "mlp.up_proj.weight") rc = mat(L.gu, I, Hd, MAP_UP, 0);
            else if (r == "mlp.down_proj.weight") rc = mat(L.down, Hd, I, MAP_OFFSET, 0);
            else if (r == "input_layernorm.weight") rc = vec(L.in_ln, Hd);
            else if (r == "post_attention_layernorm.weight") rc = vec(L.post_ln, Hd);
            else return fail(BLURR_ERR_INVALID, "set_weight: cannot parse " + key);
        }
    } else if (key == "action_encoder.linear_1.weight") {
        if (!shape_is(shape, ndim, {c.expert_hidden, c.action_dim})) return bad_shape(key);
        rc = repack(src, c.expert_hidden, c.action_dim, c.action_dim, h->ae1_w, c.action_dim, MAP_OFFSET, 0);
    } else if (key == "action_encoder.linear_1.bias") rc = vec(h->ae1_b, c.expert_hidden);
    else if (key == "action_encoder.linear_2.weight") rc = mat(h->ae2, c.expert_hidden, 2 * c.expert_hidden, MAP_OFFSET, 0);
    else if (key == "action_encoder.linear_2.bias") rc = bias_rows(h->ae2, c.expert_hidden, 0);
    else if (key == "action_encoder.linear_3.weight") rc = mat(h->ae3, c.expert_hidden, c.expert_hidden, MAP_OFFSET, 0);
    else if (key == "action_encoder.linear_3.bias") rc = bias_rows(h->ae3, c.expert_hidden, 0);
    else if (key == "proprio_encoder.weight") {
        if (!shape_is(shape, ndim, {c.expert_hidden, c.proprio_dim})) return bad_shape(key);
        rc = repack(src, c.expert_hidden, c.proprio_dim, c.proprio_dim, h->pe_w, c.proprio_dim, MAP_OFFSET, 0);
    } else if (key == "proprio_encoder.bias") rc = vec(h->pe_b, c.expert_hidden);
    else if (key == "action_decoder.weight") {
        if (!shape_is(shape, ndim, {c.action_dim, c.expert_hidden})) return bad_shape(key);
        rc = repack(src, c.action_dim, c.expert_hidden, c.expert_hidden, h->dec_w, c.expert_hidden, MAP_OFFSET, 0);
    } else if (key == "action_decoder.bias") rc = vec(h->dec_b, c.action_dim);
    else return fail(BLURR_ERR_INVALID, "set_weight: unhandled key " + key);
    if (rc) return rc;
    h->seen.insert(key);
    h->finalized = false;
    return 0;
}

extern "C" int blurr_pi0_set_rope_inv_freq(blurr_pi0_t* h, const char* mixture, const float* inv, int n) {
    if (!h || !mixture || !inv || n != 128) return fail(BLURR_ERR_INVALID, "set_rope_inv_freq: need 128 values");
    for (int m = 0; m < 3; ++m)
        if (h->mix[m].name == mixture) {
            memcpy(h->mix[m].inv_freq, inv, 128 * sizeof(float));
            h->mix[m].have_inv_freq = true;
            h->finalized = false;
            return 0;
        }
    return fail(BLURR_ERR_INVALID, std::string("unknown mixture ") + mixture);
}

extern "C" int blurr_pi0_set_time_table(blurr_pi0_t* h, const void* dev_table, int num_steps) {
    if (!h || !dev_table || num_steps < 1) return fail(BLURR_ERR_INVALID, "set_time_table: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    if (num_steps > h->time_capacity) {      // capacity is tracked apart from the step count: 1 -> 10 -> 1 -> 10 allocates once
        h->time_table = static_cast<bf16*>(dalloc(h, static_cast<size_t>(num_steps) * h->cfg.expert_hidden * 2));
        if (!h->time_table) return fail(BLURR_ERR_CUDA, "set_time_table: allocation failed");
        h->time_capacity = num_steps;
    }
    h->time_steps = num_steps;
    CUDA_TRY(cudaMemcpy(h->time_table, dev_table, static_cast<size_t>(num_steps) * h->cfg.expert_hidden * 2,
                        cudaMemcpyDeviceToDevice));
    // a different step count changes the captured schedule
    for (auto& kv : h->graphs) {
        cudaGraphExecDestroy(kv.second.exec);
        cudaGraphDestroy(kv.second.graph);
    }
    h->graphs.clear();
    return 0;
}

extern "C" int blurr_pi0_finalize_weights(blurr_pi0_t* h) {
    if (!h) return fail(BLURR_ERR_INVALID, "finalize: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    for (const auto& k : h->expected)
        if (!h->seen.count(k)) return fail(BLURR_ERR_STATE, "finalize: missing weight " + k);
    for (int m = 0; m < 3; ++m) {
        MixtureW& M = h->mix[m];
        if (!M.have_inv_freq) return fail(BLURR_ERR_STATE, "finalize: missing RoPE inv_freq for " + M.name);
        float* d_inv = static_cast<float*>(dalloc(h, 128 * 4));
        if (!d_inv) return fail(BLURR_ERR_CUDA, "finalize: allocation failed");
        CUDA_TRY(cudaMemcpy(d_inv, M.inv_freq, 128 * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(launch_rope_table(0, d_inv, kNumPos, M.cos_t, M.sin_t));
    }
    CUDA_TRY(cudaDeviceSynchronize());
    h->finalized = true;
    return 0;
}

// ---------------------------------------------------------------------------
// the schedule
// ---------------------------------------------------------------------------
static constexpr int kTraceMax = 4096;
static constexpr int kLinModeMinTokens = 1024;   // above this a single-slice GEMM hands bf16 to its consumer

struct Run {
    blurr_pi0* h;
    cudaStream_t st;
    int rc = 0;
    unsigned long long* trace_slot(const char* what) {
        if (!h->trace || h->trace_buf == nullptr) return nullptr;
        const size_t idx = h->trace_labels.size();
        if (idx >= static_cast<size_t>(kTraceMax)) return nullptr;
        h->trace_labels.push_back(label.empty() ? std::string(what) : label + ":" + what);
        h->trace_streams.push_back(st == s_main ? 0 : (st == h->s_prop ? 1 : 2));
        return h->trace_buf + idx * 4;
    }
    std::string label;               // profile mode: current phase label
    void prof_begin(const char* what) {
        if (!h->profile) return;
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
        h->prof_pending.push_back({label.empty() ? std::string(what) : label + ":" + what, {a, b}});
    }
    void prof_end() {
        if (!h->profile || h->prof_pending.empty()) return;
        cudaEventRecord(h->prof_pending.back().second.second, st);
    }
    void launched(cudaError_t e, const char* what) {
        ++h->launches;
        if (e != cudaSuccess && rc == 0) rc = fail(BLURR_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    }
    int side_stream_cap = 0;         // > 0 while the experts overlap the prefill (see pick_splitk)
    // split-K so that about one CTA per SM is in flight (each CTA owns ~all of an SM's smem)
    int pick_splitk(int T, int Nw, int K, size_t ws_floats) const {
        GemmPlan p = gemm_make_plan(T, Nw, K, 1, EPI_PARTIAL, 0);
        if (!p.valid) return 1;
        const int ctas = p.grid_x * p.grid_y;
        int s = kNumSMs / ctas;
        if (s < 1) s = 1;
        const int max_by_k = p.kb_total / 2 > 0 ? p.kb_total / 2 : 1;
        if (s > max_by_k) s = max_by_k;
        if (s > 16) s = 16;
        // Expert GEMMs that run beside the VLM prefill (proprio expert, action expert of the first flow step):
        // at most 2 K slices.  With 16 they spread over every SM and keep evicting the main stream's
        // one-CTA-per-SM GEMMs into second waves; they have slack to spare (same-box A/B: 4.374 -> 4.279 ms).
        if (side_stream_cap > 0 && T <= 32 && s > side_stream_cap) s = side_stream_cap;
        while (s > 1 && static_cast<size_t>(s) * T * Nw > ws_floats) --s;
        return s;
    }
    // Split-K GEMMs of the batch-1 main stream (128..288 tokens): how many token chunks (CTAs along the tokens) and K
    // slices?  One chunk with many slices keeps the main loop shortest but every CTA then writes a 128 x T fp32 tile
    // (a CTA stores at ~55 GB/s: 2.4-2.8 us) and the consumer re-reads slices x T x N x 4 bytes; more chunks with fewer
    // slices shrink both at the price of a longer k-loop (>= 448 clk per k-block whatever the chunk width).  Cost model
    // from the per-CTA timelines (tools/cta_timeline.py); returns the slices, *bn_override = tokens per chunk (0 = one CTA
    // holds all tokens).
    int plan_partial(int T, int Nw, int K, size_t ws_floats, int* bn_override) const {
        *bn_override = 0;
        int best_s = pick_splitk(T, Nw, K, ws_floats);
        if (!h->chunked_splitk || side_stream_cap > 0 && T <= 32) return best_s;
        if (T < 128 || T > 288 || Nw % 128 != 0) return best_s;
        const int tiles = Nw / 128, kb = (K + 63) / 64;
        // CTA budget: while the expert streams run beside the prefill a GEMM that needs every SM gets split into two
        // waves by whatever expert kernel holds a few of them (in-graph trace: o-proj with 144 CTAs 13.7 us, with 128
        // CTAs 11.3 us), so leave them some room
        const int budget = side_stream_cap > 0 ? kNumSMs - 16 : kNumSMs;
        double best_cost = 1e30;
        int best_bn = 0;
        for (int chunks = 1; chunks <= 4; ++chunks) {
            const int bn = ((T + chunks - 1) / chunks + 15) / 16 * 16;
            if (chunks > 1 && bn < 64) break;
            int sl = budget / (tiles * chunks);
            if (sl < 1) break;
            if (sl > kb / 2) sl = kb / 2 > 0 ? kb / 2 : 1;
            if (sl > 16) sl = 16;
            while (sl > 1 && static_cast<size_t>(sl) * T * Nw > ws_floats) --sl;
            const int kb_per = (kb + sl - 1) / sl;
            sl = (kb + kb_per - 1) / kb_per;
            // time per k-block (tools/sweep_splitk.py, B200): two 144-token chunks in one CTA (T > 256, one CTA holds all
            // tokens) 0.40 us (shared-memory bound); one chunk per CTA 0.25 us whatever its width (>= 112 clk per MMA)
            const double us_kb = (chunks == 1 && T > 256) ? 0.40 : 0.25;
            const double main_us = kb_per * us_kb;
            const double tile_tokens = chunks == 1 ? (T > 256 ? 288 : (T + 15) / 16 * 16) : bn;
            const double store_us = 128.0 * tile_tokens * 4.0 / 55e3;                  // one CTA's fp32 tile at ~55 GB/s
            const double read_us = static_cast<double>(sl) * T * Nw * 4.0 / 6e6;       // the consumer's slices at ~6 TB/s
            const double cost = main_us + store_us + read_us;
            if (cost < best_cost) { best_cost = cost; best_s = sl; best_bn = chunks == 1 ? 0 : bn; }
        }
        *bn_override = best_bn;
        return best_s;
    }
    // `alt` selects the split-K workspace
    float* wsp(int alt) const { return alt == 0 ? h->ws : (alt == 1 ? h->ws2 : h->ws3); }
    // streams: 0 = caller's stream (SigLIP, VLM), 1 = proprio expert, 2 = action expert
    cudaStream_t s_main = nullptr;
    bool multi = false;
    void on(int which) { st = !multi ? s_main : (which == 0 ? s_main : (which == 1 ? h->s_prop : h->s_act)); }
    void record(cudaEvent_t ev) {
        if (multi && rc == 0 && cudaEventRecord(ev, st) != cudaSuccess) rc = fail(BLURR_ERR_CUDA, "cudaEventRecord failed");
    }
    void wait(cudaEvent_t ev) {
        if (multi && rc == 0 && cudaStreamWaitEvent(st, ev, 0) != cudaSuccess) rc = fail(BLURR_ERR_CUDA, "cudaStreamWaitEvent failed");
    }
    int gemm(const Lin& L, const bf16* X, int T, int epi, bf16* out, int ldo, bool bias = true, int alt = 0) {
        if (rc) return 1;
        float* ws = wsp(alt);
        const size_t ws_floats = alt ? h->ws2_floats : h->ws_floats;
        GemmCall c{};
        c.W = L.w; c.Nw = L.Nw; c.K = L.K; c.ldw = L.ld; c.w_packed = 1;
        c.X = X; c.T = T; c.ldx = L.K;
        c.epi = epi;
        c.bn_override = 0;
        c.splitk = (epi == EPI_PARTIAL) ? plan_partial(T, L.Nw, L.K, ws_floats, &c.bn_override) : 1;
        c.bias = bias ? L.bias : nullptr;
        c.out = out; c.ldo = ldo; c.partial = ws; c.w_static = 1;
        // Batched episodes: a GEMM that needs no K split does not need the fp32 round trip either.  Its store
        // epilogue writes the Linear output itself, bf16(acc + bias), into the workspace and the consumer reads
        // that (return value 0 = "bf16 linear output, bias applied"); same bits, half the bytes.
        bool lin_mode = false;
        if (h->lin_mode && epi == EPI_PARTIAL && c.splitk == 1 && T > kLinModeMinTokens) {
            lin_mode = true;
            c.epi = EPI_STORE;
            c.bias = L.bias;            // nullptr where the Linear has none
            c.out = reinterpret_cast<bf16*>(ws); c.ldo = L.Nw; c.partial = nullptr;
        }
        if (epi == EPI_PARTIAL && static_cast<size_t>(c.splitk) * T * L.Nw > ws_floats) {
            rc = fail(BLURR_ERR_STATE, "split-K workspace too small");
            return 1;
        }
        std::string err;
        char nm[96];
        snprintf(nm, sizeof nm, "gemm[epi%d T%d N%d K%d S%d]", c.epi, T, L.Nw, L.K, c.splitk);
        c.trace = trace_slot(nm);
        prof_begin(nm);
        const int s = gemm_launch(st, c, &err);
        prof_end();
        ++h->launches;
        if (s < 0) { rc = fail(BLURR_ERR_CUDA, err); return 1; }
        return lin_mode ? 0 : s;
    }
    ConsumerArgs consumer_args(int splitk, int T, int N, int ldp, const bf16* bias, int add_mode, const bf16* res, int ldr,
                               float out_scale, bf16* x_out, int norm_mode, const bf16* nw, const bf16* nb, float eps,
                               bf16* xn_out, bool use_partial = true, int alt = 0) const {
        ConsumerArgs a{};
        a.partial = use_partial ? wsp(alt) : nullptr; a.splitk = splitk; a.T = T; a.N = N; a.ldp = ldp;
        a.bias = bias; a.add_mode = add_mode; a.res = res; a.ldr = ldr;
        if (use_partial && splitk == 0) {            // the GEMM left bf16(acc + bias) in the workspace (see gemm())
            a.lin = reinterpret_cast<const bf16*>(wsp(alt)); a.ldl = ldp;
            a.partial = nullptr; a.bias = nullptr; a.splitk = 1;
        }
        a.pos = h->pos_emb; a.pos_rows = h->cfg.num_image_tokens; a.out_scale = out_scale;
        a.x_out = x_out; a.ldx = N; a.norm_mode = norm_mode; a.norm_w = nw; a.norm_b = nb; a.eps = eps;
        a.xn_out = xn_out; a.ldn = N;
        return a;
    }
    void consumer(int splitk, int T, int N, int ldp, const bf16* bias, int add_mode, const bf16* res, int ldr,
                  float out_scale, bf16* x_out, int norm_mode, const bf16* nw, const bf16* nb, float eps,
                  bf16* xn_out, bool use_partial = true, int alt = 0) {
        if (rc) return;
        ConsumerArgs a = consumer_args(splitk, T, N, ldp, bias, add_mode, res, ldr, out_scale, x_out, norm_mode, nw, nb, eps,
                                       xn_out, use_partial, alt);
        { char nm[64]; snprintf(nm, sizeof nm, "consumer[T%d N%d S%d]", T, N, splitk); a.trace = trace_slot(nm); prof_begin(nm); launched(launch_consumer(st, a), "consumer"); prof_end(); }
    }
    // Split-K GEMM + its consumer kernel.
    void gemm_consumer(const Lin& L, const bf16* X, int T, int alt, int N, const bf16* bias, int add_mode, const bf16* res,
                       int ldr, float out_scale, bf16* x_out, int norm_mode, const bf16* nw, const bf16* nb, float eps,
                       bf16* xn_out, bool gemm_bias = false) {
        if (rc) return;
        const int s = gemm(L, X, T, EPI_PARTIAL, nullptr, 0, gemm_bias, alt);
        consumer(s, T, N, L.Nw, bias, add_mode, res, ldr, out_scale, x_out, norm_mode, nw, nb, eps, xn_out, true, alt);
    }
    void bias_act(int splitk, int T, int N, int ldp, const bf16* bias, int act, float scale, bf16* out, int ldo, int alt) {
        if (rc) return;
        { prof_begin("bias_act"); launched(launch_bias_act(st, wsp(alt), splitk, T, N, ldp, bias, act, scale, out, ldo), "bias_act"); prof_end(); }
    }
    void rope(RopeKvArgs a) {
        if (rc) return;
        { char nm[64]; snprintf(nm, sizeof nm, "rope_kv[T%d]", a.T); a.trace = trace_slot(nm); prof_begin(nm); launched(launch_rope_kv(st, a), "rope_kv"); prof_end(); }
    }
    void attn_siglip(const bf16* qkv, int ld_qkv, int B, int seq, int heads, int hidden, bf16* out, int ld_out) {
        if (rc) return;
        { prof_begin("siglip_attention"); launched(launch_siglip_attention(st, qkv, ld_qkv, B, seq, heads, hidden, out, ld_out, trace_slot("siglip_attention")), "siglip_attention"); prof_end(); }
    }
    void attn_joint(JointAttnArgs a, bool fewq) {
        if (rc) return;
        a.trace = nullptr;
        if (fewq) { char nm[64]; snprintf(nm, sizeof nm, "attention_fewq[q%d]", a.q_per_sample); a.trace = trace_slot(nm); prof_begin(nm); launched(launch_joint_attention_fewq(st, a), "attention_fewq"); prof_end(); }
        else { a.trace = trace_slot("attention_prefill"); prof_begin("attention_prefill"); launched(launch_joint_attention_prefill(st, a), "attention_prefill"); prof_end(); }
    }
    void embed_merge(const EmbedMergeArgs& a, int B) {
        if (rc) return;
        { prof_begin("embed_merge"); launched(launch_embed_merge(st, a.ids, B, a.seq, a.table, a.vocab, a.img, a.n_img, a.hidden, a.image_token,
                                         a.pad_token, a.inv_div, a.normalizer, a.out, a.err_flag), "embed_merge"); prof_end(); }
    }
    void small_k(const SmallKArgs& a) {
        if (rc) return;
        { prof_begin("small_k_linear"); launched(launch_small_k_linear(st, a.x, a.T, a.K, a.W, a.bias, a.N, a.scale, a.y, a.ldy, a.col_off, a.time_row,
                                            a.time_cols), "small_k_linear"); prof_end(); }
    }
    void action_tail(const ActionTailArgs& a) {
        if (rc) return;
        { prof_begin("action_tail"); launched(launch_action_tail(st, a.xn, a.T, a.hidden, a.W, a.bias, a.action_dim, a.dt, a.action, a.vel_tap),
                      "action_tail"); prof_end(); }
    }
    bool clips(int bit) const { return h->act_clip > 0.f && ((h->act_clip_mask >> bit) & 1); }
    // torch.clamp(x, -clip, clip) in front of a quantised Linear (int8_linear.py:73-74); dst == src clamps in place
    void clip_act(int bit, const bf16* src, bf16* dst, size_t n) {
        if (rc || !clips(bit)) return;
        { prof_begin("activation_clip"); launched(launch_clamp_copy(st, src, dst, static_cast<int>(n), 1, h->act_clip, nullptr, nullptr, nullptr), "activation_clip"); prof_end(); }
    }
    void clamp(const ClampArgs& a) {
        if (rc) return;
        { prof_begin("clamp"); launched(launch_clamp_copy(st, a.src, a.dst, a.n, a.do_clamp, a.clip, h->d_err, h->flag_gemm, h->flag_attn), "clamp"); prof_end(); }
    }
    void tap(const std::string& name, const void* src, size_t bytes) {
        if (!h->debug || rc) return;
        auto it = h->taps.find(name);
        if (it == h->taps.end() || it->second.bytes != bytes) {
            if (it != h->taps.end()) cudaFree(it->second.ptr);
            void* p = nullptr;
            if (cudaMalloc(&p, bytes) != cudaSuccess) { rc = fail(BLURR_ERR_CUDA, "tap allocation failed"); return; }
            h->taps[name] = TapBuf{p, bytes};
            it = h->taps.find(name);
        }
        cudaMemcpyAsync(it->second.ptr, src, bytes, cudaMemcpyDeviceToDevice, st);
    }
};

// SigLIP + projector + embedding merge (pizero.py:433-471; siglip.py)
static void run_vision(Run& R, int B) {
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    const int Tv = B * c.num_image_tokens, VH = c.vision_hidden;
    const float eps = c.layer_norm_eps;
    int s = R.gemm(h->patch, h->patches, Tv, EPI_PARTIAL, nullptr, 0);
    R.consumer(s, Tv, VH, h->patch.Nw, h->patch.bias, ADD_POSEMB, nullptr, 0, 1.0f, h->xs, NORM_LAYERNORM,
               h->vlayers[0].ln1w, h->vlayers[0].ln1b, eps, h->xn);
    R.tap("siglip.embeddings", h->xs, static_cast<size_t>(Tv) * VH * 2);
    for (int l = 0; l < c.vision_layers; ++l) {
        VisionLayer& L = h->vlayers[l];
        R.gemm(L.qkv, h->xn, Tv, EPI_STORE, h->sqkv, 3 * VH);
        R.attn_siglip(h->sqkv, 3 * VH, B, c.num_image_tokens, c.vision_heads, VH, h->sattn, VH);
        s = R.gemm(L.out, h->sattn, Tv, EPI_PARTIAL, nullptr, 0);
        R.consumer(s, Tv, VH, L.out.Nw, L.out.bias, ADD_RESIDUAL, h->xs, VH, 1.0f, h->xs, NORM_LAYERNORM, L.ln2w,
                   L.ln2b, eps, h->xn);
        R.gemm(L.fc1, h->xn, Tv, EPI_GELU, h->shmid, L.fc1.Nw);
        s = R.gemm(L.fc2, h->shmid, Tv, EPI_PARTIAL, nullptr, 0);
        const bool last = (l == c.vision_layers - 1);
        const bf16* nw = last ? h->post_ln_w : h->vlayers[l + 1].ln1w;
        const bf16* nb = last ? h->post_ln_b : h->vlayers[l + 1].ln1b;
        R.consumer(s, Tv, VH, L.fc2.Nw, L.fc2.bias, ADD_RESIDUAL, h->xs, VH, 1.0f, h->xs, NORM_LAYERNORM, nw, nb,
                   eps, h->xn);
        R.tap("siglip.layer" + std::to_string(l), h->xs, static_cast<size_t>(Tv) * VH * 2);
    }
    R.tap("siglip.post_layernorm", h->xn, static_cast<size_t>(Tv) * VH * 2);
    R.gemm(h->proj, h->xn, Tv, EPI_STORE, h->imgfeat, c.vlm_hidden);
    R.tap("projector", h->imgfeat, static_cast<size_t>(Tv) * c.vlm_hidden * 2);
    // image_features / (hidden ** 0.5): ATen multiplies by the fp32 reciprocal of the Python scalar;
    // normalizer: torch.tensor(hidden ** 0.5, dtype=bf16)
    const float inv_div = 1.0f / static_cast<float>(std::sqrt(static_cast<double>(c.vlm_hidden)));
    const float normalizer = __bfloat162float(__float2bfloat16(static_cast<float>(std::sqrt(static_cast<double>(c.vlm_hidden)))));
    EmbedMergeArgs em{h->d_ids, c.max_image_text_tokens, h->embed, c.vocab_size, h->imgfeat, c.num_image_tokens,
                      c.vlm_hidden, c.image_token_index, c.pad_token_id, inv_div, normalizer, h->E, h->d_err};
    R.embed_merge(em, B);
    R.tap("merged_embeds", h->E, static_cast<size_t>(B) * c.max_image_text_tokens * c.vlm_hidden * 2);
}

struct StreamBufs {
    bf16 *x, *xn, *q, *ao, *hmid;
    int tokens_per_sample, slot_base, q_row_offset;
    const int64_t* pos;
};

// One mixture's share of a joint layer (joint_model.py:24-129), split into its dependent phases so
// that the same phase of two independent streams (VLM and proprio) can share a grid barrier.
static Lin qkv_lin(blurr_pi0* h, int m, int l, bool kv_only) {
    const auto& c = h->cfg;
    MixLayer& L = h->mix[m].layers[l];
    Lin qkv = L.qkv;
    if (kv_only) {                      // last layer of vlm/proprio: only K and V are needed
        const int QW = c.num_heads * c.head_dim;
        qkv.w = L.qkv.w + static_cast<size_t>(QW) * L.qkv.K;    // tile-packed: whole 128-row tiles are contiguous
        qkv.Nw = L.qkv.Nw - QW;
    }
    return qkv;
}
static RopeKvArgs rope_args(Run& R, int m, int l, const StreamBufs& sb, int B, bool kv_only, int s, const Lin& qkv, int alt) {
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    MixtureW& M = h->mix[m];
    RopeKvArgs a{};
    a.partial = R.wsp(alt); a.splitk = s; a.T = B * sb.tokens_per_sample; a.ldp = qkv.Nw;
    if (s == 0) { a.lin = reinterpret_cast<const bf16*>(R.wsp(alt)); a.ldl = qkv.Nw; a.splitk = 1; }
    a.n_heads = kv_only ? 0 : c.num_heads;
    a.tokens_per_sample = sb.tokens_per_sample; a.position_ids = sb.pos;
    a.cos_table = M.cos_t; a.sin_table = M.sin_t; a.n_pos = kNumPos;
    a.q_out = kv_only ? nullptr : sb.q;
    const size_t layer_off = static_cast<size_t>(l) * h->max_batch * h->n_total * c.head_dim;
    a.k_cache = h->kcache + layer_off; a.v_cache = h->vcache + layer_off;
    a.n_slots = h->n_total; a.slot_base = sb.slot_base;
    return a;
}
static int phase_qkv_gemm(Run& R, int m, int l, const StreamBufs& sb, int B, bool kv_only, int alt, Lin* used) {
    *used = qkv_lin(R.h, m, l, kv_only);
    return R.gemm(*used, sb.xn, B * sb.tokens_per_sample, EPI_PARTIAL, nullptr, 0, false, alt);
}
static void phase_rope(Run& R, int m, int l, const StreamBufs& sb, int B, bool kv_only, int s, const Lin& qkv, int alt) {
    if (R.rc) return;
    R.rope(rope_args(R, m, l, sb, B, kv_only, s, qkv, alt));
}
// q/k/v projection + RoPE + cache write
static void phase_qkv_rope(Run& R, int m, int l, const StreamBufs& sb, int B, bool kv_only, int alt) {
    if (R.rc) return;
    const Lin qkv = qkv_lin(R.h, m, l, kv_only);
    const int s = R.gemm(qkv, sb.xn, B * sb.tokens_per_sample, EPI_PARTIAL, nullptr, 0, false, alt);
    if (R.rc) return;
    R.rope(rope_args(R, m, l, sb, B, kv_only, s, qkv, alt));
}

static void phase_attn(Run& R, int l, const StreamBufs& sb, int B, int n_keys, const bf16* mask, long long mbs,
                       long long mrs, bool fewq) {
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    if (R.rc) return;
    JointAttnArgs a{};
    a.q = sb.q; a.q_per_sample = sb.tokens_per_sample; a.q_row_offset = sb.q_row_offset;
    const size_t layer_off = static_cast<size_t>(l) * h->max_batch * h->n_total * c.head_dim;
    a.k_cache = h->kcache + layer_off; a.v_cache = h->vcache + layer_off;
    a.n_slots = h->n_total; a.n_keys = n_keys;
    a.mask = mask; a.mask_bstride = mbs; a.mask_rstride = mrs;
    a.batch = B; a.n_heads = c.num_heads; a.out = sb.ao;
    R.attn_joint(a, fewq);
}

static int phase_o_gemm(Run& R, int m, int l, const StreamBufs& sb, int B, int alt) {
    return R.gemm(R.h->mix[m].layers[l].o, sb.ao, B * sb.tokens_per_sample, EPI_PARTIAL, nullptr, 0, false, alt);
}
static void phase_post_attn(Run& R, int m, int l, const StreamBufs& sb, int B, int s, int alt) {
    MixtureW& M = R.h->mix[m];
    MixLayer& L = M.layers[l];
    R.consumer(s, B * sb.tokens_per_sample, M.hidden, L.o.Nw, nullptr, ADD_RESIDUAL, sb.x, M.hidden, 1.0f, sb.x,
               NORM_RMS_GEMMA, L.post_ln, nullptr, R.h->cfg.rms_norm_eps, sb.xn, true, alt);
}
static void phase_gate_up(Run& R, int m, int l, const StreamBufs& sb, int B) {
    MixtureW& M = R.h->mix[m];
    R.gemm(M.layers[l].gu, sb.xn, B * sb.tokens_per_sample, EPI_GEGLU, sb.hmid, M.inter, false);
}
static int phase_down(Run& R, int m, int l, const StreamBufs& sb, int B, int alt) {
    return R.gemm(R.h->mix[m].layers[l].down, sb.hmid, B * sb.tokens_per_sample, EPI_PARTIAL, nullptr, 0, false, alt);
}
static void phase_post_mlp(Run& R, int m, int l, const StreamBufs& sb, int B, int s, const bf16* next_norm, int alt) {
    MixtureW& M = R.h->mix[m];
    MixLayer& L = M.layers[l];
    R.consumer(s, B * sb.tokens_per_sample, M.hidden, L.down.Nw, nullptr, ADD_RESIDUAL, sb.x, M.hidden, 1.0f, sb.x,
               next_norm ? NORM_RMS_GEMMA : NORM_NONE, next_norm, nullptr, R.h->cfg.rms_norm_eps,
               next_norm ? sb.xn : nullptr, true, alt);
}

// Layer l of one expert stream (mixture m, workspace `alt`), split at the point where it needs the
// K/V of the streams before it.
static void expert_layer_head(Run& R, int m, int l, const StreamBufs& sb, int B, bool kv_only, int alt) {
    R.clip_act(m, sb.xn, sb.xn, static_cast<size_t>(B) * sb.tokens_per_sample * R.h->mix[m].hidden);     // q/k/v_proj input
    phase_qkv_rope(R, m, l, sb, B, kv_only, alt);
}
static void expert_layer_tail(Run& R, int m, int l, const StreamBufs& sb, int B, int n_keys, const bf16* mask,
                              long long mbs, long long mrs, const bf16* next_norm, int alt) {
    MixtureW& M = R.h->mix[m];
    MixLayer& L = M.layers[l];
    const int T = B * sb.tokens_per_sample;
    const float eps = R.h->cfg.rms_norm_eps;
    phase_attn(R, l, sb, B, n_keys, mask, mbs, mrs, true);
    R.clip_act(m, sb.ao, sb.ao, static_cast<size_t>(T) * L.o.K);                  // o_proj input
    R.gemm_consumer(L.o, sb.ao, T, alt, M.hidden, nullptr, ADD_RESIDUAL, sb.x, M.hidden, 1.0f, sb.x, NORM_RMS_GEMMA,
                    L.post_ln, nullptr, eps, sb.xn);
    R.clip_act(m, sb.xn, sb.xn, static_cast<size_t>(T) * M.hidden);              // gate/up_proj input
    phase_gate_up(R, m, l, sb, B);
    R.clip_act(m, sb.hmid, sb.hmid, static_cast<size_t>(T) * M.inter);           // down_proj input
    R.gemm_consumer(L.down, sb.hmid, T, alt, M.hidden, nullptr, ADD_RESIDUAL, sb.x, M.hidden, 1.0f, sb.x,
                    next_norm ? NORM_RMS_GEMMA : NORM_NONE, next_norm, nullptr, eps, next_norm ? sb.xn : nullptr);
}

static void action_encode(Run& R, int B, int s) {
    // ActionEncoder (vla/modules.py:39-53) + `*= sqrt(1024)` + input norm of layer 0
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    const int Ta = B * c.num_action_tokens;
    const float expert_norm = __bfloat162float(__float2bfloat16(static_cast<float>(std::sqrt(static_cast<double>(c.expert_hidden)))));
    const bf16* act_in = h->d_action;
    if (R.clips(3)) { R.clip_act(3, h->d_action, h->clip_action, static_cast<size_t>(Ta) * c.action_dim); act_in = h->clip_action; }
    SmallKArgs a1{act_in, Ta, c.action_dim, h->ae1_w, h->ae1_b, c.expert_hidden, 1.0f, h->X2,
                  2 * c.expert_hidden, c.expert_hidden, h->time_table + static_cast<size_t>(s) * c.expert_hidden,
                  c.expert_hidden};
    R.small_k(a1);
    R.clip_act(3, h->X2, h->X2, static_cast<size_t>(Ta) * 2 * c.expert_hidden);
    int k = R.gemm(h->ae2, h->X2, Ta, EPI_PARTIAL, nullptr, 0, true, 2);
    R.bias_act(k, Ta, c.expert_hidden, h->ae2.Nw, h->ae2.bias, ACT_SILU, 1.0f, h->A1, c.expert_hidden, 2);
    R.clip_act(3, h->A1, h->A1, static_cast<size_t>(Ta) * c.expert_hidden);
    R.gemm_consumer(h->ae3, h->A1, Ta, 2, c.expert_hidden, h->ae3.bias, ADD_NONE, nullptr, 0, expert_norm, h->Ea,
                    NORM_RMS_GEMMA, h->mix[2].layers[0].in_ln, nullptr, c.rms_norm_eps, h->Ean, true);
    R.tap("flow" + std::to_string(s) + ".action_embeds", h->Ea, static_cast<size_t>(Ta) * c.expert_hidden * 2);
}

static void action_decode(Run& R, int B, int s, float dt) {
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    const int Ta = B * c.num_action_tokens;
    bf16* vel_tap = nullptr;
    if (h->debug) {
        const std::string nm = "flow" + std::to_string(s) + ".velocity";
        R.tap(nm, h->d_action, static_cast<size_t>(Ta) * c.action_dim * 2);   // allocates the slot
        if (!R.rc) vel_tap = static_cast<bf16*>(h->taps[nm].ptr);
    }
    ActionTailArgs at{h->Ean, Ta, c.expert_hidden, h->dec_w, h->dec_b, c.action_dim, dt, h->d_action, vel_tap};
    R.action_tail(at);
}

static void run_step(Run& R, int B, int steps) {
    blurr_pi0* h = R.h;
    const auto& c = h->cfg;
    if (h->trace && h->trace_buf != nullptr) {
        // every issue / capture of the step hands out the same slot sequence; each replay re-arms it
        h->trace_labels.clear();
        h->trace_streams.clear();
        if (cudaMemsetAsync(h->trace_buf, 0xFF, static_cast<size_t>(kTraceMax) * 4 * sizeof(unsigned long long), R.st) != cudaSuccess)
            R.rc = fail(BLURR_ERR_CUDA, "trace buffer reset failed");
    }
    const int L = c.joint_layers;
    const int Tp = B * c.num_proprio_tokens, Ta = B * c.num_action_tokens, Tt = B * c.max_image_text_tokens;
    const float expert_norm = __bfloat162float(__float2bfloat16(static_cast<float>(std::sqrt(static_cast<double>(c.expert_hidden)))));
    const bool do_prefill = (h->stage_mask & 2) != 0, do_action = (h->stage_mask & 4) != 0;
    const float dt = static_cast<float>(1.0 / static_cast<double>(steps));

    StreamBufs sv{h->E, h->En, h->Qv, h->AOv, h->H, c.max_image_text_tokens, 0, 0, h->d_vpos};
    StreamBufs sp{h->Ep, h->Epn, h->Qp, h->AOp, h->Hp, c.num_proprio_tokens, c.max_image_text_tokens,
                  c.max_image_text_tokens, h->d_ppos};
    StreamBufs sa{h->Ea, h->Ean, h->Qa, h->AOa, h->Ha, c.num_action_tokens, h->n_itp, 0, h->d_apos};
    const long long itp_rs = (h->n_itp + 7) / 8 * 8, itp_bs = static_cast<long long>(h->n_itp) * itp_rs;
    const long long act_bs = static_cast<long long>(c.num_action_tokens) * h->n_total, act_rs = h->n_total;

    // fork: the expert streams start once the inputs are staged
    R.on(0);
    R.record(h->ev_fork);
    R.on(1); R.wait(h->ev_fork);
    R.on(2); R.wait(h->ev_fork);

    // ---- expert prologues (independent of the image) ----
    if (do_prefill) {
        R.on(1);
        R.label = "prefill";
        // proprio_encoder (pizero.py:493) and `*= sqrt(1024)` (joint_model.py:358-365)
        SmallKArgs pe{h->d_proprios, Tp, c.proprio_dim, h->pe_w, h->pe_b, c.expert_hidden, expert_norm, h->Ep,
                      c.expert_hidden, 0, nullptr, 0};
        R.small_k(pe);
        R.consumer(1, Tp, c.expert_hidden, 0, nullptr, ADD_NONE, h->Ep, c.expert_hidden, 1.0f, nullptr, NORM_RMS_GEMMA,
                   h->mix[1].layers[0].in_ln, nullptr, c.rms_norm_eps, h->Epn, false);
    }
    if (do_action) {
        R.on(2);
        R.label = "action";
        action_encode(R, B, 0);
    }

    // ---- SigLIP + projector + merge (main stream) ----
    R.on(0);
    R.label = "siglip";
    if (h->stage_mask & 1) run_vision(R, B);

    // ---- joint layers: VLM (main), proprio and action flow step 0 one dependency behind ----
    R.label = "prefill";
    if (do_prefill)
        R.consumer(1, Tt, c.vlm_hidden, 0, nullptr, ADD_NONE, h->E, c.vlm_hidden, 1.0f, nullptr, NORM_RMS_GEMMA,
                   h->mix[0].layers[0].in_ln, nullptr, c.rms_norm_eps, h->En, false);
    R.side_stream_cap = do_prefill ? 2 : 0;      // in every launch mode, so that all of them produce the same bits
    for (int l = 0; l < L; ++l) {
        const bool last = (l == L - 1);
        if (do_prefill) {
            // VLM: K/V of layer l first (the experts wait for them), then the rest of the layer
            R.on(0); R.label = "prefill";
            {
                Lin qv;
                const int s_v = phase_qkv_gemm(R, 0, l, sv, B, last, 0, &qv);
                phase_rope(R, 0, l, sv, B, last, s_v, qv, 0);
                R.record(h->ev_v[l]);
            }
            R.on(1);
            expert_layer_head(R, 1, l, sp, B, last, 1);
            R.record(h->ev_p[l]);
            if (!last) {
                R.on(0);
                phase_attn(R, l, sv, B, h->n_itp, h->d_mask_itp, itp_bs, itp_rs, false);
                const int o_v = phase_o_gemm(R, 0, l, sv, B, 0);
                phase_post_attn(R, 0, l, sv, B, o_v, 0);
                phase_gate_up(R, 0, l, sv, B);
                const int d_v = phase_down(R, 0, l, sv, B, 0);
                phase_post_mlp(R, 0, l, sv, B, d_v, h->mix[0].layers[l + 1].in_ln, 0);
                R.tap("prefill.L" + std::to_string(l) + ".vlm", h->E, static_cast<size_t>(Tt) * c.vlm_hidden * 2);
                R.on(1);
                R.wait(h->ev_v[l]);          // the proprio query attends over the VLM keys of this layer
                expert_layer_tail(R, 1, l, sp, B, h->n_itp, h->d_mask_itp, itp_bs, itp_rs, h->mix[1].layers[l + 1].in_ln, 1);
                R.tap("prefill.L" + std::to_string(l) + ".proprio", h->Ep, static_cast<size_t>(Tp) * c.expert_hidden * 2);
            }
        }
        if (do_action) {
            R.on(2); R.label = "action";
            expert_layer_head(R, 2, l, sa, B, false, 2);
            if (do_prefill) { R.wait(h->ev_v[l]); R.wait(h->ev_p[l]); }
            if (l == L - 1) R.side_stream_cap = 0;       // the VLM is done: the tail of the action expert runs alone
            const bf16* next = (l + 1 < L) ? h->mix[2].layers[l + 1].in_ln : h->mix[2].final_norm;
            expert_layer_tail(R, 2, l, sa, B, h->n_total, h->d_mask_act, act_bs, act_rs, next, 2);
            R.tap("flow0.L" + std::to_string(l) + ".action", h->Ea, static_cast<size_t>(Ta) * c.expert_hidden * 2);
        }
    }
    if (do_action) { R.on(2); action_decode(R, B, 0, dt); }
    // join
    R.on(1); R.record(h->ev_done_p);
    R.on(2); R.record(h->ev_done_a);
    R.on(0); R.wait(h->ev_done_p); R.wait(h->ev_done_a);

    R.side_stream_cap = 0;
    // ---- remaining Euler steps of the flow (pizero.py:516-538): sequential over the finished cache ----
    R.label = "action";
    for (int s = 1; s < (do_action ? steps : 0); ++s) {
        action_encode(R, B, s);
        for (int l = 0; l < L; ++l) {
            expert_layer_head(R, 2, l, sa, B, false, 2);
            const bf16* next = (l + 1 < L) ? h->mix[2].layers[l + 1].in_ln : h->mix[2].final_norm;
            expert_layer_tail(R, 2, l, sa, B, h->n_total, h->d_mask_act, act_bs, act_rs, next, 2);
            R.tap("flow" + std::to_string(s) + ".L" + std::to_string(l) + ".action", h->Ea,
                  static_cast<size_t>(Ta) * c.expert_hidden * 2);
        }
        action_decode(R, B, s, dt);
    }
    ClampArgs cl{h->d_action, h->d_out, Ta * c.action_dim, c.has_clip, c.final_action_clip_value};
    R.clamp(cl);
}

extern "C" int blurr_pi0_infer_action(blurr_pi0_t* h, void* cuda_stream, int batch, const blurr_pi0_inputs* in,
                                      void* actions_out) {
    if (!h || !in || !actions_out) return fail(BLURR_ERR_INVALID, "infer_action: null argument");
    if (!h->finalized) return fail(BLURR_ERR_STATE, "infer_action: call blurr_pi0_finalize_weights first");
    if (batch < 1 || batch > h->max_batch) return fail(BLURR_ERR_INVALID, "infer_action: batch out of range");
    const auto& c = h->cfg;
    const int steps = c.num_inference_steps;
    if (!h->time_table || h->time_steps != steps)
        return fail(BLURR_ERR_STATE, "infer_action: time table missing or not matching num_inference_steps");
    if (!in->input_ids || !in->pixel_values || !in->image_text_proprio_mask || !in->action_mask ||
        !in->vlm_position_ids || !in->proprio_position_ids || !in->action_position_ids || !in->proprios || !in->noise)
        return fail(BLURR_ERR_INVALID, "infer_action: null input tensor");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    h->launches = 0;

    // stage the per-call inputs (outside the graph: their addresses change from call to call)
    StageArgs sa{};
    sa.ids = in->input_ids; sa.vpos = in->vlm_position_ids; sa.ppos = in->proprio_position_ids;
    sa.apos = in->action_position_ids;
    sa.proprios = static_cast<const bf16*>(in->proprios); sa.noise = static_cast<const bf16*>(in->noise);
    sa.mask_itp = static_cast<const bf16*>(in->image_text_proprio_mask);
    sa.itp_bs = in->itp_mask_bstride; sa.itp_rs = in->itp_mask_rstride;
    sa.mask_act = static_cast<const bf16*>(in->action_mask);
    sa.act_bs = in->action_mask_bstride; sa.act_rs = in->action_mask_rstride;
    sa.d_ids = h->d_ids; sa.d_vpos = h->d_vpos; sa.d_ppos = h->d_ppos; sa.d_apos = h->d_apos;
    sa.d_proprios = h->d_proprios; sa.d_action = h->d_action; sa.d_mask_itp = h->d_mask_itp;
    sa.d_mask_act = h->d_mask_act;
    sa.n_ids = batch * c.max_image_text_tokens; sa.n_ppos = batch * c.num_proprio_tokens;
    sa.n_apos = batch * c.num_action_tokens; sa.n_prop = batch * c.num_proprio_tokens * c.proprio_dim;
    sa.n_noise = batch * c.num_action_tokens * c.action_dim;
    sa.itp_dim = h->n_itp; sa.itp_ld = (h->n_itp + 7) / 8 * 8; sa.act_rows = c.num_action_tokens; sa.act_cols = h->n_total; sa.batch = batch;
    const long long total = static_cast<long long>(sa.n_ids) + sa.n_ppos + sa.n_apos + sa.n_prop + sa.n_noise +
                            static_cast<long long>(batch) * h->n_itp * h->n_itp +
                            static_cast<long long>(batch) * c.num_action_tokens * h->n_total;
    stage_inputs_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(sa);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(launch_im2col(st, static_cast<const bf16*>(in->pixel_values), in->pixel_strides[0], in->pixel_strides[1],
                           in->pixel_strides[2], in->pixel_strides[3], batch, h->patches, h->patch.K));
    int64_t pre_launches = 2;

    Run R{h, st};
    R.s_main = st;
    R.multi = h->use_streams && !h->debug && !h->profile;
    const bool graph = h->use_graph && !h->debug && !h->profile;
    if (!graph) {
        run_step(R, batch, steps);
        if (R.rc) return R.rc;
        h->launches += pre_launches;
        if (h->profile) {
            CUDA_TRY(cudaStreamSynchronize(st));
            for (auto& p : h->prof_pending) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, p.second.first, p.second.second);
                auto& e = h->prof[p.first];
                e.count += 1; e.ms += ms;
                cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second);
            }
            h->prof_pending.clear();
        }
    } else {
        const long long key = static_cast<long long>(batch) * 4096 + steps;
        auto it = h->graphs.find(key);
        if (it == h->graphs.end()) {
            // warm run outside capture: first-use attribute setup and tensor-map creation
            run_step(R, batch, steps);
            if (R.rc) return R.rc;
            CUDA_TRY(cudaStreamSynchronize(st));
            h->launches = 0;
            cudaStream_t cs;
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
                // the captured kernel nodes inherit the capture stream's priority (highest: see s_prop / s_act)
            CUDA_TRY(cudaStreamCreateWithPriority(&cs, cudaStreamNonBlocking, prio_hi));
            Run C{h, cs};
            C.s_main = cs;
            C.multi = R.multi;
            cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) { cudaStreamDestroy(cs); return fail(BLURR_ERR_CUDA, "graph capture begin failed"); }
            run_step(C, batch, steps);
            cudaGraph_t g = nullptr;
            e = cudaStreamEndCapture(cs, &g);
            cudaStreamDestroy(cs);
            if (C.rc) { if (g) cudaGraphDestroy(g); return C.rc; }
            if (e != cudaSuccess || !g) return fail(BLURR_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
            cudaGraphExec_t ex = nullptr;
            e = cudaGraphInstantiate(&ex, g, 0);
            if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(BLURR_ERR_CUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(e)); }
            blurr_pi0::GraphEntry ge{g, ex, h->launches};
            it = h->graphs.emplace(key, ge).first;
            // the warm run already produced this call's result from the staged inputs; replay
            // anyway so the first call exercises the same path as every later one
            CUDA_TRY(launch_im2col(st, static_cast<const bf16*>(in->pixel_values), in->pixel_strides[0],
                                   in->pixel_strides[1], in->pixel_strides[2], in->pixel_strides[3], batch,
                                   h->patches, h->patch.K));
            stage_inputs_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(sa);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaGraphLaunch(it->second.exec, st));
        h->launches = it->second.launches + pre_launches;
    }
    CUDA_TRY(cudaMemcpyAsync(actions_out, h->d_out,
                             static_cast<size_t>(batch) * c.num_action_tokens * c.action_dim * 2,
                             cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int blurr_pi0_set_option(blurr_pi0_t* h, const char* name, int64_t value) {
    if (!h || !name) return fail(BLURR_ERR_INVALID, "set_option: null argument");
    const std::string n(name);
    if (n == "use_cuda_graph") h->use_graph = value != 0;
    else if (n == "use_streams") {
        h->use_streams = value != 0;
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "profile") { h->profile = value != 0; if (value == 2) h->prof.clear(); }
    else if (n == "chunked_splitk") {
        h->chunked_splitk = value != 0;
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "activation_clip_bits" || n == "activation_clip_mask") {
        // the reference's int8 fake-quant mode: value = the float32 bit pattern of the clip (0 = off) / the module mask
        if (n == "activation_clip_mask") h->act_clip_mask = static_cast<int>(value);
        else { const uint32_t bits = static_cast<uint32_t>(value); float f; memcpy(&f, &bits, 4); h->act_clip = f; }
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "lin_mode") {
        h->lin_mode = value != 0;
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "trace") {                   // in-graph per-kernel timeline (blurr_pi0_trace_report)
        if (value != 0 && h->trace_buf == nullptr &&
            cudaMalloc(&h->trace_buf, static_cast<size_t>(kTraceMax) * 4 * sizeof(unsigned long long)) != cudaSuccess)
            return fail(BLURR_ERR_CUDA, "trace buffer allocation failed");
        h->trace = value != 0;
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "debug_taps") h->debug = value != 0;
    else if (n == "stage_mask") {              // timing experiments only: run a subset of the stages
        h->stage_mask = static_cast<int>(value) & 7;
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "gemm_cluster_max") {        // activation-multicast cluster size cap (process-wide)
        gemm_set_cluster_max(static_cast<int>(value));
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "use_pdl") {                 // programmatic dependent launch (process-wide)
        pdl_set_enabled(value != 0);
        for (auto& kv : h->graphs) {
            cudaGraphExecDestroy(kv.second.exec);
            cudaGraphDestroy(kv.second.graph);
        }
        h->graphs.clear();
    }
    else if (n == "num_inference_steps") {
        if (value < 1) return fail(BLURR_ERR_INVALID, "num_inference_steps must be >= 1");
        h->cfg.num_inference_steps = static_cast<int>(value);
    } else return fail(BLURR_ERR_INVALID, "unknown option " + n);
    return 0;
}

static int g_op_bn_override = 0;      // tuning: token chunk width used by the stand-alone GEMM operator entry points

extern "C" int blurr_set_global_option(const char* name, int64_t value) {
    if (!name) return fail(BLURR_ERR_INVALID, "set_global_option: null name");
    const std::string n(name);
    if (n == "gemm_cluster_max") gemm_set_cluster_max(static_cast<int>(value));
    else if (n == "gemm_use_2cta") gemm_set_use_2cta(static_cast<int>(value));
    else if (n == "gemm_persistent") gemm_set_persistent(static_cast<int>(value));
    else if (n == "gemm_max_stages") gemm_set_max_stages(static_cast<int>(value));
    else if (n == "attn_tc") attn_set_tc(static_cast<int>(value));
    else if (n == "attn_tc_fewq") attn_set_tc_fewq(static_cast<int>(value));
    else if (n == "attn_fewq_stream") attn_set_fewq_stream(static_cast<int>(value));
    else if (n == "attn_prefill_stream") attn_set_prefill_stream(static_cast<int>(value));
    else if (n == "attn_siglip_stream") attn_set_siglip_stream(static_cast<int>(value));
    else if (n == "attn_cta_trace") { if (attn_set_cta_trace(reinterpret_cast<void*>(static_cast<intptr_t>(value)))) return fail(BLURR_ERR_CUDA, "attn_cta_trace: cudaMemcpyToSymbol failed"); }
    else if (n == "gemm_large_t_mode") gemm_set_large_t_mode(static_cast<int>(value));
    else if (n == "gemm_pair_band") gemm_set_pair_band(static_cast<int>(value));
    else if (n == "gemm_pair_small") gemm_set_pair_small(static_cast<int>(value));
    else if (n == "op_gemm_bn") g_op_bn_override = static_cast<int>(value);
    else if (n == "gemm_pair_policy") gemm_set_pair_policy(static_cast<int>(value));
    else if (n == "use_pdl") pdl_set_enabled(value != 0);
    else if (n == "gemm_cta_trace") { if (gemm_set_cta_trace(reinterpret_cast<void*>(static_cast<intptr_t>(value)))) return fail(BLURR_ERR_CUDA, "gemm_cta_trace: cudaMemcpyToSymbol failed"); }
    else return fail(BLURR_ERR_INVALID, "unknown global option " + n);
    return 0;
}

extern "C" int blurr_pi0_check(blurr_pi0_t* h, void* cuda_stream) {
    if (!h) return fail(BLURR_ERR_INVALID, "check: null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    if (int tf = gemm_take_timeout_flag())
        return fail(BLURR_ERR_CUDA, "a GEMM pipeline wait expired (role " + std::to_string(tf) + "): results are invalid");
    if (int tf = attn_take_timeout_flag())
        return fail(BLURR_ERR_CUDA, "an attention pipeline wait expired (stage " + std::to_string(tf) + "): results are invalid");
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag != 0) {
        CUDA_TRY(cudaMemset(h->d_err, 0, sizeof(int)));
        return fail(BLURR_ERR_INPUT, flag == 1 ? "more image tokens in input_ids than image features"
                                               : "token id outside the embedding table");
    }
    return 0;
}

extern "C" int blurr_pi0_debug_tap(blurr_pi0_t* h, const char* name, void* dst, size_t dst_bytes, size_t* bytes_out) {
    if (!h || !name) return fail(BLURR_ERR_INVALID, "debug_tap: null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const std::string n(name);
    const auto& c = h->cfg;
    if (n == "k_cache" || n == "v_cache") {
        // full-capacity layout [layers][max_batch][slots][head_dim]
        const size_t bytes = static_cast<size_t>(c.joint_layers) * h->max_batch * h->n_total * c.head_dim * 2;
        if (bytes_out) *bytes_out = bytes;
        if (!dst) return 0;
        if (dst_bytes < bytes) return fail(BLURR_ERR_INVALID, "debug_tap: destination too small");
        CUDA_TRY(cudaMemcpy(dst, n == "k_cache" ? h->kcache : h->vcache, bytes, cudaMemcpyDeviceToDevice));
        return 0;
    }
    auto it = h->taps.find(n);
    if (it == h->taps.end()) return fail(BLURR_ERR_STATE, "debug_tap: no such tap (enable option debug_taps and run a step): " + n);
    if (bytes_out) *bytes_out = it->second.bytes;
    if (!dst) return 0;
    if (dst_bytes < it->second.bytes) return fail(BLURR_ERR_INVALID, "debug_tap: destination too small");
    CUDA_TRY(cudaMemcpy(dst, it->second.ptr, it->second.bytes, cudaMemcpyDeviceToDevice));
    return 0;
}

extern "C" int64_t blurr_pi0_last_launch_count(const blurr_pi0_t* h) { return h ? h->launches : 0; }
extern "C" int blurr_pi0_profile_report(blurr_pi0_t* h, char* buf, size_t buf_bytes) {
    if (!h || !buf || buf_bytes == 0) return fail(BLURR_ERR_INVALID, "profile_report: bad arguments");
    std::vector<std::pair<double, std::string>> rows;
    double total = 0.0;
    for (auto& kv : h->prof) { rows.push_back({kv.second.ms, kv.first}); total += kv.second.ms; }
    std::sort(rows.begin(), rows.end(), [](const std::pair<double, std::string>& a, const std::pair<double, std::string>& b) { return a.first > b.first; });
    std::string out;
    char line[256];
    snprintf(line, sizeof line, "total %.3f ms over %zu labels (eager launches, CUDA events, warm L2)\n", total, rows.size());
    out += line;
    for (auto& r : rows) {
        const auto& e = h->prof[r.second];
        snprintf(line, sizeof line, "%9.1f us %5.1f%% %5dx avg %7.2f us  %s\n", e.ms * 1e3, 100.0 * e.ms / (total > 0 ? total : 1),
                 e.count, e.ms * 1e3 / e.count, r.second.c_str());
        out += line;
    }
    snprintf(buf, buf_bytes, "%s", out.c_str());
    return 0;
}

extern "C" int blurr_pi0_trace_report(blurr_pi0_t* h, char* buf, size_t buf_bytes) {
    if (!h || !buf || buf_bytes == 0) return fail(BLURR_ERR_INVALID, "trace_report: bad arguments");
    if (!h->trace || h->trace_buf == nullptr) return fail(BLURR_ERR_STATE, "trace_report: option trace is off");
    CUDA_TRY(cudaDeviceSynchronize());
    const size_t n = h->trace_labels.size();
    std::vector<unsigned long long> host(n * 4);
    if (n) CUDA_TRY(cudaMemcpy(host.data(), h->trace_buf, n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (size_t i = 0; i < n; ++i) if (host[i * 4] < t0) t0 = host[i * 4];
    std::string out = "# idx stream start_us waited_us end_us label   (globaltimer, relative to the first kernel start)\n";
    char line[256];
    for (size_t i = 0; i < n; ++i) {
        const unsigned long long s = host[i * 4], w = host[i * 4 + 1], e = ~host[i * 4 + 2];
        if (s == ~0ull) continue;     // never ran (stage masked off)
        snprintf(line, sizeof line, "%zu %d %.3f %.3f %.3f %s\n", i, h->trace_streams[i], (s - t0) * 1e-3,
                 w == ~0ull ? -1.0 : (w - t0) * 1e-3, (e - t0) * 1e-3, h->trace_labels[i].c_str());
        out += line;
    }
    if (out.size() + 1 > buf_bytes) return fail(BLURR_ERR_INVALID, "trace_report: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

extern "C" int64_t blurr_pi0_weight_bytes(const blurr_pi0_t* h) { return h ? static_cast<int64_t>(h->weight_bytes) : 0; }

// ---------------------------------------------------------------------------
// single-operator entry points
// ---------------------------------------------------------------------------
extern "C" int blurr_op_pack_weight(void* cuda_stream, const void* W, int N, int K, int ldw, void* packed) {
    if (N % 128 || K % 64) return fail(BLURR_ERR_INVALID, "pack_weight: N must be a multiple of 128, K of 64");
    const size_t total = static_cast<size_t>(N) * K;
    repack_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        static_cast<const bf16*>(W), N, K, ldw, static_cast<bf16*>(packed), 0, MAP_OFFSET, 0, K / 64);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int blurr_op_pair_raster(int N, int K, int T, int* band, int* n_pairs, int32_t* order, int capacity) {
    const int tiles = gemm_pair_raster(N, K, T, band, n_pairs, order, capacity);
    if (tiles < 0) return fail(BLURR_ERR_INVALID, "blurr_op_pair_raster: N must be a positive multiple of 128, K and T positive");
    return tiles;
}

extern "C" int blurr_op_gemm_async(void* cuda_stream, const void* W, int N, int K, int ldw, const void* X, int T,
                                   int ldx, int epi, int splitk, const void* bias, void* out, int ldo,
                                   float* partial) {
    GemmCall c{};
    c.W = static_cast<const bf16*>(W); c.Nw = N; c.K = K; c.ldw = ldw > 0 ? ldw : K; c.w_packed = ldw <= 0;
    c.X = static_cast<const bf16*>(X); c.T = T; c.ldx = ldx; c.epi = epi; c.splitk = splitk;
    c.bias = static_cast<const bf16*>(bias); c.out = static_cast<bf16*>(out); c.ldo = ldo; c.partial = partial;
    c.bn_override = g_op_bn_override;
    std::string err;
    const int s = gemm_launch(static_cast<cudaStream_t>(cuda_stream), c, &err);
    if (s < 0) return fail(BLURR_ERR_INVALID, err);
    return s;
}

extern "C" int blurr_op_gemm(void* cuda_stream, const void* W, int N, int K, int ldw, const void* X, int T, int ldx,
                             int epi, int splitk, const void* bias, void* out, int ldo, float* partial) {
    GemmCall c{};
    c.W = static_cast<const bf16*>(W); c.Nw = N; c.K = K; c.ldw = ldw > 0 ? ldw : K; c.w_packed = ldw <= 0;
    c.X = static_cast<const bf16*>(X); c.T = T; c.ldx = ldx; c.epi = epi; c.splitk = splitk;
    c.bias = static_cast<const bf16*>(bias); c.out = static_cast<bf16*>(out); c.ldo = ldo; c.partial = partial;
    c.bn_override = 0;
    std::string err;
    const int s = gemm_launch(static_cast<cudaStream_t>(cuda_stream), c, &err);
    if (s < 0) return fail(BLURR_ERR_INVALID, err);
    CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
    if (int tf = gemm_take_timeout_flag())
        return fail(BLURR_ERR_CUDA, "GEMM pipeline wait expired (role " + std::to_string(tf) + ")");
    return s;
}

extern "C" int blurr_op_siglip_attention(void* cuda_stream, const void* qkv, int ld_qkv, int batch, int seq, int heads,
                                         int hidden, void* out, int ld_out) {
    CUDA_TRY(launch_siglip_attention(static_cast<cudaStream_t>(cuda_stream), static_cast<const bf16*>(qkv), ld_qkv,
                                     batch, seq, heads, hidden, static_cast<bf16*>(out), ld_out));
    return 0;
}

extern "C" int blurr_op_joint_attention(void* cuda_stream, int few_query, const void* q, int q_per_sample,
                                        int q_row_offset, const void* k_cache, const void* v_cache, int n_slots,
                                        int n_keys, const void* mask, int64_t mask_bstride, int64_t mask_rstride,
                                        int batch, int n_heads, void* out) {
    JointAttnArgs a{};
    a.q = static_cast<const bf16*>(q); a.q_per_sample = q_per_sample; a.q_row_offset = q_row_offset;
    a.k_cache = static_cast<const bf16*>(k_cache); a.v_cache = static_cast<const bf16*>(v_cache);
    a.n_slots = n_slots; a.n_keys = n_keys; a.mask = static_cast<const bf16*>(mask);
    a.mask_bstride = mask_bstride; a.mask_rstride = mask_rstride; a.batch = batch; a.n_heads = n_heads;
    a.out = static_cast<bf16*>(out);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    CUDA_TRY(few_query ? launch_joint_attention_fewq(st, a) : launch_joint_attention_prefill(st, a));
    return 0;
}
