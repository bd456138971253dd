// Attention kernels of the Pi-0 path.
//
//  * attn_mma_kernel<HD_PAD, BM, GEMMA>: "full-row" attention, BM query rows per CTA (16 at batch 1,
//    32/64 for batched episodes), for the many-query cases
//      - SigLIP MHA, 16 heads x 256 tokens x head_dim 72 (siglip.py:133-152), no mask;
//      - Gemma joint-attention prefill, 8 query heads sharing one KV head (MQA), 276 query rows
//        x 277 keys x head_dim 256 with tanh soft-clamp and additive block mask
//        (joint_model.py:273-288).
//    The tiles are tiny (<= 64 x 320 logits), so they run on warp-level mma.sync (bf16, fp32
//    accumulate) with ldmatrix-fed fragments; the whole logit row is kept in shared memory so
//    the softmax is the reference's exact two-pass fp32 softmax over bf16-rounded logits.
//  * few-query mode of the same kernel: 1 (proprio) or 4 (action) query rows per sample over the
//    KV cache (joint_model.py:164-170 "append_non_active"); the (head, query) pairs of a sample
//    form the tile rows, so the sample's single K/V head is read once for all 8 query heads.
//
// Rounding points: QK^T -> bf16; every scale / tanh / mask op
// -> bf16; softmax in fp32 -> bf16; PV -> bf16.  Divisions by Python scalars are done as the
// ATen CUDA kernels do them (multiplication by the fp32 reciprocal).
#include "bodies.cuh"
#include "launch.cuh"

#include <string>

namespace blurr {

template <int HD_PAD, int BM, int GEMMA>
__global__ void __launch_bounds__(kAttnThreads) attn_mma_kernel(const AttnMmaArgs a) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    trace_stamp(a.trace, 0);
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    trace_stamp(a.trace, 1);
    attn_mma_body<HD_PAD, BM, GEMMA>(a, smem_attn, blockIdx.x, blockIdx.y, blockIdx.z);
    trace_stamp(a.trace, 2);
}

AttnMmaArgs make_siglip_attn_args(const bf16* qkv, int ld_qkv, int seq, int n_heads, int hidden, bf16* out,
                                  int ld_out) {
    AttnMmaArgs a{};
    a.trace = nullptr;
    const int hd = hidden / n_heads;           // 72
    a.q = qkv; a.ldq = ld_qkv; a.q_col0 = 0; a.q_per_sample = seq;
    a.k = qkv; a.ldk = ld_qkv; a.k_col0 = hidden; a.kv_per_sample = seq;
    a.v = qkv; a.ldv = ld_qkv; a.v_col0 = 2 * hidden;
    a.out = out; a.ldo = ld_out; a.o_col0 = 0;
    a.hd = hd; a.head_stride_q = hd; a.head_stride_kv = hd; a.n_keys = seq;
    // Python: head_dim ** -0.5 evaluated in double, then used as an fp32 scalar operand
    a.scale = static_cast<float>(pow(static_cast<double>(hd), -0.5));
    a.mask = nullptr;
    return a;
}

AttnMmaArgs make_prefill_attn_args(const JointAttnArgs& j) {
    AttnMmaArgs a{};
    a.q = j.q; a.ldq = j.n_heads * 256; a.q_col0 = 0; a.q_per_sample = j.q_per_sample;
    a.k = j.k_cache; a.ldk = 256; a.k_col0 = 0; a.kv_per_sample = j.n_slots;
    a.v = j.v_cache; a.ldv = 256; a.v_col0 = 0;
    a.out = j.out; a.ldo = j.n_heads * 256; a.o_col0 = 0;
    a.hd = 256; a.head_stride_q = 256; a.head_stride_kv = 0; a.n_keys = j.n_keys;
    a.scale = 0.f;
    a.mask = j.mask; a.mask_bstride = j.mask_bstride; a.mask_rstride = j.mask_rstride;
    a.q_row_offset = j.q_row_offset;
    a.trace = j.trace;
    return a;
}

// Few queries per sample (proprio: 1, action: 4): the (head, query) pairs of a sample form the rows of
// the same tensor-core tile kernel — all 8 query heads share the sample's single K/V head (MQA), so
// K and V are read once per 16 pairs instead of once per head.
AttnMmaArgs make_fewq_attn_args(const JointAttnArgs& j) {
    AttnMmaArgs a = make_prefill_attn_args(j);
    a.head_stride_q = 256;
    a.mqa_nq = j.q_per_sample;
    a.mqa_heads = j.n_heads;
    return a;
}

static constexpr size_t kAttnSmemMax = 215 * 1024;

template <int HD_PAD, int BM, int GEMMA>
static cudaError_t launch_attn(cudaStream_t stream, const AttnMmaArgs& a, int rows, int heads, int batch) {
    const size_t smem = attn_smem_bytes<HD_PAD, BM>(a.n_keys);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_kernel<HD_PAD, BM, GEMMA>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kAttnSmemMax));
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (smem > kAttnSmemMax) return cudaErrorInvalidValue;
    dim3 grid((rows + BM - 1) / BM, heads, batch);
    return launch_kernel(attn_mma_kernel<HD_PAD, BM, GEMMA>, grid, dim3(kAttnThreads), smem, stream, a);
}

// rows per tile: 16 while that already gives every SM a couple of tiles, larger tiles beyond
int attn_tile_rows(int rows, int heads, int batch, int max_rows) {
    const long tiles16 = static_cast<long>((rows + 15) / 16) * heads * batch;
    if (tiles16 <= 2 * 148 || max_rows <= 16) return 16;
    const long tiles32 = static_cast<long>((rows + 31) / 32) * heads * batch;
    if (tiles32 <= 4 * 148 || max_rows <= 32) return 32;
    return 64;
}

bool attn_tc_siglip_applies(int batch, int seq, int n_heads, int hidden);
cudaError_t launch_siglip_attention_tc(cudaStream_t stream, const bf16* qkv, int ld_qkv, int batch, int seq, int n_heads,
                                       int hidden, bf16* out, int ld_out, unsigned long long* trace);

cudaError_t launch_siglip_attention(cudaStream_t stream, const bf16* qkv, int ld_qkv, int batch, int seq,
                                    int n_heads, int hidden, bf16* out, int ld_out, unsigned long long* trace) {
    const int hd = hidden / n_heads;
    if (hd > 80 || seq > kAttnMaxBlocks * kBK) return cudaErrorInvalidValue;
    if (attn_tc_siglip_applies(batch, seq, n_heads, hidden))
        return launch_siglip_attention_tc(stream, qkv, ld_qkv, batch, seq, n_heads, hidden, out, ld_out, trace);
    AttnMmaArgs a = make_siglip_attn_args(qkv, ld_qkv, seq, n_heads, hidden, out, ld_out);
    a.trace = trace;
    switch (attn_tile_rows(seq, n_heads, batch, 64)) {
        case 16: return launch_attn<80, 16, false>(stream, a, seq, n_heads, batch);
        case 32: return launch_attn<80, 32, false>(stream, a, seq, n_heads, batch);
        default: return launch_attn<80, 64, false>(stream, a, seq, n_heads, batch);
    }
}

cudaError_t launch_joint_attention_prefill_tc(cudaStream_t stream, const JointAttnArgs& j, std::string* err);

cudaError_t launch_joint_attention_prefill(cudaStream_t stream, const JointAttnArgs& j) {
    if (j.n_keys > kAttnMaxBlocks * kBK) return cudaErrorInvalidValue;
    if (attn_tc_applies(j)) {
        std::string err;
        return launch_joint_attention_prefill_tc(stream, j, &err);
    }
    AttnMmaArgs a = make_prefill_attn_args(j);
    if (attn_tile_rows(j.q_per_sample, j.n_heads, j.batch, 32) == 16)
        return launch_attn<256, 16, true>(stream, a, j.q_per_sample, j.n_heads, j.batch);
    return launch_attn<256, 32, true>(stream, a, j.q_per_sample, j.n_heads, j.batch);
}

cudaError_t launch_joint_attention_fewq(cudaStream_t stream, const JointAttnArgs& j) {
    if (j.n_keys > kAttnMaxBlocks * kBK) return cudaErrorInvalidValue;
    if (attn_tc_fewq_applies(j)) {
        std::string err;
        return launch_joint_attention_prefill_tc(stream, j, &err);
    }
    AttnMmaArgs a = make_fewq_attn_args(j);
    const int pairs = j.n_heads * j.q_per_sample;
    if (attn_tile_rows(pairs, 1, j.batch, 32) == 16) return launch_attn<256, 16, true>(stream, a, pairs, 1, j.batch);
    return launch_attn<256, 32, true>(stream, a, pairs, 1, j.batch);
}

// Llama-style multi-head attention over a token-major KV cache (llm_engine.cu): head_dim 128 (or 64), one K/V head per
// query head, causal by position.  Prefill (many query rows) and decode (one row per sequence) share the kernel.
cudaError_t launch_mha_decode(cudaStream_t stream, const MhaAttnArgs& a);     // llm_kernels.cu

cudaError_t launch_mha_attention(cudaStream_t stream, const MhaAttnArgs& m) {
    if (m.n_keys > kAttnMaxBlocks * kBK || m.n_kv_heads != m.n_heads) return cudaErrorInvalidValue;
    if (m.q_per_sample == 1) return launch_mha_decode(stream, m);       // decode: K/V streaming, no tensor cores
    AttnMmaArgs a{};
    const int width = m.n_heads * m.head_dim;
    a.q = m.q; a.ldq = width; a.q_col0 = 0; a.q_per_sample = m.q_per_sample;
    a.k = m.k_cache; a.ldk = width; a.k_col0 = 0; a.kv_per_sample = m.n_slots;
    a.v = m.v_cache; a.ldv = width; a.v_col0 = 0;
    a.out = m.out; a.ldo = width; a.o_col0 = 0;
    a.hd = m.head_dim; a.head_stride_q = m.head_dim; a.head_stride_kv = m.head_dim; a.n_keys = m.n_keys;
    a.scale = m.scale; a.mask = nullptr; a.q_row_offset = m.q_pos0; a.trace = m.trace;
    const int bm = attn_tile_rows(m.q_per_sample, m.n_heads, m.batch, 64);
    if (m.head_dim == 128) {
        if (bm == 16) return launch_attn<128, 16, 2>(stream, a, m.q_per_sample, m.n_heads, m.batch);
        if (bm == 32) return launch_attn<128, 32, 2>(stream, a, m.q_per_sample, m.n_heads, m.batch);
        return launch_attn<128, 64, 2>(stream, a, m.q_per_sample, m.n_heads, m.batch);
    }
    if (m.head_dim == 64) {
        if (bm == 16) return launch_attn<64, 16, 2>(stream, a, m.q_per_sample, m.n_heads, m.batch);
        return launch_attn<64, 32, 2>(stream, a, m.q_per_sample, m.n_heads, m.batch);
    }
    return cudaErrorInvalidValue;
}

}  // namespace blurr
