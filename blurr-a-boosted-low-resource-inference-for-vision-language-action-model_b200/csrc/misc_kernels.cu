// Stand-alone launches of the small bandwidth-bound kernels at the edges of the Pi-0 control step
// (bodies in bodies.cuh): patch extraction for the SigLIP patch-embed GEMM, text/image embedding
// merge, the K<=8 encoders, the action decoder + Euler update, and the final clamp.
#include "bodies.cuh"
#include "launch.cuh"

namespace blurr {

__global__ void __launch_bounds__(128) im2col_kernel(const bf16* px, long long sb, long long sc, long long sh,
                                                     long long sw, bf16* patches, int ldp) {
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    im2col_body(px, sb, sc, sh, sw, patches, ldp, blockIdx.x, blockIdx.y);
}

cudaError_t launch_im2col(cudaStream_t stream, const bf16* pixels, int64_t sb, int64_t sc, int64_t sh,
                          int64_t sw, int batch, bf16* patches, int ldp) {
    return launch_kernel(im2col_kernel, dim3(256, batch), dim3(128), 0, stream, pixels, sb, sc, sh, sw, patches,
                         ldp);
}

__global__ void __launch_bounds__(256) embed_merge_kernel(const int64_t* ids, int seq, const bf16* table,
                                                          long long vocab, const bf16* img, int n_img, int hidden,
                                                          long long image_token, long long pad_token, float inv_div,
                                                          float normalizer, bf16* out, int* err_flag) {
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    embed_merge_body(ids, seq, table, vocab, img, n_img, hidden, image_token, pad_token, inv_div, normalizer, out,
                     err_flag, blockIdx.x, blockIdx.y);
}

cudaError_t launch_embed_merge(cudaStream_t stream, const int64_t* input_ids, int batch, int seq,
                               const bf16* embed_table, int64_t vocab, const bf16* img_feat, int n_img,
                               int hidden, int64_t image_token, int64_t pad_token, float inv_div,
                               float normalizer, bf16* out, int* err_flag) {
    return launch_kernel(embed_merge_kernel, dim3(seq, batch), dim3(256), 0, stream, input_ids, seq, embed_table,
                         vocab, img_feat, n_img, hidden, image_token, pad_token, inv_div, normalizer, out,
                         err_flag);
}

__global__ void __launch_bounds__(256) small_k_linear_kernel(const bf16* x, int T, int K, const bf16* W,
                                                             const bf16* bias, int N, float scale, bf16* y, int ldy,
                                                             int col_off, const bf16* time_row, int time_cols) {
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    small_k_linear_body(x, T, K, W, bias, N, scale, y, ldy, col_off, time_row, time_cols, blockIdx.x, blockIdx.y);
}

cudaError_t launch_small_k_linear(cudaStream_t stream, const bf16* x, int T, int K, const bf16* W,
                                  const bf16* b, int N, float scale, bf16* y, int ldy, int col_off,
                                  const bf16* time_row, int time_cols) {
    const int cols = N > time_cols ? N : time_cols;
    return launch_kernel(small_k_linear_kernel, dim3((cols + 255) / 256, T), dim3(256), 0, stream, x, T, K, W, b, N,
                         scale, y, ldy, col_off, time_row, time_cols);
}

__global__ void __launch_bounds__(256) action_tail_kernel(const bf16* xn, int T, int hidden, const bf16* W,
                                                          const bf16* bias, int action_dim, float dt, bf16* action,
                                                          bf16* vel_tap) {
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    action_tail_body(xn, T, hidden, W, bias, action_dim, dt, action, vel_tap, blockIdx.x);
}

cudaError_t launch_action_tail(cudaStream_t stream, const bf16* xn, int T, int hidden, const bf16* W,
                               const bf16* b, int action_dim, float dt, bf16* action, bf16* velocity_tap) {
    const int warps = T * action_dim;
    return launch_kernel(action_tail_kernel, dim3((warps * 32 + 255) / 256), dim3(256), 0, stream, xn, T, hidden, W,
                         b, action_dim, dt, action, velocity_tap);
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const bf16* src, int src_rows_per_sample, int row0, int rows, int width,
                                                        int lds, bf16* dst, int ldd) {
    pdl_wait();
    pdl_trigger();
    const int r = blockIdx.x, b = blockIdx.y;
    const uint4* s = reinterpret_cast<const uint4*>(src + (static_cast<size_t>(b) * src_rows_per_sample + row0 + r) * lds);
    uint4* d = reinterpret_cast<uint4*>(dst + (static_cast<size_t>(b) * rows + r) * ldd);
    for (int i = threadIdx.x; i < width / 8; i += blockDim.x) d[i] = s[i];
}

cudaError_t launch_copy_rows(cudaStream_t stream, const bf16* src, int batch, int src_rows_per_sample, int row0, int rows, int width,
                             int lds, bf16* dst, int ldd) {
    if ((width & 7) || (lds & 7) || (ldd & 7)) return cudaErrorInvalidValue;
    return launch_kernel(copy_rows_kernel, dim3(rows, batch), dim3(256), 0, stream, src, src_rows_per_sample, row0, rows, width, lds,
                         dst, ldd);
}

// Last kernel of a control step.  `flags` are the device-side sticky error words of the step (GEMM / attention pipeline
// time-outs, input validation): if any is set the step's results are not trustworthy, so the actions leave as NaN —
// a robot loop that never calls blurr_pi0_check() still cannot act on garbage (the reference would have raised).
__global__ void __launch_bounds__(256) clamp_copy_kernel(const bf16* src, bf16* dst, int n, int do_clamp, float clip,
                                                         const int* flag0, const int* flag1, const int* flag2) {
    pdl_wait();
    pdl_trigger();     // only now: a successor that is resident earlier just holds SM resources while it waits (measured)
    const bool bad = (flag0 != nullptr && *flag0 != 0) || (flag1 != nullptr && *flag1 != 0) || (flag2 != nullptr && *flag2 != 0);
    if (bad) {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i < n) dst[i] = __float2bfloat16(__int_as_float(0x7fc00000));
        return;
    }
    clamp_copy_body(src, dst, n, do_clamp, clip, blockIdx.x);
}

cudaError_t launch_clamp_copy(cudaStream_t stream, const bf16* src, bf16* dst, int n, int do_clamp,
                              float clip, const int* flag0, const int* flag1, const int* flag2) {
    return launch_kernel(clamp_copy_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, src, dst, n, do_clamp, clip, flag0,
                         flag1, flag2);
}

}  // namespace blurr
