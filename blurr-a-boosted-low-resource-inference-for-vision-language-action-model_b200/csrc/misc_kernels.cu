// Small bandwidth-bound kernels at the edges of the Pi-0 control step: patch extraction for
// the SigLIP patch-embed GEMM, text/image embedding merge, the K<=8 encoders, the action
// decoder + Euler update, and the final clamp.
#include "common.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace blurr {

// ---------------------------------------------------------------------------
// im2col for Conv2d(k=s=14) (siglip.py:42-48,69): one CTA per patch
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) im2col_kernel(const bf16* __restrict__ px, long long sb, long long sc,
                                                     long long sh, long long sw, bf16* __restrict__ patches,
                                                     int ldp) {
    pdl_wait();
    pdl_trigger();
    const int p = blockIdx.x;            // patch index within the image, row-major 16x16
    const int b = blockIdx.y;
    const int ph = p >> 4, pw = p & 15;
    bf16* dst = patches + (static_cast<size_t>(b) * 256 + p) * ldp;
    for (int idx = threadIdx.x; idx < 588; idx += blockDim.x) {
        const int c = idx / 196, rem = idx - c * 196;
        const int kh = rem / 14, kw = rem - kh * 14;
        dst[idx] = px[b * sb + c * sc + static_cast<long long>(ph * 14 + kh) * sh +
                      static_cast<long long>(pw * 14 + kw) * sw];
    }
}

cudaError_t launch_im2col(cudaStream_t stream, const bf16* pixels, int64_t sb, int64_t sc, int64_t sh,
                          int64_t sw, int batch, bf16* patches, int ldp) {
    return launch_kernel(im2col_kernel, dim3(256, batch), dim3(128), 0, stream, pixels, sb, sc, sh, sw, patches,
                         ldp);
}

// ---------------------------------------------------------------------------
// embedding merge (pizero.py:440-471) + `embeds *= sqrt(hidden)` (joint_model.py:358-365)
// one CTA per (sample, position).  The k-th image-token position of a sample receives image
// feature row k (`final_embedding[i, image_indices] = scaled[i, :num_image_tokens]`).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_merge_kernel(const int64_t* __restrict__ ids, int seq,
                                                          const bf16* __restrict__ table, long long vocab,
                                                          const bf16* __restrict__ img, int n_img, int hidden,
                                                          long long image_token, long long pad_token,
                                                          float inv_div, float normalizer,
                                                          bf16* __restrict__ out, int* err_flag) {
    __shared__ int s_rank;
    pdl_wait();
    pdl_trigger();
    const int pos = blockIdx.x, b = blockIdx.y;
    const int64_t* row = ids + static_cast<size_t>(b) * seq;
    const long long id = row[pos];
    bf16* dst = out + (static_cast<size_t>(b) * seq + pos) * hidden;
    if (id == image_token) {
        if (threadIdx.x == 0) s_rank = 0;
        __syncthreads();
        int cnt = 0;
        for (int i = threadIdx.x; i < pos; i += blockDim.x) cnt += (row[i] == image_token) ? 1 : 0;
        cnt = static_cast<int>(warp_sum(static_cast<float>(cnt)));
        if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_rank, cnt);
        __syncthreads();
        const int rank = s_rank;
        if (rank >= n_img) {             // the reference would raise a shape error here
            if (threadIdx.x == 0) *err_flag = 1;
            return;
        }
        const bf16* src = img + (static_cast<size_t>(b) * n_img + rank) * hidden;
        for (int n = threadIdx.x; n < hidden; n += blockDim.x) {
            float x = bf16_round(bf2f(src[n]) * inv_div);     // image_features / sqrt(hidden)
            x = bf16_round(x * normalizer);                   // embeds *= bf16(sqrt(hidden))
            dst[n] = f2bf(x);
        }
    } else if (id != pad_token) {
        if (id < 0 || id >= vocab) {
            if (threadIdx.x == 0) *err_flag = 2;
            return;
        }
        const bf16* src = table + static_cast<size_t>(id) * hidden;
        for (int n = threadIdx.x; n < hidden; n += blockDim.x)
            dst[n] = f2bf(bf16_round(bf2f(src[n]) * normalizer));
    } else {
        // torch.full(..., pad_token_id) rows, then *= normalizer
        const float x = bf16_round(bf16_round(static_cast<float>(pad_token)) * normalizer);
        for (int n = threadIdx.x; n < hidden; n += blockDim.x) dst[n] = f2bf(x);
    }
}

cudaError_t launch_embed_merge(cudaStream_t stream, const int64_t* input_ids, int batch, int seq,
                               const bf16* embed_table, int64_t vocab, const bf16* img_feat, int n_img,
                               int hidden, int64_t image_token, int64_t pad_token, float inv_div,
                               float normalizer, bf16* out, int* err_flag) {
    return launch_kernel(embed_merge_kernel, dim3(seq, batch), dim3(256), 0, stream, input_ids, seq, embed_table,
                         vocab, img_feat, n_img, hidden, image_token, pad_token, inv_div, normalizer, out,
                         err_flag);
}

// ---------------------------------------------------------------------------
// K <= 8 linear (proprio_encoder pizero.py:493, action_encoder.linear_1 vla/modules.py:45)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) small_k_linear_kernel(const bf16* __restrict__ x, int T, int K,
                                                             const bf16* __restrict__ W,
                                                             const bf16* __restrict__ bias, int N, float scale,
                                                             bf16* __restrict__ y, int ldy, int col_off,
                                                             const bf16* __restrict__ time_row, int time_cols) {
    pdl_wait();
    pdl_trigger();
    const int t = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += bf2f(x[t * K + k]) * bf2f(W[n * K + k]);
        acc += bf2f(bias[n]);
        float v = bf16_round(acc);
        if (scale != 1.0f) v = bf16_round(v * scale);
        y[static_cast<size_t>(t) * ldy + col_off + n] = f2bf(v);
    }
    if (time_row != nullptr && n < time_cols) y[static_cast<size_t>(t) * ldy + n] = time_row[n];
}

cudaError_t launch_small_k_linear(cudaStream_t stream, const bf16* x, int T, int K, const bf16* W,
                                  const bf16* b, int N, float scale, bf16* y, int ldy, int col_off,
                                  const bf16* time_row, int time_cols) {
    const int cols = N > time_cols ? N : time_cols;
    return launch_kernel(small_k_linear_kernel, dim3((cols + 255) / 256, T), dim3(256), 0, stream, x, T, K, W, b, N,
                         scale, y, ldy, col_off, time_row, time_cols);
}

// ---------------------------------------------------------------------------
// action_decoder (Linear hidden -> action_dim) + Euler step (pizero.py:536-537)
// one warp per (token, action component)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) action_tail_kernel(const bf16* __restrict__ xn, int T, int hidden,
                                                          const bf16* __restrict__ W,
                                                          const bf16* __restrict__ bias, int action_dim,
                                                          float dt, bf16* __restrict__ action,
                                                          bf16* __restrict__ vel_tap) {
    pdl_wait();
    pdl_trigger();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= T * action_dim) return;
    const int t = gw / action_dim, d = gw - t * action_dim;
    float acc = 0.f;
    for (int k = lane; k < hidden; k += 32)
        acc += bf2f(xn[static_cast<size_t>(t) * hidden + k]) * bf2f(W[static_cast<size_t>(d) * hidden + k]);
    acc = warp_sum(acc);
    if (lane == 0) {
        const float vel = bf16_round(acc + bf2f(bias[d]));
        if (vel_tap != nullptr) vel_tap[gw] = f2bf(vel);
        const float step = bf16_round(dt * vel);                       // delta_t * action_vel
        action[gw] = f2bf(bf16_round(bf2f(action[gw]) + step));        // action += ...
    }
}

cudaError_t launch_action_tail(cudaStream_t stream, const bf16* xn, int T, int hidden, const bf16* W,
                               const bf16* b, int action_dim, float dt, bf16* action, bf16* velocity_tap) {
    const int warps = T * action_dim;
    return launch_kernel(action_tail_kernel, dim3((warps * 32 + 255) / 256), dim3(256), 0, stream, xn, T, hidden, W,
                         b, action_dim, dt, action, velocity_tap);
}

__global__ void clamp_copy_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int n, int do_clamp,
                                  float clip) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = bf2f(src[i]);
    if (do_clamp) v = (v < -clip) ? -clip : ((v > clip) ? clip : v);   // NaN propagates like torch.clamp
    dst[i] = f2bf(v);
}

cudaError_t launch_clamp_copy(cudaStream_t stream, const bf16* src, bf16* dst, int n, int do_clamp,
                              float clip) {
    return launch_kernel(clamp_copy_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, src, dst, n, do_clamp, clip);
}

}  // namespace blurr
