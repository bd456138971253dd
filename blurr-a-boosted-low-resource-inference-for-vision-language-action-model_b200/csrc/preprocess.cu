// Device-side observation preprocessing: the per-control-step host work the
// reference does between the simulator and the model, moved onto the GPU with bit-identical results.
//
//  * frame:   uint8 HWC camera frame -> cv2.resize(INTER_LANCZOS4) to the model resolution
//             (env_adapter/simpler.py:59-64) -> VLAProcessor rescale / normalise in fp32
//             (model/vla/processing.py:27-58,112-117) -> `.to(bfloat16)` (agent/eval.py:187), written
//             as the contiguous [3][h][w] pixel_values tensor the engine's patch loader reads.
//             The resize is OpenCV's 8-tap fixed-point Lanczos (modules/imgproc/src/resize.cpp:
//             interpolateLanczos4, INTER_RESIZE_COEF_BITS = 11, HResizeLanczos4 / VResizeLanczos4 with
//             FixedPtCast<int, uchar, 22>, replicated borders); the tables are built on the host at
//             create time with the same float / double operation order, the two integer passes run in
//             one kernel (one CTA per output row, the 8 source rows staged in shared memory).
//  * proprio: normalize_bound / normalize_gaussian in float64 (env_adapter/base.py:8-18,33-40),
//             float32 (simpler.py:93-95), then bfloat16 (eval.py:194).
// Both kernels are pure bandwidth/latency work (a 640x480 frame is 0.9 MB): no tensor cores.
#include "blurr_pi0.h"
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

namespace blurr {
int record_error(int code, const std::string& msg);      // engine.cu: sets blurr_last_error()
inline int fail(int code, const std::string& msg) { return record_error(code, msg); }
}
using blurr::bf16;

struct blurr_preproc {
    int device = 0;
    int src_h = 0, src_w = 0, dst_h = 0, dst_w = 0;
    int32_t *x_ofs = nullptr, *y_ofs = nullptr;      // device
    int16_t *x_alpha = nullptr, *y_alpha = nullptr;  // device, [dst][8]
    std::vector<int32_t> hx_ofs, hy_ofs;
    std::vector<int16_t> hx_alpha, hy_alpha;
};

namespace {

constexpr int kTaps = 8;
constexpr int kCoefBits = 11;

// resize.cpp interpolateLanczos4(float x, float* coeffs), operation for operation
void lanczos4_coeffs(float x, float* coeffs) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double pi = 3.1415926535897932384626433832795;
    float sum = 0;
    const float x3 = x + 3;
    const double y0 = -static_cast<double>(x3) * pi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
    for (int i = 0; i < kTaps; ++i) {
        const float y0_ = x3 - static_cast<float>(i);
        if (std::fabs(y0_) >= 1e-6f) {
            const double y = -static_cast<double>(y0_) * pi * 0.25;
            coeffs[i] = static_cast<float>((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            coeffs[i] = 1e30f;
        }
        sum += coeffs[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < kTaps; ++i) coeffs[i] *= sum;
}

void build_tables(int src, int dst, std::vector<int32_t>* ofs, std::vector<int16_t>* alpha) {
    const double inv_scale = static_cast<double>(dst) / src;
    const double scale = 1. / inv_scale;
    ofs->resize(dst);
    alpha->resize(static_cast<size_t>(dst) * kTaps);
    for (int d = 0; d < dst; ++d) {
        float f = static_cast<float>((d + 0.5) * scale - 0.5);
        const int s = static_cast<int>(std::floor(f));
        f -= static_cast<float>(s);
        float c[kTaps];
        lanczos4_coeffs(f, c);
        for (int k = 0; k < kTaps; ++k) {
            long v = std::lrintf(c[k] * static_cast<float>(1 << kCoefBits));     // saturate_cast<short>(float)
            v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
            (*alpha)[static_cast<size_t>(d) * kTaps + k] = static_cast<int16_t>(v);
        }
        (*ofs)[d] = s;
    }
}

struct FrameArgs {
    const uint8_t* frame; long long row_stride, frame_stride;
    int src_h, src_w, dst_h, dst_w;
    const int32_t *x_ofs, *y_ofs;
    const int16_t *x_alpha, *y_alpha;
    bf16* pixel_values;       // [batch][3][dst_h][dst_w]
    uint8_t* resized;         // [batch][dst_h][dst_w][3] or nullptr
};

__global__ void __launch_bounds__(256) resize_lanczos4_normalize_kernel(const FrameArgs a) {
    extern __shared__ __align__(16) uint8_t rows[];          // [8][src_w * 3]
    const int dy = blockIdx.x, b = blockIdx.y;
    const int row_bytes = a.src_w * 3;
    const uint8_t* src = a.frame + static_cast<size_t>(b) * a.frame_stride;
    const int sy0 = a.y_ofs[dy] - (kTaps / 2 - 1);
    for (int k = 0; k < kTaps; ++k) {
        int sy = sy0 + k;
        sy = sy < 0 ? 0 : (sy >= a.src_h ? a.src_h - 1 : sy);            // replicated border
        const uint8_t* g = src + static_cast<size_t>(sy) * a.row_stride;
        uint8_t* s = rows + k * row_bytes;
        if (((reinterpret_cast<uintptr_t>(g) | static_cast<uintptr_t>(row_bytes) | reinterpret_cast<uintptr_t>(s)) & 3) == 0) {
            for (int i = threadIdx.x; i < row_bytes / 4; i += blockDim.x)
                reinterpret_cast<uint32_t*>(s)[i] = __ldg(reinterpret_cast<const uint32_t*>(g) + i);
        } else {
            for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) s[i] = __ldg(g + i);
        }
    }
    __syncthreads();
    int beta[kTaps];
#pragma unroll
    for (int k = 0; k < kTaps; ++k) beta[k] = a.y_alpha[dy * kTaps + k];
    for (int dx = threadIdx.x; dx < a.dst_w; dx += blockDim.x) {
        int xi[kTaps], al[kTaps];
        const int sx0 = a.x_ofs[dx] - (kTaps / 2 - 1);
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            int sx = sx0 + j;
            sx = sx < 0 ? 0 : (sx >= a.src_w ? a.src_w - 1 : sx);
            xi[j] = sx * 3;
            al[j] = a.x_alpha[dx * kTaps + j];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v = 0;                                   // cv2 accumulates both passes in int (WT = int)
#pragma unroll
            for (int k = 0; k < kTaps; ++k) {
                const uint8_t* r = rows + k * row_bytes + c;
                int h = 0;
#pragma unroll
                for (int j = 0; j < kTaps; ++j) h += static_cast<int>(r[xi[j]]) * al[j];
                v += h * beta[k];
            }
            v = (v + (1 << (2 * kCoefBits - 1))) >> (2 * kCoefBits);      // FixedPtCast<int, uchar, 22>
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            if (a.resized != nullptr)
                a.resized[((static_cast<size_t>(b) * a.dst_h + dy) * a.dst_w + dx) * 3 + c] = static_cast<uint8_t>(v);
            // processing.py: image * (1/255.0) -> (x - 0.5) / 0.5 in fp32 (one rounding per op), then bf16
            float f = __fmul_rn(static_cast<float>(v), static_cast<float>(1 / 255.0));
            f = __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);
            a.pixel_values[((static_cast<size_t>(b) * 3 + c) * a.dst_h + dy) * a.dst_w + dx] = __float2bfloat16_rn(f);
        }
    }
}

__global__ void normalize_proprio_kernel(const double* raw, const double* lo, const double* hi, int kind, int n,
                                         int dim, bf16* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * dim) return;
    const int d = i % dim;
    const double x = raw[i];
    double y;
    if (kind == 0) {
        // 2 * (data - min) / (max - min + eps) - 1, then clip to [-1, 1]  (base.py:17-18)
        const double num = __dmul_rn(2.0, __dsub_rn(x, lo[d]));
        const double den = __dadd_rn(__dsub_rn(hi[d], lo[d]), 1e-8);
        y = __dsub_rn(__ddiv_rn(num, den), 1.0);
        y = y < -1.0 ? -1.0 : (y > 1.0 ? 1.0 : y);
    } else {
        // (data - mean) / (std + eps)  (base.py:40)
        y = __ddiv_rn(__dsub_rn(x, lo[d]), __dadd_rn(hi[d], 1e-8));
    }
    out[i] = __float2bfloat16_rn(static_cast<float>(y));      // float64 -> float32 -> bf16, as the reference
}

}  // namespace

extern "C" int blurr_preproc_create(int device, int src_h, int src_w, int dst_h, int dst_w, blurr_preproc_t** out) {
    if (!out) return blurr::fail(BLURR_ERR_INVALID, "preproc_create: null out");
    *out = nullptr;
    if (src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1)
        return blurr::fail(BLURR_ERR_INVALID, "preproc_create: sizes must be positive");
    if (static_cast<long long>(src_w) * 3 * kTaps > 200 * 1024)
        return blurr::fail(BLURR_ERR_INVALID, "preproc_create: source rows too wide for the shared-memory stage");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
        return blurr::fail(BLURR_ERR_CUDA, "no CUDA device: the preprocessing path has no CPU fallback");
    if (device < 0 || device >= n_dev) return blurr::fail(BLURR_ERR_INVALID, "preproc_create: bad device index");
    if (cudaSetDevice(device) != cudaSuccess) return blurr::fail(BLURR_ERR_CUDA, "cudaSetDevice failed");
    blurr_preproc* p = new blurr_preproc();
    p->device = device; p->src_h = src_h; p->src_w = src_w; p->dst_h = dst_h; p->dst_w = dst_w;
    build_tables(src_w, dst_w, &p->hx_ofs, &p->hx_alpha);
    build_tables(src_h, dst_h, &p->hy_ofs, &p->hy_alpha);
    bool ok = cudaMalloc(&p->x_ofs, dst_w * sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&p->y_ofs, dst_h * sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&p->x_alpha, static_cast<size_t>(dst_w) * kTaps * sizeof(int16_t)) == cudaSuccess &&
              cudaMalloc(&p->y_alpha, static_cast<size_t>(dst_h) * kTaps * sizeof(int16_t)) == cudaSuccess;
    ok = ok && cudaMemcpy(p->x_ofs, p->hx_ofs.data(), dst_w * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(p->y_ofs, p->hy_ofs.data(), dst_h * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(p->x_alpha, p->hx_alpha.data(), p->hx_alpha.size() * sizeof(int16_t), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(p->y_alpha, p->hy_alpha.data(), p->hy_alpha.size() * sizeof(int16_t), cudaMemcpyHostToDevice) == cudaSuccess;
    const int smem = src_w * 3 * kTaps;
    if (ok && smem > 48 * 1024)
        ok = cudaFuncSetAttribute(resize_lanczos4_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    if (!ok) {
        blurr_preproc_destroy(p);
        return blurr::fail(BLURR_ERR_CUDA, "preproc_create: device allocation failed");
    }
    *out = p;
    return 0;
}

extern "C" void blurr_preproc_destroy(blurr_preproc_t* p) {
    if (!p) return;
    cudaFree(p->x_ofs); cudaFree(p->y_ofs); cudaFree(p->x_alpha); cudaFree(p->y_alpha);
    delete p;
}

extern "C" int blurr_preproc_tables(const blurr_preproc_t* p, int32_t* x_ofs, int16_t* x_alpha, int32_t* y_ofs,
                                    int16_t* y_alpha) {
    if (!p || !x_ofs || !x_alpha || !y_ofs || !y_alpha) return blurr::fail(BLURR_ERR_INVALID, "preproc_tables: null argument");
    std::copy(p->hx_ofs.begin(), p->hx_ofs.end(), x_ofs);
    std::copy(p->hx_alpha.begin(), p->hx_alpha.end(), x_alpha);
    std::copy(p->hy_ofs.begin(), p->hy_ofs.end(), y_ofs);
    std::copy(p->hy_alpha.begin(), p->hy_alpha.end(), y_alpha);
    return 0;
}

extern "C" int blurr_preproc_build_tables(int src, int dst, int32_t* ofs, int16_t* alpha) {
    if (src < 1 || dst < 1 || !ofs || !alpha) return blurr::fail(BLURR_ERR_INVALID, "preproc_build_tables: bad argument");
    std::vector<int32_t> o;
    std::vector<int16_t> a;
    build_tables(src, dst, &o, &a);
    std::copy(o.begin(), o.end(), ofs);
    std::copy(a.begin(), a.end(), alpha);
    return 0;
}

extern "C" int blurr_preproc_frame(blurr_preproc_t* p, void* cuda_stream, const void* frame_u8, int64_t row_stride_bytes,
                                   int batch, int64_t frame_stride_bytes, void* pixel_values_bf16, void* resized_u8) {
    if (!p || !frame_u8 || !pixel_values_bf16) return blurr::fail(BLURR_ERR_INVALID, "preproc_frame: null argument");
    if (batch < 1 || row_stride_bytes < static_cast<int64_t>(p->src_w) * 3)
        return blurr::fail(BLURR_ERR_INVALID, "preproc_frame: bad batch or row stride");
    if (cudaSetDevice(p->device) != cudaSuccess) return blurr::fail(BLURR_ERR_CUDA, "cudaSetDevice failed");
    FrameArgs a{};
    a.frame = static_cast<const uint8_t*>(frame_u8); a.row_stride = row_stride_bytes; a.frame_stride = frame_stride_bytes;
    a.src_h = p->src_h; a.src_w = p->src_w; a.dst_h = p->dst_h; a.dst_w = p->dst_w;
    a.x_ofs = p->x_ofs; a.y_ofs = p->y_ofs; a.x_alpha = p->x_alpha; a.y_alpha = p->y_alpha;
    a.pixel_values = static_cast<bf16*>(pixel_values_bf16); a.resized = static_cast<uint8_t*>(resized_u8);
    const size_t smem = static_cast<size_t>(p->src_w) * 3 * kTaps;
    resize_lanczos4_normalize_kernel<<<dim3(p->dst_h, batch), 256, smem, static_cast<cudaStream_t>(cuda_stream)>>>(a);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return blurr::fail(BLURR_ERR_CUDA, std::string("preproc_frame launch failed: ") + cudaGetErrorString(e));
    return 0;
}

extern "C" int blurr_op_normalize_proprio(void* cuda_stream, const double* raw, const double* lo, const double* hi,
                                          int kind, int n, int dim, void* out_bf16) {
    if (!raw || !lo || !hi || !out_bf16 || n < 1 || dim < 1 || (kind != 0 && kind != 1))
        return blurr::fail(BLURR_ERR_INVALID, "normalize_proprio: bad argument");
    const int total = n * dim;
    normalize_proprio_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        raw, lo, hi, kind, n, dim, static_cast<bf16*>(out_bf16));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return blurr::fail(BLURR_ERR_CUDA, std::string("normalize_proprio launch failed: ") + cudaGetErrorString(e));
    return 0;
}
