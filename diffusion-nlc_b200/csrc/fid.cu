// FID statistics on the device: the non-GEMM pieces of pytorch_fid's InceptionV3 pipeline, which the
// reference runs through `pytorch_fid` after a PNG -> disk -> reload round trip (src/experiments.py:210-226,
// image_sample.py:566,703, result_evaluater.py:24-27).  The convolutions themselves are tensor-core GEMMs
// (nlc_conv_tc over the patch matrices written by `im2col` here, BatchNorm folded into weights and bias, ReLU in the conv
// epilogue); this file holds
//   * fid_preprocess: sample in [-1,1] -> [0,1] -> 8-bit quantisation of the PNG round trip -> bilinear 299 x 299
//     (align_corners = False, PyTorch's index arithmetic) -> 2x - 1 -> NHWC operand
//   * im2col_nhwc: NHWC -> zero-padded [M_pad, K_pad] patch matrix (tap-major, channel-minor K order of pack_conv_weight)
//   * pool2d: 3x3 max / average (count_include_pad = False) pooling, stride 1 pad 1 or stride 2 valid
//   * global_avgpool: [B, HW, C] -> fp32 [B, C]
//   * cov_accumulate: sum and outer-product accumulation of the 2048-d features in fp64 (mu, Sigma of the FID)
// All of it is memory-bound glue: 16-byte vectors where the channel count allows.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

template <typename T>
__device__ __forceinline__ T to_op(float v, int fmt);
template <>
__device__ __forceinline__ float to_op<float>(float v, int fmt) { return op_f32(v, fmt); }
template <>
__device__ __forceinline__ __nv_bfloat16 to_op<__nv_bfloat16>(float v, int fmt) {
    if (fmt) {
        const __half h = __float2half_rn(v);
        return *reinterpret_cast<const __nv_bfloat16*>(&h);
    }
    return __float2bfloat16_rn(v);
}
template <typename T>
__device__ __forceinline__ float from_op(T v, int fmt);
template <>
__device__ __forceinline__ float from_op<float>(float v, int) { return v; }
template <>
__device__ __forceinline__ float from_op<__nv_bfloat16>(__nv_bfloat16 v, int fmt) {
    if (fmt) return __half2float(*reinterpret_cast<const __half*>(&v));
    return __bfloat162float(v);
}

// ---------------------------------------------------------------- preprocess
// PyTorch upsample_bilinear2d (align_corners = False): src = max(scale * (dst + 0.5) - 0.5, 0), i1 = min(i0 + 1, in - 1).
template <typename T>
__global__ void fid_preprocess_kernel(const float* __restrict__ x, int B, int H, int W, int from_pm1, int quantize,
                                      int resize, int normalize, int Ro, T* __restrict__ y, int ld_y, int fmt) {
    const int total = B * Ro * Ro;
    const float sh = static_cast<float>(H) / static_cast<float>(Ro), sw = static_cast<float>(W) / static_cast<float>(Ro);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / (Ro * Ro), r = i - n * Ro * Ro;
        const int ho = r / Ro, wo = r - ho * Ro;
        int h0 = ho, h1 = ho, w0 = wo, w1 = wo;
        float lh = 0.f, lw = 0.f;
        if (resize) {
            const float fh = fmaxf(sh * (static_cast<float>(ho) + 0.5f) - 0.5f, 0.f);
            const float fw = fmaxf(sw * (static_cast<float>(wo) + 0.5f) - 0.5f, 0.f);
            h0 = static_cast<int>(fh), w0 = static_cast<int>(fw);
            h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
            lh = fh - static_cast<float>(h0), lw = fw - static_cast<float>(w0);
        }
        for (int c = 0; c < 3; ++c) {
            const float* xc = x + (static_cast<size_t>(n) * 3 + c) * H * W;
            auto px = [&](int h, int w) {
                float v = __ldg(xc + static_cast<size_t>(h) * W + w);
                if (from_pm1) v = fminf(fmaxf(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), 0.f), 1.f);  // add(1).div(2).clamp(0,1)
                if (quantize)  // save_image: mul(255).add(0.5).clamp(0,255) -> uint8 (truncation); ToTensor: / 255
                    v = __fdiv_rn(floorf(fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f), 0.f), 255.f)), 255.0f);
                return v;
            };
            float v;
            if (resize) {
                const float top = __fadd_rn(__fmul_rn(1.0f - lw, px(h0, w0)), __fmul_rn(lw, px(h0, w1)));
                const float bot = __fadd_rn(__fmul_rn(1.0f - lw, px(h1, w0)), __fmul_rn(lw, px(h1, w1)));
                v = __fadd_rn(__fmul_rn(1.0f - lh, top), __fmul_rn(lh, bot));
            } else {
                v = px(ho, wo);
            }
            if (normalize) v = __fadd_rn(__fmul_rn(2.0f, v), -1.0f);
            y[static_cast<size_t>(i) * ld_y + c] = to_op<T>(v, fmt);
        }
    }
}

// ---------------------------------------------------------------- im2col
// One thread per (row m, tap, 8-channel group) when C % 8 == 0 (16-byte copies), else one per (m, k) element.
template <typename T, int VEC>
__global__ void im2col_nhwc_kernel(const T* __restrict__ x, int ld_x, int B, int H, int W, int C, int KH, int KW, int SH,
                                   int SW, int PH, int PW, int Ho, int Wo, T* __restrict__ out, int K_pad, long long M,
                                   long long M_pad) {
    const int cg = C / VEC;              // channel groups per tap
    const int kg = KH * KW * cg;         // groups of real data per row
    const int kg_pad = K_pad / VEC;      // groups per padded row
    const long long total = M_pad * kg_pad;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long m = i / kg_pad;
        const int g = static_cast<int>(i - m * kg_pad);
        T v[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = T(0.f);
        if (m < M && g < kg) {
            const int tap = g / cg, c = (g - tap * cg) * VEC;
            const int kh = tap / KW, kw = tap - kh * KW;
            const int n = static_cast<int>(m / (Ho * Wo));
            const int r = static_cast<int>(m - static_cast<long long>(n) * Ho * Wo);
            const int ho = r / Wo, wo = r - ho * Wo;
            const int h = ho * SH + kh - PH, w = wo * SW + kw - PW;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const T* src = x + ((static_cast<size_t>(n) * H + h) * W + w) * ld_x + c;
                if (VEC * sizeof(T) == 16) {
                    *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(src));
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[e] = src[e];
                }
            }
        }
        T* dst = out + m * K_pad + static_cast<size_t>(g) * VEC;
        if (VEC * sizeof(T) == 16) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = v[e];
        }
    }
}

// ---------------------------------------------------------------- pooling (3x3)
// mode 0: max;  1: average over the taps inside the image (count_include_pad = False)
template <typename T>
__global__ void pool2d_kernel(const T* __restrict__ x, int ld_x, int B, int H, int W, int C, int stride, int pad, int Ho,
                              int Wo, int mode, T* __restrict__ y, int ld_y, int fmt) {
    const long long total = static_cast<long long>(B) * Ho * Wo * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long long pix = i / C;
        const int wo = static_cast<int>(pix % Wo);
        const int ho = static_cast<int>((pix / Wo) % Ho);
        const int n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
        float acc = mode == 0 ? -INFINITY : 0.f;
        int cnt = 0;
        for (int kh = 0; kh < 3; ++kh) {
            const int h = ho * stride + kh - pad;
            if (h < 0 || h >= H) continue;
            for (int kw = 0; kw < 3; ++kw) {
                const int w = wo * stride + kw - pad;
                if (w < 0 || w >= W) continue;
                const float v = from_op<T>(x[((static_cast<size_t>(n) * H + h) * W + w) * ld_x + c], fmt);
                acc = mode == 0 ? fmaxf(acc, v) : acc + v;
                ++cnt;
            }
        }
        if (mode == 1) acc = acc / static_cast<float>(cnt);
        y[static_cast<size_t>(pix) * ld_y + c] = to_op<T>(acc, fmt);
    }
}

// [B, HW, ld] -> fp32 [B, C] mean over HW; grid (ceil(C / 128), B), 128 threads
template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ x, int ld_x, int HW, int C, float* __restrict__ y, int fmt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (c >= C) return;
    float acc = 0.f;
    const T* xp = x + static_cast<size_t>(n) * HW * ld_x + c;
    for (int p = 0; p < HW; ++p) acc += from_op<T>(xp[static_cast<size_t>(p) * ld_x], fmt);
    y[static_cast<size_t>(n) * C + c] = acc / static_cast<float>(HW);
}

// ---------------------------------------------------------------- mu / Sigma accumulation (fp64)
// sum[d] += sum_b f[b,d];  outer[i,j] += sum_b f[b,i] f[b,j].  grid (D/16, D/16), block 16 x 16; block (0, *) also does the sums.
__global__ void cov_accumulate_kernel(const float* __restrict__ f, int Bn, int D, double* __restrict__ sum,
                                      double* __restrict__ outer) {
    __shared__ float fi[16][17], fj[16][17];
    const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int ty = threadIdx.y, tx = threadIdx.x;
    double acc = 0.0, s = 0.0;
    for (int b0 = 0; b0 < Bn; b0 += 16) {
        const int b = b0 + ty;
        fi[ty][tx] = (b < Bn && i0 + tx < D) ? f[static_cast<size_t>(b) * D + i0 + tx] : 0.f;
        fj[ty][tx] = (b < Bn && j0 + tx < D) ? f[static_cast<size_t>(b) * D + j0 + tx] : 0.f;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            acc += static_cast<double>(fi[k][ty]) * static_cast<double>(fj[k][tx]);
            if (blockIdx.y == 0 && ty == 0) s += static_cast<double>(fj[k][tx]);
        }
        __syncthreads();
    }
    if (i0 + ty < D && j0 + tx < D) outer[static_cast<size_t>(i0 + ty) * D + j0 + tx] += acc;
    if (blockIdx.y == 0 && ty == 0 && j0 + tx < D) sum[j0 + tx] += s;
}

static int grid_for(long long total, int threads, int sm_count) {
    long long g = (total + threads - 1) / threads;
    const long long cap = 32LL * sm_count;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_fid_preprocess(nlc_ctx* ctx, const float* x_nchw, int B, int H, int W, int from_pm1, int quantize,
                                  int resize, int normalize, int R_out, void* y_op, int ld_y, int op_dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_nchw && y_op, "nlc_fid_preprocess: null argument");
    NLC_REQUIRE(dtype_valid(op_dtype) && B >= 1 && H >= 1 && W >= 1 && ld_y >= 3, "nlc_fid_preprocess: bad arguments");
    NLC_REQUIRE(resize || (R_out == H && R_out == W), "nlc_fid_preprocess: without resizing the output is the input size");
    const int fmt = dtype_fmt(op_dtype);
    const int grid = grid_for(static_cast<long long>(B) * R_out * R_out, 256, ctx->sm_count);
    if (dtype_is16(op_dtype))
        fid_preprocess_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x_nchw, B, H, W, from_pm1, quantize, resize, normalize,
                                                                      R_out, static_cast<__nv_bfloat16*>(y_op), ld_y, fmt);
    else
        fid_preprocess_kernel<float><<<grid, 256, 0, stream>>>(x_nchw, B, H, W, from_pm1, quantize, resize, normalize, R_out,
                                                              static_cast<float*>(y_op), ld_y, fmt);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_im2col_nhwc(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C, int KH,
                               int KW, int SH, int SW, int PH, int PW, void* out, int K_pad, long long M_pad, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_op && out, "nlc_im2col_nhwc: null argument");
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_im2col_nhwc: bad op_dtype");
    const int Ho = (H + 2 * PH - KH) / SH + 1, Wo = (W + 2 * PW - KW) / SW + 1;
    const long long M = static_cast<long long>(B) * Ho * Wo;
    const int esz = dtype_is16(op_dtype) ? 2 : 4;
    const int vec = 16 / esz;
    NLC_REQUIRE(Ho >= 1 && Wo >= 1 && M_pad >= M && K_pad >= KH * KW * C && K_pad % vec == 0,
                "nlc_im2col_nhwc: output %dx%d, K_pad %d < %d or M_pad too small", Ho, Wo, K_pad, KH * KW * C);
    const bool vec_ok = C % vec == 0 && ld_x % vec == 0 && (reinterpret_cast<uintptr_t>(x_op) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (dtype_is16(op_dtype)) {
        const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_op);
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
        if (vec_ok)
            im2col_nhwc_kernel<__nv_bfloat16, 8><<<grid_for(M_pad * (K_pad / 8), 256, ctx->sm_count), 256, 0, stream>>>(
                x, ld_x, B, H, W, C, KH, KW, SH, SW, PH, PW, Ho, Wo, o, K_pad, M, M_pad);
        else
            im2col_nhwc_kernel<__nv_bfloat16, 1><<<grid_for(M_pad * K_pad, 256, ctx->sm_count), 256, 0, stream>>>(
                x, ld_x, B, H, W, C, KH, KW, SH, SW, PH, PW, Ho, Wo, o, K_pad, M, M_pad);
    } else {
        const float* x = static_cast<const float*>(x_op);
        float* o = static_cast<float*>(out);
        if (vec_ok)
            im2col_nhwc_kernel<float, 4><<<grid_for(M_pad * (K_pad / 4), 256, ctx->sm_count), 256, 0, stream>>>(
                x, ld_x, B, H, W, C, KH, KW, SH, SW, PH, PW, Ho, Wo, o, K_pad, M, M_pad);
        else
            im2col_nhwc_kernel<float, 1><<<grid_for(M_pad * K_pad, 256, ctx->sm_count), 256, 0, stream>>>(
                x, ld_x, B, H, W, C, KH, KW, SH, SW, PH, PW, Ho, Wo, o, K_pad, M, M_pad);
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_pool2d(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C, int stride,
                          int pad, int mode, void* y_op, int ld_y, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_op && y_op, "nlc_pool2d: null argument");
    NLC_REQUIRE(dtype_valid(op_dtype) && (mode == 0 || mode == 1) && (stride == 1 || stride == 2) && (pad == 0 || pad == 1),
                "nlc_pool2d: 3x3 max (0) / average (1) pooling with stride 1 | 2 and padding 0 | 1");
    const int Ho = (H + 2 * pad - 3) / stride + 1, Wo = (W + 2 * pad - 3) / stride + 1;
    NLC_REQUIRE(Ho >= 1 && Wo >= 1, "nlc_pool2d: input %dx%d too small", H, W);
    const int fmt = dtype_fmt(op_dtype);
    const int grid = grid_for(static_cast<long long>(B) * Ho * Wo * C, 256, ctx->sm_count);
    if (dtype_is16(op_dtype))
        pool2d_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x_op), ld_x, B, H, W, C, stride,
                                                              pad, Ho, Wo, mode, static_cast<__nv_bfloat16*>(y_op), ld_y, fmt);
    else
        pool2d_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x_op), ld_x, B, H, W, C, stride, pad, Ho, Wo,
                                                      mode, static_cast<float*>(y_op), ld_y, fmt);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_global_avgpool(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int HW, int C, float* y,
                                  void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_op && y && dtype_valid(op_dtype) && B >= 1 && HW >= 1 && C >= 1, "nlc_global_avgpool: bad arguments");
    const dim3 grid((C + 127) / 128, B);
    if (dtype_is16(op_dtype))
        global_avgpool_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(x_op), ld_x, HW, C, y,
                                                                      dtype_fmt(op_dtype));
    else
        global_avgpool_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float*>(x_op), ld_x, HW, C, y,
                                                              dtype_fmt(op_dtype));
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_cov_accumulate(nlc_ctx* ctx, const float* feats, int B, int D, double* sum, double* outer, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && feats && sum && outer && B >= 1 && D >= 1, "nlc_cov_accumulate: bad arguments");
    const dim3 grid((D + 15) / 16, (D + 15) / 16);
    cov_accumulate_kernel<<<grid, dim3(16, 16), 0, stream>>>(feats, B, D, sum, outer);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
