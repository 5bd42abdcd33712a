// Direct fp32 (CUDA-core) 3x3 convolutions for the two layers whose channel counts cannot fill a tensor-core
// tile: conv_in (3 -> C, reads the sampler's NCHW image, folds the per-sample 1/sqrt(sigma^2+1) input scaling
// of ExperimentDiffusion.convert_coordinate, src/experiments.py:273-282) and conv_out (C -> 3|6, writes the
// NCHW eps the sampler consumes).  Together < 0.2 % of a forward's FLOPs; they also convert between the
// reference's NCHW fp32 boundary layout and the NHWC operand layout used internally.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

// ---------------------------------------------------------------- conv_in: NCHW fp32 -> NHWC
// thread = (pixel, 8 consecutive output channels); weights in smem as [tap*Cin + ci][Cout].
template <bool TF32>
__global__ void __launch_bounds__(256)
    conv_in_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, int B, int Cin, int H, int W,
                   const float* __restrict__ wt, const float* __restrict__ bias, int Cout, float* __restrict__ yf,
                   int ld_yf, void* __restrict__ yo, int ld_yo, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ float sw[];  // [9*Cin][Cout]
    const int K = 9 * Cin;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i / K, r = i - co * K;  // torch layout: [co][ci][kh][kw]
        const int ci = r / 9, tap = r - ci * 9;
        sw[(tap * Cin + ci) * Cout + co] = wt[i];
    }
    __syncthreads();
    const int CG = Cout >> 3;
    const int ppb = blockDim.x / CG;
    const int lp = threadIdx.x / CG, cg = threadIdx.x - lp * CG;
    const long long npix = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * ppb + lp; pix < npix;
         pix += static_cast<long long>(gridDim.x) * ppb) {
        if (lp >= ppb) break;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
        float acc[8] = {};
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xc = x + (static_cast<size_t>(n) * Cin + ci) * H * W;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int ww = w + kw - 1;
                    float v = 0.f;
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(xc + static_cast<size_t>(hh) * W + ww);
                    const float* wr = sw + ((kh * 3 + kw) * Cin + ci) * Cout + cg * 8;
                    const float4 w0 = *reinterpret_cast<const float4*>(wr);
                    const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
                    acc[0] = fmaf(v, w0.x, acc[0]), acc[1] = fmaf(v, w0.y, acc[1]);
                    acc[2] = fmaf(v, w0.z, acc[2]), acc[3] = fmaf(v, w0.w, acc[3]);
                    acc[4] = fmaf(v, w1.x, acc[4]), acc[5] = fmaf(v, w1.y, acc[5]);
                    acc[6] = fmaf(v, w1.z, acc[6]), acc[7] = fmaf(v, w1.w, acc[7]);
                }
            }
        }
        const float sc = in_scale ? in_scale[n] : 1.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], sc, bias ? bias[cg * 8 + k] : 0.f);
        if (yf) {
            float* o = yf + static_cast<size_t>(pix) * ld_yf + cg * 8;
            *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        if (yo) {
            if (TF32) {
                float* o = static_cast<float*>(yo) + static_cast<size_t>(pix) * ld_yo + cg * 8;
                *reinterpret_cast<float4*>(o) =
                    make_float4(op_f32(acc[0], rnd), op_f32(acc[1], rnd), op_f32(acc[2], rnd), op_f32(acc[3], rnd));
                *reinterpret_cast<float4*>(o + 4) =
                    make_float4(op_f32(acc[4], rnd), op_f32(acc[5], rnd), op_f32(acc[6], rnd), op_f32(acc[7], rnd));
            } else {
                __nv_bfloat16* o = static_cast<__nv_bfloat16*>(yo) + static_cast<size_t>(pix) * ld_yo + cg * 8;
                *reinterpret_cast<uint4*>(o) = make_uint4(pack_op16x2(acc[0], acc[1], rnd), pack_op16x2(acc[2], acc[3], rnd),
                                                          pack_op16x2(acc[4], acc[5], rnd), pack_op16x2(acc[6], acc[7], rnd));
            }
        }
    }
}

// ---------------------------------------------------------------- im2col of the network input
// The input convolution (3 -> C) as a tensor-core GEMM: each pixel's 3x3xCin receptive field (27 values for RGB),
// scaled by the per-sample input scale, is laid out as one K-major operand row [tap*Cin + ci], zero-padded to one
// 128-byte swizzle row (64 bf16 / 32 tf32); nlc_conv_tc then runs it as a 1x1 convolution with K = 64 | 32.
// Replaces the CUDA-core conv_in kernel above on the hot path: that one is bound by fp32 FMA issue, this one by the
// HBM write of the output.
template <bool TF32>
__global__ void __launch_bounds__(256)
    im2col_in_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, int B, int Cin, int H, int W,
                     void* __restrict__ patches, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    constexpr int KP = TF32 ? 32 : 64;
    const long long npix = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < npix;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
        const float sc = in_scale ? in_scale[n] : 1.0f;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
        const float* xn = x + static_cast<size_t>(n) * Cin * H * W;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
            if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)
                    if (ci < Cin) v[tap * Cin + ci] = __ldg(xn + (static_cast<size_t>(ci) * H + hh) * W + ww) * sc;
            }
        }
        if (TF32) {
            float4* o = reinterpret_cast<float4*>(static_cast<float*>(patches) + static_cast<size_t>(pix) * KP);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                o[i] = make_float4(op_f32(v[4 * i], rnd), op_f32(v[4 * i + 1], rnd), op_f32(v[4 * i + 2], rnd),
                                   op_f32(v[4 * i + 3], rnd));
        } else {
            uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(patches) + static_cast<size_t>(pix) * KP);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = make_uint4(pack_op16x2(v[8 * i], v[8 * i + 1], rnd), pack_op16x2(v[8 * i + 2], v[8 * i + 3], rnd),
                                  pack_op16x2(v[8 * i + 4], v[8 * i + 5], rnd), pack_op16x2(v[8 * i + 6], v[8 * i + 7], rnd));
#pragma unroll
            for (int i = 4; i < 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// ---------------------------------------------------------------- conv_out: NHWC operand -> NCHW fp32
// one warp per group of 4 consecutive pixels of a row; lanes split the input channels; weights in smem as
// [tap][co][ci].  COUT is a template parameter so the accumulators stay in registers.
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a), v[1] = __high2float(a), v[2] = __low2float(b), v[3] = __high2float(b);
}

constexpr int kOutPix = 4;
template <typename T, int COUT>
__global__ void __launch_bounds__(256)
    conv_out_kernel(const T* __restrict__ x, int ld_x, int B, int Cin, int H, int W, const float* __restrict__ wt,
                    const float* __restrict__ bias, float* __restrict__ out) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ float sw[];  // [9][COUT][Cin]
    for (int i = threadIdx.x; i < COUT * Cin * 9; i += blockDim.x) {
        const int co = i / (Cin * 9), r = i - co * Cin * 9;
        const int ci = r / 9, tap = r - ci * 9;
        sw[(tap * COUT + co) * Cin + ci] = wt[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wpr = W / kOutPix;  // pixel groups per row
    const long long ngroups = static_cast<long long>(B) * H * wpr;
    const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long g = warp0; g < ngroups; g += nwarps) {
        const int w0 = static_cast<int>(g % wpr) * kOutPix;
        const int h = static_cast<int>((g / wpr) % H);
        const int n = static_cast<int>(g / (static_cast<long long>(wpr) * H));
        float acc[kOutPix][COUT] = {};
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = h + kh - 1;
            if (hh < 0 || hh >= H) continue;
            const T* xrow = x + (static_cast<size_t>(n) * H + hh) * W * ld_x;
            for (int c = lane * 4; c < Cin; c += 128) {
                // input pixels w0-1 .. w0+4 feed the 4 outputs x 3 horizontal taps
                float xin[kOutPix + 2][4];
#pragma unroll
                for (int j = 0; j < kOutPix + 2; ++j) {
                    const int ww = w0 + j - 1;
                    if (ww >= 0 && ww < W) {
                        load4<T>(xrow + static_cast<size_t>(ww) * ld_x + c, xin[j]);
                    } else {
                        xin[j][0] = xin[j][1] = xin[j][2] = xin[j][3] = 0.f;
                    }
                }
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                    for (int co = 0; co < COUT; ++co) {
                        const float4 wv = *reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * COUT + co) * Cin + c);
#pragma unroll
                        for (int p = 0; p < kOutPix; ++p) {
                            acc[p][co] = fmaf(xin[p + kw][0], wv.x, acc[p][co]);
                            acc[p][co] = fmaf(xin[p + kw][1], wv.y, acc[p][co]);
                            acc[p][co] = fmaf(xin[p + kw][2], wv.z, acc[p][co]);
                            acc[p][co] = fmaf(xin[p + kw][3], wv.w, acc[p][co]);
                        }
                    }
                }
            }
        }
        float mine = 0.f;
#pragma unroll
        for (int co = 0; co < COUT; ++co)
#pragma unroll
            for (int p = 0; p < kOutPix; ++p) {
                const float s = warp_sum(acc[p][co]);
                if (lane == co * kOutPix + p) mine = s;
            }
        if (lane < COUT * kOutPix) {
            const int co = lane / kOutPix, p = lane - co * kOutPix;
            out[((static_cast<size_t>(n) * COUT + co) * H + h) * W + w0 + p] = mine + (bias ? bias[co] : 0.f);
        }
    }
}

template <typename T>
static int launch_conv_out(nlc_ctx* ctx, const T* x, int ld_x, int B, int Cin, int H, int W, const float* wt,
                           const float* bias, int Cout, float* out, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(9) * Cout * Cin * sizeof(float);
    const long long ngroups = static_cast<long long>(B) * H * (W / kOutPix);
    long long blocks = (ngroups + 7) / 8;
    const long long cap = static_cast<long long>(ctx->sm_count) * 8;
    if (blocks > cap) blocks = cap;
#define NLC_CO(N)                                                                                                \
    case N:                                                                                                      \
        NLC_CHECK_CUDA(cudaFuncSetAttribute(conv_out_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                            static_cast<int>(smem)));                                            \
        launch_pdl((conv_out_kernel<T, N>), dim3(static_cast<unsigned>(blocks)), dim3(256), smem, stream, x, ld_x, B, Cin, H, W, wt,    \
                                                                                    bias, out);                  \
        break;
    switch (Cout) {
        NLC_CO(1) NLC_CO(2) NLC_CO(3) NLC_CO(4) NLC_CO(6)
        default:
            return set_error(NLC_ENOTSUP, "nlc_conv_out_nchw: Cout=%d unsupported (1,2,3,4,6)", Cout);
    }
#undef NLC_CO
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_conv_in_nchw(nlc_ctx* ctx, const float* x_nchw, const float* in_scale, int B, int Cin, int H, int W,
                                const float* weight, const float* bias, int Cout, float* out_f32, int ld_out_f32,
                                void* out_op, int ld_out_op, int op_dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_nchw && weight && (out_f32 || out_op), "nlc_conv_in_nchw: null argument");
    NLC_REQUIRE(Cin >= 1 && Cin <= 16 && Cout % 8 == 0 && Cout <= 2048 && 256 % (Cout / 8) == 0,
                "nlc_conv_in_nchw: Cin=%d Cout=%d unsupported", Cin, Cout);
    NLC_REQUIRE(!out_f32 || ld_out_f32 % 4 == 0, "nlc_conv_in_nchw: ld_out_f32 %% 4");
    NLC_REQUIRE(!out_op || ld_out_op % 8 == 0, "nlc_conv_in_nchw: ld_out_op %% 8");
    const size_t smem = static_cast<size_t>(9) * Cin * Cout * sizeof(float);
    NLC_REQUIRE(smem <= 200 * 1024, "nlc_conv_in_nchw: weights do not fit shared memory");
    const int ppb = 256 / (Cout / 8);
    const long long npix = static_cast<long long>(B) * H * W;
    long long blocks = (npix + ppb - 1) / ppb;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_conv_in_nchw: bad op_dtype");
    const int rnd = dtype_fmt(op_dtype);
    if (!dtype_is16(op_dtype)) {
        NLC_CHECK_CUDA(cudaFuncSetAttribute(conv_in_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
        launch_pdl((conv_in_kernel<true>), dim3(static_cast<unsigned>(blocks)), dim3(256), smem, stream, 
            x_nchw, in_scale, B, Cin, H, W, weight, bias, Cout, out_f32, ld_out_f32, out_op, ld_out_op, rnd);
    } else {
        NLC_CHECK_CUDA(cudaFuncSetAttribute(conv_in_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
        launch_pdl((conv_in_kernel<false>), dim3(static_cast<unsigned>(blocks)), dim3(256), smem, stream, 
            x_nchw, in_scale, B, Cin, H, W, weight, bias, Cout, out_f32, ld_out_f32, out_op, ld_out_op, rnd);
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_conv_out_nchw(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int Cin, int H, int W,
                                 const float* weight, const float* bias, int Cout, float* out_nchw, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_op && weight && out_nchw, "nlc_conv_out_nchw: null argument");
    NLC_REQUIRE(Cin % 128 == 0 && W % kOutPix == 0 && ld_x % 4 == 0, "nlc_conv_out_nchw: Cin=%d W=%d unsupported", Cin,
                W);
    NLC_REQUIRE(static_cast<size_t>(9) * Cout * Cin * 4 <= 200 * 1024, "nlc_conv_out_nchw: weights exceed shared memory");
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_conv_out_nchw: bad op_dtype");
    if (!dtype_is16(op_dtype))
        return launch_conv_out<float>(ctx, static_cast<const float*>(x_op), ld_x, B, Cin, H, W, weight, bias, Cout,
                                      out_nchw, stream);
    if (op_dtype == NLC_F16)
        return launch_conv_out<__half>(ctx, static_cast<const __half*>(x_op), ld_x, B, Cin, H, W, weight, bias, Cout,
                                       out_nchw, stream);
    return launch_conv_out<__nv_bfloat16>(ctx, static_cast<const __nv_bfloat16*>(x_op), ld_x, B, Cin, H, W, weight,
                                          bias, Cout, out_nchw, stream);
}

extern "C" int nlc_im2col_in(nlc_ctx* ctx, const float* x_nchw, const float* in_scale, int B, int Cin, int H, int W,
                             void* patches_op, int op_dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_nchw && patches_op, "nlc_im2col_in: null argument");
    NLC_REQUIRE(Cin >= 1 && Cin <= 3, "nlc_im2col_in: Cin=%d unsupported (1..3: 9*Cin must fit one 32-element K row)", Cin);
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_im2col_in: bad op_dtype");
    const int rnd = dtype_fmt(op_dtype);
    NLC_REQUIRE((reinterpret_cast<uintptr_t>(patches_op) & 15) == 0, "nlc_im2col_in: patches must be 16-byte aligned");
    const long long npix = static_cast<long long>(B) * H * W;
    long long blocks = (npix + 255) / 256;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    if (!dtype_is16(op_dtype))
        launch_pdl((im2col_in_kernel<true>), dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, x_nchw, in_scale, B, Cin, H, W, patches_op, rnd);
    else
        launch_pdl((im2col_in_kernel<false>), dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, x_nchw, in_scale, B, Cin, H, W, patches_op, rnd);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
