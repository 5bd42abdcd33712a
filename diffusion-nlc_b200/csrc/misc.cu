// Small supporting kernels of the UNet executors: resampling/cast passes that produce tensor-core operands,
// the timestep-embedding sinusoid, and the small fp32 dense layers (temb MLP, per-block temb projections,
// sigma-model head).
#include "common.h"
#include "ptx.cuh"

namespace nlc {

// ---------------------------------------------------------------- resample: copy / nearest x2 / avgpool 2x2
// x: NHWC fp32 [B,H,W,C] (pitch ld_x); outputs at the resampled resolution, fp32 and/or operand dtype.
template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 ld4<__half>(const __half* p) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

template <int MODE, bool TF32, typename TIN>
__global__ void __launch_bounds__(256) resample_kernel(const TIN* __restrict__ x, int ld_x, int H, int W, int C,
                                                        float* __restrict__ yf, int ld_yf, void* __restrict__ yo,
                                                        int ld_yo, long long total4, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int Ho = MODE == 1 ? 2 * H : (MODE == 2 ? H / 2 : H);
    const int Wo = MODE == 1 ? 2 * W : (MODE == 2 ? W / 2 : W);
    const int C4 = C >> 2;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C4) << 2;
        long long pix = i / C4;
        const int wo = static_cast<int>(pix % Wo);
        pix /= Wo;
        const int ho = static_cast<int>(pix % Ho);
        const int n = static_cast<int>(pix / Ho);
        float4 v;
        if (MODE == 2) {
            const TIN* p = x + ((static_cast<size_t>(n) * H + 2 * ho) * W + 2 * wo) * ld_x + c;
            const float4 a = ld4<TIN>(p);
            const float4 b = ld4<TIN>(p + ld_x);
            const float4 d = ld4<TIN>(p + static_cast<size_t>(W) * ld_x);
            const float4 e = ld4<TIN>(p + static_cast<size_t>(W) * ld_x + ld_x);
            // same association as torch avg_pool2d: sum of the window, then divide
            v.x = ((a.x + b.x) + (d.x + e.x)) * 0.25f;
            v.y = ((a.y + b.y) + (d.y + e.y)) * 0.25f;
            v.z = ((a.z + b.z) + (d.z + e.z)) * 0.25f;
            v.w = ((a.w + b.w) + (d.w + e.w)) * 0.25f;
        } else {
            const int hi = MODE == 1 ? ho >> 1 : ho, wi = MODE == 1 ? wo >> 1 : wo;
            v = ld4<TIN>(x + ((static_cast<size_t>(n) * H + hi) * W + wi) * ld_x + c);
        }
        const size_t opix = (static_cast<size_t>(n) * Ho + ho) * Wo + wo;
        if (yf) *reinterpret_cast<float4*>(yf + opix * ld_yf + c) = v;
        if (yo) {
            if (TF32)
                *reinterpret_cast<float4*>(static_cast<float*>(yo) + opix * ld_yo + c) =
                    make_float4(op_f32(v.x, rnd), op_f32(v.y, rnd), op_f32(v.z, rnd), op_f32(v.w, rnd));
            else
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(yo) + opix * ld_yo + c) =
                    make_uint2(pack_op16x2(v.x, v.y, rnd), pack_op16x2(v.z, v.w, rnd));
        }
    }
}

// ---------------------------------------------------------------- first C channels of an NHWC fp32 tensor -> NCHW
// (output of the tensor-core conv_out, whose 3 | 6 real channels sit in a 64-channel padded tile)
__global__ void __launch_bounds__(256) nhwc_head_to_nchw_kernel(const float* __restrict__ x, int ld, long long npix,
                                                                 int HW, int C, float* __restrict__ out) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < npix;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long n = pix / HW;
        const int hw = static_cast<int>(pix - n * HW);
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + pix * ld));
        const float4 b = C > 4 ? __ldg(reinterpret_cast<const float4*>(x + pix * ld) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float* o = out + n * C * HW + hw;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < C) o[static_cast<size_t>(c) * HW] = v[c];
    }
}

// ---------------------------------------------------------------- timestep embedding
__global__ void temb_kernel(const float* __restrict__ t, int B, const float* __restrict__ freqs, int half,
                            int cos_first, float* __restrict__ out, int ld) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, k = i - b * half;
    const float a = t[b] * freqs[k];
    float s, c;
    sincosf(a, &s, &c);
    out[static_cast<size_t>(b) * ld + k] = cos_first ? c : s;
    out[static_cast<size_t>(b) * ld + half + k] = cos_first ? s : c;
}

// ---------------------------------------------------------------- small dense layer (fp32 SIMT GEMM)
__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return v / (1.0f + expf(-v));                            // SiLU
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // GELU (erf form, nn.GELU default)
    return v;
}

constexpr int kLinBM = 32, kLinBN = 64, kLinBK = 32;
// y[b,n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] )
// The layers behind this are weight-streaming (ADM's fused emb_layers: [32,1024] x [1024, ~40k] = 164 MB of weights per
// call), so the kernel is built to keep loads in flight: 128-bit global loads of the next K chunk are issued into
// registers before the FMAs of the current one (VEC: rows 16-byte aligned and K % 4 == 0; otherwise scalar loads).
template <bool VEC>
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, int ld_x, int B, int K,
                                                      const float* __restrict__ Wt, const float* __restrict__ bias,
                                                      int N, int act_in, int act_out, float* __restrict__ y, int ld_y) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ float xs[kLinBK][kLinBM + 1];
    __shared__ float ws[kLinBK][kLinBN + 1];
    const int b0 = blockIdx.y * kLinBM, n0 = blockIdx.x * kLinBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, each 2 rows x 4 cols
    float acc[2][4] = {};
    if (VEC) {
        // per thread and chunk: one float4 of x (row xr, columns xk..xk+3), two float4 of W (rows wr, wr + 32)
        const int xr = threadIdx.x >> 3, xk = (threadIdx.x & 7) << 2;
        const int wr = threadIdx.x >> 3, wk = (threadIdx.x & 7) << 2;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load = [&](int k0, float4& vx, float4& vw0, float4& vw1) {
            vx = (b0 + xr < B && k0 + xk < K) ? __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(b0 + xr) * ld_x + k0 + xk)) : zero;
            vw0 = (n0 + wr < N && k0 + wk < K) ? __ldg(reinterpret_cast<const float4*>(Wt + static_cast<size_t>(n0 + wr) * K + k0 + wk)) : zero;
            vw1 = (n0 + wr + 32 < N && k0 + wk < K) ? __ldg(reinterpret_cast<const float4*>(Wt + static_cast<size_t>(n0 + wr + 32) * K + k0 + wk)) : zero;
        };
        float4 vx, vw0, vw1;
        load(0, vx, vw0, vw1);
        for (int k0 = 0; k0 < K; k0 += kLinBK) {
            xs[xk][xr] = act_apply(vx.x, act_in), xs[xk + 1][xr] = act_apply(vx.y, act_in);
            xs[xk + 2][xr] = act_apply(vx.z, act_in), xs[xk + 3][xr] = act_apply(vx.w, act_in);
            ws[wk][wr] = vw0.x, ws[wk + 1][wr] = vw0.y, ws[wk + 2][wr] = vw0.z, ws[wk + 3][wr] = vw0.w;
            ws[wk][wr + 32] = vw1.x, ws[wk + 1][wr + 32] = vw1.y, ws[wk + 2][wr + 32] = vw1.z, ws[wk + 3][wr + 32] = vw1.w;
            __syncthreads();
            if (k0 + kLinBK < K) load(k0 + kLinBK, vx, vw0, vw1);  // in flight while the FMAs below run
#pragma unroll 8
            for (int kk = 0; kk < kLinBK; ++kk) {
                const float a0 = xs[kk][ty * 2], a1 = xs[kk][ty * 2 + 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = ws[kk][tx + 16 * j];
                    acc[0][j] = fmaf(a0, w, acc[0][j]);
                    acc[1][j] = fmaf(a1, w, acc[1][j]);
                }
            }
            __syncthreads();
        }
    } else {
        for (int k0 = 0; k0 < K; k0 += kLinBK) {
            for (int i = threadIdx.x; i < kLinBM * kLinBK; i += 256) {
                const int r = i / kLinBK, kk = i - r * kLinBK;
                float v = 0.f;
                if (b0 + r < B && k0 + kk < K) v = act_apply(x[static_cast<size_t>(b0 + r) * ld_x + k0 + kk], act_in);
                xs[kk][r] = v;
            }
            for (int i = threadIdx.x; i < kLinBN * kLinBK; i += 256) {
                const int r = i / kLinBK, kk = i - r * kLinBK;
                float v = 0.f;
                if (n0 + r < N && k0 + kk < K) v = Wt[static_cast<size_t>(n0 + r) * K + k0 + kk];
                ws[kk][r] = v;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < kLinBK; ++kk) {
                const float a0 = xs[kk][ty * 2], a1 = xs[kk][ty * 2 + 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = ws[kk][tx + 16 * j];
                    acc[0][j] = fmaf(a0, w, acc[0][j]);
                    acc[1][j] = fmaf(a1, w, acc[1][j]);
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int b = b0 + ty * 2 + i;
        if (b >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) y[static_cast<size_t>(b) * ld_y + n] = act_apply(acc[i][j] + (bias ? bias[n] : 0.f), act_out);
        }
    }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_resample(nlc_ctx* ctx, const float* x, int ld_x, int B, int H, int W, int C, int mode, float* y_f32,
                            int ld_y_f32, void* y_op, int ld_y_op, int op_dtype, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && (y_f32 || y_op), "nlc_resample: null argument");
    NLC_REQUIRE(mode >= 0 && mode <= 2 && C % 4 == 0 && ld_x % 4 == 0, "nlc_resample: bad mode/C");
    NLC_REQUIRE(mode != 2 || (H % 2 == 0 && W % 2 == 0), "nlc_resample: avgpool needs even H, W");
    NLC_REQUIRE(!y_f32 || ld_y_f32 % 4 == 0, "nlc_resample: ld_y_f32 %% 4");
    NLC_REQUIRE(!y_op || ld_y_op % 4 == 0, "nlc_resample: ld_y_op %% 4");
    const int Ho = mode == 1 ? 2 * H : (mode == 2 ? H / 2 : H);
    const int Wo = mode == 1 ? 2 * W : (mode == 2 ? W / 2 : W);
    const long long total4 = static_cast<long long>(B) * Ho * Wo * (C / 4);
    long long blocks = (total4 + 255) / 256;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_resample: bad op_dtype");
    const bool tf32 = !dtype_is16(op_dtype);
    const int rnd = dtype_fmt(op_dtype);
#define NLC_RS(M, T)                                                                                          \
    launch_pdl((resample_kernel<M, T, float>), dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, x, ld_x, H, W, C, y_f32,      \
                                                                                    ld_y_f32, y_op, ld_y_op, total4, rnd)
    if (mode == 0) {
        if (tf32) NLC_RS(0, true); else NLC_RS(0, false);
    } else if (mode == 1) {
        if (tf32) NLC_RS(1, true); else NLC_RS(1, false);
    } else {
        if (tf32) NLC_RS(2, true); else NLC_RS(2, false);
    }
#undef NLC_RS
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_resample_op(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C,
                               int mode, void* y_op, int ld_y, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_op && y_op, "nlc_resample_op: null argument");
    NLC_REQUIRE((mode == 1 || mode == 2) && C % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0, "nlc_resample_op: bad mode/C");
    NLC_REQUIRE(mode != 2 || (H % 2 == 0 && W % 2 == 0), "nlc_resample_op: avgpool needs even H, W");
    const int Ho = mode == 1 ? 2 * H : H / 2, Wo = mode == 1 ? 2 * W : W / 2;
    const long long total4 = static_cast<long long>(B) * Ho * Wo * (C / 4);
    long long blocks = (total4 + 255) / 256;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    const unsigned g = static_cast<unsigned>(blocks);
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_resample_op: bad op_dtype");
    const int rnd = dtype_fmt(op_dtype);
    if (!dtype_is16(op_dtype)) {
        const float* x = static_cast<const float*>(x_op);
        if (mode == 1) launch_pdl((resample_kernel<1, true, float>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
        else launch_pdl((resample_kernel<2, true, float>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
    } else if (op_dtype == NLC_F16) {
        const __half* x = static_cast<const __half*>(x_op);
        if (mode == 1)
            launch_pdl((resample_kernel<1, false, __half>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
        else
            launch_pdl((resample_kernel<2, false, __half>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
    } else {
        const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_op);
        if (mode == 1)
            launch_pdl((resample_kernel<1, false, __nv_bfloat16>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
        else
            launch_pdl((resample_kernel<2, false, __nv_bfloat16>), dim3(g), dim3(256), 0, stream, x, ld_x, H, W, C, nullptr, 0, y_op, ld_y, total4, rnd);
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_nhwc_head_to_nchw(nlc_ctx* ctx, const float* x, int ld, int B, int H, int W, int C, float* out_nchw,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && out_nchw && C >= 1 && C <= 8 && ld >= 8 && ld % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                "nlc_nhwc_head_to_nchw: C=%d (1..8) ld=%d (>= 8, %% 4) unsupported", C, ld);
    const long long npix = static_cast<long long>(B) * H * W;
    long long blocks = (npix + 255) / 256;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    launch_pdl((nhwc_head_to_nchw_kernel), dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, x, ld, npix, H * W, C, out_nchw);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_timestep_embedding(nlc_ctx* ctx, const float* t, int B, const float* freqs, int half,
                                      int cos_first, float* out, int ld_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && t && freqs && out && half > 0 && ld_out >= 2 * half, "nlc_timestep_embedding: bad argument");
    const int total = B * half;
    launch_pdl((temb_kernel), dim3((total + 255) / 256), dim3(256), 0, stream, t, B, freqs, half, cos_first, out, ld_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_linear(nlc_ctx* ctx, const float* x, int ld_x, int B, int K, const float* W, const float* bias,
                          int N, int act_in, int act_out, float* y, int ld_y, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && W && y && B > 0 && K > 0 && N > 0, "nlc_linear: bad argument");
    dim3 grid((N + kLinBN - 1) / kLinBN, (B + kLinBM - 1) / kLinBM);
    const bool vec = K % 4 == 0 && ld_x % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    if (vec)
        launch_pdl((linear_kernel<true>), dim3(grid), dim3(256), 0, stream, x, ld_x, B, K, W, bias, N, act_in, act_out, y, ld_y);
    else
        launch_pdl((linear_kernel<false>), dim3(grid), dim3(256), 0, stream, x, ld_x, B, K, W, bias, N, act_in, act_out, y, ld_y);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
