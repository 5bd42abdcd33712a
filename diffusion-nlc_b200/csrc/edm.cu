// EDM Heun sampler arithmetic: the reference keeps the sample x in float64 and calls the
// network in float32 (src/experiments.py:777-843, 847-918).  Every O(B*d) operation of that loop is one of the four
// kernels below; they are HBM-bound, read each [B,3,R,R] tensor once with 16/32-byte accesses and keep the
// reference's evaluation order (no FMA contraction across the torch operator boundaries).
//
// Per-sample reductions are written as NLC_EDM_PARTS fixed partial sums per sample (deterministic order), summed
// by the caller.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

constexpr int kEdmThreads = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    v = warp_sum_d(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = threadIdx.x < (kEdmThreads >> 5) ? red[threadIdx.x] : 0.0;
    if (warp == 0) v = warp_sum_d(v);
    __syncthreads();
    return v;  // valid in warp 0
}

// grid (NLC_EDM_PARTS, B): x32 = float(x64); sumsq[b][part] = sum x64^2 over the part
__global__ void __launch_bounds__(kEdmThreads)
    edm_prepare_kernel(const double* __restrict__ x, int d, float* __restrict__ x32, double* __restrict__ sumsq) {
    __shared__ double red[8];
    const int part = blockIdx.x, b = blockIdx.y;
    const int per = d / NLC_EDM_PARTS;
    const size_t base = static_cast<size_t>(b) * d + static_cast<size_t>(part) * per;
    const double2* x2 = reinterpret_cast<const double2*>(x + base);
    float2* o2 = reinterpret_cast<float2*>(x32 + base);
    double acc = 0.0;
    for (int i = threadIdx.x; i < (per >> 1); i += kEdmThreads) {
        const double2 v = x2[i];
        acc += __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y));
        o2[i] = make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
    }
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0 && sumsq) sumsq[b * NLC_EDM_PARTS + part] = acc;
}

// eps = (x64 - double(c_skip*x32 + c_out*F)) / div     (pred_edm + src/experiments.py:836-840)
__global__ void __launch_bounds__(kEdmThreads)
    edm_eps_kernel(const double* __restrict__ x, const float* __restrict__ x32, const float* __restrict__ F,
                   const float* __restrict__ cskip, const float* __restrict__ cout, const double* __restrict__ div,
                   int d, double* __restrict__ eps, double* __restrict__ denoised, double* __restrict__ sumsq) {
    __shared__ double red[8];
    const int part = blockIdx.x, b = blockIdx.y;
    const int per = d / NLC_EDM_PARTS;
    const size_t base = static_cast<size_t>(b) * d + static_cast<size_t>(part) * per;
    const float cs = cskip[b], co = cout[b];
    const double dv = div[b];
    double acc = 0.0;
    for (int i = threadIdx.x; i < (per >> 1); i += kEdmThreads) {
        const double2 v = reinterpret_cast<const double2*>(x + base)[i];
        const float2 a = reinterpret_cast<const float2*>(x32 + base)[i];
        const float2 f = reinterpret_cast<const float2*>(F + base)[i];
        const double d0 = static_cast<double>(__fadd_rn(__fmul_rn(cs, a.x), __fmul_rn(co, f.x)));
        const double d1 = static_cast<double>(__fadd_rn(__fmul_rn(cs, a.y), __fmul_rn(co, f.y)));
        const double e0 = __ddiv_rn(__dsub_rn(v.x, d0), dv);
        const double e1 = __ddiv_rn(__dsub_rn(v.y, d1), dv);
        reinterpret_cast<double2*>(eps + base)[i] = make_double2(e0, e1);
        if (denoised) reinterpret_cast<double2*>(denoised + base)[i] = make_double2(d0, d1);
        acc += __dadd_rn(__dmul_rn(e0, e0), __dmul_rn(e1, e1));
    }
    acc = block_sum_d(acc, red);
    if (threadIdx.x == 0 && sumsq) sumsq[b * NLC_EDM_PARTS + part] = acc;
}

// v_a = post_a(e_a) * s_a[b],  post(e) = sqrt_d * e / den[b] when den != NULL (normalize, src/utils.py:11-16)
// out = w1 * v_1 (+ w2 * v_2);  partial sums of out^2, v_1^2 and out.v_1 (for the cosine-similarity scale)
__global__ void __launch_bounds__(kEdmThreads)
    edm_mix_kernel(const double* __restrict__ e1, const double* __restrict__ den1, const double* __restrict__ s1,
                   const double* __restrict__ e2, const double* __restrict__ den2, const double* __restrict__ s2,
                   double w1, double w2, double sqrt_d, int d, double* __restrict__ out,
                   double* __restrict__ sums /* [B][PARTS][3] */) {
    __shared__ double red[8];
    const int part = blockIdx.x, b = blockIdx.y;
    const int per = d / NLC_EDM_PARTS;
    const size_t base = static_cast<size_t>(b) * d + static_cast<size_t>(part) * per;
    const double sa = s1 ? s1[b] : 1.0, sb = (e2 && s2) ? s2[b] : 1.0;
    const double da = den1 ? den1[b] : 0.0, db = (e2 && den2) ? den2[b] : 0.0;
    double aoo = 0.0, avv = 0.0, aov = 0.0;
    for (int i = threadIdx.x; i < per; i += kEdmThreads) {
        double v = e1[base + i];
        if (den1) v = __ddiv_rn(__dmul_rn(sqrt_d, v), da);
        if (s1) v = __dmul_rn(v, sa);
        double o = v;
        if (e2) {
            double u = e2[base + i];
            if (den2) u = __ddiv_rn(__dmul_rn(sqrt_d, u), db);
            if (s2) u = __dmul_rn(u, sb);
            o = __dadd_rn(__dmul_rn(w1, v), __dmul_rn(w2, u));
        }
        out[base + i] = o;
        aoo += __dmul_rn(o, o), avv += __dmul_rn(v, v), aov += __dmul_rn(o, v);
    }
    aoo = block_sum_d(aoo, red);
    avv = block_sum_d(avv, red);
    aov = block_sum_d(aov, red);
    if (threadIdx.x == 0 && sums) {
        double* o3 = sums + (static_cast<size_t>(b) * NLC_EDM_PARTS + part) * 3;
        o3[0] = aoo, o3[1] = avv, o3[2] = aov;
    }
}

// x_next = x_hat + coef[b] * post(e),  post(e) = [sqrt_d*e/den[b]] then [/ eps_scale | * mul[b]]
// (+ optional churn noise: x_hat itself = x + nz_coef * z, src/experiments.py:877-880, when z != NULL)
__global__ void __launch_bounds__(kEdmThreads)
    edm_axpy_kernel(const double* __restrict__ xh, const double* __restrict__ e, const double* __restrict__ den,
                    double sqrt_d, double eps_scale, const double* __restrict__ mul, const double* __restrict__ coef,
                    int d, long long total, double* __restrict__ out) {
    for (long long i = static_cast<long long>(blockIdx.x) * kEdmThreads + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * kEdmThreads) {
        const int b = static_cast<int>(i / d);
        double v = e[i];
        if (den) v = __ddiv_rn(__dmul_rn(sqrt_d, v), den[b]);
        if (eps_scale != 0.0) v = __ddiv_rn(v, eps_scale);
        if (mul) v = __dmul_rn(v, mul[b]);
        out[i] = __dadd_rn(xh[i], __dmul_rn(coef[b], v));
    }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_edm_prepare(nlc_ctx* ctx, const double* x64, int B, int d, float* x32, double* sumsq_parts,
                               void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x64 && x32 && B > 0, "nlc_edm_prepare: null argument");
    NLC_REQUIRE(d % (2 * NLC_EDM_PARTS) == 0, "nlc_edm_prepare: d=%d must be a multiple of %d", d, 2 * NLC_EDM_PARTS);
    edm_prepare_kernel<<<dim3(NLC_EDM_PARTS, B), kEdmThreads, 0, stream>>>(x64, d, x32, sumsq_parts);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_edm_eps(nlc_ctx* ctx, const double* x64, const float* x32, const float* F, const float* c_skip,
                           const float* c_out, const double* div, int B, int d, double* eps, double* denoised,
                           double* sumsq_parts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x64 && x32 && F && c_skip && c_out && div && eps && B > 0, "nlc_edm_eps: null argument");
    NLC_REQUIRE(d % (2 * NLC_EDM_PARTS) == 0, "nlc_edm_eps: d=%d must be a multiple of %d", d, 2 * NLC_EDM_PARTS);
    edm_eps_kernel<<<dim3(NLC_EDM_PARTS, B), kEdmThreads, 0, stream>>>(x64, x32, F, c_skip, c_out, div, d, eps,
                                                                        denoised, sumsq_parts);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_edm_mix(nlc_ctx* ctx, const double* e1, const double* den1, const double* s1, const double* e2,
                           const double* den2, const double* s2, double w1, double w2, int B, int d, double* out,
                           double* sums_parts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && e1 && out && B > 0, "nlc_edm_mix: null argument");
    NLC_REQUIRE(d % NLC_EDM_PARTS == 0, "nlc_edm_mix: d=%d must be a multiple of %d", d, NLC_EDM_PARTS);
    edm_mix_kernel<<<dim3(NLC_EDM_PARTS, B), kEdmThreads, 0, stream>>>(e1, den1, s1, e2, den2, s2, w1, w2,
                                                                        sqrt(static_cast<double>(d)), d, out,
                                                                        sums_parts);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_edm_axpy(nlc_ctx* ctx, const double* x_hat, const double* e, const double* den, double eps_scale,
                            const double* mul, const double* coef, int B, int d, double* x_next, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x_hat && e && coef && x_next && B > 0, "nlc_edm_axpy: null argument");
    const long long total = static_cast<long long>(B) * d;
    long long blocks = (total + kEdmThreads - 1) / kEdmThreads;
    const long long cap = static_cast<long long>(ctx->sm_count) * 16;
    if (blocks > cap) blocks = cap;
    edm_axpy_kernel<<<static_cast<unsigned>(blocks), kEdmThreads, 0, stream>>>(
        x_hat, e, den, sqrt(static_cast<double>(d)), eps_scale, mul, coef, d, total, x_next);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
