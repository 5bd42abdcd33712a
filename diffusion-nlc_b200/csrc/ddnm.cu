// DDNM+: the noisy-measurement terms of functions/svd_operators.py
//   Lambda(v)          = V (lambda o V^T v)                     (:253-274, 361-387, 464-469, 535-570, 669-693, 1016-1042)
//   Lambda_noise(v, e) = V (d1 o P v) + V (d2 o P e)            (:276-320, 389-439, 471-476, 572-623, 695-736, 1044-1091)
// and one fused reverse step of functions/svd_ddnm.py (ddnm_diffusion :40-66, ddnm_plus_diffusion :101-132).
//
// The reference rebuilds full-length lambda / d1 / d2 vectors with ~25 elementwise launches per call and applies V, V^T as
// clone / permute / index_put / matmul chains.  Here the three factors are a pure function of a component's singular value
// (`ddnm_terms`, evaluated in registers, fp32 op for op in the reference's order) and every operator gets the closed form
// its structure allows:
//   Colorization / SuperResolution   one thread per needle (pixel's channels / r x r patch): K x K rotation in registers,
//                                    the whole step = one pass: 4 reads (xt, et, z, y) + 2 writes (x0_t, x_next)
//   Inpainting / Denoising           elementwise, same traffic
//   WalshHadamardCS                  V^T V cancels between A^+ and Lambda, so a step is two 2-D FWHTs + two elementwise
//                                    passes (x_next = a x0 + FWHT(-a lambda T + d1 z + d2 et) by linearity)
//   Deblurring                       per-step factor tables over the R x R spectral positions + the separable GEMM chain
// All HBM-bound fp32.
#include <math.h>
#include <string.h>

#include "operators.h"

namespace nlc {

struct Coef {
    float a, st, sy, sy2, eta, root;
    int active;  // a != 0 and sigma_y != 0 (the reference's guard, e.g. :265)
};
static Coef make_coef(float a, float sigma_t, double sigma_y, double eta) {
    Coef c;
    c.a = a, c.st = sigma_t, c.sy = static_cast<float>(sigma_y), c.sy2 = static_cast<float>(sigma_y * sigma_y);
    c.eta = static_cast<float>(eta), c.root = static_cast<float>(pow(1.0 - eta * eta, 0.5));
    c.active = a != 0.f && sigma_y != 0.0;
    return c;
}
// lambda (Eq. 17), d1 / d2 (Eq. 51) of a component with singular value s (0 = null space)
__device__ __forceinline__ void ddnm_terms(float s, const Coef& c, float& lam, float& d1, float& d2) {
    lam = 1.f, d1 = __fmul_rn(c.st, c.eta), d2 = __fmul_rn(c.st, c.root);
    if (!c.active || s == 0.f) return;
    const float inv = __fdiv_rn(1.f, s);
    const float thr = __fmul_rn(__fmul_rn(c.a, c.sy), inv);
    if (c.st < thr) {
        lam = __fdiv_rn(__fdiv_rn(__fmul_rn(__fmul_rn(s, c.st), c.root), c.a), c.sy);
        d2 = 0.f;
    } else if (c.st > thr) {
        d1 = __fsqrt_rn(__fsub_rn(__fmul_rn(c.st, c.st),
                                  __fmul_rn(__fmul_rn(__fmul_rn(c.a, c.a), c.sy2), __fmul_rn(inv, inv))));
        d2 = 0.f;
    }
}

struct Step {
    float c_at, d_at;  // sqrt(1 - alpha_bar_t), sqrt(alpha_bar_t)
    float a;           // sqrt(alpha_bar_{t-1})
    float c1, c2;      // DDNM: sigma_t eta, sigma_t sqrt(1 - eta^2)
    int plus;
    Coef coef;
};
__device__ __forceinline__ float x0_of(float xt, float et, const Step& s) {
    return __fdiv_rn(__fsub_rn(xt, __fmul_rn(et, s.c_at)), s.d_at);  // (xt - et sqrt(1 - at)) / sqrt(at)
}

enum { M_LAMBDA = 0, M_NOISE = 1, M_STEP = 2 };

// ---------------------------------------------------------------- vector access helpers
// V consecutive floats per thread (V = 4: one 128-bit access; V = 1: the fallback for odd sizes / unaligned pointers)
template <int V> struct Pack { float v[V]; };
template <int V>
__device__ __forceinline__ Pack<V> ldv(const float* p) {
    Pack<V> r;
    if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r.v[0] = t.x, r.v[1] = t.y, r.v[2] = t.z, r.v[3] = t.w;
    } else if constexpr (V == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        r.v[0] = t.x, r.v[1] = t.y;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const Pack<V>& a) {
    if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
    else if constexpr (V == 2) *reinterpret_cast<float2*>(p) = make_float2(a.v[0], a.v[1]);
    else p[0] = a.v[0];
}
template <int V>
__device__ __forceinline__ void ldiv(const int* p, int* dst) {
    if constexpr (V == 4) {
        const int4 t = __ldg(reinterpret_cast<const int4*>(p));
        dst[0] = t.x, dst[1] = t.y, dst[2] = t.z, dst[3] = t.w;
    } else {
        dst[0] = __ldg(p);
    }
}

// ---------------------------------------------------------------- colourisation / average-pool SR
// One needle = the K entries one measurement is taken from (a pixel's channels / an r x r patch).  V_small (K x K) comes
// in as a kernel argument for K <= 16, so its entries are constant-bank operands of the FMAs (no load instruction at all);
// r = 8 (K = 64) stages it in shared memory.
template <int K> struct NeedleGeom { static constexpr int RW = K == 16 ? 4 : (K == 4 ? 2 : 1); };
template <int K> struct VMat { float v[K <= 16 ? K * K : 1]; };
struct NeedleTerms { float lam0, d10, d20, d1N, d2N; };

// in: n = v | x_t, e = eps | e_t, zz = - | z, yv = measurement.  out: o = result | x_next; n is overwritten by x0_t (M_STEP)
template <int K, typename VT>
__device__ __forceinline__ void needle_math(int mode, float* n, const float* e, float* zz, float yv, const VT& V, float u,
                                            float s, const Step& sc, const NeedleTerms& T, float* o) {
    constexpr int UN = K <= 16 ? K : 1;  // K = 64 keeps its needles in local memory instead of 256 registers
    if (mode == M_LAMBDA) {
        float w[K];
#pragma unroll UN
        for (int kp = 0; kp < K; ++kp) {
            float acc = 0.f;
#pragma unroll UN
            for (int k = 0; k < K; ++k) acc = fmaf(V[k * K + kp], n[k], acc);  // V^T n
            w[kp] = kp == 0 ? __fmul_rn(acc, T.lam0) : acc;
        }
#pragma unroll UN
        for (int j = 0; j < K; ++j) {
            float acc = 0.f;
#pragma unroll UN
            for (int kp = 0; kp < K; ++kp) acc = fmaf(V[j * K + kp], w[kp], acc);
            o[j] = acc;
        }
        return;
    }
    if (mode == M_NOISE) {
#pragma unroll UN
        for (int k = 0; k < K; ++k) zz[k] = n[k];
    } else {
        float dot = 0.f;
#pragma unroll UN
        for (int k = 0; k < K; ++k) n[k] = x0_of(n[k], e[k], sc);
#pragma unroll UN
        for (int k = 0; k < K; ++k) dot = fmaf(V[k * K], n[k], dot);  // v0 = first column of V
        const float meas = u * (s * dot);                              // A x0
        const float t = (u * (meas - yv)) * (1.0f / s);                // spectral A^+(A x0 - y)
        if (!sc.plus) {
#pragma unroll UN
            for (int k = 0; k < K; ++k) {
                const float x0h = n[k] - V[k * K] * t;
                o[k] = __fadd_rn(__fadd_rn(__fmul_rn(sc.a, x0h), __fmul_rn(sc.c1, zz[k])), __fmul_rn(sc.c2, e[k]));
            }
            return;
        }
        // A^+(A x0 - y) = v0 t lies along the first right singular vector, so Lambda (V diag(lambda) V^T) scales it by lambda_0
        const float tl = __fmul_rn(t, T.lam0);
#pragma unroll UN
        for (int k = 0; k < K; ++k) o[k] = n[k] - V[k * K] * tl;  // x0_hat
    }
    // V (d1 o z + d2 o e): channel / patch entry k stands in for component k (:575-581, 698-699)
#pragma unroll UN
    for (int k = 0; k < K; ++k)
        zz[k] = __fadd_rn(__fmul_rn(zz[k], k == 0 ? T.d10 : T.d1N), __fmul_rn(e[k], k == 0 ? T.d20 : T.d2N));
#pragma unroll UN
    for (int j = 0; j < K; ++j) {
        float acc = 0.f;
#pragma unroll UN
        for (int k = 0; k < K; ++k) acc = fmaf(V[j * K + k], zz[k], acc);
        o[j] = mode == M_NOISE ? acc : __fadd_rn(__fmul_rn(sc.a, o[j]), acc);
    }
}
__device__ __forceinline__ NeedleTerms needle_terms(float s, const Step& sc) {
    NeedleTerms T;
    float lamN;
    ddnm_terms(s, sc.coef, T.lam0, T.d10, T.d20);
    ddnm_terms(0.f, sc.coef, lamN, T.d1N, T.d2N);
    return T;
}

// One thread per needle.  in1 = v | xt, in2 = eps | et (sample stride in2_stride), out1 = result | x_next, out0 = x0_t.
// RW = contiguous floats per needle row: the r x r patch of SR is read as r row vectors (r = 2, 4), colour planes scalar.
template <int K>
__global__ void __launch_bounds__(128) needle_ddnm_kernel(int mode, const float* __restrict__ in1,
                                                           const float* __restrict__ in2, long long in2_stride,
                                                           const float* __restrict__ z, const float* __restrict__ y,
                                                           float* __restrict__ out1, float* __restrict__ out0, int B, int C,
                                                           int R, int r, int per_ch, float u, float s,
                                                           const __grid_constant__ VMat<K> Vc,
                                                           const float* __restrict__ Vfull, const Step sc, int vec_ok) {
    constexpr int UN = K <= 16 ? K : 1;
    constexpr int RW = NeedleGeom<K>::RW;
    __shared__ __align__(16) float Vs[K <= 16 ? 1 : K * K];
    if constexpr (K > 16) {
        for (int t = threadIdx.x; t < K * K; t += blockDim.x) Vs[t] = Vfull[t];
        __syncthreads();
    }
    const int yd = R / r;
    const long long plane = static_cast<long long>(R) * R;
    const long long per_sample = per_ch ? static_cast<long long>(C) * yd * yd : plane;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    size_t inner;
    if (per_ch) {
        const int j = static_cast<int>(q % yd), ii = static_cast<int>((q / yd) % yd), c = static_cast<int>(q / (yd * yd));
        inner = (static_cast<size_t>(c) * R + static_cast<size_t>(ii) * r) * R + static_cast<size_t>(j) * r;
    } else {
        inner = static_cast<size_t>(q);
    }
    const size_t base = static_cast<size_t>(b) * C * plane + inner;
    const size_t base2 = static_cast<size_t>(b) * (in2_stride ? in2_stride : C * plane) + inner;
    auto off = [&](int k) -> size_t {
        return per_ch ? static_cast<size_t>(k / r) * R + (k % r) : static_cast<size_t>(k) * plane;
    };
    auto load = [&](const float* p, size_t bs, float* dst) {
        if (RW > 1 && vec_ok) {
#pragma unroll
            for (int rr = 0; rr < (RW > 1 ? RW : 1); ++rr) {
                const Pack<RW> row = ldv<RW>(p + bs + static_cast<size_t>(rr) * R);
#pragma unroll
                for (int cc = 0; cc < RW; ++cc) dst[rr * RW + cc] = row.v[cc];
            }
        } else {
#pragma unroll UN
            for (int k = 0; k < K; ++k) dst[k] = p[bs + off(k)];
        }
    };
    auto store = [&](float* p, const float* src) {
        if (RW > 1 && vec_ok) {
#pragma unroll
            for (int rr = 0; rr < (RW > 1 ? RW : 1); ++rr) {
                Pack<RW> row;
#pragma unroll
                for (int cc = 0; cc < RW; ++cc) row.v[cc] = src[rr * RW + cc];
                stv<RW>(p + base + static_cast<size_t>(rr) * R, row);
            }
        } else {
#pragma unroll UN
            for (int k = 0; k < K; ++k) p[base + off(k)] = src[k];
        }
    };
    const NeedleTerms T = needle_terms(s, sc);
    float n[K], e[K], zz[K], o[K];
    load(in1, base, n);
    if (mode != M_LAMBDA) load(in2, base2, e);
    if (mode == M_STEP) load(z, base, zz);
    const float yv = mode == M_STEP ? y[i] : 0.f;
    if constexpr (K <= 16) needle_math<K>(mode, n, e, zz, yv, Vc.v, u, s, sc, T, o);
    else needle_math<K>(mode, n, e, zz, yv, Vs, u, s, sc, T, o);
    if (mode == M_STEP) store(out0, n);
    store(out1, o);
}
// colourisation, four adjacent pixels per thread: every plane moves as 128-bit vectors
__global__ void __launch_bounds__(256) color_ddnm_vec_kernel(int mode, const float* __restrict__ in1,
                                                              const float* __restrict__ in2, long long in2_stride,
                                                              const float* __restrict__ z, const float* __restrict__ y,
                                                              float* __restrict__ out1, float* __restrict__ out0,
                                                              long long n4, long long plane, float u, float s,
                                                              const __grid_constant__ VMat<3> Vc, const Step sc) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long b = (i * 4) / plane, p = i * 4 - b * plane;
    const size_t base = static_cast<size_t>(b) * 3 * plane + p, base2 = static_cast<size_t>(b) * in2_stride + p;
    const NeedleTerms T = needle_terms(s, sc);
    Pack<4> a1[3], a2[3], zv[3], yv, x0[3], o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a1[c] = ldv<4>(in1 + base + c * plane);
        if (mode != M_LAMBDA) a2[c] = ldv<4>(in2 + base2 + c * plane);
        if (mode == M_STEP) zv[c] = ldv<4>(z + base + c * plane);
    }
    if (mode == M_STEP) yv = ldv<4>(y + b * plane + p);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float n[3], e[3], zz[3], oo[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) n[c] = a1[c].v[k], e[c] = a2[c].v[k], zz[c] = zv[c].v[k];
        needle_math<3>(mode, n, e, zz, yv.v[k], Vc.v, u, s, sc, T, oo);
#pragma unroll
        for (int c = 0; c < 3; ++c) x0[c].v[k] = n[c], o[c].v[k] = oo[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (mode == M_STEP) stv<4>(out0 + base + c * plane, x0[c]);
        stv<4>(out1 + base + c * plane, o[c]);
    }
}

// ---------------------------------------------------------------- inpainting / denoising (V is a permutation / identity)
// pos2k == nullptr: Denoising (every entry kept, s = 1, and the class' own scalar rules :464-476); otherwise pos2k is the
// image-order table (nlc_op::idx_c).  V consecutive entries per thread.
template <int V>
__global__ void __launch_bounds__(256) mask_ddnm_kernel(int mode, const float* __restrict__ in1,
                                                         const float* __restrict__ in2, long long in2_stride,
                                                         const float* __restrict__ z, const float* __restrict__ y,
                                                         float* __restrict__ out1, float* __restrict__ out0, int B,
                                                         long long per_sample, const int* __restrict__ pos2k, int n_kept,
                                                         const Step sc) {
    const long long iv = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long i = iv * V;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    const size_t i2 = static_cast<size_t>(b) * (in2_stride ? in2_stride : per_sample) + q;
    float lamK, d1K, d2K, lamN, d1N, d2N;
    bool scale_first = false;  // Denoising's `vec * sigma_t * eta` multiplies in that order
    int ks[V];
    if (pos2k) {
        ldiv<V>(pos2k + q, ks);
        ddnm_terms(1.f, sc.coef, lamK, d1K, d2K);
        ddnm_terms(0.f, sc.coef, lamN, d1N, d2N);
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e) ks[e] = static_cast<int>(q) + e;
        const Coef& c = sc.coef;
        const float thr = __fmul_rn(c.a, c.sy);
        lamK = c.st < thr ? __fdiv_rn(__fdiv_rn(__fmul_rn(c.st, c.root), c.a), c.sy) : 1.f;
        d2K = 0.f;
        if (c.st >= thr) d1K = __fsqrt_rn(__fsub_rn(__fmul_rn(c.st, c.st), __fmul_rn(__fmul_rn(c.a, c.a), c.sy2)));
        else d1K = 0.f, scale_first = true;
        lamN = lamK, d1N = d1K, d2N = d2K;
    }
    auto noise = [&](float zv, float ev, bool kept) -> float {
        if (scale_first) return __fmul_rn(__fmul_rn(zv, sc.coef.st), sc.coef.eta);
        if (!pos2k) return __fmul_rn(zv, d1K);
        return __fadd_rn(__fmul_rn(zv, kept ? d1K : d1N), __fmul_rn(ev, kept ? d2K : d2N));
    };
    const Pack<V> a1 = ldv<V>(in1 + i);
    Pack<V> o;
    if (mode == M_LAMBDA) {
#pragma unroll
        for (int e = 0; e < V; ++e) o.v[e] = __fmul_rn(a1.v[e], ks[e] >= 0 ? lamK : lamN);
        stv<V>(out1 + i, o);
        return;
    }
    const Pack<V> a2 = ldv<V>(in2 + i2);
    if (mode == M_NOISE) {
#pragma unroll
        for (int e = 0; e < V; ++e) o.v[e] = noise(a1.v[e], a2.v[e], ks[e] >= 0);
        stv<V>(out1 + i, o);
        return;
    }
    const Pack<V> zv = ldv<V>(z + i);
    Pack<V> x0;
    const float* yb = y + static_cast<size_t>(b) * n_kept;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        x0.v[e] = x0_of(a1.v[e], a2.v[e], sc);
        const bool kept = ks[e] >= 0;
        const float resid = kept ? __fsub_rn(x0.v[e], __ldg(yb + ks[e])) : 0.f;
        if (sc.plus) {
            const float x0h = __fsub_rn(x0.v[e], __fmul_rn(resid, kept ? lamK : lamN));
            o.v[e] = __fadd_rn(__fmul_rn(sc.a, x0h), noise(zv.v[e], a2.v[e], kept));
        } else {
            const float x0h = __fsub_rn(x0.v[e], resid);
            o.v[e] = __fadd_rn(__fadd_rn(__fmul_rn(sc.a, x0h), __fmul_rn(sc.c1, zv.v[e])), __fmul_rn(sc.c2, a2.v[e]));
        }
    }
    stv<V>(out0 + i, x0);
    stv<V>(out1 + i, o);
}

// ---------------------------------------------------------------- shared elementwise pieces (WH-CS, Deblurring)
template <int V>
__global__ void __launch_bounds__(256) x0_kernel(const float* __restrict__ xt, const float* __restrict__ et,
                                                  long long et_stride, float* __restrict__ x0, long long per_sample, int B,
                                                  const Step sc) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * V;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample;
    const Pack<V> x = ldv<V>(xt + i), e = ldv<V>(et + static_cast<size_t>(b) * et_stride + (i - b * per_sample));
    Pack<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = x0_of(x.v[k], e.v[k], sc);
    stv<V>(x0 + i, o);
}
// out = g1[o] v + g2[o] e with the factors of spectral position o = i % plane:
//   invperm != nullptr (WH-CS): kept (invperm[o] < m) -> s = 1, else null space;  tab != nullptr (Deblurring): s = tab[o]
// F != nullptr adds the Lambda'd residual of the step:  -a lambda (F - y) on kept entries (WH-CS only)
template <int V>
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ v, const float* __restrict__ e,
                                                   long long e_stride, const float* __restrict__ F,
                                                   const float* __restrict__ y, float* __restrict__ out, int B, int C,
                                                   long long plane, const int* __restrict__ invperm, int m,
                                                   const float* __restrict__ tab, const Step sc) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * V;
    const long long per_sample = static_cast<long long>(C) * plane;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    const int o = static_cast<int>(q % plane), c = static_cast<int>(q / plane);
    const Pack<V> vv = ldv<V>(v + i), ee = ldv<V>(e + static_cast<size_t>(b) * e_stride + q);
    Pack<V> ff, ss, res;
    int js[V];
    if (invperm) ldiv<V>(invperm + o, js);
    else ss = ldv<V>(tab + o);
    if (F) ff = ldv<V>(F + i);
#pragma unroll
    for (int k = 0; k < V; ++k) {
        float lam, d1, d2;
        ddnm_terms(invperm ? (js[k] < m ? 1.f : 0.f) : ss.v[k], sc.coef, lam, d1, d2);
        float r = __fadd_rn(__fmul_rn(vv.v[k], d1), __fmul_rn(ee.v[k], d2));
        if (F && invperm && js[k] < m) {
            const float resid = ff.v[k] - __ldg(y + static_cast<size_t>(b) * m * C + static_cast<size_t>(js[k]) * C + c);
            r = fmaf(-sc.a * lam, resid, r);
        }
        res.v[k] = r;
    }
    stv<V>(out + i, res);
}
// WH-CS Lambda in the transform domain: F[o] *= lambda on kept entries
__global__ void __launch_bounds__(256) whcs_lambda_kernel(float* __restrict__ F, long long n, long long plane,
                                                           const int* __restrict__ invperm, int m, const Step sc) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float lam, d1, d2;
    ddnm_terms(invperm[i % plane] < m ? 1.f : 0.f, sc.coef, lam, d1, d2);
    F[i] = __fmul_rn(F[i], lam);
}
// Deblurring: per-step factor tables over the m*m spectral positions.  lam[o], and comb[c][o] = -a lambda[o] pinv[c][o]
__global__ void __launch_bounds__(256) deblur_tables_kernel(const float* __restrict__ lam_s, const float* __restrict__ pinv,
                                                             int C, int mm, float* __restrict__ lam_out,
                                                             float* __restrict__ comb, const Step sc) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= mm) return;
    float lam, d1, d2;
    ddnm_terms(lam_s[o], sc.coef, lam, d1, d2);
    lam_out[o] = lam;
    if (comb)
        for (int c = 0; c < C; ++c) comb[c * mm + o] = -sc.a * lam * pinv[c * mm + o];
}
// out = g1 z + g2 e (the plain DDNM noise terms, sample stride for e)
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ z, const float* __restrict__ e,
                                                     long long e_stride, float* __restrict__ out, long long per_sample,
                                                     int B, float g1, float g2) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample;
    out[i] = __fadd_rn(__fmul_rn(g1, z[i]), __fmul_rn(g2, e[static_cast<size_t>(b) * e_stride + (i - b * per_sample)]));
}
// x <- a x + c1 z + c2 e in place (x holds x0_hat)
__global__ void __launch_bounds__(256) assemble_kernel(float* __restrict__ x, const float* __restrict__ z,
                                                        const float* __restrict__ e, long long e_stride,
                                                        long long per_sample, int B, float a, float c1, float c2) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample;
    x[i] = __fadd_rn(__fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(c1, z[i])),
                     __fmul_rn(c2, e[static_cast<size_t>(b) * e_stride + (i - b * per_sample)]));
}
template <int V>
__global__ void __launch_bounds__(256) renoise_kernel(const float* __restrict__ x0, const float* __restrict__ z,
                                                       long long n, float a, float sig, float* __restrict__ out) {
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * V;
    if (i >= n) return;
    const Pack<V> x = ldv<V>(x0 + i), zz = ldv<V>(z + i);
    Pack<V> o;
#pragma unroll
    for (int k = 0; k < V; ++k) o.v[k] = __fadd_rn(__fmul_rn(a, x.v[k]), __fmul_rn(zz.v[k], sig));
    stv<V>(out + i, o);
}

static inline bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr,
                             const void* e = nullptr, const void* f = nullptr) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f)) & 15) == 0;
}

template <int K>
static int launch_needle(nlc_op* op, int mode, const float* in1, const float* in2, long long s2, const float* z,
                         const float* y, float* out1, float* out0, int B, const Step& sc, cudaStream_t st) {
    const int per_ch = op->task == NLC_OP_SR_AVG, r = per_ch ? op->ratio : 1;
    const long long plane = static_cast<long long>(op->R) * op->R;
    const long long n = per_ch ? static_cast<long long>(B) * op->C * (plane / (r * r)) : static_cast<long long>(B) * plane;
    const int vec_ok = aligned16(in1, in2, z, out1, out0) && s2 % 4 == 0 && op->R % 4 == 0;
    VMat<K> Vc;
    if (K <= 16) memcpy(Vc.v, op->Vfull_host, sizeof(float) * K * K);
    if constexpr (K == 3) {
        if (!per_ch && vec_ok && plane % 4 == 0 && aligned16(y)) {
            color_ddnm_vec_kernel<<<blocks_for(n / 4), 256, 0, st>>>(mode, in1, in2, s2, z, y, out1, out0, n / 4, plane,
                                                                     op->u, op->s, Vc, sc);
            NLC_CHECK_LAUNCH();
            return NLC_OK;
        }
    }
    needle_ddnm_kernel<K><<<blocks_for(n, 128), 128, 0, st>>>(mode, in1, in2, s2, z, y, out1, out0, B, op->C, op->R, r,
                                                               per_ch, op->u, op->s, Vc, op->Vfull, sc, vec_ok);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

// mode M_LAMBDA: in1 = v; M_NOISE: in1 = v, in2 = eps; M_STEP: in1 = xt, in2 = et (stride s2), z, y -> out0 = x0_t, out1
static int ddnm_apply(nlc_op* op, int mode, const float* in1, const float* in2, long long s2, const float* z,
                      const float* y, float* out1, float* out0, int B, const Step& sc, float* ws, cudaStream_t st) {
    const int C = op->C, R = op->R;
    const long long plane = static_cast<long long>(R) * R;
    const long long per_sample = op->task == NLC_OP_GENERAL ? op->nx : C * plane, total = per_sample * B;
    if (s2 == 0) s2 = per_sample;
    // 128-bit accesses when every tensor is 16-byte aligned and a sample / a plane is a whole number of float4
    const bool vec4 = aligned16(in1, in2, z, out1, out0, ws) && per_sample % 4 == 0 && s2 % 4 == 0 && plane % 4 == 0 &&
                      (op->task != NLC_OP_DENOISE || aligned16(y));
    int rc;
    switch (op->task) {
        case NLC_OP_COLOR:
        case NLC_OP_SR_AVG:
            switch (op->K) {
                case 3: return launch_needle<3>(op, mode, in1, in2, s2, z, y, out1, out0, B, sc, st);
                case 4: return launch_needle<4>(op, mode, in1, in2, s2, z, y, out1, out0, B, sc, st);
                case 9: return launch_needle<9>(op, mode, in1, in2, s2, z, y, out1, out0, B, sc, st);
                case 16: return launch_needle<16>(op, mode, in1, in2, s2, z, y, out1, out0, B, sc, st);
                case 64: return launch_needle<64>(op, mode, in1, in2, s2, z, y, out1, out0, B, sc, st);
                default:
                    return set_error(NLC_EINVAL, "DDNM+: pooling factors 2, 3, 4 and 8 are built (got needle length %d)",
                                     op->K);
            }
        case NLC_OP_INPAINT:
        case NLC_OP_DENOISE:
            if (vec4)
                mask_ddnm_kernel<4><<<blocks_for(total / 4), 256, 0, st>>>(
                    mode, in1, in2, s2, z, y, out1, out0, B, per_sample, op->task == NLC_OP_INPAINT ? op->idx_c : nullptr,
                    op->task == NLC_OP_INPAINT ? op->n_kept : static_cast<int>(per_sample), sc);
            else
                mask_ddnm_kernel<1><<<blocks_for(total), 256, 0, st>>>(
                    mode, in1, in2, s2, z, y, out1, out0, B, per_sample, op->task == NLC_OP_INPAINT ? op->idx_c : nullptr,
                    op->task == NLC_OP_INPAINT ? op->n_kept : static_cast<int>(per_sample), sc);
            NLC_CHECK_LAUNCH();
            return NLC_OK;
        case NLC_OP_WHCS: {
            NLC_REQUIRE(ws, "DDNM+: WH-CS needs a workspace (nlc_op_ws)");
            const int m = static_cast<int>(plane / op->ratio);
            float* F = ws;
            float* T = ws + total;
            const unsigned g = blocks_for(total);
            if (mode == M_LAMBDA) {
                if ((rc = fwht2d(op, in1, F, Epilogue(), B, st))) return rc;
                whcs_lambda_kernel<<<g, 256, 0, st>>>(F, total, plane, op->idx_a, m, sc);
                NLC_CHECK_LAUNCH();
                return fwht2d(op, F, out1, Epilogue(), B, st);
            }
            const unsigned g4 = blocks_for(total / 4);
            if (mode == M_NOISE) {
                if (vec4) mix_kernel<4><<<g4, 256, 0, st>>>(in1, in2, s2, nullptr, nullptr, T, B, C, plane, op->idx_a, m, nullptr, sc);
                else mix_kernel<1><<<g, 256, 0, st>>>(in1, in2, s2, nullptr, nullptr, T, B, C, plane, op->idx_a, m, nullptr, sc);
                NLC_CHECK_LAUNCH();
                return fwht2d(op, T, out1, Epilogue(), B, st);
            }
            if (vec4) x0_kernel<4><<<g4, 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            else x0_kernel<1><<<g, 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            NLC_CHECK_LAUNCH();
            Epilogue e;
            e.base = out0, e.alpha = sc.a;
            if (sc.plus) {  // x_next = a x0 + FWHT(-a lambda (F - y) + d1 z + d2 et)
                if ((rc = fwht2d(op, out0, F, Epilogue(), B, st))) return rc;
                if (vec4) mix_kernel<4><<<g4, 256, 0, st>>>(z, in2, s2, F, y, T, B, C, plane, op->idx_a, m, nullptr, sc);
                else mix_kernel<1><<<g, 256, 0, st>>>(z, in2, s2, F, y, T, B, C, plane, op->idx_a, m, nullptr, sc);
                NLC_CHECK_LAUNCH();
                e.beta = 1.f;
            } else {        // x_next = a x0 - a FWHT(kept ? F - y : 0) + c1 z + c2 et; the residual rides on the first store
                Epilogue e1;
                e1.invperm = op->idx_a, e1.m = m, e1.ymeas = y;
                if ((rc = fwht2d(op, out0, T, e1, B, st))) return rc;
                e.beta = -sc.a, e.add1 = z, e.g1 = sc.c1, e.add2 = in2, e.g2 = sc.c2, e.add2_stride = s2;
            }
            return fwht2d(op, T, out1, e, B, st);
        }
        case NLC_OP_SEPARABLE: {
            NLC_REQUIRE(ws, "DDNM+: separable operators need a workspace (nlc_op_ws)");
            const int m = op->m, n = B * C, mm = m * m;
            NLC_REQUIRE(mode == M_STEP ? (sc.plus == 0 || op->lam_s) : op->lam_s != nullptr,
                        "DDNM+: this operator class defines no Lambda / Lambda_noise in the reference (SRConv, Deblurring2D)");
            NLC_REQUIRE(!op->lam_s || m == R, "DDNM+: Lambda needs a square separable operator");
            const size_t P = static_cast<size_t>(n) * plane;
            float *W0 = ws, *W1 = ws + P, *W2 = ws + 2 * P, *diff = ws + 3 * P, *ADD = ws + 4 * P;
            float *LAM = ws + 5 * P, *COMB = LAM + 3 * plane;
            const long long pl = plane, smm = static_cast<long long>(mm);
            if (mode == M_LAMBDA) {
                deblur_tables_kernel<<<blocks_for(mm), 256, 0, st>>>(op->lam_s, op->pinv, C, mm, LAM, nullptr, sc);
                NLC_CHECK_LAUNCH();
                // V_s (LAM o (V_s^T X V_s2)) V_s2^T
                if ((rc = launch_gemm(st, n, R, R, R, op->Vs, 0, 1, R, in1, pl, R, 1, W0, nullptr, 1, nullptr, nullptr)) ||
                    (rc = launch_gemm(st, n, R, R, R, W0, pl, R, 1, op->Vs2, 0, R, 1, W1, LAM, 1, nullptr, nullptr)) ||
                    (rc = launch_gemm(st, n, R, R, R, op->Vs, 0, R, 1, W1, pl, R, 1, W2, nullptr, 1, nullptr, nullptr)))
                    return rc;
                return launch_gemm(st, n, R, R, R, W2, pl, R, 1, op->Vs2, 0, 1, R, out1, nullptr, 1, nullptr, nullptr);
            }
            if (mode == M_NOISE) {  // V_s (d1 o v + d2 o e) V_s2^T
                mix_kernel<1><<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, nullptr, nullptr, ADD, B, C, plane, nullptr, 0,
                                                              op->lam_s, sc);
                NLC_CHECK_LAUNCH();
                if ((rc = launch_gemm(st, n, R, R, R, op->Vs, 0, R, 1, ADD, pl, R, 1, W2, nullptr, 1, nullptr, nullptr)))
                    return rc;
                return launch_gemm(st, n, R, R, R, W2, pl, R, 1, op->Vs2, 0, 1, R, out1, nullptr, 1, nullptr, nullptr);
            }
            if (vec4) x0_kernel<4><<<blocks_for(total / 4), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            else x0_kernel<1><<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            NLC_CHECK_LAUNCH();
            if ((rc = separable_A(op, out0, B, diff, ws, y, st))) return rc;  // A x0 - y   (uses W0..W2)
            const float* table = op->pinv;
            const float* add_spec = nullptr;
            if (sc.plus) {
                deblur_tables_kernel<<<blocks_for(mm), 256, 0, st>>>(op->lam_s, op->pinv, C, mm, LAM, COMB, sc);
                NLC_CHECK_LAUNCH();
                mix_kernel<1><<<blocks_for(total), 256, 0, st>>>(z, in2, s2, nullptr, nullptr, ADD, B, C, plane, nullptr, 0,
                                                              op->lam_s, sc);
                NLC_CHECK_LAUNCH();
                table = COMB, add_spec = ADD;
            } else {
                axpby_kernel<<<blocks_for(total), 256, 0, st>>>(z, in2, s2, ADD, per_sample, B, sc.c1, sc.c2);
                NLC_CHECK_LAUNCH();
            }
            // U_s^T diff ; (. U_s2) o table (+ spectral noise) ; V_s . ; . V_s2^T with the x_next assembly
            if ((rc = launch_gemm(st, n, m, m, m, op->Us, 0, 1, m, diff, smm, m, 1, W0, nullptr, 1, nullptr, nullptr)) ||
                (rc = launch_gemm(st, n, m, m, m, W0, smm, m, 1, op->Us2, 0, m, 1, W1, table, C, nullptr, nullptr, 1.f,
                                  -1.f, add_spec)) ||
                (rc = launch_gemm(st, n, R, m, m, op->Vs, 0, R, 1, W1, smm, m, 1, W2, nullptr, 1, nullptr, nullptr)))
                return rc;
            if (sc.plus)
                return launch_gemm(st, n, R, R, m, W2, static_cast<long long>(R) * m, m, 1, op->Vs2, 0, 1, R, out1, nullptr,
                                   1, nullptr, out0, sc.a, 1.f, nullptr);
            return launch_gemm(st, n, R, R, m, W2, static_cast<long long>(R) * m, m, 1, op->Vs2, 0, 1, R, out1, nullptr, 1,
                               nullptr, out0, sc.a, -sc.a, ADD);
        }
        case NLC_OP_BLOCKCS:
        case NLC_OP_GENERAL: {  // no closed form kept for these: x0, the library projection, then the x_next assembly
            NLC_REQUIRE(mode == M_STEP && !sc.plus,
                        "DDNM+: this operator class defines no Lambda / Lambda_noise in the reference (CS, GeneralA)");
            if (vec4) x0_kernel<4><<<blocks_for(total / 4), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            else x0_kernel<1><<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            NLC_CHECK_LAUNCH();
            if ((rc = nlc_op_project(op, out0, y, B, out1, ws, st))) return rc;
            assemble_kernel<<<blocks_for(total), 256, 0, st>>>(out1, z, in2, s2, per_sample, B, sc.a, sc.c1, sc.c2);
            NLC_CHECK_LAUNCH();
            return NLC_OK;
        }
        default:
            return set_error(NLC_EINVAL, "DDNM+: unknown task");
    }
}

static Step make_step(float at, float at_next, double eta, double sigma_y, int plus) {
    Step s;
    s.c_at = sqrtf(1.0f - at), s.d_at = sqrtf(at);
    s.a = sqrtf(at_next);
    const float sigma_t = sqrtf(1.0f - at_next);
    s.coef = make_coef(s.a, sigma_t, sigma_y, eta);
    s.c1 = sigma_t * s.coef.eta, s.c2 = sigma_t * s.coef.root;
    s.plus = plus;
    return s;
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_op_lambda(nlc_op* op, const float* v, int B, const nlc_ddnm_coef* c, float* out, void* ws, void* stream) {
    NLC_REQUIRE(op && v && c && out && B >= 1, "nlc_op_lambda: null argument");
    Step s = make_step(1.f, 1.f, c->eta, c->sigma_y, 1);
    s.coef = make_coef(c->a, c->sigma_t, c->sigma_y, c->eta);
    s.a = c->a;
    return ddnm_apply(op, M_LAMBDA, v, nullptr, 0, nullptr, nullptr, out, nullptr, B, s, static_cast<float*>(ws),
                      static_cast<cudaStream_t>(stream));
}

extern "C" int nlc_op_lambda_noise(nlc_op* op, const float* v, const float* eps, int B, const nlc_ddnm_coef* c, float* out,
                                   void* ws, void* stream) {
    NLC_REQUIRE(op && v && eps && c && out && B >= 1, "nlc_op_lambda_noise: null argument");
    Step s = make_step(1.f, 1.f, c->eta, c->sigma_y, 1);
    s.coef = make_coef(c->a, c->sigma_t, c->sigma_y, c->eta);
    s.a = c->a;
    return ddnm_apply(op, M_NOISE, v, eps, 0, nullptr, nullptr, out, nullptr, B, s, static_cast<float*>(ws),
                      static_cast<cudaStream_t>(stream));
}

extern "C" int nlc_ddnm_step(nlc_op* op, const float* xt, const float* et, int64_t et_stride, const float* z, const float* y,
                             int B, float at, float at_next, double eta, double sigma_y, int plus, float* x0_t,
                             float* x_next, void* ws, void* stream) {
    NLC_REQUIRE(op && xt && et && z && y && x0_t && x_next && B >= 1, "nlc_ddnm_step: null argument");
    NLC_REQUIRE(at > 0.f && at <= 1.f && at_next > 0.f && at_next <= 1.f, "nlc_ddnm_step: alpha_bar out of (0, 1]");
    const long long dim = op->task == NLC_OP_GENERAL ? op->nx : static_cast<long long>(op->C) * op->R * op->R;
    NLC_REQUIRE(et_stride == 0 || et_stride >= dim, "nlc_ddnm_step: et_stride shorter than one image");
    return ddnm_apply(op, M_STEP, xt, et, et_stride, z, y, x_next, x0_t, B, make_step(at, at_next, eta, sigma_y, plus),
                      static_cast<float*>(ws), static_cast<cudaStream_t>(stream));
}

extern "C" int nlc_ddnm_renoise(nlc_ctx* ctx, const float* x0_t, const float* z, int64_t n, float at_next, float* x_next,
                                void* stream) {
    NLC_REQUIRE(ctx && x0_t && z && x_next && n >= 0, "nlc_ddnm_renoise: bad argument");
    if (n == 0) return NLC_OK;
    if (n % 4 == 0 && aligned16(x0_t, z, x_next))
        renoise_kernel<4><<<blocks_for(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            x0_t, z, n, sqrtf(at_next), sqrtf(1.0f - at_next), x_next);
    else
        renoise_kernel<1><<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0_t, z, n, sqrtf(at_next),
                                                                                      sqrtf(1.0f - at_next), x_next);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
