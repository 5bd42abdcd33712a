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

// ---------------------------------------------------------------- colourisation / average-pool SR
// One thread per needle.  in1 = v | xt, in2 = eps | et (sample stride in2_stride), out1 = result | x_next, out0 = x0_t.
template <int K>
__global__ void __launch_bounds__(128) needle_ddnm_kernel(int mode, const float* __restrict__ in1,
                                                           const float* __restrict__ in2, long long in2_stride,
                                                           const float* __restrict__ z, const float* __restrict__ y,
                                                           float* __restrict__ out1, float* __restrict__ out0, int B, int C,
                                                           int R, int r, int per_ch, float u, float s,
                                                           const float* __restrict__ Vfull, const Step sc) {
    constexpr int UN = K <= 16 ? K : 1;  // r = 8 (K = 64) keeps its needles in local memory instead of 256 registers
    __shared__ float V[K * K];
    for (int t = threadIdx.x; t < K * K; t += blockDim.x) V[t] = Vfull[t];
    __syncthreads();
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
    float lam0, d10, d20, lamN, d1N, d2N;
    ddnm_terms(s, sc.coef, lam0, d10, d20);
    ddnm_terms(0.f, sc.coef, lamN, d1N, d2N);

    float n[K];
    if (mode == M_LAMBDA) {
#pragma unroll UN
        for (int k = 0; k < K; ++k) n[k] = in1[base + off(k)];
        float w[K];
#pragma unroll UN
        for (int kp = 0; kp < K; ++kp) {
            float acc = 0.f;
#pragma unroll UN
            for (int k = 0; k < K; ++k) acc = fmaf(V[k * K + kp], n[k], acc);  // V^T n
            w[kp] = kp == 0 ? __fmul_rn(acc, lam0) : acc;
        }
#pragma unroll UN
        for (int j = 0; j < K; ++j) {
            float acc = 0.f;
#pragma unroll UN
            for (int kp = 0; kp < K; ++kp) acc = fmaf(V[j * K + kp], w[kp], acc);
            out1[base + off(j)] = acc;
        }
        return;
    }
    float e[K], zz[K];
    if (mode == M_NOISE) {
#pragma unroll UN
        for (int k = 0; k < K; ++k) zz[k] = in1[base + off(k)], e[k] = in2[base2 + off(k)];
    } else {
        float dot = 0.f;
#pragma unroll UN
        for (int k = 0; k < K; ++k) {
            e[k] = in2[base2 + off(k)];
            n[k] = x0_of(in1[base + off(k)], e[k], sc);
            out0[base + off(k)] = n[k];
            zz[k] = z[base + off(k)];
        }
#pragma unroll UN
        for (int k = 0; k < K; ++k) dot = fmaf(V[k * K], n[k], dot);  // v0 = first column of V
        const float meas = u * (s * dot);                              // A x0
        const float t = (u * (meas - y[i])) * (1.0f / s);              // spectral A^+(A x0 - y)
        if (!sc.plus) {
#pragma unroll UN
            for (int k = 0; k < K; ++k) {
                const float x0h = n[k] - V[k * K] * t;
                out1[base + off(k)] =
                    __fadd_rn(__fadd_rn(__fmul_rn(sc.a, x0h), __fmul_rn(sc.c1, zz[k])), __fmul_rn(sc.c2, e[k]));
            }
            return;
        }
        // resid = v0 t ; Lambda(resid) = V (lambda o V^T resid)
        float w[K];
#pragma unroll UN
        for (int kp = 0; kp < K; ++kp) {
            float acc = 0.f;
#pragma unroll UN
            for (int k = 0; k < K; ++k) acc = fmaf(V[k * K + kp], V[k * K] * t, acc);
            w[kp] = kp == 0 ? __fmul_rn(acc, lam0) : acc;
        }
#pragma unroll UN
        for (int j = 0; j < K; ++j) {
            float acc = 0.f;
#pragma unroll UN
            for (int kp = 0; kp < K; ++kp) acc = fmaf(V[j * K + kp], w[kp], acc);
            n[j] = n[j] - acc;  // x0_hat
        }
    }
    // V (d1 o z) + V (d2 o e): channel / patch entry k stands in for component k (:575-581, 698-699)
#pragma unroll UN
    for (int j = 0; j < K; ++j) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll UN
        for (int k = 0; k < K; ++k) {
            a1 = fmaf(V[j * K + k], __fmul_rn(zz[k], k == 0 ? d10 : d1N), a1);
            a2 = fmaf(V[j * K + k], __fmul_rn(e[k], k == 0 ? d20 : d2N), a2);
        }
        const float nz = __fadd_rn(a1, a2);
        out1[base + off(j)] = mode == M_NOISE ? nz : __fadd_rn(__fmul_rn(sc.a, n[j]), nz);
    }
}

// ---------------------------------------------------------------- inpainting / denoising (V is a permutation / identity)
// pos2k == nullptr: Denoising (every entry kept, s = 1, and the class' own scalar rules :464-476)
__global__ void __launch_bounds__(256) mask_ddnm_kernel(int mode, const float* __restrict__ in1,
                                                         const float* __restrict__ in2, long long in2_stride,
                                                         const float* __restrict__ z, const float* __restrict__ y,
                                                         float* __restrict__ out1, float* __restrict__ out0, int B, int C,
                                                         int HW, const int* __restrict__ pos2k, int n_kept, const Step sc) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_sample = static_cast<long long>(C) * HW;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    const size_t i2 = static_cast<size_t>(b) * (in2_stride ? in2_stride : per_sample) + q;
    int k;
    float lam, d1, d2;
    bool scale_first = false;  // Denoising's `vec * sigma_t * eta` multiplies in that order
    if (pos2k) {
        const int p = static_cast<int>(q % HW), c = static_cast<int>(q / HW);
        k = pos2k[p * C + c];
        ddnm_terms(k >= 0 ? 1.f : 0.f, sc.coef, lam, d1, d2);
    } else {
        k = static_cast<int>(q);
        const Coef& c = sc.coef;
        const float thr = __fmul_rn(c.a, c.sy);
        lam = c.st < thr ? __fdiv_rn(__fdiv_rn(__fmul_rn(c.st, c.root), c.a), c.sy) : 1.f;
        d2 = 0.f;
        if (c.st >= thr) d1 = __fsqrt_rn(__fsub_rn(__fmul_rn(c.st, c.st), __fmul_rn(__fmul_rn(c.a, c.a), c.sy2)));
        else d1 = 0.f, scale_first = true;
    }
    auto noise = [&](float zv, float ev) -> float {
        if (scale_first) return __fmul_rn(__fmul_rn(zv, sc.coef.st), sc.coef.eta);
        if (!pos2k) return __fmul_rn(zv, d1);
        return __fadd_rn(__fmul_rn(zv, d1), __fmul_rn(ev, d2));
    };
    if (mode == M_LAMBDA) {
        out1[i] = (pos2k || lam != 1.f) ? __fmul_rn(in1[i], lam) : in1[i];
    } else if (mode == M_NOISE) {
        out1[i] = noise(in1[i], in2[i2]);
    } else {
        const float ev = in2[i2];
        const float x0 = x0_of(in1[i], ev, sc);
        out0[i] = x0;
        const float resid = k >= 0 ? __fsub_rn(x0, y[static_cast<size_t>(b) * n_kept + k]) : 0.f;
        if (sc.plus) {
            const float x0h = __fsub_rn(x0, (pos2k || lam != 1.f) ? __fmul_rn(resid, lam) : resid);
            out1[i] = __fadd_rn(__fmul_rn(sc.a, x0h), noise(z[i], ev));
        } else {
            const float x0h = __fsub_rn(x0, resid);
            out1[i] = __fadd_rn(__fadd_rn(__fmul_rn(sc.a, x0h), __fmul_rn(sc.c1, z[i])), __fmul_rn(sc.c2, ev));
        }
    }
}

// ---------------------------------------------------------------- shared elementwise pieces (WH-CS, Deblurring)
__global__ void __launch_bounds__(256) x0_kernel(const float* __restrict__ xt, const float* __restrict__ et,
                                                  long long et_stride, float* __restrict__ x0, long long per_sample, int B,
                                                  const Step sc) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample;
    x0[i] = x0_of(xt[i], et[static_cast<size_t>(b) * et_stride + (i - b * per_sample)], sc);
}
// out = g1[o] v + g2[o] e with the factors of spectral position o = i % plane:
//   invperm != nullptr (WH-CS): kept (invperm[o] < m) -> s = 1, else null space;  tab != nullptr (Deblurring): s = tab[o]
// F != nullptr adds the Lambda'd residual of the step:  -a lambda (F - y) on kept entries (WH-CS only)
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ v, const float* __restrict__ e,
                                                   long long e_stride, const float* __restrict__ F,
                                                   const float* __restrict__ y, float* __restrict__ out, int B, int C,
                                                   long long plane, const int* __restrict__ invperm, int m,
                                                   const float* __restrict__ tab, const Step sc) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_sample = static_cast<long long>(C) * plane;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    const int o = static_cast<int>(q % plane), c = static_cast<int>(q / plane);
    float lam, d1, d2;
    int j = -1;
    if (invperm) {
        j = invperm[o];
        ddnm_terms(j < m ? 1.f : 0.f, sc.coef, lam, d1, d2);
    } else {
        ddnm_terms(tab[o], sc.coef, lam, d1, d2);
    }
    float r = __fadd_rn(__fmul_rn(v[i], d1), __fmul_rn(e[static_cast<size_t>(b) * e_stride + q], d2));
    if (F && j >= 0 && j < m) {
        const float resid = F[i] - y[static_cast<size_t>(b) * m * C + static_cast<size_t>(j) * C + c];
        r = fmaf(-sc.a * lam, resid, r);
    }
    out[i] = r;
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
// T[o] = kept ? F[o] - y : 0  (the spectral residual of the plain DDNM step)
__global__ void __launch_bounds__(256) whcs_resid_kernel(const float* __restrict__ F, const float* __restrict__ y,
                                                          float* __restrict__ T, int B, int C, long long plane,
                                                          const int* __restrict__ invperm, int m) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per_sample = static_cast<long long>(C) * plane;
    if (i >= per_sample * B) return;
    const long long b = i / per_sample, q = i - b * per_sample;
    const int o = static_cast<int>(q % plane), c = static_cast<int>(q / plane);
    const int j = invperm[o];
    T[i] = j < m ? F[i] - y[static_cast<size_t>(b) * m * C + static_cast<size_t>(j) * C + c] : 0.f;
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
__global__ void __launch_bounds__(256) renoise_kernel(const float* __restrict__ x0, const float* __restrict__ z,
                                                       long long n, float a, float sig, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(__fmul_rn(a, x0[i]), __fmul_rn(z[i], sig));
}

template <int K>
static int launch_needle(nlc_op* op, int mode, const float* in1, const float* in2, long long s2, const float* z,
                         const float* y, float* out1, float* out0, int B, const Step& sc, cudaStream_t st) {
    const int per_ch = op->task == NLC_OP_SR_AVG, r = per_ch ? op->ratio : 1;
    const long long plane = static_cast<long long>(op->R) * op->R;
    const long long n = per_ch ? static_cast<long long>(B) * op->C * (plane / (r * r)) : static_cast<long long>(B) * plane;
    needle_ddnm_kernel<K><<<blocks_for(n, 128), 128, 0, st>>>(mode, in1, in2, s2, z, y, out1, out0, B, op->C, op->R, r,
                                                               per_ch, op->u, op->s, op->Vfull, sc);
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
            mask_ddnm_kernel<<<blocks_for(total), 256, 0, st>>>(
                mode, in1, in2, s2, z, y, out1, out0, B, C, static_cast<int>(plane),
                op->task == NLC_OP_INPAINT ? op->idx_b : nullptr,
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
            if (mode == M_NOISE) {
                mix_kernel<<<g, 256, 0, st>>>(in1, in2, s2, nullptr, nullptr, T, B, C, plane, op->idx_a, m, nullptr, sc);
                NLC_CHECK_LAUNCH();
                return fwht2d(op, T, out1, Epilogue(), B, st);
            }
            x0_kernel<<<g, 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            NLC_CHECK_LAUNCH();
            if ((rc = fwht2d(op, out0, F, Epilogue(), B, st))) return rc;
            Epilogue e;
            e.base = out0, e.alpha = sc.a;
            if (sc.plus) {  // x_next = a x0 + FWHT(-a lambda (F - y) + d1 z + d2 et)
                mix_kernel<<<g, 256, 0, st>>>(z, in2, s2, F, y, T, B, C, plane, op->idx_a, m, nullptr, sc);
                e.beta = 1.f;
            } else {        // x_next = a x0 - a FWHT(F - y) + c1 z + c2 et
                whcs_resid_kernel<<<g, 256, 0, st>>>(F, y, T, B, C, plane, op->idx_a, m);
                e.beta = -sc.a, e.add1 = z, e.g1 = sc.c1, e.add2 = in2, e.g2 = sc.c2, e.add2_stride = s2;
            }
            NLC_CHECK_LAUNCH();
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
                mix_kernel<<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, nullptr, nullptr, ADD, B, C, plane, nullptr, 0,
                                                              op->lam_s, sc);
                NLC_CHECK_LAUNCH();
                if ((rc = launch_gemm(st, n, R, R, R, op->Vs, 0, R, 1, ADD, pl, R, 1, W2, nullptr, 1, nullptr, nullptr)))
                    return rc;
                return launch_gemm(st, n, R, R, R, W2, pl, R, 1, op->Vs2, 0, 1, R, out1, nullptr, 1, nullptr, nullptr);
            }
            x0_kernel<<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
            NLC_CHECK_LAUNCH();
            if ((rc = separable_A(op, out0, B, diff, ws, y, st))) return rc;  // A x0 - y   (uses W0..W2)
            const float* table = op->pinv;
            const float* add_spec = nullptr;
            if (sc.plus) {
                deblur_tables_kernel<<<blocks_for(mm), 256, 0, st>>>(op->lam_s, op->pinv, C, mm, LAM, COMB, sc);
                NLC_CHECK_LAUNCH();
                mix_kernel<<<blocks_for(total), 256, 0, st>>>(z, in2, s2, nullptr, nullptr, ADD, B, C, plane, nullptr, 0,
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
            x0_kernel<<<blocks_for(total), 256, 0, st>>>(in1, in2, s2, out0, per_sample, B, sc);
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
    renoise_kernel<<<blocks_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0_t, z, n, sqrtf(at_next),
                                                                               sqrtf(1.0f - at_next), x_next);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
