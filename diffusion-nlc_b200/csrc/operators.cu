// DDNM constraint operators: H = A, H^T = At, H^+ = A_pinv of functions/svd_operators.py and
// the fused projection  x0_hat = x0 - A^+(A x0 - y)  of image_sample.py:376-379, in closed form on the NCHW image.
//
// The reference applies U, Sigma, V^T as chains of clone / reshape / permute / index_put / matmul launches on
// flattened [B, 3R^2] rows; the permutations between the "spectral" ordering and the image cancel inside A, At and
// A^+, which leaves per task:
//   Inpainting  (:324-359)  gather / scatter of the kept (pixel, channel) entries
//   Colorization(:627-667)  per-pixel dot with v0 = V_small[:,0] and its transpose
//   SuperRes    (:479-533)  per r x r patch dot with v0 = V_small[:,0] and its transpose
//   WalshHadamardCS (:211-251)  orthonormal 2-D fast Walsh-Hadamard transform (same butterfly order as the
//                           reference's 1-D FWHT over R^2 entries) + a permutation gather
//   SRConv (:851-931), Deblurring (:934-1014)  separable:  U_s (M o (V_s^T X V_s)) U_s^T  with a spectral
//                           multiplier table M built by the host from the reference's own perm / singulars
//   CS (:101-160)           block-wise compressed sensing: every 32 x 32 patch times the first cs columns of a 1024 x 1024
//                           orthogonal V_small: patchify -> one [patches, 1024] x [1024, cs] fp32 GEMM (-> unpatchify)
//   GeneralA (:173-208)     dense U, s, V of an arbitrary small A: two fp32 GEMMs around a per-column scale
//   Denoising (:442-476)    A = I
// All kernels are HBM-bound fp32 (the separable pair adds eight small fp32 GEMMs per image-channel).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "operators.h"
#include "ptx.cuh"

namespace nlc {

// ---------------------------------------------------------------- colourisation / avg-pool SR
// mode 0: A, 1: At, 2: A_pinv, 3: project, 4: A_pinv_eta (s := s / (s^2 + eta), functions/svd_operators.py:82-91).
// One thread per (sample, low-res pixel); P = patch edge (1 for colour).
template <int MODE>
__global__ void __launch_bounds__(256) needle_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      float* __restrict__ out, int B, int C, int R, int r, int per_ch,
                                                      float u, float s, const float* __restrict__ v0) {
    // per_ch = 1: avg-pool SR (one measurement per channel and patch, v0 over the r*r patch)
    // per_ch = 0: colourisation (one measurement per pixel, v0 over the 3 channels, r == 1)
    const int yd = R / r;
    const long long n_meas = per_ch ? static_cast<long long>(B) * C * yd * yd : static_cast<long long>(B) * R * R;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_meas) return;
    const int K = per_ch ? r * r : C;
    // element k of the needle lives at base + off(k)
    size_t base;
    if (per_ch) {
        const int j = static_cast<int>(i % yd), ii = static_cast<int>((i / yd) % yd);
        const long long bc = i / (static_cast<long long>(yd) * yd);
        base = (static_cast<size_t>(bc) * R + static_cast<size_t>(ii) * r) * R + static_cast<size_t>(j) * r;
    } else {
        const long long b = i / (static_cast<long long>(R) * R), p = i - b * R * R;
        base = static_cast<size_t>(b) * C * R * R + p;
    }
    auto off = [&](int k) -> size_t {
        return per_ch ? static_cast<size_t>(k / r) * R + (k % r) : static_cast<size_t>(k) * R * R;
    };
    float meas = 0.f;
    if (MODE == 0 || MODE == 3) {
        float dot = 0.f;
        for (int k = 0; k < K; ++k) dot = fmaf(v0[k], x[base + off(k)], dot);
        meas = u * (s * dot);  // U * (singulars * Vt(x))
        if (MODE == 0) {
            out[i] = meas;
            return;
        }
    }
    float t;
    if (MODE == 1) t = s * (u * y[i]);                       // V(add_zeros(singulars * Ut(y)))
    else if (MODE == 2) t = (u * y[i]) * (1.0f / s);         // V(add_zeros(Ut(y) * 1/singulars))
    else if (MODE == 4) t = (u * y[i]) * s;                  // the caller passes s / (s^2 + eta)
    else t = (u * (meas - y[i])) * (1.0f / s);               // A^+(A x0 - y)
    for (int k = 0; k < K; ++k) {
        const float v = v0[k] * t;
        out[base + off(k)] = MODE == 3 ? x[base + off(k)] - v : v;
    }
}

// ---------------------------------------------------------------- inpainting
__global__ void inpaint_A_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, int HW,
                                 const int* __restrict__ kept, int n_kept) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * n_kept) return;
    const int b = static_cast<int>(i / n_kept), k = static_cast<int>(i - static_cast<long long>(b) * n_kept);
    const int j = kept[k], p = j / C, c = j - p * C;
    y[i] = x[(static_cast<size_t>(b) * C + c) * HW + p];
}
// mode 1/2: scatter (At == A^+, all singulars are 1); mode 3: project
template <int MODE>
__global__ void inpaint_back_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                    int B, int C, int HW, const int* __restrict__ pos2k, int n_kept, float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * C * HW) return;
    const int p = static_cast<int>(i % HW);
    const int c = static_cast<int>((i / HW) % C);
    const int b = static_cast<int>(i / (static_cast<long long>(HW) * C));
    const int k = pos2k[p * C + c];
    if (MODE == 3) {
        const float xv = x[i];
        out[i] = k >= 0 ? xv - (xv - y[static_cast<size_t>(b) * n_kept + k]) : xv;
    } else {
        out[i] = k >= 0 ? __fmul_rn(y[static_cast<size_t>(b) * n_kept + k], f) : 0.f;  // f = 1 (At, A^+), 1/(1+eta)
    }
}

// ---- 128-bit versions of the kernels above for the shapes the sampler uses (R a multiple of 4): same arithmetic per
// element, four pixels (colour, inpainting) or one r x r patch in row vectors (SR, r = 2 / 4) per thread
__device__ __forceinline__ float get4(const float4& v, int e) { return e == 0 ? v.x : (e == 1 ? v.y : (e == 2 ? v.z : v.w)); }
__device__ __forceinline__ void set4(float4& v, int e, float f) {
    if (e == 0) v.x = f; else if (e == 1) v.y = f; else if (e == 2) v.z = f; else v.w = f;
}
template <int MODE>
__global__ void __launch_bounds__(256) color_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ y,
                                                         float4* __restrict__ out, long long n4, long long plane4, float u,
                                                         float s, const float* __restrict__ v0) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long b = i / plane4, p4 = i - b * plane4;
    const float v[3] = {v0[0], v0[1], v0[2]};
    const size_t base = static_cast<size_t>(b) * 3 * plane4 + p4;
    float4 xs[3], meas = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 0 || MODE == 3) {
#pragma unroll
        for (int c = 0; c < 3; ++c) xs[c] = __ldg(x + base + c * plane4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) dot = fmaf(v[c], get4(xs[c], e), dot);
            set4(meas, e, u * (s * dot));
        }
        if (MODE == 0) {
            out[i] = meas;
            return;
        }
    }
    const float4 yv = __ldg(y + i);
    float4 t;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float ye = get4(yv, e);
        set4(t, e, MODE == 1 ? s * (u * ye) : (MODE == 2 ? (u * ye) * (1.0f / s)
                                                        : (MODE == 4 ? (u * ye) * s : (u * (get4(meas, e) - ye)) * (1.0f / s))));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float4 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float w = v[c] * get4(t, e);
            set4(o, e, MODE == 3 ? get4(xs[c], e) - w : w);
        }
        out[base + c * plane4] = o;
    }
}
template <int RW> struct RowVec;
template <> struct RowVec<2> { typedef float2 type; };
template <> struct RowVec<4> { typedef float4 type; };
template <int MODE, int RW>
__global__ void __launch_bounds__(256) sr_vec_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      float* __restrict__ out, long long n_meas, int R, float u, float s,
                                                      const float* __restrict__ v0) {
    typedef typename RowVec<RW>::type VT;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_meas) return;
    const int yd = R / RW;
    const int j = static_cast<int>(i % yd), ii = static_cast<int>((i / yd) % yd);
    const long long bc = i / (static_cast<long long>(yd) * yd);
    const size_t base = (static_cast<size_t>(bc) * R + static_cast<size_t>(ii) * RW) * R + static_cast<size_t>(j) * RW;
    float xs[RW * RW];
    float meas = 0.f;
    if (MODE == 0 || MODE == 3) {
#pragma unroll
        for (int rr = 0; rr < RW; ++rr) {
            const VT row = __ldg(reinterpret_cast<const VT*>(x + base + static_cast<size_t>(rr) * R));
            const float* rp = reinterpret_cast<const float*>(&row);
#pragma unroll
            for (int cc = 0; cc < RW; ++cc) xs[rr * RW + cc] = rp[cc];
        }
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < RW * RW; ++k) dot = fmaf(v0[k], xs[k], dot);
        meas = u * (s * dot);
        if (MODE == 0) {
            out[i] = meas;
            return;
        }
    }
    float t;
    if (MODE == 1) t = s * (u * y[i]);
    else if (MODE == 2) t = (u * y[i]) * (1.0f / s);
    else if (MODE == 4) t = (u * y[i]) * s;
    else t = (u * (meas - y[i])) * (1.0f / s);
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
        VT row;
        float* rp = reinterpret_cast<float*>(&row);
#pragma unroll
        for (int cc = 0; cc < RW; ++cc) {
            const float w = v0[rr * RW + cc] * t;
            rp[cc] = MODE == 3 ? xs[rr * RW + cc] - w : w;
        }
        *reinterpret_cast<VT*>(out + base + static_cast<size_t>(rr) * R) = row;
    }
}
// inpainting, four consecutive entries of a plane per thread; pos2k in image order (idx_c)
template <int MODE>
__global__ void __launch_bounds__(256) inpaint_back_vec_kernel(const float4* __restrict__ x, const float* __restrict__ y,
                                                                float4* __restrict__ out, long long n4, long long sample4,
                                                                const int4* __restrict__ pos2k_planar, int n_kept,
                                                                float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long b = i / sample4, q4 = i - b * sample4;
    const int4 k4 = __ldg(pos2k_planar + q4);
    const int ks[4] = {k4.x, k4.y, k4.z, k4.w};
    const float* yb = y + static_cast<size_t>(b) * n_kept;
    float4 o, xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == 3) xv = __ldg(x + i);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float xe = get4(xv, e);
        if (MODE == 3) set4(o, e, ks[e] >= 0 ? xe - (xe - yb[ks[e]]) : xe);
        else set4(o, e, ks[e] >= 0 ? __fmul_rn(yb[ks[e]], f) : 0.f);
    }
    out[i] = o;
}
__global__ void __launch_bounds__(256) identity_vec_kernel(const float4* __restrict__ in, const float4* __restrict__ y,
                                                            float4* __restrict__ out, long long n4, int mode, float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(in + i);
    float4 o = v;
    if (mode == 3) {
        const float4 w = __ldg(y + i);
        o = make_float4(__fsub_rn(v.x, __fsub_rn(v.x, w.x)), __fsub_rn(v.y, __fsub_rn(v.y, w.y)),
                        __fsub_rn(v.z, __fsub_rn(v.z, w.z)), __fsub_rn(v.w, __fsub_rn(v.w, w.w)));
    } else if (mode == 4) {
        o = make_float4(__fmul_rn(v.x, f), __fmul_rn(v.y, f), __fmul_rn(v.z, f), __fmul_rn(v.w, f));
    }
    out[i] = o;
}

// ---------------------------------------------------------------- Walsh-Hadamard
// rows: one warp per image row, E = R/32 contiguous entries per lane (128-bit loads); butterflies h < E in registers,
// h >= E across lanes by xor-shuffle — ascending h, the reference's stage order (its first log2(R) stages)
template <int E>
__global__ void __launch_bounds__(256) fwht_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         long long n_rows) {
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const size_t off = static_cast<size_t>(row) * (32 * E) + lane * E;
    float v[E];
    if constexpr (E >= 4) {
#pragma unroll
        for (int i = 0; i < E; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(in + off + i);
            v[i] = t.x, v[i + 1] = t.y, v[i + 2] = t.z, v[i + 3] = t.w;
        }
    } else if constexpr (E == 2) {
        const float2 t = *reinterpret_cast<const float2*>(in + off);
        v[0] = t.x, v[1] = t.y;
    } else {
        v[0] = in[off];
    }
#pragma unroll
    for (int h = 1; h < E; h <<= 1)
#pragma unroll
        for (int i = 0; i < E; ++i)
            if (!(i & h)) {
                const float a = v[i], b = v[i + h];
                v[i] = a + b, v[i + h] = a - b;
            }
#pragma unroll
    for (int msk = 1; msk < 32; msk <<= 1) {
        const bool upper = lane & msk;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const float p = __shfl_xor_sync(0xffffffffu, v[i], msk);
            v[i] = upper ? p - v[i] : v[i] + p;
        }
    }
    if constexpr (E >= 4) {
#pragma unroll
        for (int i = 0; i < E; i += 4)
            *reinterpret_cast<float4*>(out + off + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else if constexpr (E == 2) {
        *reinterpret_cast<float2*>(out + off) = make_float2(v[0], v[1]);
    } else {
        out[off] = v[0];
    }
}
// columns: CTA = 32 columns x R rows of one plane staged in shared memory ([R][33]: conflict-free both ways); a warp takes
// one column at a time, lane l holding rows l, l+32, ...: stages h < 32 (rows) by xor-shuffle, h >= 32 in registers —
// ascending h again (stages R .. R^2/2 of the flattened transform) — then / R.  The epilogue (operators.h) folds the
// WH-CS gather / residual, the projection's final subtraction and the DDNM step's x_{t-1} assembly into the store.
template <int E>
__global__ void __launch_bounds__(256) fwht_cols_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         const Epilogue e, int C) {
    constexpr int R = 32 * E;
    extern __shared__ float tile[];  // [R][33]
    const int plane = blockIdx.y, c0 = blockIdx.x * 32;
    const size_t pbase = static_cast<size_t>(plane) * R * R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += 8) tile[r * 33 + lane] = in[pbase + static_cast<size_t>(r) * R + c0 + lane];
    __syncthreads();
    for (int col = warp * 4; col < warp * 4 + 4; ++col) {
        float v[E];
#pragma unroll
        for (int i = 0; i < E; ++i) v[i] = tile[(lane + 32 * i) * 33 + col];
#pragma unroll
        for (int msk = 1; msk < 32; msk <<= 1) {
            const bool upper = lane & msk;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const float p = __shfl_xor_sync(0xffffffffu, v[i], msk);
                v[i] = upper ? p - v[i] : v[i] + p;
            }
        }
#pragma unroll
        for (int h = 1; h < E; h <<= 1)
#pragma unroll
            for (int i = 0; i < E; ++i)
                if (!(i & h)) {
                    const float a = v[i], b = v[i + h];
                    v[i] = a + b, v[i + h] = a - b;
                }
#pragma unroll
        for (int i = 0; i < E; ++i) tile[(lane + 32 * i) * 33 + col] = v[i];
    }
    __syncthreads();
    const float fr = static_cast<float>(R);
    const int b = plane / C, c = plane % C;
    for (int r = warp; r < R; r += 8) {
        const size_t o = pbase + static_cast<size_t>(r) * R + c0 + lane;
        float v = tile[r * 33 + lane] / fr;
        if (fwht_epilogue(e, v, b, c, C, R, r, c0 + lane, o)) out[o] = v;
    }
}
// temp[b][c][q] = invperm[q] < m ? f * y[b][invperm[q]*C + c] : 0: the measurement scattered back to its spectral positions
// (A^T, A^+ with f = 1, A_pinv_eta with f = 1 / (1 + eta)); the forward gather and the projection's residual ride on the
// transform's epilogue instead (fwht_epilogue)
__global__ void whcs_scatter_kernel(const float* __restrict__ y, float* __restrict__ temp, int B, int C, int N, int m,
                                    const int* __restrict__ invperm, float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * C * N) return;
    const int q = static_cast<int>(i % N);
    const int c = static_cast<int>((i / N) % C);
    const int b = static_cast<int>(i / (static_cast<long long>(N) * C));
    const int j = invperm[q];
    temp[i] = j < m ? __fmul_rn(y[static_cast<size_t>(b) * m * C + static_cast<size_t>(j) * C + c], f) : 0.f;
}

// ---------------------------------------------------------------- strided batched fp32 GEMM (separable operators)
// Cm[b][i][j] = epi( sum_k A[b*sab + i*sai + k*sak] * Bm[b*sbb + k*sbk + j*sbj] )
// epi: * mult[(b % nch)*M*N + i*N + j] (if mult) ; - sub[b][i][j] (if sub) ; base[b][i][j] - value (if base)
struct GemmArgs {
    const float *A, *Bm, *mult, *sub, *base, *add;
    float alpha, beta;
    float* Cm;
    long long sab, sai, sak, sbb, sbk, sbj;
    int M, N, K, nch;
};
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const GemmArgs g) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    const float* A = g.A + static_cast<long long>(b) * g.sab;
    const float* Bm = g.Bm + static_cast<long long>(b) * g.sbb;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < g.K; k0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            // pick the faster-varying index per operand so that global reads coalesce along unit strides
            int ii, kk;
            if (g.sak == 1) { kk = t & 15; ii = t >> 4; } else { ii = t & 63; kk = t >> 6; }
            float v = 0.f;
            if (i0 + ii < g.M && k0 + kk < g.K) v = A[(i0 + ii) * g.sai + (k0 + kk) * g.sak];
            As[kk][ii] = v;
            int jj, k2;
            if (g.sbk == 1) { k2 = t & 15; jj = t >> 4; } else { jj = t & 63; k2 = t >> 6; }
            v = 0.f;
            if (j0 + jj < g.N && k0 + k2 < g.K) v = Bm[(k0 + k2) * g.sbk + (j0 + jj) * g.sbj];
            Bs[k2][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = As[kk][ty + 16 * r], bb[r] = Bs[kk][tx + 16 * r];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
        }
        __syncthreads();
    }
    const size_t cb = static_cast<size_t>(b) * g.M * g.N;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 16 * r;
        if (i >= g.M) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx + 16 * c;
            if (j >= g.N) continue;
            const size_t o = static_cast<size_t>(i) * g.N + j;
            float v = acc[r][c];
            if (g.mult) v *= g.mult[static_cast<size_t>(b % g.nch) * g.M * g.N + o];
            if (g.sub) v -= g.sub[cb + o];
            if (g.base) v = g.alpha * g.base[cb + o] + g.beta * v;  // projection: alpha 1, beta -1
            if (g.add) v += g.add[cb + o];
            g.Cm[cb + o] = v;
        }
    }
}

// The same GEMM for the sizes the 256 x 256 operators produce (M, N >= 128): 128 x 128 x 16 tiles, 8 x 8 outputs per thread
// read from shared memory as float4, the next K slab prefetched into registers while the current one is multiplied.
__device__ __forceinline__ void gemm_epilogue(const GemmArgs& g, float v, int b, size_t cb, int i, int j) {
    const size_t o = static_cast<size_t>(i) * g.N + j;
    if (g.mult) v *= g.mult[static_cast<size_t>(b % g.nch) * g.M * g.N + o];
    if (g.sub) v -= g.sub[cb + o];
    if (g.base) v = g.alpha * g.base[cb + o] + g.beta * v;
    if (g.add) v += g.add[cb + o];
    g.Cm[cb + o] = v;
}
__global__ void __launch_bounds__(256, 2) sgemm_strided128_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[16][128 + 4];  // +4: the k-major fill of a unit-K operand would hit one bank 16 times
    __shared__ __align__(16) float Bs[16][128 + 4];
    const int b = blockIdx.z, i0 = blockIdx.y * 128, j0 = blockIdx.x * 128;
    const float* A = g.A + static_cast<long long>(b) * g.sab;
    const float* Bm = g.Bm + static_cast<long long>(b) * g.sbb;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    // this thread's 8 + 8 elements of a slab, element q at (i, k) = (ai0 + q*ais, ak0 + q*aks): the unit-stride index varies
    // fastest across threads (t = tid + 256 q; unit-K operand: k = t & 15, i = t >> 4; otherwise i = t & 127, k = t >> 7)
    const bool akf = g.sak == 1, bkf = g.sbk == 1;
    const int tid = threadIdx.x;
    const int ai0 = akf ? tid >> 4 : tid & 127, ais = akf ? 16 : 0, ak0 = akf ? tid & 15 : tid >> 7, aks = akf ? 0 : 2;
    const int bj0 = bkf ? tid >> 4 : tid & 127, bjs = bkf ? 16 : 0, bk0 = bkf ? tid & 15 : tid >> 7, bks = bkf ? 0 : 2;
    const float* Ap = A + static_cast<long long>(i0 + ai0) * g.sai + static_cast<long long>(ak0) * g.sak;
    const float* Bp = Bm + static_cast<long long>(bk0) * g.sbk + static_cast<long long>(j0 + bj0) * g.sbj;
    const long long aq = ais * g.sai + aks * g.sak, bq = bks * g.sbk + bjs * g.sbj;  // element q -> q + 1
    float ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            ra[q] = (i0 + ai0 + q * ais < g.M && k0 + ak0 + q * aks < g.K) ? Ap[k0 * g.sak + q * aq] : 0.f;
            rb[q] = (j0 + bj0 + q * bjs < g.N && k0 + bk0 + q * bks < g.K) ? Bp[k0 * g.sbk + q * bq] : 0.f;
        }
    };
    float acc[8][8] = {};
    fetch(0);
    for (int k0 = 0; k0 < g.K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 8; ++q) As[ak0 + q * aks][ai0 + q * ais] = ra[q], Bs[bk0 + q * bks][bj0 + q * bjs] = rb[q];
        __syncthreads();
        if (k0 + 16 < g.K) fetch(k0 + 16);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], bb[c], acc[r][c]);
        }
        __syncthreads();
    }
    const size_t cb = static_cast<size_t>(b) * g.M * g.N;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + r - 4);
        if (i >= g.M) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = j0 + (c < 4 ? tx * 4 + c : 64 + tx * 4 + c - 4);
            if (j < g.N) gemm_epilogue(g, acc[r][c], b, cb, i, j);
        }
    }
}

__global__ void l1_diff_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                    float* __restrict__ out) {
    __shared__ float red[32];
    const float* ar = a + static_cast<size_t>(blockIdx.x) * n;
    const float* br = b + static_cast<size_t>(blockIdx.x) * n;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += fabsf(ar[i] - br[i]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) out[blockIdx.x] = acc;
    }
}

// Per-sample restoration metrics of a finished batch (image_sample.py:671-679): s = clamp((x+1)/2, 0, 1) is the image
// that gets written out; mse = mean((s - orig)^2) (PSNR = 10 log10(1/mse) is taken on the host), l1 = ||(2s-1) - (2 orig
// - 1)||_1 (`cons_orig`).  One CTA per sample, one pass; `s_out` (nullable) receives s.
__global__ void image_metrics_kernel(const float* __restrict__ x, const float* __restrict__ orig, long long n,
                                     float* __restrict__ s_out, float* __restrict__ mse, float* __restrict__ l1) {
    __shared__ float red[2][32];
    const size_t base = static_cast<size_t>(blockIdx.x) * n;
    float a2 = 0.f, a1 = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float s = fminf(fmaxf(__fmul_rn(__fadd_rn(x[base + i], 1.0f), 0.5f), 0.0f), 1.0f);
        const float o = orig[base + i];
        const float d = s - o;
        a2 = fmaf(d, d, a2);
        a1 += fabsf(__fadd_rn(__fmul_rn(2.0f, s), -1.0f) - __fadd_rn(__fmul_rn(2.0f, o), -1.0f));
        if (s_out) s_out[base + i] = s;
    }
    a2 = warp_sum(a2), a1 = warp_sum(a1);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = a2, red[1][threadIdx.x >> 5] = a1;
    __syncthreads();
    if (threadIdx.x < 32) {
        a2 = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.f;
        a1 = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.f;
        a2 = warp_sum(a2), a1 = warp_sum(a1);
        if (threadIdx.x == 0) mse[blockIdx.x] = a2 / static_cast<float>(n), l1[blockIdx.x] = a1;
    }
}

int launch_gemm(cudaStream_t st, int batch, int M, int N, int K, const float* A, long long sab, long long sai,
                long long sak, const float* Bm, long long sbb, long long sbk, long long sbj, float* Cm, const float* mult,
                int nch, const float* sub, const float* base, float alpha, float beta, const float* add) {
    GemmArgs g{A, Bm, mult, sub, base, add, alpha, beta, Cm, sab, sai, sak, sbb, sbk, sbj, M, N, K, nch};
    if (M >= 128 && N >= 128) {
        sgemm_strided128_kernel<<<dim3((N + 127) / 128, (M + 127) / 128, batch), 256, 0, st>>>(g);
    } else {
        sgemm_strided_kernel<<<dim3((N + 63) / 64, (M + 63) / 64, batch), 256, 0, st>>>(g);
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

// separable forward:  out[b,c] = U_s ( mult_c o (V_s[:, :m]^T X V2_s[:, :m]) ) U2_s^T  (- sub);  U2 = U, V2 = V unless the
// operator blurs rows and columns with different kernels
int separable_A(nlc_op* op, const float* x, int B, float* y, float* ws, const float* sub, cudaStream_t st) {
    const int R = op->R, m = op->m, n = B * op->C;
    float* T1 = ws;                                        // [n][m][R]
    float* T2 = T1 + static_cast<size_t>(n) * m * R;       // [n][m][m]
    float* T3 = T2 + static_cast<size_t>(n) * m * m;       // [n][m][m]
    int rc;
    // T1 = V_s[:, :m]^T X          A(i,k) = Vs[k*R + i]
    if ((rc = launch_gemm(st, n, m, R, R, op->Vs, 0, 1, R, x, static_cast<long long>(R) * R, R, 1, T1, nullptr, 1,
                          nullptr, nullptr))) return rc;
    // T2 = (T1 V_s[:, :m]) o mult  B(k,j) = Vs[k*R + j]
    if ((rc = launch_gemm(st, n, m, m, R, T1, static_cast<long long>(m) * R, R, 1, op->Vs2, 0, R, 1, T2, op->mult, op->C,
                          nullptr, nullptr))) return rc;
    // T3 = U_s T2
    if ((rc = launch_gemm(st, n, m, m, m, op->Us, 0, m, 1, T2, static_cast<long long>(m) * m, m, 1, T3, nullptr, 1,
                          nullptr, nullptr))) return rc;
    // Y = T3 U_s^T (- sub)         B(k,j) = Us[j*m + k]
    return launch_gemm(st, n, m, m, m, T3, static_cast<long long>(m) * m, m, 1, op->Us2, 0, 1, m, y, nullptr, 1, sub,
                       nullptr);
}
// separable backward: out[b,c] = V_s[:, :m] ( table_c o (U_s^T Y U_s) ) V_s[:, :m]^T   (base - value if base)
static int separable_back(nlc_op* op, const float* y, int B, float* x, float* ws, const float* table, const float* base,
                          cudaStream_t st) {
    const int R = op->R, m = op->m, n = B * op->C;
    float* W1 = ws;                                        // [n][m][m]
    float* W2 = W1 + static_cast<size_t>(n) * m * m;       // [n][m][m]
    float* X1 = W2 + static_cast<size_t>(n) * m * m;       // [n][R][m]
    int rc;
    if ((rc = launch_gemm(st, n, m, m, m, op->Us, 0, 1, m, y, static_cast<long long>(m) * m, m, 1, W1, nullptr, 1,
                          nullptr, nullptr))) return rc;                                  // U_s^T Y
    if ((rc = launch_gemm(st, n, m, m, m, W1, static_cast<long long>(m) * m, m, 1, op->Us2, 0, m, 1, W2, table, op->C,
                          nullptr, nullptr))) return rc;                                  // (W1 U_s) o table
    if ((rc = launch_gemm(st, n, R, m, m, op->Vs, 0, R, 1, W2, static_cast<long long>(m) * m, m, 1, X1, nullptr, 1,
                          nullptr, nullptr))) return rc;                                  // V_s[:, :m] W2
    return launch_gemm(st, n, R, R, m, X1, static_cast<long long>(R) * m, m, 1, op->Vs2, 0, 1, R, x, nullptr, 1, nullptr,
                       base);                                                             // X1 V_s[:, :m]^T
}

template <typename T>
static int to_device(T** dst, const T* src, size_t n) {
    NLC_CHECK_CUDA(cudaMalloc(dst, n * sizeof(T)));
    NLC_CHECK_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return NLC_OK;
}

template <int E>
static int fwht2d_e(const float* in, float* out, const Epilogue& epi, int planes, int C, cudaStream_t st) {
    constexpr int R = 32 * E;
    const long long n_rows = static_cast<long long>(planes) * R;
    fwht_rows_kernel<E><<<blocks_for(n_rows, 8), 256, 0, st>>>(in, out, n_rows);
    NLC_CHECK_LAUNCH();
    const size_t smem = static_cast<size_t>(R) * 33 * sizeof(float);
    if (smem > 48 * 1024)
        NLC_CHECK_CUDA(cudaFuncSetAttribute(fwht_cols_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
    // a gathering epilogue writes to its own array, the transform itself stays in `out`
    fwht_cols_kernel<E><<<dim3(R / 32, planes), 256, smem, st>>>(out, out, epi, C);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
int fwht2d(nlc_op* op, const float* in, float* out, const Epilogue& epi, int B, cudaStream_t st) {
    const int planes = B * op->C;
    // NLC_FWHT_CLUSTER=1: the one-kernel cluster / DSMEM transform (fwht_cluster.cu).  Bit-identical, half the HBM traffic,
    // but measured 2x slower than the two streaming kernels below (profiles/r01m_fwht_cluster_experiment.md), so opt-in.
    const char* env = getenv("NLC_FWHT_CLUSTER");
    const bool use_cluster = env && env[0] == '1';
    if (use_cluster) {
        const int rc = fwht2d_cluster(in, out, epi, planes, op->C, op->R, st);
        if (rc != 1) return rc;
    }
    switch (op->R / 32) {
        case 1: return fwht2d_e<1>(in, out, epi, planes, op->C, st);
        case 2: return fwht2d_e<2>(in, out, epi, planes, op->C, st);
        case 4: return fwht2d_e<4>(in, out, epi, planes, op->C, st);
        case 8: return fwht2d_e<8>(in, out, epi, planes, op->C, st);
        case 16: return fwht2d_e<16>(in, out, epi, planes, op->C, st);
        case 32: return fwht2d_e<32>(in, out, epi, planes, op->C, st);
        default: return set_error(NLC_EINVAL, "WH-CS: image sizes 32 .. 1024 (powers of two) are built");
    }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_op_create(nlc_ctx* ctx, const nlc_op_desc* d, nlc_op** out) {
    NLC_REQUIRE(ctx && d && out, "nlc_op_create: null argument");
    NLC_REQUIRE(d->channels >= 1 && d->R >= 1, "nlc_op_create: bad geometry");
    nlc_op* op = new nlc_op();
    memset(op, 0, sizeof(*op));
    op->ctx = ctx, op->task = d->task, op->C = d->channels, op->R = d->R, op->ratio = d->ratio;
    const long long N = static_cast<long long>(d->R) * d->R, dim = N * d->channels;
    int rc = NLC_OK;
    switch (d->task) {
        case NLC_OP_INPAINT: {
            NLC_REQUIRE(d->idx_host && d->n_idx >= 0 && d->n_idx <= dim, "nlc_op_create: inpainting needs missing indices");
            std::vector<int> pos2k(dim, 0), kept;
            for (int64_t i = 0; i < d->n_idx; ++i) {
                NLC_REQUIRE(d->idx_host[i] >= 0 && d->idx_host[i] < dim, "nlc_op_create: missing index out of range");
                pos2k[d->idx_host[i]] = -1;
            }
            for (long long j = 0; j < dim; ++j)
                if (pos2k[j] == 0) { pos2k[j] = static_cast<int>(kept.size()); kept.push_back(static_cast<int>(j)); }
            op->n_kept = static_cast<int>(kept.size());
            op->ydim = op->n_kept;
            std::vector<int> planar(dim + 3, -1);  // image order, padded to a whole int4
            for (long long p = 0; p < N; ++p)
                for (int c = 0; c < d->channels; ++c) planar[c * N + p] = pos2k[p * d->channels + c];
            if ((rc = to_device(&op->idx_a, kept.data(), kept.size() ? kept.size() : 1)) ||
                (rc = to_device(&op->idx_b, pos2k.data(), pos2k.size())) ||
                (rc = to_device(&op->idx_c, planar.data(), planar.size()))) return rc;
        } break;
        case NLC_OP_COLOR:
        case NLC_OP_SR_AVG: {
            const int K = d->task == NLC_OP_COLOR ? d->channels : d->ratio * d->ratio;
            NLC_REQUIRE(d->U_small_host && d->V_small_host && d->sing_small_host, "nlc_op_create: SVD factors missing");
            NLC_REQUIRE(d->task == NLC_OP_COLOR || (d->ratio >= 1 && d->R % d->ratio == 0), "nlc_op_create: bad ratio");
            op->u = d->U_small_host[0], op->s = d->sing_small_host[0];
            std::vector<float> v0(K);
            for (int k = 0; k < K; ++k) v0[k] = d->V_small_host[k * K];  // first column of V_small
            if ((rc = to_device(&op->v0, v0.data(), K)) ||
                (rc = to_device(&op->Vfull, d->V_small_host, static_cast<size_t>(K) * K))) return rc;
            op->K = K;
            op->Vfull_host = static_cast<float*>(malloc(sizeof(float) * K * K));
            memcpy(op->Vfull_host, d->V_small_host, sizeof(float) * K * K);
            op->ydim = d->task == NLC_OP_COLOR ? N : d->channels * (N / (d->ratio * d->ratio));
        } break;
        case NLC_OP_WHCS: {
            NLC_REQUIRE(d->idx_host && d->n_idx == N && d->ratio >= 1, "nlc_op_create: WH-CS needs perm[R*R]");
            NLC_REQUIRE((d->R & (d->R - 1)) == 0 && d->R >= 32, "nlc_op_create: WH-CS needs R a power of two >= 32");
            std::vector<int> inv(N);
            for (long long j = 0; j < N; ++j) inv[d->idx_host[j]] = static_cast<int>(j);
            if ((rc = to_device(&op->idx_a, inv.data(), inv.size()))) return rc;
            op->ydim = dim / d->ratio;
        } break;
        case NLC_OP_SEPARABLE: {
            NLC_REQUIRE(d->U_small_host && d->V_small_host && d->mult_host && d->pinv_mult_host && d->m_small >= 1 &&
                            d->m_small <= d->R, "nlc_op_create: separable operator needs U_s, V_s and the tables");
            op->m = d->m_small;
            const size_t mm = static_cast<size_t>(op->m) * op->m;
            if ((rc = to_device(&op->Us, d->U_small_host, mm)) || (rc = to_device(&op->Vs, d->V_small_host, N)) ||
                (rc = to_device(&op->mult, d->mult_host, mm * d->channels)) ||
                (rc = to_device(&op->pinv, d->pinv_mult_host, mm * d->channels))) return rc;
            op->Us2 = op->Us, op->Vs2 = op->Vs, op->own2 = false;
            if (d->U_small2_host || d->V_small2_host) {
                NLC_REQUIRE(d->U_small2_host && d->V_small2_host, "nlc_op_create: U_small2 and V_small2 come together");
                if ((rc = to_device(&op->Us2, d->U_small2_host, mm)) || (rc = to_device(&op->Vs2, d->V_small2_host, N)))
                    return rc;
                op->own2 = true;
            }
            if (d->lambda_sing_host && (rc = to_device(&op->lam_s, d->lambda_sing_host, mm))) return rc;
            op->ydim = static_cast<int64_t>(d->channels) * mm;
        } break;
        case NLC_OP_DENOISE:
            op->ydim = dim;
            break;
        case NLC_OP_BLOCKCS: {
            const int E = d->ratio, EE = E * E;
            NLC_REQUIRE(E >= 1 && d->R % E == 0 && d->V_small_host && d->m_small >= 1 && d->m_small <= EE,
                        "nlc_op_create: block CS needs the patch edge (ratio), V_small[E^2,E^2] and 1 <= cs_size <= E^2");
            op->m = d->m_small;
            if ((rc = to_device(&op->Vs, d->V_small_host, static_cast<size_t>(EE) * EE))) return rc;
            op->ydim = static_cast<int64_t>(d->channels) * (d->R / E) * (d->R / E) * op->m;
        } break;
        case NLC_OP_GENERAL: {
            const int64_t nx = d->n_idx, ny = d->m_small;
            NLC_REQUIRE(nx >= 1 && ny >= 1 && ny <= nx && d->U_small_host && d->V_small_host && d->sing_small_host,
                        "nlc_op_create: GeneralA needs U[ny,ny], V[nx,nx], singulars[ny] with ny <= nx");
            op->m = static_cast<int>(ny), op->nx = nx;
            if ((rc = to_device(&op->Us, d->U_small_host, static_cast<size_t>(ny) * ny)) ||
                (rc = to_device(&op->Vs, d->V_small_host, static_cast<size_t>(nx) * nx)) ||
                (rc = to_device(&op->v0, d->sing_small_host, static_cast<size_t>(ny)))) return rc;
            op->ydim = ny;
        } break;
        default:
            delete op;
            return set_error(NLC_EINVAL, "nlc_op_create: unknown task %d", d->task);
    }
    *out = op;
    return NLC_OK;
}

extern "C" void nlc_op_destroy(nlc_op* op) {
    if (!op) return;
    cudaFree(op->idx_a), cudaFree(op->idx_b), cudaFree(op->idx_c), cudaFree(op->v0), cudaFree(op->Vfull), cudaFree(op->lam_s);
    cudaFree(op->Us), cudaFree(op->Vs), cudaFree(op->mult), cudaFree(op->pinv);
    if (op->own2) cudaFree(op->Us2), cudaFree(op->Vs2);
    free(op->Vfull_host);
    delete op;
}

extern "C" int64_t nlc_op_ydim(nlc_op* op) { return op ? op->ydim : 0; }

extern "C" size_t nlc_op_ws(nlc_op* op, int B) {
    if (!op) return 0;
    const size_t plane = static_cast<size_t>(op->R) * op->R, n = static_cast<size_t>(B) * op->C;
    if (op->task == NLC_OP_WHCS) return 2 * n * plane * sizeof(float);
    // separable: T1..T3 + the residual A x0 - y, + one plane set of noise terms and the per-step tables of the DDNM step
    if (op->task == NLC_OP_SEPARABLE) return (5 * n + 3 + op->C) * plane * sizeof(float);
    if (op->task == NLC_OP_BLOCKCS) return 3 * n * plane * sizeof(float);
    if (op->task == NLC_OP_GENERAL) return static_cast<size_t>(B) * (2 * op->m + op->nx) * sizeof(float);
    return 0;
}

// ---------------------------------------------------------------- block CS: 32 x 32 patches <-> rows of a [patches, 1024] matrix
// dir 0: Pm[(bc*np + p)*E*E + pi*E + pj] = x[bc][bi*E+pi][bj*E+pj];  dir 1: the inverse, out = base ? base - v : v * f
__global__ void __launch_bounds__(256) patch_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                          const float* __restrict__ base, long long total, int R, int E,
                                                          int dir, float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // image-order index
    if (i >= total) return;
    const int col = static_cast<int>(i % R), row = static_cast<int>((i / R) % R);
    const long long bc = i / (static_cast<long long>(R) * R);
    const int yd = R / E;
    const long long p = static_cast<long long>(row / E) * yd + col / E;
    const long long j = ((bc * yd * yd + p) * E + row % E) * E + col % E;
    if (dir == 0) out[j] = in[i];
    else out[i] = base ? base[i] - in[j] : __fmul_rn(in[j], f);
}
// GeneralA: T[b][j] *= f(s[j]);  kind 0: s, 1: zero-guarded 1/s, 2: s / (s^2 + eta)
__global__ void __launch_bounds__(256) scale_cols_kernel(float* __restrict__ T, long long total, int ny,
                                                          const float* __restrict__ s, int kind, float eta) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float sv = s[i % ny];
    const float f = kind == 0 ? sv : (kind == 1 ? (sv == 0.f ? 0.f : __fdiv_rn(1.f, sv))
                                                : __fdiv_rn(sv, __fadd_rn(__fmul_rn(sv, sv), eta)));
    T[i] = __fmul_rn(T[i], f);
}

// f / (f^2 + eta) per table entry (A_pinv_eta of the separable operators)
__global__ void pinv_eta_table_kernel(const float* __restrict__ mult, float eta, long long n, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fdiv_rn(mult[i], __fadd_rn(__fmul_rn(mult[i], mult[i]), eta));
}
// Denoising (A = I): mode 0/1/2 copy (x f for A_pinv_eta), 3 project = x0 - (x0 - y)
__global__ void identity_op_kernel(const float* __restrict__ in, const float* __restrict__ y, float* __restrict__ out,
                                   long long n, int mode, float f) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = in[i];
    out[i] = mode == 3 ? __fsub_rn(v, __fsub_rn(v, y[i])) : (mode == 4 ? __fmul_rn(v, f) : v);
}

// mode: 0 A, 1 At, 2 A_pinv, 3 project (in = x0, y = y, out = x0_hat), 4 A_pinv_eta
static int op_apply(nlc_op* op, int mode, const float* in, const float* y, int B, float* out, void* ws_, void* stream_,
                    double eta = 0.0) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(op && in && out && B >= 1, "nlc_op: null argument");
    float* ws = static_cast<float*>(ws_);
    const int C = op->C, R = op->R;
    const long long N = static_cast<long long>(R) * R;
    switch (op->task) {
        case NLC_OP_COLOR:
        case NLC_OP_SR_AVG: {
            const int per_ch = op->task == NLC_OP_SR_AVG, r = per_ch ? op->ratio : 1;
            const long long n_meas = per_ch ? static_cast<long long>(B) * C * (N / (r * r)) : static_cast<long long>(B) * N;
            const unsigned g = blocks_for(n_meas);
            // A reads x = in; At / A_pinv read y = in; project reads both
            const float* xin = (mode == 0 || mode == 3) ? in : nullptr;
            const float* yin = mode == 0 ? nullptr : (mode == 3 ? y : in);
            const float sv = mode == 4 ? op->s / (op->s * op->s + static_cast<float>(eta)) : op->s;
            const bool al16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) |
                                reinterpret_cast<uintptr_t>(y)) & 15) == 0;
#define NLC_MODES(KERNEL, ...)                  \
    switch (mode) {                             \
        case 0: KERNEL<0 __VA_ARGS__; break;    \
        case 1: KERNEL<1 __VA_ARGS__; break;    \
        case 2: KERNEL<2 __VA_ARGS__; break;    \
        case 3: KERNEL<3 __VA_ARGS__; break;    \
        default: KERNEL<4 __VA_ARGS__; break;   \
    }
            if (!per_ch && C == 3 && N % 4 == 0 && al16) {
                const long long n4 = n_meas / 4;
                NLC_MODES(color_vec_kernel, ><<<blocks_for(n4), 256, 0, st>>>(
                    reinterpret_cast<const float4*>(xin), reinterpret_cast<const float4*>(yin),
                    reinterpret_cast<float4*>(out), n4, N / 4, op->u, sv, op->v0))
            } else if (per_ch && r == 4 && al16) {
                NLC_MODES(sr_vec_kernel, , 4><<<g, 256, 0, st>>>(xin, yin, out, n_meas, R, op->u, sv, op->v0))
            } else if (per_ch && r == 2 && al16) {
                NLC_MODES(sr_vec_kernel, , 2><<<g, 256, 0, st>>>(xin, yin, out, n_meas, R, op->u, sv, op->v0))
            } else {
                NLC_MODES(needle_kernel, ><<<g, 256, 0, st>>>(xin, yin, out, B, C, R, r, per_ch, op->u, sv, op->v0))
            }
#undef NLC_MODES
            NLC_CHECK_LAUNCH();
        } break;
        case NLC_OP_INPAINT: {
            if (mode == 0) {
                if (op->n_kept > 0)
                    inpaint_A_kernel<<<blocks_for(static_cast<long long>(B) * op->n_kept), 256, 0, st>>>(
                        in, out, B, C, static_cast<int>(N), op->idx_a, op->n_kept);
            } else {
                const float f = mode == 4 ? 1.0f / (1.0f * 1.0f + static_cast<float>(eta)) : 1.f;
                const long long tot = static_cast<long long>(B) * C * N;
                const bool vec = (C * N) % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
                if (vec && mode == 3)
                    inpaint_back_vec_kernel<3><<<blocks_for(tot / 4), 256, 0, st>>>(
                        reinterpret_cast<const float4*>(in), y, reinterpret_cast<float4*>(out), tot / 4, C * N / 4,
                        reinterpret_cast<const int4*>(op->idx_c), op->n_kept, 1.f);
                else if (vec)
                    inpaint_back_vec_kernel<1><<<blocks_for(tot / 4), 256, 0, st>>>(
                        nullptr, in, reinterpret_cast<float4*>(out), tot / 4, C * N / 4,
                        reinterpret_cast<const int4*>(op->idx_c), op->n_kept, f);
                else if (mode == 3)
                    inpaint_back_kernel<3><<<blocks_for(tot), 256, 0, st>>>(in, y, out, B, C, static_cast<int>(N),
                                                                           op->idx_b, op->n_kept, 1.f);
                else
                    inpaint_back_kernel<1><<<blocks_for(tot), 256, 0, st>>>(nullptr, in, out, B, C, static_cast<int>(N),
                                                                           op->idx_b, op->n_kept, f);
            }
            NLC_CHECK_LAUNCH();
        } break;
        case NLC_OP_WHCS: {
            NLC_REQUIRE(ws, "nlc_op: WH-CS needs a workspace (nlc_op_ws)");
            const int m = static_cast<int>(N / op->ratio);
            float* F = ws;
            float* T = ws + static_cast<size_t>(B) * C * N;
            const unsigned g = blocks_for(B * C * N);
            int rc;
            if (mode == 0) {  // the gather through the permutation rides on the transform's store
                Epilogue e;
                e.invperm = op->idx_a, e.m = m, e.ygather = out;
                if ((rc = fwht2d(op, in, F, e, B, st))) return rc;
            } else if (mode == 3) {
                Epilogue e1;  // T = kept ? FWHT(x0) - y : 0
                e1.invperm = op->idx_a, e1.m = m, e1.ymeas = y;
                if ((rc = fwht2d(op, in, T, e1, B, st))) return rc;
                Epilogue e;
                e.base = in, e.alpha = 1.f, e.beta = -1.f;
                if ((rc = fwht2d(op, T, out, e, B, st))) return rc;  // out = x0 - fwht(T)
            } else {
                const float f = mode == 4 ? 1.0f / (1.0f * 1.0f + static_cast<float>(eta)) : 1.f;
                whcs_scatter_kernel<<<g, 256, 0, st>>>(in, T, B, C, static_cast<int>(N), m, op->idx_a, f);
                NLC_CHECK_LAUNCH();
                if ((rc = fwht2d(op, T, out, Epilogue(), B, st))) return rc;
            }
        } break;
        case NLC_OP_SEPARABLE: {
            NLC_REQUIRE(ws, "nlc_op: separable operators need a workspace (nlc_op_ws)");
            if (mode == 0) return separable_A(op, in, B, out, ws, nullptr, st);
            if (mode == 1) return separable_back(op, in, B, out, ws, op->mult, nullptr, st);
            if (mode == 2) return separable_back(op, in, B, out, ws, op->pinv, nullptr, st);
            if (mode == 4) {
                float* tab = ws + 5 * static_cast<size_t>(B) * C * N;
                const long long nt = static_cast<long long>(C) * op->m * op->m;
                pinv_eta_table_kernel<<<blocks_for(nt), 256, 0, st>>>(op->mult, static_cast<float>(eta), nt, tab);
                NLC_CHECK_LAUNCH();
                return separable_back(op, in, B, out, ws, tab, nullptr, st);
            }
            float* diff = ws + 3 * static_cast<size_t>(B) * C * N;  // A x0 - y
            int rc;
            if ((rc = separable_A(op, in, B, diff, ws, y, st))) return rc;
            return separable_back(op, diff, B, out, ws, op->pinv, in, st);
        }
        case NLC_OP_DENOISE: {
            NLC_REQUIRE(mode != 3 || y, "nlc_op: y is null");
            const long long tot = static_cast<long long>(B) * C * N;
            const float f = 1.0f / (1.0f * 1.0f + static_cast<float>(eta));
            if (tot % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) |
                                  reinterpret_cast<uintptr_t>(y)) & 15) == 0)
                identity_vec_kernel<<<blocks_for(tot / 4), 256, 0, st>>>(reinterpret_cast<const float4*>(in),
                                                                        reinterpret_cast<const float4*>(y),
                                                                        reinterpret_cast<float4*>(out), tot / 4, mode, f);
            else
                identity_op_kernel<<<blocks_for(tot), 256, 0, st>>>(in, y, out, tot, mode, f);
            NLC_CHECK_LAUNCH();
        } break;
        case NLC_OP_BLOCKCS: {
            NLC_REQUIRE(ws, "nlc_op: block CS needs a workspace (nlc_op_ws)");
            const int E = op->ratio, EE = E * E, cs = op->m;
            const long long total = static_cast<long long>(B) * C * N, np = total / EE;
            float *Pm = ws, *D = ws + total, *Pm2 = ws + 2 * total;
            const unsigned g = blocks_for(total);
            int rc;
            if (mode == 0 || mode == 3) {  // Y = patches . V[:, :cs]  (- y)
                patch_rows_kernel<<<g, 256, 0, st>>>(in, Pm, nullptr, total, R, E, 0, 1.f);
                NLC_CHECK_LAUNCH();
                if ((rc = launch_gemm(st, 1, static_cast<int>(np), cs, EE, Pm, 0, EE, 1, op->Vs, 0, EE, 1,
                                      mode == 0 ? out : D, nullptr, 1, mode == 3 ? y : nullptr, nullptr))) return rc;
                if (mode == 0) break;
            }
            // patches = Y . V[:, :cs]^T  (all singular values are 1: A^T = A^+), then back to the image
            if ((rc = launch_gemm(st, 1, static_cast<int>(np), EE, cs, mode == 3 ? D : in, 0, cs, 1, op->Vs, 0, 1, EE, Pm2,
                                  nullptr, 1, nullptr, nullptr))) return rc;
            patch_rows_kernel<<<g, 256, 0, st>>>(Pm2, out, mode == 3 ? in : nullptr, total, R, E, 1,
                                                 mode == 4 ? 1.0f / (1.0f * 1.0f + static_cast<float>(eta)) : 1.f);
            NLC_CHECK_LAUNCH();
        } break;
        case NLC_OP_GENERAL: {
            NLC_REQUIRE(ws, "nlc_op: GeneralA needs a workspace (nlc_op_ws)");
            const int ny = op->m;
            const int nx = static_cast<int>(op->nx);
            float *T = ws, *D = ws + static_cast<size_t>(B) * ny;
            const long long tn = static_cast<long long>(B) * ny;
            int rc;
            const float* yin = in;
            if (mode == 0 || mode == 3) {  // (x V[:, :ny]) o s, then . U^T (- y)
                if ((rc = launch_gemm(st, 1, B, ny, nx, in, 0, nx, 1, op->Vs, 0, nx, 1, T, nullptr, 1, nullptr, nullptr)))
                    return rc;
                scale_cols_kernel<<<blocks_for(tn), 256, 0, st>>>(T, tn, ny, op->v0, 0, 0.f);
                NLC_CHECK_LAUNCH();
                if ((rc = launch_gemm(st, 1, B, ny, ny, T, 0, ny, 1, op->Us, 0, 1, ny, mode == 0 ? out : D, nullptr, 1,
                                      mode == 3 ? y : nullptr, nullptr))) return rc;
                if (mode == 0) break;
                yin = D;
            }
            // (y U) o f(s), then . V[:, :ny]^T
            if ((rc = launch_gemm(st, 1, B, ny, ny, yin, 0, ny, 1, op->Us, 0, ny, 1, T, nullptr, 1, nullptr, nullptr)))
                return rc;
            scale_cols_kernel<<<blocks_for(tn), 256, 0, st>>>(T, tn, ny, op->v0, mode == 1 ? 0 : (mode == 4 ? 2 : 1),
                                                             static_cast<float>(eta));
            NLC_CHECK_LAUNCH();
            return launch_gemm(st, 1, B, nx, ny, T, 0, ny, 1, op->Vs, 0, 1, nx, out, nullptr, 1, nullptr,
                               mode == 3 ? in : nullptr);
        }
        default:
            return set_error(NLC_EINVAL, "nlc_op: unknown task");
    }
    return NLC_OK;
}

extern "C" int nlc_op_A(nlc_op* op, const float* x, int B, float* y, void* ws, void* stream) {
    return op_apply(op, 0, x, nullptr, B, y, ws, stream);
}
extern "C" int nlc_op_At(nlc_op* op, const float* y, int B, float* x, void* ws, void* stream) {
    return op_apply(op, 1, y, nullptr, B, x, ws, stream);
}
extern "C" int nlc_op_Apinv(nlc_op* op, const float* y, int B, float* x, void* ws, void* stream) {
    return op_apply(op, 2, y, nullptr, B, x, ws, stream);
}
extern "C" int nlc_op_Apinv_eta(nlc_op* op, const float* y, int B, double eta, float* x, void* ws, void* stream) {
    return op_apply(op, 4, y, nullptr, B, x, ws, stream, eta);
}
extern "C" int nlc_op_project(nlc_op* op, const float* x0, const float* y, int B, float* x0_hat, void* ws, void* stream) {
    NLC_REQUIRE(y, "nlc_op_project: y is null");
    return op_apply(op, 3, x0, y, B, x0_hat, ws, stream);
}

extern "C" int nlc_image_metrics(nlc_ctx* ctx, const float* x, const float* orig01, int B, int64_t n, float* sample01_out,
                                 float* mse_out, float* l1_out, void* stream_) {
    NLC_REQUIRE(ctx && x && orig01 && mse_out && l1_out && B > 0 && n > 0, "nlc_image_metrics: bad argument");
    image_metrics_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(x, orig01, n, sample01_out, mse_out, l1_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_l1_diff_rows(nlc_ctx* ctx, const float* a, const float* b, int B, int64_t n, float* out,
                                void* stream_) {
    NLC_REQUIRE(ctx && a && b && out, "nlc_l1_diff_rows: null argument");
    l1_diff_rows_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(a, b, n, out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
