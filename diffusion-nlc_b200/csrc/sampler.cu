// Sampler arithmetic around the networks: per-sample norms, the sigma refinement and
// NLC correction with its searchsorted time lookup, eps normalisation, x0 prediction with clipping and every
// pred_xprev variant of src/schedulers.py.  All of it is HBM-bound fp32 elementwise / reduction work, fused so
// that each [B,3,R,R] tensor is read once per kernel with 16-byte accesses.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

// The sampler formulas are evaluated op by op like the reference's torch expressions: the _rn intrinsics stop
// nvcc from contracting a*b+c into an FMA, which would e.g. turn sqrt(sp^2 - sqrt(sp^2)^2) from 0 into NaN.
#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define DIV(a, b) __fdiv_rn((a), (b))

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = threadIdx.x < nw ? red[threadIdx.x] : 0.f;
    if (warp == 0) {
        v = warp_sum(v);
        if (lane == 0) red[0] = v;
    }
    __syncthreads();
    v = red[0];
    __syncthreads();
    return v;
}

// torch.linalg.vector_norm accumulates in fp32; we do too (pairwise through the shuffle tree).
__device__ __forceinline__ float row_sumsq(const float* __restrict__ row, int d, float* red) {
    float acc = 0.f;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    for (int i = threadIdx.x; i < (d >> 2); i += blockDim.x) {
        const float4 v = __ldg(r4 + i);
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
    return block_sum(acc, red);
}

__global__ void __launch_bounds__(1024) row_norm_kernel(const float* __restrict__ x, int d, float* __restrict__ out) {
    __shared__ float red[32];
    const float s = row_sumsq(x + static_cast<size_t>(blockIdx.x) * d, d, red);
    if (threadIdx.x == 0) out[blockIdx.x] = sqrtf(s);
}

__global__ void __launch_bounds__(1024) normalize_rows_kernel(float* __restrict__ x, int d) {
    __shared__ float red[32];
    float* row = x + static_cast<size_t>(blockIdx.x) * d;
    const float nrm = sqrtf(row_sumsq(row, d, red));
    // sqrt(d) * x / clamp(norm, 1e-12)  (src/utils.py:11-16): multiply first, then divide, like the reference
    const float sd = sqrtf(static_cast<float>(d));
    const float den = fmaxf(nrm, 1e-12f);
    float4* r4 = reinterpret_cast<float4*>(row);
    for (int i = threadIdx.x; i < (d >> 2); i += blockDim.x) {
        float4 v = r4[i];
        v.x = sd * v.x / den, v.y = sd * v.y / den, v.z = sd * v.z / den, v.w = sd * v.w / den;
        r4[i] = v;
    }
}

// first index i in [0, n] with table[i] >= v  (torch.searchsorted default side='left')
__device__ __forceinline__ int lower_bound(const float* __restrict__ table, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (table[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// get_t_from_sigma: discrete = searchsorted (src/schedulers.py:185-190); continuous (slopes != NULL) = Interp1d of
// (sigma table -> train timestep) (src/schedulers.py:210-220, src/torchinterp1d.py:96-148): the interval index is
// clamp(searchsorted - 1, 0, n-2) and t = y[ind] + slope[ind] * (sigma - x[ind]) with y = arange(n).
__device__ __forceinline__ float time_from_sigma(const float* __restrict__ table, const float* __restrict__ slopes,
                                                 int n, float s) {
    const int lb = lower_bound(table, n, s);
    if (!slopes) return static_cast<float>(lb);
    const int ind = min(max(lb - 1, 0), n - 2);
    return ADD(static_cast<float>(ind), MUL(slopes[ind], SUB(s, table[ind])));
}

// single CTA over the B-vector (batch-global min for the time shift, src/experiments.py:411-412)
__global__ void __launch_bounds__(1024)
    refine_sigma_kernel(const float* __restrict__ norms, int B, float sqrt_d, const float* __restrict__ sigma_in,
                        int n_sigma_in, float norm_min, float norm_max, int refine, float t_fixed,
                        const float* __restrict__ table, const float* __restrict__ slopes, int n_table,
                        int time_shift, float* __restrict__ sigma_out, float* __restrict__ t_out,
                        float* __restrict__ in_scale_out) {
    __shared__ int s_pos;  // 1 while every t so far is > 0  (t.min() > 0)
    if (threadIdx.x == 0) s_pos = 1;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float s = sigma_in[n_sigma_in == 1 ? 0 : b];
        if (refine) {
            const float nx = DIV(norms[b], sqrt_d);  // vector_norm(xt) / math.sqrt(dim)
            const float lo = fmaxf(SUB(nx, norm_max), 0.f), hi = ADD(nx, norm_min);
            s = fminf(fmaxf(s, lo), hi);
            if (!(time_from_sigma(table, slopes, n_table, s) > 0.f)) atomicAnd(&s_pos, 0);
        }
        sigma_out[b] = s;
        if (in_scale_out) in_scale_out[b] = sqrtf(DIV(1.0f, ADD(MUL(s, s), 1.0f)));
        if (!refine) t_out[b] = fminf(fmaxf(t_fixed, 0.f), 1000.f);
    }
    __syncthreads();
    if (refine) {
        const float shift = s_pos ? static_cast<float>(time_shift) : 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const float t = SUB(time_from_sigma(table, slopes, n_table, sigma_out[b]), shift);
            t_out[b] = fminf(fmaxf(t, 0.f), 1000.f);
        }
    }
}

__global__ void sigma_correct_kernel(const float* __restrict__ r, const float* __restrict__ sigma,
                                     const float* __restrict__ sigma_prev, int n_prev, int B, int update_prev,
                                     const float* __restrict__ table, const float* __restrict__ slopes, int n_table,
                                     float* __restrict__ sigma_hat,
                                     float* __restrict__ sigma_prev_hat, float* __restrict__ t_hat,
                                     float* __restrict__ in_scale_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float s = sigma[b];
    const float sp = sigma_prev[n_prev == 1 ? 0 : b];
    const float dist = MUL(s, ADD(1.0f, r[b]));
    const float dist_prev = MUL(dist, DIV(sp, s));
    const float t = time_from_sigma(table, slopes, n_table, dist);
    sigma_hat[b] = dist;
    sigma_prev_hat[b] = update_prev ? dist_prev : sp;
    t_hat[b] = fminf(fmaxf(t, 0.f), 1000.f);
    if (in_scale_out) in_scale_out[b] = sqrtf(DIV(1.0f, ADD(MUL(dist, dist), 1.0f)));
}

// grid (chunks, B)
__global__ void __launch_bounds__(256) pred_xstart_kernel(const float* __restrict__ xt, const float* __restrict__ eps,
                                                           const float* __restrict__ sigma, int n_sigma, int d, int clip,
                                                           float* __restrict__ x0) {
    const int b = blockIdx.y;
    const float s = sigma[n_sigma == 1 ? 0 : b];
    const size_t base = static_cast<size_t>(b) * d;
    const float4* x4 = reinterpret_cast<const float4*>(xt + base);
    const float4* e4 = reinterpret_cast<const float4*>(eps + base);
    float4* o4 = reinterpret_cast<float4*>(x0 + base);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (d >> 2); i += gridDim.x * blockDim.x) {
        const float4 x = __ldg(x4 + i), e = __ldg(e4 + i);
        float4 o = make_float4(SUB(x.x, MUL(s, e.x)), SUB(x.y, MUL(s, e.y)), SUB(x.z, MUL(s, e.z)),
                               SUB(x.w, MUL(s, e.w)));
        if (clip == NLC_CLIP_CLAMP) {
            o.x = fminf(fmaxf(o.x, -1.f), 1.f), o.y = fminf(fmaxf(o.y, -1.f), 1.f);
            o.z = fminf(fmaxf(o.z, -1.f), 1.f), o.w = fminf(fmaxf(o.w, -1.f), 1.f);
        }
        o4[i] = o;
    }
}

struct XprevArgs {
    int sched;
    double eta;
    const float *x0, *eps, *xt, *noise, *learned_v;
    int logvar_mode;
    float min_var_coef;
    const float* sigma;
    int n_sigma;
    const float* sigma_prev;
    int n_prev;
    int d;
    float* x_prev;
    int* nan_flag;
};

// One thread-block slice of one sample. Scalars follow the reference's fp32 operation order.
__global__ void __launch_bounds__(256) pred_xprev_kernel(const XprevArgs a) {
    const int b = blockIdx.y;
    const float st = a.sigma[a.n_sigma == 1 ? 0 : b];
    const float sp = a.sigma_prev[a.n_prev == 1 ? 0 : b];
    const float eta = static_cast<float>(a.eta);
    // get_eps_logvar (src/schedulers.py:367-390)
    const float st2 = MUL(st, st), sp2 = MUL(sp, sp);
    float max_lv = 0.f, min_lv = 0.f;
    if (a.logvar_mode != 0) {
        float beta_t = DIV(SUB(st2, sp2), ADD(st2, 1.0f));
        beta_t = fmaxf(fabsf(beta_t), 1e-20f);
        const float alpha_t = DIV(1.0f, ADD(st2, 1.0f)), alpha_prev = DIV(1.0f, ADD(sp2, 1.0f));
        float coef = DIV(SUB(1.0f, alpha_prev), SUB(1.0f, alpha_t));
        coef = fminf(fmaxf(coef, 0.f), 1.f);
        const float post_var = MUL(beta_t, coef);
        max_lv = logf(beta_t);
        min_lv = logf(fmaxf(post_var, a.min_var_coef));
    }
    const float abp_sqrt = sqrtf(DIV(1.0f, ADD(sp2, 1.0f)));  // sqrt(alpha_bar_prev)
    const float mask = sp > 0.f ? 1.f : 0.f;
    const float simple_signal = MUL(static_cast<float>(sqrt(1.0 - a.eta * a.eta)), sp);
    const float eta_sp = MUL(eta, sp);
    // DDPM_orig posterior coefficients (src/schedulers.py:581-599)
    float pm1 = 0.f, pm2 = 0.f, ab_sqrt = 0.f, abp_sqrt2 = 0.f;
    if (a.sched == NLC_SCHED_DDPM_ORIG) {
        const float alpha_bar = DIV(1.0f, ADD(st2, 1.0f)), alpha_bar_prev = DIV(1.0f, ADD(sp2, 1.0f));
        const float alpha_t = DIV(alpha_bar, alpha_bar_prev), beta_t = SUB(1.0f, alpha_t);
        pm1 = DIV(MUL(beta_t, sqrtf(alpha_bar_prev)), SUB(1.0f, alpha_bar));
        pm2 = DIV(MUL(SUB(1.0f, alpha_bar_prev), sqrtf(alpha_t)), SUB(1.0f, alpha_bar));
        ab_sqrt = sqrtf(alpha_bar);
        abp_sqrt2 = sqrtf(alpha_bar_prev);
    }
    const size_t base = static_cast<size_t>(b) * a.d;
    bool saw_nan = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (a.d >> 2); i += gridDim.x * blockDim.x) {
        const float4 x0v = __ldg(reinterpret_cast<const float4*>(a.x0 + base) + i);
        float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), xtv = ev, nv = ev, lv = ev;
        const bool rederive = a.sched == NLC_SCHED_DDIM_SIMPLE_ORIG || a.sched == NLC_SCHED_DDIM_SIMPLE_DRAG ||
                              a.sched == NLC_SCHED_DDIM_ORIG;
        if (rederive || a.sched == NLC_SCHED_DDPM_ORIG) xtv = __ldg(reinterpret_cast<const float4*>(a.xt + base) + i);
        if (!rederive && a.sched != NLC_SCHED_DDPM_ORIG) ev = __ldg(reinterpret_cast<const float4*>(a.eps + base) + i);
        if (a.noise) nv = __ldg(reinterpret_cast<const float4*>(a.noise + base) + i);
        if (a.logvar_mode == 1) lv = __ldg(reinterpret_cast<const float4*>(a.learned_v + base) + i);
        const float X0[4] = {x0v.x, x0v.y, x0v.z, x0v.w};
        const float XT[4] = {xtv.x, xtv.y, xtv.z, xtv.w};
        const float NZ[4] = {nv.x, nv.y, nv.z, nv.w};
        const float LV[4] = {lv.x, lv.y, lv.z, lv.w};
        float E[4] = {ev.x, ev.y, ev.z, ev.w};
        float O[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (rederive) E[k] = DIV(SUB(XT[k], X0[k]), st);
            float logvar = 0.f;
            if (a.logvar_mode == 1) {
                const float frac = DIV(ADD(LV[k], 1.0f), 2.0f);
                logvar = ADD(MUL(frac, max_lv), MUL(SUB(1.0f, frac), min_lv));
            } else if (a.logvar_mode == 2) {
                logvar = min_lv;
            } else if (a.logvar_mode == 3) {
                logvar = max_lv;
            }
            float o;
            switch (a.sched) {
                case NLC_SCHED_DDIM:
                case NLC_SCHED_DDIM_ORIG: {
                    float noise_sigma = 0.f, nz = 0.f;
                    if (a.eta > 0) {
                        noise_sigma = DIV(MUL(eta, expf(MUL(0.5f, logvar))), abp_sqrt);
                        nz = MUL(mask, NZ[k]);
                    }
                    const float signal = sqrtf(fmaxf(SUB(sp2, MUL(noise_sigma, noise_sigma)), 0.f));
                    if (a.sched == NLC_SCHED_DDIM) noise_sigma = sqrtf(SUB(sp2, MUL(signal, signal)));
                    o = ADD(ADD(X0[k], MUL(signal, E[k])), MUL(noise_sigma, nz));
                } break;
                case NLC_SCHED_DDIM_SIMPLE:
                case NLC_SCHED_DDIM_SIMPLE_ORIG: {
                    o = ADD(X0[k], MUL(simple_signal, E[k]));
                    if (a.eta > 0) o = ADD(o, MUL(eta_sp, NZ[k]));
                } break;
                case NLC_SCHED_DDIM_SIMPLE_DRAG: {
                    o = ADD(X0[k], MUL(sp, E[k]));
                    if (a.eta > 0) o = ADD(o, MUL(eta_sp, NZ[k]));
                } break;
                case NLC_SCHED_DDPM: {
                    const float noise_sigma = DIV(expf(MUL(0.5f, logvar)), abp_sqrt);
                    const float signal = sqrtf(fmaxf(SUB(sp2, MUL(noise_sigma, noise_sigma)), 0.f));
                    o = ADD(X0[k], MUL(signal, E[k]));
                    o = ADD(o, MUL(noise_sigma, MUL(mask, NZ[k])));
                } break;
                default: {  // NLC_SCHED_DDPM_ORIG
                    const float zt = MUL(XT[k], ab_sqrt);
                    const float mean = ADD(MUL(pm1, X0[k]), MUL(pm2, zt));
                    const float zprev = ADD(mean, MUL(MUL(mask, expf(MUL(0.5f, logvar))), NZ[k]));
                    o = DIV(zprev, abp_sqrt2);
                } break;
            }
            O[k] = o;
            saw_nan |= (o != o);
        }
        reinterpret_cast<float4*>(a.x_prev + base)[i] = make_float4(O[0], O[1], O[2], O[3]);
    }
    if (a.nan_flag && saw_nan) atomicOr(a.nan_flag, 1);
}

static int row_grid(int sm_count, int B, int d) {
    int chunks = 1;
    while (static_cast<long long>(chunks) * B < 4LL * sm_count && chunks * 2 * 1024 <= d) chunks *= 2;
    return chunks;
}


// ---------------------------------------------------------------- dynamic thresholding (src/experiments.py:190-204)
// s_b = clamp(quantile(|x_b|, q), 1, max_value);  x_b <- clamp(x_b, -s_b, s_b) / s_b.
// torch.quantile (linear interpolation) = lerp(sorted[floor(rank)], sorted[ceil(rank)], frac(rank)) with
// rank = q*(n-1) evaluated in fp32.  The two order statistics are found exactly by a 3-pass radix select over the
// bit patterns of |x| (non-negative floats order like unsigned integers); one CTA per sample, the sample stays in L2.
constexpr int kSelBins = 2048;
constexpr int kSelCopies = 8;

__device__ __forceinline__ uint32_t abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }

__global__ void __launch_bounds__(1024)
    dynamic_threshold_kernel(float* __restrict__ x, int d, int rank_lo, int rank_hi, float weight, float max_value,
                             float* __restrict__ s_out) {
    extern __shared__ uint32_t hist[];  // [kSelCopies][kSelBins]
    __shared__ uint32_t sh_prefix, sh_k, sh_count_eq, sh_min_above;
    float* row = x + static_cast<size_t>(blockIdx.x) * d;
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const int n4 = d >> 2;
    const int copy = (threadIdx.x >> 5) & (kSelCopies - 1);
    uint32_t prefix = 0, mask = 0, k = static_cast<uint32_t>(rank_lo);
    const int shifts[3] = {21, 10, 0};
    const uint32_t widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = shifts[pass];
        const uint32_t bmask = (1u << widths[pass]) - 1u;
        for (int i = threadIdx.x; i < kSelCopies * kSelBins; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 v = __ldg(r4 + i);
            const uint32_t b[4] = {abs_bits(v.x), abs_bits(v.y), abs_bits(v.z), abs_bits(v.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((b[j] & mask) == prefix) atomicAdd(&hist[copy * kSelBins + ((b[j] >> shift) & bmask)], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) {
            uint32_t c = 0;
#pragma unroll
            for (int j = 0; j < kSelCopies; ++j) c += hist[j * kSelBins + i];
            hist[i] = c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cum = 0, bin = 0;
            for (; bin < bmask; ++bin) {
                if (cum + hist[bin] > k) break;
                cum += hist[bin];
            }
            sh_prefix = prefix | (bin << shift);
            sh_k = k - cum;
            sh_count_eq = hist[bin];
        }
        __syncthreads();
        prefix = sh_prefix, k = sh_k;
        mask |= bmask << shift;
        __syncthreads();
    }
    // prefix = bits of sorted[rank_lo]; k = its index among the sh_count_eq equal values
    const float v_lo = __uint_as_float(prefix);
    float v_hi = v_lo;
    if (rank_hi > rank_lo && k + 1 >= sh_count_eq) {  // the next order statistic is the smallest value above v_lo
        if (threadIdx.x == 0) sh_min_above = 0x7f800000u;
        __syncthreads();
        uint32_t m = 0x7f800000u;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 v = __ldg(r4 + i);
            const uint32_t b[4] = {abs_bits(v.x), abs_bits(v.y), abs_bits(v.z), abs_bits(v.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (b[j] > prefix && b[j] < m) m = b[j];
        }
        atomicMin(&sh_min_above, m);
        __syncthreads();
        v_hi = __uint_as_float(sh_min_above);
    }
    // Tensor.lerp: w < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float diff = SUB(v_hi, v_lo);
    float q = weight < 0.5f ? ADD(v_lo, MUL(weight, diff)) : SUB(v_hi, MUL(diff, SUB(1.0f, weight)));
    const float sc = fminf(fmaxf(q, 1.0f), max_value);
    if (threadIdx.x == 0 && s_out) s_out[blockIdx.x] = sc;
    float4* w4 = reinterpret_cast<float4*>(row);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = w4[i];
        v.x = DIV(fminf(fmaxf(v.x, -sc), sc), sc), v.y = DIV(fminf(fmaxf(v.y, -sc), sc), sc);
        v.z = DIV(fminf(fmaxf(v.z, -sc), sc), sc), v.w = DIV(fminf(fmaxf(v.w, -sc), sc), sc);
        w4[i] = v;
    }
}

// ---------------------------------------------------------------- projection_loop sigma feed-forward
// image_sample.py:483-496: cur_norm = ||x_{t-1}||/sqrt(d); cur_dist = sqrt(cur_norm^2 + N^2 - 2 cur_norm N 0.99 + 1e-8);
// sigma <- r0*sigma_prev_orig + r1*sigma_prev + r2*sigma_t*(cur_norm/last_norm) + r3*cur_dist; t <- get_t_from_sigma
__global__ void sigma_estimate_kernel(const float* __restrict__ norms, float* __restrict__ last_norm, int B,
                                      float sqrt_d, float norm_max, float sigma_prev_orig,
                                      const float* __restrict__ sigma_prev, int n_prev,
                                      const float* __restrict__ sigma_t, int n_t, float r0, float r1, float r2, float r3,
                                      const float* __restrict__ table, const float* __restrict__ slopes, int n_table,
                                      float* __restrict__ sigma_out, float* __restrict__ t_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float cur = DIV(norms[b], sqrt_d);
    const float nm2 = static_cast<float>(static_cast<double>(norm_max) * static_cast<double>(norm_max));
    float dist = ADD(MUL(cur, cur), nm2);
    dist = SUB(dist, MUL(MUL(MUL(2.0f, cur), norm_max), 0.99f));
    dist = sqrtf(ADD(dist, 1e-8f));
    const float ratio = DIV(cur, last_norm[b]);
    const float s1 = sigma_prev[n_prev == 1 ? 0 : b];
    const float s2 = MUL(sigma_t[n_t == 1 ? 0 : b], ratio);
    float s = ADD(ADD(ADD(MUL(r0, sigma_prev_orig), MUL(r1, s1)), MUL(r2, s2)), MUL(r3, dist));
    sigma_out[b] = s;
    t_out[b] = time_from_sigma(table, slopes, n_table, s);
    last_norm[b] = cur;
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_row_norm(nlc_ctx* ctx, const float* x, int B, int d, float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && out && d % 4 == 0, "nlc_row_norm: bad argument (d %% 4 == 0 required)");
    row_norm_kernel<<<B, 1024, 0, stream>>>(x, d, out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_normalize_rows(nlc_ctx* ctx, float* x, int B, int d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && d % 4 == 0, "nlc_normalize_rows: bad argument (d %% 4 == 0 required)");
    normalize_rows_kernel<<<B, 1024, 0, stream>>>(x, d);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_refine_sigma(nlc_ctx* ctx, const float* norms, int B, int d, const float* sigma_in, int n_sigma_in,
                                float norm_min, float norm_max, int refine, float t_fixed, const float* sigma_table,
                                const float* slopes, int n_table, int time_shift, float* sigma_out, float* t_out,
                                float* in_scale_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && sigma_in && sigma_out && t_out, "nlc_refine_sigma: null argument");
    NLC_REQUIRE(!refine || (norms && sigma_table), "nlc_refine_sigma: refine needs norms and the sigma table");
    NLC_REQUIRE(n_sigma_in == 1 || n_sigma_in == B, "nlc_refine_sigma: n_sigma_in must be 1 or B");
    refine_sigma_kernel<<<1, 1024, 0, stream>>>(norms, B, static_cast<float>(sqrt(static_cast<double>(d))), sigma_in, n_sigma_in,
                                                norm_min, norm_max, refine, t_fixed, sigma_table, slopes, n_table,
                                                time_shift, sigma_out, t_out, in_scale_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_sigma_correct(nlc_ctx* ctx, const float* r, const float* sigma, const float* sigma_prev, int n_prev,
                                 int B, int update_prev, const float* sigma_table, const float* slopes, int n_table,
                                 float* sigma_hat, float* sigma_prev_hat, float* t_hat, float* in_scale_out,
                                 void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && r && sigma && sigma_prev && sigma_table && sigma_hat && sigma_prev_hat && t_hat,
                "nlc_sigma_correct: null argument");
    NLC_REQUIRE(n_prev == 1 || n_prev == B, "nlc_sigma_correct: n_prev must be 1 or B");
    sigma_correct_kernel<<<(B + 127) / 128, 128, 0, stream>>>(r, sigma, sigma_prev, n_prev, B, update_prev, sigma_table,
                                                              slopes, n_table, sigma_hat, sigma_prev_hat, t_hat,
                                                              in_scale_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_pred_xstart(nlc_ctx* ctx, const float* xt, const float* eps, const float* sigma, int n_sigma, int B,
                               int d, int clip, float* x0, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && xt && eps && sigma && x0 && d % 4 == 0, "nlc_pred_xstart: bad argument");
    NLC_REQUIRE(n_sigma == 1 || n_sigma == B, "nlc_pred_xstart: n_sigma must be 1 or B");
    NLC_REQUIRE(clip == NLC_CLIP_NONE || clip == NLC_CLIP_CLAMP, "nlc_pred_xstart: clip mode %d unsupported", clip);
    pred_xstart_kernel<<<dim3(row_grid(ctx->sm_count, B, d), B), 256, 0, stream>>>(xt, eps, sigma, n_sigma, d, clip, x0);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_pred_xprev(nlc_ctx* ctx, int sched, double eta, const float* x0, const float* eps, const float* xt,
                              const float* noise, const float* learned_v, int logvar_mode, float min_var_coef,
                              const float* sigma, int n_sigma, const float* sigma_prev, int n_prev, int B, int d,
                              float* x_prev, int* nan_flag, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x0 && sigma && sigma_prev && x_prev && d % 4 == 0, "nlc_pred_xprev: bad argument");
    NLC_REQUIRE(sched >= NLC_SCHED_DDIM && sched <= NLC_SCHED_DDIM_ORIG, "nlc_pred_xprev: unknown scheduler %d", sched);
    const bool rederive = sched == NLC_SCHED_DDIM_SIMPLE_ORIG || sched == NLC_SCHED_DDIM_SIMPLE_DRAG ||
                          sched == NLC_SCHED_DDIM_ORIG;
    NLC_REQUIRE(!(rederive || sched == NLC_SCHED_DDPM_ORIG) || xt, "nlc_pred_xprev: this scheduler needs xt");
    NLC_REQUIRE(rederive || sched == NLC_SCHED_DDPM_ORIG || eps, "nlc_pred_xprev: this scheduler needs eps");
    const bool needs_noise = sched == NLC_SCHED_DDPM || sched == NLC_SCHED_DDPM_ORIG || eta > 0;
    NLC_REQUIRE(!needs_noise || noise, "nlc_pred_xprev: noise tensor required (eta > 0 or DDPM)");
    const bool needs_lv = sched == NLC_SCHED_DDPM || sched == NLC_SCHED_DDPM_ORIG ||
                          ((sched == NLC_SCHED_DDIM || sched == NLC_SCHED_DDIM_ORIG) && eta > 0);
    NLC_REQUIRE(!needs_lv || logvar_mode != 0, "nlc_pred_xprev: this scheduler needs a log-variance (sampler_var)");
    NLC_REQUIRE(logvar_mode != 1 || learned_v, "nlc_pred_xprev: learned log-variance tensor missing");
    XprevArgs a{sched, eta, x0, eps, xt, noise, learned_v, logvar_mode, min_var_coef, sigma, n_sigma, sigma_prev,
                n_prev, d, x_prev, nan_flag};
    pred_xprev_kernel<<<dim3(row_grid(ctx->sm_count, B, d), B), 256, 0, stream>>>(a);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_dynamic_threshold(nlc_ctx* ctx, float* x, int B, int d, double ratio, float max_value, float* s_out,
                                     void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && B > 0 && d >= 4 && d % 4 == 0, "nlc_dynamic_threshold: bad argument (d %% 4 == 0 required)");
    NLC_REQUIRE(ratio >= 0.0 && ratio <= 1.0, "nlc_dynamic_threshold: ratio must be in [0,1]");
    // torch.quantile evaluates q * (n - 1) in the input dtype (fp32)
    const float rank = static_cast<float>(ratio) * static_cast<float>(d - 1);
    const float lo = floorf(rank);
    const int rank_lo = static_cast<int>(lo), rank_hi = static_cast<int>(ceilf(rank));
    const size_t smem = static_cast<size_t>(kSelCopies) * kSelBins * sizeof(uint32_t);
    NLC_REQUIRE_DEVICE(ctx);
    static nlc::PerDeviceFlag configured;
    if (!configured[ctx->device]) {
        NLC_CHECK_CUDA(cudaFuncSetAttribute(dynamic_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
        configured[ctx->device] = true;
    }
    dynamic_threshold_kernel<<<B, 1024, smem, stream>>>(x, d, rank_lo, rank_hi, rank - lo, max_value, s_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_sigma_estimate(nlc_ctx* ctx, const float* norms, float* last_norm, int B, int d, float norm_max,
                                  float sigma_prev_orig, const float* sigma_prev, int n_prev, const float* sigma_t,
                                  int n_t, const float* rates4_host, const float* sigma_table, const float* slopes,
                                  int n_table, float* sigma_out, float* t_out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && norms && last_norm && sigma_prev && sigma_t && rates4_host && sigma_table && sigma_out && t_out,
                "nlc_sigma_estimate: null argument");
    NLC_REQUIRE((n_prev == 1 || n_prev == B) && (n_t == 1 || n_t == B), "nlc_sigma_estimate: vector lengths must be 1 or B");
    sigma_estimate_kernel<<<(B + 127) / 128, 128, 0, stream>>>(
        norms, last_norm, B, static_cast<float>(sqrt(static_cast<double>(d))), norm_max, sigma_prev_orig, sigma_prev,
        n_prev, sigma_t, n_t, rates4_host[0], rates4_host[1], rates4_host[2], rates4_host[3], sigma_table, slopes,
        n_table, sigma_out, t_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}


// ----------------------------------------------------------------------------------------------------------------
// Best-x0 bookkeeping of the DDIM-family loop on the device (src/experiments.py:371-376: `const_val = mean(const);
// if const_val < best_val: best_x0 = x0.clone(); best_val = const_val`).  The reference decides on the host, which costs
// one device->host read per step and keeps the step out of a CUDA graph; here the comparison and the conditional copy
// are two stream-ordered kernels.
namespace nlc {

__global__ void best_decide_kernel(const float* __restrict__ loss_sum, float inv_count, float* __restrict__ best_val,
                                   int* __restrict__ flag) {
    const float v = __fmul_rn(*loss_sum, inv_count);
    const int better = v < *best_val;  // (a NaN mean is never better, as in the reference's host comparison)
    *flag = better;
    if (better) *best_val = v;
}

__global__ void __launch_bounds__(256) copy_if_kernel(const int* __restrict__ flag, const float4* __restrict__ src,
                                                      float4* __restrict__ dst, size_t n4) {
    if (*flag == 0) return;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = src[i];
}

}  // namespace nlc

extern "C" int nlc_best_update(nlc_ctx* ctx, const float* loss_sum, float inv_count, float* best_val, int* flag,
                               const float* x0, float* best_x0, int64_t n, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && loss_sum && best_val && flag && x0 && best_x0 && n > 0 && n % 4 == 0,
                "nlc_best_update: null argument or n %% 4 != 0");
    NLC_REQUIRE(((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(best_x0)) & 15) == 0,
                "nlc_best_update: x0 / best_x0 must be 16-byte aligned");
    nlc::best_decide_kernel<<<1, 1, 0, stream>>>(loss_sum, inv_count, best_val, flag);
    NLC_CHECK_LAUNCH();
    const size_t n4 = static_cast<size_t>(n) / 4;
    size_t blocks = (n4 + 255) / 256;
    const size_t cap = static_cast<size_t>(ctx->sm_count) * 8;
    if (blocks > cap) blocks = cap;
    nlc::copy_if_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        flag, reinterpret_cast<const float4*>(x0), reinterpret_cast<float4*>(best_x0), n4);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
