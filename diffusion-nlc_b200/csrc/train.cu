// SURVEY §8f rank 3, first slice: the parts of the sigma-model training step (src/experiments.py:654-694) that are not the
// sigma-model's own forward / backward.
//   nlc_train_prepare    the perturbed-noise batch of :661-669 in one pass per sample:
//                          new_noise = noise + eta1 noise + (eta1 eta2) extra,  dist_real = ||new_noise||_2 / sqrt(d),
//                          noisy_x = x0 sqrt(alpha_bar_t) + new_noise sqrt(1 - alpha_bar_t)      (src/schedulers.py:323-329)
//                        and its EDM variant of :996-1001 (new_noise = noise + eta1 (noise + eta2 extra), x0 + sigma new_noise)
//                        (the reference: 9 elementwise launches and a norm over [B, d])
//   nlc_adamw_ema_step   torch.optim.AdamW's update (decoupled weight decay, bias-corrected moments) of :692 and the EMA
//                        of the master parameters (:233-236) fused into ONE pass over a flat fp32 parameter buffer, with the
//                        gradient scale of a data-parallel mean folded in (the reference's DDP runs under no_sync(): its
//                        ranks never average their gradients).  5 reads + 4 writes per parameter.
// The frozen UNet `encode` of the step (99 % of its FLOPs) is the sampling path's engine.  All HBM-bound fp32.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace nlc {

__global__ void __launch_bounds__(1024) train_prepare_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                              const float* __restrict__ extra,
                                                              const float* __restrict__ eta1, const float* __restrict__ eta2,
                                                              const float* __restrict__ alpha_bar, int edm, long long d,
                                                              float inv_sqrt_d, float* __restrict__ noisy_x,
                                                              float* __restrict__ new_noise_out,
                                                              float* __restrict__ dist_real) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const size_t base = static_cast<size_t>(b) * d;
    const float e1 = eta1[b], e12 = __fmul_rn(eta1[b], eta2[b]);
    const float e2 = eta2[b];
    const float ab = alpha_bar[b];  // alpha_bar_t, or sigma for the EDM step
    const float sa = edm ? 1.0f : sqrtf(ab), sn = edm ? ab : sqrtf(__fsub_rn(1.0f, ab));
    float acc = 0.f;
    auto one = [&](float xv, float nv, float ev, float& nn) -> float {
        // DDIM: noise + (eta1 * noise + (eta1 * eta2) * extra)   (src/experiments.py:666-667)
        // EDM:  noise + eta1 * (noise + eta2 * extra)            (:998-999)
        nn = edm ? __fadd_rn(nv, __fmul_rn(e1, __fadd_rn(nv, __fmul_rn(e2, ev))))
                 : __fadd_rn(nv, __fadd_rn(__fmul_rn(e1, nv), __fmul_rn(e12, ev)));
        acc = fmaf(nn, nn, acc);
        // x0 sqrt(ab) + new_noise sqrt(1 - ab)  |  x0 + sigma new_noise   (:1001)
        return edm ? __fadd_rn(xv, __fmul_rn(sn, nn)) : __fadd_rn(__fmul_rn(xv, sa), __fmul_rn(nn, sn));
    };
    if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(noise) |
                          reinterpret_cast<uintptr_t>(extra) | reinterpret_cast<uintptr_t>(noisy_x) |
                          reinterpret_cast<uintptr_t>(new_noise_out)) & 15) == 0) {
        const long long d4 = d >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x0 + base);
        const float4* n4 = reinterpret_cast<const float4*>(noise + base);
        const float4* e4 = reinterpret_cast<const float4*>(extra + base);
        for (long long i = threadIdx.x; i < d4; i += blockDim.x) {
            const float4 xv = __ldg(x4 + i), nv = __ldg(n4 + i), ev = __ldg(e4 + i);
            float4 o, nn;
            o.x = one(xv.x, nv.x, ev.x, nn.x), o.y = one(xv.y, nv.y, ev.y, nn.y);
            o.z = one(xv.z, nv.z, ev.z, nn.z), o.w = one(xv.w, nv.w, ev.w, nn.w);
            reinterpret_cast<float4*>(noisy_x + base)[i] = o;
            if (new_noise_out) reinterpret_cast<float4*>(new_noise_out + base)[i] = nn;
        }
    } else {
        for (long long i = threadIdx.x; i < d; i += blockDim.x) {
            float nn;
            noisy_x[base + i] = one(x0[base + i], noise[base + i], extra[base + i], nn);
            if (new_noise_out) new_noise_out[base + i] = nn;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) dist_real[b] = __fmul_rn(sqrtf(acc), inv_sqrt_d);
    }
}

struct AdamArgs {
    float lr_wd_keep;   // 1 - lr * weight_decay
    float beta1_c;      // 1 - beta1
    float beta2, beta2_c;
    float step_size;    // lr / (1 - beta1^t)
    float sqrt_bc2;     // sqrt(1 - beta2^t)
    float eps, grad_scale, ema_rate, ema_c;
};
// torch.optim.AdamW (single-tensor path, amsgrad off, maximize off), then ema = ema * rate + p * (1 - rate)
template <int V>
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         float* __restrict__ ema, long long n, const AdamArgs a) {
    const long long i0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * V;
    if (i0 >= n) return;
    float pv[V], gv[V], mv[V], vv[V], ev[V];
    if (V == 4) {
        const float4 t0 = *reinterpret_cast<const float4*>(p + i0), t1 = __ldg(reinterpret_cast<const float4*>(g + i0));
        const float4 t2 = *reinterpret_cast<const float4*>(m + i0), t3 = *reinterpret_cast<const float4*>(v + i0);
        pv[0] = t0.x, pv[1] = t0.y, pv[2] = t0.z, pv[3] = t0.w, gv[0] = t1.x, gv[1] = t1.y, gv[2] = t1.z, gv[3] = t1.w;
        mv[0] = t2.x, mv[1] = t2.y, mv[2] = t2.z, mv[3] = t2.w, vv[0] = t3.x, vv[1] = t3.y, vv[2] = t3.z, vv[3] = t3.w;
        if (ema) {
            const float4 t4 = *reinterpret_cast<const float4*>(ema + i0);
            ev[0] = t4.x, ev[1] = t4.y, ev[2] = t4.z, ev[3] = t4.w;
        }
    } else {
        pv[0] = p[i0], gv[0] = g[i0], mv[0] = m[i0], vv[0] = v[i0];
        if (ema) ev[0] = ema[i0];
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const float gr = __fmul_rn(gv[k], a.grad_scale);
        float pp = __fmul_rn(pv[k], a.lr_wd_keep);                                   // p.mul_(1 - lr * wd)
        mv[k] = __fadd_rn(mv[k], __fmul_rn(__fsub_rn(gr, mv[k]), a.beta1_c));        // exp_avg.lerp_(grad, 1 - beta1)
        vv[k] = __fadd_rn(__fmul_rn(vv[k], a.beta2), __fmul_rn(__fmul_rn(gr, gr), a.beta2_c));  // mul_(b2).addcmul_(g, g, 1-b2)
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv[k]), a.sqrt_bc2), a.eps);   // (sqrt(v) / sqrt(bc2)).add_(eps)
        pp = __fsub_rn(pp, __fmul_rn(a.step_size, __fdiv_rn(mv[k], denom)));         // p.addcdiv_(m, denom, value=-step_size)
        pv[k] = pp;
        if (ema) ev[k] = __fadd_rn(__fmul_rn(ev[k], a.ema_rate), __fmul_rn(pp, a.ema_c));  // mul_(rate).add_(p, alpha=1-rate)
    }
    if (V == 4) {
        *reinterpret_cast<float4*>(p + i0) = make_float4(pv[0], pv[1], pv[2], pv[3]);
        *reinterpret_cast<float4*>(m + i0) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(v + i0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (ema) *reinterpret_cast<float4*>(ema + i0) = make_float4(ev[0], ev[1], ev[2], ev[3]);
    } else {
        p[i0] = pv[0], m[i0] = mv[0], v[i0] = vv[0];
        if (ema) ema[i0] = ev[0];
    }
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_train_prepare(nlc_ctx* ctx, const float* x0, const float* noise, const float* extra, const float* eta1,
                                 const float* eta2, const float* alpha_bar, int edm, int B, int64_t d, float* noisy_x,
                                 float* new_noise_out, float* dist_real, void* stream) {
    NLC_REQUIRE(ctx && x0 && noise && extra && eta1 && eta2 && alpha_bar && noisy_x && dist_real && B >= 1 && d >= 1,
                "nlc_train_prepare: bad argument");
    train_prepare_kernel<<<B, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        x0, noise, extra, eta1, eta2, alpha_bar, edm, d, static_cast<float>(1.0 / sqrt(static_cast<double>(d))), noisy_x,
        new_noise_out, dist_real);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_adamw_ema_step(nlc_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                  float* ema, int64_t n, double lr, double beta1, double beta2, double eps,
                                  double weight_decay, int64_t step, double ema_rate, double grad_scale, void* stream) {
    NLC_REQUIRE(ctx && params && grads && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "nlc_adamw_ema_step: bad argument");
    if (n == 0) return NLC_OK;
    AdamArgs a;
    a.lr_wd_keep = static_cast<float>(1.0 - lr * weight_decay);
    a.beta1_c = static_cast<float>(1.0 - beta1);
    a.beta2 = static_cast<float>(beta2), a.beta2_c = static_cast<float>(1.0 - beta2);
    a.step_size = static_cast<float>(lr / (1.0 - pow(beta1, static_cast<double>(step))));
    a.sqrt_bc2 = static_cast<float>(sqrt(1.0 - pow(beta2, static_cast<double>(step))));
    a.eps = static_cast<float>(eps), a.grad_scale = static_cast<float>(grad_scale);
    a.ema_rate = static_cast<float>(ema_rate), a.ema_c = static_cast<float>(1.0 - ema_rate);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                                       reinterpret_cast<uintptr_t>(ema)) & 15) == 0;
    if (vec) adamw_ema_kernel<4><<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, ema, n, a);
    else adamw_ema_kernel<1><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, ema, n, a);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
