// SURVEY section 8f rank 3, second slice: the sigma-model's OWN forward and backward pass of the training step
// (src/experiments.py:683-691: `dist_res = sigma_model(feat)`, `loss = sigma_loss_fn(dist_real, dist_res + 1)`,
// `loss.backward()`), natively in fp32 - the reference leaves it to PyTorch autograd over src/unet_ddim.py:439-529
// (PureResnetBlock, AttnBlock, Downsample, Linear - BatchNorm1d - GELU - Linear).  Activations are NHWC fp32 matrices
// [B*H*W, C]; every contraction - 3x3 / 1x1 convolutions forward, data-gradient and weight-gradient, the attention products,
// the two Linear layers - is one strided batched fp32 GEMM (`nlc_sgemm`, the CUDA-core kernel of operators.cu: the strides
// express every transpose, and with the CHANNEL-major patch order of `unfold3x3` the weight matrix and its gradient are
// torch's own [Cout, Cin*9] layout); the kernels here are what sits between the GEMMs:
//   unfold3x3 / fold3x3          patch matrix of a 3x3 convolution (stride 1 pad 1, or the reference's Downsample:
//                                pad (0,1,0,1) then stride 2) and its adjoint (data gradient back to the image)
//   gn_train_fwd / gn_train_bwd  GroupNorm(32, eps 1e-6) [+ swish] with saved (mean, rstd); dx, dgamma, dbeta
//   softmax_fwd / softmax_bwd    attention rows
//   bias_add, colsum, axpby      bias, bias gradient, residual adds / gradient accumulation
//   bn1d_train_fwd / _bwd        BatchNorm1d in training mode (batch statistics, running-statistics update) [+ GELU(erf)]
//   head_loss                    dist_hat = r + 1, MSE / L1 loss and its gradient
//   nhwc_to_nchw / nchw_to_nhwc  the Flatten() order of the reference (NCHW) around the first Linear
// The sigma-model is <1 % of the step's FLOPs (the frozen UNet encoder is the rest), at 4x4 .. 8x8 pixels: these are
// latency-sized kernels, written for exactness (fp32, deterministic reductions except the atomics of dgamma / dbeta).
#include <math.h>

#include "common.h"
#include "operators.h"
#include "ptx.cuh"

namespace nlc {

// ---------------------------------------------------------------- patches
// x [B,H,W,C] -> P [B*Ho*Wo, C*9], column c*9 + (kh*3 + kw): torch's weight.view(Cout, Cin*9) is the GEMM's right operand.
// down = 0: stride 1, zero padding 1;  down = 1: F.pad(x, (0,1,0,1)) then stride 2, no padding (src/unet_ddim.py:89-94);
// down = 2: stride 2, zero padding 1 (guided-diffusion Downsample, src/unet_adm.py:143-166: the ADM sigma-model).
__global__ void unfold3x3_kernel(const float* __restrict__ x, int B, int H, int W, int C, int down, int Ho, int Wo,
                                 float* __restrict__ P) {
    const long long total = static_cast<long long>(B) * Ho * Wo * C * 9;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int tap = static_cast<int>(i % 9);
        const int c = static_cast<int>((i / 9) % C);
        const long long m = i / (9LL * C);
        const int wo = static_cast<int>(m % Wo), ho = static_cast<int>((m / Wo) % Ho), n = static_cast<int>(m / (Wo * Ho));
        const int kh = tap / 3, kw = tap - kh * 3;
        const int off = down == 1 ? 0 : 1;  // rows 2 ho + kh (down 1), 2 ho + kh - 1 (down 2), ho + kh - 1 (stride 1)
        const int h = (down ? 2 * ho : ho) + kh - off, w = (down ? 2 * wo : wo) + kw - off;
        float v = 0.f;
        if (h >= 0 && h < H && w >= 0 && w < W) v = x[((static_cast<size_t>(n) * H + h) * W + w) * C + c];
        P[i] = v;
    }
}
// adjoint: dx[n,h,w,c] = sum over the (output pixel, tap) pairs that read it of dP
__global__ void fold3x3_kernel(const float* __restrict__ dP, int B, int H, int W, int C, int down, int Ho, int Wo,
                               float* __restrict__ dx, float beta) {
    const long long total = static_cast<long long>(B) * H * W * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long long pix = i / C;
        const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H), n = static_cast<int>(pix / (W * H));
        float acc = 0.f;
        for (int kh = 0; kh < 3; ++kh) {
            int ho;
            if (down) {
                const int q = h - kh + (down == 2 ? 1 : 0);  // 2 ho = h - kh (+ 1 with padding 1)
                if (q < 0 || (q & 1)) continue;
                ho = q >> 1;
            } else {
                ho = h - kh + 1;
            }
            if (ho < 0 || ho >= Ho) continue;
            for (int kw = 0; kw < 3; ++kw) {
                int wo;
                if (down) {
                    const int q = w - kw + (down == 2 ? 1 : 0);
                    if (q < 0 || (q & 1)) continue;
                    wo = q >> 1;
                } else {
                    wo = w - kw + 1;
                }
                if (wo < 0 || wo >= Wo) continue;
                acc += dP[(((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C + c) * 9 + kh * 3 + kw];
            }
        }
        dx[i] = beta == 0.f ? acc : beta * dx[i] + acc;
    }
}

// ---------------------------------------------------------------- GroupNorm (+ swish), training
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) r += red[w];
    __syncthreads();
    return r;
}
// grid (groups, B), 256 threads.  stats[(n*G+g)*2] = mean, +1 = rstd.  act: 0 none, 1 swish
__global__ void __launch_bounds__(256) gn_train_fwd_kernel(const float* __restrict__ x, int HW, int C, int G, float eps,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           int act, float* __restrict__ y, float* __restrict__ stats) {
    __shared__ float red[8];
    const int g = blockIdx.x, n = blockIdx.y, cpg = C / G;
    const int cnt = HW * cpg;
    const float* xb = x + static_cast<size_t>(n) * HW * C + g * cpg;
    float s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s += xb[static_cast<size_t>(i / cpg) * C + i % cpg];
    const float mean = block_sum_256(s, red) / static_cast<float>(cnt);
    float m2 = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const float d = xb[static_cast<size_t>(i / cpg) * C + i % cpg] - mean;
        m2 += d * d;
    }
    const float rstd = rsqrtf(block_sum_256(m2, red) / static_cast<float>(cnt) + eps);
    if (threadIdx.x == 0) stats[(n * G + g) * 2] = mean, stats[(n * G + g) * 2 + 1] = rstd;
    float* yb = y + static_cast<size_t>(n) * HW * C + g * cpg;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int p = i / cpg, c = i % cpg;
        float z = (xb[static_cast<size_t>(p) * C + c] - mean) * rstd * gamma[g * cpg + c] + beta[g * cpg + c];
        if (act) z = z / (1.0f + expf(-z));
        yb[static_cast<size_t>(p) * C + c] = z;
    }
}
// dx (written, or accumulated when accumulate != 0), dgamma / dbeta (atomically accumulated: zero them first)
__global__ void __launch_bounds__(256) gn_train_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int HW,
                                                           int C, int G, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int act,
                                                           const float* __restrict__ stats, float* __restrict__ dx,
                                                           int accumulate, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
    __shared__ float red[8];
    extern __shared__ float ch[];  // [2 * cpg]: per-channel dgamma, dbeta partials of this (sample, group)
    const int g = blockIdx.x, n = blockIdx.y, cpg = C / G;
    const int cnt = HW * cpg;
    const float mean = stats[(n * G + g) * 2], rstd = stats[(n * G + g) * 2 + 1];
    const size_t base = static_cast<size_t>(n) * HW * C + g * cpg;
    for (int c = threadIdx.x; c < 2 * cpg; c += blockDim.x) ch[c] = 0.f;
    __syncthreads();
    // dz = dy * act'(z);  s1 = sum dz*gamma, s2 = sum dz*gamma*xhat over the group
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int p = i / cpg, c = i % cpg;
        const size_t o = base + static_cast<size_t>(p) * C + c;
        const float xh = (x[o] - mean) * rstd;
        float dz = dy[o];
        if (act) {
            const float z = xh * gamma[g * cpg + c] + beta[g * cpg + c];
            const float sg = 1.0f / (1.0f + expf(-z));
            dz *= sg * (1.0f + z * (1.0f - sg));
        }
        atomicAdd(&ch[c], dz * xh);
        atomicAdd(&ch[cpg + c], dz);
        const float dxh = dz * gamma[g * cpg + c];
        s1 += dxh, s2 += dxh * xh;
    }
    s1 = block_sum_256(s1, red) / static_cast<float>(cnt);
    s2 = block_sum_256(s2, red) / static_cast<float>(cnt);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int p = i / cpg, c = i % cpg;
        const size_t o = base + static_cast<size_t>(p) * C + c;
        const float xh = (x[o] - mean) * rstd;
        float dz = dy[o];
        if (act) {
            const float z = xh * gamma[g * cpg + c] + beta[g * cpg + c];
            const float sg = 1.0f / (1.0f + expf(-z));
            dz *= sg * (1.0f + z * (1.0f - sg));
        }
        const float v = rstd * (dz * gamma[g * cpg + c] - s1 - xh * s2);
        dx[o] = accumulate ? dx[o] + v : v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cpg; c += blockDim.x) {
        atomicAdd(&dgamma[g * cpg + c], ch[c]);
        atomicAdd(&dbeta[g * cpg + c], ch[cpg + c]);
    }
}

// ---------------------------------------------------------------- softmax rows (attention), one warp per row
__global__ void softmax_fwd_kernel(const float* __restrict__ s, int rows, int T, float scale, float* __restrict__ p) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* sr = s + static_cast<size_t>(row) * T;
    float m = -INFINITY;
    for (int j = lane; j < T; j += 32) m = fmaxf(m, sr[j] * scale);
    m = warp_max(m);
    float z = 0.f;
    for (int j = lane; j < T; j += 32) z += expf(sr[j] * scale - m);
    z = warp_sum(z);
    for (int j = lane; j < T; j += 32) p[static_cast<size_t>(row) * T + j] = expf(sr[j] * scale - m) / z;
}
// ds = scale * p * (dp - sum_j dp_j p_j)
__global__ void softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, int rows, int T, float scale,
                                   float* __restrict__ ds) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const size_t o = static_cast<size_t>(row) * T;
    float dot = 0.f;
    for (int j = lane; j < T; j += 32) dot += dp[o + j] * p[o + j];
    dot = warp_sum(dot);
    for (int j = lane; j < T; j += 32) ds[o + j] = scale * p[o + j] * (dp[o + j] - dot);
}

// ---------------------------------------------------------------- small elementwise / reductions
__global__ void bias_add_kernel(float* __restrict__ y, const float* __restrict__ b, long long rows, int C) {
    const long long total = rows * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        y[i] += b[i % C];
}
// out[c] = sum_rows x[r, c]; grid ceil(C/32) blocks of (32, 8): deterministic
__global__ void colsum_kernel(const float* __restrict__ x, long long rows, int C, float* __restrict__ out) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < C)
        for (long long r = threadIdx.y; r < rows; r += 8) acc += x[r * C + c];
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        out[c] = t;
    }
}
__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, const float* __restrict__ y,
                             float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.f);
}
// [B, HW, C] <-> [B, C, HW]
__global__ void permute_kernel(const float* __restrict__ x, int B, int HW, int C, int to_nchw, float* __restrict__ y) {
    const long long total = static_cast<long long>(B) * HW * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C), p = static_cast<int>((i / C) % HW), n = static_cast<int>(i / (static_cast<long long>(C) * HW));
        const size_t nhwc = i, nchw = (static_cast<size_t>(n) * C + c) * HW + p;
        if (to_nchw) y[nchw] = x[nhwc]; else y[nhwc] = x[nchw];
    }
}

// ---------------------------------------------------------------- BatchNorm1d (training) + GELU(erf): one block per feature
// x [B, F] -> y = gelu(bn(x)); saves xhat-statistics (mean, rstd) and updates running stats with momentum (unbiased variance)
__global__ void __launch_bounds__(128) bn1d_train_fwd_kernel(const float* __restrict__ x, int B, int F, float eps,
                                                             float momentum, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ run_mean,
                                                             float* __restrict__ run_var, float* __restrict__ stats,
                                                             float* __restrict__ y, int act) {
    __shared__ float red[8];
    const int f = blockIdx.x;
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += x[static_cast<size_t>(b) * F + f];
    const float mean = block_sum_256(s, red) / static_cast<float>(B);
    float m2 = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float d = x[static_cast<size_t>(b) * F + f] - mean;
        m2 += d * d;
    }
    m2 = block_sum_256(m2, red);
    const float var = m2 / static_cast<float>(B);
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        stats[2 * f] = mean, stats[2 * f + 1] = rstd;
        if (run_mean) {
            run_mean[f] = (1.0f - momentum) * run_mean[f] + momentum * mean;
            run_var[f] = (1.0f - momentum) * run_var[f] + momentum * (B > 1 ? m2 / static_cast<float>(B - 1) : var);
        }
    }
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float z = (x[static_cast<size_t>(b) * F + f] - mean) * rstd * gamma[f] + beta[f];
        // act 0: GELU(erf) (the DDIM / ADM sigma-models), 1: SiLU (the EDM sigma-model, src/edm_networks.py:1006-1010)
        y[static_cast<size_t>(b) * F + f] = act == 0 ? 0.5f * z * (1.0f + erff(z * 0.70710678118654752f)) : z / (1.0f + expf(-z));
    }
}
// derivative of the head's activation at z
__device__ __forceinline__ float head_act_grad(float z, int act) {
    if (act == 0) return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
    const float sg = 1.0f / (1.0f + expf(-z));
    return sg * (1.0f + z * (1.0f - sg));
}
// dy: gradient wrt gelu output -> dx, dgamma[f], dbeta[f]
__global__ void __launch_bounds__(128) bn1d_train_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B,
                                                             int F, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const float* __restrict__ stats,
                                                             float* __restrict__ dx, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int act) {
    __shared__ float red[8];
    const int f = blockIdx.x;
    const float mean = stats[2 * f], rstd = stats[2 * f + 1], g = gamma[f], bt = beta[f];
    float sg = 0.f, sb = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float xh = (x[static_cast<size_t>(b) * F + f] - mean) * rstd;
        const float z = xh * g + bt;
        // gelu'(z) = Phi(z) + z phi(z);  silu'(z) = s (1 + z (1 - s))
        const float dz = dy[static_cast<size_t>(b) * F + f] * head_act_grad(z, act);
        sg += dz * xh, sb += dz;
    }
    sg = block_sum_256(sg, red), sb = block_sum_256(sb, red);
    if (threadIdx.x == 0) dgamma[f] = sg, dbeta[f] = sb;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float xh = (x[static_cast<size_t>(b) * F + f] - mean) * rstd;
        const float z = xh * g + bt;
        const float dz = dy[static_cast<size_t>(b) * F + f] * head_act_grad(z, act);
        dx[static_cast<size_t>(b) * F + f] = rstd * g * (dz - sb / static_cast<float>(B) - xh * sg / static_cast<float>(B));
    }
}

// dist_hat = r + 1; kind 0: MSELoss(mean), 1: L1Loss(mean).  loss[0] = value, dr[b] = d loss / d r[b].  One block.
// weight == NULL: mean reduction; else loss = sum_b w_b l_b / sum_b w_b (the EDM loop's `loss_weighted`, src/experiments.py:1019-1021)
__global__ void __launch_bounds__(256) head_loss_kernel(const float* __restrict__ r, const float* __restrict__ target, int B,
                                                        int kind, const float* __restrict__ weight, float* __restrict__ dist_hat,
                                                        float* __restrict__ loss, float* __restrict__ dr) {
    __shared__ float red[8];
    float wsum = static_cast<float>(B);
    if (weight) {
        float a = 0.f;
        for (int b = threadIdx.x; b < B; b += blockDim.x) a += weight[b];
        wsum = block_sum_256(a, red);
        __syncthreads();
    }
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float dh = r[b] + 1.0f;
        dist_hat[b] = dh;
        const float e = dh - target[b];
        const float w = (weight ? weight[b] : 1.0f) / wsum;
        if (kind == 0) {
            acc += w * (e * e);
            dr[b] = 2.0f * e * w;
        } else {
            acc += w * fabsf(e);
            dr[b] = (e > 0.f ? 1.0f : (e < 0.f ? -1.0f : 0.f)) * w;
        }
    }
    acc = block_sum_256(acc, red);
    if (threadIdx.x == 0) loss[0] = acc;
}

static unsigned grid1d(long long n, int sm_count) {
    long long g = (n + 255) / 256;
    const long long cap = 16LL * sm_count;
    return static_cast<unsigned>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_sgemm(nlc_ctx* ctx, int batch, int M, int N, int K, const float* A, long long sab, long long sai,
                         long long sak, const float* Bm, long long sbb, long long sbk, long long sbj, float* Cm,
                         const float* add, void* stream_) {
    NLC_REQUIRE(ctx && A && Bm && Cm && batch >= 1 && M >= 1 && N >= 1 && K >= 1, "nlc_sgemm: bad arguments");
    return launch_gemm(static_cast<cudaStream_t>(stream_), batch, M, N, K, A, sab, sai, sak, Bm, sbb, sbk, sbj, Cm, nullptr, 1,
                       nullptr, nullptr, 1.f, -1.f, add);
}

extern "C" int nlc_unfold3x3(nlc_ctx* ctx, const float* x, int B, int H, int W, int C, int down, float* P, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && P && down >= 0 && down <= 2, "nlc_unfold3x3: bad arguments");
    const int Ho = down ? (H + down - 3) / 2 + 1 : H, Wo = down ? (W + down - 3) / 2 + 1 : W;
    unfold3x3_kernel<<<grid1d(static_cast<long long>(B) * Ho * Wo * C * 9, ctx->sm_count), 256, 0, st>>>(x, B, H, W, C, down, Ho,
                                                                                                       Wo, P);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_fold3x3(nlc_ctx* ctx, const float* dP, int B, int H, int W, int C, int down, float* dx, float beta,
                           void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && dP && dx && down >= 0 && down <= 2, "nlc_fold3x3: bad arguments");
    const int Ho = down ? (H + down - 3) / 2 + 1 : H, Wo = down ? (W + down - 3) / 2 + 1 : W;
    fold3x3_kernel<<<grid1d(static_cast<long long>(B) * H * W * C, ctx->sm_count), 256, 0, st>>>(dP, B, H, W, C, down, Ho, Wo, dx,
                                                                                                 beta);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_gn_train_fwd(nlc_ctx* ctx, const float* x, int B, int HW, int C, int groups, float eps, const float* gamma,
                                const float* beta, int act, float* y, float* stats, void* stream_) {
    NLC_REQUIRE(ctx && x && y && stats && gamma && beta && groups >= 1 && C % groups == 0, "nlc_gn_train_fwd: bad arguments");
    gn_train_fwd_kernel<<<dim3(groups, B), 256, 0, static_cast<cudaStream_t>(stream_)>>>(x, HW, C, groups, eps, gamma, beta, act,
                                                                                        y, stats);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_gn_train_bwd(nlc_ctx* ctx, const float* x, const float* dy, int B, int HW, int C, int groups,
                                const float* gamma, const float* beta, int act, const float* stats, float* dx, int accumulate,
                                float* dgamma, float* dbeta, void* stream_) {
    NLC_REQUIRE(ctx && x && dy && dx && stats && dgamma && dbeta && groups >= 1 && C % groups == 0,
                "nlc_gn_train_bwd: bad arguments");
    const size_t smem = 2 * static_cast<size_t>(C / groups) * sizeof(float);
    gn_train_bwd_kernel<<<dim3(groups, B), 256, smem, static_cast<cudaStream_t>(stream_)>>>(x, dy, HW, C, groups, gamma, beta, act,
                                                                                           stats, dx, accumulate, dgamma, dbeta);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_softmax_rows(nlc_ctx* ctx, const float* s, const float* dp, int rows, int T, float scale, float* out,
                                void* stream_) {
    NLC_REQUIRE(ctx && s && out && rows >= 1 && T >= 1, "nlc_softmax_rows: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (dp)  // backward: s holds the probabilities
        softmax_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s, dp, rows, T, scale, out);
    else
        softmax_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(s, rows, T, scale, out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_bias_add(nlc_ctx* ctx, float* y, const float* bias, long long rows, int C, void* stream_) {
    NLC_REQUIRE(ctx && y && bias && rows >= 1 && C >= 1, "nlc_bias_add: bad arguments");
    bias_add_kernel<<<grid1d(rows * C, ctx->sm_count), 256, 0, static_cast<cudaStream_t>(stream_)>>>(y, bias, rows, C);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_colsum(nlc_ctx* ctx, const float* x, long long rows, int C, float* out, void* stream_) {
    NLC_REQUIRE(ctx && x && out && rows >= 1 && C >= 1, "nlc_colsum: bad arguments");
    colsum_kernel<<<(C + 31) / 32, dim3(32, 8), 0, static_cast<cudaStream_t>(stream_)>>>(x, rows, C, out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_axpby(nlc_ctx* ctx, float a, const float* x, float b, const float* y, float* out, long long n,
                         void* stream_) {
    NLC_REQUIRE(ctx && x && out && n >= 1, "nlc_axpby: bad arguments");
    axpby_kernel<<<grid1d(n, ctx->sm_count), 256, 0, static_cast<cudaStream_t>(stream_)>>>(a, x, b, y, out, n);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_permute_nhwc(nlc_ctx* ctx, const float* x, int B, int HW, int C, int to_nchw, float* y, void* stream_) {
    NLC_REQUIRE(ctx && x && y, "nlc_permute_nhwc: bad arguments");
    permute_kernel<<<grid1d(static_cast<long long>(B) * HW * C, ctx->sm_count), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, B, HW, C, to_nchw, y);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_bn1d_act_train(nlc_ctx* ctx, const float* x, const float* dy, int B, int F, float eps, float momentum, int act,
                                  const float* gamma, const float* beta, float* run_mean, float* run_var, float* stats,
                                  float* out, float* dgamma, float* dbeta, void* stream_) {
    NLC_REQUIRE(ctx && x && gamma && beta && stats && out && B >= 1 && F >= 1 && (act == 0 || act == 1),
                "nlc_bn1d_act_train: bad arguments (act 0 = GELU, 1 = SiLU)");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (dy) {
        NLC_REQUIRE(dgamma && dbeta, "nlc_bn1d_act_train: backward needs dgamma / dbeta");
        bn1d_train_bwd_kernel<<<F, 128, 0, st>>>(x, dy, B, F, gamma, beta, stats, out, dgamma, dbeta, act);
    } else {
        bn1d_train_fwd_kernel<<<F, 128, 0, st>>>(x, B, F, eps, momentum, gamma, beta, run_mean, run_var, stats, out, act);
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_bn1d_gelu_train(nlc_ctx* ctx, const float* x, const float* dy, int B, int F, float eps, float momentum,
                                   const float* gamma, const float* beta, float* run_mean, float* run_var, float* stats,
                                   float* out, float* dgamma, float* dbeta, void* stream_) {
    return nlc_bn1d_act_train(ctx, x, dy, B, F, eps, momentum, 0, gamma, beta, run_mean, run_var, stats, out, dgamma, dbeta,
                              stream_);
}

extern "C" int nlc_head_loss(nlc_ctx* ctx, const float* r, const float* target, int B, int kind, float* dist_hat, float* loss,
                             float* dr, void* stream_) {
    NLC_REQUIRE(ctx && r && target && dist_hat && loss && dr && B >= 1 && (kind == 0 || kind == 1),
                "nlc_head_loss: kind 0 (MSE) or 1 (L1)");
    head_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(r, target, B, kind, nullptr, dist_hat, loss, dr);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

extern "C" int nlc_head_loss_weighted(nlc_ctx* ctx, const float* r, const float* target, const float* weight, int B, int kind,
                                      float* dist_hat, float* loss, float* dr, void* stream_) {
    NLC_REQUIRE(ctx && r && target && weight && dist_hat && loss && dr && B >= 1 && (kind == 0 || kind == 1),
                "nlc_head_loss_weighted: kind 0 (MSE) or 1 (L1)");
    head_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream_)>>>(r, target, B, kind, weight, dist_hat, loss, dr);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
