// SSIM of a finished batch on the device: image_sample.py:571-582 `ssim_fn` = both images rounded to
// uint8, then basicsr `_ssim_3d` (basicsr/metrics/psnr_ssim.py:171-208): ONE 11 x 11 x 11 Gaussian window (sigma 1.5) over
// the [H, W, 3] volume — the colour axis is filtered as well — with replicate padding, and the mean of the SSIM map over
// all H*W*3 entries.  The reference moves every image to the CPU, to numpy, back to the GPU and runs five cuDNN conv3d
// calls per image; here the window is applied separably (rows, columns in shared memory; the three colour planes are
// mixed per pixel by the 3 x 3 matrix the replicate-padded 11-tap filter collapses to), one CTA per 16 x 32 tile, all five
// moments (x, y, x^2, y^2, xy) in one pass: each image is read once.
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace nlc {

constexpr int SS_TW = 32, SS_TH = 16, SS_HALO = 5, SS_WW = SS_TW + 2 * SS_HALO, SS_HH = SS_TH + 2 * SS_HALO;

struct SsimArgs {
    float g[11];  // cv2.getGaussianKernel(11, 1.5)
    float M[9];   // colour-axis filter with replicate padding: M[c][c'] = sum of g[k] with clamp(c + k - 5, 0, 2) == c'
    float c1, c2;
};

__global__ void __launch_bounds__(256) ssim3d_tile_kernel(const float* __restrict__ s, const float* __restrict__ o, int H,
                                                           int W, const SsimArgs a, float* __restrict__ partial) {
    extern __shared__ float sm[];
    float* q1 = sm;                          // [3][HH][WW] sample, rounded to 0..255
    float* q2 = q1 + 3 * SS_HH * SS_WW;      // the same for the ground truth
    float* hp = q2 + 3 * SS_HH * SS_WW;      // [5 moments][3][HH][TW] after the row filter
    const int b = blockIdx.z, y0 = blockIdx.y * SS_TH, x0 = blockIdx.x * SS_TW, tid = threadIdx.x;
    for (int idx = tid; idx < 3 * SS_HH * SS_WW; idx += 256) {
        const int c = idx / (SS_HH * SS_WW), r = (idx / SS_WW) % SS_HH, cc = idx % SS_WW;
        const int gy = min(max(y0 + r - SS_HALO, 0), H - 1), gx = min(max(x0 + cc - SS_HALO, 0), W - 1);
        const size_t off = ((static_cast<size_t>(b) * 3 + c) * H + gy) * W + gx;
        q1[idx] = fminf(fmaxf(rintf(__fmul_rn(s[off], 255.f)), 0.f), 255.f);  // torch.round(x * 255).to(uint8)
        q2[idx] = fminf(fmaxf(rintf(__fmul_rn(o[off], 255.f)), 0.f), 255.f);
    }
    __syncthreads();
    constexpr int PLANE = SS_HH * SS_TW;
    for (int idx = tid; idx < 3 * PLANE; idx += 256) {
        const int c = idx / PLANE, r = (idx / SS_TW) % SS_HH, col = idx % SS_TW;
        const float* p1 = q1 + (c * SS_HH + r) * SS_WW + col;
        const float* p2 = q2 + (c * SS_HH + r) * SS_WW + col;
        float m1 = 0.f, m2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float x = p1[k], y = p2[k], w = a.g[k];
            m1 = fmaf(w, x, m1), m2 = fmaf(w, y, m2);
            m11 = fmaf(w, x * x, m11), m22 = fmaf(w, y * y, m22), m12 = fmaf(w, x * y, m12);
        }
        hp[(0 * 3 + c) * PLANE + r * SS_TW + col] = m1;
        hp[(1 * 3 + c) * PLANE + r * SS_TW + col] = m2;
        hp[(2 * 3 + c) * PLANE + r * SS_TW + col] = m11;
        hp[(3 * 3 + c) * PLANE + r * SS_TW + col] = m22;
        hp[(4 * 3 + c) * PLANE + r * SS_TW + col] = m12;
    }
    __syncthreads();
    float acc = 0.f;
    const int col = tid & 31;
    for (int r = tid >> 5; r < SS_TH; r += 8) {
        if (y0 + r >= H || x0 + col >= W) continue;
        float v[5][3];
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* p = hp + (m * 3 + c) * PLANE + r * SS_TW + col;
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 11; ++k) t = fmaf(a.g[k], p[k * SS_TW], t);
                v[m][c] = t;
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float f[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) f[m] = a.M[c * 3] * v[m][0] + a.M[c * 3 + 1] * v[m][1] + a.M[c * 3 + 2] * v[m][2];
            const float mu1_sq = f[0] * f[0], mu2_sq = f[1] * f[1], mu12 = f[0] * f[1];
            const float s1 = f[2] - mu1_sq, s2 = f[3] - mu2_sq, s12 = f[4] - mu12;
            acc += ((2.f * mu12 + a.c1) * (2.f * s12 + a.c2)) / ((mu1_sq + mu2_sq + a.c1) * (s1 + s2 + a.c2));
        }
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[(static_cast<size_t>(b) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}
// fixed-order sum of the tile sums of one image (double), / (3 H W)
__global__ void ssim3d_finalize_kernel(const float* __restrict__ partial, int tiles, double n, float* __restrict__ out) {
    __shared__ double red[32];
    double t = 0.0;
    for (int i = threadIdx.x; i < tiles; i += blockDim.x) t += partial[static_cast<size_t>(blockIdx.x) * tiles + i];
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) r += red[i];
        out[blockIdx.x] = static_cast<float>(r / n);
    }
}

}  // namespace nlc

using namespace nlc;

static inline int ssim_tiles(int H, int W) { return ((H + SS_TH - 1) / SS_TH) * ((W + SS_TW - 1) / SS_TW); }

extern "C" size_t nlc_ssim3d_ws(int B, int H, int W) { return static_cast<size_t>(B) * ssim_tiles(H, W) * sizeof(float); }

extern "C" int nlc_ssim3d(nlc_ctx* ctx, const float* sample01, const float* orig01, int B, int H, int W, void* workspace,
                          float* ssim_out, void* stream_) {
    NLC_REQUIRE(ctx && sample01 && orig01 && workspace && ssim_out && B >= 1 && H >= 1 && W >= 1,
                "nlc_ssim3d: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    SsimArgs a;
    double g[11], sum = 0.0;
    for (int i = 0; i < 11; ++i) sum += g[i] = exp(-((i - 5.0) * (i - 5.0)) / (2.0 * 1.5 * 1.5));
    for (int i = 0; i < 11; ++i) a.g[i] = static_cast<float>(g[i] /= sum);
    double M[9] = {0};
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 11; ++k) {
            int cp = c + k - 5;
            cp = cp < 0 ? 0 : (cp > 2 ? 2 : cp);
            M[c * 3 + cp] += g[k];
        }
    for (int i = 0; i < 9; ++i) a.M[i] = static_cast<float>(M[i]);
    a.c1 = static_cast<float>((0.01 * 255) * (0.01 * 255)), a.c2 = static_cast<float>((0.03 * 255) * (0.03 * 255));
    const size_t smem = (2 * 3 * SS_HH * SS_WW + 15 * SS_HH * SS_TW) * sizeof(float);
    NLC_CHECK_CUDA(cudaFuncSetAttribute(ssim3d_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    dim3 grid((W + SS_TW - 1) / SS_TW, (H + SS_TH - 1) / SS_TH, B);
    float* partial = static_cast<float*>(workspace);
    ssim3d_tile_kernel<<<grid, 256, smem, st>>>(sample01, orig01, H, W, a, partial);
    NLC_CHECK_LAUNCH();
    ssim3d_finalize_kernel<<<B, 256, 0, st>>>(partial, ssim_tiles(H, W), 3.0 * H * W, ssim_out);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
