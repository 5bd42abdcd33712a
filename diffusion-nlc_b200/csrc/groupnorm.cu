// GroupNorm statistics + fused apply (normalise, affine, optional per-sample scale/shift, optional SiLU,
// cast to the tensor-core operand dtype).  HBM-bound: the activation is read twice (stats, apply) and the
// operand written once; everything else stays in registers / shared memory.
//
// Reference semantics: torch.nn.GroupNorm(32, C, eps) in fp32 followed by x*sigmoid(x)
// (src/unet_ddim.py:54-55,139-146; src/nn_util.py:17-19,93-100; src/edm_networks.py:105-116) and ADM's
// use_scale_shift_norm h = norm(h)*(1+scale)+shift (src/unet_adm.py:248-252).
//
// Statistics are Welford/Chan (count, mean, M2) partials per (sample, pixel chunk, group) — no E[x^2]-E[x]^2
// cancellation — written to a small workspace and merged by every CTA of the apply pass.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 64;

struct Wf {
    float n, mean, m2;
};
__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
    if (b.n == 0.f) return a;
    if (a.n == 0.f) return b;
    const float n = a.n + b.n;
    const float delta = b.mean - a.mean;
    const float f = b.n / n;
    Wf r;
    r.n = n;
    r.mean = a.mean + delta * f;
    r.m2 = a.m2 + b.m2 + delta * delta * a.n * f;
    return r;
}
__device__ __forceinline__ Wf wf_of4(float4 v) {
    Wf r;
    r.n = 4.f;
    r.mean = 0.25f * ((v.x + v.y) + (v.z + v.w));
    const float a = v.x - r.mean, b = v.y - r.mean, c = v.z - r.mean, d = v.w - r.mean;
    r.m2 = (a * a + b * b) + (c * c + d * d);
    return r;
}

// grid (nchunks, B); dynamic smem: entries * sizeof(Wf)
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const float* __restrict__ x, int ld_x, int HW, int C,
                                                               int groups, int rows_per_chunk, float* __restrict__ ws) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ float gn_smem[];
    Wf* part = reinterpret_cast<Wf*>(gn_smem);
    const int C4 = C >> 2;
    const int row_lanes = C4 >= kGnThreads ? 1 : kGnThreads / C4;
    const int entries = row_lanes * C4;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const float* xb = x + (static_cast<size_t>(n) * HW + static_cast<size_t>(chunk) * rows_per_chunk) * ld_x;

    for (int e = threadIdx.x; e < entries; e += kGnThreads) {
        const int rl = e / C4, c4 = e - rl * C4;
        Wf acc = {0.f, 0.f, 0.f};
        const float* col = xb + 4 * c4;
        int r = rl;
        // two independent accumulators in flight to hide load latency
        Wf acc2 = {0.f, 0.f, 0.f};
        for (; r + row_lanes < rows_per_chunk; r += 2 * row_lanes) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(col + static_cast<size_t>(r) * ld_x));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(col + static_cast<size_t>(r + row_lanes) * ld_x));
            acc = wf_merge(acc, wf_of4(v0));
            acc2 = wf_merge(acc2, wf_of4(v1));
        }
        if (r < rows_per_chunk)
            acc = wf_merge(acc, wf_of4(__ldg(reinterpret_cast<const float4*>(col + static_cast<size_t>(r) * ld_x))));
        part[e] = wf_merge(acc, acc2);
    }
    __syncthreads();
    const int cpg4 = (C / groups) >> 2;
    for (int g = threadIdx.x; g < groups; g += kGnThreads) {
        Wf acc = {0.f, 0.f, 0.f};
        for (int rl = 0; rl < row_lanes; ++rl)
            for (int j = 0; j < cpg4; ++j) acc = wf_merge(acc, part[rl * C4 + g * cpg4 + j]);
        float* o = ws + ((static_cast<size_t>(n) * gridDim.x + chunk) * groups + g) * 3;
        o[0] = acc.n, o[1] = acc.mean, o[2] = acc.m2;
    }
}

// ---------------------------------------------------------------- statistics -> (mean, rstd) per (sample, group)
// (a) from the Welford partials of gn_stats_kernel; grid (B), one thread per group
__global__ void gn_finalize_welford_kernel(const float* __restrict__ ws, int stat_chunks, int groups, float eps,
                                           float* __restrict__ mr) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int n = blockIdx.x;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        Wf acc = {0.f, 0.f, 0.f};
        for (int k = 0; k < stat_chunks; ++k) {
            const float* o = ws + ((static_cast<size_t>(n) * stat_chunks + k) * groups + g) * 3;
            acc = wf_merge(acc, Wf{o[0], o[1], o[2]});
        }
        mr[(n * groups + g) * 2] = acc.mean;
        mr[(n * groups + g) * 2 + 1] = rsqrtf(acc.m2 / acc.n + eps);
    }
}

__device__ __forceinline__ float gn_block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    v = threadIdx.x < (kGnThreads >> 5) ? red[threadIdx.x] : 0.f;
    if (warp == 0) {
        v = warp_sum(v);
        if (lane == 0) red[0] = v;
    }
    __syncthreads();
    v = red[0];
    __syncthreads();
    return v;
}

// (b) from the (32 pixel x 4 channel) partials written by the producing conv's epilogue (conv_tc.cu, gn_partials):
// all partials have the same count (128), so the merge is "mean of means" + Chan's between-block term, evaluated in
// two passes over the (L2-resident) partials.  grid (groups, B)
__global__ void __launch_bounds__(kGnThreads)
    gn_finalize_blocks_kernel(const float* __restrict__ stats, int stats_nblk, int n_rowgroups, int blocks_per_group,
                              float eps, float* __restrict__ mr) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ float red[kGnThreads / 32];
    const int g = blockIdx.x, n = blockIdx.y;
    const float2* base = reinterpret_cast<const float2*>(stats) + static_cast<size_t>(n) * n_rowgroups * stats_nblk +
                         g * blocks_per_group;
    const int total = n_rowgroups * blocks_per_group;
    float s = 0.f;
    for (int i = threadIdx.x; i < total; i += kGnThreads) {
        const int rg = i / blocks_per_group, k = i - rg * blocks_per_group;
        s += __ldg(base + static_cast<size_t>(rg) * stats_nblk + k).x;
    }
    const float mean = gn_block_sum(s, red) / static_cast<float>(total);
    float m2 = 0.f;
    for (int i = threadIdx.x; i < total; i += kGnThreads) {
        const int rg = i / blocks_per_group, k = i - rg * blocks_per_group;
        const float2 p = __ldg(base + static_cast<size_t>(rg) * stats_nblk + k);
        const float d = p.x - mean;
        m2 += p.y + 128.0f * d * d;
    }
    m2 = gn_block_sum(m2, red);
    if (threadIdx.x == 0) {
        mr[(n * gridDim.x + g) * 2] = mean;
        mr[(n * gridDim.x + g) * 2 + 1] = rsqrtf(m2 / (128.0f * static_cast<float>(total)) + eps);
    }
}

// Same merge with one warp per (sample, group) for small images (<= 512 partials per group: no block barriers, eight
// groups per CTA instead of one CTA each; at 64x64 the block version above was 3 % of a c2 step in launch and barrier
// latency).  grid ceil(B * groups / 8)
__global__ void __launch_bounds__(256)
    gn_finalize_blocks_warp_kernel(const float* __restrict__ stats, int stats_nblk, int n_rowgroups, int blocks_per_group,
                                   int groups, int n_pairs, float eps, float* __restrict__ mr) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int n = pair / groups, g = pair - n * groups;
    const float2* base = reinterpret_cast<const float2*>(stats) + static_cast<size_t>(n) * n_rowgroups * stats_nblk +
                         g * blocks_per_group;
    const int total = n_rowgroups * blocks_per_group;
    float s = 0.f;
    for (int i = lane; i < total; i += 32) {
        const int rg = i / blocks_per_group, k = i - rg * blocks_per_group;
        s += __ldg(base + static_cast<size_t>(rg) * stats_nblk + k).x;
    }
    const float mean = warp_sum(s) / static_cast<float>(total);
    float m2 = 0.f;
    for (int i = lane; i < total; i += 32) {
        const int rg = i / blocks_per_group, k = i - rg * blocks_per_group;
        const float2 p = __ldg(base + static_cast<size_t>(rg) * stats_nblk + k);
        const float d = p.x - mean;
        m2 += p.y + 128.0f * d * d;
    }
    m2 = warp_sum(m2);
    if (lane == 0) {
        mr[pair * 2] = mean;
        mr[pair * 2 + 1] = rsqrtf(m2 / (128.0f * static_cast<float>(total)) + eps);
    }
}

// ---------------------------------------------------------------- apply
// `rnd` is the operand format flag (common.h dtype_fmt): fp32 containers -> round to tf32; 16-bit -> fp16, not bf16
template <bool TF32>
__device__ __forceinline__ void gn_store8(void* __restrict__ y, size_t off, const float (&f)[8], int rnd) {
    if (TF32) {
        float* yp = static_cast<float*>(y) + off;
        reinterpret_cast<float4*>(yp)[0] =
            make_float4(op_f32(f[0], rnd), op_f32(f[1], rnd), op_f32(f[2], rnd), op_f32(f[3], rnd));
        reinterpret_cast<float4*>(yp)[1] =
            make_float4(op_f32(f[4], rnd), op_f32(f[5], rnd), op_f32(f[6], rnd), op_f32(f[7], rnd));
    } else {
        __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y) + off;
        *reinterpret_cast<uint4*>(yp) = make_uint4(pack_op16x2(f[0], f[1], rnd), pack_op16x2(f[2], f[3], rnd),
                                                   pack_op16x2(f[4], f[5], rnd), pack_op16x2(f[6], f[7], rnd));
    }
}

__device__ __forceinline__ void gn_act8(const float4 v0, const float4 v1, const float* __restrict__ ca,
                                        const float* __restrict__ cb, int do_silu, float (&f)[8]) {
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const float4 a0 = *reinterpret_cast<const float4*>(ca), a1 = *reinterpret_cast<const float4*>(ca + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(cb), b1 = *reinterpret_cast<const float4*>(cb + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        f[k] = fmaf(x[k], a[k], b[k]);
        if (do_silu) f[k] = silu(f[k]);
    }
}

// (kGnUnroll independent 32-byte loads per thread, two CTAs per SM: ~128 KB of reads in flight per SM)
// MODE 0: y[pix] = act(x[pix]);  MODE 1: the 2x2 outputs of an input pixel get its value (nearest x2);
// MODE 2: y[opix] = mean of act over the 2x2 input window (avg_pool2d of the activated tensor).
// grid (chunks, B); dynamic smem 2*C floats: y = act(x*ca[c] + cb[c]).  Four independent 32-byte loads in flight
// per thread.
constexpr int kGnUnrollMax = 8;
// X16: the input is a 16-bit tensor in the operand dtype (the activation between a ResBlock's two convolutions, which in the
// 16-bit modes is written by the first conv's epilogue in the operand dtype only, its statistics coming from the fp32
// accumulators; in the 16-bit-activation plans every GroupNorm input at a level with >= 128 pixels): 16-byte loads of 8
// channels instead of two 16-byte fp32 loads.
template <bool TF32, int MODE, bool X16 = false>
__global__ void __launch_bounds__(kGnThreads, MODE == 2 ? 1 : 2)
    gn_apply_kernel(const float* __restrict__ x, int ld_x, int H, int W, int C, int groups,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                    const float* __restrict__ shift, int ld_ss, int do_silu, const float* __restrict__ mr,
                    void* __restrict__ y, int ld_y, int items_per_chunk, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    constexpr int kGnUnroll = MODE == 2 ? 2 : kGnUnrollMax;  // MODE 2 items are four times as wide
    extern __shared__ __align__(16) float gn_smem[];
    float* ca = gn_smem;
    float* cb = gn_smem + C;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int cpg = C / groups;
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        const int g = c / cpg;
        const float mean = mr[(n * groups + g) * 2], rstd = mr[(n * groups + g) * 2 + 1];
        float a = rstd * (gamma ? gamma[c] : 1.f);
        float b = (beta ? beta[c] : 0.f) - mean * a;
        if (scale) {
            const float sc = 1.f + scale[static_cast<size_t>(n) * ld_ss + c];
            a *= sc;
            b = b * sc + shift[static_cast<size_t>(n) * ld_ss + c];
        }
        ca[c] = a, cb[c] = b;
    }
    __syncthreads();
    const int C8 = C >> 3;
    // work items = (pixel, 8-channel block); pixels are input pixels (MODE 0, 1) or output pixels (MODE 2)
    const int Wp = MODE == 2 ? W >> 1 : W;
    const int Hp = MODE == 2 ? H >> 1 : H;
    // (items of one image fit 32 bits: the host checks H*W*C/8 < 2^31)
    const int npix = Hp * Wp;
    const int item0 = chunk * items_per_chunk;
    int item_end = item0 + items_per_chunk;
    if (item_end > npix * C8) item_end = npix * C8;
    const size_t img_in = static_cast<size_t>(n) * H * W;
    // (pixel, channel block) of each unrolled item, advanced incrementally: one integer division per thread, not per item
    int cc[kGnUnroll];
    int pp[kGnUnroll];
#pragma unroll
    for (int u = 0; u < kGnUnroll; ++u) {
        const int i = item0 + static_cast<int>(threadIdx.x) + u * kGnThreads;
        pp[u] = i / C8;
        cc[u] = (i - pp[u] * C8) << 3;
    }
    const int step_pix = (kGnUnroll * kGnThreads) / C8;
    const int step_c = ((kGnUnroll * kGnThreads) - step_pix * C8) << 3;
    for (int i0 = item0 + threadIdx.x; i0 < item_end; i0 += kGnUnroll * kGnThreads) {
        float4 v[kGnUnroll][MODE == 2 ? 8 : 2];
#pragma unroll
        for (int u = 0; u < kGnUnroll; ++u) {
            const int i = i0 + u * kGnThreads;
            if (i < item_end) {
                const int pix = pp[u];
                const int c = cc[u];
                if (MODE == 2) {
                    const int ho = pix / Wp, wo = pix - ho * Wp;
                    const float* xp = x + (img_in + static_cast<size_t>(2 * ho) * W + 2 * wo) * ld_x + c;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (X16) {
                            const __nv_bfloat16* xq = reinterpret_cast<const __nv_bfloat16*>(x) +
                                (img_in + static_cast<size_t>(2 * ho + (q >> 1)) * W + 2 * wo + (q & 1)) * ld_x + c;
                            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xq));
                            v[u][q] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                                                  __uint_as_float(raw.w));
                        } else {
                            const float* xq = xp + (static_cast<size_t>(q >> 1) * W + (q & 1)) * ld_x;
                            v[u][2 * q] = __ldg(reinterpret_cast<const float4*>(xq));
                            v[u][2 * q + 1] = __ldg(reinterpret_cast<const float4*>(xq + 4));
                        }
                    }
                } else if (X16) {
                    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x) + (img_in + pix) * ld_x + c;
                    // (raw bits only: unpacking here would make every load wait for its own data before the next issues)
                    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(xp));
                    v[u][0] = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z),
                                          __uint_as_float(raw.w));
                } else {
                    const float* xp = x + (img_in + pix) * ld_x + c;
                    v[u][0] = __ldg(reinterpret_cast<const float4*>(xp));
                    v[u][1] = __ldg(reinterpret_cast<const float4*>(xp + 4));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kGnUnroll; ++u) {
            const int i = i0 + u * kGnThreads;
            if (i < item_end) {
                const int c = cc[u];
                const int pix = pp[u];
                float f[8];
                if (MODE == 2) {
                    float t[4][8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 lo, hi;
                        if (X16)
                            unpack_op16x8(make_uint4(__float_as_uint(v[u][q].x), __float_as_uint(v[u][q].y),
                                                     __float_as_uint(v[u][q].z), __float_as_uint(v[u][q].w)),
                                          rnd, lo, hi);
                        else
                            lo = v[u][2 * q], hi = v[u][2 * q + 1];
                        gn_act8(lo, hi, ca + c, cb + c, do_silu, t[q]);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = ((t[0][k] + t[1][k]) + (t[2][k] + t[3][k])) * 0.25f;
                    gn_store8<TF32>(y, (static_cast<size_t>(n) * npix + pix) * ld_y + c, f, rnd);
                } else {
                    if (X16)
                        unpack_op16x8(make_uint4(__float_as_uint(v[u][0].x), __float_as_uint(v[u][0].y),
                                                 __float_as_uint(v[u][0].z), __float_as_uint(v[u][0].w)),
                                      rnd, v[u][0], v[u][1]);
                    gn_act8(v[u][0], v[u][1], ca + c, cb + c, do_silu, f);
                    if (MODE == 0) {
                        gn_store8<TF32>(y, (img_in + pix) * ld_y + c, f, rnd);
                    } else {
                        const int h = pix / W, w = pix - h * W;
                        const size_t o = (static_cast<size_t>(n) * 4 * H * W + static_cast<size_t>(2 * h) * 2 * W + 2 * w);
                        gn_store8<TF32>(y, o * ld_y + c, f, rnd);
                        gn_store8<TF32>(y, (o + 1) * ld_y + c, f, rnd);
                        gn_store8<TF32>(y, (o + 2 * W) * ld_y + c, f, rnd);
                        gn_store8<TF32>(y, (o + 2 * W + 1) * ld_y + c, f, rnd);
                    }
                }
            }
            cc[u] += step_c, pp[u] += step_pix;
            if (cc[u] >= C) cc[u] -= C, ++pp[u];
        }
    }
}

// ---------------------------------------------------------------- apply, 16-bit output, same resolution: the lean kernel
// ncu on the generic kernel above (16-bit input, 64x64x128, batch 256; profiles/r02j_ncu_gn.md): 168 warp instructions per
// 8-channel item - per-item index arithmetic, coefficient loads from shared memory, predication of eight unrolled items - at
// 25 % occupancy: issue slots 57 % busy, no eligible warp 43 % of the cycles, DRAM at 42 % of peak.  Here a thread owns ONE
// 8-channel block for its whole pixel range, so the 16 affine coefficients (statistics, gamma / beta, optional per-sample
// scale / shift, and for SiLU the pre-multiplied -log2(e) copies) live in registers, a pixel step is one pointer
// increment, and U loads are in flight per thread: ~65 instructions per item (48 of them the SiLU), 4 CTAs per SM.
// grid (chunks, B), block = C/8 * rows threads (rows = 256 / (C/8) pixels per pass).
template <bool X16, bool SILU, int U>
__global__ void __launch_bounds__(256, 4)
    gn_apply_lean_kernel(const void* __restrict__ x_, int ld_x, int HW, int C, int groups, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ scale, const float* __restrict__ shift,
                         int ld_ss, const float* __restrict__ mr, __nv_bfloat16* __restrict__ y, int ld_y,
                         int pix_per_chunk, int f16) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int C8 = C >> 3;
    const int n = blockIdx.y;
    const int cblk = static_cast<int>(threadIdx.x) % C8, prow = static_cast<int>(threadIdx.x) / C8;
    const int rows = static_cast<int>(blockDim.x) / C8;
    const int c0 = cblk << 3;
    const int cpg = C / groups;
    float a[8], b[8], a2[8], b2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + k, g = c / cpg;
        const float mean = __ldg(mr + (n * groups + g) * 2), rstd = __ldg(mr + (n * groups + g) * 2 + 1);
        float ak = rstd * (gamma ? __ldg(gamma + c) : 1.f);
        float bk = (beta ? __ldg(beta + c) : 0.f) - mean * ak;
        if (scale) {
            const float sc = 1.f + __ldg(scale + static_cast<size_t>(n) * ld_ss + c);
            ak *= sc;
            bk = bk * sc + __ldg(shift + static_cast<size_t>(n) * ld_ss + c);
        }
        a[k] = ak, b[k] = bk;
        a2[k] = ak * -1.4426950408889634f, b2[k] = bk * -1.4426950408889634f;  // exp(-v) = 2^(x a2 + b2)
    }
    const int p0 = blockIdx.x * pix_per_chunk;
    int p1 = p0 + pix_per_chunk;
    if (p1 > HW) p1 = HW;
    const size_t img = static_cast<size_t>(n) * HW;
    const size_t step_x = static_cast<size_t>(rows) * ld_x, step_y = static_cast<size_t>(rows) * ld_y;
    const __nv_bfloat16* x16 = static_cast<const __nv_bfloat16*>(x_) + (img + p0 + prow) * ld_x + c0;
    const float* x32 = static_cast<const float*>(x_) + (img + p0 + prow) * ld_x + c0;
    __nv_bfloat16* yp = y + (img + p0 + prow) * ld_y + c0;
    for (int p = p0 + prow; p < p1; p += rows * U) {
        uint4 raw[U][X16 ? 1 : 2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * rows < p1) {
                if (X16) {
                    raw[u][0] = __ldg(reinterpret_cast<const uint4*>(x16 + u * step_x));
                } else {
                    raw[u][0] = __ldg(reinterpret_cast<const uint4*>(x32 + u * step_x));
                    raw[u][X16 ? 0 : 1] = __ldg(reinterpret_cast<const uint4*>(x32 + u * step_x) + 1);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (p + u * rows < p1) {
                float v[8];
                if (X16) {
                    float4 lo, hi;
                    unpack_op16x8(raw[u][0], f16, lo, hi);
                    v[0] = lo.x, v[1] = lo.y, v[2] = lo.z, v[3] = lo.w, v[4] = hi.x, v[5] = hi.y, v[6] = hi.z, v[7] = hi.w;
                } else {
                    const uint4 q0 = raw[u][0], q1 = raw[u][X16 ? 0 : 1];
                    v[0] = __uint_as_float(q0.x), v[1] = __uint_as_float(q0.y), v[2] = __uint_as_float(q0.z);
                    v[3] = __uint_as_float(q0.w), v[4] = __uint_as_float(q1.x), v[5] = __uint_as_float(q1.y);
                    v[6] = __uint_as_float(q1.z), v[7] = __uint_as_float(q1.w);
                }
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    f[k] = fmaf(v[k], a[k], b[k]);
                    if (SILU) f[k] = __fdividef(f[k], 1.0f + fast_exp2(fmaf(v[k], a2[k], b2[k])));
                }
                *reinterpret_cast<uint4*>(yp + u * step_y) =
                    make_uint4(pack_op16x2(f[0], f[1], f16), pack_op16x2(f[2], f[3], f16), pack_op16x2(f[4], f[5], f16),
                               pack_op16x2(f[6], f[7], f16));
            }
        }
        x16 += U * step_x, x32 += U * step_x, yp += U * step_y;
    }
}

// (returns false when the shape is not served: more than 256 channel blocks, i.e. C > 2048)
static bool launch_apply_lean(const void* x, int x_is_op, int ld_x, int B, int HW, int C, int groups, const float* gamma,
                              const float* beta, const float* scale, const float* shift, int ld_ss, int do_silu,
                              const float* mr, void* y, int ld_y, int sm_count, int f16, cudaStream_t stream) {
    constexpr int U = 4;
    const int C8 = C / 8;
    if (C8 > 256) return false;
    const int rows = 256 / C8;
    const int threads = rows * C8;
    long long chunks = (16LL * sm_count + B - 1) / B;
    const long long max_chunks = (HW + rows * U - 1) / (rows * U);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int per = static_cast<int>((HW + chunks - 1) / chunks);
    per = (per + rows - 1) / rows * rows;
    chunks = (HW + per - 1) / per;
    const dim3 grid(static_cast<unsigned>(chunks), B);
    __nv_bfloat16* yo = static_cast<__nv_bfloat16*>(y);
#define NLC_GN_LEAN(X, S)                                                                                              \
    launch_pdl((gn_apply_lean_kernel<X, S, U>), dim3(grid), dim3(threads), 0, stream, x, ld_x, HW, C, groups, gamma, beta, scale, shift, ld_ss, mr, yo, \
                                                                ld_y, per, f16)
    if (x_is_op) {
        if (do_silu) NLC_GN_LEAN(true, true); else NLC_GN_LEAN(true, false);
    } else {
        if (do_silu) NLC_GN_LEAN(false, true); else NLC_GN_LEAN(false, false);
    }
#undef NLC_GN_LEAN
    return true;
}

// ---------------------------------------------------------------- small images: statistics + apply in one kernel
// The 8x8 / 4x4 / 2x2 levels (H*W < 128: no conv-epilogue partials) have at most 4096 values per (sample, group): one
// CTA keeps them in registers, takes exact two-pass moments and writes the activated operand.  One launch instead of
// three (statistics, finalize, apply), which is what these levels cost (latency, not bandwidth).  grid (groups, B)
constexpr int kGnSmallThreads = 128;
constexpr int kGnSmallVec = 8;  // float4 per thread -> up to 128 * 8 * 4 = 4096 values per group
template <bool TF32>
__global__ void __launch_bounds__(kGnSmallThreads)
    gn_small_kernel(const float* __restrict__ x, int ld_x, int HW, int C, int groups, float eps,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                    const float* __restrict__ shift, int ld_ss, int do_silu, void* __restrict__ y, int ld_y, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ float red[kGnSmallThreads / 32];
    const int g = blockIdx.x, n = blockIdx.y;
    const int cpg = C / groups, cpg4 = cpg >> 2;
    const int total4 = HW * cpg4;
    const float* xb = x + static_cast<size_t>(n) * HW * ld_x + g * cpg;
    float4 v[kGnSmallVec];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        const int i = threadIdx.x + u * kGnSmallThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total4) {
            const int pix = i / cpg4, c4 = i - pix * cpg4;
            v[u] = __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(pix) * ld_x + 4 * c4));
            s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
    }
    auto block_sum = [&](float t) {
        t = warp_sum(t);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
        __syncthreads();
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < kGnSmallThreads / 32; ++w) r += red[w];
        __syncthreads();
        return r;
    };
    const float cnt = static_cast<float>(total4) * 4.0f;
    const float mean = block_sum(s) / cnt;
    float m2 = 0.f;
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        if (threadIdx.x + u * kGnSmallThreads < total4) {
            const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
            m2 += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(block_sum(m2) / cnt + eps);
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        const int i = threadIdx.x + u * kGnSmallThreads;
        if (i < total4) {
            const int pix = i / cpg4, c4 = i - pix * cpg4;
            const int c = g * cpg + 4 * c4;
            const float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = rstd * (gamma ? gamma[c + k] : 1.f);
                float b = (beta ? beta[c + k] : 0.f) - mean * a;
                if (scale) {
                    const float sc = 1.f + scale[static_cast<size_t>(n) * ld_ss + c + k];
                    a *= sc;
                    b = b * sc + shift[static_cast<size_t>(n) * ld_ss + c + k];
                }
                f[k] = fmaf(in[k], a, b);
                if (do_silu) f[k] = silu(f[k]);
            }
            const size_t off = (static_cast<size_t>(n) * HW + pix) * ld_y + c;
            if (TF32) {
                *reinterpret_cast<float4*>(static_cast<float*>(y) + off) =
                    make_float4(op_f32(f[0], rnd), op_f32(f[1], rnd), op_f32(f[2], rnd), op_f32(f[3], rnd));
            } else {
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + off) =
                    make_uint2(pack_op16x2(f[0], f[1], rnd), pack_op16x2(f[2], f[3], rnd));
            }
        }
    }
}

// The same with one WARP per (sample, group) for <= 1024 values per group (the 4x4 / 8x8 levels of the c2 / c3 networks):
// shuffle reductions only, eight groups per CTA.  The block version above spends its time in per-CTA overhead there
// (8192 CTAs of 256 values each: 21 us per launch in profiles/r02r_launches_c2_summary.md, 4.4 % of a c2 step).
template <bool TF32>
__global__ void __launch_bounds__(256)
    gn_small_warp_kernel(const float* __restrict__ x, int ld_x, int HW, int C, int groups, int n_pairs, float eps,
                         const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                         const float* __restrict__ shift, int ld_ss, int do_silu, void* __restrict__ y, int ld_y, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int n = pair / groups, g = pair - n * groups;
    const int cpg = C / groups, cpg4 = cpg >> 2;
    const int total4 = HW * cpg4;
    const float* xb = x + static_cast<size_t>(n) * HW * ld_x + g * cpg;
    float4 v[kGnSmallVec];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        const int i = lane + u * 32;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < total4) {
            const int pix = i / cpg4, c4 = i - pix * cpg4;
            v[u] = __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(pix) * ld_x + 4 * c4));
            s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
    }
    const float cnt = static_cast<float>(total4) * 4.0f;
    const float mean = warp_sum(s) / cnt;
    float m2 = 0.f;
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        if (lane + u * 32 < total4) {
            const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
            m2 += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(m2) / cnt + eps);
#pragma unroll
    for (int u = 0; u < kGnSmallVec; ++u) {
        const int i = lane + u * 32;
        if (i < total4) {
            const int pix = i / cpg4, c4 = i - pix * cpg4;
            const int c = g * cpg + 4 * c4;
            const float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = rstd * (gamma ? gamma[c + k] : 1.f);
                float b = (beta ? beta[c + k] : 0.f) - mean * a;
                if (scale) {
                    const float sc = 1.f + scale[static_cast<size_t>(n) * ld_ss + c + k];
                    a *= sc;
                    b = b * sc + shift[static_cast<size_t>(n) * ld_ss + c + k];
                }
                f[k] = fmaf(in[k], a, b);
                if (do_silu) f[k] = silu(f[k]);
            }
            const size_t off = (static_cast<size_t>(n) * HW + pix) * ld_y + c;
            if (TF32) {
                *reinterpret_cast<float4*>(static_cast<float*>(y) + off) =
                    make_float4(op_f32(f[0], rnd), op_f32(f[1], rnd), op_f32(f[2], rnd), op_f32(f[3], rnd));
            } else {
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + off) =
                    make_uint2(pack_op16x2(f[0], f[1], rnd), pack_op16x2(f[2], f[3], rnd));
            }
        }
    }
}

static int pick_chunks(int B, int HW, int sm_count, int min_rows) {
    int chunks = 1;
    while (chunks < kGnMaxChunks && HW % (chunks * 2) == 0 && HW / (chunks * 2) >= min_rows &&
           static_cast<long long>(B) * chunks < 4LL * sm_count)
        chunks *= 2;
    return chunks;
}

template <bool TF32, int MODE, bool X16 = false>
static int launch_apply(const float* x, int ld_x, int B, int H, int W, int C, int groups, const float* gamma,
                        const float* beta, const float* scale, const float* shift, int ld_ss, int do_silu,
                        const float* mr, void* y, int ld_y, int sm_count, int rnd, cudaStream_t stream) {
    const long long npix = MODE == 2 ? static_cast<long long>(H / 2) * (W / 2) : static_cast<long long>(H) * W;
    const long long items = npix * (C / 8);
    // enough CTAs for ~16 per SM (4 resident), each with at least one full unrolled sweep
    long long chunks = (16LL * sm_count + B - 1) / B;
    const long long max_chunks = (items + kGnUnrollMax * kGnThreads - 1) / (kGnUnrollMax * kGnThreads);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const int per = static_cast<int>((items + chunks - 1) / chunks);
    chunks = (items + per - 1) / per;
    const size_t smem = static_cast<size_t>(2) * C * sizeof(float);
    launch_pdl((gn_apply_kernel<TF32, MODE, X16>), dim3(dim3(static_cast<unsigned>(chunks), B)), dim3(kGnThreads), smem, stream, 
        x, ld_x, H, W, C, groups, gamma, beta, scale, shift, ld_ss, do_silu, mr, y, ld_y, per, rnd);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

}  // namespace nlc

using namespace nlc;

extern "C" size_t nlc_groupnorm_ws(int B, int HW, int C, int groups) {
    (void)HW, (void)C;
    return static_cast<size_t>(B) * kGnMaxChunks * groups * 3 + static_cast<size_t>(B) * groups * 2;
}

extern "C" int nlc_groupnorm(nlc_ctx* ctx, const void* x_, int x_is_op, int ld_x, int B, int H, int W, int C, int groups, float eps,
                             const float* gamma, const float* beta, const float* scale, const float* shift,
                             int ld_ss, int do_silu, const float* stats, int stats_nblk, int resample, void* y_op,
                             int ld_y, int op_dtype, float* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const float* x = static_cast<const float*>(x_);
    NLC_REQUIRE(ctx && x && y_op && workspace, "nlc_groupnorm: null argument");
    NLC_REQUIRE(!x_is_op || (dtype_is16(op_dtype) && stats && ld_x % 8 == 0),
                "nlc_groupnorm: an operand-dtype input needs a 16-bit mode and conv-epilogue statistics");
    NLC_REQUIRE(groups >= 1 && groups <= 64 && C % groups == 0 && (C / groups) % 4 == 0 && C % 8 == 0,
                "nlc_groupnorm: C=%d groups=%d unsupported (channels per group must be a multiple of 4)", C, groups);
    NLC_REQUIRE(ld_x % 4 == 0 && ld_y % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_op) & 15) == 0,
                "nlc_groupnorm: tensors must be 16-byte aligned");
    NLC_REQUIRE((scale == nullptr) == (shift == nullptr), "nlc_groupnorm: scale and shift come together");
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_groupnorm: bad op_dtype");
    const int rnd = dtype_fmt(op_dtype);  // tf32 rounding (NLC_F32; NLC_F32X3 stays unrounded) / fp16 instead of bf16
    NLC_REQUIRE(static_cast<long long>(H) * W * (C / 8) < (1LL << 31), "nlc_groupnorm: image too large");
    NLC_REQUIRE(resample >= 0 && resample <= 2 && (resample != 2 || (H % 2 == 0 && W % 2 == 0)),
                "nlc_groupnorm: resample mode %d unsupported for %dx%d", resample, H, W);
    const int HW = H * W;
    float* mr = workspace + static_cast<size_t>(B) * kGnMaxChunks * groups * 3;  // [B][groups][2]

    if (!stats && resample == 0 && HW * ((C / groups) / 4) <= 32 * kGnSmallVec) {  // <= 1024 values per group: a warp each
        const int n_pairs = B * groups;
        if (!dtype_is16(op_dtype))
            launch_pdl((gn_small_warp_kernel<true>), dim3((n_pairs + 7) / 8), dim3(256), 0, stream, x, ld_x, HW, C, groups, n_pairs, eps, gamma, beta, scale,
                                                                              shift, ld_ss, do_silu, y_op, ld_y, rnd);
        else
            launch_pdl((gn_small_warp_kernel<false>), dim3((n_pairs + 7) / 8), dim3(256), 0, stream, x, ld_x, HW, C, groups, n_pairs, eps, gamma, beta,
                                                                               scale, shift, ld_ss, do_silu, y_op, ld_y, rnd);
        NLC_CHECK_LAUNCH();
        return NLC_OK;
    }
    if (!stats && resample == 0 && HW * ((C / groups) / 4) <= kGnSmallThreads * kGnSmallVec) {
        if (!dtype_is16(op_dtype))
            launch_pdl((gn_small_kernel<true>), dim3(dim3(groups, B)), dim3(kGnSmallThreads), 0, stream, 
                x, ld_x, HW, C, groups, eps, gamma, beta, scale, shift, ld_ss, do_silu, y_op, ld_y, rnd);
        else
            launch_pdl((gn_small_kernel<false>), dim3(dim3(groups, B)), dim3(kGnSmallThreads), 0, stream, 
                x, ld_x, HW, C, groups, eps, gamma, beta, scale, shift, ld_ss, do_silu, y_op, ld_y, rnd);
        NLC_CHECK_LAUNCH();
        return NLC_OK;
    }
    if (stats) {
        NLC_REQUIRE(HW % 32 == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0 && stats_nblk >= C / 4,
                    "nlc_groupnorm: conv-epilogue partials need H*W %% 32 == 0 and an 8-byte aligned buffer");
        const int bpg = (C / groups) / 4;
        if ((HW / 32) * bpg <= 512)
            launch_pdl((gn_finalize_blocks_warp_kernel), dim3((B * groups + 7) / 8), dim3(256), 0, stream, stats, stats_nblk, HW / 32, bpg,
                                                                                      groups, B * groups, eps, mr);
        else
            launch_pdl((gn_finalize_blocks_kernel), dim3(dim3(groups, B)), dim3(kGnThreads), 0, stream, stats, stats_nblk, HW / 32, bpg, eps, mr);
        NLC_CHECK_LAUNCH();
    } else {
        const int stat_chunks = pick_chunks(B, HW, ctx->sm_count, 1);
        const int C4 = C / 4;
        const int row_lanes = C4 >= kGnThreads ? 1 : kGnThreads / C4;
        const size_t smem_stats = static_cast<size_t>(row_lanes) * C4 * sizeof(Wf);
        launch_pdl((gn_stats_kernel), dim3(dim3(stat_chunks, B)), dim3(kGnThreads), smem_stats, stream, x, ld_x, HW, C, groups,
                                                                                   HW / stat_chunks, workspace);
        NLC_CHECK_LAUNCH();
        launch_pdl((gn_finalize_welford_kernel), dim3(B), dim3(64), 0, stream, workspace, stat_chunks, groups, eps, mr);
        NLC_CHECK_LAUNCH();
    }
#define NLC_GN_APPLY(T, M)                                                                                        \
    return launch_apply<T, M>(x, ld_x, B, H, W, C, groups, gamma, beta, scale, shift, ld_ss, do_silu, mr, y_op, ld_y, \
                              ctx->sm_count, rnd, stream)
    if (!dtype_is16(op_dtype)) {
        if (resample == 0) NLC_GN_APPLY(true, 0);
        if (resample == 1) NLC_GN_APPLY(true, 1);
        NLC_GN_APPLY(true, 2);
    }
    if (resample == 0 && launch_apply_lean(x, x_is_op, ld_x, B, HW, C, groups, gamma, beta, scale, shift, ld_ss, do_silu, mr,
                                           y_op, ld_y, ctx->sm_count, rnd, stream)) {
        NLC_CHECK_LAUNCH();
        return NLC_OK;
    }
#define NLC_GN_APPLY16(M)                                                                                                \
    return launch_apply<false, M, true>(x, ld_x, B, H, W, C, groups, gamma, beta, scale, shift, ld_ss, do_silu, mr, y_op, ld_y, \
                                        ctx->sm_count, rnd, stream)
    if (x_is_op) {
        if (resample == 0) NLC_GN_APPLY16(0);
        if (resample == 1) NLC_GN_APPLY16(1);
        NLC_GN_APPLY16(2);
    }
#undef NLC_GN_APPLY16
    if (resample == 0) NLC_GN_APPLY(false, 0);
    if (resample == 1) NLC_GN_APPLY(false, 1);
    NLC_GN_APPLY(false, 2);
#undef NLC_GN_APPLY
}
