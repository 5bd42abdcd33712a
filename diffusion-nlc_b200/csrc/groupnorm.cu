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

// grid (apply_chunks, B); dynamic smem: 2*C floats (per-channel a, b with y = x*a + b)
template <bool TF32>
__global__ void __launch_bounds__(kGnThreads)
    gn_apply_kernel(const float* __restrict__ x, int ld_x, int HW, int C, int groups, float eps,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ scale,
                    const float* __restrict__ shift, int ld_ss, int do_silu, void* __restrict__ y, int ld_y,
                    int stat_chunks, const float* __restrict__ ws, int rows_per_chunk) {
    extern __shared__ float gn_smem[];
    float* ca = gn_smem;
    float* cb = gn_smem + C;
    __shared__ float s_mean[64], s_rstd[64];
    const int chunk = blockIdx.x, n = blockIdx.y;
    for (int g = threadIdx.x; g < groups; g += kGnThreads) {
        Wf acc = {0.f, 0.f, 0.f};
        for (int k = 0; k < stat_chunks; ++k) {
            const float* o = ws + ((static_cast<size_t>(n) * stat_chunks + k) * groups + g) * 3;
            acc = wf_merge(acc, Wf{o[0], o[1], o[2]});
        }
        s_mean[g] = acc.mean;
        s_rstd[g] = rsqrtf(acc.m2 / acc.n + eps);
    }
    __syncthreads();
    const int cpg = C / groups;
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        const int g = c / cpg;
        float a = s_rstd[g] * (gamma ? gamma[c] : 1.f);
        float b = (beta ? beta[c] : 0.f) - s_mean[g] * a;
        if (scale) {
            const float sc = 1.f + scale[static_cast<size_t>(n) * ld_ss + c];
            a *= sc;
            b = b * sc + shift[static_cast<size_t>(n) * ld_ss + c];
        }
        ca[c] = a, cb[c] = b;
    }
    __syncthreads();
    const int C8 = C >> 3;
    const size_t row0 = static_cast<size_t>(n) * HW + static_cast<size_t>(chunk) * rows_per_chunk;
    const int total = rows_per_chunk * C8;
    for (int i = threadIdx.x; i < total; i += kGnThreads) {
        const int r = i / C8, c = (i - r * C8) << 3;
        const float* xp = x + (row0 + r) * ld_x + c;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(xp));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(xp + 4));
        float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            f[k] = fmaf(f[k], ca[c + k], cb[c + k]);
            if (do_silu) f[k] = silu(f[k]);
        }
        if (TF32) {
            float* yp = static_cast<float*>(y) + (row0 + r) * ld_y + c;
            reinterpret_cast<float4*>(yp)[0] =
                make_float4(round_tf32(f[0]), round_tf32(f[1]), round_tf32(f[2]), round_tf32(f[3]));
            reinterpret_cast<float4*>(yp)[1] =
                make_float4(round_tf32(f[4]), round_tf32(f[5]), round_tf32(f[6]), round_tf32(f[7]));
        } else {
            __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(y) + (row0 + r) * ld_y + c;
            *reinterpret_cast<uint4*>(yp) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                       pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
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

}  // namespace nlc

using namespace nlc;

extern "C" size_t nlc_groupnorm_ws(int B, int HW, int C, int groups) {
    (void)HW, (void)C;
    return static_cast<size_t>(B) * kGnMaxChunks * groups * 3;
}

extern "C" int nlc_groupnorm(nlc_ctx* ctx, const float* x, int ld_x, int B, int HW, int C, int groups, float eps,
                             const float* gamma, const float* beta, const float* scale, const float* shift,
                             int ld_ss, int do_silu, void* y_op, int ld_y, int op_dtype, float* workspace,
                             void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && x && y_op && workspace, "nlc_groupnorm: null argument");
    NLC_REQUIRE(groups >= 1 && groups <= 64 && C % groups == 0 && (C / groups) % 4 == 0 && C % 8 == 0,
                "nlc_groupnorm: C=%d groups=%d unsupported (channels per group must be a multiple of 4)", C, groups);
    NLC_REQUIRE(ld_x % 4 == 0 && ld_y % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_op) & 15) == 0,
                "nlc_groupnorm: tensors must be 16-byte aligned");
    NLC_REQUIRE((scale == nullptr) == (shift == nullptr), "nlc_groupnorm: scale and shift come together");
    NLC_REQUIRE(op_dtype == NLC_BF16 || op_dtype == NLC_F32, "nlc_groupnorm: bad op_dtype");

    const int stat_chunks = pick_chunks(B, HW, ctx->sm_count, 1);
    const int C4 = C / 4;
    const int row_lanes = C4 >= kGnThreads ? 1 : kGnThreads / C4;
    const size_t smem_stats = static_cast<size_t>(row_lanes) * C4 * sizeof(Wf);
    gn_stats_kernel<<<dim3(stat_chunks, B), kGnThreads, smem_stats, stream>>>(x, ld_x, HW, C, groups,
                                                                               HW / stat_chunks, workspace);
    NLC_CHECK_LAUNCH();

    const int apply_chunks = pick_chunks(B, HW, ctx->sm_count, 1);
    const size_t smem_apply = static_cast<size_t>(2) * C * sizeof(float);
    if (op_dtype == NLC_F32)
        gn_apply_kernel<true><<<dim3(apply_chunks, B), kGnThreads, smem_apply, stream>>>(
            x, ld_x, HW, C, groups, eps, gamma, beta, scale, shift, ld_ss, do_silu, y_op, ld_y, stat_chunks, workspace,
            HW / apply_chunks);
    else
        gn_apply_kernel<false><<<dim3(apply_chunks, B), kGnThreads, smem_apply, stream>>>(
            x, ld_x, HW, C, groups, eps, gamma, beta, scale, shift, ld_ss, do_silu, y_op, ld_y, stat_chunks, workspace,
            HW / apply_chunks);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}
