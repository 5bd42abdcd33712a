// Self-attention over the H*W tokens of an NHWC feature map.
// Reference: AttnBlock (src/unet_ddim.py:186-207: bmm, *C^-1/2, softmax, bmm), QKVAttentionLegacy / QKVAttention
// (src/unet_adm.py:328-389: scale ch^-1/4 on q and k == ch^-1/2 on the logits, fp32 softmax) and AttentionOp
// (src/edm_networks.py:124-130: fp32 softmax(q k^T / sqrt(C))).
//
// T >= 128 tokens: two batched tcgen05 GEMMs (S = scale*Q K^T, O = P V through nlc_conv_tc with a batched
// right-hand operand) around an fp32 row softmax; V is transposed once so that both GEMMs see K-major
// operands.  At these sizes (T <= 1024) the T x T logits are a few % of a step's HBM traffic.
// T < 128 tokens (the 4x4 / 8x8 levels): one CTA per (image, head) keeps K and V in shared memory.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v, int rnd);
template <>
__device__ __forceinline__ float from_f32<float>(float v, int rnd) { return op_f32(v, rnd); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v, int) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v, int) { return __float2half_rn(v); }

// ---------------------------------------------------------------- row softmax: one warp per row
template <typename OutT>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, OutT* __restrict__ P,
                                                            long long rows, int T, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* s = S + row * T;
    OutT* p = P + row * T;
    float v[32];  // T <= 1024
    float m = -INFINITY;
    const int per = T >> 5;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) {
            v[i] = s[lane + 32 * i];
            m = fmaxf(m, v[i]);
        }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) {
            v[i] = __expf(v[i] - m);
            sum += v[i];
        }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (i < per) p[lane + 32 * i] = from_f32<OutT>(v[i] * inv, rnd);
}

// ---------------------------------------------------------------- V [T, dh] (pitch ld) -> V^T [dh, T]
template <typename T>
__global__ void __launch_bounds__(256) transpose_heads_kernel(const T* __restrict__ v, int ld, int head_stride,
                                                               int Tn, int heads, int dh, T* __restrict__ vt) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ T tile[32][33];
    const int bh = blockIdx.z, b = bh / heads, h = bh - b * heads;
    const T* src = v + static_cast<size_t>(b) * Tn * ld + static_cast<size_t>(h) * head_stride;
    T* dst = vt + static_cast<size_t>(bh) * dh * Tn;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = src[static_cast<size_t>(t0 + r) * ld + c0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) dst[static_cast<size_t>(c0 + r) * Tn + t0 + tx] = tile[tx][r];
}

// 16-bit elements: 64 x 64 tiles moved as 32-bit pairs, so that every global access instruction of a warp covers a full
// 128-byte line on both sides (the 32 x 32 version above moves 64-byte half lines: 1.6 TB/s in
// profiles/r02r_launches_c2_summary.md).  Needs Tn % 64 == 0 and dh % 64 == 0.
__global__ void __launch_bounds__(256) transpose_heads16_kernel(const uint16_t* __restrict__ v, int ld, int head_stride, int Tn,
                                                                 int heads, int dh, uint16_t* __restrict__ vt) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ uint16_t tile[64][64 + 2];
    const int bh = blockIdx.z, b = bh / heads, h = bh - b * heads;
    const uint16_t* src = v + static_cast<size_t>(b) * Tn * ld + static_cast<size_t>(h) * head_stride;
    uint16_t* dst = vt + static_cast<size_t>(bh) * dh * Tn;
    const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 64; r += 8) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(t0 + r) * ld + c0 + 2 * tx);
        tile[r][2 * tx] = static_cast<uint16_t>(w & 0xffffu), tile[r][2 * tx + 1] = static_cast<uint16_t>(w >> 16);
    }
    __syncthreads();
    for (int r = ty; r < 64; r += 8) {
        const uint32_t w = static_cast<uint32_t>(tile[2 * tx][r]) | (static_cast<uint32_t>(tile[2 * tx + 1][r]) << 16);
        *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(c0 + r) * Tn + t0 + 2 * tx) = w;
    }
}

// ---------------------------------------------------------------- small-T fused attention
// one CTA per (image, head); K and V rows live in shared memory as fp32 with an odd pitch.
template <typename T>
__global__ void __launch_bounds__(256)
    attn_small_kernel(const T* __restrict__ qkv, int ld, int q_off, int k_off, int v_off, int head_stride, int Tn,
                      int heads, int dh, float scale, T* __restrict__ out, int ld_out, int rnd) {
    pdl_wait();  // programmatic dependent launch (ptx.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ float sm[];
    const int pitch = dh + 1;
    float* sk = sm;
    float* sv = sm + Tn * pitch;
    float* sq = sv + Tn * pitch;  // one q row per warp: 8 * dh
    const int bh = blockIdx.x, b = bh / heads, h = bh - b * heads;
    const T* base = qkv + static_cast<size_t>(b) * Tn * ld + static_cast<size_t>(h) * head_stride;
    for (int i = threadIdx.x; i < Tn * dh; i += blockDim.x) {
        const int t = i / dh, c = i - t * dh;
        sk[t * pitch + c] = to_f32(base[static_cast<size_t>(t) * ld + k_off + c]);
        sv[t * pitch + c] = to_f32(base[static_cast<size_t>(t) * ld + v_off + c]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* q = sq + warp * dh;
    const int kper = (Tn + 31) >> 5;  // keys per lane: 1..4 (Tn <= 127)
    for (int t = warp; t < Tn; t += nwarps) {
        for (int c = lane; c < dh; c += 32) q[c] = to_f32(base[static_cast<size_t>(t) * ld + q_off + c]);
        __syncwarp();
        float s[4];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i] = -INFINITY;
            const int j = lane + 32 * i;
            if (i < kper && j < Tn) {
                float acc = 0.f;
                const float* kr = sk + j * pitch;
                for (int c = 0; c < dh; ++c) acc = fmaf(q[c], kr[c], acc);
                s[i] = acc * scale;
                m = fmaxf(m, s[i]);
            }
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s[i] = (s[i] == -INFINITY) ? 0.f : __expf(s[i] - m);
            sum += s[i];
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int c0 = 0; c0 < dh; c0 += 32) {
            const int c = c0 + lane;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < kper) {
                    for (int jj = 0; jj < 32; ++jj) {
                        const int j = 32 * i + jj;
                        const float pj = __shfl_sync(0xffffffffu, s[i], jj);
                        if (j < Tn && c < dh) acc = fmaf(pj, sv[j * pitch + c], acc);
                    }
                }
            }
            if (c < dh) out[(static_cast<size_t>(b) * Tn + t) * ld_out + h * dh + c] = from_f32<T>(acc * inv, rnd);
        }
        __syncwarp();
    }
}

}  // namespace nlc

using namespace nlc;

// attention_fused.cu
int nlc_attention_fused_16(nlc_ctx* ctx, const void* qkv, int f16, int ld, int q_off, int k_off, int head_stride, int B,
                           int T, int heads, int dh, float scale, const void* vt, void* out, int ld_out,
                           cudaStream_t stream);

// The fused kernel serves bf16 / fp16, head dimension 64 (ADM) or 256 (single-head unet_ddim / SongUNet blocks) and T a
// multiple of 64; everything
// else (fp32-container accuracy modes, the single-head dh = C blocks of unet_ddim / SongUNet) takes the GEMM + softmax
// + GEMM path below.  NLC_FUSED_ATTN=0 disables it (A/B measurements).
static bool fused_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NLC_FUSED_ATTN");
        v = !(e && e[0] == '0');
    }
    return v != 0;
}
static bool use_fused(int op_dtype, int T, int dh) {
    return dtype_is16(op_dtype) && (dh == 64 || dh == 256) && T >= 64 && T % 64 == 0 && T <= 1024 && fused_enabled();
}

extern "C" size_t nlc_attention_ws(int op_dtype, int B, int T, int heads, int dh) {
    const size_t esz = dtype_is16(op_dtype) ? 2 : 4;
    const size_t bh = static_cast<size_t>(B) * heads;
    if (use_fused(op_dtype, T, dh)) return bh * dh * T * esz + 1024;  // V^T only
    if (T < 128) return 0;
    return bh * T * T * 4 + bh * T * T * esz + bh * dh * T * esz + 1024;
}

extern "C" int nlc_attention(nlc_ctx* ctx, const void* qkv, int op_dtype, int ld, int q_off, int k_off, int v_off,
                             int head_stride, int B, int T, int heads, int dh, float scale, void* out_op, int ld_out,
                             void* workspace, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && qkv && out_op, "nlc_attention: null argument");
    NLC_REQUIRE(dtype_valid(op_dtype), "nlc_attention: bad op_dtype");
    const bool f32c = !dtype_is16(op_dtype);  // fp32 containers (tf32-rounded, or plain fp32 for NLC_F32X3)
    const bool f16 = op_dtype == NLC_F16;
    const int rnd = dtype_fmt(op_dtype);
    const size_t esz = f32c ? 4 : 2;
    if (T < 128 && !use_fused(op_dtype, T, dh)) {
        const size_t smem = (static_cast<size_t>(2) * T * (dh + 1) + 8 * dh) * sizeof(float);
        NLC_REQUIRE(smem <= 227 * 1024, "nlc_attention: T=%d dh=%d needs %zu B of shared memory", T, dh, smem);
        if (f32c) {
            NLC_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                227 * 1024));
            launch_pdl((attn_small_kernel<float>), dim3(B * heads), dim3(256), smem, stream, static_cast<const float*>(qkv), ld, q_off, k_off,
                                                                       v_off, head_stride, T, heads, dh, scale,
                                                                       static_cast<float*>(out_op), ld_out, rnd);
        } else if (f16) {
            NLC_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                227 * 1024));
            launch_pdl((attn_small_kernel<__half>), dim3(B * heads), dim3(256), smem, stream, static_cast<const __half*>(qkv), ld, q_off, k_off,
                                                                       v_off, head_stride, T, heads, dh, scale,
                                                                       static_cast<__half*>(out_op), ld_out, rnd);
        } else {
            NLC_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<__nv_bfloat16>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            launch_pdl((attn_small_kernel<__nv_bfloat16>), dim3(B * heads), dim3(256), smem, stream, 
                static_cast<const __nv_bfloat16*>(qkv), ld, q_off, k_off, v_off, head_stride, T, heads, dh, scale,
                static_cast<__nv_bfloat16*>(out_op), ld_out, rnd);
        }
        NLC_CHECK_LAUNCH();
        return NLC_OK;
    }
    NLC_REQUIRE(workspace, "nlc_attention: workspace required for T >= 128");
    NLC_REQUIRE(T % 64 == 0 && T <= 1024 && dh % 64 == 0, "nlc_attention: T=%d dh=%d unsupported", T, dh);
    NLC_REQUIRE(T >= 128 || use_fused(op_dtype, T, dh), "nlc_attention: T=%d needs the small-T kernel", T);
    const size_t bh = static_cast<size_t>(B) * heads;
    const bool fused = use_fused(op_dtype, T, dh);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* S = reinterpret_cast<float*>(ws);
    void* P = ws + bh * T * T * 4;
    void* VT = fused ? static_cast<void*>(ws) : static_cast<void*>(static_cast<uint8_t*>(P) + bh * T * T * esz);
    const uint8_t* q8 = static_cast<const uint8_t*>(qkv);

    // V^T
    {
        dim3 grid(T / 32, dh / 32, static_cast<unsigned>(bh));
        if (f32c)
            launch_pdl((transpose_heads_kernel<float>), dim3(grid), dim3(256), 0, stream, reinterpret_cast<const float*>(q8 + v_off * esz), ld,
                                                                    head_stride, T, heads, dh, static_cast<float*>(VT));
        else if (T % 64 == 0 && dh % 64 == 0 && ld % 2 == 0 && v_off % 2 == 0 && head_stride % 2 == 0)
            launch_pdl((transpose_heads16_kernel), dim3(dim3(T / 64, dh / 64, static_cast<unsigned>(bh))), dim3(256), 0, stream, 
                reinterpret_cast<const uint16_t*>(q8 + v_off * esz), ld, head_stride, T, heads, dh, static_cast<uint16_t*>(VT));
        else
            launch_pdl((transpose_heads_kernel<__nv_bfloat16>), dim3(grid), dim3(256), 0, stream, 
                reinterpret_cast<const __nv_bfloat16*>(q8 + v_off * esz), ld, head_stride, T, heads, dh,
                static_cast<__nv_bfloat16*>(VT));
        NLC_CHECK_LAUNCH();
    }
    if (fused)
        return nlc_attention_fused_16(ctx, qkv, f16 ? 1 : 0, ld, q_off, k_off, head_stride, B, T, heads, dh, scale, VT,
                                      out_op, ld_out, stream);
    // S = scale * Q K^T   ("image" = sample, "row" = head, "column" = query token)
    {
        nlc_conv_desc d;
        memset(&d, 0, sizeof(d));
        d.dtype = op_dtype;
        d.nsrc = 1;
        d.src[0] = nlc_operand{q8 + q_off * esz, B, heads, T, dh, ld, heads > 1 ? head_stride : 0,
                               static_cast<int64_t>(T) * ld};
        d.nseg = 1;
        d.seg[0] = nlc_kseg{0, 0, 0, 0, dh};
        d.wbatched = nlc_operand{q8 + k_off * esz, B, heads, T, dh, ld, heads > 1 ? head_stride : 0,
                                 static_cast<int64_t>(T) * ld};
        d.Cout = T, d.stride = 1, d.B = B, d.Ho = heads, d.Wo = T;
        d.out_scale = scale;
        d.out_f32 = S, d.ld_out_f32 = T;
        int rc = nlc_conv_tc(ctx, &d, stream_);
        if (rc != NLC_OK) return rc;
    }
    {
        const long long rows = static_cast<long long>(bh) * T;
        const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
        if (f32c)
            launch_pdl((softmax_rows_kernel<float>), dim3(grid), dim3(256), 0, stream, S, static_cast<float*>(P), rows, T, rnd);
        else if (f16)
            launch_pdl((softmax_rows_kernel<__half>), dim3(grid), dim3(256), 0, stream, S, static_cast<__half*>(P), rows, T, rnd);
        else
            launch_pdl((softmax_rows_kernel<__nv_bfloat16>), dim3(grid), dim3(256), 0, stream, S, static_cast<__nv_bfloat16*>(P), rows, T, rnd);
        NLC_CHECK_LAUNCH();
    }
    // O = P V, heads merged back into [B, T, heads*dh]
    {
        nlc_conv_desc d;
        memset(&d, 0, sizeof(d));
        d.dtype = op_dtype;
        d.nsrc = 1;
        d.src[0] = nlc_operand{P, B, heads, T, T, T, 0, 0};
        d.nseg = 1;
        d.seg[0] = nlc_kseg{0, 0, 0, 0, T};
        d.wbatched = nlc_operand{VT, B, heads, dh, T, T, 0, 0};
        d.Cout = dh, d.stride = 1, d.B = B, d.Ho = heads, d.Wo = T;
        d.out_scale = 1.0f;
        d.out_op = out_op, d.ld_out_op = ld_out;
        d.out_head_split = dh;
        int rc = nlc_conv_tc(ctx, &d, stream_);
        if (rc != NLC_OK) return rc;
    }
    return NLC_OK;
}
