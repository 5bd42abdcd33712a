// Shared by the tensor-core convolution kernels (conv_tc.cu: one TMA tile per filter tap; conv_slab.cu: halo slabs shared
// by the three vertical taps): kernel parameters and the GroupNorm-partials reduction of the epilogue.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace nlc {

constexpr int kBlockM = 128;
constexpr int kChunkBytes = 128;                      // one swizzle row = one K chunk
constexpr int kAStageBytes = kBlockM * kChunkBytes;   // 16 KB
constexpr int kEpiWarps = 8;                       // two per TMEM lane quarter: they split the column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps + 32;  // warp 0 TMA (A tiles), warp 1 MMA, warps 2.. epilogue, last warp TMA (W tiles)
constexpr int kSplitWarps = 4;                     // MODE 2 only: warps 2+kEpiWarps.. split fp32 stages into hi/lo
constexpr int kThreadsX3 = 64 + 32 * kEpiWarps + 32 * kSplitWarps;  // (one producer warp; the split warps follow the epilogue)

struct ConvSegDev {
    int map, dh, dw, c0, nchunk;
};

struct ConvKParams {
    CUtensorMap mapA[NLC_MAX_SRC];
    CUtensorMap mapB;
    CUtensorMap mapOut;  // tma_epi: the 16-bit output as (channel, wo, ho, n), box = 32 channels x one warp's 32 pixels
    CUtensorMap mapRes;  // tma_epi with a residual: the 16-bit residual tensor, same box
    int B, Ho, Wo, stride;
    int BW, BH, BN;
    int tiles_w, tiles_h, tiles_n;
    int num_m_tiles, num_n_tiles, num_tiles;
    int num_m_units;  // M tiles (1-CTA kernel) or M tile pairs (CTA-pair kernel); num_tiles = num_m_units * num_n_tiles
    int Cout, nseg, total_chunks;
    ConvSegDev seg[NLC_MAX_SEG];
    const float* bias;
    const float* rowvec;
    int ld_rowvec;
    const float* resid;
    int ld_resid;
    int resid16;               // the residual is a 16-bit operand-dtype tensor (bf16, or fp16 with f16 set)
    int resid_mode;            // 0 same size, 1 nearest x2 of a half-size tensor, 2 2x2 average of a double-size tensor
    int log2_wo, log2_ho;      // (power-of-two extents: pixel index -> (n, ho, wo) by shifts)
    float out_scale;
    float* out_f32;
    int ld_out_f32;
    void* out_op;
    int ld_out_op;
    int out_head_split;
    int act;         // 1: ReLU applied last
    int out_up;      // 0, or 1 + 2a + b: output pixel (n,ho,wo) is written at (n, 2ho+a, 2wo+b) of a [B,2Ho,2Wo,.] tensor
    int w_batched;
    int f16;         // 16-bit operands are fp16 (kind::f16 with the f16 format bits), not bf16
    // split-K (small-M layers: a handful of tiles with a long serial K loop): unit u = (split = u / num_tiles, tile = u %
    // num_tiles); split s accumulates K chunks [s * chunks_per_split, ...) and writes its raw fp32 accumulators to
    // out_f32 + s * split_stride (a workspace); splitk_reduce_kernel sums the splits in a fixed order and applies the epilogue
    int ksplit, chunks_per_split, num_units;
    long long split_stride;
    int tma_epi;     // 16-bit epilogue through TMA (epi_tma_* below): output by tensor store, residual block by tensor load
    float* stats;    // GroupNorm partials of the fp32 output: [pixel/32][stats_nblk][2] = (mean, M2) over 32 px x 4 ch
    int stats_nblk;
};

// ---- 16-bit epilogue through TMA (ConvKParams.tma_epi).  A warp's 32 pixels x 32 channels are one box of the output's
// tensor map (the M tile's brick order IS the box's traversal order), 64 bytes per pixel, staged in shared memory as
// [32][64 B] with SWIZZLE_64B (16-byte chunk j of row r at chunk j ^ ((r >> 1) & 3): lanes 0..7 cover all eight 16-byte bank
// groups, so the row-per-lane 128-bit accesses are conflict-free).  Against the LSU path (fp32 staging block: 8 STS + 8 LDS
// + 4 STG per chunk, and 4 LDG + 8 STS + 8 LDS for a residual) a chunk costs 4 STS + one tensor store, and 4 LDS + one
// tensor load for the residual; rows of an image past the batch are clipped / zero-filled by the TMA unit.
constexpr int kEpiTmaBlockBytes = 32 * 64;
__device__ __forceinline__ uint32_t epi_swz64(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }

// box of one epilogue warp (32 consecutive rows of the (BN, BH, BW) brick) in (wo, ho, n)
inline void epi_tma_box(int BW, int BH, int* bw, int* bh, int* bn) {
    *bw = BW < 32 ? BW : 32;
    *bh = BH < 32 / *bw ? BH : 32 / *bw;
    *bn = 32 / (*bw * *bh);
}

// GroupNorm statistics of the tile the epilogue holds in registers, so that the consumer's GroupNorm never re-reads
// the tensor for them.  One warp = 32 consecutive pixels, f[] = 32 consecutive channels of this lane's pixel.
// For each of the 8 four-channel blocks the warp writes (mean, M2) over its 32 x 4 values: per-lane two-pass
// moments of the 4 channels, shifted by lane 0's block mean (a bf16-rounded pivot: any value near the mean removes
// the cancellation), then a transposing butterfly that reduces the 16 running sums in 16 shuffles.
__device__ __forceinline__ void gn_partials(const float (&f)[32], int lane, float* __restrict__ dst) {
    float m[8], v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m[j] = 0.25f * ((f[4 * j] + f[4 * j + 1]) + (f[4 * j + 2] + f[4 * j + 3]));
        const float a = f[4 * j] - m[j], b = f[4 * j + 1] - m[j], c = f[4 * j + 2] - m[j], d = f[4 * j + 3] - m[j];
        v[8 + j] = (a * a + b * b) + (c * c + d * d);
    }
    float pv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t pk = __shfl_sync(0xffffffffu, pack_bf16x2(m[2 * i], m[2 * i + 1]), 0);
        pv[2 * i] = __uint_as_float(pk << 16);
        pv[2 * i + 1] = __uint_as_float(pk & 0xffff0000u);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float dm = m[j] - pv[j];
        v[j] = dm;
        v[8 + j] += 4.0f * dm * dm;
    }
    // lanes end up holding: bit4 -> {sum of (m - p), sum of squares}, bits 3..1 -> block j
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
    float w8[8], w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        w8[i] = (h16 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, h16 ? v[i] : v[i + 8], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w4[i] = (h8 ? w8[i + 4] : w8[i]) + __shfl_xor_sync(0xffffffffu, h8 ? w8[i] : w8[i + 4], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i)
        w2[i] = (h4 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(0xffffffffu, h4 ? w4[i] : w4[i + 2], 4);
    float z = (h2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? w2[0] : w2[1], 2);
    z += __shfl_xor_sync(0xffffffffu, z, 1);
    const float s2 = __shfl_xor_sync(0xffffffffu, z, 16);
    if ((lane & 17) == 0) {  // bit4 == 0 (holds s1), bit0 == 0 (one of the two duplicates)
        const float p4a = h8 ? pv[4] : pv[0], p4b = h8 ? pv[5] : pv[1], p4c = h8 ? pv[6] : pv[2], p4d = h8 ? pv[7] : pv[3];
        const float p2a = h4 ? p4c : p4a, p2b = h4 ? p4d : p4b;
        const float piv = h2 ? p2b : p2a;
        const int j = (lane >> 1) & 7;
        // mean = p + s1/32;  M2 = sum (x-p)^2 - 128 (mean-p)^2 = s2 - s1^2/8
        *reinterpret_cast<float2*>(dst + 2 * j) = make_float2(piv + z * (1.0f / 32.0f), fmaxf(s2 - z * z * 0.125f, 0.0f));
    }
}


}  // namespace nlc
