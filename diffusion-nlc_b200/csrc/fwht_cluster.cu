// 2-D fast Walsh-Hadamard transform of a whole image plane in ONE kernel (WalshHadamardCS, functions/svd_operators.py:211-251).
//
// A 256 x 256 fp32 plane is 256 KB — more than one SM's shared memory — so the two-kernel transform (operators.cu) reads and
// writes every plane twice.  Here a thread-block cluster of NC = R/64 CTAs holds the plane in distributed shared memory, 64
// rows per CTA:
//   1. rows: one warp per row, 128-bit global loads, butterflies in registers / by xor-shuffle  -> the CTA's shared tile
//   2. columns, stages h = 1 .. 32: one thread per column holds its 64 entries in registers, all butterflies in registers
//   3. columns, stages h = 64 .. R/2: the partner entry lives in CTA (rank ^ h/64) of the cluster: written back to the
//      tile, cluster barrier, read through DSMEM (ld.shared::cluster), barrier
//   4. / R, the epilogue (WH-CS gather / residual, projection, x_{t-1} assembly) and the store, straight from registers
// The stage order is ascending h throughout, as in the reference's loop, so the rounding matches the two-kernel version bit
// for bit.  HBM traffic: one read + one write of the plane (plus the epilogue's operands).
#include <cooperative_groups.h>

#include "operators.h"

namespace cg = cooperative_groups;

namespace nlc {

constexpr int FC_ROWS = 64;  // rows of the plane per CTA

template <int R>
__global__ void __launch_bounds__(R) fwht2d_cluster_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            const Epilogue e, int C) {
    constexpr int NC = R / FC_ROWS;         // CTAs per plane = cluster size
    constexpr int E = R / 32;               // entries of a row per lane
    constexpr int VW = E >= 4 ? 4 : E;      // vector width of the row phase
    constexpr int NV = E / VW;              // vectors per lane: entry index = VW*lane + j + 32*VW*i
    extern __shared__ __align__(16) float tile[];  // [FC_ROWS][R]
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = NC > 1 ? static_cast<int>(cluster.block_rank()) : 0;
    const int plane = blockIdx.x / NC;
    const int row0 = rank * FC_ROWS;
    const size_t pbase = static_cast<size_t>(plane) * R * R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- 1. rows: every warp first issues the loads of all its rows (64 entries per thread: 64 KB in flight per CTA)
    constexpr int RPW = FC_ROWS / (R / 32);  // rows per warp
    float v[RPW][E];
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        const float* src = in + pbase + static_cast<size_t>(row0 + warp + q * (R / 32)) * R;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int off = VW * lane + 32 * VW * i;
            if constexpr (VW == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src + off));
                v[q][4 * i] = t.x, v[q][4 * i + 1] = t.y, v[q][4 * i + 2] = t.z, v[q][4 * i + 3] = t.w;
            } else {
                const float2 t = __ldg(reinterpret_cast<const float2*>(src + off));
                v[q][2 * i] = t.x, v[q][2 * i + 1] = t.y;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < RPW; ++q) {
        const int r = warp + q * (R / 32);
        // register index = j + VW*i ; h < VW: bits of j
#pragma unroll
        for (int h = 1; h < VW; h <<= 1)
#pragma unroll
            for (int k = 0; k < E; ++k)
                if (!(k & h)) {
                    const float a = v[q][k], b = v[q][k + h];
                    v[q][k] = a + b, v[q][k + h] = a - b;
                }
        // h = VW .. 16 VW: lane bits
#pragma unroll
        for (int msk = 1; msk < 32; msk <<= 1) {
            const bool upper = lane & msk;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const float p = __shfl_xor_sync(0xffffffffu, v[q][k], msk);
                v[q][k] = upper ? p - v[q][k] : v[q][k] + p;
            }
        }
        // h = 32 VW ..: bits of i
#pragma unroll
        for (int h = VW; h < E; h <<= 1)
#pragma unroll
            for (int k = 0; k < E; ++k)
                if (!(k & h)) {
                    const float a = v[q][k], b = v[q][k + h];
                    v[q][k] = a + b, v[q][k + h] = a - b;
                }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float* dst = tile + r * R + VW * lane + 32 * VW * i;
            if constexpr (VW == 4)
                *reinterpret_cast<float4*>(dst) = make_float4(v[q][4 * i], v[q][4 * i + 1], v[q][4 * i + 2], v[q][4 * i + 3]);
            else
                *reinterpret_cast<float2*>(dst) = make_float2(v[q][2 * i], v[q][2 * i + 1]);
        }
    }
    __syncthreads();

    // ---- 2. columns, local stages: thread = column
    const int col = threadIdx.x;
    float w[FC_ROWS];
#pragma unroll
    for (int r = 0; r < FC_ROWS; ++r) w[r] = tile[r * R + col];
#pragma unroll
    for (int h = 1; h < FC_ROWS; h <<= 1)
#pragma unroll
        for (int r = 0; r < FC_ROWS; ++r)
            if (!(r & h)) {
                const float a = w[r], b = w[r + h];
                w[r] = a + b, w[r + h] = a - b;
            }

    // ---- 3. columns, stages across the CTAs of the cluster
    if constexpr (NC > 1) {
#pragma unroll
        for (int bit = 1; bit < NC; bit <<= 1) {
#pragma unroll
            for (int r = 0; r < FC_ROWS; ++r) tile[r * R + col] = w[r];
            cluster.sync();
            const float* peer = cluster.map_shared_rank(tile, rank ^ bit);
            const bool upper = rank & bit;
#pragma unroll
            for (int r = 0; r < FC_ROWS; ++r) {
                const float p = peer[r * R + col];
                w[r] = upper ? p - w[r] : w[r] + p;
            }
            cluster.sync();  // every CTA has read its partner before anybody overwrites its tile
        }
    }

    // ---- 4. epilogue and store
    const float fr = static_cast<float>(R);
    const int b = plane / C, c = plane % C;
#pragma unroll
    for (int r = 0; r < FC_ROWS; ++r) {
        const size_t o = pbase + static_cast<size_t>(row0 + r) * R + col;
        float v = w[r] / fr;
        if (fwht_epilogue(e, v, b, c, C, R, row0 + r, col, o)) out[o] = v;
    }
}

template <int R>
static int launch_cluster(const float* in, float* out, const Epilogue& epi, int planes, int C, cudaStream_t st) {
    constexpr int NC = R / FC_ROWS;
    const size_t smem = static_cast<size_t>(FC_ROWS) * R * sizeof(float);
    NLC_CHECK_CUDA(cudaFuncSetAttribute(fwht2d_cluster_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(planes * NC), cfg.blockDim = dim3(R), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NC, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    NLC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, fwht2d_cluster_kernel<R>, in, out, epi, C));
    return NLC_OK;
}

int fwht2d_cluster(const float* in, float* out, const Epilogue& epi, int planes, int C, int R, cudaStream_t st) {
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return 1;
    switch (R) {
        case 64: return launch_cluster<64>(in, out, epi, planes, C, st);
        case 128: return launch_cluster<128>(in, out, epi, planes, C, st);
        case 256: return launch_cluster<256>(in, out, epi, planes, C, st);
        case 512: return launch_cluster<512>(in, out, epi, planes, C, st);
        default: return 1;
    }
}

}  // namespace nlc
