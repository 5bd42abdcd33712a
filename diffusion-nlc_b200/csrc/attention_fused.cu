// Fused self-attention on the tensor cores: the ADM levels (T = 64 / 256 / 1024 tokens, 64-channel heads) and the
// single-head blocks of unet_ddim / SongUNet (T = 256, head dimension = C = 256; src/unet_ddim.py:186-207,
// src/edm_networks.py:124-130).  ADM: src/unet_adm.py:328-389 QKVAttentionLegacy / QKVAttention): S = Q K^T, softmax and
// O = P V in one kernel, so the T x T logits and probabilities never touch HBM (the unfused path in attention.cu
// writes and re-reads 6 bytes per logit: 1.6 GB per 32x32 attention block at batch 32).
//
// Persistent CTAs, two per SM, walk (image, head, 128-query block) tiles.  Per tile, with key blocks of 64 tokens:
//   pass 1:  S_kb = Q K_kb^T (tcgen05, TMEM)  ->  softmax warps keep the running row maximum (no exp)
//   pass 2:  S_kb again  ->  p = exp2((s - max) * scale*log2e), row sums in registers, P as bf16 into shared memory in
//            the K-major SWIZZLE_128B operand layout  ->  O += P V_kb (tcgen05, accumulator stays in TMEM)
//   epilogue: O / rowsum -> bf16 -> out[(n, token), head*64 + c]   (heads merged back into [B, T, C])
// Recomputing S (0.27 GFLOP per tile, tensor-core time ~2 us) is cheaper than rescaling O in TMEM and keeps the exp
// count at one per logit, which is what bounds the kernel (MUFU: 16 ex2 per clock per SM).
//
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax / epilogue (one query row per thread, TMEM lane = row).
// Two CTAs share an SM (112 KB of shared memory and 256 TMEM columns each): the softmax of one hides the TMA / MMA
// latencies of the other, and every scheduler has two softmax warps to issue from (ncu on the first, one-CTA version:
// 9 % warps active, 32 % issue-active, 195 us for the 32x32 level at batch 32).
// V is consumed as V^T [B, heads, 64, T] (written by transpose_heads_kernel) so that both GEMMs see K-major operands.
#include "common.h"
#include "ptx.cuh"

namespace nlc {

constexpr int kFaBlock = 128;                   // queries per tile
constexpr int kFaKeys = 64;                     // keys per block
constexpr int kFaTileBytes = kFaBlock * 128;    // 128 rows x 64 elements = 16 KB (one Q chunk, one P block)
constexpr int kFaPBytes = kFaTileBytes;         // P block: 128 queries x 64 keys
constexpr int kFaThreads = 64 + 128;

// DH = head dimension: 64 (ADM: two CTAs per SM, 112 KB and 256 TMEM columns each) or 256 (single-head blocks: one CTA per
// SM).  Q, K are staged as DH/64 K-major chunks of 64 channels; V^T as one [DH rows x 64 keys] tile.
template <int DH>
struct FaCfg {
    static constexpr int kChunks = DH / 64;
    static constexpr int kQBytes = kChunks * kFaTileBytes;
    static constexpr int kKChunkBytes = kFaKeys * 128;              // 64 keys x 64 channels: 8 KB
    static constexpr int kKBytes = kChunks * kKChunkBytes;
    static constexpr int kVtBytes = DH * 128;                       // DH channels x 64 keys
    static constexpr int kStageBytes = kKBytes + kVtBytes;
    static constexpr int kStages = DH == 64 ? 4 : 2;
    static constexpr int kSmem = kQBytes + kStages * kStageBytes + 2 * kFaPBytes + 256;
    static constexpr int kTmemCols = DH == 64 ? 256 : 512;          // S double buffer (2 x 64) + O (DH) -> power of two
    static constexpr int kCtasPerSm = DH == 64 ? 2 : 1;
};

struct FaParams {
    CUtensorMap mapQ, mapK, mapVt;
    int B, T, heads;
    int n_qblk, n_kblk, n_tiles;
    float scale_log2e;
    int f16;  // operands, P and the output are fp16 (else bf16)
    cudaStream_t stream;
    __nv_bfloat16* out;  // (16-bit elements of either format)
    int ld_out;
};

template <int DH>
__global__ void __launch_bounds__(kFaThreads, FaCfg<DH>::kCtasPerSm) attn_fused_kernel(const __grid_constant__ FaParams p) {
    using Cfg = FaCfg<DH>;
    constexpr int kFaStages = Cfg::kStages;
    constexpr int kFaStageBytes = Cfg::kStageBytes;
    extern __shared__ __align__(1024) uint8_t smem[];  // (the swizzled tiles need 1024-byte alignment; checked below)
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + Cfg::kQBytes;
    uint8_t* sP = sKV + kFaStages * kFaStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kFaPBytes);
    uint64_t* kv_full = bars;                    // [kFaStages]
    uint64_t* kv_empty = kv_full + kFaStages;    // [kFaStages]
    uint64_t* q_full = kv_empty + kFaStages;
    uint64_t* q_empty = q_full + 1;
    uint64_t* s_full = q_empty + 1;              // [2]
    uint64_t* s_empty = s_full + 2;              // [2]
    uint64_t* p_full = s_empty + 2;              // [2]
    uint64_t* p_empty = p_full + 2;              // [2]
    uint64_t* o_full = p_empty + 2;
    uint64_t* o_empty = o_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("nlc: attn_fused_kernel dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapQ);
        tma_prefetch_desc(&p.mapK);
        tma_prefetch_desc(&p.mapVt);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kFaStages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        mbar_init(o_empty, 4);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * kFaKeys;
    const int nkb = p.n_kblk;
    pdl_wait();  // programmatic dependent launch (ptx.cuh): the set-up above overlapped the previous kernel's tail
    pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t kv_it = 0, tile_it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
                const int qb = tile % p.n_qblk;
                const int bh = tile / p.n_qblk;
                const int h = bh % p.heads, n = bh / p.heads;
                mbar_wait(q_empty, (tile_it & 1) ^ 1);
                mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
                for (int c = 0; c < Cfg::kChunks; ++c)
                    tma_load_4d(sQ + c * kFaTileBytes, &p.mapQ, q_full, c * 64, qb * kFaBlock, h, n);
                for (int pass = 0; pass < 2; ++pass) {
                    for (int kb = 0; kb < nkb; ++kb, ++kv_it) {
                        const int st = kv_it % kFaStages;
                        mbar_wait(&kv_empty[st], ((kv_it / kFaStages) & 1) ^ 1);
                        uint8_t* sk = sKV + st * kFaStageBytes;
                        mbar_expect_tx(&kv_full[st], pass ? kFaStageBytes : Cfg::kKBytes);
#pragma unroll
                        for (int c = 0; c < Cfg::kChunks; ++c)
                            tma_load_4d(sk + c * Cfg::kKChunkBytes, &p.mapK, &kv_full[st], c * 64, kb * kFaKeys, h, n);
                        if (pass) tma_load_3d(sk + Cfg::kKBytes, &p.mapVt, &kv_full[st], kb * kFaKeys, 0, bh);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the pipeline in uniform control flow; one elected lane issues the tcgen05 instructions, whose
        // operands then stay in uniform registers (ptx.cuh: elect_one_sync).  With 64-key S blocks an MMA is 32 tensor
        // clocks: the ~17-instruction per-MMA issue sequence of a `lane == 0` loop was the bound of this kernel.
        {
            const uint32_t idesc_s = umma_idesc(p.f16 ? 0 : 1, kFaBlock, kFaKeys);  // 128 queries x 64 keys
            const uint32_t idesc_o = umma_idesc(p.f16 ? 0 : 1, kFaBlock, DH);       // 128 queries x DH channels
            const uint32_t q_addr = smem_u32(sQ);
            uint32_t kv_it = 0, s_it = 0, p_it = 0, tile_it = 0;
            // S block `si` = Q K^T of the K tile in ring slot `kv`; then the commits named by the flags (same elected lane)
            auto issue_s = [&](uint32_t kv, uint32_t si, bool free_kv, bool free_q) {
                const int st = kv % kFaStages;
                const int sb = si & 1;
                mbar_wait(&s_empty[sb], ((si >> 1) & 1) ^ 1);
                mbar_wait(&kv_full[st], (kv / kFaStages) & 1);
                tc_fence_after_sync();
                const uint32_t k_addr = smem_u32(sKV + st * kFaStageBytes);
                if (elect_one_sync()) {
#pragma unroll
                    for (int c = 0; c < Cfg::kChunks; ++c) {
                        const uint64_t qdesc = umma_desc_sw128(q_addr + c * kFaTileBytes);
                        const uint64_t kdesc = umma_desc_sw128(k_addr + c * Cfg::kKChunkBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + sb * kFaKeys, qdesc + 2 * k, kdesc + 2 * k, idesc_s, (c | k) != 0);
                    }
                    if (free_kv) umma_commit(&kv_empty[st]);
                    umma_commit(&s_full[sb]);
                    if (free_q) umma_commit(q_empty);
                }
                __syncwarp();
            };
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
                mbar_wait(q_full, tile_it & 1);
                tc_fence_after_sync();
                // pass 1: logits only (row maxima)
                for (int kb = 0; kb < nkb; ++kb, ++kv_it, ++s_it) issue_s(kv_it, s_it, true, false);
                // pass 2: S(kb+1) is issued before P(kb) is awaited, so Q K^T overlaps the softmax of the previous block;
                // after the last Q K^T of the tile has been issued Q may be overwritten once it is done (q_empty)
                issue_s(kv_it, s_it, false, nkb == 1);
                mbar_wait(o_empty, (tile_it & 1) ^ 1);  // the previous tile's O has been read out of TMEM
                for (int kb = 0; kb < nkb; ++kb, ++p_it) {
                    if (kb + 1 < nkb) issue_s(kv_it + kb + 1, s_it + kb + 1, false, kb + 2 == nkb);
                    const int pb = p_it & 1;
                    const int st = (kv_it + kb) % kFaStages;
                    mbar_wait(&p_full[pb], (p_it >> 1) & 1);
                    tc_fence_after_sync();
                    const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + pb * kFaPBytes));
                    const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * kFaStageBytes + Cfg::kKBytes));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, pdesc + 2 * k, vdesc + 2 * k, idesc_o, (kb | k) != 0);
                        umma_commit(&kv_empty[st]);
                        umma_commit(&p_empty[pb]);
                        if (kb + 1 == nkb) umma_commit(o_full);
                    }
                    __syncwarp();
                }
                kv_it += nkb;
                s_it += nkb;
            }
        }
    } else {
        // ------------------------------------------------------------ softmax + epilogue: one query row per thread
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        uint32_t s_it = 0, p_it = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
            const int qb = tile % p.n_qblk;
            const int bh = tile / p.n_qblk;
            const int h = bh % p.heads, n = bh / p.heads;
            // pass 1: running maximum of the raw logits
            float m = -INFINITY;
            for (int kb = 0; kb < nkb; ++kb, ++s_it) {
                const int sb = s_it & 1;
                mbar_wait(&s_full[sb], (s_it >> 1) & 1);
                tc_fence_after_sync();
#pragma unroll
                for (int c = 0; c < kFaKeys; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + sb * kFaKeys + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sb]);
            }
            const float mc = m * p.scale_log2e;
            // pass 2: probabilities (unnormalised) -> shared memory, row sum in a register
            float sum = 0.f;
            for (int kb = 0; kb < nkb; ++kb, ++s_it, ++p_it) {
                const int sb = s_it & 1, pb = p_it & 1;
                mbar_wait(&s_full[sb], (s_it >> 1) & 1);
                tc_fence_after_sync();
                mbar_wait(&p_empty[pb], ((p_it >> 1) & 1) ^ 1);  // O += P V of two blocks ago has finished reading it
                uint8_t* prow = sP + pb * kFaPBytes + row * 128;
#pragma unroll
                for (int c = 0; c < kFaKeys; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + sb * kFaKeys + c, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a = fast_exp2(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -mc));
                        const float b = fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -mc));
                        sum += a + b;
                        pk[i] = pack_op16x2(a, b, p.f16);
                    }
                    // keys [c, c+32) = 16-byte chunks j0..j0+3 of the row; SWIZZLE_128B: chunk ^ (row & 7)
                    uint8_t* dst = prow;
                    const int j0 = c >> 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + (((j0 + j) ^ (row & 7)) << 4)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sb]);
                fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
                mbar_arrive(&p_full[pb]);
            }
            // epilogue: O / sum -> bf16 -> out[(n, q), h*64 + c]
            mbar_wait(o_full, tile_it & 1);
            tc_fence_after_sync();
            const float inv = 1.0f / sum;
            const bool row_ok = qb * kFaBlock + row < p.T;  // rows past T were zero-filled by TMA: nothing to store
            __nv_bfloat16* orow = p.out + (static_cast<size_t>(n) * p.T + qb * kFaBlock + row) * p.ld_out + h * DH;
#pragma unroll 1
            for (int c = 0; c < DH; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_o + lane_addr + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        w[i] = pack_op16x2(__uint_as_float(v[8 * j + 2 * i]) * inv, __uint_as_float(v[8 * j + 2 * i + 1]) * inv, p.f16);
                    if (row_ok) *reinterpret_cast<uint4*>(orow + c + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// One pass over the keys (online softmax).  Same tiles, rings and warp roles; per key block:
//   S_kb = Q K_kb^T  ->  the softmax warp reads the 64 logits of its rows ONCE, m_blk = their maximum;
//   the running reference m_ref only moves when a row's block maximum exceeds it by more than 2^8 (in the exp2 domain):
//   p = exp2(s * c - m_ref) then stays <= 256, exact in fp16 / bf16 range, and the row sum is carried in fp32.  When a warp
//   does move a reference it waits for O += P V of the previous block to finish (the p_empty barrier of that block), reads
//   its 32 rows of O out of TMEM, scales them by alpha = exp2(m_ref_old - m_ref_new) and writes them back (tcgen05.st)
//   before it publishes P_kb - the next P V cannot start earlier, so the tensor core never sees a half-scaled accumulator.
//   With the first block setting the reference, rescales are rare after the first one or two blocks (the decision is per
//   warp: __any_sync; no CTA-wide vote).
// Against the two-pass kernel above: K is loaded once, S is computed once and crosses the TMEM read port once per logit
// (that port bounded the 32x32 level: profiles/r01k_ncu_adm_kernels.md), one exp per logit as before.
template <int DH>
__global__ void __launch_bounds__(kFaThreads, FaCfg<DH>::kCtasPerSm) attn_fused1_kernel(const __grid_constant__ FaParams p) {
    using Cfg = FaCfg<DH>;
    constexpr int kFaStages = Cfg::kStages;
    constexpr int kFaStageBytes = Cfg::kStageBytes;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + Cfg::kQBytes;
    uint8_t* sP = sKV + kFaStages * kFaStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kFaPBytes);
    uint64_t* kv_full = bars;
    uint64_t* kv_empty = kv_full + kFaStages;
    uint64_t* q_full = kv_empty + kFaStages;
    uint64_t* q_empty = q_full + 1;
    uint64_t* s_full = q_empty + 1;              // [2]
    uint64_t* s_empty = s_full + 2;              // [2]
    uint64_t* p_full = s_empty + 2;              // [2]
    uint64_t* p_empty = p_full + 2;              // [2]
    uint64_t* o_full = p_empty + 2;
    uint64_t* o_empty = o_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("nlc: attn_fused1_kernel dynamic shared memory is not 1024-byte aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.mapQ);
        tma_prefetch_desc(&p.mapK);
        tma_prefetch_desc(&p.mapVt);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kFaStages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1);
        mbar_init(o_empty, 4);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * kFaKeys;
    const int nkb = p.n_kblk;
    pdl_wait();  // programmatic dependent launch (ptx.cuh): the set-up above overlapped the previous kernel's tail
    pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer: Q once per tile, (K, V^T) once per key block
        if (lane == 0) {
            uint32_t kv_it = 0, tile_it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
                const int qb = tile % p.n_qblk;
                const int bh = tile / p.n_qblk;
                const int h = bh % p.heads, n = bh / p.heads;
                mbar_wait(q_empty, (tile_it & 1) ^ 1);
                mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
                for (int c = 0; c < Cfg::kChunks; ++c)
                    tma_load_4d(sQ + c * kFaTileBytes, &p.mapQ, q_full, c * 64, qb * kFaBlock, h, n);
                for (int kb = 0; kb < nkb; ++kb, ++kv_it) {
                    const int st = kv_it % kFaStages;
                    mbar_wait(&kv_empty[st], ((kv_it / kFaStages) & 1) ^ 1);
                    uint8_t* sk = sKV + st * kFaStageBytes;
                    mbar_expect_tx(&kv_full[st], kFaStageBytes);
#pragma unroll
                    for (int c = 0; c < Cfg::kChunks; ++c)
                        tma_load_4d(sk + c * Cfg::kKChunkBytes, &p.mapK, &kv_full[st], c * 64, kb * kFaKeys, h, n);
                    tma_load_3d(sk + Cfg::kKBytes, &p.mapVt, &kv_full[st], kb * kFaKeys, 0, bh);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (uniform control flow, one elected lane)
        const uint32_t idesc_s = umma_idesc(p.f16 ? 0 : 1, kFaBlock, kFaKeys);
        const uint32_t idesc_o = umma_idesc(p.f16 ? 0 : 1, kFaBlock, DH);
        const uint32_t q_addr = smem_u32(sQ);
        uint32_t kv_it = 0, tile_it = 0;  // (one S block and one P block per key block: the same counter serves all rings)
        auto issue_s = [&](uint32_t it, bool free_q) {
            const int st = it % kFaStages;
            const int sb = it & 1;
            mbar_wait(&s_empty[sb], ((it >> 1) & 1) ^ 1);
            mbar_wait(&kv_full[st], (it / kFaStages) & 1);
            tc_fence_after_sync();
            const uint32_t k_addr = smem_u32(sKV + st * kFaStageBytes);
            if (elect_one_sync()) {
#pragma unroll
                for (int c = 0; c < Cfg::kChunks; ++c) {
                    const uint64_t qdesc = umma_desc_sw128(q_addr + c * kFaTileBytes);
                    const uint64_t kdesc = umma_desc_sw128(k_addr + c * Cfg::kKChunkBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + sb * kFaKeys, qdesc + 2 * k, kdesc + 2 * k, idesc_s, (c | k) != 0);
                }
                umma_commit(&s_full[sb]);
                if (free_q) umma_commit(q_empty);
            }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
            mbar_wait(q_full, tile_it & 1);
            tc_fence_after_sync();
            issue_s(kv_it, nkb == 1);
            mbar_wait(o_empty, (tile_it & 1) ^ 1);  // the previous tile's O has been read out of TMEM
            for (int kb = 0; kb < nkb; ++kb) {
                const uint32_t it = kv_it + kb;
                if (kb + 1 < nkb) issue_s(it + 1, kb + 2 == nkb);  // Q K^T of the next block overlaps this block's softmax
                const int pb = it & 1;
                const int st = it % kFaStages;
                mbar_wait(&p_full[pb], (it >> 1) & 1);
                tc_fence_after_sync();
                const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + pb * kFaPBytes));
                const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * kFaStageBytes + Cfg::kKBytes));
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, pdesc + 2 * k, vdesc + 2 * k, idesc_o, (kb | k) != 0);
                    umma_commit(&kv_empty[st]);
                    umma_commit(&p_empty[pb]);
                    if (kb + 1 == nkb) umma_commit(o_full);
                }
                __syncwarp();
            }
            kv_it += nkb;
        }
    } else {
        // ------------------------------------------------------------ online softmax + epilogue: one query row per thread
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        constexpr float kRescaleAt = 8.0f;  // move the reference when a block maximum exceeds it by 2^8
        uint32_t it = 0, tile_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_it) {
            const int qb = tile % p.n_qblk;
            const int bh = tile / p.n_qblk;
            const int h = bh % p.heads, n = bh / p.heads;
            float mref = 0.f, sum = 0.f;  // mref: reference exponent (scaled logit units); set by the first block
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int sb = it & 1, pb = it & 1;
                mbar_wait(&s_full[sb], (it >> 1) & 1);
                tc_fence_after_sync();
                uint32_t v[kFaKeys];
                {
                    uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
                    uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + sb * kFaKeys, v0);
                    tmem_ld_32x32b_x32(tmem_base + lane_addr + sb * kFaKeys + 32, v1);
                    tmem_ld_wait();
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sb]);  // the logits are in registers: S_kb+2 may overwrite the block
                float m = __uint_as_float(v[0]);
#pragma unroll
                for (int i = 1; i < kFaKeys; ++i) m = fmaxf(m, __uint_as_float(v[i]));
                const float ms = m * p.scale_log2e;
                if (kb == 0) {
                    mref = ms;
                } else {
                    const bool need = ms > mref + kRescaleAt;
                    if (__any_sync(0xffffffffu, need)) {
                        // O += P V of the previous block has completed (its commit on that block's p_empty barrier)
                        mbar_wait(&p_empty[pb ^ 1], ((it - 1) >> 1) & 1);
                        tc_fence_after_sync();
                        const float alpha = need ? fast_exp2(mref - ms) : 1.0f;
                        if (need) mref = ms;
                        sum *= alpha;
#pragma unroll 1
                        for (int c = 0; c < DH; c += 32) {
                            uint32_t o[32];
                            tmem_ld_32x32b_x32(tmem_o + lane_addr + c, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32b_x32(tmem_o + lane_addr + c, o);
                        }
                        tmem_st_wait();
                        tc_fence_before_sync();
                    }
                }
                mbar_wait(&p_empty[pb], ((it >> 1) & 1) ^ 1);  // O += P V of two blocks ago has finished reading this P block
                uint8_t* prow = sP + pb * kFaPBytes + row * 128;
#pragma unroll
                for (int c = 0; c < kFaKeys; c += 32) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float a = fast_exp2(fmaf(__uint_as_float(v[c + 2 * i]), p.scale_log2e, -mref));
                        const float b = fast_exp2(fmaf(__uint_as_float(v[c + 2 * i + 1]), p.scale_log2e, -mref));
                        sum += a + b;
                        pk[i] = pack_op16x2(a, b, p.f16);
                    }
                    const int j0 = c >> 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(prow + (((j0 + j) ^ (row & 7)) << 4)) =
                            make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
                mbar_arrive(&p_full[pb]);
            }
            // epilogue: O / sum -> 16 bit -> out[(n, q), h*DH + c]
            mbar_wait(o_full, tile_it & 1);
            tc_fence_after_sync();
            const float inv = 1.0f / sum;
            const bool row_ok = qb * kFaBlock + row < p.T;
            __nv_bfloat16* orow = p.out + (static_cast<size_t>(n) * p.T + qb * kFaBlock + row) * p.ld_out + h * DH;
#pragma unroll 1
            for (int c = 0; c < DH; c += 32) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem_o + lane_addr + c, o);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        w[i] = pack_op16x2(__uint_as_float(o[8 * j + 2 * i]) * inv, __uint_as_float(o[8 * j + 2 * i + 1]) * inv, p.f16);
                    if (row_ok) *reinterpret_cast<uint4*>(orow + c + 8 * j) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

}  // namespace nlc

using namespace nlc;

// Internal entry (declared in attention.cu): q/k inside the qkv tensor, V^T already in `vt` as [B*heads, 64, T].
template <int DH>
static int launch_fused(nlc_ctx* ctx, const FaParams& p) {
    using Cfg = FaCfg<DH>;
    NLC_REQUIRE_DEVICE(ctx);
    static PerDeviceFlag configured;
    if (!configured[ctx->device]) {
        for (auto kern : {attn_fused_kernel<DH>, attn_fused1_kernel<DH>}) {
            NLC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
            // two CTAs per SM need the whole 228 KB as shared memory (the default carveout only guarantees one)
            NLC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                cudaSharedmemCarveoutMaxShared));
        }
        configured[ctx->device] = true;
    }
    const int slots = Cfg::kCtasPerSm * ctx->sm_count;
    const int grid = p.n_tiles < slots ? p.n_tiles : slots;
    if (ctx->attn_onepass)
        launch_pdl((attn_fused1_kernel<DH>), dim3(grid), dim3(kFaThreads), Cfg::kSmem, p.stream, p);
    else
        launch_pdl((attn_fused_kernel<DH>), dim3(grid), dim3(kFaThreads), Cfg::kSmem, p.stream, p);
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

int nlc_attention_fused_16(nlc_ctx* ctx, const void* qkv, int f16, int ld, int q_off, int k_off, int head_stride, int B,
                           int T, int heads, int dh, float scale, const void* vt, void* out, int ld_out,
                           cudaStream_t stream) {
    NLC_REQUIRE(dh == 64 || dh == 256, "nlc_attention(fused): head dimension %d unsupported (64, 256)", dh);
    NLC_REQUIRE(T % kFaKeys == 0 && T >= kFaKeys, "nlc_attention(fused): T=%d must be a multiple of %d", T, kFaKeys);
    NLC_REQUIRE(ld % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0 && head_stride % 8 == 0 && ld_out % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(vt) & 15) == 0,
                "nlc_attention(fused): tensors must be 16-byte aligned");
    FaParams p;
    memset(&p, 0, sizeof(p));
    p.B = B, p.T = T, p.heads = heads;
    p.n_qblk = (T + kFaBlock - 1) / kFaBlock, p.n_kblk = T / kFaKeys;  // (a last query tile may be half empty)
    p.n_tiles = B * heads * p.n_qblk;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.f16 = f16;
    p.stream = stream;
    p.out = static_cast<__nv_bfloat16*>(out), p.ld_out = ld_out;
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(qkv);
    for (int which = 0; which < 2; ++which) {
        // (channel within head, token, head, image)
        cuuint64_t gdim[4] = {(cuuint64_t)dh, (cuuint64_t)T, (cuuint64_t)heads, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(heads > 1 ? head_stride : dh) * 2,
                              (cuuint64_t)T * ld * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(which ? kFaKeys : kFaBlock), 1, 1};  // one 64-channel chunk per load
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ctx->encode_tiled(which ? &p.mapK : &p.mapQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                       const_cast<__nv_bfloat16*>(base + (which ? k_off : q_off)), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_attention(fused): cuTensorMapEncodeTiled(%s) failed with %d",
                    which ? "K" : "Q", (int)r);
    }
    {
        // V^T: (token, channel, image*head)
        cuuint64_t gdim[3] = {(cuuint64_t)T, (cuuint64_t)dh, (cuuint64_t)B * heads};
        cuuint64_t gstr[2] = {(cuuint64_t)T * 2, (cuuint64_t)T * dh * 2};
        cuuint32_t box[3] = {(cuuint32_t)kFaKeys, (cuuint32_t)dh, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = ctx->encode_tiled(&p.mapVt, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(vt), gdim, gstr, box,
                                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_attention(fused): cuTensorMapEncodeTiled(V^T) failed with %d", (int)r);
    }
    return dh == 64 ? launch_fused<64>(ctx, p) : launch_fused<256>(ctx, p);
}
