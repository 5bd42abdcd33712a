// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA.  This is the kernel behind every 3x3 / 1x1 / strided convolution and every
// q,k,v / proj GEMM of the three UNet families and the sigma-model; it replaces
// the cuDNN / cuBLAS launches made by torch.nn.Conv2d in src/unet_ddim.py:99-211, src/unet_adm.py:143-305,
// src/edm_networks.py:52-205.
//
// GEMM view:  D[M = B*Ho*Wo, N = Cout] = A[M, K] * W[N, K]^T,  K = sum over "segments" (tap, channel range).
//   * activations are NHWC, so one output pixel's receptive-field slice for one tap is a contiguous
//     128-byte run of channels: exactly one row of a K-major SWIZZLE_128B UMMA operand tile;
//   * an M tile is a (BN images) x (BH rows) x (BW columns) brick with BN*BH*BW = 128 pixels; the TMA box
//     for tap (dh,dw) is that brick shifted by (dh,dw): zero padding costs nothing (OOB fill), stride-2
//     convolutions use the tensor map's traversal stride, and the nine taps re-read the brick from L2;
//   * a ResNet block's 1x1 shortcut and a concat are just more K segments over another tensor map.
//
// One persistent CTA per SM, 11 warps:  warp 0 = TMA producer of the A (activation) tiles, warp 1 = MMA issuer (one
// elected thread), warps 2-9 = epilogue (TMEM -> registers -> bias/temb/residual/scale -> fp32 and/or operand-dtype
// stores), warp 10 = TMA producer of the W (weight) tiles.
// Epilogue variants (template TEPI): 0 = staged through a per-warp shared-memory block and the load/store unit (any output
// combination, fp32 / resampled residuals, sub-pixel placement, head merge); 1 = 16-bit output only, through TMA (tensor
// store of the output block, tensor load of the 16-bit residual block; conv_common.cuh); 2 = the same with 256-bit global
// accesses from registers.  Small-M launches can split K over several CTAs per tile (ConvKParams.ksplit): the splits write
// raw accumulators to a workspace and splitk_reduce_kernel (below) applies the epilogue.
// smem ring of STAGES x (16 KB A + BLOCK_N*128 B W); two TMEM accumulators so the epilogue of tile i
// overlaps the main loop of tile i+1.
//
// Operand modes (template MODE): 0 = bf16 (kind::f16), 1 = tf32 (kind::tf32 on fp32 containers pre-rounded by the
// producers), 2 = fp32 (NLC_F32X3): operands arrive as plain fp32; four extra "split" warps rewrite every stage in
// shared memory into a tf32 high part (in place) and a tf32 low part (second half of the stage), and the issuer runs
// three MMAs per K step, A_hi.W_hi + A_lo.W_hi + A_hi.W_lo, which recovers fp32-accurate products (the dropped
// A_lo.W_lo term is ~2^-22 relative).  Because the tensor core's own accumulation truncates, every K chunk gets a
// fresh TMEM accumulator and the epilogue warps sum the chunks in registers.  This is the accuracy mode behind the
// <=1e-4 per-step parity tests.
#include "conv_common.cuh"

namespace nlc {

// CTA2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BLOCK_N tile; each CTA stages its own 128
// rows of A and one half of the B tile, so the weights cross L2->SM once per pair.
template <int BLOCK_N, bool CTA2, bool X3 = false>
struct ConvCfg {
    static constexpr int kBRows = CTA2 ? BLOCK_N / 2 : BLOCK_N;  // B rows staged by this CTA
    static constexpr int kBStageBytes = kBRows * kChunkBytes;
    static constexpr int kLoadBytes = kAStageBytes + kBStageBytes;     // what TMA delivers per stage
    static constexpr int kStageBytes = X3 ? 2 * kLoadBytes : kLoadBytes;  // X3: [A_hi | W_hi | A_lo | W_lo]
    static constexpr int kStages =
        X3 ? (BLOCK_N == 128 ? 3 : 4)
           : (CTA2 ? (BLOCK_N == 256 ? 6 : 8) : ((BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8)));
    static constexpr int kTmemCols = 2 * BLOCK_N;  // 128 / 256 / 512: power of two >= 32
    // mbarriers (ring, accumulators), TMEM slot, one residual-block barrier per epilogue warp (tma_epi); rounded to 512 so
    // that the epilogue staging blocks behind it are aligned for SWIZZLE_64B tensor loads / stores
    static constexpr int kBarBytes = ((3 * kStages + 4) * 8 + 16 + 8 * kEpiWarps + 511) / 512 * 512;
    static constexpr int kEpiBytes = kEpiWarps * 32 * 32 * 4;  // one 32x32 fp32 staging block per epilogue warp
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kEpiBytes + 1024;  // +1024: alignment slack
};

// TEPI: the 16-bit-only epilogues (conv_common.cuh; ConvKParams.tma_epi launches): 1 = through TMA, 2 = 256-bit global
// accesses straight from / to registers - separate instantiations, so that no epilogue carries another's registers
template <int BLOCK_N, int MODE, bool CTA2, int TEPI = 0>
__global__ void __launch_bounds__(MODE == 2 ? kThreadsX3 : kThreads, 1)
    conv_tc_kernel(const __grid_constant__ ConvKParams p) {
    static_assert(TEPI == 0 || MODE == 0, "the 16-bit-only epilogues serve the 16-bit operand modes");
    constexpr bool TF32 = MODE != 0;  // fp32 containers, kind::tf32
    constexpr bool X3 = MODE == 2;    // unrounded fp32 operands, split in shared memory, three MMAs per K step
    static_assert(!(X3 && CTA2) && !(X3 && BLOCK_N == 256), "the fp32 mode runs on the 1-CTA kernel with N <= 128");
    using Cfg = ConvCfg<BLOCK_N, CTA2, X3>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kChunkElems = TF32 ? 32 : 64;
    // work units: tiles for the 1-CTA kernel, 256-row tile pairs for the CTA-pair kernel (both CTAs of a pair walk
    // the same sequence; CTA `rank` owns M tile 2*unit + rank)
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    const int unit0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* split = bars + 2 * kStages;  // X3: stage has been split into hi/lo (one arrive per split thread)
    uint64_t* tfull = bars + 3 * kStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* rbar = tempty + 4;  // [kEpiWarps] (tma_epi: a warp's residual block has landed)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NLC_MAX_SRC; ++i) tma_prefetch_desc(&p.mapA[i]);
        tma_prefetch_desc(&p.mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&split[s], 32 * kSplitWarps);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], CTA2 ? 2 * kEpiWarps : kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
        }
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(&rbar[w], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if (CTA2) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot); else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CTA2) cluster_sync_all();  // the peer's barriers are initialised before anything remote touches them
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // (programmatic dependent launch: everything above overlapped the previous kernel's tail; nothing below may)
    pdl_wait();
    pdl_trigger();

    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int uu = unit0; uu < p.num_units; uu += unit_step) {
                const int split = uu / p.num_tiles, unit = uu - split * p.num_tiles;  // (ksplit == 1: split 0, unit = uu)
                const int k_lo = split * p.chunks_per_split;
                const int k_hi = k_lo + p.chunks_per_split < p.total_chunks ? k_lo + p.chunks_per_split : p.total_chunks;
                const int n_tile = unit / p.num_m_units;
                const int m_tile = CTA2 ? 2 * (unit - n_tile * p.num_m_units) + static_cast<int>(rank)
                                        : unit - n_tile * p.num_m_units;
                const int tn = m_tile / tiles_per_img;
                const int rem = m_tile - tn * tiles_per_img;
                const int th = rem / p.tiles_w;
                const int tw = rem - th * p.tiles_w;
                const int n0 = tn * p.BN, h0 = th * p.BH * p.stride, w0 = tw * p.BW * p.stride;
                int kchunk = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSegDev sg = p.seg[s];
                    for (int j = 0; j < sg.nchunk; ++j) {
                        if (kchunk < k_lo || kchunk >= k_hi) {  // another split's K range
                            ++kchunk;
                            continue;
                        }
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sa = smem + stage * Cfg::kStageBytes;
                        uint8_t* sb = sa + kAStageBytes;
                        if (CTA2) {
                            // both CTAs' bytes land on the leader's barrier (a tile past the end is all zero fill)
                            if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kLoadBytes);
                            tma_load_4d_pair(sa, &p.mapA[sg.map], &full[stage], sg.c0 + j * kChunkElems, w0 + sg.dw,
                                             h0 + sg.dh, n0);
                            if (X3)
                                tma_load_4d_pair(sb, &p.mapB, &full[stage], kchunk * kChunkElems,
                                                 n_tile * BLOCK_N + static_cast<int>(rank) * Cfg::kBRows, 0, 0);
                        } else {
                            mbar_expect_tx(&full[stage], Cfg::kLoadBytes);
                            tma_load_4d(sa, &p.mapA[sg.map], &full[stage], sg.c0 + j * kChunkElems, w0 + sg.dw,
                                        h0 + sg.dh, n0);
                            if (X3)  // (the other modes have a second producer warp for the W tiles, below)
                                tma_load_4d(sb, &p.mapB, &full[stage], kchunk * kChunkElems, n_tile * BLOCK_N,
                                            p.w_batched ? th : 0, p.w_batched ? n0 : 0);
                        }
                        ++kchunk;
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the pipeline (waits, stage / accumulator bookkeeping) in uniform control flow and one
        // elected lane issues the tcgen05 instructions (see elect_one_sync); the pair's leader CTA issues for both CTAs.
        if (rank == 0) {
            // operand format bits: tf32, or for the 16-bit modes bf16 / fp16 (same kind::f16 instruction)
            const uint32_t idesc = umma_idesc(TF32 ? 2 : (p.f16 ? 0 : 1), CTA2 ? 2 * kBlockM : kBlockM, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            if (X3) {
                // One TMEM accumulator per K chunk: the tensor core accumulates with truncation (round toward zero), so a
                // long in-TMEM sum drifts by ~n * 2^-25 of its magnitude (1e-5 at K = 2304).  Each chunk's 12 MMAs
                // (low-order terms first, while the accumulator is still tiny) therefore go to a fresh accumulator and
                // the epilogue warps add the chunks up in registers with round-to-nearest.
                for (int unit = unit0; unit < p.num_tiles; unit += unit_step) {
                    for (int kc = 0; kc < p.total_chunks; ++kc) {
                        mbar_wait(&tempty[acc], acc_phase ^ 1);
                        mbar_wait(&split[stage], phase);
                        tc_fence_after_sync();
                        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                        const uint64_t ahi = umma_desc_sw128(sa), bhi = umma_desc_sw128(sa + kAStageBytes);
                        const uint64_t alo = umma_desc_sw128(sa + Cfg::kLoadBytes);
                        const uint64_t blo = umma_desc_sw128(sa + Cfg::kLoadBytes + kAStageBytes);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, alo + 2 * k, bhi + 2 * k, idesc, k != 0);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, ahi + 2 * k, blo + 2 * k, idesc, 1);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, ahi + 2 * k, bhi + 2 * k, idesc, 1);
                            umma_commit(&empty[stage]);
                            umma_commit(&tfull[acc]);
                        }
                        __syncwarp();
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                        acc ^= 1;
                        if (acc == 0) acc_phase ^= 1;
                    }
                }
            } else {
                for (int uu = unit0; uu < p.num_units; uu += unit_step) {
                    const int k_lo = (uu / p.num_tiles) * p.chunks_per_split;
                    const int k_hi = k_lo + p.chunks_per_split < p.total_chunks ? k_lo + p.chunks_per_split : p.total_chunks;
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                    for (int kc = k_lo; kc < k_hi; ++kc) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after_sync();
                        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                        const uint64_t adesc = umma_desc_sw128(sa);
                        const uint64_t bdesc = umma_desc_sw128(sa + kAStageBytes);
                        if (elect_one_sync()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {  // 4 x 32 B K-steps inside the 128 B swizzle row
                                if (CTA2) {
                                    if (TF32)
                                        umma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kc - k_lo) | k) != 0);
                                    else
                                        umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kc - k_lo) | k) != 0);
                                } else if (TF32) {
                                    umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kc - k_lo) | k) != 0);
                                } else {
                                    umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kc - k_lo) | k) != 0);
                                }
                            }
                            // smem slot reusable (in both CTAs of a pair) once these MMAs retire
                            if (CTA2) umma_commit_pair(&empty[stage]); else umma_commit(&empty[stage]);
                        }
                        __syncwarp();
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (elect_one_sync()) {  // accumulator -> epilogue(s)
                        if (CTA2) umma_commit_pair(&tfull[acc]); else umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            }
        }
    } else if (!X3 && warp == 2 + kEpiWarps) {
        // ------------------------------------------------------------ second TMA producer: the W (right-hand) tiles
        // One thread has ~250 clocks per K chunk at BLOCK_N = 128; waiting for the slot, arming the barrier and issuing
        // two tensor loads does not fit, so the A and W loads of a stage are issued by two warps.  The A producer arms
        // the stage's barrier with the byte count of both (a transaction that completes before the expect is legal:
        // the phase cannot complete until the arrive that comes with it).
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int uu = unit0; uu < p.num_units; uu += unit_step) {
                const int split = uu / p.num_tiles, unit = uu - split * p.num_tiles;
                const int k_lo = split * p.chunks_per_split;
                const int k_hi = k_lo + p.chunks_per_split < p.total_chunks ? k_lo + p.chunks_per_split : p.total_chunks;
                const int n_tile = unit / p.num_m_units;
                const int m_tile = CTA2 ? 2 * (unit - n_tile * p.num_m_units) + static_cast<int>(rank)
                                        : unit - n_tile * p.num_m_units;
                const int tn = m_tile / tiles_per_img;
                const int th = (m_tile - tn * tiles_per_img) / p.tiles_w;
                const int n0 = tn * p.BN;
                for (int kchunk = k_lo; kchunk < k_hi; ++kchunk) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sb = smem + stage * Cfg::kStageBytes + kAStageBytes;
                    if (CTA2)
                        tma_load_4d_pair(sb, &p.mapB, &full[stage], kchunk * kChunkElems,
                                         n_tile * BLOCK_N + static_cast<int>(rank) * Cfg::kBRows, 0, 0);
                    else
                        tma_load_4d(sb, &p.mapB, &full[stage], kchunk * kChunkElems, n_tile * BLOCK_N,
                                    p.w_batched ? th : 0, p.w_batched ? n0 : 0);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (X3 && warp >= 2 + kEpiWarps) {
        // ------------------------------------------------------------ fp32 split (MODE 2 only)
        // hi = tf32(v) in place, lo = tf32(v - hi) at the same (swizzled) offset of the stage's second half; the
        // subtraction is exact in fp32, so hi + lo carries 21+ mantissa bits of v into the tensor core.
        const int tid = static_cast<int>(threadIdx.x) - 32 * (2 + kEpiWarps);
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = unit0; unit < p.num_tiles; unit += unit_step) {
            for (int kc = 0; kc < p.total_chunks; ++kc) {
                mbar_wait(&full[stage], phase);
                float4* hi = reinterpret_cast<float4*>(smem + stage * Cfg::kStageBytes);
                float4* lo = reinterpret_cast<float4*>(smem + stage * Cfg::kStageBytes + Cfg::kLoadBytes);
#pragma unroll 4
                for (int i = tid; i < Cfg::kLoadBytes / 16; i += 32 * kSplitWarps) {
                    const float4 v = hi[i];
                    const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
                    hi[i] = h;
                    lo[i] = make_float4(round_tf32(v.x - h.x), round_tf32(v.y - h.y), round_tf32(v.z - h.z),
                                        round_tf32(v.w - h.w));
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                mbar_arrive(&split[stage]);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        // Each warp owns TMEM lanes [32*quad, 32*quad+32) = 32 consecutive output pixels; a lane holds one pixel's
        // 32 channels of the current column chunk.  Global traffic goes through a per-warp 32x32 fp32 staging block
        // (16-byte chunks XOR-swizzled by the row, conflict-free both ways) so that every load / store instruction
        // touches 4 full 128-byte rows instead of 32 partial ones.
        const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
        const int half = (warp - 2) >> 2;  // the two warps of a lane quarter take alternate 32-column chunks
        float* stg = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + Cfg::kBarBytes) + (warp - 2) * 1024;
        // tma_epi: the same 4 KB as a 16-bit output block and a 16-bit residual block ([32][64 B], SWIZZLE_64B)
        uint8_t* ostg = reinterpret_cast<uint8_t*>(stg);
        uint8_t* rstg = ostg + kEpiTmaBlockBytes;
        uint64_t* my_rbar = &rbar[warp - 2];
        uint32_t rphase = 0;
        const int brick = p.BW * p.BH;
        const int row = quad * 32 + lane;
        const int bn = row / brick;
        // the warp's first row inside the tile (its 32 rows are consecutive pixels of one image: DESIGN.md §3)
        const int row0 = quad * 32;
        const int bn0 = row0 / brick;
        const int r20 = row0 - bn0 * brick;
        const int bh0 = r20 / p.BW;
        const int bw0 = r20 - bh0 * p.BW;
        const int sub_r4 = lane >> 3, sub_c4 = lane & 7;  // fp32 pattern: 4 rows x 8 chunks per instruction
        const int sub_r8 = lane >> 2, sub_c8 = lane & 3;  // bf16 pattern: 8 rows x 4 (8-channel) chunks
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int uu = unit0; uu < p.num_units; uu += unit_step) {
            const int split = uu / p.num_tiles, unit = uu - split * p.num_tiles;
            float* const of32 = p.out_f32 + static_cast<size_t>(split) * p.split_stride;  // (split-K: this split's workspace slice)
            const int n_tile = unit / p.num_m_units;
            const int m_tile = CTA2 ? 2 * (unit - n_tile * p.num_m_units) + static_cast<int>(rank)
                                    : unit - n_tile * p.num_m_units;
            const int tn = m_tile / tiles_per_img;
            const int rem = m_tile - tn * tiles_per_img;
            const int th = rem / p.tiles_w;
            const int tw = rem - th * p.tiles_w;
            const int n = tn * p.BN + bn;  // this lane's image (bias row of the per-sample vector)
            const bool valid = n < p.B;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            const int n0 = tn * p.BN + bn0, ho0 = th * p.BH + bh0, wo0 = tw * p.BW + bw0;
            // dense NHWC row, or (attention head merge) row (n,wo) with a channel offset of ho*split
            const size_t pix0 = p.out_head_split ? static_cast<size_t>(n0) * p.Wo + wo0
                                                 : (static_cast<size_t>(n0) * p.Ho + ho0) * p.Wo + wo0;
            const int hs_off = p.out_head_split * ho0;
            // sub-pixel placement (one phase of a nearest-x2 upsample + 3x3 conv computed at the LOW resolution)
            const int up_a = (p.out_up - 1) >> 1, up_b = (p.out_up - 1) & 1;
            auto opix = [&](int r) -> size_t {
                const size_t q = pix0 + r;
                if (!p.out_up) return q;
                const size_t wo = q & (static_cast<size_t>(p.Wo) - 1);
                const size_t ho = (q >> p.log2_wo) & (static_cast<size_t>(p.Ho) - 1);
                const size_t nn = q >> (p.log2_wo + p.log2_ho);
                return ((nn * 2 * p.Ho + 2 * ho + up_a) * 2 * p.Wo) + 2 * wo + up_b;
            };
            // GroupNorm partial block of this warp's 32 pixels: in output order, or (placement) sample-major, phase, block
            size_t stat_blk = (pix0 + lane) >> 5;
            if (p.out_up) {
                const size_t hw = static_cast<size_t>(p.Ho) * p.Wo;
                const size_t nn = pix0 >> (p.log2_wo + p.log2_ho);
                stat_blk = nn * (4 * hw >> 5) + static_cast<size_t>(p.out_up - 1) * (hw >> 5) + ((pix0 & (hw - 1)) >> 5);
            }

            // Residual rows of the chunk about to be processed are fetched one chunk ahead (mode 0): the first chunk's
            // loads are issued before the accumulator wait, every later chunk's while the previous one is being stored.
            // (a 16-bit residual - nlc_conv_desc.resid_is_op - comes as four 16-byte loads of 8 channels per lane, 8 rows x 4
            // chunks per instruction, instead of eight fp32 ones)
            float4 rpre[8];
            constexpr bool tma_epi = TEPI == 1;  // (host: 16-bit output only, dense placement, 16-bit residual or none)
            const bool pre = p.resid != nullptr && p.resid_mode == 0 && TEPI == 0;
            const __nv_bfloat16* resid16 = reinterpret_cast<const __nv_bfloat16*>(p.resid);
            auto prefetch_resid = [&](int c_next) {
                if (p.resid16) {
                    const __nv_bfloat16* rp = resid16 + pix0 * p.ld_resid + n_tile * BLOCK_N + c_next;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + sub_r8;
                        uint4 t = make_uint4(0u, 0u, 0u, 0u);
                        if ((vmask >> r) & 1) t = __ldg(reinterpret_cast<const uint4*>(rp + static_cast<size_t>(r) * p.ld_resid) + sub_c8);
                        rpre[it] = make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
                    }
                    return;
                }
                const float* rp = p.resid + pix0 * p.ld_resid + n_tile * BLOCK_N + c_next;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + sub_r4;
                    rpre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((vmask >> r) & 1) rpre[it] = __ldg(reinterpret_cast<const float4*>(rp + static_cast<size_t>(r) * p.ld_resid) + sub_c4);
                }
            };
            if (pre) prefetch_resid(32 * half);
            if (tma_epi && p.resid && lane == 0) {  // first residual block of the tile: a tensor load, before the accumulator wait
                mbar_expect_tx(my_rbar, kEpiTmaBlockBytes);
                tma_load_4d(rstg, &p.mapRes, my_rbar, n_tile * BLOCK_N + 32 * half, wo0, ho0, n0);
            }
            // TEPI 2: the lane's own 64 residual bytes of the next chunk, two 256-bit loads
            uint32_t rq[2][8];
            const __nv_bfloat16* rrow = resid16 + (pix0 + lane) * p.ld_resid + n_tile * BLOCK_N;
            auto fetch_resid256 = [&](int c_next) {
                if (valid) {
                    ldg256(rrow + c_next, rq[0]);
                    ldg256(rrow + c_next + 16, rq[1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rq[0][i] = rq[1][i] = 0u;
                }
            };
            if (TEPI == 2 && p.resid) fetch_resid256(32 * half);
            // ... and the residual block of the NEXT tile of this CTA is pulled into L2 now, a whole tile ahead (conv_slab.cu)
            if (p.resid != nullptr && p.resid_mode == 0 && !p.out_head_split && unit + unit_step < p.num_tiles) {  // (never with split-K)
                const int u2 = unit + unit_step;
                const int nt2 = u2 / p.num_m_units;
                const int mt2 = CTA2 ? 2 * (u2 - nt2 * p.num_m_units) + static_cast<int>(rank) : u2 - nt2 * p.num_m_units;
                const int tn2 = mt2 / tiles_per_img;
                const int rem2 = mt2 - tn2 * tiles_per_img;
                const int th2 = rem2 / p.tiles_w, tw2 = rem2 - th2 * p.tiles_w;
                const int n2 = tn2 * p.BN + bn0;
                if (n2 < p.B) {
                    const size_t px = (static_cast<size_t>(n2) * p.Ho + th2 * p.BH + bh0) * p.Wo + tw2 * p.BW + bw0;
                    const int esz = p.resid16 ? 2 : 4;
                    const int lines = (BLOCK_N * esz) >> 7;
                    const char* base = reinterpret_cast<const char*>(p.resid) + static_cast<size_t>(nt2) * BLOCK_N * esz;
                    for (int i = lane + 32 * half; i < 32 * lines; i += 32 * (kEpiWarps / 4))
                        prefetch_l2(base + (px + i / lines) * static_cast<size_t>(p.ld_resid) * esz + ((i % lines) << 7));
                }
            }
            // X3: the K chunks arrive one accumulator at a time and are summed here, in registers (round to nearest)
            constexpr int kNJ = X3 ? BLOCK_N / (32 * (kEpiWarps / 4)) : 1;  // column chunks owned by this warp
            float accr[kNJ][32];
            if (X3) {
#pragma unroll
                for (int j = 0; j < kNJ; ++j)
#pragma unroll
                    for (int i = 0; i < 32; ++i) accr[j][i] = 0.f;
                for (int kc = 0; kc < p.total_chunks; ++kc) {
                    mbar_wait(&tfull[acc], acc_phase);
                    tc_fence_after_sync();
                    const uint32_t ta = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + 32 * half;
#pragma unroll
                    for (int j = 0; j < kNJ; ++j) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(ta + 32 * (kEpiWarps / 4) * j, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) accr[j][i] += __uint_as_float(v[i]);
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&tempty[acc]);
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1;
                }
            } else {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after_sync();
            }
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
            if constexpr (TEPI != 0) {
#pragma unroll 1
                for (int c = 32 * half; c < BLOCK_N; c += 32 * (kEpiWarps / 4)) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c, v);
                    const int col0 = n_tile * BLOCK_N + c;
                    uint4 rr[4];
                    if (p.resid) {
                        if constexpr (TEPI == 1) {
                            // this chunk's residual block has landed: own row -> registers, then the next block's load
                            mbar_wait(my_rbar, rphase);
                            rphase ^= 1;
#pragma unroll
                            for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(rstg + epi_swz64(lane, j));
                            // (the block is handed back to the TMA unit further down, once these loads have RETURNED)
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                rr[j] = make_uint4(rq[j >> 1][4 * (j & 1)], rq[j >> 1][4 * (j & 1) + 1], rq[j >> 1][4 * (j & 1) + 2],
                                                   rq[j >> 1][4 * (j & 1) + 3]);
                            if (c + 32 * (kEpiWarps / 4) < BLOCK_N) fetch_resid256(c + 32 * (kEpiWarps / 4));
                        }
                    }
                    tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                    if (p.bias) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 t = __ldg(b4 + i);
                            f[4 * i] += t.x, f[4 * i + 1] += t.y, f[4 * i + 2] += t.z, f[4 * i + 3] += t.w;
                        }
                    }
                    if (p.rowvec && valid) {
                        const float4* b4 =
                            reinterpret_cast<const float4*>(p.rowvec + static_cast<size_t>(n) * p.ld_rowvec + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 t = __ldg(b4 + i);
                            f[4 * i] += t.x, f[4 * i + 1] += t.y, f[4 * i + 2] += t.z, f[4 * i + 3] += t.w;
                        }
                    }
                    if (p.resid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float4 a, b;
                            unpack_op16x8(rr[j], p.f16, a, b);
                            f[8 * j] += a.x, f[8 * j + 1] += a.y, f[8 * j + 2] += a.z, f[8 * j + 3] += a.w;
                            f[8 * j + 4] += b.x, f[8 * j + 5] += b.y, f[8 * j + 6] += b.z, f[8 * j + 7] += b.w;
                        }
                        if constexpr (TEPI == 1) {
                            // Every lane has CONSUMED its residual values, so its shared-memory loads have returned: only now
                            // may the TMA unit overwrite the block with the next one.  (Issuing the next load right after the
                            // loads were issued left a window - a warp barrier does not wait for outstanding LDS, and under
                            // the MMA's operand traffic an LDS can outlast the tensor load of an L2-resident block: one image
                            // in ~50 forward passes of ADM-256 came out different, scripts/repro_check.py.)
                            fence_proxy_async_smem();  // generic-proxy reads of the block -> before the async-proxy write of the next
                            __syncwarp();
                            if (c + 32 * (kEpiWarps / 4) < BLOCK_N && lane == 0) {
                                mbar_expect_tx(my_rbar, kEpiTmaBlockBytes);
                                tma_load_4d(rstg, &p.mapRes, my_rbar, col0 + 32 * (kEpiWarps / 4), wo0, ho0, n0);
                            }
                        }
                    }
                    if (p.out_scale != 1.0f) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] *= p.out_scale;
                    }
                    if (p.act) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                    }
                    if (p.stats && vmask == 0xffffffffu)
                        gn_partials(f, lane, p.stats + (stat_blk * p.stats_nblk + (col0 >> 2)) * 2);
                    if constexpr (TEPI == 1) {
                        // the previous chunk's tensor store has finished READING the staging block before it is rewritten
                        if (lane == 0) bulk_wait_group_read0();
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(ostg + epi_swz64(lane, j)) =
                                make_uint4(pack_op16x2(f[8 * j], f[8 * j + 1], p.f16), pack_op16x2(f[8 * j + 2], f[8 * j + 3], p.f16),
                                           pack_op16x2(f[8 * j + 4], f[8 * j + 5], p.f16), pack_op16x2(f[8 * j + 6], f[8 * j + 7], p.f16));
                        fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA unit's async-proxy read
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_4d(&p.mapOut, ostg, col0, wo0, ho0, n0);
                            bulk_commit_group();
                        }
                    } else if (valid) {
                        __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(p.out_op) + (pix0 + lane) * p.ld_out_op + col0;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            uint32_t w8[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w8[i] = pack_op16x2(f[16 * h2 + 2 * i], f[16 * h2 + 2 * i + 1], p.f16);
                            stg256(orow + 16 * h2, w8);
                        }
                    }
                }
            } else {
#pragma unroll 1
            for (int c = 32 * half; c < BLOCK_N; c += 32 * (kEpiWarps / 4)) {
                uint32_t v[32];
                if (!X3) tmem_ld_32x32b_x32(taddr + c, v);
                const int col0 = n_tile * BLOCK_N + c;
                const int ocol0 = col0 + hs_off;
                if (p.resid) {  // coalesced gather of the residual block while the TMEM load is in flight
                    // (one loop per mode, so that the eight loads of a lane stay back to back: with the mode test inside
                    // the loop the compiler serialised them and the residual layers lost 25 %)
                    if (p.resid_mode == 0) {
                        if (p.resid16) {
#pragma unroll
                            for (int it = 0; it < 4; ++it) {
                                const int r = it * 8 + sub_r8;
                                float4 a, b;
                                unpack_op16x8(make_uint4(__float_as_uint(rpre[it].x), __float_as_uint(rpre[it].y),
                                                         __float_as_uint(rpre[it].z), __float_as_uint(rpre[it].w)),
                                              p.f16, a, b);
                                *reinterpret_cast<float4*>(stg + r * 32 + (((2 * sub_c8) ^ (r & 7)) << 2)) = a;
                                *reinterpret_cast<float4*>(stg + r * 32 + (((2 * sub_c8 + 1) ^ (r & 7)) << 2)) = b;
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int r = it * 4 + sub_r4;
                                *reinterpret_cast<float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2)) = rpre[it];
                            }
                        }
                        if (c + 32 * (kEpiWarps / 4) < BLOCK_N) prefetch_resid(c + 32 * (kEpiWarps / 4));
                    } else if (p.resid16) {
                        // resampled skip path on a 16-bit residual: 8-byte loads of 4 channels (mode 1: the half-resolution
                        // pixel; mode 2: the 2x2 window of the double-resolution tensor, summed in fp32, then * 0.25)
#pragma unroll 2
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + sub_r4;
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((vmask >> r) & 1) {
                                const size_t q = pix0 + r;
                                const size_t wo = q & (static_cast<size_t>(p.Wo) - 1);
                                const size_t ho = (q >> p.log2_wo) & (static_cast<size_t>(p.Ho) - 1);
                                const size_t nn = q >> (p.log2_wo + p.log2_ho);
                                if (p.resid_mode == 1) {
                                    const size_t src = (nn * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1);
                                    t = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(resid16 + src * p.ld_resid + col0) + sub_c4), p.f16);
                                } else {
                                    const size_t w2 = 2 * static_cast<size_t>(p.Wo);
                                    const __nv_bfloat16* s0 = resid16 + ((nn * 2 * p.Ho + 2 * ho) * w2 + 2 * wo) * p.ld_resid + col0;
                                    const float4 a = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(s0) + sub_c4), p.f16);
                                    const float4 b = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(s0 + p.ld_resid) + sub_c4), p.f16);
                                    const float4 d = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(s0 + w2 * p.ld_resid) + sub_c4), p.f16);
                                    const float4 e = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(s0 + (w2 + 1) * p.ld_resid) + sub_c4), p.f16);
                                    t = make_float4(((a.x + b.x) + (d.x + e.x)) * 0.25f, ((a.y + b.y) + (d.y + e.y)) * 0.25f,
                                                    ((a.z + b.z) + (d.z + e.z)) * 0.25f, ((a.w + b.w) + (d.w + e.w)) * 0.25f);
                                }
                            }
                            *reinterpret_cast<float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2)) = t;
                        }
                    } else if (p.resid_mode == 1) {
                        // resampled skip path: the residual lives at half resolution (nearest x2)
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + sub_r4;
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((vmask >> r) & 1) {
                                const size_t q = pix0 + r;
                                const size_t wo = q & (static_cast<size_t>(p.Wo) - 1);
                                const size_t ho = (q >> p.log2_wo) & (static_cast<size_t>(p.Ho) - 1);
                                const size_t nn = q >> (p.log2_wo + p.log2_ho);
                                const size_t src = (nn * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1);
                                t = __ldg(reinterpret_cast<const float4*>(p.resid + src * p.ld_resid + col0) + sub_c4);
                            }
                            *reinterpret_cast<float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2)) = t;
                        }
                    } else {
                        // ... at double resolution (2x2 average; same association as avg_pool2d / resample_kernel)
#pragma unroll 2
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + sub_r4;
                            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((vmask >> r) & 1) {
                                const size_t q = pix0 + r;
                                const size_t wo = q & (static_cast<size_t>(p.Wo) - 1);
                                const size_t ho = (q >> p.log2_wo) & (static_cast<size_t>(p.Ho) - 1);
                                const size_t nn = q >> (p.log2_wo + p.log2_ho);
                                const size_t w2 = 2 * static_cast<size_t>(p.Wo);
                                const float* s0 = p.resid + ((nn * 2 * p.Ho + 2 * ho) * w2 + 2 * wo) * p.ld_resid + col0;
                                const float4 a = __ldg(reinterpret_cast<const float4*>(s0) + sub_c4);
                                const float4 b = __ldg(reinterpret_cast<const float4*>(s0 + p.ld_resid) + sub_c4);
                                const float4 d = __ldg(reinterpret_cast<const float4*>(s0 + w2 * p.ld_resid) + sub_c4);
                                const float4 e = __ldg(reinterpret_cast<const float4*>(s0 + (w2 + 1) * p.ld_resid) + sub_c4);
                                t = make_float4(((a.x + b.x) + (d.x + e.x)) * 0.25f, ((a.y + b.y) + (d.y + e.y)) * 0.25f,
                                                ((a.z + b.z) + (d.z + e.z)) * 0.25f, ((a.w + b.w) + (d.w + e.w)) * 0.25f);
                            }
                            *reinterpret_cast<float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2)) = t;
                        }
                    }
                    __syncwarp();
                }
                float f[32];
                if (X3) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = (kNJ == 2 && c >= 32 * (kEpiWarps / 4)) ? accr[kNJ - 1][i] : accr[0][i];
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                }
                if (p.bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = __ldg(b4 + i);
                        f[4 * i] += t.x, f[4 * i + 1] += t.y, f[4 * i + 2] += t.z, f[4 * i + 3] += t.w;
                    }
                }
                if (p.rowvec && valid) {
                    const float4* b4 =
                        reinterpret_cast<const float4*>(p.rowvec + static_cast<size_t>(n) * p.ld_rowvec + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = __ldg(b4 + i);
                        f[4 * i] += t.x, f[4 * i + 1] += t.y, f[4 * i + 2] += t.z, f[4 * i + 3] += t.w;
                    }
                }
                if (p.resid) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 t = *reinterpret_cast<const float4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2));
                        f[4 * i] += t.x, f[4 * i + 1] += t.y, f[4 * i + 2] += t.z, f[4 * i + 3] += t.w;
                    }
                    __syncwarp();
                }
                if (p.out_scale != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] *= p.out_scale;
                }
                if (p.act) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                if (p.stats && vmask == 0xffffffffu)  // (the odd tail tile of a CTA pair is entirely out of range)
                    gn_partials(f, lane, p.stats + (stat_blk * p.stats_nblk + (col0 >> 2)) * 2);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((i ^ (lane & 7)) << 2)) =
                        make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                __syncwarp();
                if (p.out_f32) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + sub_r4;
                        if ((vmask >> r) & 1)
                            reinterpret_cast<float4*>(of32 + opix(r) * p.ld_out_f32 + ocol0)[sub_c4] =
                                *reinterpret_cast<const float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2));
                    }
                }
                if (p.out_op) {
                    if (TF32) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + sub_r4;
                            if ((vmask >> r) & 1) {
                                const float4 t = *reinterpret_cast<const float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2));
                                reinterpret_cast<float4*>(static_cast<float*>(p.out_op) + opix(r) * p.ld_out_op +
                                                          ocol0)[sub_c4] =
                                    X3 ? t : make_float4(round_tf32(t.x), round_tf32(t.y), round_tf32(t.z),
                                                         round_tf32(t.w));
                            }
                        }
                    } else {
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int r = it * 8 + sub_r8;
                            if ((vmask >> r) & 1) {
                                const float4 a =
                                    *reinterpret_cast<const float4*>(stg + r * 32 + (((2 * sub_c8) ^ (r & 7)) << 2));
                                const float4 b =
                                    *reinterpret_cast<const float4*>(stg + r * 32 + (((2 * sub_c8 + 1) ^ (r & 7)) << 2));
                                reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_op) +
                                                         opix(r) * p.ld_out_op + ocol0)[sub_c8] =
                                    make_uint4(pack_op16x2(a.x, a.y, p.f16), pack_op16x2(a.z, a.w, p.f16),
                                               pack_op16x2(b.x, b.y, p.f16), pack_op16x2(b.z, b.w, p.f16));
                            }
                        }
                    }
                }
                __syncwarp();
            }
            }
            if (!X3) {
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (CTA2) mbar_arrive_remote_relaxed(&tempty[acc], 0); else mbar_arrive_relaxed(&tempty[acc]);
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
        if (TEPI == 1 && lane == 0) bulk_wait_group0();  // this warp's tensor stores are complete before the CTA retires
    }

    tc_fence_before_sync();
    __syncthreads();
    if (CTA2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while the peer still uses its smem / barriers
    if (warp == 2) {
        tc_fence_after_sync();
        if (CTA2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base); else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------- split-K
// Sum of the splits' raw accumulators (fixed order: deterministic) + the epilogue the main kernel would have applied, in the
// same order: bias, per-sample row, residual (fp32 or 16-bit), scale, ReLU -> fp32 and / or operand-dtype outputs.
struct SplitKReduce {
    const float* ws;
    long long stride;
    int ksplit, Cout, HW;
    long long npix;
    const float* bias;
    const float* rowvec;
    int ld_rowvec;
    const void* resid;
    int ld_resid, resid16, f16;
    float out_scale;
    int act;
    float* out_f32;
    int ld_out_f32;
    void* out_op;
    int ld_out_op, op_is_f32;  // operand output: 16-bit (f16 / bf16), or fp32 rounded to tf32
};
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const SplitKReduce a) {
    pdl_wait();
    pdl_trigger();
    const int c4n = a.Cout >> 2;
    const long long total = a.npix * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = i / c4n;
        const int c = static_cast<int>(i - pix * c4n) << 2;
        const float* w = a.ws + pix * a.Cout + c;
        float4 f = *reinterpret_cast<const float4*>(w);
        for (int s = 1; s < a.ksplit; ++s) {
            const float4 t = *reinterpret_cast<const float4*>(w + s * a.stride);
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
        }
        if (a.bias) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(a.bias + c));
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
        }
        if (a.rowvec) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(a.rowvec + (pix / a.HW) * a.ld_rowvec + c));
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
        }
        if (a.resid) {
            float4 t;
            if (a.resid16)
                t = unpack_op16x4(__ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(a.resid) +
                                                                         pix * a.ld_resid + c)), a.f16);
            else
                t = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(a.resid) + pix * a.ld_resid + c));
            f.x += t.x, f.y += t.y, f.z += t.z, f.w += t.w;
        }
        if (a.out_scale != 1.0f) f.x *= a.out_scale, f.y *= a.out_scale, f.z *= a.out_scale, f.w *= a.out_scale;
        if (a.act) f.x = fmaxf(f.x, 0.f), f.y = fmaxf(f.y, 0.f), f.z = fmaxf(f.z, 0.f), f.w = fmaxf(f.w, 0.f);
        if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + pix * a.ld_out_f32 + c) = f;
        if (a.out_op) {
            if (a.op_is_f32)
                *reinterpret_cast<float4*>(static_cast<float*>(a.out_op) + pix * a.ld_out_op + c) =
                    make_float4(round_tf32(f.x), round_tf32(f.y), round_tf32(f.z), round_tf32(f.w));
            else
                *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(a.out_op) + pix * a.ld_out_op + c) =
                    make_uint2(pack_op16x2(f.x, f.y, a.f16), pack_op16x2(f.z, f.w, a.f16));
        }
    }
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }
bool conv_slab_eligible(const nlc_ctx* ctx, const nlc_conv_desc* d, int chunk);  // conv_slab.cu

// Decide whether a launch takes the TMA epilogue and encode its output / residual tensor maps.  Called once the tile
// geometry (BW, BH, BN) is final: by nlc_conv_tc for the tap-per-tile kernel, by launch_conv_slab for the slab kernel.
int epi_tma_setup(nlc_ctx* ctx, const nlc_conv_desc* d, ConvKParams& p) {
    p.tma_epi = 0;
    if (!ctx->use_tma_epi || !dtype_is16(d->dtype) || !d->out_op || d->out_f32 || d->out_up || d->out_head_split) return NLC_OK;
    if (d->resid && (!d->resid_is_op || d->resid_mode != 0)) return NLC_OK;
    {
        static const int mask = [] { const char* e = getenv("NLC_TMA_EPI_MASK"); return e ? atoi(e) : 15; }();  // debugging
        if (!(mask & (d->resid ? 2 : 1))) return NLC_OK;
        const bool is_slab = conv_slab_eligible(ctx, d, 64);
        if (!(mask & (is_slab ? 4 : 8))) return NLC_OK;
    }
    if (!is_pow2(d->Wo) || !is_pow2(d->Ho) || d->Cout % 32 != 0) return NLC_OK;
    int bw, bh, bn;
    epi_tma_box(p.BW, p.BH, &bw, &bh, &bn);
    for (int which = 0; which < (d->resid ? 2 : 1); ++which) {
        const void* base = which ? d->resid : d->out_op;
        const cuuint64_t ld = which ? d->ld_resid : d->ld_out_op;
        if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return NLC_OK;
        cuuint64_t gdim[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->B};
        cuuint64_t gstr[3] = {ld * 2, ld * 2 * d->Wo, ld * 2 * d->Wo * d->Ho};
        cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        // (L2 promotion: the residual's 64-byte rows are fetched as whole 128-byte lines - the other half belongs to the
        // neighbouring warp's block; none for the output map.  Measured: DRAM bytes and step time are the same either way,
        // profiles/r02zk_dram_c5_promo?.csv)
        CUresult r = ctx->encode_tiled(which ? &p.mapRes : &p.mapOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                       const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_64B,
                                       which ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_conv_tc: cuTensorMapEncodeTiled(%s) failed with %d", which ? "residual" : "output",
                    (int)r);
    }
    if (!d->resid) p.mapRes = p.mapOut;
    p.tma_epi = 1;
    if (ctx->use_tma_epi == 2) {  // 256-bit accesses: every pixel row of the output (and residual) slice on a 32-byte boundary
        const bool ok = (reinterpret_cast<uintptr_t>(d->out_op) & 31) == 0 && (d->ld_out_op * 2) % 32 == 0 &&
                        (!d->resid || ((reinterpret_cast<uintptr_t>(d->resid) & 31) == 0 && (d->ld_resid * 2) % 32 == 0));
        if (ok) p.tma_epi = 2;
    }
    return NLC_OK;
}

// conv_slab.cu
bool conv_slab_eligible(const nlc_ctx* ctx, const nlc_conv_desc* d, int chunk);
int launch_conv_slab(nlc_ctx* ctx, const nlc_conv_desc* d, ConvKParams& p, int chunk, bool tf32, cudaStream_t stream);

template <int BLOCK_N, int MODE, bool CTA2, int TEPI = 0>
static int launch_conv(const ConvKParams& p, int grid, cudaStream_t stream, int device) {
    using Cfg = ConvCfg<BLOCK_N, CTA2, MODE == 2>;
    constexpr int kLaunchThreads = MODE == 2 ? kThreadsX3 : kThreads;
    if constexpr (MODE == 0 && TEPI == 0) {
        if (p.tma_epi == 1) return launch_conv<BLOCK_N, MODE, CTA2, 1>(p, grid, stream, device);
        if (p.tma_epi == 2) return launch_conv<BLOCK_N, MODE, CTA2, 2>(p, grid, stream, device);
    }
    static PerDeviceFlag configured;
    if (!configured[device]) {
        NLC_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, CTA2, TEPI>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        configured[device] = true;
    }
    if (CTA2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kLaunchThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes, cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
        NLC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, MODE, CTA2, TEPI>, p));
    } else {
        NLC_CHECK_CUDA(launch_pdl((conv_tc_kernel<BLOCK_N, MODE, CTA2, TEPI>), dim3(grid), dim3(kLaunchThreads), Cfg::kSmemBytes,
                                  stream, p));
    }
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

}  // namespace nlc

using namespace nlc;

extern "C" int nlc_conv_tc(nlc_ctx* ctx, const nlc_conv_desc* d, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    NLC_REQUIRE(ctx && d, "nlc_conv_tc: null argument");
    NLC_REQUIRE_DEVICE(ctx);
    NLC_REQUIRE(dtype_valid(d->dtype), "nlc_conv_tc: dtype must be NLC_BF16, NLC_F16, NLC_F32 or NLC_F32X3");
    const bool x3 = d->dtype == NLC_F32X3;
    const bool tf32 = !dtype_is16(d->dtype);  // fp32 containers
    const int esz = tf32 ? 4 : 2;
    const int chunk = kChunkBytes / esz;
    NLC_REQUIRE(d->nsrc >= 1 && d->nsrc <= NLC_MAX_SRC, "nlc_conv_tc: nsrc %d out of range", d->nsrc);
    NLC_REQUIRE(d->nseg >= 1 && d->nseg <= NLC_MAX_SEG, "nlc_conv_tc: nseg %d out of range", d->nseg);
    NLC_REQUIRE(d->stride == 1 || d->stride == 2, "nlc_conv_tc: stride %d unsupported", d->stride);
    // (rows of >= 128 pixels tile as BH = 1: Ho is then a plain count - e.g. the three 64-channel heads of a 192-channel
    // attention block in the unfused path - and only the modes that index pixels by shifts need a power of two)
    NLC_REQUIRE(is_pow2(d->Wo) && (is_pow2(d->Ho) || (d->Wo >= kBlockM && !d->out_up && d->resid_mode == 0)),
                "nlc_conv_tc: output %dx%d must be powers of two", d->Ho, d->Wo);
    NLC_REQUIRE(d->B >= 1 && d->Cout % 64 == 0, "nlc_conv_tc: Cout %d must be a multiple of 64", d->Cout);
    NLC_REQUIRE(d->out_f32 || d->out_op, "nlc_conv_tc: no output requested");
    NLC_REQUIRE(!d->out_f32 || (d->ld_out_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(d->out_f32) & 15) == 0),
                "nlc_conv_tc: out_f32 must be 16-byte aligned with ld %% 4 == 0");
    NLC_REQUIRE(!d->out_op || ((d->ld_out_op * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(d->out_op) & 15) == 0),
                "nlc_conv_tc: out_op must be 16-byte aligned");
    NLC_REQUIRE(!d->resid || (d->ld_resid % 4 == 0 && (reinterpret_cast<uintptr_t>(d->resid) & 15) == 0),
                "nlc_conv_tc: resid must be 16-byte aligned with ld %% 4 == 0");
    NLC_REQUIRE(!d->resid_is_op || (d->resid && dtype_is16(d->dtype) && d->ld_resid % 8 == 0),
                "nlc_conv_tc: resid_is_op needs a residual, a 16-bit operand mode and ld_resid %% 8 == 0");
    NLC_REQUIRE(!d->rowvec || (d->ld_rowvec % 4 == 0 && (reinterpret_cast<uintptr_t>(d->rowvec) & 15) == 0),
                "nlc_conv_tc: rowvec must be 16-byte aligned with ld %% 4 == 0");
    NLC_REQUIRE(!d->bias || (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0, "nlc_conv_tc: bias must be 16-byte aligned");

    ConvKParams p;
    memset(&p, 0, sizeof(p));
    p.B = d->B, p.Ho = d->Ho, p.Wo = d->Wo, p.stride = d->stride;
    p.BW = d->Wo < kBlockM ? d->Wo : kBlockM;
    p.BH = d->Ho < kBlockM / p.BW ? d->Ho : kBlockM / p.BW;
    p.BN = kBlockM / (p.BW * p.BH);
    p.tiles_w = d->Wo / p.BW;
    p.tiles_h = d->Ho / p.BH;
    p.tiles_n = (d->B + p.BN - 1) / p.BN;
    p.num_m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.Cout = d->Cout;

    // Tile configuration: BLOCK_N in {64, 128, 256} x {one CTA, CTA pair}, chosen by a small cost model instead of "the
    // widest tile that still gives every SM one".  Cost of a launch ~ waves x time per K chunk, where a chunk costs
    // max(MMA time = 2*BLOCK_N clocks, operand latency / ring depth): with ~3000 clocks of TMA latency under load the
    // narrow tiles are bound by bytes in flight, so two waves of 64-wide tiles lose to one (partial) wave of CTA pairs.
    // (The batched right-hand operand of the attention GEMMs differs per M tile and stays on the 1-CTA kernel; the
    // fp32 split mode has 1-CTA kernels for N <= 128 only.)
    int block_n = 64;
    bool pair = false;
    {
        double best = 1e300;
        for (int bn = 64; bn <= 256; bn *= 2) {
            if (d->Cout % bn != 0 || (x3 && bn == 256)) continue;
            for (int pr = 0; pr < 2; ++pr) {
                if (pr && (x3 || !ctx->use_cta_pairs || d->wbatched.ptr != nullptr || bn < 128)) continue;
                const long long units = (pr ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * static_cast<long long>(d->Cout / bn);
                const int slots = pr ? ctx->sm_count / 2 : ctx->sm_count;
                const long long waves = (units + slots - 1) / slots;
                const int stages = x3 ? (bn == 128 ? 3 : 4) : (pr ? (bn == 256 ? 6 : 8) : (bn == 256 ? 4 : (bn == 128 ? 6 : 8)));
                const double mma = (x3 ? 6.0 : (tf32 ? 4.0 : 2.0)) * bn;  // clocks per 128-byte K chunk per CTA
                const double chunk = mma > 3000.0 / stages ? mma : 3000.0 / stages;
                // prefer the wider tile on ties (fewer operand bytes per FLOP)
                const double cost = static_cast<double>(waves) * chunk * (1.0 - 1e-3 * (bn / 64) - 1e-3 * pr);
                if (cost < best) best = cost, block_n = bn, pair = pr != 0;
            }
        }
    }
    // 3x3 stride-1 layers of small images: the halo-slab kernel (conv_slab.cu; CTA pairs, 128-wide accumulators)
    const bool slab = conv_slab_eligible(ctx, d, chunk);
    if (slab) block_n = 128, pair = true;
    p.num_n_tiles = d->Cout / block_n;
    p.num_m_units = pair ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;
    p.num_tiles = p.num_m_units * p.num_n_tiles;

    int ktot = 0;
    p.nseg = d->nseg;
    for (int s = 0; s < d->nseg; ++s) {
        const nlc_kseg& sg = d->seg[s];
        NLC_REQUIRE(sg.src >= 0 && sg.src < d->nsrc, "nlc_conv_tc: segment %d names source %d", s, sg.src);
        NLC_REQUIRE(sg.nch > 0 && sg.nch % chunk == 0 && sg.c0 % 8 == 0 && sg.c0 + sg.nch <= d->src[sg.src].C,
                    "nlc_conv_tc: segment %d covers channels [%d,%d) of %d; need multiples of %d", s, sg.c0,
                    sg.c0 + sg.nch, d->src[sg.src].C, chunk);
        p.seg[s] = ConvSegDev{sg.src, sg.dh, sg.dw, sg.c0, sg.nch / chunk};
        ktot += sg.nch;
    }
    p.total_chunks = ktot / chunk;
    p.ksplit = 1, p.chunks_per_split = p.total_chunks, p.num_units = p.num_tiles, p.split_stride = 0;

    // (TMA only moves the 16-bit elements: the bf16 element type serves fp16 tensors as well)
    const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    p.f16 = d->dtype == NLC_F16;
    for (int s = 0; s < d->nsrc; ++s) {
        const nlc_operand& o = d->src[s];
        NLC_REQUIRE(o.ptr && (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0 && (static_cast<size_t>(o.ld) * esz) % 16 == 0,
                    "nlc_conv_tc: source %d must be 16-byte aligned (ptr and row pitch)", s);
        NLC_REQUIRE(o.B == d->B, "nlc_conv_tc: source %d batch %d != %d", s, o.B, d->B);
        if (slab) continue;  // (launch_conv_slab encodes the slab boxes)
        NLC_REQUIRE(o.ptr && (reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0 && (static_cast<size_t>(o.ld) * esz) % 16 == 0,
                    "nlc_conv_tc: source %d must be 16-byte aligned (ptr and row pitch)", s);
        NLC_REQUIRE(o.B == d->B, "nlc_conv_tc: source %d batch %d != %d", s, o.B, d->B);
        cuuint64_t gdim[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)o.B};
        const cuuint64_t sh = o.sh ? (cuuint64_t)o.sh : (cuuint64_t)o.W * o.ld;
        const cuuint64_t sn = o.sn ? (cuuint64_t)o.sn : (cuuint64_t)o.H * o.W * o.ld;
        NLC_REQUIRE((sh * esz) % 16 == 0 && (sn * esz) % 16 == 0, "nlc_conv_tc: source %d strides must be 16-byte multiples", s);
        cuuint64_t gstr[3] = {(cuuint64_t)o.ld * esz, sh * esz, sn * esz};
        cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)(p.BW * d->stride), (cuuint32_t)(p.BH * d->stride),
                             (cuuint32_t)p.BN};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = ctx->encode_tiled(&p.mapA[s], dt, 4, const_cast<void*>(o.ptr), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_conv_tc: cuTensorMapEncodeTiled(A%d) failed with %d", s, (int)r);
    }
    for (int s = d->nsrc; s < NLC_MAX_SRC; ++s) p.mapA[s] = p.mapA[0];
    p.w_batched = d->wbatched.ptr != nullptr;
    if (p.w_batched) {
        const nlc_operand& o = d->wbatched;
        NLC_REQUIRE(p.BW == kBlockM && d->stride == 1, "nlc_conv_tc: batched operand needs Wo >= 128 and stride 1");
        NLC_REQUIRE(o.C == ktot && o.W == d->Cout && o.H == d->Ho && o.B == d->B,
                    "nlc_conv_tc: batched operand shape [%d,%d,%d,%d] does not match K=%d Cout=%d Ho=%d B=%d", o.B, o.H,
                    o.W, o.C, ktot, d->Cout, d->Ho, d->B);
        const cuuint64_t sh = o.sh ? (cuuint64_t)o.sh : (cuuint64_t)o.W * o.ld;
        const cuuint64_t sn = o.sn ? (cuuint64_t)o.sn : (cuuint64_t)o.H * o.W * o.ld;
        NLC_REQUIRE((reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0 && ((cuuint64_t)o.ld * esz) % 16 == 0 &&
                        (sh * esz) % 16 == 0 && (sn * esz) % 16 == 0,
                    "nlc_conv_tc: batched operand must be 16-byte aligned");
        cuuint64_t gdim[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)o.B};
        cuuint64_t gstr[3] = {(cuuint64_t)o.ld * esz, sh * esz, sn * esz};
        cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)block_n, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ctx->encode_tiled(&p.mapB, dt, 4, const_cast<void*>(o.ptr), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_conv_tc: cuTensorMapEncodeTiled(batched W) failed with %d", (int)r);
    } else {
        NLC_REQUIRE(d->weight && (reinterpret_cast<uintptr_t>(d->weight) & 15) == 0, "nlc_conv_tc: weight unaligned");
        cuuint64_t gdim[4] = {(cuuint64_t)ktot, (cuuint64_t)d->Cout, 1, 1};
        cuuint64_t gstr[3] = {(cuuint64_t)ktot * esz, (cuuint64_t)ktot * esz * d->Cout, (cuuint64_t)ktot * esz * d->Cout};
        cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)(pair ? block_n / 2 : block_n), 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ctx->encode_tiled(&p.mapB, dt, 4, const_cast<void*>(d->weight), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "nlc_conv_tc: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    p.out_head_split = d->out_head_split;
    NLC_REQUIRE(d->out_head_split % 8 == 0, "nlc_conv_tc: out_head_split must be a multiple of 8");
    NLC_REQUIRE(d->out_up >= 0 && d->out_up <= 4 &&
                    (d->out_up == 0 || (is_pow2(d->Ho) && is_pow2(d->Wo) && !d->out_head_split && !d->resid &&
                                        (d->Ho * d->Wo) % 32 == 0)),
                "nlc_conv_tc: out_up (sub-pixel placement) needs power-of-two extents, a dense output and no residual");
    p.out_up = d->out_up;
    if (d->stats) {
        NLC_REQUIRE(p.BN == 1 && d->out_head_split == 0 && d->stats_nblk >= d->Cout / 4 &&
                        (reinterpret_cast<uintptr_t>(d->stats) & 7) == 0,
                    "nlc_conv_tc: fused GroupNorm partials need Ho*Wo >= 128, dense NHWC output and an 8-byte aligned "
                    "buffer of >= Cout/4 blocks per 32 pixels");
        p.stats = d->stats, p.stats_nblk = d->stats_nblk;
    }
    NLC_REQUIRE(d->resid_mode >= 0 && d->resid_mode <= 2 && (d->resid_mode == 0 || (d->resid && !d->out_head_split)) &&
                    (d->resid_mode != 1 || (d->Ho % 2 == 0 && d->Wo % 2 == 0)),
                "nlc_conv_tc: resid_mode %d needs a residual, a dense output and (mode 1) even extents", d->resid_mode);
    p.resid_mode = d->resid_mode;
    for (p.log2_wo = 0; (1 << p.log2_wo) < d->Wo; ++p.log2_wo) {}
    for (p.log2_ho = 0; (1 << p.log2_ho) < d->Ho; ++p.log2_ho) {}
    p.bias = d->bias, p.rowvec = d->rowvec, p.ld_rowvec = d->ld_rowvec;
    p.resid = static_cast<const float*>(d->resid), p.ld_resid = d->ld_resid, p.out_scale = d->out_scale;
    p.resid16 = d->resid_is_op != 0;
    NLC_REQUIRE(d->act == 0 || d->act == 1, "nlc_conv_tc: act must be 0 (none) or 1 (ReLU)");
    p.act = d->act;
    p.out_f32 = d->out_f32, p.ld_out_f32 = d->ld_out_f32, p.out_op = d->out_op, p.ld_out_op = d->ld_out_op;

    if (slab) return launch_conv_slab(ctx, d, p, chunk, tf32, stream);
    {
        const int rc = epi_tma_setup(ctx, d, p);
        if (rc != NLC_OK) return rc;
    }
    // Split-K for the small-M layers (4x4 / 8x8 levels, small batches): when the tiles fill less than half of the SMs and the
    // K loop is long, S splits share a tile's K range; their raw accumulators go to a workspace and splitk_reduce_kernel
    // applies the epilogue.  Deterministic (fixed summation order); the plain epilogues only (no GroupNorm partials,
    // sub-pixel placement, head merge or resampled residual - none of which occurs at these levels).
    SplitKReduce red;
    memset(&red, 0, sizeof(red));
    {
        const int slots = pair ? ctx->sm_count / 2 : ctx->sm_count;
        const long long npix = static_cast<long long>(d->B) * d->Ho * d->Wo;
        // (at least 8 K chunks per split, tiles on at most half of the SM slots: fewer chunks per split or fuller grids measured
        //  slower at batch 32 - 4.72 against 4.65 ms per timestep - and no different at batch 256)
        if (ctx->use_splitk && !x3 && !d->stats && !d->out_up && !d->out_head_split && d->resid_mode == 0 && !p.w_batched &&
            p.num_tiles * 2 <= slots && p.total_chunks >= 16 && d->Cout % 4 == 0) {
            int sp = slots / p.num_tiles;
            if (sp > p.total_chunks / 8) sp = p.total_chunks / 8;
            if (sp > 16) sp = 16;
            // The workspace is ONE fixed allocation per context, made at the first eligible launch outside graph capture and
            // never moved or freed before nlc_destroy: captured graphs hold its address.  A launch whose splits do not fit
            // takes fewer splits (or none).
            constexpr size_t kSplitKBytes = 64u << 20;
            if (!ctx->splitk_ws) {
                cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
                cudaStreamIsCapturing(stream, &cs);
                if (cs == cudaStreamCaptureStatusNone && cudaMalloc(&ctx->splitk_ws, kSplitKBytes) == cudaSuccess)
                    ctx->splitk_bytes = kSplitKBytes;
                else
                    ctx->splitk_ws = nullptr, (void)cudaGetLastError();
            }
            const size_t per_split = static_cast<size_t>(npix) * d->Cout * sizeof(float);
            if (static_cast<size_t>(sp) * per_split > ctx->splitk_bytes) sp = static_cast<int>(ctx->splitk_bytes / per_split);
            if (sp >= 2) {
                const int cps = (p.total_chunks + sp - 1) / sp;
                sp = (p.total_chunks + cps - 1) / cps;
                const bool ok = ctx->splitk_ws != nullptr;
                if (ok && sp >= 2) {
                    red.ws = static_cast<const float*>(ctx->splitk_ws), red.stride = npix * d->Cout, red.ksplit = sp;
                    red.Cout = d->Cout, red.HW = d->Ho * d->Wo, red.npix = npix;
                    red.bias = p.bias, red.rowvec = p.rowvec, red.ld_rowvec = p.ld_rowvec;
                    red.resid = p.resid, red.ld_resid = p.ld_resid, red.resid16 = p.resid16, red.f16 = p.f16;
                    red.out_scale = p.out_scale, red.act = p.act;
                    red.out_f32 = p.out_f32, red.ld_out_f32 = p.ld_out_f32;
                    red.out_op = p.out_op, red.ld_out_op = p.ld_out_op, red.op_is_f32 = tf32 ? 1 : 0;
                    p.bias = nullptr, p.rowvec = nullptr, p.resid = nullptr, p.resid16 = 0, p.out_scale = 1.0f, p.act = 0;
                    p.out_op = nullptr, p.out_f32 = static_cast<float*>(ctx->splitk_ws), p.ld_out_f32 = d->Cout, p.tma_epi = 0;
                    p.ksplit = sp, p.chunks_per_split = cps, p.num_units = p.num_tiles * sp, p.split_stride = red.stride;
                }
            }
        }
    }
    const int rc_main = [&]() -> int {
    if (pair) {
        const int pairs = p.num_units < ctx->sm_count / 2 ? p.num_units : ctx->sm_count / 2;
        if (tf32) {
            if (block_n == 256) return launch_conv<256, 1, true>(p, 2 * pairs, stream, ctx->device);
            return launch_conv<128, 1, true>(p, 2 * pairs, stream, ctx->device);
        }
        if (block_n == 256) return launch_conv<256, 0, true>(p, 2 * pairs, stream, ctx->device);
        return launch_conv<128, 0, true>(p, 2 * pairs, stream, ctx->device);
    }
    const int grid = p.num_units < ctx->sm_count ? p.num_units : ctx->sm_count;
    if (x3) {
        if (block_n == 128) return launch_conv<128, 2, false>(p, grid, stream, ctx->device);
        return launch_conv<64, 2, false>(p, grid, stream, ctx->device);
    }
    if (tf32) {
        if (block_n == 256) return launch_conv<256, 1, false>(p, grid, stream, ctx->device);
        if (block_n == 128) return launch_conv<128, 1, false>(p, grid, stream, ctx->device);
        return launch_conv<64, 1, false>(p, grid, stream, ctx->device);
    }
    if (block_n == 256) return launch_conv<256, 0, false>(p, grid, stream, ctx->device);
    if (block_n == 128) return launch_conv<128, 0, false>(p, grid, stream, ctx->device);
    return launch_conv<64, 0, false>(p, grid, stream, ctx->device);
    }();
    if (rc_main != NLC_OK || red.ksplit < 2) return rc_main;
    {
        const long long work = red.npix * (red.Cout >> 2);
        long long blocks = (work + 255) / 256;
        const long long cap = 8LL * ctx->sm_count;
        if (blocks > cap) blocks = cap;
        NLC_CHECK_CUDA(launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, red));
        NLC_CHECK_LAUNCH();
    }
    return NLC_OK;
}
