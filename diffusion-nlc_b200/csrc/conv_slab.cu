// 3x3 stride-1 convolution as an implicit GEMM whose ACTIVATION operand is staged as halo slabs instead of one TMA
// tile per filter tap (conv_tc.cu).  Why: the tap-per-tile kernel is bound by bytes through L2, not by the tensor pipe -
// for the 128-channel 64x64 layers of the c2 / c3 networks it moves 3.4 GB of operands L2 -> SM per launch against 1.07 GB
// of algorithmic traffic and runs at ~10-11 TB/s of combined L2 traffic whatever the epilogue does
// (profiles/r02c_ncu_conv_c2_notes.md).  Two things cut those bytes here:
//
//   * a CTA owns G vertically adjacent M tiles (G x BH image rows, BH = 128 / W rows of W pixels each): for one
//     64-channel chunk and one horizontal tap dw it loads ONE slab of (G*BH + 2) image rows x W pixels (TMA box shifted by
//     dw columns: the zero padding is TMA out-of-bounds fill).  Image rows are contiguous in the slab (no horizontal halo),
//     so the A operand of vertical tap dh for M tile g is simply the 128 pixel rows starting (g*BH + dh + 1) * W rows into
//     the slab - a UMMA descriptor offset that is a multiple of 1024 B.  One slab feeds 3 taps x G tiles: the activation
//     traffic per output row drops from 9 row loads to 3 * (G*BH + 2) / (G*BH) (4.5 at W = 64, G = 2);
//   * the G tiles share every weight tile, and the CTA pair (cta_group::2) splits it: 8 KB of weights per
//     2 x 128 x 128 x 64 MMA chunk per SM instead of 16.
//
// Warp roles as in conv_tc.cu: warp 0 = TMA producer of the slabs, warp 10 = TMA producer of the weight tiles, warp 1 =
// MMA issuer (leader CTA), warps 2-9 = epilogue.  Two rings (slabs, weight tiles), two sets of G TMEM accumulators so the
// epilogue of one unit overlaps the main loop of the next.  Serves NLC_BF16 / NLC_F16 / NLC_F32 (tf32) operands.
#include "conv_common.cuh"

namespace nlc {

int epi_tma_setup(nlc_ctx* ctx, const nlc_conv_desc* d, ConvKParams& p);  // conv_tc.cu

constexpr int kSlabN = 128;          // accumulator width (output channels per unit)
constexpr int kSlabG = 2;            // M tiles per CTA
constexpr int kSlabStages = 3;       // slab ring
constexpr int kSlabWStages = 6;      // weight-tile ring
constexpr int kSlabWBytes = (kSlabN / 2) * kChunkBytes;  // this CTA's half of a weight tile: 8 KB
constexpr int kSlabBarBytes = 512;    // 22 mbarriers, the TMEM slot, kEpiWarps residual-block barriers (tma_epi); 512: the
                                      // staging blocks behind it stay aligned for SWIZZLE_64B tensor loads / stores
constexpr int kSlabEpiBytes = kEpiWarps * 32 * 32 * 4;
constexpr int kSlabTmemCols = 2 * kSlabG * kSlabN;  // 512
constexpr int kSlabMaxSteps = 64;    // slabs per unit: 3 per 64-channel chunk (+ 1 per chunk of a fused 1x1 shortcut)

struct SlabStep {  // one slab: source map, horizontal shift, channel chunk, and the weight K coordinates of its taps
    int map, dw, c_elem, ntap;
    int k0, k1, k2;
};

struct SlabParams {
    ConvKParams k;        // maps (mapA boxes are slabs here), epilogue parameters, tile geometry (BW = W, BH = 128 / W)
    int slab_bytes;       // (G*BH + 2) * W * 128
    int row_bytes;        // W * 128: one image row inside a slab
    int sb_per_img;       // super tiles (G*BH rows) per image
    int num_super;        // B * sb_per_img
    int num_units;        // pair units x n tiles
    int num_pair_units;
    int nstep;
    int nstep3;           // the first nstep3 slabs carry three vertical taps (3x3 part), the rest one (fused 1x1 shortcut): the
                          // MMA issuer takes the tap count from here - a kernel parameter, i.e. provably warp-uniform
    const SlabStep* steps;  // device array [nstep]
};

template <int MODE, int TEPI = 0>
__global__ void __launch_bounds__(kThreads, 1) conv_slab_kernel(const __grid_constant__ SlabParams sp) {
    static_assert(TEPI == 0 || MODE == 0, "the 16-bit-only epilogues (1 TMA, 2 256-bit accesses) serve the 16-bit operand modes");
    constexpr bool TF32 = MODE != 0;
    const ConvKParams& p = sp.k;
    const uint32_t rank = cluster_ctarank();
    const int unit0 = static_cast<int>(blockIdx.x >> 1);
    const int unit_step = static_cast<int>(gridDim.x >> 1);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* slab_base = smem;
    uint8_t* w_base = smem + kSlabStages * sp.slab_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_base + kSlabWStages * kSlabWBytes);
    uint64_t* s_full = bars;
    uint64_t* s_empty = s_full + kSlabStages;
    uint64_t* w_full = s_empty + kSlabStages;
    uint64_t* w_empty = w_full + kSlabWStages;
    uint64_t* tfull = w_empty + kSlabWStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint64_t* rbar = tempty + 4;  // [kEpiWarps]
    float* stg_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kSlabBarBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < NLC_MAX_SRC; ++i) tma_prefetch_desc(&p.mapA[i]);
        tma_prefetch_desc(&p.mapB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSlabStages; ++s) mbar_init(&s_full[s], 1), mbar_init(&s_empty[s], 1);
        for (int s = 0; s < kSlabWStages; ++s) mbar_init(&w_full[s], 1), mbar_init(&w_empty[s], 1);
        for (int a = 0; a < 2; ++a) mbar_init(&tfull[a], 1), mbar_init(&tempty[a], 2 * kEpiWarps);
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(&rbar[w], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair<kSlabTmemCols>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // (programmatic dependent launch, conv_tc.cu)
    pdl_trigger();

    // CTA `rank` of pair unit u owns super tile 2u + rank = (image, block of G*BH rows)
    auto super_of = [&](int unit, int& n_tile) {
        n_tile = unit / sp.num_pair_units;
        return 2 * (unit - n_tile * sp.num_pair_units) + static_cast<int>(rank);
    };

    if (warp == 0) {
        // ------------------------------------------------------------ slab producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = unit0; unit < sp.num_units; unit += unit_step) {
                int n_tile;
                const int s = super_of(unit, n_tile);
                const int n = s / sp.sb_per_img;
                const int h0 = (s - n * sp.sb_per_img) * (kSlabG * p.BH);
                for (int i = 0; i < sp.nstep; ++i) {
                    const SlabStep st = sp.steps[i];
                    mbar_wait(&s_empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&s_full[stage], 2 * sp.slab_bytes);
                    // (a super tile past the end - odd count - has n >= B: the whole box is out of bounds, zero fill)
                    tma_load_4d_pair(slab_base + stage * sp.slab_bytes, &p.mapA[st.map], &s_full[stage], st.c_elem, st.dw,
                                     h0 - 1, n);
                    if (++stage == kSlabStages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 2 + kEpiWarps) {
        // ------------------------------------------------------------ weight-tile producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int unit = unit0; unit < sp.num_units; unit += unit_step) {
                const int n_tile = unit / sp.num_pair_units;
                const int row0 = n_tile * kSlabN + static_cast<int>(rank) * (kSlabN / 2);
                for (int i = 0; i < sp.nstep; ++i) {
                    const SlabStep st = sp.steps[i];
                    for (int t = 0; t < st.ntap; ++t) {
                        mbar_wait(&w_empty[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&w_full[stage], 2 * kSlabWBytes);
                        tma_load_4d_pair(w_base + stage * kSlabWBytes, &p.mapB, &w_full[stage],
                                         t == 0 ? st.k0 : (t == 1 ? st.k1 : st.k2), row0, 0, 0);
                        if (++stage == kSlabWStages) stage = 0, phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (the pair's leader)
        // whole warp in uniform control flow, one elected lane issues (conv_tc.cu / ptx.cuh: elect_one_sync)
        if (rank == 0) {
            const uint32_t idesc = umma_idesc(TF32 ? 2 : (p.f16 ? 0 : 1), 2 * kBlockM, kSlabN);
            int ss = 0, ws = 0, acc = 0;
            uint32_t sphase = 0, wphase = 0, acc_phase = 0;
            const uint32_t slab0 = smem_u32(slab_base), wt0 = smem_u32(w_base);
            for (int unit = unit0; unit < sp.num_units; unit += unit_step) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc * (kSlabG * kSlabN);
                for (int i = 0; i < sp.nstep; ++i) {
                    const int ntap = i < sp.nstep3 ? 3 : 1;
                    mbar_wait(&s_full[ss], sphase);
                    tc_fence_after_sync();
                    const uint32_t sa = slab0 + ss * sp.slab_bytes;
                    for (int t = 0; t < ntap; ++t) {
                        // a 3-tap slab carries dh = -1, 0, +1 at slab rows +0, +1, +2; a 1-tap slab (1x1 shortcut) the centre
                        const int dhi = ntap == 3 ? t : 1;
                        mbar_wait(&w_full[ws], wphase);
                        tc_fence_after_sync();
                        const uint64_t bdesc = umma_desc_sw128(wt0 + ws * kSlabWBytes);
                        const uint32_t first = (i | t) == 0 ? 0u : 1u;
                        if (elect_one_sync()) {
#pragma unroll
                            for (int g = 0; g < kSlabG; ++g) {
                                const uint64_t adesc = umma_desc_sw128(sa + (g * p.BH + dhi) * sp.row_bytes);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t accum = k == 0 ? first : 1u;
                                    if (TF32)
                                        umma_tf32_pair(d_tmem + g * kSlabN, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                    else
                                        umma_bf16_pair(d_tmem + g * kSlabN, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                }
                            }
                            umma_commit_pair(&w_empty[ws]);
                        }
                        __syncwarp();
                        if (++ws == kSlabWStages) ws = 0, wphase ^= 1;
                    }
                    if (elect_one_sync()) umma_commit_pair(&s_empty[ss]);
                    __syncwarp();
                    if (++ss == kSlabStages) ss = 0, sphase ^= 1;
                }
                if (elect_one_sync()) umma_commit_pair(&tfull[acc]);
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9), as in conv_tc.cu
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        float* stg = stg_base + (warp - 2) * 1024;
        uint8_t* ostg = reinterpret_cast<uint8_t*>(stg);     // tma_epi (conv_common.cuh): 16-bit output block ...
        uint8_t* rstg = ostg + kEpiTmaBlockBytes;            // ... and 16-bit residual block of this warp
        uint64_t* my_rbar = &rbar[warp - 2];
        uint32_t rphase = 0;
        constexpr bool tma_epi = TEPI == 1;
        const int brick = p.BW * p.BH;  // = 128: one M tile is BH whole image rows
        const int row0 = quad * 32;
        const int bh0 = row0 / p.BW;
        const int bw0 = row0 - bh0 * p.BW;
        const int sub_r4 = lane >> 3, sub_c4 = lane & 7;
        const int sub_r8 = lane >> 2, sub_c8 = lane & 3;
        (void)brick;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = unit0; unit < sp.num_units; unit += unit_step) {
            int n_tile;
            const int s = super_of(unit, n_tile);
            const int n = s / sp.sb_per_img;
            const int hb = s - n * sp.sb_per_img;
            const bool valid = n < p.B;
            const uint32_t vmask = valid ? 0xffffffffu : 0u;
            // Residual rows are fetched one column chunk ahead, across the M tiles of the unit: the first chunk's loads are
            // issued before the accumulator wait, every later one while the previous chunk is being stored (an exposed
            // first-chunk fetch is a DRAM round trip per tile: the top epilogue stall in profiles/r02k_ncu_epi_notes.md).
            auto pix0_of = [&](int g) -> size_t {
                return (static_cast<size_t>(n) * p.Ho + (hb * kSlabG + g) * p.BH + bh0) * p.Wo + bw0;
            };
            float4 rpre[8];
            auto prefetch_resid = [&](size_t pix_base, int c_next) {
                if (p.resid16) {  // (16-bit residual stream: four 16-byte loads of 8 channels, as in conv_tc.cu)
                    const __nv_bfloat16* rp =
                        reinterpret_cast<const __nv_bfloat16*>(p.resid) + pix_base * p.ld_resid + n_tile * kSlabN + c_next;
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + sub_r8;
                        uint4 t = make_uint4(0u, 0u, 0u, 0u);
                        if (valid) t = __ldg(reinterpret_cast<const uint4*>(rp + static_cast<size_t>(r) * p.ld_resid) + sub_c8);
                        rpre[it] = make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
                    }
                    return;
                }
                const float* rp = p.resid + pix_base * p.ld_resid + n_tile * kSlabN + c_next;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + sub_r4;
                    rpre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid)
                        rpre[it] = __ldg(reinterpret_cast<const float4*>(rp + static_cast<size_t>(r) * p.ld_resid) + sub_c4);
                }
            };
            const int ho_u = hb * kSlabG * p.BH + bh0;  // first image row of this warp's pixels in M tile 0 of the unit
            if (p.resid && TEPI == 0) prefetch_resid(pix0_of(0), 32 * half);
            // TEPI 2: the lane's own 64 residual bytes of the next chunk, two 256-bit loads
            uint32_t rq[2][8];
            auto fetch_resid256 = [&](size_t pix_base, int c_next) {
                if (valid) {
                    const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                (pix_base + lane) * p.ld_resid + n_tile * kSlabN + c_next;
                    ldg256(rrow, rq[0]);
                    ldg256(rrow + 16, rq[1]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rq[0][i] = rq[1][i] = 0u;
                }
            };
            if (TEPI == 2 && p.resid) fetch_resid256(pix0_of(0), 32 * half);
            if (p.resid && tma_epi && lane == 0) {
                mbar_expect_tx(my_rbar, kEpiTmaBlockBytes);
                tma_load_4d(rstg, &p.mapRes, my_rbar, n_tile * kSlabN + 32 * half, bw0, ho_u, n);
            }
            // ... and the residual block of the NEXT unit is pulled into L2 now (its register fetches then cost an L2 hit,
            // not a DRAM round trip per column chunk: profiles/r02r_ncu_summary.md, K 1152 with / without residual)
            if (p.resid && unit + unit_step < sp.num_units) {
                int nt2;
                const int s2 = super_of(unit + unit_step, nt2);
                const int n2 = s2 / sp.sb_per_img, hb2 = s2 - n2 * sp.sb_per_img;
                if (n2 < p.B) {
                    const int esz = p.resid16 ? 2 : 4;
                    const int lines = (kSlabN * esz) >> 7;  // 128-byte lines per row of this unit's column range
                    const char* base = reinterpret_cast<const char*>(p.resid) + static_cast<size_t>(nt2) * kSlabN * esz;
                    for (int g = 0; g < kSlabG; ++g) {
                        const size_t px = (static_cast<size_t>(n2) * p.Ho + (hb2 * kSlabG + g) * p.BH + bh0) * p.Wo + bw0;
                        for (int i = lane + 32 * half; i < 32 * lines; i += 32 * (kEpiWarps / 4))
                            prefetch_l2(base + (px + i / lines) * static_cast<size_t>(p.ld_resid) * esz + ((i % lines) << 7));
                    }
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after_sync();
#pragma unroll 1
            for (int g = 0; g < kSlabG; ++g) {
                const size_t pix0 = pix0_of(g);
                const size_t stat_blk = pix0 >> 5;
                const uint32_t taddr =
                    tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (acc * kSlabG + g) * kSlabN;
                if constexpr (TEPI != 0) {
                    const int ho_g = ho_u + g * p.BH;
#pragma unroll 1
                    for (int c = 32 * half; c < kSlabN; c += 32 * (kEpiWarps / 4)) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + c, v);
                        const int col0 = n_tile * kSlabN + c;
                        uint4 rr[4];
                        if (p.resid) {
                            if constexpr (TEPI == 1) {
                                mbar_wait(my_rbar, rphase);
                                rphase ^= 1;
#pragma unroll
                                for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(rstg + epi_swz64(lane, j));
                                // (the next block's tensor load is issued below, after these values have been consumed)
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    rr[j] = make_uint4(rq[j >> 1][4 * (j & 1)], rq[j >> 1][4 * (j & 1) + 1],
                                                       rq[j >> 1][4 * (j & 1) + 2], rq[j >> 1][4 * (j & 1) + 3]);
                                if (c + 32 * (kEpiWarps / 4) < kSlabN)
                                    fetch_resid256(pix0, c + 32 * (kEpiWarps / 4));
                                else if (g + 1 < kSlabG)
                                    fetch_resid256(pix0_of(g + 1), 32 * half);
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
                                // the residual values have been consumed (their LDS have returned): the block goes back to the
                                // TMA unit (conv_tc.cu) - next: the next column chunk, or the first one of the unit's next M tile
                                fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    if (c + 32 * (kEpiWarps / 4) < kSlabN) {
                                        mbar_expect_tx(my_rbar, kEpiTmaBlockBytes);
                                        tma_load_4d(rstg, &p.mapRes, my_rbar, col0 + 32 * (kEpiWarps / 4), bw0, ho_g, n);
                                    } else if (g + 1 < kSlabG) {
                                        mbar_expect_tx(my_rbar, kEpiTmaBlockBytes);
                                        tma_load_4d(rstg, &p.mapRes, my_rbar, n_tile * kSlabN + 32 * half, bw0, ho_g + p.BH, n);
                                    }
                                }
                            }
                        }
                        if (p.out_scale != 1.0f) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) f[i] *= p.out_scale;
                        }
                        if (p.stats && valid) gn_partials(f, lane, p.stats + (stat_blk * p.stats_nblk + (col0 >> 2)) * 2);
                        if constexpr (TEPI == 1) {
                            if (lane == 0) bulk_wait_group_read0();  // the previous tensor store has read the staging block
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(ostg + epi_swz64(lane, j)) = make_uint4(
                                    pack_op16x2(f[8 * j], f[8 * j + 1], p.f16), pack_op16x2(f[8 * j + 2], f[8 * j + 3], p.f16),
                                    pack_op16x2(f[8 * j + 4], f[8 * j + 5], p.f16), pack_op16x2(f[8 * j + 6], f[8 * j + 7], p.f16));
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_4d(&p.mapOut, ostg, col0, bw0, ho_g, n);
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
                    continue;
                }
#pragma unroll 1
                for (int c = 32 * half; c < kSlabN; c += 32 * (kEpiWarps / 4)) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c, v);
                    const int col0 = n_tile * kSlabN + c;
                    if (p.resid) {
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
                        if (c + 32 * (kEpiWarps / 4) < kSlabN)
                            prefetch_resid(pix0, c + 32 * (kEpiWarps / 4));
                        else if (g + 1 < kSlabG)
                            prefetch_resid(pix0_of(g + 1), 32 * half);
                        __syncwarp();
                    }
                    float f[32];
                    tmem_ld_wait();
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
                    if (p.stats && valid) gn_partials(f, lane, p.stats + (stat_blk * p.stats_nblk + (col0 >> 2)) * 2);
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
                                reinterpret_cast<float4*>(p.out_f32 + (pix0 + r) * p.ld_out_f32 + col0)[sub_c4] =
                                    *reinterpret_cast<const float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2));
                        }
                    }
                    if (p.out_op) {
                        if (TF32) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int r = it * 4 + sub_r4;
                                if ((vmask >> r) & 1) {
                                    const float4 t =
                                        *reinterpret_cast<const float4*>(stg + r * 32 + ((sub_c4 ^ (r & 7)) << 2));
                                    reinterpret_cast<float4*>(static_cast<float*>(p.out_op) + (pix0 + r) * p.ld_out_op +
                                                              col0)[sub_c4] =
                                        make_float4(round_tf32(t.x), round_tf32(t.y), round_tf32(t.z), round_tf32(t.w));
                                }
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < 4; ++it) {
                                const int r = it * 8 + sub_r8;
                                if ((vmask >> r) & 1) {
                                    const float4 a =
                                        *reinterpret_cast<const float4*>(stg + r * 32 + (((2 * sub_c8) ^ (r & 7)) << 2));
                                    const float4 b = *reinterpret_cast<const float4*>(stg + r * 32 +
                                                                                      (((2 * sub_c8 + 1) ^ (r & 7)) << 2));
                                    reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_op) +
                                                             (pix0 + r) * p.ld_out_op + col0)[sub_c8] =
                                        make_uint4(pack_op16x2(a.x, a.y, p.f16), pack_op16x2(a.z, a.w, p.f16),
                                                   pack_op16x2(b.x, b.y, p.f16), pack_op16x2(b.z, b.w, p.f16));
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote_relaxed(&tempty[acc], 0);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (TEPI == 1 && lane == 0) bulk_wait_group0();  // this warp's tensor stores are complete before the CTA retires
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc_pair<kSlabTmemCols>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------- host
// Device copy of a launch's step table, cached by content (a network has a handful of distinct tables; the launches of a
// plan re-use them, so nothing is allocated or copied while a CUDA graph is being captured after the first eager step).
struct StepTable {
    SlabStep steps[kSlabMaxSteps];
    int n, device;
    SlabStep* dev;
};
static StepTable g_tables[64];
static int g_ntables = 0;

static const SlabStep* step_table(const SlabStep* steps, int n, int device, cudaStream_t stream, int* rc) {
    *rc = NLC_OK;
    for (int i = 0; i < g_ntables; ++i)
        if (g_tables[i].n == n && g_tables[i].device == device && memcmp(g_tables[i].steps, steps, sizeof(SlabStep) * n) == 0) return g_tables[i].dev;
    if (g_ntables == 64) {
        *rc = set_error(NLC_ENOTSUP, "conv_slab: more than 64 distinct step tables");
        return nullptr;
    }
    StepTable& t = g_tables[g_ntables];
    memset(&t, 0, sizeof(t));
    memcpy(t.steps, steps, sizeof(SlabStep) * n);
    t.n = n, t.device = device;
    if (cudaMalloc(&t.dev, sizeof(SlabStep) * n) != cudaSuccess ||
        cudaMemcpyAsync(t.dev, t.steps, sizeof(SlabStep) * n, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
        *rc = set_error(NLC_ECUDA, "conv_slab: step table upload failed");
        return nullptr;
    }
    ++g_ntables;
    return t.dev;
}

bool conv_slab_eligible(const nlc_ctx* ctx, const nlc_conv_desc* d, int chunk) {
    if (!ctx->use_slab || d->dtype == NLC_F32X3 || d->stride != 1 || d->nseg < 9 || d->wbatched.ptr) return false;
    if (d->out_up || d->out_head_split || d->resid_mode != 0 || d->act) return false;
    if (d->Cout % kSlabN != 0 || (ctx->use_slab == 1 && d->Cout != kSlabN)) return false;
    const int W = d->Wo, H = d->Ho;
    if (W != 16 && W != 32 && W != 64) return false;
    const int BH = kBlockM / W;
    if (H % (kSlabG * BH) != 0) return false;
    const nlc_kseg& s0 = d->seg[0];
    if (s0.c0 != 0 || s0.nch % chunk != 0) return false;
    for (int i = 0; i < 9; ++i) {  // the nine taps of a padding-1 3x3 over one source, row-major
        const nlc_kseg& s = d->seg[i];
        if (s.src != s0.src || s.c0 != 0 || s.nch != s0.nch || s.dh != i / 3 - 1 || s.dw != i % 3 - 1) return false;
    }
    if (d->src[s0.src].H != H || d->src[s0.src].W != W) return false;
    for (int i = 9; i < d->nseg; ++i) {  // optional 1x1 segments (fused shortcut over another tensor)
        const nlc_kseg& s = d->seg[i];
        if (s.dh != 0 || s.dw != 0 || s.nch % chunk != 0 || d->src[s.src].H != H || d->src[s.src].W != W) return false;
    }
    int nstep = 3 * (s0.nch / chunk);
    for (int i = 9; i < d->nseg; ++i) nstep += d->seg[i].nch / chunk;
    return nstep <= kSlabMaxSteps;
}

// `p` arrives with the epilogue fields, mapB (box = 64 weight rows) and the tile geometry filled in by nlc_conv_tc
int launch_conv_slab(nlc_ctx* ctx, const nlc_conv_desc* d, ConvKParams& p, int chunk, bool tf32, cudaStream_t stream) {
    const int esz = tf32 ? 4 : 2;
    SlabParams sp;
    memset(&sp, 0, sizeof(sp));
    const int W = d->Wo, H = d->Ho;
    p.BW = W, p.BH = kBlockM / W, p.BN = 1;
    p.tiles_w = 1, p.tiles_h = H / p.BH, p.tiles_n = d->B;
    const int rows = kSlabG * p.BH + 2;
    sp.row_bytes = W * kChunkBytes;
    sp.slab_bytes = rows * sp.row_bytes;
    sp.sb_per_img = H / (kSlabG * p.BH);
    sp.num_super = d->B * sp.sb_per_img;
    sp.num_pair_units = (sp.num_super + 1) / 2;
    sp.num_units = sp.num_pair_units * (d->Cout / kSlabN);
    const size_t smem = static_cast<size_t>(kSlabStages) * sp.slab_bytes + kSlabWStages * kSlabWBytes + kSlabBarBytes +
                        kSlabEpiBytes + 1024;
    NLC_REQUIRE(smem <= 232448, "conv_slab: %zu bytes of shared memory", smem);

    const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    for (int s = 0; s < d->nsrc; ++s) {
        const nlc_operand& o = d->src[s];
        cuuint64_t gdim[4] = {(cuuint64_t)o.C, (cuuint64_t)o.W, (cuuint64_t)o.H, (cuuint64_t)o.B};
        const cuuint64_t sh = o.sh ? (cuuint64_t)o.sh : (cuuint64_t)o.W * o.ld;
        const cuuint64_t sn = o.sn ? (cuuint64_t)o.sn : (cuuint64_t)o.H * o.W * o.ld;
        cuuint64_t gstr[3] = {(cuuint64_t)o.ld * esz, sh * esz, sn * esz};
        cuuint32_t box[4] = {(cuuint32_t)chunk, (cuuint32_t)W, (cuuint32_t)rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ctx->encode_tiled(&p.mapA[s], dt, 4, const_cast<void*>(o.ptr), gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        NLC_REQUIRE(r == CUDA_SUCCESS, "conv_slab: cuTensorMapEncodeTiled(A%d) failed with %d", s, (int)r);
    }
    for (int s = d->nsrc; s < NLC_MAX_SRC; ++s) p.mapA[s] = p.mapA[0];

    // step table: channel chunk outermost, then the horizontal tap; each slab carries its three vertical taps.  Weight K
    // coordinate of tap (dh, dw), chunk c: (3 (dh+1) + (dw+1)) * Cin + c (the packing is [Cout][tap][Cin])
    SlabStep steps[kSlabMaxSteps];
    memset(steps, 0, sizeof(steps));
    int n = 0;
    const int cin = d->seg[0].nch;
    for (int c = 0; c < cin; c += chunk)
        for (int dwi = 0; dwi < 3; ++dwi) {
            SlabStep& st = steps[n++];
            st.map = d->seg[0].src, st.dw = dwi - 1, st.c_elem = c, st.ntap = 3;
            st.k0 = (0 + dwi) * cin + c, st.k1 = (3 + dwi) * cin + c, st.k2 = (6 + dwi) * cin + c;
        }
    int kbase = 9 * cin;
    for (int i = 9; i < d->nseg; ++i) {
        const nlc_kseg& sg = d->seg[i];
        for (int c = 0; c < sg.nch; c += chunk) {
            SlabStep& st = steps[n++];
            st.map = sg.src, st.dw = 0, st.c_elem = sg.c0 + c, st.ntap = 1;
            st.k0 = kbase + c, st.k1 = st.k2 = 0;
        }
        kbase += sg.nch;
    }
    int rc;
    sp.steps = step_table(steps, n, ctx->device, stream, &rc);
    if (rc != NLC_OK) return rc;
    sp.nstep = n;
    sp.nstep3 = 3 * (cin / chunk);
    rc = epi_tma_setup(ctx, d, p);  // (the geometry above is final: BW = W, BH = 128 / W, BN = 1)
    if (rc != NLC_OK) return rc;
    sp.k = p;

    static PerDeviceFlag configured[4];
    const int which = tf32 ? 1 : (p.tma_epi ? 1 + p.tma_epi : 0);
    auto kern = tf32 ? conv_slab_kernel<1>
                     : (p.tma_epi == 1 ? conv_slab_kernel<0, 1> : (p.tma_epi == 2 ? conv_slab_kernel<0, 2> : conv_slab_kernel<0>));
    if (!configured[which][ctx->device]) {
        NLC_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured[which][ctx->device] = true;
    }
    const int pairs = sp.num_units < ctx->sm_count / 2 ? sp.num_units : ctx->sm_count / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs), cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem, cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 2 : 1;
    NLC_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, sp));
    NLC_CHECK_LAUNCH();
    return NLC_OK;
}

}  // namespace nlc
