/*
 * nlc_b200 — C ABI of the B200-native NLC sampling hot path.
 *
 * Every entry point below replaces a piece of PyTorch-eager work that the reference
 * (Walleclipse/Diffusion-NLC, pure Python) performs inside its per-timestep loop; the
 * reference has no FFI of its own, so each declaration cites the Python seam it stands
 * behind (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless a name ends in _host
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*), never synchronises,
 *     never allocates caller-visible memory
 *   - returns 0 on success, a negative NLC_E* code otherwise; nlc_last_error() gives the text
 *   - activations are NHWC ("pixel rows"): element (n,h,w,c) lives at ptr[((n*H+h)*W+w)*ld + c],
 *     ld >= C lets a tensor be a channel slice of a wider (concatenated) buffer
 *   - one nlc_ctx per (process, device); not thread-safe (the reference is one thread per GPU)
 */
#ifndef NLC_B200_H
#define NLC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLC_OK 0
#define NLC_EINVAL (-1)   /* bad argument / unsupported shape */
#define NLC_ECUDA (-2)    /* CUDA runtime / driver error      */
#define NLC_ENOTSUP (-3)  /* valid but not implemented        */

/* Operand modes of the tensor-core path.  Every producer writes the "operand copy" of an activation in this type. */
#define NLC_F32 0   /* fp32 containers holding tf32-rounded values; one kind::tf32 MMA per K step              */
#define NLC_BF16 1  /* bf16; kind::f16 MMA (the throughput mode)                                              */
#define NLC_F32X3 2 /* plain fp32 operands and weights; nlc_conv_tc splits them into tf32 hi+lo parts in shared */
                    /* memory and issues three MMAs per K step: fp32-accurate products (the accuracy mode)     */
#define NLC_F16 3   /* fp16; kind::f16 MMA at the bf16 rate with 3 more mantissa bits (the reference's own      */
                    /* reduced-precision mode is fp16 too: src/fp16_util.py:15-22, UNetModel.convert_to_fp16)   */

typedef struct nlc_ctx nlc_ctx;

const char* nlc_last_error(void);
int nlc_abi_version(void);
nlc_ctx* nlc_create(int device);
void nlc_destroy(nlc_ctx* ctx);
int nlc_sm_count(nlc_ctx* ctx);
/* Kernel-selection switches of a context (defaults from the environment: NLC_CTA_PAIRS, NLC_SLAB, NLC_TMA_EPI,
 * NLC_ATTN_ONEPASS, NLC_SPLITK):
 *   "cta_pairs" 0|1  tcgen05 cta_group::2 convolution kernels;
 *   "slab" 0|1|2     halo-slab 3x3 kernel: off / layers with 128 output channels / every eligible layer;
 *   "tma_epi" 0|1|2  16-bit convolution epilogues: staged through the load/store unit / through TMA (residual block by
 *                    tensor load, output by tensor store) / 256-bit global accesses from registers;
 *   "attn_onepass" 0|1  fused attention: two passes over the keys / one pass with an online softmax;
 *   "splitk" 0|1     deterministic split-K (workspace + reducing epilogue kernel) of the small-M convolution launches. */
int nlc_ctx_set(nlc_ctx* ctx, const char* key, int value);

/* ------------------------------------------------------------------------------------------------
 * N1/N2/N3 building blocks — replace torch.nn.Conv2d / GroupNorm / attention launches inside
 * UNetModel.forward / encode (src/unet_ddim.py:323-393, src/unet_adm.py:636-693,
 * src/edm_networks.py:835-909) and SigmaModel.forward (src/unet_ddim.py:521-529).
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    const void* ptr; /* operand tensor: bf16 / fp16 (NLC_BF16 / NLC_F16) or fp32 (tf32-rounded, or plain: NLC_F32X3) */
    int B, H, W, C;  /* logical NHWC extent                                                        */
    int ld;          /* elements between consecutive pixels                                        */
    int64_t sh, sn;  /* element strides of H and B; 0 = dense (W*ld, H*W*ld). Non-dense strides let H   */
                     /* index attention heads inside a [B,T,3C] qkv tensor.                         */
} nlc_operand;

typedef struct {
    int src;    /* index into src[]                                              */
    int dh, dw; /* input pixel = stride * output pixel + (dh, dw); pad by OOB=0  */
    int c0;     /* first channel of the source covered by this segment           */
    int nch;    /* channels covered (multiple of the 128-byte K chunk)           */
} nlc_kseg;

#define NLC_MAX_SRC 3
#define NLC_MAX_SEG 24

/* Implicit-GEMM convolution on tcgen05 tensor cores:
 *   out[n,ho,wo,:] = scale * ( sum_seg W_seg . src[seg][n, s*ho+dh, s*wo+dw, c0:c0+nch]
 *                              + bias + rowvec[n,:] + resid[n,ho,wo,:] )
 * A 3x3 conv is nine segments, a 1x1 conv one, a ResNet block's 1x1 shortcut is one extra segment over a
 * second source (torch.nn.Conv2d calls at src/unet_ddim.py:109-135,141,148-156). `weight` is
 * [Cout][sum nch] in the operand dtype, K ordered like seg[]. */
typedef struct {
    int dtype; /* NLC_BF16 / NLC_F16 (kind::f16), NLC_F32 (kind::tf32) or NLC_F32X3 (3 x kind::tf32 on split fp32) */
    int nsrc;
    nlc_operand src[NLC_MAX_SRC];
    int nseg;
    nlc_kseg seg[NLC_MAX_SEG];
    const void* weight;
    /* Batched right-hand operand (attention: S = Q K^T, O = P V): when wbatched.ptr != NULL it replaces
     * `weight`; it is [B][H][Cout rows][K] with the strides given (C = K, W = Cout), and the tile at
     * (image n, row h) multiplies with its own slice.  Requires Wo >= 128. */
    nlc_operand wbatched;
    int Cout;
    int stride;
    int B, Ho, Wo;
    const float* bias;   /* [Cout] or NULL                          */
    const float* rowvec; /* [B, ld_rowvec] per-sample add or NULL   */
    int ld_rowvec;
    const void* resid;  /* NHWC fp32 [B,Ho,Wo,ld_resid] (operand dtype with resid_is_op) or NULL */
    int ld_resid;
    float out_scale;
    float* out_f32; /* NHWC fp32 or NULL                       */
    int ld_out_f32;
    void* out_op; /* NHWC operand-dtype copy (bf16 / tf32-rounded fp32) or NULL */
    int ld_out_op;
    int out_head_split; /* >0: output row (n,ho,wo) is written at pixel (n,wo), channel offset ho*out_head_split
                           (merges attention heads back into [B,T,C]); 0: dense NHWC                  */
    float* stats;       /* optional: GroupNorm partials of the fp32 result, written by the epilogue so that the next
                           GroupNorm (nlc_groupnorm with `stats`) skips its statistics pass.  Layout
                           [B*Ho*Wo/32][stats_nblk][2] = (mean, M2) over 32 consecutive pixels x 4 channels; the
                           pointer is pre-offset to this output's first 4-channel block.  Needs Ho*Wo >= 128. */
    int stats_nblk;     /* 4-channel blocks per pixel group in the stats buffer (= its total channels / 4)    */
    int resid_mode;     /* 0: resid is [B,Ho,Wo,.]; 1: resid is [B,Ho/2,Wo/2,.] and is added nearest-upsampled x2;
                           2: resid is [B,2Ho,2Wo,.] and its 2x2 average is added (sum of the window, then * 0.25).
                           Folds ADM's x_upd (src/unet_adm.py:236-243, Upsample / Downsample without conv of the
                           skip path, :81-140) into the conv that consumes it.                                   */
    int out_up;         /* 0: dense output; 1 + 2a + b (a, b in {0,1}): output pixel (n,ho,wo) is written at
                           (n, 2ho+a, 2wo+b) of a [B,2Ho,2Wo,.] tensor (and its GroupNorm partials into that tensor's
                           block range).  With the four phase-combined 2x2 tap sets of a 3x3 kernel this computes
                           "nearest-neighbour x2 upsample, then 3x3 conv" (src/unet_ddim.py:58-74) at the LOW
                           resolution: 16 instead of 36 multiplies per output and no replicated operand in HBM. */
    int resid_is_op;    /* 1: `resid` points at a tensor in the 16-bit OPERAND dtype (NLC_BF16 / NLC_F16 modes only; ld_resid
                           in elements, a multiple of 8) instead of fp32: the residual stream of the 16-bit-activation
                           plans, read 2 instead of 4 bytes per element.  All three resid_modes. */
    int act;            /* activation applied last (after bias / row / residual / out_scale): 0 none, 1 ReLU (the
                           BasicConv2d blocks of the FID InceptionV3, whose BatchNorm is folded into weight and bias) */
} nlc_conv_desc;

int nlc_conv_tc(nlc_ctx* ctx, const nlc_conv_desc* d, void* stream);

/* Direct (CUDA-core, fp32) convolution for the layers whose channel counts cannot fill a tensor-core tile:
 * conv_in (3 -> C) reading the sampler's NCHW fp32 image with the per-sample input scale 1/sqrt(sigma^2+1)
 * folded in (src/experiments.py:273-282,295-302), and conv_out (C -> 3|6) writing NCHW fp32.
 * weight is the torch layout [Cout][Cin][3][3] fp32. */
int nlc_conv_in_nchw(nlc_ctx* ctx, const float* x_nchw, const float* in_scale /*[B] or NULL*/, int B, int Cin, int H,
                     int W, const float* weight, const float* bias, int Cout, float* out_f32, int ld_out_f32,
                     void* out_op, int ld_out_op, int op_dtype, void* stream);
/* The input convolution on the tensor cores: patches[(n,h,w), tap*Cin + ci] = in_scale[n] * x[n, ci, h+kh-1, w+kw-1]
 * (zero outside the image), one 128-byte K row per pixel (64 bf16 / 32 tf32-rounded fp32 elements, zero tail), to be
 * multiplied by nlc_conv_tc as a 1x1 convolution with the 3x3 weights packed [Cout][tap*Cin + ci] (same conv as
 * nlc_conv_in_nchw; src/unet_ddim.py:301, src/unet_adm.py:480, src/edm_networks.py:786).  Cin <= 3. */
int nlc_im2col_in(nlc_ctx* ctx, const float* x_nchw, const float* in_scale /*[B] or NULL*/, int B, int Cin, int H,
                  int W, void* patches_op, int op_dtype, void* stream);
int nlc_conv_out_nchw(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int Cin, int H, int W,
                      const float* weight, const float* bias, int Cout, float* out_nchw, void* stream);
/* The output convolution on the tensor cores (16-bit modes): nlc_conv_tc with the C -> 3|6 weights zero-padded to one
 * 64-channel tile writes NHWC fp32 [B,H,W,ld]; this copies its first C (<= 8) channels into the sampler's NCHW eps
 * tensor (src/unet_ddim.py:362, src/unet_adm.py:664, src/edm_networks.py:876). */
int nlc_nhwc_head_to_nchw(nlc_ctx* ctx, const float* x, int ld, int B, int H, int W, int C, float* out_nchw,
                          void* stream);

/* GroupNorm (+ optional SiLU, + optional per-sample scale/shift) producing the next conv's operand.
 * Replaces Normalize/GroupNorm32 + nonlinearity (src/unet_ddim.py:54-55,139-146; src/nn_util.py:17-19;
 * src/unet_adm.py:236-252).  Statistics are fp32 Welford partials per (sample, pixel chunk, group), merged by
 * the apply pass.  x is NHWC fp32 with C channels in `groups` groups.
 *   y = ((x-mean)*rstd*gamma+beta) * (1+scale[n,c]) + shift[n,c]  -> SiLU (if silu) -> operand dtype
 * `stats` != NULL: the statistics pass is skipped; the per-(32 pixel, 4 channel) partials the producing
 * nlc_conv_tc wrote are merged instead (x is then read exactly once).  `resample` applies ADM's resblock_updown
 * h_upd (src/unet_adm.py:236-243) / the EDM block's conv0 resampling (src/edm_networks.py:85-93) to the activated
 * tensor while it is written: y is [B,2H,2W,C] (1) or [B,H/2,W/2,C] (2).
 * `x_is_op` != 0: x is a 16-bit tensor in the operand dtype instead of fp32 (the tensor between a ResBlock's two
 * convolutions, which the 16-bit modes keep in the operand dtype only; needs `stats`, whose partials the producing conv took
 * from its fp32 accumulators, and resample == 0): 4 instead of 6 bytes moved per element. */
int nlc_groupnorm(nlc_ctx* ctx, const void* x, int x_is_op, int ld_x, int B, int H, int W, int C, int groups, float eps,
                  const float* gamma, const float* beta, const float* scale, const float* shift, int ld_ss,
                  int silu, const float* stats /* nullable: partials written by nlc_conv_tc */, int stats_nblk,
                  int resample /* 0 none, 1 nearest x2, 2 avgpool 2x2 of the activated tensor */, void* y_op,
                  int ld_y, int op_dtype, float* workspace /* >= nlc_groupnorm_ws floats */, void* stream);
size_t nlc_groupnorm_ws(int B, int HW, int C, int groups);

/* fp32 NHWC -> operand dtype, optionally nearest-neighbour x2 upsampled (Upsample.forward,
 * src/unet_ddim.py:69-74) or 2x2 average pooled (src/unet_adm.py:134-140). mode: 0 copy, 1 up2, 2 avgpool2 */
int nlc_resample(nlc_ctx* ctx, const float* x, int ld_x, int B, int H, int W, int C, int mode, float* y_f32,
                 int ld_y_f32, void* y_op, int ld_y_op, int op_dtype, void* stream);

/* ---- The sigma-model's own forward / backward of the training step (src/experiments.py:683-691 leaves
 * it to autograd over src/unet_ddim.py:439-529).  Activations are NHWC fp32 matrices [B*H*W, C]; every contraction is nlc_sgemm,
 * the functions below are what sits between the GEMMs (csrc/sigma_train.cu; host side: nlc_b200/training.py NativeSigmaModel). */
/* C[b] = A[b] B[b] (+ add[b]): strided batched fp32 GEMM on the CUDA cores.  A element (i,k) at A + b*sab + i*sai + k*sak, B
 * element (k,j) at Bm + b*sbb + k*sbk + j*sbj, C[b] row-major [M,N] contiguous; `add` (same layout as C) or NULL. */
int nlc_sgemm(nlc_ctx* ctx, int batch, int M, int N, int K, const float* A, long long sab, long long sai, long long sak,
              const float* Bm, long long sbb, long long sbk, long long sbj, float* Cm, const float* add, void* stream);
/* x [B,H,W,C] -> patches [B*Ho*Wo, C*9] with column c*9 + kh*3 + kw (torch's weight.view(Cout, Cin*9) multiplies them);
 * down 0: stride 1, zero padding 1; down 1: the reference's Downsample, F.pad(x,(0,1,0,1)) then stride 2 (src/unet_ddim.py:89-94);
 * down 2: stride 2, zero padding 1 (guided-diffusion Downsample of the ADM sigma-model, src/unet_adm.py:143-166).
 * nlc_fold3x3 is the adjoint: dx = beta*dx + sum of the patch gradients that read each pixel. */
int nlc_unfold3x3(nlc_ctx* ctx, const float* x, int B, int H, int W, int C, int down, float* patches, void* stream);
int nlc_fold3x3(nlc_ctx* ctx, const float* d_patches, int B, int H, int W, int C, int down, float* dx, float beta, void* stream);
/* GroupNorm(groups, eps) [+ swish when act = 1] in training: forward saves stats [B*groups][2] = (mean, rstd); backward writes
 * (or, accumulate != 0, adds to) dx and ATOMICALLY accumulates dgamma / dbeta (zero them first). */
int nlc_gn_train_fwd(nlc_ctx* ctx, const float* x, int B, int HW, int C, int groups, float eps, const float* gamma,
                     const float* beta, int act, float* y, float* stats, void* stream);
int nlc_gn_train_bwd(nlc_ctx* ctx, const float* x, const float* dy, int B, int HW, int C, int groups, const float* gamma,
                     const float* beta, int act, const float* stats, float* dx, int accumulate, float* dgamma, float* dbeta,
                     void* stream);
/* dp == NULL: out = softmax(scale * s) over rows of length T; dp != NULL: s holds the probabilities p and
 * out = scale * p * (dp - sum_j dp_j p_j), the gradient with respect to the unscaled logits. */
int nlc_softmax_rows(nlc_ctx* ctx, const float* s, const float* dp, int rows, int T, float scale, float* out, void* stream);
int nlc_bias_add(nlc_ctx* ctx, float* y, const float* bias, long long rows, int C, void* stream);      /* y[r,c] += bias[c] */
int nlc_colsum(nlc_ctx* ctx, const float* x, long long rows, int C, float* out, void* stream);         /* out[c] = sum_r x[r,c] */
int nlc_axpby(nlc_ctx* ctx, float a, const float* x, float b, const float* y /* or NULL */, float* out, long long n, void* stream);
int nlc_permute_nhwc(nlc_ctx* ctx, const float* x, int B, int HW, int C, int to_nchw, float* y, void* stream);
/* BatchNorm1d(F) in training mode followed by GELU(erf) on x [B,F]: dy == NULL forward (out = gelu(bn(x)), stats [F][2] =
 * (batch mean, rstd) saved, running statistics updated with `momentum` and the unbiased variance when run_mean != NULL);
 * dy != NULL backward (out = dx, dgamma[F], dbeta[F] written). */
int nlc_bn1d_gelu_train(nlc_ctx* ctx, const float* x, const float* dy, int B, int F, float eps, float momentum,
                        const float* gamma, const float* beta, float* run_mean, float* run_var, float* stats, float* out,
                        float* dgamma, float* dbeta, void* stream);
/* The same with the activation chosen: act 0 = GELU(erf), 1 = SiLU (the head of the EDM sigma-model, src/edm_networks.py:1006-1010). */
int nlc_bn1d_act_train(nlc_ctx* ctx, const float* x, const float* dy, int B, int F, float eps, float momentum, int act,
                       const float* gamma, const float* beta, float* run_mean, float* run_var, float* stats, float* out,
                       float* dgamma, float* dbeta, void* stream);
/* dist_hat = r + 1 (src/experiments.py:689); loss = MSELoss (kind 0) | L1Loss (kind 1), mean reduction; dr = d loss / d r */
int nlc_head_loss(nlc_ctx* ctx, const float* r, const float* target, int B, int kind, float* dist_hat, float* loss, float* dr,
                  void* stream);
/* The same with per-sample weights [B]: loss = sum_b w_b l_b / sum_b w_b (`loss_weighted` of the EDM loop, src/experiments.py:1019-1021). */
int nlc_head_loss_weighted(nlc_ctx* ctx, const float* r, const float* target, const float* weight, int B, int kind,
                           float* dist_hat, float* loss, float* dr, void* stream);

/* ---- FID statistics on the device.  The reference goes through the third-party pytorch_fid package
 * after writing every sample as a PNG (src/experiments.py:210-226 fid_helper -> compute_statistics_of_path /
 * calculate_frechet_distance; image_sample.py:566,703; result_evaluater.py:24-27).  The InceptionV3 convolutions are
 * nlc_conv_tc GEMMs over patch matrices (BatchNorm folded, ReLU = nlc_conv_desc.act); these are the pieces around them. */
/* x [B,3,H,W] fp32 -> y NHWC [B,R_out,R_out,ld_y] operand dtype (3 channels written).  from_pm1: x is a sample in [-1,1],
 * mapped add(1).div(2).clamp(0,1) (image_sample.py:560); quantize: the 8-bit PNG round trip (save_image: x*255+0.5, clamp,
 * truncate; ToTensor: /255); resize: bilinear to R_out x R_out, align_corners = False (pytorch_fid InceptionV3
 * resize_input); normalize: 2x - 1 (normalize_input). */
int nlc_fid_preprocess(nlc_ctx* ctx, const float* x_nchw, int B, int H, int W, int from_pm1, int quantize, int resize,
                       int normalize, int R_out, void* y_op, int ld_y, int op_dtype, void* stream);
/* NHWC [B,H,W,ld_x] (C channels used) -> patch matrix [M_pad, K_pad], row m = (n, ho, wo), column (kh*KW + kw)*C + c; rows
 * >= B*Ho*Wo and columns >= KH*KW*C are zero.  Ho = (H + 2 PH - KH) / SH + 1 (torch.nn.Conv2d arithmetic). */
int nlc_im2col_nhwc(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C, int KH, int KW,
                    int SH, int SW, int PH, int PW, void* out, int K_pad, long long M_pad, void* stream);
/* 3x3 pooling on NHWC operand tensors: mode 0 max_pool2d, 1 avg_pool2d(count_include_pad=False); stride 1|2, pad 0|1. */
int nlc_pool2d(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C, int stride, int pad,
               int mode, void* y_op, int ld_y, void* stream);
/* adaptive_avg_pool2d(1,1): [B, HW, ld_x] operand -> fp32 [B, C] */
int nlc_global_avgpool(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int HW, int C, float* y, void* stream);
/* FID statistics: sum[D] += sum_b f[b,:], outer[D,D] += f^T f, both fp64 (mu = sum/N, Sigma = (outer - N mu mu^T)/(N-1)) */
int nlc_cov_accumulate(nlc_ctx* ctx, const float* feats, int B, int D, double* sum, double* outer, void* stream);

/* Same resampling on a tensor that is already in the operand dtype (ADM's resblock_updown pools / upsamples the
 * activated tensor GN+SiLU(x) before the block's first conv, src/unet_adm.py:236-243). mode: 1 up2, 2 avgpool2 */
int nlc_resample_op(nlc_ctx* ctx, const void* x_op, int op_dtype, int ld_x, int B, int H, int W, int C, int mode,
                    void* y_op, int ld_y, void* stream);

/* Fused softmax attention over NHWC tokens: qkv is [B, T, ld] with q at column q_off + h*dh, etc.
 * Replaces bmm/softmax/bmm (src/unet_ddim.py:193-207), QKVAttention(Legacy) (src/unet_adm.py:328-389),
 * AttentionOp (src/edm_networks.py:124-130).  softmax(scale * q k^T) v, fp32 math, output operand dtype. */
int nlc_attention(nlc_ctx* ctx, const void* qkv, int op_dtype, int ld, int q_off, int k_off, int v_off,
                  int head_stride, int B, int T, int heads, int dh, float scale, void* out_op, int ld_out,
                  void* workspace /* nlc_attention_ws bytes, may be NULL when that is 0 */, void* stream);
size_t nlc_attention_ws(int op_dtype, int B, int T, int heads, int dh);

/* Small dense layers (timestep-embedding MLP, per-block temb projections, sigma head).
 * y[b, n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] );  act: 0 none, 1 SiLU, 2 GELU(erf).
 * src/unet_ddim.py:327-330,143; src/unet_adm.py:650,199-205. */
int nlc_linear(nlc_ctx* ctx, const float* x, int ld_x, int B, int K, const float* W, const float* bias, int N,
               int act_in, int act_out, float* y, int ld_y, void* stream);

/* Sinusoidal timestep embedding out[b, :] = [sin|cos](t_b * freqs) (cos first when cos_first != 0).
 * The frequency table is built on the host exactly as the reference does, so the three variants
 * (src/unet_ddim.py:28-46 sin||cos, src/nn_util.py:103-121 cos||sin, src/edm_networks.py:212-225 cos||sin)
 * share one kernel. */
int nlc_timestep_embedding(nlc_ctx* ctx, const float* t, int B, const float* freqs, int half, int cos_first,
                           float* out, int ld_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * D1 / S2 / S3 / S4 / S5 — the sampler arithmetic around the networks.
 * ---------------------------------------------------------------------------------------------- */

/* Per-sample L2 norm over d contiguous floats: vector_norm (src/utils.py:7-9). out[b] = ||x[b,:]||_2 */
int nlc_row_norm(nlc_ctx* ctx, const float* x, int B, int d, float* out, void* stream);

/* refine_prior_sigma + searchsorted (src/experiments.py:401-419, src/schedulers.py:185-190), norms from
 * nlc_row_norm:
 *   nrm = norms_b/sqrt(d); sigma_b = clamp(sigma_in_b, max(nrm-norm_max,0), nrm+norm_min)      (refine != 0)
 *   t_b = clamp(first i with table[i] >= sigma_b  (- time_shift if min_b t > 0), 0, 1000)
 *   refine == 0: sigma_b = sigma_in_b, t_b = clamp(t_fixed, 0, 1000)
 *   in_scale_b = sqrt(1/(sigma_b^2+1))  (convert_coordinate, src/experiments.py:273-282)
 * sigma_in has n_sigma_in entries (1 = broadcast scalar, B = per sample).
 * slopes == NULL: discrete time (searchsorted).  slopes != NULL (continuous_t): the n_table-1 interval slopes of
 * Interp1d(sigma table -> arange(n_table)) built by the host exactly as src/torchinterp1d.py:137-142 does, and
 * t = ind + slopes[ind]*(sigma - table[ind]), ind = clamp(searchsorted-1, 0, n_table-2) (src/schedulers.py:210-220). */
int nlc_refine_sigma(nlc_ctx* ctx, const float* norms, int B, int d, const float* sigma_in, int n_sigma_in,
                     float norm_min, float norm_max, int refine, float t_fixed, const float* sigma_table,
                     const float* slopes, int n_table, int time_shift, float* sigma_out, float* t_out,
                     float* in_scale_out, void* stream);

/* NLC correction (src/experiments.py:424-431): sigma_hat = sigma*(1+r); sigma_prev_hat = sigma_hat*sigma_prev/sigma
 * (style pred) or sigma_prev (pred_partial); t_hat = clamp(searchsorted(table, sigma_hat)); in_scale = 1/sqrt(s^2+1) */
int nlc_sigma_correct(nlc_ctx* ctx, const float* r, const float* sigma, const float* sigma_prev, int n_prev, int B,
                      int update_prev, const float* sigma_table, const float* slopes, int n_table, float* sigma_hat,
                      float* sigma_prev_hat, float* t_hat, float* in_scale_out, void* stream);

/* Dynamic thresholding, in place (src/experiments.py:190-204 with the driver's partial(..., 0.99, 100)):
 * s_b = clamp(quantile(|x_b|, ratio), 1, max_value), x_b <- clamp(x_b, -s_b, s_b) / s_b.  The quantile is
 * torch.quantile's (linear interpolation between the two exact order statistics).  s_out [B] may be NULL. */
int nlc_dynamic_threshold(nlc_ctx* ctx, float* x, int B, int d, double ratio, float max_value, float* s_out,
                          void* stream);

/* projection_loop's sigma feed-forward (image_sample.py:483-496), norms = nlc_row_norm(x_{t-1}):
 *   cur = norms/sqrt(d); dist = sqrt(cur^2 + norm_max^2 - 2*cur*norm_max*0.99 + 1e-8)
 *   sigma_out = r0*sigma_prev_orig + r1*sigma_prev + r2*sigma_t*(cur/last_norm) + r3*dist
 *   t_out = get_t_from_sigma(sigma_out) (unclamped, like the reference); last_norm <- cur (in place) */
int nlc_sigma_estimate(nlc_ctx* ctx, const float* norms, float* last_norm, int B, int d, float norm_max,
                       float sigma_prev_orig, const float* sigma_prev, int n_prev, const float* sigma_t, int n_t,
                       const float* rates4_host, const float* sigma_table, const float* slopes, int n_table,
                       float* sigma_out, float* t_out, void* stream);

/* eps normalisation (src/utils.py:11-16): eps_b <- sqrt(d) * eps_b / max(||eps_b||, 1e-12), in place. */
int nlc_normalize_rows(nlc_ctx* ctx, float* x, int B, int d, void* stream);

#define NLC_SCHED_DDIM 0
#define NLC_SCHED_DDIM_SIMPLE 1
#define NLC_SCHED_DDIM_SIMPLE_ORIG 2
#define NLC_SCHED_DDIM_SIMPLE_DRAG 3
#define NLC_SCHED_DDPM 4
#define NLC_SCHED_DDPM_ORIG 5
#define NLC_SCHED_DDIM_ORIG 6

#define NLC_CLIP_NONE 0
#define NLC_CLIP_CLAMP 1
#define NLC_CLIP_DYNAMIC 2 /* host-side: pred_xstart without clip, then nlc_dynamic_threshold */

/* x0_hat = clip(x_t - sigma_b * eps)  — Scheduler.pred_xstart + clamp clip (src/schedulers.py:407-409,
 * src/experiments.py:186-188). */
int nlc_pred_xstart(nlc_ctx* ctx, const float* xt, const float* eps, const float* sigma, int n_sigma, int B, int d,
                    int clip, float* x0, void* stream);

/* x_{t-1} for every pred_xprev variant (src/schedulers.py:432-449,465-473,487-496,505-514,548-562,581-599,
 * 609-627) including get_eps_logvar (:367-390).  logvar_mode: 0 none, 1 learned (v given), 2 fixedsmall,
 * 3 fixedlarge.  noise (the reference's torch.randn_like draw) may be NULL when eta == 0.  nan_flag (1 int,
 * caller-zeroed, may be NULL) is OR-ed with 1 when x_{t-1} holds a NaN (src/experiments.py:389). */
int nlc_pred_xprev(nlc_ctx* ctx, int sched, double eta, const float* x0, const float* eps, const float* xt,
                   const float* noise, const float* learned_v, int logvar_mode, float min_var_coef,
                   const float* sigma, int n_sigma, const float* sigma_prev, int n_prev, int B, int d,
                   float* x_prev, int* nan_flag, void* stream);

/* Best-x0 bookkeeping of the loop on the device (src/experiments.py:371-376): v = *loss_sum * inv_count (the batch-mean
 * constraint loss; loss_sum may already be all-reduced over the ranks of a sharded run); if v < *best_val then
 * *best_val = v and best_x0 <- x0 (n floats).  `flag` is one int of scratch.  No host read: the step stays graph-able. */
int nlc_best_update(nlc_ctx* ctx, const float* loss_sum, float inv_count, float* best_val, int* flag, const float* x0,
                    float* best_x0, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * L3 / D2 — EDM Heun sampler (src/experiments.py:777-843, 847-918).  The sample x is float64, the network runs in
 * float32.  Per-sample reductions come back as NLC_EDM_PARTS partial sums per sample (fixed order).
 * ---------------------------------------------------------------------------------------------- */
#define NLC_EDM_PARTS 16

/* x32 = float(x64) (the .to(torch.float32) of encode_edm / pred_edm, :778,:789);
 * sumsq_parts[b][p] = partial sums of x64^2 (vector_norm of refine_prior_sigma, :808); may be NULL. */
int nlc_edm_prepare(nlc_ctx* ctx, const double* x64, int B, int d, float* x32, double* sumsq_parts, void* stream);

/* denoised = double(c_skip[b]*x32 + c_out[b]*F) (:801), eps = (x64 - denoised) / div[b] (:836-840);
 * sumsq_parts = partial sums of eps^2 (normalize, src/utils.py:11-16).  denoised / sumsq_parts may be NULL. */
int nlc_edm_eps(nlc_ctx* ctx, const double* x64, const float* x32, const float* F, const float* c_skip,
                const float* c_out, const double* div, int B, int d, double* eps, double* denoised,
                double* sumsq_parts, void* stream);

/* v_i = [sqrt(d) * e_i / den_i[b]] * s_i[b]   (normalize, then the eps rescale sigma_hat/sigma_hat0, :884,:903)
 * out = v_1                      when e2 == NULL
 *     = w1 * v_1 + w2 * v_2      (Heun blend eps_ratio, :907)
 * sums_parts[b][p] = partial sums of {out^2, v_1^2, out*v_1} (normalize of the blend / cosine-similarity scale,
 * :908-916).  den_i, s_i, sums_parts may be NULL. */
int nlc_edm_mix(nlc_ctx* ctx, const double* e1, const double* den1, const double* s1, const double* e2,
                const double* den2, const double* s2, double w1, double w2, int B, int d, double* out,
                double* sums_parts, void* stream);

/* x_next = x_hat + coef[b] * post(e), post(e) = [sqrt(d)*e/den[b]] [/ eps_scale when != 0] [* mul[b]]
 * (the Euler / Heun updates, :886-890,:917). */
int nlc_edm_axpy(nlc_ctx* ctx, const double* x_hat, const double* e, const double* den, double eps_scale,
                 const double* mul, const double* coef, int B, int d, double* x_next, void* stream);

/* ------------------------------------------------------------------------------------------------
 * P0-P6 — DDNM constraint operators (functions/svd_operators.py) and the fused projection
 * x0 <- x0 - A^+(A x0 - y)  (image_sample.py:376-379).  x is NCHW fp32 [B,3,R,R] flattened.
 * ---------------------------------------------------------------------------------------------- */
#define NLC_OP_INPAINT 1   /* Inpainting          functions/svd_operators.py:324-359   */
#define NLC_OP_COLOR 2     /* Colorization        :627-667                              */
#define NLC_OP_SR_AVG 3    /* SuperResolution     :479-533                              */
#define NLC_OP_WHCS 4      /* WalshHadamardCS     :211-251                              */
#define NLC_OP_SEPARABLE 5 /* SRConv :851-931 and Deblurring :934-1014 (Kronecker SVD)  */
#define NLC_OP_DENOISE 6   /* Denoising           :442-476 (A = I)                      */
#define NLC_OP_BLOCKCS 7   /* CS (block-wise compressed sensing) :101-160               */
#define NLC_OP_GENERAL 8   /* GeneralA (dense SVD of an arbitrary small A) :173-208     */

typedef struct nlc_op nlc_op;

/* Host-side description of one operator.  The small SVD factors are computed by the host exactly as the
 * reference does (torch.svd of the 1 x r^2 / 1 x 3 / R/f x R matrices, functions/svd_operators.py:486-488,
 * 632-634, 878-885, 953-961) and handed over as plain arrays. */
typedef struct {
    int task;
    int channels, R;
    int ratio;                    /* SR_AVG: pooling factor; WHCS: compression ratio; BLOCKCS: patch edge (32) */
    const int64_t* idx_host;      /* INPAINT: missing indices (pixel*3+c); WHCS: perm[R*R]                 */
    int64_t n_idx;                /* length of idx_host; GENERAL: nx, the number of columns of A            */
    const float* U_small_host;    /* COLOR / SR_AVG: [1,1]; SEPARABLE: [m,m] row-major; GENERAL: [ny,ny]   */
    const float* V_small_host;    /* COLOR: [3,3]; SR_AVG: [r^2,r^2]; SEPARABLE: [R,R]; BLOCKCS: [E^2,E^2]; */
                                  /* GENERAL: [nx,nx]; all row-major                                        */
    const float* sing_small_host; /* COLOR / SR_AVG: [1]; GENERAL: [ny] (already thresholded)              */
    int m_small;                  /* SEPARABLE: R/f (SRConv) or R (Deblurring); BLOCKCS: cs_size = measurements */
                                  /* kept per patch; GENERAL: ny, the number of rows of A                   */
    const float* mult_host;       /* SEPARABLE: [channels, m*m] spectral multipliers used by A and At      */
    const float* pinv_mult_host;  /* SEPARABLE: [channels, m*m] zero-guarded reciprocals used by A^+       */
    const float* U_small2_host;   /* SEPARABLE, optional: right-hand factors when rows and columns are blurred   */
    const float* V_small2_host;   /* by different kernels (Deblurring2D, functions/svd_operators.py:1094-1165):  */
                                  /* A x = U (mult o (V^T X V2)) U2^T; NULL = same as U_small / V_small          */
    const float* lambda_sing_host; /* SEPARABLE, optional: [m*m] singular value per spectral position (row-major) that  */
                                  /* Lambda / Lambda_noise use (Deblurring: the un-thresholded products, :957-966,     */
                                  /* 1021); NULL = the class has no Lambda (SRConv, Deblurring2D)                       */
} nlc_op_desc;

int nlc_op_create(nlc_ctx* ctx, const nlc_op_desc* d, nlc_op** out);
void nlc_op_destroy(nlc_op* op);
int64_t nlc_op_ydim(nlc_op* op);          /* length of one measurement row y                               */
size_t nlc_op_ws(nlc_op* op, int B);      /* workspace bytes the calls below need for batch B (may be 0)   */
int nlc_op_A(nlc_op* op, const float* x, int B, float* y, void* workspace, void* stream);
int nlc_op_At(nlc_op* op, const float* y, int B, float* x, void* workspace, void* stream);
int nlc_op_Apinv(nlc_op* op, const float* y, int B, float* x, void* workspace, void* stream);
/* x0_hat = x0 - A^+(A x0 - y), fused (image_sample.py:376-379) */
int nlc_op_project(nlc_op* op, const float* x0, const float* y, int B, float* x0_hat, void* workspace, void* stream);
/* V diag(s / (s^2 + eta)) U^T y: the regularised pseudo-inverse A_functions.A_pinv_eta (functions/svd_operators.py:82-91) */
int nlc_op_Apinv_eta(nlc_op* op, const float* y, int B, double eta, float* x, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY section 8(f) rank 2 — DDNM+ (noisy measurements): the operators' Lambda / Lambda_noise
 * (functions/svd_operators.py:253-320, 361-439, 464-476, 535-623, 669-736, 1016-1091) and one fused reverse step of
 * functions/svd_ddnm.py ddnm_diffusion (:40-66) / ddnm_plus_diffusion (:101-132).
 *   Lambda(v)          = V (lambda o V^T v)                       Eq. 17
 *   Lambda_noise(v, e) = V (d1 o P v) + V (d2 o P e)              Eq. 51 (P: the re-ordering half of V^T)
 * lambda, d1, d2 are per spectral component functions of its singular value and of the four scalars below; `a` and
 * `sigma_t` are fp32 at the reference's call site (0-dim tensors, functions/svd_ddnm.py:121-132), sigma_y and eta Python
 * floats.  Operators without a Lambda in the reference (SRConv, Deblurring2D) return NLC_EINVAL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float a;        /* sqrt(alpha_bar_{t-1}) */
    float sigma_t;  /* sqrt(1 - alpha_bar_{t-1}) */
    double sigma_y; /* std of the measurement noise */
    double eta;
} nlc_ddnm_coef;
int nlc_op_lambda(nlc_op* op, const float* v, int B, const nlc_ddnm_coef* c, float* out, void* workspace, void* stream);
int nlc_op_lambda_noise(nlc_op* op, const float* v, const float* eps, int B, const nlc_ddnm_coef* c, float* out,
                        void* workspace, void* stream);
/* One reverse step, fused:  x0_t = (xt - et sqrt(1 - at)) / sqrt(at);
 *   plus = 0:  x_next = sqrt(at_next) (x0_t - A^+(A x0_t - y)) + c1 z + c2 et        (:52-62)
 *   plus = 1:  x_next = sqrt(at_next) (x0_t - Lambda A^+(A x0_t - y)) + Lambda_noise(z, et)   (:118-132)
 * xt, z, x0_t, x_next are [B, C*R*R]; et is the network output, sample b at et + b * et_stride (its first C channels are
 * used: et_stride = 2*C*R*R for a learned-variance head); at / at_next are compute_alpha's fp32 values (:10-13). */
int nlc_ddnm_step(nlc_op* op, const float* xt, const float* et, int64_t et_stride, const float* z, const float* y, int B,
                  float at, float at_next, double eta, double sigma_y, int plus, float* x0_t, float* x_next,
                  void* workspace, void* stream);
/* Time-travel step (:67-73, 133-139): x_next = sqrt(at_next) x0_t + z sqrt(1 - at_next) */
int nlc_ddnm_renoise(nlc_ctx* ctx, const float* x0_t, const float* z, int64_t n, float at_next, float* x_next,
                     void* stream);

/* out[b] = sum_i |a[b,i] - b[b,i]|: the L1 constraint residuals of Constraint_Function.loss (image_sample.py:325-333) */
int nlc_l1_diff_rows(nlc_ctx* ctx, const float* a, const float* b, int B, int64_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY section 8(f) rank 1 (first slice): per-sample restoration metrics of a finished batch on the device instead
 * of the PNG -> disk -> reload round trip (image_sample.py:671-679): s = clamp((x+1)/2, 0, 1);
 * mse[b] = mean((s - orig)^2); l1[b] = ||(2s-1) - (2 orig - 1)||_1.  x is the sampler output in [-1,1] coordinates,
 * orig01 the ground truth in [0,1], both [B, n] fp32; sample01_out (nullable) receives s.
 * ---------------------------------------------------------------------------------------------- */
int nlc_image_metrics(nlc_ctx* ctx, const float* x, const float* orig01, int B, int64_t n, float* sample01_out,
                      float* mse_out, float* l1_out, void* stream);
/* SSIM as the reference evaluates it (image_sample.py:571-582 ssim_fn -> basicsr _ssim_3d, psnr_ssim.py:171-208): both
 * images [B,3,H,W] in [0,1] are rounded to 0..255, one 11x11x11 Gaussian window (sigma 1.5, replicate padding) runs
 * over the [H,W,3] volume, ssim_out[b] = mean of the SSIM map.  workspace: nlc_ssim3d_ws(B,H,W) bytes. */
size_t nlc_ssim3d_ws(int B, int H, int W);
int nlc_ssim3d(nlc_ctx* ctx, const float* sample01, const float* orig01, int B, int H, int W, void* workspace,
               float* ssim_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SURVEY section 8(f) rank 3, first slice — the sigma-model training step (src/experiments.py:654-694) around the
 * sigma-model's own forward / backward: batch preparation and the optimizer + EMA update.
 * ---------------------------------------------------------------------------------------------- */
/* edm = 0: new_noise = noise + eta1 noise + (eta1 eta2) extra;  dist_real[b] = ||new_noise[b]||_2 / sqrt(d);
 *          noisy_x = x0 sqrt(alpha_bar[b]) + new_noise sqrt(1 - alpha_bar[b])   (:661-669, src/schedulers.py:323-329).
 * edm = 1: new_noise = noise + eta1 (noise + eta2 extra);  noisy_x = x0 + sigma[b] new_noise, sigma passed as `alpha_bar`
 *          (the EDM experiment's step, :996-1001).
 * x0, noise, extra, noisy_x, new_noise_out (nullable): [B, d]; eta1, eta2, alpha_bar, dist_real: [B]. */
int nlc_train_prepare(nlc_ctx* ctx, const float* x0, const float* noise, const float* extra, const float* eta1,
                      const float* eta2, const float* alpha_bar, int edm, int B, int64_t d, float* noisy_x,
                      float* new_noise_out, float* dist_real, void* stream);
/* One torch.optim.AdamW step (:144, :692) on a flat fp32 buffer followed by the EMA of the parameters (:233-236; ema may be
 * NULL): grads are multiplied by grad_scale first (1 / world_size after a sum all-reduce).  step counts from 1. */
int nlc_adamw_ema_step(nlc_ctx* ctx, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* ema,
                       int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       double ema_rate, double grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NLC_B200_H */
