/* tpgsr_b200 — C ABI of the B200-native TPGSR hot path (libtpgsr_b200.so).
 *
 * Conventions
 *   - every entry point returns 0 (TPGSR_OK) or a negative error code; none throws, allocates, or synchronises;
 *   - all tensor pointers are CALLER-OWNED DEVICE memory (torch: tensor.data_ptr()); scalars are passed by value;
 *   - work is enqueued on the given cudaStream_t, so a whole train step can be captured in a CUDA graph;
 *   - scratch space is passed in (`workspace`, `workspace_bytes`), sized by the matching *_workspace_bytes() query;
 *   - activations inside the library are NHWC fp32 ("pixel rows": m = (b*H + h)*W + w, `ld*` = row pitch in floats);
 *     the module boundary (what the reference's nn.Module callers see) stays NCHW fp32 — see tpgsr_nchw_to_nhwc /
 *     tpgsr_nhwc_to_nchw and the TPGSR_OUT_NCHW epilogue;
 *   - parameter layouts are the reference's state_dict layouts (OIHW conv weights, [3H,I]/[4H,I] RNN weights), so
 *     reference checkpoints load unchanged;
 *   - parameter-gradient outputs take `accumulate` (0: overwrite, 1: +=) because a shared SR model is used once per
 *     cascade stage (interfaces/super_resolution.py:306-386).
 *
 * Each declaration cites the reference call site (paths relative to the reference repo) whose library kernel it replaces;
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 */
#ifndef TPGSR_B200_H
#define TPGSR_B200_H

#include <stddef.h>
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPGSR_OK 0
#define TPGSR_ERR_BAD_ARG (-1)
#define TPGSR_ERR_LAUNCH (-2)
#define TPGSR_ERR_WORKSPACE (-3)
#define TPGSR_ERR_UNSUPPORTED (-4)
#define TPGSR_ERR_DRIVER (-5)

#define TPGSR_ACT_NONE 0
#define TPGSR_ACT_RELU 1
#define TPGSR_ACT_PRELU 2
#define TPGSR_ACT_MISH 3
#define TPGSR_ACT_TANH 4

#define TPGSR_OUT_NHWC 0
#define TPGSR_OUT_NCHW 1

/* ---- convolution / linear as gather-GEMM (fp32 CUDA cores) ------------------------------------------------------
 * nn.Conv2d / nn.Linear forward + backward: model/tsrn.py:26-43,376-381,467,495; model/stn_head.py:12-53;
 * model/crnn/crnn.py:35-67,11,24 (cuDNN / cuBLAS in the reference). */
int tpgsr_pack_conv_weights(const float* w_oihw, float* wk_fwd, float* wk_dgrad, int Co, int Ci, int KH, int KW,
                            cudaStream_t stream);
int tpgsr_conv2d_igemm(const float* x, const float* wk, const float* bias, float* y, int B, int Hi, int Wi, int Ci,
                       int Ho, int Wo, int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int act,
                       const float* act_param, int out_mode, int accumulate, cudaStream_t stream);
size_t tpgsr_conv2d_igemm_workspace_bytes(int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW);
int tpgsr_conv2d_igemm_ws(const float* x, const float* wk, const float* bias, float* y, int B, int Hi, int Wi, int Ci,
                          int Ho, int Wo, int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int act,
                          const float* act_param, int out_mode, int accumulate, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream);
size_t tpgsr_conv2d_wgrad_workspace_bytes(int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW);
int tpgsr_conv2d_wgrad(const float* x, const float* dy, float* dw_oihw, int B, int Hi, int Wi, int Ci, int Ho, int Wo,
                       int Co, int KH, int KW, int ph, int pw, int ldx, int ldy, int accumulate, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream);

/* ---- convolution on tensor cores: TMA-staged NHWC tiles -> tcgen05.mma (bf16 x {1,3} products, fp32 accumulate in TMEM) -------
 * nn.Conv2d with Cin % 64 == 0, 'same' padding (model/tsrn.py:37,376,381,467,495,43); operands are bf16 hi/lo planes. */
int tpgsr_split_bf16(const float* x, int ldx, void* hi, void* lo, long M, int C, cudaStream_t stream);
int tpgsr_pack_conv_weights_bf16(const float* w_oihw, void* fwd_hi, void* fwd_lo, void* dgrad_hi, void* dgrad_lo, int Co,
                                 int Ci, int KH, int KW, cudaStream_t stream);
int tpgsr_conv2d_tc_set_debug(long long* dbg);
int tpgsr_conv2d_tc_supported(int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW);
int tpgsr_conv2d_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y,
                    int B, int Hi, int Wi, int Ci, int Ho, int Wo, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                    const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream);
int tpgsr_conv2d_wgrad_reduce(const float* ws, float* dw_oihw, int splits, int K, int N, int Ci, int KH, int KW, int accumulate,
                              cudaStream_t stream);
int tpgsr_split_bf16_pad(const float* x, int ldx, void* hi, void* lo, long M, int C, int Cout, cudaStream_t stream);
int tpgsr_conv2d_wgrad_tc_supported(int B, int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int nprod);
size_t tpgsr_conv2d_wgrad_tc_workspace_bytes(int B, int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int nprod);
int tpgsr_conv2d_wgrad_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw_oihw, int B,
                          int H, int W, int Ci, int Co, int Cpad, int KH, int KW, int ph, int pw, int accumulate, int nprod,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- "padded-flat" tensor-core convolutions: halo rows staged ONCE in shared memory, every tap = a row offset of the same buffer ----
 * Same reference ops as tpgsr_conv2d_tc (nn.Conv2d 'same', 64 input channels: model/tsrn.py:37,376,381,467,495,43), but the
 * bf16 planes are stored [B][H+P][W+P][C] with zero pad rows/columns (tpgsr_split_bf16_flat) so shifted taps are row offsets.
 * tpgsr_conv2d_flat picks a plan per geometry (tpgsr_conv2d_flat_plan reports it: out[12] = 0 flat, 2 tall = 2-D halo tiles for KH x 1
 * kernels, 3 weights-stationary = split weights resident in tensor memory for 64 -> 64 convs with 1..12 taps).
 * Developer hooks: tpgsr_conv2d_flat_set_debug(buf) = per-CTA clock64 stamps of the next launches ([grid][16] slots; the
 * weights-stationary kernel writes [grid][32]); tpgsr_conv2d_flat_set_debug_skip(bits): bits 0-1 skip epilogue phases (timing
 * experiments), bit 2 no N-stacking of the weight planes, bit 3 no tall plan, bit 4 no weights-stationary plan. */
int tpgsr_split_bf16_flat(const float* x, int ldx, void* hi, void* lo, int B, int H, int W, int C, int Cout, int P,
                          cudaStream_t stream);
int tpgsr_conv2d_flat_supported(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod);
int tpgsr_conv2d_flat_set_debug(long long* dbg);
int tpgsr_conv2d_flat_set_debug_skip(int bits);
size_t tpgsr_conv2d_flat_plan(int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int nprod, int* out);
size_t tpgsr_conv2d_wgrad_flat_plan(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod, int* out);
int tpgsr_conv2d_flat(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y, int B,
                      int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                      const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream);
int tpgsr_conv2d_flat_planes(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, float* y, void* yp_hi,
                             void* yp_lo, int B, int H, int W, int P, int Ci, int Co, int KH, int KW, int ph, int pw, int ldy, int act,
                             const float* act_param, int out_mode, int accumulate, int nprod, cudaStream_t stream);
int tpgsr_conv2d_wgrad_flat_supported(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod);
size_t tpgsr_conv2d_wgrad_flat_workspace_bytes(int B, int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int nprod);
int tpgsr_conv2d_wgrad_flat(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, float* dw_oihw, int B,
                            int H, int W, int P, int Ci, int Co, int Cpad, int KH, int KW, int ph, int pw, int accumulate,
                            int nprod, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- all weight re-packs of a train step in one launch (pack_batched.cu) ---------------------------------------------------------------
 * The reference re-reads its OIHW weights in every cuDNN call (nn.Conv2d, model/tsrn.py:26-43,376-381,467,495); here the packed operand copies
 * of tpgsr_pack_conv_weights(_bf16) / tpgsr_kwstack_pack_weights persist across steps and are refreshed after optim.Adam.step()
 * (interfaces/super_resolution.py:424).  jobs = device table of n_jobs records of tpgsr_pack_job_bytes() bytes each:
 * { const float* w; void* o0, *o1, *o2, *o3; int kind, Co, Ci, KH, KW, C, narrow_out, transposed, flips, rows; }
 * kind 0: o0..o3 = fwd hi/lo, dgrad hi/lo (bf16);  kind 1: o0/o1 = wk_fwd / wk_dgrad (fp32);  kind 2: o0/o1 = kw-stacked hi/lo. */
int tpgsr_pack_job_bytes(void);
int tpgsr_pack_weights_batched(const void* jobs, int n_jobs, int max_elems, cudaStream_t stream);

/* ---- kw-stacking: 9x9 convs with a 4-channel side (model/tsrn.py:26,43) as 9x1 tensor-core convs on 64-channel operands --------------
 * (kwstack.cu) weight packing for tpgsr_conv2d_flat, the stacked operand split, the shift-and-add epilogue, the wgrad scatter. */
int tpgsr_kwstack_pack_weights(const float* w_oihw, void* hi, void* lo, int Co, int Ci, int KH, int KW, int narrow_out, int transposed,
                               int flip_kh, int flip_kw, int rows, cudaStream_t stream);
int tpgsr_kwstack_split(const float* x, int ldx, void* hi, void* lo, int B, int H, int W, int C, int KW, int pw, int P, int sign,
                        cudaStream_t stream);
int tpgsr_kwsum(const float* p, int ldp, const float* bias, float* y, int B, int H, int W, int C, int KW, int pw, int ldy, int act,
                const float* act_param, int out_mode, int accumulate, cudaStream_t stream);
int tpgsr_kwstack_wgrad_scatter(const float* dwp, float* dw_oihw, int Co, int Ci, int KH, int KW, int layout, int accumulate,
                                cudaStream_t stream);

/* ---- batch norm, activations, pooling, reshuffles -----------------------------------------------------------------
 * nn.BatchNorm2d/1d (model/tsrn.py:377,382,38; model/stn_head.py:19,51; model/crnn/crnn.py:48), mish
 * (model/tsrn.py:480-488), nn.PReLU (model/tsrn.py:28), nn.MaxPool2d (model/stn_head.py:36-44;
 * model/crnn/crnn.py:56-66), nn.PixelShuffle + mish (model/tsrn.py:464-477), torch.tanh (model/tsrn.py:74). */
size_t tpgsr_colreduce_workspace_bytes(long M, int C);
int tpgsr_bn_stats(const float* x, long M, int C, int ldx, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, float* mean_rstd, float* scale_shift, void* workspace,
                   size_t workspace_bytes, cudaStream_t stream);
int tpgsr_bn_eval_prepare(int C, const float* gamma, const float* beta, const float* running_mean,
                          const float* running_var, float eps, float* mean_rstd, float* scale_shift, cudaStream_t stream);
int tpgsr_bn_apply(const float* x, float* y, const float* scale_shift, long M, int C, int ldx, int ldy, int act,
                   const float* residual, int ldr, cudaStream_t stream);
/* fused variants: reduction + finalize in ONE launch (the block that draws the last ticket finalizes; `ticket` = one zeroed unsigned in
 * device memory private to the launching stream, left zero; tpgsr_bn_bwd_fused with ticket == NULL keeps the stand-alone finalize kernel —
 * the ticket path measured slower on B200 and is off by default), and apply passes that also emit the result as the padded-flat bf16 hi/lo
 * operand planes [B][H+P][W+P][C] of the neighbouring tensor-core conv (replacing a tpgsr_split_bf16_flat pass).  nn.BatchNorm2d train
 * forward / backward: model/tsrn.py:377,382,38; model/crnn/crnn.py:48; model/stn_head.py:19. */
int tpgsr_bn_stats_fused(const float* x, long M, int C, int ldx, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, float* mean_rstd, float* scale_shift, void* workspace,
                         size_t workspace_bytes, unsigned* ticket, cudaStream_t stream);
int tpgsr_bn_apply_planes(const float* x, float* y, const float* scale_shift, int B, int H, int W, int C, int act, const float* residual,
                          int P, void* hi, void* lo, cudaStream_t stream);
int tpgsr_bn_bwd_fused(const float* dy, const float* x, float* dx, const float* scale_shift, const float* mean_rstd, int B, int H, int W,
                       int C, int act, int train, float* dgamma, float* dbeta, int accumulate, int P, void* hi, void* lo, void* workspace,
                       size_t workspace_bytes, unsigned* ticket, cudaStream_t stream);
size_t tpgsr_bn_bwd_workspace_bytes(long M, int C);
int tpgsr_bn_bwd(const float* dy, const float* x, float* dx, const float* scale_shift, const float* mean_rstd, long M,
                 int C, int lddy, int ldx, int lddx, int act, int train, float* dgamma, float* dbeta, int accumulate,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream);
int tpgsr_colsum(const float* x, long M, int C, int ldx, float* out, int accumulate, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);
int tpgsr_act_fwd(const float* x, float* y, long n, int act, const float* slope, cudaStream_t stream);
size_t tpgsr_act_bwd_workspace_bytes(long n);
int tpgsr_act_bwd(const float* dy, const float* ref, float* dx, long n, int act, const float* slope, float* dslope,
                  int accumulate, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int tpgsr_add(const float* a, const float* b, float* y, long n, cudaStream_t stream);
int tpgsr_copy_cols(const float* src, int lds, float* dst, int ldd, long M, int C, float alpha, int accumulate,
                    cudaStream_t stream);
int tpgsr_add_i64(long long* x, long n, long long v, cudaStream_t stream);
/* W_ih * (W_1x1 x + b_1x1) + b_ih as one linear map (GruBlock, model/tsrn.py:499-504): wout = w2 w1, bout = w2 b1 + b2 */
int tpgsr_fuse_linear(const float* w2, const float* b2, const float* w1, const float* b1, float* wout, float* bout, int N2, int N1, int K1,
                      cudaStream_t stream);
int tpgsr_maxpool_fwd(const float* x, float* y, int* argmax, int B, int Hi, int Wi, int C, int kh, int kw, int sh,
                      int sw, int ph, int pw, cudaStream_t stream);
int tpgsr_maxpool_bwd(const float* dy, const int* argmax, float* dx, int B, int Hi, int Wi, int C, int kh, int kw,
                      int sh, int sw, int ph, int pw, cudaStream_t stream);
int tpgsr_pixelshuffle_mish_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream);
int tpgsr_pixelshuffle_mish_bwd(const float* dy, const float* x, float* dx, int B, int H, int W, int C,
                                cudaStream_t stream);
int tpgsr_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, long sb, long sc, long sh, long sw,
                       int ldd, const float* tanh_out, cudaStream_t stream);
int tpgsr_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, int lds, cudaStream_t stream);
/* F.interpolate(x, scale_factor) nearest (model/srcnn.py:138) fused with the NCHW -> NHWC boundary change */
int tpgsr_upsample_nearest_nchw_to_nhwc(const float* src, float* dst, int B, int C, int Hi, int Wi, int scale, long sb, long sc,
                                        long sh, long sw, cudaStream_t stream);

/* ---- resampling / prior formatting of the text-prior branch -------------------------------------------------------------------
 * parse_crnn_data (interfaces/base.py:806-829); F.interpolate bilinear align_corners=True (model/tsrn.py:196); InfoGen's
 * ConvTranspose2d on H=1 maps as zero-upsample + conv (model/tsrn.py:81-108); softmax + permute + prior dropout
 * (interfaces/super_resolution.py:319-321,376-382). */
int tpgsr_bicubic_gray_fwd(const float* in, long sb, long sc, long sh, long sw, int B, int Hi, int Wi, float* out, int Ho,
                           int Wo, cudaStream_t stream);
int tpgsr_bicubic_gray_bwd(const float* dout, int B, int Hi, int Wi, int Ho, int Wo, float* din, long sb, long sc, long sh,
                           long sw, int accumulate, cudaStream_t stream);
int tpgsr_bilinear_rows_fwd(const float* in, float* out, int B, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream);
int tpgsr_bilinear_rows_bwd(const float* dout, float* din, int B, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream);
int tpgsr_zero_upsample_w(const float* x, float* y, int B, int W, int C, int stride, cudaStream_t stream);
int tpgsr_zero_upsample_w_bwd(const float* dy, float* dx, int B, int W, int C, int stride, cudaStream_t stream);
int tpgsr_tconv_weight_fwd(const float* wt, float* wc, int Ci, int Co, int KH, int KW, int kh_mid, cudaStream_t stream);
int tpgsr_tconv_weight_bwd(const float* dwc, float* dwt, int Ci, int Co, int KH, int KW, int kh_mid, int accumulate,
                           cudaStream_t stream);
int tpgsr_softmax_prior_fwd(const float* logits, float* probs, float* prior, int T, int B, int C, int drop_first,
                            cudaStream_t stream);
int tpgsr_softmax_prior_bwd(const float* probs, const float* dprobs, const float* dprior, float* dlogits, int T, int B, int C,
                            int drop_first, cudaStream_t stream);

/* ---- general 2-D resampling for the SRResNet_TL text-prior branch (model/srresnet.py:139-195,223-234; BASELINE config 5) -------------
 * ConvTranspose2d on H > 1 maps as zero-upsample + stride-1 conv, F.interpolate(bilinear, align_corners=True) 2-D, nn.PixelShuffle(2). */
int tpgsr_zero_upsample2d(const float* x, float* y, int B, int H, int W, int C, int sh, int sw, cudaStream_t stream);
int tpgsr_zero_upsample2d_bwd(const float* dy, float* dx, int B, int H, int W, int C, int sh, int sw, cudaStream_t stream);
int tpgsr_tconv_weight2d_fwd(const float* wt, float* wc, int Ci, int Co, int KH, int KW, cudaStream_t stream);
int tpgsr_tconv_weight2d_bwd(const float* dwc, float* dwt, int Ci, int Co, int KH, int KW, int accumulate, cudaStream_t stream);
int tpgsr_bilinear2d_fwd(const float* in, float* out, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream);
int tpgsr_bilinear2d_bwd(const float* dout, float* din, int B, int Hi, int Wi, int C, int Ho, int Wo, int ldo, cudaStream_t stream);
int tpgsr_pixelshuffle_fwd(const float* x, float* y, int B, int H, int W, int C, cudaStream_t stream);
int tpgsr_pixelshuffle_bwd(const float* dy, float* dx, int B, int H, int W, int C, cudaStream_t stream);

/* ---- bidirectional LSTM recurrence -----------------------------------------------------------------------------------------------
 * nn.LSTM(nIn, 256, bidirectional=True) in BidirectionalLSTM: model/crnn/crnn.py:5-26 (cuDNN RNN). */
int tpgsr_lstm_set_mode(int mma);               /* developer switch: 1 (default) tensor-core recurrence kernels, 0 the FFMA cluster kernels */
int tpgsr_lstm_set_debug(long long* dbg);      /* developer instrumentation: per-step clock64 stamps of the forward kernel */
int tpgsr_lstm_transpose_whh(const float* whh, float* whhT, cudaStream_t stream);
int tpgsr_lstm_fwd(const float* gx, const float* whhT, const float* bhh, float* y, float* stash, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream);
int tpgsr_lstm_bwd(const float* dy, const float* stash, const float* whh, float* dgx, int T, int B, int hidden,
                   int batch_first, cudaStream_t stream);
int tpgsr_lstm_shift_h(const float* y, float* hprev, int T, int B, int hidden, int batch_first, cudaStream_t stream);

/* ---- bidirectional GRU recurrence ----------------------------------------------------------------------------------
 * nn.GRU(64, 32, bidirectional=True, batch_first=True) in GruBlock: model/tsrn.py:491-508 (cuDNN RNN). */
int tpgsr_gru_fwd(const float* gx, const float* whh, const float* bhh, float* y, int ldy, const float* residual,
                  int ldr, float* ysum, int ldys, float* stash, int S, int T, int inner, long outer_stride,
                  long inner_stride, long t_stride, cudaStream_t stream);
size_t tpgsr_gru_bwd_workspace_bytes(int S);
int tpgsr_gru_set_debug(long long* dbg);  /* developer instrumentation: per-step clock64 stamps of the tensor-core forward kernel */
int tpgsr_gru_set_mode(int mma);      /* developer switch: 1 (default) kernel picked per shape, 0 the warp-per-sequence FFMA kernels, 2 the tensor-core kernels */
int tpgsr_gru_bwd(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx,
                  float* dwhh, float* dbhh, float* dbih, int accumulate, int S, int T, int inner, long outer_stride,
                  long inner_stride, long t_stride, void* workspace, size_t workspace_bytes, cudaStream_t stream);

int tpgsr_gru_bwd_recurrence(const float* dy, int lddy, const float* y, int ldy, const float* stash, const float* whh, float* dgx, int S,
                             int T, int inner, long outer_stride, long inner_stride, long t_stride, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream);
int tpgsr_gru_bwd_weights(const float* dgx, const float* y, int ldy, const float* stash, float* dwhh, float* dbhh, float* dbih, int accumulate,
                          int S, int T, int inner, long outer_stride, long inner_stride, long t_stride, void* workspace,
                          size_t workspace_bytes, cudaStream_t stream);

/* ---- TPS rectification + bilinear sampler ---------------------------------------------------------------------------
 * TPSSpatialTransformer.forward + F.grid_sample: model/tps_spatial_transformer.py:10-18,97-112 (2 bmm + ATen gather). */
int tpgsr_tps_sample_fwd(const float* img, const float* ctrl, const float* inv_kernel, const float* reprT, float* out,
                         float* src, int B, int H, int W, int C, cudaStream_t stream);
int tpgsr_tps_sample_bwd(const float* dout, const float* img, const float* src, const float* inv_kernel,
                         const float* reprT, float* dctrl, int B, int H, int W, int C, int accumulate,
                         cudaStream_t stream);

/* ---- losses --------------------------------------------------------------------------------------------------------
 * ImageLoss / GradientPriorLoss: loss/image_loss.py:19-51; SemanticLoss: loss/semantic_loss.py:21-38. */
size_t tpgsr_loss_workspace_bytes(long n);
int tpgsr_image_loss_fwd(const float* out, const float* target, int B, int C, int H, int W, int gradient, float w0,
                         float w1, float out_scale, float* loss, int accumulate, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
int tpgsr_image_loss_bwd(const float* out, const float* target, const float* upstream, float scale, int B, int C, int H,
                         int W, int gradient, float w0, float w1, float* dout, int accumulate, cudaStream_t stream);
int tpgsr_semantic_loss_fwd(const float* pred, const float* gt, long n, float out_scale, float* loss, int accumulate,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);
int tpgsr_semantic_loss_bwd(const float* pred, const float* gt, long n, const float* upstream, float scale, float* dpred,
                            int accumulate, cudaStream_t stream);

/* ---- evaluation metrics -----------------------------------------------------------------------------------------------
 * calculate_psnr / SSIM on RGB: utils/ssim_psnr.py:9-15,18-78 (F.conv2d gaussian windows + ATen elementwise). */
size_t tpgsr_psnr_ssim_workspace_bytes(int B, int C);
int tpgsr_psnr_ssim(const float* a, const float* b, int B, int C, int H, int W, long sab, long sac, long sah, long saw, long sbb, long sbc,
                    long sbh, long sbw, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- optimiser step --------------------------------------------------------------------------------------------------
 * clip_grad_norm_(SR params, 0.25) + optim.Adam(lr, (beta1, 0.999)): interfaces/super_resolution.py:419-424,
 * interfaces/base.py:449-450 (ATen foreach kernels). */
size_t tpgsr_optim_workspace_bytes(long n);
int tpgsr_optim_prepare(const float* grads, long n_clipped, float max_norm, float beta1, float beta2, float grad_scale,
                        float* state, int advance, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int tpgsr_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long n, long n_clipped,
                    const float* state, float lr, float beta1, float beta2, float eps, float grad_scale,
                    cudaStream_t stream);
int tpgsr_scale_inplace(float* x, long n, const float* coef, float s, cudaStream_t stream);

/* ---- eval tail: CTC greedy decode (ctc.cu) — utils/metrics.py:71-88 get_string_crnn (argmax over classes -> collapse repeats -> drop blank 0),
 * called on every recogniser output in TextSR.eval (interfaces/super_resolution.py:855,920,995,1022-1023).  logits [T][B][C] fp32;
 * ids [B][T] int32 (decoded class indices, -1 padded), lens [B] int32.  Ties -> lowest class index. */
int tpgsr_ctc_greedy_decode(const float* logits, int T, int B, int C, int* ids, int* lens, cudaStream_t stream);

/* ---- input pipeline: resizeNormalize.__call__ (dataset/dataset.py:615-632) for a batch of ragged RGB crops (resize_input.cu) -----------------
 * img.resize((W,H), Image.BICUBIC) -> ToTensor -> optional mask channel (convert('L'), 0 if x > mean else 255), bit-identical to Pillow's 8-bit
 * resampler.  src = the crops' bytes back to back ([h][w][3] uint8 each), offsets[b] = byte offset of crop b relative to src, hs/ws = crop sizes
 * (device memory); max_h/max_w = host-known upper bounds of hs/ws; out = [B][3+mask][H][W] fp32; workspace = B*max_h*W*3 bytes. */
size_t tpgsr_resize_normalize_workspace_bytes(int B, int max_h, int W);
int tpgsr_resize_normalize(const unsigned char* src, const long long* offsets, const int* hs, const int* ws, int B, int max_h, int max_w, int H,
                           int W, int mask, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ---- input pipeline: label one-hot collate (collate.cu) — the text half of alignCollate_realWTL* (dataset/dataset.py:1172-1217, 1275-1323):
 * codes [B][L] int32 = class index of character t of label b in the alphabet "-0123456789abcdefghijklmnopqrstuvwxyz" (-1 behind the end);
 * out [B][C][1][L] fp32 one-hot, zero-padded to the longest label (the reference's label_rebatches). */
int tpgsr_label_onehot(const int* codes, int B, int L, int C, float* out, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TPGSR_B200_H */
