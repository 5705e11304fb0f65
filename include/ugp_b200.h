/*
 * ugp_b200.h — C ABI of libugp_b200.so: the sm_100a kernels behind the
 * UG-PG-UNet hot path (ProgressiveUNet / PGUNet1..4 training + inference,
 * UncertaintyGuidedLoss, the trainer step).
 *
 * The reference (tridang04022004/UG-PG-UNet) has no native layer: every op
 * below replaces an ATen/cuDNN call site reached from the reference's
 * nn.Modules.  Each entry cites the reference line(s) whose arithmetic it
 * implements.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers
 *   - activations: NHWC bf16 ("void*"), 16-byte aligned, C % 64 == 0
 *   - parameters / statistics / logits / losses: fp32
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *     allocates nothing, is re-entrant, and never throws
 *   - return 0 on success, negative on error; ugp_last_error() gives the text
 *     for the calling thread
 */
#ifndef UGP_B200_H_
#define UGP_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define UGP_OK 0
#define UGP_ERR_INVALID (-1)
#define UGP_ERR_CUDA (-2)
#define UGP_ERR_DRIVER (-3)

/* Library / device introspection. */
const char* ugp_last_error(void);
int ugp_version(void);                 /* 100 * major + minor */
int ugp_device_sm_count(void);
/* Tuning switches (process-wide).  "conv_halo": 0 = per-tap TMA loads for every layer, 1/2 = halo
 * staging for the wide thin layers (see csrc/conv_halo.cu). */
int ugp_set_option(const char* name, int value);

/* ---- convolutions (tcgen05 implicit GEMM) -------------------------------
 * nn.Conv2d(k=3, padding=1) forward and input-gradient, and 1x1 GEMM.
 * Reference: UG_unet_parts.py:10,13 (DoubleConv convs), :80 (skip concat is
 * walked as two sources, never materialised).
 *   D[p, n] = sum_{tap, c} src[p + shift(tap), c] * wpack[tap][n][c]
 *   y = relu?( (D + bias) * scale + shift )          (null = identity)
 * src0/src1: NHWC bf16 [B,H,W,c0] / [B,H,W,c1] (c1 = 0 -> single source).
 * wpack: bf16 [taps][cout][c0+c1]; taps = 9 (3x3) or 1 (1x1).
 * dst0 gets output channels [0, split_n), dst1 the rest (dst1 = NULL -> all in
 * dst0); both NHWC bf16.  stats (nullable): fp32 [P][2][cout], P = ugp_device_sm_count():
 * CTA i stores its partial per-channel sum / sum of squares of the stored outputs in
 * slot i (train-mode BatchNorm2d statistics, UG_unet_parts.py:11,14); slots of CTAs that
 * do not exist are left untouched, so the buffer must be zero before the call
 * (ugp_bn_finalize leaves it zeroed).  No atomics anywhere: the statistics - and with
 * them the whole bf16 forward - are bit-reproducible from run to run.
 */
int ugp_conv_gemm_bf16(const void* src0, int c0, const void* src1, int c1,
                       int B, int H, int W, int taps, const void* wpack,
                       int cout, void* dst0, void* dst1, int split_n,
                       const float* bias, const float* scale, const float* shift,
                       int relu, float* stats, void* stream);

/* Weight gradient of the same convolution (autograd of UG_unet_parts.py:10,13):
 *   dw_packed[tap][ci][co] += sum_p src[p + shift(tap), ci] * dz[p, co]
 * dw_packed: fp32 [taps][c0+c1][cout], accumulated (caller zeroes). */
int ugp_conv_wgrad_bf16(const void* src0, int c0, const void* src1, int c1,
                        const void* dz, int cout, int B, int H, int W, int taps,
                        float* dw_packed, void* stream);

/* fp32 OIHW master weights -> bf16 GEMM operands.  wf: forward pack
 * [9][Cout][Cin]; wd (nullable): input-gradient pack [9][Cin][Cout] with the
 * taps rotated by 180 degrees.  For the first layer (Cin % 64 != 0, 9*Cin <=
 * 64) wf is [Cout][64] in im2col order k = (r*3+s)*Cin + ci and wd is unused. */
int ugp_pack_weights(const float* w_oihw, int Cout, int Cin, void* wf, void* wd,
                     void* stream);
/* Packed fp32 gradient -> OIHW fp32 (.grad layout of nn.Conv2d.weight).  dw_packed is ZEROED after it
 * has been read, so a persistent accumulation buffer is ready for the next ugp_conv_wgrad_bf16. */
int ugp_unpack_wgrad(float* dw_packed, int Cout, int Cin, float* grad_oihw,
                     int accumulate, void* stream);

/* ---- nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) ------------------
 * Reference: UpSample.up (UG_unet_parts.py:57-67), the up path of the legacy UNet (UG_unet.py:430-462).
 *   out[b, 2y+i, 2x+j, co] = bias[co] + sum_ci x[b, y, x, ci] * W[ci, co, i, j]
 * Stride == kernel size, so the layer is a 1x1 tcgen05 GEMM [B*H*W, Cin] x [Cin, (i,j,co)] whose
 * epilogue writes each output-row phase i through a strided TMA view of the NHWC result (the pixel
 * shuffle is free); the input gradient walks the two phases as the K dimension and the weight gradient
 * is one GEMM per phase.  x: NHWC bf16 [B,H,W,cin]; out / dout: NHWC bf16 [B,2H,2W,cout];
 * cin <= 1024, cout <= 512, multiples of 64.
 * wf: bf16 [(i*2+j)*cout + co][ci];  wd: bf16 [ci][(i*2+j)*cout + co]  (ugp_pack_weights_tconv).
 * dw_packed: fp32 [i][ci][j*cout + co], accumulated (caller zeroes; ugp_unpack_wgrad_tconv re-zeroes). */
int ugp_conv_transpose2x2_fwd(const void* x, int B, int H, int W, int cin, const void* wf, int cout,
                              const float* bias, void* out, void* stream);
int ugp_conv_transpose2x2_dgrad(const void* dout, int B, int H, int W, int cout, const void* wd,
                                int cin, void* dx, void* stream);
int ugp_conv_transpose2x2_wgrad(const void* x, const void* dout, int B, int H, int W, int cin,
                                int cout, float* dw_packed, void* stream);
/* fp32 [cin][cout][2][2] (nn.ConvTranspose2d.weight layout) -> the two bf16 operand packs (wd nullable),
 * and the packed fp32 gradient back to that layout (.grad of the weight). */
int ugp_pack_weights_tconv(const float* w, int cin, int cout, void* wf, void* wd, void* stream);
int ugp_unpack_wgrad_tconv(float* dw_packed, int cin, int cout, float* grad, int accumulate,
                           void* stream);
/* out[c] += sum_p x[p][c] over an NHWC bf16 tensor [pixels][C] (bias gradient of a layer with no
 * BatchNorm behind it: ConvTranspose2d.bias); out fp32 [C], caller zeroes.  C % 8 == 0, C <= 2048. */
int ugp_channel_sum(const void* x, long long pixels, int C, float* out, void* stream);

/* First-layer input packing: fp32 NCHW [B,Cin,Hin,Win] -> bf16 [B,H,W,64]
 * 3x3 neighbourhoods (zero padding), with the bilinear align_corners=True
 * resize of uncertainty_guided_trainer.py:208 / UG_unet.py:36-41 folded in
 * when (Hin,Win) != (H,W). */
int ugp_im2col_first(const float* x, int B, int Cin, int Hin, int Win, int H,
                     int W, void* out, void* stream);

/* The same packer fed with the loaders' native bytes: uint8 HWC images [B,Hin,Win,Cin] (np.array of a
 * PIL image; MoNuSegImprove/test_monuseg.py:178-181, monuseg_dataset.py:137-187) with ToTensor's /255
 * and an optional Normalize(mean, std) (Herlev/herlev_dataset.py:200-228) folded in:
 *   value = (byte / 255 - mean[c]) / std[c].   mean / std: HOST arrays of Cin floats, nullable (0 / 1). */
int ugp_im2col_first_u8(const unsigned char* x, int B, int Cin, int Hin, int Win, int H, int W,
                        const float* mean, const float* stddev, void* out, void* stream);

/* The first convolution without the packed rows in HBM (replaces ugp_im2col_first[_u8] + ugp_conv_gemm_bf16
 * with taps = 1, and ugp_conv_wgrad_bf16 on the packed rows): nn.Conv2d(n_channels, 64, 3, padding=1) at
 * UG_unet_parts.py:10 inside InConv, its autograd weight gradient, and the bilinear align_corners=True
 * resize of uncertainty_guided_trainer.py:208 / UG_unet.py:36-41 when (Hin,Win) != (H,W).  The 64-wide
 * rows are built in shared memory, directly as the tcgen05 operand tile.
 *   x        is_u8 = 0: fp32 NCHW [B,Cin,Hin,Win];  is_u8 = 1: uint8 NHWC [B,Hin,Win,Cin] loader bytes with
 *            value = (byte / 255 - mean[c]) / std[c]  (mean / std: HOST arrays, nullable, ignored for fp32)
 *   wpack    the first-layer forward pack of ugp_pack_weights: bf16 [1][cout][64], k = (r*3+s)*Cin + ci
 *   dst      bf16 NHWC [B,H,W,cout];  bias / scale / shift / relu / stats as ugp_conv_gemm_bf16
 *   dw_packed  fp32 [64][cout], accumulated (as ugp_conv_wgrad_bf16 with taps = 1, 64 source channels)
 * Supported (ugp_conv_first_supported != 0): Cin = 3, cout = 64 or 128 (the stage-4 / stage-3 networks),
 * B*H*W % 128 == 0; the weight gradient (ugp_conv_first_wgrad_supported) additionally needs cout = 64.
 * Callers keep the packer + GEMM path for everything else.  The forward results are bit-identical to that
 * path; the weight gradient differs by the order of its fp32 atomics only. */
int ugp_conv_first_supported(int Cin, int cout, int B, int H, int W);
int ugp_conv_first_wgrad_supported(int Cin, int cout, int B, int H, int W);
int ugp_conv_first_fwd(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                       const float* stddev, const void* wpack, int cout, void* dst, const float* bias,
                       const float* scale, const float* shift, int relu, float* stats, void* stream);
int ugp_conv_first_wgrad(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                         const float* stddev, const void* dz, int cout, float* dw_packed, void* stream);

/* ---- BatchNorm2d / ReLU / MaxPool2d / bilinear x2 ------------------------ */
/* Train-mode statistics -> scale/shift, running-stat update (eps, momentum as
 * nn.BatchNorm2d: biased var to normalise, unbiased into running_var,
 * num_batches_tracked += 1).  stats: [parts][2][C] partial sums written by
 * ugp_conv_gemm_bf16, summed in slot order, consumed and zeroed. */
int ugp_bn_finalize(float* stats, int parts, int C, long long count, const float* gamma,
                    const float* beta, float eps, float momentum,
                    float* running_mean, float* running_var,
                    long long* num_batches_tracked, float* scale, float* shift,
                    float* save_mean, float* save_rstd, void* stream);
/* Eval mode: scale/shift from the running statistics. */
int ugp_bn_eval_coeffs(int C, const float* gamma, const float* beta,
                       const float* running_mean, const float* running_var,
                       float eps, float* scale, float* shift, void* stream);
/* a = relu(z*scale + shift) (UG_unet_parts.py:11-12); pooled (nullable) also
 * receives MaxPool2d(2)(a) (UG_unet_parts.py:49). */
int ugp_bn_relu_apply(const void* z, const float* scale, const float* shift,
                      int B, int H, int W, int C, void* a, void* pooled,
                      void* stream);
/* Backward of conv-output -> BN(train) -> ReLU (-> MaxPool2d(2)):
 *   g = (dA + unpool(dP)) * [a > 0]; dz = BN backward of g; sums = [sum g*xhat,
 *   sum g] (= dgamma, dbeta: the grads of BatchNorm2d.weight, .bias);
 *   dbias = sum dz (overwritten).  dA or dP may be NULL. */
int ugp_bn_relu_bwd(const void* z, const void* dA, const void* dP,
                    const float* scale, const float* shift, const float* mean,
                    const float* rstd, const float* gamma, int B, int H, int W,
                    int C, float* sums, float* dbias, void* dz, void* stream);
/* Stand-alone MaxPool2d(2) with PyTorch index convention (int32 plane-local
 * h*W+w, first maximum wins); indices nullable. */
int ugp_maxpool2x2(const void* x, int B, int H, int W, int C, void* out,
                   int* indices, void* stream);
int ugp_maxpool2x2_bwd(const void* x, const void* dout, int B, int H, int W,
                       int C, void* dx, void* stream);
/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=True)
 * (UG_unet_parts.py:78) and its adjoint. */
int ugp_upsample2x(const void* x, int B, int h, int w, int C, void* out,
                   void* stream);
int ugp_upsample2x_bwd(const void* dout, int B, int h, int w, int C, void* dx,
                       void* stream);

/* ---- logit head ---------------------------------------------------------- */
/* OutConv 1x1 (UG_unet_parts.py:84-91): x NHWC bf16 [B,hw,C] -> fp32 NCHW
 * [B,nc,hw]; C in {64,128,256}, nc <= 8. */
int ugp_outconv_fwd(const void* x, const float* w, const float* bias, int B,
                    int hw, int C, int nc, float* out, void* stream);
int ugp_outconv_bwd(const void* x, const float* w, const float* dl, int B, int hw,
                    int C, int nc, int accumulate, void* dx, float* dw, float* db,
                    void* stream);
/* logits = sum_s bilinear_up(maps[s]) (UG_unet.py:294-303); maps ordered
 * coarse -> fine, sizes[s] = side of map s, the last one has side R. */
int ugp_head_combine_fwd(const float* const* maps, const int* sizes, int n,
                         int planes, int R, float* logits, void* stream);
int ugp_head_combine_bwd(const float* dlogits, int planes, int R, int size,
                         float* dl, void* stream);
/* The same for rectangular images (H != W): map s is heights[s] x widths[s], logits RH x RW. */
int ugp_head_combine_fwd_hw(const float* const* maps, const int* heights, const int* widths, int n,
                            int planes, int RH, int RW, float* logits, void* stream);
int ugp_head_combine_bwd_hw(const float* dlogits, int planes, int RH, int RW, int h, int w,
                            float* dl, void* stream);

/* ---- uncertainty + loss -------------------------------------------------- */
/* A = 1 - 2*|up(sigmoid(z_prev)) - 0.5| (UG_unet.py:44-57). */
int ugp_uncertainty_margin(const float* zprev, int planes, int h, int w, int H,
                           int W, float* A, void* stream);
/* u_b = H(softmax(z_b)) / ln K (Herlev/train_herlev.py:253-257). */
int ugp_softmax_entropy(const float* logits, int B, int K, float* u, void* stream);
/* BCEWithLogitsLoss(pos_weight, 'none') * (1 + alpha*A) summed, with every
 * metric of the trainer step (UG_unet.py:61-94;
 * uncertainty_guided_trainer.py:64-65,90-123,177-181) and dlogits =
 * grad_scale * d(sum l*w)/dz in one pass.  out: double[5 + 3*B] =
 * {sum l*w, sum l, sum A, sum A^2, #(pred != t), per image {sum pred*t,
 * sum pred, sum t}}.  A and dlogits nullable. */
int ugp_bce_uncertainty_loss(const float* z, const float* t, const float* A,
                             float alpha, float pos_weight, int B,
                             long long per_image, float grad_scale, float* dlogits,
                             double* out, void* stream);
/* The same with the mask stored as uint8 0/1 bytes (the loaders' format before `.float()`). */
int ugp_bce_uncertainty_loss_u8(const float* z, const unsigned char* t, const float* A, float alpha,
                                float pos_weight, int B, long long per_image, float grad_scale,
                                float* dlogits, double* out, void* stream);
/* mean(CE_i * (1 + alpha*u_i)) pieces (Herlev/train_herlev.py:274-281):
 * out: double[3] = {sum CE*w, sum CE, #correct}. */
int ugp_ce_weighted_loss(const float* logits, const long long* labels,
                         const float* u, float alpha, int B, int K,
                         float grad_scale, float* dlogits, double* out,
                         void* stream);

/* nn.AdaptiveAvgPool2d(1) + Flatten on an NHWC bf16 feature map [B][HW][C] -> fp32 [B][C] (the
 * Herlev classifier head input, Herlev/train_herlev.py:66-68), and its adjoint
 * dx[b][p][c] = dy[b][c] / HW (bf16).  C multiple of 64 (fwd) / 8 (bwd). */
int ugp_global_avg_pool(const void* x, int B, int HW, int C, float* out, void* stream);
int ugp_global_avg_pool_bwd(const float* dy, int B, int HW, int C, void* dx, void* stream);

/* ---- evaluation reductions ---------------------------------------------------------
 * Segmentation evaluator (MoNuSegImprove/test_monuseg.py:186-201, 244-258, 264-297): pred = z > 0
 * (sigmoid(z) > 0.5); out: double[B][4] = {TP, sum pred, sum gt, sum sigmoid(z)} per image - IoU, Dice,
 * accuracy, precision, recall, specificity and the mean-probability "confidence" follow from these and the
 * pixel count.  gt: fp32 or (gt_is_u8) uint8 0/1, nullable; mask (nullable): uint8 0/1 prediction. */
int ugp_seg_confusion(const float* z, const void* gt, int gt_is_u8, int B, long long per_image,
                      unsigned char* mask, double* out, void* stream);
/* Classification evaluator (Herlev/test_herlev.py:128-133, 196-204, 232): probs = softmax(logits)
 * [B][K], preds = argmax int64 [B], confusion[label][pred] += 1 (int64 [K][K], caller zeroes);
 * probs / preds / labels+confusion nullable. */
int ugp_cls_eval(const float* logits, const long long* labels, int B, int K, float* probs,
                 long long* preds, long long* confusion, void* stream);

/* ---- optimiser ----------------------------------------------------------- */
/* torch.optim.RMSprop step (uncertainty_guided_trainer.py:84-88; alpha,
 * eps, weight_decay, momentum 0, not centered) over a flat fp32 buffer:
 *   g = grad*grad_scale + wd*p; v = alpha*v + (1-alpha)*g*g;
 *   p -= lr * g / (sqrt(v) + eps)
 */
int ugp_rmsprop_step(float* param, const float* grad, float* square_avg,
                     long long n, float lr, float alpha, float eps,
                     float weight_decay, float grad_scale, void* stream);

/* The same update for n_tensors separate tensors in ONE launch (host arrays of device pointers and
 * element counts; at most 96 tensors per launch, more are chunked). */
int ugp_rmsprop_multi(float* const* params, const float* const* grads, float* const* square_avgs,
                      const long long* sizes, int n_tensors, float lr, float alpha, float eps,
                      float weight_decay, float grad_scale, void* stream);

/* torch.optim.Adam step (Herlev/train_herlev.py:170-176: betas (.9, .999), eps 1e-8, L2 weight decay, no
 * amsgrad) for n_tensors tensors in one launch per 64 tensors; `step` = the 1-based step count, or, when
 * step_dev (a DEVICE int64, nullable) is given, read from there at run time so that CUDA-graph replays
 * advance it with ugp_counter_add:
 *   g = grad*grad_scale + wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g
 *   p -= lr/(1-b1^step) * m / (sqrt(v)/sqrt(1-b2^step) + eps) */
int ugp_adam_multi(float* const* params, const float* const* grads, float* const* exp_avgs,
                   float* const* exp_avg_sqs, const long long* sizes, int n_tensors, double lr,
                   double beta1, double beta2, double eps, double weight_decay, long long step,
                   const long long* step_dev, float grad_scale, void* stream);
/* *counter += inc on the device (the optimiser's step count inside a captured graph). */
int ugp_counter_add(long long* counter, long long inc, void* stream);

/* RMSprop for the 3x3 convolution weights straight from the PACKED gradient that
 * ugp_conv_wgrad_bf16 accumulates (replaces ugp_unpack_wgrad + optimizer.step() of
 * uncertainty_guided_trainer.py:231 + ugp_pack_weights for the next forward).  Per tensor i:
 *   g = dw_packed[i]*grad_scale + wd*w[i]; v = alpha*v + (1-alpha)*g*g; w -= lr*g/(sqrt(v)+eps)
 * w / square_avgs: fp32 OIHW [cout][cin][3][3]; dw_packed: fp32 [9][cin][cout], ZEROED on return;
 * wf / wd: the bf16 operand packs of ugp_pack_weights, rewritten from the new weights (wd[i] may be
 * null).  cout, cin multiples of 32.  One launch per 40 tensors. */
int ugp_rmsprop_conv_packed(float* const* w, float* const* square_avgs, float* const* dw_packed,
                            void* const* wf, void* const* wd, const int* cout, const int* cin,
                            int n_tensors, float lr, float alpha, float eps, float weight_decay,
                            float grad_scale, void* stream);

/* Input gradient of the SECOND convolution of a DoubleConv (autograd of UG_unet_parts.py:12-13) fused with
 * the reduction pass of the first BatchNorm's backward (autograd of :11-12): dA = conv3x3(dz, wd) exactly as
 * ugp_conv_gemm_bf16(taps = 9), and in the epilogue, while the tile is in shared memory,
 *   raw[0][c] = sum_p dA[p][c] * [a[p][c] > 0],   raw[1][c] = sum_p dA[p][c] * [a[p][c] > 0] * z[p][c]
 * with z the stored conv output of that BatchNorm (NHWC bf16 [B][H][W][cout]) and a = z*scale + shift.
 * dz: NHWC bf16 [B][H][W][c]; wd: the input-gradient weight pack; raw: fp32 [2][cout] (zeroed here,
 * accumulated with fp32 atomics).  H, W must give power-of-two tiles (the fast kernels). */
int ugp_conv_dgrad_bnstats_bf16(const void* dz, int c, int B, int H, int W, const void* wd, int cout, void* dA,
                                const void* z, const float* scale, const float* shift, float* raw, void* stream);

/* Apply pass of ugp_bn_relu_bwd alone, taking the sums of ugp_conv_dgrad_bnstats_bf16: dz as there;
 * sums [2][C] receives dgamma = rstd*(raw[1] - mean*raw[0]) and dbeta = raw[0]; dbias accumulates sum dz. */
int ugp_bn_relu_bwd_apply(const void* z, const void* dA, const float* raw, const float* scale, const float* shift,
                          const float* mean, const float* rstd, const float* gamma, int B, int H, int W, int C,
                          float* sums, float* dbias, void* dz, void* stream);

/* Joint flips of loader bytes on the device: TF.hflip / TF.vflip of MoNuSegImprove/monuseg_dataset.py:150-158
 * (one decision per sample, applied to image and mask alike).  x, out: [B][H][W][C] bytes (out != x);
 * flags: B device bytes, bit 0 = horizontal flip, bit 1 = vertical flip.  Pure byte moves: bit-exact. */
int ugp_flip_u8(const unsigned char* x, const unsigned char* flags, int B, int H, int W, int C, unsigned char* out,
                void* stream);

/* BatchNorm + ReLU of a conv output (as ugp_bn_relu_apply) with the 1x1 head fused: the decoder outputs of
 * PGUNet2..4 each feed an OutConv (UG_unet_parts.py:84-91; UG_unet.py:214-217, 250-252, 294-297).
 * lmap[b][k][p] = bh[k] + sum_c wh[k][c] * a[p][c], fp32 NCHW [B][nc][H][W], from the fp32 activation while
 * it is in registers.  C in {64, 128, 256}, nc in {1, 2}; wh fp32 [nc][C], bh fp32 [nc]. */
int ugp_bn_relu_apply_head(const void* z, const float* scale, const float* shift, int B, int H, int W, int C,
                           const float* wh, const float* bh, int nc, void* a, float* lmap, void* stream);

/* Backward of the above (autograd of the same reference lines): ugp_bn_relu_bwd whose activation gradient is
 * dA (nullable) + sum_k head_dl[b][k][p] * head_w[k][c]; head_grads [nc*C + nc] receives the head's weight
 * gradient (sum_p dl * a) and bias gradient (sum_p dl) from the same pass (zeroed here).  No pooled gradient. */
int ugp_bn_relu_bwd_head(const void* z, const void* dA, const float* scale, const float* shift, const float* mean,
                         const float* rstd, const float* gamma, int B, int H, int W, int C, const float* head_dl,
                         const float* head_w, int nc, float* sums, float* dbias, void* dz, float* head_grads,
                         void* stream);

/* Herlev classification head, reference Herlev/train_herlev.py:66-77 (the nn.Sequential
 * AdaptiveAvgPool2d(1), Flatten, Dropout(.5), Linear(C,H1), ReLU, Dropout(.3), Linear(H1,H2), ReLU,
 * Dropout(.2), Linear(H2,K)) and :104-110 (forward = classifier(features)) as ONE kernel.
 * x: NHWC bf16 feature map [B][HW][C] (C % 64 == 0).  w*: fp32 nn.Linear weights [out][in], b*: biases.
 * u: null (eval: dropout inactive) or fp32 uniforms [B][C+H1+H2]; an element is kept iff u >= p and scaled
 * by 1/(1-p) (p0, p1, p2 = the three dropout probabilities).  Outputs: pooled [B][C] (the average pool),
 * h1 [B][H1], h2 [B][H2] (post-ReLU, pre-dropout: what backward needs), logits [B][K].
 * tickets: B zeroed uint32 counters (left zeroed).  Deterministic. */
int ugp_gap_mlp_head_fwd(const void* x, int B, int HW, int C, const float* w1, const float* b1, int H1,
                         const float* w2, const float* b2, int H2, const float* w3, const float* b3, int K,
                         const float* u, float p0, float p1, float p2, float* pooled, float* h1, float* h2,
                         float* logits, unsigned int* tickets, void* stream);

/* Backward of the three Linear layers above (autograd of train_herlev.py:104-110) in one launch (a cluster of
 * 8 CTAs): dw* / db* (overwritten, fp32, nn.Linear layouts) and dpooled [B][C], the gradient of the pooled
 * vector (feed it to ugp_global_avg_pool_bwd).  pooled / h1 / h2 / u as produced by / passed to the forward;
 * scratch: B*(H1+H2) floats.  B <= 64. */
int ugp_mlp_head_bwd(const float* dlogits, int B, int C, const float* w1, int H1, const float* w2, int H2,
                     const float* w3, int K, const float* u, float p0, float p1, float p2, const float* pooled,
                     const float* h1, const float* h2, float* dw1, float* db1, float* dw2, float* db2,
                     float* dw3, float* db3, float* dpooled, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UGP_B200_H_ */
