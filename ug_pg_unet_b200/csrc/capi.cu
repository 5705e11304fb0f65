// extern "C" surface of libugp_b200.so (declared in include/ugp_b200.h).
// Thin forwarding only: argument validation lives with each kernel launcher.
#include "../../include/ugp_b200.h"

#include "ugp_common.cuh"

namespace ugp {
const char* last_error();
int set_option(const char*, int);

int conv_gemm_bf16(const void*, int, const void*, int, int, int, int, int, const void*, int,
                   void*, void*, int, const float*, const float*, const float*, int, float*,
                   cudaStream_t);
int conv_wgrad_bf16(const void*, int, const void*, int, const void*, int, int, int, int, int,
                    float*, cudaStream_t);
int im2col_first(const float*, int, int, int, int, int, int, void*, cudaStream_t);
int bn_finalize(float*, int, int, long long, const float*, const float*, float, float, float*,
                float*, long long*, float*, float*, float*, float*, cudaStream_t);
int bn_eval_coeffs(int, const float*, const float*, const float*, const float*, float, float*,
                   float*, cudaStream_t);
int bn_relu_apply(const void*, const float*, const float*, int, int, int, int, void*, void*,
                  cudaStream_t);
int bn_relu_apply_head(const void*, const float*, const float*, int, int, int, int, const float*, const float*, int,
                       void*, float*, cudaStream_t);
int bn_relu_bwd_ex(const void*, const void*, const void*, const float*, const float*, const float*, const float*,
                   const float*, int, int, int, int, float*, float*, void*, const float*, const float*, int, float*,
                   const float*, cudaStream_t);
int conv_dgrad_bnstats_bf16(const void*, int, int, int, int, const void*, int, void*, const void*, const float*,
                            const float*, float*, cudaStream_t);
int flip_u8(const unsigned char*, const unsigned char*, int, int, int, int, unsigned char*, cudaStream_t);
int maxpool2x2(const void*, int, int, int, int, void*, int*, cudaStream_t);
int maxpool2x2_bwd(const void*, const void*, int, int, int, int, void*, cudaStream_t);
int upsample2x(const void*, int, int, int, int, void*, cudaStream_t);
int upsample2x_bwd(const void*, int, int, int, int, void*, cudaStream_t);
int bn_relu_bwd(const void*, const void*, const void*, const float*, const float*, const float*,
                const float*, const float*, int, int, int, int, float*, float*, void*,
                cudaStream_t);
int pack_weights(const float*, int, int, void*, void*, cudaStream_t);
int unpack_wgrad(float*, int, int, float*, int, cudaStream_t);
int outconv_fwd(const void*, const float*, const float*, int, int, int, int, float*,
                cudaStream_t);
int outconv_bwd(const void*, const float*, const float*, int, int, int, int, int, void*, float*,
                float*, cudaStream_t);
int head_combine_fwd(const float* const*, const int*, int, int, int, float*, cudaStream_t);
int head_combine_bwd(const float*, int, int, int, float*, cudaStream_t);
int head_combine_fwd_hw(const float* const*, const int*, const int*, int, int, int, int, float*, cudaStream_t);
int head_combine_bwd_hw(const float*, int, int, int, int, int, float*, cudaStream_t);
int uncertainty_margin(const float*, int, int, int, int, int, float*, cudaStream_t);
int softmax_entropy(const float*, int, int, float*, cudaStream_t);
int bce_uncertainty_loss(const float*, const float*, const float*, float, float, int, long long,
                         float, float*, double*, cudaStream_t);
int ce_weighted_loss(const float*, const long long*, const float*, float, int, int, float,
                     float*, double*, cudaStream_t);
int rmsprop_step(float*, const float*, float*, long long, float, float, float, float, float,
                 cudaStream_t);
int rmsprop_multi(float* const*, const float* const*, float* const*, const long long*, int, float,
                  float, float, float, float, cudaStream_t);
int global_avg_pool(const void*, int, int, int, float*, cudaStream_t);
int global_avg_pool_bwd(const float*, int, int, int, void*, cudaStream_t);
int gap_mlp_head_fwd(const void*, int, int, int, const float*, const float*, int, const float*, const float*, int,
                     const float*, const float*, int, const float*, float, float, float, float*, float*, float*,
                     float*, unsigned int*, cudaStream_t);
int mlp_head_bwd(const float*, int, int, const float*, int, const float*, int, const float*, int, const float*,
                 float, float, float, const float*, const float*, const float*, float*, float*, float*, float*,
                 float*, float*, float*, float*, cudaStream_t);
int conv_transpose2x2_fwd(const void*, int, int, int, int, const void*, int, const float*, void*, cudaStream_t);
int conv_transpose2x2_dgrad(const void*, int, int, int, int, const void*, int, void*, cudaStream_t);
int conv_transpose2x2_wgrad(const void*, const void*, int, int, int, int, int, float*, cudaStream_t);
int pack_weights_tconv(const float*, int, int, void*, void*, cudaStream_t);
int unpack_wgrad_tconv(float*, int, int, float*, int, cudaStream_t);
int channel_sum(const void*, long long, int, float*, cudaStream_t);
int adam_multi(float* const*, const float* const*, float* const*, float* const*, const long long*, int, double,
               double, double, double, double, long long, const long long*, float, cudaStream_t);
int counter_add(long long*, long long, cudaStream_t);
bool conv_first_supported(int, int, int, int, int);
bool conv_first_wgrad_supported(int, int, int, int, int);
int conv_first_fwd(const void*, int, int, int, int, int, int, int, const float*, const float*, const void*, int, void*,
                   const float*, const float*, const float*, int, float*, cudaStream_t);
int conv_first_wgrad(const void*, int, int, int, int, int, int, int, const float*, const float*, const void*, int,
                     float*, cudaStream_t);
int im2col_first_u8(const unsigned char*, int, int, int, int, int, int, const float*, const float*, void*,
                    cudaStream_t);
int bce_uncertainty_loss_u8(const float*, const unsigned char*, const float*, float, float, int, long long,
                            float, float*, double*, cudaStream_t);
int seg_confusion(const float*, const void*, int, int, long long, unsigned char*, double*, cudaStream_t);
int cls_eval(const float*, const long long*, int, int, float*, long long*, long long*, cudaStream_t);
int rmsprop_conv_packed(float* const*, float* const*, float* const*, void* const*, void* const*,
                        const int*, const int*, int, float, float, float, float, float, cudaStream_t);
}  // namespace ugp

#define S(stream) static_cast<cudaStream_t>(stream)

extern "C" {

const char* ugp_last_error(void) { return ugp::last_error(); }
int ugp_version(void) { return 100; }
int ugp_device_sm_count(void) { return ugp::device_sm_count(); }
int ugp_set_option(const char* name, int value) { return ugp::set_option(name, value); }

int ugp_conv_gemm_bf16(const void* src0, int c0, const void* src1, int c1, int B, int H, int W,
                       int taps, const void* wpack, int cout, void* dst0, void* dst1,
                       int split_n, const float* bias, const float* scale, const float* shift,
                       int relu, float* stats, void* stream) {
  return ugp::conv_gemm_bf16(src0, c0, src1, c1, B, H, W, taps, wpack, cout, dst0, dst1, split_n,
                             bias, scale, shift, relu, stats, S(stream));
}
int ugp_conv_wgrad_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz,
                        int cout, int B, int H, int W, int taps, float* dw_packed,
                        void* stream) {
  return ugp::conv_wgrad_bf16(src0, c0, src1, c1, dz, cout, B, H, W, taps, dw_packed, S(stream));
}
int ugp_pack_weights(const float* w, int Cout, int Cin, void* wf, void* wd, void* stream) {
  return ugp::pack_weights(w, Cout, Cin, wf, wd, S(stream));
}
int ugp_unpack_wgrad(float* dwp, int Cout, int Cin, float* grad, int accumulate,
                     void* stream) {
  return ugp::unpack_wgrad(dwp, Cout, Cin, grad, accumulate, S(stream));
}
int ugp_im2col_first(const float* x, int B, int Cin, int Hin, int Win, int H, int W, void* out,
                     void* stream) {
  return ugp::im2col_first(x, B, Cin, Hin, Win, H, W, out, S(stream));
}
int ugp_bn_finalize(float* stats, int parts, int C, long long count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    long long* nbt, float* scale, float* shift, float* save_mean,
                    float* save_rstd, void* stream) {
  return ugp::bn_finalize(stats, parts, C, count, gamma, beta, eps, momentum, running_mean, running_var,
                          nbt, scale, shift, save_mean, save_rstd, S(stream));
}
int ugp_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* rm,
                       const float* rv, float eps, float* scale, float* shift, void* stream) {
  return ugp::bn_eval_coeffs(C, gamma, beta, rm, rv, eps, scale, shift, S(stream));
}
int ugp_bn_relu_apply(const void* z, const float* scale, const float* shift, int B, int H, int W,
                      int C, void* a, void* pooled, void* stream) {
  return ugp::bn_relu_apply(z, scale, shift, B, H, W, C, a, pooled, S(stream));
}
int ugp_flip_u8(const unsigned char* x, const unsigned char* flags, int B, int H, int W, int C, unsigned char* out,
                void* stream) {
  return ugp::flip_u8(x, flags, B, H, W, C, out, S(stream));
}
int ugp_bn_relu_apply_head(const void* z, const float* scale, const float* shift, int B, int H, int W, int C,
                           const float* wh, const float* bh, int nc, void* a, float* lmap, void* stream) {
  return ugp::bn_relu_apply_head(z, scale, shift, B, H, W, C, wh, bh, nc, a, lmap, S(stream));
}
int ugp_bn_relu_bwd_head(const void* z, const void* dA, const float* scale, const float* shift, const float* mean,
                         const float* rstd, const float* gamma, int B, int H, int W, int C, const float* head_dl,
                         const float* head_w, int nc, float* sums, float* dbias, void* dz, float* head_grads,
                         void* stream) {
  return ugp::bn_relu_bwd_ex(z, dA, nullptr, scale, shift, mean, rstd, gamma, B, H, W, C, sums, dbias, dz, head_dl,
                             head_w, nc, head_grads, nullptr, S(stream));
}
int ugp_conv_dgrad_bnstats_bf16(const void* dz, int c, int B, int H, int W, const void* wd, int cout, void* dA,
                                const void* z, const float* scale, const float* shift, float* raw, void* stream) {
  return ugp::conv_dgrad_bnstats_bf16(dz, c, B, H, W, wd, cout, dA, z, scale, shift, raw, S(stream));
}
int ugp_bn_relu_bwd_apply(const void* z, const void* dA, const float* raw, const float* scale, const float* shift,
                          const float* mean, const float* rstd, const float* gamma, int B, int H, int W, int C,
                          float* sums, float* dbias, void* dz, void* stream) {
  return ugp::bn_relu_bwd_ex(z, dA, nullptr, scale, shift, mean, rstd, gamma, B, H, W, C, sums, dbias, dz, nullptr,
                             nullptr, 0, nullptr, raw, S(stream));
}
int ugp_bn_relu_bwd(const void* z, const void* dA, const void* dP, const float* scale,
                    const float* shift, const float* mean, const float* rstd, const float* gamma,
                    int B, int H, int W, int C, float* sums, float* dbias, void* dz,
                    void* stream) {
  return ugp::bn_relu_bwd(z, dA, dP, scale, shift, mean, rstd, gamma, B, H, W, C, sums, dbias, dz,
                          S(stream));
}
int ugp_maxpool2x2(const void* x, int B, int H, int W, int C, void* out, int* indices,
                   void* stream) {
  return ugp::maxpool2x2(x, B, H, W, C, out, indices, S(stream));
}
int ugp_maxpool2x2_bwd(const void* x, const void* dout, int B, int H, int W, int C, void* dx,
                       void* stream) {
  return ugp::maxpool2x2_bwd(x, dout, B, H, W, C, dx, S(stream));
}
int ugp_upsample2x(const void* x, int B, int h, int w, int C, void* out, void* stream) {
  return ugp::upsample2x(x, B, h, w, C, out, S(stream));
}
int ugp_upsample2x_bwd(const void* dout, int B, int h, int w, int C, void* dx, void* stream) {
  return ugp::upsample2x_bwd(dout, B, h, w, C, dx, S(stream));
}
int ugp_outconv_fwd(const void* x, const float* w, const float* bias, int B, int hw, int C, int nc,
                    float* out, void* stream) {
  return ugp::outconv_fwd(x, w, bias, B, hw, C, nc, out, S(stream));
}
int ugp_outconv_bwd(const void* x, const float* w, const float* dl, int B, int hw, int C, int nc,
                    int accumulate, void* dx, float* dw, float* db, void* stream) {
  return ugp::outconv_bwd(x, w, dl, B, hw, C, nc, accumulate, dx, dw, db, S(stream));
}
int ugp_head_combine_fwd(const float* const* maps, const int* sizes, int n, int planes, int R,
                         float* logits, void* stream) {
  return ugp::head_combine_fwd(maps, sizes, n, planes, R, logits, S(stream));
}
int ugp_head_combine_bwd(const float* dlogits, int planes, int R, int size, float* dl,
                         void* stream) {
  return ugp::head_combine_bwd(dlogits, planes, R, size, dl, S(stream));
}
int ugp_uncertainty_margin(const float* zprev, int planes, int h, int w, int H, int W, float* A,
                           void* stream) {
  return ugp::uncertainty_margin(zprev, planes, h, w, H, W, A, S(stream));
}
int ugp_softmax_entropy(const float* logits, int B, int K, float* u, void* stream) {
  return ugp::softmax_entropy(logits, B, K, u, S(stream));
}
int ugp_bce_uncertainty_loss(const float* z, const float* t, const float* A, float alpha,
                             float pos_weight, int B, long long per_image, float grad_scale,
                             float* dlogits, double* out, void* stream) {
  return ugp::bce_uncertainty_loss(z, t, A, alpha, pos_weight, B, per_image, grad_scale, dlogits,
                                   out, S(stream));
}
int ugp_ce_weighted_loss(const float* logits, const long long* labels, const float* u, float alpha,
                         int B, int K, float grad_scale, float* dlogits, double* out,
                         void* stream) {
  return ugp::ce_weighted_loss(logits, labels, u, alpha, B, K, grad_scale, dlogits, out,
                               S(stream));
}
int ugp_rmsprop_step(float* param, const float* grad, float* square_avg, long long n, float lr,
                     float alpha, float eps, float weight_decay, float grad_scale, void* stream) {
  return ugp::rmsprop_step(param, grad, square_avg, n, lr, alpha, eps, weight_decay, grad_scale,
                           S(stream));
}

int ugp_rmsprop_multi(float* const* params, const float* const* grads, float* const* square_avgs,
                      const long long* sizes, int n_tensors, float lr, float alpha, float eps,
                      float weight_decay, float grad_scale, void* stream) {
  return ugp::rmsprop_multi(params, grads, square_avgs, sizes, n_tensors, lr, alpha, eps, weight_decay,
                            grad_scale, S(stream));
}

int ugp_rmsprop_conv_packed(float* const* w, float* const* square_avgs, float* const* dw_packed,
                            void* const* wf, void* const* wd, const int* cout, const int* cin,
                            int n_tensors, float lr, float alpha, float eps, float weight_decay,
                            float grad_scale, void* stream) {
  return ugp::rmsprop_conv_packed(w, square_avgs, dw_packed, wf, wd, cout, cin, n_tensors, lr, alpha,
                                  eps, weight_decay, grad_scale, S(stream));
}
int ugp_global_avg_pool(const void* x, int B, int HW, int C, float* out, void* stream) {
  return ugp::global_avg_pool(x, B, HW, C, out, S(stream));
}
int ugp_global_avg_pool_bwd(const float* dy, int B, int HW, int C, void* dx, void* stream) {
  return ugp::global_avg_pool_bwd(dy, B, HW, C, dx, S(stream));
}
int ugp_gap_mlp_head_fwd(const void* x, int B, int HW, int C, const float* w1, const float* b1, int H1,
                         const float* w2, const float* b2, int H2, const float* w3, const float* b3, int K,
                         const float* u, float p0, float p1, float p2, float* pooled, float* h1, float* h2,
                         float* logits, unsigned int* tickets, void* stream) {
  return ugp::gap_mlp_head_fwd(x, B, HW, C, w1, b1, H1, w2, b2, H2, w3, b3, K, u, p0, p1, p2, pooled, h1, h2,
                               logits, tickets, S(stream));
}
int ugp_mlp_head_bwd(const float* dlogits, int B, int C, const float* w1, int H1, const float* w2, int H2,
                     const float* w3, int K, const float* u, float p0, float p1, float p2, const float* pooled,
                     const float* h1, const float* h2, float* dw1, float* db1, float* dw2, float* db2,
                     float* dw3, float* db3, float* dpooled, float* scratch, void* stream) {
  return ugp::mlp_head_bwd(dlogits, B, C, w1, H1, w2, H2, w3, K, u, p0, p1, p2, pooled, h1, h2, dw1, db1, dw2,
                           db2, dw3, db3, dpooled, scratch, S(stream));
}
int ugp_conv_transpose2x2_fwd(const void* x, int B, int H, int W, int cin, const void* wf, int cout,
                              const float* bias, void* out, void* stream) {
  return ugp::conv_transpose2x2_fwd(x, B, H, W, cin, wf, cout, bias, out, S(stream));
}
int ugp_conv_transpose2x2_dgrad(const void* dout, int B, int H, int W, int cout, const void* wd, int cin,
                                void* dx, void* stream) {
  return ugp::conv_transpose2x2_dgrad(dout, B, H, W, cout, wd, cin, dx, S(stream));
}
int ugp_conv_transpose2x2_wgrad(const void* x, const void* dout, int B, int H, int W, int cin, int cout,
                                float* dw_packed, void* stream) {
  return ugp::conv_transpose2x2_wgrad(x, dout, B, H, W, cin, cout, dw_packed, S(stream));
}
int ugp_pack_weights_tconv(const float* w, int cin, int cout, void* wf, void* wd, void* stream) {
  return ugp::pack_weights_tconv(w, cin, cout, wf, wd, S(stream));
}
int ugp_unpack_wgrad_tconv(float* dw_packed, int cin, int cout, float* grad, int accumulate, void* stream) {
  return ugp::unpack_wgrad_tconv(dw_packed, cin, cout, grad, accumulate, S(stream));
}
int ugp_channel_sum(const void* x, long long pixels, int C, float* out, void* stream) {
  return ugp::channel_sum(x, pixels, C, out, S(stream));
}
int ugp_conv_first_supported(int Cin, int cout, int B, int H, int W) {
  return ugp::conv_first_supported(Cin, cout, B, H, W) ? 1 : 0;
}
int ugp_conv_first_wgrad_supported(int Cin, int cout, int B, int H, int W) {
  return ugp::conv_first_wgrad_supported(Cin, cout, B, H, W) ? 1 : 0;
}
int ugp_conv_first_fwd(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                       const float* stddev, const void* wpack, int cout, void* dst, const float* bias,
                       const float* scale, const float* shift, int relu, float* stats, void* stream) {
  return ugp::conv_first_fwd(x, is_u8, B, Cin, Hin, Win, H, W, mean, stddev, wpack, cout, dst, bias, scale, shift,
                             relu, stats, S(stream));
}
int ugp_conv_first_wgrad(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                         const float* stddev, const void* dz, int cout, float* dw_packed, void* stream) {
  return ugp::conv_first_wgrad(x, is_u8, B, Cin, Hin, Win, H, W, mean, stddev, dz, cout, dw_packed, S(stream));
}
int ugp_im2col_first_u8(const unsigned char* x, int B, int Cin, int Hin, int Win, int H, int W,
                        const float* mean, const float* stddev, void* out, void* stream) {
  return ugp::im2col_first_u8(x, B, Cin, Hin, Win, H, W, mean, stddev, out, S(stream));
}
int ugp_bce_uncertainty_loss_u8(const float* z, const unsigned char* t, const float* A, float alpha,
                                float pos_weight, int B, long long per_image, float grad_scale,
                                float* dlogits, double* out, void* stream) {
  return ugp::bce_uncertainty_loss_u8(z, t, A, alpha, pos_weight, B, per_image, grad_scale, dlogits, out,
                                      S(stream));
}
int ugp_seg_confusion(const float* z, const void* gt, int gt_is_u8, int B, long long per_image,
                      unsigned char* mask, double* out, void* stream) {
  return ugp::seg_confusion(z, gt, gt_is_u8, B, per_image, mask, out, S(stream));
}
int ugp_cls_eval(const float* logits, const long long* labels, int B, int K, float* probs, long long* preds,
                 long long* confusion, void* stream) {
  return ugp::cls_eval(logits, labels, B, K, probs, preds, confusion, S(stream));
}
int ugp_head_combine_fwd_hw(const float* const* maps, const int* heights, const int* widths, int n, int planes,
                            int RH, int RW, float* logits, void* stream) {
  return ugp::head_combine_fwd_hw(maps, heights, widths, n, planes, RH, RW, logits, S(stream));
}
int ugp_head_combine_bwd_hw(const float* dlogits, int planes, int RH, int RW, int h, int w, float* dl,
                            void* stream) {
  return ugp::head_combine_bwd_hw(dlogits, planes, RH, RW, h, w, dl, S(stream));
}
int ugp_adam_multi(float* const* params, const float* const* grads, float* const* exp_avgs,
                   float* const* exp_avg_sqs, const long long* sizes, int n_tensors, double lr, double beta1,
                   double beta2, double eps, double weight_decay, long long step, const long long* step_dev,
                   float grad_scale, void* stream) {
  return ugp::adam_multi(params, grads, exp_avgs, exp_avg_sqs, sizes, n_tensors, lr, beta1, beta2, eps,
                         weight_decay, step, step_dev, grad_scale, S(stream));
}
int ugp_counter_add(long long* counter, long long inc, void* stream) {
  return ugp::counter_add(counter, inc, S(stream));
}
}  // extern "C"
