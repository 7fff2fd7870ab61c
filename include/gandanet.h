/*
 * gandanet.h -- C ABI of libgandanet_sm100.so
 *
 * B200 (sm_100a) kernels for the GAN-DANet generator/discriminator training
 * step.  The reference (/root/reference, Aster32/GAN-DANet) has no FFI: its hot
 * path is PyTorch ATen calls made from models/generator.py, models/discriminator.py,
 * models/losses.py and the step glue in GAN_DANet_train.ipynb.  Every entry point
 * below names the reference call site it replaces (file:line into /root/reference).
 *
 * Conventions
 *   - plain C: raw device pointers, ints, floats.  No C++/torch types.
 *   - activations are float32 NHWC; a tensor argument is (ptr, pitch, c0): the
 *     channel slice [c0, c0+C) of a buffer whose innermost dimension is `pitch`.
 *   - the caller owns every buffer including workspaces; the library never
 *     allocates device memory.
 *   - every call enqueues on the given stream and returns without synchronising.
 *   - return value: GDN_OK (0) or a negative gdn_status; gdn_last_error() gives
 *     a thread-local message.  Never throws, never exits.
 */
#ifndef GANDANET_H_
#define GANDANET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gdn_stream_t; /* cudaStream_t */

typedef enum {
  GDN_OK = 0,
  GDN_EINVAL = -1, /* bad shape / alignment / null pointer           */
  GDN_EARCH = -2,  /* device is not sm_100                            */
  GDN_ECUDA = -3,  /* CUDA runtime error, see gdn_last_error()        */
  GDN_EWORKSPACE = -4 /* workspace too small                          */
} gdn_status;

enum { GDN_ACT_NONE = 0, GDN_ACT_RELU = 1, GDN_ACT_LRELU = 2 };
enum { GDN_PREC_FP32 = 0, GDN_PREC_BF16 = 1, GDN_PREC_FP16 = 2, GDN_PREC_BF16X3 = 3, GDN_PREC_FP16X3 = 4 };

int gdn_version(void);
const char* gdn_last_error(void);
/* Checks the device is sm_100 and opts kernels into large dynamic shared memory. */
int gdn_init(int device);

/* ------------------------------------------------------------------ layout */
/* NCHW <-> NHWC slice (module boundary; replaces nothing in the reference, which is NCHW throughout). */
int gdn_nchw_to_nhwc(const float* src, float* dst, int dst_pitch, int dst_c0, int B, int C, int H, int W, gdn_stream_t s);
int gdn_nhwc_to_nchw(const float* src, int src_pitch, int src_c0, float* dst, int B, int C, int H, int W, gdn_stream_t s);
/* Conv weight OIHW -> [O][kh][kw][I] (forward operand) and -> [I][kh][kw][O] (dgrad operand). */
int gdn_weight_oihw_to_ohwi(const float* w, float* out, int O, int I, int kh, int kw, gdn_stream_t s);
int gdn_weight_oihw_to_ihwo(const float* w, float* out, int O, int I, int kh, int kw, gdn_stream_t s);

/* ------------------------------------------------------------- convolution */
/*
 * Implicit-GEMM convolution / linear layer, fp32 CUDA-core ("parity") engine.
 * Replaces nn.Conv2d at generator.py:20,34,63,108-110,148,188,214,218,222,228,
 * discriminator.py:62-67, the VGG19 convs of losses.py:58, and torch.bmm in CAM
 * (generator.py:138) via per-group weights.
 *
 *   y[b,ho,wo,n] = act(alpha * sum_{kh,kw,c} xin(b,ho,wo,kh,kw,c) * w[g][n][kh][kw][w_c0+c] + bias[n]) + res[b,ho,wo,n]
 *
 * transposed == 0: xin = x[b, ho*stride-pad+kh, wo*stride-pad+kw, c]            (forward)
 * transposed == 1: xin = x[b, (ho+pad-kh)/stride, (wo+pad-kw)/stride, c] when divisible (data gradient;
 *                  then (Hi,Wi) is the conv-output grid and (Ho,Wo) the conv-input grid)
 * groups > 1: weights differ per group of B/groups consecutive samples (w_group_stride floats apart).
 * alpha_ptr (device scalar, may be NULL => 1.0f).  res may alias y.
 * splits > 1: split-K through `ws` (splits*M*Cout floats).
 */
typedef struct {
  const float* x; int x_pitch, x_c0;
  const float* w; int w_k_pitch, w_c0; long long w_group_stride; int groups;
  float* y; int y_pitch, y_c0;
  const float* bias;
  const float* alpha_ptr;
  const float* res; int res_pitch, res_c0;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad, transposed;
  int act; float slope;
  int splits; float* ws; size_t ws_bytes;
} gdn_conv_args;
int gdn_conv2d(const gdn_conv_args* a, gdn_stream_t s);

/*
 * Reduction-over-pixels GEMM: weight gradient of the convolution above and the CAM Gram matrix
 * (generator.py:132).
 *   out[g][i][j] = scale * sum_{pixels m of group g} dy[m, dy_c0+i] * xin(m, j)      j = (kh,kw,c)
 * layout == 0: out is [groups][Cout][kh*kw*Cin] row-major (Gram / generic)
 * layout == 1: out is an OIHW weight gradient [Cout][out_cin_total][kh][kw], channels [out_c0, out_c0+Cin)
 * accumulate != 0 adds into out.  ws: splits*groups*Cout*K floats.
 */
typedef struct {
  const float* dy; int dy_pitch, dy_c0;
  const float* x; int x_pitch, x_c0;
  float* out; int layout, out_cin_total, out_c0, accumulate;
  const float* scale_ptr; float scale;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad, groups;
  int splits; float* ws; size_t ws_bytes;
} gdn_wgrad_args;
int gdn_conv2d_wgrad(const gdn_wgrad_args* a, gdn_stream_t s);
/* suggested split counts (pure host functions) */
int gdn_conv2d_suggest_splits(const gdn_conv_args* a);
int gdn_wgrad_suggest_splits(const gdn_wgrad_args* a);

/* ------------------------------------------- tensor-core convolution (tcgen05) */
/*
 * The same convolutions as gdn_conv2d / gdn_conv2d_wgrad (nn.Conv2d at generator.py:34,63,108-110,148,188,214,218,222,
 * discriminator.py:63-65, VGG19 of losses.py:58 and their autograd), as implicit GEMMs on the sm_100a tensor cores:
 * TMA-tiled bf16 NHWC operands, tcgen05.mma with fp32 accumulation in TMEM, fused bias / activation / residual epilogue.
 * precision GDN_PREC_BF16   : operands rounded to bf16 (the *_lo pointers are ignored)
 *           GDN_PREC_BF16X3 : operands split hi = bf16(v), lo = bf16(v - hi); hi*hi + lo*hi + hi*lo (about 16 mantissa bits)
 * Operands are produced by the two pack calls below; the caller owns them.
 */
/* fp32 NHWC slice -> bf16 [M][Cp], Cp = round_up(C, 8), zero padded.  v = act(x*scale[c] + shift[c]) when scale != NULL
 * (BatchNorm + ReLU applied while packing), else v = x.  lo may be NULL. */
int gdn_pack_act_bf16(const float* x, int x_pitch, int x_c0, long long M, int C, uint16_t* hi, uint16_t* lo,
                      const float* scale, const float* shift, int act, float slope, gdn_stream_t s);
/* dz = dy * act'(y) packed directly (activation backward of a fused conv+ReLU/LeakyReLU, generator.py / losses.py:58 VGG ReLUs, fused
 * into the operand packing of the gradient GEMMs: the fp32 dz of autograd is never materialised) */
int gdn_pack_actgrad_bf16(const float* dy, int dy_pitch, const float* y, int y_pitch, long long M, int C, uint16_t* hi, uint16_t* lo,
                          int act, float slope, gdn_stream_t s);
/* same with the activation output given as bf16 [M][y16_pitch] (only its sign matters) */
int gdn_pack_actgrad_bf16g(const float* dy, int dy_pitch, const uint16_t* y16, int y16_pitch, long long M, int C, uint16_t* hi, uint16_t* lo,
                           int act, float slope, gdn_stream_t s);
/* OIHW fp32 weight (input channels [i_c0, i_c0+I) of I_total) -> bf16 [kh*kw][R][Kp]:
 * transposed == 0: R = O, Kp = round_up(I, 8) (forward operand); transposed == 1: R = I, Kp = round_up(O, 8) (data gradient). */
size_t gdn_pack_weight_bf16_elems(int O, int I, int kh, int kw, int transposed);
int gdn_pack_weight_bf16(const float* w, int O, int I_total, int i_c0, int I, int kh, int kw, int transposed,
                         uint16_t* hi, uint16_t* lo, gdn_stream_t s);
/*
 * y[b,ho,wo,n] = act(alpha * sum_{kh,kw,c} xin(b,ho,wo,kh,kw,c) * w[kh,kw][n][c] + bias[n]) + res[b,ho,wo,n]
 * transposed == 0: forward, x is [B,Hi,Wi,Cin_p8] packed, w packed with transposed = 0.
 * transposed == 1: data gradient: x is the packed OUTPUT gradient [B,Hi,Wi,(conv Cout)_p8] (so Cin here = conv Cout,
 *                  Cout here = conv Cin, (Ho,Wo) = conv input grid), w packed with transposed = 1.  stride 1 or 2.
 * res may alias y (accumulation).
 */
typedef struct {
  const uint16_t* x_hi; const uint16_t* x_lo;
  const uint16_t* w_hi; const uint16_t* w_lo;
  float* y; int y_pitch, y_c0;
  const float* bias;
  const float* res; int res_pitch, res_c0;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad, transposed;
  int act; float slope;
  int precision;
  const float* alpha_ptr; /* device scalar multiplying the accumulator (gamma of CAM, generator.py:139); NULL = 1 */
  int groups;             /* > 1: per-sample weights [groups][kh*kw][R][Kp] for B/groups consecutive samples each (CAM's bmm) */
  uint16_t* y16; int y16_pitch; /* optional: the output also (or, with y == NULL, only) as bf16 [pixel][y16_pitch] = the packed operand of the
                                   next convolution (frozen VGG19 branch of losses.py:58-73: untapped feature maps never exist in fp32) */
} gdn_conv_tc_args;
int gdn_conv2d_tc(const gdn_conv_tc_args* a, gdn_stream_t s);
/* test hook: enable/disable the halo-reuse variant of the forward kernel (stride-1 3x3, narrow output tiles); returns the old setting */
int gdn_conv_tc_set_halo(int enabled);
/* test hook: enable/disable the COL mode of the forward kernel (data gradient of 3x3 stride-1 "same" convolutions with at most 24 output channels --
 * the DenseNet growth convolutions of generator.py:34 -- as one K = 3 x 80 product over the im2col'd gradient tile with resident weights) */
int gdn_conv_tc_set_col(int enabled);
/* test hook: enable/disable the role-swapped weight gradient for narrow outputs (Cout <= 32, the DenseNet growth convolutions of generator.py:34) */
int gdn_conv_tc_set_wgrad_swap(int enabled);
/* test hook: enable/disable the narrow-output weight-gradient kernel (3x3 stride-1 "same" convolutions with Cout <= 24 and Cin <= 192 in the bf16 precision --
 * the twelve DenseNet growth convolutions of generator.py:34: the 24-channel dy is the shifted operand, x is read once per pixel tile); returns the old setting */
int gdn_conv_tc_set_wgrad_col(int enabled);
/* test hook: within that kernel, the one-image-row variant (tiles of 128 consecutive pixels: 3 boxes of 130 pixels replace the 9 shifted boxes) */
int gdn_conv_tc_set_wgrad_col_row(int enabled);
/*
 * Weight gradient: out[co][out_c0+ci][kh][kw] (OIHW, out_cin_total input channels) (+)= scale * sum_pixels dy * x.
 * dy: packed [B,Ho,Wo,Cout_p8]; x: packed [B,Hi,Wi,Cin_p8].  Deterministic (split-K through ws, fixed-order reduction).
 */
typedef struct {
  const uint16_t* dy_hi; const uint16_t* dy_lo;
  const uint16_t* x_hi; const uint16_t* x_lo;
  float* out; int out_cin_total, out_c0, accumulate; float scale;
  int B, Hi, Wi, Cin, Ho, Wo, Cout, kh, kw, stride, pad;
  int precision;
  float* ws; size_t ws_bytes;
  int groups;              /* == B (1x1 only): per-sample Gram matrices out[b][Cout][Cin] = scale_ptr[0] * dy_b^T x_b, written directly
                              (the C x C energy of CAM, generator.py:132, and dA of its backward); ws unused */
  const float* scale_ptr;  /* optional device scalar multiplied into the result (both modes); NULL = 1.  groups == B == 1 (one sample, cfg4)
                              is the ordinary split-K path and needs ws */
} gdn_wgrad_tc_args;
size_t gdn_conv2d_wgrad_tc_ws_bytes(const gdn_wgrad_tc_args* a);
int gdn_conv2d_wgrad_tc(const gdn_wgrad_tc_args* a, gdn_stream_t s);

/* ----------------------------------------------- linear layer on tensor cores (TF32) */
/*
 * nn.Linear for Discriminator1.fc1 (discriminator.py:66,75: [B, 512*H/16*W/16] -> 1024, LeakyReLU) and its autograd, as
 * tcgen05.mma kind::tf32 GEMMs that stream the fp32 weight exactly once through TMA (HBM-bound: the weight is 1 GB at the
 * 256x512 output grid).  x: [Mb][K], w: [N][K], y/dz: [Mb][N], all fp32 row-major, 16-byte aligned.
 * Supported: Mb in {16,32,48,64}, N % 128 == 0, K % 256 == 0 (gdn_linear_tc_supported); everything else uses gdn_conv2d.
 */
int gdn_linear_tc_supported(int Mb, int N, int K);
size_t gdn_linear_tc_fwd_ws_bytes(int Mb, int N, int K);
/* y = act(x w^T + bias) */
int gdn_linear_tc_fwd(const float* x, const float* w, const float* bias, float* y, int Mb, int N, int K, int act, float slope,
                      float* ws, size_t ws_bytes, gdn_stream_t s);
/* dx = dz w */
int gdn_linear_tc_dgrad(const float* dz, const float* w, float* dx, int Mb, int N, int K, gdn_stream_t s);
/* dw = dz^T x (overwrites); additionally Mb % 32 == 0 and N % 256 == 0; ws holds dz^T */
size_t gdn_linear_tc_wgrad_ws_bytes(int Mb, int N, int K);
int gdn_linear_tc_wgrad(const float* dz, const float* x, float* dw, int Mb, int N, int K, float* ws, size_t ws_bytes, gdn_stream_t s);

/* ------------------------------------------------------------ thin convolutions */
/*
 * 3x3 convolutions with ONE channel on one side and C on the other (C = 32, 64 or 128): Discriminator1.conv1
 * (discriminator.py:62), the generator's final conv (generator.py:228), VGG19 conv1_1 on the channel-summed weight
 * (losses.py:58,64-65) and their gradients.  HBM-bound: coalesced fp32 CUDA-core kernels that read the wide tensor once.
 * Geometry: wide tensor V [B,Hv,Wv,C] (pitch), single-channel field S [B,Hs,Ws] dense; s = v*stride + k - pad.  w: [C][9].
 */
int gdn_thin_conv_supported(int C, int kh, int kw);
/* V = act(sum_k S[v*stride + k - pad] w[c][k] + bias[c]) + res   (flip != 0: S at v + pad - k: data gradient of the C->1 conv) */
int gdn_thin_conv_expand(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                         int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, gdn_stream_t st);
/* same, additionally writing V as bf16 [pixel][C] when v16 != NULL (packed operand of the next convolution) */
int gdn_thin_conv_expand_p(const float* s_in, const float* w, const float* bias, float* v_out, int v_pitch, const float* res, int res_pitch,
                           int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int flip, int act, float slope, uint16_t* v16, gdn_stream_t st);
/* S = sum_k sum_c V w[c][k] + bias[0] + res   (transposed == 0: C->1 forward, v = s + k - pad; 1: data gradient of the 1->C conv) */
int gdn_thin_conv_reduce(const float* v_in, int v_pitch, const float* w, const float* bias, float* s_out, const float* res,
                         int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t st);
/* the same with the wide tensor multiplied by act'(gate) while it is loaded (gate = the activation's output: ReLU for gate_slope 0, LeakyReLU otherwise):
 * the data gradient of a 1 -> C convolution followed by an activation (VGG19 conv1_1 + ReLU, losses.py:58; Discriminator1.conv1 + LeakyReLU,
 * discriminator.py:71) without a separate activation-backward pass */
int gdn_thin_conv_reduce_gated(const float* v_in, int v_pitch, const float* gate, int gate_pitch, float gate_slope, const float* w, const float* bias, float* s_out,
                               const float* res, int B, int Hv, int Wv, int C, int Hs, int Ws, int stride, int pad, int transposed, gdn_stream_t s);
/* A TAPPED 1 -> C conv + ReLU (3x3, stride 1, pad 1) of an image PAIR without its fp32 outputs: the perceptual loss's relu1_1 term (losses.py:60-72 with
 * conv1_1 on the channel-summed weight, losses.py:64-65).  One pass over the two single-channel images writes both maps only as the bf16 operands of
 * conv1_2, accumulates the L1 term and keeps 2 bits per element for the backward pass (ReLU gate and sign of the L1 gradient: one byte per pixel and
 * 4 channels) -- instead of 2 x 1.07 GB of fp32 maps written, 2.1 GB read by the L1 pass and 1 GB of L1 gradient written and re-read (256x512, batch 32).
 *   tap_pair : loss[0] (+)= scale * mean | relu(conv(fa) + cbias) - relu(conv(fb) + cbias) |; v16a / v16b [B*H*W][C] bf16 (either may be NULL);
 *              mask [B*H*W][C/4] bytes (may be NULL); ws: >= 8 * 148 doubles
 *   tap_dgrad: s_out = conv_transpose((dy + gcoef * sign) * gate, w) (+ res) with gate / sign from the mask; dy [B,H,W,C] = the gradient arriving from conv1_2
 * fa generated / fb target field [B,H,W]; C = 32 or 64, W % 4 == 0. */
int gdn_thin_conv_tap_supported(int C, int H, int W);
size_t gdn_thin_conv_tap_mask_bytes(int B, int H, int W, int C);
int gdn_thin_conv_tap_pair(const float* fa, const float* fb, const float* w, const float* cbias, int B, int H, int W, int C, float* loss, int loss_accumulate,
                           float scale, uint16_t* v16a, uint16_t* v16b, uint8_t* mask, void* ws, size_t ws_bytes, gdn_stream_t st);
int gdn_thin_conv_tap_dgrad(const float* dy, int dy_pitch, const uint8_t* mask, const float* w, float gcoef, float* s_out, const float* res,
                            int B, int H, int W, int C, gdn_stream_t st);
/* dw[c][k] (+)= sum_v V[v][c] S[v*stride + k - pad]   (flip: S at v + pad - k).  Deterministic two-stage reduction. */
size_t gdn_thin_conv_wgrad_ws_bytes(int B, int Hv, int Wv, int C);
int gdn_thin_conv_wgrad(const float* v, int v_pitch, const float* s_in, float* dw, int accumulate, int B, int Hv, int Wv, int C, int Hs, int Ws,
                        int stride, int pad, int flip, float* ws, size_t ws_bytes, gdn_stream_t st);

/* ------------------------------------------------------------ elementwise */
/* per-channel sums over M rows of an NHWC slice: out[0..C) = sum x, out[C..2C) = sum x*x (double).  BN statistics
 * (generator.py:32,61,149,189,219,223) and bias gradients.  ws: gdn_colstats_ws_bytes(M, C). */
size_t gdn_colstats_ws_bytes(long long M, int C);
int gdn_colstats(const float* x, int pitch, int c0, long long M, int C, double* out, void* ws, gdn_stream_t s);
/* BatchNorm2d train-mode finalize: sums -> mean/biased var -> scale = w*rsqrt(var+eps), shift = b - mean*scale;
 * saves mean and invstd; updates running stats (momentum, unbiased var). */
int gdn_bn_finalize(const double* sums, long long M, int C, const float* weight, const float* bias, float eps, float momentum,
                    float* running_mean, float* running_var, float* mean, float* invstd, float* scale, float* shift, gdn_stream_t s);
/* Single-launch forms (round 2).  The blocks of the reduction finish it themselves (two-level last-block-done scheme with a fixed summation
 * order: bitwise deterministic), so train-mode BatchNorm statistics (generator.py:32,61,149,189,219,223) are ONE launch instead of
 * gdn_colstats (2 launches) + gdn_bn_finalize + the num_batches_tracked increment, and the BatchNorm backward reduction is one launch instead of
 * gdn_bn_bwd_reduce (2) + gdn_sums_to_float.  ws: gdn_stat_fused_ws_bytes(M, C) bytes; counters: gdn_stat_fused_counters() 32-bit words that are
 * ZERO on entry (the finishing block restores them to zero; one buffer per stream). */
size_t gdn_stat_fused_ws_bytes(long long M, int C);
int gdn_stat_fused_counters(void);
/* statistics + gdn_bn_finalize (+ *num_batches_tracked += 1 when given) */
int gdn_bn_stats(const float* x, int pitch, int c0, long long M, int C, const float* weight, const float* bias, float eps, float momentum,
                 float* running_mean, float* running_var, long long* num_batches_tracked, float* mean, float* invstd, float* scale,
                 float* shift, void* ws, unsigned* counters, gdn_stream_t s);
/* gdn_colstats with the sums also (or only) as float: sums / fsums [2C], either may be NULL */
int gdn_colsums_f(const float* x, int pitch, int c0, long long M, int C, double* sums, float* fsums, void* ws, unsigned* counters, gdn_stream_t s);
/* gdn_bn_bwd_reduce with the sums also as float (fsums[0..C) = dbias, fsums[C..2C) = dweight; may be NULL) */
int gdn_bn_bwd_reduce_f(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, long long M, int C,
                        const float* mean, const float* invstd, const float* scale, const float* shift, int act, float slope,
                        double* sums, float* fsums, void* ws, unsigned* counters, gdn_stream_t s);
/* eval-mode: scale/shift from running statistics */
int gdn_bn_eval_coeffs(const float* weight, const float* bias, const float* running_mean, const float* running_var, float eps,
                       int C, float* scale, float* shift, gdn_stream_t s);
/* y = act(x*scale[c] + shift[c]) on NHWC slices */
int gdn_affine_act(const float* x, int x_pitch, int x_c0, float* y, int y_pitch, int y_c0, long long M, int C,
                   const float* scale, const float* shift, int act, float slope, gdn_stream_t s);
/* BN(+act) backward, stage 1: g = dy * act'(x*scale+shift); sums[0..C) = sum g, sums[C..2C) = sum g*xhat (double) */
int gdn_bn_bwd_reduce(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, long long M, int C,
                      const float* mean, const float* invstd, const float* scale, const float* shift, int act, float slope,
                      double* sums, void* ws, gdn_stream_t s);
/* stage 2: dx (+)= w*invstd*(g - sum_g/M - xhat*sum_gx/M); dweight = sum_gx; dbias = sum_g */
int gdn_bn_bwd_apply(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0,
                     float* dx, int dx_pitch, int dx_c0, int accumulate, long long M, int C,
                     const float* mean, const float* invstd, const float* weight, const float* scale, const float* shift,
                     int act, float slope, const double* sums, float* dweight, float* dbias, gdn_stream_t s);
/* stage 2 with dx written only as the bf16 operand [M][dx16_pitch] of the preceding bias-free tensor-core convolution's gradient GEMMs
 * (conv -> BN -> ReLU chains: generator.py:187-190,147-150,218-224): no fp32 dx.  C % 8 == 0, 16-byte aligned rows. */
int gdn_bn_bwd_apply16(const float* dy, int dy_pitch, int dy_c0, const float* x, int x_pitch, int x_c0, uint16_t* dx16, int dx16_pitch,
                       long long M, int C, const float* mean, const float* invstd, const float* weight, const float* scale, const float* shift,
                       int act, float slope, const double* sums, gdn_stream_t s);
/* dz = dy * act'(y) given the activation OUTPUT y (ReLU / LeakyReLU) */
int gdn_act_bwd(const float* dy, int dy_pitch, int dy_c0, const float* y, int y_pitch, int y_c0,
                float* dz, int dz_pitch, int dz_c0, long long M, int C, int act, float slope, gdn_stream_t s);
/* y (+)= alpha * x on NHWC slices */
int gdn_axpy(const float* x, int x_pitch, int x_c0, float* y, int y_pitch, int y_c0, long long M, int C, float alpha, int accumulate, gdn_stream_t s);
/* y (+)= scalar[0] * x with a DEVICE scalar (autograd's upstream gradient of a scalar loss) */
int gdn_scale_dev(const float* x, const float* scalar, float* y, long long n, int accumulate, gdn_stream_t s);
/* double sums -> float vector */
int gdn_sums_to_float(const double* sums, float* out, int n, float scale, gdn_stream_t s);

/* ------------------------------------------------------------- resampling */
/* nn.Upsample(scale_factor=2, mode='bicubic', align_corners=False), generator.py:221,225 (A=-0.75, clamped taps) */
int gdn_bicubic_up2_fwd(const float* x, float* y, int B, int H, int W, int C, gdn_stream_t s);
int gdn_bicubic_up2_bwd(const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s);
/* y = up2(x) + F.interpolate(skip, size=(2H, 2W), mode='bilinear', align_corners=False): the last up-sampling of generator.py:225 and the skip fusion of
 * :242-246 in one pass over the full-resolution tensor (x [B,H,W,C], skip [B,Hs,Ws,C], y [B,2H,2W,C]; C % 4 == 0).  Backward: gdn_bicubic_up2_bwd and gdn_bilinear_bwd of dy */
int gdn_bicubic_up2_bilinear_add_fwd(const float* x, const float* skip, float* y, int B, int H, int W, int Hs, int Ws, int C, gdn_stream_t s);
/* The generator's head (generator.py:225,228,242-246) without its 64-channel full-resolution tensor: final(up2(u) + resize(s)) is linear and the resamplers
 * act per channel, so the 3x3 convolution's channel reduction runs first at low resolution (a 1x1 convolution to T >= 9 "tap planes", plane t = kh*3 + kw),
 * the planes are resampled, and y[b][i][j] = bias[0] + sum_{kh,kw} Z[b][i+kh-1][j+kw-1][kh*3+kw] (zero outside the grid).  Z: [B,H,W,T]; y, dy: [B,H,W]. */
int gdn_tap_shift_sum(const float* Z, int T, const float* bias, float* y, int B, int H, int W, gdn_stream_t s);
/* adjoint: dZ[b][i][j][kh*3+kw] = dy[b][i-kh+1][j-kw+1] (0 outside the grid), planes 9..T-1 = 0 */
int gdn_tap_shift_expand(const float* dy, float* dZ, int T, int B, int H, int W, gdn_stream_t s);
/* the 1x1 convolution C -> T <= 16 tap planes of that head and its gradients, HBM-bound single passes (C % 4 == 0, C <= 256; wp: [T][C] row-major):
 * z[px][t] = sum_c u[px][c] wp[t][c];  du[px][c] (+)= sum_t dz[px][t] wp[t][c];  gwp[t][c] (+)= sum_px dz[px][t] u[px][c] (deterministic two-stage) */
int gdn_narrow_conv1x1_fwd(const float* u, int C, const float* wp, int T, float* z, long long M, gdn_stream_t s);
int gdn_narrow_conv1x1_dgrad(const float* dz, int T, const float* wp, int C, float* du, long long M, int accumulate, gdn_stream_t s);
size_t gdn_narrow_conv1x1_wgrad_ws_bytes(long long M, int C, int T);
int gdn_narrow_conv1x1_wgrad(const float* dz, int T, const float* u, int C, long long M, float* gwp, int accumulate, void* ws, size_t ws_bytes, gdn_stream_t s);
/* F.interpolate(size=(Ho,Wo), mode='bilinear', align_corners=False), generator.py:244: y (+)= resize(x) */
int gdn_bilinear_fwd(const float* x, float* y, int B, int Hi, int Wi, int Ho, int Wo, int C, int accumulate, gdn_stream_t s);
int gdn_bilinear_bwd(const float* dy, float* dx, int B, int Hi, int Wi, int Ho, int Wo, int C, int accumulate, gdn_stream_t s);
/* F.interpolate(scale_factor=1/f, mode='bicubic') for f = 2 or 4 (GAN_DANet_train.ipynb:226,231): NCHW in -> NHWC slice out */
int gdn_bicubic_down_nchw_to_nhwc(const float* x, float* y, int y_pitch, int y_c0, int B, int C, int Hi, int Wi, int f, gdn_stream_t s);
/* the same with the input field held as bf16 (the aux stack of GAN_DANet_train.ipynb:227 transported host -> device at 2 bytes per value: 24 -> 12 MB per
 * sample); taps are widened to fp32, the interpolation accumulates in fp32 */
int gdn_bicubic_down_nchw_to_nhwc_bf16(const uint16_t* x, float* y, int y_pitch, int y_c0, int B, int C, int Hi, int Wi, int f, gdn_stream_t s);
/* 2x2/2 max pool (VGG19 features idx 4,9,18; losses.py:58) */
int gdn_maxpool2_fwd(const float* x, float* y, int B, int H, int W, int C, gdn_stream_t s);
int gdn_maxpool2_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s);
/* pooling backward fused with the ReLU backward of the pooled map x (a conv + ReLU output that feeds only this pool: VGG19 conv1_2 / conv2_2 / conv3_4) and
 * with the operand packing of that convolution's data-gradient GEMM: dz16 [B*H*W][C] bf16 = route(dy) * (x > 0).  Even H, W; C % 8 == 0 */
int gdn_maxpool2_bwd_relu_pack16(const float* x, const float* dy, uint16_t* dz16, int B, int H, int W, int C, gdn_stream_t s);
/* the same pooling (VGG19 of losses.py:58) on bf16 feature maps [B,H,W,C], C % 8 == 0; the gradient stays fp32 (even H, W) */
int gdn_maxpool2_fwd_bf16(const uint16_t* x, uint16_t* y, int B, int H, int W, int C, gdn_stream_t s);
int gdn_maxpool2_bwd_bf16(const uint16_t* x, const float* dy, float* dx, int B, int H, int W, int C, gdn_stream_t s);

/* ---------------------------------------------------------------- attention */
/*
 * Position attention core (generator.py:115-122), the NxN map never reaches the caller:
 *   S = q k^T (no scaling), P = softmax_row(S), O = P v, y = gamma*O + x, lse = logsumexp_row(S)
 * q,k: [B,N,d] (pitch qk_pitch); v: [B,N,C] (pitch v_pitch); x,y: [B,N,C] slices; o: [B,N,C] dense; lse: [B,N].
 * precision: GDN_PREC_FP32 = fp32 CUDA-core parity engine (reference formulation, `chunk` samples of NxN scratch at a
 *            time in `ws`; chunk 0 = auto);
 *            GDN_PREC_FP16 = fused flash-style tcgen05/TMEM kernel fed by TMA (fp16 operands, fp32 accumulate,
 *            online softmax); N must be a multiple of 128, d <= 32, C <= 192;
 *            GDN_PREC_FP16X3 = the same kernels with the logit operands split into fp16 hi + lo pairs
 *            (S = q_hi k_hi^T + q_lo k_hi^T + q_hi k_lo^T, ~22 mantissa bits): the reference has no 1/sqrt(d) scale
 *            (generator.py:115-118), its logits reach +-100 where one fp16 product is off by 0.05.  This is the mode
 *            that meets the 1e-3 parity bar; same shape limits.
 * ws: gdn_pam_fwd_ws_bytes(a) bytes (operand packing for the tensor-core path, NxN scratch for the parity path).
 */
typedef struct {
  const float* q; const float* k; int qk_pitch; int d;
  const float* v; int v_pitch;
  const float* x; int x_pitch; const float* gamma;
  float* o; float* y; int y_pitch; float* lse;
  int B, N, C; int precision; int chunk;
  void* ws; size_t ws_bytes;
  const uint16_t* v16;  /* optional (tensor-core precisions only, may be NULL): v already packed as the kernel's bf16 operand [B*N][192] -- columns [0,C) = v,
                           column C = 1, the rest 0 -- e.g. written by the value projection's epilogue (gdn_conv_tc_args.y16, pitch 192) into a
                           buffer whose tail columns were initialised once; the packing pass then touches q and k only */
  uint16_t* y16; int y16_pitch; /* optional (GDN_PREC_FP16 only): y written as bf16 rows of pitch y16_pitch -- a column block of the packed
                                   operand of the convolution that consumes cat[PAM, CAM] (generator.py:156-157) INSTEAD of the fp32 y, which must then be NULL */
} gdn_pam_fwd_args;
size_t gdn_pam_fwd_ws_bytes(const gdn_pam_fwd_args* a);
int gdn_pam_fwd(const gdn_pam_fwd_args* a, gdn_stream_t s);
/*
 * Backward of the core (autograd of generator.py:115-122): given dy computes dq, dk, dv (dense [B,N,d] / [B,N,C]) and
 * rowdot[b,n] = sum_c dy*o (dgamma = sum rowdot; delta = gamma*rowdot).  dx of the residual is dy itself and is left
 * to the caller.
 */
typedef struct {
  const float* q; const float* k; int qk_pitch; int d;
  const float* v; int v_pitch;
  const float* o; const float* lse; const float* gamma;
  const float* dy; int dy_pitch;
  float* dq; float* dk; float* dv; float* rowdot;
  int B, N, C; int precision; int chunk;
  void* ws; size_t ws_bytes;
} gdn_pam_bwd_args;
size_t gdn_pam_bwd_ws_bytes(const gdn_pam_bwd_args* a);
int gdn_pam_bwd(const gdn_pam_bwd_args* a, gdn_stream_t s);
/*
 * Channel attention (generator.py:125-139): attn = softmax_row(rowmax(E) - E), E = X X^T per sample ([C][C]),
 * y = gamma * (attn X) + x.  x,y: [B,N,C] slices.  attn [B][C][C] is kept for the backward pass.
 */
int gdn_cam_fwd(const float* x, int x_pitch, const float* gamma, float* attn, float* y, int y_pitch, int B, int N, int C, gdn_stream_t s);
/* dx (+)= dy + gamma*A^T dy + (dE+dE^T) x with dE = -A*(dA - rowsum(A*dA)), dA = gamma dy^T x; dgamma[0] = sum dy*(A x) */
size_t gdn_cam_bwd_ws_bytes(int B, int N, int C);
int gdn_cam_bwd(const float* x, int x_pitch, const float* gamma, const float* attn, const float* dy, int dy_pitch,
                float* dx, int dx_pitch, int accumulate, float* dgamma, int B, int N, int C, void* ws, size_t ws_bytes, void* dot_ws, gdn_stream_t s);
/* The same two operations with every contraction on the tensor cores (tcgen05, bf16 hi+lo split operands): Gram matrices by the
 * grouped weight-gradient kernel, re-projections as 1x1 convolutions with per-sample weights.  C % 4 == 0. */
size_t gdn_cam_tc_ws_bytes(int B, int N, int C);
int gdn_cam_fwd_tc(const float* x, int x_pitch, const float* gamma, float* attn, float* y, int y_pitch, int B, int N, int C,
                   void* ws, size_t ws_bytes, gdn_stream_t s);
/* the same with y also (or, with y == NULL, only) written as bf16 rows of pitch y16_pitch (see gdn_pam_fwd_args.y16) */
int gdn_cam_fwd_tc16(const float* x, int x_pitch, const float* gamma, float* attn, float* y, int y_pitch, uint16_t* y16, int y16_pitch, int B, int N, int C,
                     void* ws, size_t ws_bytes, gdn_stream_t s);
int gdn_cam_bwd_tc(const float* x, int x_pitch, const float* gamma, const float* attn, const float* dy, int dy_pitch,
                   float* dx, int dx_pitch, int accumulate, float* dgamma, int B, int N, int C, void* ws, size_t ws_bytes, void* dot_ws, gdn_stream_t s);
/* in-place-capable row softmax of [rows][n] (negate: softmax(-x)); lse optional.  torch.softmax at generator.py:118,136 */
int gdn_row_softmax(const float* in, float* out, long long rows, int n, int negate, float* lse, gdn_stream_t s);
/* CAM softmax (generator.py:135-136): attn = softmax_row(rowmax(E) - E) for [rows][C] */
int gdn_cam_softmax(const float* e, float* attn, int rows, int C, gdn_stream_t s);
/* y = gamma[0]*o + x on [M][C] slices (generator.py:122,139) */
int gdn_gamma_residual(const float* o, int o_pitch, const float* x, int x_pitch, const float* gamma, float* y, int y_pitch, long long M, int C, gdn_stream_t s);
/* out[m] = sum_c a[m][c]*b[m][c] */
int gdn_rowdot(const float* a, int a_pitch, const float* b, int b_pitch, long long M, int C, float* out, gdn_stream_t s);
/* out[0] = sum over M*C of a*b  (deterministic two-stage; dgamma of PAM/CAM) */
size_t gdn_dot_ws_bytes(long long n);
int gdn_dot(const float* a, int a_pitch, int a_c0, const float* b, int b_pitch, int b_c0, long long M, int C, float* out, void* ws, gdn_stream_t s);

/* -------------------------------------------------------------------- losses */
/* All loss kernels are deterministic two-stage reductions; ws via gdn_dot_ws_bytes(numel). */
/* MSELoss (GAN_DANet_train.ipynb:262): loss[0] = mean((a-b)^2); if grad != NULL: grad (+)= gscale * 2(a-b)/n */
int gdn_mse(const float* a, const float* b, long long n, float* loss, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s);
/* F.l1_loss mean (losses.py:72): loss[0] (+)= mean|a-b|; grad_a (+)= gscale*sign(a-b)/n */
int gdn_l1(const float* a, const float* b, long long n, float* loss, int loss_accumulate, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s);
/* TVLoss (losses.py:81-87) on [B,H,W] single-channel fields */
int gdn_tv(const float* x, int B, int H, int W, float weight, float* loss, float* grad, float gscale, int accumulate, void* ws, gdn_stream_t s);
/* BCEWithLogitsLoss mean (GAN_DANet_train.ipynb:252-253,261) against target_ptr[i] (or the constant `target` when
 * target_ptr is NULL): loss[0] = mean(max(z,0) - z*t + log1p(exp(-|z|))); grad = gscale*(sigmoid(z)-t)/n */
int gdn_bce_logits(const float* z, int n, const float* target_ptr, float target, float* loss, float* grad, float gscale, gdn_stream_t s);
/* SSIM (losses.py:98-136) forward only, single channel [B,H,W]: out[0] = mean ssim map.  ws: gdn_ssim_ws_bytes */
size_t gdn_ssim_ws_bytes(int B, int H, int W);
int gdn_ssim(const float* a, const float* b, int B, int H, int W, float* out, void* ws, gdn_stream_t s);

/* ----------------------------------------------------------------- optimiser */
/* torch.optim.AdamW (GAN_DANet_train.ipynb:182-183) over a flat arena; grad_scale folds the 1/world of DP. step is 1-based. */
int gdn_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps, float wd,
              int step, float grad_scale, gdn_stream_t s);
/* the same update for `count` tensors (arrays of device pointers and element counts in HOST memory; each n < 2^31) with common hyper-parameters and
 * step: one launch per 64 tensors instead of one per tensor (the generator's ~100 small parameter tensors) */
int gdn_adamw_multi(int count, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, float lr, float beta1,
                    float beta2, float eps, float wd, int step, float grad_scale, gdn_stream_t s);
/* The same two updates with the STEP-DEPENDENT scalars read from device memory: dyn[0] = lr, dyn[1] = lr / (1 - beta1^step), dyn[2] = sqrt(1 - beta2^step)
 * (computed by the caller in double precision, exactly as gdn_adamw does from `step`).  A CUDA graph that captured the optimiser step bakes kernel
 * arguments but not device memory: the caller refreshes the three floats before each replay (GAN_DANet_train.ipynb:256,269 once per iteration). */
int gdn_adamw_dyn(float* p, const float* g, float* m, float* v, long long n, const float* dyn, float beta1, float beta2, float eps, float wd,
                  float grad_scale, gdn_stream_t s);
int gdn_adamw_multi_dyn(int count, float* const* p, const float* const* g, float* const* m, float* const* v, const long long* n, const float* dyn,
                        float beta1, float beta2, float eps, float wd, float grad_scale, gdn_stream_t s);
int gdn_fill(float* p, long long n, float value, gdn_stream_t s);

/* ------------------------------------------- deep-ensemble statistics and inference post-processing (f1, f3) */
/* out = keep ? (x + trend) * scale + shift : NaN on [rows][HW] fields: "+ trend", StandardScaler.inverse_transform and the
 * tpb_h == 0 -> NaN masking of test.ipynb:180-191 / deep_ensemble.ipynb:415-416,450-457.  trend ([rows][HW]) and keep
 * ([HW] bytes, 0 = masked) may be NULL. */
int gdn_destandardise(const float* x, const float* trend, const unsigned char* keep, float* out, long long rows, long long HW, float scale, float shift,
                      gdn_stream_t s);
/* out[r] = np.nanmean over the kept pixels of x[r]*scale + shift (deep_ensemble.ipynb:459-460: spatial mean per member and
 * month after masking); NaN inputs are skipped, a row without a valid pixel gives NaN.  keep may be NULL.
 * ws: gdn_masked_spatial_mean_ws_bytes(rows, HW) bytes, 16-byte aligned (per-row partial sums of the row splits). */
size_t gdn_masked_spatial_mean_ws_bytes(long long rows, long long HW);
int gdn_masked_spatial_mean(const float* x, const unsigned char* keep, long long rows, long long HW, float scale, float shift, float* out, void* ws,
                            size_t ws_bytes, gdn_stream_t s);
/* np.nanmean / np.nanstd (ddof 0) over M members (deep_ensemble.ipynb:463-464), member m at preds + m*member_stride, n
 * elements each, values de-standardised by scale/shift first.  stdev may be NULL. */
int gdn_ensemble_stats(const float* preds, long long member_stride, int M, long long n, float scale, float shift, float* mean, float* stdev, gdn_stream_t s);
/* mild_histogram_matching (test.ipynb:115-125; weight 1 = simple_histogram_matching :104-113) per sample: src [B][ns],
 * src_sorted / ref_sorted = ascending copies of the sample's source and reference values ([B][ns], [B][nt]),
 * out = (1-weight)*src + weight*np.interp(cdf_src(src), cdf_ref, ref_values), CDF arithmetic in float64 as numpy's. */
/* ascending sort of every row of [rows][n] (NaNs last): the np.sort / np.unique step of simple_histogram_matching (test.ipynb:117-121).
 * ws: gdn_sort_rows_ws_bytes(rows, n) bytes.  out may not alias x. */
size_t gdn_sort_rows_ws_bytes(int rows, int n);
int gdn_sort_rows(const float* x, float* out, int rows, int n, void* ws, size_t ws_bytes, gdn_stream_t s);
int gdn_hist_match(const float* src, const float* src_sorted, const float* ref_sorted, float* out, int B, int ns, int nt, float weight, gdn_stream_t s);
/* F.interpolate(scale_factor=(scale_h, scale_w), mode='bicubic', align_corners=False) on [rows][Hi][Wi] planes
 * (test.ipynb:553 x1.25, :559 x4); Ho = floor(Hi*scale_h), Wo = floor(Wi*scale_w) are passed by the caller. */
int gdn_bicubic_resize(const float* x, float* y, long long rows, int Hi, int Wi, int Ho, int Wo, float scale_h, float scale_w, gdn_stream_t s);
/* smooth_blend (test.ipynb:482-496): out = a*(1-mask) + b*mask inside the rectangle (r0, c0, h, w) of every [H][W] plane,
 * a elsewhere; mask [h][w] is the caller's feather mask; out may alias a. */
int gdn_blend_region(const float* a, const float* b, const float* mask, float* out, long long rows, int H, int W, int r0, int c0, int h, int w, gdn_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* GANDANET_H_ */
