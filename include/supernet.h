/* supernet.h -- C ABI of libsupernet_b200.so: the SUPER-Net moment-propagation hot path on B200 (sm_100a).
 *
 * The reference (/root/reference, TensorFlow/Keras) has no plugin/FFI boundary: the path sits behind
 * Keras Layer.__call__ with a (mean, sigma) -> (mean, sigma) surface, "sigma" being the VARIANCE
 * (SURVEY.md 0.3, 8b).  This header is the boundary a binding for those layers would call; each entry
 * point cites the reference interface it replaces (file:line into /root/reference).
 *
 * Conventions
 *   - plain C: POD structs, raw device pointers, sizes; no C++/torch types; nothing throws.
 *   - every function returns SN_OK (0) or a negative sn_status; sn_last_error() gives the thread-local text.
 *   - all work is enqueued on the caller's stream (sn_stream_t == cudaStream_t), asynchronously; no
 *     host synchronisation, no device allocation (the caller owns every buffer, workspace included),
 *     so a sequence of calls is CUDA-graph capturable.  No pointer is retained after return.
 *   - there is NO CPU fallback and no alternate backend: unsupported configurations are errors.
 *   - "f32" tensors: NHWC contiguous float (the reference's layout); weights HWIO
 *     [k,k,Cin,Cout] float (Brats.py:55,108), w_sigma is the RAW pre-softplus [Cout] vector (Brats.py:59-63).
 *   - "packed" tensors (FAST mode): bf16 [B][H][W][3][C]: per pixel the planes mean_hi, mean_lo (mean = hi + lo,
 *     ~16 mantissa bits) and variance.  Written by the producing kernel's epilogue, consumed directly by
 *     TMA -> tcgen05 (DESIGN.md "data layout").
 */
#ifndef SUPERNET_B200_H_
#define SUPERNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SN_ABI_VERSION 2

typedef void* sn_stream_t; /* cudaStream_t */

typedef enum sn_status {
  SN_OK = 0,
  SN_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, inconsistent geometry          */
  SN_ERR_UNSUPPORTED = -2,  /* valid but not implemented (e.g. stride != 1, k > 3 on the TC path) */
  SN_ERR_MISALIGNED = -3,   /* pointer / channel count violates a 16-byte TMA/vector requirement */
  SN_ERR_ARCH = -4,         /* device is not sm_100                                             */
  SN_ERR_LAUNCH = -5,       /* CUDA launch/runtime error                                        */
  SN_ERR_DRIVER = -6        /* tensor-map encode (driver entry point) failed                    */
} sn_status;

enum { SN_CONV_RELU = 1 };  /* fuse the ReLU moment gate (Brats.py:233-238) into the conv epilogue */

int sn_version(void);
const char* sn_last_error(void);
/* SN_OK when the current device can run the kernels (compute capability 10.x). */
int sn_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * FP32 mode (CUDA-core implicit GEMM, fp32 operands and accumulation; parity bar 1e-5).
 * ------------------------------------------------------------------------------------------------ */

/* Geometry of one stride-1 VALID moment convolution (the only mode the reference uses, Brats.py:37,89). */
typedef struct sn_conv_desc {
  int32_t batch, in_h, in_w, cin, cout, ksize;
  int32_t flags; /* SN_CONV_* */
  int32_t reserved;
} sn_conv_desc;

/* myConv_input.call (Brats.py:65-76) when var_in == NULL, myConv_intermediate.call (Brats.py:118-137)
 * otherwise:  mu_out = mu_in (*) W ;  var_out = var_in (*) W^2 + softplus(w_sigma)[n] * box_k(sum_c mu_in^2 + var_in).
 * rsum_out (optional, [B,Ho,Wo]) receives box_k(sum_c mu^2+var), which bwd_weight needs.
 * With SN_CONV_RELU the outputs are post-ReLU (Brats.py:233-238) and relu_mask-free: the gate is mu_out > 0. */
int sn_conv_moments_fwd(const sn_conv_desc* d, const float* mu_in, const float* var_in, const float* w_mu,
                        const float* w_sigma, float* mu_out, float* var_out, float* rsum_out, sn_stream_t st);

/* Data gradient of the above (what tf.GradientTape derives at Brats.py:578,593):
 *   g_mu_in = g_mu_out (*)^T W + 2 mu_in . box^T(t),  g_var_in = g_var_out (*)^T W^2 + box^T(t),  t = sum_n g_var_out s_n.
 * g_var_in may be NULL (first layer: the input is deterministic).  flags must not contain SN_CONV_RELU. */
int sn_conv_moments_bwd_data(const sn_conv_desc* d, const float* g_mu_out, const float* g_var_out,
                             const float* mu_in, const float* w_mu, const float* w_sigma, float* g_mu_in,
                             float* g_var_in, sn_stream_t st);

/* Weight gradient:  g_w_mu = corr(mu_in, g_mu_out) + 2 W . corr(var_in, g_var_out);
 *   g_w_sigma[n] = sigmoid(w_sigma[n]) * sum_p g_var_out[p,n] * rsum[p].   Outputs are overwritten.
 * var_in may be NULL (first layer). */
int sn_conv_moments_bwd_weight(const sn_conv_desc* d, const float* mu_in, const float* var_in,
                               const float* g_mu_out, const float* g_var_out, const float* rsum,
                               const float* w_mu, const float* w_sigma, float* g_w_mu, float* g_w_sigma,
                               sn_stream_t st);

/* myReLU.call + grad_ReLU (Brats.py:220-238): mu_out = max(mu,0), var_out = var * 1[mu > 0]. In-place allowed. */
int sn_relu_moments_fwd(size_t n, const float* mu, const float* var, float* mu_out, float* var_out, sn_stream_t st);
/* g_* _in = g_* _out * 1[mu_in > 0] (the gate is a constant for autodiff). */
int sn_relu_moments_bwd(size_t n, const float* mu_in, const float* g_mu_out, const float* g_var_out,
                        float* g_mu_in, float* g_var_in, sn_stream_t st);

/* mymaxpooling.call + get_pooled (Brats.py:171-174,206-216): 2x2/2 SAME max-pool of the mean, variance taken at
 * the arg-max.  Output is [B,ceil(H/2),ceil(W/2),C].  argmax_out (optional) stores the 2-bit window position
 * (dy*2+dx) per output element instead of TF's flat int64 index. */
int sn_maxpool2_moments_fwd(int32_t batch, int32_t h, int32_t w, int32_t c, const float* mu, const float* var,
                            float* mu_out, float* var_out, uint8_t* argmax_out, sn_stream_t st);
/* Routes both gradients to the stored arg-max position; g_*_in ([B,H,W,C]) are fully overwritten. */
int sn_maxpool2_moments_bwd(int32_t batch, int32_t h, int32_t w, int32_t c, const uint8_t* argmax,
                            const float* g_mu_out, const float* g_var_out, float* g_mu_in, float* g_var_in,
                            sn_stream_t st);

/* Strided window copy:
 *   dst[b, dst_y0 + dst_step*y, dst_x0 + dst_step*x, dst_c0 + ch] = src[b, src_y0 + src_step*y, src_x0 + src_step*x, src_c0 + ch]
 * for y < h, x < w, ch < c.  One primitive behind unpool (Brats.py:178-203: dst_step 2, dst offset 1), mypadding
 * (Brats.py:159-163, after sn_fill), crop_tensor + concat (Brats_functions.py:518-526, Brats.py:257-260) and all
 * of their adjoints (the adjoint of unpool reads with src_step 2). */
typedef struct sn_window {
  int32_t batch, h, w, c;                                  /* window extent */
  int32_t src_h, src_w, src_c, src_y0, src_x0, src_c0;      /* full source dims and window origin */
  int32_t dst_h, dst_w, dst_c, dst_y0, dst_x0, dst_c0;      /* full destination dims and window origin */
  int32_t dst_step;                                        /* 1, or 2 for zero-stuffing */
  int32_t src_step;                                        /* 1 (0 is read as 1), or 2 for the adjoint of zero-stuffing */
} sn_window;
int sn_window_copy(const sn_window* win, const float* src, float* dst, sn_stream_t st);
int sn_fill(float* dst, size_t n, float value, sn_stream_t st);

/* mysoftmax.call (Brats.py:269-283) on [rows, C] (rows = B*H*W, C <= 8):
 *   p = softmax(mu);  var_out_i = sum_j (p_i (delta_ij - p_j))^2 var_j. */
int sn_softmax_moments_fwd(size_t rows, int32_t c, const float* mu, const float* var, float* p_out,
                           float* var_out, sn_stream_t st);
/* VJP of both outputs w.r.t. both inputs. */
int sn_softmax_moments_bwd(size_t rows, int32_t c, const float* p, const float* var_in, const float* g_p,
                           const float* g_var_out, float* g_mu, float* g_var_in, sn_stream_t st);

/* nll_gaussian (Brats.py:293-311) applied to clip(var, clip_lo, clip_hi) (Brats.py:573-574: [1e-12,1e3];
 * Brats.py:588-589: [-1e4,1e3]).  acc: 2 doubles of caller workspace (zeroed by the call); loss_out: 1 float:
 *   0.5 * ( finite_or_0( mean_rows sum_c (p-y)^2/(v+1e-3) ) + mean_rows log prod_c (v+1e-3) ). */
int sn_nll_gaussian_fwd(size_t rows, int32_t c, const float* y, const float* p, const float* var, float clip_lo,
                        float clip_hi, double* acc, float* loss_out, sn_stream_t st);
/* Gradients w.r.t. p and the UNclipped var (zero outside [clip_lo, clip_hi]); g_loss is a device scalar;
 * acc is the workspace filled by the forward (needed for the NaN/Inf -> 0 rule). */
int sn_nll_gaussian_bwd(size_t rows, int32_t c, const float* y, const float* p, const float* var, float clip_lo,
                        float clip_hi, const double* acc, const float* g_loss, float* g_p, float* g_var,
                        sn_stream_t st);

/* One conv's contribution to add_n(model.losses) (Brats.py:575): l2(1.)(w_mu) (Brats.py:56,109) plus
 * sigma_regularizer(k*k)(w_sigma) (Brats.py:314-320).  Accumulates into *acc (double, caller-zeroed). */
int sn_kl_regularizer_fwd(const float* w_mu, size_t n_w, const float* w_sigma, int32_t cout, int32_t ksize,
                          double* acc, sn_stream_t st);
/* g_w_mu += scale * 2 w_mu ;  g_w_sigma[n] += scale * (-k^2/cout) (1/s_n - 1) sigmoid(w_sigma[n]). */
int sn_kl_regularizer_bwd(const float* w_mu, size_t n_w, const float* w_sigma, int32_t cout, int32_t ksize,
                          float scale, float* g_w_mu, float* g_w_sigma, sn_stream_t st);

/* ------------------------------------------------------------------------------------------------
 * FAST mode (inference): tcgen05 / TMEM implicit GEMM fed by TMA im2col tiles; the mean is carried as a
 * bf16 hi/lo pair (3 UMMAs: hi*Whi + lo*Whi + hi*Wlo, ~2^-16 relative), the variance as one bf16
 * (1 UMMA against bf16(W^2)), fp32 accumulation in TMEM, variance clamped >= 0 in the epilogue.
 *
 * "packed" moment tensor: [n][h][w][3][c] bf16 -- per pixel the three planes mean_hi, mean_lo, variance,
 * each c channels (c % 32 == 0 for tensors the tensor-core conv reads).  A sn_packed_view addresses a
 * window of such a buffer, which is how mypadding (Brats.py:159-163), crop_tensor + concat
 * (Brats_functions.py:518-526, Brats.py:257-260) and unpool's scatter (Brats.py:178-203) are fused away:
 * producers write straight into the interior / channel slice / strided positions of the consumer's buffer.
 * ------------------------------------------------------------------------------------------------ */
typedef struct sn_packed_view {
  void* base;           /* plane 0, element (0,0,0,0) of the FULL buffer; 16-byte aligned          */
  int32_t n, h, w, c;   /* FULL buffer dims                                                        */
  int32_t y0, x0, c0;   /* window origin inside the buffer (crop / pad interior / concat slice)    */
  int32_t reserved;
} sn_packed_view;

size_t sn_packed_bytes(int32_t n, int32_t h, int32_t w, int32_t c);

/* fp32 [pixels][c] mean/variance <-> packed [pixels][3][c] (var == NULL packs a zero variance). */
int sn_pack_moments(size_t pixels, int32_t c, const float* mu, const float* var, void* packed, sn_stream_t st);
int sn_unpack_moments(size_t pixels, int32_t c, const void* packed, float* mu, float* var, sn_stream_t st);
/* Fill a whole packed buffer with (mean 0, variance var_fill): the mypadding border (Brats.py:159-163,370-372). */
int sn_packed_fill(void* packed, size_t pixels, int32_t c, float var_fill, sn_stream_t st);

/* Once per weight update: HWIO fp32 w_mu -> tensor-core operands [3][taps][cout][cin] bf16 (plane 0 = W_hi,
 * 1 = W_lo, 2 = bf16(W^2)) and s_out[n] = softplus(w_sigma[n]) (Brats.py:67,120).  upconv != 0 (requires
 * ksize == 2): the four "taps" are the output parities (a,b) of the stride-2 transposed convolution that unpool
 * (Brats.py:178-203) followed by the 2x2 VALID conv (Brats.py:349,414-415) amounts to: tap 2a+b holds W[1-a,1-b]. */
size_t sn_prepared_weight_bytes(int32_t ksize, int32_t cin, int32_t cout);
int sn_prepare_weights(const float* w_mu, const float* w_sigma, int32_t ksize, int32_t cin, int32_t cout,
                       int32_t upconv, void* w_packed, float* s_out, sn_stream_t st);

enum { SN_TC_RELU = 1, SN_TC_UPCONV = 2, SN_TC_DST_F32 = 4, SN_TC_IM2COL = 8, SN_TC_ROWS = 16, SN_TC_EXACT = 32,
       SN_TC_KWC = 64, SN_TC_NO_KWC = 128, SN_TC_CTA2 = 256, SN_TC_NO_CTA2 = 512 };

/* One fused moment convolution on the tensor cores: myConv_intermediate.call (Brats.py:118-137), optionally
 * with the ReLU gate of Brats.py:233-238 (SN_TC_RELU), reading the channel-concat of up to two packed windows
 * (src[0] = decoder first, src[1] = cropped encoder: myConc, Brats.py:247-261) and writing a packed window.
 *   - ksize in {1,2,3}, stride 1, VALID over the in_h x in_w window; src_c[i] % 32 == 0, cout % 32 == 0;
 *   - SN_TC_UPCONV (ksize must be 2, weights prepared with upconv = 1): myupsampling + the 2x2 conv
 *     (Brats.py:414-415) as four parity GEMMs; output pixel (2y+a, 2x+b) of a 2*in_h x 2*in_w window;
 *   - SN_TC_DST_F32: write fp32 NHWC mean/variance (dst_mu, dst_var: contiguous [batch,out_h,out_w,cout])
 *     instead of the packed window dst;
 *   - SN_TC_IM2COL: run the first-generation kernel (one CTA per 128-pixel tile, TMA im2col loads per tap)
 *   - SN_TC_KWC / SN_TC_NO_KWC: force / forbid the kw-concatenated variant of the halo kernel (32 output columns per
 *     tile, k = 3, width >= 32: the three taps of a filter row are N columns of one UMMA and are summed with a lane
 *     shift in the epilogue; packed rows leave through TMA stores).  Default: the library's choice (layers with >= 2
 *     input channel blocks).  Same results to rounding either way; an A/B and test switch like SN_TC_IM2COL.
 *   - SN_TC_CTA2 / SN_TC_NO_CTA2: force / forbid the CTA-pair variant (128 output columns per tile with streamed weights;
 *     64 output columns per tile, k = 3, with half of every weight slot RESIDENT per SM): clusters of two CTAs issue
 *     cta_group::2 UMMAs of M = 256 over two pixel tiles and each SM holds half of every weight slot.  Default: the
 *     library's choice (enough tile pairs to fill the SM pairs; 64 columns: layers whose weights do not fit one SM).
 *     Bit-identical results.
 *     instead of the default persistent halo-tiled kernel; same results, kept for A/B measurements. */
typedef struct sn_tc_conv_desc {
  sn_packed_view src[2];
  int32_t src_c[2];       /* channels taken from each source (src_c[1] == 0: single source) */
  int32_t batch, in_h, in_w, ksize, cout, flags;
  const void* w_packed;   /* from sn_prepare_weights */
  const float* s;         /* softplus(w_sigma) [cout], from sn_prepare_weights */
  sn_packed_view dst;
  float* dst_mu;
  float* dst_var;
  float* rsum_out;        /* optional fp32 [batch, Ho, Wo] (SN_TC_UPCONV: [batch, in_h, in_w]): the rank-1 statistic
                             box_k(sum_c mu^2 + var) per pixel, which sn_conv_moments_bwd_weight_tc needs; may be NULL */
} sn_tc_conv_desc;
int sn_conv_moments_fwd_tc(const sn_tc_conv_desc* d, sn_stream_t st);

/* The network's LAST 3x3 convolution fused with conv_final (k = 1, Brats.py:367,454) and mysoftmax (Brats.py:269-283):
 * Density_prop_with_pad_UNET.call ends `up4_conv2 -> ReLU -> conv_final -> softmax` (Brats.py:451-455,
 * Hippocampus.py:417-421), and the 32 channels of a pixel that the conv's epilogue has just produced are all conv_final
 * needs -- so the epilogue finishes the forward and sn_final_conv_softmax_packed is not launched.
 *   - d: as for sn_conv_moments_fwd_tc, restricted to what sn_tc_head_fusable() accepts (ksize 3, SN_TC_RELU, one packed
 *     32-channel source, cout 32, no SN_TC_UPCONV / DST_F32 / IM2COL / KWC, rsum_out NULL); d->dst.base may be NULL: the
 *     32-channel tensor is then never written (inference), otherwise it is stored as usual (32-byte aligned window);
 *   - h: conv_final's raw parameters (w_mu fp32 [32][n_labels], w_sigma [n_labels], 2 <= n_labels <= 5) and the outputs
 *     of sn_final_conv_softmax_packed: fp32 [batch*Ho*Wo, n_labels] probabilities / variances, optional pre-softmax
 *     moments (both or neither).
 * Results are bit-identical to sn_conv_moments_fwd_tc followed by sn_final_conv_softmax_packed: the head reads the
 * bf16-rounded (hi, lo, var) values the conv would have stored and runs the same per-pixel chain in channel order. */
typedef struct sn_tc_head_desc {
  int32_t n_labels;
  const float* w_mu;
  const float* w_sigma;
  float* p_out;
  float* var_out;
  float* presoftmax_mu;
  float* presoftmax_var;
} sn_tc_head_desc;
int sn_tc_head_fusable(const sn_tc_conv_desc* d, int32_t n_labels);   /* 1 / 0; no launch, no error state */
int sn_conv_moments_fwd_tc_head(const sn_tc_conv_desc* d, const sn_tc_head_desc* h, sn_stream_t st);

/* myConv_input.call (Brats.py:65-76) (+ ReLU with SN_TC_RELU) on fp32 NHWC x (cin <= 8), written as a packed window.
 * k = 3, cout = 32, cin in {1, 4} runs on the tensor cores (bf16 hi/lo image and weights, ~1e-5 relative); with
 * SN_TC_EXACT the fp32 CUDA-core kernel is used instead: the gradient engine wants this layer's ReLU gates -- the
 * largest tensor of the network -- decided at fp32 accuracy, because every flipped gate is a 100 % error of that
 * element's gradient.  SN_TC_ROWS selects the warp-specialised form of the tensor-core kernel (builder / UMMA / epilogue
 * roles of one persistent CTA, TMA-store epilogue), SN_TC_IM2COL forces the default one-role-per-CTA form: bit-identical
 * results and the same measured time; an A/B switch. */
int sn_first_conv_fwd_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout, int32_t ksize,
                             const float* x, const float* w_mu, const float* w_sigma, const sn_packed_view* dst,
                             int32_t flags, sn_stream_t st);
/* mymaxpooling (Brats.py:171-174) on an in_h x in_w x c packed window -> packed window (2x2/2, SAME). */
int sn_maxpool2_packed(const sn_packed_view* src, int32_t batch, int32_t in_h, int32_t in_w, int32_t c,
                       const sn_packed_view* dst, sn_stream_t st);
/* myReLU.call (Brats.py:233-238) on a packed window: mean = hi + lo; where mean > 0 (strict, TF ReluGrad) the three
 * planes are copied, elsewhere all three are zero.  gate = 0: plain window-to-window copy (mypadding / crop of a tensor
 * that is already in memory).  src and dst windows are h x w x c (c % 8 == 0); in place is allowed when they coincide.
 * The engines never call this (the gate is a conv-epilogue flag); the layer-by-layer FAST API does. */
int sn_relu_packed(const sn_packed_view* src, int32_t batch, int32_t h, int32_t w, int32_t c,
                   const sn_packed_view* dst, int32_t gate, sn_stream_t stream);

/* conv_final (k = 1, Brats.py:367,454) fused with mysoftmax (Brats.py:269-283): packed window in (cin == 32),
 * fp32 [batch*in_h*in_w, n_labels] probabilities and variances out; presoftmax_{mu,var} optional (may be NULL). */
int sn_final_conv_softmax_packed(const sn_packed_view* src, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                 int32_t n_labels, const float* w_mu, const float* w_sigma, float* p_out,
                                 float* var_out, float* presoftmax_mu, float* presoftmax_var, sn_stream_t st);

/* ------------------------------------------------------------------------------------------------
 * FAST mode backward: the input-gradient chain tf.GradientTape derives for create_adversarial_pattern
 * (Brats.py:582-596) and, layer by layer, for train_on_batch (Brats.py:569-580); formulas in SURVEY.md A.3.
 * A gradient tensor uses the packed layout of the tensor it belongs to: planes g_mean_hi, g_mean_lo
 * (g_mean = hi + lo) and g_variance, bf16.  Every kernel recomputes gates / arg-max routing from the SAVED
 * forward activations (the packed buffers the forward wrote), so no masks or indices are stored.
 * ------------------------------------------------------------------------------------------------ */

/* Once per weight update: HWIO fp32 w_mu -> the data-gradient operands [3][taps][cin][K] bf16 (W_hi, W_lo, bf16(W^2)):
 * regular conv: taps = k*k, K = cout, tap (kh',kw') holds W[k-1-kh', k-1-kw', ci, n] (the flipped, transposed filter
 * of the full correlation g (*)^T W); upconv != 0 (ksize == 2): taps = 1, K = 4*cout ordered (parity 2a+b, n),
 * holding W[1-a, 1-b, ci, n] -- the adjoint of the four parity GEMMs of sn_conv_moments_fwd_tc(SN_TC_UPCONV).
 * Same byte size as sn_prepared_weight_bytes(). */
int sn_prepare_weights_bwd(const float* w_mu, int32_t ksize, int32_t cin, int32_t cout, int32_t upconv,
                           void* wt_packed, sn_stream_t st);

/* Data gradient of sn_conv_moments_fwd_tc on the tensor cores (same persistent halo kernel, gradient as the A operand):
 *   g_mu_in = g_mu_out (*)^T W + 2 mu_in . box^T(t),  g_var_in = g_var_out (*)^T W^2 + box^T(t),  t = sum_n g_var_out s_n,
 * followed, per forward source i with gate[i] != 0, by the ReLU gate of the layer that produced that source
 * (g = 0 where the saved mean is not > 0; Brats.py:233-238 -- the gate is a constant for autodiff).
 *   g_out : window (out_h x out_w x cout) of the packed gradient w.r.t. the conv output (SN_TC_UPCONV: 2in_h x 2in_w);
 *   in[i] : the forward's source windows (saved activations), in_c[i] channels each (in_c[1] == 0: single source);
 *   g_in[i]: destination windows (in_h x in_w x in_c[i]), fully overwritten.
 * The adjoints of mypadding / crop_tensor / myConc / unpool are address arithmetic: pass the matching windows. */
typedef struct sn_tc_dgrad_desc {
  sn_packed_view g_out;
  sn_packed_view in[2];
  sn_packed_view g_in[2];
  int32_t in_c[2];
  int32_t gate[2];
  int32_t batch, in_h, in_w, ksize, cout, flags; /* forward geometry; flags: SN_TC_UPCONV only */
  const void* wt_packed;                         /* from sn_prepare_weights_bwd */
  const float* s;                                /* softplus(w_sigma) [cout] */
} sn_tc_dgrad_desc;
int sn_conv_moments_bwd_data_tc(const sn_tc_dgrad_desc* d, sn_stream_t st);

/* Weight gradient of sn_conv_moments_fwd_tc on the tensor cores (train_on_batch, Brats.py:569-580):
 *   g_w_mu = corr(mu_in, g_mu_out) + 2 W . corr(var_in, g_var_out);  g_w_sigma[n] = sigmoid(w_sigma[n]) sum_p g_var_out[p,n] rsum[p].
 * Two pixel-axis GEMMs (tcgen05, MN-major operands straight from the NHWC planes, split-K with fp32 atomics into
 * `workspace`, sn_wgrad_workspace_bytes() bytes, zeroed by the call), then a finalize pass.  Outputs are overwritten.
 * in[]/in_c[]/g_out/flags as in sn_tc_dgrad_desc; rsum = the forward's rsum_out; w_mu HWIO fp32, w_sigma raw.
 * k = 3 layers run the row-halo kernel (one tiled TMA band of full-width rows, taps as K-row shifts of the MN-major
 * tile) unless both channel counts are >= 256 (few pixels: the general kernel is faster); SN_TC_IM2COL forces the
 * general kernel (TMA im2col per tap) that k = 1 / 2 and the up-conv always use, SN_TC_ROWS forces the row-halo one. */
typedef struct sn_tc_wgrad_desc {
  sn_packed_view g_out;
  sn_packed_view in[2];
  int32_t in_c[2];
  int32_t batch, in_h, in_w, ksize, cout, flags;
  const float* rsum;
  const float* w_mu;
  const float* w_sigma;
  void* workspace;
  float* g_w_mu;
  float* g_w_sigma;
} sn_tc_wgrad_desc;
size_t sn_wgrad_workspace_bytes(int32_t ksize, int32_t cin, int32_t cout);
int sn_conv_moments_bwd_weight_tc(const sn_tc_wgrad_desc* d, sn_stream_t st);
/* rsum[b,y,x] = sum over the k x k window and the channels of x^2: myConv_input's rank-1 statistic (Brats.py:69-73),
 * the `rsum` argument of sn_conv_moments_bwd_weight for the first layer. */
int sn_first_conv_rsum(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t ksize, const float* x,
                       float* rsum, sn_stream_t st);

/* Weight gradients of the two layers too thin for the tensor cores, on CUDA cores (same formulas):
 *   myConv_input (k = 3, cout = 32, cin 1 or 4): x fp32 NHWC, g_out = packed gradient w.r.t. its (pre-ReLU) output;
 *   conv_final (k = 1, cin = 32): in = its packed input, g_logit_* / rsum = the optional outputs of sn_head_bwd_packed.
 * workspace: at least sn_wgrad_workspace_bytes(ksize, cin, cout) bytes.  Outputs are overwritten. */
int sn_first_conv_bwd_weight_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout,
                                    int32_t ksize, const float* x, const float* w_sigma, const sn_packed_view* g_out,
                                    void* workspace, float* g_w_mu, float* g_w_sigma, sn_stream_t st);
int sn_final_conv_bwd_weight_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                    int32_t n_labels, const float* w_mu, const float* w_sigma, const float* g_logit_mu,
                                    const float* g_logit_var, const float* rsum, void* workspace, float* g_w_mu,
                                    float* g_w_sigma, sn_stream_t st);

/* Adjoint of sn_maxpool2_packed (Brats.py:171-174,206-216): routes g_out (ceil(in_h/2) x ceil(in_w/2) x c) to the
 * arg-max position recomputed from the saved pool input `in` (first maximum in row-major window order).  g_in
 * (in_h x in_w x c window) is overwritten, EXCEPT inside the sub-window (keep_y0, keep_x0, keep_h, keep_w), where
 * the routed gradient is added to what is already there: that is how the skip connection's second consumer
 * (the cropped encoder half of myConc, Brats.py:247-261) is summed in.  keep_h == 0: plain overwrite. */
int sn_maxpool2_bwd_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t c,
                           const sn_packed_view* g_out, const sn_packed_view* g_in, int32_t keep_y0, int32_t keep_x0,
                           int32_t keep_h, int32_t keep_w, sn_stream_t st);

/* Backward of the fused head: loss_scale * nll_gaussian(y, p, clip(var)) (Brats.py:293-311; clips :573-574 / :588-589)
 * -> mysoftmax (Brats.py:269-283) -> conv_final (k = 1) -> ReLU gate of the tensor feeding conv_final, in one pass.
 * The forward quantities are recomputed from the saved packed input `in`; `acc` is the workspace sn_nll_gaussian_fwd
 * filled for this batch (NaN/Inf -> 0 rule of Brats.py:304-305).  Writes the packed gradient window g_in and,
 * when the three optional fp32 outputs are given (all or none), what the weight gradient of conv_final needs:
 * the gradients w.r.t. the pre-softmax mean / variance [rows, n_labels] and rsum [rows] = sum_c mu^2 + var. */
int sn_head_bwd_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                       int32_t n_labels, const float* w_mu, const float* w_sigma, const float* y, float clip_lo,
                       float clip_hi, const double* acc, float loss_scale, const sn_packed_view* g_in,
                       float* g_logit_mu, float* g_logit_var, float* rsum_out, sn_stream_t st);

/* Same chain for GIVEN upstream gradients w.r.t. the two outputs of mysoftmax (g_var_out may be NULL = 0): what
 * create_saliency_map needs (Brats.py:598-609: d sum(masked p) / d x, no loss involved). */
int sn_head_bwd_upstream_packed(const sn_packed_view* in, int32_t batch, int32_t in_h, int32_t in_w, int32_t cin,
                                int32_t n_labels, const float* w_mu, const float* w_sigma, const float* g_p,
                                const float* g_var_out, const sn_packed_view* g_in, sn_stream_t st);

/* Input gradient of myConv_input (Brats.py:65-76; what create_adversarial_pattern returns the sign of):
 *   g_x = g_mu_out (*)^T W + 2 x . box^T(t),  t = sum_n g_var_out s_n;  g_out: packed (in_h-k+1 x in_w-k+1 x cout)
 * window (already gated by the first ReLU), g_x: fp32 NHWC [batch,in_h,in_w,cin], cin <= 8. */
int sn_first_conv_bwd_data_packed(int32_t batch, int32_t in_h, int32_t in_w, int32_t cin, int32_t cout, int32_t ksize,
                                  const float* x, const float* w_mu, const float* w_sigma,
                                  const sn_packed_view* g_out, float* g_x, sn_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* SUPERNET_B200_H_ */
