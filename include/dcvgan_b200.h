/*
 * dcvgan_b200.h — C ABI of libdcvgan_b200.so
 *
 * Drop-in boundary for the DCVGAN training step (reference: raahii/dcvgan).
 * The reference has no FFI of its own: its hot path is the PyTorch module surface
 * src/generator.py, src/discriminator.py, src/loss.py, src/trainer.py:271-363 and the
 * torch.optim.Adam steps built at src/train.py:167-176.  Every entry point below names the
 * reference call site whose arithmetic it replaces.  The Python host side
 * (dcvgan_b200/*.py) binds these through ctypes; see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless marked host
 *   - the library never allocates or frees device memory: the caller owns every buffer
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *   - return 0 on success, negative on error; dcv_last_error() gives the text
 *   - activations are channels-last: (N, [T,] H, W, C) with an explicit pixel stride `ld`
 *     (elements between consecutive pixels) so a tensor may be a channel slice of a wider
 *     concat buffer (replaces torch.cat at generator.py:114,393,398,400,
 *     discriminator.py:124,228)
 *   - dtype: DCV_F32 (exact mode, CUDA-core fp32) or DCV_BF16 (fast mode; tcgen05 where the
 *     shape allows, fp32 accumulate everywhere)
 */
#ifndef DCVGAN_B200_H
#define DCVGAN_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCV_ABI_VERSION 5

enum { DCV_F32 = 0, DCV_BF16 = 1 };
/* activation selectors (generator.py:63,76,78,175,206,243,276; discriminator.py:82-99) */
enum { DCV_ACT_NONE = 0, DCV_ACT_LEAKY = 1, DCV_ACT_TANH = 2 };
/* implementation selector for the convolution entry points */
/* SIMT: CUDA-core fp32 kernels (exact mode, any shape); TC: tcgen05 kind::f16 on bf16 tensors; TC_TF32: tcgen05 kind::tf32 on
 * fp32 tensors (forward / data gradient of layers whose channel counts are multiples of 32) */
enum { DCV_IMPL_SIMT = 0, DCV_IMPL_TC = 1, DCV_IMPL_TC_TF32 = 2 };
/* direction of a correlation w.r.t. the geometry below */
enum { DCV_DIR_GATHER = 0, DCV_DIR_SCATTER = 1 };
/* loss kinds (loss.py:93-99,123-131,163-166,190-193) */
enum {
  DCV_LOSS_BCE_ONES = 0,   /* mean BCEWithLogits(y, 1) = mean softplus(-y)      */
  DCV_LOSS_BCE_ZEROS = 1,  /* mean BCEWithLogits(y, 0) = mean softplus(+y)      */
  DCV_LOSS_HINGE_REAL = 2, /* mean relu(1 - y)                                  */
  DCV_LOSS_HINGE_FAKE = 3, /* mean relu(1 + y)                                  */
  DCV_LOSS_SOFTPLUS_NEG = 4/* mean softplus(-y)  (HingeLoss.compute_gen_loss)   */
};

/*
 * Geometry of one strided correlation between a "large" side L and a "small" side S:
 *   S[n, ts, hs, ws]  <->  L[n, ts*st - pt + a, hs*sh - ph + b, ws*sw - pw + c]   for taps (a,b,c)
 * nn.Conv2d / nn.Conv3d forward reads L and writes S (DCV_DIR_GATHER); their data gradient is
 * DCV_DIR_SCATTER.  nn.ConvTranspose2d forward reads S and writes L (DCV_DIR_SCATTER); its data
 * gradient is DCV_DIR_GATHER.  2-D layers use T=1, kt=1, st=1, pt=0.
 */
typedef struct dcv_geom {
  int32_t N;
  int32_t Tl, Hl, Wl, Cl;
  int32_t Ts, Hs, Ws, Cs;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
  /* Channels present in the master weight (0 = same as Cl / Cs).  Cl / Cs may be larger: the activation then
   * carries zero-filled padding channels (up to a multiple of 16) so that tiny channel counts (1, 2, 3, 25, 50,
   * 266 ...) still run on the tensor cores.  Packing zero-fills the padding rows; wgrad never writes them. */
  int32_t wCl, wCs;
} dcv_geom;

int dcv_abi_version(void);
const char* dcv_last_error(void);
/* 1 if the device behind the current context is sm_100 (B200); the host side refuses to run otherwise */
int dcv_device_ok(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
long long dcv_launch_count(void);
/* Result-preserving tuning switches of the tcgen05 kernels (process-wide; never read from the environment):
 * "nohalo", "mt" (0 = automatic, 1/2/4 forced M tiles per work item), "no_tma_store", "no_narrow_tma_store",
 * "wgrad_waves", "no_gemv", "no_tapgroup", "no_fused_stats", "sm_reserve" (SMs the persistent kernels leave free,
 * for a collective running beside them), "pdl" (1 = launch with the programmatic-dependent-launch attribute; measured slower,
 * off by default), "no_nsplit" (1 = keep the widest N tile even when a convolution has fewer work items than half the SMs),
 * weight-gradient plans: "wgrad_g" (accumulator tiles per CTA: 0 = default 1; 2 / 4 = the older plans), "wgrad_ns" (S-tile
 * columns, 0 = 256), "no_wgrad_halo" (one TMA box per tap instead of a row-halo box per tap pair), "no_wgrad_whalo" (no
 * column halo in the 2- / 4-tile plans), "wgrad_halo4" (row halo for 4-pixel-wide maps as well).  The parity tests flip
 * them to cover every kernel path. */
int dcv_set_tuning(const char* key, int value);

/* ---- weight packing -------------------------------------------------------------------------
 * Re-lays a PyTorch fp32 master weight for one direction of a layer.  `w` is addressed as
 * w[cl*s_l + cs*s_s + tap*s_tap] (tap = (a*kh + b)*kw + c), which covers both
 * nn.Conv*d (Cout,Cin,k..) and nn.ConvTranspose2d (Cin,Cout,kh,kw) layouts.
 *   impl SIMT : out is fp32 [tap][K][Nout]           (K = reduction channels, Nout = produced channels)
 *   impl TC   : out is bf16, GATHER : [Nout_pad][tap][K]
 *                            SCATTER: [phase][Nout_pad][tap_in_phase][K]   (sub-pixel phases)
 * dcv_packed_weight_bytes returns the size the caller must provide.
 */
int64_t dcv_packed_weight_bytes(const dcv_geom* g, int dir, int impl);
int dcv_pack_weight(const dcv_geom* g, int dir, int impl, const float* w, int64_t s_l, int64_t s_s,
                    int64_t s_tap, void* out, void* stream);

/* n tcgen05 packs (whole weights, DCV_IMPL_TC) in one kernel launch: all layers of a network after its optimizer step
 * (train.py:167-176 builds one Adam per network).  Arrays of length n; outs[i] has dcv_packed_weight_bytes(geoms[i], dirs[i], TC) bytes. */
int dcv_pack_weight_batch(int n, const dcv_geom* const* geoms, const int* dirs, const float* const* w, const int64_t* s_l,
                          const int64_t* s_s, const int64_t* s_tap, void* const* outs, void* stream);

/* Same, for ONE master weight that occupies only a window of the packed matrix - activation channels
 * [cl_off, cl_off+cl_cnt) x [cs_off, cs_off+cs_cnt) - as when the two stem convolutions of a discriminator
 * (conv_g on the geometry channels, conv_c on the colour channels, discriminator.py:79-90,180-193) are run as a single
 * convolution over the concatenated input [xg | xc] producing [hc | hg].  `w` is indexed with window-local channels.
 * fill_outside != 0 zero-fills everything outside the window (first call), 0 leaves it untouched (further calls). */
int dcv_pack_weight_sub(const dcv_geom* g, int dir, int impl, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                        int cl_off, int cl_cnt, int cs_off, int cs_cnt, int fill_outside, void* out, void* stream);

/* tcgen05 packing of a matrix assembled from n (<= 16) master weights with disjoint windows, zeros elsewhere, one launch */
int dcv_pack_weight_multi(const dcv_geom* g, int dir, int n, const float* const* w, const int64_t* s_l, const int64_t* s_s,
                          const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off, const int* cs_cnt,
                          void* out, void* stream);

/* ---- convolution family ---------------------------------------------------------------------
 * dcv_conv: y = act(correlate(x, wp)); replaces every nn.Conv2d / nn.ConvTranspose2d / nn.Conv3d
 * forward on the path (generator.py:61-73,174,204,240,274; discriminator.py:81-101,182-206,288-305)
 * and, with the opposite `dir`, the data-gradient half of their autograd backward.
 * dcv_conv_tc_supported tells whether the tcgen05 kernel covers (geom, dir).
 */
int dcv_conv_tc_supported(const dcv_geom* g, int dir);
/* Discriminator head (Conv -> 1 channel, discriminator.py:101,206,305) with its adversarial-loss term fused in (loss.py:93-99,
 * 123-131,163-166,190-193): y = correlate(x, wp) as dcv_conv does for a single output channel (bf16, packed DCV_IMPL_TC weight),
 * then - in the same launch, by the block that finishes last - loss_out[0] (+)= mean loss(kind) over all logits and, if dy is
 * not NULL, dy = grad_scale * dL/dlogit.  Same numbers as dcv_conv followed by dcv_loss_fwd_bwd.  counter: one zero-initialised
 * uint32 the kernel leaves zero again (dcv_bn_tail_counters() buffers qualify). */
int dcv_head_loss_supported(const dcv_geom* g, int dir);
int dcv_head_loss(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int kind,
                  float* loss_out, int accumulate, void* dy, int64_t lddy, float grad_scale, void* counter, void* stream);
/* same question for DCV_IMPL_TC_TF32 (fp32 activations and packed weights, tf32 products, fp32 accumulation and output) */
int dcv_conv_tf32_supported(const dcv_geom* g, int dir);
int dcv_conv(const dcv_geom* g, int dir, int impl, int dtype, const void* x, int64_t ldx,
             const void* wp, void* y, int64_t ldy, int act, float slope, void* stream);

/* Convolution with the BatchNorm batch statistics fused into its epilogue (Conv/ConvTranspose + BatchNorm pairs:
 * generator.py:62-71,205-211,242-248; discriminator.py:94-99,197-204,289-302).  dcv_conv_stats_slots returns the number
 * of fp32 [2][Cout_padded] partial-sum slots (sum, sum of squares per output channel of the bf16-rounded outputs) that
 * dcv_conv_stats writes for this geometry and these pixel strides - 0 if the statistics cannot be fused (output channels
 * not a multiple of 64, CUDA-core path, ...), in which case the caller runs dcv_conv + dcv_bn_stats.  The slots have the
 * layout dcv_bn_finalize expects (nblk = slots, C = Cout_padded).  tcgen05 / bf16 only. */
/* y += correlate(x, wp) for DCV_IMPL_TC / bf16: the epilogue adds into the output with TMA reduce-add instead of overwriting it
 * (gradient fan-in of the U-Net skip connections, generator.py:398-400: the data gradient of a down block lands directly on top
 * of the skip gradient, no separate a + b pass).  Needs a 16-byte aligned output whose channel count is a multiple of 16. */
int dcv_conv_accumulate(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, void* stream);
int dcv_conv_stats_slots(const dcv_geom* g, int dir, int64_t ldx, int64_t ldy);
int dcv_conv_stats(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int act,
                   float slope, float* stats, int slots, void* stream);

/* Weight gradient: dw[cl*s_l + cs*s_s + tap*s_tap] (+)= sum_m S[m,cs] * L[gather(m,tap),cl].
 * `ws` is scratch of dcv_wgrad_workspace_bytes(); split partial sums are reduced deterministically. */
int64_t dcv_wgrad_workspace_bytes(const dcv_geom* g, int impl);
int dcv_wgrad_tc_supported(const dcv_geom* g);
int dcv_wgrad(const dcv_geom* g, int impl, int dtype, const void* xl, int64_t ldl, const void* xs,
              int64_t lds, float* dw, int64_t s_l, int64_t s_s, int64_t s_tap, int accumulate,
              void* ws, int64_t ws_bytes, void* stream);

/* Two-step form of dcv_wgrad for merged layers: dcv_wgrad_partial leaves the split partial sums in `ws`;
 * dcv_wgrad_reduce_sub reduces the window of one master weight out of them (call once per master weight). */
int dcv_wgrad_partial(const dcv_geom* g, int impl, int dtype, const void* xl, int64_t ldl, const void* xs, int64_t lds,
                      void* ws, int64_t ws_bytes, void* stream);
int dcv_wgrad_reduce_sub(const dcv_geom* g, int impl, const void* ws, float* dw, int64_t s_l, int64_t s_s, int64_t s_tap,
                         int cl_off, int cl_cnt, int cs_off, int cs_cnt, int accumulate, void* stream);
/* n windows (<= 16, disjoint) in one launch: every partial sum is read once and routed to the weight that owns it */
int dcv_wgrad_reduce_multi(const dcv_geom* g, int impl, const void* ws, int n, float* const* dw, const int64_t* s_l,
                           const int64_t* s_s, const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off,
                           const int* cs_cnt, const int* accumulate, void* stream);

/* ---- BatchNorm (training + eval), activation, dropout, noise --------------------------------
 * Replaces nn.BatchNorm2d/3d (+ReLU/LeakyReLU, +Dropout2d between them, +Noise before the next
 * conv): generator.py:62-71,205-211,242-248; discriminator.py:30-39,94-99,197-204,289-302.
 * rows = N*T*H*W pixels.  partials is fp32 [nblk][2][C]; nblk = dcv_bn_stats_blocks(rows, C).
 */
int dcv_bn_stats_blocks(int64_t rows, int C);
int dcv_bn_stats(int dtype, const void* z, int64_t ldz, int64_t rows, int C, float* partials, void* stream);
/* mean/invstd from partials (biased var), running stats updated with unbiased var, momentum;
 * *num_batches_tracked += 1 on the device (nn.BatchNorm's counter) when the pointer is not NULL. */
int dcv_bn_finalize(const float* partials, int nblk, int C, int64_t count, float eps, float momentum,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean,
                    float* invstd, void* stream);
/* dcv_bn_stats + dcv_bn_finalize in ONE launch: the reduction blocks finish the job themselves ("last block done",
 * hierarchical, fixed summation order -> deterministic).  ws: dcv_bn_tail_workspace_bytes(rows, C) bytes of scratch;
 * counters: dcv_bn_tail_counters() zero-initialised uint32 that the kernel leaves zero again (one buffer serves every
 * call on a stream).  Same outputs as the two-launch form up to the rounding of a different summation tree. */
int64_t dcv_bn_tail_workspace_bytes(int64_t rows, int C);
int dcv_bn_tail_counters(void);
int dcv_bn_stats_finalize(int dtype, const void* z, int64_t ldz, int64_t rows, int C, float eps, float momentum,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                          void* ws, void* counters, void* stream);
/* eval mode: mean = running_mean, invstd = rsqrt(running_var + eps) */
int dcv_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps,
                      float* mean, float* invstd, void* stream);
/* a = act( ((z-mean)*invstd*gamma+beta) * drop[n,c] ) ; any of mean/gamma/drop may be NULL (identity).
 * rows_per_n = pixels per sample (dropout scale index n = row / rows_per_n). */
int dcv_bn_act(int dtype, const void* z, int64_t ldz, int64_t rows, int C, const float* mean,
               const float* invstd, const float* gamma, const float* beta, const float* drop,
               int64_t rows_per_n, int act, float slope, void* a, int64_t lda, void* stream);
/* backward: pass 1 reduces sum(du), sum(du*xhat) into partials [nblk][2][C]; pass 2 writes dz and,
 * from the reduced sums, dgamma/dbeta (fp32, (+)= if accumulate). du = da*act'(a)*drop.
 * For (Leaky)ReLU the vectorised kernels take the sign of the pre-activation gamma*xhat+beta recomputed from z and
 * do not read `a` (it must still be a valid pointer for the other activations and the scalar fallback kernels). */
int dcv_bn_act_bwd_reduce(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda,
                          const void* z, int64_t ldz, int64_t rows, int C, const float* mean,
                          const float* invstd, const float* gamma, const float* beta, const float* drop,
                          int64_t rows_per_n, int act, float slope, float* partials, void* stream);
/* dcv_bn_act_bwd_reduce + dcv_bn_bwd_finalize in ONE launch (same scheme as dcv_bn_stats_finalize) */
int dcv_bn_act_bwd_reduce_finalize(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda,
                                   const void* z, int64_t ldz, int64_t rows, int C, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, const float* drop,
                                   int64_t rows_per_n, int act, float slope, float* sums /*[2][C]*/, float* dgamma,
                                   float* dbeta, int accumulate, void* ws, void* counters, void* stream);
int dcv_bn_bwd_finalize(const float* partials, int nblk, int C, float* sums /*[2][C]*/, float* dgamma,
                        float* dbeta, int accumulate, void* stream);
int dcv_bn_act_bwd_apply(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda,
                         const void* z, int64_t ldz, int64_t rows, int C, const float* mean,
                         const float* invstd, const float* gamma, const float* beta, const float* drop,
                         int64_t rows_per_n, int act, float slope, const float* sums, int64_t count, void* dz,
                         int64_t lddz, void* stream);
/* plain activation backward (no BN): dz = da * act'(a) ; a is the activated output */
int dcv_act_bwd(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, int64_t rows, int C,
                int act, float slope, void* dz, int64_t lddz, void* stream);
/* Tap-unrolled form of a transposed convolution with <= 4 output channels (generator.py:73 ggen main.12 ConvTranspose2d(ngf, C,
 * 4, 2, 1) + Tanh; generator.py:274 outconv ConvTranspose2d(2*ngf, 3, 3, 1, 1) + Tanh): a 1x1 dcv_conv with the weight viewed
 * as (Cin) x (Cout*kh*kw) produces P[n][ih][iw][co*kh*kw + tap]; dcv_col2im_act sums the taps that land on each output pixel
 * and applies the activation.  2-D only (T == 1). */
int dcv_col2im_act(int dtype, const void* P, int64_t ldp, int N, int Ih, int Iw, int Cout, int KH, int KW, int stride, int pad,
                   int act, float slope, void* y, int64_t ldy, int Oh, int Ow, void* stream);
/* HBM-bound kernels (warp-level mma.sync, operands assembled in registers) for the 3x3 / stride 1 / pad 1 image-side layers
 * of the colour generator whose small side has <= 3 real channels: `Inconv` Conv2d(C, 64, 3, 1, 1) + LeakyReLU(0.01)
 * (generator.py:158-176) and `Outconv` ConvTranspose2d(128, 3, 3, 1, 1) (generator.py:256-282).  bf16 activations, fp32
 * MASTER weight addressed like dcv_pack_weight (w[cl*s_l + cs*s_s + tap], no packed copy).  In dcv_geom terms the small
 * tensor is L, the 64- or 128-channel tensor is S.
 * dcv_img_conv_supported(g, what): what = 0 forward-type pass L -> S (Inconv forward, Outconv data gradient),
 *   1 full Inconv backward (S has 64 channels, <= 2 real L channels), 2 weight gradient only, 3 scatter pass S -> L
 *   (Outconv forward).
 * dcv_img_conv_fwd:  y(S) = act(correlate(x(L), w)).
 * dcv_img_conv_scatter: y(L) = act(transposed correlation of xb(S) with w) - the forward of a ConvTranspose2d whose OUTPUT is
 *   the small tensor; the big tensor is read once, no intermediate.
 * dcv_img_conv_bwd:  one pass over da (gradient w.r.t. the activated S tensor) and a (the activated S tensor; unused for
 *   ACT_NONE): applies the activation derivative (NONE / LEAKY), writes the weight gradient to dw (accumulate: +=;
 *   dw == NULL: skipped) and the data gradient w.r.t. x to dx (NULL: skipped).  ws: dcv_img_conv_bwd_workspace_bytes(g). */
/* BatchNorm + (Leaky)ReLU applied on load ("pre-BN", NULL = off): the 64 channels [c0, c0 + 64) of the 128-channel input xb of
 * dcv_img_conv_scatter hold the PRE-BatchNorm convolution output z of the producing block (cgen up_blocks.5,
 * generator.py:239-248), and the kernel uses a = leaky_relu(((z - mean) * invstd) * gamma + beta, slope) rounded to bf16 -
 * exactly what dcv_bn_act would have written, minus the pass over the tensor.  For forward passes that are not differentiated
 * (the fake batch of the D-phase, trainer.py:303-309; sampling).  mean / invstd / gamma / beta: fp32 [64] (gamma and beta
 * may both be NULL). */
typedef struct dcv_prebn {
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  int c0, count;      /* count must be 64, c0 a multiple of 64 */
  float slope;
} dcv_prebn;
int dcv_img_conv_supported(const dcv_geom* g, int what);
int64_t dcv_img_conv_bwd_workspace_bytes(const dcv_geom* g);
int dcv_img_conv_fwd(const dcv_geom* g, const void* x, int64_t ldx, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                     void* y, int64_t ldy, int act, float slope, void* stream);
int dcv_img_conv_scatter(const dcv_geom* g, const void* xb, int64_t ldb, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                         void* y, int64_t ldy, int act, float slope, const dcv_prebn* pre, void* stream);
int dcv_img_conv_bwd(const dcv_geom* g, const void* da, int64_t ldda, const void* a, int64_t lda, const void* x, int64_t ldx,
                     const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, int act, float slope, float* dw, int accumulate,
                     void* dx, int64_t lddx, void* ws, int64_t ws_bytes, void* stream);
/* w-tap folding for the image-like stem inputs of the discriminators (discriminator.py:79-90,180-193: conv_g on the
 * geometry channels, conv_c on the colour channels, kernel 4, stride 2, pad 1 along w):
 *   out[line][ow][k*(cg+cc) + c] = [xg | xc][line][ow*sw - pw + k][c]  (+ sigma*noise, the Noise layer), 0 outside the row
 * lines = N*T*H rows of W pixels; out has Ow = (W + 2*pw - kw)/sw + 1 positions per line with pitch ldo.
 * The stem convolution then runs with kw = 1 on kw*(cg+cc) real channels.  dcv_unfold_w is the adjoint (gradient w.r.t.
 * xg and xc from the gradient w.r.t. the folded tensor). */
int dcv_fold_w(int dtype, const void* xg, int64_t ldg, int cg, const float* noise_g, const void* xc, int64_t ldc, int cc,
               const float* noise_c, float sigma, int64_t lines, int W, int kw, int sw, int pw, void* out, int64_t ldo,
               void* stream);
int dcv_unfold_w(int dtype, const void* d2, int64_t ld2, int64_t lines, int W, int kw, int sw, int pw, void* dxg, int64_t ldg,
                 int cg, void* dxc, int64_t ldc, int cc, void* stream);
/* Dataset-side conversions on the device.  The reference converts on the host: uint8 frames -> float32 `x / 127.5 - 1`
 * after a transpose to channels-first (dataset.py:128-131,157-167), class indices -> one-hot `np.eye(25)[segm]`
 * (dataset.py:177-181), generated videos -> uint8 `clip(v,-1,1); (v+1)/2*255; astype(uint8)` (util.py:74-79).  Frames on
 * disk are channels-last uint8, i.e. already in this library's activation layout:
 *   dcv_ingest_u8     : dst[row][c] = src[row*C + c] / 127.5 - 1       (src uint8 [rows][C] dense, dst pitch ld)
 *   dcv_ingest_onehot : dst[row][c] = (idx[row] == c)                  (idx uint8 or int64, idx_bytes = 1 or 8)
 *   dcv_export_u8     : channels-last (N,T,hw,C) -> planar uint8 (N,C,T,hw) with numpy's arithmetic (bit-exact in fp32) */
int dcv_ingest_u8(int dtype, const void* src, int64_t rows, int C, void* dst, int64_t ld, void* stream);
int dcv_ingest_onehot(int dtype, const void* idx, int idx_bytes, int64_t rows, int C, void* dst, int64_t ld, void* stream);
int dcv_export_u8(int dtype, const void* src, int64_t ld, int N, int C, int T, int64_t hw, void* dst, void* stream);
/* Noise layer (discriminator.py:30-39): out = x + sigma*noise (noise fp32, dense [rows][C]) */
int dcv_add_noise(int dtype, const void* x, int64_t ldx, const float* noise, float sigma, int64_t rows,
                  int C, void* out, int64_t ldo, void* stream);
/* out (+)= x over a channel slice (gradient fan-in at U-Net skips and D inputs) */
int dcv_axpy(int dtype, const void* x, int64_t ldx, int64_t rows, int C, void* out, int64_t ldo,
             int accumulate, void* stream);
/* temporal finite difference of the gradient discriminator (discriminator.py:330-331) and its adjoint */
int dcv_tdiff(int dtype, const void* x, int64_t ldx, int N, int T, int64_t hw, int C, void* out,
              int64_t ldo, void* stream);
int dcv_tdiff_bwd(int dtype, const void* dy, int64_t lddy, int N, int T, int64_t hw, int C, void* dx,
                  int64_t lddx, int accumulate, void* stream);
/* Softmax(dim=channel) forward/backward (generator.py:75-76) */
int dcv_softmax(int dtype, const void* z, int64_t ldz, int64_t rows, int C, void* y, int64_t ldy, void* stream);
int dcv_softmax_bwd(int dtype, const void* dy, int64_t lddy, const void* y, int64_t ldy, int64_t rows,
                    int C, void* dz, int64_t lddz, void* stream);
/* segmentation remap argmax -> {-1,+1} one-hot (generator.py:378-385) */
int dcv_segm_remap(int dtype, const void* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, void* stream);

/* ---- layout conversion at the module boundary -----------------------------------------------
 * src fp32 with arbitrary element strides (n, c, t, h, w) <-> channels-last (N,T,H,W,ld) of `dtype`.
 * Replaces the permute/view/contiguous copies at generator.py:136-139,425-433 and the frame slice
 * x[:, :, t] at trainer.py:299,307,347 (pass T=1 and a pointer to frame t).
 */
int dcv_to_channels_last(int dtype, const float* src, int64_t sn, int64_t sc, int64_t st, int64_t sh,
                         int64_t sw, int N, int C, int T, int H, int W, void* dst, int64_t ld, void* stream);
int dcv_from_channels_last(int dtype, const void* src, int64_t ld, int N, int C, int T, int H, int W,
                           float* dst, int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw,
                           int accumulate, void* stream);
/* dtype cast/copy of a channels-last block, optionally selecting frame t of (N,T,HW,C) */
int dcv_copy_cl(int src_dtype, const void* src, int64_t lds, int dst_dtype, void* dst, int64_t ldd,
                int64_t rows, int C, void* stream);

/* frame t of every clip: reverse=0: dst(N,1,HW,C) = clips(N,T,HW,C)[:, t]; reverse=1: clips[:, t] (+)= dst
 * (the image discriminator's x[:, :, t] slice, trainer.py:299,307,347, and its gradient) */
int dcv_frame_copy(int dtype, void* clips, int64_t ldc, int N, int T, int64_t hw, int C, int t, const int* t_dev,
                   void* frames, int64_t ldf, int reverse, int accumulate, void* stream);
/* t_dev (device pointer, may be NULL) overrides t: the frame index then lives on the device, so a captured CUDA graph
 * of the step can be replayed with a fresh np.random.randint(T) every iteration (trainer.py:279) */

/* ---- GRU latent trajectory (generator.py:58,84-101) -----------------------------------------
 * h0 [B][D], eps [T][B][D] fp32; weights in nn.GRUCell layout (3D x D, gate order r,z,n).
 * hs out [B][T][D].  Backward = BPTT producing fp32 weight/bias grads ((+)= if accumulate).
 */
int dcv_gru_traj_fwd(const float* h0, const float* eps, const float* w_ih, const float* w_hh,
                     const float* b_ih, const float* b_hh, int B, int T, int D, float* hs, void* stream);
int dcv_gru_traj_bwd(const float* h0, const float* eps, const float* hs, const float* dhs,
                     const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B,
                     int T, int D, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                     int accumulate, void* stream);

/* ---- losses (loss.py) ------------------------------------------------------------------------
 * loss_out[0] (+)= mean_i l(y_i); dy_i = grad_scale * dl/dy_i / n.  y / dy hold n logits of `dtype`, one every
 * ldy / lddy elements (1 = dense; 16 for a single-channel activation with zero padding channels) */
int dcv_loss_fwd_bwd(int dtype, const void* y, int64_t ldy, int64_t n, int kind, float* loss_out, int accumulate,
                     void* dy, int64_t lddy, float grad_scale, void* stream);

/* ---- Adam (train.py:167-176, trainer.py:320-322,357-359) ---------------------------------------
 * One launch over `ntensors` tensors.  ptrs are HOST arrays of device pointers; numel host array.
 * L2-style weight decay (grad += wd*p), bias correction from `step` (1-based, already incremented). */
int dcv_adam_multi(int ntensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                   const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay,
                   int64_t step, float grad_scale, void* stream);
/* Flat-buffer Adam whose step counter lives on the device: step_state is 16 bytes {int64 step; float step_size;
 * float 1/sqrt(bias_correction2)}; the call first advances it (step += 1, bias corrections in fp64), then updates.
 * No host value changes between iterations, so the whole training step can be captured in a CUDA graph. */
int dcv_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, void* step_state, float grad_scale, void* stream);
/* same over one flat buffer (the data-parallel path keeps params/grads/state flat per network) */
int dcv_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCVGAN_B200_H */
