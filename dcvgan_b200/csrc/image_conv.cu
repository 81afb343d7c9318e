// HBM-bound kernels for the image-side 3x3 convolution of the colour generator:
// `Inconv` = Conv2d(C -> 64, k3, s1, p1, no bias) + LeakyReLU(0.01) with C = 1 (depth) or 2 (optical flow) input
// channels (reference: src/generator.py:158-176).
//
// As an implicit GEMM the layer has K = 9*C <= 18.  Through the tcgen05 / TMA path it needs its input padded to 16
// channels (nine 32-byte-row boxes per tile) and runs at 14 TFLOP/s: 0.167 ms at batch 32 where the HBM time of its
// 268 MB bf16 output is 0.041 ms; its backward took three passes over that tensor (act_bwd, wgrad, dgrad: 0.42 ms).
// A first CUDA-core version of these kernels was instruction-bound (0.165 / 0.83 ms, profiles/r2a_layers.md).  Here the
// arithmetic runs on warp-level mma.sync (m16n8k16, bf16 -> fp32), whose operands are assembled IN REGISTERS from
// plain coalesced loads - the tiny K makes TMA / shared-memory operand staging pure overhead - so the kernels are
// bound by the bytes of the big tensor they stream:
//   * forward : one warp = 16 consecutive pixels x 64 channels per step.  A fragments (im2col rows) are gathered from a
//               bf16 copy of the input window in shared memory, B fragments (the 9*C x 64 weight matrix, bf16-rounded
//               from the fp32 master weight) live in registers for the whole kernel; the result is staged through a
//               warp-private shared-memory tile so that global stores are 16-byte vectors covering full 128-byte rows.
//   * backward: ONE pass over (da, a).  dz = da * LeakyReLU'(a) is formed in registers directly in mma A-fragment layout
//               (the K permutation of a per-thread 2 x 16-byte load is applied to the B operand instead), used as
//                 - A of  P[pixel][tap,ci] = sum_co dz[pixel][co] * w[co][ci][tap]   (data-gradient partial products),
//                 - after a movmatrix transpose, A of  dW^T[co][tap,ci] += sum_pixel dz[pixel][co] * x[pixel + tap][ci],
//               P is staged in shared memory for a band of 8 rows + 1 halo row each side and summed over the nine taps
//               (col2im inside the block); dW is accumulated per warp in 32 registers, reduced per block, and the
//               per-block partials are summed in a fixed order by wgrad_reduce.
// Algorithmic bytes: forward 2 B/output element (+ the tiny input); backward 4 B/output element (da and a read once).
#include "common.cuh"
#include "conv_geom.cuh"

namespace dcv {

int wgrad_reduce(const float* partial, int splits, const dcv_geom* g, float* dw, int64_t s_l, int64_t s_s,
                 int64_t s_tap, int accumulate, cudaStream_t s);

constexpr int IMG_CO = 64;          // output channels
constexpr int IMG_BAND = 8;         // image rows per block
constexpr int IMG_WARPS = 8;

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---- BatchNorm + (Leaky)ReLU applied on load ("pre-BN") ------------------------------------------------------------------
// The 128-channel input of Outconv is the concat [up_blocks.5 output | Inconv skip] (generator.py:401-402).  Its first half is
// BatchNorm + ReLU of a convolution output z; materialising a = relu(bn(z)) costs a pass that reads and writes the 268 MB
// tensor (bn_act, 0.09 ms at batch 32).  With a PreBn the scatter kernel (Outconv forward) reads z itself from that half of
// the concat buffer and applies a = act(z * P + Q), P = invstd * gamma, Q = beta - mean * invstd * gamma, rounded to bf16 - the
// arithmetic and rounding of bn_act_bf16_kernel, so the operand the MMAs see is bit-identical to the materialised tensor.
// Used for forward passes that are not differentiated (the fake batch of the D-phase, sampling): the weight-gradient kernel
// with the same transform measured 0.33 ms against 0.19 + 0.09 ms for bn_act + the plain kernel, so when a backward pass
// follows the tensor is materialised as before.
struct PreBn {
  const float* mean; const float* invstd; const float* gamma; const float* beta;   // [64] each; gamma / beta may be NULL
  int half;                                                                        // 64-channel half of the input they apply to
  float slope;                                                                     // LeakyReLU slope (0 = ReLU)
};
// coefficients of the 16 channels a lane touches in a half: 8q .. 8q+7 and 32+8q .. 32+8q+7
__device__ __forceinline__ void prebn_coeffs(const PreBn& pre, int q, float (&P)[16], float (&Q)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = (i < 8 ? 8 * q + i : 32 + 8 * q + (i - 8));
    const float is = pre.invstd[c], mu = pre.mean[c];
    const float g = pre.gamma ? pre.gamma[c] : 1.f, b = pre.gamma ? pre.beta[c] : 0.f;
    P[i] = is * g; Q[i] = b - mu * is * g;
  }
}
// word p (0..7) of a lane's two 16-byte vectors holds the channel pair 2p, 2p+1 of those 16
__device__ __forceinline__ uint32_t prebn_word(uint32_t v, int p, const float (&P)[16], const float (&Q)[16], float slope) {
  float a = fmaf(bf16_lo(v), P[2 * p], Q[2 * p]), b = fmaf(bf16_hi(v), P[2 * p + 1], Q[2 * p + 1]);
  a = a > 0.f ? a : a * slope; b = b > 0.f ? b : b * slope;
  return pack_bf16x2(a, b);
}

// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col).  Fragment layout (g = lane / 4, q = lane % 4):
//   a0 = A[g][2q,2q+1]  a1 = A[g+8][2q,2q+1]  a2 = A[g][2q+8,2q+9]  a3 = A[g+8][2q+8,2q+9]
//   b0 = B[2q,2q+1][g]  b1 = B[2q+8,2q+9][g]      d0,d1 = D[g][2q,2q+1]  d2,d3 = D[g+8][2q,2q+1]
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// transpose of an 8x8 b16 matrix held one 32-bit register per thread (row g, columns 2q, 2q+1)
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// input window of one band in shared memory: rows r0-1 .. r0+BAND, columns -1 .. W, C channels, zero outside the image
template <int C>
__device__ __forceinline__ void stage_window(__nv_bfloat16* xs, const __nv_bfloat16* __restrict__ x, int64_t ldx, int n, int r0,
                                             int H, int W) {
  const int WP = W + 2;
  for (int i = threadIdx.x; i < (IMG_BAND + 2) * WP; i += blockDim.x) {
    const int lr = i / WP, lc = i % WP;
    const int h = r0 - 1 + lr, w = lc - 1;
    const bool ok = h >= 0 && h < H && w >= 0 && w < W;
    const __nv_bfloat16* p = x + ((int64_t)(n * H + h) * W + w) * ldx;
#pragma unroll
    for (int ci = 0; ci < C; ++ci) xs[i * C + ci] = ok ? p[ci] : __float2bfloat16_rn(0.f);
  }
}
// element k = tap * C + ci of the im2col row of pixel (lr, w) (lr: band-local row, window row lr + 1 is the pixel's own row)
template <int C>
__device__ __forceinline__ uint16_t im2col_at(const __nv_bfloat16* xs, int WP, int lr, int w, int k) {
  if (k >= 9 * C) return 0;
  const int tap = k / C, ci = k % C;
  const int dy = tap / 3, dx = tap % 3;
  return reinterpret_cast<const uint16_t*>(xs)[((lr + dy) * WP + (w + dx)) * C + ci];
}

// ------------------------------------------------------------------------------------------ forward
// C real input channels (1..3), NCO output channels (64 or 128, processed as halves of 64)
template <int C, int NCO, int ACT>
__global__ void __launch_bounds__(256, 2)
img_conv3x3_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                       int N, int H, int W, __nv_bfloat16* __restrict__ y, int64_t ldy, float slope) {
  auto actf = [&](float v) { return ACT == DCV_ACT_LEAKY ? (v > 0.f ? v : v * slope) : v; };
  pdl_wait(); pdl_trigger();
  constexpr int KS = (9 * C + 15) / 16;                        // k16 steps
  constexpr int NH = NCO / 64;                                 // halves of 64 output channels
  __shared__ __align__(16) __nv_bfloat16 xs[(IMG_BAND + 2) * 66 * C];
  __shared__ __align__(16) uint32_t stg[IMG_WARPS][16][36];     // warp-private 16 x 64 bf16 tile, 144-byte row pitch
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, g = lane / 4, q = lane % 4;
  const int bands = H / IMG_BAND;
  const int WP = W + 2;
  // B fragments: B[k][co] = w[co][ci][tap] (k = tap*C + ci), bf16; kept in shared memory in fragment order (one word per lane:
  // conflict-free), 64-128 registers otherwise
  __shared__ uint32_t bw[NH][KS][8][2][32];
#pragma unroll
  for (int hf = 0; hf < NH; ++hf)
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if ((((hf * KS + s) * 8 + j) * 2 + h) % IMG_WARPS != warp) continue;      // the fragments are spread over the warps
          float v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 16 * s + 2 * q + 8 * h + e;
            v[e] = k < 9 * C ? w[(64 * hf + 8 * j + g) * s_s + (k % C) * s_l + (k / C)] : 0.f;
          }
          bw[hf][s][j][h][lane] = pack_bf16x2(v[0], v[1]);
        }
  const int tiles_w = W / 16, tiles = IMG_BAND * tiles_w;
  // persistent over (image, band) items: the weight fragments are built once per block
  for (int item = blockIdx.x; item < N * bands; item += gridDim.x) {
    const int n = item / bands, r0 = (item % bands) * IMG_BAND;
    __syncthreads();                               // the previous item's readers of xs are done
    stage_window<C>(xs, x, ldx, n, r0, H, W);
    __syncthreads();
    for (int t = warp; t < tiles; t += IMG_WARPS) {
      const int lr = t / tiles_w, w0 = (t % tiles_w) * 16;
      uint32_t a[KS][4];
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)              // column half (k, k + 8)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {       // row half (g, g + 8)
            const int k = 16 * s + 2 * q + 8 * h;
            const uint32_t lo = im2col_at<C>(xs, WP, lr, w0 + g + 8 * rr, k), hi = im2col_at<C>(xs, WP, lr, w0 + g + 8 * rr, k + 1);
            a[s][2 * h + rr] = lo | (hi << 16);
          }
      const int64_t pix0 = (int64_t)(n * H + r0 + lr) * W + w0;
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
          for (int j = 0; j < 8; ++j) mma16816(acc[j], a[s][0], a[s][1], a[s][2], a[s][3], bw[hf][s][j][0][lane], bw[hf][s][j][1][lane]);
        // activation, bf16, warp-private staging (conflict-free: bank = 4 * row + 4 * j + q), then 16-byte global stores
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          stg[warp][g][4 * j + q] = pack_bf16x2(actf(acc[j][0]), actf(acc[j][1]));
          stg[warp][g + 8][4 * j + q] = pack_bf16x2(actf(acc[j][2]), actf(acc[j][3]));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = 4 * i + lane / 8, chunk = lane % 8;
          const uint4 v = *reinterpret_cast<const uint4*>(&stg[warp][row][4 * chunk]);
          *reinterpret_cast<uint4*>(y + (pix0 + row) * ldy + 64 * hf + 8 * chunk) = v;
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
// channel that K slot sigma (0..15) of k16 step s carries: thread q loads channels 8q..8q+7 and 32+8q..32+8q+7 of a pixel
// (two 16-byte loads, 64 contiguous bytes per row and instruction across the quad) and feeds value i = 4s + 2h + e into
// slot 2q + 8h + e
__device__ __forceinline__ int slot_channel(int s, int sigma) {
  const int qq = (sigma % 8) / 2, e = sigma % 2, h = sigma / 8;
  const int i = 4 * s + 2 * h + e;
  return i < 8 ? 8 * qq + i : 32 + 8 * qq + (i - 8);
}

// NCO = channels of the big tensor (64; 128 = weight gradient only, the two channel halves are owned by warps 0-3 / 4-7),
// DGRAD: also form the data gradient dx
template <int C, int NCO, bool DGRAD>
__global__ void __launch_bounds__(256, 2)
img_conv3x3_bwd_kernel(const __nv_bfloat16* __restrict__ da, int64_t ldda, const __nv_bfloat16* __restrict__ a, int64_t lda,
                       const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                       int N, int H, int W, int act, float slope, float* __restrict__ partial, __nv_bfloat16* __restrict__ dx,
                       int64_t lddx) {
  pdl_wait(); pdl_trigger();
  constexpr int NT = 9 * C;                    // (tap, ci) pairs
  constexpr int NJ = (NT + 7) / 8;             // n8 tiles over them
  constexpr int NP = NJ * 8 + 8;               // row length of P: 8 floats of padding keep the float2 fragment stores conflict-free
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* Ps = reinterpret_cast<float*>(sm_raw);                                  // [(BAND+2) * W][NP]; later [8 warps][NP][64]
  // region 0: the P tile (data gradient only) and, after the loop, the per-warp weight-gradient tiles; region 1: input window
  const int tile_floats = DGRAD ? (IMG_BAND + 2) * W * NP : 0, red_floats = IMG_WARPS * NP * IMG_CO;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(Ps + (tile_floats > red_floats ? tile_floats : red_floats));
  static_assert(NCO == 64 || !DGRAD, "the data gradient needs all channels of a pixel in one warp");
  constexpr int NH = NCO / 64, WPH = IMG_WARPS / NH;             // channel halves, warps per half
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, g = lane / 4, q = lane % 4;
  const int half = warp / WPH, wsub = warp % WPH;
  const int bands = H / IMG_BAND;
  const int WP = W + 2;
  // B fragments of the data-gradient GEMM: B[k slot][nn] = w[co(s, slot)][ci][tap], nn = tap*C + ci (0 beyond 9*C)
  uint32_t bw[DGRAD ? 4 : 1][NJ][2];
  if (DGRAD) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[2];
          const int nn = 8 * j + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = slot_channel(s, 2 * q + 8 * h + e);
            v[e] = nn < NT ? w[co * s_s + (nn % C) * s_l + (nn / C)] : 0.f;
          }
          bw[DGRAD ? s : 0][j][h] = pack_bf16x2(v[0], v[1]);
        }
  }
  float acc[4][NJ][4];                         // dW^T: m tile s (16 channels), n tile j (8 (tap, ci) pairs)
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < NJ; ++j) { acc[s][j][0] = acc[s][j][1] = acc[s][j][2] = acc[s][j][3] = 0.f; }

  const int tiles_w = W / 16, tiles = (IMG_BAND + 2) * tiles_w;
  // persistent over (image, band) items: weight fragments built once, dW accumulated across all items of the block
  for (int item = blockIdx.x; item < N * bands; item += gridDim.x) {
  const int n = item / bands, r0 = (item % bands) * IMG_BAND;
  __syncthreads();                               // the previous item's readers of xs / Ps are done
  stage_window<C>(xs, x, ldx, n, r0, H, W);
  __syncthreads();
  for (int t = wsub; t < tiles; t += WPH) {
    const int prow = t / tiles_w, w0 = (t % tiles_w) * 16;       // prow 0 / BAND+1 = halo rows
    const int h = r0 - 1 + prow;
    const bool live = h >= 0 && h < H && (DGRAD || (prow >= 1 && prow <= IMG_BAND));   // warp-uniform; halo rows only feed the data gradient
    const bool inband = prow >= 1 && prow <= IMG_BAND;
    // dz in A-fragment layout: reg[s][0] rows g (slots 2q,2q+1), [s][1] rows g+8, [s][2] rows g (slots 2q+8,+9), [s][3] rows g+8
    uint32_t dzr[4][4];
    if (live) {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t pix = (int64_t)(n * H + h) * W + w0 + g + 8 * rr;
        uint4 d[2], o[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          d[hf] = *reinterpret_cast<const uint4*>(da + pix * ldda + 64 * half + 32 * hf + 8 * q);
          if (act != DCV_ACT_NONE) o[hf] = *reinterpret_cast<const uint4*>(a + pix * lda + 64 * half + 32 * hf + 8 * q);
        }
        const uint32_t dw_[8] = {d[0].x, d[0].y, d[0].z, d[0].w, d[1].x, d[1].y, d[1].z, d[1].w};
        const uint32_t ow_[8] = {o[0].x, o[0].y, o[0].z, o[0].w, o[1].x, o[1].y, o[1].z, o[1].w};
#pragma unroll
        for (int p = 0; p < 8; ++p) {          // value pair i = 2p, 2p+1  ->  step s = p / 2, column half = p % 2
          uint32_t v = dw_[p];
          if (act != DCV_ACT_NONE) {
            const float g0 = act_grad_from_out(bf16_lo(ow_[p]), act, slope), g1 = act_grad_from_out(bf16_hi(ow_[p]), act, slope);
            v = pack_bf16x2(bf16_lo(v) * g0, bf16_hi(v) * g1);
          }
          dzr[p / 2][2 * (p % 2) + rr] = v;
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s) { dzr[s][0] = dzr[s][1] = dzr[s][2] = dzr[s][3] = 0u; }
    }
    // ---- data gradient partial products P[pixel][nn]
    if (DGRAD && dx != nullptr) {
      float pacc[NJ][4];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { pacc[j][0] = pacc[j][1] = pacc[j][2] = pacc[j][3] = 0.f; }
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma16816(pacc[j], dzr[s][0], dzr[s][1], dzr[s][2], dzr[s][3], bw[DGRAD ? s : 0][j][0], bw[DGRAD ? s : 0][j][1]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        *reinterpret_cast<float2*>(Ps + (size_t)(prow * W + w0 + g) * NP + 8 * j + 2 * q) = make_float2(pacc[j][0], pacc[j][1]);
        *reinterpret_cast<float2*>(Ps + (size_t)(prow * W + w0 + g + 8) * NP + 8 * j + 2 * q) = make_float2(pacc[j][2], pacc[j][3]);
      }
    }
    // ---- weight gradient: dW^T[co][nn] += dz^T[co][pixel] * X[pixel][nn]   (K = the 16 pixels of the tile)
    if (inband && live && partial != nullptr) {
      uint32_t xb[NJ][2];                       // B[k = pixel 2q+e (+8)][nn = 8j + g]
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int nn = 8 * j + g;
          const uint32_t lo = im2col_at<C>(xs, WP, prow - 1, w0 + 2 * q + 8 * hh, nn), hi = im2col_at<C>(xs, WP, prow - 1, w0 + 2 * q + 8 * hh + 1, nn);
          xb[j][hh] = lo | (hi << 16);
        }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        // transposed 8x8 blocks: [s][0] = (pixels 0-7, slots 0-7), [1] = (pixels 8-15, slots 0-7), [2] = (0-7, 8-15), [3] = (8-15, 8-15)
        const uint32_t t0 = movmatrix_trans(dzr[s][0]), t1 = movmatrix_trans(dzr[s][1]);
        const uint32_t t2 = movmatrix_trans(dzr[s][2]), t3 = movmatrix_trans(dzr[s][3]);
        // A of the m tile: rows 0-7 = slots 0-7, rows 8-15 = slots 8-15; k 0-7 = pixels 0-7, k 8-15 = pixels 8-15
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma16816(acc[s][j], t0, t2, t1, t3, xb[j][0], xb[j][1]);
      }
    }
  }
  __syncthreads();
  // ---- data gradient of the band: dx[h][w][ci] = sum_taps P[h - ky + 1][w - kx + 1][tap][ci]
  if (DGRAD && dx != nullptr) {
    for (int p = threadIdx.x; p < IMG_BAND * W; p += blockDim.x) {
      const int hb = p / W, wq = p % W;
      float s[C];
#pragma unroll
      for (int ci = 0; ci < C; ++ci) s[ci] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int pr = hb + 2 - ky, pc = wq + 1 - kx;          // P row (tile-local: band row hb is P row hb + 1), column
          if (pc < 0 || pc >= W) continue;
          const float* qp = Ps + (size_t)(pr * W + pc) * NP + (ky * 3 + kx) * C;
#pragma unroll
          for (int ci = 0; ci < C; ++ci) s[ci] += qp[ci];
        }
      __nv_bfloat16* d = dx + ((int64_t)(n * H + r0 + hb) * W + wq) * lddx;
#pragma unroll
      for (int ci = 0; ci < C; ++ci) d[ci] = __float2bfloat16_rn(s[ci]);
    }
  }
  }   // items
  if (partial == nullptr) return;
  __syncthreads();
  // ---- weight gradient: per-warp accumulators -> shared [warp][nn][co] -> fixed-order sum over the warps
  float* red = Ps;
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int co = slot_channel(s, g + 8 * rr);
#pragma unroll
        for (int e = 0; e < 2; ++e) red[((size_t)warp * NP + 8 * j + 2 * q + e) * IMG_CO + co] = acc[s][j][2 * rr + e];
      }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * NCO; i += blockDim.x) {
    const int nn = i / NCO, co = i % NCO;
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < WPH; ++wq) s += red[((size_t)((co / 64) * WPH + wq) * NP + nn) * IMG_CO + (co % 64)];
    partial[(size_t)blockIdx.x * NT * NCO + i] = s;
  }
}

// ------------------------------------------------------------------------------------------ scatter pass (Outconv forward)
// y (the small L tensor, C real channels) = act( sum over taps of the partial products of the big S tensor ):
//   P[pixel][tap, cl] = sum_cs S[pixel][cs] * w[cl][cs][tap]        one pass over S, A fragments straight from global loads
//   y[h][w][cl]       = act( sum_taps P[h - ky + 1][w - kx + 1][tap][cl] )     col2im inside the block (band + halo rows)
// This is the forward of `Outconv` = ConvTranspose2d(128, 3, 3, 1, 1) + Tanh (generator.py:272-277), which ran as a 1x1
// tensor-core GEMM onto 27 (padded 32) columns plus a col2im pass over a 134 MB intermediate (0.204 ms at batch 32); here the
// 537 MB input is read once and nothing but the 3-channel output is written.
template <int C, int NCO, bool PRE = false>
__global__ void __launch_bounds__(256, 2)
img_conv3x3_scatter_kernel(const __nv_bfloat16* __restrict__ xb, int64_t ldb, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                           int N, int H, int W, int act, float slope, __nv_bfloat16* __restrict__ y, int64_t ldy, const PreBn pre) {
  pdl_wait(); pdl_trigger();
  constexpr int NT = 9 * C, NJ = (NT + 7) / 8, NP = NJ * 8 + 8, NH = NCO / 64;   // NP: P row with 8 floats of padding (conflict-free)
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* Ps = reinterpret_cast<float*>(sm_raw);                                  // [(BAND+2) * W][NP]
  uint32_t* bws = reinterpret_cast<uint32_t*>(Ps + (size_t)(IMG_BAND + 2) * W * NP);   // [NH][4][NJ][2][32] B fragments, one word per lane
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, g = lane / 4, q = lane % 4;
  const int bands = H / IMG_BAND;
#pragma unroll
  for (int hf = 0; hf < NH; ++hf)
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if ((((hf * 4 + s) * NJ + j) * 2 + h) % IMG_WARPS != warp) continue;   // the fragments are spread over the warps
          float v[2];
          const int nn = 8 * j + g;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = 64 * hf + slot_channel(s, 2 * q + 8 * h + e);
            v[e] = nn < NT ? w[co * s_s + (nn % C) * s_l + (nn / C)] : 0.f;
          }
          bws[((((hf * 4 + s) * NJ + j) * 2 + h) << 5) + lane] = pack_bf16x2(v[0], v[1]);
        }
  const int tiles_w = W / 16, tiles = (IMG_BAND + 2) * tiles_w;
  float preP[PRE ? 16 : 1], preQ[PRE ? 16 : 1];
  if constexpr (PRE) prebn_coeffs(pre, q, preP, preQ);
  for (int item = blockIdx.x; item < N * bands; item += gridDim.x) {
    const int n = item / bands, r0 = (item % bands) * IMG_BAND;
    __syncthreads();                               // B fragments written / the previous item's readers of Ps are done
    // two tiles per warp and step: every B fragment read from shared memory feeds two MMAs
    for (int t = warp; t < tiles; t += 2 * IMG_WARPS) {
      int prow[2], w0[2]; bool live[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tt = t + u * IMG_WARPS;
        prow[u] = tt / tiles_w; w0[u] = (tt % tiles_w) * 16;
        const int h = r0 - 1 + prow[u];
        live[u] = tt < tiles && h >= 0 && h < H;   // warp-uniform
      }
      float pacc[2][NJ][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { pacc[u][j][0] = pacc[u][j][1] = pacc[u][j][2] = pacc[u][j][3] = 0.f; }
#pragma unroll
      for (int hf = 0; hf < NH; ++hf) {
        uint32_t ar[2][4][4];                      // A fragments of the 64 channels of this half (see slot_channel)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            uint4 d0 = make_uint4(0u, 0u, 0u, 0u), d1 = d0;
            if (live[u]) {
              const int64_t pix = (int64_t)(n * H + r0 - 1 + prow[u]) * W + w0[u] + g + 8 * rr;
              d0 = *reinterpret_cast<const uint4*>(xb + pix * ldb + 64 * hf + 8 * q);
              d1 = *reinterpret_cast<const uint4*>(xb + pix * ldb + 64 * hf + 32 + 8 * q);
            }
            uint32_t dw_[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            if constexpr (PRE) {
              if (hf == pre.half && live[u]) {     // padding rows stay zero AFTER the activation
#pragma unroll
                for (int p = 0; p < 8; ++p) dw_[p] = prebn_word(dw_[p], p, preP, preQ, pre.slope);
              }
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) ar[u][p / 2][2 * (p % 2) + rr] = dw_[p];
          }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const uint32_t b0 = bws[((((hf * 4 + s) * NJ + j) * 2 + 0) << 5) + lane], b1 = bws[((((hf * 4 + s) * NJ + j) * 2 + 1) << 5) + lane];
            mma16816(pacc[0][j], ar[0][s][0], ar[0][s][1], ar[0][s][2], ar[0][s][3], b0, b1);
            mma16816(pacc[1][j], ar[1][s][0], ar[1][s][1], ar[1][s][2], ar[1][s][3], b0, b1);
          }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (t + u * IMG_WARPS >= tiles) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          *reinterpret_cast<float2*>(Ps + (size_t)(prow[u] * W + w0[u] + g) * NP + 8 * j + 2 * q) = make_float2(pacc[u][j][0], pacc[u][j][1]);
          *reinterpret_cast<float2*>(Ps + (size_t)(prow[u] * W + w0[u] + g + 8) * NP + 8 * j + 2 * q) = make_float2(pacc[u][j][2], pacc[u][j][3]);
        }
      }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < IMG_BAND * W; p += blockDim.x) {
      const int hb = p / W, wq = p % W;
      float sacc[C];
#pragma unroll
      for (int ci = 0; ci < C; ++ci) sacc[ci] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int pr = hb + 2 - ky, pc = wq + 1 - kx;          // P row (band row hb is P row hb + 1), column
          if (pc < 0 || pc >= W) continue;
          const float* qp = Ps + (size_t)(pr * W + pc) * NP + (ky * 3 + kx) * C;
#pragma unroll
          for (int ci = 0; ci < C; ++ci) sacc[ci] += qp[ci];
        }
      __nv_bfloat16* d = y + ((int64_t)(n * H + r0 + hb) * W + wq) * ldy;
#pragma unroll
      for (int ci = 0; ci < C; ++ci) d[ci] = __float2bfloat16_rn(apply_act(sacc[ci], act, slope));
    }
  }
}

// ------------------------------------------------------------------------------------------ host side
// what = 0: forward-type pass (small-channel L tensor -> S tensor with 64 or 128 channels; Inconv forward, Outconv data
//           gradient), 1: full backward of Inconv (da, a -> dW, dx; S has 64 channels),
//           2: weight gradient only (S has 64 or 128 channels; Outconv), 3: scatter pass S -> L (Outconv forward).
//           k3 / s1 / p1, 2-D, bf16.
int img_conv_supported_for(const dcv_geom* g, int what) {
  const int wcl = g->wCl > 0 ? g->wCl : g->Cl, wcs = g->wCs > 0 ? g->wCs : g->Cs;
  if (g->kt != 1 || g->kh != 3 || g->kw != 3 || g->st != 1 || g->sh != 1 || g->sw != 1 || g->pt != 0 || g->ph != 1 || g->pw != 1) return 0;
  if (g->Tl != 1 || g->Ts != 1 || g->Hl != g->Hs || g->Wl != g->Ws) return 0;
  if (wcs != g->Cs || (g->Cs != 64 && g->Cs != 128)) return 0;
  if (g->Wl % 16 || g->Hl % IMG_BAND || g->Wl > 64) return 0;
  if (what == 1) return wcl >= 1 && wcl <= 2 && g->Cs == 64;
  return wcl >= 1 && wcl <= 3;
}
int img_conv_supported(const dcv_geom* g) { return img_conv_supported_for(g, 1); }

static int bwd_smem_bytes(int C, int W, bool dgrad) {
  const int np = (9 * C + 7) / 8 * 8 + 8;
  int tile = dgrad ? (IMG_BAND + 2) * W * np : 0, red = IMG_WARPS * np * IMG_CO;
  return (tile > red ? tile : red) * (int)sizeof(float) + (IMG_BAND + 2) * (W + 2) * C * 2 + 16;
}

static int img_items(const dcv_geom* g) { return g->N * (g->Hl / IMG_BAND); }
// persistent grids: forward 2-3 blocks / SM, backward 2 blocks / SM (116 registers, ~42 KB)
static int img_fwd_blocks(const dcv_geom* g) { const int it = img_items(g); return it < 148 * 3 ? it : 148 * 3; }
int img_conv_bwd_blocks(const dcv_geom* g) { const int it = img_items(g); return it < 148 * 2 ? it : 148 * 2; }

int64_t img_conv_bwd_ws_bytes(const dcv_geom* g) {
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  return (int64_t)img_conv_bwd_blocks(g) * 9 * C * g->Cs * sizeof(float);
}

template <int C, int NCO>
static void launch_fwd(const dcv_geom* g, unsigned blocks, const void* x, int64_t ldx, const float* w, int64_t s_l, int64_t s_s, void* y,
                       int64_t ldy, int act, float slope, cudaStream_t s) {
  if (act == DCV_ACT_LEAKY)
    launch_k(img_conv3x3_fwd_kernel<C, NCO, DCV_ACT_LEAKY>, blocks, 256, 0, s, (const __nv_bfloat16*)x, ldx, w, s_l, s_s, g->N, g->Hl, g->Wl, (__nv_bfloat16*)y, ldy, slope);
  else
    launch_k(img_conv3x3_fwd_kernel<C, NCO, DCV_ACT_NONE>, blocks, 256, 0, s, (const __nv_bfloat16*)x, ldx, w, s_l, s_s, g->N, g->Hl, g->Wl, (__nv_bfloat16*)y, ldy, slope);
}

int img_conv_fwd(const dcv_geom* g, const void* x, int64_t ldx, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, void* y,
                 int64_t ldy, int act, float slope, cudaStream_t s) {
  DCV_REQUIRE(img_conv_supported_for(g, 0), "img_conv_fwd: geometry not supported");
  DCV_REQUIRE(act == DCV_ACT_NONE || act == DCV_ACT_LEAKY, "img_conv_fwd: activation %d", act);
  DCV_REQUIRE(s_tap == 1, "img_conv_fwd: taps of the master weight must be contiguous");
  DCV_REQUIRE((((uintptr_t)y) & 15) == 0 && ldy % 8 == 0, "img_conv_fwd: output must be 16-byte aligned");
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  const unsigned blocks = (unsigned)img_fwd_blocks(g);
#define DCV_IMG_FWD(C_, N_) if (C == C_ && g->Cs == N_) launch_fwd<C_, N_>(g, blocks, x, ldx, w, s_l, s_s, y, ldy, act, slope, s);
  DCV_IMG_FWD(1, 64) DCV_IMG_FWD(2, 64) DCV_IMG_FWD(3, 64) DCV_IMG_FWD(1, 128) DCV_IMG_FWD(2, 128) DCV_IMG_FWD(3, 128)
#undef DCV_IMG_FWD
  return check_launch("img_conv3x3_fwd");
}

template <int C, int NCO, bool DGRAD>
static int launch_bwd(const dcv_geom* g, int blocks, int smem, const void* da, int64_t ldda, const void* a, int64_t lda, const void* x,
                      int64_t ldx, const float* w, int64_t s_l, int64_t s_s, int act, float slope, float* partial, void* dx, int64_t lddx,
                      cudaStream_t s) {
  static int smem_set = 0;
  if (smem > smem_set) {
    DCV_CUDA(cudaFuncSetAttribute(img_conv3x3_bwd_kernel<C, NCO, DGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  launch_k(img_conv3x3_bwd_kernel<C, NCO, DGRAD>, blocks, 256, smem, s, (const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)a, lda,
           (const __nv_bfloat16*)x, ldx, w, s_l, s_s, g->N, g->Hl, g->Wl, act, slope, partial, (__nv_bfloat16*)dx, lddx);
  return 0;
}

template <int C, int NCO, bool PRE = false>
static int launch_scatter(const dcv_geom* g, int blocks, const void* xb, int64_t ldb, const float* w, int64_t s_l, int64_t s_s, int act,
                          float slope, void* y, int64_t ldy, const PreBn& pre, cudaStream_t s) {
  const int nj = (9 * C + 7) / 8, np = nj * 8 + 8;
  const int smem = (IMG_BAND + 2) * g->Wl * np * (int)sizeof(float) + (NCO / 64) * 4 * nj * 2 * 32 * 4 + 16;
  static int smem_set = 0;
  if (smem > smem_set) {
    DCV_CUDA(cudaFuncSetAttribute(img_conv3x3_scatter_kernel<C, NCO, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  launch_k(img_conv3x3_scatter_kernel<C, NCO, PRE>, blocks, 256, smem, s, (const __nv_bfloat16*)xb, ldb, w, s_l, s_s, g->N, g->Hl, g->Wl, act, slope,
           (__nv_bfloat16*)y, ldy, pre);
  return 0;
}

static int make_prebn(const dcv_geom* g, const dcv_prebn* pre, PreBn* out) {
  memset(out, 0, sizeof(*out));
  out->half = -1;
  if (!pre) return 0;
  DCV_REQUIRE(pre->mean && pre->invstd && (!pre->gamma == !pre->beta), "pre-BN: null statistics / gamma without beta");
  DCV_REQUIRE(pre->count == 64 && pre->c0 % 64 == 0 && pre->c0 >= 0 && pre->c0 + 64 <= g->Cs,
              "pre-BN: channels [%d, %d) are not a 64-channel half of the %d-channel input", pre->c0, pre->c0 + pre->count, g->Cs);
  DCV_REQUIRE(pre->slope >= 0.f, "pre-BN: negative LeakyReLU slope");
  out->mean = pre->mean; out->invstd = pre->invstd; out->gamma = pre->gamma; out->beta = pre->beta;
  out->half = pre->c0 / 64; out->slope = pre->slope;
  return 0;
}

int img_conv_scatter(const dcv_geom* g, const void* xb, int64_t ldb, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, void* y,
                     int64_t ldy, int act, float slope, const dcv_prebn* prebn, cudaStream_t s) {
  PreBn pre;
  if (int prc = make_prebn(g, prebn, &pre)) return prc;
  DCV_REQUIRE(img_conv_supported_for(g, 3), "img_conv_scatter: geometry not supported");
  DCV_REQUIRE(s_tap == 1, "img_conv_scatter: taps of the master weight must be contiguous");
  DCV_REQUIRE((((uintptr_t)xb) & 15) == 0 && ldb % 8 == 0, "img_conv_scatter: input must be 16-byte aligned");
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  const int blocks = img_conv_bwd_blocks(g);
  int rc = -1;
#define DCV_IMG_SC(C_, N_) if (C == C_ && g->Cs == N_ && pre.half < 0) rc = launch_scatter<C_, N_>(g, blocks, xb, ldb, w, s_l, s_s, act, slope, y, ldy, pre, s);
  DCV_IMG_SC(1, 64) DCV_IMG_SC(2, 64) DCV_IMG_SC(3, 64) DCV_IMG_SC(1, 128) DCV_IMG_SC(2, 128) DCV_IMG_SC(3, 128)
#undef DCV_IMG_SC
  if (pre.half >= 0 && C == 3 && g->Cs == 128) rc = launch_scatter<3, 128, true>(g, blocks, xb, ldb, w, s_l, s_s, act, slope, y, ldy, pre, s);
  DCV_REQUIRE(rc != -1, "img_conv_scatter: no kernel for C %d, %d channels", C, g->Cs);
  if (rc) return rc;
  return check_launch("img_conv3x3_scatter");
}

int img_conv_bwd(const dcv_geom* g, const void* da, int64_t ldda, const void* a, int64_t lda, const void* x, int64_t ldx,
                 const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, int act, float slope, float* dw, int accumulate,
                 void* dx, int64_t lddx, void* ws, int64_t ws_bytes, cudaStream_t s) {
  const bool dgrad = dx != nullptr;
  DCV_REQUIRE(img_conv_supported_for(g, dgrad ? 1 : 2), "img_conv_bwd: geometry not supported");
  DCV_REQUIRE(s_tap == 1, "img_conv_bwd: taps of the master weight must be contiguous");
  DCV_REQUIRE(act == DCV_ACT_NONE || act == DCV_ACT_LEAKY, "img_conv_bwd: activation %d", act);
  DCV_REQUIRE(dgrad || dw, "img_conv_bwd: nothing to compute");
  DCV_REQUIRE(!dw || ws_bytes >= img_conv_bwd_ws_bytes(g), "img_conv_bwd: workspace too small");
  DCV_REQUIRE((((uintptr_t)da) & 15) == 0 && ldda % 8 == 0, "img_conv_bwd: da must be 16-byte aligned");
  DCV_REQUIRE(act == DCV_ACT_NONE || ((((uintptr_t)a) & 15) == 0 && lda % 8 == 0), "img_conv_bwd: a must be 16-byte aligned");
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  const int blocks = img_conv_bwd_blocks(g);
  const int smem = bwd_smem_bytes(C, g->Wl, dgrad);
  float* partial = dw ? (float*)ws : nullptr;
  int rc = -1;
#define DCV_IMG_BWD(C_, N_, D_) if (C == C_ && g->Cs == N_ && dgrad == D_) rc = launch_bwd<C_, N_, D_>(g, blocks, smem, da, ldda, a, lda, x, ldx, w, s_l, s_s, act, slope, partial, dx, lddx, s);
  DCV_IMG_BWD(1, 64, true) DCV_IMG_BWD(2, 64, true)
  DCV_IMG_BWD(1, 64, false) DCV_IMG_BWD(2, 64, false) DCV_IMG_BWD(3, 64, false) DCV_IMG_BWD(1, 128, false) DCV_IMG_BWD(2, 128, false) DCV_IMG_BWD(3, 128, false)
#undef DCV_IMG_BWD
  DCV_REQUIRE(rc != -1, "img_conv_bwd: no kernel for C %d, %d channels, dgrad %d", C, g->Cs, (int)dgrad);
  if (rc) return rc;
  if (int r2 = check_launch("img_conv3x3_bwd")) return r2;
  if (!dw) return 0;
  // partial sums [blocks][9][C][Cs] == the split layout of wgrad_reduce for a geometry with Cl = C real channels
  dcv_geom g2 = *g;
  g2.Cl = C; g2.wCl = C;
  return wgrad_reduce((const float*)ws, blocks, &g2, dw, s_l, s_s, s_tap, accumulate, s);
}

}  // namespace dcv
