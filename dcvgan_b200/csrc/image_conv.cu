// Direct (CUDA-core, HBM-bound) kernels for the image-side 3x3 convolution of the colour generator:
// `Inconv` = Conv2d(C -> 64, k3, s1, p1, no bias) + LeakyReLU(0.01) with C = 1 (depth) or 2 (optical flow) input
// channels (reference: src/generator.py:158-176).
//
// As an implicit GEMM this layer has K = 9*C <= 18: on the tensor cores it needs its input padded to 16 channels and
// runs at 14 TFLOP/s (0.167 ms at batch 32) while the HBM time of its 268 MB bf16 output is 0.041 ms.  Here
//   * forward : one thread = 4 consecutive pixels x 8 output channels; the 9*C weights rows it needs sit in shared
//               memory as fp32 [tap][ci][co] (read straight from the fp32 master weight - no packing pass), the
//               18*C input values in registers; output written as 16-byte vectors (128 contiguous bytes per pixel).
//   * backward: ONE pass over (da, a) produces everything the layer's backward needs - the LeakyReLU derivative is
//               applied on the fly (no separate act_bwd pass over the 268 MB tensor), the weight gradient is
//               accumulated in registers per thread (8 output channels x 9*C taps) and reduced per block into fp32
//               partials, and the data gradient is formed per block from per-pixel partial products
//               P[pixel][tap][ci] = sum_co dz[pixel][co] * w[co][ci][tap] (8 lanes per pixel, butterfly reduction)
//               staged in shared memory for a band of 8 rows plus one halo row on each side, then summed over the
//               nine taps (col2im inside the block).
// Algorithmic bytes: forward 2 B/output element (+ the tiny input); backward 4 B/output element (da and a read once).
#include "common.cuh"
#include "conv_geom.cuh"

namespace dcv {

int wgrad_reduce(const float* partial, int splits, const dcv_geom* g, float* dw, int64_t s_l, int64_t s_s,
                 int64_t s_tap, int accumulate, cudaStream_t s);

constexpr int IMG_CO = 64;          // output channels (8 lanes x 8)
constexpr int IMG_PX = 4;           // consecutive pixels per thread
constexpr int IMG_BAND = 8;         // rows per block in the backward kernel

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(u[i] << 16); f[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]); u[i] = *reinterpret_cast<uint32_t*>(&h); }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

// fp32 master weight w[co*s_s + ci*s_l + tap] -> shared [tap*C + ci][co]
template <int C>
__device__ __forceinline__ void load_weights(float* ws, const float* __restrict__ w, int64_t s_l, int64_t s_s) {
  for (int i = threadIdx.x; i < 9 * C * IMG_CO; i += blockDim.x) {
    const int co = i % IMG_CO, tc = i / IMG_CO, tap = tc / C, ci = tc % C;
    ws[i] = w[co * s_s + ci * s_l + tap];
  }
}

// the 3 x 6 x C input window of a group of 4 pixels (zero outside the image)
template <int C>
__device__ __forceinline__ void load_window(const __nv_bfloat16* __restrict__ x, int64_t ldx, int n, int h, int w0, int H, int W,
                                            float (&xin)[3][IMG_PX + 2][C]) {
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int hh = h + dy - 1;
#pragma unroll
    for (int dx = 0; dx < IMG_PX + 2; ++dx) {
      const int ww = w0 + dx - 1;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
      const __nv_bfloat16* p = x + ((int64_t)(n * H + hh) * W + ww) * ldx;
      if (C == 2) {
        uint32_t v = ok ? *reinterpret_cast<const uint32_t*>(p) : 0u;
        xin[dy][dx][0] = __uint_as_float(v << 16);
        xin[dy][dx][C - 1] = __uint_as_float(v & 0xFFFF0000u);
      } else {
#pragma unroll
        for (int ci = 0; ci < C; ++ci) xin[dy][dx][ci] = ok ? __bfloat162float(p[ci]) : 0.f;
      }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(256)
img_conv3x3_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                       int N, int H, int W, __nv_bfloat16* __restrict__ y, int64_t ldy, int act, float slope) {
  __shared__ __align__(16) float ws[9 * C * IMG_CO];
  load_weights<C>(ws, w, s_l, s_s);
  __syncthreads();
  const int cg = threadIdx.x % 8, pg = threadIdx.x / 8;
  const int gpr = W / IMG_PX;
  const int64_t groups = (int64_t)N * H * gpr;
  const int64_t gid = (int64_t)blockIdx.x * 32 + pg;
  if (gid >= groups) return;
  const int w0 = (int)(gid % gpr) * IMG_PX;
  const int64_t row = gid / gpr;
  const int h = (int)(row % H), n = (int)(row / H);
  float xin[3][IMG_PX + 2][C];
  load_window<C>(x, ldx, n, h, w0, H, W, xin);
  float acc[IMG_PX][8];
#pragma unroll
  for (int j = 0; j < IMG_PX; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[j][k] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < C; ++ci) {
        const float4* wp = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * C + ci) * IMG_CO + cg * 8);
        const float4 wa = wp[0], wb = wp[1];
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < IMG_PX; ++j) {
          const float xv = xin[ky][j + kx][ci];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j][k] = fmaf(xv, wv[k], acc[j][k]);
        }
      }
#pragma unroll
  for (int j = 0; j < IMG_PX; ++j) {
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = apply_act(acc[j][k], act, slope);
    *reinterpret_cast<uint4*>(y + ((int64_t)(n * H + h) * W + w0 + j) * ldy + cg * 8) = pack8(o);
  }
}

// total over the 8 lanes of an aligned lane group: lane l (0..7) returns the group total of v[l]   (7 shuffles)
__device__ __forceinline__ float butterfly8(const float (&v)[8], int l) {
  float t[4], u[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = (l & 4) ? v[i + 4] : v[i], send = (l & 4) ? v[i] : v[i + 4];
    t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = (l & 2) ? t[i + 2] : t[i], send = (l & 2) ? t[i] : t[i + 2];
    u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const float keep = (l & 1) ? u[1] : u[0], send = (l & 1) ? u[0] : u[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// One block = one image x one band of IMG_BAND rows (+ one halo row above and below for the data gradient).
// da: gradient w.r.t. the activated output, a: the activated output (its sign gives the LeakyReLU derivative; act NONE:
// dz = da), x: the layer input.  partial: [gridDim.x][9*C][64] fp32 weight-gradient partial sums.  dx: data gradient
// (C real channels at pixel stride lddx) or NULL.
template <int C>
__global__ void __launch_bounds__(256, 1)
img_conv3x3_bwd_kernel(const __nv_bfloat16* __restrict__ da, int64_t ldda, const __nv_bfloat16* __restrict__ a, int64_t lda,
                       const __nv_bfloat16* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                       int N, int H, int W, int act, float slope, float* __restrict__ partial, __nv_bfloat16* __restrict__ dx,
                       int64_t lddx) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NT = 9 * C;                              // (tap, ci) pairs
  float* ws = sm;                                        // [NT][64]
  float* P = sm + NT * IMG_CO;                           // [(BAND+2) * W][NT]   (re-used as the wgrad reduction buffer)
  load_weights<C>(ws, w, s_l, s_s);
  __syncthreads();
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int cg = threadIdx.x % 8, pg = threadIdx.x / 8;
  const int bands = H / IMG_BAND;
  const int n = blockIdx.x / bands, r0 = (blockIdx.x % bands) * IMG_BAND;
  const int gpr = W / IMG_PX;
  const int ngroups = (IMG_BAND + 2) * gpr;
  float acc[NT][8];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;

  for (int g0 = 0; g0 < ngroups; g0 += 32) {             // warp-uniform trip count (the shuffles below need every lane)
    const int g = g0 + pg;
    const int prow = g / gpr;                            // row inside the P tile: 0 = halo above the band
    const int h = r0 - 1 + prow, w0 = (g % gpr) * IMG_PX;
    const bool live = g < ngroups && h >= 0 && h < H;
    const bool inband = live && prow >= 1 && prow <= IMG_BAND;
    float dz[IMG_PX][8];
#pragma unroll
    for (int j = 0; j < IMG_PX; ++j) {
      uint4 qd = make_uint4(0, 0, 0, 0), qa = make_uint4(0, 0, 0, 0);
      if (live) {
        const int64_t pix = (int64_t)(n * H + h) * W + w0 + j;
        qd = *reinterpret_cast<const uint4*>(da + pix * ldda + cg * 8);
        if (act != DCV_ACT_NONE) qa = *reinterpret_cast<const uint4*>(a + pix * lda + cg * 8);
      }
      float fa[8];
      unpack8(qd, dz[j]);
      unpack8(qa, fa);
      if (act != DCV_ACT_NONE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dz[j][k] *= act_grad_from_out(fa[k], act, slope);
      }
    }
    // ---- data gradient: per-pixel partial products, reduced over the 8 lanes that share the pixel
    if (dx != nullptr) {
#pragma unroll
      for (int j = 0; j < IMG_PX; ++j) {
        float part[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const float4* wp = reinterpret_cast<const float4*>(ws + i * IMG_CO + cg * 8);
          const float4 wa = wp[0], wb = wp[1];
          part[i] = dz[j][0] * wa.x + dz[j][1] * wa.y + dz[j][2] * wa.z + dz[j][3] * wa.w + dz[j][4] * wb.x + dz[j][5] * wb.y +
                    dz[j][6] * wb.z + dz[j][7] * wb.w;
        }
        float* prow_p = P + (size_t)(prow * W + w0 + j) * NT;
#pragma unroll
        for (int b = 0; b + 8 <= NT; b += 8) {
          float v8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v8[i] = part[b + i];
          const float tot = butterfly8(v8, cg);
          if (g < ngroups) prow_p[b + cg] = tot;
        }
#pragma unroll
        for (int i = NT / 8 * 8; i < NT; ++i) {
          const float tot = group8_sum(part[i]);
          if (cg == 0 && g < ngroups) prow_p[i] = tot;
        }
      }
    }
    // ---- weight gradient: rows of the band only
    if (inband) {
      float xin[3][IMG_PX + 2][C];
      load_window<C>(x, ldx, n, h, w0, H, W, xin);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ci = 0; ci < C; ++ci)
#pragma unroll
            for (int j = 0; j < IMG_PX; ++j) {
              const float xv = xin[ky][j + kx][ci];
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[(ky * 3 + kx) * C + ci][k] = fmaf(xv, dz[j][k], acc[(ky * 3 + kx) * C + ci][k]);
            }
    }
  }
  __syncthreads();
  // ---- data gradient of the band: dx[h][w][ci] = sum_taps P[h - ky + 1][w - kx + 1][tap][ci]
  if (dx != nullptr) {
    for (int p = threadIdx.x; p < IMG_BAND * W; p += blockDim.x) {
      const int hb = p / W, wq = p % W;                  // band-local row, column
      float s[C];
#pragma unroll
      for (int ci = 0; ci < C; ++ci) s[ci] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int pr = hb + 1 - (ky - 1), pc = wq - (kx - 1);   // P row (tile-local: band row hb is P row hb + 1), column
          if (pc < 0 || pc >= W) continue;
          const float* q = P + (size_t)(pr * W + pc) * NT + (ky * 3 + kx) * C;
#pragma unroll
          for (int ci = 0; ci < C; ++ci) s[ci] += q[ci];
        }
      __nv_bfloat16* d = dx + ((int64_t)(n * H + r0 + hb) * W + wq) * lddx;
#pragma unroll
      for (int ci = 0; ci < C; ++ci) d[ci] = __float2bfloat16_rn(s[ci]);
    }
    __syncthreads();
  }
  // ---- weight gradient: sum over the 4 pixel groups of a warp (lanes l, l^8, l^16), then over the 8 warps through smem
  float* red = P;                                        // [8 warps][NT][64]
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = acc[i][k];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) red[(warp * NT + i) * IMG_CO + cg * 8 + k] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < NT * IMG_CO; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) s += red[wq * NT * IMG_CO + i];
    partial[(size_t)blockIdx.x * NT * IMG_CO + i] = s;
  }
}

// ------------------------------------------------------------------------------------------ host side
// geometry test: Conv2d(k3, s1, p1) with <= 2 real input channels and exactly 64 output channels, bf16
int img_conv_supported(const dcv_geom* g) {
  const int wcl = g->wCl > 0 ? g->wCl : g->Cl, wcs = g->wCs > 0 ? g->wCs : g->Cs;
  if (g->kt != 1 || g->kh != 3 || g->kw != 3 || g->st != 1 || g->sh != 1 || g->sw != 1 || g->pt != 0 || g->ph != 1 || g->pw != 1) return 0;
  if (g->Tl != 1 || g->Ts != 1 || g->Hl != g->Hs || g->Wl != g->Ws) return 0;
  if (wcl < 1 || wcl > 2 || wcs != IMG_CO || g->Cs != IMG_CO) return 0;
  if (g->Wl % IMG_PX || g->Hl % IMG_BAND || g->Wl > 64) return 0;
  return 1;
}

static int bwd_smem_bytes(int C, int W) {
  const int nt = 9 * C;
  int tile = (IMG_BAND + 2) * W * nt, red = 8 * nt * IMG_CO;
  return (nt * IMG_CO + (tile > red ? tile : red)) * (int)sizeof(float);
}

int img_conv_bwd_blocks(const dcv_geom* g) { return g->N * (g->Hl / IMG_BAND); }

int64_t img_conv_bwd_ws_bytes(const dcv_geom* g) {
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  return (int64_t)img_conv_bwd_blocks(g) * 9 * C * IMG_CO * sizeof(float);
}

int img_conv_fwd(const dcv_geom* g, const void* x, int64_t ldx, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, void* y,
                 int64_t ldy, int act, float slope, cudaStream_t s) {
  DCV_REQUIRE(img_conv_supported(g), "img_conv_fwd: geometry not supported");
  DCV_REQUIRE(s_tap == 1, "img_conv_fwd: taps of the master weight must be contiguous");
  DCV_REQUIRE((((uintptr_t)y) & 15) == 0 && ldy % 8 == 0, "img_conv_fwd: output must be 16-byte aligned");
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  DCV_REQUIRE(C == 1 || ((((uintptr_t)x) & 3) == 0 && ldx % 2 == 0), "img_conv_fwd: 2-channel input must be 4-byte aligned");
  const int64_t groups = (int64_t)g->N * g->Hl * (g->Wl / IMG_PX);
  const unsigned blocks = (unsigned)((groups + 31) / 32);
  if (C == 1)
    img_conv3x3_fwd_kernel<1><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, w, s_l, s_s, g->N, g->Hl, g->Wl, (__nv_bfloat16*)y, ldy, act, slope);
  else
    img_conv3x3_fwd_kernel<2><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, w, s_l, s_s, g->N, g->Hl, g->Wl, (__nv_bfloat16*)y, ldy, act, slope);
  return check_launch("img_conv3x3_fwd");
}

int img_conv_bwd(const dcv_geom* g, const void* da, int64_t ldda, const void* a, int64_t lda, const void* x, int64_t ldx,
                 const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, int act, float slope, float* dw, int accumulate,
                 void* dx, int64_t lddx, void* ws, int64_t ws_bytes, cudaStream_t s) {
  DCV_REQUIRE(img_conv_supported(g), "img_conv_bwd: geometry not supported");
  DCV_REQUIRE(s_tap == 1, "img_conv_bwd: taps of the master weight must be contiguous");
  DCV_REQUIRE(act == DCV_ACT_NONE || act == DCV_ACT_LEAKY, "img_conv_bwd: activation %d", act);
  DCV_REQUIRE(ws_bytes >= img_conv_bwd_ws_bytes(g), "img_conv_bwd: workspace too small");
  DCV_REQUIRE((((uintptr_t)da) & 15) == 0 && ldda % 8 == 0 && (((uintptr_t)a) & 15) == 0 && lda % 8 == 0, "img_conv_bwd: da / a must be 16-byte aligned");
  const int C = g->wCl > 0 ? g->wCl : g->Cl;
  DCV_REQUIRE(C == 1 || ((((uintptr_t)x) & 3) == 0 && ldx % 2 == 0), "img_conv_bwd: 2-channel input must be 4-byte aligned");
  const int blocks = img_conv_bwd_blocks(g);
  const int smem = bwd_smem_bytes(C, g->Wl);
  static int smem_set[3] = {0, 0, 0};
  if (C == 1) {
    if (smem > smem_set[1]) { DCV_CUDA(cudaFuncSetAttribute(img_conv3x3_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); smem_set[1] = smem; }
    img_conv3x3_bwd_kernel<1><<<blocks, 256, smem, s>>>((const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)x, ldx,
                                                       w, s_l, s_s, g->N, g->Hl, g->Wl, act, slope, (float*)ws, (__nv_bfloat16*)dx, lddx);
  } else {
    if (smem > smem_set[2]) { DCV_CUDA(cudaFuncSetAttribute(img_conv3x3_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); smem_set[2] = smem; }
    img_conv3x3_bwd_kernel<2><<<blocks, 256, smem, s>>>((const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)x, ldx,
                                                       w, s_l, s_s, g->N, g->Hl, g->Wl, act, slope, (float*)ws, (__nv_bfloat16*)dx, lddx);
  }
  if (int rc = check_launch("img_conv3x3_bwd")) return rc;
  if (!dw) return 0;
  // partial sums [blocks][9][C][64] == the split layout of wgrad_reduce for a geometry with Cl = C real channels
  dcv_geom g2 = *g;
  g2.Cl = C; g2.wCl = C;
  return wgrad_reduce((const float*)ws, blocks, &g2, dw, s_l, s_s, s_tap, accumulate, s);
}

}  // namespace dcv
