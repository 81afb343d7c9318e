// HBM-bound layers of the DCVGAN step: BatchNorm statistics / apply / backward, activations,
// Dropout2d scaling, Noise, softmax, layout conversion, temporal difference.
// Everything is channels-last with an explicit pixel stride, fp32 math, fp32 or bf16 storage.
#include "common.cuh"
#include <stdlib.h>
#include <initializer_list>
#include <string.h>

namespace dcv {

static inline int ew_blocks(int64_t total, int per_thread = 1) {
  int64_t b = (total + 256 * per_thread - 1) / (256 * per_thread);
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------
// In-kernel finalize of a per-channel reduction ("last block done").  The reduction kernels below leave one partial
// row [2][C] per block; turning those into mean / invstd (forward) or sum(du), sum(du*xhat), dgamma, dbeta (backward)
// used to be a second launch of a few microseconds per BatchNorm site - 72 launches per training iteration.  With a
// BnTail the blocks finish the job themselves, hierarchically and in a FIXED summation order (deterministic):
//   * blocks are grouped by 16; the block of a group that arrives last (atomic ticket) adds the group's rows, in row
//     order, into group_ws[group][2][C];
//   * the group-finishing block that arrives last overall adds the group rows in double precision, in group order, and
//     writes the final per-channel values (and updates the running statistics exactly like bn_finalize_kernel).
// Tickets are reset by the block that consumes them, so the same counter buffer serves every call on a stream.
constexpr int BN_TAIL_GROUP = 16;
constexpr int BN_TAIL_MAX_GROUPS = 63;
struct BnTail {
  unsigned* counters;       // [1 + BN_TAIL_MAX_GROUPS] zero-initialised, self-resetting; NULL = no in-kernel finalize
  float* group_ws;          // [ngroups][2][C]
  int backward;             // 0: BatchNorm statistics, 1: BatchNorm backward sums
  double count; float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches_tracked; float* mean; float* invstd;
  float* sums; float* dgamma; float* dbeta; int accumulate;
};
static inline BnTail no_tail() { BnTail t; memset(&t, 0, sizeof(t)); return t; }

__device__ __forceinline__ void bn_tail(const float* __restrict__ partials, int C, const BnTail& t) {
  if (t.counters == nullptr) return;
  __shared__ int s_role;
  const int nblk = (int)gridDim.x, tid = (int)threadIdx.x;
  const int group = (int)blockIdx.x / BN_TAIL_GROUP, ngroups = (nblk + BN_TAIL_GROUP - 1) / BN_TAIL_GROUP;
  const int g0 = group * BN_TAIL_GROUP, gsize = nblk - g0 < BN_TAIL_GROUP ? nblk - g0 : BN_TAIL_GROUP;
  __threadfence();                                   // this block's partial row is visible device-wide ...
  __syncthreads();
  if (tid == 0) s_role = atomicAdd(&t.counters[1 + group], 1u) == (unsigned)(gsize - 1);   // ... before its ticket is
  __syncthreads();
  if (!s_role) return;
  __threadfence();
  for (int i = tid; i < 2 * C; i += (int)blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < gsize; ++r) a += __ldcg(partials + (size_t)(g0 + r) * 2 * C + i);
    t.group_ws[(size_t)group * 2 * C + i] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    t.counters[1 + group] = 0u;
    s_role = atomicAdd(&t.counters[0], 1u) == (unsigned)(ngroups - 1);
  }
  __syncthreads();
  if (!s_role) return;
  __threadfence();
  for (int c = tid; c < C; c += (int)blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    for (int g = 0; g < ngroups; ++g) {
      s0 += (double)__ldcg(t.group_ws + (size_t)g * 2 * C + c);
      s1 += (double)__ldcg(t.group_ws + (size_t)g * 2 * C + C + c);
    }
    if (!t.backward) {
      const double m = s0 / t.count;
      double var = s1 / t.count - m * m;
      if (var < 0.0) var = 0.0;
      t.mean[c] = (float)m;
      t.invstd[c] = (float)(1.0 / sqrt(var + (double)t.eps));
      if (t.running_mean) {
        const double unbiased = t.count > 1.0 ? var * t.count / (t.count - 1.0) : var;
        t.running_mean[c] = (float)((1.0 - t.momentum) * t.running_mean[c] + t.momentum * m);
        t.running_var[c] = (float)((1.0 - t.momentum) * t.running_var[c] + t.momentum * unbiased);
      }
    } else {
      t.sums[c] = (float)s0; t.sums[C + c] = (float)s1;
      if (t.dbeta) t.dbeta[c] = t.accumulate ? t.dbeta[c] + (float)s0 : (float)s0;
      if (t.dgamma) t.dgamma[c] = t.accumulate ? t.dgamma[c] + (float)s1 : (float)s1;
    }
  }
  if (tid == 0) {
    t.counters[0] = 0u;
    if (!t.backward && t.num_batches_tracked) *t.num_batches_tracked += 1;      // nn.BatchNorm's counter
  }
}

// ------------------------------------------------------------------------------------------
// per-channel reductions.  Block b reduces rows [b*rpb, (b+1)*rpb) into partials[b][k][C].
// F(row, c, &v0, &v1) yields the two addends.
template <typename F>
__device__ __forceinline__ void channel_reduce2(int64_t rows, int C, float* __restrict__ partials, F f) {
  __shared__ float s0[256];
  __shared__ float s1[256];
  const int64_t rpb = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t rb = (int64_t)blockIdx.x * rpb;
  int64_t re = rb + rpb; if (re > rows) re = rows;
  float* out = partials + (int64_t)blockIdx.x * 2 * C;
  const int tid = threadIdx.x;
  if (C <= 256) {
    const int lanes = 256 / C;
    const int c = tid % C, r = tid / C;
    float a0 = 0.f, a1 = 0.f;
    if (r < lanes) {
      for (int64_t row = rb + r; row < re; row += lanes) {
        float v0, v1; f(row, c, v0, v1); a0 += v0; a1 += v1;
      }
    }
    s0[tid] = a0; s1[tid] = a1;
    __syncthreads();
    if (r == 0) {
      for (int rr = 1; rr < lanes; ++rr) { a0 += s0[rr * C + c]; a1 += s1[rr * C + c]; }
      out[c] = a0; out[C + c] = a1;
    }
  } else {
    for (int c = tid; c < C; c += 256) {
      float a0 = 0.f, a1 = 0.f;
      for (int64_t row = rb; row < re; ++row) { float v0, v1; f(row, c, v0, v1); a0 += v0; a1 += v1; }
      out[c] = a0; out[C + c] = a1;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ z, int64_t ldz, int64_t rows, int C,
                                                        float* __restrict__ partials, const BnTail tail) {
  pdl_wait(); pdl_trigger();
  channel_reduce2(rows, C, partials, [&](int64_t row, int c, float& v0, float& v1) {
    const float v = ldf(z + row * ldz + c); v0 = v; v1 = v * v;
  });
  bn_tail(partials, C, tail);
}

// one warp per channel: lanes stride over the partial blocks, double accumulation, fixed shuffle tree
__device__ __forceinline__ void warp_sum_partials(const float* __restrict__ partials, int nblk, int C, int c, double& s0, double& s1) {
  const int lane = threadIdx.x % 32;
  s0 = 0.0; s1 = 0.0;
  // eight independent loads per round trip (the partials of one channel are 2*C floats apart: every load is its own
  // sector, and a dependent load-add chain over up to 19 blocks per lane made these tiny kernels 5-9 us each)
  int b = lane;
  for (; b + 96 < nblk; b += 128) {
    float v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { v0[u] = partials[(int64_t)(b + 32 * u) * 2 * C + c]; v1[u] = partials[(int64_t)(b + 32 * u) * 2 * C + C + c]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { s0 += v0[u]; s1 += v1[u]; }
  }
  for (; b < nblk; b += 32) { s0 += partials[(int64_t)b * 2 * C + c]; s1 += partials[(int64_t)b * 2 * C + C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
}

// One block per 32 channels, 32 warps: lane = channel (coalesced 128-byte rows of the partials), warp w adds the partial rows
// w, w + 32, ... with four independent loads in flight, the warps are then added in warp order (fixed order: deterministic).
// The first version gave every channel ONE warp whose lanes strode over the rows - every load its own 32-byte sector and ~19
// dependent round trips per lane: 8.9 us per launch for the 592 slots of a convolution epilogue, 42 launches per iteration.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partials, int nblk, int C, double count, float eps,
                   float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                   float* __restrict__ mean, float* __restrict__ invstd) {
  pdl_wait(); pdl_trigger();
  __shared__ double red[2][32][33];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + lane;
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;   // nn.BatchNorm's counter
  double s0 = 0.0, s1 = 0.0;
  if (c < C) {
    int b = warp;
    for (; b + 96 < nblk; b += 128) {
      float v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { v0[u] = partials[(int64_t)(b + 32 * u) * 2 * C + c]; v1[u] = partials[(int64_t)(b + 32 * u) * 2 * C + C + c]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s0 += v0[u]; s1 += v1[u]; }
    }
    for (; b < nblk; b += 32) { s0 += partials[(int64_t)b * 2 * C + c]; s1 += partials[(int64_t)b * 2 * C + C + c]; }
  }
  red[0][warp][lane] = s0; red[1][warp][lane] = s1;
  __syncthreads();
  if (warp != 0 || c >= C) return;
  s0 = 0.0; s1 = 0.0;
  for (int w = 0; w < 32; ++w) { s0 += red[0][w][lane]; s1 += red[1][w][lane]; }
  const double m = s0 / count;
  double var = s1 / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  pdl_wait(); pdl_trigger();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  invstd[c] = 1.0f / sqrtf(rv[c] + eps);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_kernel(const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
              const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
              const float* __restrict__ drop, int64_t rows_per_n, int act, float slope, T* __restrict__ a, int64_t lda) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    float v = ldf(z + row * ldz + c);
    if (mean) {
      v = (v - mean[c]) * invstd[c];
      if (gamma) v = v * gamma[c] + beta[c];
    }
    if (drop) v *= drop[(row / rows_per_n) * C + c];
    stf(a + row * lda + c, apply_act(v, act, slope));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda,
                         const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ drop, int64_t rows_per_n,
                         int act, float slope, float* __restrict__ partials, const BnTail tail) {
  pdl_wait(); pdl_trigger();
  channel_reduce2(rows, C, partials, [&](int64_t row, int c, float& v0, float& v1) {
    float du = ldf(da + row * ldda + c) * act_grad_from_out(ldf(a + row * lda + c), act, slope);
    if (drop) du *= drop[(row / rows_per_n) * C + c];
    const float xhat = (ldf(z + row * ldz + c) - mean[c]) * invstd[c];
    v0 = du; v1 = du * xhat;
  });
  bn_tail(partials, C, tail);
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int nblk, int C, float* __restrict__ sums,
                                       float* dgamma, float* dbeta, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  if (c >= C) return;
  double s0, s1;
  warp_sum_partials(partials, nblk, C, c, s0, s1);
  if (threadIdx.x % 32 != 0) return;
  sums[c] = (float)s0; sums[C + c] = (float)s1;
  if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s0 : (float)s0;
  if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)s1 : (float)s1;
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda,
                        const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ drop, int64_t rows_per_n, int act, float slope,
                        const float* __restrict__ sums, float inv_count, T* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    float du = ldf(da + row * ldda + c) * act_grad_from_out(ldf(a + row * lda + c), act, slope);
    if (drop) du *= drop[(row / rows_per_n) * C + c];
    const float is = invstd[c];
    const float xhat = (ldf(z + row * ldz + c) - mean[c]) * is;
    const float g = gamma ? gamma[c] : 1.f;
    const float v = g * is * (du - sums[c] * inv_count - xhat * sums[C + c] * inv_count);
    stf(dz + row * lddz + c, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda, int64_t rows, int C,
               int act, float slope, T* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    stf(dz + row * lddz + c, ldf(da + row * ldda + c) * act_grad_from_out(ldf(a + row * lda + c), act, slope));
  }
}

// planar fp32 (N, C, T, H, W) -> channels-last bf16, one thread per pixel: the reads of a warp are contiguous in every channel plane
__global__ void __launch_bounds__(256)
to_cl_px_bf16_kernel(const float* __restrict__ src, int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw, int N, int C,
                     int Tn, int H, int W, __nv_bfloat16* __restrict__ dst, int64_t ld) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * Tn * H * W;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(r % W); int64_t q = r / W;
    const int h = (int)(q % H); q /= H;
    const int t = (int)(q % Tn); const int n = (int)(q / Tn);
    const float* p = src + n * sn + t * st + h * sh + w * sw;
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = c < C ? p[c * sc] : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) dst[r * ld + c] = __float2bfloat16_rn(v[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
add_noise_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ noise, float sigma, int64_t rows,
                 int C, T* __restrict__ out, int64_t ldo) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    stf(out + row * ldo + c, ldf(x + row * ldx + c) + sigma * noise[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
axpy_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int C, T* __restrict__ out, int64_t ldo, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    float v = ldf(x + row * ldx + c);
    if (accumulate) v += ldf(out + row * ldo + c);
    stf(out + row * ldo + c, v);
  }
}

// out[n, t, p, c] = x[n, t+1, p, c] - x[n, t, p, c], t in [0, T-1)
template <typename T>
__global__ void __launch_bounds__(256)
tdiff_kernel(const T* __restrict__ x, int64_t ldx, int N, int Tn, int64_t hw, int C, T* __restrict__ out, int64_t ldo) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * (Tn - 1) * hw * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t r = i / C;
    const int64_t p = r % hw; r /= hw;
    const int t = (int)(r % (Tn - 1)); const int n = (int)(r / (Tn - 1));
    const int64_t src = ((int64_t)n * Tn + t) * hw + p;
    const float v = ldf(x + (src + hw) * ldx + c) - ldf(x + src * ldx + c);
    stf(out + (((int64_t)n * (Tn - 1) + t) * hw + p) * ldo + c, v);
  }
}

// dx[n,t] (+)= dy[n,t-1] - dy[n,t]  (terms dropped where out of range)
template <typename T>
__global__ void __launch_bounds__(256)
tdiff_bwd_kernel(const T* __restrict__ dy, int64_t lddy, int N, int Tn, int64_t hw, int C, T* __restrict__ dx,
                 int64_t lddx, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * Tn * hw * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t r = i / C;
    const int64_t p = r % hw; r /= hw;
    const int t = (int)(r % Tn); const int n = (int)(r / Tn);
    float v = 0.f;
    if (t >= 1) v += ldf(dy + (((int64_t)n * (Tn - 1) + (t - 1)) * hw + p) * lddy + c);
    if (t < Tn - 1) v -= ldf(dy + (((int64_t)n * (Tn - 1) + t) * hw + p) * lddy + c);
    T* d = dx + (((int64_t)n * Tn + t) * hw + p) * lddx + c;
    if (accumulate) v += ldf(d);
    stf(d, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_kernel(const T* __restrict__ z, int64_t ldz, int64_t rows, int C, T* __restrict__ y, int64_t ldy) {
  pdl_wait(); pdl_trigger();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const T* zr = z + row * ldz; T* yr = y + row * ldy;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, ldf(zr + c));
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(ldf(zr + c) - mx);
    const float inv = 1.f / s;
    for (int c = 0; c < C; ++c) stf(yr + c, expf(ldf(zr + c) - mx) * inv);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ y, int64_t ldy, int64_t rows, int C,
                   T* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const T* dyr = dy + row * lddy; const T* yr = y + row * ldy; T* dzr = dz + row * lddz;
    float dot = 0.f;
    for (int c = 0; c < C; ++c) dot += ldf(dyr + c) * ldf(yr + c);
    for (int c = 0; c < C; ++c) stf(dzr + c, ldf(yr + c) * (ldf(dyr + c) - dot));
  }
}

// torch.argmax returns the first maximal index; scatter +1 there, -1 elsewhere
template <typename T>
__global__ void __launch_bounds__(256)
segm_remap_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int C, T* __restrict__ y, int64_t ldy) {
  pdl_wait(); pdl_trigger();
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (int64_t)gridDim.x * blockDim.x) {
    const T* xr = x + row * ldx; T* yr = y + row * ldy;
    int best = 0; float bv = ldf(xr);
    for (int c = 1; c < C; ++c) { const float v = ldf(xr + c); if (v > bv) { bv = v; best = c; } }
    for (int c = 0; c < C; ++c) stf(yr + c, c == best ? 1.f : -1.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
to_cl_kernel(const float* __restrict__ src, int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw, int N, int C,
             int Tn, int H, int W, T* __restrict__ dst, int64_t ld) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * Tn * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t r = i / C;
    const int w = (int)(r % W); int64_t q = r / W;
    const int h = (int)(q % H); q /= H;
    const int t = (int)(q % Tn); const int n = (int)(q / Tn);
    stf(dst + r * ld + c, src[n * sn + c * sc + t * st + h * sh + w * sw]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
from_cl_kernel(const T* __restrict__ src, int64_t ld, int N, int C, int Tn, int H, int W, float* __restrict__ dst,
               int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * Tn * H * W * C;
  // iterate with w fastest, then h, t, c, n: the usual dense NC(T)HW destination is then written coalesced
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W); int64_t q = i / W;
    const int h = (int)(q % H); q /= H;
    const int t = (int)(q % Tn); q /= Tn;
    const int c = (int)(q % C); const int n = (int)(q / C);
    const int64_t row = (((int64_t)n * Tn + t) * H + h) * W + w;
    float v = ldf(src + row * ld + c);
    float* d = dst + n * sn + c * sc + t * st + h * sh + w * sw;
    *d = accumulate ? *d + v : v;
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
copy_cl_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd, int64_t rows, int C) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    stf(dst + row * ldd + c, ldf(src + row * lds + c));
  }
}

// narrow copies (C <= 8: image-like tensors into / out of channel slices): one thread per pixel, no per-element division
template <typename TS, typename TD>
__global__ void __launch_bounds__(256)
copy_cl_rows_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd, int64_t rows, int C) {
  pdl_wait(); pdl_trigger();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const TS* sp = src + r * lds; TD* dp = dst + r * ldd;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < C) v[c] = ldf(sp + c);
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < C) stf(dp + c, v[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
frame_copy_kernel(T* __restrict__ clips, int64_t ldc, int N, int Tn, int64_t hw, int C, int t, const int* __restrict__ t_dev,
                  T* __restrict__ frames, int64_t ldfr, int reverse, int accumulate) {
  pdl_wait(); pdl_trigger();
  if (t_dev) t = min(max(*t_dev, 0), Tn - 1);   // frame index kept on the device (CUDA-graph replay)
  const int64_t total = (int64_t)N * hw * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t r = i / C;
    const int64_t p = r % hw; const int n = (int)(r / hw);
    T* cp = clips + (((int64_t)n * Tn + t) * hw + p) * ldc + c;
    T* fp = frames + r * ldfr + c;
    if (!reverse) stf(fp, ldf(cp));
    else stf(cp, accumulate ? ldf(cp) + ldf(fp) : ldf(fp));
  }
}


// ------------------------------------------------------------------------------------------
// 16-byte vectorised variants (8 bf16 / 4 fp32 channels per thread).  Thread -> (channel group cg, row lane);
// a block walks its rows with the per-channel constants held in registers, so every global access is a
// coalesced 16-byte load/store.  Used when C and every pixel stride are multiples of the vector width and all
// pointers are 16-byte aligned; the scalar kernels above remain the general case.
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
};
template <> struct VecIO<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
  }
};

struct VecMap {           // thread -> (cg, lane) mapping shared by the vector kernels
  int cg, lane, lanes; bool active;
  int64_t rb, re;
};
template <int N>
__device__ __forceinline__ VecMap vec_map(int64_t rows, int C) {
  VecMap m;
  const int CG = C / N;
  m.lanes = 256 / CG;
  m.cg = threadIdx.x % CG; m.lane = threadIdx.x / CG;
  m.active = m.lane < m.lanes;
  const int64_t rpb = (rows + gridDim.x - 1) / gridDim.x;
  m.rb = (int64_t)blockIdx.x * rpb;
  m.re = m.rb + rpb; if (m.re > rows) m.re = rows;
  return m;
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_vec_kernel(const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                  const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  const float* __restrict__ drop, int64_t rows_per_n, int act, float slope, T* __restrict__ a, int64_t lda) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  const VecMap m = vec_map<N>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * N;
  float sc[N], sh[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    sc[j] = 1.f; sh[j] = 0.f;
    if (mean) {
      const float is = invstd[c0 + j];
      const float g = gamma ? gamma[c0 + j] : 1.f;
      sc[j] = is * g; sh[j] = (gamma ? beta[c0 + j] : 0.f) - mean[c0 + j] * is * g;
    }
  }
  float mu[N], is[N], ga[N], be[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    mu[j] = mean ? mean[c0 + j] : 0.f; is[j] = mean ? invstd[c0 + j] : 1.f;
    ga[j] = (mean && gamma) ? gamma[c0 + j] : 1.f; be[j] = (mean && gamma) ? beta[c0 + j] : 0.f;
  }
  constexpr int U = 4;   // rows in flight per thread (independent 16-byte loads)
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    float v[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row < m.re) VecIO<T>::load(z + row * ldz + c0, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      const float* dr = drop ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float t = v[u][j];
        if (mean) { t = (t - mu[j]) * is[j]; if (gamma) t = t * ga[j] + be[j]; }   // same operation order as the scalar kernel
        if (dr) t *= dr[j];
        v[u][j] = apply_act(t, act, slope);
      }
      VecIO<T>::store(a + row * lda + c0, v[u]);
    }
  }
  (void)sc; (void)sh;
}

// partials[b][0][c] = sum f0, partials[b][1][c] = sum f1 over the block's rows; F fills two N-vectors per row
template <typename T, typename F>
__device__ __forceinline__ void channel_reduce2_vec(int64_t rows, int C, float* __restrict__ partials, F f) {
  constexpr int N = VecIO<T>::N;
  __shared__ float red[2][256][N + 1];
  const VecMap m = vec_map<N>(rows, C);
  float a0[N], a1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (m.active) {
    int64_t row = m.rb + m.lane;
    for (; row + m.lanes < m.re; row += 2 * (int64_t)m.lanes) {      // two rows in flight
      float v0[N], v1[N], u0[N], u1[N];
      f(row, m.cg * N, v0, v1);
      f(row + m.lanes, m.cg * N, u0, u1);
#pragma unroll
      for (int j = 0; j < N; ++j) { a0[j] += v0[j]; a1[j] += v1[j]; }
#pragma unroll
      for (int j = 0; j < N; ++j) { a0[j] += u0[j]; a1[j] += u1[j]; }
    }
    for (; row < m.re; row += m.lanes) {
      float v0[N], v1[N];
      f(row, m.cg * N, v0, v1);
#pragma unroll
      for (int j = 0; j < N; ++j) { a0[j] += v0[j]; a1[j] += v1[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) { red[0][threadIdx.x][j] = a0[j]; red[1][threadIdx.x][j] = a1[j]; }
  __syncthreads();
  const int CG = C / N;
  if (m.active && m.lane == 0) {
    float* out = partials + (int64_t)blockIdx.x * 2 * C;
    for (int l = 1; l < m.lanes; ++l) {
#pragma unroll
      for (int j = 0; j < N; ++j) { a0[j] += red[0][l * CG + m.cg][j]; a1[j] += red[1][l * CG + m.cg][j]; }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) { out[m.cg * N + j] = a0[j]; out[C + m.cg * N + j] = a1[j]; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_vec_kernel(const T* __restrict__ z, int64_t ldz, int64_t rows, int C,
                                                            float* __restrict__ partials, const BnTail tail) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  channel_reduce2_vec<T>(rows, C, partials, [&](int64_t row, int c0, float (&v0)[N], float (&v1)[N]) {
    VecIO<T>::load(z + row * ldz + c0, v0);
#pragma unroll
    for (int j = 0; j < N; ++j) v1[j] = v0[j] * v0[j];
  });
  bn_tail(partials, C, tail);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_vec_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda,
                             const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                             const float* __restrict__ invstd, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ drop, int64_t rows_per_n,
                             int act, float slope, float* __restrict__ partials, const BnTail tail) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  // (Leaky)ReLU: the sign of the activated output equals the sign of the pre-activation gamma * xhat + beta, which is
  // recomputed from z with exactly the forward's arithmetic - the output tensor `a` is then not read at all (one of
  // the three streamed tensors).  A dropped element (drop scale 0) has zero gradient whatever the sign says.
  const bool from_z = act == DCV_ACT_LEAKY;
  channel_reduce2_vec<T>(rows, C, partials, [&](int64_t row, int c0, float (&v0)[N], float (&v1)[N]) {
    float g[N], o[N], zz[N];
    VecIO<T>::load(da + row * ldda + c0, g);
    if (!from_z) VecIO<T>::load(a + row * lda + c0, o);
    VecIO<T>::load(z + row * ldz + c0, zz);
    const float* dr = drop ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float xhat = (zz[j] - mean[c0 + j]) * invstd[c0 + j];
      if (from_z) { float t = xhat; if (gamma) t = t * gamma[c0 + j] + beta[c0 + j]; o[j] = t; }
      float du = g[j] * act_grad_from_out(o[j], act, slope);
      if (dr) du *= dr[j];
      v0[j] = du; v1[j] = du * xhat;
    }
  });
  bn_tail(partials, C, tail);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_vec_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda,
                            const T* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                            const float* __restrict__ invstd, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const float* __restrict__ drop, int64_t rows_per_n, int act,
                            float slope, const float* __restrict__ sums, float inv_count, T* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  const VecMap m = vec_map<N>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * N;
  const bool from_z = act == DCV_ACT_LEAKY;       // see bn_act_bwd_reduce_vec_kernel
  float mu[N], is[N], ga[N], be[N], gi[N], k0[N], k1[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    mu[j] = mean[c0 + j]; is[j] = invstd[c0 + j];
    ga[j] = gamma ? gamma[c0 + j] : 1.f; be[j] = gamma ? beta[c0 + j] : 0.f;
    gi[j] = ga[j] * is[j];
    k0[j] = sums[c0 + j] * inv_count; k1[j] = sums[C + c0 + j] * inv_count;
  }
  constexpr int U = 2;   // rows in flight per thread
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    float g[U][N], o[U][N], zz[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      VecIO<T>::load(da + row * ldda + c0, g[u]);
      if (!from_z) VecIO<T>::load(a + row * lda + c0, o[u]);
      VecIO<T>::load(z + row * ldz + c0, zz[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      const float* dr = drop ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float xhat = (zz[u][j] - mu[j]) * is[j];
        if (from_z) { float t = xhat; if (gamma) t = t * ga[j] + be[j]; o[u][j] = t; }
        float du = g[u][j] * act_grad_from_out(o[u][j], act, slope);
        if (dr) du *= dr[j];
        g[u][j] = gi[j] * (du - k0[j] - xhat * k1[j]);
      }
      VecIO<T>::store(dz + row * lddz + c0, g[u]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const T* __restrict__ da, int64_t ldda, const T* __restrict__ a, int64_t lda, int64_t rows, int C,
                   int act, float slope, T* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  const VecMap m = vec_map<N>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * N;
  constexpr int U = 4;   // rows in flight per thread
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    uint4 rg[U], ro[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row < m.re) {
        rg[u] = *reinterpret_cast<const uint4*>(da + row * ldda + c0);
        ro[u] = *reinterpret_cast<const uint4*>(a + row * lda + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      float g[N], o[N];
      VecIO<T>::unpack(rg[u], g); VecIO<T>::unpack(ro[u], o);
#pragma unroll
      for (int j = 0; j < N; ++j) g[j] *= act_grad_from_out(o[j], act, slope);
      VecIO<T>::store(dz + row * lddz + c0, g);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
axpy_vec_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int C, T* __restrict__ out, int64_t ldo, int accumulate) {
  pdl_wait(); pdl_trigger();
  constexpr int N = VecIO<T>::N;
  const VecMap m = vec_map<N>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * N;
  constexpr int U = 4;   // rows in flight per thread
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    uint4 rx[U], ro[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row < m.re) {
        rx[u] = *reinterpret_cast<const uint4*>(x + row * ldx + c0);
        if (accumulate) ro[u] = *reinterpret_cast<const uint4*>(out + row * ldo + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      float v[N], o[N];
      VecIO<T>::unpack(rx[u], v);
      if (accumulate) {
        VecIO<T>::unpack(ro[u], o);
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] += o[j];
      }
      VecIO<T>::store(out + row * ldo + c0, v);
    }
  }
}

// ------------------------------------------------------------------------------------------ tap-unrolled transposed convolution
// Transposed convolutions with a tiny output channel count (outconv 128 -> 3, ggen main.12 64 -> C: generator.py:73,274)
// make poor tensor-core tiles: N = 16 padded output columns, and every tap re-reads its 128x16 A slice from shared
// memory.  They run instead as ONE 1x1 convolution x[pixels][Cin] x W[Cin][Cout*taps] -> P[pixels][(co, kh, kw)] (the
// transposed-convolution weight (Cin, Cout, kh, kw) already is that matrix) followed by this col2im gather:
//   y[n][oh][ow][co] = act( sum over (kh, kw) with oh = ih*s - p + kh, ow = iw*s - p + kw of P[n][ih][iw][co*taps + kh*KW + kw] )
// Tiled version (bf16, 16-byte aligned rows): a block owns a 16x16 output tile, stages the input pixels of P it depends
// on in shared memory with coalesced 16-byte loads (P is read from HBM exactly once; the per-pixel gather of
// Cout*taps 2-byte values then hits shared memory), one thread per output pixel.
template <int S>
__global__ void __launch_bounds__(256)
col2im_act_tiled_kernel(const __nv_bfloat16* __restrict__ P, int64_t ldp, int Ih, int Iw, int Cout, int KH, int KW, int pad,
                        int Oh, int Ow, int rowv /*16-byte vectors per staged row*/, int RH, int RW /*staged region*/, int TW /*tile width*/,
                        int act, float slope, __nv_bfloat16* __restrict__ y, int64_t ldy) {
  pdl_wait(); pdl_trigger();
  extern __shared__ uint32_t tile[];
  // staged pixel pitch = rowv*4 + 1 words: neighbouring output pixels read the same column of neighbouring staged pixels,
  // and a 64-byte pitch put 16 of them on 2 banks (ncu: 8 shared-memory wavefronts per load)
  const int pitch = rowv * 4 + 1;
  const int n = blockIdx.z, oh0 = blockIdx.y * 16, ow0 = blockIdx.x * TW;
  // first input row / column any output pixel of the tile can see: ih = (oh + pad - kh) / S, kh = KH-1 .. 0
  const int ih0 = max(0, (oh0 + pad - (KH - 1) + (S - 1)) / S), iw0 = max(0, (ow0 + pad - (KW - 1) + (S - 1)) / S);
  const int nvec = RH * RW * rowv;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const int v = i % rowv; const int r = i / rowv;
    const int rw = r % RW, rh = r / RW;
    const int ih = ih0 + rh, iw = iw0 + rw;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (ih < Ih && iw < Iw) val = *reinterpret_cast<const uint4*>(P + (((int64_t)n * Ih + ih) * Iw + iw) * ldp + v * 8);
    uint32_t* d = tile + r * pitch + v * 4;
    d[0] = val.x; d[1] = val.y; d[2] = val.z; d[3] = val.w;
  }
  __syncthreads();
  const int taps = KH * KW;
  // a 16-row x TW-column output tile per block (TW = 64 when the staged region fits: 4x fewer, 4x longer blocks - 8192 blocks of
  // 256 outputs each spent their time in launch / barrier latency, 58 us for the 64x64 frames of a batch of 32 clips)
  for (int o = threadIdx.x; o < 16 * TW; o += 256) {
    const int oh = oh0 + o / TW, ow = ow0 + o % TW;
    if (oh >= Oh || ow >= Ow) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < KH; ++kh) {
      const int th = oh + pad - kh;
      if (th < 0 || th % S) continue;
      const int ih = th / S;
      if (ih >= Ih) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tw = ow + pad - kw;
        if (tw < 0 || tw % S) continue;
        const int iw = tw / S;
        if (iw >= Iw) continue;
        const uint32_t* row = tile + ((ih - ih0) * RW + (iw - iw0)) * pitch;
        const int e0 = kh * KW + kw;
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          if (co < Cout) {
            const int e = e0 + co * taps;
            const uint32_t wv = row[e >> 1];
            acc[co] += __uint_as_float((e & 1) ? (wv & 0xFFFF0000u) : (wv << 16));
          }
        }
      }
    }
    __nv_bfloat16* yo = y + (((int64_t)n * Oh + oh) * Ow + ow) * ldy;
#pragma unroll
    for (int co = 0; co < 4; ++co) if (co < Cout) yo[co] = __float2bfloat16_rn(apply_act(acc[co], act, slope));
  }
}

template <typename T, int S>      // S = stride known at compile time (1, 2) or 0 = runtime
__global__ void __launch_bounds__(256)
col2im_act_kernel(const T* __restrict__ P, int64_t ldp, int N, int Ih, int Iw, int Cout, int KH, int KW, int s_rt, int pad, int Oh,
                  int Ow, int act, float slope, T* __restrict__ y, int64_t ldy) {
  pdl_wait(); pdl_trigger();
  const int s = S ? S : s_rt;
  const int taps = KH * KW;
  const int64_t total = (int64_t)N * Oh * Ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % Ow); const int64_t r = i / Ow;
    const int oh = (int)(r % Oh); const int n = (int)(r / Oh);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < KH; ++kh) {
      const int th = oh + pad - kh;
      if (th < 0 || th % s) continue;
      const int ih = th / s;
      if (ih >= Ih) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tw = ow + pad - kw;
        if (tw < 0 || tw % s) continue;
        const int iw = tw / s;
        if (iw >= Iw) continue;
        const T* row = P + (((int64_t)n * Ih + ih) * Iw + iw) * ldp + kh * KW + kw;
#pragma unroll
        for (int co = 0; co < 4; ++co) if (co < Cout) acc[co] += ldf(row + co * taps);
      }
    }
#pragma unroll
    for (int co = 0; co < 4; ++co) if (co < Cout) stf(y + i * ldy + co, apply_act(acc[co], act, slope));
  }
}

// ------------------------------------------------------------------------------------------ w-tap folding
// Image-like stem inputs (1-5 real channels): the kw taps of a strided convolution are moved into the channel dimension
// by a cheap pre-pass, X2[line][ow][k*cin + c] = x[line][ow*sw - pw + k][c] (zero outside the row), so that the
// convolution proper has kw = 1, kt*kh taps and kw*cin real channels per tap instead of kt*kh*kw taps of `cin` real
// channels in a 16-channel pitch: 4x fewer TMA boxes / MMAs for the 4x4(x4) stride-2 stems (they were bound by the
// TMA request rate of 32-byte rows, not by bytes).  cin = cg + cc channels come from two tensors (geometry | colour);
// the Noise layer of the image discriminator (x + sigma * N(0,1), fp32 noise dense [pixel][c]) is applied on the way.
template <typename T>
__global__ void __launch_bounds__(256)
fold_w_kernel(const T* __restrict__ xg, int64_t ldg, int cg, const float* __restrict__ ng, const T* __restrict__ xc, int64_t ldc,
              int cc, const float* __restrict__ nc, float sigma, int64_t lines, int W, int Ow, int kw, int sw, int pw,
              T* __restrict__ out, int64_t ldo) {
  pdl_wait(); pdl_trigger();
  const int cin = cg + cc;
  const int64_t total = lines * Ow;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t line = i / Ow; const int ow = (int)(i - line * Ow);
    T* o = out + i * ldo;
    if (sizeof(T) == 2 && kw == 4 && cin == 4 && ldo % 8 == 0 && ((uintptr_t)out & 15) == 0) {
      // 4 taps x 4 channels = one 32-byte row: build it in registers, two 16-byte stores (16 scalar 2-byte stores per thread
      // made this pass 56 us for the video stem)
      float v[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int iw = ow * sw - pw + k;
        const bool ok = iw >= 0 && iw < W;
        const int64_t px = line * W + iw;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float t = 0.f;
          if (ok) {
            if (c < cg) { t = ldf(xg + px * ldg + c); if (ng) t += sigma * ng[px * cg + c]; }
            else { t = ldf(xc + px * ldc + (c - cg)); if (nc) t += sigma * nc[px * cc + (c - cg)]; }
          }
          v[k * 4 + c] = t;
        }
      }
      uint32_t w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]); w[q] = *reinterpret_cast<uint32_t*>(&h); }
      uint4* dst = reinterpret_cast<uint4*>(o);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
      continue;
    }
    for (int k = 0; k < kw; ++k) {
      const int iw = ow * sw - pw + k;
      const bool ok = iw >= 0 && iw < W;
      const int64_t px = line * W + iw;
      for (int c = 0; c < cg; ++c) {
        float v = 0.f;
        if (ok) { v = ldf(xg + px * ldg + c); if (ng) v += sigma * ng[px * cg + c]; }
        stf(o + k * cin + c, v);
      }
      for (int c = 0; c < cc; ++c) {
        float v = 0.f;
        if (ok) { v = ldf(xc + px * ldc + c); if (nc) v += sigma * nc[px * cc + c]; }
        stf(o + k * cin + cg + c, v);
      }
    }
  }
}

// adjoint: dx[line][w][c] = sum over (k, ow) with ow*sw - pw + k == w of dX2[line][ow][k*cin + c]
template <typename T>
__global__ void __launch_bounds__(256)
unfold_w_kernel(const T* __restrict__ d2, int64_t ld2, int64_t lines, int W, int Ow, int kw, int sw, int pw, T* __restrict__ dxg,
                int64_t ldg, int cg, T* __restrict__ dxc, int64_t ldc, int cc) {
  pdl_wait(); pdl_trigger();
  const int cin = cg + cc;
  const int64_t total = lines * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t line = i / W; const int w = (int)(i - line * W);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    for (int k = 0; k < kw; ++k) {
      const int num = w + pw - k;
      if (num < 0 || num % sw) continue;
      const int ow = num / sw;
      if (ow >= Ow) continue;
      const T* src = d2 + (line * Ow + ow) * ld2 + k * cin;
#pragma unroll
      for (int c = 0; c < 8; ++c) if (c < cin) acc[c] += ldf(src + c);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < cg) stf(dxg + i * ldg + c, acc[c]);
      else if (c < cin) stf(dxc + i * ldc + (c - cg), acc[c]);
    }
  }
}

// ------------------------------------------------------------------------------------------ bf16 streaming variants
// The generic vector kernels above hold every loaded row as 8 unpacked floats plus four per-channel constant vectors:
// 95-127 registers, 2 CTAs per SM, ~32 KB of loads in flight per SM - they reached 2.8-3.0 TB/s where a plain copy
// runs at 6.0 TB/s (tools/exp_bn.py).  These bf16 variants keep the rows in flight as RAW 16-byte vectors (4
// registers each), unpack one row at a time, and fold the per-channel arithmetic into 2-5 fused constants:
//   xhat = z*A + B            A = invstd, B = -mean*invstd
//   pre  = z*P + Q            P = A*gamma, Q = B*gamma + beta      (sign of the (Leaky)ReLU argument)
//   dz   = G*du - R - S*z     G = gamma*invstd, R = G*(k0 + k1*B), S = G*k1*A   (k = reduced sums / count)
__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&h); }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256, 3)
bn_act_bf16_kernel(const __nv_bfloat16* __restrict__ z, int64_t ldz, int64_t rows, int C, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ drop, int64_t rows_per_n, int act, float slope, __nv_bfloat16* __restrict__ a,
                   int64_t lda) {
  pdl_wait(); pdl_trigger();
  const VecMap m = vec_map<8>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * 8;
  float P[8], Q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float is = mean ? invstd[c0 + j] : 1.f, mu = mean ? mean[c0 + j] : 0.f;
    const float g = (mean && gamma) ? gamma[c0 + j] : 1.f, b = (mean && gamma) ? beta[c0 + j] : 0.f;
    P[j] = is * g; Q[j] = b - mu * is * g;
  }
  constexpr int U = 8;
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row < m.re) raw[u] = *reinterpret_cast<const uint4*>(z + row * ldz + c0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      float v[8];
      unpack8(raw[u], v);
      const float* dr = drop ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(v[j], P[j], Q[j]);
        if (dr) t *= dr[j];
        v[j] = apply_act(t, act, slope);
      }
      *reinterpret_cast<uint4*>(a + row * lda + c0) = pack8(v);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
bn_act_bwd_apply_bf16_kernel(const __nv_bfloat16* __restrict__ da, int64_t ldda, const __nv_bfloat16* __restrict__ z, int64_t ldz,
                             int64_t rows, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                             const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ drop,
                             int64_t rows_per_n, float slope, const float* __restrict__ sums, float inv_count,
                             __nv_bfloat16* __restrict__ dz, int64_t lddz) {
  pdl_wait(); pdl_trigger();
  const VecMap m = vec_map<8>(rows, C);
  if (!m.active) return;
  const int c0 = m.cg * 8;
  float P[8], Q[8], G[8], R[8], S[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float A = invstd[c0 + j], B = -mean[c0 + j] * A;
    const float ga = gamma ? gamma[c0 + j] : 1.f, be = gamma ? beta[c0 + j] : 0.f;
    const float k0 = sums[c0 + j] * inv_count, k1 = sums[C + c0 + j] * inv_count;
    P[j] = A * ga; Q[j] = B * ga + be;
    G[j] = ga * A; R[j] = G[j] * (k0 + k1 * B); S[j] = G[j] * k1 * A;
  }
  constexpr int U = 4;
  for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
    uint4 rg[U], rz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row < m.re) {
        rg[u] = *reinterpret_cast<const uint4*>(da + row * ldda + c0);
        rz[u] = *reinterpret_cast<const uint4*>(z + row * ldz + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = row0 + (int64_t)u * m.lanes;
      if (row >= m.re) continue;
      float g[8], zz[8];
      unpack8(rg[u], g); unpack8(rz[u], zz);
      const float* dr = drop ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pre = fmaf(zz[j], P[j], Q[j]);
        float du = pre > 0.f ? g[j] : g[j] * slope;
        if (dr) du *= dr[j];
        g[j] = fmaf(G[j], du, -fmaf(S[j], zz[j], R[j]));
      }
      *reinterpret_cast<uint4*>(dz + row * lddz + c0) = pack8(g);
    }
  }
}

// block-level reduction tail shared by the two bf16 reduce kernels: partials[b][0][c], partials[b][1][c]
__device__ __forceinline__ void reduce2_tail(const VecMap& m, int C, float (&a0)[8], float (&a1)[8], float* __restrict__ partials) {
  // tree over the row lanes of each channel group (fixed order -> deterministic); a serial loop in the lane-0 threads
  // took ~7 us per block, the whole cost of the small layers' launches
  __shared__ float red[2][256][9];
  const int CG = C / 8;
  int span = 1;
  while (span < m.lanes) span <<= 1;
  for (int sft = span >> 1; sft >= 1; sft >>= 1) {
    if (m.active && m.lane >= sft && m.lane < 2 * sft) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { red[0][threadIdx.x][j] = a0[j]; red[1][threadIdx.x][j] = a1[j]; }
    }
    __syncthreads();
    if (m.active && m.lane < sft && m.lane + sft < m.lanes) {
      const int src = (m.lane + sft) * CG + m.cg;
#pragma unroll
      for (int j = 0; j < 8; ++j) { a0[j] += red[0][src][j]; a1[j] += red[1][src][j]; }
    }
    __syncthreads();
  }
  if (m.active && m.lane == 0) {
    float* out = partials + (int64_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) { out[m.cg * 8 + j] = a0[j]; out[C + m.cg * 8 + j] = a1[j]; }
  }
}

__global__ void __launch_bounds__(256, 3)
bn_stats_bf16_kernel(const __nv_bfloat16* __restrict__ z, int64_t ldz, int64_t rows, int C, float* __restrict__ partials,
                     const BnTail tail) {
  pdl_wait(); pdl_trigger();
  const VecMap m = vec_map<8>(rows, C);
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (m.active) {
    const int c0 = m.cg * 8;
    constexpr int U = 8;
    for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = row0 + (int64_t)u * m.lanes;
        raw[u] = row < m.re ? *reinterpret_cast<const uint4*>(z + row * ldz + c0) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[8];
        unpack8(raw[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
      }
    }
  }
  reduce2_tail(m, C, a0, a1, partials);
  bn_tail(partials, C, tail);
}

__global__ void __launch_bounds__(256, 2)
bn_act_bwd_reduce_bf16_kernel(const __nv_bfloat16* __restrict__ da, int64_t ldda, const __nv_bfloat16* __restrict__ z, int64_t ldz,
                              int64_t rows, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ drop,
                              int64_t rows_per_n, float slope, float* __restrict__ partials, const BnTail tail) {
  pdl_wait(); pdl_trigger();
  const VecMap m = vec_map<8>(rows, C);
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (m.active) {
    const int c0 = m.cg * 8;
    float A[8], B[8], P[8], Q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      A[j] = invstd[c0 + j]; B[j] = -mean[c0 + j] * A[j];
      const float ga = gamma ? gamma[c0 + j] : 1.f, be = gamma ? beta[c0 + j] : 0.f;
      P[j] = A[j] * ga; Q[j] = B[j] * ga + be;
    }
    constexpr int U = 4;
    for (int64_t row0 = m.rb + m.lane; row0 < m.re; row0 += (int64_t)m.lanes * U) {
      uint4 rg[U], rz[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = row0 + (int64_t)u * m.lanes;
        if (row < m.re) {
          rg[u] = *reinterpret_cast<const uint4*>(da + row * ldda + c0);
          rz[u] = *reinterpret_cast<const uint4*>(z + row * ldz + c0);
        } else {
          rg[u] = make_uint4(0u, 0u, 0u, 0u); rz[u] = rg[u];      // zero upstream gradient: contributes nothing
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t row = row0 + (int64_t)u * m.lanes;
        float g[8], zz[8];
        unpack8(rg[u], g); unpack8(rz[u], zz);
        const float* dr = (drop && row < m.re) ? drop + (row / rows_per_n) * C + c0 : nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = fmaf(zz[j], P[j], Q[j]);
          float du = pre > 0.f ? g[j] : g[j] * slope;
          if (dr) du *= dr[j];
          a0[j] += du; a1[j] = fmaf(du, fmaf(zz[j], A[j], B[j]), a1[j]);
        }
      }
    }
  }
  reduce2_tail(m, C, a0, a1, partials);
  bn_tail(partials, C, tail);
}

static inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }
template <typename T>
static inline bool vec_ok(int C, std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds) {
  constexpr int N = 16 / (int)sizeof(T);
  if (C % N != 0 || C / N > 256) return false;
  for (const void* p : ptrs) if (!al16(p)) return false;
  for (int64_t l : lds) if (l % N != 0) return false;
  return true;
}
static inline int vec_blocks(int64_t rows, int C, int N) {
  const int lanes = 256 / (C / N);
  int64_t b = (rows + (int64_t)lanes * 8 - 1) / ((int64_t)lanes * 8);   // ~8 rows per thread
  if (b > 148 * 8) b = 148 * 8;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace dcv

using namespace dcv;

#define DISPATCH_T(dtype, ...)                                 \
  do {                                                         \
    if ((dtype) == DCV_F32) { using T = float; __VA_ARGS__; }  \
    else { using T = __nv_bfloat16; __VA_ARGS__; }             \
  } while (0)

extern "C" {

int dcv_bn_stats_blocks(int64_t rows, int C) {
  (void)C;
  int64_t b = rows / 64;
  if (b > 148 * 4) b = 148 * 4;
  if (b < 1) b = 1;
  return (int)b;
}

static int launch_bn_stats(int dtype, const void* z, int64_t ldz, int64_t rows, int C, float* partials, const BnTail& tail, void* stream) {
  const int nblk = dcv_bn_stats_blocks(rows, C);
  DISPATCH_T(dtype, {
    if (sizeof(T) == 2 && vec_ok<T>(C, {z}, {ldz}))
      launch_k(bn_stats_bf16_kernel, nblk, 256, 0, as_stream(stream), (const __nv_bfloat16*)z, ldz, rows, C, partials, tail);
    else if (vec_ok<T>(C, {z}, {ldz})) launch_k(bn_stats_vec_kernel<T>, nblk, 256, 0, as_stream(stream), (const T*)z, ldz, rows, C, partials, tail);
    else launch_k(bn_stats_kernel<T>, nblk, 256, 0, as_stream(stream), (const T*)z, ldz, rows, C, partials, tail);
  });
  return check_launch("bn_stats");
}

int dcv_bn_stats(int dtype, const void* z, int64_t ldz, int64_t rows, int C, float* partials, void* stream) {
  return launch_bn_stats(dtype, z, ldz, rows, C, partials, no_tail(), stream);
}

int64_t dcv_bn_tail_workspace_bytes(int64_t rows, int C) {
  const int nblk = dcv_bn_stats_blocks(rows, C);
  const int ngroups = (nblk + BN_TAIL_GROUP - 1) / BN_TAIL_GROUP;
  return (int64_t)(nblk + ngroups) * 2 * C * sizeof(float);
}
int dcv_bn_tail_counters(void) { return 1 + BN_TAIL_MAX_GROUPS; }

static int tail_ws(int64_t rows, int C, float* ws, unsigned* counters, float** partials, BnTail* t) {
  DCV_REQUIRE(ws && counters, "bn (fused finalize): null workspace / counters");
  const int nblk = dcv_bn_stats_blocks(rows, C);
  const int ngroups = (nblk + BN_TAIL_GROUP - 1) / BN_TAIL_GROUP;
  DCV_REQUIRE(ngroups <= BN_TAIL_MAX_GROUPS, "bn (fused finalize): %d block groups", ngroups);
  *partials = ws;
  *t = no_tail();
  t->counters = counters;
  t->group_ws = ws + (size_t)nblk * 2 * C;
  return 0;
}

int dcv_bn_stats_finalize(int dtype, const void* z, int64_t ldz, int64_t rows, int C, float eps, float momentum,
                          float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                          void* ws, void* counters, void* stream) {
  float* partials; BnTail t;
  if (int rc = tail_ws(rows, C, (float*)ws, (unsigned*)counters, &partials, &t)) return rc;
  t.backward = 0; t.count = (double)rows; t.eps = eps; t.momentum = momentum;
  t.running_mean = running_mean; t.running_var = running_var; t.num_batches_tracked = (long long*)num_batches_tracked;
  t.mean = mean; t.invstd = invstd;
  return launch_bn_stats(dtype, z, ldz, rows, C, partials, t, stream);
}

int dcv_bn_finalize(const float* partials, int nblk, int C, int64_t count, float eps, float momentum,
                    float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd,
                    void* stream) {
  launch_k(bn_finalize_kernel, ceil_div(C, 32), 1024, 0, as_stream(stream), partials, nblk, C, (double)count, eps, momentum,
                                                                    running_mean, running_var, (long long*)num_batches_tracked,
                                                                    mean, invstd);
  return check_launch("bn_finalize");
}

int dcv_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean,
                      float* invstd, void* stream) {
  launch_k(bn_eval_stats_kernel, ceil_div(C, 128), 128, 0, as_stream(stream), running_mean, running_var, C, eps, mean, invstd);
  return check_launch("bn_eval_stats");
}

int dcv_bn_act(int dtype, const void* z, int64_t ldz, int64_t rows, int C, const float* mean, const float* invstd,
               const float* gamma, const float* beta, const float* drop, int64_t rows_per_n, int act, float slope,
               void* a, int64_t lda, void* stream) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, {
    if (sizeof(T) == 2 && vec_ok<T>(C, {z, a}, {ldz, lda}))
      launch_k(bn_act_bf16_kernel, vec_blocks(rows, C, 8), 256, 0, as_stream(stream), 
          (const __nv_bfloat16*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope, (__nv_bfloat16*)a, lda);
    else if (vec_ok<T>(C, {z, a}, {ldz, lda}))
      launch_k(bn_act_vec_kernel<T>, vec_blocks(rows, C, 16 / (int)sizeof(T)), 256, 0, as_stream(stream), 
          (const T*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope, (T*)a, lda);
    else
      launch_k(bn_act_kernel<T>, ew_blocks(rows * C, 4), 256, 0, as_stream(stream), 
          (const T*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope, (T*)a, lda);
  });
  return check_launch("bn_act");
}

static int launch_bn_bwd_reduce(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, const void* z,
                                int64_t ldz, int64_t rows, int C, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, const float* drop, int64_t rows_per_n, int act, float slope, float* partials,
                                const BnTail& tail, void* stream) {
  const int nblk = dcv_bn_stats_blocks(rows, C);
  DISPATCH_T(dtype, {
    if (sizeof(T) == 2 && act == DCV_ACT_LEAKY && vec_ok<T>(C, {da, z}, {ldda, ldz}))
      launch_k(bn_act_bwd_reduce_bf16_kernel, nblk, 256, 0, as_stream(stream), 
          (const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, slope,
          partials, tail);
    else if (vec_ok<T>(C, {da, a, z}, {ldda, lda, ldz}))
      launch_k(bn_act_bwd_reduce_vec_kernel<T>, nblk, 256, 0, as_stream(stream), 
          (const T*)da, ldda, (const T*)a, lda, (const T*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope,
          partials, tail);
    else
      launch_k(bn_act_bwd_reduce_kernel<T>, nblk, 256, 0, as_stream(stream), 
          (const T*)da, ldda, (const T*)a, lda, (const T*)z, ldz, rows, C, mean, invstd, drop, rows_per_n, act, slope, partials, tail);
  });
  return check_launch("bn_act_bwd_reduce");
}

int dcv_bn_act_bwd_reduce(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, const void* z,
                          int64_t ldz, int64_t rows, int C, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, const float* drop, int64_t rows_per_n, int act, float slope, float* partials,
                          void* stream) {
  return launch_bn_bwd_reduce(dtype, da, ldda, a, lda, z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope,
                              partials, no_tail(), stream);
}

int dcv_bn_act_bwd_reduce_finalize(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, const void* z,
                                   int64_t ldz, int64_t rows, int C, const float* mean, const float* invstd, const float* gamma,
                                   const float* beta, const float* drop, int64_t rows_per_n, int act, float slope, float* sums,
                                   float* dgamma, float* dbeta, int accumulate, void* ws, void* counters, void* stream) {
  float* partials; BnTail t;
  if (int rc = tail_ws(rows, C, (float*)ws, (unsigned*)counters, &partials, &t)) return rc;
  DCV_REQUIRE(sums, "bn_act_bwd_reduce_finalize: null sums");
  t.backward = 1; t.sums = sums; t.dgamma = dgamma; t.dbeta = dbeta; t.accumulate = accumulate;
  return launch_bn_bwd_reduce(dtype, da, ldda, a, lda, z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope,
                              partials, t, stream);
}

int dcv_bn_bwd_finalize(const float* partials, int nblk, int C, float* sums, float* dgamma, float* dbeta,
                        int accumulate, void* stream) {
  launch_k(bn_bwd_finalize_kernel, ceil_div(C, 8), 256, 0, as_stream(stream), partials, nblk, C, sums, dgamma, dbeta, accumulate);
  return check_launch("bn_bwd_finalize");
}

int dcv_bn_act_bwd_apply(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, const void* z,
                         int64_t ldz, int64_t rows, int C, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, const float* drop, int64_t rows_per_n, int act, float slope, const float* sums,
                         int64_t count, void* dz, int64_t lddz, void* stream) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, {
    if (sizeof(T) == 2 && act == DCV_ACT_LEAKY && vec_ok<T>(C, {da, z, dz}, {ldda, ldz, lddz}))
      launch_k(bn_act_bwd_apply_bf16_kernel, vec_blocks(rows, C, 8), 256, 0, as_stream(stream), 
          (const __nv_bfloat16*)da, ldda, (const __nv_bfloat16*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, slope,
          sums, 1.0f / (float)count, (__nv_bfloat16*)dz, lddz);
    else if (vec_ok<T>(C, {da, a, z, dz}, {ldda, lda, ldz, lddz}))
      launch_k(bn_act_bwd_apply_vec_kernel<T>, vec_blocks(rows, C, 16 / (int)sizeof(T)), 256, 0, as_stream(stream), 
          (const T*)da, ldda, (const T*)a, lda, (const T*)z, ldz, rows, C, mean, invstd, gamma, beta, drop, rows_per_n, act, slope,
          sums, 1.0f / (float)count, (T*)dz, lddz);
    else
      launch_k(bn_act_bwd_apply_kernel<T>, ew_blocks(rows * C, 4), 256, 0, as_stream(stream), 
          (const T*)da, ldda, (const T*)a, lda, (const T*)z, ldz, rows, C, mean, invstd, gamma, drop, rows_per_n, act, slope,
          sums, 1.0f / (float)count, (T*)dz, lddz);
  });
  return check_launch("bn_act_bwd_apply");
}

int dcv_act_bwd(int dtype, const void* da, int64_t ldda, const void* a, int64_t lda, int64_t rows, int C, int act,
                float slope, void* dz, int64_t lddz, void* stream) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, {
    if (vec_ok<T>(C, {da, a, dz}, {ldda, lda, lddz}))
      launch_k(act_bwd_vec_kernel<T>, vec_blocks(rows, C, 16 / (int)sizeof(T)), 256, 0, as_stream(stream), 
          (const T*)da, ldda, (const T*)a, lda, rows, C, act, slope, (T*)dz, lddz);
    else
      launch_k(act_bwd_kernel<T>, ew_blocks(rows * C, 4), 256, 0, as_stream(stream), (const T*)da, ldda, (const T*)a, lda, rows, C, act,
                                                                               slope, (T*)dz, lddz);
  });
  return check_launch("act_bwd");
}

int dcv_add_noise(int dtype, const void* x, int64_t ldx, const float* noise, float sigma, int64_t rows, int C,
                  void* out, int64_t ldo, void* stream) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, launch_k(add_noise_kernel<T>, ew_blocks(rows * C, 4), 256, 0, as_stream(stream), 
                        (const T*)x, ldx, noise, sigma, rows, C, (T*)out, ldo));
  return check_launch("add_noise");
}

int dcv_axpy(int dtype, const void* x, int64_t ldx, int64_t rows, int C, void* out, int64_t ldo, int accumulate,
             void* stream) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, {
    if (vec_ok<T>(C, {x, out}, {ldx, ldo}))
      launch_k(axpy_vec_kernel<T>, vec_blocks(rows, C, 16 / (int)sizeof(T)), 256, 0, as_stream(stream), (const T*)x, ldx, rows, C, (T*)out,
                                                                                                  ldo, accumulate);
    else
      launch_k(axpy_kernel<T>, ew_blocks(rows * C, 4), 256, 0, as_stream(stream), (const T*)x, ldx, rows, C, (T*)out, ldo, accumulate);
  });
  return check_launch("axpy");
}

int dcv_tdiff(int dtype, const void* x, int64_t ldx, int N, int T_, int64_t hw, int C, void* out, int64_t ldo,
              void* stream) {
  const int64_t total = (int64_t)N * (T_ - 1) * hw * C;
  if (total <= 0) return 0;
  DISPATCH_T(dtype, launch_k(tdiff_kernel<T>, ew_blocks(total, 4), 256, 0, as_stream(stream), (const T*)x, ldx, N, T_, hw, C,
                                                                                        (T*)out, ldo));
  return check_launch("tdiff");
}

int dcv_tdiff_bwd(int dtype, const void* dy, int64_t lddy, int N, int T_, int64_t hw, int C, void* dx, int64_t lddx,
                  int accumulate, void* stream) {
  const int64_t total = (int64_t)N * T_ * hw * C;
  if (total <= 0) return 0;
  DISPATCH_T(dtype, launch_k(tdiff_bwd_kernel<T>, ew_blocks(total, 4), 256, 0, as_stream(stream), 
                        (const T*)dy, lddy, N, T_, hw, C, (T*)dx, lddx, accumulate));
  return check_launch("tdiff_bwd");
}

int dcv_softmax(int dtype, const void* z, int64_t ldz, int64_t rows, int C, void* y, int64_t ldy, void* stream) {
  if (rows == 0) return 0;
  DISPATCH_T(dtype, launch_k(softmax_kernel<T>, ew_blocks(rows), 256, 0, as_stream(stream), (const T*)z, ldz, rows, C, (T*)y, ldy));
  return check_launch("softmax");
}

int dcv_softmax_bwd(int dtype, const void* dy, int64_t lddy, const void* y, int64_t ldy, int64_t rows, int C, void* dz,
                    int64_t lddz, void* stream) {
  if (rows == 0) return 0;
  DISPATCH_T(dtype, launch_k(softmax_bwd_kernel<T>, ew_blocks(rows), 256, 0, as_stream(stream), (const T*)dy, lddy, (const T*)y,
                                                                                          ldy, rows, C, (T*)dz, lddz));
  return check_launch("softmax_bwd");
}

int dcv_segm_remap(int dtype, const void* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, void* stream) {
  if (rows == 0) return 0;
  DISPATCH_T(dtype, launch_k(segm_remap_kernel<T>, ew_blocks(rows), 256, 0, as_stream(stream), (const T*)x, ldx, rows, C, (T*)y, ldy));
  return check_launch("segm_remap");
}

int dcv_to_channels_last(int dtype, const float* src, int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw, int N,
                         int C, int T_, int H, int W, void* dst, int64_t ld, void* stream) {
  const int64_t total = (int64_t)N * T_ * H * W * C;
  if (total == 0) return 0;
  if (dtype == DCV_BF16 && C <= 4) {
    launch_k(to_cl_px_bf16_kernel, ew_blocks(total / C, 2), 256, 0, as_stream(stream), src, sn, sc, st, sh, sw, N, C, T_, H, W,
             (__nv_bfloat16*)dst, ld);
    return check_launch("to_channels_last_px");
  }
  DISPATCH_T(dtype, launch_k(to_cl_kernel<T>, ew_blocks(total, 4), 256, 0, as_stream(stream), src, sn, sc, st, sh, sw, N, C, T_,
                                                                                        H, W, (T*)dst, ld));
  return check_launch("to_channels_last");
}

int dcv_from_channels_last(int dtype, const void* src, int64_t ld, int N, int C, int T_, int H, int W, float* dst,
                           int64_t sn, int64_t sc, int64_t st, int64_t sh, int64_t sw, int accumulate, void* stream) {
  const int64_t total = (int64_t)N * T_ * H * W * C;
  if (total == 0) return 0;
  DISPATCH_T(dtype, launch_k(from_cl_kernel<T>, ew_blocks(total, 4), 256, 0, as_stream(stream), 
                        (const T*)src, ld, N, C, T_, H, W, dst, sn, sc, st, sh, sw, accumulate));
  return check_launch("from_channels_last");
}

int dcv_frame_copy(int dtype, void* clips, int64_t ldc, int N, int T_, int64_t hw, int C, int t, const int* t_dev, void* frames,
                   int64_t ldf, int reverse, int accumulate, void* stream) {
  const int64_t total = (int64_t)N * hw * C;
  if (total == 0) return 0;
  DCV_REQUIRE(t_dev || (t >= 0 && t < T_), "frame_copy: frame %d out of range [0,%d)", t, T_);
  DISPATCH_T(dtype, launch_k(frame_copy_kernel<T>, ew_blocks(total, 4), 256, 0, as_stream(stream), 
                        (T*)clips, ldc, N, T_, hw, C, t, t_dev, (T*)frames, ldf, reverse, accumulate));
  return check_launch("frame_copy");
}

int dcv_col2im_act(int dtype, const void* P, int64_t ldp, int N, int Ih, int Iw, int Cout, int KH, int KW, int stride, int pad,
                   int act, float slope, void* y, int64_t ldy, int Oh, int Ow, void* stream) {
  DCV_REQUIRE(P && y, "col2im_act: null pointer");
  DCV_REQUIRE(Cout >= 1 && Cout <= 4 && KH >= 1 && KW >= 1 && stride >= 1 && pad >= 0, "col2im_act: bad shape (Cout %d k %dx%d s %d p %d)", Cout, KH, KW, stride, pad);
  DCV_REQUIRE(Oh == (Ih - 1) * stride - 2 * pad + KH && Ow == (Iw - 1) * stride - 2 * pad + KW, "col2im_act: output %dx%d does not match input %dx%d", Oh, Ow, Ih, Iw);
  DCV_REQUIRE(ldp >= (int64_t)Cout * KH * KW, "col2im_act: pitch %lld < %d columns", (long long)ldp, Cout * KH * KW);
  if ((int64_t)N * Oh * Ow == 0) return 0;
  if (dtype == DCV_BF16 && (stride == 1 || stride == 2) && ldp % 8 == 0 && (((uintptr_t)P) & 15) == 0 && N <= 65535) {
    const int rowv = (Cout * KH * KW + 7) / 8;
    // staged region: input rows [ih0, ih_last] for 16 output rows (ih_last = (oh0 + 15 + pad) / S)
    const int RH = (16 + KH - 2) / stride + 2;
    int TW = 64, RW = (TW + KW - 2) / stride + 2;
    if (Ow < 64 || (size_t)RH * RW * (rowv * 4 + 1) * 4 > 48 * 1024) { TW = 16; RW = (TW + KW - 2) / stride + 2; }
    const size_t smem = (size_t)RH * RW * (rowv * 4 + 1) * 4;
    if (smem <= 48 * 1024) {
      dim3 grid(ceil_div(Ow, TW), ceil_div(Oh, 16), N);
      if (stride == 1)
        launch_k(col2im_act_tiled_kernel<1>, grid, 256, smem, as_stream(stream), (const __nv_bfloat16*)P, ldp, Ih, Iw, Cout, KH, KW, pad, Oh, Ow,
                                                                           rowv, RH, RW, TW, act, slope, (__nv_bfloat16*)y, ldy);
      else
        launch_k(col2im_act_tiled_kernel<2>, grid, 256, smem, as_stream(stream), (const __nv_bfloat16*)P, ldp, Ih, Iw, Cout, KH, KW, pad, Oh, Ow,
                                                                           rowv, RH, RW, TW, act, slope, (__nv_bfloat16*)y, ldy);
      return check_launch("col2im_act_tiled");
    }
  }
  const int nb = ew_blocks((int64_t)N * Oh * Ow, 1);
  DISPATCH_T(dtype, {
    if (stride == 1)
      launch_k(col2im_act_kernel<T, 1>, nb, 256, 0, as_stream(stream), (const T*)P, ldp, N, Ih, Iw, Cout, KH, KW, stride, pad, Oh, Ow, act, slope, (T*)y, ldy);
    else if (stride == 2)
      launch_k(col2im_act_kernel<T, 2>, nb, 256, 0, as_stream(stream), (const T*)P, ldp, N, Ih, Iw, Cout, KH, KW, stride, pad, Oh, Ow, act, slope, (T*)y, ldy);
    else
      launch_k(col2im_act_kernel<T, 0>, nb, 256, 0, as_stream(stream), (const T*)P, ldp, N, Ih, Iw, Cout, KH, KW, stride, pad, Oh, Ow, act, slope, (T*)y, ldy);
  });
  return check_launch("col2im_act");
}

int dcv_fold_w(int dtype, const void* xg, int64_t ldg, int cg, const float* noise_g, const void* xc, int64_t ldc, int cc,
               const float* noise_c, float sigma, int64_t lines, int W, int kw, int sw, int pw, void* out, int64_t ldo,
               void* stream) {
  DCV_REQUIRE(xg && xc && out, "fold_w: null pointer");
  DCV_REQUIRE(cg >= 1 && cc >= 1 && cg + cc <= 8 && kw >= 1 && sw >= 1 && pw >= 0 && W + 2 * pw >= kw, "fold_w: bad shape (cg %d cc %d kw %d sw %d pw %d W %d)", cg, cc, kw, sw, pw, W);
  DCV_REQUIRE(ldo >= (int64_t)kw * (cg + cc), "fold_w: output pitch %lld < %d folded channels", (long long)ldo, kw * (cg + cc));
  const int Ow = (W + 2 * pw - kw) / sw + 1;
  if (lines * Ow == 0) return 0;
  DISPATCH_T(dtype, launch_k(fold_w_kernel<T>, ew_blocks(lines * Ow, 1), 256, 0, as_stream(stream), 
                        (const T*)xg, ldg, cg, noise_g, (const T*)xc, ldc, cc, noise_c, sigma, lines, W, Ow, kw, sw, pw, (T*)out, ldo));
  return check_launch("fold_w");
}

int dcv_unfold_w(int dtype, const void* d2, int64_t ld2, int64_t lines, int W, int kw, int sw, int pw, void* dxg, int64_t ldg,
                 int cg, void* dxc, int64_t ldc, int cc, void* stream) {
  DCV_REQUIRE(d2 && dxg && dxc, "unfold_w: null pointer");
  DCV_REQUIRE(cg >= 1 && cc >= 1 && cg + cc <= 8 && kw >= 1 && sw >= 1 && pw >= 0 && W + 2 * pw >= kw, "unfold_w: bad shape");
  const int Ow = (W + 2 * pw - kw) / sw + 1;
  if (lines * W == 0) return 0;
  DISPATCH_T(dtype, launch_k(unfold_w_kernel<T>, ew_blocks(lines * W, 1), 256, 0, as_stream(stream), 
                        (const T*)d2, ld2, lines, W, Ow, kw, sw, pw, (T*)dxg, ldg, cg, (T*)dxc, ldc, cc));
  return check_launch("unfold_w");
}

int dcv_copy_cl(int src_dtype, const void* src, int64_t lds, int dst_dtype, void* dst, int64_t ldd, int64_t rows, int C,
                void* stream) {
  const int64_t total = rows * C;
  if (total == 0) return 0;
  const int nb = ew_blocks(total, 4);
  cudaStream_t s = as_stream(stream);
  if (src_dtype == DCV_F32 && dst_dtype == DCV_F32)
    launch_k(copy_cl_kernel<float, float>, nb, 256, 0, s, (const float*)src, lds, (float*)dst, ldd, rows, C);
  else if (src_dtype == DCV_F32)
    launch_k(copy_cl_kernel<float, __nv_bfloat16>, nb, 256, 0, s, (const float*)src, lds, (__nv_bfloat16*)dst, ldd, rows, C);
  else if (dst_dtype == DCV_F32)
    launch_k(copy_cl_kernel<__nv_bfloat16, float>, nb, 256, 0, s, (const __nv_bfloat16*)src, lds, (float*)dst, ldd, rows, C);
  else if (C <= 8)
    launch_k(copy_cl_rows_kernel<__nv_bfloat16, __nv_bfloat16>, ew_blocks(rows, 2), 256, 0, s, (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst,
                                                                                          ldd, rows, C);
  else
    launch_k(copy_cl_kernel<__nv_bfloat16, __nv_bfloat16>, nb, 256, 0, s, (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, rows, C);
  return check_launch("copy_cl");
}

}  // extern "C"
