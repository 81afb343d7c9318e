// GRU latent trajectory, adversarial/hinge loss heads and fused multi-tensor Adam.
#include "common.cuh"

namespace dcv {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// ------------------------------------------------------------------------------------------
// GRU trajectory (generator.py:84-101): h_t = GRUCell(eps_t, h_{t-1}), t = 1..T, gate order r,z,n.
// One warp walks one batch row through all T steps; lanes own gate rows j = lane, lane+32, ...
// Shared memory per warp: h[D] e[D] gi[3D] gh[3D] (+ backward: dgi[3D] dgh[3D] dh[D] dW_ih[3D*D] dW_hh[3D*D] db[6D])
__global__ void gru_fwd_kernel(const float* __restrict__ h0, const float* __restrict__ eps, const float* __restrict__ w_ih,
                               const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                               const float* __restrict__ b_hh, int B, int T, int D, float* __restrict__ hs) {
  pdl_wait(); pdl_trigger();
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  float* h = sm + warp * (8 * D);
  float* e = h + D; float* gi = e + D; float* gh = gi + 3 * D;
  for (int b = blockIdx.x * nw + warp; b < B; b += gridDim.x * nw) {
    for (int d = lane; d < D; d += 32) h[d] = h0[b * D + d];
    __syncwarp();
    for (int t = 0; t < T; ++t) {
      for (int d = lane; d < D; d += 32) e[d] = eps[((int64_t)t * B + b) * D + d];
      __syncwarp();
      for (int j = lane; j < 3 * D; j += 32) {
        float a = b_ih[j], c = b_hh[j];
        for (int k = 0; k < D; ++k) { a = fmaf(w_ih[j * D + k], e[k], a); c = fmaf(w_hh[j * D + k], h[k], c); }
        gi[j] = a; gh[j] = c;
      }
      __syncwarp();
      for (int d = lane; d < D; d += 32) {
        const float r = sigmoidf_(gi[d] + gh[d]);
        const float z = sigmoidf_(gi[D + d] + gh[D + d]);
        const float n = tanhf(gi[2 * D + d] + r * gh[2 * D + d]);
        const float hn = (1.f - z) * n + z * h[d];
        hs[((int64_t)b * T + t) * D + d] = hn;
        gi[d] = hn;  // stash; h[] is still read by other lanes' nothing at this point, but keep it simple
      }
      __syncwarp();
      for (int d = lane; d < D; d += 32) h[d] = gi[d];
      __syncwarp();
    }
  }
}

// Backward through time.  A single block (nw warps); warp w handles rows w, w+nw, ... and keeps its
// weight-gradient accumulators in shared memory; the block then sums the warps in a fixed order, so
// the result is deterministic.
__global__ void gru_bwd_kernel(const float* __restrict__ h0, const float* __restrict__ eps, const float* __restrict__ hs,
                               const float* __restrict__ dhs, const float* __restrict__ w_ih,
                               const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                               const float* __restrict__ b_hh, int B, int T, int D, float* __restrict__ dw_ih,
                               float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh,
                               int accumulate) {
  pdl_wait(); pdl_trigger();
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  const int G = 3 * D;
  const int per_warp = 3 * D /*h,e,dh*/ + 4 * G /*gi,gh,dgi,dgh*/ + 2 * G * D + 2 * G;
  float* base = sm + warp * per_warp;
  float* h = base; float* e = h + D; float* dh = e + D;
  float* gi = dh + D; float* gh = gi + G; float* dgi = gh + G; float* dgh = dgi + G;
  float* aWi = dgh + G; float* aWh = aWi + G * D; float* abi = aWh + G * D; float* abh = abi + G;
  for (int i = lane; i < 2 * G * D + 2 * G; i += 32) aWi[i] = 0.f;
  __syncwarp();
  for (int b = warp; b < B; b += nw) {
    for (int d = lane; d < D; d += 32) dh[d] = 0.f;
    __syncwarp();
    for (int t = T - 1; t >= 0; --t) {
      for (int d = lane; d < D; d += 32) {
        h[d] = t == 0 ? h0[b * D + d] : hs[((int64_t)b * T + (t - 1)) * D + d];
        e[d] = eps[((int64_t)t * B + b) * D + d];
        dh[d] += dhs[((int64_t)b * T + t) * D + d];
      }
      __syncwarp();
      for (int j = lane; j < G; j += 32) {
        float a = b_ih[j], c = b_hh[j];
        for (int k = 0; k < D; ++k) { a = fmaf(w_ih[j * D + k], e[k], a); c = fmaf(w_hh[j * D + k], h[k], c); }
        gi[j] = a; gh[j] = c;
      }
      __syncwarp();
      for (int d = lane; d < D; d += 32) {
        const float r = sigmoidf_(gi[d] + gh[d]);
        const float z = sigmoidf_(gi[D + d] + gh[D + d]);
        const float n = tanhf(gi[2 * D + d] + r * gh[2 * D + d]);
        const float g = dh[d];
        const float dn_pre = g * (1.f - z) * (1.f - n * n);
        const float dz_pre = g * (h[d] - n) * z * (1.f - z);
        const float dr_pre = dn_pre * gh[2 * D + d] * r * (1.f - r);
        dgi[d] = dr_pre; dgi[D + d] = dz_pre; dgi[2 * D + d] = dn_pre;
        dgh[d] = dr_pre; dgh[D + d] = dz_pre; dgh[2 * D + d] = dn_pre * r;
        dh[d] = g * z;  // direct path to h_{t-1}
      }
      __syncwarp();
      for (int j = lane; j < G; j += 32) {
        const float a = dgi[j], c = dgh[j];
        for (int k = 0; k < D; ++k) { aWi[j * D + k] = fmaf(a, e[k], aWi[j * D + k]); aWh[j * D + k] = fmaf(c, h[k], aWh[j * D + k]); }
        abi[j] += a; abh[j] += c;
      }
      for (int k = lane; k < D; k += 32) {
        float s = 0.f;
        for (int j = 0; j < G; ++j) s = fmaf(dgh[j], w_hh[j * D + k], s);
        dh[k] += s;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  const int nacc = 2 * G * D + 2 * G;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sm[w * per_warp + 3 * D + 4 * G + i];
    float* dst;
    if (i < G * D) dst = dw_ih + i;
    else if (i < 2 * G * D) dst = dw_hh + (i - G * D);
    else if (i < 2 * G * D + G) dst = db_ih + (i - 2 * G * D);
    else dst = db_hh + (i - 2 * G * D - G);
    *dst = accumulate ? *dst + s : s;
  }
}

// ---- dim_z_motion <= 10 (every shipped configuration): 3 D <= 32 gate rows, ONE per lane.  The generic kernels above re-read the
// weights from global memory inside every time step, keep the weight-gradient accumulators in shared memory (read-modify-write
// per step) and fetch the step's inputs with dependent global loads: 79 us for the 16-step BPTT of a batch of 32, all of it
// latency.  Here a lane keeps its gate row of both weight matrices (and, backward, its accumulators and its column of W_hh) in
// registers and the whole trajectory of the row (eps, h, dh) is staged in shared memory up front.  Same operations in the same
// order as the generic kernels: bit-identical results.
template <int D>
__global__ void gru_fwd_small_kernel(const float* __restrict__ h0, const float* __restrict__ eps, const float* __restrict__ w_ih,
                                     const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                     const float* __restrict__ b_hh, int B, int T, float* __restrict__ hs) {
  pdl_wait(); pdl_trigger();
  constexpr int G = 3 * D;
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  float* es = sm + warp * (T * D + D + 2 * G);          // eps of every step
  float* h = es + T * D; float* gi = h + D; float* gh = gi + G;
  float wi[D], wh[D], bi = 0.f, bh = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) { wi[k] = lane < G ? w_ih[lane * D + k] : 0.f; wh[k] = lane < G ? w_hh[lane * D + k] : 0.f; }
  if (lane < G) { bi = b_ih[lane]; bh = b_hh[lane]; }
  for (int b = blockIdx.x * nw + warp; b < B; b += gridDim.x * nw) {
    __syncwarp();
    for (int i = lane; i < T * D; i += 32) es[i] = eps[((int64_t)(i / D) * B + b) * D + i % D];
    if (lane < D) h[lane] = h0[b * D + lane];
    __syncwarp();
    for (int t = 0; t < T; ++t) {
      if (lane < G) {
        float a = bi, c = bh;
#pragma unroll
        for (int k = 0; k < D; ++k) { a = fmaf(wi[k], es[t * D + k], a); c = fmaf(wh[k], h[k], c); }
        gi[lane] = a; gh[lane] = c;
      }
      __syncwarp();
      float hn = 0.f;
      if (lane < D) {
        const float r = sigmoidf_(gi[lane] + gh[lane]);
        const float z = sigmoidf_(gi[D + lane] + gh[D + lane]);
        const float n = tanhf(gi[2 * D + lane] + r * gh[2 * D + lane]);
        hn = (1.f - z) * n + z * h[lane];
        hs[((int64_t)b * T + t) * D + lane] = hn;
      }
      __syncwarp();
      if (lane < D) h[lane] = hn;
      __syncwarp();
    }
  }
}

template <int D>
__global__ void __launch_bounds__(1024, 1) gru_bwd_small_kernel(const float* __restrict__ h0, const float* __restrict__ eps, const float* __restrict__ hs,
                                     const float* __restrict__ dhs, const float* __restrict__ w_ih,
                                     const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                     const float* __restrict__ b_hh, int B, int T, float* __restrict__ dw_ih,
                                     float* __restrict__ dw_hh, float* __restrict__ db_ih, float* __restrict__ db_hh,
                                     int accumulate) {
  pdl_wait(); pdl_trigger();
  constexpr int G = 3 * D, NACC = 2 * G * D + 2 * G;
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  const int per_warp = 3 * T * D + 2 * D + 4 * G + NACC;
  float* whs = sm + nw * per_warp;                                                // W_hh, shared by the warps (read by column)
  for (int i = threadIdx.x; i < G * D; i += blockDim.x) whs[i] = w_hh[i];
  __syncthreads();
  float* base = sm + warp * per_warp;
  float* es = base; float* hp = es + T * D; float* dhs_s = hp + T * D;          // eps_t, h_{t-1}, dL/dh_t for every step
  float* dh = dhs_s + T * D; float* pad = dh + D;
  float* gi = pad + D; float* gh = gi + G; float* dgi = gh + G; float* dgh = dgi + G;
  float* acc = dgh + G;                                                           // this warp's sums, dumped at the end
  float wi[D], wh[D], aWi[D], aWh[D], bi = 0.f, bh = 0.f, abi = 0.f, abh = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    wi[k] = lane < G ? w_ih[lane * D + k] : 0.f; wh[k] = lane < G ? w_hh[lane * D + k] : 0.f;
    aWi[k] = 0.f; aWh[k] = 0.f;
  }
  if (lane < G) { bi = b_ih[lane]; bh = b_hh[lane]; }
  for (int b = warp; b < B; b += nw) {
    __syncwarp();
    for (int i = lane; i < T * D; i += 32) {
      const int t = i / D, d = i % D;
      es[i] = eps[((int64_t)t * B + b) * D + d];
      hp[i] = t == 0 ? h0[b * D + d] : hs[((int64_t)b * T + (t - 1)) * D + d];
      dhs_s[i] = dhs[((int64_t)b * T + t) * D + d];
    }
    if (lane < D) dh[lane] = 0.f;
    __syncwarp();
    for (int t = T - 1; t >= 0; --t) {
      const float* e = es + t * D; const float* h = hp + t * D;
      if (lane < D) dh[lane] += dhs_s[t * D + lane];
      if (lane < G) {
        float a = bi, c = bh;
#pragma unroll
        for (int k = 0; k < D; ++k) { a = fmaf(wi[k], e[k], a); c = fmaf(wh[k], h[k], c); }
        gi[lane] = a; gh[lane] = c;
      }
      __syncwarp();
      if (lane < D) {
        const int d = lane;
        const float r = sigmoidf_(gi[d] + gh[d]);
        const float z = sigmoidf_(gi[D + d] + gh[D + d]);
        const float n = tanhf(gi[2 * D + d] + r * gh[2 * D + d]);
        const float g = dh[d];
        const float dn_pre = g * (1.f - z) * (1.f - n * n);
        const float dz_pre = g * (h[d] - n) * z * (1.f - z);
        const float dr_pre = dn_pre * gh[2 * D + d] * r * (1.f - r);
        dgi[d] = dr_pre; dgi[D + d] = dz_pre; dgi[2 * D + d] = dn_pre;
        dgh[d] = dr_pre; dgh[D + d] = dz_pre; dgh[2 * D + d] = dn_pre * r;
        dh[d] = g * z;  // direct path to h_{t-1}
      }
      __syncwarp();
      if (lane < G) {
        const float a = dgi[lane], c = dgh[lane];
#pragma unroll
        for (int k = 0; k < D; ++k) { aWi[k] = fmaf(a, e[k], aWi[k]); aWh[k] = fmaf(c, h[k], aWh[k]); }
        abi += a; abh += c;
      }
      if (lane < D) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < G; ++j) s = fmaf(dgh[j], whs[j * D + lane], s);
        dh[lane] += s;
      }
      __syncwarp();
    }
  }
  if (lane < G) {
#pragma unroll
    for (int k = 0; k < D; ++k) { acc[lane * D + k] = aWi[k]; acc[G * D + lane * D + k] = aWh[k]; }
    acc[2 * G * D + lane] = abi; acc[2 * G * D + G + lane] = abh;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sm[w * per_warp + (per_warp - NACC) + i];
    float* dst;
    if (i < G * D) dst = dw_ih + i;
    else if (i < 2 * G * D) dst = dw_hh + (i - G * D);
    else if (i < 2 * G * D + G) dst = db_ih + (i - 2 * G * D);
    else dst = db_hh + (i - 2 * G * D - G);
    *dst = accumulate ? *dst + s : s;
  }
}

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
loss_kernel(const T* __restrict__ y, int64_t ldy, int64_t n, int kind, float* __restrict__ loss_out, int accumulate,
            T* __restrict__ dy, int64_t lddy, float grad_scale) {
  pdl_wait(); pdl_trigger();
  __shared__ float red[256];
  float acc = 0.f;
  const float inv_n = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const float v = ldf(y + i * ldy);
    float l, g;
    switch (kind) {
      case DCV_LOSS_BCE_ONES:
      case DCV_LOSS_SOFTPLUS_NEG: l = softplusf_(-v); g = sigmoidf_(v) - 1.f; break;
      case DCV_LOSS_BCE_ZEROS: l = softplusf_(v); g = sigmoidf_(v); break;
      case DCV_LOSS_HINGE_REAL: l = fmaxf(1.f - v, 0.f); g = (1.f - v > 0.f) ? -1.f : 0.f; break;
      default: l = fmaxf(1.f + v, 0.f); g = (1.f + v > 0.f) ? 1.f : 0.f; break;
    }
    acc += l;
    if (dy) stf(dy + i * lddy, g * inv_n * grad_scale);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = red[0] * inv_n;
    loss_out[0] = accumulate ? loss_out[0] + l : l;
  }
}

// ------------------------------------------------------------------------------------------
// Adam, one launch for many tensors.  Arithmetic follows torch.optim.Adam (single-tensor path):
//   g += wd*p ; m = m + (g-m)(1-b1) ; v = b2*v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
constexpr int ADAM_MAXT = 40;
constexpr int ADAM_CHUNK = 4096;  // elements per block
struct AdamArgs {
  float* p[ADAM_MAXT];
  const float* g[ADAM_MAXT];
  float* m[ADAM_MAXT];
  float* v[ADAM_MAXT];
  int64_t numel[ADAM_MAXT];
  int block_start[ADAM_MAXT + 1];
  int ntensors;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float b1, float b2, float eps,
                                            float wd, float step_size, float inv_bc2_sqrt, float gscale) {
  g = g * gscale + wd * p;
  m = m + (g - m) * (1.f - b1);
  v = v * b2 + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

// Device-side step bookkeeping (CUDA-graph friendly): state = {int64 step; float step_size; float inv_bc2_sqrt}.
__global__ void adam_advance_kernel(long long* state, float lr, float b1, float b2) {
  pdl_wait(); pdl_trigger();
  const long long step = ++state[0];
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  float* f = reinterpret_cast<float*>(state + 1);
  f[0] = (float)((double)lr / bc1);
  f[1] = (float)(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(AdamArgs a, float b1, float b2, float eps, float wd, float step_size, float inv_bc2_sqrt, float gscale,
                  const float* __restrict__ dev_scalars) {
  pdl_wait(); pdl_trigger();
  if (dev_scalars) { step_size = dev_scalars[0]; inv_bc2_sqrt = dev_scalars[1]; }
  int t = 0;
  while (t + 1 < a.ntensors && (int)blockIdx.x >= a.block_start[t + 1]) ++t;
  const int64_t off = (int64_t)(blockIdx.x - a.block_start[t]) * ADAM_CHUNK;
  const int64_t n = a.numel[t];
  float* p = a.p[t] + off; const float* g = a.g[t] + off; float* m = a.m[t] + off; float* v = a.v[t] + off;
  const int64_t len = (n - off) < ADAM_CHUNK ? (n - off) : ADAM_CHUNK;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
  if (aligned && len == ADAM_CHUNK) {
#pragma unroll
    for (int it = 0; it < ADAM_CHUNK / (256 * 4); ++it) {
      const int i = (it * 256 + threadIdx.x) * 4;
      float4 pp = *reinterpret_cast<float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      adam_update(pp.x, gg.x, mm.x, vv.x, b1, b2, eps, wd, step_size, inv_bc2_sqrt, gscale);
      adam_update(pp.y, gg.y, mm.y, vv.y, b1, b2, eps, wd, step_size, inv_bc2_sqrt, gscale);
      adam_update(pp.z, gg.z, mm.z, vv.z, b1, b2, eps, wd, step_size, inv_bc2_sqrt, gscale);
      adam_update(pp.w, gg.w, mm.w, vv.w, b1, b2, eps, wd, step_size, inv_bc2_sqrt, gscale);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    }
  } else {
    for (int64_t i = threadIdx.x; i < len; i += 256) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(pp, g[i], mm, vv, b1, b2, eps, wd, step_size, inv_bc2_sqrt, gscale);
      p[i] = pp; m[i] = mm; v[i] = vv;
    }
  }
}

}  // namespace dcv

using namespace dcv;

// ------------------------------------------------------------------------------------------ dataset-side conversions
// The reference's dataset delivers float32 (C,T,H,W) clips made on the host from uint8 frames (dataset.py:128-131,
// 157-167: `astype(float32) / 127.5 - 1.0` after a transpose to channels-first) or from class indices (:177-181:
// `np.eye(25)[segm]`), and its sampler converts generated videos back on the host (util.py:74-79).  The frames on disk
// are channels-LAST uint8 - exactly this library's activation layout - so the conversions run on the device instead:
// a quarter of the PCIe bytes in, a quarter out, and no host-side transpose.
template <typename T>
__global__ void __launch_bounds__(256)
ingest_u8_kernel(const uint8_t* __restrict__ src, int64_t rows, int C, T* __restrict__ dst, int64_t ld) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    stf(dst + row * ld + c, (float)src[i] / 127.5f - 1.0f);
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
ingest_onehot_kernel(const I* __restrict__ idx, int64_t rows, int C, T* __restrict__ dst, int64_t ld) {
  pdl_wait(); pdl_trigger();
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    stf(dst + row * ld + c, (int)idx[row] == c ? 1.0f : 0.0f);
  }
}

// channels-last (N, T, HW, C) activations in [-1, 1] -> planar uint8 (N, C, T, HW), numpy's clip / (v + 1) / 2 * 255 / astype
template <typename T>
__global__ void __launch_bounds__(256)
export_u8_kernel(const T* __restrict__ src, int64_t ld, int N, int C, int T_, int64_t hw, uint8_t* __restrict__ dst) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)N * T_ * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % hw; const int64_t nt = i / hw; const int t = (int)(nt % T_); const int64_t n = nt / T_;
    for (int c = 0; c < C; ++c) {
      float v = ldf(src + i * ld + c);
      v = fminf(fmaxf(v, -1.0f), 1.0f);
      v = (v + 1.0f) / 2.0f * 255.0f;
      dst[((n * C + c) * T_ + t) * hw + p] = (uint8_t)v;
    }
  }
}

extern "C" {

int dcv_gru_traj_fwd(const float* h0, const float* eps, const float* w_ih, const float* w_hh, const float* b_ih,
                     const float* b_hh, int B, int T, int D, float* hs, void* stream) {
  DCV_REQUIRE(D >= 1 && D <= 64, "gru: dim_z_motion %d out of range [1,64]", D);
  if (B == 0) return 0;
  const int nw = 4;
  const int blocks = ceil_div(B, nw);
  if (D == 10 && (size_t)nw * (T * D + D + 6 * D) * sizeof(float) <= 48 * 1024) {
    launch_k(gru_fwd_small_kernel<10>, blocks, nw * 32, nw * (T * D + D + 6 * D) * sizeof(float), as_stream(stream), h0, eps, w_ih, w_hh,
             b_ih, b_hh, B, T, hs);
    return check_launch("gru_fwd_small");
  }
  launch_k(gru_fwd_kernel, blocks, nw * 32, nw * 8 * D * sizeof(float), as_stream(stream), h0, eps, w_ih, w_hh, b_ih, b_hh, B, T, D, hs);
  return check_launch("gru_fwd");
}

int dcv_gru_traj_bwd(const float* h0, const float* eps, const float* hs, const float* dhs, const float* w_ih,
                     const float* w_hh, const float* b_ih, const float* b_hh, int B, int T, int D, float* dw_ih,
                     float* dw_hh, float* db_ih, float* db_hh, int accumulate, void* stream) {
  DCV_REQUIRE(D >= 1 && D <= 64, "gru: dim_z_motion %d out of range [1,64]", D);
  const int G = 3 * D;
  if (D == 10) {
    const int pw = 3 * T * D + 2 * D + 4 * G + 2 * G * D + 2 * G;
    int nws = 32;
    while (nws > 1 && (nws > B || ((size_t)nws * pw + G * D) * sizeof(float) > 200 * 1024)) nws /= 2;
    const size_t smem_s = ((size_t)nws * pw + G * D) * sizeof(float);
    if (smem_s <= 200 * 1024) {
      static size_t smem_s_set = 48 * 1024;
      if (smem_s > smem_s_set) {
        DCV_CUDA(cudaFuncSetAttribute(gru_bwd_small_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        smem_s_set = smem_s;
      }
      launch_k(gru_bwd_small_kernel<10>, 1, nws * 32, smem_s, as_stream(stream), h0, eps, hs, dhs, w_ih, w_hh, b_ih, b_hh, B, T, dw_ih,
               dw_hh, db_ih, db_hh, accumulate);
      return check_launch("gru_bwd_small");
    }
  }
  const int per_warp = 3 * D + 4 * G + 2 * G * D + 2 * G;
  // one warp per batch row when the per-warp accumulators fit (the time steps are serial and latency-bound: 8 warps took
  // 170 us for B = 32, T = 16); the warps are summed in a fixed order, so the result stays deterministic
  int nw = 32;
  while (nw > 1 && (nw > B || (size_t)nw * per_warp * sizeof(float) > 200 * 1024)) nw /= 2;
  DCV_REQUIRE((size_t)nw * per_warp * sizeof(float) <= 200 * 1024, "gru: dim_z_motion %d too large for BPTT kernel", D);
  const size_t gru_smem = (size_t)nw * per_warp * sizeof(float);
  static size_t gru_smem_set = 48 * 1024;
  if (gru_smem > gru_smem_set) {
    DCV_CUDA(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_smem));
    gru_smem_set = gru_smem;
  }
  launch_k(gru_bwd_kernel, 1, nw * 32, gru_smem, as_stream(stream), 
      h0, eps, hs, dhs, w_ih, w_hh, b_ih, b_hh, B, T, D, dw_ih, dw_hh, db_ih, db_hh, accumulate);
  return check_launch("gru_bwd");
}

int dcv_loss_fwd_bwd(int dtype, const void* y, int64_t ldy, int64_t n, int kind, float* loss_out, int accumulate, void* dy,
                     int64_t lddy, float grad_scale, void* stream) {
  DCV_REQUIRE(n > 0, "loss: empty logits");
  DCV_REQUIRE(kind >= 0 && kind <= 4, "loss: unknown kind %d", kind);
  if (dtype == DCV_F32)
    launch_k(loss_kernel<float>, 1, 256, 0, as_stream(stream), (const float*)y, ldy, n, kind, loss_out, accumulate, (float*)dy, lddy, grad_scale);
  else
    launch_k(loss_kernel<__nv_bfloat16>, 1, 256, 0, as_stream(stream), (const __nv_bfloat16*)y, ldy, n, kind, loss_out, accumulate,
                                                                 (__nv_bfloat16*)dy, lddy, grad_scale);
  return check_launch("loss");
}

int dcv_adam_multi(int ntensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                   const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                   float grad_scale, void* stream) {
  DCV_REQUIRE(step >= 1, "adam: step must be >= 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int i = 0;
  while (i < ntensors) {
    AdamArgs a;
    int k = 0, blocks = 0;
    while (i < ntensors && k < ADAM_MAXT) {
      if (numel[i] > 0) {
        a.p[k] = p[i]; a.g[k] = g[i]; a.m[k] = m[i]; a.v[k] = v[i]; a.numel[k] = numel[i];
        a.block_start[k] = blocks;
        blocks += (int)((numel[i] + ADAM_CHUNK - 1) / ADAM_CHUNK);
        ++k;
      }
      ++i;
    }
    if (k == 0) break;
    a.block_start[k] = blocks;
    a.ntensors = k;
    launch_k(adam_multi_kernel, blocks, 256, 0, as_stream(stream), a, beta1, beta2, eps, weight_decay, step_size, inv_bc2_sqrt, grad_scale, nullptr);
    int rc = check_launch("adam_multi");
    if (rc) return rc;
  }
  return 0;
}

int dcv_adam_flat_dev(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, void* step_state, float grad_scale, void* stream) {
  DCV_REQUIRE(step_state && (((uintptr_t)step_state) & 7) == 0, "adam: step_state must be an 8-byte aligned 16-byte device buffer");
  if (n <= 0) return 0;
  launch_k(adam_advance_kernel, 1, 1, 0, as_stream(stream), (long long*)step_state, lr, beta1, beta2);
  int rc = check_launch("adam_advance");
  if (rc) return rc;
  AdamArgs a;
  a.p[0] = p; a.g[0] = g; a.m[0] = m; a.v[0] = v; a.numel[0] = n;
  a.block_start[0] = 0;
  const int blocks = (int)((n + ADAM_CHUNK - 1) / ADAM_CHUNK);
  a.block_start[1] = blocks;
  a.ntensors = 1;
  launch_k(adam_multi_kernel, blocks, 256, 0, as_stream(stream), a, beta1, beta2, eps, weight_decay, 0.f, 0.f, grad_scale,
                                                           reinterpret_cast<const float*>((long long*)step_state + 1));
  return check_launch("adam_flat_dev");
}

int dcv_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int64_t step, float grad_scale, void* stream) {
  float* pp[1] = {p}; const float* gg[1] = {g}; float* mm[1] = {m}; float* vv[1] = {v}; int64_t nn[1] = {n};
  return dcv_adam_multi(1, pp, gg, mm, vv, nn, lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream);
}

static inline int io_blocks(int64_t total) { int64_t b = (total + 1023) / 1024; if (b > 148 * 16) b = 148 * 16; if (b < 1) b = 1; return (int)b; }

int dcv_ingest_u8(int dtype, const void* src, int64_t rows, int C, void* dst, int64_t ld, void* stream) {
  DCV_REQUIRE(src && dst && C >= 1 && ld >= C, "ingest_u8: bad arguments");
  if (rows * C == 0) return 0;
  if (dtype == DCV_F32) launch_k(ingest_u8_kernel<float>, io_blocks(rows * C), 256, 0, as_stream(stream), (const uint8_t*)src, rows, C, (float*)dst, ld);
  else launch_k(ingest_u8_kernel<__nv_bfloat16>, io_blocks(rows * C), 256, 0, as_stream(stream), (const uint8_t*)src, rows, C, (__nv_bfloat16*)dst, ld);
  return check_launch("ingest_u8");
}

int dcv_ingest_onehot(int dtype, const void* idx, int idx_bytes, int64_t rows, int C, void* dst, int64_t ld, void* stream) {
  DCV_REQUIRE(idx && dst && C >= 1 && ld >= C, "ingest_onehot: bad arguments");
  DCV_REQUIRE(idx_bytes == 1 || idx_bytes == 8, "ingest_onehot: class indices must be uint8 or int64 (element size %d)", idx_bytes);
  if (rows * C == 0) return 0;
  const int nb = io_blocks(rows * C);
  cudaStream_t st = as_stream(stream);
  if (dtype == DCV_F32) {
    if (idx_bytes == 1) launch_k(ingest_onehot_kernel<float, uint8_t>, nb, 256, 0, st, (const uint8_t*)idx, rows, C, (float*)dst, ld);
    else launch_k(ingest_onehot_kernel<float, long long>, nb, 256, 0, st, (const long long*)idx, rows, C, (float*)dst, ld);
  } else {
    if (idx_bytes == 1) launch_k(ingest_onehot_kernel<__nv_bfloat16, uint8_t>, nb, 256, 0, st, (const uint8_t*)idx, rows, C, (__nv_bfloat16*)dst, ld);
    else launch_k(ingest_onehot_kernel<__nv_bfloat16, long long>, nb, 256, 0, st, (const long long*)idx, rows, C, (__nv_bfloat16*)dst, ld);
  }
  return check_launch("ingest_onehot");
}

int dcv_export_u8(int dtype, const void* src, int64_t ld, int N, int C, int T_, int64_t hw, void* dst, void* stream) {
  DCV_REQUIRE(src && dst && C >= 1 && ld >= C, "export_u8: bad arguments");
  const int64_t total = (int64_t)N * T_ * hw;
  if (total == 0) return 0;
  if (dtype == DCV_F32) launch_k(export_u8_kernel<float>, io_blocks(total * 2), 256, 0, as_stream(stream), (const float*)src, ld, N, C, T_, hw, (uint8_t*)dst);
  else launch_k(export_u8_kernel<__nv_bfloat16>, io_blocks(total * 2), 256, 0, as_stream(stream), (const __nv_bfloat16*)src, ld, N, C, T_, hw, (uint8_t*)dst);
  return check_launch("export_u8");
}

}  // extern "C"
