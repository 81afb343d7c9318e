// CUDA-core (fp32 accumulate) implicit-GEMM correlation kernels.
//
// These cover every convolution on the DCVGAN path in exact (fp32) mode and the layers whose
// channel counts are too small for tensor cores in fast (bf16) mode:
//   nn.Conv2d / nn.Conv3d forward            -> DCV_DIR_GATHER   (generator.py:174,204; discriminator.py:81-101,182-206,288-305)
//   nn.ConvTranspose2d forward               -> DCV_DIR_SCATTER  (generator.py:61-73,240,274)
//   their data gradients                     -> the opposite direction
//   their weight gradients                   -> wgrad_simt
// Scatter (transposed) correlations are decomposed into stride^d sub-pixel phases so that no
// multiply ever touches a structural zero.
#include "common.cuh"
#include "conv_geom.cuh"

namespace dcv {

constexpr int BK = 16;

template <typename T, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvP p, const T* __restrict__ x, int64_t ldx, const float* __restrict__ wp,
                 T* __restrict__ y, int64_t ldy, int act, float slope) {
  pdl_wait(); pdl_trigger();
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  constexpr int EA = BM * BK / 256;          // A elements per thread
  constexpr int TPR = BK / EA;               // threads per A row
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const PhaseInfo f = make_phase(p, blockIdx.z);
  const int64_t M = (int64_t)p.N * f.Qt * f.Qh * f.Qw;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  if (m0 >= M) return;
  const int nblk = blockIdx.y * BN;
  const int tid = threadIdx.x;

  // the row this thread gathers for A
  const int arow = tid / TPR;
  const int ak0 = (tid % TPR) * EA;
  int64_t am = m0 + arow;
  const bool arow_valid = am < M;
  int an = 0, aot = 0, aoh = 0, aow = 0;
  if (arow_valid) {
    aow = (int)(am % f.Qw); am /= f.Qw;
    aoh = (int)(am % f.Qh); am /= f.Qh;
    aot = (int)(am % f.Qt); an = (int)(am / f.Qt);
  }

  const int tm = tid / (BN / TN);
  const int tn = tid % (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int jt = 0; jt < f.nt; ++jt) {
    const int it = aot * f.mult + f.offt + f.sgn * jt;
    for (int jh = 0; jh < f.nh; ++jh) {
      const int ih = aoh * f.mulh + f.offh + f.sgn * jh;
      for (int jw = 0; jw < f.nw; ++jw) {
        const int iw = aow * f.mulw + f.offw + f.sgn * jw;
        const bool inb = arow_valid && it >= 0 && it < p.It && ih >= 0 && ih < p.Ih && iw >= 0 && iw < p.Iw;
        const T* xrow = x;
        if (inb) xrow = x + ((((int64_t)an * p.It + it) * p.Ih + ih) * p.Iw + iw) * ldx;
        const int tap = ((f.a0t + f.ast * jt) * p.kh + (f.a0h + f.ash * jh)) * p.kw + (f.a0w + f.asw * jw);
        const float* wt = wp + (int64_t)tap * p.Kc * p.Nc;
        for (int c0 = 0; c0 < p.Kc; c0 += BK) {
#pragma unroll
          for (int i = 0; i < EA; ++i) {
            const int c = c0 + ak0 + i;
            float v = 0.f;
            if (inb && c < p.Kc) v = ldf(xrow + c);
            As[ak0 + i][arow] = v;
          }
          if (BN * BK >= 256) {
            constexpr int EB = (BN * BK >= 256) ? BN * BK / 256 : 1;
            constexpr int TPK = BN / EB;
            const int kb = tid / TPK;
            const int nb = (tid % TPK) * EB;
            const int c = c0 + kb;
#pragma unroll
            for (int j = 0; j < EB; ++j) {
              const int n = nblk + nb + j;
              float v = 0.f;
              if (c < p.Kc && n < p.Nc) v = wt[(int64_t)c * p.Nc + n];
              Bs[kb][nb + j] = v;
            }
          } else {
            if (tid < BN * BK) {
              const int kb = tid / BN, nb = tid % BN;
              const int c = c0 + kb, n = nblk + nb;
              float v = 0.f;
              if (c < p.Kc && n < p.Nc) v = wt[(int64_t)c * p.Nc + n];
              Bs[kb][nb] = v;
            }
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][tm * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tn * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
              for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          }
          __syncthreads();
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + tm * TM + i;
    if (m >= M) continue;
    const int ow = (int)(m % f.Qw); m /= f.Qw;
    const int oh = (int)(m % f.Qh); m /= f.Qh;
    const int ot = (int)(m % f.Qt); const int n = (int)(m / f.Qt);
    const int64_t pos = (((int64_t)n * p.Ot + (ot * f.ost + f.rt)) * p.Oh + (oh * f.osh + f.rh)) * p.Ow + (ow * f.osw + f.rw);
    T* yrow = y + pos * ldy;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = nblk + tn * TN + j;
      if (c < p.Nc) stf(yrow + c, apply_act(acc[i][j], act, slope));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: partial[z][tap][cl][cs] = sum_{m in split z} L[gather(m,tap)][cl] * S[m][cs]
struct WgradP {
  dcv_geom g;
  int64_t M;          // N*Ts*Hs*Ws
  int64_t m_per_split;
  int tilesB;
};

template <typename T, int BA, int BB, int TA, int TB>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(WgradP p, const T* __restrict__ xl, int64_t ldl, const T* __restrict__ xs, int64_t lds,
                  float* __restrict__ partial) {
  pdl_wait(); pdl_trigger();
  static_assert((BA / TA) * (BB / TB) == 256, "256 threads");
  __shared__ float As[BK][BA + 4];
  __shared__ float Bs[BK][BB + 4];
  __shared__ int64_t offL[BK];
  __shared__ int64_t offS[BK];
  const dcv_geom& g = p.g;
  const int tileA = blockIdx.x / p.tilesB, tileB = blockIdx.x % p.tilesB;
  const int tap = blockIdx.y;
  const int tc = tap % g.kw, tb = (tap / g.kw) % g.kh, ta_ = tap / (g.kw * g.kh);
  const int64_t mb = (int64_t)blockIdx.z * p.m_per_split;
  int64_t me = mb + p.m_per_split; if (me > p.M) me = p.M;
  const int tid = threadIdx.x;
  const int ta = tid / (BB / TB), tbb = tid % (BB / TB);
  float acc[TA][TB];
#pragma unroll
  for (int i = 0; i < TA; ++i)
#pragma unroll
    for (int j = 0; j < TB; ++j) acc[i][j] = 0.f;

  for (int64_t mc = mb; mc < me; mc += BK) {
    if (tid < BK) {
      int64_t m = mc + tid;
      int64_t ol = -1, os = -1;
      if (m < me) {
        os = m * lds;
        const int ws = (int)(m % g.Ws); m /= g.Ws;
        const int hs = (int)(m % g.Hs); m /= g.Hs;
        const int ts = (int)(m % g.Ts); const int n = (int)(m / g.Ts);
        const int tl = ts * g.st - g.pt + ta_, hl = hs * g.sh - g.ph + tb, wl = ws * g.sw - g.pw + tc;
        if (tl >= 0 && tl < g.Tl && hl >= 0 && hl < g.Hl && wl >= 0 && wl < g.Wl)
          ol = ((((int64_t)n * g.Tl + tl) * g.Hl + hl) * g.Wl + wl) * ldl;
      }
      offL[tid] = ol; offS[tid] = os;
    }
    __syncthreads();
    for (int e = tid; e < BK * BA; e += 256) {
      const int pos = e / BA, c = e % BA;
      const int cl = tileA * BA + c;
      float v = 0.f;
      const int64_t o = offL[pos];
      if (o >= 0 && cl < g.Cl) v = ldf(xl + o + cl);
      As[pos][c] = v;
    }
    for (int e = tid; e < BK * BB; e += 256) {
      const int pos = e / BB, c = e % BB;
      const int cs = tileB * BB + c;
      float v = 0.f;
      const int64_t o = offS[pos];
      if (o >= 0 && cs < g.Cs) v = ldf(xs + o + cs);
      Bs[pos][c] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TA], b[TB];
#pragma unroll
      for (int i = 0; i < TA; ++i) a[i] = As[k][ta * TA + i];
#pragma unroll
      for (int j = 0; j < TB; ++j) b[j] = Bs[k][tbb * TB + j];
#pragma unroll
      for (int i = 0; i < TA; ++i)
#pragma unroll
        for (int j = 0; j < TB; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + ((int64_t)blockIdx.z * (g.kt * g.kh * g.kw) + tap) * g.Cl * g.Cs;
#pragma unroll
  for (int i = 0; i < TA; ++i) {
    const int cl = tileA * BA + ta * TA + i;
    if (cl >= g.Cl) continue;
#pragma unroll
    for (int j = 0; j < TB; ++j) {
      const int cs = tileB * BB + tbb * TB + j;
      if (cs < g.Cs) out[(int64_t)cl * g.Cs + cs] = acc[i][j];
    }
  }
}

// dw[cl*s_l + cs*s_s + tap*s_tap] (+)= sum_z partial[z][tap][cl][cs]   (fixed order => deterministic)
// few splits (large weights): one element per thread
__global__ void __launch_bounds__(256)
wgrad_reduce_flat_kernel(const float* __restrict__ partial, int splits, int taps, int Cl, int Cs, WeightWin win,
                         float* __restrict__ dw, int64_t s_l, int64_t s_s, int64_t s_tap, int accumulate) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)taps * Cl * Cs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cs = (int)(i % Cs); const int cl = (int)((i / Cs) % Cl); const int tap = (int)(i / ((int64_t)Cs * Cl));
    if (!win.has(cl, cs)) continue;    // zero-padding channels / other sub-weights have no entry in this master weight
    float s = 0.f;
    const float* pp = partial + i;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {          // eight independent loads in flight, added in split order (deterministic)
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = pp[(int64_t)(z + u) * total];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += pp[(int64_t)z * total];
    float* d = dw + (cl - win.cl_off) * s_l + (cs - win.cs_off) * s_s + tap * s_tap;
    *d = accumulate ? (*d + s) : s;
  }
}

// Transposing form of the flat reduction for master weights whose taps are contiguous (every Conv / ConvTranspose weight:
// [cs][cl][tap] with tap innermost) - the partial sums are [split][tap][cl][cs] with cs innermost, so the flat kernel above,
// one thread per element in partial order, writes 4-byte words Cl*taps floats apart (a 32-byte sector per word:
// wgrad_reduce_flat was 0.49 ms per training iteration, mostly these scattered writes).  Here a block owns one cl and 32
// consecutive cs: it reads the partials with the lanes along cs (128-byte segments), sums the splits in split order
// (deterministic), transposes through shared memory and writes rows of `taps` contiguous floats.
__global__ void __launch_bounds__(256)
wgrad_reduce_tr_kernel(const float* __restrict__ partial, int splits, int taps, int Cl, int Cs, WeightWin win,
                       float* __restrict__ dw, int64_t s_l, int64_t s_s, int accumulate) {
  pdl_wait(); pdl_trigger();
  __shared__ float t[64][33];
  const int cl = blockIdx.x % Cl, cs0 = (blockIdx.x / Cl) * 32;
  if (cl < win.cl_off || cl >= win.cl_off + win.cl_cnt) return;          // zero-padding channel: no entry in the master weight
  const int lane = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int cs = cs0 + lane;
  const int64_t total = (int64_t)taps * Cl * Cs;
  for (int tap = ty; tap < taps; tap += 8) {
    float s = 0.f;
    if (cs < Cs) {
      const float* pp = partial + ((int64_t)tap * Cl + cl) * Cs + cs;
      int z = 0;
      for (; z + 8 <= splits; z += 8) {          // eight independent loads in flight, added in split order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = pp[(int64_t)(z + u) * total];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; z < splits; ++z) s += pp[(int64_t)z * total];
    }
    t[tap][lane] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * taps; i += 256) {
    const int r = i / taps, tp = i % taps;
    const int c = cs0 + r;
    if (c >= Cs || c < win.cs_off || c >= win.cs_off + win.cs_cnt) continue;
    float* d = dw + (int64_t)(cl - win.cl_off) * s_l + (int64_t)(c - win.cs_off) * s_s + tp;
    *d = accumulate ? (*d + t[tp][r]) : t[tp][r];
  }
}

// several master weights that each own a window of the partial matrix (merged / folded stems): every element is summed
// once over the splits and routed to the weight whose window contains it (8 separate reductions re-read the partials 8x)
constexpr int REDUCE_MULTI_MAX = 16;
struct ReducePart { float* dw; int64_t s_l, s_s, s_tap; WeightWin win; int accumulate; };
struct ReduceParts { int n; ReducePart p[REDUCE_MULTI_MAX]; };

__global__ void __launch_bounds__(256)
wgrad_reduce_multi_kernel(const float* __restrict__ partial, int splits, int taps, int Cl, int Cs, const __grid_constant__ ReduceParts parts) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)taps * Cl * Cs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cs = (int)(i % Cs); const int cl = (int)((i / Cs) % Cl); const int tap = (int)(i / ((int64_t)Cs * Cl));
    int pi = -1;
    for (int q = 0; q < parts.n; ++q) if (parts.p[q].win.has(cl, cs)) { pi = q; break; }
    if (pi < 0) continue;
    float s = 0.f;
    const float* pp = partial + i;
    int z = 0;
    for (; z + 8 <= splits; z += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = pp[(int64_t)(z + u) * total];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; ++z) s += pp[(int64_t)z * total];
    const ReducePart& r = parts.p[pi];
    float* d = r.dw + (cl - r.win.cl_off) * r.s_l + (cs - r.win.cs_off) * r.s_s + tap * r.s_tap;
    *d = r.accumulate ? (*d + s) : s;
  }
}

int wgrad_reduce_multi(const float* partial, int splits, const dcv_geom* g, int n, float* const* dw, const int64_t* s_l,
                       const int64_t* s_s, const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off,
                       const int* cs_cnt, const int* accumulate, cudaStream_t s) {
  DCV_REQUIRE(n >= 1 && n <= REDUCE_MULTI_MAX, "wgrad_reduce_multi: %d parts (max %d)", n, REDUCE_MULTI_MAX);
  ReduceParts parts; parts.n = n;
  for (int i = 0; i < n; ++i) {
    ReducePart& r = parts.p[i];
    r.dw = dw[i]; r.s_l = s_l[i]; r.s_s = s_s[i]; r.s_tap = s_tap[i]; r.accumulate = accumulate[i];
    r.win.cl_off = cl_off[i]; r.win.cl_cnt = cl_cnt[i]; r.win.cs_off = cs_off[i]; r.win.cs_cnt = cs_cnt[i]; r.win.fill = 0;
  }
  const int taps = g->kt * g->kh * g->kw;
  const int64_t total = (int64_t)taps * g->Cl * g->Cs;
  int fb = (int)((total + 255) / 256); if (fb > 148 * 16) fb = 148 * 16;
  launch_k(wgrad_reduce_multi_kernel, fb, 256, 0, s, partial, splits, taps, g->Cl, g->Cs, parts);
  return check_launch("wgrad_reduce_multi");
}

// many splits (small weights, huge pixel counts): block = 8 consecutive elements x 32 split lanes; lane y sums splits
// y, y+32, ... (32-byte sectors), the 32 partial sums are then added in a fixed order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int Cl, int Cs, WeightWin win,
                    float* __restrict__ dw, int64_t s_l, int64_t s_s, int64_t s_tap, int accumulate) {
  pdl_wait(); pdl_trigger();
  __shared__ float red[32][9];
  const int64_t total = (int64_t)taps * Cl * Cs;
  const int tx = threadIdx.x % 8, ty = threadIdx.x / 8;
  for (int64_t base = (int64_t)blockIdx.x * 8; base < total; base += (int64_t)gridDim.x * 8) {
    const int64_t i = base + tx;
    float s0 = 0.f, s1 = 0.f;
    if (i < total) {
      int z = ty;
      for (; z + 32 < splits; z += 64) { s0 += partial[(int64_t)z * total + i]; s1 += partial[(int64_t)(z + 32) * total + i]; }
      if (z < splits) s0 += partial[(int64_t)z * total + i];
    }
    red[ty][tx] = s0 + s1;
    __syncthreads();
    if (ty == 0 && i < total) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) t += red[k][tx];
      const int cs = (int)(i % Cs); const int cl = (int)((i / Cs) % Cl); const int tap = (int)(i / ((int64_t)Cs * Cl));
      if (win.has(cl, cs)) {
        float* d = dw + (cl - win.cl_off) * s_l + (cs - win.cs_off) * s_s + tap * s_tap;
        *d = accumulate ? (*d + t) : t;
      }
    }
    __syncthreads();
  }
}

__global__ void pack_weight_simt_kernel(const float* __restrict__ w, int64_t s_l, int64_t s_s, int64_t s_tap,
                                        int Cl, int Cs, WeightWin win, int taps, int scatter, float* __restrict__ out) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)taps * Cl * Cs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int tap, cl, cs;
    if (!scatter) { cs = (int)(i % Cs); cl = (int)((i / Cs) % Cl); tap = (int)(i / ((int64_t)Cs * Cl)); }
    else          { cl = (int)(i % Cl); cs = (int)((i / Cl) % Cs); tap = (int)(i / ((int64_t)Cs * Cl)); }
    if (win.has(cl, cs)) out[i] = w[(cl - win.cl_off) * s_l + (cs - win.cs_off) * s_s + tap * s_tap];
    else if (win.fill) out[i] = 0.f;
  }
}

template <typename T>
static int launch_conv_simt(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y,
                            int64_t ldy, int act, float slope, cudaStream_t s) {
  ConvP p = make_convp(g, dir);
  int phases = p.scatter ? g->st * g->sh * g->sw : 1;
  int64_t Mmax;
  if (!p.scatter) Mmax = (int64_t)p.N * p.Ot * p.Oh * p.Ow;
  else Mmax = (int64_t)p.N * ceil_div(p.Ot, g->st) * ceil_div(p.Oh, g->sh) * ceil_div(p.Ow, g->sw);
  if (Mmax == 0) return 0;
  if (p.Nc <= 8) {
    dim3 grid(ceil_div(Mmax, 256), ceil_div(p.Nc, 4), phases);
    launch_k(conv_simt_kernel<T, 256, 4, 1, 4>, grid, 256, 0, s, p, (const T*)x, ldx, (const float*)wp, (T*)y, ldy, act, slope);
  } else {
    dim3 grid(ceil_div(Mmax, 64), ceil_div(p.Nc, 64), phases);
    launch_k(conv_simt_kernel<T, 64, 64, 4, 4>, grid, 256, 0, s, p, (const T*)x, ldx, (const float*)wp, (T*)y, ldy, act, slope);
  }
  return check_launch("conv_simt");
}

int conv_simt(const dcv_geom* g, int dir, int dtype, const void* x, int64_t ldx, const void* wp, void* y,
              int64_t ldy, int act, float slope, cudaStream_t s) {
  if (dtype == DCV_F32) return launch_conv_simt<float>(g, dir, x, ldx, wp, y, ldy, act, slope, s);
  return launch_conv_simt<__nv_bfloat16>(g, dir, x, ldx, wp, y, ldy, act, slope, s);
}

int wgrad_reduce_win(const float* partial, int splits, const dcv_geom* g, WeightWin win, float* dw, int64_t s_l,
                     int64_t s_s, int64_t s_tap, int accumulate, cudaStream_t s) {
  const int taps = g->kt * g->kh * g->kw;
  const int64_t total = (int64_t)taps * g->Cl * g->Cs;
  if (s_tap == 1 && taps <= 64 && total >= 16384) {      // tap-contiguous master weight: transposing reduction, coalesced both ways
    launch_k(wgrad_reduce_tr_kernel, g->Cl * ceil_div(g->Cs, 32), 256, 0, s, partial, splits, taps, g->Cl, g->Cs, win, dw, s_l, s_s, accumulate);
    return check_launch("wgrad_reduce");
  }
  if (splits <= 16 || total >= 16384) {      // enough elements to fill the machine with one thread per element
    int fb = (int)((total + 255) / 256); if (fb > 148 * 16) fb = 148 * 16;
    launch_k(wgrad_reduce_flat_kernel, fb, 256, 0, s, partial, splits, taps, g->Cl, g->Cs, win, dw, s_l, s_s, s_tap, accumulate);
    return check_launch("wgrad_reduce");
  }
  int blocks = (int)((total + 7) / 8); if (blocks > 148 * 16) blocks = 148 * 16;
  launch_k(wgrad_reduce_kernel, blocks, 256, 0, s, partial, splits, taps, g->Cl, g->Cs, win, dw, s_l, s_s, s_tap, accumulate);
  return check_launch("wgrad_reduce");
}

int wgrad_reduce(const float* partial, int splits, const dcv_geom* g, float* dw, int64_t s_l, int64_t s_s,
                 int64_t s_tap, int accumulate, cudaStream_t s) {
  return wgrad_reduce_win(partial, splits, g, full_window(g), dw, s_l, s_s, s_tap, accumulate, s);
}

int wgrad_simt_splits(const dcv_geom* g) {
  const int taps = g->kt * g->kh * g->kw;
  int tilesA, tilesB;
  if (g->Cl <= 8) { tilesA = ceil_div(g->Cl, 4); tilesB = ceil_div(g->Cs, 64); }
  else if (g->Cs <= 8) { tilesA = ceil_div(g->Cl, 64); tilesB = ceil_div(g->Cs, 4); }
  else if (g->Cl <= 16) { tilesA = 1; tilesB = ceil_div(g->Cs, 64); }
  else if (g->Cs <= 16) { tilesA = ceil_div(g->Cl, 64); tilesB = 1; }
  else { tilesA = ceil_div(g->Cl, 64); tilesB = ceil_div(g->Cs, 64); }
  const int64_t M = (int64_t)g->N * g->Ts * g->Hs * g->Ws;
  int64_t base = (int64_t)tilesA * tilesB * taps;
  int64_t splits = (148 * 4 + base - 1) / base;
  int64_t maxs = (M + 255) / 256;
  if (splits > maxs) splits = maxs;
  if (splits > 64) splits = 64;
  if (splits < 1) splits = 1;
  return (int)splits;
}

int64_t wgrad_simt_ws_bytes(const dcv_geom* g) {
  return (int64_t)wgrad_simt_splits(g) * g->kt * g->kh * g->kw * g->Cl * g->Cs * sizeof(float);
}

template <typename T>
static int launch_wgrad_simt(const dcv_geom* g, const void* xl, int64_t ldl, const void* xs, int64_t lds,
                             float* partial, int splits, cudaStream_t s) {
  WgradP p;
  p.g = *g;
  p.M = (int64_t)g->N * g->Ts * g->Hs * g->Ws;
  p.m_per_split = ((p.M + splits - 1) / splits + BK - 1) / BK * BK;
  const int taps = g->kt * g->kh * g->kw;
  if (g->Cl <= 8) {
    p.tilesB = ceil_div(g->Cs, 64);
    dim3 grid(ceil_div(g->Cl, 4) * p.tilesB, taps, splits);
    launch_k(wgrad_simt_kernel<T, 4, 64, 1, 1>, grid, 256, 0, s, p, (const T*)xl, ldl, (const T*)xs, lds, partial);
  } else if (g->Cs <= 8) {
    p.tilesB = ceil_div(g->Cs, 4);
    dim3 grid(ceil_div(g->Cl, 64) * p.tilesB, taps, splits);
    launch_k(wgrad_simt_kernel<T, 64, 4, 1, 1>, grid, 256, 0, s, p, (const T*)xl, ldl, (const T*)xs, lds, partial);
  } else if (g->Cl <= 16) {
    p.tilesB = ceil_div(g->Cs, 64);
    dim3 grid(p.tilesB, taps, splits);
    launch_k(wgrad_simt_kernel<T, 16, 64, 2, 2>, grid, 256, 0, s, p, (const T*)xl, ldl, (const T*)xs, lds, partial);
  } else if (g->Cs <= 16) {
    p.tilesB = 1;
    dim3 grid(ceil_div(g->Cl, 64), taps, splits);
    launch_k(wgrad_simt_kernel<T, 64, 16, 2, 2>, grid, 256, 0, s, p, (const T*)xl, ldl, (const T*)xs, lds, partial);
  } else {
    p.tilesB = ceil_div(g->Cs, 64);
    dim3 grid(ceil_div(g->Cl, 64) * p.tilesB, taps, splits);
    launch_k(wgrad_simt_kernel<T, 64, 64, 4, 4>, grid, 256, 0, s, p, (const T*)xl, ldl, (const T*)xs, lds, partial);
  }
  return check_launch("wgrad_simt");
}

int wgrad_simt_partial(const dcv_geom* g, int dtype, const void* xl, int64_t ldl, const void* xs, int64_t lds, void* ws,
                       int64_t ws_bytes, cudaStream_t s) {
  const int splits = wgrad_simt_splits(g);
  DCV_REQUIRE(ws_bytes >= wgrad_simt_ws_bytes(g), "wgrad workspace too small: %lld < %lld",
              (long long)ws_bytes, (long long)wgrad_simt_ws_bytes(g));
  return dtype == DCV_F32 ? launch_wgrad_simt<float>(g, xl, ldl, xs, lds, (float*)ws, splits, s)
                          : launch_wgrad_simt<__nv_bfloat16>(g, xl, ldl, xs, lds, (float*)ws, splits, s);
}

int wgrad_simt(const dcv_geom* g, int dtype, const void* xl, int64_t ldl, const void* xs, int64_t lds,
               float* dw, int64_t s_l, int64_t s_s, int64_t s_tap, int accumulate, void* ws,
               int64_t ws_bytes, cudaStream_t s) {
  const int splits = wgrad_simt_splits(g);
  DCV_REQUIRE(ws_bytes >= wgrad_simt_ws_bytes(g), "wgrad workspace too small: %lld < %lld",
              (long long)ws_bytes, (long long)wgrad_simt_ws_bytes(g));
  int rc = dtype == DCV_F32 ? launch_wgrad_simt<float>(g, xl, ldl, xs, lds, (float*)ws, splits, s)
                            : launch_wgrad_simt<__nv_bfloat16>(g, xl, ldl, xs, lds, (float*)ws, splits, s);
  if (rc) return rc;
  return wgrad_reduce((const float*)ws, splits, g, dw, s_l, s_s, s_tap, accumulate, s);
}

int pack_weight_simt(const dcv_geom* g, int dir, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                     WeightWin win, float* out, cudaStream_t s) {
  const int taps = g->kt * g->kh * g->kw;
  const int64_t total = (int64_t)taps * g->Cl * g->Cs;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(pack_weight_simt_kernel, blocks, 256, 0, s, w, s_l, s_s, s_tap, g->Cl, g->Cs, win, taps, dir == DCV_DIR_SCATTER, out);
  return check_launch("pack_weight_simt");
}

}  // namespace dcv
