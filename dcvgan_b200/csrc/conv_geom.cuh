// Index arithmetic shared by the CUDA-core and tcgen05 correlation kernels.
#pragma once
#include "common.cuh"

namespace dcv {

struct ConvP {
  int N;
  int It, Ih, Iw;     // extents of the tensor being read
  int Ot, Oh, Ow;     // extents of the tensor being written
  int kt, kh, kw;
  int st, sh, sw;
  int pt, ph, pw;
  int Kc, Nc;         // reduction channels, produced channels (incl. zero padding)
  int wK, wN;         // of which present in the master weight
  int scatter;
};

// Window of a (possibly zero-padded, possibly merged) weight matrix that one master weight tensor occupies:
// activation channels cl in [cl_off, cl_off+cl_cnt) x cs in [cs_off, cs_off+cs_cnt).  fill: packers write zeros outside.
struct WeightWin {
  int cl_off, cl_cnt, cs_off, cs_cnt, fill;
  __host__ __device__ bool has(int cl, int cs) const {
    return cl >= cl_off && cl < cl_off + cl_cnt && cs >= cs_off && cs < cs_off + cs_cnt;
  }
};
static inline WeightWin full_window(const dcv_geom* g) {
  WeightWin w;
  w.cl_off = 0; w.cl_cnt = g->wCl > 0 ? g->wCl : g->Cl;
  w.cs_off = 0; w.cs_cnt = g->wCs > 0 ? g->wCs : g->Cs;
  w.fill = 1;
  return w;
}

struct PhaseInfo {
  int Qt, Qh, Qw;       // index-space extents of this phase
  int nt, nh, nw;       // tap counts
  int a0t, a0h, a0w;    // first tap
  int ast, ash, asw;    // tap step
  int mult, mulh, mulw; // input coord = o*mul + off + sgn*j
  int offt, offh, offw;
  int sgn;
  int rt, rh, rw;       // residues (scatter) for the output coordinate
  int ost, osh, osw;    // output coord = o*ost + r
};

__host__ __device__ __forceinline__ PhaseInfo make_phase(const ConvP& p, int phase) {
  PhaseInfo f;
  if (!p.scatter) {
    f.Qt = p.Ot; f.Qh = p.Oh; f.Qw = p.Ow;
    f.nt = p.kt; f.nh = p.kh; f.nw = p.kw;
    f.a0t = f.a0h = f.a0w = 0;
    f.ast = f.ash = f.asw = 1;
    f.mult = p.st; f.mulh = p.sh; f.mulw = p.sw;
    f.offt = -p.pt; f.offh = -p.ph; f.offw = -p.pw;
    f.sgn = 1;
    f.rt = f.rh = f.rw = 0;
    f.ost = f.osh = f.osw = 1;
  } else {
    int rw = phase % p.sw; int rh = (phase / p.sw) % p.sh; int rt = phase / (p.sw * p.sh);
    f.rt = rt; f.rh = rh; f.rw = rw;
    f.Qt = (p.Ot - rt + p.st - 1) / p.st; f.Qh = (p.Oh - rh + p.sh - 1) / p.sh; f.Qw = (p.Ow - rw + p.sw - 1) / p.sw;
    f.a0t = (rt + p.pt) % p.st; f.a0h = (rh + p.ph) % p.sh; f.a0w = (rw + p.pw) % p.sw;
    f.ast = p.st; f.ash = p.sh; f.asw = p.sw;
    f.nt = f.a0t < p.kt ? (p.kt - f.a0t + p.st - 1) / p.st : 0;
    f.nh = f.a0h < p.kh ? (p.kh - f.a0h + p.sh - 1) / p.sh : 0;
    f.nw = f.a0w < p.kw ? (p.kw - f.a0w + p.sw - 1) / p.sw : 0;
    f.mult = f.mulh = f.mulw = 1;
    f.offt = (rt + p.pt - f.a0t) / p.st; f.offh = (rh + p.ph - f.a0h) / p.sh; f.offw = (rw + p.pw - f.a0w) / p.sw;
    f.sgn = -1;
    f.ost = p.st; f.osh = p.sh; f.osw = p.sw;
  }
  return f;
}

static inline ConvP make_convp(const dcv_geom* g, int dir) {
  ConvP p;
  p.N = g->N;
  p.kt = g->kt; p.kh = g->kh; p.kw = g->kw;
  p.st = g->st; p.sh = g->sh; p.sw = g->sw;
  p.pt = g->pt; p.ph = g->ph; p.pw = g->pw;
  p.scatter = dir == DCV_DIR_SCATTER;
  if (!p.scatter) {
    p.It = g->Tl; p.Ih = g->Hl; p.Iw = g->Wl; p.Ot = g->Ts; p.Oh = g->Hs; p.Ow = g->Ws;
    p.Kc = g->Cl; p.Nc = g->Cs;
    p.wK = g->wCl > 0 ? g->wCl : g->Cl; p.wN = g->wCs > 0 ? g->wCs : g->Cs;
  } else {
    p.It = g->Ts; p.Ih = g->Hs; p.Iw = g->Ws; p.Ot = g->Tl; p.Oh = g->Hl; p.Ow = g->Wl;
    p.Kc = g->Cs; p.Nc = g->Cl;
    p.wK = g->wCs > 0 ? g->wCs : g->Cs; p.wN = g->wCl > 0 ? g->wCl : g->Cl;
  }
  return p;
}


}  // namespace dcv
