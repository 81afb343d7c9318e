// C-ABI glue: error reporting and the convolution-family dispatchers.
#include "common.cuh"
#include "conv_geom.cuh"
#include <string.h>

namespace dcv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

// implemented in simt_conv.cu / tc_conv.cu
int conv_simt(const dcv_geom*, int, int, const void*, int64_t, const void*, void*, int64_t, int, float, cudaStream_t);
int wgrad_simt(const dcv_geom*, int, const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int64_t, int,
               void*, int64_t, cudaStream_t);
int64_t wgrad_simt_ws_bytes(const dcv_geom*);
int pack_weight_simt(const dcv_geom*, int, const float*, int64_t, int64_t, int64_t, WeightWin, float*, cudaStream_t);
int wgrad_simt_partial(const dcv_geom*, int, const void*, int64_t, const void*, int64_t, void*, int64_t, cudaStream_t);
int wgrad_simt_splits(const dcv_geom*);
int wgrad_tc_splits(const dcv_geom*);
int wgrad_reduce_win(const float*, int, const dcv_geom*, WeightWin, float*, int64_t, int64_t, int64_t, int, cudaStream_t);
int wgrad_reduce_multi(const float*, int, const dcv_geom*, int, float* const*, const int64_t*, const int64_t*, const int64_t*, const int*,
                       const int*, const int*, const int*, const int*, cudaStream_t);
int conv_tc_supported(const dcv_geom*, int);
int set_tuning(const char*, int);
int img_conv_supported_for(const dcv_geom*, int);
int img_conv_scatter(const dcv_geom*, const void*, int64_t, const float*, int64_t, int64_t, int64_t, void*, int64_t, int, float, const dcv_prebn*, cudaStream_t);
int64_t img_conv_bwd_ws_bytes(const dcv_geom*);
int img_conv_fwd(const dcv_geom*, const void*, int64_t, const float*, int64_t, int64_t, int64_t, void*, int64_t, int, float, cudaStream_t);
int img_conv_bwd(const dcv_geom*, const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, int64_t, int64_t,
                 int64_t, int, float, float*, int, void*, int64_t, void*, int64_t, cudaStream_t);
int pack_weight_tc_multi(const dcv_geom*, int, int, const float* const*, const int64_t*, const int64_t*, const int64_t*, const int*,
                         const int*, const int*, const int*, void*, cudaStream_t);
int pack_weight_tc_batch(int, const dcv_geom* const*, const int*, const float* const*, const int64_t*, const int64_t*, const int64_t*,
                         void* const*, cudaStream_t);
int64_t packed_weight_tc_bytes(const dcv_geom*, int, int);
int conv_tf32_supported(const dcv_geom*, int);
int head_loss_supported(const dcv_geom*, int);
int head_loss(const dcv_geom*, int, const void*, int64_t, const void*, void*, int64_t, int, float*, int, void*, int64_t, float, unsigned*, cudaStream_t);
int pack_weight_tc(const dcv_geom*, int, const float*, int64_t, int64_t, int64_t, WeightWin, void*, cudaStream_t, int);
int conv_tc(const dcv_geom*, int, const void*, int64_t, const void*, void*, int64_t, int, float, float*, int*, cudaStream_t, int, int);
int wgrad_tc_supported(const dcv_geom*);
int64_t wgrad_tc_ws_bytes(const dcv_geom*);
int wgrad_tc(const dcv_geom*, const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int64_t, int, void*,
             int64_t, cudaStream_t);

static int check_geom(const dcv_geom* g) {
  DCV_REQUIRE(g, "null geometry");
  DCV_REQUIRE(g->N >= 0 && g->Cl > 0 && g->Cs > 0, "bad geometry: N=%d Cl=%d Cs=%d", g->N, g->Cl, g->Cs);
  DCV_REQUIRE(g->wCl >= 0 && g->wCl <= g->Cl && g->wCs >= 0 && g->wCs <= g->Cs, "bad weight channel counts wCl=%d wCs=%d", g->wCl, g->wCs);
  DCV_REQUIRE(g->kt > 0 && g->kh > 0 && g->kw > 0 && g->st > 0 && g->sh > 0 && g->sw > 0, "bad kernel/stride");
  DCV_REQUIRE(g->Tl > 0 && g->Hl > 0 && g->Wl > 0 && g->Ts > 0 && g->Hs > 0 && g->Ws > 0, "bad extents");
  // S extent must be what a convolution of L produces: floor((L + 2p - k)/s) + 1
  DCV_REQUIRE(g->Ts == (g->Tl + 2 * g->pt - g->kt) / g->st + 1 && g->Hs == (g->Hl + 2 * g->ph - g->kh) / g->sh + 1 &&
              g->Ws == (g->Wl + 2 * g->pw - g->kw) / g->sw + 1, "inconsistent L/S extents");
  return 0;
}

}  // namespace dcv

using namespace dcv;

extern "C" {

int dcv_abi_version(void) { return DCV_ABI_VERSION; }
const char* dcv_last_error(void) { return g_err; }

long long dcv_launch_count(void) { return g_launches; }

int dcv_head_loss_supported(const dcv_geom* g, int dir) { return g && check_geom(g) == 0 ? head_loss_supported(g, dir) : 0; }
int dcv_head_loss(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int kind, float* loss_out,
                  int accumulate, void* dy, int64_t lddy, float grad_scale, void* counter, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(x && wp && y && loss_out && counter, "head_loss: null pointer");
  return head_loss(g, dir, x, ldx, wp, y, ldy, kind, loss_out, accumulate, dy, lddy, grad_scale, (unsigned*)counter, as_stream(stream));
}
int dcv_conv_tf32_supported(const dcv_geom* g, int dir) { return g && check_geom(g) == 0 ? conv_tf32_supported(g, dir) : 0; }
int dcv_img_conv_supported(const dcv_geom* g, int what) { return g && check_geom(g) == 0 ? img_conv_supported_for(g, what) : 0; }
int64_t dcv_img_conv_bwd_workspace_bytes(const dcv_geom* g) { return g ? img_conv_bwd_ws_bytes(g) : -1; }
int dcv_img_conv_fwd(const dcv_geom* g, const void* x, int64_t ldx, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                     void* y, int64_t ldy, int act, float slope, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(x && w && y, "img_conv_fwd: null pointer");
  return img_conv_fwd(g, x, ldx, w, s_l, s_s, s_tap, y, ldy, act, slope, as_stream(stream));
}
int dcv_img_conv_scatter(const dcv_geom* g, const void* xb, int64_t ldb, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                         void* y, int64_t ldy, int act, float slope, const dcv_prebn* pre, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(xb && w && y, "img_conv_scatter: null pointer");
  return img_conv_scatter(g, xb, ldb, w, s_l, s_s, s_tap, y, ldy, act, slope, pre, as_stream(stream));
}
int dcv_img_conv_bwd(const dcv_geom* g, const void* da, int64_t ldda, const void* a, int64_t lda, const void* x, int64_t ldx,
                     const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, int act, float slope, float* dw, int accumulate,
                     void* dx, int64_t lddx, void* ws, int64_t ws_bytes, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(da && a && x && w && ws, "img_conv_bwd: null pointer");
  return img_conv_bwd(g, da, ldda, a, lda, x, ldx, w, s_l, s_s, s_tap, act, slope, dw, accumulate, dx, lddx, ws, ws_bytes,
                      as_stream(stream));
}

int dcv_set_tuning(const char* key, int value) { DCV_REQUIRE(key, "dcv_set_tuning: null key"); return set_tuning(key, value); }

int dcv_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return 0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return 0; }
  if (prop.major != 10) { set_error("device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor); return 0; }
  return 1;
}

int64_t dcv_packed_weight_bytes(const dcv_geom* g, int dir, int impl) {
  if (check_geom(g)) return -1;
  if (impl == DCV_IMPL_TC || impl == DCV_IMPL_TC_TF32) return packed_weight_tc_bytes(g, dir, impl == DCV_IMPL_TC_TF32);
  return (int64_t)g->kt * g->kh * g->kw * g->Cl * g->Cs * sizeof(float);
}

int dcv_pack_weight(const dcv_geom* g, int dir, int impl, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                    void* out, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(w && out, "pack_weight: null pointer");
  if (impl == DCV_IMPL_TC || impl == DCV_IMPL_TC_TF32)
    return pack_weight_tc(g, dir, w, s_l, s_s, s_tap, full_window(g), out, as_stream(stream), impl == DCV_IMPL_TC_TF32);
  return pack_weight_simt(g, dir, w, s_l, s_s, s_tap, full_window(g), (float*)out, as_stream(stream));
}

static int check_window(const dcv_geom* g, int cl_off, int cl_cnt, int cs_off, int cs_cnt) {
  DCV_REQUIRE(cl_off >= 0 && cl_cnt > 0 && cl_off + cl_cnt <= g->Cl && cs_off >= 0 && cs_cnt > 0 && cs_off + cs_cnt <= g->Cs,
              "weight window [%d,+%d) x [%d,+%d) outside %d x %d", cl_off, cl_cnt, cs_off, cs_cnt, g->Cl, g->Cs);
  return 0;
}

int dcv_pack_weight_sub(const dcv_geom* g, int dir, int impl, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap,
                        int cl_off, int cl_cnt, int cs_off, int cs_cnt, int fill_outside, void* out, void* stream) {
  if (int rc = check_geom(g)) return rc;
  if (int rc = check_window(g, cl_off, cl_cnt, cs_off, cs_cnt)) return rc;
  DCV_REQUIRE(w && out, "pack_weight_sub: null pointer");
  WeightWin win; win.cl_off = cl_off; win.cl_cnt = cl_cnt; win.cs_off = cs_off; win.cs_cnt = cs_cnt; win.fill = fill_outside;
  if (impl == DCV_IMPL_TC) return pack_weight_tc(g, dir, w, s_l, s_s, s_tap, win, out, as_stream(stream), 0);
  return pack_weight_simt(g, dir, w, s_l, s_s, s_tap, win, (float*)out, as_stream(stream));
}

int dcv_pack_weight_batch(int n, const dcv_geom* const* geoms, const int* dirs, const float* const* w, const int64_t* s_l,
                          const int64_t* s_s, const int64_t* s_tap, void* const* outs, void* stream) {
  DCV_REQUIRE(n >= 0 && (n == 0 || (geoms && dirs && w && s_l && s_s && s_tap && outs)), "pack_weight_batch: null array");
  for (int i = 0; i < n; ++i) {
    if (int rc = check_geom(geoms[i])) return rc;
    DCV_REQUIRE(w[i] && outs[i], "pack_weight_batch: null pointer in job %d", i);
  }
  return pack_weight_tc_batch(n, geoms, dirs, w, s_l, s_s, s_tap, outs, as_stream(stream));
}

int dcv_pack_weight_multi(const dcv_geom* g, int dir, int n, const float* const* w, const int64_t* s_l, const int64_t* s_s,
                          const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off, const int* cs_cnt,
                          void* out, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(w && s_l && s_s && s_tap && cl_off && cl_cnt && cs_off && cs_cnt && out, "pack_weight_multi: null pointer");
  for (int i = 0; i < n; ++i) {
    if (int rc = check_window(g, cl_off[i], cl_cnt[i], cs_off[i], cs_cnt[i])) return rc;
    DCV_REQUIRE(w[i], "pack_weight_multi: null weight pointer in part %d", i);
  }
  return pack_weight_tc_multi(g, dir, n, w, s_l, s_s, s_tap, cl_off, cl_cnt, cs_off, cs_cnt, out, as_stream(stream));
}

int dcv_conv_tc_supported(const dcv_geom* g, int dir) {
  if (check_geom(g)) return 0;
  return conv_tc_supported(g, dir);
}

int dcv_conv(const dcv_geom* g, int dir, int impl, int dtype, const void* x, int64_t ldx, const void* wp, void* y,
             int64_t ldy, int act, float slope, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(x && wp && y, "conv: null pointer");
  DCV_REQUIRE(dtype == DCV_F32 || dtype == DCV_BF16, "conv: bad dtype %d", dtype);
  if (g->N == 0) return 0;
  if (impl == DCV_IMPL_TC) {
    DCV_REQUIRE(dtype == DCV_BF16, "conv: DCV_IMPL_TC computes in bf16 (DCV_IMPL_TC_TF32 takes fp32 tensors)");
    return conv_tc(g, dir, x, ldx, wp, y, ldy, act, slope, nullptr, nullptr, as_stream(stream), 0, 0);
  }
  if (impl == DCV_IMPL_TC_TF32) {
    DCV_REQUIRE(dtype == DCV_F32, "conv: DCV_IMPL_TC_TF32 takes fp32 tensors");
    return conv_tc(g, dir, x, ldx, wp, y, ldy, act, slope, nullptr, nullptr, as_stream(stream), 1, 0);
  }
  return conv_simt(g, dir, dtype, x, ldx, wp, y, ldy, act, slope, as_stream(stream));
}

int dcv_conv_accumulate(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(x && wp && y, "conv_accumulate: null pointer");
  if (g->N == 0) return 0;
  return conv_tc(g, dir, x, ldx, wp, y, ldy, DCV_ACT_NONE, 0.f, nullptr, nullptr, as_stream(stream), 0, 1);
}

int dcv_conv_stats_slots(const dcv_geom* g, int dir, int64_t ldx, int64_t ldy) {
  if (check_geom(g)) return -1;
  if (g->N == 0 || !conv_tc_supported(g, dir)) return 0;
  int slots = 0;
  // planning query: aligned dummy pointers, nothing is dereferenced or launched
  if (conv_tc(g, dir, (const void*)16, ldx, nullptr, (void*)16, ldy, 0, 0.f, nullptr, &slots, nullptr, 0, 0)) return -1;
  return slots;
}

int dcv_conv_stats(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int act,
                   float slope, float* stats, int slots, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(x && wp && y && stats, "conv_stats: null pointer");
  DCV_REQUIRE(g->N > 0, "conv_stats: empty batch");
  DCV_REQUIRE(((uintptr_t)y & 15) == 0, "conv_stats: output must be 16-byte aligned");
  const int want = dcv_conv_stats_slots(g, dir, ldx, ldy);
  DCV_REQUIRE(want > 0 && want == slots, "conv_stats: this geometry writes %d statistic slots, caller provided %d", want, slots);
  return conv_tc(g, dir, x, ldx, wp, y, ldy, act, slope, stats, nullptr, as_stream(stream), 0, 0);
}

int64_t dcv_wgrad_workspace_bytes(const dcv_geom* g, int impl) {
  if (check_geom(g)) return -1;
  return impl == DCV_IMPL_TC ? wgrad_tc_ws_bytes(g) : wgrad_simt_ws_bytes(g);
}

int dcv_wgrad_tc_supported(const dcv_geom* g) {
  if (check_geom(g)) return 0;
  return wgrad_tc_supported(g);
}

int dcv_wgrad_partial(const dcv_geom* g, int impl, int dtype, const void* xl, int64_t ldl, const void* xs, int64_t lds,
                      void* ws, int64_t ws_bytes, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(xl && xs && ws, "wgrad_partial: null pointer");
  DCV_REQUIRE(g->N > 0, "wgrad_partial: empty batch");
  if (impl == DCV_IMPL_TC) {
    DCV_REQUIRE(dtype == DCV_BF16, "wgrad: the tcgen05 kernel computes in bf16");
    return wgrad_tc(g, xl, ldl, xs, lds, nullptr, 0, 0, 0, 0, ws, ws_bytes, as_stream(stream));
  }
  return wgrad_simt_partial(g, dtype, xl, ldl, xs, lds, ws, ws_bytes, as_stream(stream));
}

int dcv_wgrad_reduce_sub(const dcv_geom* g, int impl, const void* ws, float* dw, int64_t s_l, int64_t s_s, int64_t s_tap,
                         int cl_off, int cl_cnt, int cs_off, int cs_cnt, int accumulate, void* stream) {
  if (int rc = check_geom(g)) return rc;
  if (int rc = check_window(g, cl_off, cl_cnt, cs_off, cs_cnt)) return rc;
  DCV_REQUIRE(ws && dw, "wgrad_reduce_sub: null pointer");
  WeightWin win; win.cl_off = cl_off; win.cl_cnt = cl_cnt; win.cs_off = cs_off; win.cs_cnt = cs_cnt; win.fill = 0;
  const int splits = impl == DCV_IMPL_TC ? wgrad_tc_splits(g) : wgrad_simt_splits(g);
  return wgrad_reduce_win((const float*)ws, splits, g, win, dw, s_l, s_s, s_tap, accumulate, as_stream(stream));
}

int dcv_wgrad_reduce_multi(const dcv_geom* g, int impl, const void* ws, int n, float* const* dw, const int64_t* s_l,
                           const int64_t* s_s, const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off,
                           const int* cs_cnt, const int* accumulate, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(ws && dw && s_l && s_s && s_tap && cl_off && cl_cnt && cs_off && cs_cnt && accumulate, "wgrad_reduce_multi: null pointer");
  for (int i = 0; i < n; ++i) {
    if (int rc = check_window(g, cl_off[i], cl_cnt[i], cs_off[i], cs_cnt[i])) return rc;
    DCV_REQUIRE(dw[i], "wgrad_reduce_multi: null gradient pointer in part %d", i);
  }
  const int splits = impl == DCV_IMPL_TC ? wgrad_tc_splits(g) : wgrad_simt_splits(g);
  return wgrad_reduce_multi((const float*)ws, splits, g, n, dw, s_l, s_s, s_tap, cl_off, cl_cnt, cs_off, cs_cnt, accumulate,
                            as_stream(stream));
}

int dcv_wgrad(const dcv_geom* g, int impl, int dtype, const void* xl, int64_t ldl, const void* xs, int64_t lds, float* dw,
              int64_t s_l, int64_t s_s, int64_t s_tap, int accumulate, void* ws, int64_t ws_bytes, void* stream) {
  if (int rc = check_geom(g)) return rc;
  DCV_REQUIRE(xl && xs && dw && ws, "wgrad: null pointer");
  DCV_REQUIRE(g->N > 0, "wgrad: empty batch");
  if (impl == DCV_IMPL_TC) {
    DCV_REQUIRE(dtype == DCV_BF16, "wgrad: the tcgen05 kernel computes in bf16");
    return wgrad_tc(g, xl, ldl, xs, lds, dw, s_l, s_s, s_tap, accumulate, ws, ws_bytes, as_stream(stream));
  }
  return wgrad_simt(g, dtype, xl, ldl, xs, lds, dw, s_l, s_s, s_tap, accumulate, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
