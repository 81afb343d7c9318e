// tcgen05 / TMEM / TMA implicit-GEMM correlation kernels for sm_100a (bf16 in, fp32 accumulate).
//
//   conv_tc_pers_kernel<KS, MT, HG>
//               : forward / data-gradient of every Conv2d, ConvTranspose2d and Conv3d on the path whose reduction
//                 channel count is a multiple of 16.  Persistent (one CTA per SM, contiguous item ranges, odometer tile
//                 iteration), double-buffered TMEM accumulators, warp-uniform lean MMA issue, TMA-store epilogue.
//                 A-operand tiles (128 output positions x CBLK channels) are fetched straight from the channels-last
//                 activation by one 5-D TMA box per (tap, channel block) - traversal strides implement the (1,2,2)
//                 convolution stride, out-of-bounds coordinates implement zero padding - so no im2col matrix ever exists
//                 in HBM.  Transposed (scatter) correlations run as stride^d sub-pixel phases.  HG > 1: the h taps of one
//                 lattice share a box with halo rows; KS == 0: four 16-channel taps per stage (image-like operands).
//   wgrad_tc    : weight gradient.  Both operands are MN-major (channels contiguous, pixels = GEMM K), again
//                 straight from the activations by TMA; one wave of CTAs split over pixel ranges, fp32 partials reduced
//                 in a fixed order by wgrad_reduce (simt_conv.cu).
//   pack_weight_tc_* : fp32 master weights -> bf16 K-major operand (single, windowed, multi-window, batched).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (tcgen05.ld -> registers -> swizzled smem -> TMA store).
// What limited these kernels and in which order it was found: DESIGN.md section 5, profiles/r1_summary.md.
#include "common.cuh"
#include "conv_geom.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace dcv {

// Tuning / timing-experiment hooks (environment overrides, "stop loading A / B", "skip the stores") exist only in the
// -DDCV_EXPERIMENTS build that tools/exp_*.py use (csrc/build.sh exp); the product library has none of them, so no
// environment variable can change what a production kernel loads or stores.
// Result-preserving tuning switches (row-halo sharing off, forced M-tile count, direct-store epilogue, extra wgrad split
// waves ...) are explicit process state set through dcv_set_tuning(), never read from the environment; the parity tests
// flip them to cover every code path of the kernels (tests/test_ops_gpu.py::test_conv_tcgen05_kernel_variants).
struct Tuning { int nohalo, mt, no_tma_store, no_narrow_tma_store, wgrad_waves, no_gemv, no_tapgroup, no_fused_stats, sm_reserve, pdl, no_nsplit, no_wgrad_halo, wgrad_halo4, no_wgrad_whalo, wgrad_g, wgrad_ns; };
static Tuning g_tune = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
int set_tuning(const char* key, int value) {
  struct { const char* k; int* v; } tab[] = {{"nohalo", &g_tune.nohalo}, {"mt", &g_tune.mt}, {"no_tma_store", &g_tune.no_tma_store},
      {"no_narrow_tma_store", &g_tune.no_narrow_tma_store}, {"wgrad_waves", &g_tune.wgrad_waves}, {"no_gemv", &g_tune.no_gemv},
      {"no_tapgroup", &g_tune.no_tapgroup}, {"no_fused_stats", &g_tune.no_fused_stats}, {"sm_reserve", &g_tune.sm_reserve},
      {"pdl", &g_tune.pdl}, {"no_nsplit", &g_tune.no_nsplit}, {"no_wgrad_halo", &g_tune.no_wgrad_halo}, {"wgrad_halo4", &g_tune.wgrad_halo4}, {"no_wgrad_whalo", &g_tune.no_wgrad_whalo}, {"wgrad_g", &g_tune.wgrad_g}, {"wgrad_ns", &g_tune.wgrad_ns}};
  for (auto& t : tab) if (!strcmp(t.k, key)) { *t.v = value; return 0; }
  DCV_REQUIRE(false, "dcv_set_tuning: unknown key '%s'", key);
}
int tuning_sm_reserve() { return g_tune.sm_reserve; }
bool pdl_enabled() { return g_tune.pdl != 0; }

#ifdef DCV_EXPERIMENTS
static inline const char* exp_env(const char* name) { return getenv(name); }
#define TC_DBG(p) ((p).dbg)
#else
static inline const char* exp_env(const char*) { return nullptr; }
#define TC_DBG(p) 0
#endif

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// same box, but the destination is read-modify-written: dst += src (bf16 / fp32 add performed at L2, element type from the map)
__device__ __forceinline__ void tma_reduce_add_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- lean MMA issue ------------------------------------------------------------------------------
// Measured with tools/mma_probe.cu (profiles/r1h_mma_probe.md): a tcgen05.mma stream issued from inside `if (lane == 0)`
// with 64-bit descriptors rebuilt per instruction costs ~100-270 cycles of the issuing thread per MMA (R2UR moves and a
// compiler-generated "for each distinct lane value" loop around every UTCHMMA) - 25-40 % of the tensor pipe - while the
// same stream issued from warp-uniform code (every lane runs the loop, one elected lane issues, descriptors kept as
// 32-bit halves in uniform registers, K steps unrolled) retires back to back: 2160 TFLOP/s from one warp at N = 128.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// descriptor halves: lo = start>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version 1 (bit 46) | layout (bits 61..63)
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t sdesc_hi(uint32_t sbo_bytes, uint32_t layout) { return (sbo_bytes >> 4) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
// same with kind::tf32 (fp32 operands in shared memory, 10-bit mantissa products, fp32 accumulate; K = 8 per instruction = the
// same 32 bytes per K step, so descriptors and step arithmetic are shared with the bf16 path)
__device__ __forceinline__ void umma_lohi_tf32(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate) : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate) {
  if (TF32) umma_lohi_tf32(tmem_d, alo, ahi, blo, bhi, idesc, accumulate);
  else umma_lohi(tmem_d, alo, ahi, blo, bhi, idesc, accumulate);
}
// one pipeline stage of conv_tc: KS K-steps of 16 channels (32 bytes = +2 in descriptor units) for MT stacked M tiles
template <int KS>
__device__ __forceinline__ void conv_issue_stage(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                                 uint32_t accum, int mt, uint32_t a_tile16, uint32_t bnt) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    umma_lohi(tmem_d, alo + 2 * k, hi, blo + 2 * k, hi, idesc, k == 0 ? accum : 1u);
    if (mt == 2) umma_lohi(tmem_d + bnt, alo + a_tile16 + 2 * k, hi, blo + 2 * k, hi, idesc, k == 0 ? accum : 1u);
  }
}

// one pixel stage of wgrad_tc for one accumulator tile: KS steps of 16 pixels
// slabA16 / slab_shift: row-halo mode - the A box holds one extra tile row per (t, n) slab of slab = 1 << slab_shift pixels
template <int KS>
__device__ __forceinline__ void wgrad_issue(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                            uint32_t kstepA16, uint32_t kstepB16, uint32_t idesc, uint32_t accum,
                                            uint32_t slabA16, int slab_shift) {
#pragma unroll
  for (int k = 0; k < KS; ++k)
    umma_lohi(tmem_d, alo + k * kstepA16 + (uint32_t)((k * 16) >> slab_shift) * slabA16, ahi, blo + k * kstepB16, bhi, idesc,
              k == 0 ? accum : 1u);
}

template <int KS, int G>
__device__ __forceinline__ void wgrad_mma_loop(uint64_t* full_bar, uint64_t* empty_bar, int nst, int stages, uint32_t stage16,
                                               uint32_t lo_a0, uint32_t lo_b0, uint32_t ahi, uint32_t bhi, uint32_t kstepA16,
                                               uint32_t kstepB16, uint32_t pairA16, uint32_t oddA16, uint32_t tmem_base, uint32_t Ns,
                                               uint32_t idesc, bool leader, uint32_t slabA16, int slab_shift) {
  int stage = 0; uint32_t phase = 0, accum = 0, alo = lo_a0, blo = lo_b0;
  for (int it = 0; it < nst; ++it) {
    mbar_wait(&full_bar[stage], phase);
    tc_fence_after();
    if (leader) {
#pragma unroll
      for (int gi = 0; gi < G; ++gi)
        wgrad_issue<KS>(tmem_base + gi * Ns, alo + (gi >> 1) * pairA16 + (gi & 1) * oddA16, ahi, blo, bhi, kstepA16, kstepB16, idesc, accum,
                        slabA16, slab_shift);
      umma_commit(&empty_bar[stage]);
    }
    accum = 1;
    alo += stage16; blo += stage16;
    if (++stage == stages) { stage = 0; phase ^= 1u; alo = lo_a0; blo = lo_b0; }
  }
}

// shared-memory matrix descriptor (sm_100 "version 1"); see DESIGN.md for the field map
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle mode
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: D format f32 (bit 4), A / B format bf16 = 1 (kind::f16) or tf32 = 2 (kind::tf32) at bits 7 / 10
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int tf32 = 0) {
  const uint32_t fmt = tf32 ? 2u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int TC_THREADS = 192;
constexpr int MAX_STAGES = 8;
constexpr int TAPG = 4;     // grouped-tap mode: taps per pipeline stage

// ------------------------------------------------------------------------------------------ conv_tc
struct TcConvP {
  ConvP c;
  int bw, bh, bt, bn;                    // box of 128 output index positions
  int mt;                                // 128-row M tiles per CTA, stacked along the batch dimension (one TMA box)
  int tiles_w, tiles_h, tiles_t, tiles_n;
  int cblk, kchunks, bnt, stages;
  int swz_layout;                        // UMMA layout code (2 = SW128, 4 = SW64, 6 = SW32)
  int a_bytes, b_bytes, tx_bytes;
  int64_t ldy;
  int act; float slope;
  int accumulate;                        // epilogue adds into the output (TMA reduce-add) instead of overwriting it
  int vec_ok;
  int tmem_cols;
  int ntn, items, acc_cols;              // persistent kernel: N tiles, work items, TMEM columns of one accumulator set
  int tma_store;                         // persistent kernel: epilogue through shared memory + TMA store
  PhaseInfo phs[8];                      // persistent kernel: geometry of every sub-pixel phase (host-computed)
  float* stats;                          // fused BatchNorm statistics: [4 * grid][2][npad] fp32 partial sums (NULL = off)
  int npad;                              // padded output channels (stats row length)
  int hg, hcls;                          // row-halo sharing: taps per h class that share one A box (1 = off), h classes (nh / hg)
  int a_tile16;                          // M-tile stride inside the A box, in 16-byte units (rows incl. halo x row bytes)
  int dbg;                               // DCV_TC_DBG (timing experiments only): 1 = stop loading A, 2 = stop loading B after the first ring fill
};

// ------------------------------------------------------------------------------------------ conv_tc, persistent
// One CTA per SM walks the work items (tile = blockIdx.x + i * gridDim.x); the shared-memory ring runs continuously
// across items and TMEM holds two accumulator sets, so the store epilogue of item i overlaps the main loop of item i+1
// and TMEM allocation / barrier set-up happen once per SM instead of once per tile.  With the lean issue path a single
// MMA warp saturates the tensor pipe, which the first attempt at this kernel (profiles/experiments, r1d) could not.
struct TileCoord { int ph, ntile, w0, h0, t0, n0; };

// Work-item iterator of the persistent kernel.  item = ((((ph * ntn + ntile) * tiles_n + tn) * tiles_t + tt) * tiles_h + th)
// * tiles_w + tw; a CTA owns a CONTIGUOUS range of items, so stepping to the next one is an odometer increment instead of
// five integer divisions, neighbouring tiles (shared halo rows, same weight tile) follow each other in L2, and the
// phase geometry (PhaseInfo, precomputed on the host in TcConvP::phs) only changes a few times per CTA.  ncu of the
// first persistent version: ~35 % of all warp samples sat in the division sequences of decode_tile / make_phase, on
// the MMA warp that is tensor-pipe idle time.
struct TileIter { int tw, th, tt, tn, ntile, ph; };

__device__ __forceinline__ TileIter tile_iter_init(const TcConvP& p, int item) {
  TileIter t;
  int r = item;
  t.tw = r % p.tiles_w; r /= p.tiles_w;
  t.th = r % p.tiles_h; r /= p.tiles_h;
  t.tt = r % p.tiles_t; r /= p.tiles_t;
  t.tn = r % p.tiles_n; r /= p.tiles_n;
  t.ntile = r % p.ntn; t.ph = r / p.ntn;
  return t;
}
__device__ __forceinline__ void tile_iter_next(const TcConvP& p, TileIter& t) {
  if (++t.tw < p.tiles_w) return; t.tw = 0;
  if (++t.th < p.tiles_h) return; t.th = 0;
  if (++t.tt < p.tiles_t) return; t.tt = 0;
  if (++t.tn < p.tiles_n) return; t.tn = 0;
  if (++t.ntile < p.ntn) return; t.ntile = 0;
  ++t.ph;
}
__device__ __forceinline__ TileCoord tile_coord(const TcConvP& p, const TileIter& t) {
  TileCoord c;
  c.ph = t.ph; c.ntile = t.ntile; c.w0 = t.tw * p.bw; c.h0 = t.th * p.bh; c.t0 = t.tt * p.bt; c.n0 = t.tn * p.bn * p.mt;
  return c;
}
// contiguous item range of this CTA
__device__ __forceinline__ void cta_item_range(int items, int& first, int& count) {
  const int per = items / (int)gridDim.x, rem = items % (int)gridDim.x, b = (int)blockIdx.x;
  first = b * per + (b < rem ? b : rem);
  count = per + (b < rem ? 1 : 0);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 consecutive output channels of one row: activation, bf16, two 16-byte stores (or guarded scalar stores)
__device__ __forceinline__ void store_row16(const uint32_t* v, __nv_bfloat16* yrow, int c0, int Nc, int vec_ok, int act, float slope) {
  if (c0 >= Nc) return;
  if (vec_ok && c0 + 16 <= Nc) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = apply_act(__uint_as_float(v[2 * i]), act, slope);
      const float b = apply_act(__uint_as_float(v[2 * i + 1]), act, slope);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
      pk[i] = *reinterpret_cast<uint32_t*>(&h2);
    }
    uint4* dst = reinterpret_cast<uint4*>(yrow + c0);
    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < Nc) yrow[c0 + i] = __float2bfloat16_rn(apply_act(__uint_as_float(v[i]), act, slope));
  }
}

// number of taps of a tile that touch real pixels (the skip test is separable per dimension); >= 1 by convention
// (a tile whose every tap lies in the padding still runs its last tap on an all-zero box so that TMEM is written)
__device__ __forceinline__ int live_taps(const TcConvP& p, const PhaseInfo& f, const TileCoord& tc) {
  int lt = 0, lh = 0, lw = 0;
  for (int jt = 0; jt < f.nt; ++jt) { const int c = tc.t0 * f.mult + f.offt + f.sgn * jt; lt += !((c + (p.bt - 1) * f.mult < 0) || (c >= p.c.It)); }
  for (int jh = 0; jh < f.nh; ++jh) { const int c = tc.h0 * f.mulh + f.offh + f.sgn * jh; lh += !((c + (p.bh - 1) * f.mulh < 0) || (c >= p.c.Ih)); }
  for (int jw = 0; jw < f.nw; ++jw) { const int c = tc.w0 * f.mulw + f.offw + f.sgn * jw; lw += !((c + (p.bw - 1) * f.mulw < 0) || (c >= p.c.Iw)); }
  const int n = lt * lh * lw;
  return n > 0 ? n : 1;
}

// same with row-halo sharing: an h "tap" is a class of HG taps whose common box (bh + HG - 1 rows) has to touch real pixels
template <int HG>
__device__ __forceinline__ int live_stages(const TcConvP& p, const PhaseInfo& f, const TileCoord& tc) {
  if (HG == 1) return live_taps(p, f, tc);
  int lt = 0, lh = 0, lw = 0;
  for (int jt = 0; jt < f.nt; ++jt) { const int c = tc.t0 * f.mult + f.offt + f.sgn * jt; lt += !((c + (p.bt - 1) * f.mult < 0) || (c >= p.c.It)); }
  for (int jh = 0; jh < p.hcls; ++jh) {
    int c = tc.h0 * f.mulh + f.offh + f.sgn * jh;
    if (f.sgn < 0) c -= (HG - 1) * f.mulh;
    lh += !((c + (p.bh + HG - 2) * f.mulh < 0) || (c >= p.c.Ih));
  }
  for (int jw = 0; jw < f.nw; ++jw) { const int c = tc.w0 * f.mulw + f.offw + f.sgn * jw; lw += !((c + (p.bw - 1) * f.mulw < 0) || (c >= p.c.Iw)); }
  const int n = lt * lh * lw;
  return n > 0 ? n : 1;
}

// KS = K steps of 16 channels per stage (cblk / 16), MT = 128-row M tiles per work item
// HG = taps of one h class that share a single A box with HG-1 halo rows (1 = every tap loads its own box)
// TF32: fp32 activations and weights, kind::tf32 MMAs, fp32 output through the TMA-store epilogue (32-channel chunks)
template <int KS, int MT, int HG, bool TF32 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_pers_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapY, const TcConvP p, __nv_bfloat16* __restrict__ y) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t stage_info[MAX_STAGES];   // grouped-tap mode: bits 0..3 = taps of the stage that were loaded, bit 8 = last stage of the item

  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x / 32, 0), lane = threadIdx.x % 32;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = p.a_bytes + HG * p.b_bytes;
  const int items = p.items;

  if (warp == 0 && lane == 0) { tmap_prefetch(&mapA); tmap_prefetch(&mapB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // global memory is only touched from here on
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    const bool leader = elect_one();
    int stage = 0; uint32_t phase = 0; int issued = 0;
    int item_first, item_count;
    cta_item_range(items, item_first, item_count);
    TileIter ti = tile_iter_init(p, item_first);
    int cur_ph = ti.ph;
    PhaseInfo f = p.phs[cur_ph];
    for (int itn = 0; itn < item_count; ++itn, tile_iter_next(p, ti)) {
      if (ti.ph != cur_ph) { cur_ph = ti.ph; f = p.phs[cur_ph]; }
      const TileCoord tc = tile_coord(p, ti);
      if (tc.w0 >= f.Qw || tc.h0 >= f.Qh || tc.t0 >= f.Qt) continue;        // tile outside this phase
      const int ntaps = f.nt * f.nh * f.nw;
      const int bcol = tc.ntile * p.bnt;
      if constexpr (KS == 0) {
        // Grouped taps (16-channel, image-like operands): a stage carries FOUR taps - four A boxes of 32-byte rows issued
        // by four lanes and one B box of 4 taps x 16 channels (128-byte rows).  Which taps were loaded travels to the MMA
        // warp in stage_info (written before the releasing expect_tx arrive).
        auto tap_coords = [&](int j, int& ct, int& ch, int& cw) -> bool {
          const int jw = j % f.nw, jh = (j / f.nw) % f.nh, jt = j / (f.nw * f.nh);
          ct = tc.t0 * f.mult + f.offt + f.sgn * jt;
          ch = tc.h0 * f.mulh + f.offh + f.sgn * jh;
          cw = tc.w0 * f.mulw + f.offw + f.sgn * jw;
          const bool skt = (ct + (p.bt - 1) * f.mult < 0) || (ct >= p.c.It);
          const bool skh = (ch + (p.bh - 1) * f.mulh < 0) || (ch >= p.c.Ih);
          const bool skw = (cw + (p.bw - 1) * f.mulw < 0) || (cw >= p.c.Iw);
          return !(skt || skh || skw);
        };
        int ct = 0, ch = 0, cw = 0;
        uint32_t lo = __ballot_sync(0xffffffffu, lane < ntaps && tap_coords(lane, ct, ch, cw));
        uint32_t hi = __ballot_sync(0xffffffffu, lane + 32 < ntaps && tap_coords(lane + 32, ct, ch, cw));
        if ((lo | hi) == 0u) { if (ntaps > 32) hi = 1u << (ntaps - 33); else lo = 1u << (ntaps - 1); }   // all padding: run the last tap on zeros
        const uint64_t live = ((uint64_t)hi << 32) | lo;
        const int glast = (63 - __clzll((long long)live)) / TAPG;
        const int a_tap_bytes = p.a_bytes / TAPG;
        for (int gq = 0; gq <= glast; ++gq) {
          const uint32_t mask = (uint32_t)(live >> (gq * TAPG)) & 0xFu;
          if (mask == 0u) continue;
          const bool mine = lane < TAPG && ((mask >> lane) & 1u);
          if (mine) tap_coords(gq * TAPG + lane, ct, ch, cw);
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t a_dst = sbase + stage * stage_bytes;
          if (lane == 0) {
            stage_info[stage] = mask | (gq == glast ? 0x100u : 0u);
            mbar_expect_tx(&full_bar[stage], (uint32_t)(__popc(mask) * a_tap_bytes + p.bnt * 128));
          }
          __syncwarp();
          if (mine) tma_load_5d(a_dst + lane * a_tap_bytes, &mapA, &full_bar[stage], 0, cw, ch, ct, tc.n0);
          if (lane == TAPG) tma_load_3d(a_dst + p.a_bytes, &mapB, &full_bar[stage], gq * TAPG * 16, bcol, tc.ph);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      int executed = 0;
      const int nhc = HG > 1 ? p.hcls : f.nh;                 // h classes (HG taps each) or plain h taps
      const int nlast = f.nt * nhc * f.nw - 1;
      int jc = 0;
      for (int jt = 0; jt < f.nt; ++jt) {
        const int ct = tc.t0 * f.mult + f.offt + f.sgn * jt;
        const bool skt = (ct + (p.bt - 1) * f.mult < 0) || (ct >= p.c.It);
        for (int jh = 0; jh < nhc; ++jh) {
          // first lattice row of the box: tap jh itself, or with halo sharing the lowest row of the class' HG taps
          // (tap i of class jh reads rows shifted by +i (gather) / -i (scatter) in units of the traversal stride)
          int ch = tc.h0 * f.mulh + f.offh + f.sgn * jh;
          if (HG > 1 && f.sgn < 0) ch -= (HG - 1) * f.mulh;
          const bool skh = (ch + (p.bh + HG - 2) * f.mulh < 0) || (ch >= p.c.Ih);
          for (int jw = 0; jw < f.nw; ++jw, ++jc) {
            const int cw = tc.w0 * f.mulw + f.offw + f.sgn * jw;
            const bool skw = (cw + (p.bw - 1) * f.mulw < 0) || (cw >= p.c.Iw);
            if ((skt || skh || skw) && !(jc == nlast && executed == 0)) continue;   // box entirely in the padding
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              const bool ldA = !(TC_DBG(p) & 1) || issued < p.stages, ldB = !(TC_DBG(p) & 2) || issued < p.stages;
              ++issued;
              if (leader) {
                mbar_expect_tx(&full_bar[stage], (uint32_t)((ldA ? p.a_bytes : 0) + (ldB ? p.tx_bytes - p.a_bytes : 0)));
                const uint32_t a_dst = sbase + stage * stage_bytes;
                if (ldA) tma_load_5d(a_dst, &mapA, &full_bar[stage], kc * p.cblk, cw, ch, ct, tc.n0);
                if (ldB) {
#pragma unroll
                  for (int i = 0; i < HG; ++i) {
                    const int jhh = HG > 1 ? jh + i * p.hcls : jh;             // h tap i of this class
                    const int j = (jt * f.nh + jhh) * f.nw + jw;
                    tma_load_3d(a_dst + p.a_bytes + i * p.b_bytes, &mapB, &full_bar[stage], j * p.c.Kc + kc * p.cblk, bcol, tc.ph);
                  }
                }
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            ++executed;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // MMA issuer.  The geometry is reduced to "how many stages does this item have" before its flat issue loop:
    // a single warp executes the loop control serially (~5 cycles per dependent instruction), so everything that is
    // not wait / issue / commit has to stay out of it (profiles/r1h_mma_probe.md: ~250 cycles of control per stage
    // already cost 30 % at 4 MMAs per stage).
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, p.bnt, 0, 0, TF32 ? 1 : 0);
    const uint32_t dhi = sdesc_hi(8u * (uint32_t)((KS ? KS : 1) * 16) * 2u, (uint32_t)p.swz_layout);
    const uint32_t a_tile16 = HG > 1 ? (uint32_t)p.a_tile16 : ((128u * (uint32_t)((KS ? KS : 1) * 16) * 2u) >> 4);
    const uint32_t row16 = ((uint32_t)p.bw * (uint32_t)((KS ? KS : 1) * 16) * 2u) >> 4;   // one tile row (bw pixels) in 16-byte units
    const uint32_t bt16 = (uint32_t)p.b_bytes >> 4;
    const bool sgn_pos = p.phs[0].sgn > 0;
    const uint32_t g4_bhi = sdesc_hi(1024, 2);              // grouped taps: B rows are 128 bytes (4 taps x 16 channels), 128B swizzle
    const uint32_t g4_tap16 = (uint32_t)(p.a_bytes / TAPG) >> 4;
    const uint32_t stage16 = (uint32_t)stage_bytes >> 4, b_off16 = (uint32_t)p.a_bytes >> 4;
    const uint32_t lo0 = sdesc_lo(sbase, 16), lo_end = lo0 + (uint32_t)p.stages * stage16;
    const uint32_t bnt = (uint32_t)p.bnt;
    uint32_t alo = lo0; int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t tempty_phase0 = 0u, tempty_phase1 = 0u;
    int item_first, item_count;
    cta_item_range(items, item_first, item_count);
    TileIter ti = tile_iter_init(p, item_first);
    int cur_ph = ti.ph;
    PhaseInfo f = p.phs[cur_ph];
    for (int itn = 0; itn < item_count; ++itn, tile_iter_next(p, ti)) {
      if (ti.ph != cur_ph) { cur_ph = ti.ph; f = p.phs[cur_ph]; }
      const TileCoord tc = tile_coord(p, ti);
      if (tc.w0 >= f.Qw || tc.h0 >= f.Qh || tc.t0 >= f.Qt) continue;
      const int nst = KS == 0 ? 0 : live_stages<HG>(p, f, tc) * p.kchunks;
      mbar_wait(&tempty_bar[acc], (acc ? tempty_phase1 : tempty_phase0) ^ 1u);      // epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(acc * p.acc_cols);
      uint32_t accum = 0;
      if constexpr (KS == 0) {
        uint32_t info;
        do {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          info = __shfl_sync(0xffffffffu, *((volatile uint32_t*)&stage_info[stage]), 0);
          if (leader) {
            const uint32_t blo = alo + b_off16;
#pragma unroll
            for (int jj = 0; jj < TAPG; ++jj) {
              if (!((info >> jj) & 1u)) continue;
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_lohi(d_base + m * bnt, alo + jj * g4_tap16 + m * ((128u * 32u) >> 4), dhi, blo + 2 * jj, g4_bhi, idesc, accum);
              accum = 1;
            }
            umma_commit(&empty_bar[stage]);
          }
          accum = 1;
          alo += stage16;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; alo = lo0; }
        } while (!(info & 0x100u));
      }
      for (int it = 0; it < nst; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (leader) {
          const uint32_t blo = alo + b_off16;
#pragma unroll
          for (int i = 0; i < HG; ++i) {
            // tap i of the class: A window shifted by whole tile rows inside the halo box (a multiple of the swizzle
            // atom, so the descriptor only changes its start address), B tile i of the stage
            const uint32_t ai = alo + (HG > 1 ? (sgn_pos ? (uint32_t)i : (uint32_t)(HG - 1 - i)) * row16 : 0u);
            const uint32_t bi = blo + (uint32_t)i * bt16;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_issue<TF32>(d_base + m * bnt, ai + m * a_tile16 + 2 * k, dhi, bi + 2 * k, dhi, idesc, (i == 0 && k == 0) ? accum : 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        accum = 1;
        alo += stage16;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; alo = lo0; }
      }
      (void)lo_end;
      if (leader) umma_commit(&tfull_bar[acc]);
      if (acc) tempty_phase1 ^= 1u; else tempty_phase0 ^= 1u;
      acc ^= 1;
    }
    __syncwarp();
  } else {
    // epilogue: TMEM lane quarter = warp % 4 (hardware restriction), row = quarter*32 + lane
    const int quarter = warp % 4;
    const int row = quarter * 32 + lane;
    int r = row;
    const int dw = r % p.bw; r /= p.bw;
    const int dh = r % p.bh; r /= p.bh;
    const int dt = r % p.bt; const int dn = r / p.bt;
    int acc = 0; uint32_t tfull_phase0 = 0u, tfull_phase1 = 0u;
    const uint32_t stg_base = sbase + (uint32_t)(p.stages * stage_bytes);      // 4 warps x 2 x 4 KB staging buffers (TMA-store path)
    int nstore = 0;
    // fused BatchNorm batch statistics (sum, sum of squares per output channel over the bf16-rounded outputs): every warp
    // owns one [2][npad] slot of p.stats and accumulates the channel pair (2*lane, 2*lane+1) of up to four 64-channel chunks
    float* const st_slot = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + quarter) * 2 * p.npad : nullptr;
    float bs[4][2], bq[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bs[i][0] = bs[i][1] = bq[i][0] = bq[i][1] = 0.f; }
    int st_ntile = -1;
    if (st_slot) {
      for (int i = lane; i < 2 * p.npad; i += 32) st_slot[i] = 0.f;
      __syncwarp();
    }
    auto st_flush = [&](int ntile) {
      if (!st_slot || ntile < 0) return;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = ntile * p.bnt + i * 64 + 2 * lane;
        if (i * 64 < p.bnt) {
          st_slot[c] += bs[i][0]; st_slot[c + 1] += bs[i][1];
          st_slot[p.npad + c] += bq[i][0]; st_slot[p.npad + c + 1] += bq[i][1];
        }
        bs[i][0] = bs[i][1] = bq[i][0] = bq[i][1] = 0.f;
      }
    };
    int sub_w, sub_h, sub_t, sub_n;                                            // tile-local origin of this warp's 32 rows
    {
      int r0 = quarter * 32;
      sub_w = r0 % p.bw; r0 /= p.bw;
      sub_h = r0 % p.bh; r0 /= p.bh;
      sub_t = r0 % p.bt; sub_n = r0 / p.bt;
    }
    int item_first, item_count;
    cta_item_range(items, item_first, item_count);
    TileIter ti = tile_iter_init(p, item_first);
    int cur_ph = ti.ph;
    PhaseInfo f = p.phs[cur_ph];
    for (int itn = 0; itn < item_count; ++itn, tile_iter_next(p, ti)) {
      if (ti.ph != cur_ph) { cur_ph = ti.ph; f = p.phs[cur_ph]; }
      const TileCoord tc = tile_coord(p, ti);
      if (tc.w0 >= f.Qw || tc.h0 >= f.Qh || tc.t0 >= f.Qt) continue;
      const int qw = tc.w0 + dw, qh = tc.h0 + dh, qt = tc.t0 + dt;
      const int nbase = tc.ntile * p.bnt;
      mbar_wait(&tfull_bar[acc], acc ? tfull_phase1 : tfull_phase0);
      tc_fence_after();
      const uint32_t d_base = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.acc_cols);
      if constexpr (TF32) {
        // fp32 output: 32-channel chunks = 128-byte rows (128B swizzle), same per-warp staging + TMA store as the bf16 path
        const uint32_t swz_row = (uint32_t)lane * 128u, swz_x = (uint32_t)(lane & 7);
        const uint32_t wbuf = stg_base + (uint32_t)quarter * 8192u;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          for (int cb = 0; cb < p.bnt; cb += 32) {
            const uint32_t buf = wbuf + (uint32_t)(nstore & 1) * 4096u;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            uint32_t v[32];
            tmem_ld32_nowait(d_base + (uint32_t)(m * p.bnt + cb), v);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = __float_as_uint(apply_act(__uint_as_float(v[4 * c + i]), p.act, p.slope));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + swz_row + (((uint32_t)c ^ swz_x) << 4)),
                           "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.accumulate) tma_reduce_add_5d(&mapY, buf, nbase + cb, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                                                  (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
              else tma_store_5d(&mapY, buf, nbase + cb, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                                (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
              tma_store_commit();
            }
            ++nstore;
          }
        }
      } else if (p.tma_store == 2) {
        // narrow outputs (bnt = 16 or 32 channels: image-like tensors, tap-unrolled partial products): same staging + TMA
        // store with 32- / 64-byte rows (32B / 64B swizzle), one TMEM load per row chunk
        const int cw = p.bnt;                                   // 16 or 32
        const uint32_t rb = (uint32_t)cw * 2u;
        const uint32_t swz_x = cw == 32 ? (uint32_t)((lane >> 1) & 3) : (uint32_t)((lane >> 2) & 1);
        const uint32_t wbuf = stg_base + (uint32_t)quarter * 8192u;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint32_t buf = wbuf + (uint32_t)(nstore & 1) * 4096u;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint32_t v[32];
          if (cw == 32) { tmem_ld32_nowait(d_base + (uint32_t)(m * p.bnt), v); tmem_ld_wait(); }
          else { uint32_t t16[16]; tmem_ld16(d_base + (uint32_t)(m * p.bnt), t16);
#pragma unroll
                 for (int i = 0; i < 16; ++i) v[i] = t16[i]; }
          if (!(TC_DBG(p) & 4)) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c * 8 >= cw) break;
              uint32_t pk[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float a = apply_act(__uint_as_float(v[8 * c + 2 * i]), p.act, p.slope);
                const float b = apply_act(__uint_as_float(v[8 * c + 2 * i + 1]), p.act, p.slope);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * rb + (((uint32_t)c ^ swz_x) << 4)),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && !(TC_DBG(p) & 4)) {
            if (p.accumulate) tma_reduce_add_5d(&mapY, buf, nbase, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                                                (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
            else tma_store_5d(&mapY, buf, nbase, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                              (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
            tma_store_commit();
          }
          ++nstore;
        }
      } else if (p.tma_store) {
        // Coalesced path (bnt % 64 == 0): every epilogue warp stages ITS 32 rows of a 64-channel chunk in shared memory
        // (128-byte rows, 128B swizzle -> conflict-free 16-byte writes) and writes them with its own TMA store; the tensor
        // map does the sub-pixel scatter and clips rows / channels outside the tensor.  No cross-warp barrier: the four
        // warps drain TMEM independently.  (Direct 32-byte row stores from the TMEM registers ran at ~1.6 TB/s and cost
        // half of the kernel time on the 64x64-pixel layers; one store per 128-row tile with two block barriers per
        // chunk still limited the 1x1-like layers to 2.8 TB/s of output.)
        const uint32_t swz_row = (uint32_t)lane * 128u, swz_x = (uint32_t)(lane & 7);
        const uint32_t wbuf = stg_base + (uint32_t)quarter * 8192u;
        if (st_slot && tc.ntile != st_ntile) { st_flush(st_ntile); st_ntile = tc.ntile; }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          // rows of this warp that are real output positions (TMA clips the others on store; the statistics must too)
          const uint32_t rowmask = st_slot ? __ballot_sync(0xffffffffu, qw < f.Qw && qh < f.Qh && qt < f.Qt && tc.n0 + m * p.bn + dn < p.c.N) : 0u;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const int cb = ci * 64;
            if (cb >= p.bnt) break;
            const uint32_t buf = wbuf + (uint32_t)(nstore & 1) * 4096u;
            if (lane == 0) tma_store_wait_read<1>();                   // this warp's store that last used the buffer has read it
            __syncwarp();
            uint32_t v[32], w[32];
            tmem_ld32_nowait(d_base + (uint32_t)(m * p.bnt + cb), v);
            tmem_ld32_nowait(d_base + (uint32_t)(m * p.bnt + cb + 32), w);
            tmem_ld_wait();
            if (!(TC_DBG(p) & 4)) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t* src = c < 4 ? v + 8 * c : w + 8 * (c - 4);
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float a = apply_act(__uint_as_float(src[2 * i]), p.act, p.slope);
                  const float b = apply_act(__uint_as_float(src[2 * i + 1]), p.act, p.slope);
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                  pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + swz_row + (((uint32_t)c ^ swz_x) << 4)),
                             "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
              }
            }
            fence_async_smem();
            __syncwarp();
            if (st_slot) {
              // column sums over this warp's 32 staged rows: lane l reads the bf16 pair (2l, 2l+1) of every row - 32
              // distinct words per row (the swizzle only permutes 16-byte chunks inside the row), so no bank conflicts
              const uint32_t cx = (uint32_t)(lane >> 2), co = (uint32_t)(lane & 3) * 4u;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                if (!((rowmask >> r) & 1u)) continue;
                uint32_t wv;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wv) : "r"(buf + (uint32_t)r * 128u + ((cx ^ (uint32_t)(r & 7)) << 4) + co));
                const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xFFFF0000u);
                bs[ci][0] += f0; bs[ci][1] += f1;
                bq[ci][0] = fmaf(f0, f0, bq[ci][0]); bq[ci][1] = fmaf(f1, f1, bq[ci][1]);
              }
            }
            if (lane == 0 && !(TC_DBG(p) & 4)) {
              if (p.accumulate) tma_reduce_add_5d(&mapY, buf, nbase + cb, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                                                  (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
              else tma_store_5d(&mapY, buf, nbase + cb, (tc.w0 + sub_w) * f.osw + f.rw, (tc.h0 + sub_h) * f.osh + f.rh,
                                (tc.t0 + sub_t) * f.ost + f.rt, tc.n0 + m * p.bn + sub_n);
              tma_store_commit();
            }
            ++nstore;
          }
        }
      } else
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int n = tc.n0 + m * p.bn + dn;
        const bool valid = qw < f.Qw && qh < f.Qh && qt < f.Qt && n < p.c.N;
        const int64_t pos = (((int64_t)n * p.c.Ot + (qt * f.ost + f.rt)) * p.c.Oh + (qh * f.osh + f.rh)) * p.c.Ow + (qw * f.osw + f.rw);
        __nv_bfloat16* yrow = y + pos * p.ldy;
        int cb = 0;
        if (TC_DBG(p) & 8) continue;                                // timing experiment: no TMEM reads, no stores
        for (; cb + 32 <= p.bnt; cb += 32) {                    // 32 columns per TMEM round trip
          uint32_t v[32];
          tmem_ld32_nowait(d_base + (uint32_t)(m * p.bnt + cb), v);
          tmem_ld_wait();
          if (!valid || (TC_DBG(p) & 4)) continue;
          store_row16(v, yrow, nbase + cb, p.c.Nc, p.vec_ok, p.act, p.slope);
          store_row16(v + 16, yrow, nbase + cb + 16, p.c.Nc, p.vec_ok, p.act, p.slope);
        }
        for (; cb < p.bnt; cb += 16) {
          uint32_t v[16];
          tmem_ld16(d_base + (uint32_t)(m * p.bnt + cb), v);
          if (!valid || (TC_DBG(p) & 4)) continue;
          store_row16(v, yrow, nbase + cb, p.c.Nc, p.vec_ok, p.act, p.slope);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);                            // this warp is done reading the set
      if (acc) tfull_phase1 ^= 1u; else tfull_phase0 ^= 1u;
      acc ^= 1;
    }
    st_flush(st_ntile);
    if (p.tma_store && lane == 0) tma_store_wait_all();                         // every warp's bulk stores must complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

typedef void (*ConvPersFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcConvP, __nv_bfloat16*);
static ConvPersFn conv_pers_variant_tf32(int mt, int hg) {      // 32 fp32 channels per stage = 128-byte rows = 4 K steps
#define DCV_V(MT_, HG_) if (mt == MT_ && hg == HG_) return conv_tc_pers_kernel<4, MT_, HG_, true>;
  DCV_V(1, 1) DCV_V(2, 1) DCV_V(4, 1) DCV_V(1, 2) DCV_V(2, 2) DCV_V(4, 2) DCV_V(1, 3) DCV_V(2, 3) DCV_V(4, 3)
#undef DCV_V
  return nullptr;
}
static ConvPersFn conv_pers_variant(int ks, int mt, int hg) {
#define DCV_V(KS_, MT_, HG_) if (ks == KS_ && mt == MT_ && hg == HG_) return conv_tc_pers_kernel<KS_, MT_, HG_>;
  DCV_V(4, 1, 1) DCV_V(4, 2, 1) DCV_V(4, 4, 1) DCV_V(2, 1, 1) DCV_V(2, 2, 1) DCV_V(2, 4, 1) DCV_V(1, 1, 1) DCV_V(1, 2, 1) DCV_V(1, 4, 1)
  DCV_V(0, 1, 1) DCV_V(0, 2, 1) DCV_V(0, 4, 1)                                      // grouped taps (16-channel operands)
  DCV_V(4, 1, 2) DCV_V(4, 2, 2) DCV_V(4, 4, 2) DCV_V(2, 1, 2) DCV_V(2, 2, 2) DCV_V(2, 4, 2)   // row-halo sharing, 2 taps per class
  DCV_V(4, 1, 3) DCV_V(4, 2, 3) DCV_V(4, 4, 3) DCV_V(2, 1, 3) DCV_V(2, 2, 3) DCV_V(2, 4, 3)   // 3 taps per class (3x3 stride 1)
#undef DCV_V
  return nullptr;
}

// ------------------------------------------------------------------------------------------ single-channel heads
// The discriminators end in a convolution with ONE output channel (discriminator.py:101,206,305): a GEMV, HBM-bound
// (every input element is used once per overlapping window).  As a 128 x 16 MMA tile it ran 16 CTAs with 256 serial
// K iterations (175 us for 0.07 GFLOP); here one warp owns one logit, lanes stride over the channels with 16-byte
// loads and the bf16 weight row (row 0 of the tensor-core packing [n][tap][k]) is read through the same index.
__global__ void __launch_bounds__(256)
head_gemv_kernel(ConvP c, const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                 __nv_bfloat16* __restrict__ y, int64_t ldy, int act, float slope) {
  pdl_wait(); pdl_trigger();
  const int lane = threadIdx.x % 32;
  const int64_t M = (int64_t)c.N * c.Ot * c.Oh * c.Ow;
  int64_t m = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (m >= M) return;
  const int64_t pos = m;
  const int ow = (int)(m % c.Ow); m /= c.Ow;
  const int oh = (int)(m % c.Oh); m /= c.Oh;
  const int ot = (int)(m % c.Ot); const int n = (int)(m / c.Ot);
  float acc = 0.f;
  for (int a = 0; a < c.kt; ++a) {
    const int it = ot * c.st - c.pt + a;
    if (it < 0 || it >= c.It) continue;
    for (int b = 0; b < c.kh; ++b) {
      const int ih = oh * c.sh - c.ph + b;
      if (ih < 0 || ih >= c.Ih) continue;
      for (int d = 0; d < c.kw; ++d) {
        const int iw = ow * c.sw - c.pw + d;
        if (iw < 0 || iw >= c.Iw) continue;
        const __nv_bfloat16* xp = x + ((((int64_t)n * c.It + it) * c.Ih + ih) * c.Iw + iw) * ldx;
        const __nv_bfloat16* wt = w + (int64_t)((a * c.kh + b) * c.kw + d) * c.Kc;
        for (int k = lane * 8; k < c.Kc; k += 256) {
          const uint4 xv = *reinterpret_cast<const uint4*>(xp + k);
          const uint4 wv = *reinterpret_cast<const uint4*>(wt + k);
          const uint32_t xa[4] = {xv.x, xv.y, xv.z, xv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc = fmaf(__uint_as_float(xa[i] << 16), __uint_as_float(wa[i] << 16), acc);
            acc = fmaf(__uint_as_float(xa[i] & 0xFFFF0000u), __uint_as_float(wa[i] & 0xFFFF0000u), acc);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[pos * ldy] = __float2bfloat16_rn(apply_act(acc, act, slope));
}

// The same GEMV with the adversarial-loss term of the head fused in (loss.py:93-99,123-131,163-166,190-193 on the logits of
// discriminator.py:101,206,305): every block writes its logits, takes a ticket, and the block that arrives last computes
// the mean loss over all M logits (in the summation order of loss_kernel, from the bf16-rounded logits it re-reads through
// L2: identical numbers to the two-launch form) and dL/dlogit for the backward pass.  kind: DCV_LOSS_*.
__device__ __forceinline__ float hl_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }   // == misc.cu softplusf_
__device__ __forceinline__ float hl_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
head_gemv_loss_kernel(ConvP c, const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ w,
                      __nv_bfloat16* __restrict__ y, int64_t ldy, int kind, float* __restrict__ loss_out, int accumulate,
                      __nv_bfloat16* __restrict__ dy, int64_t lddy, float grad_scale, unsigned* __restrict__ counter) {
  pdl_wait(); pdl_trigger();
  __shared__ float red[256];
  __shared__ int s_last;
  const int lane = threadIdx.x % 32;
  const int64_t M = (int64_t)c.N * c.Ot * c.Oh * c.Ow;
  int64_t m = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (m < M) {
    const int64_t pos = m;
    const int ow = (int)(m % c.Ow); m /= c.Ow;
    const int oh = (int)(m % c.Oh); m /= c.Oh;
    const int ot = (int)(m % c.Ot); const int n = (int)(m / c.Ot);
    float acc = 0.f;
    for (int a = 0; a < c.kt; ++a) {
      const int it = ot * c.st - c.pt + a;
      if (it < 0 || it >= c.It) continue;
      for (int b = 0; b < c.kh; ++b) {
        const int ih = oh * c.sh - c.ph + b;
        if (ih < 0 || ih >= c.Ih) continue;
        for (int d = 0; d < c.kw; ++d) {
          const int iw = ow * c.sw - c.pw + d;
          if (iw < 0 || iw >= c.Iw) continue;
          const __nv_bfloat16* xp = x + ((((int64_t)n * c.It + it) * c.Ih + ih) * c.Iw + iw) * ldx;
          const __nv_bfloat16* wt = w + (int64_t)((a * c.kh + b) * c.kw + d) * c.Kc;
          for (int k = lane * 8; k < c.Kc; k += 256) {
            const uint4 xv = *reinterpret_cast<const uint4*>(xp + k);
            const uint4 wv = *reinterpret_cast<const uint4*>(wt + k);
            const uint32_t xa[4] = {xv.x, xv.y, xv.z, xv.w}, wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc = fmaf(__uint_as_float(xa[i] << 16), __uint_as_float(wa[i] << 16), acc);
              acc = fmaf(__uint_as_float(xa[i] & 0xFFFF0000u), __uint_as_float(wa[i] & 0xFFFF0000u), acc);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[pos * ldy] = __float2bfloat16_rn(acc);
  }
  // ---- last block: the loss term and its gradient over all logits
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float lacc = 0.f;
  const float inv_n = 1.f / (float)M;
  for (int64_t i = threadIdx.x; i < M; i += 256) {
    const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(y + i * ldy));
    const float v = __uint_as_float((uint32_t)raw << 16);
    float l, g;
    switch (kind) {
      case DCV_LOSS_BCE_ONES:
      case DCV_LOSS_SOFTPLUS_NEG: l = hl_softplus(-v); g = hl_sigmoid(v) - 1.f; break;
      case DCV_LOSS_BCE_ZEROS: l = hl_softplus(v); g = hl_sigmoid(v); break;
      case DCV_LOSS_HINGE_REAL: l = fmaxf(1.f - v, 0.f); g = (1.f - v > 0.f) ? -1.f : 0.f; break;
      default: l = fmaxf(1.f + v, 0.f); g = (1.f + v > 0.f) ? 1.f : 0.f; break;
    }
    lacc += l;
    if (dy) dy[i * lddy] = __float2bfloat16_rn(g * inv_n * grad_scale);
  }
  red[threadIdx.x] = lacc;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = red[0] * inv_n;
    loss_out[0] = accumulate ? loss_out[0] + l : l;
    *counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------ wgrad_tc
// D[128 rows = (128/cbA) gathered L blocks of cbA channels, Ns columns = S channels] += A^T B over pixel blocks.
// A block = (tap, channel chunk of cbA) of the gathered L tensor, B = S tile; both MN-major (channels contiguous),
// swizzle chosen by the block width (64 ch -> 128B, 32 -> 64B, 16 -> 32B).  G accumulators of Ns columns share one
// S tile per stage.
struct TcWgradP {
  dcv_geom g;
  int pix;                               // pixels (GEMM K) per stage
  int bw, bh, bt, bn;
  int tiles_w, tiles_h, tiles_t, tiles_n;
  int64_t ptiles_total, ptiles_per_split;
  int cbA, nA, clchunks, blocksA_total;  // A: block width, blocks per 128-row tile, channel chunks per tap, taps*clchunks
  int cbB, nbB, blocksB_total;           // B: block width, blocks per CTA tile, total blocks
  int G, Ns, stages, tmem_cols;
  int layA, layB;                        // UMMA layout codes
  int dbg;                               // DCV_TC_DBG & 3: stop issuing TMA loads after the first ring fill (timing experiments)
  // Row-halo sharing (k = 4, stride 2 along h, 64-channel blocks): the taps kh and kh + 2 read the same stride-2 row lattice of
  // L, one tile row apart.  A 128-row accumulator tile is then the PAIR (kh, kh + 2) of one (kt, kh % 2, kw, channel chunk),
  // both halves served by ONE box with a halo row - the second half's descriptor starts bw pixels (a multiple of the 1024-byte
  // swizzle atom) further down - so a stage carries (bh + 1) / (2 bh) of the per-tap boxes.
  // Mode 2 (k = 4, stride 2 along w as well, 8-pixel-wide tiles) shares along w too: the FOUR taps (kh % 2 + {0, 2},
  // kw % 2 + {0, 2}) of a parity class read one box of (bh + 1) x (bw + 1) lattice pixels; the kw + 2 pair is the same box one
  // PIXEL (128 bytes) further - the 128-byte swizzle is a function of the absolute shared-memory address, so a descriptor that
  // starts inside a swizzle atom still reads what TMA wrote (checked on the GPU with the half-atom row shift of 4-pixel tiles).
  // Two accumulator tiles per box: a stage carries (bh + 1)(bw + 1) / (4 bh bw) of the per-tap boxes.
  int halo;                              // 0: one box per (tap, chunk); 1: row halo; 2: row + column halo
  int boxA_bytes;                        // shared-memory pitch of one halo box: slabs * (bh + 1) * rpA pixels * 128, rounded up to 1024
  int boxA_tx;                           // bytes TMA delivers per box (the same without the rounding)
  int rpA;                               // pixels per box row: bw (mode 1) or bw + 1 (mode 2)
  int slab_shift;                        // log2(bw * bh): pixels of one (t, n) slab of the tile
};

// halo modes: tile index -> (kt, kh class, kw, channel chunk); the tile's row block d (0 / 1) is the tap kh = class + 2 d.
// Mode 2 orders the tiles (kt, kh class, kw class, chunk, dw) so that the two tiles of a box (kw = class + 2 dw) are neighbours.
__device__ __forceinline__ void wgrad_halo_tile(const dcv_geom& g, int clchunks, int mode, int tile, int& ta, int& cls, int& tc, int& clc) {
  if (mode == 2) {
    const int dw = tile % 2; tile /= 2;
    clc = tile % clchunks; tile /= clchunks;
    tc = tile % 2 + 2 * dw; tile /= 2;
  } else {
    clc = tile % clchunks; tile /= clchunks;
    tc = tile % g.kw; tile /= g.kw;
  }
  cls = tile % 2; ta = tile / 2;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapS, const TcWgradP p,
                float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[MAX_STAGES];
  __shared__ uint64_t empty_bar[MAX_STAGES];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;

  const dcv_geom& g = p.g;
  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform for the compiler
  const int tile0 = blockIdx.x * p.G;                       // first 128-row tile of this CTA
  const int tiles_total = (p.blocksA_total + p.nA - 1) / p.nA;
  int Gcur = tiles_total - tile0; if (Gcur > p.G) Gcur = p.G;
  const int bB0 = blockIdx.y * p.nbB;                       // first S block of this CTA
  int nbB = p.blocksB_total - bB0; if (nbB > p.nbB) nbB = p.nbB;
  const int64_t pt_begin = (int64_t)blockIdx.z * p.ptiles_per_split;
  int64_t pt_end = pt_begin + p.ptiles_per_split; if (pt_end > p.ptiles_total) pt_end = p.ptiles_total;

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int blkA_bytes = p.pix * p.cbA * 2;
  const int blkB_bytes = p.pix * p.cbB * 2;
  // shared memory of the A operand of a PAIR of accumulator tiles, and where the odd tile of a pair starts
  const int pairA_bytes = p.halo == 2 ? p.boxA_bytes : 2 * (p.halo ? p.boxA_bytes : blkA_bytes * p.nA);
  const int oddA_bytes = p.halo == 2 ? p.cbA * 2 : pairA_bytes / 2;
  const int stage_bytes = blkB_bytes * p.nbB + (p.halo == 2 ? (p.G / 2) * p.boxA_bytes : (pairA_bytes / 2) * p.G);
  int blocksA_here = p.blocksA_total - tile0 * p.nA;        // valid A blocks of this CTA
  if (blocksA_here > Gcur * p.nA) blocksA_here = Gcur * p.nA;
  const int loadsA = p.halo == 2 ? Gcur / 2 : (p.halo ? Gcur : blocksA_here);          // TMA boxes per stage for A
  const int bytesA = p.halo ? loadsA * p.boxA_tx : blkA_bytes * blocksA_here;

  if (warp == 0 && lane == 0) { tmap_prefetch(&mapL); tmap_prefetch(&mapS); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(&tmem_full_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail;
  // global memory is only touched from here on
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // TMA producer.  A stage is many small boxes (one per (tap, channel chunk) block); the 32 lanes issue them in
    // parallel after lane 0 has armed the barrier, otherwise the per-instruction TMA latency serialises.  Everything
    // that does not depend on the pixel tile - which box a lane loads, its channel / tap offsets, its shared-memory
    // offset - is computed once; the pixel-tile origin advances like an odometer.  (The first version re-derived the
    // tap coordinates with ~10 integer divisions per lane per stage: the MMA warp then waited on this warp for 58 % of
    // the kernel, profiles/r1j_ncu_full_summary.md.)
    int stage = 0; uint32_t phase = 0;
    const int nloads = nbB + loadsA;
    int l_isB[2], l_c[2], l_dw[2], l_dh[2], l_dt[2]; uint32_t l_off[2]; bool l_on[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      const int i = lane + 32 * sl;
      l_on[sl] = i < nloads;
      l_isB[sl] = i < nbB;
      l_c[sl] = 0; l_dw[sl] = l_dh[sl] = l_dt[sl] = 0; l_off[sl] = 0;
      if (!l_on[sl]) continue;
      if (l_isB[sl]) {
        l_c[sl] = (bB0 + i) * p.cbB; l_off[sl] = (uint32_t)(i * blkB_bytes);
      } else if (p.halo) {
        const int b = i - nbB;
        int ta, cls, tc, clc;
        wgrad_halo_tile(g, p.clchunks, p.halo, p.halo == 2 ? tile0 + 2 * b : tile0 + b, ta, cls, tc, clc);   // mode 2: box b = tiles 2b, 2b + 1
        l_c[sl] = clc * p.cbA; l_dw[sl] = tc - g.pw; l_dh[sl] = cls - g.ph; l_dt[sl] = ta - g.pt;
        l_off[sl] = (uint32_t)(p.nbB * blkB_bytes + b * p.boxA_bytes);
      } else {
        const int b = i - nbB;
        const int blk = tile0 * p.nA + b;
        const int tap = blk / p.clchunks, clc = blk % p.clchunks;
        const int tc = tap % g.kw, tb = (tap / g.kw) % g.kh, ta = tap / (g.kw * g.kh);
        l_c[sl] = clc * p.cbA; l_dw[sl] = tc - g.pw; l_dh[sl] = tb - g.ph; l_dt[sl] = ta - g.pt;
        l_off[sl] = (uint32_t)(p.nbB * blkB_bytes + b * blkA_bytes);
      }
    }
    int w0, h0, t0, n0;
    {
      int64_t q = pt_begin;
      w0 = (int)(q % p.tiles_w) * p.bw; q /= p.tiles_w;
      h0 = (int)(q % p.tiles_h) * p.bh; q /= p.tiles_h;
      t0 = (int)(q % p.tiles_t) * p.bt; n0 = (int)(q / p.tiles_t) * p.bn;
    }
    const int wend = p.tiles_w * p.bw, hend = p.tiles_h * p.bh, tend = p.tiles_t * p.bt;
    for (int64_t pt = pt_begin; pt < pt_end; ++pt) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      const bool skip = (TC_DBG(p) & 3) && (pt - pt_begin) >= p.stages;
      if (lane == 0) mbar_expect_tx(&full_bar[stage], skip ? 0u : (uint32_t)(blkB_bytes * nbB + bytesA));
      __syncwarp();
      const uint32_t s_dst = sbase + stage * stage_bytes;
      if (!skip) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          if (!l_on[sl]) continue;
          if (l_isB[sl]) tma_load_5d(s_dst + l_off[sl], &mapS, &full_bar[stage], l_c[sl], w0, h0, t0, n0);
          else tma_load_5d(s_dst + l_off[sl], &mapL, &full_bar[stage], l_c[sl], w0 * g.sw + l_dw[sl], h0 * g.sh + l_dh[sl],
                           t0 * g.st + l_dt[sl], n0);
        }
      }
      w0 += p.bw;
      if (w0 >= wend) { w0 = 0; h0 += p.bh; if (h0 >= hend) { h0 = 0; t0 += p.bt; if (t0 >= tend) { t0 = 0; n0 += p.bn; } } }
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // MMA issuer: uniform control flow, elected lane issues (see "lean MMA issue")
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(128, p.Ns, 1, 1);
    // K step = 16 pixels, in 16-byte units; SBO = distance between consecutive 8-pixel groups.  Mode 2: a group is one 8-pixel
    // tile row inside a 9-pixel box row, so both follow the box row pitch
    const uint32_t kstepA16 = (p.halo == 2 ? 2u * (uint32_t)p.rpA : 16u) * (uint32_t)p.cbA * 2u >> 4, kstepB16 = (16u * (uint32_t)p.cbB * 2u) >> 4;
    const uint32_t ahi = sdesc_hi((p.halo == 2 ? (uint32_t)p.rpA : 8u) * (uint32_t)p.cbA * 2u, (uint32_t)p.layA),
                   bhi = sdesc_hi(8u * (uint32_t)p.cbB * 2u, (uint32_t)p.layB);
    const uint32_t pairA16 = (uint32_t)pairA_bytes >> 4, oddA16 = (uint32_t)oddA_bytes >> 4;
    // halo modes: the tile's second 64-channel block is the same box one box row further down
    const uint32_t rowA_bytes = (uint32_t)(p.rpA * p.cbA * 2);
    const uint32_t lo_b0 = sdesc_lo(sbase, (uint32_t)blkB_bytes),
                   lo_a0 = sdesc_lo(sbase + p.nbB * blkB_bytes, p.halo ? rowA_bytes : (uint32_t)blkA_bytes);
    const uint32_t slabA16 = p.halo ? rowA_bytes >> 4 : 0u;      // extra offset per (t, n) slab of the tile: its halo row
    const int nst = (int)(pt_end - pt_begin);
    // the hot loop is instantiated per (K steps per stage, accumulator tiles) so that a stage is straight-line code:
    // wait, fence, G*KS MMAs, commit (a single warp pays ~5 cycles per dependent control instruction)
#define DCV_WG_LOOP(KS_, G_)                                                                                                     \
    wgrad_mma_loop<KS_, G_>(full_bar, empty_bar, nst, p.stages, (uint32_t)stage_bytes >> 4, lo_a0, lo_b0, ahi, bhi, kstepA16,      \
                            kstepB16, pairA16, oddA16, tmem_base, (uint32_t)p.Ns, idesc, leader, slabA16, p.slab_shift)
#define DCV_WG_KS(G_)                                                                                                            \
    do { if (p.pix == 128) DCV_WG_LOOP(8, G_); else if (p.pix == 64) DCV_WG_LOOP(4, G_); else DCV_WG_LOOP(2, G_); } while (0)
    if (Gcur == 4) DCV_WG_KS(4); else if (Gcur == 3) DCV_WG_KS(3); else if (Gcur == 2) DCV_WG_KS(2); else DCV_WG_KS(1);
#undef DCV_WG_KS
#undef DCV_WG_LOOP
    if (leader) umma_commit(&tmem_full_bar);
    __syncwarp();
  } else {
    const int quarter = warp % 4;
    const int row = quarter * 32 + lane;
    const int taps = g.kt * g.kh * g.kw;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const bool has_work = pt_end > pt_begin;
    const int cs0 = bB0 * p.cbB;
    const int ncols = nbB * p.cbB;                           // valid columns of this CTA
    for (int gi = 0; gi < Gcur; ++gi) {
      const int blk = (tile0 + gi) * p.nA + row / p.cbA;
      const bool row_ok = blk < p.blocksA_total;
      int tap = row_ok ? blk / p.clchunks : 0, clc = row_ok ? blk % p.clchunks : 0;
      if (p.halo) {
        int ta, cls, tc;
        wgrad_halo_tile(g, p.clchunks, p.halo, tile0 + gi, ta, cls, tc, clc);
        tap = (ta * g.kh + cls + 2 * (row / p.cbA)) * g.kw + tc;
      }
      const int cl = clc * p.cbA + row % p.cbA;
      float* out = partial + (((int64_t)blockIdx.z * taps + tap) * g.Cl + cl) * g.Cs + cs0;
      auto store16 = [&](const uint32_t* v, int cb) {
        if (!row_ok || cb >= ncols) return;
        float4* dst = reinterpret_cast<float4*>(out + cb);
        if (has_work) {
          dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
          dst[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
          dst[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
        } else {
          dst[0] = dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      int cb = 0;
      for (; cb + 32 <= p.Ns; cb += 32) {                       // 32 columns per TMEM round trip
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(gi * p.Ns + cb), v);
        tmem_ld_wait();
        store16(v, cb);
        store16(v + 16, cb + 16);
      }
      for (; cb < p.Ns; cb += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(gi * p.Ns + cb), v);
        store16(v, cb);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ packing
// gather : out[n][tap][k]                 (n < npad, k < Kc)
// scatter: out[phase][n][j][k]            (j = tap-in-phase index in the kernel's loop order)
template <typename OT>
__device__ __forceinline__ void pack_weight_tc_elem(const ConvP& c, const float* __restrict__ w, int64_t s_l, int64_t s_s,
                                                    int64_t s_tap, const WeightWin& win, int npad, int64_t i,
                                                    OT* __restrict__ out) {
  // one thread per (phase, n, k): it walks the taps of its phase, so the fp32 reads of a thread fall into one or two
  // cache lines (taps are the innermost dimension of the PyTorch layouts) and the bf16 writes of a warp are contiguous in k
  const int64_t per_phase_nk = (int64_t)npad * c.Kc;
  const int ph = (int)(i / per_phase_nk);
  const int64_t r = i - (int64_t)ph * per_phase_nk;
  const int k = (int)(r % c.Kc); const int n = (int)(r / c.Kc);
  const PhaseInfo f = make_phase(c, ph);
  const int ntaps = f.nt * f.nh * f.nw;
  // gather: reduction channel k is an L channel, produced channel n is an S channel; scatter: swapped
  const int cl = c.scatter ? n : k, cs = c.scatter ? k : n;
  const bool real = win.has(cl, cs);
  if (!real && !win.fill) return;
  const float* wb = w + (cl - win.cl_off) * s_l + (cs - win.cs_off) * s_s;
  OT* ob = out + ((int64_t)ph * npad + n) * ntaps * c.Kc + k;
  int j = 0;
  for (int jt = 0; jt < f.nt; ++jt)
    for (int jh = 0; jh < f.nh; ++jh)
      for (int jw = 0; jw < f.nw; ++jw, ++j) {
        const int tap = ((f.a0t + f.ast * jt) * c.kh + (f.a0h + f.ash * jh)) * c.kw + (f.a0w + f.asw * jw);
        stf(ob + (int64_t)j * c.Kc, real ? wb[tap * s_tap] : 0.f);
      }
}

template <typename OT>
__global__ void __launch_bounds__(256)
pack_weight_tc_kernel(ConvP c, const float* __restrict__ w, int64_t s_l, int64_t s_s, int64_t s_tap, WeightWin win,
                      int npad, int phases, OT* __restrict__ out) {
  pdl_wait(); pdl_trigger();
  const int64_t total = (int64_t)npad * c.Kc * phases;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    pack_weight_tc_elem(c, w, s_l, s_s, s_tap, win, npad, i, out);
}

// One packed matrix assembled from several master weights that each own a window of it (merged / folded stems), one launch
constexpr int PACK_MULTI_MAX = 16;
struct PackPart { const float* w; int64_t s_l, s_s, s_tap; WeightWin win; };
struct PackParts { int n; PackPart p[PACK_MULTI_MAX]; };

__global__ void __launch_bounds__(256)
pack_weight_tc_multi_kernel(ConvP c, const __grid_constant__ PackParts parts, int npad, int phases, __nv_bfloat16* __restrict__ out) {
  pdl_wait(); pdl_trigger();
  const int64_t per_phase_nk = (int64_t)npad * c.Kc;
  const int64_t total = per_phase_nk * phases;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ph = (int)(i / per_phase_nk);
    const int64_t r = i - (int64_t)ph * per_phase_nk;
    const int k = (int)(r % c.Kc); const int n = (int)(r / c.Kc);
    const int cl = c.scatter ? n : k, cs = c.scatter ? k : n;
    int pi = -1;
    for (int q = 0; q < parts.n; ++q) if (parts.p[q].win.has(cl, cs)) { pi = q; break; }
    WeightWin win;                                   // the owning part's window, or an empty zero-filling one
    const float* w = nullptr; int64_t s_l = 0, s_s = 0, s_tap = 0;
    if (pi >= 0) { win = parts.p[pi].win; w = parts.p[pi].w; s_l = parts.p[pi].s_l; s_s = parts.p[pi].s_s; s_tap = parts.p[pi].s_tap; }
    else { win.cl_off = win.cs_off = 0; win.cl_cnt = win.cs_cnt = 0; }
    win.fill = 1;
    pack_weight_tc_elem(c, w, s_l, s_s, s_tap, win, npad, i, out);
  }
}

// Many weights in one launch (all layers of a network after its Adam step): the jobs travel as a kernel parameter
// (<= 4 KB), a block finds its job by scanning the block offsets.  ~90 single packs per iteration were ~90 tiny launches.
constexpr int PACK_BATCH_MAX = 20;
struct PackJob {
  ConvP c; const float* w; int64_t s_l, s_s, s_tap; WeightWin win; int npad, phases; __nv_bfloat16* out; int64_t total; int block0;
};
struct PackBatch { int n; PackJob jobs[PACK_BATCH_MAX]; };

__global__ void __launch_bounds__(256)
pack_weight_tc_batch_kernel(const __grid_constant__ PackBatch b) {
  pdl_wait(); pdl_trigger();
  int ji = 0;
  while (ji + 1 < b.n && (int)blockIdx.x >= b.jobs[ji + 1].block0) ++ji;
  const PackJob& j = b.jobs[ji];
  const int nblocks = (ji + 1 < b.n ? b.jobs[ji + 1].block0 : (int)gridDim.x) - j.block0;
  for (int64_t i = (int64_t)(blockIdx.x - j.block0) * blockDim.x + threadIdx.x; i < j.total; i += (int64_t)nblocks * blockDim.x)
    pack_weight_tc_elem(j.c, j.w, j.s_l, j.s_s, j.s_tap, j.win, j.npad, i, j.out);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// activation map: dims {C, W, H, T, N}, box {cbox, bw*ew, bh*eh, bt*et, bn}, traversal strides {1, ew, eh, et, 1}
static int make_act_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int T, int N, int64_t ld, int cbox,
                        int bw, int bh, int bt, int bn, int ew, int eh, int et, CUtensorMapSwizzle swz, int esize = 2) {
  EncodeTiledFn enc = get_encode();
  DCV_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
  DCV_REQUIRE(((uintptr_t)ptr & 15) == 0 && ((ld * esize) % 16) == 0, "TMA operand must be 16-byte aligned (ptr %p, ld %lld)", ptr, (long long)ld);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  const cuuint64_t es = (cuuint64_t)esize;
  cuuint64_t strides[4] = {(cuuint64_t)ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es, (cuuint64_t)T * H * W * ld * es};
  cuuint32_t box[5] = {(cuuint32_t)cbox, (cuuint32_t)(bw * ew), (cuuint32_t)(bh * eh), (cuuint32_t)(bt * et), (cuuint32_t)bn};
  cuuint32_t estr[5] = {1, (cuuint32_t)ew, (cuuint32_t)eh, (cuuint32_t)et, 1};
  for (int i = 0; i < 5; ++i) DCV_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box dim %d = %u out of range", i, box[i]);
  CUresult r = enc(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with %d", (int)r);
  return 0;
}

static int make_weight_map(CUtensorMap* m, const void* ptr, int64_t K, int npad, int phases, int cbox, int nbox,
                           CUtensorMapSwizzle swz, int esize = 2) {
  EncodeTiledFn enc = get_encode();
  DCV_REQUIRE(enc, "cuTensorMapEncodeTiled entry point unavailable");
  DCV_REQUIRE(((uintptr_t)ptr & 15) == 0, "packed weight must be 16-byte aligned");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)npad, (cuuint64_t)phases};
  cuuint64_t strides[2] = {(cuuint64_t)K * esize, (cuuint64_t)K * npad * esize};
  cuuint32_t box[3] = {(cuuint32_t)cbox, (cuuint32_t)nbox, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed with %d", (int)r);
  return 0;
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

// choose a box (bw,bh,bt,bn) with product `target` over an index space (Qw,Qh,Qt,N)
static void choose_box(int target, int Qw, int Qh, int Qt, int N, int* bw, int* bh, int* bt, int* bn) {
  int rem = target;
  *bw = pow2_floor(Qw < rem ? Qw : rem); rem /= *bw;
  *bh = pow2_floor(Qh < rem ? Qh : rem); rem /= *bh;
  int t = 1;
  while (t * 2 <= rem && Qt % (t * 2) == 0) t *= 2;
  *bt = t; rem /= t;
  *bn = rem;
  (void)N;
}

int tc_npad(int Nc) { return (Nc + 15) / 16 * 16; }
static int tc_bnt(int npad, int cap = 256) {
  for (int b = cap; b >= 16; b -= 16) if (npad % b == 0) return b;
  return 16;
}

int conv_tc_supported(const dcv_geom* g, int dir) {
  const ConvP c = make_convp(g, dir);
  if (c.Kc % 16 != 0) return 0;
  if (c.scatter) {
    if (g->kt % g->st || g->kh % g->sh || g->kw % g->sw) return 0;
  } else {
    if (g->st > 8 || g->sh > 8 || g->sw > 8) return 0;
  }
  return 1;
}

int64_t packed_weight_tc_bytes(const dcv_geom* g, int dir, int tf32) {
  const ConvP c = make_convp(g, dir);
  const int taps = g->kt * g->kh * g->kw;   // sum over phases of taps-in-phase == taps when k % s == 0
  return (int64_t)tc_npad(c.Nc) * taps * c.Kc * (tf32 ? 4 : 2);
}

int pack_weight_tc(const dcv_geom* g, int dir, const float* w, int64_t s_l, int64_t s_s, int64_t s_tap, WeightWin win,
                   void* out, cudaStream_t s, int tf32) {
  DCV_REQUIRE(conv_tc_supported(g, dir), "pack_weight_tc: geometry not supported by the tcgen05 kernel");
  const ConvP c = make_convp(g, dir);
  const int phases = c.scatter ? g->st * g->sh * g->sw : 1;
  const int64_t total = (int64_t)tc_npad(c.Nc) * c.Kc * phases;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  if (tf32) launch_k(pack_weight_tc_kernel<float>, blocks, 256, 0, s, c, w, s_l, s_s, s_tap, win, tc_npad(c.Nc), phases, (float*)out);
  else launch_k(pack_weight_tc_kernel<__nv_bfloat16>, blocks, 256, 0, s, c, w, s_l, s_s, s_tap, win, tc_npad(c.Nc), phases, (__nv_bfloat16*)out);
  return check_launch("pack_weight_tc");
}

int pack_weight_tc_multi(const dcv_geom* g, int dir, int n, const float* const* w, const int64_t* s_l, const int64_t* s_s,
                         const int64_t* s_tap, const int* cl_off, const int* cl_cnt, const int* cs_off, const int* cs_cnt, void* out,
                         cudaStream_t s) {
  DCV_REQUIRE(conv_tc_supported(g, dir), "pack_weight_multi: geometry not supported by the tcgen05 kernel");
  DCV_REQUIRE(n >= 1 && n <= PACK_MULTI_MAX, "pack_weight_multi: %d parts (max %d)", n, PACK_MULTI_MAX);
  PackParts parts; parts.n = n;
  for (int i = 0; i < n; ++i) {
    PackPart& q = parts.p[i];
    q.w = w[i]; q.s_l = s_l[i]; q.s_s = s_s[i]; q.s_tap = s_tap[i];
    q.win.cl_off = cl_off[i]; q.win.cl_cnt = cl_cnt[i]; q.win.cs_off = cs_off[i]; q.win.cs_cnt = cs_cnt[i]; q.win.fill = 1;
  }
  const ConvP c = make_convp(g, dir);
  const int phases = c.scatter ? g->st * g->sh * g->sw : 1;
  const int64_t total = (int64_t)tc_npad(c.Nc) * c.Kc * phases;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  launch_k(pack_weight_tc_multi_kernel, blocks, 256, 0, s, c, parts, tc_npad(c.Nc), phases, (__nv_bfloat16*)out);
  return check_launch("pack_weight_tc_multi");
}

int pack_weight_tc_batch(int n, const dcv_geom* const* geoms, const int* dirs, const float* const* w, const int64_t* s_l,
                         const int64_t* s_s, const int64_t* s_tap, void* const* outs, cudaStream_t s) {
  int i = 0;
  while (i < n) {
    PackBatch b; b.n = 0;
    int blocks = 0;
    for (; i < n && b.n < PACK_BATCH_MAX; ++i) {
      const dcv_geom* g = geoms[i];
      DCV_REQUIRE(conv_tc_supported(g, dirs[i]), "pack_weight_batch: job %d not supported by the tcgen05 kernel", i);
      PackJob& j = b.jobs[b.n++];
      j.c = make_convp(g, dirs[i]);
      j.w = w[i]; j.s_l = s_l[i]; j.s_s = s_s[i]; j.s_tap = s_tap[i]; j.win = full_window(g);
      j.npad = tc_npad(j.c.Nc); j.phases = j.c.scatter ? g->st * g->sh * g->sw : 1;
      j.out = (__nv_bfloat16*)outs[i];
      j.total = (int64_t)j.npad * j.c.Kc * j.phases;
      j.block0 = blocks;
      int nb = (int)((j.total + 255) / 256); if (nb > 148 * 2) nb = 148 * 2;
      blocks += nb;
    }
    launch_k(pack_weight_tc_batch_kernel, blocks, 256, 0, s, b);
    if (int rc = check_launch("pack_weight_tc_batch")) return rc;
  }
  return 0;
}

// stats != NULL: also accumulate per-channel sum / sum of squares of the outputs (fused BatchNorm statistics, persistent
// TMA-store path only).  slots_out != NULL: planning query - store the number of [2][npad] partial-sum slots such a launch
// writes (0 = this geometry cannot fuse the statistics) and return without launching anything.
int head_loss_supported(const dcv_geom* g, int dir) {
  if (!conv_tc_supported(g, dir)) return 0;
  const ConvP c = make_convp(g, dir);
  return !c.scatter && c.wN == 1 && c.Kc % 8 == 0;
}

int head_loss(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int kind, float* loss_out,
              int accumulate, void* dy, int64_t lddy, float grad_scale, unsigned* counter, cudaStream_t s) {
  DCV_REQUIRE(head_loss_supported(g, dir) && (ldx % 8) == 0 && (((uintptr_t)x) & 15) == 0, "head_loss: not a single-channel head on an aligned bf16 input");
  DCV_REQUIRE(kind >= 0 && kind <= 4, "head_loss: unknown loss kind %d", kind);
  const ConvP c = make_convp(g, dir);
  const int64_t M = (int64_t)c.N * c.Ot * c.Oh * c.Ow;
  DCV_REQUIRE(M > 0, "head_loss: empty logits");
  launch_k(head_gemv_loss_kernel, (unsigned)((M + 7) / 8), 256, 0, s, c, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)wp,
           (__nv_bfloat16*)y, ldy, kind, loss_out, accumulate, (__nv_bfloat16*)dy, lddy, grad_scale, counter);
  return check_launch("head_gemv_loss");
}

int conv_tf32_supported(const dcv_geom* g, int dir) {
  if (!conv_tc_supported(g, dir)) return 0;
  const ConvP c = make_convp(g, dir);
  return c.Kc % 32 == 0 && c.Nc % 32 == 0 && c.wN > 1;
}

// tf32 != 0: fp32 activations / packed weights / output, kind::tf32 MMAs (x, wp, y are float buffers, ldx / ldy in floats)
// accumulate != 0: y += result (TMA reduce-add epilogue; needs the TMA-store path, fails otherwise)
int conv_tc(const dcv_geom* g, int dir, const void* x, int64_t ldx, const void* wp, void* y, int64_t ldy, int act,
            float slope, float* stats, int* slots_out, cudaStream_t s, int tf32, int accumulate) {
  if (slots_out) *slots_out = 0;
  DCV_REQUIRE(conv_tc_supported(g, dir), "conv_tc: geometry not supported by the tcgen05 kernel");
  DCV_REQUIRE(!tf32 || (conv_tf32_supported(g, dir) && !stats && !slots_out), "conv_tc: geometry not supported by the tf32 variant");
  const int esz = tf32 ? 4 : 2;
  TcConvP p;
  p.c = make_convp(g, dir);
  const ConvP& c = p.c;
  if (!tf32 && !accumulate && !c.scatter && c.wN == 1 && c.Kc % 8 == 0 && (ldx % 8) == 0 && (((uintptr_t)x) & 15) == 0 && !g_tune.no_gemv) {
    if (slots_out) return 0;
    DCV_REQUIRE(!stats, "conv_tc: fused statistics are not available for the single-channel head kernel");
    const int64_t M = (int64_t)c.N * c.Ot * c.Oh * c.Ow;          // one warp per logit
    launch_k(head_gemv_kernel, (unsigned)((M + 7) / 8), 256, 0, s, c, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)wp,
                                                           (__nv_bfloat16*)y, ldy, act, slope);
    return check_launch("head_gemv");
  }
  const int phases = c.scatter ? g->st * g->sh * g->sw : 1;
  const PhaseInfo f0 = make_phase(c, 0);
  int bnt_cap = 256;
replan:
  choose_box(128, f0.Qw, f0.Qh, f0.Qt, c.N, &p.bw, &p.bh, &p.bt, &p.bn);
  p.tiles_w = ceil_div(f0.Qw, p.bw); p.tiles_h = ceil_div(f0.Qh, p.bh); p.tiles_t = ceil_div(f0.Qt, p.bt);
  p.tiles_n = ceil_div(c.N, p.bn);
  p.cblk = tf32 ? 32 : (c.Kc % 64 == 0 ? 64 : (c.Kc % 32 == 0 ? 32 : 16));   // 128-byte rows whenever the channel count allows
  p.kchunks = c.Kc / p.cblk;
  const int npad = tc_npad(c.Nc);
  p.bnt = tc_bnt(npad, bnt_cap);
  const int row_bytes = p.cblk * esz;
  p.swz_layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  p.b_bytes = (p.bnt * p.cblk * esz + 1023) / 1024 * 1024;
  const int ntaps0 = f0.nt * f0.nh * f0.nw;
  { const char* e = exp_env("DCV_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.ldy = ldy; p.act = act; p.slope = slope;
  p.vec_ok = (((uintptr_t)y & 15) == 0) && ((ldy * esz) % 16 == 0);

  CUtensorMap mapA, mapB;
  int rc = 0;
  const bool g4 = !tf32 && c.Kc == 16 && ntaps0 >= 4 && !g_tune.no_tapgroup;   // grouped-tap mode for 16-channel (image-like) operands
  const int64_t Kph = (int64_t)ntaps0 * c.Kc;  // K extent of one phase

  {
    // persistent kernel: one CTA per SM, two accumulator sets in TMEM
    static int num_sms = 0;
    if (num_sms == 0) {
      int dev = 0;
      DCV_CUDA(cudaGetDevice(&dev));
      DCV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    p.ntn = npad / p.bnt;
    DCV_REQUIRE(phases <= 8, "conv_tc: %d sub-pixel phases", phases);
    for (int ph = 0; ph < phases; ++ph) p.phs[ph] = make_phase(c, ph);
    p.tma_store = 0;                                         // 1: 64-channel chunks, 2: one 16- / 32-channel chunk
    if (tf32) {
      DCV_REQUIRE(p.vec_ok, "conv_tc (tf32): the output must be 16-byte aligned");
      p.tma_store = 1;
    } else if ((((uintptr_t)y & 15) == 0) && (ldy % 8 == 0) && !g_tune.no_tma_store) {
      if (p.bnt % 64 == 0) p.tma_store = 1;
      else if ((p.bnt == 16 || p.bnt == 32) && c.Nc >= p.bnt && !g_tune.no_narrow_tma_store) p.tma_store = 2;
    }
    const int stg_bytes = p.tma_store ? 2 * 16384 : 0;
    p.b_bytes = g4 ? (p.bnt * 128 + 1023) / 1024 * 1024 : p.b_bytes;

    // Row-halo sharing: the h taps that read the same strided lattice (every sh-th tap of a strided gather, all nh taps of
    // a sub-pixel phase / stride-1 correlation) differ by whole tile rows, so ONE A box with hg-1 extra rows serves all of
    // them - the MMA descriptors just start hg-1 .. 0 rows further down, which is a multiple of the swizzle atom when a
    // tile row holds a multiple of 8 pixels.  L2 -> SM traffic for A drops to (bh + hg - 1) / (hg * bh) of the per-tap
    // boxes; the tile is re-shaped to 16 x 8 so that the halo stays small.  Needs a tile that lies in one (n, t) slice.
    p.hg = 1; p.hcls = f0.nh;
    if (!g4 && !g_tune.nohalo) {
      const int hg = c.scatter ? f0.nh : (g->kh % g->sh == 0 ? g->kh / g->sh : 1);
      const int hbw = f0.Qw >= 16 ? 16 : (f0.Qw >= 8 ? 8 : 0);
      const int hbh = hbw ? 128 / hbw : 0;
      bool same = true;                                       // every phase has the same number of h taps
      for (int ph = 1; ph < phases; ++ph) same = same && (make_phase(c, ph).nh == f0.nh);
      const int bsz = hg * p.b_bytes;
      if ((hg == 2 || hg == 3) && hbw && f0.Qh >= hbh && same && row_bytes >= 64 && bsz <= 96 * 1024) {
        p.hg = hg; p.hcls = f0.nh / hg;
        p.bw = hbw; p.bh = hbh; p.bt = 1; p.bn = 1;
        p.tiles_w = ceil_div(f0.Qw, p.bw); p.tiles_h = ceil_div(f0.Qh, p.bh); p.tiles_t = f0.Qt;
      }
    }
    const int box_rows = (p.bh + p.hg - 1) * p.bw * p.bt * p.bn;                       // A rows of one M tile incl. halo
    p.a_tile16 = (box_rows * p.cblk * esz) >> 4;
    const int64_t sp_tiles = (int64_t)p.tiles_w * p.tiles_h * p.tiles_t * p.ntn * phases;
    auto balance = [&](int mt) {
      const int64_t items = sp_tiles * ceil_div(c.N, p.bn * mt);
      return (double)items / ((double)num_sms * (double)((items + num_sms - 1) / num_sms));
    };
    auto a_bytes_of = [&](int mt) { return g4 ? TAPG * mt * 128 * 32 : mt * box_rows * p.cblk * esz; };
    auto stages_of = [&](int mt) { return (222 * 1024 - stg_bytes) / (a_bytes_of(mt) + p.hg * p.b_bytes); };
    // M tiles per work item: more tiles share every weight tile and amortise the per-stage barrier handshake over more
    // MMAs (>= 512 tensor cycles per stage wanted: hg * mt * bnt >= 256), as long as two accumulator sets fit in TMEM, the
    // TMA box stays <= 256 samples, >= 3 ring stages remain and the load balance over the SMs does not suffer
    p.mt = 1;
    for (int mt = 2; mt <= 4; mt *= 2) {
      if (2 * mt * p.bnt > 512 || p.bn * mt > 256 || c.N < mt * p.bn) break;
      if (p.hg * mt * p.bnt > 256 && p.hg * p.mt * p.bnt >= 256) break;
      if (stages_of(mt) < 3) break;
      if (balance(mt) < 0.9 * balance(p.mt)) break;
      p.mt = mt;
    }
    if (g_tune.mt) {                                 // tuning override: force 1 / 2 / 4 when legal
      const int mt = g_tune.mt;
      if ((mt == 1 || mt == 2 || mt == 4) && 2 * mt * p.bnt <= 512 && p.bn * mt <= 256 && c.N >= mt * p.bn && stages_of(mt) >= 2) p.mt = mt;
    }
    p.tiles_n = ceil_div(c.N, p.bn * p.mt);
    p.items = (int)(sp_tiles * p.tiles_n);
    // Few work items (the deep 1x1 .. 4x4 layers: 4-32 items, each a K loop of 4096-8192 fed at ONE SM's L2 bandwidth): a
    // narrower N tile spreads the same work over 2-4x as many SMs; the A tile is then fetched by more CTAs, from L2.
    if (!g4 && !g_tune.no_nsplit && p.items * 2 <= num_sms && p.bnt % 128 == 0) { bnt_cap = p.bnt / 2; goto replan; }
    p.a_bytes = a_bytes_of(p.mt);
    p.tx_bytes = p.a_bytes + p.hg * p.bnt * p.cblk * esz;
    p.acc_cols = p.mt * p.bnt;
    p.tmem_cols = pow2_ceil(2 * p.acc_cols < 32 ? 32 : 2 * p.acc_cols);
    int st = stages_of(p.mt);
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (st < 1) st = 1;
    p.stages = st;
    // sm_reserve > 0 (data-parallel runs, while a gradient bucket is in flight): leave a few SMs to NCCL's CTAs
    const int sms_avail = num_sms - g_tune.sm_reserve > 8 ? num_sms - g_tune.sm_reserve : 8;
    const int grid_p = p.items < sms_avail ? p.items : sms_avail;
    const bool can_stats = !tf32 && p.tma_store == 1 && !g_tune.no_fused_stats;
    if (slots_out) { *slots_out = can_stats ? 4 * grid_p : 0; return 0; }
    DCV_REQUIRE(!stats || can_stats, "conv_tc: fused statistics need the TMA-store epilogue (output channels %% 64 == 0)");
    p.stats = stats; p.npad = npad;
    p.accumulate = accumulate;
    DCV_REQUIRE(!accumulate || (p.tma_store && !stats && act == DCV_ACT_NONE), "conv_tc: accumulating into the output needs the TMA-store epilogue, no activation and no fused statistics");
    rc = make_act_map(&mapA, x, c.Kc, c.Iw, c.Ih, c.It, c.N, ldx, p.cblk, p.bw, p.bh + p.hg - 1, p.bt, p.bn * p.mt, f0.mulw, f0.mulh, f0.mult, swz, esz);
    if (rc) return rc;
    rc = g4 ? make_weight_map(&mapB, wp, Kph, npad, phases, 64, p.bnt, CU_TENSOR_MAP_SWIZZLE_128B)
            : make_weight_map(&mapB, wp, Kph, npad, phases, p.cblk, p.bnt, swz, esz);
    if (rc) return rc;
    CUtensorMap mapY = mapA;
    if (p.tma_store) {
      // output map: {Nc channels, Ow, Oh, Ot, N}; box = one M tile x 64 channels; the traversal strides are the sub-pixel
      // phase strides of a transposed convolution / strided data gradient (1 for plain convolutions)
      // box = the 32 tile rows one epilogue warp owns (tile extents are powers of two, so 32 consecutive rows are a box)
      int rem = 32;
      const int yw = p.bw < rem ? p.bw : rem; rem /= yw;
      const int yh = p.bh < rem ? p.bh : rem; rem /= yh;
      const int yt = p.bt < rem ? p.bt : rem; rem /= yt;
      const int yn = p.bn < rem ? p.bn : rem; rem /= yn;
      DCV_REQUIRE(rem == 1, "conv_tc: tile %dx%dx%dx%d has no 32-row sub-box", p.bw, p.bh, p.bt, p.bn);
      const int ycw = tf32 ? 32 : (p.tma_store == 1 ? 64 : p.bnt);      // 128-byte rows: 64 bf16 or 32 fp32 channels
      const int ybytes = ycw * esz;
      rc = make_act_map(&mapY, y, c.Nc, c.Ow, c.Oh, c.Ot, c.N, ldy, ycw, yw, yh, yt, yn, f0.osw, f0.osh, f0.ost,
                        ybytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (ybytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B), esz);
      if (rc) return rc;
    }
    const int smem_p = st * (p.a_bytes + p.hg * p.b_bytes) + stg_bytes + 1024;
    const int ks = g4 ? 0 : row_bytes / 32;                       // 32-byte K steps per stage
    ConvPersFn fn = tf32 ? conv_pers_variant_tf32(p.mt, p.hg) : conv_pers_variant(ks, p.mt, p.hg);
    DCV_REQUIRE(fn != nullptr, "conv_tc: no persistent kernel variant for cblk %d mt %d hg %d tf32 %d", p.cblk, p.mt, p.hg, tf32);
    static int smem_set_p[512] = {0};
    int& set = smem_set_p[((tf32 ? 8 : ks) * 8 + p.mt) * 4 + p.hg];
    if (smem_p > set) {
      DCV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p));
      set = smem_p;
    }
    launch_k(fn, grid_p, TC_THREADS, smem_p, s, mapA, mapB, mapY, p, (__nv_bfloat16*)y);
    return check_launch("conv_tc_pers");
  }

  return 0;
}

// ---- wgrad
static int block_width(int C) { return C % 64 == 0 ? 64 : (C % 32 == 0 ? 32 : 16); }
static int layout_code(int cb) { return cb == 64 ? 2 : (cb == 32 ? 4 : 6); }
static CUtensorMapSwizzle swizzle_of(int cb) {
  return cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int wgrad_tc_supported(const dcv_geom* g) {
  if (g->Cl % 16 || g->Cs % 16) return 0;
  if (g->st > 8 || g->sh > 8 || g->sw > 8) return 0;
  return 1;
}

static void wgrad_tc_plan(const dcv_geom* g, TcWgradP* p, int* splits) {
  p->g = *g;
  { const char* e = exp_env("DCV_TC_DBG"); p->dbg = e ? atoi(e) : 0; }
  const int taps = g->kt * g->kh * g->kw;
  p->cbA = block_width(g->Cl); p->nA = 128 / p->cbA; p->clchunks = g->Cl / p->cbA;
  p->blocksA_total = taps * p->clchunks;
  p->cbB = block_width(g->Cs); p->blocksB_total = g->Cs / p->cbB;
  p->nbB = (g_tune.wgrad_ns > 0 ? g_tune.wgrad_ns : 256) / p->cbB; if (p->nbB < 1) p->nbB = 1;
  if (p->nbB > p->blocksB_total) p->nbB = p->blocksB_total;
  p->Ns = p->nbB * p->cbB;
  p->layA = layout_code(p->cbA); p->layB = layout_code(p->cbB);
  const int tiles_total = ceil_div(p->blocksA_total, p->nA);
  p->G = 512 / pow2_ceil(p->Ns < 32 ? 32 : p->Ns);
  if (p->G > 4) p->G = 4;
  // Accumulator tiles per CTA: fewer tiles = more CTAs per pixel range = fewer pixel splits, and every split writes (and the
  // reduction re-reads) a full fp32 copy of the weight gradient - 39 MB per launch for a 64 -> 128 channel layer with 74 splits.
  // Measured per iteration (mug-depth, batch 32, same session): four tiles per CTA 9.70 ms, two (one column-halo box pair)
  // 9.59 ms, ONE (row halo only) 9.47 ms; narrowing the S tile as well (128 / 64 columns) 9.47 / 9.66 ms.  One tile per CTA is
  // the default; dcv_set_tuning("wgrad_g", 2 | 4) selects the wider plans (parity-tested, incl. the column halo).
  const int gcap = g_tune.wgrad_g > 0 ? g_tune.wgrad_g : 1;
  if (p->G > gcap) p->G = gcap;
  if (p->G > tiles_total) p->G = tiles_total;
  // pixels per stage.  Every (tap, channel chunk) block is its own TMA box, so with narrow blocks (16 or 32
  // channels: image-like tensors) a 32-pixel stage is dozens of 1-2 KB boxes and the kernel becomes bound by the TMA
  // issue rate, not by bytes: take 128-pixel stages there and fewer accumulator tiles per CTA instead.
  p->pix = 32;
  if (p->cbA <= 32) {
    p->pix = 128;
    while (p->G > 1 && (200 * 1024) / (p->pix * 2 * (p->Ns + 128 * p->G)) < 4) --p->G;
    while (p->pix > 32 && (200 * 1024) / (p->pix * 2 * (p->Ns + 128 * p->G)) < 3) p->pix /= 2;
  } else {
    while (p->pix < 128 && (200 * 1024) / (2 * p->pix * 2 * (p->Ns + 128 * p->G)) >= 4) p->pix *= 2;
  }
  p->tmem_cols = pow2_ceil(p->G * p->Ns < 32 ? 32 : p->G * p->Ns);
  choose_box(p->pix, g->Ws, g->Hs, g->Ts, g->N, &p->bw, &p->bh, &p->bt, &p->bn);
  int stage_bytes = p->pix * 2 * (p->Ns + 128 * p->G);
  p->halo = 0; p->boxA_bytes = 0; p->boxA_tx = 0; p->slab_shift = 0; p->rpA = 0;
  // tile width: 8 pixels (a tile row = one 1024-byte swizzle atom).  4-pixel-wide tiles (4x4 maps) shift the second block by
  // HALF an atom: the swizzle is a function of the absolute shared-memory address bits, so the shifted view stays consistent.
  const int hbw = g->Ws >= 8 ? 8 : (g->Ws >= 4 && g_tune.wgrad_halo4 ? 4 : 0);
  if (!g_tune.no_wgrad_halo && p->cbA == 64 && g->kh == 4 && g->sh == 2 && hbw && g->Hs * hbw >= 16) {
    // row-halo sharing (see TcWgradP): narrow tiles with as many rows as fit, the largest pixel count that leaves >= 4
    // ring stages (else >= 3)
    auto bh_of = [&](int pix) { return pow2_floor(g->Hs < pix / hbw ? g->Hs : pix / hbw); };
    auto stages_of = [&](int pix) {
      const int bh = bh_of(pix);
      return (200 * 1024) / (pix * 2 * p->Ns + p->G * (pix / (hbw * bh)) * (bh + 1) * hbw * 128);
    };
    int best = 0;
    for (int want = 4; want >= 3 && !best; --want)
      for (int pix = 128; pix >= 32 && !best; pix /= 2) if (stages_of(pix) >= want) best = pix;
    if (best) {
      p->halo = 1; p->pix = best;
      p->bw = hbw; p->bh = bh_of(best);
      int rem = best / (hbw * p->bh), t = 1;
      while (t * 2 <= rem && g->Ts % (t * 2) == 0) t *= 2;
      p->bt = t; p->bn = rem / t;
      p->rpA = hbw;
      p->boxA_bytes = p->boxA_tx = p->bt * p->bn * (p->bh + 1) * hbw * 128;
      int sh = 0; while ((1 << sh) < hbw * p->bh) ++sh;
      p->slab_shift = sh;
      stage_bytes = p->pix * 2 * p->Ns + p->G * p->boxA_bytes;
    }
    // column halo on top (mode 2): 8-pixel-wide tiles, kw = 4 with stride 2, an even number of tiles per CTA
    if (best && hbw == 8 && g->kw == 4 && g->sw == 2 && p->G % 2 == 0 && tiles_total % p->G == 0 && !g_tune.no_wgrad_whalo) {
      auto box2 = [&](int pix) { const int bh = bh_of(pix); return ((pix / (8 * bh)) * (bh + 1) * 9 * 128 + 1023) / 1024 * 1024; };
      auto stages2 = [&](int pix) { return (200 * 1024) / (pix * 2 * p->Ns + (p->G / 2) * box2(pix)); };
      int best2 = 0;
      for (int want = 4; want >= 3 && !best2; --want)
        for (int pix = 128; pix >= 32 && !best2; pix /= 2) if (stages2(pix) >= want) best2 = pix;
      if (best2) {
        p->halo = 2; p->pix = best2;
        p->bh = bh_of(best2);
        int rem = best2 / (8 * p->bh), t = 1;
        while (t * 2 <= rem && g->Ts % (t * 2) == 0) t *= 2;
        p->bt = t; p->bn = rem / t;
        p->rpA = 9;
        p->boxA_bytes = box2(best2);
        p->boxA_tx = p->bt * p->bn * (p->bh + 1) * 9 * 128;
        int sh = 0; while ((1 << sh) < 8 * p->bh) ++sh;
        p->slab_shift = sh;
        stage_bytes = p->pix * 2 * p->Ns + (p->G / 2) * p->boxA_bytes;
      }
    }
  }
  p->tiles_w = ceil_div(g->Ws, p->bw); p->tiles_h = ceil_div(g->Hs, p->bh); p->tiles_t = ceil_div(g->Ts, p->bt);
  p->tiles_n = ceil_div(g->N, p->bn);
  p->ptiles_total = (int64_t)p->tiles_w * p->tiles_h * p->tiles_t * p->tiles_n;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 1) stages = 1;
  p->stages = stages;
  const int64_t tiles = (int64_t)ceil_div(tiles_total, p->G) * ceil_div(p->blocksB_total, p->nbB);
  // the ring takes the whole shared memory, so exactly one CTA is resident per SM: aim for ONE wave of <= 148 CTAs
  // (a second wave only adds a second non-overlapped epilogue and doubles the partial sums that have to be reduced)
  const int sms = 148 - g_tune.sm_reserve > 8 ? 148 - g_tune.sm_reserve : 8;   // sm_reserve: SMs left to a collective
  int64_t sp = tiles >= sms ? 1 : sms / tiles;
  if (g_tune.wgrad_waves > 0) sp *= g_tune.wgrad_waves;
  const int64_t maxs = (p->ptiles_total + 7) / 8;
  if (sp > maxs) sp = maxs;
  if (sp > 296) sp = 296;
  if (sp < 1) sp = 1;
  p->ptiles_per_split = (p->ptiles_total + sp - 1) / sp;
  *splits = (int)((p->ptiles_total + p->ptiles_per_split - 1) / p->ptiles_per_split);
}

int64_t wgrad_tc_ws_bytes(const dcv_geom* g) {
  TcWgradP p; int splits;
  wgrad_tc_plan(g, &p, &splits);
  return (int64_t)splits * g->kt * g->kh * g->kw * g->Cl * g->Cs * sizeof(float);
}

int wgrad_reduce(const float* partial, int splits, const dcv_geom* g, float* dw, int64_t s_l, int64_t s_s,
                 int64_t s_tap, int accumulate, cudaStream_t s);

int wgrad_tc_splits(const dcv_geom* g) {
  TcWgradP p; int splits;
  wgrad_tc_plan(g, &p, &splits);
  return splits;
}

int wgrad_tc(const dcv_geom* g, const void* xl, int64_t ldl, const void* xs, int64_t lds, float* dw, int64_t s_l,
             int64_t s_s, int64_t s_tap, int accumulate, void* ws, int64_t ws_bytes, cudaStream_t s) {
  DCV_REQUIRE(wgrad_tc_supported(g), "wgrad_tc: geometry not supported by the tcgen05 kernel");
  TcWgradP p; int splits;
  wgrad_tc_plan(g, &p, &splits);
  DCV_REQUIRE(ws_bytes >= wgrad_tc_ws_bytes(g), "wgrad_tc workspace too small");
  CUtensorMap mapL, mapS;
  int rc = make_act_map(&mapL, xl, g->Cl, g->Wl, g->Hl, g->Tl, g->N, ldl, p.cbA, p.bw + (p.halo == 2), p.bh + (p.halo != 0), p.bt, p.bn,
                        g->sw, g->sh, g->st, swizzle_of(p.cbA));
  if (rc) return rc;
  rc = make_act_map(&mapS, xs, g->Cs, g->Ws, g->Hs, g->Ts, g->N, lds, p.cbB, p.bw, p.bh, p.bt, p.bn, 1, 1, 1,
                    swizzle_of(p.cbB));
  if (rc) return rc;
  const int smem = p.stages * (p.halo == 2 ? p.pix * 2 * p.Ns + (p.G / 2) * p.boxA_bytes
                                           : (p.halo ? p.pix * 2 * p.Ns + p.G * p.boxA_bytes : p.pix * 2 * (p.Ns + 128 * p.G))) + 1024;
  static int smem_set = 0;
  if (smem > smem_set) {
    DCV_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    smem_set = smem;
  }
  dim3 grid(ceil_div(ceil_div(p.blocksA_total, p.nA), p.G), ceil_div(p.blocksB_total, p.nbB), splits);
  launch_k(wgrad_tc_kernel, grid, TC_THREADS, smem, s, mapL, mapS, p, (float*)ws);
  rc = check_launch("wgrad_tc");
  if (rc) return rc;
  if (!dw) return 0;   // partial sums only; the caller reduces them (dcv_wgrad_reduce_sub)
  return wgrad_reduce((const float*)ws, splits, g, dw, s_l, s_s, s_tap, accumulate, s);
}

}  // namespace dcv
