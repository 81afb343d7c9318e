// Raw tcgen05.mma issue-rate probe for sm_100a (timing experiment, not part of the library).
//
// Question it answers: how fast does ONE CTA (or a CTA pair) retire kind::f16 MMAs of shape M=128 (256 for a pair) x N x 16
// with both operands in shared memory (128B-swizzled K-major tiles, K block = 64 like conv_tc), as a function of N, of
// the number of accumulator tiles per CTA, of the CTAs per SM, and of the smem-ring handshake conv_tc wraps around it?
//
//   mode 0: one thread issues every MMA back to back, one commit at the end                      (pure pipe rate)
//   mode 1: ring handshake: "producer" thread waits empty -> arrives full (no TMA), MMA thread waits full -> 4*MT
//           MMAs -> commit(empty)                                                                 (conv_tc control flow)
//   cg 2  : same as mode 0 with tcgen05.mma.cta_group::2 on a 2-CTA cluster (M = 256, each CTA holds half of B)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_probe tools/mma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// lean issue: descriptors as {lo, hi} 32-bit halves (hi is constant per layout, lo advances by 2 per 16-element K step)
__device__ __forceinline__ void umma_lohi(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
               ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}

struct P { int N, MT, iters, mode, stages, tmem_cols, kblk; };

template <int CG>
__global__ void __launch_bounds__(256, 1) probe(const P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int a_bytes = p.MT * 128 * 128;                                  // MT tiles of 128 rows x 64 bf16
  const int b_rows = CG == 2 ? p.N / 2 : p.N;
  const int stage_bytes = a_bytes + b_rows * 128;
  // zero the operand ring (NaN-free, data independent timing)
  for (int i = threadIdx.x * 16; i < p.stages * stage_bytes; i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sbase + i), "r"(0));
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < 8; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(&done_bar, (p.mode >= 3 && p.mode <= 5) ? p.MT : 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = make_idesc(CG == 2 ? 256 : 128, p.N);
  const int ksteps = p.kblk / 16;

  if (p.mode == 6) {
    // lean issue + the conv_tc ring handshake: warp 2 = "producer" (all lanes wait empty, elected lane arrives full),
    // warp 1 = MMA warp (all lanes wait full, elected lane issues 4*MT MMAs and commits empty); other warps spin on done_bar
    const int w = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    if (w == 1 && rank == 0) {
      const bool leader = elect_one();
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      int stage = 0; uint32_t phase = 0; uint32_t accum = 0;
      for (int it = 0; it < p.iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_src = sbase + stage * stage_bytes, b_src = a_src + a_bytes;
        const uint32_t alo = ((a_src >> 4) & 0x3FFFu) | (1u << 16), blo = ((b_src >> 4) & 0x3FFFu) | (1u << 16);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_lohi(tmem_base, alo + 2 * k, blo + 2 * k, hi, idesc, k == 0 ? accum : 1u);
            if (p.MT == 2) umma_lohi(tmem_base + p.N, alo + 1024 + 2 * k, blo + 2 * k, hi, idesc, k == 0 ? accum : 1u);
          }
          commit<1>(&empty_bar[stage]);
        }
        accum = 1;
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (leader) commit<1>(&done_bar);
      __syncwarp();
    } else if (w == 2 && rank == 0) {
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < p.iters; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (leader) mbar_arrive(&full_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      __syncwarp();
    }
  } else if (p.mode == 5) {
    // as mode 4 but both operands MN-major (wgrad_tc): 128 rows = 2 blocks of 64 channels, K = pixels, 32-pixel stage
    const int w = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    if (w < p.MT && rank == 0) {
      const uint32_t blk = 32 * 128;                                // one 64-channel block of 32 pixels
      const uint32_t a_src = sbase + w * (2 * blk), b_src = sbase + 8 * blk;
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t alo = ((a_src >> 4) & 0x3FFFu) | ((blk >> 4) << 16), blo = ((b_src >> 4) & 0x3FFFu) | ((blk >> 4) << 16);
      const uint32_t d = tmem_base + w * p.N;
      const uint32_t idesc_mn = idesc | (1u << 15) | (1u << 16);
      const bool leader = elect_one();
      if (leader) {
        umma_lohi(d, alo, blo, hi, idesc_mn, 0u);
        umma_lohi(d, alo + 128, blo + 128, hi, idesc_mn, 1u);
      }
      for (int it = 1; it < 2 * p.iters; ++it) {
        if (leader) {
          umma_lohi(d, alo, blo, hi, idesc_mn, 1u);
          umma_lohi(d, alo + 128, blo + 128, hi, idesc_mn, 1u);
        }
      }
      if (leader) commit<1>(&done_bar);
      __syncwarp();
    }
  } else if (p.mode == 4) {
    // warp-uniform issue loop: every lane runs the loop, one elected lane issues; 32-bit descriptor arithmetic
    const int w = __shfl_sync(0xffffffffu, (int)threadIdx.x >> 5, 0);
    if (w < p.MT && rank == 0) {
      const uint32_t a_src = sbase + w * (128 * 128), b_src = sbase + a_bytes;
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t alo = ((a_src >> 4) & 0x3FFFu) | (1u << 16), blo = ((b_src >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t d = tmem_base + w * p.N;
      const bool leader = elect_one();
      if (leader) {
        umma_lohi(d, alo, blo, hi, idesc, 0u);
        umma_lohi(d, alo + 2, blo + 2, hi, idesc, 1u);
        umma_lohi(d, alo + 4, blo + 4, hi, idesc, 1u);
        umma_lohi(d, alo + 6, blo + 6, hi, idesc, 1u);
      }
      for (int it = 1; it < p.iters; ++it) {
        if (leader) {
          umma_lohi(d, alo, blo, hi, idesc, 1u);
          umma_lohi(d, alo + 2, blo + 2, hi, idesc, 1u);
          umma_lohi(d, alo + 4, blo + 4, hi, idesc, 1u);
          umma_lohi(d, alo + 6, blo + 6, hi, idesc, 1u);
        }
      }
      if (leader) commit<1>(&done_bar);
      __syncwarp();
    }
  } else if (p.mode == 3) {
    // MT issuing threads (one per warp), each with its own accumulator tile and A tile, sharing B
    if (warp < p.MT && lane == 0 && rank == 0) {
      uint32_t accum = 0;
      const uint32_t a_src = sbase + warp * (128 * 128), b_src = sbase + a_bytes;
      for (int it = 0; it < p.iters; ++it)
        for (int k = 0; k < ksteps; ++k) {
          umma<CG>(tmem_base + warp * p.N, make_sdesc(a_src + k * 32, 16, 1024, 2), make_sdesc(b_src + k * 32, 16, 1024, 2), idesc, accum);
          accum = 1;
        }
      commit<CG>(&done_bar);
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    int stage = 0; uint32_t phase = 0; uint32_t accum = 0;
    for (int it = 0; it < p.iters; ++it) {
      if (p.mode == 1) { mbar_wait(&full_bar[stage], phase); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      const uint32_t a_src = sbase + stage * stage_bytes, b_src = a_src + a_bytes;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t bd = make_sdesc(b_src + k * 32, 16, 1024, 2);
        for (int m = 0; m < p.MT; ++m) {
          const uint64_t ad = make_sdesc(a_src + m * (128 * 128) + k * 32, 16, 1024, 2);
          umma<CG>(tmem_base + m * p.N, ad, bd, idesc, accum);
        }
        accum = 1;
      }
      if (p.mode == 1) commit<CG>(&empty_bar[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
    commit<CG>(&done_bar);
  } else if (warp == 2 && lane == 0 && p.mode == 1 && rank == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      mbar_arrive(&full_bar[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1u; }
    }
  }
  __syncwarp();
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

static int pow2_ceil(int v) { int q = 32; while (q < v) q *= 2; return q; }

template <int CG>
static float run(P p, int cps, int smem_pad) {
  p.tmem_cols = pow2_ceil(p.MT * p.N);
  const int b_rows = CG == 2 ? p.N / 2 : p.N;
  int smem = p.stages * (p.MT * 128 * 128 + b_rows * 128) + 1024;
  // occupy 1/cps of the SM's shared memory so that exactly cps CTAs are co-resident
  const int want = (220 * 1024) / cps - 2048;
  if (smem < want && smem_pad) smem = want;
  if (smem > 227 * 1024) return -1.f;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148 * cps;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (CG == 1) {
      probe<1><<<grid, 256, smem>>>(p);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, probe<2>, p);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = (double)grid * p.iters * (p.kblk / 16) * p.MT * 2.0 * 128 * p.N * 16;
  return (float)(flops / best / 1e9);   // TFLOP/s
}

int main() {
  printf("| cg | mode | N | MT | CTAs/SM | stages | TFLOP/s (148 SMs) |\n|---|---|---|---|---|---|---|\n");
  const int Ns[] = {64, 128, 256};
  for (int cg = 1; cg <= 2; ++cg)
    for (int mode = 0; mode <= (cg == 1 ? 1 : 0); ++mode)
      for (int N : Ns)
        for (int MT = 1; MT <= 8; MT *= 2)
          for (int cps = 1; cps <= 4; cps *= 2) {
            if (MT * N > 512) continue;
            if (cps * pow2_ceil(MT * N) > 512) continue;
            if (cps * MT * N < 256 && !(cps == 1 && MT == 1)) continue;      // the interesting region: >= 256 columns busy
            P p; p.N = N; p.MT = MT; p.iters = 2000; p.mode = mode; p.stages = 2; p.kblk = 64;
            if (2 * (MT * 16384 + N * 128) + 1024 > (220 * 1024) / cps - 2048) p.stages = 1;   // keep the ring inside the smem share
            const float tf = cg == 1 ? run<1>(p, cps, 1) : run<2>(p, cps, 1);
            printf("| %d | %d | %d | %d | %d | %d | %.0f |%s\n", cg, mode, N, MT, cps, p.stages, tf, "");
            fflush(stdout);
          }
  // several issuing threads in ONE CTA (mode 3): does the per-CTA ordering go away?
  for (int N : Ns)
    for (int MT = 1; MT <= 8; MT *= 2) {
      if (MT * N > 512) continue;
      P p; p.N = N; p.MT = MT; p.iters = 2000; p.mode = 3; p.stages = 1; p.kblk = 64;
      printf("| 1 | 3 | %d | %d | 1 | 1 | %.0f | %d issuing warps\n", N, MT, run<1>(p, 1, 1), MT);
    }
  for (int N : Ns)
    for (int MT = 1; MT <= 4; MT *= 2) {
      if (MT * N > 512) continue;
      P p; p.N = N; p.MT = MT; p.iters = 2000; p.mode = 4; p.stages = 1; p.kblk = 64;
      printf("| 1 | 4 | %d | %d | 1 | 1 | %.0f | %d issuing warps, lean uniform issue\n", N, MT, run<1>(p, 1, 1), MT);
    }
  for (int N : {128, 256})
    for (int MT = 1; MT <= 2; MT *= 2) {
      P p; p.N = N; p.MT = MT; p.iters = 2000; p.mode = 5; p.stages = 2; p.kblk = 64;
      printf("| 1 | 5 | %d | %d | 1 | 1 | %.0f | %d issuing warps, lean issue, MN-major operands\n", N, MT, run<1>(p, 1, 1), MT);
    }
  for (int N : {128, 256})
    for (int MT = 1; MT <= 2; MT *= 2)
      for (int st = 1; st <= 4; ++st) {
        if (MT * N > 512) continue;
        P p; p.N = N; p.MT = MT; p.iters = 2000; p.mode = 6; p.stages = st; p.kblk = 64;
        printf("| 1 | 6 | %d | %d | 1 | %d | %.0f | lean issue + ring handshake\n", N, MT, st, run<1>(p, 1, 1));
      }
  // stage-count sensitivity of the handshake (N=128, MT=2: the vdis main.1 configuration)
  for (int st = 1; st <= 4; ++st) {
    P p; p.N = 128; p.MT = 2; p.iters = 2000; p.mode = 1; p.stages = st; p.kblk = 64;
    printf("| 1 | 1 | 128 | 2 | 1 | %d | %.0f |\n", st, run<1>(p, 1, 1));
  }
  return 0;
}
