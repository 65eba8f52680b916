// screen_sm100.cu — tensor-core screening pass of the nearest-code search (sm_100a, tcgen05).
//
// Reference: quantize.py:45-50 computes d[n,k] = ||z_n||^2 + ||e_k||^2 - 2 z_n.e_k for all N x K
// pairs in FP32 and takes the row argmin.  argmin_k d[n,k] == argmax_k s[n,k] with
//     s[n,k] = z_n.e_k - 0.5*||e_k||^2                     (||z_n||^2 is constant per row)
// This kernel evaluates s on the 5th-gen tensor cores with BF16 operands / FP32 accumulation and
// keeps, per row, every code whose score is within `row_margin[n]` of the row maximum (the margin
// bounds the BF16 rounding noise).  The survivors are re-scored in FP32 by ccvsq_rescore, so the
// final index equals the FP32 argmin whenever the FP32 winner is inside the margin.
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised):
//   warp 0      TMA producer: A tile (128 latents x D, BF16, stationary for the whole code sweep)
//               and a NST-deep ring of B stages (BN codes x 64 dims), 128B-swizzled, mbarrier-signalled
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16; accumulators
//               double-buffered in TMEM (2 x BN columns) so the epilogue of code tile j overlaps the
//               MMAs of tile j+1
//   warp 2      TMEM allocator
//   warps 4-7   epilogue group 0 (even code tiles): tcgen05.ld 32 columns at a time, one latent row
//   warps 8-11  epilogue group 1 (odd code tiles)   per thread, bias add, running max, candidate ring
// After the last code tile of a row tile the two groups' candidate rings are merged, filtered by
// (row max - margin), sorted by (score desc, index asc) and written out.
#include <cuda.h>
#include <float.h>
#include "common.cuh"

namespace ccvsq {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, 128B-swizzled shared-memory matrix descriptor (sm_100 format, version 1):
// 8-row x 128-byte swizzle atoms, stride between 8-row groups (SBO) = 1024 B, LBO unused.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address  [0,14)
  d |= (uint64_t)0 << 16;                            // LBO            [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO            [32,46)
  d |= (uint64_t)1 << 46;                            // version = 1    [46,48)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B   [61,64)
  return d;
}
// kind::f16 instruction descriptor: FP32 accum, BF16 x BF16, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;            // latent rows per CTA tile (UMMA M)
constexpr int BK = 64;             // dims per smem block (128 bytes of BF16 = one swizzle atom row)
constexpr int LCAP = 6;                 // candidate-list entries per (row, epilogue group)
constexpr uint32_t SC_STRIDE = 128 * 16;   // entry e of a row: 4 raw scores at sc_base + e*SC_STRIDE ...
constexpr uint32_t CO_STRIDE = 128 * 4;    // ... and the first code of the group at co_base + e*CO_STRIDE
constexpr uint32_t LIST_BYTES = LCAP * (SC_STRIDE + CO_STRIDE);   // per epilogue group
constexpr int SCREEN_THREADS = 384;
constexpr int A_BLOCK_BYTES = BM * BK * 2;   // 16 KiB

struct ScreenSmem {
  // byte offsets inside the 1024-aligned dynamic shared memory
  uint32_t a, b, ring, bias, bars, total;
};
__host__ __device__ inline ScreenSmem screen_smem_layout(int dblk, int BN, int nst) {
  ScreenSmem s;
  uint32_t off = 0;
  s.a = off;    off += (uint32_t)dblk * A_BLOCK_BYTES;
  s.b = off;    off += (uint32_t)nst * BN * BK * 2;
  s.ring = off; off += 2u * LIST_BYTES;             // [group]{ [slot][row] float4 | [slot][row] u32 }
  s.bias = off; off += 2u * 2u * BN * 4;            // [group][parity][BN]
  s.bars = off; off += 512;
  s.total = off;
  return s;
}

// Candidate list maintenance (rare path).  Drops entries whose best score fell below `thr`; if
// more than LCAP-2 survive, the entries with the lowest best score go too and the best dropped
// score is remembered, so the row is flagged only if a dropped code could still be inside the
// final margin.
__device__ __noinline__ void list_compact(uint32_t sc_base, uint32_t co_base, uint32_t& psc, uint32_t& pco,
                                          float thr, float& dropped_max) {
  const uint32_t n = (pco - co_base) / CO_STRIDE;
  uint32_t w = 0;
  for (uint32_t e = 0; e < n; ++e) {
    float a, b, c, d;
    uint32_t code;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(sc_base + e * SC_STRIDE));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(co_base + e * CO_STRIDE));
    if (fmaxf(fmaxf(a, b), fmaxf(c, d)) >= thr) {
      if (w != e) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sc_base + w * SC_STRIDE), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(co_base + w * CO_STRIDE), "r"(code) : "memory");
      }
      ++w;
    }
  }
  while (w > LCAP - 2) {
    float lo = INFINITY;              // victim: lowest best score, highest code among equals
    uint32_t lo_code = 0, lo_e = 0;
    for (uint32_t e = 0; e < w; ++e) {
      float a, b, c, d;
      uint32_t code;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(sc_base + e * SC_STRIDE));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(co_base + e * CO_STRIDE));
      const float mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
      if (mx < lo || (mx == lo && code > lo_code)) { lo = mx; lo_code = code; lo_e = e; }
    }
    dropped_max = fmaxf(dropped_max, lo);
    --w;
    for (uint32_t e = lo_e; e < w; ++e) {   // close the hole, keeping entries in increasing code order
      float a, b, c, d;
      uint32_t code;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(sc_base + (e + 1) * SC_STRIDE));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(co_base + (e + 1) * CO_STRIDE));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sc_base + e * SC_STRIDE), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(co_base + e * CO_STRIDE), "r"(code) : "memory");
    }
  }
  psc = sc_base + w * SC_STRIDE;
  pco = co_base + w * CO_STRIDE;
}

// TRACE: diagnostic instantiation (ccvsq_screen_trace) — CTA 0 records (clock64, event) pairs per role.
#define CCVSQ_TRACE_EVENT(role, code)                                                            \
  do {                                                                                           \
    if (TRACE && blockIdx.x == 0 && trace_n[role] < 4000) {                                      \
      trace[(role) * 4000 + trace_n[role]] = (clock64() << 8) | (long long)(code);               \
      ++trace_n[role];                                                                           \
    }                                                                                            \
  } while (0)

template <int BN, int NST, bool TRACE>
__global__ void __launch_bounds__(SCREEN_THREADS, 1)
screen_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const float* __restrict__ bias, const float* __restrict__ row_margin, int64_t N,
              int num_row_tiles, int K, int K_pad, int dblk, int n_cand,
              int32_t* __restrict__ cand_idx, float* __restrict__ cand_score,
              uint8_t* __restrict__ flags, float* __restrict__ dbg_scores, long long* __restrict__ trace) {
  int trace_n[4] = {0, 0, 0, 0};
  (void)trace_n;
  extern __shared__ __align__(1024) uint8_t smem[];
  const ScreenSmem lay = screen_smem_layout(dblk, BN, NST);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n_tiles = K_pad / BN;
  constexpr uint32_t B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  // barrier addresses
  const uint32_t bar0 = smem_base + lay.bars;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NST + s); };
  auto tmem_full = [&](int b) { return bar0 + 8u * (2 * NST + b); };
  auto tmem_empty = [&](int b) { return bar0 + 8u * (2 * NST + 2 + b); };
  auto bias_rdy = [&](int g) { return bar0 + 8u * (2 * NST + 4 + g); };
  // the stationary A tile is handed over per 64-dim block, so the next row tile's blocks stream in
  // while the last code tile of the current row tile is still being multiplied
  auto a_full = [&](int kb) { return bar0 + 8u * (2 * NST + 6 + kb); };
  auto a_empty = [&](int kb) { return bar0 + 8u * (2 * NST + 14 + kb); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + lay.bars + 8u * (2 * NST + 22));

  if ((smem_base & 1023u) != 0) __trap();   // SWIZZLE_128B needs 1024-byte aligned tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int kb = 0; kb < 8; ++kb) { mbar_init(a_full(kb), 1); mbar_init(a_empty(kb), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 128);
      mbar_init(bias_rdy(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int tile = blockIdx.x; tile < num_row_tiles; tile += gridDim.x) {
        for (int j = 0; j < num_n_tiles; ++j) {
          for (int kb = 0; kb < dblk; ++kb) {
            if (j == 0) {   // A block kb of this row tile, as soon as the previous tile's MMAs released it
              mbar_wait(a_empty(kb), a_phase ^ 1);
              mbar_arrive_expect_tx(a_full(kb), A_BLOCK_BYTES);
              tma_load_2d(smem_base + lay.a + kb * A_BLOCK_BYTES, &map_a, a_full(kb), kb * BK, tile * BM);
            }
            mbar_wait(empty_bar(stage), phase ^ 1);
            CCVSQ_TRACE_EVENT(0, 1);                       // B stage issued
            mbar_arrive_expect_tx(full_bar(stage), B_STAGE_BYTES);
            tma_load_2d(smem_base + lay.b + stage * B_STAGE_BYTES, &map_b, full_bar(stage), kb * BK,
                        j * BN);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
        a_phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      uint32_t uses[2] = {0, 0};
      for (int tile = blockIdx.x; tile < num_row_tiles; tile += gridDim.x) {
        for (int j = 0; j < num_n_tiles; ++j) {
          const int b = j & 1;
          CCVSQ_TRACE_EVENT(1, 1);                         // start waiting for the accumulator buffer
          mbar_wait(tmem_empty(b), (uses[b] & 1) ^ 1);
          CCVSQ_TRACE_EVENT(1, 2);                         // accumulator buffer free
          ++uses[b];
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
          for (int kb = 0; kb < dblk; ++kb) {
            if (j == 0) mbar_wait(a_full(kb), a_phase);
            mbar_wait(full_bar(stage), phase);
            CCVSQ_TRACE_EVENT(1, 3);                       // operands of this k-block landed
            tc_fence_after();
            const uint32_t a_addr = smem_base + lay.a + kb * A_BLOCK_BYTES;
            const uint32_t b_addr = smem_base + lay.b + stage * B_STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t da = make_sw128_desc(a_addr + k * 32);
              const uint64_t db = make_sw128_desc(b_addr + k * 32);
              umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));      // smem stage reusable once these MMAs retire
            if (j == num_n_tiles - 1) umma_commit(a_empty(kb));   // last reader of A block kb
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
          umma_commit(tmem_full(b));            // accumulator tile j complete
          CCVSQ_TRACE_EVENT(1, 4);                         // tile issued
        }
        a_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue groups ===========================
    // The two groups are fully decoupled: group g owns the code tiles j == g (mod 2) and the TMEM
    // buffer g, keeps its own running maximum and candidate list and writes its own half of the
    // output; ccvsq_rescore merges the two halves.
    const int g = (warp - 4) >> 2;             // 0: even code tiles, 1: odd code tiles
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row_in_tile = q * 32 + lane;
    const int tg = threadIdx.x - 128 - g * 128;   // 0..127 inside the group
    float* bias_s = reinterpret_cast<float*>(smem + lay.bias) + g * 2 * BN;
    const uint32_t sc_base = smem_base + lay.ring + (uint32_t)g * LIST_BYTES + (uint32_t)row_in_tile * 16u;
    const uint32_t co_base = smem_base + lay.ring + (uint32_t)g * LIST_BYTES + LCAP * SC_STRIDE + (uint32_t)row_in_tile * 4u;
    const uint32_t co_limit = co_base + (LCAP - 2) * CO_STRIDE;   // appending 2 groups needs pco <= co_limit
    uint32_t full_phase = 0;
    uint32_t it = 0;                           // tiles processed by this group (bias buffer parity)
    const bool has_tiles = g < num_n_tiles;

    // bias of the next tile this group will process, prefetched one tile ahead (also across row tiles)
    float nb[BN / 128];
    if (has_tiles && (int)blockIdx.x < num_row_tiles) {
#pragma unroll
      for (int u = 0; u < BN / 128; ++u) nb[u] = __ldg(bias + (size_t)g * BN + u * 128 + tg);
    }

    for (int tile = blockIdx.x; tile < num_row_tiles; tile += gridDim.x) {
      const int64_t row = (int64_t)tile * BM + row_in_tile;
      const float margin = __ldg(row_margin + row);
      float runmax = -FLT_MAX;
      float dropped_max = -FLT_MAX;
      uint32_t psc = sc_base, pco = co_base;    // next free list entry

      for (int j = g; j < num_n_tiles; j += 2, ++it) {
        float* bs = bias_s + (it & 1) * BN;
#pragma unroll
        for (int u = 0; u < BN / 128; ++u) bs[u * 128 + tg] = nb[u];
        mbar_arrive(bias_rdy(g));               // (waited on below, after the accumulator wait)
        {
          int jn = j + 2;                       // next tile of this group: same row tile, or the
          if (jn >= num_n_tiles) jn = g;        // first one of the next row tile
          if (j + 2 < num_n_tiles || tile + (int)gridDim.x < num_row_tiles) {
#pragma unroll
            for (int u = 0; u < BN / 128; ++u) nb[u] = __ldg(bias + (size_t)jn * BN + u * 128 + tg);
          }
        }
        if (tg == 0) CCVSQ_TRACE_EVENT(2 + g, 1);          // waiting for accumulator tile
        mbar_wait(tmem_full(g), full_phase);
        full_phase ^= 1;
        tc_fence_after();
        mbar_wait(bias_rdy(g), it & 1);
        if (tg == 0) CCVSQ_TRACE_EVENT(2 + g, 2);          // tile available         // all 128 threads of the group stored their bias slice
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN);
        const int col0 = j * BN;

        // One 32-column chunk of this thread's row.  Fast path: bias add + max tree (FMNMX3).  A
        // chunk can only contribute candidates if its maximum reaches (running max - margin); then
        // the threshold is refreshed and every group of 4 columns whose maximum reaches it is
        // appended (its 4 raw scores + first code) with predicated, branch-free stores.
        auto process = [&](uint32_t (&ra)[32], const int cbase) {
          float v[32];
          const float4* b4 = reinterpret_cast<const float4*>(bs + cbase);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            v[4 * i + 0] = __uint_as_float(ra[4 * i + 0]) + bb.x;
            v[4 * i + 1] = __uint_as_float(ra[4 * i + 1]) + bb.y;
            v[4 * i + 2] = __uint_as_float(ra[4 * i + 2]) + bb.z;
            v[4 * i + 3] = __uint_as_float(ra[4 * i + 3]) + bb.w;
          }
          if (dbg_scores) {   // diagnostic dump of the raw score tile (ccvsq_screen_dump only)
#pragma unroll
            for (int i = 0; i < 32; ++i) dbg_scores[row * K_pad + col0 + cbase + i] = v[i];
          }
          float m4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            m4[i] = fmaxf(fmaxf(v[4 * i], v[4 * i + 1]), fmaxf(v[4 * i + 2], v[4 * i + 3]));
          const float m = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])),
                                fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
          if (m >= runmax - margin) {
            // a maximum that beats the old one by more than the margin makes every listed entry stale
            if (m > runmax + margin) { psc = sc_base; pco = co_base; }
            runmax = fmaxf(runmax, m);
            const float thr = runmax - margin;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if ((i & 1) == 0) {
                if (pco > co_limit) list_compact(sc_base, co_base, psc, pco, thr, dropped_max);
              }
              asm volatile(
                  "{\n\t"
                  ".reg .pred p;\n\t"
                  "setp.ge.f32 p, %2, %3;\n\t"
                  "@p st.shared.v4.f32 [%0], {%4, %5, %6, %7};\n\t"
                  "@p st.shared.u32 [%1], %8;\n\t"
                  "@p add.u32 %0, %0, %9;\n\t"
                  "@p add.u32 %1, %1, %10;\n\t"
                  "}"
                  : "+r"(psc), "+r"(pco)
                  : "f"(m4[i]), "f"(thr), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]),
                    "r"((uint32_t)(col0 + cbase + 4 * i)), "n"(SC_STRIDE), "n"(CO_STRIDE)
                  : "memory");
            }
          }
        };

        // two 32-column chunks in flight: the tcgen05.ld of chunk c+1 overlaps the arithmetic on chunk c
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr0, ra);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c += 2) {
          tmem_ld_wait();
          tmem_ld32(taddr0 + (c + 1) * 32, rb);
          process(ra, c * 32);
          tmem_ld_wait();
          if (c + 2 < BN / 32) tmem_ld32(taddr0 + (c + 2) * 32, ra);
          process(rb, (c + 1) * 32);
        }
        tc_fence_before();
        mbar_arrive(tmem_empty(g));
        if (tg == 0) CCVSQ_TRACE_EVENT(2 + g, 3);          // tile consumed
      }

      // ---- this group's candidates for the row: every listed code with score >= runmax - margin,
      //      sorted by (score desc, code asc), at most n_cand of them.  The list is in increasing
      //      code order, so a strict '>' scan keeps the lowest code among equal scores.
      if (row < N) {
        const int64_t obase = (row * 2 + g) * n_cand;
        const float thr = runmax - margin;
        const uint32_t n = (pco - co_base) / CO_STRIDE;
        uint32_t within = 0;
        float best_s = -INFINITY;
        int best_i = -1;
        for (uint32_t e = 0; e < n; ++e) {
          float sc[4];
          uint32_t code;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3]) : "r"(sc_base + e * SC_STRIDE));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(co_base + e * CO_STRIDE));
#pragma unroll
          for (int u = 0; u < 4; ++u) {      // branch-free (padding codes carry -inf and never pass)
            within += (sc[u] >= thr) ? 1u : 0u;
            const bool better = sc[u] > best_s;
            best_s = better ? sc[u] : best_s;
            best_i = better ? (int)code + u : best_i;
          }
        }
        if (within <= 1 && n_cand == 4) {   // the overwhelmingly common case: two 16-byte stores
          *reinterpret_cast<int4*>(cand_idx + obase) = make_int4(within ? best_i : -1, -1, -1, -1);
          *reinterpret_cast<float4*>(cand_score + obase) = make_float4(best_s, -INFINITY, -INFINITY, -INFINITY);
        } else {
          int written = 0;
          if (within >= 1) {
            cand_idx[obase] = best_i;
            cand_score[obase] = best_s;
            written = 1;
          }
          float prev_s = best_s;
          int prev_i = best_i;
          for (int c = 1; c < n_cand && within > 1; ++c) {   // near-ties: repeated selection
            float bs_ = -INFINITY;
            int bi = 0x7fffffff;
            for (uint32_t e = 0; e < n; ++e) {
              float sc[4];
              uint32_t code;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3]) : "r"(sc_base + e * SC_STRIDE));
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(code) : "r"(co_base + e * CO_STRIDE));
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = (int)code + u;
                const float x = sc[u];
                const bool after_prev = (x < prev_s) || (x == prev_s && i > prev_i);
                const bool better = (x > bs_) || (x == bs_ && i < bi);
                if (x >= thr && after_prev && better) { bs_ = x; bi = i; }
              }
            }
            if (bi == 0x7fffffff) break;
            cand_idx[obase + c] = bi;
            cand_score[obase + c] = bs_;
            prev_s = bs_;
            prev_i = bi;
            ++written;
          }
          for (int c = written; c < n_cand; ++c) {
            cand_idx[obase + c] = -1;
            cand_score[obase + c] = -INFINITY;
          }
        }
        if (tg == 0) CCVSQ_TRACE_EVENT(2 + g, 4);          // row tile finalised
        flags[row * 2 + g] = (uint8_t)((within > (uint32_t)n_cand ? 1 : 0) | (dropped_max >= thr ? 2 : 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CCVSQ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CCVSQ_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, CCVSQ_CUDA_ERROR,
                  "cuTensorMapEncodeTiled entry point not available (driver too old?)");
    cached = (EncodeTiledFn)fn;
  }
  *out = cached;
  return CCVSQ_OK;
}

// row-major [rows, D] BF16, box = 64 columns x box_rows rows, 128-byte swizzle
static int make_map(EncodeTiledFn enc, CUtensorMap* map, const void* base, int64_t rows, int D,
                    int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CCVSQ_REQUIRE(r == CUDA_SUCCESS, CCVSQ_CUDA_ERROR, "cuTensorMapEncodeTiled failed with CUresult %d",
                (int)r);
  return CCVSQ_OK;
}

template <int BN, int NST, bool TRACE = false>
static int launch_screen(const CUtensorMap& ma, const CUtensorMap& mb, const float* bias,
                         const float* row_margin, int64_t N, int tiles, int K, int K_pad, int dblk,
                         int n_cand, int32_t* cand_idx, float* cand_score, uint8_t* flags,
                         float* dbg_scores, cudaStream_t st, long long* trace = nullptr) {
  const ScreenSmem lay = screen_smem_layout(dblk, BN, NST);
  const size_t smem = lay.total;
  CCVSQ_REQUIRE(smem <= 227 * 1024, CCVSQ_UNSUPPORTED, "screen: %zu bytes of shared memory needed", smem);
  auto kern = screen_kernel<BN, NST, TRACE>;
  if (int rc = enable_smem(kern, smem)) return rc;
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  kern<<<grid, SCREEN_THREADS, smem, st>>>(ma, mb, bias, row_margin, N, tiles, K, K_pad, dblk, n_cand,
                                         cand_idx, cand_score, flags, dbg_scores, trace);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

}  // namespace ccvsq

using namespace ccvsq;

static int screen_impl(const void* z_bf16, const float* row_margin, const void* E_bf16, const float* bias,
                       int64_t N, int K, int D, int n_cand, int32_t* cand_idx, float* cand_score,
                       uint8_t* flags, float* dbg_scores, void* stream, long long* trace = nullptr) {
  CCVSQ_REQUIRE(z_bf16 && row_margin && E_bf16 && bias && cand_idx && cand_score && flags,
                CCVSQ_NULL_POINTER, "screen: null pointer");
  CCVSQ_REQUIRE(N > 0 && K > 0, CCVSQ_BAD_SHAPE, "screen: N=%lld K=%d", (long long)N, K);
  CCVSQ_REQUIRE(D % 64 == 0 && D >= 64 && D <= 512, CCVSQ_UNSUPPORTED,
                "screen: D=%d unsupported by the tensor-core path (need 64 <= D <= 512, D %% 64 == 0)", D);
  CCVSQ_REQUIRE(n_cand >= 1 && n_cand <= CCVSQ_MAX_CAND, CCVSQ_BAD_SHAPE, "screen: n_cand=%d", n_cand);
  CCVSQ_REQUIRE((((uintptr_t)z_bf16 | (uintptr_t)E_bf16 | (uintptr_t)cand_idx | (uintptr_t)cand_score) & 15) == 0,
                CCVSQ_MISALIGNED, "screen: BF16 operands and candidate arrays must be 16-byte aligned");
  const int64_t N_pad = ((N + BM - 1) / BM) * BM;
  const int K_pad = ((K + 255) / 256) * 256;
  const int64_t tiles64 = N_pad / BM;
  CCVSQ_REQUIRE(tiles64 < (1ll << 24), CCVSQ_BAD_SHAPE, "screen: N=%lld too large for one launch",
                (long long)N);
  const int dblk = D / 64;
  EncodeTiledFn enc;
  if (int rc = get_encode_fn(&enc)) return rc;
  CUtensorMap ma, mb;
  if (int rc = make_map(enc, &ma, z_bf16, N_pad, D, BM)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dblk <= 4) {
    if (int rc = make_map(enc, &mb, E_bf16, K_pad, D, 256)) return rc;
    if (trace)
      return launch_screen<256, 4, true>(ma, mb, bias, row_margin, N, (int)tiles64, K, K_pad, dblk, n_cand,
                                         cand_idx, cand_score, flags, dbg_scores, st, trace);
    return launch_screen<256, 4>(ma, mb, bias, row_margin, N, (int)tiles64, K, K_pad, dblk, n_cand,
                                 cand_idx, cand_score, flags, dbg_scores, st);
  }
  if (int rc = make_map(enc, &mb, E_bf16, K_pad, D, 128)) return rc;
  return launch_screen<128, 4>(ma, mb, bias, row_margin, N, (int)tiles64, K, K_pad, dblk, n_cand, cand_idx,
                               cand_score, flags, dbg_scores, st);
}

extern "C" int ccvsq_screen(const void* z_bf16, const float* row_margin, const void* E_bf16,
                            const float* bias, int64_t N, int K, int D, int n_cand, int32_t* cand_idx,
                            float* cand_score, uint8_t* flags, void* stream) {
  return screen_impl(z_bf16, row_margin, E_bf16, bias, N, K, D, n_cand, cand_idx, cand_score, flags, nullptr,
                     stream);
}

extern "C" int ccvsq_screen_dump(const void* z_bf16, const float* row_margin, const void* E_bf16,
                                 const float* bias, int64_t N, int K, int D, int n_cand, int32_t* cand_idx,
                                 float* cand_score, uint8_t* flags, float* scores, void* stream) {
  CCVSQ_REQUIRE(scores, CCVSQ_NULL_POINTER, "screen_dump: scores must be non-null");
  return screen_impl(z_bf16, row_margin, E_bf16, bias, N, K, D, n_cand, cand_idx, cand_score, flags, scores,
                     stream);
}

// Diagnostic: CTA 0 records a per-role event timeline, trace is int64 [4][4000] zero-filled by the
// caller (roles: 0 TMA producer, 1 MMA issuer, 2/3 epilogue groups; value = clock64 << 8 | event).
extern "C" int ccvsq_screen_trace(const void* z_bf16, const float* row_margin, const void* E_bf16,
                                  const float* bias, int64_t N, int K, int D, int n_cand, int32_t* cand_idx,
                                  float* cand_score, uint8_t* flags, long long* trace, void* stream) {
  CCVSQ_REQUIRE(trace, CCVSQ_NULL_POINTER, "screen_trace: trace must be non-null");
  CCVSQ_REQUIRE(D <= 256, CCVSQ_UNSUPPORTED, "screen_trace: only the BN=256 configuration is traced");
  return screen_impl(z_bf16, row_margin, E_bf16, bias, N, K, D, n_cand, cand_idx, cand_score, flags, nullptr,
                     stream, trace);
}
