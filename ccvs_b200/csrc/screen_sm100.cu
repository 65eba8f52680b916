// screen_sm100.cu — tensor-core screening pass of the nearest-code search (sm_100a, tcgen05).
//
// Reference: quantize.py:45-50 computes d[n,k] = ||z_n||^2 + ||e_k||^2 - 2 z_n.e_k for all N x K
// pairs in FP32 and takes the row argmin.  argmin_k d[n,k] == argmax_k s[n,k] with
//     s[n,k] = z_n.e_k - 0.5*||e_k||^2                     (||z_n||^2 is constant per row)
// This kernel evaluates s on the 5th-gen tensor cores with BF16 operands / FP32 accumulation and
// keeps, per row, every code whose score is within a margin of the row maximum (the margin bounds
// the BF16 rounding noise).  Rows with a single survivor are final; the others are queued for FP32
// re-scoring (ccvsq_rescore), so the result equals the FP32 argmin whenever the FP32 winner is
// inside the margin.
//
// Design (one persistent CTA per SM, CTAs paired into 2-CTA clusters, 512 threads, warp-specialised):
//   * A operand (128 latent rows per CTA) lives in TENSOR MEMORY for the whole code sweep.  Loader
//     warps read the FP32 latents straight from the caller's tensor (any ccvsq_layout: the
//     NCHW->rows transpose of quantize.py:40-42 is pure addressing), round to BF16 and write the
//     packed pairs with tcgen05.st.  No packing pass, no BF16 copy of z in HBM, no A traffic in
//     shared memory.  Two A buffers (D <= 256) let the next row tile stream in under the MMAs.
//   * B operand: BN=96 codes x (D+16) dims per tile, TMA-loaded (128B swizzle) into a ring of
//     whole-tile slots.  With tcgen05.mma.cta_group::2 each CTA of the pair holds half the codes of
//     a tile, so L2->SM and shared-memory traffic per FLOP are half those of a single-CTA MMA.
//   * The bias -0.5||e||^2 is folded into the GEMM: the codebook shadow carries 16 extra columns
//     (a 3-term BF16 split hi+mid+lo of the bias, exact to FP32), matched by a constant A block
//     (1,1,1,0,...) in tensor memory.  The accumulator IS the score: the epilogue has no bias loads
//     or adds, only max trees.
//   * Accumulators (128 x 96 FP32) double-buffered in tensor memory; one epilogue warp per TMEM
//     lane quadrant, one latent row per thread over ALL codes -> one candidate list per row.
//   warp 0      TMA producer (both CTAs, each loads its half of every B tile)
//   warp 1      MMA issuer (leader CTA of the pair only)
//   warp 2      TMEM allocator
//   warp 3      L2 prefetcher for the latents (one row tile ahead of the A loaders)
//   warps 4-7   epilogue: tcgen05.ld 32 columns at a time, max tree, candidate list, outputs
//   warps 8-15  A loaders: global FP32 -> BF16 -> tcgen05.st, row norms for the margin
#include <cuda.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ccvsq {

// ------------------------------------------------------------------------------------------------
// kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128;              // latent rows per CTA (TMEM lanes)
// Codes per accumulator tile (UMMA N) is a template parameter:
//   128  wide-tile plan (short code sweeps, the BASELINE K = 1024 shapes): next to concurrent tensor-memory readers a
//        128-column MMA stream keeps 87 % of its rate where a 64-column one keeps 59 % (the per-instruction arbitration
//        loss is the same, the instruction twice as long: profiles/r01_tmem_mma_microbench.txt).  2 accumulators +
//        2 A buffers fill the 512 columns exactly, so the constant bias operand moves to shared memory (one SS-form
//        MMA per tile), and TWO epilogue warpgroups split the columns of every tile (64 each): the accumulator is
//        handed back as soon as both have issued their loads, and each has two tile times for its max trees / lists.
//   96   long sweeps (K >= 2304): three accumulators, one A buffer
//   64   D = 512 / fallback plans
// Candidate store in shared memory ("chunk parking").  A 32-column chunk whose maximum reaches (running max - margin)
// is parked whole: 8 unpredicated 128-bit stores + one header {first code, chunk maximum}.  Every lane of a warp sets
// ~ln(chunks) running-max records per sweep, so SOME lane is in that case for most chunks and the warp pays the
// parking path almost always: it has to be short (about 25 instructions; the per-group lists of round 1 cost ~100
// with prefix sums, predicated stores and compaction).  A new maximum that beats the old one by more than the margin
// makes everything parked so far stale (ring restarts); otherwise the ring overwrites its oldest entry and remembers
// the largest maximum it dropped, so a row is only flagged if a dropped chunk could still hold a code inside the final
// margin.  Entry e of a group: scores [8 x float4][128 rows] (a warp's store of one float4 is 512 contiguous bytes),
// then headers [128 rows] x {code, max}.
constexpr uint32_t ENTRY_BYTES = 8 * BM * 16 + BM * 8;
constexpr uint32_t HDR_OFFSET = 8 * BM * 16;
// entries per row and epilogue group (the same shared memory per row for the one-group and the two-group plans)
// (the plan with the second A buffer in shared memory gives one entry per group back: 64 KiB of staging)
__host__ __device__ constexpr int park_cap(int bn, bool a_smem = false) { return bn == 128 ? (a_smem ? 2 : 3) : 6; }
constexpr int MAX_SLOTS = 16;
// tensor-memory columns (32-bit):
//   BN < 128 : [0, nacc*BN) accumulators | [nacc*BN, +8) bias-extension A block | then abuf_n A buffers of D/2
//   BN = 128 : [0, 2*128) accumulators | abuf_n A buffers of D/2 (the bias-extension A block is in shared memory)
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_ACC = 8;

template <int BN>
struct ScreenCfg {
  static constexpr bool WIDE = BN == 128;
  static constexpr int NEPG = WIDE ? 2 : 1;              // epilogue warpgroups
  static constexpr int EP_WARP0 = 4;                     // warps 0-3: TMA, MMA, allocator, L2 prefetcher
  static constexpr int LD_WARP0 = EP_WARP0 + 4 * NEPG;   // 8 loader warps follow the epilogue warps
  static constexpr int THREADS = (LD_WARP0 + 8) * 32;    // 512 / 640
};
__host__ __device__ constexpr int screen_threads(int bn) { return bn == 128 ? 640 : 512; }

struct ScreenSmem {
  uint32_t slots, astage, list, norm, epst, cblk, bars, total;   // byte offsets inside the 1024-aligned dynamic smem
  uint32_t block_bytes, ext_off, slot_bytes, slot_tx;
  int nslots;
};
__host__ __device__ inline ScreenSmem screen_smem_layout(int dblk, int cg, int bn, bool a_smem = false) {
  ScreenSmem s;
  const uint32_t nepg = bn == 128 ? 2 : 1;
  const uint32_t rows = bn / cg;                    // codes of a tile held by one CTA
  s.block_bytes = rows * 128;                       // one 64-dim block, 128B-swizzled
  s.ext_off = dblk * s.block_bytes;                 // bias extension: [2 K-chunks][rows][16 B]
  s.slot_tx = s.ext_off + rows * 32;
  s.slot_bytes = (s.slot_tx + 1023u) & ~1023u;
  const uint32_t norm_bytes = 2 * 2 * 2 * BM * 4 + 2 * BM * 8;   // [A buffer][loader half][||z||^2, ||z - bf16(z)||^2][row] | [group][row] {live running max, sweep}
  const uint32_t cblk_bytes = bn == 128 ? BM * 32 : 0;
  const uint32_t list_bytes = (uint32_t)park_cap(bn, a_smem) * ENTRY_BYTES;      // per epilogue group
  // second A buffer (BF16, K-major without swizzle: 16-byte chunk kc of row r at kc * BM*16 + r*16) of the plan that
  // alternates the A operand between tensor memory and shared memory
  const uint32_t astage_bytes = a_smem ? (uint32_t)dblk * 64u * 2u * BM : 0u;
  const uint32_t fixed = nepg * list_bytes + norm_bytes + nepg * BM * 16 + cblk_bytes + 512 + astage_bytes;
  int n = (int)((227u * 1024u - fixed) / s.slot_bytes);
  s.nslots = n > MAX_SLOTS ? MAX_SLOTS : n;
  uint32_t off = 0;
  s.slots = off; off += (uint32_t)s.nslots * s.slot_bytes;
  s.astage = off; off += astage_bytes;
  s.list = off;  off += nepg * list_bytes;          // per epilogue group: park_cap entries (see ENTRY_BYTES)
  s.norm = off;  off += norm_bytes;
  s.epst = off;  off += nepg * BM * 16;             // [group][row] {running max, best code, codes inside the margin, entries}: end-of-sweep exchange
  s.cblk = off;  off += cblk_bytes;                 // constant A block (1,1,1,0,...) of the bias step: [2 K-chunks][128 rows][16 B]
  s.bars = off;  off += 512;
  s.total = off;
  return s;
}

struct ScreenOut {
  int64_t* idx;            // [N] final index of unambiguous rows (provisional best for queued rows)
  int32_t* q_count;        // [1] rows queued for FP32 re-scoring
  int32_t* q_rows;         // [N]
  int32_t* q_cand;         // [N, n_cand]
  uint8_t* q_flags;        // [N]
  bool z_stable;           // z was complete before this launch chain began (set by the composite forward, whose first
                           // operations are ordinary launches): the latent loads may start before pdl_wait()
  // diagnostics (all may be null)
  int32_t* dbg_cand;       // [N, n_cand]
  float* dbg_score;        // [N, n_cand]
  uint8_t* dbg_flags;      // [N]
  float* dbg_margin;       // [N]
  float* dbg_scores;       // [N, K_pad]
  long long* trace;        // [gridDim.x, TRACE_SWEEPS, 8] SM clock stamps of the pipeline hand-offs (DBG kernel only)
};
constexpr int TRACE_SWEEPS = 32;

// the parked chunks of one row: one ring per epilogue group (the wide-tile plan keeps two, each over half the columns)
struct RowLists {
  uint32_t base[2];      // shared-memory address of the group's entry 0, already offset to this row's float4 column
  uint32_t hdr[2];       // ... and of its header
  uint32_t n[2];         // valid entries
};

// Fast end-of-sweep scan of one store: how many parked codes are inside the margin, and — valid only when that number
// is exactly one — which code it is.  Per live chunk: the 8 group maxima (FMNMX3 trees), the number of groups that reach
// the threshold, and, if that is one, the four scores of that group again (one more shared-memory load).  Two or more
// qualifying groups are reported as "at least two codes" — the caller takes the general path then.
__device__ __forceinline__ void scan_unique(uint32_t base, uint32_t hdr, uint32_t n, float thr, uint32_t& within,
                                            uint32_t& code_sum) {
#pragma unroll 1
  for (uint32_t e = 0; e < n; ++e) {
    uint32_t code;
    float mx;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(code), "=f"(mx) : "r"(hdr + e * ENTRY_BYTES));
    if (!(mx >= thr)) continue;
    const uint32_t eb = base + e * ENTRY_BYTES;
    uint32_t ng = 0, gi = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float sc[4];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3])
                   : "r"(eb + (uint32_t)i * (BM * 16)));
      const bool p = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), sc[2]), sc[3]) >= thr;
      ng += p ? 1u : 0u;
      gi += p ? (uint32_t)i : 0u;
    }
    if (ng == 1u) {
      float sc[4];
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3])
                   : "r"(eb + gi * (BM * 16)));
      uint32_t w = 0, off = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool p = sc[u] >= thr;
        w += p ? 1u : 0u;
        off += p ? (uint32_t)u : 0u;
      }
      within += w;
      code_sum += w * (code + 4u * gi) + off;
    } else {
      within += 2u;            // (ng >= 2; ng == 0 cannot happen: the chunk maximum reached the threshold)
    }
  }
}

// The same scan for the wide plan, with the liveness test on the REGISTER copy of the chunk maxima (the header is only
// read for a live chunk: the row's first code of that chunk) and the loop over the entries fully unrolled.
template <int CAPN>
__device__ __forceinline__ void scan_unique_reg(uint32_t base, uint32_t hdr, uint32_t n, const float (&pmax)[CAPN], float thr,
                                                uint32_t& within, uint32_t& code_sum) {
#pragma unroll
  for (uint32_t e = 0; e < (uint32_t)CAPN; ++e) {
    if (e < n && pmax[e] >= thr) {
      const uint32_t eb = base + e * ENTRY_BYTES;
      float sc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[i][0]), "=f"(sc[i][1]), "=f"(sc[i][2]), "=f"(sc[i][3])
                     : "r"(eb + (uint32_t)i * (BM * 16)));
      uint32_t code;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(code) : "r"(hdr + e * ENTRY_BYTES));
      uint32_t w = 0, pos = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool p = sc[i][u] >= thr;
          w += p ? 1u : 0u;
          pos += p ? (uint32_t)(4 * i + u) : 0u;
        }
      within += w;
      code_sum += w * code + pos;          // (meaningful only when the total count is exactly one)
    }
  }
}

// General end-of-sweep path of one row (rows with more than one code inside the margin, dropped chunks, diagnostics):
// every parked code with score >= runmax - margin, sorted by (score desc, code asc), at most n_cand of them, goes to the
// re-scoring queue.  ONE pass over the live entries: every qualifying code is counted, compared with the best, and
// inserted into a sorted list of CCVSQ_MAX_CAND (score, code) pairs held in registers (an unrolled compare-exchange chain;
// round 2's first version re-scanned the stores once per candidate slot — five passes at n_cand = 4, which made the
// screen three times slower on the fresh-init distribution where a third of the rows comes here).
// Kept out of line so the per-tile loop stays small in the instruction cache.
template <int NL>
__device__ __noinline__ void finalize_row(const RowLists L2, float runmax, float margin, bool dropped, int n_cand,
                                          int64_t row, const ScreenOut out) {
  const float thr = runmax - margin;
  uint32_t within = 0;
  float cs[CCVSQ_MAX_CAND];
  int ci[CCVSQ_MAX_CAND];
#pragma unroll
  for (int t = 0; t < CCVSQ_MAX_CAND; ++t) { cs[t] = -INFINITY; ci[t] = 0x7fffffff; }
#pragma unroll
  for (int l = 0; l < NL; ++l) {
#pragma unroll 1
    for (uint32_t e = 0; e < L2.n[l]; ++e) {
      uint32_t code;
      float mx;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(code), "=f"(mx) : "r"(L2.hdr[l] + e * ENTRY_BYTES));
      if (!(mx >= thr)) continue;
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        float sc[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3])
                     : "r"(L2.base[l] + e * ENTRY_BYTES + (uint32_t)i * (BM * 16)));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!(sc[u] >= thr)) continue;           // (padding codes carry -3e38 and never pass)
          ++within;
          float vx = sc[u];
          int vc = (int)code + 4 * i + u;
#pragma unroll
          for (int t = 0; t < CCVSQ_MAX_CAND; ++t) {   // insertion: (score desc, code asc)
            const bool before = (vx > cs[t]) || (vx == cs[t] && vc < ci[t]);
            const float ts = cs[t];
            const int tc = ci[t];
            cs[t] = before ? vx : ts;
            ci[t] = before ? vc : tc;
            vx = before ? ts : vx;
            vc = before ? tc : vc;
          }
        }
      }
    }
  }
  const int best_i = ci[0] == 0x7fffffff ? -1 : ci[0];
  const uint8_t flag = (uint8_t)((within > (uint32_t)n_cand ? 1 : 0) | (dropped ? 2 : 0));
  const bool final_row = within == 1 && flag == 0;
  // (a row of NaN / Inf latents parks nothing: every returned index stays inside [0, K) like torch.argmin's)
  if (out.idx) out.idx[row] = best_i < 0 ? 0 : best_i;
  if (final_row && !out.dbg_cand) return;

  int slot = -1;
  if (!final_row && out.q_count) {
    slot = atomicAdd(out.q_count, 1);
    out.q_rows[slot] = (int32_t)row;
    out.q_flags[slot] = flag;
  }
  if (out.dbg_cand) {
    out.dbg_flags[row] = flag;
    out.dbg_margin[row] = margin;
  }
#pragma unroll
  for (int c = 0; c < CCVSQ_MAX_CAND; ++c) {
    if (c < n_cand) {
      const int code_out = ci[c] == 0x7fffffff ? -1 : ci[c];
      if (slot >= 0) out.q_cand[(int64_t)slot * n_cand + c] = code_out;
      if (out.dbg_cand) {
        out.dbg_cand[row * n_cand + c] = code_out;
        out.dbg_score[row * n_cand + c] = cs[c];
      }
    }
  }
}

// 32 consecutive dims of one latent row -> registers (S == 1: contiguous row, else stride S)
__device__ __forceinline__ void load_chunk(float (&v)[32], const float* __restrict__ p, int64_t S, bool valid) {
  if (!valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    return;
  }
  if (S == 1) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(p4 + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else if (S == 256) {        // the BASELINE layouts (16x16 and 8x8 latent frames): the 32 offsets become immediates
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __ldg(p + i * 256);
  } else if (S == 64) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __ldg(p + i * 64);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __ldg(p + (int64_t)i * S);
  }
}

// 32 FP32 values -> 16 packed BF16 pairs (RN), accumulating ||v||^2 and the squared rounding error ||v - bf16(v)||^2
__device__ __forceinline__ void pack_chunk(const float (&v)[32], uint32_t (&pk)[16], float& ss, float& dd, bool with_dd = true) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float a = v[2 * i], b = v[2 * i + 1];
    ss = fmaf(a, a, ss);
    ss = fmaf(b, b, ss);
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);                           // .x (low half) = even k
    const uint32_t bits = *reinterpret_cast<uint32_t*>(&t);
    pk[i] = bits;
    if (with_dd) {
      const float da = a - __uint_as_float(bits << 16), db = b - __uint_as_float(bits & 0xffff0000u);
      dd = fmaf(da, da, dd);
      dd = fmaf(db, db, dd);
    }
  }
}

// per-tile stamps of ONE sweep (the fourth) for the DBG kernel: trace[cta][16 + tile][slot]
#define CCVSQ_TILE_STAMP(TL, J, SLOT)                                                                              \
  do {                                                                                                             \
    if constexpr (DBG) {                                                                                           \
      if (out.trace && (TL) == 3 && (J) < 16 && lane == 0)                                                         \
        out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + 16 + (J)) * 8 + (SLOT)] = clock64();                        \
    }                                                                                                              \
  } while (0)

template <int CG, bool DBG, int BN, bool ASM>
__global__ void __launch_bounds__(screen_threads(BN), 1)
screen_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_ext,
              const float* __restrict__ z, const Lay L, const float* __restrict__ e_max, float tau,
              int K_pad, int n_tiles, int dblk, int nacc, int abuf_n, int n_cand, int num_group_tiles,
              const ScreenOut out, const int ablate_arg) {
  // Ablation switches for timing experiments (results are WRONG when set; compiled in only with -DCCVSQ_ENABLE_ABLATE,
  // then read from CCVSQ_SCREEN_ABLATE): bit 0 = no end-of-sweep work, bit 1 = no chunk parking, bit 2 = loaders skip the
  // rounding-error norm, bit 3 / bit 7 = loaders skip all / half of the tensor-memory stores, bit 4 = loaders skip the
  // global loads, bit 5 = no end-of-sweep scan, bit 6 = no end-of-sweep barriers.  profiles/r02_screen_history.md
#ifdef CCVSQ_ENABLE_ABLATE
  const int ablate = ablate_arg;
#else
  constexpr int ablate = 0;
#endif
  using Cfg = ScreenCfg<BN>;
  constexpr bool WIDE = Cfg::WIDE;
  constexpr int NEPG = Cfg::NEPG;
  extern __shared__ __align__(1024) uint8_t smem[];
  static_assert(!ASM || BN == 128, "the shared-memory A buffer belongs to the wide-tile plan");
  const ScreenSmem lay = screen_smem_layout(dblk, CG, BN, ASM);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 1) ? 0u : cluster_ctarank();
  const int group = (int)blockIdx.x / CG, num_groups = (int)gridDim.x / CG;
  const int nslots = lay.nslots;
  // tensor-memory column map
  const uint32_t TM_EXT = (uint32_t)nacc * BN, TM_A = WIDE ? TM_EXT : TM_EXT + 8;
  const uint32_t a_cols = (uint32_t)dblk * 32;      // 32-bit columns per A buffer (D/2)
  constexpr int ROWS = BN / CG;                     // codes of a tile in this CTA's shared memory

  // barrier addresses (this CTA's copies; the leader's full / tmem_empty / a_full collect both CTAs)
  const uint32_t bar0 = smem_base + lay.bars;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_SLOTS + s); };
  auto tmem_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + b); };
  auto tmem_empty = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + MAX_ACC + b); };
  auto a_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 2 * MAX_ACC + b); };
  auto a_empty = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 2 * MAX_ACC + 2 + b); };
  auto norm_full = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 2 * MAX_ACC + 4 + b); };
  auto norm_empty = [&](int b) { return bar0 + 8u * (2 * MAX_SLOTS + 2 * MAX_ACC + 6 + b); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + lay.bars + 8u * (2 * MAX_SLOTS + 2 * MAX_ACC + 8));
  float* norm_s = reinterpret_cast<float*>(smem + lay.norm);      // [(ab*2 + h)*2 + {0: ||z||^2, 1: ||dz||^2}][row]

  if ((smem_base & 1023u) != 0) __trap();   // SWIZZLE_128B needs 1024-byte aligned tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_ext);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_SLOTS; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < MAX_ACC; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 4 * NEPG * CG);   // one arrive per epilogue warp of every CTA in the group
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(a_full(b), 8 * CG);         // one arrive per loader warp of every CTA in the group
      mbar_init(a_empty(b), 1);
      mbar_init(norm_full(b), 8);
      mbar_init(norm_empty(b), 4 * NEPG);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(smem_u32(tmem_ptr_smem), TMEM_COLS);
  if constexpr (WIDE) {
    // constant A block of the bias step in shared memory, K-major without swizzle: chunk kc (8 BF16 = 16 B) of row r
    // at kc*2048 + r*16; row = (1, 1, 1, 0, ..., 0)
    if (warp == 3) {
      uint4* cb4 = reinterpret_cast<uint4*>(smem + lay.cblk);
      for (int i = lane; i < 2 * BM; i += 32)
        cb4[i] = i < BM ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's reads
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if constexpr (!WIDE) {
    if (warp >= 4 && warp < 8) {
      // constant A block of the bias extension: (1, 1, 1, 0, ..., 0) in BF16, 16 K-elements = 8 columns
      uint32_t ext[8] = {0x3F803F80u, 0x00003F80u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TM_EXT, ext);
      tmem_st_wait();
      tc_fence_before();
    }
  }
  if constexpr (WIDE) {
    // live running-max words of the two epilogue groups: no sweep yet (uninitialised shared memory could carry tag 0)
    if (warp >= Cfg::EP_WARP0 && warp < Cfg::LD_WARP0) {
      const uint32_t w = smem_base + lay.norm + 2 * 2 * 2 * BM * 4 + (uint32_t)(threadIdx.x - Cfg::EP_WARP0 * 32) * 8u;
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(w), "r"(__float_as_uint(-FLT_MAX)), "r"(0xffffffffu) : "memory");
    }
  }
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  // Programmatic dependent launch: everything above (barriers, tensor-memory allocation, descriptor prefetch)
  // touched nothing a predecessor writes, and neither do the latent loads of the A loaders / the L2 prefetcher —
  // all of that overlaps the tail of prepare_codebook.  The roles that read the codebook shadow (TMA), max ||e||
  // (epilogue) or consume them through barriers (MMA) wait first; the others wait before they exit, so that this
  // grid's completion still implies the predecessor's.
  pdl_launch_dependents();
  if (!out.z_stable || warp < 3 || (warp >= Cfg::EP_WARP0 && warp < Cfg::LD_WARP0)) pdl_wait();

  // (WIDE: 640 threads at 96 registers each.  setmaxnreg re-balancing was tried and dropped: the values every role keeps
  //  live from the common prologue do not fit a shrunken control warpgroup, and at 96 the kernel spills ~12 words outside
  //  the per-tile loops)

  if (warp == 0) {
    // =========================== TMA producer (every CTA: its half of each B tile) ===============
    // The whole warp runs the loop (warp-uniform control flow keeps addresses in uniform registers);
    // one elected lane issues.
    int slot = 0;
    uint32_t phase = 0;
    const uint32_t fb0 = (CG == 1) ? full_bar(0) : mapa(full_bar(0), 0);
    uint32_t tl = 0;
    for (int gt = group; gt < num_group_tiles; gt += num_groups, ++tl) {
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(empty_bar(slot), phase ^ 1);
        CCVSQ_TILE_STAMP(tl, j, 6);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(slot), lay.slot_tx * CG);
          const uint32_t fb = fb0 + 8u * slot;
          const uint32_t dst = smem_base + lay.slots + (uint32_t)slot * lay.slot_bytes;
          const int row0 = j * BN + (int)rank * ROWS;
          for (int kb = 0; kb < dblk; ++kb)
            tma_load_2d<CG>(dst + kb * lay.block_bytes, &map_b, fb, kb * 64, row0);
          tma_load_2d<CG>(dst + lay.ext_off, &map_ext, fb, dblk * 64, row0);
          tma_load_2d<CG>(dst + lay.ext_off + ROWS * 16, &map_ext, fb, dblk * 64 + 8, row0);
        }
        __syncwarp();
        CCVSQ_TILE_STAMP(tl, j, 7);
        if (++slot == nslots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN);
      int slot = 0;
      uint32_t phase = 0, tl = 0, b = 0, b_phase = 0;
      for (int gt = group; gt < num_group_tiles; gt += num_groups, ++tl) {
        const uint32_t ab = tl % abuf_n, a_phase = (tl / abuf_n) & 1;
        const uint32_t a_tmem = tmem_base + TM_A + (ASM ? 0u : ab * a_cols);
        for (int j = 0; j < n_tiles; ++j) {
          if (j == 0) { if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 0] = clock64(); } }
          mbar_wait(tmem_empty(b), b_phase ^ 1);
          CCVSQ_TILE_STAMP(tl, j, 0);
          if (j == 0) {
            mbar_wait(a_full(ab), a_phase);
            if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 1] = clock64(); }
          }
          mbar_wait(full_bar(slot), phase);
          CCVSQ_TILE_STAMP(tl, j, 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + b * BN;
            const uint32_t sbase = smem_base + lay.slots + (uint32_t)slot * lay.slot_bytes;
            // bias first (overwrites the accumulator), then the D/16 K steps of the dot product
            if constexpr (WIDE)
              umma_ss<CG>(d_tmem, desc_lo(smem_base + lay.cblk, BM * 16), DESC_HI_NOSW, desc_lo(sbase + lay.ext_off, ROWS * 16),
                          DESC_HI_NOSW, idesc, 0u);
            else
              umma_ts<CG>(d_tmem, tmem_base + TM_EXT, desc_lo(sbase + lay.ext_off, ROWS * 16), DESC_HI_NOSW, idesc, 0u);
            uint32_t lo = desc_lo(sbase);
            if (ASM && ab == 1) {
              // odd row tiles: A from the shared-memory buffer (SS form), one 16-element K step = two 16-byte chunks
              // BM*16 bytes apart
              uint32_t alo = desc_lo(smem_base + lay.astage, BM * 16);
              for (int kb = 0; kb < dblk; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ss<CG>(d_tmem, alo + k * ((2 * BM * 16) >> 4), DESC_HI_NOSW, lo + k * 2, DESC_HI_SW128, idesc, 1u);
                lo += lay.block_bytes >> 4;
                alo += (8 * BM * 16) >> 4;
              }
            } else {
              uint32_t a_addr = a_tmem;
              for (int kb = 0; kb < dblk; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_ts<CG>(d_tmem, a_addr + k * 8, lo + k * 2, DESC_HI_SW128, idesc, 1u);
                lo += lay.block_bytes >> 4;
                a_addr += 32;
              }
            }
            umma_commit<CG>(empty_bar(slot));      // B slot reusable once these MMAs retire
            umma_commit<CG>(tmem_full(b));         // accumulator tile complete
            if (j == n_tiles - 1) umma_commit<CG>(a_empty(ab));   // last reader of this A buffer
          }
          __syncwarp();
          CCVSQ_TILE_STAMP(tl, j, 2);
          if (j == n_tiles - 1) { if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 2] = clock64(); } }
          if (++slot == nslots) { slot = 0; phase ^= 1; }
          if (++b == (uint32_t)nacc) { b = 0; b_phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // =========================== L2 prefetcher ===================================================
    // The A loaders are latency-bound (one latent row per lane, 32 loads in flight per lane): on short
    // code sweeps (K = 1024) a 128 KiB row tile per CTA has to arrive within one sweep.  This warp runs
    // one row tile ahead of them and pulls the tile after next into L2 (fire-and-forget prefetches, no
    // registers or shared memory), so the loaders' global loads become L2 hits.
    // (the first row tile is requested here as well: the loaders take it in four dependent rounds, and only
    // their first round would otherwise be in flight while every CTA of the chip starts cold)
    uint32_t tl = 0;
    for (int gt = group - num_groups; gt < num_group_tiles; gt += num_groups) {
      const int gn = gt + num_groups;                     // the tile the loaders take after this one
      if (gn >= num_group_tiles) break;
      if (gt >= group) {
        const uint32_t ab = tl % abuf_n, a_phase = (tl / abuf_n) & 1;
        mbar_wait_sleep(a_empty(ab), a_phase ^ 1);        // the loaders are starting on tile gt now
        ++tl;
      }
      const int64_t n0 = ((int64_t)gn * CG + rank) * BM;
      if (n0 >= L.N) break;
      const int64_t n1 = (n0 + BM < L.N ? n0 + BM : L.N) - 1;
      if (L.S == 1) {                                     // contiguous rows
        const char* base = reinterpret_cast<const char*>(z + n0 * L.D);
        const int64_t bytes = (n1 - n0 + 1) * (int64_t)L.D * 4;
        for (int64_t off = (int64_t)lane * 128; off < bytes; off += 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
      } else {                                            // 32 consecutive positions of a channel share a line
        const int64_t pos_lo = n0 / L.mult, pos_hi = n1 / L.mult;
        for (int64_t pg = pos_lo; pg <= pos_hi; pg += 32) {
          const float* pz = z + pos_base(L, pg);
          for (int c = lane; c < L.C; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pz + (int64_t)c * L.S));
        }
      }
    }
    pdl_wait();
  } else if (warp >= Cfg::LD_WARP0) {
    // =========================== A loaders: FP32 global -> BF16 -> tensor memory ================
    const int q = warp & 3, h = (warp - Cfg::LD_WARP0) >> 2;
    const int r = q * 32 + lane;
    const int half = dblk * 32;                 // dims handled by this loader half
    const int nchunk = dblk;                    // chunks of 32 dims
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + TM_A + (uint32_t)(h * half / 2);
    const uint32_t af_bar = (CG == 1) ? a_full(0) : mapa(a_full(0), 0);
    uint32_t tl = 0;
    for (int gt = group; gt < num_group_tiles; gt += num_groups, ++tl) {
      const uint32_t ab = tl % abuf_n, a_phase = (tl / abuf_n) & 1;
      const int64_t n = ((int64_t)gt * CG + rank) * BM + r;
      const bool valid = n < L.N && !(ablate & 16);
      const float* p = z;
      if (valid) {
        const int64_t pos = n / L.mult;
        const int m = (int)(n - pos * L.mult);
        p = z + pos_base(L, pos) + ((int64_t)m * L.D + (int64_t)h * half) * L.S;
      }
      float va[32], vb[32];
      load_chunk(va, p, L.S, valid);            // in flight while waiting for the buffer
      if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && warp == Cfg::LD_WARP0 && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 3] = clock64(); }
      mbar_wait_sleep(a_empty(ab), a_phase ^ 1);
      if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && warp == Cfg::LD_WARP0 && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 4] = clock64(); }
      tc_fence_after();
      float ss = 0.f, dd = 0.f;                 // ||z||^2 and ||z - bf16(z)||^2 of this half row (screening margin)
      const uint32_t dst = lane_base + (ASM ? 0u : ab * a_cols);
      const bool to_smem = ASM && ab == 1;
      // shared-memory A buffer: chunk kc (8 dims, 16 bytes) of row r at kc * BM*16 + r*16 — a warp's store of one
      // chunk is 512 contiguous bytes
      const uint32_t sdst = smem_base + lay.astage + (uint32_t)(h * (half / 8)) * (BM * 16) + (uint32_t)r * 16u;
      auto put = [&](const uint32_t (&pk)[16], int c) {
        if (to_smem) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + (uint32_t)(c * 4 + i) * (BM * 16)),
                         "r"(pk[4 * i]), "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
        } else {
          tmem_st16(dst + c * 16, pk);
        }
      };
      for (int c = 0; c < nchunk; c += 2) {
        if (c + 1 < nchunk) load_chunk(vb, p + (int64_t)(c + 1) * 32 * L.S, L.S, valid);
        {
          uint32_t pk[16];
          pack_chunk(va, pk, ss, dd, !(ablate & 4));
          if (!(ablate & 8)) put(pk, c);
        }
        if (c + 1 < nchunk) {
          if (c + 2 < nchunk) load_chunk(va, p + (int64_t)(c + 2) * 32 * L.S, L.S, valid);
          uint32_t pk[16];
          pack_chunk(vb, pk, ss, dd, !(ablate & 4));
          if (!(ablate & (8 | 128))) put(pk, c + 1);
        }
      }
      if (to_smem) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> the MMA's reads
      else tmem_st_wait();
      if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && warp == Cfg::LD_WARP0 && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 5] = clock64(); }
      mbar_wait_sleep(norm_empty(ab), a_phase ^ 1);   // the epilogue has read the previous norms of this buffer
      norm_s[((ab * 2 + h) * 2 + 0) * BM + r] = ss;
      norm_s[((ab * 2 + h) * 2 + 1) * BM + r] = dd;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(norm_full(ab));
        if (CG == 1) mbar_arrive(a_full(ab));
        else if (to_smem) mbar_arrive_cluster_release(af_bar + 8u * ab);   // (shared-memory A: read by an MMA the leader CTA issues)
        else mbar_arrive_cluster(af_bar + 8u * ab);
      }
    }
    pdl_wait();
  } else if (warp >= Cfg::EP_WARP0) {
    // =========================== epilogue: one latent row per thread and group ===================
    // NEPG == 1: the thread sees all codes of its row.  NEPG == 2 (wide tiles): group g sees columns [64g, 64g + 64)
    // of every tile and keeps its own running maximum / candidate list; the two lists meet in finalize_row.
    const int q = warp & 3, g = (warp - Cfg::EP_WARP0) >> 2;
    const int row_in_tile = q * 32 + lane;
    constexpr uint32_t CAP = (uint32_t)park_cap(BN, ASM);
    constexpr uint32_t LIST_BYTES = CAP * ENTRY_BYTES;            // one group's store
    const uint32_t pk_base = smem_base + lay.list + (uint32_t)g * LIST_BYTES + (uint32_t)row_in_tile * 16u;   // entry 0, float4 0
    const uint32_t pk_hdr = smem_base + lay.list + (uint32_t)g * LIST_BYTES + HDR_OFFSET + (uint32_t)row_in_tile * 8u;
    // wide plan: the two groups of a row tell each other their running maxima once per tile (plain shared-memory words):
    // a group stops parking noise as soon as the OTHER half of the columns has produced the row's winner
    // (each word carries the sweep number it belongs to: the two groups are not in lock step across sweep boundaries,
    //  and a value from another sweep is another row's)
    const uint32_t live_own = smem_base + lay.norm + 2 * 2 * 2 * BM * 4 + (uint32_t)(g * BM + row_in_tile) * 8u;
    const uint32_t live_par = smem_base + lay.norm + 2 * 2 * 2 * BM * 4 + (uint32_t)((1 - g) * BM + row_in_tile) * 8u;
    auto peer_running_max = [&](uint32_t sweep) {
      uint32_t vb, tag;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(vb), "=r"(tag) : "r"(live_par));
      return tag == sweep ? __uint_as_float(vb) : -FLT_MAX;
    };
    const uint32_t te_bar = (CG == 1) ? tmem_empty(0) : mapa(tmem_empty(0), 0);
    const float emax = e_max ? __ldg(e_max) : 1.f;
    const float demax = e_max ? __ldg(e_max + 1) : 0.00390625f;
    uint32_t tl = 0, b = 0, b_phase = 0;
    uint32_t ra[32], rb[32];                    // score chunks of the tile in flight (wide plan: both chunks of the half tile)
    // wide plan: wait for the next accumulator, pull this group's 64 columns into registers, hand the accumulator back
    auto load_half_tile = [&](uint32_t sweep, int j) {
      mbar_wait(tmem_full(b), b_phase);
      if (warp == 4) CCVSQ_TILE_STAMP(sweep, j, 3);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + b * BN + (uint32_t)g * 64u;
      tmem_ld32(taddr0, ra);
      tmem_ld32(taddr0 + 32, rb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(tmem_empty(b)); else mbar_arrive_cluster(te_bar + 8u * b);
      }
      if (warp == 4) CCVSQ_TILE_STAMP(sweep, j, 4);
      if (++b == (uint32_t)nacc) { b = 0; b_phase ^= 1; }
    };

    for (int gt = group; gt < num_group_tiles; gt += num_groups, ++tl) {
      const uint32_t ab = tl % abuf_n, a_phase = (tl / abuf_n) & 1;
      const int64_t row = ((int64_t)gt * CG + rank) * BM + row_in_tile;
      mbar_wait(norm_full(ab), a_phase);
      if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && warp == 4 && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 6] = clock64(); }
      // proven margin: 2 (||dz|| max||e|| (1 + 2^-8) + ||z|| max||de||) bounds the error of the DIFFERENCE of two
      // BF16-operand scores; 2^-13 ||z|| max||e|| covers the FP32 accumulation (see ccvsq_screen in ccvsq.h)
      float margin;
      {
        const float zn = sqrtf(norm_s[((ab * 2 + 0) * 2 + 0) * BM + row_in_tile] + norm_s[((ab * 2 + 1) * 2 + 0) * BM + row_in_tile]);
        const float dn = sqrtf(norm_s[((ab * 2 + 0) * 2 + 1) * BM + row_in_tile] + norm_s[((ab * 2 + 1) * 2 + 1) * BM + row_in_tile]);
        margin = tau * 2.f * (dn * emax * 1.00390625f + zn * demax) + 0.0001220703125f * zn * emax;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(norm_empty(ab));
      float runmax = -FLT_MAX, dropmax = -FLT_MAX;   // running maximum; largest maximum of a chunk the store dropped

      uint32_t cnt = 0;                         // parked chunks (valid entries: [0, cnt))
      float pmax[CAP];                          // their maxima (registers: the victim of a full store is found without loads)
#pragma unroll
      for (uint32_t e = 0; e < CAP; ++e) pmax[e] = -FLT_MAX;

      for (int j = 0; j < n_tiles; ++j) {
        if constexpr (!WIDE) {
          mbar_wait(tmem_full(b), b_phase);
          if (warp == 4) CCVSQ_TILE_STAMP(tl, j, 3);
          tc_fence_after();
        }
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + b * BN;     // (narrow plans)
        const int col0 = j * BN + (WIDE ? g * 64 : 0);
        float peer_max = -FLT_MAX;              // wide plan: the other group's running maximum (a tile old at most)
        if constexpr (WIDE) peer_max = peer_running_max(tl);

        // One 32-column chunk of this thread's row.  Fast path: maxima of the 8 groups of 4 columns (16 ops) and
        // their maximum (4 ops).  A chunk can only contribute candidates if its maximum reaches (running max - margin);
        // then it is parked whole (see ENTRY_BYTES).
        // maximum of one 32-column chunk: maxima of the 8 groups of 4 (FMNMX3 pairs) and their maximum
        auto chunk_max = [&](const uint32_t (&r)[32]) {
          float gm[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            gm[i] = fmaxf(fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), __uint_as_float(r[4 * i + 2])),
                          __uint_as_float(r[4 * i + 3]));
          return fmaxf(fmaxf(fmaxf(gm[0], gm[1]), gm[2]), fmaxf(fmaxf(fmaxf(gm[3], gm[4]), gm[5]), fmaxf(gm[6], gm[7])));
        };
        // candidate path of one chunk with maximum m
        auto park_chunk = [&](const uint32_t (&r)[32], const float m, const int cbase) {
          if constexpr (DBG) {
            if (out.dbg_scores && row < L.N) {   // diagnostic dump of the raw score tile
#pragma unroll
              for (int i = 0; i < 32; ++i) out.dbg_scores[row * K_pad + col0 + cbase + i] = __uint_as_float(r[i]);
            }
          }
          if (m >= fmaxf(runmax, peer_max) - margin && !(ablate & 2)) {
            if (m > runmax + margin) { cnt = 0; dropmax = -FLT_MAX; }   // everything parked so far is stale
            runmax = fmaxf(runmax, m);
            uint32_t slot = cnt;
            bool park = true;
            if (cnt == CAP) {                   // store full: the chunk with the smallest maximum goes (maybe this one)
              float lo = pmax[0];
              slot = 0;
#pragma unroll
              for (uint32_t e = 1; e < CAP; ++e)
                if (pmax[e] < lo) { lo = pmax[e]; slot = e; }
              park = m > lo;
              dropmax = fmaxf(dropmax, park ? lo : m);   // the largest maximum ever dropped decides whether the row gets flagged
            } else {
              ++cnt;
            }
            if (park) {
              const uint32_t eoff = slot * ENTRY_BYTES;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pk_base + eoff + (uint32_t)i * (BM * 16)),
                             "r"(r[4 * i]), "r"(r[4 * i + 1]), "r"(r[4 * i + 2]), "r"(r[4 * i + 3]) : "memory");
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pk_hdr + eoff), "r"((uint32_t)(col0 + cbase)), "r"(__float_as_uint(m)) : "memory");
#pragma unroll
              for (uint32_t e = 0; e < CAP; ++e) pmax[e] = (slot == e) ? m : pmax[e];
            }
          }
        };
        // One 32-column chunk of this thread's row.  Fast path: its maximum (20 ops).  A chunk can only contribute
        // candidates if its maximum reaches (running max - margin); then it is parked whole (see ENTRY_BYTES).
        auto process = [&](const uint32_t (&r)[32], const int cbase) { park_chunk(r, chunk_max(r), cbase); };

        // (Measured and dropped with three accumulators: rolling the two chunks through the register sets — the tcgen05.ld
        //  of one in flight while the other is processed, the next tile's first chunk requested before this tile's second
        //  is looked at.  It takes the tensor-memory load latency off the epilogue's path but hands the accumulator back
        //  half a tile later: c2 123.0 -> 129.2 us on the same box.)
        if constexpr (WIDE) {
          // both 32-column chunks of this group's half tile at once; the accumulator goes back to the MMA warp as
          // soon as they are in registers (the other group does the same with its half), the max trees run after
          load_half_tile(tl, j);
          // (both maxima first: 36 independent max operations in flight instead of two dependent trees separated by
          //  the divergent candidate path)
          const float m_a = chunk_max(ra), m_b = chunk_max(rb);
          park_chunk(ra, m_a, 0);
          park_chunk(rb, m_b, 32);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(live_own), "r"(__float_as_uint(runmax)), "r"(tl) : "memory");   // (racy by design: any earlier value of THIS sweep is a valid lower bound)
        } else {
          // BN/32 chunks of 32 columns; the tcgen05.ld of the next chunk overlaps the max tree of this
          // one, and the accumulator is handed back as soon as the last chunk is in registers
          tmem_ld32(taddr0, ra);
          tmem_ld_wait();
          tmem_ld32(taddr0 + 32, rb);
          process(ra, 0);
          tmem_ld_wait();
          if constexpr (BN == 96) {
            tmem_ld32(taddr0 + 64, ra);
            process(rb, 32);
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 1) mbar_arrive(tmem_empty(b)); else mbar_arrive_cluster(te_bar + 8u * b);
          }
          if (warp == 4) CCVSQ_TILE_STAMP(tl, j, 4);
          if constexpr (BN == 96) process(ra, 64); else process(rb, 32);
          if (++b == (uint32_t)nacc) { b = 0; b_phase ^= 1; }
        }
        if (warp == 4) CCVSQ_TILE_STAMP(tl, j, 5);
      }
      // (Measured and dropped: reading tile 0 of the NEXT sweep into the free score registers here, before the end-of-sweep
      //  work, so that its accumulator goes back to the MMA warp ~2 400 cycles earlier: c2 125.6 -> 147.5 us on the same
      //  box, although the structure alone costs nothing — profiles/r02_screen_history.md.)

      if constexpr (DBG) { if (out.trace && tl < TRACE_SWEEPS && warp == 4 && lane == 0) out.trace[((size_t)blockIdx.x * TRACE_SWEEPS + tl) * 8 + 7] = clock64(); }
      if (ablate & 1) continue;
      if constexpr (WIDE) {
        // End of the sweep.  Each group scans its OWN store against its own threshold (all 32 lanes, inline; typically
        // one live chunk) and publishes {running max, the code if exactly one is inside its margin, that count, entries};
        // after ONE barrier the rows are split between the two warps of the lane quadrant: the group holding the larger
        // maximum decides — if it has exactly one code inside the margin, dropped nothing that matters, and the other
        // group's maximum is below the threshold, the row is final right here (the common case); anything else takes
        // the general two-store finalize_row.
        const uint32_t st_own = smem_base + lay.epst + (uint32_t)(g * BM + row_in_tile) * 16u;
        const uint32_t st_par = smem_base + lay.epst + (uint32_t)((1 - g) * BM + row_in_tile) * 16u;
        uint32_t within = 0, code1 = 0;
        {
          // threshold from the larger of the two running maxima as far as this group knows it (the peer's value may be a
          // tile old: a LOWER threshold, i.e. a superset — and exact for the group that holds the row's maximum, whose
          // result is the one that decides below)
          const float thr_s = fmaxf(runmax, peer_running_max(tl)) - margin;
          if (!(ablate & 32)) scan_unique_reg<(int)CAP>(pk_base, pk_hdr, cnt, pmax, thr_s, within, code1); else within = 1;
          if (dropmax >= thr_s) within |= 0x80000000u;             // a dropped chunk may hold a code inside the margin
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_own), "r"(__float_as_uint(runmax)), "r"(code1), "r"(within), "r"(cnt) : "memory");
        if (q == 0) CCVSQ_TILE_STAMP(tl, 8 + 2 * g, 3);
        if (!(ablate & 64)) named_bar_sync(1 + q, 64);
        if (q == 0) CCVSQ_TILE_STAMP(tl, 8 + 2 * g, 4);
        bool slow_row;
        {
          // every lane evaluates its row (both warps of the pair hold the same 32 rows and see the same two states, so
          // they agree on which rows need the general path); the rows are then split between the two warps
          uint32_t rm_pb, code_p, w_p, n_p;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rm_pb), "=r"(code_p), "=r"(w_p), "=r"(n_p) : "r"(st_par));
          const float rm_p = __uint_as_float(rm_pb);
          const bool own_max = runmax >= rm_p;
          const float rmax = fmaxf(runmax, rm_p), thr = rmax - margin;
          const uint32_t w_top = own_max ? within : w_p;            // count (and drop bit) of the group holding the maximum
          const bool fast = w_top == 1u && fminf(runmax, rm_p) < thr && !out.dbg_cand;
          slow_row = !fast && row < L.N;
          if ((lane >> 4) == g && row < L.N) {
            if (fast) {
              if (out.idx) out.idx[row] = (int64_t)(own_max ? code1 : code_p);
            } else {
              RowLists rl;
              const uint32_t b0 = smem_base + lay.list + (uint32_t)row_in_tile * 16u;
              const uint32_t h0 = smem_base + lay.list + HDR_OFFSET + (uint32_t)row_in_tile * 8u;
              rl.base[0] = b0; rl.base[1] = b0 + LIST_BYTES;
              rl.hdr[0] = h0;  rl.hdr[1] = h0 + LIST_BYTES;
              rl.n[g] = cnt;   rl.n[1 - g] = n_p;
              finalize_row<2>(rl, rmax, margin, ((within | w_p) & 0x80000000u) != 0, n_cand, row, out);
            }
          }
        }
        if (q == 0) CCVSQ_TILE_STAMP(tl, 8 + 2 * g, 5);
        // the partner has read this group's store and state: they may be reused.  (Measured: taking this barrier only when
        // some row of the pair went the general way — both warps can compute that vote — changes nothing, 121.1 vs 121.3 us.)
        (void)slow_row;
        if (!(ablate & 64)) named_bar_sync(1 + q, 64);
        if (q == 0) CCVSQ_TILE_STAMP(tl, 9 + 2 * g, 3);
      } else {
        if (row < L.N) {
          const float thr = runmax - margin;
          uint32_t within = 0, best_i = 0;
          scan_unique(pk_base, pk_hdr, cnt, thr, within, best_i);
          const bool dropped = dropmax >= thr;
          if (within == 1u && !dropped && !out.dbg_cand) {
            if (out.idx) out.idx[row] = (int64_t)best_i;
          } else {
            RowLists rl;
            rl.base[0] = pk_base; rl.hdr[0] = pk_hdr; rl.n[0] = cnt;
            rl.base[1] = rl.hdr[1] = rl.n[1] = 0;
            finalize_row<1>(rl, runmax, margin, dropped, n_cand, row, out);
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
template <int CG, int BN, bool ASM = false>
static int launch_screen(const void* E_bf16, const float* z, const Lay& L, const float* e_max,
                         float tau, int K, int K_pad, int D, int n_cand, int nacc, int abuf,
                         const ScreenOut& out, cudaStream_t st) {
  static_assert(BN == 128 || BN == 96 || BN == 64, "epilogue is written for 2 or 3 chunks of 32 columns per group (narrower / odd widths were measured and dropped: profiles/r01_screen_history.md)");
  const int dblk = D / 64;
  const ScreenSmem lay = screen_smem_layout(dblk, CG, BN, ASM);
  CCVSQ_REQUIRE(lay.nslots >= 1, CCVSQ_UNSUPPORTED, "screen: D=%d leaves room for %d B slots", D, lay.nslots);
  EncodeTiledFn enc;
  if (int rc = get_encode_fn(&enc)) return rc;
  CUtensorMap mb, mx;
  if (int rc = make_map(enc, &mb, E_bf16, K_pad, D + SCREEN_EXT, 64, BN / CG, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map(enc, &mx, E_bf16, K_pad, D + SCREEN_EXT, 8, BN / CG, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  auto kern = (out.dbg_cand || out.trace) ? screen_kernel<CG, true, BN, ASM> : screen_kernel<CG, false, BN, ASM>;
  if (int rc = enable_smem(kern, lay.total)) return rc;
  const int64_t rows_per_group = (int64_t)BM * CG;
  const int64_t group_tiles = (L.N + rows_per_group - 1) / rows_per_group;
  CCVSQ_REQUIRE(group_tiles < (1ll << 24), CCVSQ_BAD_SHAPE, "screen: N=%lld too large for one launch",
                (long long)L.N);
  const int max_groups = kNumSMs / CG;
  const int groups = (int)(group_tiles < max_groups ? group_tiles : max_groups);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(screen_threads(BN));
  cfg.dynamicSmemBytes = lay.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl_enabled() && !(pdl_off_mask() & 1)) ? 2 : 1;
  const int n_tiles = (K + BN - 1) / BN;          // rows [K, n_tiles*BN) of the shadow are padding (bias -3e38)
  CCVSQ_REQUIRE(n_tiles * BN <= K_pad, CCVSQ_BAD_SHAPE, "screen: codebook shadow has %d rows, the sweep needs %d",
                K_pad, n_tiles * BN);
  static const int ablate = [] { const char* e = getenv("CCVSQ_SCREEN_ABLATE"); return e ? atoi(e) : 0; }();
  CCVSQ_CUDA(cudaLaunchKernelEx(&cfg, kern, mb, mx, z, L, e_max, tau, K_pad, n_tiles, dblk, nacc, abuf,
                                n_cand, (int)group_tiles, out, ablate));
  return CCVSQ_OK;
}

// Tensor-memory budget (512 columns).  Narrow tiles: nacc accumulators of BN columns, 8 columns of bias extension, abuf A
// buffers of D/2; the MMA -> commit -> epilogue -> arrive -> MMA chain of one accumulator is longer than two tile
// times, so three accumulators come first.  Wide tiles (BN = 128): two accumulators + abuf A buffers, nothing else.
// Short code sweeps take the wide plan (see the note at `BM`); long sweeps (K >= 2304) keep 96-column tiles with three
// accumulators and one A buffer (fewest hand-offs per code; MMA-bound at 86-90 % of the BF16 burst peak).
// Wide tiles with THREE accumulators (nacc = 3, abuf = 2): one A buffer in tensor memory, the other in shared memory; row
// tiles alternate between them (TS-form / SS-form MMAs) — the third accumulator lets the MMA warp run through the
// epilogue's end-of-sweep work.
struct ScreenPlan { int bn, nacc, abuf; };
static bool plan_fits(int D, int bn, int nacc, int abuf) {
  const int a_cols = D / 2;
  if (bn == 128 && nacc == 3)
    return abuf == 2 && D <= 256 && 3 * 128 + a_cols <= (int)TMEM_COLS && screen_smem_layout(D / 64, 2, 128, true).nslots >= 2;
  if (bn == 128) return nacc == 2 && D <= 256 && 2 * 128 + abuf * a_cols <= (int)TMEM_COLS;   // (D = 512: one B slot left)
  return nacc * bn + 8 + abuf * a_cols <= (int)TMEM_COLS;
}
static ScreenPlan plan_screen(int K, int D) {
  auto fits = [&](int bn, int nacc, int abuf) { return plan_fits(D, bn, nacc, abuf); };
  static const int forced_bn = [] { const char* e = getenv("CCVSQ_SCREEN_BN"); return e ? atoi(e) : 0; }();
  // CCVSQ_SCREEN_PLAN=bn,nacc,abuf forces a plan (A/B runs)
  static const ScreenPlan forced = [] {
    ScreenPlan p = {0, 0, 0};
    const char* e = getenv("CCVSQ_SCREEN_PLAN");
    if (e) sscanf(e, "%d,%d,%d", &p.bn, &p.nacc, &p.abuf);
    return p;
  }();
  if (forced.bn && (forced.bn == 64 || forced.bn == 96 || forced.bn == 128) && forced.nacc >= 1 && forced.nacc <= MAX_ACC &&
      forced.abuf >= 1 && forced.abuf <= 2 && fits(forced.bn, forced.nacc, forced.abuf))
    return forced;
  const bool long_sweep = (K + 95) / 96 >= 24;
  ScreenPlan best = {96, 2, 1};
  const ScreenPlan order_long[] = {{96, 3, 2}, {96, 3, 1}, {64, 3, 2}, {64, 3, 1}, {96, 2, 2}, {96, 2, 1}};
  const ScreenPlan order_short[] = {{128, 3, 2}, {128, 2, 2}, {128, 2, 1}, {96, 3, 2}, {64, 3, 2}, {96, 3, 1}, {64, 3, 1}, {96, 2, 2}, {96, 2, 1}};
  const ScreenPlan* order = long_sweep ? order_long : order_short;
  const int n_order = long_sweep ? 6 : 9;
  for (int i = 0; i < n_order; ++i) {
    const ScreenPlan& p = order[i];
    if (forced_bn && p.bn != forced_bn) continue;
    if (fits(p.bn, p.nacc, p.abuf)) { best = p; break; }
  }
  if (forced_bn == 128 && long_sweep) {       // A/B: wide tiles on long sweeps
    if (fits(128, 2, 2)) best = {128, 2, 2};
    else if (fits(128, 2, 1)) best = {128, 2, 1};
  }
  return best;
}

}  // namespace ccvsq

using namespace ccvsq;

static int screen_impl(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                       float margin_tau, int n_cand, const ScreenOut& out, int cta_group, void* stream) {
  CCVSQ_REQUIRE(z && E_bf16, CCVSQ_NULL_POINTER, "screen: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "screen: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  const int D = L.D;
  CCVSQ_REQUIRE(D % 64 == 0 && D >= 64 && D <= 512, CCVSQ_UNSUPPORTED,
                "screen: D=%d unsupported by the tensor-core path (need 64 <= D <= 512, D %% 64 == 0)", D);
  CCVSQ_REQUIRE(n_cand >= 1 && n_cand <= CCVSQ_MAX_CAND, CCVSQ_BAD_SHAPE, "screen: n_cand=%d", n_cand);
  CCVSQ_REQUIRE(L.N < (1ll << 31), CCVSQ_BAD_SHAPE, "screen: N=%lld rows exceed int32 row ids", (long long)L.N);
  CCVSQ_REQUIRE((((uintptr_t)z | (uintptr_t)E_bf16) & 15) == 0, CCVSQ_MISALIGNED,
                "screen: z and the codebook shadow must be 16-byte aligned");
  CCVSQ_REQUIRE(cta_group == 1 || cta_group == 2, CCVSQ_BAD_SHAPE, "screen: cta_group=%d", cta_group);
  const int K_pad = ccvsq_codebook_rows(K);
  const float margin_scale = margin_tau;
  const ScreenPlan pl = plan_screen(K, D);
  CCVSQ_REQUIRE(plan_fits(D, pl.bn, pl.nacc, pl.abuf), CCVSQ_UNSUPPORTED,
                "screen: D=%d does not fit the tensor-memory budget", D);
  cudaStream_t st = (cudaStream_t)stream;
  if (cta_group == 2) {
    if (pl.bn == 128 && pl.nacc == 3) return launch_screen<2, 128, true>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
    if (pl.bn == 128) return launch_screen<2, 128>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
    if (pl.bn == 64) return launch_screen<2, 64>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
    return launch_screen<2, 96>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
  }
  if (pl.bn == 128 && pl.nacc == 3) return launch_screen<1, 128, true>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
  if (pl.bn == 128) return launch_screen<1, 128>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
  if (pl.bn == 64) return launch_screen<1, 64>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
  return launch_screen<1, 96>(E_bf16, z, L, e_max, margin_scale, K, K_pad, D, n_cand, pl.nacc, pl.abuf, out, st);
}

// rows of the BF16 codebook shadow: enough for a sweep with any tile width (128 covers 64)
extern "C" int ccvsq_codebook_rows(int K) {
  const int a = (K + 95) / 96 * 96, b = (K + 127) / 128 * 128;
  return a > b ? a : b;
}

extern "C" int ccvsq_screen(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                            float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count,
                            int32_t* queue_rows, int32_t* queue_cand, uint8_t* queue_flags, void* stream) {
  CCVSQ_REQUIRE(idx && queue_count && queue_rows && queue_cand && queue_flags, CCVSQ_NULL_POINTER,
                "screen: null output pointer");
  ScreenOut out = {};
  out.idx = idx;
  out.q_count = queue_count;
  out.q_rows = queue_rows;
  out.q_cand = queue_cand;
  out.q_flags = queue_flags;
  // CCVSQ_SCREEN_CG=1: single-CTA MMA (M = 128 per CTA, no pairing) instead of the 2-CTA form, for A/B runs
  static const int cg = [] { const char* e = getenv("CCVSQ_SCREEN_CG"); return (e && atoi(e) == 1) ? 1 : 2; }();
  return screen_impl(z, lay, E_bf16, e_max, K, margin_tau, n_cand, out, cg, stream);
}

// internal (composite.cu): same as ccvsq_screen, with the promise that z is not written by the kernel right in front
namespace ccvsq {
int screen_launch_stable_z(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                           float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count, int32_t* queue_rows,
                           int32_t* queue_cand, uint8_t* queue_flags, void* stream) {
  CCVSQ_REQUIRE(idx && queue_count && queue_rows && queue_cand && queue_flags, CCVSQ_NULL_POINTER,
                "screen: null output pointer");
  ScreenOut out = {};
  out.idx = idx;
  out.q_count = queue_count;
  out.q_rows = queue_rows;
  out.q_cand = queue_cand;
  out.q_flags = queue_flags;
  out.z_stable = true;
  return screen_impl(z, lay, E_bf16, e_max, K, margin_tau, n_cand, out, 2, stream);
}
}  // namespace ccvsq

extern "C" int ccvsq_screen_trace(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                                  float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count,
                                  int32_t* queue_rows, int32_t* queue_cand, uint8_t* queue_flags, int64_t* trace,
                                  void* stream) {
  CCVSQ_REQUIRE(idx && queue_count && queue_rows && queue_cand && queue_flags && trace, CCVSQ_NULL_POINTER,
                "screen_trace: null pointer");
  ScreenOut out = {};
  out.idx = idx;
  out.q_count = queue_count;
  out.q_rows = queue_rows;
  out.q_cand = queue_cand;
  out.q_flags = queue_flags;
  out.trace = reinterpret_cast<long long*>(trace);
  return screen_impl(z, lay, E_bf16, e_max, K, margin_tau, n_cand, out, 2, stream);
}

extern "C" int ccvsq_screen_debug(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max,
                                  int K, float margin_tau, int n_cand, int cta_group, int64_t* idx,
                                  int32_t* queue_count, int32_t* queue_rows, int32_t* queue_cand,
                                  uint8_t* queue_flags, int32_t* cand_idx, float* cand_score, uint8_t* flags,
                                  float* row_margin, float* scores, void* stream) {
  CCVSQ_REQUIRE(cand_idx && cand_score && flags && row_margin, CCVSQ_NULL_POINTER,
                "screen_debug: candidate dump pointers must be non-null");
  CCVSQ_REQUIRE((queue_count == nullptr) == (queue_rows == nullptr) && (queue_count == nullptr) == (queue_cand == nullptr) &&
                    (queue_count == nullptr) == (queue_flags == nullptr),
                CCVSQ_NULL_POINTER, "screen_debug: queue pointers must be given together");
  ScreenOut out = {};
  out.idx = idx;
  out.q_count = queue_count;
  out.q_rows = queue_rows;
  out.q_cand = queue_cand;
  out.q_flags = queue_flags;
  out.dbg_cand = cand_idx;
  out.dbg_score = cand_score;
  out.dbg_flags = flags;
  out.dbg_margin = row_margin;
  out.dbg_scores = scores;
  return screen_impl(z, lay, E_bf16, e_max, K, margin_tau, n_cand, out, cta_group, stream);
}
