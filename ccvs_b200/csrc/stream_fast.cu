// stream_fast.cu — 128-bit fast paths of the HBM-bound kernels (assign / decode gather / backward /
// per-code scatter-reduce).  Same arithmetic as the generic kernels in stream_kernels.cu; these are
// taken whenever the layout allows 16-byte accesses (the generic ones stay as the fallback for odd
// shapes: e_dim = 1, S % 4 != 0, C % 32 != 0, unaligned views).
//
// Channel-major latents (S > 1: NCHW / NTCHW, element (g,c,s) at (g*C + c)*S + s)
//   A CTA owns 32 consecutive positions x one slab of <= 256 channels.  Every thread works on 4x4
//   blocks (4 consecutive positions x 4 consecutive channels):
//     * the latent side of a block is 4 LDG.128 along the spatial index (coalesced: 8 lanes cover the
//       128 contiguous bytes of one channel), issued first so they are in flight during the gather;
//     * the codebook side is gathered once per tile: E[idx[p]] rows are read with LDG.128 along the
//       channel index and parked in shared memory as [position][16-byte chunk], chunk index XOR-
//       swizzled with (position/4) so both the row-wise fill and the block-wise read (4 LDS.128, one
//       per position) are bank-conflict free without padding;
//     * the 4x4 transpose between "channels of a position" (codebook rows, atomics) and "positions of
//       a channel" (tensor I/O) happens in registers.
//   Per 64 KiB of HBM traffic the LSU sees 5 full-width passes (z load, E load, E park, E read, store)
//   instead of the 7 passes + 4-byte accesses of the generic tile kernels.
// Row-major latents (S == 1): rows are contiguous; a flat loop over 16-byte chunks, no shared memory.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace ccvsq {

constexpr int FPT = 32;    // positions per tile
constexpr int FNT = 256;   // threads per CTA
constexpr int FSLAB = 256; // channels per CTA (a tile is FPT x slab)

// 16-byte global -> shared copy that bypasses the register file (LDGSTS).  Cached in L1 as well (.ca): with a
// skewed code distribution (codebook collapse: most latents on a few codes) every SM keeps asking for the same few
// lines, and L2-only caching (.cg) serialises all of them on the L2 slices that own those lines (measured: 12x
// slower when all latents share one code).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#ifndef CCVSQ_PREFETCH_WAVES
#define CCVSQ_PREFETCH_WAVES 1
#endif
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// loss / perplexity by the last CTA to finish (replaces the separate finalize launch in forward)
__device__ void fold_finalize(const StreamArgs& a, unsigned total_ctas, bool* s_last, float* s_red) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_last = atomicAdd(a.fin.ticket, 1) == (int)(total_ctas - 1);
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (a.fin.loss && threadIdx.x == 0) {
    const double sq = atomicAdd(a.sq_err, 0.0);         // coherent read of the accumulator
    const float mse = (float)(sq / a.fin.M);            // quantize.py:60-61
    *a.fin.loss = mse + a.fin.beta * mse;
  }
  if (a.fin.perplexity && a.counts) {
    float acc = 0.f;
    for (int k = threadIdx.x; k < a.K; k += FNT) {
      const int32_t c = __ldcg(a.counts + k);
      if (a.fin.counts_f32) a.fin.counts_f32[k] = (float)c;   // exact below 2^24 per code
      const float p = (float)((double)c / a.fin.N);
      acc += p * logf(p + 1e-10f);                      // quantize.py:68
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < FNT / 32; ++w) s += s_red[w];
      *a.fin.perplexity = expf(-s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// channel-major tiles
// ------------------------------------------------------------------------------------------------
template <int MODE, int MAXIT>
#ifndef CCVSQ_CM4_MINBLOCKS
#define CCVSQ_CM4_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(FNT, MODE == MODE_BACKWARD ? 2 : CCVSQ_CM4_MINBLOCKS) cm4_kernel(const StreamArgs a, const Lay L, const int CS) {
  extern __shared__ float4 tile4[];                 // [FPT][CS/4], chunk index ^ (position/4)
  __shared__ float s_red[FNT / 32];
  __shared__ bool s_last;
  pdl_launch_dependents();
  const bool late_wait = IS_ASSIGN(MODE) && a.x_stable;   // the latents were complete before the chain began: their
  if (!late_wait) pdl_wait();                            // loads may start before the wait (the codes may not)
  const int tid = threadIdx.x;
  // assign walks the tiles from the end: it runs right after the search, which read the latents front to back, so
  // the tail of z is what the L2 still holds (126 MB) — those tiles are re-read on chip instead of from HBM
  #ifdef CCVSQ_ASSIGN_FORWARD_ORDER
  const uint32_t bx = blockIdx.x;
#else
  const uint32_t bx = IS_ASSIGN(MODE) ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
#endif
  const int64_t p0 = (int64_t)bx * FPT;
  const int np = (int)min((int64_t)FPT, L.P - p0);
  const int c_lo = (int)blockIdx.y * CS;
  const int QS = CS >> 2;                           // 16-byte chunks per tile row
  const int nblk = 8 * QS;                          // 4x4 blocks in the tile
  constexpr bool HAS_X = MODE != MODE_GATHER;
  const bool has_g = MODE == MODE_BACKWARD && a.g != nullptr;
  const bool need_e = MODE != MODE_STATS || a.sub != 0.f;

  // ---- streaming loads first (independent of the gather below).  Addressing avoids per-thread 64-bit
  // divisions: with S % 32 == 0 a tile lies inside one group g, so (g, s0) is one 32-bit division per CTA.
  float xa[MAXIT][4][4], ga[MAXIT][4][4];           // [it][channel j][position i]
  int64_t base[MAXIT];
  bool valid[MAXIT];
  const bool tile_in_group = (L.S & 31) == 0;
  int64_t tile_base = 0;
  if (tile_in_group) {
    const uint32_t tpg = (uint32_t)L.S >> 5;        // tiles per group
    const uint32_t g = bx / tpg;
    const uint32_t s0 = (bx - g * tpg) << 5;
    tile_base = ((int64_t)g * L.C + c_lo) * L.S + s0;
  }
  // ---- L2 prefetch of the tile one wave of CTAs ahead (fire and forget, no registers): the demand loads of
  // that tile then hit in L2, and HBM sees twice the requests in flight (these kernels are latency-bound: every
  // CTA goes idx -> codebook rows -> compute -> store, and only 4 CTAs fit an SM)
  if (tile_in_group) {
    const int64_t ahead = (int64_t)CCVSQ_PREFETCH_WAVES * kNumSMs * (MODE == MODE_BACKWARD ? 2 : 4);
    const int64_t nb = IS_ASSIGN(MODE) ? (int64_t)bx - ahead : (int64_t)bx + ahead;
    if (nb >= 0 && nb < (int64_t)gridDim.x) {
      const uint32_t tpg = (uint32_t)L.S >> 5;
      const uint32_t g = (uint32_t)nb / tpg;
      const uint32_t s0 = ((uint32_t)nb - g * tpg) << 5;
      const int64_t nbase = ((int64_t)g * L.C + c_lo) * L.S + s0;
      for (int c = tid; c < CS; c += FNT) {
        if (HAS_X) prefetch_l2(a.x + nbase + (int64_t)c * L.S);
        if (has_g) prefetch_l2(a.g + nbase + (int64_t)c * L.S);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < MAXIT; ++it) {
    const int b = tid + it * FNT;
    const int p4 = b & 7, cq = b >> 3;
    valid[it] = b < nblk && 4 * p4 < np;
    base[it] = 0;
    if (valid[it]) {
      if (tile_in_group) {
        base[it] = tile_base + (int64_t)(4 * cq) * L.S + 4 * p4;
      } else {
        const int64_t pos = p0 + 4 * p4;
        const int64_t g = pos / L.S;
        const int s = (int)(pos - g * L.S);
        base[it] = (g * L.C + c_lo + 4 * cq) * (int64_t)L.S + s;
      }
      if (HAS_X) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(xa[it][j]) = __ldcs(reinterpret_cast<const float4*>(a.x + base[it] + (int64_t)j * L.S));
      }
      if (has_g) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(ga[it][j]) = __ldcs(reinterpret_cast<const float4*>(a.g + base[it] + (int64_t)j * L.S));
      }
    }
  }

  // ---- everything below depends on the codes, which the search kernels in front of an assign write
  if (late_wait) pdl_wait();
  if (tile_in_group) {                               // codes of the tile one wave ahead -> L2
    const int64_t ahead = (int64_t)CCVSQ_PREFETCH_WAVES * kNumSMs * (MODE == MODE_BACKWARD ? 2 : 4);
    const int64_t nb = IS_ASSIGN(MODE) ? (int64_t)bx - ahead : (int64_t)bx + ahead;
    if (nb >= 0 && nb < (int64_t)gridDim.x && tid < (FPT * L.mult + 15) / 16)
      prefetch_l2(a.idx + nb * FPT * L.mult + tid * 16);
  }
  // ---- gather the codebook rows of this tile into shared memory: one warp per position (4 positions per warp),
  // lanes over the 16-byte chunks of the row (QS <= 64: at most two per lane).  The four codes of a warp are read
  // by four lanes in ONE load and broadcast, and the rows go global -> shared with cp.async: the whole 32 KiB gather
  // of the tile is in flight at once next to the latent loads above, without a register round trip.
  if (need_e) {
    const int warp = tid >> 5, lane = tid & 31;
    if (L.mult == 1) {
      int64_t kk = 0;
      {
        const int p = warp + (FNT / 32) * lane;
        if (lane < FPT / (FNT / 32) && p < np) {
          kk = __ldg(a.idx + p0 + p);
          if (kk < 0 || kk >= a.K) {
            if (MODE == MODE_GATHER) { if (a.err_flag) atomicOr(a.err_flag, 1); kk = 0; }
            else kk = kk < 0 ? 0 : a.K - 1;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < FPT / (FNT / 32); ++i) {
        const int p = warp + (FNT / 32) * i;
        const int64_t k = __shfl_sync(0xffffffffu, kk, i);
        if (p < np) {
          const int sw = (p >> 2) & 7;
          const float4* src = reinterpret_cast<const float4*>(a.E + (size_t)k * L.D + c_lo);
          for (int q = lane; q < QS; q += 32) cp_async16(&tile4[p * QS + (q ^ sw)], src + q);
        }
      }
    } else {
      for (int p = warp; p < np; p += FNT / 32) {
        const int sw = (p >> 2) & 7;
        for (int q = lane; q < QS; q += 32) {
          const int c = c_lo + 4 * q;
          const int m = c / L.D, j = c - m * L.D;
          int64_t k = __ldg(a.idx + (p0 + p) * L.mult + m);
          if (k < 0 || k >= a.K) {
            if (MODE == MODE_GATHER) { if (a.err_flag) atomicOr(a.err_flag, 1); k = 0; }
            else k = k < 0 ? 0 : a.K - 1;
          }
          cp_async16(&tile4[p * QS + (q ^ sw)], a.E + (size_t)k * L.D + j);
        }
      }
    }
  }
  if (a.counts && blockIdx.y == 0 && (IS_ASSIGN(MODE) || MODE == MODE_STATS)) {
    // usage counts: equal codes inside a warp are combined first (one atomic per distinct code; a collapsed
    // codebook would otherwise put every latent's atomic on the same address)
    const int rows = np * L.mult;
    for (int r0 = tid & ~31; r0 < rows; r0 += FNT) {            // warp-uniform trip count
      const int r = r0 + (tid & 31);
      int64_t k = -1;
      if (r < rows) {
        k = __ldg(a.idx + p0 * L.mult + r);
        if (IS_ASSIGN(MODE)) k = k < 0 ? 0 : (k >= a.K ? a.K - 1 : k);
        if (k < 0 || k >= a.K) k = -1;
      }
      const unsigned peers = __match_any_sync(0xffffffffu, (unsigned long long)k);
      if (k >= 0 && (tid & 31) == __ffs(peers) - 1) atomicAdd(a.counts + k, __popc(peers));
    }
  }
  cp_async_wait_all();
  __syncthreads();

  float coef = 0.f;
  if (MODE == MODE_BACKWARD) coef = __ldg(a.g_loss) * a.coef_scale;
  float acc = 0.f;
#pragma unroll
  for (int it = 0; it < MAXIT; ++it) {
    if (!valid[it]) continue;
    const int b = tid + it * FNT;
    const int p4 = b & 7, cq = b >> 3;
    float ea[4][4];                                 // [position i][channel j]
    if (need_e) {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(ea[i]) = tile4[(4 * p4 + i) * QS + (cq ^ p4)];
    }
    float oa[4][4];                                 // [channel j][position i]
    if (IS_ASSIGN(MODE)) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float diff = __fsub_rn(ea[i][j], xa[it][j][i]);     // fl(E[idx] - z)
          oa[j][i] = __fadd_rn(xa[it][j][i], diff);                 // fl(z + fl(E[idx] - z))  (quantize.py:64)
          acc = fmaf(diff, diff, acc);
        }
    } else if (MODE == MODE_BACKWARD) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float t = __fmul_rn(coef, __fsub_rn(xa[it][j][i], ea[i][j]));
          oa[j][i] = has_g ? __fadd_rn(t, ga[it][j][i]) : t;
        }
    } else if (MODE == MODE_GATHER) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) oa[j][i] = ea[i][j];
    }
    if (MODE != MODE_STATS && a.out) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4* dst = reinterpret_cast<float4*>(a.out + base[it] + (int64_t)j * L.S);
        const float4 v = *reinterpret_cast<const float4*>(oa[j]);
#ifdef CCVSQ_STREAM_ALL_STORES
        __stcs(dst, v);
#else
        if (MODE == MODE_GATHER) __stcs(dst, v); else *dst = v;
#endif
      }
    }
    if ((MODE == MODE_STATS || MODE == MODE_BACKWARD || MODE == MODE_ASSIGN_STATS) && a.resid) {
      const int c = c_lo + 4 * cq;
      const int m = L.mult == 1 ? 0 : c / L.D;
      const int j0 = c - m * L.D;
      const float sub = MODE == MODE_STATS ? a.sub : 1.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t k = __ldg(a.idx + (p0 + 4 * p4 + i) * L.mult + m);
        if (k < 0 || k >= a.K) continue;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = need_e ? xa[it][j][i] - sub * ea[i][j] : xa[it][j][i];
        red_add_v4(a.resid + (size_t)k * L.D + j0, r[0], r[1], r[2], r[3]);
      }
    }
  }

  if (IS_ASSIGN(MODE) && a.sq_err) {
    acc = warp_sum(acc);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < FNT / 32; ++w) s += s_red[w];
      atomicAdd(a.sq_err, (double)s);
    }
  }
  if (IS_ASSIGN(MODE) && a.fin.ticket) fold_finalize(a, gridDim.x * gridDim.y, &s_last, s_red);
}

// ------------------------------------------------------------------------------------------------
// row-major latents: flat loop over 16-byte chunks, U chunks per thread
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(FNT) rows4_kernel(const StreamArgs a, const int64_t N, const int D) {
  __shared__ float s_red[FNT / 32];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int U = 4;
  const int Q = D >> 2;
  const int64_t total = N * Q;
  const int64_t cb = (int64_t)blockIdx.x * (U * FNT);      // first chunk of this CTA
  const int64_t row_b = cb / Q;
  const int q_b = (int)(cb - row_b * Q);
  const bool has_g = MODE == MODE_BACKWARD && a.g != nullptr;
  const bool need_e = MODE != MODE_STATS || a.sub != 0.f;

  int64_t row[U];
  int q[U];
  int64_t k[U];
  bool valid[U];
  float4 xv[U], gv[U], ev[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int o = (int)threadIdx.x + u * FNT;
    valid[u] = cb + o < total;
    const int oo = o + q_b;
    const int dr = oo / Q;
    row[u] = row_b + dr;
    q[u] = oo - dr * Q;
    k[u] = 0;
    if (valid[u]) {
      if (MODE != MODE_GATHER) xv[u] = __ldcs(reinterpret_cast<const float4*>(a.x) + cb + o);
      if (has_g) gv[u] = __ldcs(reinterpret_cast<const float4*>(a.g) + cb + o);
      k[u] = __ldg(a.idx + row[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!valid[u]) continue;
    bool in_range = k[u] >= 0 && k[u] < a.K;
    if (!in_range) {
      if (MODE == MODE_GATHER) { if (a.err_flag) atomicOr(a.err_flag, 1); k[u] = 0; }
      else if (MODE != MODE_STATS) { k[u] = k[u] < 0 ? 0 : a.K - 1; in_range = true; }
    }
    if (need_e && (in_range || MODE == MODE_GATHER))
      ev[u] = __ldg(reinterpret_cast<const float4*>(a.E + (size_t)k[u] * D) + q[u]);
    if (MODE == MODE_STATS && !in_range) valid[u] = false;
  }
  float coef = 0.f;
  if (MODE == MODE_BACKWARD) coef = __ldg(a.g_loss) * a.coef_scale;
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!valid[u]) continue;
    const int o = (int)threadIdx.x + u * FNT;
    float x[4], e[4], g[4], out[4];
    if (MODE != MODE_GATHER) *reinterpret_cast<float4*>(x) = xv[u];
    if (need_e) *reinterpret_cast<float4*>(e) = ev[u];
    if (has_g) *reinterpret_cast<float4*>(g) = gv[u];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (IS_ASSIGN(MODE)) {
        const float diff = __fsub_rn(e[j], x[j]);
        out[j] = __fadd_rn(x[j], diff);
        acc = fmaf(diff, diff, acc);
      } else if (MODE == MODE_BACKWARD) {
        const float t = __fmul_rn(coef, __fsub_rn(x[j], e[j]));
        out[j] = has_g ? __fadd_rn(t, g[j]) : t;
      } else if (MODE == MODE_GATHER) {
        out[j] = e[j];
      }
    }
    if (MODE == MODE_GATHER && a.pos) {
      const float4 pv = __ldg(reinterpret_cast<const float4*>(a.pos) + (row[u] % a.pos_period) * Q + q[u]);
      out[0] = __fadd_rn(out[0], pv.x); out[1] = __fadd_rn(out[1], pv.y); out[2] = __fadd_rn(out[2], pv.z); out[3] = __fadd_rn(out[3], pv.w);
    }
    if (MODE != MODE_STATS && a.out) {
      float4* dst = reinterpret_cast<float4*>(a.out) + cb + o;
      const float4 v = *reinterpret_cast<const float4*>(out);
      if (MODE == MODE_GATHER) __stcs(dst, v); else *dst = v;
    }
    if ((MODE == MODE_STATS || MODE == MODE_BACKWARD || MODE == MODE_ASSIGN_STATS) && a.resid) {
      const float sub = MODE == MODE_STATS ? a.sub : 1.f;
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = need_e ? x[j] - sub * e[j] : x[j];
      red_add_v4(a.resid + (size_t)k[u] * D + 4 * q[u], r[0], r[1], r[2], r[3]);
    }
    if (a.counts && q[u] == 0 && (IS_ASSIGN(MODE) || MODE == MODE_STATS)) atomicAdd(a.counts + k[u], 1);
  }
  if (IS_ASSIGN(MODE) && a.sq_err) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < FNT / 32; ++w) s += s_red[w];
      atomicAdd(a.sq_err, (double)s);
    }
  }
  if (IS_ASSIGN(MODE) && a.fin.ticket) fold_finalize(a, gridDim.x, &s_last, s_red);
}

// Row-major, D % 128 == 0: one warp per RPW consecutive rows, lanes over the chunks of a row.  No
// divisions or per-chunk index arithmetic: ~8 (gather) to ~25 (assign) instructions per 16 bytes.
template <int MODE, int RPW>
__global__ void __launch_bounds__(FNT, MODE == MODE_BACKWARD ? 3 : 4) rowsw_kernel(const StreamArgs a, const int64_t N, const int D) {
  __shared__ float s_red[FNT / 32];
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int Q = D >> 2, QL = D >> 7;                        // chunks per row / per lane
  const int64_t n_base = ((int64_t)blockIdx.x * (FNT / 32) + (threadIdx.x >> 5)) * RPW;
  const bool has_g = MODE == MODE_BACKWARD && a.g != nullptr;
  const bool need_e = MODE != MODE_STATS || a.sub != 0.f;
  int64_t k[RPW];
  bool valid[RPW];
  {  // codes of the CTA one wave ahead -> L2 (a CTA covers 32 rows = two 128-byte lines of idx)
    const int64_t nb = ((int64_t)blockIdx.x + (int64_t)CCVSQ_PREFETCH_WAVES * kNumSMs * 8) * (FNT / 32) * RPW;
    if (threadIdx.x < 2 && nb + 16 * threadIdx.x < N) prefetch_l2(a.idx + nb + 16 * threadIdx.x);
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t n = n_base + i;
    valid[i] = n < N;
    int64_t kk = valid[i] ? __ldg(a.idx + n) : 0;
    if (kk < 0 || kk >= a.K) {
      if (MODE == MODE_GATHER) { if (a.err_flag && lane == 0) atomicOr(a.err_flag, 1); kk = 0; }
      else if (MODE == MODE_STATS) { valid[i] = false; kk = 0; }
      else kk = kk < 0 ? 0 : a.K - 1;
    }
    k[i] = kk;
    if (a.counts && valid[i] && lane == 0 && (IS_ASSIGN(MODE) || MODE == MODE_STATS)) atomicAdd(a.counts + kk, 1);
  }
  if constexpr (MODE == MODE_GATHER && RPW <= 2) {
    // decode of short inputs (the launcher picks RPW <= 2 there): every chunk of the warp's rows in flight before the
    // first store — with one row per warp and a chunk at a time a lane has ONE 16-byte load outstanding and the kernel is
    // bound by the L2 round trip (the D = 512 Kinetics shard: 0.62 of the copy bandwidth where a device fill reaches 0.89)
    if (QL <= 4 && !a.pos) {
      float4 ev[4][RPW];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < RPW; ++i)
          if (c < QL && valid[i]) ev[c][i] = __ldg(reinterpret_cast<const float4*>(a.E) + (size_t)k[i] * Q + lane + 32 * c);
      if (a.out) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < RPW; ++i)
            if (c < QL && valid[i]) __stcs(reinterpret_cast<float4*>(a.out) + (n_base + i) * Q + lane + 32 * c, ev[c][i]);
      }
      return;
    }
  }
  float coef = 0.f;
  if (MODE == MODE_BACKWARD) coef = __ldg(a.g_loss) * a.coef_scale;
  float acc = 0.f;
  for (int c = 0; c < QL; ++c) {
    const int j = lane + 32 * c;
    float4 xv[RPW], gv[RPW], ev[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      if (!valid[i]) continue;
      if (MODE != MODE_GATHER) xv[i] = __ldcs(reinterpret_cast<const float4*>(a.x) + (n_base + i) * Q + j);
      if (has_g) gv[i] = __ldcs(reinterpret_cast<const float4*>(a.g) + (n_base + i) * Q + j);
      if (need_e) ev[i] = __ldg(reinterpret_cast<const float4*>(a.E) + (size_t)k[i] * Q + j);
    }
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      if (!valid[i]) continue;
      float x[4], e[4], g[4], out[4];
      if (MODE != MODE_GATHER) *reinterpret_cast<float4*>(x) = xv[i];
      if (need_e) *reinterpret_cast<float4*>(e) = ev[i];
      if (has_g) *reinterpret_cast<float4*>(g) = gv[i];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (IS_ASSIGN(MODE)) {
          const float diff = __fsub_rn(e[t], x[t]);
          out[t] = __fadd_rn(x[t], diff);
          acc = fmaf(diff, diff, acc);
        } else if (MODE == MODE_BACKWARD) {
          const float v = __fmul_rn(coef, __fsub_rn(x[t], e[t]));
          out[t] = has_g ? __fadd_rn(v, g[t]) : v;
        } else if (MODE == MODE_GATHER) {
          out[t] = e[t];
        }
      }
      if (MODE == MODE_GATHER && a.pos) {
        const float4 pv = __ldg(reinterpret_cast<const float4*>(a.pos) + ((n_base + i) % a.pos_period) * Q + j);
        out[0] = __fadd_rn(out[0], pv.x); out[1] = __fadd_rn(out[1], pv.y); out[2] = __fadd_rn(out[2], pv.z); out[3] = __fadd_rn(out[3], pv.w);
      }
      if (MODE != MODE_STATS && a.out) {
        float4* dst = reinterpret_cast<float4*>(a.out) + (n_base + i) * Q + j;
        const float4 v = *reinterpret_cast<const float4*>(out);
        if (MODE == MODE_GATHER) __stcs(dst, v); else *dst = v;
      }
      if ((MODE == MODE_STATS || MODE == MODE_BACKWARD || MODE == MODE_ASSIGN_STATS) && a.resid) {
        const float sub = MODE == MODE_STATS ? a.sub : 1.f;
        float r[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) r[t] = need_e ? x[t] - sub * e[t] : x[t];
        red_add_v4(a.resid + ((size_t)k[i] * Q + j) * 4, r[0], r[1], r[2], r[3]);
      }
    }
  }
  if (IS_ASSIGN(MODE) && a.sq_err) {
    acc = warp_sum(acc);
    if (lane == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < FNT / 32; ++w) s += s_red[w];
      atomicAdd(a.sq_err, (double)s);
    }
  }
  if (IS_ASSIGN(MODE) && a.fin.ticket) fold_finalize(a, gridDim.x, &s_last, s_red);
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

bool stream_fast_supported(const StreamArgs& a, const Lay& L) {
  if (!aligned16(a.x) || !aligned16(a.g) || !aligned16(a.E) || !aligned16(a.out) || !aligned16(a.resid) || !aligned16(a.pos)) return false;
  if (L.D % 4 != 0) return false;
  if (L.S == 1) return true;
  if (L.S % 4 != 0 || L.C % 32 != 0) return false;
  const int nslab = (L.C + FSLAB - 1) / FSLAB;
  if (L.C % nslab != 0) return false;
  const int CS = L.C / nslab;
  return CS % 32 == 0;
}

template <int MODE>
static int launch_cm4(const StreamArgs& a, const Lay& L, cudaStream_t st) {
  const int nslab = (L.C + FSLAB - 1) / FSLAB;
  const int CS = L.C / nslab;
  const size_t smem = (size_t)FPT * CS * sizeof(float);
  const dim3 grid((unsigned)cdiv(L.P, FPT), (unsigned)nslab);
  if (CS <= 128) {
    CCVSQ_CUDA(launch_pdl(cm4_kernel<MODE, 1>, grid, dim3(FNT), smem, st, a, L, CS));
  } else {
    CCVSQ_CUDA(launch_pdl(cm4_kernel<MODE, 2>, grid, dim3(FNT), smem, st, a, L, CS));
  }
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

template <int MODE>
static int launch_rows4(const StreamArgs& a, const Lay& L, cudaStream_t st) {
  if (L.D % 128 == 0) {
    // rows per warp: 4 keeps the most loads in flight per lane; small inputs take 2 or 1 so that the grid still spans
    // several waves of CTAs (at 4 rows per warp the Kinetics shard at D = 512 is 1.7 waves: the last one 0.7 full)
    const int64_t slots = (int64_t)kNumSMs * 8;
    int rpw = 4;
    while (rpw > 1 && (L.N + rpw * (FNT / 32) - 1) / (rpw * (FNT / 32)) < 4 * slots) rpw >>= 1;
    static const int forced_rpw = [] { const char* e = getenv("CCVSQ_ROWS_RPW"); return e ? atoi(e) : 0; }();   // (A/B runs)
    if (forced_rpw == 1 || forced_rpw == 2 || forced_rpw == 4) rpw = forced_rpw;
    const int64_t blocks = (L.N + rpw * (FNT / 32) - 1) / (rpw * (FNT / 32));
    CCVSQ_REQUIRE(blocks < (1ll << 31), CCVSQ_BAD_SHAPE, "stream kernel: %lld CTAs exceed the grid limit", (long long)blocks);
    if (rpw == 4) CCVSQ_CUDA(launch_pdl(rowsw_kernel<MODE, 4>, dim3((unsigned)blocks), dim3(FNT), 0, st, a, L.N, L.D));
    else if (rpw == 2) CCVSQ_CUDA(launch_pdl(rowsw_kernel<MODE, 2>, dim3((unsigned)blocks), dim3(FNT), 0, st, a, L.N, L.D));
    else CCVSQ_CUDA(launch_pdl(rowsw_kernel<MODE, 1>, dim3((unsigned)blocks), dim3(FNT), 0, st, a, L.N, L.D));
    CCVSQ_LAUNCH_CHECK();
    return CCVSQ_OK;
  }
  const int64_t total = L.N * (L.D >> 2);
  const int64_t blocks = (total + 4 * FNT - 1) / (4 * FNT);
  CCVSQ_REQUIRE(blocks < (1ll << 31), CCVSQ_BAD_SHAPE, "stream kernel: %lld CTAs exceed the grid limit", (long long)blocks);
  CCVSQ_CUDA(launch_pdl(rows4_kernel<MODE>, dim3((unsigned)blocks), dim3(FNT), 0, st, a, L.N, L.D));
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

int stream_fast_launch(int mode, const StreamArgs& a, const Lay& L, cudaStream_t st) {
  const bool cm = L.S > 1;
  switch (mode) {
    case MODE_ASSIGN:
      if (a.resid) return cm ? launch_cm4<MODE_ASSIGN_STATS>(a, L, st) : launch_rows4<MODE_ASSIGN_STATS>(a, L, st);
      return cm ? launch_cm4<MODE_ASSIGN>(a, L, st) : launch_rows4<MODE_ASSIGN>(a, L, st);
    case MODE_BACKWARD: return cm ? launch_cm4<MODE_BACKWARD>(a, L, st) : launch_rows4<MODE_BACKWARD>(a, L, st);
    case MODE_GATHER:   return cm ? launch_cm4<MODE_GATHER>(a, L, st) : launch_rows4<MODE_GATHER>(a, L, st);
    case MODE_STATS:    return cm ? launch_cm4<MODE_STATS>(a, L, st) : launch_rows4<MODE_STATS>(a, L, st);
  }
  set_error("stream_fast_launch: bad mode %d", mode);
  return CCVSQ_BAD_SHAPE;
}

}  // namespace ccvsq
