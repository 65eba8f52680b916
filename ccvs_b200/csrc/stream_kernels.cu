// stream_kernels.cu — the HBM-bound kernels of the quantizer path (everything except the
// tcgen05 screening GEMM): codebook preparation, FP32 rescoring of queued rows, assignment
// (gather + straight-through value + squared error), decode gather, backward, per-code
// scatter-reduce, finalize, EMA update.
//
// Common structure ("position tile"): a CTA stages PT=32 consecutive positions x C channels of
// the latent tensor in shared memory.  For channel-major inputs (S>1, i.e. NCHW / NTCHW) the
// global accesses are coalesced along the spatial index (lane = position), for row-contiguous
// inputs (S==1) along the channel index.  The tile row stride is C+1 floats so that both the
// position-major fill and the channel-major per-row sweeps are bank-conflict free.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace ccvsq {

constexpr int PT = 32;        // positions per tile
constexpr int NT = 256;       // threads per CTA
constexpr int NW = NT / 32;   // warps per CTA

// ------------------------------------------------------------------------------------------------
// tile movers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_load(const float* __restrict__ x, const Lay& L, int64_t p0,
                                          int np, float* __restrict__ tile) {
  const int CP = L.C + 1;
  if (L.S == 1) {
    const float* src = x + p0 * L.C;
    const int total = np * L.C;
    for (int i = threadIdx.x; i < total; i += NT) {
      int p = i / L.C, c = i - p * L.C;
      tile[p * CP + c] = __ldg(src + i);
    }
  } else {
    const int p = threadIdx.x % PT;
    if (p < np) {
      const float* src = x + pos_base(L, p0 + p);
      float* dst = tile + p * CP;
      int c = threadIdx.x / PT;
#pragma unroll 4
      for (; c < L.C; c += NT / PT) dst[c] = __ldg(src + (int64_t)c * L.S);
    }
  }
}

// out[...] = tile (+ addend[...] if addend != nullptr)
__device__ __forceinline__ void tile_store(float* __restrict__ out, const float* __restrict__ addend,
                                           const Lay& L, int64_t p0, int np,
                                           const float* __restrict__ tile) {
  const int CP = L.C + 1;
  if (L.S == 1) {
    float* dst = out + p0 * L.C;
    const float* add = addend ? addend + p0 * L.C : nullptr;
    const int total = np * L.C;
    for (int i = threadIdx.x; i < total; i += NT) {
      int p = i / L.C, c = i - p * L.C;
      float v = tile[p * CP + c];
      if (add) v += __ldg(add + i);
      dst[i] = v;
    }
  } else {
    const int p = threadIdx.x % PT;
    if (p < np) {
      const int64_t base = pos_base(L, p0 + p);
      const float* src = tile + p * CP;
      int c = threadIdx.x / PT;
#pragma unroll 4
      for (; c < L.C; c += NT / PT) {
        float v = src[c];
        if (addend) v += __ldg(addend + base + (int64_t)c * L.S);
        out[base + (int64_t)c * L.S] = v;
      }
    }
  }
}

static inline size_t tile_smem_bytes(const Lay& L) { return (size_t)PT * (L.C + 1) * sizeof(float); }

// ------------------------------------------------------------------------------------------------
// prepare_codebook: one warp per (padded) code row.  BF16 shadow row = [ bf16(e_0..e_{D-1}) |
// hi, mid, lo, 0 x 13 ] where hi+mid+lo is a 3-term BF16 split of the bias -0.5||e||^2 (exact to
// FP32): the screen multiplies these 16 extra columns with a constant (1,1,1,0,...) block, so the
// tensor-core accumulator already holds z.e - 0.5||e||^2.  Padding rows carry a bias of -3e38.
// ------------------------------------------------------------------------------------------------
__global__ void prepare_codebook_kernel(const float* __restrict__ E, int K, int K_pad, int D,
                                        float* __restrict__ e_sq, __nv_bfloat16* __restrict__ Eb,
                                        unsigned int* __restrict__ e_max_bits) {
  pdl_launch_dependents();   // the screen may set itself up (barriers, tensor memory) while the shadow is built
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K_pad) return;
  const int DE = D + SCREEN_EXT;
  float acc = 0.f;
  if (k < K) {
    const float* e = E + (size_t)k * D;
    float dacc = 0.f;                         // ||e - bf16(e)||^2: the code's own rounding error (screening margin)
    for (int j = lane; j < D; j += 32) {
      float v = e[j];
      acc = fmaf(v, v, acc);
      const __nv_bfloat16 r = __float2bfloat16_rn(v);
      const float dv = v - __bfloat162float(r);
      dacc = fmaf(dv, dv, dacc);
      if (Eb) Eb[(size_t)k * DE + j] = r;
    }
    acc = warp_sum(acc);
    dacc = warp_sum(dacc);
    if (lane == 0) {
      e_sq[k] = acc;
      if (e_max_bits) {                       // both >= 0: uint order == float order
        atomicMax(e_max_bits, __float_as_uint(sqrtf(acc)));
        atomicMax(e_max_bits + 1, __float_as_uint(sqrtf(dacc)));
      }
    }
  } else if (Eb) {
    for (int j = lane; j < D; j += 32) Eb[(size_t)k * DE + j] = __float2bfloat16_rn(0.f);
  }
  if (Eb && lane < SCREEN_EXT) {
    float t = 0.f;
    if (k >= K) {
      t = lane == 0 ? -3.0e38f : 0.f;
    } else {
      const float bias = -0.5f * acc;
      const __nv_bfloat16 hi = __float2bfloat16_rn(bias);
      const float r1 = bias - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(mid);
      t = lane == 0 ? __bfloat162float(hi) : lane == 1 ? __bfloat162float(mid) : lane == 2 ? r2 : 0.f;
    }
    Eb[(size_t)k * DE + D + lane] = __float2bfloat16_rn(t);
  }
}

// ------------------------------------------------------------------------------------------------
// rescore: FP32 re-evaluation of the rows the screen queued (more than one code inside the margin),
// reference formula and lowest-index tie-break.  One warp per queued row; the queue length is read
// on the device.  Rows whose candidate set is incomplete (flags != 0) go to the exact fallback.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) rescore_queue_kernel(const float* __restrict__ z, Lay L,
                                                           const float* __restrict__ E,
                                                           const float* __restrict__ e_sq, int K, int n_cand,
                                                           const int32_t* __restrict__ q_count,
                                                           const int32_t* __restrict__ q_rows,
                                                           const int32_t* __restrict__ q_cand,
                                                           const uint8_t* __restrict__ q_flags,
                                                           int64_t* __restrict__ idx,
                                                           int64_t* __restrict__ fb_rows,
                                                           int32_t* __restrict__ fb_count, int64_t fb_cap) {
  // (launched the ordinary way by default, see ccvsq_rescore: the wait is a no-op then, the release lets the exact fallback —
  //  which IS a programmatic dependent of this kernel — set itself up while the queue is worked off; under
  //  CCVSQ_RESCORE_PDL=1 that release comes only after the screen has completed)
  pdl_wait();                               // the queue is written by the screen kernel
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * NW;
  const int total = *q_count;
  for (int64_t e = (int64_t)blockIdx.x * NW + (threadIdx.x >> 5); e < total; e += nwarps) {
    const int64_t n = q_rows[e];
    if (q_flags[e] != 0 && fb_rows) {
      if (lane == 0) {
        const int slot = atomicAdd(fb_count, 1);
        if (slot < fb_cap) {
          fb_rows[slot] = n;
          fb_rows[fb_cap + slot] = -1;     // packed (distance, code) key, all ones = +inf
        }
      }
      continue;
    }
    const int64_t pos = n / L.mult;
    const int m = (int)(n - pos * L.mult);
    const float* zr = z + pos_base(L, pos) + (int64_t)m * L.D * L.S;
    float zv[16];                          // D <= 512 on the tensor-core path
    float zz = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int j = lane + 32 * i;
      zv[i] = j < L.D ? __ldg(zr + (int64_t)j * L.S) : 0.f;
      zz = fmaf(zv[i], zv[i], zz);
    }
    zz = warp_sum(zz);
    float best_d = INFINITY;
    int best_k = 0x7fffffff;
    for (int c = 0; c < n_cand; ++c) {
      const int k = q_cand[e * n_cand + c];
      if (k < 0 || k >= K) continue;
      const float* er = E + (size_t)k * L.D;
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int j = lane + 32 * i;
        if (j < L.D) dot = fmaf(zv[i], __ldg(er + j), dot);
      }
      dot = warp_sum(dot);
      const float d = (zz + __ldg(e_sq + k)) - 2.f * dot;   // quantize.py:45-47 association
      if (d < best_d || (d == best_d && k < best_k)) { best_d = d; best_k = k; }
    }
    if (lane == 0 && best_k != 0x7fffffff) idx[n] = best_k;
  }
}

// ------------------------------------------------------------------------------------------------
// assign: zq = fl(z + fl(E[idx]-z)), sq_err += sum (E[idx]-z)^2, counts[idx]++
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) assign_kernel(const float* __restrict__ z, Lay L,
                                                    const float* __restrict__ E, int K,
                                                    const int64_t* __restrict__ idx,
                                                    float* __restrict__ zq, double* __restrict__ sq_err,
                                                    int32_t* __restrict__ counts) {
  extern __shared__ float tile[];
  __shared__ float red[NW];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = np * L.mult;
  const int64_t n0 = p0 * L.mult;
  tile_load(z, L, p0, np, tile);
  __syncthreads();
  float acc = 0.f;
  for (int r = warp; r < rows; r += NW) {
    const int64_t n = n0 + r;
    int64_t k = idx[n];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const int p = r / L.mult, m = r - p * L.mult;
    float* t = tile + p * CP + m * L.D;
    const float* e = E + (size_t)k * L.D;
    for (int j = lane; j < L.D; j += 32) {
      const float zv = t[j];
      const float diff = __fsub_rn(__ldg(e + j), zv);   // fl(E[idx] - z)
      t[j] = __fadd_rn(zv, diff);                       // fl(z + fl(E[idx] - z))   (quantize.py:64)
      acc = fmaf(diff, diff, acc);
    }
    if (counts && lane == 0) atomicAdd(counts + k, 1);
  }
  if (sq_err) {
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
  }
  __syncthreads();
  if (sq_err && threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w];
    atomicAdd(sq_err, (double)s);
  }
  if (zq) tile_store(zq, nullptr, L, p0, np, tile);
}

// ------------------------------------------------------------------------------------------------
// backward: dz = g_zq + (2 g_loss / M) (z - E[idx])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) backward_dz_kernel(const float* __restrict__ z, Lay L,
                                                         const float* __restrict__ E, int K,
                                                         const int64_t* __restrict__ idx,
                                                         const float* __restrict__ g_zq,
                                                         const float* __restrict__ g_loss,
                                                         float inv_M2, float* __restrict__ dz) {
  extern __shared__ float tile[];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = np * L.mult;
  const int64_t n0 = p0 * L.mult;
  const float coef = __ldg(g_loss) * inv_M2;
  tile_load(z, L, p0, np, tile);
  __syncthreads();
  for (int r = warp; r < rows; r += NW) {
    const int64_t n = n0 + r;
    int64_t k = idx[n];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const int p = r / L.mult, m = r - p * L.mult;
    float* t = tile + p * CP + m * L.D;
    const float* e = E + (size_t)k * L.D;
    for (int j = lane; j < L.D; j += 32) t[j] = coef * (t[j] - __ldg(e + j));
  }
  __syncthreads();
  tile_store(dz, g_zq, L, p0, np, tile);
}

// ------------------------------------------------------------------------------------------------
// normalize=True (quantize.py:56-57) with any `mult`: the quantized vector of a position is the CONCATENATION of its
// `mult` code rows divided by its L2 norm over all C channels,  u = y / ||y||,  y = [E[k_0] | ... | E[k_{mult-1}]].
// A position tile holds all sub-rows of its positions, so the norm, the straight-through value, the squared error
// and — in the backward — the chain rule through the normalisation are per-warp work on one position:
//   forward   z_q = fl(z + fl(u - z)),  sq_err += sum (u - z)^2,  counts[k]++
//   backward  dz  = g_zq + (2 g/M)(z - u)
//             dL/dy = (2 beta g/M)(1/||y||)(u (u.z) - z)   ->   resid[k,:] += (z - (u.z) u) / ||y||   (dE = -(2 beta g/M) resid)
// One warp per position: pass 1 over the code rows for ||y||^2 (and u.z in the backward), pass 2 for the outputs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) assign_norm_kernel(const float* __restrict__ z, Lay L, const float* __restrict__ E, int K,
                                                         const int64_t* __restrict__ idx, float* __restrict__ zq,
                                                         double* __restrict__ sq_err, int32_t* __restrict__ counts) {
  extern __shared__ float tile[];
  __shared__ float red[NW];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  tile_load(z, L, p0, np, tile);
  __syncthreads();
  float acc = 0.f;
  for (int p = warp; p < np; p += NW) {
    const int64_t n0 = (p0 + p) * L.mult;
    float ss = 0.f;
    for (int m = 0; m < L.mult; ++m) {
      int64_t k = idx[n0 + m];
      k = k < 0 ? 0 : (k >= K ? K - 1 : k);
      const float* e = E + (size_t)k * L.D;
      for (int j = lane; j < L.D; j += 32) { const float v = __ldg(e + j); ss = fmaf(v, v, ss); }
      if (counts && lane == 0) atomicAdd(counts + k, 1);
    }
    const float nrm = sqrtf(warp_sum(ss));
    for (int m = 0; m < L.mult; ++m) {
      int64_t k = idx[n0 + m];
      k = k < 0 ? 0 : (k >= K ? K - 1 : k);
      const float* e = E + (size_t)k * L.D;
      float* t = tile + p * CP + m * L.D;
      for (int j = lane; j < L.D; j += 32) {
        const float zv = t[j];
        const float diff = __fsub_rn(__fdiv_rn(__ldg(e + j), nrm), zv);   // fl(y/||y|| - z)
        t[j] = __fadd_rn(zv, diff);                                       // quantize.py:64 on the normalised vector
        acc = fmaf(diff, diff, acc);
      }
    }
  }
  if (sq_err) {
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
  }
  __syncthreads();
  if (sq_err && threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w];
    atomicAdd(sq_err, (double)s);
  }
  if (zq) tile_store(zq, nullptr, L, p0, np, tile);
}

__global__ void __launch_bounds__(NT) backward_norm_kernel(const float* __restrict__ z, Lay L, const float* __restrict__ E, int K,
                                                           const int64_t* __restrict__ idx, const float* __restrict__ g_zq,
                                                           const float* __restrict__ g_loss, float inv_M2,
                                                           float* __restrict__ dz, float* __restrict__ resid) {
  extern __shared__ float tile[];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float coef = __ldg(g_loss) * inv_M2;
  tile_load(z, L, p0, np, tile);
  __syncthreads();
  for (int p = warp; p < np; p += NW) {
    const int64_t n0 = (p0 + p) * L.mult;
    float ss = 0.f, yz = 0.f;
    for (int m = 0; m < L.mult; ++m) {
      int64_t k = idx[n0 + m];
      k = k < 0 ? 0 : (k >= K ? K - 1 : k);
      const float* e = E + (size_t)k * L.D;
      const float* t = tile + p * CP + m * L.D;
      for (int j = lane; j < L.D; j += 32) {
        const float v = __ldg(e + j);
        ss = fmaf(v, v, ss);
        yz = fmaf(v, t[j], yz);
      }
    }
    const float nrm = sqrtf(warp_sum(ss));
    const float q = warp_sum(yz) / nrm;                 // u . z
    for (int m = 0; m < L.mult; ++m) {
      int64_t k = idx[n0 + m];
      const bool in_range = k >= 0 && k < K;
      k = k < 0 ? 0 : (k >= K ? K - 1 : k);
      const float* e = E + (size_t)k * L.D;
      float* t = tile + p * CP + m * L.D;
      for (int j = lane; j < L.D; j += 32) {
        const float zv = t[j], u = __ldg(e + j) / nrm;
        if (resid && in_range) atomicAdd(resid + (size_t)k * L.D + j, (zv - q * u) / nrm);
        t[j] = coef * (zv - u);
      }
    }
  }
  __syncthreads();
  if (dz) tile_store(dz, g_zq, L, p0, np, tile);
}

// ------------------------------------------------------------------------------------------------
// code_stats: resid[k,:] += x_n - sub*E[k]; counts[k]++      (global fp32 reductions)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) code_stats_kernel(const float* __restrict__ x, Lay L,
                                                        const float* __restrict__ E, int K,
                                                        const int64_t* __restrict__ idx, float sub,
                                                        float* __restrict__ resid,
                                                        int32_t* __restrict__ counts) {
  extern __shared__ float tile[];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = np * L.mult;
  const int64_t n0 = p0 * L.mult;
  tile_load(x, L, p0, np, tile);
  __syncthreads();
  for (int r = warp; r < rows; r += NW) {
    const int64_t n = n0 + r;
    int64_t k = idx[n];
    if (k < 0 || k >= K) continue;
    const int p = r / L.mult, m = r - p * L.mult;
    const float* t = tile + p * CP + m * L.D;
    float* dst = resid + (size_t)k * L.D;
    if (sub != 0.f) {
      const float* e = E + (size_t)k * L.D;
      for (int j = lane; j < L.D; j += 32) atomicAdd(dst + j, t[j] - sub * __ldg(e + j));
    } else {
      for (int j = lane; j < L.D; j += 32) atomicAdd(dst + j, t[j]);
    }
    if (counts && lane == 0) atomicAdd(counts + k, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Deterministic per-code sums: 64-bit fixed point.  FP32 atomics make resid / dE depend on the order in which CTAs
// reach a code row (run-to-run differences in the last bits).  Integer addition is associative: every term
// t = x - sub*E[k] is scaled by a power of two 2^s chosen from the data (so that N terms cannot overflow 62 bits),
// rounded to an integer ONCE, and accumulated with 64-bit integer atomics; the sum is exact in that grid and the
// result bit-identical for any launch order.  Resolution 2^-s = 2^(e + ceil(log2 N) - 60) with 2^e > max|t|.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ a, size_t na, const float* __restrict__ b,
                                                     size_t nb, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(a[i]));
  if (b)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(b[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // NaN / Inf never enter: fmaxf drops NaN, Inf is clamped to FLT_MAX by the consumer
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // m >= 0: uint order == float order
}

__device__ __forceinline__ int fixed_shift(float amax, float sub, int64_t N) {
  int e = 0;
  const float bound = fminf(amax, 1.0e38f) * (1.f + fabsf(sub));
  frexpf(fmaxf(bound, 1e-37f), &e);             // bound < 2^e
  int lg = 0;
  while (((int64_t)1 << lg) < N) ++lg;
  return 60 - e - lg;
}

__global__ void __launch_bounds__(NT) code_stats_fixed_kernel(const float* __restrict__ x, Lay L,
                                                              const float* __restrict__ E, int K,
                                                              const int64_t* __restrict__ idx, float sub,
                                                              const float* __restrict__ amax,
                                                              unsigned long long* __restrict__ acc,
                                                              int32_t* __restrict__ counts) {
  extern __shared__ float tile[];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = np * L.mult;
  const int64_t n0 = p0 * L.mult;
  const int sh = fixed_shift(__ldg(amax), sub, L.N);
  tile_load(x, L, p0, np, tile);
  __syncthreads();
  for (int r = warp; r < rows; r += NW) {
    const int64_t n = n0 + r;
    int64_t k = idx[n];
    if (k < 0 || k >= K) continue;
    const int p = r / L.mult, m = r - p * L.mult;
    const float* t = tile + p * CP + m * L.D;
    unsigned long long* dst = acc + (size_t)k * L.D;
    const float* e = (sub != 0.f) ? E + (size_t)k * L.D : nullptr;
    for (int j = lane; j < L.D; j += 32) {
      const float v = e ? __fsub_rn(t[j], __fmul_rn(sub, __ldg(e + j))) : t[j];
      const long long q = __float2ll_rn(ldexpf(v, sh));
      atomicAdd(dst + j, (unsigned long long)q);
    }
    if (counts && lane == 0) atomicAdd(counts + k, 1);
  }
}

__global__ void __launch_bounds__(256) fixed_to_float_kernel(const long long* __restrict__ acc, size_t n, float sub, int64_t N,
                                                             const float* __restrict__ amax, float* __restrict__ out) {
  const int sh = fixed_shift(__ldg(amax), sub, N);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)ldexp((double)acc[i], -sh);
}

// ------------------------------------------------------------------------------------------------
// decode gather, row-major output: one warp per row, 128-bit loads/stores when D % 4 == 0
// ------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(NT) gather_rows_kernel(const int64_t* __restrict__ code,
                                                         const float* __restrict__ E, int K, int D,
                                                         int64_t N, float* __restrict__ out,
                                                         int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * NW + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * NW;
  constexpr int RPW = 4;  // rows in flight per warp
  for (int64_t n_base = warp_global * RPW; n_base < N; n_base += nwarps * RPW) {
    int64_t k[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int64_t n = n_base + i;
      int64_t kk = (n < N) ? __ldg(code + n) : 0;
      if (kk < 0 || kk >= K) {
        if (err_flag && lane == 0 && n < N) atomicOr(err_flag, 1);
        kk = 0;
      }
      k[i] = kk;
    }
    if (VEC4) {
      const int D4 = D >> 2;
      for (int j = lane; j < D4; j += 32) {
        float4 v[RPW];
#pragma unroll
        for (int i = 0; i < RPW; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(E + (size_t)k[i] * D) + j);
#pragma unroll
        for (int i = 0; i < RPW; ++i)
          if (n_base + i < N) __stcs(reinterpret_cast<float4*>(out + (size_t)(n_base + i) * D) + j, v[i]);
      }
    } else {
      for (int j = lane; j < D; j += 32) {
#pragma unroll
        for (int i = 0; i < RPW; ++i)
          if (n_base + i < N) out[(size_t)(n_base + i) * D + j] = __ldg(E + (size_t)k[i] * D + j);
      }
    }
  }
}

// decode gather, channel-major output (fused NHWC->NCHW of quantized_video_model.py:833)
__global__ void __launch_bounds__(NT) gather_cm_kernel(const int64_t* __restrict__ code,
                                                       const float* __restrict__ E, int K, Lay L,
                                                       float* __restrict__ out,
                                                       int32_t* __restrict__ err_flag) {
  extern __shared__ float tile[];
  const int CP = L.C + 1;
  const int64_t p0 = (int64_t)blockIdx.x * PT;
  const int np = (int)min((int64_t)PT, L.P - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = np * L.mult;
  const int64_t n0 = p0 * L.mult;
  for (int r = warp; r < rows; r += NW) {
    int64_t k = __ldg(code + n0 + r);
    if (k < 0 || k >= K) {
      if (err_flag && lane == 0) atomicOr(err_flag, 1);
      k = 0;
    }
    const int p = r / L.mult, m = r - p * L.mult;
    float* t = tile + p * CP + m * L.D;
    const float* e = E + (size_t)k * L.D;
    for (int j = lane; j < L.D; j += 32) t[j] = __ldg(e + j);
  }
  __syncthreads();
  tile_store(out, nullptr, L, p0, np, tile);
}

// ------------------------------------------------------------------------------------------------
// finalize: dE, loss, perplexity
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(const float* resid,   /* may alias dE */
                                                       const int32_t* __restrict__ counts,
                                                       const double* __restrict__ sq_err,
                                                       const float* __restrict__ g_loss, int K, int D,
                                                       double M, double N, float beta,
                                                       float* dE, float* __restrict__ loss,
                                                       float* __restrict__ perplexity,
                                                       float* __restrict__ counts_f32 = nullptr) {
  if (dE && resid) {
    const float coef = -(float)(2.0 * (double)beta / M) * __ldg(g_loss);
    const size_t total = (size_t)K * D;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
      dE[i] = coef * resid[i];
  }
  if (blockIdx.x != 0) return;
  if (loss && sq_err && threadIdx.x == 0) {
    // quantize.py:60-61: mean((zq-z)^2) + beta*mean((zq-z)^2)
    const float mse = (float)(*sq_err / M);
    *loss = mse + beta * mse;
  }
  if (perplexity && counts) {
    __shared__ float red[8];
    float acc = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      if (counts_f32) counts_f32[k] = (float)counts[k];
      const float p = (float)((double)counts[k] / N);
      acc += p * logf(p + 1e-10f);   // quantize.py:68
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w];
      *perplexity = expf(-s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// EMA codebook update (extension, see header)
// ------------------------------------------------------------------------------------------------
template <typename CT>
__global__ void ema_counts_kernel(float* __restrict__ n_ema, const CT* __restrict__ counts, int K,
                                  float decay, float* __restrict__ n_total) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = decay * n_ema[k] + (1.f - decay) * (float)counts[k];
    n_ema[k] = v;
    acc += v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    *n_total = s;
  }
}

template <typename CT>
__global__ void ema_embed_kernel(float* __restrict__ E, const float* __restrict__ n_ema,
                                 float* __restrict__ sum_ema, const float* __restrict__ resid,
                                 const CT* __restrict__ counts, int K, int D, float decay,
                                 float eps, const float* __restrict__ n_total) {
  const size_t total = (size_t)K * D;
  const float nt = *n_total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / D);
    const float s_batch = resid[i] + (float)counts[k] * E[i];   // sum of latents assigned to k
    const float s = decay * sum_ema[i] + (1.f - decay) * s_batch;
    sum_ema[i] = s;
    const float n_smooth = (n_ema[k] + eps) / (nt + (float)K * eps) * nt;   // Laplace smoothing
    E[i] = s / n_smooth;
  }
}

// Polyak average of the codebook (quantized_video_model.py:951-964):
//   par_ema.data.mul_(decay).add_(par.data, alpha=1-decay)
__global__ void polyak_kernel(float* __restrict__ ema, const float* __restrict__ live, size_t n, float decay,
                              float alpha) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    ema[i] = fmaf(alpha, live[i], __fmul_rn(ema[i], decay));
}

}  // namespace ccvsq

// =================================================================================================
// C ABI
// =================================================================================================
using namespace ccvsq;

namespace ccvsq {
// e_max[0..1] (if given) must already be zero
int prepare_codebook_launch(const float* E, int K, int D, float* e_sq, void* E_bf16, float* e_max, cudaStream_t st) {
  const int K_pad = E_bf16 ? ccvsq_codebook_rows(K) : K;
  const int wpb = 8;
  prepare_codebook_kernel<<<cdiv(K_pad, wpb), wpb * 32, 0, st>>>(E, K, K_pad, D, e_sq, (__nv_bfloat16*)E_bf16,
                                                                 (unsigned int*)e_max);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}
}  // namespace ccvsq

extern "C" int ccvsq_prepare_codebook(const float* E, int K, int D, float* e_sq, void* E_bf16, float* e_max,
                                      void* stream) {
  CCVSQ_REQUIRE(E && e_sq, CCVSQ_NULL_POINTER, "prepare_codebook: E and e_sq must be non-null");
  CCVSQ_REQUIRE(K > 0 && D > 0, CCVSQ_BAD_SHAPE, "prepare_codebook: K=%d D=%d", K, D);
  cudaStream_t st = (cudaStream_t)stream;
  if (e_max) CCVSQ_CUDA(cudaMemsetAsync(e_max, 0, 2 * sizeof(float), st));
  return prepare_codebook_launch(E, K, D, e_sq, E_bf16, e_max, st);
}

extern "C" int ccvsq_rescore(const float* z, ccvsq_layout lay, const float* E, const float* e_sq, int K,
                             int n_cand, const int32_t* queue_count, const int32_t* queue_rows,
                             const int32_t* queue_cand, const uint8_t* queue_flags, int64_t* idx,
                             int64_t* fallback_ws, int32_t* fallback_count, int64_t fallback_capacity,
                             void* stream) {
  CCVSQ_REQUIRE(z && E && e_sq && queue_count && queue_rows && queue_cand && queue_flags && idx,
                CCVSQ_NULL_POINTER, "rescore: null pointer");
  CCVSQ_REQUIRE(n_cand >= 1 && n_cand <= CCVSQ_MAX_CAND, CCVSQ_BAD_SHAPE, "rescore: n_cand=%d", n_cand);
  CCVSQ_REQUIRE((fallback_ws == nullptr) == (fallback_count == nullptr), CCVSQ_NULL_POINTER,
                "rescore: fallback_ws and fallback_count must be given together");
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  CCVSQ_REQUIRE(L.D <= 512, CCVSQ_UNSUPPORTED, "rescore: D=%d > 512", L.D);
  // Grid = what the chip holds at this kernel's occupancy (a grid-stride loop over the queue).
  // NOT launched as a programmatic dependent of the screen (the rest of the chain is): its CTAs, parked early on the SMs
  // the screen leaves first, cost the fresh-init regime — a third of the rows queued, thousands in the exact fallback —
  // a quarter of the whole forward (1 032 -> 763 us at the c2 shape, same-box A/B per kernel: CCVSQ_NO_PDL_KERNELS),
  // while the trained-codebook regime gains only the ~1 us of launch latency the early start hides (230.7 -> 231.8 us).
  // CCVSQ_RESCORE_PDL=1 restores the early launch for A/B runs.
  static const bool rescore_pdl = [] { const char* e = getenv("CCVSQ_RESCORE_PDL"); return e && atoi(e) != 0; }();
  static int per_sm = 0;
  if (per_sm == 0) {
    int n = 0;
    CCVSQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rescore_queue_kernel, NT, 0));
    per_sm = n > 0 ? n : 4;
  }
  int64_t blocks = (L.N + NW - 1) / NW;
  if (blocks > (int64_t)kNumSMs * per_sm) blocks = (int64_t)kNumSMs * per_sm;
  CCVSQ_CUDA(launch_pdl_if(rescore_pdl && !(pdl_off_mask() & 2), rescore_queue_kernel, dim3((unsigned)blocks), dim3(NT), 0, (cudaStream_t)stream, z, L, E, e_sq, K,
                        n_cand, queue_count, queue_rows, queue_cand, queue_flags, idx, fallback_ws, fallback_count,
                        fallback_capacity));
  return CCVSQ_OK;
}

// generic (any layout) or 128-bit fast path (stream_fast.cu), same arithmetic
namespace ccvsq {
int stream_launch(int mode, const StreamArgs& a, const Lay& L, cudaStream_t st) {
  if (stream_fast_supported(a, L)) return stream_fast_launch(mode, a, L, st);
  const unsigned tiles = (unsigned)cdiv(L.P, PT);
  const size_t smem = tile_smem_bytes(L);
  switch (mode) {
    case MODE_ASSIGN:
      if (int rc = enable_smem(assign_kernel, smem)) return rc;
      assign_kernel<<<tiles, NT, smem, st>>>(a.x, L, a.E, a.K, a.idx, a.out, a.sq_err, a.counts);
      if (a.resid) {   // generic layouts: the per-code residual sums take their own pass
        CCVSQ_LAUNCH_CHECK();
        if (int rc = enable_smem(code_stats_kernel, smem)) return rc;
        code_stats_kernel<<<tiles, NT, smem, st>>>(a.x, L, a.E, a.K, a.idx, 1.f, a.resid, nullptr);
      }
      break;
    case MODE_BACKWARD:
      if (a.out) {
        if (int rc = enable_smem(backward_dz_kernel, smem)) return rc;
        backward_dz_kernel<<<tiles, NT, smem, st>>>(a.x, L, a.E, a.K, a.idx, a.g, a.g_loss, a.coef_scale, a.out);
        CCVSQ_LAUNCH_CHECK();
      }
      if (a.resid) {
        if (int rc = enable_smem(code_stats_kernel, smem)) return rc;
        code_stats_kernel<<<tiles, NT, smem, st>>>(a.x, L, a.E, a.K, a.idx, 1.f, a.resid, nullptr);
      }
      break;
    case MODE_STATS:
      if (int rc = enable_smem(code_stats_kernel, smem)) return rc;
      code_stats_kernel<<<tiles, NT, smem, st>>>(a.x, L, a.E, a.K, a.idx, a.sub, a.resid, a.counts);
      break;
    case MODE_GATHER:
      if (L.S == 1) {
        const int64_t warps_needed = (L.N + 3) / 4;
        const int64_t blocks = (warps_needed + NW - 1) / NW;
        CCVSQ_REQUIRE(blocks < (1ll << 31), CCVSQ_BAD_SHAPE, "gather: N=%lld too large", (long long)L.N);
        gather_rows_kernel<false><<<(unsigned)blocks, NT, 0, st>>>(a.idx, a.E, a.K, L.D, L.N, a.out, a.err_flag);
      } else {
        if (int rc = enable_smem(gather_cm_kernel, smem)) return rc;
        gather_cm_kernel<<<tiles, NT, smem, st>>>(a.idx, a.E, a.K, L, a.out, a.err_flag);
      }
      break;
    default:
      set_error("stream_launch: bad mode %d", mode);
      return CCVSQ_BAD_SHAPE;
  }
  CCVSQ_LAUNCH_CHECK();
  // the generic assign kernel has no folded finalize: one extra (tiny) launch
  if (mode == MODE_ASSIGN && a.fin.ticket && (a.fin.loss || a.fin.perplexity)) {
    finalize_kernel<<<1, 256, 0, st>>>(nullptr, a.counts, a.sq_err, nullptr, a.K, L.D, a.fin.M, a.fin.N, a.fin.beta,
                                      nullptr, a.fin.loss, a.fin.perplexity, a.fin.counts_f32);
    CCVSQ_LAUNCH_CHECK();
  }
  return CCVSQ_OK;
}
}  // namespace ccvsq

extern "C" int ccvsq_assign(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                            float* zq_out, double* sq_err, int32_t* counts, void* stream) {
  CCVSQ_REQUIRE(z && E && idx, CCVSQ_NULL_POINTER, "assign: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "assign: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  StreamArgs a = {};
  a.x = z; a.E = E; a.idx = idx; a.out = zq_out; a.sq_err = sq_err; a.counts = counts; a.K = K;
  return stream_launch(MODE_ASSIGN, a, L, (cudaStream_t)stream);
}

extern "C" int ccvsq_assign_normalized(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                                       float* zq_out, double* sq_err, int32_t* counts, void* stream) {
  CCVSQ_REQUIRE(z && E && idx, CCVSQ_NULL_POINTER, "assign_normalized: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "assign_normalized: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  const size_t smem = tile_smem_bytes(L);
  if (int rc = enable_smem(assign_norm_kernel, smem)) return rc;
  assign_norm_kernel<<<(unsigned)cdiv(L.P, PT), NT, smem, (cudaStream_t)stream>>>(z, L, E, K, idx, zq_out, sq_err, counts);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_backward_normalized(const float* z, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                                         const float* g_zq, const float* g_loss, float* dz, float* resid, void* stream) {
  CCVSQ_REQUIRE(z && E && idx && g_loss, CCVSQ_NULL_POINTER, "backward_normalized: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "backward_normalized: K=%d", K);
  if (!dz && !resid) return CCVSQ_OK;
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (resid) CCVSQ_CUDA(cudaMemsetAsync(resid, 0, (size_t)K * L.D * sizeof(float), st));
  const size_t smem = tile_smem_bytes(L);
  if (int rc = enable_smem(backward_norm_kernel, smem)) return rc;
  backward_norm_kernel<<<(unsigned)cdiv(L.P, PT), NT, smem, st>>>(z, L, E, K, idx, g_zq, g_loss,
                                                                  (float)(2.0 / ((double)L.P * L.C)), dz, resid);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_backward_dz(const float* z, ccvsq_layout lay, const float* E, int K,
                                 const int64_t* idx, const float* g_zq, const float* g_loss, float* dz,
                                 void* stream) {
  CCVSQ_REQUIRE(z && E && idx && g_loss && dz, CCVSQ_NULL_POINTER, "backward_dz: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "backward_dz: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  StreamArgs a = {};
  a.x = z; a.g = g_zq; a.E = E; a.idx = idx; a.out = dz; a.g_loss = g_loss; a.K = K;
  a.coef_scale = (float)(2.0 / ((double)L.P * L.C));
  return stream_launch(MODE_BACKWARD, a, L, (cudaStream_t)stream);
}

extern "C" int ccvsq_code_stats(const float* x, ccvsq_layout lay, const float* E, int K,
                                const int64_t* idx, float sub, float* resid, int32_t* counts,
                                void* stream) {
  CCVSQ_REQUIRE(x && idx && resid, CCVSQ_NULL_POINTER, "code_stats: null pointer");
  CCVSQ_REQUIRE(sub == 0.f || E, CCVSQ_NULL_POINTER, "code_stats: E required when sub != 0");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "code_stats: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  StreamArgs a = {};
  a.x = x; a.E = E; a.idx = idx; a.resid = resid; a.counts = counts; a.sub = sub; a.K = K;
  return stream_launch(MODE_STATS, a, L, (cudaStream_t)stream);
}

extern "C" int ccvsq_code_stats_fixed(const float* x, ccvsq_layout lay, const float* E, int K, const int64_t* idx,
                                      float sub, int64_t* acc, float* amax_scratch, float* resid, int32_t* counts,
                                      void* stream) {
  CCVSQ_REQUIRE(x && idx && acc && amax_scratch && resid, CCVSQ_NULL_POINTER, "code_stats_fixed: null pointer");
  CCVSQ_REQUIRE(sub == 0.f || E, CCVSQ_NULL_POINTER, "code_stats_fixed: E required when sub != 0");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "code_stats_fixed: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t kd = (size_t)K * L.D;
  CCVSQ_CUDA(cudaMemsetAsync(acc, 0, kd * sizeof(int64_t), st));
  CCVSQ_CUDA(cudaMemsetAsync(amax_scratch, 0, sizeof(float), st));
  const size_t nx = (size_t)L.P * L.C;
  int blocks = cdiv((int64_t)nx, 256 * 8);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  absmax_kernel<<<blocks, 256, 0, st>>>(x, nx, sub != 0.f ? E : nullptr, kd, (unsigned int*)amax_scratch);
  CCVSQ_LAUNCH_CHECK();
  const size_t smem = tile_smem_bytes(L);
  if (int rc = enable_smem(code_stats_fixed_kernel, smem)) return rc;
  code_stats_fixed_kernel<<<(unsigned)cdiv(L.P, PT), NT, smem, st>>>(x, L, E, K, idx, sub, amax_scratch,
                                                                   (unsigned long long*)acc, counts);
  CCVSQ_LAUNCH_CHECK();
  int cb = cdiv((int64_t)kd, 256 * 4);
  if (cb > kNumSMs * 8) cb = kNumSMs * 8;
  fixed_to_float_kernel<<<cb, 256, 0, st>>>((const long long*)acc, kd, sub, L.N, amax_scratch, resid);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_gather(const int64_t* code, const float* E, int K, ccvsq_layout out_lay, float* out,
                            int32_t* err_flag, void* stream) {
  CCVSQ_REQUIRE(code && E && out, CCVSQ_NULL_POINTER, "gather: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "gather: K=%d", K);
  Lay L;
  if (int rc = make_lay(out_lay, &L)) return rc;
  StreamArgs a = {};
  a.E = E; a.idx = code; a.out = out; a.err_flag = err_flag; a.K = K;
  return stream_launch(MODE_GATHER, a, L, (cudaStream_t)stream);
}

extern "C" int ccvsq_gather_add(const int64_t* code, const float* table, int K, int D, int64_t N, const float* pos,
                                int64_t pos_period, float* out, int32_t* err_flag, void* stream) {
  CCVSQ_REQUIRE(code && table && pos && out, CCVSQ_NULL_POINTER, "gather_add: null pointer");
  CCVSQ_REQUIRE(K > 0 && D > 0 && N > 0 && pos_period > 0, CCVSQ_BAD_SHAPE, "gather_add: K=%d D=%d N=%lld period=%lld", K,
                D, (long long)N, (long long)pos_period);
  Lay L;
  ccvsq_layout lay = {N, D, 1, 1};
  if (int rc = make_lay(lay, &L)) return rc;
  StreamArgs a = {};
  a.E = table; a.idx = code; a.out = out; a.err_flag = err_flag; a.K = K; a.pos = pos; a.pos_period = pos_period;
  CCVSQ_REQUIRE(stream_fast_supported(a, L), CCVSQ_UNSUPPORTED,
                "gather_add: needs D %% 4 == 0 and 16-byte aligned table / pos / out (D=%d)", D);
  return stream_fast_launch(MODE_GATHER, a, L, (cudaStream_t)stream);
}

extern "C" int ccvsq_polyak(float* ema, const float* live, int64_t n, double decay, void* stream) {
  CCVSQ_REQUIRE(ema && live, CCVSQ_NULL_POINTER, "polyak: null pointer");
  CCVSQ_REQUIRE(n > 0, CCVSQ_BAD_SHAPE, "polyak: n=%lld", (long long)n);
  int blocks = cdiv(n, 256 * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  polyak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ema, live, (size_t)n, (float)decay, (float)(1.0 - decay));
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_finalize(const float* resid, const int32_t* counts, const double* sq_err,
                              const float* g_loss, int K, int D, double M, double N, float beta,
                              float* dE, float* loss, float* perplexity, void* stream) {
  CCVSQ_REQUIRE(K > 0 && D > 0 && M > 0 && N > 0, CCVSQ_BAD_SHAPE, "finalize: K=%d D=%d M=%g N=%g", K,
                D, M, N);
  CCVSQ_REQUIRE(!dE || (resid && g_loss), CCVSQ_NULL_POINTER, "finalize: dE needs resid and g_loss");
  CCVSQ_REQUIRE(!loss || sq_err, CCVSQ_NULL_POINTER, "finalize: loss needs sq_err");
  CCVSQ_REQUIRE(!perplexity || counts, CCVSQ_NULL_POINTER, "finalize: perplexity needs counts");
  int blocks = dE ? cdiv((int64_t)K * D, 256 * 8) : 1;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(resid, counts, sq_err, g_loss, K, D, M, N,
                                                         beta, dE, loss, perplexity);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

template <typename CT>
static int ema_update_impl(float* E, float* n_ema, float* sum_ema, const float* resid, const CT* counts, int K, int D,
                           float decay, float eps, float* scratch, void* stream) {
  CCVSQ_REQUIRE(E && n_ema && sum_ema && resid && counts && scratch, CCVSQ_NULL_POINTER,
                "ema_update: null pointer");
  CCVSQ_REQUIRE(K > 0 && D > 0, CCVSQ_BAD_SHAPE, "ema_update: K=%d D=%d", K, D);
  cudaStream_t st = (cudaStream_t)stream;
  ema_counts_kernel<CT><<<1, 256, 0, st>>>(n_ema, counts, K, decay, scratch);
  CCVSQ_LAUNCH_CHECK();
  int blocks = cdiv((int64_t)K * D, 256 * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  ema_embed_kernel<CT><<<blocks, 256, 0, st>>>(E, n_ema, sum_ema, resid, counts, K, D, decay, eps, scratch);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_ema_update(float* E, float* n_ema, float* sum_ema, const float* resid,
                                const int32_t* counts, int K, int D, float decay, float eps,
                                float* scratch, void* stream) {
  return ema_update_impl<int32_t>(E, n_ema, sum_ema, resid, counts, K, D, decay, eps, scratch, stream);
}

extern "C" int ccvsq_ema_update_packed(float* E, float* n_ema, float* sum_ema, const float* packed, int K, int D,
                                       float decay, float eps, float* scratch, void* stream) {
  CCVSQ_REQUIRE(packed, CCVSQ_NULL_POINTER, "ema_update_packed: null pointer");
  return ema_update_impl<float>(E, n_ema, sum_ema, packed, packed + (size_t)K * D, K, D, decay, eps, scratch, stream);
}
