// search_exact.cu — exact FP32 nearest-code search (no tensor cores).
//
// Reference: quantize.py:45-50.  d[n,k] = fl(fl(||z_n||^2 + ||e_k||^2) - 2*<z_n, e_k>) and a
// first-occurrence argmin, without ever materialising the N x K matrix.  This is the parity
// anchor on the GPU (pure FP32 FMA arithmetic), the path for shapes the tensor-core screen does
// not cover (e_dim = 1 state quantizer, D not a multiple of 64, ...) and the fallback for rows
// the screen flags as "too many candidates inside the margin".
//
// Tiling: a CTA owns 64 latent rows whose full D-vector stays resident in shared memory
// (zs[D][64]); the codebook streams through in 64-code x 16-dim slabs.  256 threads, each a
// 4 rows x 4 codes register tile.
#include <math.h>
#include "common.cuh"

namespace ccvsq {

constexpr int XR = 64;    // rows per CTA
constexpr int XC = 64;    // codes per slab
constexpr int XK = 16;    // dims per slab
constexpr int XT = 256;   // threads

// order-preserving map float -> uint32, packed with the code so that a 64-bit unsigned minimum is the
// lexicographic (distance, code) minimum, i.e. the first-occurrence argmin
__device__ __forceinline__ unsigned long long pack_key(float d, int k) {
  uint32_t b = __float_as_uint(d);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (uint32_t)k;
}

// row_list == nullptr: all L.N rows, one K sweep per CTA, indices written directly.
// row_list != nullptr: only the listed rows (count on the device); blockIdx.y selects a slice of the
//   codebook, partial results meet in keys[] via 64-bit atomicMin and the last CTA to finish turns
//   the keys into indices (single launch, no host sync).
__global__ void __launch_bounds__(XT) search_exact_kernel(const float* __restrict__ z, Lay L,
                                                          const float* __restrict__ E,
                                                          const float* __restrict__ e_sq, int K,
                                                          const int64_t* __restrict__ row_list,
                                                          unsigned long long* __restrict__ keys,
                                                          int32_t* __restrict__ row_count,
                                                          int64_t max_rows, int64_t* __restrict__ idx) {
  extern __shared__ float smem[];
  const int Dp = ((L.D + XK - 1) / XK) * XK;      // D padded to the slab depth
  float* zs = smem;                               // [Dp][XR]
  float* es = zs + (size_t)Dp * XR;               // [XK][XC]
  float* zz = es + XK * XC;                       // [XR]
  float* red_d = zz + XR;                         // [XR][16]
  int* red_k = reinterpret_cast<int*>(red_d + XR * 16);   // [XR][16]
  __shared__ int64_t row_id[XR];

  // Programmatic dependent launch: the whole-codebook search lets its successor in as early as possible.  The FALLBACK
  // launch must not — its successor in the forward chain is the assign kernel, whose early CTAs (4 per SM, 32 KiB of
  // shared memory each, parked in griddepcontrol.wait) would leave this kernel one CTA per SM instead of three exactly
  // when it has real work (fresh-init codebook: 280 -> 560 us).  So the fallback decides AFTER it has seen the count:
  // nothing flagged (the usual case) -> release the successor and leave; rows to search -> release it at the end.
  if (!row_list) pdl_launch_dependents();
  pdl_wait();                                     // the fallback list / the codebook norms come from predecessors
  const int64_t total_rows = row_list ? min((int64_t)row_count[0], max_rows) : L.N;
  if (row_list && total_rows == 0) {         // (every CTA sees the same count, so the ticket below is skipped consistently)
    pdl_launch_dependents();
    return;
  }
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  // gridDim.y CTAs share a row tile's codebook sweep ("code slices") or take different row tiles ("row lanes"), decided
  // HERE from the row count, which only the device knows: few flagged rows (the usual fallback: a few thousand rows on
  // a fresh-init codebook) -> every CTA of the column takes a slice of the codebook, so that ~100 row tiles still fill the
  // chip several CTAs deep; many rows (a collapsed codebook flags all of them) -> only as many code slices as the
  // codebook size asks for (1024 codes each), the other CTAs become row lanes and nothing is staged more often than needed.
  const int ny = (int)gridDim.y;
  const int need = min(ny, (K + 1023) / 1024);                    // code slices the codebook size asks for
  const int64_t tiles = (total_rows + XR - 1) / XR;
  const bool by_code = !row_list || ny % need != 0 || tiles < 2 * (int64_t)gridDim.x * (ny / need);
  const int n_code = by_code ? ny : need;                        // code slices in use
  const int n_lane = ny / n_code;                                // row lanes in use
  const int code_slice = (int)blockIdx.y % n_code, lane_id = (int)blockIdx.y / n_code;
  // this CTA's slice of the codebook (whole codebook when there is one slice), aligned to the slab width
  const int slice = ((K + n_code - 1) / n_code + XC - 1) / XC * XC;
  const int k_begin = code_slice * slice;
  const int k_end = min(K, k_begin + slice);

  for (int64_t r0 = ((int64_t)blockIdx.x * n_lane + lane_id) * XR; r0 < total_rows; r0 += (int64_t)gridDim.x * n_lane * XR) {
    const int nr = (int)min((int64_t)XR, total_rows - r0);
    __syncthreads();   // previous iteration done with smem
    if (tid < XR) row_id[tid] = (tid < nr) ? (row_list ? row_list[r0 + tid] : r0 + tid) : -1;
    __syncthreads();

    // ---- stage the 64 rows: zs[j][r]
    if (L.S == 1) {
      for (int i = tid; i < XR * Dp; i += XT) {
        const int r = i / Dp, j = i - r * Dp;
        float v = 0.f;
        const int64_t n = row_id[r];
        if (n >= 0 && j < L.D) v = __ldg(z + n * L.D + j);   // S==1: row n is contiguous (mult folds in)
        zs[j * XR + r] = v;
      }
    } else {
      const int r = tid & (XR - 1);
      const int64_t n = row_id[r];
      int64_t base = 0;
      if (n >= 0) {
        const int64_t p = n / L.mult;
        const int m = (int)(n - p * L.mult);
        base = pos_base(L, p) + (int64_t)m * L.D * L.S;
      }
      for (int j = tid / XR; j < Dp; j += XT / XR) {
        float v = 0.f;
        if (n >= 0 && j < L.D) v = __ldg(z + base + (int64_t)j * L.S);
        zs[j * XR + r] = v;
      }
    }
    __syncthreads();
    if (tid < XR) {
      float acc = 0.f;
      for (int j = 0; j < L.D; ++j) acc = fmaf(zs[j * XR + tid], zs[j * XR + tid], acc);
      zz[tid] = acc;
    }
    // (zz is consumed after the next __syncthreads inside the slab loop)

    float best_d[4];
    int best_k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best_d[i] = INFINITY; best_k[i] = 0x7fffffff; }

    for (int k0 = k_begin; k0 < k_end; k0 += XC) {
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;

      for (int j0 = 0; j0 < Dp; j0 += XK) {
        __syncthreads();
        // slab: es[jj][code] for 64 codes x 16 dims; thread -> (code = tid/4, 4 dims)
        {
          const int code = tid >> 2, jq = (tid & 3) * 4;
          const int k = k0 + code;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + jq + u;
            float v = 0.f;
            if (k < k_end && j < L.D) v = __ldg(E + (size_t)k * L.D + j);
            es[(jq + u) * XC + code] = v;
          }
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < XK; ++jj) {
          const float4 zv = *reinterpret_cast<const float4*>(zs + (j0 + jj) * XR + ty * 4);
          const float4 ev = *reinterpret_cast<const float4*>(es + jj * XC + tx * 4);
          const float zr[4] = {zv.x, zv.y, zv.z, zv.w};
          const float er[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(zr[i], er[c], acc[i][c]);
        }
      }
      // distances for this slab of codes (increasing k within the thread -> strict '<' keeps first)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = k0 + tx * 4 + c;
        if (k < k_end) {
          const float ee = __ldg(e_sq + k);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float d = (zz[ty * 4 + i] + ee) - 2.f * acc[i][c];
            if (d < best_d[i] || (d == best_d[i] && k < best_k[i])) { best_d[i] = d; best_k[i] = k; }
          }
        }
      }
    }
    // ---- merge the 16 threads that share a row (lexicographic (d, k) minimum)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      red_d[(ty * 4 + i) * 16 + tx] = best_d[i];
      red_k[(ty * 4 + i) * 16 + tx] = best_k[i];
    }
    __syncthreads();
    if (tid < nr) {
      float bd = red_d[tid * 16];
      int bk = red_k[tid * 16];
      for (int t = 1; t < 16; ++t) {
        const float d = red_d[tid * 16 + t];
        const int k = red_k[tid * 16 + t];
        if (d < bd || (d == bd && k < bk)) { bd = d; bk = k; }
      }
      if (row_list) {
        if (bk != 0x7fffffff) atomicMin(keys + r0 + tid, pack_key(bd, bk));
      } else {
        idx[row_id[tid]] = (bk == 0x7fffffff) ? 0 : bk;
      }
    }
  }
  if (row_list) {
    // last CTA to finish converts the packed keys of all listed rows into plain indices
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(row_count + 1, 1) == (int)(gridDim.x * gridDim.y) - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (int64_t i = tid; i < total_rows; i += XT) {
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(keys + i);
        // (a row whose distances are all NaN / +inf never beats the all-ones key: give it code 0, in range like
        //  torch.argmin's answer, instead of 2^32 - 1)
        idx[row_list[i]] = key == ~0ull ? 0 : (int64_t)(key & 0xffffffffull);
      }
    }
  }
}

static size_t exact_smem_bytes(int D) {
  const int Dp = ((D + XK - 1) / XK) * XK;
  return ((size_t)Dp * XR + XK * XC + XR + XR * 16 * 2) * sizeof(float);
}

static int launch_exact(const float* z, const Lay& L, const float* E, const float* e_sq, int K,
                        const int64_t* rows, unsigned long long* keys, int32_t* row_count,
                        int64_t max_rows, int64_t* idx, cudaStream_t st) {
  const size_t smem = exact_smem_bytes(L.D);
  CCVSQ_REQUIRE(smem <= 227 * 1024, CCVSQ_UNSUPPORTED,
                "search_exact: D=%d needs %zu bytes of shared memory (> 227 KiB)", L.D, smem);
  if (int rc = enable_smem(search_exact_kernel, smem)) return rc;
  const int64_t work = rows ? max_rows : L.N;
  int64_t blocks = (work + XR - 1) / XR;
  int slices = 1;
  if (rows) {
    // few rows, possibly a large codebook: split K so that a handful of rows still fills the chip (the kernel turns
    // surplus code slices into row lanes when the list is long, see there)
    const int need = (K + 1023) / 1024 > 16 ? 16 : (K + 1023) / 1024;
    slices = need;
    while (slices < 4 && K / (slices + need) >= XC) slices += need;     // a multiple of `need`, at least XC codes per slice
    const int64_t cap = (4 * kNumSMs) / slices;
    if (blocks > cap) blocks = cap;
  } else {
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
  }
  if (blocks < 1) blocks = 1;
  CCVSQ_CUDA(launch_pdl_if(!(rows && (pdl_off_mask() & 4)), search_exact_kernel, dim3((unsigned)blocks, slices), dim3(XT), smem, st, z, L, E, e_sq, K, rows, keys,
                        row_count, max_rows, idx));
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

}  // namespace ccvsq

using namespace ccvsq;

extern "C" int ccvsq_search_exact(const float* z, ccvsq_layout lay, const float* E, const float* e_sq,
                                  int K, int64_t* idx, void* stream) {
  CCVSQ_REQUIRE(z && E && e_sq && idx, CCVSQ_NULL_POINTER, "search_exact: null pointer");
  CCVSQ_REQUIRE(K > 0, CCVSQ_BAD_SHAPE, "search_exact: K=%d", K);
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  return launch_exact(z, L, E, e_sq, K, nullptr, nullptr, nullptr, 0, idx, (cudaStream_t)stream);
}

extern "C" int ccvsq_search_exact_rows(const float* z, ccvsq_layout lay, const float* E,
                                       const float* e_sq, int K, int64_t* fallback_ws,
                                       int32_t* fallback_count, int64_t fallback_capacity, int64_t* idx,
                                       void* stream) {
  CCVSQ_REQUIRE(z && E && e_sq && idx && fallback_ws && fallback_count, CCVSQ_NULL_POINTER,
                "search_exact_rows: null pointer");
  CCVSQ_REQUIRE(K > 0 && fallback_capacity >= 0, CCVSQ_BAD_SHAPE, "search_exact_rows: K=%d capacity=%lld",
                K, (long long)fallback_capacity);
  if (fallback_capacity == 0) return CCVSQ_OK;
  Lay L;
  if (int rc = make_lay(lay, &L)) return rc;
  return launch_exact(z, L, E, e_sq, K, fallback_ws,
                      reinterpret_cast<unsigned long long*>(fallback_ws + fallback_capacity), fallback_count,
                      fallback_capacity, idx, (cudaStream_t)stream);
}
