// peer_exchange.cu — the training collective of the EMA codebook update as two kernels over NVLink peer memory.
//
// What it replaces: one NCCL all-reduce of the packed statistics buffer [resid: K*D | counts: K] (1.05 MB at K = 1024,
// D = 256) followed by the EMA update kernels (ccvsq_ema_update_packed).  At that size the all-reduce is pure latency
// (80 us on 2 GPUs, 125 us exposed on 8 even when issued under the backward) and its host-side issue cost makes the
// training step host-bound.  Here every rank PUSHES its statistics into an inbox slot on every peer right after its
// forward (posted NVLink writes), and the EMA update reads the W local slots, sums them in rank order and rewrites the
// codebook — compute and collective in one kernel, no reduction tree, no second exchange:
//     publish:  stats -> peer[p].inbox[parity][rank]  for all p;   then  peer[p].flag[parity][rank] = seq + 1  (release.sys)
//     update :  wait until flag[parity][r] == seq + 1 for all r (acquire.sys);  S = sum_r inbox[parity][r]  (fixed order:
//               every rank computes bit-identical sums, so the replicated codebooks cannot drift apart);  EMA update;
//               the last CTA advances seq.
// Two parities make one barrier per step enough: rank p writes my slot of parity s&1 only after it has finished its own
// update s-1, which waited for my publish s-1, which I issued after my update s-2 — the last reader of that parity.
// seq lives in device memory, so the pair is replayable from a CUDA graph (no host-side step counter in the arguments).
//
// The exchange area is ordinary cudaMalloc memory shared with cudaIpc handles (one process per GPU, same node).
#include <string.h>
#include "common.cuh"

namespace ccvsq {

constexpr int PEER_MAX_WORLD = CCVSQ_PEER_MAX_WORLD;
constexpr uint32_t PEER_HEADER_BYTES = 1024;
// header words (uint32): [0] seq | [1] publish ticket | [2] update ticket | [3] fp32 sum(n_ema) | [4] error code
//                        [32 + parity*PEER_MAX_WORLD + r] flag of rank r
constexpr int HDR_SEQ = 0, HDR_TICKET_PUB = 1, HDR_TICKET_UPD = 2, HDR_NTOTAL = 3, HDR_FLAGS = 32;

struct PeerPtrs { uint8_t* p[PEER_MAX_WORLD]; };

__host__ __device__ inline size_t peer_stride_floats(int K, int D) { return ((size_t)K * D + K + 63) / 64 * 64; }
__host__ __device__ inline float* peer_inbox(uint8_t* area, int K, int D, int world, uint32_t parity, int r) {
  return reinterpret_cast<float*>(area + PEER_HEADER_BYTES) + ((size_t)parity * world + r) * peer_stride_floats(K, D);
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// every rank's statistics into every peer's inbox (own slot included), then the flags
__global__ void __launch_bounds__(256) peer_publish_kernel(const float* __restrict__ stats, const PeerPtrs peers, int K,
                                                           int D, int rank, int world) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(peers.p[rank]);
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(hdr + HDR_SEQ);   // (stable: only the update kernel's last CTA writes it)
  const uint32_t parity = seq & 1u;
  const size_t n = (size_t)K * D + K, n4 = n / 4;
  const float4* src4 = reinterpret_cast<const float4*>(stats);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  // (fully unrolled over the rank slots so that every peer base pointer is read straight from the parameter bank)
  const size_t slot_off = PEER_HEADER_BYTES + ((size_t)parity * world + rank) * peer_stride_floats(K, D) * sizeof(float);
  for (size_t i = tid; i < n4; i += nth) {
    const float4 v = src4[i];
#pragma unroll
    for (int p = 0; p < PEER_MAX_WORLD; ++p)
      if (p < world) *reinterpret_cast<float4*>(peers.p[p] + slot_off + i * 16) = v;
  }
  for (size_t i = n4 * 4 + tid; i < n; i += nth) {
    const float v = stats[i];
#pragma unroll
    for (int p = 0; p < PEER_MAX_WORLD; ++p)
      if (p < world) *reinterpret_cast<float*>(peers.p[p] + slot_off + i * 4) = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(hdr + HDR_TICKET_PUB, 1u);
    s_last = t == gridDim.x - 1 ? 1u : 0u;
    if (s_last) {                      // all CTAs of this grid have written and fenced
      hdr[HDR_TICKET_PUB] = 0;
      __threadfence_system();          // ONE system-scope fence, then the W flags in parallel (a release store per peer would
    }                                  // serialise W NVLink round trips: the last peer would see its flag ~15 us late)
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < world)
    st_relaxed_sys(reinterpret_cast<uint32_t*>(peers.p[threadIdx.x]) + HDR_FLAGS + parity * PEER_MAX_WORLD + rank, seq + 1u);
}

// wait (bounded) until every rank's statistics of this step have landed in the local inbox
__device__ __forceinline__ void peer_wait_all(const uint32_t* hdr, uint32_t parity, uint32_t want, int world) {
  const long long t0 = clock64();
  for (int r = 0; r < world; ++r) {
    const uint32_t* f = hdr + HDR_FLAGS + parity * PEER_MAX_WORLD + r;
    while (ld_acquire_sys(f) != want) {
      __nanosleep(200);
      if (clock64() - t0 > 120000000000ll) {      // ~60 s: a rank never published (crashed or out of step) — fail loudly
        printf("ccvsq peer exchange: rank %d never published step %u (flag %u)\n", r, want, ld_acquire_sys(f));
        __trap();
      }
    }
  }
}

// counts: n_ema <- decay n_ema + (1-decay) sum_r counts_r ; sum(n_ema) -> header   (one CTA, like ema_counts_kernel)
__global__ void __launch_bounds__(256) peer_ema_counts_kernel(float* __restrict__ n_ema, uint8_t* area, int K, int D, int world,
                                                              float decay) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(area);
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(hdr + HDR_SEQ), parity = seq & 1u;
  if (threadIdx.x == 0) peer_wait_all(hdr, parity, seq + 1u, world);
  __syncthreads();
  __shared__ float red[8];
  float acc = 0.f;
  const float* cnt0 = peer_inbox(area, K, D, world, parity, 0) + (size_t)K * D;
  const size_t stride = peer_stride_floats(K, D);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    // all W loads in flight, THEN the sum in rank order (a load-add loop is W dependent round trips to L2 / HBM)
    float v[PEER_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r) v[r] = r < world ? __ldcg(cnt0 + r * stride + k) : 0.f;
    float c = 0.f;
#pragma unroll
    for (int r = 0; r < PEER_MAX_WORLD; ++r) c += v[r];          // (+ 0.f for r >= world: exact)
    const float nv = decay * n_ema[k] + (1.f - decay) * c;
    n_ema[k] = nv;
    acc += nv;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    reinterpret_cast<float*>(hdr)[HDR_NTOTAL] = s;
  }
}

// sums: sum_ema <- decay sum_ema + (1-decay) (sum_r resid_r + sum_r counts_r * E) ; E <- sum_ema / smooth(n_ema)
__global__ void __launch_bounds__(256) peer_ema_embed_kernel(float* __restrict__ E, const float* __restrict__ n_ema,
                                                             float* __restrict__ sum_ema, uint8_t* area, int K, int D,
                                                             int world, float decay, float eps) {
  uint32_t* hdr = reinterpret_cast<uint32_t*>(area);
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(hdr + HDR_SEQ), parity = seq & 1u;
  const float nt = reinterpret_cast<const float*>(hdr)[HDR_NTOTAL];
  const size_t total4 = (size_t)K * D / 4;         // (D % 4 == 0 is required by the launcher)
  const float* in0 = peer_inbox(area, K, D, world, parity, 0);
  const size_t stride = peer_stride_floats(K, D);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i * 4 / D);
    // 8 ranks' loads in flight at a time, then their sums in rank order (fixed order: bit-identical on every rank)
    float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
    float c = 0.f;
    for (int r0 = 0; r0 < world; r0 += 8) {
      float4 v[8];
      float cv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const bool on = r0 + r < world;
        v[r] = on ? __ldcg(reinterpret_cast<const float4*>(in0 + (r0 + r) * stride) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        cv[r] = on ? __ldcg(in0 + (r0 + r) * stride + (size_t)K * D + k) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        rs.x += v[r].x; rs.y += v[r].y; rs.z += v[r].z; rs.w += v[r].w;
        c += cv[r];
      }
    }
    float4 e = reinterpret_cast<float4*>(E)[i], s = reinterpret_cast<float4*>(sum_ema)[i];
    s.x = decay * s.x + (1.f - decay) * (rs.x + c * e.x);
    s.y = decay * s.y + (1.f - decay) * (rs.y + c * e.y);
    s.z = decay * s.z + (1.f - decay) * (rs.z + c * e.z);
    s.w = decay * s.w + (1.f - decay) * (rs.w + c * e.w);
    reinterpret_cast<float4*>(sum_ema)[i] = s;
    const float n_smooth = (n_ema[k] + eps) / (nt + (float)K * eps) * nt;   // Laplace smoothing
    e.x = s.x / n_smooth; e.y = s.y / n_smooth; e.z = s.z / n_smooth; e.w = s.w / n_smooth;
    reinterpret_cast<float4*>(E)[i] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(hdr + HDR_TICKET_UPD, 1u);
    if (t == gridDim.x - 1) {          // every CTA has read seq (at its start) and its inbox slots: the step is over
      hdr[HDR_TICKET_UPD] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(hdr + HDR_SEQ) = seq + 1u;
    }
  }
}

}  // namespace ccvsq

using namespace ccvsq;

extern "C" uint64_t ccvsq_peer_exchange_bytes(int K, int D, int world) {
  if (K <= 0 || D <= 0 || world <= 0 || world > PEER_MAX_WORLD) return 0;
  return PEER_HEADER_BYTES + 2ull * world * peer_stride_floats(K, D) * sizeof(float);
}

extern "C" int ccvsq_peer_alloc(uint64_t bytes, void** ptr, void* handle64) {
  CCVSQ_REQUIRE(ptr && handle64 && bytes > 0, CCVSQ_NULL_POINTER, "peer_alloc: null pointer / zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == CCVSQ_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  CCVSQ_CUDA(cudaMalloc(&p, bytes));
  CCVSQ_CUDA(cudaMemset(p, 0, bytes));
  CCVSQ_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return CCVSQ_CUDA_ERROR;
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return CCVSQ_OK;
}

extern "C" int ccvsq_peer_open(const void* handle64, void** ptr) {
  CCVSQ_REQUIRE(ptr && handle64, CCVSQ_NULL_POINTER, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  CCVSQ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CCVSQ_OK;
}

extern "C" int ccvsq_peer_close(void* ptr) {
  if (ptr) CCVSQ_CUDA(cudaIpcCloseMemHandle(ptr));
  return CCVSQ_OK;
}

extern "C" int ccvsq_peer_free(void* ptr) {
  if (ptr) CCVSQ_CUDA(cudaFree(ptr));
  return CCVSQ_OK;
}

static int peer_args_ok(void* const* areas, int K, int D, int rank, int world, const char* what) {
  CCVSQ_REQUIRE(areas, CCVSQ_NULL_POINTER, "%s: null pointer", what);
  CCVSQ_REQUIRE(K > 0 && D > 0 && D % 4 == 0, CCVSQ_BAD_SHAPE, "%s: K=%d D=%d (D must be a multiple of 4)", what, K, D);
  CCVSQ_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, CCVSQ_BAD_SHAPE,
                "%s: rank %d of %d (at most %d ranks)", what, rank, world, PEER_MAX_WORLD);
  for (int p = 0; p < world; ++p) CCVSQ_REQUIRE(areas[p], CCVSQ_NULL_POINTER, "%s: exchange area of rank %d is null", what, p);
  return CCVSQ_OK;
}

extern "C" int ccvsq_peer_publish(const float* stats, int K, int D, void* const* areas, int rank, int world, void* stream) {
  CCVSQ_REQUIRE(stats, CCVSQ_NULL_POINTER, "peer_publish: null pointer");
  if (int rc = peer_args_ok(areas, K, D, rank, world, "peer_publish")) return rc;
  CCVSQ_REQUIRE(((uintptr_t)stats & 15) == 0, CCVSQ_MISALIGNED, "peer_publish: the statistics buffer must be 16-byte aligned");
  PeerPtrs pp = {};
  for (int p = 0; p < world; ++p) pp.p[p] = (uint8_t*)areas[p];
  // A small grid on purpose: the pushes run NEXT TO the caller's backward pass (side stream) and have its whole duration
  // to finish; 32 CTAs move W x 1 MB in ~10 us without taking an SM slot on most of the chip from the HBM-bound kernel.
  const size_t n4 = ((size_t)K * D + K) / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 32) blocks = 32;
  if (blocks < 1) blocks = 1;
  peer_publish_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(stats, pp, K, D, rank, world);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_peer_ema_update(float* E, float* n_ema, float* sum_ema, void* area, int K, int D, int world,
                                     float decay, float eps, void* stream) {
  CCVSQ_REQUIRE(E && n_ema && sum_ema && area, CCVSQ_NULL_POINTER, "peer_ema_update: null pointer");
  CCVSQ_REQUIRE(K > 0 && D > 0 && D % 4 == 0, CCVSQ_BAD_SHAPE, "peer_ema_update: K=%d D=%d (D must be a multiple of 4)", K, D);
  CCVSQ_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD, CCVSQ_BAD_SHAPE, "peer_ema_update: world=%d", world);
  CCVSQ_REQUIRE((((uintptr_t)E | (uintptr_t)sum_ema) & 15) == 0, CCVSQ_MISALIGNED, "peer_ema_update: E and sum_ema must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  peer_ema_counts_kernel<<<1, 256, 0, st>>>(n_ema, (uint8_t*)area, K, D, world, decay);
  CCVSQ_LAUNCH_CHECK();
  const size_t total4 = (size_t)K * D / 4;
  int blocks = (int)((total4 + 255) / 256);
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
  peer_ema_embed_kernel<<<blocks, 256, 0, st>>>(E, n_ema, sum_ema, (uint8_t*)area, K, D, world, decay, eps);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}
