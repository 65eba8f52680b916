// common.cuh — shared device/host helpers for libccvsq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ccvsq.h"

namespace ccvsq {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message, negative status codes; nothing throws across the ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define CCVSQ_REQUIRE(cond, code, ...)        \
  do {                                        \
    if (!(cond)) {                            \
      ::ccvsq::set_error(__VA_ARGS__);        \
      return (code);                          \
    }                                         \
  } while (0)

#define CCVSQ_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::ccvsq::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return CCVSQ_CUDA_ERROR;                                                            \
    }                                                                                     \
  } while (0)

#define CCVSQ_LAUNCH_CHECK()                                                                 \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::ccvsq::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                          \
      return CCVSQ_CUDA_ERROR;                                                               \
    }                                                                                        \
  } while (0)

// ---------------------------------------------------------------------------------------------
// latent layout (see include/ccvsq.h): element (g,c,s) at ((g*C + c)*S + s)
// ---------------------------------------------------------------------------------------------
struct Lay {
  int64_t G;
  int64_t P;  // positions = G*S
  int64_t N;  // latent rows = P*mult
  int C, S, mult, D;
};

inline int make_lay(const ccvsq_layout& l, Lay* out) {
  CCVSQ_REQUIRE(l.G > 0 && l.C > 0 && l.S > 0 && l.mult > 0, CCVSQ_BAD_SHAPE,
                "layout: G=%lld C=%d S=%d mult=%d must all be positive", (long long)l.G, l.C, l.S,
                l.mult);
  CCVSQ_REQUIRE(l.C % l.mult == 0, CCVSQ_BAD_SHAPE, "layout: C=%d not divisible by mult=%d", l.C,
                l.mult);
  out->G = l.G;
  out->C = l.C;
  out->S = l.S;
  out->mult = l.mult;
  out->D = l.C / l.mult;
  out->P = l.G * (int64_t)l.S;
  out->N = out->P * l.mult;
  return CCVSQ_OK;
}

// position p -> offset of (g, c=0, s)
__device__ __forceinline__ int64_t pos_base(const Lay& L, int64_t p) {
  if (L.S == 1) return p * L.C;
  int64_t g = p / L.S;
  int s = (int)(p - g * L.S);
  return g * (int64_t)L.C * L.S + s;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Opt a kernel in to > 48 KiB of dynamic shared memory.  The attribute is per (kernel, device) and
// sticky, so it is raised once instead of on every launch (the table lives in api.cu).
int enable_smem_impl(const void* kern, size_t bytes);
template <typename Kern>
inline int enable_smem(Kern kern, size_t bytes) {
  return enable_smem_impl(reinterpret_cast<const void*>(kern), bytes);
}

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): the kernels of one forward run back to back on one stream and each
// boundary costs a few microseconds of launch latency + prologue — a quarter of the whole call at the 1k-5k
// latents the reference trainer issues.  Every kernel of the chain calls pdl_launch_dependents() first (the next
// kernel's CTAs may become resident as this one's drain) and pdl_wait() before it touches anything a predecessor
// wrote (blocks until the preceding grid has completed and its memory is visible).  Both are no-ops for a kernel
// launched the ordinary way.  EVERY thread of a PDL-launched kernel must pass pdl_wait() before it exits, so that
// "grid complete" keeps implying "all earlier grids complete" along the chain.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_off_mask();    // api.cu: CCVSQ_NO_PDL_KERNELS=<mask> (diagnosis: 1 screen, 2 rescore, 4 exact fallback)
bool pdl_enabled();   // api.cu: CCVSQ_NO_PDL=1 switches the launch attribute off (A/B runs)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool on, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (on && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// tensor-core screen: codes per accumulator tile, and the K-extension that carries the bias
// (the BF16 codebook shadow is [ccvsq_codebook_rows(K), D + SCREEN_EXT])
constexpr int SCREEN_BN = 96;
constexpr int SCREEN_EXT = 16;

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// HBM-bound stream kernels: argument block shared by the generic (stream_kernels.cu) and the
// 128-bit fast paths (stream_fast.cu)
// ---------------------------------------------------------------------------------------------
enum { MODE_ASSIGN = 0, MODE_BACKWARD = 1, MODE_GATHER = 2, MODE_STATS = 3,
       MODE_ASSIGN_STATS = 4 };   // assign with the per-code residual sums on the same pass (internal to the fast kernels)
#define IS_ASSIGN(MODE) ((MODE) == MODE_ASSIGN || (MODE) == MODE_ASSIGN_STATS)

struct FinArgs {          // loss / perplexity folded into the assign kernel (last CTA); ticket NULL = off
  int32_t* ticket;        // zeroed by the caller
  float* loss;
  float* perplexity;
  float* counts_f32;      // optional [K]: the per-code counts as FP32 (tail of a packed all-reduce buffer)
  double M, N;
  float beta;
};

struct StreamArgs {
  const float* x;         // latents (assign / backward / stats); unused by gather
  const float* g;         // backward: upstream gradient of z_q (may be NULL)
  const float* E;         // codebook [K, D]
  const int64_t* idx;     // codes, one per latent row
  float* out;             // z_q / dz / decoded latents (may be NULL)
  double* sq_err;         // assign
  int32_t* counts;        // assign / stats
  float* resid;           // stats, or backward with the scatter-reduce fused in
  const float* g_loss;    // backward: device scalar
  float coef_scale;       // backward: 2 / M
  float sub;              // stats: resid += x - sub * E[idx]
  int32_t* err_flag;      // gather
  const float* pos;       // gather (row-major only): rows added to the gathered rows, out[n] = E[idx[n]] + pos[n % pos_period]
  int64_t pos_period;
  int K;
  bool x_stable;          // assign only: x was complete before this launch chain began (the composite forward starts with
                          // ordinary launches), so its loads may be issued before pdl_wait(); false = wait first
  FinArgs fin;
};

// ccvsq_screen for the composite forward (screen_sm100.cu): z is known to be complete before the chain starts
int screen_launch_stable_z(const float* z, ccvsq_layout lay, const void* E_bf16, const float* e_max, int K,
                           float margin_tau, int n_cand, int64_t* idx, int32_t* queue_count, int32_t* queue_rows,
                           int32_t* queue_cand, uint8_t* queue_flags, void* stream);

bool stream_fast_supported(const StreamArgs& a, const Lay& L);
int stream_fast_launch(int mode, const StreamArgs& a, const Lay& L, cudaStream_t st);
// generic launches + finalize (stream_kernels.cu), used by the ABI entry points and the composites
int stream_launch(int mode, const StreamArgs& a, const Lay& L, cudaStream_t st);
int prepare_codebook_launch(const float* E, int K, int D, float* e_sq, void* E_bf16, float* e_max, cudaStream_t st);

}  // namespace ccvsq
