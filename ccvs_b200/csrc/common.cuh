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

// tensor-core screen: codes per accumulator tile, and the K-extension that carries the bias
// (the BF16 codebook shadow is [ccvsq_codebook_rows(K), D + SCREEN_EXT])
constexpr int SCREEN_BN = 96;
constexpr int SCREEN_EXT = 16;

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace ccvsq
