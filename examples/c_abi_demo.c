/* c_abi_demo.c — the quantizer hot path driven from plain C through include/ccvsq.h: no Python, no torch.
 *
 * What a non-Python host (or another language's FFI) does: own every device buffer, fill a
 * ccvsq_forward_args, make ONE call per forward, one per decode.  The demo quantizes a small
 * [G, C, h, w] latent tensor (the layout of quantize.py:40-42, never transposed), decodes the indices
 * again and checks both against a brute-force double-precision search on the host.
 *
 *   build:  make -C ccvs_b200/csrc demo        (gcc + libcudart; links ../lib/libccvsq.so)
 *   run:    build/c_abi_demo                   (needs a B200; exit code 0 = all checks passed)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "../include/ccvsq.h"

#define CHECK_CUDA(x)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                      \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)
#define CHECK_VQ(x)                                                                        \
  do {                                                                                     \
    int rc_ = (x);                                                                         \
    if (rc_ != CCVSQ_OK) {                                                                 \
      fprintf(stderr, "%s failed: %d (%s)\n", #x, rc_, ccvsq_last_error());                \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

static uint32_t rng_state = 12345u;
static float rnd(void) { /* xorshift, roughly N(0,1) by summing uniforms */
  float s = 0.f;
  for (int i = 0; i < 4; ++i) {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 17;
    rng_state ^= rng_state << 5;
    s += (float)(rng_state & 0xFFFFFF) / 16777216.0f;
  }
  return (s - 2.0f) * 1.7320508f;
}

int main(void) {
  const int G = 8, C = 256, H = 16, W = 16, K = 1024, D = C, S = H * W;
  const int64_t N = (int64_t)G * S;
  const size_t zn = (size_t)G * C * S;
  printf("libccvsq version %d; %lld latents, K=%d, D=%d\n", ccvsq_version(), (long long)N, K, D);
  if (ccvsq_version() != CCVSQ_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 1; }

  /* host data: codebook ~ N(0,1); latent (g, :, s) = a random code + noise, stored channel-major [G, C, S] */
  float* hE = (float*)malloc((size_t)K * D * sizeof(float));
  float* hz = (float*)malloc(zn * sizeof(float));
  for (size_t i = 0; i < (size_t)K * D; ++i) hE[i] = rnd();
  for (int g = 0; g < G; ++g)
    for (int s = 0; s < S; ++s) {
      rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5;
      const int k = (int)(rng_state % (uint32_t)K);
      for (int c = 0; c < C; ++c) hz[((size_t)g * C + c) * S + s] = hE[(size_t)k * D + c] + 0.5f * rnd();
    }

  cudaStream_t st;
  CHECK_CUDA(cudaStreamCreate(&st));
  float *dz, *dE, *dzq, *dloss, *dperp, *ddec;
  int64_t* didx;
  int32_t* dhdr;
  void* dws;
  const ccvsq_layout lay = {G, C, S, 1};
  const uint64_t ws_bytes = ccvsq_forward_workspace_bytes(N, K, D, CCVSQ_SEARCH_AUTO, 4, 1);
  CHECK_CUDA(cudaMalloc((void**)&dz, zn * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dzq, zn * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&ddec, zn * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dE, (size_t)K * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&didx, (size_t)N * sizeof(int64_t)));
  CHECK_CUDA(cudaMalloc((void**)&dhdr, (size_t)(CCVSQ_HEADER_INTS + K) * sizeof(int32_t)));
  CHECK_CUDA(cudaMalloc((void**)&dloss, sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dperp, sizeof(float)));
  CHECK_CUDA(cudaMalloc(&dws, ws_bytes));
  CHECK_CUDA(cudaMemcpyAsync(dz, hz, zn * sizeof(float), cudaMemcpyHostToDevice, st));
  CHECK_CUDA(cudaMemcpyAsync(dE, hE, (size_t)K * D * sizeof(float), cudaMemcpyHostToDevice, st));

  /* forward: quantize.py:32-74 in one call */
  ccvsq_forward_args a;
  memset(&a, 0, sizeof(a));
  a.struct_size = (uint32_t)sizeof(a);   /* lets the library reject a binding built against another header */
  a.z = dz; a.lay = lay; a.E = dE; a.K = K; a.beta = 0.25f;
  a.search_mode = CCVSQ_SEARCH_AUTO; a.n_cand = 4; a.margin_tau = 1.0f; a.exact_fallback = 1;
  a.header = dhdr; a.workspace = dws; a.workspace_bytes = ws_bytes;
  a.idx = didx; a.zq = dzq; a.loss = dloss; a.perplexity = dperp;
  CHECK_VQ(ccvsq_quantize_forward(&a, st));
  /* decode: embed_code written channel-major, the decoder's layout (quantize.py:76-83 + caller transpose) */
  CHECK_VQ(ccvsq_gather(didx, dE, K, lay, ddec, NULL, st));

  int64_t* hidx = (int64_t*)malloc((size_t)N * sizeof(int64_t));
  float* hzq = (float*)malloc(zn * sizeof(float));
  float* hdec = (float*)malloc(zn * sizeof(float));
  float hloss = 0.f, hperp = 0.f;
  CHECK_CUDA(cudaMemcpyAsync(hidx, didx, (size_t)N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaMemcpyAsync(hzq, dzq, zn * sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaMemcpyAsync(hdec, ddec, zn * sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaMemcpyAsync(&hloss, dloss, sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaMemcpyAsync(&hperp, dperp, sizeof(float), cudaMemcpyDeviceToHost, st));
  CHECK_CUDA(cudaStreamSynchronize(st));

  /* host check: brute-force nearest code in double precision (row n = g*S + s, quantize.py:40-42) */
  int64_t bad_idx = 0, bad_val = 0;
  double sq = 0.0;
  double* zrow = (double*)malloc((size_t)D * sizeof(double));
  for (int64_t n = 0; n < N; ++n) {
    const int g = (int)(n / S), s = (int)(n % S);
    for (int c = 0; c < C; ++c) zrow[c] = hz[((size_t)g * C + c) * S + s];
    double best = INFINITY, ours = 0.0;
    int bk = -1;
    for (int k = 0; k < K; ++k) {
      double d = 0.0;
      for (int c = 0; c < C; ++c) { const double t = zrow[c] - hE[(size_t)k * D + c]; d += t * t; }
      if (d < best) { best = d; bk = k; }
      if (k == (int)hidx[n]) ours = d;
    }
    if (hidx[n] != bk && fabs(ours - best) > 1e-6 * best) ++bad_idx;   /* near-ties may differ */
    sq += ours;
    for (int c = 0; c < C; ++c) {
      const size_t o = ((size_t)g * C + c) * S + s;
      const float e = hE[(size_t)hidx[n] * D + c], zz = hz[o];
      const float diff = e - zz;                       /* fl(E[idx] - z)          */
      if (hzq[o] != zz + diff) ++bad_val;              /* fl(z + fl(E[idx] - z)), quantize.py:64 */
      if (hdec[o] != e) ++bad_val;                     /* decode is a pure gather */
    }
  }
  const double ref_loss = 1.25 * sq / (double)zn;      /* (1 + beta) * mse, quantize.py:60-61 */
  const int loss_ok = fabs(hloss - ref_loss) <= 1e-5 * ref_loss;
  printf("index mismatches outside near-ties: %lld, value mismatches: %lld, loss %.6f (host %.6f), perplexity %.2f\n",
         (long long)bad_idx, (long long)bad_val, hloss, ref_loss, hperp);
  const int ok = bad_idx == 0 && bad_val == 0 && loss_ok && hperp > 1.f && hperp <= (float)K;
  printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 4;
}
