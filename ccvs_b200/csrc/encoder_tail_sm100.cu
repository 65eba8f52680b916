// encoder_tail_sm100.cu — the encoder's last block, the 1x1 convolution that PRODUCES the latents z:
//
//   reference: models/skip_vid_generator/models/skip_autoencoder.py:331  blocks.append(ConvLayer(block_out, z_size, 1))
//              :40-58   EqualConv2d.forward:  F.conv2d(x, weight * (1 / sqrt(C_in * k^2)), bias)          (k = 1)
//              :66-101  ConvLayer: ... + nn.LeakyReLU(negative_slope=0.1)
//              :346-349 out = blocks[-1](out); if normalize_out: out = out / ||out||_2 over the channel dim
//
// A 1x1 convolution is a GEMM over positions:  z[g, o, s] = lrelu( sum_c Ws[o, c] x[g, c, s] + b[o] ),  Ws = W * scale.
// The quantizer downstream must reproduce the FP32 argmin, so the product has to be FP32-accurate; the tensor cores
// take BF16.  Both operands are therefore split into THREE BF16 terms (x = x0 + x1 + x2 exactly to 2^-24 relative),
// and the six products whose magnitude is above 2^-24 of the result are accumulated in FP32 tensor memory:
//     x0 w0 + x0 w1 + x1 w0 + x0 w2 + x1 w1 + x2 w0
// — six tcgen05.mma per 16-element K step (the "BF16x6" scheme), ~3x the throughput of the FP32 CUDA-core GEMM the
// reference's F.conv2d runs.
//
// Structure (2-CTA clusters, cta_group::2, M = 256 positions x N <= 256 output channels per tile; 384 threads per CTA):
//   warp 0      TMA producer: this CTA's half of the three weight-term tiles of a K block (64 channels, 128B swizzle)
//   warp 1      MMA issuer (leader CTA): 4 K steps x 6 term pairs per K block, SS form
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld, + bias, LeakyReLU, coalesced stores straight into the NCHW output
//   warps 8-11  A loaders: FP32 activations straight from the NCHW input (coalesced along the spatial index), split into
//               three BF16 terms, written into shared memory in the tensor core's K-major 128B-swizzled layout
// Two pipeline stages of (A terms | B terms) = 2 x 96 KiB of shared memory.  Tensor memory holds TWO accumulators of
// the same tile: the leading product x0 w0 goes to one, the five correction products (2^-8 ... 2^-16 of it) to the other,
// and the epilogue adds them in FP32.  The tensor core's accumulator keeps only ~FP32 alignment per K step: mixed into
// one chain, the small terms lose their low bits against the large partial sum 6 x C_in/16 times (measured: 6.5e-7 of
// the accumulated magnitude); kept apart, each chain adds terms of one magnitude class.  (One tile in flight instead
// of two: the epilogue of a tile, ~10 % of its MMA time, is exposed — accuracy first, parity is the gate.)
#include <cuda.h>
#include <float.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace ccvsq {

constexpr int ET_BM = 128;          // positions per CTA (TMEM lanes)
constexpr int ET_BK = 64;           // input channels per pipeline stage (one 128-byte swizzle atom of BF16)
constexpr int ET_THREADS = 384;
constexpr int ET_STAGES = 2;
constexpr uint32_t ET_TERM_A = ET_BM * 128;          // bytes of one A term tile (128 rows x 64 BF16)

struct EtSmem { uint32_t a, b, bias, bars, total, term_b, stage; };
__host__ __device__ inline EtSmem et_smem_layout(int bn) {
  EtSmem s;
  s.term_b = (uint32_t)(bn / 2) * 128;               // this CTA's half of the output channels x 64 BF16
  s.stage = 3 * ET_TERM_A + 3 * s.term_b;
  uint32_t off = 0;
  s.a = off;    off += ET_STAGES * s.stage;          // per stage: [3 A terms][3 B terms], each 1024-byte aligned
  s.b = 0;                                           // (B terms follow the A terms inside a stage)
  s.bias = off; off += 256 * 4;
  s.bars = off; off += 256;
  s.total = off;
  return s;
}

// three-term BF16 split of an FP32 value: v = t0 + t1 + t2 up to 2^-24 |v|
__device__ __forceinline__ void split3(float v, __nv_bfloat16& t0, __nv_bfloat16& t1, __nv_bfloat16& t2) {
  t0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(t0);
  t1 = __float2bfloat16_rn(r1);
  t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
}

// weights -> three BF16 term matrices [3][C_out][C_in] of Ws = fl(W * scale) (the reference multiplies in FP32 first)
__global__ void et_prepare_kernel(const float* __restrict__ W, int64_t n, float scale, __nv_bfloat16* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 a, b, c;
    split3(__fmul_rn(W[i], scale), a, b, c);
    out[i] = a; out[n + i] = b; out[2 * n + i] = c;
  }
}

__global__ void __launch_bounds__(ET_THREADS, 1)
encoder_tail_kernel(const __grid_constant__ CUtensorMap map_w, const float* __restrict__ x, const float* __restrict__ bias,
                    float* __restrict__ z, int64_t P, int C_in, int C_out, int S, int BN, float slope, int num_row_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const EtSmem lay = et_smem_layout(BN);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)blockIdx.x / 2, num_pairs = (int)gridDim.x / 2;
  const int n_tiles = C_out / BN, k_blocks = C_in / ET_BK;
  const int total_tiles = num_row_tiles * n_tiles;     // tile t: row tile t / n_tiles, channel tile t % n_tiles
  const int ROWS_B = BN / 2;

  const uint32_t bar0 = smem_base + lay.bars;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };               // leader: A of both CTAs written + B bytes landed
  auto empty_bar = [&](int s) { return bar0 + 8u * (ET_STAGES + s); }; // MMAs of the stage retired (multicast commit)
  auto tmem_full = [&](int b) { return bar0 + 8u * (2 * ET_STAGES + b); };
  auto tmem_empty = [&](int b) { return bar0 + 8u * (2 * ET_STAGES + 2 + b); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + lay.bars + 8u * (2 * ET_STAGES + 4));
  float* bias_s = reinterpret_cast<float*>(smem + lay.bias);

  if ((smem_base & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_w);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < ET_STAGES; ++s) {
      mbar_init(full_bar(s), 4 * 2 + 1);      // 4 loader warps of each CTA + the TMA producer's expect_tx arrive
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4 * 2); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2>(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // =========================== TMA producer: weight terms ===========================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t fb0 = mapa(full_bar(0), 0);
    for (int t = pair; t < total_tiles; t += num_pairs) {
      const int n0 = (t % n_tiles) * BN + (int)rank * ROWS_B;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 3 * lay.term_b * 2);
          const uint32_t dst = smem_base + lay.a + (uint32_t)stage * lay.stage + 3 * ET_TERM_A;
          for (int term = 0; term < 3; ++term)
            tma_load_2d<2>(dst + term * lay.term_b, &map_w, fb0 + 8u * stage, kb * ET_BK, term * C_out + n0);
        }
        __syncwarp();
        if (++stage == ET_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(ET_BM * 2, BN);
      int stage = 0;
      uint32_t phase = 0, b = 0, b_phase = 0;
      for (int t = pair; t < total_tiles; t += num_pairs) {
        mbar_wait(tmem_empty(b), b_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base, c_tmem = tmem_base + 256;      // leading products | correction products
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + lay.a + (uint32_t)stage * lay.stage;
            const uint32_t sb = sa + 3 * ET_TERM_A;
            uint32_t da[3], db[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) { da[i] = desc_lo(sa + i * ET_TERM_A); db[i] = desc_lo(sb + i * lay.term_b); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // 4 K steps of 16 channels; term pairs in decreasing magnitude
              const uint32_t o = (uint32_t)k * 2;
              const uint32_t acc = (kb | k) ? 1u : 0u;
              umma_ss<2>(d_tmem, da[0] + o, DESC_HI_SW128, db[0] + o, DESC_HI_SW128, idesc, acc);
              umma_ss<2>(c_tmem, da[0] + o, DESC_HI_SW128, db[1] + o, DESC_HI_SW128, idesc, acc);
              umma_ss<2>(c_tmem, da[1] + o, DESC_HI_SW128, db[0] + o, DESC_HI_SW128, idesc, 1u);
              umma_ss<2>(c_tmem, da[0] + o, DESC_HI_SW128, db[2] + o, DESC_HI_SW128, idesc, 1u);
              umma_ss<2>(c_tmem, da[1] + o, DESC_HI_SW128, db[1] + o, DESC_HI_SW128, idesc, 1u);
              umma_ss<2>(c_tmem, da[2] + o, DESC_HI_SW128, db[0] + o, DESC_HI_SW128, idesc, 1u);
            }
            umma_commit<2>(empty_bar(stage));
            if (kb == k_blocks - 1) umma_commit<2>(tmem_full(b));
          }
          __syncwarp();
          if (++stage == ET_STAGES) { stage = 0; phase ^= 1; }
        }
        b_phase ^= 1;                              // (one accumulator pair: b stays 0)
      }
    }
  } else if (warp >= 8) {
    // =========================== A loaders: FP32 NCHW -> three BF16 terms in swizzled shared memory ===========
    const int r = (warp - 8) * 32 + lane;                 // row (position) of this CTA's tile
    const uint32_t fb0 = mapa(full_bar(0), 0);
    int stage = 0;
    uint32_t phase = 0;
    // K-major 128B swizzle: row r, 16-byte chunk j (8 channels) at (r / 8) * 1024 + (r % 8) * 128 + ((j ^ (r % 8)) * 16)
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    for (int t = pair; t < total_tiles; t += num_pairs) {
      const int64_t p = ((int64_t)(t / n_tiles) * 2 + rank) * ET_BM + r;
      const bool valid = p < P;
      const float* xr = x;
      if (valid) {
        const int64_t g = p / S;
        xr = x + (g * C_in) * (int64_t)S + (p - g * S);
      }
      for (int kb = 0; kb < k_blocks; ++kb) {
        float v[ET_BK];
#pragma unroll
        for (int c = 0; c < ET_BK; ++c) v[c] = valid ? __ldg(xr + (int64_t)(kb * ET_BK + c) * S) : 0.f;
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + lay.a + (uint32_t)stage * lay.stage + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t w0[4], w1[4], w2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __nv_bfloat16 a0, a1, a2, b0, b1, b2;
            split3(v[8 * j + 2 * q], a0, a1, a2);
            split3(v[8 * j + 2 * q + 1], b0, b1, b2);
            w0[q] = (uint32_t)__bfloat16_as_ushort(a0) | ((uint32_t)__bfloat16_as_ushort(b0) << 16);
            w1[q] = (uint32_t)__bfloat16_as_ushort(a1) | ((uint32_t)__bfloat16_as_ushort(b1) << 16);
            w2[q] = (uint32_t)__bfloat16_as_ushort(a2) | ((uint32_t)__bfloat16_as_ushort(b2) << 16);
          }
          const uint32_t co = (uint32_t)((j ^ (r & 7)) * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + co), "r"(w0[0]), "r"(w0[1]), "r"(w0[2]), "r"(w0[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + ET_TERM_A + co), "r"(w1[0]), "r"(w1[1]), "r"(w1[2]), "r"(w1[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa + 2 * ET_TERM_A + co), "r"(w2[0]), "r"(w2[1]), "r"(w2[2]), "r"(w2[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> the MMA's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(fb0 + 8u * stage);
        if (++stage == ET_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue: + bias, LeakyReLU, NCHW store ===========================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t te_bar = mapa(tmem_empty(0), 0);
    uint32_t b = 0, b_phase = 0;
    int last_n_tile = -1;
    for (int t = pair; t < total_tiles; t += num_pairs) {
      const int nt = t % n_tiles;
      const int64_t p = ((int64_t)(t / n_tiles) * 2 + rank) * ET_BM + r;
      const bool valid = p < P;
      float* zr = z;
      if (valid) {
        const int64_t g = p / S;
        zr = z + (g * C_out + (int64_t)nt * BN) * (int64_t)S + (p - g * S);
      }
      if (nt != last_n_tile) {                 // this tile's bias slice -> shared memory (the four epilogue warps only)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = threadIdx.x - 128; i < BN; i += 128) bias_s[i] = bias ? __ldg(bias + nt * BN + i) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        last_n_tile = nt;
      }
      mbar_wait(tmem_full(b), b_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < BN; c += 32) {
        uint32_t ra[32], rc[32];
        tmem_ld32(taddr + c, ra);
        tmem_ld32(taddr + 256 + c, rc);
        tmem_ld_wait();
        if (c + 32 >= BN) {                    // last chunk in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(te_bar + 8u * b);
        }
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (c + i >= BN) break;                // (BN is a multiple of 16, not necessarily of 32)
            const float a = __fadd_rn(__fadd_rn(__uint_as_float(ra[i]), __uint_as_float(rc[i])), bias_s[c + i]);
            zr[(int64_t)(c + i) * S] = a > 0.f ? a : __fmul_rn(a, slope);       // 32 lanes = 32 consecutive positions: coalesced
          }
        }
      }
      b_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

// in-place  z[g, :, s] /= ||z[g, :, s]||_2  (skip_autoencoder.py:348-349); one thread per position, coalesced along s
__global__ void et_normalize_kernel(float* __restrict__ z, int64_t P, int C, int S) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int64_t g = p / S;
  float* zr = z + g * (int64_t)C * S + (p - g * S);
  float ss = 0.f;
  for (int c = 0; c < C; ++c) { const float v = zr[(int64_t)c * S]; ss = fmaf(v, v, ss); }
  const float n = sqrtf(ss);
  for (int c = 0; c < C; ++c) zr[(int64_t)c * S] = __fdiv_rn(zr[(int64_t)c * S], n);
}

}  // namespace ccvsq

using namespace ccvsq;

extern "C" int ccvsq_encoder_tail_prepare(const float* W, int C_out, int C_in, float scale, void* W_terms, void* stream) {
  CCVSQ_REQUIRE(W && W_terms, CCVSQ_NULL_POINTER, "encoder_tail_prepare: null pointer");
  CCVSQ_REQUIRE(C_out > 0 && C_in > 0, CCVSQ_BAD_SHAPE, "encoder_tail_prepare: C_out=%d C_in=%d", C_out, C_in);
  const int64_t n = (int64_t)C_out * C_in;
  int blocks = cdiv(n, 256 * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  et_prepare_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, n, scale, (__nv_bfloat16*)W_terms);
  CCVSQ_LAUNCH_CHECK();
  return CCVSQ_OK;
}

extern "C" int ccvsq_encoder_tail(const float* x, int64_t G, int C_in, int S, const void* W_terms, const float* bias,
                                  int C_out, float negative_slope, int normalize, float* z, void* stream) {
  CCVSQ_REQUIRE(x && W_terms && z, CCVSQ_NULL_POINTER, "encoder_tail: null pointer");
  CCVSQ_REQUIRE(G > 0 && S > 0, CCVSQ_BAD_SHAPE, "encoder_tail: G=%lld S=%d", (long long)G, S);
  CCVSQ_REQUIRE(C_in >= 64 && C_in % 64 == 0, CCVSQ_UNSUPPORTED, "encoder_tail: C_in=%d (need a multiple of 64)", C_in);
  CCVSQ_REQUIRE(C_out >= 16 && C_out % 16 == 0 && (C_out <= 256 || C_out % 256 == 0), CCVSQ_UNSUPPORTED,
                "encoder_tail: C_out=%d (need a multiple of 16 up to 256, or a multiple of 256)", C_out);
  CCVSQ_REQUIRE((((uintptr_t)x | (uintptr_t)W_terms | (uintptr_t)z) & 15) == 0, CCVSQ_MISALIGNED,
                "encoder_tail: x, the weight terms and z must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int BN = C_out <= 256 ? C_out : 256;
  const int64_t P = G * (int64_t)S;
  const int64_t row_tiles = (P + 2 * ET_BM - 1) / (2 * ET_BM);
  CCVSQ_REQUIRE(row_tiles * (C_out / BN) < (1ll << 30), CCVSQ_BAD_SHAPE, "encoder_tail: too many tiles");
  const EtSmem lay = et_smem_layout(BN);
  EncodeTiledFn enc;
  if (int rc = get_encode_fn(&enc)) return rc;
  CUtensorMap mw;   // weight terms: [3 * C_out rows, C_in cols] bf16; box = 64 channels x (BN / 2) output channels
  if (int rc = make_map(enc, &mw, W_terms, 3ll * C_out, C_in, ET_BK, BN / 2, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = enable_smem(encoder_tail_kernel, lay.total)) return rc;
  const int64_t total_tiles = row_tiles * (C_out / BN);
  const int pairs = (int)(total_tiles < kNumSMs / 2 ? total_tiles : kNumSMs / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(ET_THREADS);
  cfg.dynamicSmemBytes = lay.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CCVSQ_CUDA(cudaLaunchKernelEx(&cfg, encoder_tail_kernel, mw, x, bias, z, P, C_in, C_out, S, BN, negative_slope, (int)row_tiles));
  if (normalize) {
    et_normalize_kernel<<<cdiv(P, 256), 256, 0, st>>>(z, P, C_out, S);
    CCVSQ_LAUNCH_CHECK();
  }
  return CCVSQ_OK;
}
