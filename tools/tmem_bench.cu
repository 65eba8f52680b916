// tmem_bench.cu — how fast can one SM read tensor memory?  (the floor under the screen kernel's epilogue)
// Each CTA allocates 512 TMEM columns; `nwarps` warps (warp w reads lane quadrant w % 4) issue `iters` rounds of
// tcgen05.ld.32x32b.xN over the whole column range and time themselves with clock64().  Reports bytes per SM clock.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tmem_bench.cu -o tools/bin/tmem_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: ld.x32 + wait each (latency-exposed); mode 1: two ld.x32 in flight per wait; mode 2: ld.x16, four in flight
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int nwarps, int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    // warps sharing a quadrant read different halves of the column range
    const int share = (nwarps + 3) / 4, part = warp >> 2;
    const int c_lo = 512 / share * part, c_hi = 512 / share * (part + 1);
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t r[32];
          tmem_ld32(base + c, r);
          tmem_ld_wait();
          acc ^= r[0] ^ r[31];
        }
      } else if (MODE == 1) {
        for (int c = c_lo; c < c_hi; c += 64) {
          uint32_t ra[32], rb[32];
          tmem_ld32(base + c, ra);
          tmem_ld32(base + c + 32, rb);
          tmem_ld_wait();
          acc ^= ra[0] ^ rb[31];
        }
      } else {
        for (int c = c_lo; c < c_hi; c += 64) {
          uint32_t r0[16], r1[16], r2[16], r3[16];
          tmem_ld16(base + c, r0);
          tmem_ld16(base + c + 16, r1);
          tmem_ld16(base + c + 32, r2);
          tmem_ld16(base + c + 48, r3);
          tmem_ld_wait();
          acc ^= r0[0] ^ r1[1] ^ r2[2] ^ r3[3];
        }
      }
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < nwarps) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

// ---- the same reads while one thread keeps the tensor core busy (TS-form tcgen05.mma: A from tensor memory, B from
// shared memory, FP32 accumulators in tensor memory — the screen kernel's instruction; operand contents are
// irrelevant).  Columns: [0,128) two 64-column accumulators written by the MMAs, [128,256) read by the timing warps,
// [256,384) the A operand.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(512, 1) tmem_read_under_mma_kernel(int nwarps, int iters, int mma_on, int n_cols, long long* cycles,
                                                                     uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ unsigned long long bar;
  __shared__ int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp == 15) {
    // MMA issuer: M = 128, N = n_cols, K = 16 per instruction, 16 instructions per "tile", alternating accumulators
    if (mma_on && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t lo0 = (((uint32_t)__cvta_generic_to_shared(smem)) >> 4) & 0x3FFFu;
      const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
      uint32_t phase = 0;
      long long n_tiles = 0;
      while (*reinterpret_cast<volatile int*>(&done) < nwarps) {
        for (int t = 0; t < 2; ++t) {
          for (int k = 0; k < 16; ++k) umma_ts(tb + (n_cols == 64 ? t * 64 : 0), tb + 256 + (k & 3) * 8, lo0 + (k & 3) * 2, hi, idesc, k ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(barp), "r"(phase) : "memory");
        }
        phase ^= 1;
        n_tiles += 2;
      }
      cycles[blockIdx.x * 16 + 15] = n_tiles;
    }
  } else if (warp < nwarps) {
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + 128;
    const int share = (nwarps + 3) / 4, part = warp >> 2;
    const int c_lo = 128 / share * part, c_hi = 128 / share * (part + 1);
    __syncwarp();
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t r[32];
        tmem_ld32(base + c, r);
        tmem_ld_wait();
        acc ^= r[0] ^ r[31];
      }
    }
    t1 = clock64();
    if (lane == 0) {
      cycles[blockIdx.x * 16 + warp] = t1 - t0;
      atomicAdd(&done, 1);                        // the MMA loop stops when every reader is through
    }
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}

static void run_under_mma(int nwarps, int mma_on, int n_cols, int grid) {
  const int iters = 2000;
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, grid * 16 * sizeof(long long));
  cudaMalloc(&d_sink, grid * 512 * sizeof(uint32_t));
  cudaMemset(d_cycles, 0, grid * 16 * sizeof(long long));
  cudaFuncSetAttribute(tmem_read_under_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  tmem_read_under_mma_kernel<<<grid, 512, 65536>>>(nwarps, iters, mma_on, n_cols, d_cycles, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("under-mma run: %s\n", cudaGetErrorString(e)); return; }
  long long h[16];
  cudaMemcpy(h, d_cycles, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long worst = 0;
  for (int i = 0; i < nwarps; ++i) worst = h[i] > worst ? h[i] : worst;
  const double bytes = 128.0 * 128 * 4 * iters;      // 128 lanes x 128 columns per iteration
  const double mma_cyc_per_tile = h[15] ? (double)worst / (double)h[15] : 0.0;
  printf("readers=%2d warps  MMA %s (N=%3d)  %9lld cycles -> %6.1f B/clk per SM read", nwarps, mma_on ? "on " : "off", n_cols, worst,
         bytes / worst);
  if (mma_on) printf("   | MMA: %.0f cycles per 128xNx256 tile (ideal %.0f)", mma_cyc_per_tile, 16.0 * n_cols / 2.0);
  printf("\n");
  cudaFree(d_cycles);
  cudaFree(d_sink);
}

// ---- tcgen05.mma issue rate alone: M = 128 (cta_group::1), K = 16 per instruction, 16 instructions per tile, 8 tiles
// between commits.  TS form = A from tensor memory (what the screen kernel uses), SS form = A from shared memory.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int ss_form, int n_cols, int rounds, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  if (__shfl_sync(0xffffffffu, warp, 0) == 1) {
    // the whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform registers: issuing under
    // `if (lane == 0)` costs ~110 cycles per MMA in uniformisation code); one elected lane issues
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t blo = (((uint32_t)__cvta_generic_to_shared(smem)) >> 4) & 0x3FFFu;
    const uint32_t alo = (((uint32_t)__cvta_generic_to_shared(smem + 32768)) >> 4) & 0x3FFFu;
    const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
        for (int t = 0; t < 8; ++t) {
          const uint32_t d = tb + (uint32_t)((t & 1) * n_cols);          // two accumulators (n_cols <= 128)
          if (ss_form) {
#pragma unroll
            for (int k = 0; k < 16; ++k) umma_ss(d, alo + (k & 3) * 2, hi, blo + (k & 3) * 2, hi, idesc, k ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) umma_ts(d, tb + 256 + (k & 3) * 8, blo + (k & 3) * 2, hi, idesc, k ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(barp), "r"(phase) : "memory");
      }
      phase ^= 1;
    }
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_ptr) : "memory");
}
static void run_mma_rate(int ss_form, int n_cols) {
  const int rounds = 200, grid = 148;
  long long* d_cycles;
  cudaMalloc(&d_cycles, grid * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  mma_rate_kernel<<<grid, 128, 65536>>>(ss_form, n_cols, rounds, d_cycles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mma rate run: %s\n", cudaGetErrorString(e)); return; }
  long long h = 0;
  cudaMemcpy(&h, d_cycles, sizeof(long long), cudaMemcpyDeviceToHost);
  const double per_instr = (double)h / (rounds * 8.0 * 16.0);
  printf("%s form  N=%3d : %6.1f cycles per tcgen05.mma (M=128, K=16)   ideal %5.1f   -> %5.1f %% of the tensor peak\n",
         ss_form ? "SS" : "TS", n_cols, per_instr, n_cols / 2.0, 100.0 * (n_cols / 2.0) / per_instr);
  cudaFree(d_cycles);
}

template <int MODE>
static void run(const char* name, int nwarps, int grid) {
  const int iters = 200;
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, grid * 16 * sizeof(long long));
  cudaMalloc(&d_sink, grid * 512 * sizeof(uint32_t));
  cudaMemset(d_cycles, 0, grid * 16 * sizeof(long long));
  tmem_read_kernel<MODE><<<grid, 512>>>(nwarps, iters, d_cycles, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[16 * 148];
  cudaMemcpy(h, d_cycles, grid * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long worst = 0;
  for (int i = 0; i < nwarps; ++i) worst = h[i] > worst ? h[i] : worst;     // CTA 0
  // every quadrant's 512 columns x 32 lanes x 4 B are read once per iteration, by share warps together
  const double bytes = 128.0 * 512 * 4 * iters;
  printf("%-44s warps=%2d  %8lld cycles  -> %6.1f B/clk per SM  (%.0f cycles per 128x64 fp32 tile)\n", name, nwarps, worst,
         bytes / worst, 128.0 * 64 * 4 / (bytes / worst));
  cudaFree(d_cycles);
  cudaFree(d_sink);
}

int main() {
  printf("--- tcgen05.mma issue rate by tile width (one issuing thread per SM, all SMs)\n");
  for (int ss = 0; ss < 2; ++ss)
    for (int n : {32, 64, 96, 128}) run_mma_rate(ss, n);
  printf("--- tensor-memory reads next to a running tcgen05.mma stream (1 CTA per SM, all SMs)\n");
  for (int n_cols : {64, 128}) {
    for (int nw : {4, 8}) {
      run_under_mma(nw, 0, n_cols, 148);
      run_under_mma(nw, 1, n_cols, 148);
    }
  }
  for (int grid : {1, 148}) {
    printf("--- %d CTA(s)\n", grid);
    run<0>("ld.x32, wait after each", 4, grid);
    run<1>("2 x ld.x32 in flight", 4, grid);
    run<2>("4 x ld.x16 in flight", 4, grid);
    run<0>("ld.x32, wait after each, 2 warps/quadrant", 8, grid);
    run<1>("2 x ld.x32 in flight, 2 warps/quadrant", 8, grid);
    run<1>("2 x ld.x32 in flight, 4 warps/quadrant", 16, grid);
  }
  return 0;
}
