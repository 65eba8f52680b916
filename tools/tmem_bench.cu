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
