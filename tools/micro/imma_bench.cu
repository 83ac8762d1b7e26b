// Micro-benchmark: issue rate of legacy warp-level integer MMA (mma.sync.m16n8k32 s8 x s8 -> s32) on B200,
// next to the scalar IMAD rate - the two candidates for the exact long-term autocorrelation (E6a).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_bench imma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_imma(int iters, int* out)
{
  int c[8][4];
  for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) c[t][i] = 0;
  unsigned a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x5a5a5a5au, b1 = b0 + 7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 8; t++)
      asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+r"(c[t][0]), "+r"(c[t][1]), "+r"(c[t][2]), "+r"(c[t][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + t), "r"(b1 + t));
  }
  int s = 0;
  for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_imad(int iters, int* out)
{
  int acc[16];
  for (int i = 0; i < 16; i++) acc[i] = i;
  int x = threadIdx.x | 1, y = blockIdx.x | 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = acc[i] + x * (y + i);
    x += 2;
  }
  int s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  int* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(int));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 20000, grid = 148 * 2, threads = warps * 32 / 2;
    k_imma<<<grid, threads>>>(100, out);
    cudaEventRecord(e0); k_imma<<<grid, threads>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)grid * (threads / 32) * iters * 8.0;
    printf("IMMA m16n8k32 s8: %d warps/SM: %.1f G mma/s = %.1f T MAC/s (%.3f ms)\n", warps, mmas / ms / 1e6, mmas * 4096 / ms / 1e9, ms);
  }
  {
    const int iters = 20000, grid = 148 * 4, threads = 256;
    k_imad<<<grid, threads>>>(100, out);
    cudaEventRecord(e0); k_imad<<<grid, threads>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double macs = (double)grid * threads * iters * 16.0;
    printf("scalar IMAD: %.2f T MAC/s (%.3f ms)\n", macs / ms / 1e9, ms);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
