// Development microbenchmark: conflict-free shared-memory lookups per clock per SM, by access width.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bench lds_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int WIDTH>
__device__ __forceinline__ unsigned int lds(unsigned int a) {
  unsigned int v;
  if (WIDTH == 8) asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  else if (WIDTH == 16) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
template <int WIDTH, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(unsigned int* out, int iters, long long* cyc) {
  extern __shared__ unsigned char sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < 32768; i += THREADS) reinterpret_cast<unsigned int*>(sm)[i] = i * 2654435761u;
  __syncthreads();
  unsigned int base = (unsigned int)__cvta_generic_to_shared(sm);
  unsigned int a[16];
  unsigned int x = tid * 747796405u + 1u;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    x = x * 1664525u + 1013904223u;
    a[j] = base + j * 8192 + ((x >> 20) & 63u) * 128 + lane * 4 + (WIDTH == 8 ? ((x >> 10) & 3u) : WIDTH == 16 ? ((x >> 10) & 2u) : 0u);
  }
  unsigned int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      a[j] ^= 128u;  // another line of the same lane-private bank: the load cannot be merged with the previous one
      acc += lds<WIDTH>(a[j]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int W, int T>
void run(const char* name) {
  unsigned int* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k<W, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  const int iters = 4000;
  k<W, T><<<148, T, 131072>>>(out, iters, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<W, T><<<148, T, 131072>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("[%.1f us] ", ms * 1e3);
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  printf("%s threads %d: %.0f cycles, %.3f LDS warp-instr per clk per SM (err %s)\n", name, T, c, (T / 32.0) * iters * 16 / c, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<8, 1024>("LDS.U8 "); run<16, 1024>("LDS.U16"); run<32, 1024>("LDS.32 ");
  run<8, 512>("LDS.U8 "); run<32, 512>("LDS.32 "); run<8, 256>("LDS.U8 "); run<32, 256>("LDS.32 ");
  return 0;
}
