// scan_bw.cu -- development microbenchmark (not product code): read-bandwidth ceilings of the
// access patterns considered for the K1 scan on B200.  nvcc -arch=sm_100a -O3 -o scan_bw scan_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float4 ldnc(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldnc128(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// (1) pure coalesced stream: each warp reads 512 contiguous bytes per instruction, UN in flight
template <int UN>
__global__ void k_coalesced(const float4* __restrict__ x, size_t n4, float* out) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = tid; i + (UN - 1) * stride < n4; i += UN * stride) {
    float4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) v[u] = ldnc(x + i + u * stride);
#pragma unroll
    for (int u = 0; u < UN; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

// (2) 4 threads per 512-byte row, 8 loads of 16 B at stride 64 B per thread (the lane-exact pattern)
template <int U, int MINB, bool L2HINT>
__global__ void __launch_bounds__(256, MINB) k_rows4(const float* __restrict__ X, int64_t n, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int t = lane & 3, g = lane >> 2;
  const int64_t rpb = 8 * U;
  const int64_t nb = n / rpb;
  float acc = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nb; b += (int64_t)gridDim.x * nw) {
    float4 v[U][8];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const float4* p = reinterpret_cast<const float4*>(X + (size_t)(b * rpb + u * 8 + g) * 128) + t;
#pragma unroll
      for (int j = 0; j < 8; j++) v[u][j] = L2HINT ? ldnc128(p + j * 4) : ldnc(p + j * 4);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(v[u][j].x, v[u][j].x, fmaf(v[u][j].y, v[u][j].y, fmaf(v[u][j].z, v[u][j].z, fmaf(v[u][j].w, v[u][j].w, acc))));
  }
  if (acc == 123.456f) out[0] = acc;
}

// (3) per-warp private ring of bulk async copies (TMA 1D), consumer reads smem with the 4-threads-per-row pattern
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TR, int NS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_bulk(const float* __restrict__ X, int64_t n, float* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int STAGE = TR * 512;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane & 3, g = lane >> 2;
  unsigned char* ring = smem_raw + (size_t)warp * NS * STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * NS * STAGE) + warp * NS;
  if (lane == 0) {
    for (int s = 0; s < NS; s++) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t ntiles = n / TR;
  const int64_t first = (int64_t)blockIdx.x * NW + warp, step = (int64_t)gridDim.x * NW;
  // prologue
  if (lane == 0) {
    for (int s = 0; s < NS; s++) {
      int64_t tile = first + s * step;
      if (tile < ntiles) {
        mbar_expect_tx(bars + s, STAGE);
        bulk_g2s(ring + s * STAGE, X + (size_t)tile * TR * 128, STAGE, bars + s);
      }
    }
  }
  float acc = 0.f;
  int s = 0;
  uint32_t parity = 0;
  for (int64_t tile = first; tile < ntiles; tile += step) {
    mbar_wait(bars + s, parity);
    const unsigned char* st = ring + s * STAGE;
#pragma unroll
    for (int r0 = 0; r0 < TR; r0 += 8) {
      const float4* p = reinterpret_cast<const float4*>(st + (size_t)(r0 + g) * 512) + t;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = p[j * 4];
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, acc))));
    }
    __syncwarp();
    if (lane == 0) {
      int64_t nt = tile + (int64_t)NS * step;
      if (nt < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bars + s, STAGE);
        bulk_g2s(ring + s * STAGE, X + (size_t)nt * TR * 128, STAGE, bars + s);
      }
    }
    if (++s == NS) { s = 0; parity ^= 1; }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float timeit(F f, int iters = 20) {
  for (int i = 0; i < 3; i++) f();
  CK(cudaDeviceSynchronize());
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < iters; i++) f();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 1000000;
  const size_t bytes = (size_t)n * 512;
  float *X, *out;
  CK(cudaMalloc(&X, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(X, 0, bytes));
  auto rep = [&](const char* name, float ms) { printf("%-34s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  const size_t n4 = bytes / 16;
  rep("coalesced UN=4 g=148*8", timeit([&] { k_coalesced<4><<<148 * 8, 256>>>((const float4*)X, n4, out); }));
  rep("coalesced UN=8 g=148*8", timeit([&] { k_coalesced<8><<<148 * 8, 256>>>((const float4*)X, n4, out); }));
  rep("coalesced UN=8 g=148*4", timeit([&] { k_coalesced<8><<<148 * 4, 256>>>((const float4*)X, n4, out); }));
  rep("coalesced UN=16 g=148*4", timeit([&] { k_coalesced<16><<<148 * 4, 256>>>((const float4*)X, n4, out); }));
  rep("rows4 U=2 occ2 L2hint", timeit([&] { k_rows4<2, 2, true><<<148 * 2, 256>>>(X, n, out); }));
  rep("rows4 U=2 occ2 nohint", timeit([&] { k_rows4<2, 2, false><<<148 * 2, 256>>>(X, n, out); }));
  rep("rows4 U=1 occ4 L2hint", timeit([&] { k_rows4<1, 4, true><<<148 * 4, 256>>>(X, n, out); }));
  rep("rows4 U=1 occ4 nohint", timeit([&] { k_rows4<1, 4, false><<<148 * 4, 256>>>(X, n, out); }));
  rep("rows4 U=1 occ6 nohint", timeit([&] { k_rows4<1, 6, false><<<148 * 6, 256>>>(X, n, out); }));
  rep("rows4 U=1 occ8 nohint", timeit([&] { k_rows4<1, 8, false><<<148 * 8, 256>>>(X, n, out); }));
  rep("rows4 U=2 occ3 nohint", timeit([&] { k_rows4<2, 3, false><<<148 * 3, 256>>>(X, n, out); }));
  rep("rows4 U=4 occ1 nohint", timeit([&] { k_rows4<4, 1, false><<<148 * 1, 256>>>(X, n, out); }));
  {
    auto run = [&](auto kern, int nw, int tr, int ns, const char* name) {
      size_t smem = (size_t)nw * ns * tr * 512 + nw * ns * 8 + 128;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      rep(name, timeit([&] { kern<<<148, nw * 32, smem>>>(X, n, out); }));
    };
    run(k_bulk<16, 3, 8>, 8, 16, 3, "bulk TR=16 NS=3 NW=8 (192K)");
    run(k_bulk<16, 2, 8>, 8, 16, 2, "bulk TR=16 NS=2 NW=8 (128K)");
    run(k_bulk<8, 4, 8>, 8, 8, 4, "bulk TR=8 NS=4 NW=8 (128K)");
    run(k_bulk<8, 6, 8>, 8, 8, 6, "bulk TR=8 NS=6 NW=8 (192K)");
    run(k_bulk<32, 3, 4>, 4, 32, 3, "bulk TR=32 NS=3 NW=4 (192K)");
    run(k_bulk<16, 3, 4>, 4, 16, 3, "bulk TR=16 NS=3 NW=4 (96K)");
    run(k_bulk<8, 3, 16>, 16, 8, 3, "bulk TR=8 NS=3 NW=16 (192K)");
  }
  return 0;
}
