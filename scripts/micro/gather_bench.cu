// Micro-benchmark (GPU box only): the cost of the sweep kernels' two L1TEX-side operations --
// random 1-byte gathers from a 10 MB chain array (L2 resident) and random LDS.64 lookups in a
// 1000-entry weight table -- in the production kernel's launch shape (296 CTAs x 256 threads).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

enum { V_PLAIN, V_CG, V_NC, V_TEX, V_WORD, V_NOALLOC, V_CV, V_N };
static const char *names[] = {"ld.global.u8 (default)", "ld.global.cg.u8", "ld.global.nc.u8", "tex1Dfetch u8", "ld.global.nc.u32 word", "ld.global.L1::no_allocate.u8", "ld.volatile.u8"};

template <int V, int B>
__global__ void __launch_bounds__(256, 2) k_gather(const uint8_t *a, cudaTextureObject_t tex, const uint32_t *idx, uint32_t n, uint32_t iters, uint32_t *out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t ix[B], v[B];
#pragma unroll
    for (int b = 0; b < B; ++b) ix[b] = __ldg(idx + (size_t)(it * B + b) * stride + tid);  // coalesced index stream (like the others words)
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const uint32_t i = ix[b];
      if (V == V_PLAIN) v[b] = a[i];
      else if (V == V_CG) v[b] = __ldcg(a + i);
      else if (V == V_NC) v[b] = __ldg(a + i);
      else if (V == V_TEX) v[b] = tex1Dfetch<unsigned char>(tex, (int)i);
      else if (V == V_WORD) v[b] = (__ldg(reinterpret_cast<const uint32_t *>(a) + (i >> 2)) >> (8u * (i & 3u))) & 0xffu;
      else if (V == V_NOALLOC) { uint32_t t; asm volatile("ld.global.L1::no_allocate.u8 %0, [%1];" : "=r"(t) : "l"(a + i)); v[b] = t; }
      else { v[b] = *reinterpret_cast<const volatile uint8_t *>(a + i); }
    }
#pragma unroll
    for (int b = 0; b < B; ++b) acc += v[b];
  }
  if (acc == 0xffffffffu) out[0] = acc;
}

// random LDS.64 lookups in a table of n entries (3 planes like the signed table), ids from a coalesced stream
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_lds(const uint32_t *idx, uint32_t ntab, uint32_t iters, double *out) {
  extern __shared__ double tab[];
  for (uint32_t i = threadIdx.x; i < 3 * ntab; i += blockDim.x) tab[i] = (double)i;
  __syncthreads();
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  double acc = 0;
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t ix[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) ix[b] = __ldg(idx + (size_t)(it * 4 + b) * stride + tid);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      uint32_t i = ix[b] % ntab;
      if (MODE == 1) i = (i & ~15u) | (threadIdx.x & 15u);  // conflict-free: lane-aligned bank pair
      if (MODE == 2) {  // two 32-bit planes
        const uint32_t *t32 = reinterpret_cast<const uint32_t *>(tab);
        acc += __hiloint2double(t32[i], t32[ntab + i]);
      } else acc += tab[i];
    }
  }
  if (acc == -1.0) out[0] = acc;
}

int main() {
  const uint32_t n = 10u << 20, blocks = 296, threads = 256, iters = 64;
  uint8_t *a; uint32_t *idx, *out; double *dout;
  CK(cudaMalloc(&a, n)); CK(cudaMemset(a, 1, n));
  const size_t nidx = (size_t)blocks * threads * iters * 4;
  std::vector<uint32_t> h(nidx);
  uint32_t s = 12345;
  for (size_t i = 0; i < nidx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (uint32_t)(((uint64_t)(s >> 4) * n) >> 28); }
  CK(cudaMalloc(&idx, nidx * 4)); CK(cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&dout, 8));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = a;
  rd.res.linear.desc = cudaCreateChannelDesc<unsigned char>(); rd.res.linear.sizeInBytes = n;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char *name, int B, float ms, double per_sm_ops) {
    printf("%-34s B=%d  %.3f ms  %.1f cycles per warp-op per SM (at %d MHz nominal)\n", name, B, ms, ms * 1e-3 * clk * 1e3 / per_sm_ops, clk / 1000);
  };
#define RUN(V, B) do { \
    const uint32_t it = iters * 4 / B; \
    for (int w = 0; w < 2; ++w) k_gather<V, B><<<blocks, threads>>>(a, tex, idx, n, it, out); \
    cudaEventRecord(e0); k_gather<V, B><<<blocks, threads>>>(a, tex, idx, n, it, out); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); \
    float ms; cudaEventElapsedTime(&ms, e0, e1); report(names[V], B, ms, (double)blocks * 8 * it * B / 148.0); } while (0)
  RUN(V_PLAIN, 2); RUN(V_PLAIN, 4); RUN(V_PLAIN, 8);
  RUN(V_CG, 2); RUN(V_CG, 4); RUN(V_CG, 8);
  RUN(V_NC, 4); RUN(V_TEX, 2); RUN(V_TEX, 4); RUN(V_TEX, 8); RUN(V_WORD, 4); RUN(V_NOALLOC, 4); RUN(V_CV, 4);
#define RUNL(M, name, ntab) do { \
    CK(cudaFuncSetAttribute(k_lds<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * ntab * 8)); \
    for (int w = 0; w < 2; ++w) k_lds<M><<<blocks, threads, 3 * ntab * 8>>>(idx, ntab, iters, dout); \
    cudaEventRecord(e0); k_lds<M><<<blocks, threads, 3 * ntab * 8>>>(idx, ntab, iters, dout); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); \
    float ms; cudaEventElapsedTime(&ms, e0, e1); report(name, 4, ms, (double)blocks * 8 * iters * 4 / 148.0); } while (0)
  RUNL(0, "LDS.64 random of 1000", 1000); RUNL(1, "LDS.64 lane-aligned banks", 1000); RUNL(2, "2 x LDS.32 random of 1000", 1000);
  return 0;
}
