// Micro-benchmark (GPU box only): how fast can the staged sweep's data path stream a blob?
// Persistent grid (2 CTAs per SM), every warp owns a ring of `stages` buffers and fetches its
// slices (one contiguous run of `bytes`) with one cp.async.bulk per slice, waits on the mbarrier,
// reads the slice from shared memory (one LDS.128 per 16 B) and moves on.  Variants:
//   lookahead 0: wait for slice k, consume it, refill its stage with slice k + stages
//   lookahead 1: the production kernel's order -- wait for slice k + 1, consume slice k, refill
// against a plain coalesced LDG.128 stream of the same bytes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol, bool hint) {
  if (hint) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
  else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int LOOKAHEAD>
__global__ void __launch_bounds__(512, 1) k_ring(const uint4 *blob, uint32_t nslice, uint32_t bytes, uint32_t stages, uint32_t work, int hint, uint32_t *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarp = blockDim.x >> 5;
  const uint32_t bar0 = smem_addr(smem) + warp * stages * 8u;
  const uint32_t ring_off = ((nwarp * stages * 8u + 127u) & ~127u) + warp * stages * bytes;
  const uint32_t ring0 = smem_addr(smem) + ring_off;
  if (lane == 0) for (uint32_t i = 0; i < stages; ++i) mbar_init(bar0 + 8u * i, 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const uint32_t step = (gridDim.x * blockDim.x) >> 5, s0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s0 >= nslice) return;
  const uint32_t n_my = (nslice - 1 - s0) / step + 1u;
  auto issue = [&](uint32_t k, uint32_t stage) {
    if (lane == 0) {
      mbar_expect(bar0 + 8u * stage, bytes);
      bulk_g2s(ring0 + stage * bytes, reinterpret_cast<const unsigned char *>(blob) + (size_t)(s0 + (size_t)k * step) * bytes, bytes, bar0 + 8u * stage, pol, hint);
    }
  };
  const uint32_t pre = n_my < stages ? n_my : stages;
  for (uint32_t k = 0; k < pre; ++k) issue(k, k);
  uint32_t stage = 0, parity = 0, acc = 0;
  if (LOOKAHEAD) mbar_wait(bar0, 0);
  for (uint32_t k = 0; k < n_my; ++k) {
    uint32_t nstage = stage + 1, nparity = parity;
    if (nstage == stages) { nstage = 0; nparity ^= 1u; }
    if (LOOKAHEAD) { if (k + 1 < n_my) mbar_wait(bar0 + 8u * nstage, nparity); }
    else mbar_wait(bar0 + 8u * stage, parity);
    const uint4 *b = reinterpret_cast<const uint4 *>(smem + ring_off + stage * bytes);
    for (uint32_t i = lane; i < bytes / 16; i += 32) { const uint4 v = b[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    for (uint32_t w = 0; w < work; ++w) acc = acc * 1664525u + 1013904223u;  // dependent ALU work per slice
    __syncwarp();
    if (k + stages < n_my) issue(k + stages, stage);
    stage = nstage; parity = nparity;
  }
  if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(512, 2) k_ldg(const uint4 *blob, size_t n16, uint32_t *out) {
  uint32_t acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(blob + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n16; i += stride) { const uint4 v = __ldcs(blob + i); acc += v.x ^ v.y; }
  if (acc == 0x12345u) out[0] = acc;
}

int main() {
  const size_t total = 600ull << 20;
  uint4 *blob; uint32_t *out;
  CK(cudaMalloc(&blob, total)); CK(cudaMemset(blob, 1, total)); CK(cudaMalloc(&out, 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  {
    for (int w = 0; w < 2; ++w) k_ldg<<<sms * 2, 512>>>(blob, total / 16, out);
    cudaEventRecord(e0); k_ldg<<<sms * 2, 512>>>(blob, total / 16, out); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("LDG.128 stream                                   %.3f ms  %.0f GB/s\n", ms, total / ms * 1e-6);
  }
  const uint32_t sizes[] = {2304, 4608};
  for (uint32_t bytes : sizes)
    for (int la = 0; la < 2; ++la)
      for (uint32_t nwarp : {8u, 16u})
        for (uint32_t stages : {2u, 3u, 4u})
          for (uint32_t work : {0u, 400u})
            for (int hint : {1, 0}) {
              if (hint == 0 && !(stages == 3 && work == 0)) continue;
              const uint32_t nslice = (uint32_t)(total / bytes);
              const size_t smem = ((nwarp * stages * 8u + 127u) & ~127u) + (size_t)nwarp * stages * bytes;
              if (smem > 110 * 1024) continue;
              auto kern = la ? k_ring<1> : k_ring<0>;
              CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
              for (int w = 0; w < 2; ++w) kern<<<sms * 2, nwarp * 32, smem>>>(blob, nslice, bytes, stages, work, hint, out);
              cudaEventRecord(e0); kern<<<sms * 2, nwarp * 32, smem>>>(blob, nslice, bytes, stages, work, hint, out); cudaEventRecord(e1);
              CK(cudaDeviceSynchronize());
              float ms; cudaEventElapsedTime(&ms, e0, e1);
              printf("bulk ring bytes=%u lookahead=%d warps/CTA=%2u stages=%u work=%3u hint=%d  %.3f ms  %.0f GB/s\n", bytes, la, nwarp, stages, work, hint, ms, (double)nslice * bytes / ms * 1e-6);
            }
  return 0;
}
