// Write-bandwidth probes for the k_decode output pattern (scratch; not product code).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void k_fill4(float4* p, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += st) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
// 8 SoA columns, one warp writes `cnt` consecutive elements per "block" at a running offset.
// mode 0: cnt = 32 (aligned); mode 1: cnt = popc(random mask ~95%) (unaligned compaction)
template <int MODE>
__global__ void k_cols(float* x, float* y, float* z, uint32_t* t, uint16_t* a, uint16_t* d, uint8_t* in,
                       uint8_t* l, const unsigned* masks, const unsigned long long* offs, int nblk_per_warp,
                       int total_warps) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= total_warps) return;
  const unsigned lt = (1u << lane) - 1u;
  unsigned long long o = offs[gw];
  for (int b = 0; b < nblk_per_warp; ++b) {
    const unsigned m = MODE ? masks[(size_t)gw * nblk_per_warp + b] : 0xffffffffu;
    if ((m >> lane) & 1u) {
      const unsigned long long q = o + __popc(m & lt);
      x[q] = 1.f; y[q] = 2.f; z[q] = 3.f; t[q] = 4u; a[q] = 5; d[q] = 6; in[q] = 7; l[q] = 8;
    }
    o += __popc(m);
  }
}
int main() {
  const size_t npts = 384ull << 20;  // 402M points
  float *x, *y, *z; uint32_t* t; uint16_t *a, *d; uint8_t *in, *l;
  CK(cudaMalloc(&x, npts * 4)); CK(cudaMalloc(&y, npts * 4)); CK(cudaMalloc(&z, npts * 4)); CK(cudaMalloc(&t, npts * 4));
  CK(cudaMalloc(&a, npts * 2)); CK(cudaMalloc(&d, npts * 2)); CK(cudaMalloc(&in, npts)); CK(cudaMalloc(&l, npts));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // 1. pure fill
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_fill4<<<148 * 8, 256>>>((float4*)x, npts / 4);
    k_fill4<<<148 * 8, 256>>>((float4*)y, npts / 4);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("fill4: %.3f ms  %.1f GB/s\n", ms, 2.0 * npts * 4 / ms / 1e6);
  }
  // column writers: each warp handles nblk consecutive "blocks"
  for (int nblk : {48, 12 * 32}) {
    const int total_warps = (int)(npts / 32 / nblk);
    std::vector<unsigned> hm((size_t)total_warps * nblk);
    std::vector<unsigned long long> ho(total_warps), ho0(total_warps);
    uint64_t s = 88172645463325252ull; unsigned long long run = 0;
    for (int w = 0; w < total_warps; ++w) {
      ho[w] = run; ho0[w] = (unsigned long long)w * nblk * 32;
      for (int b = 0; b < nblk; ++b) {
        unsigned m = 0;
        for (int k = 0; k < 32; ++k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; if ((s % 100) >= 5) m |= 1u << k; }
        hm[(size_t)w * nblk + b] = m; run += __builtin_popcount(m);
      }
    }
    unsigned* dm; unsigned long long *dof, *dof0;
    CK(cudaMalloc(&dm, hm.size() * 4)); CK(cudaMalloc(&dof, ho.size() * 8)); CK(cudaMalloc(&dof0, ho.size() * 8));
    CK(cudaMemcpy(dm, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dof, ho.data(), ho.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dof0, ho0.data(), ho0.size() * 8, cudaMemcpyHostToDevice));
    const int blocks = (total_warps * 32 + 255) / 256;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_cols<0><<<blocks, 256>>>(x, y, z, t, a, d, in, l, dm, dof0, nblk, total_warps);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("cols aligned   nblk=%d: %.3f ms  %.1f GB/s\n", nblk, ms, 22.0 * npts / ms / 1e6);
      cudaEventRecord(e0);
      k_cols<1><<<blocks, 256>>>(x, y, z, t, a, d, in, l, dm, dof, nblk, total_warps);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      printf("cols compacted nblk=%d: %.3f ms  %.1f GB/s\n", nblk, ms, 22.0 * run / ms / 1e6);
    }
    cudaFree(dm); cudaFree(dof); cudaFree(dof0);
  }
  return 0;
}
