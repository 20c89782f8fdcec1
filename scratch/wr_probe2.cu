// Write-bandwidth probes, part 2 (scratch): unaligned / compacted column stores without loads.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)
__device__ __forceinline__ unsigned hash32(unsigned v) { v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16; return v; }
// each warp owns a contiguous region of nblk*32 slots; MODE 0 aligned full; 1 shifted by `shift` full;
// 2 compacted with ~95% masks (register hash); NCOL columns
template <int MODE, int NCOL>
__global__ void k_cols(float* x, float* y, float* z, uint32_t* t, uint16_t* a, uint16_t* d, uint8_t* in,
                       uint8_t* l, int nblk, int total_warps, int shift) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < total_warps; gw += (gridDim.x * blockDim.x) >> 5) {
    unsigned long long o = (unsigned long long)gw * nblk * 32 + (MODE == 1 ? shift : 0);
    for (int b = 0; b < nblk; ++b) {
      unsigned m = 0xffffffffu;
      if (MODE == 2) { const unsigned h = hash32(gw * 977 + b); m = ~(1u << (h & 31)) & ~(((h >> 8) & 1) << ((h >> 9) & 31)); }
      if ((m >> lane) & 1u) {
        const unsigned long long q = o + __popc(m & lt);
        x[q] = 1.f; if (NCOL > 1) { y[q] = 2.f; z[q] = 3.f; t[q] = 4u; } if (NCOL > 4) { a[q] = 5; d[q] = 6; in[q] = 7; l[q] = 8; }
      }
      o += __popc(m);
    }
  }
}
int main() {
  const size_t npts = 384ull << 20;
  float *x, *y, *z; uint32_t* t; uint16_t *a, *d; uint8_t *in, *l;
  CK(cudaMalloc(&x, npts * 4 + 256)); CK(cudaMalloc(&y, npts * 4 + 256)); CK(cudaMalloc(&z, npts * 4 + 256)); CK(cudaMalloc(&t, npts * 4 + 256));
  CK(cudaMalloc(&a, npts * 2 + 256)); CK(cudaMalloc(&d, npts * 2 + 256)); CK(cudaMalloc(&in, npts + 256)); CK(cudaMalloc(&l, npts + 256));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int nblk = 48;
  const int total_warps = (int)(npts / 32 / nblk);
  for (int grid : {296, 148 * 8, 148 * 32}) {
    for (int rep = 0; rep < 2; ++rep) {
#define RUN(MODE, NCOL, BYTES, SH, NAME) \
      cudaEventRecord(e0); k_cols<MODE, NCOL><<<grid, 256>>>(x, y, z, t, a, d, in, l, nblk, total_warps, SH); \
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); \
      printf("grid %5d %-28s %.3f ms  %.1f GB/s\n", grid, NAME, ms, (double)BYTES * npts / ms / 1e6);
      RUN(0, 8, 22, 0, "8col aligned")
      RUN(1, 8, 22, 5, "8col shifted+5")
      RUN(2, 8, 22 * 0.95, 0, "8col compacted(~95%)")
      RUN(0, 4, 16, 0, "4col(f32) aligned")
      RUN(1, 4, 16, 5, "4col(f32) shifted+5")
      RUN(1, 4, 16, 8, "4col(f32) shifted+8")
      RUN(2, 4, 16 * 0.95, 0, "4col(f32) compacted")
      RUN(1, 1, 4, 5, "1col(f32) shifted+5")
    }
  }
  return 0;
}
