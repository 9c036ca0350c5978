// Development aid: ceiling of a pure HBM write stream on B200 (16 B per lane, fully coalesced),
// for different store cache operators and grid shapes.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k_write(uint4* __restrict__ dst, size_t n16, uint32_t v) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 val = make_uint4(v, v + 1, v + 2, v + 3);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    if (MODE == 0) dst[i] = val;
    if (MODE == 1) __stcs(dst + i, val);
    if (MODE == 2) __stwt(dst + i, val);
    if (MODE == 3) __stcg(dst + i, val);
  }
}
// chunked like the rollout kernel: each warp owns a contiguous 6400-byte span per step, T steps apart by row
template <int MODE>
__global__ void k_write_tiles(uint4* __restrict__ dst, int T, size_t row16, uint32_t v) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint4 val = make_uint4(v, v + 1, v + 2, v + 3);
  for (int t = 0; t < T; ++t) {
    uint4* d = dst + (size_t)t * row16 + warp * 400;
    for (int k = lane; k < 400; k += 32) {
      if (MODE == 1) __stcs(d + k, val); else d[k] = val;
    }
  }
}
int main() {
  const size_t bytes = (size_t)1 << 30, n16 = bytes / 16;
  uint4* d; cudaMalloc(&d, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms, double b) { printf("%-34s %.3f ms  %.0f GB/s\n", name, ms, b / ms / 1e6); };
  int grids[] = {148 * 2, 148 * 8, 148 * 16, 148 * 64};
  for (int g : grids) {
    for (int mode = 0; mode < 4; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_write<0><<<g, 256>>>(d, n16, rep);
        if (mode == 1) k_write<1><<<g, 256>>>(d, n16, rep);
        if (mode == 2) k_write<2><<<g, 256>>>(d, n16, rep);
        if (mode == 3) k_write<3><<<g, 256>>>(d, n16, rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
      }
      char name[64]; snprintf(name, 64, "grid-stride grid=%d mode=%d", g, mode);
      report(name, best, (double)bytes);
    }
  }
  {  // rollout-shaped: 4096 warps x 32 steps x 6400 B = 838 MB
    const int T = 32; const size_t row16 = (size_t)4096 * 400;
    for (int mode = 0; mode < 2; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_write_tiles<0><<<1024, 128>>>(d, T, row16, rep); else k_write_tiles<1><<<1024, 128>>>(d, T, row16, rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
      }
      report(mode ? "rollout-shaped tiles st.cs" : "rollout-shaped tiles st", best, (double)T * row16 * 16);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
