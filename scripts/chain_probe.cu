// Development probe (round 2): what sits on the per-step dependent chain of the tile kernels at small batches?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/chain_probe scripts/chain_probe.cu && scripts/chain_probe
// Measures, per iteration and in SM clocks, for one warp per CTA (1 CTA per SM) and for 8 warps per CTA:
//   0  fence.proxy.async alone                       (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC.S)
//   1  STS.U8 + fence
//   2  STG.32 (st.global.cs) + fence                 does the fence wait for the global store's acknowledgement?
//   3  LDG.U8 issued, consumed 2 iterations later + fence   ... or for loads in flight?
//   4  UBLKCP of B bytes + commit + wait_group.read 0        TMA issue -> shared memory read
//   5  STS + fence + UBLKCP + commit + wait_group.read 1  (two tiles alternating)
//   6  like 5 plus STG.32 + STG.U8 before the fence (the step of the lane-per-env kernel without its table look-ups)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk(void* g, const void* s, uint32_t b) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(b) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int MODE>
__global__ void k(uint8_t* out, float* rw, const uint8_t* act, long long* cyc, int iters, int bytes, size_t stride) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  uint8_t* tile = sm + (size_t)warp * 2 * bytes;
  for (int i = lane; i < 2 * bytes; i += 32) tile[i] = (uint8_t)i;
  __syncthreads();
  const size_t wid = (size_t)blockIdx.x * nw + warp;
  uint8_t* dst = out + wid * bytes;
  uint32_t a0 = 0, a1 = 0, acc = 0;
  long long t0 = clock64();
  for (int t = 0; t < iters; ++t) {
    if (MODE == 0) fence_async();
    if (MODE == 1) { tile[lane * 25 + (t % 25)] = (uint8_t)t; fence_async(); }
    if (MODE == 2) { __stcs(rw + (size_t)t * stride + wid * 32 + lane, (float)t); fence_async(); }
    if (MODE == 3) { acc += a0; a0 = a1; a1 = act[(size_t)t * stride + wid * 32 + lane]; fence_async(); }
    if (MODE == 4) {
      if (lane == 0) { bulk(dst + (size_t)t * stride * 25, tile, bytes); asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
      __syncwarp();
    }
    if (MODE == 5 || MODE == 6) {
      uint8_t* tl = tile + (t & 1) * bytes;
      if (MODE == 6) {
        __stcs(rw + (size_t)t * stride + wid * 32 + lane, (float)t);
        reinterpret_cast<uint8_t*>(rw)[(size_t)(iters + 1) * stride * 4 + (size_t)t * stride + wid * 32 + lane] = (uint8_t)t;
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      tl[lane * 25 + (t % 25)] = (uint8_t)t;
      fence_async();
      __syncwarp();
      if (lane == 0) bulk(dst + (size_t)t * stride * 25, tl, bytes);
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) + (acc == 12345678u);
}
template <int MODE>
void run(const char* name, int warps, int bytes, uint8_t* out, float* rw, uint8_t* act, long long* cyc, size_t stride) {
  const int iters = 64;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<MODE><<<148, warps * 32, (size_t)warps * 2 * bytes>>>(out, rw, act, cyc, iters, bytes, stride);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-52s warps/CTA %d bytes %5d: %7.1f clk/iter  (%s)\n", name, warps, bytes, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const size_t stride = 148 * 8 * 32;          // "n": envs of the whole grid at 8 warps per CTA
  uint8_t *out, *act; float* rw; long long* cyc;
  cudaMalloc(&out, stride * 25 * 66 * 8 + (1 << 20)); cudaMalloc(&rw, stride * 66 * 5 + (1 << 20)); cudaMalloc(&act, stride * 66);
  cudaMalloc(&cyc, 8);
  cudaMemset(act, 1, stride * 66);
  for (int w : {1, 4, 8}) {
    run<0>("fence.proxy.async alone", w, 800, out, rw, act, cyc, stride);
    run<1>("STS.U8 + fence", w, 800, out, rw, act, cyc, stride);
    run<2>("STG.32 + fence", w, 800, out, rw, act, cyc, stride);
    run<3>("LDG.U8 (used 2 iterations later) + fence", w, 800, out, rw, act, cyc, stride);
    for (int b : {800, 1600, 6400}) run<4>("UBLKCP + wait_group.read 0", w, b, out, rw, act, cyc, stride);
    for (int b : {800, 1600, 6400}) run<5>("STS + fence + UBLKCP, 2 tiles (wait_group.read 1)", w, b, out, rw, act, cyc, stride);
    for (int b : {800, 1600}) run<6>("the same + STG.32 + STG.U8 before the fence", w, b, out, rw, act, cyc, stride);
  }
  return 0;
}
