// Development aid: which of the rollout kernel's HBM streams costs the gap to a pure board writer?
// Same geometry as k_agent_rollout: 1024 CTAs x 4 warps, each warp owns 256 envs for T steps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <bool BOARD, bool REWARD, bool FLAGS, bool ACTIONS, int PF = 0>
__global__ void __launch_bounds__(128, 7) k(uint4* board, float4* reward, uint32_t* flags, const uint32_t* actions, int T, size_t n, uint32_t v) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t env0 = warp * 256;
  uint32_t acc = v;
  if (PF == 1) {  // pull this warp's whole action column into L2 up front (T rows x 256 B)
    for (int t = lane; t < 2 * T; t += 32) {
      const char* a = (const char*)actions + ((size_t)(t >> 1) * n + env0) + (t & 1) * 128;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    }
  }
  if (PF == 2) {
    for (int t = lane; t < 2 * T; t += 32) {
      const char* a = (const char*)actions + ((size_t)(t >> 1) * n + env0) + (t & 1) * 128;
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
    }
  }
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)t * n + env0;
    if (ACTIONS) { acc += __ldcs(actions + (row + lane * 4) / 4); acc += __ldcs(actions + (row + 128 + lane * 4) / 4); }
    if (REWARD) { float f = __uint_as_float(acc & 0x3fffffff); __stcs(reward + (row + lane * 4) / 4, make_float4(f, f, f, f)); __stcs(reward + (row + 128 + lane * 4) / 4, make_float4(f, f, f, f)); }
    if (FLAGS) { __stcs(flags + (row + lane * 4) / 4, acc); __stcs(flags + (row + 128 + lane * 4) / 4, acc); }
    if (BOARD) { uint4* d = board + row * 25 / 16; const uint4 val = make_uint4(acc, acc + 1, acc + 2, acc + 3);
      for (int k2 = lane; k2 < 400; k2 += 32) __stcs(d + k2, val); }
  }
}
// reads batched: every B steps the warp loads its next B action rows into registers
template <int B, int PF>
__global__ void __launch_bounds__(128, 7) kb(uint4* board, float4* reward, uint32_t* flags, const uint32_t* actions, int T, size_t n, uint32_t v) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t env0 = warp * 256;
  if (PF) for (int t = lane; t < 2 * T; t += 32) {
      const char* a = (const char*)actions + ((size_t)(t >> 1) * n + env0) + (t & 1) * 128;
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
  }
  uint32_t acc = v, a0[B], a1[B];
  for (int t0 = 0; t0 < T; t0 += B) {
#pragma unroll
    for (int b = 0; b < B; ++b) { const size_t row = (size_t)(t0 + b) * n + env0; a0[b] = __ldcs(actions + (row + lane * 4) / 4); a1[b] = __ldcs(actions + (row + 128 + lane * 4) / 4); }
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const size_t row = (size_t)(t0 + b) * n + env0;
      acc += a0[b] + a1[b];
      float f = __uint_as_float(acc & 0x3fffffff); __stcs(reward + (row + lane * 4) / 4, make_float4(f, f, f, f)); __stcs(reward + (row + 128 + lane * 4) / 4, make_float4(f, f, f, f));
      __stcs(flags + (row + lane * 4) / 4, acc); __stcs(flags + (row + 128 + lane * 4) / 4, acc);
      uint4* d = board + row * 25 / 16; const uint4 val = make_uint4(acc, acc + 1, acc + 2, acc + 3);
      for (int k2 = lane; k2 < 400; k2 += 32) __stcs(d + k2, val);
    }
  }
}
int main() {
  const int T = 32; const size_t n = 1 << 20;
  uint4* board; float4* reward; uint32_t* flags; uint32_t* actions;
  cudaMalloc(&board, T * n * 25); cudaMalloc(&reward, T * n * 4); cudaMalloc(&flags, T * n); cudaMalloc(&actions, T * n);
  cudaMemset(actions, 1, T * n);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto kern, double bytes) {
    float best = 1e9;
    for (int rep = 0; rep < 8; ++rep) {
      cudaEventRecord(e0); kern<<<1024, 128>>>(board, reward, flags, actions, T, n, rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 1 && ms < best) best = ms;
    }
    printf("%-40s %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
  };
  const double N = (double)T * n;
  run("board only (25 B)", k<true, false, false, false>, N * 25);
  run("board + reward (29 B)", k<true, true, false, false>, N * 29);
  run("board + reward + flags (30 B)", k<true, true, true, false>, N * 30);
  run("board + reward + flags + actions (31 B)", k<true, true, true, true>, N * 31);
  run("board + actions (26 B)", k<true, false, false, true>, N * 26);
  run("all 4 streams + L2 prefetch of actions", k<true, true, true, true, 1>, N * 31);
  run("all 4 streams + L2 evict_last prefetch", k<true, true, true, true, 2>, N * 31);
  run("batched reads B=4", kb<4, 0>, N * 31);
  run("batched reads B=8", kb<8, 0>, N * 31);
  run("batched reads B=16", kb<16, 0>, N * 31);
  run("batched reads B=32 (all up front)", kb<32, 0>, N * 31);
  run("batched reads B=8 + evict_last prefetch", kb<8, 1>, N * 31);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
