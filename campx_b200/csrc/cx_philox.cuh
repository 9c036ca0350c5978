// Counter-based synthetic actions shared by cx_fill_actions and the in-kernel generators of cx_rollout_synth.
// Philox4x32-10 (Salmon et al., SC'11): key = seed, counter = (global_env >> 2, step); the four output
// words serve the four envs of a quad: action = mulhi(word[global_env & 3], n_actions).  Any (env, step)
// can therefore be regenerated independently, on the device or on the host (tests do both).
#pragma once
#include <stdint.h>

struct CxPhilox4 {
  uint32_t w[4];
};

__device__ __forceinline__ CxPhilox4 cx_philox4(uint64_t seed, uint64_t quad, uint64_t step) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)quad, c1 = (uint32_t)(quad >> 32), c2 = (uint32_t)step, c3 = (uint32_t)(step >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  CxPhilox4 out;
  out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

// packed actions (one byte each) of the quad of global envs [4q, 4q+3] at `step`
__device__ __forceinline__ uint32_t cx_synth_actions_quad(uint64_t seed, uint64_t quad, uint64_t step, uint32_t A) {
  const CxPhilox4 p = cx_philox4(seed, quad, step);
  return __umulhi(p.w[0], A) | (__umulhi(p.w[1], A) << 8) | (__umulhi(p.w[2], A) << 16) | (__umulhi(p.w[3], A) << 24);
}

__device__ __forceinline__ uint32_t cx_synth_action(uint64_t seed, uint64_t env, uint64_t step, uint32_t A) {
  const CxPhilox4 p = cx_philox4(seed, env >> 2, step);
  return __umulhi(p.w[env & 3], A);
}
