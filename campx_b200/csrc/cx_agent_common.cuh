// Device helpers shared by the single-agent kernels (cx_agent_kernels.cu, cx_agent_obs_kernels.cu).
#pragma once
#include "cx_internal.cuh"

namespace {

template <bool VEC>
__device__ __forceinline__ uint32_t ld_u8x4(const uint8_t* p, int64_t i, int64_t end, uint32_t fill) {
  if (VEC) return __ldcs(reinterpret_cast<const unsigned int*>(p + i));
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) v |= (uint32_t)(i + k < end ? p[i + k] : (uint8_t)fill) << (8 * k);
  return v;
}

template <bool VEC>
__device__ __forceinline__ void st_u8x4(uint8_t* p, int64_t i, int64_t end, uint32_t v) {
  if (VEC) {
    __stcs(reinterpret_cast<unsigned int*>(p + i), v);
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i + k < end) p[i + k] = (uint8_t)(v >> (8 * k));
}

template <bool VEC>
__device__ __forceinline__ void st_f32x4(float* p, int64_t i, int64_t end, const float (&v)[4]) {
  if (VEC) {
    __stcs(reinterpret_cast<float4*>(p + i), make_float4(v[0], v[1], v[2], v[3]));
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i + k < end) p[i + k] = v[k];
}

// GW-wide variants (GW = 4: the quad helpers above; GW = 2: half quads, for the 64-envs-per-warp build)
template <bool VEC, int GW>
__device__ __forceinline__ uint32_t ld_u8xg(const uint8_t* p, int64_t i, int64_t end, uint32_t fill) {
  if (GW == 4) return ld_u8x4<VEC>(p, i, end, fill);
  if (VEC) return (uint32_t)__ldcs(reinterpret_cast<const unsigned short*>(p + i));
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < GW; ++k) v |= (uint32_t)(i + k < end ? p[i + k] : (uint8_t)fill) << (8 * k);
  return v;
}
template <bool VEC, int GW>
__device__ __forceinline__ void st_u8xg(uint8_t* p, int64_t i, int64_t end, uint32_t v) {
  if (GW == 4) {
    st_u8x4<VEC>(p, i, end, v);
    return;
  }
  if (VEC) {
    __stcs(reinterpret_cast<unsigned short*>(p + i), (unsigned short)v);
    return;
  }
#pragma unroll
  for (int k = 0; k < GW; ++k)
    if (i + k < end) p[i + k] = (uint8_t)(v >> (8 * k));
}
template <bool VEC, int GW>
__device__ __forceinline__ void st_f32xg(float* p, int64_t i, int64_t end, const float (&v)[4]) {
  if (GW == 4) {
    st_f32x4<VEC>(p, i, end, v);
    return;
  }
  if (VEC) {
    __stcs(reinterpret_cast<float2*>(p + i), make_float2(v[0], v[1]));
    return;
  }
#pragma unroll
  for (int k = 0; k < GW; ++k)
    if (i + k < end) p[i + k] = v[k];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---- bulk asynchronous copy shared -> global (TMA engine, no tensor map: the tile is a flat byte range) ----
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#ifndef CX_OPT_L2HINT
#define CX_OPT_L2HINT 1   // bulk stores carry the L2 evict-first policy (0: default policy; development knob)
#endif
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes, uint64_t pol) {
#if CX_OPT_L2HINT
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the source bytes of every committed bulk store have been read: the tile may be modified again
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... of all but the N most recently committed groups (a ring of N + 1 tiles)
template <int N>
__device__ __forceinline__ void bulk_wait_read_pending() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// make this thread's generic-proxy shared-memory writes visible to the async proxy (the TMA engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// atomic max on a double that is only ever raised (stats slots start at -inf)
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

// Episode bookkeeping of one lane.  Lives in shared memory (touched only when an episode ends, about
// once per hundred steps) so that it costs no registers in the step loop; folded into the global
// statistics once per launch.
struct LaneStats {
  double sum, sumsq;
  uint32_t cnt, len;
  float mx, negmn;
  __device__ __forceinline__ void clear() {
    sum = sumsq = 0.0;
    cnt = len = 0;
    mx = negmn = -INFINITY;
  }
  __device__ __forceinline__ void episode(float ret, uint32_t steps) {
    cnt += 1;
    len += steps;
    sum += (double)ret;
    sumsq += (double)ret * (double)ret;
    mx = fmaxf(mx, ret);
    negmn = fmaxf(negmn, -ret);
  }
};

}  // namespace
