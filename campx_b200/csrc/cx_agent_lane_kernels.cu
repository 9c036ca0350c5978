// Single-agent fast path for SMALL batches: fused T-step Engine.play(), lane = env.
//
// Same step as k_agent_rollout (cx_agent_kernels.cu: one (action, cell) table look-up = action dispatch, toroidal move,
// wall gate, entry rewards, directives -- examples/boat_race.py:35-91, campx/plot.py:161-211, engine.py:285-290 --
// then the painter's algorithm collapsed to two byte pokes into a pre-tiled image of the static scene,
// engine.py:306-321), written for the regime where that kernel is NOT bandwidth-bound.
//
// Why a third build.  BASELINE config 1 (Demo 1, 65,536 envs) gives an SM 443 envs: one step of the whole batch is
// 2 MB, 0.31 us at the HBM peak = 620 SM clocks.  The tile kernels spend more than that on the step's dependent
// instruction chain: the 64-envs-per-warp build issues 214 warp instructions per warp-step with 7 warps per SM
// (ncu: issue slots 30 % busy, 4.4 stall cycles per issue, `wait` = fixed-latency dependencies on top), and every TMA
// bulk store costs its issuing warp ~240 clocks whatever its size (scripts/chain_probe.cu: STS + fence.proxy.async +
// UBLKCP + wait_group.read with two tiles = 239 clocks per iteration for 800, 1,600 or 6,400 bytes; the fence alone is
// 13-27 clocks and does not wait for global stores).  So for small batches this kernel
//   * gives every env its own lane (32 envs per warp: twice the warps of the 64-env build, no per-quad unpacking),
//   * keeps the step's common case to a handful of instructions: the rare events (episode end, time limit, bad action,
//     frozen env) are ONE test and one out-of-line branch,
//   * ships the warp's board tile (32 * cells bytes, a 32-byte-aligned run of whole sectors) with LDS.128 -> STG.128
//     straight from the generic proxy: no proxy fence, no TMA issue latency, no read-wait; two tiles alternate so that
//     one __syncwarp per step orders the pokes of step t+2 behind the reads of step t.
// Large batches stay on k_agent_rollout (TMA bulk stores: fewer LSU instructions once the SM is full of warps).
#include <stdlib.h>

#include "cx_agent_common.cuh"
#include "cx_philox.cuh"

namespace {

struct LaneParams {
  CxAgentHeader h;
  const uint8_t* blob;
  uint8_t* cell;
  uint16_t* tstep;
  float* ret;
  double* stats;
  const uint8_t* actions;  // [T, n]
  float* reward;           // [T, n]
  float* discount;         // [T, n] or null
  uint8_t* flags;          // [T, n]
  uint8_t* board;          // [T, n, cells]
  int64_t n;               // a multiple of 32
  int32_t T;
  uint64_t seed, env_offset, t0;
  uint8_t* actions_out;    // SYNTH: [T, n] or null
};

constexpr int LANE_MAX_WARPS = 4;
constexpr int LANE_UNR = 4;   // steps per unrolled group (tile parity and action registers are indexed statically)

template <bool TRACK, bool SYNTH>
__global__ void __launch_bounds__(LANE_MAX_WARPS * 32) k_agent_rollout_lane(const __grid_constant__ LaneParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const CxAgentHeader& H = P.h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, WARPS = blockDim.x >> 5;
  const int cells = H.cells;

  asm volatile("griddepcontrol.launch_dependents;");
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < H.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();  // the only block barrier

  const int64_t n = P.n;
  const int64_t env0 = ((int64_t)blockIdx.x * WARPS + warp) * 32;
  if (env0 >= n) return;
  const int64_t env = env0 + lane;

  const uint32_t* __restrict__ s_tt = reinterpret_cast<const uint32_t*>(smem + H.off_tt);
  const float* __restrict__ s_tr = reinterpret_cast<const float*>(smem + H.off_tr);
  const float* __restrict__ s_td = reinterpret_cast<const float*>(smem + H.off_td);
  const uint8_t* __restrict__ s_basech = smem + H.off_basech;
  const uint8_t* __restrict__ s_shown = smem + H.off_shown;
  const uint32_t stride = H.stride, n_actions = H.n_actions;
  const uint8_t agent_char = (uint8_t)H.agent_char;
  const uint32_t none = cells;
  // the step counter's common case ends one short of the time limit / of saturation
  const uint32_t limit = min(H.max_steps > 0 ? (uint32_t)H.max_steps : 0xFFFFFFFFu, (uint32_t)CX_STEP_MAX);
  const uint32_t max_steps = H.max_steps > 0 ? (uint32_t)H.max_steps : 0xFFFFFFFFu;
  const bool want_discount = P.discount != nullptr;

  // two tiles per warp, back to back: 2 * 32 boards of the static scene, written as 16-byte pattern chunks
  const int tile_bytes = 32 * cells, nch = tile_bytes / 16;   // 32 * cells is a multiple of 32
  const int tile0 = H.blob_bytes + warp * 2 * tile_bytes;     // byte offset of this warp's tile 0 in smem[]
  {
    const uint4* pat = reinterpret_cast<const uint4*>(smem + H.off_pat);
    uint4* t16 = reinterpret_cast<uint4*>(smem + tile0);
    for (int k = lane; k < 2 * nch; k += 32) t16[k] = pat[(16 * k) % cells];
  }
  __syncwarp();

  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint32_t cell = min((uint32_t)P.cell[env], none);
  uint32_t ts = 0;
  float rt = 0.0f;
  if (TRACK) {
    ts = P.tstep[env];
    rt = P.ret[env];
  }
  const int mine0 = tile0 + lane * cells;                 // this env's board in tile 0; tile 1 is tile_bytes further
  uint32_t drawn[2];
  drawn[0] = drawn[1] = s_shown[cell];
  if (drawn[0] != none) {
    smem[mine0 + drawn[0]] = agent_char;
    smem[mine0 + tile_bytes + drawn[0]] = agent_char;
  }
  LaneStats& stats = reinterpret_cast<LaneStats*>(smem + H.blob_bytes + (size_t)WARPS * 2 * tile_bytes)[tid];
  if (TRACK) stats.clear();

  auto fetch_action = [&](int t) -> uint32_t {
    if (t >= P.T) return 0u;
    if (SYNTH) {
      const uint32_t a = cx_synth_action(P.seed, P.env_offset + (uint64_t)env, P.t0 + (uint64_t)t, n_actions);
      if (P.actions_out) P.actions_out[(int64_t)t * n + env] = (uint8_t)a;
      return a;
    }
    return __ldcs(P.actions + (int64_t)t * n + env);
  };
  uint32_t act[LANE_UNR], act_next[LANE_UNR];
#pragma unroll
  for (int u = 0; u < LANE_UNR; ++u) act[u] = fetch_action(u);

  // running pointers: one 64-bit add per stream and step
  float* p_rw = P.reward + env;
  float* p_dc = want_discount ? P.discount + env : nullptr;
  uint8_t* p_fl = P.flags + env;
  uint4* p_bd = reinterpret_cast<uint4*>(P.board + env0 * cells) + lane;
  const int64_t bd_step = n * cells / 16;   // uint4 per [n, cells] row: n % 32 == 0
  const bool small_tile = nch <= 64;

  for (int t0 = 0; t0 < P.T; t0 += LANE_UNR) {
#pragma unroll
    for (int u = 0; u < LANE_UNR; ++u) act_next[u] = fetch_action(t0 + LANE_UNR + u);
#pragma unroll
    for (int u = 0; u < LANE_UNR; ++u) {
      if (t0 + u >= P.T) break;
      // ---- the env's step: one table look-up ----
      const uint32_t a = min(act[u], n_actions);
      const uint32_t idx = a * stride + cell;
      uint32_t e = s_tt[idx];
      float rw = s_tr[idx];
      float dc = 1.0f;
      if (want_discount) dc = s_td[H.td_per_cell ? idx : a];
      uint32_t f = e >> 16;
      uint32_t p = e & 0xFF, show = (e >> 8) & 0xFF;
      if (TRACK) {
        const uint32_t steps = ts + 1u;
        // rare: episode end / bad action / frozen env (flag bits or bit 15 of the counter), time limit, saturation
        const uint32_t special = (f & (CX_FLAG_TERMINATED | CX_FLAG_BAD_ACTION | CX_FLAG_ALREADY_OVER)) | (ts & CX_OVER_BIT);
        if (special == 0 && steps < limit) {
          ts = steps;
          rt += rw;
        } else {
          if (ts & CX_OVER_BIT) {  // auto_reset == 0 and the episode ended: frozen env
            p = cell;
            show = drawn[(u + 1) & 1];   // where the last frame drew it
            f = CX_FLAG_ALREADY_OVER | CX_FLAG_REWARD_NONE;
            rw = 0.0f;
            dc = 0.0f;
          }
          if (!(f & (CX_FLAG_BAD_ACTION | CX_FLAG_ALREADY_OVER))) {
            const uint32_t st = min(steps, (uint32_t)CX_STEP_MAX);
            rt += rw;
            if (!(f & CX_FLAG_TERMINATED) && st >= max_steps) f |= CX_FLAG_TRUNCATED;
            ts = st;
            if (f & (CX_FLAG_TERMINATED | CX_FLAG_TRUNCATED)) {
              stats.episode(rt, st);
              if (H.auto_reset) {
                p = H.init_cell;
                ts = 0;
                rt = 0.0f;
              } else {
                ts |= CX_OVER_BIT;
              }
            }
          }
        }
      }
      cell = p;
      __stcs(p_rw, rw);
      p_rw += n;
      if (want_discount) {
        __stcs(p_dc, dc);
        p_dc += n;
      }
      *p_fl = (uint8_t)f;
      p_fl += n;

      // ---- the boards: poke tile (t & 1), then the warp copies it out, 512 contiguous bytes per instruction ----
      const int mine = mine0 + (u & 1) * tile_bytes;
      const uint32_t was = drawn[u & 1];
      if (was != show) {
        if (was != none) smem[mine + was] = s_basech[was];
        if (show != none) smem[mine + show] = agent_char;
        drawn[u & 1] = show;
      }
      __syncwarp();
      const uint4* t16 = reinterpret_cast<const uint4*>(smem + tile0 + (u & 1) * tile_bytes) + lane;
      if (small_tile) {   // boards of up to 32 cells: at most two chunks per lane
        if (lane < nch) __stcs(p_bd, t16[0]);
        if (lane + 32 < nch) __stcs(p_bd + 32, t16[32]);
      } else {
        for (int k = 0; k + lane < nch; k += 32) __stcs(p_bd + k, t16[k]);
      }
      p_bd += bd_step;
    }
#pragma unroll
    for (int u = 0; u < LANE_UNR; ++u) act[u] = act_next[u];
  }

  P.cell[env] = (uint8_t)cell;
  if (TRACK) {
    P.tstep[env] = (uint16_t)ts;
    P.ret[env] = rt;
    const double cnt = warp_sum((double)stats.cnt), len = warp_sum((double)stats.len);
    const double sum = warp_sum(stats.sum), sumsq = warp_sum(stats.sumsq);
    const float mx = warp_max(stats.mx), ngmn = warp_max(stats.negmn);
    if (lane == 0) {
      double* sp = cx_stat_stripe(P.stats, (uint32_t)(blockIdx.x * WARPS + warp));
      if (cnt > 0.0) {
        atomicAdd(sp + CX_STAT_EPISODES, cnt);
        atomicAdd(sp + CX_STAT_RETURN_SUM, sum);
        atomicAdd(sp + CX_STAT_RETURN_SUMSQ, sumsq);
        atomicAdd(sp + CX_STAT_LENGTH_SUM, len);
        atomic_max_double(sp + CX_STAT_RETURN_MAX, (double)mx);
        atomic_max_double(sp + CX_STAT_NEG_RETURN_MIN, (double)ngmn);
      }
      atomicAdd(sp + CX_STAT_ENV_STEPS, 32.0 * (double)P.T);
    }
  }
}

size_t lane_smem_bytes(const cx_game* g, int warps) {
  return (size_t)g->ah.blob_bytes + (size_t)warps * 2 * 32 * g->ah.cells +
         (g->ah.track ? (size_t)warps * 32 * sizeof(LaneStats) : 0);
}

template <bool TRACK, bool SYNTH>
int launch_lane(const LaneParams& P, unsigned grid, unsigned block, size_t smem, cudaStream_t s) {
  static CxPerDevice configured;
  if (configured.need()) {
    CX_CUDA_OK(cudaFuncSetAttribute(k_agent_rollout_lane<TRACK, SYNTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    227 * 1024));
    configured.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CX_CUDA_OK(cudaLaunchKernelEx(&cfg, k_agent_rollout_lane<TRACK, SYNTH>, P));
  return CX_OK;
}

}  // namespace

// whole warps of 32 envs, every [T, n] / [T, n, cells] row 16-byte aligned, tiles that fit shared memory
bool cx_agent_lane_applies(const cx_game* g, int64_t n, const void* d_actions, const void* d_actions_out,
                           const void* d_reward, const void* d_discount, const void* d_flags, const void* d_board) {
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return g->path == CX_PATH_AGENT && n % 32 == 0 && n > 0 && lane_smem_bytes(g, 1) <= 64 * 1024 && al16(d_actions) &&
         al16(d_actions_out) && al16(d_reward) && al16(d_discount) && al16(d_flags) && al16(d_board);
}

int cx_launch_agent_rollout_lane(const cx_game* g, void* d_state, int64_t n, int32_t T, const uint8_t* d_actions,
                                 const CxSynth& synth, float* d_reward, float* d_discount, uint8_t* d_flags,
                                 uint8_t* d_board, cudaStream_t s) {
  const CxStateLayout L = cx_layout(g, n);
  uint8_t* base = static_cast<uint8_t*>(d_state);
  LaneParams P;
  P.h = g->ah;
  P.blob = g->d_blob;
  P.cell = base + L.off_dyn;
  P.tstep = reinterpret_cast<uint16_t*>(base + L.off_tstep);
  P.ret = reinterpret_cast<float*>(base + L.off_ret);
  P.stats = reinterpret_cast<double*>(base + L.off_stats);
  P.actions = d_actions;
  P.reward = d_reward;
  P.discount = d_discount;
  P.flags = d_flags;
  P.board = d_board;
  P.n = n;
  P.T = T;
  P.seed = synth.seed;
  P.env_offset = synth.env_offset;
  P.t0 = synth.t0;
  P.actions_out = synth.actions_out;
  const int64_t warps = n / 32;
  // warps per CTA: 4, fewer while that leaves the grid under ~8 CTAs per SM (small grids spread more evenly)
  int wpc = LANE_MAX_WARPS;
  while (wpc > 1 && (warps + wpc - 1) / wpc < (int64_t)g->sm_count * 8) wpc >>= 1;
  if (const char* dbg = getenv("CX_LANE_WPC")) {
    const int w = atoi(dbg);
    if (w == 1 || w == 2 || w == 4) wpc = w;
  }
  const int64_t grid = (warps + wpc - 1) / wpc;
  if (grid > 0x7fffffff) {
    cx_set_error("cx_rollout: too many environments for one launch");
    return CX_ERR_INVALID_ARG;
  }
  const size_t smem = lane_smem_bytes(g, wpc);
  const bool track = g->ah.track != 0;
  if (synth.on)
    return track ? launch_lane<true, true>(P, (unsigned)grid, 32u * wpc, smem, s)
                 : launch_lane<false, true>(P, (unsigned)grid, 32u * wpc, smem, s);
  return track ? launch_lane<true, false>(P, (unsigned)grid, 32u * wpc, smem, s)
               : launch_lane<false, false>(P, (unsigned)grid, 32u * wpc, smem, s);
}
