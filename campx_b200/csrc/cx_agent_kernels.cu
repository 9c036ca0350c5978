// Single-agent fast path (boat_race, Demo 1-5): fused T-step Engine.play() for a batch of envs.
//
// What one env-step does (reference lines restated on compressed state):
//   action dispatch            act[a] * shifted masks           examples/boat_race.py:40-49
//   candidate move             toroidal unit shift               boat_race.py:42-45 (quirk Q5)
//   wall gate                  b = gate*b + prev_pos*(1-gate)    boat_race.py:52-56, last-render layers
//   first-entry reward         (A.curtain * layers[tile]).sum()  boat_race.py:76-90 (quirk Q4)
//   plot directives            reward sum / game over / discount campx/plot.py:161-211, engine.py:285-290
//   z-ordered composition      painter's algorithm               campx/engine.py:306-321, rendering.py:128-178
//   time limit + auto reset    fresh make_game() per episode     examples/actor_critic.py:56,146-173
//
// B200 mapping.  The kernel is an HBM write stream (31 of the 33 algorithmic bytes per env-step are
// stores), so the design goal is: every global access is a full 16-byte-per-lane, 512-byte-per-warp
// coalesced transaction, and nothing ever waits on a block barrier.
//   * one WARP owns 256 consecutive envs for all T steps; its env state lives in registers;
//   * the static scene (backdrop + fixed drapes composed in z-order) is staged once in shared memory as
//     a pre-tiled byte image of the warp's 256 boards (256*cells bytes);  a step only un-pokes the
//     agent's old cell and pokes its new one (2 byte stores per env), then the warp streams the tile to
//     HBM with LDS.128 -> STG.128 (st.global.cs), 512 contiguous bytes per instruction;
//   * rewards / flags / actions are float4 / uchar4 per lane, laid out so that each instruction covers a
//     contiguous 512 B / 128 B span;
//   * warps only __syncwarp(); CTAs of 4 warps exist just to share the staged tables, so 1024 CTAs cover
//     2^20 envs in one resident wave on 148 SMs (7 CTAs/SM, 25.6 KB of tile per CTA).
#include "cx_internal.cuh"

namespace {

struct AgentParams {
  CxAgentHeader h;
  const uint8_t* blob;
  uint8_t* cell;     // [n]
  uint16_t* tstep;   // [n] (track)
  float* ret;        // [n] (track)
  double* stats;     // [CX_STATS_DOUBLES]
  const uint8_t* actions;  // [T, n]
  float* reward;           // [T, n]
  float* discount;         // [T, n] or null
  uint8_t* flags;          // [T, n]
  uint8_t* board;          // [T, n, cells]
  int64_t n;
  int32_t T;
  int32_t vec;  // all pointers 16B aligned and n % 16 == 0: use the vector path
};

constexpr int WT = CX_WARP_TILE_ENVS;       // envs per warp
constexpr int QUADS = WT / 128;             // quads (4 consecutive envs) per lane
constexpr int WARPS = CX_AGENT_CTA_THREADS / 32;

struct Tables {
  const uint8_t* nxt;
  const uint8_t* info;
  const uint8_t* basech;
  const float* rwc;
  const CxActionTable* act;
  int cells, A, K, agent_idx, agent_char, self_blocks, uses_old;
};

__device__ __forceinline__ uint32_t ld_u8x4(const uint8_t* p, int64_t i, int64_t end, bool vec, uint32_t fill) {
  if (vec) return __ldcs(reinterpret_cast<const unsigned int*>(p + i));
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) v |= (uint32_t)(i + k < end ? p[i + k] : (uint8_t)fill) << (8 * k);
  return v;
}

__device__ __forceinline__ void st_u8x4(uint8_t* p, int64_t i, int64_t end, bool vec, uint32_t v, bool stream) {
  if (vec) {
    if (stream)
      __stcs(reinterpret_cast<unsigned int*>(p + i), v);
    else
      *reinterpret_cast<unsigned int*>(p + i) = v;
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i + k < end) p[i + k] = (uint8_t)(v >> (8 * k));
}

__device__ __forceinline__ void st_f32x4(float* p, int64_t i, int64_t end, bool vec, const float (&v)[4]) {
  if (vec) {
    __stcs(reinterpret_cast<float4*>(p + i), make_float4(v[0], v[1], v[2], v[3]));
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (i + k < end) p[i + k] = v[k];
}

// One env, one Engine.play().  p: agent cell (CX_EMPTY_CELL = empty mask).
__device__ __forceinline__ void agent_env_step(const Tables& S, uint32_t a, uint32_t& p,
                                               float& reward, uint32_t& flags, float& disc) {
  if (a >= (uint32_t)S.A) {  // outside the action set: the reference would fail inside update(); leave the env alone
    reward = 0.0f;
    disc = 1.0f;
    flags = CX_FLAG_BAD_ACTION | CX_FLAG_REWARD_NONE;
    return;
  }
  uint32_t ko = S.K - 1, kn = S.K - 1;  // "no cell"
  if (p != CX_EMPTY_CELL) {
    const uint32_t ip = S.info[p];
    const bool visp = ip >> 7;                        // the agent was visible in the last render
    const uint32_t t = S.nxt[a * S.cells + p];        // shifted mask (boat_race.py:42-49)
    const uint32_t it = S.info[t];
    const bool onto_self = (t == p) && visp;          // the last render showed the agent itself there
    const bool blocked = onto_self ? (S.self_blocks != 0) : ((it >> 5) & 1);  // layers[c] at the target (:54)
    ko = visp ? S.agent_idx : (ip & 31);
    if (blocked) {
      // b = prev_pos: the agent layer of the last render -- empty when the agent was occluded (:55-56)
      if (visp)
        kn = S.agent_idx;
      else
        p = CX_EMPTY_CELL;
    } else {
      p = t;
      kn = onto_self ? S.agent_idx : (it & 31);
    }
  }
  if (!S.uses_old) ko = 0;
  reward = S.rwc[(a * S.K + ko) * S.K + kn];
  disc = S.act->discount[a];
  flags = (S.act->over[a] ? CX_FLAG_TERMINATED : 0) | (S.act->reward_none[a] ? CX_FLAG_REWARD_NONE : 0);
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

template <bool TRACK>
__global__ void __launch_bounds__(CX_AGENT_CTA_THREADS, 7)  // 7 CTAs/SM: 1024 CTAs (2^20 envs) in one wave
k_agent_rollout(const __grid_constant__ AgentParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const CxAgentHeader& H = P.h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cells = H.cells;

  // ---- stage the static tables (next-cell, cell info, base board, tile pattern, reward table) ----
  {
    const uint4* src = reinterpret_cast<const uint4*>(P.blob);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < H.blob_bytes / 16; i += CX_AGENT_CTA_THREADS) dst[i] = src[i];
  }
  __syncthreads();  // the only block barrier; warps are independent from here on

  const int64_t env0 = ((int64_t)blockIdx.x * WARPS + warp) * WT;
  if (env0 >= P.n) return;
  const int nenv = (int)min((int64_t)WT, P.n - env0);
  const bool vec = P.vec != 0;

  Tables S;
  S.nxt = smem + H.off_nxt;
  S.info = smem + H.off_info;
  S.basech = smem + H.off_basech;
  S.rwc = reinterpret_cast<const float*>(smem + H.off_rwc);
  S.act = reinterpret_cast<const CxActionTable*>(smem + H.off_act);
  S.cells = cells;
  S.A = H.n_actions;
  S.K = H.n_chars + 1;
  S.agent_idx = H.agent_idx;
  S.agent_char = H.agent_char;
  S.self_blocks = H.self_blocks;
  S.uses_old = H.uses_old;

  uint8_t* tile = smem + H.blob_bytes + (size_t)warp * (WT * cells);
  {  // pre-tile the static scene: 256 copies of the base board, written as 16-byte pattern chunks
    const uint4* pat = reinterpret_cast<const uint4*>(smem + H.off_pat);
    uint4* t16 = reinterpret_cast<uint4*>(tile);
    const int nchunks = WT * cells / 16;
    for (int k = lane; k < nchunks; k += 32) t16[k] = pat[(16 * k) % cells];
  }
  __syncwarp();

  // ---- load env state into registers; paint the agents into the tile ----
  uint32_t cellq[QUADS];   // 4 agent cells per quad, one byte each
  uint32_t shownq[QUADS];  // cell currently drawn in the tile (CX_EMPTY_CELL: none)
  uint16_t ts[QUADS][4];
  float rt[QUADS][4];
#pragma unroll
  for (int j = 0; j < QUADS; ++j) {
    const int el = j * 128 + lane * 4;
    cellq[j] = ld_u8x4(P.cell, env0 + el, P.n, vec && el < nenv, CX_EMPTY_CELL);
    shownq[j] = 0xFFFFFFFFu;
    if (TRACK && vec && el < nenv) {  // 8-byte / 16-byte state loads
      const uint2 tv = *reinterpret_cast<const uint2*>(P.tstep + env0 + el);
      const float4 rv = *reinterpret_cast<const float4*>(P.ret + env0 + el);
      ts[j][0] = (uint16_t)tv.x; ts[j][1] = (uint16_t)(tv.x >> 16);
      ts[j][2] = (uint16_t)tv.y; ts[j][3] = (uint16_t)(tv.y >> 16);
      rt[j][0] = rv.x; rt[j][1] = rv.y; rt[j][2] = rv.z; rt[j][3] = rv.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = el + i < nenv;
      if (TRACK && !(vec && el < nenv)) {
        ts[j][i] = valid ? P.tstep[env0 + el + i] : (uint16_t)0;
        rt[j][i] = valid ? P.ret[env0 + el + i] : 0.0f;
      }
      const uint32_t p = (cellq[j] >> (8 * i)) & 0xFF;
      if (valid && p != CX_EMPTY_CELL && (S.info[p] >> 7)) {
        tile[(el + i) * cells + p] = (uint8_t)S.agent_char;
        shownq[j] = (shownq[j] & ~(0xFFu << (8 * i))) | (p << (8 * i));
      }
    }
  }

  // episode statistics: per lane here, one warp reduction + a handful of atomics at the end
  uint32_t ep_cnt = 0, ep_len = 0;
  double ep_sum = 0.0, ep_sumsq = 0.0;
  float ep_max = -INFINITY, ep_negmin = -INFINITY;

  uint32_t actq[QUADS];
#pragma unroll
  for (int j = 0; j < QUADS; ++j) {
    const int el = j * 128 + lane * 4;
    actq[j] = el < nenv ? ld_u8x4(P.actions, env0 + el, P.n, vec, 0) : 0u;
  }

  for (int t = 0; t < P.T; ++t) {
    const int64_t row = (int64_t)t * P.n + env0;  // index of this warp's first env in [T, n] arrays
    const int64_t row_end = (int64_t)(t + 1) * P.n;
#pragma unroll
    for (int j = 0; j < QUADS; ++j) {
      const int el = j * 128 + lane * 4;
      if (el < nenv) {
        float rw[4], dc[4];
        uint32_t fl = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!vec && el + i >= nenv) {  // tail quad on the scalar path: env does not exist
            rw[i] = 0.0f;
            dc[i] = 0.0f;
            continue;
          }
          const uint32_t a = (actq[j] >> (8 * i)) & 0xFF;
          uint32_t p = (cellq[j] >> (8 * i)) & 0xFF;
          uint32_t f;
          if (TRACK && (ts[j][i] & CX_OVER_BIT)) {  // auto_reset == 0 and the episode ended: frozen env
            rw[i] = 0.0f;
            dc[i] = 0.0f;
            f = CX_FLAG_ALREADY_OVER | CX_FLAG_REWARD_NONE;
          } else {
            agent_env_step(S, a, p, rw[i], f, dc[i]);
          }
          const uint32_t show = p;  // the cell this step's board shows (terminal board on a terminal step)
          if (TRACK && !(f & (CX_FLAG_BAD_ACTION | CX_FLAG_ALREADY_OVER))) {
            const uint32_t steps = ts[j][i] + 1u;
            rt[j][i] += rw[i];
            if (!(f & CX_FLAG_TERMINATED) && H.max_steps > 0 && steps >= (uint32_t)H.max_steps)
              f |= CX_FLAG_TRUNCATED;  // time limit: done, discount untouched
            ts[j][i] = (uint16_t)steps;
            if (f & (CX_FLAG_TERMINATED | CX_FLAG_TRUNCATED)) {
              const float r = rt[j][i];
              ep_cnt += 1;
              ep_len += steps;
              ep_sum += (double)r;
              ep_sumsq += (double)r * (double)r;
              ep_max = fmaxf(ep_max, r);
              ep_negmin = fmaxf(ep_negmin, -r);
              if (H.auto_reset) {  // the next play() starts from the its_showtime state
                p = H.init_cell;
                ts[j][i] = 0;
                rt[j][i] = 0.0f;
              } else {
                ts[j][i] |= CX_OVER_BIT;
              }
            }
          }
          fl |= f << (8 * i);
          cellq[j] = (cellq[j] & ~(0xFFu << (8 * i))) | (p << (8 * i));
          // re-compose this env's board: base character back where the agent was drawn, agent character
          // where it is visible now (painter's algorithm collapsed to two byte stores)
          const uint32_t drawn = (shownq[j] >> (8 * i)) & 0xFF;
          const uint32_t now = (show != CX_EMPTY_CELL && (S.info[show] >> 7)) ? show : CX_EMPTY_CELL;
          if (drawn != now) {
            uint8_t* b = tile + (el + i) * cells;
            if (drawn != CX_EMPTY_CELL) b[drawn] = S.basech[drawn];
            if (now != CX_EMPTY_CELL) b[now] = (uint8_t)S.agent_char;
            shownq[j] = (shownq[j] & ~(0xFFu << (8 * i))) | (now << (8 * i));
          }
        }
        st_f32x4(P.reward, row + el, row_end, vec, rw);
        if (P.discount) st_f32x4(P.discount, row + el, row_end, vec, dc);
        st_u8x4(P.flags, row + el, row_end, vec, fl, true);
      }
    }
    // next step's actions: issue the loads before streaming the tile so their latency is hidden
    if (t + 1 < P.T) {
#pragma unroll
      for (int j = 0; j < QUADS; ++j) {
        const int el = j * 128 + lane * 4;
        if (el < nenv) actq[j] = ld_u8x4(P.actions, row + P.n + el, row_end + P.n, vec, 0);
      }
    }
    __syncwarp();
    // ---- stream the finished boards of this warp's envs to HBM ----
    {
      uint8_t* dst = P.board + row * cells;
      const int nbytes = nenv * cells;
      if (vec) {
        const uint4* t16 = reinterpret_cast<const uint4*>(tile);
        uint4* d16 = reinterpret_cast<uint4*>(dst);
        const int nchunks = nbytes / 16;  // exact: nenv % 16 == 0 on the vector path
#pragma unroll 4
        for (int k = lane; k < nchunks; k += 32) __stcs(d16 + k, t16[k]);
      } else {
        for (int k = lane; k < nbytes; k += 32) dst[k] = tile[k];
      }
    }
    __syncwarp();
  }

  // ---- write the env state back ----
#pragma unroll
  for (int j = 0; j < QUADS; ++j) {
    const int el = j * 128 + lane * 4;
    if (el < nenv) {
      st_u8x4(P.cell, env0 + el, P.n, vec, cellq[j], false);
      if (TRACK && vec) {
        *reinterpret_cast<uint2*>(P.tstep + env0 + el) =
            make_uint2((uint32_t)ts[j][0] | ((uint32_t)ts[j][1] << 16), (uint32_t)ts[j][2] | ((uint32_t)ts[j][3] << 16));
        *reinterpret_cast<float4*>(P.ret + env0 + el) = make_float4(rt[j][0], rt[j][1], rt[j][2], rt[j][3]);
      } else if (TRACK) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (el + i < nenv) {
            P.tstep[env0 + el + i] = ts[j][i];
            P.ret[env0 + el + i] = rt[j][i];
          }
      }
    }
  }
  if (TRACK) {
    const double cnt = warp_sum((double)ep_cnt), len = warp_sum((double)ep_len);
    const double sum = warp_sum(ep_sum), sumsq = warp_sum(ep_sumsq);
    const float mx = warp_max(ep_max), ngmn = warp_max(ep_negmin);
    if (lane == 0) {
      if (cnt > 0.0) {
        atomicAdd(P.stats + CX_STAT_EPISODES, cnt);
        atomicAdd(P.stats + CX_STAT_RETURN_SUM, sum);
        atomicAdd(P.stats + CX_STAT_RETURN_SUMSQ, sumsq);
        atomicAdd(P.stats + CX_STAT_LENGTH_SUM, len);
        atomic_max_double(P.stats + CX_STAT_RETURN_MAX, (double)mx);
        atomic_max_double(P.stats + CX_STAT_NEG_RETURN_MIN, (double)ngmn);
      }
      atomicAdd(P.stats + CX_STAT_ENV_STEPS, (double)nenv * (double)P.T);
    }
  }
}

}  // namespace

int cx_launch_agent_rollout(const cx_game* g, void* d_state, int64_t n, int32_t T, const uint8_t* d_actions,
                            float* d_reward, float* d_discount, uint8_t* d_flags, uint8_t* d_board,
                            cudaStream_t s) {
  const CxStateLayout L = cx_layout(g, n);
  uint8_t* base = static_cast<uint8_t*>(d_state);
  AgentParams P;
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
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  P.vec = (n % 16 == 0) && al16(d_actions) && al16(d_reward) && al16(d_discount) && al16(d_flags) && al16(d_board);

  const size_t smem = (size_t)g->ah.blob_bytes + (size_t)WARPS * WT * g->ah.cells;
  const int64_t warps = (n + WT - 1) / WT;
  const int64_t grid = (warps + WARPS - 1) / WARPS;
  if (grid > 0x7fffffff) {
    cx_set_error("cx_rollout: too many environments for one launch");
    return CX_ERR_INVALID_ARG;
  }
  static bool configured = false;  // raise the dynamic shared memory cap once (it reserves nothing)
  if (!configured) {
    const int cap = 227 * 1024;
    CX_CUDA_OK(cudaFuncSetAttribute(k_agent_rollout<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    CX_CUDA_OK(cudaFuncSetAttribute(k_agent_rollout<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    configured = true;
  }
  if (g->ah.track)
    k_agent_rollout<true><<<(unsigned)grid, CX_AGENT_CTA_THREADS, smem, s>>>(P);
  else
    k_agent_rollout<false><<<(unsigned)grid, CX_AGENT_CTA_THREADS, smem, s>>>(P);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}
