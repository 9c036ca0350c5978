// Generic path: any game made of the four entity primitives (static drape, one-cell drape, rolling
// drape, sprite), several moving entities, several update groups, and the backdrop-stamping quirk Q1.
// Used by Hello World (13x36: RollingDrape + 4 SlidingSprites, `Hello World Example.ipynb` cells 3-4).
//
// Reference lines restated per env-step:
//   update order / groups      campx/engine.py:184-208 (re-render after every group)
//   RollingDrape / SlidingSprite   Hello World notebook cell 3 (np.roll of the mask; (row,col) += d mod size)
//   one-cell drape with wall gate  examples/boat_race.py:35-59
//   entry rewards              boat_race.py:76-90, Demo 3 cell 3
//   painter's algorithm        engine.py:306-321 + rendering.py:111,128,150,173-178 -- including the
//                              storage aliasing: sprites painted before the first drape write into the
//                              backdrop itself (per-env backdrop plane `dynbd`), and a game without any
//                              drape has its canvas zeroed at every render
//   plot directives            campx/plot.py:161-257, engine.py:285-290
//
// Layout: one CTA owns CX_GEN_TILE_ENVS consecutive envs for all T steps.  Phase 1 (one thread per env)
// advances the few bytes of entity state held in shared memory; phase 2 (all threads) composes the
// boards 16 bytes at a time and streams them out with coalesced 128-bit stores.
#include "cx_internal.cuh"

namespace {

constexpr int G = CX_GEN_TILE_ENVS;
constexpr int NT = CX_GEN_CTA_THREADS;

struct GenParams {
  CxGenHeader h;
  const uint8_t* blob;
  uint16_t* dyn;     // [n_dyn][n]
  uint16_t* tstep;   // [n]
  float* ret;        // [n]
  uint8_t* dynbd;    // [n][cells]
  double* stats;
  const uint8_t* actions;
  float* reward;
  float* discount;
  uint8_t* flags;
  uint8_t* board;
  int64_t n;
  int32_t T;
  int32_t vec;
};

struct Ctx {
  const CxGenHeader* H;
  const uint32_t* masks;
  const uint8_t* backdrop;
  const float* entry;
  const uint16_t* rc;
  const uint8_t* chidx;  // [256] char code -> game char index (0xFF: not a game char)
};

__device__ __forceinline__ bool mask_bit(const Ctx& X, int z, int cell) {
  return (X.masks[z * X.H->mask_words + (cell >> 5)] >> (cell & 31)) & 1u;
}

// Painter's algorithm at one cell on top of backdrop byte v (engine.py:310-321).
__device__ __forceinline__ uint8_t overlay(const Ctx& X, const uint16_t* st, int cell, uint8_t v) {
  const CxGenHeader& H = *X.H;
  for (int z = 0; z < H.n_ent; ++z) {
    const CxGenEntity& e = H.ent[z];
    bool covers;
    switch (e.kind) {
      case CX_KIND_STATIC:
        covers = mask_bit(X, z, cell);
        break;
      case CX_KIND_ROLL: {
        const uint32_t off = st[e.dyn_slot], rcv = X.rc[cell];
        int r = (int)(rcv >> 8) - (int)(off >> 8), c = (int)(rcv & 255) - (int)(off & 255);
        if (r < 0) r += H.rows;
        if (c < 0) c += H.cols;
        covers = mask_bit(X, z, r * H.cols + c);
        break;
      }
      default:  // CELL drape or SPRITE: one cell
        covers = e.visible && st[e.dyn_slot] == cell;
    }
    if (covers) v = e.ch;
  }
  return v;
}

__device__ __forceinline__ uint8_t backdrop_at(const Ctx& X, const GenParams& P, int64_t env, int cell) {
  return X.H->has_dynbd ? P.dynbd[env * X.H->cells + cell] : X.backdrop[cell];
}

// rendering.py:150 through the alias of :128 -- sprites behind the first drape paint into the backdrop
__device__ __forceinline__ void stamp(const Ctx& X, const GenParams& P, const uint16_t* st, int64_t env) {
  const CxGenHeader& H = *X.H;
  if (!H.has_dynbd) return;
  for (int z = 0; z < H.n_ent; ++z) {
    const CxGenEntity& e = H.ent[z];
    if (e.stamps) P.dynbd[env * H.cells + st[e.dyn_slot]] = e.ch;
  }
}

__device__ __forceinline__ uint32_t wrap_move(const CxGenHeader& H, uint32_t r, uint32_t c, int dr, int dc,
                                              uint32_t* rr, uint32_t* cc) {
  int nr = (int)r + dr, nc = (int)c + dc;
  nr %= H.rows;
  nc %= H.cols;
  if (nr < 0) nr += H.rows;
  if (nc < 0) nc += H.cols;
  *rr = nr;
  *cc = nc;
  return nr * H.cols + nc;
}

// One env, one Engine.play(): entity updates in schedule order with a render after every update group.
// st: current state (updated in place); prev: scratch snapshot of the state at the last render.
__device__ void generic_env_step(const Ctx& X, const GenParams& P, int64_t env, uint32_t a, uint16_t* st,
                                 uint16_t* prev, float& reward, uint32_t& flags, float& disc) {
  const CxGenHeader& H = *X.H;
  if (a >= (uint32_t)H.n_actions) {
    reward = 0.0f;
    disc = 1.0f;
    flags = CX_FLAG_BAD_ACTION | CX_FLAG_REWARD_NONE;
    return;
  }
  for (int d = 0; d < H.n_dyn; ++d) prev[d] = st[d];
  float summed = 0.0f;
  bool first = true;
  int group = H.n_ent > 0 ? H.ent[H.update_order[0]].group : 0;
  for (int i = 0; i < H.n_ent; ++i) {
    const int z = H.update_order[i];
    const CxGenEntity& e = H.ent[z];
    if (e.group != group) {  // engine.py:208: the board is re-rendered before the next group updates
      stamp(X, P, st, env);
      for (int d = 0; d < H.n_dyn; ++d) prev[d] = st[d];
      group = e.group;
    }
    const int dr = e.dr[a], dc = e.dc[a];
    if (e.kind == CX_KIND_CELL) {
      const uint32_t p = st[e.dyn_slot];
      if (p != CX_EMPTY_CELL16) {
        uint32_t rr, cc;
        const uint32_t rcv = X.rc[p];
        uint32_t t = wrap_move(H, rcv >> 8, rcv & 255, dr, dc, &rr, &cc);
        if (e.blockers) {
          const uint8_t seen = overlay(X, prev, (int)t, backdrop_at(X, P, env, (int)t));
          const uint8_t k = X.chidx[seen];
          if (k != 0xFF && ((e.blockers >> k) & 1u)) {
            // fall back to the agent layer of the last render (boat_race.py:55-56)
            const uint32_t pp = prev[e.dyn_slot];
            const bool vis = pp != CX_EMPTY_CELL16 &&
                             overlay(X, prev, (int)pp, backdrop_at(X, P, env, (int)pp)) == e.ch;
            t = vis ? pp : CX_EMPTY_CELL16;
          }
        }
        st[e.dyn_slot] = (uint16_t)t;
      }
    } else if (e.kind == CX_KIND_SPRITE) {
      const uint32_t rcv = X.rc[st[e.dyn_slot]];
      uint32_t rr, cc;
      st[e.dyn_slot] = (uint16_t)wrap_move(H, rcv >> 8, rcv & 255, dr, dc, &rr, &cc);
    } else if (e.kind == CX_KIND_ROLL) {
      const uint32_t off = st[e.dyn_slot];
      uint32_t rr, cc;
      wrap_move(H, off >> 8, off & 255, dr, dc, &rr, &cc);
      st[e.dyn_slot] = (uint16_t)((rr << 8) | cc);
    }
    if ((e.reward_actions >> a) & 1u) {
      float r = e.step_reward[a];
      if (e.watch != 0xFF) {
        const uint32_t wc = st[H.ent[e.watch].dyn_slot];  // things[...] is current, not last-render, state
        if (wc != CX_EMPTY_CELL16) {
          const uint8_t k = X.chidx[overlay(X, prev, (int)wc, backdrop_at(X, P, env, (int)wc))];
          if (k != 0xFF) r = __fadd_rn(r, X.entry[(z * H.n_actions + a) * H.n_chars + k]);
        }
      }
      summed = first ? r : __fadd_rn(r, summed);  // plot.py:208-211
      first = false;
    }
  }
  stamp(X, P, st, env);  // the render that produces this step's observation
  reward = summed;
  disc = H.act.discount[a];
  flags = (H.act.over[a] ? CX_FLAG_TERMINATED : 0) | (H.act.reward_none[a] ? CX_FLAG_REWARD_NONE : 0);
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
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

__device__ __forceinline__ void setup_ctx(Ctx& X, const GenParams& P, uint8_t* smem, uint8_t* s_chidx) {
  X.H = &P.h;
  X.masks = reinterpret_cast<const uint32_t*>(smem + P.h.off_masks);
  X.backdrop = smem + P.h.off_backdrop;
  X.entry = reinterpret_cast<const float*>(smem + P.h.off_entry);
  X.rc = reinterpret_cast<const uint16_t*>(smem + P.h.off_rc);
  X.chidx = s_chidx;
}

__device__ __forceinline__ void stage_tables(const GenParams& P, uint8_t* smem, uint8_t* s_chidx) {
  const uint4* src = reinterpret_cast<const uint4*>(P.blob);
  uint4* dst = reinterpret_cast<uint4*>(smem);
  for (int i = threadIdx.x; i < P.h.blob_bytes / 16; i += blockDim.x) dst[i] = src[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    uint8_t k = 0xFF;
    for (int j = 0; j < P.h.n_chars; ++j)
      if (P.h.chars[j] == i) k = (uint8_t)j;
    s_chidx[i] = k;
  }
}

__global__ void __launch_bounds__(NT) k_generic_rollout(const __grid_constant__ GenParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint8_t s_chidx[256];
  __shared__ uint16_t s_dyn[G][CX_MAX_DYN];
  __shared__ uint16_t s_prev[G][CX_MAX_DYN];
  __shared__ uint8_t s_reset[G];
  const CxGenHeader& H = P.h;
  const int tid = threadIdx.x;
  stage_tables(P, smem, s_chidx);
  Ctx X;
  setup_ctx(X, P, smem, s_chidx);

  const int64_t env0 = (int64_t)blockIdx.x * G;
  const int nenv = (int)min((int64_t)G, P.n - env0);
  const int cells = H.cells;
  const bool mine = tid < nenv;
  const int64_t env = env0 + tid;
  uint32_t ts = 0;
  float rt = 0.0f;
  if (mine) {
    for (int d = 0; d < H.n_dyn; ++d) s_dyn[tid][d] = P.dyn[(int64_t)d * P.n + env];
    if (H.track) {
      ts = P.tstep[env];
      rt = P.ret[env];
    }
  }
  if (tid < G) s_reset[tid] = 0;
  __syncthreads();

  uint32_t ep_cnt = 0, ep_len = 0;
  double ep_sum = 0.0, ep_sumsq = 0.0;
  float ep_max = -INFINITY, ep_negmin = -INFINITY;

  for (int t = 0; t < P.T; ++t) {
    const int64_t row = (int64_t)t * P.n + env0;
    if (mine) {
      const uint32_t a = P.actions[row + tid];
      float rw, dc;
      uint32_t f;
      if (H.track && (ts & CX_OVER_BIT)) {
        rw = 0.0f;
        dc = 0.0f;
        f = CX_FLAG_ALREADY_OVER | CX_FLAG_REWARD_NONE;
      } else {
        generic_env_step(X, P, env, a, s_dyn[tid], s_prev[tid], rw, f, dc);
      }
      if (H.track && !(f & (CX_FLAG_BAD_ACTION | CX_FLAG_ALREADY_OVER))) {
        const uint32_t steps = ts + 1u;
        rt += rw;
        if (!(f & CX_FLAG_TERMINATED) && H.max_steps > 0 && steps >= (uint32_t)H.max_steps) f |= CX_FLAG_TRUNCATED;
        ts = steps;
        if (f & (CX_FLAG_TERMINATED | CX_FLAG_TRUNCATED)) {
          ep_cnt += 1;
          ep_len += steps;
          ep_sum += (double)rt;
          ep_sumsq += (double)rt * (double)rt;
          ep_max = fmaxf(ep_max, rt);
          ep_negmin = fmaxf(ep_negmin, -rt);
          if (H.auto_reset) {
            s_reset[tid] = 1;  // state is reset after this step's (terminal) board has been composed
            ts = 0;
            rt = 0.0f;
          } else {
            ts |= CX_OVER_BIT;
          }
        }
      }
      P.reward[row + tid] = rw;
      if (P.discount) P.discount[row + tid] = dc;
      P.flags[row + tid] = (uint8_t)f;
    }
    __syncthreads();  // entity state and backdrop stamps of all envs of the tile are visible

    // ---- compose and stream out the boards ----
    uint8_t* dst = P.board + row * cells;
    const int nbytes = nenv * cells;
    if (P.vec) {
      const int nchunks = nbytes / 16;
      for (int k = tid; k < nchunks; k += NT) {
        const int g0 = 16 * k;
        int e = g0 / cells, cell = g0 - e * cells;
        uint32_t in[4] = {0, 0, 0, 0}, out[4] = {0, 0, 0, 0};
        if (H.has_dynbd) {
          const uint4 v = *reinterpret_cast<const uint4*>(P.dynbd + env0 * cells + g0);
          in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          const uint8_t bd = H.has_dynbd ? (uint8_t)(in[b >> 2] >> (8 * (b & 3))) : X.backdrop[cell];
          out[b >> 2] |= (uint32_t)overlay(X, s_dyn[e], cell, bd) << (8 * (b & 3));
          if (++cell == cells) {
            cell = 0;
            ++e;
          }
        }
        __stcs(reinterpret_cast<uint4*>(dst) + k, make_uint4(out[0], out[1], out[2], out[3]));
      }
    } else {
      for (int b = tid; b < nbytes; b += NT) {
        const int e = b / cells, cell = b - e * cells;
        dst[b] = overlay(X, s_dyn[e], cell, backdrop_at(X, P, env0 + e, cell));
      }
    }
    __syncthreads();

    // ---- auto reset: back to the its_showtime state (fresh make_game(), actor_critic.py:146) ----
    bool any_reset = false;
    for (int e = 0; e < nenv; ++e) any_reset |= s_reset[e] != 0;
    if (any_reset) {
      if (H.has_dynbd)
        for (int b = tid; b < nbytes; b += NT) {
          const int e = b / cells;
          if (s_reset[e]) P.dynbd[env0 * cells + b] = X.backdrop[b - e * cells];
        }
      if (mine && s_reset[tid])
        for (int z = 0; z < H.n_ent; ++z)
          if (H.ent[z].dyn_slot != 0xFF) s_dyn[tid][H.ent[z].dyn_slot] = H.ent[z].init_state;
      __syncthreads();
      if (tid < G) s_reset[tid] = 0;
      __syncthreads();
    }
  }

  if (mine) {
    for (int d = 0; d < H.n_dyn; ++d) P.dyn[(int64_t)d * P.n + env] = s_dyn[tid][d];
    if (H.track) {
      P.tstep[env] = (uint16_t)ts;
      P.ret[env] = rt;
    }
  }
  if (H.track && tid < 32) {  // G == 32: all env threads live in warp 0
    const double cnt = warp_sum((double)ep_cnt), len = warp_sum((double)ep_len);
    const double sum = warp_sum(ep_sum), sumsq = warp_sum(ep_sumsq);
    const float mx = warp_max(ep_max), ngmn = warp_max(ep_negmin);
    if (tid == 0) {
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

// Render the current state of every env (first frame after its_showtime / reset).
__global__ void __launch_bounds__(NT) k_generic_render(const __grid_constant__ GenParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint8_t s_chidx[256];
  __shared__ uint16_t s_dyn[G][CX_MAX_DYN];
  const CxGenHeader& H = P.h;
  const int tid = threadIdx.x;
  stage_tables(P, smem, s_chidx);
  Ctx X;
  setup_ctx(X, P, smem, s_chidx);
  const int64_t env0 = (int64_t)blockIdx.x * G;
  const int nenv = (int)min((int64_t)G, P.n - env0);
  if (tid < nenv)
    for (int d = 0; d < H.n_dyn; ++d) s_dyn[tid][d] = P.dyn[(int64_t)d * P.n + env0 + tid];
  __syncthreads();
  const int cells = H.cells;
  for (int b = tid; b < nenv * cells; b += NT) {
    const int e = b / cells, cell = b - e * cells;
    P.board[env0 * cells + b] = overlay(X, s_dyn[e], cell, backdrop_at(X, P, env0 + e, cell));
  }
}

GenParams make_params(const cx_game* g, void* d_state, int64_t n) {
  const CxStateLayout L = cx_layout(g, n);
  uint8_t* base = static_cast<uint8_t*>(d_state);
  GenParams P;
  memset(&P, 0, sizeof(P));
  P.h = g->gh;
  P.blob = g->d_blob;
  P.dyn = reinterpret_cast<uint16_t*>(base + L.off_dyn);
  P.tstep = reinterpret_cast<uint16_t*>(base + L.off_tstep);
  P.ret = reinterpret_cast<float*>(base + L.off_ret);
  P.dynbd = base + L.off_dynbd;
  P.stats = reinterpret_cast<double*>(base + L.off_stats);
  P.n = n;
  return P;
}

}  // namespace

int cx_launch_generic_rollout(const cx_game* g, void* d_state, int64_t n, int32_t T, const uint8_t* d_actions,
                              float* d_reward, float* d_discount, uint8_t* d_flags, uint8_t* d_board,
                              cudaStream_t s) {
  GenParams P = make_params(g, d_state, n);
  P.actions = d_actions;
  P.reward = d_reward;
  P.discount = d_discount;
  P.flags = d_flags;
  P.board = d_board;
  P.T = T;
  P.vec = (n % 16 == 0) && (reinterpret_cast<uintptr_t>(d_board) & 15) == 0;
  const int64_t grid = (n + G - 1) / G;
  if (grid > 0x7fffffff) {
    cx_set_error("cx_rollout: too many environments for one launch");
    return CX_ERR_INVALID_ARG;
  }
  static bool configured = false;
  if (!configured) {
    CX_CUDA_OK(cudaFuncSetAttribute(k_generic_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CX_CUDA_OK(cudaFuncSetAttribute(k_generic_render, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  k_generic_rollout<<<(unsigned)grid, NT, g->gh.blob_bytes, s>>>(P);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}

int cx_launch_generic_render(const cx_game* g, const void* d_state, int64_t n, uint8_t* d_board, cudaStream_t s) {
  GenParams P = make_params(g, const_cast<void*>(d_state), n);
  P.board = d_board;
  const int64_t grid = (n + G - 1) / G;
  static bool configured = false;
  if (!configured) {
    CX_CUDA_OK(cudaFuncSetAttribute(k_generic_render, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  k_generic_render<<<(unsigned)grid, NT, g->gh.blob_bytes, s>>>(P);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}
