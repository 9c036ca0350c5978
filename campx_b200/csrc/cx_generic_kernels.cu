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
//                              backdrop itself (per-env backdrop plane), and a game without any drape
//                              has its canvas zeroed at every render
//   plot directives            campx/plot.py:161-257, engine.py:285-290
//
// B200 mapping.  One WARP owns `tile_envs` (8..32) consecutive envs for all T fused steps,
// and keeps two byte tiles in shared memory: the per-env BACKDROP PLANE (static scenery + quirk-Q1
// sprite stamps; loaded once per launch, written back once) and the composed BOARD.  Warps never wait
// on a block barrier after the tables are staged.  A step is
//   phase 1  lane = env: entity state update in update order (a few bytes in shared memory), stamps
//            into the plane tile;
//   phase 2a whole warp: board tile = plane tile (128-bit shared-memory copies);
//   phase 2b one lane per (env, board row): each entity's row as a 64-bit bitset (static rows from a
//            table, rolled rows by a 64-bit rotate, one-cell entities by a shift), z-order resolved with
//            bit masks front to back, visible bits written as bytes into the board tile;
//   phase 2c whole warp: stream the board tile to HBM with LDS.128 -> STG.128 (st.global.cs).
// Per env-step HBM traffic is the observation contract only (board + reward + flags + discount + action);
// entity state and the plane move once per launch.  Boards wider than 64 columns use the per-cell
// composition fallback.
#include "cx_internal.cuh"
#include "cx_philox.cuh"

namespace {

constexpr int GMAX = CX_GEN_TILE_ENVS;
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
  int32_t synth;                  // actions generated in the kernel (cx_philox.cuh)
  uint64_t seed, env_offset, t0;
  uint8_t* actions_out;
};

struct Ctx {
  const CxGenHeader* H;
  const uint32_t* masks;
  const uint8_t* backdrop;
  const float* entry;
  const uint16_t* rc;
  const uint64_t* rowbits;
  const uint8_t* chidx;  // [256] char code -> game char index (0xFF: not a game char)
};

__device__ __forceinline__ bool mask_bit(const Ctx& X, int z, int cell) {
  return (X.masks[z * X.H->mask_words + (cell >> 5)] >> (cell & 31)) & 1u;
}

// Painter's algorithm at ONE cell on top of backdrop byte v (engine.py:310-321).  Used for the few
// point queries of the last render (wall gate, entry rewards) and by the wide-board fallback.
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

// One board row of entity z as a bitset (bit c = column c), cols <= 64.
__device__ __forceinline__ uint64_t entity_row(const Ctx& X, const uint16_t* st, int z, int r) {
  const CxGenHeader& H = *X.H;
  const CxGenEntity& e = H.ent[z];
  if (e.kind == CX_KIND_STATIC) return X.rowbits[z * H.rows + r];
  const uint32_t s = st[e.dyn_slot];
  if (e.kind == CX_KIND_ROLL) {  // np.roll: content moves down/right by the accumulated offset
    int sr = r - (int)(s >> 8);
    if (sr < 0) sr += H.rows;
    const uint64_t x = X.rowbits[z * H.rows + sr];
    const uint32_t k = s & 255, C = H.cols;
    if (k == 0) return x;
    const uint64_t full = C == 64 ? ~0ull : ((1ull << C) - 1ull);
    return ((x << k) | (x >> (C - k))) & full;
  }
  if (!e.visible || s == CX_EMPTY_CELL16) return 0ull;
  const uint32_t rcv = X.rc[s];
  return (rcv >> 8) == (uint32_t)r ? 1ull << (rcv & 255) : 0ull;
}

// rendering.py:150 through the alias of :128 -- sprites behind the first drape paint into the backdrop
__device__ __forceinline__ void stamp(const Ctx& X, const uint16_t* st, uint8_t* plane) {
  const CxGenHeader& H = *X.H;
  if (!H.has_dynbd) return;
  for (int z = 0; z < H.n_ent; ++z) {
    const CxGenEntity& e = H.ent[z];
    if (e.stamps) plane[st[e.dyn_slot]] = e.ch;
  }
}

__device__ __forceinline__ uint32_t wrap_move(const CxGenHeader& H, uint32_t r, uint32_t c, int dr, int dc,
                                              uint32_t* rr, uint32_t* cc) {
  int nr = (int)r + dr, nc = (int)c + dc;  // |dr| < rows, |dc| < cols (checked by cx_game_create)
  if (nr >= H.rows) nr -= H.rows;
  if (nc >= H.cols) nc -= H.cols;
  if (nr < 0) nr += H.rows;
  if (nc < 0) nc += H.cols;
  *rr = nr;
  *cc = nc;
  return nr * H.cols + nc;
}

// One env, one Engine.play(): entity updates in schedule order with a render after every update group.
// st: current state (updated in place); prev: scratch snapshot of the state at the last render;
// plane: this env's backdrop plane (shared memory).
__device__ void generic_env_step(const Ctx& X, uint32_t a, uint16_t* st, uint16_t* prev, uint8_t* plane,
                                 float& reward, uint32_t& flags, float& disc) {
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
      stamp(X, st, plane);
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
          const uint8_t k = X.chidx[overlay(X, prev, (int)t, plane[t])];
          if (k != 0xFF && ((e.blockers >> k) & 1u)) {
            // fall back to the agent layer of the last render (boat_race.py:55-56)
            const uint32_t pp = prev[e.dyn_slot];
            const bool vis = pp != CX_EMPTY_CELL16 && overlay(X, prev, (int)pp, plane[pp]) == e.ch;
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
          const uint8_t k = X.chidx[overlay(X, prev, (int)wc, plane[wc])];
          if (k != 0xFF) r = __fadd_rn(r, X.entry[(z * H.n_actions + a) * H.n_chars + k]);
        }
      }
      summed = first ? r : __fadd_rn(r, summed);  // plot.py:208-211
      first = false;
    }
  }
  stamp(X, st, plane);  // the render that produces this step's observation
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
  X.rowbits = P.h.off_rowbits >= 0 ? reinterpret_cast<const uint64_t*>(smem + P.h.off_rowbits) : nullptr;
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

// Compose the boards of one warp's envs from `plane` into `out` (both shared memory, [G][cells]).
__device__ __forceinline__ void compose_warp(const Ctx& X, const uint16_t (*dyn)[CX_MAX_DYN], const uint8_t* plane,
                                             uint8_t* out, int nenv, int tile_bytes16, int lane) {
  const CxGenHeader& H = *X.H;
  const int cells = H.cells;
  {  // 2a: board = backdrop plane
    const uint4* s16 = reinterpret_cast<const uint4*>(plane);
    uint4* d16 = reinterpret_cast<uint4*>(out);
    for (int k = lane; k < tile_bytes16; k += 32) d16[k] = s16[k];
  }
  __syncwarp();
  if (X.rowbits) {  // 2b: one lane per (env, row); z-order resolved on 64-bit row bitsets, front to back
    const int R = H.rows, C = H.cols;
    const uint32_t inv_r = 0xFFFFFFFFu / (uint32_t)R + 1u;  // exact for task < 2^16
    for (int task = lane; task < nenv * R; task += 32) {
      const int e = (int)__umulhi((uint32_t)task, inv_r), r = task - e * R;
      uint8_t* row = out + e * cells + r * C;
      uint64_t covered = 0ull;
      for (int z = H.n_ent - 1; z >= 0; --z) {
        const uint64_t bits = entity_row(X, dyn[e], z, r);
        uint64_t vis = bits & ~covered;
        covered |= bits;
        const uint8_t ch = H.ent[z].ch;
        uint32_t lo = (uint32_t)vis, hi = (uint32_t)(vis >> 32);
        while (lo) {
          const int c = __ffs(lo) - 1;
          lo &= lo - 1;
          row[c] = ch;
        }
        while (hi) {
          const int c = __ffs(hi) - 1;
          hi &= hi - 1;
          row[32 + c] = ch;
        }
      }
    }
  } else {  // wide boards: per-cell painter's algorithm, whole warp over the tile
    for (int b = lane; b < nenv * cells; b += 32) {
      const int e = b / cells, cell = b - e * cells;
      out[b] = overlay(X, dyn[e], cell, plane[b]);
    }
  }
  __syncwarp();
}

struct WarpTile {
  uint16_t dyn[GMAX][CX_MAX_DYN];
  uint16_t prev[GMAX][CX_MAX_DYN];
};

__global__ void __launch_bounds__(NT) k_generic_rollout(const __grid_constant__ GenParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint8_t s_chidx[256];
  __shared__ WarpTile s_w[NT / 32];
  const CxGenHeader& H = P.h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpc = blockDim.x >> 5;
  const int G = H.tile_envs, cells = H.cells;
  stage_tables(P, smem, s_chidx);
  Ctx X;
  setup_ctx(X, P, smem, s_chidx);
  __syncthreads();  // tables staged; warps are independent from here on

  const int64_t env0 = ((int64_t)blockIdx.x * wpc + warp) * G;
  if (env0 >= P.n) return;
  const int nenv = (int)min((int64_t)G, P.n - env0);
  const int tile_bytes16 = (G * cells + 15) / 16;
  uint8_t* plane = smem + H.blob_bytes + (size_t)warp * 2 * tile_bytes16 * 16;
  uint8_t* out = plane + tile_bytes16 * 16;
  uint16_t (*dyn)[CX_MAX_DYN] = s_w[warp].dyn;
  const bool mine = lane < nenv;
  const int64_t env = env0 + lane;
  uint32_t ts = 0;
  float rt = 0.0f;
  if (mine) {
    for (int d = 0; d < H.n_dyn; ++d) dyn[lane][d] = P.dyn[(int64_t)d * P.n + env];
    if (H.track) {
      ts = P.tstep[env];
      rt = P.ret[env];
    }
  }
  // backdrop plane of the tile: per-env plane from HBM (quirk Q1 games) or the static scenery
  for (int b = lane; b < nenv * cells; b += 32) {
    const int e = b / cells;
    plane[b] = H.has_dynbd ? P.dynbd[env0 * cells + b] : X.backdrop[b - e * cells];
  }
  __syncwarp();

  uint32_t ep_cnt = 0, ep_len = 0;
  double ep_sum = 0.0, ep_sumsq = 0.0;
  float ep_max = -INFINITY, ep_negmin = -INFINITY;

  for (int t = 0; t < P.T; ++t) {
    const int64_t row = (int64_t)t * P.n + env0;
    bool reset_me = false;
    if (mine) {
      uint32_t a;
      if (P.synth) {
        a = cx_synth_action(P.seed, P.env_offset + (uint64_t)env, P.t0 + (uint64_t)t, (uint32_t)H.n_actions);
        if (P.actions_out) P.actions_out[row + lane] = (uint8_t)a;
      } else {
        a = P.actions[row + lane];
      }
      float rw, dc;
      uint32_t f;
      if (H.track && (ts & CX_OVER_BIT)) {
        rw = 0.0f;
        dc = 0.0f;
        f = CX_FLAG_ALREADY_OVER | CX_FLAG_REWARD_NONE;
      } else {
        generic_env_step(X, a, dyn[lane], s_w[warp].prev[lane], plane + lane * cells, rw, f, dc);
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
            reset_me = true;  // state is reset after this step's (terminal) board has been composed
            ts = 0;
            rt = 0.0f;
          } else {
            ts |= CX_OVER_BIT;
          }
        }
      }
      P.reward[row + lane] = rw;
      if (P.discount) P.discount[row + lane] = dc;
      P.flags[row + lane] = (uint8_t)f;
    }
    __syncwarp();  // entity state and backdrop stamps of the warp's envs are visible

    compose_warp(X, dyn, plane, out, nenv, tile_bytes16, lane);

    // ---- 2c: stream the boards out ----
    uint8_t* dst = P.board + row * cells;
    const int nbytes = nenv * cells;
    if (P.vec) {
      const uint4* s16 = reinterpret_cast<const uint4*>(out);
      uint4* d16 = reinterpret_cast<uint4*>(dst);
#pragma unroll 4
      for (int k = lane; k < nbytes / 16; k += 32) __stcs(d16 + k, s16[k]);
    } else {
      for (int b = lane; b < nbytes; b += 32) dst[b] = out[b];
    }

    // ---- auto reset: back to the its_showtime state (fresh make_game(), actor_critic.py:146) ----
    uint32_t rmask = __ballot_sync(0xffffffffu, reset_me);
    if (rmask) {
      if (reset_me)
        for (int z = 0; z < H.n_ent; ++z)
          if (H.ent[z].dyn_slot != 0xFF) dyn[lane][H.ent[z].dyn_slot] = H.ent[z].init_state;
      while (rmask) {  // the plane of each finished env goes back to the its_showtime backdrop
        const int e = __ffs(rmask) - 1;
        rmask &= rmask - 1;
        uint8_t* pl = plane + e * cells;
        for (int b = lane; b < cells; b += 32) pl[b] = X.backdrop[b];
      }
    }
    __syncwarp();
  }

  if (mine) {
    for (int d = 0; d < H.n_dyn; ++d) P.dyn[(int64_t)d * P.n + env] = dyn[lane][d];
    if (H.track) {
      P.tstep[env] = (uint16_t)ts;
      P.ret[env] = rt;
    }
  }
  if (H.has_dynbd)
    for (int b = lane; b < nenv * cells; b += 32) P.dynbd[env0 * cells + b] = plane[b];
  if (H.track) {
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

// Render the current state of every env (first frame after its_showtime / reset).
__global__ void __launch_bounds__(NT) k_generic_render(const __grid_constant__ GenParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint8_t s_chidx[256];
  __shared__ WarpTile s_w[NT / 32];
  const CxGenHeader& H = P.h;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wpc = blockDim.x >> 5;
  const int G = H.tile_envs, cells = H.cells;
  stage_tables(P, smem, s_chidx);
  Ctx X;
  setup_ctx(X, P, smem, s_chidx);
  __syncthreads();
  const int64_t env0 = ((int64_t)blockIdx.x * wpc + warp) * G;
  if (env0 >= P.n) return;
  const int nenv = (int)min((int64_t)G, P.n - env0);
  const int tile_bytes16 = (G * cells + 15) / 16;
  uint8_t* plane = smem + H.blob_bytes + (size_t)warp * 2 * tile_bytes16 * 16;
  uint8_t* out = plane + tile_bytes16 * 16;
  if (lane < nenv)
    for (int d = 0; d < H.n_dyn; ++d) s_w[warp].dyn[lane][d] = P.dyn[(int64_t)d * P.n + env0 + lane];
  for (int b = lane; b < nenv * cells; b += 32) {
    const int e = b / cells;
    plane[b] = H.has_dynbd ? P.dynbd[env0 * cells + b] : X.backdrop[b - e * cells];
  }
  __syncwarp();
  compose_warp(X, s_w[warp].dyn, plane, out, nenv, tile_bytes16, lane);
  for (int b = lane; b < nenv * cells; b += 32) P.board[env0 * cells + b] = out[b];
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

size_t gen_tile_bytes(const cx_game* g) {
  return ((size_t)g->gh.tile_envs * g->gh.cells + 15) / 16 * 16;
}
// warps per CTA: share the staged tables between warps while keeping a CTA below ~100 KB of shared memory
int gen_warps_per_cta(const cx_game* g) {
  int w = NT / 32;
  while (w > 1 && (size_t)g->gh.blob_bytes + (size_t)w * 2 * gen_tile_bytes(g) > 100 * 1024) w >>= 1;
  return w;
}
size_t gen_smem_bytes(const cx_game* g, int wpc) {
  return (size_t)g->gh.blob_bytes + (size_t)wpc * 2 * gen_tile_bytes(g);
}

int configure_once() {
  static bool configured = false;
  if (!configured) {
    CX_CUDA_OK(cudaFuncSetAttribute(k_generic_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CX_CUDA_OK(cudaFuncSetAttribute(k_generic_render, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  return CX_OK;
}

}  // namespace

int cx_launch_generic_rollout(const cx_game* g, void* d_state, int64_t n, int32_t T, const uint8_t* d_actions,
                              const CxSynth& synth, float* d_reward, float* d_discount, uint8_t* d_flags,
                              uint8_t* d_board, cudaStream_t s) {
  GenParams P = make_params(g, d_state, n);
  P.actions = d_actions;
  P.synth = synth.on;
  P.seed = synth.seed;
  P.env_offset = synth.env_offset;
  P.t0 = synth.t0;
  P.actions_out = synth.actions_out;
  P.reward = d_reward;
  P.discount = d_discount;
  P.flags = d_flags;
  P.board = d_board;
  P.T = T;
  const int G = g->gh.tile_envs;
  // vector stores need every tile and every [T, n] row to start 16-byte aligned and hold whole chunks
  P.vec = ((int64_t)G * g->gh.cells % 16 == 0) && (n % G == 0) && ((n * g->gh.cells) % 16 == 0) &&
          (reinterpret_cast<uintptr_t>(d_board) & 15) == 0;
  const int wpc = gen_warps_per_cta(g);
  const int64_t warps = (n + G - 1) / G;
  const int64_t grid = (warps + wpc - 1) / wpc;
  if (grid > 0x7fffffff) {
    cx_set_error("cx_rollout: too many environments for one launch");
    return CX_ERR_INVALID_ARG;
  }
  int rc = configure_once();
  if (rc) return rc;
  k_generic_rollout<<<(unsigned)grid, wpc * 32, gen_smem_bytes(g, wpc), s>>>(P);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}

int cx_launch_generic_render(const cx_game* g, const void* d_state, int64_t n, uint8_t* d_board, cudaStream_t s) {
  GenParams P = make_params(g, const_cast<void*>(d_state), n);
  P.board = d_board;
  const int G = g->gh.tile_envs;
  const int wpc = gen_warps_per_cta(g);
  const int64_t warps = (n + G - 1) / G;
  const int64_t grid = (warps + wpc - 1) / wpc;
  int rc = configure_once();
  if (rc) return rc;
  k_generic_render<<<(unsigned)grid, wpc * 32, gen_smem_bytes(g, wpc), s>>>(P);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}
