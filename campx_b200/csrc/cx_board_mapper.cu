// Board -> array conversion for a batch of finished boards: the device side of
// ObservationToArray (campx/rendering.py:461-594: every board character is replaced by a scalar or a 1-D
// vector from a value mapping, e.g. RGB rendering) and ObservationToFeatureArray (rendering.py:597-712:
// float 0/1 planes of chosen characters, zero planes for characters the game does not have), including the
// `permute` argument of both (any ordering of the vector / row / column axes).
//
// Both are one table lookup per output element: out[board, i0, i1, i2] = values[board byte at (row, col)][d]
// with (d, row, col) a permutation of (i0, i1, i2).  One CTA stages EB boards and the value table in shared
// memory; every thread then produces 16 consecutive output BYTES per iteration in output order (so stores are
// coalesced 16-byte vectors whatever the permutation) by walking an (i0, i1, i2) cursor.
// HBM traffic = cells bytes read + depth * cells * elem_size bytes written per board.
#include <string.h>

#include <new>

#include "cx_internal.cuh"

struct cx_board_mapper {
  int depth, esz;
  uint8_t* d_values;  // [256][depth] elements of esz bytes
  uint32_t known[8];  // bit b: byte value b has an entry in the value mapping
  int device;
};

namespace {

constexpr int TB = 256;
constexpr int MAX_ROW_BYTES = 128;  // depth * elem_size: the value table is at most 32 KB of shared memory

struct KnownBits {
  uint32_t w[8];
};

struct MapGeom {
  int cells, depth, per;  // per = depth * cells output elements per board
  int n1, n2;             // sizes of output axes 1 and 2 (axis 0 is whatever is left)
  int a0, a1, a2;         // board-cell stride of output axis k (0 for the vector axis)
  int b0, b1, b2;         // vector-component stride of output axis k (1 for the vector axis, else 0)
};

enum { LAYOUT_ANY = 0, LAYOUT_CHW = 1, LAYOUT_HWC = 2 };

template <typename E, int LAYOUT>
__global__ void __launch_bounds__(TB) k_board_map(const uint8_t* __restrict__ board, E* __restrict__ out,
                                                  const E* __restrict__ values, KnownBits known, MapGeom g, int EB,
                                                  int64_t n_boards, int32_t* __restrict__ unknown) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int V = 16 / (int)sizeof(E);  // elements per 16-byte store
  E* s_val = reinterpret_cast<E*>(smem);  // [256][depth]
  uint8_t* s_board = smem + ((256 * g.depth * (int)sizeof(E) + 15) & ~15);
  const int64_t b0 = (int64_t)blockIdx.x * EB;
  const int nb = (int)min((int64_t)EB, n_boards - b0);
  const int tile_bytes = nb * g.cells;
  const uint8_t* src = board + b0 * g.cells;
  for (int i = threadIdx.x; i < 256 * g.depth; i += TB) s_val[i] = values[i];
  bool bad = false;
  auto check = [&](uint32_t b) { bad |= !((known.w[b >> 5] >> (b & 31)) & 1u); };
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* s16 = reinterpret_cast<const uint4*>(src);
    uint4* d16 = reinterpret_cast<uint4*>(s_board);
    for (int i = threadIdx.x; i < tile_bytes / 16; i += TB) {
      const uint4 q = __ldcs(s16 + i);
      d16[i] = q;
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 16; ++j) check((w[j >> 2] >> (8 * (j & 3))) & 255u);
    }
    for (int i = (tile_bytes & ~15) + threadIdx.x; i < tile_bytes; i += TB) {
      s_board[i] = src[i];
      check(src[i]);
    }
  } else {
    for (int i = threadIdx.x; i < tile_bytes; i += TB) {
      s_board[i] = src[i];
      check(src[i]);
    }
  }
  if (bad && unknown) atomicAdd(unknown, 1);
  __syncthreads();
  const int total = nb * g.per;  // output elements of this CTA (< 2^31)
  E* dst = out + b0 * g.per;
  const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  const int last = tile_bytes - 1;  // cursors that run past the CTA's output are clamped; those elements are not stored
  for (int o = threadIdx.x * V; o < total; o += TB * V) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};  // the 16 output bytes, packed in registers
    auto put = [&](int j, E val) {
      if (sizeof(E) == 8) {
        w[(2 * j) & 3] = (uint32_t)((uint64_t)val);
        w[(2 * j + 1) & 3] = (uint32_t)((uint64_t)val >> 32);
      } else {
        w[((j * (int)sizeof(E)) >> 2) & 3] |= (uint32_t)val << (8 * ((j * (int)sizeof(E)) & 3));
      }
    };
    if (LAYOUT == LAYOUT_HWC) {
      // [board, row, col, component]: element o of the CTA is component o % depth of tile byte o / depth
      int B = o / g.depth;
      int d = o - B * g.depth;
      int base = (int)s_board[B] * g.depth;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        put(j, s_val[base + d]);
        if (++d == g.depth) {
          d = 0;
          ++B;
          base = (int)s_board[min(B, last)] * g.depth;
        }
      }
    } else if (LAYOUT == LAYOUT_CHW) {
      // [board, component, row, col]: runs of `cells` consecutive tile bytes, repeated once per component
      const int e = o / g.per;
      const int rem = o - e * g.per;
      int d = rem / g.cells;
      int cell = rem - d * g.cells;
      int B = e * g.cells + cell;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        put(j, s_val[(int)s_board[min(B, last)] * g.depth + d]);
        ++B;
        if (++cell == g.cells) {
          cell = 0;
          if (++d == g.depth)
            d = 0;  // next board: B already points at its first cell
          else
            B -= g.cells;
        }
      }
    } else {
      // any axis order: board e, output coordinates (i0, i1, i2) of element o; divided out once, then advanced
      int e = o / g.per;
      int rem = o - e * g.per;
      const int n12 = g.n1 * g.n2;
      int i0 = rem / n12;
      rem -= i0 * n12;
      int i1 = rem / g.n2;
      int i2 = rem - i1 * g.n2;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int cell = i0 * g.a0 + i1 * g.a1 + i2 * g.a2;
        const int d = i0 * g.b0 + i1 * g.b1 + i2 * g.b2;
        put(j, s_val[(int)s_board[min(e * g.cells + cell, last)] * g.depth + d]);
        if (++i2 == g.n2) {
          i2 = 0;
          if (++i1 == g.n1) {
            i1 = 0;
            if (++i0 * n12 == g.per) {
              i0 = 0;
              ++e;
            }
          }
        }
      }
    }
    if (vec && o + V <= total) {
      __stcs(reinterpret_cast<uint4*>(dst + o), make_uint4(w[0], w[1], w[2], w[3]));
    } else {
      uint8_t* d8 = reinterpret_cast<uint8_t*>(dst + o);
      const int nbytes = min(V, total - o) * (int)sizeof(E);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nbytes) d8[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
    }
  }
}

// boards per CTA: a multiple of 16 (every CTA's tiles start 16-byte aligned), <= 32 KB of boards
inline int boards_per_cta(int cells) {
  int eb = (32 * 1024 / cells) / 16 * 16;
  if (eb > 256) eb = 256;
  return eb < 16 ? 16 : eb;
}

template <typename E, int LAYOUT>
int launch_map(const cx_board_mapper* m, const uint8_t* d_board, int64_t n_boards, const MapGeom& g, void* d_out,
               int32_t* d_unknown, cudaStream_t s) {
  static CxPerDevice configured;
  if (configured.need()) {
    CX_CUDA_OK(cudaFuncSetAttribute(k_board_map<E, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured.mark();
  }
  const int EB = boards_per_cta(g.cells);
  const int64_t grid = (n_boards + EB - 1) / EB;
  if (grid > 0x7fffffff) {
    cx_set_error("cx_board_mapper_apply: too many boards for one launch");
    return CX_ERR_INVALID_ARG;
  }
  KnownBits kb;
  memcpy(kb.w, m->known, sizeof(kb.w));
  const size_t smem = ((size_t)256 * g.depth * sizeof(E) + 15) / 16 * 16 + ((size_t)EB * g.cells + 15) / 16 * 16;
  k_board_map<E, LAYOUT><<<(unsigned)grid, TB, smem, s>>>(d_board, static_cast<E*>(d_out),
                                                  reinterpret_cast<const E*>(m->d_values), kb, g, EB, n_boards,
                                                  d_unknown);
  CX_CUDA_OK(cudaGetLastError());
  return CX_OK;
}

}  // namespace

extern "C" int cx_board_mapper_create(const void* h_values, const uint8_t* h_known, int32_t depth, int32_t elem_size,
                                      cx_board_mapper** out) {
  if (!h_values || !h_known || !out || depth < 1 ||
      (elem_size != 1 && elem_size != 2 && elem_size != 4 && elem_size != 8)) {
    cx_set_error("cx_board_mapper_create: bad argument (elem_size must be 1, 2, 4 or 8)");
    return CX_ERR_INVALID_ARG;
  }
  if (depth * elem_size > MAX_ROW_BYTES) {
    cx_set_error("cx_board_mapper_create: depth * elem_size = %d exceeds %d bytes per character", depth * elem_size,
                 MAX_ROW_BYTES);
    return CX_ERR_UNSUPPORTED;
  }
  cx_board_mapper* m = new (std::nothrow) cx_board_mapper();
  if (!m) {
    cx_set_error("cx_board_mapper_create: out of host memory");
    return CX_ERR_NOMEM;
  }
  m->depth = depth;
  m->esz = elem_size;
  m->d_values = nullptr;
  memset(m->known, 0, sizeof(m->known));
  for (int b = 0; b < 256; ++b)
    if (h_known[b]) m->known[b >> 5] |= 1u << (b & 31);
  const size_t bytes = (size_t)256 * depth * elem_size;
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_values, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_values, h_values, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cx_set_error("cx_board_mapper_create: %s", cudaGetErrorString(e));
    if (m->d_values) cudaFree(m->d_values);
    delete m;
    return e == cudaErrorMemoryAllocation ? CX_ERR_NOMEM : CX_ERR_CUDA;
  }
  *out = m;
  return CX_OK;
}

extern "C" int cx_board_mapper_destroy(cx_board_mapper* m) {
  if (!m) return CX_OK;
  if (m->d_values) cudaFree(m->d_values);
  delete m;
  return CX_OK;
}

extern "C" int cx_board_mapper_apply(const cx_board_mapper* m, const uint8_t* d_board, int64_t n_boards, int32_t rows,
                                     int32_t cols, const int32_t* permute, void* d_out, int32_t* d_unknown,
                                     void* stream) {
  if (!m || !d_board || !d_out || n_boards < 1 || rows < 1 || cols < 1 || (int64_t)rows * cols > CX_MAX_CELLS) {
    cx_set_error("cx_board_mapper_apply: bad argument");
    return CX_ERR_INVALID_ARG;
  }
  int perm[3] = {0, 1, 2};
  if (permute) {
    int seen = 0;
    for (int k = 0; k < 3; ++k) {
      if (permute[k] < 0 || permute[k] > 2) {
        cx_set_error("cx_board_mapper_apply: permute must be a permutation of 0, 1, 2");
        return CX_ERR_INVALID_ARG;
      }
      perm[k] = permute[k];
      seen |= 1 << permute[k];
    }
    if (seen != 7) {
      cx_set_error("cx_board_mapper_apply: permute must be a permutation of 0, 1, 2");
      return CX_ERR_INVALID_ARG;
    }
  }
  // source axes: 0 = vector component, 1 = row, 2 = column (rendering.py:478-489)
  const int size[3] = {m->depth, rows, cols};
  const int cell_stride[3] = {0, cols, 1};
  const int comp_stride[3] = {1, 0, 0};
  MapGeom g;
  g.cells = rows * cols;
  g.depth = m->depth;
  g.per = g.depth * g.cells;
  g.n1 = size[perm[1]];
  g.n2 = size[perm[2]];
  g.a0 = cell_stride[perm[0]];
  g.a1 = cell_stride[perm[1]];
  g.a2 = cell_stride[perm[2]];
  g.b0 = comp_stride[perm[0]];
  g.b1 = comp_stride[perm[1]];
  g.b2 = comp_stride[perm[2]];
  cudaStream_t s = (cudaStream_t)stream;
  // the two layouts users ask for -- (vector, row, col) and channels-last (row, col, vector); a scalar mapping
  // (depth 1) is both whenever rows precede columns -- have their own index arithmetic
  const bool rc = perm[0] == 1 && perm[1] == 2, cr = perm[1] == 1 && perm[2] == 2;
  const int layout = (perm[0] == 0 && cr) ? LAYOUT_CHW
                     : (perm[2] == 0 && rc) ? LAYOUT_HWC
                     : (m->depth == 1 && (rc || cr || (perm[0] == 1 && perm[2] == 2))) ? LAYOUT_CHW : LAYOUT_ANY;
#define CX_MAP_DISPATCH(E)                                                                              \
  (layout == LAYOUT_CHW   ? launch_map<E, LAYOUT_CHW>(m, d_board, n_boards, g, d_out, d_unknown, s)     \
   : layout == LAYOUT_HWC ? launch_map<E, LAYOUT_HWC>(m, d_board, n_boards, g, d_out, d_unknown, s)     \
                          : launch_map<E, LAYOUT_ANY>(m, d_board, n_boards, g, d_out, d_unknown, s))
  switch (m->esz) {
    case 1: return CX_MAP_DISPATCH(uint8_t);
    case 2: return CX_MAP_DISPATCH(uint16_t);
    case 4: return CX_MAP_DISPATCH(uint32_t);
    default: return CX_MAP_DISPATCH(uint64_t);
  }
#undef CX_MAP_DISPATCH
}
