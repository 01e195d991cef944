// Row F -- bev_pool_v2 forward, cell-stationary kernel (the product path for ranks whose cells are
// sorted, i.e. everything prepare emits and everything rcb_pool_validate proves sorted).
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48; the zero-fill and the permute copy of
// mmdet3d/ops/bev_pool_v2/bev_pool.py:27,91 are folded in (every cell written once, final layout).
//
// k_fwd_cells: one CTA per kPatchX x kPatchY patch of BEV cells; a WARP pools one cell at a time.
// The patch's cells are ranked by length and the warps take them longest first from a shared counter
// (a static longest-first schedule, RCB_FWD_STATIC, measured 1.8x slower: a warp that runs out of work
// idles while its CTA keeps the SM's slots).  Between the opening barrier (cell bounds + order) and
// the closing one (write-out) the warps never wait for each other: the kernel's critical path is one
// cell (<= 656 points on the R50 grid), not one patch (6 k points).
//
// A warp's cells are cut into chunks of 32 points and run through a three-stage software pipeline,
// lane <-> point: (1) coalesced loads of ranks_feat / ranks_depth of chunk t + 2, (2) the depth
// gather of chunk t + 1, (3) chunk t: the depth bins of one (cell, pixel) pair -- adjacent in
// prepare's (pixel, depth) order, 1.42 per pair on the R50 grid -- are merged in registers (groups
// of <= 4, weights summed in point order) and the surviving (row, weight) entries are compacted
// into the warp's 256 bytes of shared memory.  Then lane <-> 128-bit channel quad: every entry
// costs half a broadcast LDS.128, one IMAD.WIDE, one LDG.128 of the context row through L1 -- the
// 2-D patch makes neighbouring cells, which see the same pixels along a ray, share rows there --
// and two packed FMAs; kRows rows are in flight per warp.  No partial sums, no atomics: the order
// of additions inside a cell is the point order (groups folded first), bit-reproducible.
//
// The patch leaves through a shared-memory transpose as (B, C, Z*Y*X) runs of kPatchX cells per
// channel row, or directly as channels-last rows.  Empty cells get their zeros here.
#include <cstdlib>

#include "common.cuh"

namespace rcb {

#ifndef RCB_FWD_PX
#define RCB_FWD_PX 8
#endif
#ifndef RCB_FWD_PY
#define RCB_FWD_PY 4
#endif
#ifndef RCB_FWD_WARPS
#define RCB_FWD_WARPS 8
#endif
#ifndef RCB_FWD_ROWS
#define RCB_FWD_ROWS 4  // context rows in flight per warp
#endif
#ifndef RCB_FWD_CARVEOUT
#define RCB_FWD_CARVEOUT 29  // preferred shared-memory share of the SM's 228 KB, per cent
#endif
#ifndef RCB_FWD_CTAS
#define RCB_FWD_CTAS 5  // CTAs per SM the register budget is cut for
#endif
constexpr int kPatchX = RCB_FWD_PX;
constexpr int kPatchY = RCB_FWD_PY;
constexpr int kPatchCells = kPatchX * kPatchY;
constexpr int kFwdWarps = RCB_FWD_WARPS;
constexpr int kCellWarps = kPatchCells / 32;           // warps that do the bookkeeping, lane <-> cell
constexpr int kFwdRounds = kPatchCells / kFwdWarps;    // cells per warp
constexpr int kTilePitch = kPatchCells + 1;
constexpr int kRowsInFlight = RCB_FWD_ROWS;
static_assert(kPatchCells % 32 == 0 && kPatchCells % kFwdWarps == 0 && kCellWarps <= kFwdWarps, "patch shape");

// CTA shapes.  The kernel's duration is bounded from below by its heaviest patch (the patches next to
// the cameras hold tens of times the mean), which its CTA works off with `warps` warps and `rows`
// context rows in flight per warp.  With many samples in the launch the SMs stay full while those
// patches run and the densest packing wins (shape 0: 40 warps per SM); with few samples the heavy
// patches ARE the kernel and a CTA that throws more warps and more loads at its patch wins, although
// fewer CTAs fit an SM.  Measured forward (us), R50 grid B = 1 / 2 / 4 / 8 and 900x1600 / 256^2 grid
// B = 1 / 2 / 4:   shape 0: 48 / 52 / 60 / 76 and 238 / 249 / 288
//                  shape 1: 29 / 35 / 51 / 97 and 145 / 154 / 288
//                  shape 2: 22 / 38 / 74 / 144 and 116 / 204 / 418
// Every shape pools a cell with one warp in point order: the results are bit-identical.
template <int kShape> struct FwdShape;
template <> struct FwdShape<0> { static constexpr int warps = kFwdWarps, rows = kRowsInFlight, ctas = RCB_FWD_CTAS; };
template <> struct FwdShape<1> { static constexpr int warps = 16, rows = 4, ctas = 2; };   // 64 registers
template <> struct FwdShape<2> { static constexpr int warps = 16, rows = 8, ctas = 1; };   // 124 registers
static_assert(kPatchCells % 16 == 0, "patch shape");
static_assert((kPatchX & (kPatchX - 1)) == 0, "kPatchX is a power of two");

struct FwdCellsParams {
  const float *depth;
  const void *feat;
  const int *ranks_depth, *ranks_feat;
  const int *cell_start;
  float *out;
  int C;
  int X, R;              // cells per row, rows per sample (Z*Y)
  int tiles_x, tiles_r;  // patches per sample
  int cells_per_sample;
  int layout;
  int B;
  unsigned row_bytes;
  FastDiv by_B, by_tiles_x;
  const int *gate;  // launch gate (common.cuh)
};

// i-th element of {c, c-1, c+1, c-2, c+2, ...} clipped to [0, n), c = n / 2
__device__ __forceinline__ int cells_zigzag(int i, int n) {
  const int c = n >> 1;
  const int lo_side = c, hi_side = n - 1 - c;
  const int paired = 2 * min(lo_side, hi_side) + 1;
  if (i < paired) return (i & 1) ? c - ((i + 1) >> 1) : c + (i >> 1);
  const int rest = i - paired;
  return lo_side > hi_side ? c - hi_side - 1 - rest : c + lo_side + 1 + rest;
}

__device__ __forceinline__ void cells_fma(float4 &acc, const float4 v, const float w) {
  const float2 ww = make_float2(w, w);
  const float2 lo = __ffma2_rn(make_float2(v.x, v.y), ww, make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(v.z, v.w), ww, make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

// one chunk (<= 32 consecutive points of one cell) on its way through the pipeline
struct FwdChunk {
  int n;      // valid points, -1: no chunk
  int cell;   // patch-local cell
  int last;   // the cell ends with this chunk
  int rf;     // lane's ranks_feat
  int rd;     // lane's ranks_depth
  float w;    // lane's depth weight
};

// kQ channel quads per lane (quad j of lane l = l + j * lanes); kLanes lanes carry a row
// (C = 4 * kQ * lanes), kLanes == 0: run time.
template <typename FeatT, int kLanes, int kQ, int kShape>
__global__ void __launch_bounds__(32 * FwdShape<kShape>::warps, FwdShape<kShape>::ctas) k_fwd_cells(FwdCellsParams p) {
  pdl_prologue();
  if (gate_closed(p.gate)) return;
  constexpr int kFwdWarps = FwdShape<kShape>::warps;        // (shadow the namespace-scope defaults)
  constexpr int kRowsInFlight = FwdShape<kShape>::rows;
  constexpr int kFwdRounds = kPatchCells / kFwdWarps;
  extern __shared__ __align__(16) float cells_ts[];  // [C][kTilePitch] write-out tile (B_C_CELLS layout only)
  __shared__ __align__(16) uint2 s_ent[kFwdWarps][32];
  __shared__ int s_lo[kPatchCells], s_hi[kPatchCells];
  __shared__ unsigned char s_order[kPatchCells];
#ifndef RCB_FWD_STATIC
  __shared__ int s_next;
  if (threadIdx.x == 0) s_next = 0;
#endif
  constexpr int kRows = kRowsInFlight / kQ;  // per warp, whatever the quads per lane
  const int lanes = kLanes ? kLanes : p.C / (4 * kQ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool to_tile = p.layout == RCB_LAYOUT_B_C_CELLS;
  // Launch order: patches nearest the grid centre first, samples interleaved (point density peaks
  // around the ego vehicle, so the heavy patches start at once).  Any order is correct.
  const int t = (int)p.by_B.div(blockIdx.x);
  const int b = (int)blockIdx.x - t * p.B;
  const int t_r = (int)p.by_tiles_x.div((unsigned)t);
  const int x0 = cells_zigzag(t - t_r * p.tiles_x, p.tiles_x) * kPatchX;
  const int r0 = cells_zigzag(t_r, p.tiles_r) * kPatchY;
  const int nx = min(kPatchX, p.X - x0), nr = min(kPatchY, p.R - r0);
  const int cell_base = b * p.cells_per_sample;

  if (warp < kCellWarps) {
    // lane <-> cell: every bookkeeping warp reads all bounds (its own cells' rank needs all lengths)
    int len[kCellWarps], lo_mine = 0;
#pragma unroll
    for (int m = 0; m < kCellWarps; ++m) {
      const int c = m * 32 + lane, ty = c / kPatchX, tx = c % kPatchX;
      int s = 0, e = 0;
      if (ty < nr && tx < nx) {
        const int g = cell_base + (r0 + ty) * p.X + x0 + tx;
        s = __ldg(p.cell_start + g);
        e = __ldg(p.cell_start + g + 1);
      }
      len[m] = e - s;
      if (m == warp) lo_mine = s;
    }
    int len_mine = 0;
#pragma unroll
    for (int m = 0; m < kCellWarps; ++m)
      if (m == warp) len_mine = len[m];
    const int mine = warp * 32 + lane;
    int rank = 0;  // position in the order: length descending, ties by cell index
#pragma unroll
    for (int m = 0; m < kCellWarps; ++m) {
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int o = __shfl_sync(kFull, len[m], k);
        rank += (o > len_mine || (o == len_mine && m * 32 + k < mine)) ? 1 : 0;
      }
    }
    s_lo[mine] = lo_mine, s_hi[mine] = lo_mine + len_mine;
    s_order[rank] = (unsigned char)mine;
  }
  if (to_tile) {  // empty cells are never visited: their zeros are written here
    float4 *t4 = reinterpret_cast<float4 *>(cells_ts);
    for (int i = tid; i < (p.C * kTilePitch + 3) / 4; i += 32 * kFwdWarps) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const bool active = lane < lanes;
  const char *feat_q = static_cast<const char *>(p.feat) + (size_t)(active ? lane : 0) * 4 * sizeof(FeatT);
  asm volatile("" : "+l"(feat_q));  // the per-lane base stays in a register pair: one IMAD.WIDE per row
  const unsigned qstep = (unsigned)lanes * 4 * sizeof(FeatT);
  const unsigned row_bytes = p.row_bytes;
  const unsigned lt = lanemask_lt();
  uint2 *my_ent = s_ent[warp];

  // ---- this warp's chunks, in order: cells of the snake schedule, 32 points at a time -----------
  int it_round = 0, it_cell = 0, it_base = 0, it_end = 0;
  auto fetch = [&]() -> FwdChunk {  // stage 1: next chunk + its index loads
    FwdChunk c;
    c.n = -1, c.cell = 0, c.last = 0, c.rf = -1 - lane, c.rd = 0, c.w = 0.f;
    if (it_base >= it_end) {  // next non-empty cell (the order ends with the empty ones)
      if (it_round >= kFwdRounds) return c;
#ifndef RCB_FWD_STATIC
      int item = 0;
      if (lane == 0)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(item) : "r"((unsigned)__cvta_generic_to_shared(&s_next)) : "memory");
      item = __shfl_sync(kFull, item, 0);
      if (item >= kPatchCells) {
        it_round = kFwdRounds;
        return c;
      }
#else
      const int item = it_round * kFwdWarps + ((it_round & 1) ? kFwdWarps - 1 - warp : warp);
      ++it_round;
#endif
      it_cell = s_order[item];
      it_base = s_lo[it_cell], it_end = s_hi[it_cell];
      if (it_base >= it_end) {
        it_round = kFwdRounds;
        return c;
      }
    }
    c.cell = it_cell;
    c.n = min(32, it_end - it_base);
    if (lane < c.n) {
      c.rf = ld_stream_s32(p.ranks_feat + it_base + lane);
      c.rd = ld_stream_s32(p.ranks_depth + it_base + lane);
    }
    it_base += 32;
    c.last = it_base >= it_end;
#ifdef RCB_FWD_NOPIPE
    if (lane < c.n) c.w = ld_stream_f32(p.depth + c.rd);
#endif
    return c;
  };
  auto gather = [&](FwdChunk &c) {  // stage 2: depth weights
#ifndef RCB_FWD_NOPIPE
    if (lane < c.n) c.w = ld_stream_f32(p.depth + c.rd);
#endif
  };

  float4 acc[kQ];
#pragma unroll
  for (int j = 0; j < kQ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

#ifdef RCB_FWD_NOPIPE
  FwdChunk c1 = fetch();
  while (c1.n >= 0) {
    const FwdChunk c0 = c1;
#else
  FwdChunk c2 = fetch();
  FwdChunk c1 = c2;
  gather(c1);
  c2 = fetch();
  while (c1.n >= 0) {
    const FwdChunk c0 = c1;
    c1 = c2;
    gather(c1);
    c2 = fetch();
#endif

    // ---- stage 3, lane <-> point: merge the points of one row, compact the entries --------------
    const int rf = c0.rf;
    const int rf_prev = __shfl_up_sync(kFull, rf, 1);
    const unsigned starts = __ballot_sync(kFull, lane == 0 || rf != rf_prev);
    const int run_head = 31 - __clz((int)(starts & (0xffffffffu >> (31 - lane))));
    const int pos = (lane - run_head) & 3;  // position inside a group of <= 4 points of one row
    // followers of a group head: the run goes on to the next start (or the end of the warp)
    const unsigned later = starts & ~(0xffffffffu >> (31 - lane));
    const int run_end = later ? __ffs(later) - 1 : 32;
    const int followers = min(3, run_end - lane - 1);
    float w_sum = c0.w;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float wk = __shfl_down_sync(kFull, c0.w, k);
      if (k <= followers) w_sum += wk;  // (only heads use the sum: pos == 0)
    }
    const bool is_head = lane < c0.n && pos == 0;
    const unsigned heads = __ballot_sync(kFull, is_head);
    const int n_ent = __popc(heads);
    if (is_head) my_ent[__popc(heads & lt)] = make_uint2((unsigned)rf, __float_as_uint(w_sum));
    __syncwarp();

    // ---- lane <-> channel quad: rows through L1, kRows in flight ----------------------------------
    if (active) {
      int j = 0;
      for (; j + kRows <= n_ent; j += kRows) {
        uint2 e1[kRows];
        float4 v[kRows][kQ];
#pragma unroll
        for (int u = 0; u < kRows; u += 2) {
          const uint4 e2 = *reinterpret_cast<const uint4 *>(my_ent + j + u);
          e1[u] = make_uint2(e2.x, e2.y), e1[u + 1] = make_uint2(e2.z, e2.w);
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const char *row = feat_q + (size_t)e1[u].x * row_bytes;
#pragma unroll
          for (int q = 0; q < kQ; ++q) v[u][q] = Row4<FeatT>::load_bytes(row + q * qstep);
        }
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
#pragma unroll
          for (int q = 0; q < kQ; ++q) cells_fma(acc[q], v[u][q], __uint_as_float(e1[u].y));
        }
      }
      if (j < n_ent) {  // the last, partial group: same shape, loads predicated (warp-uniform)
        uint2 e1[kRows];
        float4 v[kRows][kQ];
#pragma unroll
        for (int u = 0; u < kRows - 1; ++u) {
          if (j + u < n_ent) {
            e1[u] = my_ent[j + u];
            const char *row = feat_q + (size_t)e1[u].x * row_bytes;
#pragma unroll
            for (int q = 0; q < kQ; ++q) v[u][q] = Row4<FeatT>::load_bytes(row + q * qstep);
          }
        }
#pragma unroll
        for (int u = 0; u < kRows - 1; ++u) {
          if (j + u < n_ent) {
#pragma unroll
            for (int q = 0; q < kQ; ++q) cells_fma(acc[q], v[u][q], __uint_as_float(e1[u].y));
          }
        }
      }
      if (c0.last) {  // the cell's row: straight out (channels last) or into the transpose tile
        if (to_tile) {
#pragma unroll
          for (int q = 0; q < kQ; ++q) {
            // tile row of channel 4 * quad + k is k * (C / 4) + quad: the lanes of one store hit
            // consecutive rows, i.e. distinct banks (rows of 4 * quad + k put lanes 8 apart on one bank)
            float *col = cells_ts + (lane + q * lanes) * kTilePitch + c0.cell;
            const int kstep = (p.C >> 2) * kTilePitch;
            col[0] = acc[q].x, col[kstep] = acc[q].y, col[2 * kstep] = acc[q].z, col[3 * kstep] = acc[q].w;
          }
        } else {
          const int cty = c0.cell / kPatchX, ctx = c0.cell % kPatchX;
          float4 *dst = reinterpret_cast<float4 *>(p.out + (size_t)(cell_base + (r0 + cty) * p.X + x0 + ctx) * p.C);
#pragma unroll
          for (int q = 0; q < kQ; ++q) st_stream_f4(dst + lane + q * lanes, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < kQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncwarp();  // the next chunk rewrites the entries
#ifdef RCB_FWD_NOPIPE
    c1 = fetch();
#endif
  }

  if (!to_tile) {  // channels last: the empty cells' zeros, cell <-> warp round-robin
    if (active) {
      for (int c = warp; c < kPatchCells; c += kFwdWarps) {
        const int cty = c / kPatchX, ctx = c % kPatchX;
        if (s_hi[c] > s_lo[c] || cty >= nr || ctx >= nx) continue;
        float4 *dst = reinterpret_cast<float4 *>(p.out + (size_t)(cell_base + (r0 + cty) * p.X + x0 + ctx) * p.C);
#pragma unroll
        for (int q = 0; q < kQ; ++q) st_stream_f4(dst + lane + q * lanes, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    return;
  }

  // (B, C, cells): every store instruction writes 32 / kPatchX runs of kPatchX consecutive cells of
  // one channel row
  __syncthreads();
  float *dst_b = p.out + (size_t)b * p.C * p.cells_per_sample;
#pragma unroll
  for (int m = 0; m < kCellWarps; ++m) {
    const int c = m * 32 + lane, ty = c / kPatchX, tx = c % kPatchX;
    if (ty >= nr || tx >= nx) continue;
    float *dst = dst_b + (size_t)(r0 + ty) * p.X + x0 + tx;
    const float *src = cells_ts + c;
    const int C4 = p.C >> 2;
    for (int ch = warp; ch < p.C; ch += kFwdWarps)  // channel ch sits in tile row (ch % 4) * C / 4 + ch / 4
      st_stream_f32(dst + (size_t)ch * p.cells_per_sample, src[((ch & 3) * C4 + (ch >> 2)) * kTilePitch]);
  }
}

template <typename FeatT, int kLanes, int kQ, int kShape>
static int launch_cells_s(const FwdCellsParams &p, long long grid, cudaStream_t s) {
  const size_t smem = p.layout == RCB_LAYOUT_B_C_CELLS ? align_up((size_t)p.C * kTilePitch * 4, 16) : 0;
  if (smem > 40 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(k_fwd_cells<FeatT, kLanes, kQ, kShape>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // the kernel lives on L1-resident context rows: prefer a small shared-memory partition (measured
  // on the R50 workload: 29 % -> L1 hit rate 45 -> 49 %, 78.4 -> 76.7 us; the driver rounds up to
  // what the resident CTAs need)
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_fwd_cells<FeatT, kLanes, kQ, kShape>, cudaFuncAttributePreferredSharedMemoryCarveout, RCB_FWD_CARVEOUT));
  RCB_CUDA_TRY(launch_pdl(k_fwd_cells<FeatT, kLanes, kQ, kShape>, (unsigned)grid, 32 * FwdShape<kShape>::warps, smem, s, p));
  return RCB_OK;
}

// few samples: the heavy patches bound the kernel (see FwdShape); RCB_FWD_SHAPE=0|1|2 overrides
static int fwd_shape_for(int B) {
  static const int forced = [] {
    const char *e = getenv("RCB_FWD_SHAPE");
    return e && e[0] >= '0' && e[0] <= '2' ? e[0] - '0' : -1;
  }();
  if (forced >= 0) return forced;
  return B <= 1 ? 2 : (B <= 4 ? 1 : 0);
}

template <typename FeatT, int kLanes, int kQ>
static int launch_cells_t(const FwdCellsParams &p, long long grid, cudaStream_t s) {
  switch (fwd_shape_for(p.B)) {
    case 2: return launch_cells_s<FeatT, kLanes, kQ, 2>(p, grid, s);
    case 1: return launch_cells_s<FeatT, kLanes, kQ, 1>(p, grid, s);
    default: return launch_cells_s<FeatT, kLanes, kQ, 0>(p, grid, s);
  }
}

template <typename FeatT>
static int launch_cells(const FwdCellsParams &p, long long grid, cudaStream_t s) {
  const int C = p.C;
  if (C <= 128) {  // one quad per lane
    switch (C / 4) {
      case 16: return launch_cells_t<FeatT, 16, 1>(p, grid, s);
      case 20: return launch_cells_t<FeatT, 20, 1>(p, grid, s);
      case 32: return launch_cells_t<FeatT, 32, 1>(p, grid, s);
      default: return launch_cells_t<FeatT, 0, 1>(p, grid, s);
    }
  }
  return launch_cells_t<FeatT, 0, 2>(p, grid, s);  // two quads per lane
}

// Sorted cells with a CSR, whole 128-bit quads per lane, 32-bit row offsets.
bool fwd_cells_eligible(const rcb_pool_desc *d, const void *feat, const int *cell_start) {
  if (!cell_start || !(d->flags & RCB_PLAN_SORTED_CELLS)) return false;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if ((d->C % 4) != 0 || d->C <= 0 || d->C > 256 || (d->C > 128 && (d->C % 8) != 0)) return false;
  if ((long long)d->n_pixels * d->C * elem >= 0xffffffffll) return false;
  if ((((uintptr_t)feat) % (4 * elem)) != 0) return false;
  return true;
}

int fwd_cells_launch(const rcb_pool_desc *d, const float *depth, const void *feat, const int *ranks_depth,
                     const int *ranks_feat, const int *cell_start, float *out, cudaStream_t s) {
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  FwdCellsParams p;
  p.depth = depth, p.feat = feat, p.ranks_depth = ranks_depth, p.ranks_feat = ranks_feat;
  p.cell_start = cell_start, p.out = out;
  p.C = d->C, p.X = d->X, p.R = d->Z * d->Y;
  p.tiles_x = ceil_div(p.X, kPatchX), p.tiles_r = ceil_div(p.R, kPatchY);
  p.cells_per_sample = d->Z * d->Y * d->X, p.layout = d->layout, p.B = d->B;
  p.row_bytes = (unsigned)d->C * elem;
  p.gate = launch_gate();
  p.by_B = FastDiv::make((unsigned)p.B), p.by_tiles_x = FastDiv::make((unsigned)p.tiles_x);
  const long long grid = (long long)d->B * p.tiles_r * p.tiles_x;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: return launch_cells<float>(p, grid, s);
    case RCB_DTYPE_BF16: return launch_cells<__nv_bfloat16>(p, grid, s);
    default: return launch_cells<__half>(p, grid, s);
  }
}

}  // namespace rcb
