// Rows F and B -- bev_pool_v2 forward / backward, STRIP-stationary kernels.
// Reference semantics: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (forward), :67-121
// (backward); host sequence mmdet3d/ops/bev_pool_v2/bev_pool.py:16-83.
//
// Geometry the design rests on: the pixels of one image column look along (nearly) the same azimuth,
// so at a given depth bin they fall into the same BEV cell (Z is collapsed), and consecutive depth
// bins of the column walk along a line of cells.  A STRIP is a vertical run of <= 16 pixels of one
// image column.  On the R50 grid the 1888 frustum points of a strip touch ~70-110 distinct cells:
// the (cell, strip) SEGMENTS are 9-13x fewer than the (cell, pixel) pairs the cell-stationary kernel
// gathers a 320-byte context row for (19-28 k against 253 k per sample).
//
//   out[cell] = sum over the cell's segments of  ( sum over the segment's entries (d, mask) of
//                                                  sum_v mask_v * depth[d, v] * feat[v, :] )
//
// so the inner two sums are a small DENSE product with the strip's 16 context rows held in
// registers (each row is read from memory exactly once per launch), and only one partial row per
// segment crosses memory:
//
//   plan (once per set of ranks; k_strip_plan + k_cellseg_sort)
//       per strip : ENTRIES (depth bin, 16-bit pixel mask, last-of-segment flag), sorted by
//                   (cell, depth bin); segment j of strip s owns row s * seg_cap + j of a workspace
//       per cell  : the sorted list of its segments' row indices (lives in the cell's own slice of
//                   the point-sorted index space, cell_start[c] + k: no second CSR)
//   forward  k_fwd_strips  : strip-stationary; registers hold feat[8 px][C/16 ch] per lane, the
//                            strip's depth slab sits in shared memory; one partial row per segment
//            k_fwd_combine : cell-stationary; sums a cell's segment rows in list order (fixed ->
//                            bit-reproducible) and writes the final layout, zeros included
//   backward k_bwd_spread  : the mirror of combine: out_grad row of a cell -> its segments' rows
//            k_bwd_strips  : strip-stationary, lane <-> pixel: per segment ONE dot product per pixel
//                            (depth_grad of every bin of the segment is that value) and one
//                            rank-1 update of the pixel's feat_grad with the segment's summed weight
//
// Nothing here uses float atomics; integer atomics only hand out list slots, and the lists are
// sorted afterwards, so every output is bit-reproducible.  Inputs the plan cannot hold (more
// entries / segments per strip than reserved, which needs cells scattered at random along a
// column) raise the plan's status word: every kernel here then exits at once and the
// cell-/pixel-stationary kernels (pool_fwd_cells.cu, pool_bwd.cu), enqueued behind with the same
// word as their gate, do the work instead.
#include "common.cuh"

namespace rcb {

constexpr int kStripV = 16;    // pixels per strip
constexpr int kStripCols = 4;  // adjacent image columns per CTA (16-byte runs of depth / point_cell)
constexpr int kCombineCells = 64;

struct StripGeom {
  int n_img, D, H, W, HW;
  int VC;  // strips per image column
  int UG;  // column groups per image row
  int seg_cap, ent_cap;
  int n_strips;
  int n_cells;
  int n_list;  // capacity of the per-cell lists (= points)
};

static inline int next_pow2(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

static bool make_geom(const rcb_strip_desc *d, StripGeom *g) {
  if (!d || d->n_img <= 0 || d->D <= 0 || d->D > 256 || d->H <= 0 || d->W <= 0 || d->n_cells <= 0 ||
      d->n_cells > (1 << 24))
    return false;
  g->n_img = d->n_img, g->D = d->D, g->H = d->H, g->W = d->W, g->HW = d->H * d->W;
  g->VC = ceil_div(d->H, kStripV), g->UG = ceil_div(d->W, kStripCols);
  g->ent_cap = next_pow2(4 * d->D);
  g->seg_cap = (int)align_up((size_t)3 * d->D, 8);
  const long long strips = (long long)d->n_img * g->VC * d->W;
  const long long points = (long long)d->n_img * d->D * g->HW;
  if (strips * g->seg_cap >= (1ll << 31) || points >= (1ll << 31) || strips * g->ent_cap >= (1ll << 31))
    return false;
  g->n_strips = (int)strips, g->n_cells = d->n_cells, g->n_list = (int)points;
  return true;
}

// plan buffer: header | info | entries | cell_nseg | cellseg_raw | cellseg
struct PlanView {
  int *status;
  int4 *info;
  unsigned *ent;
  int *cell_nseg, *raw, *list;
  size_t bytes;
};

static PlanView plan_view(const StripGeom &g, void *base) {
  PlanView v;
  char *p = static_cast<char *>(base);
  size_t off = 0;
  v.status = reinterpret_cast<int *>(p + off), off += 256;
  v.info = reinterpret_cast<int4 *>(p + off), off += align_up((size_t)g.n_strips * 16, 256);
  v.ent = reinterpret_cast<unsigned *>(p + off), off += align_up((size_t)g.n_strips * g.ent_cap * 4, 256);
  v.cell_nseg = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_cells * 4, 256);
  v.raw = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_list * 4, 256);
  v.list = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_list * 4, 256);
  v.bytes = off;
  return v;
}

__device__ __forceinline__ void strip_block(const StripGeom &g, int bid, int &img, int &vc, int &u0) {
  const int ug = bid % g.UG, t = bid / g.UG;
  vc = t % g.VC, img = t / g.VC, u0 = ug * kStripCols;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct StripPlanParams {
  const int *point_cell, *cell_start;
  int *status;
  int4 *info;
  unsigned *ent;
  int *cell_nseg, *raw;
  StripGeom g;
  int vec4;
};

__global__ void __launch_bounds__(32 * kStripCols) k_strip_plan(StripPlanParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char plan_smem[];
  const StripGeom &g = p.g;
  const int D = g.D;
  int *s_cells = reinterpret_cast<int *>(plan_smem);  // [col][D][16]
  unsigned long long *s_buf = reinterpret_cast<unsigned long long *>(plan_smem + (size_t)kStripCols * D * kStripV * 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int img, vc, u0;
  strip_block(g, blockIdx.x, img, vc, u0);
  const int v0 = vc * kStripV;

  for (int idx = tid; idx < D * kStripV; idx += 32 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, v = v0 + i;
    int c[kStripCols] = {-1, -1, -1, -1};
    if (v < g.H) {
      const int *src = p.point_cell + ((size_t)(img * D + d) * g.HW + (size_t)v * g.W + u0);
      if (p.vec4) {
        const int4 q = *reinterpret_cast<const int4 *>(src);
        c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < kStripCols; ++k)
          if (u0 + k < g.W) c[k] = src[k];
      }
    }
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) s_cells[(k * D + d) * kStripV + i] = c[k];
  }
  __syncthreads();

  const int u = u0 + warp;
  if (u >= g.W) return;
  const int strip = (img * g.VC + vc) * g.W + u;
  const int *cells = s_cells + warp * D * kStripV;
  unsigned long long *buf = s_buf + (size_t)warp * g.ent_cap;

  // entries: per depth bin the distinct cells among the strip's pixels, two bins per round
  int n = 0;
  const int half = lane >> 4, i16 = lane & 15;
  for (int d0 = 0; d0 < D; d0 += 2) {
    const int dd = d0 + half;
    const int cell = dd < D ? cells[dd * kStripV + i16] : -1;
    const unsigned grp = __match_any_sync(kFull, (unsigned)cell ^ ((unsigned)half << 30));
    const bool emit = cell >= 0 && lane == __ffs(grp) - 1;
    const unsigned em = __ballot_sync(kFull, emit);
    if (emit) {
      const int pos = n + __popc(em & lanemask_lt());
      if (pos < g.ent_cap)
        buf[pos] = ((unsigned long long)cell << 24) | ((unsigned long long)dd << 16) |
                   (unsigned long long)((grp >> (lane & 16)) & 0xffffu);
    }
    n += __popc(em);
  }
  if (n > g.ent_cap) {
    if (lane == 0) {
      atomicOr(p.status, 1);
      p.info[strip] = make_int4(0, 0, 0, 0);
    }
    return;
  }

  // sort by (cell, depth bin): bitonic network over the warp's shared-memory slice
  int n2 = 32;
  while (n2 < n) n2 <<= 1;
  for (int i = n + lane; i < n2; i += 32) buf[i] = ~0ull;
  __syncwarp();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n2 >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const unsigned long long a = buf[i], b = buf[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) buf[i] = b, buf[l] = a;
      }
      __syncwarp();
    }
  }

  // segments = runs of equal cell
  int segs = 0, e_half = n, j_half = 0;
  bool found = false;
  const int half_target = n >> 1;
  for (int e0 = 0; e0 < n; e0 += 32) {
    const int e = e0 + lane;
    const bool valid = e < n;
    const unsigned long long cur = valid ? buf[e] : 0ull;
    const int cell = (int)(cur >> 24);
    const bool first = valid && (e == 0 || (int)(buf[e - 1] >> 24) != cell);
    const bool last = valid && (e + 1 >= n || (int)(buf[e + 1] >> 24) != cell);
    const unsigned fm = __ballot_sync(kFull, first);
    const int j = segs + __popc(fm & (lanemask_lt() | (1u << lane))) - 1;
    if (valid)
      p.ent[(size_t)strip * g.ent_cap + e] =
          (unsigned)((cur >> 16) & 0xffu) | ((unsigned)(cur & 0xffffu) << 8) | ((unsigned)last << 24);
    if (first && j < g.seg_cap) {
      const int k = atomicAdd(&p.cell_nseg[cell], 1);
      p.raw[p.cell_start[cell] + k] = strip * g.seg_cap + j;
    }
    const unsigned hm = __ballot_sync(kFull, first && e >= half_target);
    if (hm && !found) {
      const int src = __ffs(hm) - 1;
      e_half = e0 + src, j_half = __shfl_sync(kFull, j, src), found = true;
    }
    segs += __popc(fm);
  }
  if (!found) j_half = segs;
  if (lane == 0) {
    if (segs > g.seg_cap) {
      atomicOr(p.status, 2);
      p.info[strip] = make_int4(0, 0, 0, 0);
    } else {
      p.info[strip] = make_int4(n, segs, e_half, j_half);
    }
  }
}

constexpr int kMaxListSort = 1024;

// per cell: claimed order (atomics) -> ascending row index = (strip, segment) order
__global__ void __launch_bounds__(256) k_cellseg_sort(const int *cell_nseg, const int *cell_start, const int *raw,
                                                      int *list, int n_cells, int *status) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * 256 + threadIdx.x);
  int cnt = 0, cs = 0;
  if (c < n_cells) {
    cnt = cell_nseg[c];
    if (cnt) cs = cell_start[c];
  }
  if (cnt == 1) list[cs] = raw[cs];
  unsigned multi = __ballot_sync(kFull, cnt >= 2);
  while (multi) {
    const int src = __ffs(multi) - 1;
    multi &= multi - 1;
    const int n = __shfl_sync(kFull, cnt, src), b = __shfl_sync(kFull, cs, src);
    if (n <= 32) {
      const int id = lane < n ? raw[b + lane] : 0x7fffffff;
      int rank = 0;
      for (int m = 0; m < n; ++m) rank += __shfl_sync(kFull, id, m) < id;
      if (lane < n) list[b + rank] = id;
    } else if (n <= kMaxListSort) {
      for (int e = lane; e < n; e += 32) {
        const int id = raw[b + e];
        int rank = 0;
        for (int m = 0; m < n; ++m) rank += raw[b + m] < id;
        list[b + rank] = id;
      }
    } else if (lane == 0) {
      atomicOr(status, 4);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct FwdStripsParams {
  const float *depth;
  const void *feat;
  const int *status;
  const int4 *info;
  const unsigned *ent;
  float *rows;
  StripGeom g;
  int vec4;
};

// lane = (h, j): h = lane / 16 owns pixels 8h .. 8h+7 of the strip, j = lane % 16 owns channels
// j + 16k, k < CPL (C = 16 * CPL).
template <typename FeatT, int CPL>
__global__ void __launch_bounds__(64 * kStripCols, 3) k_fwd_strips(FwdStripsParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  constexpr int CP = CPL / 2;
  extern __shared__ __align__(16) unsigned char fs_smem[];
  const StripGeom &g = p.g;
  const int D = g.D;
  float *s_w = reinterpret_cast<float *>(fs_smem);  // [col][D][16]
  unsigned *s_ent = reinterpret_cast<unsigned *>(fs_smem + (size_t)kStripCols * D * kStripV * 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int img, vc, u0;
  strip_block(g, blockIdx.x, img, vc, u0);
  const int v0 = vc * kStripV;

  for (int idx = tid; idx < D * kStripV; idx += 64 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, v = v0 + i;
    float c[kStripCols] = {0.f, 0.f, 0.f, 0.f};
    if (v < g.H) {
      const float *src = p.depth + ((size_t)(img * D + d) * g.HW + (size_t)v * g.W + u0);
      if (p.vec4) {
        const float4 q = ld_stream_f4(reinterpret_cast<const float4 *>(src));
        c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < kStripCols; ++k)
          if (u0 + k < g.W) c[k] = ld_stream_f32(src + k);
      }
    }
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) s_w[(k * D + d) * kStripV + i] = c[k];
  }

  const int col = warp >> 1, part = warp & 1;
  const int u = u0 + col;
  const bool active = u < g.W;
  const int strip = (img * g.VC + vc) * g.W + (active ? u : 0);
  const int4 inf = active ? p.info[strip] : make_int4(0, 0, 0, 0);
  const int e_lo = part ? inf.z : 0, e_hi = part ? inf.x : inf.z;
  int seg = part ? inf.w : 0;
  unsigned *ents = s_ent + col * g.ent_cap;
  for (int e = e_lo + lane; e < e_hi; e += 32) ents[e] = p.ent[(size_t)strip * g.ent_cap + e];

  const int h = lane >> 4, j = lane & 15;
  float f[8][CPL];
  const FeatT *feat = static_cast<const FeatT *>(p.feat);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int v = v0 + 8 * h + t;
    const bool ok = active && v < g.H && e_lo < e_hi;
    const FeatT *row = feat + ((size_t)img * g.HW + (size_t)(ok ? v : 0) * g.W + (ok ? u : 0)) * C + j;
#pragma unroll
    for (int k = 0; k < CPL; ++k) f[t][k] = ok ? to_f32<FeatT>(row[16 * k]) : 0.f;
  }
  __syncthreads();
  if (e_lo >= e_hi) return;

  const float *wcol = s_w + col * D * kStripV + 8 * h;
  float2 acc[CP > 0 ? CP : 1];
  float2 acc_l = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < CP; ++k) acc[k] = make_float2(0.f, 0.f);
  float *dst = p.rows + ((size_t)strip * g.seg_cap + seg) * C + j;
  const int mshift = 8 + 8 * h;

  for (int e = e_lo; e < e_hi; ++e) {
    const unsigned word = ents[e];
    const int d = word & 255;
    const unsigned m8 = (word >> mshift) & 255u;
    const float4 wa = *reinterpret_cast<const float4 *>(wcol + d * kStripV);
    const float4 wb = *reinterpret_cast<const float4 *>(wcol + d * kStripV + 4);
    float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) ww[t] = (m8 >> t) & 1u ? ww[t] : 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float2 w2 = make_float2(ww[t], ww[t]);
#pragma unroll
      for (int k = 0; k < CP; ++k) acc[k] = __ffma2_rn(make_float2(f[t][2 * k], f[t][2 * k + 1]), w2, acc[k]);
    }
    if (CPL & 1) {
#pragma unroll
      for (int t = 0; t < 8; t += 2)
        acc_l = __ffma2_rn(make_float2(f[t][CPL - 1], f[t + 1][CPL - 1]), make_float2(ww[t], ww[t + 1]), acc_l);
    }
    if (word >> 24) {  // last entry of the segment: fold the two pixel halves, write the partial row
      float a[CPL];
#pragma unroll
      for (int k = 0; k < CP; ++k) a[2 * k] = acc[k].x, a[2 * k + 1] = acc[k].y, acc[k] = make_float2(0.f, 0.f);
      if (CPL & 1) a[CPL - 1] = acc_l.x + acc_l.y, acc_l = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CPL; ++k) a[k] += __shfl_xor_sync(kFull, a[k], 16);
#pragma unroll
      for (int k = 0; k < CPL; ++k)
        if ((k & 1) == h) dst[16 * k] = a[k];
      dst += C;
    }
  }
}

struct CombineParams {
  const int *status;
  const int *cell_nseg, *cell_start, *list;
  float *rows;        // forward: read; backward spread: written
  float *out;         // forward: the pooled tensor; backward spread: out_grad (read)
  int cps, tiles_per_sample, layout;
};

// lane = (h, j): half-warp h pools one cell, lane j owns channels j + 16k
template <int CPL>
__global__ void __launch_bounds__(256) k_fwd_combine(CombineParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  constexpr int kPitch = kCombineCells + 1;
  extern __shared__ __align__(16) float cb_tile[];  // [C][kPitch]   (B, C, cells) layout only
  __shared__ int s_cnt[kCombineCells], s_cs[kCombineCells];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.tiles_per_sample, c0 = (blockIdx.x % p.tiles_per_sample) * kCombineCells;
  if (tid < kCombineCells) {
    int cnt = 0, cs = 0;
    if (c0 + tid < p.cps) {
      const size_t gc = (size_t)b * p.cps + c0 + tid;
      cnt = p.cell_nseg[gc];
      cs = p.cell_start[gc];
    }
    s_cnt[tid] = cnt, s_cs[tid] = cs;
  }
  __syncthreads();
  const int h = lane >> 4, j = lane & 15;
  const bool to_tile = p.layout == RCB_LAYOUT_B_C_CELLS;
#pragma unroll 1
  for (int r = 0; r < kCombineCells / 16; ++r) {
    const int slot = warp * (kCombineCells / 8) + 2 * r + h;
    const int cnt = s_cnt[slot], cs = s_cs[slot];
    float a[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) a[k] = 0.f;
    for (int k0 = 0; k0 < cnt; k0 += 4) {
      int id[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) id[q] = k0 + q < cnt ? p.list[cs + k0 + q] : -1;
      float v[4][CPL];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float *src = p.rows + (size_t)(id[q] < 0 ? 0 : id[q]) * C + j;
#pragma unroll
        for (int k = 0; k < CPL; ++k) v[q][k] = id[q] >= 0 ? src[16 * k] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < CPL; ++k) a[k] += v[q][k];
    }
    if (to_tile) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) cb_tile[(j + 16 * k) * kPitch + slot] = a[k];
    } else if (c0 + slot < p.cps) {
      float *dst = p.out + ((size_t)b * p.cps + c0 + slot) * C + j;
#pragma unroll
      for (int k = 0; k < CPL; ++k) st_stream_f32(dst + 16 * k, a[k]);
    }
  }
  if (!to_tile) return;
  __syncthreads();
  for (int ch = warp; ch < C; ch += 8) {
    float *dst = p.out + ((size_t)b * C + ch) * p.cps + c0;
    if (c0 + lane < p.cps) st_stream_f32(dst + lane, cb_tile[ch * kPitch + lane]);
    if (c0 + 32 + lane < p.cps) st_stream_f32(dst + 32 + lane, cb_tile[ch * kPitch + 32 + lane]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(256) k_bwd_spread(CombineParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  constexpr int kPitch = kCombineCells + 1;
  extern __shared__ __align__(16) float cb_tile[];
  __shared__ int s_cnt[kCombineCells], s_cs[kCombineCells];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.tiles_per_sample, c0 = (blockIdx.x % p.tiles_per_sample) * kCombineCells;
  const bool from_tile = p.layout == RCB_LAYOUT_B_C_CELLS;
  if (tid < kCombineCells) {
    int cnt = 0, cs = 0;
    if (c0 + tid < p.cps) {
      const size_t gc = (size_t)b * p.cps + c0 + tid;
      cnt = p.cell_nseg[gc];
      cs = p.cell_start[gc];
    }
    s_cnt[tid] = cnt, s_cs[tid] = cs;
  }
  if (from_tile) {
    for (int ch = warp; ch < C; ch += 8) {
      const float *src = p.out + ((size_t)b * C + ch) * p.cps + c0;
      cb_tile[ch * kPitch + lane] = c0 + lane < p.cps ? ld_stream_f32(src + lane) : 0.f;
      cb_tile[ch * kPitch + 32 + lane] = c0 + 32 + lane < p.cps ? ld_stream_f32(src + 32 + lane) : 0.f;
    }
  }
  __syncthreads();
  const int h = lane >> 4, j = lane & 15;
#pragma unroll 1
  for (int r = 0; r < kCombineCells / 16; ++r) {
    const int slot = warp * (kCombineCells / 8) + 2 * r + h;
    const int cnt = s_cnt[slot], cs = s_cs[slot];
    if (cnt == 0) continue;
    float a[CPL];
    if (from_tile) {
#pragma unroll
      for (int k = 0; k < CPL; ++k) a[k] = cb_tile[(j + 16 * k) * kPitch + slot];
    } else {
      const float *src = p.out + ((size_t)b * p.cps + c0 + slot) * C + j;
#pragma unroll
      for (int k = 0; k < CPL; ++k) a[k] = ld_stream_f32(src + 16 * k);
    }
    for (int k0 = 0; k0 < cnt; ++k0) {
      float *dst = p.rows + (size_t)p.list[cs + k0] * C + j;
#pragma unroll
      for (int k = 0; k < CPL; ++k) dst[16 * k] = a[k];
    }
  }
}

struct BwdStripsParams {
  const float *depth;
  const void *feat;
  const int *status;
  const int4 *info;
  const unsigned *ent;
  const float *rows;
  float *depth_grad, *feat_grad;
  StripGeom g;
  int vec4;
};

// lane = (hc, v): v = lane % 16 is the strip's pixel, hc = lane / 16 owns channels [hc*C/2, (hc+1)*C/2)
// as Q = C/8 float4's.
template <typename FeatT, int Q>
__global__ void __launch_bounds__(64 * kStripCols, 2) k_bwd_strips(BwdStripsParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 8 * Q;
  constexpr int NL = (C + 31) / 32;  // row elements per lane when a warp moves one row
  constexpr int kFgPitch = C + 4;
  extern __shared__ __align__(16) unsigned char bs_smem[];
  const StripGeom &g = p.g;
  const int D = g.D;
  float *s_w = reinterpret_cast<float *>(bs_smem);  // [col][D][16]: depth on the way in, depth_grad on the way out
  size_t off = (size_t)kStripCols * D * kStripV * 4;
  unsigned *s_ent = reinterpret_cast<unsigned *>(bs_smem + off);
  off += (size_t)kStripCols * g.ent_cap * 4;
  float *s_g = reinterpret_cast<float *>(bs_smem + off);  // [warp][C]
  off += (size_t)2 * kStripCols * C * 4;
  float *s_fg = reinterpret_cast<float *>(bs_smem + off);  // [col][16][kFgPitch]
  off += (size_t)kStripCols * kStripV * kFgPitch * 4;
  unsigned *s_kept = reinterpret_cast<unsigned *>(bs_smem + off);  // [col][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int img, vc, u0;
  strip_block(g, blockIdx.x, img, vc, u0);
  const int v0 = vc * kStripV;

  for (int idx = tid; idx < D * kStripV; idx += 64 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, v = v0 + i;
    float c[kStripCols] = {0.f, 0.f, 0.f, 0.f};
    if (v < g.H) {
      const float *src = p.depth + ((size_t)(img * D + d) * g.HW + (size_t)v * g.W + u0);
      if (p.vec4) {
        const float4 q = ld_stream_f4(reinterpret_cast<const float4 *>(src));
        c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < kStripCols; ++k)
          if (u0 + k < g.W) c[k] = ld_stream_f32(src + k);
      }
    }
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) s_w[(k * D + d) * kStripV + i] = c[k];
  }
  for (int idx = tid; idx < kStripCols * D; idx += 64 * kStripCols) s_kept[idx] = 0u;

  const int col = warp >> 1, part = warp & 1;
  const int u = u0 + col;
  const bool active = u < g.W;
  const int strip = (img * g.VC + vc) * g.W + (active ? u : 0);
  const int4 inf = active ? p.info[strip] : make_int4(0, 0, 0, 0);
  const int e_lo = part ? inf.z : 0, e_hi = part ? inf.x : inf.z;
  const int seg_lo = part ? inf.w : 0, seg_hi = part ? inf.y : inf.w;
  const int nseg = seg_hi - seg_lo;
  unsigned *ents = s_ent + col * g.ent_cap;
  for (int e = e_lo + lane; e < e_hi; e += 32) ents[e] = p.ent[(size_t)strip * g.ent_cap + e];

  const int hc = lane >> 4, v = lane & 15;
  const bool pix_ok = active && v0 + v < g.H;
  const size_t pixel = (size_t)img * g.HW + (size_t)(pix_ok ? v0 + v : 0) * g.W + (pix_ok ? u : 0);
  float4 f[Q], fg[Q];
  {
    const char *row = static_cast<const char *>(p.feat) + (pixel * C + (size_t)hc * (C / 2)) * sizeof(FeatT);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      f[q] = pix_ok && nseg > 0 ? Row4<FeatT>::load_bytes(row + (size_t)q * 4 * sizeof(FeatT)) : make_float4(0.f, 0.f, 0.f, 0.f);
      fg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // segment rows two ahead in registers
  const float *rowp = p.rows + ((size_t)strip * g.seg_cap + seg_lo) * C;
  float r0[NL], r1[NL], r2[NL];
  auto load_row = [&](int jj, float(&r)[NL]) {
#pragma unroll
    for (int k = 0; k < NL; ++k) r[k] = (jj < nseg && 32 * k + lane < C) ? rowp[(size_t)jj * C + 32 * k + lane] : 0.f;
  };
  load_row(0, r0), load_row(1, r1), load_row(2, r2);
  __syncthreads();

  float *wcol = s_w + col * D * kStripV;
  float *gs = s_g + warp * C;
  int e = e_lo;
  for (int jj = 0; jj < nseg; ++jj) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < NL; ++k)
      if (32 * k + lane < C) gs[32 * k + lane] = r0[k];
#pragma unroll
    for (int k = 0; k < NL; ++k) r0[k] = r1[k], r1[k] = r2[k];
    load_row(jj + 3, r2);
    __syncwarp();
    // the segment's summed weight of this lane's pixel
    float wsum = 0.f;
    int e1 = e;
    for (;; ++e1) {
      const unsigned word = ents[e1];
      if ((word >> (8 + v)) & 1u) wsum += wcol[(word & 255u) * kStripV + v];
      if (word >> 24) break;
    }
    // dot(out_grad row, feat row of the pixel) and feat_grad += wsum * out_grad row
    float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
    const float2 w2 = make_float2(wsum, wsum);
    const float4 *g4 = reinterpret_cast<const float4 *>(gs + hc * (C / 2));
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float4 gq = g4[q];
      const float2 glo = make_float2(gq.x, gq.y), ghi = make_float2(gq.z, gq.w);
      d0 = __ffma2_rn(glo, make_float2(f[q].x, f[q].y), d0);
      d1 = __ffma2_rn(ghi, make_float2(f[q].z, f[q].w), d1);
      const float2 a = __ffma2_rn(glo, w2, make_float2(fg[q].x, fg[q].y));
      const float2 bq = __ffma2_rn(ghi, w2, make_float2(fg[q].z, fg[q].w));
      fg[q] = make_float4(a.x, a.y, bq.x, bq.y);
    }
    float dot = (d0.x + d0.y) + (d1.x + d1.y);
    dot += __shfl_xor_sync(kFull, dot, 16);
    // depth_grad of every bin of the segment that holds this pixel
    for (;; ++e) {
      const unsigned word = ents[e];
      const int d = word & 255u;
      if (hc == 0 && ((word >> (8 + v)) & 1u)) wcol[d * kStripV + v] = dot;
      if (lane == 0) atomicOr(&s_kept[col * D + d], (word >> 8) & 0xffffu);
      if (word >> 24) {
        ++e;
        break;
      }
    }
  }

  // feat_grad: the second warp of the strip hands its partial sums over
  float *fgs = s_fg + ((size_t)(col * kStripV + v) * kFgPitch + hc * (C / 2));
  if (part == 1) {
#pragma unroll
    for (int q = 0; q < Q; ++q) *reinterpret_cast<float4 *>(fgs + 4 * q) = fg[q];
  }
  __syncthreads();
  if (part == 0 && pix_ok) {
    float4 *dst = reinterpret_cast<float4 *>(p.feat_grad + pixel * C + (size_t)hc * (C / 2));
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float4 o = *reinterpret_cast<const float4 *>(fgs + 4 * q);
      st_stream_f4(dst + q, make_float4(fg[q].x + o.x, fg[q].y + o.y, fg[q].z + o.z, fg[q].w + o.w));
    }
  }
  for (int idx = tid; idx < D * kStripV; idx += 64 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, vv = v0 + i;
    if (vv >= g.H) continue;
    float c[kStripCols];
#pragma unroll
    for (int k = 0; k < kStripCols; ++k)
      c[k] = (s_kept[k * D + d] >> i) & 1u ? s_w[(k * D + d) * kStripV + i] : 0.f;
    float *dst = p.depth_grad + ((size_t)(img * D + d) * g.HW + (size_t)vv * g.W + u0);
    if (p.vec4) {
      st_stream_f4(reinterpret_cast<float4 *>(dst), make_float4(c[0], c[1], c[2], c[3]));
    } else {
#pragma unroll
      for (int k = 0; k < kStripCols; ++k)
        if (u0 + k < g.W) st_stream_f32(dst + k, c[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool strips_channels_ok(int C) { return C == 64 || C == 80 || C == 128; }
static bool strips_bwd_channels_ok(int C) { return C == 64 || C == 80; }

static size_t slab_bytes(const StripGeom &g) { return (size_t)kStripCols * g.D * kStripV * 4; }

template <typename K>
static int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return RCB_OK;
}

template <typename FeatT, int CPL>
static int launch_fwd_strips_t(const FwdStripsParams &p, cudaStream_t s) {
  const size_t smem = slab_bytes(p.g) + (size_t)kStripCols * p.g.ent_cap * 4;
  int rc = set_smem(k_fwd_strips<FeatT, CPL>, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(k_fwd_strips<FeatT, CPL>, (unsigned)(p.g.n_img * p.g.VC * p.g.UG), 64 * kStripCols, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_fwd_strips(const FwdStripsParams &p, int C, cudaStream_t s) {
  switch (C) {
    case 64: return launch_fwd_strips_t<FeatT, 4>(p, s);
    case 80: return launch_fwd_strips_t<FeatT, 5>(p, s);
    default: return launch_fwd_strips_t<FeatT, 8>(p, s);
  }
}

template <int CPL>
static int launch_combine_t(const CombineParams &p, int B, bool spread, cudaStream_t s) {
  const size_t smem = p.layout == RCB_LAYOUT_B_C_CELLS ? (size_t)16 * CPL * (kCombineCells + 1) * 4 : 0;
  auto kernel = spread ? k_bwd_spread<CPL> : k_fwd_combine<CPL>;
  int rc = set_smem(kernel, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(kernel, (unsigned)(B * p.tiles_per_sample), 256, smem, s, p));
  return RCB_OK;
}

static int launch_combine(const CombineParams &p, int B, int C, bool spread, cudaStream_t s) {
  switch (C) {
    case 64: return launch_combine_t<4>(p, B, spread, s);
    case 80: return launch_combine_t<5>(p, B, spread, s);
    default: return launch_combine_t<8>(p, B, spread, s);
  }
}

template <typename FeatT, int Q>
static int launch_bwd_strips_t(const BwdStripsParams &p, cudaStream_t s) {
  constexpr int C = 8 * Q;
  const size_t smem = slab_bytes(p.g) + (size_t)kStripCols * p.g.ent_cap * 4 + (size_t)2 * kStripCols * C * 4 +
                      (size_t)kStripCols * kStripV * (C + 4) * 4 + (size_t)kStripCols * p.g.D * 4;
  int rc = set_smem(k_bwd_strips<FeatT, Q>, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(k_bwd_strips<FeatT, Q>, (unsigned)(p.g.n_img * p.g.VC * p.g.UG), 64 * kStripCols, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_bwd_strips(const BwdStripsParams &p, int C, cudaStream_t s) {
  return C == 64 ? launch_bwd_strips_t<FeatT, 8>(p, s) : launch_bwd_strips_t<FeatT, 10>(p, s);
}

static bool geom_matches(const rcb_pool_desc *d, const StripGeom &g) {
  return d->n_depth == g.n_list && d->n_pixels == g.n_img * g.HW && d->B * d->Z * d->Y * d->X == g.n_cells;
}

}  // namespace rcb

using namespace rcb;

extern "C" size_t rcb_strip_plan_bytes(const rcb_strip_desc *d) {
  StripGeom g;
  if (!make_geom(d, &g)) return 0;
  return plan_view(g, nullptr).bytes;
}

extern "C" size_t rcb_strip_rows_bytes(const rcb_strip_desc *d, int C) {
  StripGeom g;
  if (!make_geom(d, &g) || C <= 0) return 0;
  return align_up((size_t)g.n_strips * g.seg_cap * C * 4, 256);
}

extern "C" int rcb_strip_plan_build(const rcb_strip_desc *d, const int *point_cell, const int *cell_start,
                                    void *plan, size_t plan_bytes, int device, rcb_stream_t stream) {
  StripGeom g;
  if (!make_geom(d, &g)) return RCB_ERR_UNSUPPORTED;
  if (!point_cell || !cell_start || !plan) return RCB_ERR_ARG;
  PlanView pv = plan_view(g, plan);
  if (plan_bytes < pv.bytes) return RCB_ERR_WORKSPACE;
  if (((uintptr_t)plan % 16) != 0) return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  RCB_CUDA_TRY(cudaMemsetAsync(pv.status, 0, 256, s));
  RCB_CUDA_TRY(cudaMemsetAsync(pv.cell_nseg, 0, (size_t)g.n_cells * 4, s));
  StripPlanParams p;
  p.point_cell = point_cell, p.cell_start = cell_start, p.status = pv.status, p.info = pv.info, p.ent = pv.ent;
  p.cell_nseg = pv.cell_nseg, p.raw = pv.raw, p.g = g;
  p.vec4 = (g.W % 4) == 0 && (((uintptr_t)point_cell) % 16) == 0;
  const size_t smem = slab_bytes(g) + (size_t)kStripCols * g.ent_cap * 8;
  int rc = set_smem(k_strip_plan, smem);
  if (rc != RCB_OK) return rc;
  k_strip_plan<<<(unsigned)(g.n_img * g.VC * g.UG), 32 * kStripCols, smem, s>>>(p);
  RCB_LAUNCH_CHECK();
  RCB_CUDA_TRY(launch_pdl(k_cellseg_sort, (unsigned)ceil_div(g.n_cells, 256), 256, 0, s, (const int *)pv.cell_nseg,
                          cell_start, (const int *)pv.raw, pv.list, g.n_cells, pv.status));
  return RCB_OK;
}

extern "C" int rcb_bev_pool_v2_fwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                                          const int *cell_start, const float *depth, const void *feat,
                                          float *out, void *rows, size_t rows_bytes, int device,
                                          rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  StripGeom g;
  if (!make_geom(sd, &g) || !strips_channels_ok(d->C) || !geom_matches(d, g)) return RCB_ERR_UNSUPPORTED;
  if (!plan || !cell_start || !depth || !feat || !out || !rows) return RCB_ERR_ARG;
  if (rows_bytes < (size_t)g.n_strips * g.seg_cap * d->C * 4) return RCB_ERR_WORKSPACE;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if ((((uintptr_t)feat) % elem) != 0 || (((uintptr_t)rows) % 16) != 0) return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  PlanView pv = plan_view(g, const_cast<void *>(plan));
  FwdStripsParams p;
  p.depth = depth, p.feat = feat, p.status = pv.status, p.info = pv.info, p.ent = pv.ent;
  p.rows = static_cast<float *>(rows), p.g = g;
  p.vec4 = (g.W % 4) == 0 && (((uintptr_t)depth) % 16) == 0;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: rc = launch_fwd_strips<float>(p, d->C, s); break;
    case RCB_DTYPE_BF16: rc = launch_fwd_strips<__nv_bfloat16>(p, d->C, s); break;
    default: rc = launch_fwd_strips<__half>(p, d->C, s);
  }
  if (rc != RCB_OK) return rc;
  CombineParams c;
  c.status = pv.status, c.cell_nseg = pv.cell_nseg, c.cell_start = cell_start, c.list = pv.list;
  c.rows = static_cast<float *>(rows), c.out = out;
  c.cps = d->Z * d->Y * d->X, c.tiles_per_sample = ceil_div(c.cps, kCombineCells), c.layout = d->layout;
  return launch_combine(c, d->B, d->C, false, s);
}

extern "C" int rcb_bev_pool_v2_bwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                                          const int *cell_start, const float *out_grad, const float *depth,
                                          const void *feat, float *depth_grad, float *feat_grad, void *rows,
                                          size_t rows_bytes, int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  StripGeom g;
  if (!make_geom(sd, &g) || !strips_bwd_channels_ok(d->C) || !geom_matches(d, g)) return RCB_ERR_UNSUPPORTED;
  if (!plan || !cell_start || !out_grad || !depth || !feat || !depth_grad || !feat_grad || !rows) return RCB_ERR_ARG;
  if (rows_bytes < (size_t)g.n_strips * g.seg_cap * d->C * 4) return RCB_ERR_WORKSPACE;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if ((((uintptr_t)feat) % (4 * elem)) != 0 || (((uintptr_t)rows) % 16) != 0 || (((uintptr_t)feat_grad) % 16) != 0)
    return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  PlanView pv = plan_view(g, const_cast<void *>(plan));
  CombineParams c;
  c.status = pv.status, c.cell_nseg = pv.cell_nseg, c.cell_start = cell_start, c.list = pv.list;
  c.rows = static_cast<float *>(rows), c.out = const_cast<float *>(out_grad);
  c.cps = d->Z * d->Y * d->X, c.tiles_per_sample = ceil_div(c.cps, kCombineCells), c.layout = d->layout;
  rc = launch_combine(c, d->B, d->C, true, s);
  if (rc != RCB_OK) return rc;
  BwdStripsParams p;
  p.depth = depth, p.feat = feat, p.status = pv.status, p.info = pv.info, p.ent = pv.ent;
  p.rows = static_cast<const float *>(rows), p.depth_grad = depth_grad, p.feat_grad = feat_grad, p.g = g;
  p.vec4 = (g.W % 4) == 0 && (((uintptr_t)depth) % 16) == 0 && (((uintptr_t)depth_grad) % 16) == 0;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: return launch_bwd_strips<float>(p, d->C, s);
    case RCB_DTYPE_BF16: return launch_bwd_strips<__nv_bfloat16>(p, d->C, s);
    default: return launch_bwd_strips<__half>(p, d->C, s);
  }
}
