// Rows F and B -- bev_pool_v2 forward / backward, STRIP-stationary kernels.
// Reference semantics: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (forward), :67-121
// (backward); host sequence mmdet3d/ops/bev_pool_v2/bev_pool.py:16-83.
//
// Geometry the design rests on: the pixels of one image column look along (nearly) the same azimuth,
// so at a given depth bin they fall into the same BEV cell (Z is collapsed), and consecutive depth
// bins of the column walk along a line of cells.  A STRIP is a vertical run of <= 16 pixels of one
// image column.  On the R50 grid the 1888 frustum points of a strip touch ~70-110 distinct cells:
// the (cell, strip) SEGMENTS are 9-13x fewer than the (cell, pixel) pairs the cell-stationary kernel
// gathers a 320-byte context row for (19-28 k against 253 k per sample).  With
//
//   W[s][v]  = sum of depth[d][v] over the depth bins d whose point (d, v) lies in segment s's cell
//   out[cell] = sum over the cell's segments s of  W[s][0..15] . feat[strip pixels 0..15][:]
//
// the pooling of a strip is a small DENSE product W (S x 16) . F (16 x C): the strip's 16 context
// rows sit in registers (each row is read from memory once per launch), W is built in shared memory
// by one sequential pass per pixel (deterministic), and only one partial row per segment crosses
// memory.  The backward is the same product transposed: per segment one out_grad row, 16 dot products
// (every depth bin of a pixel in the segment gets the pixel's dot product as depth_grad) and one
// rank-1 update W[s] x row of the strip's feat_grad registers.
//
//   plan (from point_cell alone; k_strip_plan + k_seg_prefix / k_seg_fill / k_seg_assign)
//       per point : LABEL = index of its segment inside its strip (0xffff: dropped)
//       per segment: the ROW it owns in a workspace
//       per cell  : seg_start[c] .. seg_start[c+1]: the cell's segments own consecutive rows, in
//                   (strip, segment) order -- the workspace is cell-major, so the cell-stationary
//                   passes stream it without any indirection
//   forward  k_fwd_strips  : W build + dense product, one partial row per segment
//            k_fwd_combine : cell-stationary; sums a cell's rows in order (fixed -> bit-reproducible)
//                            and writes the final layout, zeros included
//   backward k_bwd_spread  : the mirror of combine: out_grad row of a cell -> its segments' rows
//            k_bwd_strips  : W build, then per segment dot products + feat_grad update
//
// Nothing here uses float atomics; integer atomics only hand out list slots and count, and the
// lists are sorted before rows are assigned, so every output is bit-reproducible.  Inputs the plan
// cannot hold (more segments per strip / per group of eight strips than reserved, which needs cells
// scattered at random along a column) raise the plan's status word: every kernel here then exits at
// once and the caller runs the cell-/pixel-stationary kernels (pool_fwd_cells.cu, pool_bwd.cu).
#include "common.cuh"

namespace rcb {

constexpr int kStripV = 16;    // pixels per strip
constexpr int kStripCols = 8;  // adjacent image columns per CTA: 32-byte runs of depth / point_cell = whole sectors
constexpr int kGroupPix = kStripV * kStripCols;  // (pixel, column) pairs of a CTA per depth bin
#ifndef RCB_COMBINE_CELLS
#define RCB_COMBINE_CELLS 64
#endif
constexpr int kCombineCells = RCB_COMBINE_CELLS;  // cells per CTA of the combine / spread kernels (64 per lane-group slot)
constexpr int kSpreadCells = 64;
#ifndef RCB_WB_CHUNK
#define RCB_WB_CHUNK 16
#endif
#ifndef RCB_STRIP_CTAS
#define RCB_STRIP_CTAS 2
#endif
constexpr int kWChunk = RCB_WB_CHUNK;  // depth bins a W-build thread loads as one batch
constexpr int kIndexBlock = 256;  // cells per CTA of k_seg_prefix / k_seg_assign
constexpr int kNoLabel = 0xffff;
#ifndef RCB_COMBINE_CTAS
#define RCB_COMBINE_CTAS 1  // minimum resident CTAs per SM the combine / spread kernels are compiled for
#endif

struct StripGeom {
  int n_img, D, H, W, HW;
  int VC;  // strips per image column
  int UG;  // column groups per image row
  int seg_cap;    // segments a strip may have (stride of seg_dst)
  int group_cap;  // segments the strips of a CTA may have together (rows of W in shared memory)
  int ent_cap;    // (cell, depth bin) entries a strip may have while the plan is built
  int n_strips, n_groups;
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
  g->group_cap = 1536;  // rows of W a CTA can hold: 96 KB of shared memory, two CTAs of eight warps per SM
  const long long strips = (long long)d->n_img * g->VC * d->W;
  const long long groups = (long long)d->n_img * g->VC * g->UG;
  const long long points = (long long)d->n_img * d->D * g->HW;
  if (strips * g->seg_cap >= (1ll << 31) || points >= (1ll << 31) || groups * d->D * kGroupPix >= (1ll << 31)) return false;
  g->n_strips = (int)strips, g->n_groups = (int)groups, g->n_cells = d->n_cells, g->n_list = (int)points;
  return true;
}

// plan buffer: status | cell_nseg | block_sum (these three zeroed by one memset) | nseg | label |
//              seg_dst | seg_start | raw
struct PlanView {
  int *status;
  int *cell_nseg, *block_sum;
  int *nseg;                // [n_strips]
  unsigned short *label;    // [n_groups][D][16][kStripCols]
  int *seg_dst;             // [n_strips][seg_cap]
  int *seg_start, *raw;
  size_t zero_bytes, bytes;
};

static PlanView plan_view(const StripGeom &g, void *base) {
  PlanView v;
  char *p = static_cast<char *>(base);
  size_t off = 0;
  v.status = reinterpret_cast<int *>(p + off), off += 256;
  v.cell_nseg = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_cells * 4, 256);
  v.block_sum = reinterpret_cast<int *>(p + off), off += align_up((size_t)ceil_div(g.n_cells, kIndexBlock) * 4, 256);
  v.zero_bytes = off;
  v.nseg = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_strips * 4, 256);
  v.label = reinterpret_cast<unsigned short *>(p + off), off += align_up((size_t)g.n_groups * g.D * kGroupPix * 2, 256);
  v.seg_dst = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_strips * g.seg_cap * 4, 256);
  v.seg_start = reinterpret_cast<int *>(p + off), off += align_up((size_t)(g.n_cells + 1) * 4, 256);
  v.raw = reinterpret_cast<int *>(p + off), off += align_up((size_t)g.n_list * 4, 256);
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
  const int *point_cell;
  int *status;
  int *nseg;
  unsigned short *label;
  int *cell_nseg, *block_sum;
  int *seg_dst;  // plan kernel: the segment's CELL (k_seg_assign replaces it by the segment's row)
  StripGeom g;
  int vec4;
};

// One CTA per group of eight adjacent strips, one warp per strip.  The strip's (cell, depth bin)
// entries (distinct cells among the 16 pixels of a bin, with the pixel mask) are sorted by cell in
// shared memory; runs of equal cell are the segments; every point gets its segment's index as label.
__global__ void __launch_bounds__(32 * kStripCols) k_strip_plan(StripPlanParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char plan_smem[];
  __shared__ int s_segs[kStripCols];
  const StripGeom &g = p.g;
  const int D = g.D;
  int *s_cells = reinterpret_cast<int *>(plan_smem);  // [col][D][16]: cells on the way in, labels on the way out
  unsigned long long *s_buf = reinterpret_cast<unsigned long long *>(plan_smem + (size_t)kStripCols * D * kStripV * 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int img, vc, u0;
  strip_block(g, blockIdx.x, img, vc, u0);
  const int v0 = vc * kStripV;

  for (int idx = tid; idx < D * kStripV; idx += 32 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, v = v0 + i;
    int c[kStripCols];
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) c[k] = -1;
    if (v < g.H) {
      const int *src = p.point_cell + ((size_t)(img * D + d) * g.HW + (size_t)v * g.W + u0);
      if (p.vec4) {
#pragma unroll
        for (int k = 0; k < kStripCols; k += 4) {
          if (u0 + k >= g.W) break;   // W % 4 == 0: a quad is inside the image or outside
          const int4 q = *reinterpret_cast<const int4 *>(src + k);
          c[k] = q.x, c[k + 1] = q.y, c[k + 2] = q.z, c[k + 3] = q.w;
        }
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
  int segs = 0;
  if (u < g.W) {
    const int strip = (img * g.VC + vc) * g.W + u;
    int *cells = s_cells + warp * D * kStripV;
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
      if (lane == 0) atomicOr(p.status, 1);
      n = 0;
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

    // segments = runs of equal cell; label the entry's pixels; register the segment with its cell
    for (int e0 = 0; e0 < n; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < n;
      const unsigned long long cur = valid ? buf[e] : 0ull;
      const int cell = (int)(cur >> 24);
      const bool first = valid && (e == 0 || (int)(buf[e - 1] >> 24) != cell);
      const unsigned fm = __ballot_sync(kFull, first);
      const int j = segs + __popc(fm & (lanemask_lt() | (1u << lane))) - 1;
      if (valid) {
        int *row = cells + (int)((cur >> 16) & 0xffu) * kStripV;
        for (unsigned m = (unsigned)(cur & 0xffffu); m; m &= m - 1) row[__ffs(m) - 1] = j;
      }
      if (first && j < g.seg_cap) {
        p.seg_dst[(size_t)strip * g.seg_cap + j] = cell;
        atomicAdd(&p.cell_nseg[cell], 1);
        atomicAdd(&p.block_sum[cell / kIndexBlock], 1);
      }
      segs += __popc(fm);
    }
    if (lane == 0) {
      if (segs > g.seg_cap) atomicOr(p.status, 2);
      p.nseg[strip] = segs;
    }
    // every (segment, pixel) must be ONE run of consecutive depth bins (what a straight ray gives; the
    // strip kernels build their weights run by run): mark the run starts in a bitmap, a second start
    // of the same (segment, pixel) refuses the plan
    __syncwarp();
    unsigned *seen = reinterpret_cast<unsigned *>(buf);
    const int nwords = min(segs, g.seg_cap);
    for (int i = lane; i < nwords; i += 32) seen[i] = 0u;
    __syncwarp();
    bool bad = false;
    for (int d0 = 0; d0 < D; d0 += 2) {
      const int dd = d0 + half;
      if (dd >= D) continue;
      const int l = cells[dd * kStripV + i16];
      if (l < 0 || l >= nwords) continue;
      if (dd > 0 && cells[(dd - 1) * kStripV + i16] == l) continue;
      bad |= (atomicOr(&seen[l], 1u << i16) >> i16) & 1u;
    }
    if (bad) atomicOr(p.status, 16);
  }
  if (lane == 0) s_segs[warp] = segs;
  __syncthreads();
  if (tid == 0) {
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) sum += s_segs[k];
    if (sum > g.group_cap) atomicOr(p.status, 8);
  }

  // labels leave as [group][d][pixel][column]: 16 bytes per (d, pixel), consecutive
  static_assert(kStripCols == 8, "label rows are written as one 128-bit word");
  unsigned short *lab = p.label + (size_t)blockIdx.x * D * kGroupPix;
  for (int idx = tid; idx < D * kStripV; idx += 32 * kStripCols) {
    unsigned q[kStripCols / 2];
#pragma unroll
    for (int k = 0; k < kStripCols; k += 2)
      q[k / 2] = (unsigned)(unsigned short)s_cells[(k * D) * kStripV + idx] |
                 ((unsigned)(unsigned short)s_cells[((k + 1) * D) * kStripV + idx] << 16);
    reinterpret_cast<uint4 *>(lab)[idx] = make_uint4(q[0], q[1], q[2], q[3]);
  }
}

constexpr int kMaxListSort = 1024;

__device__ __forceinline__ void cswap(int &a, int &b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo, b = hi;
}

// The per-cell segment lists need no input besides the plan kernel's own counts (in particular not the
// sorted ranks' CSR: the plan is built from point_cell alone, so a chain that never exposes ranks can
// skip the sort altogether -- rcb_voxel_pooling chain in view_pool.py):
//   k_seg_prefix : seg_start = exclusive prefix of the per-cell segment counts (block totals were
//                  counted by the plan kernel, so a CTA only sums the totals of the blocks before it)
//   k_seg_fill   : every segment claims a slot of its cell's list (integer countdown of the cell's count)
//   k_seg_assign : the cell's list sorted ascending by (strip, segment); the segment at rank r owns the
//                  workspace row seg_start[c] + r -- the claimed order of the atomics never reaches an
//                  output.  Lists of <= 8 (all but a few hundred cells next to the cameras) are sorted
//                  by their own thread in registers, longer ones by the warp.
__global__ void __launch_bounds__(kIndexBlock) k_seg_prefix(const int *cell_nseg, const int *block_sum, int *seg_start,
                                                            int n_cells) {
  pdl_prologue();
  __shared__ int s_part[kIndexBlock / 32], s_wsum[kIndexBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x * kIndexBlock + tid;
  const int cnt = c < n_cells ? cell_nseg[c] : 0;
  int part = 0;
  for (int i = tid; i < (int)blockIdx.x; i += kIndexBlock) part += block_sum[i];
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
    part += __shfl_xor_sync(kFull, part, o);
  }
  if (lane == 31) s_wsum[warp] = incl;
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  int start = incl - cnt;
#pragma unroll
  for (int w = 0; w < kIndexBlock / 32; ++w) start += s_part[w] + (w < warp ? s_wsum[w] : 0);
  if (c < n_cells) seg_start[c] = start;
  if (c == n_cells - 1) seg_start[n_cells] = start + cnt;
}

constexpr int kFillThreads = 128;

// one CTA per strip
__global__ void __launch_bounds__(kFillThreads) k_seg_fill(const int *nseg, const int *seg_dst, const int *seg_start,
                                                           int *cell_nseg, int *raw, int seg_cap) {
  pdl_prologue();
  const int strip = blockIdx.x;
  const int n = min(nseg[strip], seg_cap);
  for (int j = threadIdx.x; j < n; j += kFillThreads) {
    const int id = strip * seg_cap + j;
    const int cell = seg_dst[id];
    const int k = atomicSub(&cell_nseg[cell], 1) - 1;
    raw[seg_start[cell] + k] = id;
  }
}

__global__ void __launch_bounds__(kIndexBlock) k_seg_assign(const int *seg_start, const int *raw, int *seg_dst,
                                                            int n_cells, int *status) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * kIndexBlock + threadIdx.x;
  int cnt = 0, cs = 0;
  if (c < n_cells) {
    cs = seg_start[c];
    cnt = seg_start[c + 1] - cs;
  }
  const int start = cs;
  if (cnt >= 1 && cnt <= 8) {
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = k < cnt ? raw[cs + k] : 0x7fffffff;
    if (cnt >= 2) {  // Batcher's odd-even merge network for 8 keys (19 exchanges)
      cswap(v[0], v[1]), cswap(v[2], v[3]), cswap(v[4], v[5]), cswap(v[6], v[7]);
      cswap(v[0], v[2]), cswap(v[1], v[3]), cswap(v[4], v[6]), cswap(v[5], v[7]);
      cswap(v[1], v[2]), cswap(v[5], v[6]);
      cswap(v[0], v[4]), cswap(v[1], v[5]), cswap(v[2], v[6]), cswap(v[3], v[7]);
      cswap(v[2], v[4]), cswap(v[3], v[5]);
      cswap(v[1], v[2]), cswap(v[3], v[4]), cswap(v[5], v[6]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < cnt) seg_dst[v[k]] = start + k;
  }
  unsigned multi = __ballot_sync(kFull, cnt > 8);
  while (multi) {
    const int src = __ffs(multi) - 1;
    multi &= multi - 1;
    const int n = __shfl_sync(kFull, cnt, src), b = __shfl_sync(kFull, cs, src), st = __shfl_sync(kFull, start, src);
    if (n <= 32) {
      const int id = lane < n ? raw[b + lane] : 0x7fffffff;
      int rank = 0;
      for (int m = 0; m < n; ++m) rank += __shfl_sync(kFull, id, m) < id;
      if (lane < n) seg_dst[id] = st + rank;
    } else if (n <= kMaxListSort) {
      for (int e = lane; e < n; e += 32) {
        const int id = raw[b + e];
        int rank = 0;
        for (int m = 0; m < n; ++m) rank += raw[b + m] < id;
        seg_dst[id] = st + rank;
      }
    } else if (lane == 0) {
      atomicOr(status, 4);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// strip passes
// ------------------------------------------------------------------------------------------------
struct StripsParams {
  const float *depth;
  const void *feat;
  const int *status;
  const int *nseg;
  const unsigned short *label;
  const int *seg_dst;
  float *rows;  // forward: written; backward: read
  float *depth_grad, *feat_grad;
  StripGeom g;
  int vec4;
};

// Shared prologue of the two strip kernels.  After it: s_W[row][16] holds, for the CTA's eight strips
// back to back (strip k's segment s is row base[k] + s), the summed depth weight of every (segment,
// pixel); s_dst[row] the workspace row of the segment.
struct StripCta {
  int img, vc, u0, v0, strip0;
  int base[kStripCols + 1];
  // base[i] without dynamic indexing (keeps the array in registers)
  __device__ __forceinline__ int at(int i) const {
    int r = base[0];
#pragma unroll
    for (int k = 1; k <= kStripCols; ++k) r = i == k ? base[k] : r;
    return r;
  }
};

__device__ __forceinline__ void strip_prologue(const StripsParams &p, float *s_W, int *s_dst, StripCta &c) {
  const StripGeom &g = p.g;
  const int D = g.D;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  strip_block(g, blockIdx.x, c.img, c.vc, c.u0);
  c.v0 = c.vc * kStripV;
  c.strip0 = (c.img * g.VC + c.vc) * g.W + c.u0;
  c.base[0] = 0;
#pragma unroll
  for (int k = 0; k < kStripCols; ++k) c.base[k + 1] = c.base[k] + (c.u0 + k < g.W ? p.nseg[c.strip0 + k] : 0);
  const int total = c.base[kStripCols];
  for (int i = tid; i < total * 4; i += 32 * kStripCols) reinterpret_cast<float4 *>(s_W)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    const int b0 = c.at(warp), n = c.at(warp + 1) - b0;
    for (int q = lane; q < n; q += 32) s_dst[b0 + q] = p.seg_dst[(size_t)(c.strip0 + warp) * g.seg_cap + q];
  }
  __syncthreads();
  // W[segment][pixel] = sum of the pixel's depth weights over its run of bins with that label.  A
  // straight ray is inside a (convex) cell over ONE interval of depth, so every (segment, pixel) has
  // exactly one run (the plan checks it and refuses inputs where it does not hold) and a run's sum is
  // simply stored.  Two threads per (pixel, column), each taking half of the depth bins in chunks of
  // 32 that are loaded as a batch (one memory latency per chunk) and then walked in registers.  A run
  // belongs to the thread in whose half it starts; that thread follows it across the boundary.
  {
    const int dh = tid / kGroupPix, vk = tid % kGroupPix, v = vk / kStripCols, k = vk % kStripCols;
    if (c.v0 + v < g.H && c.u0 + k < g.W) {
      const int dlen = (D + 1) >> 1, d_begin = dh * dlen, d_end = min(D, d_begin + dlen);
      const unsigned short *lab = p.label + (size_t)blockIdx.x * D * kGroupPix + vk;
      const float *w = p.depth + ((size_t)c.img * D * g.HW + (size_t)(c.v0 + v) * g.W + c.u0 + k);
      float *wrow = s_W + (size_t)c.at(k) * kStripV + v;
      int cur = d_begin > 0 ? (int)lab[(size_t)(d_begin - 1) * kGroupPix] : kNoLabel;
      bool own = false;
      float acc = 0.f;
      for (int d0 = d_begin; d0 < d_end; d0 += kWChunk) {
        int l[kWChunk];
        float x[kWChunk];
#pragma unroll
        for (int i = 0; i < kWChunk; ++i) {
          const bool ok = d0 + i < d_end;
          l[i] = ok ? (int)lab[(size_t)(d0 + i) * kGroupPix] : -1;
          x[i] = ok ? ld_stream_f32(w + (size_t)(d0 + i) * g.HW) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kWChunk; ++i) {
          if (l[i] < 0) break;
          if (l[i] != cur) {
            if (own && cur != kNoLabel) wrow[cur * kStripV] = acc;
            cur = l[i], acc = 0.f, own = true;
          }
          acc += x[i];
        }
      }
      if (own && cur != kNoLabel) {
        for (int dd = d_end; dd < D && (int)lab[(size_t)dd * kGroupPix] == cur; ++dd) acc += ld_stream_f32(w + (size_t)dd * g.HW);
        wrow[cur * kStripV] = acc;
      }
    }
  }
  __syncthreads();
}

// lane = (h, j): h = lane / 16 owns pixels 8h .. 8h+7 of the strip, j = lane % 16 owns channels
// j + 16k, k < CPL (C = 16 * CPL).  fp[tp][k] = (feat[pixel 2tp][channel], feat[pixel 2tp+1][channel]):
// the packed FMAs pair two PIXELS, so their weight operand is a register pair straight out of the
// 128-bit load of the W row.
template <typename FeatT, int CPL>
__device__ __forceinline__ void load_strip_feat(const StripsParams &p, const StripCta &c, int col, bool wanted,
                                                float2 (&fp)[4][CPL]) {
  constexpr int C = 16 * CPL;
  const StripGeom &g = p.g;
  const int lane = threadIdx.x & 31, h = lane >> 4, j = lane & 15;
  const FeatT *feat = static_cast<const FeatT *>(p.feat);
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int v = c.v0 + 8 * h + t;
    const bool ok = wanted && v < g.H;
    const FeatT *row = feat + ((size_t)c.img * g.HW + (size_t)(ok ? v : 0) * g.W + (ok ? c.u0 + col : 0)) * C + j;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const float x = ok ? to_f32<FeatT>(row[16 * k]) : 0.f;
      if (t & 1) fp[t >> 1][k].y = x;
      else fp[t >> 1][k].x = x;
    }
  }
}

template <typename FeatT, int CPL>
__global__ void __launch_bounds__(32 * kStripCols, CPL <= 5 ? RCB_STRIP_CTAS : 1) k_fwd_strips(StripsParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  extern __shared__ __align__(16) unsigned char fs_smem[];
  float *s_W = reinterpret_cast<float *>(fs_smem);
  int *s_dst = reinterpret_cast<int *>(fs_smem + (size_t)p.g.group_cap * kStripV * 4);
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  StripCta c;
  // the feature rows are requested before the W build so that their latency hides behind it
  strip_block(p.g, blockIdx.x, c.img, c.vc, c.u0);
  c.v0 = c.vc * kStripV;
  float2 fp[4][CPL];
  load_strip_feat<FeatT, CPL>(p, c, col, c.u0 + col < p.g.W, fp);
  strip_prologue(p, s_W, s_dst, c);

  const int h = lane >> 4, j = lane & 15;
  const int b0 = c.at(col), n = c.at(col + 1) - b0;
  const float *wrow = s_W + (size_t)b0 * kStripV + 8 * h;
  const int *dsts = s_dst + b0;
#pragma unroll 4
  for (int sg = 0; sg < n; ++sg) {
    const float4 wa = *reinterpret_cast<const float4 *>(wrow + sg * kStripV);
    const float4 wb = *reinterpret_cast<const float4 *>(wrow + sg * kStripV + 4);
    const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
    float a[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) acc = __ffma2_rn(fp[tp][k], w2[tp], acc);
      a[k] = acc.x + acc.y;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) a[k] += __shfl_xor_sync(kFull, a[k], 16);
    float *dst = p.rows + (size_t)dsts[sg] * C + j;
#pragma unroll
    for (int k = 0; k < CPL; ++k)
      if ((k & 1) == h) dst[16 * k] = a[k];
  }
}

// Backward: same layout, fgp = the strip's feat_grad as pixel pairs.  Per segment the out_grad row
// arrives as CPL coalesced words per lane (prefetched one segment ahead); the dot products of all 16
// pixels cost CPL * 4 packed FMAs + a transposed butterfly over the 16 channel lanes (8 shuffles: lane
// j ends up with pixel (j >> 1) % 8), and overwrite the segment's W row once it has been consumed by
// the feat_grad update.  At the end depth_grad[d][v] = W[label[d][v]][v], written coalesced.
template <typename FeatT, int CPL>
__global__ void __launch_bounds__(32 * kStripCols, RCB_STRIP_CTAS) k_bwd_strips(StripsParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  extern __shared__ __align__(16) unsigned char bs_smem[];
  float *s_W = reinterpret_cast<float *>(bs_smem);
  int *s_dst = reinterpret_cast<int *>(bs_smem + (size_t)p.g.group_cap * kStripV * 4);
  const StripGeom &g = p.g;
  const int tid = threadIdx.x, lane = tid & 31, col = tid >> 5;
  StripCta c;
  strip_block(g, blockIdx.x, c.img, c.vc, c.u0);
  c.v0 = c.vc * kStripV;
  const bool active = c.u0 + col < g.W;
  float2 fp[4][CPL], fgp[4][CPL];
  load_strip_feat<FeatT, CPL>(p, c, col, active, fp);
#pragma unroll
  for (int tp = 0; tp < 4; ++tp)
#pragma unroll
    for (int k = 0; k < CPL; ++k) fgp[tp][k] = make_float2(0.f, 0.f);
  strip_prologue(p, s_W, s_dst, c);

  const int h = lane >> 4, j = lane & 15;
  const int b0 = c.at(col), n = c.at(col + 1) - b0;
  if (n > 0) {
    float *wrow = s_W + (size_t)b0 * kStripV + 8 * h;
    const int *dsts = s_dst + b0;
    const int my_t = (j >> 1) & 7;  // the pixel (of this half) whose dot product ends up in this lane
    const bool b8 = j & 8, b4 = j & 4, b2 = j & 2, writer = (j & 1) == 0;
    float gr[CPL];
    {
      const float *src = p.rows + (size_t)dsts[0] * C + j;
#pragma unroll
      for (int k = 0; k < CPL; ++k) gr[k] = src[16 * k];
    }
    for (int sg = 0; sg < n; ++sg) {
      float gn[CPL];
      {
        const float *src = p.rows + (size_t)dsts[sg + 1 < n ? sg + 1 : sg] * C + j;
#pragma unroll
        for (int k = 0; k < CPL; ++k) gn[k] = src[16 * k];
      }
      const float4 wa = *reinterpret_cast<const float4 *>(wrow + sg * kStripV);
      const float4 wb = *reinterpret_cast<const float4 *>(wrow + sg * kStripV + 4);
      const float2 w2[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
      float2 dp[4];
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) dp[tp] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const float2 g2 = make_float2(gr[k], gr[k]);
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
          dp[tp] = __ffma2_rn(fp[tp][k], g2, dp[tp]);
          fgp[tp][k] = __ffma2_rn(w2[tp], g2, fgp[tp][k]);
        }
      }
      // transposed butterfly over the 16 channel lanes of the half
      const float2 sa = b8 ? dp[0] : dp[2], sb = b8 ? dp[1] : dp[3];
      float2 ka = b8 ? dp[2] : dp[0], kb = b8 ? dp[3] : dp[1];
      ka.x += __shfl_xor_sync(kFull, sa.x, 8), ka.y += __shfl_xor_sync(kFull, sa.y, 8);
      kb.x += __shfl_xor_sync(kFull, sb.x, 8), kb.y += __shfl_xor_sync(kFull, sb.y, 8);
      const float2 s2 = b4 ? ka : kb;
      float2 k2 = b4 ? kb : ka;
      k2.x += __shfl_xor_sync(kFull, s2.x, 4), k2.y += __shfl_xor_sync(kFull, s2.y, 4);
      const float s3 = b2 ? k2.x : k2.y;
      float dot = b2 ? k2.y : k2.x;
      dot += __shfl_xor_sync(kFull, s3, 2);
      dot += __shfl_xor_sync(kFull, dot, 1);
      if (writer) wrow[sg * kStripV + my_t] = dot;  // every lane of the half read the row above (shuffles in between)
#pragma unroll
      for (int k = 0; k < CPL; ++k) gr[k] = gn[k];
    }
  }

  if (active) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int v = c.v0 + 8 * h + t;
      if (v >= g.H) continue;
      float *row = p.feat_grad + ((size_t)c.img * g.HW + (size_t)v * g.W + c.u0 + col) * C + j;
#pragma unroll
      for (int k = 0; k < CPL; ++k) st_stream_f32(row + 16 * k, (t & 1) ? fgp[t >> 1][k].y : fgp[t >> 1][k].x);
    }
  }
  __syncthreads();
  const int D = g.D;
  const uint4 *lab = reinterpret_cast<const uint4 *>(p.label + (size_t)blockIdx.x * D * kGroupPix);
  for (int idx = tid; idx < D * kStripV; idx += 32 * kStripCols) {
    const int d = idx >> 4, i = idx & 15, vv = c.v0 + i;
    if (vv >= g.H) continue;
    const uint4 q = lab[idx];
    const unsigned qq[4] = {q.x, q.y, q.z, q.w};
    float o[kStripCols];
#pragma unroll
    for (int k = 0; k < kStripCols; ++k) {
      const int l = (qq[k >> 1] >> (16 * (k & 1))) & 0xffff;
      o[k] = l != kNoLabel ? s_W[(size_t)(c.base[k] + l) * kStripV + i] : 0.f;
    }
    float *dst = p.depth_grad + ((size_t)(c.img * D + d) * g.HW + (size_t)vv * g.W + c.u0);
    if (p.vec4) {
#pragma unroll
      for (int k = 0; k < kStripCols; k += 4)
        if (c.u0 + k < g.W) st_stream_f4(reinterpret_cast<float4 *>(dst + k), make_float4(o[k], o[k + 1], o[k + 2], o[k + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < kStripCols; ++k)
        if (c.u0 + k < g.W) st_stream_f32(dst + k, o[k]);
    }
  }
}

struct CombineParams {
  const int *status;
  const int *seg_start;
  float *rows;  // forward: read; backward spread: written
  float *out;   // forward: the pooled tensor; backward spread: out_grad (read)
  int cps, tiles_per_sample, layout;
};

// Four lanes per cell (eight cells per warp): lane `sub` owns the 128-bit channel quads sub + 4q of the
// cell's rows.  The rows of consecutive cells are consecutive in the workspace: the CTA streams one
// contiguous piece of it.  A (B, C, cells) store instruction writes four channel rows x eight
// consecutive cells (whole 32-byte sectors).  No shared memory, no barrier, no indirection.
template <int CPL>
__global__ void __launch_bounds__(256, RCB_COMBINE_CTAS) k_fwd_combine(CombineParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  constexpr int kPer = kCombineCells / 64;  // cells per lane group: their loads are in flight together
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.tiles_per_sample, c0 = (blockIdx.x % p.tiles_per_sample) * kCombineCells;
  const int sub = lane & 3;
  int cell[kPer], r0[kPer], n[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    cell[u] = c0 + (warp * kPer + u) * 8 + (lane >> 2);
    const bool ok = cell[u] < p.cps;
    const size_t gc = (size_t)b * p.cps + (ok ? cell[u] : 0);
    r0[u] = p.seg_start[gc];
    n[u] = ok ? p.seg_start[gc + 1] - r0[u] : -1;
  }
  float4 acc[kPer][CPL];
  int nmax = 0;
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    nmax = max(nmax, n[u]);
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = 0; i < nmax; ++i) {
    float4 v[kPer][CPL];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const float4 *src = reinterpret_cast<const float4 *>(p.rows + (size_t)(r0[u] + (i < n[u] ? i : 0)) * C) + sub;
#pragma unroll
      for (int q = 0; q < CPL; ++q) v[u][q] = i < n[u] ? src[4 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u)
#pragma unroll
      for (int q = 0; q < CPL; ++q)
        acc[u][q].x += v[u][q].x, acc[u][q].y += v[u][q].y, acc[u][q].z += v[u][q].z, acc[u][q].w += v[u][q].w;
  }
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    if (n[u] < 0) continue;
    if (p.layout == RCB_LAYOUT_B_C_CELLS) {
      float *dst = p.out + (size_t)b * C * p.cps + cell[u];
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const size_t ch = (size_t)(sub + 4 * q) * 4;
        st_stream_f32(dst + (ch + 0) * p.cps, acc[u][q].x);
        st_stream_f32(dst + (ch + 1) * p.cps, acc[u][q].y);
        st_stream_f32(dst + (ch + 2) * p.cps, acc[u][q].z);
        st_stream_f32(dst + (ch + 3) * p.cps, acc[u][q].w);
      }
    } else {
      float4 *dst = reinterpret_cast<float4 *>(p.out + ((size_t)b * p.cps + cell[u]) * C) + sub;
#pragma unroll
      for (int q = 0; q < CPL; ++q) st_stream_f4(dst + 4 * q, acc[u][q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// the mirror image of k_fwd_combine: the out_grad row of a cell goes to each of the cell's segments
template <int CPL>
__global__ void __launch_bounds__(256, RCB_COMBINE_CTAS) k_bwd_spread(CombineParams p) {
  pdl_prologue();
  if (*p.status != 0) return;
  constexpr int C = 16 * CPL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / p.tiles_per_sample, c0 = (blockIdx.x % p.tiles_per_sample) * kSpreadCells;
  const int cell = c0 + warp * 8 + (lane >> 2), sub = lane & 3;
  if (cell >= p.cps) return;
  const size_t gc = (size_t)b * p.cps + cell;
  const int r0 = p.seg_start[gc], r1 = p.seg_start[gc + 1];
  if (r0 >= r1) return;
  float4 a[CPL];
  if (p.layout == RCB_LAYOUT_B_C_CELLS) {
    const float *src = p.out + (size_t)b * C * p.cps + cell;
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      const size_t ch = (size_t)(sub + 4 * q) * 4;
      a[q] = make_float4(ld_stream_f32(src + (ch + 0) * p.cps), ld_stream_f32(src + (ch + 1) * p.cps),
                         ld_stream_f32(src + (ch + 2) * p.cps), ld_stream_f32(src + (ch + 3) * p.cps));
    }
  } else {
    const float4 *src = reinterpret_cast<const float4 *>(p.out + gc * C) + sub;
#pragma unroll
    for (int q = 0; q < CPL; ++q) a[q] = ld_stream_f4(src + 4 * q);
  }
  float4 *dst = reinterpret_cast<float4 *>(p.rows + (size_t)r0 * C) + sub;
  for (int r = r0; r < r1; ++r, dst += C / 4) {
#pragma unroll
    for (int q = 0; q < CPL; ++q) dst[4 * q] = a[q];
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool strips_channels_ok(int C) { return C == 64 || C == 80 || C == 128; }
static bool strips_bwd_channels_ok(int C) { return C == 64 || C == 80; }

template <typename K>
static int set_smem(K kernel, size_t smem) {
  if (smem > 48 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return RCB_OK;
}

static size_t strips_smem(const StripGeom &g) { return (size_t)g.group_cap * (kStripV * 4 + 4); }

template <typename FeatT, int CPL>
static int launch_fwd_strips_t(const StripsParams &p, cudaStream_t s) {
  const size_t smem = strips_smem(p.g);
  int rc = set_smem(k_fwd_strips<FeatT, CPL>, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(k_fwd_strips<FeatT, CPL>, (unsigned)p.g.n_groups, 32 * kStripCols, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_fwd_strips(const StripsParams &p, int C, cudaStream_t s) {
  switch (C) {
    case 64: return launch_fwd_strips_t<FeatT, 4>(p, s);
    case 80: return launch_fwd_strips_t<FeatT, 5>(p, s);
    default: return launch_fwd_strips_t<FeatT, 8>(p, s);
  }
}

template <int CPL>
static int launch_combine_t(const CombineParams &p, int B, bool spread, cudaStream_t s) {
  auto kernel = spread ? k_bwd_spread<CPL> : k_fwd_combine<CPL>;
  RCB_CUDA_TRY(launch_pdl(kernel, (unsigned)(B * p.tiles_per_sample), 256, 0, s, p));
  return RCB_OK;
}

static int launch_combine(const CombineParams &p, int B, int C, bool spread, cudaStream_t s) {
  switch (C) {
    case 64: return launch_combine_t<4>(p, B, spread, s);
    case 80: return launch_combine_t<5>(p, B, spread, s);
    default: return launch_combine_t<8>(p, B, spread, s);
  }
}

template <typename FeatT, int CPL>
static int launch_bwd_strips_t(const StripsParams &p, cudaStream_t s) {
  const size_t smem = strips_smem(p.g);
  int rc = set_smem(k_bwd_strips<FeatT, CPL>, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(k_bwd_strips<FeatT, CPL>, (unsigned)p.g.n_groups, 32 * kStripCols, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_bwd_strips(const StripsParams &p, int C, cudaStream_t s) {
  return C == 64 ? launch_bwd_strips_t<FeatT, 4>(p, s) : launch_bwd_strips_t<FeatT, 5>(p, s);
}

static bool geom_matches(const rcb_pool_desc *d, const StripGeom &g) {
  return d->n_depth == g.n_list && d->n_pixels == g.n_img * g.HW && d->B * d->Z * d->Y * d->X == g.n_cells;
}

static StripsParams strips_params(const StripGeom &g, const PlanView &pv, const float *depth, const void *feat,
                                  float *rows) {
  StripsParams p;
  p.depth = depth, p.feat = feat, p.status = pv.status, p.nseg = pv.nseg, p.label = pv.label, p.seg_dst = pv.seg_dst;
  p.rows = rows, p.depth_grad = nullptr, p.feat_grad = nullptr, p.g = g;
  p.vec4 = 0;
  return p;
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
  (void)cell_start;  // kept in the signature for callers of the first revision; the plan no longer reads it
  if (!point_cell || !plan) return RCB_ERR_ARG;
  PlanView pv = plan_view(g, plan);
  if (plan_bytes < pv.bytes) return RCB_ERR_WORKSPACE;
  if (((uintptr_t)plan % 16) != 0) return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  RCB_CUDA_TRY(cudaMemsetAsync(plan, 0, pv.zero_bytes, s));  // status, segment counts per cell and per block
  StripPlanParams p;
  p.point_cell = point_cell, p.status = pv.status, p.nseg = pv.nseg, p.label = pv.label;
  p.cell_nseg = pv.cell_nseg, p.block_sum = pv.block_sum, p.seg_dst = pv.seg_dst, p.g = g;
  p.vec4 = (g.W % 4) == 0 && (((uintptr_t)point_cell) % 16) == 0;
  const size_t smem = (size_t)kStripCols * g.D * kStripV * 4 + (size_t)kStripCols * g.ent_cap * 8;
  int rc = set_smem(k_strip_plan, smem);
  if (rc != RCB_OK) return rc;
  RCB_CUDA_TRY(launch_pdl(k_strip_plan, (unsigned)g.n_groups, 32 * kStripCols, smem, s, p));
  const unsigned index_blocks = (unsigned)ceil_div(g.n_cells, kIndexBlock);
  RCB_CUDA_TRY(launch_pdl(k_seg_prefix, index_blocks, kIndexBlock, 0, s, (const int *)pv.cell_nseg,
                          (const int *)pv.block_sum, pv.seg_start, g.n_cells));
  RCB_CUDA_TRY(launch_pdl(k_seg_fill, (unsigned)g.n_strips, kFillThreads, 0, s, (const int *)pv.nseg,
                          (const int *)pv.seg_dst, (const int *)pv.seg_start, pv.cell_nseg, pv.raw, g.seg_cap));
  RCB_CUDA_TRY(launch_pdl(k_seg_assign, index_blocks, kIndexBlock, 0, s, (const int *)pv.seg_start,
                          (const int *)pv.raw, pv.seg_dst, g.n_cells, pv.status));
  return RCB_OK;
}

extern "C" int rcb_bev_pool_v2_fwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                                          const float *depth, const void *feat, float *out, void *rows,
                                          size_t rows_bytes, int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  StripGeom g;
  if (!make_geom(sd, &g) || !strips_channels_ok(d->C) || !geom_matches(d, g)) return RCB_ERR_UNSUPPORTED;
  if (!plan || !depth || !feat || !out || !rows) return RCB_ERR_ARG;
  if (rows_bytes < (size_t)g.n_strips * g.seg_cap * d->C * 4) return RCB_ERR_WORKSPACE;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if ((((uintptr_t)feat) % elem) != 0 || (((uintptr_t)rows) % 16) != 0) return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  PlanView pv = plan_view(g, const_cast<void *>(plan));
  StripsParams p = strips_params(g, pv, depth, feat, static_cast<float *>(rows));
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: rc = launch_fwd_strips<float>(p, d->C, s); break;
    case RCB_DTYPE_BF16: rc = launch_fwd_strips<__nv_bfloat16>(p, d->C, s); break;
    default: rc = launch_fwd_strips<__half>(p, d->C, s);
  }
  if (rc != RCB_OK) return rc;
  CombineParams c;
  c.status = pv.status, c.seg_start = pv.seg_start;
  c.rows = static_cast<float *>(rows), c.out = out;
  c.cps = d->Z * d->Y * d->X, c.tiles_per_sample = ceil_div(c.cps, kCombineCells), c.layout = d->layout;
  return launch_combine(c, d->B, d->C, false, s);
}

extern "C" int rcb_bev_pool_v2_bwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                                          const float *out_grad, const float *depth, const void *feat,
                                          float *depth_grad, float *feat_grad, void *rows, size_t rows_bytes,
                                          int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  StripGeom g;
  if (!make_geom(sd, &g) || !strips_bwd_channels_ok(d->C) || !geom_matches(d, g)) return RCB_ERR_UNSUPPORTED;
  if (!plan || !out_grad || !depth || !feat || !depth_grad || !feat_grad || !rows) return RCB_ERR_ARG;
  if (rows_bytes < (size_t)g.n_strips * g.seg_cap * d->C * 4) return RCB_ERR_WORKSPACE;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if ((((uintptr_t)feat) % elem) != 0 || (((uintptr_t)rows) % 16) != 0) return RCB_ERR_ALIGN;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  PlanView pv = plan_view(g, const_cast<void *>(plan));
  CombineParams c;
  c.status = pv.status, c.seg_start = pv.seg_start;
  c.rows = static_cast<float *>(rows), c.out = const_cast<float *>(out_grad);
  c.cps = d->Z * d->Y * d->X, c.tiles_per_sample = ceil_div(c.cps, kSpreadCells), c.layout = d->layout;
  rc = launch_combine(c, d->B, d->C, true, s);
  if (rc != RCB_OK) return rc;
  StripsParams p = strips_params(g, pv, depth, feat, static_cast<float *>(rows));
  p.depth_grad = depth_grad, p.feat_grad = feat_grad;
  p.vec4 = (g.W % 4) == 0 && (((uintptr_t)depth_grad) % 16) == 0;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: return launch_bwd_strips<float>(p, d->C, s);
    case RCB_DTYPE_BF16: return launch_bwd_strips<__nv_bfloat16>(p, d->C, s);
    default: return launch_bwd_strips<__half>(p, d->C, s);
  }
}
