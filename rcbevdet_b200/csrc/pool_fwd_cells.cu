// Row F -- bev_pool_v2 forward, cell-stationary kernel pair (the product path for ranks whose cells
// are sorted, i.e. everything prepare emits and everything rcb_pool_validate proves sorted).
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48; the zero-fill and the permute copy of
// mmdet3d/ops/bev_pool_v2/bev_pool.py:27,91 are folded in (every cell written once, final layout).
//
//   k_fwd_entries  one thread per sorted point: (context-row byte offset, depth weight) pairs, 8 bytes
//                  each, in sorted order.  Consecutive points of one cell that read the SAME row -- the
//                  depth bins of one (cell, pixel) pair, adjacent in prepare's (pixel, depth) order,
//                  1.42 per pair on the R50 grid -- are merged here: the first carries the sum of their
//                  depth weights (added in point order), the others are marked "skip".  This is also
//                  where the depth gather happens, with one independent load per thread instead of a
//                  dependent one inside the pooling loop.
//   k_fwd_cells    one CTA per 8 x 4 patch of BEV cells, thread <-> (cell, 128-bit channel quad).
//                  A thread walks the entries of ITS cell straight from global memory (broadcast
//                  128-bit loads of two entries, L1-resident lines), loads its quad of every
//                  non-skipped row through L1 -- the patch shape makes neighbouring cells, which see
//                  the same pixels along a ray, share rows there -- and accumulates in registers.  No
//                  shared-memory staging, no partial sums, no barrier before the write-out; the patch
//                  leaves through a shared-memory transpose as (B, C, Z*Y*X) runs of 8 cells per
//                  channel row, or directly as channels-last rows.  Empty cells get their zeros here.
//
// The order of additions inside a cell is the point order (runs folded first): bit-reproducible.
#include "common.cuh"

namespace rcb {

constexpr int kCellsTileX = 8;
constexpr int kCellsTileY = 4;
constexpr unsigned kSkipEntry = 0xffffffffu;

struct FwdEntriesParams {
  const float *depth;
  const int *ranks_depth, *ranks_feat, *ranks_bev;
  const int *n_ptr;  // number of valid points (cell_start[n_cells]) -- a device value in the fused chain
  uint2 *entries;
  unsigned row_bytes;
};

__global__ void __launch_bounds__(256) k_fwd_entries(FwdEntriesParams p) {
  pdl_prologue();
  const int n = __ldg(p.n_ptr);
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int rb = __ldg(p.ranks_bev + i), rf = __ldg(p.ranks_feat + i);
    if (i > 0 && __ldg(p.ranks_bev + i - 1) == rb && __ldg(p.ranks_feat + i - 1) == rf) {
      p.entries[i] = make_uint2(kSkipEntry, 0u);
      continue;
    }
    float w = ld_stream_f32(p.depth + __ldg(p.ranks_depth + i));
    for (int j = i + 1; j < n && __ldg(p.ranks_bev + j) == rb && __ldg(p.ranks_feat + j) == rf; ++j)
      w += ld_stream_f32(p.depth + __ldg(p.ranks_depth + j));
    p.entries[i] = make_uint2((unsigned)rf * p.row_bytes, __float_as_uint(w));
  }
}

struct FwdCellsParams {
  const void *feat;
  const uint2 *entries;
  const int *cell_start;
  float *out;
  int C;
  int X, R;              // cells per row, rows per sample (Z*Y)
  int tiles_x, tiles_r;  // patches per sample
  int cells_per_sample;
  int layout;
  int B;
  FastDiv by_B, by_tiles_x;
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

// kQ channel quads per lane (quad j of lane l = l + j * kLanes); kLanes lanes per cell
// (C = 4 * kQ * kLanes), kLanes == 0: run time.  blockDim.x == 32 * lanes.
template <typename FeatT, int kLanes, int kQ>
__global__ void __launch_bounds__(kLanes ? 32 * kLanes : 1024, kLanes ? (kLanes <= 20 ? 2 : 1) : 1)
    k_fwd_cells(FwdCellsParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) float cells_ts[];  // [C][33] write-out tile (B_C_CELLS layout only)
  const int lanes = kLanes ? kLanes : p.C / (4 * kQ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  // Launch order: patches nearest the grid centre first, samples interleaved (point density peaks
  // around the ego vehicle, so the long cells start at once).  Any order is correct.
  const int t = (int)p.by_B.div(blockIdx.x);
  const int b = (int)blockIdx.x - t * p.B;
  const int t_r = (int)p.by_tiles_x.div((unsigned)t);
  const int x0 = cells_zigzag(t - t_r * p.tiles_x, p.tiles_x) * kCellsTileX;
  const int r0 = cells_zigzag(t_r, p.tiles_r) * kCellsTileY;
  const int nx = min(kCellsTileX, p.X - x0), nr = min(kCellsTileY, p.R - r0);
  const int cell_base = b * p.cells_per_sample;

  const int cell = tid / lanes, l = tid - cell * lanes;
  const int cty = cell / kCellsTileX, ctx = cell % kCellsTileX;
  const bool cell_ok = cty < nr && ctx < nx;
  const int gcell = cell_base + (r0 + cty) * p.X + x0 + ctx;
  int s = 0, e = 0;
  if (cell_ok) {
    s = __ldg(p.cell_start + gcell);
    e = __ldg(p.cell_start + gcell + 1);
  }
  const char *feat_q = static_cast<const char *>(p.feat) + (size_t)l * 4 * sizeof(FeatT);
  asm volatile("" : "+l"(feat_q));  // keep the per-lane base in a register pair (one IMAD.WIDE per row)
  const size_t qstep = (size_t)lanes * 4 * sizeof(FeatT);

  float4 acc[kQ];
#pragma unroll
  for (int j = 0; j < kQ; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  auto take = [&](const uint2 ev, float4 (&v)[kQ]) {  // issue the row loads of one entry
    if (ev.x != kSkipEntry) {
#pragma unroll
      for (int j = 0; j < kQ; ++j) v[j] = Row4<FeatT>::load_bytes(feat_q + ev.x + j * qstep);
    }
  };
  auto fold = [&](const uint2 ev, const float4 (&v)[kQ]) {
    if (ev.x != kSkipEntry) {
      const float w = __uint_as_float(ev.y);
#pragma unroll
      for (int j = 0; j < kQ; ++j) cells_fma(acc[j], v[j], w);
    }
  };

  int i = s;
  if ((i & 1) && i < e) {  // entries are fetched in 16-byte pairs: peel an odd first one
    const uint2 ev = __ldg(p.entries + i);
    float4 v[kQ];
    take(ev, v);
    fold(ev, v);
    ++i;
  }
  constexpr int kPairs = 2;  // 4 entries (rows) in flight per thread
  for (; i + 2 * kPairs <= e; i += 2 * kPairs) {
    uint4 pr[kPairs];
    float4 v[2 * kPairs][kQ];
#pragma unroll
    for (int u = 0; u < kPairs; ++u) pr[u] = __ldg(reinterpret_cast<const uint4 *>(p.entries + i) + u);
#pragma unroll
    for (int u = 0; u < kPairs; ++u) {
      take(make_uint2(pr[u].x, pr[u].y), v[2 * u]);
      take(make_uint2(pr[u].z, pr[u].w), v[2 * u + 1]);
    }
#pragma unroll
    for (int u = 0; u < kPairs; ++u) {
      fold(make_uint2(pr[u].x, pr[u].y), v[2 * u]);
      fold(make_uint2(pr[u].z, pr[u].w), v[2 * u + 1]);
    }
  }
  for (; i < e; ++i) {
    const uint2 ev = __ldg(p.entries + i);
    float4 v[kQ];
    take(ev, v);
    fold(ev, v);
  }

  // ---- write the whole patch, empty cells included ------------------------------------------------
  if (p.layout == RCB_LAYOUT_CELLS_C) {
    if (cell_ok) {
      float4 *dst = reinterpret_cast<float4 *>(p.out + (size_t)gcell * p.C);
#pragma unroll
      for (int j = 0; j < kQ; ++j) st_stream_f4(dst + l + j * lanes, acc[j]);
    }
    return;
  }
  // (B, C, cells): transpose through shared memory, then every store instruction writes kCellsTileY
  // runs of kCellsTileX consecutive cells of one channel row
#pragma unroll
  for (int j = 0; j < kQ; ++j) {
    const int c0 = 4 * (l + j * lanes);
    cells_ts[(c0 + 0) * 33 + cell] = acc[j].x, cells_ts[(c0 + 1) * 33 + cell] = acc[j].y;
    cells_ts[(c0 + 2) * 33 + cell] = acc[j].z, cells_ts[(c0 + 3) * 33 + cell] = acc[j].w;
  }
  __syncthreads();
  const int ty = lane / kCellsTileX, tx = lane % kCellsTileX;
  if (ty < nr && tx < nx) {
    float *dst = p.out + (size_t)b * p.C * p.cells_per_sample + (size_t)(r0 + ty) * p.X + x0 + tx;
    for (int ch = warp; ch < p.C; ch += n_warps) st_stream_f32(dst + (size_t)ch * p.cells_per_sample, cells_ts[ch * 33 + lane]);
  }
}

template <typename FeatT, int kLanes, int kQ>
static int launch_cells_t(const FwdCellsParams &p, int lanes, long long grid, cudaStream_t s) {
  const size_t smem = p.layout == RCB_LAYOUT_B_C_CELLS ? (size_t)p.C * 33 * 4 : 0;
  if (smem > 48 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(k_fwd_cells<FeatT, kLanes, kQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RCB_CUDA_TRY(launch_pdl(k_fwd_cells<FeatT, kLanes, kQ>, (unsigned)grid, 32 * lanes, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_cells(const FwdCellsParams &p, long long grid, cudaStream_t s) {
  const int C = p.C;
  if (C <= 128) {  // one quad per lane
    switch (C / 4) {
      case 16: return launch_cells_t<FeatT, 16, 1>(p, 16, grid, s);
      case 20: return launch_cells_t<FeatT, 20, 1>(p, 20, grid, s);
      case 32: return launch_cells_t<FeatT, 32, 1>(p, 32, grid, s);
      default: return launch_cells_t<FeatT, 0, 1>(p, C / 4, grid, s);
    }
  }
  return launch_cells_t<FeatT, 0, 2>(p, C / 8, grid, s);  // two quads per lane
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

size_t fwd_cells_workspace_bytes(const rcb_pool_desc *d) { return align_up((size_t)max(d->n_points, 1) * 8 + 16, 256); }

int fwd_cells_launch(const rcb_pool_desc *d, const float *depth, const void *feat, const int *ranks_depth,
                     const int *ranks_feat, const int *ranks_bev, const int *cell_start, float *out,
                     void *workspace, size_t workspace_bytes, int sms, cudaStream_t s) {
  if (!workspace || workspace_bytes < fwd_cells_workspace_bytes(d)) return RCB_ERR_WORKSPACE;
  if (((uintptr_t)workspace) % 16) return RCB_ERR_ALIGN;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  const int n_cells = d->B * d->Z * d->Y * d->X;
  FwdEntriesParams ep;
  ep.depth = depth, ep.ranks_depth = ranks_depth, ep.ranks_feat = ranks_feat, ep.ranks_bev = ranks_bev;
  ep.n_ptr = cell_start + n_cells;
  ep.entries = static_cast<uint2 *>(workspace);
  ep.row_bytes = (unsigned)d->C * elem;
  if (d->n_points > 0) {
    const int grid = max(1, min(ceil_div(d->n_points, 256), sms * 16));
    RCB_CUDA_TRY(launch_pdl(k_fwd_entries, grid, 256, 0, s, ep));
  }
  FwdCellsParams p;
  p.feat = feat, p.entries = ep.entries, p.cell_start = cell_start, p.out = out;
  p.C = d->C, p.X = d->X, p.R = d->Z * d->Y;
  p.tiles_x = ceil_div(p.X, kCellsTileX), p.tiles_r = ceil_div(p.R, kCellsTileY);
  p.cells_per_sample = d->Z * d->Y * d->X, p.layout = d->layout, p.B = d->B;
  p.by_B = FastDiv::make((unsigned)p.B), p.by_tiles_x = FastDiv::make((unsigned)p.tiles_x);
  const long long grid = (long long)d->B * p.tiles_r * p.tiles_x;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: return launch_cells<float>(p, grid, s);
    case RCB_DTYPE_BF16: return launch_cells<__nv_bfloat16>(p, grid, s);
    default: return launch_cells<__half>(p, grid, s);
  }
}

}  // namespace rcb
