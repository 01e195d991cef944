// Row F -- bev_pool_v2 forward: out[cell, c] = sum_i depth[ranks_depth[i]] * feat[ranks_feat[i], c]
// over the points i of the cell's interval.
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (kernel), :125-131 (launch),
// mmdet3d/ops/bev_pool_v2/bev_pool.py:16-41,86-92 (zero-fill before, permute copy after).
//
// Two kernels:
//
//  k_pool_fwd_tile  (the product path; needs the dense per-cell CSR `cell_start`, i.e. intervals
//      sorted by cell and tiling [0, K) -- what prepare emits and what rcb_pool_validate proves):
//      one CTA per TX x TY patch of BEV cells.  The patch's points are TY contiguous slices of the
//      sorted arrays: they are staged once, coalesced, into shared memory as (depth weight, feat
//      row) pairs -- the depth gather happens here with full memory-level parallelism.  The
//      patch's work is cut into items of <= L consecutive points of one cell; lane-groups of C/4
//      threads (one 128-bit load of the context row per thread) take items round-robin, so a
//      656-point cell near the ego vehicle is shared by many groups instead of serialising one.
//      Partial sums are combined in a fixed order (deterministic, no atomics).  Every cell of the
//      patch is written -- zeros included -- directly in the layout the caller wants:
//      channels-last (what bev_pool_v2_forward produces) or (B, C, Z*Y*X) (what bev_pool_v2()
//      returns after its permute+contiguous).  No memset pass, no permute pass.
//      A patch is 2-D so that neighbouring cells, which see the same camera pixels along a ray,
//      share context rows through L1: ~9x fewer L2 fetches than per-interval gathering.
//
//  k_pool_fwd_intervals (general path, any ranks the reference accepts): one warp per interval,
//      writes only non-empty cells into a pre-zeroed `out`.
#include <cstdlib>

#include "common.cuh"

namespace rcb {

constexpr int kTileX = 8;
constexpr int kTileY = 4;
constexpr int kTileCells = kTileX * kTileY;   // 32: one warp describes the patch, lane <-> cell
constexpr int kGroups = 16;                   // lane-groups (of C/4 threads) per CTA
constexpr int kStagePerThread = 4;            // points staged per thread and round
constexpr int kMinItem = 32;
constexpr int kMaxItems = kTileCells + kGroups;
constexpr int kFwdTableBytes = 1024;

struct FwdTileParams {
  const float *depth;
  const void *feat;
  const int *ranks_depth;
  const int *ranks_feat;
  const int *cell_start;
  float *out;
  int C, C4;
  int X, R;              // cells per row, rows per sample (Z*Y)
  int tiles_x, tiles_r;  // patches per sample
  int cells_per_sample;
  int layout;
  int B;
  FastDiv by_B, by_tiles_x;  // block index -> (sample, patch column, patch row) without IDIV
};

struct __align__(8) StagePoint {
  float w;
  unsigned row;  // ranks_feat: index of the context row
};

// i-th element of {c, c-1, c+1, c-2, c+2, ...} clipped to [0, n), c = n / 2
__device__ __forceinline__ int zigzag_from_centre(int i, int n) {
  const int c = n >> 1;
  const int lo_side = c, hi_side = n - 1 - c;           // elements below / above the centre
  const int paired = 2 * min(lo_side, hi_side) + 1;     // prefix that alternates
  if (i < paired) return (i & 1) ? c - ((i + 1) >> 1) : c + (i >> 1);
  const int rest = i - paired;                          // one side is exhausted
  return lo_side > hi_side ? c - hi_side - 1 - rest : c + lo_side + 1 + rest;
}

__host__ __device__ inline int fwd_part_pitch(int C) { return C + 4; }  // floats; +4 keeps 16-byte rows and
                                                                        // spreads cells over banks
__host__ __device__ inline size_t fwd_tile_smem_bytes(int C) {
  const size_t stage = (size_t)kStagePerThread * kGroups * (C / 4) * sizeof(StagePoint);
  const size_t part = (size_t)kMaxItems * fwd_part_pitch(C) * 4;
  return ((stage + part + 15) / 16) * 16 + kFwdTableBytes;
}

__device__ __forceinline__ void fma_row(float4 &acc, const float4 v, const float w) {
  const float2 ww = make_float2(w, w);
  float2 lo = __ffma2_rn(make_float2(v.x, v.y), ww, make_float2(acc.x, acc.y));
  float2 hi = __ffma2_rn(make_float2(v.z, v.w), ww, make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

#ifdef RCB_PROFILE_PHASES  // debug build only: per-CTA phase stamps read by tools/prof_fwd_phases.py
__device__ long long g_fwd_prof[16384 * 8];
__device__ __forceinline__ long long rcb_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int rcb_smid() {
  int v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
}
#define RCB_T(k) if (threadIdx.x == 0 && blockIdx.x < 16384) g_fwd_prof[blockIdx.x * 8 + (k)] = rcb_gtime();
#else
#define RCB_T(k)
#endif

// sum_i w_i * row_i over staged points [i, hi): fused multiply-adds in point order (the
// reference's order), kUnroll independent 128-bit row loads in flight
template <typename FeatT, int kUnroll>
__device__ __forceinline__ float4 accumulate_range(const StagePoint *__restrict__ stage,
                                                   const char *__restrict__ feat_q, unsigned row_stride,
                                                   int i, int hi) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i + kUnroll <= hi; i += kUnroll) {
    StagePoint sp[kUnroll];
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) sp[u] = stage[i + u];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      v[u] = Row4<FeatT>::load_bytes(feat_q + (size_t)sp[u].row * row_stride);  // one IMAD.WIDE
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) fma_row(acc, v[u], sp[u].w);
  }
  for (; i < hi; ++i) {
    const StagePoint s0 = stage[i];
    fma_row(acc, Row4<FeatT>::load_bytes(feat_q + (size_t)s0.row * row_stride), s0.w);
  }
  return acc;
}

// kC4 > 0: channels/4 known at compile time; 0: run time (C4 even).  blockDim.x == kGroups * C4,
// i.e. C4 / 2 warps, and every warp combines / writes two 128-bit channel quads of all 32 cells.
template <typename FeatT, int kC4>
#ifndef RCB_FWD_MINCTAS
#define RCB_FWD_MINCTAS 4  // measured: 3 -> 104 us, 4 -> 99 us, 5+ -> no gain (the row loads need the registers)
#endif
#ifndef RCB_FWD_UNROLL
#define RCB_FWD_UNROLL 6  // with the lane base pinned: 6 -> 96.4 us, 7 -> 97.0, 8 -> 99.1 (spills at 48 registers)
#endif
__global__ void __launch_bounds__(kC4 ? kC4 * kGroups : 1024, kC4 ? (kC4 <= 20 ? RCB_FWD_MINCTAS : 2) : 1)
    k_pool_fwd_tile(FwdTileParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C4 = kC4 ? kC4 : p.C4;
  const int C = C4 * 4;
  const int pitch = fwd_part_pitch(C);
  const int stage_cap = kStagePerThread * blockDim.x;
  StagePoint *stage = reinterpret_cast<StagePoint *>(smem_raw);
  float *part = reinterpret_cast<float *>(stage + stage_cap);
  int *tab = reinterpret_cast<int *>(smem_raw + fwd_tile_smem_bytes(C) - kFwdTableBytes);
  int *cell_lo = tab;          // [32] patch-local point range of every cell
  int *cell_hi = tab + 32;     // [32]
  int *seg_off = tab + 64;     // [kTileY + 1] patch-local prefix of the row slices
  int *seg_g = tab + 72;       // [kTileY] global start of each row slice
  int *cell_item0 = tab + 80;  // [33] first item of each cell in this round
  int *item_lo = tab + 116;    // [kMaxItems] staged range of each item
  int *item_hi = tab + 116 + kMaxItems;

  RCB_T(0)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
  // Launch order: patches nearest the grid centre first, samples interleaved.  Point density
  // peaks around the ego vehicle (a patch there holds ~8x the average), so the long patches start
  // at once instead of forming the kernel's tail.  Any order is correct; this one is a heuristic.
  const int t = (int)p.by_B.div(blockIdx.x);
  const int b = (int)blockIdx.x - t * p.B;
  const int t_r = (int)p.by_tiles_x.div((unsigned)t);
  const int tx_i = zigzag_from_centre(t - t_r * p.tiles_x, p.tiles_x);
  const int tr_i = zigzag_from_centre(t_r, p.tiles_r);
  const int x0 = tx_i * kTileX, r0 = tr_i * kTileY;
  const int nx = min(kTileX, p.X - x0), nr = min(kTileY, p.R - r0);
  const int cell_base = b * p.cells_per_sample;

  // ---- patch geometry, by warp 0: lane <-> cell ----------------------------------------------
  if (tid < 32) {
    const int ty = tid / kTileX, tx = tid % kTileX;
    int s = 0, e = 0;
    if (ty < nr && tx < nx) {
      const int c = cell_base + (r0 + ty) * p.X + x0 + tx;
      s = __ldg(p.cell_start + c);
      e = __ldg(p.cell_start + c + 1);
    }
    // row slice = [start of its first cell, end of its last valid cell)
    int off = 0, my_off = 0, my_g = 0;
#pragma unroll
    for (int r = 0; r < kTileY; ++r) {
      const int g = __shfl_sync(kFull, s, r * kTileX);
      const int ge = __shfl_sync(kFull, e, r * kTileX + nx - 1);
      const int len = r < nr ? ge - g : 0;
      if (r == ty) my_off = off, my_g = g;
      if (tid == r) seg_off[r] = off, seg_g[r] = g;
      off += len;
    }
    if (tid == 0) seg_off[kTileY] = off;
    cell_lo[tid] = s - my_g + my_off;
    cell_hi[tid] = e - my_g + my_off;
  }
  __syncthreads();
  RCB_T(1)
  const int total = seg_off[kTileY];
#ifdef RCB_PROFILE_PHASES
  if (threadIdx.x == 0 && blockIdx.x < 16384) g_fwd_prof[blockIdx.x * 8 + 7] = total, g_fwd_prof[blockIdx.x * 8 + 6] = rcb_smid();
#endif

  const int group = tid / C4, q = tid - group * C4;
  const char *feat_q = static_cast<const char *>(p.feat) + (size_t)q * 4 * sizeof(FeatT);
  // pin the full per-lane base in a register pair: otherwise the compiler keeps q * 16 and re-adds
  // the uniform tensor base with a 64-bit IADD3 pair in front of EVERY row load
  asm volatile("" : "+l"(feat_q));
  const unsigned row_stride = (unsigned)C * sizeof(FeatT);
  // combine / write role: lane <-> cell, this warp's quads are warp and warp + n_warps
  float4 racc[2];
  racc[0] = racc[1] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int cb = 0; cb < total; cb += stage_cap) {
    const int n = min(stage_cap, total - cb);
    // ---- stage (depth weight, context row offset): all index loads first, then the gathers ----
    {
      int g[kStagePerThread], rd[kStagePerThread];
      unsigned rf[kStagePerThread];
#pragma unroll
      for (int k = 0; k < kStagePerThread; ++k) {
        const int pt = cb + tid + k * blockDim.x;
        int ty = 0;
#pragma unroll
        for (int r = 1; r < kTileY; ++r) ty += (pt >= seg_off[r]);
        g[k] = tid + k * blockDim.x < n ? seg_g[ty] + (pt - seg_off[ty]) : -1;
      }
#pragma unroll
      for (int k = 0; k < kStagePerThread; ++k) {
        rd[k] = g[k] >= 0 ? ld_stream_s32(p.ranks_depth + g[k]) : 0;
        rf[k] = g[k] >= 0 ? (unsigned)ld_stream_s32(p.ranks_feat + g[k]) : 0u;
      }
#pragma unroll
      for (int k = 0; k < kStagePerThread; ++k) {
        if (g[k] >= 0) {
          StagePoint sp;
          sp.row = rf[k];
          sp.w = ld_stream_f32(p.depth + rd[k]);  // used once per cell: keep it out of L1
          stage[tid + k * blockDim.x] = sp;
        }
      }
    }
    // ---- work items: every cell is cut into equal pieces of <= L points -----------------------
    if (tid < 32) {
      const int a = max(cell_lo[tid], cb), bnd = min(cell_hi[tid], cb + n);
      // as many items as the part buffer holds (non-empty cells + extra cuts <= kMaxItems), but no
      // shorter than kMinItem points: a dense round (7 cells x 190 points) becomes ~40 pieces of 32
      // instead of 20 of 64, which the 16 lane-groups share far more evenly
      const int n_cells_here = __popc(__ballot_sync(kFull, bnd > a));
      const int L = max(kMinItem, ceil_div(n, max(kGroups, kMaxItems - n_cells_here)));
      const int mine = bnd > a ? ceil_div(bnd - a, L) : 0;
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (tid >= o) incl += v;
      }
      int off = incl - mine;
      cell_item0[tid] = off;
      if (tid == 31) cell_item0[32] = incl;
      // equal pieces (a 190-point cell with L = 80 becomes 64+63+63, not 80+80+30): the round ends
      // with its longest lane-group
      const int piece = mine > 0 ? ceil_div(bnd - a, mine) : 0;
      for (int k = 0; k < mine; ++k, ++off) {
        item_lo[off] = a + k * piece - cb;
        item_hi[off] = min(bnd, a + (k + 1) * piece) - cb;
      }
    }
    __syncthreads();
    if (cb == 0) { RCB_T(2) }
    // ---- items round-robin over the lane-groups ---------------------------------------------
    const int n_items = cell_item0[32];
    for (int it = group; it < n_items; it += kGroups) {
      const float4 acc = accumulate_range<FeatT, RCB_FWD_UNROLL>(stage, feat_q, row_stride, item_lo[it], item_hi[it]);
      *reinterpret_cast<float4 *>(part + (size_t)it * pitch + q * 4) = acc;
    }
    __syncthreads();
    if (cb == 0) { RCB_T(3) }
    // ---- fixed-order combine into the writer's registers -------------------------------------
    {
      const int i0 = cell_item0[lane], i1 = cell_item0[lane + 1];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int qq = warp + k * n_warps;
        for (int it = i0; it < i1; ++it) {
          const float4 v = *reinterpret_cast<const float4 *>(part + (size_t)it * pitch + qq * 4);
          racc[k].x += v.x, racc[k].y += v.y, racc[k].z += v.z, racc[k].w += v.w;
        }
      }
    }
    if (cb + stage_cap < total) __syncthreads();  // stage / part / tables are rewritten next round
  }

  RCB_T(4)
  // ---- write the whole patch, empty cells included: lane <-> cell ---------------------------
  const int ty = lane / kTileX, tx = lane % kTileX;
  if (ty >= nr || tx >= nx) return;
  const size_t cell_in_sample = (size_t)(r0 + ty) * p.X + x0 + tx;
  if (p.layout == RCB_LAYOUT_CELLS_C) {
    float4 *dst = reinterpret_cast<float4 *>(p.out + ((size_t)cell_base + cell_in_sample) * C);
#pragma unroll
    for (int k = 0; k < 2; ++k) st_stream_f4(dst + warp + k * n_warps, racc[k]);
  } else {
    // (B, C, cells): every store instruction writes kTileY runs of kTileX consecutive cells
    float *dst = p.out + (size_t)b * C * p.cells_per_sample + cell_in_sample;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float *d4 = dst + (size_t)(warp + k * n_warps) * 4 * p.cells_per_sample;
      st_stream_f32(d4, racc[k].x);
      st_stream_f32(d4 + p.cells_per_sample, racc[k].y);
      st_stream_f32(d4 + 2 * (size_t)p.cells_per_sample, racc[k].z);
      st_stream_f32(d4 + 3 * (size_t)p.cells_per_sample, racc[k].w);
    }
  }
  RCB_T(5)
}

// ---------------------------------------------------------------------------------------------
// General path: the reference's contract verbatim (any ranks, only interval cells written).
// ---------------------------------------------------------------------------------------------
template <typename FeatT>
__global__ void __launch_bounds__(256)
    k_pool_fwd_intervals(int n_intervals, int C, int cells_per_sample, int layout,
                         const float *__restrict__ depth, const FeatT *__restrict__ feat,
                         const int *__restrict__ ranks_depth, const int *__restrict__ ranks_feat,
                         const int *__restrict__ ranks_bev, const int *__restrict__ interval_starts,
                         const int *__restrict__ interval_lengths, float *__restrict__ out) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; iv < n_intervals; iv += warps) {
    const int start = __ldg(interval_starts + iv), len = __ldg(interval_lengths + iv);
    const int cell = __ldg(ranks_bev + start);
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      for (int base = 0; base < len; base += 32) {
        const int m = min(32, len - base);
        int row = 0;
        float w = 0.f;
        if (lane < m) {
          row = __ldg(ranks_feat + start + base + lane);
          w = __ldg(depth + __ldg(ranks_depth + start + base + lane));
        }
        for (int k = 0; k < m; ++k) {
          const int rk = __shfl_sync(kFull, row, k);
          const float wk = __shfl_sync(kFull, w, k);
          if (c < C) acc = fmaf(to_f32<FeatT>(feat[(size_t)rk * C + c]), wk, acc);
        }
      }
      if (c < C) {
        if (layout == RCB_LAYOUT_B_C_CELLS) {
          const int bb = cell / cells_per_sample, cc = cell - bb * cells_per_sample;
          out[((size_t)bb * C + c) * cells_per_sample + cc] = acc;
        } else {
          out[(size_t)cell * C + c] = acc;
        }
      }
    }
  }
}

template <typename FeatT, int kC4>
static int launch_tile_c4(const rcb_pool_desc *d, FwdTileParams &p, cudaStream_t s) {
  const int threads = kGroups * p.C4;
  const size_t smem = fwd_tile_smem_bytes(p.C);
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_tile<FeatT, kC4>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)d->B * p.tiles_r * p.tiles_x;
  RCB_CUDA_TRY(launch_pdl(k_pool_fwd_tile<FeatT, kC4>, (unsigned)grid, threads, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_tile(const rcb_pool_desc *d, FwdTileParams &p, cudaStream_t s) {
  switch (p.C4) {
    case 16: return launch_tile_c4<FeatT, 16>(d, p, s);
    case 20: return launch_tile_c4<FeatT, 20>(d, p, s);
    case 32: return launch_tile_c4<FeatT, 32>(d, p, s);
    default: return launch_tile_c4<FeatT, 0>(d, p, s);
  }
}

template <typename FeatT>
static int launch_intervals(const rcb_pool_desc *d, const float *depth, const void *feat,
                            const int *rd, const int *rf, const int *rb, const int *starts,
                            const int *lengths, float *out, int sms, cudaStream_t s) {
  const int grid = max(1, min(ceil_div(d->n_intervals, 8), sms * 16));
  k_pool_fwd_intervals<FeatT><<<grid, 256, 0, s>>>(d->n_intervals, d->C, d->Z * d->Y * d->X, d->layout,
                                                   depth, static_cast<const FeatT *>(feat), rd, rf, rb,
                                                   starts, lengths, out);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

bool fwd_cells_eligible(const rcb_pool_desc *d, const void *feat, const int *cell_start);
size_t fwd_cells_workspace_bytes(const rcb_pool_desc *d);
int fwd_cells_launch(const rcb_pool_desc *d, const float *depth, const void *feat, const int *ranks_depth,
                     const int *ranks_feat, const int *ranks_bev, const int *cell_start, float *out,
                     void *workspace, size_t workspace_bytes, int sms, cudaStream_t s);

// Diagnostic knob (A/B timing only): RCB_FWD_KERNEL=tile keeps round 1's lane-group kernel.
static bool fwd_cells_disabled() {
  static const bool off = [] {
    const char *v = getenv("RCB_FWD_KERNEL");
    return v != nullptr && v[0] == 't';
  }();
  return off;
}

int check_pool_desc(const rcb_pool_desc *d) {
  if (!d) return RCB_ERR_ARG;
  if (d->n_points < 0 || d->n_intervals < 0 || d->C <= 0 || d->B <= 0 || d->Z <= 0 || d->Y <= 0 ||
      d->X <= 0 || d->n_depth < 0 || d->n_pixels < 0)
    return RCB_ERR_ARG;
  if ((long long)d->B * d->Z * d->Y * d->X >= (1ll << 31) / 4) return RCB_ERR_UNSUPPORTED;
  if (d->layout != RCB_LAYOUT_CELLS_C && d->layout != RCB_LAYOUT_B_C_CELLS) return RCB_ERR_ARG;
  if (d->feat_dtype != RCB_DTYPE_F32 && d->feat_dtype != RCB_DTYPE_BF16 && d->feat_dtype != RCB_DTYPE_F16)
    return RCB_ERR_ARG;
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

extern "C" size_t rcb_pool_fwd_workspace_bytes(const rcb_pool_desc *d) {
  if (check_pool_desc(d) != RCB_OK) return 0;
  return fwd_cells_workspace_bytes(d);
}

extern "C" int rcb_bev_pool_v2_fwd(const rcb_pool_desc *d, const float *depth, const void *feat,
                                   const int *ranks_depth, const int *ranks_feat,
                                   const int *ranks_bev, const int *interval_lengths,
                                   const int *interval_starts, const int *cell_start, float *out,
                                   void *workspace, size_t workspace_bytes, int device,
                                   rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!out) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  const int cps = d->Z * d->Y * d->X;
  const size_t out_bytes = (size_t)d->B * cps * d->C * 4;

  if (!fwd_cells_disabled() && fwd_cells_eligible(d, feat, cell_start)) {
    if (d->n_points > 0 && (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev)) return RCB_ERR_ARG;
    return fwd_cells_launch(d, depth, feat, ranks_depth, ranks_feat, ranks_bev, cell_start, out, workspace,
                            workspace_bytes, sm_count_cached(device), s);
  }
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  const bool tile_ok = cell_start != nullptr && (d->C % 8) == 0 && d->C <= 256 && (long long)d->n_pixels * d->C * elem < (1ll << 32) &&
                       (((uintptr_t)feat) % (4 * elem)) == 0;
  if (tile_ok) {
    if (d->n_points > 0 && (!depth || !feat || !ranks_depth || !ranks_feat)) return RCB_ERR_ARG;
    FwdTileParams p;
    p.depth = depth, p.feat = feat, p.ranks_depth = ranks_depth, p.ranks_feat = ranks_feat;
    p.cell_start = cell_start, p.out = out;
    p.C = d->C, p.C4 = d->C / 4;
    p.X = d->X, p.R = d->Z * d->Y;
    p.tiles_x = ceil_div(p.X, kTileX), p.tiles_r = ceil_div(p.R, kTileY);
    p.cells_per_sample = cps, p.layout = d->layout, p.B = d->B;
    p.by_B = FastDiv::make((unsigned)p.B), p.by_tiles_x = FastDiv::make((unsigned)p.tiles_x);
    switch (d->feat_dtype) {
      case RCB_DTYPE_F32: return launch_tile<float>(d, p, s);
      case RCB_DTYPE_BF16: return launch_tile<__nv_bfloat16>(d, p, s);
      default: return launch_tile<__half>(d, p, s);
    }
  }
  // A CSR without interval arrays is the sync-free fused chain (n_points is only an upper bound
  // there): it cannot fall back to the interval kernel, so say so instead of returning zeros.
  if (cell_start != nullptr && d->n_points > 0 && (!interval_lengths || !interval_starts)) return RCB_ERR_UNSUPPORTED;
  // general path: zero-fill (bev_pool.py:27) then one warp per interval
  RCB_CUDA_TRY(cudaMemsetAsync(out, 0, out_bytes, s));
  if (d->n_intervals == 0) return RCB_OK;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !interval_lengths ||
      !interval_starts)
    return RCB_ERR_ARG;
  const int sms = sm_count_cached(device);
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32:
      return launch_intervals<float>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                     interval_starts, interval_lengths, out, sms, s);
    case RCB_DTYPE_BF16:
      return launch_intervals<__nv_bfloat16>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                             interval_starts, interval_lengths, out, sms, s);
    default:
      return launch_intervals<__half>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                      interval_starts, interval_lengths, out, sms, s);
  }
}

#ifdef RCB_PROFILE_PHASES
extern "C" int rcb_debug_fwd_prof(long long *host, int n) {
  return (int)cudaMemcpyFromSymbol(host, rcb::g_fwd_prof, sizeof(long long) * n);
}
extern "C" int rcb_debug_fwd_prof_reset() {
  void *p = nullptr;
  cudaGetSymbolAddress(&p, rcb::g_fwd_prof);
  return (int)cudaMemset(p, 0, sizeof(long long) * 16384 * 8);
}
#endif
