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
#include "common.cuh"

namespace rcb {

constexpr int kTileX = 8;
constexpr int kTileY = 8;
constexpr int kTileCells = kTileX * kTileY;
constexpr int kStageCap = 2048;  // points staged per round
constexpr int kMinItem = 32;
constexpr int kFwdMinCtas = 2;

struct FwdTileParams {
  const float *depth;
  const void *feat;
  const int *ranks_depth;
  const int *ranks_feat;
  const int *cell_start;
  float *out;
  int C, C4;
  int n_groups;          // lane-groups per CTA = blockDim.x / C4
  int X, R;              // cells per row, rows per sample (Z*Y)
  int tiles_x, tiles_r;  // patches per sample
  int cells_per_sample;
  int layout;
};

struct StagePoint {
  float w;
  int row;
};

__host__ __device__ inline size_t fwd_tile_smem_bytes(int C, int n_groups) {
  const int max_items = n_groups + kTileCells;
  size_t b = 0;
  b += (size_t)kStageCap * sizeof(StagePoint);
  b += (size_t)max_items * C * 4;          // part
  b += (size_t)kTileCells * (C + 1) * 4;   // res
  b += (size_t)max_items * 3 * 4;          // item_cell/lo/hi
  b += (size_t)kTileCells * 4 * 4;         // cell_lo, cell_hi, cell_item0, cell_items
  b += 64 * 4;                             // seg tables + scalars
  return b;
}

// kC4 > 0: channels/4 known at compile time (index arithmetic by constants); 0: runtime.
template <typename FeatT, int kC4>
__global__ void __launch_bounds__(kC4 ? kC4 * 32 : 1024, (kC4 && kC4 * 32 <= 640) ? kFwdMinCtas : 1)
    k_pool_fwd_tile(FwdTileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C4 = kC4 ? kC4 : p.C4;
  const int C = C4 * 4, NG = p.n_groups;
  const int max_items = NG + kTileCells;
  StagePoint *stage = reinterpret_cast<StagePoint *>(smem_raw);
  float *part = reinterpret_cast<float *>(stage + kStageCap);
  float *res = part + (size_t)max_items * C;
  int *item_cell = reinterpret_cast<int *>(res + kTileCells * (C + 1));
  int *item_lo = item_cell + max_items;
  int *item_hi = item_lo + max_items;
  int *cell_lo = item_hi + max_items;
  int *cell_hi = cell_lo + kTileCells;
  int *cell_item0 = cell_hi + kTileCells;
  int *cell_items = cell_item0 + kTileCells;
  int *seg_off = cell_items + kTileCells;   // [kTileY + 1] patch-local prefix of row slices
  int *seg_g = seg_off + kTileY + 1;        // [kTileY] global start of each row slice
  int *s_nitems = seg_g + kTileY;

  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx_i = t % p.tiles_x;
  t /= p.tiles_x;
  const int tr_i = t % p.tiles_r;
  const int b = t / p.tiles_r;
  const int x0 = tx_i * kTileX, r0 = tr_i * kTileY;
  const int nx = min(kTileX, p.X - x0), nr = min(kTileY, p.R - r0);
  const int cell_base = b * p.cells_per_sample;

  // ---- patch geometry: one thread per cell reads its CSR range -----------------------------
  if (tid < kTileCells) {
    const int ty = tid / kTileX, tx = tid % kTileX;
    int s = 0, e = 0;
    if (ty < nr && tx < nx) {
      const int c = cell_base + (r0 + ty) * p.X + x0 + tx;
      s = __ldg(p.cell_start + c);
      e = __ldg(p.cell_start + c + 1);
    }
    cell_lo[tid] = s;  // global for now
    cell_hi[tid] = e;
  }
  __syncthreads();
  if (tid == 0) {
    int off = 0;
    for (int ty = 0; ty < kTileY; ++ty) {
      seg_off[ty] = off;
      int g = 0, n = 0;
      if (ty < nr) {
        g = cell_lo[ty * kTileX];
        n = cell_hi[ty * kTileX + nx - 1] - g;
      }
      seg_g[ty] = g;
      off += n;
    }
    seg_off[kTileY] = off;
  }
  __syncthreads();
  const int total = seg_off[kTileY];
  if (tid < kTileCells) {  // global -> patch-local coordinates
    const int ty = tid / kTileX;
    const int shift = seg_off[ty] - seg_g[ty];
    cell_lo[tid] += shift;
    cell_hi[tid] += shift;
  }
  for (int i = tid; i < kTileCells * (C + 1); i += blockDim.x) res[i] = 0.f;
  __syncthreads();

  const int group = tid / C4, q = tid - group * C4;
  const FeatT *feat = static_cast<const FeatT *>(p.feat);

  for (int cb = 0; cb < total; cb += kStageCap) {
    const int n = min(kStageCap, total - cb);
    // ---- stage (weight, row) of the round's points; build the item list ----------------------
    for (int i = tid; i < n; i += blockDim.x) {
      const int pt = cb + i;
      int ty = 0;
#pragma unroll
      for (int k = 1; k < kTileY; ++k) ty += (pt >= seg_off[k]);
      const int g = seg_g[ty] + (pt - seg_off[ty]);
      const int rd = ld_stream_s32(p.ranks_depth + g);
      StagePoint sp;
      sp.row = ld_stream_s32(p.ranks_feat + g);
      sp.w = __ldg(p.depth + rd);
      stage[i] = sp;
    }
    const int L = max(kMinItem, ceil_div(n, NG));
    if (tid < 32) {  // warp 0: items per cell (2 cells per lane), exclusive scan, emit
      int a[2], bnd[2], cnt[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int j = tid * 2 + k;
        a[k] = max(cell_lo[j], cb);
        bnd[k] = min(cell_hi[j], cb + n);
        cnt[k] = bnd[k] > a[k] ? ceil_div(bnd[k] - a[k], L) : 0;
      }
      const int mine = cnt[0] + cnt[1];
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, o);
        if (tid >= o) incl += v;
      }
      int off = incl - mine;
      if (tid == 31) *s_nitems = incl;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int j = tid * 2 + k;
        cell_item0[j] = off;
        cell_items[j] = cnt[k];
        for (int s = 0; s < cnt[k]; ++s) {
          item_cell[off] = j;
          item_lo[off] = a[k] + s * L - cb;
          item_hi[off] = min(bnd[k], a[k] + (s + 1) * L) - cb;
          ++off;
        }
      }
    }
    __syncthreads();
    // ---- items: sequential fused multiply-adds in point order (the reference's order) --------
    const int n_items = *s_nitems;
    if (group < NG) {
      for (int it = group; it < n_items; it += NG) {
        int i = item_lo[it];
        const int hi = item_hi[it];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (; i + 4 <= hi; i += 4) {
          const StagePoint s0 = stage[i], s1 = stage[i + 1], s2 = stage[i + 2], s3 = stage[i + 3];
          const float4 v0 = Row4<FeatT>::load(feat, (size_t)s0.row * C4 + q);
          const float4 v1 = Row4<FeatT>::load(feat, (size_t)s1.row * C4 + q);
          const float4 v2 = Row4<FeatT>::load(feat, (size_t)s2.row * C4 + q);
          const float4 v3 = Row4<FeatT>::load(feat, (size_t)s3.row * C4 + q);
          acc.x = fmaf(v0.x, s0.w, acc.x), acc.y = fmaf(v0.y, s0.w, acc.y);
          acc.z = fmaf(v0.z, s0.w, acc.z), acc.w = fmaf(v0.w, s0.w, acc.w);
          acc.x = fmaf(v1.x, s1.w, acc.x), acc.y = fmaf(v1.y, s1.w, acc.y);
          acc.z = fmaf(v1.z, s1.w, acc.z), acc.w = fmaf(v1.w, s1.w, acc.w);
          acc.x = fmaf(v2.x, s2.w, acc.x), acc.y = fmaf(v2.y, s2.w, acc.y);
          acc.z = fmaf(v2.z, s2.w, acc.z), acc.w = fmaf(v2.w, s2.w, acc.w);
          acc.x = fmaf(v3.x, s3.w, acc.x), acc.y = fmaf(v3.y, s3.w, acc.y);
          acc.z = fmaf(v3.z, s3.w, acc.z), acc.w = fmaf(v3.w, s3.w, acc.w);
        }
        for (; i < hi; ++i) {
          const StagePoint s0 = stage[i];
          const float4 v0 = Row4<FeatT>::load(feat, (size_t)s0.row * C4 + q);
          acc.x = fmaf(v0.x, s0.w, acc.x), acc.y = fmaf(v0.y, s0.w, acc.y);
          acc.z = fmaf(v0.z, s0.w, acc.z), acc.w = fmaf(v0.w, s0.w, acc.w);
        }
        *reinterpret_cast<float4 *>(part + (size_t)it * C + q * 4) = acc;
      }
    }
    __syncthreads();
    // ---- fixed-order combine of each cell's items ---------------------------------------------
    for (int idx = tid; idx < kTileCells * C4; idx += blockDim.x) {
      const int j = idx / C4, qq = idx - j * C4;
      const int k0 = cell_item0[j], kn = cell_items[j];
      if (kn > 0) {
        float4 s = *reinterpret_cast<const float4 *>(part + (size_t)k0 * C + qq * 4);
        for (int k = 1; k < kn; ++k) {
          const float4 v = *reinterpret_cast<const float4 *>(part + (size_t)(k0 + k) * C + qq * 4);
          s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
        }
        float *r = res + j * (C + 1) + qq * 4;
        r[0] += s.x, r[1] += s.y, r[2] += s.z, r[3] += s.w;
      }
    }
    __syncthreads();
  }

  // ---- write the whole patch, empty cells included ------------------------------------------
  if (p.layout == RCB_LAYOUT_B_C_CELLS) {
    const int per_c = nr * nx;
    float *out_b = p.out + (size_t)b * C * p.cells_per_sample + (size_t)r0 * p.X + x0;
    for (int idx = tid; idx < C * per_c; idx += blockDim.x) {
      const int c = idx / per_c, rem = idx - c * per_c;
      const int ty = rem / nx, tx = rem - ty * nx;
      st_stream_f32(out_b + (size_t)c * p.cells_per_sample + ty * p.X + tx,
                    res[(ty * kTileX + tx) * (C + 1) + c]);
    }
  } else {
    const int per_row = nx * C;
    for (int idx = tid; idx < nr * per_row; idx += blockDim.x) {
      const int ty = idx / per_row, rem = idx - ty * per_row;
      const int tx = rem / C, c = rem - tx * C;
      st_stream_f32(p.out + ((size_t)cell_base + (size_t)(r0 + ty) * p.X + x0 + tx) * C + c,
                    res[(ty * kTileX + tx) * (C + 1) + c]);
    }
  }
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
  const int threads = p.n_groups * p.C4;
  const size_t smem = fwd_tile_smem_bytes(p.C, p.n_groups);
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_tile<FeatT, kC4>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long grid = (long long)d->B * p.tiles_r * p.tiles_x;
  k_pool_fwd_tile<FeatT, kC4><<<(unsigned)grid, threads, smem, s>>>(p);
  RCB_LAUNCH_CHECK();
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

extern "C" int rcb_bev_pool_v2_fwd(const rcb_pool_desc *d, const float *depth, const void *feat,
                                   const int *ranks_depth, const int *ranks_feat,
                                   const int *ranks_bev, const int *interval_lengths,
                                   const int *interval_starts, const int *cell_start, float *out,
                                   int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!out) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  const int cps = d->Z * d->Y * d->X;
  const size_t out_bytes = (size_t)d->B * cps * d->C * 4;

  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  const bool tile_ok = cell_start != nullptr && (d->C % 4) == 0 && d->C <= 256 &&
                       (((uintptr_t)feat) % (4 * elem)) == 0;
  if (tile_ok) {
    if (d->n_points > 0 && (!depth || !feat || !ranks_depth || !ranks_feat)) return RCB_ERR_ARG;
    FwdTileParams p;
    p.depth = depth, p.feat = feat, p.ranks_depth = ranks_depth, p.ranks_feat = ranks_feat;
    p.cell_start = cell_start, p.out = out;
    p.C = d->C, p.C4 = d->C / 4;
    p.n_groups = max(1, min(32, 1024 / p.C4));
    p.X = d->X, p.R = d->Z * d->Y;
    p.tiles_x = ceil_div(p.X, kTileX), p.tiles_r = ceil_div(p.R, kTileY);
    p.cells_per_sample = cps, p.layout = d->layout;
    switch (d->feat_dtype) {
      case RCB_DTYPE_F32: return launch_tile<float>(d, p, s);
      case RCB_DTYPE_BF16: return launch_tile<__nv_bfloat16>(d, p, s);
      default: return launch_tile<__half>(d, p, s);
    }
  }
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
