// Row R -- PointPillarsScatterRCS.forward up to (not including) its two convolutions.
// Reference: mmdet3d/middle_encoders/pillar_scatter.py:64-104 (pillar scatter), :115-131 (RCS
// heat-maps: a Python loop over pillars with two .item() syncs and a numpy Gaussian each),
// mmdet3d/core/utils/gaussian.py:6-23 (gaussian_2d, float64), :26-55 (max-draw), :57-81 (fill).
//
// The sequential loop has two order-dependent effects, both expressible as order-free maxima:
//   heatmap[cell]      = max over pillars covering the cell of G_r(dx, dy)        (max is commutative)
//   heatmap_feat[cell] = rcs value of the LAST pillar (largest index) covering the cell
// so the loop becomes one splat kernel with integer atomicMax (non-negative floats order like
// their bit patterns) -- deterministic -- followed by one dense write kernel that produces all
// three outputs, zeros included, with coalesced rows (no memset of the outputs, no per-pillar
// host sync).
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace rcb {

struct RadarParams {
  int V, Cin, rcs_dim, B, ny, nx, cells;  // cells = ny*nx
};

// pillar_scatter.py:122-126,130: r = x^2 + y^2; relu(rcs * r) + 1 in fp32, then Python int().
__device__ __forceinline__ int rcs_radius(const float *rcs_row, int rcs_dim) {
  const float x = rcs_row[0], y = rcs_row[1];
  const float r = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const float t = fmaxf(__fmul_rn(rcs_row[rcs_dim - 2], r), 0.f);
  const float rad = __fadd_rn(t, 1.f);
  if (!(rad < 2147483520.f)) return 0x7fffff00;
  return (int)rad;
}

// ------------------------------------------------------------------------------------------------
// Tile (gather) form of the two maxima -- the default.  Both reductions are maxima over the pillars
// whose window covers a cell, so a CELL can own them: a CTA takes a tile of 32 x 8R cells of one
// sample, a thread R cells of one column, and the maxima live in registers -- no atomics, no L2
// round trip per update, no memset, and heatmap / heatmap_feat / pillar_at leave as coalesced rows.
//   k_radar_records : per pillar {x, y, radius, sample} (radius -1 = coordinates outside the grid)
//                     + per sample the index range its pillars span (integer atomicMax: order-free)
//   k_radar_tiles<R>: the CTA scans its sample's range of records in chunks of 1024, compacts the
//                     pillars whose window meets the tile into shared memory (list order is
//                     irrelevant: maxima), and every thread walks the list: |dx| <= r and |dy| <= r
//                     is the clipped window of gaussian.py:40-47; G = float32(E_r[|dx|] * E_r[|dy|]),
//                     the float64 product of the factors the splat kernel tabulates per pillar, comes
//                     from a table over (radius, |dy|, |dx|) filled once per device (radii >= kTabR:
//                     two exponentials, out of line).
// Work: (cells x pillars near the tile) integer tests + one float64 product per covered cell, against
// two L2 loads + up to two atomics per covered cell in the splat form (kept below for reference
// measurements: RCB_RADAR_SPLAT=1).
// ------------------------------------------------------------------------------------------------
constexpr int kTabR = 64;  // tabulated radii / offsets: G[r][|dy|][|dx|], r, |dx|, |dy| < kTabR
// float32(E_r[|dy|] * E_r[|dx|]) with the eps cut of gaussian.py:21 applied -- what a covered cell takes
// the maximum of -- and EXACTLY ZERO outside the window (|dx| > r or |dy| > r).  Inside the window the
// value is never zero (its minimum, at the corners, is exp(-36 r^2 / (2r+1)^2) > exp(-9)), so
// "value > 0" IS the window test: the hot loop needs no comparison against the radius.  1 MB, L2
// resident, the part within reach of the common radii L1 resident.
constexpr int kTabW = 128;                       // offsets per table row / rows per radius: every offset a tile can see
constexpr int kTabStride = kTabW * kTabW;        // (|dx| <= r + 31, |dy| <= r + 8 R - 1 <= 126 for r < 64) has an entry, zero outside the window
__device__ float g_radar_tab[kTabR * kTabStride];

__device__ __forceinline__ double radar_denom(int radius) {
  // gaussian.py:17-23 with sigma = diameter / 6 (gaussian.py:38-39), all float64
  const double diameter = 2.0 * (double)radius + 1.0;
  const double sigma = diameter / 6.0;
  return 2.0 * sigma * sigma;
}
__device__ __forceinline__ double radar_factor(int k, double denom) {
  return exp(-(double)((long long)k * k) / denom);
}
__device__ __forceinline__ float radar_value(double ex, double ey) {
  double g = ex * ey;
  if (g < 2.220446049250313e-16) g = 0.0;  // np.finfo(float64).eps * h.max(), h.max() == 1
  return (float)g;
}
// radii beyond the table (rare): kept out of line so that the float64 division and exponentials stay
// off the tile kernel's hot loop
__device__ __noinline__ float radar_value_big(int ax, int ay, int radius) {
  const double denom = radar_denom(radius);
  return radar_value(radar_factor(ax, denom), radar_factor(ay, denom));
}

__global__ void __launch_bounds__(256) k_radar_table() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTabR * kTabStride) return;
  const int r = i / kTabStride, o = i % kTabStride, ay = o / kTabW, ax = o % kTabW;
  const double denom = radar_denom(r);
  g_radar_tab[i] = (ax > r || ay > r) ? 0.f : radar_value(radar_factor(ax, denom), radar_factor(ay, denom));  // o == 4096: ay = 64 > r
}

// range[2b] = max(V - v), range[2b + 1] = max(v + 1) over the valid pillars of sample b (zeroed before)
__global__ void __launch_bounds__(256) k_radar_records(RadarParams p, const float *__restrict__ rcs,
                                                       const int *__restrict__ coors, int4 *__restrict__ rec,
                                                       int *__restrict__ range) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  int4 q = make_int4(0, 0, -1, -1);
  if (v < p.V) {
    const int4 c = __ldg(reinterpret_cast<const int4 *>(coors) + v);  // [b, z, y, x]
    if (c.x >= 0 && c.x < p.B && c.z >= 0 && c.z < p.ny && c.w >= 0 && c.w < p.nx)
      q = make_int4(c.w, c.z, rcs_radius(rcs + (size_t)v * p.rcs_dim, p.rcs_dim), c.x);
    rec[v] = q;
  }
  // one pair of atomics per run of equal samples inside the warp (pillars arrive grouped by sample:
  // usually one run); lanes without a valid pillar carry sample -1
  const int lane = lane_id();
  const int prev = __shfl_up_sync(kFull, q.w, 1);
  const unsigned heads = __ballot_sync(kFull, lane == 0 || prev != q.w);
  if ((heads >> lane) & 1u) {
    const unsigned after = heads & ~((2u << lane) - 1u);          // heads above this lane
    const int run_last = after ? __ffs(after) - 2 : 31;           // last lane of this run
    if (q.w >= 0) {
      atomicMax(range + 2 * q.w, p.V - v);                        // v ascends with the lane: the head has the smallest
      atomicMax(range + 2 * q.w + 1, v + (run_last - lane) + 1);  // ... and the run's last lane the largest
    }
  }
}

constexpr int kScanPerThread = 4;                    // records a thread inspects per round
constexpr int kScanChunk = 256 * kScanPerThread;     // = capacity of the CTA's pillar lists

template <int R>
__global__ void __launch_bounds__(256) k_radar_tiles(RadarParams p, const float *__restrict__ rcs,
                                                     const int4 *__restrict__ rec, const int *__restrict__ range,
                                                     int *__restrict__ pillar_at, float *__restrict__ heatmap,
                                                     float *__restrict__ heatmap_feat, int tiles_x, int tiles_y,
                                                     const float *__restrict__ table,
                                                     const float *__restrict__ point_features,
                                                     float *__restrict__ features) {
  pdl_prologue();  // launched with programmatic stream serialisation behind k_radar_records
  __shared__ int4 s_list[kScanChunk];
  __shared__ int s_own[R * 256];  // pillar at each cell of the tile (index + 1), thread-major like the registers
  __shared__ int s_n, s_nbig;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
  const int x0 = (t % tiles_x) * 32, y0 = (t / tiles_x) * (8 * R);
  const int x1 = min(x0 + 31, p.nx - 1), y1 = min(y0 + 8 * R - 1, p.ny - 1);
  const int v_lo = p.V - range[2 * b], v_hi = range[2 * b + 1];  // empty sample: v_lo = V, v_hi = 0
  const int cx = x0 + lane, cy = y0 + warp;                      // cell j of the thread: (cx, cy + 8j)
  const int cy64 = cy * kTabW;   // (row offsets are pre-scaled by the table's row pitch)
  float best[R];
  int last[R];
#pragma unroll
  for (int j = 0; j < R; ++j) best[j] = 0.f, last[j] = 0, s_own[j * 256 + tid] = 0;

  for (int v0 = v_lo; v0 < v_hi; v0 += kScanChunk) {
    __syncthreads();  // the previous chunk's list is no longer read (first trip: s_own is zeroed)
    if (tid == 0) s_n = 0, s_nbig = 0;
    __syncthreads();
    int4 q[kScanPerThread];
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
      const int v = v0 + k * 256 + tid;
      q[k] = v < v_hi ? __ldg(rec + v) : make_int4(0, 0, -1, -1);
    }
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
      // the window [x - r, x + r] x [y - r, y + r] meets the tile
      const int ddx = q[k].x - min(max(q[k].x, x0), x1), ddy = q[k].y - min(max(q[k].y, y0), y1);
      if (q[k].w == b && q[k].z >= 0 && abs(ddx) <= q[k].z && abs(ddy) <= q[k].z) {
        const int v1 = v0 + k * 256 + tid + 1;
        // small radii fill the list from the front, radii beyond the table from the back
        if (q[k].z < kTabR) s_list[atomicAdd(&s_n, 1)] = make_int4(q[k].x, q[k].y * kTabW, q[k].z * kTabStride, v1);
        else s_list[kScanChunk - 1 - atomicAdd(&s_nbig, 1)] = make_int4(q[k].x, q[k].y, q[k].z, v1);
        if ((ddx | ddy) == 0) {  // the pillar's own cell is in this tile
          const int ly = q[k].y - y0;
          atomicMax(&s_own[(ly >> 3) * 256 + (ly & 7) * 32 + (q[k].x - x0)], v1);
        }
      }
    }
    __syncthreads();
    const int n = s_n;
    for (int i = 0; i < n; ++i) {
      const int4 e = s_list[i];                      // x, y * kTabW, radius * kTabStride, index + 1
      const unsigned ax = __sad(cx, e.x, (unsigned)e.z);   // |dx| + the radius' table base
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const unsigned idx = __sad(cy64 + 8 * kTabW * j, e.y, ax);   // base + |dy| * kTabW + |dx|
        // every pillar of the list meets the tile: |dx| <= r + 31, |dy| <= r + 8 R - 1 -- inside the table
        const float g = __ldg(table + idx);
        best[j] = fmaxf(best[j], g);
        if (g > 0.f) last[j] = max(last[j], e.w);    // inside the window (see g_radar_tab)
      }
    }
    const int nbig = s_nbig;
    for (int i = 0; i < nbig; ++i) {
      const int4 e = s_list[kScanChunk - 1 - i];
      const int ax = abs(cx - e.x);
      if (ax > e.z) continue;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int ay = abs(cy + 8 * j - e.y);
        if (ay > e.z) continue;
        best[j] = fmaxf(best[j], radar_value_big(ax, ay, e.z));
        last[j] = max(last[j], e.w);
      }
    }
  }
  __syncthreads();  // (no trip at all: s_own zeroed by its own thread, no barrier needed, but harmless)
  if (cx >= p.nx) return;
  const int col = p.rcs_dim - 2;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int y = cy + 8 * j;
    if (y >= p.ny) break;
    const size_t g = (size_t)b * p.cells + (size_t)y * p.nx + cx;
    const int own = s_own[j * 256 + tid];
    if (features != nullptr) {
      // small grids: the dense feature planes leave from here too (a warp writes 128-byte rows of one
      // channel; the pillar's row of point_features is 2 cache lines, read once and served from L1)
      float *dst = features + (size_t)b * p.Cin * p.cells + (size_t)y * p.nx + cx;
      const float *src = point_features + (size_t)(own > 0 ? own - 1 : 0) * p.Cin;
#pragma unroll 8
      for (int c = 0; c < p.Cin; ++c) st_stream_f32(dst + (size_t)c * p.cells, own > 0 ? __ldg(src + c) : 0.f);
    } else {
      pillar_at[g] = own;
    }
    st_stream_f32(heatmap + g, best[j]);
    st_stream_f32(heatmap_feat + g, last[j] > 0 ? __ldg(rcs + (size_t)(last[j] - 1) * p.rcs_dim + col) : 0.f);
  }
}

// Splat: one CTA per group of kSplatPillars pillars; for each pillar the ROWS of its window are dealt
// to the CTA's eight warps.  A pillar's window is (2r+1)^2 cells with r in [1, 56] on the config-4
// distribution (780 cells on average): a warp per pillar (round 1) left a 3000x imbalance between
// warps.  Here a CTA's warps share every pillar, and the hardware scheduler balances the groups.
//
// Two cuts of the per-cell cost:
//  * exp(-(dx^2 + dy^2) / 2 sigma^2) = E[|dx|] * E[|dy|] with E[k] = exp(-k^2 / 2 sigma^2) tabulated
//    per pillar in float64 (r + 1 exponentials instead of (2r+1)^2): the float64 product differs
//    from the direct exponential by ~2 ulp(float64), i.e. it rounds to a different float32 for
//    ~1e-8 of the values -- inside the <= 1 ulp(float32) contract of the heat-map.
//    Offsets beyond the table (window clipped by a huge grid) take the direct exponential.
//  * a cell is covered by ~136 pillars, and both reductions are maxima: most updates lose.  A plain
//    load at L2 (possibly stale: values only grow, so a stale read can only cause a redundant
//    atomic, never a wrong skip) filters them; groups run from the LAST pillar down, so that for
//    last_cover (larger index wins) nearly every later update is filtered, too.
// Both planes store value + 1 / bit patterns with 0 = "nothing yet": one memset clears the workspace.
constexpr int kSplatPillars = 4;
constexpr int kSplatTable = 512;  // tabulated offsets per pillar

__global__ void __launch_bounds__(256) k_radar_splat(RadarParams p, const float *__restrict__ rcs,
                                                     const int *__restrict__ coors,
                                                     int *__restrict__ pillar_at,
                                                     int *__restrict__ last_cover,
                                                     int *__restrict__ heat_bits) {
  __shared__ double s_exp[kSplatPillars][kSplatTable];
  __shared__ int s_x[kSplatPillars], s_y[kSplatPillars], s_left[kSplatPillars], s_top[kSplatPillars];
  __shared__ int s_w[kSplatPillars], s_h[kSplatPillars], s_base[kSplatPillars], s_v[kSplatPillars];
  __shared__ double s_denom[kSplatPillars];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int n_groups = ceil_div(p.V, kSplatPillars);
  const double eps = 2.220446049250313e-16;  // np.finfo(float64).eps * h.max(), h.max() == 1
  for (int group = n_groups - 1 - (int)blockIdx.x; group >= 0; group -= gridDim.x) {
    __syncthreads();  // the previous group's tables are no longer read
    if (warp < kSplatPillars) {
      const int v = group * kSplatPillars + warp;
      int rows = 0;
      if (v < p.V) {
        const int b = __ldg(coors + v * 4), y = __ldg(coors + v * 4 + 2), x = __ldg(coors + v * 4 + 3);
        if (b >= 0 && b < p.B && y >= 0 && y < p.ny && x >= 0 && x < p.nx) {
          const int base = b * p.cells;
          if (lane == 0) atomicMax(pillar_at + base + y * p.nx + x, v + 1);
          const int radius = rcs_radius(rcs + (size_t)v * p.rcs_dim, p.rcs_dim);
          // gaussian.py:40-47: clipped window
          const int left = min(x, radius), right = min(p.nx - x, radius + 1);
          const int top = min(y, radius), bottom = min(p.ny - y, radius + 1);
          // gaussian.py:17-23 with sigma = diameter / 6 (gaussian.py:38-39), all float64
          const double diameter = 2.0 * (double)radius + 1.0;
          const double sigma = diameter / 6.0;
          const double denom = 2.0 * sigma * sigma;
          const int reach = min(kSplatTable - 1, max(max(left, right - 1), max(top, bottom - 1)));
          for (int k = lane; k <= reach; k += 32) s_exp[warp][k] = exp(-(double)(k * k) / denom);
          if (lane == 0) {
            s_x[warp] = x, s_y[warp] = y, s_left[warp] = left, s_top[warp] = top, s_w[warp] = left + right;
            s_base[warp] = base, s_v[warp] = v + 1, s_denom[warp] = denom;
          }
          rows = top + bottom;
        }
      }
      if (lane == 0) s_h[warp] = rows;
    }
    __syncthreads();
    // pillar by pillar, the rows of its window dealt to the eight warps: everything that depends on the
    // pillar lives in registers across its rows, a row costs its factor E[|dy|] and one address, a
    // cell one table read, one product and two filtered maxima
#pragma unroll 1
    for (int i = 0; i < kSplatPillars; ++i) {
      const int h = s_h[i];
      if (h == 0) continue;
      const int w = s_w[i], left = s_left[i], top = s_top[i];
      const int v1 = s_v[i];
      const double *tab = s_exp[i];
      int *heat0 = heat_bits + s_base[i] + (s_y[i] - top) * p.nx + (s_x[i] - left);
      int *last0 = last_cover + (heat0 - heat_bits);
      for (int iy = warp; iy < h; iy += 8) {
        const int dy = iy - top;
        const int ay = abs(dy);
        const double ey = ay < kSplatTable ? tab[ay] : -1.0;
        int *heat = heat0 + iy * p.nx, *last = last0 + iy * p.nx;
        for (int ix = lane; ix < w; ix += 32) {
          const int dx = ix - left;
          const int ax = abs(dx);
          double g;
          if (ey >= 0.0 && ax < kSplatTable) g = tab[ax] * ey;
          else g = exp(-(double)((long long)dx * dx + (long long)dy * dy) / s_denom[i]);
          if (g < eps) g = 0.0;
          const int bits = __float_as_int((float)g);  // non-negative floats order like their bit patterns
          // (read at L2, where the atomics execute: a line cached in this SM's L1 would stay stale and
          // let every update through)
          if (bits > __ldcg(heat + ix)) atomicMax(heat + ix, bits);
          if (v1 > __ldcg(last + ix)) atomicMax(last + ix, v1);
        }
      }
    }
  }
}

// Dense write of the three outputs, zeros included: a CTA owns 128 consecutive cells of one sample,
// lane <-> four consecutive cells, warp <-> every 8th channel; every store is a 128-bit quad, a warp
// writes 512 contiguous bytes of a channel row.  The (rare) pillar rows are gathered straight from
// point_features: no staging, no barrier.  Needs cells % 4 == 0 (else the scalar kernel below).
__global__ void __launch_bounds__(256)
    k_radar_write4(RadarParams p, const float *__restrict__ point_features, const float *__restrict__ rcs,
                   const int *__restrict__ pillar_at, const int *__restrict__ last_cover,
                   const int *__restrict__ heat_bits, float *__restrict__ features,
                   float *__restrict__ heatmap, float *__restrict__ heatmap_feat) {
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int tiles_per_sample = ceil_div(p.cells, 128);
  const int b = blockIdx.x / tiles_per_sample;
  const int cell = (blockIdx.x - b * tiles_per_sample) * 128 + lane * 4;
  if (cell >= p.cells) return;
  const size_t g = (size_t)b * p.cells + cell;
  const int4 own = *reinterpret_cast<const int4 *>(pillar_at + g);  // (stored + 1: 0 = no pillar)
  if (heat_bits == nullptr) {
    // heatmap / heatmap_feat were written by k_radar_tiles
  } else if (warp == 0) {
    const int4 hb = *reinterpret_cast<const int4 *>(heat_bits + g);
    st_stream_f4(reinterpret_cast<float4 *>(heatmap + g),
                 make_float4(__int_as_float(hb.x), __int_as_float(hb.y), __int_as_float(hb.z), __int_as_float(hb.w)));
  } else if (warp == 1) {
    const int4 lc = *reinterpret_cast<const int4 *>(last_cover + g);
    const int col = p.rcs_dim - 2;
    float4 v;
    v.x = lc.x > 0 ? __ldg(rcs + (size_t)(lc.x - 1) * p.rcs_dim + col) : 0.f;
    v.y = lc.y > 0 ? __ldg(rcs + (size_t)(lc.y - 1) * p.rcs_dim + col) : 0.f;
    v.z = lc.z > 0 ? __ldg(rcs + (size_t)(lc.z - 1) * p.rcs_dim + col) : 0.f;
    v.w = lc.w > 0 ? __ldg(rcs + (size_t)(lc.w - 1) * p.rcs_dim + col) : 0.f;
    st_stream_f4(reinterpret_cast<float4 *>(heatmap_feat + g), v);
  }
  float *dst = features + (size_t)b * p.Cin * p.cells + cell;
  const bool any = (own.x | own.y | own.z | own.w) != 0;
  for (int c = warp; c < p.Cin; c += 8) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (any) {
      if (own.x > 0) v.x = __ldg(point_features + (size_t)(own.x - 1) * p.Cin + c);
      if (own.y > 0) v.y = __ldg(point_features + (size_t)(own.y - 1) * p.Cin + c);
      if (own.z > 0) v.z = __ldg(point_features + (size_t)(own.z - 1) * p.Cin + c);
      if (own.w > 0) v.w = __ldg(point_features + (size_t)(own.w - 1) * p.Cin + c);
    }
    st_stream_f4(reinterpret_cast<float4 *>(dst + (size_t)c * p.cells), v);
  }
}

// Scalar variant (cells % 4 != 0): 32 consecutive cells x all channels per CTA; pillar rows are
// read coalesced, transposed in shared memory, written as 128-byte channel rows.
__global__ void __launch_bounds__(256)
    k_radar_write(RadarParams p, const float *__restrict__ point_features, const float *__restrict__ rcs,
                  const int *__restrict__ pillar_at, const int *__restrict__ last_cover,
                  const int *__restrict__ heat_bits, float *__restrict__ features,
                  float *__restrict__ heatmap, float *__restrict__ heatmap_feat) {
  extern __shared__ float tile[];  // [Cin][33]
  __shared__ int s_owner[32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int tiles_per_sample = ceil_div(p.cells, 32);
  const int b = blockIdx.x / tiles_per_sample;
  const int cell0 = (blockIdx.x - b * tiles_per_sample) * 32;
  const int n = min(32, p.cells - cell0);
  if (threadIdx.x < 32) {
    int o = -1;
    if (threadIdx.x < n) {
      const int g = b * p.cells + cell0 + threadIdx.x;
      o = pillar_at[g] - 1;  // (stored + 1: 0 = no pillar)
      if (heat_bits != nullptr) {
        heatmap[g] = __int_as_float(heat_bits[g]);
        const int lc = last_cover[g] - 1;
        heatmap_feat[g] = lc >= 0 ? __ldg(rcs + (size_t)lc * p.rcs_dim + p.rcs_dim - 2) : 0.f;
      }
    }
    s_owner[threadIdx.x] = o;
  }
  __syncthreads();
  for (int j = warp; j < 32; j += 8) {
    const int o = s_owner[j];
    for (int c = lane; c < p.Cin; c += 32)
      tile[c * 33 + j] = o >= 0 ? __ldg(point_features + (size_t)o * p.Cin + c) : 0.f;
  }
  __syncthreads();
  float *dst = features + (size_t)b * p.Cin * p.cells + cell0;
  for (int c = warp; c < p.Cin; c += 8)
    if (lane < n) st_stream_f32(dst + (size_t)c * p.cells + lane, tile[c * 33 + lane]);
}

__global__ void __launch_bounds__(256)
    k_radar_scatter_bwd(RadarParams p, const float *__restrict__ features_grad,
                        const int *__restrict__ coors, float *__restrict__ point_features_grad) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < p.V; v += warps) {
    const int b = __ldg(coors + v * 4), y = __ldg(coors + v * 4 + 2), x = __ldg(coors + v * 4 + 3);
    const bool ok = b >= 0 && b < p.B && y >= 0 && y < p.ny && x >= 0 && x < p.nx;
    for (int c = lane; c < p.Cin; c += 32)
      point_features_grad[(size_t)v * p.Cin + c] =
          ok ? __ldg(features_grad + ((size_t)b * p.Cin + c) * p.cells + y * p.nx + x) : 0.f;
  }
}

static int fill_radar(const rcb_radar_desc *d, RadarParams *p) {
  if (!d) return RCB_ERR_ARG;
  if (d->V < 0 || d->Cin <= 0 || d->rcs_dim < 3 || d->B <= 0 || d->ny <= 0 || d->nx <= 0) return RCB_ERR_ARG;
  if ((long long)d->B * d->ny * d->nx * d->Cin >= (1ll << 40)) return RCB_ERR_UNSUPPORTED;
  if ((long long)d->B * d->ny * d->nx >= (1ll << 31)) return RCB_ERR_UNSUPPORTED;
  p->V = d->V, p->Cin = d->Cin, p->rcs_dim = d->rcs_dim, p->B = d->B, p->ny = d->ny, p->nx = d->nx;
  p->cells = d->ny * d->nx;
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

struct RadarWorkspace {
  size_t plane, off_rec, off_range, total;
};
static RadarWorkspace radar_workspace(const RadarParams &p) {
  RadarWorkspace w;
  w.plane = align_up((size_t)p.B * p.cells * 4, 256);
  w.off_rec = w.plane;                                                      // tile form: pillar_at | records | ranges
  w.off_range = w.off_rec + align_up((size_t)max(p.V, 1) * 16, 256);
  const size_t tiles_total = w.off_range + align_up((size_t)p.B * 8, 256);
  w.total = tiles_total > 3 * w.plane ? tiles_total : 3 * w.plane;         // splat form: three planes
  return w;
}

extern "C" size_t rcb_radar_workspace_bytes(const rcb_radar_desc *d) {
  RadarParams p;
  if (fill_radar(d, &p) != RCB_OK) return 0;
  return radar_workspace(p).total;
}

// the (radius, offset) table of the tile kernel: filled once per device
static int radar_table_ready(int device, cudaStream_t s) {
  static std::mutex mu;
  static bool done[64];
  int dev = device;
  if (dev < 0) RCB_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && done[dev]) return RCB_OK;
  k_radar_table<<<ceil_div(kTabR * kTabStride, 256), 256, 0, s>>>();
  RCB_LAUNCH_CHECK();
  if (dev < 64) done[dev] = true;
  return RCB_OK;
}

template <int R>
static int launch_radar_tiles(const RadarParams &p, const float *rcs, const int4 *rec, const int *range, int *pillar_at,
                              float *heatmap, float *heatmap_feat, const float *point_features, float *features,
                              cudaStream_t s) {
  const int tiles_x = ceil_div(p.nx, 32), tiles_y = ceil_div(p.ny, 8 * R);
  float *table = nullptr;
  RCB_CUDA_TRY(cudaGetSymbolAddress((void **)&table, g_radar_tab));
  RCB_CUDA_TRY(launch_pdl(k_radar_tiles<R>, (unsigned)(p.B * tiles_x * tiles_y), 256, 0, s, p, rcs, rec, range, pillar_at,
                          heatmap, heatmap_feat, tiles_x, tiles_y, (const float *)table, point_features, features));
  return RCB_OK;
}

extern "C" int rcb_radar_rcs_scatter(const rcb_radar_desc *d, const float *point_features,
                                     const float *rcs, const int *coors, float *features,
                                     float *heatmap, float *heatmap_feat, void *workspace,
                                     size_t workspace_bytes, int device, rcb_stream_t stream) {
  RadarParams p;
  int rc = fill_radar(d, &p);
  if (rc != RCB_OK) return rc;
  if (!features || !heatmap || !heatmap_feat || !workspace) return RCB_ERR_ARG;
  if (p.V > 0 && (!point_features || !rcs || !coors)) return RCB_ERR_ARG;
  const RadarWorkspace w = radar_workspace(p);
  if (workspace_bytes < w.total) return RCB_ERR_WORKSPACE;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  char *ws = static_cast<char *>(workspace);
  int *pillar_at = (int *)ws;
  const int sms = sm_count_cached(device);
  const char *force_splat = getenv("RCB_RADAR_SPLAT");
  const bool tiles = !(force_splat && force_splat[0] == '1') && (((uintptr_t)coors | (uintptr_t)workspace) % 16) == 0 &&
                     (long long)p.B * ceil_div(p.nx, 32) * ceil_div(p.ny, 8) < (1ll << 31) && p.nx < (1 << 24) && p.ny < (1 << 24);
  const int *last_cover = nullptr, *heat_bits = nullptr;
  if (tiles) {
    int4 *rec = (int4 *)(ws + w.off_rec);
    int *range = (int *)(ws + w.off_range);
    rc = radar_table_ready(device, s);
    if (rc != RCB_OK) return rc;
    RCB_CUDA_TRY(cudaMemsetAsync(range, 0, (size_t)p.B * 8, s));
    if (p.V > 0) {
      k_radar_records<<<ceil_div(p.V, 256), 256, 0, s>>>(p, rcs, coors, rec, range);
      RCB_LAUNCH_CHECK();
    }
    // rows per thread: the fewest for which the grid stays within ~16 CTAs per SM (every CTA scans its
    // sample's records once; measured at 512^2, B = 8: R = 8 161 us per call, R = 4 148, R = 2 149, R = 1 170)
    const long long base = (long long)p.B * ceil_div(p.nx, 32);
    long long budget = (long long)sms * 16;
    if (const char *e = getenv("RCB_RADAR_BUDGET")) budget = (long long)sms * atoi(e);  // development knob
    if (base * ceil_div(p.ny, 8) <= budget) {
      // one row per thread (small grids): the tile kernel writes the feature planes as well
      rc = launch_radar_tiles<1>(p, rcs, rec, range, pillar_at, heatmap, heatmap_feat, point_features, features, s);
      return rc;
    }
    else if (base * ceil_div(p.ny, 16) <= budget) rc = launch_radar_tiles<2>(p, rcs, rec, range, pillar_at, heatmap, heatmap_feat, nullptr, nullptr, s);
    else if (base * ceil_div(p.ny, 32) <= budget) rc = launch_radar_tiles<4>(p, rcs, rec, range, pillar_at, heatmap, heatmap_feat, nullptr, nullptr, s);
    else rc = launch_radar_tiles<8>(p, rcs, rec, range, pillar_at, heatmap, heatmap_feat, nullptr, nullptr, s);
    if (rc != RCB_OK) return rc;
  } else {
    int *lc = (int *)(ws + w.plane), *hb = (int *)(ws + 2 * w.plane);
    RCB_CUDA_TRY(cudaMemsetAsync(ws, 0, 3 * w.plane, s));  // indices are stored + 1, heat as bit patterns: 0 = nothing
    if (p.V > 0) {
      k_radar_splat<<<max(1, min(ceil_div(p.V, kSplatPillars), sms * 16)), 256, 0, s>>>(p, rcs, coors, pillar_at, lc, hb);
      RCB_LAUNCH_CHECK();
    }
    last_cover = lc, heat_bits = hb;
  }
  if (p.cells % 4 == 0 && (((uintptr_t)features | (uintptr_t)heatmap | (uintptr_t)heatmap_feat | (uintptr_t)workspace) % 16) == 0) {
    k_radar_write4<<<p.B * ceil_div(p.cells, 128), 256, 0, s>>>(p, point_features, rcs, pillar_at, last_cover, heat_bits,
                                                                features, heatmap, heatmap_feat);
    RCB_LAUNCH_CHECK();
    return RCB_OK;
  }
  const size_t smem = (size_t)p.Cin * 33 * 4;
  if (smem > 200 * 1024) return RCB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(k_radar_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_radar_write<<<p.B * ceil_div(p.cells, 32), 256, smem, s>>>(p, point_features, rcs, pillar_at,
                                                               last_cover, heat_bits, features, heatmap,
                                                               heatmap_feat);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

extern "C" int rcb_radar_scatter_bwd(const rcb_radar_desc *d, const float *features_grad,
                                     const int *coors, float *point_features_grad, int device,
                                     rcb_stream_t stream) {
  RadarParams p;
  int rc = fill_radar(d, &p);
  if (rc != RCB_OK) return rc;
  if (p.V == 0) return RCB_OK;
  if (!features_grad || !coors || !point_features_grad) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  const int sms = sm_count_cached(device);
  k_radar_scatter_bwd<<<max(1, min(ceil_div(p.V, 8), sms * 32)), 256, 0, (cudaStream_t)stream>>>(
      p, features_grad, coors, point_features_grad);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}
