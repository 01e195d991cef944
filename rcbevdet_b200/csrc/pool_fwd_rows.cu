// Row F -- bev_pool_v2 forward, row-staging kernel (the product path for ranks whose cells are
// sorted and whose context rows live in the cell's own sample -- everything prepare emits).
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48; zero-fill + permute of
// mmdet3d/ops/bev_pool_v2/bev_pool.py:27,91 are folded in (every cell written once, final layout).
//
// One CTA per 8 x 4 patch of BEV cells; thread <-> (cell, lane), a lane owning two 128-bit channel
// quads of that cell's output row (C / 8 lanes per cell: 320 threads at C = 80).
//
// Neighbouring cells see the same camera pixels along a ray: a patch with ~700 points touches only
// ~110 distinct context rows.  Round 1's kernel fetched the 320-byte row once per POINT through
// L1/LSU (920 MB per launch against 88 MB algorithmic).  Here the rows a patch needs are brought
// from L2 ONCE, by asynchronous 16-byte copies (cp.async, no registers), into a double-buffered
// shared-memory stage and consumed with LDS.128 + packed FMAs:
//
//   A  stage the chunk's points: (pixel, depth weight) pairs -- all index loads, then the depth
//      gathers -- and mark the pixels in a shared-memory bitmap over the sample's pixels
//   B  prefix of the bitmap's popcounts -> every distinct pixel gets a slot (in pixel order), the
//      slot list is written, every point's pixel is replaced by its slot's byte offset
//   C  rounds of `sh` slots: while round r is accumulated from one half of the stage, round r + 1
//      is already arriving in the other.  A thread walks the points of ITS cell (its accumulators
//      never leave registers: no partial sums, no combine pass) and takes those whose slot lies in
//      the round.
//   D  the patch is written through a shared-memory transpose: (B, C, Z*Y*X) in runs of 8 cells per
//      channel row, or channels-last.  Empty cells get their zeros here.
//
// Patches with more than kEMax points repeat A-C per chunk of kEMax points; accumulators persist.
// The order of additions inside a cell depends only on the data (slot range, then point order):
// bit-reproducible, and identical to the reference's point order whenever a patch needs one round.
#include "common.cuh"

namespace rcb {

constexpr int kRowsTileX = 8;
constexpr int kRowsTileY = 4;
constexpr int kRowsCells = kRowsTileX * kRowsTileY;  // 32
constexpr int kEMax = 2048;                          // points staged per chunk
constexpr int kRowsTabBytes = 1024;

struct FwdRowsParams {
  const float *depth;
  const void *feat;
  const int *ranks_depth;
  const int *ranks_feat;
  const int *cell_start;
  float *out;
  int C, L;              // channels, lanes per cell (C / 8)
  int X, R;              // cells per row, rows per sample (Z*Y)
  int tiles_x, tiles_r;  // patches per sample
  int cells_per_sample;
  int layout;
  int B;
  int npix_sample;       // context rows (pixels) per sample
  int bw;                // bitmap words = ceil(npix_sample / 32)
  int pixel_major;       // inside a cell the points ascend by pixel (what prepare emits): cursor walk
  int sh;                // slots per half of the row stage
  unsigned pitch;        // bytes between staged rows (row bytes + 16: conflict-free quads)
  unsigned row_bytes;    // C * sizeof(FeatT)
  FastDiv by_B, by_tiles_x;
};

// bytes of the double-buffered row stage; the [C][33] write-out tile aliases it
__host__ __device__ inline size_t fwd_rows_stage_bytes(int sh, unsigned pitch, int C) {
  const size_t stage = (size_t)2 * sh * pitch;
  const size_t transpose = ((size_t)C * 33 * 4 + 15) / 16 * 16;
  return stage < transpose ? transpose : stage;
}
__host__ __device__ inline size_t fwd_rows_smem_bytes(int sh, unsigned pitch, int bw, int C) {
  return fwd_rows_stage_bytes(sh, pitch, C) + (size_t)kEMax * 8 + (size_t)kEMax * 2 + (size_t)bw * 8 +
         (size_t)(kEMax / 32) * 4 + (size_t)(kEMax + 8) * 2 + kRowsTabBytes;
}

// i-th element of {c, c-1, c+1, c-2, c+2, ...} clipped to [0, n), c = n / 2
__device__ __forceinline__ int rows_zigzag(int i, int n) {
  const int c = n >> 1;
  const int lo_side = c, hi_side = n - 1 - c;
  const int paired = 2 * min(lo_side, hi_side) + 1;
  if (i < paired) return (i & 1) ? c - ((i + 1) >> 1) : c + (i >> 1);
  const int rest = i - paired;
  return lo_side > hi_side ? c - hi_side - 1 - rest : c + lo_side + 1 + rest;
}

template <typename T>
__device__ __forceinline__ float4 lds_row4(const unsigned char *p);
template <>
__device__ __forceinline__ float4 lds_row4<float>(const unsigned char *p) {
  return *reinterpret_cast<const float4 *>(p);
}
template <>
__device__ __forceinline__ float4 lds_row4<__nv_bfloat16>(const unsigned char *p) {
  const uint2 raw = *reinterpret_cast<const uint2 *>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 lds_row4<__half>(const unsigned char *p) {
  const uint2 raw = *reinterpret_cast<const uint2 *>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ void rows_fma(float4 &acc, const float4 v, const float w) {
  const float2 ww = make_float2(w, w);
  const float2 lo = __ffma2_rn(make_float2(v.x, v.y), ww, make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(v.z, v.w), ww, make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ void cp_async16(unsigned dst_shared, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// Points [a, e) of this thread's cell against the staged rows of one round.  kCheck = false: the
// chunk needs a single round, every point's row is resident -> straight-line groups of four points,
// eight 128-bit shared-memory loads in flight.  kCheck = true: take the points whose slot lies in
// this round's range (offset relative to the round's first slot < bytes of one half).
template <typename FeatT, bool kCheck>
__device__ __forceinline__ void rows_accumulate(const uint2 *__restrict__ ent, int a, int e,
                                                const unsigned char *__restrict__ bufp, unsigned rbase,
                                                unsigned half_bytes, unsigned qoff0, unsigned qoff1,
                                                float4 &acc0, float4 &acc1) {
  int i = a;
  if (!kCheck) {
    for (; i + 4 <= e; i += 4) {
      uint2 ev[4];
      float4 v0[4], v1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ev[u] = ent[i + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v0[u] = lds_row4<FeatT>(bufp + ev[u].x + qoff0);
        v1[u] = lds_row4<FeatT>(bufp + ev[u].x + qoff1);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        rows_fma(acc0, v0[u], __uint_as_float(ev[u].y));
        rows_fma(acc1, v1[u], __uint_as_float(ev[u].y));
      }
    }
    for (; i < e; ++i) {
      const uint2 ev = ent[i];
      rows_fma(acc0, lds_row4<FeatT>(bufp + ev.x + qoff0), __uint_as_float(ev.y));
      rows_fma(acc1, lds_row4<FeatT>(bufp + ev.x + qoff1), __uint_as_float(ev.y));
    }
  } else {
    for (; i < e; i += 4) {
      uint2 ev[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ev[u] = (i + u < e) ? ent[i + u] : make_uint2(0xffffffffu, 0u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned rel = ev[u].x - rbase;
        if (rel < half_bytes) {
          rows_fma(acc0, lds_row4<FeatT>(bufp + rel + qoff0), __uint_as_float(ev[u].y));
          rows_fma(acc1, lds_row4<FeatT>(bufp + rel + qoff1), __uint_as_float(ev[u].y));
        }
      }
    }
  }
}

// Multi-round walk when the points of a cell ascend by pixel (= by slot): the points of round r are
// a contiguous piece of the cell's list; `cur` remembers where the previous round stopped.
template <typename FeatT>
__device__ __forceinline__ void rows_accumulate_cursor(const uint2 *__restrict__ ent, int &cur, int e,
                                                       const unsigned char *__restrict__ bufp, unsigned rbase,
                                                       unsigned half_bytes, unsigned qoff0, unsigned qoff1,
                                                       float4 &acc0, float4 &acc1) {
  while (cur < e) {
    uint2 ev[4];
    int taken = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) ev[u] = (cur + u < e) ? ent[cur + u] : make_uint2(0xffffffffu, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned rel = ev[u].x - rbase;
      if (rel < half_bytes) {
        rows_fma(acc0, lds_row4<FeatT>(bufp + rel + qoff0), __uint_as_float(ev[u].y));
        rows_fma(acc1, lds_row4<FeatT>(bufp + rel + qoff1), __uint_as_float(ev[u].y));
        ++taken;
      }
    }
    cur += taken;
    if (taken < 4) break;
  }
}

// kL > 0: lanes per cell known at compile time (blockDim.x == 32 * kL); 0: run time (p.L).
template <typename FeatT, int kL>
__global__ void __launch_bounds__(kL ? 32 * kL : 1024, kL ? (kL <= 10 ? 3 : (kL <= 16 ? 2 : 1)) : 1)
    k_pool_fwd_rows(FwdRowsParams p) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem[];
  const int L = kL ? kL : p.L;
  const int T = 32 * L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = T >> 5;
  const unsigned half_bytes = (unsigned)p.sh * p.pitch;
  const size_t stage_bytes = fwd_rows_stage_bytes(p.sh, p.pitch, p.C);
  unsigned char *rowbuf = smem;
  uint2 *ent = reinterpret_cast<uint2 *>(smem + stage_bytes);            // [kEMax] (slot offset | pixel, weight)
  unsigned short *upix = reinterpret_cast<unsigned short *>(ent + kEMax);  // [kEMax] slot -> pixel of the sample
  unsigned *bitmap = reinterpret_cast<unsigned *>(upix + kEMax);         // [bw]
  unsigned *wpre = bitmap + p.bw;                                        // [bw] slots before each word
  unsigned *cs_bits = wpre + p.bw;                                       // [kEMax / 32] first point of a cell?
  unsigned short *pos_map = reinterpret_cast<unsigned short *>(cs_bits + kEMax / 32);  // [kEMax + 8] raw -> merged index
  int *tab = reinterpret_cast<int *>(pos_map + kEMax + 8);
  int *cell_lo = tab;        // [32] patch-local point range of every cell
  int *cell_hi = tab + 32;   // [32]
  int *seg_off = tab + 64;   // [kRowsTileY + 1] patch-local prefix of the row slices
  int *seg_g = tab + 72;     // [kRowsTileY] global start of each row slice
  unsigned *s_warp = reinterpret_cast<unsigned *>(tab + 80);  // [32] scan scratch
  unsigned *s_nslots = reinterpret_cast<unsigned *>(tab + 112);
  int *new_lo = tab + 120;   // [32] chunk-local entry range of every cell after merging
  int *new_hi = tab + 152;   // [32]

  // Launch order: patches nearest the grid centre first, samples interleaved (point density peaks
  // around the ego vehicle, so the long patches start at once).  Any order is correct.
  const int t = (int)p.by_B.div(blockIdx.x);
  const int b = (int)blockIdx.x - t * p.B;
  const int t_r = (int)p.by_tiles_x.div((unsigned)t);
  const int tx_i = rows_zigzag(t - t_r * p.tiles_x, p.tiles_x);
  const int tr_i = rows_zigzag(t_r, p.tiles_r);
  const int x0 = tx_i * kRowsTileX, r0 = tr_i * kRowsTileY;
  const int nx = min(kRowsTileX, p.X - x0), nr = min(kRowsTileY, p.R - r0);
  const int cell_base = b * p.cells_per_sample;

  // ---- patch geometry, by warp 0: lane <-> cell ------------------------------------------------
  if (tid < 32) {
    const int ty = tid / kRowsTileX, tx = tid % kRowsTileX;
    int s = 0, e = 0;
    if (ty < nr && tx < nx) {
      const int c = cell_base + (r0 + ty) * p.X + x0 + tx;
      s = __ldg(p.cell_start + c);
      e = __ldg(p.cell_start + c + 1);
    }
    int off = 0, my_off = 0, my_g = 0;
#pragma unroll
    for (int r = 0; r < kRowsTileY; ++r) {
      const int g = __shfl_sync(kFull, s, r * kRowsTileX);
      const int ge = __shfl_sync(kFull, e, r * kRowsTileX + nx - 1);
      const int len = r < nr ? ge - g : 0;
      if (r == ty) my_off = off, my_g = g;
      if (tid == r) seg_off[r] = off, seg_g[r] = g;
      off += len;
    }
    if (tid == 0) seg_off[kRowsTileY] = off;
    cell_lo[tid] = s - my_g + my_off;
    cell_hi[tid] = e - my_g + my_off;
  }
  __syncthreads();
  const int total = seg_off[kRowsTileY];

  const int cell = tid / L, l = tid - cell * L;
  const unsigned qoff0 = (unsigned)l * 4u * (unsigned)sizeof(FeatT);
  const unsigned qoff1 = (unsigned)(l + L) * 4u * (unsigned)sizeof(FeatT);
  const int pix_base = b * p.npix_sample;
  const unsigned char *feat_s = static_cast<const unsigned char *>(p.feat) + (size_t)pix_base * p.row_bytes;
  const unsigned rowbuf_sa = (unsigned)__cvta_generic_to_shared(rowbuf);
  const int chunks_per_row = (int)(p.row_bytes >> 4);
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;

  for (int cb = 0; cb < total; cb += kEMax) {
    const int n = min(kEMax, total - cb);
    // ---- A: points of the chunk -> (pixel of the sample, depth weight); mark the pixels ----------
    for (int i = tid; i < p.bw; i += T) bitmap[i] = 0u;
    for (int i = tid; i < kEMax / 32; i += T) cs_bits[i] = 0u;
    __syncthreads();
    if (tid < 32) {  // first point (inside this chunk) of every non-empty cell: a merged run never crosses it
      const int lo = cell_lo[tid], hi = cell_hi[tid];
      if (hi > lo && hi > cb && lo < cb + n) {
        const int idx = max(lo, cb) - cb;
        atomicOr(&cs_bits[idx >> 5], 1u << (idx & 31));
      }
    }
    for (int i0 = tid; i0 < n; i0 += 4 * T) {
      int g[4], rd[4], rf[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * T;
        const int pt = cb + i;
        int ty = 0;
#pragma unroll
        for (int r = 1; r < kRowsTileY; ++r) ty += (pt >= seg_off[r]);
        g[k] = i < n ? seg_g[ty] + (pt - seg_off[ty]) : -1;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        rd[k] = g[k] >= 0 ? ld_stream_s32(p.ranks_depth + g[k]) : 0;
        rf[k] = g[k] >= 0 ? ld_stream_s32(p.ranks_feat + g[k]) : 0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (g[k] >= 0) {
          const float w = ld_stream_f32(p.depth + rd[k]);
          // (the plan guarantees 0 <= pl < npix_sample; the clamp only keeps a violated guarantee
          // from writing outside the bitmap)
          const unsigned pl = min((unsigned)(rf[k] - pix_base), (unsigned)(p.npix_sample - 1));
          ent[i0 + k * T] = make_uint2(pl, __float_as_uint(w));
          atomicOr(&bitmap[pl >> 5], 1u << (pl & 31u));
        }
      }
    }
    __syncthreads();
    // ---- B: slots = exclusive prefix of the bitmap's popcounts; the slot -> pixel list ------------
    {
      const int per = (p.bw + T - 1) / T;
      const int w0 = tid * per;
      unsigned mine = 0;
      for (int j = 0; j < per; ++j)
        if (w0 + j < p.bw) mine += __popc(bitmap[w0 + j]);
      unsigned incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      unsigned run = incl - mine;
      for (int w = 0; w < warp; ++w) run += s_warp[w];
      for (int j = 0; j < per; ++j) {
        const int wd = w0 + j;
        if (wd >= p.bw) break;
        unsigned bits = bitmap[wd];
        wpre[wd] = run;
        while (bits) {
          const int bit = __ffs(bits) - 1;
          bits &= bits - 1;
          upix[run++] = (unsigned short)(wd * 32 + bit);
        }
      }
      if (tid == T - 1) *s_nslots = run;
    }
    __syncthreads();
    const int n_slots = (int)*s_nslots;
    const int n_rounds = (n_slots + p.sh - 1) / p.sh;

    // rows of slots [r * sh, (r + 1) * sh) -> half (r & 1) of the stage, 16 bytes per copy
    auto stage_round = [&](int r) {
      const int s0 = r * p.sh;
      const int cnt = min(p.sh, n_slots - s0);
      const unsigned dst0 = rowbuf_sa + (unsigned)(r & 1) * half_bytes;
      const int items = cnt * chunks_per_row;
      for (int it = tid; it < items; it += T) {
        const int s = it / chunks_per_row, ch = it - s * chunks_per_row;
        const unsigned pix = upix[s0 + s];
        cp_async16(dst0 + (unsigned)s * p.pitch + (unsigned)ch * 16u, feat_s + (size_t)pix * p.row_bytes + (size_t)ch * 16);
      }
      cp_async_commit();
    };
    stage_round(0);
    auto off_of = [&](int i) -> unsigned {  // byte offset of the slot of raw point i
      const unsigned pl = ent[i].x;
      const unsigned wd = pl >> 5;
      return (wpre[wd] + __popc(bitmap[wd] & ((1u << (pl & 31u)) - 1u))) * p.pitch;
    };
    if (kL > 0) {
      // ---- merge: consecutive points of one cell that read the same row (the depth bins of one
      // (cell, pixel) pair, adjacent in prepare's order) become ONE entry whose weight is the sum of
      // their depth weights, added in point order.  Thread t owns raw points [t*per, (t+1)*per); a
      // run belongs to the thread that owns its first point.  Everything is held in registers across
      // the barrier, so the compaction happens in place.
      constexpr int kPer = kL > 0 ? (kEMax + 32 * kL - 1) / (32 * kL) : 1;
      auto is_cs = [&](int i) -> bool { return (cs_bits[i >> 5] >> (i & 31)) & 1u; };
      const int per = (n + T - 1) / T;
      const int i0 = min(n, tid * per), i1 = min(n, i0 + per), cnt = i1 - i0;
      unsigned off[kPer];
      float ws[kPer];
      unsigned heads = 0, cs_mask = 0;
      unsigned prev = (cnt > 0 && i0 > 0) ? off_of(i0 - 1) : 0xffffffffu;
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        off[k] = 0u, ws[k] = 0.f;
        if (k < cnt) {
          off[k] = off_of(i0 + k);
          ws[k] = __uint_as_float(ent[i0 + k].y);
          const bool cs = is_cs(i0 + k);
          if (cs || off[k] != prev) heads |= 1u << k;
          if (cs) cs_mask |= 1u << k;
          prev = off[k];
        }
      }
      // running sums inside a run, left to right (point order) ...
#pragma unroll
      for (int k = 1; k < kPer; ++k)
        if (k < cnt && !((heads >> k) & 1u)) ws[k] = ws[k - 1] + ws[k];
      // ... and the run's total copied back to its head
#pragma unroll
      for (int k = kPer - 2; k >= 0; --k)
        if (k + 1 < cnt && !((heads >> (k + 1)) & 1u)) ws[k] = ws[k + 1];
      // the last run of the range may continue in the next threads' points
      if (heads != 0u && i1 < n) {
        const int last = 31 - __clz(heads);
        float tot = 0.f;
        unsigned o = 0u;
#pragma unroll
        for (int k = 0; k < kPer; ++k)
          if (k == last) tot = ws[k], o = off[k];
        for (int i = i1; i < n && !is_cs(i) && off_of(i) == o; ++i) tot += __uint_as_float(ent[i].y);
#pragma unroll
        for (int k = 0; k < kPer; ++k)
          if (k == last) ws[k] = tot;
      }
      // block-wide exclusive prefix of the head counts (every raw read above happens before its barrier)
      const unsigned mine = __popc(heads);
      unsigned incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      unsigned base = incl - mine;
      for (int w = 0; w < warp; ++w) base += s_warp[w];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        if ((heads >> k) & 1u) {
          const unsigned pos = base + __popc(heads & ((1u << k) - 1u));
          ent[pos] = make_uint2(off[k], __float_as_uint(ws[k]));
          if ((cs_mask >> k) & 1u) pos_map[i0 + k] = (unsigned short)pos;
        }
      }
      if (tid == T - 1) pos_map[n] = (unsigned short)(base + mine);
      __syncthreads();
      if (tid < 32) {
        const int lo = cell_lo[tid], hi = cell_hi[tid];
        new_lo[tid] = pos_map[min(max(lo - cb, 0), n)];
        new_hi[tid] = pos_map[min(max(hi - cb, 0), n)];
      }
    } else {
      for (int i = tid; i < n; i += T) ent[i].x = off_of(i);
      if (tid < 32) {
        new_lo[tid] = min(max(cell_lo[tid] - cb, 0), n);
        new_hi[tid] = min(max(cell_hi[tid] - cb, 0), n);
      }
    }
    // ---- C: rounds ---------------------------------------------------------------------------------
    int a = 0, e = 0, cur = 0;
    for (int r = 0; r < n_rounds; ++r) {
      if (r + 1 < n_rounds) {
        stage_round(r + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();  // round r's rows (every thread's copies) and, for r == 0, the entries and cell ranges
      if (r == 0) a = cur = new_lo[cell], e = new_hi[cell];
      const unsigned char *bufp = rowbuf + (size_t)(r & 1) * half_bytes;
      const unsigned rbase = (unsigned)r * half_bytes;  // offset of the round's first slot
      if (n_rounds == 1)
        rows_accumulate<FeatT, false>(ent, a, e, bufp, 0u, half_bytes, qoff0, qoff1, acc0, acc1);
      else if (p.pixel_major)
        rows_accumulate_cursor<FeatT>(ent, cur, e, bufp, rbase, half_bytes, qoff0, qoff1, acc0, acc1);
      else
        rows_accumulate<FeatT, true>(ent, a, e, bufp, rbase, half_bytes, qoff0, qoff1, acc0, acc1);
      __syncthreads();  // this half is refilled two rounds on; the chunk's tables are rebuilt after the last
    }
  }

  // ---- D: write the whole patch, empty cells included ----------------------------------------------
  const int cty = cell / kRowsTileX, ctx = cell % kRowsTileX;
  if (p.layout == RCB_LAYOUT_CELLS_C) {
    if (cty < nr && ctx < nx) {
      float4 *dst = reinterpret_cast<float4 *>(p.out + ((size_t)cell_base + (size_t)(r0 + cty) * p.X + x0 + ctx) * p.C);
      st_stream_f4(dst + l, acc0);
      st_stream_f4(dst + l + L, acc1);
    }
    return;
  }
  // (B, C, cells): transpose through shared memory (the stage is free: the loop ends on a barrier, or
  // never ran), then every store instruction writes kRowsTileY runs of kRowsTileX consecutive cells
  float *ts = reinterpret_cast<float *>(rowbuf);  // [C][33]
  {
    const int c0 = 4 * l, c1 = 4 * (l + L);
    ts[(c0 + 0) * 33 + cell] = acc0.x, ts[(c0 + 1) * 33 + cell] = acc0.y;
    ts[(c0 + 2) * 33 + cell] = acc0.z, ts[(c0 + 3) * 33 + cell] = acc0.w;
    ts[(c1 + 0) * 33 + cell] = acc1.x, ts[(c1 + 1) * 33 + cell] = acc1.y;
    ts[(c1 + 2) * 33 + cell] = acc1.z, ts[(c1 + 3) * 33 + cell] = acc1.w;
  }
  __syncthreads();
  const int ty = lane / kRowsTileX, tx = lane % kRowsTileX;
  if (ty >= nr || tx >= nx) return;
  float *dst = p.out + (size_t)b * p.C * p.cells_per_sample + (size_t)(r0 + ty) * p.X + x0 + tx;
  for (int ch = warp; ch < p.C; ch += n_warps) st_stream_f32(dst + (size_t)ch * p.cells_per_sample, ts[ch * 33 + lane]);
}

template <typename FeatT, int kL>
static int launch_rows_l(FwdRowsParams &p, long long grid, cudaStream_t s) {
  const size_t smem = fwd_rows_smem_bytes(p.sh, p.pitch, p.bw, p.C);
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_rows<FeatT, kL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RCB_CUDA_TRY(launch_pdl(k_pool_fwd_rows<FeatT, kL>, (unsigned)grid, 32 * p.L, smem, s, p));
  return RCB_OK;
}

template <typename FeatT>
static int launch_rows(FwdRowsParams &p, long long grid, cudaStream_t s) {
  switch (p.L) {
    case 8: return launch_rows_l<FeatT, 8>(p, grid, s);
    case 10: return launch_rows_l<FeatT, 10>(p, grid, s);
    case 16: return launch_rows_l<FeatT, 16>(p, grid, s);
    default: return launch_rows_l<FeatT, 0>(p, grid, s);
  }
}

// Can the row-staging kernel take this problem?  (Sorted cells with a CSR, context rows inside the
// cell's own sample, whole 128-bit quad pairs per lane, 16-byte aligned rows, a sample's pixels
// addressable by the 16-bit slot list.)
bool fwd_rows_eligible(const rcb_pool_desc *d, const void *feat, const int *cell_start) {
  if (!cell_start || !(d->flags & RCB_PLAN_SORTED_CELLS) || !(d->flags & RCB_PLAN_SAMPLE_LOCAL)) return false;
  if ((d->C % 8) != 0 || d->C > 256 || d->B <= 0 || d->n_pixels <= 0 || (d->n_pixels % d->B) != 0) return false;
  if (d->n_pixels / d->B > 65536) return false;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  if (((size_t)d->C * elem) % 16 != 0 || (((uintptr_t)feat) % 16) != 0) return false;
  return true;
}

int fwd_rows_launch(const rcb_pool_desc *d, const float *depth, const void *feat, const int *ranks_depth,
                    const int *ranks_feat, const int *cell_start, float *out, cudaStream_t s) {
  FwdRowsParams p;
  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  p.depth = depth, p.feat = feat, p.ranks_depth = ranks_depth, p.ranks_feat = ranks_feat;
  p.cell_start = cell_start, p.out = out;
  p.C = d->C, p.L = d->C / 8;
  p.X = d->X, p.R = d->Z * d->Y;
  p.tiles_x = ceil_div(p.X, kRowsTileX), p.tiles_r = ceil_div(p.R, kRowsTileY);
  p.cells_per_sample = d->Z * d->Y * d->X, p.layout = d->layout, p.B = d->B;
  p.npix_sample = d->n_pixels / d->B;
  p.bw = ceil_div(p.npix_sample, 32);
  p.row_bytes = (unsigned)d->C * elem;
  p.pitch = p.row_bytes + 16;
  // ~43 KB of row stage per CTA: 64 slots per half at C = 80 fp32
  p.sh = max(8, min(64, (int)(44032u / (2u * p.pitch))));
  p.pixel_major = (d->flags & RCB_PLAN_PIXEL_MAJOR) ? 1 : 0;
  p.by_B = FastDiv::make((unsigned)p.B), p.by_tiles_x = FastDiv::make((unsigned)p.tiles_x);
  const long long grid = (long long)d->B * p.tiles_r * p.tiles_x;
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32: return launch_rows<float>(p, grid, s);
    case RCB_DTYPE_BF16: return launch_rows<__nv_bfloat16>(p, grid, s);
    default: return launch_rows<__half>(p, grid, s);
  }
}

}  // namespace rcb
