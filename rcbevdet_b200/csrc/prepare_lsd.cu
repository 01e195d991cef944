// Row P, grids of more than 2^22 cells (up to the 2^24 the reference's fp32 ranks can address,
// view_transformer.py:246-249): plain LSD radix passes over the cells k_cells left in point_cell
// (10 bits per pass, three passes), a binary-search CSR and a look-back scan for the intervals.
// Every pass is stable and the keys are taken in point order, so inside a cell the points come out
// in ascending point index (depth-major; the two-level sort of prepare.cu orders a cell by (pixel,
// depth bin) -- the reference leaves the order unspecified, and the pooling kernels accept either).
// Nobody's hot path: ~3x the time of the two-level sort per point.
#include "prepare_common.cuh"

namespace rcb {

// ---------------------------------------------------------------------------------------------
// Single-pass scan helpers (decoupled look-back, warp-parallel probe), used by K6.
// Tile state word: bits 63..62 flag (1 = aggregate, 2 = inclusive prefix), low bits value.
// total_out (optional) receives the grand total.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct ScanCtl {  // lives in the workspace, zeroed before every run
  unsigned ticket;
  unsigned pad[3];
};

__device__ __forceinline__ unsigned long long lookback_prefix(volatile unsigned long long *st, unsigned tile,
                                                             unsigned long long block_total) {
  // called by warp 0 of the tile; returns the exclusive prefix of this tile (all lanes)
  constexpr unsigned long long kMask = (1ull << 62) - 1;
  const int lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st[0] = (2ull << 62) | block_total;
    return 0;
  }
  if (lane == 0) st[tile] = (1ull << 62) | block_total;
  unsigned long long run = 0;
  int window_end = (int)tile - 1;
  while (true) {
    const int look = window_end - lane;
    unsigned long long v = 0;
    unsigned flag = 3;  // before tile 0: nothing to add
    if (look >= 0) {
      v = st[look];
      flag = (unsigned)(v >> 62);
    }
    const unsigned not_ready = __ballot_sync(kFull, flag == 0);
    const unsigned inclusive = __ballot_sync(kFull, flag == 2 || flag == 3);
    const int first_incl = inclusive ? __ffs(inclusive) - 1 : 32;
    const unsigned needed = first_incl >= 31 ? kFull : ((2u << first_incl) - 1);
    if (not_ready & needed) continue;  // a needed predecessor has not published yet
    unsigned long long add = (lane <= first_incl && flag != 3) ? (v & kMask) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(kFull, add, o);
    run += add;
    if (first_incl < 32) break;
    window_end -= 32;
  }
  if (lane == 0) st[tile] = (2ull << 62) | (run + block_total);
  return run;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long local, unsigned long long *s_warp,
                                                                  unsigned long long *total) {
  unsigned long long incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(kFull, incl, o);
    if (lane_id() >= o) incl += t;
  }
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long warp_off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    const unsigned long long v = s_warp[w];
    if (w < warp) warp_off += v;
    tot += v;
  }
  *total = tot;
  return warp_off + incl - local;
}

// Exclusive scan of the (digit, block) count matrix in digit-major order, without a chain: warp d
// owns digit d's row; its base is the sum of the global digit totals below d (accumulated by the
// histogram kernels), the rest is a scan along the row.  total_out (optional) = grand total.
__global__ void __launch_bounds__(256)
    k_digit_offsets(int n_blocks, unsigned *__restrict__ hist, const unsigned *__restrict__ digit_total,
                    int *__restrict__ total_out) {
  pdl_prologue();
  __shared__ unsigned s_base[kRadixBins];
  __shared__ unsigned s_warp[8];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int d = blockIdx.x * 8 + warp;
  {  // every CTA scans the 1024 digit totals (one coalesced 4 KB read): thread t owns digits 4t .. 4t+3
    const uint4 t4 = reinterpret_cast<const uint4 *>(digit_total)[threadIdx.x];
    const unsigned mine = t4.x + t4.y + t4.z + t4.w;
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned excl = incl - mine;
#pragma unroll
    for (int w = 0; w < 8; ++w)
      if (w < warp) excl += s_warp[w];
    uint4 e4;
    e4.x = excl, e4.y = e4.x + t4.x, e4.z = e4.y + t4.y, e4.w = e4.z + t4.z;
    reinterpret_cast<uint4 *>(s_base)[threadIdx.x] = e4;
    if (total_out != nullptr && blockIdx.x == 0 && threadIdx.x == 255) *total_out = (int)(e4.w + t4.w);
    __syncthreads();
  }
  const unsigned base = s_base[d];
  unsigned *row = hist + (size_t)d * n_blocks;
  unsigned run = base;
  constexpr int kDepth = 8;  // row words in flight per lane
  for (int c0 = 0; c0 < n_blocks; c0 += 32 * kDepth) {
    unsigned v[kDepth];
#pragma unroll
    for (int k = 0; k < kDepth; ++k) {
      const int c = c0 + 32 * k + lane;
      v[k] = c < n_blocks ? __ldcg(row + c) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kDepth; ++k) {
      unsigned incl = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const int c = c0 + 32 * k + lane;
      if (c < n_blocks) row[c] = run + incl - v[k];
      run += __shfl_sync(kFull, incl, 31);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Radix pass, histogram half: per-block counts of digit (key >> shift) over the first *n_ptr keys.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRadixThreads)
    k_radix_hist(const int *__restrict__ keys, const int *__restrict__ n_ptr, int n_fixed, int shift,
                 unsigned *__restrict__ hist, unsigned *__restrict__ digit_total, int n_blocks) {
  pdl_prologue();
  __shared__ unsigned s_hist[kRadixBins];
  for (int i = threadIdx.x; i < kRadixBins; i += kRadixThreads) s_hist[i] = 0;
  __syncthreads();
  const int n = n_ptr ? __ldg(n_ptr) : n_fixed;
  const int base = blockIdx.x * kRadixTile;
  if (base < n) {
    int key[kRadixRounds];
#pragma unroll
    for (int k = 0; k < kRadixRounds; ++k) {
      const int i = base + k * kRadixThreads + threadIdx.x;
      key[k] = i < n ? ld_stream_s32(keys + i) : -1;
    }
#pragma unroll
    for (int k = 0; k < kRadixRounds; ++k)
      if (key[k] >= 0) atomicAdd(&s_hist[((unsigned)key[k] >> shift) & (kRadixBins - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadixBins; i += kRadixThreads) {
    const unsigned c = s_hist[i];
    hist[(size_t)i * n_blocks + blockIdx.x] = c;
    if (c) atomicAdd(digit_total + i, c);  // integer sums: order-independent
  }
}

// ---------------------------------------------------------------------------------------------
// Radix pass, scatter half (stable).  Element order inside a block is (warp, round, lane): warp w
// owns 512 consecutive elements, 32 per round.  Rank of an element among the block's earlier
// elements with the same digit = (same-digit count of earlier warps) + (count of this warp's
// earlier rounds) + (lower lanes of this round with the same digit, via match.any).
//   kFirst: keys = point_cell in point order (dropped points, key < 0, are not emitted), value = index
//   kLast : also emits ranks_feat = pixel of the point index
// ---------------------------------------------------------------------------------------------
template <bool kFirst, bool kLast>
#ifndef RCB_SCATTER_MINCTAS
#define RCB_SCATTER_MINCTAS 3  // measured: 3 (80 registers) 113.5 us of prepare, 4 (64, spills) 116.6, 5 (48) 116.5
#endif
__global__ void __launch_bounds__(kRadixThreads, RCB_SCATTER_MINCTAS)
    k_radix_scatter(const int *__restrict__ keys_in, const int *__restrict__ vals_in, int n_first,
                    const int *__restrict__ n_ptr, int shift, const unsigned *__restrict__ offsets,
                    int n_blocks, int *__restrict__ keys_out, int *__restrict__ vals_out,
                    int *__restrict__ feat_out, PixelMap pm) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char radix_smem[];
  // the counters and the locally grouped tile share 32 KB: the tile is written only after every
  // thread has turned its counters into local positions
  unsigned(*s_cnt)[kRadixBins] = reinterpret_cast<unsigned(*)[kRadixBins]>(radix_smem);  // [warps][bins]
  int *s_key = reinterpret_cast<int *>(radix_smem);                                         // [tile]
  int *s_val = s_key + kRadixTile;                                                          // [tile]
  unsigned *s_gbase = reinterpret_cast<unsigned *>(radix_smem) + kRadixWarps * kRadixBins;  // [bins]
  static_assert(kRadixWarps * kRadixBins * 4 == kRadixTile * 8, "counters and tile alias exactly");
  __shared__ unsigned s_warp_tot[kRadixWarps];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int n = kFirst ? n_first : __ldg(n_ptr);
  const int base = blockIdx.x * kRadixTile + warp * kRadixWarpSpan;
  if (blockIdx.x * kRadixTile >= n) return;
  for (int i = threadIdx.x; i < kRadixWarps * kRadixBins; i += kRadixThreads) (&s_cnt[0][0])[i] = 0;
  __syncthreads();

  int key[kRadixRounds];
  unsigned short rank[kRadixRounds];
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int k = 0; k < kRadixRounds; ++k) {
    const int i = base + k * 32 + lane;
    key[k] = i < n ? ld_stream_s32(keys_in + i) : -1;
  }
  // Peer masks of all rounds first -- ballots, issued back to back (match.any would serialise the
  // SM's warps: 39 -> 30 us for this kernel) -- each reduced to (same-digit lanes below me, group
  // size) in rank[k]; only the counter update is a chain across rounds.
#pragma unroll
  for (int k = 0; k < kRadixRounds; ++k) {
    const bool valid = key[k] >= 0;
    const unsigned digit = ((unsigned)key[k] >> shift) & (kRadixBins - 1);
    const unsigned peers = peers_by_ballot<kRadixBits>(digit, valid, kRadixBits);
    rank[k] = (unsigned short)(__popc(peers & lt) | (__popc(peers) << 8));
  }
#pragma unroll
  for (int k = 0; k < kRadixRounds; ++k) {
    const bool valid = key[k] >= 0;
    const unsigned digit = ((unsigned)key[k] >> shift) & (kRadixBins - 1);
    const unsigned lower = rank[k] & 0xffu, group = rank[k] >> 8;
    unsigned before = 0;
    if (valid) before = s_cnt[warp][digit];
    __syncwarp();
    rank[k] = (unsigned short)(before + lower);
    if (valid && lower == 0) s_cnt[warp][digit] = before + group;
    __syncwarp();
  }
  __syncthreads();
  // Per digit (thread t owns digits t, t + 256, ...: bank-conflict-free): block total, exclusive
  // prefix over the warps, then an exclusive prefix over the digits IN THAT THREAD-MAJOR ORDER ->
  // where every (digit, warp) group sits in the block's locally grouped tile (the order of the
  // digit groups inside the tile is irrelevant, each goes to its own global range).  s_cnt[w][d] becomes that local start; s_gbase[d] = global start - local
  // start of the digit, so that global position = s_gbase[digit] + local position.
  {
    unsigned tot[kRadixBins / kRadixThreads], mine = 0, goff[kRadixBins / kRadixThreads];
#pragma unroll
    for (int j = 0; j < kRadixBins / kRadixThreads; ++j)
      goff[j] = __ldg(offsets + (size_t)(threadIdx.x + j * kRadixThreads) * n_blocks + blockIdx.x);
#pragma unroll
    for (int j = 0; j < kRadixBins / kRadixThreads; ++j) {
      // (loads batched in front of the stores: through the aliased shared-memory pointer the
      // compiler would otherwise keep every load behind the previous store)
      unsigned c[kRadixWarps];
#pragma unroll
      for (int w = 0; w < kRadixWarps; ++w) c[w] = s_cnt[w][threadIdx.x + j * kRadixThreads];
      unsigned run = 0;
#pragma unroll
      for (int w = 0; w < kRadixWarps; ++w) {
        s_cnt[w][threadIdx.x + j * kRadixThreads] = run;  // warp offset inside the digit, for now
        run += c[w];
      }
      tot[j] = run;
      mine += run;
    }
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    unsigned digit_start = incl - mine;
#pragma unroll
    for (int w = 0; w < kRadixWarps; ++w)
      if (w < warp) digit_start += s_warp_tot[w];
#pragma unroll
    for (int j = 0; j < kRadixBins / kRadixThreads; ++j) {
      const int d = threadIdx.x + j * kRadixThreads;
      unsigned c[kRadixWarps];
#pragma unroll
      for (int w = 0; w < kRadixWarps; ++w) c[w] = s_cnt[w][d];
#pragma unroll
      for (int w = 0; w < kRadixWarps; ++w) s_cnt[w][d] = c[w] + digit_start;
      s_gbase[d] = goff[j] - digit_start;
      digit_start += tot[j];
    }
  }
  __syncthreads();
  // locally sorted tile in shared memory ...
#pragma unroll
  for (int k = 0; k < kRadixRounds; ++k) {
    if (key[k] < 0) continue;
    const unsigned digit = ((unsigned)key[k] >> shift) & (kRadixBins - 1);
    rank[k] = (unsigned short)(s_cnt[warp][digit] + rank[k]);  // local position
  }
  const unsigned n_valid = s_warp_tot[0] + s_warp_tot[1] + s_warp_tot[2] + s_warp_tot[3] + s_warp_tot[4] +
                           s_warp_tot[5] + s_warp_tot[6] + s_warp_tot[7];
  __syncthreads();  // counters are dead from here: their memory becomes the tile
#pragma unroll
  for (int k = 0; k < kRadixRounds; ++k) {
    if (key[k] >= 0) {
      const int i = base + k * 32 + lane;
      s_key[rank[k]] = key[k];
      s_val[rank[k]] = kFirst ? i : ld_stream_s32(vals_in + i);
    }
  }
  __syncthreads();
  // ... written out in sorted order: a digit's elements go to consecutive global addresses
  for (unsigned l = threadIdx.x; l < n_valid; l += kRadixThreads) {
    const int kk = s_key[l], vv = s_val[l];
    const unsigned pos = s_gbase[((unsigned)kk >> shift) & (kRadixBins - 1)] + l;
    keys_out[pos] = kk;
    vals_out[pos] = vv;
    if (kLast) feat_out[pos] = pixel_of_point(vv, pm);
  }
}

// ---------------------------------------------------------------------------------------------
// K5: dense CSR over BEV cells from the sorted cells: cell_start[c] = lower_bound(ranks_bev, c).
// One thread per cell (+1), a 24-step binary search over keys that sit in L2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_bounds(int n_cells, const int *__restrict__ sorted_cells,
                                                     const int *__restrict__ n_ptr,
                                                     int *__restrict__ cell_start) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_cells) return;
  const int n = __ldg(n_ptr);
  int lo = 0, hi = n;  // first index with key >= c
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(sorted_cells + mid) < c) lo = mid + 1;
    else hi = mid;
  }
  cell_start[c] = lo;
}

// ---------------------------------------------------------------------------------------------
// K6: intervals = the non-empty cells, compacted (view_transformer.py:254-262).  Exclusive scan of
// the non-empty flags (same single-pass scan), interval_starts / interval_lengths straight from the
// CSR, counts = {n_kept, n_intervals}.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads)
    k_intervals_scan(int n_cells, const int *__restrict__ cell_start, int *__restrict__ interval_starts,
                int *__restrict__ interval_lengths, unsigned long long *__restrict__ tile_state,
                ScanCtl *__restrict__ ctl, int *__restrict__ counts) {
  pdl_prologue();
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_warp[kScanThreads / 32];
  __shared__ unsigned long long s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(&ctl->ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  int start[kScanItems + 1];
#pragma unroll
  for (int k = 0; k <= kScanItems; ++k) start[k] = __ldg(cell_start + min(base + k, n_cells));
  unsigned long long local = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) local += (base + k < n_cells && start[k + 1] > start[k]) ? 1u : 0u;
  unsigned long long block_total;
  const unsigned long long excl = block_exclusive_scan(local, s_warp, &block_total);
  if (threadIdx.x < 32) {
    const unsigned long long pre = lookback_prefix(tile_state, tile, block_total);
    if (threadIdx.x == 0) s_prefix = pre;
  }
  __syncthreads();
  int iv = (int)(s_prefix + excl);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n_cells && start[k + 1] > start[k]) {
      interval_starts[iv] = start[k];
      interval_lengths[iv] = start[k + 1] - start[k];
      ++iv;
    }
  }
  if (tile == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) {
    counts[0] = __ldg(cell_start + n_cells);
    counts[1] = iv;
    counts[2] = 0;
    counts[3] = 0;
  }
}

constexpr size_t kRadixScatterSmem = (size_t)(kRadixWarps * kRadixBins + kRadixBins) * 4;  // 36 KB

struct LsdWorkspace {
  size_t off_ctl, off_state, off_totals, zero_bytes, off_hist, off_keys, off_vals, total;
  int n_blocks, n_cell_tiles, n_passes;
};

static LsdWorkspace lsd_layout(const PrepParams &p) {
  LsdWorkspace w;
  w.n_blocks = max(1, ceil_div(p.P, kRadixTile));
  int bits = 1;
  while ((1ll << bits) < (long long)p.n_cells) ++bits;
  w.n_passes = ceil_div(bits, kRadixBits);
  w.n_cell_tiles = ceil_div(p.n_cells, kScanTile);
  size_t o = 0;
  w.off_ctl = o, o += 256;                                                   // ScanCtl of k_intervals_scan
  w.off_state = o, o += align_up((size_t)w.n_cell_tiles * 8, 256);           // its tile states
  w.off_totals = o, o += (size_t)3 * kRadixBins * 4;                         // digit totals, one set per pass
  w.zero_bytes = o;                                                          // all of the above start at zero
  w.off_hist = o, o += align_up((size_t)kRadixBins * w.n_blocks * 4, 256);
  w.off_keys = o, o += align_up((size_t)(p.P + 4) * 4, 256);
  w.off_vals = o, o += align_up((size_t)(p.P + 4) * 4, 256);
  w.total = o;
  return w;
}

size_t lsd_workspace_bytes(const PrepParams &p) { return lsd_layout(p).total; }

int lsd_prepare(const PrepParams &p, const int *point_cell, int *ranks_bev, int *ranks_depth, int *ranks_feat,
                int *interval_starts, int *interval_lengths, int *cell_start, int *counts, void *workspace,
                size_t workspace_bytes, cudaStream_t s) {
  const LsdWorkspace w = lsd_layout(p);
  if (workspace_bytes < w.total) return RCB_ERR_WORKSPACE;
  char *ws = (char *)workspace;
  ScanCtl *ctl = (ScanCtl *)(ws + w.off_ctl);
  unsigned long long *state = (unsigned long long *)(ws + w.off_state);
  unsigned *hist = (unsigned *)(ws + w.off_hist), *totals = (unsigned *)(ws + w.off_totals);
  int *tmp_keys = (int *)(ws + w.off_keys), *tmp_vals = (int *)(ws + w.off_vals);
  RCB_CUDA_TRY(cudaMemsetAsync(ws, 0, w.zero_bytes, s));
  PixelMap pm;
  pm.by_dhw = FastDiv::make((unsigned)p.DHW);
  pm.by_hw = FastDiv::make((unsigned)p.HW);
  const int nb = w.n_blocks;
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRadixScatterSmem));
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRadixScatterSmem));
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRadixScatterSmem));
  RCB_CUDA_TRY(cudaFuncSetAttribute(k_radix_scatter<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRadixScatterSmem));
  // ping-pong so that the last pass lands in the caller's arrays
  const int *in_keys = point_cell, *in_vals = nullptr;
  for (int pass = 0; pass < w.n_passes; ++pass) {
    const bool first = pass == 0, last = pass == w.n_passes - 1;
    const bool to_final = ((w.n_passes - 1 - pass) % 2) == 0;
    int *out_keys = to_final ? ranks_bev : tmp_keys, *out_vals = to_final ? ranks_depth : tmp_vals;
    const int shift = pass * kRadixBits;
    // pass 0 counts all P slots of point_cell (negative = dropped, not counted); later passes the n_kept keys
    k_radix_hist<<<nb, kRadixThreads, 0, s>>>(in_keys, first ? nullptr : counts, first ? p.P : 0, shift, hist,
                                              totals + pass * kRadixBins, nb);
    RCB_LAUNCH_CHECK();
    // the first scan's grand total is n_kept: later passes and the CSR / interval kernels read it from counts[0]
    k_digit_offsets<<<kRadixBins / 8, 256, 0, s>>>(nb, hist, totals + pass * kRadixBins, first ? counts : nullptr);
    RCB_LAUNCH_CHECK();
    if (first && last)
      k_radix_scatter<true, true><<<nb, kRadixThreads, kRadixScatterSmem, s>>>(in_keys, in_vals, p.P, counts, shift, hist, nb,
                                                                               out_keys, out_vals, ranks_feat, pm);
    else if (first)
      k_radix_scatter<true, false><<<nb, kRadixThreads, kRadixScatterSmem, s>>>(in_keys, in_vals, p.P, counts, shift, hist, nb,
                                                                                out_keys, out_vals, ranks_feat, pm);
    else if (last)
      k_radix_scatter<false, true><<<nb, kRadixThreads, kRadixScatterSmem, s>>>(in_keys, in_vals, p.P, counts, shift, hist, nb,
                                                                                out_keys, out_vals, ranks_feat, pm);
    else
      k_radix_scatter<false, false><<<nb, kRadixThreads, kRadixScatterSmem, s>>>(in_keys, in_vals, p.P, counts, shift, hist, nb,
                                                                                 out_keys, out_vals, ranks_feat, pm);
    RCB_LAUNCH_CHECK();
    in_keys = out_keys, in_vals = out_vals;
  }
  k_cell_bounds<<<ceil_div(p.n_cells + 1, 256), 256, 0, s>>>(p.n_cells, ranks_bev, counts, cell_start);
  RCB_LAUNCH_CHECK();
  k_intervals_scan<<<w.n_cell_tiles, kScanThreads, 0, s>>>(p.n_cells, cell_start, interval_starts, interval_lengths, state,
                                                           ctl, counts);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

}  // namespace rcb
