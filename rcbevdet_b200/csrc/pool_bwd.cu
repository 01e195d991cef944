// Row B -- bev_pool_v2 backward.
// Reference: mmdet3d/ops/bev_pool_v2/bev_pool.py:43-83 (re-sort by ranks_feat, rebuild intervals,
// zero-filled gradients) and src/bev_pool_cuda.cu:67-121 (one thread per pixel, serial over its
// points and channels):
//   depth_grad[ranks_depth[p]]   = sum_c out_grad[ranks_bev[p], c] * feat[ranks_feat[p], c]
//   feat_grad[ranks_feat[p], c] += out_grad[ranks_bev[p], c] * depth[ranks_depth[p]]
//
//  Structured ranks (ranks_depth unique, ranks_feat = pixel of ranks_depth -- what prepare emits):
//      no sort at all.  The points of pixel (b, n, h, w) are the D depth bins of that pixel, and
//      the inverse map point_cell[P] (BEV cell of each frustum point, -1 = dropped; a by-product
//      of prepare) says where each went.  A pixel's context row stays in registers while its depth
//      column is walked; every kept bin gathers its out_grad row; feat_grad accumulates in
//      registers (written once, zeros for unseen pixels), depth_grad is a transposed warp
//      reduction of the per-bin dot products.  Both outputs are fully written, deterministic, no
//      atomics, no memset.
//      k_pool_bwd_pixels16 (C = 64 / 80 / 128, the hot one): 16 lanes per pixel, 16 pixels per
//      CTA, columns staged through shared memory, merged (cell, pixel) runs -- see its comment.
//      k_pool_bwd_pixels (other C): one warp per pixel, 32 bins at a time.
//
//  k_pool_bwd_points (general ranks): one warp per point; depth_grad exact, feat_grad by float
//      atomics into a zero-filled buffer.
#include "common.cuh"

namespace rcb {

// sum over lanes of v[k] for every k in [0,32): lane l returns the total of index bitrev-free
// mapping k = l (see the selects below: at each step the half of the values a lane does not own
// is sent to its partner).  31 shuffles.
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const bool up = lane & 16;
    const float send = up ? v[k] : v[k + 16];
    const float keep = up ? v[k + 16] : v[k];
    v[k] = keep + __shfl_xor_sync(kFull, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up = lane & 8;
    const float send = up ? v[k] : v[k + 8];
    const float keep = up ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(kFull, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 4;
    const float send = up ? v[k] : v[k + 4];
    const float keep = up ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(kFull, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 2;
    const float send = up ? v[k] : v[k + 2];
    const float keep = up ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(kFull, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(kFull, send, 1);
  }
  // lane l now holds the total of index k with bits (16,8,4,2,1) of l selecting the upper half at
  // each step, i.e. k == l.
  return v[0];
}

struct BwdPixelParams {
  const float *out_grad_rows;  // (n_cells, C) channels last
  const float *depth;
  const void *feat;
  const int *point_cell;
  float *depth_grad;
  float *feat_grad;
  int n_pixels, D, HW, DHW, C4;
  int H, W;  // 0 when unknown (then pixels are walked in memory order)
  const int *gate;  // launch gate (common.cuh)
};

// kQ = 128-bit quads per lane (C <= 128*kQ)
template <typename FeatT, int kQ>
__global__ void __launch_bounds__(256) k_pool_bwd_pixels(BwdPixelParams p) {
  if (gate_closed(p.gate)) return;
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const FeatT *feat = static_cast<const FeatT *>(p.feat);
  const float4 *og_rows = reinterpret_cast<const float4 *>(p.out_grad_rows);
  for (int pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < p.n_pixels; pix += warps) {
    const int bn = pix / p.HW, hw = pix - bn * p.HW;
    const int col = bn * p.DHW + hw;  // point index of depth bin 0
    float4 f[kQ], fg[kQ];
#pragma unroll
    for (int j = 0; j < kQ; ++j) {
      const int q = lane + 32 * j;
      f[j] = q < p.C4 ? Row4<FeatT>::load(feat, (size_t)pix * p.C4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      fg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int d0 = 0; d0 < p.D; d0 += 32) {
      const int my_d = d0 + lane;
      int my_cell = -1;
      float my_w = 0.f;
      if (my_d < p.D) {
        my_cell = __ldg(p.point_cell + col + my_d * p.HW);
        my_w = __ldg(p.depth + col + my_d * p.HW);
      }
      float dot[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const int cell = __shfl_sync(kFull, my_cell, k);
        const float w = __shfl_sync(kFull, my_w, k);
        float s = 0.f;
        if (cell >= 0) {
#pragma unroll
          for (int j = 0; j < kQ; ++j) {
            const int q = lane + 32 * j;
            if (q < p.C4) {
              const float4 g = __ldg(og_rows + (size_t)cell * p.C4 + q);
              s = fmaf(g.x, f[j].x, s), s = fmaf(g.y, f[j].y, s);
              s = fmaf(g.z, f[j].z, s), s = fmaf(g.w, f[j].w, s);
              fg[j].x = fmaf(g.x, w, fg[j].x), fg[j].y = fmaf(g.y, w, fg[j].y);
              fg[j].z = fmaf(g.z, w, fg[j].z), fg[j].w = fmaf(g.w, w, fg[j].w);
            }
          }
        }
        dot[k] = s;
      }
      const float total = transpose_reduce32(dot, lane);
      if (my_d < p.D) p.depth_grad[col + my_d * p.HW] = total;
    }
#pragma unroll
    for (int j = 0; j < kQ; ++j) {
      const int q = lane + 32 * j;
      if (q < p.C4) reinterpret_cast<float4 *>(p.feat_grad)[(size_t)pix * p.C4 + q] = fg[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Half-warp variant (C a multiple of 16): 16 lanes per pixel, two adjacent pixels per warp, each
// lane owning kNQ 128-bit quads (channels 64*j + 4*l .. +3) and kNS scalars (channels
// 64*kNQ + 16*s + l).  C = 80 -> one quad + one scalar per lane: no idle lanes (the 32-lane
// kernel above runs 20 of 32).
//
// The kernel is bound by the L1 data pipe (ncu: 71 % of its wavefronts, shuffles included), so
// everything that is not an out_grad row goes through it as cheaply as possible:
//  * a CTA owns 16 consecutive pixels; their point_cell / depth columns (stride HW words per
//    depth bin) are fetched as 64-byte runs across the 16 pixels into shared memory and read back
//    column-wise, depth_grad leaves the same way: 2 wavefronts per 32 words instead of ~20;
//  * the compacted (cell, weight) slots of a half-warp are exchanged through 144 bytes of shared
//    memory, two slots per 128-bit load, instead of two shuffles per slot;
//  * the per-bin dot products are reduced with a transposed butterfly sized to the number of
//    slots in use (8 shuffles for <= 8 slots, 15 otherwise).
// ---------------------------------------------------------------------------------------------
template <int kSlots>
__device__ __forceinline__ float transpose_reduce16(float (&v)[kSlots], int l16) {
  static_assert(kSlots == 4 || kSlots == 8 || kSlots == 16, "slot tiers");
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    constexpr int kTop = kSlots;
    const int n = (kTop * h) / 16;  // values still held per lane after this step
    if (n == 0) {  // fewer slots than lanes: the lanes of a pair hold the same slot
      v[0] += __shfl_xor_sync(kFull, v[0], h);
    } else {
#pragma unroll
      for (int k = 0; k < kSlots / 2; ++k) {
        if (k >= n) break;
        const bool up = l16 & h;
        const float send = up ? v[k] : v[k + n];
        const float keep = up ? v[k + n] : v[k];
        v[k] = keep + __shfl_xor_sync(kFull, send, h);
      }
    }
  }
  return v[0];  // slot s ends up in the 16 / kSlots lanes starting at s * 16 / kSlots
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static __device__ __forceinline__ float load(const char *p) { return __ldg(reinterpret_cast<const float *>(p)); }
};
template <>
struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float load(const char *p) {
    return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16 *>(p)));
  }
};
template <>
struct Elem<__half> {
  static __device__ __forceinline__ float load(const char *p) {
    return __half2float(__ldg(reinterpret_cast<const __half *>(p)));
  }
};

constexpr int kBwdTilePitch = 17;        // words per depth bin in the staged tiles (16 pixels + 1)
constexpr int kBwdSlotPitch = 18;        // int2 per half-warp slot table (16 slots + 16 bytes)
constexpr int kBwdMaxChunk = 128;        // depth bins staged at once
__host__ __device__ inline size_t bwd16_smem_bytes(int d_chunk) {
  return (size_t)3 * d_chunk * kBwdTilePitch * 4 + (size_t)8 * 2 * kBwdSlotPitch * 8;
}

template <typename FeatT, int kNQ, int kNS>
__global__ void __launch_bounds__(256, 4) k_pool_bwd_pixels16(BwdPixelParams p, int d_chunk) {
  pdl_prologue();
  if (gate_closed(p.gate)) return;
  constexpr int kC = 64 * kNQ + 16 * kNS;
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  int *s_cell = reinterpret_cast<int *>(bwd_smem);            // [d_chunk][17]
  float *s_w = reinterpret_cast<float *>(s_cell + d_chunk * kBwdTilePitch);
  float *s_dg = s_w + d_chunk * kBwdTilePitch;
  int2 *s_slot = reinterpret_cast<int2 *>(s_dg + d_chunk * kBwdTilePitch);  // [8 warps][2 halves][18]
  const int lane = lane_id(), l16 = lane & 15, half = lane >> 4, warp = threadIdx.x >> 5;
  int2 *my_slots = s_slot + (warp * 2 + half) * kBwdSlotPitch;
  const char *og = reinterpret_cast<const char *>(p.out_grad_rows);
  const char *feat = static_cast<const char *>(p.feat);
  const int px_mine = warp * 2 + half;  // my pixel inside the CTA's group of 16
  const int px_load = threadIdx.x & 15;  // the pixel column this thread stages
  for (int group = blockIdx.x; group * 16 < p.n_pixels; group += gridDim.x) {
    const int pix = group * 16 + px_mine;
    const bool live = pix < p.n_pixels;
    const int pix_c = live ? pix : p.n_pixels - 1;
    float4 fq[kNQ > 0 ? kNQ : 1], gq[kNQ > 0 ? kNQ : 1];
    float fs[kNS > 0 ? kNS : 1], gs[kNS > 0 ? kNS : 1];
    const char *frow = feat + (size_t)pix_c * kC * sizeof(FeatT);
#pragma unroll
    for (int j = 0; j < kNQ; ++j) {
      fq[j] = Row4<FeatT>::load_bytes(frow + (64 * j + 4 * l16) * sizeof(FeatT));
      gq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kNS; ++j) {
      fs[j] = Elem<FeatT>::load(frow + (64 * kNQ + 16 * j + l16) * sizeof(FeatT));
      gs[j] = 0.f;
    }
    // staging role: this thread moves depth bins (threadIdx.x >> 4) + 16 k of pixel column px_load
    const int pix_l = group * 16 + px_load;
    const bool live_l = pix_l < p.n_pixels;
    const int bn_l = (live_l ? pix_l : 0) / p.HW;
    const int col_l = bn_l * p.DHW + ((live_l ? pix_l : 0) - bn_l * p.HW);  // point index of depth bin 0
    // per-lane bases of the quad part and of the scalar tail of an out_grad row, pinned in
    // registers so that each row address is ONE 32x32+64 multiply-add of the cell index
    const char *og_lane = og + (size_t)l16 * 16;
    const char *og_tail = og + 256 * kNQ + (size_t)l16 * 4;
    asm volatile("" : "+l"(og_lane), "+l"(og_tail));

    for (int dc0 = 0; dc0 < p.D; dc0 += d_chunk) {
      const int n_d = min(d_chunk, p.D - dc0);
      // eight depth bins per thread and step, all sixteen loads issued before the first store (the
      // trip count is a run-time value: left alone, the loop keeps two loads in flight per trip)
      for (int ds = threadIdx.x >> 4; ds < d_chunk; ds += 128) {
        int c8[8];
        float w8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int d = ds + 16 * u;
          c8[u] = -1, w8[u] = 0.f;
          if (live_l && d < n_d) {
            c8[u] = __ldg(p.point_cell + col_l + (dc0 + d) * p.HW);
            w8[u] = __ldg(p.depth + col_l + (dc0 + d) * p.HW);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int d = ds + 16 * u;
          if (d < d_chunk) {
            s_cell[d * kBwdTilePitch + px_load] = c8[u];
            s_w[d * kBwdTilePitch + px_load] = w8[u];
          }
        }
      }
      __syncthreads();

      for (int d0 = 0; d0 < n_d; d0 += 16) {
        const int my_cell = s_cell[(d0 + l16) * kBwdTilePitch + px_mine];
        const float my_w = s_w[(d0 + l16) * kBwdTilePitch + px_mine];
        if (__ballot_sync(kFull, my_cell >= 0) == 0u) {  // all 32 bins dropped (rays that left the grid)
          s_dg[(d0 + l16) * kBwdTilePitch + px_mine] = 0.f;
          continue;
        }
        // Consecutive depth bins of a pixel often fall into the same BEV cell (1.42 bins per (cell,
        // pixel) pair on average: a ray crosses a 0.8 m cell in ~1.6 steps of 0.5 m).  Such a run
        // shares its out_grad row and its dot product, so only the run's first bin (at most 4 bins
        // per group) loads the row: it carries the group's summed depth weight for feat_grad and
        // hands its dot product to the followers afterwards.  depth_grad is bit-identical to the
        // unmerged computation; feat_grad differs by the rounding of (w1 + w2) * g vs w1 * g + w2 * g.
        const int prev_cell = __shfl_up_sync(kFull, my_cell, 1, 16);
        const bool starts_run = l16 == 0 || my_cell != prev_cell;
        const unsigned starts16 = (__ballot_sync(kFull, starts_run) >> (half * 16)) & 0xffffu;
        const int run_head = 31 - __clz((int)(starts16 & ((2u << l16) - 1u)));  // lane where my run starts
        const int pos = (l16 - run_head) & 3;                                   // position inside a group of <= 4
        const bool is_head = my_cell >= 0 && pos == 0;
        const int head_lane = l16 - pos;
        float w_group = my_w;
#pragma unroll
        for (int j = 1; j < 4; ++j) {
          const float wn = __shfl_down_sync(kFull, my_w, j, 16);
          const int pn = __shfl_down_sync(kFull, pos, j, 16);
          if (l16 + j < 16 && pn == j) w_group += wn;  // lane + j is the j-th follower of my group
        }
        // The heads are compacted: slot j of a half-warp is its j-th head, published through the
        // half-warp's slot table (unused slots read as cell -1).
        const unsigned heads16 = (__ballot_sync(kFull, is_head) >> (half * 16)) & 0xffffu;
        const int n_heads = __popc(heads16);
        const int n_heads_max = max(n_heads, __shfl_xor_sync(kFull, n_heads, 16));
        __syncwarp();  // the previous batch's slot reads are done
        if (is_head) my_slots[__popc(heads16 & ((1u << l16) - 1u))] = make_int2(my_cell, __float_as_int(w_group));
        if (l16 >= n_heads) my_slots[l16] = make_int2(-1, 0);
        __syncwarp();

        auto slot_pair = [&](int k2, float &dot_a, float &dot_b) {  // slots 2*k2 and 2*k2 + 1
          const int4 sl = *reinterpret_cast<const int4 *>(my_slots + 2 * k2);
          const int cells[2] = {sl.x, sl.z};
          const float ws[2] = {__int_as_float(sl.y), __int_as_float(sl.w)};
          float dots[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            float s = 0.f;
            if (cells[t] >= 0) {
              // one IMAD.WIDE each: 32-bit cell index x row bytes + pinned 64-bit lane base
              const char *row = og_lane + (size_t)(unsigned)cells[t] * (size_t)(kC * 4);
              const char *tail = og_tail + (size_t)(unsigned)cells[t] * (size_t)(kC * 4);
              const float2 ww = make_float2(ws[t], ws[t]);
              float2 acc = make_float2(0.f, 0.f);
#pragma unroll
              for (int j = 0; j < kNQ; ++j) {
                const float4 g = __ldg(reinterpret_cast<const float4 *>(row + 256 * j));
                const float2 glo = make_float2(g.x, g.y), ghi = make_float2(g.z, g.w);
                acc = __ffma2_rn(glo, make_float2(fq[j].x, fq[j].y), acc);
                acc = __ffma2_rn(ghi, make_float2(fq[j].z, fq[j].w), acc);
                const float2 a = __ffma2_rn(glo, ww, make_float2(gq[j].x, gq[j].y));
                const float2 b = __ffma2_rn(ghi, ww, make_float2(gq[j].z, gq[j].w));
                gq[j] = make_float4(a.x, a.y, b.x, b.y);
              }
              s = acc.x + acc.y;
#pragma unroll
              for (int j = 0; j < kNS; ++j) {
                // scalar tail: channel 64*kNQ + 16*j + l16
                const float g = __ldg(reinterpret_cast<const float *>(tail + 64 * j));
                s = fmaf(g, fs[j], s);
                gs[j] = fmaf(g, ws[t], gs[j]);
              }
            }
            dots[t] = s;
          }
          dot_a = dots[0], dot_b = dots[1];
        };
        // a bin's value sits in the slot of its group head: slot = rank of the head lane among the heads
        const int my_slot = __popc(heads16 & ((1u << head_lane) - 1u));
        float total;
        if (n_heads_max <= 4) {  // warp-uniform tiers: fewer predicated bodies, shorter reduction
          float dot[4];
#pragma unroll
          for (int k2 = 0; k2 < 2; ++k2) slot_pair(k2, dot[2 * k2], dot[2 * k2 + 1]);
          total = transpose_reduce16<4>(dot, l16);
          total = __shfl_sync(kFull, total, 4 * my_slot, 16);
        } else if (n_heads_max <= 8) {
          float dot[8];
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) slot_pair(k2, dot[2 * k2], dot[2 * k2 + 1]);
          total = transpose_reduce16<8>(dot, l16);
          total = __shfl_sync(kFull, total, 2 * my_slot, 16);
        } else {
          float dot[16];
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) slot_pair(k2, dot[2 * k2], dot[2 * k2 + 1]);
#pragma unroll
          for (int k = 12; k < 16; ++k) dot[k] = 0.f;
          if (n_heads_max > 12) {
#pragma unroll
            for (int k2 = 6; k2 < 8; ++k2) slot_pair(k2, dot[2 * k2], dot[2 * k2 + 1]);
          }
          total = transpose_reduce16<16>(dot, l16);
          total = __shfl_sync(kFull, total, my_slot, 16);
        }
        s_dg[(d0 + l16) * kBwdTilePitch + px_mine] = my_cell >= 0 ? total : 0.f;
      }
      __syncthreads();
      for (int d = threadIdx.x >> 4; d < n_d; d += 16)
        if (live_l) p.depth_grad[col_l + (dc0 + d) * p.HW] = s_dg[d * kBwdTilePitch + px_load];
      // (the next chunk's staging writes s_cell / s_w only; s_dg is rewritten after its barrier)
    }
    if (live) {
      float *grow = p.feat_grad + (size_t)pix * kC;
#pragma unroll
      for (int j = 0; j < kNQ; ++j) *reinterpret_cast<float4 *>(grow + 64 * j + 4 * l16) = gq[j];
#pragma unroll
      for (int j = 0; j < kNS; ++j) grow[64 * kNQ + 16 * j + l16] = gs[j];
    }
  }
}

// General ranks: warp per point.
template <typename FeatT>
__global__ void __launch_bounds__(256)
    k_pool_bwd_points(int n_points, int C, const float *__restrict__ out_grad_rows,
                      const float *__restrict__ depth, const FeatT *__restrict__ feat,
                      const int *__restrict__ ranks_depth, const int *__restrict__ ranks_feat,
                      const int *__restrict__ ranks_bev, float *__restrict__ depth_grad,
                      float *__restrict__ feat_grad) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_points; i += warps) {
    const int rd = __ldg(ranks_depth + i), rf = __ldg(ranks_feat + i), rb = __ldg(ranks_bev + i);
    const float w = __ldg(depth + rd);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = __ldg(out_grad_rows + (size_t)rb * C + c);
      s = fmaf(g, to_f32<FeatT>(feat[(size_t)rf * C + c]), s);
      atomicAdd(feat_grad + (size_t)rf * C + c, g * w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) depth_grad[rd] = s;
  }
}

int planes_to_rows_launch(const void *src, void *dst, int n_img, int C, int HW,
                          long long src_img_stride, int elem_bytes, cudaStream_t s, const int *gate);

template <typename FeatT>
static int launch_pixels(BwdPixelParams &p, int sms, cudaStream_t s) {
  const int C = p.C4 * 4;
  const int grid16 = max(1, min(ceil_div(p.n_pixels, 16), sms * 24));
  if (C == 80 || C == 64 || C == 128) {
    const int d_chunk = min(kBwdMaxChunk, ceil_div(p.D, 16) * 16);
    const size_t smem = bwd16_smem_bytes(d_chunk);  // <= 28 KB
    if (C == 80) RCB_CUDA_TRY(launch_pdl(k_pool_bwd_pixels16<FeatT, 1, 1>, grid16, 256, smem, s, p, d_chunk));
    else if (C == 64) RCB_CUDA_TRY(launch_pdl(k_pool_bwd_pixels16<FeatT, 1, 0>, grid16, 256, smem, s, p, d_chunk));
    else RCB_CUDA_TRY(launch_pdl(k_pool_bwd_pixels16<FeatT, 2, 0>, grid16, 256, smem, s, p, d_chunk));
    return RCB_OK;
  }
  const int grid = max(1, min(ceil_div(p.n_pixels, 8), sms * 32));
  if (p.C4 <= 32)
    k_pool_bwd_pixels<FeatT, 1><<<grid, 256, 0, s>>>(p);
  else
    k_pool_bwd_pixels<FeatT, 2><<<grid, 256, 0, s>>>(p);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

extern "C" size_t rcb_pool_bwd_workspace_bytes(const rcb_pool_desc *d) {
  if (check_pool_desc(d) != RCB_OK) return 0;
  if (d->layout != RCB_LAYOUT_B_C_CELLS) return 256;
  return align_up((size_t)d->B * d->Z * d->Y * d->X * d->C * 4, 256);
}

extern "C" int rcb_bev_pool_v2_bwd(const rcb_pool_desc *d, const float *out_grad, const float *depth,
                                   const void *feat, const int *ranks_depth, const int *ranks_feat,
                                   const int *ranks_bev, const int *point_cell, float *depth_grad,
                                   float *feat_grad, void *workspace, size_t workspace_bytes,
                                   int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!out_grad || !depth || !feat || !depth_grad || !feat_grad) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  const int cps = d->Z * d->Y * d->X;
  const int sms = sm_count_cached(device);

  // out_grad as channels-last rows (the reference gets them from permute-backward + .contiguous(),
  // bev_pool.py:69,91)
  const float *og_rows = out_grad;
  if (d->layout == RCB_LAYOUT_B_C_CELLS) {
    const size_t need = (size_t)d->B * cps * d->C * 4;
    if (!workspace || workspace_bytes < need) return RCB_ERR_WORKSPACE;
    rc = planes_to_rows_launch(out_grad, workspace, d->B, d->C, cps, (long long)d->C * cps, 4, s, launch_gate());
    if (rc != RCB_OK) return rc;
    og_rows = static_cast<const float *>(workspace);
  }

  const int elem = d->feat_dtype == RCB_DTYPE_F32 ? 4 : 2;
  const bool structured = (d->flags & RCB_PLAN_STRUCTURED) && point_cell != nullptr && d->D > 0 &&
                          d->HW > 0 && (d->C % 4) == 0 && d->C <= 256 &&
                          (((uintptr_t)feat) % (4 * elem)) == 0 && (((uintptr_t)og_rows) % 16) == 0 &&
                          (((uintptr_t)feat_grad) % 16) == 0 &&
                          (long long)d->n_pixels * d->D == (long long)d->n_depth &&
                          d->n_pixels % d->HW == 0;  // whole camera images: col + d*HW stays inside n_depth
  if (structured) {
    BwdPixelParams p;
    p.out_grad_rows = og_rows, p.depth = depth, p.feat = feat, p.point_cell = point_cell;
    p.depth_grad = depth_grad, p.feat_grad = feat_grad;
    p.n_pixels = d->n_pixels, p.D = d->D, p.HW = d->HW, p.DHW = d->D * d->HW, p.C4 = d->C / 4;
    p.H = d->H > 0 && d->HW % d->H == 0 ? d->H : 1;
    p.W = d->HW / p.H;
    p.gate = launch_gate();
    switch (d->feat_dtype) {
      case RCB_DTYPE_F32: return launch_pixels<float>(p, sms, s);
      case RCB_DTYPE_BF16: return launch_pixels<__nv_bfloat16>(p, sms, s);
      default: return launch_pixels<__half>(p, sms, s);
    }
  }
  if (launch_gate()) return RCB_ERR_UNSUPPORTED;  // the point kernel takes no launch gate
  // general path (bev_pool.py:67-68 zero-fill, then accumulate)
  RCB_CUDA_TRY(cudaMemsetAsync(depth_grad, 0, (size_t)d->n_depth * 4, s));
  RCB_CUDA_TRY(cudaMemsetAsync(feat_grad, 0, (size_t)d->n_pixels * d->C * 4, s));
  if (d->n_points == 0) return RCB_OK;
  if (!ranks_depth || !ranks_feat || !ranks_bev) return RCB_ERR_ARG;
  const int grid = max(1, min(ceil_div(d->n_points, 8), sms * 32));
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32:
      k_pool_bwd_points<float><<<grid, 256, 0, s>>>(d->n_points, d->C, og_rows, depth,
                                                    static_cast<const float *>(feat), ranks_depth,
                                                    ranks_feat, ranks_bev, depth_grad, feat_grad);
      break;
    case RCB_DTYPE_BF16:
      k_pool_bwd_points<__nv_bfloat16><<<grid, 256, 0, s>>>(
          d->n_points, d->C, og_rows, depth, static_cast<const __nv_bfloat16 *>(feat), ranks_depth,
          ranks_feat, ranks_bev, depth_grad, feat_grad);
      break;
    default:
      k_pool_bwd_points<__half><<<grid, 256, 0, s>>>(d->n_points, d->C, og_rows, depth,
                                                     static_cast<const __half *>(feat), ranks_depth,
                                                     ranks_feat, ranks_bev, depth_grad, feat_grad);
  }
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}
