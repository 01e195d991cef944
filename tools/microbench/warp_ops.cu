// Microbenchmark: per-SM throughput of match.any / ballot / shfl / smem read-modify-write on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o warp_ops warp_ops.cu ; ./warp_ops
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kIters = 4096;
template <int kOp>
__global__ void __launch_bounds__(256) k(unsigned *out, int seed) {
  __shared__ unsigned s[8][1024];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 1024; i += 256) (&s[0][0])[i] = 0;
  __syncthreads();
  unsigned v = (threadIdx.x * 2654435761u + seed) >> 27;  // ~5 bits: a few peers per value
  unsigned acc = 0;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < kIters; ++it) {
    if (kOp == 0) acc += __match_any_sync(0xffffffffu, v + (acc & 1));
    if (kOp == 1) acc += __ballot_sync(0xffffffffu, (v + acc) & 1);
    if (kOp == 2) acc += __shfl_xor_sync(0xffffffffu, v + acc, 1);
    if (kOp == 3) { unsigned b = s[warp][(v + acc) & 1023]; __syncwarp(); s[warp][(v + acc) & 1023] = b + 1; __syncwarp(); acc += b; }
    if (kOp == 4) {  // ten ballots + logic (peer mask by bits)
      unsigned d = v + (acc & 1), p = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 10; ++b) { const bool bit = (d >> b) & 1; const unsigned m = __ballot_sync(0xffffffffu, bit); p &= bit ? m : ~m; }
      acc += p;
    }
    if (kOp == 5) acc += __match_any_sync(0xffffffffu, v);  // independent matches (pipelined)
    if (kOp == 6) acc += __reduce_or_sync(0xffffffffu, v + acc) ^ __reduce_and_sync(0xffffffffu, v + acc);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x * 2] = (unsigned)(t1 - t0);
  out[blockIdx.x * 2 + 1] = acc + lane;
}
template <int kOp>
void run(const char *name, int ctas_per_sm) {
  unsigned *out;
  cudaMalloc(&out, 148 * 8 * 2 * 4);
  k<kOp><<<148 * ctas_per_sm, 256>>>(out, 1);
  cudaDeviceSynchronize();
  k<kOp><<<148 * ctas_per_sm, 256>>>(out, 2);
  cudaDeviceSynchronize();
  unsigned h[2];
  cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h[0];
  printf("%-34s %d CTAs/SM: %8.1f cycles per op per warp; %6.2f SM-cycles per warp-op\n", name, ctas_per_sm,
         cyc / kIters, cyc / kIters / (8.0 * ctas_per_sm));
  cudaFree(out);
}
int main() {
  for (int c : {1, 3}) {
    if (c == 1) { run<0>("match.any (dependent)", 1); run<5>("match.any (independent)", 1); run<1>("ballot (dependent)", 1); run<2>("shfl (dependent)", 1); run<3>("smem RMW + 2 syncwarp", 1); run<4>("10 ballots + logic", 1); run<6>("redux or + and (dependent)", 1); }
    else { run<0>("match.any (dependent)", 3); run<5>("match.any (independent)", 3); run<1>("ballot (dependent)", 3); run<2>("shfl (dependent)", 3); run<3>("smem RMW + 2 syncwarp", 3); run<4>("10 ballots + logic", 3); run<6>("redux or + and (dependent)", 3); }
  }
  return 0;
}
