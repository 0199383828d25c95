// microbench3.cu -- how do the warps of a CTA map onto the SM's sub-partitions, and what does it cost when the heavy
// warps of several serial CTAs share one?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb3 microbench3.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smid() { uint32_t v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }
__device__ __forceinline__ uint32_t warpid() { uint32_t v; asm volatile("mov.u32 %0, %%warpid;" : "=r"(v)); return v; }

struct Rec { uint32_t sm, wid, block, warp; long long cycles; };

// mode 0: the heavy warp is warp 0 of every CTA; mode 1: warp (blockIdx / 148) & 3; mode 2: the warp whose %warpid & 3 equals (blockIdx / 148) & 3
// ilp: 1 = one dependent chain (like the range coder), 4 = four independent chains
__global__ void __launch_bounds__(128) probe(Rec *out, int mode, int ilp, int iters, uint32_t *sink) {
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t k = (blockIdx.x / 148) & 3;
  __shared__ uint32_t heavy;
  if (threadIdx.x == 0) heavy = mode == 0 ? 0 : k;
  __syncthreads();
  if (mode == 2 && lane == 0 && (warpid() & 3) == k) heavy = w;      // racy only if two warps share a sub-partition
  __syncthreads();
  long long t0 = clock64();
  uint32_t a = threadIdx.x, b = 3, c = 5, d = 7;
  if (w == heavy) {
    if (ilp == 1) for (int i = 0; i < iters; i++) { a = a * 1664525u + 1013904223u; a ^= a >> 7; a = a * 22695477u + 1u; a ^= a >> 11; }
    else for (int i = 0; i < iters; i++) { a = a * 1664525u + 1013904223u; b = b * 22695477u + 1u; c = c * 1103515245u + 12345u; d = d * 134775813u + 1u; }
  }
  long long t1 = clock64();
  if (lane == 0) { Rec r; r.sm = smid(); r.wid = warpid(); r.block = blockIdx.x; r.warp = w; r.cycles = (w == heavy) ? t1 - t0 : -1; out[blockIdx.x * 4 + w] = r; }
  if (a + b + c + d == 0x12345) *sink = a;
}

int main() {
  const int iters = 200000;
  uint32_t *sink; cudaMalloc(&sink, 4);
  for (int per_sm : {1, 4, 8}) {
    const int nb = 148 * per_sm;
    Rec *d; cudaMalloc(&d, sizeof(Rec) * nb * 4);
    std::vector<Rec> h(nb * 4);
    for (int ilp : {1, 4}) for (int mode = 0; mode < 3; mode++) {
      cudaMemset(d, 0, sizeof(Rec) * nb * 4);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      probe<<<nb, 128>>>(d, mode, ilp, 1000, sink);  // warm
      cudaEventRecord(e0);
      probe<<<nb, 128>>>(d, mode, ilp, iters, sink);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h.data(), d, sizeof(Rec) * nb * 4, cudaMemcpyDeviceToHost);
      long long mx = 0, mn = 1ll << 60; double sum = 0; int cnt = 0;
      for (auto &r : h) if (r.cycles >= 0) { mx = std::max(mx, r.cycles); mn = std::min(mn, r.cycles); sum += r.cycles; cnt++; }
      printf("per_sm %d ilp %d mode %d: kernel %.3f ms | heavy-warp cycles/iter min %.2f avg %.2f max %.2f (n=%d)\n", per_sm, ilp, mode, ms, (double)mn / iters, sum / cnt / iters, (double)mx / iters, cnt);
      if (ilp == 1 && mode == 0) {
        // mapping: CTAs resident on SM 0 and SM 1: warp-in-block -> %warpid; and how many SMs got exactly per_sm CTAs
        std::vector<int> cnt_sm(256, 0);
        for (int b = 0; b < nb; b++) cnt_sm[h[b * 4].sm]++;
        int exact = 0, mxc = 0; for (int s = 0; s < 256; s++) { if (cnt_sm[s] == per_sm) exact++; mxc = std::max(mxc, cnt_sm[s]); }
        printf("  placement: %d SMs hold exactly %d CTAs, max on one SM %d\n", exact, per_sm, mxc);
        for (int b = 0; b < nb; b++) if (h[b * 4].sm == h[0].sm) {
          printf("  block %4d on sm %3u: warpids", b, h[b * 4].sm);
          for (int w = 0; w < 4; w++) printf(" %2u", h[b * 4 + w].wid);
          printf("\n");
        }
      }
    }
    cudaFree(d);
  }
  return 0;
}
