// True dependent-op latencies (chains unrolled x32 inside a loop so the loop branch is amortised).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
#define U 32
template <int K> __global__ void k(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
  uint32_t x = a + threadIdx.x * 0, y = b; uint64_t z64 = a;
  extern __shared__ uint32_t sm[];
  sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 32] = 1; __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (K == 0) { x = x * y + 1; }
      if (K == 1) { x = __umulhi(x, y) ^ 0x9E3779B9u; }
      if (K == 2) { x = x + y; }
      if (K == 3) { x = x ^ (x >> 3); }            // SHF + LOP
      if (K == 4) { x = (x < y) ? x + 7 : x - 3; } // 2 adds + ISETP + SEL
      if (K == 5) { x = __reduce_max_sync(0xffffffffu, x) + 1; }
      if (K == 6) { x = sm[x & 31] + 1; }
      if (K == 7) { x = __funnelshift_l(x, y, x & 7) + 1; }
      if (K == 8) { z64 = (z64 << (z64 & 7)) + 1; }
      if (K == 9) { x = (x >= 0x1000000u ? 0u : 8u) + (x >= 0x10000u ? 0u : 8u) + (x >= 0x100u ? 0u : 8u) + (x << 3) + 1; }
      if (K == 10) { x = __shfl_sync(0xffffffffu, x, 3) + 1; }
      if (K == 11) { if (x == 0xdeadbeef) { out[5] = x; x = 5; } x = x * 3 + 1; }   // never-taken branch + IMAD
      if (K == 12) { x = __byte_perm(x, y, 0x0123) + 1; }
      if (K == 13) { x = (x * y + 1); if (x & 0x40000000) out[x & 1023] = x; }   // IMAD + predicated store
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2048 + K] = x + (uint32_t)z64; cyc[K] = t1 - t0; }
}
int main() {
  uint32_t *out; long long *cyc;
  cudaMalloc(&out, 1 << 16); cudaMemset(out, 0, 1 << 16); cudaMallocManaged(&cyc, 32 * 8);
  const char *names[] = {"IMAD", "IMAD.HI+LOP", "IADD", "SHF+LOP", "2xIADD+ISETP+SEL", "CREDUX+IADD", "LOP+LDS+IADD", "LOP+SHF.funnel+IADD", "shl64 var + add64", "rc_equal_bits+shl+adds", "SHFL(const)+IADD", "ISETP+BRA(not taken)+IMAD", "PRMT+IADD", "IMAD+pred store"};
#define RUN(K) k<K><<<1, 32, 1024>>>(out, 12345, 7, cyc); k<K><<<1, 32, 1024>>>(out, 12345, 7, cyc);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13)
  cudaDeviceSynchronize();
  for (int i = 0; i < 14; i++) printf("%-28s %6.2f cycles/op-group\n", names[i], (double)cyc[i] / N / U);
  return 0;
}
