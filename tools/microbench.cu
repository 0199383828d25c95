// Latency microbenchmarks for the serial (one-warp) kernels: cycles per dependent operation on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int K> __global__ void k(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
  uint32_t x = a + threadIdx.x * 0, y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; i++) {
    if (K == 0) { x = x * y + 1; }                                        // IMAD
    if (K == 1) { x = __umulhi(x, y) + 0x9E3779B9u; }                       // IMAD.HI (+IADD)
    if (K == 2) { x = (31 - __clz(x | 1)) + y; }                          // FLO
    if (K == 3) { x = __popc(x) + y; }                                     // POPC
    if (K == 4) { x = __shfl_sync(0xffffffffu, x, (x + 1) & 31) + 1; }     // SHFL
    if (K == 5) { x = __ballot_sync(0xffffffffu, x & 1) + y; }             // VOTE
    if (K == 6) { x = (x ^ y) + 1; }                                       // LOP+IADD
    if (K == 7) { x = x << (y & 7); x += 1; }                             // SHF + IADD
    if (K == 8) { if (x & 1) x = x * 3 + 1; else x = x >> 1; }             // data-dependent branch (collatz)
    if (K == 9) { x = out[x & 1023] + 1; }                                 // dependent global load (L1 hit)
    if (K == 10) { extern __shared__ uint32_t sm[]; x = sm[x & 255] + 1; }  // dependent LDS
    if (K == 11) { x = (x < y) ? x + 7 : x - 3; }                          // ISETP+SEL
    if (K == 12) { x = __reduce_max_sync(0xffffffffu, x) + 1; }            // REDUX
    if (K == 13) { x = __byte_perm(x, y, 0x2103) + 1; }                    // PRMT
    if (K == 14) { uint64_t z = ((uint64_t)x << 32 | y) << (x & 15); x = (uint32_t)(z >> 32) + 1; } // 64-bit shift
    if (K == 15) { x = (uint32_t)((double)x * 0.333) + 1; }                // I2F.F64, DMUL, F2I
    if (K == 16) { x = (uint32_t)__fdividef((float)x, 3.0f) + 1; }         // I2F, MUFU.RCP, FMUL, F2I
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2048 + K] = x; cyc[K] = t1 - t0; }
}
int main() {
  uint32_t *out; long long *cyc;
  cudaMalloc(&out, 1 << 16); cudaMemset(out, 0, 1 << 16); cudaMallocManaged(&cyc, 32 * 8);
  const char *names[] = {"IMAD", "IMAD.HI+IADD", "FLO+IADD", "POPC+IADD", "SHFL+IADD", "VOTE+IADD", "LOP+IADD", "SHF+IADD", "branchy collatz", "LDG(L1)+IADD", "LDS+IADD", "ISETP+SEL(2 adds)", "REDUX+IADD", "PRMT+IADD", "shl64+IADD", "I2D DMUL D2I", "I2F RCP FMUL F2I"};
#define RUN(K) k<K><<<1, 32, 1024>>>(out, 12345, 7, cyc); k<K><<<1, 32, 1024>>>(out, 12345, 7, cyc);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16)
  cudaDeviceSynchronize();
  for (int i = 0; i < 17; i++) printf("%-22s %6.1f cycles/iter\n", names[i], (double)cyc[i] / N);
  return 0;
}
