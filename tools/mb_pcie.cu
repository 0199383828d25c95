// mb_pcie.cu -- design inputs for the streaming pipeline (round 2):
//   1. PCIe ceiling of the box: pinned cudaMemcpyAsync H2D alone, D2H alone, both directions at once.
//   2. zero-copy stores: a kernel that writes 32-byte records straight into pinned host memory (what the decoder's point
//      kernel does when the caller's output buffer is pinned), alone and while the copy engine runs H2D.
//   3. green contexts: can the device be split into two SM partitions (cuGreenCtxCreate), and does a kernel launched
//      into a green-context stream stay on its partition?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_pcie mb_pcie.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define DK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_ = nullptr; cuGetErrorString(r_, &s_); printf("%s: %s\n", #x, s_ ? s_ : "?"); return; } } while (0)

__global__ void __launch_bounds__(256) store_records(uint4 *dst, const uint4 *src, size_t nrec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 a = src[2 * i], b = src[2 * i + 1];
    dst[2 * i] = a; dst[2 * i + 1] = b;
  }
}
__global__ void sm_probe(uint32_t *out) {
  uint32_t v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  if (threadIdx.x == 0) out[blockIdx.x] = v;
  for (volatile int k = 0; k < 20000; k++) { }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

static void green_probe() {
  CUdevice dev; DK(cuDeviceGet(&dev, 0));
  CUdevResource all; DK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("green: device has %u SMs in its SM resource\n", all.sm.smCount);
  CUdevResource groups[2]; unsigned int ng = 1; CUdevResource rest;
  DK(cuDevSmResourceSplitByCount(&groups[0], &ng, &all, &rest, 0, 32));
  printf("green: split -> group of %u SMs, remainder %u SMs (ng %u)\n", groups[0].sm.smCount, rest.sm.smCount, ng);
  CUdevResourceDesc d0, d1; DK(cuDevResourceGenerateDesc(&d0, &groups[0], 1)); DK(cuDevResourceGenerateDesc(&d1, &rest, 1));
  CUgreenCtx g0, g1; DK(cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM)); DK(cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream s0, s1; DK(cuGreenCtxStreamCreate(&s0, g0, CU_STREAM_NON_BLOCKING, 0)); DK(cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0));
  uint32_t *d; CK(cudaMalloc(&d, 4096 * 4));
  for (int k = 0; k < 2; k++) {
    CK(cudaMemset(d, 0xFF, 4096 * 4));
    sm_probe<<<1024, 64, 0, (cudaStream_t)(k ? s1 : s0)>>>(d);
    CK(cudaStreamSynchronize((cudaStream_t)(k ? s1 : s0)));
    std::vector<uint32_t> h(1024); CK(cudaMemcpy(h.data(), d, 4096, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end()); int distinct = (int)(std::unique(h.begin(), h.end()) - h.begin());
    printf("green: stream %d -> kernel ran on %d distinct SMs (smid %u..%u)\n", k, distinct, h[0], h[distinct - 1]);
  }
  // runtime-API launch into green streams works if we got here; concurrent run of both partitions
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, (cudaStream_t)s0));
  sm_probe<<<4096, 64, 0, (cudaStream_t)s0>>>(d); sm_probe<<<4096, 64, 0, (cudaStream_t)s1>>>(d);
  CK(cudaEventRecord(e1, (cudaStream_t)s0)); CK(cudaDeviceSynchronize());
  printf("green: ok (both partitions launched)\n");
  cuStreamDestroy(s0); cuStreamDestroy(s1); cuGreenCtxDestroy(g0); cuGreenCtxDestroy(g1);
}

int main(int argc, char **argv) {
  const size_t MB = 1 << 20, bytes = (argc > 1 ? atoi(argv[1]) : 1024) * MB;
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  uint8_t *h_in, *h_out, *d_a, *d_b;
  CK(cudaMallocHost(&h_in, bytes)); CK(cudaMallocHost(&h_out, bytes));
  CK(cudaMalloc(&d_a, bytes)); CK(cudaMalloc(&d_b, bytes));
  for (size_t i = 0; i < bytes; i += 4096) { h_in[i] = (uint8_t)i; h_out[i] = 0; }
  CK(cudaMemset(d_b, 7, bytes));
  cudaStream_t s_in, s_out; CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  cudaEvent_t a0, a1, b0, b1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
  const size_t chunk = 32 * MB;
  auto h2d = [&]() { for (size_t o = 0; o < bytes; o += chunk) CK(cudaMemcpyAsync(d_a + o, h_in + o, chunk, cudaMemcpyHostToDevice, s_in)); };
  auto d2h = [&]() { for (size_t o = 0; o < bytes; o += chunk) CK(cudaMemcpyAsync(h_out + o, d_b + o, chunk, cudaMemcpyDeviceToHost, s_out)); };
  auto zc = [&](int ctas) { uint4 *hd; CK(cudaHostGetDevicePointer((void **)&hd, h_out, 0)); store_records<<<ctas, 256, 0, s_out>>>(hd, (const uint4 *)d_b, bytes / 32); };
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(a0, s_in)); h2d(); CK(cudaEventRecord(a1, s_in)); CK(cudaDeviceSynchronize());
    printf("H2D alone: %.1f GB/s\n", bytes / 1e6 / time_ms(a0, a1));
    CK(cudaEventRecord(b0, s_out)); d2h(); CK(cudaEventRecord(b1, s_out)); CK(cudaDeviceSynchronize());
    printf("D2H alone (copy engine): %.1f GB/s\n", bytes / 1e6 / time_ms(b0, b1));
    CK(cudaEventRecord(a0, s_in)); CK(cudaEventRecord(b0, s_out)); h2d(); d2h(); CK(cudaEventRecord(a1, s_in)); CK(cudaEventRecord(b1, s_out)); CK(cudaDeviceSynchronize());
    printf("both (copy engines): H2D %.1f GB/s, D2H %.1f GB/s\n", bytes / 1e6 / time_ms(a0, a1), bytes / 1e6 / time_ms(b0, b1));
    for (int ctas : {8, 32, 148, 592}) {
      CK(cudaEventRecord(b0, s_out)); zc(ctas); CK(cudaEventRecord(b1, s_out)); CK(cudaDeviceSynchronize());
      printf("zero-copy stores alone, %d CTAs: %.1f GB/s\n", ctas, bytes / 1e6 / time_ms(b0, b1));
    }
    for (int ctas : {32, 148}) {
      CK(cudaEventRecord(a0, s_in)); CK(cudaEventRecord(b0, s_out)); h2d(); zc(ctas); CK(cudaEventRecord(a1, s_in)); CK(cudaEventRecord(b1, s_out)); CK(cudaDeviceSynchronize());
      printf("H2D copy engine + zero-copy stores (%d CTAs): H2D %.1f GB/s, stores %.1f GB/s\n", ctas, bytes / 1e6 / time_ms(a0, a1), bytes / 1e6 / time_ms(b0, b1));
    }
  }
  uint64_t bad = 0; for (size_t i = 0; i < bytes; i += 4097) bad += h_out[i] != 7;
  printf("zero-copy result check: %llu mismatches\n", (unsigned long long)bad);
  // registering pageable memory (what the library does with a caller's malloc'ed buffers?)
  {
    void *p = aligned_alloc(4096, 256 * MB); memset(p, 1, 256 * MB);
    cudaEvent_t r0, r1; CK(cudaEventCreate(&r0)); CK(cudaEventCreate(&r1));
    timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    cudaError_t e = cudaHostRegister(p, 256 * MB, cudaHostRegisterDefault);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    printf("cudaHostRegister 256 MB: %s, %.2f ms\n", cudaGetErrorString(e), (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    if (e == cudaSuccess) cudaHostUnregister(p);
    free(p);
  }
  if (cuInit(0) == CUDA_SUCCESS) green_probe();
  return 0;
}
