// Microbenchmark: sustained LOP3 issue rate of the whole chip (SURVEY.md section 8d asks for the
// box's LOP3 peak, since MEASURED_PEAKS.json only has HBM and bf16 GEMM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lop3_peak scripts/lop3_peak.cu && ./lop3_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_lop3(uint32_t *out, int iters, uint32_t seed) {
  uint32_t a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) a[i] = seed + threadIdx.x * 31u + i;
  uint32_t b = seed ^ 0x9e3779b9u, c = seed * 0x85ebca6bu + blockIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
      }
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) x ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  constexpr int CH = 8;
  uint32_t *d;
  cudaMalloc(&d, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_lop3<CH><<<blocks, threads>>>(d, iters, 12345u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * (double)iters * 8 * CH;
    const double rate = ops / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  printf("{\"sms\": %d, \"lop3_thread_ops_per_s\": %.4e, \"lop3_warp_instr_per_s\": %.4e, "
         "\"per_sm_per_clock_at_1965MHz\": %.2f}\n", p.multiProcessorCount, best, best / 32.0,
         best / p.multiProcessorCount / 1.965e9);
  return 0;
}
