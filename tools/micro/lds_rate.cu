// Micro-benchmark: shared-memory load issue rate per SM for 32/64/128-bit loads (conflict-free, independent loads).
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void k(float* out, int iters, int stride) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float* base = sm + (threadIdx.x >> 5) * 64;
  float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int o = ((it + u) & 7) * 1024;
      if (W == 1) { acc0 += base[o + lane * stride]; }
      if (W == 2) { float2 v = *reinterpret_cast<const float2*>(base + o + lane * 2); acc0 += v.x; acc1 += v.y; }
      if (W == 4) { float4 v = *reinterpret_cast<const float4*>(base + o + lane * 4); acc0 += v.x; acc1 += v.y; acc2 += v.z; acc3 += v.w; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int W>
void run(int threads, int stride, const char* name) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int iters = 2000;
  cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<W><<<148, threads, 96 * 1024>>>(out, iters, stride);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<W><<<148, threads, 96 * 1024>>>(out, iters, stride);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  float cyc; cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
  double warp_instr = (double)iters * 16 * (threads / 32);
  printf("%-10s threads %4d stride %d: %.0f cycles, %.3f LDS instr/clk/SM, %.1f B/clk/SM (%s)\n", name, threads, stride, cyc,
         warp_instr / cyc, warp_instr * 32 * 4 * W / cyc, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<1>(th, 1, "LDS.32"); run<2>(th, 1, "LDS.64"); run<4>(th, 1, "LDS.128");
  }
  run<1>(512, 0, "LDS.32 bc");
  run<1>(512, 2, "LDS.32 2way");
  return 0;
}
