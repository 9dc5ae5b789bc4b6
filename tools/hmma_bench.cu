// Throughput of the legacy warp-level tensor path (mma.sync m16n8k8 tf32) on sm_100a, per SM -- decides whether the
// message kernel's filter contraction (K = 12) can use it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void hmma_kernel(float* out, int iters) {
  float d[8][4];
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u};
  unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  for (int c = 0; c < 8; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
  for (int c = 0; c < 8; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_kernel(float* out, int iters) {
  float d[32];
  float a = 1.0f + threadIdx.x * 1e-9f, b = 0.5f;
  for (int c = 0; c < 32; ++c) d[c] = (float)c;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 32; ++c) d[c] = fmaf(d[c], a, b);
  }
  float s = 0.f;
  for (int c = 0; c < 32; ++c) s += d[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out;
  cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 4000, blocks = 148 * 2;
    hmma_kernel<<<blocks, warps * 32>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    hmma_kernel<<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * warps * iters * 8;
    const double flops = mmas * 16 * 8 * 8 * 2;
    printf("hmma m16n8k8 tf32: %2d warps/CTA x %d CTAs: %.3f ms, %.1f TFLOP/s, %.2f SM-cycles per mma at %d MHz nominal\n", warps,
           blocks, ms, flops / ms / 1e9, ms * 1e-3 * clk_khz * 1e3 * 148 / mmas, clk_khz / 1000);
    ffma_kernel<<<blocks, warps * 32>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ffma_kernel<<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    const double ffma = (double)blocks * warps * iters * 32;
    printf("ffma            : %2d warps/CTA x %d CTAs: %.3f ms, %.1f TFLOP/s, %.2f SM-cycles per warp-FFMA\n", warps, blocks, ms,
           ffma * 32 * 2 / ms / 1e9, ms * 1e-3 * clk_khz * 1e3 * 148 / ffma);
  }
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
