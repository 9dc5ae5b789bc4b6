// Microbenchmark: issue and completion time of sequences of tcgen05.mma (M = 128, one CTA) on one SM as a function of N,
// kind (tf32 K=8 / f16 K=16), A operand source and ISSUE STYLE:
//   style 0: `if (lane == 0)` divergent branch        style 1: `if (elect_one())` in a converged warp
//   style 2: one asm block `elect.sync; @p tcgen05.mma` (no C++ branch)
// Decides the batch shape of the tensor-core message kernels (csrc/message_tc.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I coarsegrainingvae_b200/csrc -o tools/bin/tc_latency tools/tc_latency.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace cgvae::tcx;

__device__ __forceinline__ void mma_pred(uint32_t d, uint64_t a, uint64_t b, uint32_t acc, uint32_t idesc, int f16) {
  if (f16)
    asm volatile("{\n.reg .pred pe, pa;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %4, 0;\n@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred pe, pa;\nelect.sync _|pe, 0xffffffff;\nsetp.ne.b32 pa, %4, 0;\n@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_pred(uint64_t* bar) {
  asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
               ::"r"(smem_u32(bar)) : "memory");
}

template <int STYLE>
__global__ void __launch_bounds__(128, 1) lat_kernel(int N, int n_mma, int chains, int f16, int reps, long long* out) {
  extern __shared__ __align__(1024) char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 16 + 256 * 16); i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 0 && (STYLE == 0 ? tid == 0 : true)) {
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 128 * 16 * 4;
    // f16: D fp32, A/B f16 (format 0), K = 16
    const uint32_t idesc = f16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24)) : idesc_tf32(128, N);
    long long best = 1ll << 60, total = 0, best_issue = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (STYLE == 2) {
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t d = tm + (uint32_t)((i % chains) * N);
          const uint32_t ko = (uint32_t)(i & 1) * 256;
          mma_pred(d, make_desc(a_s + ko, 128, 512), make_desc(b_s + ko, 128, 512), i >= chains ? 1u : 0u, idesc, f16);
        }
      } else if (STYLE == 0 || elect_one()) {
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t d = tm + (uint32_t)((i % chains) * N);
          const uint32_t ko = (uint32_t)(i & 1) * 256;
          if (f16)
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(d), "l"(make_desc(a_s + ko, 128, 512)), "l"(make_desc(b_s + ko, 128, 512)), "r"(idesc), "r"(i >= chains ? 1u : 0u) : "memory");
          else
            umma_tf32_ss(d, make_desc(a_s + ko, 128, 512), make_desc(b_s + ko, 128, 512), i >= chains ? 1u : 0u, idesc);
        }
      }
      const long long t1 = clock64();
      if (STYLE == 2) commit_pred(&bar);
      else if (STYLE == 0 || elect_one()) umma_commit(&bar);
      if (STYLE != 0) __syncwarp();
      mbar_wait(&bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      tc_fence_after();
      if (t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
      total += t2 - t0;
    }
    if (tid == 0) {
      out[0] = best;
      out[1] = best_issue;
      out[2] = total / reps;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

// operand-layout sweep (round 2, gemm_tc.cu): M = 128, N = 128, tf32, SS form, descriptors of the swizzled layouts the GEMM uses
//   lay 0: A, B K-major SWIZZLE_128B (k-advance 32 B inside the swizzle row)      lay 1: A K-major, B MN-major SWIZZLE_128B_BASE32B
//   lay 2: A, B MN-major                                                           lay 3: A, B K-major SWIZZLE_NONE (round-1 layout)
__device__ __forceinline__ uint64_t desc_lay(uint32_t addr, bool kmajor, bool swz) {
  const uint32_t lbo = !swz ? 128u : (kmajor ? 16u : 4096u), sbo = !swz ? 1024u : (kmajor ? 1024u : 512u);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(!swz ? 0 : (kmajor ? 2 : 1)) << 61;
  return d;
}
__global__ void __launch_bounds__(128, 1) lay_kernel(int lay, int n_mma, int reps, long long* out) {
  extern __shared__ __align__(1024) char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 128 * 32; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 0) {
    const bool a_k = (lay == 0 || lay == 1 || lay == 3), b_k = (lay == 0 || lay == 3), swz = lay != 3;
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 128 * 32 * 4;
    const uint32_t idesc = idesc_tf32(128, 128, !a_k, !b_k);
    const uint32_t a_step = !swz ? 256u : (a_k ? 32u : 1024u), b_step = !swz ? 256u : (b_k ? 32u : 1024u);
    long long best = 1ll << 60;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t ks = (uint32_t)(i & 3);
          umma_tf32_ss(tm + (uint32_t)((i % 3) * 128), desc_lay(a_s + ks * a_step, a_k, swz), desc_lay(b_s + ks * b_step, b_k, swz),
                       i >= 3 ? 1u : 0u, idesc);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      tc_fence_after();
      if (t2 - t0 < best) best = t2 - t0;
    }
    if (tid == 0) out[0] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  cudaFuncSetAttribute(lay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  printf("operand layout sweep, 128x128x8 tf32 SS MMAs round-robin over 3 accumulators\n%-4s %-6s %10s %12s\n", "lay", "n_mma", "best_cyc", "cyc_per_mma");
  for (int lay = 0; lay < 4; ++lay)
    for (int n_mma : {12, 48, 192}) {
      lay_kernel<<<1, 128, 40 * 1024>>>(lay, n_mma, 20, out);
      cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h;
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("%-4d %-6d %10lld %12.1f\n", lay, n_mma, h, (double)h / n_mma);
    }
  printf("%-6s %-5s %-4s %-6s %-6s %10s %10s %10s %12s\n", "style", "kind", "N", "n_mma", "chains", "best_cyc", "issue_cyc", "avg_cyc", "cyc_per_mma");
  for (int style = 0; style < 3; ++style)
    for (int f16 = 0; f16 < 2; ++f16)
      for (int N : {32, 128, 256})
        for (int n_mma : {1, 4, 16, 64}) {
          const int chains = 1;
          if (style == 0) lat_kernel<0><<<1, 128, 40 * 1024>>>(N, n_mma, chains, f16, 30, out);
          if (style == 1) lat_kernel<1><<<1, 128, 40 * 1024>>>(N, n_mma, chains, f16, 30, out);
          if (style == 2) lat_kernel<2><<<1, 128, 40 * 1024>>>(N, n_mma, chains, f16, 30, out);
          cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[3];
          cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
          printf("%-6d %-5s %-4d %-6d %-6d %10lld %10lld %10lld %12.1f\n", style, f16 ? "f16" : "tf32", N, n_mma, chains, h[0], h[1], h[2],
                 (double)h[0] / n_mma);
        }
  return 0;
}
