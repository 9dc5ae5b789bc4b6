// Sample-quality metrics of the ensemble-sampling loop (scripts/sampling.py:120-194 get_bond_graphs / count_valid_graphs,
// :220-239 compute_rmsd, :324-333 eval_sample_qualities): bond graphs from covalent-radius cutoffs on the dense distance
// matrix, their difference to the reference conformation's graph, and RMSDs -- for all S samples of an ensemble in two
// launches, instead of 4 S dense [N,N] torch passes on the host.
//
// Bit-exact adjacency: the reference evaluates, in fp32, dist = sqrt(((dx^2 + dy^2) + dz^2)) and bond = dist < (r_i + r_j) * scale
// with the diagonal cleared; the same unfused operations in the same order are used here (no FMA contraction, IEEE sqrt).
#include "common.cuh"

namespace cgvae {

__device__ __forceinline__ bool bonded(const float* __restrict__ xyz, const float* __restrict__ radius, int i, int j, float scale) {
  const float dx = __fsub_rn(xyz[3 * i + 0], xyz[3 * j + 0]);
  const float dy = __fsub_rn(xyz[3 * i + 1], xyz[3 * j + 1]);
  const float dz = __fsub_rn(xyz[3 * i + 2], xyz[3 * j + 2]);
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float cut = __fmul_rn(__fadd_rn(radius[j], radius[i]), scale);      // (vdw[None, :] + vdw[:, None]) * scale
  return (i != j) && (__fsqrt_rn(d2) < cut);
}

// bond[i][j] (uint8) of one conformation
__global__ void __launch_bounds__(256) bond_graph_kernel(const float* __restrict__ xyz, const float* __restrict__ radius, int n, float scale,
                                                         uint8_t* __restrict__ bond) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * n) return;
  const int i = (int)(idx / n), j = (int)(idx % n);
  bond[idx] = bonded(xyz, radius, i, j, scale) ? 1 : 0;
}

// counts[s] = { #(gen != ref), sum(ref - gen), sum(ref) } over all ordered pairs, and the same over heavy-atom pairs
__global__ void __launch_bounds__(256) graph_diff_kernel(const float* __restrict__ ref_xyz, const float* __restrict__ samples,
                                                         const float* __restrict__ radius, const uint8_t* __restrict__ heavy, int n,
                                                         float scale, int32_t* __restrict__ counts) {
  CGVAE_KERNEL_PROLOGUE();
  const int s = blockIdx.y;
  const float* xyz = samples + (int64_t)s * n * 3;
  int c[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (int64_t)n * n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const int r = bonded(ref_xyz, radius, i, j, scale) ? 1 : 0;
    const int g = bonded(xyz, radius, i, j, scale) ? 1 : 0;
    const int hv = (heavy[i] && heavy[j]) ? 1 : 0;
    c[0] += (r != g);
    c[1] += r - g;
    c[2] += r;
    c[3] += hv * (r != g);
    c[4] += hv * (r - g);
    c[5] += hv * r;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    int v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(&counts[s * 6 + k], v);     // integer sums: exact in any order
  }
}

// rmsd[s] = { sqrt(mean_i |x_i - ref_i|^2) over all atoms, over heavy atoms }; one block per sample, fixed order
__global__ void __launch_bounds__(256) rmsd_kernel(const float* __restrict__ ref_xyz, const float* __restrict__ samples,
                                                   const uint8_t* __restrict__ heavy, int n, float* __restrict__ rmsd) {
  CGVAE_KERNEL_PROLOGUE();
  __shared__ double red[2][8];
  const int s = blockIdx.x;
  const float* xyz = samples + (int64_t)s * n * 3;
  double a = 0.0, h = 0.0;
  int nh = 0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const double dx = (double)xyz[3 * i] - (double)ref_xyz[3 * i], dy = (double)xyz[3 * i + 1] - (double)ref_xyz[3 * i + 1],
                 dz = (double)xyz[3 * i + 2] - (double)ref_xyz[3 * i + 2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    a += d2;
    if (heavy[i]) h += d2;
  }
  for (int i = threadIdx.x; i < n; i += 256) nh += heavy[i] ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    h += __shfl_xor_sync(0xffffffffu, h, o);
    nh += __shfl_xor_sync(0xffffffffu, nh, o);
  }
  __shared__ int nh_sh[8];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = h;
    nh_sh[threadIdx.x >> 5] = nh;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, th = 0.0;
    int tn = 0;
    for (int w = 0; w < 8; ++w) {
      ta += red[0][w];
      th += red[1][w];
      tn += nh_sh[w];
    }
    rmsd[2 * s + 0] = (float)sqrt(ta / (double)max(n, 1));
    rmsd[2 * s + 1] = (float)sqrt(th / (double)max(tn, 1));
  }
}

}  // namespace cgvae

using namespace cgvae;

extern "C" {

int cgvae_bond_graph(const float* xyz, const float* radius, int64_t n, float scale, uint8_t* bond, cgvae_stream_t stream) {
  if (n == 0) return 0;
  CGVAE_REQUIRE(xyz && radius && bond && n < 46341, "bond_graph: bad arguments");
  launch_kernel(bond_graph_kernel, dim3((unsigned)ceil_div(n * n, 256)), dim3(256), 0, (cudaStream_t)stream, xyz, radius, (int)n, scale, bond);
  return launched("bond_graph");
}

int cgvae_sample_quality(const float* ref_xyz, const float* samples, const float* radius, const uint8_t* heavy, int64_t n,
                         int64_t n_samples, float scale, int32_t* counts, float* rmsd, cgvae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0 || n_samples == 0) return 0;
  CGVAE_REQUIRE(ref_xyz && samples && radius && heavy && counts && rmsd && n < 46341 && n_samples < 65536, "sample_quality: bad arguments");
  CGVAE_ZERO(counts, sizeof(int32_t) * 6 * (size_t)n_samples, st);
  const unsigned bx = (unsigned)std::min<int64_t>(ceil_div(n * n, 256), std::max<int64_t>(1, (int64_t)kNumSM * 8 / n_samples));
  launch_kernel(graph_diff_kernel, dim3(bx, (unsigned)n_samples), dim3(256), 0, st, ref_xyz, samples, radius, heavy, (int)n, scale, counts);
  if (int rc = launched("graph_diff")) return rc;
  launch_kernel(rmsd_kernel, dim3((unsigned)n_samples), dim3(256), 0, st, ref_xyz, samples, heavy, (int)n, rmsd);
  return launched("rmsd");
}

}  // extern "C"
