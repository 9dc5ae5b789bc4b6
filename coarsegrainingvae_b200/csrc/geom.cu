// Edge geometry: distance, unit vector, sinc radial basis x cosine envelope, once per graph
// (the reference recomputes these ~12 element-wise launches in every layer).
#include "common.cuh"

namespace cgvae {

__device__ __forceinline__ int row_of_slot(const int32_t* __restrict__ rowptr, int64_t n_rows, int slot) {
  int64_t lo = 0, hi = n_rows;  // largest row with rowptr[row] <= slot
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= slot) lo = mid; else hi = mid;
  }
  return (int)lo;
}

__global__ void __launch_bounds__(256) edge_geometry_kernel(
    const float* __restrict__ xs, const float* __restrict__ xr, const float* __restrict__ r_edge,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_recv, int64_t E,
    const float* __restrict__ coef, int R, int RB,
    float cutoff, const float* __restrict__ edge_wgt, const int32_t* __restrict__ eid, float* __restrict__ basis,
    float* __restrict__ unit) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  if (e >= rowptr[n_recv]) {
    // tail of a static-capacity graph (CUDA-graph replay): neutral records, so padded rows contribute exact zeros
    reinterpret_cast<float4*>(unit)[e] = make_float4(0.f, 0.f, 0.f, 1.f);
    for (int r = 0; r < RB; ++r) basis[e * RB + r] = 0.f;
    return;
  }
  float rx, ry, rz;
  if (r_edge != nullptr) {
    // block-level API: the caller supplies r_ij per edge in its original edge order
    const int64_t o = eid ? eid[e] : e;
    rx = r_edge[o * 3 + 0]; ry = r_edge[o * 3 + 1]; rz = r_edge[o * 3 + 2];
  } else {
    // r_ij = x_j - x_i (cgvae.py:276)
    const int i = row_of_slot(rowptr, n_recv, (int)e);
    const int j = col[e];
    rx = __fsub_rn(xs[(int64_t)j * 3 + 0], xr[(int64_t)i * 3 + 0]);
    ry = __fsub_rn(xs[(int64_t)j * 3 + 1], xr[(int64_t)i * 3 + 1]);
    rz = __fsub_rn(xs[(int64_t)j * 3 + 2], xr[(int64_t)i * 3 + 2]);
  }
  // d = sqrt(sum_c (r_c^2 + 1e-8)) (conv.py:26), fp32, unfused
  const float ax = __fadd_rn(__fmul_rn(rx, rx), 1e-8f);
  const float ay = __fadd_rn(__fmul_rn(ry, ry), 1e-8f);
  const float az = __fadd_rn(__fmul_rn(rz, rz), 1e-8f);
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(ax, ay), az));
  float4 u;
  u.x = __fdiv_rn(rx, d);
  u.y = __fdiv_rn(ry, d);
  u.z = __fdiv_rn(rz, d);
  u.w = d;
  reinterpret_cast<float4*>(unit)[e] = u;
  const bool outside = d >= cutoff;
  // CosineEnvelope modules.py:54-56: 0.5 * (cos(pi * d / cutoff) + 1), fp32 pi
  float env = 0.f;
  if (!outside) env = 0.5f * (cosf(__fdiv_rn(__fmul_rn(3.14159265358979323846f, d), cutoff)) + 1.0f);
  float wgt = 1.0f;
  if (edge_wgt != nullptr) wgt = edge_wgt[eid ? eid[e] : e];
  float* b = basis + e * RB;
  for (int r = 0; r < RB; ++r) {
    float val = 0.f;
    if (r < R) {
      // PainnRadialBasis modules.py:161-170
      if (!outside) {
        const float c = coef[r];
        const float num = (d == 0.f) ? c : sinf(__fmul_rn(c, d));
        const float den = (d == 0.f) ? 1.0f : d;
        val = __fmul_rn(__fdiv_rn(num, den), env);
      }
    } else if (r == R) {
      val = env;  // bias column of DistanceEmbed (modules.py:192-197)
    }
    b[r] = val * wgt;
  }
}

// reference layout v[N][F][3] <-> planar v[N][3][F]
__global__ void vec_to_planar_kernel(const float* __restrict__ in, int64_t N, int F, float* __restrict__ out) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over N*3*F outputs
  if (idx >= N * 3 * (int64_t)F) return;
  const int f = (int)(idx % F);
  const int c = (int)((idx / F) % 3);
  const int64_t n = idx / (3 * (int64_t)F);
  out[idx] = in[(n * F + f) * 3 + c];
}
__global__ void vec_from_planar_kernel(const float* __restrict__ in, int64_t N, int F, float* __restrict__ out) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over N*F*3 outputs
  if (idx >= N * 3 * (int64_t)F) return;
  const int c = (int)(idx % 3);
  const int f = (int)((idx / 3) % F);
  const int64_t n = idx / (3 * (int64_t)F);
  out[idx] = in[(n * 3 + c) * F + f];
}

}  // namespace cgvae

using namespace cgvae;

extern "C" {

int cgvae_edge_geometry(const float* xyz_send, const float* xyz_recv, const float* r_edge, const int32_t* rowptr,
                        const int32_t* col, int64_t n_recv, int64_t n_edges, const float* coef, int R, int RB, float cutoff,
                        const float* edge_wgt, const int32_t* eid, float* basis, float* unit, cgvae_stream_t stream) {
  if (n_edges == 0) return 0;
  CGVAE_REQUIRE(((xyz_send && xyz_recv) || r_edge) && rowptr && col && coef && basis && unit, "edge_geometry: null pointer");
  CGVAE_REQUIRE(R >= 1 && RB >= R + 1 && (RB % 4) == 0, "edge_geometry: need RB %% 4 == 0 and RB >= R+1 (R=%d RB=%d)", R, RB);
  CGVAE_REQUIRE(aligned16(unit) && aligned16(basis), "edge_geometry: outputs must be 16-byte aligned");
  launch_kernel(edge_geometry_kernel, dim3((unsigned)ceil_div(n_edges, 256)), dim3(256), 0, (cudaStream_t)stream, 
      xyz_send, xyz_recv, r_edge, rowptr, col, n_recv, n_edges, coef, R, RB, cutoff, edge_wgt, eid, basis, unit);
  return launched("edge_geometry");
}

int cgvae_vec_to_planar(const float* v_nf3, int64_t N, int F, float* v_n3f, cgvae_stream_t stream) {
  if (N == 0) return 0;
  launch_kernel(vec_to_planar_kernel, dim3((unsigned)ceil_div(N * 3 * (int64_t)F, 256)), dim3(256), 0, (cudaStream_t)stream, v_nf3, N, F, v_n3f);
  return launched("vec_to_planar");
}
int cgvae_vec_from_planar(const float* v_n3f, int64_t N, int F, float* v_nf3, cgvae_stream_t stream) {
  if (N == 0) return 0;
  launch_kernel(vec_from_planar_kernel, dim3((unsigned)ceil_div(N * 3 * (int64_t)F, 256)), dim3(256), 0, (cudaStream_t)stream, v_n3f, N, F, v_nf3);
  return launched("vec_from_planar");
}

}  // extern "C"
