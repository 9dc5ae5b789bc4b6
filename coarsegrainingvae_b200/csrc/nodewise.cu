// Node-wise kernels: UpdateBlock element-wise stages (conv.py:588-616), bead pooling
// (scatter_mean / scatter_add, cgvae.py:297-298), embedding gather and bead -> atom lifting
// (cgvae.py:466-482, 556-576).  All deterministic (segment loops, no atomics).
#include "common.cuh"

namespace cgvae {

// x[n] = [ s[n] | sqrt(sum_c (Vv[n][c]^2 + 1e-10)) ]
__global__ void __launch_bounds__(256) update_norm_fwd_kernel(const float* __restrict__ s, const float* __restrict__ Vv, int64_t N,
                                                              int F, float* __restrict__ x) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * F) return;
  const int64_t n = idx / F;
  const int f = (int)(idx % F);
  const float* p = Vv + n * 3 * F + f;
  const float a = __fadd_rn(__fmul_rn(p[0], p[0]), 1e-10f);
  const float b = __fadd_rn(__fmul_rn(p[F], p[F]), 1e-10f);
  const float c = __fadd_rn(__fmul_rn(p[2 * (int64_t)F], p[2 * (int64_t)F]), 1e-10f);
  x[n * 2 * F + f] = s[idx];
  x[n * 2 * F + F + f] = __fsqrt_rn(__fadd_rn(__fadd_rn(a, b), c));
}

__global__ void __launch_bounds__(256) update_combine_fwd_kernel(const float* __restrict__ s, const float* __restrict__ v,
                                                                 const float* __restrict__ Uv, const float* __restrict__ Vv,
                                                                 const float* __restrict__ q, int64_t N, int F, int residual,
                                                                 float* __restrict__ s_out, float* __restrict__ v_out) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * F) return;
  const int64_t n = idx / F;
  const int f = (int)(idx % F);
  const int64_t o = n * 3 * F + f;
  const float a_vv = q[o], a_sv = q[o + F], a_ss = q[o + 2 * (int64_t)F];
  float inner = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float u = Uv[o + (int64_t)c * F], w = Vv[o + (int64_t)c * F];
    inner = fmaf(u, w, inner);
    v_out[o + (int64_t)c * F] = (residual ? v[o + (int64_t)c * F] : 0.f) + u * a_vv;
  }
  s_out[idx] = (residual ? s[idx] : 0.f) + (inner * a_sv + a_ss);
}

__global__ void __launch_bounds__(256) update_combine_bwd_kernel(const float* __restrict__ Uv, const float* __restrict__ Vv,
                                                                 const float* __restrict__ q, const float* __restrict__ g_s,
                                                                 const float* __restrict__ g_v, int64_t N, int F,
                                                                 float* __restrict__ gq, float* __restrict__ gUv,
                                                                 float* __restrict__ gVv, int64_t ldg) {
  // ldg: row pitch (floats) of gUv / gVv viewed as [3N][.]: F for two separate tensors, 2F when they are the two column
  // halves of ONE [3N][2F] matrix (the input gradient through u_mat and v_mat is then a single contraction over 2F)
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * F) return;
  const int64_t n = idx / F;
  const int f = (int)(idx % F);
  const int64_t o = n * 3 * F + f;
  const int64_t og = n * 3 * ldg + f;
  const float a_vv = q[o], a_sv = q[o + F];
  const float gs = g_s[idx];
  float inner = 0.f, g_avv = 0.f;
  const float t = gs * a_sv;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float u = Uv[o + (int64_t)c * F], w = Vv[o + (int64_t)c * F], g = g_v[o + (int64_t)c * F];
    inner = fmaf(u, w, inner);
    g_avv = fmaf(g, u, g_avv);
    gUv[og + (int64_t)c * ldg] = g * a_vv + t * w;
    gVv[og + (int64_t)c * ldg] = t * u;
  }
  gq[o] = g_avv;
  gq[o + F] = gs * inner;
  gq[o + 2 * (int64_t)F] = gs;
}

__global__ void __launch_bounds__(256) update_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ Vv,
                                                              const float* __restrict__ gx, const float* __restrict__ g_s,
                                                              int64_t N, int F, int residual, float* __restrict__ gs_in,
                                                              float* __restrict__ gVv, int64_t ldg) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * F) return;
  const int64_t n = idx / F;
  const int f = (int)(idx % F);
  gs_in[idx] = (residual ? g_s[idx] : 0.f) + gx[n * 2 * F + f];
  const float scale = gx[n * 2 * F + F + f] / x[n * 2 * F + F + f];
  const int64_t o = n * 3 * F + f, og = n * 3 * ldg + f;
#pragma unroll
  for (int c = 0; c < 3; ++c) gVv[og + (int64_t)c * ldg] = fmaf(scale, Vv[o + (int64_t)c * F], gVv[og + (int64_t)c * ldg]);
}

// out[b][w] = sum over the bead's atoms (ascending) of X[a][w], optionally / max(count, 1)
__global__ void __launch_bounds__(256) segment_reduce_fwd_kernel(const float* __restrict__ X, const int32_t* __restrict__ rowptr_b,
                                                                 const int32_t* __restrict__ atoms, int64_t W, int mean,
                                                                 float* __restrict__ out) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t b = blockIdx.x;
  const int64_t w = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int beg = rowptr_b[b], end = rowptr_b[b + 1];
  float acc = 0.f;
  for (int t = beg; t < end; ++t) acc += X[(int64_t)atoms[t] * W + w];
  if (mean) acc = acc / (float)max(end - beg, 1);
  out[b * W + w] = acc;
}
__global__ void __launch_bounds__(256) segment_reduce_bwd_kernel(const float* __restrict__ g_out, const int64_t* __restrict__ mapping,
                                                                 const int32_t* __restrict__ rowptr_b, int64_t W, int mean,
                                                                 float* __restrict__ g_X) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t a = blockIdx.x;
  const int64_t w = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int64_t b = mapping[a];
  float g = g_out[b * W + w];
  if (mean) g = g / (float)max(rowptr_b[b + 1] - rowptr_b[b], 1);
  g_X[a * W + w] = g;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t W,
                                                          int64_t n_rows, float* __restrict__ out, int32_t* __restrict__ err) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t n = blockIdx.x;
  const int64_t w = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (w >= W) return;
  const int64_t r = idx[n];
  if (n_rows > 0 && (r < 0 || r >= n_rows)) {                  // nn.Embedding raises IndexError here
    if (w == 0) raise_flag(err, CGVAE_ERR_EMBED_INDEX);
    out[n * W + w] = __int_as_float(0x7fc00000);
    return;
  }
  out[n * W + w] = table[r * W + w];
}

// warp per bead.  mode 0: plain gather; 1: subtract the bead mean; 2: zero pinned atoms
__global__ void __launch_bounds__(128) lift_fwd_kernel(const float* __restrict__ V, const float* __restrict__ cg_xyz,
                                                       const int64_t* __restrict__ rank, const int32_t* __restrict__ rowptr_b,
                                                       const int32_t* __restrict__ atoms, const uint8_t* __restrict__ pin,
                                                       int64_t n_beads, int F, int mode, float* __restrict__ xyz_out,
                                                       int32_t* __restrict__ err) {
  CGVAE_KERNEL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= n_beads) return;
  const int beg = rowptr_b[b], end = rowptr_b[b + 1];
  const float* Vb = V + b * 3 * F;
  float mean[3] = {0.f, 0.f, 0.f};
  if (mode == 1) {
    for (int t = beg + lane; t < end; t += 32) {
      const int64_t ch = rank[atoms[t]];
      if (ch >= F) continue;                                   // more atoms in the bead than channels: flagged below
#pragma unroll
      for (int c = 0; c < 3; ++c) mean[c] += Vb[(int64_t)c * F + ch];
    }
    const float inv = 1.0f / (float)max(end - beg, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) mean[c] = warp_sum(mean[c]) * inv;
  }
  for (int t = beg + lane; t < end; t += 32) {
    const int a = atoms[t];
    const int64_t ch = rank[a];
    const bool pinned = (mode == 2) && pin != nullptr && pin[a] != 0;
    if (ch >= F) {
      // the reference raises IndexError at cg_v[mapping, CG2atomChannel] (cgvae.py:473): no read, NaN out, flag
      raise_flag(err, CGVAE_ERR_LIFT_RANK);
#pragma unroll
      for (int c = 0; c < 3; ++c) xyz_out[(int64_t)a * 3 + c] = __int_as_float(0x7fc00000);
      continue;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float rel = pinned ? 0.f : Vb[(int64_t)c * F + ch] - mean[c];
      xyz_out[(int64_t)a * 3 + c] = rel + cg_xyz[b * 3 + c];
    }
  }
}
__global__ void __launch_bounds__(128) lift_bwd_kernel(const float* __restrict__ g_xyz, const int64_t* __restrict__ rank,
                                                       const int32_t* __restrict__ rowptr_b, const int32_t* __restrict__ atoms,
                                                       const uint8_t* __restrict__ pin, int64_t n_beads, int F, int mode,
                                                       float* __restrict__ g_V) {
  CGVAE_KERNEL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= n_beads) return;
  const int beg = rowptr_b[b], end = rowptr_b[b + 1];
  float* gVb = g_V + b * 3 * F;
  float mean[3] = {0.f, 0.f, 0.f};
  if (mode == 1) {
    for (int t = beg + lane; t < end; t += 32) {
      const int a = atoms[t];
#pragma unroll
      for (int c = 0; c < 3; ++c) mean[c] += g_xyz[(int64_t)a * 3 + c];
    }
    const float inv = 1.0f / (float)max(end - beg, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) mean[c] = warp_sum(mean[c]) * inv;
  }
  for (int t = beg + lane; t < end; t += 32) {
    const int a = atoms[t];
    const int64_t ch = rank[a];
    const bool pinned = (mode == 2) && pin != nullptr && pin[a] != 0;
    if (ch >= F) continue;                                     // flagged by lift_fwd_kernel; never write past the bead row
#pragma unroll
    for (int c = 0; c < 3; ++c) gVb[(int64_t)c * F + ch] = pinned ? 0.f : g_xyz[(int64_t)a * 3 + c] - mean[c];
  }
}

// PCN C-alpha pin mask (cgvae.py:569-571: `if ca_idx[-1] < n_atoms: xyz_rel[ca_idx] = 0`) with the predicate evaluated on
// the device: no host read of ca_idx[-1], so the PCN step can be captured in a CUDA graph
__global__ void __launch_bounds__(256) pin_mask_kernel(const int64_t* __restrict__ idx, int64_t n_idx, const int64_t* __restrict__ n_live,
                                                       int64_t n_atoms, uint8_t* __restrict__ pin, int32_t* __restrict__ err) {
  CGVAE_KERNEL_PROLOGUE();
  const int64_t n = n_live ? min(*n_live, n_idx) : n_idx;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || n <= 0) return;
  if (idx[n - 1] >= n_atoms) return;                 // the reference skips the re-anchoring altogether
  const int64_t a = idx[i];
  if (a < 0 || a >= n_atoms) {
    raise_flag(err, CGVAE_ERR_NODE_INDEX);
    return;
  }
  pin[a] = 1;
}

}  // namespace cgvae

using namespace cgvae;

extern "C" {

int cgvae_pin_mask(const int64_t* idx, int64_t n_idx, const int64_t* n_live, int64_t n_atoms, uint8_t* pin, size_t pin_bytes,
                   cgvae_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n_atoms == 0) return 0;
  CGVAE_REQUIRE(pin && pin_bytes >= (size_t)n_atoms && pin_bytes % 4 == 0, "pin_mask: pin buffer must hold n_atoms bytes, padded to 4");
  CGVAE_ZERO(pin, pin_bytes, st);
  if (n_idx == 0) return 0;
  CGVAE_REQUIRE(idx, "pin_mask: null index list");
  launch_kernel(pin_mask_kernel, dim3((unsigned)ceil_div(n_idx, 256)), dim3(256), 0, st, idx, n_idx, n_live, n_atoms, pin, err_flags());
  return launched("pin_mask");
}

int cgvae_update_norm_fwd(const float* s, const float* Vv, int64_t N, int F, float* x, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(s && Vv && x, "update_norm_fwd: null pointer");
  launch_kernel(update_norm_fwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, s, Vv, N, F, x);
  return launched("update_norm_fwd");
}
int cgvae_update_combine_fwd(const float* s, const float* v, const float* Uv, const float* Vv, const float* q, int64_t N, int F,
                             int residual, float* s_out, float* v_out, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(Uv && Vv && q && s_out && v_out && (!residual || (s && v)), "update_combine_fwd: null pointer");
  launch_kernel(update_combine_fwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, s, v, Uv, Vv, q, N, F, residual, s_out,
                                                                                             v_out);
  return launched("update_combine_fwd");
}
int cgvae_update_combine_bwd(const float* Uv, const float* Vv, const float* q, const float* g_s, const float* g_v, int64_t N, int F,
                             float* gq, float* gUv, float* gVv, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(Uv && Vv && q && g_s && g_v && gq && gUv && gVv, "update_combine_bwd: null pointer");
  launch_kernel(update_combine_bwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, Uv, Vv, q, g_s, g_v, N, F, gq, gUv, gVv, (int64_t)F);
  return launched("update_combine_bwd");
}
int cgvae_update_norm_bwd(const float* x, const float* Vv, const float* gx, const float* g_s, int64_t N, int F, int residual,
                          float* gs_in, float* gVv, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(x && Vv && gx && gs_in && gVv && (!residual || g_s), "update_norm_bwd: null pointer");
  launch_kernel(update_norm_bwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, x, Vv, gx, g_s, N, F, residual, gs_in, gVv, (int64_t)F);
  return launched("update_norm_bwd");
}

int cgvae_update_combine_bwd_ld(const float* Uv, const float* Vv, const float* q, const float* g_s, const float* g_v, int64_t N, int F,
                                float* gq, float* gUv, float* gVv, int64_t ldg, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(Uv && Vv && q && g_s && g_v && gq && gUv && gVv && ldg >= F, "update_combine_bwd_ld: bad arguments");
  launch_kernel(update_combine_bwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, Uv, Vv, q, g_s, g_v, N, F, gq,
                gUv, gVv, ldg);
  return launched("update_combine_bwd");
}
int cgvae_update_norm_bwd_ld(const float* x, const float* Vv, const float* gx, const float* g_s, int64_t N, int F, int residual,
                             float* gs_in, float* gVv, int64_t ldg, cgvae_stream_t stream) {
  if (N == 0) return 0;
  CGVAE_REQUIRE(x && Vv && gx && gs_in && gVv && (!residual || g_s) && ldg >= F, "update_norm_bwd_ld: bad arguments");
  launch_kernel(update_norm_bwd_kernel, dim3((unsigned)ceil_div(N * F, 256)), dim3(256), 0, (cudaStream_t)stream, x, Vv, gx, g_s, N, F, residual,
                gs_in, gVv, ldg);
  return launched("update_norm_bwd");
}

int cgvae_segment_reduce_fwd(const float* X, const int32_t* rowptr_b, const int32_t* atoms, int64_t n_beads, int64_t W, int mean,
                             float* out, cgvae_stream_t stream) {
  if (n_beads == 0 || W == 0) return 0;
  CGVAE_REQUIRE(X && rowptr_b && atoms && out, "segment_reduce_fwd: null pointer");
  dim3 grid((unsigned)n_beads, (unsigned)ceil_div(W, 256));
  launch_kernel(segment_reduce_fwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, X, rowptr_b, atoms, W, mean, out);
  return launched("segment_reduce_fwd");
}
int cgvae_segment_reduce_bwd(const float* g_out, const int64_t* mapping, const int32_t* rowptr_b, int64_t N, int64_t W, int mean,
                             float* g_X, cgvae_stream_t stream) {
  if (N == 0 || W == 0) return 0;
  CGVAE_REQUIRE(g_out && mapping && rowptr_b && g_X, "segment_reduce_bwd: null pointer");
  dim3 grid((unsigned)N, (unsigned)ceil_div(W, 256));
  launch_kernel(segment_reduce_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, g_out, mapping, rowptr_b, W, mean, g_X);
  return launched("segment_reduce_bwd");
}
int cgvae_gather_rows(const float* table, const int64_t* idx, int64_t N, int64_t W, int64_t n_rows, float* out,
                      cgvae_stream_t stream) {
  if (N == 0 || W == 0) return 0;
  CGVAE_REQUIRE(table && idx && out, "gather_rows: null pointer");
  dim3 grid((unsigned)N, (unsigned)ceil_div(W, 256));
  launch_kernel(gather_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, table, idx, W, n_rows, out, err_flags());
  return launched("gather_rows");
}

int cgvae_lift_fwd(const float* V, const float* cg_xyz, const int64_t* mapping, const int64_t* rank, const int32_t* rowptr_b,
                   const int32_t* atoms, const uint8_t* pin, int64_t N, int64_t n_beads, int F, int mode, float* xyz_out,
                   cgvae_stream_t stream) {
  (void)mapping;
  if (N == 0 || n_beads == 0) return 0;
  CGVAE_REQUIRE(V && cg_xyz && rank && rowptr_b && atoms && xyz_out, "lift_fwd: null pointer");
  CGVAE_REQUIRE(mode >= 0 && mode <= 2, "lift_fwd: bad mode %d", mode);
  launch_kernel(lift_fwd_kernel, dim3((unsigned)ceil_div(n_beads, 4)), dim3(128), 0, (cudaStream_t)stream, V, cg_xyz, rank, rowptr_b, atoms, pin, n_beads, F,
                                                                                   mode, xyz_out, err_flags());
  return launched("lift_fwd");
}
int cgvae_lift_bwd(const float* g_xyz, const int64_t* mapping, const int64_t* rank, const int32_t* rowptr_b, const int32_t* atoms,
                   const uint8_t* pin, int64_t N, int64_t n_beads, int F, int mode, float* g_V, cgvae_stream_t stream) {
  (void)mapping;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_beads == 0) return 0;
  CGVAE_REQUIRE(g_xyz && rank && rowptr_b && atoms && g_V, "lift_bwd: null pointer");
  CGVAE_ZERO(g_V, sizeof(float) * (size_t)n_beads * 3 * (size_t)F, st);
  if (N == 0) return 0;
  launch_kernel(lift_bwd_kernel, dim3((unsigned)ceil_div(n_beads, 4)), dim3(128), 0, st, g_xyz, rank, rowptr_b, atoms, pin, n_beads, F, mode, g_V);
  return launched("lift_bwd");
}

}  // extern "C"
