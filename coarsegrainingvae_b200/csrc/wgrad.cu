// Grouped parameter-gradient contraction for small node counts (decoder graphs of the molecule configs: 12 beads).
//
// dW[n_out][n_in] = gy[rows][n_out]^T x[rows][n_in] with rows = 12..96 is an outer-product accumulation whose cost is
// WRITING dW: at the chignolin config 55 M weight gradients (222 MB) per step.  As one launch per Dense layer (92 TN
// GEMMs of 1.4..13 MB each + 56 bias column sums) the step spent 1.2 ms on them -- every launch is a single wave of
// CTAs whose life is one load -> 12 FMA -> store latency chain (180 GB/s of stores).  The weight gradients do not feed
// the rest of the backward pass, so the host defers them (ops.py: deferred list) and this kernel produces ALL of them
// in one launch per 64 problems: the grid is the concatenation of the 64x256 output strips of every problem, thousands
// of CTAs are in flight and the stores stream at HBM rate.  The problem table travels in the kernel parameters
// (__grid_constant__, 3.9 KB): no table memory, no copies, CUDA-graph capturable.
//
// Summation over rows is sequential in row order: deterministic, independent of the grouping.
#include "common.cuh"

namespace cgvae {

constexpr int WG_MAXP = 48;   // problems per launch (72-byte entries: the table stays below the 4 KB parameter space)
constexpr int WG_ROWS = 32;   // rows staged per pass
constexpr int WG_T = 64;      // output tile edge
constexpr int WG_STRIP = 4;   // output tiles along n_in per CTA
constexpr int WG_FAST_ROWS = 16;   // up to this many rows the whole strip of x is staged at once

struct WgradBatch {
  cgvae_wgrad_problem p[WG_MAXP];
  int tile_begin[WG_MAXP + 1];
  int n;
};
static_assert(sizeof(WgradBatch) <= 4000, "problem table must fit the 4 KB kernel parameter space");

// Row R of an operand: one [rows x cols] matrix, or -- seg_rows > 0 -- the row-wise concatenation of several such
// matrices that live seg_stride elements apart (the all-gathered factors of the data-parallel ranks: dW = sum over ranks
// of gy_r^T x_r is ONE contraction over world * rows rows, in rank order).
__device__ __forceinline__ void stage_rows(float (*dst)[WG_T], const float* __restrict__ src, int ld, int r0, int rows,
                                           int c0, int ncols, bool vec, int tid, int seg_rows, int64_t seg_stride) {
  // dst[r][c] = row (r0+r), column c0 + c for r0 + r < rows (at most WG_ROWS of them), c < 64; zero for c0 + c >= ncols
  const int rn = min(WG_ROWS, rows - r0);
  for (int idx = tid; idx < rn * (WG_T / 4); idx += 256) {
    const int r = idx / (WG_T / 4), c = 4 * (idx % (WG_T / 4));
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      const int R = r0 + r;
      const float* q;
      if (seg_rows > 0) {
        const int sgm = R / seg_rows;
        q = src + (int64_t)sgm * seg_stride + (int64_t)(R - sgm * seg_rows) * ld + c0 + c;
      } else {
        q = src + (int64_t)R * ld + c0 + c;
      }
      if (vec && c0 + c + 3 < ncols) {
        val = __ldg(reinterpret_cast<const float4*>(q));
      } else {
        if (c0 + c + 0 < ncols) val.x = __ldg(q + 0);
        if (c0 + c + 1 < ncols) val.y = __ldg(q + 1);
        if (c0 + c + 2 < ncols) val.z = __ldg(q + 2);
        if (c0 + c + 3 < ncols) val.w = __ldg(q + 3);
      }
    }
    *reinterpret_cast<float4*>(&dst[r][c]) = val;
  }
}

__global__ void __launch_bounds__(256) wgrad_grouped_kernel(const __grid_constant__ WgradBatch batch) {
  CGVAE_KERNEL_PROLOGUE();
  __shared__ __align__(16) float gs[WG_ROWS][WG_T];
  __shared__ __align__(16) float xs[WG_ROWS][WG_T];
  const int tid = threadIdx.x, bid = blockIdx.x;
  int lo = 0, hi = batch.n;                     // tile_begin[lo] <= bid < tile_begin[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (batch.tile_begin[mid] <= bid) lo = mid; else hi = mid;
  }
  const cgvae_wgrad_problem& P = batch.p[lo];
  const float* __restrict__ gy = P.gy;
  const float* __restrict__ x = P.x;
  float* __restrict__ dW = P.dW;
  float* __restrict__ db = P.db;
  const int rows = P.rows, n_out = P.n_out, n_in = P.n_in, ldg = P.ldg, ldx = P.ldx, seg_rows = P.seg_rows;
  const int64_t sg = P.seg_stride_g, sx = P.seg_stride_x;
  const bool has_w = (dW != nullptr);
  // a CTA owns a 64 x (WG_STRIP * 64) strip of dW: WG_STRIP output tiles along n_in share one staged gy tile (13 600
  // one-tile CTAs per chignolin step were bound by CTA turnover, not by the stores)
  const int strips_k = has_w ? (n_in + WG_STRIP * WG_T - 1) / (WG_STRIP * WG_T) : 1;
  const int t = bid - batch.tile_begin[lo];
  const int n0 = (t / strips_k) * WG_T, kstrip = (t % strips_k) * WG_STRIP * WG_T;
  const bool g_vec = ((reinterpret_cast<uintptr_t>(gy) & 15) == 0) && (ldg % 4 == 0) && (seg_rows == 0 || sg % 4 == 0);
  const bool x_vec = has_w && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0) && (seg_rows == 0 || sx % 4 == 0);
  const bool w_vec = has_w && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0) && (n_in % 4 == 0);
  const int ty = tid >> 4, tx = tid & 15;       // outputs n0 + 4*ty + a, k0 + 4*tx + b
  const bool one_chunk = rows <= WG_ROWS;       // the gy tile is staged once for the whole strip

  if (one_chunk) stage_rows(gs, gy, ldg, 0, rows, n0, n_out, g_vec, tid, seg_rows, sg);
  if (db != nullptr && kstrip == 0) {
    // bias gradient: column sums of gy, rows in order
    float bsum = 0.f;
    for (int r0 = 0; r0 < rows; r0 += WG_ROWS) {
      if (!one_chunk) {
        __syncthreads();
        stage_rows(gs, gy, ldg, r0, rows, n0, n_out, g_vec, tid, seg_rows, sg);
      }
      __syncthreads();
      const int rn = min(WG_ROWS, rows - r0);
      if (tid < WG_T)
        for (int r = 0; r < rn; ++r) bsum += gs[r][tid];
    }
    if (tid < WG_T && n0 + tid < n_out) db[n0 + tid] = bsum;
  }
  if (!has_w) return;
  if (rows <= WG_FAST_ROWS) {
    // few rows (12-bead layers): the x columns of ALL WG_STRIP tiles of the strip are staged at once -- one global-load latency
    // per CTA instead of one per tile (each tile was a load -> sync -> 12 FMA -> store chain)
    __shared__ __align__(16) float xw[WG_FAST_ROWS][WG_STRIP * WG_T];
    const int ncol4 = WG_STRIP * WG_T / 4;
    for (int idx = tid; idx < rows * ncol4; idx += 256) {
      const int r = idx / ncol4, c = 4 * (idx % ncol4);
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* q;
      if (seg_rows > 0) {
        const int sgm = r / seg_rows;
        q = x + (int64_t)sgm * sx + (int64_t)(r - sgm * seg_rows) * ldx + kstrip + c;
      } else {
        q = x + (int64_t)r * ldx + kstrip + c;
      }
      if (x_vec && kstrip + c + 3 < n_in) {
        val = __ldg(reinterpret_cast<const float4*>(q));
      } else {
        if (kstrip + c + 0 < n_in) val.x = __ldg(q + 0);
        if (kstrip + c + 1 < n_in) val.y = __ldg(q + 1);
        if (kstrip + c + 2 < n_in) val.z = __ldg(q + 2);
        if (kstrip + c + 3 < n_in) val.w = __ldg(q + 3);
      }
      *reinterpret_cast<float4*>(&xw[r][c]) = val;
    }
    __syncthreads();                                 // gs (staged above: one_chunk) and xw complete
#pragma unroll
    for (int kk = 0; kk < WG_STRIP; ++kk) {
      const int k0 = kstrip + kk * WG_T;
      if (k0 >= n_in) break;
      float acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
      for (int r = 0; r < rows; ++r) {               // rows in order: the same sums as the general path
        const float4 g4 = *reinterpret_cast<const float4*>(&gs[r][4 * ty]);
        const float4 x4 = *reinterpret_cast<const float4*>(&xw[r][kk * WG_T + 4 * tx]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(g[a], xv[b], acc[a][b]);
      }
      const int k = k0 + 4 * tx;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int n = n0 + 4 * ty + a;
        if (n >= n_out) continue;
        float* dst = dW + (int64_t)n * n_in + k;
        if (w_vec && k + 3 < n_in) {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (k + b < n_in) dst[b] = acc[a][b];
        }
      }
    }
    return;
  }
  for (int kk = 0; kk < WG_STRIP; ++kk) {
    const int k0 = kstrip + kk * WG_T;
    if (k0 >= n_in) break;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int r0 = 0; r0 < rows; r0 += WG_ROWS) {
      __syncthreads();                           // previous readers of xs (and of gs when it is re-staged) are done
      if (!one_chunk) stage_rows(gs, gy, ldg, r0, rows, n0, n_out, g_vec, tid, seg_rows, sg);
      stage_rows(xs, x, ldx, r0, rows, k0, n_in, x_vec, tid, seg_rows, sx);
      __syncthreads();
      const int rn = min(WG_ROWS, rows - r0);
#pragma unroll 4
      for (int r = 0; r < rn; ++r) {
        const float4 g4 = *reinterpret_cast<const float4*>(&gs[r][4 * ty]);
        const float4 x4 = *reinterpret_cast<const float4*>(&xs[r][4 * tx]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(g[a], xv[b], acc[a][b]);
      }
    }
    const int k = k0 + 4 * tx;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int n = n0 + 4 * ty + a;
      if (n >= n_out) continue;
      float* dst = dW + (int64_t)n * n_in + k;
      if (w_vec && k + 3 < n_in) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (k + b < n_in) dst[b] = acc[a][b];
      }
    }
  }
}

}  // namespace cgvae

using namespace cgvae;

extern "C" {

int cgvae_wgrad_grouped(const cgvae_wgrad_problem* problems, int n_problems, cgvae_stream_t stream) {
  CGVAE_REQUIRE(n_problems >= 0, "wgrad_grouped: negative problem count");
  if (n_problems == 0) return 0;
  CGVAE_REQUIRE(problems != nullptr, "wgrad_grouped: null problem table");
  for (int i = 0; i < n_problems; ++i) {
    const cgvae_wgrad_problem& q = problems[i];
    CGVAE_REQUIRE(q.gy && (q.dW || q.db), "wgrad_grouped: problem %d has no operand / no output", i);
    CGVAE_REQUIRE(!q.dW || q.x, "wgrad_grouped: problem %d: dW without x", i);
    CGVAE_REQUIRE(q.rows >= 0 && q.n_out >= 1 && q.ldg >= q.n_out, "wgrad_grouped: problem %d: bad gy shape", i);
    CGVAE_REQUIRE(!q.dW || (q.n_in >= 1 && q.ldx >= q.n_in), "wgrad_grouped: problem %d: bad x shape", i);
    CGVAE_REQUIRE(q.seg_rows >= 0 && (q.seg_rows == 0 || q.rows % q.seg_rows == 0),
                  "wgrad_grouped: problem %d: rows must be a multiple of seg_rows", i);
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n_problems; base += WG_MAXP) {
    WgradBatch batch;
    batch.n = std::min(WG_MAXP, n_problems - base);
    int64_t tiles = 0;
    for (int i = 0; i < batch.n; ++i) {
      batch.p[i] = problems[base + i];
      batch.tile_begin[i] = (int)tiles;
      const cgvae_wgrad_problem& q = batch.p[i];
      tiles += ceil_div(q.n_out, WG_T) * (q.dW ? ceil_div(q.n_in, WG_STRIP * WG_T) : 1);
      CGVAE_REQUIRE(tiles < (int64_t)1 << 30, "wgrad_grouped: too many output tiles");
    }
    for (int i = batch.n; i <= WG_MAXP; ++i) batch.tile_begin[i] = (int)tiles;
    for (int i = batch.n; i < WG_MAXP; ++i) batch.p[i] = cgvae_wgrad_problem{};
    launch_kernel(wgrad_grouped_kernel, dim3((unsigned)tiles), dim3(256), 0, st, batch);
    if (int rc = launched("wgrad_grouped")) return rc;
  }
  return 0;
}

}  // extern "C"
