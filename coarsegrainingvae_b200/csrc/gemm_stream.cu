// Weight-streaming contractions for small node sets (decoder graphs of the molecule configs: 12 beads, 36 vector rows).
//
// With M <= 48 rows a Dense layer is M simultaneous matrix-vector products: the cost is reading the weight matrix
// ONCE from HBM (1.4 .. 13 MB at F = 600), the flops are noise.  The tiled kernels of gemm.cu spend their time in
// shared-memory staging and short dependent chains (9 .. 18 us per launch = 0.15 .. 0.7 TB/s).  These two kernels keep
// the weights in registers straight from global memory, many independent 16-byte loads in flight per thread, and keep
// the tiny activation operand in shared memory:
//
//  NT  Y[M,N] = X[M,K] W[N,K]^T   one warp per RW weight rows, lanes along K (coalesced 512-byte row segments), the
//      K-range optionally split over the warps of the CTA; X staged in smem; M x RW accumulators per lane reduced with a
//      transposing butterfly (n/2 shuffles per halving step instead of 5 per value).
//  NN  C[M,N] = G[M,K] W[K,N]     lanes along the output columns, the contraction (weight rows) split over the warps of
//      a CTA and over a thread-block cluster (1,1,S<=8); G^T staged in smem (broadcast LDS.128); warp partials reduced in
//      smem, CTA partials through distributed shared memory in rank order.
//
// Both are deterministic (fixed reduction trees) and carry the epilogue of cgvae_gemm.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cgvae {

struct StreamEpilogue {
  const float* bias;
  int act;
  float* z_out;
  const float* z_in;
  int dact;
  const float* add;
};

__device__ __forceinline__ float stream_epilogue(const StreamEpilogue& ep, float v, int64_t m, int64_t n, int64_t ldc) {
  if (ep.bias) v += ep.bias[n];
  if (ep.z_out) ep.z_out[m * ldc + n] = v;
  v = act_fwd(ep.act, v);
  if (ep.z_in) v *= act_bwd(ep.dact, ep.z_in[m * ldc + n]);
  if (ep.add) v += ep.add[m * ldc + n];
  return v;
}

// ---- transposing warp reduction ---------------------------------------------------------------------------------
// Every lane holds N partial values; afterwards the warp-wide sums are spread over the lanes: lane l holds
// wtr_count(N) sums whose original indices are base .. base + count - 1; lanes with (l & wtr_dupmask(N)) != 0 hold
// duplicates.  While the count is even a step sends one half and keeps the other (N/2 shuffles); an odd remainder is
// finished with plain xor-shuffles.
__host__ __device__ constexpr int wtr_count(int n, int off) { return (off >= 1 && n % 2 == 0) ? wtr_count(n / 2, off / 2) : n; }
__host__ __device__ constexpr int wtr_dupmask(int n, int off) { return (off >= 1 && n % 2 == 0) ? wtr_dupmask(n / 2, off / 2) : (off >= 1 ? 2 * off - 1 : 0); }

template <int N, int OFF>
__device__ __forceinline__ void wtr_step(float* v, int lane, int& base) {
  if constexpr (OFF >= 1 && (N % 2 == 0)) {
    constexpr int H = N / 2;
    const bool up = (lane & OFF) != 0;
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const float send = up ? v[i] : v[i + H];
      const float keep = up ? v[i + H] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    if (up) base += H;
    wtr_step<H, OFF / 2>(v, lane, base);
  } else if constexpr (OFF >= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int o = OFF; o >= 1; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NT: Y[m][n] = sum_k X[m][k] W[n][k]
// ------------------------------------------------------------------------------------------------------------------
constexpr int NT_WARPS = 8;
constexpr int NT_KB = 512;   // floats of K staged per pass (4 chunks of 128 = one float4 per lane)

template <int MR, int RW>
__global__ void __launch_bounds__(NT_WARPS * 32) gemm_nt_stream_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw, float* __restrict__ C, int64_t ldc,
    int M, int N, int K, int KS, StreamEpilogue ep) {
  CGVAE_KERNEL_PROLOGUE();
  extern __shared__ __align__(16) float smem[];
  constexpr int NV = RW * MR;
  float* Xs = smem;                    // [MR][NT_KB]
  float* red = smem + MR * NT_KB;      // [NT_WARPS][NV]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tasks_per_cta = NT_WARPS / KS;
  const int task = blockIdx.x * tasks_per_cta + warp / KS;
  const int slice = warp % KS;
  const int n0 = task * RW;
  const bool task_ok = n0 < N;

  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;

  for (int kb = 0; kb < K; kb += NT_KB) {
    const int klen = min(NT_KB, K - kb);
    const int kpad = (klen + 127) & ~127;
    if (kb > 0) __syncthreads();
    for (int idx = tid; idx < MR * (kpad / 4); idx += NT_WARPS * 32) {
      const int m = idx / (kpad / 4), k4 = 4 * (idx % (kpad / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k4 < klen) v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)m * ldx + kb + k4));
      *reinterpret_cast<float4*>(Xs + m * NT_KB + k4) = v;
    }
    __syncthreads();
    if (task_ok) {
      const int nchunks = kpad / 128;
      float4 wcur[RW], wnext[RW];
      auto fetch = [&](float4 (&w)[RW], int c) {
        const int k = kb + c * 128 + lane * 4;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          w[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (n0 + r < N && k < K) w[r] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(n0 + r) * ldw + k));
        }
      };
      if (slice < nchunks) fetch(wcur, slice);
      for (int c = slice; c < nchunks; c += KS) {
        const bool more = c + KS < nchunks;
        if (more) fetch(wnext, c + KS);
        const float* xrow = Xs + c * 128 + lane * 4;
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(xrow + m * NT_KB);
#pragma unroll
          for (int r = 0; r < RW; ++r) {
            float a = acc[r * MR + m];
            a = fmaf(wcur[r].x, xv.x, a);
            a = fmaf(wcur[r].y, xv.y, a);
            a = fmaf(wcur[r].z, xv.z, a);
            a = fmaf(wcur[r].w, xv.w, a);
            acc[r * MR + m] = a;
          }
        }
        if (more) {
#pragma unroll
          for (int r = 0; r < RW; ++r) wcur[r] = wnext[r];
        }
      }
    }
  }
  // lanes -> sums
  int base = 0;
  wtr_step<NV, 16>(acc, lane, base);
  constexpr int CNT = wtr_count(NV, 16);
  constexpr int DUP = wtr_dupmask(NV, 16);
  const bool writer = (lane & DUP) == 0;
  if (KS > 1) {
    if (writer) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) red[warp * NV + base + i] = acc[i];
    }
    __syncthreads();
    if (slice != 0) return;
    if (writer) {
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        float s = 0.f;
        for (int z = 0; z < KS; ++z) s += red[(warp + z) * NV + base + i];   // slice order: deterministic
        acc[i] = s;
      }
    }
  }
  if (!task_ok || !writer) return;
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const int idx = base + i;
    const int r = idx / MR, m = idx % MR;
    const int n = n0 + r;
    if (m < M && n < N) C[(int64_t)m * ldc + n] = stream_epilogue(ep, acc[i], m, n, ldc);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NN: C[m][n] = sum_k G[m][k] W[k][n]      (k = weight rows: the contraction; n = weight columns: the outputs)
// ------------------------------------------------------------------------------------------------------------------
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int MR, int VEC, int NW, int U>
__global__ void __launch_bounds__(NW * 32) gemm_nn_stream_kernel(
    const float* __restrict__ G, int64_t ldg, const float* __restrict__ W, int64_t ldw, float* __restrict__ C, int64_t ldc,
    int M, int N, int K, int k_per_cta, StreamEpilogue ep) {
  CGVAE_KERNEL_PROLOGUE();
  namespace cg = cooperative_groups;
  using vec_t = typename VecT<VEC>::type;
  constexpr int COLS = 32 * VEC;         // output columns per CTA
  constexpr int GS = MR + 4;             // row stride of the transposed G tile (float4-aligned, spreads banks)
  extern __shared__ __align__(16) float smem[];
  float* part = smem;                    // [MR][COLS] CTA partial (read by the cluster)
  float* Gt = smem + MR * COLS;          // [k_per_cta][GS]     (main loop)
  float* red = smem + MR * COLS;         // [NW][MR][COLS]      (after the main loop; aliases Gt)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = blockIdx.x * COLS + lane * VEC;
  const bool col_ok = col < N;
  const int kbeg = blockIdx.z * k_per_cta;
  const int cnt = max(0, min(K, kbeg + k_per_cta) - kbeg);

  for (int idx = tid; idx < MR * cnt; idx += NW * 32) {
    const int m = idx / cnt, j = idx - m * cnt;
    Gt[j * GS + m] = (m < M) ? __ldg(G + (int64_t)m * ldg + kbeg + j) : 0.f;
  }
  __syncthreads();

  float acc[MR][VEC];
#pragma unroll
  for (int m = 0; m < MR; ++m)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[m][v] = 0.f;

  for (int j0 = warp; j0 < cnt; j0 += NW * U) {
    float w[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NW;
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[u][v] = 0.f;
      if (j < cnt && col_ok) {
        const vec_t t = __ldg(reinterpret_cast<const vec_t*>(W + (int64_t)(kbeg + j) * ldw + col));
        if constexpr (VEC == 1) { w[u][0] = t; }
        if constexpr (VEC == 2) { w[u][0] = t.x; w[u][1] = t.y; }
        if constexpr (VEC == 4) { w[u][0] = t.x; w[u][1] = t.y; w[u][2] = t.z; w[u][3] = t.w; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + u * NW;
      if (j < cnt) {                       // warp-uniform
        const float* g = Gt + j * GS;
#pragma unroll
        for (int m4 = 0; m4 < MR / 4; ++m4) {
          const float4 g4 = *reinterpret_cast<const float4*>(g + 4 * m4);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            acc[4 * m4 + 0][v] = fmaf(g4.x, w[u][v], acc[4 * m4 + 0][v]);
            acc[4 * m4 + 1][v] = fmaf(g4.y, w[u][v], acc[4 * m4 + 1][v]);
            acc[4 * m4 + 2][v] = fmaf(g4.z, w[u][v], acc[4 * m4 + 2][v]);
            acc[4 * m4 + 3][v] = fmaf(g4.w, w[u][v], acc[4 * m4 + 3][v]);
          }
        }
      }
    }
  }
  __syncthreads();                         // Gt is dead: its storage becomes the warp-partial buffer
#pragma unroll
  for (int m = 0; m < MR; ++m)
#pragma unroll
    for (int v = 0; v < VEC; ++v) red[(warp * MR + m) * COLS + lane * VEC + v] = acc[m][v];
  __syncthreads();
  for (int o = tid; o < MR * COLS; o += NW * 32) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) s += red[wv * MR * COLS + o];   // warp order: deterministic
    part[o] = s;
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();
  const int S = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  // rank r finishes rows r, r + S, ...: sum of the S CTA partials in rank order, epilogue, store
  for (int o = tid; o < MR * COLS; o += NW * 32) {
    const int m = o / COLS, c = o - m * COLS;
    if (m % S != rank) continue;
    const int n = blockIdx.x * COLS + c;
    if (m < M && n < N) {
      float s = 0.f;
      for (int z = 0; z < S; ++z) s += *cluster.map_shared_rank(part + o, z);
      C[(int64_t)m * ldc + n] = stream_epilogue(ep, s, m, n, ldc);
    }
  }
  cluster.sync();
}

template <int MR, int RW>
static bool launch_nt(const float* X, int64_t ldx, const float* W, int64_t ldw, float* C, int64_t ldc, int M, int N, int K,
                      const StreamEpilogue& ep, cudaStream_t st) {
  const int64_t tasks = ceil_div(N, RW);
  const int nchunks = (int)ceil_div(std::min<int64_t>(K, NT_KB), 128);
  int KS = 1;
  while (KS < 8 && KS * 2 <= nchunks && ceil_div(tasks * KS, NT_WARPS) < 2 * kNumSM) KS *= 2;
  const size_t smem = sizeof(float) * ((size_t)MR * NT_KB + (size_t)NT_WARPS * RW * MR);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_nt_stream_kernel<MR, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    attr_set = true;
  }
  const int tasks_per_cta = NT_WARPS / KS;
  launch_kernel(gemm_nt_stream_kernel<MR, RW>, dim3((unsigned)ceil_div(tasks, tasks_per_cta)), dim3(NT_WARPS * 32), smem, st, X, ldx, W,
                ldw, C, ldc, M, N, K, KS, ep);
  return true;
}

template <int MR, int VEC, int NW, int U>
static bool launch_nn(const float* G, int64_t ldg, const float* W, int64_t ldw, float* C, int64_t ldc, int M, int N, int K,
                      const StreamEpilogue& ep, cudaStream_t st) {
  constexpr int COLS = 32 * VEC;
  int S = (int)std::min<int64_t>(8, std::max<int64_t>(1, K / (2 * NW)));
  const int k_per_cta = (int)ceil_div(K, S);
  S = (int)ceil_div(K, k_per_cta);
  const size_t smem = sizeof(float) * ((size_t)MR * COLS + std::max<size_t>((size_t)k_per_cta * (MR + 4), (size_t)NW * MR * COLS));
  if (smem > 200 * 1024) return false;
  static size_t attr_bytes = 0;
  if (smem > attr_bytes) {
    if (cudaFuncSetAttribute(gemm_nn_stream_kernel<MR, VEC, NW, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    attr_bytes = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ceil_div(N, COLS), 1, (unsigned)S);
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)S;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, gemm_nn_stream_kernel<MR, VEC, NW, U>, G, ldg, W, ldw, C, ldc, M, N, K, k_per_cta, ep);
  return true;
}

// returns 1 when the problem was launched here, 0 when the caller should use the tiled kernels
int launch_gemm_stream(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                       int64_t N, int64_t K, const float* bias, int act, float* z_out, const float* z_in, int dact,
                       const float* add, cudaStream_t st) {
  static const bool enabled = [] { const char* e = getenv("CGVAE_STREAM_GEMM"); return !(e && e[0] == '0'); }();
  if (!enabled || M > 48 || M < 1 || form == CGVAE_GEMM_TN) return 0;
  if (N * K < 64 * 64 || N >= (1 << 24) || K >= (1 << 24)) return 0;
  const StreamEpilogue ep{bias, act, z_out, z_in, dact, add};
  const int m = (int)M, n = (int)N, k = (int)K;
  if (form == CGVAE_GEMM_NT) {
    // lanes run along K in float4
    if (K % 4 != 0 || K < 128 || lda % 4 != 0 || ldb % 4 != 0 || !aligned16(A) || !aligned16(B)) return 0;
    // few weight rows: two rows per warp double the number of warps that stream
    if (M <= 12) return N <= 1024 ? launch_nt<12, 2>(A, lda, B, ldb, C, ldc, m, n, k, ep, st)
                                  : launch_nt<12, 4>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
    if (M <= 16) return N <= 1024 ? launch_nt<16, 2>(A, lda, B, ldb, C, ldc, m, n, k, ep, st)
                                  : launch_nt<16, 4>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
    if (M <= 36) return launch_nt<36, 2>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
    return launch_nt<48, 2>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
  }
  // NN: C[M,N] = A[M,K] B[K,N]; lanes run along N
  if (K < 64) return 0;
  if (M <= 12) return launch_nn<12, 1, 16, 16>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
  if (M <= 16) return launch_nn<16, 1, 16, 16>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
  if (M <= 36) return launch_nn<36, 1, 8, 8>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
  return launch_nn<48, 1, 8, 8>(A, lda, B, ldb, C, ldc, m, n, k, ep, st);
}

}  // namespace cgvae
