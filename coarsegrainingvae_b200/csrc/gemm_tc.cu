// 3xTF32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) for the large node GEMMs
// (phi-MLPs and UpdateBlock mixes of the protein / large-graph configurations: M = 16 000 ... 128 000 rows).
//
// fp32 parity needs more than one TF32 pass, so every operand element is split on the fly into
//   hi = tf32(x) (round to nearest),  lo = tf32(x - hi)
// and the tile product is accumulated as  A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (error ~2^-21 relative per product).
// The split is SIMT work, so operands are staged global -> registers -> shared by the producer warps (no TMA): they write
// both halves straight into the canonical K-major SWIZZLE_NONE core-matrix layout the UMMA shared-memory descriptor
// expects, which also lets all three operand forms (NT / NN / TN) share one MMA pipeline -- the producers transpose.
//
// Layout of one operand tile [ROWS=128][BK=32] fp32 (units of 16 bytes = 4 tf32):
//   core matrix = 8 rows x 16 B, rows contiguous;  element (r, k) at  (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4
//   LBO = 128 B (next k-chunk), SBO = (BK/4)*128 B = 1024 B (next 8-row group)      [cute: ((8,n),2):((1,SBO),LBO)]
//
// Warp roles (one 128x128 output tile per CTA): warps 0..7 producers (+ warps 0..3 epilogue: TMEM -> registers ->
// global with the fused bias / activation / residual epilogue), warp 8 lane 0 issues tcgen05.mma and commits to mbarriers.
#include "common.cuh"

namespace cgvae {

namespace tc {

constexpr int BM = 128, BN = 64, BK = 32, STAGES = 2;   // 96 KB smem + 64 TMEM columns per CTA: two CTAs per SM
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int NUM_THREADS = (NUM_PRODUCER_WARPS + 1) * 32;
constexpr uint32_t A_TILE_BYTES = BM * BK * 4;               // 16 KB
constexpr uint32_t B_TILE_BYTES = BN * BK * 4;               //  8 KB
constexpr uint32_t STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo
constexpr uint32_t LBO = 128, SBO = (BK / 4) * 128;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024; // + barriers / tmem pointer, +align slack
constexpr uint32_t TMEM_COLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;          // leading byte offset
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;          // stride byte offset
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  return d;                                            // layout type 0 = SWIZZLE_NONE, base offset 0
}

// kind::tf32, fp32 accumulate, M = 128, N = BN; bit 15 / 16 (A / B operand MN-major) stay 0
constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t IDESC) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// one elected lane of a converged warp (see csrc/tc_common.cuh::elect_one)
__device__ __forceinline__ bool elect_one_lane() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Ep {
  const float* bias;
  int act;
  float* z_out;
  const float* z_in;
  int dact;
  const float* add;
};

__device__ __forceinline__ float ep_apply(const Ep& ep, float v, int64_t m, int64_t n, int64_t ldc) {
  if (ep.bias) v += ep.bias[n];
  if (ep.z_out) ep.z_out[m * ldc + n] = v;
  v = act_fwd(ep.act, v);
  if (ep.z_in) v *= act_bwd(ep.dact, ep.z_in[m * ldc + n]);
  if (ep.add) v += ep.add[m * ldc + n];
  return v;
}

// hi = tf32(x) rounded to nearest, lo = tf32(x - hi) rounded to nearest: unbiased halves whatever the tensor core does
// with the 13 low mantissa bits (a truncating split showed a one-sided error growing linearly with K)
__device__ __forceinline__ float to_tf32(float x) {
  // round to nearest (ties away) on the integer ALU: add half a tf32 ulp to the magnitude, clear the 13 low bits
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_store(float x, float* hi, float* lo) {
  const float h = to_tf32(x);
  *hi = h;
  *lo = to_tf32(x - h);
}

// One [128 x BK] operand tile = 1024 float4; the 128 threads of a producer group own 8 each (v = t + 128*it).
//
// K-contiguous source (element (row, k) at ptr[row*ld + k]) -> K-major smem tile:
//     r_lo = v & 7, chunk c = (v >> 3) & 7, row group g = v >> 6;  row = 8g + r_lo, k = 4c..4c+3
//     offset = g*SBO + c*LBO + r_lo*16            (a quarter-warp = 8 rows of one chunk = 8 distinct 16-byte bank slots)
// MN-contiguous source (element (row, k) at ptr[k*ld + row]): the float4 holds rows 4q..4q+3 of ONE k, so it is
// transposed into the same K-major tile with four scalar stores.  Lane bits: k&3 = lane&3, q&1 = (lane>>2)&1,
// rho = (q>>1)&3 = lane>>3, and in store step s a lane writes component j = (s + rho) & 3: the 32 lanes then hit the
// 32 distinct words (r&7, k&3) of a 128-byte bank row -> conflict free.  (The tensor core's MN-major operand mode
// would avoid the transpose, but its tf32 canonical layout could not be validated here; K-major is.)
constexpr int NGROUPS = 2;                                    // producer groups alternate k-steps
constexpr int NPROD = NUM_PRODUCER_WARPS * 32 / NGROUPS;      // threads staging one k-step

template <bool KCONTIG, int ROWS>
__device__ __forceinline__ void fetch_tile(const float* __restrict__ ptr, int64_t ld, int64_t row0, int64_t nrows, int64_t k0,
                                           int64_t K, bool vec_ok, int tid, float4 (&reg)[(ROWS * BK / 4) / NPROD]) {
#pragma unroll
  for (int it = 0; it < (ROWS * BK / 4) / NPROD; ++it) {
    const int v = tid + it * NPROD;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
      const int r = ((v >> 6) << 3) + (v & 7), c = (v >> 3) & 7;
      const int64_t row = row0 + r, k = k0 + 4 * c;
      if (row < nrows) {
        const float* src = ptr + row * ld + k;
        if (vec_ok && k + 3 < K) {
          x = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (k + 0 < K) x.x = src[0];
          if (k + 1 < K) x.y = src[1];
          if (k + 2 < K) x.z = src[2];
          if (k + 3 < K) x.w = src[3];
        }
      }
    } else {
      const int kk = (((v >> 5) & 7) << 2) + (v & 3), q = (((v >> 8) & 3) << 3) + (((v >> 3) & 3) << 1) + ((v >> 2) & 1);
      const int64_t k = k0 + kk, row = row0 + 4 * q;
      if (k < K) {
        const float* src = ptr + k * ld + row;
        if (vec_ok && row + 3 < nrows) {
          x = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (row + 0 < nrows) x.x = src[0];
          if (row + 1 < nrows) x.y = src[1];
          if (row + 2 < nrows) x.z = src[2];
          if (row + 3 < nrows) x.w = src[3];
        }
      }
    }
    reg[it] = x;
  }
}

template <bool KCONTIG, int ROWS>
__device__ __forceinline__ void store_tile(const float4 (&reg)[(ROWS * BK / 4) / NPROD], char* hi_tile, char* lo_tile, int tid) {
#pragma unroll
  for (int it = 0; it < (ROWS * BK / 4) / NPROD; ++it) {
    const int v = tid + it * NPROD;
    const float4 x = reg[it];
    if (KCONTIG) {
      float4 h, l;
      split_store(x.x, &h.x, &l.x);
      split_store(x.y, &h.y, &l.y);
      split_store(x.z, &h.z, &l.z);
      split_store(x.w, &h.w, &l.w);
      const uint32_t off = (uint32_t)(v >> 6) * SBO + (uint32_t)((v >> 3) & 7) * LBO + (uint32_t)(v & 7) * 16;
      *reinterpret_cast<float4*>(hi_tile + off) = h;
      *reinterpret_cast<float4*>(lo_tile + off) = l;
    } else {
      const int kk = (((v >> 5) & 7) << 2) + (v & 3), q = (((v >> 8) & 3) << 3) + (((v >> 3) & 3) << 1) + ((v >> 2) & 1);
      const int rho = (v >> 3) & 3;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int j = (s + rho) & 3;
        const float val = (j == 0) ? x.x : (j == 1) ? x.y : (j == 2) ? x.z : x.w;
        const int r = 4 * q + j;
        const uint32_t off = (uint32_t)(r >> 3) * SBO + (uint32_t)(kk >> 2) * LBO + (uint32_t)(r & 7) * 16 + (uint32_t)(kk & 3) * 4;
        split_store(val, reinterpret_cast<float*>(hi_tile + off), reinterpret_cast<float*>(lo_tile + off));
      }
    }
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(NUM_THREADS, 2) gemm_tc_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                                 int64_t ldb, float* __restrict__ C, int64_t ldc, int64_t M, int64_t N,
                                                                 int64_t K, Ep ep, bool a_vec, bool b_vec) {
  CGVAE_KERNEL_PROLOGUE();
  // NB: index the extern array directly -- rounding the pointer through an integer makes the compiler lose the shared
  // state space and emit generic ST.E instead of STS (seen in the ncu source page: 'stall_lg' on every tile store)
  extern __shared__ __align__(1024) char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_sh = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int num_kt = (int)((K + BK - 1) / BK);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NUM_PRODUCER_WARPS / NGROUPS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NUM_PRODUCER_WARPS) {   // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_sh)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_sh;

  if (warp < NUM_PRODUCER_WARPS) {
    // ---------------- producers: global -> (hi, lo) -> canonical smem layout ----------------
    // Two producer groups (warps 0-3 / 4-7) take alternate k-steps: while one group waits for its global loads the
    // other splits and stores.  (fence.proxy.async waits for ALL of a thread's outstanding loads, so prefetching the
    // next tile inside one thread does not overlap -- seen as long-scoreboard stalls on the fence in the ncu source page.)
    const int group = warp / (NUM_PRODUCER_WARPS / NGROUPS);
    const int t = tid - group * NPROD;
    float4 a_reg[(BM * BK / 4) / NPROD], b_reg[(BN * BK / 4) / NPROD];
    for (int kt = group; kt < num_kt; kt += NGROUPS) {
      const int s = kt % STAGES;
      const uint32_t round = (uint32_t)(kt / STAGES);
      fetch_tile<A_KC, BM>(A, lda, m0, M, (int64_t)kt * BK, K, a_vec, t, a_reg);
      fetch_tile<B_KC, BN>(B, ldb, n0, N, (int64_t)kt * BK, K, b_vec, t, b_reg);
      mbar_wait(&empty_bar[s], (round & 1u) ^ 1u);
      char* stage = smem + (uint32_t)s * STAGE_BYTES;
      store_tile<A_KC, BM>(a_reg, stage, stage + A_TILE_BYTES, t);
      store_tile<B_KC, BN>(b_reg, stage + 2 * A_TILE_BYTES, stage + 2 * A_TILE_BYTES + B_TILE_BYTES, t);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  } else {
    // ---------------- MMA issuer: the whole warp stays in the loop, ONE ELECTED lane drives the tensor core ----------------
    // (issued under a divergent `if (lane == 0)` every tcgen05.mma is wrapped by ptxas in an ELECT / R2UR / BRA.U.ANY
    // serialisation loop: ~205 cycles per instruction whatever its shape, measured with tools/tc_latency.cu; under
    // elect.sync in warp-uniform control flow a 128x64x8 tf32 MMA issues every ~85 cycles)
    for (int kt = 0; kt < num_kt; ++kt) {
      const int s = kt % STAGES;
      const uint32_t round = (uint32_t)(kt / STAGES);
      mbar_wait(&full_bar[s], round & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one_lane()) {
        const uint32_t a_hi = smem_u32(smem + (uint32_t)s * STAGE_BYTES), a_lo = a_hi + A_TILE_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
        constexpr uint32_t idesc = IDESC_BASE;                 // both smem tiles are K-major (the producers transpose)
        // one MMA consumes K = 8: two 16-byte chunks of a K-major tile
        constexpr uint32_t a_step = 2 * LBO, a_lbo = LBO, a_sbo = SBO;
        constexpr uint32_t b_step = 2 * LBO, b_lbo = LBO, b_sbo = SBO;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint32_t ao = (uint32_t)ks * a_step, bo = (uint32_t)ks * b_step;
          const uint32_t acc = (kt > 0 || ks > 0) ? 1u : 0u;
          umma_tf32(tmem_base, make_desc(a_lo + ao, a_lbo, a_sbo), make_desc(b_hi + bo, b_lbo, b_sbo), acc, idesc);   // small terms first
          umma_tf32(tmem_base, make_desc(a_hi + ao, a_lbo, a_sbo), make_desc(b_lo + bo, b_lbo, b_sbo), 1u, idesc);
          umma_tf32(tmem_base, make_desc(a_hi + ao, a_lbo, a_sbo), make_desc(b_hi + bo, b_lbo, b_sbo), 1u, idesc);
        }
        umma_commit(&empty_bar[s]);          // smem slot reusable once these MMAs have read it
        if (kt == num_kt - 1) umma_commit(acc_bar);   // accumulator complete
      }
      __syncwarp();
    }
  }

  if (warp < NUM_PRODUCER_WARPS) {
    // ---------------- epilogue: TMEM -> registers -> smem transpose -> coalesced global stores ----------------
    // warp w reads TMEM lanes 32*(w%4)..+31 (= output rows) and the column half (w/4); the 32x32 block goes through
    // shared memory (the pipeline stages are free by now) so that a warp writes 128 contiguous bytes of one row at a time
    mbar_wait(acc_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    uint32_t r[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * half);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(r[j]);   // row = lane, column = j
    __syncwarp();
    const int64_t n = n0 + 32 * half + lane;
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
      const int64_t m = m0 + 32 * q + rr;
      if (m < M && n < N) C[m * ldc + n] = ep_apply(ep, tr[rr * 33 + lane], m, n, ldc);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == NUM_PRODUCER_WARPS) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

bool tcgen05_enabled() {
  static const bool on = [] {
    const char* e = getenv("CGVAE_TCGEN05");
    return !(e && e[0] == '0');
  }();
  return on;
}

// returns 1 when the problem was launched on the tensor-core path, 0 when the caller should use the SIMT kernel
int launch_gemm_tcgen05(int form, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                        int64_t N, int64_t K, const float* bias, int act, float* z_out, const float* z_in, int dact,
                        const float* add, cudaStream_t st) {
  if (!tcgen05_enabled()) return 0;
  // worth it only when the 128x128 tiles fill the machine and the k-loop amortises the pipeline fill
  // K <= 4096: the accumulator lives in TMEM for the whole k-loop; reductions over node counts (weight gradients,
  // K = 1e4..1e5 rows) stay on the SIMT kernel with its two-level fp32 accumulation
  // accuracy: TMEM accumulation is not round-to-nearest, the error grows linearly with the length of the k-loop
  // (3e-6 at K = 512, 1.5e-5 at K = 2048 on B200) -> the 1e-5 gate allows K <= 1280 here
  static const int min_m = [] { const char* e = getenv("CGVAE_TC_MIN_M"); return e ? atoi(e) : 256; }();
  static const int min_tiles_nt = [] { const char* e = getenv("CGVAE_TC_MIN_TILES_NT"); return e ? atoi(e) : 64; }();
  static const int min_tiles_nn = [] { const char* e = getenv("CGVAE_TC_MIN_TILES_NN"); return e ? atoi(e) : 2 * kNumSM; }();
  if (M < min_m || N < 64 || K < 64 || K > 1280) return 0;
  // measured on B200 against the SIMT tiles (tools/check_tc.py): NT 1.3-1.55x faster once the grid fills the machine;
  // NN (B needs the transposing stage) only pays with >= one full wave of tiles; TN (both operands transposed, and its
  // k-loop runs over node counts) stays on the SIMT kernel
  const int64_t n_tiles = ceil_div(M, tc::BM) * ceil_div(N, tc::BN);
  if (form == CGVAE_GEMM_TN) return 0;
  if (form == CGVAE_GEMM_NT && n_tiles < min_tiles_nt) return 0;
  if (form == CGVAE_GEMM_NN && n_tiles < min_tiles_nn) return 0;
  tc::Ep ep{bias, act, z_out, z_in, dact, add};
  const bool a_kc = (form != CGVAE_GEMM_TN), b_kc = (form == CGVAE_GEMM_NT);
  const bool a_vec = aligned16(A) && (lda % 4 == 0), b_vec = aligned16(B) && (ldb % 4 == 0);
  dim3 grid((unsigned)ceil_div(N, tc::BN), (unsigned)ceil_div(M, tc::BM));
  static bool attr_done[3] = {false, false, false};
  auto set_attr = [&](auto kernel, int idx) {
    if (!attr_done[idx]) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
      attr_done[idx] = true;
    }
  };
  if (a_kc && b_kc) {
    set_attr(tc::gemm_tc_kernel<true, true>, 0);
    launch_kernel(tc::gemm_tc_kernel<true, true>, grid, dim3(tc::NUM_THREADS), tc::SMEM_BYTES, st, A, lda, B, ldb, C, ldc, M, N, K, ep,
                  a_vec, b_vec);
  } else if (a_kc) {
    set_attr(tc::gemm_tc_kernel<true, false>, 1);
    launch_kernel(tc::gemm_tc_kernel<true, false>, grid, dim3(tc::NUM_THREADS), tc::SMEM_BYTES, st, A, lda, B, ldb, C, ldc, M, N, K, ep,
                  a_vec, b_vec);
  } else {
    set_attr(tc::gemm_tc_kernel<false, false>, 2);
    launch_kernel(tc::gemm_tc_kernel<false, false>, grid, dim3(tc::NUM_THREADS), tc::SMEM_BYTES, st, A, lda, B, ldb, C, ldc, M, N, K, ep,
                  a_vec, b_vec);
  }
  return 1;
}

}  // namespace cgvae
