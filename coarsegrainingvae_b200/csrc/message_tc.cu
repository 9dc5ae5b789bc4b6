// Fused message layers on the 5th-generation tensor cores (EquiMessageBlock conv.py:505-563, EquiMessageCross
// conv.py:358-402 and their autograd).
//
// The edge filter  w_k[e][f] = sum_r basis[e][r] * Wf'[k][f][r]   (DistanceEmbed, modules.py:192-197: rbf @ W_filter +
// bias, times the envelope -- the bias rides as the basis column that holds the envelope) is the dense contraction of
// the layer.  Here it is a 3xTF32 tcgen05.mma per batch of 32 edge columns:
//      D_k[128 channels x 32 columns] = Wf'_k[128 x 16] * basis^T[16 x 32]          (accumulators in TMEM)
// with channels on the TMEM lanes, so that the thread that owns channel f reads w_k[.][f] of 32 edges with tcgen05.ld
// and keeps its gathers coalesced (a warp = 32 consecutive channels of one sender row) and its receiver sums in registers.
//
// Column tiles.  Edges are re-ordered into COLUMN TILES: a chunk of RC consecutive receivers x the sorted union of their
// senders.  Column (g, rr) = edge (receiver chunk*RC + rr  <-  sender gcol[g]); pairs that are not edges keep an all-zero
// basis row, so their filter value is exactly 0.  The thread gathers the six sender values (phi_0..2, v_x..z) ONCE per
// group and re-uses them from registers for the RC receivers of the chunk -- on the molecule graphs (chignolin: degree
// 159 of 175) that is ~7x fewer gathers than one gather set per edge, which is what the SIMT kernel of message.cu is
// bound by (26 % address arithmetic, 62 % issue utilisation, profiles/r1b_message_kernels_hotspots.txt).
// A batch record (32 columns: both tf32 halves of the basis tile in the canonical K-major core-matrix layout, unit
// vectors, sender ids) is ONE 4.6 KB cp.async.bulk; records are produced once per graph (cgvae_msg_tiles_build) and
// shared by every layer that runs on that graph.
//
// Warp roles (192 threads): warps 0-3 = 128 channel threads (gather, tcgen05.ld, mix, reduce), warp 4 lane 0 = bulk-copy
// producer (4-stage mbarrier ring), warp 5 lane 0 = MMA issuer (+ TMEM allocation: 2 accumulator buffers).
// Two CTAs per SM (256 TMEM columns, <= 84 KB shared memory each).
//
// Backward: the same structure on the transposed tiles (chunk of RC SENDERS x union of their receivers); the filter
// gradient dWf'_k[f][r] = sum_cols G_k[f][col] * basis[col][r] is a second tcgen05.mma whose A operand G_k (produced by
// the channel threads) is written to TMEM with tcgen05.st and whose B operand is the SAME shared-memory basis tile read
// MN-major.
#include "tc_common.cuh"

namespace cgvae {
namespace mtc {

using namespace tcx;

constexpr int NB = 32;                 // columns per batch = MMA N
constexpr int KT = 16;                 // padded filter K (R + 1 <= 16)
constexpr uint32_t TILE_SBO = (KT / 4) * kLBO;   // 512
constexpr int REC_BHI = 0;             // [32 x 16] tf32 hi, canonical layout (2048 B)
constexpr int REC_BLO = 2048;          // [32 x 16] tf32 lo
constexpr int REC_UNIT = 4096;         // unit vectors, component-major: ux[32] uy[32] uz[32]
constexpr int REC_GCOL = 4480;         // partner (sender) index of each group of the batch: NB / RC <= 8 ints
constexpr int REC_BYTES = 4608;        // 36 x 128
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 192;
constexpr int A_TILE_BYTES = 128 * KT * 4;       // 8 KB per (split, half)
constexpr int TMEM_COLS = 256;

// ------------------------------------------------------------------------------------------------------------------
// tile builder (integer work, exact, deterministic)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;

// shared: bitmap[W] | seg_pref[kTileThreads + 1]
__device__ __forceinline__ int build_bitmap(uint32_t* bm, int W, const int32_t* __restrict__ col, int e0, int e1) {
  for (int w = threadIdx.x; w < W; w += kTileThreads) bm[w] = 0u;
  __syncthreads();
  for (int e = e0 + threadIdx.x; e < e1; e += kTileThreads) {
    const int j = col[e];
    atomicOr(&bm[j >> 5], 1u << (j & 31));
  }
  __syncthreads();
  const int SEG = (W + kTileThreads - 1) / kTileThreads;
  int cnt = 0;
  const int w0 = threadIdx.x * SEG, w1 = min(W, w0 + SEG);
  for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
  return cnt;
}

// exclusive scan of one int per thread over the block into pref[0 .. kTileThreads] (pref[kTileThreads] = total)
__device__ __forceinline__ void block_scan(int v, int* pref) {
  __shared__ int warp_tot[kTileThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  pref[threadIdx.x] = base + x - v;
  if (threadIdx.x == kTileThreads - 1) pref[kTileThreads] = base + x;
  __syncthreads();
}

__global__ void __launch_bounds__(kTileThreads) tile_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                 int64_t n_rows, int64_t n_partners, int RC,
                                                                 int32_t* __restrict__ nbatch, int32_t* __restrict__ ngroups) {
  CGVAE_KERNEL_PROLOGUE();
  extern __shared__ uint32_t tile_sm[];
  const int W = (int)((n_partners + 31) / 32);
  uint32_t* bm = tile_sm;
  int* pref = reinterpret_cast<int*>(tile_sm + W);
  const int64_t r0 = (int64_t)blockIdx.x * RC, r1 = min(n_rows, r0 + RC);
  const int cnt = build_bitmap(bm, W, col, rowptr[r0], rowptr[r1]);
  block_scan(cnt, pref);
  if (threadIdx.x == 0) {
    const int ng = pref[kTileThreads], GB = NB / RC;
    ngroups[blockIdx.x] = ng;
    nbatch[blockIdx.x] = (ng + GB - 1) / GB;
  }
}

__global__ void __launch_bounds__(kTileThreads) tile_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                const int32_t* __restrict__ slot_map, int64_t n_rows,
                                                                int64_t n_partners, int RC, const int32_t* __restrict__ bptr,
                                                                const float* __restrict__ basis, const float* __restrict__ unit,
                                                                int RB, char* __restrict__ rec, int64_t n_batches_cap) {
  CGVAE_KERNEL_PROLOGUE();
  extern __shared__ uint32_t tile_sm[];
  const int W = (int)((n_partners + 31) / 32);
  uint32_t* bm = tile_sm;
  int* pref = reinterpret_cast<int*>(tile_sm + W);
  const int64_t r0 = (int64_t)blockIdx.x * RC, r1 = min(n_rows, r0 + RC);
  const int cnt = build_bitmap(bm, W, col, rowptr[r0], rowptr[r1]);
  block_scan(cnt, pref);
  const int SEG = (W + kTileThreads - 1) / kTileThreads;
  const int GB = NB / RC;
  const int64_t b0 = bptr[blockIdx.x];
  // sender id of every group (sorted ascending: bit order)
  {
    int g = pref[threadIdx.x];
    const int w0 = threadIdx.x * SEG, w1 = min(W, w0 + SEG);
    for (int w = w0; w < w1; ++w) {
      uint32_t bits = bm[w];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int64_t batch = b0 + g / GB;
        if (batch < n_batches_cap)
          reinterpret_cast<int32_t*>(rec + batch * REC_BYTES + REC_GCOL)[g % GB] = w * 32 + b;
        ++g;
      }
    }
  }
  // every edge of the chunk -> its column
  const int RBQ = RB / 4;
  for (int64_t i = r0; i < r1; ++i) {
    const int rr = (int)(i - r0);
    for (int e = rowptr[i] + threadIdx.x; e < rowptr[i + 1]; e += kTileThreads) {
      const int j = col[e];
      const int w = j >> 5, seg = w / SEG;
      int g = pref[seg];
      for (int ww = seg * SEG; ww < w; ++ww) g += __popc(bm[ww]);
      g += __popc(bm[w] & ((1u << (j & 31)) - 1u));
      const int64_t batch = b0 + g / GB;
      if (batch >= n_batches_cap) continue;
      const int c = (g % GB) * RC + rr;                  // column inside the batch
      char* r = rec + batch * REC_BYTES;
      const int64_t src = slot_map ? slot_map[e] : e;
      const float4* brow = reinterpret_cast<const float4*>(basis) + src * RBQ;
      for (int q = 0; q < RBQ; ++q) {
        const float4 x = __ldg(brow + q);
        float4 h, l;
        split_tf32(x.x, h.x, l.x);
        split_tf32(x.y, h.y, l.y);
        split_tf32(x.z, h.z, l.z);
        split_tf32(x.w, h.w, l.w);
        const uint32_t off = tile_off(c, 4 * q, TILE_SBO);
        *reinterpret_cast<float4*>(r + REC_BHI + off) = h;
        *reinterpret_cast<float4*>(r + REC_BLO + off) = l;
      }
      const float4 u = __ldg(reinterpret_cast<const float4*>(unit) + src);
      float* un = reinterpret_cast<float*>(r + REC_UNIT);
      un[c] = u.x;
      un[NB + c] = u.y;
      un[2 * NB + c] = u.z;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// shared pieces of the forward / backward kernels
// ------------------------------------------------------------------------------------------------------------------
struct Smem {
  char* a_tiles;        // [KS][hi | lo][8 KB]
  char* stages;         // [NSTAGE][REC_BYTES]
  uint64_t* stage_full;
  uint64_t* stage_empty;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  uint32_t* tmem_ptr;
};

template <int KS>
__device__ __forceinline__ Smem carve(char* smem) {
  Smem s;
  s.a_tiles = smem;
  s.stages = smem + KS * 2 * A_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.stages + NSTAGE * REC_BYTES);
  s.stage_full = bars;
  s.stage_empty = bars + NSTAGE;
  s.tmem_full = bars + 2 * NSTAGE;
  s.tmem_empty = bars + 2 * NSTAGE + 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  return s;
}
template <int KS>
constexpr size_t smem_bytes() { return (size_t)KS * 2 * A_TILE_BYTES + NSTAGE * REC_BYTES + (2 * NSTAGE + 4) * 8 + 64; }

// Wf'[k][f0 + row][0..15] = [Wf[(k*F + f)*R + r] (r < R) | bf[k*F + f] | 0 ...], both tf32 halves, canonical K-major tiles
template <int KS>
__device__ __forceinline__ void stage_filter_tiles(char* a_tiles, const float* __restrict__ Wf, const float* __restrict__ bf, int F,
                                                   int R, int f0) {
  for (int idx = threadIdx.x; idx < KS * 128 * (KT / 4); idx += NTHREADS) {
    const int c4 = idx & 3, row = (idx >> 2) & 127, k = idx >> 9;
    const int ch = f0 + row;
    float x[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * c4 + j;
      float val = 0.f;
      if (ch < F) {
        if (r < R) val = __ldg(Wf + ((int64_t)k * F + ch) * R + r);
        else if (r == R) val = __ldg(bf + (int64_t)k * F + ch);
      }
      x[j] = val;
    }
    float4 h, l;
    split_tf32(x[0], h.x, l.x);
    split_tf32(x[1], h.y, l.y);
    split_tf32(x[2], h.z, l.z);
    split_tf32(x[3], h.w, l.w);
    const uint32_t off = tile_off(row, 4 * c4, TILE_SBO);
    *reinterpret_cast<float4*>(a_tiles + (2 * k) * A_TILE_BYTES + off) = h;
    *reinterpret_cast<float4*>(a_tiles + (2 * k + 1) * A_TILE_BYTES + off) = l;
  }
}

// bulk-copy producer: one thread, NSTAGE-deep ring of batch records
__device__ __forceinline__ void producer_loop(const Smem& sm, const char* __restrict__ rec, int64_t b0, int nb) {
  const char* src = rec + b0 * REC_BYTES;
  for (int b = 0; b < nb; ++b) {
    const int s = b % NSTAGE;
    mbar_wait(&sm.stage_empty[s], (((uint32_t)(b / NSTAGE)) & 1u) ^ 1u);
    mbar_expect_tx(&sm.stage_full[s], REC_BYTES);
    bulk_g2s(sm.stages + s * REC_BYTES, src + (int64_t)b * REC_BYTES, REC_BYTES, &sm.stage_full[s]);
  }
}

// filter MMAs of one batch: D_k = A_k * B^T for every split, 3xTF32 (small terms first), K = 16 in two steps
template <int KS>
__device__ __forceinline__ void issue_filter_mmas(const Smem& sm, int s, uint32_t tmem_d) {
  constexpr uint32_t idesc = idesc_tf32(128, NB);
  const uint32_t b_hi = smem_u32(sm.stages + s * REC_BYTES + REC_BHI), b_lo = b_hi + (REC_BLO - REC_BHI);
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const uint32_t a_hi = smem_u32(sm.a_tiles + (2 * k) * A_TILE_BYTES), a_lo = a_hi + A_TILE_BYTES;
    const uint32_t d = tmem_d + (uint32_t)(k * NB);
#pragma unroll
    for (int ks = 0; ks < KT / 8; ++ks) {
      const uint32_t o = (uint32_t)ks * 2 * kLBO;
      umma_tf32_ss(d, make_desc(a_lo + o, kLBO, TILE_SBO), make_desc(b_hi + o, kLBO, TILE_SBO), ks > 0 ? 1u : 0u, idesc);
      umma_tf32_ss(d, make_desc(a_hi + o, kLBO, TILE_SBO), make_desc(b_lo + o, kLBO, TILE_SBO), 1u, idesc);
      umma_tf32_ss(d, make_desc(a_hi + o, kLBO, TILE_SBO), make_desc(b_hi + o, kLBO, TILE_SBO), 1u, idesc);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
template <int KS, int RC>
__global__ void __launch_bounds__(NTHREADS, 2) message_tc_fwd_kernel(
    const float* __restrict__ phi, const float* __restrict__ v_send, const float* __restrict__ v_recv,
    const int32_t* __restrict__ bptr, const int32_t* __restrict__ ngroups, const char* __restrict__ rec,
    const float* __restrict__ Wf, const float* __restrict__ bf, int64_t n_recv, int F, int R,
    const float* __restrict__ res_s, const float* __restrict__ res_v, int v_is_zero, float* __restrict__ out_s,
    float* __restrict__ out_v, float* __restrict__ q_out) {
  CGVAE_KERNEL_PROLOGUE();
  constexpr int GB = NB / RC;
  extern __shared__ __align__(1024) char smem_raw[];
  const Smem sm = carve<KS>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x, f0 = blockIdx.y * 128;
  const int64_t b0 = bptr[chunk];
  const int nb = bptr[chunk + 1] - bptr[chunk];
  const int ng = ngroups[chunk];

  if (nb > 0) {
    if (tid == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(&sm.stage_full[s], 1);
        mbar_init(&sm.stage_empty[s], 4);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&sm.tmem_full[b], 1);
        mbar_init(&sm.tmem_empty[b], 4);
      }
      mbar_fence_init();
    }
    if (warp == 5) tmem_alloc<TMEM_COLS>(sm.tmem_ptr);
    stage_filter_tiles<KS>(sm.a_tiles, Wf, bf, F, R, f0);
    fence_async_smem();                      // generic-proxy tile writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = nb > 0 ? *sm.tmem_ptr : 0u;

  if (warp == 4) {
    if (lane == 0 && nb > 0) producer_loop(sm, rec, b0, nb);
  } else if (warp == 5) {
    if (lane == 0) {
      for (int b = 0; b < nb; ++b) {
        const int s = b % NSTAGE, buf = b & 1;
        mbar_wait(&sm.stage_full[s], ((uint32_t)(b / NSTAGE)) & 1u);
        mbar_wait(&sm.tmem_empty[buf], (((uint32_t)(b >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        issue_filter_mmas<KS>(sm, s, tmem_base + (uint32_t)(buf * KS * NB));
        umma_commit(&sm.tmem_full[buf]);
      }
    }
  } else {
    // ---------------- channel threads ----------------
    const int f = f0 + tid;
    const bool active = f < F;
    const int fc = active ? f : F - 1;       // inactive lanes compute on a valid channel, never store
    float acc_s[RC], acc_v[3][RC], acc_q[KS == 4 ? 3 : 1][RC];
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
      acc_s[rr] = 0.f;
      acc_v[0][rr] = acc_v[1][rr] = acc_v[2][rr] = 0.f;
#pragma unroll
      for (int c = 0; c < (KS == 4 ? 3 : 1); ++c) acc_q[c][rr] = 0.f;
    }
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int b = 0; b < nb; ++b) {
      const int s = b % NSTAGE, buf = b & 1;
      mbar_wait(&sm.stage_full[s], ((uint32_t)(b / NSTAGE)) & 1u);
      const char* st = sm.stages + s * REC_BYTES;
      const int32_t* gcol = reinterpret_cast<const int32_t*>(st + REC_GCOL);
      const float* un = reinterpret_cast<const float*>(st + REC_UNIT);
      const int n_live = min(GB, ng - b * GB);
      // gathers of every live group of the batch: in flight while the tensor core finishes the batch
      float ph[GB][KS], vv[GB][3];
#pragma unroll
      for (int g = 0; g < GB; ++g) {
        const int j = (g < n_live) ? gcol[g] : 0;
        const float* pj = phi + (int64_t)j * KS * F + fc;
#pragma unroll
        for (int k = 0; k < KS; ++k) ph[g][k] = (g < n_live) ? __ldg(pj + (int64_t)k * F) : 0.f;
        if (!v_is_zero && g < n_live) {
          const float* vj = v_send + (int64_t)j * 3 * F + fc;
          vv[g][0] = __ldg(vj); vv[g][1] = __ldg(vj + F); vv[g][2] = __ldg(vj + 2 * (int64_t)F);
        } else {
          vv[g][0] = vv[g][1] = vv[g][2] = 0.f;
        }
      }
      mbar_wait(&sm.tmem_full[buf], ((uint32_t)(b >> 1)) & 1u);
      tc_fence_after();
      const uint32_t t_buf = t_lane + (uint32_t)(buf * KS * NB);
#pragma unroll
      for (int g = 0; g < GB; ++g) {
        if (g < n_live) {
          uint32_t w[KS][RC];
#pragma unroll
          for (int k = 0; k < KS; ++k) tmem_ld<RC>(t_buf + (uint32_t)(k * NB + g * RC), w[k]);
          tmem_wait_ld();
          const float p0 = ph[g][0] * vv[g][0], p1 = ph[g][0] * vv[g][1], p2 = ph[g][0] * vv[g][2];
          float x0 = 0.f, x1 = 0.f, x2 = 0.f;
          if constexpr (KS == 4) {
            x0 = ph[g][3] * vv[g][0]; x1 = ph[g][3] * vv[g][1]; x2 = ph[g][3] * vv[g][2];
          }
#pragma unroll
          for (int rr = 0; rr < RC; ++rr) {
            const int c = g * RC + rr;
            const float w0 = __uint_as_float(w[0][rr]), w1 = __uint_as_float(w[1][rr]), w2 = __uint_as_float(w[2][rr]);
            acc_s[rr] = fmaf(ph[g][1], w1, acc_s[rr]);                     // dS_i += m1            (conv.py:526,558)
            const float t = ph[g][2] * w2;                                  // m2
            acc_v[0][rr] = fmaf(t, un[c], acc_v[0][rr]);                    // dV_i += m2 * u_ij     (conv.py:524)
            acc_v[1][rr] = fmaf(t, un[NB + c], acc_v[1][rr]);
            acc_v[2][rr] = fmaf(t, un[2 * NB + c], acc_v[2][rr]);
            acc_v[0][rr] = fmaf(w0, p0, acc_v[0][rr]);                      //        + m0 * v_j     (conv.py:525)
            acc_v[1][rr] = fmaf(w0, p1, acc_v[1][rr]);
            acc_v[2][rr] = fmaf(w0, p2, acc_v[2][rr]);
            if constexpr (KS == 4) {
              const float w3 = __uint_as_float(w[3][rr]);                   // q_i = sum m3 * v_j    (conv.py:379)
              acc_q[0][rr] = fmaf(w3, x0, acc_q[0][rr]);
              acc_q[1][rr] = fmaf(w3, x1, acc_q[1][rr]);
              acc_q[2][rr] = fmaf(w3, x2, acc_q[2][rr]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sm.tmem_empty[buf]);
        mbar_arrive(&sm.stage_empty[s]);
      }
    }
    // ---------------- receiver epilogue: single store per output element, residual fused ----------------
    if (active) {
#pragma unroll
      for (int rr = 0; rr < RC; ++rr) {
        const int64_t i = (int64_t)chunk * RC + rr;
        if (i < n_recv) {
          const int64_t so = i * F + f, vo = i * 3 * F + f;
          float a0 = acc_v[0][rr], a1 = acc_v[1][rr], a2 = acc_v[2][rr];
          if constexpr (KS == 4) {
            // sum_e m3 (v_i x v_j) = v_i x q_i
            float vi0 = 0.f, vi1 = 0.f, vi2 = 0.f;
            if (!v_is_zero) {
              vi0 = v_recv[vo]; vi1 = v_recv[vo + F]; vi2 = v_recv[vo + 2 * (int64_t)F];
            }
            a0 += vi1 * acc_q[2][rr] - vi2 * acc_q[1][rr];
            a1 += vi2 * acc_q[0][rr] - vi0 * acc_q[2][rr];
            a2 += vi0 * acc_q[1][rr] - vi1 * acc_q[0][rr];
            if (q_out) {
              q_out[vo] = acc_q[0][rr]; q_out[vo + F] = acc_q[1][rr]; q_out[vo + 2 * (int64_t)F] = acc_q[2][rr];
            }
          }
          out_s[so] = (res_s ? res_s[so] : 0.f) + acc_s[rr];
          out_v[vo] = (res_v ? res_v[vo] : 0.f) + a0;
          out_v[vo + F] = (res_v ? res_v[vo + F] : 0.f) + a1;
          out_v[vo + 2 * (int64_t)F] = (res_v ? res_v[vo + 2 * (int64_t)F] : 0.f) + a2;
        }
      }
    }
  }
  if (nb > 0) {
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
      tc_fence_after();
      tmem_dealloc<TMEM_COLS>(tmem_base);
    }
  }
}

}  // namespace mtc
}  // namespace cgvae

using namespace cgvae;

extern "C" {

int64_t cgvae_msg_tiles_batches_cap(int64_t n_rows, int64_t n_partners, int64_t n_edge_slots, int RC) {
  if (RC != 4 && RC != 8 && RC != 16) return -1;
  const int64_t n_chunks = ceil_div(n_rows, RC), GB = mtc::NB / RC;
  int64_t groups = std::min(n_edge_slots, n_chunks * n_partners);   // every group holds at least one edge
  return ceil_div(groups, GB) + n_chunks + 1;
}
size_t cgvae_msg_tiles_rec_bytes(void) { return mtc::REC_BYTES; }

int cgvae_msg_tiles_build(const int32_t* rowptr, const int32_t* col, const int32_t* slot_map, int64_t n_rows, int64_t n_partners,
                          int64_t n_edge_slots, const float* basis, const float* unit, int RB, int RC, int32_t* nbatch,
                          int32_t* bptr, int32_t* ngroups, void* rec, int64_t n_batches_cap, cgvae_stream_t stream) {
  CGVAE_REQUIRE(RC == 4 || RC == 8 || RC == 16, "msg_tiles_build: RC must be 4, 8 or 16 (got %d)", RC);
  CGVAE_REQUIRE(RB == 8 || RB == 12 || RB == 16, "msg_tiles_build: RB must be 8, 12 or 16 (got %d)", RB);
  CGVAE_REQUIRE(n_partners <= 262144, "msg_tiles_build: at most 262144 partner nodes (got %lld)", (long long)n_partners);
  CGVAE_REQUIRE(rowptr && col && basis && unit && nbatch && bptr && ngroups && rec, "msg_tiles_build: null pointer");
  CGVAE_REQUIRE(n_batches_cap >= cgvae_msg_tiles_batches_cap(n_rows, n_partners, n_edge_slots, RC),
                "msg_tiles_build: record capacity too small");
  CGVAE_REQUIRE(aligned16(basis) && aligned16(unit) && aligned16(rec), "msg_tiles_build: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_chunks = ceil_div(n_rows, RC);
  if (n_chunks == 0) return 0;
  const size_t W = (size_t)((n_partners + 31) / 32);
  const size_t sm_bytes = sizeof(uint32_t) * W + sizeof(int) * (mtc::kTileThreads + 1) + 16;
  static std::atomic<size_t> attr_bytes{0};
  if (sm_bytes > 48 * 1024 && attr_bytes.load() < sm_bytes) {
    CGVAE_CUDA(cudaFuncSetAttribute(mtc::tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes));
    CGVAE_CUDA(cudaFuncSetAttribute(mtc::tile_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bytes));
    attr_bytes.store(sm_bytes);
  }
  launch_kernel(mtc::tile_count_kernel, dim3((unsigned)n_chunks), dim3(mtc::kTileThreads), sm_bytes, st, rowptr, col, n_rows, n_partners, RC,
                nbatch, ngroups);
  if (int rc = launched("msg_tile_count")) return rc;
  if (int rc = cgvae_scan_i32(nbatch, n_chunks, bptr, stream)) return rc;
  CGVAE_ZERO(rec, (size_t)n_batches_cap * mtc::REC_BYTES, st);
  launch_kernel(mtc::tile_fill_kernel, dim3((unsigned)n_chunks), dim3(mtc::kTileThreads), sm_bytes, st, rowptr, col, slot_map, n_rows,
                n_partners, RC, (const int32_t*)bptr, basis, unit, RB, reinterpret_cast<char*>(rec), n_batches_cap);
  return launched("msg_tile_fill");
}

int cgvae_message_tc_fwd(int n_split, const float* phi, const float* v_send, const float* v_recv, const int32_t* bptr,
                         const int32_t* ngroups, const void* rec, int RC, const float* Wf, const float* bf, int64_t n_recv, int F,
                         int R, const float* res_s, const float* res_v, int v_is_zero, float* out_s, float* out_v, float* q,
                         cgvae_stream_t stream) {
  CGVAE_REQUIRE(n_split == 3 || n_split == 4, "message_tc_fwd: n_split must be 3 or 4 (got %d)", n_split);
  CGVAE_REQUIRE(RC == 4 || RC == 8 || RC == 16, "message_tc_fwd: RC must be 4, 8 or 16 (got %d)", RC);
  CGVAE_REQUIRE(R >= 1 && R + 1 <= mtc::KT && F >= 1, "message_tc_fwd: need 1 <= R <= 15");
  if (n_recv == 0) return 0;
  CGVAE_REQUIRE(phi && bptr && ngroups && rec && Wf && bf && out_s && out_v, "message_tc_fwd: null pointer");
  CGVAE_REQUIRE(v_is_zero || v_send, "message_tc_fwd: v_send missing");
  CGVAE_REQUIRE(n_split != 4 || v_is_zero || v_recv, "message_tc_fwd: v_recv missing for the cross block");
  CGVAE_REQUIRE(aligned16(rec), "message_tc_fwd: records must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)ceil_div(n_recv, RC), (unsigned)ceil_div(F, 128));
  const char* recp = reinterpret_cast<const char*>(rec);
#define LAUNCH_TC_FWD(KS, RCV)                                                                                                   \
  do {                                                                                                                           \
    static bool attr_done = false;                                                                                               \
    constexpr size_t smb = mtc::smem_bytes<KS>();                                                                                \
    if (!attr_done) {                                                                                                            \
      CGVAE_CUDA(cudaFuncSetAttribute(mtc::message_tc_fwd_kernel<KS, RCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb)); \
      attr_done = true;                                                                                                          \
    }                                                                                                                            \
    launch_kernel(mtc::message_tc_fwd_kernel<KS, RCV>, grid, dim3(mtc::NTHREADS), smb, st, phi, v_send, v_recv, bptr, ngroups, recp, Wf, \
                  bf, n_recv, F, R, res_s, res_v, v_is_zero, out_s, out_v, q);                                                   \
  } while (0)
  if (n_split == 3) {
    if (RC == 4) LAUNCH_TC_FWD(3, 4); else if (RC == 8) LAUNCH_TC_FWD(3, 8); else LAUNCH_TC_FWD(3, 16);
  } else {
    if (RC == 4) LAUNCH_TC_FWD(4, 4); else if (RC == 8) LAUNCH_TC_FWD(4, 8); else LAUNCH_TC_FWD(4, 16);
  }
#undef LAUNCH_TC_FWD
  return launched("message_tc_fwd");
}

}  // extern "C"
