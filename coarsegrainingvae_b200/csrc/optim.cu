// Fused gradient clipping + Adam on flat buffers: torch.nn.utils.clip_grad_norm_(params, max_norm) followed by
// torch.optim.Adam.step() (scripts/utils.py:151-157 of the reference) in two launches over one contiguous parameter /
// gradient / moment layout -- one read of the gradient for the norm, one streaming pass for the update.
#include "common.cuh"

namespace cgvae {

constexpr int kNormBlocks = 1024;

// partial[b] = sum of squares of a fixed, contiguous slice; block 0 also clears the "blocks done" ticket of the update pass
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial,
                                                            unsigned int* __restrict__ ticket) {
  CGVAE_KERNEL_PROLOGUE();
  __shared__ float red[8];
  const int64_t per = (n + kNormBlocks - 1) / kNormBlocks;
  const int64_t beg = (int64_t)blockIdx.x * per, end = min(n, beg + per);
  float s = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) {
    const float x = g[i];
    s = fmaf(x, x, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
    if (blockIdx.x == 0) *ticket = 0u;
  }
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ partial,
                                                        float max_norm, float grad_scale, float lr, float beta1, float beta2,
                                                        float eps, float* __restrict__ step, float* __restrict__ norm_out,
                                                        const float* __restrict__ loss, float loss_scale, float loss_limit,
                                                        float* __restrict__ skipped, unsigned int* __restrict__ ticket) {
  CGVAE_KERNEL_PROLOGUE();
  // every block re-reduces the 1024 partial sums in the same fixed order: identical clip coefficient everywhere
  __shared__ float red[8];
  __shared__ float coef_sh;
  __shared__ int skip_sh;
  float s = 0.f;
  for (int i = threadIdx.x; i < kNormBlocks; i += 256) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    const float norm = grad_scale * sqrtf(t);                // g holds the SUM over ranks: the mean's norm is scale * |g|
    // clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1; the 1/world of the data-parallel mean rides along
    coef_sh = grad_scale * fminf(max_norm / (norm + 1e-6f), 1.0f);
    if (blockIdx.x == 0 && norm_out) norm_out[0] = norm;
    // Skip guard of the reference loop (scripts/utils.py:145-148: `loss >= gamma*200 or isnan(loss)` -> `continue`, no
    // optimiser step), decided on the device from the (all-reduced) loss so that the step stays graph-capturable and
    // every data-parallel rank takes the same decision.  A non-finite gradient norm is skipped as well: fminf() would
    // turn a NaN norm into coef = 1 and write the NaN gradients into the moments and the parameters for good.
    bool skip = !isfinite(norm);
    if (loss != nullptr) {
      const float l = loss[0] * loss_scale;
      skip = skip || isnan(l) || l >= loss_limit;
    }
    skip_sh = skip ? 1 : 0;
  }
  __syncthreads();
  const float coef = coef_sh;
  const bool skip = skip_sh != 0;
  // the step counter is read by every block and advanced by the LAST block to finish (ticket): all other blocks have
  // read it by then.  A skipped step leaves p / m / v and the counter untouched.
  const float t = step[0] + 1.0f;
  if (skip) {
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(ticket, 1u) == gridDim.x - 1 && skipped) skipped[0] += 1.0f;
    }
    return;
  }
  const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 gp = *reinterpret_cast<const float4*>(g + i);
      float4 mp = *reinterpret_cast<const float4*>(m + i);
      float4 vp = *reinterpret_cast<const float4*>(v + i);
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      float* ga = &gp.x; float* ma = &mp.x; float* va = &vp.x; float* pa = &pp.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gg = ga[k] * coef;
        ma[k] = beta1 * ma[k] + (1.0f - beta1) * gg;
        va[k] = beta2 * va[k] + (1.0f - beta2) * gg * gg;
        pa[k] -= step_size * ma[k] / (sqrtf(va[k]) * inv_sqrt_bc2 + eps);
      }
      *reinterpret_cast<float4*>(m + i) = mp;
      *reinterpret_cast<float4*>(v + i) = vp;
      *reinterpret_cast<float4*>(p + i) = pp;
    } else {
      for (int64_t j = i; j < n; ++j) {
        const float gg = g[j] * coef;
        const float mm = beta1 * m[j] + (1.0f - beta1) * gg;
        const float vv = beta2 * v[j] + (1.0f - beta2) * gg * gg;
        m[j] = mm; v[j] = vv;
        p[j] -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) step[0] = t;
  }
}

}  // namespace cgvae

using namespace cgvae;

extern "C" {

size_t cgvae_adam_ws_bytes(void) { return sizeof(float) * (kNormBlocks + 4); }

int cgvae_adam_clip_step(float* p, const float* g, float* m, float* v, int64_t n, float max_norm, float grad_scale, float lr,
                         float beta1, float beta2, float eps, float* step, float* norm_out, const float* loss,
                         float loss_scale, float loss_limit, float* skipped, void* ws, size_t ws_bytes,
                         cgvae_stream_t stream) {
  if (n == 0) return 0;
  CGVAE_REQUIRE(p && g && m && v && step && ws && ws_bytes >= cgvae_adam_ws_bytes(), "adam_clip_step: bad arguments");
  CGVAE_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adam_clip_step: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + kNormBlocks + 1);      // [1024] = loss slot of the sharded form
  launch_kernel(sumsq_partial_kernel, dim3(kNormBlocks), dim3(256), 0, st, g, n, partial, ticket);
  if (int rc = launched("sumsq_partial")) return rc;
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 8);
  launch_kernel(adam_clip_kernel, dim3(blocks), dim3(256), 0, st, p, g, m, v, n, (const float*)partial, max_norm, grad_scale, lr, beta1, beta2, eps,
                step, norm_out, loss, loss_scale, loss_limit, skipped, ticket);
  return launched("adam_clip");
}

// Sharded form (data parallel, reduce-scatter -> Adam on 1/world of the flat buffers -> all-gather of the parameters):
// phase 1 leaves the 1024 partial sums of squares of THIS rank's gradient shard in ws[0..1024) and a copy of the local loss
// in ws[1024]; the caller all-reduces ws[0..1025) (SUM) so that every rank holds the partials of the whole gradient and the
// summed loss; phase 2 re-reduces them in every block exactly like the unsharded kernel and updates the shard.
int cgvae_grad_sumsq(const float* g, int64_t n, const float* loss, void* ws, size_t ws_bytes, cgvae_stream_t stream) {
  CGVAE_REQUIRE(ws && ws_bytes >= cgvae_adam_ws_bytes() && (n == 0 || g), "grad_sumsq: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + kNormBlocks + 1);
  launch_kernel(sumsq_partial_kernel, dim3(kNormBlocks), dim3(256), 0, st, g, n, partial, ticket);
  if (int rc = launched("sumsq_partial")) return rc;
  if (loss != nullptr) CGVAE_CUDA(cudaMemcpyAsync(partial + kNormBlocks, loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}
int cgvae_adam_apply(float* p, const float* g, float* m, float* v, int64_t n, float max_norm, float grad_scale, float lr, float beta1,
                     float beta2, float eps, float* step, float* norm_out, int use_ws_loss, float loss_scale, float loss_limit,
                     float* skipped, void* ws, size_t ws_bytes, cgvae_stream_t stream) {
  if (n == 0) return 0;
  CGVAE_REQUIRE(p && g && m && v && step && ws && ws_bytes >= cgvae_adam_ws_bytes(), "adam_apply: bad arguments");
  CGVAE_REQUIRE(aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v), "adam_apply: buffers must be 16-byte aligned");
  float* partial = reinterpret_cast<float*>(ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(partial + kNormBlocks + 1);
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 4), 148 * 8);
  launch_kernel(adam_clip_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, p, g, m, v, n, (const float*)partial, max_norm, grad_scale, lr,
                beta1, beta2, eps, step, norm_out, use_ws_loss ? (const float*)(partial + kNormBlocks) : (const float*)nullptr, loss_scale,
                loss_limit, skipped, ticket);
  return launched("adam_clip");
}

}  // extern "C"
