// Gradient step of data-parallel training on ONE flat fp32 bucket (the same bucket NCCL all-reduces):
// global-norm clip (model.py:645-650, gradMaxNorm), Adam (model.py:618) and the EMA shadow weights (model.py:658-667),
// fused into one pass.  The norm is a fixed-order two-stage reduction; its clip factor stays on the device.
#include "common.cuh"

using namespace mac;

namespace mac {
constexpr int NORM_BLOCKS = 592;   // 4 per SM

__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ g, long long n,
                                                            float* __restrict__ partial) {
  __shared__ float s_red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float v = g[i];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s_red[i];
    partial[blockIdx.x] = t;
  }
}

// out[0] = global norm, out[1] = clip factor = max_norm / max(norm, max_norm)   (tf.clip_by_global_norm)
__global__ void norm_finalize_kernel(const float* __restrict__ partial, int nblocks, float grad_scale, float max_norm,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nblocks; ++i) t += (double)partial[i];
  const float norm = sqrtf((float)t) * grad_scale;
  out[0] = norm;
  out[1] = max_norm > 0.f ? max_norm / fmaxf(norm, max_norm) : 1.f;
}

__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, float* __restrict__ ema, const float* __restrict__ norm_out,
                                float grad_scale, float lr, float b1, float b2, float eps, float bc1, float bc2,
                                float ema_decay, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale * norm_out[1];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  // TF AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  p -= lr_t * m / (sqrt(v) + eps)
  const float lr_t = lr * sqrtf(bc2) / bc1;
  const float pn = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  p[i] = pn;
  if (ema) ema[i] = ema_decay * ema[i] + (1.f - ema_decay) * pn;   // tf.train.ExponentialMovingAverage.apply
}
}  // namespace mac

extern "C" size_t mac_optimizer_workspace_bytes(void) { return (NORM_BLOCKS + 8) * sizeof(float); }

extern "C" int mac_clip_adam_ema_step(float* params, const float* grads, float* adam_m, float* adam_v, float* ema,
                                      long long n, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                      float eps, int step, float ema_decay, float* norm_out, void* workspace,
                                      size_t workspace_bytes, mac_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!params || !grads || !adam_m || !adam_v || !norm_out || !workspace || n <= 0 || step < 1) return MAC_ERR_INVALID;
  if (workspace_bytes < mac_optimizer_workspace_bytes()) return MAC_ERR_WORKSPACE;
  float* partial = reinterpret_cast<float*>(workspace);
  sqnorm_partial_kernel<<<NORM_BLOCKS, 256, 0, stream>>>(grads, n, partial);
  MAC_LAUNCH_CHECK();
  norm_finalize_kernel<<<1, 32, 0, stream>>>(partial, NORM_BLOCKS, grad_scale, max_norm, norm_out);
  MAC_LAUNCH_CHECK();
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adam_ema_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(params, grads, adam_m, adam_v, ema, norm_out, grad_scale,
                                                                  lr, beta1, beta2, eps, bc1, bc2, ema_decay, n);
  MAC_LAUNCH_CHECK();
  return MAC_OK;
}
