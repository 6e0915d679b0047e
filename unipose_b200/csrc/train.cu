// Training glue of the hot path (unipose.py:113-124): MSE loss forward+backward and the Adam update,
// each one pass over flat fp32 buffers.
#include "up_internal.h"

namespace up {

// loss (mean) accumulated in double via one atomicAdd per block; grad written in the same pass.
__global__ void mse_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                   double* __restrict__ loss_acc, float* __restrict__ grad, long long count,
                                   float gcoef) {
  double local = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = pred[i] - target[i];
    local += static_cast<double>(d) * d;
    if (grad) grad[i] = gcoef * d;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
  __shared__ double s[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s[warp] = local;
  __syncthreads();
  if (warp == 0) {
    local = lane < (blockDim.x >> 5) ? s[lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
    if (lane == 0) atomicAdd(loss_acc, local);
  }
}

__global__ void mse_finish_kernel(const double* __restrict__ loss_acc, float* __restrict__ loss, long long count) {
  loss[0] = static_cast<float>(loss_acc[0] / static_cast<double>(count));
}

// torch.optim.Adam (no weight decay, no amsgrad, eps added after the bias-corrected sqrt):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long count, float b1, float b2, float eps, float step_size,
                            float inv_sqrt_bc2) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= count) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
  p[i] -= step_size * (mi / denom);
}

}  // namespace up

using namespace up;

extern "C" int up_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* grad, double* scratch,
                              int64_t count, float gscale, void* stream) {
  UP_CHECK_ARG(pred && target && loss && scratch && count > 0, "up_mse_fwd_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(scratch, 0, sizeof(double), st), "cudaMemsetAsync(loss scratch)");
  if (rc) return rc;
  long long blocks = (count + 255) / 256;
  if (blocks > 2048) blocks = 2048;
  mse_fwd_bwd_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(pred, target, scratch, grad, count,
                                                               2.0f * gscale / static_cast<float>(count));
  UP_CHECK_LAUNCH("mse_fwd_bwd_kernel");
  mse_finish_kernel<<<1, 1, 0, st>>>(scratch, loss, count);
  UP_CHECK_LAUNCH("mse_finish_kernel");
  return 0;
}

extern "C" int up_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count,
                            float lr, float beta1, float beta2, float eps, int step, void* stream) {
  UP_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && count > 0 && step >= 1, "up_adam_step: bad argument");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  adam_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, count, beta1, beta2, eps, step_size, inv_sqrt_bc2);
  UP_CHECK_LAUNCH("adam_kernel");
  return 0;
}
