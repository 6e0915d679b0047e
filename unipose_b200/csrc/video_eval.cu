// Latency-bound fp32 kernels of the video variant (ConvLSTM cells, centre-map pooling) and of the
// evaluation path (per-joint arg-max, PCK distances).  All tensors here are the reference's own fp32 NCHW
// user-facing tensors, so results are fp32-exact up to summation order.
#include "up_internal.h"

namespace up {

// ------------------------------------------------------------------------------------------
// nn.AvgPool2d(kernel_size=9, stride=8, padding=1), count_include_pad=True (divide by 81 always)
// ------------------------------------------------------------------------------------------
__global__ void avgpool9s8p1_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int c, int h, int w,
                                    int ho, int wo, int yct, int ycoff) {
  const long long total = static_cast<long long>(n) * c * ho * wo;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ox = static_cast<int>(i % wo);
  long long t = i / wo;
  const int oy = static_cast<int>(t % ho);
  t /= ho;
  const int ch = static_cast<int>(t % c);
  const int b = static_cast<int>(t / c);
  const float* p = x + (static_cast<long long>(b) * c + ch) * h * w;
  float s = 0.f;
  for (int dy = 0; dy < 9; ++dy) {
    const int iy = oy * 8 - 1 + dy;
    if (iy < 0 || iy >= h) continue;
    for (int dx = 0; dx < 9; ++dx) {
      const int ix = ox * 8 - 1 + dx;
      if (ix < 0 || ix >= w) continue;
      s += p[iy * w + ix];
    }
  }
  y[((static_cast<long long>(b) * yct + ycoff + ch) * ho + oy) * wo + ox] = s / 81.0f;
}

// ------------------------------------------------------------------------------------------
// ConvLSTM cells: all gate convolutions (3x3, pad 1, with bias) + gate non-linearities + state update
// in one kernel.  A block owns a 16x16 pixel tile of one image and a group of kLstmCoG output channels (all gates
// of those channels, so the point-wise update stays local): batch 8 x 9 tiles x 3 channel groups = 216 blocks.
// One thread per pixel keeps kGates x kLstmCoG accumulators IN REGISTERS (compile-time trip counts); the input tile
// (+halo) and the block's slice of the filters are staged in shared memory (broadcast reads).
// Accumulation order per output = (input channel, tap), the same as a direct fp32 convolution loop.
// ------------------------------------------------------------------------------------------
constexpr int kLstmTile = 16;
constexpr int kLstmCMax = 16;
constexpr int kLstmCoG = 5;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// wsm: [kGates][kLstmCoG][cin][9] slice of wts [kGates][c][cin][3][3] for channels co0 .. co0+kLstmCoG-1 (zeros beyond c)
template <int kGates>
__device__ __forceinline__ void lstm_stage_weights(const float* __restrict__ wts, float* wsm, int cin, int c, int co0) {
  const int total = kGates * kLstmCoG * cin * 9;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int t = e % 9;
    int r = e / 9;
    const int ci = r % cin;
    r /= cin;
    const int j = r % kLstmCoG, g = r / kLstmCoG;
    const int co = co0 + j;
    wsm[e] = co < c ? wts[((static_cast<long long>(g) * c + co) * cin + ci) * 9 + t] : 0.f;
  }
}

template <int kGates>
__device__ __forceinline__ void lstm_accumulate(const float* __restrict__ src, int cin, int h, int w,
                                                const float* wsm, float (&acc)[kGates][kLstmCoG], float* tile, int ty0,
                                                int tx0) {
  // src: [cin, h, w] of one image
  constexpr int tw = kLstmTile + 2;
  const int ly = threadIdx.x / kLstmTile, lx = threadIdx.x % kLstmTile;
  for (int ci = 0; ci < cin; ++ci) {
    __syncthreads();
    for (int e = threadIdx.x; e < tw * tw; e += blockDim.x) {
      const int yy = ty0 - 1 + e / tw, xx = tx0 - 1 + e % tw;
      tile[e] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? src[(static_cast<long long>(ci) * h + yy) * w + xx] : 0.f;
    }
    __syncthreads();
    float v[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = tile[(ly + t / 3) * tw + lx + t % 3];
#pragma unroll
    for (int g = 0; g < kGates; ++g) {
#pragma unroll
      for (int j = 0; j < kLstmCoG; ++j) {
        const float* wp = wsm + ((g * kLstmCoG + j) * cin + ci) * 9;
        float a = acc[g][j];
#pragma unroll
        for (int t = 0; t < 9; ++t) a = fmaf(wp[t], v[t], a);
        acc[g][j] = a;
      }
    }
  }
}

// LSTM_0: gates g,i,o from x only.
__global__ void __launch_bounds__(kLstmTile * kLstmTile)
    convlstm_cell0_kernel(const float* __restrict__ x, const float* __restrict__ w3, const float* __restrict__ b3,
                          float* __restrict__ cell, float* __restrict__ hide, int cin, int c, int h, int w, int ngroups) {
  __shared__ float tile[(kLstmTile + 2) * (kLstmTile + 2)];
  extern __shared__ float wsm[];
  const int b = blockIdx.z / ngroups, co0 = (blockIdx.z % ngroups) * kLstmCoG;
  const int ty0 = blockIdx.y * kLstmTile, tx0 = blockIdx.x * kLstmTile;
  lstm_stage_weights<3>(w3, wsm, cin, c, co0);
  float acc[3][kLstmCoG];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int j = 0; j < kLstmCoG; ++j) acc[g][j] = 0.f;
  lstm_accumulate<3>(x + static_cast<long long>(b) * cin * h * w, cin, h, w, wsm, acc, tile, ty0, tx0);
  const int oy = ty0 + threadIdx.x / kLstmTile, ox = tx0 + threadIdx.x % kLstmTile;
  if (oy >= h || ox >= w) return;
#pragma unroll
  for (int j = 0; j < kLstmCoG; ++j) {
    const int co = co0 + j;
    if (co >= c) break;
    const float g = tanhf(acc[0][j] + b3[0 * c + co]);
    const float i = sigmoidf_(acc[1][j] + b3[1 * c + co]);
    const float o = sigmoidf_(acc[2][j] + b3[2 * c + co]);
    const float cl = tanhf(g * i);
    const long long idx = ((static_cast<long long>(b) * c + co) * h + oy) * w + ox;
    cell[idx] = cl;
    hide[idx] = o * cl;
  }
}

// LSTM: gates g,i,o,f = conv_x(x) + conv_h(h_prev) (+ both biases).
__global__ void __launch_bounds__(kLstmTile * kLstmTile)
    convlstm_cell_kernel(const float* __restrict__ x, const float* __restrict__ hp, const float* __restrict__ cp,
                         const float* __restrict__ wx, const float* __restrict__ bx, const float* __restrict__ wh,
                         const float* __restrict__ bh, float* __restrict__ cell, float* __restrict__ hide, int cin,
                         int c, int h, int w, int ngroups) {
  __shared__ float tile[(kLstmTile + 2) * (kLstmTile + 2)];
  extern __shared__ float wsm[];   // [x filters | h filters]
  const int b = blockIdx.z / ngroups, co0 = (blockIdx.z % ngroups) * kLstmCoG;
  const int ty0 = blockIdx.y * kLstmTile, tx0 = blockIdx.x * kLstmTile;
  float* wsm_h = wsm + 4 * kLstmCoG * cin * 9;
  lstm_stage_weights<4>(wx, wsm, cin, c, co0);
  lstm_stage_weights<4>(wh, wsm_h, c, c, co0);
  float acc[4][kLstmCoG], acch[4][kLstmCoG];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int j = 0; j < kLstmCoG; ++j) acc[g][j] = acch[g][j] = 0.f;
  lstm_accumulate<4>(x + static_cast<long long>(b) * cin * h * w, cin, h, w, wsm, acc, tile, ty0, tx0);
  lstm_accumulate<4>(hp + static_cast<long long>(b) * c * h * w, c, h, w, wsm_h, acch, tile, ty0, tx0);
  const int oy = ty0 + threadIdx.x / kLstmTile, ox = tx0 + threadIdx.x % kLstmTile;
  if (oy >= h || ox >= w) return;
#pragma unroll
  for (int j = 0; j < kLstmCoG; ++j) {
    const int co = co0 + j;
    if (co >= c) break;
    // same association as the reference: (conv_x + bias_x) + (conv_h + bias_h)
    const float gs = (acc[0][j] + bx[0 * c + co]) + (acch[0][j] + bh[0 * c + co]);
    const float is = (acc[1][j] + bx[1 * c + co]) + (acch[1][j] + bh[1 * c + co]);
    const float os = (acc[2][j] + bx[2 * c + co]) + (acch[2][j] + bh[2 * c + co]);
    const float fs = (acc[3][j] + bx[3 * c + co]) + (acch[3][j] + bh[3 * c + co]);
    const long long idx = ((static_cast<long long>(b) * c + co) * h + oy) * w + ox;
    const float cl = sigmoidf_(fs) * cp[idx] + sigmoidf_(is) * tanhf(gs);
    cell[idx] = cl;
    hide[idx] = sigmoidf_(os) * tanhf(cl);
  }
}

// ------------------------------------------------------------------------------------------
// evaluation
// ------------------------------------------------------------------------------------------
// numpy argmax semantics: first occurrence of the maximum in row-major order; a NaN beats every number.
__device__ __forceinline__ bool arg_better(float v, int i, float bv, int bi) {
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (vn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

__global__ void argmax2d_kernel(const float* __restrict__ heat, int32_t* __restrict__ idx, float* __restrict__ preds,
                                float* __restrict__ maxvals, int hw, int w) {
  const long long map = blockIdx.x;
  const float* p = heat + map * hw;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    const float v = p[i];
    if (bi == 0x7fffffff || arg_better(v, i, bv, bi)) {
      bv = v;
      bi = i;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, bv, off);
    const int oi = __shfl_down_sync(0xffffffffu, bi, off);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || arg_better(ov, oi, bv, bi))) {
      bv = ov;
      bi = oi;
    }
  }
  __shared__ float sv[32];
  __shared__ int si[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sv[warp] = bv;
    si[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    bv = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ov = __shfl_down_sync(0xffffffffu, bv, off);
      const int oi = __shfl_down_sync(0xffffffffu, bi, off);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || arg_better(ov, oi, bv, bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      idx[map] = bi;
      maxvals[map] = bv;
      const float m = bv > 0.0f ? 1.0f : 0.0f;  // np.greater(maxvals, 0.0): False for NaN
      preds[map * 2 + 0] = static_cast<float>(bi % w) * m;
      preds[map * 2 + 1] = static_cast<float>(bi / w) * m;
    }
  }
}

__global__ void calc_dists_kernel(const float* __restrict__ preds, const float* __restrict__ target,
                                  double* __restrict__ dists, int n, int k, double nx, double ny) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * k) return;
  const int b = i / k, c = i % k;
  const float tx = target[i * 2 + 0], ty = target[i * 2 + 1];
  double d = -1.0;
  if (tx > 1.0f && ty > 1.0f) {
    const double dx = static_cast<double>(preds[i * 2 + 0]) / nx - static_cast<double>(tx) / nx;
    const double dy = static_cast<double>(preds[i * 2 + 1]) / ny - static_cast<double>(ty) / ny;
    d = sqrt(dx * dx + dy * dy);
  }
  dists[static_cast<long long>(c) * n + b] = d;
}

__global__ void dist_acc_kernel(const double* __restrict__ dists, double* __restrict__ acc, int n, int k,
                                double threshold) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  int valid = 0, hit = 0;
  for (int b = 0; b < n; ++b) {
    const double d = dists[static_cast<long long>(c) * n + b];
    if (d != -1.0) {
      ++valid;
      hit += d < threshold ? 1 : 0;
    }
  }
  acc[c] = valid > 0 ? static_cast<double>(hit) * 1.0 / static_cast<double>(valid) : -1.0;
}

}  // namespace up

using namespace up;

extern "C" int up_avgpool9s8p1_f32(const float* x, float* y, int n, int c, int h, int w, int ho, int wo,
                                   int y_c_total, int y_c_off, void* stream) {
  UP_CHECK_ARG(x && y && n > 0 && c > 0, "up_avgpool9s8p1_f32: bad argument");
  UP_CHECK_ARG(ho == (h + 2 - 9) / 8 + 1 && wo == (w + 2 - 9) / 8 + 1, "up_avgpool9s8p1_f32: ho/wo mismatch");
  if (y_c_total <= 0) y_c_total = c;
  UP_CHECK_ARG(y_c_off >= 0 && y_c_off + c <= y_c_total, "up_avgpool9s8p1_f32: bad output channel slice");
  const long long total = static_cast<long long>(n) * c * ho * wo;
  avgpool9s8p1_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, c, h, w, ho,
                                                                                                wo, y_c_total, y_c_off);
  UP_CHECK_LAUNCH("avgpool9s8p1_kernel");
  return 0;
}

extern "C" int up_convlstm_cell0_fwd(const float* x, const float* w3, const float* b3, float* cell, float* hide,
                                     int b, int cin, int c, int h, int w, void* stream) {
  UP_CHECK_ARG(x && w3 && b3 && cell && hide, "up_convlstm_cell0_fwd: null argument");
  UP_CHECK_ARG(b > 0 && cin > 0 && c > 0 && c <= kLstmCMax, "up_convlstm_cell0_fwd: c must be <= %d", kLstmCMax);
  const int ngroups = (c + kLstmCoG - 1) / kLstmCoG;
  UP_CHECK_ARG(cin <= 64, "up_convlstm_cell0_fwd: cin must be <= 64");
  dim3 grid((w + kLstmTile - 1) / kLstmTile, (h + kLstmTile - 1) / kLstmTile, b * ngroups);
  const size_t wbytes = static_cast<size_t>(3) * kLstmCoG * cin * 9 * sizeof(float);
  convlstm_cell0_kernel<<<grid, kLstmTile * kLstmTile, wbytes, (cudaStream_t)stream>>>(x, w3, b3, cell, hide, cin, c, h,
                                                                                      w, ngroups);
  UP_CHECK_LAUNCH("convlstm_cell0_kernel");
  return 0;
}

extern "C" int up_convlstm_cell_fwd(const float* x, const float* h_prev, const float* c_prev, const float* wx,
                                    const float* bx, const float* wh, const float* bh, float* cell, float* hide, int b,
                                    int cin, int c, int h, int w, void* stream) {
  UP_CHECK_ARG(x && h_prev && c_prev && wx && bx && wh && bh && cell && hide, "up_convlstm_cell_fwd: null argument");
  UP_CHECK_ARG(b > 0 && cin > 0 && c > 0 && c <= kLstmCMax, "up_convlstm_cell_fwd: c must be <= %d", kLstmCMax);
  const int ngroups = (c + kLstmCoG - 1) / kLstmCoG;
  UP_CHECK_ARG(cin <= 32, "up_convlstm_cell_fwd: cin must be <= 32 (filters are staged in 48 KB of shared memory)");
  dim3 grid((w + kLstmTile - 1) / kLstmTile, (h + kLstmTile - 1) / kLstmTile, b * ngroups);
  const size_t wbytes = static_cast<size_t>(4) * kLstmCoG * (cin + c) * 9 * sizeof(float);
  convlstm_cell_kernel<<<grid, kLstmTile * kLstmTile, wbytes, (cudaStream_t)stream>>>(x, h_prev, c_prev, wx, bx, wh, bh,
                                                                                     cell, hide, cin, c, h, w, ngroups);
  UP_CHECK_LAUNCH("convlstm_cell_kernel");
  return 0;
}

extern "C" int up_argmax2d(const float* heat, int32_t* idx, float* preds, float* maxvals, int n, int k, int h, int w,
                           void* stream) {
  UP_CHECK_ARG(heat && idx && preds && maxvals, "up_argmax2d: null argument");
  UP_CHECK_ARG(n > 0 && k > 0 && h > 0 && w > 0 && static_cast<long long>(h) * w < 0x7fffffffLL, "up_argmax2d: bad dims");
  argmax2d_kernel<<<n * k, 256, 0, (cudaStream_t)stream>>>(heat, idx, preds, maxvals, h * w, w);
  UP_CHECK_LAUNCH("argmax2d_kernel");
  return 0;
}

extern "C" int up_calc_dists(const float* preds, const float* target, double* dists, int n, int k, double norm_x,
                             double norm_y, void* stream) {
  UP_CHECK_ARG(preds && target && dists && n > 0 && k > 0, "up_calc_dists: bad argument");
  calc_dists_kernel<<<(n * k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(preds, target, dists, n, k, norm_x, norm_y);
  UP_CHECK_LAUNCH("calc_dists_kernel");
  return 0;
}

extern "C" int up_dist_acc(const double* dists, double* acc, int n, int k, double threshold, void* stream) {
  UP_CHECK_ARG(dists && acc && n > 0 && k > 0, "up_dist_acc: bad argument");
  dist_acc_kernel<<<(k + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dists, acc, n, k, threshold);
  UP_CHECK_LAUNCH("dist_acc_kernel");
  return 0;
}
