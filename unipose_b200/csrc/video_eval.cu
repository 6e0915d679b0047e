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
                          float* __restrict__ cell, float* __restrict__ hide, float* __restrict__ gates, int cin, int c,
                          int h, int w, int ngroups) {
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
    if (gates) {       // training: activated gates [b][3][c][h][w] for the backward pass
      const long long plane = static_cast<long long>(c) * h * w;
      float* gp = gates + static_cast<long long>(b) * 3 * plane + (static_cast<long long>(co) * h + oy) * w + ox;
      gp[0] = g;
      gp[plane] = i;
      gp[2 * plane] = o;
    }
  }
}

// LSTM: gates g,i,o,f = conv_x(x) + conv_h(h_prev) (+ both biases).
__global__ void __launch_bounds__(kLstmTile * kLstmTile)
    convlstm_cell_kernel(const float* __restrict__ x, const float* __restrict__ hp, const float* __restrict__ cp,
                         const float* __restrict__ wx, const float* __restrict__ bx, const float* __restrict__ wh,
                         const float* __restrict__ bh, float* __restrict__ cell, float* __restrict__ hide,
                         float* __restrict__ gates, int cin, int c, int h, int w, int ngroups) {
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
    const float fg = sigmoidf_(fs), ig = sigmoidf_(is), gg = tanhf(gs), og = sigmoidf_(os);
    const float cl = fg * cp[idx] + ig * gg;
    cell[idx] = cl;
    hide[idx] = og * tanhf(cl);
    if (gates) {       // training: activated gates [b][4][c][h][w] in the order g, i, o, f
      const long long plane = static_cast<long long>(c) * h * w;
      float* gp = gates + static_cast<long long>(b) * 4 * plane + (static_cast<long long>(co) * h + oy) * w + ox;
      gp[0] = gg;
      gp[plane] = ig;
      gp[2 * plane] = og;
      gp[3 * plane] = fg;
    }
  }
}


// ------------------------------------------------------------------------------------------
// ConvLSTM cell backward (what loss.backward() runs through LSTM_0 / LSTM, uniposeLSTM.py:16-24,40-64,132)
// ------------------------------------------------------------------------------------------
// Pointwise part: gradients of the gate PRE-activations (and of c_prev) from dcell / dhide.
//   LSTM_0: cell = tanh(g*i), hide = o*cell              LSTM: cell = f*c_prev + i*g, hide = o*tanh(cell)
// gates: activated [b][G][c][h][w] (g, i, o(, f));  dpre: same shape.
__global__ void convlstm_gate_grad_kernel(const float* __restrict__ gates, const float* __restrict__ cell,
                                          const float* __restrict__ c_prev, const float* __restrict__ dcell,
                                          const float* __restrict__ dhide, float* __restrict__ dpre,
                                          float* __restrict__ dc_prev, long long per_image, int ngates, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const long long b = i / per_image, r = i - b * per_image;
  const float* gp = gates + b * ngates * per_image + r;
  float* dp = dpre + b * ngates * per_image + r;
  const float g = gp[0], ig = gp[per_image], o = gp[2 * per_image];
  const float dh = dhide ? dhide[i] : 0.f;
  const float dc_in = dcell ? dcell[i] : 0.f;
  if (ngates == 3) {
    const float cl = cell[i];
    const float dct = dc_in + dh * o;                 // d loss / d cell
    const float dgi = dct * (1.f - cl * cl);          // through tanh(g*i)
    dp[0] = dgi * ig * (1.f - g * g);
    dp[per_image] = dgi * g * ig * (1.f - ig);
    dp[2 * per_image] = dh * cl * o * (1.f - o);
  } else {
    const float f = gp[3 * per_image];
    const float tc = tanhf(cell[i]);
    const float dct = dc_in + dh * o * (1.f - tc * tc);
    dp[0] = dct * ig * (1.f - g * g);
    dp[per_image] = dct * g * ig * (1.f - ig);
    dp[2 * per_image] = dh * tc * o * (1.f - o);
    dp[3 * per_image] = dct * c_prev[i] * f * (1.f - f);
    dc_prev[i] = dct * f;
  }
}

// dIn[b][ci][y][x] = sum_{gate,co,ky,kx} dpre[b][gate][co][y+1-ky][x+1-kx] * w[gate][co][ci][ky][kx]   (3x3, padding 1)
__global__ void __launch_bounds__(256)
    convlstm_input_grad_kernel(const float* __restrict__ dpre, const float* __restrict__ wts, float* __restrict__ din,
                               int ngates, int c, int cin, int h, int w, long long total) {
  extern __shared__ float wsm[];        // [ngates][c][cin][9]
  const int nw = ngates * c * cin * 9;
  for (int e = threadIdx.x; e < nw; e += blockDim.x) wsm[e] = wts[e];
  __syncthreads();
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % w);
  const int y = static_cast<int>((i / w) % h);
  const int ci = static_cast<int>((i / (static_cast<long long>(w) * h)) % cin);
  const long long b = i / (static_cast<long long>(w) * h * cin);
  const float* dp = dpre + b * ngates * c * h * w;
  float acc = 0.f;
  for (int gc = 0; gc < ngates * c; ++gc) {
    const float* plane = dp + static_cast<long long>(gc) * h * w;
    const float* wp = wsm + (gc * cin + ci) * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + 1 - ky;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + 1 - kx;
        if (xx < 0 || xx >= w) continue;
        acc = fmaf(plane[yy * w + xx], wp[ky * 3 + kx], acc);
      }
    }
  }
  din[i] = acc;
}

// One block per (gate*c + co, ci): dW[gate][co][ci][ky][kx] = sum_{b,y,x} dpre[b][gate][co][y][x] * in[b][ci][y+ky-1][x+kx-1];
// blocks with ci == 0 also produce db[gate][co] = sum dpre.  Fixed-order block reduction: deterministic.
__global__ void __launch_bounds__(256)
    convlstm_weight_grad_kernel(const float* __restrict__ dpre, const float* __restrict__ in, float* __restrict__ dw,
                                float* __restrict__ db, int nb, int ngates, int c, int cin, int h, int w) {
  const int gc = blockIdx.x, ci = blockIdx.y;
  float acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = 0.f;
  const int hw = h * w;
  for (int e = threadIdx.x; e < nb * hw; e += blockDim.x) {
    const int b = e / hw, r = e - b * hw;
    const int y = r / w, x = r - y * w;
    const float d = dpre[(static_cast<long long>(b) * ngates * c + gc) * hw + r];
    const float* ip = in + (static_cast<long long>(b) * cin + ci) * hw;
    acc[9] += d;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= w) continue;
        acc[ky * 3 + kx] = fmaf(d, ip[yy * w + xx], acc[ky * 3 + kx]);
      }
    }
  }
  __shared__ float red[10][256];
#pragma unroll
  for (int t = 0; t < 10; ++t) red[t][threadIdx.x] = acc[t];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int t = 0; t < 10; ++t) red[t][threadIdx.x] += red[t][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < 9) dw[(static_cast<long long>(gc) * cin + ci) * 9 + threadIdx.x] = red[threadIdx.x][0];
  if (threadIdx.x == 9 && ci == 0 && db) db[gc] = red[9][0];
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
                                     int b, int cin, int c, int h, int w, float* gates, void* stream) {
  UP_CHECK_ARG(x && w3 && b3 && cell && hide, "up_convlstm_cell0_fwd: null argument");
  UP_CHECK_ARG(b > 0 && cin > 0 && c > 0 && c <= kLstmCMax, "up_convlstm_cell0_fwd: c must be <= %d", kLstmCMax);
  const int ngroups = (c + kLstmCoG - 1) / kLstmCoG;
  UP_CHECK_ARG(cin <= 64, "up_convlstm_cell0_fwd: cin must be <= 64");
  dim3 grid((w + kLstmTile - 1) / kLstmTile, (h + kLstmTile - 1) / kLstmTile, b * ngroups);
  const size_t wbytes = static_cast<size_t>(3) * kLstmCoG * cin * 9 * sizeof(float);
  convlstm_cell0_kernel<<<grid, kLstmTile * kLstmTile, wbytes, (cudaStream_t)stream>>>(x, w3, b3, cell, hide, gates, cin,
                                                                                      c, h, w, ngroups);
  UP_CHECK_LAUNCH("convlstm_cell0_kernel");
  return 0;
}

extern "C" int up_convlstm_cell_fwd(const float* x, const float* h_prev, const float* c_prev, const float* wx,
                                    const float* bx, const float* wh, const float* bh, float* cell, float* hide, int b,
                                    int cin, int c, int h, int w, float* gates, void* stream) {
  UP_CHECK_ARG(x && h_prev && c_prev && wx && bx && wh && bh && cell && hide, "up_convlstm_cell_fwd: null argument");
  UP_CHECK_ARG(b > 0 && cin > 0 && c > 0 && c <= kLstmCMax, "up_convlstm_cell_fwd: c must be <= %d", kLstmCMax);
  const int ngroups = (c + kLstmCoG - 1) / kLstmCoG;
  UP_CHECK_ARG(cin <= 32, "up_convlstm_cell_fwd: cin must be <= 32 (filters are staged in 48 KB of shared memory)");
  dim3 grid((w + kLstmTile - 1) / kLstmTile, (h + kLstmTile - 1) / kLstmTile, b * ngroups);
  const size_t wbytes = static_cast<size_t>(4) * kLstmCoG * (cin + c) * 9 * sizeof(float);
  convlstm_cell_kernel<<<grid, kLstmTile * kLstmTile, wbytes, (cudaStream_t)stream>>>(
      x, h_prev, c_prev, wx, bx, wh, bh, cell, hide, gates, cin, c, h, w, ngroups);
  UP_CHECK_LAUNCH("convlstm_cell_kernel");
  return 0;
}

extern "C" int up_convlstm_cell_bwd(const float* x, const float* h_prev, const float* c_prev, const float* gates,
                                    const float* cell, const float* dcell, const float* dhide, const float* wx,
                                    const float* wh, float* dx, float* dh_prev, float* dc_prev, float* dwx, float* dbx,
                                    float* dwh, float* dbh, float* dpre, int b, int cin, int c, int h, int w,
                                    void* stream) {
  UP_CHECK_ARG(x && gates && cell && wx && dx && dwx && dbx && dpre, "up_convlstm_cell_bwd: null argument");
  UP_CHECK_ARG(dcell || dhide, "up_convlstm_cell_bwd: at least one of dcell / dhide is required");
  const bool full = wh != nullptr;      // LSTM (4 gates, recurrent inputs) vs LSTM_0 (3 gates)
  if (full) UP_CHECK_ARG(h_prev && c_prev && dh_prev && dc_prev && dwh && dbh, "up_convlstm_cell_bwd: recurrent arguments missing");
  UP_CHECK_ARG(b > 0 && cin > 0 && c > 0 && c <= kLstmCMax && cin <= 32, "up_convlstm_cell_bwd: bad dims");
  const int ng = full ? 4 : 3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long per_image = static_cast<long long>(c) * h * w;
  const long long total = per_image * b;
  convlstm_gate_grad_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(gates, cell, c_prev, dcell, dhide, dpre,
                                                                                  dc_prev, per_image, ng, total);
  UP_CHECK_LAUNCH("convlstm_gate_grad_kernel");
  {
    const long long tx = static_cast<long long>(b) * cin * h * w;
    const size_t wb = static_cast<size_t>(ng) * c * cin * 9 * sizeof(float);
    UP_CHECK_ARG(wb <= 48 * 1024, "up_convlstm_cell_bwd: filters exceed 48 KB of shared memory");
    convlstm_input_grad_kernel<<<static_cast<int>((tx + 255) / 256), 256, wb, st>>>(dpre, wx, dx, ng, c, cin, h, w, tx);
    UP_CHECK_LAUNCH("convlstm_input_grad_kernel(x)");
    convlstm_weight_grad_kernel<<<dim3(ng * c, cin), 256, 0, st>>>(dpre, x, dwx, dbx, b, ng, c, cin, h, w);
    UP_CHECK_LAUNCH("convlstm_weight_grad_kernel(x)");
  }
  if (full) {
    const long long th = static_cast<long long>(b) * c * h * w;
    const size_t wb = static_cast<size_t>(ng) * c * c * 9 * sizeof(float);
    convlstm_input_grad_kernel<<<static_cast<int>((th + 255) / 256), 256, wb, st>>>(dpre, wh, dh_prev, ng, c, c, h, w, th);
    UP_CHECK_LAUNCH("convlstm_input_grad_kernel(h)");
    convlstm_weight_grad_kernel<<<dim3(ng * c, c), 256, 0, st>>>(dpre, h_prev, dwh, dbh, b, ng, c, c, h, w);
    UP_CHECK_LAUNCH("convlstm_weight_grad_kernel(h)");
  }
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
