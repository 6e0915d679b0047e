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

// =============================================================================================
// Train-mode BatchNorm + backward bandwidth kernels (NHWC 16-bit activations / gradients)
// =============================================================================================
namespace up {

// A loaded channel octet stays packed (4 registers, 8 in split mode) until it is used: the streaming kernels below keep
// several octets per thread in flight, and 8 converted floats per octet would halve the occupancy.
template <int kMode>
struct Raw8 {
  uint4 h, l;   // l: the low plane of the bf16 pair format (unused otherwise)
};

template <int kMode>
__device__ __forceinline__ Raw8<kMode> t_load_raw(const uint16_t* p, long long plane) {
  Raw8<kMode> r;
  r.h = __ldg(reinterpret_cast<const uint4*>(p));
  if constexpr (kMode == 2) r.l = __ldg(reinterpret_cast<const uint4*>(p + plane));
  else r.l = make_uint4(0, 0, 0, 0);
  return r;
}

template <int kMode>
__device__ __forceinline__ Raw8<kMode> raw_zero() {
  Raw8<kMode> r;
  r.h = make_uint4(0, 0, 0, 0);
  r.l = r.h;
  return r;
}

template <int kMode>
__device__ __forceinline__ void t_cvt8(const Raw8<kMode>& r, float (&v)[8]) {
  const uint32_t w[4] = {r.h.x, r.h.y, r.h.z, r.h.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e + 0] = cvt16_to_f32<(kMode == 0 ? 0 : 1)>(static_cast<uint16_t>(w[e] & 0xFFFFu));
    v[2 * e + 1] = cvt16_to_f32<(kMode == 0 ? 0 : 1)>(static_cast<uint16_t>(w[e] >> 16));
  }
  if constexpr (kMode == 2) {
    const uint32_t lw[4] = {r.l.x, r.l.y, r.l.z, r.l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e + 0] += cvt16_to_f32<1>(static_cast<uint16_t>(lw[e] & 0xFFFFu));
      v[2 * e + 1] += cvt16_to_f32<1>(static_cast<uint16_t>(lw[e] >> 16));
    }
  }
}

template <int kMode>
__device__ __forceinline__ void t_load8(const uint16_t* p, long long plane, float (&v)[8]) {
  t_cvt8<kMode>(t_load_raw<kMode>(p, plane), v);
}

__device__ __forceinline__ void load_f8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

template <int kMode>
__device__ __forceinline__ void t_store8(uint16_t* p, long long plane, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if constexpr (kMode == 2) {
      uint16_t h0, l0, h1, l1;
      split_bf16(v[2 * e], h0, l0);
      split_bf16(v[2 * e + 1], h1, l1);
      h[e] = h0 | (uint32_t(h1) << 16);
      l[e] = l0 | (uint32_t(l1) << 16);
    } else {
      h[e] = cvt_f32_to16<kMode>(v[2 * e]) | (uint32_t(cvt_f32_to16<kMode>(v[2 * e + 1])) << 16);
    }
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  if constexpr (kMode == 2) *reinterpret_cast<uint4*>(p + plane) = make_uint4(l[0], l[1], l[2], l[3]);
}

struct TView {           // NHWC channel-slice view
  const uint16_t* p;     // base of the buffer (plane 0)
  int cs, coff;
  long long plane;
};
struct TViewW {
  uint16_t* p;
  int cs, coff;
  long long plane;
};

// Per-channel reduction skeleton: every thread owns one channel octet (c/8 is a power of two <= 256) and strides
// over pixels; partial sums are combined in shared memory and added to the global double accumulators.
// kWhat 0: sums of (x, x^2).  kWhat 1: sums of (dy', dy' * xhat) with dy' = dy * (y > 0 if relu).
template <int kMode, int kWhat>
__global__ void __launch_bounds__(512)
    channel_reduce_kernel(TView a, TView b, TView c, const float* __restrict__ mean, const float* __restrict__ invstd,
                          float* __restrict__ rows, long long npix, int ch, int relu) {
  pdl_enter();
  const int octs = ch / 8;
  const int oct = threadIdx.x % octs;
  const int pstride = blockDim.x / octs;
  const int plane_lane = threadIdx.x / octs;
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  float mu[8], is[8];
  if (kWhat == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = mean[oct * 8 + e];
      is[e] = invstd[oct * 8 + e];
    }
  }
  const long long per_block = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per_block;
  const long long p1 = min(p0 + per_block, npix);
  // pixels in flight per thread: the loads of one iteration are all issued before any use and stay packed until then
  // (measured with tools/bn_microbench.cu: 4 in flight, 512 threads: 5.0-6.4 TB/s on the three-tensor reduction)
#ifndef UP_RED_UNROLL
#define UP_RED_UNROLL 4
#endif
  constexpr int kUnroll = kMode == 2 ? 2 : UP_RED_UNROLL;
  for (long long px0 = p0 + plane_lane; px0 < p1; px0 += static_cast<long long>(kUnroll) * pstride) {
    Raw8<kMode> ra[kUnroll], rz[kUnroll], ry[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long px = px0 + static_cast<long long>(u) * pstride;
      ra[u] = raw_zero<kMode>();     // a zero octet adds nothing to either sum
      rz[u] = ra[u];
      ry[u] = ra[u];
      if (px < p1) {
        ra[u] = t_load_raw<kMode>(a.p + px * a.cs + a.coff + oct * 8, a.plane);
        if (kWhat == 1) {
          rz[u] = t_load_raw<kMode>(c.p + px * c.cs + c.coff + oct * 8, c.plane);
          if (relu) ry[u] = t_load_raw<kMode>(b.p + px * b.cs + b.coff + oct * 8, b.plane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float va[8];
      t_cvt8<kMode>(ra[u], va);
      if (kWhat == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s0[e] += va[e];
          s1[e] = fmaf(va[e], va[e], s1[e]);
        }
      } else {
        float vz[8], vy[8];
        t_cvt8<kMode>(rz[u], vz);
        if (relu) t_cvt8<kMode>(ry[u], vy);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float d = va[e];
          if (relu) d = vy[e] > 0.f ? d : 0.f;
          s0[e] += d;
          s1[e] = fmaf(d, (vz[e] - mu[e]) * is[e], s1[e]);
        }
      }
    }
  }
  extern __shared__ float red[];  // [blockDim.x][16]
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[threadIdx.x * 16 + e] = s0[e];
    red[threadIdx.x * 16 + 8 + e] = s1[e];
  }
  __syncthreads();
  // every block owns one row of partial sums (no atomics, deterministic); reduce_rows_kernel adds the rows up
  for (int idx = threadIdx.x; idx < octs * 16; idx += blockDim.x) {
    const int o = idx / 16, e = idx % 16;
    double s = 0.0;
    for (int q = 0; q < pstride; ++q) s += red[(q * octs + o) * 16 + e];
    const int chn = o * 8 + (e & 7);
    rows[static_cast<long long>(blockIdx.x) * 2 * ch + (e < 8 ? 0 : ch) + chn] = static_cast<float>(s);
  }
}

// blockDim = (32, 16): 32 consecutive sums per block, the rows split 16 ways, combined in shared memory in a fixed
// order (deterministic).
__global__ void reduce_rows_kernel(const float* __restrict__ rows, int nrows, int c2, double* __restrict__ sums) {
  pdl_enter();
  __shared__ double part[16][33];
  const int i = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (i < c2) {
    for (int r = threadIdx.y; r < nrows; r += 16) s += rows[static_cast<long long>(r) * c2 + i];
  }
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && i < c2) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += part[q][threadIdx.x];
    sums[i] = t;
  }
}

struct BnFinalizeArgs {
  double count;
  const float *gamma, *beta;
  float *rmean, *rvar;
  float momentum, eps;
  float *scale, *shift, *save_mean, *save_invstd;
  int c_real, c;
};

__device__ __forceinline__ void bn_finalize_channel(int i, double sum, double sumsq, const BnFinalizeArgs& a) {
  if (i >= a.c_real) {
    a.scale[i] = 0.f;
    a.shift[i] = 0.f;
    a.save_mean[i] = 0.f;
    a.save_invstd[i] = 0.f;
    return;
  }
  const double mean = sum / a.count;
  double var = sumsq / a.count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps)));
  const float sc = a.gamma[i] * invstd;
  a.scale[i] = sc;
  a.shift[i] = a.beta[i] - static_cast<float>(mean) * sc;
  a.save_mean[i] = static_cast<float>(mean);
  a.save_invstd[i] = invstd;
  if (a.rmean) {
    const double unbiased = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
    a.rmean[i] = (1.f - a.momentum) * a.rmean[i] + a.momentum * static_cast<float>(mean);
    a.rvar[i] = (1.f - a.momentum) * a.rvar[i] + a.momentum * static_cast<float>(unbiased);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const BnFinalizeArgs a) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.c) return;
  bn_finalize_channel(i, sums[i], sums[a.c + i], a);
}

// k1 = gamma*invstd, k2 = -k1*invstd*m1, k3 = -k1*m0 - k2*mean with m0 = sum_dy/M, m1 = sum_dy_xhat/M; padded
// channels get zeros.  coef = [k1 | k2 | k3], c floats each.
// frozen (eval-mode BatchNorm inside a training step: the statistics are constants): dz = gamma*invstd*dy'.
struct BnBwdCoefArgs {
  double count;
  const float *mean, *invstd, *gamma;
  float *coef, *dgamma, *dbeta;
  int c_real, c, frozen;
};

__device__ __forceinline__ void bn_bwd_coef_channel(int i, double sum_dy, double sum_dy_xhat, const BnBwdCoefArgs& a) {
  double k1 = 0.0, k2 = 0.0, k3 = 0.0;
  if (i < a.c_real) {
    const double m0 = sum_dy / a.count, m1 = sum_dy_xhat / a.count;
    k1 = static_cast<double>(a.gamma[i]) * a.invstd[i];
    if (!a.frozen) {
      k2 = -k1 * a.invstd[i] * m1;
      k3 = -k1 * m0 - k2 * a.mean[i];
    }
    if (a.dgamma) {
      a.dbeta[i] = static_cast<float>(sum_dy);
      a.dgamma[i] = static_cast<float>(sum_dy_xhat);
    }
  }
  a.coef[i] = static_cast<float>(k1);
  a.coef[a.c + i] = static_cast<float>(k2);
  a.coef[2 * a.c + i] = static_cast<float>(k3);
}

// The partial rows of a per-channel reduction -> the two sums of every channel (same order as reduce_rows_kernel) ->
// what the streaming kernel that follows needs, in ONE launch: the forward statistics and epilogue constants
// (kWhat 0) or the backward coefficients and dgamma / dbeta (kWhat 1).  blockDim = (32, 16).
template <int kWhat, class Args>
__global__ void bn_finish_kernel(const float* __restrict__ rows, int nrows, double* __restrict__ sums, const Args a) {
  pdl_enter();
  __shared__ double part[2][16][33];
  const int c = a.c;
  const int i = blockIdx.x * 32 + threadIdx.x;
  double s0 = 0.0, s1 = 0.0;
  if (i < c) {
    for (int r = threadIdx.y; r < nrows; r += 16) {
      s0 += rows[static_cast<long long>(r) * 2 * c + i];
      s1 += rows[static_cast<long long>(r) * 2 * c + c + i];
    }
  }
  part[0][threadIdx.y][threadIdx.x] = s0;
  part[1][threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && i < c) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      t0 += part[0][q][threadIdx.x];
      t1 += part[1][q][threadIdx.x];
    }
    sums[i] = t0;
    sums[c + i] = t1;
    if constexpr (kWhat == 0) bn_finalize_channel(i, t0, t1, a);
    else bn_bwd_coef_channel(i, t0, t1, a);
  }
}

// y = [relu]( z * scale[c] + shift[c] (+ res) ) (* mask)
// kElemUnroll octets per thread, `stride` (= total threads) apart, all loads issued before the first use.  More than one
// octet per thread measured SLOWER on B200 (occupancy drops faster than memory-level parallelism rises): default 1.
// The two streaming BatchNorm maps come in two forms.  "fast": c/8 is a power of two <= 256, so a block of 256 threads
// owning 256*kU consecutive octets gives every thread ONE channel octet for all its kU octets: the per-channel
// constants are loaded once into registers, the kU loads per tensor are issued back to back and stay packed.
// Measured (tools/bn_microbench.cu, bf16, B200): 5.8-6.9 TB/s against 1.7-2.9 TB/s for the one-octet-per-thread form,
// whose 16 scalar constant loads per octet were the bottleneck.  "generic": any c % 8 == 0, one octet per thread.
template <int kMode, int kU>
__global__ void __launch_bounds__(256)
    scale_shift_act_fast_kernel(TView z, TViewW y, TView res, TView mask, const float* __restrict__ scale,
                                const float* __restrict__ shift, long long total, int lg, int relu, int has_res,
                                int has_mask) {
  pdl_enter();
  const int g = threadIdx.x & ((1 << lg) - 1);
  const long long base = blockIdx.x * (256LL * kU) + threadIdx.x;
  Raw8<kMode> rz[kU], rr[kU], rm[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      const long long px = i >> lg;
      rz[u] = t_load_raw<kMode>(z.p + px * z.cs + z.coff + g * 8, z.plane);
      if (has_res) rr[u] = t_load_raw<kMode>(res.p + px * res.cs + res.coff + g * 8, res.plane);
      if (has_mask) rm[u] = t_load_raw<kMode>(mask.p + px * mask.cs + mask.coff + g * 8, mask.plane);
    }
  }
  float sc[8], sh[8];
  load_f8(scale + g * 8, sc);
  load_f8(shift + g * 8, sh);
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      float v[8], r[8], m[8];
      t_cvt8<kMode>(rz[u], v);
      if (has_res) t_cvt8<kMode>(rr[u], r);
      if (has_mask) t_cvt8<kMode>(rm[u], m);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = fmaf(v[e], sc[e], sh[e]);
        if (has_res) t += r[e];
        if (relu) t = fmaxf(t, 0.f);
        if (has_mask) t *= m[e];
        v[e] = t;
      }
      t_store8<kMode>(y.p + (i >> lg) * y.cs + y.coff + g * 8, y.plane, v);
    }
  }
}

template <int kMode>
__global__ void scale_shift_act_kernel(TView z, TViewW y, TView res, TView mask, const float* __restrict__ scale,
                                       const float* __restrict__ shift, long long npix, int ch, int relu, int has_res,
                                       int has_mask) {
  pdl_enter();
  const int c8 = ch / 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= npix * c8) return;
  const int g = static_cast<int>(i % c8);
  const long long px = i / c8;
  float v[8], r[8], m[8], sc[8], sh[8];
  t_load8<kMode>(z.p + px * z.cs + z.coff + g * 8, z.plane, v);
  if (has_res) t_load8<kMode>(res.p + px * res.cs + res.coff + g * 8, res.plane, r);
  if (has_mask) t_load8<kMode>(mask.p + px * mask.cs + mask.coff + g * 8, mask.plane, m);
  load_f8(scale + g * 8, sc);
  load_f8(shift + g * 8, sh);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float t = fmaf(v[e], sc[e], sh[e]);
    if (has_res) t += r[e];
    if (relu) t = fmaxf(t, 0.f);
    if (has_mask) t *= m[e];
    v[e] = t;
  }
  t_store8<kMode>(y.p + px * y.cs + y.coff + g * 8, y.plane, v);
}

// dz = gamma*invstd * ( dy' - sum_dy/M - xhat * sum_dy_xhat/M ),  dy' = dy * (y > 0 if relu);  optional dres = dy'
// The per-channel part is folded once (bn_bwd_coef_kernel, double arithmetic) into dz = k1*dy' + k2*z + k3.
template <int kMode>
__device__ __forceinline__ void bn_bwd_apply_octet(const Raw8<kMode>& rd, const Raw8<kMode>& rz, const Raw8<kMode>& ry,
                                                   const float (&k1)[8], const float (&k2)[8], const float (&k3)[8],
                                                   int relu, int has_dres, uint16_t* dzp, long long dz_plane,
                                                   uint16_t* drp, long long dr_plane) {
  float vd[8], vz[8], vy[8], o[8];
  t_cvt8<kMode>(rd, vd);
  t_cvt8<kMode>(rz, vz);
  if (relu) {
    t_cvt8<kMode>(ry, vy);
#pragma unroll
    for (int e = 0; e < 8; ++e) vd[e] = vy[e] > 0.f ? vd[e] : 0.f;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = fmaf(k1[e], vd[e], fmaf(k2[e], vz[e], k3[e]));
  t_store8<kMode>(dzp, dz_plane, o);
  if (has_dres) t_store8<kMode>(drp, dr_plane, vd);
}

template <int kMode, int kU>
__global__ void __launch_bounds__(256)
    bn_bwd_apply_fast_kernel(TView dy, TView y, TView z, TViewW dz, TViewW dres, const float* __restrict__ coef,
                             long long total, int lg, int ch, int relu, int has_dres) {
  pdl_enter();
  const int g = threadIdx.x & ((1 << lg) - 1);
  const long long base = blockIdx.x * (256LL * kU) + threadIdx.x;
  Raw8<kMode> rd[kU], rz[kU], ry[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      const long long px = i >> lg;
      rd[u] = t_load_raw<kMode>(dy.p + px * dy.cs + dy.coff + g * 8, dy.plane);
      rz[u] = t_load_raw<kMode>(z.p + px * z.cs + z.coff + g * 8, z.plane);
      if (relu) ry[u] = t_load_raw<kMode>(y.p + px * y.cs + y.coff + g * 8, y.plane);
    }
  }
  float k1[8], k2[8], k3[8];
  load_f8(coef + g * 8, k1);
  load_f8(coef + ch + g * 8, k2);
  load_f8(coef + 2 * ch + g * 8, k3);
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      const long long px = i >> lg;
      bn_bwd_apply_octet<kMode>(rd[u], rz[u], ry[u], k1, k2, k3, relu, has_dres, dz.p + px * dz.cs + dz.coff + g * 8,
                                dz.plane, has_dres ? dres.p + px * dres.cs + dres.coff + g * 8 : nullptr, dres.plane);
    }
  }
}

// out = (accumulate ? out : 0) + a * b_mask   /  generic masked ReLU-gate / plain add:  op 0: out = a (*mask)
template <int kMode>
__global__ void ew_mul_kernel(TView a, TView m, TViewW out, long long npix, int ch, int has_mask, int accumulate,
                              int relu_gate) {
  // relu_gate: `m` is the forward OUTPUT y; pass a where y > 0.  has_mask: multiply by m.
  const int c8 = ch / 8;
  const long long total = npix * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long px = i / c8;
  float v[8];
  t_load8<kMode>(a.p + px * a.cs + a.coff + g * 8, a.plane, v);
  if (has_mask || relu_gate) {
    float mm[8];
    t_load8<kMode>(m.p + px * m.cs + m.coff + g * 8, m.plane, mm);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = relu_gate ? (mm[e] > 0.f ? v[e] : 0.f) : v[e] * mm[e];
  }
  uint16_t* op = out.p + px * out.cs + out.coff + g * 8;
  if (accumulate) {
    float o[8];
    t_load8<kMode>(op, out.plane, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += o[e];
  }
  t_store8<kMode>(op, out.plane, v);
}

// Two-pass max-pool backward (used when the caller provides `idx` scratch, n*ho*wo*c bytes):
//   pass 1: every output window records the position (0..8, row-major) of its FIRST maximum per channel;
//   pass 2: every input pixel sums dy of the (at most 4) windows that point at it.
template <int kMode>
__global__ void maxpool3x3s2_argmax_kernel(TView x, uint8_t* __restrict__ idx, int n, int h, int w, int ho, int wo,
                                           int ch) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long opix = i / c8;
  const int ox = static_cast<int>(opix % wo);
  long long t = opix / wo;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  float best[8];
  uint32_t pos[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    best[e] = -INFINITY;
    pos[e] = 4;   // the centre is always inside the image
  }
#pragma unroll
  for (int d = 0; d < 9; ++d) {
    const int yy = 2 * oy - 1 + d / 3, xx = 2 * ox - 1 + d % 3;
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
    float v[8];
    t_load8<kMode>(x.p + ((static_cast<long long>(b) * h + yy) * w + xx) * x.cs + x.coff + g * 8, x.plane, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (v[e] > best[e]) {   // strict: the first maximum in scan order wins, like ATen
        best[e] = v[e];
        pos[e] = d;
      }
    }
  }
  uint2 o;
  o.x = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24);
  o.y = pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24);
  *reinterpret_cast<uint2*>(idx + opix * ch + g * 8) = o;
}

template <int kMode>
__global__ void maxpool3x3s2_bwd_idx_kernel(const uint8_t* __restrict__ idx, TView dy, TViewW dx, int n, int h, int w,
                                            int ho, int wo, int ch, int accumulate) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long ipix = i / c8;
  const int ix = static_cast<int>(ipix % w);
  long long t = ipix / w;
  const int iy = static_cast<int>(t % h);
  const int b = static_cast<int>(t / h);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // windows covering (iy, ix): oy = iy/2 (+1 when iy is odd), same for ox
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int oy = iy / 2 + a;
    if ((a == 1 && (iy & 1) == 0) || oy >= ho) continue;
    const int py = iy - (2 * oy - 1);
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const int ox = ix / 2 + bb;
      if ((bb == 1 && (ix & 1) == 0) || ox >= wo) continue;
      const uint32_t code = static_cast<uint32_t>(py * 3 + (ix - (2 * ox - 1)));
      const long long opix = (static_cast<long long>(b) * ho + oy) * wo + ox;
      const uint2 k = __ldg(reinterpret_cast<const uint2*>(idx + opix * ch + g * 8));
      float d[8];
      t_load8<kMode>(dy.p + opix * dy.cs + dy.coff + g * 8, dy.plane, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t ke = ((e < 4 ? k.x : k.y) >> (8 * (e & 3))) & 0xFFu;
        acc[e] += ke == code ? d[e] : 0.f;
      }
    }
  }
  uint16_t* op = dx.p + ipix * dx.cs + dx.coff + g * 8;
  if (accumulate) {
    float o[8];
    t_load8<kMode>(op, dx.plane, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += o[e];
  }
  t_store8<kMode>(op, dx.plane, acc);
}

// max-pool 3/2/1 backward (gather form): an input pixel receives dy of every window whose FIRST maximum
// (row-major scan, like ATen) it is.
template <int kMode>
__global__ void maxpool3x3s2_bwd_kernel(TView x, TView dy, TViewW dx, int n, int h, int w, int ho, int wo, int ch,
                                        int accumulate) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long ipix = i / c8;
  const int ix = static_cast<int>(ipix % w);
  long long t = ipix / w;
  const int iy = static_cast<int>(t % h);
  const int b = static_cast<int>(t / h);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  float self[8];
  t_load8<kMode>(x.p + ipix * x.cs + x.coff + g * 8, x.plane, self);
  for (int oy = (iy + 1) / 2 - ((iy + 1) % 2 == 0 ? 1 : 0); oy <= (iy + 1) / 2; ++oy) {
    if (oy < 0 || oy >= ho) continue;
    for (int ox = (ix + 1) / 2 - ((ix + 1) % 2 == 0 ? 1 : 0); ox <= (ix + 1) / 2; ++ox) {
      if (ox < 0 || ox >= wo) continue;
      // is (iy, ix) the first maximum of window (oy, ox)?
      bool first[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) first[e] = true;
      for (int dyy = 0; dyy < 3; ++dyy) {
        const int yy = 2 * oy - 1 + dyy;
        if (yy < 0 || yy >= h) continue;
        for (int dxx = 0; dxx < 3; ++dxx) {
          const int xx = 2 * ox - 1 + dxx;
          if (xx < 0 || xx >= w) continue;
          if (yy == iy && xx == ix) continue;
          float v[8];
          t_load8<kMode>(x.p + ((static_cast<long long>(b) * h + yy) * w + xx) * x.cs + x.coff + g * 8, x.plane, v);
          const bool before = (yy < iy) || (yy == iy && xx < ix);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (before ? (v[e] >= self[e]) : (v[e] > self[e])) first[e] = false;
          }
        }
      }
      float d[8];
      t_load8<kMode>(dy.p + ((static_cast<long long>(b) * ho + oy) * wo + ox) * dy.cs + dy.coff + g * 8, dy.plane, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += first[e] ? d[e] : 0.f;
    }
  }
  uint16_t* op = dx.p + ipix * dx.cs + dx.coff + g * 8;
  if (accumulate) {
    float o[8];
    t_load8<kMode>(op, dx.plane, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += o[e];
  }
  t_store8<kMode>(op, dx.plane, acc);
}

// adjoint of the align_corners bilinear up-sampling (gather form, separable weights)
template <int kMode>
__global__ void upsample_bilinear_ac_bwd_kernel(TView dy, TViewW dx, int n, int h, int w, int ho, int wo, int ch,
                                                float sh, float sw, int accumulate) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long ipix = i / c8;
  const int ix = static_cast<int>(ipix % w);
  long long t = ipix / w;
  const int iy = static_cast<int>(t % h);
  const int b = static_cast<int>(t / h);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const int oy_lo = sh > 0.f ? max(0, static_cast<int>(floorf((iy - 1) / sh)) - 1) : 0;
  const int oy_hi = sh > 0.f ? min(ho - 1, static_cast<int>(ceilf((iy + 1) / sh)) + 1) : ho - 1;
  const int ox_lo = sw > 0.f ? max(0, static_cast<int>(floorf((ix - 1) / sw)) - 1) : 0;
  const int ox_hi = sw > 0.f ? min(wo - 1, static_cast<int>(ceilf((ix + 1) / sw)) + 1) : wo - 1;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const float fy = sh * oy;
    const int y0 = static_cast<int>(fy);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
    const float ly1 = fy - y0, ly0 = 1.f - ly1;
    const float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
    if (wy == 0.f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const float fx = sw * ox;
      const int x0 = static_cast<int>(fx);
      const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float lx1 = fx - x0, lx0 = 1.f - lx1;
      const float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
      if (wx == 0.f) continue;
      float d[8];
      t_load8<kMode>(dy.p + ((static_cast<long long>(b) * ho + oy) * wo + ox) * dy.cs + dy.coff + g * 8, dy.plane, d);
      const float ww = wy * wx;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(ww, d[e], acc[e]);
    }
  }
  uint16_t* op = dx.p + ipix * dx.cs + dx.coff + g * 8;
  if (accumulate) {
    float o[8];
    t_load8<kMode>(op, dx.plane, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += o[e];
  }
  t_store8<kMode>(op, dx.plane, acc);
}

// dx[n,h,w,:] (+)= g[n,:] * mult   (adjoint of the global average pool: mult = 1/(h*w))
template <int kMode>
__global__ void add_broadcast_kernel(TView gsrc, TViewW dx, int n, int hw, int ch, float mult, int accumulate) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * hw * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long px = i / c8;
  const int b = static_cast<int>(px / hw);
  float v[8];
  t_load8<kMode>(gsrc.p + static_cast<long long>(b) * gsrc.cs + gsrc.coff + g * 8, gsrc.plane, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] *= mult;
  uint16_t* op = dx.p + px * dx.cs + dx.coff + g * 8;
  if (accumulate) {
    float o[8];
    t_load8<kMode>(op, dx.plane, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += o[e];
  }
  t_store8<kMode>(op, dx.plane, v);
}

// zero insertion: y[n, 2i, 2j, :] = x[n, i, j, :], everything else 0 (turns a stride-2 dgrad into a stride-1 conv)
template <int kMode>
__global__ void zero_insert2x_kernel(TView x, TViewW y, int n, int h, int w, int ch) {
  const int c8 = ch / 8;
  const long long total = static_cast<long long>(n) * (2 * h) * (2 * w) * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long opix = i / c8;
  const int ox = static_cast<int>(opix % (2 * w));
  long long t = opix / (2 * w);
  const int oy = static_cast<int>(t % (2 * h));
  const int b = static_cast<int>(t / (2 * h));
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if ((ox & 1) == 0 && (oy & 1) == 0) {
    t_load8<kMode>(x.p + ((static_cast<long long>(b) * h + oy / 2) * w + ox / 2) * x.cs + x.coff + g * 8, x.plane, v);
  }
  t_store8<kMode>(y.p + opix * y.cs + y.coff + g * 8, y.plane, v);
}

}  // namespace up

#define UP_T_DISPATCH(dtype, ...)                                      \
  do {                                                                 \
    if ((dtype) == UP_FP16) {                                          \
      constexpr int kMode = 0;                                         \
      __VA_ARGS__;                                                     \
    } else if ((dtype) == UP_BF16) {                                   \
      constexpr int kMode = 1;                                         \
      __VA_ARGS__;                                                     \
    } else if ((dtype) == UP_SPLIT) {                                  \
      constexpr int kMode = 2;                                         \
      __VA_ARGS__;                                                     \
    } else {                                                           \
      return ::up::fail(UP_ERR_INVALID, "bad dtype %d", (int)(dtype)); \
    }                                                                  \
  } while (0)

static inline up::TView tv(const UpView* v) {
  up::TView t{};
  if (v) {
    t.p = static_cast<const uint16_t*>(v->ptr);
    t.cs = v->cstride;
    t.coff = v->coff;
    t.plane = v->plane_stride;
  }
  return t;
}
static inline up::TViewW tvw(const UpView* v) {
  up::TViewW t{};
  if (v) {
    t.p = static_cast<uint16_t*>(v->ptr);
    t.cs = v->cstride;
    t.coff = v->coff;
    t.plane = v->plane_stride;
  }
  return t;
}
static int check_view(const char* who, const UpView* v, int c) {
  UP_CHECK_ARG(v && v->ptr, "%s: null view", who);
  UP_CHECK_ARG(c % 8 == 0 && v->cstride % 8 == 0 && v->coff % 8 == 0 && v->coff + c <= v->cstride,
               "%s: bad channel view (c %d cstride %d coff %d)", who, c, v->cstride, v->coff);
  UP_CHECK_ARG((reinterpret_cast<uintptr_t>(v->ptr) & 15) == 0, "%s: view not 16-byte aligned", who);
  return 0;
}
static inline int blocks_for(long long total) { return static_cast<int>((total + 255) / 256); }
// Per-channel reductions: blocks of 512 threads, ~16 pixel-octets per thread (4 iterations of 4 loads in flight), at
// most kBnRows = 296 blocks (2 per SM).  Block b writes its partial sums to row b of the work buffer.
constexpr int kBnRows = 296;
#ifndef UP_RED_PER_THREAD
#define UP_RED_PER_THREAD 8
#endif
static inline int reduce_grid(long long npix, int octs) {
  // UP_RED_PER_THREAD octets per thread: 16 left the 24x24 / 48x48 layers (two thirds of the BatchNorms) with 36-72
  // blocks on 148 SMs
  long long g = (npix * octs + 512LL * UP_RED_PER_THREAD - 1) / (512LL * UP_RED_PER_THREAD);
  if (g > kBnRows) g = kBnRows;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
// work buffer layout (doubles): [0, 2c) sums | [2c, 4c) coefficient scratch | [4c, 4c + kBnRows*c) partial rows
extern "C" int64_t up_bn_work_doubles(int c) { return static_cast<int64_t>(4 + kBnRows) * c; }

extern "C" int up_bn_stats(const UpView* x, int64_t npix, int c, int dtype, double* sums, void* stream) {
  int rc = check_view("up_bn_stats", x, c);
  if (rc) return rc;
  UP_CHECK_ARG(sums && npix > 0, "up_bn_stats: bad argument");
  const int octs = c / 8;
  UP_CHECK_ARG(octs <= 256 && (octs & (octs - 1)) == 0, "up_bn_stats: c/8 must be a power of two <= 256 (c = %d)", c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = reduce_grid(npix, octs);
  float* rows = reinterpret_cast<float*>(sums + 4 * c);
  UP_T_DISPATCH(dtype, (up::launch_pdl(up::channel_reduce_kernel<kMode, 0>, grid, 512, 512 * 16 * sizeof(float), st, tv(x),
                                          up::TView{}, up::TView{}, nullptr, nullptr, rows, npix, c, 0)));
  UP_CHECK_LAUNCH("channel_reduce_kernel<stats>");
  up::launch_pdl(up::reduce_rows_kernel, (2 * c + 31) / 32, dim3(32, 16), 0, st, rows, grid, 2 * c, sums);
  UP_CHECK_LAUNCH("reduce_rows_kernel");
  return 0;
}

extern "C" int up_bn_finalize(const double* sums, int64_t count, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float momentum, float eps, float* scale,
                              float* shift, float* save_mean, float* save_invstd, int c_real, int c, void* stream) {
  UP_CHECK_ARG(sums && gamma && beta && scale && shift && save_mean && save_invstd && count > 0 && c >= c_real,
               "up_bn_finalize: bad argument");
  const up::BnFinalizeArgs a{static_cast<double>(count), gamma, beta, running_mean, running_var, momentum, eps,
                             scale, shift, save_mean, save_invstd, c_real, c};
  up::launch_pdl(up::bn_finalize_kernel, (c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), sums, a);
  UP_CHECK_LAUNCH("bn_finalize_kernel");
  return 0;
}

extern "C" int up_bn_stats_finalize(const UpView* x, int64_t npix, int c, int dtype, double* work, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var, float momentum,
                                    float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                                    int c_real, void* stream) {
  int rc = check_view("up_bn_stats_finalize", x, c);
  if (rc) return rc;
  const int octs = c / 8;
  UP_CHECK_ARG(octs <= 256 && (octs & (octs - 1)) == 0, "up_bn_stats_finalize: c/8 must be a power of two <= 256 (c = %d)",
               c);
  UP_CHECK_ARG(work && gamma && beta && scale && shift && save_mean && save_invstd && npix > 0 && c >= c_real,
               "up_bn_stats_finalize: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = reduce_grid(npix, octs);
  float* rows = reinterpret_cast<float*>(work + 4 * c);
  UP_T_DISPATCH(dtype, (up::launch_pdl(up::channel_reduce_kernel<kMode, 0>, grid, 512, 512 * 16 * sizeof(float), st, tv(x),
                                          up::TView{}, up::TView{}, nullptr, nullptr, rows, npix, c, 0)));
  UP_CHECK_LAUNCH("channel_reduce_kernel<stats>");
  const up::BnFinalizeArgs a{static_cast<double>(npix), gamma, beta, running_mean, running_var, momentum, eps,
                             scale, shift, save_mean, save_invstd, c_real, c};
  up::launch_pdl(up::bn_finish_kernel<0, up::BnFinalizeArgs>, (c + 31) / 32, dim3(32, 16), 0, st, rows, grid, work, a);
  UP_CHECK_LAUNCH("bn_finish_kernel<stats>");
  return 0;
}

namespace up {
__global__ void bn_eval_prepare_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ rmean, const float* __restrict__ rvar, float eps,
                                       float* __restrict__ scale, float* __restrict__ shift,
                                       float* __restrict__ save_mean, float* __restrict__ save_invstd, int c_real, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float sc = 0.f, sh = 0.f, m = 0.f, is = 0.f;
  if (i < c_real) {
    is = 1.0f / sqrtf(rvar[i] + eps);     // ATen's inference order: invstd, then gamma * invstd
    m = rmean[i];
    sc = gamma[i] * is;
    sh = beta[i] - m * sc;
  }
  scale[i] = sc;
  shift[i] = sh;
  save_mean[i] = m;
  save_invstd[i] = is;
}
}  // namespace up

extern "C" int up_bn_eval_prepare(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, float* save_mean,
                                  float* save_invstd, int c_real, int c, void* stream) {
  UP_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && save_mean && save_invstd,
               "up_bn_eval_prepare: null argument");
  UP_CHECK_ARG(c_real > 0 && c >= c_real, "up_bn_eval_prepare: bad channel counts");
  up::bn_eval_prepare_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      gamma, beta, running_mean, running_var, eps, scale, shift, save_mean, save_invstd, c_real, c);
  UP_CHECK_LAUNCH("bn_eval_prepare_kernel");
  return 0;
}

extern "C" int up_scale_shift_act(const UpView* z, const UpView* y, const UpView* residual, const UpView* mask,
                                  const float* scale, const float* shift, int64_t npix, int c, int relu, int dtype,
                                  void* stream) {
  int rc = check_view("up_scale_shift_act(z)", z, c);
  if (rc) return rc;
  rc = check_view("up_scale_shift_act(y)", y, c);
  if (rc) return rc;
  if (residual && (rc = check_view("up_scale_shift_act(residual)", residual, c))) return rc;
  if (mask && (rc = check_view("up_scale_shift_act(mask)", mask, c))) return rc;
  UP_CHECK_ARG(scale && shift && npix > 0, "up_scale_shift_act: bad argument");
  const int c8 = c / 8;
  const long long total = npix * c8;
  if (c8 <= 256 && (c8 & (c8 - 1)) == 0) {
    int lg = 0;
    while ((1 << lg) < c8) ++lg;
    UP_T_DISPATCH(dtype, ({
                    constexpr int kU = kMode == 2 ? 2 : 4;
                    up::launch_pdl(up::scale_shift_act_fast_kernel<kMode, kU>,
                                   static_cast<unsigned>((total + 256 * kU - 1) / (256 * kU)), 256, 0,
                                   (cudaStream_t)stream, tv(z), tvw(y), tv(residual), tv(mask), scale, shift, total, lg,
                                   relu, residual != nullptr, mask != nullptr);
                  }));
  } else {
    UP_T_DISPATCH(dtype, up::launch_pdl(up::scale_shift_act_kernel<kMode>, blocks_for(total), 256, 0, (cudaStream_t)stream,
                                            tv(z), tvw(y), tv(residual), tv(mask), scale, shift, npix, c, relu,
                                            residual != nullptr, mask != nullptr));
  }
  UP_CHECK_LAUNCH("scale_shift_act_kernel");
  return 0;
}

extern "C" int up_bn_bwd(const UpView* dy, const UpView* y, const UpView* z, const UpView* dz, const UpView* dres,
                         const float* save_mean, const float* save_invstd, const float* gamma, double* work,
                         int64_t npix, int c_real, int c, int flags, int dtype, float* dgamma, float* dbeta,
                         void* stream) {
  // flags: bit 0 = the forward applied a ReLU (gate dy by y > 0), bit 1 = frozen statistics (eval-mode BatchNorm)
  const int relu = flags & 1, frozen = (flags & 2) ? 1 : 0;
  int rc = check_view("up_bn_bwd(dy)", dy, c);
  if (rc) return rc;
  rc = check_view("up_bn_bwd(z)", z, c);
  if (rc) return rc;
  rc = check_view("up_bn_bwd(dz)", dz, c);
  if (rc) return rc;
  if (relu && (rc = check_view("up_bn_bwd(y)", y, c))) return rc;
  if (dres && (rc = check_view("up_bn_bwd(dres)", dres, c))) return rc;
  const int octs = c / 8;
  UP_CHECK_ARG(octs <= 256 && (octs & (octs - 1)) == 0, "up_bn_bwd: c/8 must be a power of two <= 256");
  UP_CHECK_ARG(save_mean && save_invstd && gamma && work && npix > 0 && c_real <= c, "up_bn_bwd: bad argument");
  UP_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "up_bn_bwd: dgamma and dbeta come together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 1. per-channel sums of (dy', dy' * xhat): one partial row per block
  const int grid = reduce_grid(npix, octs);
  float* rows = reinterpret_cast<float*>(work + 4 * c);
  UP_T_DISPATCH(dtype, (up::launch_pdl(up::channel_reduce_kernel<kMode, 1>, grid, 512, 512 * 16 * sizeof(float), st, tv(dy),
                                          tv(y), tv(z), save_mean, save_invstd, rows, npix, c, relu)));
  UP_CHECK_LAUNCH("channel_reduce_kernel<bn bwd>");
  // 2. rows -> sums -> coefficients (3*c floats right behind the 2*c doubles) + dgamma / dbeta
  float* coef = reinterpret_cast<float*>(work + 2 * c);
  const up::BnBwdCoefArgs a{static_cast<double>(npix), save_mean, save_invstd, gamma, coef, dgamma, dbeta, c_real, c,
                            frozen};
  up::launch_pdl(up::bn_finish_kernel<1, up::BnBwdCoefArgs>, (c + 31) / 32, dim3(32, 16), 0, st, rows, grid, work, a);
  UP_CHECK_LAUNCH("bn_finish_kernel<bn bwd>");
  // 3. dz = k1*dy' + k2*z + k3 (and dres = dy')
  const long long total = npix * octs;
  int lg = 0;
  while ((1 << lg) < octs) ++lg;
  constexpr int kU = 2;
  UP_T_DISPATCH(dtype, up::launch_pdl(up::bn_bwd_apply_fast_kernel<kMode, kU>,
                                       static_cast<unsigned>((total + 256 * kU - 1) / (256 * kU)), 256, 0, st, tv(dy),
                                       tv(y), tv(z), tvw(dz), tvw(dres), coef, total, lg, c, relu, dres != nullptr));
  UP_CHECK_LAUNCH("bn_bwd_apply_fast_kernel");
  return 0;
}

extern "C" int up_ew_mul(const UpView* a, const UpView* m, const UpView* out, int64_t npix, int c, int mode_op,
                         int accumulate, int dtype, void* stream) {
  // mode_op: 0 copy/add, 1 multiply by m, 2 ReLU gate (pass where m > 0)
  int rc = check_view("up_ew_mul(a)", a, c);
  if (rc) return rc;
  rc = check_view("up_ew_mul(out)", out, c);
  if (rc) return rc;
  if (mode_op != 0 && (rc = check_view("up_ew_mul(m)", m, c))) return rc;
  UP_T_DISPATCH(dtype, (up::ew_mul_kernel<kMode><<<blocks_for(npix*(c / 8)), 256, 0, (cudaStream_t)stream>>>(
                           tv(a), tv(m), tvw(out), npix, c, mode_op == 1, accumulate, mode_op == 2)));
  UP_CHECK_LAUNCH("ew_mul_kernel");
  return 0;
}

extern "C" int up_maxpool3x3s2_bwd(const UpView* x, const UpView* dy, const UpView* dx, int n, int h, int w, int c,
                                   int accumulate, int dtype, void* idx_scratch, void* stream) {
  int rc = check_view("up_maxpool3x3s2_bwd(x)", x, c);
  if (rc) return rc;
  rc = check_view("up_maxpool3x3s2_bwd(dy)", dy, c);
  if (rc) return rc;
  rc = check_view("up_maxpool3x3s2_bwd(dx)", dx, c);
  if (rc) return rc;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  if (idx_scratch) {
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(idx_scratch) & 7) == 0, "up_maxpool3x3s2_bwd: idx_scratch alignment");
    uint8_t* idx = static_cast<uint8_t*>(idx_scratch);
    UP_T_DISPATCH(dtype, (up::maxpool3x3s2_argmax_kernel<kMode><<<blocks_for((long long)n * ho * wo * (c / 8)), 256, 0,
                                                                  (cudaStream_t)stream>>>(tv(x), idx, n, h, w, ho, wo,
                                                                                          c)));
    UP_CHECK_LAUNCH("maxpool3x3s2_argmax_kernel");
    UP_T_DISPATCH(dtype, (up::maxpool3x3s2_bwd_idx_kernel<kMode><<<blocks_for((long long)n * h * w * (c / 8)), 256, 0,
                                                                   (cudaStream_t)stream>>>(idx, tv(dy), tvw(dx), n, h,
                                                                                           w, ho, wo, c, accumulate)));
    UP_CHECK_LAUNCH("maxpool3x3s2_bwd_idx_kernel");
    return 0;
  }
  UP_T_DISPATCH(dtype, (up::maxpool3x3s2_bwd_kernel<kMode><<<blocks_for((long long)n * h * w * (c / 8)), 256, 0,
                                                             (cudaStream_t)stream>>>(tv(x), tv(dy), tvw(dx), n, h, w, ho,
                                                                                     wo, c, accumulate)));
  UP_CHECK_LAUNCH("maxpool3x3s2_bwd_kernel");
  return 0;
}

extern "C" int up_upsample_bilinear_ac_bwd(const UpView* dy, const UpView* dx, int n, int h, int w, int ho, int wo,
                                           int c, int accumulate, int dtype, void* stream) {
  int rc = check_view("up_upsample_bilinear_ac_bwd(dy)", dy, c);
  if (rc) return rc;
  rc = check_view("up_upsample_bilinear_ac_bwd(dx)", dx, c);
  if (rc) return rc;
  const float sh = ho > 1 ? static_cast<float>(h - 1) / static_cast<float>(ho - 1) : 0.f;
  const float sw = wo > 1 ? static_cast<float>(w - 1) / static_cast<float>(wo - 1) : 0.f;
  UP_T_DISPATCH(dtype, (up::upsample_bilinear_ac_bwd_kernel<kMode><<<blocks_for((long long)n * h * w * (c / 8)), 256, 0,
                                                                     (cudaStream_t)stream>>>(
                           tv(dy), tvw(dx), n, h, w, ho, wo, c, sh, sw, accumulate)));
  UP_CHECK_LAUNCH("upsample_bilinear_ac_bwd_kernel");
  return 0;
}

extern "C" int up_add_broadcast(const UpView* g, const UpView* dx, int n, int hw, int c, float mult, int accumulate,
                                int dtype, void* stream) {
  int rc = check_view("up_add_broadcast(g)", g, c);
  if (rc) return rc;
  rc = check_view("up_add_broadcast(dx)", dx, c);
  if (rc) return rc;
  UP_T_DISPATCH(dtype, (up::add_broadcast_kernel<kMode><<<blocks_for((long long)n * hw * (c / 8)), 256, 0,
                                                          (cudaStream_t)stream>>>(tv(g), tvw(dx), n, hw, c, mult,
                                                                                  accumulate)));
  UP_CHECK_LAUNCH("add_broadcast_kernel");
  return 0;
}

extern "C" int up_zero_insert2x(const UpView* x, const UpView* y, int n, int h, int w, int c, int dtype, void* stream) {
  int rc = check_view("up_zero_insert2x(x)", x, c);
  if (rc) return rc;
  rc = check_view("up_zero_insert2x(y)", y, c);
  if (rc) return rc;
  UP_T_DISPATCH(dtype, (up::zero_insert2x_kernel<kMode><<<blocks_for((long long)n * 4 * h * w * (c / 8)), 256, 0,
                                                          (cudaStream_t)stream>>>(tv(x), tvw(y), n, h, w, c)));
  UP_CHECK_LAUNCH("zero_insert2x_kernel");
  return 0;
}
