// Bandwidth-bound kernels of the UniPose hot path: layout packing, BatchNorm folding, max-pool,
// bilinear up-sampling (align_corners=True), global average pool, broadcast.
// Activations are NHWC 16-bit (bf16 / fp16 / bf16 hi+lo planes); every kernel moves 8 channels
// (16 bytes) per thread so that global accesses are 128-bit and coalesced along C.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "up_internal.h"

namespace up {

// kMode: 0 = fp16, 1 = bf16, 2 = bf16 hi+lo planes
template <int kMode>
__device__ __forceinline__ void load8(const uint16_t* p, long long plane, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e + 0] = cvt16_to_f32<(kMode == 0 ? 0 : 1)>(static_cast<uint16_t>(w[e] & 0xFFFFu));
    v[2 * e + 1] = cvt16_to_f32<(kMode == 0 ? 0 : 1)>(static_cast<uint16_t>(w[e] >> 16));
  }
  if constexpr (kMode == 2) {
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + plane));
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e + 0] += cvt16_to_f32<1>(static_cast<uint16_t>(lw[e] & 0xFFFFu));
      v[2 * e + 1] += cvt16_to_f32<1>(static_cast<uint16_t>(lw[e] >> 16));
    }
  }
}

template <int kMode>
__device__ __forceinline__ void store8(uint16_t* p, long long plane, const float (&v)[8]) {
  if constexpr (kMode == 2) {
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
    uint4 h, l;
    h.x = hi[0] | (uint32_t(hi[1]) << 16);
    h.y = hi[2] | (uint32_t(hi[3]) << 16);
    h.z = hi[4] | (uint32_t(hi[5]) << 16);
    h.w = hi[6] | (uint32_t(hi[7]) << 16);
    l.x = lo[0] | (uint32_t(lo[1]) << 16);
    l.y = lo[2] | (uint32_t(lo[3]) << 16);
    l.z = lo[4] | (uint32_t(lo[5]) << 16);
    l.w = lo[6] | (uint32_t(lo[7]) << 16);
    *reinterpret_cast<uint4*>(p) = h;
    *reinterpret_cast<uint4*>(p + plane) = l;
  } else {
    uint4 h;
    h.x = cvt_f32_to16<kMode>(v[0]) | (uint32_t(cvt_f32_to16<kMode>(v[1])) << 16);
    h.y = cvt_f32_to16<kMode>(v[2]) | (uint32_t(cvt_f32_to16<kMode>(v[3])) << 16);
    h.z = cvt_f32_to16<kMode>(v[4]) | (uint32_t(cvt_f32_to16<kMode>(v[5])) << 16);
    h.w = cvt_f32_to16<kMode>(v[6]) | (uint32_t(cvt_f32_to16<kMode>(v[7])) << 16);
    *reinterpret_cast<uint4*>(p) = h;
  }
}

#define UP_DISPATCH_MODE(dtype, ...)                                   \
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

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  return static_cast<int>(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [plane][tap][cout][cin] 16-bit
// ------------------------------------------------------------------------------------------
template <int kMode>
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int cout_real,
                                   int cin_real, int kh, int kw, int cout, int cin, long long plane) {
  const long long total = static_cast<long long>(kh) * kw * cout * cin;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ci = static_cast<int>(i % cin);
  long long t = i / cin;
  const int co = static_cast<int>(t % cout);
  const int tap = static_cast<int>(t / cout);
  float v = 0.f;
  if (co < cout_real && ci < cin_real) {
    v = w[(static_cast<long long>(co) * cin_real + ci) * kh * kw + tap];
  }
  if constexpr (kMode == 2) {
    uint16_t hi, lo;
    split_bf16(v, hi, lo);
    out[i] = hi;
    out[i + plane] = lo;
  } else {
    out[i] = cvt_f32_to16<kMode>(v);
  }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int c_real, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float s = 0.f, b = 0.f;
  if (i < c_real) {
    // same operation order as ATen's batch_norm inference path: invstd = 1/sqrt(var+eps)
    const float invstd = 1.0f / sqrtf(var[i] + eps);
    s = gamma[i] * invstd;
    b = beta[i] - mean[i] * s;
  }
  scale[i] = s;
  shift[i] = b;
}

// ------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------
template <int kMode>
__global__ void pack_input_s2d_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int n, int h, int w,
                                      long long plane, int wpitch, int wpad) {
  const int hq = h / 2, wq = w / 2;
  const long long total = static_cast<long long>(n) * hq * wq;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int xq = static_cast<int>(i % wq);
  long long t = i / wq;
  const int yq = static_cast<int>(t % hq);
  const int b = static_cast<int>(t / hq);
  float v0[8], v1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v0[e] = 0.f;
    v1[e] = 0.f;
  }
  // channel order (ph, pw, c): index = (ph*2 + pw)*3 + c
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(
          x + ((static_cast<long long>(b) * 3 + c) * h + (2 * yq + ph)) * w + 2 * xq));
      const int i0 = (ph * 2 + 0) * 3 + c;
      const int i1 = (ph * 2 + 1) * 3 + c;
      if (i0 < 8) v0[i0] = f.x; else v1[i0 - 8] = f.x;
      if (i1 < 8) v0[i1] = f.y; else v1[i1 - 8] = f.y;
    }
  }
  uint16_t* o = y + ((static_cast<long long>(b) * hq + yq) * wpitch + xq + wpad) * 16;
  store8<kMode>(o, plane, v0);
  store8<kMode>(o + 8, plane, v1);
}

template <int kMode>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int n, int c_real, int h,
                                    int w, int c, int cs, int coff, long long plane) {
  const int c8 = c / 8;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long pix = i / c8;
  const int xw = static_cast<int>(pix % w);
  long long t = pix / w;
  const int yh = static_cast<int>(t % h);
  const int b = static_cast<int>(t / h);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = g * 8 + e;
    v[e] = ch < c_real ? x[((static_cast<long long>(b) * c_real + ch) * h + yh) * w + xw] : 0.f;
  }
  store8<kMode>(y + pix * cs + coff + g * 8, plane, v);
}

template <int kMode>
__global__ void nhwc_to_nchw_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, int n, int c_real, int h,
                                    int w, int cs, int coff, long long plane) {
  const int c8 = (c_real + 7) / 8;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  // w fastest so that the fp32 NCHW writes of a warp are contiguous
  const int xw = static_cast<int>(i % w);
  long long t = i / w;
  const int yh = static_cast<int>(t % h);
  t /= h;
  const int g = static_cast<int>(t % c8);
  const int b = static_cast<int>(t / c8);
  const long long pix = (static_cast<long long>(b) * h + yh) * w + xw;
  float v[8];
  load8<kMode>(x + pix * cs + coff + g * 8, plane, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = g * 8 + e;
    if (ch < c_real) y[((static_cast<long long>(b) * c_real + ch) * h + yh) * w + xw] = v[e];
  }
}

// ------------------------------------------------------------------------------------------
// pooling / resampling
// ------------------------------------------------------------------------------------------
template <int kMode>
__global__ void maxpool3x3s2_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int n, int h, int w,
                                    int ho, int wo, int c, int xcs, int xcoff, int ycs, int ycoff, long long xplane,
                                    long long yplane) {
  const int c8 = c / 8;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long opix = i / c8;
  const int ox = static_cast<int>(opix % wo);
  long long t = opix / wo;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = 2 * oy - 1 + dy;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = 2 * ox - 1 + dx;
      if (ix < 0 || ix >= w) continue;
      float v[8];
      load8<kMode>(x + ((static_cast<long long>(b) * h + iy) * w + ix) * xcs + xcoff + g * 8, xplane, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], v[e]);
    }
  }
  store8<kMode>(y + opix * ycs + ycoff + g * 8, yplane, m);
}

// 16-bit fast path (fp16 / bf16 storage): the maximum is exact in the storage format, so it is taken on packed pairs
// (HMNMX2) without widening; one thread produces TWO horizontally adjacent outputs from a 3 x 5 window (15 loads of
// 16 bytes instead of 18).  kFmt: 0 = fp16, 1 = bf16.
template <int kFmt>
__device__ __forceinline__ uint32_t max2_16(uint32_t a, uint32_t b) {
  if constexpr (kFmt == 0) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  } else {
    const __nv_bfloat162 r =
        __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
}
template <int kFmt>
__device__ __forceinline__ uint4 max8_16(const uint4& a, const uint4& b) {
  return make_uint4(max2_16<kFmt>(a.x, b.x), max2_16<kFmt>(a.y, b.y), max2_16<kFmt>(a.z, b.z), max2_16<kFmt>(a.w, b.w));
}

template <int kFmt>
__global__ void maxpool3x3s2_pair_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int n, int h, int w,
                                         int ho, int wo, int c, int xcs, int xcoff, int ycs, int ycoff) {
  const int c8 = c / 8;
  const int wo2 = wo / 2;
  const long long total = static_cast<long long>(n) * ho * wo2 * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  long long t = i / c8;
  const int op = static_cast<int>(t % wo2);
  t /= wo2;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  const int ox = 2 * op;
  const int ix0 = 2 * ox - 1;   // window columns ix0 .. ix0+4: outputs ox (cols 0..2) and ox+1 (cols 2..4)
  // Out-of-image taps are clamped onto the nearest valid row / column: that element already belongs to the same
  // window and max() is idempotent, so the 15 loads are unconditional and all in flight together.
  uint4 v[3][5];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = min(max(2 * oy - 1 + dy, 0), h - 1);
    const uint16_t* row = x + (static_cast<long long>(b) * h + iy) * w * xcs + xcoff + g * 8;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) {
      const int ix = min(max(ix0 + dx, 0), w - 1);
      v[dy][dx] = __ldg(reinterpret_cast<const uint4*>(row + static_cast<long long>(ix) * xcs));
    }
  }
  uint4 c2 = max8_16<kFmt>(max8_16<kFmt>(v[0][2], v[1][2]), v[2][2]);
  uint4 m0 = c2, m1 = c2;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    m0 = max8_16<kFmt>(m0, max8_16<kFmt>(v[dy][0], v[dy][1]));
    m1 = max8_16<kFmt>(m1, max8_16<kFmt>(v[dy][3], v[dy][4]));
  }
  uint16_t* o = y + ((static_cast<long long>(b) * ho + oy) * wo + ox) * ycs + ycoff + g * 8;
  *reinterpret_cast<uint4*>(o) = m0;
  *reinterpret_cast<uint4*>(o + ycs) = m1;
}

// ATen upsample_bilinear2d, align_corners=True:  scale = (in-1)/(out-1) (0 when out == 1), src = scale*dst,
// i0 = int(src), i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1.
template <int kMode>
__global__ void upsample_bilinear_ac_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int n, int h,
                                            int w, int ho, int wo, int c, int xcs, int xcoff, int ycs, int ycoff,
                                            long long xplane, long long yplane, float sh, float sw) {
  const int c8 = c / 8;
  const long long total = static_cast<long long>(n) * ho * wo * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long opix = i / c8;
  const int ox = static_cast<int>(opix % wo);
  long long t = opix / wo;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  const float fy = sh * oy;
  const float fx = sw * ox;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - y0, lx1 = fx - x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const uint16_t* base = x + static_cast<long long>(b) * h * w * xcs + xcoff + g * 8;
  float v00[8], v01[8], v10[8], v11[8], o[8];
  load8<kMode>(base + (static_cast<long long>(y0) * w + x0) * xcs, xplane, v00);
  load8<kMode>(base + (static_cast<long long>(y0) * w + x1) * xcs, xplane, v01);
  load8<kMode>(base + (static_cast<long long>(y1) * w + x0) * xcs, xplane, v10);
  load8<kMode>(base + (static_cast<long long>(y1) * w + x1) * xcs, xplane, v11);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] = ly0 * (lx0 * v00[e] + lx1 * v01[e]) + ly1 * (lx0 * v10[e] + lx1 * v11[e]);
  }
  store8<kMode>(y + opix * ycs + ycoff + g * 8, yplane, o);
}

__global__ void upsample_bilinear_ac_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int nc, int h,
                                                 int w, int ho, int wo, float sh, float sw) {
  const long long total = static_cast<long long>(nc) * ho * wo;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ox = static_cast<int>(i % wo);
  long long t = i / wo;
  const int oy = static_cast<int>(t % ho);
  const long long pc = t / ho;
  const float fy = sh * oy, fx = sw * ox;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* p = x + pc * h * w;
  y[i] = ly0 * (lx0 * p[y0 * w + x0] + lx1 * p[y0 * w + x1]) + ly1 * (lx0 * p[y1 * w + x0] + lx1 * p[y1 * w + x1]);
}

// one block per (image, 64 channels): 8 channel-octets x 32 pixel lanes
template <int kMode>
__global__ void global_avgpool_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int hw, int c, int xcs,
                                      int xcoff, int ycs, int ycoff, long long xplane, long long yplane, int sum_only) {
  const int blocks_per_img = c / 64;
  const int b = blockIdx.x / blocks_per_img;
  const int cg = blockIdx.x % blocks_per_img;
  const int oct = threadIdx.x & 7;
  const int pl = threadIdx.x >> 3;  // 0..31
  const int ch = cg * 64 + oct * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const uint16_t* base = x + static_cast<long long>(b) * hw * xcs + xcoff + ch;
  for (int p = pl; p < hw; p += 32) {
    float v[8];
    load8<kMode>(base + static_cast<long long>(p) * xcs, xplane, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
  __shared__ float red[32][65];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[pl][oct * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int q = 0; q < 32; ++q) s += red[q][threadIdx.x];
    red[0][threadIdx.x] = sum_only ? s : s / static_cast<float>(hw);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = red[0][threadIdx.x * 8 + e];
    store8<kMode>(y + static_cast<long long>(b) * ycs + ycoff + cg * 64 + threadIdx.x * 8, yplane, o);
  }
}

template <int kMode>
__global__ void broadcast_hw_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int n, int hw, int c,
                                    int xcs, int xcoff, int ycs, int ycoff, long long xplane, long long yplane) {
  const int c8 = c / 8;
  const long long total = static_cast<long long>(n) * hw * c8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int g = static_cast<int>(i % c8);
  const long long opix = i / c8;
  const int b = static_cast<int>(opix / hw);
  const uint16_t* src = x + static_cast<long long>(b) * xcs + xcoff + g * 8;
  uint16_t* dst = y + opix * ycs + ycoff + g * 8;
  *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(src));
  if constexpr (kMode == 2) {
    *reinterpret_cast<uint4*>(dst + yplane) = __ldg(reinterpret_cast<const uint4*>(src + xplane));
  }
}

}  // namespace up

using namespace up;

#define UP_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int up_pack_conv_weight(const float* w_oihw, void* w_packed, int cout_real, int cin_real, int kh, int kw,
                                   int cout, int cin, int dtype, int64_t w_plane_stride, void* stream) {
  UP_CHECK_ARG(w_oihw && w_packed, "up_pack_conv_weight: null argument");
  UP_CHECK_ARG(cout_real > 0 && cin_real > 0 && cout >= cout_real && cin >= cin_real && kh > 0 && kw > 0,
               "up_pack_conv_weight: bad dims");
  const long long total = static_cast<long long>(kh) * kw * cout * cin;
  UP_DISPATCH_MODE(dtype, (pack_weight_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              w_oihw, static_cast<uint16_t*>(w_packed), cout_real, cin_real, kh, kw, cout, cin,
                              w_plane_stride)));
  UP_CHECK_LAUNCH("pack_weight_kernel");
  return 0;
}

extern "C" int up_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                          float* scale, float* shift, int c_real, int c, void* stream) {
  UP_CHECK_ARG(gamma && beta && mean && var && scale && shift, "up_bn_fold: null argument");
  UP_CHECK_ARG(c_real > 0 && c >= c_real, "up_bn_fold: bad channel counts");
  bn_fold_kernel<<<grid_for(c, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, scale, shift, c_real,
                                                                     c);
  UP_CHECK_LAUNCH("bn_fold_kernel");
  return 0;
}

extern "C" int up_pack_input_s2d(const float* x_nchw, void* y, int n, int h, int w, int dtype, int64_t y_plane_stride,
                                 int y_wpitch, int y_wpad_left, void* stream) {
  UP_CHECK_ARG(x_nchw && y, "up_pack_input_s2d: null argument");
  UP_CHECK_ARG(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "up_pack_input_s2d: h, w must be even");
  UP_CHECK_ARG((reinterpret_cast<uintptr_t>(x_nchw) & 7) == 0 && UP_ALIGNED16(y), "up_pack_input_s2d: alignment");
  if (y_wpitch <= 0) y_wpitch = w / 2;
  UP_CHECK_ARG(y_wpad_left >= 0 && y_wpad_left + w / 2 <= y_wpitch, "up_pack_input_s2d: bad row pitch / padding");
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2);
  UP_DISPATCH_MODE(dtype, (pack_input_s2d_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              x_nchw, static_cast<uint16_t*>(y), n, h, w, y_plane_stride, y_wpitch, y_wpad_left)));
  UP_CHECK_LAUNCH("pack_input_s2d_kernel");
  return 0;
}

extern "C" int up_nchw_f32_to_nhwc(const float* x, void* y, int n, int c_real, int h, int w, int c, int y_cstride,
                                   int y_coff, int dtype, int64_t y_plane_stride, void* stream) {
  UP_CHECK_ARG(x && y, "up_nchw_f32_to_nhwc: null argument");
  UP_CHECK_ARG(c % 8 == 0 && c >= c_real && y_cstride % 8 == 0 && y_coff % 8 == 0 && y_coff + c <= y_cstride,
               "up_nchw_f32_to_nhwc: bad channel view");
  UP_CHECK_ARG(UP_ALIGNED16(y), "up_nchw_f32_to_nhwc: y alignment");
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  UP_DISPATCH_MODE(dtype, (nchw_to_nhwc_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              x, static_cast<uint16_t*>(y), n, c_real, h, w, c, y_cstride, y_coff, y_plane_stride)));
  UP_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int up_nhwc_to_nchw_f32(const void* x, float* y, int n, int c_real, int h, int w, int x_cstride,
                                   int x_coff, int dtype, int64_t x_plane_stride, void* stream) {
  UP_CHECK_ARG(x && y, "up_nhwc_to_nchw_f32: null argument");
  UP_CHECK_ARG(x_cstride % 8 == 0 && x_coff % 8 == 0 && x_coff + ((c_real + 7) / 8) * 8 <= x_cstride,
               "up_nhwc_to_nchw_f32: bad channel view");
  UP_CHECK_ARG(UP_ALIGNED16(x), "up_nhwc_to_nchw_f32: x alignment");
  const long long total = static_cast<long long>(n) * h * w * ((c_real + 7) / 8);
  UP_DISPATCH_MODE(dtype, (nhwc_to_nchw_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), y, n, c_real, h, w, x_cstride, x_coff, x_plane_stride)));
  UP_CHECK_LAUNCH("nhwc_to_nchw_kernel");
  return 0;
}

static int check_views(const char* who, const void* x, const void* y, int c, int xcs, int xcoff, int ycs, int ycoff) {
  UP_CHECK_ARG(x && y, "%s: null argument", who);
  UP_CHECK_ARG(c > 0 && c % 8 == 0, "%s: c (%d) must be a multiple of 8", who, c);
  UP_CHECK_ARG(xcs % 8 == 0 && xcoff % 8 == 0 && xcoff + c <= xcs, "%s: bad x channel view", who);
  UP_CHECK_ARG(ycs % 8 == 0 && ycoff % 8 == 0 && ycoff + c <= ycs, "%s: bad y channel view", who);
  UP_CHECK_ARG(UP_ALIGNED16(x) && UP_ALIGNED16(y), "%s: pointers must be 16-byte aligned", who);
  return 0;
}

extern "C" int up_maxpool3x3s2(const void* x, void* y, int n, int h, int w, int c, int x_cstride, int x_coff,
                               int y_cstride, int y_coff, int dtype, int64_t x_plane_stride, int64_t y_plane_stride,
                               void* stream) {
  int rc = check_views("up_maxpool3x3s2", x, y, c, x_cstride, x_coff, y_cstride, y_coff);
  if (rc) return rc;
  const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  if (dtype != UP_SPLIT && wo % 2 == 0) {
    const long long total2 = static_cast<long long>(n) * ho * (wo / 2) * (c / 8);
    if (dtype == UP_FP16) {
      maxpool3x3s2_pair_kernel<0><<<grid_for(total2, 256), 256, 0, (cudaStream_t)stream>>>(
          static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, h, w, ho, wo, c, x_cstride, x_coff, y_cstride,
          y_coff);
    } else {
      maxpool3x3s2_pair_kernel<1><<<grid_for(total2, 256), 256, 0, (cudaStream_t)stream>>>(
          static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, h, w, ho, wo, c, x_cstride, x_coff, y_cstride,
          y_coff);
    }
    UP_CHECK_LAUNCH("maxpool3x3s2_pair_kernel");
    return 0;
  }
  UP_DISPATCH_MODE(dtype, (maxpool3x3s2_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, h, w, ho, wo, c, x_cstride,
                              x_coff, y_cstride, y_coff, x_plane_stride, y_plane_stride)));
  UP_CHECK_LAUNCH("maxpool3x3s2_kernel");
  return 0;
}

static inline float ac_scale(int in, int out) { return out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f; }

extern "C" int up_upsample_bilinear_ac(const void* x, void* y, int n, int h, int w, int ho, int wo, int c,
                                       int x_cstride, int x_coff, int y_cstride, int y_coff, int dtype,
                                       int64_t x_plane_stride, int64_t y_plane_stride, void* stream) {
  int rc = check_views("up_upsample_bilinear_ac", x, y, c, x_cstride, x_coff, y_cstride, y_coff);
  if (rc) return rc;
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  UP_DISPATCH_MODE(dtype, (upsample_bilinear_ac_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, h, w, ho, wo, c, x_cstride,
                              x_coff, y_cstride, y_coff, x_plane_stride, y_plane_stride, ac_scale(h, ho),
                              ac_scale(w, wo))));
  UP_CHECK_LAUNCH("upsample_bilinear_ac_kernel");
  return 0;
}

extern "C" int up_upsample_bilinear_ac_nchw_f32(const float* x, float* y, int n, int c, int h, int w, int ho, int wo,
                                                void* stream) {
  UP_CHECK_ARG(x && y && n > 0 && c > 0, "up_upsample_bilinear_ac_nchw_f32: bad argument");
  const long long total = static_cast<long long>(n) * c * ho * wo;
  upsample_bilinear_ac_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n * c, h, w, ho, wo,
                                                                                           ac_scale(h, ho),
                                                                                           ac_scale(w, wo));
  UP_CHECK_LAUNCH("upsample_bilinear_ac_nchw_kernel");
  return 0;
}

extern "C" int up_global_avgpool(const void* x, void* y, int n, int h, int w, int c, int x_cstride, int x_coff,
                                 int y_cstride, int y_coff, int dtype, int64_t x_plane_stride, int64_t y_plane_stride,
                                 void* stream) {
  int rc = check_views("up_global_avgpool", x, y, c, x_cstride, x_coff, y_cstride, y_coff);
  if (rc) return rc;
  UP_CHECK_ARG(c % 64 == 0, "up_global_avgpool: c must be a multiple of 64");
  UP_DISPATCH_MODE(dtype, (global_avgpool_kernel<kMode><<<n*(c / 64), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), h * w, c, x_cstride, x_coff,
                              y_cstride, y_coff, x_plane_stride, y_plane_stride, 0)));
  UP_CHECK_LAUNCH("global_avgpool_kernel");
  return 0;
}

extern "C" int up_global_sumpool(const void* x, void* y, int n, int h, int w, int c, int x_cstride, int x_coff,
                                 int y_cstride, int y_coff, int dtype, int64_t x_plane_stride, int64_t y_plane_stride,
                                 void* stream) {
  int rc = check_views("up_global_sumpool", x, y, c, x_cstride, x_coff, y_cstride, y_coff);
  if (rc) return rc;
  UP_CHECK_ARG(c % 64 == 0, "up_global_sumpool: c must be a multiple of 64");
  UP_DISPATCH_MODE(dtype, (global_avgpool_kernel<kMode><<<n*(c / 64), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), h * w, c, x_cstride, x_coff,
                              y_cstride, y_coff, x_plane_stride, y_plane_stride, 1)));
  UP_CHECK_LAUNCH("global_sumpool_kernel");
  return 0;
}

extern "C" int up_broadcast_hw(const void* x, void* y, int n, int ho, int wo, int c, int x_cstride, int x_coff,
                               int y_cstride, int y_coff, int dtype, int64_t x_plane_stride, int64_t y_plane_stride,
                               void* stream) {
  int rc = check_views("up_broadcast_hw", x, y, c, x_cstride, x_coff, y_cstride, y_coff);
  if (rc) return rc;
  const long long total = static_cast<long long>(n) * ho * wo * (c / 8);
  UP_DISPATCH_MODE(dtype, (broadcast_hw_kernel<kMode><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
                              static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, ho * wo, c, x_cstride,
                              x_coff, y_cstride, y_coff, x_plane_stride, y_plane_stride)));
  UP_CHECK_LAUNCH("broadcast_hw_kernel");
  return 0;
}
