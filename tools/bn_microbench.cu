// Stand-alone microbenchmark for the streaming BatchNorm kernels of the training step (tools only, not shipped):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_bn_microbench tools/bn_microbench.cu
// Variants of "y = relu(z*scale + shift + res)" (bf16, dense NHWC) and of the per-channel (sum, sum^2) reduction, timed
// as CUDA graphs of 32 launches over rotating buffers (so that small layers are not served from L2 by accident).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void ld8(const uint16_t* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void st8(uint16_t* p, const float (&v)[8]) {
  uint32_t h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    h[e] = *reinterpret_cast<uint32_t*>(&b);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
}

// E0: one octet per thread, 64-bit div/mod, per-use constant loads (the shipped kernel)
__global__ void e0(const uint16_t* z, uint16_t* y, const uint16_t* res, const float* scale, const float* shift,
                   long long npix, int ch, int has_res) {
  const int c8 = ch / 8;
  const long long total = npix * c8;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % c8);
  const long long px = i / c8;
  float v[8], r[8];
  ld8(z + px * ch + g * 8, v);
  if (has_res) ld8(res + px * ch + g * 8, r);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float t = fmaf(v[e], scale[g * 8 + e], shift[g * 8 + e]);
    if (has_res) t += r[e];
    v[e] = fmaxf(t, 0.f);
  }
  st8(y + px * ch + g * 8, v);
}

// E1<U>: a block owns 256*U consecutive octets; thread t takes t, t+256, ...; its channel octet is fixed (c/8 | 256)
template <int kU>
__global__ void __launch_bounds__(256) e1(const uint16_t* z, uint16_t* y, const uint16_t* res, const float* scale,
                                          const float* shift, long long total, int lg, int ch, int has_res) {
  const int g = threadIdx.x & ((1 << lg) - 1);
  const long long base = blockIdx.x * (256LL * kU) + threadIdx.x;
  float v[kU][8], r[kU][8];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      ld8(z + i * 8, v[u]);
      if (has_res) ld8(res + i * 8, r[u]);
    }
  }
  float sc[8], sh[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(scale + g * 8) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(shift + g * 8) + q);
    sc[4 * q] = a.x, sc[4 * q + 1] = a.y, sc[4 * q + 2] = a.z, sc[4 * q + 3] = a.w;
    sh[4 * q] = b.x, sh[4 * q + 1] = b.y, sh[4 * q + 2] = b.z, sh[4 * q + 3] = b.w;
  }
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = fmaf(v[u][e], sc[e], sh[e]);
        if (has_res) t += r[u][e];
        v[u][e] = fmaxf(t, 0.f);
      }
      st8(y + i * 8, v[u]);
    }
  }
}

// E2<U>: persistent, grid = SMs * 8, the block walks chunks of 256*U octets
template <int kU>
__global__ void __launch_bounds__(256) e2(const uint16_t* z, uint16_t* y, const uint16_t* res, const float* scale,
                                          const float* shift, long long total, int lg, int ch, int has_res) {
  const int g = threadIdx.x & ((1 << lg) - 1);
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sc[e] = scale[g * 8 + e], sh[e] = shift[g * 8 + e];
  for (long long c0 = blockIdx.x * (256LL * kU); c0 < total; c0 += gridDim.x * (256LL * kU)) {
    float v[kU][8], r[kU][8];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = c0 + u * 256 + threadIdx.x;
      if (i < total) {
        ld8(z + i * 8, v[u]);
        if (has_res) ld8(res + i * 8, r[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = c0 + u * 256 + threadIdx.x;
      if (i < total) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = fmaf(v[u][e], sc[e], sh[e]);
          if (has_res) t += r[u][e];
          v[u][e] = fmaxf(t, 0.f);
        }
        st8(y + i * 8, v[u]);
      }
    }
  }
}

// copy: the ceiling for a 2-tensor pass
__global__ void cp(const uint4* a, uint4* b, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) b[i] = __ldg(a + i);
}

// R<kThreads, kU>: per-channel (sum, sum^2); thread owns octet tid % octs, strides pixels; rows out
template <int kThreads, int kU>
__global__ void __launch_bounds__(kThreads) red(const uint16_t* x, float* rows, long long npix, int ch) {
  const int octs = ch / 8;
  const int oct = threadIdx.x % octs, pstride = kThreads / octs, lane = threadIdx.x / octs;
  float s0[8], s1[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(p0 + per, npix);
  for (long long q = p0 + lane; q < p1; q += (long long)kU * pstride) {
    float v[kU][8];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long px = q + (long long)u * pstride;
      if (px < p1) ld8(x + px * ch + oct * 8, v[u]);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[u][e] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) s0[e] += v[u][e], s1[e] = fmaf(v[u][e], v[u][e], s1[e]);
  }
  extern __shared__ float sm[];
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[threadIdx.x * 16 + e] = s0[e], sm[threadIdx.x * 16 + 8 + e] = s1[e];
  __syncthreads();
  for (int idx = threadIdx.x; idx < octs * 16; idx += kThreads) {
    const int o = idx / 16, e = idx % 16;
    float s = 0.f;
    for (int k = 0; k < pstride; ++k) s += sm[(k * octs + o) * 16 + e];
    rows[(long long)blockIdx.x * 2 * ch + (e < 8 ? 0 : ch) + o * 8 + (e & 7)] = s;
  }
}

// A0: shipped backward apply (float4 constant loads per octet);  A1<U>: block-contiguous, constants hoisted
__global__ void a0(const uint16_t* dy, const uint16_t* y, const uint16_t* z, uint16_t* dz, uint16_t* dres,
                   const float* coef, long long npix, int ch, int has_dres) {
  const int c8 = ch / 8;
  const long long total = npix * c8;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = (int)(i % c8);
  const long long px = i / c8;
  float vd[8], vz[8], vy[8], o[8];
  ld8(dy + px * ch + g * 8, vd);
  ld8(z + px * ch + g * 8, vz);
  ld8(y + px * ch + g * 8, vy);
#pragma unroll
  for (int e = 0; e < 8; ++e) vd[e] = vy[e] > 0.f ? vd[e] : 0.f;
  const float4* k1 = reinterpret_cast<const float4*>(coef + g * 8);
  const float4* k2 = reinterpret_cast<const float4*>(coef + ch + g * 8);
  const float4* k3 = reinterpret_cast<const float4*>(coef + 2 * ch + g * 8);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = __ldg(k1 + q), b = __ldg(k2 + q), c = __ldg(k3 + q);
    o[4 * q + 0] = fmaf(a.x, vd[4 * q + 0], fmaf(b.x, vz[4 * q + 0], c.x));
    o[4 * q + 1] = fmaf(a.y, vd[4 * q + 1], fmaf(b.y, vz[4 * q + 1], c.y));
    o[4 * q + 2] = fmaf(a.z, vd[4 * q + 2], fmaf(b.z, vz[4 * q + 2], c.z));
    o[4 * q + 3] = fmaf(a.w, vd[4 * q + 3], fmaf(b.w, vz[4 * q + 3], c.w));
  }
  st8(dz + px * ch + g * 8, o);
  if (has_dres) st8(dres + px * ch + g * 8, vd);
}

__device__ __forceinline__ void cv8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = __uint_as_float(w[e] << 16);
    v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
  }
}

template <int kU>
__global__ void __launch_bounds__(256) a1(const uint16_t* dy, const uint16_t* y, const uint16_t* z, uint16_t* dz,
                                          uint16_t* dres, const float* coef, long long total, int lg, int ch,
                                          int has_dres) {
  const int g = threadIdx.x & ((1 << lg) - 1);
  const long long base = blockIdx.x * (256LL * kU) + threadIdx.x;
  uint4 rd[kU], rz[kU], ry[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      rd[u] = __ldg(reinterpret_cast<const uint4*>(dy + i * 8));
      rz[u] = __ldg(reinterpret_cast<const uint4*>(z + i * 8));
      ry[u] = __ldg(reinterpret_cast<const uint4*>(y + i * 8));
    }
  }
  float k1[8], k2[8], k3[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(coef + g * 8) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(coef + ch + g * 8) + q);
    const float4 c = __ldg(reinterpret_cast<const float4*>(coef + 2 * ch + g * 8) + q);
    k1[4 * q] = a.x, k1[4 * q + 1] = a.y, k1[4 * q + 2] = a.z, k1[4 * q + 3] = a.w;
    k2[4 * q] = b.x, k2[4 * q + 1] = b.y, k2[4 * q + 2] = b.z, k2[4 * q + 3] = b.w;
    k3[4 * q] = c.x, k3[4 * q + 1] = c.y, k3[4 * q + 2] = c.z, k3[4 * q + 3] = c.w;
  }
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      float vd[8], vz[8], vy[8], o[8];
      cv8(rd[u], vd);
      cv8(rz[u], vz);
      cv8(ry[u], vy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        vd[e] = vy[e] > 0.f ? vd[e] : 0.f;
        o[e] = fmaf(k1[e], vd[e], fmaf(k2[e], vz[e], k3[e]));
      }
      st8(dz + i * 8, o);
      if (has_dres) st8(dres + i * 8, vd);
    }
  }
}

// E3<U>: like E1 but the loaded octets stay packed until they are used (register budget)
template <int kU>
__global__ void __launch_bounds__(256) e3(const uint16_t* z, uint16_t* y, const uint16_t* res, const float* scale,
                                          const float* shift, long long total, int lg, int ch, int has_res) {
  const int g = threadIdx.x & ((1 << lg) - 1);
  const long long base = blockIdx.x * (256LL * kU) + threadIdx.x;
  uint4 rv[kU], rr[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      rv[u] = __ldg(reinterpret_cast<const uint4*>(z + i * 8));
      if (has_res) rr[u] = __ldg(reinterpret_cast<const uint4*>(res + i * 8));
    }
  }
  float sc[8], sh[8];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(scale + g * 8) + q);
    const float4 b = __ldg(reinterpret_cast<const float4*>(shift + g * 8) + q);
    sc[4 * q] = a.x, sc[4 * q + 1] = a.y, sc[4 * q + 2] = a.z, sc[4 * q + 3] = a.w;
    sh[4 * q] = b.x, sh[4 * q + 1] = b.y, sh[4 * q + 2] = b.z, sh[4 * q + 3] = b.w;
  }
#pragma unroll
  for (int u = 0; u < kU; ++u) {
    const long long i = base + u * 256;
    if (i < total) {
      float v[8], r[8];
      cv8(rv[u], v);
      if (has_res) cv8(rr[u], r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = fmaf(v[e], sc[e], sh[e]);
        if (has_res) t += r[e];
        v[e] = fmaxf(t, 0.f);
      }
      st8(y + i * 8, v);
    }
  }
}

// RB<kThreads, kU>: backward reduction (sum dy', sum dy' * xhat), three tensors, loads kept packed
template <int kThreads, int kU>
__global__ void __launch_bounds__(kThreads) redb(const uint16_t* dy, const uint16_t* y, const uint16_t* z,
                                                 const float* mean, const float* invstd, float* rows, long long npix,
                                                 int ch) {
  const int octs = ch / 8;
  const int oct = threadIdx.x % octs, pstride = kThreads / octs, lane = threadIdx.x / octs;
  float s0[8], s1[8], mu[8], is[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s0[e] = s1[e] = 0.f, mu[e] = mean[oct * 8 + e], is[e] = invstd[oct * 8 + e];
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = blockIdx.x * per, p1 = min(p0 + per, npix);
  for (long long q = p0 + lane; q < p1; q += (long long)kU * pstride) {
    uint4 rd[kU], rz[kU], ry[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long px = q + (long long)u * pstride;
      if (px < p1) {
        rd[u] = __ldg(reinterpret_cast<const uint4*>(dy + px * ch + oct * 8));
        rz[u] = __ldg(reinterpret_cast<const uint4*>(z + px * ch + oct * 8));
        ry[u] = __ldg(reinterpret_cast<const uint4*>(y + px * ch + oct * 8));
      } else {
        rd[u] = make_uint4(0, 0, 0, 0);
        rz[u] = rd[u];
        ry[u] = rd[u];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float vd[8], vz[8], vy[8];
      cv8(rd[u], vd);
      cv8(rz[u], vz);
      cv8(ry[u], vy);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = vy[e] > 0.f ? vd[e] : 0.f;
        s0[e] += d;
        s1[e] = fmaf(d, (vz[e] - mu[e]) * is[e], s1[e]);
      }
    }
  }
  extern __shared__ float sm[];
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[threadIdx.x * 16 + e] = s0[e], sm[threadIdx.x * 16 + 8 + e] = s1[e];
  __syncthreads();
  for (int idx = threadIdx.x; idx < octs * 16; idx += kThreads) {
    const int o = idx / 16, e = idx % 16;
    float s = 0.f;
    for (int k = 0; k < pstride; ++k) s += sm[(k * octs + o) * 16 + e];
    rows[(long long)blockIdx.x * 2 * ch + (e < 8 ? 0 : ch) + o * 8 + (e & 7)] = s;
  }
}

struct Shape { long long npix; int c; };

template <class F>
static float time_graph(F&& launch, int nbuf) {
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaGraph_t g;
  cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  for (int k = 0; k < 32; ++k) launch(st, k % nbuf);
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int w = 0; w < 2; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(a, st));
  for (int w = 0; w < 4; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGraphExecDestroy(ge));
  CK(cudaGraphDestroy(g));
  CK(cudaStreamDestroy(st));
  return ms * 1e3f / (32 * 4);   // us per launch
}

int main() {
  const Shape shapes[] = {{18432, 256}, {18432, 1024}, {73728, 512}, {294912, 64}, {294912, 256}, {18432, 2048}};
  const long long max_elems = 294912LL * 256;
  const int kBuf = 4;   // 4 x 3 tensors x 151 MB at the largest shape
  uint16_t *z[kBuf], *y[kBuf], *r[kBuf];
  for (int k = 0; k < kBuf; ++k) {
    CK(cudaMalloc(&z[k], max_elems * 2));
    CK(cudaMalloc(&y[k], max_elems * 2));
    CK(cudaMalloc(&r[k], max_elems * 2));
    CK(cudaMemset(z[k], 0x3c, max_elems * 2));
    CK(cudaMemset(r[k], 0x3c, max_elems * 2));
  }
  float *scale, *shift, *rows, *coef;
  CK(cudaMalloc(&coef, 3 * 2048 * 4));
  CK(cudaMemset(coef, 0, 3 * 2048 * 4));
  CK(cudaMalloc(&scale, 2048 * 4));
  CK(cudaMalloc(&shift, 2048 * 4));
  CK(cudaMalloc(&rows, 2048LL * 2 * 2048 * 4));
  CK(cudaMemset(scale, 0, 2048 * 4));
  CK(cudaMemset(shift, 0, 2048 * 4));
  CK(cudaFuncSetAttribute(red<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 64));
  for (const Shape& s : shapes) {
    const int c8 = s.c / 8;
    int lg = 0;
    while ((1 << lg) < c8) ++lg;
    const long long total = s.npix * c8;
    const double mb = s.npix * s.c * 2 / 1e6;
    // small layers: the working set of a real step is one layer (L2-warm from the producing conv), so report both
    // the rotating (cold-ish) and the single-buffer (warm) figure
    for (int nbuf : {kBuf, 1}) {
      printf("\nshape npix %lld c %d (%.1f MB per tensor), %s buffers\n", s.npix, s.c, mb, nbuf == 1 ? "one set of" : "rotating");
      auto rep = [&](const char* name, float us, int tensors) {
        printf("  %-22s %7.2f us  %6.0f GB/s\n", name, us, mb * tensors / us * 1e3);
      };
      for (int has_res : {0, 1}) {
        const int t = 2 + has_res;
        char nm[64];
        snprintf(nm, 64, "e0 res%d", has_res);
        rep(nm, time_graph([&](cudaStream_t st, int k) { e0<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z[k], y[k], r[k], scale, shift, s.npix, s.c, has_res); }, nbuf), t);
#define E1(U) snprintf(nm, 64, "e1<%d> res%d", U, has_res); \
        rep(nm, time_graph([&](cudaStream_t st, int k) { e1<U><<<(unsigned)((total + 256 * U - 1) / (256 * U)), 256, 0, st>>>(z[k], y[k], r[k], scale, shift, total, lg, s.c, has_res); }, nbuf), t);
        E1(1) E1(2) E1(4) E1(8)
#define E2(U) snprintf(nm, 64, "e2<%d> res%d", U, has_res); \
        rep(nm, time_graph([&](cudaStream_t st, int k) { e2<U><<<148 * 8, 256, 0, st>>>(z[k], y[k], r[k], scale, shift, total, lg, s.c, has_res); }, nbuf), t);
        E2(2) E2(4)
      }

      {
        char nm[64];
#define E3(U) for (int has_res : {0, 1}) { snprintf(nm, 64, "e3<%d> res%d", U, has_res); \
        rep(nm, time_graph([&](cudaStream_t st, int k) { e3<U><<<(unsigned)((total + 256 * U - 1) / (256 * U)), 256, 0, st>>>(z[k], y[k], r[k], scale, shift, total, lg, s.c, has_res); }, nbuf), 2 + has_res); }
        E3(2) E3(4) E3(8)
        for (int has_dres : {0, 1}) {
          snprintf(nm, 64, "a0 dres%d", has_dres);
          rep(nm, time_graph([&](cudaStream_t st, int k) { a0<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z[k], r[k], z[(k + 1) % kBuf], y[k], y[(k + 1) % kBuf], coef, s.npix, s.c, has_dres); }, nbuf), 4 + has_dres);
#define A1(U) snprintf(nm, 64, "a1<%d> dres%d", U, has_dres); \
          rep(nm, time_graph([&](cudaStream_t st, int k) { a1<U><<<(unsigned)((total + 256 * U - 1) / (256 * U)), 256, 0, st>>>(z[k], r[k], z[(k + 1) % kBuf], y[k], y[(k + 1) % kBuf], coef, total, lg, s.c, has_dres); }, nbuf), 4 + has_dres);
          A1(1) A1(2) A1(4)
        }
        for (int per : {8, 16, 32}) {
          long long g512 = (total + 512LL * per - 1) / (512LL * per);
          if (g512 > 296) g512 = 296;
          long long g256 = (total + 256LL * per - 1) / (256LL * per);
          if (g256 > 592) g256 = 592;
#define RB(T, U, G) snprintf(nm, 64, "redb<%d,%d> per%d g%lld", T, U, per, G); \
          rep(nm, time_graph([&](cudaStream_t st, int k) { redb<T, U><<<(unsigned)G, T, T * 64, st>>>(z[k], r[k], z[(k + 1) % kBuf], scale, shift, rows, s.npix, s.c); }, nbuf), 3);
          RB(512, 1, g512) RB(512, 2, g512) RB(512, 4, g512) RB(256, 2, g256) RB(256, 4, g256)
        }
      }
      rep("copy", time_graph([&](cudaStream_t st, int k) { cp<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)z[k], (uint4*)y[k], total); }, nbuf), 2);
      for (int per : {2, 4, 8, 16}) {
        char nm[64];
        long long g512 = (total + 512LL * per - 1) / (512LL * per);
        if (g512 > 2048) g512 = 2048;
        snprintf(nm, 64, "red<512,4> per%d g%lld", per, g512);
        rep(nm, time_graph([&](cudaStream_t st, int k) { red<512, 4><<<(unsigned)g512, 512, 512 * 64, st>>>(z[k], rows, s.npix, s.c); }, nbuf), 1);
        long long g256 = (total + 256LL * per - 1) / (256LL * per);
        if (g256 > 2048) g256 = 2048;
        snprintf(nm, 64, "red<256,4> per%d g%lld", per, g256);
        rep(nm, time_graph([&](cudaStream_t st, int k) { red<256, 4><<<(unsigned)g256, 256, 256 * 64, st>>>(z[k], rows, s.npix, s.c); }, nbuf), 1);
        snprintf(nm, 64, "red<256,8> per%d g%lld", per, g256);
        rep(nm, time_graph([&](cudaStream_t st, int k) { red<256, 8><<<(unsigned)g256, 256, 256 * 64, st>>>(z[k], rows, s.npix, s.c); }, nbuf), 1);
      }
    }
  }
  return 0;
}
