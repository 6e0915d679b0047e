// Implicit-GEMM convolution for sm_100a: TMA (im2col-free, one box per filter tap with the tap's
// dilated offset, OOB zero fill = padding) -> shared memory (128B / 32B swizzle) -> tcgen05.mma with
// fp32 accumulators in TMEM -> fused scale/shift(/residual)(/ReLU) epilogue -> TMA store (NHWC)
// or direct fp32 NCHW store.
//
// GEMM view:  M = output pixels (tile = bn x bh x bw = 128 pixels of one/several images),
//             N = output channels (tile = block_n in {32,64,128,256}),
//             K = taps x input channels (k-block = one tap x `ck` channels, ck in {16,64}).
//
// Persistent, warp-specialised CTA (256 threads, 1 CTA / SM):
//   warp 0 lane 0 : TMA producer            (smem ring: full/empty mbarriers)
//   warp 1 lane 0 : tcgen05.mma issuer      (TMEM double buffer: tmem_full/tmem_empty mbarriers)
//   warp 2        : TMEM allocator / deallocator
//   warps 4..7    : epilogue (TMEM -> regs -> smem staging -> TMA store), overlaps the next tile's MMAs
//
// Filter taps whose whole input box lies outside the image contribute exact zeros and are skipped by
// producer and issuer alike (large-dilation WASP convs on small maps: wasp.py:47-49).
//
// UP_SPLIT ("fp32-grade") mode: activations and weights are bf16 hi+lo planes; every k-block is issued
// three times (hi*hi, lo*hi, hi*lo) into the same fp32 accumulator.
#include <cuda.h>

#include "up_internal.h"
#include "up_ptx.cuh"

namespace up {

constexpr int kMaxStages = 8;
constexpr int kTileM = 128;
constexpr int kStagingBytes = 2 * 16384;  // two 128-row x 128-byte output staging buffers
constexpr int kEpiThreads = 128;
constexpr int kEpiWarp0 = 4;

struct ConvKParams {
  int N, Hq, Wq;  // input extent in box coordinates (H/stride, W/stride)
  int Ho, Wo;
  int taps_h, taps_w, dil, pad_h, pad_w, stride;
  int ck, chunks, chunks_per_group, group_nstride;
  int x_coff, x_cs;
  int bn, bh, bw;
  int tiles_w, tiles_h, tiles_n;
  int n_tiles, block_n;
  int cout;
  int nterms;
  int stages;
  uint32_t a_bytes, b_bytes;
  uint32_t idesc;
  uint32_t tmem_cols;
  int flags, fmt, split;
  int cout_valid, out_c_total;
  int y_coff;
  int r_cs, r_coff;
  long long r_plane;
  const float* scale;
  const float* shift;
  const uint16_t* res;
  float* out_f32;
  float* stats;
};

struct TileCoord {
  int n0, h0, w0, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  t.nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  int th = mt % p.tiles_h;
  int tn = mt / p.tiles_h;
  t.n0 = tn * p.bn;
  t.h0 = th * p.bh;
  t.w0 = tw * p.bw;
  return t;
}

// Box origin (ch, cw) and parity plane (ph, pw) of filter tap (kh, kw) for the tile at (h0, w0);
// returns false when the box cannot touch the image.
__device__ __forceinline__ bool tap_box(const ConvKParams& p, int h0, int w0, int kh, int kw, int& ch, int& cw,
                                        int& ph, int& pw) {
  int oh = kh * p.dil - p.pad_h;
  int ow = kw * p.dil - p.pad_w;
  ph = 0;
  pw = 0;
  if (p.stride == 2) {
    ph = oh & 1;
    pw = ow & 1;
    oh = (oh - ph) >> 1;
    ow = (ow - pw) >> 1;
  }
  ch = h0 + oh;
  cw = w0 + ow;
  return (ch + p.bh > 0) && (ch < p.Hq) && (cw + p.bw > 0) && (cw < p.Wq);
}

__device__ __forceinline__ float load16(const uint16_t* p, int fmt) { return cvt16_to_f32_rt(*p, fmt); }

__global__ void __launch_bounds__(256, 1)
    conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY0,
                        const __grid_constant__ CUtensorMap tmY1, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms of TMA and the UMMA descriptors.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const uint32_t staging = smem_base + p.stages * stage_bytes;
  const uint32_t bars = staging + kStagingBytes;
  // barrier layout (8 bytes each): full[kMaxStages] empty[kMaxStages] tmem_full[2] tmem_empty[2] ; then tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4);
  // generic pointer to the tmem slot for reading it back
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (threadIdx.x == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t phase = 0;
    const int taps = p.taps_h * p.taps_w;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      for (int kh = 0; kh < p.taps_h; ++kh) {
        for (int kw = 0; kw < p.taps_w; ++kw) {
          int ch, cw, ph, pw;
          if (!tap_box(p, t.h0, t.w0, kh, kw, ch, cw, ph, pw)) continue;
          const int tap = kh * p.taps_w + kw;
          for (int chunk = 0; chunk < p.chunks; ++chunk) {
            const int g = chunk / p.chunks_per_group;
            const int cc = chunk - g * p.chunks_per_group;
            const int c = p.x_coff + cc * p.ck + pw * p.x_cs;
            const int n = t.n0 + g * p.group_nstride;
            for (int term = 0; term < p.nterms; ++term) {
              mbar_wait(empty_bar(s), phase ^ 1u);
              const uint32_t a_dst = smem_base + s * stage_bytes;
              const uint32_t b_dst = a_dst + p.a_bytes;
              mbar_arrive_expect_tx(full_bar(s), stage_bytes);
              tma_load_5d(term == 1 ? &tmA1 : &tmA0, a_dst, full_bar(s), c, cw, ph, ch, n);
              const int brow = ((term == 2 ? taps : 0) + tap) * p.cout + t.nt * p.block_n;
              tma_load_2d(&tmB, b_dst, full_bar(s), chunk * p.ck, brow);
              if (++s == p.stages) {
                s = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (threadIdx.x == 32) {
    // ===================== MMA issuer =====================
    int s = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t sw_bytes = p.ck * 2;
    const int kk = p.ck / 16;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tcgen05_after_thread_sync();
      const uint32_t tmem_d = tmem_base + acc * p.block_n;
      uint32_t accumulate = 0;
      for (int kh = 0; kh < p.taps_h; ++kh) {
        for (int kw = 0; kw < p.taps_w; ++kw) {
          int ch, cw, ph, pw;
          if (!tap_box(p, t.h0, t.w0, kh, kw, ch, cw, ph, pw)) continue;
          const int nkb = p.chunks * p.nterms;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(full_bar(s), phase);
            tcgen05_after_thread_sync();
            const uint32_t a_addr = smem_base + s * stage_bytes;
            const uint64_t adesc = make_smem_desc_kmajor(a_addr, sw_bytes);
            const uint64_t bdesc = make_smem_desc_kmajor(a_addr + p.a_bytes, sw_bytes);
            for (int k = 0; k < kk; ++k) {
              // advance 16 elements (32 bytes) along K inside the swizzle row: +2 in 16-byte units
              umma_f16(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(empty_bar(s));  // frees the smem slot once these MMAs have read it
            if (++s == p.stages) {
              s = 0;
              phase ^= 1u;
            }
          }
        }
      }
      umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue =====================
    const int ew = warp - kEpiWarp0;  // == warp % 4 -> TMEM lane quarter
    const int row = ew * 32 + lane;
    const bool issuer = (threadIdx.x == kEpiWarp0 * 32);
    const int fmt = p.fmt;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t store_seq = 0;
    const int bhw = p.bh * p.bw;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int rn = row / bhw;
      const int rem = row - rn * bhw;
      const int rh = rem / p.bw;
      const int rw = rem - rh * p.bw;
      const int n = t.n0 + rn, h = t.h0 + rh, w = t.w0 + rw;
      const bool valid = (n < p.N) && (h < p.Ho) && (w < p.Wo);
      const long long pix = (static_cast<long long>(n) * p.Ho + h) * p.Wo + w;

      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_after_thread_sync();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * p.block_n;

      if (p.flags & UP_FLAG_OUT_NCHW_F32) {
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(taddr0 + c0, r);
          tmem_ld_wait();
          const int colbase = t.nt * p.block_n + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = colbase + j;
            float v = __uint_as_float(r[j]) * __ldg(p.scale + col) + __ldg(p.shift + col);
            if (p.flags & UP_FLAG_RELU) v = fmaxf(v, 0.f);
            if (valid && col < p.cout_valid) {
              p.out_f32[((static_cast<long long>(n) * p.out_c_total + col) * p.Ho + h) * p.Wo + w] = v;
            }
          }
        }
      } else {
        const int groups = p.block_n / 64;
        for (int g = 0; g < groups; ++g) {
          // staging buffer(s) for this group
          uint32_t buf0, buf1;
          if (p.split) {
            if (issuer) tma_store_wait_read<0>();
            buf0 = staging;
            buf1 = staging + 16384;
          } else {
            if (issuer) tma_store_wait_read<1>();
            buf0 = staging + (store_seq & 1u) * 16384;
            buf1 = buf0;
          }
          named_bar_sync(1, kEpiThreads);
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr0 + g * 64 + half * 32, r);
            tmem_ld_wait();
            const int colbase = t.nt * p.block_n + g * 64 + half * 32;
            float v[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + colbase) + j4);
              const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + colbase) + j4);
              v[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), sc.x, sh.x);
              v[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), sc.y, sh.y);
              v[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), sc.z, sh.z);
              v[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), sc.w, sh.w);
            }
            if ((p.flags & UP_FLAG_RESIDUAL) && valid) {
              const uint16_t* rp = p.res + pix * p.r_cs + p.r_coff + colbase;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + q);
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[8 * q + 2 * e + 0] += cvt16_to_f32_rt(static_cast<uint16_t>(uw[e] & 0xFFFFu), fmt);
                  v[8 * q + 2 * e + 1] += cvt16_to_f32_rt(static_cast<uint16_t>(uw[e] >> 16), fmt);
                }
              }
              if (p.split) {
                const uint16_t* rl = rp + p.r_plane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(rl) + q);
                  const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    v[8 * q + 2 * e + 0] += cvt16_to_f32<1>(static_cast<uint16_t>(uw[e] & 0xFFFFu));
                    v[8 * q + 2 * e + 1] += cvt16_to_f32<1>(static_cast<uint16_t>(uw[e] >> 16));
                  }
                }
              }
            }
            if (p.flags & UP_FLAG_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            // pack to 16-bit and write this thread's 64 bytes (4 x 16B chunks) into the swizzled staging row
            const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t chunk = static_cast<uint32_t>(half * 4 + q) ^ (static_cast<uint32_t>(row) & 7u);
              uint32_t w0, w1, w2, w3;
              if (p.split) {
                uint16_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16(v[8 * q + e], hi[e], lo[e]);
                w0 = hi[0] | (static_cast<uint32_t>(hi[1]) << 16);
                w1 = hi[2] | (static_cast<uint32_t>(hi[3]) << 16);
                w2 = hi[4] | (static_cast<uint32_t>(hi[5]) << 16);
                w3 = hi[6] | (static_cast<uint32_t>(hi[7]) << 16);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf0 + rowoff + (chunk << 4)),
                             "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
                w0 = lo[0] | (static_cast<uint32_t>(lo[1]) << 16);
                w1 = lo[2] | (static_cast<uint32_t>(lo[3]) << 16);
                w2 = lo[4] | (static_cast<uint32_t>(lo[5]) << 16);
                w3 = lo[6] | (static_cast<uint32_t>(lo[7]) << 16);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf1 + rowoff + (chunk << 4)),
                             "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
              } else {
                w0 = pack2_rt(v[8 * q + 0], v[8 * q + 1], fmt);
                w1 = pack2_rt(v[8 * q + 2], v[8 * q + 3], fmt);
                w2 = pack2_rt(v[8 * q + 4], v[8 * q + 5], fmt);
                w3 = pack2_rt(v[8 * q + 6], v[8 * q + 7], fmt);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf0 + rowoff + (chunk << 4)),
                             "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (issuer) {
            const int c = p.y_coff + t.nt * p.block_n + g * 64;
            tma_store_5d(&tmY0, buf0, c, t.w0, 0, t.h0, t.n0);
            if (p.split) tma_store_5d(&tmY1, buf1, c, t.w0, 0, t.h0, t.n0);
            tma_store_commit();
          }
          ++store_seq;
        }
      }
      // all TMEM reads of this accumulator are done -> hand it back to the MMA issuer
      tcgen05_before_thread_sync();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tcgen05_before_thread_sync();
  __syncthreads();
  if (warp == 2) {
    tcgen05_after_thread_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
}  // namespace up
#include "up_conv_host.h"  // tensor-map encoding + tile picking helpers (shared with the wgrad kernel)
namespace up {

static int g_sm_count = 0;
static size_t g_max_smem = 0;
static bool g_attr_set = false;

static int ensure_device() {
  if (g_sm_count == 0) {
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    cudaDeviceProp prop;
    rc = check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    if (rc) return rc;
    if (prop.major != 10) {
      return fail(UP_ERR_UNSUPPORTED, "unipose_b200 needs an sm_100 class GPU (found sm_%d%d)", prop.major, prop.minor);
    }
    g_sm_count = prop.multiProcessorCount;
    g_max_smem = prop.sharedMemPerBlockOptin;
  }
  if (!g_attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(g_max_smem)),
                        "cudaFuncSetAttribute(max dynamic smem)");
    if (rc) return rc;
    g_attr_set = true;
  }
  return 0;
}

}  // namespace up

using namespace up;

extern "C" int up_conv2d_fwd(const UpConvDesc* d, const void* x, const void* w_packed, const float* scale,
                             const float* shift, const void* residual, void* y, float* stats, void* stream) {
  UP_CHECK_ARG(d && x && w_packed && scale && shift && y, "up_conv2d_fwd: null argument");
  UP_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->ho > 0 && d->wo > 0, "up_conv2d_fwd: bad spatial dims");
  UP_CHECK_ARG(d->stride == 1 || d->stride == 2, "up_conv2d_fwd: stride must be 1 or 2 (got %d)", d->stride);
  UP_CHECK_ARG(d->stride == 1 || (d->h % 2 == 0 && d->w % 2 == 0), "up_conv2d_fwd: stride 2 needs even h, w");
  UP_CHECK_ARG(d->kh >= 1 && d->kw >= 1 && d->dil >= 1 && d->pad_h >= 0 && d->pad_w >= 0, "up_conv2d_fwd: bad filter");
  UP_CHECK_ARG(d->cin > 0 && d->cin % 16 == 0, "up_conv2d_fwd: cin (%d) must be a positive multiple of 16", d->cin);
  UP_CHECK_ARG(d->cout > 0 && d->cout % 32 == 0, "up_conv2d_fwd: cout (%d) must be a positive multiple of 32", d->cout);
  UP_CHECK_ARG(d->dtype == UP_BF16 || d->dtype == UP_FP16 || d->dtype == UP_SPLIT, "up_conv2d_fwd: bad dtype");
  UP_CHECK_ARG(!(d->flags & UP_FLAG_STATS), "up_conv2d_fwd: UP_FLAG_STATS is handled by up_bn_stats");
  const int groups = d->x_groups > 0 ? d->x_groups : 1;
  UP_CHECK_ARG(d->cin % groups == 0, "up_conv2d_fwd: cin not divisible by x_groups");
  const int cin_g = d->cin / groups;
  const int ck = (cin_g % 64 == 0) ? 64 : 16;
  UP_CHECK_ARG(cin_g % ck == 0, "up_conv2d_fwd: per-group cin (%d) must be a multiple of 16", cin_g);
  UP_CHECK_ARG(d->x_cstride % 8 == 0 && d->x_coff % 8 == 0 && d->x_coff + cin_g <= d->x_cstride,
               "up_conv2d_fwd: bad x channel view (cstride %d coff %d cin/group %d)", d->x_cstride, d->x_coff, cin_g);
  UP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "up_conv2d_fwd: x / w must be 16-byte aligned");
  const bool nchw = (d->flags & UP_FLAG_OUT_NCHW_F32) != 0;
  const bool split = d->dtype == UP_SPLIT;
  if (nchw) {
    UP_CHECK_ARG(d->cout_valid > 0 && d->cout_valid <= d->cout, "up_conv2d_fwd: bad cout_valid");
    UP_CHECK_ARG(d->out_c_total == 0 || d->out_c_total >= d->cout_valid, "up_conv2d_fwd: bad out_c_total");
    UP_CHECK_ARG(!(d->flags & UP_FLAG_RESIDUAL), "up_conv2d_fwd: residual not supported with NCHW fp32 output");
  } else {
    UP_CHECK_ARG(d->cout % 64 == 0, "up_conv2d_fwd: NHWC output needs cout %% 64 == 0 (got %d)", d->cout);
    UP_CHECK_ARG(d->y_cstride % 8 == 0 && d->y_coff % 8 == 0 && d->y_coff + d->cout <= d->y_cstride,
                 "up_conv2d_fwd: bad y channel view");
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "up_conv2d_fwd: y must be 16-byte aligned");
    if (split) UP_CHECK_ARG(d->y_plane_stride % 8 == 0 && d->y_plane_stride > 0, "up_conv2d_fwd: bad y_plane_stride");
  }
  if (split) UP_CHECK_ARG(d->x_plane_stride % 8 == 0 && d->x_plane_stride > 0, "up_conv2d_fwd: bad x_plane_stride");
  if (d->flags & UP_FLAG_RESIDUAL) {
    UP_CHECK_ARG(residual != nullptr, "up_conv2d_fwd: residual pointer missing");
    UP_CHECK_ARG(d->r_cstride % 8 == 0 && d->r_coff % 8 == 0 && d->r_coff + d->cout <= d->r_cstride,
                 "up_conv2d_fwd: bad residual channel view");
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0, "up_conv2d_fwd: residual must be 16B aligned");
    if (split) UP_CHECK_ARG(d->r_plane_stride % 8 == 0 && d->r_plane_stride > 0, "up_conv2d_fwd: bad r_plane_stride");
  }
  // every tile must see at least one in-bounds tap: the centre of the receptive field has to hit the image
  UP_CHECK_ARG(d->pad_h <= (d->kh - 1) * d->dil && d->pad_w <= (d->kw - 1) * d->dil,
               "up_conv2d_fwd: padding larger than the filter extent");

  int rc = ensure_device();
  if (rc) return rc;

  const int fmt = fmt_of_dtype(d->dtype);
  ConvKParams p{};
  p.N = d->n;
  p.Hq = d->h / d->stride;
  p.Wq = d->w / d->stride;
  p.Ho = d->ho;
  p.Wo = d->wo;
  p.taps_h = d->kh;
  p.taps_w = d->kw;
  p.dil = d->dil;
  p.pad_h = d->pad_h;
  p.pad_w = d->pad_w;
  p.stride = d->stride;
  p.ck = ck;
  p.chunks = d->cin / ck;
  p.chunks_per_group = cin_g / ck;
  p.group_nstride = groups > 1 ? d->x_group_nstride : 0;
  p.x_coff = d->x_coff;
  p.x_cs = d->x_cstride;
  pick_tile(d->n, d->ho, d->wo, p.bn, p.bh, p.bw);
  p.tiles_w = (d->wo + p.bw - 1) / p.bw;
  p.tiles_h = (d->ho + p.bh - 1) / p.bh;
  p.tiles_n = (d->n + p.bn - 1) / p.bn;
  const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  int block_n = (d->cout % 256 == 0) ? 256 : (d->cout % 128 == 0) ? 128 : (d->cout % 64 == 0) ? 64 : 32;
  while (block_n > 64 && static_cast<long long>(m_tiles) * (d->cout / block_n) < g_sm_count) block_n /= 2;
  p.block_n = block_n;
  p.n_tiles = d->cout / block_n;
  p.cout = d->cout;
  p.nterms = split ? 3 : 1;
  p.a_bytes = kTileM * ck * 2;
  p.b_bytes = block_n * ck * 2;
  const size_t fixed = 1024 + kStagingBytes + 8 * (2 * kMaxStages + 4) + 16;
  int stages = static_cast<int>((g_max_smem - fixed) / (p.a_bytes + p.b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  UP_CHECK_ARG(stages >= 2, "up_conv2d_fwd: not enough shared memory for 2 pipeline stages");
  p.stages = stages;
  p.idesc = make_idesc_f16(static_cast<uint32_t>(fmt), kTileM, static_cast<uint32_t>(block_n));
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * block_n)) cols *= 2;
  p.tmem_cols = cols;
  p.flags = d->flags;
  p.fmt = fmt;
  p.split = split ? 1 : 0;
  p.cout_valid = d->cout_valid;
  p.out_c_total = d->out_c_total > 0 ? d->out_c_total : d->cout_valid;
  p.y_coff = d->y_coff;
  p.r_cs = d->r_cstride;
  p.r_coff = d->r_coff;
  p.r_plane = d->r_plane_stride;
  p.scale = scale;
  p.shift = shift;
  p.res = static_cast<const uint16_t*>(residual);
  p.out_f32 = nchw ? static_cast<float*>(y) : nullptr;
  p.stats = stats;

  // ---- tensor maps ----
  CUtensorMap tmA0, tmA1, tmB, tmY0, tmY1;
  const int sw = ck * 2;
  const uint32_t abox[5] = {static_cast<uint32_t>(ck), static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh),
                            static_cast<uint32_t>(p.bn)};
  const int n_total = d->n + (groups - 1) * p.group_nstride;
  rc = encode_act_map(&tmA0, fmt, x, n_total, d->h, d->w, d->x_cstride, d->stride, abox, sw, "x");
  if (rc) return rc;
  if (split) {
    rc = encode_act_map(&tmA1, fmt, static_cast<const uint16_t*>(x) + d->x_plane_stride, n_total, d->h, d->w,
                        d->x_cstride, d->stride, abox, sw, "x.lo");
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    const int taps = d->kh * d->kw;
    const uint64_t dims[2] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(split ? 2 : 1) * taps * d->cout};
    const uint64_t st[1] = {static_cast<uint64_t>(d->cin) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(ck), static_cast<uint32_t>(block_n)};
    if (split) {
      UP_CHECK_ARG(d->w_plane_stride == static_cast<int64_t>(taps) * d->cout * d->cin,
                   "up_conv2d_fwd: split weights must have contiguous planes (w_plane_stride = taps*cout*cin)");
    }
    rc = encode_map(&tmB, fmt, 2, w_packed, dims, st, box, sw, "w");
    if (rc) return rc;
  }
  if (!nchw) {
    const uint32_t ybox[5] = {64u, static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh),
                              static_cast<uint32_t>(p.bn)};
    rc = encode_act_map(&tmY0, fmt, y, d->n, d->ho, d->wo, d->y_cstride, 1, ybox, 128, "y");
    if (rc) return rc;
    if (split) {
      rc = encode_act_map(&tmY1, fmt, static_cast<uint16_t*>(y) + d->y_plane_stride, d->n, d->ho, d->wo,
                          d->y_cstride, 1, ybox, 128, "y.lo");
      if (rc) return rc;
    } else {
      tmY1 = tmY0;
    }
  } else {
    tmY0 = tmA0;
    tmY1 = tmA0;
  }

  const long long total_tiles = static_cast<long long>(m_tiles) * p.n_tiles;
  const int grid = static_cast<int>(total_tiles < g_sm_count ? total_tiles : g_sm_count);
  const size_t smem = fixed + static_cast<size_t>(stages) * (p.a_bytes + p.b_bytes);
  conv_tcgen05_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(tmA0, tmA1, tmB, tmY0, tmY1, p);
  UP_CHECK_LAUNCH("conv_tcgen05_kernel launch");
  return 0;
}
