// Weight-gradient convolution (wgrad) on tcgen05 tensor cores.
//
//   dW[tap][co][ci] = sum over output pixels p of  dZ[p][co] * X[p (+) tap][ci]
//
// i.e. per filter tap a GEMM with M = cout, N = cin and K = N*Ho*Wo pixels.  Both operands are
// channel-contiguous NHWC activations, so they are *MN-major* UMMA operands: a TMA box
// [64 pixels x 64 channels] lands in shared memory exactly as the 128B-swizzled MN-major canonical
// layout (8-pixel K-atoms of 1024 B; 64-channel MN blocks one box apart).  No transposes, no im2col:
// the X box of a tap is the dZ pixel tile shifted by the tap's dilated offset (same coordinates as
// the forward A-operand, incl. the stride-2 parity folding and whole-box skipping of OOB taps).
//
// Work item = (pixel split, tap, group of `cpi` 128-wide cout blocks, <=256-wide cin block).  cpi = 2 wherever cout has
// an even number of blocks: the X box of a k-step then feeds TWO accumulators (neither operand has any reuse across
// k-steps, so the X stream is half of the L2 -> SM bytes: 64 instead of 96 bytes per clock and SM at full MMA rate).
// The fp32 accumulators [128 x block_n] live in TMEM (double buffered when two sets fit); each item writes its partial to
// scratch[split][tap][cout][cin] with plain vector stores and a second kernel reduces the splits and
// scatters to the OIHW fp32 gradient (no atomics).
//
// Mirrors what autograd's conv backward (cuDNN wgrad) computes for every nn.Conv2d of the hot path
// (reference call sites: unipose.py:123 loss.backward()).
#include <cuda.h>

#include "up_conv_host.h"
#include "up_internal.h"
#include "up_ptx.cuh"

namespace up {

constexpr int kWgMaxStages = 8;
constexpr int kWgPix = 64;  // pixels (GEMM-K) per pipeline stage

struct WgradKParams {
  int N, Hq, Wq;  // x extent in box coordinates
  int Ho, Wo;
  int taps_h, taps_w, dil, pad_h, pad_w, stride;
  int x_coff, x_cs;
  int cin_per_group, group_nstride;
  int bn, bh, bw;
  int tiles_w, tiles_h, tiles_n, m_tiles;
  int splits, tiles_per_split;
  int co_blocks, ci_blocks, block_n;
  int cpi;         // cout blocks per work item (1 or 2): they share the X box of every k-step
  int co_items;    // co_blocks / cpi
  int nacc;        // accumulator SETS in TMEM (2 = double buffered)
  int cout, cin;
  int ckx;     // channels per X box (64, or 16 for the 16-channel inputs)
  int nterms;
  int stages;
  uint32_t a_bytes, b_bytes, a_box_bytes, b_box_bytes;
  uint32_t idesc;
  uint32_t tmem_cols;
  int acc_stride;  // TMEM columns between the two accumulator buffers (>= 32)
  float* scratch;
};

__device__ __forceinline__ uint64_t make_smem_desc_mnmajor(uint32_t smem_addr, uint32_t swizzle_bytes,
                                                           uint32_t lbo_bytes) {
  const uint64_t layout = (swizzle_bytes == 128) ? 2ull : (swizzle_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;        // stride between MN blocks
  d |= static_cast<uint64_t>((8u * swizzle_bytes) >> 4) << 32;         // stride between 8-row K atoms
  d |= static_cast<uint64_t>(1) << 46;
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ bool wg_tap_box(const WgradKParams& p, int h0, int w0, int kh, int kw, int& ch, int& cw,
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

struct WgItem {
  int split, tap, cob, cib;
};
__device__ __forceinline__ WgItem wg_decode(const WgradKParams& p, int item) {
  WgItem it;
  it.cib = item % p.ci_blocks;
  item /= p.ci_blocks;
  it.cob = (item % p.co_items) * p.cpi;
  item /= p.co_items;
  const int taps = p.taps_h * p.taps_w;
  it.tap = item % taps;
  it.split = item / taps;
  return it;
}

__global__ void __launch_bounds__(256, 1)
    conv_wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ0, const __grid_constant__ CUtensorMap tmZ1,
                              const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                              const WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const uint32_t bars = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kWgMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kWgMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kWgMaxStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kWgMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.taps_h * p.taps_w;
  const int total_items = p.splits * taps * p.co_items * p.ci_blocks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmZ0);
    tma_prefetch_desc(&tmX0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // programmatic dependent launch: barrier init and TMEM allocation above overlapped the previous kernel's tail
  pdl_enter();
  const int a_boxes = 2 * p.cpi;                 // 128 cout = 2 x 64 per cout block
  const int b_boxes = p.block_n / p.ckx;         // cin block = b_boxes x ckx

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    int s = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const int kh = it.tap / p.taps_w, kw = it.tap - kh * p.taps_w;
      const int t0 = it.split * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.m_tiles);
      const int ci0 = it.cib * p.block_n;
      const int g = ci0 / p.cin_per_group;
      const int cc0 = ci0 - g * p.cin_per_group;
      for (int mt = t0; mt < t1; ++mt) {
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int n0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
        int ch, cw, ph, pw;
        if (!wg_tap_box(p, h0, w0, kh, kw, ch, cw, ph, pw)) continue;
        for (int term = 0; term < p.nterms; ++term) {
          mbar_wait(empty_bar(s), phase ^ 1u);
          const uint32_t a_dst = smem_base + s * stage_bytes;
          const uint32_t b_dst = a_dst + p.a_bytes;
          if (elect_one()) {
            mbar_arrive_expect_tx(full_bar(s), stage_bytes);
            const CUtensorMap* mz = (term == 1) ? &tmZ1 : &tmZ0;
            const CUtensorMap* mx = (term == 2) ? &tmX1 : &tmX0;
            for (int bx = 0; bx < a_boxes; ++bx) {
              tma_load_5d(mz, a_dst + bx * p.a_box_bytes, full_bar(s), it.cob * 128 + bx * 64, w0, 0, h0, n0);
            }
            for (int bx = 0; bx < b_boxes; ++bx) {
              tma_load_5d(mx, b_dst + bx * p.b_box_bytes, full_bar(s), p.x_coff + cc0 + bx * p.ckx + pw * p.x_cs, cw,
                          ph, ch, n0 + g * p.group_nstride);
            }
          }
          __syncwarp();
          if (++s == p.stages) {
            s = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    int s = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t swz_b = p.ckx * 2;            // 128 or 32 byte swizzle rows for X
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const int kh = it.tap / p.taps_w, kw = it.tap - kh * p.taps_w;
      const int t0 = it.split * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.m_tiles);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tcgen05_after_thread_sync();
      const uint32_t tmem_d = tmem_base + acc * p.cpi * p.acc_stride;
      uint32_t accumulate = 0;
      for (int mt = t0; mt < t1; ++mt) {
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        int ch, cw, ph, pw;
        if (!wg_tap_box(p, th * p.bh, tw * p.bw, kh, kw, ch, cw, ph, pw)) continue;
        for (int term = 0; term < p.nterms; ++term) {
          mbar_wait(full_bar(s), phase);
          tcgen05_after_thread_sync();
          const uint32_t a_addr = smem_base + s * stage_bytes;
          const uint32_t b_addr = a_addr + p.a_bytes;
          const uint64_t adesc = make_smem_desc_mnmajor(a_addr, 128, p.a_box_bytes);
          const uint64_t bdesc = make_smem_desc_mnmajor(b_addr, swz_b, p.b_box_bytes);
          // 16 pixels (two 8-row K atoms) per MMA
          const uint32_t a_step = (2u * 8u * 128u) >> 4;
          const uint32_t b_step = (2u * 8u * swz_b) >> 4;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < p.cpi; ++j) {
              // cout block j of the item: its two A boxes follow block j-1's, its accumulator sits block_n columns on
              const uint64_t aj = adesc + static_cast<uint64_t>((2u * p.a_box_bytes * j) >> 4);
              const uint32_t dj = tmem_d + static_cast<uint32_t>(j * p.acc_stride);
#pragma unroll
              for (int k = 0; k < kWgPix / 16; ++k) {
                umma_f16(dj, aj + a_step * k, bdesc + b_step * k, p.idesc, (accumulate | k) ? 1u : 0u);
              }
            }
            umma_commit(empty_bar(s));
          }
          __syncwarp();
          accumulate = 1;
          if (++s == p.stages) {
            s = 0;
            phase ^= 1u;
          }
        }
      }
      if (accumulate == 0) {
        // every pixel tile of this (tap, split) was out of bounds: nothing was accumulated, the
        // epilogue must write zeros -> flag it through the (otherwise unused) top bit of acc slot
      }
      if (elect_one()) umma_commit(tfull_bar(acc));
      __syncwarp();
      if (++acc == p.nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> fp32 partial dW =====================
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      // recompute whether any pixel tile contributed (same predicate as producer / issuer)
      const int kh = it.tap / p.taps_w, kw = it.tap - kh * p.taps_w;
      const int t0 = it.split * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.m_tiles);
      bool any = false;
      for (int mt = t0; mt < t1 && !any; ++mt) {
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        int ch, cw, ph, pw;
        any = wg_tap_box(p, th * p.bh, tw * p.bw, kh, kw, ch, cw, ph, pw);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_after_thread_sync();
      for (int j = 0; j < p.cpi; ++j) {
        const int co = (it.cob + j) * 128 + ew * 32 + lane;
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (acc * p.cpi + j) * p.acc_stride;
        float* dst = p.scratch + ((static_cast<long long>(it.split) * taps + it.tap) * p.cout + co) * p.cin +
                     it.cib * p.block_n;
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t r[32];
          // block_n == 16: only 16 valid columns; read 32 (the allocation is >= 32 columns) and store 16
          tmem_ld_32x32b_x32(taddr0 + c0, r);
          tmem_ld_wait();
          if (co < p.cout) {
            const int ncol = min(32, p.block_n - c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (q * 4 < ncol) {
                float4 v;
                v.x = any ? __uint_as_float(r[4 * q + 0]) : 0.f;
                v.y = any ? __uint_as_float(r[4 * q + 1]) : 0.f;
                v.z = any ? __uint_as_float(r[4 * q + 2]) : 0.f;
                v.w = any ? __uint_as_float(r[4 * q + 3]) : 0.f;
                *reinterpret_cast<float4*>(dst + c0 + 4 * q) = v;
              }
            }
          }
        }
      }
      tcgen05_before_thread_sync();
      mbar_arrive(tempty_bar(acc));
      if (++acc == p.nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tcgen05_before_thread_sync();
  __syncthreads();
  if (warp == 2) {
    tcgen05_after_thread_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Sum the pixel splits and scatter [tap][cout][cin] -> OIHW fp32 [cout_real][cin_real][kh][kw].
__global__ void wgrad_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int splits, int taps,
                                    int cout, int cin, int cout_real, int cin_real, int accumulate) {
  pdl_enter();
  const long long total = static_cast<long long>(cout_real) * cin_real * taps;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  // thread order: ci fastest (coalesced scratch reads), then tap, then co
  const int ci = static_cast<int>(i % cin_real);
  long long t = i / cin_real;
  const int tap = static_cast<int>(t % taps);
  const int co = static_cast<int>(t / taps);
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    s += scratch[((static_cast<long long>(sp) * taps + tap) * cout + co) * cin + ci];
  }
  float* o = dw + (static_cast<long long>(co) * cin_real + ci) * taps + tap;
  *o = accumulate ? (*o + s) : s;
}

static int wg_ensure_device(DeviceInfo*& di) {
  di = device_info();
  if (!di) return UP_ERR_CUDA;
  if (!di->wgrad_attr) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(di->max_smem)),
                        "cudaFuncSetAttribute(wgrad smem)");
    if (rc) return rc;
    di->wgrad_attr = true;
  }
  return 0;
}

static int wgrad_plan(const UpConvDesc* d, WgradKParams& p, int sm_count) {
  const int groups = d->x_groups > 0 ? d->x_groups : 1;
  const int cin_g = d->cin / groups;
  p.ckx = (cin_g % 64 == 0) ? 64 : 16;
  int block_n;
  if (p.ckx == 16) {
    block_n = 16;
    if (cin_g != 16) return fail(UP_ERR_UNSUPPORTED, "up_conv2d_wgrad: cin per group must be 16 or a multiple of 64");
  } else {
    block_n = cin_g % 256 == 0 ? 256 : (cin_g % 128 == 0 ? 128 : 64);
  }
  p.block_n = block_n;
  p.ci_blocks = d->cin / block_n;
  p.co_blocks = (d->cout + 127) / 128;
  pick_tile(d->n, d->ho, d->wo, p.bn, p.bh, p.bw, kWgPix);
  p.tiles_w = (d->wo + p.bw - 1) / p.bw;
  p.tiles_h = (d->ho + p.bh - 1) / p.bh;
  p.tiles_n = (d->n + p.bn - 1) / p.bn;
  p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  // (cout blocks per item, pixel splits): minimise a time model of the persistent grid, in microseconds, calibrated on
  // B200 (profiles/train_r2_profile.txt, ncu of the layer3 shapes):
  //   k-step: the main loop runs at the L2 -> SM limit, 11.5 ps per staged byte (48 KB: 0.55 us);
  //   item:   prologue / pipeline fill / epilogue, 5 us per cout block of the item;
  //   splits: the scratch is written and read once per split at ~3 TB/s, + 1.5 us for the reduction launch.
  // cpi = 2 halves the X stream but doubles the splits needed to fill the SMs: it wins for the big layers (decoder 3x3
  // 320->256: 402 -> 238 us, 3x3 512->512: 111 -> 99 us) and loses for layer3's (3x3 256->256: 44 -> 56 us).
  // ceil(SMs / items) alone had put 162 items on 148 SMs for the 3x3 256->256 layers: two rounds of 32 k-steps.
  const double red_us = static_cast<double>(d->kh) * d->kw * d->cout * d->cin * 4.0 * 2.0 / 3e6;
  const int cpi_hi = (p.co_blocks % 2 == 0 && p.ckx == 64) ? 2 : 1;
  int cpi_lo = 1;
  if (const char* e = getenv("UP_WGRAD_CPI")) {     // tuning: force 1 or 2
    if (e[0] == '2' && cpi_hi == 2) cpi_lo = 2;
  }
  double best_cost = 1e30;
  int best = 1, best_cpi = 1;
  const int max_splits = p.m_tiles < 4 * sm_count ? p.m_tiles : 4 * sm_count;
  for (int cpi = cpi_lo; cpi <= cpi_hi; ++cpi) {
    if (getenv("UP_WGRAD_CPI") && getenv("UP_WGRAD_CPI")[0] == '1' && cpi == 2) break;
    const int base_items = d->kh * d->kw * (p.co_blocks / cpi) * p.ci_blocks;
    const double stage_bytes = static_cast<double>(kWgPix) * 2.0 * (128.0 * cpi + p.block_n);
    const double tk = 11.5e-6 * stage_bytes, t_item = 5.0 * cpi;
    for (int sp = 1; sp <= max_splits; ++sp) {
      const int tps = (p.m_tiles + sp - 1) / sp;
      const int eff = (p.m_tiles + tps - 1) / tps;
      if (eff != sp) continue;
      const long long items = static_cast<long long>(eff) * base_items;
      const long long rounds = (items + sm_count - 1) / sm_count;
      const double cost = static_cast<double>(rounds) * (tps * tk + t_item) + 1.5 + red_us * eff;
      if (cost < best_cost) {
        best_cost = cost;
        best = sp;
        best_cpi = cpi;
      }
    }
  }
  p.cpi = best_cpi;
  p.co_items = p.co_blocks / p.cpi;
  p.tiles_per_split = (p.m_tiles + best - 1) / best;
  p.splits = (p.m_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return 0;
}

}  // namespace up

using namespace up;

extern "C" int64_t up_conv2d_wgrad_scratch_bytes(const UpConvDesc* d) {
  if (!d || d->cin <= 0 || d->cout <= 0) return -1;
  WgradKParams p{};
  DeviceInfo* di = device_info();   // no device (CPU-only host): size for a 148-SM part
  int sm = di ? di->sm_count : 148;
  if (wgrad_plan(d, p, sm) != 0) return -1;
  // upper bound independent of the SM count actually found later: splits <= 148-ish; be generous
  const int64_t per_split = static_cast<int64_t>(d->kh) * d->kw * d->cout * d->cin * 4;
  return per_split * (p.splits + 1);
}

extern "C" int up_conv2d_wgrad(const UpConvDesc* d, const void* x, const void* dz, float* dw_oihw, int cout_real,
                               int cin_real, float* scratch, int64_t scratch_bytes, int accumulate, void* stream) {
  UP_CHECK_ARG(d && x && dz && dw_oihw && scratch, "up_conv2d_wgrad: null argument");
  UP_CHECK_ARG(d->stride == 1 || d->stride == 2, "up_conv2d_wgrad: stride must be 1 or 2");
  UP_CHECK_ARG(d->stride == 1 || (d->h % 2 == 0 && d->w % 2 == 0), "up_conv2d_wgrad: stride 2 needs even h, w");
  UP_CHECK_ARG(d->cin % 16 == 0 && d->cout % 64 == 0, "up_conv2d_wgrad: cin %% 16, cout %% 64 required (cin %d cout %d)",
               d->cin, d->cout);
  UP_CHECK_ARG(cout_real > 0 && cout_real <= d->cout && cin_real > 0 && cin_real <= d->cin,
               "up_conv2d_wgrad: bad real channel counts");
  UP_CHECK_ARG(d->dtype == UP_BF16 || d->dtype == UP_FP16 || d->dtype == UP_SPLIT, "up_conv2d_wgrad: bad dtype");
  UP_CHECK_ARG(d->y_cstride == d->cout && d->y_coff == 0, "up_conv2d_wgrad: dz must be a dense [n,ho,wo,cout] tensor");
  const int groups = d->x_groups > 0 ? d->x_groups : 1;
  UP_CHECK_ARG(d->cin % groups == 0, "up_conv2d_wgrad: cin not divisible by x_groups");
  DeviceInfo* di = nullptr;
  int rc = wg_ensure_device(di);
  if (rc) return rc;
  const int g_wg_sm_count = di->sm_count;
  const size_t g_wg_max_smem = di->max_smem;

  const bool split = d->dtype == UP_SPLIT;
  const int fmt = fmt_of_dtype(d->dtype);
  WgradKParams p{};
  rc = wgrad_plan(d, p, g_wg_sm_count);
  if (rc) return rc;
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
  p.x_coff = d->x_coff;
  p.x_cs = d->x_cstride;
  p.cin_per_group = d->cin / groups;
  p.group_nstride = groups > 1 ? d->x_group_nstride : 0;
  p.cout = d->cout;
  p.cin = d->cin;
  p.nterms = split ? 3 : 1;
  p.a_box_bytes = kWgPix * 64 * 2;
  p.b_box_bytes = kWgPix * p.ckx * 2;
  p.a_bytes = 2 * p.cpi * p.a_box_bytes;
  p.b_bytes = (p.block_n / p.ckx) * p.b_box_bytes;
  if (p.b_bytes < 1024) p.b_bytes = 1024;  // keep every stage 1024-byte aligned (16-channel X boxes are 2 KB anyway)
  const size_t fixed = 1024 + 8 * (2 * kWgMaxStages + 4) + 16;
  int stages = static_cast<int>((g_wg_max_smem - fixed) / (p.a_bytes + p.b_bytes));
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  UP_CHECK_ARG(stages >= 2, "up_conv2d_wgrad: not enough shared memory");
  p.stages = stages;
  // idesc: fp32 accumulate, both operands MN-major (bits 15, 16)
  p.idesc = make_idesc_f16(static_cast<uint32_t>(fmt), 128, static_cast<uint32_t>(p.block_n)) | (1u << 15) | (1u << 16);
  p.acc_stride = p.block_n < 32 ? 32 : p.block_n;
  uint32_t cols = 32;
  p.nacc = (2 * p.cpi * p.acc_stride <= 512) ? 2 : 1;
  while (cols < static_cast<uint32_t>(p.nacc * p.cpi * p.acc_stride)) cols *= 2;
  p.tmem_cols = cols;
  p.scratch = scratch;
  const int64_t need = static_cast<int64_t>(p.splits) * d->kh * d->kw * d->cout * d->cin * 4;
  UP_CHECK_ARG(scratch_bytes >= need, "up_conv2d_wgrad: scratch too small (%lld < %lld)", (long long)scratch_bytes,
               (long long)need);

  CUtensorMap tmZ0, tmZ1, tmX0, tmX1;
  const uint32_t zbox[5] = {64u, static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh),
                            static_cast<uint32_t>(p.bn)};
  rc = encode_act_map(&tmZ0, fmt, dz, d->n, d->ho, d->wo, d->cout, 1, zbox, 128, "dz");
  if (rc) return rc;
  const uint32_t xbox[5] = {static_cast<uint32_t>(p.ckx), static_cast<uint32_t>(p.bw), 1u,
                            static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
  const int n_total = d->n + (groups - 1) * p.group_nstride;
  rc = encode_act_map(&tmX0, fmt, x, n_total, d->h, d->w, d->x_cstride, d->stride, xbox, p.ckx * 2, "x", d->x_cextent,
                      d->x_wpitch);
  if (rc) return rc;
  if (split) {
    rc = encode_act_map(&tmZ1, fmt, static_cast<const uint16_t*>(dz) + d->y_plane_stride, d->n, d->ho, d->wo, d->cout,
                        1, zbox, 128, "dz.lo");
    if (rc) return rc;
    rc = encode_act_map(&tmX1, fmt, static_cast<const uint16_t*>(x) + d->x_plane_stride, n_total, d->h, d->w,
                        d->x_cstride, d->stride, xbox, p.ckx * 2, "x.lo", d->x_cextent, d->x_wpitch);
    if (rc) return rc;
  } else {
    tmZ1 = tmZ0;
    tmX1 = tmX0;
  }
  const long long total_items = static_cast<long long>(p.splits) * d->kh * d->kw * p.co_items * p.ci_blocks;
  const int grid = static_cast<int>(total_items < g_wg_sm_count ? total_items : g_wg_sm_count);
  const size_t smem = fixed + static_cast<size_t>(stages) * (p.a_bytes + p.b_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = check_cuda(launch_pdl(conv_wgrad_tcgen05_kernel, grid, 256, smem, st, tmZ0, tmZ1, tmX0, tmX1, p),
                      "conv_wgrad_tcgen05_kernel launch");
  if (rc) return rc;
  const long long total = static_cast<long long>(cout_real) * cin_real * d->kh * d->kw;
  rc = check_cuda(launch_pdl(wgrad_reduce_kernel, static_cast<int>((total + 255) / 256), 256, 0, st, scratch, dw_oihw,
                             p.splits, d->kh * d->kw, d->cout, d->cin, cout_real, cin_real, accumulate),
                  "wgrad_reduce_kernel launch");
  if (rc) return rc;
  return 0;
}
