// The whole WASP block (wasp.forward, model/modules/wasp.py:66-90, eval mode) as ONE persistent kernel.
//
//   x [N,h,w,2048] --aspp1 1x1--> x1 --aspp2 3x3 d--> x2 --aspp3--> x3 --aspp4--> x4
//   out = ReLU( bn1( conv1( cat( conv2(conv2(x1..x4)), broadcast(ReLU(bn(conv_gap(mean_hw(x))))) ) ) ) )
//
// with the inference-time algebra of the module mirror: conv2 o conv2 folded into conv1's filter (W1'_i = W1_i W2 W2),
// every BatchNorm scale folded into its filter, and the pooling branch - constant over h x w - folded into a
// per-image bias of conv1.
//
// Structure: the dependencies of the cascade are PER IMAGE (aspp_{s+1} of an image needs aspp_s of the same image
// only), so there is no grid-wide barrier.  Every CTA owns one 128-pixel tile (bn x bh x bw, the conv kernel's
// tiling) for the whole chain; CTA pairs (tcgen05 cta_group::2, M = 256) take the same tile position of two
// consecutive image groups.  Per stage s:
//     main GEMM   acc[128 x 256] = sum over in-bounds taps / 64-channel chunks  (TMA ring -> tcgen05.mma -> TMEM)
//     epilogue    x_s = ReLU(acc + shift_s) -> 16-bit -> the four 128B-swizzled staging buffers
//                   (a) TMA store to the stack S[s] in global memory (the halo source of stage s+1), then a
//                       release-increment of the image group's stage counter;
//                   (b) the SAME staging buffers are the A operand of a second GEMM  D2 += x_s * W1'_s
//                       (conv1's K-split group s) into a second TMEM accumulator - conv1 never re-reads x1..x4.
//     stage s+1   its TMA producer acquires the counter (all tiles of the image group stored x_s) before loading halos.
// After stage 4:  out = ReLU(D2 + bias_img), stored by TMA.
// The pooling branch rides on the otherwise idle epilogue warps: per-tile channel sums of x while the tensor pipe
// runs aspp1 (direct 16-byte loads, the lines are being pulled through L2 by the TMA anyway), then - once the image
// group's counter says all tiles are in - a 1/#tiles slice of the 2048->256 GEMV per CTA, exchanged through global
// memory, and the small 256->256 GEMV that yields bias_img.
//
// Warp roles (384 threads): 0 = TMA producer, 1 = MMA issuer (leader CTA), 2 = TMEM alloc + store/DMA thread,
// 3 = idle, 4..11 = epilogue / pooling math.
#include <cuda.h>
#include <stdlib.h>

#include "up_internal.h"
#include "up_ptx.cuh"

namespace up {

constexpr int kWcThreads = 384;
constexpr int kWcEpiWarp0 = 4;
constexpr int kWcEpiThreads = 256;
constexpr int kWcMaxSlots = 6;
constexpr uint32_t kWcABytes = 16384;      // 128 px x 64 ch
constexpr uint32_t kWcBBytes = 16384;      // this CTA's half (128 of 256 rows) of a filter tile, 64 ch
constexpr uint32_t kWcSlotBytes = kWcABytes + kWcBBytes;
constexpr uint32_t kWcBuf = 16384;         // staging buffer: 128 px x 64 ch
constexpr int kWcGroups = 4;               // 256 output channels = 4 groups of 64
constexpr int kWcCout = 256;
constexpr int kWcMaxCin = 2048;
constexpr int kWcCtrStride = 8;            // counters per image group: [0..3] stage done, [4] pooling slices done
constexpr long long kWcSpinLimit = 6000000000LL;

struct WcStage {
  int taps;            // 1 (1x1) or 3 (3x3, padding = dilation)
  int dil;
  int chunks;          // input channels / 64
  const float* shift;  // [256]
};

struct WcParams {
  int N, H, W, cin;
  int bn, bh, bw;
  int tiles_h, tiles_w, tiles_n;
  int pairs_per_wave, waves;
  int slots;
  int c_per;                  // pooling-branch output channels computed per tile (power of two >= 8)
  int l2_prefetch;            // 1: bulk L2 prefetch of the tile's x rows before aspp1's k-loop
  int gap_smem;               // 1: the pooling sums read aspp1's activation tiles from the TMA ring (no second pass over x)
  uint32_t idesc;
  WcStage st[4];
  const uint16_t* x;
  const uint16_t* wg_t;       // pooling 1x1 filter, transposed [cin][256], BN scale folded
  const float* shift_gap;     // [256]
  const uint16_t* w5_t;       // conv1' pooling group, transposed [256][256], bn1 scale folded
  const float* shift1;        // [256]
  float* gsum;                // [tiles_h*tiles_w][N][cin] per-tile channel sums
  float* g2;                  // [N][256] pooling-branch activations
  unsigned int* counters;     // [tiles_n][kWcCtrStride] + exit counter
  float inv_hw;
  unsigned long long* dbg;    // optional [grid][32] phase timestamps (UP_DEBUG_TIMING=1)
};

#define WC_STAMP(slot)                                                                     \
  do {                                                                                     \
    if (p.dbg) {                                                                           \
      unsigned long long _t;                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                              \
      p.dbg[blockIdx.x * 32 + (slot)] = _t;                                                \
    }                                                                                      \
  } while (0)

struct WcTile {
  int n0, h0, w0, tn, t;
  bool active;
};

__device__ __forceinline__ WcTile wc_tile(const WcParams& p, int wave, int cluster_id, int crank) {
  const int per = p.tiles_h * p.tiles_w;
  const int pair = wave * p.pairs_per_wave + cluster_id / per;
  WcTile r;
  r.t = cluster_id % per;
  r.active = 2 * pair < p.tiles_n;
  r.tn = 2 * pair + crank;
  r.n0 = r.tn * p.bn;
  r.h0 = (r.t / p.tiles_w) * p.bh;
  r.w0 = (r.t % p.tiles_w) * p.bw;
  return r;
}

// inclusive range of taps (offset (k-1)*dil) whose box [x0+off, x0+off+ext) touches [0, limit)
__device__ __forceinline__ void wc_taps(int taps, int dil, int x0, int ext, int limit, int& lo, int& hi) {
  if (taps == 1) {
    lo = hi = 0;
    return;
  }
  lo = 3;
  hi = -1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int c = x0 + (k - 1) * dil;
    if (c + ext > 0 && c < limit) {
      lo = min(lo, k);
      hi = max(hi, k);
    }
  }
}

// ring k-blocks one tile consumes per wave: per stage the in-bounds taps x channel chunks + 4 second-GEMM chunks
__device__ __forceinline__ int wc_wave_kblocks(const WcParams& p, const WcTile& t) {
  int total = 0;
  for (int s = 0; s < 4; ++s) {
    int kh_lo, kh_hi, kw_lo, kw_hi;
    wc_taps(p.st[s].taps, p.st[s].dil, t.h0, p.bh, p.H, kh_lo, kh_hi);
    wc_taps(p.st[s].taps, p.st[s].dil, t.w0, p.bw, p.W, kw_lo, kw_hi);
    const int ntaps = (kh_hi - kh_lo + 1) * (kw_hi - kw_lo + 1);
    // stage 0: one slot per chunk; later stages: the centre tap takes two merged slots; second GEMM: two merged slots
    total += (s == 0 ? ntaps * p.st[s].chunks : (ntaps - 1) * p.st[s].chunks + 2) + 2;
  }
  return total;
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spin until *ctr >= target (another CTA's release-increment); a watchdog turns a protocol bug into a trap.
__device__ __forceinline__ void wc_wait_counter(const unsigned int* ctr, unsigned int target) {
  if (ld_acquire_gpu(ctr) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(ctr) < target) {
    __nanosleep(64);
    if (clock64() - t0 > kWcSpinLimit) {
      printf("up: wasp chain counter watchdog: block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

template <int kFmt>
__device__ __forceinline__ void wc_add8(const uint4& v, float (&a)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    a[2 * e + 0] += cvt16_to_f32<kFmt>(static_cast<uint16_t>(w[e] & 0xFFFFu));
    a[2 * e + 1] += cvt16_to_f32<kFmt>(static_cast<uint16_t>(w[e] >> 16));
  }
}

template <int kFmt>
__device__ __forceinline__ uint32_t wc_pack2_relu(float lo_elem, float hi_elem) {
  uint32_t d;
  if constexpr (kFmt == 0) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

template <int kFmt>
__global__ void __launch_bounds__(kWcThreads, 1)
    wasp_chain_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmS,
                      const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmB0,
                      const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                      const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmBc,
                      const __grid_constant__ CUtensorMap tmXp, const WcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t staging = smem_base + p.slots * kWcSlotBytes;
  const uint32_t pool = staging + kWcGroups * kWcBuf;           // mean[2][cin] | g2[2][256] | bias[2][256] | red2[8][256]
  const uint32_t pool_bytes = (2u * kWcMaxCin + 2u * kWcCout + 2u * kWcCout + 8u * kWcCout) * 4u;
  const uint32_t bars = pool + pool_bytes;
  float* s_red = reinterpret_cast<float*>(smem_al + (staging - smem_base));       // [8][cin], aliases the staging buffers
  float* s_mean = reinterpret_cast<float*>(smem_al + (pool - smem_base));         // [2][cin]
  float* s_g2 = s_mean + 2 * kWcMaxCin;                                           // [2][256]
  float* s_bias = s_g2 + 2 * kWcCout;                                             // [2][256]
  float* s_red2 = s_bias + 2 * kWcCout;                                           // [8][256]
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kWcMaxSlots + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * kWcMaxSlots);
  const uint32_t tempty_bar = tfull_bar + 8u;
  const uint32_t d2full_bar = tfull_bar + 16u;
  const uint32_t d2empty_bar = tfull_bar + 24u;
  auto avail_bar = [&](int g) { return tfull_bar + 32u + 8u * g; };
  auto ready_bar = [&](int g) { return tfull_bar + 32u + 8u * (kWcGroups + g); };
  auto s2ready_bar = [&](int g) { return tfull_bar + 32u + 8u * (2 * kWcGroups + g); };
  auto gapdone_bar = [&](int sl) { return tfull_bar + 32u + 8u * (3 * kWcGroups + sl); };
  const uint32_t tmem_slot = tfull_bar + 32u + 8u * (3 * kWcGroups + kWcMaxSlots);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int cluster_id = blockIdx.x >> 1;
  const int per = p.tiles_h * p.tiles_w;
  if (threadIdx.x == 0) WC_STAMP(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmB0);
    tma_prefetch_desc(&tmBc);
    for (int s = 0; s < p.slots; ++s) {
      mbar_init(full_bar(s), 2);     // both CTAs' producers arrive on the leader's barrier
      mbar_init(empty_bar(s), 1);    // one multicast commit from the leader's issuer
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * (kWcEpiThreads / 32));
    mbar_init(d2full_bar, 1);
    mbar_init(d2empty_bar, 2 * (kWcEpiThreads / 32));
    for (int g = 0; g < kWcGroups; ++g) {
      mbar_init(avail_bar(g), 2);                               // store drained + second GEMM done reading
      mbar_init(ready_bar(g), kWcEpiThreads / 32);              // local epilogue warps -> store thread
      mbar_init(s2ready_bar(g), 2 * (kWcEpiThreads / 32));      // epilogue warps of BOTH CTAs -> MMA issuer
    }
    for (int sl = 0; sl < p.slots; ++sl) mbar_init(gapdone_bar(sl), kWcEpiThreads / 32);   // pooling warps -> producer
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, 512);
  tcgen05_before_thread_sync();
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_d2 = tmem_base + kWcCout;
  // the chain reads the output of the previous kernel in the stream
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) WC_STAMP(1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (p.l2_prefetch && lane == 0) {
      // Pull this CTA's whole aspp1 tile (128 pixels x cin channels) towards L2 in 512-byte rows before the k-loop
      // starts reading it in 128-byte rows: when x comes from DRAM, 64-channel boxes 4 KB apart open a DRAM page per
      // pixel and k-block; the prefetch touches every page once with 4x longer bursts, and all requests are in flight
      // at the same time.
      const WcTile t0 = wc_tile(p, 0, cluster_id, crank);
      if (t0.active) {
        for (int c = 0; c < p.cin; c += 256) {
          asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmXp)),
                       "r"(c), "r"(t0.w0), "r"(0), "r"(t0.h0), "r"(t0.n0)
                       : "memory");
        }
      }
    }
    __syncwarp();
    uint32_t slot = 0, par = 1;
    auto advance = [&]() {
      if (++slot == static_cast<uint32_t>(p.slots)) {
        slot = 0;
        par ^= 1u;
      }
    };
    uint32_t nwave = 0;
    const int chunks0 = p.st[0].chunks;
    for (int wave = 0; wave < p.waves; ++wave) {
      const WcTile t = wc_tile(p, wave, cluster_id, crank);
      if (!t.active) continue;
      int kbw = 0;      // k-block index within this wave
      // a slot last filled by an aspp1 k-block is also read by the pooling warps: wait for them before refilling it
      auto wait_slot = [&]() {
        mbar_wait(empty_bar(slot), par, 16000000000LL);
        if (p.gap_smem && kbw >= p.slots && kbw < chunks0 + p.slots) {
          const uint32_t use = nwave * static_cast<uint32_t>(chunks0 / p.slots) + static_cast<uint32_t>((kbw - p.slots) / p.slots);
          mbar_wait(gapdone_bar(slot), use & 1u, 16000000000LL);
        }
        ++kbw;
      };
      for (int s = 0; s < 4; ++s) {
        const WcStage st = p.st[s];
        int kh_lo, kh_hi, kw_lo, kw_hi;
        wc_taps(st.taps, st.dil, t.h0, p.bh, p.H, kh_lo, kh_hi);
        wc_taps(st.taps, st.dil, t.w0, p.bw, p.W, kw_lo, kw_hi);
        const CUtensorMap* amap = s == 0 ? &tmX : &tmS;
        const CUtensorMap* bmap = s == 0 ? &tmB0 : (s == 1 ? &tmB1 : (s == 2 ? &tmB2 : &tmB3));
        const int an = s == 0 ? t.n0 : t.n0 + (s - 1) * p.N;
        // two 128-row filter chunks (64 channels each) in one slot: the B operands of two B-only k-steps
        auto issue_b2 = [&](const CUtensorMap* m, int c0, int row) {
          wait_slot();
          if (elect_one()) {
            const uint32_t dst = smem_base + slot * kWcSlotBytes;
            if (crank == 0) mbar_arrive_expect_tx(full_bar(slot), 2u * kWcSlotBytes);
            else mbar_arrive_remote(full_bar(slot), 0u);
            tma_load_2d_2cta(m, dst, full_bar(slot), c0, row);
            tma_load_2d_2cta(m, dst + kWcABytes, full_bar(slot), c0 + 64, row);
          }
          __syncwarp();
          advance();
        };
        if (s > 0) {
          // centre tap first: its activation tile is this CTA's own x_s tile, still in the staging buffers - filter
          // chunks only, and nothing here waits for the neighbours
          const int crow = 4 * kWcCout + static_cast<int>(crank) * (kWcCout / 2);
          issue_b2(bmap, 0, crow);
          issue_b2(bmap, 128, crow);
          // the other taps: every tile of this image group has stored x_s (the halo source) to global memory
          if (lane == 0) wc_wait_counter(p.counters + t.tn * kWcCtrStride + (s - 1), per);
          __syncwarp();
          fence_proxy_async_all();
        }
        if (wave == 0 && lane == 0) WC_STAMP(4 + 6 * s);      // dependencies of stage s satisfied
        for (int kh = kh_lo; kh <= kh_hi; ++kh) {
          const int oh = st.taps == 1 ? 0 : (kh - 1) * st.dil;
          for (int kw = kw_lo; kw <= kw_hi; ++kw) {
            if (s > 0 && kh == 1 && kw == 1) continue;
            const int ow = st.taps == 1 ? 0 : (kw - 1) * st.dil;
            const int brow = (st.taps == 1 ? 0 : (kh * 3 + kw)) * kWcCout + static_cast<int>(crank) * (kWcCout / 2);
            for (int chunk = 0; chunk < st.chunks; ++chunk) {
              wait_slot();
              if (elect_one()) {
                const uint32_t dst = smem_base + slot * kWcSlotBytes;
                if (crank == 0) mbar_arrive_expect_tx(full_bar(slot), 2u * kWcSlotBytes);
                else mbar_arrive_remote(full_bar(slot), 0u);
                tma_load_5d_2cta(amap, dst, full_bar(slot), chunk * 64, t.w0 + ow, 0, t.h0 + oh, an);
                tma_load_2d_2cta(bmap, dst + kWcABytes, full_bar(slot), chunk * 64, brow);
              }
              __syncwarp();
              advance();
            }
          }
        }
        // conv1' K-split group s: four 64-channel filter chunks for the second GEMM (B operand only), two per slot
        issue_b2(&tmBc, s * kWcCout, static_cast<int>(crank) * (kWcCout / 2));
        issue_b2(&tmBc, s * kWcCout + 128, static_cast<int>(crank) * (kWcCout / 2));
      }
      ++nwave;
    }
  } else if (warp == 1 && crank == 0) {
    // ===================== MMA issuer (leader CTA) =====================
    uint32_t slot = 0, phase = 0;
    auto advance = [&]() {
      if (++slot == static_cast<uint32_t>(p.slots)) {
        slot = 0;
        phase ^= 1u;
      }
    };
    const uint64_t adesc0 = make_smem_desc_kmajor(smem_base, 128);
    const uint64_t bdesc0 = make_smem_desc_kmajor(smem_base + kWcABytes, 128);
    const uint64_t sdesc0 = make_smem_desc_kmajor(staging, 128);
    const uint32_t slot_step = kWcSlotBytes >> 4;
    uint32_t nst = 0;     // stages processed (tempty / s2ready parity)
    uint32_t nwave = 0;   // waves processed (d2empty parity)
    for (int wave = 0; wave < p.waves; ++wave) {
      const WcTile t = wc_tile(p, wave, cluster_id, crank);
      if (!t.active) continue;
      for (int s = 0; s < 4; ++s) {
        const WcStage st = p.st[s];
        int kh_lo, kh_hi, kw_lo, kw_hi;
        wc_taps(st.taps, st.dil, t.h0, p.bh, p.H, kh_lo, kh_hi);
        wc_taps(st.taps, st.dil, t.w0, p.bw, p.W, kw_lo, kw_hi);
        const int ntaps = (kh_hi - kh_lo + 1) * (kw_hi - kw_lo + 1);
        const int nkb = (s == 0 ? ntaps : ntaps - 1) * st.chunks;
        mbar_wait(tempty_bar, (nst & 1u) ^ 1u);     // the epilogue of the previous stage has drained the accumulator
        tcgen05_after_thread_sync();
        if (s > 0) {
          // centre tap: A = this CTA's x_s tile in the staging buffers (their s2ready phases were awaited by the second
          // GEMM of the previous stage), B = two filter chunks per slot
          for (int s2 = 0; s2 < 2; ++s2) {
            mbar_wait(full_bar(slot), phase);
            tcgen05_after_thread_sync();
            if (elect_one()) {
              for (int cc = 0; cc < 2; ++cc) {
                const int c = 2 * s2 + cc;
                const uint64_t ad = sdesc0 + static_cast<uint64_t>((kWcBuf >> 4) * c);
                const uint64_t bd = adesc0 + static_cast<uint64_t>(slot_step * slot) + static_cast<uint64_t>((kWcABytes >> 4) * cc);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_2cta(tmem_acc, ad + 2u * k, bd + 2u * k, p.idesc, (c | k) ? 1u : 0u);
              }
              umma_commit_2cta_mc(empty_bar(slot), 3);
              if (s2 == 1) {
                for (int g = 0; g < kWcGroups; ++g) umma_commit_2cta_mc(avail_bar(g), 3);   // x_s tile consumed
                if (nkb == 0) umma_commit_2cta_mc(tfull_bar, 3);
              }
            }
            __syncwarp();
            advance();
          }
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(slot), phase);
          tcgen05_after_thread_sync();
          if (elect_one()) {
            const uint64_t ad = adesc0 + static_cast<uint64_t>(slot_step * slot);
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(slot_step * slot);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2cta(tmem_acc, ad + 2u * k, bd + 2u * k, p.idesc, (s > 0 || (kb | k)) ? 1u : 0u);
            umma_commit_2cta_mc(empty_bar(slot), 3);
            if (kb == nkb - 1) {
              umma_commit_2cta_mc(tfull_bar, 3);
              if (wave == 0) WC_STAMP(5 + 6 * s);              // main loop of stage s issued
            }
          }
          __syncwarp();
          advance();
        }
        // second GEMM: D2 += x_s (the staging buffers the epilogue is filling) x W1'_s, two filter chunks per slot
        if (s == 0) {
          mbar_wait(d2empty_bar, (nwave & 1u) ^ 1u);
          tcgen05_after_thread_sync();
        }
        for (int s2 = 0; s2 < 2; ++s2) {
          mbar_wait(full_bar(slot), phase);
          for (int cc = 0; cc < 2; ++cc) {
            const int g = 2 * s2 + cc;
            mbar_wait(s2ready_bar(g), nst & 1u);
            tcgen05_after_thread_sync();
            if (elect_one()) {
              const uint64_t ad = sdesc0 + static_cast<uint64_t>((kWcBuf >> 4) * g);
              const uint64_t bd = adesc0 + static_cast<uint64_t>(slot_step * slot) + static_cast<uint64_t>((kWcABytes >> 4) * cc);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_2cta(tmem_d2, ad + 2u * k, bd + 2u * k, p.idesc, (s | g | k) ? 1u : 0u);
              // the staging buffer is released by the centre tap of the next stage; after the last stage, here
              if (s == 3) umma_commit_2cta_mc(avail_bar(g), 3);
              if (cc == 1) {
                umma_commit_2cta_mc(empty_bar(slot), 3);
                if (s == 3 && g == kWcGroups - 1) umma_commit_2cta_mc(d2full_bar, 3);
                if (wave == 0 && g == kWcGroups - 1) WC_STAMP(9 + 6 * s);   // second GEMM of stage s issued
              }
            }
            __syncwarp();
          }
          advance();
        }
        ++nst;
      }
      ++nwave;
    }
  } else if (threadIdx.x == 64) {
    // ===================== store / DMA thread =====================
    uint32_t nuse = 0;
    for (int wave = 0; wave < p.waves; ++wave) {
      const WcTile t = wc_tile(p, wave, cluster_id, crank);
      if (!t.active) continue;
      for (int s = 0; s < 4; ++s) {
        for (int g = 0; g < kWcGroups; ++g) {
          mbar_wait(ready_bar(g), nuse & 1u);
          tma_store_5d(&tmS, staging + g * kWcBuf, g * 64, t.w0, 0, t.h0, t.n0 + s * p.N);
          tma_store_commit();
        }
        tma_store_wait_read<0>();
        for (int g = 0; g < kWcGroups; ++g) mbar_arrive(avail_bar(g));
        tma_store_wait_all<0>();          // x_s of this tile is in global memory ...
        fence_proxy_async_all();
        __threadfence();
        red_release_gpu_add(p.counters + t.tn * kWcCtrStride + s, 1u);   // ... tell the image group
        if (wave == 0) WC_STAMP(8 + 6 * s);
        ++nuse;
      }
      for (int g = 0; g < kWcGroups; ++g) {
        mbar_wait(ready_bar(g), nuse & 1u);
        tma_store_5d(&tmO, staging + g * kWcBuf, g * 64, t.w0, 0, t.h0, t.n0);
        tma_store_commit();
      }
      tma_store_wait_read<0>();
      for (int g = 0; g < kWcGroups; ++g) {
        mbar_arrive(avail_bar(g));        // no second GEMM reads the output tile: both arrivals are ours
        mbar_arrive(avail_bar(g));
      }
      ++nuse;
    }
    tma_store_wait_all<0>();
  } else if (warp >= kWcEpiWarp0) {
    // ===================== epilogue + pooling-branch math (8 warps) =====================
    const int ew = warp - kWcEpiWarp0;
    const int etid = threadIdx.x - kWcEpiWarp0 * 32;
    const int quarter = ew & 3;    // TMEM lane quarter this warp may read
    const int half = ew >> 2;      // which 32 of a group's 64 columns
    const int row = quarter * 32 + lane;
    const uint32_t row_smem = staging + static_cast<uint32_t>(row) * 128u;
    const uint32_t row7 = static_cast<uint32_t>(row) & 7u;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    const int rpi = p.bh * p.bw;               // tile rows per image
    const int wpi = rpi / 16;                  // epilogue warps (16 rows each) per image
    uint32_t nuse = 0, nst = 0, nwave = 0;
    uint32_t kbase = 0;       // ring k-blocks of the previous waves (the pooling sums follow aspp1's slots)

    // one 128 x 256 accumulator -> ReLU(acc + shift) -> 16-bit -> staging buffers.  `sh` points at this thread's row of
    // shifts (column 0); second = also hand the buffers to the MMA issuer as the A operand of the second GEMM.
    auto epilogue = [&](uint32_t tmem_col0, const float* sh, bool second) {
      uint32_t r[32];
      const uint32_t taddr = tmem_col0 + tlane + static_cast<uint32_t>(half * 32);
      tmem_ld_32x32b_x32(taddr, r);
      for (int g = 0; g < kWcGroups; ++g) {
        const float4* s4 = reinterpret_cast<const float4*>(sh + g * 64 + half * 32);
        float v[32];
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 h4 = s4[j4];
          v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + h4.x;
          v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + h4.y;
          v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + h4.z;
          v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + h4.w;
        }
        if (g + 1 < kWcGroups) tmem_ld_32x32b_x32(taddr + (g + 1) * 64, r);
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = wc_pack2_relu<kFmt>(v[2 * e], v[2 * e + 1]);
        mbar_wait(avail_bar(g), (nuse & 1u) ^ 1u);       // previous store drained, previous second GEMM done
        const uint32_t rowaddr = row_smem + g * kWcBuf;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t addr = rowaddr + (((static_cast<uint32_t>(half) * 4u + c4) ^ row7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * c4]), "r"(w[4 * c4 + 1]),
                       "r"(w[4 * c4 + 2]), "r"(w[4 * c4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(ready_bar(g));
          if (second) {
            if (crank == 0) mbar_arrive(s2ready_bar(g));
            else mbar_arrive_remote(s2ready_bar(g), 0u);
          }
        }
      }
      ++nuse;
    };

    for (int wave = 0; wave < p.waves; ++wave) {
      const WcTile t = wc_tile(p, wave, cluster_id, crank);
      if (!t.active) continue;
      unsigned int* ctr = p.counters + t.tn * kWcCtrStride;

      // ---- pooling branch, part 1: channel sums of this tile's pixels (nn.AdaptiveAvgPool2d, wasp.py:51) ----
      // The partial-sum scratch aliases the staging buffers: they must be drained (previous wave's output stores).
      for (int g = 0; g < kWcGroups; ++g) mbar_wait(avail_bar(g), (nuse & 1u) ^ 1u);
      if (p.gap_smem) {
        // aspp1's activation tiles pass through the TMA ring: once the MMAs of a k-block have consumed a slot (its
        // `empty` barrier completes, in both CTAs of the pair) and before the producer refills it (it waits for
        // `gapdone`), the eight epilogue warps add up their 16 rows of the 128 x 64 tile straight from shared memory.
        const int c = lane & 7;                    // logical 16-byte chunk = 8 channels
        const int rsub = lane >> 3;                // rows rsub, rsub + 4, rsub + 8, rsub + 12 of this warp's 16
        for (int kb = 0; kb < p.st[0].chunks; ++kb) {
          const uint32_t g = kbase + static_cast<uint32_t>(kb);
          const uint32_t sl = g % static_cast<uint32_t>(p.slots), use = g / static_cast<uint32_t>(p.slots);
          mbar_wait(empty_bar(sl), use & 1u, 16000000000LL);
          const uint32_t tile = smem_base + sl * kWcSlotBytes;
          float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t r = static_cast<uint32_t>(ew * 16 + rsub + 4 * i);
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(tile + r * 128u + ((static_cast<uint32_t>(c) ^ (r & 7u)) << 4)));
            wc_add8<kFmt>(v, a);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            a[e] += __shfl_xor_sync(0xffffffffu, a[e], 8);
            a[e] += __shfl_xor_sync(0xffffffffu, a[e], 16);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(gapdone_bar(sl));           // this warp is done with the slot
          if (rsub == 0) {
            float4* dst = reinterpret_cast<float4*>(s_red + ew * p.cin + kb * 64 + c * 8);
            dst[0] = make_float4(a[0], a[1], a[2], a[3]);
            dst[1] = make_float4(a[4], a[5], a[6], a[7]);
          }
        }
        kbase += static_cast<uint32_t>(wc_wave_kblocks(p, t));
      } else {
        const int nl = (ew * 16) / rpi;            // image of this warp's 16 rows
        const int n = t.n0 + nl;
        const int rem0 = ew * 16 - nl * rpi;
        // 16 pixel rows per warp; all 16 loads of a 256-channel slab are issued before the first add (memory-level
        // parallelism: the loop is latency bound otherwise)
        const uint16_t* prow[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int rem = rem0 + i;
          const int h = t.h0 + rem / p.bw, w = t.w0 + rem % p.bw;
          const bool ok = h < p.H && w < p.W && n < p.N;
          prow[i] = ok ? p.x + ((static_cast<size_t>(n) * p.H + h) * p.W + w) * p.cin + lane * 8 : nullptr;
        }
        for (int cc = 0; cc < p.cin / 256; ++cc) {
          uint4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = prow[i] ? __ldg(reinterpret_cast<const uint4*>(prow[i] + cc * 256)) : make_uint4(0u, 0u, 0u, 0u);
          }
          float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) wc_add8<kFmt>(v[i], a);
          float4* dst = reinterpret_cast<float4*>(s_red + ew * p.cin + cc * 256 + lane * 8);
          dst[0] = make_float4(a[0], a[1], a[2], a[3]);
          dst[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
      }
      {
        named_bar_sync(1, kWcEpiThreads);
        for (int idx = etid; idx < p.bn * p.cin; idx += kWcEpiThreads) {
          const int img = idx / p.cin, ch = idx - img * p.cin;
          float s = 0.f;
          for (int q = 0; q < wpi; ++q) s += s_red[(img * wpi + q) * p.cin + ch];
          if (t.n0 + img < p.N) p.gsum[(static_cast<size_t>(t.t) * p.N + t.n0 + img) * p.cin + ch] = s;
        }
        __threadfence();
        named_bar_sync(1, kWcEpiThreads);          // the scratch is dead: the epilogue may fill the staging buffers
        if (wave == 0 && etid == 0) WC_STAMP(2);
      }

      for (int s = 0; s < 4; ++s) {
        mbar_wait(tfull_bar, nst & 1u);
        if (wave == 0 && etid == 0) WC_STAMP(6 + 6 * s);       // accumulator of stage s complete
        tcgen05_after_thread_sync();
        epilogue(tmem_acc, p.st[s].shift, true);
        tcgen05_before_thread_sync();
        __syncwarp();
        if (lane == 0) {
          if (crank == 0) mbar_arrive(tempty_bar);
          else mbar_arrive_remote(tempty_bar, 0u);
        }
        if (wave == 0 && etid == 0) WC_STAMP(7 + 6 * s);       // epilogue of stage s done
        ++nst;

        if (s == 1) {
          // ---- pooling branch, part 2 (after the stage-1 epilogue: it then runs in the shadow of aspp3's k-loop): mean over the whole image, then this tile's slice of the 1x1 conv
          // (global_avg_pool[1..3], wasp.py:51-54).  All tiles of the image group have published their sums once
          // the group's stage-0 counter is complete (the store thread increments it after the sums were fenced).
          if (etid == 0) wc_wait_counter(ctr + 0, per);
          named_bar_sync(1, kWcEpiThreads);
          for (int idx = etid; idx < p.bn * (p.cin / 4); idx += kWcEpiThreads) {
            const int img = idx / (p.cin / 4), c4 = idx - img * (p.cin / 4);
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t.n0 + img < p.N) {
              const float4* src = reinterpret_cast<const float4*>(p.gsum + static_cast<size_t>(t.n0 + img) * p.cin) + c4;
              const size_t tstride = static_cast<size_t>(p.N) * p.cin / 4;     // float4 elements between tiles
              for (int q0 = 0; q0 < per; q0 += 8) {
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  v[q] = (q0 + q < per) ? __ldcg(src + static_cast<size_t>(q0 + q) * tstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  m.x += v[q].x;
                  m.y += v[q].y;
                  m.z += v[q].z;
                  m.w += v[q].w;
                }
              }
            }
            m.x *= p.inv_hw;
            m.y *= p.inv_hw;
            m.z *= p.inv_hw;
            m.w *= p.inv_hw;
            reinterpret_cast<float4*>(s_mean + img * kWcMaxCin)[c4] = m;
          }
          named_bar_sync(1, kWcEpiThreads);
          const int c0 = t.t * p.c_per;            // this tile's output channels [c0, c0 + c_per)
          if (c0 < kWcCout) {
            const int octs = p.c_per / 8;          // 16-byte groups of output channels (power of two <= 32)
            const int o = lane % octs, ksub = lane / octs, ksubs = 32 / octs;
            const int kper = p.cin / 8;            // this warp's K range
            float a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int k0 = ew * kper + ksub; k0 < (ew + 1) * kper; k0 += 8 * ksubs) {
              uint4 wv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = k0 + u * ksubs;
                wv[u] = (k < (ew + 1) * kper)
                            ? __ldg(reinterpret_cast<const uint4*>(p.wg_t + static_cast<size_t>(k) * kWcCout + c0 + o * 8))
                            : make_uint4(0u, 0u, 0u, 0u);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int k = min(k0 + u * ksubs, (ew + 1) * kper - 1);     // zero filter words beyond the range
                const float m0 = s_mean[k], m1 = s_mean[kWcMaxCin + k];
                const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float w0 = cvt16_to_f32<kFmt>(static_cast<uint16_t>(ww[e] & 0xFFFFu));
                  const float w1 = cvt16_to_f32<kFmt>(static_cast<uint16_t>(ww[e] >> 16));
                  a0[2 * e] = fmaf(w0, m0, a0[2 * e]);
                  a0[2 * e + 1] = fmaf(w1, m0, a0[2 * e + 1]);
                  a1[2 * e] = fmaf(w0, m1, a1[2 * e]);
                  a1[2 * e + 1] = fmaf(w1, m1, a1[2 * e + 1]);
                }
              }
            }
            for (int m = octs; m < 32; m <<= 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                a0[e] += __shfl_xor_sync(0xffffffffu, a0[e], m);
                a1[e] += __shfl_xor_sync(0xffffffffu, a1[e], m);
              }
            }
            if (ksub == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                s_red2[ew * kWcCout + o * 8 + e] = a0[e];
                if (p.bn == 2) s_red2[ew * kWcCout + 128 + o * 8 + e] = a1[e];      // c_per <= 128 when bn == 2
              }
            }
          }
          named_bar_sync(1, kWcEpiThreads);
          if (c0 < kWcCout && etid < p.bn * p.c_per) {
            const int img = etid / p.c_per, c = etid - img * p.c_per;
            float a = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) a += s_red2[q * kWcCout + img * 128 + c];
            a = fmaxf(a + __ldg(p.shift_gap + c0 + c), 0.f);          // (BatchNorm shift) + ReLU, wasp.py:53-54
            if (c0 + c < kWcCout && t.n0 + img < p.N) p.g2[static_cast<size_t>(t.n0 + img) * kWcCout + c0 + c] = a;
          }
          __threadfence();
          named_bar_sync(1, kWcEpiThreads);
          if (etid == 0) red_release_gpu_add(ctr + 4, 1u);
          if (wave == 0 && etid == 0) WC_STAMP(28);
        }
        if (s == 2) {
          // ---- pooling branch, part 3 (in the shadow of aspp4's k-loop): bias_img = W1'_5 . g2 + bn1 shift (the broadcast branch through conv1) ----
          if (etid == 0) wc_wait_counter(ctr + 4, per);
          named_bar_sync(1, kWcEpiThreads);
          for (int idx = etid; idx < p.bn * kWcCout; idx += kWcEpiThreads) {
            const int img = idx / kWcCout, c = idx - img * kWcCout;
            s_g2[idx] = (t.n0 + img < p.N) ? __ldcg(p.g2 + static_cast<size_t>(t.n0 + img) * kWcCout + c) : 0.f;
          }
          named_bar_sync(1, kWcEpiThreads);
          {
            // thread = (32-row slice of j = warp, 8 output channels = lane); partial sums meet in a scratch that
            // aliases s_mean (dead since the first GEMV)
            const int o = lane, j0 = ew * (kWcCout / 8);
            float a0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float a1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int jb = 0; jb < kWcCout / 8; jb += 8) {
              uint4 wv[8];
#pragma unroll
              for (int u = 0; u < 8; ++u)
                wv[u] = __ldg(reinterpret_cast<const uint4*>(p.w5_t + static_cast<size_t>(j0 + jb + u) * kWcCout + o * 8));
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float g0 = s_g2[j0 + jb + u], g1 = s_g2[kWcCout + j0 + jb + u];
                const uint32_t ww[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float w0 = cvt16_to_f32<kFmt>(static_cast<uint16_t>(ww[e] & 0xFFFFu));
                  const float w1 = cvt16_to_f32<kFmt>(static_cast<uint16_t>(ww[e] >> 16));
                  a0[2 * e] = fmaf(w0, g0, a0[2 * e]);
                  a0[2 * e + 1] = fmaf(w1, g0, a0[2 * e + 1]);
                  a1[2 * e] = fmaf(w0, g1, a1[2 * e]);
                  a1[2 * e + 1] = fmaf(w1, g1, a1[2 * e + 1]);
                }
              }
            }
            float* scr = s_mean;      // [8 warps][2 images][256]
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              scr[(ew * 2 + 0) * kWcCout + o * 8 + e] = a0[e];
              scr[(ew * 2 + 1) * kWcCout + o * 8 + e] = a1[e];
            }
            named_bar_sync(1, kWcEpiThreads);
            const int c = etid;
            float b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              b0 += scr[(q * 2 + 0) * kWcCout + c];
              b1 += scr[(q * 2 + 1) * kWcCout + c];
            }
            const float sh = __ldg(p.shift1 + c);
            s_bias[c] = b0 + sh;
            s_bias[kWcCout + c] = b1 + sh;
          }
          named_bar_sync(1, kWcEpiThreads);
          if (wave == 0 && etid == 0) WC_STAMP(29);
        }
      }

      // ---- conv1 (+ bn1 + ReLU): D2 holds the four cascade groups, the pooling group is the per-image bias ----
      mbar_wait(d2full_bar, nwave & 1u);
      tcgen05_after_thread_sync();
      epilogue(tmem_d2, s_bias + (row / rpi) * kWcCout, false);
      tcgen05_before_thread_sync();
      __syncwarp();
      if (lane == 0) {
        if (crank == 0) mbar_arrive(d2empty_bar);
        else mbar_arrive_remote(d2empty_bar, 0u);
      }
      if (wave == 0 && etid == 0) WC_STAMP(30);
      ++nwave;
    }
  }

  tcgen05_before_thread_sync();
  __syncthreads();
  if (threadIdx.x == 0) WC_STAMP(31);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 2) {
    tcgen05_after_thread_sync();
    tmem_dealloc_2cta(tmem_base, 512);
  }
  if (threadIdx.x == 0) {
    // the last CTA to leave re-arms the counters for the next launch (nobody waits on them any more)
    const int n_ctr = p.tiles_n * kWcCtrStride;
    __threadfence();
    const unsigned int old = atomicAdd(p.counters + n_ctr, 1u);
    if (old == gridDim.x - 1) {
      for (int i = 0; i <= n_ctr; ++i) p.counters[i] = 0u;
      __threadfence();
    }
  }
}

}  // namespace up
#include "up_conv_host.h"
namespace up {

struct WcPlan {
  int bn, bh, bw, tiles_h, tiles_w, tiles_n, per, c_per;
};

static int wc_plan(const UpWaspChainDesc* d, WcPlan& pl) {
  if (!d) return fail(UP_ERR_INVALID, "up_wasp_chain: null descriptor");
  if (d->dtype != UP_FP16 && d->dtype != UP_BF16)
    return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: fp16 / bf16 only (the fp32-grade split mode runs the layer-wise plan)");
  if (d->n <= 0 || d->h <= 0 || d->w <= 0) return fail(UP_ERR_INVALID, "up_wasp_chain: bad dims");
  if (d->cin % 256 != 0 || d->cin <= 0 || d->cin > kWcMaxCin)
    return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: cin must be a multiple of 256, <= %d (got %d)", kWcMaxCin, d->cin);
  if (d->conv1_cin < 5 * kWcCout) return fail(UP_ERR_INVALID, "up_wasp_chain: conv1_cin must cover five groups of 256");
  for (int i = 0; i < 3; ++i)
    if (d->dil[i] < 1) return fail(UP_ERR_INVALID, "up_wasp_chain: bad dilation");
  pick_tile(d->n, d->h, d->w, pl.bn, pl.bh, pl.bw);
  if (pl.bn > 2 || (pl.bh * pl.bw) % 16 != 0)
    return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: map %dx%d too small for the fused chain (tile %dx%dx%d)", d->h, d->w,
                pl.bn, pl.bh, pl.bw);
  pl.tiles_w = (d->w + pl.bw - 1) / pl.bw;
  pl.tiles_h = (d->h + pl.bh - 1) / pl.bh;
  if (d->n % (2 * pl.bn) != 0)
    return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: batch %d is not a multiple of %d (CTA pairs take two image groups)",
                d->n, 2 * pl.bn);
  pl.tiles_n = d->n / pl.bn;
  pl.per = pl.tiles_h * pl.tiles_w;
  int c = (kWcCout + pl.per - 1) / pl.per;
  int c_per = 8;
  while (c_per < c) c_per *= 2;
  pl.c_per = c_per;
  if (pl.bn * c_per > kWcEpiThreads || (pl.bn == 2 && c_per > 128))
    return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: too few tiles per image for the pooling-branch split");
  return 0;
}

static size_t wc_align(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }
static unsigned long long* g_wc_dbg = nullptr;

}  // namespace up

using namespace up;

// Debug only (UP_DEBUG_TIMING=1): phase timestamps (globaltimer ns) of the last chain launch, 160 CTAs x 32 slots.
extern "C" int up_debug_chain_timing(unsigned long long* h_out) {
  if (!g_wc_dbg) return up::fail(UP_ERR_INVALID, "no timing buffer (set UP_DEBUG_TIMING=1)");
  return up::check_cuda(cudaMemcpy(h_out, g_wc_dbg, 160 * 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost),
                        "cudaMemcpy(chain timing)");
}

extern "C" int up_wasp_chain_supported(const UpWaspChainDesc* d) {
  WcPlan pl;
  int rc = wc_plan(d, pl);
  if (rc) return rc;
  DeviceInfo* di = device_info();
  if (di && pl.per > di->sm_count / 2) return fail(UP_ERR_UNSUPPORTED, "up_wasp_chain: more tiles per image than CTA pairs");
  return 0;
}

extern "C" int64_t up_wasp_chain_workspace_bytes(const UpWaspChainDesc* d) {
  WcPlan pl;
  if (wc_plan(d, pl)) return -1;
  const size_t ctr = wc_align(static_cast<size_t>(pl.tiles_n * kWcCtrStride + 1) * 4);
  const size_t g2 = wc_align(static_cast<size_t>(d->n) * kWcCout * 4);
  const size_t gsum = wc_align(static_cast<size_t>(pl.per) * d->n * d->cin * 4);
  return static_cast<int64_t>(ctr + g2 + gsum);
}

extern "C" int up_wasp_chain_fwd(const UpWaspChainDesc* d, const UpWaspChainWeights* w, const void* x, void* s_stack,
                                 void* out, void* workspace, int64_t workspace_bytes, void* stream) {
  UP_CHECK_ARG(d && w && x && s_stack && out && workspace, "up_wasp_chain_fwd: null argument");
  WcPlan pl;
  int rc = wc_plan(d, pl);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) UP_CHECK_ARG(w->aspp[i] && w->shift[i], "up_wasp_chain_fwd: missing stage %d weights", i);
  UP_CHECK_ARG(w->conv1 && w->shift1 && w->gap_t && w->shift_gap && w->conv1_pool_t, "up_wasp_chain_fwd: missing weights");
  UP_CHECK_ARG(workspace_bytes >= up_wasp_chain_workspace_bytes(d), "up_wasp_chain_fwd: workspace too small");
  UP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(s_stack) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "up_wasp_chain_fwd: alignment");
  DeviceInfo* di = device_info();
  if (!di) return UP_ERR_CUDA;
  const int fmt = fmt_of_dtype(d->dtype);
  if (!di->chain_attr) {
    rc = check_cuda(cudaFuncSetAttribute(wasp_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(di->max_smem)),
                    "cudaFuncSetAttribute(wasp chain)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(wasp_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(di->max_smem)),
                    "cudaFuncSetAttribute(wasp chain bf16)");
    if (rc) return rc;
    di->chain_attr = true;
  }
  WcParams p{};
  p.N = d->n;
  p.H = d->h;
  p.W = d->w;
  p.cin = d->cin;
  p.bn = pl.bn;
  p.bh = pl.bh;
  p.bw = pl.bw;
  p.tiles_h = pl.tiles_h;
  p.tiles_w = pl.tiles_w;
  p.tiles_n = pl.tiles_n;
  p.c_per = pl.c_per;
  // co-resident CTA pairs: every tile of a wave's image groups must be resident at once (they wait for each other)
  int max_clusters = di->sm_count / 2;
  {
    int* cached = di->max_clusters;
    if (cached[0] == 0) {     // slot 0: this kernel
      cudaLaunchConfig_t occ{};
      occ.gridDim = dim3(di->sm_count / 2 * 2);
      occ.blockDim = dim3(kWcThreads);
      occ.dynamicSmemBytes = di->max_smem;
      cudaLaunchAttribute oa[1];
      oa[0].id = cudaLaunchAttributeClusterDimension;
      oa[0].val.clusterDim.x = 2;
      oa[0].val.clusterDim.y = 1;
      oa[0].val.clusterDim.z = 1;
      occ.attrs = oa;
      occ.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, wasp_chain_kernel<0>, &occ) == cudaSuccess && nc > 0) cached[0] = nc;
      else {
        (void)cudaGetLastError();
        cached[0] = di->sm_count / 2;
      }
    }
    if (cached[0] < max_clusters) max_clusters = cached[0];
  }
  if (const char* e = getenv("UP_CHAIN_MAX_CLUSTERS")) {   // tests: force several waves
    const int v = atoi(e);
    if (v >= 1 && v < max_clusters) max_clusters = v;
  }
  UP_CHECK_ARG(pl.per <= max_clusters, "up_wasp_chain_fwd: %d tiles per image exceed the %d co-resident CTA pairs", pl.per,
               max_clusters);
  const int pairs = pl.tiles_n / 2;
  p.pairs_per_wave = max_clusters / pl.per;
  if (p.pairs_per_wave > pairs) p.pairs_per_wave = pairs;
  p.waves = (pairs + p.pairs_per_wave - 1) / p.pairs_per_wave;
  const size_t pool_bytes = (2u * kWcMaxCin + 2u * kWcCout + 2u * kWcCout + 8u * kWcCout) * 4u;
  const size_t fixed = 1024 + kWcGroups * kWcBuf + pool_bytes + 8 * (3 * kWcMaxSlots + 4 + 3 * kWcGroups) + 16;
  int slots = static_cast<int>((di->max_smem - fixed) / kWcSlotBytes);
  if (slots > kWcMaxSlots) slots = kWcMaxSlots;
  if (const char* e = getenv("UP_CHAIN_SLOTS")) {
    const int v = atoi(e);
    if (v >= 2 && v < slots) slots = v;
  }
  UP_CHECK_ARG(slots >= 2, "up_wasp_chain_fwd: not enough shared memory");
  p.slots = slots;
  // pooling sums from the ring need every slot to see the same number of aspp1 k-blocks per wave
  // (measured slower than direct loads - the slot turnaround grows by the read: off unless UP_CHAIN_GAP_SMEM=1)
  p.gap_smem = 0;
  if (const char* e = getenv("UP_CHAIN_GAP_SMEM")) p.gap_smem = (e[0] == '1' && (d->cin / 64) % slots == 0) ? 1 : 0;
  p.l2_prefetch = 0;      // measured: no gain (x is DRAM-bound either way), slightly slower
  if (const char* e = getenv("UP_CHAIN_L2_PREFETCH")) p.l2_prefetch = e[0] == '1' ? 1 : 0;
  p.idesc = make_idesc_f16(static_cast<uint32_t>(fmt), 256u, 256u);
  p.st[0] = WcStage{1, 1, d->cin / 64, w->shift[0]};
  for (int i = 1; i < 4; ++i) p.st[i] = WcStage{3, d->dil[i - 1], kWcCout / 64, w->shift[i]};
  p.x = static_cast<const uint16_t*>(x);
  p.wg_t = static_cast<const uint16_t*>(w->gap_t);
  p.shift_gap = w->shift_gap;
  p.w5_t = static_cast<const uint16_t*>(w->conv1_pool_t);
  p.shift1 = w->shift1;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const size_t ctr_bytes = wc_align(static_cast<size_t>(pl.tiles_n * kWcCtrStride + 1) * 4);
  const size_t g2_bytes = wc_align(static_cast<size_t>(d->n) * kWcCout * 4);
  p.counters = reinterpret_cast<unsigned int*>(ws);
  p.g2 = reinterpret_cast<float*>(ws + ctr_bytes);
  p.gsum = reinterpret_cast<float*>(ws + ctr_bytes + g2_bytes);
  p.inv_hw = 1.0f / static_cast<float>(d->h * d->w);
  p.dbg = nullptr;
  if (getenv("UP_DEBUG_TIMING")) {
    if (!g_wc_dbg) cudaMalloc(&g_wc_dbg, 160 * 32 * sizeof(unsigned long long));
    cudaMemsetAsync(g_wc_dbg, 0, 160 * 32 * sizeof(unsigned long long), static_cast<cudaStream_t>(stream));
    p.dbg = g_wc_dbg;
  }

  CUtensorMap tmX, tmS, tmO, tmB[4], tmBc, tmXp;
  const uint32_t abox[5] = {64u, static_cast<uint32_t>(pl.bw), 1u, static_cast<uint32_t>(pl.bh),
                            static_cast<uint32_t>(pl.bn)};
  rc = encode_act_map(&tmX, fmt, x, d->n, d->h, d->w, d->cin, 1, abox, 128, "wasp.x");
  if (rc) return rc;
  {
    // prefetch view of x: 256-channel (512-byte) rows, no swizzle
    const uint32_t pbox[5] = {256u, static_cast<uint32_t>(pl.bw), 1u, static_cast<uint32_t>(pl.bh),
                              static_cast<uint32_t>(pl.bn)};
    rc = encode_act_map(&tmXp, fmt, x, d->n, d->h, d->w, d->cin, 1, pbox, 0, "wasp.x.prefetch");
    if (rc) return rc;
  }
  rc = encode_act_map(&tmS, fmt, s_stack, 4 * d->n, d->h, d->w, kWcCout, 1, abox, 128, "wasp.stack");
  if (rc) return rc;
  rc = encode_act_map(&tmO, fmt, out, d->n, d->h, d->w, kWcCout, 1, abox, 128, "wasp.out");
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) {
    const int cin_s = i == 0 ? d->cin : kWcCout;
    const int taps = i == 0 ? 1 : 9;
    const uint64_t dims[2] = {static_cast<uint64_t>(cin_s), static_cast<uint64_t>(taps) * kWcCout};
    const uint64_t st[1] = {static_cast<uint64_t>(cin_s) * 2};
    const uint32_t box[2] = {64u, kWcCout / 2};
    rc = encode_map(&tmB[i], fmt, 2, w->aspp[i], dims, st, box, 128, "wasp.filter");
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->conv1_cin), kWcCout};
    const uint64_t st[1] = {static_cast<uint64_t>(d->conv1_cin) * 2};
    const uint32_t box[2] = {64u, kWcCout / 2};
    rc = encode_map(&tmBc, fmt, 2, w->conv1, dims, st, box, 128, "wasp.conv1");
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * p.pairs_per_wave * pl.per);
  cfg.blockDim = dim3(kWcThreads);
  cfg.dynamicSmemBytes = fixed + static_cast<size_t>(slots) * kWcSlotBytes;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  rc = check_cuda(fmt == 0 ? cudaLaunchKernelEx(&cfg, wasp_chain_kernel<0>, tmX, tmS, tmO, tmB[0], tmB[1], tmB[2], tmB[3], tmBc, tmXp, p)
                           : cudaLaunchKernelEx(&cfg, wasp_chain_kernel<1>, tmX, tmS, tmO, tmB[0], tmB[1], tmB[2], tmB[3], tmBc, tmXp, p),
                  "wasp_chain_kernel launch");
  return rc;
}
