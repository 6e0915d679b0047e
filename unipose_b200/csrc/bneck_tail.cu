// The second half of a ResNet bottleneck - 3x3 conv (+BN+ReLU) -> 1x1 expansion (+BN) + residual + ReLU
// (Bottleneck.forward, model/modules/backbone/resnet.py:28-41) - as ONE kernel per block for the bandwidth-bound
// high-resolution layers (layer1: planes 64 at H/4, layer2: planes 128 at H/8).  The 3x3 output t2 never leaves the
// SM: its epilogue writes it (16-bit, 128B-swizzled) into shared memory, where it is the A operand of the expansion.
// Per 128-pixel tile:
//   conv2   acc2[128 x P]   = sum over taps / 64-channel chunks of t1 (TMA, out-of-image rows zero-filled) x W2
//   epi A   t2 = ReLU(acc2 + shift2) -> staging set T
//   conv3   acc3[128 x 256] = T x W3[n-tile]  (+ residual tile x I: identity MMAs, exact)     for each 256-channel N-tile
//   epi B   out = ReLU(acc3 + shift3) -> staging set O -> TMA store
// Persistent CTAs (one per SM, no clusters: these layers are HBM-bound, halving the weight traffic buys nothing) walk
// the tiles; the MMA issuer runs conv2 of tile i+1 BEFORE conv3 of tile i (acc2 is double buffered), so the tensor
// pipe works on the next tile while the epilogue warps stage t2, and the 256-column output epilogue of tile i overlaps
// conv2 of tile i+2.  Bytes per block: t1 (+halo) + residual + output - the t2 round trip (2 x N*H*W*P*2 B) and one
// launch disappear.
//
// Projection mode (the first block of a stage, stride 1: layer1 block 0): the shortcut is downsample(x) = Wd x (1x1 conv
// + BN, resnet.py:36-37) - instead of a separate launch that writes N*H*W*4P*2 bytes which the tail reads back, the x
// tile [128 x Cx] and the Wd N-tile ride through the ring and Wd x is accumulated straight into acc3 (replaces the
// identity MMAs); the epilogue adds both shifts.
//
// Next-conv1 mode (planes 64, `next_planes` = 64 / 128): the FOLLOWING bottleneck's conv1 (1x1, 4P -> N1, + BN + ReLU,
// resnet.py:25-27) is a per-pixel GEMM over exactly the tile this kernel has just produced, so it runs here: every
// 64-channel output group, while it sits in its staging buffer for the TMA store, is also the A operand of
// D1[128 x N1] += O_g x W1n[:, g]  (TMEM columns [128, 128 + N1), filter chunks through the ring); after the fourth group
// epilogue C turns D1 into the next block's t1 tile and stores it.  That block then starts at its 3x3: its conv1 launch
// and the re-read of this block's output (151 MB per layer1 block at 384^2 x 32) disappear.
//
// Warp roles (384 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc + store thread, 3 idle, 4..11 epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "up_internal.h"
#include "up_ptx.cuh"

namespace up {

constexpr int kBtThreads = 384;
constexpr int kBtEpiWarp0 = 4;
constexpr int kBtEpiThreads = 256;
constexpr int kBtMaxSlots = 4;
constexpr uint32_t kBtABytes = 16384;
constexpr uint32_t kBtSlotBytes = 32768;   // a filter N-tile (256 rows x 64 ch) or two 64-channel activation chunks;
                                           // also the ring slot size unless the tall conv2 boxes need more
constexpr uint32_t kBtBuf = 16384;

struct BtParams {
  int N, H, W;
  int bn, bh, bw;
  int tiles_h, tiles_w, tiles_n, total_tiles;
  int P;                 // planes: 64 or 128
  int pchunks;           // P / 64
  int ntiles;            // 4P / 256 output N-tiles
  int dil;
  int slots;
  uint32_t idesc2;       // M128 x N=P
  uint32_t idesc3;       // M128 x N=256
  uint32_t idesc_res;    // M128 x N=64
  const float* shift2;   // [P]
  const float* shift3;   // [4P]
  int obufs;             // output staging buffers of 64 channels (2: leaves room for a 4th ring slot; 4: one whole N-tile)
  int xchunks;           // projection mode: Cx / 64 input-channel chunks of the block input x (0 = identity shortcut)
  const float* shiftd;   // projection mode: [4P] shift of the downsample BatchNorm
  // Filter-row reuse for conv2 (the conv kernel's "tall" boxes): single-image tiles with bw % 8 == 0 fetch ONE
  // activation box of bh + 2*dil rows per (kw, chunk); filter row r reads it r*dil*bw pixel rows further down (a
  // multiple of the 1024-byte swizzle atom) against its own filter tile - 3 slots of (tall box + 3 filter tiles)
  // instead of 9 of (16 KB + 1 filter tile): the nine halo taps no longer re-read the tile from L2
  int n1;                  // next-conv1 mode: output channels of the following block's conv1 (0 = off, 64, 128)
  int n1_cps;              //   its 64-channel K-chunks [n1 x 64] per ring slot (32 KB): 4 or 2
  uint32_t idesc1;         //   M128 x N=n1
  const float* shift1n;    //   [n1] shift of the following block's bn1
  int l2pf;                // bulk L2 prefetch of the NEXT tile's shortcut operand (UP_TAIL_L2PF)
  int tall;
  uint32_t slot_bytes;     // ring slot stride
  uint32_t tall_a_bytes;   // (bh + 2*dil) * bw * 128
  uint32_t tall_a_step;    // dil * bw * 128 >> 4: descriptor step between filter rows
};

struct BtTile {
  int n0, h0, w0;
};
__device__ __forceinline__ BtTile bt_tile(const BtParams& p, int t) {
  BtTile r;
  r.w0 = (t % p.tiles_w) * p.bw;
  t /= p.tiles_w;
  r.h0 = (t % p.tiles_h) * p.bh;
  r.n0 = (t / p.tiles_h) * p.bn;
  return r;
}
__device__ __forceinline__ void bt_taps(int dil, int x0, int ext, int limit, int& lo, int& hi) {
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
template <int kFmt>
__device__ __forceinline__ uint32_t bt_pack2_relu(float lo_elem, float hi_elem) {
  uint32_t d;
  if constexpr (kFmt == 0) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}

template <int kFmt>
__global__ void __launch_bounds__(kBtThreads, 1)
    bneck_tail_kernel(const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmR,
                      const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW2,
                      const __grid_constant__ CUtensorMap tmW3, const __grid_constant__ CUtensorMap tmWd,
                      const __grid_constant__ CUtensorMap tmT1t, const __grid_constant__ CUtensorMap tmW1n,
                      const __grid_constant__ CUtensorMap tmT1n, const BtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stgT = smem_base + p.slots * p.slot_bytes;     // P/64 buffers: t2 tile
  const uint32_t stgO = stgT + p.pchunks * kBtBuf;               // obufs buffers: the stream of 64-channel output groups
  const uint32_t ident = stgO + p.obufs * kBtBuf;                // 64 x 64 identity (K-major, 128B swizzle), 8 KB
  const uint32_t bars = ident + 8192u;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kBtMaxSlots + s); };
  const uint32_t b0 = bars + 8u * (2 * kBtMaxSlots);
  auto full2_bar = [&](int a) { return b0 + 8u * a; };            // [2] conv2 accumulators
  auto empty2_bar = [&](int a) { return b0 + 8u * (2 + a); };
  const uint32_t full3_bar = b0 + 8u * 4, empty3_bar = b0 + 8u * 5;
  auto availT = [&](int g) { return b0 + 8u * (6 + g); };         // [2] MMA commit -> epilogue may refill
  auto s2readyT = [&](int g) { return b0 + 8u * (8 + g); };       // [2] epilogue -> MMA
  auto availO = [&](int g) { return b0 + 8u * (10 + g); };        // [4] store drained
  auto readyO = [&](int g) { return b0 + 8u * (14 + g); };        // [4] epilogue -> store thread
  const uint32_t full1_bar = b0 + 8u * 18, empty1_bar = b0 + 8u * 19;   // next-conv1 accumulator D1
  const uint32_t tmem_slot = b0 + 8u * 20;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int first = blockIdx.x, step = gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmT1);
    tma_prefetch_desc(&tmT1t);
    if (p.n1) {
      tma_prefetch_desc(&tmW1n);
      tma_prefetch_desc(&tmT1n);
    }
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW3);
    tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.slots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(full2_bar(a), 1);
      mbar_init(empty2_bar(a), kBtEpiThreads / 32);
      mbar_init(availT(a), 1);
      mbar_init(s2readyT(a), kBtEpiThreads / 32);
    }
    mbar_init(full3_bar, 1);
    mbar_init(empty3_bar, kBtEpiThreads / 32);
    for (int g = 0; g < 4; ++g) {
      // next-conv1 mode: a buffer is free again once the TMA store AND the D1 MMAs have read it (two arrivals)
      mbar_init(availO(g), p.n1 ? 2 : 1);
      mbar_init(readyO(g), kBtEpiThreads / 32);
    }
    mbar_init(full1_bar, 1);
    mbar_init(empty1_bar, kBtEpiThreads / 32);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  {
    // K-major, 128B-swizzled 64 x 64 identity: row n holds a single 1.0 at k = n
    const uint32_t one = kFmt == 1 ? 0x3F80u : 0x3C00u;
    for (uint32_t i = threadIdx.x; i < 8192u / 16u; i += blockDim.x) {
      const uint32_t n = i >> 3, chunk = i & 7u;
      const uint32_t src_chunk = chunk ^ (n & 7u);
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (src_chunk == (n >> 3)) {
        const uint32_t e = n & 7u;
        w[e >> 1] = one << ((e & 1u) * 16u);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ident + i * 16u), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                   "r"(w[3])
                   : "memory");
    }
    fence_proxy_async_smem();
  }
  tcgen05_before_thread_sync();
  __syncthreads();
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: conv2 accumulators at [0, P) and [P, 2P); the 256-column output accumulator at [256, 512)
  const uint32_t tmem_acc3 = tmem_base + 256u;
  const uint32_t tmem_d1 = tmem_base + 128u;     // next-conv1 accumulator (planes 64 only: conv2 uses [0, 128))
  const int c1 = p.n1 >> 6;                      // 64-channel groups of the next block's t1 tile
  const int gpt = 4 * p.ntiles + c1;             // staging-buffer groups per tile in the output stream
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // Work order shared by producer and MMA issuer: conv2(tile 0), conv2(tile 1); then for i >= 0: conv3(tile i),
  // conv2(tile i+2) [if any], next-conv1(tile i) [n1 mode]
  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t slot = 0, par = 1;
    auto advance = [&]() {
      if (++slot == static_cast<uint32_t>(p.slots)) {
        slot = 0;
        par ^= 1u;
      }
    };
    auto conv2_loads = [&](int tile) {
      const BtTile t = bt_tile(p, tile);
      int kh_lo, kh_hi, kw_lo, kw_hi;
      bt_taps(p.dil, t.h0, p.bh, p.H, kh_lo, kh_hi);
      bt_taps(p.dil, t.w0, p.bw, p.W, kw_lo, kw_hi);
      const uint32_t bbytes = static_cast<uint32_t>(p.P) * 128u;
      if (p.tall) {
        for (int kw = kw_lo; kw <= kw_hi; ++kw)
          for (int c = 0; c < p.pchunks; ++c) {
            mbar_wait(empty_bar(slot), par, 16000000000LL);
            if (elect_one()) {
              const uint32_t dst = smem_base + slot * p.slot_bytes;
              mbar_arrive_expect_tx(full_bar(slot), p.tall_a_bytes + 3u * bbytes);
              tma_load_5d(&tmT1t, dst, full_bar(slot), c * 64, t.w0 + (kw - 1) * p.dil, 0, t.h0 - p.dil, t.n0);
              for (int r = 0; r < 3; ++r)
                tma_load_2d(&tmW2, dst + p.tall_a_bytes + r * bbytes, full_bar(slot), c * 64, (r * 3 + kw) * p.P);
            }
            __syncwarp();
            advance();
          }
        return;
      }
      for (int kh = kh_lo; kh <= kh_hi; ++kh)
        for (int kw = kw_lo; kw <= kw_hi; ++kw)
          for (int c = 0; c < p.pchunks; ++c) {
            mbar_wait(empty_bar(slot), par, 16000000000LL);
            if (elect_one()) {
              const uint32_t dst = smem_base + slot * p.slot_bytes;
              mbar_arrive_expect_tx(full_bar(slot), kBtABytes + bbytes);
              tma_load_5d(&tmT1, dst, full_bar(slot), c * 64, t.w0 + (kw - 1) * p.dil, 0, t.h0 + (kh - 1) * p.dil, t.n0);
              tma_load_2d(&tmW2, dst + kBtABytes, full_bar(slot), c * 64, (kh * 3 + kw) * p.P);
            }
            __syncwarp();
            advance();
          }
    };
    auto conv3_loads = [&](int tile) {
      const BtTile t = bt_tile(p, tile);
      for (int nt = 0; nt < p.ntiles; ++nt) {
        // the N-tile's filter: 256 rows x 64 channels per K-chunk = 32 KB = one whole slot
        for (int c = 0; c < p.pchunks; ++c) {
          mbar_wait(empty_bar(slot), par, 16000000000LL);
          if (elect_one()) {
            const uint32_t dst = smem_base + slot * p.slot_bytes;
            mbar_arrive_expect_tx(full_bar(slot), kBtSlotBytes);
            tma_load_2d(&tmW3, dst, full_bar(slot), c * 64, nt * 256);
          }
          __syncwarp();
          advance();
        }
        if (p.xchunks > 0) {
          // projection shortcut: per 64-channel chunk of x one slot with the x tile (A) and one with the Wd N-tile (B)
          for (int xc = 0; xc < p.xchunks; ++xc) {
            mbar_wait(empty_bar(slot), par, 16000000000LL);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar(slot), kBtABytes);
              tma_load_5d(&tmR, smem_base + slot * p.slot_bytes, full_bar(slot), xc * 64, t.w0, 0, t.h0, t.n0);
            }
            __syncwarp();
            advance();
            mbar_wait(empty_bar(slot), par, 16000000000LL);
            if (elect_one()) {
              mbar_arrive_expect_tx(full_bar(slot), kBtSlotBytes);
              tma_load_2d(&tmWd, smem_base + slot * p.slot_bytes, full_bar(slot), xc * 64, nt * 256);
            }
            __syncwarp();
            advance();
          }
          continue;
        }
        // the residual tile of this N-tile: 256 channels = two slots of two 64-channel chunks
        for (int r2 = 0; r2 < 2; ++r2) {
          mbar_wait(empty_bar(slot), par, 16000000000LL);
          if (elect_one()) {
            const uint32_t dst = smem_base + slot * p.slot_bytes;
            mbar_arrive_expect_tx(full_bar(slot), kBtSlotBytes);
            for (int cc = 0; cc < 2; ++cc)
              tma_load_5d(&tmR, dst + cc * kBtABytes, full_bar(slot), nt * 256 + (2 * r2 + cc) * 64, t.w0, 0, t.h0, t.n0);
          }
          __syncwarp();
          advance();
        }
      }
    };
    auto conv1n_loads = [&]() {
      // the following block's conv1 filter, 64-channel K-chunks [n1 rows x 64] side by side, 32 KB per slot
      const uint32_t cbytes = static_cast<uint32_t>(p.n1) * 128u;
      for (int g0 = 0; g0 < 4; g0 += p.n1_cps) {
        mbar_wait(empty_bar(slot), par, 16000000000LL);
        if (elect_one()) {
          const uint32_t dst = smem_base + slot * p.slot_bytes;
          mbar_arrive_expect_tx(full_bar(slot), kBtSlotBytes);
          for (int j = 0; j < p.n1_cps; ++j) tma_load_2d(&tmW1n, dst + j * cbytes, full_bar(slot), (g0 + j) * 64, 0);
        }
        __syncwarp();
        advance();
      }
    };
    // The shortcut operand (residual tile / projection input) is the kernel's big DRAM stream and enters the ring only
    // right before its MMAs: a bulk L2 prefetch one tile ahead turns its DRAM latency into L2 latency
    auto prefetch_shortcut = [&](int tile) {
      const BtTile t = bt_tile(p, tile);
      if (elect_one()) {
        const int nch = p.xchunks > 0 ? p.xchunks : 4 * p.ntiles;
        for (int c = 0; c < nch; ++c)
          asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmR)),
                       "r"(c * 64), "r"(t.w0), "r"(0), "r"(t.h0), "r"(t.n0)
                       : "memory");
      }
      __syncwarp();
    };
    if (first < p.total_tiles) conv2_loads(first);
    if (first + step < p.total_tiles) conv2_loads(first + step);
    for (int tile = first; tile < p.total_tiles; tile += step) {
      if (p.l2pf && tile + step < p.total_tiles) prefetch_shortcut(tile + step);
      conv3_loads(tile);
      if (tile + 2 * step < p.total_tiles) conv2_loads(tile + 2 * step);
      if (p.n1) conv1n_loads();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t slot = 0, phase = 0;
    auto advance = [&]() {
      if (++slot == static_cast<uint32_t>(p.slots)) {
        slot = 0;
        phase ^= 1u;
      }
    };
    const uint64_t adesc0 = make_smem_desc_kmajor(smem_base, 128);
    const uint64_t tdesc0 = make_smem_desc_kmajor(stgT, 128);
    const uint64_t identdesc = make_smem_desc_kmajor(ident, 128);
    const uint32_t slot_step = p.slot_bytes >> 4;
    uint32_t n2[2] = {0u, 0u};      // uses of the conv2 accumulators
    uint32_t n3 = 0;                // uses of the output accumulator
    uint32_t nT = 0;                // t2 tiles consumed (s2readyT parity)
    auto conv2_mma = [&](int tile, int it) {
      const BtTile t = bt_tile(p, tile);
      int kh_lo, kh_hi, kw_lo, kw_hi;
      bt_taps(p.dil, t.h0, p.bh, p.H, kh_lo, kh_hi);
      bt_taps(p.dil, t.w0, p.bw, p.W, kw_lo, kw_hi);
      const int nkb = (p.tall ? 1 : (kh_hi - kh_lo + 1)) * (kw_hi - kw_lo + 1) * p.pchunks;
      const int a = it & 1;
      mbar_wait(empty2_bar(a), (n2[a] & 1u) ^ 1u);
      ++n2[a];
      tcgen05_after_thread_sync();
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(a * p.P);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(slot), phase);
        tcgen05_after_thread_sync();
        if (elect_one()) {
          const uint64_t ad = adesc0 + static_cast<uint64_t>(slot_step * slot);
          if (p.tall) {
            // three filter rows from ONE activation box (all three always: out-of-image rows are zero-filled)
            const uint64_t bd0 = ad + static_cast<uint64_t>(p.tall_a_bytes >> 4);
            for (int r = 0; r < 3; ++r) {
              const uint64_t ar = ad + static_cast<uint64_t>(p.tall_a_step * r);
              const uint64_t br = bd0 + static_cast<uint64_t>((static_cast<uint32_t>(p.P) * 128u >> 4) * r);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tacc, ar + 2u * k, br + 2u * k, p.idesc2, (kb | r | k) ? 1u : 0u);
            }
          } else {
            const uint64_t bd = ad + static_cast<uint64_t>(kBtABytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tacc, ad + 2u * k, bd + 2u * k, p.idesc2, (kb | k) ? 1u : 0u);
          }
          umma_commit(empty_bar(slot));
          if (kb == nkb - 1) umma_commit(full2_bar(a));
        }
        __syncwarp();
        advance();
      }
    };
    auto conv3_mma = [&]() {
      for (int g = 0; g < p.pchunks; ++g) mbar_wait(s2readyT(g), nT & 1u);     // t2 of this tile is staged
      for (int nt = 0; nt < p.ntiles; ++nt) {
        mbar_wait(empty3_bar, (n3 & 1u) ^ 1u);
        ++n3;
        tcgen05_after_thread_sync();
        for (int c = 0; c < p.pchunks; ++c) {
          mbar_wait(full_bar(slot), phase);
          tcgen05_after_thread_sync();
          if (elect_one()) {
            const uint64_t ad = tdesc0 + static_cast<uint64_t>((kBtBuf >> 4) * c);
            const uint64_t bd = adesc0 + static_cast<uint64_t>(slot_step * slot);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem_acc3, ad + 2u * k, bd + 2u * k, p.idesc3, (c | k) ? 1u : 0u);
            umma_commit(empty_bar(slot));
            if (nt == p.ntiles - 1 && c == p.pchunks - 1)
              for (int g = 0; g < p.pchunks; ++g) umma_commit(availT(g));          // t2 consumed
          }
          __syncwarp();
          advance();
        }
        if (p.xchunks > 0) {
          for (int xc = 0; xc < p.xchunks; ++xc) {
            const uint32_t slot_a = slot;
            mbar_wait(full_bar(slot), phase);
            advance();
            mbar_wait(full_bar(slot), phase);
            tcgen05_after_thread_sync();
            if (elect_one()) {
              const uint64_t ad = adesc0 + static_cast<uint64_t>(slot_step * slot_a);
              const uint64_t bd = adesc0 + static_cast<uint64_t>(slot_step * slot);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(tmem_acc3, ad + 2u * k, bd + 2u * k, p.idesc3, 1u);
              umma_commit(empty_bar(slot_a));
              umma_commit(empty_bar(slot));
              if (xc == p.xchunks - 1) umma_commit(full3_bar);
            }
            __syncwarp();
            advance();
          }
          continue;
        }
        for (int r2 = 0; r2 < 2; ++r2) {
          mbar_wait(full_bar(slot), phase);
          tcgen05_after_thread_sync();
          if (elect_one()) {
            for (int cc = 0; cc < 2; ++cc) {
              // residual: D[:, g*64 .. g*64+63] += R chunk x I   (exact: products with 1.0)
              const uint64_t rd = adesc0 + static_cast<uint64_t>(slot_step * slot) + static_cast<uint64_t>((kBtABytes >> 4) * cc);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_acc3 + static_cast<uint32_t>(2 * r2 + cc) * 64u, rd + 2u * k, identdesc + 2u * k, p.idesc_res, 1u);
            }
            umma_commit(empty_bar(slot));
            if (r2 == 1) umma_commit(full3_bar);
          }
          __syncwarp();
          advance();
        }
      }
      ++nT;
    };
    const uint64_t odesc0 = make_smem_desc_kmajor(stgO, 128);
    uint32_t n1u = 0;               // uses of the next-conv1 accumulator
    auto conv1n_mma = [&](int it) {
      // D1 += O_g x W1n[:, g]: the A operand is the output group in its staging buffer (the store thread reads the
      // same bytes); the buffer returns to the epilogue when both readers are done (availO counts 2)
      const uint32_t nb = static_cast<uint32_t>(p.obufs);
      const uint32_t q0 = static_cast<uint32_t>(it) * static_cast<uint32_t>(gpt);
      const uint32_t cstep = (static_cast<uint32_t>(p.n1) * 128u) >> 4;
      mbar_wait(empty1_bar, (n1u & 1u) ^ 1u);
      ++n1u;
      tcgen05_after_thread_sync();
      for (int g = 0; g < 4; ++g) {
        const int j = g % p.n1_cps;
        if (j == 0) mbar_wait(full_bar(slot), phase);
        const uint32_t q = q0 + static_cast<uint32_t>(g);
        const uint32_t b = q % nb;
        mbar_wait(readyO(b), (q / nb) & 1u);
        tcgen05_after_thread_sync();
        if (elect_one()) {
          const uint64_t ad = odesc0 + static_cast<uint64_t>((kBtBuf >> 4) * b);
          const uint64_t bd = adesc0 + static_cast<uint64_t>(slot_step * slot) + static_cast<uint64_t>(cstep * j);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_d1, ad + 2u * k, bd + 2u * k, p.idesc1, (g | k) ? 1u : 0u);
          umma_commit(availO(b));
          if (j == p.n1_cps - 1) umma_commit(empty_bar(slot));
          if (g == 3) umma_commit(full1_bar);
        }
        __syncwarp();
        if (j == p.n1_cps - 1) advance();
      }
    };
    // conv2 runs two tiles ahead of the output epilogue; the D1 MMAs of tile i (which wait for the epilogue's output
    // groups) come after conv2 of tile i+2, so the ring already holds conv3(i+1)'s operands while the issuer waits
    int it = 0;
    if (first < p.total_tiles) conv2_mma(first, 0);
    if (first + step < p.total_tiles) conv2_mma(first + step, 1);
    for (int tile = first; tile < p.total_tiles; tile += step, ++it) {
      conv3_mma();
      if (tile + 2 * step < p.total_tiles) conv2_mma(tile + 2 * step, it + 2);
      if (p.n1) conv1n_mma(it);
    }
  } else if (threadIdx.x == 64) {
    // ===================== store thread =====================
    // the 64-channel output groups form one stream q = 0, 1, ...: group q lives in buffer q % obufs; a buffer goes back
    // to the epilogue as soon as the store that read it has finished reading (one younger store may still be pending)
    uint32_t q = 0;
    bool prev_c = false;            // the previous group of the stream was a t1 group (next-conv1 mode)
    const uint32_t nb = static_cast<uint32_t>(p.obufs);
    for (int tile = first; tile < p.total_tiles; tile += step) {
      const BtTile t = bt_tile(p, tile);
      for (int nt = 0; nt < p.ntiles; ++nt) {
        for (int g = 0; g < 4; ++g, ++q) {
          const uint32_t b = q % nb;
          mbar_wait(readyO(b), (q / nb) & 1u);
          tma_store_5d(&tmY, stgO + b * kBtBuf, nt * 256 + g * 64, t.w0, 0, t.h0, t.n0);
          tma_store_commit();
          if (q > 0) {
            tma_store_wait_read<1>();
            mbar_arrive(availO((q - 1) % nb));
            if (prev_c) mbar_arrive(availO((q - 1) % nb));   // a t1 group has no MMA reader: second arrival
          }
          prev_c = false;
        }
      }
      // next-conv1 mode: the groups of the following block's t1 tile ride in the same buffer stream
      for (int g = 0; g < c1; ++g, ++q) {
        const uint32_t b = q % nb;
        mbar_wait(readyO(b), (q / nb) & 1u);
        tma_store_5d(&tmT1n, stgO + b * kBtBuf, g * 64, t.w0, 0, t.h0, t.n0);
        tma_store_commit();
        tma_store_wait_read<1>();                              // q > 0 here
        mbar_arrive(availO((q - 1) % nb));
        if (prev_c) mbar_arrive(availO((q - 1) % nb));
        prev_c = true;
      }
    }
    tma_store_wait_all<0>();
  } else if (warp >= kBtEpiWarp0) {
    // ===================== epilogue (8 warps) =====================
    const int ew = warp - kBtEpiWarp0;
    const int quarter = ew & 3;
    const int half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
    const uint32_t row7 = static_cast<uint32_t>(row) & 7u;
    const uint32_t tlane = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t nT = 0, qO = 0, n3 = 0, n1e = 0;   // qO: next group of the output stream (see the store thread)
    uint32_t n2[2] = {0u, 0u};
    // `groups` x 64 accumulator columns -> ReLU(acc + shift) -> 16-bit -> staging buffers
    auto epilogue = [&](uint32_t tmem_col0, int groups, const float* sh, const float* sh2, bool setO) {
      uint32_t r[32];
      const uint32_t taddr = tmem_col0 + tlane + static_cast<uint32_t>(half * 32);
      const uint32_t nb = static_cast<uint32_t>(p.obufs);
      tmem_ld_32x32b_x32(taddr, r);
      for (int g = 0; g < groups; ++g) {
        const float4* s4 = reinterpret_cast<const float4*>(sh + g * 64 + half * 32);
        float v[32];
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 h4 = __ldg(s4 + j4);
          v[4 * j4 + 0] = __uint_as_float(r[4 * j4 + 0]) + h4.x;
          v[4 * j4 + 1] = __uint_as_float(r[4 * j4 + 1]) + h4.y;
          v[4 * j4 + 2] = __uint_as_float(r[4 * j4 + 2]) + h4.z;
          v[4 * j4 + 3] = __uint_as_float(r[4 * j4 + 3]) + h4.w;
        }
        if (sh2) {
          const float4* d4 = reinterpret_cast<const float4*>(sh2 + g * 64 + half * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 h4 = __ldg(d4 + j4);
            v[4 * j4 + 0] += h4.x;
            v[4 * j4 + 1] += h4.y;
            v[4 * j4 + 2] += h4.z;
            v[4 * j4 + 3] += h4.w;
          }
        }
        if (g + 1 < groups) tmem_ld_32x32b_x32(taddr + (g + 1) * 64, r);
        uint32_t w[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) w[e] = bt_pack2_relu<kFmt>(v[2 * e], v[2 * e + 1]);
        const uint32_t q = qO + static_cast<uint32_t>(g);
        const uint32_t b = setO ? q % nb : static_cast<uint32_t>(g);
        mbar_wait(setO ? availO(b) : availT(g), ((setO ? q / nb : nT) & 1u) ^ 1u);
        const uint32_t rowaddr = (setO ? stgO : stgT) + b * kBtBuf + rowoff;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t addr = rowaddr + (((static_cast<uint32_t>(half) * 4u + c4) ^ row7) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * c4]), "r"(w[4 * c4 + 1]),
                       "r"(w[4 * c4 + 2]), "r"(w[4 * c4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(setO ? readyO(b) : s2readyT(g));
      }
      if (setO) qO += static_cast<uint32_t>(groups); else ++nT;
    };
    auto epilogue_a = [&](int it) {      // t2 of the tile with sequence number `it`
      const int a = it & 1;
      mbar_wait(full2_bar(a), n2[a] & 1u);
      ++n2[a];
      tcgen05_after_thread_sync();
      epilogue(tmem_base + static_cast<uint32_t>(a * p.P), p.pchunks, p.shift2, nullptr, false);
      tcgen05_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty2_bar(a));
    };
    int it = 0;
    if (first < p.total_tiles) epilogue_a(0);      // ---- epilogue A: t2, always one tile ahead of epilogue B ----
    for (int tile = first; tile < p.total_tiles; tile += step, ++it) {
      // ---- epilogue B: the block's output, 256 channels per N-tile ----
      for (int nt = 0; nt < p.ntiles; ++nt) {
        mbar_wait(full3_bar, n3 & 1u);
        ++n3;
        tcgen05_after_thread_sync();
        epilogue(tmem_acc3, 4, p.shift3 + nt * 256, p.xchunks > 0 ? p.shiftd + nt * 256 : nullptr, true);
        tcgen05_before_thread_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty3_bar);
      }
      // t2 of the next tile (its conv2 finished long ago): the issuer can go on to that tile's conv3
      if (tile + step < p.total_tiles) epilogue_a(it + 1);
      if (p.n1) {
        // ---- epilogue C: the following block's t1 tile = ReLU(D1 + shift1n) ----
        mbar_wait(full1_bar, n1e & 1u);
        ++n1e;
        tcgen05_after_thread_sync();
        epilogue(tmem_d1, c1, p.shift1n, nullptr, true);
        tcgen05_before_thread_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty1_bar);
      }
    }
  }

  tcgen05_before_thread_sync();
  __syncthreads();
  if (warp == 2) {
    tcgen05_after_thread_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace up
#include "up_conv_host.h"

using namespace up;

static int bt_check(const UpBneckTailDesc* d) {
  if (!d) return fail(UP_ERR_INVALID, "up_bneck_tail: null descriptor");
  if (d->dtype != UP_FP16 && d->dtype != UP_BF16) return fail(UP_ERR_UNSUPPORTED, "up_bneck_tail: fp16 / bf16 only");
  if (d->planes != 64 && d->planes != 128) return fail(UP_ERR_UNSUPPORTED, "up_bneck_tail: planes must be 64 or 128 (got %d)", d->planes);
  if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->dil < 1) return fail(UP_ERR_INVALID, "up_bneck_tail: bad dims");
  if (d->proj_cin < 0 || d->proj_cin % 64 != 0 || d->proj_cin > 1024)
    return fail(UP_ERR_UNSUPPORTED, "up_bneck_tail: proj_cin must be 0 or a multiple of 64 <= 1024 (got %d)", d->proj_cin);
  if (d->next_planes != 0 && (d->planes != 64 || (d->next_planes != 64 && d->next_planes != 128)))
    return fail(UP_ERR_UNSUPPORTED, "up_bneck_tail: next_planes (%d) needs planes 64 and must be 64 or 128", d->next_planes);
  return 0;
}

extern "C" int up_bneck_tail_supported(const UpBneckTailDesc* d) { return bt_check(d); }

extern "C" int up_bneck_tail_fwd(const UpBneckTailDesc* d, const void* t1, const void* w2, const float* shift2,
                                 const void* w3, const float* shift3, const void* residual, const void* wd,
                                 const float* shiftd, void* y, const void* w1n, const float* shift1n, void* t1n,
                                 void* stream) {
  UP_CHECK_ARG(d && t1 && w2 && shift2 && w3 && shift3 && residual && y, "up_bneck_tail_fwd: null argument");
  int rc = bt_check(d);
  if (rc) return rc;
  UP_CHECK_ARG((d->proj_cin > 0) == (wd != nullptr) && (wd != nullptr) == (shiftd != nullptr),
               "up_bneck_tail_fwd: wd / shiftd go with proj_cin > 0");
  UP_CHECK_ARG((d->next_planes > 0) == (w1n != nullptr) && (w1n != nullptr) == (shift1n != nullptr) &&
                   (w1n != nullptr) == (t1n != nullptr),
               "up_bneck_tail_fwd: w1n / shift1n / t1n go with next_planes > 0");
  DeviceInfo* di = device_info();
  if (!di) return UP_ERR_CUDA;
  if (!di->tail_attr) {
    rc = check_cuda(cudaFuncSetAttribute(bneck_tail_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(di->max_smem)), "cudaFuncSetAttribute(bneck tail)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(bneck_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(di->max_smem)), "cudaFuncSetAttribute(bneck tail bf16)");
    if (rc) return rc;
    di->tail_attr = true;
  }
  const int fmt = fmt_of_dtype(d->dtype);
  BtParams p{};
  p.N = d->n;
  p.H = d->h;
  p.W = d->w;
  pick_tile(d->n, d->h, d->w, p.bn, p.bh, p.bw);
  p.tiles_w = (d->w + p.bw - 1) / p.bw;
  p.tiles_h = (d->h + p.bh - 1) / p.bh;
  p.tiles_n = (d->n + p.bn - 1) / p.bn;
  p.total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.P = d->planes;
  p.pchunks = d->planes / 64;
  p.ntiles = (4 * d->planes) / 256;
  p.dil = d->dil;
  static const int obufs_env = []() {
    const char* e = getenv("UP_TAIL_OBUFS");      // tuning: 2 .. 4 output staging buffers (0 = default)
    return (e && e[0] >= '2' && e[0] <= '4') ? e[0] - '0' : 0;
  }();
  // planes 64: the t2 set is one buffer, the 16 KB it leaves buy a third output buffer (ncu: the epilogue warps spent
  // 21 % of their samples waiting for a free buffer with two, see profiles/ncu_r2_tail_conv1_*.txt); planes 128 keeps
  // two buffers and a fourth ring slot
  const int obufs = obufs_env ? obufs_env : (d->planes == 64 ? 3 : 2);
  p.obufs = obufs;
  const size_t fixed = 1024 + static_cast<size_t>(d->planes / 64 + obufs) * kBtBuf + 8192 + 8 * (2 * kBtMaxSlots + 20) + 16;
  p.tall = 0;
  p.slot_bytes = kBtSlotBytes;
  p.tall_a_bytes = static_cast<uint32_t>(p.bh + 2 * d->dil) * p.bw * 128u;
  p.tall_a_step = (static_cast<uint32_t>(d->dil) * p.bw * 128u) >> 4;
  {
    static const bool want = []() {
      const char* e = getenv("UP_TAIL_TALL");
      return !(e && e[0] == '0');
    }();
    // one tall box must be smaller than the two extra 16 KB boxes it replaces, and three slots must still fit
    const uint32_t slot = p.tall_a_bytes + 3u * static_cast<uint32_t>(d->planes) * 128u;
    if (want && p.bn == 1 && p.bw % 8 == 0 && p.bh + 2 * d->dil <= 256 && p.tall_a_bytes <= 2u * kBtABytes &&
        slot >= kBtSlotBytes && fixed + 3u * static_cast<size_t>(slot) <= di->max_smem) {
      p.tall = 1;
      p.slot_bytes = slot;
    }
  }
  int slots = static_cast<int>((di->max_smem - fixed) / p.slot_bytes);
  if (slots > kBtMaxSlots) slots = kBtMaxSlots;
  UP_CHECK_ARG(slots >= 2, "up_bneck_tail_fwd: not enough shared memory");
  p.slots = slots;
  p.idesc2 = make_idesc_f16(static_cast<uint32_t>(fmt), 128u, static_cast<uint32_t>(d->planes));
  p.idesc3 = make_idesc_f16(static_cast<uint32_t>(fmt), 128u, 256u);
  p.idesc_res = make_idesc_f16(static_cast<uint32_t>(fmt), 128u, 64u);
  p.shift2 = shift2;
  p.shift3 = shift3;
  p.xchunks = d->proj_cin / 64;
  p.shiftd = shiftd;
  static const int l2pf = []() {
    const char* e = getenv("UP_TAIL_L2PF");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  p.l2pf = l2pf;
  p.n1 = d->next_planes;
  p.n1_cps = p.n1 ? static_cast<int>(kBtSlotBytes / (static_cast<uint32_t>(p.n1) * 128u)) : 1;
  p.idesc1 = make_idesc_f16(static_cast<uint32_t>(fmt), 128u, static_cast<uint32_t>(p.n1 ? p.n1 : 64));
  p.shift1n = shift1n;
  CUtensorMap tmT1, tmR, tmY, tmW2, tmW3, tmWd, tmT1t, tmW1n, tmT1n;
  const uint32_t abox[5] = {64u, static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh), static_cast<uint32_t>(p.bn)};
  rc = encode_act_map(&tmT1, fmt, t1, d->n, d->h, d->w, d->planes, 1, abox, 128, "tail.t1");
  if (rc) return rc;
  tmT1t = tmT1;
  if (p.tall) {
    const uint32_t tbox[5] = {64u, static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh + 2 * d->dil), 1u};
    rc = encode_act_map(&tmT1t, fmt, t1, d->n, d->h, d->w, d->planes, 1, tbox, 128, "tail.t1 (tall box)");
    if (rc) return rc;
  }
  rc = encode_act_map(&tmR, fmt, residual, d->n, d->h, d->w, d->proj_cin > 0 ? d->proj_cin : 4 * d->planes, 1, abox, 128,
                      "tail.residual");
  if (rc) return rc;
  rc = encode_act_map(&tmY, fmt, y, d->n, d->h, d->w, 4 * d->planes, 1, abox, 128, "tail.y");
  if (rc) return rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->planes), static_cast<uint64_t>(9) * d->planes};
    const uint64_t st[1] = {static_cast<uint64_t>(d->planes) * 2};
    const uint32_t box[2] = {64u, static_cast<uint32_t>(d->planes)};
    rc = encode_map(&tmW2, fmt, 2, w2, dims, st, box, 128, "tail.w2");
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->planes), static_cast<uint64_t>(4) * d->planes};
    const uint64_t st[1] = {static_cast<uint64_t>(d->planes) * 2};
    const uint32_t box[2] = {64u, 256u};
    rc = encode_map(&tmW3, fmt, 2, w3, dims, st, box, 128, "tail.w3");
    if (rc) return rc;
  }
  tmWd = tmW3;
  if (d->proj_cin > 0) {
    const uint64_t dims[2] = {static_cast<uint64_t>(d->proj_cin), static_cast<uint64_t>(4) * d->planes};
    const uint64_t st[1] = {static_cast<uint64_t>(d->proj_cin) * 2};
    const uint32_t box[2] = {64u, 256u};
    rc = encode_map(&tmWd, fmt, 2, wd, dims, st, box, 128, "tail.wd");
    if (rc) return rc;
  }
  tmW1n = tmW3;
  tmT1n = tmY;
  if (p.n1) {
    const uint64_t dims[2] = {static_cast<uint64_t>(4) * d->planes, static_cast<uint64_t>(p.n1)};
    const uint64_t st[1] = {static_cast<uint64_t>(4) * d->planes * 2};
    const uint32_t box[2] = {64u, static_cast<uint32_t>(p.n1)};
    rc = encode_map(&tmW1n, fmt, 2, w1n, dims, st, box, 128, "tail.w1n");
    if (rc) return rc;
    rc = encode_act_map(&tmT1n, fmt, t1n, d->n, d->h, d->w, p.n1, 1, abox, 128, "tail.t1n");
    if (rc) return rc;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.total_tiles < di->sm_count ? p.total_tiles : di->sm_count);
  cfg.blockDim = dim3(kBtThreads);
  cfg.dynamicSmemBytes = fixed + static_cast<size_t>(slots) * p.slot_bytes;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  rc = check_cuda(fmt == 0 ? cudaLaunchKernelEx(&cfg, bneck_tail_kernel<0>, tmT1, tmR, tmY, tmW2, tmW3, tmWd, tmT1t, tmW1n, tmT1n, p)
                           : cudaLaunchKernelEx(&cfg, bneck_tail_kernel<1>, tmT1, tmR, tmY, tmW2, tmW3, tmWd, tmT1t, tmW1n, tmT1n, p),
                  "bneck_tail_kernel launch");
  return rc;
}
