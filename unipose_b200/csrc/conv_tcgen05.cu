// Implicit-GEMM convolution for sm_100a: TMA (im2col-free, one box per filter tap with the tap's
// dilated offset, OOB zero fill = padding) -> shared memory (128B / 32B swizzle) -> tcgen05.mma with
// fp32 accumulators in TMEM -> fused scale/shift(/residual)(/ReLU) epilogue -> TMA store (NHWC)
// or direct fp32 NCHW store.
//
// GEMM view:  M = output pixels (tile = bn x bh x bw = 128 pixels of one/several images),
//             N = output channels (tile = block_n in {32,64,128,256}),
//             K = taps x input channels (k-block = one tap x `ck` channels, ck in {16,64}).
//
// Persistent, warp-specialised CTA (384 threads, 1 CTA / SM).  By default two CTAs form a cluster and issue ONE
// tcgen05.mma.cta_group::2 of M = 256 per k-step (kPair): each CTA loads its own 128 activation rows and HALF of the
// weight tile, i.e. 2/3 of the L2->SM bytes per FLOP of two independent CTAs.
//   warps 0 and 3     : TMA producers (even / odd k-blocks).  Each warp walks the (tap, channel chunk, term) loop nest
//                       CONVERGED with warp-uniform values - increments only, no divisions - and one elect.sync lane
//                       issues the mbarrier expect_tx + the TMA loads.  Residual tiles ride in the same ring, up to
//                       (a_bytes + b_bytes) / 16 KB of them per slot.
//   warp 1            : tcgen05.mma issuer, converged as well (descriptors stay in uniform registers, the four MMAs
//                       of a k-block issue back to back); smem ring: full/empty mbarriers, TMEM double buffer:
//                       tfull/tempty.  The residual is added as D += R x I (identity B tile) inside the same pipe.
//   warp 2            : TMEM alloc/dealloc; lane 0 = epilogue DMA thread: TMA-stores each finished 64-channel group
//                       from its staging buffer (2..4 buffers, avail/ready mbarriers, one group of look-ahead)
//   warps 4..11       : epilogue math, templated on the storage format: tcgen05.ld of group g+1 is in flight while
//                       group g gets shift (+scale), cvt.rn[.relu].{f16,bf16}x2 and the swizzled st.shared.
//                       Two warps per TMEM lane quarter, each takes 32 of a group's 64 columns.
// Measured (profiles/conv_phase_timeline_r1_*.txt): a k-block costs ~0.28 us of unique activation bytes (16 KB at
// ~57 KB/us/SM from L2) + weight bytes at ~200 KB/us/SM (all SMs read the same tile); the tensor pipe needs 0.26 us.
//
// Filter taps whose whole input box lies outside the image contribute exact zeros and are skipped by producer and
// issuer alike (large-dilation WASP convs on small maps: wasp.py:47-49); the in-bounds taps of a tile always form a
// rectangle [kh_lo,kh_hi] x [kw_lo,kw_hi].
//
// UP_SPLIT ("fp32-grade") mode: activations and weights are bf16 hi+lo planes; every k-block is issued three times
// (hi*hi, lo*hi, hi*lo) into the same fp32 accumulator.
#include <cuda.h>
#include <stdlib.h>

#include "up_internal.h"
#include "up_ptx.cuh"

namespace up {

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define UP_STAMP(slot)                                                       \
  do {                                                                       \
    if (p.dbg) p.dbg[blockIdx.x * 16 + (slot)] = gtimer();                   \
  } while (0)

constexpr int kMaxStages = 8;
constexpr int kMaxBufs = 4;
constexpr int kTileM = 128;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kPlaneBytes = 16384;  // one 128-row x 128-byte staging plane

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
  int stages, nbuf;
  int pair;        // 1: the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 (M = 256) per k-step; each holds
                   //    its own 128 activation rows and HALF of the weight tile -> half the smem fill + operand reads
  int cluster;     // CTAs per cluster (1, 2 or 4): they share one weight tile per k-block via TMA multicast
  int res_terms;   // residual tiles (128 px x 64 ch, 16 KB) per 64-channel group (0 = none, 1, or 2 in split mode)
  int res_per_slot;  // how many of them share one ring slot (the slot is a_bytes + b_bytes wide)
  int wide;          // N = 512 tile (CTA pairs only): ONE accumulator of 512 TMEM columns, two N = 256 UMMAs per k-step
                     // over the same activation slot - the activation tile is fetched once per 512 output channels
  int proj_taps;     // UP_FLAG_PROJ: extra K - a second input (tensor map A1) projected by proj_taps more "taps" of the
  int proj_coff;     //   packed filter (cin columns each); proj_coff = first channel of that input's view
  int bsplit;        // experiment (UP_DEBUG_BSPLIT): fetch the weight tile with this many TMA instructions
  int tall;          // 3x3 stride-1 convs on single-image tiles: ONE activation box of bh + 2*dil rows per (kw, chunk)
                     // serves the three filter rows through row-offset descriptors (a_bytes = tall box, b_bytes = 3 taps)
  uint32_t tall_a_step;   // dil * bw * 128 bytes >> 4: descriptor step between filter rows
  uint32_t tall_b_bytes;  // bytes of ONE tap's weight tile in the slot (this CTA's half in pair mode)
  uint32_t idesc_res;
  uint32_t a_bytes, b_bytes, buf_bytes;
  uint32_t idesc;
  uint32_t tmem_cols;
  int flags, fmt, split;
  int cout_valid, out_c_total;
  int y_coff, r_coff;
  const float* scale;
  const float* shift;
  float* out_f32;
  unsigned long long* dbg;   // optional per-CTA phase timestamps (globaltimer ns): UP_DEBUG_TIMING=1
};

struct TileCoord {
  int n0, h0, w0, nt;
};

// Work item -> tile.  The CTAs of a cluster take the SAME (nt, tw, th) and consecutive image groups tn, so they
// walk identical k-block sequences (same in-bounds taps, same weight tiles) - the precondition for multicast.
__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int work, int crank) {
  TileCoord t;
  t.nt = work % p.n_tiles;
  int mt = work / p.n_tiles;
  int tw = mt % p.tiles_w;
  mt /= p.tiles_w;
  int th = mt % p.tiles_h;
  int tn = (mt / p.tiles_h) * p.cluster + crank;
  t.n0 = tn * p.bn;
  t.h0 = th * p.bh;
  t.w0 = tw * p.bw;
  return t;
}

// Offset of filter tap k along one axis in box coordinates (+ parity plane for stride 2).
__device__ __forceinline__ void tap_offset(const ConvKParams& p, int k, int pad, int& off, int& par) {
  int o = k * p.dil - pad;
  par = 0;
  if (p.stride == 2) {
    par = o & 1;
    o = (o - par) >> 1;
  }
  off = o;
}

// Inclusive range of taps along one axis whose box [x0+off, x0+off+ext) touches [0, limit).
__device__ __forceinline__ void tap_range(const ConvKParams& p, int taps, int pad, int x0, int ext, int limit, int& lo,
                                          int& hi) {
  lo = taps;
  hi = -1;
  for (int k = 0; k < taps; ++k) {
    int off, par;
    tap_offset(p, k, pad, off, par);
    const int c = x0 + off;
    if (c + ext > 0 && c < limit) {
      lo = min(lo, k);
      hi = max(hi, k);
    }
  }
}

// ---- epilogue of one 128 x block_n tile, NHWC 16-bit output ---------------------------------------------------
// kMode: 0 = fp16, 1 = bf16, 2 = bf16 hi/lo split.  cvt.*x2 packs two values per instruction and applies the ReLU.
template <int kMode, bool kRelu>
__device__ __forceinline__ uint32_t epi_pack2(float lo_elem, float hi_elem) {
  uint32_t d;
  if constexpr (kMode == 0) {
    if constexpr (kRelu) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  } else {
    if constexpr (kRelu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  }
  return d;
}

struct EpiCtx {
  const float* scale;    // first of this warp's 32 columns of group 0
  const float* shift;
  uint32_t taddr;        // TMEM address of the same columns (lane quarter in the upper half)
  uint32_t row_smem;     // staging buffer 0 + this thread's 128-byte row
  uint32_t buf_bytes, nbuf;
  uint32_t chunk0;       // first 16-byte chunk of the row this warp fills (0 or 4)
  uint32_t row7;         // row & 7: the 128B-swizzle XOR
  uint32_t avail0;       // avail[0] barrier; ready[b] = avail0 + 8 * (kMaxBufs + b)
  int lane;
  int groups;
};

// Groups of 64 columns.  The accumulator registers are dead once scale/shift have been applied, so the tcgen05.ld of
// group g+1 is issued right there and flies while group g is packed, written to the staging buffer and handed over.
template <int kMode, bool kRelu>
__device__ __forceinline__ void epi_tile(const EpiCtx& c, uint32_t& q) {
  uint32_t r[32];
  tmem_ld_32x32b_x32(c.taddr, r);
  for (int g = 0; g < c.groups; ++g, ++q) {
    const float4* sc = reinterpret_cast<const float4*>(c.scale + g * 64);
    const float4* sh = reinterpret_cast<const float4*>(c.shift + g * 64);
    float v[32];
    tmem_ld_wait();
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      const float4 s4 = __ldg(sc + j4);
      const float4 h4 = __ldg(sh + j4);
      v[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), s4.x, h4.x);
      v[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), s4.y, h4.y);
      v[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), s4.z, h4.z);
      v[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), s4.w, h4.w);
    }
    if (g + 1 < c.groups) tmem_ld_32x32b_x32(c.taddr + (g + 1) * 64, r);
    uint32_t w[16], wl[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      if constexpr (kMode == 2) {
        const float a = kRelu ? fmaxf(v[2 * e], 0.f) : v[2 * e];
        const float b = kRelu ? fmaxf(v[2 * e + 1], 0.f) : v[2 * e + 1];
        w[e] = epi_pack2<1, false>(a, b);
        wl[e] = epi_pack2<1, false>(a - __uint_as_float(w[e] << 16), b - __uint_as_float(w[e] & 0xFFFF0000u));
      } else {
        w[e] = epi_pack2<kMode, kRelu>(v[2 * e], v[2 * e + 1]);
      }
    }
    const uint32_t b = q % c.nbuf;
    // the staging buffer becomes ours (its previous TMA store has drained)
    mbar_wait(c.avail0 + 8u * b, (q / c.nbuf) & 1u);
    const uint32_t rowaddr = c.row_smem + b * c.buf_bytes;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const uint32_t addr = rowaddr + (((c.chunk0 + c4) ^ c.row7) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * c4]), "r"(w[4 * c4 + 1]),
                   "r"(w[4 * c4 + 2]), "r"(w[4 * c4 + 3])
                   : "memory");
      if constexpr (kMode == 2) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + kPlaneBytes), "r"(wl[4 * c4]),
                     "r"(wl[4 * c4 + 1]), "r"(wl[4 * c4 + 2]), "r"(wl[4 * c4 + 3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();   // make the generic-proxy writes visible to the TMA store
    __syncwarp();
    if (c.lane == 0) mbar_arrive(c.avail0 + 8u * (kMaxBufs + b));  // one arrival per warp (8) -> the DMA thread stores
  }
}

// kPair selects tcgen05 cta_group::2 everywhere: PTX requires ONE cta_group per kernel, hence two instantiations.
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
    conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                        const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmY0,
                        const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmR0,
                        const __grid_constant__ CUtensorMap tmR1, const ConvKParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atoms of TMA and the UMMA descriptors.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
  const uint32_t staging = smem_base + p.stages * stage_bytes;
  const uint32_t ident = staging + p.nbuf * p.buf_bytes;  // 64x64 identity B tile (residual add as an MMA), 8 KB
  const uint32_t bars = ident + (p.res_terms ? 8192u : 0u);
  // barriers (8 bytes each): full[8] empty[8] tfull[2] tempty[2] avail[4] ready[4]; then the TMEM base slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * kMaxStages + 2 + a); };
  auto avail_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 4 + b); };
  auto ready_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 4 + kMaxBufs + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 4 + 2 * kMaxBufs);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) UP_STAMP(0);   // kernel entry
  uint32_t crank = 0;
  if (p.cluster > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int first_work = blockIdx.x / p.cluster;
  const int work_step = gridDim.x / p.cluster;
  const int total_tiles = (p.tiles_n / p.cluster) * p.tiles_h * p.tiles_w * p.n_tiles;
  const bool nchw = (p.flags & UP_FLAG_OUT_NCHW_F32) != 0;
  const bool has_res = (p.flags & UP_FLAG_RESIDUAL) != 0;
  const int nacc = p.wide ? 1 : 2;      // accumulator buffers in TMEM (the 512-column tile has no second one)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (!nchw) tma_prefetch_desc(&tmY0);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), kPair ? 2 : 1);                 // pair: both CTAs' producers arrive on the leader's
      mbar_init(empty_bar(s), kPair ? 1 : p.cluster);        // pair: one multicast commit from the leader's issuer
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), (kEpiThreads / 32) * (kPair ? 2 : 1));   // pair: epilogue warps of both CTAs
    }
    for (int b = 0; b < kMaxBufs; ++b) {
      mbar_init(avail_bar(b), 1);
      mbar_init(ready_bar(b), kEpiThreads / 32);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (kPair) tmem_alloc_2cta(tmem_slot, p.tmem_cols);
    else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  if (p.res_terms) {
    // K-major, 128B-swizzled identity: row n holds a single 1.0 at k = n (16-byte chunk n/8 lands at (n/8)^(n&7))
    const uint32_t one = p.fmt == 1 ? 0x3F80u : 0x3C00u;
    for (uint32_t i = threadIdx.x; i < 8192u / 16u; i += blockDim.x) {
      const uint32_t n = i >> 3, chunk = i & 7u;
      const uint32_t src_chunk = chunk ^ (n & 7u);  // logical chunk stored at this physical position
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      // pair mode: this CTA supplies rows [32*rank, 32*rank+32) of the 64x64 identity as its half of the B operand
      const uint32_t gn = kPair ? (n + 32u * crank) : n;
      if ((!kPair || n < 32u) && src_chunk == (gn >> 3)) {
        const uint32_t e = gn & 7u;
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
  if (p.cluster > 1) {
    // peers multicast into our smem and arrive on our mbarriers: nobody may start before all are initialised
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tcgen05_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, identity tile) overlapped the
  // tail of the previous kernel; from here on we read its output.  Let our own dependents get scheduled early.
  if (threadIdx.x == 0) UP_STAMP(1);   // prologue done
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) UP_STAMP(2);   // dependencies resolved
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 || warp == 3) {
    // ===================== TMA producers (two warps: even / odd k-blocks) =====================
    // Each warp walks the whole (tap, channel chunk, term) loop nest with warp-uniform values - pure increments, no
    // divisions, no cross-lane traffic - and one elected lane issues the copies of every second k-block.  A k-block's slot
    // and barrier parity follow from its sequence number alone, and a producer cannot run more than `stages`
    // k-blocks ahead of the MMA issuer, so the two issue streams need no ordering between them.
    const int which = warp == 3 ? 1 : 0;
    uint32_t s = 0, wait_par = 1;   // slot and the parity of its `empty` barrier to wait for (fresh barrier: passes)
    int issued = 0;
    const int taps = p.taps_h * p.taps_w;
    const int res_units = (p.block_n >> 6) * p.res_terms;
    for (int tile = first_work; tile < total_tiles; tile += work_step) {
      const TileCoord t = decode_tile(p, tile, crank);
      int kh_lo, kh_hi, kw_lo, kw_hi;
      tap_range(p, p.taps_h, p.pad_h, t.h0, p.bh, p.Hq, kh_lo, kh_hi);
      tap_range(p, p.taps_w, p.pad_w, t.w0, p.bw, p.Wq, kw_lo, kw_hi);
      const int brow_nt = t.nt * p.block_n;
      if (p.tall) {   // the tall activation box starts at filter row 0 and covers all three rows
        kh_lo = 0;
        kh_hi = 0;
      }
      for (int kh = kh_lo; kh <= kh_hi; ++kh) {
        int oh, ph;
        tap_offset(p, kh, p.pad_h, oh, ph);
        for (int kw = kw_lo; kw <= kw_hi; ++kw) {
          int ow, pw;
          tap_offset(p, kw, p.pad_w, ow, pw);
          const int brow_tap = (kh * p.taps_w + kw) * p.cout + brow_nt;
          const int cbase = p.x_coff + pw * p.x_cs;
          int g = 0, cc = 0;
          for (int chunk = 0; chunk < p.chunks; ++chunk) {
            const int c = cbase + cc * p.ck;
            const int n = t.n0 + g * p.group_nstride;
            for (int term = 0; term < p.nterms; ++term) {
              // split mode: hi*hi (A0, B0), lo*hi (A1, B0), hi*lo (A0, B1)
              const CUtensorMap* amap = (term == 1) ? &tmA1 : &tmA0;
              const int brow = brow_tap + (term == 2 ? taps * p.cout : 0);
              const uint32_t dst = smem_base + s * stage_bytes;
              const bool mine = (issued & 1) == which;
              if (mine) mbar_wait(empty_bar(s), wait_par, 16000000000LL);
              if (mine && elect_one()) {
                if constexpr (kPair) {
                  // CTA pair: bytes of BOTH CTAs are credited to the leader's barrier (count 2: one arrive each)
                  const uint32_t mine = p.a_bytes + p.b_bytes;   // b_bytes = this CTA's half of the weight tile
                  if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * mine);
                  else mbar_arrive_remote(full_bar(s), 0u);
                  tma_load_5d_2cta(amap, dst, full_bar(s), c, t.w0 + ow, ph, t.h0 + oh, n);
                  if (p.tall) {
                    for (int r = 0; r < 3; ++r)
                      tma_load_2d_2cta(&tmB, dst + p.a_bytes + r * p.tall_b_bytes, full_bar(s), chunk * p.ck,
                                       brow + r * p.taps_w * p.cout + static_cast<int>(crank) * (p.block_n >> 1));
                  } else if (p.wide) {
                    // each N = 256 UMMA takes 128 filter rows from either CTA: this CTA holds rows
                    // [j*256 + rank*128, +128) of the 512-row tile, j = 0, 1
                    for (int j = 0; j < 2; ++j)
                      tma_load_2d_2cta(&tmB, dst + p.a_bytes + j * (p.b_bytes >> 1), full_bar(s), chunk * p.ck,
                                       brow + j * 256 + static_cast<int>(crank) * 128);
                  } else {
                    tma_load_2d_2cta(&tmB, dst + p.a_bytes, full_bar(s), chunk * p.ck,
                                     brow + static_cast<int>(crank) * (p.block_n >> 1));
                  }
                } else {
                  mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                  tma_load_5d(amap, dst, full_bar(s), c, t.w0 + ow, ph, t.h0 + oh, n);
                  if (p.cluster > 1) {
                    // each CTA fetches 1/cluster of the weight rows and multicasts them to every CTA of the cluster
                    const uint32_t sub_rows = static_cast<uint32_t>(p.block_n / p.cluster);
                    tma_load_2d_mc(&tmB, dst + p.a_bytes + crank * sub_rows * static_cast<uint32_t>(p.ck) * 2u,
                                   full_bar(s), chunk * p.ck, brow + static_cast<int>(crank * sub_rows),
                                   static_cast<uint16_t>((1u << p.cluster) - 1u));
                  } else if (p.tall) {
                    for (int r = 0; r < 3; ++r)
                      tma_load_2d(&tmB, dst + p.a_bytes + r * p.tall_b_bytes, full_bar(s), chunk * p.ck,
                                  brow + r * p.taps_w * p.cout);
                  } else {
                    const uint32_t part_rows = static_cast<uint32_t>(p.block_n / p.bsplit);
                    for (int j = 0; j < p.bsplit; ++j)
                      tma_load_2d(&tmB, dst + p.a_bytes + j * part_rows * static_cast<uint32_t>(p.ck) * 2u, full_bar(s),
                                  chunk * p.ck, brow + static_cast<int>(j * part_rows));
                  }
                }
                if (issued == 0) UP_STAMP(10);   // first k-block issued
                if (issued == 4) UP_STAMP(11);   // fifth k-block issued (slots were free)
              }
              ++issued;
              if (++s == static_cast<uint32_t>(p.stages)) {
                s = 0;
                wait_par ^= 1u;
              }
            }
            if (++cc == p.chunks_per_group) {
              cc = 0;
              ++g;
            }
          }
        }
      }
      // projection shortcut (UP_FLAG_PROJ): k-blocks over the SECOND input (tensor map A1; a stride-2 map reads parity
      // plane 0 at offset 0 = pixel (2h, 2w)) against the filter rows that follow the main taps
      for (int j = 0; j < p.proj_taps; ++j) {
        const int brow = (taps + j) * p.cout + brow_nt;
        for (int chunk = 0; chunk < p.chunks; ++chunk) {
          const int c = p.proj_coff + (j * p.chunks + chunk) * p.ck;
          const uint32_t dst = smem_base + s * stage_bytes;
          const bool mine = (issued & 1) == which;
          if (mine) mbar_wait(empty_bar(s), wait_par, 16000000000LL);
          if (mine && elect_one()) {
            if constexpr (kPair) {
              if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * stage_bytes);
              else mbar_arrive_remote(full_bar(s), 0u);
              tma_load_5d_2cta(&tmA1, dst, full_bar(s), c, t.w0, 0, t.h0, t.n0);
              tma_load_2d_2cta(&tmB, dst + p.a_bytes, full_bar(s), chunk * p.ck,
                               brow + static_cast<int>(crank) * (p.block_n >> 1));
            } else {
              mbar_arrive_expect_tx(full_bar(s), stage_bytes);
              tma_load_5d(&tmA1, dst, full_bar(s), c, t.w0, 0, t.h0, t.n0);
              tma_load_2d(&tmB, dst + p.a_bytes, full_bar(s), chunk * p.ck, brow);
            }
          }
          ++issued;
          if (++s == static_cast<uint32_t>(p.stages)) {
            s = 0;
            wait_par ^= 1u;
          }
        }
      }
      // residual slots: up to res_per_slot tiles [128 px x 64 ch] side by side; their B operand is the resident identity
      for (int u0 = 0; u0 < res_units; u0 += p.res_per_slot) {
        const int u_end = min(u0 + p.res_per_slot, res_units);
        const uint32_t dst = smem_base + s * stage_bytes;
        const bool mine = (issued & 1) == which;
        if (mine) mbar_wait(empty_bar(s), wait_par, 16000000000LL);
        if (mine && elect_one()) {
          const uint32_t bytes = static_cast<uint32_t>(u_end - u0) * p.a_bytes;
          if constexpr (kPair) {
            if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * bytes);
            else mbar_arrive_remote(full_bar(s), 0u);
          } else {
            mbar_arrive_expect_tx(full_bar(s), bytes);
          }
          for (int u = u0; u < u_end; ++u) {
            const int rg = u / p.res_terms;
            const CUtensorMap* rmap = (u - rg * p.res_terms) ? &tmR1 : &tmR0;   // hi / lo plane of the residual
            const int rc = p.r_coff + brow_nt + rg * 64;
            const uint32_t rdst = dst + static_cast<uint32_t>(u - u0) * p.a_bytes;
            if constexpr (kPair) tma_load_5d_2cta(rmap, rdst, full_bar(s), rc, t.w0, 0, t.h0, t.n0);
            else tma_load_5d(rmap, rdst, full_bar(s), rc, t.w0, 0, t.h0, t.n0);
          }
        }
        ++issued;
        if (++s == static_cast<uint32_t>(p.stages)) {
          s = 0;
          wait_par ^= 1u;
        }
      }
    }
  } else if (warp == 1 && (!kPair || crank == 0)) {
    // ===================== MMA issuer (pair mode: leader CTA only) =====================
    // The whole warp runs the loop converged (warp-uniform descriptors stay in uniform registers); one elected lane
    // issues the tcgen05 instructions.
    int s = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t sw_bytes = p.ck * 2;
    const int kk = p.ck / 16;
    const uint64_t adesc0 = make_smem_desc_kmajor(smem_base, sw_bytes);
    const uint64_t bdesc0 = make_smem_desc_kmajor(smem_base + p.a_bytes, sw_bytes);
    const uint32_t stage_step = stage_bytes >> 4;
    const uint64_t identdesc = make_smem_desc_kmajor(ident, 128);
    for (int tile = first_work; tile < total_tiles; tile += work_step) {
      const TileCoord t = decode_tile(p, tile, crank);
      int kh_lo, kh_hi, kw_lo, kw_hi;
      tap_range(p, p.taps_h, p.pad_h, t.h0, p.bh, p.Hq, kh_lo, kh_hi);
      tap_range(p, p.taps_w, p.pad_w, t.w0, p.bw, p.Wq, kw_lo, kw_hi);
      const int nkb_conv = (p.tall ? 1 : (kh_hi - kh_lo + 1)) * (kw_hi - kw_lo + 1) * p.chunks * p.nterms +
                           p.proj_taps * p.chunks;   // projection k-blocks look like ordinary ones to the issuer
      const int res_units = (p.block_n >> 6) * p.res_terms;
      const int nkb = nkb_conv + (p.res_terms ? (res_units + p.res_per_slot - 1) / p.res_per_slot : 0);
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tcgen05_after_thread_sync();
      const uint32_t tmem_d = tmem_base + acc * p.block_n;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), phase);
        tcgen05_after_thread_sync();
        const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage_step * s);
        if (elect_one()) {
          if (tile == first_work && kb == 0) UP_STAMP(3);   // first operands landed
          if (tile == first_work && kb == 2) UP_STAMP(12);  // third k-block landed
          if (kb >= nkb_conv) {
            // residual tiles of this slot: D[:, g*64 .. g*64+63] += R_tile x I  (exact: products with 1.0)
            const int u0 = (kb - nkb_conv) * p.res_per_slot;
            const int u_end = min(u0 + p.res_per_slot, res_units);
            for (int u = u0; u < u_end; ++u) {
              const int rg = u / p.res_terms;
              const uint64_t rdesc = adesc + static_cast<uint64_t>((p.a_bytes >> 4) * (u - u0));
              for (int k = 0; k < 4; ++k) {
                if constexpr (kPair) umma_f16_2cta(tmem_d + rg * 64, rdesc + 2u * k, identdesc + 2u * k, p.idesc_res, 1u);
                else umma_f16(tmem_d + rg * 64, rdesc + 2u * k, identdesc + 2u * k, p.idesc_res, 1u);
              }
            }
          } else if (p.tall) {
            // three filter rows from ONE activation box: row r reads it dil*bw pixel rows further down (a multiple
            // of the 1024-byte swizzle atom) against its own weight tile
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage_step * s);
            for (int r = 0; r < 3; ++r) {
              const uint64_t ar = adesc + static_cast<uint64_t>(p.tall_a_step * r);
              const uint64_t br = bdesc + static_cast<uint64_t>((p.tall_b_bytes >> 4) * r);
              for (int k = 0; k < kk; ++k) {
                if constexpr (kPair) umma_f16_2cta(tmem_d, ar + 2u * k, br + 2u * k, p.idesc, (kb | r | k) ? 1u : 0u);
                else umma_f16(tmem_d, ar + 2u * k, br + 2u * k, p.idesc, (kb | r | k) ? 1u : 0u);
              }
            }
          } else {
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(stage_step * s);
            for (int k = 0; k < kk; ++k) {
              // advance 16 elements (32 bytes) along K inside the swizzle row: +2 in 16-byte units
              if constexpr (kPair) {
                umma_f16_2cta(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc, (kb | k) ? 1u : 0u);
                if (p.wide)   // second half of the 512 output channels: same activation slot, next 128 filter rows
                  umma_f16_2cta(tmem_d + 256u, adesc + 2u * k, bdesc + static_cast<uint64_t>(p.b_bytes >> 5) + 2u * k,
                                p.idesc, (kb | k) ? 1u : 0u);
              } else {
                umma_f16(tmem_d, adesc + 2u * k, bdesc + 2u * k, p.idesc, (kb | k) ? 1u : 0u);
              }
            }
          }
          // frees the smem slot once these MMAs have read it - in every CTA of the cluster (peers multicast into it)
          if constexpr (kPair) umma_commit_2cta_mc(empty_bar(s), 3);
          else if (p.cluster > 1) umma_commit_mc(empty_bar(s), static_cast<uint16_t>((1u << p.cluster) - 1u));
          else umma_commit(empty_bar(s));
          if (tile == first_work && kb == 2) UP_STAMP(13);  // ... and its MMAs + commit issued
          if (kb == nkb - 1) {
            if (tile == first_work) UP_STAMP(4);                 // all MMAs of the first tile issued
            if constexpr (kPair) umma_commit_2cta_mc(tfull_bar(acc), 3);   // accumulator halves complete in both CTAs
            else umma_commit(tfull_bar(acc));                     // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++s == p.stages) {
          s = 0;
          phase ^= 1u;
        }
      }
      if (++acc == nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (threadIdx.x == 64) {
    // ===================== epilogue DMA thread =====================
    if (!nchw) {
      const int groups = p.block_n / 64;
      int my_tiles = 0;
      for (int tile = first_work; tile < total_tiles; tile += work_step) ++my_tiles;
      const int total_q = my_tiles * groups;
      auto coords = [&](int q, int& c, TileCoord& t) {
        const int tile = first_work + (q / groups) * work_step;
        t = decode_tile(p, tile, crank);
        c = t.nt * p.block_n + (q % groups) * 64;
      };
      // nbuf staging buffers, one group of look-ahead: while the epilogue warps fill buffer q % nbuf, this thread
      // waits until the TMA store of group q+1-nbuf has finished READING its buffer (at most nbuf-2 younger stores
      // still pending) and hands that buffer out for group q+1.
      mbar_arrive(avail_bar(0));
      for (int q = 0; q < total_q; ++q) {
        if (q + 1 < total_q) {
          if (q + 1 >= p.nbuf) {
            if (p.nbuf == 2) tma_store_wait_read<0>();
            else if (p.nbuf == 3) tma_store_wait_read<1>();
            else tma_store_wait_read<2>();
          }
          mbar_arrive(avail_bar((q + 1) % p.nbuf));
        }
        const int b = q % p.nbuf;
        mbar_wait(ready_bar(b), (q / p.nbuf) & 1u);
        int c;
        TileCoord t;
        coords(q, c, t);
        const uint32_t src = staging + b * p.buf_bytes;
        tma_store_5d(&tmY0, src, p.y_coff + c, t.w0, 0, t.h0, t.n0);
        if (p.split) tma_store_5d(&tmY1, src + kPlaneBytes, p.y_coff + c, t.w0, 0, t.h0, t.n0);
        tma_store_commit();
      }
      UP_STAMP(7);                  // last store issued
      tma_store_wait_all<0>();
      UP_STAMP(8);                  // stores complete
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue math (8 warps) =====================
    const int ew = warp - kEpiWarp0;
    const int quarter = ew & 3;  // == warp % 4 -> TMEM lane quarter this warp may read
    const int half = ew >> 2;    // which 32 of a group's 64 columns
    const int row = quarter * 32 + lane;
    const int fmt = p.fmt;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t q = 0;  // staging-buffer sequence number (shared convention with the DMA thread)
    const int bhw = p.bh * p.bw;
    const uint32_t rowoff = static_cast<uint32_t>(row) * 128u;
    for (int tile = first_work; tile < total_tiles; tile += work_step) {
      const TileCoord t = decode_tile(p, tile, crank);
      if (!nchw && lane < 2 * (p.block_n / 64)) {
        // pull this warp's scale / shift lines (128 B per 32 columns) into L1 while the MMAs of the tile still run
        const float* line = ((lane & 1) ? p.shift : p.scale) + t.nt * p.block_n + (lane >> 1) * 64 + half * 32;
        float sink;
        asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(sink) : "l"(line));
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      if (tile == first_work && threadIdx.x == kEpiWarp0 * 32) UP_STAMP(5);   // first accumulator complete
      tcgen05_after_thread_sync();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * p.block_n;

      if (nchw) {
        const int rn = row / bhw;
        const int rem = row - rn * bhw;
        const int rh = rem / p.bw;
        const int rw = rem - rh * p.bw;
        const int n = t.n0 + rn, h = t.h0 + rh, w = t.w0 + rw;
        const bool valid = (n < p.N) && (h < p.Ho) && (w < p.Wo);
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
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
        const EpiCtx ec{p.scale + t.nt * p.block_n + half * 32, p.shift + t.nt * p.block_n + half * 32,
                        taddr0 + static_cast<uint32_t>(half * 32), staging + rowoff, p.buf_bytes,
                        static_cast<uint32_t>(p.nbuf), static_cast<uint32_t>(half * 4),
                        static_cast<uint32_t>(row) & 7u, bars + 8u * (2 * kMaxStages + 4), lane, p.block_n / 64};
        const bool relu = (p.flags & UP_FLAG_RELU) != 0;
        if (p.split) {
          if (relu) epi_tile<2, true>(ec, q); else epi_tile<2, false>(ec, q);
        } else if (fmt == 1) {
          if (relu) epi_tile<1, true>(ec, q); else epi_tile<1, false>(ec, q);
        } else {
          if (relu) epi_tile<0, true>(ec, q); else epi_tile<0, false>(ec, q);
        }
      }
      // all TMEM reads of this accumulator are done -> hand it back to the MMA issuer
      if (tile == first_work && threadIdx.x == kEpiWarp0 * 32) UP_STAMP(6);     // first tile's epilogue math done
      if (tile == first_work + work_step && threadIdx.x == kEpiWarp0 * 32) UP_STAMP(14);      // second tile's
      if (tile == first_work + 2 * work_step && threadIdx.x == kEpiWarp0 * 32) UP_STAMP(15);  // third tile's
      tcgen05_before_thread_sync();
      __syncwarp();
      if (lane == 0) {
        if (kPair && crank != 0) mbar_arrive_remote(tempty_bar(acc), 0u);
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == nacc) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tcgen05_before_thread_sync();
  __syncthreads();
  if (threadIdx.x == 0) UP_STAMP(9);   // all roles done
  if (p.cluster > 1) {
    // a CTA must not exit while a peer can still multicast into its smem or arrive on its barriers
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 2) {
    tcgen05_after_thread_sync();
    if constexpr (kPair) tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace up
#include "up_conv_host.h"  // tensor-map encoding + tile picking helpers (shared with the wgrad kernel)
namespace up {

static unsigned long long* g_dbg_last = nullptr;

static int ensure_device(DeviceInfo*& di) {
  di = device_info();
  if (!di) return UP_ERR_CUDA;
  if (!di->conv_attr) {
    int rc = check_cuda(cudaFuncSetAttribute(conv_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(di->max_smem)),
                        "cudaFuncSetAttribute(max dynamic smem)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(conv_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(di->max_smem)),
                    "cudaFuncSetAttribute(max dynamic smem, pair)");
    if (rc) return rc;
    di->conv_attr = true;
  }
  return 0;
}

}  // namespace up

using namespace up;

// Debug only (UP_DEBUG_TIMING=1): copies the phase timestamps of the last conv launch (160 CTAs x 16 slots) to the host.
extern "C" int up_debug_conv_timing(unsigned long long* h_out) {
  if (!g_dbg_last) return up::fail(UP_ERR_INVALID, "no timing buffer (set UP_DEBUG_TIMING=1)");
  return up::check_cuda(cudaMemcpy(h_out, g_dbg_last, 160 * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost),
                        "cudaMemcpy(timing)");
}

extern "C" int up_conv2d_fwd(const UpConvDesc* d, const void* x, const void* w_packed, const float* scale,
                             const float* shift, const void* residual, void* y, float* stats, void* stream) {
  (void)stats;
  UP_CHECK_ARG(d && x && w_packed && scale && shift && y, "up_conv2d_fwd: null argument");
  UP_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0 && d->ho > 0 && d->wo > 0, "up_conv2d_fwd: bad spatial dims");
  UP_CHECK_ARG(d->stride == 1 || d->stride == 2, "up_conv2d_fwd: stride must be 1 or 2 (got %d)", d->stride);
  UP_CHECK_ARG(d->stride == 1 || (d->h % 2 == 0 && d->w % 2 == 0), "up_conv2d_fwd: stride 2 needs even h, w");
  UP_CHECK_ARG(d->kh >= 1 && d->kw >= 1 && d->dil >= 1 && d->pad_h >= 0 && d->pad_w >= 0, "up_conv2d_fwd: bad filter");
  UP_CHECK_ARG(d->cin > 0 && d->cin % 16 == 0, "up_conv2d_fwd: cin (%d) must be a positive multiple of 16", d->cin);
  UP_CHECK_ARG(d->cout > 0 && d->cout % 32 == 0, "up_conv2d_fwd: cout (%d) must be a positive multiple of 32", d->cout);
  UP_CHECK_ARG(d->dtype == UP_BF16 || d->dtype == UP_FP16 || d->dtype == UP_SPLIT, "up_conv2d_fwd: bad dtype");
  UP_CHECK_ARG(!(d->flags & UP_FLAG_STATS), "up_conv2d_fwd: UP_FLAG_STATS is not fused; use up_bn_stats");
  const int groups = d->x_groups > 0 ? d->x_groups : 1;
  UP_CHECK_ARG(d->cin % groups == 0, "up_conv2d_fwd: cin not divisible by x_groups");
  const int cin_g = d->cin / groups;
  const int ck = (cin_g % 64 == 0) ? 64 : 16;
  UP_CHECK_ARG(cin_g % ck == 0, "up_conv2d_fwd: per-group cin (%d) must be a multiple of 16", cin_g);
  const int cext = d->x_cextent > 0 ? d->x_cextent : d->x_cstride;
  UP_CHECK_ARG(d->x_cstride % 8 == 0 && d->x_coff % 8 == 0 && d->x_coff + cin_g <= cext,
               "up_conv2d_fwd: bad x channel view (cstride %d coff %d cin/group %d extent %d)", d->x_cstride, d->x_coff,
               cin_g, cext);
  UP_CHECK_ARG(d->x_cextent == 0 || (d->stride == 1 && groups == 1),
               "up_conv2d_fwd: overlapping channel windows need stride 1 and no groups");
  UP_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0,
               "up_conv2d_fwd: x / w must be 16-byte aligned");
  const bool nchw = (d->flags & UP_FLAG_OUT_NCHW_F32) != 0;
  const bool split = d->dtype == UP_SPLIT;
  const bool has_res = (d->flags & UP_FLAG_RESIDUAL) != 0;
  if (nchw) {
    UP_CHECK_ARG(d->cout_valid > 0 && d->cout_valid <= d->cout, "up_conv2d_fwd: bad cout_valid");
    UP_CHECK_ARG(d->out_c_total == 0 || d->out_c_total >= d->cout_valid, "up_conv2d_fwd: bad out_c_total");
    UP_CHECK_ARG(!has_res, "up_conv2d_fwd: residual not supported with NCHW fp32 output");
  } else {
    UP_CHECK_ARG(d->cout % 64 == 0, "up_conv2d_fwd: NHWC output needs cout %% 64 == 0 (got %d)", d->cout);
    UP_CHECK_ARG(d->y_cstride % 8 == 0 && d->y_coff % 8 == 0 && d->y_coff + d->cout <= d->y_cstride,
                 "up_conv2d_fwd: bad y channel view");
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "up_conv2d_fwd: y must be 16-byte aligned");
    if (split) UP_CHECK_ARG(d->y_plane_stride % 8 == 0 && d->y_plane_stride > 0, "up_conv2d_fwd: bad y_plane_stride");
  }
  if (split) UP_CHECK_ARG(d->x_plane_stride % 8 == 0 && d->x_plane_stride > 0, "up_conv2d_fwd: bad x_plane_stride");
  if (has_res) {
    UP_CHECK_ARG(residual != nullptr, "up_conv2d_fwd: residual pointer missing");
    UP_CHECK_ARG(d->r_cstride % 8 == 0 && d->r_coff % 8 == 0 && d->r_coff + d->cout <= d->r_cstride,
                 "up_conv2d_fwd: bad residual channel view");
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0, "up_conv2d_fwd: residual must be 16B aligned");
    if (split) UP_CHECK_ARG(d->r_plane_stride % 8 == 0 && d->r_plane_stride > 0, "up_conv2d_fwd: bad r_plane_stride");
  }
  const bool has_proj = (d->flags & UP_FLAG_PROJ) != 0;
  if (has_proj) {
    // second input x2 (passed in `residual`, viewed through r_cstride / r_coff): its 1x1 projection of stride
    // proj_stride is accumulated into the same tile; the packed filter holds proj_cin / cin extra "taps" of cin columns
    UP_CHECK_ARG(residual != nullptr && !has_res && !split && !nchw, "up_conv2d_fwd: UP_FLAG_PROJ needs the second input "
                 "in `residual`, 16-bit NHWC output, no UP_FLAG_RESIDUAL, no split mode");
    UP_CHECK_ARG(d->kh == 1 && d->kw == 1 && d->stride == 1 && groups == 1 && ck == 64 && d->x_cextent == 0,
                 "up_conv2d_fwd: UP_FLAG_PROJ needs a 1x1 stride-1 main filter with cin %% 64 == 0");
    UP_CHECK_ARG(d->proj_cin > 0 && d->proj_cin % d->cin == 0, "up_conv2d_fwd: proj_cin (%d) must be a multiple of cin (%d)",
                 d->proj_cin, d->cin);
    UP_CHECK_ARG(d->proj_stride == 1 || d->proj_stride == 2, "up_conv2d_fwd: proj_stride must be 1 or 2");
    UP_CHECK_ARG(d->r_cstride % 8 == 0 && d->r_coff % 8 == 0 && d->r_coff + d->proj_cin <= d->r_cstride,
                 "up_conv2d_fwd: bad projection input channel view");
    UP_CHECK_ARG((reinterpret_cast<uintptr_t>(residual) & 15) == 0, "up_conv2d_fwd: projection input must be 16B aligned");
  }
  // every tile must see at least one in-bounds tap: the centre of the receptive field has to hit the image
  UP_CHECK_ARG(d->pad_h <= (d->kh - 1) * d->dil && d->pad_w <= (d->kw - 1) * d->dil,
               "up_conv2d_fwd: padding larger than the filter extent");

  DeviceInfo* di = nullptr;
  int rc = ensure_device(di);
  if (rc) return rc;
  const int g_sm_count = di->sm_count;
  const size_t g_max_smem = di->max_smem;

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
  // keep at least ~2 tiles per SM so the epilogue of one tile overlaps the main loop of the next
  // The kernel is bound by the bytes each SM has to ingest per MMA (A 16 KB + B block_n*128 B per k-block at ~48 B/clk),
  // so wide N tiles win even when that leaves a single wave; only shrink when more than ~40% of the SMs would idle.
  while (block_n > 64 && static_cast<long long>(m_tiles) * (d->cout / block_n) * 10 < 6LL * g_sm_count) block_n /= 2;
  if (const char* e = getenv("UP_DEBUG_BLOCKN")) {
    const int v = atoi(e);
    if (v >= 32 && d->cout % v == 0 && (nchw || v % 64 == 0)) block_n = v;
  }
  p.block_n = block_n;
  p.n_tiles = d->cout / block_n;
  p.cout = d->cout;
  p.nterms = split ? 3 : 1;
  p.a_bytes = kTileM * ck * 2;
  p.b_bytes = block_n * ck * 2;
  p.buf_bytes = split ? 2 * kPlaneBytes : kPlaneBytes;
  p.nbuf = nchw ? 0 : 2;
  p.res_terms = has_res ? (split ? 2 : 1) : 0;
  p.res_per_slot = 1;
  p.bsplit = 1;
  // cluster along the image-group direction: largest of 4 / 2 that divides tiles_n and leaves >= 8 weight rows per CTA
  p.cluster = 1;
  {
    int want = 2;   // 148 SMs pack perfectly into 74 pairs; clusters of 4 strand SMs (GPC sizes 16/18/20)
    if (const char* e = getenv("UP_CLUSTER")) want = atoi(e);
    for (int c = 4; c >= 2; c /= 2) {
      if (c <= want && p.tiles_n % c == 0 && (block_n / c) % 8 == 0 &&
          static_cast<long long>(m_tiles) * (d->cout / block_n) >= 2LL * c) {
        p.cluster = c;
        break;
      }
    }
  }
  p.pair = 0;
  if (p.cluster == 2 && block_n >= 64) {
    // CTA pair (tcgen05 cta_group::2, M = 256): each CTA fetches only half of the weight tile -> fewer L2->SM bytes
    // per FLOP.  Measured faster than independent CTAs on every layer shape of the network except the HBM-bound
    // 64->256 expansion of layer1 (-1%), so it is the default wherever a pair can form; UP_PAIR=0 / 1 overrides.
    p.pair = nchw ? 0 : 1;
    if (const char* e = getenv("UP_PAIR")) p.pair = (e[0] == '1') ? 1 : 0;
  }
  if (!p.pair && (has_proj || !getenv("UP_CLUSTER"))) p.cluster = 1;   // plain multicast clusters measured slower than independent CTAs
  if (const char* e = getenv("UP_DEBUG_BSPLIT")) {
    const int v = atoi(e);
    if (p.cluster == 1 && !has_proj && (v == 2 || v == 4) && block_n % (8 * v) == 0) p.bsplit = v;
  }
  if (p.pair) p.b_bytes = static_cast<uint32_t>(block_n / 2) * ck * 2;   // each CTA of the pair holds half of the weight tile
  // N = 512 tiles (CTA pairs, cout % 512 == 0): the activation tile - the unique bytes that bound these kernels at
  // ~57 KB/us per SM - is fetched once per 512 output channels instead of once per 256.  The single 512-column
  // accumulator cannot overlap a tile's epilogue with the next tile's MMAs, so only when the whole launch is ONE round
  // of work items and the k-loop is long enough to carry it.  UP_WIDE_N: 0 off, 1 = 1x1 filters only, 2 = all (default).
  // Measured at batch 32 (profiles/segments_r2_wide.txt): layer4 conv1 2048 -> 512 45.1 -> 41.0 us, conv2 3x3 512 -> 512
  // 75.8 -> 65.5 us.
  p.wide = 0;
  {
    static const int want = []() {
      const char* e = getenv("UP_WIDE_N");
      return e ? atoi(e) : 2;
    }();
    const long long items = static_cast<long long>(m_tiles / 2) * (d->cout / 512 > 0 ? d->cout / 512 : 1);
    const int kblocks = d->kh * d->kw * (d->cin / ck);
    // want == 3 (experiment): also the short-K 1x1 expansions with a residual that need several rounds anyway
    // (layer4 conv3 512 -> 2048: a quarter fewer bytes through the ring per output channel, no epilogue overlap)
    const bool single_round = items <= g_sm_count / 2 && 2 * items > g_sm_count / 2 && !has_res && kblocks >= 16;
    const bool multi_round = want >= 3 && has_res && d->kh == 1 && d->kw == 1 && items > g_sm_count / 2 && kblocks >= 8;
    if (want > 0 && p.pair && block_n == 256 && d->cout % 512 == 0 && !split && !has_proj && groups == 1 &&
        ck == 64 && p.bsplit == 1 && (want >= 2 || (d->kh == 1 && d->kw == 1)) &&
        (single_round || multi_round)) {   // one round instead of two; small launches keep the
                                           // N = 256 tiles that spread over twice the SMs
      p.wide = 1;
      block_n = 512;
      p.block_n = 512;
      p.n_tiles = d->cout / 512;
      p.b_bytes = 256u * ck * 2;     // this CTA's two 128-row halves
    }
  }
  // Filter-row reuse ("tall" activation box): 3x3, stride 1, same padding, single-image tiles with bw a multiple of 8
  // (row offsets stay aligned to the 1024-byte swizzle atom), and a slot (tall box + three weight tiles) small
  // enough for >= 3 pipeline stages.  Cuts the activation bytes of the k-loop by 3*bh / (bh + 2*dil).
  p.tall = 0;
  p.tall_a_step = 0;
  p.tall_b_bytes = 0;
  {
    const bool shape_ok = !p.wide && d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad_h == d->dil && ck == 64 && p.bn == 1 &&
                          p.bw % 8 == 0 && p.cluster * p.bsplit == (p.pair ? 2 : 1) && d->x_cextent == 0 && !has_res;
    const uint32_t tall_a = static_cast<uint32_t>(p.bh + 2 * d->dil) * p.bw * 128u;
    const uint32_t slot = tall_a + 3u * p.b_bytes;
    const char* e = getenv("UP_TALL");
    const bool want = e ? (e[0] == '1') : true;
    uint32_t min_stages = 3;
    if (const char* ms = getenv("UP_TALL_MIN_STAGES")) min_stages = atoi(ms) == 2 ? 2u : 3u;
    if (want && shape_ok && p.bh + 2 * d->dil <= 256 && min_stages * slot + 2u * p.buf_bytes + 4096u <= g_max_smem) {
      p.tall = 1;
      p.tall_a_step = (static_cast<uint32_t>(d->dil) * p.bw * 128u) >> 4;
      p.tall_b_bytes = p.b_bytes;
      p.a_bytes = tall_a;
      p.b_bytes = 3u * p.b_bytes;
    }
  }
  p.idesc = make_idesc_f16(static_cast<uint32_t>(fmt), p.pair ? 256u : kTileM, static_cast<uint32_t>(p.wide ? 256 : block_n));
  p.idesc_res = make_idesc_f16(static_cast<uint32_t>(fmt), p.pair ? 256u : kTileM, 64u);
  if (has_res) UP_CHECK_ARG(ck == 64, "up_conv2d_fwd: residual needs cin to be a multiple of 64");
  if (const char* e = getenv("UP_DEBUG_NBUF")) {
    const int v = atoi(e);
    if (!nchw && v >= 2 && v <= kMaxBufs) p.nbuf = v;
  }
  const size_t fixed = 1024 + 8 * (2 * kMaxStages + 4 + 2 * kMaxBufs) + 16 + static_cast<size_t>(p.nbuf) * p.buf_bytes +
                       (has_res ? 8192 : 0);
  int stages = static_cast<int>((g_max_smem - fixed) / (p.a_bytes + p.b_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("UP_DEBUG_STAGES")) {
    const int v = atoi(e);
    if (v >= 2 && v < stages) stages = v;
  }
  UP_CHECK_ARG(stages >= 2, "up_conv2d_fwd: not enough shared memory for 2 pipeline stages");
  p.stages = stages;
  if (has_res) {
    // several 16 KB residual tiles ride in one ring slot (fewer slot round trips per tile)
    p.res_per_slot = static_cast<int>((p.a_bytes + p.b_bytes) / p.a_bytes);
    if (const char* e = getenv("UP_DEBUG_RES_PER_SLOT")) p.res_per_slot = atoi(e) >= 1 ? std::min(atoi(e), p.res_per_slot) : 1;
  }
  size_t fixed_all = fixed;
  if (!nchw && !getenv("UP_DEBUG_NBUF")) {
    // shared memory the operand ring cannot use becomes extra staging buffers (more TMA stores in flight)
    while (p.nbuf < kMaxBufs &&
           fixed_all + p.buf_bytes + static_cast<size_t>(stages) * (p.a_bytes + p.b_bytes) <= g_max_smem) {
      fixed_all += p.buf_bytes;
      ++p.nbuf;
    }
  }
  // NOTE: filled again below once the cluster / pair decision is known
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.wide ? block_n : 2 * block_n)) cols *= 2;
  p.tmem_cols = cols;
  p.flags = d->flags;
  p.fmt = fmt;
  p.split = split ? 1 : 0;
  p.cout_valid = d->cout_valid;
  p.out_c_total = d->out_c_total > 0 ? d->out_c_total : d->cout_valid;
  p.y_coff = d->y_coff;
  p.r_coff = d->r_coff;
  p.proj_taps = has_proj ? d->proj_cin / d->cin : 0;
  p.proj_coff = d->r_coff;
  p.scale = scale;
  p.shift = shift;
  p.out_f32 = nchw ? static_cast<float*>(y) : nullptr;
  p.dbg = nullptr;
  if (getenv("UP_DEBUG_TIMING")) {
    static unsigned long long* dbuf = nullptr;
    if (!dbuf) cudaMalloc(&dbuf, 160 * 16 * sizeof(unsigned long long));
    cudaMemsetAsync(dbuf, 0, 160 * 16 * sizeof(unsigned long long), static_cast<cudaStream_t>(stream));
    p.dbg = dbuf;
    g_dbg_last = dbuf;
  }

  // ---- tensor maps ----
  CUtensorMap tmA0, tmA1, tmB, tmY0, tmY1, tmR0, tmR1;
  const int sw = ck * 2;
  const uint32_t abox[5] = {static_cast<uint32_t>(ck), static_cast<uint32_t>(p.bw), 1u,
                            static_cast<uint32_t>(p.tall ? p.bh + 2 * d->dil : p.bh), static_cast<uint32_t>(p.bn)};
  const int n_total = d->n + (groups - 1) * p.group_nstride;
  rc = encode_act_map(&tmA0, fmt, x, n_total, d->h, d->w, d->x_cstride, d->stride, abox, sw, "x", d->x_cextent,
                      d->x_wpitch);
  if (rc) return rc;
  if (split) {
    rc = encode_act_map(&tmA1, fmt, static_cast<const uint16_t*>(x) + d->x_plane_stride, n_total, d->h, d->w,
                        d->x_cstride, d->stride, abox, sw, "x.lo", d->x_cextent, d->x_wpitch);
    if (rc) return rc;
  } else if (has_proj) {
    rc = encode_act_map(&tmA1, fmt, residual, d->n, d->ho * d->proj_stride, d->wo * d->proj_stride, d->r_cstride,
                        d->proj_stride, abox, sw, "x2 (projection input)");
    if (rc) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    const int taps = d->kh * d->kw + p.proj_taps;   // the projection's filter rows follow the main taps
    const uint64_t dims[2] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(split ? 2 : 1) * taps * d->cout};
    const uint64_t st[1] = {static_cast<uint64_t>(d->cin) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(ck), static_cast<uint32_t>(p.wide ? 128 : block_n / p.cluster / p.bsplit)};
    if (split) {
      UP_CHECK_ARG(d->w_plane_stride == static_cast<int64_t>(taps) * d->cout * d->cin,
                   "up_conv2d_fwd: split weights must have contiguous planes (w_plane_stride = taps*cout*cin)");
    }
    rc = encode_map(&tmB, fmt, 2, w_packed, dims, st, box, sw, "w");
    if (rc) return rc;
  }
  tmY0 = tmA0;
  tmY1 = tmA0;
  tmR0 = tmA0;
  tmR1 = tmA0;
  if (!nchw) {
    const uint32_t ybox[5] = {64u, static_cast<uint32_t>(p.bw), 1u, static_cast<uint32_t>(p.bh),
                              static_cast<uint32_t>(p.bn)};
    rc = encode_act_map(&tmY0, fmt, y, d->n, d->ho, d->wo, d->y_cstride, 1, ybox, 128, "y");
    if (rc) return rc;
    tmY1 = tmY0;
    if (split) {
      rc = encode_act_map(&tmY1, fmt, static_cast<uint16_t*>(y) + d->y_plane_stride, d->n, d->ho, d->wo,
                          d->y_cstride, 1, ybox, 128, "y.lo");
      if (rc) return rc;
    }
    if (has_res) {
      rc = encode_act_map(&tmR0, fmt, residual, d->n, d->ho, d->wo, d->r_cstride, 1, ybox, 128, "residual");
      if (rc) return rc;
      tmR1 = tmR0;
      if (split) {
        rc = encode_act_map(&tmR1, fmt, static_cast<const uint16_t*>(residual) + d->r_plane_stride, d->n, d->ho,
                            d->wo, d->r_cstride, 1, ybox, 128, "residual.lo");
        if (rc) return rc;
      }
    }
  }

  const long long total_work = static_cast<long long>(m_tiles / p.cluster) * p.n_tiles;
  const size_t smem = fixed_all + static_cast<size_t>(stages) * (p.a_bytes + p.b_bytes);
  int max_clusters = g_sm_count / p.cluster;
  if (p.cluster > 1) {
    // persistent kernel: never launch more clusters than can be co-resident
    int* cached = di->max_clusters;
    if (cached[p.cluster] == 0) {
      cudaLaunchConfig_t occ{};
      occ.gridDim = dim3(g_sm_count / p.cluster * p.cluster);
      occ.blockDim = dim3(kThreads);
      occ.dynamicSmemBytes = g_max_smem;
      cudaLaunchAttribute oa[1];
      oa[0].id = cudaLaunchAttributeClusterDimension;
      oa[0].val.clusterDim.x = p.cluster;
      oa[0].val.clusterDim.y = 1;
      oa[0].val.clusterDim.z = 1;
      occ.attrs = oa;
      occ.numAttrs = 1;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, conv_tcgen05_kernel<true>, &occ) == cudaSuccess && nc > 0) {
        cached[p.cluster] = nc;
      } else {
        (void)cudaGetLastError();
        cached[p.cluster] = g_sm_count / p.cluster;
      }
    }
    if (cached[p.cluster] < max_clusters) max_clusters = cached[p.cluster];
  }
  const int grid = static_cast<int>(total_work < max_clusters ? total_work : max_clusters) * p.cluster;
  static const bool use_pdl = []() {
    const char* e = getenv("UP_PDL");
    return !(e && e[0] == '0');
  }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  rc = check_cuda(p.pair ? cudaLaunchKernelEx(&cfg, conv_tcgen05_kernel<true>, tmA0, tmA1, tmB, tmY0, tmY1, tmR0, tmR1, p)
                         : cudaLaunchKernelEx(&cfg, conv_tcgen05_kernel<false>, tmA0, tmA1, tmB, tmY0, tmY1, tmR0, tmR1, p),
                  "conv_tcgen05_kernel launch");
  if (rc) return rc;
  return 0;
}
