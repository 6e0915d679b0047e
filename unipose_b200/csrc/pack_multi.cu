// Table-driven weight preparation: ONE launch (re)packs every convolution filter of a plan and ONE launch derives every
// epilogue constant, instead of ~8 small launches per layer (116 layers, and twice that many packed filters in a
// training step whose weights change every step).
//
//   up_epilogue_consts   per layer: eval-mode BatchNorm fold (scale = gamma / sqrt(var + eps), shift = beta - mean * scale:
//                        resnet.py:26-34, wasp.py:18,86, decoder.py:40 in .eval()) or the conv bias -> the fp32
//                        [cout_pad] scale / shift vectors of the conv epilogue, zero padded
//   up_pack_conv_weights per filter: OIHW fp32 -> 16-bit [plane][tap][rows][cols] (zero padded), optionally scaled per
//                        output channel (the folded BatchNorm scale), optionally the dgrad layout
//                        (rows = input channels of a slice, cols = output channels, taps flipped)
// Both read their job tables from DEVICE memory (the host side uploads a table once per plan and whenever a parameter
// is re-allocated).
#include "up_internal.h"

namespace up {

constexpr int kPackTile = 32;      // co x ci tile of the transposing path
constexpr int kPackMaxTaps = 9;    // filters with more taps (7x7 stem, 11x11 video) take the direct path
constexpr int kPackThreads = 256;

__device__ __forceinline__ void store_packed(uint16_t* out, long long idx, long long plane, int dtype, float v) {
  if (dtype == UP_SPLIT) {
    uint16_t hi, lo;
    split_bf16(v, hi, lo);
    out[idx] = hi;
    out[idx + plane] = lo;
  } else {
    out[idx] = cvt_f32_to16_rt(v, dtype == UP_FP16 ? 0 : 1);
  }
}

// Block b works on tile (b - job.tile_start) of the job found by binary search over tile_start.
__global__ void __launch_bounds__(kPackThreads)
    pack_conv_weights_kernel(const UpPackJob* __restrict__ jobs, int njobs) {
  __shared__ float tile[kPackTile * (kPackTile * kPackMaxTaps + 1)];
  int lo = 0, hi = njobs - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_start <= b) lo = mid;
    else hi = mid - 1;
  }
  const UpPackJob j = jobs[lo];
  const long long t = b - j.tile_start;
  const int taps = j.kh * j.kw;
  const float* __restrict__ w = j.w;
  uint16_t* __restrict__ out = static_cast<uint16_t*>(j.out);
  // packed geometry: rows x cols per tap; forward: rows = cout, cols = cin; dgrad: rows = cin slice, cols = cout
  const int rows = j.rows, cols = j.cols;
  if (taps > kPackMaxTaps) {
    // direct path: one thread per packed element of a 1024-element chunk
    const long long total = static_cast<long long>(taps) * rows * cols;
    const long long base = t * 1024;
    for (long long i = base + threadIdx.x; i < base + 1024 && i < total; i += kPackThreads) {
      const int c = static_cast<int>(i % cols);
      const long long q = i / cols;
      const int r = static_cast<int>(q % rows);
      const int tap = static_cast<int>(q / rows);
      const int co = j.transpose ? c : r;
      const int ci = j.transpose ? r : c;
      float v = 0.f;
      if (co < j.cout_real && ci < j.cin_slice) {
        const int st = j.transpose ? (taps - 1 - tap) : tap;
        v = w[(static_cast<long long>(co) * j.cin_total + j.ci_off + ci) * taps + st];
        if (j.row_scale) v *= j.row_scale[co % j.scale_period];
      }
      store_packed(out, i, j.plane_stride, j.dtype, v);
    }
    return;
  }
  // transposing path: a 32 (co) x 32 (ci) tile; every source row segment of 32*taps floats is contiguous in OIHW
  const int tiles_ci = (max(j.transpose ? rows : cols, 1) + kPackTile - 1) / kPackTile;
  const int co0 = static_cast<int>(t / tiles_ci) * kPackTile;
  const int ci0 = static_cast<int>(t % tiles_ci) * kPackTile;
  const int pitch = kPackTile * taps + 1;
  const int nci = min(kPackTile, j.cin_slice - ci0);       // real input channels in this tile (may be <= 0)
  const int seg = max(nci, 0) * taps;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the kernel is issue-bound (ncu: 67 % issue slots, 8 % of the DRAM peak), so no per-element divisions: a warp owns
  // source rows warp, warp + 8, ... and walks the contiguous segment
  for (int r = warp; r < kPackTile; r += kPackThreads / 32) {
    const int co = co0 + r;
    const bool live = co < j.cout_real;
    const float* src = w + (static_cast<long long>(co) * j.cin_total + j.ci_off + ci0) * taps;
    const float rs = (live && j.row_scale) ? __ldg(j.row_scale + co % j.scale_period) : 1.f;
    for (int c = lane; c < seg; c += 32) tile[r * pitch + c] = live ? __ldg(src + c) * rs : 0.f;
  }
  __syncthreads();
  // a warp owns packed rows (tap, slow) = warp, warp + 8, ...; lane = the fast index, 32 consecutive 16-bit outputs
  const bool tr = j.transpose != 0;
  const int fast_ext = tr ? j.cols - co0 : j.cols - ci0;     // packed columns left from this tile's first column
  for (int q = warp; q < taps * kPackTile; q += kPackThreads / 32) {
    const int slow = q & (kPackTile - 1);
    const int tap = q >> 5;
    // forward: fast = ci (cols), slow = co (rows);  dgrad: fast = co (cols), slow = ci (rows)
    const int r_co = tr ? lane : slow;
    const int c_ci = tr ? slow : lane;
    const int row = tr ? ci0 + slow : co0 + slow;
    if (row >= rows || lane >= fast_ext) continue;
    float v = 0.f;
    if (co0 + r_co < j.cout_real && c_ci < nci) v = tile[r_co * pitch + c_ci * taps + (tr ? (taps - 1 - tap) : tap)];
    const long long o = (static_cast<long long>(tap) * rows + row) * cols + (tr ? co0 : ci0) + lane;
    store_packed(out, o, j.plane_stride, j.dtype, v);
  }
}

__global__ void epilogue_consts_kernel(const UpEpilogueJob* __restrict__ jobs, int njobs) {
  const UpEpilogueJob j = jobs[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < j.c_bn && j.fold_scale) {
    // same operation order as ATen's batch_norm inference path: invstd = 1/sqrt(var+eps)
    j.fold_scale[c] = j.gamma[c] * (1.0f / sqrtf(j.var[c] + j.eps));
  }
  if (c >= j.cout) return;
  float s = 0.f, b = 0.f;
  if (c < j.cout_real) {
    s = 1.0f;
    if (j.c_bn > 0) {
      const int k = c % j.c_bn;
      const float fs = j.gamma[k] * (1.0f / sqrtf(j.var[k] + j.eps));
      b = j.beta[k] - j.mean[k] * fs;
      if (!j.fold_into_weights) s = fs;
    } else if (j.bias) {
      b = j.bias[c % j.bias_len];
    }
  }
  j.scale[c] = s;
  j.shift[c] = b;
}

}  // namespace up

using namespace up;

extern "C" int64_t up_pack_job_tiles(const UpPackJob* h_job) {
  if (!h_job || h_job->kh <= 0 || h_job->kw <= 0 || h_job->rows <= 0 || h_job->cols <= 0) return -1;
  const int taps = h_job->kh * h_job->kw;
  if (taps > kPackMaxTaps) return (static_cast<int64_t>(taps) * h_job->rows * h_job->cols + 1023) / 1024;
  const int co_ext = h_job->transpose ? h_job->cols : h_job->rows;
  const int ci_ext = h_job->transpose ? h_job->rows : h_job->cols;
  return static_cast<int64_t>((co_ext + kPackTile - 1) / kPackTile) * ((ci_ext + kPackTile - 1) / kPackTile);
}

extern "C" int up_pack_conv_weights(const UpPackJob* d_jobs, int njobs, int64_t total_tiles, void* stream) {
  UP_CHECK_ARG(d_jobs && njobs > 0 && total_tiles > 0 && total_tiles < (1ll << 31), "up_pack_conv_weights: bad job table");
  pack_conv_weights_kernel<<<static_cast<unsigned>(total_tiles), kPackThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_jobs, njobs);
  UP_CHECK_LAUNCH("pack_conv_weights_kernel");
  return 0;
}

extern "C" int up_epilogue_consts(const UpEpilogueJob* d_jobs, int njobs, int max_channels, void* stream) {
  UP_CHECK_ARG(d_jobs && njobs > 0 && max_channels > 0, "up_epilogue_consts: bad job table");
  const dim3 grid((max_channels + 127) / 128, njobs);
  epilogue_consts_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_jobs, njobs);
  UP_CHECK_LAUNCH("epilogue_consts_kernel");
  return 0;
}
