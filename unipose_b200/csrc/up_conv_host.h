// Host-side helpers shared by the conv forward/dgrad and wgrad launchers: cuTensorMapEncodeTiled access,
// NHWC activation tensor maps (rank 5, stride-2 parity folding) and output-pixel tile selection.
#pragma once
#include <cuda.h>

#include "up_internal.h"

namespace up {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

static int encode_map(CUtensorMap* m, int fmt, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes, const char* what) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(UP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5];
  cuuint64_t s[5];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 0  ? CU_TENSOR_MAP_SWIZZLE_NONE
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(UP_ERR_CUDA,
                "cuTensorMapEncodeTiled(%s) failed: %d (rank %d dims %llu,%llu,%llu,%llu,%llu box %u,%u,%u,%u,%u)",
                what, static_cast<int>(r), rank, (unsigned long long)d[0], (unsigned long long)d[1],
                (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
                (unsigned long long)(rank > 4 ? d[4] : 0), b[0], b[1], rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0,
                rank > 4 ? b[4] : 0);
  }
  return 0;
}

// NHWC activation view as a rank-5 tensor map (c, w, parity|1, h, n); stride 2 folds the row/column
// parity into dims 2 / 0 so that a stride-2 tap is still a dense box.
//
// `cextent` / `wpitch` (stride 1 only) describe OVERLAPPING channel windows: the innermost dimension then spans
// `cextent` (> cs) consecutive elements, i.e. several neighbouring pixels, while consecutive "pixels" are still cs
// elements apart and rows are `wpitch` pixels apart in memory.  The space-to-depth stem uses this to fetch 4
// horizontal taps x 16 channels as one 64-element K-chunk.
static int encode_act_map(CUtensorMap* m, int fmt, const void* base, int n_total, int h, int w, int cs, int stride,
                          const uint32_t* box, int swizzle_bytes, const char* what, int cextent = 0, int wpitch = 0) {
  uint64_t dims[5];
  uint64_t st[4];
  const uint64_t es = 2;
  if (stride == 1) {
    const uint64_t wp = wpitch > 0 ? wpitch : w;
    dims[0] = cextent > 0 ? cextent : cs;
    dims[1] = w;
    dims[2] = 1;
    dims[3] = h;
    dims[4] = n_total;
    st[0] = cs * es;
    st[1] = wp * cs * es;
    st[2] = wp * cs * es;
    st[3] = static_cast<uint64_t>(h) * wp * cs * es;
  } else {
    dims[0] = 2ull * cs;
    dims[1] = w / 2;
    dims[2] = 2;
    dims[3] = h / 2;
    dims[4] = n_total;
    st[0] = 2ull * cs * es;
    st[1] = static_cast<uint64_t>(w) * cs * es;
    st[2] = 2ull * w * cs * es;
    st[3] = static_cast<uint64_t>(h) * w * cs * es;
  }
  return encode_map(m, fmt, 5, base, dims, st, box, swizzle_bytes, what);
}

static void pick_tile(int n, int ho, int wo, int& bn, int& bh, int& bw, int tile_px = 128) {
  long long best = -1;
  bn = 1;
  bh = 8;
  bw = 16;
  for (int cw = 1; cw <= tile_px; cw *= 2) {
    for (int chh = 1; cw * chh <= tile_px; chh *= 2) {
      const int cn = tile_px / (cw * chh);
      if (cn > 256) continue;
      const long long cost = static_cast<long long>((wo + cw - 1) / cw) * ((ho + chh - 1) / chh) * ((n + cn - 1) / cn);
      // prefer fewer tiles, then wider rows (TMA efficiency), then fewer images per tile
      const bool better = best < 0 || cost < best || (cost == best && (cw > bw || (cw == bw && chh > bh)));
      if (better) {
        best = cost;
        bn = cn;
        bh = chh;
        bw = cw;
      }
    }
  }
}


}  // namespace up
