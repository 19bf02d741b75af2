// contours.cu — cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) on the device
// (replaces /root/reference/chips/chips.py:382-386).
//
// All the logic lives in contours_core.h as per-index passes; this file runs each pass as one
// grid-stride kernel on the caller's stream.  HBM-bound integer work over the border pixels only
// (a few 1e5 of the 1.7e7 pixels of a 4096^2 image): the full image is read once by the flagging
// pass and once more (neighbour bits) by the scatter pass, everything after touches 8 states per
// border pixel.  Grid = a multiple of the SM count, 256 threads per CTA.
#include "common.cuh"
#include "contours_core.h"

namespace chips {

template <class F>
__global__ void __launch_bounds__(256) k_contour_pass(F f, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

struct CudaExec {
  cudaStream_t stream;
  int launches = 0;
  bool failed = false;
  template <class F>
  void each(long n, const F& f) {
    if (n <= 0 || failed) return;
    long blocks = (n + 255) / 256;
    const long max_blocks = (long)kNumSM * 32;
    if (blocks > max_blocks) blocks = max_blocks;
    k_contour_pass<F><<<(unsigned)blocks, 256, 0, stream>>>(f, n);
    if (cudaGetLastError() != cudaSuccess) failed = true;
    ++launches;
  }
};

}  // namespace chips

static_assert(sizeof(chips_contour_rec) == sizeof(ctr::Rec), "record layout");
static_assert(sizeof(chips_contour_counts) == sizeof(ctr::Counts), "counts layout");

extern "C" size_t chips_contours_workspace_bytes(int w, int h, long long max_border_px) {
  if (w < 1 || h < 1 || max_border_px < 1) return 0;
  ctr::Ws ws;
  return ctr::carve(ws, nullptr, w, h, (long)max_border_px);
}

extern "C" int chips_contours(const uint8_t* img, int w, int h, int pitch, int mode, int thr,
                              long long max_border_px, int start_rounds, chips_contour_rec* recs,
                              long long max_contours, int32_t* points, long long max_points,
                              chips_contour_counts* counts, void* ws, size_t ws_bytes, void* stream) {
  if (!img || !recs || !points || !counts || !ws || w < 1 || h < 1 || pitch < w || (mode != 0 && mode != 1) ||
      max_border_px < 1 || max_border_px >= (1ll << 28) || (long long)w * h >= (1ll << 30) ||
      start_rounds < 2 || max_contours < 1 || max_points < 1)
    return CHIPS_E_ARG;
  ctr::Ws cw;
  if (ctr::carve(cw, (char*)ws, w, h, (long)max_border_px) > ws_bytes) return CHIPS_E_WS;
  ctr::Args a{img, w, h, pitch, mode, thr, (long)max_border_px, start_rounds, (ctr::Rec*)recs,
              (long)max_contours, points, (long)max_points, (ctr::Counts*)counts};
  chips::CudaExec ex{(cudaStream_t)stream};
  ctr::run(ex, a, cw);
  chips::count_launch(ex.launches);
  return ex.failed ? CHIPS_E_CUDA : CHIPS_OK;
}
