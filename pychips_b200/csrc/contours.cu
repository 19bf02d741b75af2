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

// the same over [0, min(n_max, mult * *n_ptr)): the count lives on the device
template <class F>
__global__ void __launch_bounds__(256) k_contour_pass_n(F f, long n_max, const int32_t* __restrict__ n_ptr,
                                                        int mult) {
  const long m = (long)mult * (long)*n_ptr;
  const long n = m < n_max ? m : n_max;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// each_n for a functor that returns "raise the flag": one store per CTA at most, and only while the
// flag still reads 0 (same-address stores from every thread serialise in L2: ~10 ns each)
template <class F>
__global__ void __launch_bounds__(256) k_contour_pass_flag(F f, long n_max, const int32_t* __restrict__ n_ptr,
                                                           int mult, int32_t* flag) {
  const long m = (long)mult * (long)*n_ptr;
  const long n = m < n_max ? m : n_max;
  const long stride = (long)gridDim.x * blockDim.x;
  bool any = false;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) any |= f(i);
  if (__syncthreads_or(any) && threadIdx.x == 0 && __ldcg(flag) == 0) *flag = 1;
}

// ---- exclusive prefix sums of count(i): tile sums, one CTA over the tile sums, tile scans ------
constexpr int kScanThreads = 256, kScanItems = ctr::kScanTile / kScanThreads;

__device__ __forceinline__ uint32_t warp_inclusive(uint32_t x, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
    if (lane >= o) x += y;
  }
  return x;
}

template <class Count>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(Count c, long n_max,
                                                                 const int32_t* __restrict__ n_ptr,
                                                                 uint32_t* __restrict__ sums) {
  __shared__ uint32_t part[kScanThreads / 32];
  const long n = n_ptr ? ((long)*n_ptr < n_max ? (long)*n_ptr : n_max) : n_max;
  const long base = (long)blockIdx.x * ctr::kScanTile;
  uint32_t s = 0;
#pragma unroll 4
  for (int k = 0; k < kScanItems; ++k) {
    const long i = base + k * kScanThreads + threadIdx.x;       // coalesced: order is free in a sum
    if (i < n) s += c(i);
  }
  s = __reduce_add_sync(0xFFFFFFFFu, s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += part[w];
    sums[blockIdx.x] = t;
  }
}

template <class Total>
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* __restrict__ sums, int ntiles, Total total) {
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? sums[i] : 0u;
    const uint32_t x = warp_inclusive(v, lane);
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) wsum[lane] = warp_inclusive(wsum[lane], lane);
    __syncthreads();
    if (i < ntiles) sums[i] = carry + (warp ? wsum[warp - 1] : 0u) + x - v;
    carry += wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) total(carry);
}

template <class Count>
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_final(Count c, long n_max,
                                                                  const int32_t* __restrict__ n_ptr,
                                                                  const uint32_t* __restrict__ sums,
                                                                  uint32_t* __restrict__ out) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  const long n = n_ptr ? ((long)*n_ptr < n_max ? (long)*n_ptr : n_max) : n_max;
  const long base = (long)blockIdx.x * ctr::kScanTile + (long)threadIdx.x * kScanItems;
  if ((long)blockIdx.x * ctr::kScanTile >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t v[kScanItems], s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < n ? c(base + k) : 0u;
    s += v[k];
  }
  const uint32_t x = warp_inclusive(s, lane);
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  uint32_t run = sums[blockIdx.x] + x - s;
  for (int w = 0; w < warp; ++w) run += wsum[w];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

struct CudaExec {
  cudaStream_t stream;
  int launches = 0;
  bool failed = false;
  static unsigned grid_for(long n) {
    long blocks = (n + 255) / 256;
    const long max_blocks = (long)kNumSM * 8;
    return (unsigned)(blocks > max_blocks ? max_blocks : blocks);
  }
  void after() {
    if (cudaGetLastError() != cudaSuccess) failed = true;
    ++launches;
  }
  template <class F>
  void each(long n, const F& f) {
    if (n <= 0 || failed) return;
    k_contour_pass<F><<<grid_for(n), 256, 0, stream>>>(f, n);
    after();
  }
  template <class F>
  void each_n(long n_max, const int32_t* n_ptr, int mult, const F& f) {
    if (n_max <= 0 || failed) return;
    k_contour_pass_n<F><<<grid_for(n_max), 256, 0, stream>>>(f, n_max, n_ptr, mult);
    after();
  }
  template <class F>
  void each_flag(long n_max, const int32_t* n_ptr, int mult, const F& f, int32_t* flag) {
    if (n_max <= 0 || failed) return;
    k_contour_pass_flag<F><<<grid_for(n_max), 256, 0, stream>>>(f, n_max, n_ptr, mult, flag);
    after();
  }
  template <class Count, class Total>
  void scan(Count c, long n_max, const int32_t* n_ptr, uint32_t* out, uint32_t* sums, Total total) {
    if (n_max <= 0 || failed) return;
    const int tiles = (int)((n_max + ctr::kScanTile - 1) / ctr::kScanTile);
    k_scan_tile_sums<Count><<<tiles, kScanThreads, 0, stream>>>(c, n_max, n_ptr, sums);
    after();
    k_scan_top<Total><<<1, 1024, 0, stream>>>(sums, tiles, total);
    after();
    k_scan_tile_final<Count><<<tiles, kScanThreads, 0, stream>>>(c, n_max, n_ptr, sums, out);
    after();
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
      start_rounds < 2 || start_rounds > ctr::kMaxStartRounds || max_contours < 1 || max_points < 1)
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
