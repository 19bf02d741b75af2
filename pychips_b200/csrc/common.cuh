// common.cuh — shared device helpers for the CHIPS sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/chips_b200.h"

#define CHIPS_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return CHIPS_E_CUDA;               \
  } while (0)

#define CHIPS_CUDA(call)                                       \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return CHIPS_E_CUDA;               \
  } while (0)

// -DCHIPS_DEBUG_BOUNDS (python -m pychips_b200.build with CHIPS_EXTRA_NVCC_FLAGS): index asserts on the
// union-find, list and candidate buffers; a violation prints the site and traps (the call then
// fails with a CUDA error).  compute-sanitizer is not available on this pool: the full GPU suite
// is run once against such a build instead (profiles/r02_debug_bounds.txt).
#ifdef CHIPS_DEBUG_BOUNDS
#include <stdio.h>
#define CHIPS_BOUNDS(cond)                                                                   \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      printf("CHIPS_BOUNDS failed: %s  (%s:%d)\n", #cond, __FILE__, __LINE__);               \
      __trap();                                                                              \
    }                                                                                        \
  } while (0)
#else
#define CHIPS_BOUNDS(cond) do { } while (0)
#endif

namespace chips {

// number of kernels this library has launched (bench.py reports it as gpu_launches)
extern std::atomic<long long> g_launches;
static inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// SMs of the current device (148 on B200: 2 dies x 74), queried once per thread; grids of the
// grid-stride kernels are multiples of it
static inline int num_sms() {
  static thread_local int cached = 0;
  if (cached <= 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
  }
  return cached;
}
#define kNumSM (::chips::num_sms())

// ---- order-preserving integer keys of IEEE floats (NaN-free input: the median's input pass
//      counts non-finite pixels in chips_result.n_nonfinite and the host classes raise) ----------
template <typename T> struct Key;
template <> struct Key<float> {
  using type = uint32_t;
  static __device__ __forceinline__ uint32_t enc(float v) {
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
  static __device__ __forceinline__ float dec(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
  }
  static constexpr uint32_t kMax = 0xFFFFFFFFu;
};
template <> struct Key<double> {
  using type = uint64_t;
  static __device__ __forceinline__ uint64_t enc(double v) {
    uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
  }
  static __device__ __forceinline__ double dec(uint64_t k) {
    uint64_t u = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
  }
  static constexpr uint64_t kMax = 0xFFFFFFFFFFFFFFFFull;
};

// NaN / +-Inf: all exponent bits set
__device__ __forceinline__ bool nonfinite(float v) { return (__float_as_uint(v) & 0x7F800000u) == 0x7F800000u; }
__device__ __forceinline__ bool nonfinite(double v) {
  return ((uint32_t)((uint64_t)__double_as_longlong(v) >> 32) & 0x7FF00000u) == 0x7FF00000u;
}

// filled cv2.circle == integer test
__device__ __forceinline__ bool on_disk(int x, int y, int cx, int cy, int r) {
  if (r < 0) return true;
  const unsigned ax = (unsigned)abs(x - cx), ay = (unsigned)abs(y - cy);
  // exact in 32 bits for every image up to 32768 pixels on a side (2 * 32767^2 < 2^32)
  if ((ax | ay) < 32768u && (unsigned)r < 46341u) return ax * ax + ay * ay <= (unsigned)r * (unsigned)r;
  long long dx = x - cx, dy = y - cy;
  return dx * dx + dy * dy <= (long long)r * (long long)r;
}

// on-disk columns [xa, xb) of image row y inside the ROI columns [rx0, rx0+rw); false if none
__device__ __forceinline__ bool row_chord(int y, int cx, int cy, int r, int rx0, int rw, int& xa,
                                          int& xb) {
  xa = rx0;
  xb = rx0 + rw;
  if (r < 0) return true;
  long long dy = y - cy;
  long long rem = (long long)r * r - dy * dy;
  if (rem < 0) return false;
  long long dx = (long long)sqrt((double)rem);
  while ((dx + 1) * (dx + 1) <= rem) ++dx;
  while (dx * dx > rem) --dx;
  xa = max(xa, cx - (int)dx);
  xb = min(xb, cx + (int)dx + 1);
  return xa < xb;
}

// on-disk columns [xa, xb) of row y, computed by lane 0 of the warp (an fp64 square root) and
// broadcast; xa = xb when the row misses the disk
__device__ __forceinline__ void warp_row_chord(int y, int cx, int cy, int r, int rx0, int rw, int& xa, int& xb) {
  xa = 0;
  xb = 0;
  if ((threadIdx.x & 31) == 0 && !row_chord(y, cx, cy, r, rx0, rw, xa, xb)) xa = xb = 0;
  xa = __shfl_sync(0xFFFFFFFFu, xa, 0);
  xb = __shfl_sync(0xFFFFFFFFu, xb, 0);
}

// np.linspace edge i (fp64, multiply THEN add, never fused), last edge forced to `last`
__device__ __forceinline__ double linspace_edge(int i, int nbins, double first, double last,
                                                double step) {
  if (i >= nbins) return last;
  if (step == 0.0) {   // numpy's denormal branch: (i / div) * delta + start
    double t = __ddiv_rn((double)i, (double)nbins);
    t = __dmul_rn(t, __dsub_rn(last, first));
    return __dadd_rn(t, first);
  }
  return __dadd_rn(__dmul_rn((double)i, step), first);
}

// Bin edges of np.histogram(bins=n): np.linspace(first, last, n + 1) in the dtype of the data.
// float64 data (and everything widened to it): fp64.  float32 DATA (f32 = 1: numpy keeps
// bin_type = float32, numpy/lib/_histograms_impl.py `_get_bin_edges`): every operation of
// linspace in float32 - delta = last - first, step = delta / n, edge_i = i * step + first, last
// edge forced - and the values returned widened (exactly) to double.
struct Edges {
  int nbins, f32;
  double first, last, step;
  float ffirst, flast, fstep, fdelta;
  __device__ __forceinline__ double operator()(int i) const {
    if (!f32) return linspace_edge(i, nbins, first, last, step);
    if (i >= nbins) return (double)flast;
    const float y = (fstep == 0.0f) ? __fmul_rn(__fdiv_rn((float)i, (float)nbins), fdelta)
                                    : __fmul_rn((float)i, fstep);
    return (double)__fadd_rn(y, ffirst);
  }
  // np.diff(edges)[i] as float64 (numpy subtracts in the edge dtype, then converts)
  __device__ __forceinline__ double width(double e0, double e1) const {
    return f32 ? (double)__fsub_rn((float)e1, (float)e0) : __dsub_rn(e1, e0);
  }
  // edges[:-1] + np.diff(edges) in the edge dtype
  __device__ __forceinline__ double upper(double e0, double db) const {
    return f32 ? (double)__fadd_rn((float)e0, (float)db) : __dadd_rn(e0, db);
  }
};

// outer edges of a selection with minimum `first` and maximum `last` (np.histogram's
// _get_outer_edges: a.min(), a.max(), widened by 0.5 on both sides when they are equal; (0, 1)
// for an empty selection: empty = true)
__device__ __forceinline__ Edges make_edges(double first, double last, bool empty, int nbins, int f32) {
  Edges e;
  e.nbins = nbins;
  e.f32 = f32;
  if (empty) {
    first = 0.0;
    last = 1.0;
  }
  if (first == last) {
    if (f32) {
      first = (double)__fsub_rn((float)first, 0.5f);
      last = (double)__fadd_rn((float)last, 0.5f);
    } else {
      first = first - 0.5;
      last = last + 0.5;
    }
  }
  e.first = first;
  e.last = last;
  e.step = __ddiv_rn(__dsub_rn(last, first), (double)nbins);
  e.ffirst = (float)first;
  e.flast = (float)last;
  e.fdelta = __fsub_rn(e.flast, e.ffirst);
  e.fstep = __fdiv_rn(e.fdelta, (float)nbins);
  return e;
}

__device__ __forceinline__ void atomic_min_key(uint32_t* p, uint32_t v) { atomicMin(p, v); }
__device__ __forceinline__ void atomic_max_key(uint32_t* p, uint32_t v) { atomicMax(p, v); }
__device__ __forceinline__ void atomic_min_key(uint64_t* p, uint64_t v) {
  atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ void atomic_max_key(uint64_t* p, uint64_t v) {
  atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace chips
