// median.cu — exact zero-padded k x k median filter.
//
// Replaces scipy.signal.medfilt2d at /root/reference/chips/chips.py:214-216 and
// /root/reference/chips/syn_chips.py:106.  A median is a pure order statistic, so any
// exact selection reproduces scipy to the bit.  The B200 design is a two-tier selection:
//
//   1. k_bounds      one CTA: 4096 samples of the ROI, bitonic sort in shared memory,
//                    255 quantile boundaries  ->  256 value buckets.
//   2. k_bucketize   image -> padded uint8 bucket image (zero padding = bucket of 0.0).
//   3. k_huang       tier 1: every thread slides its own k x k window down one image
//                    column (Huang's O(k) update: +k / -k bucket counts per step) with a
//                    thread-private 256-bin histogram kept in shared memory in a
//                    bank-conflict-free interleaved layout.  Output per pixel: the bucket
//                    that holds the median and the residual rank inside that bucket.
//   4. k_refine      tier 2: one CTA per 16x16 output tile gathers, from the tile's
//                    (16+k-1)^2 input region, only the pixels that fall in a bucket some
//                    pixel of the tile needs, sorts them by value in shared memory
//                    (bitonic), and every thread walks its bucket's sorted segment counting
//                    in-window candidates until it reaches its residual rank.  The fused
//                    epilogue applies the disk mask and reduces min/max for the histogram.
//
// k_median_brute is an independent per-thread radix-select used only as the at-scale
// checker in the tests.
#include <math.h>
#include "common.cuh"

namespace chips {

constexpr int kBuckets = 256;
constexpr int kSamples = 4096;
constexpr int kHuangThreads = 128;
constexpr int kTile = 16;

// ------------------------------------------------------------------ shared bitonic sort
template <typename K, bool HAS_POS>
__device__ __forceinline__ void bitonic_sort(K* key, uint16_t* pos, int n, int tid, int nthreads) {
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (n >> 1); t += nthreads) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        K a = key[i], b = key[l];
        bool up = (i & k2) == 0;
        if ((a > b) == up) {
          key[i] = b;
          key[l] = a;
          if (HAS_POS) {
            uint16_t pa = pos[i];
            pos[i] = pos[l];
            pos[l] = pa;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ 1. boundaries
template <typename T>
__global__ void __launch_bounds__(1024) k_bounds(const T* __restrict__ img, int W, int rx0, int ry0,
                                                 int rw, int rh,
                                                 typename Key<T>::type* __restrict__ bounds,
                                                 typename Key<T>::type* __restrict__ minmax) {
  using K = typename Key<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* s = reinterpret_cast<K*>(smem_raw);
  for (int i = threadIdx.x; i < kSamples; i += blockDim.x) {
    int sy = i >> 6, sx = i & 63;
    int y = ry0 + (int)(((long long)(2 * sy + 1) * rh) >> 7);
    int x = rx0 + (int)(((long long)(2 * sx + 1) * rw) >> 7);
    s[i] = Key<T>::enc(img[(size_t)y * W + x]);
  }
  __syncthreads();
  bitonic_sort<K, false>(s, nullptr, kSamples, threadIdx.x, blockDim.x);
  for (int i = threadIdx.x; i < kBuckets; i += blockDim.x)
    bounds[i] = (i < kBuckets - 1) ? s[(i + 1) * (kSamples / kBuckets)] : Key<T>::kMax;
  if (threadIdx.x == 0) {
    minmax[0] = Key<T>::kMax;   // running min
    minmax[1] = 0;              // running max
  }
}

template <typename K>
__device__ __forceinline__ int bucket_of(const K* __restrict__ bnd, K key) {
  // number of boundaries bnd[0..254] that are <= key  (bnd sorted ascending)
  int b = 0;
#pragma unroll
  for (int step = 128; step > 0; step >>= 1)
    if (bnd[b + step - 1] <= key) b += step;
  return b;
}

// ------------------------------------------------------------------ 2. bucket image
template <typename T>
__global__ void __launch_bounds__(256) k_bucketize(const T* __restrict__ img, int H, int W, int half,
                                                   int pitch, int prow0, int pcol0_w, int ncol_w,
                                                   const typename Key<T>::type* __restrict__ bounds,
                                                   uint8_t* __restrict__ Bp) {
  using K = typename Key<T>::type;
  __shared__ K sb[kBuckets];
  for (int i = threadIdx.x; i < kBuckets; i += blockDim.x) sb[i] = bounds[i];
  __syncthreads();
  int cw = blockIdx.x * blockDim.x + threadIdx.x;   // word column inside the span
  if (cw >= ncol_w) return;
  int py = prow0 + blockIdx.y;
  int px = (pcol0_w + cw) * 4;
  int y = py - half;
  uint32_t word = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int x = px + i - half;
    T v = (y >= 0 && y < H && x >= 0 && x < W) ? img[(size_t)y * W + x] : (T)0;
    word |= (uint32_t)bucket_of<K>(sb, Key<T>::enc(v)) << (8 * i);
  }
  *reinterpret_cast<uint32_t*>(Bp + (size_t)py * pitch + px) = word;
}

// ------------------------------------------------------------------ 3. tier 1: Huang
// hist layout (uint16 units): bin * NT + slot(tid), slot = lane*2 + (warp&1) + (warp>>1)*64,
// so that lane l of every warp always addresses bank l (NT/2 is a multiple of 32).
__device__ __forceinline__ int huang_slot(int tid) {
  int lane = tid & 31, warp = tid >> 5;
  return lane * 2 + (warp & 1) + (warp >> 1) * 64;
}

template <int DELTA>
__device__ __forceinline__ void huang_row(uint16_t* __restrict__ my, const uint8_t* __restrict__ row,
                                          int k, int med, int& lt) {
  constexpr int NT = kHuangThreads;
  uintptr_t a = reinterpret_cast<uintptr_t>(row);
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  int sh = (int)(a & 3) * 8;
  uint32_t w0 = __ldg(wp);
  int j = 0;
  for (; j + 4 <= k; j += 4) {
    uint32_t w1 = __ldg(wp + (j >> 2) + 1);
    uint32_t v = __funnelshift_r(w0, w1, sh);
    w0 = w1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int b = (v >> (8 * i)) & 255;
      uint16_t* c = my + b * NT;
      *c = (uint16_t)(*c + DELTA);
      lt += (b < med) ? DELTA : 0;
    }
  }
  if (j < k) {
    uint32_t w1 = __ldg(wp + (j >> 2) + 1);
    uint32_t v = __funnelshift_r(w0, w1, sh);
    for (int i = 0; i < k - j; ++i) {
      int b = (v >> (8 * i)) & 255;
      uint16_t* c = my + b * NT;
      *c = (uint16_t)(*c + DELTA);
      lt += (b < med) ? DELTA : 0;
    }
  }
}

__global__ void __launch_bounds__(kHuangThreads) k_huang(const uint8_t* __restrict__ Bp, int pitch,
                                                         int half, int k, int W, int rx0, int ry0,
                                                         int rw, int rh, int L, int cx, int cy, int r,
                                                         uint8_t* __restrict__ bstar,
                                                         uint16_t* __restrict__ rres) {
  constexpr int NT = kHuangThreads;
  extern __shared__ __align__(16) uint16_t hist[];
  const int tid = threadIdx.x;
  const int x = rx0 + blockIdx.x * NT + tid;
  int ya = ry0 + blockIdx.y * L;
  int yb = min(ya + L, ry0 + rh);
  bool valid = x < rx0 + rw;
  if (valid && r >= 0) {
    long long dx = x - cx;
    long long rem = (long long)r * r - dx * dx;
    if (rem < 0) {
      valid = false;
    } else {
      long long dy = (long long)sqrt((double)rem);
      while ((dy + 1) * (dy + 1) <= rem) ++dy;
      while (dy * dy > rem) --dy;
      ya = max(ya, cy - (int)dy);
      yb = min(yb, cy + (int)dy + 1);
    }
  }
  if (!valid || ya >= yb) return;
  uint16_t* my = hist + huang_slot(tid);
#pragma unroll 8
  for (int b = 0; b < kBuckets; ++b) my[b * NT] = 0;
  const int R = (k * k - 1) >> 1;
  int med = 0, lt = 0;
  // padded coordinates: image (y, x) lives at Bp[(y+half)*pitch + x+half]; the window of
  // centre column x starts at padded column x.
  const uint8_t* col = Bp + x;
  for (int yy = ya - half; yy < yb + half; ++yy) {
    huang_row<+1>(my, col + (size_t)(yy + half) * pitch, k, med, lt);
    int yo = yy - k;
    if (yo >= ya - half) huang_row<-1>(my, col + (size_t)(yo + half) * pitch, k, med, lt);
    int y = yy - half;
    if (y >= ya) {
      while (lt > R) {
        --med;
        lt -= my[med * NT];
      }
      while (true) {
        int c = my[med * NT];
        if (lt + c > R) break;
        lt += c;
        ++med;
      }
      size_t o = (size_t)y * W + x;
      bstar[o] = (uint8_t)med;
      rres[o] = (uint16_t)(R - lt);
    }
  }
}

// ------------------------------------------------------------------ 4. tier 2: tile refine
template <typename T>
__global__ void __launch_bounds__(kTile* kTile) k_refine(
    const T* __restrict__ img, const uint8_t* __restrict__ Bp, int pitch,
    const uint8_t* __restrict__ bstar, const uint16_t* __restrict__ rres,
    const typename Key<T>::type* __restrict__ bounds, int H, int W, int half, int rx0, int ry0,
    int rw, int rh, int cx, int cy, int r, int cap, T* __restrict__ out,
    typename Key<T>::type* __restrict__ minmax, int* __restrict__ err) {
  using K = typename Key<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  K* ckey = reinterpret_cast<K*>(smem_raw);
  uint16_t* cpos = reinterpret_cast<uint16_t*>(smem_raw + (size_t)cap * sizeof(K));
  __shared__ uint32_t need[8];
  __shared__ int ccount;
  __shared__ int nactive;

  const int tid = threadIdx.x;
  const int tx = tid & (kTile - 1), ty = tid >> 4;
  const int X0 = rx0 + blockIdx.x * kTile, Y0 = ry0 + blockIdx.y * kTile;
  const int x = X0 + tx, y = Y0 + ty;
  const bool active = (x < rx0 + rw) && (y < ry0 + rh) && on_disk(x, y, cx, cy, r);
  if (tid < 8) need[tid] = 0;
  if (tid == 0) {
    ccount = 0;
    nactive = 0;
  }
  __syncthreads();
  int bs = 0, rr = 0;
  if (active) {
    size_t o = (size_t)y * W + x;
    bs = bstar[o];
    rr = rres[o];
    atomicOr(&need[bs >> 5], 1u << (bs & 31));
    nactive = 1;
  }
  __syncthreads();
  if (nactive == 0) return;

  // gather candidates of the needed buckets from the (kTile+2*half)^2 input region
  const int RS = kTile + 2 * half;
  const int lane = tid & 31;
  for (int ryb = 0; ryb < RS; ryb += kTile) {      // uniform trip counts: ballots below
    const int ry = ryb + ty;
    const uint8_t* brow = Bp + (size_t)(Y0 + min(ry, RS - 1)) * pitch + X0;
    for (int rxb = 0; rxb < RS; rxb += kTile) {
      const int rx = rxb + tx;
      bool hit = false;
      if (rx < RS && ry < RS) {
        int b = brow[rx];
        hit = (need[b >> 5] >> (b & 31)) & 1u;
      }
      unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
      if (m) {
        int base = 0;
        int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(&ccount, __popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (hit) {
          int slot = base + __popc(m & ((1u << lane) - 1));
          int iy = Y0 + ry - half, ix = X0 + rx - half;
          T v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? img[(size_t)iy * W + ix] : (T)0;
          if (slot < cap) {
            ckey[slot] = Key<T>::enc(v);
            cpos[slot] = (uint16_t)((ry << 8) | rx);
          }
        }
      }
    }
  }
  __syncthreads();
  const int C = min(ccount, cap);
  int n2 = 2;
  while (n2 < C) n2 <<= 1;
  for (int i = C + tid; i < n2; i += kTile * kTile) {
    ckey[i] = Key<T>::kMax;
    cpos[i] = 0xFFFF;
  }
  __syncthreads();
  bitonic_sort<K, true>(ckey, cpos, n2, tid, kTile * kTile);

  K kmin = Key<T>::kMax, kmax = 0;
  if (active) {
    K lowkey = (bs == 0) ? (K)0 : bounds[bs - 1];
    int lo = 0, hi = C;   // first index with ckey >= lowkey
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (ckey[mid] < lowkey) lo = mid + 1; else hi = mid;
    }
    const unsigned span = 2u * (unsigned)half;
    int cnt = 0, found = -1;
    for (int i = lo; i < C; ++i) {
      unsigned p = cpos[i];
      if ((p >> 8) - (unsigned)ty <= span && (p & 255u) - (unsigned)tx <= span) {
        if (cnt == rr) {
          found = i;
          break;
        }
        ++cnt;
      }
    }
    K res = 0;
    if (found >= 0) {
      res = ckey[found];
    } else {
      atomicAdd(err, 1);
      res = Key<T>::enc((T)0);
    }
    out[(size_t)y * W + x] = Key<T>::dec(res);
    kmin = res;
    kmax = res;
  }
  // min/max of the masked output for the histogram range (np.histogram's a.min()/a.max())
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    K a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o);
    K b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if (lane == 0 && kmin <= kmax) {
    atomic_min_key(&minmax[0], kmin);
    atomic_max_key(&minmax[1], kmax);
  }
}

// ------------------------------------------------------------------ brute-force checker
template <typename T>
__global__ void __launch_bounds__(kHuangThreads) k_median_brute(const T* __restrict__ img,
                                                                T* __restrict__ out, int H, int W,
                                                                int k, int cx, int cy, int r) {
  using K = typename Key<T>::type;
  constexpr int NT = kHuangThreads;
  extern __shared__ __align__(16) uint16_t hist[];
  const int tid = threadIdx.x;
  const int x = blockIdx.x * NT + tid, y = blockIdx.y;
  if (x >= W) return;
  if (!on_disk(x, y, cx, cy, r)) {
    out[(size_t)y * W + x] = (T)0;
    return;
  }
  uint16_t* my = hist + huang_slot(tid);
  const int half = k >> 1;
  int rank = (k * k - 1) >> 1;
  K prefix = 0, mask = 0;
  for (int shift = (int)sizeof(K) * 8 - 8; shift >= 0; shift -= 8) {
    for (int b = 0; b < 256; ++b) my[b * NT] = 0;
    for (int dy = -half; dy <= half; ++dy) {
      int iy = y + dy;
      for (int dx = -half; dx <= half; ++dx) {
        int ix = x + dx;
        T v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? img[(size_t)iy * W + ix] : (T)0;
        K key = Key<T>::enc(v);
        if ((key & mask) == prefix) {
          int b = (int)((key >> shift) & 255);
          my[b * NT] = (uint16_t)(my[b * NT] + 1);
        }
      }
    }
    int b = 0;
    while (rank >= (int)my[b * NT]) {
      rank -= my[b * NT];
      ++b;
    }
    prefix |= (K)b << shift;
    mask |= (K)255 << shift;
  }
  out[(size_t)y * W + x] = Key<T>::dec(prefix);
}

template <typename T>
__global__ void k_copy_masked(const T* __restrict__ in, T* __restrict__ out, int H, int W, int cx,
                              int cy, int r) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * W) return;
  int y = (int)(i / W), x = (int)(i % W);
  out[i] = on_disk(x, y, cx, cy, r) ? in[i] : (T)0;
}

// ------------------------------------------------------------------ host-side launchers
static int refine_cap(int k) {
  int rs = kTile + (k - 1);
  int n = rs * rs, c = 2;
  while (c < n) c <<= 1;
  return c;
}

template <typename T>
static int medfilt_impl(const T* in, T* out, const chips_plan* p, int cx, int cy, int r,
                        unsigned char* ws, cudaStream_t st, int stages) {
  using K = typename Key<T>::type;
  const int H = p->H, W = p->W, k = p->k, half = k >> 1;
  K* bounds = reinterpret_cast<K*>(ws + p->off_bounds);
  K* minmax = reinterpret_cast<K*>(ws + p->off_minmax);
  uint8_t* Bp = ws + p->off_bucket;
  uint8_t* bstar = ws + p->off_bstar;
  uint16_t* rres = reinterpret_cast<uint16_t*>(ws + p->off_rres);
  int* err = &reinterpret_cast<chips_result*>(ws + p->off_result)->median_errors;
  const int rx0 = p->roi_x0, ry0 = p->roi_y0, rw = p->roi_w, rh = p->roi_h;

  if (stages & 1) {
    CHIPS_CUDA(cudaMemsetAsync(out, 0, (size_t)H * W * sizeof(T), st));
    CHIPS_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    CHIPS_CUDA(cudaFuncSetAttribute(k_bounds<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSamples * sizeof(K))));
    k_bounds<T><<<1, 1024, kSamples * sizeof(K), st>>>(in, W, rx0, ry0, rw, rh, bounds, minmax);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (k == 1) {
    // medfilt2d with a 1x1 kernel is the identity (then masked); min/max via k_minmax later
    size_t n = (size_t)H * W;
    k_copy_masked<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, H, W, cx, cy, r);
    CHIPS_CHECK_LAUNCH();
    count_launch();
    return CHIPS_OK;
  }
  if (stages & 2) {  // bucket image over the ROI expanded by the halo (padded coordinates)
    int prow0 = ry0, prow1 = ry0 + rh + 2 * half;            // padded rows [ry0, ry0+rh+2h)
    int pcol0_w = rx0 / 4;
    int pcol1_w = (rx0 + rw + 2 * half + 8 + 3) / 4;         // + slack for the funnel loads
    if (pcol1_w * 4 > p->bp_pitch) pcol1_w = p->bp_pitch / 4;
    int ncol_w = pcol1_w - pcol0_w;
    dim3 grid((ncol_w + 255) / 256, prow1 - prow0);
    k_bucketize<T><<<grid, 256, 0, st>>>(in, H, W, half, p->bp_pitch, prow0, pcol0_w, ncol_w,
                                         bounds, Bp);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (stages & 4) {  // tier 1
    // run length: aim at >= 4 CTAs per SM worth of work, but keep the (k-1)-row warm-up
    // of every run amortised (L >= 2k).
    int nbx = (rw + kHuangThreads - 1) / kHuangThreads;
    int L = 2 * k;
    while (L < rh && (long long)nbx * ((rh + L - 1) / L) > 8LL * kNumSM) L += k;
    if (L > rh) L = rh;
    dim3 grid(nbx, (rh + L - 1) / L);
    size_t smem = (size_t)kBuckets * kHuangThreads * sizeof(uint16_t);
    CHIPS_CUDA(cudaFuncSetAttribute(k_huang, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_huang<<<grid, kHuangThreads, smem, st>>>(Bp, p->bp_pitch, half, k, W, rx0, ry0, rw, rh, L, cx,
                                               cy, r, bstar, rres);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (stages & 8) {  // tier 2
    int cap = refine_cap(k);
    size_t smem = (size_t)cap * (sizeof(K) + sizeof(uint16_t));
    CHIPS_CUDA(cudaFuncSetAttribute(k_refine<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((rw + kTile - 1) / kTile, (rh + kTile - 1) / kTile);
    k_refine<T><<<grid, kTile * kTile, smem, st>>>(in, Bp, p->bp_pitch, bstar, rres, bounds, H, W,
                                                   half, rx0, ry0, rw, rh, cx, cy, r, cap, out,
                                                   minmax, err);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  return CHIPS_OK;
}

}  // namespace chips

extern "C" int chips_medfilt2d_stages(const void* in, void* out, const chips_plan* plan, int cx,
                                      int cy, int r, void* ws, void* stream, int stages);

extern "C" int chips_medfilt2d(const void* in, void* out, const chips_plan* plan, int cx, int cy,
                               int r, void* ws, void* stream) {
  return chips_medfilt2d_stages(in, out, plan, cx, cy, r, ws, stream, 15);
}

extern "C" int chips_medfilt2d_stages(const void* in, void* out, const chips_plan* plan, int cx,
                                      int cy, int r, void* ws, void* stream, int stages) {
  if (!in || !out || !plan || !ws) return CHIPS_E_ARG;
  if (plan->k < 1 || !(plan->k & 1) || plan->k > CHIPS_MAX_K) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->elem == 4)
    return chips::medfilt_impl<float>((const float*)in, (float*)out, plan, cx, cy, r,
                                      (unsigned char*)ws, st, stages);
  if (plan->elem == 8)
    return chips::medfilt_impl<double>((const double*)in, (double*)out, plan, cx, cy, r,
                                       (unsigned char*)ws, st, stages);
  return CHIPS_E_ARG;
}

extern "C" int chips_medfilt2d_bruteforce(const void* in, void* out, int H, int W, int k, int elem,
                                          int cx, int cy, int r, void* stream) {
  if (!in || !out || k < 1 || !(k & 1)) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((W + chips::kHuangThreads - 1) / chips::kHuangThreads, H);
  size_t smem = (size_t)256 * chips::kHuangThreads * sizeof(uint16_t);
  if (elem == 4) {
    CHIPS_CUDA(cudaFuncSetAttribute(chips::k_median_brute<float>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chips::k_median_brute<float><<<grid, chips::kHuangThreads, smem, st>>>(
        (const float*)in, (float*)out, H, W, k, cx, cy, r);
  } else if (elem == 8) {
    CHIPS_CUDA(cudaFuncSetAttribute(chips::k_median_brute<double>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chips::k_median_brute<double><<<grid, chips::kHuangThreads, smem, st>>>(
        (const double*)in, (double*)out, H, W, k, cx, cy, r);
  } else {
    return CHIPS_E_ARG;
  }
  CHIPS_CHECK_LAUNCH();
  chips::count_launch();
  return CHIPS_OK;
}
