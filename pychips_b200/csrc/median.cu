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
#include <cuda.h>
#include <cuda_pipeline.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
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

// Streams the k bucket bytes that start at `row` (any alignment) through aligned 32-bit
// loads + funnel shifts.
struct RowReader {
  const uint32_t* wp;
  int sh;
  uint32_t w0;
  __device__ __forceinline__ RowReader(const uint8_t* row) {
    uintptr_t a = reinterpret_cast<uintptr_t>(row);
    wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    sh = (int)(a & 3) * 8;
    w0 = __ldg(wp);
  }
  __device__ __forceinline__ uint32_t next(int word) {   // bytes [4*word, 4*word+4)
    uint32_t w1 = __ldg(wp + word + 1);
    uint32_t v = __funnelshift_r(w0, w1, sh);
    w0 = w1;
    return v;
  }
};

// number of the 4 bytes of w that are < m (m in 0..255), branch-free SWAR: the low 7 bits are
// compared through a borrow-free subtraction, the top bits by boolean algebra.
__device__ __forceinline__ int count_bytes_below(uint32_t w, uint32_t m4, uint32_t m4lo) {
  const uint32_t Hm = 0x80808080u;
  uint32_t t = ((w & ~Hm) | Hm) - m4lo;           // bit 7 of each byte: (w&0x7f) >= (m&0x7f)
  uint32_t lt = ((~w & m4) | (~(w ^ m4) & ~t)) & Hm;
  return __popc(lt);
}

// +1 for two entering elements; the two read-modify-writes are issued together (the loads
// are independent); if both hit the same bin the second store (program order) carries the +2.
__device__ __forceinline__ void huang_add2(uint16_t* __restrict__ my, int a, int b) {
  constexpr int NT = kHuangThreads;
  uint16_t* pa = my + a * NT;
  uint16_t* pb = my + b * NT;
  uint16_t ca = *pa, cb = *pb;
  uint16_t na = (uint16_t)(ca + 1);
  *pa = na;
  *pb = (uint16_t)((a == b ? na : cb) + 1);
}
__device__ __forceinline__ void huang_add1(uint16_t* __restrict__ my, int a) {
  uint16_t* pa = my + a * kHuangThreads;
  *pa = (uint16_t)(*pa + 1);
}
// +1 for the entering element a, -1 for the leaving element b (a == b: net zero, the second
// store writes the original count back)
__device__ __forceinline__ void huang_swap(uint16_t* __restrict__ my, int a, int b) {
  constexpr int NT = kHuangThreads;
  uint16_t* pa = my + a * NT;
  uint16_t* pb = my + b * NT;
  uint16_t ca = *pa, cb = *pb;
  uint16_t na = (uint16_t)(ca + 1);
  *pa = na;
  *pb = (uint16_t)((a == b ? na : cb) - 1);
}

__device__ __forceinline__ void huang_add_row(uint16_t* __restrict__ my, const uint8_t* row, int k,
                                              int med, int& lt) {
  RowReader rd(row);
  const uint32_t m4 = (uint32_t)med * 0x01010101u, m4lo = m4 & 0x7F7F7F7Fu;
  int j = 0, acc = 0;
  for (; j + 4 <= k; j += 4) {
    uint32_t v = rd.next(j >> 2);
    acc += count_bytes_below(v, m4, m4lo);
    huang_add2(my, v & 255, (v >> 8) & 255);
    huang_add2(my, (v >> 16) & 255, v >> 24);
  }
  if (j < k) {
    uint32_t v = rd.next(j >> 2);
    acc += count_bytes_below(v | (0xFFFFFFFFu << (8 * (k - j))), m4, m4lo);
    for (int i = 0; i < k - j; ++i) huang_add1(my, (v >> (8 * i)) & 255);
  }
  lt += acc;
}

__device__ __forceinline__ void huang_swap_row(uint16_t* __restrict__ my, const uint8_t* add,
                                               const uint8_t* rem, int k, int med, int& lt) {
  RowReader ra(add), rb(rem);
  const uint32_t m4 = (uint32_t)med * 0x01010101u, m4lo = m4 & 0x7F7F7F7Fu;
  int j = 0, acc = 0;
  for (; j + 4 <= k; j += 4) {
    uint32_t va = ra.next(j >> 2), vb = rb.next(j >> 2);
    acc += count_bytes_below(va, m4, m4lo) - count_bytes_below(vb, m4, m4lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) huang_swap(my, (va >> (8 * i)) & 255, (vb >> (8 * i)) & 255);
  }
  if (j < k) {
    uint32_t va = ra.next(j >> 2), vb = rb.next(j >> 2);
    const uint32_t inval = 0xFFFFFFFFu << (8 * (k - j));
    acc += count_bytes_below(va | inval, m4, m4lo) - count_bytes_below(vb | inval, m4, m4lo);
    for (int i = 0; i < k - j; ++i) huang_swap(my, (va >> (8 * i)) & 255, (vb >> (8 * i)) & 255);
  }
  lt += acc;
}

__global__ void __launch_bounds__(kHuangThreads) k_huang(const uint8_t* __restrict__ Bp, int pitch,
                                                         int half, int k, int W, int rx0, int ry0,
                                                         int rw, int rh, int L, int cx, int cy, int r,
                                                         uint8_t* __restrict__ bstar,
                                                         uint32_t* __restrict__ rres) {
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
  // centre column x starts at padded column x, the window row of image row yy is padded row
  // yy + half.
  const uint8_t* col = Bp + x;
  auto prow = [&](int yy) { return col + (size_t)(yy + half) * pitch; };
  for (int yy = ya - half; yy < ya + half; ++yy) huang_add_row(my, prow(yy), k, med, lt);
  for (int y = ya; y < yb; ++y) {
    if (y == ya) huang_add_row(my, prow(y + half), k, med, lt);
    else huang_swap_row(my, prow(y + half), prow(y - half - 1), k, med, lt);
    while (lt > R) {
      --med;
      lt -= my[med * NT];
    }
    int c;
    while (true) {
      c = my[med * NT];
      if (lt + c > R) break;
      lt += c;
      ++med;
    }
    size_t o = (size_t)y * W + x;
    bstar[o] = (uint8_t)med;
    rres[o] = (uint32_t)(R - lt) | ((uint32_t)c << 16);   // residual rank | bucket population
  }
}

// ------------------------------------------------------------------ 4. tier 2: tile refine
// One CTA per 16x16 output tile.
//   a. bitmap of the buckets the tile's pixels need (tier-1 output);
//   b. the tile's (16+k-1)^2 input region (bucket bytes + ordered keys) is staged in shared
//      memory with coalesced, mutually independent loads;
//   c. a counting sort collects the positions of the region pixels whose bucket is needed,
//      grouped by (bucket, leading digit of the key inside the bucket's key range);
//   d. each warp finishes the sort inside the small sub-groups (shuffle rank sort, in place);
//   e. every thread walks the sorted candidates of ITS bucket from the nearer end, counting
//      those inside its own k x k window, until it reaches its residual rank.
// Epilogue: masked store + min/max reduction (histogram range).
constexpr int kMaxSub = 2048;   // sub-group counters per tile: groups * digits <= kMaxSub

template <typename K>
__device__ __forceinline__ int bit_length(K v) {
  return (sizeof(K) == 8) ? 64 - __clzll((long long)v) : 32 - __clz((int)v);
}

__device__ __forceinline__ bool in_window(unsigned p, unsigned bias, unsigned win) {
  // p = (ry << 8) | rx, bias = (ty << 8) | tx.  A borrow out of the low byte (rx < tx) leaves a
  // low byte >= 256 - 15 > win, a negative row difference wraps the whole word: both rejected.
  unsigned t = p - bias;
  return ((t >> 8) <= win) & ((t & 255u) <= win);
}

template <typename T>
__device__ __forceinline__ void refine_tile_staged(
    const int tile_x, const int tile_y, uint32_t& tma_phase, const uint32_t bar,
    const T* __restrict__ img, const uint8_t* __restrict__ Bp, int pitch,
    const uint8_t* __restrict__ bstar, const uint32_t* __restrict__ rres,
    const typename Key<T>::type* __restrict__ bounds, int H, int W, int half, int rx0, int ry0,
    int rw, int rh, int cx, int cy, int r, int cap, T* __restrict__ out,
    typename Key<T>::type* __restrict__ minmax, int* __restrict__ err,
    int use_tma, const CUtensorMap* tm_img, const CUtensorMap* tm_bkt) {
  using K = typename Key<T>::type;
  constexpr int NT = kTile * kTile;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(128) unsigned char smem_tile[];
  K* skey = reinterpret_cast<K*>(smem_tile);                                    // [RS*vpitch] keys
  uint16_t* rsg = reinterpret_cast<uint16_t*>(smem_tile + (size_t)cap * sizeof(K));   // [RS*RS] sub-group of a region pixel
  uint16_t* cpos = rsg + cap;      // [C] candidate positions grouped by sub-group (unsorted)
  uint16_t* csg = cpos + cap;      // [C] sub-group of a candidate
  uint16_t* spos = rsg;            // [C] candidate positions, fully sorted (aliases rsg)
  __shared__ uint32_t need[8], needpfx[9];
  __shared__ uint16_t glut[kBuckets];          // bucket -> group (0xFFFF: not needed by this tile)
  __shared__ K gklo[kBuckets];
  __shared__ uint8_t gshift[kBuckets];
  __shared__ uint32_t sgstart[kMaxSub], sgcur[kMaxSub];
  __shared__ uint32_t wsum[NW];
  __shared__ int nactive;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTile - 1), ty = tid >> 4;
  const int X0 = rx0 + tile_x * kTile, Y0 = ry0 + tile_y * kTile;
  const int x = X0 + tx, y = Y0 + ty;
  const bool active = (x < rx0 + rw) && (y < ry0 + rh) && on_disk(x, y, cx, cy, r);
  if (tid < 8) need[tid] = 0;
  if (tid == 0) nactive = 0;
  __syncthreads();
  int bs = 0, rr = 0, mpop = 0;
  if (active) {
    size_t o = (size_t)y * W + x;
    bs = bstar[o];
    uint32_t pk = rres[o];
    rr = (int)(pk & 0xFFFFu);
    mpop = (int)(pk >> 16);
    atomicOr(&need[bs >> 5], 1u << (bs & 31));
    nactive = 1;
  }
  __syncthreads();
  if (nactive == 0) return;
  if (tid < 9) {
    uint32_t p = 0;
    for (int i = 0; i < tid; ++i) p += __popc(need[i]);
    needpfx[tid] = p;
  }
  __syncthreads();
  const int ngroups = (int)needpfx[8];
  // digits per group: as many as the counter budget allows (64 for <= 32 buckets, the common
  // case): finer sub-groups make the rank sort below cheaper
  const int dlog = ngroups <= kMaxSub / 64 ? 6 : ngroups <= kMaxSub / 32 ? 5
                 : ngroups <= kMaxSub / 16 ? 4 : 3;
  const int nsg = ngroups << dlog;
  const int ndm1 = (1 << dlog) - 1;
  {   // key range of bucket `tid` -> digit shift of its group; clear the sub-group counters
    const int b = tid;
    uint16_t gl = 0xFFFF;
    if ((need[b >> 5] >> (b & 31)) & 1u) {
      K klo = (b == 0) ? (K)0 : bounds[b - 1];
      K span = bounds[b] - klo;               // bucket = [klo, bounds[b]) ; bounds[255] = kMax
      if (b == kBuckets - 1 && span != Key<T>::kMax) span += 1;   // the last bucket is closed
      int g = (int)(needpfx[b >> 5] + __popc(need[b >> 5] & ((1u << (b & 31)) - 1u)));
      gl = (uint16_t)g;
      gklo[g] = klo;
      gshift[g] = (uint8_t)max(0, bit_length<K>(span - 1) - dlog);
    }
    glut[b] = gl;
    for (int i = tid; i < nsg; i += NT) sgstart[i] = 0;
  }
  __syncthreads();

  // ---- b. stage the region asynchronously, then one shared-memory pass turns values into
  //         ordered keys, looks up each pixel's sub-group and counts the members.
  //         TMA path: two 2-D tensor-tile loads (cp.async.bulk.tensor) issued by one thread and
  //         tracked by an mbarrier; out-of-image coordinates are zero-filled by the TMA unit,
  //         which IS medfilt2d's zero padding.  Fallback: per-lane LDGSTS copies.
  const int RS = kTile + 2 * half;
  constexpr int EPA = 16 / (int)sizeof(T);          // elements per 16 bytes
  // TMA needs the first byte of a box row 16-byte aligned: start the boxes at the aligned
  // column at or left of the region (vmis / bmis = how far the region sits inside the box)
  const int vpitch = (RS + 2 * EPA - 2) & ~(EPA - 1);
  const int vx0 = use_tma ? ((X0 - half) & ~(EPA - 1)) : (X0 - half);   // & on negatives rounds down
  const int vmis = (X0 - half) - vx0;
  const int bpitch = use_tma ? ((RS + 30) & ~15) : ((RS + 6) & ~3);
  const int bmis = use_tma ? (X0 & 15) : (X0 & 3);  // LDGSTS rows start at word (X0 & ~3)
  {
    T* sraw = reinterpret_cast<T*>(skey);           // raw values first, keys in place afterwards
    uint8_t* sbyte = reinterpret_cast<uint8_t*>(csg);   // dead before csg is written
    if (use_tma) {
      if (tid == 0) {
        // earlier tiles of this CTA touched the buffers through the generic proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = (uint32_t)(RS * vpitch * (int)sizeof(T) + RS * bpitch);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3}], [%4];" ::"r"((uint32_t)__cvta_generic_to_shared(sraw)),
            "l"(reinterpret_cast<uint64_t>(tm_img)), "r"(vx0), "r"(Y0 - half), "r"(bar)
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3}], [%4];" ::"r"((uint32_t)__cvta_generic_to_shared(sbyte)),
            "l"(reinterpret_cast<uint64_t>(tm_bkt)), "r"(X0 & ~15), "r"(Y0), "r"(bar)
            : "memory");
      }
      uint32_t ok = 0;
      for (int spin = 0; !ok && spin < (1 << 22); ++spin) {   // bounded: a bad descriptor must not hang
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok) : "r"(bar), "r"(tma_phase) : "memory");
      }
      tma_phase ^= 1u;
      if (!ok && tid == 0) atomicAdd(err, 1 << 20);
    } else {
      const bool interior = (X0 - half >= 0) && (Y0 - half >= 0) && (X0 - half + RS <= W) &&
                            (Y0 - half + RS <= H);
      const int bwords = (bmis + RS + 3) >> 2;
      for (int ry = warp; ry < RS; ry += NW) {
        const int iy = Y0 + ry - half;
        const T* irow = img + ((long long)iy * W + (X0 - half));
        for (int rx = lane; rx < RS; rx += 32) {
          int ix = X0 + rx - half;
          if (interior || (iy >= 0 && iy < H && ix >= 0 && ix < W))
            __pipeline_memcpy_async(&sraw[ry * vpitch + rx], &irow[rx], sizeof(T));   // vmis == 0 here
          else
            sraw[ry * vpitch + rx] = (T)0;          // zero padding of medfilt2d
        }
        const uint32_t* bw = reinterpret_cast<const uint32_t*>(Bp + (size_t)(Y0 + ry) * pitch + (X0 & ~3));
        if (lane < bwords) __pipeline_memcpy_async(sbyte + ry * bpitch + 4 * lane, bw + lane, 4);
      }
      __pipeline_commit();
      __pipeline_wait_prior(0);
      __syncthreads();
    }
    for (int ry = ty; ry < RS; ry += kTile) {
      for (int rx = tx; rx < RS; rx += kTile) {
        const int ri = ry * vpitch + vmis + rx;
        K key = Key<T>::enc(sraw[ri]);
        skey[ri] = key;
        const int g = glut[sbyte[ry * bpitch + bmis + rx]];
        uint16_t sg = 0xFFFF;
        if (g != 0xFFFF) {
          K dg = (key - gklo[g]) >> gshift[g];
          sg = (uint16_t)((g << dlog) + (int)(dg < (K)ndm1 ? dg : (K)ndm1));
          atomicAdd(&sgstart[sg], 1u);
        }
        rsg[ri] = sg;
      }
    }
  }
  __syncthreads();
  {   // exclusive scan of the nsg counters (contiguous chunk per thread)
    const int per = (nsg + NT - 1) / NT;
    const int lo = tid * per, hi = min(lo + per, nsg);
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += sgstart[i];
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int wv = 0; wv < warp; ++wv) base += wsum[wv];
    uint32_t run = base + inc - s;
    for (int i = lo; i < hi; ++i) {
      uint32_t c = sgstart[i];
      sgstart[i] = run;
      sgcur[i] = run;
      run += c;
    }
  }
  __syncthreads();
  // ---- c. scatter candidate positions into their sub-groups
  for (int ry = ty; ry < RS; ry += kTile) {
    for (int rx = tx; rx < RS; rx += kTile) {
      const unsigned sg = rsg[ry * vpitch + vmis + rx];
      if (sg != 0xFFFFu) {
        uint32_t slot = atomicAdd(&sgcur[sg], 1u);
        cpos[slot] = (uint16_t)((ry << 8) | rx);
        csg[slot] = (uint16_t)sg;
      }
    }
  }
  __syncthreads();
  auto key_at = [&](unsigned p) -> K { return skey[(p >> 8) * vpitch + vmis + (p & 255u)]; };
  const int C = (int)sgcur[nsg - 1];

  // ---- d. finish the sort: every candidate finds its rank inside its sub-group (rsg is dead
  //         from here on; spos aliases it)
  for (int e = tid; e < C; e += NT) {
    const unsigned sg = csg[e];
    const int s0 = (int)sgstart[sg], s1 = (int)sgcur[sg];
    const unsigned p = cpos[e];
    const K key = key_at(p);
    int rank = s0;
    for (int j = s0; j < s1; ++j) {
      const K o = key_at(cpos[j]);
      rank += (o < key || (o == key && j < e)) ? 1 : 0;
    }
    spos[rank] = (uint16_t)p;
  }
  __syncthreads();

  // ---- e. walk the sorted candidates of my bucket from the nearer end
  K kmin = Key<T>::kMax, kmax = 0;
  if (active) {
    const int g = glut[bs];
    const int beg = (int)sgstart[g << dlog], end = (int)sgcur[(g << dlog) + ndm1];
    const unsigned win = 2u * (unsigned)half;
    const unsigned bias = ((unsigned)ty << 8) | (unsigned)tx;
    const bool fwd = 2 * rr < mpop;
    const int want = fwd ? rr : mpop - 1 - rr;     // in-window candidates to skip
    const int step = fwd ? 1 : -1;
    int i = fwd ? beg : end - 1;
    int left = end - beg;
    int cnt = 0, found = -1;
    for (; left >= 4; left -= 4, i += 4 * step) {
      unsigned p0 = spos[i], p1 = spos[i + step], p2 = spos[i + 2 * step], p3 = spos[i + 3 * step];
      int c0 = in_window(p0, bias, win), c1 = in_window(p1, bias, win);
      int c2 = in_window(p2, bias, win), c3 = in_window(p3, bias, win);
      int c = c0 + c1 + c2 + c3;
      if (cnt + c > want) {                        // the target is one of these four
        if (c0 && cnt == want) found = i;
        cnt += c0;
        if (found < 0 && c1 && cnt == want) found = i + step;
        cnt += c1;
        if (found < 0 && c2 && cnt == want) found = i + 2 * step;
        cnt += c2;
        if (found < 0) found = i + 3 * step;
        break;
      }
      cnt += c;
    }
    if (found < 0) {
      for (; left > 0; --left, i += step) {
        if (in_window(spos[i], bias, win)) {
          if (cnt == want) {
            found = i;
            break;
          }
          ++cnt;
        }
      }
    }
    K res;
    if (found >= 0) {
      res = key_at(spos[found]);
    } else {                                   // inconsistent tier-1 data: must never happen
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

// Slow path of tier 2: tiles whose candidate set does not fit the fast path's fixed buffers
// (large flat areas: every region pixel shares the median's bucket) are refined with the whole
// region staged in shared memory.  A fixed grid walks the overflow list left by k_refine_fast.
template <typename T>
__global__ void __launch_bounds__(kTile* kTile) k_refine_list(
    const uint32_t* __restrict__ ovf, int gx, const T* __restrict__ img,
    const uint8_t* __restrict__ Bp, int pitch, const uint8_t* __restrict__ bstar,
    const uint32_t* __restrict__ rres, const typename Key<T>::type* __restrict__ bounds, int H, int W,
    int half, int rx0, int ry0, int rw, int rh, int cx, int cy, int r, int cap, T* __restrict__ out,
    typename Key<T>::type* __restrict__ minmax, int* __restrict__ err, int use_tma,
    const __grid_constant__ CUtensorMap tm_img, const __grid_constant__ CUtensorMap tm_bkt) {
  __shared__ __align__(8) uint64_t tma_bar;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
  const uint32_t n = ovf[0];
  if (blockIdx.x >= n) return;
  if (use_tma) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  uint32_t phase = 0;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t t = ovf[1 + i];
    refine_tile_staged<T>((int)(t % (uint32_t)gx), (int)(t / (uint32_t)gx), phase, bar, img, Bp, pitch,
                          bstar, rres, bounds, H, W, half, rx0, ry0, rw, rh, cx, cy, r, cap, out,
                          minmax, err, use_tma, &tm_img, &tm_bkt);
    __syncthreads();      // the next tile reuses every shared buffer
  }
}

// Fast path of tier 2.  One CTA per 16x16 output tile:
//   a. bitmap + range [bmin, bmax] of the buckets the tile's pixels need (tier-1 output);
//   b. the bucket bytes of the tile's (16+k-1)^2 region are streamed from the padded bucket
//      image with aligned 128-bit loads; bytes in a needed bucket become candidates
//      (position + group), typically ~5 % of the region;
//   c. only the candidates' values are fetched (zero padding outside the image), keyed and
//      counting-sorted by (bucket, leading key digits); a rank sort finishes each sub-group;
//   d. every thread walks the sorted candidates of ITS bucket from the nearer end, counting
//      those inside its own k x k window, until it reaches its residual rank.
// Tiles with more than kFastCap candidates go to the overflow list (k_refine_list).
constexpr int kFastCap = 2048;   // candidates per tile on the fast path
constexpr int kFastSub = 1024;   // (bucket, digit) sub-groups per tile

__device__ __forceinline__ uint32_t sum_u16x8(const uint4 c) {
  return (c.x & 0xFFFFu) + (c.x >> 16) + (c.y & 0xFFFFu) + (c.y >> 16) + (c.z & 0xFFFFu) + (c.z >> 16) +
         (c.w & 0xFFFFu) + (c.w >> 16);
}

template <typename T>
__global__ void __launch_bounds__(kTile* kTile, 4) k_refine_fast(
    const T* __restrict__ img, const uint8_t* __restrict__ Bp, int pitch,
    const uint8_t* __restrict__ bstar, const uint32_t* __restrict__ rres,
    const typename Key<T>::type* __restrict__ bounds, int H, int W, int half, int rx0, int ry0,
    int rw, int rh, int cx, int cy, int r, T* __restrict__ out,
    typename Key<T>::type* __restrict__ minmax, int* __restrict__ err, uint32_t* __restrict__ ovf,
    int force_overflow, uint32_t* __restrict__ dbg) {
  using K = typename Key<T>::type;
  constexpr int NT = kTile * kTile;
  constexpr int NW = NT / 32;
  static_assert(NW == 8, "one uint4 of u16 counters per sub-group");
  extern __shared__ __align__(128) unsigned char smem_tile[];
  K* tkey = reinterpret_cast<K*>(smem_tile);
  K* ckey = tkey + kFastCap;
  uint16_t* tpos = reinterpret_cast<uint16_t*>(ckey + kFastCap);
  uint16_t* cpos = tpos + kFastCap;
  uint16_t* tsg = cpos + kFastCap;
  uint16_t* csg = tsg + kFastCap;
  K* skey = tkey;               // sorted keys / positions reuse the unsorted first-pass arrays
  uint16_t* spos = tpos;
  uint8_t* tg = reinterpret_cast<uint8_t*>(csg);   // group of a first-pass candidate (dead before csg)
  __shared__ uint32_t need[8], needpfx[9];
  __shared__ uint16_t glut[kBuckets];          // bucket -> group (0xFFFF: not needed by this tile)
  __shared__ K gklo[kBuckets];
  __shared__ uint8_t gshift[kBuckets];
  // warp-private sub-group counters [sub-group][warp] (shared-memory atomics cost 2 cycles per
  // lane on the one LSU of the SM; a warp resolves its own collisions with match.any instead)
  uint16_t* wcnt = csg + kFastCap;                 // [kFastSub][NW], 16-byte aligned
  __shared__ uint16_t sgoff[kFastSub + 1];
  __shared__ uint32_t wsum[NW];
  __shared__ int nactive;
  __shared__ uint32_t ncand;
#ifdef CHIPS_DEBUG_TIMING
  long long t_0 = clock64(), t_1 = 0, t_2 = 0;
#endif

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int tx = tid & (kTile - 1), ty = tid >> 4;
  const int X0 = rx0 + blockIdx.x * kTile, Y0 = ry0 + blockIdx.y * kTile;
  const int x = X0 + tx, y = Y0 + ty;
  const bool active = (x < rx0 + rw) && (y < ry0 + rh) && on_disk(x, y, cx, cy, r);
  if (tid < 8) need[tid] = 0;
  if (tid == 0) {
    nactive = 0;
    ncand = 0;
  }
  __syncthreads();
  int bs = 0, rr = 0, mpop = 0;
  if (active) {
    size_t o = (size_t)y * W + x;
    bs = bstar[o];
    uint32_t pk = rres[o];
    rr = (int)(pk & 0xFFFFu);
    mpop = (int)(pk >> 16);
    atomicOr(&need[bs >> 5], 1u << (bs & 31));
    nactive = 1;
  }
  __syncthreads();
  if (nactive == 0) return;
  if (tid < 9) {
    uint32_t p = 0;
    for (int i = 0; i < tid; ++i) p += __popc(need[i]);
    needpfx[tid] = p;
  }
  __syncthreads();
  const int ngroups = (int)needpfx[8];
  {   // key range of bucket `tid` -> group tables (gshift holds the span's bit length for now)
    const int b = tid;
    uint16_t gl = 0xFFFF;
    if ((need[b >> 5] >> (b & 31)) & 1u) {
      K klo = (b == 0) ? (K)0 : bounds[b - 1];
      K span = bounds[b] - klo;               // bucket = [klo, bounds[b]) ; bounds[255] = kMax
      if (b == kBuckets - 1 && span != Key<T>::kMax) span += 1;   // the last bucket is closed
      int g = (int)(needpfx[b >> 5] + __popc(need[b >> 5] & ((1u << (b & 31)) - 1u)));
      gl = (uint16_t)g;
      gklo[g] = klo;
      gshift[g] = (uint8_t)bit_length<K>(span - 1);
    }
    glut[b] = gl;
  }
  __syncthreads();

  // ---- b. stream the region's bucket bytes; collect (position, group) of the candidates
  const int RS = kTile + 2 * half;
  {
    const int xoff = X0 & 15;
    const int nseg = (xoff + RS + 15) >> 4;
    const int ntask = RS * nseg;
    const uint8_t* base = Bp + (size_t)Y0 * pitch + (X0 & ~15);
    for (int t0 = 0; t0 < ntask; t0 += NT) {
      const int task = t0 + tid;
      uint32_t m = 0;
      uint64_t wlo = 0, whi = 0;
      int ry = 0, rxb = 0;
      if (task < ntask) {
        ry = task / nseg;
        const int seg = task - ry * nseg;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)ry * pitch + 16 * seg));
        wlo = ((uint64_t)v.y << 32) | v.x;
        whi = ((uint64_t)v.w << 32) | v.z;
        rxb = 16 * seg - xoff;                               // region column of byte 0
        const int lo = max(0, -rxb), hi = min(16, RS - rxb);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const unsigned b = (unsigned)((i < 8 ? wlo : whi) >> (8 * (i & 7))) & 255u;
          m |= ((need[b >> 5] >> (b & 31)) & 1u) << i;       // 8 words, 8 banks: conflict-free
        }
        m &= (0xFFFFu >> (16 - hi)) & (0xFFFFu << lo);
      }
      // one shared-memory atomic per warp: exclusive scan of the hit counts
      const uint32_t n = (uint32_t)__popc(m);
      uint32_t inc = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += u;
      }
      uint32_t wbase = 0;
      if (lane == 31 && inc) wbase = atomicAdd(&ncand, inc);
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
      uint32_t slot = wbase + inc - n;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        if (slot < (uint32_t)kFastCap) {
          const unsigned b = (unsigned)((i < 8 ? wlo : whi) >> (8 * (i & 7))) & 255u;
          tpos[slot] = (uint16_t)((ry << 8) | (rxb + i));
          tg[slot] = (uint8_t)glut[b];
        }
        ++slot;
      }
    }
  }
  __syncthreads();
  const int C = (int)ncand;
  if (C > kFastCap || force_overflow) {
    if (tid == 0) {
      uint32_t i = atomicAdd(&ovf[0], 1u);
      ovf[1 + i] = blockIdx.y * gridDim.x + blockIdx.x;
    }
    return;
  }
  // digits per group: enough sub-groups that the rank sort below sees ~1 candidate each
  int dlog = 0;
  while (dlog < 6 && (ngroups << (dlog + 1)) <= kFastSub && (ngroups << dlog) < C) ++dlog;
  const int nsg = ngroups << dlog;
  const int ndm1 = (1 << dlog) - 1;
  if (tid < ngroups) gshift[tid] = (uint8_t)max(0, (int)gshift[tid] - dlog);
  uint4* wcnt4 = reinterpret_cast<uint4*>(wcnt);
  for (int i = tid; i < nsg; i += NT) wcnt4[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();

  // ---- c. fetch + key the candidates, count the sub-groups (warp-private counters)
  for (int e0 = warp * 32; e0 < C; e0 += NT) {
    const int e = e0 + lane;
    unsigned sg = 0xFFFFu;
    if (e < C) {
      const unsigned p = tpos[e];
      const int iy = Y0 + (int)(p >> 8) - half, ix = X0 + (int)(p & 255u) - half;
      const T v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(&img[(size_t)iy * W + ix]) : (T)0;
      const K key = Key<T>::enc(v);
      const int g = tg[e];
      const K dg = (key - gklo[g]) >> gshift[g];
      sg = (unsigned)((g << dlog) + (int)(dg < (K)ndm1 ? dg : (K)ndm1));
      tkey[e] = key;
      tsg[e] = (uint16_t)sg;
    }
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, sg);
    if (e < C && (peers & lt_mask) == 0) {
      uint16_t* c = &wcnt[sg * NW + warp];
      *c = (uint16_t)(*c + __popc(peers));
    }
    __syncwarp();
  }
  __syncthreads();
  {   // exclusive scan over (sub-group, warp): contiguous sub-groups per thread
    const int per = (nsg + NT - 1) / NT;
    const int lo = tid * per, hi = min(lo + per, nsg);
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += sum_u16x8(wcnt4[i]);
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int wv = 0; wv < warp; ++wv) base += wsum[wv];
    uint32_t run = base + inc - s;
    for (int i = lo; i < hi; ++i) {
      const uint4 c = wcnt4[i];
      sgoff[i] = (uint16_t)run;
      uint4 o;
      uint32_t a0 = run, a1 = a0 + (c.x & 0xFFFFu), a2 = a1 + (c.x >> 16), a3 = a2 + (c.y & 0xFFFFu);
      uint32_t a4 = a3 + (c.y >> 16), a5 = a4 + (c.z & 0xFFFFu), a6 = a5 + (c.z >> 16);
      uint32_t a7 = a6 + (c.w & 0xFFFFu);
      run = a7 + (c.w >> 16);
      o.x = a0 | (a1 << 16);
      o.y = a2 | (a3 << 16);
      o.z = a4 | (a5 << 16);
      o.w = a6 | (a7 << 16);
      wcnt4[i] = o;
    }
    if (tid == 0) sgoff[nsg] = (uint16_t)C;
  }
  __syncthreads();
  for (int e0 = warp * 32; e0 < C; e0 += NT) {     // scatter into the sub-groups
    const int e = e0 + lane;
    const unsigned sg = e < C ? (unsigned)tsg[e] : 0xFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, sg);
    uint32_t slot = 0;
    if (e < C) slot = wcnt[sg * NW + warp];
    __syncwarp();
    if (e < C) {
      if ((peers & lt_mask) == 0) wcnt[sg * NW + warp] = (uint16_t)(slot + __popc(peers));
      slot += __popc(peers & lt_mask);
      cpos[slot] = tpos[e];
      ckey[slot] = tkey[e];
      csg[slot] = (uint16_t)sg;
    }
    __syncwarp();
  }
  __syncthreads();
#ifdef CHIPS_DEBUG_TIMING
  t_1 = clock64();
#endif
  // rank sort inside each sub-group -> fully sorted (skey, spos)
  for (int e = tid; e < C; e += NT) {
    const unsigned sg = csg[e];
    const int s0 = (int)sgoff[sg], s1 = (int)sgoff[sg + 1];
    const K key = ckey[e];
    int rank = s0;
    for (int j = s0; j < s1; ++j) {
      const K o = ckey[j];
      rank += (o < key || (o == key && j < e)) ? 1 : 0;
    }
    spos[rank] = cpos[e];
    skey[rank] = key;
  }
  __syncthreads();
#ifdef CHIPS_DEBUG_TIMING
  t_2 = clock64();
#endif

  // ---- d. walk the sorted candidates of my bucket from the nearer end
  K kmin = Key<T>::kMax, kmax = 0;
  if (active) {
    const int g = glut[bs];
    const int beg = (int)sgoff[g << dlog], end = (int)sgoff[(g + 1) << dlog];
    const unsigned win = 2u * (unsigned)half;
    const unsigned bias = ((unsigned)ty << 8) | (unsigned)tx;
    const bool fwd = 2 * rr < mpop;
    const int want = fwd ? rr : mpop - 1 - rr;     // in-window candidates to skip
    const int step = fwd ? 1 : -1;
    int i = fwd ? beg : end - 1;
    int left = end - beg;
    int cnt = 0, found = -1;
    for (; left >= 4; left -= 4, i += 4 * step) {
      unsigned p0 = spos[i], p1 = spos[i + step], p2 = spos[i + 2 * step], p3 = spos[i + 3 * step];
      int c0 = in_window(p0, bias, win), c1 = in_window(p1, bias, win);
      int c2 = in_window(p2, bias, win), c3 = in_window(p3, bias, win);
      int c = c0 + c1 + c2 + c3;
      if (cnt + c > want) {                        // the target is one of these four
        if (c0 && cnt == want) found = i;
        cnt += c0;
        if (found < 0 && c1 && cnt == want) found = i + step;
        cnt += c1;
        if (found < 0 && c2 && cnt == want) found = i + 2 * step;
        cnt += c2;
        if (found < 0) found = i + 3 * step;
        break;
      }
      cnt += c;
    }
    if (found < 0) {
      for (; left > 0; --left, i += step) {
        if (in_window(spos[i], bias, win)) {
          if (cnt == want) {
            found = i;
            break;
          }
          ++cnt;
        }
      }
    }
    K res;
    if (found >= 0) {
      res = skey[found];
    } else {                                   // inconsistent tier-1 data: must never happen
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
#ifdef CHIPS_DEBUG_TIMING
  __syncthreads();
  if (tid == 0 && dbg) {
    uint32_t* d = dbg + 4 * (size_t)(blockIdx.y * gridDim.x + blockIdx.x);
    long long t_3 = clock64();
    d[0] = (uint32_t)(t_1 - t_0);
    d[1] = (uint32_t)(t_2 - t_1);
    d[2] = (uint32_t)(t_3 - t_2);
    d[3] = (uint32_t)C;
  }
#endif
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
static int refine_vpitch(int k, int elem) {
  int rs = kTile + (k - 1), epa = 16 / elem;
  return (rs + 2 * epa - 2) & ~(epa - 1);
}
static int refine_cap(int k, int elem) {
  int rs = kTile + (k - 1);
  return (rs * refine_vpitch(k, elem) + 31) & ~31;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows][cols] of `elem`-byte words, box [box_rows][box_cols], OOB -> 0
static bool make_tile_map(CUtensorMap* tm, const void* base, int elem, long long cols, long long rows,
                          long long row_bytes, int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || getenv("CHIPS_NO_TMA")) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_bytes & 15) || ((box_cols * elem) & 15) ||
      box_cols > 256 || box_rows > 256 || box_cols > cols || box_rows > rows)
    return false;     // (a box larger than the tensor faults on sm_100a: tiny images use LDGSTS)
  CUtensorMapDataType dt = elem == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                         : elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
  uint32_t* rres = reinterpret_cast<uint32_t*>(ws + p->off_rres);
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
    // run length L: every run pays a (k-1)-row warm-up; 3 CTAs (64 KB of histograms each) are
    // resident per SM.  Minimise waves(L) * (L + k).
    int nbx = (rw + kHuangThreads - 1) / kHuangThreads;
    const long long resident = 3LL * kNumSM;
    int L = rh;
    long long best = -1;
    for (int cand = (k < 8 ? 8 : k); cand <= rh; ++cand) {
      long long blocks = (long long)nbx * ((rh + cand - 1) / cand);
      long long waves = (blocks + resident - 1) / resident;
      long long cost = waves * (cand + k);
      if (best < 0 || cost < best) {
        best = cost;
        L = cand;
      }
    }
    dim3 grid(nbx, (rh + L - 1) / L);
    size_t smem = (size_t)kBuckets * kHuangThreads * sizeof(uint16_t);
    CHIPS_CUDA(cudaFuncSetAttribute(k_huang, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_huang<<<grid, kHuangThreads, smem, st>>>(Bp, p->bp_pitch, half, k, W, rx0, ry0, rw, rh, L, cx,
                                               cy, r, bstar, rres);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (stages & 8) {  // tier 2: fast path over every tile, then the overflow list
    uint32_t* ovf = reinterpret_cast<uint32_t*>(ws + p->off_ovf);
    const int gx = (rw + kTile - 1) / kTile, gy = (rh + kTile - 1) / kTile;
    static const int force_ovf = getenv("CHIPS_REFINE_STAGED") ? 1 : 0;   // tests: slow path only
    CHIPS_CUDA(cudaMemsetAsync(ovf, 0, sizeof(uint32_t), st));
    const size_t smem_fast = (size_t)kFastCap * (2 * sizeof(K) + 4 * sizeof(uint16_t)) +
                             (size_t)kFastSub * (kTile * kTile / 32) * sizeof(uint16_t);
    CHIPS_CUDA(cudaFuncSetAttribute(k_refine_fast<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_fast));
    k_refine_fast<T><<<dim3(gx, gy), kTile * kTile, smem_fast, st>>>(
        in, Bp, p->bp_pitch, bstar, rres, bounds, H, W, half, rx0, ry0, rw, rh, cx, cy, r, out, minmax,
        err, ovf, force_ovf, reinterpret_cast<uint32_t*>(ws + p->off_parent_c));
    CHIPS_CHECK_LAUNCH();
    count_launch();
    int cap = refine_cap(k, (int)sizeof(T));
    size_t smem = (size_t)cap * (sizeof(K) + 3 * sizeof(uint16_t));
    const int rs = kTile + 2 * half;
    const int vp = refine_vpitch(k, (int)sizeof(T));
    CUtensorMap tm_img, tm_bkt;
    memset(&tm_img, 0, sizeof(tm_img));
    memset(&tm_bkt, 0, sizeof(tm_bkt));
    // rs >= 24: the staged bucket tile (rs x ((rs+30)&~15) bytes) must fit the area it aliases
    int use_tma = rs >= 24 &&
                  make_tile_map(&tm_img, in, (int)sizeof(T), W, H, (long long)W * sizeof(T), vp, rs) &&
                  make_tile_map(&tm_bkt, Bp, 1, p->bp_pitch, p->bp_rows, p->bp_pitch, (rs + 30) & ~15, rs);
    CHIPS_CUDA(cudaFuncSetAttribute(k_refine_list<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (long long)gx * gy;
    const int nlist = (int)(ntiles < 3LL * kNumSM ? ntiles : 3LL * kNumSM);
    k_refine_list<T><<<nlist, kTile * kTile, smem, st>>>(ovf, gx, in, Bp, p->bp_pitch, bstar, rres, bounds,
                                                         H, W, half, rx0, ry0, rw, rh, cx, cy, r, cap, out,
                                                         minmax, err, use_tma, tm_img, tm_bkt);
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
