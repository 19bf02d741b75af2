// median.cu — exact zero-padded k x k median filter.
//
// Replaces scipy.signal.medfilt2d at /root/reference/chips/chips.py:214-216 and
// /root/reference/chips/syn_chips.py:106.  A median is a pure order statistic, so any
// exact selection reproduces scipy to the bit.  The B200 design is a block sort followed by a
// two-tier selection on RANKS:
//
//   0. the ROI is cut into blocks of 128 columns x L rows (L a multiple of the 16-pixel
//      select tile, chosen by chips_plan_init so that one wave of tier-1 warps covers the ROI).
//   1. k_block_sort  one CTA per block: the block's halo-extended region arrives as one TMA tensor
//                    tile (cp.async.bulk.tensor; the unit's zero fill outside the image IS the
//                    zero padding of medfilt2d) and is radix-sorted in shared memory (LSD, 4-bit
//                    digits over the significant bits of key - min, thread-private counters, no
//                    atomics); the two outputs leave as bulk shared -> global copies.
//                    Output: the block's pixel POSITIONS in key order (u16 list; a pixel's index in
//                    it is its rank, unique - ties are broken by the sort) and every pixel's bucket
//                    = rank / S (u8 image, 128 buckets of exactly equal population).  From here on
//                    no kernel compares values.
//   2. k_huang4      tier 1: one warp per block, every thread slides FOUR adjacent columns
//                    down the block (Huang's O(k) update) with a shared core histogram of the
//                    common window columns plus packed per-column fringe counters, private to
//                    the thread, in shared memory (bank-conflict-free interleaved layout).
//                    Output per pixel: the bucket that holds the median, the residual rank
//                    inside that bucket and the bucket's population in the window.
//   3. k_rank_select tier 2: one CTA per 16x16 output tile.  A bucket is a contiguous slice of S
//                    entries of the block's sorted list, so "the sorted candidates of a bucket"
//                    need no sort: one warp per bucket some pixel of the tile needs streams the
//                    slice and keeps the positions inside the tile's (16+k-1)^2 region (ballot
//                    compaction, order preserved).  Every thread then walks the candidates of ITS
//                    bucket from the nearer end, counting those inside its own k x k window until
//                    it reaches its residual rank, and fetches that one value.  At most
//                    kSelSlots / S buckets per round, so there is no overflow path: flat or
//                    quantised images cost the same as continuous ones.  The fused epilogue applies
//                    the disk mask and reduces min/max for the histogram.
//
// k_median_brute is an independent per-thread radix-select used only as the at-scale
// checker in the tests.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace chips {

constexpr int kBuckets = 128;   // 8 resident tier-1 warps per SM (24 KB of counters each)
constexpr int kHuangThreads = 128;
constexpr int kTile = 16;

// ------------------------------------------------------------------ 0. block geometry
// Block b = (bx, by) covers image columns [rx0 + 128 bx, +128) and rows [ry0 + L by, +L); its
// rank / bucket images hold the block extended by `half` on every side: image (x, y) lives at
// local (x - X0b + half, y - Y0b + half), pitch LP elements (multiple of 16), LR rows.
struct BlockGeom {
  int rx0, ry0, L, nbx, LP, LR, half;
};

// does the rectangle [x0, x0+w) x [y0, y0+h) hold an on-disk pixel?
__device__ __forceinline__ bool rect_on_disk(int x0, int y0, int w, int h, int cx, int cy, int r) {
  if (r < 0) return true;
  int nx = min(max(cx, x0), x0 + w - 1), ny = min(max(cy, y0), y0 + h - 1);
  return on_disk(nx, ny, cx, cy, r);
}

template <typename K>
__device__ __forceinline__ int bit_length(K v) {
  return (sizeof(K) == 8) ? 64 - __clzll((long long)v) : 32 - __clz((int)v);
}

// ------------------------------------------------------------------ 1. block sort
// LSD radix sort of one block's halo-extended region in shared memory, 4-bit digits over the
// significant bits of key - min.  Shared memory: keys[n] (ordered keys, row-major, never moved),
// two u16 index arrays (the current order and the one being scattered) and 16 x NT u16 counters.
// Every thread owns a CONTIGUOUS chunk of the current order and private digit counters, so a
// pass needs no cross-lane matching and no atomics:
//   1. walk my chunk: digit d, my running count of d (= rank among my own elements with digit d),
//      kept packed in a register per element;
//   2. exclusive scan of all counters in (digit, thread) order: packed warp scans + one scan of the
//      16 x warps warp totals; my counters become my start offsets;
//   3. scatter: slot = start[d] + my running count.  Stable, hence LSD-correct.
// (A first version used 8-bit digits with warp-private counters and match.any / eight ballots per
// round of 32 elements: 700 instructions per element against ~250 here.)
template <typename T> struct SortCfg;
// NT threads, IT elements per thread at most, VEC elements per vector load, DB bits per digit
// (float32 with 512 threads x 48 elements and 5-bit digits - one pass fewer - measured 0.39 ms against 0.37 ms)
template <> struct SortCfg<float> { static constexpr int NT = 1024, IT = 24, VEC = 8, DB = 4; };    // 24576 pixels
template <> struct SortCfg<double> { static constexpr int NT = 512, IT = 36, VEC = 4, DB = 4; };    // 17600 pixels
static_assert(1024 * 24 >= CHIPS_MAX_BLOCK_ELEMS_F32 && 512 * 36 >= CHIPS_MAX_BLOCK_ELEMS_F64, "block sort chunks");
template <typename T> static int sort_chunk(int n) {
  return (((n + SortCfg<T>::NT - 1) / SortCfg<T>::NT) + SortCfg<T>::VEC - 1) / SortCfg<T>::VEC * SortCfg<T>::VEC;
}

template <typename T>
static size_t block_sort_smem(int n) {   // keys[n] (padded to 128 bytes: the orders double as the TMA landing area) + two orders of NT * chunk slots
  const size_t npad = ((size_t)n + 15) & ~(size_t)7;
  return ((npad * sizeof(T) + 127) & ~(size_t)127) + (size_t)4 * SortCfg<T>::NT * sort_chunk<T>(n);
}

template <typename T>
__global__ void __launch_bounds__(SortCfg<T>::NT, 1) k_block_sort(
    const T* __restrict__ img, int H, int W, BlockGeom g, int cx, int cy, int r, int S,
    uint16_t* __restrict__ Ord, uint8_t* __restrict__ Bp, typename Key<T>::type* __restrict__ minmax,
    int* __restrict__ n_nonfinite, const int tma_boxw, const __grid_constant__ CUtensorMap tm_img) {
  using K = typename Key<T>::type;
  constexpr int NT = SortCfg<T>::NT, IT = SortCfg<T>::IT, VEC = SortCfg<T>::VEC, NW = NT / 32;
  constexpr int DB = SortCfg<T>::DB, ND = 1 << DB;
  extern __shared__ __align__(128) unsigned char sort_smem[];
  __shared__ __align__(8) uint64_t tma_bar;
  __shared__ K s_min[NW], s_max[NW];
  __shared__ uint16_t s_wtot[ND * NW];          // [digit][warp] totals, then exclusive starts
  __shared__ uint16_t s_cnt[ND * NT];           // [digit][thread] counts, then start offsets
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blk = blockIdx.x, bx = blk % g.nbx, by = blk / g.nbx;
  if (blk == 0 && tid == 0) {
    minmax[0] = Key<T>::kMax;   // running min of the masked output (k_rank_select epilogue)
    minmax[1] = 0;              // running max
  }
  const int X = g.rx0 + kHuangThreads * bx, Y = g.ry0 + g.L * by;
  if (!rect_on_disk(X, Y, kHuangThreads, g.L, cx, cy, r)) return;
  const int ew = kHuangThreads + 2 * g.half, eh = g.LR, n = ew * eh;
  const int npad = (n + 15) & ~7;
  K* keys = reinterpret_cast<K*>(sort_smem);
  const int ncap = NT * ((((n + NT - 1) / NT) + VEC - 1) / VEC * VEC);      // slots of one order
  uint16_t* ia = reinterpret_cast<uint16_t*>(sort_smem + (((size_t)npad * sizeof(K) + 127) & ~(size_t)127));
  uint16_t* ib = ia + ncap;
  // my counter of digit d: s_cnt[d * NT + slot], slot = 2 * lane + (warp & 1) + 64 * (warp >> 1): lane l
  // of every warp addresses bank l whatever the digit (with slot = tid, lanes 2j and 2j+1 share a
  // bank and differ in the digit row: a 2-way conflict on every counter access)
  uint16_t* mycnt = s_cnt + (2 * lane + (warp & 1) + 64 * (warp >> 1));

  // ---- load the region as ordered keys, identity order, key range.
  //      TMA path (tma_boxw > 0): ONE 2-D tensor-tile load (cp.async.bulk.tensor, tracked by an
  //      mbarrier) lands the whole halo-extended region in the area of the two order arrays; the
  //      TMA unit zero-fills coordinates outside the image, which IS medfilt2d's zero padding.  The
  //      box starts at the 16-byte aligned column at or left of the region (`vmis` columns early).
  //      Fallback (unaligned rows, tiny images, no room for the box): coalesced row loads.
  K lmin = Key<T>::kMax, lmax = 0;
  bool bad_value = false;
  if (tma_boxw > 0) {
    constexpr int EPA = 16 / (int)sizeof(T);
    const int vx0 = (X - g.half) & ~(EPA - 1);          // (& on a negative rounds down)
    const int vmis = (X - g.half) - vx0;
    const T* sraw = reinterpret_cast<const T*>(ia);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(tma_boxw * eh * (int)sizeof(T));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
          "[%0], [%1, {%2, %3}], [%4];" ::"r"((uint32_t)__cvta_generic_to_shared(ia)),
          "l"(reinterpret_cast<uint64_t>(&tm_img)), "r"(vx0), "r"(Y - g.half), "r"(bar)
          : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; !ok && spin < (1 << 22); ++spin) {      // bounded: a bad descriptor must not hang
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
          : "=r"(ok) : "r"(bar) : "memory");
    }
    if (!ok) bad_value = true;                                 // reported as non-finite input: the host raises
    for (int row = warp; row < eh; row += NW) {
      for (int lx = lane; lx < ew; lx += 32) {
        const T v = sraw[row * tma_boxw + vmis + lx];
        bad_value |= nonfinite(v);
        const K key = Key<T>::enc(v);
        keys[row * ew + lx] = key;
        lmin = key < lmin ? key : lmin;
        lmax = key > lmax ? key : lmax;
      }
    }
    __syncthreads();                                           // the landing area becomes the order arrays
  } else {
    for (int row = warp; row < eh; row += NW) {
      const int y = Y - g.half + row;
      const bool yin = y >= 0 && y < H;
      const T* irow = img + (size_t)(yin ? y : 0) * W;
      for (int lx = lane; lx < ew; lx += 32) {
        const int x = X - g.half + lx;
        const T v = (yin && x >= 0 && x < W) ? __ldg(irow + x) : (T)0;   // medfilt2d zero padding
        bad_value |= nonfinite(v);
        const K key = Key<T>::enc(v);
        const int e = row * ew + lx;
        keys[e] = key;
        lmin = key < lmin ? key : lmin;
        lmax = key > lmax ? key : lmax;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const K a = __shfl_xor_sync(0xFFFFFFFFu, lmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, lmax, o);
    lmin = a < lmin ? a : lmin;
    lmax = b > lmax ? b : lmax;
  }
  if (lane == 0) {
    s_min[warp] = lmin;
    s_max[warp] = lmax;
  }
  if (bad_value) atomicAdd(n_nonfinite, 1);     // (pixels in block halos are met more than once)
  __syncthreads();
  K kmin = s_min[lane % NW], kmax = s_max[lane % NW];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const K a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  const int nbits = bit_length<K>(kmax - kmin);                  // only the bits that differ in this block
  // float64 keys: the first attempt sorts the TOP 32 significant bits only (8 passes instead of
  // ~13).  The stable sort leaves pixels that agree in those bits in position order; if the full
  // keys are then in order - the rule for continuous data: 17 k pixels in 2^32 cells - the block
  // is done, else it is sorted again on all its bits.
  int drop = (sizeof(K) == 8 && nbits > 32) ? nbits - 32 : 0;

  // Every thread owns exactly `chunk` consecutive slots of the current order (a multiple of VEC:
  // vector loads, no per-element guards).  The slots past n hold virtual elements e >= n whose
  // digit is always 15: they stay behind every real element (the sort is stable).
  const int chunk = (((n + NT - 1) / NT) + VEC - 1) / VEC * VEC;          // <= IT
  const int beg = tid * chunk;
  uint16_t *in = ia, *out = ib;
  while (true) {
  // (at least one pass: it creates the order, whatever the keys)
  const int npass = max(1, (nbits - drop + DB - 1) / DB);
  for (int pass = 0; pass < npass; ++pass) {
    const int shift = drop + DB * pass;
    // 1. my chunk.  All loads first (independent: nothing is stored in between), then the digit
    //    counters (a static array: the compiler knows they alias neither the keys nor the orders).
    //    item = element | digit << 16 | my running count << 20.  (Counters packed in registers
    //    cost three times the instructions of the shared-memory read-modify-write.)
    uint32_t item[IT];
    if (pass == 0) {
      // The first pass starts from ANY order (LSD only needs the later passes to be stable), so no
      // order array is read: slot beg + j of the virtual initial order holds element j * NT + tid -
      // the keys are read in place, consecutive lanes consecutive addresses (no index load, no
      // bank conflict, no identity array to initialise).
#pragma unroll
      for (int j = 0; j < IT; ++j) {
        if (j < chunk) {
          const uint32_t e = (uint32_t)(j * NT + tid);
          const uint32_t d = (uint32_t)((keys[e] - kmin) >> shift) & (uint32_t)(ND - 1);      // (e >= n reads slack)
          item[j] = e | ((e < (uint32_t)n ? d : (uint32_t)(ND - 1)) << 16);
        }
      }
    } else {
#pragma unroll
    for (int jv = 0; jv < IT / VEC; ++jv) {
      if (VEC * jv < chunk) {
        if (VEC == 8) {
          const uint4 v = *reinterpret_cast<const uint4*>(in + beg + VEC * jv);
          item[VEC * jv + 0] = v.x & 0xFFFFu; item[VEC * jv + 1] = v.x >> 16;
          item[VEC * jv + 2] = v.y & 0xFFFFu; item[VEC * jv + 3] = v.y >> 16;
          item[VEC * jv + (VEC > 4 ? 4 : 0)] = v.z & 0xFFFFu; item[VEC * jv + (VEC > 4 ? 5 : 1)] = v.z >> 16;
          item[VEC * jv + (VEC > 4 ? 6 : 2)] = v.w & 0xFFFFu; item[VEC * jv + (VEC > 4 ? 7 : 3)] = v.w >> 16;
        } else {
          const uint2 v = *reinterpret_cast<const uint2*>(in + beg + VEC * jv);
          item[VEC * jv + 0] = v.x & 0xFFFFu; item[VEC * jv + 1] = v.x >> 16;
          item[VEC * jv + 2] = v.y & 0xFFFFu; item[VEC * jv + 3] = v.y >> 16;
        }
      }
    }
#pragma unroll
    for (int jv = 0; jv < IT / VEC; ++jv) {
      if (VEC * jv < chunk) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          const int j = VEC * jv + u;
          const uint32_t e = item[j];
          CHIPS_BOUNDS(e < (uint32_t)ncap);
          const uint32_t d = (uint32_t)((keys[e] - kmin) >> shift) & (uint32_t)(ND - 1);      // (e >= n reads slack)
          item[j] = e | ((e < (uint32_t)n ? d : (uint32_t)(ND - 1)) << 16);
        }
      }
    }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) mycnt[d * NT] = 0;
#pragma unroll
    for (int jv = 0; jv < IT / VEC; ++jv) {
      if (VEC * jv < chunk) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          const int j = VEC * jv + u;
          uint16_t* c = mycnt + (item[j] >> 16) * NT;
          const uint32_t cur = *c;
          *c = (uint16_t)(cur + 1);
          item[j] |= cur << (16 + DB);
        }
      }
    }
    // 2. exclusive scan in (digit, thread) order; two digits per 32-bit word (warp sums < 2^16)
    uint32_t inc[ND / 2];
#pragma unroll
    for (int q = 0; q < ND / 2; ++q)
      inc[q] = (uint32_t)mycnt[(2 * q) * NT] | ((uint32_t)mycnt[(2 * q + 1) * NT] << 16);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int q = 0; q < ND / 2; ++q) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc[q], o);
        if (lane >= o) inc[q] += u;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int q = 0; q < ND / 2; ++q) {
        s_wtot[(2 * q) * NW + warp] = (uint16_t)(inc[q] & 0xFFFFu);
        s_wtot[(2 * q + 1) * NW + warp] = (uint16_t)(inc[q] >> 16);
      }
    }
    __syncthreads();
    if (warp == 0) {   // ND * NW totals, consecutive entries per lane
      constexpr int E = ND * NW / 32;
      uint32_t v[E], sum = 0;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        v[i] = s_wtot[lane * E + i];
        sum += v[i];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += u;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        s_wtot[lane * E + i] = (uint16_t)run;
        run += v[i];
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ND / 2; ++q) {      // my start offsets, looked up by digit below
      const uint32_t c0 = mycnt[(2 * q) * NT], c1 = mycnt[(2 * q + 1) * NT];      // my own counts
      mycnt[(2 * q) * NT] = (uint16_t)(s_wtot[(2 * q) * NW + warp] + (inc[q] & 0xFFFFu) - c0);
      mycnt[(2 * q + 1) * NT] = (uint16_t)(s_wtot[(2 * q + 1) * NW + warp] + (inc[q] >> 16) - c1);
    }
    // 3. stable scatter: all slots first (loads), then the stores
#pragma unroll
    for (int jv = 0; jv < IT / VEC; ++jv) {
      if (VEC * jv < chunk) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          const int j = VEC * jv + u;
          item[j] = (item[j] & 0xFFFFu) | ((mycnt[((item[j] >> 16) & (uint32_t)(ND - 1)) * NT] + (item[j] >> (16 + DB))) << 16);
        }
      }
    }
#pragma unroll
    for (int jv = 0; jv < IT / VEC; ++jv) {
      if (VEC * jv < chunk) {
#pragma unroll
        for (int u = 0; u < VEC; ++u) {
          CHIPS_BOUNDS((item[VEC * jv + u] >> 16) < (uint32_t)ncap);
          out[item[VEC * jv + u] >> 16] = (uint16_t)(item[VEC * jv + u] & 0xFFFFu);
        }
      }
    }
    __syncthreads();   // `out` is complete; s_wtot and the counters are free again
    uint16_t* t = in;
    in = out;
    out = t;
  }
  if (drop == 0) break;
  bool bad = false;                                            // is the order exact on the full keys?
  for (int i = tid; i + 1 < n; i += NT) bad |= keys[in[i]] > keys[in[i + 1]];
  if (!__syncthreads_or(bad)) break;
  drop = 0;                                                    // again, on every bit
  }
  // ---- outputs: the block's pixels in key order as (ly << 8 | lx), and the bucket image: bucket =
  //      rank / S by position.  Both are assembled in shared memory (the order array in place, the
  //      bucket image at its global pitch in the other order array) and leave the SM as two bulk
  //      asynchronous copies (cp.async.bulk shared -> global) issued by one thread.
  {
    uint16_t* ord = Ord + (size_t)blk * ((n + 7) & ~7);
    uint8_t* bkt = reinterpret_cast<uint8_t*>(out);          // [eh][LP] bucket by position
    uint8_t* bp = Bp + (size_t)blk * g.LR * g.LP;
    // e / ew == umulhi(e, magic) for e < 2^16 (ew, S < 2^16)
    const uint32_t magic = 0xFFFFFFFFu / (uint32_t)ew + 1u, magic_s = 0xFFFFFFFFu / (uint32_t)S + 1u;
    const bool bulk = (size_t)eh * g.LP <= (size_t)2 * ncap;  // the pitched bucket image fits the spare order array
    const int bpitch = bulk ? g.LP : ew;
    for (int i = tid; i < n; i += NT) {
      const uint32_t e = in[i];
      CHIPS_BOUNDS(e < (uint32_t)n);                                  // the virtual elements sort last
      const uint32_t ly = __umulhi(e, magic), lx = e - ly * ew;
      CHIPS_BOUNDS(ly == e / (uint32_t)ew && ly < (uint32_t)eh && __umulhi((uint32_t)i, magic_s) == (uint32_t)i / (uint32_t)S);
      const uint16_t pos = (uint16_t)((ly << 8) | lx);
      if (bulk) in[i] = pos;
      else ord[i] = pos;
      bkt[ly * bpitch + lx] = (uint8_t)__umulhi((uint32_t)i, magic_s);
    }
    if (bulk) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy reads
      __syncthreads();
      if (tid == 0) {
        const uint32_t nb_ord = (uint32_t)(((n + 7) & ~7) * 2), nb_bkt = (uint32_t)(eh * g.LP);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ord),
                     "r"((uint32_t)__cvta_generic_to_shared(in)), "r"(nb_ord) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(bp),
                     "r"((uint32_t)__cvta_generic_to_shared(bkt)), "r"(nb_bkt) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads
      }
    } else {
      __syncthreads();
      const int groups = (ew + 3) >> 2;
      for (int row = warp; row < eh; row += NW) {
        for (int gq = lane; gq < groups; gq += 32) {
          const int lx = 4 * gq;
          const uint8_t* src = bkt + row * ew + lx;                // (reads <= 3 bytes past a row: slack columns)
          *reinterpret_cast<uint32_t*>(bp + (size_t)row * g.LP + lx) =
              (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ 3. tier 1: Huang
// hist layout (uint16 units): bin * NT + slot(tid), slot = lane*2 + (warp&1) + (warp>>1)*64,
// so that lane l of every warp always addresses bank l (NT/2 is a multiple of 32).
__device__ __forceinline__ int huang_slot(int tid) {
  int lane = tid & 31, warp = tid >> 5;
  return lane * 2 + (warp & 1) + (warp >> 1) * 64;
}

// number of the 4 bytes of w that are < m (m in 0..255), branch-free SWAR: the low 7 bits are
// compared through a borrow-free subtraction, the top bits by boolean algebra.
__device__ __forceinline__ int count_bytes_below(uint32_t w, uint32_t m4, uint32_t m4lo) {
  const uint32_t Hm = 0x80808080u;
  uint32_t t = ((w & ~Hm) | Hm) - m4lo;           // bit 7 of each byte: (w&0x7f) >= (m&0x7f)
  uint32_t lt = ((~w & m4) | (~(w ^ m4) & ~t)) & Hm;
  return __popc(lt);
}

// ------------------------------------------------------------------ 3. tier 1: Huang, 4 columns per thread
// Adjacent columns share 50 of the 51 window columns, so a thread that owns FOUR adjacent
// output columns x0..x0+3 keeps
//   core   u16[kBuckets]  histogram of the 48 window columns common to all four windows, and
//   fringe u32[kBuckets]  four u8 histograms packed in one word: byte j counts the (at most 3 x k)
//                    pixels that only column j's window adds to the core
// and pays 48 + 6 counter updates per row for four outputs instead of 4 x 51.  With r[i] the
// k+3 bucket bytes of one row of the union window (r[0] = column x0 - half), column j's window
// is r[j .. j+k-1]: r[3 .. k-1] is the core, r[i<3] enters the fringes of columns 0..i (one
// packed add), r[k+i] the fringes of columns i+1..3.
// The rank bookkeeping is relative to one pivot bucket P per thread: ltc / ltf count the core /
// fringe pixels below P (SWAR byte compares while streaming the row); after each row P is
// walked through the four medians in turn.
// One warp = one 128-column block; a CTA holds two warps (two horizontally adjacent blocks) so
// that the u16 core counters of lane l of both warps share bank l (conflict-free).
constexpr int kH4Warps = 2;
constexpr int kH4Stride = 32 * kH4Warps;   // counters of one bucket, all threads of the CTA

template <int KT>      // compile-time window size (0: run-time k) — lets the byte roles fold
__global__ void __launch_bounds__(32 * kH4Warps) k_huang4(const uint8_t* __restrict__ Bp, BlockGeom g,
                                                          int k_rt, int W, int rw, int rh, int cx, int cy,
                                                          int r, uint8_t* __restrict__ bstar,
                                                          uint32_t* __restrict__ rres, uint32_t* __restrict__ dbg) {
  extern __shared__ __align__(16) uint32_t h4_smem[];
  const int k = KT ? KT : k_rt;
#ifdef CHIPS_DEBUG_TIMING
  const long long t_begin = clock64();
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
  uint16_t* core_all = reinterpret_cast<uint16_t*>(h4_smem);          // [kBuckets][kH4Stride] u16
  uint32_t* fr_all = h4_smem + kBuckets * kH4Stride / 2;               // [kBuckets][kH4Stride] u32
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bx = blockIdx.x * kH4Warps + warp;
  if (bx >= g.nbx) return;                                             // no CTA-wide barrier below
  const int half = g.half, pitch = g.LP, rx0 = g.rx0, ry0 = g.ry0;
  uint16_t* myc = core_all + lane * kH4Warps + warp;                   // bank = lane
  uint32_t* myf = fr_all + warp * 32 + lane;
  const int x0 = rx0 + bx * kHuangThreads + 4 * lane;
  const int Yb = ry0 + blockIdx.y * g.L;
  int ya[4], yb[4], ya_min = 0x7FFFFFFF, yb_max = -0x7FFFFFFF;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + j;
    int a = Yb, b = min(Yb + g.L, ry0 + rh);
    if (x >= rx0 + rw) {
      b = a;
    } else if (r >= 0) {
      long long dx = x - cx;
      long long rem = (long long)r * r - dx * dx;
      if (rem < 0) {
        b = a;
      } else {
        long long dy = (long long)sqrt((double)rem);
        while ((dy + 1) * (dy + 1) <= rem) ++dy;
        while (dy * dy > rem) --dy;
        a = max(a, cy - (int)dy);
        b = min(b, cy + (int)dy + 1);
      }
    }
    ya[j] = a;
    yb[j] = b;
    if (a < b) {
      ya_min = min(ya_min, a);
      yb_max = max(yb_max, b);
    }
  }
  if (ya_min >= yb_max) return;
#pragma unroll 8
  for (int b = 0; b < kBuckets; ++b) {
    myc[b * kH4Stride] = 0;
    myf[b * kH4Stride] = 0;
  }
  const int R = (k * k - 1) >> 1;
  const int nw = (k + 3 + 3) >> 2;            // 32-bit words covering r[0 .. k+2]
  int P = 0, ltc = 0;                         // pivot bucket, core pixels below it
  uint32_t ltf = 0;                           // fringe pixels below it, one byte per column
  // block-local coordinates: r[0] of this thread is local column 4*lane (4-byte aligned), the
  // window row of image row yy is local row yy - Yb + half
  const uint8_t* col = Bp + (size_t)(blockIdx.y * g.nbx + bx) * g.LR * pitch + 4 * lane;
  auto prow = [&](int yy) {
    return reinterpret_cast<const uint32_t*>(col + (size_t)(yy - Yb + half) * pitch);
  };
  // one bucket byte `e` at union-window position i entering (sign = +1) or leaving (-1)
  auto edge_byte = [&](int i, unsigned e, int sign) {
    if (i >= k + 3) return;
    if (i >= 3 && i < k) {
      uint16_t* c = myc + e * kH4Stride;
      *c = (uint16_t)(*c + sign);
      ltc += ((int)e < P) ? sign : 0;
    } else {
      const uint32_t pat = i < 3 ? (0x01010101u >> (8 * (3 - i))) : (0x01010101u << (8 * (i - k + 1)));
      uint32_t* f = myf + e * kH4Stride;
      *f = sign > 0 ? *f + pat : *f - pat;
      if ((int)e < P) ltf = sign > 0 ? ltf + pat : ltf - pat;
    }
  };
  // A row is k+3 <= 78 bytes: all its words are fetched into registers by independent loads (the
  // row of the NEXT step while the current one is being counted), so no load latency is exposed
  // even with only four resident warps per SM.
  constexpr int kMaxW = (CHIPS_MAX_K + 6) / 4;
  auto load_row = [&](const uint32_t* row, uint32_t (&v)[kMaxW]) {
#pragma unroll
    for (int w = 0; w < kMaxW; ++w)
      if (w < nw) v[w] = __ldg(row + w);
  };
  auto add_row = [&](const uint32_t (&va)[kMaxW]) {
    const uint32_t m4 = (uint32_t)P * 0x01010101u, m4lo = m4 & 0x7F7F7F7Fu;
    int acc = 0;
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) {
      if (w >= nw) break;
      const uint32_t v = va[w];
      if (w >= 1 && 4 * w + 3 < k) {          // four core pixels
        acc += count_bytes_below(v, m4, m4lo);
        const unsigned a0 = v & 255u, a1 = (v >> 8) & 255u, a2 = (v >> 16) & 255u, a3 = v >> 24;
        uint16_t *p0 = myc + a0 * kH4Stride, *p1 = myc + a1 * kH4Stride;
        uint16_t c0 = *p0, c1 = *p1;
        uint16_t n0 = (uint16_t)(c0 + 1);
        *p0 = n0;
        *p1 = (uint16_t)((a0 == a1 ? n0 : c1) + 1);
        uint16_t *p2 = myc + a2 * kH4Stride, *p3 = myc + a3 * kH4Stride;
        uint16_t c2 = *p2, c3 = *p3;
        uint16_t n2 = (uint16_t)(c2 + 1);
        *p2 = n2;
        *p3 = (uint16_t)((a2 == a3 ? n2 : c3) + 1);
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) edge_byte(4 * w + b, (v >> (8 * b)) & 255u, +1);
      }
    }
    ltc += acc;
  };
  auto swap_row = [&](const uint32_t (&xa)[kMaxW], const uint32_t (&xb)[kMaxW]) {
    const uint32_t m4 = (uint32_t)P * 0x01010101u, m4lo = m4 & 0x7F7F7F7Fu;
    int acc = 0;
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) {
      if (w >= nw) break;
      const uint32_t va = xa[w], vb = xb[w];
      if (w >= 1 && 4 * w + 3 < k) {
        acc += count_bytes_below(va, m4, m4lo) - count_bytes_below(vb, m4, m4lo);
#pragma unroll
        for (int b = 0; b < 4; ++b) {         // +1 entering, -1 leaving (same bucket: net zero)
          const unsigned ea = (va >> (8 * b)) & 255u, eb = (vb >> (8 * b)) & 255u;
          uint16_t *pa = myc + ea * kH4Stride, *pb = myc + eb * kH4Stride;
          uint16_t ca = *pa, cb = *pb;
          uint16_t na = (uint16_t)(ca + 1);
          *pa = na;
          *pb = (uint16_t)((ea == eb ? na : cb) - 1);
        }
      } else {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          edge_byte(4 * w + b, (va >> (8 * b)) & 255u, +1);
          edge_byte(4 * w + b, (vb >> (8 * b)) & 255u, -1);
        }
      }
    }
    ltc += acc;
  };

  uint32_t ra[kMaxW], rb[kMaxW];
  load_row(prow(ya_min - half), ra);
  for (int yy = ya_min - half; yy < ya_min + half; ++yy) {
    uint32_t nx[kMaxW];
    load_row(prow(yy + 1), nx);               // rows up to ya_min + half exist in the block image
    add_row(ra);
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) ra[w] = nx[w];
  }
  // ra = row ya_min + half (entering at the first output row)
  for (int y = ya_min; y < yb_max; ++y) {
    uint32_t na[kMaxW], nb[kMaxW];
    if (y + 1 < yb_max) {                     // operands of the next step
      load_row(prow(y + 1 + half), na);
      load_row(prow(y - half), nb);
    }
    if (y == ya_min) add_row(ra);
    else swap_row(ra, rb);
#pragma unroll
    for (int w = 0; w < kMaxW; ++w) {
      ra[w] = na[w];
      rb[w] = nb[w];
    }
    // Walk the pivot through the four columns' medians, keeping the exact counts below it for
    // all four (packed) on the way; odd rows visit the columns right to left, so a thread whose
    // columns straddle a jump of the median (limb, hole boundary) crosses it once per row.
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = ((y - ya_min) & 1) ? 3 - jj : jj;
      if (y < ya[j] || y >= yb[j]) continue;
      const int sh = 8 * j;
      while (ltc + (int)((ltf >> sh) & 255u) > R) {
        --P;
        ltc -= myc[P * kH4Stride];
        ltf -= myf[P * kH4Stride];
      }
      int c, lt;
      while (true) {
        const int cc = myc[P * kH4Stride];
        const uint32_t ff = myf[P * kH4Stride];
        lt = ltc + (int)((ltf >> sh) & 255u);
        c = cc + (int)((ff >> sh) & 255u);
        if (lt + c > R) break;
        ltc += cc;
        ltf += ff;
        ++P;
      }
      const size_t o = (size_t)y * W + (x0 + j);
      CHIPS_BOUNDS(P >= 0 && P < kBuckets && c > 0 && R - lt >= 0 && R - lt < c);
      bstar[o] = (uint8_t)P;
      rres[o] = (uint32_t)(R - lt) | ((uint32_t)c << 16);   // residual rank | bucket population
    }
  }
#ifdef CHIPS_DEBUG_TIMING
  if (lane == 0 && dbg) {
    uint32_t* d = dbg + 4 * (size_t)(blockIdx.y * g.nbx + bx);
    d[0] = (uint32_t)(clock64() - t_begin);
    d[1] = smid;
    d[2] = (uint32_t)(yb_max - ya_min);
    d[3] = (uint32_t)(t_begin & 0xFFFFFFFFu);
  }
#endif
}

// ------------------------------------------------------------------ 3. tier 2: select
constexpr int kSelSlots = 8192;   // candidate slots per round (16 KB per CTA)

template <typename T, int KT>      // KT: compile-time window size (0 = run time)
__global__ void __launch_bounds__(kTile* kTile, 8) k_rank_select(
    const T* __restrict__ img, const uint16_t* __restrict__ Ord_all, const BlockGeom g, const int S,
    const int per_round, const uint32_t magic_L, const uint8_t* __restrict__ bstar, const uint32_t* __restrict__ rres, int H, int W, int rw, int rh,
    int cx, int cy, int r, T* __restrict__ out, typename Key<T>::type* __restrict__ minmax,
    int* __restrict__ err) {
  using K = typename Key<T>::type;
  constexpr int NT = kTile * kTile;
  constexpr int NW = NT / 32;
  __shared__ __align__(16) uint16_t cand[kSelSlots];   // [needed bucket of this round][Sp] candidates in key order
  __shared__ uint16_t ncand[kBuckets];          // candidates per needed bucket of this round
  __shared__ uint8_t blist[kBuckets];           // the needed buckets, ascending
  __shared__ uint32_t need[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const int tx = tid & (kTile - 1), ty = tid >> 4;
  const int half = KT ? KT / 2 : g.half, rx0 = g.rx0, ry0 = g.ry0;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int by = (int)__umulhi((uint32_t)(tile_y * kTile), magic_L), blk = by * g.nbx + (tile_x >> 3);   // / g.L
  const int nelem = (kHuangThreads + 2 * half) * g.LR;       // pixels of one block
  const uint16_t* __restrict__ ord = Ord_all + (size_t)blk * ((nelem + 7) & ~7);
  // block-frame position of the tile region's corner (the +half of the frame cancels the -half)
  const unsigned origin = ((unsigned)(tile_y * kTile - by * g.L) << 8) | (unsigned)((tile_x & 7) * kTile);
  const int X0 = rx0 + tile_x * kTile, Y0 = ry0 + tile_y * kTile;
  const int x = X0 + tx, y = Y0 + ty;
  const bool active = (x < rx0 + rw) && (y < ry0 + rh) && on_disk(x, y, cx, cy, r);
  if (tid < 4) need[tid] = 0;
  __syncthreads();
  int bs = 0, rr = 0, mpop = 0;
  {
    uint32_t m[4] = {0, 0, 0, 0};
    if (active) {
      const size_t o = (size_t)y * W + x;
      bs = bstar[o];
      const uint32_t pk = rres[o];
      rr = (int)(pk & 0xFFFFu);
      mpop = (int)(pk >> 16);
#pragma unroll
      for (int wi = 0; wi < 4; ++wi) m[wi] = (bs >> 5) == wi ? 1u << (bs & 31) : 0u;
    }
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
      const uint32_t v = __reduce_or_sync(0xFFFFFFFFu, m[wi]);
      if (lane == 0 && v) atomicOr(&need[wi], v);
    }
  }
  __syncthreads();
  const uint32_t nd[4] = {need[0], need[1], need[2], need[3]};
  const int nneeded = __popc(nd[0]) + __popc(nd[1]) + __popc(nd[2]) + __popc(nd[3]);
  if (nneeded == 0) return;                                   // no on-disk pixel in this tile
  // ordinal of bucket b among the needed ones
  auto ordinal = [&](int b) {
    int o = 0;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi)
      o += (wi < (b >> 5)) ? __popc(nd[wi]) : (wi == (b >> 5)) ? __popc(nd[wi] & ((1u << (b & 31)) - 1u)) : 0;
    return o;
  };
  const int myord = ordinal(bs);
  if (tid < kBuckets && ((nd[tid >> 5] >> (tid & 31)) & 1u)) blist[ordinal(tid)] = (uint8_t)tid;

  const unsigned RSm1 = (unsigned)(kTile + 2 * half - 1);
  const unsigned win = 2u * (unsigned)half;
  const unsigned bias = ((unsigned)ty << 8) | (unsigned)tx;
  // A slice of the candidate buffer holds S entries + a sentinel, at an even stride: the walk reads
  // two u16 candidates per 32-bit word.  Region coordinates are < 128 (k <= 73), so bits 7 and 15 of
  // an entry are free guard bits: four fields per word are tested against [b, b + win] by two
  // borrow-free subtractions (SWAR) instead of two compares per field.
  const int Sp = (S + 2) & ~1;
  const uint32_t kG = 0x80808080u;
  const uint32_t B2 = bias | (bias << 16);
  const uint32_t BH = (B2 + win * 0x01010101u) | kG;      // b + win <= 15 + 72: no carry between fields
  unsigned found = 0xFFFFFFFFu;
  for (int o0 = 0; o0 < nneeded; o0 += per_round) {
    const int o1 = min(o0 + per_round, nneeded);
    __syncthreads();                           // blist is written / the previous round's walks are over
    // ---- a. one warp per needed bucket: stream the bucket's S block pixels in key order (two per
    //         lane and load), keep the ones inside the tile's (16+k-1)^2 region (ballot compaction:
    //         still in key order)
    for (int o = o0 + warp; o < o1; o += NW) {
      const int b = blist[o];
      CHIPS_BOUNDS(o < kBuckets && b < kBuckets && b * S < nelem);
      const int rb = b * S, re = min(rb + S, nelem);
      uint16_t* dst = cand + (o - o0) * Sp;
      int nout = 0;
      for (int i0 = rb & ~1; i0 < re; i0 += 64) {
        const int i = i0 + 2 * lane;
        unsigned ta = 0, tb = 0;
        bool ha = false, hb = false;
        if (i < re) {                                     // (the list is padded to an even length)
          const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(ord + i));
          ta = (w & 0xFFFFu) - origin;
          tb = (w >> 16) - origin;
          ha = (i >= rb) & ((ta >> 8) <= RSm1) & ((ta & 255u) <= RSm1);
          hb = (i + 1 < re) & ((tb >> 8) <= RSm1) & ((tb & 255u) <= RSm1);
        }
        const uint32_t ma = __ballot_sync(0xFFFFFFFFu, ha), mb = __ballot_sync(0xFFFFFFFFu, hb);
        // key order: a0 b0 a1 b1 ...: before my a come the a's and b's of the lower lanes
        const int before = __popc(ma & lt_mask) + __popc(mb & lt_mask);
        CHIPS_BOUNDS((!ha || (o - o0) * Sp + nout + before < kSelSlots) &&
                     (!hb || (o - o0) * Sp + nout + before + (ha ? 1 : 0) < kSelSlots));
        if (ha) dst[nout + before] = (uint16_t)ta;
        if (hb) dst[nout + before + (ha ? 1 : 0)] = (uint16_t)tb;
        nout += __popc(ma) + __popc(mb);
      }
      if (lane == 0) {
        ncand[o - o0] = (uint16_t)nout;
        dst[nout] = 0x7F7Fu;                       // sentinel: in nobody's window (completes the last pair)
      }
    }
    __syncthreads();
    // ---- b. walk the candidates of my bucket from the nearer end, two per 32-bit word, counting
    //         those inside my window.  Walking backwards the halves of a word are swapped first, so
    //         "low half, then high half" is the walking order in both directions.
    if (active && myord >= o0 && myord < o1) {
      const uint32_t* sp32 = reinterpret_cast<const uint32_t*>(cand + (myord - o0) * Sp);
      const int end = ncand[myord - o0];
      CHIPS_BOUNDS(end <= S && (myord - o0 + 1) * Sp <= kSelSlots && rr < mpop && mpop <= end);
      const bool fwd = 2 * rr < mpop;
      const int want = fwd ? rr : mpop - 1 - rr;     // in-window candidates to skip
      const int npairs = (end + 1) >> 1;
      const int step = fwd ? 1 : -1;
      const uint32_t sel = fwd ? 0x3210u : 0x1032u;
      // bit 7 / bit 23 of the result: the low / high candidate of the word lies in my window
      auto test = [&](uint32_t w) -> uint32_t {
        const uint32_t ok = ((w | kG) - B2) & (BH - w) & kG;   // per field: f >= b and f <= b + win
        return ok & (ok >> 8) & 0x00800080u;
      };
      int p = fwd ? 0 : npairs - 1;
      int left = npairs, cnt = 0;
      uint32_t hit = 0xFFFFFFFFu;
      if (want >= 0) {
        for (; left >= 2; left -= 2, p += 2 * step) {
          const uint32_t w0 = __byte_perm(sp32[p], 0u, sel), w1 = __byte_perm(sp32[p + step], 0u, sel);
          const uint32_t m0 = test(w0), m1 = test(w1);
          const int c0 = __popc(m0), c1 = __popc(m1);
          if (cnt + c0 + c1 > want) {                  // the target is one of these four
            uint32_t w = w0, m = m0;
            if (cnt + c0 <= want) {
              cnt += c0;
              w = w1;
              m = m1;
            }
            hit = ((m & 0x80u) && cnt == want) ? (w & 0xFFFFu) : (w >> 16);
            break;
          }
          cnt += c0 + c1;
        }
        if (hit == 0xFFFFFFFFu && left == 1) {
          const uint32_t w = __byte_perm(sp32[p], 0u, sel);
          const uint32_t m = test(w);
          if (cnt + __popc(m) > want) hit = ((m & 0x80u) && cnt == want) ? (w & 0xFFFFu) : (w >> 16);
        }
      }
      if (hit != 0xFFFFFFFFu) found = hit;
    }
  }
  // ---- epilogue: fetch the one value, mask, min/max of the masked output
  K kmin = Key<T>::kMax, kmax = 0;
  if (active) {
    K res;
    if (found != 0xFFFFFFFFu) {
      const int iy = Y0 + (int)(found >> 8) - half, ix = X0 + (int)(found & 255u) - half;
      const T v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(&img[(size_t)iy * W + ix]) : (T)0;
      res = Key<T>::enc(v);
    } else {                                   // inconsistent tier-1 data: must never happen
      atomicAdd(err, 1);
      res = Key<T>::enc((T)0);
    }
    out[(size_t)y * W + x] = Key<T>::dec(res);
    kmin = kmax = res;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const K a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  // every warp of the grid would hit the same two addresses: only values that improve on the
  // (possibly stale, therefore conservative) current bounds go to the L2 atomic unit
  if (lane == 0 && kmin <= kmax) {
    if (kmin < __ldcg(&minmax[0])) atomic_min_key(&minmax[0], kmin);
    if (kmax > __ldcg(&minmax[1])) atomic_max_key(&minmax[1], kmax);
  }
}

// ------------------------------------------------------------------ small kernels (k <= 5)
// For a 3x3 or 5x5 window the two-tier machinery is all overhead (the synoptic default is k = 3):
// one thread per pixel holds the window's ordered keys in registers and runs a fixed selection
// network (compare-exchange with compile-time indices: after pass i, a[i] is the i-th smallest),
// stopping at the median's rank (30 / 234 exchanges).  At 7x7 the network (900 exchanges) is
// already slower than the two-tier path.
template <typename T, int KS>
__global__ void __launch_bounds__(256) k_median_small(const T* __restrict__ img, T* __restrict__ out, int H,
                                                      int W, int rx0, int ry0, int rw, int rh, int cx, int cy,
                                                      int r, typename Key<T>::type* __restrict__ minmax,
                                                      int* __restrict__ n_nonfinite) {
  using K = typename Key<T>::type;
  constexpr int N = KS * KS, R = (N - 1) / 2, HALF = KS / 2;
  const int x = rx0 + blockIdx.x * 32 + (threadIdx.x & 31), y = ry0 + blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool active = x < rx0 + rw && y < ry0 + rh && on_disk(x, y, cx, cy, r);
  K kmin = Key<T>::kMax, kmax = 0;
  if (active) {
    K a[N];
    bool bad_value = false;
#pragma unroll
    for (int dy = 0; dy < KS; ++dy) {
      const int iy = y + dy - HALF;
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        const int ix = x + dx - HALF;
        const T v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(&img[(size_t)iy * W + ix]) : (T)0;   // zero padding
        bad_value |= nonfinite(v);
        a[dy * KS + dx] = Key<T>::enc(v);
      }
    }
    if (bad_value) atomicAdd(n_nonfinite, 1);
#pragma unroll
    for (int i = 0; i <= R; ++i) {
#pragma unroll
      for (int j = i + 1; j < N; ++j) {
        const K lo = a[i] < a[j] ? a[i] : a[j], hi = a[i] < a[j] ? a[j] : a[i];
        a[i] = lo;
        a[j] = hi;
      }
    }
    out[(size_t)y * W + x] = Key<T>::dec(a[R]);
    kmin = kmax = a[R];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    K p = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), q = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = p < kmin ? p : kmin;
    kmax = q > kmax ? q : kmax;
  }
  if ((threadIdx.x & 31) == 0 && kmin <= kmax) {     // (see reduce_minmax in the refine kernels)
    if (kmin < __ldcg(&minmax[0])) atomic_min_key(&minmax[0], kmin);
    if (kmax > __ldcg(&minmax[1])) atomic_max_key(&minmax[1], kmax);
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
                              int cy, int r, int* __restrict__ n_nonfinite) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * W) return;
  int y = (int)(i / W), x = (int)(i % W);
  const bool on = on_disk(x, y, cx, cy, r);
  const T v = on ? in[i] : (T)0;
  if (nonfinite(v)) atomicAdd(n_nonfinite, 1);
  out[i] = v;
}

// ------------------------------------------------------------------ TMA descriptor (host)
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

// 2-D row-major image [rows][cols] of `elem`-byte words, box [box_rows][box_cols], out of bounds -> 0.
// false: the image cannot be described (rows not 16-byte aligned, a box larger than the image -
// which faults on sm_100a -, no driver entry point, CHIPS_NO_TMA=1 for A/B timing): the caller falls
// back to plain loads.
static bool make_tile_map(CUtensorMap* tm, const void* base, int elem, long long cols, long long rows,
                          int box_cols, int box_rows) {
  static const bool disabled = getenv("CHIPS_NO_TMA") != nullptr;
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || disabled) return false;
  const long long row_bytes = cols * elem;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_bytes & 15) || ((box_cols * elem) & 15) ||
      box_cols > 256 || box_rows > 256 || box_cols > cols || box_rows > rows)
    return false;
  const CUtensorMapDataType dt = elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ host-side launchers
template <typename T>
static int medfilt_impl(const T* in, T* out, const chips_plan* p, int cx, int cy, int r,
                        unsigned char* ws, cudaStream_t st, int stages) {
  using K = typename Key<T>::type;
  const int H = p->H, W = p->W, k = p->k, half = k >> 1;
  K* minmax = reinterpret_cast<K*>(ws + p->off_minmax);
  uint16_t* Ord = reinterpret_cast<uint16_t*>(ws + p->off_order);
  uint8_t* Bp = ws + p->off_bucket;
  uint8_t* bstar = ws + p->off_bstar;
  uint32_t* rres = reinterpret_cast<uint32_t*>(ws + p->off_rres);
  int* err = &reinterpret_cast<chips_result*>(ws + p->off_result)->median_errors;
  const int rx0 = p->roi_x0, ry0 = p->roi_y0, rw = p->roi_w, rh = p->roi_h;
  const BlockGeom g = {rx0, ry0, p->blk_rows, p->blk_nbx, p->bp_pitch, p->bp_rows, half};
  const int nblocks = p->blk_nbx * p->blk_nby;

  if (stages & 1) {
    CHIPS_CUDA(cudaMemsetAsync(out, 0, (size_t)H * W * sizeof(T), st));
    CHIPS_CUDA(cudaMemsetAsync(err, 0, 2 * sizeof(int), st));   // median_errors, n_nonfinite
  }
  if (k == 1) {
    // medfilt2d with a 1x1 kernel is the identity (then masked); min/max via k_minmax later
    size_t n = (size_t)H * W;
    k_copy_masked<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, H, W, cx, cy, r, err + 1);
    CHIPS_CHECK_LAUNCH();
    count_launch();
    return CHIPS_OK;
  }
  if (k <= 5) {      // direct selection in registers (all of it counts as stage 1)
    if (stages & 1) {
      CHIPS_CUDA(cudaMemsetAsync(&minmax[0], 0xFF, sizeof(K), st));     // running min = kMax
      CHIPS_CUDA(cudaMemsetAsync(&minmax[1], 0x00, sizeof(K), st));     // running max = 0
      dim3 grid((rw + 31) / 32, (rh + 7) / 8);
      if (k == 3) k_median_small<T, 3><<<grid, 256, 0, st>>>(in, out, H, W, rx0, ry0, rw, rh, cx, cy, r, minmax, err + 1);
      else k_median_small<T, 5><<<grid, 256, 0, st>>>(in, out, H, W, rx0, ry0, rw, rh, cx, cy, r, minmax, err + 1);
      CHIPS_CHECK_LAUNCH();
      count_launch();
    }
    return CHIPS_OK;
  }
  const int nelem = (kHuangThreads + 2 * half) * p->bp_rows;
  const int S = p->blk_bucket;
  if (S < 1 || (nelem + S - 1) / S > kBuckets || nelem > 65535 || S + 2 > kSelSlots) return CHIPS_E_ARG;
  if (stages & 1) {  // rank + bucket images of every block's halo-extended region
    const size_t smem = block_sort_smem<T>(nelem);
    {   // 227 KB per CTA, minus the kernel's static arrays (counters, warp totals, min/max)
      constexpr size_t kStatic = ((size_t)SortCfg<T>::NT * 2 + (SortCfg<T>::NT / 32) * 2) * (1 << SortCfg<T>::DB) +
                                 2 * (SortCfg<T>::NT / 32) * sizeof(K) + 64;
      if (sort_chunk<T>(nelem) > SortCfg<T>::IT || smem + kStatic > 227 * 1024) return CHIPS_E_ARG;
    }
    CHIPS_CUDA(cudaFuncSetAttribute(k_block_sort<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the region of a block as one TMA box, landed in the two order arrays (4 * NT * chunk bytes):
    // ew columns + the 16-byte alignment of the box start, rounded to whole 16-byte groups
    CUtensorMap tm_img;
    memset(&tm_img, 0, sizeof(tm_img));
    constexpr int EPA = 16 / (int)sizeof(T);
    const int ew = kHuangThreads + 2 * half;
    int boxw = (ew + 2 * EPA - 2) & ~(EPA - 1);
    if ((size_t)boxw * p->bp_rows * sizeof(T) > (size_t)4 * SortCfg<T>::NT * sort_chunk<T>(nelem) ||
        !make_tile_map(&tm_img, in, (int)sizeof(T), W, H, boxw, p->bp_rows))
      boxw = 0;
    k_block_sort<T><<<nblocks, SortCfg<T>::NT, smem, st>>>(in, H, W, g, cx, cy, r, S, Ord, Bp, minmax, err + 1, boxw,
                                                           tm_img);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (stages & 4) {  // tier 1: one warp per block, two blocks per CTA
    dim3 grid((p->blk_nbx + kH4Warps - 1) / kH4Warps, p->blk_nby);
    size_t smem = (size_t)kBuckets * kH4Stride * (sizeof(uint16_t) + sizeof(uint32_t));
    uint32_t* dbg = reinterpret_cast<uint32_t*>(ws + p->off_parent_c);
#define CHIPS_LAUNCH_HUANG(KT)                                                                              \
    do {                                                                                                    \
      CHIPS_CUDA(cudaFuncSetAttribute(k_huang4<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_huang4<KT><<<grid, 32 * kH4Warps, smem, st>>>(Bp, g, k, W, rw, rh, cx, cy, r, bstar, rres, dbg);      \
    } while (0)
    switch (k) {       // the kernel sizes the reference's scripts use; anything else runs the generic copy
      case 11: CHIPS_LAUNCH_HUANG(11); break;
      case 31: CHIPS_LAUNCH_HUANG(31); break;
      case 51: CHIPS_LAUNCH_HUANG(51); break;
      case 71: CHIPS_LAUNCH_HUANG(71); break;
      default: CHIPS_LAUNCH_HUANG(0); break;
    }
#undef CHIPS_LAUNCH_HUANG
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (stages & 8) {  // tier 2: one CTA per 16x16 tile
    const dim3 grid((rw + kTile - 1) / kTile, (rh + kTile - 1) / kTile);
#define CHIPS_LAUNCH_SELECT(KT)                                                                             \
    k_rank_select<T, KT><<<grid, kTile * kTile, 0, st>>>(in, Ord, g, S, kSelSlots / ((S + 2) & ~1),                     \
                                                         0xFFFFFFFFu / (uint32_t)p->blk_rows + 1u, bstar, rres, H, W, rw, rh, cx, cy, r, out, \
                                                    minmax, err)
    switch (k) {
      case 11: CHIPS_LAUNCH_SELECT(11); break;
      case 31: CHIPS_LAUNCH_SELECT(31); break;
      case 51: CHIPS_LAUNCH_SELECT(51); break;
      case 71: CHIPS_LAUNCH_SELECT(71); break;
      default: CHIPS_LAUNCH_SELECT(0); break;
    }
#undef CHIPS_LAUNCH_SELECT
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
