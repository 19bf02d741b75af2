// fl.cu — CHIMERA-style coronal-hole candidate bitmaps.
//
// Replaces CleanFl.create_coronal_hole_candidates, /root/reference/chips/cleanup.py:69-200, and
// the hook `tmp_data = tmp_data * self.cf.candidate_bitmap`, /root/reference/chips/chips.py:376-377.
//
// Per pixel, in the reference's own order and precision (float64 until the float32 cast of
// `scale_limits`, float32 ratio / sum afterwards, comparison against a float64 threshold as numpy 2
// promotes it):
//     d  = v <= clip_negative ? 0 : v                          cleanup.py:131-133
//     s  = float32(255 * (clip(log10 d, lo, hi) - lo) / (hi - lo))     :141-152
//     bmmix  = s171 / s211 >= mean171 * bmmix / mean211         :155-164
//     bmhot  = s211 + s193 <  bmhot * (mean193 + mean211)       :165-172
//     bmcool = s171 / s193 >= mean171 * bmcool / mean193        :173-181
//     candidate = bmcool * bmmix * bmhot ; stored copies are multiplied by the disk mask :184-190
// The three image means are float64 sums in numpy's pairwise order for images whose pixel count is
// 128 * 2^m (every power-of-two AIA size): 128-element leaves with 8 strided accumulators, then a
// perfect binary tree of adjacent pairs.  Other sizes use the same tree, which differs from numpy's
// (unbalanced) one in the last bits only; pixels within one float32 ulp of a threshold are counted.
//
// Two streaming kernels (HBM-bound: 3 images read twice, 1 byte per pixel written):
//   k_fl_leaf_sums   one 8-lane group per 128-element leaf, leaves reduced pairwise per CTA
//   k_fl_masks       every CTA first finishes the (tiny) tree over the CTA partials, then streams
#include <math.h>
#include "common.cuh"

namespace chips {

constexpr int kLeaf = 128;            // numpy PW_BLOCKSIZE
constexpr int kLeavesPerCta = 256;    // 256 leaves = 32768 elements per CTA (1024 threads, 4 rounds)

template <typename T>
__device__ __forceinline__ double fl_clipneg(T v, double clip_negative) {
  double d = (double)v;
  return d <= clip_negative ? 0.0 : d;
}

// partial[img][cta] = pairwise sum of the CTA's 256 leaves (adjacent-pair tree)
template <typename T>
__global__ void __launch_bounds__(1024) k_fl_leaf_sums(const T* __restrict__ a0, const T* __restrict__ a1,
                                                       const T* __restrict__ a2, size_t n,
                                                       double clip_negative, double* __restrict__ partial,
                                                       int ncta) {
  __shared__ double leaf[3][kLeavesPerCta];
  const int tid = threadIdx.x, sub = tid & 7, grp = tid >> 3;      // 128 groups of 8 lanes
  const T* src[3] = {a0, a1, a2};
#pragma unroll
  for (int im = 0; im < 3; ++im) {
    for (int round = 0; round < kLeavesPerCta / 128; ++round) {
      const int lf = round * 128 + grp;
      const size_t base = ((size_t)blockIdx.x * kLeavesPerCta + lf) * kLeaf;
      double r = 0.0;
      bool any = false;
      // numpy: r[j] = a[j]; for i in 8..120 step 8: r[j] += a[i + j]   (full leaves)
      if (base + kLeaf <= n) {
        r = fl_clipneg(src[im][base + sub], clip_negative);
#pragma unroll
        for (int i = 8; i < kLeaf; i += 8) r = __dadd_rn(r, fl_clipneg(src[im][base + i + sub], clip_negative));
        any = true;
      } else if (base < n) {   // ragged tail leaf: same lanes, then the rest sequentially below
        const size_t m = n - base, m8 = m & ~(size_t)7;
        if (m8 >= 8) {
          r = fl_clipneg(src[im][base + sub], clip_negative);
          for (size_t i = 8; i < m8; i += 8) r = __dadd_rn(r, fl_clipneg(src[im][base + i + sub], clip_negative));
          any = true;
        }
      }
      // ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7))
      double t = __dadd_rn(r, __shfl_xor_sync(0xFFFFFFFFu, r, 1));
      t = __dadd_rn(t, __shfl_xor_sync(0xFFFFFFFFu, t, 2));
      t = __dadd_rn(t, __shfl_xor_sync(0xFFFFFFFFu, t, 4));
      if (sub == 0) {
        if (base < n && base + kLeaf > n) {          // tail elements one by one (numpy's remainder loop)
          const size_t m = n - base, m8 = m & ~(size_t)7;
          double s = (m8 >= 8 && any) ? t : 0.0;
          for (size_t i = (m8 >= 8 ? m8 : 0); i < m; ++i) s = __dadd_rn(s, fl_clipneg(src[im][base + i], clip_negative));
          t = s;
        } else if (base >= n) {
          t = 0.0;
        }
        leaf[im][lf] = t;
      }
    }
  }
  __syncthreads();
  // adjacent-pair tree over the CTA's leaves, the three images side by side
  for (int stride = 1; stride < kLeavesPerCta; stride <<= 1) {
    if (tid < 3 * kLeavesPerCta) {
      const int im = tid / kLeavesPerCta, j = (tid - im * kLeavesPerCta) * 2 * stride;
      if (j + stride < kLeavesPerCta) leaf[im][j] = __dadd_rn(leaf[im][j], leaf[im][j + stride]);
    }
    __syncthreads();
  }
  if (tid < 3) partial[(size_t)tid * ncta + blockIdx.x] = leaf[tid][0];
}

struct FlThresholds {
  double mean[3], thr_mix, thr_hot, thr_cool;
};

// adjacent-pair tree over the CTA partials (ncta <= 4096 handled in shared memory by 1024 threads)
__device__ __forceinline__ void fl_finish_means(const double* __restrict__ partial, int ncta, double npix,
                                                const chips_fl_params& prm, FlThresholds& th,
                                                double* sh /* [4096] */) {
  for (int im = 0; im < 3; ++im) {
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = i < ncta ? partial[(size_t)im * ncta + i] : 0.0;
    __syncthreads();
    for (int stride = 1; stride < 4096; stride <<= 1) {
      for (int j = threadIdx.x * 2 * stride; j + stride < 4096; j += blockDim.x * 2 * stride)
        if (j + stride < ncta) sh[j] = __dadd_rn(sh[j], sh[j + stride]);
      __syncthreads();
    }
    th.mean[im] = __ddiv_rn(sh[0], npix);
    __syncthreads();
  }
  // order of the images: 0 = 171, 1 = 193, 2 = 211
  th.thr_mix = __ddiv_rn(__dmul_rn(th.mean[0], prm.bmmix_value), th.mean[2]);
  th.thr_hot = __dmul_rn(prm.bmhot_value, __dadd_rn(th.mean[1], th.mean[2]));
  th.thr_cool = __ddiv_rn(__dmul_rn(th.mean[0], prm.bmcool_value), th.mean[1]);
}

// One wavelength's clip / scale constants.  p10lo / p10hi bracket the range in which log10 has to
// be evaluated: below p10lo the clipped logarithm is `lo` exactly (s = 0), above p10hi it is `hi`
// exactly (s = s_hi); the brackets sit 1e-9 (relative) inside the safe side, far more than the
// error of any log10, so the shortcut can never change a result.
struct FlBand {
  double lo, hi, span, p10lo, p10hi;
  float s_hi;
};

__device__ __forceinline__ FlBand fl_band(const double lims[2]) {
  FlBand b;
  b.lo = lims[0];
  b.hi = lims[1];
  b.span = __dsub_rn(lims[1], lims[0]);
  b.p10lo = pow(10.0, b.lo) * (1.0 - 1e-9);
  b.p10hi = pow(10.0, b.hi) * (1.0 + 1e-9);
  b.s_hi = __double2float_rn(__ddiv_rn(__dmul_rn(255.0, __dsub_rn(b.hi, b.lo)), b.span));
  return b;
}

template <typename T>
__device__ __forceinline__ float fl_scaled(T v, double clip_negative, const FlBand& b) {
  const double d = fl_clipneg(v, clip_negative);
  if (d >= 0.0 && d <= b.p10lo) return 0.0f;   // includes log10(0) = -inf -> clipped to lo
  if (d >= b.p10hi) return b.s_hi;             // (+inf included)
  double lg = log10(d);                        // negative d (clip_negative < 0 only) -> NaN
  // np.clip == minimum(maximum(x, lo), hi): NaN propagates
  if (lg == lg) lg = fmin(fmax(lg, b.lo), b.hi);
  const double s = __ddiv_rn(__dmul_rn(255.0, __dsub_rn(lg, b.lo)), b.span);
  return __double2float_rn(s);
}

__device__ __forceinline__ bool fl_near(float q, double thr) {
  // is thr within ~one float32 ulp of q?  (the only pixels a last-bit difference in a mean can flip)
  const double dq = (double)q;
  return fabs(dq - thr) <= fabs(dq) * 1.1920928955078125e-7;     // NaN / inf -> false
}

template <typename T>
__global__ void __launch_bounds__(1024) k_fl_masks(const T* __restrict__ a171, const T* __restrict__ a193,
                                                   const T* __restrict__ a211, int H, int W, int cx, int cy,
                                                   int r, chips_fl_params prm, const double* __restrict__ partial,
                                                   int ncta, uint8_t* __restrict__ bits,
                                                   chips_fl_stats* __restrict__ stats) {
  __shared__ double sh[4096];
  __shared__ unsigned long long cnt[5];
  FlThresholds th;
  const size_t n = (size_t)H * W;
  fl_finish_means(partial, ncta, (double)n, prm, th, sh);
  if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
  __syncthreads();
  const FlBand b211 = fl_band(prm.clip_211_log), b193 = fl_band(prm.clip_193_log),
               b171 = fl_band(prm.clip_171_log);
  unsigned c_mix = 0, c_hot = 0, c_cool = 0, c_cand = 0, c_near = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float f171 = fl_scaled(a171[i], prm.clip_negative, b171);
    const float f193 = fl_scaled(a193[i], prm.clip_negative, b193);
    const float f211 = fl_scaled(a211[i], prm.clip_negative, b211);
    const float qmix = __fdiv_rn(f171, f211), qhot = __fadd_rn(f211, f193), qcool = __fdiv_rn(f171, f193);
    const bool mix = (double)qmix >= th.thr_mix;
    const bool hot = (double)qhot < th.thr_hot;
    const bool cool = (double)qcool >= th.thr_cool;
    const int y = (int)((uint32_t)i / (uint32_t)W), x = (int)((uint32_t)i - (uint32_t)y * (uint32_t)W);
    const bool disk = on_disk(x, y, cx, cy, r);
    const bool cand = mix && hot && cool;
    bits[i] = (uint8_t)((mix ? CHIPS_FL_MIX : 0) | (hot ? CHIPS_FL_HOT : 0) | (cool ? CHIPS_FL_COOL : 0) |
                        (cand ? CHIPS_FL_CAND : 0) | (disk ? CHIPS_FL_DISK : 0));
    c_mix += mix && disk;
    c_hot += hot && disk;
    c_cool += cool && disk;
    c_cand += cand && disk;
    c_near += fl_near(qmix, th.thr_mix) || fl_near(qhot, th.thr_hot) || fl_near(qcool, th.thr_cool);
  }
  unsigned v[5] = {c_mix, c_hot, c_cool, c_cand, c_near};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    unsigned s = __reduce_add_sync(0xFFFFFFFFu, v[j]);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&cnt[j], (unsigned long long)s);
  }
  __syncthreads();
  if (threadIdx.x < 5 && cnt[threadIdx.x])
    atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_mix) + threadIdx.x, cnt[threadIdx.x]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats->mean171 = th.mean[0];
    stats->mean193 = th.mean[1];
    stats->mean211 = th.mean[2];
    stats->thr_mix = th.thr_mix;
    stats->thr_hot = th.thr_hot;
    stats->thr_cool = th.thr_cool;
  }
}

// chips.py:376-377: a pixel that is not a candidate leaves every threshold mask
__global__ void k_apply_candidates(uint8_t* __restrict__ level, const uint8_t* __restrict__ bits, size_t n,
                                   uint8_t T) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n) {
    uint4 l = *reinterpret_cast<const uint4*>(level + i);
    const uint4 b = *reinterpret_cast<const uint4*>(bits + i);
    uint32_t* lw = reinterpret_cast<uint32_t*>(&l);
    const uint32_t* bw = reinterpret_cast<const uint32_t*>(&b);
    const uint32_t T4 = (uint32_t)T * 0x01010101u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // byte mask 0xFF where the candidate bit is set
      const uint32_t c = ((bw[j] >> 3) & 0x01010101u) * 0xFFu;
      lw[j] = (lw[j] & c) | (T4 & ~c);
    }
    *reinterpret_cast<uint4*>(level + i) = l;
  } else {
    for (; i < n; ++i)
      if (!(bits[i] & CHIPS_FL_CAND)) level[i] = T;
  }
}

template <typename T>
static int fl_impl(const T* a171, const T* a193, const T* a211, int H, int W, int cx, int cy, int r,
                   const chips_fl_params* prm, uint8_t* bits, chips_fl_stats* stats, double* partial,
                   int ncta, cudaStream_t st) {
  const size_t n = (size_t)H * W;
  CHIPS_CUDA(cudaMemsetAsync(stats, 0, sizeof(chips_fl_stats), st));
  k_fl_leaf_sums<T><<<ncta, 1024, 0, st>>>(a171, a193, a211, n, prm->clip_negative, partial, ncta);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  k_fl_masks<T><<<2 * kNumSM, 1024, 0, st>>>(a171, a193, a211, H, W, cx, cy, r, *prm, partial, ncta, bits,
                                             stats);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

}  // namespace chips

using namespace chips;

extern "C" size_t chips_fl_workspace_bytes(int H, int W) {
  if (H < 1 || W < 1) return 0;
  const size_t n = (size_t)H * W, per = (size_t)kLeaf * kLeavesPerCta;
  return 3 * ((n + per - 1) / per) * sizeof(double);
}

extern "C" int chips_fl_candidates(const void* d171, const void* d193, const void* d211, int H, int W,
                                   int elem, int cx, int cy, int r, const chips_fl_params* prm,
                                   uint8_t* bits, chips_fl_stats* stats, void* ws, void* stream) {
  if (!d171 || !d193 || !d211 || !prm || !bits || !stats || !ws || H < 1 || W < 1) return CHIPS_E_ARG;
  const size_t n = (size_t)H * W, per = (size_t)kLeaf * kLeavesPerCta;
  const size_t ncta = (n + per - 1) / per;
  if (ncta > 4096) return CHIPS_E_ARG;        // 128 Mi pixels: far above any AIA / synoptic size
  cudaStream_t st = (cudaStream_t)stream;
  if (elem == 4)
    return fl_impl<float>((const float*)d171, (const float*)d193, (const float*)d211, H, W, cx, cy, r, prm,
                          bits, stats, (double*)ws, (int)ncta, st);
  if (elem == 8)
    return fl_impl<double>((const double*)d171, (const double*)d193, (const double*)d211, H, W, cx, cy, r,
                           prm, bits, stats, (double*)ws, (int)ncta, st);
  return CHIPS_E_ARG;
}

extern "C" int chips_apply_candidates(const chips_plan* plan, const uint8_t* bits, void* ws, void* stream) {
  if (!plan || !bits || !ws) return CHIPS_E_ARG;
  const size_t n = (size_t)plan->H * plan->W;
  uint8_t* level = (uint8_t*)ws + plan->off_level;
  k_apply_candidates<<<(unsigned)((n / 16 + 256) / 256), 256, 0, (cudaStream_t)stream>>>(level, bits, n,
                                                                                        (uint8_t)plan->T);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}
