// stats.cu — histogram, threshold selection, per-threshold levels and probabilities.
//
//   k_minmax / k_hist      np.histogram(bins=h_bins) of the on-disk filtered image
//                          (/root/reference/chips/chips.py:235-239): numpy's uniform-bin
//                          index formula in fp64 with its +-1 edge correction, counts kept
//                          in warp-private shared-memory sub-histograms.
//   k_select               density, sticky h_thresh, scipy.signal.find_peaks(height=) with
//                          the plateau-midpoint rule, bound, limit_0 and the rounded limits
//                          (chips.py:239-244, :356-357) — one CTA, fp64, no host round trip.
//   k_threshold_prob       one pass over the filtered image produces the level image
//                          (first threshold each pixel satisfies == all T masks of
//                          chips.py:360-375) and the [T+1][20] table behind all T
//                          calculate_prob results (chips.py:432-450).
//   k_prob_epilogue        Σ_{b>x_tau} b·h / Σ b·h in numpy's pairwise-sum order, rounded.
#include <math.h>
#include "common.cuh"

namespace chips {

// ------------------------------------------------------------------ min / max
template <typename T>
__global__ void k_minmax_init(typename Key<T>::type* minmax) {
  minmax[0] = Key<T>::kMax;
  minmax[1] = 0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_minmax(const T* __restrict__ filt, int W, int rx0, int ry0,
                                                int rw, int rh, int cx, int cy, int r,
                                                typename Key<T>::type* __restrict__ minmax) {
  using K = typename Key<T>::type;
  K kmin = Key<T>::kMax, kmax = 0;
  for (int y = ry0 + blockIdx.x; y < ry0 + rh; y += gridDim.x) {
    for (int x = rx0 + threadIdx.x; x < rx0 + rw; x += blockDim.x) {
      if (!on_disk(x, y, cx, cy, r)) continue;
      K key = Key<T>::enc(filt[(size_t)y * W + x]);
      kmin = key < kmin ? key : kmin;
      kmax = key > kmax ? key : kmax;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    K a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o);
    K b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
    atomic_min_key(&minmax[0], kmin);
    atomic_max_key(&minmax[1], kmax);
  }
}

// ------------------------------------------------------------------ histogram
// numpy's bin of v: trunc(((v-first)/denom)*nbins) then at most one step down / up against the
// linspace edges.  Since that guess is within one bin of the bin whose edges bracket v, the
// result IS the bracketing bin; we reach the same bin from a cheap fp32 guess and exact fp64
// edge comparisons (no fp64 divide per pixel).
__device__ __forceinline__ int hist_bin(double v, int nbins, const Edges& E, float first_f, float scale_f) {
  int g = (int)(((float)v - first_f) * scale_f);
  g = min(max(g, 0), nbins - 1);
  while (g > 0 && v < E(g)) --g;
  while (g < nbins - 1 && v >= E(g + 1)) ++g;
  CHIPS_BOUNDS(g >= 0 && g < nbins && v >= E(g) && (g == nbins - 1 || v < E(g + 1)));
  return g;
}

// the same with the edges read from a shared-memory table E[0..nbins] (two loads and two fp64
// compares in the common case instead of two edge evaluations)
__device__ __forceinline__ int hist_bin_tab(double v, int nbins, const double* __restrict__ E, float first_f,
                                            float scale_f) {
  int g = (int)(((float)v - first_f) * scale_f);
  g = min(max(g, 0), nbins - 1);
  const double e0 = E[g], e1 = E[g + 1];
  if (v < e0) {
    do --g; while (g > 0 && v < E[g]);
  } else if (v >= e1 && g < nbins - 1) {
    do ++g; while (g < nbins - 1 && v >= E[g + 1]);
  }
  return g;
}

// Fast path of the bin search: t = (v - first) * nbins / (last - first) in float32 is within
// `eps` of the exact value (hist_eps bounds every rounding: v and first to float32, the
// subtraction, the scale and the product), so when t is farther than eps from both integers
// around it, trunc(t) IS the bin whose float64 edges bracket v and no edge is read.  ~99.8 % of
// the pixels at 500 bins; the rest (and eps >= 0.5: a range tiny against its offset) take the exact
// edge comparison.  Returns -1 when the fast path cannot decide.
__device__ __forceinline__ int hist_bin_fast(float vf, int nbins, float first_f, float scale_f, float eps) {
  const float t = (vf - first_f) * scale_f;
  const float fl = floorf(t);
  const float fr = t - fl;
  return (fr >= eps && fr <= 1.0f - eps && fl >= 0.0f && fl < (float)nbins) ? (int)fl : -1;
}

__device__ __forceinline__ float hist_eps(double first, double last, int nbins) {
  const double denom = __dsub_rn(last, first);
  const double M = fmax(fabs(first), fabs(last)) / denom;
  const double e = 8.0 * (double)nbins * 5.9604644775390625e-8 * (3.0 + 2.0 * M);     // 8 x the bound, 2^-24
  return (e < 0.25) ? (float)e * 1.0001f : 2.0f;                                    // NaN / huge: never fast
}

// row[x .. x+3], x a non-negative multiple of 4: one 128-bit load (two for float64) when the rows are
// 16-byte aligned and the group lies inside the image, guarded scalar loads otherwise
template <typename T>
__device__ __forceinline__ void load4(const T* __restrict__ row, int x, int W, bool vec, T v[4]) {
  if (vec && x + 4 <= W) {
    if (sizeof(T) == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(row + x));
      v[0] = (T)q.x; v[1] = (T)q.y; v[2] = (T)q.z; v[3] = (T)q.w;
    } else {
      const double2 a = __ldg(reinterpret_cast<const double2*>(row + x));
      const double2 b = __ldg(reinterpret_cast<const double2*>(row + x + 2));
      v[0] = (T)a.x; v[1] = (T)a.y; v[2] = (T)b.x; v[3] = (T)b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (x + i < W) ? row[x + i] : (T)0;
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) k_hist(const T* __restrict__ filt, int W, int rx0, int ry0,
                                              int rw, int rh, int cx, int cy, int r,
                                              const typename Key<T>::type* __restrict__ minmax,
                                              int nbins, int copies, int use_table, int f32_math,
                                              unsigned long long* __restrict__ counts,
                                              chips_result* __restrict__ res) {
  extern __shared__ __align__(16) uint32_t sh[];
  for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh[i] = 0;
  const typename Key<T>::type k0 = minmax[0], k1 = minmax[1];
  const Edges edge = make_edges((double)Key<T>::dec(k0), (double)Key<T>::dec(k1), k0 > k1, nbins, f32_math);
  const double first = edge.first, last = edge.last;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    res->first_edge = first;
    res->last_edge = last;
  }
  const double denom = __dsub_rn(last, first);
  const float first_f = (float)first, scale_f = (float)((double)nbins / denom);
  // edge table behind the sub-histograms (8-byte aligned: copies * nbins is rounded up to even)
  double* E = reinterpret_cast<double*>(sh + ((copies * nbins + 1) & ~1));
  if (use_table)
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) E[i] = edge(i);
  __syncthreads();
  uint32_t* mine = sh + ((threadIdx.x >> 5) % copies) * nbins;
  const float eps = hist_eps(first, last, nbins);
  const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(filt) & 15) == 0;
  constexpr int UN = 2;        // 128-bit loads of a thread in flight together (4 CTAs of 64 registers per SM)
  for (int y = ry0 + blockIdx.x; y < ry0 + rh; y += gridDim.x) {
    int xa, xb;
    warp_row_chord(y, cx, cy, r, rx0, rw, xa, xb);
    if (xa >= xb) continue;
    const T* row = filt + (size_t)y * W;
    for (int x0 = (xa & ~3) + 4 * (int)threadIdx.x; x0 < xb; x0 += 4 * UN * (int)blockDim.x) {
      T v[UN][4];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int x = x0 + 4 * u * (int)blockDim.x;
        if (x < xb) load4(row, x, W, vec, v[u]);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int x = x0 + 4 * u * (int)blockDim.x;
        if (x >= xb) break;
        int idx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int xi = x + i;
          idx[i] = -1;
          if (xi >= xa && xi < xb) {
            int b = hist_bin_fast((float)v[u][i], nbins, first_f, scale_f, eps);
            if (b < 0)
              b = use_table ? hist_bin_tab((double)v[u][i], nbins, E, first_f, scale_f)
                            : hist_bin((double)v[u][i], nbins, edge, first_f, scale_f);
            idx[i] = b;
          }
        }
        // neighbouring pixels of the median-filtered image mostly share a bin: merge first
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (idx[i] < 0) continue;
          unsigned c = 1;
#pragma unroll
          for (int j = i + 1; j < 4; ++j)
            if (idx[j] == idx[i]) {
              ++c;
              idx[j] = -1;
            }
          atomicAdd(&mine[idx[i]], c);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
    unsigned long long s = 0;
    for (int c = 0; c < copies; ++c) s += sh[c * nbins + i];
    if (s) atomicAdd(&counts[i], s);
  }
}

// ------------------------------------------------------------------ threshold selection
struct ThresholdOffsets {   // np.arange(threshold_range[0], threshold_range[1]) (chips.py:352), by value
  double v[CHIPS_MAX_T];
};

__global__ void __launch_bounds__(1024) k_select(const long long* __restrict__ counts, int nbins,
                                                 double ratio, double h_thresh_in,
                                                 const ThresholdOffsets offs, int T, int f32_math,
                                                 double* __restrict__ density,
                                                 double* __restrict__ edges,
                                                 double* __restrict__ bound, int* __restrict__ peaks,
                                                 chips_result* __restrict__ res) {
  __shared__ long long s_ll[32];
  __shared__ double s_d[32];
  __shared__ int s_cnt[1024];
  __shared__ long long total_s;
  __shared__ double hthr_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Edges edge = make_edges(res->first_edge, res->last_edge, false, nbins, f32_math);

  long long part = 0;
  for (int i = tid; i < nbins; i += blockDim.x) part += counts[i];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
  if (lane == 0) s_ll[warp] = part;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int w = 0; w < 32; ++w) t += s_ll[w];
    total_s = t;
  }
  __syncthreads();
  const double total = (double)total_s;

  double dmax = -INFINITY;
  for (int i = tid; i <= nbins; i += blockDim.x) {
    double e0 = edge(i);
    edges[i] = e0;
    if (i < nbins) {
      double e1 = edge(i + 1);
      double db = edge.width(e0, e1);
      double d = __ddiv_rn(__ddiv_rn((double)counts[i], db), total);   // n / db / n.sum()
      density[i] = d;
      bound[i] = edge.upper(e0, db);                                    // be[:-1] + diff(be)
      dmax = fmax(dmax, d);
    }
  }
  for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xFFFFFFFFu, dmax, o));
  if (lane == 0) s_d[warp] = dmax;
  __syncthreads();
  if (tid == 0) {
    double m = -INFINITY;
    for (int w = 0; w < 32; ++w) m = fmax(m, s_d[w]);
    hthr_s = isnan(h_thresh_in) ? __ddiv_rn(m, ratio) : h_thresh_in;
  }
  __syncthreads();
  const double hthr = hthr_s;

  // find_peaks(height=hthr): contiguous chunk per thread keeps the output ordered
  const int chunk = (nbins + blockDim.x - 1) / blockDim.x;
  const int lo = tid * chunk, hi = min(lo + chunk, nbins);
  auto peak_at = [&](int i) -> int {     // returns peak index started by rising edge at i, or -1
    if (i < 1 || i >= nbins - 1) return -1;
    double di = density[i];
    if (!(density[i - 1] < di)) return -1;
    int ia = i + 1;
    while (ia < nbins - 1 && density[ia] == di) ++ia;
    if (density[ia] < di) {
      int p = (i + ia - 1) / 2;
      return (density[p] >= hthr) ? p : -1;
    }
    return -1;
  };
  int c = 0;
  for (int i = lo; i < hi; ++i) c += (peak_at(i) >= 0);
  s_cnt[tid] = c;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) {
      int v = s_cnt[t];
      s_cnt[t] = run;
      run += v;
    }
    res->n_peaks = run;
    res->n_thresholds = T;
    res->n_samples = total_s;
    res->h_thresh = hthr;
  }
  __syncthreads();
  int o = s_cnt[tid];
  for (int i = lo; i < hi; ++i) {
    int p = peak_at(i);
    if (p >= 0) peaks[o++] = p;
  }
  __syncthreads();
  if (tid < T) {
    if (res->n_peaks > 0) {
      double l0 = bound[peaks[0]];
      if (tid == 0) res->limit_0 = l0;
      // np.round(limit_0 + threshold_range, 4) = rint(x * 1e4) / 1e4
      double xv = __dadd_rn(l0, offs.v[tid]);
      res->limits[tid] = __ddiv_rn(rint(__dmul_rn(xv, 10000.0)), 10000.0);
    } else {
      if (tid == 0) res->limit_0 = nan("");
      res->limits[tid] = nan("");
    }
  }
}

// ------------------------------------------------------------------ levels + prob table
// d-space cut points of calculate_prob's 20 bins: bin(d) = #{j=1..19 : d <= kProbCut[j]}
// (largest float64 d with rint(exp(d)*1e4) <= k_j, k_j the largest integer whose
// ps = 1/(1+k/1e4) still lies in bin >= j; generated and re-checked against numpy by
// tests/test_first_principles.py::test_prob_cut_constants).
__constant__ double kProbCut[20] = {
    0.0, 0x1.78e376738f1eep+1, 0x1.193ed6453eac1p+1, 0x1.bc0e9f3c39428p+0,
    0x1.62e501a66511cp+0, 0x1.193fbf4901473p+0, 0x1.b1d1f61d1fc89p-1, 0x1.3cf336143d45fp-1,
    0x1.9f3afbb8bc957p-2, 0x1.9b052730799f3p-3, 0x1.a36b7f85c4e5fp-15, -0x1.9b0da0803bfd9p-3,
    -0x1.9f38cc8a12b56p-2, -0x1.3cf5840e17938p-1, -0x1.b1d79434228fap-1, -0x1.193b60d3d1253p+0,
    -0x1.62d714d4115dep+0, -0x1.bc1676093fb01p+0, -0x1.1933302b0a04ap+1, -0x1.78d7e8e08506cp+1};

template <typename T>
__global__ void __launch_bounds__(256, 8) k_threshold_prob(const T* __restrict__ filt, int W, int rx0,
                                                        int ry0, int rw, int rh, int cx, int cy,
                                                        int r, int f32_math, const chips_result* __restrict__ res,
                                                        uint8_t* __restrict__ lev,
                                                        uint32_t* __restrict__ probtab) {
  __shared__ double slim[CHIPS_MAX_T];
  __shared__ uint32_t tab[(CHIPS_MAX_T + 1) * CHIPS_PROB_BINS];
  const int nT = res->n_thresholds;
  const bool have = res->n_peaks > 0 && nT > 0;
  for (int i = threadIdx.x; i < nT; i += blockDim.x) slim[i] = res->limits[i];
  for (int i = threadIdx.x; i < (nT + 1) * CHIPS_PROB_BINS; i += blockDim.x) tab[i] = 0;
  __syncthreads();
  if (!have) return;                 // lev is pre-filled with T: no regions (chips.py:355)
  int t_inv = 0;                     // limits ascend: the ones < -1 are a prefix
  for (int t = 0; t < nT; ++t) t_inv += (slim[t] < -1.0) ? 1 : 0;
  const double limit_0 = res->limit_0;
  const double top = slim[nT - 1];
  const bool vec = (W & 3) == 0;
  const bool vecl = vec && (reinterpret_cast<uintptr_t>(filt) & 15) == 0;
  constexpr int UN = 1;        // (8 resident CTAs per SM need 32 registers: one 128-bit load per thread)
  for (int y = ry0 + blockIdx.x; y < ry0 + rh; y += gridDim.x) {
    int xa, xb;
    warp_row_chord(y, cx, cy, r, rx0, rw, xa, xb);
    if (xa >= xb) continue;
    const T* row = filt + (size_t)y * W;
    uint8_t* lrow = lev + (size_t)y * W;
    for (int x0 = (xa & ~3) + 4 * (int)threadIdx.x; x0 < xb; x0 += 4 * UN * (int)blockDim.x) {
      T vv[UN][4];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int x = x0 + 4 * u * (int)blockDim.x;
        if (x < xb) load4(row, x, W, vecl, vv[u]);
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int x = x0 + 4 * u * (int)blockDim.x;
        if (x >= xb) break;
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int xi = x + i;
          int level = nT;
          if (xi >= xa && xi < xb) {
            double v = (double)vv[u][i];
            if (v <= top) {
              level = 0;
              for (int t = 0; t < nT; ++t) level += (slim[t] < v) ? 1 : 0;   // first t with v <= lim_t
              int bin = 0;
              if (f32_math) {
                // float32 data: numpy evaluates 1/(1 + exp(data - limit_0).round(4)) in float32
                // (chips.py:446) and bins it against float64 edges j/20.  expf here and numpy's SIMD
                // float32 exp may differ in the last bit: a pixel on a rounding boundary may land in
                // the neighbouring bin (probabilities agree to ~1e-7, the stated tolerance is 1e-5).
                const float e = expf(__fsub_rn((float)v, (float)limit_0));
                const float rr = __fdiv_rn(rintf(__fmul_rn(e, 10000.0f)), 10000.0f);
                const double ps = (double)__fdiv_rn(1.0f, __fadd_rn(1.0f, rr));
#pragma unroll
                for (int j = 1; j < CHIPS_PROB_BINS; ++j)
                  bin += (ps >= __dadd_rn(__dmul_rn((double)j, __ddiv_rn(1.0, 20.0)), 0.0)) ? 1 : 0;
              } else {
                const double d = __dsub_rn(v, limit_0);
#pragma unroll
                for (int j = 1; j < CHIPS_PROB_BINS; ++j) bin += (d <= kProbCut[j]) ? 1 : 0;
              }
              atomicAdd(&tab[level * CHIPS_PROB_BINS + bin], 1u);
              level = max(level, t_inv);             // chips.py:372-374: lim < -1 wipes the mask
            }
          }
          packed |= (uint32_t)level << (8 * i);
        }
        if (vec) {                      // bytes outside the chord carry T, the pre-filled value
          *reinterpret_cast<uint32_t*>(lrow + x) = packed;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (x + i >= xa && x + i < xb) lrow[x + i] = (uint8_t)(packed >> (8 * i));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nT * CHIPS_PROB_BINS; i += blockDim.x)
    if (tab[i]) atomicAdd(&probtab[i], tab[i]);
}

__device__ double pairwise_sum20(const double* a, int n) {
  // numpy add.reduce for a contiguous float64 run with n < 128
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

__global__ void k_prob_epilogue(const uint32_t* __restrict__ probtab, double porb_threshold,
                                chips_result* __restrict__ res) {
  // the (T+1) x 20 table is read once, coalesced, into shared memory and prefix-summed over
  // the thresholds there (per-thread strided global reads made this tiny kernel take 14 us)
  __shared__ uint32_t stab[(CHIPS_MAX_T + 1) * CHIPS_PROB_BINS];
  const int T = res->n_thresholds;
  for (int i = threadIdx.x; i < (T + 1) * CHIPS_PROB_BINS; i += blockDim.x) stab[i] = probtab[i];
  __syncthreads();
  if (threadIdx.x < CHIPS_PROB_BINS) {
    uint32_t run = 0;
    for (int l = 0; l <= T; ++l) {
      run += stab[l * CHIPS_PROB_BINS + threadIdx.x];
      stab[l * CHIPS_PROB_BINS + threadIdx.x] = run;
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= T) return;
  if (res->n_peaks <= 0) {
    res->probs[t] = nan("");
    return;
  }
  // b = (e + diff(e)[0]/2)[:-1], e = np.linspace(0, 1, 21)
  const double stepe = __ddiv_rn(1.0, 20.0);
  const double halfw = __ddiv_rn(__dsub_rn(__dadd_rn(__dmul_rn(1.0, stepe), 0.0), 0.0), 2.0);
  double all[CHIPS_PROB_BINS], sel[CHIPS_PROB_BINS];
  int nsel = 0;
  for (int j = 0; j < CHIPS_PROB_BINS; ++j) {
    const long long h = stab[t * CHIPS_PROB_BINS + j];      // counts of levels 0..t
    double e = (j == 20) ? 1.0 : __dadd_rn(__dmul_rn((double)j, stepe), 0.0);
    double b = __dadd_rn(e, halfw);
    double bh = __dmul_rn(b, (double)h);
    all[j] = bh;
    if (b > porb_threshold) sel[nsel++] = bh;
  }
  double p = __ddiv_rn(pairwise_sum20(sel, nsel), pairwise_sum20(all, CHIPS_PROB_BINS));
  res->probs[t] = __ddiv_rn(rint(__dmul_rn(p, 10000.0)), 10000.0);
}

// ------------------------------------------------------------------ f64 -> f32 demotion
__global__ void k_demote(const double* __restrict__ in, float* __restrict__ out, size_t n,
                         int* __restrict__ all_exact) {
  bool ok = true;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    double v = in[i];
    float f = (float)v;
    out[i] = f;
    ok = ok && ((double)f == v);
  }
  if (!__all_sync(0xFFFFFFFFu, ok) && (threadIdx.x & 31) == 0) atomicAnd(all_exact, 0);
}

__global__ void k_set_int(int* p, int v) { *p = v; }

template <typename T>
__global__ void k_disk_mask(T* __restrict__ n_mask, T* __restrict__ i_mask, int H, int W, int cx,
                            int cy, int r) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)H * W) return;
  int y = (int)(i / W), x = (int)(i % W);
  bool on = on_disk(x, y, cx, cy, r);
  if (n_mask) n_mask[i] = on ? (T)1 : (T)NAN;
  if (i_mask) i_mask[i] = on ? (T)1 : (T)0;
}

static int hist_copies(int nbins) {
  int c = 65536 / (nbins * 4);
  if (c < 1) c = 1;
  if (c > 8) c = 8;
  return c;
}

template <typename T>
static int minmax_impl(const T* filt, const chips_plan* p, int cx, int cy, int r, unsigned char* ws,
                       cudaStream_t st) {
  using K = typename Key<T>::type;
  K* mm = reinterpret_cast<K*>(ws + p->off_minmax);
  k_minmax_init<T><<<1, 1, 0, st>>>(mm);
  CHIPS_CHECK_LAUNCH();
  int grid = min(p->roi_h, 4 * kNumSM);
  k_minmax<T><<<grid, 256, 0, st>>>(filt, p->W, p->roi_x0, p->roi_y0, p->roi_w, p->roi_h, cx, cy, r, mm);
  CHIPS_CHECK_LAUNCH();
  count_launch(2);
  return CHIPS_OK;
}

template <typename T>
static int hist_impl(const T* filt, const chips_plan* p, int cx, int cy, int r, unsigned char* ws,
                     cudaStream_t st) {
  using K = typename Key<T>::type;
  const int nbins = p->h_bins;
  int copies = hist_copies(nbins);
  const int use_table = nbins <= 4096;          // edges in shared memory (32 KB at most)
  size_t smem = (size_t)((copies * nbins + 1) & ~1) * sizeof(uint32_t) +
                (use_table ? (size_t)(nbins + 1) * sizeof(double) : 0);
  if (smem > 200 * 1024) return CHIPS_E_ARG;
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(ws + p->off_counts);
  CHIPS_CUDA(cudaMemsetAsync(counts, 0, (size_t)nbins * sizeof(long long), st));
  CHIPS_CUDA(cudaFuncSetAttribute(k_hist<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = smem > 100 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
  int grid = min(p->roi_h, per_sm * kNumSM);
  k_hist<T><<<grid, 256, smem, st>>>(filt, p->W, p->roi_x0, p->roi_y0, p->roi_w, p->roi_h, cx, cy, r,
                                     reinterpret_cast<const K*>(ws + p->off_minmax), nbins, copies, use_table,
                                     p->f32_math && sizeof(T) == 4, counts, reinterpret_cast<chips_result*>(ws + p->off_result));
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

template <typename T>
static int thr_impl(const T* filt, const chips_plan* p, int cx, int cy, int r, double porb,
                    unsigned char* ws, cudaStream_t st) {
  uint8_t* lev = ws + p->off_level;
  uint32_t* tab = reinterpret_cast<uint32_t*>(ws + p->off_probtab);
  chips_result* res = reinterpret_cast<chips_result*>(ws + p->off_result);
  CHIPS_CUDA(cudaMemsetAsync(lev, p->T, (size_t)p->H * p->W, st));
  CHIPS_CUDA(cudaMemsetAsync(tab, 0, (size_t)(p->T + 1) * CHIPS_PROB_BINS * sizeof(uint32_t), st));
  int grid = min(p->roi_h, 8 * kNumSM);
  k_threshold_prob<T><<<grid, 256, 0, st>>>(filt, p->W, p->roi_x0, p->roi_y0, p->roi_w, p->roi_h, cx,
                                            cy, r, p->f32_math && sizeof(T) == 4, res, lev, tab);
  CHIPS_CHECK_LAUNCH();
  k_prob_epilogue<<<1, CHIPS_MAX_T, 0, st>>>(tab, porb, res);
  CHIPS_CHECK_LAUNCH();
  count_launch(2);
  return CHIPS_OK;
}

// ------------------------------------------------------------------ sliding-window histograms
// SURVEY.md §8(f) row 3 / north_star (2)-(3): the reference's extract_sliding_histograms
// (chips.py:321-339) is an empty stub ("sliding a window (size determined by xsplit, ysplit)
// through the solar disk"); the per-window arithmetic is the loop body of extract_histograms
// (chips.py:282-300): np.histogram(bins, density=True) of the window's on-disk pixels over the
// window's OWN min/max range, a sticky h_thresh fixed by the first window, find_peaks, bound.
// PARITY UNPINNED (no reference output exists); pinned to oracle/chips_oracle.py::sliding_histograms.
// All windows in three launches: min/max (grid.y = window), histogram (warp-private shared-memory
// sub-histograms, grid.y = window), selection (ONE WARP per window).
struct WinRect { int x0, y0, x1, y1; };

template <typename T>
__global__ void k_win_init(typename Key<T>::type* minmax, int nwin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nwin) {
    minmax[2 * i] = Key<T>::kMax;
    minmax[2 * i + 1] = 0;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_win_minmax(const T* __restrict__ filt, int W, const WinRect* __restrict__ rects,
                                                    int cx, int cy, int r,
                                                    typename Key<T>::type* __restrict__ minmax) {
  using K = typename Key<T>::type;
  const WinRect q = rects[blockIdx.y];
  K kmin = Key<T>::kMax, kmax = 0;
  for (int y = q.y0 + blockIdx.x; y < q.y1; y += gridDim.x) {
    int xa, xb;
    if (!row_chord(y, cx, cy, r, q.x0, q.x1 - q.x0, xa, xb)) continue;
    const T* row = filt + (size_t)y * W;
    for (int x = xa + threadIdx.x; x < xb; x += blockDim.x) {
      const K key = Key<T>::enc(row[x]);
      kmin = key < kmin ? key : kmin;
      kmax = key > kmax ? key : kmax;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const K a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = b > kmax ? b : kmax;
  }
  if ((threadIdx.x & 31) == 0 && kmin <= kmax) {
    atomic_min_key(&minmax[2 * blockIdx.y], kmin);
    atomic_max_key(&minmax[2 * blockIdx.y + 1], kmax);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_win_hist(const T* __restrict__ filt, int W, const WinRect* __restrict__ rects,
                                                  int cx, int cy, int r,
                                                  const typename Key<T>::type* __restrict__ minmax, int nbins,
                                                  int copies, int f32_math,
                                                  unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(16) uint32_t sh[];
  const WinRect q = rects[blockIdx.y];
  for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh[i] = 0;
  const typename Key<T>::type k0 = minmax[2 * blockIdx.y], k1 = minmax[2 * blockIdx.y + 1];
  const Edges edge = make_edges((double)Key<T>::dec(k0), (double)Key<T>::dec(k1), k0 > k1, nbins, f32_math);
  const double denom = __dsub_rn(edge.last, edge.first);
  const float first_f = (float)edge.first, scale_f = (float)((double)nbins / denom);
  __syncthreads();
  uint32_t* mine = sh + ((threadIdx.x >> 5) % copies) * nbins;      // warp-private bins
  for (int y = q.y0 + blockIdx.x; y < q.y1; y += gridDim.x) {
    int xa, xb;
    if (!row_chord(y, cx, cy, r, q.x0, q.x1 - q.x0, xa, xb)) continue;
    const T* row = filt + (size_t)y * W;
    for (int x0 = xa; x0 < xb; x0 += blockDim.x) {
      const int x = x0 + threadIdx.x;
      int bin = -1;
      if (x < xb) bin = hist_bin((double)row[x], nbins, edge, first_f, scale_f);
      // neighbouring pixels of the filtered image mostly share a bin: one shared-memory add per
      // run of equal bins inside the warp (shuffle with the left neighbour, ballot of run starts)
      const int left = __shfl_up_sync(0xFFFFFFFFu, bin, 1);
      const bool head = bin >= 0 && ((threadIdx.x & 31) == 0 || left != bin);
      const uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
      const uint32_t valid = __ballot_sync(0xFFFFFFFFu, bin >= 0);
      if (head) {
        const int lane = threadIdx.x & 31;
        const uint32_t later = heads & (0xFFFFFFFEu << lane);          // run heads after me
        const int end = later ? __ffs(later) - 1 : 32;
        const uint32_t run = valid & (0xFFFFFFFFu >> (32 - end)) & (0xFFFFFFFFu << lane);
        atomicAdd(&mine[bin], (unsigned)__popc(run));
      }
    }
  }
  __syncthreads();
  unsigned long long* out = counts + (size_t)blockIdx.y * nbins;
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
    unsigned long long s = 0;
    for (int c = 0; c < copies; ++c) s += sh[c * nbins + i];
    if (s) atomicAdd(&out[i], s);
  }
}

// one warp per window
template <typename T>
__global__ void __launch_bounds__(128) k_win_select(const unsigned long long* __restrict__ counts,
                                                    const typename Key<T>::type* __restrict__ minmax, int nwin,
                                                    int nbins, double ratio, double h_thresh_in, int f32_math,
                                                    double* __restrict__ density, double* __restrict__ bound,
                                                    int* __restrict__ peaks, chips_window_rec* __restrict__ recs) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nwin) return;
  // density of window `v` into `dst` (or nowhere); returns max density and the sample count
  auto fill = [&](int v, double* dst, double* bnd, double& first, double& last, long long& total) {
    const typename Key<T>::type k0 = minmax[2 * v], k1 = minmax[2 * v + 1];
    const Edges edge = make_edges((double)Key<T>::dec(k0), (double)Key<T>::dec(k1), k0 > k1, nbins, f32_math);
    first = edge.first;
    last = edge.last;
    const unsigned long long* c = counts + (size_t)v * nbins;
    long long part = 0;
    for (int i = lane; i < nbins; i += 32) part += (long long)c[i];
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    total = part;
    double dmax = -INFINITY;
    for (int i = lane; i < nbins; i += 32) {
      const double e0 = edge(i), e1 = edge(i + 1);
      const double db = edge.width(e0, e1);
      const double d = __ddiv_rn(__ddiv_rn((double)c[i], db), (double)part);   // n / db / n.sum()
      if (dst) {
        dst[i] = d;
        bnd[i] = edge.upper(e0, db);                                            // be[:-1] + diff(be)
      }
      dmax = fmax(dmax, d);          // NaN densities (empty window: 0 / 0) are skipped, as np.max would not
    }
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xFFFFFFFFu, dmax, o));
    return dmax;
  };
  double first, last;
  long long total;
  double hthr = h_thresh_in;
  if (isnan(hthr)) {     // sticky: fixed by the first window of the loop (chips.py:295-296)
    double f0, l0;
    long long t0;
    const double m0 = fill(0, nullptr, nullptr, f0, l0, t0);
    hthr = t0 > 0 ? __ddiv_rn(m0, ratio) : nan("");    // np.max of an all-NaN density is NaN
  }
  double* dst = density + (size_t)w * nbins;
  double* bnd = bound + (size_t)w * nbins;
  fill(w, dst, bnd, first, last, total);
  __syncwarp();
  if (lane == 0) {       // scipy.signal.find_peaks(h, height=hthr): plateau midpoints, in order
    int np_ = 0;
    int* pk = peaks + (size_t)w * nbins;
    int i = 1;
    const int imax = nbins - 1;
    while (i < imax) {
      const double di = dst[i];
      if (dst[i - 1] < di) {
        int ia = i + 1;
        while (ia < imax && dst[ia] == di) ++ia;
        if (dst[ia] < di) {
          const int p = (i + ia - 1) / 2;
          if (dst[p] >= hthr) pk[np_++] = p;
          i = ia;
        }
      }
      ++i;
    }
    chips_window_rec rc;
    rc.first_edge = first;
    rc.last_edge = last;
    rc.h_thresh = hthr;
    rc.n_samples = total;
    rc.n_peaks = np_;
    rc.pad_ = 0;
    recs[w] = rc;
  }
}

template <typename T>
static int windows_impl(const T* filt, int H, int W, int f32_math, int cx, int cy, int r, int nwin,
                        const int32_t* rects_host,
                        int nbins, double ratio, double h_in, unsigned long long* counts, double* density,
                        double* bound, int32_t* peaks, chips_window_rec* recs, unsigned char* ws, cudaStream_t st) {
  using K = typename Key<T>::type;
  K* minmax = reinterpret_cast<K*>(ws);
  WinRect* rects = reinterpret_cast<WinRect*>(ws + (size_t)nwin * 16);
  for (int i = 0; i < nwin; ++i) {
    const int32_t* q = rects_host + 4 * i;
    if (q[0] < 0 || q[1] < 0 || q[2] > W || q[3] > H || q[2] < q[0] || q[3] < q[1]) return CHIPS_E_ARG;
  }
  CHIPS_CUDA(cudaMemcpyAsync(rects, rects_host, (size_t)nwin * sizeof(WinRect), cudaMemcpyHostToDevice, st));
  CHIPS_CUDA(cudaMemsetAsync(counts, 0, (size_t)nwin * nbins * sizeof(long long), st));
  k_win_init<T><<<(nwin + 127) / 128, 128, 0, st>>>(minmax, nwin);
  CHIPS_CHECK_LAUNCH();
  int copies = hist_copies(nbins);
  size_t smem = (size_t)copies * nbins * sizeof(uint32_t);
  if (smem > 200 * 1024) return CHIPS_E_ARG;
  int gx = (4 * kNumSM + nwin - 1) / nwin;
  if (gx > H) gx = H;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)nwin);
  k_win_minmax<T><<<grid, 256, 0, st>>>(filt, W, rects, cx, cy, r, minmax);
  CHIPS_CHECK_LAUNCH();
  CHIPS_CUDA(cudaFuncSetAttribute(k_win_hist<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_win_hist<T><<<grid, 256, smem, st>>>(filt, W, rects, cx, cy, r, minmax, nbins, copies, f32_math, counts);
  CHIPS_CHECK_LAUNCH();
  k_win_select<T><<<(nwin + 3) / 4, 128, 0, st>>>(counts, minmax, nwin, nbins, ratio, h_in, f32_math, density, bound,
                                                  peaks, recs);
  CHIPS_CHECK_LAUNCH();
  count_launch(4);
  return CHIPS_OK;
}

}  // namespace chips

using namespace chips;

extern "C" size_t chips_window_workspace_bytes(int nwin) { return (size_t)(nwin > 0 ? nwin : 0) * 32 + 256; }

extern "C" int chips_window_histograms(const void* filt, int H, int W, int elem, int f32_math, int cx, int cy,
                                       int r, int nwin, const int32_t* rects, int h_bins, double ht_peak_ratio,
                                       double h_thresh_in,
                                       long long* counts, double* density, double* bound, int32_t* peaks,
                                       chips_window_rec* recs, void* ws, void* stream) {
  if (!filt || !rects || !counts || !density || !bound || !peaks || !recs || !ws) return CHIPS_E_ARG;
  if (H < 1 || W < 1 || nwin < 1 || nwin > 65535 || h_bins < 1) return CHIPS_E_ARG;
  unsigned long long* c = reinterpret_cast<unsigned long long*>(counts);
  if (elem == 4)
    return windows_impl<float>((const float*)filt, H, W, f32_math != 0, cx, cy, r, nwin, rects, h_bins, ht_peak_ratio, h_thresh_in,
                               c, density, bound, peaks, recs, (unsigned char*)ws, (cudaStream_t)stream);
  if (elem == 8)
    return windows_impl<double>((const double*)filt, H, W, 0, cx, cy, r, nwin, rects, h_bins, ht_peak_ratio,
                                h_thresh_in, c, density, bound, peaks, recs, (unsigned char*)ws, (cudaStream_t)stream);
  return CHIPS_E_ARG;
}

extern "C" int chips_minmax(const void* filt, const chips_plan* plan, int cx, int cy, int r,
                            void* ws, void* stream) {
  if (!filt || !plan || !ws) return CHIPS_E_ARG;
  if (plan->elem == 4) return minmax_impl<float>((const float*)filt, plan, cx, cy, r, (unsigned char*)ws, (cudaStream_t)stream);
  if (plan->elem == 8) return minmax_impl<double>((const double*)filt, plan, cx, cy, r, (unsigned char*)ws, (cudaStream_t)stream);
  return CHIPS_E_ARG;
}

extern "C" int chips_histogram(const void* filt, const chips_plan* plan, int cx, int cy, int r,
                               void* ws, void* stream) {
  if (!filt || !plan || !ws || plan->h_bins < 1) return CHIPS_E_ARG;
  if (plan->elem == 4) return hist_impl<float>((const float*)filt, plan, cx, cy, r, (unsigned char*)ws, (cudaStream_t)stream);
  if (plan->elem == 8) return hist_impl<double>((const double*)filt, plan, cx, cy, r, (unsigned char*)ws, (cudaStream_t)stream);
  return CHIPS_E_ARG;
}

extern "C" int chips_select_threshold(const chips_plan* plan, double ht_peak_ratio,
                                      double h_thresh_in, const double* offsets, void* ws, void* stream) {
  if (!plan || !ws || !offsets) return CHIPS_E_ARG;
  const int T = plan->T;
  if (T < 1 || T > CHIPS_MAX_T) return CHIPS_E_ARG;
  ThresholdOffsets offs;
  for (int i = 0; i < CHIPS_MAX_T; ++i) offs.v[i] = i < T ? offsets[i] : 0.0;
  unsigned char* w = (unsigned char*)ws;
  k_select<<<1, 1024, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(w + plan->off_counts), plan->h_bins, ht_peak_ratio,
      h_thresh_in, offs, T, plan->f32_math, reinterpret_cast<double*>(w + plan->off_density),
      reinterpret_cast<double*>(w + plan->off_edges), reinterpret_cast<double*>(w + plan->off_bound),
      reinterpret_cast<int*>(w + plan->off_peaks), reinterpret_cast<chips_result*>(w + plan->off_result));
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_threshold_prob(const void* filt, const chips_plan* plan, int cx, int cy, int r,
                                    double porb_threshold, void* ws, void* stream) {
  if (!filt || !plan || !ws) return CHIPS_E_ARG;
  if (plan->elem == 4) return thr_impl<float>((const float*)filt, plan, cx, cy, r, porb_threshold, (unsigned char*)ws, (cudaStream_t)stream);
  if (plan->elem == 8) return thr_impl<double>((const double*)filt, plan, cx, cy, r, porb_threshold, (unsigned char*)ws, (cudaStream_t)stream);
  return CHIPS_E_ARG;
}

extern "C" int chips_disk_mask(void* n_mask, void* i_mask, int H, int W, int elem, int cx, int cy,
                               int r, void* stream) {
  size_t n = (size_t)H * W;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (elem == 4)
    k_disk_mask<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)n_mask, (float*)i_mask, H, W, cx, cy, r);
  else if (elem == 8)
    k_disk_mask<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)n_mask, (double*)i_mask, H, W, cx, cy, r);
  else
    return CHIPS_E_ARG;
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_demote_f64(const double* in, float* out, size_t n, int32_t* all_exact,
                                void* stream) {
  if (!in || !out || !all_exact) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  k_set_int<<<1, 1, 0, st>>>(all_exact, 1);
  CHIPS_CHECK_LAUNCH();
  k_demote<<<8 * kNumSM, 256, 0, st>>>(in, out, n, all_exact);
  CHIPS_CHECK_LAUNCH();
  count_launch(2);
  return CHIPS_OK;
}
