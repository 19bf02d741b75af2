// export.cu — output-side helpers of the detection path.
//
//   k_pack_cube      the `[res, res, L]` float32 region cube that Chips.to_netcdf stores
//                    (/root/reference/chips/chips.py:570-586: data[:, :, i] = region.map), written
//                    straight from the level-encoded result image: HBM-write bound, one warp
//                    = 32 pixels = 32*T consecutive floats stored as aligned 128-bit words.
//   k_row_cosine     Chips.compute_similarity_measures (chips.py:592-617): for every row i with
//                    sum(map[i]) > 0 the cosine similarity of map[i] and map0[i] as
//                    sklearn.metrics.pairwise.cosine_similarity computes it (rows L2-normalised,
//                    zero norms replaced by 1, then the dot product); k_cosine_mean = np.nanmean.
#include <math.h>
#include "common.cuh"

namespace chips {

__global__ void __launch_bounds__(256) k_pack_cube(const uint8_t* __restrict__ src, size_t npix, int T,
                                                   float* __restrict__ cube) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t p0 = warp * 32;
  if (p0 >= npix) return;
  const size_t p = p0 + lane;
  const int mine = p < npix ? (int)src[p] : 255;          // first threshold whose map holds the pixel
  const int nvalid = (int)(npix - p0 < 32 ? npix - p0 : 32);
  const int nquad = nvalid * T / 4;                        // whole float4 words of this warp's run
  float* base = cube + p0 * (size_t)T;                     // 32*T floats per warp: 128-byte aligned
  for (int q0 = 0; q0 < 8 * T; q0 += 32) {                // uniform trip count: shuffles below
    const int q = q0 + lane;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = 4 * q + j, px = e / T, t = e - px * T;
      const int first = __shfl_sync(0xFFFFFFFFu, mine, px & 31);
      v[j] = first <= t ? 1.0f : 0.0f;
    }
    if (q < nquad) {
      *reinterpret_cast<float4*>(base + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (q < 8 * T) {                                // ragged tail of the last warp
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < nvalid * T) base[4 * q + j] = v[j];
    }
  }
}

// one CTA per row: out[row] = cosine similarity, NaN for rows the reference skips
template <typename T>
__global__ void __launch_bounds__(256) k_row_cosine(const T* __restrict__ a, const T* __restrict__ b, int W,
                                                    double* __restrict__ out) {
  __shared__ double red[3][8];
  __shared__ double na_s, nb_s;
  const T* ra = a + (size_t)blockIdx.x * W;
  const T* rb = b + (size_t)blockIdx.x * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto block_sum3 = [&](double& x, double& y, double& z) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
      y += __shfl_xor_sync(0xFFFFFFFFu, y, o);
      z += __shfl_xor_sync(0xFFFFFFFFu, z, o);
    }
    __syncthreads();
    if (lane == 0) {
      red[0][warp] = x;
      red[1][warp] = y;
      red[2][warp] = z;
    }
    __syncthreads();
    x = y = z = 0.0;
    for (int i = 0; i < 8; ++i) {
      x += red[0][i];
      y += red[1][i];
      z += red[2][i];
    }
  };
  double s = 0.0, aa = 0.0, bb = 0.0;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const double x = (double)ra[i], y = (double)rb[i];
    s += x;
    aa += x * x;
    bb += y * y;
  }
  block_sum3(s, aa, bb);
  if (threadIdx.x == 0) {
    double na = sqrt(aa), nb = sqrt(bb);
    na_s = na == 0.0 ? 1.0 : na;         // sklearn.preprocessing.normalize: zero norms -> 1
    nb_s = nb == 0.0 ? 1.0 : nb;
  }
  __syncthreads();
  const double na = na_s, nb = nb_s;
  double dot = 0.0, z1 = 0.0, z2 = 0.0;
  for (int i = threadIdx.x; i < W; i += blockDim.x) dot += ((double)ra[i] / na) * ((double)rb[i] / nb);
  block_sum3(dot, z1, z2);
  if (threadIdx.x == 0) out[blockIdx.x] = (s > 0.0) ? dot : nan("");   // `if np.sum(map[i, :]) > 0`
}

// np.nanmean of the appended similarities: out[H] = mean, out[H+1] = number of rows used
__global__ void __launch_bounds__(1024) k_cosine_mean(double* __restrict__ out, int H) {
  __shared__ double sred[32];
  __shared__ int cred[32];
  double s = 0.0;
  int c = 0;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const double v = out[i];
    if (v == v) {
      s += v;
      ++c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sred[threadIdx.x >> 5] = s;
    cred[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.0;
    c = 0;
    for (int i = 0; i < 32; ++i) {
      s += sred[i];
      c += cred[i];
    }
    out[H] = c ? s / (double)c : nan("");
    out[H + 1] = (double)c;
  }
}

// a device-sized copy: `*count` words (at most cap) from device memory to mapped pinned host memory
// (or anywhere else) with a handful of CTAs - 128-bit coalesced stores, posted over PCIe - so that
// the copy neither needs the size on the host nor takes SMs from the detections in flight
__global__ void __launch_bounds__(256) k_copy_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                    const uint32_t* __restrict__ count, uint32_t cap,
                                                    uint32_t* __restrict__ count_out) {
  uint32_t n = *count;
  if (blockIdx.x == 0 && threadIdx.x == 0 && count_out) *count_out = n;
  if (n > cap) n = cap;
  const uint32_t n4 = n >> 2;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) d4[i] = s4[i];
  for (uint32_t i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// ------------------------------------------------------------------ sparse result map
// The level-encoded result `keep` is T ("in no map") on ~97 % of a 4096^2 image; only the
// pixels with keep < T are worth a device -> host copy.  entry = (pixel index << 6) | keep.
// Warp-aggregated append (one atomic per warp that has anything); the order of the entries is
// arbitrary, the decoder scatters them.  `entries` may be mapped pinned host memory: the appends
// are coalesced runs of up to 32 words.
__global__ void __launch_bounds__(256) k_pack_keep(const uint8_t* __restrict__ keep, int W, int rx0, int ry0,
                                                   int rw, int rh, int T, uint32_t* __restrict__ entries,
                                                   uint32_t cap, uint32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int words = (rw + (rx0 & 3) + 3) >> 2;               // aligned 4-pixel words per ROI row
  const long long total = (long long)words * rh;
  for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < total;
       i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + lane;
    uint32_t m = 0, v = 0;
    int x = 0, y = 0;
    if (i < total) {
      y = ry0 + (int)(i / words);
      x = (rx0 & ~3) + 4 * (int)(i % words);
      v = *reinterpret_cast<const uint32_t*>(keep + (size_t)y * W + x);   // W % 4 == 0 is checked by the host
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xb = x + b;
        if (xb >= rx0 && xb < rx0 + rw && ((v >> (8 * b)) & 255u) < (uint32_t)T) m |= 1u << b;
      }
    }
    const uint32_t any = __ballot_sync(0xFFFFFFFFu, m != 0);
    if (!any) continue;
    const int c = __popc(m);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
      if (lane >= o) inc += u;
    }
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(count, (uint32_t)inc);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + (uint32_t)(inc - c);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (m & (1u << b)) {
        if (base < cap) entries[base] = ((uint32_t)(y * W + x + b) << 6) | ((v >> (8 * b)) & 255u);
        ++base;
      }
    }
  }
}

}  // namespace chips

using namespace chips;

extern "C" int chips_pack_cube(const chips_plan* plan, int which, float* cube, void* ws, void* stream) {
  if (!plan || !cube || !ws || which < 0 || which > 2) return CHIPS_E_ARG;
  const size_t n = (size_t)plan->H * plan->W;
  const unsigned char* w = (const unsigned char*)ws;
  const uint8_t* src = w + (which == 1 ? plan->off_level : (which == 2 ? plan->off_fill : plan->off_keep));
  const size_t warps = (n + 31) / 32;
  k_pack_cube<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(src, n, plan->T, cube);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_row_cosine(const void* map, const void* map0, int H, int W, int elem, double* out,
                                void* stream) {
  if (!map || !map0 || !out || H < 1 || W < 1) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (elem == 4)
    k_row_cosine<float><<<H, 256, 0, st>>>((const float*)map, (const float*)map0, W, out);
  else if (elem == 8)
    k_row_cosine<double><<<H, 256, 0, st>>>((const double*)map, (const double*)map0, W, out);
  else
    return CHIPS_E_ARG;
  CHIPS_CHECK_LAUNCH();
  k_cosine_mean<<<1, 1024, 0, st>>>(out, H);
  CHIPS_CHECK_LAUNCH();
  count_launch(2);
  return CHIPS_OK;
}

namespace chips {
// more entries than `cap` words: the words buffer receives the dense map instead (H * W bytes fit
// when cap >= H * W / 4); *count > cap tells the reader which of the two it holds
__global__ void __launch_bounds__(256) k_keep_dense_if_overflow(const uint8_t* __restrict__ keep, size_t nbytes,
                                                                uint32_t* __restrict__ words, uint32_t cap,
                                                                const uint32_t* __restrict__ count) {
  if (*count <= cap) return;
  const uint4* s = reinterpret_cast<const uint4*>(keep);
  uint4* d = reinterpret_cast<uint4*>(words);
  const size_t n16 = nbytes >> 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    d[i] = s[i];
  for (size_t i = (n16 << 4) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes;
       i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint8_t*>(words)[i] = keep[i];
}
}  // namespace chips

extern "C" int chips_pack_keep(const chips_plan* plan, uint32_t* entries, uint32_t cap, uint32_t* count,
                               void* ws, void* stream) {
  if (!plan || !entries || !count || !ws) return CHIPS_E_ARG;
  if ((long long)plan->H * plan->W > (1LL << 26) || plan->T > 64 || (plan->W & 3)) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CHIPS_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), st));
  const uint8_t* keep = (const uint8_t*)ws + plan->off_keep;
  k_pack_keep<<<4 * kNumSM, 256, 0, st>>>(keep, plan->W, plan->roi_x0, plan->roi_y0, plan->roi_w,
                                                plan->roi_h, plan->T, entries, cap, count);
  CHIPS_CHECK_LAUNCH();
  chips::count_launch();
  const size_t nbytes = (size_t)plan->H * plan->W;
  if ((size_t)cap * 4 >= nbytes && !(((uintptr_t)entries) & 15)) {
    chips::k_keep_dense_if_overflow<<<4 * kNumSM, 256, 0, st>>>(keep, nbytes, entries, cap, count);
    CHIPS_CHECK_LAUNCH();
    chips::count_launch();
  }
  return CHIPS_OK;
}

extern "C" int chips_copy_words(const uint32_t* src, uint32_t* dst, const uint32_t* count, uint32_t cap,
                                uint32_t* count_out, void* stream) {
  if (!src || !dst || !count) return CHIPS_E_ARG;
  if (((uintptr_t)src | (uintptr_t)dst) & 15) return CHIPS_E_ARG;
  chips::k_copy_words<<<8, 256, 0, (cudaStream_t)stream>>>(src, dst, count, cap, count_out);
  CHIPS_CHECK_LAUNCH();
  chips::count_launch();
  return CHIPS_OK;
}
