// pipeline.cu — workspace plan and the fused per-image detection (rows B..I of
// SURVEY.md §8(a)): every stage is enqueued on one stream, all intermediate scalars
// (histogram range, h_thresh, limit_0, limits, probabilities) stay in device memory, so a
// detection is a fixed launch sequence with no host round trip (CUDA-graph capturable).
#include "common.cuh"

using namespace chips;

namespace chips {
std::atomic<long long> g_launches{0};
}

extern "C" long long chips_launch_count(void) { return g_launches.load(); }

extern "C" int chips_version(void) { return 100; }

extern "C" int chips_plan_init(chips_plan* p, int H, int W, int k, int h_bins, int T, int elem,
                               int cx, int cy, int r) {
  if (!p || H < 1 || W < 1 || k < 1 || !(k & 1) || k > CHIPS_MAX_K) return CHIPS_E_ARG;
  if (h_bins < 1 || T < 1 || T > CHIPS_MAX_T || (elem != 4 && elem != 8)) return CHIPS_E_ARG;
  if ((long long)H * W >= (1LL << 31)) return CHIPS_E_ARG;
  p->H = H; p->W = W; p->k = k; p->h_bins = h_bins; p->T = T; p->elem = elem;
  p->masked = r >= 0; p->f32_math = 0;
  if (r >= 0) {
    int x0 = cx - r, x1 = cx + r + 1, y0 = cy - r, y1 = cy + r + 1;
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > W) x1 = W;
    if (y1 > H) y1 = H;
    if (x1 <= x0 || y1 <= y0) return CHIPS_E_ARG;
    p->roi_x0 = x0; p->roi_y0 = y0; p->roi_w = x1 - x0; p->roi_h = y1 - y0;
  } else {
    p->roi_x0 = 0; p->roi_y0 = 0; p->roi_w = W; p->roi_h = H;
  }
  const int half = k >> 1;
  {  // median blocks (median.cu): 128 columns x L rows, L a multiple of the 16-pixel select tile.
     // Every tier-1 run pays a (k-1)-row warm-up; one warp owns one block and 8 warps (24 KB of
     // histograms each) are resident per SM: minimise waves(L) * (L + k).  The block sort holds a
     // block's halo-extended region in shared memory: (128 + k - 1) * (L + k - 1) pixels at most
     // CHIPS_MAX_BLOCK_ELEMS_F32 / _F64.
    const int nbx = (p->roi_w + 127) / 128;
    const long long resident = 8LL * kNumSM;
    const int emax = elem == 4 ? CHIPS_MAX_BLOCK_ELEMS_F32 : CHIPS_MAX_BLOCK_ELEMS_F64;
    int lcap = (emax / (128 + 2 * half) - 2 * half) & ~15;
    if (lcap < 16) return CHIPS_E_ARG;
    int lmax = (int)align_up((size_t)p->roi_h, 16);
    if (lmax > lcap) lmax = lcap;
    int L = lmax;
    long long best = -1;
    int cand0 = (int)align_up((size_t)(k < 16 ? 16 : k), 16);
    if (cand0 > lmax) cand0 = lmax;
    for (int cand = cand0; cand <= lmax; cand += 16) {
      long long blocks = (long long)nbx * ((p->roi_h + cand - 1) / cand);
      long long waves = (blocks + resident - 1) / resident;
      long long cost = waves * (cand + k);
      if (best < 0 || cost < best) {
        best = cost;
        L = cand;
      }
    }
    p->blk_rows = L;
    p->blk_nbx = nbx;
    p->blk_nby = (p->roi_h + L - 1) / L;
    // + 16 elements of slack: funnel loads of tier 1, 8-element row reads of tier 2
    p->bp_pitch = (int)align_up((size_t)128 + 2 * half + 16, 16);
    p->bp_rows = L + 2 * half;
    // 128 buckets of equal population: bucket = rank / blk_bucket
    p->blk_bucket = ((128 + 2 * half) * (L + 2 * half) + 127) / 128;
  }
  const size_t nblk = (size_t)p->blk_nbx * p->blk_nby;
  const size_t n = (size_t)H * W, nr = (size_t)p->roi_w * p->roi_h;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p->off_order = take(nblk * align_up((size_t)(128 + 2 * half) * p->bp_rows, 8) * 2);
  p->off_bucket = take(nblk * p->bp_pitch * p->bp_rows);
  p->off_bstar = take(n);
  p->off_rres = take(n * 4);
  p->off_filt = take(n * elem);
  p->off_minmax = take(16);
  p->off_counts = take((size_t)h_bins * 8);
  p->off_density = take((size_t)h_bins * 8);
  p->off_edges = take((size_t)(h_bins + 1) * 8);
  p->off_bound = take((size_t)h_bins * 8);
  p->off_peaks = take((size_t)h_bins * 4);
  p->off_result = take(sizeof(chips_result));
  p->off_level = take(n);
  p->off_fill = take(n);
  p->off_keep = take(n);
  p->off_probtab = take((size_t)(T + 1) * CHIPS_PROB_BINS * 4);
  if (p->masked) {
    p->off_parent_b = take((nr + 1) * 4);
    p->off_parent_c = take(nr * 4 * 2);       // union-find + merge-tree parents
    p->off_area = take(nr * 4 + nr * 3);      // 2*area accumulators; hook level, kept level, local W
    p->off_pending = take(nr * 4);
    p->off_list2 = take(nr * 4);
    p->off_lvl = take(512 * 4);
    p->off_tiles = take(256);                  // (reserved)
  } else {   // the synoptic variant has no contour/cleanup step (syn_chips.py:237-251)
    p->off_parent_b = p->off_parent_c = p->off_area = off;
    p->off_pending = p->off_list2 = p->off_lvl = p->off_tiles = off;
  }
  p->off_ovf = take(256);                      // (reserved)
  p->bytes = off;
  return CHIPS_OK;
}

__global__ void k_copy_level_to_keep(const uint8_t* __restrict__ lev, uint8_t* __restrict__ keep,
                                     size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i + 16 <= n) {
    *reinterpret_cast<uint4*>(keep + i) = *reinterpret_cast<const uint4*>(lev + i);
  } else {
    for (; i < n; ++i) keep[i] = lev[i];
  }
}

extern "C" int chips_detect(const void* image, const chips_plan* plan, int cx, int cy, int r,
                            double ht_peak_ratio, double h_thresh_in, const double* offsets,
                            double porb_threshold, double area_threshold, uint8_t* maps_or_null,
                            void* ws, void* stream) {
  if (!image || !plan || !ws || !offsets) return CHIPS_E_ARG;
  unsigned char* w = (unsigned char*)ws;
  void* filt = w + plan->off_filt;
  int rc;
  if ((rc = chips_medfilt2d(image, filt, plan, cx, cy, r, ws, stream))) return rc;
  if (plan->k == 1 && (rc = chips_minmax(filt, plan, cx, cy, r, ws, stream))) return rc;
  if ((rc = chips_histogram(filt, plan, cx, cy, r, ws, stream))) return rc;
  if ((rc = chips_select_threshold(plan, ht_peak_ratio, h_thresh_in, offsets, ws, stream))) return rc;
  if ((rc = chips_threshold_prob(filt, plan, cx, cy, r, porb_threshold, ws, stream))) return rc;
  if (plan->masked) {
    double disk_area = 3.141592653589793 * (double)((long long)r * r);   // np.pi * r**2
    if ((rc = chips_cleanup(plan, cx, cy, r, disk_area, area_threshold, ws, stream))) return rc;
  } else {
    size_t n = (size_t)plan->H * plan->W;
    k_copy_level_to_keep<<<(unsigned)((n / 16 + 256) / 256), 256, 0, (cudaStream_t)stream>>>(
        w + plan->off_level, w + plan->off_keep, n);
    CHIPS_CHECK_LAUNCH();
    count_launch();
  }
  if (maps_or_null && (rc = chips_expand_maps(plan, 0, maps_or_null, ws, stream))) return rc;
  return CHIPS_OK;
}

// ---- host <-> device copies of only what the path reads ------------------------------------
// Every kernel that reads the input image stays inside the halo-extended median blocks (block
// sort, the selected medians, the k <= 5 windows), clipped to the image: that rectangle
// is all a caller has to upload.  4096^2, r = 1577, k = 51: 3250 x 3218 of 4096 x 4096 = 62 %.
extern "C" int chips_input_rect(const chips_plan* plan, int32_t* rect) {
  if (!plan || !rect) return CHIPS_E_ARG;
  const int half = plan->k / 2;
  const int bw = plan->blk_nbx * 128 > plan->roi_w ? plan->blk_nbx * 128 : plan->roi_w;
  const int bh = plan->blk_nby * plan->blk_rows > plan->roi_h ? plan->blk_nby * plan->blk_rows : plan->roi_h;
  const int x0 = plan->roi_x0 - half, y0 = plan->roi_y0 - half;
  const int x1 = plan->roi_x0 + bw + half, y1 = plan->roi_y0 + bh + half;
  rect[0] = x0 < 0 ? 0 : x0;
  rect[1] = y0 < 0 ? 0 : y0;
  rect[2] = x1 > plan->W ? plan->W : x1;
  rect[3] = y1 > plan->H ? plan->H : y1;
  return CHIPS_OK;
}

// The same, tighter: row intervals [y0, y1) with their own column range [x0, x1).  Only the
// halo-extended regions of the blocks that hold an on-disk pixel are read (k_block_sort, and the
// medians k_rank_select fetches from them): per band of blk_rows rows the on-disk blocks are a
// contiguous run, so the union is one column range per row interval - 52 % of a 4096^2 image at
// r = 1577, k = 51 against 63 % for the rectangle.  Returns the number of spans (<= max_spans), or
// a negative error; spans are disjoint and ascending in y.
static bool host_rect_on_disk(int x0, int y0, int w, int h, int cx, int cy, int r) {   // median.cu rect_on_disk
  if (r < 0) return true;
  const long long nx = cx < x0 ? x0 : (cx > x0 + w - 1 ? x0 + w - 1 : cx);
  const long long ny = cy < y0 ? y0 : (cy > y0 + h - 1 ? y0 + h - 1 : cy);
  return (nx - cx) * (nx - cx) + (ny - cy) * (ny - cy) <= (long long)r * r;
}

extern "C" int chips_input_spans(const chips_plan* plan, int cx, int cy, int r, int32_t* spans, int max_spans) {
  if (!plan || !spans || max_spans < 1) return CHIPS_E_ARG;
  int32_t rect[4];
  chips_input_rect(plan, rect);
  const int half = plan->k / 2, L = plan->blk_rows;
  if (!plan->masked || plan->k <= 5 || r < 0) {          // whole rectangle (the k <= 5 kernels read windows, not blocks)
    spans[0] = rect[1]; spans[1] = rect[3]; spans[2] = rect[0]; spans[3] = rect[2];
    return 1;
  }
  // column range of every band, then the union over the (at most two or three) bands a row belongs to
  const int nby = plan->blk_nby, nbx = plan->blk_nbx;
  if (nby > 4096) return CHIPS_E_ARG;
  static thread_local int band_xa[4096], band_xb[4096];
  for (int by = 0; by < nby; ++by) {
    const int Y = plan->roi_y0 + L * by;
    band_xa[by] = 1 << 30;
    band_xb[by] = -1;
    for (int bx = 0; bx < nbx; ++bx) {
      const int X = plan->roi_x0 + 128 * bx;
      if (!host_rect_on_disk(X, Y, 128, L, cx, cy, r)) continue;
      if (X - half < band_xa[by]) band_xa[by] = X - half;
      if (X + 128 + half > band_xb[by]) band_xb[by] = X + 128 + half;
    }
  }
  int n = 0;
  int cur_x0 = 0, cur_x1 = 0, cur_y0 = 0;
  bool open = false;
  for (int y = rect[1]; y <= rect[3]; ++y) {
    int xa = 1 << 30, xb = -1;
    if (y < rect[3]) {
      int b0 = (y - half - plan->roi_y0) / L - 1, b1 = (y + half - plan->roi_y0) / L + 1;
      if (b0 < 0) b0 = 0;
      if (b1 >= nby) b1 = nby - 1;
      for (int by = b0; by <= b1; ++by) {
        const int Y = plan->roi_y0 + L * by;
        if (y < Y - half || y >= Y + L + half || band_xb[by] < 0) continue;
        if (band_xa[by] < xa) xa = band_xa[by];
        if (band_xb[by] > xb) xb = band_xb[by];
      }
      if (xa < rect[0]) xa = rect[0];
      if (xb > rect[2]) xb = rect[2];
    }
    const bool has = xb > xa;
    if (open && (!has || xa != cur_x0 || xb != cur_x1)) {
      if (n >= max_spans) return CHIPS_E_WS;
      spans[4 * n + 0] = cur_y0; spans[4 * n + 1] = y; spans[4 * n + 2] = cur_x0; spans[4 * n + 3] = cur_x1;
      ++n;
      open = false;
    }
    if (has && !open) {
      open = true;
      cur_y0 = y; cur_x0 = xa; cur_x1 = xb;
    }
  }
  return n;
}

extern "C" int chips_copy_rect(void* dst, const void* src, size_t pitch_bytes, size_t x0_bytes, int y0,
                               size_t width_bytes, int rows, int to_device, void* stream) {
  if (!dst || !src || rows < 1 || y0 < 0 || width_bytes < 1 || x0_bytes + width_bytes > pitch_bytes)
    return CHIPS_E_ARG;
  const size_t off = (size_t)y0 * pitch_bytes + x0_bytes;
  CHIPS_CUDA(cudaMemcpy2DAsync((char*)dst + off, pitch_bytes, (const char*)src + off, pitch_bytes, width_bytes,
                               (size_t)rows, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                               (cudaStream_t)stream));
  return CHIPS_OK;
}
