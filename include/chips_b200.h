/* chips_b200.h — C ABI of the B200-native CHIPS detection hot path.
 *
 * The reference (shibaji7/pyCHIPS) has no FFI layer: its hot path is Python calling
 * numpy / scipy.signal / OpenCV directly.  Each entry point below
 * replaces one of those call sites; the citation after "replaces:" is the reference
 * line(s) a maintainer would rebind (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it is named h_*;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - all calls only ENQUEUE work on `stream` and never synchronise;
 *   - return value: 0 = OK, negative = CHIPS_E_* ; nothing throws across the ABI;
 *   - images are row-major [H][W]; `elem` is 4 (float32) or 8 (float64) bytes;
 *   - disk geometry: pixel (x,y) is on-disk iff (x-cx)^2+(y-cy)^2 <= r^2 in integers,
 *     which is exactly filled cv2.circle (chips.py:186-192); r < 0 means "no mask"
 *     (the synoptic variant, syn_chips.py).
 */
#ifndef CHIPS_B200_H
#define CHIPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHIPS_OK 0
#define CHIPS_E_ARG (-1)      /* bad argument (k even, k too large, T too large, ...) */
#define CHIPS_E_CUDA (-2)     /* a CUDA runtime call / launch failed                  */
#define CHIPS_E_WS (-3)       /* workspace too small                                  */

#define CHIPS_MAX_T 64        /* max number of thresholds per detection               */
#define CHIPS_MAX_K 73        /* max median kernel size (block sort: a 16-row block with its halo fits shared memory) */
/* pixels of one halo-extended median block that the block sort holds in shared memory
 * (keys + two index arrays): float32 / float64 storage                                  */
#define CHIPS_MAX_BLOCK_ELEMS_F32 24576
#define CHIPS_MAX_BLOCK_ELEMS_F64 17600
#define CHIPS_PROB_BINS 20    /* chips.py:445 np.linspace(0,1,21)                     */

/* Per-image result record written by the device epilogues (one per detection). */
typedef struct chips_result {
  double first_edge;          /* histogram range = min / max of the filtered on-disk data */
  double last_edge;
  double h_thresh;            /* chips.py:240-241 (sticky: in = NaN means "compute")      */
  double limit_0;             /* chips.py:356 peaks[0]                                    */
  double limits[CHIPS_MAX_T]; /* chips.py:357 np.round(limit_0 + arange, 4)               */
  double probs[CHIPS_MAX_T];  /* chips.py:378 / :432-450                                  */
  int64_t n_samples;          /* number of histogrammed pixels                            */
  int32_t n_peaks;            /* len(find_peaks(...)) ; 0 => no regions (chips.py:355)    */
  int32_t n_thresholds;       /* T                                                        */
  int32_t n_kept[CHIPS_MAX_T];/* kept top-level components per threshold                  */
  int32_t n_comp[CHIPS_MAX_T];/* top-level components per threshold before the area cut   */
  int32_t median_errors;      /* internal consistency counter of the median (must be 0)  */
  int32_t n_nonfinite;        /* NaN / +-Inf pixels met by the median's input pass (the reference's
                               * scipy.signal.medfilt2d result is undefined for them): the host
                               * classes raise when it is not 0                               */
} chips_result;

/* Byte offsets of every buffer inside one workspace blob (all 256-byte aligned). */
typedef struct chips_plan {
  int32_t H, W, k, h_bins, T, elem, masked;
  int32_t f32_math;   /* 1: the caller's arrays are float32 and numpy would keep float32 for the bin edges,
                       * bound / limit_0 and the sigmoid of calculate_prob (np.histogram's bin_type,
                       * chips.py:235-244, :446); 0 (chips_plan_init): float64 arithmetic.  elem must be 4. */
  int32_t roi_x0, roi_y0, roi_w, roi_h;   /* bounding box of the disk (whole image if unmasked) */
  /* median blocks: the ROI is cut into blk_nbx x blk_nby blocks of 128 columns x blk_rows rows
   * (blk_rows a multiple of 16).  Each block sorts its halo-extended region ((128+k-1) x bp_rows
   * pixels, zero padding outside the image) and keeps the pixels' positions in key order (u16
   * lists) and every pixel's bucket = rank / blk_bucket (u8, < 128; images of bp_rows x bp_pitch) */
  int32_t bp_pitch, bp_rows;
  int32_t blk_rows, blk_nbx, blk_nby, blk_bucket;
  size_t off_order;      /* u16 [blocks][align8(pixels per block)]  the block's pixels in key order, (ly << 8 | lx) */
  size_t off_bucket;     /* u8  [blocks][bp_rows][bp_pitch]   block-local bucket images   */
  size_t off_bstar;      /* u8  [H][W]   median bucket per pixel                          */
  size_t off_rres;       /* u32 [H][W]   residual rank | (bucket population << 16)        */
  size_t off_filt;       /* elem[H][W]   masked median (filt_disk)                        */
  size_t off_minmax;     /* 2 ordered keys (8 B each)                                     */
  size_t off_counts;     /* i64 [h_bins]                                                  */
  size_t off_density;    /* f64 [h_bins]   hbase                                          */
  size_t off_edges;      /* f64 [h_bins+1] hbins                                          */
  size_t off_bound;      /* f64 [h_bins]   bound                                          */
  size_t off_peaks;      /* i32 [h_bins]   peak indices (first n_peaks valid)             */
  size_t off_result;     /* chips_result                                                  */
  size_t off_level;      /* u8  [H][W]   first threshold index with v<=lim (T = none)     */
  size_t off_fill;       /* u8  [H][W]   hole-filled level W                              */
  size_t off_keep;       /* u8  [H][W]   first threshold at which the pixel is in the final map */
  size_t off_probtab;    /* u32 [T+1][20]                                                 */
  size_t off_parent_b;   /* u32 [roi_h*roi_w + 1]   background union-find                 */
  size_t off_parent_c;   /* u32 [2][roi_h*roi_w]    component tree: union-find, merge parents */
  size_t off_area;       /* u32 [roi] 2*area at roots; u8 [3][roi] hook level, kept level, W  */
  size_t off_pending;    /* u32 [roi_h*roi_w]  hole-fill work list grouped by level       */
  size_t off_list2;      /* u32 [roi_h*roi_w]  unsorted list (phase B) / foreground list grouped by W (phase C) */
  size_t off_lvl;        /* u32 [512]          per-level counts / offsets / cursors, phase-C counters */
  size_t off_tiles;      /* (reserved)                                                    */
  size_t off_ovf;        /* (reserved)                                                    */
  size_t bytes;          /* total workspace size                                          */
} chips_plan;

int chips_version(void);
/* kernels launched by this library so far (process-wide counter) */
long long chips_launch_count(void);

/* Fill `plan` for one image geometry.  cx,cy,r as in the conventions (r<0: unmasked). */
int chips_plan_init(chips_plan* plan, int H, int W, int k, int h_bins, int T, int elem,
                    int cx, int cy, int r);

/* ---- row C  replaces: cv2.circle canvas + NaN/1 assigns, chips.py:183-200 ---------- */
int chips_disk_mask(void* n_mask, void* i_mask, int H, int W, int elem, int cx, int cy,
                    int r, void* stream);

/* ---- row D  replaces: i_mask * scipy.signal.medfilt2d(normalized, k), chips.py:214-216;
 *             scipy.signal.medfilt2d(raw, k), syn_chips.py:106.
 * Zero-padded k x k median (rank (k*k-1)/2), bit-exact; with r >= 0 the output is the
 * median on-disk and 0 off-disk, and min/max keys of the on-disk output are left at
 * plan->off_minmax.  `out` may be ws + plan->off_filt.
 * Fast path of the float32 block sort: `in` 16-byte aligned and W a multiple of 4 (any cudaMalloc'd
 * image of an even-sized instrument frame) - the kernel then stages each block with one TMA tensor
 * tile load; anything else takes plain loads, same result (CHIPS_NO_TMA=1 forces them: A/B timing). */
int chips_medfilt2d(const void* in, void* out, const chips_plan* plan, int cx, int cy, int r,
                    void* ws, void* stream);
/* the same, restricted to a subset of its kernels (bit 0 block sort, 2 tier-1 sliding
 * histograms, 3 tier-2 rank select; bit 1 unused) so a caller can time each one */
int chips_medfilt2d_stages(const void* in, void* out, const chips_plan* plan, int cx, int cy,
                           int r, void* ws, void* stream, int stages);
/* brute-force radix-select reference of the same filter (tests / at-scale checker) */
int chips_medfilt2d_bruteforce(const void* in, void* out, int H, int W, int k, int elem,
                               int cx, int cy, int r, void* stream);

/* ---- row E  replaces: np.histogram(h_data[~isnan], bins=h_bins), chips.py:235-239 --- */
int chips_minmax(const void* filt, const chips_plan* plan, int cx, int cy, int r, void* ws,
                 void* stream);
int chips_histogram(const void* filt, const chips_plan* plan, int cx, int cy, int r, void* ws,
                    void* stream);
/* replaces: density, h_thresh, scipy.signal.find_peaks, bound, peaks, limits
 *           chips.py:239-244, :356-357.  h_thresh_in: NaN = compute max(h)/ratio.
 * offsets: plan->T HOST doubles = np.arange(threshold_range[0], threshold_range[1]) (chips.py:352;
 *           List[float] in the reference: any start, unit steps), read before the call returns. */
int chips_select_threshold(const chips_plan* plan, double ht_peak_ratio, double h_thresh_in,
                           const double* offsets, void* ws, void* stream);

/* ---- rows F+G  replaces: the `for lim in limits` mask assigns chips.py:360-375 and
 *                calculate_prob chips.py:378,432-450 (all T in one pass) ------------- */
int chips_threshold_prob(const void* filt, const chips_plan* plan, int cx, int cy, int r,
                         double porb_threshold, void* ws, void* stream);

/* ---- row H  replaces: cv2.findContours + clean_small_scale_structures
 *             chips.py:382-389, :405-430 (hole fill + 8-conn CCL + bit-quad area) ---- */
int chips_cleanup(const chips_plan* plan, int cx, int cy, int r, double disk_area,
                  double area_threshold, void* ws, void* stream);

/* Expand the level-encoded result into T uint8 maps [T][H][W] (map_t = keep <= t), or
 * the raw threshold masks when which = 1 (level <= t), or filled masks which = 2. */
int chips_expand_maps(const chips_plan* plan, int which, uint8_t* maps, void* ws, void* stream);
/* One float64/float32 map (the reference's `dxmap/255`, chips.py:430) for threshold t. */
int chips_expand_map(const chips_plan* plan, int which, int t, void* out, int out_elem,
                     void* ws, void* stream);
/* Root image of the per-threshold labelling: roots[y][x] = ROI-local index of the first
 * pixel (raster order) of the pixel's filled top-level component at threshold t, -1 for
 * background; twice_area (optional) = 2*contourArea of that component.  Numbering the
 * distinct roots in ascending order gives the canonical label image
 * (== scipy.ndimage.label(filled, ones((3,3)))). */
int chips_roots(const chips_plan* plan, int cx, int cy, int r, int t, int32_t* roots,
                uint32_t* twice_area, void* ws, void* stream);

/* ---- row J  replaces: probability-map stack plots.py:305-322 ---------------------- */
int chips_prob_stack(const chips_plan* plan, double lower_lim, void* out, int out_elem,
                     int accumulate, void* ws, void* stream);

/* ---- rows B..I fused: one full detection enqueued on `stream` --------------------- */
int chips_detect(const void* image, const chips_plan* plan, int cx, int cy, int r,
                 double ht_peak_ratio, double h_thresh_in, const double* offsets,
                 double porb_threshold, double area_threshold, uint8_t* maps_or_null,
                 void* ws, void* stream);

/* ---- SURVEY.md §8(f) row 3 / sliding windows.  replaces: the loop body of Chips.extract_histograms,
 *      chips.py:282-300, for nwin windows at once (extract_sliding_histograms, chips.py:321-339, is
 *      an empty stub in the reference: PARITY UNPINNED, pinned to oracle/chips_oracle.py).
 *      rects: HOST [nwin][4] = x0, y0, x1, y1 (half-open) inside the image; per window the
 *      histogram of its on-disk pixels over its own min/max range, density, bound = be[:-1] +
 *      diff(be), find_peaks(height = h_thresh) indices (first n_peaks of row w of `peaks`).
 *      h_thresh_in = NaN: max(density of window 0) / ht_peak_ratio for every window (sticky).
 *      counts/density/bound/peaks: device [nwin][h_bins]; recs: device [nwin]. ------------- */
typedef struct chips_window_rec {
  double first_edge, last_edge, h_thresh;
  int64_t n_samples;
  int32_t n_peaks, pad_;
} chips_window_rec;
size_t chips_window_workspace_bytes(int nwin);
int chips_window_histograms(const void* filt, int H, int W, int elem, int f32_math, int cx, int cy, int r,
                            int nwin, const int32_t* rects, int h_bins, double ht_peak_ratio, double h_thresh_in,
                            long long* counts, double* density, double* bound, int32_t* peaks,
                            chips_window_rec* recs, void* ws, void* stream);

/* ---- BASELINE config 4 "wrap-around CCL": 8-connected components of a threshold mask whose
 *      longitude axis is periodic (column W-1 is adjacent to column 0).  The reference's
 *      SynopticChips has no labelling step (syn_chips.py:237-251): an extension, PARITY UNPINNED,
 *      pinned to scipy.ndimage.label + a seam merge in the tests.  Foreground: img <= thr (mode 0,
 *      e.g. the level image with thr = threshold index) or img != 0 (mode 1).  labels: int32 [H][W],
 *      0 = background, components numbered 1.. in raster order of their first pixel; areas
 *      (optional, max_labels ints) = pixel counts; *n_labels = number of components. ------------ */
size_t chips_label_workspace_bytes(int H, int W);
int chips_label_periodic(const uint8_t* img, int H, int W, int mode, int thr, int periodic,
                         int32_t* labels, int32_t* areas, long long max_labels, int32_t* n_labels,
                         void* ws, void* stream);

/* ---- batched detections (SURVEY.md §8(b) "batched variants", §8(e) time series).  The reference has
 *      no batching: one Chips object handles one RegisterAIA (chips.py:110-150) and the caller
 *      loops over dates; this is that loop for a non-Python host.  The handle owns n_slots
 *      in-flight detections (stream, device image, workspace, one instantiated CUDA graph per
 *      geometry).  images[i]: H x W x elem bytes, host (pinned for asynchronous copies) or device;
 *      only chips_input_rect of each is read.  results[i]: the record.  keep_or_null[i] (may be
 *      NULL per image): H x W bytes, the level-encoded final map (map_t = keep <= t).
 *      prob_sum_or_null: H x W floats, sum over the batch of the plots.py:305-322 probability
 *      maps (prob_lower_lim as in chips_prob_stack), accumulated in image order.
 *      Returns when every output is on the host.  masked handles take r[i] >= 0, others r[i] < 0. */
int chips_batch_create(void** handle, int H, int W, int k, int h_bins, int T, int elem, int masked,
                       int n_slots);
int chips_detect_batch(void* handle, int B, const void* const* images, int images_on_device,
                       const int32_t* cx, const int32_t* cy, const int32_t* r,
                       double ht_peak_ratio, double h_thresh_in, const double* offsets,
                       double porb_threshold, double area_threshold, chips_result* results,
                       uint8_t* const* keep_or_null, float* prob_sum_or_null, double prob_lower_lim);
int chips_batch_destroy(void* handle);

/* ---- SURVEY.md §8(f) row 1  replaces: CleanFl.create_coronal_hole_candidates,
 *      cleanup.py:69-200 (three co-registered wavelengths of the same size; the reference's
 *      skimage resize is the identity then), and the mask hook chips.py:376-377. ---------- */
typedef struct chips_fl_params {   /* cleanup.py:52-60 */
  double clip_negative;
  double clip_211_log[2], clip_193_log[2], clip_171_log[2];
  double bmmix_value, bmhot_value, bmcool_value;
} chips_fl_params;

typedef struct chips_fl_stats {    /* written by the device */
  double mean171, mean193, mean211;       /* np.mean of the clipped images, cleanup.py:158-178 */
  double thr_mix, thr_hot, thr_cool;      /* right-hand sides of the three comparisons          */
  int64_t n_mix, n_hot, n_cool, n_cand;   /* on-disk pixels set in each stored bitmap           */
  int64_t n_near;                         /* pixels within one float32 ulp of a threshold       */
} chips_fl_stats;

#define CHIPS_FL_MIX 1    /* bmmix  (before the disk mask)          */
#define CHIPS_FL_HOT 2    /* bmhot                                   */
#define CHIPS_FL_COOL 4   /* bmcool                                  */
#define CHIPS_FL_CAND 8   /* bmcool * bmmix * bmhot                  */
#define CHIPS_FL_DISK 16  /* filled cv2.circle disk mask             */

size_t chips_fl_workspace_bytes(int H, int W);
/* bits[H][W]: OR of CHIPS_FL_* per pixel; stats, ws: device memory */
int chips_fl_candidates(const void* d171, const void* d193, const void* d211, int H, int W,
                        int elem, int cx, int cy, int r, const chips_fl_params* h_params,
                        uint8_t* bits, chips_fl_stats* stats, void* ws, void* stream);
/* level[i] = T (in no mask) wherever bits[i] lacks CHIPS_FL_CAND; call between
 * chips_threshold_prob and chips_cleanup */
int chips_apply_candidates(const chips_plan* plan, const uint8_t* bits, void* ws, void* stream);

/* ---- SURVEY.md §8(f) row 2  replaces: the `data[:, :, i] = region.map` cube of Chips.to_netcdf,
 *      chips.py:570-586.  cube is float32 [H][W][T] (T innermost), 1.0 / 0.0; `which` as in
 *      chips_expand_maps. ------------------------------------------------------------------ */
int chips_pack_cube(const chips_plan* plan, int which, float* cube, void* ws, void* stream);

/* ---- SURVEY.md §8(f) row 4  replaces: Chips.compute_similarity_measures, chips.py:592-617.
 *      map, map0: [H][W] of `elem` bytes; out: double[H + 2] = per-row cosine similarity (NaN for
 *      rows with sum(map[i]) <= 0), then np.nanmean of them and the number of rows used. ---- */
int chips_row_cosine(const void* map, const void* map0, int H, int W, int elem, double* out,
                     void* stream);

/* ---- SURVEY.md §8(f) row 4  replaces: cv2.findContours(tmp_data.astype(np.uint8), cv2.RETR_TREE,
 *      cv2.CHAIN_APPROX_SIMPLE), chips.py:382-386 — point lists, order of discovery and parents of
 *      every border (Suzuki-Abe border following as data-parallel passes, csrc/contours_core.h).
 *      Foreground of `img` ([h] rows of `pitch` bytes): img <= thr (mode 0, e.g. the level image at
 *      plan->off_level with thr = threshold index) or img != 0 (mode 1).
 *      recs[b], b = 0 .. n_contours-1 in ORDER OF DISCOVERY (border number b + 2): start pixel,
 *      hole flag, parent border number (1 = the image frame), number of CHAIN_APPROX_SIMPLE points,
 *      offset of its first point in `points` (x, y pairs), and the shoelace sum of the border
 *      (|area2| == 2 * cv2.contourArea).  OpenCV's output order and hierarchy rows follow from the
 *      parents on the host (pychips_b200/contours.py::output_order).
 *      counts->overflow: bit 0 more than max_border_px border pixels (nothing else is valid then;
 *      n_border_px holds the number needed), bit 1 n_contours > max_contours, bit 2 n_points >
 *      max_points (records / points beyond the capacity are not written).  counts->changed != 0:
 *      the start positions had not settled after `start_rounds` rounds — call again with more. */
typedef struct chips_contour_rec {
  int32_t x, y, is_hole, parent, n_points, offset;
  int64_t area2;
} chips_contour_rec;

typedef struct chips_contour_counts {
  int32_t n_border_px, n_contours, n_points, changed, overflow, pad_;
} chips_contour_counts;

size_t chips_contours_workspace_bytes(int w, int h, long long max_border_px);
int chips_contours(const uint8_t* img, int w, int h, int pitch, int mode, int thr,
                   long long max_border_px, int start_rounds, chips_contour_rec* recs,
                   long long max_contours, int32_t* points, long long max_points,
                   chips_contour_counts* counts, void* ws, size_t ws_bytes, void* stream);

/* ---- uploads of only what the path reads.  rect = {x0, y0, x1, y1}: the pixels of the input image
 *      any kernel of a detection with this plan can touch (bounding box of the disk, extended to the
 *      median's blocks and halo, clipped to the image); everything outside may stay uninitialised.
 *      chips_copy_rect: cudaMemcpy2DAsync of rows [y0, y0+rows) x bytes [x0_bytes, x0_bytes+width_bytes)
 *      between two images of the same pitch (to_device: host -> device, else device -> host). */
int chips_input_rect(const chips_plan* plan, int32_t* rect);
/* tighter: disjoint row intervals, spans[i] = {y0, y1, x0, x1} (half-open), ascending in y - the
 * halo-extended regions of the median blocks that hold an on-disk pixel (52 % of a 4096^2 image at
 * r = 1577, k = 51; the rectangle is 63 %).  Returns the number of spans or a negative error
 * (CHIPS_E_WS: more than max_spans).  One chips_copy_rect per span uploads an image. */
int chips_input_spans(const chips_plan* plan, int cx, int cy, int r, int32_t* spans, int max_spans);
int chips_copy_rect(void* dst, const void* src, size_t pitch_bytes, size_t x0_bytes, int y0,
                    size_t width_bytes, int rows, int to_device, void* stream);

/* ---- downloads of only what carries information.  The level-encoded result map (plan->off_keep)
 *      is T ("in no map") almost everywhere: chips_pack_keep appends one word per pixel with
 *      keep < T, (y * W + x) << 6 | keep, in arbitrary order, to `entries` (device memory or
 *      mapped pinned host memory, `cap` words) and leaves their number in *count (device; entries
 *      beyond cap are counted, not written).  When *count > cap and cap * 4 >= H * W (entries
 *      16-byte aligned), the buffer receives the dense H * W bytes of the map instead: the reader
 *      tells the two apart by *count > cap.  Needs H * W <= 2^26, T <= 64, W % 4 == 0. -------- */
int chips_pack_keep(const chips_plan* plan, uint32_t* entries, uint32_t cap, uint32_t* count, void* ws,
                    void* stream);
/* copy of min(*count, cap) 32-bit words whose number only the device knows (count: device memory),
 * e.g. the entries of chips_pack_keep into mapped pinned host memory, by 8 CTAs on `stream` (a copy
 * stream: it takes no SMs worth mentioning from the detections); *count_out (optional, may be
 * mapped host memory) = *count.  src, dst 16-byte aligned. */
int chips_copy_words(const uint32_t* src, uint32_t* dst, const uint32_t* count, uint32_t cap,
                     uint32_t* count_out, void* stream);

/* helper: f64 -> f32 demotion with an exactness flag (1 = every value representable) */
int chips_demote_f64(const double* in, float* out, size_t n, int32_t* all_exact, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHIPS_B200_H */
