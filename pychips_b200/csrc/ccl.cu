// ccl.cu — replaces cv2.findContours(RETR_TREE) + cv2.contourArea + cv2.drawContours(fill)
// of /root/reference/chips/chips.py:382-389 and :405-430 for ALL T thresholds at once.
//
// Formulation (executable spec: oracle/first_principles.py):
//   level(p)  = first threshold index whose mask contains p  (masks are nested)
//   W(p)      = greyscale hole fill of `level`: max over 4-connected paths from p to the
//               outside of the disk/image of the min level on the path.  {W <= t} is mask_t
//               with every hole (and whatever is nested inside) filled == the union of the
//               filled top-level contours of mask_t.
//   per t     : 8-connected components of {W <= t}; 2*contourArea = 2*Q4 + Q3 (bit quads);
//               keep if (2A/2)/(pi r^2) >= area_threshold.
//   keep(p)   = first t at which p's component is kept (areas only grow with t) — the
//               level-encoded stack of the T final maps: map_t = (keep <= t).
//
// Phase B (W): ONE union-find over the disk, thresholds visited in DEscending order so the
// outer background only ever grows; a pixel's W is the round in which it first joins the
// border set.  Phase C: the T foreground labellings are independent -> gridDim.z = T.
// Union-find = min-index roots, atomicMin hooking (Komura/Playne), path halving in find.
#include "common.cuh"

namespace chips {

constexpr uint32_t kBorder = 0;   // phase-B node 0 = "outside"; pixel li is node li+1

__device__ __forceinline__ uint32_t uf_find(uint32_t* __restrict__ parent, uint32_t i) {
  uint32_t p = parent[i];
  if (p == i) return i;
  uint32_t prev = i;
  uint32_t gp = parent[p];
  while (p != gp) {
    parent[prev] = gp;          // path halving; any ancestor is a valid parent
    prev = p;
    p = gp;
    gp = parent[p];
  }
  return p;
}

__device__ __forceinline__ void uf_unite(uint32_t* __restrict__ parent, uint32_t a, uint32_t b) {
  a = uf_find(parent, a);
  b = uf_find(parent, b);
  while (a != b) {
    if (a < b) {
      uint32_t t = a;
      a = b;
      b = t;
    }                            // a > b : hook a under b
    uint32_t old = atomicMin(&parent[a], b);
    a = (old == a) ? b : old;    // a was not a root any more: continue with its old parent
  }
}

struct Roi {
  int x0, y0, w, h, W, H, cx, cy, r;
  __device__ __forceinline__ bool inside(int x, int y) const {
    return x >= x0 && x < x0 + w && y >= y0 && y < y0 + h && on_disk(x, y, cx, cy, r);
  }
  __device__ __forceinline__ uint32_t li(int x, int y) const {
    return (uint32_t)((y - y0) * w + (x - x0));
  }
};

// ------------------------------------------------------------------ phase B
__global__ void __launch_bounds__(256) k_fill_init(Roi g, uint8_t* __restrict__ Wimg,
                                                   uint32_t* __restrict__ parent) {
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) parent[0] = kBorder;
  if (x >= g.x0 + g.w) return;
  uint32_t n = g.li(x, y) + 1;
  parent[n] = n;
  if (on_disk(x, y, g.cx, g.cy, g.r)) Wimg[(size_t)y * g.W + x] = 0;
}

// Row-run linking for the pixels of one level set (`sel`): every pixel points at the start
// of its run inside the warp's 32-pixel span, a run that continues across the span boundary
// points at the last pixel of the previous span (chains hop 32 pixels at a time).
template <typename Sel>
__device__ __forceinline__ void link_row_runs(const Roi& g, uint32_t* __restrict__ parent,
                                              uint32_t node_off, Sel sel) {
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  int lane = threadIdx.x & 31;
  bool act = (x < g.x0 + g.w) && sel(x, y);
  unsigned m = __ballot_sync(0xFFFFFFFFu, act);
  bool left_act = false;
  if (lane == 0 && act) left_act = (x - 1 >= g.x0) && sel(x - 1, y);
  left_act = __shfl_sync(0xFFFFFFFFu, left_act, 0);
  if (!act) return;
  unsigned z = ~m & ((1u << lane) - 1u);
  int s = z ? (32 - __clz(z)) : 0;
  uint32_t me = g.li(x, y) + node_off;
  uint32_t tgt;
  if (s == 0 && left_act) tgt = me - (uint32_t)lane - 1u;     // last pixel of previous span
  else tgt = me - (uint32_t)(lane - s);
  parent[me] = tgt;
}

__global__ void __launch_bounds__(256) k_fill_bulk_link(Roi g, const uint8_t* __restrict__ lev, int T,
                                                        uint32_t* __restrict__ parent) {
  link_row_runs(g, parent, 1u, [&](int x, int y) {
    return on_disk(x, y, g.cx, g.cy, g.r) && lev[(size_t)y * g.W + x] == T;
  });
}

// unions of the pixels that enter the background at round t (level == t+1) with their
// already-active 4-neighbours (level >= t+1) and with the border node.
template <bool BULK>
__global__ void __launch_bounds__(256) k_fill_union(Roi g, const uint8_t* __restrict__ lev, int t,
                                                    uint32_t* __restrict__ parent) {
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (x >= g.x0 + g.w) return;
  const int lv = t + 1;
  if (lev[(size_t)y * g.W + x] != lv || !on_disk(x, y, g.cx, g.cy, g.r)) return;
  uint32_t me = g.li(x, y) + 1;
  auto level_at = [&](int xx, int yy) -> int {   // -1 = outside (border)
    if (!g.inside(xx, yy)) return -1;
    return lev[(size_t)yy * g.W + xx];
  };
  int l_l = level_at(x - 1, y), l_r = level_at(x + 1, y);
  int l_u = level_at(x, y - 1), l_d = level_at(x, y + 1);
  if (l_l < 0 || l_r < 0 || l_u < 0 || l_d < 0) uf_unite(parent, me, kBorder);
  if (BULK) {
    // left edge done by the run links; vertical edge unless implied by left & up-left
    if (l_u == lv) {
      bool implied = (l_l == lv) && (level_at(x - 1, y - 1) == lv);
      if (!implied) uf_unite(parent, me, me - (uint32_t)g.w);
    }
  } else {
    if (l_l >= lv) uf_unite(parent, me, me - 1u);
    if (l_u >= lv) uf_unite(parent, me, me - (uint32_t)g.w);
    if (l_r > lv) uf_unite(parent, me, me + 1u);
    if (l_d > lv) uf_unite(parent, me, me + (uint32_t)g.w);
  }
}

__global__ void __launch_bounds__(256) k_fill_assign(Roi g, const uint8_t* __restrict__ lev, int t,
                                                     uint8_t* __restrict__ Wimg,
                                                     uint32_t* __restrict__ parent) {
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (x >= g.x0 + g.w) return;
  size_t o = (size_t)y * g.W + x;
  if (lev[o] <= t || Wimg[o] != 0 || !on_disk(x, y, g.cx, g.cy, g.r)) return;
  if (uf_find(parent, g.li(x, y) + 1) == kBorder) Wimg[o] = (uint8_t)(t + 1);
}

// ------------------------------------------------------------------ phase C
__global__ void __launch_bounds__(256) k_ccl_link(Roi g, const uint8_t* __restrict__ Wimg,
                                                  uint32_t* __restrict__ parent_all,
                                                  uint32_t* __restrict__ area_all) {
  const int t = blockIdx.z;
  const size_t n = (size_t)g.w * g.h;
  uint32_t* parent = parent_all + (size_t)t * n;
  uint32_t* area = area_all + (size_t)t * n;
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (x < g.x0 + g.w && Wimg[(size_t)y * g.W + x] <= t) area[g.li(x, y)] = 0;
  link_row_runs(g, parent, 0u, [&](int xx, int yy) { return Wimg[(size_t)yy * g.W + xx] <= t; });
}

__global__ void __launch_bounds__(256) k_ccl_union(Roi g, const uint8_t* __restrict__ Wimg,
                                                   uint32_t* __restrict__ parent_all) {
  const int t = blockIdx.z;
  uint32_t* parent = parent_all + (size_t)t * g.w * g.h;
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (x >= g.x0 + g.w || y == g.y0) return;
  auto fg = [&](int xx, int yy) -> bool {
    return xx >= g.x0 && xx < g.x0 + g.w && yy >= g.y0 && yy < g.y0 + g.h &&
           Wimg[(size_t)yy * g.W + xx] <= t;
  };
  if (!fg(x, y)) return;
  uint32_t me = g.li(x, y);
  bool up = fg(x, y - 1), ul = fg(x - 1, y - 1), ur = fg(x + 1, y - 1);
  bool lf = fg(x - 1, y), rt = fg(x + 1, y);
  if (up) {
    if (!(lf && ul)) uf_unite(parent, me, me - (uint32_t)g.w);
  } else {
    if (ul && !lf) uf_unite(parent, me, me - (uint32_t)g.w - 1u);
    if (ur && !rt) uf_unite(parent, me, me - (uint32_t)g.w + 1u);
  }
}

// 2*area = 2*Q4 + Q3 over every 2x2 quad of the zero-padded filled set
__global__ void __launch_bounds__(256) k_ccl_area(Roi g, const uint8_t* __restrict__ Wimg,
                                                  uint32_t* __restrict__ parent_all,
                                                  uint32_t* __restrict__ area_all) {
  const int t = blockIdx.z;
  const size_t n = (size_t)g.w * g.h;
  uint32_t* parent = parent_all + (size_t)t * n;
  uint32_t* area = area_all + (size_t)t * n;
  int qx = g.x0 - 1 + blockIdx.x * blockDim.x + threadIdx.x;   // quad = (qx..qx+1, qy..qy+1)
  int qy = g.y0 - 1 + blockIdx.y;
  auto fg = [&](int xx, int yy) -> bool {
    return xx >= g.x0 && xx < g.x0 + g.w && yy >= g.y0 && yy < g.y0 + g.h &&
           Wimg[(size_t)yy * g.W + xx] <= t;
  };
  uint32_t root = 0xFFFFFFFFu, add = 0;
  if (qx < g.x0 + g.w) {
    bool a = fg(qx, qy), b = fg(qx + 1, qy), c = fg(qx, qy + 1), d = fg(qx + 1, qy + 1);
    int s = (int)a + (int)b + (int)c + (int)d;
    if (s >= 3) {
      int px = a ? qx : qx + 1, py = qy;   // a, else b: one of the top pair is set when s >= 3
      root = uf_find(parent, g.li(px, py));
      add = (s == 4) ? 2u : 1u;
    }
  }
  // warp-aggregate per root before the global atomic
  unsigned active = __ballot_sync(0xFFFFFFFFu, add != 0);
  if (add) {
    unsigned peers = __match_any_sync(active, root);
    uint32_t sum = 0;
    for (unsigned mm = peers; mm; mm &= mm - 1) sum += __shfl_sync(peers, add, __ffs(mm) - 1);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&area[root], sum);
  }
}

__device__ __forceinline__ bool area_kept(uint32_t a2, double disk_area, double thr) {
  // chips.py:424-426: cv2.contourArea(cnt) / self.disk_area >= self.area_threshold
  return __ddiv_rn(__dmul_rn((double)a2, 0.5), disk_area) >= thr;
}

__global__ void __launch_bounds__(256) k_ccl_stats(Roi g, const uint8_t* __restrict__ Wimg,
                                                   const uint32_t* __restrict__ parent_all,
                                                   const uint32_t* __restrict__ area_all,
                                                   double disk_area, double thr,
                                                   chips_result* __restrict__ res) {
  const int t = blockIdx.z;
  const size_t n = (size_t)g.w * g.h;
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  bool root = false, kept = false;
  if (x < g.x0 + g.w && Wimg[(size_t)y * g.W + x] <= t) {
    uint32_t me = g.li(x, y);
    if (parent_all[(size_t)t * n + me] == me) {
      root = true;
      kept = area_kept(area_all[(size_t)t * n + me], disk_area, thr);
    }
  }
  unsigned mr = __ballot_sync(0xFFFFFFFFu, root), mk = __ballot_sync(0xFFFFFFFFu, kept);
  if ((threadIdx.x & 31) == 0) {
    if (mr) atomicAdd(&res->n_comp[t], __popc(mr));
    if (mk) atomicAdd(&res->n_kept[t], __popc(mk));
  }
}

__global__ void __launch_bounds__(256) k_ccl_keep(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                  uint32_t* __restrict__ parent_all,
                                                  const uint32_t* __restrict__ area_all,
                                                  double disk_area, double thr,
                                                  uint8_t* __restrict__ keep) {
  const size_t n = (size_t)g.w * g.h;
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (x >= g.x0 + g.w) return;
  size_t o = (size_t)y * g.W + x;
  int w = Wimg[o];
  int kk = T;
  uint32_t me = g.li(x, y);
  for (int t = w; t < T; ++t) {
    uint32_t root = uf_find(parent_all + (size_t)t * n, me);
    if (area_kept(area_all[(size_t)t * n + root], disk_area, thr)) {
      kk = t;
      break;
    }
  }
  keep[o] = (uint8_t)kk;
}

__global__ void k_clear_counts(chips_result* res) {
  int t = threadIdx.x;
  if (t < CHIPS_MAX_T) {
    res->n_kept[t] = 0;
    res->n_comp[t] = 0;
  }
}

// ------------------------------------------------------------------ expansion / debugging
__global__ void __launch_bounds__(256) k_expand_all(const uint8_t* __restrict__ src, size_t n, int T,
                                                    uint8_t* __restrict__ maps) {
  // 4 pixels per thread, one 32-bit store per plane
  size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n && (n & 3) == 0) {      // plane stride n keeps the 32-bit stores aligned
    uint32_t v = *reinterpret_cast<const uint32_t*>(src + i4);
    for (int t = 0; t < T; ++t) {
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) o |= ((((v >> (8 * b)) & 255u) <= (uint32_t)t) ? 1u : 0u) << (8 * b);
      *reinterpret_cast<uint32_t*>(maps + (size_t)t * n + i4) = o;
    }
  } else {
    for (size_t i = i4; i < n && i < i4 + 4; ++i)
      for (int t = 0; t < T; ++t) maps[(size_t)t * n + i] = src[i] <= t;
  }
}

template <typename O>
__global__ void __launch_bounds__(256) k_expand_one(const uint8_t* __restrict__ src, size_t n, int t,
                                                    O* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (src[i] <= t) ? (O)1 : (O)0;
}

__global__ void __launch_bounds__(256) k_roots(Roi g, const uint8_t* __restrict__ Wimg, int t,
                                               uint32_t* __restrict__ parent_all,
                                               const uint32_t* __restrict__ area_all,
                                               int32_t* __restrict__ roots,
                                               uint32_t* __restrict__ area_out) {
  const size_t n = (size_t)g.w * g.h;
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= g.W) return;
  size_t o = (size_t)y * g.W + x;
  int32_t rr = -1;
  uint32_t aa = 0;
  if (x >= g.x0 && x < g.x0 + g.w && y >= g.y0 && y < g.y0 + g.h && Wimg[o] <= t) {
    uint32_t root = uf_find(parent_all + (size_t)t * n, g.li(x, y));
    rr = (int32_t)root;
    aa = area_all[(size_t)t * n + root];
  }
  roots[o] = rr;
  if (area_out) area_out[o] = aa;
}

template <typename O>
__global__ void __launch_bounds__(256) k_prob_stack(const uint8_t* __restrict__ keep, size_t n,
                                                    const chips_result* __restrict__ res,
                                                    double lower_lim, int accumulate,
                                                    O* __restrict__ out) {
  // plots.py:305-322: n_probs = (p - min)/(max - min); stacked = max_t(map_t * n_prob_t)
  __shared__ double suffix[CHIPS_MAX_T + 1];
  __shared__ int poisoned;
  const int T = res->n_thresholds;
  if (threadIdx.x == 0) {
    double pmin = INFINITY, pmax = -INFINITY;
    bool bad = res->n_peaks <= 0;
    for (int t = 0; t < T; ++t) {
      double p = res->probs[t];
      if (isnan(p)) bad = true;
      pmin = fmin(pmin, p);
      pmax = fmax(pmax, p);
    }
    double run = 0.0;
    suffix[T] = 0.0;
    for (int t = T - 1; t >= 0; --t) {
      double np_ = __ddiv_rn(__dsub_rn(res->probs[t], pmin), __dsub_rn(pmax, pmin));
      if (isnan(np_)) bad = true;
      if (np_ >= lower_lim) run = fmax(run, np_);
      suffix[t] = run;
    }
    poisoned = bad ? 1 : 0;
  }
  __syncthreads();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int kk = keep[i];
  double v = poisoned ? NAN : suffix[kk > T ? T : kk];
  if (accumulate) {
    // running sum for the batch uncertainty map: NaN (= "no detection") counts as 0
    out[i] += (O)((poisoned || v == 0.0) ? 0.0 : v);
  } else {
    out[i] = (O)((v == 0.0) ? NAN : v);
  }
}

static Roi make_roi(const chips_plan* p, int cx, int cy, int r) {
  Roi g;
  g.x0 = p->roi_x0; g.y0 = p->roi_y0; g.w = p->roi_w; g.h = p->roi_h;
  g.W = p->W; g.H = p->H; g.cx = cx; g.cy = cy; g.r = r;
  return g;
}

}  // namespace chips

using namespace chips;

extern "C" int chips_cleanup(const chips_plan* plan, int cx, int cy, int r, double disk_area,
                             double area_threshold, void* ws, void* stream) {
  if (!plan || !ws) return CHIPS_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)ws;
  const int T = plan->T;
  const size_t n_img = (size_t)plan->H * plan->W;
  const uint8_t* lev = w + plan->off_level;
  uint8_t* Wimg = w + plan->off_fill;
  uint8_t* keep = w + plan->off_keep;
  uint32_t* parent_b = reinterpret_cast<uint32_t*>(w + plan->off_parent_b);
  uint32_t* parent_c = reinterpret_cast<uint32_t*>(w + plan->off_parent_c);
  uint32_t* area = reinterpret_cast<uint32_t*>(w + plan->off_area);
  chips_result* res = reinterpret_cast<chips_result*>(w + plan->off_result);
  Roi g = make_roi(plan, cx, cy, r);
  dim3 blk(256);
  dim3 grid((g.w + 255) / 256, g.h);
  dim3 gridq((g.w + 1 + 255) / 256, g.h + 1);

  CHIPS_CUDA(cudaMemsetAsync(Wimg, T, n_img, st));
  CHIPS_CUDA(cudaMemsetAsync(keep, T, n_img, st));
  k_clear_counts<<<1, CHIPS_MAX_T, 0, st>>>(res);
  k_fill_init<<<grid, blk, 0, st>>>(g, Wimg, parent_b);
  CHIPS_CHECK_LAUNCH();
  // round T-1: the bulk background (pixels in no mask at all)
  k_fill_bulk_link<<<grid, blk, 0, st>>>(g, lev, T, parent_b);
  k_fill_union<true><<<grid, blk, 0, st>>>(g, lev, T - 1, parent_b);
  k_fill_assign<<<grid, blk, 0, st>>>(g, lev, T - 1, Wimg, parent_b);
  CHIPS_CHECK_LAUNCH();
  for (int t = T - 2; t >= 0; --t) {
    k_fill_union<false><<<grid, blk, 0, st>>>(g, lev, t, parent_b);
    k_fill_assign<<<grid, blk, 0, st>>>(g, lev, t, Wimg, parent_b);
  }
  CHIPS_CHECK_LAUNCH();
  dim3 gridz(grid.x, grid.y, T), gridqz(gridq.x, gridq.y, T);
  k_ccl_link<<<gridz, blk, 0, st>>>(g, Wimg, parent_c, area);
  k_ccl_union<<<gridz, blk, 0, st>>>(g, Wimg, parent_c);
  k_ccl_area<<<gridqz, blk, 0, st>>>(g, Wimg, parent_c, area);
  k_ccl_stats<<<gridz, blk, 0, st>>>(g, Wimg, parent_c, area, disk_area, area_threshold, res);
  k_ccl_keep<<<grid, blk, 0, st>>>(g, Wimg, T, parent_c, area, disk_area, area_threshold, keep);
  CHIPS_CHECK_LAUNCH();
  count_launch(10 + 2 * (T - 1));
  return CHIPS_OK;
}

static const uint8_t* pick_src(const chips_plan* plan, int which, unsigned char* w) {
  return w + (which == 1 ? plan->off_level : (which == 2 ? plan->off_fill : plan->off_keep));
}

extern "C" int chips_expand_maps(const chips_plan* plan, int which, uint8_t* maps, void* ws,
                                 void* stream) {
  if (!plan || !maps || !ws) return CHIPS_E_ARG;
  size_t n = (size_t)plan->H * plan->W;
  unsigned grid = (unsigned)((n / 4 + 256) / 256);
  k_expand_all<<<grid, 256, 0, (cudaStream_t)stream>>>(pick_src(plan, which, (unsigned char*)ws), n,
                                                       plan->T, maps);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_expand_map(const chips_plan* plan, int which, int t, void* out, int out_elem,
                                void* ws, void* stream) {
  if (!plan || !out || !ws || t < 0 || t >= plan->T) return CHIPS_E_ARG;
  size_t n = (size_t)plan->H * plan->W;
  unsigned grid = (unsigned)((n + 255) / 256);
  const uint8_t* src = pick_src(plan, which, (unsigned char*)ws);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_elem == 1) k_expand_one<uint8_t><<<grid, 256, 0, st>>>(src, n, t, (uint8_t*)out);
  else if (out_elem == 4) k_expand_one<float><<<grid, 256, 0, st>>>(src, n, t, (float*)out);
  else if (out_elem == 8) k_expand_one<double><<<grid, 256, 0, st>>>(src, n, t, (double*)out);
  else return CHIPS_E_ARG;
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_roots(const chips_plan* plan, int cx, int cy, int r, int t, int32_t* roots,
                           uint32_t* twice_area, void* ws, void* stream) {
  if (!plan || !roots || !ws || t < 0 || t >= plan->T) return CHIPS_E_ARG;
  unsigned char* w = (unsigned char*)ws;
  Roi g = make_roi(plan, cx, cy, r);
  dim3 grid((plan->W + 255) / 256, plan->H);
  k_roots<<<grid, 256, 0, (cudaStream_t)stream>>>(
      g, w + plan->off_fill, t, reinterpret_cast<uint32_t*>(w + plan->off_parent_c),
      reinterpret_cast<const uint32_t*>(w + plan->off_area), roots, twice_area);
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}

extern "C" int chips_prob_stack(const chips_plan* plan, double lower_lim, void* out, int out_elem,
                                int accumulate, void* ws, void* stream) {
  if (!plan || !out || !ws) return CHIPS_E_ARG;
  unsigned char* w = (unsigned char*)ws;
  size_t n = (size_t)plan->H * plan->W;
  unsigned grid = (unsigned)((n + 255) / 256);
  const uint8_t* keep = w + plan->off_keep;
  const chips_result* res = reinterpret_cast<const chips_result*>(w + plan->off_result);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_elem == 4) k_prob_stack<float><<<grid, 256, 0, st>>>(keep, n, res, lower_lim, accumulate, (float*)out);
  else if (out_elem == 8) k_prob_stack<double><<<grid, 256, 0, st>>>(keep, n, res, lower_lim, accumulate, (double*)out);
  else return CHIPS_E_ARG;
  CHIPS_CHECK_LAUNCH();
  count_launch();
  return CHIPS_OK;
}
