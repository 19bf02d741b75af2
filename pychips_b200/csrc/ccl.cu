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
// Row-run linking for the pixels of one level set (`sel`): every pixel points at the start
// of its run inside the warp's 32-pixel span, a run that continues across the span boundary
// points at the last pixel of the previous span (chains hop 32 pixels at a time).  Returns
// the link target (or `me` when not selected).  Must be called by all 32 lanes; lanes of a
// warp hold 32 consecutive pixels of one row.
template <typename Sel>
__device__ __forceinline__ uint32_t row_run_target(const Roi& g, int x, int y, bool valid,
                                                   uint32_t me, Sel sel) {
  int lane = threadIdx.x & 31;
  bool act = valid && sel(x, y);
  unsigned m = __ballot_sync(0xFFFFFFFFu, act);
  bool left_act = false;
  if (lane == 0 && act) left_act = (x - 1 >= g.x0) && sel(x - 1, y);
  left_act = __shfl_sync(0xFFFFFFFFu, left_act, 0);
  if (!act) return me;
  unsigned z = ~m & ((1u << lane) - 1u);
  int s = z ? (32 - __clz(z)) : 0;
  if (s == 0 && left_act) return me - (uint32_t)lane - 1u;   // last pixel of previous span
  return me - (uint32_t)(lane - s);
}

// B1: parent init for every ROI pixel (run links for the bulk background = pixels in no
// mask at all, self otherwise) and W = 0 on-disk / T off-disk.
__global__ void __launch_bounds__(256) k_fill_bulk_link(Roi g, const uint8_t* __restrict__ lev, int T,
                                                        uint8_t* __restrict__ Wimg,
                                                        uint32_t* __restrict__ parent) {
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) parent[0] = kBorder;
  bool valid = x < g.x0 + g.w;
  uint32_t me = valid ? g.li(x, y) + 1u : 0u;
  uint32_t tgt = row_run_target(g, x, y, valid, me, [&](int xx, int yy) {
    return on_disk(xx, yy, g.cx, g.cy, g.r) && lev[(size_t)yy * g.W + xx] == T;
  });
  if (!valid) return;
  parent[me] = tgt;
  Wimg[(size_t)y * g.W + x] = on_disk(x, y, g.cx, g.cy, g.r) ? 0 : (uint8_t)T;
}

__device__ __forceinline__ int level_at(const Roi& g, const uint8_t* __restrict__ lev, int xx, int yy) {
  if (!g.inside(xx, yy)) return -1;            // outside the disk / image: the border set
  return lev[(size_t)yy * g.W + xx];
}

// B2: vertical + border unions of the bulk background (left edges are the run links).
__device__ __forceinline__ void bulk_union_pixel(const Roi& g, const uint8_t* __restrict__ lev, int T,
                                                 uint32_t* __restrict__ parent, int x, int y) {
  if (lev[(size_t)y * g.W + x] != T || !on_disk(x, y, g.cx, g.cy, g.r)) return;
  uint32_t me = g.li(x, y) + 1;
  int l_l = level_at(g, lev, x - 1, y), l_r = level_at(g, lev, x + 1, y);
  int l_u = level_at(g, lev, x, y - 1), l_d = level_at(g, lev, x, y + 1);
  if (l_l < 0 || l_r < 0 || l_u < 0 || l_d < 0) uf_unite(parent, me, kBorder);
  if (l_u == T) {
    bool implied = (l_l == T) && (level_at(g, lev, x - 1, y - 1) == T);
    if (!implied) uf_unite(parent, me, me - (uint32_t)g.w);
  }
}

// 4 pixels per thread.  Deep inside the quiet Sun nothing has to be done: if the 6x3 pixel
// neighbourhood of the group is entirely bulk and entirely on the disk, every vertical edge
// is implied by the row-run links and no pixel touches the border.
__global__ void __launch_bounds__(256) k_fill_bulk_union(Roi g, const uint8_t* __restrict__ lev, int T,
                                                         uint32_t* __restrict__ parent) {
  const int x4 = (g.x0 & ~3) + 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int y = g.y0 + blockIdx.y;
  if (x4 >= g.x0 + g.w) return;
  if ((g.W & 3) == 0 && x4 >= g.x0 + 1 && x4 + 5 <= g.x0 + g.w && y > g.y0 && y + 1 < g.y0 + g.h) {
    // the disk is convex: its 4 corners on-disk => the whole 6x3 box is
    if (on_disk(x4 - 1, y - 1, g.cx, g.cy, g.r) && on_disk(x4 + 4, y - 1, g.cx, g.cy, g.r) &&
        on_disk(x4 - 1, y + 1, g.cx, g.cy, g.r) && on_disk(x4 + 4, y + 1, g.cx, g.cy, g.r)) {
      const uint32_t T4 = (uint32_t)T * 0x01010101u;
      const uint8_t* c = lev + (size_t)y * g.W + x4;
      const uint32_t cur = *reinterpret_cast<const uint32_t*>(c);
      const uint32_t up = *reinterpret_cast<const uint32_t*>(c - g.W);
      const uint32_t dn = *reinterpret_cast<const uint32_t*>(c + g.W);
      if (cur == T4 && up == T4 && dn == T4 && c[-1] == T && c[4] == T && c[-g.W - 1] == T) return;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x4 + i;
    if (x >= g.x0 && x < g.x0 + g.w) bulk_union_pixel(g, lev, T, parent, x, y);
  }
}

// B3/B5: bulk pixels that reached the border get W = T; everything else on the disk (pixels
// of some mask, and bulk pixels enclosed by masks) is "pending".  PASS 0 counts pending pixels
// per level, PASS 1 scatters their ROI-local indices into a list grouped by level.
template <int PASS>
__global__ void __launch_bounds__(256) k_fill_pending(Roi g, const uint8_t* __restrict__ lev, int T,
                                                      uint8_t* __restrict__ Wimg,
                                                      uint32_t* __restrict__ parent,
                                                      uint32_t* __restrict__ lvl_count,
                                                      uint32_t* __restrict__ lvl_cursor,
                                                      uint32_t* __restrict__ pending) {
  __shared__ uint32_t scnt[CHIPS_MAX_T + 1];
  if (PASS == 0) {
    for (int i = threadIdx.x; i <= T; i += blockDim.x) scnt[i] = 0;
    __syncthreads();
  }
  int x = g.x0 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = g.y0 + blockIdx.y;
  int mylev = -1;
  bool bulk = false;
  if (x < g.x0 + g.w && on_disk(x, y, g.cx, g.cy, g.r)) {
    size_t o = (size_t)y * g.W + x;
    int l = lev[o];
    if (PASS == 0) {
      if (l == T) bulk = true;
      else mylev = l;
    } else {
      if (!(l == T && Wimg[o] == T)) mylev = l;
    }
  }
  if (PASS == 0) {
    // bulk pixels of one row run share their root: the first lane of each run inside the warp's
    // span finds it, the others copy it.  Everybody is flattened (later finds are one hop).
    const int lane = threadIdx.x & 31;
    const unsigned mb = __ballot_sync(0xFFFFFFFFu, bulk);
    const unsigned z = ~mb & ((1u << lane) - 1u);
    const int s = bulk ? (z ? (32 - __clz(z)) : 0) : lane;
    uint32_t root = 0;
    const uint32_t me = bulk ? g.li(x, y) + 1 : 0u;
    if (bulk && s == lane) root = uf_find(parent, me);
    root = __shfl_sync(0xFFFFFFFFu, root, s);
    if (bulk) {
      if (root != me) parent[me] = root;
      if (root == kBorder) Wimg[(size_t)y * g.W + x] = (uint8_t)T;
      else mylev = T;
    }
  }
  unsigned act = __ballot_sync(0xFFFFFFFFu, mylev >= 0);
  if (mylev >= 0) {
    unsigned peers = __match_any_sync(act, mylev);
    int lane = threadIdx.x & 31;
    int leader = __ffs(peers) - 1;
    if (PASS == 0) {
      if (lane == leader) atomicAdd(&scnt[mylev], __popc(peers));
    } else {
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&lvl_cursor[mylev], __popc(peers));
      base = __shfl_sync(peers, base, leader);
      pending[base + __popc(peers & ((1u << lane) - 1u))] = g.li(x, y);
    }
  }
  if (PASS == 0) {
    __syncthreads();
    for (int i = threadIdx.x; i <= T; i += blockDim.x)
      if (scnt[i]) atomicAdd(&lvl_count[i], scnt[i]);
  }
}

// B4: offsets[l] = first list slot of level l (offsets[T+1] = total); cursor = copy.
__global__ void k_fill_offsets(int T, uint32_t* __restrict__ lvl_count,
                               uint32_t* __restrict__ lvl_offset, uint32_t* __restrict__ lvl_cursor) {
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int l = 0; l <= T; ++l) {
      lvl_offset[l] = run;
      lvl_cursor[l] = run;
      run += lvl_count[l];
    }
    lvl_offset[T + 1] = run;
  }
}

// Round t (descending): the pending pixels of level t+1 join the background.  The roots of
// the pixel and of its (up to 5) union partners are chased together so the dependent global
// loads of the finds overlap; the hooks themselves go through uf_unite.
__global__ void __launch_bounds__(256) k_fill_round_union(Roi g, const uint8_t* __restrict__ lev, int t,
                                                          const uint32_t* __restrict__ lvl_offset,
                                                          const uint32_t* __restrict__ pending,
                                                          uint32_t* __restrict__ parent) {
  const int lv = t + 1;
  const uint32_t beg = lvl_offset[lv], end = lvl_offset[lv + 1];
  for (uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    uint32_t li = pending[i];
    int y = g.y0 + (int)(li / (uint32_t)g.w), x = g.x0 + (int)(li % (uint32_t)g.w);
    uint32_t me = li + 1;
    int l_l = level_at(g, lev, x - 1, y), l_r = level_at(g, lev, x + 1, y);
    int l_u = level_at(g, lev, x, y - 1), l_d = level_at(g, lev, x, y + 1);
    uint32_t node[5];
    bool use[5];
    use[0] = (l_l < 0 || l_r < 0 || l_u < 0 || l_d < 0); node[0] = kBorder;
    use[1] = l_l >= lv;  node[1] = me - 1u;
    use[2] = l_u >= lv;  node[2] = me - (uint32_t)g.w;
    use[3] = l_r > lv;   node[3] = me + 1u;
    use[4] = l_d > lv;   node[4] = me + (uint32_t)g.w;
    if (!(use[0] | use[1] | use[2] | use[3] | use[4])) continue;
    uint32_t rm = me;
#pragma unroll
    for (int j = 0; j < 5; ++j) if (!use[j]) node[j] = me;
    for (int hop = 0; hop < 64; ++hop) {        // bounded: uf_unite finishes whatever is left
      uint32_t pm = parent[rm], p0 = parent[node[0]], p1 = parent[node[1]], p2 = parent[node[2]],
               p3 = parent[node[3]], p4 = parent[node[4]];
      bool done = (pm == rm) & (p0 == node[0]) & (p1 == node[1]) & (p2 == node[2]) &
                  (p3 == node[3]) & (p4 == node[4]);
      rm = pm; node[0] = p0; node[1] = p1; node[2] = p2; node[3] = p3; node[4] = p4;
      if (done) break;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (use[j] && node[j] != rm) uf_unite(parent, rm, node[j]);
  }
}

// ... and every active (level > t), still unassigned pending pixel that now reaches the
// border gets W = t+1.  Active pending pixels are the list suffix starting at level t+1.
__global__ void __launch_bounds__(256) k_fill_round_assign(Roi g, int t, int T,
                                                           const uint32_t* __restrict__ lvl_offset,
                                                           const uint32_t* __restrict__ pending,
                                                           uint8_t* __restrict__ Wimg,
                                                           uint32_t* __restrict__ parent) {
  const uint32_t beg = lvl_offset[t + 1], end = lvl_offset[T + 1];
  for (uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += gridDim.x * blockDim.x) {
    uint32_t li = pending[i];
    int y = g.y0 + (int)(li / (uint32_t)g.w), x = g.x0 + (int)(li % (uint32_t)g.w);
    size_t o = (size_t)y * g.W + x;
    if (Wimg[o] != 0) continue;
    if (uf_find(parent, li + 1) == kBorder) Wimg[o] = (uint8_t)(t + 1);
  }
}

// ------------------------------------------------------------------ phase C
// Work is restricted to "active" 32x8 tiles (tiles that hold at least one pixel of some
// filled mask, W < T); each of the T labellings walks the same tile list (blockIdx.y = t).
constexpr int kTileW = 32, kTileH = 8;

__global__ void __launch_bounds__(256) k_tile_list(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                   int tiles_x, int tiles_y,
                                                   uint32_t* __restrict__ tile_count,
                                                   uint32_t* __restrict__ tiles) {
  // one warp per tile: lane = column, loop over the 8 rows
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= tiles_x * tiles_y) return;
  int tj = warp % tiles_x, ti = warp / tiles_x;
  int x = g.x0 + tj * kTileW + lane;
  bool any = false;
  if (x < g.x0 + g.w) {
    for (int dy = 0; dy < kTileH; ++dy) {
      int y = g.y0 + ti * kTileH + dy;
      if (y < g.y0 + g.h) any |= Wimg[(size_t)y * g.W + x] < T;
    }
  }
  if (__any_sync(0xFFFFFFFFu, any) && lane == 0) tiles[atomicAdd(tile_count, 1u)] = (uint32_t)warp;
}

struct TileCtx {
  int x, y;
  bool valid;
};
__device__ __forceinline__ TileCtx tile_pixel(const Roi& g, uint32_t tile, int tiles_x) {
  int tj = tile % tiles_x, ti = tile / tiles_x;
  TileCtx c;
  c.x = g.x0 + tj * kTileW + (threadIdx.x & 31);
  c.y = g.y0 + ti * kTileH + (threadIdx.x >> 5);
  c.valid = c.x < g.x0 + g.w && c.y < g.y0 + g.h;
  return c;
}

__device__ __forceinline__ bool fg_at(const Roi& g, const uint8_t* __restrict__ Wimg, int t, int xx,
                                      int yy) {
  return xx >= g.x0 && xx < g.x0 + g.w && yy >= g.y0 && yy < g.y0 + g.h &&
         Wimg[(size_t)yy * g.W + xx] <= t;
}

// W of a pixel as an int, 255 (= never foreground) outside the ROI
__device__ __forceinline__ int w_at(const Roi& g, const uint8_t* __restrict__ Wimg, int xx, int yy) {
  return (xx >= g.x0 && xx < g.x0 + g.w && yy >= g.y0 && yy < g.y0 + g.h)
             ? (int)Wimg[(size_t)yy * g.W + xx] : 255;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  return v;
}

// All T labellings in one sweep over the active tiles: every thread loads the W values it needs
// ONCE and loops over the thresholds t >= W (a pixel is foreground from its W upwards).  The
// loop variable is warp-uniform (starts at the smallest W of the warp's 32-pixel row span) so
// the ballots / match_any inside see consistent participants.
__global__ void __launch_bounds__(256) k_ccl_link(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                  int tiles_x, const uint32_t* __restrict__ tile_count,
                                                  const uint32_t* __restrict__ tiles,
                                                  uint32_t* __restrict__ parent_all,
                                                  uint32_t* __restrict__ area_all) {
  const size_t n = (size_t)g.w * g.h;
  const uint32_t nt = *tile_count;
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    TileCtx c = tile_pixel(g, tiles[i], tiles_x);
    const int w = c.valid ? (int)Wimg[(size_t)c.y * g.W + c.x] : 255;
    int wl = 255;                                  // pixel left of the span (lane 0 only)
    if (lane == 0) wl = w_at(g, Wimg, c.x - 1, c.y);
    wl = __shfl_sync(0xFFFFFFFFu, wl, 0);
    const uint32_t me = c.valid ? g.li(c.x, c.y) : 0u;
    for (int t = warp_min(w); t < T; ++t) {
      const bool act = w <= t;
      const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
      if (!act) continue;
      const unsigned z = ~m & ((1u << lane) - 1u);
      const int s = z ? (32 - __clz(z)) : 0;
      uint32_t tgt = (s == 0 && wl <= t) ? me - (uint32_t)lane - 1u : me - (uint32_t)(lane - s);
      parent_all[(size_t)t * n + me] = tgt;
      area_all[(size_t)t * n + me] = 0;
    }
  }
}

__global__ void __launch_bounds__(256) k_ccl_union(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                   int tiles_x, const uint32_t* __restrict__ tile_count,
                                                   const uint32_t* __restrict__ tiles,
                                                   uint32_t* __restrict__ parent_all) {
  const size_t n = (size_t)g.w * g.h;
  const uint32_t nt = *tile_count;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    TileCtx c = tile_pixel(g, tiles[i], tiles_x);
    if (!c.valid || c.y == g.y0) continue;
    const int x = c.x, y = c.y;
    const int w = Wimg[(size_t)y * g.W + x];
    if (w >= T) continue;
    const int wu = w_at(g, Wimg, x, y - 1), wul = w_at(g, Wimg, x - 1, y - 1);
    const int wur = w_at(g, Wimg, x + 1, y - 1), wlf = w_at(g, Wimg, x - 1, y);
    const int wrt = w_at(g, Wimg, x + 1, y);
    if (min(wu, min(wul, wur)) >= T) continue;     // nothing above in any mask
    const uint32_t me = g.li(x, y);
    for (int t = w; t < T; ++t) {
      uint32_t* parent = parent_all + (size_t)t * n;
      const bool up = wu <= t, ul = wul <= t, ur = wur <= t, lf = wlf <= t, rt = wrt <= t;
      if (up) {
        if (!(lf && ul)) uf_unite(parent, me, me - (uint32_t)g.w);
      } else {
        if (ul && !lf) uf_unite(parent, me, me - (uint32_t)g.w - 1u);
        if (ur && !rt) uf_unite(parent, me, me - (uint32_t)g.w + 1u);
      }
    }
  }
}

// 2*area = 2*Q4 + Q3 over every 2x2 quad of the zero-padded filled set.  Each quad is charged
// to its first set pixel in raster order, so only two quads per foreground pixel p matter:
//   A = {p, right, down, down-right}              (p is the top-left pixel)
//   B = {left, p, down-left, down} with left unset (p is the top-right pixel; s = 3 at most)
__global__ void __launch_bounds__(256) k_ccl_area(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                  int tiles_x, const uint32_t* __restrict__ tile_count,
                                                  const uint32_t* __restrict__ tiles,
                                                  uint32_t* __restrict__ parent_all,
                                                  uint32_t* __restrict__ area_all) {
  const size_t n = (size_t)g.w * g.h;
  const uint32_t nt = *tile_count;
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    TileCtx c = tile_pixel(g, tiles[i], tiles_x);
    const int x = c.x, y = c.y;
    const int w = c.valid ? (int)Wimg[(size_t)y * g.W + x] : 255;
    int wr = 255, wd = 255, wdr = 255, wl = 255, wdl = 255;
    if (w < T) {
      wr = w_at(g, Wimg, x + 1, y);
      wd = w_at(g, Wimg, x, y + 1);
      wdr = w_at(g, Wimg, x + 1, y + 1);
      wl = w_at(g, Wimg, x - 1, y);
      wdl = w_at(g, Wimg, x - 1, y + 1);
    }
    const uint32_t me = c.valid ? g.li(x, y) : 0u;
    for (int t = warp_min(w); t < T; ++t) {
      uint32_t root = 0xFFFFFFFFu, add = 0;
      if (w <= t) {
        const int sa = 1 + (int)(wr <= t) + (int)(wd <= t) + (int)(wdr <= t);
        add = (sa == 4) ? 2u : (sa == 3 ? 1u : 0u);
        if (wd <= t && wl > t && wdl <= t) add += 1u;
        // flatten while we are here: stats / keep / roots then find every root in one hop
        uint32_t* parent = parent_all + (size_t)t * n;
        root = uf_find(parent, me);
        if (root != me) parent[me] = root;
      }
      const unsigned active = __ballot_sync(0xFFFFFFFFu, add != 0);
      if (add) {      // warp-aggregate per root before the global atomic
        const unsigned peers = __match_any_sync(active, root);
        const uint32_t sum = __reduce_add_sync(peers, add);
        if (lane == __ffs(peers) - 1) atomicAdd(&area_all[(size_t)t * n + root], sum);
      }
    }
  }
}

__device__ __forceinline__ bool area_kept(uint32_t a2, double disk_area, double thr) {
  // chips.py:424-426: cv2.contourArea(cnt) / self.disk_area >= self.area_threshold
  return __ddiv_rn(__dmul_rn((double)a2, 0.5), disk_area) >= thr;
}

__global__ void __launch_bounds__(256) k_ccl_stats(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                   int tiles_x, const uint32_t* __restrict__ tile_count,
                                                   const uint32_t* __restrict__ tiles,
                                                   const uint32_t* __restrict__ parent_all,
                                                   const uint32_t* __restrict__ area_all,
                                                   double disk_area, double thr,
                                                   chips_result* __restrict__ res) {
  __shared__ int s_root[CHIPS_MAX_T], s_kept[CHIPS_MAX_T];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    s_root[t] = 0;
    s_kept[t] = 0;
  }
  __syncthreads();
  const size_t n = (size_t)g.w * g.h;
  const uint32_t nt = *tile_count;
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    TileCtx c = tile_pixel(g, tiles[i], tiles_x);
    const int w = c.valid ? (int)Wimg[(size_t)c.y * g.W + c.x] : 255;
    const uint32_t me = c.valid ? g.li(c.x, c.y) : 0u;
    for (int t = warp_min(w); t < T; ++t) {
      bool root = false, kept = false;
      if (w <= t && parent_all[(size_t)t * n + me] == me) {
        root = true;
        kept = area_kept(area_all[(size_t)t * n + me], disk_area, thr);
      }
      const unsigned mr = __ballot_sync(0xFFFFFFFFu, root), mk = __ballot_sync(0xFFFFFFFFu, kept);
      if (lane == 0) {
        if (mr) atomicAdd(&s_root[t], __popc(mr));
        if (mk) atomicAdd(&s_kept[t], __popc(mk));
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    if (s_root[t]) atomicAdd(&res->n_comp[t], s_root[t]);
    if (s_kept[t]) atomicAdd(&res->n_kept[t], s_kept[t]);
  }
}

__global__ void __launch_bounds__(256) k_ccl_keep(Roi g, const uint8_t* __restrict__ Wimg, int T,
                                                  int tiles_x, const uint32_t* __restrict__ tile_count,
                                                  const uint32_t* __restrict__ tiles,
                                                  uint32_t* __restrict__ parent_all,
                                                  const uint32_t* __restrict__ area_all,
                                                  double disk_area, double thr,
                                                  uint8_t* __restrict__ keep) {
  const size_t n = (size_t)g.w * g.h;
  const uint32_t nt = *tile_count;
  for (uint32_t i = blockIdx.x; i < nt; i += gridDim.x) {
    TileCtx c = tile_pixel(g, tiles[i], tiles_x);
    if (!c.valid) continue;
    size_t o = (size_t)c.y * g.W + c.x;
    int w = Wimg[o];
    if (w >= T) continue;                      // keep[] is pre-filled with T
    int kk = T;
    uint32_t me = g.li(c.x, c.y);
    for (int t = w; t < T; ++t) {
      uint32_t root = uf_find(parent_all + (size_t)t * n, me);
      if (area_kept(area_all[(size_t)t * n + root], disk_area, thr)) {
        kk = t;
        break;
      }
    }
    keep[o] = (uint8_t)kk;
  }
}

__global__ void k_cleanup_reset(chips_result* res, uint32_t* lvl_count, uint32_t* tile_count) {
  int t = threadIdx.x;
  if (t < CHIPS_MAX_T) {
    res->n_kept[t] = 0;
    res->n_comp[t] = 0;
  }
  if (t <= CHIPS_MAX_T + 1) lvl_count[t] = 0;
  if (t == 0) *tile_count = 0;
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
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    int kk = keep[i];
    double v = poisoned ? NAN : suffix[kk > T ? T : kk];
    if (accumulate) {
      // running sum for the batch uncertainty map: "no detection" counts as 0
      if (!poisoned && v != 0.0) out[i] += (O)v;
    } else {
      out[i] = (O)((v == 0.0) ? NAN : v);
    }
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
  if (!plan || !ws || !plan->masked) return CHIPS_E_ARG;
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
  uint32_t* pending = reinterpret_cast<uint32_t*>(w + plan->off_pending);
  uint32_t* lvl = reinterpret_cast<uint32_t*>(w + plan->off_lvl);       // count | offset | cursor
  uint32_t* lvl_count = lvl, *lvl_offset = lvl + 128, *lvl_cursor = lvl + 256;
  uint32_t* tile_count = lvl + 384;
  uint32_t* tiles = reinterpret_cast<uint32_t*>(w + plan->off_tiles);
  chips_result* res = reinterpret_cast<chips_result*>(w + plan->off_result);
  Roi g = make_roi(plan, cx, cy, r);
  dim3 blk(256);
  dim3 grid((g.w + 255) / 256, g.h);
  const int tiles_x = (g.w + kTileW - 1) / kTileW, tiles_y = (g.h + kTileH - 1) / kTileH;
  const int ntiles = tiles_x * tiles_y;

  CHIPS_CUDA(cudaMemsetAsync(Wimg, T, n_img, st));
  CHIPS_CUDA(cudaMemsetAsync(keep, T, n_img, st));
  k_cleanup_reset<<<1, 128, 0, st>>>(res, lvl_count, tile_count);
  // round T-1: the bulk background (pixels in no mask at all)
  k_fill_bulk_link<<<grid, blk, 0, st>>>(g, lev, T, Wimg, parent_b);
  dim3 grid4((g.w / 4 + 2 + 255) / 256, g.h);
  k_fill_bulk_union<<<grid4, blk, 0, st>>>(g, lev, T, parent_b);
  k_fill_pending<0><<<grid, blk, 0, st>>>(g, lev, T, Wimg, parent_b, lvl_count, lvl_cursor, pending);
  k_fill_offsets<<<1, 32, 0, st>>>(T, lvl_count, lvl_offset, lvl_cursor);
  k_fill_pending<1><<<grid, blk, 0, st>>>(g, lev, T, Wimg, parent_b, lvl_count, lvl_cursor, pending);
  CHIPS_CHECK_LAUNCH();
  // the bulk pixels that did not reach the border in round T-1 stay pending at level T and
  // are re-examined by every later round (suffix of the list)
  const int gl = 2 * kNumSM;
  for (int t = T - 2; t >= 0; --t) {
    k_fill_round_union<<<gl, blk, 0, st>>>(g, lev, t, lvl_offset, pending, parent_b);
    k_fill_round_assign<<<gl, blk, 0, st>>>(g, t, T, lvl_offset, pending, Wimg, parent_b);
  }
  CHIPS_CHECK_LAUNCH();
  k_tile_list<<<(ntiles * 32 + 255) / 256, blk, 0, st>>>(g, Wimg, T, tiles_x, tiles_y, tile_count, tiles);
  const int gridt = 8 * kNumSM;
  k_ccl_link<<<gridt, blk, 0, st>>>(g, Wimg, T, tiles_x, tile_count, tiles, parent_c, area);
  k_ccl_union<<<gridt, blk, 0, st>>>(g, Wimg, T, tiles_x, tile_count, tiles, parent_c);
  k_ccl_area<<<gridt, blk, 0, st>>>(g, Wimg, T, tiles_x, tile_count, tiles, parent_c, area);
  k_ccl_stats<<<gridt, blk, 0, st>>>(g, Wimg, T, tiles_x, tile_count, tiles, parent_c, area,
                                     disk_area, area_threshold, res);
  k_ccl_keep<<<8 * kNumSM, blk, 0, st>>>(g, Wimg, T, tiles_x, tile_count, tiles, parent_c, area,
                                         disk_area, area_threshold, keep);
  CHIPS_CHECK_LAUNCH();
  count_launch(6 + 2 * (T - 1) + 6);
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
  unsigned grid = 8 * kNumSM;
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
