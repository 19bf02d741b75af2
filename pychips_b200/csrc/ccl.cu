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
// border set.  Phase C: the filled masks are nested too, so ONE union-find grown through the
// ASCENDING levels gives all T labellings as a component tree (see "phase C" below).
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
    CHIPS_BOUNDS(x >= x0 && x < x0 + w && y >= y0 && y < y0 + h);
    return (uint32_t)((y - y0) * w + (x - x0));
  }
  __device__ __forceinline__ uint32_t npix() const { return (uint32_t)w * (uint32_t)h; }
};

// ------------------------------------------------------------------ phase B
// B1: parent init for every ROI pixel, one CTA per ROI row.  A bulk-background pixel (on the disk,
// in no mask at all) points at the FIRST pixel of its row run - warp ballots, the masks of the
// CTA's eight warps and a carry across the 256-pixel chunks of the row - so every later find on
// a bulk pixel is one hop to its run start; everything else is a singleton.  (Round 1 linked runs
// inside 32-pixel spans only: a 3000-pixel run of quiet Sun was a chain of ~100 dependent hops.)
__global__ void __launch_bounds__(256) k_fill_bulk_link(Roi g, const uint8_t* __restrict__ lev, int T,
                                                        uint32_t* __restrict__ parent) {
  __shared__ uint32_t s_mask[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int y = g.y0 + blockIdx.x;
  if (blockIdx.x == 0 && tid == 0) parent[0] = kBorder;
  int xa, xb;                                           // on-disk columns of this row (xa = xb: none)
  warp_row_chord(y, g.cx, g.cy, g.r, g.x0, g.w, xa, xb);
  const int la = xa - g.x0, lb = xb - g.x0;
  const uint8_t* __restrict__ row = lev + (size_t)y * g.W + g.x0;
  const uint32_t node0 = (uint32_t)blockIdx.x * (uint32_t)g.w + 1u;      // node of local column 0
  int carry = 0;                 // local column after the last non-bulk pixel left of this chunk
  constexpr int U = 8;           // chunks whose loads are in flight together
  int it = 0;
  for (int base0 = 0; base0 < g.w; base0 += 256 * U) {
    int v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int lx = base0 + 256 * u + tid;
      v[u] = (lx >= la && lx < lb) ? (int)row[lx] : 255;       // [la, lb) lies inside [0, g.w)
    }
#pragma unroll
    for (int u = 0; u < U; ++u, ++it) {
      const int base = base0 + 256 * u;
      if (base >= g.w) break;
      const int lx = base + tid;
      const bool valid = lx < g.w;
      const bool act = v[u] == T;
      const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
      uint32_t* sm = s_mask[it & 1];
      if (lane == 0) sm[warp] = m;
      __syncthreads();
      // the eight masks, one per lane (replicated four times): which warps hold a non-bulk pixel,
      // the last one before mine (my run starts behind its last non-bulk pixel) and the last one
      // of the chunk (the carry), without a loop over the masks
      const uint32_t mq = sm[lane & 7];
      const unsigned nf = __ballot_sync(0xFFFFFFFFu, mq != 0xFFFFFFFFu) & 0xFFu;
      const unsigned below = nf & ((1u << warp) - 1u);
      const int wv = below ? 31 - __clz(below) : 0;
      const uint32_t mv = __shfl_sync(0xFFFFFFFFu, mq, wv);
      const int wl = nf ? 31 - __clz(nf) : 0;
      const uint32_t ml = __shfl_sync(0xFFFFFFFFu, mq, wl);
      const unsigned z = ~m & ((1u << lane) - 1u);
      const int start = z ? base + 32 * warp + (32 - __clz(z))
                          : (below ? base + 32 * wv + (32 - __clz(~mv)) : carry);
      CHIPS_BOUNDS(!act || (start >= 0 && start <= lx));
      if (valid) parent[node0 + (uint32_t)lx] = node0 + (uint32_t)(act ? start : lx);
      if (nf) carry = base + 32 * wl + (32 - __clz(~ml));
    }
  }
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
    // the disk is convex: its 4 corners on-disk => the whole 6x3 box is, no pixel of the group
    // touches the border and every neighbour level is in the three row words + three bytes
    if (on_disk(x4 - 1, y - 1, g.cx, g.cy, g.r) && on_disk(x4 + 4, y - 1, g.cx, g.cy, g.r) &&
        on_disk(x4 - 1, y + 1, g.cx, g.cy, g.r) && on_disk(x4 + 4, y + 1, g.cx, g.cy, g.r)) {
      const uint32_t T4 = (uint32_t)T * 0x01010101u;
      const uint8_t* c = lev + (size_t)y * g.W + x4;
      const uint32_t cur = *reinterpret_cast<const uint32_t*>(c);
      const uint32_t up = *reinterpret_cast<const uint32_t*>(c - g.W);
      const uint32_t dn = *reinterpret_cast<const uint32_t*>(c + g.W);
      const uint32_t cl = c[-1], ul0 = c[-g.W - 1];
      if (cur == T4 && up == T4 && dn == T4 && cl == (uint32_t)T && c[4] == T && ul0 == (uint32_t)T) return;
      const uint32_t me0 = g.li(x4, y) + 1u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (((cur >> (8 * i)) & 255u) != (uint32_t)T || ((up >> (8 * i)) & 255u) != (uint32_t)T) continue;
        const uint32_t l_l = i ? (cur >> (8 * (i - 1))) & 255u : cl;
        const uint32_t l_ul = i ? (up >> (8 * (i - 1))) & 255u : ul0;
        if (!(l_l == (uint32_t)T && l_ul == (uint32_t)T)) uf_unite(parent, me0 + (uint32_t)i, me0 + (uint32_t)i - (uint32_t)g.w);
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x4 + i;
    if (x >= g.x0 && x < g.x0 + g.w) bulk_union_pixel(g, lev, T, parent, x, y);
  }
}

// B3: bulk pixels that reached the border get W = T; everything else on the disk (pixels of some
// mask, and bulk pixels enclosed by masks) is "pending": W = 0, counted per level and appended (in
// no particular order) to `unsorted`.  Four pixels per thread, 1024 per CTA.  A bulk pixel's parent
// is the start of its row run (k_fill_bulk_link), so its root is two cached hops away: no leader
// election, the four chases of a thread are independent; the run starts themselves are flattened,
// other parents are not rewritten.
__global__ void __launch_bounds__(256) k_fill_pending(Roi g, const uint8_t* __restrict__ lev, int T,
                                                      uint8_t* __restrict__ Wimg,
                                                      uint32_t* __restrict__ parent,
                                                      uint32_t* __restrict__ lvl_count,
                                                      uint32_t* __restrict__ total,
                                                      uint32_t* __restrict__ unsorted) {
  __shared__ uint32_t scnt[CHIPS_MAX_T + 1];
  __shared__ uint32_t s_total, s_base;
  const int y = g.y0 + blockIdx.y;
  for (int i = threadIdx.x; i <= T; i += blockDim.x) scnt[i] = 0;
  if (threadIdx.x == 0) s_total = 0;
  int xa, xb;                                           // on-disk columns of this row, per warp (no
  warp_row_chord(y, g.cx, g.cy, g.r, g.x0, g.w, xa, xb);   // CTA waits on one thread's square root)
  __syncthreads();
  const int x4 = (g.x0 & ~3) + 4 * (int)(blockIdx.x * blockDim.x + threadIdx.x);
  const int lane = threadIdx.x & 31;
  const uint8_t* __restrict__ row = lev + (size_t)y * g.W;
  uint8_t* __restrict__ wrow = Wimg + (size_t)y * g.W;
  const bool some = x4 + 4 > xa && x4 < xb;           // the group holds an on-disk pixel
  const bool vec = (g.W & 3) == 0 && x4 + 4 <= g.W;    // x4 >= 0 is a multiple of 4
  uint32_t lw = 0;
  int lleft = -1;
  if (some) {
    if (vec) {
      lw = *reinterpret_cast<const uint32_t*>(row + x4);
    } else {
      for (int i = 0; i < 4; ++i)
        if (x4 + i < g.W) lw |= (uint32_t)row[x4 + i] << (8 * i);
    }
    if (x4 - 1 >= xa) lleft = row[x4 - 1];
  }
  bool disk[4], bulk[4], start[4];
  uint32_t me[4], s0[4], sp[4], rp[4];
  int mylev[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int x = x4 + i;
    const int l = (int)((lw >> (8 * i)) & 255u);
    disk[i] = some && x >= xa && x < xb;
    bulk[i] = disk[i] && l == T;
    const int left = i ? (int)((lw >> (8 * (i - 1))) & 255u) : lleft;
    start[i] = bulk[i] && !(x - 1 >= xa && left == T);
    me[i] = bulk[i] ? g.li(x, y) + 1u : 0u;
    mylev[i] = (disk[i] && !bulk[i]) ? l : -1;
  }
  // quiet Sun: all 128 pixels of the warp are bulk, i.e. one piece of one row run - lane 0 chases
  // its root for everybody
  const bool wfull = __all_sync(0xFFFFFFFFu, bulk[0] && bulk[1] && bulk[2] && bulk[3]);
  if (wfull) {
    uint32_t root = 0;
    if (lane == 0) {
      const uint32_t first = parent[me[0]];
      uint32_t sp0 = first, rp0 = parent[sp0];
      while (rp0 != sp0) {
        sp0 = rp0;
        rp0 = parent[sp0];
      }
      if (start[0] && rp0 != first) parent[me[0]] = rp0;
      root = rp0;
    }
    root = __shfl_sync(0xFFFFFFFFu, root, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rp[i] = root;
      if (root != kBorder) mylev[i] = T;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) s0[i] = sp[i] = bulk[i] ? parent[me[i]] : 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) rp[i] = parent[sp[i]];          // (node 0 is its own parent)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      while (rp[i] != sp[i]) {
        sp[i] = rp[i];
        rp[i] = parent[sp[i]];
      }
      if (start[i] && rp[i] != s0[i]) parent[me[i]] = rp[i];
      if (bulk[i] && rp[i] != kBorder) mylev[i] = T;
    }
  }
  // W: T for the bulk background that reaches the border (and off the disk: the memset), 0 = not
  // assigned yet for every other on-disk pixel
  if (some) {
    uint32_t ww = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) ww |= (uint32_t)((!disk[i] || (bulk[i] && rp[i] == kBorder)) ? T : 0) << (8 * i);
    if (vec) {
      *reinterpret_cast<uint32_t*>(wrow + x4) = ww;           // (off-disk bytes of the group: T as before)
    } else {
      for (int i = 0; i < 4; ++i)
        if (disk[i]) wrow[x4 + i] = (uint8_t)(ww >> (8 * i));
    }
  }
  // per-level counts and list slots: one shared atomic per warp, one global atomic per CTA
  const int np = (mylev[0] >= 0) + (mylev[1] >= 0) + (mylev[2] >= 0) + (mylev[3] >= 0);
  uint32_t slot = 0;
  if (__ballot_sync(0xFFFFFFFFu, np > 0)) {
    uint32_t incl = (uint32_t)np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += u;
    }
    uint32_t woff = 0;
    if (lane == 31) woff = atomicAdd(&s_total, incl);
    woff = __shfl_sync(0xFFFFFFFFu, woff, 31);
    slot = woff + incl - (uint32_t)np;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned act = __ballot_sync(0xFFFFFFFFu, mylev[i] >= 0);
      if (mylev[i] >= 0) {
        const unsigned peers = __match_any_sync(act, mylev[i]);
        if (lane == __ffs(peers) - 1) atomicAdd(&scnt[mylev[i]], __popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= T; i += blockDim.x)
    if (scnt[i]) atomicAdd(&lvl_count[i], scnt[i]);
  if (threadIdx.x == 0 && s_total) s_base = atomicAdd(total, s_total);
  __syncthreads();
  if (np) {
    slot += s_base;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (mylev[i] >= 0) {
        CHIPS_BOUNDS(slot < g.npix());
        unsorted[slot++] = g.li(x4 + i, y);
      }
  }
}

// B5 / C1b: regroup a pixel list by level: entry -> slot of its level's segment (cursor[] starts
// at the segment offsets).  `mode` 0: level = lev image value (phase B: pending pixels by lev),
// 1: level = ROI-local W (phase C: foreground pixels by W; 255 = not foreground, skipped).
// A CTA ranks the 512 entries of a tile per level in shared memory and reserves each level's
// slots with ONE global atomic (per-warp atomics on the ~20 cursors serialised in L2: 16 us).
constexpr int kScatterThreads = 512;
__global__ void __launch_bounds__(kScatterThreads) k_list_scatter(Roi g, const uint8_t* __restrict__ levsrc, int mode,
                                                                  const uint32_t* __restrict__ in,
                                                                  const uint32_t* __restrict__ n_ptr,
                                                                  uint32_t* __restrict__ cursor,
                                                                  uint32_t* __restrict__ out) {
  __shared__ uint32_t scnt[CHIPS_MAX_T + 1], sbase[CHIPS_MAX_T + 1];
  const uint32_t n = *n_ptr;
  const int lane = threadIdx.x & 31;
  for (uint32_t t0 = blockIdx.x * kScatterThreads; t0 < n; t0 += gridDim.x * kScatterThreads) {
    if (threadIdx.x <= CHIPS_MAX_T) scnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = t0 + threadIdx.x;
    int level = -1;
    uint32_t li = 0, rank = 0;
    if (i < n) {
      li = in[i];
      if (mode == 0) {
        const uint32_t ly = li / (uint32_t)g.w, lx = li - ly * (uint32_t)g.w;
        level = levsrc[(size_t)(g.y0 + (int)ly) * g.W + (g.x0 + (int)lx)];
      } else {
        const int w = levsrc[li];
        level = w == 255 ? -1 : w;
      }
    }
    CHIPS_BOUNDS(level <= CHIPS_MAX_T);
    const unsigned act = __ballot_sync(0xFFFFFFFFu, level >= 0);
    if (level >= 0) {
      const unsigned peers = __match_any_sync(act, level);
      const int leader = __ffs(peers) - 1;
      if (lane == leader) rank = atomicAdd(&scnt[level], (uint32_t)__popc(peers));
      rank = __shfl_sync(peers, rank, leader) + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x <= CHIPS_MAX_T && scnt[threadIdx.x]) sbase[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], scnt[threadIdx.x]);
    __syncthreads();
    if (level >= 0) {
      CHIPS_BOUNDS(sbase[level] + rank < g.npix() && li < g.npix());
      out[sbase[level] + rank] = li;
    }
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
    CHIPS_BOUNDS(i < g.npix() && li < g.npix());
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
    CHIPS_BOUNDS(i < g.npix() && li < g.npix());
    int y = g.y0 + (int)(li / (uint32_t)g.w), x = g.x0 + (int)(li % (uint32_t)g.w);
    size_t o = (size_t)y * g.W + x;
    if (Wimg[o] != 0) continue;
    if (uf_find(parent, li + 1) == kBorder) Wimg[o] = (uint8_t)(t + 1);
  }
}

// ------------------------------------------------------------------ phase C: component tree
// The T filled masks {W <= t} are nested, so their 8-connected components form a tree: ONE
// union-find is grown through the levels t = 0 .. T-1 instead of T independent labellings
// (T x less union-find work and one parent array instead of T).  Two small launches per level
// (the launch boundary is the grid barrier; other streams fill the GPU meanwhile):
//   k_tree_union(t)   the pixels with W == t (a list grouped by W) are united with their
//                     8-neighbours that are already foreground; every OLD root that gets hooked
//                     is appended to `hooked` (a new pixel that joins a component is not).
//   k_tree_merge(t)   each node hooked in this level hands its accumulated 2*area to the root
//                     of the merged component and records (mparent, mlevel) — the merge tree,
//                     whose chains have strictly increasing levels, i.e. at most T links; each
//                     new pixel records its root of this level and adds its bit-quad events (a
//                     2x2 quad adds 1 to 2*contourArea when its third pixel appears and 1 more
//                     with the fourth: 2*Q4 + Q3).  The atomicAdd that lifts a root over the
//                     area threshold records klevel[root] = t; component counts are running sums.
//   k_tree_keep       keep(p) = first level at which p's component is kept: walk p's merge
//                     chain to the first node with a finite klevel.
struct Tree {
  uint32_t* parent;    // [nr] union-find (path-halved)
  uint32_t* mparent;   // [nr] root of the merged component at the end of the level of the hook
  uint32_t* area;      // [nr] 2*area accumulated while the node was a root
  uint8_t* mlevel;     // [nr] level in which the node stopped being a root (255: still a root)
  uint8_t* klevel;     // [nr] level at which the node, as a root, reached the threshold (255: never)
  uint8_t* wl;         // [nr] ROI-local copy of W (255 outside every mask)
  uint32_t* plist;     // foreground pixels (ROI-local index) grouped by W, ascending
  uint32_t* hooked;    // hooked old roots in hook order
  uint32_t* offset;    // [T+2] list offsets by W (lvl_offset)
  uint32_t* ctr;       // [1] hook cursor, [2] kept components, [3] components, [8+t] hook cursor after level t
  uint32_t npix;       // ROI pixels = capacity of every array above
};

__device__ __forceinline__ int wl_at(const Roi& g, const uint8_t* __restrict__ wl, int lx, int ly) {
  return (lx >= 0 && lx < g.w && ly >= 0 && ly < g.h) ? (int)wl[(size_t)ly * g.w + lx] : 255;
}

__device__ __forceinline__ void uf_unite_rec(const Tree& tr, int t, uint32_t a, uint32_t b) {
  uint32_t* __restrict__ parent = tr.parent;
  a = uf_find(parent, a);
  b = uf_find(parent, b);
  while (a != b) {
    if (a < b) {
      uint32_t x = a;
      a = b;
      b = x;
    }                            // a > b : hook a under b
    uint32_t old = atomicMin(&parent[a], b);
    if (old == a) {              // a was a root until now: it is hooked exactly once
      if (tr.wl[a] != t) {                                           // an older component merges
        const uint32_t slot = atomicAdd(&tr.ctr[1], 1u);
        CHIPS_BOUNDS(slot < tr.npix && a < tr.npix);
        tr.hooked[slot] = a;
      }
      a = b;
    } else {
      a = old;                   // hooked by somebody else in this level: go on with its parent
    }
  }
}

// C1: ROI-local W, per-level counts of the foreground pixels and singleton initialisation (the
// list itself is regrouped from phase B's by k_list_scatter).  A CTA covers 1024 pixels of a row
// in four 256-pixel chunks whose W bytes are loaded together.
__global__ void __launch_bounds__(256) k_tree_list(Roi g, const uint8_t* __restrict__ Wimg, int T, Tree tr,
                                                   uint32_t thr2, uint32_t* __restrict__ lvl_count) {
  __shared__ uint32_t scnt[CHIPS_MAX_T + 1];
  for (int i = threadIdx.x; i <= T; i += blockDim.x) scnt[i] = 0;
  __syncthreads();
  const int y = g.y0 + blockIdx.y;
  const uint8_t* __restrict__ wrow = Wimg + (size_t)y * g.W;
  const int lane = threadIdx.x & 31;
  int vv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int x = g.x0 + (4 * (int)blockIdx.x + c) * 256 + (int)threadIdx.x;
    vv[c] = x < g.x0 + g.w ? (int)wrow[x] : 255;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int x = g.x0 + (4 * (int)blockIdx.x + c) * 256 + (int)threadIdx.x;
    if (x - (int)threadIdx.x >= g.x0 + g.w) break;            // the whole chunk lies right of the ROI
    const int v = vv[c];
    const int w = (x < g.x0 + g.w && v < T) ? v : -1;
    if (x < g.x0 + g.w) tr.wl[g.li(x, y)] = (uint8_t)(v < T ? v : 255);
    const unsigned act = __ballot_sync(0xFFFFFFFFu, w >= 0);
    if (w >= 0) {
      const uint32_t me = g.li(x, y);
      const unsigned peers = __match_any_sync(act, w);
      const int leader = __ffs(peers) - 1;
      const int lane_x0 = x - lane;                  // image column of lane 0 of this warp
      // pixels of one row run that appear at the same level belong to one component from that
      // level on: link each to the start of its run inside the warp's 32-pixel span (a run that
      // continues across the span boundary points at the last pixel of the previous span)
      const unsigned z = ~peers & ((1u << lane) - 1u);
      const int s = z ? (32 - __clz(z)) : 0;
      uint32_t tgt = me - (uint32_t)(lane - s);
      if (s == 0 && lane_x0 > g.x0 && wrow[lane_x0 - 1] == w) tgt = me - (uint32_t)lane - 1u;
      tr.parent[me] = tgt;
      tr.area[me] = 0;
      tr.mlevel[me] = 255;
      tr.klevel[me] = thr2 == 0 ? (uint8_t)w : 255;   // threshold <= 0: kept from birth
      if (lane == leader) atomicAdd(&scnt[w], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= T; i += blockDim.x)
    if (scnt[i]) atomicAdd(&lvl_count[i], scnt[i]);
}

__device__ __forceinline__ void tree_stats(const Tree& tr, int t, uint32_t thr2, chips_result* res) {
  const int ncomp = (int)tr.ctr[3];
  res->n_comp[t] = ncomp;
  res->n_kept[t] = thr2 == 0 ? ncomp : (int32_t)tr.ctr[2];
}

// C2a: unions of the pixels that appear at level t
__global__ void __launch_bounds__(256) k_tree_union(Roi g, int t, Tree tr, uint32_t thr2,
                                                    chips_result* __restrict__ res) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid == 0 && t > 0) tree_stats(tr, t - 1, thr2, res);      // counters of the finished level
  const uint32_t beg = tr.offset[t], end = tr.offset[t + 1];
  const uint8_t* __restrict__ wl = tr.wl;
  for (uint32_t i = beg + gtid; i < end; i += nthreads) {
    const uint32_t me = tr.plist[i];
    CHIPS_BOUNDS(i < tr.npix && me < tr.npix);
    const int ly = (int)(me / (uint32_t)g.w), lx = (int)(me - (uint32_t)ly * (uint32_t)g.w);
    const int w_l = wl_at(g, wl, lx - 1, ly), w_r = wl_at(g, wl, lx + 1, ly);
    const int w_u = wl_at(g, wl, lx, ly - 1), w_d = wl_at(g, wl, lx, ly + 1);
    const int w_ul = wl_at(g, wl, lx - 1, ly - 1), w_ur = wl_at(g, wl, lx + 1, ly - 1);
    const int w_dl = wl_at(g, wl, lx - 1, ly + 1), w_dr = wl_at(g, wl, lx + 1, ly + 1);
    // Smaller raster index: any foreground neighbour; larger index: strictly older ones only (a
    // same-level neighbour with a larger index unites with us from its side).  A same-level left
    // neighbour is already linked (row runs); a vertical edge is implied when the horizontal
    // neighbour and the diagonal next to it are both foreground (they are connected to each
    // other and to us); a diagonal is implied by the vertical neighbour between.
    if (w_l < t) uf_unite_rec(tr, t, me, me - 1u);
    if (w_u <= t) {
      if (!(w_l <= t && w_ul <= t)) uf_unite_rec(tr, t, me, me - (uint32_t)g.w);
    } else {
      if (w_ul <= t && !(w_l <= t)) uf_unite_rec(tr, t, me, me - (uint32_t)g.w - 1u);
      if (w_ur <= t && !(w_r < t)) uf_unite_rec(tr, t, me, me - (uint32_t)g.w + 1u);
    }
    if (w_r < t) uf_unite_rec(tr, t, me, me + 1u);
    if (w_d < t) {
      if (!(w_r < t && w_dr < t)) uf_unite_rec(tr, t, me, me + (uint32_t)g.w);
    } else {
      if (w_dl < t && !(w_l < t)) uf_unite_rec(tr, t, me, me + (uint32_t)g.w - 1u);
      if (w_dr < t && !(w_r < t)) uf_unite_rec(tr, t, me, me + (uint32_t)g.w + 1u);
    }
  }
}

// C2b: area hand-over of the hooked roots, merge-tree edges, bit-quad events of the new pixels
__global__ void __launch_bounds__(256) k_tree_merge(Roi g, int t, Tree tr, uint32_t thr2) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t hook_begin = t > 0 ? tr.ctr[8 + t - 1] : 0u, hook_end = tr.ctr[1];
  if (gtid == 0) tr.ctr[8 + t] = hook_end;
  int dk = 0, dc = 0;                             // change of the kept / total component counts
  for (uint32_t i = hook_begin + gtid; i < hook_end; i += nthreads) {
    const uint32_t a = tr.hooked[i];
    CHIPS_BOUNDS(i < tr.npix && a < tr.npix);
    const uint32_t r = uf_find(tr.parent, a);
    tr.mparent[a] = r;
    tr.mlevel[a] = (uint8_t)t;
    --dc;
    const uint32_t aa = tr.area[a];
    if (aa >= thr2) --dk;                         // a kept component stops being a component
    if (aa) {
      const uint32_t old = atomicAdd(&tr.area[r], aa);
      if (old < thr2 && old + aa >= thr2) {
        tr.klevel[r] = (uint8_t)t;
        ++dk;
      }
    }
  }
  const uint32_t beg = tr.offset[t], end = tr.offset[t + 1];
  const uint8_t* __restrict__ wl = tr.wl;
  for (uint32_t i = beg + gtid; i < end; i += nthreads) {
    const uint32_t me = tr.plist[i];
    CHIPS_BOUNDS(i < tr.npix && me < tr.npix);
    const int ly = (int)(me / (uint32_t)g.w), lx = (int)(me - (uint32_t)ly * (uint32_t)g.w);
    int wn[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) wn[dy + 1][dx + 1] = (dx | dy) ? wl_at(g, wl, lx + dx, ly + dy) : t;
    uint32_t add = 0;
    // the four quads that contain p; inside a quad positions are ordered TL, TR, BL, BR
#pragma unroll
    for (int qy = 0; qy < 2; ++qy)
#pragma unroll
      for (int qx = 0; qx < 2; ++qx) {
        const int pme = (1 - qy) * 2 + (1 - qx);           // p's position inside this quad
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j == pme) continue;
          const int wq = wn[qy + (j >> 1)][qx + (j & 1)];
          rank += (wq < t || (wq == t && j < pme)) ? 1 : 0;
        }
        add += (rank >= 2) ? 1u : 0u;
      }
    const uint32_t r = uf_find(tr.parent, me);
    if (r != me) {                                         // joined a component in this level
      tr.parent[me] = r;                                   // flat: the next level's finds start here
      tr.mparent[me] = r;
      tr.mlevel[me] = (uint8_t)t;
    } else {
      ++dc;                                                // a new component
    }
    // one atomic per root and warp (the pixels of a large component all add to one address)
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, add ? r : 0xFFFFFFFFu);
    if (add) {
      const uint32_t sum = __reduce_add_sync(peers, add);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) {
        const uint32_t old = atomicAdd(&tr.area[r], sum);
        if (old < thr2 && old + sum >= thr2) {
          tr.klevel[r] = (uint8_t)t;
          ++dk;
        }
      }
    }
  }
  dk = __reduce_add_sync(0xFFFFFFFFu, dk);
  dc = __reduce_add_sync(0xFFFFFFFFu, dc);
  if ((threadIdx.x & 31) == 0) {
    if (dk) atomicAdd(&tr.ctr[2], (uint32_t)dk);
    if (dc) atomicAdd(&tr.ctr[3], (uint32_t)dc);
  }
}

// C3: keep(p) = first level at which p's component is kept
__global__ void __launch_bounds__(256) k_tree_keep(Roi g, int T, Tree tr, uint32_t thr2,
                                                   chips_result* __restrict__ res,
                                                   uint8_t* __restrict__ keep) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid == 0) tree_stats(tr, T - 1, thr2, res);
  const uint32_t total = tr.offset[T];
  for (uint32_t i = gtid; i < total; i += nthreads) {
    const uint32_t me = tr.plist[i];
    CHIPS_BOUNDS(i < tr.npix && me < tr.npix);
    const int ly = (int)(me / (uint32_t)g.w), lx = (int)(me - (uint32_t)ly * (uint32_t)g.w);
    int tjoin = tr.wl[me], kk = T;
    uint32_t n = me;
    for (int hop = 0; hop <= T; ++hop) {
      const int kl = tr.klevel[n];
      if (kl != 255) {
        kk = max(kl, tjoin);
        break;
      }
      const int ml = tr.mlevel[n];
      if (ml == 255) break;                       // a final root that never reached the threshold
      tjoin = ml;
      n = tr.mparent[n];
    }
    keep[(size_t)(g.y0 + ly) * g.W + (g.x0 + lx)] = (uint8_t)kk;
  }
}

// chips.py:424-426: cv2.contourArea(cnt) / self.disk_area >= self.area_threshold
static inline bool area_kept(uint32_t a2, double disk_area, double thr) {
  return ((double)a2 * 0.5) / disk_area >= thr;
}

__global__ void k_cleanup_reset(chips_result* res, uint32_t* lvl_count, uint32_t* ctr) {
  int t = threadIdx.x;
  if (t < CHIPS_MAX_T) {
    res->n_kept[t] = 0;
    res->n_comp[t] = 0;
  }
  if (t <= CHIPS_MAX_T + 1) lvl_count[t] = 0;
  if (t < 8 + CHIPS_MAX_T) ctr[t] = 0;
}

__global__ void k_zero_counts(uint32_t* lvl_count) {
  if (threadIdx.x <= CHIPS_MAX_T + 1) lvl_count[threadIdx.x] = 0;
}

// ------------------------------------------------------------------ expansion / debugging
__global__ void __launch_bounds__(256) k_expand_all(const uint8_t* __restrict__ src, size_t n, int T,
                                                    uint8_t* __restrict__ maps) {
  // 16 pixels per thread: one 128-bit load, one 128-bit store per plane; map_t = (keep <= t) for
  // four bytes at a time (SIMD byte compare)
  size_t i16 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i16 >= n) return;
  if (i16 + 16 <= n && (n & 15) == 0 && (reinterpret_cast<uintptr_t>(maps) & 15) == 0) {     // aligned 128-bit stores in every plane
    const uint4 v = *reinterpret_cast<const uint4*>(src + i16);
    for (int t = 0; t < T; ++t) {
      const uint32_t t4 = (uint32_t)t * 0x01010101u;
      uint4 o;
      o.x = __vcmpleu4(v.x, t4) & 0x01010101u;
      o.y = __vcmpleu4(v.y, t4) & 0x01010101u;
      o.z = __vcmpleu4(v.z, t4) & 0x01010101u;
      o.w = __vcmpleu4(v.w, t4) & 0x01010101u;
      *reinterpret_cast<uint4*>(maps + (size_t)t * n + i16) = o;
    }
  } else {
    for (size_t i = i16; i < n && i < i16 + 16; ++i)
      for (int t = 0; t < T; ++t) maps[(size_t)t * n + i] = src[i] <= t;
  }
}

template <typename O>
__global__ void __launch_bounds__(256) k_expand_one(const uint8_t* __restrict__ src, size_t n, int t,
                                                    O* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (src[i] <= t) ? (O)1 : (O)0;
}

// Root image of level t from the merge tree (chains have increasing levels: stop at the first
// node that was still a root at level t) and the component's 2*area at that level, re-counted
// from the bit quads into `acc` (the dead union-find array, zeroed by the caller); each quad is
// charged to its first set pixel in raster order.
__global__ void __launch_bounds__(256) k_roots(Roi g, int t, Tree tr, uint32_t* __restrict__ acc,
                                               int32_t* __restrict__ roots) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= g.W) return;
  size_t o = (size_t)y * g.W + x;
  int32_t rr = -1;
  const int lx = x - g.x0, ly = y - g.y0;
  if (wl_at(g, tr.wl, lx, ly) <= t) {
    uint32_t n = (uint32_t)ly * (uint32_t)g.w + (uint32_t)lx;
    while (tr.mlevel[n] <= t) n = tr.mparent[n];
    rr = (int32_t)n;
    const int wr = wl_at(g, tr.wl, lx + 1, ly), wd = wl_at(g, tr.wl, lx, ly + 1);
    const int wdr = wl_at(g, tr.wl, lx + 1, ly + 1), wl_ = wl_at(g, tr.wl, lx - 1, ly);
    const int wdl = wl_at(g, tr.wl, lx - 1, ly + 1);
    const int sa = 1 + (int)(wr <= t) + (int)(wd <= t) + (int)(wdr <= t);
    uint32_t add = (sa == 4) ? 2u : (sa == 3 ? 1u : 0u);
    if (wd <= t && wl_ > t && wdl <= t) add += 1u;
    if (add) atomicAdd(&acc[n], add);
  }
  roots[o] = rr;
}

__global__ void __launch_bounds__(256) k_roots_area(const int32_t* __restrict__ roots,
                                                    const uint32_t* __restrict__ acc, size_t n,
                                                    uint32_t* __restrict__ area_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) area_out[i] = roots[i] >= 0 ? acc[roots[i]] : 0u;
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
  // 16 level bytes per thread and step (one 128-bit load); in accumulate mode only the pixels
  // of some map (~10 % of the disk) touch the sum image
  for (size_t i16 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i16 < n;
       i16 += (size_t)gridDim.x * blockDim.x * 16) {
    uint8_t kb[16];
    if (i16 + 16 <= n) {
      *reinterpret_cast<uint4*>(kb) = *reinterpret_cast<const uint4*>(keep + i16);
    } else {
      for (int j = 0; j < 16; ++j) kb[j] = i16 + j < n ? keep[i16 + j] : (uint8_t)T;
    }
    if (accumulate) {
      // running sum for the batch uncertainty map: "no detection" counts as 0
      if (poisoned) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int kk = kb[j];
        if (kk < T && i16 + j < n) {
          const double v = suffix[kk];
          if (v != 0.0) out[i16 + j] += (O)v;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (i16 + j >= n) break;
        const int kk = kb[j];
        const double v = poisoned ? NAN : suffix[kk > T ? T : kk];
        out[i16 + j] = (O)((v == 0.0) ? NAN : v);
      }
    }
  }
}

static Roi make_roi(const chips_plan* p, int cx, int cy, int r) {
  Roi g;
  g.x0 = p->roi_x0; g.y0 = p->roi_y0; g.w = p->roi_w; g.h = p->roi_h;
  g.W = p->W; g.H = p->H; g.cx = cx; g.cy = cy; g.r = r;
  return g;
}

static Tree make_tree(const chips_plan* p, unsigned char* w) {
  const size_t nr = (size_t)p->roi_w * p->roi_h;
  Tree t;
  t.parent = reinterpret_cast<uint32_t*>(w + p->off_parent_c);
  t.mparent = t.parent + nr;
  t.area = reinterpret_cast<uint32_t*>(w + p->off_area);
  t.mlevel = reinterpret_cast<uint8_t*>(t.area + nr);
  t.klevel = t.mlevel + nr;
  t.wl = t.klevel + nr;
  t.plist = reinterpret_cast<uint32_t*>(w + p->off_list2);
  t.hooked = reinterpret_cast<uint32_t*>(w + p->off_parent_b);
  uint32_t* lvl = reinterpret_cast<uint32_t*>(w + p->off_lvl);
  t.offset = lvl + 128;
  t.ctr = lvl + 384;
  t.npix = (uint32_t)nr;
  return t;
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
  uint32_t* pending = reinterpret_cast<uint32_t*>(w + plan->off_pending);
  uint32_t* lvl = reinterpret_cast<uint32_t*>(w + plan->off_lvl);       // count | offset | cursor
  uint32_t* lvl_count = lvl, *lvl_offset = lvl + 128, *lvl_cursor = lvl + 256;
  uint32_t* ctr = lvl + 384;                   // phase C counters
  chips_result* res = reinterpret_cast<chips_result*>(w + plan->off_result);
  Roi g = make_roi(plan, cx, cy, r);
  dim3 blk(256);
  dim3 grid((g.w + 255) / 256, g.h);

  CHIPS_CUDA(cudaMemsetAsync(Wimg, T, n_img, st));
  CHIPS_CUDA(cudaMemsetAsync(keep, T, n_img, st));
  k_cleanup_reset<<<1, 128, 0, st>>>(res, lvl_count, ctr);
  // round T-1: the bulk background (pixels in no mask at all)
  k_fill_bulk_link<<<g.h, blk, 0, st>>>(g, lev, T, parent_b);
  dim3 grid4((g.w / 4 + 2 + 255) / 256, g.h);
  k_fill_bulk_union<<<grid4, blk, 0, st>>>(g, lev, T, parent_b);
  uint32_t* list2 = reinterpret_cast<uint32_t*>(w + plan->off_list2);
  k_fill_pending<<<grid4, blk, 0, st>>>(g, lev, T, Wimg, parent_b, lvl_count, &ctr[4], list2);
  k_fill_offsets<<<1, 32, 0, st>>>(T, lvl_count, lvl_offset, lvl_cursor);
  k_list_scatter<<<8 * kNumSM, kScatterThreads, 0, st>>>(g, lev, 0, list2, &ctr[4], lvl_cursor, pending);
  CHIPS_CHECK_LAUNCH();
  // the bulk pixels that did not reach the border in round T-1 stay pending at level T and
  // are re-examined by every later round (suffix of the list)
  const int gl = kNumSM;      // (2 x SMs: 4 % shorter alone, 1.5 % less throughput with 16 detections in flight)
  for (int t = T - 2; t >= 0; --t) {
    k_fill_round_union<<<gl, blk, 0, st>>>(g, lev, t, lvl_offset, pending, parent_b);
    k_fill_round_assign<<<gl, blk, 0, st>>>(g, t, T, lvl_offset, pending, Wimg, parent_b);
  }
  CHIPS_CHECK_LAUNCH();
  // ---- phase C: one union-find grown through the levels (component tree)
  Tree tr = make_tree(plan, w);
  uint32_t thr2 = 0xFFFFFFFFu;                 // smallest 2*area that passes chips.py:424-426
  {
    const double guess = 2.0 * area_threshold * disk_area;
    if (guess <= 0.0) {
      if (area_kept(0u, disk_area, area_threshold)) thr2 = 0;
    } else if (guess < 4.0e9) {
      long long a2 = (long long)guess;
      while (a2 > 0 && area_kept((uint32_t)(a2 - 1), disk_area, area_threshold)) --a2;
      for (int i = 0; i < 8 && !area_kept((uint32_t)a2, disk_area, area_threshold); ++i) ++a2;
      if (area_kept((uint32_t)a2, disk_area, area_threshold)) thr2 = (uint32_t)a2;
    }                                          // NaN / huge threshold: nothing is ever kept
  }
  k_zero_counts<<<1, 128, 0, st>>>(lvl_count);
  k_tree_list<<<dim3((g.w + 1023) / 1024, g.h), blk, 0, st>>>(g, Wimg, T, tr, thr2, lvl_count);
  k_fill_offsets<<<1, 32, 0, st>>>(T, lvl_count, lvl_offset, lvl_cursor);
  // the foreground pixels are exactly phase B's pending pixels: regroup that list by W
  k_list_scatter<<<8 * kNumSM, kScatterThreads, 0, st>>>(g, tr.wl, 1, pending, &ctr[4], lvl_cursor, list2);
  CHIPS_CHECK_LAUNCH();
  for (int t = 0; t < T; ++t) {
    k_tree_union<<<gl, blk, 0, st>>>(g, t, tr, thr2, res);
    k_tree_merge<<<gl, blk, 0, st>>>(g, t, tr, thr2);
  }
  k_tree_keep<<<16 * kNumSM, blk, 0, st>>>(g, T, tr, thr2, res, keep);
  CHIPS_CHECK_LAUNCH();
  count_launch(6 + 2 * (T - 1) + 5 + 2 * T);
  return CHIPS_OK;
}

static const uint8_t* pick_src(const chips_plan* plan, int which, unsigned char* w) {
  return w + (which == 1 ? plan->off_level : (which == 2 ? plan->off_fill : plan->off_keep));
}

extern "C" int chips_expand_maps(const chips_plan* plan, int which, uint8_t* maps, void* ws,
                                 void* stream) {
  if (!plan || !maps || !ws) return CHIPS_E_ARG;
  size_t n = (size_t)plan->H * plan->W;
  unsigned grid = (unsigned)((n / 16 + 256) / 256);
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
  Tree tr = make_tree(plan, w);
  const size_t nr = (size_t)plan->roi_w * plan->roi_h, npix = (size_t)plan->H * plan->W;
  cudaStream_t st = (cudaStream_t)stream;
  CHIPS_CUDA(cudaMemsetAsync(tr.parent, 0, nr * sizeof(uint32_t), st));   // dead after chips_cleanup
  dim3 grid((plan->W + 255) / 256, plan->H);
  k_roots<<<grid, 256, 0, st>>>(g, t, tr, tr.parent, roots);
  if (twice_area)
    k_roots_area<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(roots, tr.parent, npix, twice_area);
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
