// contours_core.h — border following (cv2.findContours RETR_TREE) as data-parallel passes.
//
// Replaces: cv2.findContours(tmp_data.astype(np.uint8), RETR_TREE, CHAIN_APPROX_SIMPLE),
// chips/chips.py:382-386.  Formulation pinned against OpenCV by
// oracle/contours_parallel_spec.py and tests/test_contours_spec.py:
//   * a state (p, s) = "at border pixel p, arrived from its neighbour in direction s"; one
//     border-following step is a permutation of the states, so borders are its cycles;
//   * whether a visit marks p negative / keeps p under CHAIN_APPROX_SIMPLE is local to a state;
//   * the raster scan only decides WHERE each cycle starts: the first candidate pixel (west or
//     east neighbour 0) whose mark condition holds given the starts of the other cycles through
//     that pixel — a fixed point reached in a handful of Jacobi rounds (O(image side) rounds on
//     lattices of 1-pixel walls: the call reports `changed`, the caller re-enqueues with more);
//   * border number = rank of the start position, parent from the nearest earlier mark in the row.
//
// Every pass is a functor applied to an index range.  contours.cu runs them as CUDA kernels;
// tests/emu/contours_emu.cpp runs the same functors in host loops so the logic is checked on
// the CPU-only build box (test infrastructure, never loaded by the product).
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CTR_HD __host__ __device__ __forceinline__
#else
#define CTR_HD inline
#endif

#ifdef __CUDA_ARCH__
#define CTR_POPC(v) __popc(v)
#define CTR_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define CTR_ATOMIC_ADD64(p, v) atomicAdd((unsigned long long*)(p), (unsigned long long)(v))
#else
#define CTR_POPC(v) __builtin_popcount(v)
#define CTR_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define CTR_ATOMIC_ADD64(p, v) (*(p) += (v))
#endif

namespace ctr {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kScanTile = 4096;      // elements per CTA of the scan kernels

// counters in Ws::ctr
enum { C_NB = 0, C_NBORDERS = 2, C_NPOINTS = 3, C_OVERFLOW = 4, C_NB_RAW = 5,
       C_LABEL = 16,      // [32] "round r changed a label"  (a round with no change => converged, both
       C_START = 80,      // [3] "round r moved a start", slot r % 3           buffers equal, later rounds return)
       C_COUNT = 256 };
constexpr int kMaxStartRounds = 1 << 20;   // lattices of 1-pixel walls need O(image side) rounds

// directions: 0 = east, then counter-clockwise on the screen (y grows downwards)
CTR_HD int dx(int s) { return (int)((0x901Au >> (2 * s)) & 3u) - 1; }
CTR_HD int dy(int s) { return (int)((0xA901u >> (2 * s)) & 3u) - 1; }

struct Geo {
  const uint8_t* img;
  int w, h, pitch, gw;   // gw = groups of 8 pixels per row
  int mode, thr;         // mode 0: foreground = img <= thr ; mode 1: foreground = img != 0
};

CTR_HD bool fg(const Geo& g, int x, int y) {
  if ((unsigned)x >= (unsigned)g.w || (unsigned)y >= (unsigned)g.h) return false;
  const uint8_t v = g.img[(size_t)y * g.pitch + x];
  return g.mode ? v != 0 : (int)v <= g.thr;
}

struct Rec {              // == chips_contour_rec
  int32_t x, y, is_hole, parent, n_points, offset;
  int64_t area2;
};

struct Counts {           // == chips_contour_counts
  int32_t n_border_px, n_contours, n_points, changed, overflow, pad_;
};

struct Ws {
  int32_t* ctr;
  uint8_t* gmask;   uint32_t* goff;   uint32_t* s1;
  uint32_t* px;     uint8_t* nbm;
  uint32_t* nxt;    uint8_t* sfl;
  uint32_t *labA, *labB, *jmpA, *jmpB, *cntA, *cntB, *ecA, *ecB;
  long long* area;
  uint32_t *stOld, *stNew, *nbd;
  uint8_t* sflag;   uint32_t* srank;  uint32_t* cyc_at;
  int32_t *btype, *blnbd, *bparent, *bnpts, *bbi;
  uint32_t *boff, *bss;
  long long* barea;
};

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// carve the workspace; returns the bytes used (base may be null to only size it)
inline size_t carve(Ws& ws, char* base, int w, int h, long cap) {
  const size_t G = (size_t)((w + 7) / 8) * h;
  const size_t S = (size_t)cap * 8;
  const size_t nmax = G > (size_t)cap ? G : (size_t)cap;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  ws.ctr = (int32_t*)take(C_COUNT * 4);
  ws.gmask = (uint8_t*)take(G);
  ws.goff = (uint32_t*)take(G * 4);
  ws.s1 = (uint32_t*)take((nmax / kScanTile + 2) * 4);
  ws.px = (uint32_t*)take((size_t)cap * 4);
  ws.nbm = (uint8_t*)take((size_t)cap);
  ws.nxt = (uint32_t*)take(S * 4);
  ws.sfl = (uint8_t*)take(S);
  uint32_t** s32[] = {&ws.labA, &ws.labB, &ws.jmpA, &ws.jmpB, &ws.cntA, &ws.cntB, &ws.ecA, &ws.ecB,
                      &ws.stOld, &ws.stNew, &ws.nbd};
  for (auto p : s32) *p = (uint32_t*)take(S * 4);
  ws.area = (long long*)take(S * 8);
  ws.sflag = (uint8_t*)take((size_t)cap);
  ws.srank = (uint32_t*)take((size_t)cap * 4);
  ws.cyc_at = (uint32_t*)take((size_t)cap * 4);
  int32_t** b32[] = {&ws.btype, &ws.blnbd, &ws.bparent, &ws.bnpts, &ws.bbi};
  for (auto p : b32) *p = (int32_t*)take((size_t)cap * 4);
  ws.boff = (uint32_t*)take((size_t)cap * 4);
  ws.bss = (uint32_t*)take((size_t)cap * 4);
  ws.barea = (long long*)take((size_t)cap * 8);
  return off;
}

// ------------------------------------------------------------------------------------------
// generic passes
struct Fill32 {
  uint32_t* p; uint32_t v;
  CTR_HD void operator()(long i) const { p[i] = v; }
};
// per-state / per-border-pixel tables of the start resolution, in one pass over the states
struct InitCycles {
  const int32_t* ctr; long long* area; uint32_t *stOld, *stNew, *nbd; uint8_t* sflag;
  CTR_HD void operator()(long i) const {
    area[i] = 0; stOld[i] = kNone; stNew[i] = kNone; nbd[i] = 0u;
    if (i < ctr[C_NB]) sflag[i] = 0;
  }
};

// counts fed to Exec::scan (exclusive prefix sums; the CUDA executor has tiled scan kernels for it)
struct CountGroups { const uint8_t* g; CTR_HD uint32_t operator()(long i) const { return (uint32_t)CTR_POPC((unsigned)g[i]); } };
struct CountFlags  { const uint8_t* f; CTR_HD uint32_t operator()(long i) const { return f[i]; } };
struct CountI32    { const int32_t* v; CTR_HD uint32_t operator()(long i) const { return (uint32_t)v[i]; } };

struct ScanTotal {          // what a scan does with its grand total
  int32_t* ctr; int slot, raw_slot, overflow_bit; long limit;
  CTR_HD void operator()(uint32_t run) const {
    if (raw_slot >= 0) ctr[raw_slot] = (int32_t)run;
    if (limit >= 0 && (long)run > limit) { ctr[C_OVERFLOW] |= overflow_bit; run = 0; }   // clamp: nothing downstream runs
    ctr[slot] = (int32_t)run;
  }
};

// ------------------------------------------------------------------------------------------
// border pixels: foreground with a background pixel among the 8 neighbours
CTR_HD bool fg_value(const Geo& g, unsigned v) { return g.mode ? v != 0 : (int)v <= g.thr; }

// 10 foreground bits of row y for x0-1 .. x0+8; the 8 pixels of the group come from one aligned
// 64-bit load when the row allows it (`words`: image base and pitch are multiples of 8)
CTR_HD unsigned row_window(const Geo& g, bool words, int x0, int y) {
  if ((unsigned)y >= (unsigned)g.h) return 0u;
  unsigned b = 0;
  if (words && x0 + 8 <= g.w) {
    const uint8_t* p = g.img + (size_t)y * g.pitch + x0;
#ifndef __CUDA_ARCH__
    if ((uintptr_t)p & 7) __builtin_trap();          // the host emulation checks the alignment rule
#endif
    const uint64_t v = *reinterpret_cast<const uint64_t*>(p);
    for (int k = 0; k < 8; ++k) b |= (unsigned)fg_value(g, (unsigned)(v >> (8 * k)) & 0xFFu) << (k + 1);
    b |= (unsigned)fg(g, x0 - 1, y);
    b |= (unsigned)fg(g, x0 + 8, y) << 9;
    return b;
  }
  for (int k = 0; k < 10; ++k) b |= (unsigned)fg(g, x0 - 1 + k, y) << k;
  return b;
}

struct FlagGroups {
  Geo g; uint8_t* gmask; bool words;
  CTR_HD void operator()(long i) const {
    const int y = (int)(i / g.gw), x0 = (int)(i % g.gw) * 8;
    unsigned inner = 0xFFu, mid = 0;
    for (int r = 0; r < 3; ++r) {
      const unsigned row = row_window(g, words, x0, y - 1 + r);
      if (r == 1) mid = row;
      inner &= row & (row >> 1) & (row >> 2);
    }
    gmask[i] = (uint8_t)((mid >> 1) & 0xFFu & ~inner);
  }
};

CTR_HD bool is_border(const Geo& g, const uint8_t* gmask, int x, int y) {
  if ((unsigned)x >= (unsigned)g.w || (unsigned)y >= (unsigned)g.h) return false;
  return (gmask[(size_t)y * g.gw + (x >> 3)] >> (x & 7)) & 1;
}
CTR_HD uint32_t border_index(const Geo& g, const uint8_t* gmask, const uint32_t* goff, int x, int y) {
  const size_t gi = (size_t)y * g.gw + (x >> 3);
  return goff[gi] + (uint32_t)CTR_POPC((unsigned)gmask[gi] & ((1u << (x & 7)) - 1u));
}

struct ScatterBorder {
  Geo g; const uint8_t* gmask; const uint32_t* goff; const int32_t* ctr; uint32_t* px; uint8_t* nbm;
  CTR_HD void operator()(long i) const {
    unsigned m = gmask[i];
    if (!m) return;
    const uint32_t nB = (uint32_t)ctr[C_NB];
    const int y = (int)(i / g.gw), x0 = (int)(i % g.gw) * 8;
    uint32_t bi = goff[i];
    for (int k = 0; k < 8; ++k) {
      if (!((m >> k) & 1)) continue;
      if (bi < nB) {
        const int x = x0 + k;
        unsigned nb = 0;
        for (int s = 0; s < 8; ++s) nb |= (unsigned)fg(g, x + dx(s), y + dy(s)) << s;
        px[bi] = (uint32_t)y * (uint32_t)g.w + (uint32_t)x;
        nbm[bi] = (uint8_t)nb;
      }
      ++bi;
    }
  }
};

// state flags
enum { F_VALID = 1, F_NEG = 2, F_EMIT = 4 };

struct InitStates {
  Geo g; const uint8_t* gmask; const uint32_t* goff; const int32_t* ctr;
  const uint32_t* px; const uint8_t* nbm;
  uint32_t* nxt; uint8_t* sfl; uint32_t* lab; uint32_t* jmp; uint32_t *dist, *em, *ew;
  CTR_HD void operator()(long i) const {
    if (i >= 8L * ctr[C_NB]) return;
    const uint32_t bi = (uint32_t)(i >> 3);
    const int s = (int)(i & 7);
    const unsigned m = nbm[bi];
    uint32_t to = (uint32_t)i, label = kNone;
    uint8_t fl = 0;
    if ((m >> s) & 1) {
      const int x = (int)(px[bi] % (uint32_t)g.w), y = (int)(px[bi] / (uint32_t)g.w);
      for (int t = 1; t <= 8; ++t) {
        const int s2 = (s + t) & 7;
        if (!((m >> s2) & 1)) continue;
        const int x2 = x + dx(s2), y2 = y + dy(s2);
        if (is_border(g, gmask, x2, y2)) {
          to = border_index(g, gmask, goff, x2, y2) * 8u + (uint32_t)((s2 + 4) & 7);
          label = (uint32_t)i;
          fl = F_VALID;
          if (s != 0 && s + t >= 9) fl |= F_NEG;          // the east neighbour was examined as 0
          if (s2 != ((s + 4) & 7)) fl |= F_EMIT;          // direction changes here
        }
        break;
      }
    }
    nxt[i] = to; sfl[i] = fl; lab[i] = label; jmp[i] = to;
    dist[i] = 0; em[i] = 0; ew[i] = (fl & F_EMIT) ? 1u : 0u;
  }
};

// Pointer doubling over the window W_r(i) = {i, T(i), .., T^(2^r - 1)(i)}; per state
//   lab  = smallest state id in the window (kNone absorbs: the walk leaves the border pixels),
//   dist = steps from i to that state, em = kept points among those steps, ew = kept points in W_r(i).
// Once the window covers the cycle: lab = the cycle's label m, dist / em = list rank of i against m.
// A round that changes no label has converged (windows at stride 2^r tile the cycle, so equal
// minima along every stride orbit are the global minimum); both ping-pong buffers then hold the
// same lab / dist / em and the remaining rounds return at once.
struct LabelRound {
  int32_t* ctr; int r; const uint32_t *lab, *jmp, *dist, *em, *ew; uint32_t *lab2, *jmp2, *dist2, *em2, *ew2;
  CTR_HD bool operator()(long i) const {          // true: raise ctr[C_LABEL + r] (Exec::each_flag)
    if (i >= 8L * ctr[C_NB] || (r > 0 && !ctr[C_LABEL + r - 1])) return false;
    const uint32_t a = lab[i], j = jmp[i], b = lab[j];
    uint32_t v = a, d = dist[i], e = em[i];
    const bool better = a != kNone && (b == kNone || b < a);
    if (better) { v = b; d = (1u << r) + dist[j]; e = ew[i] + em[j]; }
    lab2[i] = v; dist2[i] = d; em2[i] = e;
    ew2[i] = ew[i] + ew[j];
    jmp2[i] = jmp[j];
    return better;
  }
};

// shoelace sum of a cycle: 2 * signed area of the polygon through the pixel centres
struct AreaAccum {
  Geo g; const int32_t* ctr; const uint32_t *lab, *nxt, *px; long long* area;
  CTR_HD void operator()(long i) const {
    if (i >= 8L * ctr[C_NB]) return;
    const uint32_t l = lab[i];
    if (l == kNone) return;
    const uint32_t p = px[i >> 3], q = px[nxt[i] >> 3];
    const long long x = p % (uint32_t)g.w, y = p / (uint32_t)g.w;
    const long long x2 = q % (uint32_t)g.w, y2 = q / (uint32_t)g.w;
    const long long v = x * y2 - x2 * y;
    if (v) CTR_ATOMIC_ADD64(&area[l], v);
  }
};

// first neighbour clockwise from direction s0 (exclusive) — the start state of a new border
CTR_HD int start_dir(unsigned m, int s0) {
  int s = s0;
  for (int k = 0; k < 8; ++k) {
    s = (s - 1) & 7;
    if ((m >> s) & 1) return s;
  }
  return -1;
}

// one Jacobi round of "where does every cycle start": start key = 2 * position + is_hole
struct StartRound {
  int32_t* ctr; int r; const uint32_t* px; const uint8_t* nbm; const uint32_t* lab;
  const uint8_t* sfl; const uint32_t* stOld; uint32_t* stNew;
  CTR_HD void operator()(long bi) const {
    if (bi == 0) ctr[C_START + (r + 1) % 3] = 0;      // the slot of the next round (nobody reads it now)
    if (bi >= ctr[C_NB] || (r > 0 && !ctr[C_START + (r - 1) % 3])) return;
    const unsigned m = nbm[bi];
    const bool w0 = !((m >> 4) & 1), e0 = !(m & 1);
    if (!w0 && !e0) return;
    const uint32_t pos = px[bi];
    if (m == 0) { CTR_ATOMIC_MIN(&stNew[bi * 8], pos * 2u); return; }   // isolated pixel
    bool earlier_any = false, earlier_neg = false;
    for (int s = 0; s < 8; ++s) {
      if (!((m >> s) & 1)) continue;
      const uint32_t c = lab[bi * 8 + s];
      if (c == kNone) continue;
      const uint32_t st = stOld[c];
      if (st == kNone || (st >> 1) >= pos) continue;
      earlier_any = true;
      for (int s2 = 0; s2 < 8; ++s2)
        if (((m >> s2) & 1) && lab[bi * 8 + s2] == c && (sfl[bi * 8 + s2] & F_NEG)) earlier_neg = true;
    }
    if (w0 && !earlier_any) {
      const uint32_t c = lab[bi * 8 + start_dir(m, 4)];
      if (c != kNone) CTR_ATOMIC_MIN(&stNew[c], pos * 2u);
    } else if (e0 && !earlier_neg) {
      const uint32_t c = lab[bi * 8 + start_dir(m, 0)];
      if (c != kNone) CTR_ATOMIC_MIN(&stNew[c], pos * 2u + 1u);
    }
  }
};
struct StartCommit {
  int32_t* ctr; int r; uint32_t *stOld, *stNew;
  CTR_HD bool operator()(long i) const {          // true: raise ctr[C_START + r]
    if (i >= 8L * ctr[C_NB] || (r > 0 && !ctr[C_START + (r - 1) % 3])) return false;
    const uint32_t v = stNew[i];
    const bool moved = v != stOld[i];
    if (moved) stOld[i] = v;
    stNew[i] = kNone;
    return moved;
  }
};

struct MarkStarts {
  Geo g; const int32_t* ctr; const uint8_t* gmask; const uint32_t* goff; const uint32_t* stOld;
  uint8_t* sflag; uint32_t* cyc_at;
  CTR_HD void operator()(long i) const {
    if (i >= 8L * ctr[C_NB]) return;
    const uint32_t st = stOld[i];
    if (st == kNone) return;
    const uint32_t pos = st >> 1;
    const uint32_t bi = border_index(g, gmask, goff, (int)(pos % (uint32_t)g.w), (int)(pos / (uint32_t)g.w));
    sflag[bi] = 1; cyc_at[bi] = (uint32_t)i;
  }
};

struct BorderTable {
  const int32_t* ctr; const uint8_t* sflag; const uint32_t *srank, *cyc_at, *stOld, *nxt, *ec;
  const uint8_t *nbm, *sfl; const long long* area;
  uint32_t* nbd; int32_t *btype, *bnpts, *bbi; uint32_t* bss; long long* barea;
  CTR_HD void operator()(long bi) const {
    if (bi >= ctr[C_NB] || !sflag[bi]) return;
    const uint32_t b = srank[bi], c = cyc_at[bi];
    const int type = (int)(stOld[c] & 1u);
    const unsigned m = nbm[bi];
    nbd[c] = b + 2;
    btype[b] = type; bbi[b] = (int32_t)bi;
    if (m == 0) { bss[b] = kNone; bnpts[b] = 1; barea[b] = 0; return; }
    bss[b] = (uint32_t)bi * 8u + (uint32_t)start_dir(m, type ? 0 : 4);
    bnpts[b] = (int32_t)(ec[nxt[c]] + ((sfl[c] & F_EMIT) ? 1u : 0u));
    barea[b] = area[c];
  }
};

// |f(q)| as the raster scan sees it at `time` (cycles started at positions <= time); 0 = unmarked
CTR_HD uint32_t mark_at(long q, long time, const uint8_t* nbm, const uint32_t* lab, const uint8_t* sfl,
                        const uint32_t* stOld, const uint32_t* nbd) {
  const unsigned m = nbm[q];
  if (m == 0) {
    const uint32_t st = stOld[q * 8];
    return (st != kNone && (long)(st >> 1) <= time) ? nbd[q * 8] : 0u;
  }
  uint32_t neg_st = 0, neg_c = kNone, pos_st = kNone, pos_c = kNone;
  for (int s = 0; s < 8; ++s) {
    if (!((m >> s) & 1)) continue;
    const uint32_t c = lab[q * 8 + s];
    if (c == kNone) continue;
    const uint32_t st = stOld[c];
    if (st == kNone || (long)(st >> 1) > time) continue;
    bool neg = false;
    for (int s2 = 0; s2 < 8; ++s2)
      if (((m >> s2) & 1) && lab[q * 8 + s2] == c && (sfl[q * 8 + s2] & F_NEG)) neg = true;
    if (neg && (neg_c == kNone || st > neg_st)) { neg_st = st; neg_c = c; }
    if (pos_c == kNone || st < pos_st) { pos_st = st; pos_c = c; }
  }
  if (neg_c != kNone) return nbd[neg_c];
  if (pos_c != kNone) return nbd[pos_c];
  return 0u;
}

struct LastMark {
  int w; const int32_t* ctr; const uint32_t* px; const uint8_t* nbm; const uint32_t* lab;
  const uint8_t* sfl; const uint32_t *stOld, *nbd; const int32_t *btype, *bbi; int32_t* blnbd;
  CTR_HD void operator()(long b) const {
    if (b >= ctr[C_NBORDERS]) return;
    const long bi = bbi[b];
    const long pos = px[bi];
    uint32_t l = 0;
    if (btype[b]) l = mark_at(bi, pos - 1, nbm, lab, sfl, stOld, nbd);
    if (!l) {
      const long row = pos / w;
      for (long q = bi - 1; q >= 0 && (long)px[q] / w == row; --q) {
        l = mark_at(q, (long)px[q], nbm, lab, sfl, stOld, nbd);
        if (l) break;
      }
    }
    blnbd[b] = l ? (int32_t)l : 1;
  }
};

// Suzuki's parent table: the last border met if its type differs, else that border's parent
struct Parents {
  const int32_t* ctr; const int32_t *btype, *blnbd; int32_t* bparent;
  CTR_HD void operator()(long b) const {
    const long n = ctr[C_NBORDERS];
    if (b >= n) return;
    const int tb = btype[b];
    int l = blnbd[b], parent = 0;
    for (long it = 0; it <= n; ++it) {
      if (l == 0) break;
      const int tl = l == 1 ? 1 : btype[l - 2];
      if (tl != tb) { parent = l; break; }
      l = l == 1 ? 0 : blnbd[l - 2];
    }
    bparent[b] = parent;
  }
};

struct EmitPoints {
  Geo g; const int32_t* ctr; const uint32_t* lab; const uint8_t* sfl; const uint32_t *nbd, *bss, *boff;
  const int32_t* bnpts; const uint32_t *ec, *px; int32_t* points; long max_points;
  CTR_HD uint32_t before(uint32_t k, uint32_t c, uint32_t total) const { return k == c ? 0u : total - ec[k]; }
  CTR_HD void operator()(long i) const {
    if (i >= 8L * ctr[C_NB]) return;
    const uint32_t c = lab[i];
    if (c == kNone || !(sfl[i] & F_EMIT)) return;
    const uint32_t nb = nbd[c];
    if (!nb) return;
    const uint32_t b = nb - 2, total = (uint32_t)bnpts[b];
    long r = (long)before((uint32_t)i, c, total) - (long)before(bss[b], c, total);
    if (r < 0) r += total;
    const long at = (long)boff[b] + r;
    if (at >= max_points) return;
    const uint32_t p = px[i >> 3];
    points[2 * at] = (int32_t)(p % (uint32_t)g.w);
    points[2 * at + 1] = (int32_t)(p / (uint32_t)g.w);
  }
};

struct WriteRecs {
  int w; const int32_t* ctr; const uint32_t* px; const int32_t *bbi, *btype, *bparent, *bnpts;
  const uint32_t *boff, *bss; const long long* barea; Rec* recs; long max_contours;
  int32_t* points; long max_points;
  CTR_HD void operator()(long b) const {
    if (b >= ctr[C_NBORDERS]) return;
    const uint32_t p = px[bbi[b]];
    const int x = (int)(p % (uint32_t)w), y = (int)(p / (uint32_t)w);
    if (bss[b] == kNone && (long)boff[b] < max_points) {     // isolated pixel: its single point
      points[2 * (long)boff[b]] = x; points[2 * (long)boff[b] + 1] = y;
    }
    if (b >= max_contours) return;
    Rec r;
    r.x = x; r.y = y; r.is_hole = btype[b]; r.parent = bparent[b];
    r.n_points = bnpts[b]; r.offset = (int32_t)boff[b]; r.area2 = barea[b];
    recs[b] = r;
  }
};

struct Finish {
  int32_t* ctr; Counts* out; long max_contours, max_points; int start_rounds;
  CTR_HD void operator()(long) const {
    int ov = ctr[C_OVERFLOW];
    if (ctr[C_NBORDERS] > max_contours) ov |= 2;
    if (ctr[C_NPOINTS] > max_points) ov |= 4;
    out->n_border_px = ctr[C_NB_RAW]; out->n_contours = ctr[C_NBORDERS]; out->n_points = ctr[C_NPOINTS];
    out->changed = ctr[C_START + (start_rounds - 1) % 3]; out->overflow = ov; out->pad_ = 0;
  }
};

// ------------------------------------------------------------------------------------------
struct Args {
  const uint8_t* img; int w, h, pitch, mode, thr;
  long cap;                 // max border pixels
  int start_rounds;         // Jacobi rounds of the start resolution (>= 2)
  Rec* recs; long max_contours;
  int32_t* points; long max_points;
  Counts* counts;
};

inline int ceil_log2(unsigned long long v) { int r = 0; while ((1ull << r) < v) ++r; return r; }

// The whole pass list.  Exec::each(n, f) applies f(i) for i in [0, n); each_n bounds the range by
// mult * *n_ptr read on the device (the counts are never known on the host: nothing synchronises);
// each_flag is each_n for a functor returning bool and sets *flag = 1 if any call returned true;
// Exec::scan(count, n_max, n_ptr, out, scratch, total) writes exclusive prefix sums of count(i).
template <class Exec>
void run(Exec& ex, const Args& a, Ws& ws) {
  Geo g{a.img, a.w, a.h, a.pitch, (a.w + 7) / 8, a.mode, a.thr};
  const long G = (long)g.gw * a.h, cap = a.cap, S = cap * 8;
  const int32_t *nB = ws.ctr + C_NB, *nBorders = ws.ctr + C_NBORDERS;
  ex.each(C_COUNT, Fill32{(uint32_t*)ws.ctr, 0u});
  const bool words = ((uintptr_t)a.img & 7) == 0 && (a.pitch & 7) == 0;
  ex.each(G, FlagGroups{g, ws.gmask, words});
  ex.scan(CountGroups{ws.gmask}, G, nullptr, ws.goff, ws.s1, ScanTotal{ws.ctr, C_NB, C_NB_RAW, 1, cap});
  ex.each(G, ScatterBorder{g, ws.gmask, ws.goff, ws.ctr, ws.px, ws.nbm});
  // ws.stOld / ws.stNew double as the window counts of the doubling rounds (re-filled below)
  uint32_t *lab = ws.labA, *lab2 = ws.labB, *jmp = ws.jmpA, *jmp2 = ws.jmpB;
  uint32_t *cnt = ws.cntA, *cnt2 = ws.cntB, *ec = ws.ecA, *ec2 = ws.ecB, *ew = ws.stOld, *ew2 = ws.stNew;
  ex.each_n(S, nB, 8, InitStates{g, ws.gmask, ws.goff, ws.ctr, ws.px, ws.nbm, ws.nxt, ws.sfl, lab, jmp, cnt, ec, ew});
  int rounds = ceil_log2((unsigned long long)S) + 1;
  if (rounds > 31) rounds = 31;
  for (int r = 0; r < rounds; ++r) {
    ex.each_flag(S, nB, 8, LabelRound{ws.ctr, r, lab, jmp, cnt, ec, ew, lab2, jmp2, cnt2, ec2, ew2},
                 ws.ctr + C_LABEL + r);
    uint32_t* t = lab; lab = lab2; lab2 = t;
    t = jmp; jmp = jmp2; jmp2 = t;
    t = cnt; cnt = cnt2; cnt2 = t;
    t = ec; ec = ec2; ec2 = t;
    t = ew; ew = ew2; ew2 = t;
  }
  ex.each_n(S, nB, 8, InitCycles{ws.ctr, ws.area, ws.stOld, ws.stNew, ws.nbd, ws.sflag});
  ex.each_n(S, nB, 8, AreaAccum{g, ws.ctr, lab, ws.nxt, ws.px, ws.area});
  for (int r = 0; r < a.start_rounds; ++r) {
    ex.each_n(cap, nB, 1, StartRound{ws.ctr, r, ws.px, ws.nbm, lab, ws.sfl, ws.stOld, ws.stNew});
    ex.each_flag(S, nB, 8, StartCommit{ws.ctr, r, ws.stOld, ws.stNew}, ws.ctr + C_START + r % 3);
  }
  ex.each_n(S, nB, 8, MarkStarts{g, ws.ctr, ws.gmask, ws.goff, ws.stOld, ws.sflag, ws.cyc_at});
  ex.scan(CountFlags{ws.sflag}, cap, nB, ws.srank, ws.s1, ScanTotal{ws.ctr, C_NBORDERS, -1, 0, -1});
  ex.each_n(cap, nB, 1, BorderTable{ws.ctr, ws.sflag, ws.srank, ws.cyc_at, ws.stOld, ws.nxt, ec, ws.nbm, ws.sfl,
                                    ws.area, ws.nbd, ws.btype, ws.bnpts, ws.bbi, ws.bss, ws.barea});
  ex.each_n(cap, nBorders, 1, LastMark{a.w, ws.ctr, ws.px, ws.nbm, lab, ws.sfl, ws.stOld, ws.nbd, ws.btype,
                                       ws.bbi, ws.blnbd});
  ex.each_n(cap, nBorders, 1, Parents{ws.ctr, ws.btype, ws.blnbd, ws.bparent});
  ex.scan(CountI32{ws.bnpts}, cap, nBorders, ws.boff, ws.s1, ScanTotal{ws.ctr, C_NPOINTS, -1, 0, -1});
  ex.each_n(S, nB, 8, EmitPoints{g, ws.ctr, lab, ws.sfl, ws.nbd, ws.bss, ws.boff, ws.bnpts, ec, ws.px, a.points,
                                 a.max_points});
  ex.each_n(cap, nBorders, 1, WriteRecs{a.w, ws.ctr, ws.px, ws.bbi, ws.btype, ws.bparent, ws.bnpts, ws.boff,
                                        ws.bss, ws.barea, a.recs, a.max_contours, a.points, a.max_points});
  ex.each(1, Finish{ws.ctr, a.counts, a.max_contours, a.max_points, a.start_rounds});
}

}  // namespace ctr
