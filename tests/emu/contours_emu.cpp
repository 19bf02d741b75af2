// TEST INFRASTRUCTURE — host emulation of pychips_b200/csrc/contours_core.h.
// Runs exactly the passes contours.cu launches as kernels, as plain loops over host memory, so
// tests/test_contours_emu.py can check the pass logic against OpenCV on the CPU-only build box.
// Built by the test with g++ into a temporary directory; never loaded by the product.
#include <stdlib.h>
#include <string.h>
#include "../../pychips_b200/csrc/contours_core.h"

// Order in which a pass visits its indices: 0 ascending, 1 descending, 2 a fixed pseudo-random
// permutation.  A pass whose result depended on the order would be a race on the GPU.
static int g_order = 0;
extern "C" void emu_set_order(int order) { g_order = order; }

struct HostExec {
  template <class F>
  void each(long n, const F& f) {
    if (g_order == 1) {
      for (long i = n - 1; i >= 0; --i) f(i);
    } else if (g_order == 2 && n > 2) {
      long step = 2654435761L % n;                   // any step coprime to n walks every index once
      auto gcd = [](long a, long b) { while (b) { long t = a % b; a = b; b = t; } return a; };
      while (step < 2 || gcd(step, n) != 1) ++step;
      long i = n / 3;
      for (long k = 0; k < n; ++k) { f(i); i = (i + step) % n; }
    } else {
      for (long i = 0; i < n; ++i) f(i);
    }
  }
  template <class F>
  void each_n(long n_max, const int32_t* n_ptr, int mult, const F& f) {
    const long m = (long)mult * *n_ptr;
    each(m < n_max ? m : n_max, f);
  }
  template <class F>
  void each_flag(long n_max, const int32_t* n_ptr, int mult, const F& f, int32_t* flag) {
    const long m = (long)mult * *n_ptr;
    const long n = m < n_max ? m : n_max;
    each(n, [&](long i) { if (f(i)) *flag = 1; });
  }
  template <class Count, class Total>
  void scan(Count c, long n_max, const int32_t* n_ptr, uint32_t* out, uint32_t*, Total total) {
    const long n = n_ptr && *n_ptr < n_max ? *n_ptr : n_max;
    uint32_t run = 0;
    for (long i = 0; i < n; ++i) {
      out[i] = run;
      run += c(i);
    }
    total(run);
  }
};

extern "C" size_t emu_contours_workspace_bytes(int w, int h, long long cap) {
  ctr::Ws ws;
  return ctr::carve(ws, nullptr, w, h, (long)cap);
}

extern "C" int emu_contours(const uint8_t* img, int w, int h, int pitch, int mode, int thr, long long cap,
                            int start_rounds, void* recs, long long max_contours, int32_t* points,
                            long long max_points, void* counts, void* wsmem, size_t ws_bytes) {
  ctr::Ws ws;
  if (ctr::carve(ws, (char*)wsmem, w, h, (long)cap) > ws_bytes) return -3;
  ctr::Args a{img, w, h, pitch, mode, thr, (long)cap, start_rounds, (ctr::Rec*)recs, (long)max_contours,
              points, (long)max_points, (ctr::Counts*)counts};
  HostExec ex;
  ctr::run(ex, a, ws);
  return 0;
}
