// Host-side schedule for the fused operator + gather-scatter kernel (ax8_fused_kernel): the segments of
// the direct-stiffness sum grouped by the element that CLOSES them, i.e. the element holding the
// segment's highest local index.  When elements are processed in increasing order, every other copy of
// a segment closed by element e lives in an element with a smaller index, so the worker that finishes
// e (or a later one) can complete those segments as soon as the listed dependency elements are done.
// Pure C++ (no CUDA) so that it can be unit-tested on the CPU (tests/test_closure_tables.py).
#pragma once

#include <stdint.h>

#include <algorithm>
#include <vector>

namespace semb {

struct ClosureTables {
  std::vector<int32_t> base;       // E + 1: start of element e's block in `ids`
  std::vector<uint16_t> n2, n4, n8;  // E each: segments of length 2 / 4 / 8 closed by element e
  std::vector<int32_t> ids;        // per element [n2 pairs | n4 quads | n8 octets], ascending inside a segment
  std::vector<int32_t> dep_start;  // E + 1
  std::vector<int32_t> dep_elem;   // elements (all < e) that hold copies of e's closed segments
  int64_t other_segments = 0;      // segments of any other length (left to the CSR kernel)
};

// Returns 0, or a negative code: -1 bad input, -2 more than 65535 segments of one class on an element.
inline int build_closure_tables(int64_t E, int64_t Np, const int64_t* g2l, const int64_t* gstart, int64_t nglobal,
                                ClosureTables& t) {
  if (E <= 0 || Np <= 0 || !g2l || !gstart || nglobal <= 0) return -1;
  struct Seg {
    int32_t closer;  // element of the highest local index
    int32_t cls;     // 0, 1, 2 for length 2, 4, 8
    int64_t seg;
  };
  std::vector<Seg> segs;
  for (int64_t s = 0; s < nglobal; ++s) {
    const int64_t lo = gstart[s], len = gstart[s + 1] - lo;
    if (len < 2) continue;
    const int cls = len == 2 ? 0 : len == 4 ? 1 : len == 8 ? 2 : -1;
    if (cls < 0) {
      t.other_segments++;
      continue;
    }
    int64_t last = g2l[lo];
    for (int64_t q = lo + 1; q < lo + len; ++q) last = std::max(last, g2l[q]);
    if (last < 0 || last >= E * Np) return -1;
    segs.push_back(Seg{(int32_t)(last / Np), (int32_t)cls, s});
  }
  std::stable_sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) {
    return a.closer != b.closer ? a.closer < b.closer : a.cls < b.cls;
  });
  t.base.assign((size_t)E + 1, 0);
  t.n2.assign((size_t)E, 0);
  t.n4.assign((size_t)E, 0);
  t.n8.assign((size_t)E, 0);
  t.dep_start.assign((size_t)E + 1, 0);
  t.ids.clear();
  t.dep_elem.clear();
  size_t at = 0;
  std::vector<int32_t> deps;
  for (int64_t e = 0; e < E; ++e) {
    t.base[(size_t)e] = (int32_t)t.ids.size();
    t.dep_start[(size_t)e] = (int32_t)t.dep_elem.size();
    deps.clear();
    int64_t cnt[3] = {0, 0, 0};
    while (at < segs.size() && segs[at].closer == e) {
      const Seg& sg = segs[at++];
      const int64_t lo = gstart[sg.seg], len = gstart[sg.seg + 1] - lo;
      const size_t first = t.ids.size();
      for (int64_t q = lo; q < lo + len; ++q) {
        t.ids.push_back((int32_t)g2l[q]);
        const int32_t el = (int32_t)(g2l[q] / Np);
        if (el != e) deps.push_back(el);
      }
      std::sort(t.ids.begin() + first, t.ids.end());  // fixed summation order: ascending local index
      cnt[sg.cls]++;
    }
    if (cnt[0] > 65535 || cnt[1] > 65535 || cnt[2] > 65535) return -2;
    t.n2[(size_t)e] = (uint16_t)cnt[0];
    t.n4[(size_t)e] = (uint16_t)cnt[1];
    t.n8[(size_t)e] = (uint16_t)cnt[2];
    std::sort(deps.begin(), deps.end());
    deps.erase(std::unique(deps.begin(), deps.end()), deps.end());
    t.dep_elem.insert(t.dep_elem.end(), deps.begin(), deps.end());
  }
  t.base[(size_t)E] = (int32_t)t.ids.size();
  t.dep_start[(size_t)E] = (int32_t)t.dep_elem.size();
  return 0;
}

}  // namespace semb
