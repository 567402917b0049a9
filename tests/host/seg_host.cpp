// TEST INFRASTRUCTURE: runs the scalar logic of deepgrp_b200/csrc/seg_core.cuh on the CPU, emulating the grid of
// segments.cu (256 threads x 16 labels per tile laid over the 16-byte aligned address below the label pointer;
// flags -> tile counts -> exclusive scan -> rows through the tile's start / end lists) one "thread" at a time, so the
// segment kernels and the TSV number formatting can be checked against the oracle / Python without a GPU.
// Build: g++ -O2 -shared -fPIC -o tests/host/libseg_host.so tests/host/seg_host.cpp
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../deepgrp_b200/csrc/seg_core.cuh"

using namespace dgrp::seg;

static const int THREADS = 256;
static const int TILE = THREADS * PER;

// seg_word of segments.cu: the thread's 16 labels (0 outside the array) and their neighbours
static void word(const uint8_t *lab, int64_t n, int mis, int64_t v0, bool open, uint32_t *q, unsigned &ms, unsigned &me) {
  const int64_t p0 = v0 - mis;
  q[0] = q[1] = q[2] = q[3] = 0u;
  for (int k = 0; k < PER; ++k) {
    const int64_t p = p0 + k;
    if (p >= 0 && p < n) q[k >> 2] |= (uint32_t)lab[p] << (8 * (k & 3));
  }
  const unsigned prev = (p0 - 1 >= 0 && p0 - 1 < n) ? lab[p0 - 1] : 0u;
  const unsigned next = (p0 + PER >= 0 && p0 + PER < n) ? lab[p0 + PER] : 0u;
  flags16(q, prev, next, p0, n, open, ms, me);
}

extern "C" {

// (start, end, label) triples of lab[0, n) as segv_count_kernel + seg_scan_kernel + segv_scatter_kernel produce
// them; `mis` = the label pointer's offset from 16-byte alignment.  Returns the number of rows (all of them: with
// `open` the caller drops a last row that ends at n); -1 if a row was left incomplete or cap is too small.
int64_t seg_host_rows(const uint8_t *lab, int64_t n, int mis, int open, int64_t offset, int64_t *tri, int64_t cap) {
  if (n <= 0) return 0;
  const int64_t nblk = (n + mis + TILE - 1) / TILE;
  std::vector<int64_t> bs(nblk + 1, 0), be(nblk + 1, 0);
  for (int64_t b = 0; b < nblk; ++b) {
    int cs = 0, ce = 0;
    for (int t = 0; t < THREADS; ++t) {
      uint32_t q[4]; unsigned ms, me;
      word(lab, n, mis, b * TILE + (int64_t)t * PER, open != 0, q, ms, me);
      cs += __builtin_popcount(ms); ce += __builtin_popcount(me);
    }
    bs[b + 1] = bs[b] + cs; be[b + 1] = be[b] + ce;
  }
  const int64_t nrun = bs[nblk];
  if (nrun != be[nblk] || nrun > cap) return -1;
  for (int64_t i = 0; i < 3 * nrun; ++i) tri[i] = INT64_MIN;
  for (int64_t b = 0; b < nblk; ++b) {
    std::vector<int> spos, epos;
    std::vector<uint8_t> tile(TILE, 0);
    for (int t = 0; t < THREADS; ++t) {
      uint32_t q[4]; unsigned ms, me;
      word(lab, n, mis, b * TILE + (int64_t)t * PER, open != 0, q, ms, me);
      memcpy(&tile[t * PER], q, 16);
      for (unsigned m = ms; m; m &= m - 1) spos.push_back(t * PER + __builtin_ctz(m));
      for (unsigned m = me; m; m &= m - 1) epos.push_back(t * PER + __builtin_ctz(m));
    }
    const int64_t S0 = bs[b], E0 = be[b], ns = (int64_t)spos.size(), ne = (int64_t)epos.size();
    if (S0 - E0 < 0 || S0 - E0 > 1) return -1;
    const int64_t o_end = S0 + ns > E0 + ne ? S0 + ns : E0 + ne;
    const int64_t pos0 = b * TILE - mis + offset;
    for (int64_t f = 0; f < (o_end - E0) * 3; ++f) {
      const int64_t o = E0 + f / 3, fld = f % 3, si = o - S0, ei = o - E0;
      if (fld == 1) { if (ei < ne) tri[3 * o + 1] = pos0 + epos[ei] + 1; }
      else if (si >= 0 && si < ns) tri[3 * o + fld] = fld == 0 ? pos0 + spos[si] : (int64_t)tile[spos[si]];
    }
  }
  for (int64_t i = 0; i < 3 * nrun; ++i) if (tri[i] == INT64_MIN) return -1;
  return nrun;
}

// decimal text of v as tsv.cu writes it; returns the length (and checks fmt_len against it: -1 on a mismatch)
int seg_host_fmt(long long v, uint8_t *out) {
  uint8_t *e = fmt_put(out, v);
  const int len = (int)(e - out);
  return len == fmt_len(v) ? len : -1;
}

}  // extern "C"
