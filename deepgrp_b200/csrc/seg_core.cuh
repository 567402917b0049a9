// Scalar building blocks of the run-length segment kernels (K8, segments.cu) and of the TSV writer (tsv.cu), as
// __host__ __device__ code so that tests/host/seg_host.cpp can run the exact kernel logic on the CPU next to the
// oracle (the product only calls them from kernels).  Replaces deepgrp/sequence.pyx:40-53 (get_segments),
// :79-85 (yield_segments) and the "{}".format of deepgrp/__main__.py:288-292.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DGRP_SEG_HD __host__ __device__ __forceinline__
#else
#define DGRP_SEG_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DGRP_SEG_UNROLL _Pragma("unroll")
#else
#define DGRP_SEG_UNROLL
#endif

namespace dgrp {
namespace seg {

constexpr int PER = 16;   // labels per thread

// 0x80 in every byte of x that is not zero
DGRP_SEG_HD uint32_t nz_bytes(uint32_t x) {
  return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// the 0x80 flags of four bytes -> bits 0..3
DGRP_SEG_HD unsigned flags4(uint32_t f) { return (f * 0x00204081u) >> 28; }

// Start / end flags of 16 consecutive labels q[0..3] (little endian, 0 outside the array) at array index p0 ..
// p0 + 15 of lab[0, n); prev / next = the labels at p0 - 1 / p0 + 16 (0 outside).  Bit k of ms: a non-zero run
// starts at p0 + k; bit k of me: a run ends after p0 + k.  A run starts where the label differs from the one
// before it and ends where it differs from the one after it; the reference's loop bounds (`size - 1`,
// sequence.pyx:43-52) additionally split a run that reaches the record's last element into [s, n-1) and
// [n-1, n).  `open`: the labels are a PREFIX of a record that goes on -- that special case does not apply (a run
// that reaches n is still closed there; the caller drops it).
DGRP_SEG_HD void flags16(const uint32_t *q, unsigned prev, unsigned next, int64_t p0, int64_t n, bool open,
                         unsigned &ms_out, unsigned &me_out) {
  unsigned ms = 0u, me = 0u;
  if (p0 + PER <= n - 2 || (open && p0 + PER <= n)) {
    // Four labels per 32-bit operation.  Outside the array the labels read as 0, which makes position 0 a start
    // and position n - 1 an end by the ordinary rule; only positions n - 2 and n - 1 of a closed record need
    // the byte loop below.
    DGRP_SEG_UNROLL
    for (int i = 0; i < 4; ++i) {
      const uint32_t before = (q[i] << 8) | (i ? q[i - 1] >> 24 : prev);
      const uint32_t after = (q[i] >> 8) | ((i < 3 ? q[i + 1] : next) << 24);
      const uint32_t cur = nz_bytes(q[i]);
      ms |= flags4(cur & nz_bytes(q[i] ^ before)) << (4 * i);
      me |= flags4(cur & nz_bytes(q[i] ^ after)) << (4 * i);
    }
  } else {
    DGRP_SEG_UNROLL
    for (int k = 0; k < PER; ++k) {
      const int64_t p = p0 + k;
      const unsigned cur = (q[k >> 2] >> (8 * (k & 3))) & 0xffu;
      const unsigned pv = k ? (q[(k - 1) >> 2] >> (8 * ((k - 1) & 3))) & 0xffu : prev;
      const unsigned nx = k + 1 < PER ? (q[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu : next;
      const bool in = p >= 0 && p < n && cur != 0u;
      const bool st = in && (p == 0 || pv != cur || (!open && p == n - 1));
      const bool en = in && (p + 1 == n || nx != cur || (!open && p + 1 == n - 1));
      ms |= (unsigned)st << k;
      me |= (unsigned)en << k;
    }
  }
  ms_out = ms; me_out = me;
}

// ---- decimal text of a row's numbers ------------------------------------------------------------------
// Positions are below 2^32 for every record the path accepts (2^31 - 1 bases, mss.h:16): decimal digits with 32-bit
// arithmetic (a division by the constant 10 is a multiply-high), the 64-bit loop only beyond that.  The 64-bit
// divisions of the first version were ~4 000 instructions per row and bounded both TSV kernels.
DGRP_SEG_HD int n_digits32(uint32_t v) {
  return 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) +
         (v >= 10000000u) + (v >= 100000000u) + (v >= 1000000000u);
}
DGRP_SEG_HD int n_digits(unsigned long long v) {
  if (v <= 0xffffffffull) return n_digits32((uint32_t)v);
  int d = 1;
  while (v >= 10ull) { v /= 10ull; ++d; }
  return d;
}
DGRP_SEG_HD int fmt_len(long long v) {
  return v < 0 ? 1 + n_digits(0ull - (unsigned long long)v) : n_digits((unsigned long long)v);
}
DGRP_SEG_HD uint8_t *fmt_put(uint8_t *p, long long v) {
  unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
  if (v < 0) *p++ = '-';
  if (u <= 0xffffffffull) {
    uint32_t w = (uint32_t)u;
    const int d = n_digits32(w);
    for (int k = d - 1; k >= 0; --k) { const uint32_t q = w / 10u; p[k] = (uint8_t)('0' + (w - q * 10u)); w = q; }
    return p + d;
  }
  const int d = n_digits(u);
  for (int k = d - 1; k >= 0; --k) { p[k] = (uint8_t)('0' + (u % 10ull)); u /= 10ull; }
  return p + d;
}

}  // namespace seg
}  // namespace dgrp
