// Scalar building blocks of the GPU FASTA decode (K1, csrc/fasta.cu): the 3-state line machine of
// _read_multi_fasta (deepgrp/__main__.py:20-43) as composable state functions.
//
// The functions are __host__ __device__ so that tests/host/fasta_host.cpp can compile this header with g++
// and run the exact kernel logic (per-thread items, tile scans, classification) on the CPU against the
// Python reader, without a GPU; the product only calls them from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DGRP_FA_HD __host__ __device__ __forceinline__
#else
#define DGRP_FA_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DGRP_FA_UNROLL _Pragma("unroll")
#else
#define DGRP_FA_UNROLL
#endif

namespace dgrp {
namespace fa {

enum : unsigned { ST_START = 0, ST_SEQ = 1, ST_HDR = 2 };
constexpr unsigned FN_IDENT = (0u) | (1u << 2) | (2u << 4);

DGRP_FA_HD bool is_ws(unsigned b) { return (b >= 9 && b <= 13) || (b >= 28 && b <= 32); }

// terminator: '\n', or '\r' not followed by '\n' (the '\r' of "\r\n" is plain whitespace)
DGRP_FA_HD bool is_term(const uint8_t *raw, int64_t n, int64_t i) {
  const unsigned b = raw[i];
  if (b == '\n') return true;
  if (b == '\r') return !(i + 1 < n && raw[i + 1] == '\n');
  return false;
}

DGRP_FA_HD unsigned step_state(unsigned s, unsigned b, bool term) {
  if (term) return ST_START;
  if (s == ST_START) return is_ws(b) ? ST_START : (b == '>' ? ST_HDR : ST_SEQ);
  return s;
}
DGRP_FA_HD unsigned byte_fn(unsigned b, bool term) {
  return step_state(0, b, term) | (step_state(1, b, term) << 2) | (step_state(2, b, term) << 4);
}
// (f then g)
DGRP_FA_HD unsigned compose(unsigned f, unsigned g) {
  return ((g >> (2 * (f & 3))) & 3) | (((g >> (2 * ((f >> 2) & 3))) & 3) << 2) |
         (((g >> (2 * ((f >> 4) & 3))) & 3) << 4);
}
DGRP_FA_HD unsigned apply_fn(unsigned f, unsigned s) { return (f >> (2 * s)) & 3; }

// f followed by one byte: a terminator sends every start state to ST_START, whitespace changes nothing,
// any other byte moves the fields that are ST_START (== 0) to ST_HDR ('>') or ST_SEQ.  Equal to
// compose(f, byte_fn(b, term)) for every f and byte, without the three table look-ups.
DGRP_FA_HD unsigned push_byte(unsigned f, unsigned b, bool term) {
  const unsigned z = ~(f | (f >> 1)) & 0x15u;        // low bit of every field that holds ST_START
  unsigned g = f | (b == '>' ? (z << 1) : z);
  g = is_ws(b) ? f : g;
  return term ? 0u : g;
}

// classification of byte i given the state in front of it
struct ByteClass {
  bool seq;      // sequence byte that survives strip()
  bool hdr;      // the '>' that starts a record
  bool blank;    // terminator of a line that is empty after strip()
};
DGRP_FA_HD ByteClass classify(const uint8_t *raw, int64_t n, int64_t i, unsigned b, bool term,
                                              unsigned s) {
  ByteClass c = {false, false, false};
  if (term) { c.blank = (s == ST_START); return c; }
  if (s == ST_START) {
    if (is_ws(b)) return c;
    if (b == '>') c.hdr = true; else c.seq = true;
    return c;
  }
  if (s == ST_SEQ) {
    if (!is_ws(b)) { c.seq = true; return c; }
    // whitespace inside a sequence line is kept unless only whitespace follows up to the line end
    // (rare: walks the text in global memory)
    int64_t j = i + 1;
    while (j < n && !is_term(raw, n, j) && is_ws(raw[j])) ++j;
    c.seq = !(j >= n || is_term(raw, n, j));
  }
  return c;
}

// The bytes one thread owns, in registers, plus one look-ahead byte (0 past the end of the buffer).
constexpr int FA_ITEMS = 16;
struct Items {
  unsigned b[FA_ITEMS + 1];
  int cnt;                                           // bytes of this thread that are inside the buffer
  // terminator: '\n', or '\r' not followed by '\n' (is_term on the registers)
  DGRP_FA_HD bool term(int k) const {
    return b[k] == '\n' || (b[k] == '\r' && b[k + 1] != '\n');
  }
};
// state function of the thread's bytes
DGRP_FA_HD unsigned items_fn(const Items &it) {
  unsigned f = FN_IDENT;
DGRP_FA_UNROLL
  for (int k = 0; k < FA_ITEMS; ++k)
    if (k < it.cnt) f = push_byte(f, it.b[k], it.term(k));
  return f;
}

}  // namespace fa
}  // namespace dgrp
