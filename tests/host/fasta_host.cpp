// TEST INFRASTRUCTURE: runs the scalar logic of deepgrp_b200/csrc/fasta_core.cuh on the CPU, emulating the grid
// of fasta.cu (256 threads x 16 bytes per tile; tile functions -> scan over tiles -> classification + count ->
// ordered scatter) one "thread" at a time, so the GPU FASTA decode can be checked against the Python reader
// without a GPU.  Build: g++ -O2 -shared -fPIC -o tests/host/libfasta_host.so tests/host/fasta_host.cpp
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../deepgrp_b200/csrc/fasta_core.cuh"

using namespace dgrp::fa;

static const int THREADS = 256;
static const int TILE = THREADS * FA_ITEMS;

// load_items of fasta.cu: the thread's bytes, zeros past the end, look-ahead = first byte of the next thread
static void load(const uint8_t *raw, int64_t n, int64_t base, Items &it) {
  const int64_t left = n - base;
  it.cnt = left <= 0 ? 0 : (left < FA_ITEMS ? (int)left : FA_ITEMS);
  for (int k = 0; k < FA_ITEMS; ++k) it.b[k] = k < it.cnt ? raw[base + k] : 0u;
  it.b[FA_ITEMS] = left > FA_ITEMS ? raw[base + FA_ITEMS] : 0u;
}

extern "C" {

// Returns 0, or 1 when the reference would raise IndexError (blank line).  seq_out has room for n bytes,
// hdr_pos / hdr_seq for n entries.
int fasta_host_decode(const uint8_t *raw, int64_t n, uint8_t *seq_out, int64_t *n_seq, int64_t *hdr_pos,
                      int64_t *hdr_seq, int64_t *n_hdr) {
  *n_seq = 0; *n_hdr = 0;
  if (n <= 0) return 0;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  // pass 1 (fa_tile_fn_kernel): function of every tile
  std::vector<unsigned> tile_fn(ntiles);
  for (int64_t t = 0; t < ntiles; ++t) {
    unsigned f = FN_IDENT;
    for (int th = 0; th < THREADS; ++th) {
      Items it; load(raw, n, t * TILE + (int64_t)th * FA_ITEMS, it);
      f = compose(f, items_fn(it));
    }
    tile_fn[t] = f;
  }
  // fa_tile_state_kernel: state in front of every tile
  std::vector<unsigned> tile_state(ntiles);
  unsigned carry = ST_START;
  for (int64_t t = 0; t < ntiles; ++t) { tile_state[t] = carry; carry = apply_fn(tile_fn[t], carry); }
  const unsigned final_state = carry;
  // pass 2 + 3 (fa_count_kernel, fa_scatter_kernel): classification from the scanned state, ordered output
  bool blank = false;
  int64_t ps = 0, ph = 0;
  for (int64_t t = 0; t < ntiles; ++t) {
    unsigned excl = FN_IDENT;
    for (int th = 0; th < THREADS; ++th) {
      const int64_t base = t * TILE + (int64_t)th * FA_ITEMS;
      Items it; load(raw, n, base, it);
      unsigned s = apply_fn(excl, tile_state[t]);
      for (int k = 0; k < FA_ITEMS; ++k) {
        if (k < it.cnt) {
          const bool term = it.term(k);
          const ByteClass c = classify(raw, n, base + k, it.b[k], term, s);
          blank |= c.blank;
          if (c.seq) seq_out[ps++] = (uint8_t)it.b[k];
          if (c.hdr) { hdr_pos[ph] = base + k; hdr_seq[ph] = ps; ++ph; }
          s = step_state(s, it.b[k], term);
        }
      }
      excl = compose(excl, items_fn(it));
    }
  }
  *n_seq = ps; *n_hdr = ph;
  const bool last_is_term = raw[n - 1] == '\n' || raw[n - 1] == '\r';
  return (blank || (!last_is_term && final_state == ST_START)) ? 1 : 0;   // run_fasta_decode's blank rule
}

// push_byte against the table form it replaces, for every function value and byte; returns the mismatches
int fasta_host_push_byte_mismatches(void) {
  int bad = 0;
  for (unsigned f = 0; f < 64; ++f) {
    if ((f & 3) == 3 || ((f >> 2) & 3) == 3 || ((f >> 4) & 3) == 3) continue;
    for (unsigned b = 0; b < 256; ++b)
      for (int term = 0; term < 2; ++term)
        bad += compose(f, byte_fn(b, term != 0)) != push_byte(f, b, term != 0);
  }
  return bad;
}

}  // extern "C"
