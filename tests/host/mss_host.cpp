// TEST INFRASTRUCTURE: runs the scalar kernel logic of deepgrp_b200/csrc/mss_core.cuh on the CPU,
// emulating the grid of mss.cu one "thread" at a time, so the chunked algorithm can be checked
// against the oracle without a GPU.  Build: g++ -O2 -shared -fPIC -o tests/host/libmss_host.so
// tests/host/mss_host.cpp
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../deepgrp_b200/csrc/mss_core.cuh"

using namespace dgrp::mss;

struct Seg { int st, en; double sc; };

// Hierarchical summary chain (emulates mss_group_compose / mss_super_compose / mss_group_chain /
// mss_super_fill / mss_group_fill of mss.cu): groups of G chunks and super-groups of G groups are composed
// in parallel, one sequential pass walks the super-groups, the start states are filled in level by
// level.  Returns the number of stale chunks.
static int chain_grouped(int NC, int G, const std::vector<ScanState> &used, const std::vector<ScanState> &out,
                         const std::vector<ChunkSummary> &sum, std::vector<ScanState> &pred,
                         std::vector<uint8_t> &dirty, double L0 = 0.0, bool force = false) {
  const int NG = (NC + G - 1) / G;
  std::vector<Composite> comp(NG);
  for (int g = 0; g < NG; ++g) {
    const int c0 = g * G, c1 = c0 + G < NC ? c0 + G : NC;
    Composite acc{sum[c0], used[c0], out[c0]};
    for (int c = c0 + 1; c < c1; ++c) acc = compose(acc, Composite{sum[c], used[c], out[c]});
    comp[g] = acc;
  }
  // second level, as on the GPU: super-groups of G groups are composed, one sequential pass over the
  // super-groups, then every super-group fills in its groups' start states
  const int NS = (NG + G - 1) / G;
  std::vector<Composite> super(NS);
  for (int q = 0; q < NS; ++q) {
    const int g0 = q * G, g1 = g0 + G < NG ? g0 + G : NG;
    Composite acc = comp[g0];
    for (int g = g0 + 1; g < g1; ++g) acc = compose(acc, comp[g]);
    super[q] = acc;
  }
  std::vector<ScanState> sstart(NS), gstart(NG);
  ScanState s; state_initial(s, L0);
  for (int q = 0; q < NS; ++q) {
    sstart[q] = s;
    s = state_equal(super[q].x_in, s) ? super[q].x_out : apply_summary(super[q].sum, super[q].x_in, super[q].x_out, s);
  }
  for (int q = 0; q < NS; ++q) {
    const int g0 = q * G, g1 = g0 + G < NG ? g0 + G : NG;
    ScanState t = sstart[q];
    for (int g = g0; g < g1; ++g) {
      gstart[g] = t;
      t = state_equal(comp[g].x_in, t) ? comp[g].x_out : apply_summary(comp[g].sum, comp[g].x_in, comp[g].x_out, t);
    }
  }
  int n_dirty = 0;
  for (int g = 0; g < NG; ++g) {
    const int c0 = g * G, c1 = c0 + G < NC ? c0 + G : NC;
    ScanState t = gstart[g];
    for (int c = c0; c < c1; ++c) {
      pred[c] = t;
      const bool same = state_equal(used[c], t);
      if (same && !force) dirty[c] = 0; else { dirty[c] = 1; ++n_dirty; }
      t = same ? out[c] : apply_summary(sum[c], used[c], out[c], t);
    }
  }
  return n_dirty;
}

// L0 / open_end / restart mirror run_mss_t of mss.cu: the scan starts from running sum L0; with open_end the
// array is a PREFIX of a record, only the candidates flushed by the last FLUSH run (or before) are returned
// and restart[0] / restart_L[0] receive that run's start and the running sum before it (restart[0] = -1: no run).
template <typename T>
static int run_all(int n, const T *S, double min_sc, double xdrop, int CH, Seg *segs_out, int cap,
                   int *rounds_out, int max_rounds, int group = 0, int *chain_mismatch = nullptr,
                   double L0 = 0.0, bool open_end = false, int *restart = nullptr, double *restart_L = nullptr) {
  if (restart) *restart = -1;
  if (n <= 0) { *rounds_out = 0; return 0; }
  const int min_sc_int = (int)min_sc;
  const int NC = (n + CH - 1) / CH;
  std::vector<int> base(NC + 1, 0);
  for (int c = 0; c < NC; ++c) {
    int cnt = 0;
    for (int i = c * CH; i < n && i < (c + 1) * CH; ++i)
      if ((double)S[i] > 0 && (i == 0 || !((double)S[i - 1] > 0))) ++cnt;
    base[c + 1] = base[c] + cnt;
  }
  int NR = base[NC];
  std::vector<int> st(NR + 1), en(NR + 1), pre(NR + 1);
  std::vector<double> L(NR + 1), R(NR + 1);
  std::vector<uint8_t> kind(NR + 1);
  RunTable rt{st.data(), en.data(), L.data(), R.data(), pre.data(), kind.data()};
  std::vector<ScanState> used(NC), out(NC), pred(NC);
  std::vector<ChunkSummary> sum(NC);
  std::vector<uint8_t> dirty(NC, 0);
  // round 1: chunk 0 from the initial state, every other chunk from the canonical state (emulates
  // mss_scan_kernel, one thread per chunk).  With more than one chunk the first round is speculative almost
  // everywhere, so it does not write its run records ("lean") and the first chain pass marks every chunk stale.
  const bool lean = NC > 1 && (max_rounds <= 0 || max_rounds >= 2);
  for (int c = 0; c < NC; ++c) {
    ScanState s;
    if (c == 0) state_initial(s, L0); else state_canonical(s);
    used[c] = s;
    int b = c * CH, e = b + CH < n ? b + CH : n;
    scan_chunk(S, n, xdrop, b, e, base[c], s, rt, sum[c], !lean);
    out[c] = s;
  }
  int rounds = 1;
  for (;;) {
    // chain pass (emulates mss_chain_kernel): predict every start state, mark stale chunks
    const bool force = lean && rounds == 1;
    int n_dirty = 0;
    ScanState s; state_initial(s, L0);
    for (int c = 0; c < NC; ++c) {
      pred[c] = s;
      const bool same = state_equal(used[c], s);
      if (same && !force) dirty[c] = 0; else { dirty[c] = 1; ++n_dirty; }
      s = same ? out[c] : apply_summary(sum[c], used[c], out[c], s);
    }
    if (group > 0) {
      // the grouped chain must predict what the sequential chain predicts (counted for the tests;
      // its predictions are then the ones used, as on the GPU)
      std::vector<ScanState> pred2(NC);
      std::vector<uint8_t> dirty2(NC, 0);
      const int nd2 = chain_grouped(NC, group, used, out, sum, pred2, dirty2, L0, force);
      if (chain_mismatch)
        for (int c = 0; c < NC; ++c)
          if (!state_equal(pred2[c], pred[c])) ++*chain_mismatch;
      pred = pred2; dirty = dirty2; n_dirty = nd2;
    }
    if (!n_dirty) break;
    if (max_rounds > 0 && rounds >= max_rounds) {
      // sequential completion (emulates mss_complete_kernel)
      ScanState t; state_initial(t, L0);
      for (int c = 0; c < NC; ++c) {
        if (!state_equal(used[c], t)) {
          used[c] = t;
          int b = c * CH, e = b + CH < n ? b + CH : n;
          scan_chunk(S, n, xdrop, b, e, base[c], t, rt, sum[c]);
          out[c] = t;
        } else t = out[c];
      }
      rounds = -rounds;
      break;
    }
    for (int c = 0; c < NC; ++c) {
      if (!dirty[c]) continue;
      ScanState t = pred[c];
      used[c] = t;
      int b = c * CH, e = b + CH < n ? b + CH : n;
      scan_chunk(S, n, xdrop, b, e, base[c], t, rt, sum[c]);
      out[c] = t;
    }
    ++rounds;
    // parallel verification (emulates mss_verify_kernel)
    int bad = 0;
    for (int c = 0; c < NC; ++c) {
      ScanState expect;
      if (c == 0) state_initial(expect, L0); else expect = out[c - 1];
      if (!state_equal(used[c], expect)) ++bad;
    }
    if (!bad) break;
  }
  *rounds_out = rounds;
  if (open_end) {
    // the last FLUSH run: the stack is emptied there whatever follows, so the candidates flushed up to it are
    // final and the record can be resumed at its first element
    int kF = -1;
    for (int k = 0; k < NR; ++k) if (kind[k] == RUN_FLUSH) kF = k;
    if (kF < 0) return 0;
    if (restart) *restart = st[kF];
    if (restart_L) *restart_L = L[kF];
    NR = kF;
  }
  // regions
  std::vector<int> ev;
  for (int k = 0; k < NR; ++k) if (kind[k] != RUN_PLAIN) ev.push_back(k);
  std::vector<uint8_t> live(NR + 1, 0);
  for (size_t r = 0; r < ev.size(); ++r) {
    int k0 = ev[r], k1 = r + 1 < ev.size() ? ev[r + 1] : NR;
    int depth = process_region(k0, k1, rt);
    bool flushed = (k1 == NR) || kind[k1] == RUN_FLUSH;
    for (int j = 0; j < depth; ++j)
      live[k0 + j] = flushed && (R[k0 + j] - L[k0 + j] >= min_sc_int);
  }
  int m = 0;
  for (int k = 0; k < NR; ++k)
    if (live[k]) { if (m < cap) { segs_out[m].st = st[k]; segs_out[m].en = en[k]; segs_out[m].sc = R[k] - L[k]; } ++m; }
  return m;
}

extern "C" int host_mss_grouped_f64(int n, const double *S, double min_sc, double xdrop, int CH, int group, Seg *out,
                                    int cap, int *rounds, int max_rounds, int *chain_mismatch) {
  *chain_mismatch = 0;
  return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds, group, chain_mismatch);
}
extern "C" int host_mss_grouped_f32(int n, const float *S, double min_sc, double xdrop, int CH, int group, Seg *out,
                                    int cap, int *rounds, int max_rounds, int *chain_mismatch) {
  *chain_mismatch = 0;
  return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds, group, chain_mismatch);
}
// prefix of a record, resumable (see run_all)
extern "C" int host_mss_open_f32(int n, const float *S, double min_sc, double xdrop, int CH, int group, double L0,
                                 int open_end, Seg *out, int cap, int *rounds, int max_rounds, int *restart,
                                 double *restart_L) {
  return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds, group, nullptr, L0, open_end != 0, restart,
                 restart_L);
}
extern "C" int host_mss_open_f64(int n, const double *S, double min_sc, double xdrop, int CH, int group, double L0,
                                 int open_end, Seg *out, int cap, int *rounds, int max_rounds, int *restart,
                                 double *restart_L) {
  return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds, group, nullptr, L0, open_end != 0, restart,
                 restart_L);
}
extern "C" int host_mss_f64(int n, const double *S, double min_sc, double xdrop, int CH, Seg *out,
                            int cap, int *rounds, int max_rounds) { return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds); }
extern "C" int host_mss_f32(int n, const float *S, double min_sc, double xdrop, int CH, Seg *out,
                            int cap, int *rounds, int max_rounds) { return run_all(n, S, min_sc, xdrop, CH, out, cap, rounds, max_rounds); }
