// Scalar building blocks of the chunked maximal-scoring-segment scan (K7).  Replaces
// deepgrp/_mss/mss.c:50-101 (mss_find_all) and :35-47 (move_segs).
//
// The functions are __host__ __device__ so that tests/ can compile this header with g++ and run
// the exact kernel logic on the CPU next to the oracle (the product only calls them from kernels).
//
// How the sequential algorithm is split ("K7"):
//
//  1. The reference walks S once keeping (L, max, candidate stack).  Every stack operation happens
//     at the END of a maximal run of positive scores; the run enters as candidate
//     t = {st, en, L = prefix sum before the run, R = prefix sum after it}.
//  2. Two invariants hold whenever the stack is non-empty:
//        (i)  the bottom candidate has the strictly smallest L on the stack;
//        (ii) max == bottom.R  (every other candidate's R is <= the R of its `pre` chain, which
//             ends at the bottom).
//     Hence what happens to the stack as a whole at a run end is decided by (L, max, bottom.L):
//        A: stack empty or !(bottom.L < t.L)  -> mss.c:78-81: flush (move_segs), t is the new bottom
//        B: else R > max                      -> t absorbs every candidate down to the bottom
//                                                (mss.c:73-76 repeatedly), then j < 0: the stack is
//                                                {st = bottom.st, L = bottom.L, R, en}; nothing flushed
//        0: else                              -> ordinary push / partial merges, bottom unchanged
//     and the x-drop reset (mss.c:89-92) also only reads (L, max).
//  3. So a REDUCED scan over S carrying (L, max, bottom.L, bottom.st) -- no stack -- yields, per run,
//     a record {st, en, L, R, kind} with bit-identical doubles, because it performs the same
//     additions in the same order.  After an A or B run the stack is exactly one known candidate,
//     so the runs between two consecutive A/B runs form a REGION whose stack evolution is
//     independent of everything before it.  Regions are processed in parallel (one thread each)
//     by the unmodified push/merge loop; a region is flushed if the next event is A (or the end)
//     and dropped if it is B (absorbed).
//  4. The reduced scan itself is chunked, one thread per chunk.  A chunk's execution is an exact,
//     deterministic function of the state it starts from, so the scan is correct as soon as every
//     chunk has been executed from exactly the state its predecessor ended in (chunk 0 starts from
//     the true initial state): a fixed point of  state_in[c] = state_out[c-1],  verified by bitwise
//     comparison.  To reach it in O(1) parallel rounds instead of walking the chain, the start states
//     are PREDICTED: in exact arithmetic the machine is shift-invariant in L, so the effect of a chunk
//     on (L, max, bottom) is summarised by a few numbers (ChunkSummary) and can be applied to any
//     start state in O(1) (apply_summary).  A cheap sequential pass over the chunk summaries predicts
//     all start states, every chunk whose prediction differs from what it last ran from is re-run
//     in parallel, and the pass repeats until nothing is stale.  Scores that are float32 values
//     (the fused path) add exactly in double, so predictions are exact and two rounds suffice;
//     where double rounding does occur a prediction can miss by an ulp, the bitwise check catches
//     it and the affected chunks simply run again (in the worst case the chain is walked
//     sequentially -- always bit-exact, never approximate).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DGRP_HD __host__ __device__ __forceinline__
#else
#define DGRP_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define DGRP_UNROLL _Pragma("unroll")
#else
#define DGRP_UNROLL
#endif

namespace dgrp {
namespace mss {

constexpr double kNegInf = -1e30;  // NEG_INF, mss.c:33

enum RunKind : uint8_t { RUN_PLAIN = 0, RUN_FLUSH = 1 /* A */, RUN_ABSORB = 2 /* B */ };

// State of the reduced scan between two elements.  Unused fields are kept at fixed values so that
// two states can be compared field by field.
struct ScanState {
  double L;        // running sum ("L" of mss.c:56)
  double maxv;     // "max" of mss.c:56
  double botL;     // bottom candidate's L   (0 when empty)
  double run_L0;   // L before the run in progress (0 when !in_run)
  int bot_st;      // bottom candidate's st  (-1 when empty)
  int run_st;      // start of the run in progress (-1 when !in_run)
  int run_ord;     // ordinal of the run in progress (-1 when !in_run or unknown)
  int flags;       // bit 0: stack empty, bit 1: in a positive run
};

DGRP_HD void state_canonical(ScanState &s) {
  s.L = 0.0; s.maxv = kNegInf; s.botL = 0.0; s.run_L0 = 0.0;
  s.bot_st = -1; s.run_st = -1; s.run_ord = -1; s.flags = 1;
}

DGRP_HD bool state_equal(const ScanState &a, const ScanState &b) {
  // compare bit patterns of the doubles through integer views (so that -0.0 != 0.0, NaN == NaN)
  union V { double d; long long i; };
  V x, y;
  x.d = a.L; y.d = b.L; if (x.i != y.i) return false;
  x.d = a.maxv; y.d = b.maxv; if (x.i != y.i) return false;
  x.d = a.botL; y.d = b.botL; if (x.i != y.i) return false;
  x.d = a.run_L0; y.d = b.run_L0; if (x.i != y.i) return false;
  return a.bot_st == b.bot_st && a.run_st == b.run_st && a.run_ord == b.run_ord && a.flags == b.flags;
}

// Per-run records, struct of arrays indexed by run ordinal.  After region processing the same
// slots hold the candidate stacks (in place), `pre` added.
struct RunTable {
  int *st, *en;
  double *L, *R;
  int *pre;
  uint8_t *kind;
};

// Close the run in progress at index k (exclusive end); mss.c:64-87 reduced to (L, max, bottom).
DGRP_HD void finish_run(ScanState &s, int k, const RunTable &rt) {
  const double R = s.L, tL = s.run_L0;
  const double old_max = s.maxv;
  if (R > s.maxv) s.maxv = R;                       // mss.c:64
  uint8_t kind;
  int rec_st;
  double rec_L;
  if ((s.flags & 1) || !(s.botL < tL)) {            // no candidate with L < t.L  ->  j < 0
    kind = RUN_FLUSH;
    s.botL = tL; s.bot_st = s.run_st; s.maxv = R;   // mss.c:79-80, t becomes the bottom
    s.flags &= ~1;
    rec_st = s.run_st; rec_L = tL;
  } else if (R > old_max) {                         // absorbs down to and including the bottom
    kind = RUN_ABSORB;
    s.maxv = R;
    rec_st = s.bot_st; rec_L = s.botL;
  } else {
    kind = RUN_PLAIN;
    rec_st = s.run_st; rec_L = tL;
  }
  if (s.run_ord >= 0) {
    rt.st[s.run_ord] = rec_st; rt.en[s.run_ord] = k;
    rt.L[s.run_ord] = rec_L; rt.R[s.run_ord] = R;
    rt.kind[s.run_ord] = kind;
  }
  s.flags &= ~2;
  s.run_st = -1; s.run_ord = -1; s.run_L0 = 0.0;
}

// What a chunk did to (max, bottom) after its last x-drop reset, in shift-invariant form.
struct ChunkSummary {
  double carryR;    // R of the run carried in from the previous chunk, if it ended in this chunk
  double minT;      // min t.L over runs that started AND ended in the chunk (latest on ties)
  double maxAfter;  // max R from that run on
  double maxIn;     // max R over runs that started and ended in the chunk
  int min_st;       // st of the minT run
  int flags;        // 1: a reset happened, 2: carried run ended here, 4: minT.. valid,
                    // 8: the run in progress at the chunk end started inside the chunk
};

// Reduced scan over elements [b, e) of S (e <= n) from state `s`.  `first_ord` is the ordinal of
// the first run that STARTS in [b, e).  Writes the run records and the chunk summary.
template <typename ScoreT>
DGRP_HD void scan_chunk(const ScoreT *S, int n, double xdrop, int b, int e, int first_ord,
                        ScanState &s, const RunTable &rt, ChunkSummary &sum) {
  int next_ord = first_ord;
  sum.carryR = 0.0; sum.minT = 0.0; sum.maxAfter = 0.0; sum.maxIn = 0.0; sum.min_st = -1; sum.flags = 0;
  bool carried = (s.flags & 2) != 0;       // the run in progress came from before the chunk
  // Scores are consumed in blocks of BL values (+1 look-ahead) held in registers, so that the loads of
  // a block are independent of the sequential state machine and overlap each other.
  constexpr int BL = 8;
  bool prev_pos = b > 0 && ((double)S[b - 1] > 0);
  for (int i0 = b; i0 < e; i0 += BL) {
    const int m = e - i0 < BL ? e - i0 : BL;
    double blk[BL + 1];
    DGRP_UNROLL
    for (int j = 0; j <= BL; ++j) {
      const int idx = i0 + j;
      blk[j] = (j <= m && idx < n) ? (double)S[idx] : 0.0;
    }
    DGRP_UNROLL
    for (int j = 0; j < BL; ++j) {
      if (j >= m) break;
      const int i = i0 + j;
      const double v = blk[j];
      if (v > 0) {
        if (!(s.flags & 2)) {
          s.flags |= 2;
          s.run_st = i; s.run_L0 = s.L;
          // a "run" that begins at a speculative chunk start in the middle of a true run has no
          // ordinal; it only shapes the (to be discarded) speculative state
          const bool true_start = !prev_pos;
          s.run_ord = true_start ? next_ord++ : -1;
          carried = !true_start;
        }
        s.L += v;                                   // R = L + S[i]; R += S[k]  (mss.c:61-63)
        if (i + 1 == n || !(blk[j + 1] > 0)) {
          const double tL = s.run_L0;
          const int st = s.run_st;
          finish_run(s, i + 1, rt);
          const double R = s.L;
          if (carried) {
            sum.flags |= 2; sum.carryR = R;
          } else {
            if (!(sum.flags & 4) || !(sum.minT < tL)) { sum.minT = tL; sum.min_st = st; sum.maxAfter = R; }
            else if (R > sum.maxAfter) sum.maxAfter = R;
            if (!(sum.flags & 4) || R > sum.maxIn) sum.maxIn = R;
            sum.flags |= 4;
          }
          carried = false;
        }
        prev_pos = true;
      } else {
        if (xdrop > 0.0 && s.L + v + xdrop < s.maxv) {  // mss.c:89
          s.L = 0.0; s.maxv = kNegInf;              // mss.c:91 (the flush happens at the next run)
          s.botL = 0.0; s.bot_st = -1; s.flags |= 1;
          sum.flags = 1;                            // summary restarts after a reset
        }
        s.L += v;                                   // mss.c:93
        prev_pos = false;
      }
    }
  }
  if ((s.flags & 2) && !carried) sum.flags |= 8;
}

// Predict the end state of a chunk for start state `y` from one known execution x_in -> x_out with
// summary `sum`.  Exact when all additions involved are exact; otherwise only a guess that the
// bitwise verification will reject.
DGRP_HD ScanState apply_summary(const ChunkSummary &sum, const ScanState &x_in,
                                const ScanState &x_out, const ScanState &y) {
  if (sum.flags & 1) return x_out;                  // after a reset the frame is absolute
  const double d = y.L - x_in.L;
  ScanState o = x_out;
  o.L = x_out.L + d;
  if (x_out.flags & 2) {
    if (sum.flags & 8) o.run_L0 = x_out.run_L0 + d;
    else if (y.flags & 2) { o.run_L0 = y.run_L0; o.run_st = y.run_st; o.run_ord = y.run_ord; }
    else o.run_L0 = x_out.run_L0 + d;
  }
  double bot = y.botL, mx = y.maxv;
  int bst = y.bot_st;
  bool empty = (y.flags & 1) != 0;
  if (sum.flags & 2) {                              // the carried run ended here
    const double tL = (y.flags & 2) ? y.run_L0 : x_in.run_L0 + d;
    const int st = (y.flags & 2) ? y.run_st : x_in.run_st;
    const double R = sum.carryR + d;
    if (empty || !(bot < tL)) { bot = tL; bst = st; mx = R; empty = false; }
    else if (R > mx) mx = R;
  }
  if (sum.flags & 4) {
    const double mT = sum.minT + d;
    if (empty || !(bot < mT)) { bot = mT; bst = sum.min_st; mx = sum.maxAfter + d; empty = false; }
    else if (sum.maxIn + d > mx) mx = sum.maxIn + d;
  }
  o.botL = empty ? 0.0 : bot;
  o.bot_st = empty ? -1 : bst;
  o.maxv = mx;
  o.flags = (x_out.flags & 2) | (empty ? 1 : 0);
  return o;
}

// ---- composition of chunk effects (parallel summary chain) -------------------------------------
// A chunk's recorded execution x_in -> x_out plus its summary predicts its end state for any start
// state (apply_summary).  Two consecutive chunks compose into one object of the same kind: the runs
// that started and ended inside either chunk, plus the run that crosses the boundary, fold into one
// (latest minimum t.L, maximum R from there on, maximum R overall) triple -- the fold the state machine
// itself applies run by run (finish_run: a run whose t.L is not above the bottom replaces the bottom).
// Like apply_summary this is exact in exact arithmetic and only a prediction otherwise; every
// predicted start state is verified bitwise by the caller.
struct Composite {
  ChunkSummary sum;
  ScanState x_in, x_out;
};

// append the summary (minT, min_st, maxAfter, maxIn) of a later run list to p's
DGRP_HD void fold_runs(ChunkSummary &p, double minT, int min_st, double maxAfter, double maxIn) {
  if (!(p.flags & 4) || !(p.minT < minT)) { p.minT = minT; p.min_st = min_st; p.maxAfter = maxAfter; }
  else if (maxIn > p.maxAfter) p.maxAfter = maxIn;
  if (!(p.flags & 4) || maxIn > p.maxIn) p.maxIn = maxIn;
  p.flags |= 4;
}

DGRP_HD Composite compose(const Composite &A, const Composite &B) {
  Composite C;
  C.x_in = A.x_in;
  C.x_out = apply_summary(B.sum, B.x_in, B.x_out, A.x_out);
  ChunkSummary c;
  c.carryR = 0.0; c.minT = 0.0; c.maxAfter = 0.0; c.maxIn = 0.0; c.min_st = -1; c.flags = 0;
  if ((A.sum.flags & 1) || (B.sum.flags & 1)) {   // a reset inside: the end state is absolute
    c.flags = 1;
    C.sum = c;
    return C;
  }
  const double dB = A.x_out.L - B.x_in.L;         // B's recorded frame -> the frame of A's execution
  c = A.sum;
  c.flags = A.sum.flags & (2 | 4);
  const bool a_in_run = (A.x_out.flags & 2) != 0, a_run_inside = (A.sum.flags & 8) != 0;
  bool open_inside = false;                       // run in progress at the end started inside A+B
  if (a_in_run) {
    if (B.sum.flags & 2) {                        // the run crossing the boundary ends in B
      const double R = B.sum.carryR + dB;
      if (a_run_inside) fold_runs(c, A.x_out.run_L0, A.x_out.run_st, R, R);
      else { c.flags |= 2; c.carryR = R; }        // it was carried into A as well
    } else if (a_run_inside) {
      open_inside = true;                         // still running at the end of B
    }
  }
  if (B.sum.flags & 4) fold_runs(c, B.sum.minT + dB, B.sum.min_st, B.sum.maxAfter + dB, B.sum.maxIn + dB);
  if (B.sum.flags & 8) open_inside = true;
  if (open_inside) c.flags |= 8;
  C.sum = c;
  return C;
}

// Stack evolution of one region: runs [k0, k1) where k0 is a FLUSH/ABSORB run (its record is the
// single candidate on the stack) and no other event lies inside.  The push/merge loop of
// mss.c:65-86, with the stack stored in place at slots k0.. .  Returns the final depth.
DGRP_HD int process_region(int k0, int k1, const RunTable &rt) {
  int depth = 1;
  rt.pre[k0] = -1;
  for (int k = k0 + 1; k < k1; ++k) {
    int t_st = rt.st[k], t_en = rt.en[k], t_pre;
    double t_L = rt.L[k];
    const double t_R = rt.R[k];
    for (;;) {
      int j = depth - 1;
      while (j >= 0) {
        if (rt.L[k0 + j] < t_L) break;
        const int pj = rt.pre[k0 + j];
        j = pj >= 0 ? pj : j - 1;
      }
      if (j >= 0 && rt.R[k0 + j] < t_R) {
        t_st = rt.st[k0 + j]; t_L = rt.L[k0 + j]; t_pre = rt.pre[k0 + j];
        depth = j;
        (void)t_pre;
      } else {
        // j < 0 cannot happen inside a region (it would have been classified FLUSH/ABSORB)
        t_pre = j;
        const int slot = k0 + depth;
        rt.st[slot] = t_st; rt.en[slot] = t_en; rt.L[slot] = t_L; rt.R[slot] = t_R;
        rt.pre[slot] = t_pre;
        ++depth;
        break;
      }
    }
  }
  return depth;
}

}  // namespace mss
}  // namespace dgrp
