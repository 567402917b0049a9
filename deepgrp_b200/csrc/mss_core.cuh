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

// The state a scan STARTS from: nothing on the stack, running sum L0.  L0 = 0 is the start of a record
// (mss.c:59); a non-zero L0 resumes a record at the first element of a FLUSH run (see run_mss_segments in
// mss.cu, "open end"): there the reference's stack has just been emptied and everything the machine still
// reads is the running sum, so the run classifies as FLUSH again and every later double is the same.
DGRP_HD void state_initial(ScanState &s, double L0) {
  state_canonical(s);
  s.L = L0;
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
// `store` = false only computes the state (a speculative first round whose records would be rewritten anyway).
DGRP_HD void finish_run(ScanState &s, int k, const RunTable &rt, bool store = true) {
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
  if (store && s.run_ord >= 0) {
    rt.st[s.run_ord] = rec_st; rt.en[s.run_ord] = k;
    rt.L[s.run_ord] = rec_L; rt.R[s.run_ord] = R;
    rt.kind[s.run_ord] = kind;
  }
  s.flags &= ~2;
  s.run_st = -1; s.run_ord = -1; s.run_L0 = 0.0;
}

// What a chunk did to (max, bottom) after its last x-drop reset, in shift-invariant form: the values the
// running sum L had at the points that matter (trajectory samples).
struct Samples {
  double outL;      // L at the chunk end
  double runL0;     // L before the run in progress at the chunk end, if that run started inside the chunk
  double carryR;    // R of the run carried in from the previous chunk, if it ended in this chunk
  double minT;      // min t.L over runs that started AND ended in the chunk (latest on ties)
  double maxAfter;  // max R from that run on
  double maxIn;     // max R over runs that started and ended in the chunk
};
// Shift invariance under rounding.  While L is small every addition is exact and a chunk executed from
// L0 + d is the execution from L0 shifted by d.  Once |L| is large (a record of 10^7..10^8 bases whose scores
// drift upwards: no x-drop reset ever fires, L reaches 2^28) the sums round to the grid g = ulp(L) of L's
// binade, and rn(x + d) = rn(x) + d still holds for every d that is an EVEN multiple of g -- but for an odd
// multiple the ties (x exactly between two grid points: a float32 score whose last set bit is g/2, several per
// cent of all scores) resolve the other way, once, after which the two trajectories are an even multiple
// apart.  So the chunk's effect is a function of the PARITY of the start value on its grid: every chunk is
// therefore executed on two trajectories, the main one from the given start state and a shadow one from
// L + ulp(L) that follows the main one's decisions, and a prediction for start value L0 + m g uses the main
// samples for even m and the shadow samples (shifted by (m - 1) g) for odd m.  Without this the predicted
// start states are off by one ulp after every other chunk and the fixed point is only reached chunk by chunk.
struct ChunkSummary {
  Samples m, s;     // main / shadow trajectory
  int min_st;       // st of the minT run
  int flags;        // 1: a reset happened, 2: carried run ended here, 4: minT.. valid,
                    // 8: the run in progress at the chunk end started inside the chunk
};

// spacing of the doubles at |x|; 0 for values so small that every sum of scores is exact
DGRP_HD double ulp_of(double x) {
  union V { double d; unsigned long long u; } v;
  v.d = x;
  const int e = (int)((v.u >> 52) & 0x7ffu);
  if (e <= 64 || e == 0x7ff) return 0.0;
  v.u = (unsigned long long)(e - 52) << 52;
  return v.d;
}
// is d an odd multiple of the grid spacing g (a power of two, so d / g is exact)?
DGRP_HD bool odd_multiple(double d, double g) {
  if (g == 0.0 || d == 0.0) return false;
  const double m = d / g;
  if (!(m > -9007199254740992.0 && m < 9007199254740992.0)) return false;
  const double h = m * 0.5;
  return m == (double)(long long)m && h != (double)(long long)h;
}
// the trajectory of `sum` (recorded from start value x_inL) that predicts the execution from yL, and the
// shift to apply to its samples
struct Pick {
  const Samples *p;
  double dd;
};
DGRP_HD Pick pick_trajectory(const ChunkSummary &sum, double x_inL, double yL) {
  const double d = yL - x_inL, g = ulp_of(x_inL);
  Pick r;
  if (odd_multiple(d, g)) { r.p = &sum.s; r.dd = d - g; }
  else { r.p = &sum.m; r.dd = d; }
  return r;
}

// Reduced scan over elements [b, e) of S (e <= n) from state `s`, as an object so that the caller decides where the
// scores come from (global memory directly, or a shared-memory tile loaded cooperatively: mss.cu): begin(), then
// block() for consecutive groups of up to BL scores (+1 look-ahead value), then end().  `first_ord` is the ordinal
// of the first run that STARTS in [b, e).  Writes the run records and the chunk summary.
constexpr int SCAN_BL = 8;

struct ChunkScan {
  ScanState &s;
  const RunTable &rt;
  ChunkSummary &sum;
  double xdrop;
  int n, next_ord;
  double Lb;      // the shadow trajectory (see ChunkSummary)
  double runb;    // its L before the run in progress, for a run that started in the chunk
  bool carried;   // the run in progress came from before the chunk
  bool prev_pos;
  bool store;     // write the run records (false: a speculative round that only wants the summary)

  DGRP_HD ChunkScan(ScanState &s_, const RunTable &rt_, ChunkSummary &sum_, double xdrop_, int n_, int first_ord,
                    bool prev_positive, bool store_ = true)
      : s(s_), rt(rt_), sum(sum_), xdrop(xdrop_), n(n_), next_ord(first_ord), prev_pos(prev_positive), store(store_) {
    sum.m.outL = sum.m.runL0 = sum.m.carryR = sum.m.minT = sum.m.maxAfter = sum.m.maxIn = 0.0;
    sum.s = sum.m;
    sum.min_st = -1; sum.flags = 0;
    Lb = s.L + ulp_of(s.L);
    runb = 0.0;
    carried = (s.flags & 2) != 0;
  }

  // scores blk[0..m) are elements i0.. of S; blk[m] is the look-ahead value S[i0 + m] (anything when i0 + m >= n)
  DGRP_HD void block(int i0, int m, const double *blk) {
    DGRP_UNROLL
    for (int j = 0; j < SCAN_BL; ++j) {
      if (j >= m) break;
      const int i = i0 + j;
      const double v = blk[j];
      if (v > 0) {
        if (!(s.flags & 2)) {
          s.flags |= 2;
          s.run_st = i; s.run_L0 = s.L; runb = Lb;
          // a "run" that begins at a speculative chunk start in the middle of a true run has no
          // ordinal; it only shapes the (to be discarded) speculative state
          const bool true_start = !prev_pos;
          s.run_ord = true_start ? next_ord++ : -1;
          carried = !true_start;
        }
        s.L += v;                                   // R = L + S[i]; R += S[k]  (mss.c:61-63)
        Lb += v;
        if (i + 1 == n || !(blk[j + 1] > 0)) {
          const double tL = s.run_L0;
          const int st = s.run_st;
          finish_run(s, i + 1, rt, store);
          const double R = s.L;
          if (carried) {
            sum.flags |= 2; sum.m.carryR = R; sum.s.carryR = Lb;
          } else {
            if (!(sum.flags & 4) || !(sum.m.minT < tL)) {
              sum.m.minT = tL; sum.s.minT = runb; sum.min_st = st; sum.m.maxAfter = R; sum.s.maxAfter = Lb;
            } else if (R > sum.m.maxAfter) { sum.m.maxAfter = R; sum.s.maxAfter = Lb; }
            if (!(sum.flags & 4) || R > sum.m.maxIn) { sum.m.maxIn = R; sum.s.maxIn = Lb; }
            sum.flags |= 4;
          }
          carried = false;
        }
        prev_pos = true;
      } else {
        if (xdrop > 0.0 && s.L + v + xdrop < s.maxv) {  // mss.c:89
          s.L = 0.0; s.maxv = kNegInf;              // mss.c:91 (the flush happens at the next run)
          s.botL = 0.0; s.bot_st = -1; s.flags |= 1;
          Lb = 0.0;                                 // from here on the frame is absolute: both trajectories agree
          sum.flags = 1;                            // summary restarts after a reset
        }
        s.L += v;                                   // mss.c:93
        Lb += v;
        prev_pos = false;
      }
    }
  }

  DGRP_HD void end() {
    if ((s.flags & 2) && !carried) sum.flags |= 8;
    sum.m.outL = s.L; sum.s.outL = Lb;
    sum.m.runL0 = s.run_L0; sum.s.runL0 = runb;
  }
};

template <typename ScoreT>
DGRP_HD void scan_chunk(const ScoreT *S, int n, double xdrop, int b, int e, int first_ord,
                        ScanState &s, const RunTable &rt, ChunkSummary &sum, bool store = true) {
  ChunkScan sc(s, rt, sum, xdrop, n, first_ord, b > 0 && ((double)S[b - 1] > 0), store);
  // Scores are consumed in blocks of BL values (+1 look-ahead) held in registers, so that the loads of
  // a block are independent of the sequential state machine and overlap each other.
  constexpr int BL = SCAN_BL;
  int i0 = b;
#if defined(__CUDA_ARCH__)
  // Device, float32 scores: a thread reads its own chunk, so every lane of a load instruction touches a different
  // line -- nine scalar loads per block were 32 sector accesses each and the L1 was the busiest unit of the kernel
  // (86 % of its peak, profiles/r02s2_hbm_kernels_ncu.md).  Two 128-bit loads per block instead, issued one block
  // ahead: the next block's first score is this block's look-ahead, so no ninth load is needed.
  if (sizeof(ScoreT) == 4 && (reinterpret_cast<uintptr_t>(S) & 15u) == 0 && (b & 3) == 0 && b + 2 * BL <= n &&
      b + BL <= e) {
    const float4 *F = reinterpret_cast<const float4 *>(S);
    float4 c0 = F[i0 >> 2], c1 = F[(i0 >> 2) + 1];
    for (; i0 + BL <= e && i0 + 2 * BL <= n; i0 += BL) {
      const float4 n0 = F[(i0 >> 2) + 2], n1 = F[(i0 >> 2) + 3];
      const double blk[BL + 1] = {(double)c0.x, (double)c0.y, (double)c0.z, (double)c0.w, (double)c1.x,
                                  (double)c1.y, (double)c1.z, (double)c1.w, (double)n0.x};
      sc.block(i0, BL, blk);
      c0 = n0; c1 = n1;
    }
  }
#endif
  for (; i0 < e; i0 += BL) {
    const int m = e - i0 < BL ? e - i0 : BL;
    double blk[BL + 1];
    DGRP_UNROLL
    for (int j = 0; j <= BL; ++j) {
      const int idx = i0 + j;
      blk[j] = (j <= m && idx < n) ? (double)S[idx] : 0.0;
    }
    sc.block(i0, m, blk);
  }
  sc.end();
}

// Predict the end state of a chunk for start state `y` from one known execution x_in -> x_out with
// summary `sum`.  Exact when the additions involved are exact or round the same way on the trajectory
// picked by the parity of the shift (see ChunkSummary); otherwise only a guess that the bitwise
// verification will reject.
DGRP_HD ScanState apply_summary(const ChunkSummary &sum, const ScanState &x_in,
                                const ScanState &x_out, const ScanState &y) {
  if (sum.flags & 1) return x_out;                  // after a reset the frame is absolute
  const double d = y.L - x_in.L;
  const Pick pk = pick_trajectory(sum, x_in.L, y.L);
  const Samples &t = *pk.p;
  const double dd = pk.dd;
  ScanState o = x_out;
  o.L = t.outL + dd;
  if (x_out.flags & 2) {
    if (sum.flags & 8) o.run_L0 = t.runL0 + dd;
    else if (y.flags & 2) { o.run_L0 = y.run_L0; o.run_st = y.run_st; o.run_ord = y.run_ord; }
    else o.run_L0 = x_out.run_L0 + d;
  }
  double bot = y.botL, mx = y.maxv;
  int bst = y.bot_st;
  bool empty = (y.flags & 1) != 0;
  if (sum.flags & 2) {                              // the carried run ended here
    const double tL = (y.flags & 2) ? y.run_L0 : x_in.run_L0 + d;
    const int st = (y.flags & 2) ? y.run_st : x_in.run_st;
    const double R = t.carryR + dd;
    if (empty || !(bot < tL)) { bot = tL; bst = st; mx = R; empty = false; }
    else if (R > mx) mx = R;
  }
  if (sum.flags & 4) {
    const double mT = t.minT + dd;
    if (empty || !(bot < mT)) { bot = mT; bst = sum.min_st; mx = t.maxAfter + dd; empty = false; }
    else if (t.maxIn + dd > mx) mx = t.maxIn + dd;
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
// Both trajectories are composed: A's main (shadow) end value picks, by the parity of its distance from
// B's recorded start, which of B's trajectories continues it.  Like apply_summary this is a prediction;
// every predicted start state is verified bitwise by the caller.
struct Composite {
  ChunkSummary sum;
  ScanState x_in, x_out;
};

// append the summary (minT, min_st, maxAfter, maxIn) of a later run list to p's (decisions on the main values)
DGRP_HD void fold_runs(ChunkSummary &p, double minT, double minTs, int min_st, double maxAfter, double maxAfters,
                       double maxIn, double maxIns) {
  if (!(p.flags & 4) || !(p.m.minT < minT)) {
    p.m.minT = minT; p.s.minT = minTs; p.min_st = min_st; p.m.maxAfter = maxAfter; p.s.maxAfter = maxAfters;
  } else if (maxIn > p.m.maxAfter) { p.m.maxAfter = maxIn; p.s.maxAfter = maxIns; }
  if (!(p.flags & 4) || maxIn > p.m.maxIn) { p.m.maxIn = maxIn; p.s.maxIn = maxIns; }
  p.flags |= 4;
}

DGRP_HD Composite compose(const Composite &A, const Composite &B) {
  Composite C;
  C.x_in = A.x_in;
  C.x_out = apply_summary(B.sum, B.x_in, B.x_out, A.x_out);
  ChunkSummary c = A.sum;
  if ((A.sum.flags & 1) || (B.sum.flags & 1)) {   // a reset inside: the end state is absolute
    c.flags = 1;
    c.m.outL = c.s.outL = C.x_out.L;
    C.sum = c;
    return C;
  }
  // B's recorded frame -> the frames of A's two trajectories
  const Pick pm = pick_trajectory(B.sum, B.x_in.L, A.sum.m.outL);
  const Pick ps = pick_trajectory(B.sum, B.x_in.L, A.sum.s.outL);
  c.flags = A.sum.flags & (2 | 4);
  const bool a_in_run = (A.x_out.flags & 2) != 0, a_run_inside = (A.sum.flags & 8) != 0;
  bool open_inside = false;                       // run in progress at the end started inside A+B
  if (a_in_run) {
    if (B.sum.flags & 2) {                        // the run crossing the boundary ends in B
      const double R = pm.p->carryR + pm.dd, Rs = ps.p->carryR + ps.dd;
      if (a_run_inside) fold_runs(c, A.sum.m.runL0, A.sum.s.runL0, A.x_out.run_st, R, Rs, R, Rs);
      else { c.flags |= 2; c.m.carryR = R; c.s.carryR = Rs; }   // it was carried into A as well
    } else if (a_run_inside) {
      open_inside = true;                         // still running at the end of B: runL0 stays A's
    }
  }
  if (B.sum.flags & 4)
    fold_runs(c, pm.p->minT + pm.dd, ps.p->minT + ps.dd, B.sum.min_st, pm.p->maxAfter + pm.dd,
              ps.p->maxAfter + ps.dd, pm.p->maxIn + pm.dd, ps.p->maxIn + ps.dd);
  if (B.sum.flags & 8) {
    open_inside = true;
    c.m.runL0 = pm.p->runL0 + pm.dd; c.s.runL0 = ps.p->runL0 + ps.dd;
  }
  if (open_inside) c.flags |= 8;
  c.m.outL = pm.p->outL + pm.dd; c.s.outL = ps.p->outL + ps.dd;
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
