// K2-K6, tcgen05 form: the recurrent product h.R of both GRU passes on the 5th-generation tensor
// cores, attention + FF + softmax in a second phase of the same persistent kernel, the max-vote as
// plain stores + a gather pass (vote.cu).  forward.cu holds the fp32 FFMA form of the same math.
//
// Precision.  The parity bar (>= 99.99 % identical labels against a float32 CPU run, on random-init
// weights whose class margins are ~1e-4) needs fp32-faithful pre-activations through a 150..512 step
// recurrence; plain bf16 misses it by orders of magnitude.  Both operands are
// therefore split into pieces and the product is formed from the piece products that matter, each a
// kind::f16 tcgen05.mma accumulating in fp32 into the same TMEM tile, smallest terms first:
//   NP = 2 (default)  two fp16 pieces of the scaled value (state x 2^8, weights x 2^shift so that the
//                     low pieces stay normal): hi.hi + hi.lo + lo.hi, relative error ~2^-21 per product;
//   NP = 3            three bf16 pieces (8+8+8 bits = the fp32 significand), the six products of weight
//                     >= 2^-16: hi.hi + hi.mid + mid.hi + mid.mid + hi.lo + lo.hi.
// Both measure the same distance from the float64 oracle as the fp32 kernel (tools/accuracy_study.py).
// The gates, sigmoid/tanh and the state update stay in fp32 registers.
//
// One CTA = 20 warps, persistent, one per SM; a unit of work = two tiles of 64 windows x 2 directions
// (M = 128 rows; row r = window r/2, direction r%2) in flight:
//   warps 0..15  gate warps: TMEM lane quadrant w % 4 (rows 32*(w%4) ..+31, one row per lane), unit
//                quarter w / 4; they wait for the tile's "done" mbarrier, read the accumulator
//                (tcgen05.ld), add the input projection (a table row: x_t is one-hot), apply the gates,
//                write the new state as operand pieces into A (shared memory), fence, arrive on the
//                tile's "ready" mbarrier, and only then store the step's scratch;
//   warp 16      MMA issuer: waits for "ready" (with back-off), issues the step's NP(NP+1)/2 x UP/16
//                tcgen05.mma (M128, N = 3*UP + 16, K16) and commits them to "done".
//                (Letting the last-arriving gate warp issue instead was measured 14 % slower.)
//   warps 17..19 exit after setmaxnreg: the CTA launches with 96 registers per thread, the issuer's
//                warpgroup drops to 24 and the gate warps rise to 112 (their loop spills at 96).
// While the tensor core multiplies tile X's new state by R, the gate warps work on tile Y.  A unit starts
// with a priming MMA round on an all-zero A operand (h[-1] = 0; two rounds when T is even), so the step
// loop has no special case and every unit runs an even number of rounds (barrier parities are a
// function of the step).  The z, r (h) columns of R, the input table and the biases are pre-scaled by
// -log2(e) (2 log2(e)) so that the accumulator feeds ex2 directly; reciprocals are shared four ways.
//   TMEM   2 x [128 lanes x (3*UP + 16) columns] fp32 accumulators: z | r | h gate blocks, then 16
//          projection columns h.(K/2) (the FF layer's two halves): avg[t].K = h_fwd[t].K/2 + h_rc[t].K/2
//          comes out of the same MMA and the second phase needs no FFMA over the units
//   smem   B = [R | K/2]^T pieces, [NP][(3*UP+16) x UP], K-major core matrices, resident
//          A = state pieces, [2 tiles][NP][128 x UP], rewritten every step by the gate warps
//          input table [10 rows][3*UP + 4] (A,C,G,T of either direction 4 banks apart, then 'N')
//          staged bases of the unit's tiles as table rows, [2 tiles][forward | reversed+complemented]
//          scores of one pass of the second phase, [wpp windows][T]
// Operand layout (no swizzle, K-major): 8-row x 16-byte core matrices, 128 B each; core matrices
// adjacent along K are 128 B apart (leading byte offset), along M/N they are UP/8 * 128 B apart
// (stride byte offset); one MMA consumes K = 16 = two core matrices.
//
// Scratch in HBM (per CTA, rewritten by every unit; the attention query avg[T-1] is only known after
// the last step, so the scores need a second pass over all avg[t]):
//   sum    ST[2][64/wpp][T][UP/8][wpp][8]  h_fwd[t] + h_rc[t] (= 2 avg[t], one lane shuffle) -- feeds ONLY the
//                                 additive-attention scores; ST = half by default (|sum| < 2; a
//                                 probability moves by ~1e-7 on random-init weights, <= 3e-5 on sharp
//                                 ones), float with forward_sum16 = 0
//   q      f32[2][64][UP]         avg[T-1] in full precision (the query)
//   proj   f32[2][64][T][16]      avg[t].K, both FF halves (the logits come from these, full precision)
// = 192 B per (window, t), written once and read once.  All CTAs run the same amount of work, so without
// care their second phases coincide and saturate HBM while the recurrence phases leave it idle: CTA b
// starts with (b mod 4) single-tile units, which shifts its phase by ~3/4 of a period each.  By the time
// the scratch is read back it has left the L2, and one SM's demand loads sustain only ~17 B/clk against
// HBM latency, so the second phase asks the L2 for it a few rows ahead (cp.async.bulk.prefetch.L2).
// The window probabilities go to [window][T][C] with plain stores; vote_gather_kernel max-merges them.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "forward_common.cuh"

namespace dgrp {

constexpr int TC_GATE_WARPS = 16;
constexpr int TC_THREADS = (TC_GATE_WARPS + 4) * 32;   // + one warpgroup: the issuer warp and three idle warps
constexpr int TC_GATE_REGS = 112, TC_AUX_REGS = 24;     // setmaxnreg split: the inc (16 warps x 16) must fit into what the dec releases (4 warps x 72)
constexpr float kNegLog2e = -1.4426950408889634f;   // z, r columns are pre-scaled: ex2(arg) = e^{-x}
constexpr float kTwoLog2e = 2.8853900817779268f;    // h columns are pre-scaled:    ex2(arg) = e^{2x}

#ifdef DGRP_TC_TRACE
// Developer instrumentation (not compiled into the product): cycles of CTA 0's gate warps per segment
// of the step loop, [warp][tile slot][segment]: 0 wait for "done", 1 TMEM loads, 2 gates + A operand
// stores, 3 fences + arrive, 4 scratch stores + loop; [16][.][5..6] the issuer's wait / issue.
__device__ unsigned long long g_tc_trace[17][2][8];
__shared__ unsigned int s_tc_trace[17][2][8];   // accumulated with fire-and-forget shared atomics
#define TC_TRACE_DECL unsigned int tr_t = (unsigned int)clock64()
#define TC_TRACE(seg)                                                         \
  do {                                                                        \
    const unsigned int now__ = (unsigned int)clock64();                       \
    if (lane == 0) atomicAdd(&s_tc_trace[warp][s][seg], now__ - tr_t);        \
    tr_t = now__;                                                             \
  } while (0)
#define TC_TRACE2(seg)                                                        \
  do {                                                                        \
    const unsigned int now__ = (unsigned int)clock64();                       \
    if (lane == 0) atomicAdd(&s_tc_trace2[warp][seg], now__ - tr_t);          \
    tr_t = now__;                                                             \
  } while (0)
__shared__ unsigned int s_tc_trace2[16][8];   // second phase: 0 scores, 1 barrier, 2 softmax over t, 3 vote, 4 barrier
__device__ unsigned long long g_tc_trace2[16][8];
#else
#define TC_TRACE_DECL
#define TC_TRACE(seg)
#define TC_TRACE2(seg)
#endif

template <int UP, int NP>
struct TCfg {
  static constexpr int NG = 3 * UP;            // gate columns (z | r | h)
  static constexpr int N = 3 * UP + 16;        // + 16 projection columns: h.[FF ctx half(5) pad(3) | FF avg half(5) pad(3)]
  static constexpr int ROWS = 128, WT = 64;
  static constexpr int UPT = UP / 4;           // units per gate thread
  // NP = 2 folds the z / r input projection into the MMA: 16 more K columns hold one_hot(x_t) * 2^8 in A
  // (5 used) and, in B, the z / r rows of the input table plus the h gate's recurrent bias, so the gate
  // warps read only the h-gate row of the table (a quarter of their former shared-memory loads).
  static constexpr bool FOLD = NP == 2;
  static constexpr int KP = UP + (FOLD ? 16 : 0);   // K of the operands
  static constexpr int KC = KP / 8;            // core matrices along K
  static constexpr int SBO = KC * 128;         // bytes between 8-row groups
  static constexpr int A_BYTES = ROWS * KP * 2;  // one piece of one tile
  static constexpr int B_BYTES = N * KP * 2;     // one piece
  static constexpr int PSTRIDE = 3 * UP + 4;     // floats per code row of the input table
  static constexpr int TCOLS = N <= 64 ? 64 : (N <= 128 ? 128 : 256);  // TMEM columns per tile
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// the issuer's wait: back off between polls so that the spinning warp does not eat issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= 1ull << 46;  // descriptor version (sm_100)
  return d;         // layout_type = 0: no swizzle
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                 "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// named barrier over the gate warps only (the issuer warp is in its own loop)
__device__ __forceinline__ void gate_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(TC_GATE_WARPS * 32) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// (a, b) = hi + mid + lo in bf16 pieces; each output word packs a's piece (low half) and b's (high half).
// The residuals are formed with one packed FFMA2 (x - hi exactly, as the scalar subtraction would).
__device__ __forceinline__ void split3(float2 ab, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
  const float2 neg1 = make_float2(-1.0f, -1.0f);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(ab.x, ab.y);
  hi = *reinterpret_cast<uint32_t *>(&h2);
  const float2 r = __ffma2_rn(make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)), neg1, ab);
  __nv_bfloat162 m2 = __floats2bfloat162_rn(r.x, r.y);
  mid = *reinterpret_cast<uint32_t *>(&m2);
  const float2 q = __ffma2_rn(make_float2(__uint_as_float(mid << 16), __uint_as_float(mid & 0xffff0000u)), neg1, r);
  __nv_bfloat162 l2 = __floats2bfloat162_rn(q.x, q.y);
  lo = *reinterpret_cast<uint32_t *>(&l2);
}

// fp16 x2 form: (a, b) * 2^8 = hi + lo in half-precision pieces (11 + 11 significant bits; the scale keeps
// the low piece a normal number for every |h| > 5e-4 and exact-to-2^-33 below)
constexpr float kStateScale = 256.0f;
__device__ __forceinline__ void split2h(float2 ab, uint32_t &hi, uint32_t &lo) {
  const float2 x = __fmul2_rn(ab, make_float2(kStateScale, kStateScale));
  const __half2 h2 = __floats2half2_rn(x.x, x.y);
  hi = *reinterpret_cast<const uint32_t *>(&h2);
  const float2 hf = __half22float2(h2);
  const float2 r = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), x);
  const __half2 l2 = __floats2half2_rn(r.x, r.y);
  lo = *reinterpret_cast<const uint32_t *>(&l2);
}

// FOLD: the 16-byte core-matrix row that holds one_hot(x) * 2^8 (half precision, 0x5C00 = 256) for the
// input-table row `trow` (0..3 forward A,C,G,T; 4..7 the rc pass, already complemented; 8, 9 'N'):
// K position p = the base the table row stands for.
__device__ __forceinline__ uint4 onehot_row(int trow) {
  const int p = trow < 4 ? trow : (trow < 8 ? 7 - trow : 4);
  const uint32_t v = 0x5C00u << ((p & 1) * 16);
  const int j = p >> 1;
  return make_uint4(j == 0 ? v : 0u, j == 1 ? v : 0u, j == 2 ? v : 0u, 0u);
}

// GRU cell update for four units at once.  Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 issue one
// instruction per pair, with the same IEEE rounding as the scalar forms); the accumulator and table
// values arrive pre-scaled (z, r by -log2 e; h by 2 log2 e):
//   z = 1/(1+2^sz) = 1/A, r = 1/(1+2^sr) = 1/B, hh = tanh = 1 - 2/D with D = 2^arg + 1, h' = hh + z (h - hh).
// The MUFU pipe (16 lanes/clk/SM) is what bounds this kernel, so reciprocals are shared: the four
// sigmoid denominators of two units cost ONE rcp (1/(A0 B0 A1 B1), then multiplied back), and so
// do the four tanh denominators of the quad: 3 ex2 + 0.75 rcp per unit instead of 3 + 2.  The clamps
// keep the products finite: sz, sr <= 28.85 is x >= -20 (sigmoid error < 2.1e-9), arg <= 30 is
// tanh = 1 - 2/(2^30 + 1), which rounds to 1.
__device__ __forceinline__ void sigmoid_pair(float2 sz, float2 sr, float2 &z, float2 &r) {
  const float2 one = make_float2(1.0f, 1.0f);
  const float2 ez = make_float2(ex2_approx(fminf(sz.x, 28.85f)), ex2_approx(fminf(sz.y, 28.85f)));
  const float2 er = make_float2(ex2_approx(fminf(sr.x, 28.85f)), ex2_approx(fminf(sr.y, 28.85f)));
  const float2 a = __fadd2_rn(ez, one), b = __fadd2_rn(er, one);
  const float2 ab = __fmul2_rn(a, b);
  const float inv = rcp_approx(ab.x * ab.y);
  const float2 iab = make_float2(inv * ab.y, inv * ab.x);   // 1/(a.x b.x), 1/(a.y b.y)
  z = __fmul2_rn(iab, b);
  r = __fmul2_rn(iab, a);
}
// SCALED: the accumulator holds (state scale * weight scale) h.R and is multiplied by `us` on the way in
// (an FFMA2 in place of the FADD2, so the scaling is free).
template <bool SCALED>
__device__ __forceinline__ float2 acc_add(float2 acc, float2 x, float2 us) {
  return SCALED ? __ffma2_rn(acc, us, x) : __fadd2_rn(x, acc);
}
// FOLD: the z / r pre-activations already contain the input projection (one-hot K columns of the MMA).
template <bool SCALED, bool FOLD>
__device__ __forceinline__ void gru_cell4(const float4 xz, const float4 xr, const float4 xh, const float4 bh,
                                          const float *az, const float *ar, const float *ah, float *hp,
                                          float2 us, float2 &h01, float2 &h23) {
  const float2 one = make_float2(1.0f, 1.0f);
  float2 z0, r0, z1, r1;
  if (FOLD) {
    sigmoid_pair(__fmul2_rn(make_float2(az[0], az[1]), us), __fmul2_rn(make_float2(ar[0], ar[1]), us), z0, r0);
    sigmoid_pair(__fmul2_rn(make_float2(az[2], az[3]), us), __fmul2_rn(make_float2(ar[2], ar[3]), us), z1, r1);
  } else {
    sigmoid_pair(acc_add<SCALED>(make_float2(az[0], az[1]), make_float2(xz.x, xz.y), us),
                 acc_add<SCALED>(make_float2(ar[0], ar[1]), make_float2(xr.x, xr.y), us), z0, r0);
    sigmoid_pair(acc_add<SCALED>(make_float2(az[2], az[3]), make_float2(xz.z, xz.w), us),
                 acc_add<SCALED>(make_float2(ar[2], ar[3]), make_float2(xr.z, xr.w), us), z1, r1);
  }
  // FOLD: the h-gate accumulator already contains its recurrent bias (the one-hot K rows of B)
  const float2 g0 = FOLD ? __fmul2_rn(make_float2(ah[0], ah[1]), us)
                         : acc_add<SCALED>(make_float2(ah[0], ah[1]), make_float2(bh.x, bh.y), us);
  const float2 g1 = FOLD ? __fmul2_rn(make_float2(ah[2], ah[3]), us)
                         : acc_add<SCALED>(make_float2(ah[2], ah[3]), make_float2(bh.z, bh.w), us);
  const float2 t0 = __ffma2_rn(r0, g0, make_float2(xh.x, xh.y));
  const float2 t1 = __ffma2_rn(r1, g1, make_float2(xh.z, xh.w));
  const float2 d0 = __fadd2_rn(make_float2(ex2_approx(fminf(t0.x, 30.0f)), ex2_approx(fminf(t0.y, 30.0f))), one);
  const float2 d1 = __fadd2_rn(make_float2(ex2_approx(fminf(t1.x, 30.0f)), ex2_approx(fminf(t1.y, 30.0f))), one);
  const float2 m = __fmul2_rn(d0, d1);              // {d0.x d1.x, d0.y d1.y}
  const float inv = rcp_approx(m.x * m.y);
  const float2 j = make_float2(inv * m.y, inv * m.x);
  const float2 i0 = __fmul2_rn(j, d1), i1 = __fmul2_rn(j, d0);   // 1/d0, 1/d1
  const float2 m2 = make_float2(-2.0f, -2.0f), neg1 = make_float2(-1.0f, -1.0f);
  const float2 hh0 = __ffma2_rn(m2, i0, one), hh1 = __ffma2_rn(m2, i1, one);
  h01 = __ffma2_rn(z0, __ffma2_rn(hh0, neg1, make_float2(hp[0], hp[1])), hh0);   // hh + z (h - hh)
  h23 = __ffma2_rn(z1, __ffma2_rn(hh1, neg1, make_float2(hp[2], hp[3])), hh1);
  hp[0] = h01.x; hp[1] = h01.y; hp[2] = h23.x; hp[3] = h23.y;
}

// Second phase for one tile, in passes of WPP windows (WPP * T floats of scores fit in shared memory).
//   sum   [T][UP/8 chunks][64 windows][8 units]  h_fwd[t] + h_rc[t]; the layout the gate warps can
//         write with full 128-byte lines (a warp's 32 rows are 16 windows x 2 directions x 4 units)
//   q     [64][UP]       avg[T-1]
//   proj  [64][T][16]    avg[t].K: ctx half in 0..4, avg half in 8..12
// (a) scores: score[t] = sum_u scale[u] tanh(q[u] + avg[t][u]) = S - 2 sum_u scale[u] / D[t][u] with
//     D = e^{2(q+avg)} + 1 and S = sum_u scale[u] the same for every t, so the softmax over t only needs
//     the second term.  A warp takes 8 adjacent windows x a slice of t; a lane takes one window and every
//     fourth chunk, so each load instruction reads whole 128-byte lines; four denominators share one rcp.
// (b) one warp per window, lanes over t: softmax over t, logits, class softmax and the max-vote.
// Ask the L2 to fetch `bytes` (multiple of 16) from HBM: one instruction, no registers, no L1 miss entries.
// The second phase reads scratch that left the L2 long ago; a single SM's demand loads sustain only
// ~17 B/clk against HBM latency, but ~4x that against L2 hits.
__device__ __forceinline__ void l2_prefetch(const void *ptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
// One 32-byte sector in one request (LDG.256): the softmax / vote passes read one sector of a `proj` row
// per lane, every lane a different line, and the LSU resolves one line per cycle -- the request count,
// not the bytes, bounds those passes.  `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float *p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ float exp_fast(float x) {   // e^x through ex2.approx (2 ulp), e^{-inf} = 0
  return ex2_approx(x * 1.4426950408889634f);
}
__device__ __forceinline__ float inv4_dot(const float *d, const float *sc) {
  // sum_k sc[k] / d[k], k < 4, with one reciprocal (d[k] <= 2^30 + 1, so the product is finite)
  const float2 d01 = make_float2(d[0], d[1]), d23 = make_float2(d[2], d[3]);
  const float2 m = __fmul2_rn(d01, d23);
  const float inv = rcp_approx(m.x * m.y);
  const float2 j = make_float2(inv * m.y, inv * m.x);
  const float2 i01 = __fmul2_rn(j, d23), i23 = __fmul2_rn(j, d01);
  return fmaf(sc[0], i01.x, fmaf(sc[1], i01.y, fmaf(sc[2], i23.x, sc[3] * i23.y)));
}

template <int UP, int WT, int NWARPS, typename ST>
__device__ __forceinline__ void attention_vote_sum_tile(const FwdParams &p, const ST *sum, const float *qbuf,
                                                        const float *proj, int64_t w_tile0, int wpp,
                                                        const float *s_scale, float *s_score) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, C = p.C;
  constexpr int NCH = UP / 8;          // 8-unit chunks per avg row
  constexpr int CPL = NCH / 4;         // chunks per lane: cj, cj + 4, ...
  constexpr bool H16 = sizeof(ST) == 2;
  constexpr int NV = H16 ? 1 : 2;      // 16-byte loads per chunk
  constexpr int UNR = H16 ? 4 : 2;     // t's in flight per lane (scores)
  constexpr int PU = 11;               // rows in flight per lane (softmax / vote passes): all of T <= 352 at once
  TC_TRACE_DECL;
  for (int w0 = 0; w0 < WT; w0 += wpp) {
    if (w_tile0 + w0 >= p.w_end) break;
    if (p.attention) {
      const int ngrp = wpp >> 3, grp = warp % ngrp, sl = warp / ngrp, nsl = NWARPS / ngrp;
      const int wi = lane & 7, cj = lane >> 3;
      const int wl = w0 + grp * 8 + wi;
      const int t_begin = (int)((int64_t)T * sl / nsl), t_end = (int)((int64_t)T * (sl + 1) / nsl);
      float q2[CPL][8], sc2[CPL][8];   // 2 log2(e) q[u], -2 scale[u]
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int u0 = (cj + 4 * i) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          q2[i][j] = qbuf[(size_t)wl * UP + u0 + j] * kTwoLog2e;
          sc2[i][j] = -2.0f * s_scale[u0 + j];
        }
      }
      // the pass's windows are contiguous per (t, chunk): [pass][t][chunk][wpp windows][8 units]
      const ST *sum_pass = sum + (size_t)(w0 / wpp) * T * NCH * wpp * 8;
      const ST *base = sum_pass + ((size_t)cj * wpp + (grp * 8 + wi)) * 8;
      // one warp per t-slice keeps the L2 PF_AHEAD rows ahead of the demand loads, PF_BLOCK rows at a time
      // (a t-row of the tile is NCH * WT * 8 contiguous elements)
      constexpr int PF_BLOCK = 4, PF_AHEAD = 12;
      const uint32_t ROW_BYTES = (uint32_t)(NCH * wpp * 8 * sizeof(ST));
      if (grp == 0 && lane == 0) {
        const int n0 = min(PF_AHEAD, t_end - t_begin);
        l2_prefetch(sum_pass + (size_t)t_begin * NCH * wpp * 8, (uint32_t)n0 * ROW_BYTES);
      }
      for (int t0 = t_begin; t0 < t_end; t0 += UNR) {
        if (grp == 0 && lane == 0 && (t0 - t_begin) % PF_BLOCK == 0) {
          const int ta = t0 + PF_AHEAD;
          if (ta < t_end) l2_prefetch(sum_pass + (size_t)ta * NCH * wpp * 8, (uint32_t)min(PF_BLOCK, t_end - ta) * ROW_BYTES);
        }
        uint4 v[UNR][CPL][NV];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          const int t = min(t0 + k, t_end - 1);
#pragma unroll
          for (int i = 0; i < CPL; ++i)
#pragma unroll
            for (int n = 0; n < NV; ++n)
              v[k][i][n] = __ldcs(reinterpret_cast<const uint4 *>(base + ((size_t)t * NCH + 4 * i) * wpp * 8) + n);
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
          float sacc = 0.f;
#pragma unroll
          for (int i = 0; i < CPL; ++i) {
            float x[8];   // h_fwd[t] + h_rc[t] of 8 units
            if (H16) {
              const uint32_t wv[4] = {v[k][i][0].x, v[k][i][0].y, v[k][i][0].z, v[k][i][0].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&wv[j]));
                x[2 * j] = f.x; x[2 * j + 1] = f.y;
              }
            } else {
              const uint32_t wv[8] = {v[k][i][0].x, v[k][i][0].y, v[k][i][0].z, v[k][i][0].w,
                                      v[k][i][NV - 1].x, v[k][i][NV - 1].y, v[k][i][NV - 1].z, v[k][i][NV - 1].w};
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(wv[j]);
            }
            float d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)   // e^{2 (q + sum/2)} + 1 = 2^{2 log2e q + log2e sum} + 1
              d[j] = ex2_approx(fminf(fmaf(x[j], 1.4426950408889634f, q2[i][j]), 30.0f)) + 1.0f;
            sacc += inv4_dot(d, sc2[i]) + inv4_dot(d + 4, sc2[i] + 4);
          }
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
          sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
          if (cj == 0 && t0 + k < t_end) s_score[(size_t)(grp * 8 + wi) * T + t0 + k] = sacc;
        }
      }
    }
    if (lane == 0 && warp < wpp) l2_prefetch(proj + (size_t)(w0 + warp) * T * 16, (uint32_t)T * 64u);
    TC_TRACE2(0);
    asm volatile("bar.sync 1, %0;" ::"n"(NWARPS * 32) : "memory");
    TC_TRACE2(1);
    for (int wq = warp; wq < wpp; wq += NWARPS) {
      const int wl = w0 + wq;
      if (lane == 0 && wq + NWARPS < wpp)   // the next window of this warp
        l2_prefetch(proj + (size_t)(wl + NWARPS) * T * 16, (uint32_t)T * 64u);
      const int64_t w = w_tile0 + wl;
      if (w >= p.w_end) break;
      const float *pr = proj + (size_t)wl * T * 16;
      const float *sc = s_score + (size_t)wq * T;
      float ctxk[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.attention) {
        // softmax over t and ctx.K1 = sum_t a_t (avg[t].K1); four rows per lane in flight
        float m_run = -INFINITY, l_run = 0.f, cacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t0 = lane; t0 < T; t0 += 32 * PU) {
          float sv[PU], k1[PU][5];
#pragma unroll
          for (int k = 0; k < PU; ++k) {
            const int t = t0 + 32 * k;
            const bool ok = t < T;
            float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok) ldg256(pr + (size_t)t * 16, row);
#pragma unroll
            for (int c = 0; c < 5; ++c) k1[k][c] = row[c];
            sv[k] = ok ? sc[t] : -INFINITY;
          }
          float m_new = m_run;
#pragma unroll
          for (int k = 0; k < PU; ++k) m_new = fmaxf(m_new, sv[k]);
          if (m_new > -INFINITY) {
            const float corr = exp_fast(m_run - m_new);   // exp(-inf) = 0 on the first rows
            l_run *= corr;
#pragma unroll
            for (int c = 0; c < 5; ++c) cacc[c] *= corr;
#pragma unroll
            for (int k = 0; k < PU; ++k) {
              const float e = exp_fast(sv[k] - m_new);
              l_run += e;
#pragma unroll
              for (int c = 0; c < 5; ++c) cacc[c] = fmaf(e, k1[k][c], cacc[c]);
            }
            m_run = m_new;
          }
        }
        float m_all = m_run;
        for (int off = 16; off > 0; off >>= 1) m_all = fmaxf(m_all, __shfl_xor_sync(0xffffffffu, m_all, off));
        const float f = (m_run == -INFINITY) ? 0.f : exp_fast(m_run - m_all);
        float l = l_run * f;
#pragma unroll
        for (int c = 0; c < 5; ++c) cacc[c] *= f;
        for (int off = 16; off > 0; off >>= 1) {
          l += __shfl_xor_sync(0xffffffffu, l, off);
#pragma unroll
          for (int c = 0; c < 5; ++c) cacc[c] += __shfl_xor_sync(0xffffffffu, cacc[c], off);
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) ctxk[c] = cacc[c] / l;
      }
      TC_TRACE2(2);
      // logits[t] = ctx.K1 + avg[t].K2 + b ; softmax over classes ; vote
      const int64_t place = (w < p.full_windows ? w * (int64_t)p.step
                                                : p.tail_base + (w - p.full_windows) * (int64_t)p.step) -
                            p.pred_row0;
      float cb[5];
#pragma unroll
      for (int c = 0; c < 5; ++c) cb[c] = c < C ? ctxk[c] + p.ffb[c] : 0.f;
      for (int t0 = lane; t0 < T; t0 += 32 * PU) {
        float k2[PU][5];
#pragma unroll
        for (int k = 0; k < PU; ++k) {
          const int t = t0 + 32 * k;
          const bool ok = t < T;
          float row[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (ok) ldg256(pr + (size_t)t * 16 + 8, row);
#pragma unroll
          for (int c = 0; c < 5; ++c) k2[k][c] = row[c];
        }
#pragma unroll
        for (int k = 0; k < PU; ++k) {
          const int t = t0 + 32 * k;
          float lg[5], mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 5; ++c) {
            lg[c] = c < C ? k2[k][c] + cb[c] : -INFINITY;
            mx = fmaxf(mx, lg[c]);
          }
          float sum_e = 0.f;
#pragma unroll
          for (int c = 0; c < 5; ++c) { lg[c] = c < C ? exp_fast(lg[c] - mx) : 0.f; sum_e += lg[c]; }
          const float inv = 1.0f / sum_e;
          if (p.win_probs) {
            // plain stores of the window's probabilities; vote_gather_kernel max-merges them
            if (t < T) {
              float *dst = p.win_probs + ((size_t)(w - p.w_begin) * T + t) * C;
#pragma unroll
              for (int c = 0; c < 5; ++c)
                if (c < C) dst[c] = lg[c] * inv;
            }
          } else {
            const int64_t r = place + t;
            if (t < T && r >= 0 && r < p.pred_rows) {
              int *dst = reinterpret_cast<int *>(p.pred + (size_t)r * C);
#pragma unroll
              for (int c = 0; c < 5; ++c)
                if (c < C) atomicMax(dst + c, __float_as_int(lg[c] * inv));   // probs > 0
            }
          }
        }
      }
    }
    TC_TRACE2(3);
    asm volatile("bar.sync 1, %0;" ::"n"(NWARPS * 32) : "memory");   // s_score is rewritten by the next pass
    TC_TRACE2(4);
  }
}

// NP = 3: bf16 hi|mid|lo pieces, 6 products (fp32-faithful).  NP = 2: fp16 hi|lo pieces of the scaled
// operands, 3 products (hi.hi + hi.lo + lo.hi, relative error ~2^-21 per product): half the tensor
// work, half the operand traffic out of shared memory.
template <int UP, typename ST, int NP>
__global__ void __launch_bounds__(TC_THREADS, 1) gru_tc_attention_vote_kernel(const FwdParams p) {
  using K = TCfg<UP, NP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *s_B = smem_raw;                               // [3][B_BYTES]
  unsigned char *s_A = s_B + NP * K::B_BYTES;                  // [2][NP][A_BYTES]
  float *s_P = reinterpret_cast<float *>(s_A + 2 * NP * K::A_BYTES);  // [10][PSTRIDE]: rows 0..3 A,C,G,T of the forward
                                                               // pass, 4..7 of the rc pass (complemented), 8, 9
                                                               // 'N' of either: the eight common rows start 4
                                                               // banks apart, so a warp's mixed reads do not conflict
  float *s_bh = s_P + 10 * K::PSTRIDE;                         // [UP] recurrent bias of the h gate
  float *s_scale = s_bh + UP;                                  // [UP] attention scale
  float *s_score = s_scale + UP;                               // [wpp][T]
  uint8_t *s_codes = reinterpret_cast<uint8_t *>(s_score + (size_t)p.wpp * p.T);   // [2 tiles][fwd | rc][code_span] staged bases
  __shared__ __align__(8) unsigned long long s_ready[2], s_done[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = p.T;

  // ---- one-time setup -------------------------------------------------------------------------
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.Bsplit);
    uint4 *dst = reinterpret_cast<uint4 *>(s_B);
    for (int i = tid; i < NP * K::B_BYTES / 16; i += TC_THREADS) dst[i] = src[i];
    for (int i = tid; i < 10 * K::PSTRIDE; i += TC_THREADS) {
      const int row = i / K::PSTRIDE, j = i % K::PSTRIDE;
      // reverse complement: [3,2,1,0,4] (model.py:277)
      const int c = row >= 8 ? 4 : (row >= 4 ? 7 - row : row);
      // (x.W + b_in) + b_rec for z and r; the h gate keeps b_rec inside r * (h.R + b_rec)
      float v = 0.f;
      if (j < 2 * UP) v = (p.P[c * 3 * UP + j] + p.b1[j]) * kNegLog2e;
      else if (j < 3 * UP) v = p.P[c * 3 * UP + j] * kTwoLog2e;
      s_P[i] = v;
    }
    for (int i = tid; i < UP; i += TC_THREADS) {
      s_bh[i] = p.b1[2 * UP + i] * kTwoLog2e;
      s_scale[i] = (p.attention && i < p.U) ? p.scale[i] : 0.f;
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&s_tmem)),
                 "r"(2 * K::TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
#ifdef DGRP_TC_TRACE
  for (int i = tid; i < 17 * 2 * 8; i += TC_THREADS) (&s_tc_trace[0][0][0])[i] = 0u;
  for (int i = tid; i < 16 * 8; i += TC_THREADS) (&s_tc_trace2[0][0])[i] = 0u;
#endif
  if (tid == 0) {
    mbar_init(smem_u32(&s_ready[0]), TC_GATE_WARPS * 32);
    mbar_init(smem_u32(&s_ready[1]), TC_GATE_WARPS * 32);
    mbar_init(smem_u32(&s_done[0]), 1);
    mbar_init(smem_u32(&s_done[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  // 20 warps launch with 96 registers each; the issuer's warpgroup hands most of its share to the gate
  // warps, whose loop otherwise spills.  ptxas budgets registers per role only when each role's whole
  // code sits inside the setmaxnreg branch, so the two roles have their own copy of the unit loop and
  // meet only through the mbarriers; the gate warps synchronise among themselves on named barrier 1.
  const bool is_gate = warp < TC_GATE_WARPS;
  // Rounds per unit and tile slot: 1 priming round (2 when T is even) + T steps = an EVEN number, so that
  // every unit starts with both mbarriers of a slot at phase parity 0 and the parity of a wait is a
  // function of the step alone (a per-slot parity register was spilled and reloaded inside the step loop).
  const int extra = (T & 1) ? 0 : 1;
  const uint32_t bar_ready = smem_u32(&s_ready[0]), bar_done = smem_u32(&s_done[0]);
  const int64_t n_windows = p.w_end - p.w_begin;
  const int64_t n_tiles = (n_windows + K::WT - 1) / K::WT;
  // this CTA's contiguous tile range; the first `lead` units are single tiles (phase shift, see top)
  const int64_t tile_lo = n_tiles * blockIdx.x / gridDim.x, tile_hi = n_tiles * (blockIdx.x + 1) / gridDim.x;
  const int64_t my_tiles = tile_hi - tile_lo;   // a single-tile unit costs ~3/4 of a period for half the work
  const int lead = my_tiles >= 32 ? (int)(blockIdx.x & 3) : (my_tiles >= 16 ? (int)(blockIdx.x & 1) : 0);
  // instruction descriptor: D fp32, A/B bf16, both K-major, N = 3*UP + 16, M = 128
  const uint32_t ab_fmt = NP == 3 ? 1u : 0u;   // operand format: 1 = bf16, 0 = fp16
  const uint32_t idesc = (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((uint32_t)(K::N >> 3) << 17) |
                         ((uint32_t)(128 >> 4) << 24);

  if (!is_gate) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_AUX_REGS));
    if (warp != TC_GATE_WARPS) return;   // the three idle warps of the issuer's warpgroup
    int unit = 0;
    for (int64_t tile = tile_lo; tile < tile_hi; ++unit) {
      const int nt = (unit < lead || tile + 1 >= tile_hi) ? 1 : 2;
      const bool live1 = nt == 2;
      // ===================== MMA issuer: T + 1 rounds per tile (round 0 = priming on h = 0) =====
      TC_TRACE_DECL;
      for (int k = 0; k <= T + extra; ++k) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s == 1 && !live1) continue;
          TC_TRACE(6);
          mbar_wait_backoff(bar_ready + 8 * s, (uint32_t)(k & 1));
          tc_fence_after();
          TC_TRACE(5);
          if (lane == 0) {
            const uint32_t a0 = smem_u32(s_A + (size_t)s * NP * K::A_BYTES), b0 = smem_u32(s_B);
            const uint32_t d = tmem_base + (uint32_t)(s * K::TCOLS);
            // (A piece, B piece), smallest products first: lo.hi, hi.lo, mid.mid, mid.hi, hi.mid, hi.hi
            // (fp16 x2: lo.hi, hi.lo, hi.hi)
            const int pa[6] = {NP == 3 ? 2 : 1, 0, NP == 3 ? 1 : 0, 1, 0, 0};
            const int pb[6] = {0, NP == 3 ? 2 : 1, NP == 3 ? 1 : 0, 0, 1, 0};
            uint32_t acc = 0;
#pragma unroll
            for (int q = 0; q < (NP == 3 ? 6 : 3); ++q) {
              // the one-hot K chunk only exists in the hi piece of A (its lo piece is zero)
              const int nkc = UP / 16 + ((K::FOLD && pa[q] == 0) ? 1 : 0);
#pragma unroll
              for (int kc = 0; kc < UP / 16 + 1; ++kc) {
                if (kc >= nkc) break;
                const uint64_t ad = umma_desc(a0 + pa[q] * K::A_BYTES + kc * 256, 128, K::SBO);
                const uint64_t bd = umma_desc(b0 + pb[q] * K::B_BYTES + kc * 256, 128, K::SBO);
                umma_bf16(d, ad, bd, idesc, acc);
                acc = 1;
              }
            }
            umma_commit(bar_done + 8 * s);
          }
          __syncwarp();
        }
      }
      tile += nt;
    }
#ifdef DGRP_TC_TRACE
    __syncwarp();
    if (blockIdx.x == 0 && lane < 16) g_tc_trace[16][lane >> 3][lane & 7] += s_tc_trace[16][lane >> 3][lane & 7];
#endif
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_GATE_REGS));
  ST *sum0 = reinterpret_cast<ST *>(p.scratch) + (size_t)blockIdx.x * 2 * K::WT * T * UP;         // [2][WT][T][UP]
  float *proj0 = p.ff2 + (size_t)blockIdx.x * 2 * K::WT * T * 16;                                 // [2][WT][T][16]
  float *q0 = p.qbuf + (size_t)blockIdx.x * 2 * K::WT * UP;                                       // [2][WT][UP]
  const float2 us = make_float2(p.b_unscale, p.b_unscale);
  int unit = 0;
  for (int64_t tile = tile_lo; tile < tile_hi; ++unit) {
    const int nt = (unit < lead || tile + 1 >= tile_hi) ? 1 : 2;
    const bool live1 = nt == 2;
    {
    // ===================== gate warps =====================
    const int quad = warp & 3, uq = warp >> 2;
    const int row = quad * 32 + lane;        // row of the tile = TMEM lane
    const int wl = row >> 1, dir = row & 1;  // window in tile, direction
    const float *tblp = s_P + uq * K::UPT;
    const float *bhp = s_bh + uq * K::UPT;
    const uint32_t a_off = (uint32_t)((row >> 3) * K::SBO + (row & 7) * 16 + ((uq * K::UPT) >> 3) * 128);
    // this thread's constant part of the `sum` index: pass of its window, window within the pass, half
    const size_t sum_thr = ((size_t)(wl / p.wpp) * T * (UP / 8) * p.wpp + (size_t)(wl % p.wpp)) * 8 + (dir ? 4 : 0);
    const uint32_t oh_off = (uint32_t)((row >> 3) * K::SBO + (row & 7) * 16 + (UP / 8) * 128);   // FOLD: one-hot K chunk
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    // The bases of a tile are one contiguous span of 63 * step + T codes: staged in shared memory once
    // per unit, so that the step loop has no global load (an L2 miss there stalls the whole tile).  Two
    // copies, already translated to input-table rows: forward (rows 0..3, 8 for 'N') and reversed +
    // complemented for the rc pass (rows 4..7, 9), so that both directions read position base + t.
    // Offset of this row's first base in its copy of tile slot 0's span; slot 1 is 2 * code_span further.
    // The reversed copy is laid out against the FULL span (63 * step + T) so that the offset does not
    // depend on the slot; what a short last tile does not cover is filled with 'N' rows (the windows
    // past the end compute on them; their results are never used).
    const int full_span = (K::WT - 1) * p.step + T;
    const int cbase = dir * p.code_span + (dir ? full_span - wl * p.step - T : wl * p.step);
    {
      const int nthr = TC_GATE_WARPS * 32;
      for (int s = 0; s < nt; ++s) {
        const int64_t w_first = p.w_begin + (tile + s) * K::WT;
        int64_t w_last = w_first + K::WT - 1;
        w_last = w_last < p.w_end ? w_last : p.w_end - 1;
        const int span = (int)((w_last - w_first) * p.step) + T;
        const uint8_t *src = p.codes + (w_first * (int64_t)p.step - p.codes_base);
        uint8_t *fwd_copy = s_codes + (size_t)(2 * s) * p.code_span, *rc_copy = fwd_copy + p.code_span;
        for (int i = tid; i < full_span; i += nthr) {
          const int c = i < span ? src[i] : 4;
          fwd_copy[i] = (uint8_t)(c < 4 ? c : 8);
          rc_copy[full_span - 1 - i] = (uint8_t)(c < 4 ? c + 4 : 9);
        }
      }
      gate_bar_sync();
    }
    float hprev[2][K::UPT];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
#pragma unroll
      for (int j = 0; j < K::UPT; ++j) hprev[s][j] = 0.f;
      if (s == 1 && !live1) continue;
      // h[-1] = 0: zero this thread's slots of the A operand and let the issuer prime the accumulator
      unsigned char *a_tile = s_A + (size_t)s * NP * K::A_BYTES + a_off;
#pragma unroll
      for (int c8 = 0; c8 < K::UPT / 8; ++c8)
#pragma unroll
        for (int pc = 0; pc < NP; ++pc)
          *reinterpret_cast<uint4 *>(a_tile + pc * K::A_BYTES + c8 * 128) = make_uint4(0u, 0u, 0u, 0u);
      if (K::FOLD && uq == 0) {   // one_hot(x_0) in the hi piece's extra K chunk (its second half stays zero)
        unsigned char *oh = s_A + (size_t)s * NP * K::A_BYTES + oh_off;
        *reinterpret_cast<uint4 *>(oh) = onehot_row(s_codes[cbase + s * 2 * p.code_span]);
        *reinterpret_cast<uint4 *>(oh + 128) = make_uint4(0u, 0u, 0u, 0u);
      }
      tc_fence_before();
      fence_async_smem();
      mbar_arrive(bar_ready + 8 * s);
      if (extra) {   // second priming round on the same all-zero operand
        mbar_wait(bar_done + 8 * s, 0u);
        mbar_arrive(bar_ready + 8 * s);
      }
    }
    TC_TRACE_DECL;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (s == 1 && !live1) continue;   // uniform over the CTA
        const int trow = s_codes[cbase + s * 2 * p.code_span + t];   // input-table row of this step's base
        TC_TRACE(4);
        mbar_wait(bar_done + 8 * s, (uint32_t)((t + extra) & 1));
        tc_fence_after();
        TC_TRACE(0);
        const float *prow = tblp + trow * K::PSTRIDE;
        const uint32_t t_tile = t_lane + (uint32_t)(s * K::TCOLS);
        unsigned char *a_tile = s_A + (size_t)s * NP * K::A_BYTES + a_off;
        const size_t rt = (size_t)(s * K::WT + wl) * T + t;   // (window, t) row of the scratch arrays
        float pj[4];   // projection of h[t-1] (4 of the 16 extra columns per unit quarter)
        tmem_ld4(t_tile + (uint32_t)(K::NG + 4 * uq), pj);
        float2 hn2[K::UPT / 8][4];   // the new state of this thread's units
#pragma unroll
        for (int c8 = 0; c8 < K::UPT / 8; ++c8) {
          float az[8], ar[8], ah[8];
          {
            const uint32_t taddr = t_tile + (uint32_t)(uq * K::UPT + c8 * 8);
            tmem_ld8(taddr, az);
            tmem_ld8(taddr + UP, ar);
            tmem_ld8(taddr + 2 * UP, ah);
            tmem_ld_wait();
            if (c8 == 0) TC_TRACE(1);
          }
#pragma unroll
          for (int j4 = 0; j4 < 2; ++j4) {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 xz = K::FOLD ? zero4 : *reinterpret_cast<const float4 *>(prow + c8 * 8 + 4 * j4);
            const float4 xr = K::FOLD ? zero4 : *reinterpret_cast<const float4 *>(prow + UP + c8 * 8 + 4 * j4);
            const float4 xh = *reinterpret_cast<const float4 *>(prow + 2 * UP + c8 * 8 + 4 * j4);
            const float4 bh = K::FOLD ? zero4 : *reinterpret_cast<const float4 *>(bhp + c8 * 8 + 4 * j4);
            gru_cell4<NP == 2, K::FOLD>(xz, xr, xh, bh, az + 4 * j4, ar + 4 * j4, ah + 4 * j4,
                               &hprev[s][c8 * 8 + 4 * j4], us, hn2[c8][2 * j4], hn2[c8][2 * j4 + 1]);
          }
          // new state -> operand pieces in A (one 16-byte core-matrix row per piece)
          uint32_t hi[4], mid[4], lo[4];
          if (NP == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) split3(hn2[c8][j], hi[j], mid[j], lo[j]);
            *reinterpret_cast<uint4 *>(a_tile + c8 * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(a_tile + K::A_BYTES + c8 * 128) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
            *reinterpret_cast<uint4 *>(a_tile + 2 * K::A_BYTES + c8 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) split2h(hn2[c8][j], hi[j], lo[j]);
            *reinterpret_cast<uint4 *>(a_tile + c8 * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(a_tile + K::A_BYTES + c8 * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (K::FOLD && uq == 0 && t + 1 < T)   // the next step's base for the MMA's one-hot K columns
          *reinterpret_cast<uint4 *>(s_A + (size_t)s * NP * K::A_BYTES + oh_off) =
              onehot_row(s_codes[cbase + s * 2 * p.code_span + t + 1]);
        TC_TRACE(2);
        // Hand the new A operand to the tensor core (the last step's MMA only feeds the projection).
        // The release fence waits for this thread's outstanding stores, so it comes BEFORE the scratch
        // stores of this step: only the A operand writes are in flight here.
        tc_fence_before();
        fence_async_smem();
        mbar_arrive(bar_ready + 8 * s);
        TC_TRACE(3);
        {
          // avg[t-1].K = h_fwd.(K/2) + h_rc.(K/2), stored by the fwd lane
          float4 o;
          o.x = pj[0] + __shfl_xor_sync(0xffffffffu, pj[0], 1);
          o.y = pj[1] + __shfl_xor_sync(0xffffffffu, pj[1], 1);
          o.z = pj[2] + __shfl_xor_sync(0xffffffffu, pj[2], 1);
          o.w = pj[3] + __shfl_xor_sync(0xffffffffu, pj[3], 1);
          if (NP == 2) { o.x *= us.x; o.y *= us.x; o.z *= us.x; o.w *= us.x; }
          if (dir == 0 && t > 0) *reinterpret_cast<float4 *>(proj0 + (rt - 1) * 16 + 4 * uq) = o;
        }
        TC_TRACE(5);
        if (p.attention) {
#pragma unroll
          for (int c8 = 0; c8 < K::UPT / 8; ++c8) {
            // h_fwd[t] + h_rc[t]: the partner row is the neighbouring lane; the fwd lane stores
            // units u0..u0+3, the rc lane u0+4..u0+7 (u0 = first unit of this block)
            float2 sm2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const float2 send = dir ? hn2[c8][j] : hn2[c8][j + 2];
              const float2 mine = dir ? hn2[c8][j + 2] : hn2[c8][j];
              const float2 recv = make_float2(__shfl_xor_sync(0xffffffffu, send.x, 1),
                                              __shfl_xor_sync(0xffffffffu, send.y, 1));
              sm2[j] = __fadd2_rn(mine, recv);
            }
            const int u = uq * K::UPT + c8 * 8 + (dir ? 4 : 0);
            // [slot][pass][t][chunk][wpp windows][8 units]: the warp's 32 rows write 32 consecutive 4-unit groups
            ST *dst = sum0 + (size_t)s * K::WT * T * UP + sum_thr + ((size_t)t * (UP / 8) + (uq * (K::UPT / 8) + c8)) * p.wpp * 8;
            if (sizeof(ST) == 2) {
              const __half2 h0 = __floats2half2_rn(sm2[0].x, sm2[0].y), h1 = __floats2half2_rn(sm2[1].x, sm2[1].y);
              *reinterpret_cast<uint2 *>(dst) =
                  make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
            } else {
              *reinterpret_cast<float4 *>(dst) = make_float4(sm2[0].x, sm2[0].y, sm2[1].x, sm2[1].y);
            }
            if (t == T - 1)   // the query avg[T-1] in full precision
              *reinterpret_cast<float4 *>(q0 + (size_t)(s * K::WT + wl) * UP + u) =
                  make_float4(0.5f * sm2[0].x, 0.5f * sm2[0].y, 0.5f * sm2[1].x, 0.5f * sm2[1].y);
          }
        }
        TC_TRACE(6);
      }
    }
    // projection of the last state h[T-1]
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s == 1 && !live1) continue;
      mbar_wait(bar_done + 8 * s, (uint32_t)((T + extra) & 1));
      tc_fence_after();
      float pj[4];
      tmem_ld4(t_lane + (uint32_t)(s * K::TCOLS + K::NG + 4 * uq), pj);
      tmem_ld_wait();
      float4 o;
      o.x = pj[0] + __shfl_xor_sync(0xffffffffu, pj[0], 1);
      o.y = pj[1] + __shfl_xor_sync(0xffffffffu, pj[1], 1);
      o.z = pj[2] + __shfl_xor_sync(0xffffffffu, pj[2], 1);
      o.w = pj[3] + __shfl_xor_sync(0xffffffffu, pj[3], 1);
      if (NP == 2) { o.x *= us.x; o.y *= us.x; o.z *= us.x; o.w *= us.x; }
      if (dir == 0)
        *reinterpret_cast<float4 *>(proj0 + ((size_t)(s * K::WT + wl) * T + (T - 1)) * 16 + 4 * uq) = o;
    }
    tc_fence_before();
    }
    // ---- attention + FF + softmax + vote for the unit's tiles ------------------------------------
    __threadfence_block();
    gate_bar_sync();
#pragma unroll 1
    for (int s = 0; s < nt; ++s)
      attention_vote_sum_tile<UP, K::WT, TC_GATE_WARPS, ST>(
          p, sum0 + (size_t)s * K::WT * T * UP, q0 + (size_t)s * K::WT * UP,
          proj0 + (size_t)s * K::WT * T * 16, p.w_begin + (tile + s) * K::WT, p.wpp, s_scale, s_score);
    tile += nt;
  }

  // ---- teardown (the last "done" wait of every tile saw its MMAs complete) -----------------------
  tc_fence_before();
  gate_bar_sync();
#ifdef DGRP_TC_TRACE
  if (blockIdx.x == 0 && tid < 16 * 16) g_tc_trace[tid >> 4][(tid >> 3) & 1][tid & 7] += s_tc_trace[tid >> 4][(tid >> 3) & 1][tid & 7];
  if (blockIdx.x == 0 && tid < 16 * 8) g_tc_trace2[tid >> 3][tid & 7] += s_tc_trace2[tid >> 3][tid & 7];
#endif
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(2 * K::TCOLS)
                 : "memory");
  }
}

template <int UP, int NP>
static size_t tc_smem_bytes(int T, int wpp, int code_span) {
  using K = TCfg<UP, NP>;
  return (size_t)NP * K::B_BYTES + 2 * NP * K::A_BYTES +
         sizeof(float) * ((size_t)10 * K::PSTRIDE + 2 * UP + (size_t)wpp * T) + 4 * (size_t)code_span + 128;
}

template <int UP, typename ST, int NP>
static int launch_tc_t(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  using K = TCfg<UP, NP>;
  const int64_t n_windows = p.w_end - p.w_begin;
  if (n_windows <= 0) return DGRP_OK;
  // windows per pass of the second phase: as many score rows as shared memory holds
  // staged base codes of a tile: 63 * step + T bytes; very large steps have no tcgen05 form
  const int64_t span = (int64_t)(K::WT - 1) * p.step + p.T;
  if (span > 16384) return DGRP_E_UNSUPPORTED;
  p.code_span = (int)((span + 15) & ~(int64_t)15);
  int wpp = K::WT;
  while (wpp > 8 && tc_smem_bytes<UP, NP>(p.T, wpp, p.code_span) > 227 * 1024) wpp >>= 1;
  const size_t smem = tc_smem_bytes<UP, NP>(p.T, wpp, p.code_span);
  if (smem > 227 * 1024) return DGRP_E_UNSUPPORTED;   // caller falls back to the fp32 kernel
  p.wpp = wpp;
  auto kern = gru_tc_attention_vote_kernel<UP, ST, NP>;
  DGRP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (n_windows + K::WT - 1) / K::WT;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const int grid = (int)(n_pairs < c->sm_count ? n_pairs : c->sm_count);
  const size_t rows = (size_t)grid * 2 * K::WT;   // window slots in flight
  DGRP_CHECK(c->avg.reserve(rows * p.T * UP * sizeof(ST) + rows * UP * sizeof(float)));
  DGRP_CHECK(c->io_c.reserve(rows * p.T * 16 * sizeof(float)));
  p.scratch = c->avg.as<float>();
  p.qbuf = reinterpret_cast<float *>(c->avg.as<unsigned char>() + rows * p.T * UP * sizeof(ST));
  p.ff2 = c->io_c.as<float>();
  // the vote: window probabilities + gather pass (5 plain stores per window-step instead of 5 atomics,
  // whose issue rate bounded the second phase), when the [windows][T][C] buffer is affordable
  p.win_probs = nullptr;
  const size_t win_bytes = (size_t)n_windows * p.T * p.C * sizeof(float);
  if (c->forward_gather && win_bytes <= ((size_t)40 << 30)) {
    if (c->winprobs.reserve(win_bytes) == DGRP_OK) p.win_probs = c->winprobs.as<float>();
    else cudaGetLastError();   // out of memory: keep the atomic vote
  }
  kern<<<grid, TC_THREADS, smem, c->stream>>>(p);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  if (p.win_probs)
    DGRP_CHECK(launch_vote_gather(c, p.win_probs, p.w_begin, p.w_end, p.T, p.C, p.full_windows, p.tail_base,
                                  p.step, p.pred, p.pred_row0, p.pred_rows));
  return DGRP_OK;
}

// Returns DGRP_E_UNSUPPORTED when this model/shape has no tcgen05 form (the caller then uses the
// fp32 kernel): units > 64, or a window too long for the score buffer in shared memory.
#ifdef DGRP_TC_TRACE
extern "C" int dgrp_debug_tc_trace(unsigned long long *out, int reset) {
  if (reset) {
    static unsigned long long zero[17][2][8];
    return (int)cudaMemcpyToSymbol(g_tc_trace, zero, sizeof(zero));
  }
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 17 * 2 * 8);
}
extern "C" int dgrp_debug_tc_trace2(unsigned long long *out, int reset) {
  if (reset) {
    static unsigned long long zero[16][8];
    return (int)cudaMemcpyToSymbol(g_tc_trace2, zero, sizeof(zero));
  }
  return (int)cudaMemcpyFromSymbol(out, g_tc_trace2, sizeof(unsigned long long) * 16 * 8);
}
#endif

template <int UP>
static int launch_tc_up(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  if (c->forward_fp16x2)
    return c->forward_sum16 ? launch_tc_t<UP, __half, 2>(c, m, p) : launch_tc_t<UP, float, 2>(c, m, p);
  return c->forward_sum16 ? launch_tc_t<UP, __half, 3>(c, m, p) : launch_tc_t<UP, float, 3>(c, m, p);
}

int launch_forward_tc(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  if (!m->d_Bsplit || !m->d_Bsplit16) return DGRP_E_UNSUPPORTED;
  p.Bsplit = c->forward_fp16x2 ? m->d_Bsplit16 : m->d_Bsplit;
  p.b_unscale = c->forward_fp16x2 ? ldexpf(1.0f, -(8 + m->b16_shift)) : 1.0f;
  switch (m->UP) {
    case 32: return launch_tc_up<32>(c, m, p);
    case 64: return launch_tc_up<64>(c, m, p);
    default: return DGRP_E_UNSUPPORTED;
  }
}

}  // namespace dgrp
