// K2-K6 (fp32 form): windows -> shared GRU both directions -> attention -> FF -> softmax -> max-vote.
//
// Replaces, for one tile of windows per CTA, what the reference runs per batch through Keras:
//   fetch_validation_batch (deepgrp/prediction.py:28-37)      windows = strided views of the codes
//   ReverseComplement      (deepgrp/model.py:277-279)         index arithmetic on the same codes
//   GRU x2, shared weights (deepgrp/model.py:225-229,309-310) persistent recurrence, both passes
//   Average, AdditiveAttention, RepeatVector/Concatenate, Dense FF, Softmax (model.py:311-329)
//   get_max                (deepgrp/maxcalc.c:10-24)          integer atomicMax on positive floats
//
// Layout in HBM
//   codes   u8[L]                 base codes 0..4 (A,C,G,T,other) of the trimmed record
//   scratch f32[grid][WT][T][UP]  avg[t] = (h_fwd[t] + h_rc[t]) / 2 of the CTA's current tile
//   ff2     f32[grid][WT][T][C]   avg[t] . FF_kernel[avg half] (second half of the concat)
//   pred    f32[L][C]             zero-initialised, windows merged by max
//
// This file is the fp32 CUDA-core form (the correctness anchor of SURVEY.md section 7 step 5):
// every product of the recurrence is an FFMA in fp32, so results differ from a float32 CPU run
// only by summation order and the last-ulp behaviour of expf/tanhf.
#include "forward_common.cuh"

namespace dgrp {

Placement make_placement(int64_t length, int T, int step, int batch_size, int compat) {
  Placement p;
  p.step = step;
  const int64_t W = length > T ? (length - T + step - 1) / step : 0;  // len(range(0, L-T, step))
  p.n_windows = W;
  p.full_windows = W;
  p.tail_base = 0;
  if (compat == DGRP_COMPAT_REFERENCE && batch_size > 0 && W % batch_size != 0) {
    // prediction.py:105: index = i * batch.shape[0] * step_size with the SHORT batch's size
    const int64_t nb = W / batch_size, b_last = W % batch_size;
    p.full_windows = nb * batch_size;
    p.tail_base = nb * b_last * (int64_t)step;
  }
  return p;
}


int launch_forward_tc(dgrp_ctx *c, dgrp_model *m, FwdParams &p);              // forward_tc.cu
int launch_forward_tcw(dgrp_ctx *c, dgrp_model *m, FwdParams &p, int which);  // forward_tcw.cu

// RNN: 0 = GRU (reset_after, gates z, r, h), 1 = LSTM (gates i, f, c, o; one bias)
template <int UP, int RNN = 0>
struct Cfg {
  static constexpr int G = RNN ? 4 : 3;               // gates
  static constexpr int UG = UP / 4;                   // unit groups (4 units each)
  static constexpr int RG = FWD_THREADS / UG;         // row groups
  static constexpr int ROWS = (UP >= 128 || (RNN && UP >= 64)) ? 64 : 128;  // rows = windows x 2 directions
  static constexpr int RPT = ROWS / RG;               // rows per thread
  static constexpr int WPT = RPT / 2;                 // windows per thread
  static constexpr int WT = ROWS / 2;                 // windows per tile
  static constexpr int HS = UP + 4;                   // h row stride (floats)
  static constexpr bool R_SMEM = UP <= 64;            // recurrent weights resident in smem
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int UP, bool DENSE, int RNN>
__global__ void __launch_bounds__(FWD_THREADS, 1) gru_attention_vote_kernel(const FwdParams p) {
  using K = Cfg<UP, RNN>;
  constexpr int G = K::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_R = reinterpret_cast<float *>(smem_raw);                      // [UP][G][UP] or unused
  float *s_h = s_R + (K::R_SMEM ? UP * G * UP : 0);                      // [2][ROWS][HS]
  float *s_P = s_h + 2 * K::ROWS * K::HS;                                // [5][G][UP] (P or Wk)
  float *s_b0 = s_P + 5 * G * UP;                                        // [G][UP] (dense mode)
  float *s_b1 = s_b0 + G * UP;                                           // [G][UP]
  float *s_att = s_b1 + G * UP;                                          // [UP][12]: K1[5] K2[5] scale 0
  float *s_q = s_att + UP * 12;                                          // [8 warps][UP]
  float *s_score = s_q + 8 * UP;                                         // [8 warps][T]

  const int tid = threadIdx.x;
  const int tu = tid % K::UG, tr = tid / K::UG;
  const int T = p.T, U = p.U;
  const int KU = (U + 3) & ~3;

  // ---- one-time: stage weights ----
  if (K::R_SMEM)
    for (int i = tid; i < UP * G * UP; i += FWD_THREADS) s_R[i] = p.Rp[i];
  for (int i = tid; i < 5 * G * UP; i += FWD_THREADS) s_P[i] = DENSE ? p.Wk[i] : p.P[i];
  for (int i = tid; i < G * UP; i += FWD_THREADS) { s_b0[i] = p.b0[i]; s_b1[i] = p.b1[i]; }
  stage_attention_table<UP>(p, s_att, tid);
  __syncthreads();

  float *scratch = p.scratch + (size_t)blockIdx.x * K::WT * T * UP;
  float *ff2 = p.ff2 + (size_t)blockIdx.x * K::WT * T * 5;

  const int64_t n_windows = p.w_end - p.w_begin;
  const int64_t n_tiles = (n_windows + K::WT - 1) / K::WT;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t w_tile0 = p.w_begin + tile * K::WT;

    // ---------------- phase 1: recurrence, both directions in lock step -----------------
    for (int i = tid; i < 2 * K::ROWS * K::HS; i += FWD_THREADS) s_h[i] = 0.f;

    // per-thread window bookkeeping
    int64_t wpos[K::WPT];   // record position of the window start (code mode) / window index
    bool wvalid[K::WPT];
#pragma unroll
    for (int q = 0; q < K::WPT; ++q) {
      const int64_t w = w_tile0 + tr * K::WPT + q;
      wvalid[q] = w < p.w_end;
      wpos[q] = DENSE ? w : (w * (int64_t)p.step - p.codes_base);
    }
    float hprev[K::RPT][4], cprev[RNN ? K::RPT : 1][4];   // cell state (LSTM only)
#pragma unroll
    for (int i = 0; i < K::RPT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { hprev[i][j] = 0.f; if (RNN) cprev[RNN ? i : 0][j] = 0.f; }
    __syncthreads();

    int cur = 0;
    for (int t = 0; t < T; ++t) {
      const float *hc = s_h + cur * K::ROWS * K::HS;
      float *hn = s_h + (cur ^ 1) * K::ROWS * K::HS;

      // input part (x_t . kernel + b_in): a table row for one-hot codes
      float xg[G][K::RPT][4];
#pragma unroll
      for (int i = 0; i < K::RPT; ++i) {
        const int q = i >> 1, dir = i & 1;
        if (!DENSE) {
          int code = 4;
          if (wvalid[q]) {
            const int64_t pos = wpos[q] + (dir ? (T - 1 - t) : t);
            code = p.codes[pos];
            if (dir && code < 4) code = 3 - code;   // complement A<->T, C<->G (model.py:233-237)
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float4 v = *reinterpret_cast<const float4 *>(s_P + (code * G + g) * UP + 4 * tu);
            xg[g][i][0] = v.x; xg[g][i][1] = v.y; xg[g][i][2] = v.z; xg[g][i][3] = v.w;
          }
        } else {
          float xv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          if (wvalid[q]) {
            const float *src = p.dense + ((size_t)wpos[q] * T + (dir ? (T - 1 - t) : t)) * 5;
#pragma unroll
            for (int c = 0; c < 5; ++c) xv[c] = src[c];
            if (dir) {  // gather [3,2,1,0,4]
              float a = xv[0], b = xv[1];
              xv[0] = xv[3]; xv[1] = xv[2]; xv[2] = b; xv[3] = a;
            }
          }
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float acc = 0.f;
#pragma unroll
              for (int c = 0; c < 5; ++c) acc = fmaf(xv[c], s_P[(c * G + g) * UP + 4 * tu + j], acc);
              xg[g][i][j] = acc + s_b0[g * UP + 4 * tu + j];
            }
        }
      }

      // recurrent part h . R (fp32 FFMA, k ascending)
      float ag[G][K::RPT][4];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < K::RPT; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) ag[g][i][j] = 0.f;
      if (t > 0) {
        for (int k = 0; k < KU; k += 4) {
          float hv[K::RPT][4];
#pragma unroll
          for (int i = 0; i < K::RPT; ++i) {
            const float4 v = *reinterpret_cast<const float4 *>(hc + (tr * K::RPT + i) * K::HS + k);
            hv[i][0] = v.x; hv[i][1] = v.y; hv[i][2] = v.z; hv[i][3] = v.w;
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
              float4 rw;
              if (K::R_SMEM) rw = *reinterpret_cast<const float4 *>(s_R + ((k + kk) * G + g) * UP + 4 * tu);
              else rw = __ldg(reinterpret_cast<const float4 *>(p.Rp + ((size_t)(k + kk) * G + g) * UP + 4 * tu));
#pragma unroll
              for (int i = 0; i < K::RPT; ++i) {
                const float h = hv[i][kk];
                ag[g][i][0] = fmaf(h, rw.x, ag[g][i][0]); ag[g][i][1] = fmaf(h, rw.y, ag[g][i][1]);
                ag[g][i][2] = fmaf(h, rw.z, ag[g][i][2]); ag[g][i][3] = fmaf(h, rw.w, ag[g][i][3]);
              }
            }
          }
        }
      }

      float bg[G][4];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 b4 = *reinterpret_cast<const float4 *>(s_b1 + g * UP + 4 * tu);
        bg[g][0] = b4.x; bg[g][1] = b4.y; bg[g][2] = b4.z; bg[g][3] = b4.w;
      }
#pragma unroll
      for (int i = 0; i < K::RPT; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (RNN == 0) {
            // Keras GRU, reset_after=True, gate order z, r, h: h' = z*h + (1-z)*tanh(x_h + r*(hR_h + b))
            const float z = sigmoid_f(xg[0][i][j] + (ag[0][i][j] + bg[0][j]));
            const float r = sigmoid_f(xg[1][i][j] + (ag[1][i][j] + bg[1][j]));
            const float hh = tanhf(xg[2][i][j] + r * (ag[2][i][j] + bg[2][j]));
            hprev[i][j] = z * hprev[i][j] + (1.0f - z) * hh;
          } else {
            // Keras LSTM, gate order i, f, c, o (one bias, already in the input part):
            // c' = sig(f)*c + sig(i)*tanh(c~); h' = sig(o)*tanh(c')
            const float gi = sigmoid_f(xg[0][i][j] + ag[0][i][j]);
            const float gf = sigmoid_f(xg[1][i][j] + ag[1][i][j]);
            const float gc = tanhf(xg[2][i][j] + ag[2][i][j]);
            const float go = sigmoid_f(xg[G - 1][i][j] + ag[G - 1][i][j]);
            const float c = gf * cprev[RNN ? i : 0][j] + gi * gc;
            cprev[RNN ? i : 0][j] = c;
            hprev[i][j] = go * tanhf(c);
          }
        }
        *reinterpret_cast<float4 *>(hn + (tr * K::RPT + i) * K::HS + 4 * tu) =
            make_float4(hprev[i][0], hprev[i][1], hprev[i][2], hprev[i][3]);
      }
      // avg[t] = (fwd[t] + rev[t]) / 2 -- the reverse pass is NOT re-reversed (model.py:312)
#pragma unroll
      for (int q = 0; q < K::WPT; ++q) {
        const int wl = tr * K::WPT + q;
        float4 a;
        a.x = (hprev[2 * q][0] + hprev[2 * q + 1][0]) * 0.5f;
        a.y = (hprev[2 * q][1] + hprev[2 * q + 1][1]) * 0.5f;
        a.z = (hprev[2 * q][2] + hprev[2 * q + 1][2]) * 0.5f;
        a.w = (hprev[2 * q][3] + hprev[2 * q + 1][3]) * 0.5f;
        *reinterpret_cast<float4 *>(scratch + ((size_t)wl * T + t) * UP + 4 * tu) = a;
      }
      __syncthreads();
      cur ^= 1;
    }
    // make this CTA's scratch writes visible to its other warps (global memory, same CTA)
    __threadfence_block();
    __syncthreads();

    attention_vote_tile<UP, DENSE, K::WT>(p, scratch, ff2, w_tile0, s_att, s_q, s_score);
    __syncthreads();
  }
}

template <int UP, int RNN>
static size_t fwd_smem_bytes(int T) {
  using K = Cfg<UP, RNN>;
  size_t f = (K::R_SMEM ? (size_t)UP * K::G * UP : 0) + 2 * (size_t)K::ROWS * K::HS + 5 * K::G * UP +
             K::G * UP + K::G * UP + (size_t)UP * 12 + 8 * UP + 8 * (size_t)T;
  return f * sizeof(float);
}

template <int UP, bool DENSE, int RNN>
static int launch_fwd_t(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  using K = Cfg<UP, RNN>;
  const int64_t n_windows = p.w_end - p.w_begin;
  if (n_windows <= 0) return DGRP_OK;
  const size_t smem = fwd_smem_bytes<UP, RNN>(p.T);
  if (smem > 227 * 1024) {
    set_error("vecsize %d needs %zu bytes of shared memory (limit 232448)", p.T, smem);
    return DGRP_E_UNSUPPORTED;
  }
  auto kern = gru_attention_vote_kernel<UP, DENSE, RNN>;
  DGRP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (n_windows + K::WT - 1) / K::WT;
  const int grid = (int)(n_tiles < c->sm_count ? n_tiles : c->sm_count);
  DGRP_CHECK(c->avg.reserve((size_t)grid * K::WT * p.T * UP * sizeof(float)));
  DGRP_CHECK(c->io_c.reserve((size_t)grid * K::WT * p.T * 5 * sizeof(float)));
  p.scratch = c->avg.as<float>();
  p.ff2 = c->io_c.as<float>();
  kern<<<grid, FWD_THREADS, smem, c->stream>>>(p);
  c->launches++;
  DGRP_CUDA(cudaGetLastError());
  return DGRP_OK;
}

template <bool DENSE>
static int launch_fwd(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  if (m->rnn == 1) {
    switch (m->UP) {
      case 32: return launch_fwd_t<32, DENSE, 1>(c, m, p);
      case 64: return launch_fwd_t<64, DENSE, 1>(c, m, p);
      case 128: return launch_fwd_t<128, DENSE, 1>(c, m, p);
      default: break;
    }
  }
  switch (m->UP) {
    case 32: return launch_fwd_t<32, DENSE, 0>(c, m, p);
    case 64: return launch_fwd_t<64, DENSE, 0>(c, m, p);
    case 128: return launch_fwd_t<128, DENSE, 0>(c, m, p);
    default:
      set_error("units=%d not supported by the CUDA forward (max 128)", m->U);
      return DGRP_E_UNSUPPORTED;
  }
}

static void fill_model(FwdParams &p, const dgrp_model *m) {
  p.T = m->T; p.U = m->U; p.C = m->C; p.attention = m->attention ? 1 : 0;
  p.P = m->d_P; p.Wk = m->d_Wk; p.b0 = m->d_b0; p.Rp = m->d_Rp; p.b1 = m->d_b1;
  p.scale = m->d_scale; p.ffk = m->d_ffk; p.ffb = m->d_ffb;
}

int run_forward_vote(dgrp_ctx *c, dgrp_model *m, const uint8_t *d_codes, int64_t codes_base,
                     int64_t w_begin, int64_t w_end, const Placement &pl, float *d_pred,
                     int64_t pred_row0, int64_t pred_rows, uint8_t *d_label, float *d_score, bool *fused,
                     int64_t w2_begin, int64_t w2_end) {
  if (fused) *fused = false;
  // A second window range (the displaced last batch of a position range, api.cu): the tcgen05 kernels with the
  // vote in shared memory run it in the same launch; every other route runs it as a call of its own afterwards.
  bool has2 = w2_end > w2_begin;
  if (has2 && w_end <= w_begin) { w_begin = w2_begin; w_end = w2_end; has2 = false; }
  auto second = [&]() -> int {
    return has2 ? run_forward_vote(c, m, d_codes, codes_base, w2_begin, w2_end, pl, d_pred, pred_row0, pred_rows) : DGRP_OK;
  };
  FwdParams p = {};
  fill_model(p, m);
  p.codes = d_codes; p.codes_base = codes_base;
  p.w_begin = w_begin; p.w_end = w_end;
  p.step = pl.step; p.full_windows = pl.full_windows; p.tail_base = pl.tail_base;
  p.pred = d_pred; p.pred_row0 = pred_row0; p.pred_rows = pred_rows;
  if (w_end <= w_begin) {   // no window in this range (a record no longer than the window; an empty placement family)
    if (fused && pred_rows > 0)   // the caller left the zero-fill to us
      DGRP_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)pred_rows * m->C * sizeof(float), c->stream));
    return DGRP_OK;               // (forward_used_tc keeps what the last real launch used)
  }
  c->forward_used_tc = 0;
  if (c->forward_tc) {
    // The tcgen05 kernels leave the window probabilities in a [windows][T][C] buffer that a gather pass
    // max-merges into pred; the windows run in slabs so that the buffer stays below forward_slab_bytes
    // (33.9 GB for a chr1-sized record otherwise).  The gather folds pred's current value, so slabs compose.
    auto dispatch = [&](FwdParams &q, int *used) -> int {
      int rc = DGRP_E_UNSUPPORTED;
      if (c->forward_wide == 0) { rc = launch_forward_tc(c, m, q); *used = 1; }
      if (rc == DGRP_E_UNSUPPORTED && c->forward_wide != 2) { rc = launch_forward_tcw(c, m, q, 1); *used = 2; }
      if (rc == DGRP_E_UNSUPPORTED) { rc = launch_forward_tcw(c, m, q, 2); *used = 3; }
      return rc;
    };
    // Will the kernel max-merge a tile's windows in shared memory (the usual case)?  Then no window probabilities
    // exist in HBM: one launch over all windows, the tile spans go straight into the zero-initialised pred.
    bool smem_vote = false;
    if (w_end > w_begin) {
      FwdParams q = p;
      q.query = 1;
      int used = 0;
      smem_vote = dispatch(q, &used) == DGRP_OK && q.smem_vote;
    }
    const int64_t per_win = (int64_t)m->T * m->C * (int64_t)sizeof(float);
    int64_t slab = c->forward_slab_bytes / per_win;
    slab = slab < 4096 ? 4096 : slab & ~(int64_t)4095;   // whole tiles of 64 windows, pairs of them per SM
    p.win_probs = nullptr;
    if (c->forward_gather && !smem_vote) {
      const int64_t nw = w_end - w_begin < slab ? w_end - w_begin : slab;
      if (nw > 0 && c->winprobs.reserve((size_t)(nw * per_win)) == DGRP_OK) p.win_probs = c->winprobs.as<float>();
      else cudaGetLastError();   // out of memory: the kernels vote with atomicMax instead
    }
    if (!p.win_probs) slab = w_end - w_begin;
    if (has2 && smem_vote && !fused) {   // one launch for both ranges (4 tiles of their own would hold the GPU for a whole
      p.w2_begin = w2_begin; p.w2_end = w2_end;   // unit time: 1.3 ms at the defaults)
      has2 = false;
    }
    // one slab holds every window of the record: the gather pass sees each row's final vote and can apply the
    // score transform at once (label + score out, the predictions never written)
    const bool fuse = fused && d_label && d_score && p.win_probs && w_end - w_begin <= slab && pred_row0 == 0 &&
                      c->forward_fuse_score;
    if (fuse && w_end <= w_begin) {   // no window at all (L <= T): every row is a never-covered row
      DGRP_CHECK(launch_vote_gather(c, p.win_probs, 0, 0, m->T, m->C, pl.full_windows, pl.tail_base, pl.step, d_pred, 0,
                                    pred_rows, d_label, d_score));
      *fused = true;
      return DGRP_OK;
    }
    if (!fuse && fused && pred_rows > 0)   // the caller left the zero-fill to us
      DGRP_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)pred_rows * m->C * sizeof(float), c->stream));
    for (int64_t w0 = w_begin; w0 < w_end; w0 += slab) {
      const int64_t w1 = w0 + slab < w_end ? w0 + slab : w_end;
      p.w_begin = w0; p.w_end = w1;
      int used = 0;
      int rc = dispatch(p, &used);
      if (rc == DGRP_E_UNSUPPORTED) {
        if (w0 != w_begin) { set_error("forward: tcgen05 form lost between slabs"); return DGRP_E_CUDA; }
        p.w_begin = w_begin; p.w_end = w_end; p.win_probs = nullptr;
        if (fuse && pred_rows > 0)
          DGRP_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)pred_rows * m->C * sizeof(float), c->stream));
        if (p.w2_end > p.w2_begin) { has2 = true; p.w2_begin = p.w2_end = 0; }   // (not reached: the query found a form)
        DGRP_CHECK(launch_fwd<false>(c, m, p));
        return second();
      }
      if (rc != DGRP_OK) return rc;
      c->forward_used_tc = used;
      if (fuse) {
        DGRP_CHECK(launch_vote_gather(c, p.win_probs, w0, w1, m->T, m->C, pl.full_windows, pl.tail_base, pl.step, d_pred,
                                      0, pred_rows, d_label, d_score));
        *fused = true;
      } else if (p.win_probs) {
        // rows the slab's placed windows cover (both placement families, prediction.py:105)
        int64_t lo = INT64_MAX, hi = INT64_MIN;
        const int64_t f0 = w0, f1 = w1 < pl.full_windows ? w1 : pl.full_windows;
        if (f0 < f1) { lo = f0 * pl.step; hi = (f1 - 1) * pl.step + m->T; }
        const int64_t t0 = w0 > pl.full_windows ? w0 : pl.full_windows;
        if (t0 < w1) {
          const int64_t a = pl.tail_base + (t0 - pl.full_windows) * pl.step;
          const int64_t b = pl.tail_base + (w1 - 1 - pl.full_windows) * pl.step + m->T;
          lo = a < lo ? a : lo; hi = b > hi ? b : hi;
        }
        lo = lo > pred_row0 ? lo : pred_row0;
        hi = hi < pred_row0 + pred_rows ? hi : pred_row0 + pred_rows;
        if (lo < hi)
          DGRP_CHECK(launch_vote_gather(c, p.win_probs, w0, w1, m->T, m->C, pl.full_windows, pl.tail_base, pl.step,
                                        d_pred + (size_t)(lo - pred_row0) * m->C, lo, hi - lo));
      }
    }
    return second();
  }
  if (fused && pred_rows > 0)
    DGRP_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)pred_rows * m->C * sizeof(float), c->stream));
  DGRP_CHECK(launch_fwd<false>(c, m, p));
  return second();
}

int run_forward_dense(dgrp_ctx *c, dgrp_model *m, const float *d_batch, int64_t nbatch,
                      float *d_probs) {
  FwdParams p = {};
  fill_model(p, m);
  p.dense = d_batch; p.probs_out = d_probs;
  p.w_begin = 0; p.w_end = nbatch; p.step = 0;
  p.full_windows = nbatch;
  return launch_fwd<true>(c, m, p);
}

}  // namespace dgrp
