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


int launch_forward_tc(dgrp_ctx *c, dgrp_model *m, FwdParams &p);   // forward_tc.cu

template <int UP>
struct Cfg {
  static constexpr int UG = UP / 4;                   // unit groups (4 units each)
  static constexpr int RG = FWD_THREADS / UG;         // row groups
  static constexpr int ROWS = (UP >= 128) ? 64 : 128; // rows = windows x 2 directions
  static constexpr int RPT = ROWS / RG;               // rows per thread
  static constexpr int WPT = RPT / 2;                 // windows per thread
  static constexpr int WT = ROWS / 2;                 // windows per tile
  static constexpr int HS = UP + 4;                   // h row stride (floats)
  static constexpr bool R_SMEM = UP <= 64;            // recurrent weights resident in smem
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int UP, bool DENSE>
__global__ void __launch_bounds__(FWD_THREADS, 1) gru_attention_vote_kernel(const FwdParams p) {
  using K = Cfg<UP>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_R = reinterpret_cast<float *>(smem_raw);                      // [UP][3][UP] or unused
  float *s_h = s_R + (K::R_SMEM ? UP * 3 * UP : 0);                      // [2][ROWS][HS]
  float *s_P = s_h + 2 * K::ROWS * K::HS;                                // [5][3][UP] (P or Wk)
  float *s_b0 = s_P + 5 * 3 * UP;                                        // [3][UP] (dense mode)
  float *s_b1 = s_b0 + 3 * UP;                                           // [3][UP]
  float *s_att = s_b1 + 3 * UP;                                          // [UP][12]: K1[5] K2[5] scale 0
  float *s_q = s_att + UP * 12;                                          // [8 warps][UP]
  float *s_score = s_q + 8 * UP;                                         // [8 warps][T]

  const int tid = threadIdx.x;
  const int tu = tid % K::UG, tr = tid / K::UG;
  const int T = p.T, U = p.U;
  const int KU = (U + 3) & ~3;

  // ---- one-time: stage weights ----
  if (K::R_SMEM)
    for (int i = tid; i < UP * 3 * UP; i += FWD_THREADS) s_R[i] = p.Rp[i];
  for (int i = tid; i < 5 * 3 * UP; i += FWD_THREADS) s_P[i] = DENSE ? p.Wk[i] : p.P[i];
  for (int i = tid; i < 3 * UP; i += FWD_THREADS) { s_b0[i] = p.b0[i]; s_b1[i] = p.b1[i]; }
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
    float hprev[K::RPT][4];
#pragma unroll
    for (int i = 0; i < K::RPT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) hprev[i][j] = 0.f;
    __syncthreads();

    int cur = 0;
    for (int t = 0; t < T; ++t) {
      const float *hc = s_h + cur * K::ROWS * K::HS;
      float *hn = s_h + (cur ^ 1) * K::ROWS * K::HS;

      // input part (x_t . kernel + b_in): a table row for one-hot codes
      float xz[K::RPT][4], xr[K::RPT][4], xh[K::RPT][4];
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
          const float4 vz = *reinterpret_cast<const float4 *>(s_P + (code * 3 + 0) * UP + 4 * tu);
          const float4 vr = *reinterpret_cast<const float4 *>(s_P + (code * 3 + 1) * UP + 4 * tu);
          const float4 vh = *reinterpret_cast<const float4 *>(s_P + (code * 3 + 2) * UP + 4 * tu);
          xz[i][0] = vz.x; xz[i][1] = vz.y; xz[i][2] = vz.z; xz[i][3] = vz.w;
          xr[i][0] = vr.x; xr[i][1] = vr.y; xr[i][2] = vr.z; xr[i][3] = vr.w;
          xh[i][0] = vh.x; xh[i][1] = vh.y; xh[i][2] = vh.z; xh[i][3] = vh.w;
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
          for (int j = 0; j < 4; ++j) {
            float az = 0.f, ar = 0.f, ah = 0.f;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
              az = fmaf(xv[c], s_P[(c * 3 + 0) * UP + 4 * tu + j], az);
              ar = fmaf(xv[c], s_P[(c * 3 + 1) * UP + 4 * tu + j], ar);
              ah = fmaf(xv[c], s_P[(c * 3 + 2) * UP + 4 * tu + j], ah);
            }
            xz[i][j] = az + s_b0[0 * UP + 4 * tu + j];
            xr[i][j] = ar + s_b0[1 * UP + 4 * tu + j];
            xh[i][j] = ah + s_b0[2 * UP + 4 * tu + j];
          }
        }
      }

      // recurrent part h . R (fp32 FFMA, k ascending)
      float az[K::RPT][4], ar[K::RPT][4], ah[K::RPT][4];
#pragma unroll
      for (int i = 0; i < K::RPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { az[i][j] = 0.f; ar[i][j] = 0.f; ah[i][j] = 0.f; }
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
            float4 rz, rr, rh;
            if (K::R_SMEM) {
              rz = *reinterpret_cast<const float4 *>(s_R + ((k + kk) * 3 + 0) * UP + 4 * tu);
              rr = *reinterpret_cast<const float4 *>(s_R + ((k + kk) * 3 + 1) * UP + 4 * tu);
              rh = *reinterpret_cast<const float4 *>(s_R + ((k + kk) * 3 + 2) * UP + 4 * tu);
            } else {
              rz = __ldg(reinterpret_cast<const float4 *>(p.Rp + ((size_t)(k + kk) * 3 + 0) * UP + 4 * tu));
              rr = __ldg(reinterpret_cast<const float4 *>(p.Rp + ((size_t)(k + kk) * 3 + 1) * UP + 4 * tu));
              rh = __ldg(reinterpret_cast<const float4 *>(p.Rp + ((size_t)(k + kk) * 3 + 2) * UP + 4 * tu));
            }
#pragma unroll
            for (int i = 0; i < K::RPT; ++i) {
              const float h = hv[i][kk];
              az[i][0] = fmaf(h, rz.x, az[i][0]); az[i][1] = fmaf(h, rz.y, az[i][1]);
              az[i][2] = fmaf(h, rz.z, az[i][2]); az[i][3] = fmaf(h, rz.w, az[i][3]);
              ar[i][0] = fmaf(h, rr.x, ar[i][0]); ar[i][1] = fmaf(h, rr.y, ar[i][1]);
              ar[i][2] = fmaf(h, rr.z, ar[i][2]); ar[i][3] = fmaf(h, rr.w, ar[i][3]);
              ah[i][0] = fmaf(h, rh.x, ah[i][0]); ah[i][1] = fmaf(h, rh.y, ah[i][1]);
              ah[i][2] = fmaf(h, rh.z, ah[i][2]); ah[i][3] = fmaf(h, rh.w, ah[i][3]);
            }
          }
        }
      }

      // gates (Keras GRU, reset_after=True; order z, r, h): h' = z*h + (1-z)*tanh(x_h + r*(hR_h+b))
      const float4 bz = *reinterpret_cast<const float4 *>(s_b1 + 0 * UP + 4 * tu);
      const float4 br = *reinterpret_cast<const float4 *>(s_b1 + 1 * UP + 4 * tu);
      const float4 bh = *reinterpret_cast<const float4 *>(s_b1 + 2 * UP + 4 * tu);
      const float bzv[4] = {bz.x, bz.y, bz.z, bz.w};
      const float brv[4] = {br.x, br.y, br.z, br.w};
      const float bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
      for (int i = 0; i < K::RPT; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float z = sigmoid_f(xz[i][j] + (az[i][j] + bzv[j]));
          const float r = sigmoid_f(xr[i][j] + (ar[i][j] + brv[j]));
          const float hh = tanhf(xh[i][j] + r * (ah[i][j] + bhv[j]));
          hprev[i][j] = z * hprev[i][j] + (1.0f - z) * hh;
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

template <int UP>
static size_t fwd_smem_bytes(int T) {
  using K = Cfg<UP>;
  size_t f = (K::R_SMEM ? (size_t)UP * 3 * UP : 0) + 2 * (size_t)K::ROWS * K::HS + 5 * 3 * UP +
             3 * UP + 3 * UP + (size_t)UP * 12 + 8 * UP + 8 * (size_t)T;
  return f * sizeof(float);
}

template <int UP, bool DENSE>
static int launch_fwd_t(dgrp_ctx *c, dgrp_model *m, FwdParams &p) {
  using K = Cfg<UP>;
  const int64_t n_windows = p.w_end - p.w_begin;
  if (n_windows <= 0) return DGRP_OK;
  const size_t smem = fwd_smem_bytes<UP>(p.T);
  if (smem > 227 * 1024) {
    set_error("vecsize %d needs %zu bytes of shared memory (limit 232448)", p.T, smem);
    return DGRP_E_UNSUPPORTED;
  }
  auto kern = gru_attention_vote_kernel<UP, DENSE>;
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
  switch (m->UP) {
    case 32: return launch_fwd_t<32, DENSE>(c, m, p);
    case 64: return launch_fwd_t<64, DENSE>(c, m, p);
    case 128: return launch_fwd_t<128, DENSE>(c, m, p);
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
                     int64_t pred_row0, int64_t pred_rows) {
  FwdParams p = {};
  fill_model(p, m);
  p.codes = d_codes; p.codes_base = codes_base;
  p.w_begin = w_begin; p.w_end = w_end;
  p.step = pl.step; p.full_windows = pl.full_windows; p.tail_base = pl.tail_base;
  p.pred = d_pred; p.pred_row0 = pred_row0; p.pred_rows = pred_rows;
  c->forward_used_tc = 0;
  if (c->forward_tc) {
    const int rc = launch_forward_tc(c, m, p);
    if (rc != DGRP_E_UNSUPPORTED) {
      c->forward_used_tc = 1;
      return rc;
    }
  }
  return launch_fwd<false>(c, m, p);
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
