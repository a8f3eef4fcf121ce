// fwdback_blk.cuh -- lane-blocked Forward / Backward engines: lane l keeps model positions l*Q+1 .. l*Q+Q of the
// current row in registers, so one row costs three boundary shuffles, one affine warp scan for the D chain and one
// warp sum -- instead of that per 32-position chunk (fwdback.cuh).  Same arithmetic (scaled odds ratios, SURVEY.md
// A.5), same special-state bookkeeping, so the two engines are interchangeable to within fp32 summation order.
// Transitions live in registers (Q <= 8) or in the warp's shared-memory slice (TSMEM, Q = 16/32).
// Full matrices use the blocked layout  full[((row*3 + state)*Q + q)*32 + lane], state 0=M 1=D 2=I.
#pragma once
#include "device_utils.cuh"
#include "engine.hpp"
#include "fwdback.cuh"

namespace ckm {

constexpr int BLK_LOOKAHEAD_MAXQ = 16;

template <int Q, bool TSMEM>
struct BlkModel {
  int M;
  const float *rfb;            // emissions [KPAD][Q][32] (+ lane already added)
  const float4 *tsm;           // TSMEM: warp's shared-memory copy [2][Q][32] (T0 plane, T1 plane: 16-byte lane stride, conflict-free LDS.128)
  float4 t0[TSMEM ? 1 : Q], t1[TSMEM ? 1 : Q];
  int lane;
  __device__ __forceinline__ float4 T0(int q) const { return TSMEM ? tsm[q * 32 + lane] : t0[TSMEM ? 0 : q]; }             // BM MM IM DM
  __device__ __forceinline__ float4 T1(int q) const { return TSMEM ? tsm[(Q + q) * 32 + lane] : t1[TSMEM ? 0 : q]; }       // MD MI II DD
};

template <int Q, bool TSMEM>
__device__ __forceinline__ void blk_model_load(BlkModel<Q, TSMEM> &bm, const ModelScalars &ms, const float4 *tfb, const float *rfb,
                                               float4 *tsm, int lane) {
  bm.M = ms.M; bm.lane = lane;
  bm.rfb = rfb + ms.blk_off * 32 * KPAD + lane;
  bm.tsm = tsm;
  const float4 *src = tfb + ms.blk_off * 32 * 2;
  if (TSMEM) {
    __syncwarp();
    for (int q = 0; q < Q; ++q) { tsm[q * 32 + lane] = __ldg(src + (q * 32 + lane) * 2); tsm[(Q + q) * 32 + lane] = __ldg(src + (q * 32 + lane) * 2 + 1); }
    __syncwarp();
  } else {
#pragma unroll
    for (int q = 0; q < Q; ++q) { bm.t0[TSMEM ? 0 : q] = __ldg(src + (q * 32 + lane) * 2); bm.t1[TSMEM ? 0 : q] = __ldg(src + (q * 32 + lane) * 2 + 1); }
  }
}

// Forward.  xmx: optional (L+1) x 6 special-state rows; full: optional blocked matrix.  Returns the score (nats).
// STORE_D = false: the D plane of the full matrix is not written (posterior decoding reads M and I only).
// The emission row of residue i+1 is fetched while row i is computed (one row of look-ahead in registers), so the
// table latency (L2 for the wide classes) is off the row-to-row critical path.
template <int Q, bool TSMEM, bool FULL, bool STORE_D = true>
__device__ __forceinline__ float forward_blk(const BlkModel<Q, TSMEM> &bm, const uint8_t *__restrict__ res, int L, const Specials sp,
                                             float *xmx, float *full) {
  const int lane = bm.lane;
  float Mx[Q], Ix[Q], Dx[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) { Mx[q] = 0.0f; Ix[q] = 0.0f; Dx[q] = 0.0f; }
  if (FULL) {
#pragma unroll
    for (int z = 0; z < 3 * Q; ++z) full[z * 32 + lane] = 0.0f;
  }
  float xE = 0.0f, xN = 1.0f, xJ = 0.0f, xB = sp.nmove, xC = 0.0f, totscale = 0.0f;
  if (xmx != nullptr && lane == 0) { xmx[X_E] = xE; xmx[X_N] = xN; xmx[X_J] = xJ; xmx[X_B] = xB; xmx[X_C] = xC; xmx[X_SCALE] = 1.0f; }
  constexpr bool LA = (Q <= BLK_LOOKAHEAD_MAXQ);      // the look-ahead row costs Q registers; the widest classes cannot afford it
  float ecur[Q];
  int xnext;
  if (LA) {
    const float *rp = bm.rfb + (size_t)((L >= 1) ? __ldg(res) : 0) * Q * 32;
#pragma unroll
    for (int q = 0; q < Q; ++q) ecur[q] = __ldg(rp + q * 32);
    xnext = (L >= 2) ? __ldg(res + 1) : 0;
  } else xnext = (L >= 1) ? __ldg(res) : 0;
  for (int i = 1; i <= L; ++i) {
    float enext[LA ? Q : 1];
    {
      const float *rp = bm.rfb + (size_t)xnext * Q * 32;
      if (LA) {
#pragma unroll
        for (int q = 0; q < Q; ++q) enext[LA ? q : 0] = __ldg(rp + q * 32);
        xnext = (i + 2 <= L) ? __ldg(res + i + 1) : 0;
      } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) ecur[q] = __ldg(rp + q * 32);
        xnext = (i + 1 <= L) ? __ldg(res + i) : 0;
      }
    }
    float pm_in = __shfl_up_sync(0xffffffffu, Mx[Q - 1], 1), pi_in = __shfl_up_sync(0xffffffffu, Ix[Q - 1], 1), pd_in = __shfl_up_sync(0xffffffffu, Dx[Q - 1], 1);
    if (lane == 0) { pm_in = 0.0f; pi_in = 0.0f; pd_in = 0.0f; }
    float md[Q], esum = 0.0f;
#pragma unroll
    for (int q = Q - 1; q >= 0; --q) {
      const float4 t0 = bm.T0(q), t1 = bm.T1(q);
      const float pm = (q > 0) ? Mx[q - 1] : pm_in, pi = (q > 0) ? Ix[q - 1] : pi_in, pd = (q > 0) ? Dx[q - 1] : pd_in;
      float sv = xB * t0.x;
      sv += pm * t0.y;
      sv += pi * t0.z;
      sv += pd * t0.w;
      sv *= ecur[q];
      const float nI = Mx[q] * t1.y + Ix[q] * t1.z;
      md[q] = sv * t1.x;
      Mx[q] = sv; Ix[q] = nI;
      esum += sv;
    }
    // D(k+1) = md(k) + D(k) tDD(k): compose my block, scan across lanes, then replay inside the block
    float Bb = 0.0f, Tb = 1.0f;
#pragma unroll
    for (int q = 0; q < Q; ++q) { const float tdd = bm.T1(q).w; Bb = md[q] + Bb * tdd; Tb *= tdd; }
    float Bs = Bb, Ts = Tb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float Bl = __shfl_up_sync(0xffffffffu, Bs, o), Tl = __shfl_up_sync(0xffffffffu, Ts, o);
      if (lane >= o) { Bs = Bs + Bl * Ts; Ts = Ts * Tl; }
    }
    float d = __shfl_up_sync(0xffffffffu, Bs, 1);
    if (lane == 0) d = 0.0f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const bool in = (lane * Q + q + 1) <= bm.M;
      Dx[q] = in ? d : 0.0f;
      esum += Dx[q];
      d = md[q] + d * bm.T1(q).w;
    }
    xE = warp_sum_float(esum);
    xN = xN * sp.nloop;
    xC = (xC * sp.nloop) + (xE * sp.emove);
    xJ = (xJ * sp.nloop) + (xE * sp.eloop);
    xB = (xJ * sp.nmove) + (xN * sp.nmove);
    float scale = 1.0f;
    if (xE > 1.0e4f) {
      scale = xE;
      const float inv = __fdiv_rn(1.0f, xE);
      xN = __fdiv_rn(xN, xE); xC = __fdiv_rn(xC, xE); xJ = __fdiv_rn(xJ, xE); xB = __fdiv_rn(xB, xE);
#pragma unroll
      for (int q = 0; q < Q; ++q) { Mx[q] *= inv; Dx[q] *= inv; Ix[q] *= inv; }
      totscale += (float)log((double)xE);
      xE = 1.0f;
    }
    if (FULL) {
      float *fr = full + (size_t)i * 3 * Q * 32 + lane;
#pragma unroll
      for (int q = 0; q < Q; ++q) { fr[q * 32] = Mx[q]; if (STORE_D) fr[(Q + q) * 32] = Dx[q]; fr[(2 * Q + q) * 32] = Ix[q]; }
    }
    if (xmx != nullptr && lane == 0) {
      float *xr = xmx + (size_t)i * X_NX;
      xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = scale;
    }
    if (LA) {
#pragma unroll
      for (int q = 0; q < Q; ++q) ecur[q] = enext[LA ? q : 0];
    }
  }
  return totscale + (float)log((double)(xC * sp.nmove));
}

// Backward with the Forward pass's per-row scale factors.  bxmx: (L+1) x 6 rows out.
// MODE 0: no matrix.  MODE 1: the Backward matrix is stored in `full` (blocked layout).  MODE 2: `full` holds the
// Forward matrix (M and I planes); row i is replaced IN PLACE by the products F(i,k) B(i,k) -- the posterior
// probabilities up to the per-row factor totr, which the consumers apply -- so no Backward matrix is ever written.
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int Q, bool TSMEM, int MODE>
__device__ __forceinline__ void backward_blk(const BlkModel<Q, TSMEM> &bm, const uint8_t *__restrict__ res, int L, const Specials sp,
                                             const float *fxmx, float *bxmx, float *full) {
  const int lane = bm.lane, M = bm.M;
  float Mx[Q], Ix[Q], Dx[Q];     // row i+1 on entry to row i
#pragma unroll
  for (int q = 0; q < Q; ++q) { Mx[q] = 0.0f; Ix[q] = 0.0f; Dx[q] = 0.0f; }
  float xC = 0.0f, xE = 0.0f, xJ = 0.0f, xN = 0.0f, xB = 0.0f;
  constexpr bool LA = (Q <= BLK_LOOKAHEAD_MAXQ);
  float ecur[Q];          // emission row of residue x_{i+1}; fetched one iteration ahead when registers allow
#pragma unroll
  for (int q = 0; q < Q; ++q) ecur[q] = 0.0f;
  int xnext = (L >= 1) ? __ldg(res + L - 1) : 0;
  for (int i = L; i >= 0; --i) {
    float enext[LA ? Q : 1];
    if (LA) {
      const float *rp = bm.rfb + (size_t)xnext * Q * 32;
#pragma unroll
      for (int q = 0; q < Q; ++q) enext[LA ? q : 0] = __ldg(rp + q * 32);
      xnext = (i >= 2) ? __ldg(res + i - 2) : 0;
    } else if (i < L) {
      const float *rp = bm.rfb + (size_t)xnext * Q * 32;      // xnext == res[i] here
#pragma unroll
      for (int q = 0; q < Q; ++q) ecur[q] = __ldg(rp + q * 32);
      xnext = (i >= 1) ? __ldg(res + i - 1) : 0;
    }
    if (MODE == 2 && i >= 3) {     // pull row i-2 of the Forward matrix towards L2 (M plane, I plane: Q lines each)
      const float *fr = full + (size_t)(i - 2) * 3 * Q * 32;
      if (lane < Q) { prefetch_l2(fr + lane * 32); prefetch_l2(fr + (2 * Q + lane) * 32); }
    }
    float em[Q];                                                                  // e(k, x_{i+1}) M(i+1, k)
    if (i == L) {
      xC = sp.nmove; xE = xC * sp.emove; xB = 0.0f; xJ = 0.0f; xN = 0.0f;
#pragma unroll
      for (int q = 0; q < Q; ++q) em[q] = 0.0f;
    } else {
      float part = 0.0f;
#pragma unroll
      for (int q = 0; q < Q; ++q) { em[q] = Mx[q] * ecur[q]; part += em[q] * bm.T0(q).x; }
      xB = warp_sum_float(part);
      xC = xC * sp.nloop;
      xJ = (xB * sp.nmove) + (xJ * sp.nloop);
      xN = (xB * sp.nmove) + (xN * sp.nloop);
      xE = (xC * sp.emove) + (xJ * sp.eloop);
    }
    const float s = (i >= 1) ? fxmx[(size_t)i * X_NX + X_SCALE] : 1.0f;
    if (i >= 1) {
      const float inv = (s > 1.0f) ? __fdiv_rn(1.0f, s) : 1.0f;
      // position k+1 of my last cell lives in lane+1's first cell
      float em_next = __shfl_down_sync(0xffffffffu, em[0], 1);
      float4 tn_next;       // transitions entering the first position of lane+1's block: MM(k), IM(k), DM(k) for my last k
      {
        const float4 mine = bm.T0(0);
        tn_next.x = 0.0f;
        tn_next.y = __shfl_down_sync(0xffffffffu, mine.y, 1); tn_next.z = __shfl_down_sync(0xffffffffu, mine.z, 1); tn_next.w = __shfl_down_sync(0xffffffffu, mine.w, 1);
      }
      if (lane == 31) { em_next = 0.0f; tn_next.y = 0.0f; tn_next.z = 0.0f; tn_next.w = 0.0f; }
      // D(k) = (xE + mnext tDM(k)) + D(k+1) tDD(k): reverse composite of my block, reverse scan, replay
      float cb[Q];       // xE + mnext*tdm per cell (0 outside the model)
      float Bb = 0.0f, Tb = 1.0f;
#pragma unroll
      for (int q = Q - 1; q >= 0; --q) {
        const int k = lane * Q + q + 1;
        const float mnext = (q < Q - 1) ? em[q + 1] : em_next;
        const float tdm = (q < Q - 1) ? bm.T0(q + 1).w : tn_next.w;
        const bool in = (k <= M);
        cb[q] = in ? (xE + ((k < M) ? mnext * tdm : 0.0f)) : 0.0f;
        const float tdd = in ? bm.T1(q).w : 0.0f;
        Bb = cb[q] + Bb * tdd; Tb *= tdd;
      }
      float Bs = Bb, Ts = Tb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float Br = __shfl_down_sync(0xffffffffu, Bs, o), Tr = __shfl_down_sync(0xffffffffu, Ts, o);
        if (lane + o < 32) { Bs = Bs + Br * Ts; Ts = Ts * Tr; }
      }
      float dnext = __shfl_down_sync(0xffffffffu, Bs, 1);       // D(i, first k of lane+1)
      if (lane == 31) dnext = 0.0f;
      float nM[Q], nI[Q], nD[Q];
#pragma unroll
      for (int q = Q - 1; q >= 0; --q) {
        const int k = lane * Q + q + 1;
        const bool in = (k <= M);
        const float4 t1 = bm.T1(q);
        const float mnext = (q < Q - 1) ? em[q + 1] : em_next;
        const float4 tn = (q < Q - 1) ? bm.T0(q + 1) : tn_next;
        const float tmm = (k < M) ? tn.y : 0.0f, tim = (k < M) ? tn.z : 0.0f;
        const float inext = (i < L) ? Ix[q] : 0.0f;
        const float dv = in ? (cb[q] + dnext * t1.w) : 0.0f;
        float mv = xE + mnext * tmm + inext * t1.y + dnext * t1.x;
        float iv = mnext * tim + inext * t1.z;
        if (!in) { mv = 0.0f; iv = 0.0f; }
        nM[q] = mv * inv; nI[q] = iv * inv; nD[q] = dv * inv;
        dnext = dv;
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) { Mx[q] = nM[q]; Ix[q] = nI[q]; Dx[q] = nD[q]; }
      if (s > 1.0f) { xE = __fdiv_rn(xE, s); xN = __fdiv_rn(xN, s); xJ = __fdiv_rn(xJ, s); xB = __fdiv_rn(xB, s); xC = __fdiv_rn(xC, s); }
    } else {
      xC = 0.0f; xJ = 0.0f; xE = 0.0f;
#pragma unroll
      for (int q = 0; q < Q; ++q) { Mx[q] = 0.0f; Ix[q] = 0.0f; Dx[q] = 0.0f; }
    }
    if (MODE == 1) {
      float *fr = full + (size_t)i * 3 * Q * 32 + lane;
#pragma unroll
      for (int q = 0; q < Q; ++q) { fr[q * 32] = Mx[q]; fr[(Q + q) * 32] = Dx[q]; fr[(2 * Q + q) * 32] = Ix[q]; }
    }
    if (MODE == 2 && i >= 1) {
      float *fr = full + (size_t)i * 3 * Q * 32 + lane;
      float fm[Q], fi[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) { fm[q] = fr[q * 32]; fi[q] = fr[(2 * Q + q) * 32]; }      // all loads first, then the stores
#pragma unroll
      for (int q = 0; q < Q; ++q) { fr[q * 32] = fm[q] * Mx[q]; fr[(2 * Q + q) * 32] = fi[q] * Ix[q]; }
    }
    if (bxmx != nullptr && lane == 0) {
      float *xr = bxmx + (size_t)i * X_NX;
      xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = s;
    }
    if (LA) {
#pragma unroll
      for (int q = 0; q < Q; ++q) ecur[q] = enext[LA ? q : 0];
    }
  }
}

}  // namespace ckm
