// fwdback.cuh -- warp-per-pair Forward / Backward row engines in scaled odds-ratio space.
// Lane l owns model positions k = 32*c + l + 1.  The previous (Forward) or next (Backward) DP row lives in shared
// memory; the diagonal neighbour comes by warp shuffle; the in-row D->D chain D(k+1) = M(k) tMD(k) + D(k) tDD(k) is a
// warp scan over affine maps d -> B + d*T, carried from chunk to chunk.
// Optional outputs: the special-state columns of every row (xmx: E,N,J,B,C,scale) and the full M/D/I matrix in a
// planar layout  full[(i*3 + {0:M,1:D,2:I}) * Mpad + k].
#pragma once
#include "device_utils.cuh"
#include "engine.hpp"

namespace ckm {

enum { X_E = 0, X_N, X_J, X_B, X_C, X_SCALE, X_NX };

struct Specials { float nloop, nmove, eloop, emove; };

__device__ __forceinline__ Specials make_specials(int Lcfg, bool multihit) {
  Specials s;
  const float nj = multihit ? 1.0f : 0.0f;
  s.nmove = __fdiv_rn(__fadd_rn(2.0f, nj), __fadd_rn(__fadd_rn((float)Lcfg, 2.0f), nj));
  s.nloop = __fsub_rn(1.0f, s.nmove);
  s.eloop = multihit ? 0.5f : 0.0f;
  s.emove = multihit ? 0.5f : 1.0f;
  return s;
}

struct FwdModel {
  int M, Mpad;
  const float *rfv;        // [KPAD][Mpad] match odds ratios (0 for k = 0 and k > M)
  const float4 *tfv;       // [Mpad][2]: {BM, MM, IM, DM}, {MD, MI, II, DD}
};

// Forward.  rows: 3 shared-memory arrays of >= 32*ceil(M/32)+1 floats.  Returns the Forward score in nats.
// EXACT: evaluate the D chain and the E sum strictly in model order (one lane after another), reproducing a scalar
// left-to-right evaluation bit for bit; used where sampling decisions hang on the matrix (the trace ensemble).
template <bool FULL, bool EXACT = false>
__device__ __forceinline__ void forward_rows(const FwdModel &fm, const uint8_t *__restrict__ res, int L, const Specials sp,
                                             float *rowM, float *rowI, float *rowD, int lane,
                                             float *xmx, float *full, int /*unused*/, float *ret_sc) {
  const int M = fm.M, nchunk = (M + 31) >> 5;
  for (int k = lane; k < nchunk * 32 + 1; k += 32) { rowM[k] = 0.0f; rowI[k] = 0.0f; rowD[k] = 0.0f; }
  if (FULL) for (int k = lane; k < 3 * fm.Mpad; k += 32) full[k] = 0.0f;
  __syncwarp();
  float xE = 0.0f, xN = 1.0f, xJ = 0.0f, xB = sp.nmove, xC = 0.0f, totscale = 0.0f;
  if (xmx != nullptr && lane == 0) { xmx[X_E] = xE; xmx[X_N] = xN; xmx[X_J] = xJ; xmx[X_B] = xB; xmx[X_C] = xC; xmx[X_SCALE] = 1.0f; }
  for (int i = 1; i <= L; ++i) {
    const int x = res[i - 1];
    const float *rp = fm.rfv + (int64_t)x * fm.Mpad;
    float esum = 0.0f, cM = 0.0f, cI = 0.0f, cD = 0.0f, dcarry = 0.0f;
    float *frow = FULL ? full + (int64_t)i * 3 * fm.Mpad : nullptr;
    for (int ch = 0; ch < nchunk; ++ch) {
      const int k = ch * 32 + lane + 1;
      const float oM = rowM[k], oI = rowI[k], oD = rowD[k];
      float pm = __shfl_up_sync(0xffffffffu, oM, 1), pi = __shfl_up_sync(0xffffffffu, oI, 1), pd = __shfl_up_sync(0xffffffffu, oD, 1);
      if (lane == 0) { pm = cM; pi = cI; pd = cD; }
      cM = __shfl_sync(0xffffffffu, oM, 31); cI = __shfl_sync(0xffffffffu, oI, 31); cD = __shfl_sync(0xffffffffu, oD, 31);
      const float4 t0 = __ldg(fm.tfv + 2 * k), t1 = __ldg(fm.tfv + 2 * k + 1);
      float sv = xB * t0.x;
      sv += pm * t0.y;
      sv += pi * t0.z;
      sv += pd * t0.w;
      sv *= __ldg(rp + k);
      const float nI = oM * t1.y + oI * t1.z;
      float dk;
      if (EXACT) {
        dk = dcarry;                                  // lane 0 holds D(k_first); the others receive theirs in turn
        for (int l = 0; l < 32; ++l) {
          const float out = __fadd_rn(__fmul_rn(sv, t1.x), __fmul_rn(dk, t1.w));   // D(k+1) from lane l's M(k), D(k)
          const float pass = __shfl_sync(0xffffffffu, out, l);
          if (lane == l + 1) dk = pass;
          if (l == 31) dcarry = pass;
          esum = __fadd_rn(esum, __shfl_sync(0xffffffffu, sv, l));                  // running sum of M's in k order (all lanes alike)
        }
        if (k > M) dk = 0.0f;
      } else {
        float B = sv * t1.x, T = t1.w;                 // D(k+1) = B + D(k) * T
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float Bl = __shfl_up_sync(0xffffffffu, B, o), Tl = __shfl_up_sync(0xffffffffu, T, o);
          if (lane >= o) { B = B + Bl * T; T = T * Tl; }
        }
        const float dnext = B + dcarry * T;
        dk = __shfl_up_sync(0xffffffffu, dnext, 1);
        if (lane == 0) dk = dcarry;
        dcarry = __shfl_sync(0xffffffffu, dnext, 31);
        if (k > M) dk = 0.0f;
        esum += sv + dk;
      }
      rowM[k] = sv; rowI[k] = nI; rowD[k] = dk;
      if (FULL) { frow[k] = sv; frow[fm.Mpad + k] = dk; frow[2 * fm.Mpad + k] = nI; }
    }
    if (EXACT) {
      __syncwarp();
      for (int k = 1; k <= M; ++k) esum = __fadd_rn(esum, rowD[k]);       // then the D's, again in k order
      xE = esum;
    } else xE = warp_sum_float(esum);
    xN = xN * sp.nloop;
    xC = (xC * sp.nloop) + (xE * sp.emove);
    xJ = (xJ * sp.nloop) + (xE * sp.eloop);
    xB = (xJ * sp.nmove) + (xN * sp.nmove);
    float scale = 1.0f;
    if (xE > 1.0e4f) {
      scale = xE;
      const float inv = __fdiv_rn(1.0f, xE);
      xN = __fdiv_rn(xN, xE); xC = __fdiv_rn(xC, xE); xJ = __fdiv_rn(xJ, xE); xB = __fdiv_rn(xB, xE);
      __syncwarp();
      for (int k = lane + 1; k <= nchunk * 32; k += 32) {
        const float a = rowM[k] * inv, b = rowD[k] * inv, c = rowI[k] * inv;
        rowM[k] = a; rowD[k] = b; rowI[k] = c;
        if (FULL) { frow[k] = a; frow[fm.Mpad + k] = b; frow[2 * fm.Mpad + k] = c; }
      }
      totscale += (float)log((double)xE);
      xE = 1.0f;
    }
    if (FULL && lane == 0) { frow[0] = 0.0f; frow[fm.Mpad] = 0.0f; frow[2 * fm.Mpad] = 0.0f; }
    if (xmx != nullptr && lane == 0) {
      float *xr = xmx + (int64_t)i * X_NX;
      xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = scale;
    }
    __syncwarp();
  }
  if (ret_sc != nullptr) *ret_sc = totscale + (float)log((double)(xC * sp.nmove));
}

// Backward, scaled by the Forward pass's per-row scale factors (fxmx).  rows hold row i+1 on entry to row i.
template <bool FULL>
__device__ __forceinline__ void backward_rows(const FwdModel &fm, const uint8_t *__restrict__ res, int L, const Specials sp,
                                              float *rowM, float *rowI, float *rowD, int lane,
                                              const float *fxmx, float *bxmx, float *full) {
  const int M = fm.M, nchunk = (M + 31) >> 5;
  for (int k = lane; k < nchunk * 32 + 2; k += 32) { rowM[k] = 0.0f; rowI[k] = 0.0f; rowD[k] = 0.0f; }
  __syncwarp();
  float xC = 0.0f, xE = 0.0f, xJ = 0.0f, xN = 0.0f, xB = 0.0f;
  for (int i = L; i >= 0; --i) {
    const float *rp = (i < L) ? fm.rfv + (int64_t)res[i] * fm.Mpad : nullptr;    // residue x_{i+1}
    if (i == L) {
      xC = sp.nmove; xE = xC * sp.emove; xB = 0.0f; xJ = 0.0f; xN = 0.0f;
    } else {
      float part = 0.0f;
      for (int k = lane + 1; k <= M; k += 32) part += rowM[k] * __ldg(rp + k) * __ldg(fm.tfv + 2 * k).x;
      xB = warp_sum_float(part);
      xC = xC * sp.nloop;
      xJ = (xB * sp.nmove) + (xJ * sp.nloop);
      xN = (xB * sp.nmove) + (xN * sp.nloop);
      xE = (xC * sp.emove) + (xJ * sp.eloop);
    }
    const float s = (i >= 1) ? fxmx[(int64_t)i * X_NX + X_SCALE] : 1.0f;
    float *frow = FULL ? full + (int64_t)i * 3 * fm.Mpad : nullptr;
    if (i >= 1) {
      const float inv = (s > 1.0f) ? __fdiv_rn(1.0f, s) : 1.0f;
      float cMn = 0.0f;        // e(k+1) M(i+1,k+1) for the first position of the chunk to the right
      float dcarry = 0.0f;     // D(i, first k of the chunk to the right)
      for (int ch = nchunk - 1; ch >= 0; --ch) {
        const int k = ch * 32 + lane + 1;
        const float nM = rowM[k], nI = rowI[k];
        const float em = (i < L && k <= M) ? nM * __ldg(rp + k) : 0.0f;     // e(k,x_{i+1}) M(i+1,k)
        float mnext = __shfl_down_sync(0xffffffffu, em, 1);
        if (lane == 31) mnext = cMn;
        cMn = __shfl_sync(0xffffffffu, em, 0);
        const float4 tn = __ldg(fm.tfv + 2 * (k + 1));   // transitions entering node k+1: MM(k), IM(k), DM(k)
        const float4 t1 = __ldg(fm.tfv + 2 * k + 1);     // leaving node k: MD, MI, II, DD
        const float inext = (i < L) ? nI : 0.0f;
        const bool in = (k <= M);
        const float tmm = (k < M) ? tn.y : 0.0f, tim = (k < M) ? tn.z : 0.0f, tdm = (k < M) ? tn.w : 0.0f;
        // D(k) = (xE + mnext*tDM) + D(k+1)*tDD : reverse scan over affine maps
        float B = in ? (xE + mnext * tdm) : 0.0f, T = in ? t1.w : 0.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float Br = __shfl_down_sync(0xffffffffu, B, o), Tr = __shfl_down_sync(0xffffffffu, T, o);
          if (lane + o < 32) { B = B + Br * T; T = T * Tr; }
        }
        const float dv = B + dcarry * T;                  // D(i,k)
        float dnext = __shfl_down_sync(0xffffffffu, dv, 1);
        if (lane == 31) dnext = dcarry;
        dcarry = __shfl_sync(0xffffffffu, dv, 0);
        float mv = xE + mnext * tmm + inext * t1.y + dnext * t1.x;
        float iv = mnext * tim + inext * t1.z;
        float dvv = dv;
        if (!in) { mv = 0.0f; iv = 0.0f; dvv = 0.0f; }
        mv *= inv; iv *= inv; dvv *= inv;
        __syncwarp();
        rowM[k] = mv; rowI[k] = iv; rowD[k] = dvv;
        if (FULL) { frow[k] = mv; frow[fm.Mpad + k] = dvv; frow[2 * fm.Mpad + k] = iv; }
      }
      if (s > 1.0f) { xE = __fdiv_rn(xE, s); xN = __fdiv_rn(xN, s); xJ = __fdiv_rn(xJ, s); xB = __fdiv_rn(xB, s); xC = __fdiv_rn(xC, s); }
    } else {
      xC = 0.0f; xJ = 0.0f; xE = 0.0f;
      if (FULL) for (int k = lane; k < 3 * fm.Mpad; k += 32) frow[k] = 0.0f;
    }
    if (FULL && lane == 0 && i >= 1) { frow[0] = 0.0f; frow[fm.Mpad] = 0.0f; frow[2 * fm.Mpad] = 0.0f; }
    if (bxmx != nullptr && lane == 0) {
      float *xr = bxmx + (int64_t)i * X_NX;
      xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = s;
    }
    __syncwarp();
  }
}

}  // namespace ckm
