// domdef_common.cuh -- pieces shared by the chunked (kernels_domdef.cu) and lane-blocked (kernels_blk.cu) domain stages
#pragma once
#include "device_utils.cuh"
#include "stages.hpp"
#include "fwdback.cuh"

namespace ckm {

enum { ST_M = 1, ST_D, ST_I, ST_S, ST_N, ST_B, ST_E, ST_C, ST_T, ST_J };

// domain decoding of the special-state columns + the region walk by posterior heuristics
__device__ __forceinline__ void regions_tail(const DomdefParams &p, int pi, int L, const Specials sp, const float *xf, const float *xb,
                                             float *btot, float *etot, float *mocc, float *n2sc, int lane) {
    const float scaleproduct = __fdiv_rn(1.0f, xb[X_N]);
    for (int i = lane; i <= L; i += 32) {
      n2sc[i] = 0.0f;
      if (i == 0) { btot[0] = 0.0f; etot[0] = 0.0f; mocc[0] = 0.0f; continue; }
      const float *f0 = xf + (int64_t)(i - 1) * X_NX, *f1 = xf + (int64_t)i * X_NX;
      const float *b0 = xb + (int64_t)(i - 1) * X_NX, *b1 = xb + (int64_t)i * X_NX;
      btot[i] = (f0[X_B] * b0[X_B]) * f0[X_SCALE] * scaleproduct;      // per-row terms; prefix-summed below
      etot[i] = (f1[X_E] * b1[X_E]) * f1[X_SCALE] * scaleproduct;
      float njcp;
      njcp = f0[X_N] * b1[X_N] * sp.nloop * scaleproduct;
      njcp += f0[X_J] * b1[X_J] * sp.nloop * scaleproduct;
      njcp += f0[X_C] * b1[X_C] * sp.nloop * scaleproduct;
      mocc[i] = 1.0f - njcp;
    }
    __syncwarp();
    if (lane == 0) {
      float bt = 0.0f, et = 0.0f;
      for (int i = 1; i <= L; ++i) { bt = bt + btot[i]; et = et + etot[i]; btot[i] = bt; etot[i] = et; }
      int i = -1; bool triggered = false;
      for (int j = 1; j <= L; ++j) {
        if (!triggered) {
          if (mocc[j] - (btot[j] - btot[j - 1]) < 0.10f) i = j;
          else if (i == -1) i = j;
          if (mocc[j] >= 0.25f) triggered = true;
        } else if (mocc[j] - (etot[j] - etot[j - 1]) < 0.10f) {
          float mx = -1.0f;
          for (int z = i; z <= j; ++z) {
            const float a = etot[z] - etot[i - 1], b = btot[j] - btot[z - 1];
            mx = fmaxf(mx, fminf(a, b));
          }
          const int pos = atomicAdd(p.region_count, 1);
          if (pos < p.region_cap) { Region r; r.pair = pi; r.i = i; r.j = j; r.multi = (mx >= 0.20f) ? 1 : 0; p.regions[pos] = r; }
          i = -1; triggered = false;
        }
      }
    }
}

}  // namespace ckm
