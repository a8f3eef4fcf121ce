// kernels_vitp.cu -- ViterbiFilter on packed int16x2 lanes (two model positions per 32-bit register, one DPX
// VIADDMNMX.S16x2 per add+max of both).  Stage 3 of the cascade behind checkm/hmmer.py:70-71; same result as the
// int32 kernels of kernels_filters.cu (the int16-saturating recurrence of SURVEY.md A.5 step 3), at ~4.5 ALU instructions per DP
// cell instead of ~11.
//
// Layout.  W = vq/2 words per lane, K = 64 W cells.  Word w of lane l holds position k0 = l*W + w + 1 in its low
// half and k1 = 32*W + k0 in its high half, so the (k-1) neighbour of BOTH halves is word w-1 of the same lane (no
// intra-register shift); word 0 takes it from lane l-1 by one SHFL per state, and lane 0 patches its two halves with
// one PRMT (low: the k = 0 boundary; high: the low half of lane 31, i.e. position 32*W).
//
// Arithmetic.  The reference filter saturates int16 adds at -32768; VIADDMNMX wraps.  We keep every DP value
// >= FLOOR = -10240 (third operand of the instruction) and clamp every table entry at -22528 when the tables are built
// (models.cu), so a + b >= -32768 always.  Raising low values changes nothing that can reach the final score when
//   C1:  |E->J| + |tmove(L)| + max_k |tBM(k)| + 64 <= 22528      C2:  |tmove(L)| + max_k |tBM(k)| + 64 <= 22240
// because then (a) a path that crossed a clamped ("-inf") entry sits >= 22528 below a value already banked in xJ, and
// re-entering the same cell through E->J->B->M costs less than that; (b) a path restarted from FLOOR is beaten by the
// plain B->M entry of the same cell (>= 12000 - |tmove| - |tBM| + e).  Rows whose every match cell sits below FLOOR
// report xE = FLOOR; if the final xC is not above FLOOR + E->C such a row may have set it and the pair is re-scored.
// Upper end: a row maximum >= 32767 - max emission could wrap on the next add -> re-scored (these are the strong
// hits, a few percent).  The full D->D evaluation floors block sums of tDD at -16384, which is only safe under
//   C1': |E->J| + |tmove(L)| + max_k |tBM(k)| + 64 <= 16384;
// pairs outside C1' are re-scored if the lazy-F test ever asks for the full evaluation.  "Re-scored" = appended to
// p.redo, which search.cu runs through the int32 kernels.  Pairs outside C1/C2 go there directly.
#include "engine.hpp"
#include "device_utils.cuh"
#include "stages.hpp"

namespace ckm {

constexpr uint32_t VP_FLOORW = 0xD800D800u;      // -10240 | -10240
constexpr int      VP_FLOOR  = -10240;
constexpr uint32_t VP_TFLOORW = 0xC000C000u;     // -16384 | -16384 : floor of tDD block sums

__device__ __forceinline__ int vp_lo(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
__device__ __forceinline__ int vp_hi(uint32_t w) { return (int)w >> 16; }

// Work distribution.  Every class kernel reads the one survivor list of the bias filter.  A warp takes 32 consecutive
// candidates at a time (from the class's cursor p.vit_work[cls]: dynamic, so the long ORFs at the end of the list do not
// leave a tail; without a cursor, in static strides), each lane looks up the class of one of them -- one coalesced pass
// over the list per kernel instead of a dependent two-load round trip per candidate and warp -- and the warp then scores
// the ones that are its own, one after the other.
template <int W, bool TSMEM>
__global__ void __launch_bounds__(128) vitp_kernel(FilterParams p, int cls) {
  extern __shared__ __align__(16) uint8_t vsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  uint4 *tws = reinterpret_cast<uint4 *>(vsm) + (size_t)warp * W * 64;
  const uint32_t sel = (lane == 0) ? 0x1054u : 0x3210u;
  const int n = min(*p.in_count, p.in_cap);
  int32_t *cursor = (p.vit_work != nullptr) ? p.vit_work + cls : nullptr;
  for (int64_t it = 0;; ++it) {
    int64_t base64;
    if (cursor != nullptr) {
      int b = 0;
      if (lane == 0) b = atomicAdd(cursor, 32);
      base64 = __shfl_sync(0xffffffffu, b, 0);
    } else {
      base64 = ((it * gridDim.x + blockIdx.x) * wpb + warp) * 32;
    }
    if (base64 >= n) break;
    const int base = (int)base64;
    bool mine = false;
    if (base + lane < n) {
      const int vq = p.ms[p.in[base + lane].model].vq;
      mine = (vq == 2 * W) || (W == 1 && vq == 0);
    }
    unsigned todo = __ballot_sync(0xffffffffu, mine);
  while (todo != 0u) {
    const int c = base + __ffs(todo) - 1;
    todo &= todo - 1u;
    Candidate cd = p.in[c];
    const int m = cd.model;
    const ModelScalars ms = p.ms[m];
    const bool unclassed = (ms.vq == 0 && W == 1);      // the first class also forwards models without a class
    if (ms.vq != 2 * W && !unclassed) continue;
    const int s = cd.seq, L = p.len[s];
    bool pass = true, redo = false;
    if (cd.P > p.F2) {
      const int tmove = p.tmove_w[s];
      const int cost = -tmove - (int)ms.vit_tbm + 64;
      const bool c1 = (cost - (int)ms.xw_e_loop <= 22528) && (cost <= 22240);
      const bool c1p = (cost - (int)ms.xw_e_loop <= 16384);
      redo = unclassed || !c1;
      float vsc = 0.0f;
      if (!redo) {
        uint4 tr0[TSMEM ? 1 : W], tr1[TSMEM ? 1 : W];
        const uint4 *tsrc = p.twp + ms.blk_off * 32;
        if (TSMEM) {
          __syncwarp();
          for (int z = lane; z < W * 64; z += 32) tws[(z & 1) * (W * 32) + (z >> 1)] = __ldg(tsrc + z);     // [half][w][lane]: conflict-free LDS.128
          __syncwarp();
        } else {
#pragma unroll
          for (int w = 0; w < W; ++w) { tr0[TSMEM ? 0 : w] = __ldg(tsrc + (w * 32 + lane) * 2); tr1[TSMEM ? 0 : w] = __ldg(tsrc + (w * 32 + lane) * 2 + 1); }
        }
#define TR0(w) (TSMEM ? tws[(w) * 32 + lane] : tr0[TSMEM ? 0 : (w)])
#define TR1(w) (TSMEM ? tws[W * 32 + (w) * 32 + lane] : tr1[TSMEM ? 0 : (w)])
        const uint32_t *rwp = p.rwp + ms.blk_off * 32 * (KPAD / 2) + lane;
        uint32_t Mx[W], Ix[W], Dx[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { Mx[w] = VP_FLOORW; Ix[w] = VP_FLOORW; Dx[w] = VP_FLOORW; }
        const int ddbound = ms.ddbound_w, cap = 32767 - (int)ms.vit_emax;
        const int e_move = ms.xw_e_move, e_loop = ms.xw_e_loop;
        int xN = ms.base_w, xB = xN + tmove, xJ = -32768, xC = -32768;
        const uint4 *rp = reinterpret_cast<const uint4 *>(p.res + p.off[s]);
        const int nblk = (L + 15) >> 4;
        uint4 r16 = (nblk > 0) ? __ldg(rp) : make_uint4(0, 0, 0, 0);
        uint32_t ecur[W];
        {
          const uint32_t x0 = r16.x & 0xffu;
#pragma unroll
          for (int w = 0; w < W; ++w) ecur[w] = __ldg(rwp + (x0 * W + w) * 32);
        }
        bool stop = false;
        for (int b = 0; b < nblk && !stop; ++b) {
          const uint4 rnext = (b + 1 < nblk) ? __ldg(rp + b + 1) : make_uint4(0, 0, 0, 0);
          for (int j = 0; j < 4 && !stop; ++j) {
            const uint32_t wcur = (j == 0) ? r16.x : (j == 1) ? r16.y : (j == 2) ? r16.z : r16.w;
            const uint32_t wnxt = (j == 0) ? r16.y : (j == 1) ? r16.z : (j == 2) ? r16.w : rnext.x;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              if (b * 16 + j * 4 + rr >= L) { stop = true; break; }
              // emission words of the NEXT row are requested now and consumed one iteration later
              uint32_t enext[W];
              {
                const uint32_t xn = (rr < 3) ? ((wcur >> (8 * (rr + 1))) & 0xffu) : (wnxt & 0xffu);
#pragma unroll
                for (int w = 0; w < W; ++w) enext[w] = __ldg(rwp + (xn * W + w) * 32);
              }
              // row i-1 values of the position just below my block, both halves
              const uint32_t shM = __shfl_sync(0xffffffffu, Mx[W - 1], (lane + 31) & 31);
              const uint32_t shI = __shfl_sync(0xffffffffu, Ix[W - 1], (lane + 31) & 31);
              const uint32_t shD = __shfl_sync(0xffffffffu, Dx[W - 1], (lane + 31) & 31);
              const uint32_t pm0 = __byte_perm(shM, VP_FLOORW, sel), pi0 = __byte_perm(shI, VP_FLOORW, sel), pd0 = __byte_perm(shD, VP_FLOORW, sel);
              const uint32_t xBw = __byte_perm((uint32_t)xB, 0u, 0x1010u);
              uint32_t md[W];
              uint32_t xEw = VP_FLOORW, dmw = VP_FLOORW;
#pragma unroll
              for (int w = W - 1; w >= 0; --w) {
                const uint32_t pm = (w > 0) ? Mx[w - 1] : pm0, pi = (w > 0) ? Ix[w - 1] : pi0, pd = (w > 0) ? Dx[w - 1] : pd0;
                const uint4 t0 = TR0(w), t1 = TR1(w);
                uint32_t sv = __viaddmax_s16x2(xBw, t0.x, VP_FLOORW);
                sv = __viaddmax_s16x2(pm, t0.y, sv);
                sv = __viaddmax_s16x2(pi, t0.z, sv);
                sv = __viaddmax_s16x2(pd, t0.w, sv);
                sv = __viaddmax_s16x2(sv, ecur[w], VP_FLOORW);
                const uint32_t nI = __viaddmax_s16x2(Ix[w], t1.z, __viaddmax_s16x2(Mx[w], t1.y, VP_FLOORW));
                md[w] = __viaddmax_s16x2(sv, t1.x, VP_FLOORW);
                Mx[w] = sv; Ix[w] = nI;
              }
#pragma unroll
              for (int w = 0; w + 1 < W; w += 2) { xEw = __vimax3_s16x2(xEw, Mx[w], Mx[w + 1]); dmw = __vimax3_s16x2(dmw, md[w], md[w + 1]); }
              if (W & 1) { xEw = __vimax3_s16x2(xEw, Mx[W - 1], Mx[W - 1]); dmw = __vimax3_s16x2(dmw, md[W - 1], md[W - 1]); }
              const int xE = __reduce_max_sync(0xffffffffu, max(vp_lo(xEw), vp_hi(xEw)));
              if (xE >= cap) { redo = true; stop = true; break; }
              xC = max(xC, xE + e_move);
              xJ = max(xJ, xE + e_loop);
              xB = max(xJ + tmove, xN + tmove);
              const int Dmax = __reduce_max_sync(0xffffffffu, max(vp_lo(dmw), vp_hi(dmw)));
              if (Dmax + ddbound > xB) {
                if (!c1p) { redo = true; stop = true; break; }
                // full D->D.  Per lane and half: composite f(d) = max(Bb, d + Tb) of my W cells; inclusive max-plus scan over
                // the lanes (the two halves are two independent chains here); then the high chain takes the low chain's exit.
                uint32_t Bb = VP_FLOORW, Tb = 0u;
#pragma unroll
                for (int w = 0; w < W; ++w) { const uint32_t tdd = TR1(w).w; Bb = __viaddmax_s16x2(Bb, tdd, md[w]); Tb = __viaddmax_s16x2(Tb, __vimax3_s16x2(tdd, VP_TFLOORW, VP_TFLOORW), VP_TFLOORW); }
                uint32_t Bs = Bb, Ts = Tb;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const uint32_t Bl = __shfl_up_sync(0xffffffffu, Bs, o), Tl = __shfl_up_sync(0xffffffffu, Ts, o);
                  if (lane >= o) { Bs = __viaddmax_s16x2(Bl, Ts, Bs); Ts = __viaddmax_s16x2(Ts, Tl, VP_TFLOORW); }
                }
                uint32_t din = __shfl_up_sync(0xffffffffu, Bs, 1), Tex = __shfl_up_sync(0xffffffffu, Ts, 1);
                if (lane == 0) { din = VP_FLOORW; Tex = 0u; }
                const uint32_t lowexit = __shfl_sync(0xffffffffu, Bs, 31);                 // low half: D(i, 32W+1)
                const uint32_t dmid = __byte_perm(lowexit, VP_FLOORW, 0x1054u);            // (low: FLOOR, high: that exit)
                din = __viaddmax_s16x2(dmid, Tex, din);
                uint32_t d = din;
#pragma unroll
                for (int w = 0; w < W; ++w) { Dx[w] = d; d = __viaddmax_s16x2(d, TR1(w).w, md[w]); }
              } else {
                // lazy F: no D->D path can beat entering from B; D(i,k) = M(i,k-1) + tMD(k-1)
                const uint32_t shd = __shfl_sync(0xffffffffu, md[W - 1], (lane + 31) & 31);
#pragma unroll
                for (int w = W - 1; w >= 1; --w) Dx[w] = md[w - 1];
                Dx[0] = __byte_perm(shd, VP_FLOORW, sel);
              }
#pragma unroll
              for (int w = 0; w < W; ++w) ecur[w] = enext[w];
            }
          }
          r16 = rnext;
        }
        if (!redo && xC <= VP_FLOOR + e_move) redo = true;      // a floored row may have set xC (or nothing scored at all)
        if (!redo) {
          vsc = __fsub_rn(__fadd_rn((float)xC, (float)tmove), (float)ms.base_w);
          vsc = __fdiv_rn(vsc, ms.scale_w);
          vsc = __fsub_rn(vsc, 3.0f);
        }
#undef TR0
#undef TR1
      }
      if (redo) {
        if (lane == 0) {
          const int pos = atomicAdd(p.redo_count, 1);
          if (pos < p.redo_cap) p.redo[pos] = cd;
        }
        continue;
      }
      cd.vitsc = vsc;
      const float seq_score = __fdiv_rn(__fsub_rn(vsc, cd.filtersc), 0.69314718055994529f);
      const double P = gumbel_surv((double)seq_score, (double)ms.evparam[2], (double)ms.evparam[3]);
      cd.P = P;
      pass = (P <= p.F2);
      if (lane == 0 && p.dense_vit != nullptr) p.dense_vit[(int64_t)p.model_slot[m] * p.nseq + s] = vsc;
    }
    if (lane == 0 && pass) {
      const int pos = atomicAdd(p.out_count, 1);
      if (pos < p.out_cap) p.out[pos] = cd;
      if (p.dense_passed != nullptr) atomicOr_u8(p.dense_passed, (int64_t)p.model_slot[m] * p.nseq + s, 4);
    }
  }   // candidates of this group of 32
  }   // groups
}

template <int W, bool TSMEM>
static int launch_vitp_w(const FilterParams &p, int cls, int grid, cudaStream_t st) {
  const int sm = TSMEM ? 4 * W * 64 * (int)sizeof(uint4) : 0;
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(vitp_kernel<W, TSMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(vitp)");
  }
  vitp_kernel<W, TSMEM><<<grid, 128, sm, st>>>(p, cls);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? CKM_OK : cuda_fail(e, "vitp_kernel launch");
}

// every (model slot, sequence) pair as a candidate that still needs the Viterbi filter (parity entry point ckm_viterbi_scores)
__global__ void all_pairs_kernel(Candidate *out, int32_t *count, const int32_t *slot_model, int32_t nslots, int32_t nseq) {
  const int64_t n = (int64_t)nslots * nseq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    Candidate cd;
    cd.seq = (int32_t)(i % nseq); cd.model = slot_model[i / nseq]; cd.usc = 0.0f; cd.filtersc = 0.0f; cd.vitsc = 0.0f; cd.fwdsc = 0.0f; cd.P = 1.0;
    out[i] = cd;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = (int32_t)n;
}
int launch_all_pairs(Candidate *out, int32_t *count, const int32_t *slot_model, int32_t nslots, int32_t nseq, cudaStream_t st) {
  all_pairs_kernel<<<592, 256, 0, st>>>(out, count, slot_model, nslots, nseq);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? CKM_OK : cuda_fail(e, "all_pairs_kernel launch");
}

int launch_vitp(const FilterParams &p, int cls, int grid, cudaStream_t st) {
  switch (cls) {
    case 0: return launch_vitp_w<1, false>(p, cls, grid, st);
    case 1: return launch_vitp_w<2, false>(p, cls, grid, st);
    case 2: return launch_vitp_w<3, false>(p, cls, grid, st);
    case 3: return launch_vitp_w<4, false>(p, cls, grid, st);
    case 4: return launch_vitp_w<6, true>(p, cls, grid, st);
    case 5: return launch_vitp_w<8, true>(p, cls, grid, st);
    case 6: return launch_vitp_w<10, true>(p, cls, grid, st);
    case 7: return launch_vitp_w<12, true>(p, cls, grid, st);
    case 8: return launch_vitp_w<14, true>(p, cls, grid, st);
    case 9: return launch_vitp_w<16, true>(p, cls, grid, st);
  }
  set_error("launch_vitp: bad class"); return CKM_EINVAL;
}

}  // namespace ckm
