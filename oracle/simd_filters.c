/*
 * simd_filters.c -- SSE2 striped MSV and Viterbi filters for the CPU BASELINE of bench.py (--impl reference, cpu_baseline).
 *
 * TEST / BENCH INFRASTRUCTURE ONLY, like the rest of oracle/: nothing under checkm_b200/ includes, links or calls this.
 *
 * Why it exists: the reference (CheckM) spends its time inside the external `hmmsearch`, whose filters are hand-vectorised
 * (HMMER 3.1 impl_sse: 16 uint8 lanes for MSV, 8 int16 lanes for the Viterbi filter, striped over the model as in
 * Farrar 2007).  A scalar restatement is 20-40x slower per core than that binary, so timing it next to the GPU says
 * nothing.  These two functions restate the striped filters with the same lane counts, so that the CPU arm of the bench
 * runs at the speed class of the real tool.  They return exactly the scores of the scalar orc_msv / orc_vitfilter
 * (tests/test_oracle_cpu.py::test_simd_filters_equal_scalar), which are themselves pinned to HMMER's own calibration.
 *
 * Layout: Q vectors cover the model; position k (1-based) sits in vector (k-1) % Q, lane (k-1) / Q.
 */
#include <emmintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "hmmer_oracle.h"

typedef struct {
  int      M, Qb, Qw;
  __m128i *rbv;      /* [KP][Qb]  MSV costs                                         */
  __m128i *rwv;      /* [KP][Qw]  Viterbi emission words                            */
  __m128i *twv;      /* [7*Qw + Qw] per q: BM MM IM DM MD MI II, then the DD block   */
} orc_striped;

enum { T_BM = 0, T_MM, T_IM, T_DM, T_MD, T_MI, T_II, T_DD };

static void *amalloc(size_t n) { void *p = NULL; if (posix_memalign(&p, 16, n ? n : 16)) return NULL; return p; }

void *orc_striped_create(const orc_profile *p)
{
  orc_striped *s = (orc_striped *)calloc(1, sizeof(orc_striped));
  const int M = p->M;
  s->M = M;
  s->Qb = (M - 1) / 16 + 1; if (s->Qb < 2) s->Qb = 2;
  s->Qw = (M - 1) / 8 + 1;  if (s->Qw < 2) s->Qw = 2;
  s->rbv = (__m128i *)amalloc(sizeof(__m128i) * ORC_KP * s->Qb);
  s->rwv = (__m128i *)amalloc(sizeof(__m128i) * ORC_KP * s->Qw);
  s->twv = (__m128i *)amalloc(sizeof(__m128i) * 8 * s->Qw);
  for (int x = 0; x < ORC_KP; x++) {
    for (int q = 0; q < s->Qb; q++) {
      uint8_t v[16];
      for (int z = 0; z < 16; z++) { int k = z * s->Qb + q + 1; v[z] = (k <= M) ? p->rbv[(size_t)x * (M + 1) + k] : 255; }
      memcpy(&s->rbv[x * s->Qb + q], v, 16);
    }
    for (int q = 0; q < s->Qw; q++) {
      int16_t v[8];
      for (int z = 0; z < 8; z++) { int k = z * s->Qw + q + 1; v[z] = (k <= M) ? p->rwv[(size_t)x * (M + 1) + k] : -32768; }
      memcpy(&s->rwv[x * s->Qw + q], v, 16);
    }
  }
  for (int q = 0; q < s->Qw; q++) {
    for (int t = 0; t < 7; t++) {
      int16_t v[8];
      for (int z = 0; z < 8; z++) { int k = z * s->Qw + q + 1; v[z] = (k <= M) ? p->twv[k * 8 + t] : -32768; }
      memcpy(&s->twv[q * 7 + t], v, 16);
    }
    int16_t v[8];
    for (int z = 0; z < 8; z++) { int k = z * s->Qw + q + 1; v[z] = (k <= M) ? p->twv[k * 8 + T_DD] : -32768; }
    memcpy(&s->twv[7 * s->Qw + q], v, 16);
  }
  return s;
}

void orc_striped_free(void *h)
{
  orc_striped *s = (orc_striped *)h;
  if (!s) return;
  free(s->rbv); free(s->rwv); free(s->twv); free(s);
}

static inline uint8_t hmax_epu8(__m128i a)
{
  a = _mm_max_epu8(a, _mm_srli_si128(a, 8));
  a = _mm_max_epu8(a, _mm_srli_si128(a, 4));
  a = _mm_max_epu8(a, _mm_srli_si128(a, 2));
  a = _mm_max_epu8(a, _mm_srli_si128(a, 1));
  return (uint8_t)_mm_extract_epi16(a, 0);
}
static inline int16_t hmax_epi16(__m128i a)
{
  a = _mm_max_epi16(a, _mm_srli_si128(a, 8));
  a = _mm_max_epi16(a, _mm_srli_si128(a, 4));
  a = _mm_max_epi16(a, _mm_srli_si128(a, 2));
  return (int16_t)_mm_extract_epi16(a, 0);
}
static inline uint8_t byteify_u(float scale, float sc) { sc = -1.0f * roundf(scale * sc); return (sc > 255.0f) ? 255 : (uint8_t)sc; }
static inline int16_t wordify_(float scale, float sc)
{
  sc = roundf(scale * sc);
  if (sc >= 32767.0f) return 32767;
  if (sc <= -32768.0f) return -32768;
  return (int16_t)sc;
}

/* 8-bit MSV filter, 16 lanes.  Same return convention as orc_msv.  `dp` is caller scratch of Qb vectors (or NULL). */
int orc_msv_simd(const orc_profile *p, const void *h, const uint8_t *dsq, int L, float *ret_sc, int *ret_xJ)
{
  const orc_striped *s = (const orc_striped *)h;
  const int Q = s->Qb;
  __m128i dpbuf[64], *dp = (Q <= 64) ? dpbuf : (__m128i *)amalloc(sizeof(__m128i) * Q);
  for (int q = 0; q < Q; q++) dp[q] = _mm_setzero_si128();
  const int tjb = byteify_u(p->scale_b, logf(3.0f / (float)(L + 3)));
  int tjbm = tjb + p->tbm_b; if (tjbm > 255) tjbm = 255;
  const __m128i biasv = _mm_set1_epi8((char)p->bias_b);
  const int overflow_at = 255 - p->bias_b;
  int xJ = 0;
  int xB = (int)p->base_b - tjbm; if (xB < 0) xB = 0;
  for (int i = 0; i < L; i++) {
    const __m128i *rsc = s->rbv + (size_t)dsq[i] * Q;
    const __m128i xBv = _mm_set1_epi8((char)xB);
    __m128i xEv = _mm_setzero_si128();
    __m128i mpv = _mm_slli_si128(dp[Q - 1], 1);
    for (int q = 0; q < Q; q++) {
      __m128i sv = _mm_max_epu8(mpv, xBv);
      sv = _mm_adds_epu8(sv, biasv);
      sv = _mm_subs_epu8(sv, rsc[q]);
      xEv = _mm_max_epu8(xEv, sv);
      mpv = dp[q];
      dp[q] = sv;
    }
    int xE = hmax_epu8(xEv);
    if (xE >= overflow_at) { if (dp != dpbuf) free(dp); *ret_sc = INFINITY; if (ret_xJ) *ret_xJ = 256; return 1; }
    xE -= p->tec_b; if (xE < 0) xE = 0;
    if (xE > xJ) xJ = xE;
    xB = (p->base_b > xJ) ? p->base_b : xJ;
    xB -= tjbm; if (xB < 0) xB = 0;
  }
  if (dp != dpbuf) free(dp);
  float sc = ((float)(xJ - tjb) - (float)p->base_b);
  sc /= p->scale_b;
  sc -= 3.0f;
  *ret_sc = sc;
  if (ret_xJ) *ret_xJ = xJ;
  return 0;
}

/* 16-bit Viterbi filter, 8 lanes, with the delayed evaluation of the D->D chain (it is completed only on rows where a
 * delete path could beat re-entering through B: max D + ddbound > xB). */
int orc_vitfilter_simd(const orc_profile *p, const void *h, const uint8_t *dsq, int L, float *ret_sc)
{
  const orc_striped *s = (const orc_striped *)h;
  const int Q = s->Qw;
  __m128i buf[3 * 140], *mem = (Q <= 140) ? buf : (__m128i *)amalloc(sizeof(__m128i) * 3 * Q);
  __m128i *MMX = mem, *DMX = mem + Q, *IMX = mem + 2 * Q;
  const __m128i negv = _mm_set1_epi16(-32768);
  const __m128i neg0 = _mm_insert_epi16(_mm_setzero_si128(), -32768, 0);     /* -inf into lane 0 after a shift */
  for (int q = 0; q < Q; q++) MMX[q] = DMX[q] = IMX[q] = negv;
  const int16_t tmove = wordify_(p->scale_w, logf(3.0f / (float)(L + 3)));
  int xN = p->base_w, xB = xN + tmove, xJ = -32768, xC = -32768;
  for (int i = 0; i < L; i++) {
    const __m128i *rsc = s->rwv + (size_t)dsq[i] * Q;
    const __m128i *tsc = s->twv;
    __m128i dcv = negv, xEv = negv, Dmaxv = negv;
    const __m128i xBv = _mm_set1_epi16((short)xB);
    __m128i mpv = _mm_or_si128(_mm_slli_si128(MMX[Q - 1], 2), neg0);
    __m128i dpv = _mm_or_si128(_mm_slli_si128(DMX[Q - 1], 2), neg0);
    __m128i ipv = _mm_or_si128(_mm_slli_si128(IMX[Q - 1], 2), neg0);
    for (int q = 0; q < Q; q++) {
      __m128i sv = _mm_adds_epi16(xBv, tsc[T_BM]);
      sv = _mm_max_epi16(sv, _mm_adds_epi16(mpv, tsc[T_MM]));
      sv = _mm_max_epi16(sv, _mm_adds_epi16(ipv, tsc[T_IM]));
      sv = _mm_max_epi16(sv, _mm_adds_epi16(dpv, tsc[T_DM]));
      sv = _mm_adds_epi16(sv, rsc[q]);
      xEv = _mm_max_epi16(xEv, sv);
      mpv = MMX[q]; dpv = DMX[q]; ipv = IMX[q];
      MMX[q] = sv;
      DMX[q] = dcv;
      dcv = _mm_adds_epi16(sv, tsc[T_MD]);
      Dmaxv = _mm_max_epi16(dcv, Dmaxv);
      sv = _mm_adds_epi16(mpv, tsc[T_MI]);
      sv = _mm_max_epi16(sv, _mm_adds_epi16(ipv, tsc[T_II]));
      IMX[q] = sv;
      tsc += 7;
    }
    int xE = hmax_epi16(xEv);
    if (xE >= 32767) { if (mem != buf) free(mem); *ret_sc = INFINITY; return 1; }
    { int b = xE + p->xw_e_move; if (b > xC) xC = b; }
    { int b = xE + p->xw_e_loop; if (b > xJ) xJ = b; }
    { int a = xJ + tmove, b = xN + tmove; xB = a > b ? a : b; }
    if (xC < -32768) xC = -32768;
    if (xJ < -32768) xJ = -32768;
    if (xB < -32768) xB = -32768;
    /* the D->D chain */
    const int Dmax = hmax_epi16(Dmaxv);
    if (Dmax + (int)p->ddbound_w > xB) {
      const __m128i *tdd = s->twv + 7 * Q;
      dcv = _mm_or_si128(_mm_slli_si128(dcv, 2), neg0);
      for (int q = 0; q < Q; q++) { DMX[q] = _mm_max_epi16(dcv, DMX[q]); dcv = _mm_adds_epi16(DMX[q], tdd[q]); }
      int q;
      do {
        dcv = _mm_or_si128(_mm_slli_si128(dcv, 2), neg0);
        for (q = 0; q < Q; q++) {
          if (_mm_movemask_epi8(_mm_cmpgt_epi16(dcv, DMX[q])) == 0) break;
          DMX[q] = _mm_max_epi16(dcv, DMX[q]);
          dcv = _mm_adds_epi16(DMX[q], tdd[q]);
        }
      } while (q == Q);
    } else {
      DMX[0] = _mm_or_si128(_mm_slli_si128(dcv, 2), neg0);
    }
  }
  if (mem != buf) free(mem);
  if (xC > -32768) {
    float sc = (float)xC + (float)tmove - (float)p->base_w;
    sc /= p->scale_w;
    sc -= 3.0f;
    *ret_sc = sc;
  } else *ret_sc = -INFINITY;
  return 0;
}
