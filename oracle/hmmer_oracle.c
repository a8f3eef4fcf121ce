/*
 * hmmer_oracle.c -- CPU restatement of the HMMER3 hmmsearch acceleration pipeline.
 * TEST INFRASTRUCTURE ONLY (see hmmer_oracle.h).  PARITY UNPINNED for the search arithmetic.
 *
 * Reference call sites this stands in for (relative to /root/reference):
 *   checkm/hmmer.py:61-74            HMMERRunner.search -> os.system('hmmsearch --domtblout ...')
 *   checkm/markerGeneFinder.py:137-142  flags: --cpu N --notextw -E 0.1 --domE 0.1 [--noali]
 *   checkm/hmmer.py:184-200,255-285  the domtblout columns CheckM reads back
 * The algorithm itself lives in the third-party HMMER 3.1b2 (unpinned apt dependency, docker/Dockerfile:5);
 * it is restated here from its published design: MSV -> bias filter ->
 * ViterbiFilter -> ForwardParser -> BackwardParser -> domain definition by posterior heuristics ->
 * per-envelope Forward/Backward/decoding/null2/optimal-accuracy alignment -> scores, E-values, thresholds.
 *
 * Plain scalar C, one cell at a time, no SIMD: it is the checker, written for clarity.
 */
#define _GNU_SOURCE
#include "hmmer_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <pthread.h>

#define LOG2C      0.69314718055994529
#define F1_DEFAULT 0.02
#define F2_DEFAULT 1e-3
#define F3_DEFAULT 1e-5
#define OMEGA      (1.0f / 256.0f)

static const char AMINO[] = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";

/* Swiss-Prot 50.8 background frequencies (HMMER p7_AminoFrequencies) */
static const float BGF[ORC_K] = {
  0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f, 0.0590092f,
  0.0594422f, 0.0963728f, 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f, 0.0540978f, 0.0683364f,
  0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f };

/* degeneracy table: B=ND, J=IL, Z=QE, O=K, U=C, X=any (Easel amino alphabet) */
static int degen_has(int x, int r)
{
  switch (x) {
  case 21: return (r == 11 || r == 2);
  case 22: return (r == 7  || r == 9);
  case 23: return (r == 13 || r == 3);
  case 24: return (r == 8);
  case 25: return (r == 1);
  case 26: return 1;
  default: return (x == r);
  }
}

int orc_digitize(const char *seq, int n, uint8_t *dsq)
{
  static int8_t map[256]; static int init = 0;
  int bad = 0;
  if (!init) {
    memset(map, -1, sizeof(map));
    for (int i = 0; i < ORC_KP; i++) {
      map[(unsigned char)AMINO[i]] = (int8_t)i;
      if (AMINO[i] >= 'A' && AMINO[i] <= 'Z') map[(unsigned char)(AMINO[i] + 32)] = (int8_t)i;
    }
    map['.'] = 20;
    init = 1;
  }
  for (int i = 0; i < n; i++) {
    int c = map[(unsigned char)seq[i]];
    if (c < 0) { c = 26; bad++; }
    dsq[i] = (uint8_t)c;
  }
  return bad;
}

/* =====================================================================================
 * HMMER3/f ASCII reader.  Numbers are -ln p, '*' is p = 0.
 * ===================================================================================== */
static float parse_prob(const char *tok)
{
  if (tok[0] == '*') return 0.0f;
  return expf(-1.0f * (float)atof(tok));
}

static int read_model(FILE *fp, orc_hmm *h, char **linebuf, size_t *cap)
{
  ssize_t n;
  int in_header = 0;
  memset(h, 0, sizeof(*h));
  while ((n = getline(linebuf, cap, fp)) > 0) {
    char *line = *linebuf;
    if (!in_header) {
      if (strncmp(line, "HMMER3", 6) == 0) { in_header = 1; }
      continue;
    }
    char tag[32]; int off = 0;
    if (sscanf(line, "%31s%n", tag, &off) != 1) continue;
    char *rest = line + off;
    while (*rest == ' ' || *rest == '\t') rest++;
    size_t rl = strlen(rest);
    while (rl > 0 && (rest[rl - 1] == '\n' || rest[rl - 1] == '\r' || rest[rl - 1] == ' ')) rest[--rl] = 0;
    if      (!strcmp(tag, "NAME")) { strncpy(h->name, rest, sizeof(h->name) - 1); }
    else if (!strcmp(tag, "ACC"))  { strncpy(h->acc,  rest, sizeof(h->acc) - 1); }
    else if (!strcmp(tag, "DESC")) { strncpy(h->desc, rest, sizeof(h->desc) - 1); }
    else if (!strcmp(tag, "LENG")) { h->M = atoi(rest); }
    else if (!strcmp(tag, "GA"))   { if (sscanf(rest, "%f %f", &h->ga[0], &h->ga[1]) == 2) h->has_ga = 1; }
    else if (!strcmp(tag, "TC"))   { if (sscanf(rest, "%f %f", &h->tc[0], &h->tc[1]) == 2) h->has_tc = 1; }
    else if (!strcmp(tag, "NC"))   { if (sscanf(rest, "%f %f", &h->nc[0], &h->nc[1]) == 2) h->has_nc = 1; }
    else if (!strcmp(tag, "STATS")) {
      char loc[16], kind[16]; float a, b;
      if (sscanf(rest, "%15s %15s %f %f", loc, kind, &a, &b) == 4) {
        if      (!strcmp(kind, "MSV"))     { h->evparam[ORC_MMU]  = a; h->evparam[ORC_MLAMBDA] = b; h->has_stats |= 1; }
        else if (!strcmp(kind, "VITERBI")) { h->evparam[ORC_VMU]  = a; h->evparam[ORC_VLAMBDA] = b; h->has_stats |= 2; }
        else if (!strcmp(kind, "FORWARD")) { h->evparam[ORC_FTAU] = a; h->evparam[ORC_FLAMBDA] = b; h->has_stats |= 4; }
      }
    }
    else if (!strcmp(tag, "HMM")) {
      /* body */
      int M = h->M;
      if (M <= 0) return -2;
      h->mat = (float *)calloc((size_t)(M + 1) * ORC_K, sizeof(float));
      h->ins = (float *)calloc((size_t)(M + 1) * ORC_K, sizeof(float));
      h->t   = (float *)calloc((size_t)(M + 1) * ORC_NT, sizeof(float));
      if (getline(linebuf, cap, fp) <= 0) return -3;           /* transition label line */
      for (int k = 0; k <= M; k++) {
        char *save, *tok;
        /* match line (or COMPO for k=0) */
        if (getline(linebuf, cap, fp) <= 0) return -3;
        tok = strtok_r(*linebuf, " \t\n", &save);
        if (k == 0) {
          if (tok && !strcmp(tok, "COMPO")) {
            for (int x = 0; x < ORC_K; x++) { tok = strtok_r(NULL, " \t\n", &save); if (!tok) return -4; h->compo[x] = parse_prob(tok); }
            h->has_compo = 1;
            if (getline(linebuf, cap, fp) <= 0) return -3;
            tok = strtok_r(*linebuf, " \t\n", &save);
          }
          /* tok is now the first insert emission of node 0 */
          for (int x = 0; x < ORC_K; x++) { if (!tok) return -4; h->ins[x] = parse_prob(tok); tok = strtok_r(NULL, " \t\n", &save); }
        } else {
          if (!tok || atoi(tok) != k) return -5;
          for (int x = 0; x < ORC_K; x++) { tok = strtok_r(NULL, " \t\n", &save); if (!tok) return -4; h->mat[k * ORC_K + x] = parse_prob(tok); }
          if (getline(linebuf, cap, fp) <= 0) return -3;
          tok = strtok_r(*linebuf, " \t\n", &save);
          for (int x = 0; x < ORC_K; x++) { if (!tok) return -4; h->ins[k * ORC_K + x] = parse_prob(tok); tok = strtok_r(NULL, " \t\n", &save); }
        }
        if (getline(linebuf, cap, fp) <= 0) return -3;
        tok = strtok_r(*linebuf, " \t\n", &save);
        for (int z = 0; z < ORC_NT; z++) { if (!tok) return -4; h->t[k * ORC_NT + z] = parse_prob(tok); tok = strtok_r(NULL, " \t\n", &save); }
      }
      if (getline(linebuf, cap, fp) <= 0) return -3;           /* "//" */
      return 1;
    }
  }
  return 0; /* EOF */
}

int orc_hmmfile_read(const char *path, orc_hmm **ret_hmms, int *ret_n)
{
  FILE *fp = fopen(path, "r");
  if (!fp) return -1;
  int n = 0, cap_n = 64, st;
  orc_hmm *hs = (orc_hmm *)calloc(cap_n, sizeof(orc_hmm));
  char *line = NULL; size_t cap = 0;
  while (1) {
    if (n == cap_n) { cap_n *= 2; hs = (orc_hmm *)realloc(hs, cap_n * sizeof(orc_hmm)); }
    st = read_model(fp, &hs[n], &line, &cap);
    if (st <= 0) break;
    n++;
  }
  free(line); fclose(fp);
  if (st < 0) { orc_hmms_free(hs, n + 1); return st; }
  *ret_hmms = hs; *ret_n = n;
  return 0;
}

void orc_hmms_free(orc_hmm *hmms, int n)
{
  if (!hmms) return;
  for (int i = 0; i < n; i++) { free(hmms[i].mat); free(hmms[i].ins); free(hmms[i].t); }
  free(hmms);
}
orc_hmm *orc_hmm_at(orc_hmm *hmms, int i) { return &hmms[i]; }

/* =====================================================================================
 * Profile configuration: multihit local; MSV bytes, ViterbiFilter words, Forward odds.
 * ===================================================================================== */
static uint8_t unbiased_byteify(float scale_b, float sc)
{
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.0f) ? 255 : (uint8_t)sc;
}
static uint8_t biased_byteify(float scale_b, uint8_t bias_b, float sc)
{
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.0f - (float)bias_b) ? 255 : (uint8_t)((uint8_t)sc + bias_b);
}
static int16_t wordify(float scale_w, float sc)
{
  sc = roundf(scale_w * sc);
  if (sc >= 32767.0f) return 32767;
  if (sc <= -32768.0f) return -32768;
  return (int16_t)sc;
}

enum { T_BM = 0, T_MM, T_IM, T_DM, T_MD, T_MI, T_II, T_DD };

orc_profile *orc_profile_create(const orc_hmm *hmm)
{
  int M = hmm->M;
  orc_profile *p = (orc_profile *)calloc(1, sizeof(orc_profile));
  p->M = M; p->hmm = hmm;
  p->tsc = (float *)malloc(sizeof(float) * (M + 1) * ORC_NT);
  p->bm  = (float *)malloc(sizeof(float) * (M + 1));
  p->msc = (float *)malloc(sizeof(float) * ORC_KP * (M + 1));
  p->rbv = (uint8_t *)malloc((size_t)ORC_KP * (M + 1));
  p->rwv = (int16_t *)malloc(sizeof(int16_t) * ORC_KP * (M + 1));
  p->twv = (int16_t *)malloc(sizeof(int16_t) * (M + 1) * 8);
  p->rfv = (float *)malloc(sizeof(float) * ORC_KP * (M + 1));
  p->tfv = (float *)malloc(sizeof(float) * (M + 1) * 8);

  /* local entry: occ[k] / sum_j occ[j]*(M-j+1) */
  float *occ = (float *)malloc(sizeof(float) * (M + 1));
  occ[0] = 0.0f;
  occ[1] = hmm->t[0 * ORC_NT + ORC_MI] + hmm->t[0 * ORC_NT + ORC_MM];
  for (int k = 2; k <= M; k++)
    occ[k] = occ[k - 1] * (hmm->t[(k - 1) * ORC_NT + ORC_MM] + hmm->t[(k - 1) * ORC_NT + ORC_MI]) +
             (1.0f - occ[k - 1]) * hmm->t[(k - 1) * ORC_NT + ORC_DM];
  float Z = 0.0f;
  for (int k = 1; k <= M; k++) Z += occ[k] * (float)(M - k + 1);
  p->bm[0] = -INFINITY;
  for (int k = 1; k <= M; k++) p->bm[k] = (float)log(occ[k] / Z);
  free(occ);

  for (int z = 0; z < ORC_NT; z++) { p->tsc[z] = -INFINITY; p->tsc[M * ORC_NT + z] = -INFINITY; }
  for (int k = 1; k < M; k++)
    for (int z = 0; z < ORC_NT; z++)
      p->tsc[k * ORC_NT + z] = (float)log(hmm->t[k * ORC_NT + z]);

  /* match emission log-odds; degenerate residues get the expected score; gap/'*'/'~' are -inf */
  for (int x = 0; x < ORC_KP; x++) p->msc[x * (M + 1)] = -INFINITY;
  for (int k = 1; k <= M; k++) {
    float sc[ORC_KP];
    for (int x = 0; x < ORC_K; x++) sc[x] = (float)log((double)hmm->mat[k * ORC_K + x] / BGF[x]);
    sc[20] = -INFINITY; sc[27] = -INFINITY; sc[28] = -INFINITY;
    for (int x = ORC_K + 1; x <= ORC_KP - 3; x++) {
      float result = 0.0f, denom = 0.0f;
      for (int i = 0; i < ORC_K; i++)
        if (degen_has(x, i)) { result += sc[i] * BGF[i]; denom += BGF[i]; }
      sc[x] = result / denom;
    }
    for (int x = 0; x < ORC_KP; x++) p->msc[x * (M + 1) + k] = sc[x];
  }

  /* ---- MSV: 1/3-bit units, base 190 ---- */
  {
    float max = 0.0f;                /* insert scores are hard-wired 0, so the maximum is >= 0 */
    if (M < 2) max = -INFINITY;
    for (int x = 0; x < ORC_K; x++)
      for (int k = 1; k <= M; k++)
        if (p->msc[x * (M + 1) + k] > max) max = p->msc[x * (M + 1) + k];
    p->scale_b = (float)(3.0 / LOG2C);
    p->base_b  = 190;
    p->bias_b  = unbiased_byteify(p->scale_b, -1.0f * max);
    for (int x = 0; x < ORC_KP; x++) {
      p->rbv[x * (M + 1)] = 255;
      for (int k = 1; k <= M; k++) p->rbv[x * (M + 1) + k] = biased_byteify(p->scale_b, p->bias_b, p->msc[x * (M + 1) + k]);
    }
    p->tbm_b = unbiased_byteify(p->scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
    p->tec_b = unbiased_byteify(p->scale_b, logf(0.5f));
  }

  /* ---- ViterbiFilter: 1/500-bit units, base 12000 ---- */
  {
    p->scale_w = (float)(500.0 / LOG2C);
    p->base_w  = 12000;
    for (int x = 0; x < ORC_KP; x++) {
      p->rwv[x * (M + 1)] = -32768;
      for (int k = 1; k <= M; k++) p->rwv[x * (M + 1) + k] = wordify(p->scale_w, p->msc[x * (M + 1) + k]);
    }
    for (int k = 0; k <= M; k++) {
      int16_t *tw = p->twv + k * 8;
      /* entering node k: B->M_k, and M/I/D_{k-1} -> M_k */
      tw[T_BM] = (k >= 1) ? wordify(p->scale_w, p->bm[k]) : -32768;
      tw[T_MM] = (k >= 1) ? wordify(p->scale_w, p->tsc[(k - 1) * ORC_NT + ORC_MM]) : -32768;
      tw[T_IM] = (k >= 1) ? wordify(p->scale_w, p->tsc[(k - 1) * ORC_NT + ORC_IM]) : -32768;
      tw[T_DM] = (k >= 1) ? wordify(p->scale_w, p->tsc[(k - 1) * ORC_NT + ORC_DM]) : -32768;
      /* leaving node k */
      tw[T_MD] = (k >= 1 && k < M) ? wordify(p->scale_w, p->tsc[k * ORC_NT + ORC_MD]) : -32768;
      tw[T_MI] = (k >= 1 && k < M) ? wordify(p->scale_w, p->tsc[k * ORC_NT + ORC_MI]) : -32768;
      tw[T_II] = (k >= 1 && k < M) ? wordify(p->scale_w, p->tsc[k * ORC_NT + ORC_II]) : -32768;
      tw[T_DD] = (k >= 1 && k < M) ? wordify(p->scale_w, p->tsc[k * ORC_NT + ORC_DD]) : -32768;
      if (tw[T_BM] > 0) tw[T_BM] = 0;
      if (tw[T_MM] > 0) tw[T_MM] = 0;
      if (tw[T_IM] > 0) tw[T_IM] = 0;
      if (tw[T_DM] > 0) tw[T_DM] = 0;
      if (tw[T_MD] > 0) tw[T_MD] = 0;
      if (tw[T_MI] > 0) tw[T_MI] = 0;
      if (tw[T_II] > -1) tw[T_II] = -1;     /* an II cost of 0 is never allowed */
    }
    p->xw_e_loop = wordify(p->scale_w, (float)-LOG2C);
    p->xw_e_move = wordify(p->scale_w, (float)-LOG2C);
    p->ddbound_w = -32768;
    for (int k = 2; k < M - 1; k++) {
      int dd = (int)wordify(p->scale_w, p->tsc[k * ORC_NT + ORC_DD]);
      dd += (int)wordify(p->scale_w, p->tsc[(k + 1) * ORC_NT + ORC_DM]);
      dd -= (int)wordify(p->scale_w, p->bm[k + 2]);
      if (dd > p->ddbound_w) p->ddbound_w = (int16_t)dd;
    }
  }

  /* ---- Forward/Backward odds ratios ---- */
  for (int x = 0; x < ORC_KP; x++)
    for (int k = 0; k <= M; k++) p->rfv[x * (M + 1) + k] = expf(p->msc[x * (M + 1) + k]);
  for (int k = 0; k <= M; k++) {
    float *tf = p->tfv + k * 8;
    tf[T_BM] = (k >= 1) ? expf(p->bm[k]) : 0.0f;
    tf[T_MM] = (k >= 1) ? expf(p->tsc[(k - 1) * ORC_NT + ORC_MM]) : 0.0f;
    tf[T_IM] = (k >= 1) ? expf(p->tsc[(k - 1) * ORC_NT + ORC_IM]) : 0.0f;
    tf[T_DM] = (k >= 1) ? expf(p->tsc[(k - 1) * ORC_NT + ORC_DM]) : 0.0f;
    tf[T_MD] = expf(p->tsc[k * ORC_NT + ORC_MD]);
    tf[T_MI] = expf(p->tsc[k * ORC_NT + ORC_MI]);
    tf[T_II] = expf(p->tsc[k * ORC_NT + ORC_II]);
    tf[T_DD] = expf(p->tsc[k * ORC_NT + ORC_DD]);
  }
  return p;
}

void orc_profile_enable_simd(orc_profile *p) { if (p && !p->striped) p->striped = orc_striped_create(p); }

void orc_profile_free(orc_profile *p)
{
  if (!p) return;
  free(p->tsc); free(p->bm); free(p->msc); free(p->rbv); free(p->rwv); free(p->twv); free(p->rfv); free(p->tfv);
  orc_striped_free(p->striped);
  free(p);
}

/* =====================================================================================
 * Statistics helpers
 * ===================================================================================== */
static double gumbel_surv(double x, double mu, double lambda)
{
  double y  = lambda * (x - mu);
  double ey = -exp(-y);
  if (fabs(ey) < 5e-9) return -ey;
  return 1.0 - exp(ey);
}
static double exp_surv(double x, double mu, double lambda)    { return (x < mu) ? 1.0 : exp(-lambda * (x - mu)); }
static double exp_logsurv(double x, double mu, double lambda) { return (x < mu) ? 0.0 : -lambda * (x - mu); }

#define LOGSUM_TBL 16000
static float flogsum_lookup[LOGSUM_TBL];
static pthread_once_t logsum_once = PTHREAD_ONCE_INIT;
static void flogsum_init(void)
{
  for (int i = 0; i < LOGSUM_TBL; i++) flogsum_lookup[i] = (float)log(1.0 + exp((double)-i / 1000.0));
}
static float flogsum(float a, float b)
{
  pthread_once(&logsum_once, flogsum_init);
  float max = (a > b) ? a : b, min = (a > b) ? b : a;
  return (min == -INFINITY || (max - min) >= 15.7f) ? max : max + flogsum_lookup[(int)((max - min) * 1000.0f)];
}

float orc_null1(int L)
{
  float p1 = (float)L / (float)(L + 1);
  return (float)L * logf(p1) + logf(1.0f - p1);
}

/* =====================================================================================
 * MSV filter: uint8 costs, saturating arithmetic.
 * ===================================================================================== */
static inline int addus8(int a, int b) { int s = a + b; return s > 255 ? 255 : s; }
static inline int subus8(int a, int b) { int s = a - b; return s < 0 ? 0 : s; }

int orc_msv(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc, int *ret_xJ)
{
  int M = p->M;
  uint8_t *dp = (uint8_t *)calloc(M + 1, 1);
  int tjb  = unbiased_byteify(p->scale_b, logf(3.0f / (float)(L + 3)));
  int tjbm = addus8(tjb, p->tbm_b);       /* (int8_t)+(int8_t) in the 8-bit original; sums stay < 128 for real models */
  int xJ = 0;
  int xB = subus8(p->base_b, tjbm);
  for (int i = 0; i < L; i++) {
    const uint8_t *rsc = p->rbv + (size_t)dsq[i] * (M + 1);
    int xE = 0, diag = 0;                 /* diag = dp[k-1] of the previous row; column 0 is -inf (0) */
    for (int k = 1; k <= M; k++) {
      int sv = diag > xB ? diag : xB;
      sv = addus8(sv, p->bias_b);
      sv = subus8(sv, rsc[k]);
      diag = dp[k];
      dp[k] = (uint8_t)sv;
      if (sv > xE) xE = sv;
    }
    if (addus8(xE, p->bias_b) == 255) { free(dp); *ret_sc = INFINITY; if (ret_xJ) *ret_xJ = 256; return 1; }
    xE = subus8(xE, p->tec_b);
    if (xE > xJ) xJ = xE;
    xB = (p->base_b > xJ) ? p->base_b : xJ;
    xB = subus8(xB, tjbm);
  }
  free(dp);
  float sc = ((float)(xJ - tjb) - (float)p->base_b);
  sc /= p->scale_b;
  sc -= 3.0f;
  *ret_sc = sc;
  if (ret_xJ) *ret_xJ = xJ;
  return 0;
}

/* max over all cells of the MSV recurrence with the J state switched off (xB fixed at base - tjb - tbm).
 * Equals the MSV's running max xE whenever max xE - tec <= base; used to check the GPU's SSV pass. */
int orc_ssv_xe(const orc_profile *p, const uint8_t *dsq, int L)
{
  int M = p->M;
  uint8_t *dp = (uint8_t *)calloc(M + 1, 1);
  int tjb  = unbiased_byteify(p->scale_b, logf(3.0f / (float)(L + 3)));
  int tjbm = addus8(tjb, p->tbm_b);
  int xB = subus8(p->base_b, tjbm), xEmax = 0;
  for (int i = 0; i < L; i++) {
    const uint8_t *rsc = p->rbv + (size_t)dsq[i] * (M + 1);
    int diag = 0;
    for (int k = 1; k <= M; k++) {
      int sv = diag > xB ? diag : xB;
      sv = addus8(sv, p->bias_b);
      sv = subus8(sv, rsc[k]);
      diag = dp[k];
      dp[k] = (uint8_t)sv;
      if (sv > xEmax) xEmax = sv;
    }
  }
  free(dp);
  return xEmax;
}

/* =====================================================================================
 * Bias filter (A.5 step 2): two-state HMM Forward; state 0 = background, state 1 = model composition.
 * ===================================================================================== */
float orc_biasfilter(const orc_profile *p, const uint8_t *dsq, int L)
{
  const orc_hmm *h = p->hmm;
  float p1 = (float)L / (float)(L + 1);
  float t[2][3];
  float L1 = (float)p->M / 8.0f;
  t[0][0] = p1;                 t[0][1] = 1.0f - p1;          t[0][2] = 1.0f;
  t[1][0] = 1.0f / (L1 + 1.0f); t[1][1] = L1 / (L1 + 1.0f);   t[1][2] = 1.0f;
  float pi[2] = { 0.999f, 0.001f };
  float eo[ORC_KP][2];
  for (int x = 0; x < ORC_K; x++) { eo[x][0] = BGF[x] / BGF[x]; eo[x][1] = h->compo[x] / BGF[x]; }
  eo[20][0] = eo[20][1] = 1.0f; eo[27][0] = eo[27][1] = 1.0f; eo[28][0] = eo[28][1] = 1.0f;
  for (int x = ORC_K + 1; x <= ORC_KP - 3; x++)
    for (int k = 0; k < 2; k++) {
      float num = 0.0f, denom = 0.0f;
      for (int y = 0; y < ORC_K; y++)
        if (degen_has(x, y)) { num += (k == 0 ? BGF[y] : h->compo[y]); denom += BGF[y]; }
      eo[x][k] = (denom > 0.0f) ? num / denom : 0.0f;
    }
  if (L == 0) return 0.0f + (float)L * logf(p1) + logf(1.0f - p1);
  float dp0, dp1, max, logsc = 0.0f;
  dp0 = eo[dsq[0]][0] * pi[0];
  dp1 = eo[dsq[0]][1] * pi[1];
  max = dp0 > dp1 ? dp0 : dp1;
  if (max < 0.0f) max = 0.0f;
  dp0 /= max; dp1 /= max;
  logsc += (float)log(max);
  for (int i = 1; i < L; i++) {
    float n0 = 0.0f, n1 = 0.0f;
    n0 += dp0 * t[0][0]; n0 += dp1 * t[1][0]; n0 *= eo[dsq[i]][0];
    n1 += dp0 * t[0][1]; n1 += dp1 * t[1][1]; n1 *= eo[dsq[i]][1];
    max = 0.0f;
    if (n0 > max) max = n0;
    if (n1 > max) max = n1;
    dp0 = n0 / max; dp1 = n1 / max;
    logsc += (float)log(max);
  }
  float last = 0.0f;
  last += dp0 * t[0][2];
  last += dp1 * t[1][2];
  logsc += (float)log(last);
  return logsc + (float)L * logf(p1) + logf(1.0f - p1);
}

/* =====================================================================================
 * ViterbiFilter (A.5 step 3): int16, saturating, N/C/J loops free with a -3 nat correction.
 * D->D paths are fully evaluated (the original's "lazy F" shortcut is score-preserving).
 * ===================================================================================== */
static inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

int orc_vitfilter(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc)
{
  int M = p->M;
  int16_t *mmx = (int16_t *)malloc(sizeof(int16_t) * (M + 1) * 3);
  int16_t *imx = mmx + (M + 1), *dmx = imx + (M + 1);
  for (int k = 0; k <= M; k++) mmx[k] = imx[k] = dmx[k] = -32768;
  int16_t tmove = wordify(p->scale_w, logf(3.0f / (float)(L + 3)));
  int xN = p->base_w, xB = xN + tmove, xJ = -32768, xC = -32768, xE;
  for (int i = 0; i < L; i++) {
    const int16_t *rsc = p->rwv + (size_t)dsq[i] * (M + 1);
    int mpv = -32768, ipv = -32768, dpv = -32768;     /* row i-1, column k-1 */
    int dcv = -32768;                                  /* D(i,k) being assembled */
    xE = -32768;
    for (int k = 1; k <= M; k++) {
      const int16_t *tw = p->twv + k * 8;
      int sv = sat16(xB + tw[T_BM]);
      int c;
      c = sat16(mpv + tw[T_MM]); if (c > sv) sv = c;
      c = sat16(ipv + tw[T_IM]); if (c > sv) sv = c;
      c = sat16(dpv + tw[T_DM]); if (c > sv) sv = c;
      sv = sat16(sv + rsc[k]);
      if (sv > xE) xE = sv;
      mpv = mmx[k]; ipv = imx[k]; dpv = dmx[k];
      mmx[k] = (int16_t)sv;
      dmx[k] = (int16_t)dcv;
      /* D(i,k+1) = max(M(i,k)+tMD(k), D(i,k)+tDD(k)) */
      int md = sat16(sv + tw[T_MD]);
      int dd = sat16(dcv + tw[T_DD]);
      dcv = md > dd ? md : dd;
      int mi = sat16(mpv + tw[T_MI]);
      int ii = sat16(ipv + tw[T_II]);
      imx[k] = (int16_t)(mi > ii ? mi : ii);
    }
    if (xE >= 32767) { free(mmx); *ret_sc = INFINITY; return 1; }
    /* N/C/J loop scores are 0 */
    { int a = xC, b = xE + p->xw_e_move; xC = a > b ? a : b; }
    { int a = xJ, b = xE + p->xw_e_loop; xJ = a > b ? a : b; }
    { int a = xJ + tmove, b = xN + tmove; xB = a > b ? a : b; }
    if (xC < -32768) xC = -32768;
    if (xJ < -32768) xJ = -32768;
    if (xB < -32768) xB = -32768;
  }
  free(mmx);
  if (xC > -32768) {
    float sc = (float)xC + (float)tmove - (float)p->base_w;
    sc /= p->scale_w;
    sc -= 3.0f;
    *ret_sc = sc;
  } else *ret_sc = -INFINITY;
  return 0;
}

/* =====================================================================================
 * Forward / Backward in scaled odds-ratio space (A.5 steps 4-5).
 * A "matrix" keeps specials for every row and, optionally, the full M/I/D rows.
 * ===================================================================================== */
enum { X_E = 0, X_N, X_J, X_B, X_C, X_SCALE, X_NX };

typedef struct {
  int    M, L;
  int    full;
  float *dp;        /* full ? (L+1)*(M+1)*3 : 2*(M+1)*3 ; [row][k][M,D,I]            */
  float *xmx;       /* (L+1)*X_NX                                                     */
  float  totscale;
} fmx;

enum { C_M = 0, C_D = 1, C_I = 2 };

static fmx *fmx_create(int M, int L, int full)
{
  fmx *x = (fmx *)calloc(1, sizeof(fmx));
  x->M = M; x->L = L; x->full = full;
  size_t rows = full ? (size_t)(L + 1) : 2;
  x->dp  = (float *)calloc(rows * (M + 1) * 3, sizeof(float));
  x->xmx = (float *)calloc((size_t)(L + 1) * X_NX, sizeof(float));
  return x;
}
static void fmx_free(fmx *x) { if (x) { free(x->dp); free(x->xmx); free(x); } }
static inline float *fmx_row(const fmx *x, int i) { return x->dp + (size_t)(x->full ? i : (i & 1)) * (x->M + 1) * 3; }

typedef struct { float nloop, nmove, eloop, emove; } specials;
static specials make_specials(int Lcfg, int multihit)
{
  specials s;
  float nj = multihit ? 1.0f : 0.0f;
  s.nmove = (2.0f + nj) / ((float)Lcfg + 2.0f + nj);
  s.nloop = 1.0f - s.nmove;
  s.eloop = multihit ? 0.5f : 0.0f;
  s.emove = multihit ? 0.5f : 1.0f;
  return s;
}

static int forward_engine_seq(const orc_profile *p, const uint8_t *dsq, int L, specials sp, fmx *ox, float *ret_sc)
{
  int M = p->M;
  float *dpc = fmx_row(ox, 0), *dpp;
  for (int k = 0; k <= M; k++) dpc[k * 3 + C_M] = dpc[k * 3 + C_D] = dpc[k * 3 + C_I] = 0.0f;
  float xE = 0.0f, xN = 1.0f, xJ = 0.0f, xB = sp.nmove, xC = 0.0f;
  ox->xmx[X_E] = xE; ox->xmx[X_N] = xN; ox->xmx[X_J] = xJ; ox->xmx[X_B] = xB; ox->xmx[X_C] = xC; ox->xmx[X_SCALE] = 1.0f;
  ox->totscale = 0.0f;
  for (int i = 1; i <= L; i++) {
    dpp = dpc; dpc = fmx_row(ox, i);
    const float *rp = p->rfv + (size_t)dsq[i - 1] * (M + 1);
    dpc[C_M] = dpc[C_D] = dpc[C_I] = 0.0f;
    float dcv = 0.0f;
    xE = 0.0f;
    for (int k = 1; k <= M; k++) {
      const float *tf = p->tfv + k * 8;
      float sv = xB * tf[T_BM];
      sv += dpp[(k - 1) * 3 + C_M] * tf[T_MM];
      sv += dpp[(k - 1) * 3 + C_I] * tf[T_IM];
      sv += dpp[(k - 1) * 3 + C_D] * tf[T_DM];
      sv *= rp[k];
      xE += sv;
      dpc[k * 3 + C_M] = sv;
      dpc[k * 3 + C_D] = dcv;
      /* D(i,k+1) = M(i,k)*tMD(k) + D(i,k)*tDD(k) */
      dcv = sv * tf[T_MD] + dcv * tf[T_DD];
      dpc[k * 3 + C_I] = dpp[k * 3 + C_M] * tf[T_MI] + dpp[k * 3 + C_I] * tf[T_II];
    }
    for (int k = 1; k <= M; k++) xE += dpc[k * 3 + C_D];
    xN = xN * sp.nloop;
    xC = (xC * sp.nloop) + (xE * sp.emove);
    xJ = (xJ * sp.nloop) + (xE * sp.eloop);
    xB = (xJ * sp.nmove) + (xN * sp.nmove);
    if (xE > 1.0e4f) {
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      float inv = 1.0f / xE;
      for (int k = 1; k <= M; k++) { dpc[k * 3 + C_M] *= inv; dpc[k * 3 + C_D] *= inv; dpc[k * 3 + C_I] *= inv; }
      ox->xmx[i * X_NX + X_SCALE] = xE;
      ox->totscale += (float)log((double)xE);
      xE = 1.0f;
    } else ox->xmx[i * X_NX + X_SCALE] = 1.0f;
    float *xr = ox->xmx + i * X_NX;
    xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC;
  }
  if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { if (ret_sc) *ret_sc = -INFINITY; return 1; }
  if (ret_sc) *ret_sc = ox->totscale + (float)log((double)(xC * sp.nmove));
  return 0;
}

/* Backward, scaled with the Forward matrix's per-row scale factors. */
static int backward_engine_seq(const orc_profile *p, const uint8_t *dsq, int L, specials sp, const fmx *fwd, fmx *bck, float *ret_sc)
{
  int M = p->M;
  float *dpc = NULL, *dpp = NULL;
  float xC = 0.0f, xE = 0.0f, xJ = 0.0f, xN = 0.0f, xB = 0.0f;
  bck->totscale = 0.0f;
  for (int i = L; i >= 0; i--) {
    dpp = dpc; dpc = fmx_row(bck, i);             /* dpp = row i+1 (NULL at i = L) */
    const float *rp = (i < L) ? p->rfv + (size_t)dsq[i] * (M + 1) : NULL;   /* residue x_{i+1} */
    if (i == L) {
      xC = sp.nmove;                              /* C <- T */
      xE = xC * sp.emove;                         /* E <- C */
      xB = xJ = xN = 0.0f;
    } else {
      xB = 0.0f;                                  /* B(i) = sum_k tBM(k) e(k,x_{i+1}) M(i+1,k) */
      for (int k = 1; k <= M; k++) xB += dpp[k * 3 + C_M] * rp[k] * p->tfv[k * 8 + T_BM];
      xC = xC * sp.nloop;
      xJ = (xB * sp.nmove) + (xJ * sp.nloop);
      xN = (xB * sp.nmove) + (xN * sp.nloop);
      xE = (xC * sp.emove) + (xJ * sp.eloop);
    }
    dpc[C_M] = dpc[C_D] = dpc[C_I] = 0.0f;
    if (i >= 1) {
      float dnext = 0.0f;                         /* D(i,k+1) */
      for (int k = M; k >= 1; k--) {
        const float *tf = p->tfv + k * 8;         /* transitions leaving node k */
        float mnext = (k < M && i < L) ? rp[k + 1] * dpp[(k + 1) * 3 + C_M] : 0.0f;
        float tmm = (k < M) ? p->tfv[(k + 1) * 8 + T_MM] : 0.0f;
        float tim = (k < M) ? p->tfv[(k + 1) * 8 + T_IM] : 0.0f;
        float tdm = (k < M) ? p->tfv[(k + 1) * 8 + T_DM] : 0.0f;
        float inext = (i < L) ? dpp[k * 3 + C_I] : 0.0f;    /* I(i+1,k), insert odds 1 */
        float mv = xE + mnext * tmm + inext * tf[T_MI] + dnext * tf[T_MD];
        float iv = mnext * tim + inext * tf[T_II];
        float dv = xE + mnext * tdm + dnext * tf[T_DD];
        dpc[k * 3 + C_M] = mv; dpc[k * 3 + C_I] = iv; dpc[k * 3 + C_D] = dv;
        dnext = dv;
      }
    } else {
      for (int k = 1; k <= M; k++) dpc[k * 3 + C_M] = dpc[k * 3 + C_I] = dpc[k * 3 + C_D] = 0.0f;
      xC = 0.0f; xJ = 0.0f; xE = 0.0f;            /* at i = 0 only N and B are reachable */
    }
    float s = (i >= 1) ? fwd->xmx[i * X_NX + X_SCALE] : 1.0f;
    if (s > 1.0f) {
      xE /= s; xN /= s; xJ /= s; xB /= s; xC /= s;
      float inv = 1.0f / s;
      for (int k = 1; k <= M; k++) { dpc[k * 3 + C_M] *= inv; dpc[k * 3 + C_D] *= inv; dpc[k * 3 + C_I] *= inv; }
    }
    bck->totscale += (float)log(s);
    float *xr = bck->xmx + i * X_NX;
    xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = s;
  }
  if (ret_sc) *ret_sc = bck->totscale + (float)log(xN);
  return 0;
}


/* =====================================================================================
 * Evaluation order of the fp32 sums.
 *
 * fp32 addition is not associative, so the last bits of a Forward score depend on the order in which the cells of a row
 * are added up -- HMMER's own results differ between its SSE, VMX, NEON and generic builds for exactly this reason.
 * Two orders are restated here:
 *   ORC_ORDER_SEQUENTIAL  the textbook order: k = 1..M left to right (HMMER's generic implementation).
 *   ORC_ORDER_CANONICAL   the order the engine under test is specified to use (DESIGN.md "canonical order"): the model is cut
 *                         into 32 blocks of Q consecutive positions (Q = orc_block_width(M)); a row sum is the sum of the 32
 *                         block sums taken in a butterfly (pairs 16 apart, then 8, 4, 2, 1); the delete chain
 *                         D(k+1) = M(k) tMD(k) + D(k) tDD(k) is carried across blocks by composing the 32 affine block maps in
 *                         a doubling scan and then replayed inside each block.  Models longer than 1024 positions have no
 *                         blocked form; for them this order IS the sequential one.
 * Same recurrences, same per-cell arithmetic; only the association of the sums differs (tests bound the difference
 * between the two orders below 1e-3 bits, typically 1e-5).  With the canonical order the GPU's floats are reproduced bit for
 * bit, which is what lets the parity tests compare printed domtblout text.
 * ===================================================================================== */
static int g_order = ORC_ORDER_CANONICAL;
void orc_set_order(int order) { g_order = order; }
int  orc_get_order(void) { return g_order; }
int  orc_block_width(int M)
{
  return (M <= 64) ? 2 : (M <= 128) ? 4 : (M <= 192) ? 6 : (M <= 256) ? 8 : (M <= 384) ? 12 : (M <= 512) ? 16 : (M <= 640) ? 20 :
         (M <= 768) ? 24 : (M <= 896) ? 28 : (M <= 1024) ? 32 : 0;
}
#define NB 32            /* blocks */
#define QMAX 32

/* butterfly sum of 32 block values: every block ends up with the same total */
static float butterfly_sum(const float *v)
{
  float a[NB], b[NB];
  memcpy(a, v, sizeof(a));
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < NB; l++) b[l] = a[l] + a[l ^ o];
    memcpy(a, b, sizeof(a));
  }
  return a[0];
}

static int forward_engine_blk(const orc_profile *p, const uint8_t *dsq, int L, specials sp, fmx *ox, float *ret_sc, int Q)
{
  const int M = p->M;
  static const float zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float (*Mx)[QMAX] = (float (*)[QMAX])calloc(NB * 3, sizeof(float[QMAX]));
  float (*Ix)[QMAX] = Mx + NB, (*Dx)[QMAX] = Ix + NB;
  float md[NB][QMAX], esum[NB], Bs[NB], Ts[NB], Bn[NB], Tn[NB];
  float *dpc = fmx_row(ox, 0);
  for (int k = 0; k <= M; k++) dpc[k * 3 + C_M] = dpc[k * 3 + C_D] = dpc[k * 3 + C_I] = 0.0f;
  float xE = 0.0f, xN = 1.0f, xJ = 0.0f, xB = sp.nmove, xC = 0.0f;
  ox->xmx[X_E] = xE; ox->xmx[X_N] = xN; ox->xmx[X_J] = xJ; ox->xmx[X_B] = xB; ox->xmx[X_C] = xC; ox->xmx[X_SCALE] = 1.0f;
  ox->totscale = 0.0f;
  for (int i = 1; i <= L; i++) {
    const float *rp = p->rfv + (size_t)dsq[i - 1] * (M + 1);
    float pm_in[NB], pi_in[NB], pd_in[NB];
    for (int l = 0; l < NB; l++) {
      pm_in[l] = l ? Mx[l - 1][Q - 1] : 0.0f; pi_in[l] = l ? Ix[l - 1][Q - 1] : 0.0f; pd_in[l] = l ? Dx[l - 1][Q - 1] : 0.0f;
    }
    for (int l = 0; l < NB; l++) {
      float es = 0.0f;
      for (int q = Q - 1; q >= 0; q--) {
        const int k = l * Q + q + 1;
        const float *tf = (k <= M) ? p->tfv + k * 8 : zero8;
        const float e = (k <= M) ? rp[k] : 0.0f;
        const float pm = (q > 0) ? Mx[l][q - 1] : pm_in[l], pi = (q > 0) ? Ix[l][q - 1] : pi_in[l], pd = (q > 0) ? Dx[l][q - 1] : pd_in[l];
        float sv = xB * tf[T_BM];
        sv += pm * tf[T_MM];
        sv += pi * tf[T_IM];
        sv += pd * tf[T_DM];
        sv *= e;
        const float nI = Mx[l][q] * tf[T_MI] + Ix[l][q] * tf[T_II];
        md[l][q] = sv * tf[T_MD];
        Mx[l][q] = sv; Ix[l][q] = nI;
        es += sv;
      }
      esum[l] = es;
      float Bb = 0.0f, Tb = 1.0f;
      for (int q = 0; q < Q; q++) {
        const int k = l * Q + q + 1;
        const float tdd = (k <= M) ? p->tfv[k * 8 + T_DD] : 0.0f;
        Bb = md[l][q] + Bb * tdd; Tb *= tdd;
      }
      Bs[l] = Bb; Ts[l] = Tb;
    }
    for (int o = 1; o < NB; o <<= 1) {
      for (int l = 0; l < NB; l++) {
        if (l >= o) { Bn[l] = Bs[l] + Bs[l - o] * Ts[l]; Tn[l] = Ts[l] * Ts[l - o]; }
        else { Bn[l] = Bs[l]; Tn[l] = Ts[l]; }
      }
      memcpy(Bs, Bn, sizeof(Bs)); memcpy(Ts, Tn, sizeof(Ts));
    }
    for (int l = 0; l < NB; l++) {
      float d = l ? Bs[l - 1] : 0.0f;
      for (int q = 0; q < Q; q++) {
        const int k = l * Q + q + 1;
        const float tdd = (k <= M) ? p->tfv[k * 8 + T_DD] : 0.0f;
        Dx[l][q] = (k <= M) ? d : 0.0f;
        esum[l] += Dx[l][q];
        d = md[l][q] + d * tdd;
      }
    }
    xE = butterfly_sum(esum);
    xN = xN * sp.nloop;
    xC = (xC * sp.nloop) + (xE * sp.emove);
    xJ = (xJ * sp.nloop) + (xE * sp.eloop);
    xB = (xJ * sp.nmove) + (xN * sp.nmove);
    float scale = 1.0f;
    if (xE > 1.0e4f) {
      scale = xE;
      const float inv = 1.0f / xE;
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      for (int l = 0; l < NB; l++) for (int q = 0; q < Q; q++) { Mx[l][q] *= inv; Dx[l][q] *= inv; Ix[l][q] *= inv; }
      ox->totscale += (float)log((double)xE);
      xE = 1.0f;
    }
    if (ox->full) {
      dpc = fmx_row(ox, i);
      dpc[C_M] = dpc[C_D] = dpc[C_I] = 0.0f;
      for (int k = 1; k <= M; k++) { const int l = (k - 1) / Q, q = (k - 1) % Q; dpc[k * 3 + C_M] = Mx[l][q]; dpc[k * 3 + C_D] = Dx[l][q]; dpc[k * 3 + C_I] = Ix[l][q]; }
    }
    float *xr = ox->xmx + i * X_NX;
    xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = scale;
  }
  free(Mx);
  if (isnan(xC) || (L > 0 && xC == 0.0f) || isinf(xC)) { if (ret_sc) *ret_sc = -INFINITY; return 1; }
  if (ret_sc) *ret_sc = ox->totscale + (float)log((double)(xC * sp.nmove));
  return 0;
}

static int backward_engine_blk(const orc_profile *p, const uint8_t *dsq, int L, specials sp, const fmx *fwd, fmx *bck, float *ret_sc, int Q)
{
  const int M = p->M;
  static const float zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float (*Mx)[QMAX] = (float (*)[QMAX])calloc(NB * 3, sizeof(float[QMAX]));
  float (*Ix)[QMAX] = Mx + NB, (*Dx)[QMAX] = Ix + NB;
  float em[NB][QMAX], cb[NB][QMAX], part[NB], Bs[NB], Ts[NB], Bn[NB], Tn[NB];
  float xC = 0.0f, xE = 0.0f, xJ = 0.0f, xN = 0.0f, xB = 0.0f;
  bck->totscale = 0.0f;
  for (int i = L; i >= 0; i--) {
    const float *rp = (i < L) ? p->rfv + (size_t)dsq[i] * (M + 1) : NULL;
    if (i == L) {
      xC = sp.nmove; xE = xC * sp.emove; xB = 0.0f; xJ = 0.0f; xN = 0.0f;
      for (int l = 0; l < NB; l++) for (int q = 0; q < Q; q++) em[l][q] = 0.0f;
    } else {
      for (int l = 0; l < NB; l++) {
        float pt = 0.0f;
        for (int q = 0; q < Q; q++) {
          const int k = l * Q + q + 1;
          const float e = (k <= M) ? rp[k] : 0.0f, tbm = (k <= M) ? p->tfv[k * 8 + T_BM] : 0.0f;
          em[l][q] = Mx[l][q] * e; pt += em[l][q] * tbm;
        }
        part[l] = pt;
      }
      xB = butterfly_sum(part);
      xC = xC * sp.nloop;
      xJ = (xB * sp.nmove) + (xJ * sp.nloop);
      xN = (xB * sp.nmove) + (xN * sp.nloop);
      xE = (xC * sp.emove) + (xJ * sp.eloop);
    }
    const float s = (i >= 1) ? fwd->xmx[i * X_NX + X_SCALE] : 1.0f;
    float *dpc = fmx_row(bck, i);
    if (i >= 1) {
      const float inv = (s > 1.0f) ? 1.0f / s : 1.0f;
      for (int l = 0; l < NB; l++) {
        const float em_next = (l < NB - 1) ? em[l + 1][0] : 0.0f;
        const int kn = (l + 1) * Q + 1;                        /* first position of the next block */
        const float *tn = (l < NB - 1 && kn <= M) ? p->tfv + kn * 8 : zero8;
        float Bb = 0.0f, Tb = 1.0f;
        for (int q = Q - 1; q >= 0; q--) {
          const int k = l * Q + q + 1;
          const float mnext = (q < Q - 1) ? em[l][q + 1] : em_next;
          const float *t0n = (q < Q - 1) ? ((k + 1 <= M) ? p->tfv + (k + 1) * 8 : zero8) : tn;
          const float tdm = t0n[T_DM];
          const int in = (k <= M);
          cb[l][q] = in ? (xE + ((k < M) ? mnext * tdm : 0.0f)) : 0.0f;
          const float tdd = in ? p->tfv[k * 8 + T_DD] : 0.0f;
          Bb = cb[l][q] + Bb * tdd; Tb *= tdd;
        }
        Bs[l] = Bb; Ts[l] = Tb;
      }
      for (int o = 1; o < NB; o <<= 1) {
        for (int l = 0; l < NB; l++) {
          if (l + o < NB) { Bn[l] = Bs[l] + Bs[l + o] * Ts[l]; Tn[l] = Ts[l] * Ts[l + o]; }
          else { Bn[l] = Bs[l]; Tn[l] = Ts[l]; }
        }
        memcpy(Bs, Bn, sizeof(Bs)); memcpy(Ts, Tn, sizeof(Ts));
      }
      float nM[QMAX], nI[QMAX], nD[QMAX];
      for (int l = 0; l < NB; l++) {
        const float em_next = (l < NB - 1) ? em[l + 1][0] : 0.0f;
        const int kn = (l + 1) * Q + 1;
        const float *tn = (l < NB - 1 && kn <= M) ? p->tfv + kn * 8 : zero8;
        float dnext = (l < NB - 1) ? Bs[l + 1] : 0.0f;
        for (int q = Q - 1; q >= 0; q--) {
          const int k = l * Q + q + 1;
          const int in = (k <= M);
          const float *t1 = in ? p->tfv + k * 8 : zero8;
          const float mnext = (q < Q - 1) ? em[l][q + 1] : em_next;
          const float *t0n = (q < Q - 1) ? ((k + 1 <= M) ? p->tfv + (k + 1) * 8 : zero8) : tn;
          const float tmm = (k < M) ? t0n[T_MM] : 0.0f, tim = (k < M) ? t0n[T_IM] : 0.0f;
          const float inext = (i < L) ? Ix[l][q] : 0.0f;
          const float dv = in ? (cb[l][q] + dnext * t1[T_DD]) : 0.0f;
          float mv = xE + mnext * tmm + inext * t1[T_MI] + dnext * t1[T_MD];
          float iv = mnext * tim + inext * t1[T_II];
          if (!in) { mv = 0.0f; iv = 0.0f; }
          nM[q] = mv * inv; nI[q] = iv * inv; nD[q] = dv * inv;
          dnext = dv;
        }
        for (int q = 0; q < Q; q++) { Mx[l][q] = nM[q]; Ix[l][q] = nI[q]; Dx[l][q] = nD[q]; }
      }
      if (s > 1.0f) { xE = xE / s; xN = xN / s; xJ = xJ / s; xB = xB / s; xC = xC / s; }
    } else {
      xC = 0.0f; xJ = 0.0f; xE = 0.0f;
      for (int l = 0; l < NB; l++) for (int q = 0; q < Q; q++) { Mx[l][q] = 0.0f; Ix[l][q] = 0.0f; Dx[l][q] = 0.0f; }
    }
    dpc[C_M] = dpc[C_D] = dpc[C_I] = 0.0f;
    for (int k = 1; k <= M; k++) { const int l = (k - 1) / Q, q = (k - 1) % Q; dpc[k * 3 + C_M] = Mx[l][q]; dpc[k * 3 + C_D] = Dx[l][q]; dpc[k * 3 + C_I] = Ix[l][q]; }
    bck->totscale += (float)log((double)s);
    float *xr = bck->xmx + i * X_NX;
    xr[X_E] = xE; xr[X_N] = xN; xr[X_J] = xJ; xr[X_B] = xB; xr[X_C] = xC; xr[X_SCALE] = s;
  }
  free(Mx);
  if (ret_sc) *ret_sc = bck->totscale + (float)log((double)xN);
  return 0;
}

static int forward_engine(const orc_profile *p, const uint8_t *dsq, int L, specials sp, fmx *ox, float *ret_sc)
{
  const int Q = (g_order == ORC_ORDER_CANONICAL) ? orc_block_width(p->M) : 0;
  return Q ? forward_engine_blk(p, dsq, L, sp, ox, ret_sc, Q) : forward_engine_seq(p, dsq, L, sp, ox, ret_sc);
}
static int backward_engine(const orc_profile *p, const uint8_t *dsq, int L, specials sp, const fmx *fwd, fmx *bck, float *ret_sc)
{
  const int Q = (g_order == ORC_ORDER_CANONICAL) ? orc_block_width(p->M) : 0;
  return Q ? backward_engine_blk(p, dsq, L, sp, fwd, bck, ret_sc, Q) : backward_engine_seq(p, dsq, L, sp, fwd, bck, ret_sc);
}

int orc_forward_parser(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc)
{
  fmx *f = fmx_create(p->M, L, 0);
  int st = forward_engine(p, dsq, L, make_specials(L, 1), f, ret_sc);
  fmx_free(f);
  return st;
}
int orc_backward_parser(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc)
{
  fmx *f = fmx_create(p->M, L, 0), *b = fmx_create(p->M, L, 0);
  float fsc;
  forward_engine(p, dsq, L, make_specials(L, 1), f, &fsc);
  int st = backward_engine(p, dsq, L, make_specials(L, 1), f, b, ret_sc);
  fmx_free(f); fmx_free(b);
  return st;
}

/* =====================================================================================
 * Filters in pipeline order (A.5 steps 1-4)
 * ===================================================================================== */
int orc_filters(const orc_profile *p, const uint8_t *dsq, int L, orc_filter_result *r)
{
  const float *ev = p->hmm->evparam;
  memset(r, 0, sizeof(*r));
  r->vit_sc = NAN; r->fwd_sc = NAN; r->filtersc = NAN;
  if (L == 0) return 0;
  r->nullsc = orc_null1(L);
  if (p->striped) orc_msv_simd(p, p->striped, dsq, L, &r->msv_sc, &r->msv_xJ);
  else            orc_msv(p, dsq, L, &r->msv_sc, &r->msv_xJ);
  float seq_score = (r->msv_sc - r->nullsc) / (float)LOG2C;
  double P = gumbel_surv(seq_score, ev[ORC_MMU], ev[ORC_MLAMBDA]);
  if (P > F1_DEFAULT) return 0;
  r->passed_msv = 1;
  r->filtersc = orc_biasfilter(p, dsq, L);
  seq_score = (r->msv_sc - r->filtersc) / (float)LOG2C;
  P = gumbel_surv(seq_score, ev[ORC_MMU], ev[ORC_MLAMBDA]);
  if (P > F1_DEFAULT) return 0;
  r->passed_bias = 1;
  if (P > F2_DEFAULT) {
    if (p->striped) orc_vitfilter_simd(p, p->striped, dsq, L, &r->vit_sc);
    else            orc_vitfilter(p, dsq, L, &r->vit_sc);
    seq_score = (r->vit_sc - r->filtersc) / (float)LOG2C;
    P = gumbel_surv(seq_score, ev[ORC_VMU], ev[ORC_VLAMBDA]);
    if (P > F2_DEFAULT) return 0;
  }
  r->passed_vit = 1;
  orc_forward_parser(p, dsq, L, &r->fwd_sc);
  seq_score = (r->fwd_sc - r->filtersc) / (float)LOG2C;
  P = exp_surv(seq_score, ev[ORC_FTAU], ev[ORC_FLAMBDA]);
  if (P > F3_DEFAULT) return 0;
  r->passed_fwd = 1;
  return 1;
}

/* =====================================================================================
 * Domain definition by posterior heuristics (A.5 step 5)
 * ===================================================================================== */
typedef struct {
  int    L;
  float *btot, *etot, *mocc, *n2sc;
  int    nregions, nclustered, nenvelopes, noverlaps, ndom, dom_alloc;
  orc_domain *dcl;
} ddef_t;

static void domain_decoding(const fmx *oxf, const fmx *oxb, specials sp, ddef_t *dd)
{
  int L = oxf->L;
  float scaleproduct = 1.0f / oxb->xmx[X_N];
  dd->btot[0] = dd->etot[0] = dd->mocc[0] = 0.0f;
  for (int i = 1; i <= L; i++) {
    const float *f0 = oxf->xmx + (i - 1) * X_NX, *f1 = oxf->xmx + i * X_NX;
    const float *b0 = oxb->xmx + (i - 1) * X_NX, *b1 = oxb->xmx + i * X_NX;
    dd->btot[i] = dd->btot[i - 1] + (f0[X_B] * b0[X_B]) * f0[X_SCALE] * scaleproduct;
    dd->etot[i] = dd->etot[i - 1] + (f1[X_E] * b1[X_E]) * f1[X_SCALE] * scaleproduct;
    float njcp;
    njcp  = f0[X_N] * b1[X_N] * sp.nloop * scaleproduct;
    njcp += f0[X_J] * b1[X_J] * sp.nloop * scaleproduct;
    njcp += f0[X_C] * b1[X_C] * sp.nloop * scaleproduct;
    dd->mocc[i] = 1.0f - njcp;
  }
}

static int is_multidomain_region(const ddef_t *dd, int i, int j)
{
  float max = -1.0f;
  for (int z = i; z <= j; z++) {
    float a = dd->etot[z] - dd->etot[i - 1], b = dd->btot[j] - dd->btot[z - 1];
    float e = a < b ? a : b;
    if (e > max) max = e;
  }
  return max >= 0.20f;
}

/* posterior decoding of a full matrix pair; pp overwrites bck.  Returns 1 on range error. */
static int decoding(specials sp, const fmx *oxf, fmx *oxb)
{
  int M = oxf->M, L = oxf->L;
  float scaleproduct = 1.0f / oxb->xmx[X_N];
  float *pp0 = fmx_row(oxb, 0);
  for (int k = 0; k <= M; k++) pp0[k * 3 + C_M] = pp0[k * 3 + C_D] = pp0[k * 3 + C_I] = 0.0f;
  /* specials need b(i) untouched while computing row i, and f(i-1): walk forward, overwrite in place */
  float *ppx = (float *)calloc((size_t)(L + 1) * X_NX, sizeof(float));
  for (int i = 1; i <= L; i++) {
    const float *fr = fmx_row(oxf, i);
    float *br = fmx_row(oxb, i);
    float totr = scaleproduct * oxf->xmx[i * X_NX + X_SCALE];
    for (int k = 1; k <= M; k++) {
      br[k * 3 + C_M] = fr[k * 3 + C_M] * br[k * 3 + C_M] * totr;
      br[k * 3 + C_D] = 0.0f;
      br[k * 3 + C_I] = fr[k * 3 + C_I] * br[k * 3 + C_I] * totr;
    }
    br[C_M] = br[C_D] = br[C_I] = 0.0f;
    const float *f0 = oxf->xmx + (i - 1) * X_NX, *b1 = oxb->xmx + i * X_NX;
    ppx[i * X_NX + X_N] = f0[X_N] * b1[X_N] * sp.nloop * scaleproduct;
    ppx[i * X_NX + X_J] = f0[X_J] * b1[X_J] * sp.nloop * scaleproduct;
    ppx[i * X_NX + X_C] = f0[X_C] * b1[X_C] * sp.nloop * scaleproduct;
  }
  memcpy(oxb->xmx, ppx, sizeof(float) * (size_t)(L + 1) * X_NX);
  free(ppx);
  return isinf(scaleproduct) ? 1 : 0;
}

/* null2 by expectation over the posterior matrix pp (rows 1..Ld) */
static void null2_by_expectation(const orc_profile *p, const fmx *pp, float *null2)
{
  int M = p->M, Ld = pp->L;
  float *em = (float *)calloc((size_t)(M + 1) * 2, sizeof(float)), *ei = em + (M + 1);
  float xn = 0, xc = 0, xj = 0;
  {
    const float *r = fmx_row(pp, 1);
    for (int k = 1; k <= M; k++) { em[k] = r[k * 3 + C_M]; ei[k] = r[k * 3 + C_I]; }
    xn = pp->xmx[1 * X_NX + X_N]; xc = pp->xmx[1 * X_NX + X_C]; xj = pp->xmx[1 * X_NX + X_J];
  }
  for (int i = 2; i <= Ld; i++) {
    const float *r = fmx_row(pp, i);
    for (int k = 1; k <= M; k++) { em[k] += r[k * 3 + C_M]; ei[k] += r[k * 3 + C_I]; }
    xn += pp->xmx[i * X_NX + X_N]; xc += pp->xmx[i * X_NX + X_C]; xj += pp->xmx[i * X_NX + X_J];
  }
  float norm = 1.0f / (float)Ld;
  for (int k = 1; k <= M; k++) { em[k] *= norm; ei[k] *= norm; }
  xn *= norm; xc *= norm; xj *= norm;
  float xfactor = xn + xc + xj;
  const int Q = (g_order == ORC_ORDER_CANONICAL) ? orc_block_width(M) : 0;
  for (int x = 0; x < ORC_K; x++) {
    const float *rp = p->rfv + (size_t)x * (M + 1);
    float sv = 0.0f;
    if (Q) {                                             /* 32 block sums, then the butterfly */
      float part[NB];
      for (int l = 0; l < NB; l++) {
        float pt = 0.0f;
        for (int q = 0; q < Q; q++) { const int k = l * Q + q + 1; if (k <= M) { pt += em[k] * rp[k]; pt += ei[k]; } else { pt += 0.0f; pt += 0.0f; } }
        part[l] = pt;
      }
      sv = butterfly_sum(part);
    } else {
      for (int k = 1; k <= M; k++) { sv += em[k] * rp[k]; sv += ei[k]; }
    }
    null2[x] = sv + xfactor;
  }
  for (int x = ORC_K + 1; x <= ORC_KP - 3; x++) {      /* degenerate: plain average of the odds */
    float result = 0.0f; int n = 0;
    for (int i = 0; i < ORC_K; i++) if (degen_has(x, i)) { result += null2[i]; n++; }
    null2[x] = result / (float)n;
  }
  null2[20] = 1.0f; null2[27] = 1.0f; null2[28] = 1.0f;
  free(em);
}

/* optimal accuracy fill; pp in, ox out (full matrices, same shape). returns score */
static float optimal_accuracy(const orc_profile *p, specials sp, const fmx *pp, fmx *ox)
{
  int M = p->M, L = pp->L;
  const float ninf = -INFINITY;
  float *r0 = fmx_row(ox, 0);
  for (int k = 0; k <= M; k++) r0[k * 3 + C_M] = r0[k * 3 + C_D] = r0[k * 3 + C_I] = ninf;
  float *x0 = ox->xmx;
  x0[X_E] = ninf; x0[X_N] = 0.0f; x0[X_J] = ninf; x0[X_C] = ninf;
  x0[X_B] = (sp.nmove > 0.0f) ? 0.0f : ninf;
  for (int i = 1; i <= L; i++) {
    const float *dpp = fmx_row(ox, i - 1), *ppr = fmx_row(pp, i);
    float *dpc = fmx_row(ox, i);
    float xBp = ox->xmx[(i - 1) * X_NX + X_B];
    dpc[C_M] = dpc[C_D] = dpc[C_I] = ninf;
    float xE = ninf, dcv = ninf;
    for (int k = 1; k <= M; k++) {
      const float *tf = p->tfv + k * 8;
      float sv = (tf[T_BM] > 0.0f) ? xBp : 0.0f;            /* masked-out paths contribute 0, not -inf */
      float c;
      c = (tf[T_MM] > 0.0f) ? dpp[(k - 1) * 3 + C_M] : 0.0f; if (c > sv) sv = c;
      c = (tf[T_IM] > 0.0f) ? dpp[(k - 1) * 3 + C_I] : 0.0f; if (c > sv) sv = c;
      c = (tf[T_DM] > 0.0f) ? dpp[(k - 1) * 3 + C_D] : 0.0f; if (c > sv) sv = c;
      sv += ppr[k * 3 + C_M];
      if (sv > xE) xE = sv;
      dpc[k * 3 + C_M] = sv;
      dpc[k * 3 + C_D] = dcv;
      { float a = (tf[T_MD] > 0.0f) ? sv : 0.0f, b = (tf[T_DD] > 0.0f) ? dcv : 0.0f;
        /* D(i,k+1) = max(M(i,k), D(i,k)) through allowed transitions */
        dcv = (k < M) ? (a > b ? a : b) : ninf; }
      { float a = (tf[T_MI] > 0.0f) ? dpp[k * 3 + C_M] : 0.0f, b = (tf[T_II] > 0.0f) ? dpp[k * 3 + C_I] : 0.0f;
        dpc[k * 3 + C_I] = (a > b ? a : b) + ppr[k * 3 + C_I]; }
    }
    for (int k = 1; k <= M; k++) if (dpc[k * 3 + C_D] > xE) xE = dpc[k * 3 + C_D];
    const float *xp = ox->xmx + (i - 1) * X_NX, *ppx = pp->xmx + i * X_NX;
    float *xc = ox->xmx + i * X_NX;
    float t1, t2;
    xc[X_E] = xE;
    t1 = (sp.nloop == 0.0f) ? FLT_MIN : 1.0f; t2 = (sp.eloop == 0.0f) ? FLT_MIN : 1.0f;
    { float a = t1 * (xp[X_J] + ppx[X_J]), b = t2 * xE; xc[X_J] = a > b ? a : b; }
    t2 = (sp.emove == 0.0f) ? FLT_MIN : 1.0f;
    { float a = t1 * (xp[X_C] + ppx[X_C]), b = t2 * xE; xc[X_C] = a > b ? a : b; }
    xc[X_N] = t1 * (xp[X_N] + ppx[X_N]);
    t1 = (sp.nmove == 0.0f) ? FLT_MIN : 1.0f;
    { float a = t1 * xc[X_N], b = t1 * xc[X_J]; xc[X_B] = a > b ? a : b; }
  }
  return ox->xmx[L * X_NX + X_C];
}

enum { ST_M = 1, ST_D, ST_I, ST_S, ST_N, ST_B, ST_E, ST_C, ST_T, ST_J };

/* OA traceback; we only need the first B..E segment: first/last M state. */
static int oa_trace(const orc_profile *p, specials sp, const fmx *pp, const fmx *ox,
                    int *hmmfrom, int *hmmto, int *sqfrom, int *sqto, int *trace /* optional, [L+1]: +k match, -k insert */)
{
  int M = p->M, i = ox->L, k = 0, s0 = ST_C, s1;
  int firstM_i = 0, firstM_k = 0, lastM_i = 0, lastM_k = 0, have_last = 0, guard = 0;
  int ndom_seen = 0;
  (void)M;
  while (s0 != ST_S) {
    if (++guard > 4 * (ox->L + p->M) + 16) return 1;
    const float *xc = ox->xmx + i * X_NX;
    switch (s0) {
    case ST_M: {
      const float *dpp = fmx_row(ox, i - 1); const float *tf = p->tfv + k * 8;
      float path[4];
      path[0] = (tf[T_MM] > 0.0f) ? dpp[(k - 1) * 3 + C_M] : -INFINITY;
      path[1] = (tf[T_IM] > 0.0f) ? dpp[(k - 1) * 3 + C_I] : -INFINITY;
      path[2] = (tf[T_DM] > 0.0f) ? dpp[(k - 1) * 3 + C_D] : -INFINITY;
      path[3] = (tf[T_BM] > 0.0f) ? ox->xmx[(i - 1) * X_NX + X_B] : -INFINITY;
      int best = 0; for (int z = 1; z < 4; z++) if (path[z] > path[best]) best = z;
      static const int st[4] = { ST_M, ST_I, ST_D, ST_B };
      s1 = st[best]; k--; i--; break; }
    case ST_D: {
      const float *dpc = fmx_row(ox, i); const float *tf = p->tfv + (k - 1) * 8;
      float a = (tf[T_MD] > 0.0f) ? dpc[(k - 1) * 3 + C_M] : -INFINITY;
      float b = (tf[T_DD] > 0.0f) ? dpc[(k - 1) * 3 + C_D] : -INFINITY;
      s1 = (a >= b) ? ST_M : ST_D; k--; break; }
    case ST_I: {
      const float *dpp = fmx_row(ox, i - 1); const float *tf = p->tfv + k * 8;
      float a = (tf[T_MI] > 0.0f) ? dpp[k * 3 + C_M] : -INFINITY;
      float b = (tf[T_II] > 0.0f) ? dpp[k * 3 + C_I] : -INFINITY;
      s1 = (a >= b) ? ST_M : ST_I; i--; break; }
    case ST_N: s1 = (i == 0) ? ST_S : ST_N; break;
    case ST_C: {
      float t1 = (sp.nloop == 0.0f) ? FLT_MIN : 1.0f, t2 = (sp.emove == 0.0f) ? FLT_MIN : 1.0f;
      float a = (i > 0) ? t1 * (ox->xmx[(i - 1) * X_NX + X_C] + pp->xmx[i * X_NX + X_C]) : -INFINITY;
      float b = t2 * xc[X_E];
      s1 = (a > b) ? ST_C : ST_E; break; }
    case ST_J: {
      float t1 = (sp.nloop == 0.0f) ? FLT_MIN : 1.0f, t2 = (sp.eloop == 0.0f) ? FLT_MIN : 1.0f;
      float a = (i > 0) ? t1 * (ox->xmx[(i - 1) * X_NX + X_J] + pp->xmx[i * X_NX + X_J]) : -INFINITY;
      float b = t2 * xc[X_E];
      s1 = (a > b) ? ST_J : ST_E; break; }
    case ST_E: {
      const float *dpc = fmx_row(ox, i);
      float max = -INFINITY; int smax = -1, kmax = -1;
      for (int kk = 1; kk <= p->M; kk++) {
        if (dpc[kk * 3 + C_M] >= max) { max = dpc[kk * 3 + C_M]; smax = ST_M; kmax = kk; }
      }
      for (int kk = 1; kk <= p->M; kk++)
        if (dpc[kk * 3 + C_D] > max) { max = dpc[kk * 3 + C_D]; smax = ST_D; kmax = kk; }
      s1 = smax; k = kmax; ndom_seen++; break; }
    case ST_B: {
      float t1 = (sp.nmove == 0.0f) ? FLT_MIN : 1.0f;
      float a = t1 * xc[X_N], b = t1 * xc[X_J];
      s1 = (a > b) ? ST_N : ST_J; break; }
    default: return 1;
    }
    if (s1 == -1) return 1;
    /* traceback runs right to left: the LAST domain in the sequence is seen first; the alignment
     * display takes domain 0 = the leftmost, so keep overwriting until the walk ends. */
    if (s1 == ST_M) {
      if (!have_last || s0 == ST_E) { lastM_i = i; lastM_k = k; have_last = 1; }
      firstM_i = i; firstM_k = k;
      if (trace) trace[i] = k;
    } else if (s1 == ST_I) {
      if (trace) trace[i] = -k;
    }
    if ((s1 == ST_N || s1 == ST_J || s1 == ST_C) && s1 == s0) i--;
    s0 = s1;
  }
  if (!have_last) return 1;
  *hmmfrom = firstM_k; *hmmto = lastM_k; *sqfrom = firstM_i; *sqto = lastM_i;
  return 0;
}

/* =====================================================================================
 * Stochastic traceback ensemble + single-linkage clustering (multi-domain regions only)
 * ===================================================================================== */
typedef struct { uint32_t x, seed; } lcg_t;
static uint32_t mix3(uint32_t a, uint32_t b, uint32_t c)
{
  a -= b; a -= c; a ^= (c >> 13);
  b -= c; b -= a; b ^= (a << 8);
  c -= a; c -= b; c ^= (b >> 13);
  a -= b; a -= c; a ^= (c >> 12);
  b -= c; b -= a; b ^= (a << 16);
  c -= a; c -= b; c ^= (b >> 5);
  a -= b; a -= c; a ^= (c >> 3);
  b -= c; b -= a; b ^= (a << 10);
  c -= a; c -= b; c ^= (b >> 15);
  return c;
}
static void lcg_init(lcg_t *r, uint32_t seed) { r->seed = seed; r->x = mix3(seed, 87654321u, 12345678u); if (r->x == 0) r->x = 42; }
static double lcg_random(lcg_t *r) { r->x *= 69069u; r->x += 1u; return (double)r->x / 4294967296.0; }

static int fchoose(lcg_t *r, float *p, int N)
{
  float sum = 0.0f;
  for (int i = 0; i < N; i++) sum += p[i];
  if (sum != 0.0f) { float s = (float)(1.0 / sum); for (int i = 0; i < N; i++) p[i] *= s; }
  else for (int i = 0; i < N; i++) p[i] = 1.0f / (float)N;
  float roll = (float)lcg_random(r);
  sum = 0.0f;
  for (int i = 0; i < N; i++) { sum += p[i]; if (roll < sum) return i; }
  int i;
  do { i = (int)(lcg_random(r) * N); } while (p[i] == 0.0f);
  return i;
}

typedef struct { int N, alloc; int8_t *st; int *k, *i; } trace_t;
static void trace_append(trace_t *t, int st, int k, int i)
{
  if (t->N == t->alloc) {
    t->alloc = t->alloc ? t->alloc * 2 : 256;
    t->st = (int8_t *)realloc(t->st, t->alloc);
    t->k = (int *)realloc(t->k, sizeof(int) * t->alloc);
    t->i = (int *)realloc(t->i, sizeof(int) * t->alloc);
  }
  t->st[t->N] = (int8_t)st;
  /* only emitting/model states carry coordinates */
  t->k[t->N] = (st == ST_M || st == ST_D || st == ST_I) ? k : 0;
  t->i[t->N] = i;
  t->N++;
}

static int stochastic_trace(lcg_t *rng, const orc_profile *p, specials sp, const fmx *ox, trace_t *tr)
{
  int M = p->M, i = ox->L, k = 0, s0, s1;
  int Q = (M + 3) / 4; if (Q < 2) Q = 2;
  tr->N = 0;
  trace_append(tr, ST_T, 0, i);
  trace_append(tr, ST_C, 0, i);
  s0 = ST_C;
  long guard = 0;
  while (s0 != ST_S) {
    if (++guard > 8L * (ox->L + 2) * (M + 2)) return 1;
    float path[4];
    switch (s0) {
    case ST_M: {
      const float *dpp = fmx_row(ox, i - 1); const float *tf = p->tfv + k * 8;
      path[0] = ox->xmx[(i - 1) * X_NX + X_B] * tf[T_BM];
      path[1] = dpp[(k - 1) * 3 + C_M] * tf[T_MM];
      path[2] = dpp[(k - 1) * 3 + C_I] * tf[T_IM];
      path[3] = dpp[(k - 1) * 3 + C_D] * tf[T_DM];
      static const int st[4] = { ST_B, ST_M, ST_I, ST_D };
      s1 = st[fchoose(rng, path, 4)]; k--; i--; break; }
    case ST_D: {
      const float *dpc = fmx_row(ox, i); const float *tf = p->tfv + (k - 1) * 8;
      path[0] = dpc[(k - 1) * 3 + C_M] * tf[T_MD];
      path[1] = dpc[(k - 1) * 3 + C_D] * tf[T_DD];
      s1 = fchoose(rng, path, 2) == 0 ? ST_M : ST_D; k--; break; }
    case ST_I: {
      const float *dpp = fmx_row(ox, i - 1); const float *tf = p->tfv + k * 8;
      path[0] = dpp[k * 3 + C_M] * tf[T_MI];
      path[1] = dpp[k * 3 + C_I] * tf[T_II];
      s1 = fchoose(rng, path, 2) == 0 ? ST_M : ST_I; i--; break; }
    case ST_N: s1 = (i == 0) ? ST_S : ST_N; break;
    case ST_C:
      path[0] = (i > 0) ? ox->xmx[(i - 1) * X_NX + X_C] * sp.nloop : 0.0f;
      path[1] = ox->xmx[i * X_NX + X_E] * sp.emove * ox->xmx[i * X_NX + X_SCALE];   /* row i may have been rescaled */
      s1 = fchoose(rng, path, 2) == 0 ? ST_C : ST_E; break;
    case ST_J:
      path[0] = (i > 0) ? ox->xmx[(i - 1) * X_NX + X_J] * sp.nloop : 0.0f;
      path[1] = ox->xmx[i * X_NX + X_E] * sp.eloop * ox->xmx[i * X_NX + X_SCALE];
      s1 = fchoose(rng, path, 2) == 0 ? ST_J : ST_E; break;
    case ST_B:
      path[0] = ox->xmx[i * X_NX + X_N] * sp.nmove;
      path[1] = ox->xmx[i * X_NX + X_J] * sp.nmove;
      s1 = fchoose(rng, path, 2) == 0 ? ST_N : ST_J; break;
    case ST_E: {
      const float *dpc = fmx_row(ox, i);
      double sum = 0.0, roll = lcg_random(rng);
      float norm = (float)(1.0 / ox->xmx[i * X_NX + X_E]);
      s1 = -1;
      for (int pass = 0; pass < 2 && s1 == -1; pass++) {
        for (int q = 0; q < Q && s1 == -1; q++) {           /* striped enumeration order of the SIMD original */
          for (int r = 0; r < 4; r++) { int kk = r * Q + q + 1; float v = (kk <= M) ? dpc[kk * 3 + C_M] * norm : 0.0f;
            sum += v; if (roll < sum) { k = kk; s1 = ST_M; break; } }
          if (s1 != -1) break;
          for (int r = 0; r < 4; r++) { int kk = r * Q + q + 1; float v = (kk <= M) ? dpc[kk * 3 + C_D] * norm : 0.0f;
            sum += v; if (roll < sum) { k = kk; s1 = ST_D; break; } }
        }
        if (s1 == -1 && sum < 0.99) return 1;
      }
      if (s1 == -1) return 1;
      break; }
    default: return 1;
    }
    trace_append(tr, s1, k, i);
    if ((s1 == ST_N || s1 == ST_J || s1 == ST_C) && s1 == s0) i--;
    s0 = s1;
  }
  /* reverse */
  for (int a = 0, b = tr->N - 1; a < b; a++, b--) {
    int8_t s = tr->st[a]; tr->st[a] = tr->st[b]; tr->st[b] = s;
    int t = tr->k[a]; tr->k[a] = tr->k[b]; tr->k[b] = t;
    t = tr->i[a]; tr->i[a] = tr->i[b]; tr->i[b] = t;
  }
  /* N/C/J emit on transition: in a run of n visits the first is silent and visit v emits the residue
   * recorded with visit v-1 of the (reversed) walk; shift the recorded coordinates right by one. */
  for (int z = 0; z < tr->N; z++) {
    int s = tr->st[z];
    if (s == ST_S || s == ST_B || s == ST_E || s == ST_T || s == ST_D) tr->i[z] = 0;
  }
  for (int z = 0; z < tr->N; ) {
    int s = tr->st[z];
    if (s == ST_N || s == ST_C || s == ST_J) {
      int z2 = z; while (z2 + 1 < tr->N && tr->st[z2 + 1] == s) z2++;
      for (int y = z2; y > z; y--) tr->i[y] = tr->i[y - 1];
      tr->i[z] = 0;
      z = z2 + 1;
    } else z++;
  }
  return 0;
}

typedef struct { int idx, i, j, k, m; float prob; } spcoord;

static int sp_link(const spcoord *h1, const spcoord *h2)
{
  int nov, n, d1, d2;
  nov = (h1->j < h2->j ? h1->j : h2->j) - (h1->i > h2->i ? h1->i : h2->i) + 1;
  { int a = h1->j - h1->i + 1, b = h2->j - h2->i + 1; n = a < b ? a : b; }
  if ((float)nov / (float)n < 0.8f) return 0;
  nov = (h1->m < h2->m ? h1->m : h2->m) - (h1->k > h2->k ? h1->k : h2->k);
  { int a = h1->m - h1->k + 1, b = h2->m - h2->k + 1; n = a < b ? a : b; }
  if ((float)nov / (float)n < 0.8f) return 0;
  d1 = h1->i - h1->k; d2 = h2->i - h2->k; if (abs(d1 - d2) > 4) return 0;
  d1 = h1->j - h1->m; d2 = h2->j - h2->m; if (abs(d1 - d2) > 4) return 0;
  return 1;
}

static int cmp_sigc(const void *a, const void *b)
{
  const spcoord *x = (const spcoord *)a, *y = (const spcoord *)b;
  return (x->i > y->i) - (x->i < y->i);
}

/* clusters the sampled segments; returns the significant clusters sorted by start, *ret_n of them */
static spcoord *sp_cluster(const spcoord *sp, int n, int nsamples, int *ret_n)
{
  int *assign = (int *)malloc(sizeof(int) * (n + 1)), *a = (int *)malloc(sizeof(int) * (n + 1)), *b = (int *)malloc(sizeof(int) * (n + 1));
  int na = n, nb = 0, nc = 0;
  for (int v = 0; v < n; v++) a[v] = v;
  while (na > 0) {
    int v = a[na - 1]; na--;
    b[nb++] = v;
    while (nb > 0) {
      v = b[nb - 1]; nb--;
      assign[v] = nc;
      for (int i = na - 1; i >= 0; i--)
        if (sp_link(&sp[v], &sp[a[i]])) { int w = a[i]; a[i] = a[na - 1]; na--; b[nb++] = w; }
    }
    nc++;
  }
  spcoord *sig = (spcoord *)malloc(sizeof(spcoord) * (nc + 1));
  int nsig = 0;
  for (int c = 0; c < nc; c++) {
    int idx_of_last = -1, ninc = 0;
    for (int h = 0; h < n; h++) if (assign[h] == c) { if (sp[h].idx != idx_of_last) ninc++; idx_of_last = sp[h].idx; }
    if ((float)ninc / (float)nsamples < 0.25f) continue;
    int imin = INT_MAX, jmin = INT_MAX, kmin = INT_MAX, mmin = INT_MAX, imax = 0, jmax = 0, kmax = 0, mmax = 0;
    for (int h = 0; h < n; h++) if (assign[h] == c) {
      if (sp[h].i < imin) imin = sp[h].i; if (sp[h].i > imax) imax = sp[h].i;
      if (sp[h].j < jmin) jmin = sp[h].j; if (sp[h].j > jmax) jmax = sp[h].j;
      if (sp[h].k < kmin) kmin = sp[h].k; if (sp[h].k > kmax) kmax = sp[h].k;
      if (sp[h].m < mmin) mmin = sp[h].m; if (sp[h].m > mmax) mmax = sp[h].m;
    }
    int len = imax - imin; if (jmax - jmin > len) len = jmax - jmin; if (kmax - kmin > len) len = kmax - kmin; if (mmax - mmin > len) len = mmax - mmin;
    int *epc = (int *)malloc(sizeof(int) * (len + 2));
    int cm, best_i, best_j, best_k, best_m;
    memset(epc, 0, sizeof(int) * (len + 2));
    for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].i - imin]++;
    for (cm = 0, best_i = imin; best_i <= imax; best_i++) { cm += epc[best_i - imin]; if ((float)cm / (float)ninc >= 0.02f) break; }
    memset(epc, 0, sizeof(int) * (len + 2));
    for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].j - jmin]++;
    for (cm = 0, best_j = jmax; best_j >= jmin; best_j--) { cm += epc[best_j - jmin]; if ((float)cm / (float)ninc >= 0.02f) break; }
    memset(epc, 0, sizeof(int) * (len + 2));
    for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].k - kmin]++;
    for (cm = 0, best_k = kmin; best_k <= kmax; best_k++) { cm += epc[best_k - kmin]; if ((float)cm / (float)ninc >= 0.02f) break; }
    memset(epc, 0, sizeof(int) * (len + 2));
    for (int h = 0; h < n; h++) if (assign[h] == c) epc[sp[h].m - mmin]++;
    for (cm = 0, best_m = mmax; best_m >= mmin; best_m--) { cm += epc[best_m - mmin]; if ((float)cm / (float)ninc >= 0.02f) break; }
    free(epc);
    if (best_i > best_j) continue;
    sig[nsig].idx = c; sig[nsig].i = best_i; sig[nsig].j = best_j; sig[nsig].k = best_k; sig[nsig].m = best_m;
    sig[nsig].prob = (float)ninc / (float)nsamples;
    nsig++;
  }
  qsort(sig, nsig, sizeof(spcoord), cmp_sigc);
  free(assign); free(a); free(b);
  *ret_n = nsig;
  return sig;
}

/* null2 from one traced domain segment [z1..z2] (B..E) */
static void null2_by_trace(const orc_profile *p, const trace_t *tr, int z1, int z2, float *cnt, float *null2)
{
  int M = p->M, Ld = 0;
  float *cm = cnt, *ci = cnt + (M + 1);
  float xn = 0, xc = 0, xj = 0;
  memset(cnt, 0, sizeof(float) * 2 * (M + 1));
  for (int z = z1; z <= z2; z++) {
    if (tr->i[z] == 0) continue;
    Ld++;
    if (tr->k[z] > 0) { if (tr->st[z] == ST_M) cm[tr->k[z]] += 1.0f; else ci[tr->k[z]] += 1.0f; }
    else { if (tr->st[z] == ST_N) xn += 1.0f; else if (tr->st[z] == ST_C) xc += 1.0f; else if (tr->st[z] == ST_J) xj += 1.0f; }
  }
  float norm = 1.0f / (float)Ld;
  for (int k = 1; k <= M; k++) { cm[k] *= norm; ci[k] *= norm; }
  xn *= norm; xc *= norm; xj *= norm;
  float xfactor = xn + xc + xj;
  for (int x = 0; x < ORC_K; x++) {
    const float *rp = p->rfv + (size_t)x * (M + 1);
    float sv = 0.0f;
    if (g_order == ORC_ORDER_CANONICAL) {                /* 32 interleaved partial sums (k = l+1, l+33, ...), then the butterfly */
      float part[NB];
      for (int l = 0; l < NB; l++) { float pt = 0.0f; for (int k = l + 1; k <= M; k += NB) { pt += cm[k] * rp[k]; pt += ci[k]; } part[l] = pt; }
      sv = butterfly_sum(part);
    } else {
      for (int k = 1; k <= M; k++) { sv += cm[k] * rp[k]; sv += ci[k]; }
    }
    null2[x] = sv + xfactor;
  }
  for (int x = ORC_K + 1; x <= ORC_KP - 3; x++) {
    float result = 0.0f; int n = 0;
    for (int i = 0; i < ORC_K; i++) if (degen_has(x, i)) { result += null2[i]; n++; }
    null2[x] = result / (float)n;
  }
  null2[20] = 1.0f; null2[27] = 1.0f; null2[28] = 1.0f;
}

#define NSAMPLES 200

/* =====================================================================================
 * Envelope rescoring (unihit, original length model) and the region walk
 * ===================================================================================== */
static int rescore_isolated_domain(ddef_t *dd, const orc_profile *p, const uint8_t *dsq, int Lseq, int i, int j, int null2_is_done)
{
  int Ld = j - i + 1, M = p->M;
  specials sp = make_specials(Lseq, 0);
  fmx *ox1 = fmx_create(M, Ld, 1), *ox2 = fmx_create(M, Ld, 1);
  float envsc, oasc, null2[ORC_KP];
  forward_engine(p, dsq + i - 1, Ld, sp, ox1, &envsc);
  backward_engine(p, dsq + i - 1, Ld, sp, ox1, ox2, NULL);
  if (decoding(sp, ox1, ox2)) { fmx_free(ox1); fmx_free(ox2); return 1; }
  /* null2 needs the posteriors; the OA fill then reuses ox1 */
  if (!null2_is_done) {
    null2_by_expectation(p, ox2, null2);
    for (int pos = i; pos <= j; pos++) dd->n2sc[pos] = (float)log((double)null2[dsq[pos - 1]]);
  }
  oasc = optimal_accuracy(p, sp, ox2, ox1);
  int hf, ht, sf, st;
  if (oa_trace(p, sp, ox2, ox1, &hf, &ht, &sf, &st, NULL)) { fmx_free(ox1); fmx_free(ox2); return 1; }
  float domcorrection = 0.0f;
  for (int pos = i; pos <= j; pos++) domcorrection += dd->n2sc[pos];
  if (dd->ndom == dd->dom_alloc) { dd->dom_alloc = dd->dom_alloc ? dd->dom_alloc * 2 : 4; dd->dcl = (orc_domain *)realloc(dd->dcl, sizeof(orc_domain) * dd->dom_alloc); }
  orc_domain *dom = &dd->dcl[dd->ndom++];
  memset(dom, 0, sizeof(*dom));
  dom->ienv = i; dom->jenv = j; dom->envsc = envsc; dom->domcorrection = domcorrection; dom->oasc = oasc;
  dom->hmmfrom = hf; dom->hmmto = ht; dom->sqfrom = sf + i - 1; dom->sqto = st + i - 1;
  fmx_free(ox1); fmx_free(ox2);
  return 0;
}

static void region_trace_ensemble(ddef_t *dd, const orc_profile *p, const uint8_t *dsq, int Lseq, int ireg, int jreg,
                                  spcoord **ret_sig, int *ret_nc)
{
  int Lr = jreg - ireg + 1, M = p->M;
  specials sp = make_specials(Lseq, 1);
  fmx *fwd = fmx_create(M, Lr, 1);
  forward_engine_seq(p, dsq + ireg - 1, Lr, sp, fwd, NULL);      /* the sampled matrix is evaluated left to right in both orders */
  lcg_t rng; lcg_init(&rng, 42);
  trace_t tr; memset(&tr, 0, sizeof(tr));
  float *cnt = (float *)malloc(sizeof(float) * 2 * (M + 1));
  spcoord *sps = NULL; int nsp = 0, sp_alloc = 0;
  float null2[ORC_KP];
  for (int t = 0; t < NSAMPLES; t++) {
    if (stochastic_trace(&rng, p, sp, fwd, &tr)) continue;
    int pos = 1, z = 0;
    while (z < tr.N) {
      if (tr.st[z] != ST_B) { z++; continue; }
      int z1 = z, z2 = z, sqfrom = 0, sqto = 0, hmmfrom = 0, hmmto = 0;
      for (z2 = z1; z2 < tr.N && tr.st[z2] != ST_E; z2++) {
        if (tr.st[z2] == ST_M) { if (!sqfrom) sqfrom = tr.i[z2]; if (!hmmfrom) hmmfrom = tr.k[z2]; sqto = tr.i[z2]; hmmto = tr.k[z2]; }
        else if (tr.st[z2] == ST_D) { if (!hmmfrom) hmmfrom = tr.k[z2]; hmmto = tr.k[z2]; }
      }
      if (nsp == sp_alloc) { sp_alloc = sp_alloc ? sp_alloc * 2 : 256; sps = (spcoord *)realloc(sps, sizeof(spcoord) * sp_alloc); }
      sps[nsp].idx = t; sps[nsp].i = sqfrom + ireg - 1; sps[nsp].j = sqto + ireg - 1; sps[nsp].k = hmmfrom; sps[nsp].m = hmmto; sps[nsp].prob = 0; nsp++;
      null2_by_trace(p, &tr, z1, z2, cnt, null2);
      for (; pos <= sqfrom; pos++) dd->n2sc[ireg + pos - 1] += 1.0f;
      for (; pos <= sqto;   pos++) dd->n2sc[ireg + pos - 1] += null2[dsq[ireg + pos - 2]];
      z = z2 + 1;
    }
    for (; pos <= Lr; pos++) dd->n2sc[ireg + pos - 1] += 1.0f;
  }
  for (int pos = ireg; pos <= jreg; pos++) dd->n2sc[pos] = (float)log((double)(dd->n2sc[pos] / (float)NSAMPLES));
  *ret_sig = sp_cluster(sps, nsp, NSAMPLES, ret_nc);
  free(sps); free(cnt); free(tr.st); free(tr.k); free(tr.i);
  fmx_free(fwd);
}

static void domaindef_by_posterior_heuristics(ddef_t *dd, const orc_profile *p, const uint8_t *dsq, int L,
                                              const fmx *oxf, const fmx *oxb)
{
  specials spm = make_specials(L, 1);
  domain_decoding(oxf, oxb, spm, dd);
  for (int i = 0; i <= L; i++) dd->n2sc[i] = 0.0f;
  int i = -1, triggered = 0;
  for (int j = 1; j <= L; j++) {
    if (!triggered) {
      if (dd->mocc[j] - (dd->btot[j] - dd->btot[j - 1]) < 0.10f) i = j;
      else if (i == -1) i = j;
      if (dd->mocc[j] >= 0.25f) triggered = 1;
    } else if (dd->mocc[j] - (dd->etot[j] - dd->etot[j - 1]) < 0.10f) {
      dd->nregions++;
      if (is_multidomain_region(dd, i, j)) {
        dd->nclustered++;
        spcoord *sig; int nc, last_j2 = 0;
        region_trace_ensemble(dd, p, dsq, L, i, j, &sig, &nc);
        for (int d = 0; d < nc; d++) {
          if (sig[d].i <= last_j2) dd->noverlaps++;
          dd->nenvelopes++;
          if (rescore_isolated_domain(dd, p, dsq, L, sig[d].i, sig[d].j, 1) == 0) last_j2 = sig[d].j;
        }
        free(sig);
      } else {
        dd->nenvelopes++;
        rescore_isolated_domain(dd, p, dsq, L, i, j, 0);
      }
      i = -1; triggered = 0;
    }
  }
}

/* =====================================================================================
 * The per-target pipeline (A.5 steps 1-6)
 * ===================================================================================== */
int orc_pipeline(const orc_profile *p, const uint8_t *dsq, int L, orc_hit *hit)
{
  orc_filter_result fr;
  const float *ev = p->hmm->evparam;
  if (L == 0) return 0;
  if (!orc_filters(p, dsq, L, &fr)) return 0;
  float nullsc = fr.nullsc, fwdsc = fr.fwd_sc;

  specials spm = make_specials(L, 1);
  fmx *oxf = fmx_create(p->M, L, 0), *oxb = fmx_create(p->M, L, 0);
  forward_engine(p, dsq, L, spm, oxf, NULL);
  backward_engine(p, dsq, L, spm, oxf, oxb, NULL);
  ddef_t dd; memset(&dd, 0, sizeof(dd));
  dd.L = L;
  dd.btot = (float *)calloc((size_t)(L + 1) * 4, sizeof(float));
  dd.etot = dd.btot + (L + 1); dd.mocc = dd.etot + (L + 1); dd.n2sc = dd.mocc + (L + 1);
  domaindef_by_posterior_heuristics(&dd, p, dsq, L, oxf, oxb);
  fmx_free(oxf); fmx_free(oxb);
  if (dd.nregions == 0 || dd.nenvelopes == 0 || dd.ndom == 0) { free(dd.btot); free(dd.dcl); return 0; }

  float seqbias = 0.0f;
  for (int i = 0; i <= L; i++) seqbias += dd.n2sc[i];
  seqbias = flogsum(0.0f, (float)log(OMEGA) + seqbias);
  float pre_score = (fwdsc - nullsc) / (float)LOG2C;
  float seq_score = (fwdsc - (nullsc + seqbias)) / (float)LOG2C;

  float sum_score = 0.0f; int Ld = 0;
  seqbias = 0.0f;
  for (int d = 0; d < dd.ndom; d++)
    if (dd.dcl[d].envsc - dd.dcl[d].domcorrection > 0.0f) {
      sum_score += dd.dcl[d].envsc;
      Ld        += dd.dcl[d].jenv - dd.dcl[d].ienv + 1;
      seqbias   += dd.dcl[d].domcorrection;
    }
  seqbias = flogsum(0.0f, (float)log(OMEGA) + seqbias);
  sum_score = (float)(sum_score + (L - Ld) * log((float)L / (float)(L + 3)));
  float pre2_score = (sum_score - nullsc) / (float)LOG2C;
  sum_score = (sum_score - (nullsc + seqbias)) / (float)LOG2C;
  if (Ld > 0 && sum_score > seq_score) { seq_score = sum_score; pre_score = pre2_score; }

  double lnP = exp_logsurv(seq_score, ev[ORC_FTAU], ev[ORC_FLAMBDA]);
  memset(hit, 0, sizeof(*hit));
  hit->L = L;
  hit->ndom = dd.ndom; hit->nregions = dd.nregions; hit->nclustered = dd.nclustered; hit->nenvelopes = dd.nenvelopes;
  hit->pre_score = pre_score; hit->score = seq_score; hit->sum_score = sum_score; hit->lnP = lnP;
  hit->dcl = dd.dcl;
  for (int d = 0; d < hit->ndom; d++) {
    orc_domain *dom = &hit->dcl[d];
    int Ldd = dom->jenv - dom->ienv + 1;
    dom->bitscore = (float)(dom->envsc + (L - Ldd) * log((float)L / (float)(L + 3)));
    dom->dombias  = flogsum(0.0f, (float)log(OMEGA) + dom->domcorrection);
    dom->bitscore = (dom->bitscore - (nullsc + dom->dombias)) / (float)LOG2C;
    dom->lnP      = exp_logsurv(dom->bitscore, ev[ORC_FTAU], ev[ORC_FLAMBDA]);
  }
  free(dd.btot);
  return 1;
}

/* =====================================================================================
 * hmmalign (checkm/hmmer.py:76-95 -> `hmmalign`): one sequence against one model, unihit local, the whole sequence as the
 * envelope: Forward, Backward, posterior decoding, optimal-accuracy fill and traceback.  state[i-1] for residue i:
 * +k = emitted by match state k, -k = by insert state k, 0 = unaligned flank (N / C).  Returns 0 on success.
 * ===================================================================================== */
/* optimal-accuracy trace of the envelope [i..j] of a sequence of length L (unihit, length model of the full sequence) */
static int align_range(const orc_profile *p, const uint8_t *dsq, int L, int i, int j, int *state, float *ret_oasc)
{
  int Ld = j - i + 1;
  specials sp = make_specials(L, 0);
  fmx *ox1 = fmx_create(p->M, Ld, 1), *ox2 = fmx_create(p->M, Ld, 1);
  float envsc;
  forward_engine(p, dsq + i - 1, Ld, sp, ox1, &envsc);
  backward_engine(p, dsq + i - 1, Ld, sp, ox1, ox2, NULL);
  int rc = 1;
  if (!decoding(sp, ox1, ox2)) {
    float oasc = optimal_accuracy(p, sp, ox2, ox1);
    int hf, ht, sf, st;
    int *trace = (int *)calloc((size_t)Ld + 1, sizeof(int));
    if (!oa_trace(p, sp, ox2, ox1, &hf, &ht, &sf, &st, trace)) {
      for (int r = 1; r <= Ld; r++) state[i - 1 + r - 1] = trace[r];
      if (ret_oasc) *ret_oasc = oasc;
      rc = 0;
    }
    free(trace);
  }
  fmx_free(ox1); fmx_free(ox2);
  return rc;
}

int orc_pipeline(const orc_profile *p, const uint8_t *dsq, int L, orc_hit *hit);

int orc_align(const orc_profile *p, const uint8_t *dsq, int L, int *state, float *ret_oasc)
{
  for (int i = 0; i < L; i++) state[i] = 0;
  if (ret_oasc) *ret_oasc = 0.0f;
  if (L == 0) return 0;
  if (align_range(p, dsq, L, 1, L, state, ret_oasc) == 0) return 0;
  /* One unihit envelope cannot hold a sequence with a second strong copy of the domain (scaled fp32: the Backward pass overflows
   * where the Forward pass has underflowed).  Such a sequence is aligned over the envelope of its best-scoring domain as the
   * search pipeline defines it; the rest of it is flank. */
  for (int i = 0; i < L; i++) state[i] = 0;
  orc_hit hit;
  if (!orc_pipeline(p, dsq, L, &hit)) return 1;
  int best = -1;
  for (int d = 0; d < hit.ndom; d++) if (best < 0 || hit.dcl[d].bitscore > hit.dcl[best].bitscore) best = d;
  int rc = 1;
  if (best >= 0) rc = align_range(p, dsq, L, hit.dcl[best].ienv, hit.dcl[best].jenv, state, ret_oasc);
  free(hit.dcl);
  return rc;
}

/* =====================================================================================
 * Search driver: every model against every sequence; Z = nseq; thresholds; per-model sort.
 * ===================================================================================== */
typedef struct {
  orc_profile **profs; int nmodels;
  const uint8_t *residues; const int64_t *offsets; int nseq;
  int next; pthread_mutex_t mu;
  orc_hit *hits; int nhits, alloc;
} search_ctx;

static void *search_worker(void *arg)
{
  search_ctx *c = (search_ctx *)arg;
  const int chunk = 16;
  long total = (long)c->nmodels * c->nseq;
  while (1) {
    pthread_mutex_lock(&c->mu);
    long begin = c->next; c->next += chunk;
    pthread_mutex_unlock(&c->mu);
    if (begin >= total) break;
    long end = begin + chunk; if (end > total) end = total;
    for (long w = begin; w < end; w++) {
      int m = (int)(w / c->nseq), s = (int)(w % c->nseq);
      int L = (int)(c->offsets[s + 1] - c->offsets[s]);
      orc_hit h;
      if (orc_pipeline(c->profs[m], c->residues + c->offsets[s], L, &h)) {
        h.seqidx = s; h.model = m;
        pthread_mutex_lock(&c->mu);
        if (c->nhits == c->alloc) { c->alloc = c->alloc ? c->alloc * 2 : 256; c->hits = (orc_hit *)realloc(c->hits, sizeof(orc_hit) * c->alloc); }
        c->hits[c->nhits++] = h;
        pthread_mutex_unlock(&c->mu);
      }
    }
  }
  return NULL;
}

static int cmp_hits(const void *a, const void *b)
{
  const orc_hit *x = (const orc_hit *)a, *y = (const orc_hit *)b;
  if (x->model != y->model) return (x->model > y->model) - (x->model < y->model);
  if (x->lnP != y->lnP) return (x->lnP > y->lnP) - (x->lnP < y->lnP);      /* sortkey = -lnP, descending */
  return (x->seqidx > y->seqidx) - (x->seqidx < y->seqidx);
}

orc_results *orc_search(orc_profile **profs, int nmodels, const uint8_t *residues, const int64_t *offsets, int nseq,
                        double Ecut, double domEcut, int nthreads)
{
  search_ctx c; memset(&c, 0, sizeof(c));
  c.profs = profs; c.nmodels = nmodels; c.residues = residues; c.offsets = offsets; c.nseq = nseq;
  pthread_mutex_init(&c.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, search_worker, &c);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&c.mu);

  orc_results *r = (orc_results *)calloc(1, sizeof(orc_results));
  r->hits = c.hits; r->nhits = c.nhits; r->Z = (double)nseq; r->nmodels = nmodels;
  r->domZ = (double *)calloc(nmodels > 0 ? nmodels : 1, sizeof(double));
  qsort(r->hits, r->nhits, sizeof(orc_hit), cmp_hits);
  for (int h = 0; h < r->nhits; h++) {
    orc_hit *hit = &r->hits[h];
    if (exp(hit->lnP) * r->Z <= Ecut) { hit->is_reported = 1; r->domZ[hit->model] += 1.0; }
  }
  for (int h = 0; h < r->nhits; h++) {
    orc_hit *hit = &r->hits[h];
    if (!hit->is_reported) continue;
    for (int d = 0; d < hit->ndom; d++)
      if (exp(hit->dcl[d].lnP) * r->domZ[hit->model] <= domEcut) { hit->dcl[d].is_reported = 1; hit->nreported++; }
  }
  return r;
}

void orc_results_free(orc_results *r)
{
  if (!r) return;
  for (int h = 0; h < r->nhits; h++) free(r->hits[h].dcl);
  free(r->hits); free(r->domZ); free(r);
}
int orc_results_nhits(const orc_results *r) { return r->nhits; }
const orc_hit *orc_results_hit(const orc_results *r, int i) { return &r->hits[i]; }
const orc_domain *orc_hit_domain(const orc_hit *h, int d) { return &h->dcl[d]; }

/* Raw filter scores (nats) of every (model, sequence) pair, for the calibration tests: msv / vit / fwd are
 * [nmodels*nseq] or NULL.  No thresholds are applied; an MSV/Viterbi overflow is reported as +inf. */
typedef struct {
  orc_profile **profs; int nmodels;
  const uint8_t *residues; const int64_t *offsets; int nseq;
  float *msv, *vit, *fwd;
  long next; pthread_mutex_t mu;
} stage_ctx;

static void *stage_worker(void *arg)
{
  stage_ctx *c = (stage_ctx *)arg;
  const long chunk = 32, total = (long)c->nmodels * c->nseq;
  while (1) {
    pthread_mutex_lock(&c->mu);
    long begin = c->next; c->next += chunk;
    pthread_mutex_unlock(&c->mu);
    if (begin >= total) break;
    long end = begin + chunk; if (end > total) end = total;
    for (long w = begin; w < end; w++) {
      int m = (int)(w / c->nseq), s = (int)(w % c->nseq);
      int L = (int)(c->offsets[s + 1] - c->offsets[s]);
      const uint8_t *dsq = c->residues + c->offsets[s];
      int xj;
      if (c->msv) orc_msv(c->profs[m], dsq, L, &c->msv[w], &xj);
      if (c->vit) orc_vitfilter(c->profs[m], dsq, L, &c->vit[w]);
      if (c->fwd) orc_forward_parser(c->profs[m], dsq, L, &c->fwd[w]);
    }
  }
  return NULL;
}

int orc_stage_scores(orc_profile **profs, int nmodels, const uint8_t *residues, const int64_t *offsets, int nseq,
                     float *msv, float *vit, float *fwd, int nthreads)
{
  stage_ctx c; memset(&c, 0, sizeof(c));
  c.profs = profs; c.nmodels = nmodels; c.residues = residues; c.offsets = offsets; c.nseq = nseq;
  c.msv = msv; c.vit = vit; c.fwd = fwd;
  pthread_mutex_init(&c.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, stage_worker, &c);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&c.mu);
  return 0;
}

/* domtblout: one row per reported domain of each reported target */
int orc_write_domtblout(const orc_results *r, orc_profile **profs, const char **seqnames, const char **seqdescs, const char *path)
{
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  int tnamew = 20, qnamew = 20, qaccw = 10, taccw = 10;
  for (int h = 0; h < r->nhits; h++) if (r->hits[h].is_reported) {
    int n = (int)strlen(seqnames[r->hits[h].seqidx]); if (n > tnamew) tnamew = n;
    const orc_hmm *hm = profs[r->hits[h].model]->hmm;
    n = (int)strlen(hm->name); if (n > qnamew) qnamew = n;
    n = (int)strlen(hm->acc);  if (n > qaccw)  qaccw = n;
  }
  fprintf(fp, "#%*s %22s %40s %11s %11s %11s\n", tnamew + qnamew - 1 + 15 + taccw + qaccw, "", "--- full sequence ---",
          "-------------- this domain -------------", "hmm coord", "ali coord", "env coord");
  fprintf(fp, "#%-*s %-*s %5s %-*s %-*s %5s %9s %6s %5s %3s %3s %9s %9s %6s %5s %5s %5s %5s %5s %5s %5s %4s %s\n",
          tnamew - 1, " target name", taccw, "accession", "tlen", qnamew, "query name", qaccw, "accession", "qlen",
          "E-value", "score", "bias", "#", "of", "c-Evalue", "i-Evalue", "score", "bias", "from", "to", "from", "to", "from", "to", "acc", "description of target");
  fprintf(fp, "#------------------- ---------- ----- -------------------- ---------- ----- --------- ------ ----- --- --- --------- --------- ------ ----- ----- ----- ----- ----- ----- ----- ---- ---------------------\n");
  for (int h = 0; h < r->nhits; h++) {
    const orc_hit *hit = &r->hits[h];
    if (!hit->is_reported) continue;
    const orc_hmm *hm = profs[hit->model]->hmm;
    int nd = 0;
    for (int d = 0; d < hit->ndom; d++) {
      const orc_domain *dom = &hit->dcl[d];
      if (!dom->is_reported) continue;
      nd++;
      fprintf(fp, "%-*s %-*s %5d %-*s %-*s %5d %9.2g %6.1f %5.1f %3d %3d %9.2g %9.2g %6.1f %5.1f %5d %5d %5d %5d %5d %5d %4.2f %s\n",
              tnamew, seqnames[hit->seqidx], taccw, "-", hit->L,
              qnamew, hm->name, qaccw, (hm->acc[0] ? hm->acc : "-"), hm->M,
              exp(hit->lnP) * r->Z, hit->score, hit->pre_score - hit->score,
              nd, hit->nreported,
              exp(dom->lnP) * r->domZ[hit->model], exp(dom->lnP) * r->Z,
              dom->bitscore, dom->dombias * (float)(1.0 / LOG2C),
              dom->hmmfrom, dom->hmmto, dom->sqfrom, dom->sqto, dom->ienv, dom->jenv,
              (dom->oasc / (1.0 + fabs((float)(dom->jenv - dom->ienv)))),
              (seqdescs && seqdescs[hit->seqidx] && seqdescs[hit->seqidx][0]) ? seqdescs[hit->seqidx] : "-");
    }
  }
  fclose(fp);
  return 0;
}
