/*
 * hmmer_oracle.h -- CPU restatement of HMMER3 `hmmsearch` (the arithmetic CheckM reaches through
 * os.system at checkm/hmmer.py:70-71 with the flags at checkm/markerGeneFinder.py:141).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under checkm_b200/ may include, link or call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY: HMMER's sources are not under /root/reference and no hmmsearch binary exists in the build container
 * (SURVEY.md section 8c); the algorithm is restated from the published HMMER 3.1b2 design (Eddy 2011, "Accelerated
 * profile HMM searches"; Eddy 2008) -- see SURVEY.md Appendix A.
 *   PINNED against numbers the real HMMER 3.1b2 produced: the HMM file parser (fixture custom_marker_sets/cpr_43_markers.hmm),
 *     and the MSV, ViterbiFilter and ForwardParser SCORES -- replaying hmmbuild's calibration (same generator, same 3 x 200
 *     sequences) reproduces the STATS LOCAL lines of all 43 fixture models (tests/test_oracle_cpu.py).
 *   PARITY UNPINNED: the domain-definition arithmetic (regions, trace ensemble, null2, optimal accuracy, reported scores).
 *     tools/validate_against_hmmer.sh diffs domtblout against a real hmmsearch wherever one is installed.
 */
#ifndef HMMER_ORACLE_H
#define HMMER_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_K   20
#define ORC_KP  29

enum { ORC_MM = 0, ORC_MI, ORC_MD, ORC_IM, ORC_II, ORC_DM, ORC_DD, ORC_NT };
enum { ORC_MMU = 0, ORC_MLAMBDA, ORC_VMU, ORC_VLAMBDA, ORC_FTAU, ORC_FLAMBDA };

typedef struct {
  char   name[128];
  char   acc[64];
  char   desc[256];
  int    M;
  float *mat;      /* (M+1)*20 match emission probabilities, node 0 unused          */
  float *ins;      /* (M+1)*20 insert emission probabilities                       */
  float *t;        /* (M+1)*7  transition probabilities, file order MM MI MD IM II DM DD */
  float  compo[ORC_K];
  int    has_compo;
  float  ga[2], tc[2], nc[2];
  int    has_ga, has_tc, has_nc;
  float  evparam[6];
  int    has_stats;
} orc_hmm;

typedef struct {
  int      M;
  const orc_hmm *hmm;
  /* generic log-odds profile (local, length-independent part) */
  float   *tsc;        /* (M+1)*7 log transitions; rows 0 and M are -inf              */
  float   *bm;         /* (M+1)   log local entry B->M_k                              */
  float   *msc;        /* ORC_KP*(M+1) match log-odds, [x*(M+1)+k]                    */
  /* MSV filter (8-bit costs) */
  uint8_t *rbv;        /* ORC_KP*(M+1)                                                */
  uint8_t  tbm_b, tec_b, base_b, bias_b;
  float    scale_b;
  /* Viterbi filter (16-bit scores) */
  int16_t *rwv;        /* ORC_KP*(M+1)                                                */
  int16_t *twv;        /* (M+1)*8: [k*8 + {BM,MM,IM,DM,MD,MI,II,DD}]                  */
  int16_t  base_w, xw_e_loop, xw_e_move, ddbound_w;
  float    scale_w;
  /* Forward/Backward (odds ratios) */
  float   *rfv;        /* ORC_KP*(M+1)                                                */
  float   *tfv;        /* (M+1)*8 same order as twv                                   */
  void    *striped;    /* SSE2 striped copies (simd_filters.c), NULL until orc_profile_enable_simd() */
} orc_profile;

/* one reported (or unreported) domain of a hit */
typedef struct {
  int    ienv, jenv;
  int    hmmfrom, hmmto, sqfrom, sqto;
  float  envsc;          /* nats */
  float  domcorrection;  /* nats */
  float  dombias;        /* nats */
  float  oasc;
  float  bitscore;       /* bits */
  double lnP;
  int    is_reported;
} orc_domain;

typedef struct {
  int    seqidx;
  int    model;
  int    L;
  float  pre_score, score, sum_score;   /* bits */
  double lnP;
  int    ndom, nreported, nregions, nclustered, nenvelopes;
  int    is_reported;
  orc_domain *dcl;
} orc_hit;

/* filter outcomes for one (profile, sequence) pair; used by parity tests */
typedef struct {
  int    msv_xJ;          /* final xJ byte, or 256 on overflow                        */
  float  msv_sc;          /* nats (INFINITY on overflow)                              */
  float  nullsc, filtersc;
  float  vit_sc, fwd_sc;  /* nats; NAN when the stage was not reached                 */
  int    passed_msv, passed_bias, passed_vit, passed_fwd;
} orc_filter_result;

/* ---- alphabet ---- */
int  orc_digitize(const char *seq, int n, uint8_t *dsq);   /* returns #invalid symbols */

/* ---- HMM file ---- */
int  orc_hmmfile_read(const char *path, orc_hmm **ret_hmms, int *ret_n);
void orc_hmms_free(orc_hmm *hmms, int n);
orc_hmm *orc_hmm_at(orc_hmm *hmms, int i);

/* ---- profile ---- */
orc_profile *orc_profile_create(const orc_hmm *hmm);
void         orc_profile_free(orc_profile *p);

/* ---- SSE2 striped MSV / Viterbi filters for the CPU baseline of bench.py (simd_filters.c): same scores as the scalar
 * functions below; once enabled on a profile, orc_filters / orc_pipeline / orc_search use them ---- */
void  orc_profile_enable_simd(orc_profile *p);
void *orc_striped_create(const orc_profile *p);
void  orc_striped_free(void *h);
int   orc_msv_simd(const orc_profile *p, const void *h, const uint8_t *dsq, int L, float *ret_sc, int *ret_xJ);
int   orc_vitfilter_simd(const orc_profile *p, const void *h, const uint8_t *dsq, int L, float *ret_sc);

/* ---- evaluation order of the fp32 row sums (see hmmer_oracle.c): sequential (textbook) or canonical (the engine's
 * specified blocked order, the default) ---- */
enum { ORC_ORDER_SEQUENTIAL = 0, ORC_ORDER_CANONICAL = 1 };
void orc_set_order(int order);
int  orc_get_order(void);
int  orc_block_width(int M);

/* ---- individual stages (dsq is 0-based, length L) ---- */
float orc_null1(int L);
int   orc_msv(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc, int *ret_xJ);
int   orc_ssv_xe(const orc_profile *p, const uint8_t *dsq, int L);   /* max xE with J disabled */
float orc_biasfilter(const orc_profile *p, const uint8_t *dsq, int L);
int   orc_vitfilter(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc);
int   orc_forward_parser(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc);
int   orc_backward_parser(const orc_profile *p, const uint8_t *dsq, int L, float *ret_sc);
int   orc_filters(const orc_profile *p, const uint8_t *dsq, int L, orc_filter_result *r);

/* ---- full per-target pipeline; returns 1 and fills *hit if the target reaches the hit list ---- */
int   orc_pipeline(const orc_profile *p, const uint8_t *dsq, int L, orc_hit *hit);

/* ---- hmmalign: optimal-accuracy alignment of one sequence to one model (unihit local); state[i-1] = +k match / -k insert / 0 flank ---- */
int   orc_align(const orc_profile *p, const uint8_t *dsq, int L, int *state, float *ret_oasc);

/* ---- search: all models x all sequences, thresholds E<=Ecut, domE<=domEcut, Z=nseq ---- */
typedef struct {
  int      nhits;
  orc_hit *hits;       /* sorted per model by lnP, thresholded flags set                */
  double   Z;
  double  *domZ;       /* per model                                                    */
  int      nmodels;
} orc_results;

orc_results *orc_search(orc_profile **profs, int nmodels,
                        const uint8_t *residues, const int64_t *offsets, int nseq,
                        double Ecut, double domEcut, int nthreads);
void orc_results_free(orc_results *r);
int  orc_results_nhits(const orc_results *r);
const orc_hit *orc_results_hit(const orc_results *r, int i);
const orc_domain *orc_hit_domain(const orc_hit *h, int d);

/* raw filter scores (nats) of every (model, sequence) pair -- calibration tests; each output is [nmodels*nseq] or NULL */
int orc_stage_scores(orc_profile **profs, int nmodels, const uint8_t *residues, const int64_t *offsets, int nseq,
                     float *msv, float *vit, float *fwd, int nthreads);

/* domtblout text for one search (names/descriptions supplied by the caller) */
int orc_write_domtblout(const orc_results *r, orc_profile **profs,
                        const char **seqnames, const char **seqdescs, const char *path);

#ifdef __cplusplus
}
#endif
#endif
