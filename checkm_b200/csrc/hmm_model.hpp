// hmm_model.hpp -- host-side model database: HMMER3/f reader (header + body) and the three
// score systems the device kernels consume (8-bit MSV costs, 16-bit Viterbi scores, fp32 odds ratios).
// Stands where hmmsearch's own HMM reading and profile configuration stood behind
// checkm/hmmer.py:70-71; the header fields are the ones checkm/hmmerModelParser.py:46-83 scrapes.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ckm {

constexpr int K   = 20;   // canonical residues
constexpr int KP  = 29;   // "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"
constexpr int KPAD = 30;  // + code 29: row padding, scores -inf everywhere
constexpr int CODE_PAD = 29;

// transition slots of the device tables, node-k centric:
//   entering M_k : BM (B->M_k), MM (M_{k-1}->M_k), IM (I_{k-1}->M_k), DM (D_{k-1}->M_k)
//   leaving node k: MD (M_k->D_{k+1}), MI (M_k->I_k), II (I_k->I_k), DD (D_k->D_{k+1})
enum { T_BM = 0, T_MM, T_IM, T_DM, T_MD, T_MI, T_II, T_DD, T_N };
// file order
enum { H_MM = 0, H_MI, H_MD, H_IM, H_II, H_DM, H_DD, H_N };

struct Model {
  std::string name, acc, desc;
  std::string text;             // verbatim HMMER3/f record, for ckm_models_write (hmmfetch replacement)
  int M = 0;
  bool has_ga = false, has_tc = false, has_nc = false, has_compo = false;
  float ga[2] = {0, 0}, tc[2] = {0, 0}, nc[2] = {0, 0};
  double ga_d[2] = {0, 0}, tc_d[2] = {0, 0}, nc_d[2] = {0, 0};   // as Python's float() reads them (hmmerModelParser.py:76)
  float evparam[6] = {0, 0, 0, 0, 0, 0};
  float compo[K];
  std::vector<float> mat, ins, t;      // probabilities: (M+1)*20, (M+1)*20, (M+1)*7

  // ---- configured profile (multihit local; length-dependent specials are applied per sequence) ----
  std::vector<float>   msc;            // KP*(M+1) match log-odds
  std::vector<float>   bm;             // (M+1) log B->M_k
  std::vector<float>   tsc;            // (M+1)*7 log transitions (file order), rows 0 and M = -inf
  // MSV
  std::vector<uint8_t> rbv;            // KP*(M+1) biased costs
  uint8_t tbm_b = 0, tec_b = 0, base_b = 190, bias_b = 0;
  float   scale_b = 0;
  // Viterbi filter
  std::vector<int16_t> rwv;            // KP*(M+1)
  std::vector<int16_t> twv;            // (M+1)*8
  int16_t base_w = 12000, xw_e_loop = 0, xw_e_move = 0;
  int32_t ddbound_w = -32768;
  float   scale_w = 0;
  // Forward / Backward
  std::vector<float>   rfv;            // KP*(M+1) match odds ratios
  std::vector<float>   tfv;            // (M+1)*8 transition probabilities
  // bias filter emission odds, [x][2]
  float   bias_eo[KP][2];
};

// Reads every model of a HMMER3/f file.  Throws std::runtime_error with a message on malformed input.
std::vector<Model> read_hmm_file(const std::string &path);
// Fills the configured-profile members of m.
void configure_profile(Model &m);

extern const float BG_F[K];
bool degen_has(int x, int r);
int  digitize_char(unsigned char c);     // -1 for symbols outside the alphabet

}  // namespace ckm
