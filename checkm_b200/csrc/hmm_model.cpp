// hmm_model.cpp -- HMMER3/f reader and profile configuration (host side of libckm.so).
#include "hmm_model.hpp"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace ckm {

const float BG_F[K] = {0.0787945f, 0.0151600f, 0.0535222f, 0.0668298f, 0.0397062f, 0.0695071f, 0.0229198f,
                       0.0590092f, 0.0594422f, 0.0963728f, 0.0237718f, 0.0414386f, 0.0482904f, 0.0395639f,
                       0.0540978f, 0.0683364f, 0.0540687f, 0.0673417f, 0.0114135f, 0.0304133f};

static const char *kAlphabet = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";
static const double kLog2 = 0.69314718055994529;

bool degen_has(int x, int r) {
  switch (x) {
    case 21: return r == 11 || r == 2;   // B = N|D
    case 22: return r == 7 || r == 9;    // J = I|L
    case 23: return r == 13 || r == 3;   // Z = Q|E
    case 24: return r == 8;              // O -> K
    case 25: return r == 1;              // U -> C
    case 26: return true;                // X
    default: return x == r;
  }
}

int digitize_char(unsigned char c) {
  static int8_t map[256];
  static bool init = false;
  if (!init) {
    std::memset(map, -1, sizeof(map));
    for (int i = 0; i < KP; ++i) {
      unsigned char ch = (unsigned char)kAlphabet[i];
      map[ch] = (int8_t)i;
      if (ch >= 'A' && ch <= 'Z') map[ch + 32] = (int8_t)i;
    }
    map[(unsigned char)'.'] = 20;
    init = true;
  }
  return map[c];
}

static inline float neglog_to_prob(const std::string &tok) {
  if (!tok.empty() && tok[0] == '*') return 0.0f;
  return expf(-1.0f * (float)std::atof(tok.c_str()));
}

static std::vector<std::string> split_ws(const std::string &s) {
  std::vector<std::string> out;
  std::istringstream is(s);
  std::string w;
  while (is >> w) out.push_back(w);
  return out;
}

static std::string rest_after_tag(const std::string &line) {
  size_t p = line.find_first_of(" \t");
  if (p == std::string::npos) return "";
  p = line.find_first_not_of(" \t", p);
  if (p == std::string::npos) return "";
  size_t e = line.find_last_not_of(" \t\r\n");
  return line.substr(p, e - p + 1);
}

std::vector<Model> read_hmm_file(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open HMM file " + path);
  std::vector<Model> models;
  std::string line;
  auto need = [&](std::string &l, Model &m) {
    if (!std::getline(in, l)) throw std::runtime_error("unexpected end of HMM file " + path);
    m.text += l;
    m.text += '\n';
  };
  while (std::getline(in, line)) {
    if (line.compare(0, 6, "HMMER3") != 0) continue;
    Model m;
    m.text = line + "\n";
    bool in_body = false;
    while (!in_body) {
      need(line, m);
      std::vector<std::string> w = split_ws(line);
      if (w.empty()) continue;
      const std::string &tag = w[0];
      if (tag == "NAME") m.name = rest_after_tag(line);
      else if (tag == "ACC") m.acc = rest_after_tag(line);
      else if (tag == "DESC") m.desc = rest_after_tag(line);
      else if (tag == "LENG") m.M = std::atoi(w.at(1).c_str());
      else if (tag == "GA" && w.size() >= 3) { for (int z = 0; z < 2; ++z) { m.ga_d[z] = std::atof(w[1 + z].c_str()); m.ga[z] = (float)m.ga_d[z]; } m.has_ga = true; }
      else if (tag == "TC" && w.size() >= 3) { for (int z = 0; z < 2; ++z) { m.tc_d[z] = std::atof(w[1 + z].c_str()); m.tc[z] = (float)m.tc_d[z]; } m.has_tc = true; }
      else if (tag == "NC" && w.size() >= 3) { for (int z = 0; z < 2; ++z) { m.nc_d[z] = std::atof(w[1 + z].c_str()); m.nc[z] = (float)m.nc_d[z]; } m.has_nc = true; }
      else if (tag == "STATS" && w.size() >= 5) {
        float a = (float)std::atof(w[3].c_str()), b = (float)std::atof(w[4].c_str());
        if (w[2] == "MSV") { m.evparam[0] = a; m.evparam[1] = b; }
        else if (w[2] == "VITERBI") { m.evparam[2] = a; m.evparam[3] = b; }
        else if (w[2] == "FORWARD") { m.evparam[4] = a; m.evparam[5] = b; }
      } else if (tag == "HMM") in_body = true;
    }
    if (m.M <= 0) throw std::runtime_error("model " + m.name + " has no LENG in " + path);
    const int M = m.M;
    m.mat.assign((size_t)(M + 1) * K, 0.0f);
    m.ins.assign((size_t)(M + 1) * K, 0.0f);
    m.t.assign((size_t)(M + 1) * H_N, 0.0f);
    need(line, m);  // transition labels
    for (int k = 0; k <= M; ++k) {
      need(line, m);
      std::vector<std::string> w = split_ws(line);
      if (k == 0) {
        if (!w.empty() && w[0] == "COMPO") {
          if (w.size() < 21) throw std::runtime_error("short COMPO line in " + m.name);
          for (int x = 0; x < K; ++x) m.compo[x] = neglog_to_prob(w[1 + x]);
          m.has_compo = true;
          need(line, m);
          w = split_ws(line);
        }
        if (w.size() < (size_t)K) throw std::runtime_error("short insert line (node 0) in " + m.name);
        for (int x = 0; x < K; ++x) m.ins[x] = neglog_to_prob(w[x]);
      } else {
        if (w.size() < (size_t)K + 1 || std::atoi(w[0].c_str()) != k)
          throw std::runtime_error("bad match line for node " + std::to_string(k) + " of " + m.name);
        for (int x = 0; x < K; ++x) m.mat[(size_t)k * K + x] = neglog_to_prob(w[1 + x]);
        need(line, m);
        w = split_ws(line);
        if (w.size() < (size_t)K) throw std::runtime_error("short insert line in " + m.name);
        for (int x = 0; x < K; ++x) m.ins[(size_t)k * K + x] = neglog_to_prob(w[x]);
      }
      need(line, m);
      w = split_ws(line);
      if (w.size() < (size_t)H_N) throw std::runtime_error("short transition line in " + m.name);
      for (int z = 0; z < H_N; ++z) m.t[(size_t)k * H_N + z] = neglog_to_prob(w[z]);
    }
    need(line, m);  // "//"
    if (!m.has_compo) {
      // no COMPO line: fall back to the average match emission, which is what the composition is
      for (int x = 0; x < K; ++x) {
        float s = 0.0f;
        for (int k = 1; k <= M; ++k) s += m.mat[(size_t)k * K + x];
        m.compo[x] = s / (float)M;
      }
    }
    configure_profile(m);
    models.push_back(std::move(m));
  }
  if (models.empty()) throw std::runtime_error("no HMMER3 models found in " + path);
  return models;
}

// ---- limited-precision conversions --------------------------------------------------------------
static inline uint8_t unbiased_byteify(float scale_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.0f) ? 255 : (uint8_t)sc;
}
static inline uint8_t biased_byteify(float scale_b, uint8_t bias_b, float sc) {
  sc = -1.0f * roundf(scale_b * sc);
  return (sc > 255.0f - (float)bias_b) ? 255 : (uint8_t)((uint8_t)sc + bias_b);
}
static inline int16_t wordify(float scale_w, float sc) {
  sc = roundf(scale_w * sc);
  if (sc >= 32767.0f) return 32767;
  if (sc <= -32768.0f) return -32768;
  return (int16_t)sc;
}

void configure_profile(Model &m) {
  const int M = m.M;
  const float NINF = -std::numeric_limits<float>::infinity();
  const size_t W = (size_t)M + 1;
  m.msc.assign((size_t)KP * W, NINF);
  m.bm.assign(W, NINF);
  m.tsc.assign(W * H_N, NINF);

  // local entry distribution from match-state occupancy
  {
    std::vector<float> occ(W, 0.0f);
    occ[1] = m.t[0 * H_N + H_MI] + m.t[0 * H_N + H_MM];
    for (int k = 2; k <= M; ++k)
      occ[k] = occ[k - 1] * (m.t[(size_t)(k - 1) * H_N + H_MM] + m.t[(size_t)(k - 1) * H_N + H_MI]) +
               (1.0f - occ[k - 1]) * m.t[(size_t)(k - 1) * H_N + H_DM];
    float Z = 0.0f;
    for (int k = 1; k <= M; ++k) Z += occ[k] * (float)(M - k + 1);
    for (int k = 1; k <= M; ++k) m.bm[k] = (float)std::log(occ[k] / Z);
  }
  for (int k = 1; k < M; ++k)
    for (int z = 0; z < H_N; ++z) m.tsc[(size_t)k * H_N + z] = (float)std::log(m.t[(size_t)k * H_N + z]);

  for (int k = 1; k <= M; ++k) {
    float sc[KP];
    for (int x = 0; x < K; ++x) sc[x] = (float)std::log((double)m.mat[(size_t)k * K + x] / BG_F[x]);
    sc[20] = NINF; sc[27] = NINF; sc[28] = NINF;
    for (int x = K + 1; x <= KP - 3; ++x) {
      float result = 0.0f, denom = 0.0f;
      for (int i = 0; i < K; ++i)
        if (degen_has(x, i)) { result += sc[i] * BG_F[i]; denom += BG_F[i]; }
      sc[x] = result / denom;
    }
    for (int x = 0; x < KP; ++x) m.msc[(size_t)x * W + k] = sc[x];
  }

  // ---- MSV bytes ----
  {
    float mx = (M >= 2) ? 0.0f : NINF;
    for (int x = 0; x < K; ++x)
      for (int k = 1; k <= M; ++k) mx = std::fmax(mx, m.msc[(size_t)x * W + k]);
    m.scale_b = (float)(3.0 / kLog2);
    m.base_b = 190;
    m.bias_b = unbiased_byteify(m.scale_b, -1.0f * mx);
    m.rbv.assign((size_t)KP * W, 255);
    for (int x = 0; x < KP; ++x)
      for (int k = 1; k <= M; ++k) m.rbv[(size_t)x * W + k] = biased_byteify(m.scale_b, m.bias_b, m.msc[(size_t)x * W + k]);
    m.tbm_b = unbiased_byteify(m.scale_b, logf(2.0f / ((float)M * (float)(M + 1))));
    m.tec_b = unbiased_byteify(m.scale_b, logf(0.5f));
  }

  // ---- Viterbi words ----
  {
    m.scale_w = (float)(500.0 / kLog2);
    m.base_w = 12000;
    m.rwv.assign((size_t)KP * W, -32768);
    for (int x = 0; x < KP; ++x)
      for (int k = 1; k <= M; ++k) m.rwv[(size_t)x * W + k] = wordify(m.scale_w, m.msc[(size_t)x * W + k]);
    m.twv.assign(W * T_N, -32768);
    for (int k = 1; k <= M; ++k) {
      int16_t *tw = &m.twv[(size_t)k * T_N];
      tw[T_BM] = wordify(m.scale_w, m.bm[k]);
      tw[T_MM] = wordify(m.scale_w, m.tsc[(size_t)(k - 1) * H_N + H_MM]);
      tw[T_IM] = wordify(m.scale_w, m.tsc[(size_t)(k - 1) * H_N + H_IM]);
      tw[T_DM] = wordify(m.scale_w, m.tsc[(size_t)(k - 1) * H_N + H_DM]);
      if (k < M) {
        tw[T_MD] = wordify(m.scale_w, m.tsc[(size_t)k * H_N + H_MD]);
        tw[T_MI] = wordify(m.scale_w, m.tsc[(size_t)k * H_N + H_MI]);
        tw[T_II] = wordify(m.scale_w, m.tsc[(size_t)k * H_N + H_II]);
        tw[T_DD] = wordify(m.scale_w, m.tsc[(size_t)k * H_N + H_DD]);
      }
      for (int z = T_BM; z <= T_MI; ++z) if (tw[z] > 0) tw[z] = 0;
      if (tw[T_II] > -1) tw[T_II] = -1;
    }
    m.xw_e_loop = wordify(m.scale_w, (float)-kLog2);
    m.xw_e_move = wordify(m.scale_w, (float)-kLog2);
    // lazy-F bound: the best a D->D->M detour can do relative to entering the same match state from B
    m.ddbound_w = -32768;
    for (int k = 2; k < M - 1; ++k) {
      int dd = (int)wordify(m.scale_w, m.tsc[(size_t)k * H_N + H_DD]);
      dd += (int)wordify(m.scale_w, m.tsc[(size_t)(k + 1) * H_N + H_DM]);
      dd -= (int)wordify(m.scale_w, m.bm[k + 2]);
      if (dd > m.ddbound_w) m.ddbound_w = dd;
    }
  }

  // ---- Forward/Backward odds ----
  m.rfv.assign((size_t)KP * W, 0.0f);
  for (int x = 0; x < KP; ++x)
    for (int k = 0; k <= M; ++k) m.rfv[(size_t)x * W + k] = expf(m.msc[(size_t)x * W + k]);
  m.tfv.assign(W * T_N, 0.0f);
  for (int k = 0; k <= M; ++k) {
    float *tf = &m.tfv[(size_t)k * T_N];
    tf[T_BM] = (k >= 1) ? expf(m.bm[k]) : 0.0f;
    tf[T_MM] = (k >= 1) ? expf(m.tsc[(size_t)(k - 1) * H_N + H_MM]) : 0.0f;
    tf[T_IM] = (k >= 1) ? expf(m.tsc[(size_t)(k - 1) * H_N + H_IM]) : 0.0f;
    tf[T_DM] = (k >= 1) ? expf(m.tsc[(size_t)(k - 1) * H_N + H_DM]) : 0.0f;
    tf[T_MD] = expf(m.tsc[(size_t)k * H_N + H_MD]);
    tf[T_MI] = expf(m.tsc[(size_t)k * H_N + H_MI]);
    tf[T_II] = expf(m.tsc[(size_t)k * H_N + H_II]);
    tf[T_DD] = expf(m.tsc[(size_t)k * H_N + H_DD]);
  }

  // ---- bias-filter emission odds: state 0 = background, state 1 = model composition ----
  for (int x = 0; x < K; ++x) { m.bias_eo[x][0] = BG_F[x] / BG_F[x]; m.bias_eo[x][1] = m.compo[x] / BG_F[x]; }
  for (int s = 0; s < 2; ++s) { m.bias_eo[20][s] = 1.0f; m.bias_eo[27][s] = 1.0f; m.bias_eo[28][s] = 1.0f; }
  for (int x = K + 1; x <= KP - 3; ++x)
    for (int s = 0; s < 2; ++s) {
      float num = 0.0f, denom = 0.0f;
      for (int y = 0; y < K; ++y)
        if (degen_has(x, y)) { num += (s == 0 ? BG_F[y] : m.compo[y]); denom += BG_F[y]; }
      m.bias_eo[x][s] = (denom > 0.0f) ? num / denom : 0.0f;
    }
}

}  // namespace ckm
