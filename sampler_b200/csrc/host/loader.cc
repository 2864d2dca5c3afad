// loader.cc -- reads the reference's on-disk factor graph (doc/binary_format.md).
//
// Replaces read_meta / FactorGraph::load_{weights,variables,domains,factors}
// (src/binary_format.cc:23-226).  Big-endian records; every input may be a pipe
// (test/run_end_to_end.sh:18-23 passes <(cat ...)), so files are slurped sequentially and
// decoded from memory, one thread per file like parallel_load (:38-46).  With several
// factor files the reference numbers factors in thread-arrival order (:146); here ids
// follow the order of the -f flags, which is the single-file order when one file is given.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/dwg.h"
#include "../graph_image.h"

namespace dwg {
namespace {

bool slurp(const std::string &path, std::vector<uint8_t> *buf, std::string *msg) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) {
    *msg = "cannot open " + path;
    return false;
  }
  buf->clear();
  size_t cap = 1 << 20, len = 0;
  buf->resize(cap);
  for (;;) {
    size_t n = std::fread(buf->data() + len, 1, cap - len, f);
    len += n;
    if (n == 0) break;
    if (len == cap) {
      cap *= 2;
      buf->resize(cap);
    }
  }
  bool bad = std::ferror(f);
  std::fclose(f);
  buf->resize(len);
  if (bad) *msg = "read error on " + path;
  return !bad;
}

inline uint64_t be64(const uint8_t *p) {
  uint64_t x;
  std::memcpy(&x, p, 8);
  return __builtin_bswap64(x);
}
inline uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }
inline double bef64(const uint8_t *p) {
  uint64_t x = be64(p);
  double d;
  std::memcpy(&d, &x, 8);
  return d;
}

}  // namespace

// numWeights,numVariables,numFactors,numEdges[,ignored...]; atoll on the first four
// comma-separated fields (src/binary_format.cc:23-36)
std::string read_meta(const std::string &path, MetaCounts *m) {
  std::vector<uint8_t> buf;
  std::string msg;
  if (!slurp(path, &buf, &msg)) return msg;
  std::string s(buf.begin(), buf.end());
  uint64_t vals[4] = {0, 0, 0, 0};
  size_t pos = 0;
  for (int i = 0; i < 4; ++i) {
    size_t c = s.find(',', pos);
    std::string field = s.substr(pos, c == std::string::npos ? std::string::npos : c - pos);
    vals[i] = (uint64_t)std::atoll(field.c_str());
    if (c == std::string::npos) break;
    pos = c + 1;
  }
  m->nweight = vals[0]; m->nvar = vals[1]; m->nfac = vals[2]; m->nedge = vals[3];
  return "";
}

std::string load_raw(const MetaCounts &meta, const std::vector<std::string> &var_files,
                     const std::vector<std::string> &dom_files,
                     const std::vector<std::string> &fac_files,
                     const std::vector<std::string> &wgt_files, RawGraph *out, int *err) {
  *err = DWG_ERR_IO;
  // slurp everything, one thread per file
  std::vector<const std::vector<std::string> *> kinds = {&var_files, &dom_files, &fac_files, &wgt_files};
  std::vector<std::vector<std::vector<uint8_t>>> data(4);
  std::vector<std::string> msgs;
  {
    std::vector<std::thread> th;
    std::vector<std::string> m;
    size_t total = 0;
    for (auto *k : kinds) total += k->size();
    m.resize(total);
    size_t idx = 0;
    for (int k = 0; k < 4; ++k) {
      data[k].resize(kinds[k]->size());
      for (size_t i = 0; i < kinds[k]->size(); ++i, ++idx) {
        const std::string &path = (*kinds[k])[i];
        std::vector<uint8_t> *dst = &data[k][i];
        std::string *mm = &m[idx];
        th.emplace_back([path, dst, mm]() { slurp(path, dst, mm); });
      }
    }
    for (auto &t : th) t.join();
    for (auto &s : m)
      if (!s.empty()) return s;
  }
  *err = DWG_ERR_INVALID;
  RawGraph &r = *out;
  r = RawGraph();

  // variables: 27-byte records placed by id (:86-111)
  r.var_role.assign(meta.nvar, 0); r.var_init.assign(meta.nvar, 0);
  r.var_dtype.assign(meta.nvar, 0); r.var_card.assign(meta.nvar, 0);
  {
    std::vector<uint8_t> seen(meta.nvar, 0);
    uint64_t count = 0;
    for (auto &b : data[0]) {
      if (b.size() % 27) return "variables file: truncated record";
      for (size_t p = 0; p < b.size(); p += 27) {
        uint64_t vid = be64(&b[p]);
        if (vid >= meta.nvar) return "variable id " + std::to_string(vid) + " out of range";
        r.var_role[vid] = b[p + 8];
        r.var_init[vid] = be64(&b[p + 9]);
        r.var_dtype[vid] = be16(&b[p + 17]);
        r.var_card[vid] = be64(&b[p + 19]);
        seen[vid] = 1;
        ++count;
      }
    }
    // safety_check: counts equal meta and ids dense (src/factor_graph.cc:203-218)
    if (count != meta.nvar) return "number of variables (" + std::to_string(count) + ") != meta (" + std::to_string(meta.nvar) + ")";
    for (uint64_t v = 0; v < meta.nvar; ++v)
      if (!seen[v]) return "variable ids are not dense: missing " + std::to_string(v);
  }
  // weights: 17-byte records (:58-63)
  r.w_init.assign(meta.nweight, 0.0); r.w_fixed.assign(meta.nweight, 0);
  {
    std::vector<uint8_t> seen(meta.nweight, 0);
    uint64_t count = 0;
    for (auto &b : data[3]) {
      if (b.size() % 17) return "weights file: truncated record";
      for (size_t p = 0; p < b.size(); p += 17) {
        uint64_t wid = be64(&b[p]);
        if (wid >= meta.nweight) return "weight id " + std::to_string(wid) + " out of range";
        r.w_fixed[wid] = b[p + 8] != 0;  // Weight::isfixed is a bool
        r.w_init[wid] = bef64(&b[p + 9]);
        seen[wid] = 1;
        ++count;
      }
    }
    if (count != meta.nweight) return "number of weights (" + std::to_string(count) + ") != meta (" + std::to_string(meta.nweight) + ")";
    for (uint64_t w = 0; w < meta.nweight; ++w)
      if (!seen[w]) return "weight ids are not dense: missing " + std::to_string(w);
  }
  // domains (:202-214)
  for (auto &b : data[1]) {
    size_t p = 0;
    while (p < b.size()) {
      if (p + 16 > b.size()) return "domains file: truncated block header";
      uint64_t vid = be64(&b[p]), card = be64(&b[p + 8]);
      p += 16;
      if (card > (b.size() - p) / 16) return "domains file: truncated block";
      r.dom_vid.push_back(vid);
      for (uint64_t i = 0; i < card; ++i, p += 16) {
        r.dom_val.push_back(be64(&b[p]));
        double t = bef64(&b[p + 8]);
        if (!(t >= 0 && t <= 1)) return "domains file: truthiness outside [0,1]";  // :213
        r.dom_truth.push_back(t);
      }
      r.dom_ptr.push_back(r.dom_val.size());
    }
  }
  // factors (:143-181)
  r.fac_func.reserve(meta.nfac); r.fac_wid.reserve(meta.nfac); r.fac_fv.reserve(meta.nfac);
  r.fac_ptr.reserve(meta.nfac + 1); r.edge_vid.reserve(meta.nedge); r.edge_eq.reserve(meta.nedge);
  for (auto &b : data[2]) {
    size_t p = 0;
    while (p < b.size()) {
      if (p + 10 > b.size()) return "factors file: truncated record";
      uint16_t func = be16(&b[p]);
      uint64_t arity = be64(&b[p + 2]);
      p += 10;
      if (arity > (b.size() - p) / 16 || b.size() - p - 16 * arity < 16) return "factors file: truncated record";
      r.fac_func.push_back((int32_t)(int16_t)func == -1 ? -1 : (int32_t)func);
      for (uint64_t i = 0; i < arity; ++i, p += 16) {
        r.edge_vid.push_back(be64(&b[p]));
        r.edge_eq.push_back(be64(&b[p + 8]));
      }
      r.fac_ptr.push_back(r.edge_vid.size());
      r.fac_wid.push_back(be64(&b[p]));
      r.fac_fv.push_back(bef64(&b[p + 8]));
      p += 16;
    }
  }
  if (r.fac_func.size() != meta.nfac) return "number of factors (" + std::to_string(r.fac_func.size()) + ") != meta (" + std::to_string(meta.nfac) + ")";
  if (r.edge_vid.size() != meta.nedge) return "number of edges (" + std::to_string(r.edge_vid.size()) + ") != meta (" + std::to_string(meta.nedge) + ")";
  *err = DWG_OK;
  return "";
}

}  // namespace dwg
