// text2bin.cc -- `dw text2bin` and `dw bin2text`: the TSV <-> binary
// converters of doc/text_format.md, behaviour as src/text2bin.cc:19-261 and
// src/bin2text.cc:23-152.  Host-only, buffered I/O; not on the sweep.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/dwg.h"
#include "../graph_image.h"
#include "cmd_parser.h"

namespace dwg {
namespace {

struct BeWriter {
  std::vector<uint8_t> buf;
  FILE *f;
  explicit BeWriter(FILE *f_) : f(f_) { buf.reserve(1 << 20); }
  void flush() {
    if (!buf.empty()) std::fwrite(buf.data(), 1, buf.size(), f);
    buf.clear();
  }
  void maybe_flush() { if (buf.size() > (1 << 20) - 64) flush(); }
  void u8(uint8_t v) { buf.push_back(v); }
  void u16(uint16_t v) { buf.push_back(v >> 8); buf.push_back(v & 0xff); }
  void u64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) buf.push_back((v >> s) & 0xff); }
  void f64(double d) { uint64_t v; std::memcpy(&v, &d, 8); u64(v); }
};

std::vector<std::string> split_tabs(const std::string &line) {
  std::vector<std::string> out;
  size_t pos = 0;
  for (;;) {
    size_t t = line.find('\t', pos);
    out.push_back(line.substr(pos, t == std::string::npos ? std::string::npos : t - pos));
    if (t == std::string::npos) break;
    pos = t + 1;
  }
  return out;
}

// "{a,b,c}" -> elements (parse_pgarray, src/text2bin.cc:77-113)
bool parse_pgarray(const std::string &s, std::vector<std::string> *out) {
  out->clear();
  if (s.size() < 2 || s.front() != '{' || s.back() != '}') return false;
  std::string body = s.substr(1, s.size() - 2);
  std::stringstream ss(body);
  std::string el;
  while (std::getline(ss, el, ',')) out->push_back(el);
  return true;
}

int fail(const std::string &msg) {
  std::cerr << "[ERROR] text2bin: " << msg << std::endl;
  return 1;
}

}  // namespace

int text2bin_main(const CmdParser &a) {
  std::ifstream fin(a.text2bin_input);
  if (!fin) return fail("cannot open " + a.text2bin_input);
  FILE *fo = std::fopen(a.text2bin_output.c_str(), "wb");
  if (!fo) return fail("cannot write " + a.text2bin_output);
  BeWriter w(fo);
  uint64_t count = 0;
  std::string line;
  int rc = 0;
  if (a.text2bin_mode == "variable") {  // :19-48: vid role init type cardinality
    while (std::getline(fin, line)) {
      std::istringstream ss(line);
      uint64_t vid, role, init, type, card;
      if (!(ss >> vid >> role >> init >> type >> card)) { rc = fail("bad variable line: " + line); break; }
      w.u64(vid); w.u8((uint8_t)role); w.u64(init); w.u16((uint16_t)type); w.u64(card);
      w.maybe_flush();
      ++count;
    }
  } else if (a.text2bin_mode == "weight") {  // :51-73: wid isfixed init
    while (std::getline(fin, line)) {
      std::istringstream ss(line);
      uint64_t wid, fixed;
      double init;
      if (!(ss >> wid >> fixed >> init)) { rc = fail("bad weight line: " + line); break; }
      w.u64(wid); w.u8((uint8_t)fixed); w.f64(init);
      w.maybe_flush();
      ++count;
    }
  } else if (a.text2bin_mode == "factor") {  // :115-183; the count output is the EDGE total
    const uint64_t arity = a.text2bin_factor_arity;
    const bool cat = a.text2bin_factor_func_id == DWG_FUNC_AND_CATEGORICAL;
    while (std::getline(fin, line)) {
      std::vector<std::string> f = split_tabs(line);
      const size_t need = arity * (cat ? 2 : 1) + 2;
      if (f.size() < need) { rc = fail("bad factor line: " + line); break; }
      w.u16((uint16_t)a.text2bin_factor_func_id);
      w.u64(arity);
      for (uint64_t i = 0; i < arity; ++i) {
        w.u64((uint64_t)std::atol(f[i].c_str()));
        uint64_t eq;
        if (cat) eq = (uint64_t)std::atol(f[arity + i].c_str());
        else {
          if (i >= a.text2bin_factor_variables_should_equal_to.size()) { rc = fail("missing var_is_positive for position"); break; }
          eq = a.text2bin_factor_variables_should_equal_to[i];
        }
        w.u64(eq);
      }
      if (rc) break;
      size_t k = arity * (cat ? 2 : 1);
      w.u64((uint64_t)std::atol(f[k].c_str()));
      w.f64(std::atof(f[k + 1].c_str()));
      w.maybe_flush();
      count += arity;
    }
  } else if (a.text2bin_mode == "domain") {  // :186-236: vid card {values} {truthiness}
    std::vector<std::string> vals, truth;
    while (std::getline(fin, line)) {
      std::istringstream ss(line);
      uint64_t vid, card;
      std::string ds, ts;
      if (!(ss >> vid >> card >> ds >> ts)) { rc = fail("bad domain line: " + line); break; }
      if (!parse_pgarray(ds, &vals) || !parse_pgarray(ts, &truth) || vals.size() != card || truth.size() != card) {
        rc = fail("Expected arrays of " + std::to_string(card) + " elements: " + line);
        break;
      }
      w.u64(vid); w.u64(card);
      for (uint64_t i = 0; i < card; ++i) {
        w.u64((uint64_t)std::atol(vals[i].c_str()));
        w.f64(std::atof(truth[i].c_str()));
      }
      w.maybe_flush();
      ++count;
    }
  } else {
    std::cerr << "Unsupported type" << std::endl;  // :256-258
    rc = 1;
  }
  w.flush();
  std::fclose(fo);
  if (rc == 0) {
    std::ofstream fc(a.text2bin_count_output);
    fc << count << std::endl;
  }
  return rc;
}

int bin2text_main(const CmdParser &a) {
  MetaCounts meta;
  std::string msg = read_meta(a.fg_file, &meta);
  if (!msg.empty()) return fail(msg);
  RawGraph r;
  int err = 0;
  msg = load_raw(meta, a.variable_file, a.domain_file, a.factor_file, a.weight_file, &r, &err);
  if (!msg.empty()) return fail(msg);
  const std::string dir = a.output_folder;
  const uint64_t nvar = r.var_role.size();
  // dense value <-> sparse value like the loaded FactorGraph would print them
  std::vector<int64_t> dom_of(nvar, -1);
  for (size_t b = 0; b < r.dom_vid.size(); ++b)
    if (r.dom_vid[b] < nvar) dom_of[r.dom_vid[b]] = (int64_t)b;
  auto dense_of = [&](uint64_t v, uint64_t sparse) -> uint64_t {
    if (dom_of[v] < 0) return sparse;
    uint64_t b = dom_of[v], found = sparse;
    for (uint64_t i = r.dom_ptr[b]; i < r.dom_ptr[b + 1]; ++i)
      if (r.dom_val[i] == sparse) found = i - r.dom_ptr[b];
    return found;
  };
  {
    std::ofstream f(dir + "/variables.tsv");  // bin2text.cc:23-50 prints the DENSE evidence value
    for (uint64_t v = 0; v < nvar; ++v) {
      bool ev = r.var_role[v] >= 1;
      uint64_t val = ev ? r.var_init[v] : 0;
      if (val) val = dense_of(v, val);
      f << v << '\t' << (ev ? "1" : "0") << '\t' << val << '\t' << r.var_dtype[v] << '\t' << r.var_card[v] << '\n';
    }
  }
  {
    std::ofstream f(dir + "/domains.tsv");  // :52-69 (values only, in domain order)
    for (uint64_t v = 0; v < nvar; ++v) {
      if (r.var_dtype[v] == 0) continue;
      f << v << '\t' << r.var_card[v] << '\t' << '{';
      for (uint64_t j = 0; j < r.var_card[v]; ++j) {
        if (j) f << ',';
        f << (dom_of[v] >= 0 ? r.dom_val[r.dom_ptr[dom_of[v]] + j] : j);
      }
      f << "}\n";
    }
  }
  {
    std::ofstream f(dir + "/factors.tsv");  // :71-98
    for (size_t i = 0; i < r.fac_func.size(); ++i) {
      for (uint64_t e = r.fac_ptr[i]; e < r.fac_ptr[i + 1]; ++e) f << r.edge_vid[e] << '\t';
      if (r.fac_func[i] == DWG_FUNC_AND_CATEGORICAL)
        for (uint64_t e = r.fac_ptr[i]; e < r.fac_ptr[i + 1]; ++e) f << r.edge_eq[e] << '\t';
      f << r.fac_wid[i] << '\t' << r.fac_fv[i] << '\n';
    }
  }
  {
    std::ofstream f(dir + "/weights.tsv");  // :100-112
    for (size_t i = 0; i < r.w_init.size(); ++i)
      f << i << '\t' << (int)r.w_fixed[i] << '\t' << r.w_init[i] << '\n';
  }
  {
    std::ofstream f(dir + "/graph.meta");  // :114-130
    f << r.w_init.size() << ',' << nvar << ',' << r.fac_func.size() << ',' << r.edge_vid.size()
      << ",graph.weights,graph.variables,graph.factors" << std::endl;
  }
  return 0;
}

}  // namespace dwg
