// loader.cc -- reads the reference's on-disk factor graph (doc/binary_format.md).
//
// Replaces read_meta / FactorGraph::load_{weights,variables,domains,factors}
// (src/binary_format.cc:23-226).  Big-endian records; every input may be a pipe
// (test/run_end_to_end.sh:18-23 passes <(cat ...)), so files are slurped sequentially, one
// thread per file like parallel_load (:38-46), and decoded from memory by all host threads:
// fixed-size records by ranges, the variable-length factor records after one light pass that
// finds block boundaries (the arity is the only variable-length field).  With several
// factor files the reference numbers factors in thread-arrival order (:146); here ids
// follow the order of the -f flags, which is the single-file order when one file is given.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../../include/dwg.h"
#include "../graph_image.h"
#include "parallel.h"

namespace dwg {
namespace {

// One input, in memory: a regular file is mapped (no copy: records are decoded straight from
// the page cache), anything else (pipe, /dev/fd/N) is read to its end into a heap buffer.
struct FileBuf {
  const uint8_t *ptr = nullptr;
  size_t len = 0;
  bool mapped = false;
  std::vector<uint8_t> heap;
  FileBuf() = default;
  FileBuf(const FileBuf &) = delete;
  FileBuf &operator=(const FileBuf &) = delete;
  ~FileBuf() {
    if (mapped && ptr) munmap(const_cast<uint8_t *>(ptr), len);
  }
  size_t size() const { return len; }
  const uint8_t *data() const { return ptr; }
  const uint8_t &operator[](size_t i) const { return ptr[i]; }
};

bool slurp(const std::string &path, FileBuf *buf, std::string *msg) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    *msg = "cannot open " + path;
    return false;
  }
  struct stat st;
  if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
    if (st.st_size == 0) { close(fd); return true; }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (m != MAP_FAILED) {
      madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
      buf->ptr = static_cast<const uint8_t *>(m);
      buf->len = (size_t)st.st_size;
      buf->mapped = true;
      close(fd);
      return true;
    }
  }
  size_t cap = 1 << 20, len = 0;
  buf->heap.resize(cap);
  for (;;) {
    ssize_t n = read(fd, buf->heap.data() + len, cap - len);
    if (n < 0) {
      close(fd);
      *msg = "read error on " + path;
      return false;
    }
    if (n == 0) break;
    len += (size_t)n;
    if (len == cap) {
      cap *= 2;
      buf->heap.resize(cap);
    }
  }
  close(fd);
  buf->heap.resize(len);
  buf->ptr = buf->heap.data();
  buf->len = len;
  return true;
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
  FileBuf buf;
  std::string msg;
  if (!slurp(path, &buf, &msg)) return msg;
  std::string s(buf.data(), buf.data() + buf.size());
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
  const bool timing = getenv("DWG_COMPILE_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    auto n = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "[load]    %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t_prev).count());
    t_prev = n;
  };
  // slurp everything, one thread per file
  std::vector<const std::vector<std::string> *> kinds = {&var_files, &dom_files, &fac_files, &wgt_files};
  std::vector<std::vector<FileBuf>> data(4);
  std::vector<std::string> msgs;
  {
    std::vector<std::thread> th;
    std::vector<std::string> m;
    size_t total = 0;
    for (auto *k : kinds) total += k->size();
    m.resize(total);
    size_t idx = 0;
    for (int k = 0; k < 4; ++k) {
      data[k] = std::vector<FileBuf>(kinds[k]->size());
      for (size_t i = 0; i < kinds[k]->size(); ++i, ++idx) {
        const std::string &path = (*kinds[k])[i];
        FileBuf *dst = &data[k][i];
        std::string *mm = &m[idx];
        th.emplace_back([path, dst, mm]() { slurp(path, dst, mm); });
      }
    }
    for (auto &t : th) t.join();
    for (auto &s : m)
      if (!s.empty()) return s;
  }
  lap("read files");
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
      const uint64_t nrec = b.size() / 27;
      std::atomic<uint64_t> bad{~0ull};  // lowest record index with an id out of range
      parallel_for(nrec, [&](uint64_t i0, uint64_t i1) {
        for (uint64_t i = i0; i < i1; ++i) {
          const uint8_t *q = &b[i * 27];
          uint64_t vid = be64(q);
          if (vid >= meta.nvar) {
            uint64_t cur = bad.load();
            while (i < cur && !bad.compare_exchange_weak(cur, i)) {}
            return;
          }
          r.var_role[vid] = q[8];
          r.var_init[vid] = be64(q + 9);
          r.var_dtype[vid] = be16(q + 17);
          r.var_card[vid] = be64(q + 19);
          seen[vid] = 1;
        }
      });
      if (bad.load() != ~0ull) return "variable id " + std::to_string(be64(&b[bad.load() * 27])) + " out of range";
      count += nrec;
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
  lap("variables, weights, domains");
  // factors (:143-181).  One thread walks the arities and hands out blocks of 2^16 records (byte
  // offset, factor index, edge index) as it finds them; the other host threads decode the blocks
  // while the walk goes on.  The arrays are sized from the meta counts, which the walk confirms
  // (a block that would overrun them is never handed out).
  struct FBlock { const FileBuf *buf; size_t p; uint64_t f, e, nrec; };
  r.fac_func.resize(meta.nfac); r.fac_wid.resize(meta.nfac); r.fac_fv.resize(meta.nfac);
  r.fac_ptr.resize(meta.nfac + 1); r.edge_vid.resize(meta.nedge); r.edge_eq.resize(meta.nedge);
  r.fac_ptr[0] = 0;
  auto decode = [&](const FBlock &B) {
    const uint8_t *q = B.buf->data() + B.p;
    uint64_t e = B.e;
    for (uint64_t f = B.f; f < B.f + B.nrec; ++f) {
      uint16_t func = be16(q);
      uint64_t arity = be64(q + 2);
      q += 10;
      r.fac_func[f] = (int32_t)(int16_t)func == -1 ? -1 : (int32_t)func;
      for (uint64_t i = 0; i < arity; ++i, q += 16, ++e) {
        r.edge_vid[e] = be64(q);
        r.edge_eq[e] = be64(q + 8);
      }
      r.fac_ptr[f + 1] = e;
      r.fac_wid[f] = be64(q);
      r.fac_fv[f] = bef64(q + 8);
      q += 16;
    }
  };
  std::mutex qmu;
  std::condition_variable qcv;
  std::vector<FBlock> queue;  // blocks found so far
  size_t next_block = 0;      // first block no worker has taken
  bool walk_done = false;
  std::vector<std::thread> workers;
  const unsigned nworkers = std::max(1u, host_threads() - 1);
  for (unsigned t = 0; t < nworkers; ++t)
    workers.emplace_back([&] {
      for (;;) {
        FBlock B;
        {
          std::unique_lock<std::mutex> lk(qmu);
          qcv.wait(lk, [&] { return next_block < queue.size() || walk_done; });
          if (next_block >= queue.size()) return;
          B = queue[next_block++];
        }
        decode(B);
      }
    });
  std::string walk_error;
  uint64_t nfac = 0, nedge = 0;
  {
    FBlock cur{nullptr, 0, 0, 0, 0};
    auto hand_out = [&]() {
      if (!cur.nrec) return;
      {
        std::lock_guard<std::mutex> lk(qmu);
        queue.push_back(cur);
      }
      qcv.notify_one();
      cur.nrec = 0;
    };
    for (auto &b : data[2]) {
      size_t p = 0;
      while (p < b.size() && walk_error.empty()) {
        if (!cur.nrec) cur = FBlock{&b, p, nfac, nedge, 0};
        if (p + 10 > b.size()) { walk_error = "factors file: truncated record"; break; }
        uint64_t arity = be64(&b[p + 2]);
        p += 10;
        if (arity > (b.size() - p) / 16 || b.size() - p - 16 * arity < 16) { walk_error = "factors file: truncated record"; break; }
        if (nfac + 1 > meta.nfac || nedge + arity > meta.nedge) {  // more than the meta file announced: reported below
          nfac += 1; nedge += arity;
          p += 16 * arity + 16;
          cur.nrec = 0;  // never decoded: it would not fit the arrays
          continue;
        }
        p += 16 * arity + 16;
        nedge += arity;
        ++nfac;
        if (++cur.nrec == (1u << 16)) hand_out();
      }
      if (nfac <= meta.nfac && nedge <= meta.nedge) hand_out();  // a block does not span files
      cur.nrec = 0;
    }
  }
  {
    std::lock_guard<std::mutex> lk(qmu);
    walk_done = true;
  }
  qcv.notify_all();
  lap("factor boundaries (decode running)");
  for (auto &t : workers) t.join();
  if (!walk_error.empty()) return walk_error;
  if (nfac != meta.nfac) return "number of factors (" + std::to_string(nfac) + ") != meta (" + std::to_string(meta.nfac) + ")";
  if (nedge != meta.nedge) return "number of edges (" + std::to_string(nedge) + ") != meta (" + std::to_string(meta.nedge) + ")";
  lap("factor decode");
  *err = DWG_OK;
  return "";
}

}  // namespace dwg
