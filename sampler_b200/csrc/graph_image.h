// graph_image.h -- the factor graph in the layout the sweep kernels stream.
//
// Replaces the reference's AoS FactorGraph (src/factor_graph.h:58-70: Variable 64 B,
// VariableToFactor 32 B, Factor 48 B, FactorToVariable 16 B, size_t everywhere) by a
// 32-bit SoA CSR in SCHEDULE ORDER: variables are permuted so that each colour class
// (or the whole graph under hogwild) is one contiguous range, and the
// variable -> factor_index -> Factor -> vifs pointer chain (src/factor_graph.h:137-142)
// is flattened into one edge stream per (variable, value) row.
#ifndef DWG_GRAPH_IMAGE_H_
#define DWG_GRAPH_IMAGE_H_
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace dwg {

// compact function codes (4 bits of the edge tag); ISTRUE aliases AND (src/factor.h:73)
enum FuncCode : uint8_t {
  FC_IMPLY_NATURAL = 0, FC_OR = 1, FC_AND = 2, FC_EQUAL = 3, FC_LINEAR = 4, FC_RATIO = 5,
  FC_LOGICAL = 6, FC_AND_CATEGORICAL = 7, FC_IMPLY_MLN = 8
};

// edge tag byte: [3:0] FuncCode | [4] self equal_to == 1 (Boolean rows) | [5] self is the
// head (last position) | [7:6] form: 0 unary, 1 one other, 2 two others, 3 general
// (all positions, self included, spelled out in the others stream after a length word)
constexpr uint8_t TAG_FUNC_MASK = 0x0f;
constexpr uint8_t TAG_SELF_EQ1 = 0x10;
constexpr uint8_t TAG_SELF_HEAD = 0x20;
constexpr int TAG_FORM_SHIFT = 6;
constexpr uint8_t FORM_GENERAL = 3;

// var_flags
constexpr uint8_t VF_EVID = 1;
constexpr uint8_t VF_CATEGORICAL = 2;
constexpr uint8_t VF_TRUTHINESS = 4;  // has_truthiness(), src/variable.h:105
constexpr uint8_t VF_PAD = 0x80;      // filler slot that rounds a colour class up to 32 variables

// ---- sliced-ELL blob (Boolean graphs, parallel schedules) ---------------------------------
// Slice s = the 32 variables d = 32s .. 32s+31 (lane = d & 31).  All data of a slice is ONE
// contiguous run of 32-bit words starting at word 8 * sl_off[s] of `blob`:
//   header  8 words (SH_*: counts, evidence / padding / evidence-value masks, flags, own offset)
//   unary   n_un columns x 32 words      record of lane l in column j at  8 + 32 j + l
//   (fv)    n_un columns x 32 f32 | f64  feature values of the unary records (FV_F32 | FV_F64)
//   multi   n_mu columns x 32 words
//   (fv)    n_mu columns x 32 f32 | f64
//   others  n_oc columns x 32 words      (d << 1 | equal_to) references, or a length word
// so a warp streams its slice with fully coalesced 128 B requests.  Within a row the unary
// records come first, then the multi-variable ones, each group in ascending factor id (the
// CSR stream keeps construct_index's pure factor-id order; the two sums agree to fp64
// reassociation).  Sign codes: 0 -> -1, 1 -> 0, 2 -> +1.
//   unary record:  code(variable = 1) | weight_is_fixed << 2 | wid << 3 | code(variable = 0) << 26,
//               so that  rec & SELL_UN_WID8_MASK  is the byte offset of the weight in an f64 table and
//               code * plane + that offset its entry in the staged kernel's signed weight table (three
//               planes: -w, 0, +w).  Padding records are SELL_UNARY_NOP (weight 0, both codes 1: adds +0.0).
//   multi record:  wid[22:0] | weight_is_fixed << 23 | tag << 24
//   multi tag : slices flagged SF_BINARY (every multi record has exactly one other variable,
//               record j owns others column j): the factor's whole sign table,
//               code(x, sat) at bits 2 * (2 x + sat), x = this variable's value, sat = the
//               other variable satisfies its equal_to; padding = SELL_BINARY_NOP.
//               other slices: the CSR tag byte above (FuncCode | SELF_EQ1 | SELF_HEAD | form),
//               records consume others words through a per-lane cursor; form 0 = padding.
constexpr int SH_N_UNARY = 0, SH_N_MULTI = 1, SH_N_OTH = 2;
constexpr int SH_EVID_MASK = 3;   // bit l: variable of lane l is evidence
constexpr int SH_PAD_MASK = 4;    // bit l: lane l is a filler slot
constexpr int SH_EVVAL_MASK = 5;  // bit l: Variable::assignment_dense of lane l is 1
constexpr int SH_FLAGS = 6;       // SF_*
constexpr int SH_SELF_OFF = 7;    // sl_off of this slice: a staged copy of the header can find the rest in HBM
constexpr int SH_WORDS = 8;
constexpr uint32_t SF_ANTISYM = 1;  // every unary record has sign(1) == -sign(0): neg == -pos
constexpr uint32_t SF_BINARY = 2;   // every multi record has exactly one other variable
constexpr uint32_t SF_WIDE = 4;     // ONE variable (lane 0's slot), its records dealt round-robin to the 32 lanes:
                                    // record i of a section at column i / 32, lane i % 32; the lanes' partial sums
                                    // are combined by shuffles (warp per variable: hubs)
constexpr uint64_t SELL_WIDE_MIN = 32;   // incident records from which a variable may get a slice of its own (compile.cc)
// A hub with more than SELL_PART_RECORDS records is cut into PARTS: consecutive SF_WIDE | SF_PART
// slices, part p holding an equal share of the unary and of the multi records.  Any warp sweeps
// any part; a part's lane sums are combined, stored, and the warp that finishes LAST adds the
// parts' sums in index order and draws (deterministic).  flags: part index << 4 (14 bits), number
// of parts << 18 (14 bits).  Every part's header repeats the variable's evidence bits at lane 0;
// the variable itself sits in slot 0 of the first part.
constexpr uint32_t SF_PART = 8;
constexpr uint64_t SELL_PART_RECORDS = 2048;
constexpr uint32_t SELL_MAX_PARTS = 1u << 14;
constexpr uint32_t SELL_WID_MASK = 0x7fffffu, SELL_FIXED_BIT = 0x800000u;  // multi records
constexpr uint32_t SELL_BINARY_NOP = 0x55u;                                // tag of a padding binary record
constexpr uint32_t SELL_UN_WID8_MASK = 0x3fffff8u, SELL_UN_FIXED_BIT = 1u << 2;
constexpr uint32_t SELL_UNARY_NOP = 1u | (1u << 26);                       // a whole padding unary record
inline constexpr uint32_t sell_unary_record(uint32_t wid, bool fixed, uint32_t code1, uint32_t code0) {
  return code1 | (fixed ? SELL_UN_FIXED_BIT : 0u) | (wid << 3) | (code0 << 26);
}
inline constexpr uint32_t sell_unary_wid(uint32_t rec) { return (rec & SELL_UN_WID8_MASK) >> 3; }
constexpr uint64_t SELL_MAX_WEIGHTS = 1ull << 23;

// ---- categorical sliced-ELL blob (graphs whose variables are ALL categorical with cardinality
// <= CELL_MAX_CARD and whose factors touch at most two variables; parallel schedules) ----------
// Slice = 32 variables, lane = variable, one contiguous run of words at 8 * sl_off[s]:
//   header  8 words  CH_*: record columns, others columns, evidence / padding masks, own offset, max cardinality
//   meta    32 words per lane: cardinality | evidence value << 8
//   row0    32 words per lane: first tally row of the variable (rows of a variable are consecutive)
//   records n_rec columns x 32 words, per lane the records of value 0, then of value 1, ... each
//           row in factor-id order (the reference's order, src/factor_graph.cc:179-195):
//             wid[21:0] | table[29:22] | consumes-an-others-word << 30 | first-record-of-a-row << 31
//           table = sign codes (0 -> -1, 1 -> 0, 2 -> +1) at bits 2 * (2 * self_sat + other_sat); a
//           row without factors holds one CELL_NOP_ROW record, a lane's tail CELL_NOP records
//   others  n_oc columns x 32 words: (d << 8 | dense equal_to) of the other variable, consumed in order
constexpr int CH_N_REC = 0, CH_N_OTH = 1, CH_EVID_MASK = 2, CH_PAD_MASK = 3, CH_KMAX = 4, CH_SELF_OFF = 7;
constexpr uint32_t CELL_MAX_CARD = 32, CELL_MAX_OTH = 32;
constexpr uint64_t CELL_MAX_WEIGHTS = 1ull << 22, CELL_MAX_VARS = 1ull << 24;
constexpr uint32_t CELL_WID_MASK = 0x3fffffu, CELL_MULTI_BIT = 1u << 30, CELL_ROW_BIT = 1u << 31;
constexpr uint32_t CELL_NOP = 0x55u << 22, CELL_NOP_ROW = CELL_NOP | CELL_ROW_BIT;  // every code 1: adds +0.0

// weights with internal id < n_hot are staged in shared memory by the sweep kernels
#define DWG_HOT_WEIGHTS 4096

enum FvMode : int { FV_ONE = 0, FV_F32 = 1, FV_F64 = 2 };

struct GraphImage {
  // sizes (reference FactorGraphDescriptor, src/factor_graph.h:18-46)
  uint64_t nvar = 0, nfac = 0, nweight = 0, nedge = 0, nrow = 0;
  uint64_t nvar_evid = 0, nvar_query = 0;
  uint64_t nindex = 0;  // edge records after dedupe  (E')
  uint64_t noth = 0;    // entries in the others stream
  int schedule = 0;     // resolved DWG_SCHED_*
  bool has_categorical = false, has_multi = false;
  bool oth_packed = true;  // others as (d<<1 | eq) words; else (d, eq) pairs
  int fv_mode = FV_ONE;
  int assign_bytes = 1;  // 1: u8 chains (max cardinality <= 256), 4: u32
  uint64_t nvar_pad = 0;  // variables in schedule order incl. class padding (== nvar unless sell)
  uint64_t ndrow = 0;     // device rows: nrow, or nvar_pad for sell graphs
  bool sell = false;      // Boolean-only + parallel schedule: sliced-ELL arrays below are built
  bool csr_built = true;  // false: a streamed synthetic image (synth.cc) -- no CSR stream, no per-value host tables
  bool cell = false;      // categorical-only + parallel schedule: the categorical sliced-ELL blob (CELL_*) is built
  uint32_t cell_kmax = 0; // largest cardinality (CELL graphs)

  // schedule order
  std::vector<uint32_t> perm;         // [nvar]   d -> original vid
  std::vector<uint32_t> inv;          // [nvar]   original vid -> d
  std::vector<uint32_t> color_start;  // [ncolor+1] in d space

  // per variable (d order)
  std::vector<uint8_t> var_flags;   // [nvar]
  std::vector<uint32_t> evid_val;   // [nvar] Variable::assignment_dense
  std::vector<uint8_t> evid_val8;   // the same as bytes when assign_bytes == 1 (what the device holds)
  std::vector<uint32_t> var_row;    // [nvar+1] first row of each variable (only if has_categorical)

  // per row = (variable, dense value); Boolean variables own one row
  std::vector<uint32_t> row_ptr;    // [nrow+1] into the edge arrays
  std::vector<uint32_t> row_oth;    // [nrow+1] into oth (only if has_multi)
  std::vector<double> truthiness;   // [nrow]   (only if any is non-zero)
  bool has_truthiness = false;

  // per edge record, in (row, factor id) order == construct_index order
  std::vector<uint32_t> e_wid;
  std::vector<float> e_fv32;
  std::vector<double> e_fv64;
  std::vector<uint8_t> e_tag;
  std::vector<uint32_t> oth;        // others stream (words; pairs take two)

  // sliced-ELL blob (only if sell): see the layout comment above
  std::vector<uint32_t> sl_off;    // [nslice+1] first word of each slice / 8
  std::vector<uint32_t> blob;
  uint32_t sell_max_words = 0;     // longest slice, in words
  // Learning sweeps visit the slices in ascending id block of their variables (classes
  // interleaved: Hogwild, like the reference's threads), inference sweeps class by class.
  uint32_t id_block = 0;               // variables per id block (compile.cc)
  std::vector<uint32_t> learn_visit;   // [nslice] slice indices in learning order
  std::vector<uint32_t> part_slices;   // slices that are parts of a hub (SF_PART), ascending
  bool has_wide = false;               // some slice is SF_WIDE: the kernels with the warp-per-variable code run

  // weights: ABI order is the original id; the device uses internal ids, hottest first
  std::vector<double> w_init;      // original order
  std::vector<uint8_t> w_fixed;    // original order
  std::vector<uint32_t> w_perm;    // internal -> original
  std::vector<uint32_t> w_inv;     // original -> internal
  uint32_t n_hot = 0;              // internal ids below this are staged in shared memory
  // per (colour class, internal weight): SGD updates per learning launch and their curvature
  // bound; *_evid = evidence variables learn (default), *_all = --learn_non_evidence
  bool cls_shared = false;         // one class-averaged row instead of ncolor rows
  // N = number of SGD steps, H = their own curvature bound (diagonal), Q = the row-sum
  // majorizer of the joint curvature (compile.cc)
  std::vector<float> cls_N_evid, cls_Q_evid, cls_H_evid, cls_N_all, cls_Q_all, cls_H_all;

  // host-only, original id order, for the ABI getters / writers
  std::vector<uint8_t> var_dtype, var_is_evid;
  std::vector<uint64_t> value_base;  // [nvar+1] var_val_base in ORIGINAL order
  std::vector<uint64_t> values;      // [nrow] sparse domain value of each (vid, k)
  std::vector<uint32_t> row_of_value;  // [nrow] original value index -> device row

  // ALGORITHMIC bytes per sweep: all variables sampled / query variables only / learning
  double bytes_infer = 0, bytes_infer_query = 0, bytes_learn = 0;
  uint32_t ncolor() const { return (uint32_t)color_start.size() - 1; }
};

// std::vector whose resize() leaves new elements uninitialised: the big per-factor / per-edge
// arrays are sized once and then filled by parallel ranges, so a zero-fill would be a second,
// single-threaded pass over gigabytes (and would fault every page in on one thread)
template <typename T>
struct NoInitAlloc : std::allocator<T> {
  template <typename U> struct rebind { using other = NoInitAlloc<U>; };
  template <typename U> void construct(U *p) noexcept { ::new (static_cast<void *>(p)) U; }
  template <typename U, typename... A> void construct(U *p, A &&...a) {
    ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
  }
};
template <typename T>
using BulkVec = std::vector<T, NoInitAlloc<T>>;

// decoded input records (owning)
struct RawGraph {
  std::vector<uint8_t> var_role;
  std::vector<uint64_t> var_init;
  std::vector<uint16_t> var_dtype;
  std::vector<uint64_t> var_card;
  std::vector<uint64_t> dom_vid, dom_ptr{0}, dom_val;
  std::vector<double> dom_truth;
  BulkVec<int32_t> fac_func;
  BulkVec<uint64_t> fac_ptr{0}, edge_vid, edge_eq, fac_wid;
  BulkVec<double> fac_fv;
  std::vector<double> w_init;
  std::vector<uint8_t> w_fixed;
};

// non-owning view (what dwg_raw_t carries)
struct RawView {
  uint64_t nvar, ndom, nfac, nweight;
  const uint8_t *var_role; const uint64_t *var_init; const uint16_t *var_dtype;
  const uint64_t *var_card; const uint64_t *dom_vid, *dom_ptr, *dom_val; const double *dom_truth;
  const int32_t *fac_func; const uint64_t *fac_ptr, *edge_vid, *edge_eq, *fac_wid;
  const double *fac_fv; const double *w_init; const uint8_t *w_fixed;
};
RawView view_of(const RawGraph &r);

// compile.cc: returns "" on success, else the error message (code in *err)
std::string compile_graph(const RawView &raw, int schedule, GraphImage *out, int *err);

// synth.cc: the image of a synthetic unary Boolean graph, streamed (never holds the records)
std::string build_synthetic_unary(uint64_t nvar, uint32_t degree, uint64_t nweight, double zipf, bool real_features,
                                  double p_evid, uint64_t seed, GraphImage *out, std::vector<double> *cdf_out, int *err);

// loader.cc: reads the reference's binary files; returns "" or an error message
struct MetaCounts { uint64_t nweight = 0, nvar = 0, nfac = 0, nedge = 0; };
std::string read_meta(const std::string &path, MetaCounts *m);
std::string load_raw(const MetaCounts &meta, const std::vector<std::string> &var_files,
                     const std::vector<std::string> &dom_files,
                     const std::vector<std::string> &fac_files,
                     const std::vector<std::string> &wgt_files, RawGraph *out, int *err);

}  // namespace dwg
#endif
