// TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the reference's jump-aware aligner.
//
// Nothing under stitch_b200/ may include, link or call this.  Only tests/, the smoke check in
// __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs use it, as the checker
// and as the reported CPU baseline.
//
// The reference (fulcrumgenomics/stitch, safe Rust) cannot be built here (no cargo/rustc), so this
// file restates its algorithm in C++.  Parity is PINNED against the reference's own known-answer
// tests (tests/golden/*.json, generated from the reference test modules by
// tests/golden/extract_reference_kats.py): 63 single-contig, 9 multi-contig (13 assertions),
// 7 split_at_y and 1 API case.  Pre-alignment (bio 1.1.0 banded aligner) is not restated:
// parity unpinned for that row.
//
// Citations: LIB = fg-stitch-lib/src of the reference checkout.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace oracle {

// LIB/align/aligners/constants.rs:7
constexpr int32_t MIN_SCORE = -858993459;

// LIB/align/traceback/mod.rs:47-57
enum : uint8_t {
    TB_START = 0, TB_INS = 1, TB_DEL = 2, TB_SUBST = 3, TB_MATCH = 4, TB_XCLIP_PREFIX = 5,
    TB_XCLIP_SUFFIX = 6, TB_YCLIP_PREFIX = 7, TB_YCLIP_SUFFIX = 8, TB_XJUMP = 9
};

// LIB/align/aligners/constants.rs:96-108
enum Mode : int { LOCAL = 0, QUERY_LOCAL = 1, TARGET_LOCAL = 2, GLOBAL = 3, CUSTOM = 4 };

// LIB/align/scoring.rs:11-23 (match_fn is bio's MatchParams: equal bytes -> match score)
struct Scoring {
    int32_t match = 1, mismatch = -1;
    int32_t gap_open = -5, gap_extend = -1;
    int32_t jump_same = -10, jump_opp = -10, jump_inter = -10;
    int32_t xclip_prefix = MIN_SCORE, xclip_suffix = MIN_SCORE;
    int32_t yclip_prefix = MIN_SCORE, yclip_suffix = MIN_SCORE;
    int32_t sub(uint8_t a, uint8_t b) const { return a == b ? match : mismatch; }
    void set_clips_for(Mode m);   // mod.rs:123-131
};

// LIB/align/aligners/constants.rs:20-29
enum OpKind : uint8_t { MATCH = 0, SUBST = 1, DEL = 2, INS = 3, XCLIP = 4, YCLIP = 5, XJUMP = 6, YJUMP = 7 };
struct Op {
    uint8_t kind;
    uint32_t a = 0, b = 0;   // Xclip/Yclip/Yjump: a = len; Xjump: a = contig, b = offset
    bool operator==(const Op &o) const { return kind == o.kind && a == o.a && b == o.b; }
    bool is_special() const { return kind == XCLIP || kind == YCLIP || kind == XJUMP; }
    int64_t len_x(int64_t x_index) const;   // constants.rs:61-72
    int64_t len_y() const;                  // constants.rs:75-84
};

// LIB/align/alignment.rs:16-51
struct Alignment {
    int32_t score = 0;
    int64_t ystart = 0, xstart = 0, yend = 0, xend = 0, ylen = 0, xlen = 0;
    int64_t start_contig_idx = 0, end_contig_idx = 0;
    std::vector<Op> ops;
    Mode mode = CUSTOM;
    int64_t length = 0;
    std::string cigar() const;                    // alignment.rs:105-149
    Alignment split_at_y(int64_t y_pivot) const;  // alignment.rs:207-360
};

struct JumpInfo { int32_t score; uint32_t len, idx, from; };   // mod.rs:56-62

struct SValue { uint8_t tb; uint32_t len, idx, from; };

// 16-byte traceback cell with the same capacity as the reference's PackedLengthCell
// (4-bit move + 27-bit length per layer, 8-bit contig index, 27-bit from;
//  LIB/align/traceback/packed_length_cell.rs:25-30, 107-181).
struct Cell {
    uint64_t w0 = 0, w1 = 0;
    void set_s(uint8_t tb, uint32_t len);
    void set_i(uint8_t tb, uint32_t len);
    void set_d(uint8_t tb, uint32_t len);
    void set_s_all(uint8_t tb, uint32_t len, uint32_t idx, uint32_t from);
    void set_all(uint8_t tb, uint32_t len) { set_i(tb, len); set_d(tb, len); set_s(tb, len); }
    uint8_t s_tb() const; uint32_t s_len() const;
    uint8_t i_tb() const; uint32_t i_len() const;
    uint8_t d_tb() const; uint32_t d_len() const;
    uint32_t idx() const; uint32_t from() const;
    SValue s() const { return SValue{s_tb(), s_len(), idx(), from()}; }
};

// LIB/align/aligners/single_contig_aligner.rs:72-83
struct SingleContig {
    std::vector<int32_t> I[2], D[2], S[2];
    std::vector<int64_t> Lx, Ly;
    std::vector<int32_t> Sn;
    std::vector<Cell> tb;      // (m+1) x (n+1), row-major in i (traceback/mod.rs:102-114)
    int64_t rows = 0, cols = 0;
    // The reference stores cell (i, j) at i * cols + j and fills column by column (a stride of 16 (n+1) bytes per inner
    // step): that is the layout the CPU BASELINE is timed on.  The parity tests on full-length reads set
    // g_checker_layout (oracle_set_checker_layout): cell (i, j) at j * rows + i, same values, several times faster.
    static bool g_checker_layout;
    Scoring sc;
    uint32_t contig_idx = 0;
    bool circular = false;

    Cell &cell(int64_t i, int64_t j) { return tb[(size_t)(g_checker_layout ? j * rows + i : i * cols + j)]; }
    const Cell &cell(int64_t i, int64_t j) const { return tb[(size_t)(g_checker_layout ? j * rows + i : i * cols + j)]; }

    void init_matrices(int64_t m, int64_t n);                                      // :97-186
    void init_column(int64_t j, int curr, int64_t m, int64_t n);                   // :188-239
    JumpInfo jump_info(int64_t m, int64_t j, int32_t jump_score) const;            // :677-697
    void fill_column(const uint8_t *x, const uint8_t *y, int64_t m, int64_t n, int64_t j,
                     int prev, int curr, JumpInfo jump);                           // :292-451
    void fill_last_column(int64_t m, int64_t n);                                   // :453-555
    Alignment custom(const uint8_t *x, int64_t m, const uint8_t *y, int64_t n);    // :705-729
    Alignment with_mode(Mode mode, const uint8_t *x, int64_t m, const uint8_t *y, int64_t n); // :733-872
};

// LIB/align/traceback/mod.rs:129-373.  `ok` is false where the reference returns None.
Alignment traceback_best(const std::vector<const SingleContig *> &al, int64_t n);
bool traceback_from(const std::vector<const SingleContig *> &al, int64_t n, uint32_t contig_index, Alignment &out);
std::vector<Alignment> traceback_all(const std::vector<const SingleContig *> &al, int64_t n,
                                     const std::vector<bool> &consider, size_t n_consider);

// LIB/align/aligners/multi_contig_aligner.rs
struct MultiContig {
    struct Contig { std::string name; bool is_forward; SingleContig aligner; const uint8_t *seq; int64_t len; };
    std::vector<Contig> contigs;
    void add_contig(const std::string &name, bool is_forward, const uint8_t *seq, int64_t len,
                    bool circular, const Scoring &sc);                              // :93-133
    Alignment custom(const uint8_t *y, int64_t n);                                  // :231-361
    Alignment custom_with_subset(const uint8_t *y, int64_t n, const std::vector<bool> *subset); // :178-223
    std::vector<Alignment> traceback_all(int64_t n, const std::vector<bool> *subset);          // :363-378
    bool traceback_from(int64_t n, uint32_t contig_index, Alignment &out);                     // :380-387
    uint64_t cells_filled = 0;   // bookkeeping for GCUPS (not in the reference)
    uint64_t fills = 0;
};

// LIB/align/aligners/mod.rs:65-116 (fields that reach the path)
struct Options {
    Mode mode = LOCAL;
    int32_t match_score = 1, mismatch_score = -4, gap_open = -6, gap_extend = -2;
    int32_t jump_same = -10, jump_opp = -10, jump_inter = -10;
    bool double_strand = false, circular = false, suboptimal = false;
    int64_t circular_slop = 20;
    float suboptimal_pct = 20.0f;
    Scoring contig_scoring() const;   // mod.rs:143-167
};

// LIB/align/aligners/mod.rs:227-553 without the pre-alignment branch
struct Aligners {
    Options opts;
    std::vector<std::string> names;
    std::vector<std::vector<uint8_t>> fwd, rev;
    MultiContig mc;
    Aligners(const Options &o, const std::vector<std::string> &names, const std::vector<std::vector<uint8_t>> &fwd);
    std::vector<Alignment> align(const uint8_t *read, int64_t n, const std::vector<bool> *subset); // :237-340
    Alignment remove_clipping(Alignment a) const;                                                  // :343-353
  private:
    Alignment multi_contig_align(const uint8_t *q, int64_t n, const std::vector<bool> *subset);    // :355-363
    Alignment realign_origin(const std::vector<uint8_t> &q, const Alignment &a, int64_t slop);     // :442-553
    bool realign_and_split(const std::vector<uint8_t> &q, const Alignment &best,
                           const std::vector<bool> &subset, int64_t contig_idx, int64_t y_pivot, Alignment &out); // :412-431
};

std::vector<uint8_t> reverse_complement(const std::vector<uint8_t> &s);   // LIB/util/dna.rs:5-41

}  // namespace oracle
