// dp_core.h — the arithmetic of stitch's jump-aware affine-gap DP, decomposed for a GPU.
//
// Everything here is __host__ __device__ so that the exact same code runs inside the CUDA
// kernels (fill_kernel.cu) and inside the CPU lane/tile emulator the tests fuzz against the
// oracle (tests/emul/).  Citations: SCA = fg-stitch-lib/src/align/aligners/single_contig_aligner.rs,
// MCA = .../multi_contig_aligner.rs, TB = fg-stitch-lib/src/align/traceback/mod.rs of the reference.
//
// Decomposition (DESIGN.md has the derivations):
//  * one column j of one contig is cut into warp TILES of 32 lanes x STRIP consecutive rows;
//  * pass A (per lane, no dependency inside the column): D, diagonal, jump, clips -> H' = the
//    best S candidate that does not come from the I layer, plus "H' > A" (A = best of
//    {start, diag, D}), which is all the S/I merge needs;
//  * the in-column insertion chain I(i) = max_k<i H'(k)+o+e(i-k) (earliest k on ties) is an
//    associative max-plus scan over lane aggregates (warp shuffles on the GPU);
//  * pass B (per lane, sequential over its STRIP rows): I, S = merge(H', I), trackers, packed
//    traceback byte;
//  * row m of every contig doubles as the x-suffix-clip accumulator (SCA:407-429), so it is
//    finished per contig after a reduction over rows < m ("finalize");
//  * the per-column best cell of every contig feeds the jump selection of the next column
//    (MCA:279-331).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SHD __host__ __device__ __forceinline__
#define STITCH_UNROLL _Pragma("unroll")
#else
#define SHD inline
#define STITCH_UNROLL
#endif

namespace stitch {

constexpr int32_t MIN_SCORE = -858993459;   // aligners/constants.rs:7
#ifndef STITCH_STRIP
#define STITCH_STRIP 8
#endif
constexpr int STRIP = STITCH_STRIP;          // rows per lane (tests shrink it to force multi-tile contigs)
constexpr int TILE = 32 * STRIP;             // rows per warp tile
constexpr uint32_t MAX_STRANDS = 256;        // contig-strands ONE READ may be aligned against: packed_length_cell.rs:112-114 (8-bit
                                             // contig index of the reference's cell); the layout position of a contig is 8-bit here too
constexpr uint32_t MAX_TABLE_STRANDS = 65536;   // contig-strands the table may hold (a read then needs a subset: pre-alignment)
constexpr uint32_t MAX_CONTIG_LEN = 134217727u;

// Reference traceback codes (TB:47-57).
enum : uint32_t {
    TB_START = 0, TB_INS = 1, TB_DEL = 2, TB_SUBST = 3, TB_MATCH = 4, TB_XCLIP_PREFIX = 5,
    TB_XCLIP_SUFFIX = 6, TB_YCLIP_PREFIX = 7, TB_YCLIP_SUFFIX = 8, TB_XJUMP = 9
};

// Packed traceback byte of an interior cell (1 <= j <= n, 1 <= i <= m):
//   bits 0-3 S move (MV_*), bit 4 D pointer is an extension, bit 5 I pointer is an extension.
// An "open" I/D pointer is resolved at walk time as the S move of the cell it left
// (SCA:324-325, 336-337), which is final by then except in column n (kept in LastCell).
enum : uint32_t {
    MV_START = 0, MV_INS = 1, MV_DEL = 2, MV_DIAG = 3, MV_JUMP = 4, MV_XCLIP_PREFIX = 5,
    MV_XCLIP_SUFFIX = 6, MV_YCLIP_PREFIX = 7, MV_WRAP = 8
};
constexpr uint32_t TBB_DEXT = 16, TBB_IEXT = 32;

struct Scoring {
    int32_t match, mismatch, o, e;           // o = gap_open, e = gap_extend
    int32_t g_same, g_opp, g_inter;
    int32_t xp, xs, yp, ys;                  // clip penalties (AMOD:123-131)
};

// Rolling column state of one cell (what column j+1 needs from column j).
struct CellState { int32_t S, D; uint32_t sl, dl; };

struct JumpInfo { int32_t score; uint32_t len, idx, from; };

// Per (contig, column) record: the jump selected for the column (so that one contig can be
// re-filled alone from a checkpoint) and Lx[j] of the x-suffix tracker (walk).
struct ColRec { int32_t jscore; uint32_t jlen, jidx, jfrom, lx, pad0, pad1, pad2; };

// What a re-fill needs about row m of a contig at a checkpoint column.
struct CkSum { int32_t Sm; uint32_t slm, tbm, pad; };

// Column-n cell in reference form (the end-of-read fix-up edits these, SCA:453-555).
struct LastCell {
    int32_t S, I;
    uint32_t sl, il, idx, from;
    uint8_t s_tb, i_tb, flags, pad;   // flags: bit0 I pointer is an extension, bit1 D pointer is an extension
    uint32_t pad2;
};

// y-suffix tracker of one row (Sn / Ly of SCA:432-447 plus what SCA:482-490 later reads).
struct SnRec { int32_t sn; uint32_t len, ly, idx; };

// One contig-strand of a read's layout.
struct ContigEntry {
    uint32_t contig_idx;   // index among all contig-strands (forward contigs, then reverse)
    uint32_t m;
    uint32_t tile_start, ntiles;
    int32_t opp;           // position (in this layout) of the opposite strand, or -1
    uint32_t seq_off;      // offset of the bases in the contig blob
    uint32_t circular;
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------
// Row 0 of column j (SCA:188-239).  Identical for every contig.  Sn[0]/Ly[0] never move off
// their initial (ys, n): S(0,j) <= 0 = S(0,0) for every j.
// ---------------------------------------------------------------------------------------------
struct Row0 { int32_t S, D; uint32_t sl, dl; uint32_t s_tb, d_tb; };

SHD Row0 row0_at(const Scoring &sc, uint32_t j, uint32_t n) {
    Row0 r;
    if (j == 0) { r.S = 0; r.D = MIN_SCORE; r.sl = 0; r.dl = 0; r.s_tb = TB_START; r.d_tb = TB_START; return r; }
    if (j == 1) {
        r.D = sc.o + sc.e; r.d_tb = TB_START; r.dl = 1;
    } else {
        int32_t d = sc.o + sc.e * (int32_t)j, c = sc.yp + sc.o + sc.e;
        if (d > c) { r.D = d; r.d_tb = TB_DEL; r.dl = j; }
        else { r.D = c; r.d_tb = TB_YCLIP_PREFIX; r.dl = 0; }
    }
    if (r.D > sc.yp) { r.S = r.D; r.s_tb = TB_DEL; r.sl = j; }
    else { r.S = sc.yp; r.s_tb = TB_YCLIP_PREFIX; r.sl = 0; }
    if (j == n && sc.ys > r.S) { r.S = sc.ys; r.s_tb = TB_YCLIP_SUFFIX; r.sl = 0; }   // Sn[0] == ys
    return r;
}

// ---------------------------------------------------------------------------------------------
// Column 0 (SCA:97-186).
// ---------------------------------------------------------------------------------------------
struct Col0 { int32_t S, I; uint32_t sl, il; uint32_t s_tb, i_tb; };

// Value S[.][m] holds after rows < m of column 0 (x-suffix tracker); lx = Lx[0].
SHD void col0_tracker(const Scoring &sc, uint32_t m, int32_t &tm, uint32_t &lx) {
    tm = MIN_SCORE; lx = 0;
    if (m >= 2) {
        int32_t s1 = sc.o + sc.e;               // S(1,0) = max(I(1,0), xp)
        if (sc.xp > s1) s1 = sc.xp;
        if (s1 + sc.xs > MIN_SCORE) { tm = s1 + sc.xs; lx = m - 1; }
    }
}

SHD Col0 col0_at(const Scoring &sc, uint32_t i, uint32_t m) {   // i >= 1
    Col0 c;
    if (i == 1) { c.I = sc.o + sc.e; c.i_tb = TB_START; c.il = 1; }
    else {
        int32_t is = sc.o + sc.e * (int32_t)i, cs = sc.xp + sc.o + sc.e;
        if (is > cs) { c.I = is; c.i_tb = TB_INS; c.il = i; }
        else { c.I = cs; c.i_tb = TB_XCLIP_PREFIX; c.il = 0; }
    }
    c.S = MIN_SCORE; c.s_tb = TB_START; c.sl = 0;
    if (i == m) { uint32_t lx; col0_tracker(sc, m, c.S, lx); c.s_tb = TB_XCLIP_SUFFIX; }
    if (c.I > c.S) { c.S = c.I; c.s_tb = TB_INS; c.sl = i; }
    if (sc.xp > c.S) { c.S = sc.xp; c.s_tb = TB_XCLIP_PREFIX; c.sl = 0; }
    return c;
}

// ---------------------------------------------------------------------------------------------
// Insertion-chain carry: the best candidate for I arriving at some row, with its length and
// whether it opens from the row immediately before (the only case the I pointer is "open").
// ---------------------------------------------------------------------------------------------
struct ICarry { int32_t v; uint32_t il; uint32_t open; };

// `a` covers earlier rows and arrives `dist` rows before `b`'s arrival row; earlier wins ties
// (extension is preferred, SCA:321).
SHD ICarry icarry_combine(ICarry a, uint32_t dist, int32_t e, ICarry b) {
    int32_t av = a.v + e * (int32_t)dist;
    if (av >= b.v) { ICarry r; r.v = av; r.il = a.il + dist; r.open = 0; return r; }
    return b;
}

// I candidate arriving at row 1 of a contig in column j (from row 0, SCA:317-326 with i = 1).
SHD ICarry icarry_row1(const Scoring &sc, const Row0 &r0) {
    int32_t ext = MIN_SCORE + sc.e, open = r0.S + sc.o + sc.e;
    ICarry c;
    if (ext >= open) { c.v = ext; c.il = 1; c.open = 0; }
    else { c.v = open; c.il = r0.sl + 1; c.open = 1; }
    return c;
}

// ---------------------------------------------------------------------------------------------
// Per-cell results of pass A.
// ---------------------------------------------------------------------------------------------
struct PassA {
    int32_t D; uint32_t dl;
    int32_t H; uint32_t hl;                  // best non-I candidate and its length
    uint32_t fl;                             // bits 0-3 move of H, bit 4 D is an extension,
                                             // bit 6 "H > best of {start, diag, D}"
};
constexpr uint32_t PA_GTA = 64;

// Reference (idx, from) fields of the S pointer a packed move stands for (SCA:360-398); the
// untouched start value keeps the zeros of Cell::default().
SHD void ptr_of_move(uint32_t mv, uint32_t self_idx, uint32_t i, uint32_t m, const JumpInfo &J,
                     uint32_t &idx, uint32_t &from) {
    switch (mv) {
    case MV_JUMP: idx = J.idx; from = J.from; break;
    case MV_WRAP: idx = self_idx; from = m; break;
    case MV_DEL: case MV_YCLIP_PREFIX: idx = self_idx; from = i; break;
    case MV_XCLIP_PREFIX: idx = self_idx; from = 0; break;
    case MV_XCLIP_SUFFIX: idx = 0; from = 0; break;
    default: idx = self_idx; from = i - 1; break;   // MV_DIAG, MV_INS
    }
}

// Everything row m needs to be finished once the x-suffix tracker of rows < m is known.
struct RowM {
    int32_t diag; uint32_t dgl;
    int32_t D; uint32_t dl, dext;
    int32_t I; uint32_t il, iext;
    JumpInfo jp;
    int32_t xclip; uint32_t xclip_len;
    int32_t yclip; uint32_t yclip_len;
    uint32_t is_match;
};

struct ColConst {          // per column, per read
    uint32_t j, n;
    int32_t xclip_score;   // xp + max(yp, o + e*j)   (SCA:304-308)
    uint32_t sl0j;         // s_len of cell (0, j)
    uint8_t q;             // read base y[j-1]
};

SHD uint32_t col0_slen(const Scoring &sc, uint32_t i, uint32_t m) { return col0_at(sc, i, m).sl; }

// D layer (SCA:329-338).
SHD void d_layer(const Scoring &sc, const CellState &up, int32_t &D, uint32_t &dl, uint32_t &dext) {
    int32_t ext = up.D + sc.e, open = up.S + sc.o + sc.e;
    if (ext >= open) { D = ext; dl = up.dl + 1; dext = 1; }
    else { D = open; dl = up.sl + 1; dext = 0; }
}

// Jump candidate of a cell (SCA:242-290): J + substitution score, with the circular zero-cost
// wrap from row m of the previous column at i == 1.
SHD JumpInfo jump_for_cell(JumpInfo J, int32_t addend, bool wrap_ok, int32_t Sm_prev, uint32_t slm_prev,
                           uint32_t self_idx, uint32_t m, uint32_t &is_wrap) {
    J.score += addend;
    is_wrap = 0;
    if (wrap_ok) {
        int32_t w = Sm_prev + addend;
        if (!(J.score > w)) {
            uint32_t wl = slm_prev + 1;
            if (!(w == J.score && wl <= J.len)) {
                J.score = w; J.len = wl; J.idx = self_idx; J.from = m; is_wrap = 1;
            }
        }
    }
    return J;
}

// Pass A for a row i < m (SCA:340-399 without the I layer).
//   up   = state of (i, j-1);  dg = state of (i-1, j-1)
SHD PassA pass_a(const Scoring &sc, const ColConst &cc, const CellState &up, int32_t dgS, uint32_t dgsl,
                 uint8_t p, JumpInfo J, bool wrap_ok, int32_t Sm_prev, uint32_t slm_prev,
                 uint32_t self_idx, uint32_t i, uint32_t m) {
    PassA r;
    uint32_t dext;
    d_layer(sc, up, r.D, r.dl, dext);
    const bool eq = (p == cc.q);
    const int32_t addend = eq ? sc.match : sc.mismatch;
    const int32_t diag = dgS + addend;
    const uint32_t dgl = dgsl + 1;
    int32_t best = MIN_SCORE; uint32_t mv = MV_XCLIP_SUFFIX, len = 0;
    bool is_diag = false;
    if (diag >= best) { best = diag; mv = MV_DIAG; len = dgl; is_diag = true; }
    if (r.D > best) { best = r.D; mv = MV_DEL; len = r.dl; is_diag = false; }
    const int32_t A = best;
    uint32_t is_wrap;
    JumpInfo jp = jump_for_cell(J, addend, wrap_ok, Sm_prev, slm_prev, self_idx, m, is_wrap);
    if (jp.score > best || (jp.score == best && is_diag && jp.len > dgl)) {
        best = jp.score; mv = is_wrap ? MV_WRAP : MV_JUMP; len = jp.len;
    }
    if (cc.xclip_score > best) { best = cc.xclip_score; mv = MV_XCLIP_PREFIX; len = cc.sl0j; }
    const int32_t yclip = sc.yp + sc.o + sc.e * (int32_t)i;
    if (yclip > best) { best = yclip; mv = MV_YCLIP_PREFIX; len = col0_slen(sc, i, m); }
    r.H = best; r.hl = len; r.fl = mv | (dext ? TBB_DEXT : 0u) | (best > A ? PA_GTA : 0u);
    return r;
}

// Candidates of row m (everything except the start value, which is the tracker).
SHD RowM pass_a_rowm(const Scoring &sc, const ColConst &cc, const CellState &up, int32_t dgS, uint32_t dgsl,
                     uint8_t p, JumpInfo J, bool wrap_ok, int32_t Sm_prev, uint32_t slm_prev,
                     uint32_t self_idx, uint32_t m) {
    RowM r;
    d_layer(sc, up, r.D, r.dl, r.dext);
    const bool eq = (p == cc.q);
    const int32_t addend = eq ? sc.match : sc.mismatch;
    r.is_match = eq;
    r.diag = dgS + addend; r.dgl = dgsl + 1;
    uint32_t is_wrap;
    r.jp = jump_for_cell(J, addend, wrap_ok, Sm_prev, slm_prev, self_idx, m, is_wrap);
    if (is_wrap) r.jp.idx |= 0x80000000u;   // remember it is the wrap (for the packed move)
    r.xclip = cc.xclip_score; r.xclip_len = cc.sl0j;
    r.yclip = sc.yp + sc.o + sc.e * (int32_t)m; r.yclip_len = 0;   // filled lazily below
    r.I = MIN_SCORE; r.il = 0; r.iext = 0;
    return r;
}

// S/I merge for a row < m (see header comment); returns the packed move.
struct CellOut { int32_t S; uint32_t sl, mv, idx, from; };   // idx/from only filled by finish_rowm

SHD void merge_si(const PassA &a, int32_t I, uint32_t il, int32_t &S, uint32_t &sl, uint32_t &mv) {
    if (I > a.H || (I == a.H && (a.fl & PA_GTA))) { S = I; sl = il; mv = MV_INS; }
    else { S = a.H; sl = a.hl; mv = a.fl & 15u; }
}

// Next row's I from this row's final S (SCA:317-326).
SHD void i_step(const Scoring &sc, int32_t S, uint32_t sl, int32_t &I, uint32_t &il, uint32_t &iext) {
    int32_t ext = I + sc.e, open = S + sc.o + sc.e;
    if (ext >= open) { I = ext; il = il + 1; iext = 1; }
    else { I = open; il = sl + 1; iext = 0; }
}

// x-suffix tracker partial over rows < m of one contig in one column (SCA:407-429):
// maximal (t, len), FIRST row on full ties.  Start state: (MIN_SCORE, 0) with no row.
struct XsPart { int32_t t; uint32_t len, row; };   // row == 0: nothing beat the start state
SHD void xs_init(XsPart &p) { p.t = MIN_SCORE; p.len = 0; p.row = 0; }
SHD void xs_add(XsPart &p, int32_t t, uint32_t len, uint32_t row) {     // rows arrive in increasing order
    if (t > p.t || (t == p.t && len > p.len)) { p.t = t; p.len = len; p.row = row; }
}
SHD XsPart xs_merge(const XsPart &a, const XsPart &b) {                 // order-independent
    if (a.row == 0) return b;
    if (b.row == 0) return a;
    if (a.t != b.t) return a.t > b.t ? a : b;
    if (a.len != b.len) return a.len > b.len ? a : b;
    return a.row < b.row ? a : b;
}

// Column-best partial for the next column's jump (SCA:677-697): max S, FIRST row on ties.
struct CmPart { int32_t S; uint32_t row, sl; uint32_t valid; };
SHD void cm_init(CmPart &p) { p.S = 0; p.row = 0; p.sl = 0; p.valid = 0; }
SHD void cm_add(CmPart &p, int32_t S, uint32_t sl, uint32_t row) {      // rows in increasing order
    if (!p.valid || S > p.S) { p.S = S; p.row = row; p.sl = sl; p.valid = 1; }
}
SHD CmPart cm_merge(const CmPart &a, const CmPart &b) {
    if (!a.valid) return b;
    if (!b.valid) return a;
    if (a.S != b.S) return a.S > b.S ? a : b;
    return a.row < b.row ? a : b;
}

// y-suffix tracker update of one row (SCA:432-447).  The tie-break compares with the s_len of
// cell (i, n), which is still the zero of Traceback::init whenever this runs.
SHD void sn_update(const Scoring &sc, SnRec &r, int32_t S, uint32_t sl, uint32_t idx, uint32_t j, uint32_t n) {
    int32_t u = S + sc.ys;
    if (u > r.sn || (u == r.sn && sl > 0)) { r.sn = u; r.ly = n - j; r.len = sl; r.idx = idx; }
}

SHD SnRec sn_init(const Scoring &sc, int32_t S0, uint32_t sl0, uint32_t self_idx, uint32_t n) {   // SCA:179-183
    SnRec r; r.sn = MIN_SCORE; r.ly = 0; r.len = 0; r.idx = 0;
    if (S0 + sc.ys > MIN_SCORE) { r.sn = S0 + sc.ys; r.ly = n; r.len = sl0; r.idx = self_idx; }
    return r;
}

// Windowed y-suffix tracking: with G(j) = best cell of column j over all
// contigs, every row i >= 1 ends with Sn(i) - ys >= Gmax - W', W' = max(match, mismatch, 0) - min jump
// score - min(match, mismatch), because every cell of column j+1 can be entered by a jump from the
// best cell of column j.  Columns with G(j) < Gmax - W' therefore never hold the final (Sn, Ly) of any
// row; the trackers only need the columns from the first candidate on (column 0 is their start value).
//
// `best_only` (reads walked with `traceback`, TB:129-150, i.e. only from the best end) with an x-suffix clip
// (xs finite and <= 0: the local mode): the margin is -xs.  The end of the read (SCA:453-555) folds the rows of a
// contig into its row m by "replace if greater, or equal and longer" scans in row order, and `traceback`
// compares the row-m cells of the contigs the same way: the outcome is decided by the rows whose final value is
// the maximum, in the contigs that reach the best final value V.  A row's tracker reaches row m through the
// x-suffix clip (+ xs; the x-jump reads the cell before its y-clip), row m's own tracker directly, the
// insertion re-computation with o + e < 0 on top.  The row of the best cell gives V >= Gmax + ys + xs, so a row
// that decides anything through its tracker has Sn(i) - ys >= Gmax + xs: its tracker is set by cells within -xs
// of the best score of the whole matrix, and those are tracked.  Every other row keeps a tracker that is too
// low (never too high), loses the same scans it loses with the exact value, and is never walked.  (Without an
// x-suffix clip the trackers of the rows below m never reach row m and V has no such bound: full margin.)
SHD int32_t track_margin(const Scoring &sc, bool best_only) {
    if (best_only && sc.xs != MIN_SCORE && sc.xs <= 0) return -sc.xs;
    int32_t submax = sc.match > sc.mismatch ? sc.match : sc.mismatch;
    if (submax < 0) submax = 0;
    const int32_t submin = sc.match < sc.mismatch ? sc.match : sc.mismatch;
    int32_t gmin = sc.g_same < sc.g_opp ? sc.g_same : sc.g_opp;
    gmin = gmin < sc.g_inter ? gmin : sc.g_inter;
    return submax - gmin - submin;
}
SHD uint32_t first_candidate_column(const Scoring &sc, const int32_t *gcol, uint32_t n, bool best_only = false) {
    int32_t gmax = gcol[0];
    for (uint32_t j = 1; j <= n; ++j) gmax = gcol[j] > gmax ? gcol[j] : gmax;
    const int32_t thr = gmax - track_margin(sc, best_only);
    for (uint32_t j = 1; j <= n; ++j) if (gcol[j] >= thr) return j;
    return n + 1;
}

// Finishes row m of a contig in column j given the tracker over rows < m (SCA:350-429 at i == m).
struct RowMOut { CellOut c; uint32_t lx; uint32_t s_tb; };
SHD RowMOut finish_rowm(const Scoring &sc, const RowM &r, const XsPart &tr, uint32_t self_idx, uint32_t m) {
    RowMOut o;
    int32_t best = tr.t; uint32_t mv = MV_XCLIP_SUFFIX, len = tr.len, idx = 0, from = 0;
    bool is_diag = false;
    if (r.diag >= best) { best = r.diag; mv = MV_DIAG; len = r.dgl; idx = self_idx; from = m - 1; is_diag = true; }
    if (r.D > best) { best = r.D; mv = MV_DEL; len = r.dl; idx = self_idx; from = m; is_diag = false; }
    if (r.I > best) { best = r.I; mv = MV_INS; len = r.il; idx = self_idx; from = m - 1; is_diag = false; }
    const bool wrap = (r.jp.idx & 0x80000000u) != 0;
    const uint32_t jidx = r.jp.idx & 0x7fffffffu;
    if (r.jp.score > best || (r.jp.score == best && is_diag && best == r.diag && r.jp.len > r.dgl)) {
        best = r.jp.score; mv = wrap ? MV_WRAP : MV_JUMP; len = r.jp.len; idx = jidx; from = r.jp.from;
    }
    if (r.xclip > best) { best = r.xclip; mv = MV_XCLIP_PREFIX; len = r.xclip_len; idx = self_idx; from = 0; }
    if (r.yclip > best) { best = r.yclip; mv = MV_YCLIP_PREFIX; len = col0_slen(sc, m, m); idx = self_idx; from = m; }
    o.c.S = best; o.c.sl = len; o.c.mv = mv; o.c.idx = idx; o.c.from = from;
    // tracker step at i == m compares the new cell with the stored tracker cell (SCA:407-429)
    uint32_t lx = tr.row ? (m - tr.row) : 0;
    if (sc.xs == 0 && len > tr.len) lx = 0;
    o.lx = lx;
    o.s_tb = (mv == MV_DIAG || mv == MV_JUMP || mv == MV_WRAP) ? (r.is_match ? TB_MATCH : TB_SUBST) : mv;
    return o;
}

// Reference TB code of a packed move.
SHD uint32_t tb_of_move(uint32_t mv, bool is_match) {
    if (mv == MV_DIAG || mv == MV_JUMP || mv == MV_WRAP) return is_match ? TB_MATCH : TB_SUBST;
    return mv;   // MV_START..MV_DEL and the clip codes share the reference's numbering
}

// ---------------------------------------------------------------------------------------------
// Jump selection for one contig of the layout (MCA:279-331).  cm/len/from: per layout position,
// the column best of the previous column (score WITHOUT the jump cost).
// ---------------------------------------------------------------------------------------------
SHD JumpInfo select_jump(const Scoring &sc, const ContigEntry *ent, uint32_t C, uint32_t a,
                         const int32_t *cm, const uint32_t *cml, const uint32_t *cmk) {
    JumpInfo best;
    best.score = cm[a] + sc.g_same; best.len = cml[a] + 1; best.idx = ent[a].contig_idx; best.from = cmk[a];
    const int32_t opp = ent[a].opp;
    if (opp >= 0) {
        int32_t s = cm[opp] + sc.g_opp;
        if (s > best.score) { best.score = s; best.len = cml[opp] + 1; best.idx = ent[opp].contig_idx; best.from = cmk[opp]; }
    }
    bool have = false; int32_t is = 0; uint32_t il = 0, ib = 0;
    for (uint32_t b = 0; b < C; ++b) {
        if (b == a || (int32_t)b == opp) continue;
        int32_t s = cm[b] + sc.g_inter; uint32_t l = cml[b] + 1;
        if (!have || s > is || (s == is && l >= il)) { have = true; is = s; il = l; ib = b; }   // last max wins
    }
    if (have && is > best.score) { best.score = is; best.len = il; best.idx = ent[ib].contig_idx; best.from = cmk[ib]; }
    return best;
}


// =============================================================================================
// Lane-level passes over one strip of STRIP consecutive rows of one contig in one column.
// Index math: padded cell p = tile*TILE + r (r = row-1 within the contig's tiles).  The rolling
// state is stored tile-transposed (k*32 + lane) so that a warp's loads are 512 contiguous bytes.
// =============================================================================================
SHD uint32_t state_index(uint32_t tile, uint32_t lane, uint32_t k) { return tile * TILE + k * 32 + lane; }
SHD uint32_t cell_index(uint32_t tile, uint32_t lane, uint32_t k) { return tile * TILE + lane * STRIP + k; }
// row i (1-based) of a contig -> tile-transposed index (rolling state, SnRec, LastCell)
SHD uint32_t row_index(const ContigEntry &en, uint32_t i) {
    const uint32_t r = i - 1;
    return state_index(en.tile_start + r / TILE, (r % TILE) / STRIP, r % STRIP);
}
// row i -> linear index (packed traceback bytes)
SHD uint32_t row_linear(const ContigEntry &en, uint32_t i) { return en.tile_start * TILE + i - 1; }

struct TileCtx {          // per tile, per column (uniform over the warp)
    uint32_t a;           // layout position of the contig
    uint32_t self_idx;    // its contig index
    uint32_t m;
    uint32_t tile;        // global tile index
    uint32_t tile_in_contig;
    JumpInfo J;           // best jump into this contig for this column (MCA:319-330)
    bool circular;
    bool wrap_src_ok;     // s_tb(m, j-1) != XCLIP_SUFFIX   (SCA:263-267)
    int32_t Sm_prev; uint32_t slm_prev;   // S / s_len of (m, j-1)
};

struct LaneA {
    PassA a[STRIP];
    ICarry agg;           // I candidate leaving this strip (arrives at the row after it)
    uint32_t has_m;       // row m is in this strip (its candidates went to *rowm)
};

// Pass A of one lane.  `up[k]` = state of row (row0+k) at column j-1; (dgS, dgsl) = S, s_len of
// row (row0-1) at column j-1.  Row m's candidates are written to *rowm (shared memory on the GPU).
SHD void lane_pass_a(const Scoring &sc, const ColConst &cc, const TileCtx &tc, uint32_t row0,
                     const CellState *up, int32_t dgS, uint32_t dgsl, const uint8_t *x, LaneA &out, RowM *rowm) {
    out.has_m = 0;
    out.agg.v = MIN_SCORE; out.agg.il = 0; out.agg.open = 0;
    STITCH_UNROLL
    for (int k = 0; k < STRIP; ++k) {
        const uint32_t i = row0 + (uint32_t)k;
        const bool wrap_ok = tc.circular && i == 1 && tc.wrap_src_ok;
        if (i < tc.m) {
            PassA pa = pass_a(sc, cc, up[k], dgS, dgsl, x[k], tc.J, wrap_ok, tc.Sm_prev, tc.slm_prev, tc.self_idx, i, tc.m);
            out.a[k] = pa;
            const int32_t open = pa.H + sc.o + sc.e;
            if (k == 0) { out.agg.v = open; out.agg.il = pa.hl + 1; out.agg.open = 1; }
            else {
                const int32_t ext = out.agg.v + sc.e;
                if (ext >= open) { out.agg.v = ext; out.agg.il += 1; out.agg.open = 0; }
                else { out.agg.v = open; out.agg.il = pa.hl + 1; out.agg.open = 1; }
            }
        } else if (i == tc.m) {
            *rowm = pass_a_rowm(sc, cc, up[k], dgS, dgsl, x[k], tc.J, wrap_ok, tc.Sm_prev, tc.slm_prev, tc.self_idx, tc.m);
            out.has_m = 1;
        }
        dgS = up[k].S; dgsl = up[k].sl;
    }
}

struct LaneB {
    XsPart xs;            // x-suffix tracker partial over this strip's rows < m
    CmPart cm;            // column-best partial over this strip's rows < m
};

// Pass B of one lane.  `cin` = I arriving at row0.  Writes the new rolling state, the packed
// traceback bytes and (optionally) the y-suffix trackers / column-n records.
SHD void lane_pass_b(const Scoring &sc, const ColConst &cc, const TileCtx &tc, uint32_t row0, uint32_t lane,
                     const LaneA &la, ICarry cin, CellState *state_curr, CellState *ck_state, uint8_t *tb_col,
                     bool track, SnRec *sn, bool lastcol, LastCell *last, const uint8_t *x, LaneB &out, RowM *rowm) {
    xs_init(out.xs); cm_init(out.cm);
    int32_t I = cin.v; uint32_t il = cin.il; uint32_t iext = cin.open ? 0u : 1u;
    STITCH_UNROLL
    for (int k = 0; k < STRIP; ++k) {
        const uint32_t i = row0 + (uint32_t)k;
        if (i == tc.m) { rowm->I = I; rowm->il = il; rowm->iext = iext; }
        if (i < tc.m) {
            const PassA &pa = la.a[k];
            int32_t S; uint32_t sl, mv;
            merge_si(pa, I, il, S, sl, mv);
            CellState st; st.S = S; st.D = pa.D; st.sl = sl; st.dl = pa.dl;
            const uint32_t si = state_index(tc.tile, lane, (uint32_t)k);
            state_curr[si] = st;
            if (ck_state) ck_state[si] = st;
            if (tb_col) tb_col[cell_index(tc.tile, lane, (uint32_t)k)] = (uint8_t)(mv | (pa.fl & TBB_DEXT) | (iext ? TBB_IEXT : 0u));
            xs_add(out.xs, S + sc.xs, sl, i);
            cm_add(out.cm, S, sl, i);
            if (track || lastcol) {
                uint32_t idx, from;
                ptr_of_move(mv, tc.self_idx, i, tc.m, tc.J, idx, from);
                if (track) sn_update(sc, sn[si], S, sl, idx, cc.j, cc.n);
                if (lastcol) {
                    LastCell lc; lc.S = S; lc.I = I; lc.sl = sl; lc.il = il; lc.idx = idx; lc.from = from;
                    lc.s_tb = (uint8_t)tb_of_move(mv, x[k] == cc.q); lc.i_tb = 0;
                    lc.flags = (uint8_t)((iext ? 1 : 0) | ((pa.fl & TBB_DEXT) ? 2 : 0)); lc.pad = 0; lc.pad2 = 0;
                    last[si] = lc;
                }
            }
            i_step(sc, S, sl, I, il, iext);
        }
    }
}

// Per-contig end of column: finishes row m, returns the column best (for the next jump).
struct ContigColOut { CmPart cm; int32_t Sm; uint32_t slm; uint32_t s_tb_m; };
SHD ContigColOut contig_finalize(const Scoring &sc, const ColConst &cc, const ContigEntry &en, uint32_t a, uint32_t C,
                                 const RowM &rm, XsPart xs, CmPart cm_rows, const Row0 &r0, JumpInfo J,
                                 CellState *state_curr, CellState *ck_state, uint8_t *tb_col, ColRec *colrec_col,
                                 bool track, SnRec *sn, bool lastcol, LastCell *last) {
    (void)C;
    RowMOut ro = finish_rowm(sc, rm, xs, en.contig_idx, en.m);
    const uint32_t r = en.m - 1;
    const uint32_t tile = en.tile_start + r / TILE, lane = (r % TILE) / STRIP, k = r % STRIP;
    CellState st; st.S = ro.c.S; st.D = rm.D; st.sl = ro.c.sl; st.dl = rm.dl;
    const uint32_t p = state_index(tile, lane, k);
    state_curr[p] = st;
    if (ck_state) ck_state[p] = st;
    if (tb_col) tb_col[cell_index(tile, lane, k)] = (uint8_t)(ro.c.mv | (rm.dext ? TBB_DEXT : 0u) | (rm.iext ? TBB_IEXT : 0u));
    if (colrec_col) {
        ColRec cr; cr.jscore = J.score; cr.jlen = J.len; cr.jidx = J.idx; cr.jfrom = J.from; cr.lx = ro.lx;
        cr.pad0 = cr.pad1 = cr.pad2 = 0;
        colrec_col[a] = cr;
    }
    if (track) sn_update(sc, sn[p], ro.c.S, ro.c.sl, ro.c.idx, cc.j, cc.n);
    if (lastcol) {
        LastCell lc; lc.S = ro.c.S; lc.I = rm.I; lc.sl = ro.c.sl; lc.il = rm.il; lc.idx = ro.c.idx; lc.from = ro.c.from;
        lc.s_tb = (uint8_t)ro.s_tb; lc.i_tb = 0; lc.flags = (uint8_t)((rm.iext ? 1 : 0) | (rm.dext ? 2 : 0)); lc.pad = 0; lc.pad2 = 0;
        last[p] = lc;
    }
    // column best over rows 0..m, first row on ties (SCA:680-687)
    CmPart cm; cm_init(cm);
    cm_add(cm, r0.S, r0.sl, 0);
    cm = cm_merge(cm, cm_rows);
    CmPart top; top.S = ro.c.S; top.row = en.m; top.sl = ro.c.sl; top.valid = 1;
    cm = cm_merge(cm, top);
    ContigColOut o; o.cm = cm; o.Sm = ro.c.S; o.slm = ro.c.sl; o.s_tb_m = ro.s_tb;
    return o;
}

// =============================================================================================
// End-of-read fix-up of one contig (SCA:453-555) on the column-n records.
// =============================================================================================
// `track` false: the y-suffix trackers were not kept (ys == MIN_SCORE: Sn = S + MIN can never
// exceed S(i, n) inside the score range the API accepts), so Sn reads as MIN_SCORE.
SHD void fixup_contig(const Scoring &sc, const ContigEntry &en, uint32_t n, LastCell *last /* base of the read */,
                      const SnRec *sn, bool track, uint32_t *lx_n) {
    const uint32_t m = en.m;
    #define ROWIDX(i) row_index(en, (i))
    const Row0 r0 = row0_at(sc, n, n);
    LastCell row0c; row0c.S = r0.S; row0c.I = MIN_SCORE; row0c.sl = r0.sl; row0c.il = 0; row0c.idx = en.contig_idx;
    row0c.from = 0; row0c.s_tb = (uint8_t)r0.s_tb; row0c.i_tb = TB_START; row0c.flags = 0; row0c.pad = 0; row0c.pad2 = 0;
    // resolve the open I pointers against the S moves as they were when the column was filled
    {
        uint8_t prev_tb = row0c.s_tb;
        for (uint32_t i = 1; i <= m; ++i) {
            LastCell &c = last[ROWIDX(i)];
            c.i_tb = (c.flags & 1) ? (uint8_t)TB_INS : prev_tb;
            prev_tb = c.s_tb;
        }
    }
    LastCell &cm_ = last[ROWIDX(m)];
    // One row of SCA:455-517.  Row 0 lives in a local record, rows >= 1 in global memory; the two
    // cases are kept as separate calls (a pointer select between the address spaces made nvcc
    // 12.9 drop the `from` store of the selected record).
    auto row_step = [&](LastCell &c, const SnRec &s, uint32_t i) {
        if (c.S + sc.g_same > cm_.S) {                                   // SCA:460-466
            cm_.S = c.S + sc.g_same;
            const uint32_t l = c.sl, ix = c.idx;
            cm_.s_tb = TB_XJUMP; cm_.sl = l; cm_.idx = ix; cm_.from = i;
        }
        if (s.sn > c.S) {                                                // SCA:469-491
            c.S = s.sn;
            c.s_tb = TB_YCLIP_SUFFIX; c.sl = s.len; c.idx = s.idx; c.from = i;
        }
        {                                                                // SCA:494-516
            const int32_t t = c.S + sc.xs;
            if (t > cm_.S || (t == cm_.S && c.sl > cm_.sl)) {
                cm_.S = t; *lx_n = m - i;
                const uint32_t l = c.sl, ix = c.idx;
                cm_.s_tb = TB_XCLIP_SUFFIX; cm_.sl = l; cm_.idx = ix; cm_.from = i;
            }
        }
    };
    {
        SnRec s0; s0.sn = sc.ys; s0.ly = n; s0.len = 0; s0.idx = en.contig_idx;
        row_step(row0c, s0, 0);
    }
    for (uint32_t i = 1; i <= m; ++i) {
        SnRec s;
        if (track) s = sn[ROWIDX(i)];
        else { s.sn = MIN_SCORE; s.len = 0; s.ly = 0; s.idx = 0; }
        row_step(last[ROWIDX(i)], s, i);
    }
    for (uint32_t i = 1; i <= m; ++i) {                                  // SCA:521-554
        int32_t pS; uint32_t psl; uint8_t ptb;
        if (i == 1) { pS = row0c.S; psl = row0c.sl; ptb = row0c.s_tb; }
        else { const LastCell &pc = last[ROWIDX(i - 1)]; pS = pc.S; psl = pc.sl; ptb = pc.s_tb; }
        LastCell &c = last[ROWIDX(i)];
        const int32_t is = pS + sc.o + sc.e;
        if (is > c.I) { c.I = is; c.i_tb = ptb; c.il = psl + 1; }
        if (is > c.S) {
            c.S = is;
            const uint32_t pl = c.il;
            c.s_tb = TB_INS; c.sl = pl; c.idx = en.contig_idx; c.from = i - 1;
            if (c.S + sc.xs > cm_.S) {
                cm_.S = c.S + sc.xs; *lx_n = m - i;
                cm_.s_tb = TB_XCLIP_SUFFIX; cm_.sl = pl; cm_.idx = en.contig_idx; cm_.from = i;
            }
        }
    }
    #undef ROWIDX
}

// =============================================================================================
// Traceback walk (TB:219-373) over a virtual reference-cell view of the packed stores.
// =============================================================================================
struct OutOp { uint32_t kind, a, b; };   // same layout as stitch_op
enum : uint32_t { OP_MATCH = 0, OP_SUBST = 1, OP_DEL = 2, OP_INS = 3, OP_XCLIP = 4, OP_YCLIP = 5, OP_XJUMP = 6, OP_YJUMP = 7 };

struct ChainHdr {
    int32_t score;
    uint32_t xstart, xend, ystart, yend, xlen, ylen, start_contig_idx, end_contig_idx, length, n_ops, status;
};
enum : uint32_t { WALK_OK = 0, WALK_NONE = 1, WALK_OVERFLOW = 2, WALK_PANIC = 3 };

// The packed traceback bytes of ONE contig over one block of columns (jb, je], re-filled from
// a column-state checkpoint when the walk needs them (checkpoint-and-recompute).
//
// A unit is re-filled for the cell (i_hi, je) the walk enters it at.  The walk only moves up and left from there, and a
// cell depends only on cells above / left of it within `slope` rows per column (one row for the diagonal plus the reach of
// the insertion chain, dp_packed.h), so a packed re-fill may restrict itself to the CONE of the entry cell: rows
// [i_hi - slope * (je - j), i_hi] at column j.  Cells outside the cone hold stale values; the walk never reads them
// (has() says "not loaded" for them and the unit is re-filled for the new entry cell).  Full units: i_hi = 0xffffffff.
struct TbUnit {
    const uint8_t *bytes;   // column j (jb < j <= je), row i (1..m): bytes[(j - jb - 1) * pm + i - 1]
    const ColRec *cr;       // per column of the unit: cr[j - jb - 1].lx = Lx[j] of this contig (SCA:407-429); full units only
    uint32_t a;             // layout position; 0xffffffff = nothing loaded
    uint32_t jb, je, pm;    // je: the last column filled (the column of the entry cell)
    uint32_t i_hi, slope;   // cone (see above)
    SHD bool has(uint32_t a_, uint32_t i, uint32_t j) const {
        if (a != a_ || j <= jb || j > je) return false;
        if (i_hi == 0xffffffffu) return true;
        return i <= i_hi && (int64_t)i >= (int64_t)i_hi - (int64_t)slope * (int64_t)(je - j);
    }
    SHD uint32_t at(uint32_t i, uint32_t j) const { return bytes[(uint64_t)(j - jb - 1) * pm + (i - 1)]; }
};

struct ReadView {
    Scoring sc;
    const ContigEntry *ent; uint32_t C; uint32_t n;
    const ColRec *colrec;       // [(j)*C + a], j = 0..n: jump sources; lx only for j = n
    const LastCell *last;       // column n, tile-transposed order
    const SnRec *sn;            // tile-transposed order
    const uint8_t *contig_bases;
    const uint8_t *read;
    TbUnit unit;

    // layout position of contig-strand `idx`, or -1 (the layout's contigs are in ascending contig_idx order, MCA:178-223)
    SHD int32_t pos_of(uint32_t idx) const {
        uint32_t lo = 0, hi = C;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ent[mid].contig_idx < idx) lo = mid + 1; else hi = mid; }
        return lo < C && ent[lo].contig_idx == idx ? (int32_t)lo : -1;
    }

    SHD uint32_t pidx(uint32_t a, uint32_t i) const { return row_index(ent[a], i); }
    SHD bool is_match(uint32_t a, uint32_t i, uint32_t j) const {
        return contig_bases[ent[a].seq_off + i - 1] == read[j - 1];
    }
    SHD bool interior(uint32_t i, uint32_t j) const { return i >= 1 && j >= 1 && j < n; }
    // S move of a cell as a reference TB code (interior cells need the loaded unit)
    SHD uint32_t s_tb(uint32_t a, uint32_t i, uint32_t j) const {
        if (i == 0) return row0_at(sc, j, n).s_tb;
        if (j == 0) return col0_at(sc, i, ent[a].m).s_tb;
        if (j == n) return last[pidx(a, i)].s_tb;
        return tb_of_move(unit.at(i, j) & 15u, is_match(a, i, j));
    }
    // (idx, from) of the S pointer of a MATCH/SUBST cell
    SHD void s_ptr(uint32_t a, uint32_t i, uint32_t j, uint32_t &idx, uint32_t &from) const {
        if (j == n) { const LastCell &c = last[pidx(a, i)]; idx = c.idx; from = c.from; return; }
        const uint32_t mv = unit.at(i, j) & 15u;
        if (mv == MV_JUMP) { const ColRec &r = colrec[(uint64_t)j * C + a]; idx = r.jidx; from = r.jfrom; }
        else if (mv == MV_WRAP) { idx = ent[a].contig_idx; from = ent[a].m; }
        else { idx = ent[a].contig_idx; from = i - 1; }
    }
    SHD uint32_t lx(uint32_t a, uint32_t j) const {
        if (j == 0) { int32_t t; uint32_t l; col0_tracker(sc, ent[a].m, t, l); return l; }
        if (j == n) return colrec[(uint64_t)j * C + a].lx;
        return unit.cr[j - unit.jb - 1].lx;   // interior columns: from the re-filled unit
    }
    SHD uint32_t ly(uint32_t a, uint32_t i) const { return i == 0 ? n : sn[pidx(a, i)].ly; }
};

struct OpWriter {
    OutOp *ops; uint32_t cap, n; bool overflow;
    uint32_t first_kind;   // kind of the first op pushed (operations.first() in the reference)
    uint32_t n_pushed;     // un-encoded count
    bool only_special;
    SHD void init(OutOp *o, uint32_t c) { ops = o; cap = c; n = 0; overflow = false; first_kind = 0xff; n_pushed = 0; only_special = true; }
    SHD void push(uint32_t kind, uint32_t a, uint32_t b) {
        if (n_pushed == 0) first_kind = kind;
        ++n_pushed;
        if (kind <= OP_INS) {
            only_special = false;
            if (n > 0 && ops[n - 1].kind == kind) { ops[n - 1].a += 1; return; }
            a = 1; b = 0;
        }
        if (n >= cap) { overflow = true; return; }
        ops[n].kind = kind; ops[n].a = a; ops[n].b = b; ++n;
    }
};

// Resumable walk (TB:219-373): runs until the chain is complete or the packed bytes of a
// (contig, column block) that is not loaded are needed (WALK_NEED_UNIT: load the unit holding
// column st.j of layout position st.a and call walk_run again).
constexpr uint32_t WL_LOOKUP = 0xffu;   // "S move of the current cell", resolved lazily
struct WalkState {
    uint32_t a, i, j, layer, cur_idx;
    uint32_t xstart, ystart, yend, xend;
    OpWriter w;
};
enum : uint32_t { WALK_NEED_UNIT = 4 };

SHD void walk_begin(const ReadView &v, uint32_t a_end, OutOp *ops, uint32_t cap, WalkState &st, ChainHdr &h) {
    st.w.init(ops, cap);
    st.a = a_end; st.j = v.n; st.i = v.ent[a_end].m;
    st.xstart = 0; st.ystart = 0; st.yend = v.n; st.xend = v.ent[a_end].m;
    const LastCell &endc = v.last[v.pidx(a_end, st.i)];
    h.score = endc.S; h.length = endc.sl; h.end_contig_idx = v.ent[a_end].contig_idx; h.xlen = v.ent[a_end].m; h.ylen = v.n;
    st.cur_idx = v.ent[a_end].contig_idx;
    st.layer = WL_LOOKUP;
    h.status = WALK_OK;
}

SHD uint32_t walk_run(const ReadView &v, WalkState &st, ChainHdr &h) {
    const uint32_t n = v.n;
    OpWriter &w = st.w;
    uint32_t a = st.a, i = st.i, j = st.j, layer = st.layer, cur_idx = st.cur_idx;
    uint32_t status = WALK_OK;
    for (;;) {
        if (w.overflow) { status = WALK_OVERFLOW; break; }
        if (v.interior(i, j) && !v.unit.has(a, i, j) &&
            (layer == WL_LOOKUP || layer == TB_INS || layer == TB_DEL || layer == TB_MATCH || layer == TB_SUBST ||
             layer == TB_XCLIP_SUFFIX)) {
            st.a = a; st.i = i; st.j = j; st.layer = layer; st.cur_idx = cur_idx;
            return WALK_NEED_UNIT;
        }
        if (layer == WL_LOOKUP) { layer = v.s_tb(a, i, j); continue; }
        if (layer == TB_START) break;
        if (layer == TB_INS) {
            w.push(OP_INS, 0, 0);
            uint32_t next;
            if (i == 0) next = TB_START;
            else if (j == 0) next = col0_at(v.sc, i, v.ent[a].m).i_tb;
            else if (j == n) next = v.last[v.pidx(a, i)].i_tb;
            else next = (v.unit.at(i, j) & TBB_IEXT) ? (uint32_t)TB_INS : WL_LOOKUP;
            if (i == 0) { status = WALK_PANIC; break; }
            i -= 1; layer = next;
        } else if (layer == TB_DEL) {
            w.push(OP_DEL, 0, 0);
            uint32_t next;
            if (i == 0) next = row0_at(v.sc, j, n).d_tb;
            else if (j == 0) next = TB_START;
            else if (j == n) next = (v.last[v.pidx(a, i)].flags & 2) ? (uint32_t)TB_DEL : WL_LOOKUP;
            else next = (v.unit.at(i, j) & TBB_DEXT) ? (uint32_t)TB_DEL : WL_LOOKUP;
            if (j == 0) { status = WALK_PANIC; break; }
            j -= 1; layer = next;
        } else if (layer == TB_MATCH || layer == TB_SUBST) {
            w.push(layer == TB_MATCH ? OP_MATCH : OP_SUBST, 0, 0);
            if (i == 0 || j == 0) { status = WALK_PANIC; break; }
            uint32_t sidx, sfrom;
            v.s_ptr(a, i, j, sidx, sfrom);
            if (sidx != cur_idx || sfrom != i - 1) {
                w.push(OP_XJUMP, cur_idx, i - 1);
                cur_idx = sidx;
                const int32_t na = v.pos_of(sidx);
                if (na < 0) { status = WALK_NONE; break; }
                a = (uint32_t)na;
            }
            i = sfrom; j -= 1;
            if (i > v.ent[a].m) { status = WALK_PANIC; break; }
            layer = WL_LOOKUP;
        } else if (layer == TB_XCLIP_PREFIX) {
            const uint32_t next = v.s_tb(a, 0, j);
            if (next == TB_START || next == TB_YCLIP_PREFIX) { w.push(OP_XCLIP, i, 0); st.xstart = i; }
            i = 0; layer = next;
        } else if (layer == TB_XCLIP_SUFFIX) {
            const uint32_t l = v.lx(a, j);
            if (w.n_pushed == 0 || w.first_kind == OP_YCLIP) { w.push(OP_XCLIP, l, 0); st.xend = i - l; }
            if (l > i) { status = WALK_PANIC; break; }
            if (l == 0) { status = WALK_PANIC; break; }   // the reference would spin forever on the same cell
            i -= l; layer = WL_LOOKUP;
        } else if (layer == TB_YCLIP_PREFIX) {
            w.push(OP_YCLIP, j, 0);
            st.ystart = j; j = 0; layer = WL_LOOKUP;
        } else if (layer == TB_YCLIP_SUFFIX) {
            const uint32_t l = v.ly(a, i);
            w.push(OP_YCLIP, l, 0);
            const uint32_t sfrom = (i == 0) ? 0u : v.last[v.pidx(a, i)].from;   // only column n holds this move
            if (l > j) { status = WALK_PANIC; break; }
            j -= l;
            if (sfrom != i) { w.push(OP_XJUMP, cur_idx, i); i = sfrom; }
            st.yend = j; layer = WL_LOOKUP;
        } else if (layer == TB_XJUMP) {
            const LastCell &c = v.last[v.pidx(a, i)];                              // only (m, n) holds this move
            w.push(OP_XJUMP, cur_idx, i);
            cur_idx = c.idx;
            const int32_t na = v.pos_of(c.idx);
            if (na < 0) { status = WALK_NONE; break; }
            a = (uint32_t)na;
            i = c.from;
            if (i > v.ent[a].m) { status = WALK_PANIC; break; }
            layer = WL_LOOKUP;
        } else { status = WALK_PANIC; break; }
    }
    if (w.overflow && status == WALK_OK) status = WALK_OVERFLOW;
    for (uint32_t l = 0; l < w.n / 2; ++l) {
        OutOp t = w.ops[l]; w.ops[l] = w.ops[w.n - 1 - l]; w.ops[w.n - 1 - l] = t;
    }
    if (w.only_special) { st.xstart = 0; st.xend = 0; st.ystart = 0; st.yend = 0; }
    h.xstart = st.xstart; h.xend = st.xend; h.ystart = st.ystart; h.yend = st.yend;
    h.start_contig_idx = cur_idx; h.n_ops = w.n; h.status = status;
    return status;
}

// End contig of the best chain (TB:129-150): max S[m_c], then longer length, else first.
// `skip` (optional, by layout position): contigs to leave out (traceback_all's seen / not-considered).
SHD uint32_t pick_end(const ReadView &v, const uint8_t *skip) {
    uint32_t off = 0; int32_t score = MIN_SCORE; uint32_t alen = 0;
    for (uint32_t a = 0; a < v.C; ++a) {
        if (skip && skip[a]) continue;
        const LastCell &c = v.last[v.pidx(a, v.ent[a].m)];
        if (c.S > score || (c.S == score && c.sl > alen)) { off = a; score = c.S; alen = c.sl; }
    }
    return off;
}

}  // namespace stitch
