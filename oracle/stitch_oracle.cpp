// TEST INFRASTRUCTURE ONLY — see stitch_oracle.hpp.  CPU restatement of the reference aligner.
#include "stitch_oracle.hpp"

#include <algorithm>
#include <cassert>
#include <stdexcept>

namespace oracle {

static inline void panic(const char *what) { throw std::runtime_error(what); }

// ---------------------------------------------------------------------------------------------
// Scoring / ops
// ---------------------------------------------------------------------------------------------
void Scoring::set_clips_for(Mode m) {   // mod.rs:123-131
    switch (m) {
    case LOCAL: xclip_prefix = xclip_suffix = yclip_prefix = yclip_suffix = 0; break;
    case QUERY_LOCAL: xclip_prefix = xclip_suffix = MIN_SCORE; yclip_prefix = yclip_suffix = 0; break;
    case TARGET_LOCAL: xclip_prefix = xclip_suffix = 0; yclip_prefix = yclip_suffix = MIN_SCORE; break;
    case GLOBAL: xclip_prefix = xclip_suffix = yclip_prefix = yclip_suffix = MIN_SCORE; break;
    default: panic("Custom alignment mode not supported");
    }
}

int64_t Op::len_x(int64_t x_index) const {   // constants.rs:61-72
    switch (kind) {
    case MATCH: case SUBST: case INS: return 1;
    case DEL: case YCLIP: case YJUMP: return 0;
    case XCLIP: return (int64_t)a;
    default: return (int64_t)b - x_index;   // Xjump(_, to)
    }
}
int64_t Op::len_y() const {   // constants.rs:75-84
    switch (kind) {
    case MATCH: case SUBST: case DEL: return 1;
    case YCLIP: case YJUMP: return (int64_t)a;
    default: return 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Cell: w0 = s_tb[0:4) s_len[4:31) i_tb[31:35) i_len[35:62) idx_hi[62:64)
//       w1 = d_tb[0:4) d_len[4:31) from[31:58) idx_lo[58:64)
// ---------------------------------------------------------------------------------------------
static constexpr uint64_t L27 = (1ull << 27) - 1;
void Cell::set_s(uint8_t tb, uint32_t len) {
    if (tb > TB_XJUMP) panic("tb > TB_MAX");
    w0 = (w0 & ~((0xFull) | (L27 << 4))) | (uint64_t)tb | (((uint64_t)len & L27) << 4);
}
void Cell::set_i(uint8_t tb, uint32_t len) {
    if (tb > TB_XJUMP) panic("tb > TB_MAX");
    w0 = (w0 & ~((0xFull << 31) | (L27 << 35))) | ((uint64_t)tb << 31) | (((uint64_t)len & L27) << 35);
}
void Cell::set_d(uint8_t tb, uint32_t len) {
    if (tb > TB_XJUMP) panic("tb > TB_MAX");
    w1 = (w1 & ~((0xFull) | (L27 << 4))) | (uint64_t)tb | (((uint64_t)len & L27) << 4);
}
void Cell::set_s_all(uint8_t tb, uint32_t len, uint32_t idx, uint32_t from) {
    if (idx > 255) panic("idx > max_num_contigs");          // packed_length_cell.rs:139
    if (from > 134217727u) panic("from > max_target_len");  // packed_length_cell.rs:140
    set_s(tb, len);
    w0 = (w0 & ~(3ull << 62)) | ((uint64_t)(idx >> 6) << 62);
    w1 = (w1 & ~((L27 << 31) | (0x3Full << 58))) | ((uint64_t)from << 31) | ((uint64_t)(idx & 63) << 58);
}
uint8_t Cell::s_tb() const { return (uint8_t)(w0 & 0xF); }
uint32_t Cell::s_len() const { return (uint32_t)((w0 >> 4) & L27); }
uint8_t Cell::i_tb() const { return (uint8_t)((w0 >> 31) & 0xF); }
uint32_t Cell::i_len() const { return (uint32_t)((w0 >> 35) & L27); }
uint8_t Cell::d_tb() const { return (uint8_t)(w1 & 0xF); }
uint32_t Cell::d_len() const { return (uint32_t)((w1 >> 4) & L27); }
uint32_t Cell::idx() const { return (uint32_t)(((w0 >> 62) << 6) | (w1 >> 58)); }
uint32_t Cell::from() const { return (uint32_t)((w1 >> 31) & L27); }

// ---------------------------------------------------------------------------------------------
// SingleContig
// ---------------------------------------------------------------------------------------------
bool SingleContig::g_checker_layout = false;

void SingleContig::init_matrices(int64_t m, int64_t n) {   // SCA:97-186
    // traceback.init: every cell = START / len 0 / idx 0 / from 0 (traceback/mod.rs:93-100)
    rows = m + 1;
    cols = n + 1;
    tb.assign((size_t)(rows * cols), Cell{});

    for (int k = 0; k < 2; ++k) {
        I[k].assign((size_t)m + 1, MIN_SCORE);
        D[k].assign((size_t)m + 1, MIN_SCORE);
        S[k].assign((size_t)m + 1, MIN_SCORE);
        S[k][0] = 0;

        if (k == 0) {
            Cell c;
            c.set_all(TB_START, 0);
            c.set_s_all(TB_START, 0, contig_idx, 0);
            cell(0, 0) = c;
            Lx.assign((size_t)n + 1, 0);
            Ly.assign((size_t)m + 1, 0);
            Sn.assign((size_t)m + 1, MIN_SCORE);
            Sn[0] = sc.yclip_suffix;
            Ly[0] = n;
        }

        for (int64_t i = 1; i <= m; ++i) {
            Cell c;
            c.set_all(TB_START, 0);
            c.set_s_all(TB_START, 0, contig_idx, 0);
            if (i == 1) {
                I[k][i] = sc.gap_open + sc.gap_extend;
                c.set_i(TB_START, 1);
            } else {
                // one long insertion, or x-prefix clip followed by a fresh insertion (clip wins ties)
                int32_t i_score = sc.gap_open + sc.gap_extend * (int32_t)i;
                int32_t c_score = sc.xclip_prefix + sc.gap_open + sc.gap_extend;
                if (i_score > c_score) {
                    I[k][i] = i_score;
                    c.set_i(TB_INS, (uint32_t)i);
                } else {
                    I[k][i] = c_score;
                    c.set_i(TB_XCLIP_PREFIX, 0);
                }
            }

            if (i == m) {
                c.set_s(TB_XCLIP_SUFFIX, 0);   // S[k][m] keeps the tracker value
            } else {
                S[k][i] = MIN_SCORE;
            }
            if (I[k][i] > S[k][i]) {
                S[k][i] = I[k][i];
                c.set_s(TB_INS, (uint32_t)i);
            }
            if (sc.xclip_prefix > S[k][i]) {
                S[k][i] = sc.xclip_prefix;
                c.set_s(TB_XCLIP_PREFIX, 0);
            }
            // x-suffix tracker
            if (i != m && S[k][i] + sc.xclip_suffix > S[k][m]) {
                S[k][m] = S[k][i] + sc.xclip_suffix;
                Lx[0] = m - i;
            }
            if (k == 0) cell(i, 0) = c;
            // y-suffix tracker
            if (S[k][i] + sc.yclip_suffix > Sn[i]) {
                Sn[i] = S[k][i] + sc.yclip_suffix;
                Ly[i] = n;
            }
        }
    }
}

void SingleContig::init_column(int64_t j, int curr, int64_t m, int64_t n) {   // SCA:188-239
    Cell c;
    c.set_s_all(TB_START, 0, contig_idx, 0);
    I[curr][0] = MIN_SCORE;
    if (j == 1) {
        D[curr][0] = sc.gap_open + sc.gap_extend;
        c.set_d(TB_START, 1);
    } else {
        int32_t d_score = sc.gap_open + sc.gap_extend * (int32_t)j;
        int32_t c_score = sc.yclip_prefix + sc.gap_open + sc.gap_extend;
        if (d_score > c_score) {
            D[curr][0] = d_score;
            c.set_d(TB_DEL, (uint32_t)j);
        } else {
            D[curr][0] = c_score;
            c.set_d(TB_YCLIP_PREFIX, 0);
        }
    }
    if (D[curr][0] > sc.yclip_prefix) {
        S[curr][0] = D[curr][0];
        c.set_s(TB_DEL, (uint32_t)j);
    } else {
        S[curr][0] = sc.yclip_prefix;
        c.set_s(TB_YCLIP_PREFIX, 0);
    }
    if (j == n && Sn[0] > S[curr][0]) {
        S[curr][0] = Sn[0];
        c.set_s(TB_YCLIP_SUFFIX, 0);
    } else if (S[curr][0] + sc.yclip_suffix > Sn[0]) {
        Sn[0] = S[curr][0] + sc.yclip_suffix;
        Ly[0] = n - j;
    }
    cell(0, j) = c;
    for (int64_t i = 1; i <= m; ++i) S[curr][i] = MIN_SCORE;
}

JumpInfo SingleContig::jump_info(int64_t m, int64_t j, int32_t jump_score) const {   // SCA:677-697
    const std::vector<int32_t> &col = S[j % 2];
    int32_t best = col[0] + jump_score;
    int64_t from = 0;
    for (int64_t k = 1; k <= m; ++k) {
        int32_t v = col[k] + jump_score;
        if (best < v) { best = v; from = k; }
    }
    return JumpInfo{best, cell(from, j).s_len() + 1, contig_idx, (uint32_t)from};
}

void SingleContig::fill_column(const uint8_t *x, const uint8_t *y, int64_t m, int64_t n, int64_t j,
                               int prev, int curr, JumpInfo jump) {   // SCA:292-451
    const uint8_t q = y[j - 1];
    const int32_t o = sc.gap_open, e = sc.gap_extend;
    const int32_t xclip_score = sc.xclip_prefix + std::max(sc.yclip_prefix, o + e * (int32_t)j);

    for (int64_t i = 1; i <= m; ++i) {
        const uint8_t p = x[i - 1];
        Cell c;

        // I layer: extension wins ties; an opening records the S move of the cell it leaves
        int32_t i_ext = I[curr][i - 1] + e;
        int32_t i_open = S[curr][i - 1] + o + e;
        int32_t best_i = std::max(i_ext, i_open);
        if (i_ext == best_i) {
            c.set_i(TB_INS, cell(i - 1, j).i_len() + 1);
        } else {
            SValue s = cell(i - 1, j).s();
            c.set_i(s.tb, s.len + 1);
        }

        // D layer
        int32_t d_ext = D[prev][i] + e;
        int32_t d_open = S[prev][i] + o + e;
        int32_t best_d = std::max(d_ext, d_open);
        if (d_ext == best_d) {
            c.set_d(TB_DEL, cell(i, j - 1).d_len() + 1);
        } else {
            SValue s = cell(i, j - 1).s();
            c.set_d(s.tb, s.len + 1);
        }

        // S layer.  Start from whatever S[curr][i] holds (MIN, or the x-suffix tracker at i == m).
        c.set_s(TB_XCLIP_SUFFIX, cell(i, j).s_len());
        int32_t best = S[curr][i];
        const int32_t addend = sc.sub(p, q);
        const int32_t diag = S[prev][i - 1] + addend;
        const uint32_t diag_len = cell(i - 1, j - 1).s_len() + 1;
        const uint8_t mtb = (p == q) ? TB_MATCH : TB_SUBST;
        if (diag >= best) {
            best = diag;
            c.set_s_all(mtb, diag_len, contig_idx, (uint32_t)(i - 1));
        }
        if (best_d > best) {
            best = best_d;
            c.set_s_all(TB_DEL, c.d_len(), contig_idx, (uint32_t)i);
        }
        if (best_i > best) {
            best = best_i;
            c.set_s_all(TB_INS, c.i_len(), contig_idx, (uint32_t)(i - 1));
        }
        // jump move (SCA:242-290): add the substitution score; circular zero-cost wrap at i == 1
        JumpInfo jp = jump;
        jp.score += addend;
        if (circular && i == 1) {
            const Cell &endc = cell(m, j - 1);
            if (endc.s_tb() != TB_XCLIP_SUFFIX) {
                int32_t w = S[prev][m] + addend;
                if (!(jp.score > w)) {
                    uint32_t wl = endc.s_len() + 1;
                    if (!(w == jp.score && wl <= jp.len)) {
                        jp = JumpInfo{w, wl, contig_idx, (uint32_t)m};
                    }
                }
            }
        }
        if (jp.score > best || (jp.score == best && best == diag && jp.len > diag_len)) {
            best = jp.score;
            c.set_s_all(mtb, jp.len, jp.idx, jp.from);
        }
        if (xclip_score > best) {
            best = xclip_score;
            c.set_s_all(TB_XCLIP_PREFIX, cell(0, j).s_len(), contig_idx, 0);
        }
        int32_t yclip_score = sc.yclip_prefix + o + e * (int32_t)i;
        if (yclip_score > best) {
            uint32_t pl = cell(i, 0).s_len();
            best = yclip_score;
            c.set_s_all(TB_YCLIP_PREFIX, pl, contig_idx, (uint32_t)i);
        }

        S[curr][i] = best;
        I[curr][i] = best_i;
        D[curr][i] = best_d;

        // x-suffix tracker into S[curr][m] / cell(m, j) / Lx[j]
        {
            int32_t t = S[curr][i] + sc.xclip_suffix;
            bool clip = t > S[curr][m] || (t == S[curr][m] && c.s_len() > cell(m, j).s_len());
            if (clip) {
                S[curr][m] = t;
                SValue ps = c.s();
                cell(m, j).set_s_all(TB_XCLIP_SUFFIX, ps.len, ps.idx, (uint32_t)i);
                Lx[j] = m - i;
            }
        }
        // y-suffix tracker into Sn[i] / Ly[i]; the tie-break reads the LAST column's cell
        {
            int32_t u = S[curr][i] + sc.yclip_suffix;
            bool clip = u > Sn[i] || (u == Sn[i] && c.s_len() > cell(i, n).s_len());
            if (clip) {
                Sn[i] = u;
                Ly[i] = n - j;
            }
        }
        cell(i, j) = c;
    }
}

void SingleContig::fill_last_column(int64_t m, int64_t n) {   // SCA:453-555
    const int64_t j = n;
    const int curr = (int)(j % 2);
    for (int64_t i = 0; i <= m; ++i) {
        // jump over the remaining contig bases to (m, n)
        if (S[curr][i] + sc.jump_same > S[curr][m]) {
            S[curr][m] = S[curr][i] + sc.jump_same;
            SValue ps = cell(i, j).s();
            cell(m, j).set_s_all(TB_XJUMP, ps.len, ps.idx, (uint32_t)i);
        }
        // y-suffix clip (a tie compares a cell's length with itself: never)
        if (Sn[i] > S[curr][i]) {
            S[curr][i] = Sn[i];
            if (j - Ly[i] < 0) panic("j - Ly[i] underflow");
            SValue sv = cell(i, j - Ly[i]).s();
            cell(i, j).set_s_all(TB_YCLIP_SUFFIX, sv.len, sv.idx, (uint32_t)i);
        }
        // x-suffix clip
        {
            int32_t t = S[curr][i] + sc.xclip_suffix;
            bool clip = t > S[curr][m] || (t == S[curr][m] && cell(i, j).s_len() > cell(m, j).s_len());
            if (clip) {
                S[curr][m] = t;
                Lx[j] = m - i;
                SValue ps = cell(i, j).s();
                cell(m, j).set_s_all(TB_XCLIP_SUFFIX, ps.len, ps.idx, (uint32_t)i);
            }
        }
    }
    // S may have changed in the last column: repair I
    for (int64_t i = 1; i <= m; ++i) {
        int32_t i_score = S[curr][i - 1] + sc.gap_open + sc.gap_extend;
        if (i_score > I[curr][i]) {
            I[curr][i] = i_score;
            SValue sv = cell(i - 1, j).s();
            cell(i, j).set_i(sv.tb, sv.len + 1);
        }
        if (i_score > S[curr][i]) {
            S[curr][i] = i_score;
            uint32_t pl = cell(i, j).i_len();
            cell(i, j).set_s_all(TB_INS, pl, contig_idx, (uint32_t)(i - 1));
            if (S[curr][i] + sc.xclip_suffix > S[curr][m]) {
                S[curr][m] = S[curr][i] + sc.xclip_suffix;
                Lx[j] = m - i;
                cell(m, j).set_s_all(TB_XCLIP_SUFFIX, pl, contig_idx, (uint32_t)i);
            }
        }
    }
}

Alignment SingleContig::custom(const uint8_t *x, int64_t m, const uint8_t *y, int64_t n) {   // SCA:705-729
    init_matrices(m, n);
    for (int64_t j = 1; j <= n; ++j) {
        int curr = (int)(j % 2), prev = 1 - curr;
        init_column(j, curr, m, n);
        JumpInfo ji = jump_info(m, j - 1, sc.jump_same);
        fill_column(x, y, m, n, j, prev, curr, ji);
    }
    fill_last_column(m, n);
    std::vector<const SingleContig *> al{this};
    return traceback_best(al, n);
}

Alignment SingleContig::with_mode(Mode mode, const uint8_t *x, int64_t m, const uint8_t *y, int64_t n) {
    // SCA:733-872: temporarily set the clip penalties, align, filter the clip ops of the free ends
    Scoring saved = sc;
    sc.set_clips_for(mode);
    Alignment a = custom(x, m, y, n);
    a.mode = mode;
    auto drop = [&](bool dx, bool dy) {
        std::vector<Op> kept;
        for (const Op &op : a.ops) {
            if ((dx && op.kind == XCLIP) || (dy && op.kind == YCLIP)) continue;
            kept.push_back(op);
        }
        a.ops.swap(kept);
    };
    if (mode == QUERY_LOCAL) drop(false, true);
    else if (mode == TARGET_LOCAL) drop(true, false);
    else if (mode == LOCAL) drop(true, true);
    sc = saved;
    return a;
}

// ---------------------------------------------------------------------------------------------
// Traceback (traceback/mod.rs:129-373)
// ---------------------------------------------------------------------------------------------
static size_t pick_end(const std::vector<const SingleContig *> &al, int64_t n,
                       const std::vector<bool> *consider, const std::vector<bool> *seen) {
    size_t off = 0;
    int32_t score = MIN_SCORE;
    uint32_t alen = 0;
    for (size_t a = 0; a < al.size(); ++a) {
        const SingleContig *c = al[a];
        if (consider && !((*consider).size() > c->contig_idx && (*consider)[c->contig_idx])) continue;
        if (seen && (*seen)[c->contig_idx]) continue;
        int64_t m = c->rows - 1;
        int32_t cs = c->S[n % 2][(size_t)m];
        uint32_t cl = c->cell(m, n).s_len();
        if (cs > score || (cs == score && cl > alen)) { off = a; score = cs; alen = cl; }
    }
    return off;
}

Alignment traceback_best(const std::vector<const SingleContig *> &al, int64_t n) {   // :129-150
    size_t off = pick_end(al, n, nullptr, nullptr);
    Alignment out;
    if (!traceback_from(al, n, al[off]->contig_idx, out)) panic("traceback_from returned None");
    return out;
}

bool traceback_from(const std::vector<const SingleContig *> &al, int64_t n, uint32_t contig_index,
                    Alignment &out) {   // :219-373
    if (al.empty()) panic("no aligners");
    uint32_t max_idx = 0;
    for (auto *a : al) max_idx = std::max(max_idx, a->contig_idx);
    std::vector<int64_t> map((size_t)max_idx + 1, -1);
    for (size_t a = 0; a < al.size(); ++a)
        if (!al[a]->tb.empty()) map[al[a]->contig_idx] = (int64_t)a;
    auto lookup = [&](uint32_t idx) -> const SingleContig * {
        if (idx >= map.size()) panic("IndexMap::get out of range");   // index_map.rs:47
        return map[idx] < 0 ? nullptr : al[(size_t)map[idx]];
    };
    if (contig_index >= map.size() || map[contig_index] < 0) return false;   // :239-241

    const SingleContig *cur = lookup(contig_index);
    int64_t j = n;
    std::vector<Op> ops;
    int64_t xstart = 0, ystart = 0, yend = n;
    const int32_t score = cur->S[n % 2][(size_t)(cur->rows - 1)];
    const uint32_t alen = cur->cell(cur->rows - 1, n).s_len();
    const uint32_t end_idx = cur->contig_idx;
    const int64_t xlen = cur->rows - 1;
    uint32_t cur_idx = end_idx;
    int64_t i = cur->rows - 1, xend = cur->rows - 1;
    uint8_t layer = cur->cell(i, j).s_tb();
    auto need = [&](bool ok) { if (!ok) panic("traceback index underflow"); };

    for (;;) {
        cur = lookup(cur_idx);
        if (!cur) return false;
        uint8_t next;
        switch (layer) {
        case TB_START: goto done;
        case TB_INS:
            ops.push_back(Op{INS});
            next = cur->cell(i, j).i_tb();
            need(i >= 1); i -= 1;
            break;
        case TB_DEL:
            ops.push_back(Op{DEL});
            next = cur->cell(i, j).d_tb();
            need(j >= 1); j -= 1;
            break;
        case TB_MATCH: case TB_SUBST: {
            ops.push_back(Op{layer == TB_MATCH ? (uint8_t)MATCH : (uint8_t)SUBST});
            SValue sv = cur->cell(i, j).s();
            need(i >= 1);
            if (sv.idx != cur_idx || (int64_t)sv.from != i - 1) {
                ops.push_back(Op{XJUMP, cur_idx, (uint32_t)(i - 1)});
                cur_idx = sv.idx;
                cur = lookup(cur_idx);
                if (!cur) return false;
            }
            i = sv.from;
            need(j >= 1); j -= 1;
            need(i < cur->rows);
            next = cur->cell(i, j).s_tb();
            break;
        }
        case TB_XCLIP_PREFIX:
            next = cur->cell(0, j).s_tb();
            if (next == TB_START || next == TB_YCLIP_PREFIX) {
                ops.push_back(Op{XCLIP, (uint32_t)i});
                xstart = i;
            }
            i = 0;
            break;
        case TB_XCLIP_SUFFIX:
            if (ops.empty() || ops.front().kind == YCLIP) {
                ops.push_back(Op{XCLIP, (uint32_t)cur->Lx[(size_t)j]});
                xend = i - cur->Lx[(size_t)j];
            }
            i -= cur->Lx[(size_t)j];
            need(i >= 0);
            next = cur->cell(i, j).s_tb();
            break;
        case TB_YCLIP_PREFIX:
            ops.push_back(Op{YCLIP, (uint32_t)j});
            ystart = j;
            j = 0;
            next = cur->cell(i, 0).s_tb();
            break;
        case TB_YCLIP_SUFFIX: {
            ops.push_back(Op{YCLIP, (uint32_t)cur->Ly[(size_t)i]});
            int64_t sfrom = cur->cell(i, j).from();
            j -= cur->Ly[(size_t)i];
            need(j >= 0);
            if (sfrom != i) {
                ops.push_back(Op{XJUMP, cur_idx, (uint32_t)i});
                i = sfrom;
            }
            yend = j;
            next = cur->cell(i, j).s_tb();
            break;
        }
        case TB_XJUMP: {
            SValue sv = cur->cell(i, j).s();
            ops.push_back(Op{XJUMP, cur_idx, (uint32_t)i});
            cur_idx = sv.idx;
            cur = lookup(cur_idx);
            if (!cur) return false;
            i = sv.from;
            need(i < cur->rows);
            next = cur->cell(i, j).s_tb();
            break;
        }
        default: panic("unexpected traceback layer");
        }
        layer = next;
    }
done:
    std::reverse(ops.begin(), ops.end());
    bool only_special = true;
    for (const Op &op : ops) if (!op.is_special()) { only_special = false; break; }
    if (only_special) xstart = xend = ystart = yend = 0;
    out = Alignment{};
    out.score = score; out.ystart = ystart; out.xstart = xstart; out.yend = yend; out.xend = xend;
    out.xlen = xlen; out.ylen = n; out.start_contig_idx = cur_idx; out.end_contig_idx = end_idx;
    out.ops = std::move(ops); out.mode = CUSTOM; out.length = alen;
    return true;
}

std::vector<Alignment> traceback_all(const std::vector<const SingleContig *> &al, int64_t n,
                                     const std::vector<bool> &consider, size_t n_consider) {   // :152-217
    std::vector<Alignment> outs;
    std::vector<bool> seen(consider.size(), false);
    size_t n_seen = 0;
    auto mark = [&](int64_t idx) {
        if (idx >= 0 && (size_t)idx < consider.size() && consider[(size_t)idx] && !seen[(size_t)idx]) {
            seen[(size_t)idx] = true; ++n_seen;
        }
    };
    while (n_seen < n_consider) {
        size_t off = pick_end(al, n, &consider, &seen);
        Alignment a;
        if (!traceback_from(al, n, al[off]->contig_idx, a)) {
            mark(al[off]->contig_idx);
            continue;
        }
        mark(a.start_contig_idx);
        mark(a.end_contig_idx);
        for (const Op &op : a.ops) if (op.kind == XJUMP) mark(op.a);
        outs.push_back(std::move(a));
    }
    return outs;
}

// ---------------------------------------------------------------------------------------------
// MultiContig (multi_contig_aligner.rs)
// ---------------------------------------------------------------------------------------------
void MultiContig::add_contig(const std::string &name, bool is_forward, const uint8_t *seq, int64_t len,
                             bool circular, const Scoring &sc) {   // :93-133
    for (auto &c : contigs)
        if (c.is_forward == is_forward && c.name == name) panic("Contig already added");
    Contig c;
    c.name = name; c.is_forward = is_forward; c.seq = seq; c.len = len;
    c.aligner.sc = sc;
    c.aligner.contig_idx = (uint32_t)contigs.size();
    c.aligner.circular = circular;
    contigs.push_back(std::move(c));
}

Alignment MultiContig::custom(const uint8_t *y, int64_t n) {   // :231-361
    const size_t C = contigs.size();
    if (C == 0) panic("no contigs");
    // opposite strand by (same name, other strand) among the contigs present; values are POSITIONS
    std::vector<int64_t> opp(C, -1);
    for (size_t a = 0; a < C; ++a) {
        if (opp[a] >= 0) continue;
        for (size_t b = a + 1; b < C; ++b) {
            if (contigs[a].name == contigs[b].name && contigs[a].is_forward != contigs[b].is_forward) {
                if (opp[a] >= 0) panic("more than one opposite strand");   // assert at :255-257
                opp[a] = (int64_t)b;
                opp[b] = (int64_t)a;
            }
        }
    }
    for (auto &c : contigs) c.aligner.init_matrices(c.len, n);
    ++fills;
    for (auto &c : contigs) cells_filled += (uint64_t)c.len * (uint64_t)n;

    std::vector<JumpInfo> inter(C), best(C);
    for (int64_t j = 1; j <= n; ++j) {
        int curr = (int)(j % 2), prev = 1 - curr;
        for (auto &c : contigs) c.aligner.init_column(j, curr, c.len, n);
        for (size_t a = 0; a < C; ++a)
            inter[a] = contigs[a].aligner.jump_info(contigs[a].len, j - 1, contigs[a].aligner.sc.jump_inter);
        for (size_t a = 0; a < C; ++a) {
            const Contig &c = contigs[a];
            JumpInfo bj = c.aligner.jump_info(c.len, j - 1, c.aligner.sc.jump_same);
            uint32_t opp_idx = c.aligner.contig_idx;
            if (opp[a] >= 0) {
                const Contig &oc = contigs[(size_t)opp[a]];
                JumpInfo fl = oc.aligner.jump_info(oc.len, j - 1, oc.aligner.sc.jump_opp);
                opp_idx = oc.aligner.contig_idx;
                if (fl.score > bj.score) bj = fl;
            }
            // inter-contig: max by (score, len), LAST maximal element wins (Iterator::max_by_key)
            bool have = false;
            JumpInfo ic{};
            for (size_t b = 0; b < C; ++b) {
                const JumpInfo &cand = inter[b];
                if (cand.idx == c.aligner.contig_idx || cand.idx == opp_idx) continue;
                if (!have || cand.score > ic.score || (cand.score == ic.score && cand.len >= ic.len)) {
                    ic = cand; have = true;
                }
            }
            if (have && ic.score > bj.score) bj = ic;
            best[a] = bj;
        }
        for (size_t a = 0; a < C; ++a) {
            Contig &c = contigs[a];
            c.aligner.fill_column(c.seq, y, c.len, n, j, prev, curr, best[a]);
        }
    }
    for (auto &c : contigs) c.aligner.fill_last_column(c.len, n);
    std::vector<const SingleContig *> al;
    for (auto &c : contigs) al.push_back(&c.aligner);
    return traceback_best(al, n);
}

Alignment MultiContig::custom_with_subset(const uint8_t *y, int64_t n, const std::vector<bool> *subset) {   // :178-223
    if (!subset) return custom(y, n);
    bool any = false;
    for (bool b : *subset) any = any || b;
    if (!any) panic("Subsetted to an empty set of contigs");
    std::vector<Contig> included, excluded;
    for (auto &c : contigs) {
        uint32_t idx = c.aligner.contig_idx;
        if (idx < subset->size() && (*subset)[idx]) included.push_back(std::move(c));
        else excluded.push_back(std::move(c));
    }
    if (included.empty()) panic("no contig included");
    contigs = std::move(included);
    Alignment a = custom(y, n);
    for (auto &c : excluded) contigs.push_back(std::move(c));
    std::stable_sort(contigs.begin(), contigs.end(),
                     [](const Contig &l, const Contig &r) { return l.aligner.contig_idx < r.aligner.contig_idx; });
    return a;
}

std::vector<Alignment> MultiContig::traceback_all(int64_t n, const std::vector<bool> *subset) {   // :363-378
    std::vector<bool> consider(contigs.size(), false);
    size_t cnt = 0;
    size_t sub_cnt = 0;
    if (subset) for (bool b : *subset) sub_cnt += b;
    if (subset && sub_cnt < contigs.size()) {
        consider.assign(std::max(contigs.size(), subset->size()), false);
        for (size_t k = 0; k < subset->size(); ++k) if ((*subset)[k]) { consider[k] = true; ++cnt; }
    } else {
        for (auto &c : contigs) { consider[c.aligner.contig_idx] = true; ++cnt; }
    }
    std::vector<const SingleContig *> al;
    for (auto &c : contigs) al.push_back(&c.aligner);
    return oracle::traceback_all(al, n, consider, cnt);
}

bool MultiContig::traceback_from(int64_t n, uint32_t contig_index, Alignment &out) {   // :380-387
    std::vector<const SingleContig *> al;
    for (auto &c : contigs) al.push_back(&c.aligner);
    return oracle::traceback_from(al, n, contig_index, out);
}

// ---------------------------------------------------------------------------------------------
// Alignment helpers (alignment.rs)
// ---------------------------------------------------------------------------------------------
static std::string op_string(const Op &op, int64_t contig_idx, int64_t x_index) {   // constants.rs:37-59
    switch (op.kind) {
    case MATCH: return "=";
    case SUBST: return "X";
    case DEL: return "D";
    case INS: return "I";
    case XCLIP: return std::to_string(op.a) + "A";
    case YCLIP: return std::to_string(op.a) + "B";
    case XJUMP: {
        std::string s;
        int64_t nc = op.a, nx = op.b;
        if (nc > contig_idx) s = std::to_string(nc - contig_idx) + "C";
        else if (nc < contig_idx) s = std::to_string(contig_idx - nc) + "c";
        if (nx >= x_index) s += std::to_string(nx - x_index) + "J";
        else s += std::to_string(x_index - nx) + "j";
        return s;
    }
    default: return std::to_string(op.a) + "S";
    }
}

std::string Alignment::cigar() const {   // alignment.rs:105-149
    std::string out;
    if (ops.empty()) return out;
    int64_t contig = start_contig_idx;
    int64_t x_index = xstart;
    const Op *last = &ops.front();
    int64_t last_len = 0;
    for (const Op &op : ops) {
        if ((op.is_special() || !(op == *last)) && last_len > 0)
            out += std::to_string(last_len) + op_string(*last, contig, x_index);
        if (op.is_special()) {
            out += op_string(op, contig, x_index);
            x_index += op.len_x(x_index);
            last = &op;
            last_len = 0;
            if (op.kind == XJUMP) contig = op.a;
        } else if (op == *last) {
            x_index += op.len_x(x_index);
            last_len += 1;
        } else {
            x_index += op.len_x(x_index);
            last = &op;
            last_len = 1;
        }
    }
    if (last_len > 0) out += std::to_string(last_len) + op_string(*last, contig, x_index);
    return out;
}

static inline bool is_aligned_op(const Op &op) { return op.kind == MATCH || op.kind == SUBST || op.kind == DEL || op.kind == INS; }

Alignment Alignment::split_at_y(int64_t y_pivot) const {   // alignment.rs:207-360
    if (ops.empty()) return *this;
    if (ops.front().kind == XCLIP || ops.front().kind == YCLIP) panic("leading clip in split_at_y");
    if (ops.back().kind == XCLIP || ops.back().kind == YCLIP) panic("trailing clip in split_at_y");

    int64_t x_index = xstart, y_index = ystart, contig = start_contig_idx;
    size_t k = 0;
    // leading special ops
    for (const Op &op : ops) {
        if (is_aligned_op(op)) break;
        if (op.kind == XJUMP) contig = op.a;
        y_index += op.len_y();
        x_index += op.len_x(x_index);
        ++k;
    }
    // up to the pivot
    for (size_t t = k; t < ops.size(); ++t) {
        const Op &op = ops[t];
        if (y_index + op.len_y() >= y_pivot) break;
        if (op.kind == XJUMP) contig = op.a;
        y_index += op.len_y();
        x_index += op.len_x(x_index);
        ++k;
    }
    if (k >= ops.size()) panic("split_at_y: op index out of range");   // slice [..=op_index]
    Alignment pre;
    pre.xstart = xstart; pre.xend = x_index + 1; pre.ystart = ystart; pre.yend = y_index + 1;
    pre.start_contig_idx = start_contig_idx; pre.end_contig_idx = contig;
    pre.ops.assign(ops.begin(), ops.begin() + (std::ptrdiff_t)k + 1);
    pre.mode = mode;
    if (!(y_pivot >= pre.yend)) panic("y_pivot < pre.yend");

    // special ops at the pivot
    {
        size_t start = k;
        for (size_t t = start; t < ops.size(); ++t) {
            const Op &op = ops[t];
            if (y_index >= y_pivot && is_aligned_op(op)) break;
            if (op.kind == XJUMP) contig = op.a;
            y_index += op.len_y();
            x_index += op.len_x(x_index);
            ++k;
        }
    }
    Alignment post;
    post.xstart = x_index; post.xend = xend; post.ystart = y_index; post.yend = yend;
    post.start_contig_idx = contig; post.end_contig_idx = end_contig_idx;
    post.ops.assign(ops.begin() + (std::ptrdiff_t)std::min(k, ops.size()), ops.end());
    post.mode = mode;

    Alignment a;
    a.start_contig_idx = post.start_contig_idx;
    a.end_contig_idx = pre.end_contig_idx;
    a.xstart = post.xstart;
    a.ystart = post.ystart - y_pivot;
    a.xend = pre.xend;
    a.yend = pre.yend + ylen - y_pivot;
    a.ylen = ylen; a.xlen = xlen; a.score = score; a.mode = mode; a.length = length;
    if (a.ystart < 0) panic("split_at_y: ystart underflow");

    const bool x_clip = (mode == GLOBAL || mode == QUERY_LOCAL);
    const bool y_clip = (mode == GLOBAL || mode == TARGET_LOCAL);
    if (x_clip && a.xstart > 0) { a.ops.push_back(Op{XCLIP, (uint32_t)a.xstart}); a.xstart = 0; }
    if (y_clip && a.ystart > 0) { a.ops.push_back(Op{YCLIP, (uint32_t)a.ystart}); a.ystart = 0; }
    a.ops.insert(a.ops.end(), post.ops.begin(), post.ops.end());
    if (pre.start_contig_idx != post.end_contig_idx || pre.xstart != post.xend)
        a.ops.push_back(Op{XJUMP, (uint32_t)pre.start_contig_idx, (uint32_t)pre.xstart});
    int64_t yjump = a.ylen + pre.ystart - post.yend;
    if (yjump < 0) panic("split_at_y: yjump underflow");
    if (yjump > 0) a.ops.push_back(Op{YJUMP, (uint32_t)yjump});
    a.ops.insert(a.ops.end(), pre.ops.begin(), pre.ops.end());
    if (x_clip && a.xend < a.xlen) { a.ops.push_back(Op{XCLIP, (uint32_t)(a.xlen - a.xend)}); a.xend = a.xlen; }
    if (y_clip && a.yend < a.ylen) { a.ops.push_back(Op{XCLIP, (uint32_t)(a.ylen - a.yend)}); a.yend = a.ylen; }   // sic: Xclip, alignment.rs:355
    return a;
}

std::vector<uint8_t> reverse_complement(const std::vector<uint8_t> &s) {   // dna.rs:5-41
    static uint8_t comp[256];
    static bool init = false;
    if (!init) {
        for (int v = 0; v < 256; ++v) comp[v] = (uint8_t)v;
        const char *a = "AGCTYRWSKMDVHBN", *b = "TCGARYWSMKHBDVN";
        for (int k = 0; k < 15; ++k) {
            comp[(uint8_t)a[k]] = (uint8_t)b[k];
            comp[(uint8_t)a[k] + 32] = (uint8_t)(b[k] + 32);
        }
        init = true;
    }
    std::vector<uint8_t> r(s.size());
    for (size_t k = 0; k < s.size(); ++k) r[k] = comp[s[s.size() - 1 - k]];
    return r;
}

// ---------------------------------------------------------------------------------------------
// Options / Aligners (aligners/mod.rs)
// ---------------------------------------------------------------------------------------------
Scoring Options::contig_scoring() const {   // mod.rs:143-167
    Scoring s;
    s.match = match_score; s.mismatch = mismatch_score;
    s.gap_open = gap_open; s.gap_extend = gap_extend;
    s.jump_same = jump_same; s.jump_opp = jump_opp; s.jump_inter = jump_inter;
    s.set_clips_for(mode);
    return s;
}

Aligners::Aligners(const Options &o, const std::vector<std::string> &nm,
                   const std::vector<std::vector<uint8_t>> &f) : opts(o), names(nm), fwd(f) {   // mod.rs:171-211
    if (o.gap_open > 0 || o.gap_extend > 0) panic("gap scores can't be positive");
    if (o.jump_same > 0 || o.jump_opp > 0 || o.jump_inter > 0) panic("jump scores can't be positive");
    Scoring sc = opts.contig_scoring();
    for (auto &s : fwd) rev.push_back(reverse_complement(s));
    for (size_t k = 0; k < fwd.size(); ++k)
        mc.add_contig(names[k], true, fwd[k].data(), (int64_t)fwd[k].size(), opts.circular, sc);
    if (opts.double_strand)
        for (size_t k = 0; k < fwd.size(); ++k)
            mc.add_contig(names[k], false, rev[k].data(), (int64_t)rev[k].size(), opts.circular, sc);
}

Alignment Aligners::remove_clipping(Alignment a) const {   // mod.rs:343-353
    if (opts.mode == LOCAL || opts.mode == QUERY_LOCAL || opts.mode == TARGET_LOCAL) {
        std::vector<Op> kept;
        for (const Op &op : a.ops)
            if (is_aligned_op(op) || op.kind == XJUMP) kept.push_back(op);
        a.ops.swap(kept);
    }
    return a;
}

Alignment Aligners::multi_contig_align(const uint8_t *q, int64_t n, const std::vector<bool> *subset) {
    return remove_clipping(mc.custom_with_subset(q, n, subset));
}

bool Aligners::realign_and_split(const std::vector<uint8_t> &q, const Alignment &best,
                                 const std::vector<bool> &subset, int64_t contig_idx, int64_t y_pivot,
                                 Alignment &out) {   // mod.rs:412-431
    multi_contig_align(q.data(), (int64_t)q.size(), &subset);
    Alignment na;
    if (!mc.traceback_from((int64_t)q.size(), (uint32_t)contig_idx, na)) return false;
    if (na.score > best.score && na.start_contig_idx == contig_idx && best.end_contig_idx == contig_idx) {
        out = remove_clipping(na).split_at_y(y_pivot);
        return true;
    }
    return false;
}

Alignment Aligners::realign_origin(const std::vector<uint8_t> &query, const Alignment &alignment, int64_t slop) {   // mod.rs:442-553
    // mod.rs:365-410
    int64_t at_start = -1, at_end = -1;
    auto circ = [&](int64_t idx) { return mc.contigs[(size_t)idx].aligner.circular; };
    if (alignment.xstart <= slop && circ(alignment.start_contig_idx)) at_start = alignment.start_contig_idx;
    if (alignment.xlen <= alignment.xend + slop && circ(alignment.end_contig_idx)) at_end = alignment.end_contig_idx;
    if (at_start >= 0 && at_end >= 0 && at_start == at_end) return alignment;
    if (at_start < 0 && at_end < 0) return alignment;
    if (at_start >= 0 && alignment.yend == alignment.ylen) at_start = -1;
    if (at_end >= 0 && alignment.ystart == 0) at_end = -1;
    if (at_start < 0 && at_end < 0) return alignment;

    std::vector<bool> subset(mc.contigs.size(), false);
    subset[(size_t)alignment.start_contig_idx] = true;
    subset[(size_t)alignment.end_contig_idx] = true;
    for (const Op &op : alignment.ops) if (op.kind == XJUMP) subset[op.a] = true;

    auto rotate = [&](int64_t at) {
        std::vector<uint8_t> r(query.begin() + at, query.end());
        r.insert(r.end(), query.begin(), query.begin() + at);
        return r;
    };
    Alignment best = alignment;
    if (at_start >= 0) {
        int64_t y1 = alignment.yend;
        int64_t y2 = alignment.ystart;
        for (const Op &op : alignment.ops) {
            if (op.kind == XJUMP && (int64_t)op.a != at_start) break;
            y2 += op.len_y();
        }
        for (int64_t yend : {y1, y2}) {
            Alignment cand;
            if (realign_and_split(rotate(yend), best, subset, at_start, alignment.ylen - yend, cand)) best = cand;
        }
    }
    if (at_end >= 0) {
        int64_t y1 = alignment.ystart;
        int64_t y2 = alignment.ystart, ycur = alignment.ystart, xidx = alignment.start_contig_idx;
        for (const Op &op : alignment.ops) {
            if (op.kind == XJUMP) {
                if ((int64_t)op.a == at_end && xidx != at_end) y2 = ycur;
                xidx = op.a;
            }
            ycur += op.len_y();
        }
        for (int64_t ystart : {y1, y2}) {
            Alignment cand;
            if (realign_and_split(rotate(ystart), best, subset, at_end, alignment.ylen - ystart, cand)) best = cand;
        }
    }
    return best;
}

std::vector<Alignment> Aligners::align(const uint8_t *read, int64_t n, const std::vector<bool> *subset) {   // mod.rs:237-340
    std::vector<uint8_t> query(read, read + n);
    for (auto &b : query) if (b >= 'a' && b <= 'z') b = (uint8_t)(b - 32);   // io.rs:64
    Alignment original = multi_contig_align(query.data(), n, subset);
    std::vector<Alignment> out;
    if (opts.suboptimal) {
        std::vector<Alignment> all = mc.traceback_all(n, subset);
        for (Alignment &a : all) out.push_back(realign_origin(query, remove_clipping(a), opts.circular_slop));
        if (out.size() > 1) {
            std::stable_sort(out.begin(), out.end(), [](const Alignment &l, const Alignment &r) { return -l.score < -r.score; });
            float min_score = (float)out[0].score * opts.suboptimal_pct / 100.0f;
            std::vector<Alignment> kept;
            for (Alignment &a : out) if ((float)a.score >= min_score) kept.push_back(std::move(a));
            out.swap(kept);
        }
    } else {
        out.push_back(realign_origin(query, original, opts.circular_slop));
    }
    return out;
}

}  // namespace oracle
