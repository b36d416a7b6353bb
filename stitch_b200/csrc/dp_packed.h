// dp_packed.h — the fast arithmetic of the fill: one 32-bit key per DP value.
//
//   key = [ relative score | 4-bit priority | alignment length ]      (signed 32-bit compare)
//           32-SH bits        bits LB..LB+3    LB bits        SH = LB + 4
//
// The reference (SCA = single_contig_aligner.rs, MCA = multi_contig_aligner.rs) keeps every DP
// value as an i32 score plus a 27-bit alignment length and decides every cell by a fixed sequence
// of "replace if strictly greater" comparisons (SCA:350-399) with two length tie-breaks
// (SCA:373-377 jump vs diagonal, SCA:272-281 circular wrap vs jump).  Packing (score, priority,
// length) into one integer turns that sequence into plain integer max operations:
//   * a candidate that must WIN score ties gets the higher priority, so max() reproduces the
//     "strictly greater" order diag > D > I > jump > x-clip > y-clip;
//   * the jump candidate inherits the diagonal's priority when (and only when) the running best is
//     the diagonal, so that exactly then the longer alignment wins an equal score (SCA:373-377);
//   * lengths ride along in the low bits: one integer add updates score and length together.
// Scores are stored relative to the best cell of the previous column (B_j = G(j-1), MCA:279-331
// makes a jump from that cell available to every cell, so every live value lies in a band of
// width max(match,mismatch,0) - min(match,mismatch) - min(jump scores) below B_j + max(...)); the
// change of base is folded into the per-column add constants.  Values below the band (they exist
// only in column 0 and in D of column 1) are clamped to NEG; a clamped value can never win or tie
// against the jump candidate, so its exact value never matters.
//
// The in-column insertion chain I(i) = max_{k<i} H'(k) + o + e(i-k) (SCA:317-326) has a bounded
// reach R = floor(band/|e|) + 1 rows: a source further back is strictly worse than opening from
// row i-1.  When R <= STRIP a lane's chain needs nothing beyond the previous lane's strip, so the
// cross-lane dependency is ONE warp shuffle (no scan); the host only selects this path for such
// scorings (pk_plan) and everything else runs on the exact wide kernels.
//
// Everything here is __host__ __device__: the CUDA kernels (kernels_packed.cuh) and the CPU
// emulator the tests fuzz against the oracle run the same code.
#pragma once
#include "dp_core.h"

namespace stitch {

enum : int32_t { PP_YC = 1, PP_XC = 2, PP_JMP = 3, PP_INC = 4, PP_ICARRY = 5, PP_D = 6, PP_DIAG = 11, PP_DIAGBIT = 8 };

struct PK {                 // per job
    int32_t LB, SH;
    int32_t NEG;            // lowest relative score kept exactly
    int32_t NEGKEY;         // NEG << SH
    int32_t LMASK, NPM;     // (1 << LB) - 1,  ~(15 << LB)
    int32_t PD6, PI4, PI5, PB8, P1;
};

SHD int32_t pk_band(const Scoring &sc) {
    int32_t submax = sc.match > sc.mismatch ? sc.match : sc.mismatch;
    if (submax < 0) submax = 0;
    const int32_t submin = sc.match < sc.mismatch ? sc.match : sc.mismatch;
    int32_t gmin = sc.g_same < sc.g_opp ? sc.g_same : sc.g_opp;
    gmin = gmin < sc.g_inter ? gmin : sc.g_inter;
    return submax - submin - gmin;
}
SHD int32_t pk_submax0(const Scoring &sc) {
    int32_t submax = sc.match > sc.mismatch ? sc.match : sc.mismatch;
    return submax < 0 ? 0 : submax;
}
SHD int32_t pk_neg(const Scoring &sc) {
    // every live D is >= (gmin + submin) + o + e - submax0 relative to its column base
    return -(pk_band(sc) - sc.o - sc.e + 2 * pk_submax0(sc) + 2);
}

// Can the packed path run this (scoring, read length, longest contig)?  Returns LB or 0.
SHD uint32_t pk_plan(const Scoring &sc, uint32_t n, uint32_t m_max) {
    if (sc.e >= 0) return 0;
    const int32_t band = pk_band(sc), ae = -sc.e;
    if (band / ae + 1 > STRIP) return 0;                       // insertion-chain reach
    const int64_t submax0 = pk_submax0(sc);
    int32_t submin = sc.match < sc.mismatch ? sc.match : sc.mismatch;
    if (submin > 0) submin = 0;
    int32_t gmin = sc.g_same < sc.g_opp ? sc.g_same : sc.g_opp;
    gmin = gmin < sc.g_inter ? gmin : sc.g_inter;
    // y-prefix clip (SCA:391-399) can only win in the first warp tile of a contig
    if (sc.yp != MIN_SCORE && sc.xp == MIN_SCORE && sc.o + sc.e * (int64_t)(TILE + 1) > (int64_t)gmin + submin) return 0;
    // length field: any cell's alignment length is at most j + #insertions, and a path that stays
    // inside the band cannot afford more than j*submax0/|e| + const insertions
    const int64_t len_bound = (int64_t)n * (2 + (submax0 + ae - 1) / ae) + (int64_t)(-gmin) - sc.o - submin + 64;
    int64_t need = len_bound > (int64_t)m_max + 2 ? len_bound : (int64_t)m_max + 2;
    uint32_t LB = 4;
    while (((int64_t)1 << LB) <= need) ++LB;
    if (LB > 24) return 0;
    const uint32_t SB = 32 - 4 - LB;                           // score bits
    const int64_t lim = 2 * ((int64_t)-pk_neg(sc) + band + submax0 - submin - sc.o + (int64_t)(STRIP + 1) * ae) + 16;
    if (SB < 4 || lim >= ((int64_t)1 << (SB - 1))) return 0;
    return LB;
}

SHD PK pk_make(const Scoring &sc, uint32_t LB) {
    PK p;
    p.LB = (int32_t)LB; p.SH = (int32_t)LB + 4;
    p.NEG = pk_neg(sc);
    p.NEGKEY = (int32_t)((uint32_t)p.NEG << p.SH);
    p.LMASK = (int32_t)((1u << LB) - 1u);
    p.NPM = (int32_t)~(15u << LB);
    p.PD6 = PP_D << LB; p.PI4 = PP_INC << LB; p.PI5 = PP_ICARRY << LB; p.PB8 = PP_DIAGBIT << LB; p.P1 = 1 << LB;
    return p;
}
SHD int32_t pk_shl(const PK &p, int32_t v) { return (int32_t)((uint32_t)v << p.SH); }   // v * 2^SH, wrap-around
SHD int32_t pk_key(const PK &p, int64_t rel, int32_t prio, uint32_t len) {
    if (rel < p.NEG) rel = p.NEG;
    return pk_shl(p, (int32_t)rel) + (prio << p.LB) + (int32_t)(len & (uint32_t)p.LMASK);
}
SHD int32_t pk_rel(const PK &p, int32_t key) { return key >> p.SH; }
SHD uint32_t pk_len(const PK &p, int32_t key) { return (uint32_t)(key & p.LMASK); }
SHD uint32_t pk_prio(const PK &p, int32_t key) { return (uint32_t)(key >> p.LB) & 15u; }
SHD int32_t pk_max(int32_t a, int32_t b) { return a > b ? a : b; }
#ifdef STITCH_PK_FLOORS   // (build variant for A/B measurements: floors in the bulk pass too)
constexpr bool PK_FLOORS = true;
#else
constexpr bool PK_FLOORS = false;
#endif
SHD int32_t pk_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
#if defined(__CUDA_ARCH__)
// DPX instructions of sm_90+/sm_100: VIMNMX3 and VIADDMNMX (one issue slot each)
SHD int32_t pk_max3(int32_t a, int32_t b, int32_t c) { return __vimax3_s32(a, b, c); }
SHD int32_t pk_addmax(int32_t a, int32_t b, int32_t c) { return __viaddmax_s32(a, b, c); }   // max(a + b, c)
#else
SHD int32_t pk_max3(int32_t a, int32_t b, int32_t c) { return pk_max(pk_max(a, b), c); }
SHD int32_t pk_addmax(int32_t a, int32_t b, int32_t c) { return pk_max((int32_t)((uint32_t)a + (uint32_t)b), c); }
#endif

// Wide <-> packed state conversion (checkpoints, hand-over to the wide tail, row m).
SHD int32_t pk_from_wide(const PK &p, int32_t B, int32_t score_abs, uint32_t len, int32_t prio) {
    return pk_key(p, (int64_t)score_abs - B, prio, len);
}
SHD int32_t pk_abs(const PK &p, int32_t B, int32_t key) {
    const int32_t r = pk_rel(p, key);
    return r <= p.NEG ? MIN_SCORE : B + r;   // a clamped value stands for "far below everything live"
}

struct PCol {               // per column of a read (uniform over the CTA)
    int32_t cM, cX;         // diagonal add (match / mismatch), includes the change of base, PP_DIAG and len+1
    int32_t cE, cOE;        // D layer: extension of a stored D (PP_D -> 1) / open from a stored S (-> 0)
    int32_t cEi, cOEi;      // in-column chain: extension (keeps PP_INC) / open from a clean H (-> PP_JMP)
    int32_t XC;             // x-prefix clip candidate (PP_XC) or NEGKEY
    int32_t B, delta;       // base of this column, B_j - B_{j-1}
    uint8_t q;              // read base y[j-1]
};

SHD PCol pk_col(const PK &p, const Scoring &sc, int32_t B, int32_t Bprev, uint32_t j, uint32_t n, uint8_t q) {
    PCol c;
    c.B = B; c.delta = B - Bprev; c.q = q;
    c.cM = pk_shl(p, sc.match - c.delta) + (PP_DIAG << p.LB) + 1;
    c.cX = pk_shl(p, sc.mismatch - c.delta) + (PP_DIAG << p.LB) + 1;
    c.cE = pk_shl(p, sc.e - c.delta) + (1 - PP_D) * (1 << p.LB) + 1;
    c.cOE = pk_shl(p, sc.o + sc.e - c.delta) + 1;
    c.cEi = pk_shl(p, sc.e) + 1;
    c.cOEi = pk_shl(p, sc.o + sc.e) + (PP_JMP << p.LB) + 1;
    if (sc.xp != MIN_SCORE) {   // SCA:304-308: xp + max(yp, o + e*j), length = s_len(0, j)
        const int64_t dj = (int64_t)sc.o + (int64_t)sc.e * j;
        const int64_t xs = (int64_t)sc.xp + ((int64_t)sc.yp > dj ? (int64_t)sc.yp : dj);
        c.XC = pk_key(p, xs - B, PP_XC, row0_at(sc, j, n).sl);
    } else c.XC = p.NEGKEY;
    return c;
}

// Jump candidate of a contig for this column, pre-adjusted so that jp = Jc + (cM | cX).
SHD int32_t pk_jc(const PK &p, const PCol &c, int32_t jscore_abs, uint32_t jlen) {
    return pk_key(p, (int64_t)jscore_abs - c.B, PP_JMP, jlen) + pk_shl(p, c.delta) - (PP_DIAG << p.LB) - 1;
}
// Circular wrap source (SCA:263-281): S(m, j-1) as stored (relative to B_{j-1}, priority 0).
SHD int32_t pk_wbase(const PK &p, int32_t sm_prev_key) { return sm_prev_key - p.PB8; }

struct PStrip {             // a lane's working set between pass 1 and pass 2
    int32_t A[STRIP];       // best of {diag, D}             (row m: the diagonal candidate alone)
    int32_t jp[STRIP];      // jump candidate, PP_JMP
    int32_t D6[STRIP];      // D of the cell, PP_D (what is stored)
    int32_t Inc[STRIP];     // chain value arriving at row k from the strip's own rows (no carry-in), PP_INC
    int32_t YC[STRIP];      // y-prefix clip candidates (special tiles only)
    int32_t exit;           // chain value leaving the strip, PP_INC
    uint32_t fl[STRIP];     // traceback variant: bit0 D is an extension, bit1 Inc[k] is an extension, bit2 wrap won
    uint32_t exit_open;     // traceback variant: exit opens from the strip's last row
};

// Pass 1 of one lane: everything that does not depend on the insertion chain of earlier lanes.
//   Sup/Dup[k] : stored S / D keys of row (row0+k) at column j-1;  Sdg0 : S key of row (row0-1) at j-1
//   SPECIAL    : first or last tile of a contig (rows >= nv are not ordinary cells; row m at index nv
//                when has_m; wrap / y-clip candidates are live)
template <bool SPECIAL, bool TB>
SHD void pk_pass1(const PK &p, const PCol &c, const int32_t *Sup, const int32_t *Dup, int32_t Sdg0, const uint8_t *x,
                  int32_t Jc, bool wrap0, int32_t wbase, int nv, bool has_m, PStrip &o) {
    int32_t Iacc = p.NEGKEY + p.PI4;
    uint32_t iacc_ext = 0;
    int32_t Sdg = Sdg0;
    STITCH_UNROLL
    for (int k = 0; k < STRIP; ++k) {
        const bool ordinary = !SPECIAL || k < nv;
        const bool rowm = SPECIAL && has_m && k == nv;
        if (ordinary || rowm) {
            const int32_t cc = (x[k] == c.q) ? c.cM : c.cX;
            // (the floors at NEG keep the traceback variant's I / D pointers canonical; the bulk pass only needs the values that
            // can still win, and an unfloored value stays within (STRIP + 1) extensions of NEG, which pk_plan's range covers)
            const int32_t ext = (TB || PK_FLOORS) ? pk_addmax(Dup[k], c.cE, p.NEGKEY) : pk_add(Dup[k], c.cE);
            const int32_t Dp = pk_addmax(Sup[k], c.cOE, ext);
            const int32_t D6 = (Dp & p.NPM) | p.PD6;
            int32_t jp = Jc + cc;
            uint32_t fl = 0;
            if (TB) fl = (Dp & p.P1) ? 1u : 0u;
            if (SPECIAL && k == 0 && wrap0) {
                const int32_t w = wbase + cc;
                if (w > jp) { jp = w; if (TB) fl |= 4u; }
            }
            o.D6[k] = D6; o.jp[k] = jp;
            o.Inc[k] = Iacc;
            if (TB) { fl |= iacc_ext ? 2u : 0u; o.fl[k] = fl; }
            if (ordinary) {
                const int32_t A = pk_addmax(Sdg, cc, D6);
                o.A[k] = A;
                int32_t H = pk_max3(A, jp | (A & p.PB8), c.XC);
                if (SPECIAL) H = pk_max(H, o.YC[k]);
                const int32_t Hc = H & p.NPM;
                const int32_t exti = (TB || PK_FLOORS) ? pk_addmax(Iacc, c.cEi, p.NEGKEY + p.PI4) : pk_add(Iacc, c.cEi);
                const int32_t Ip = pk_addmax(Hc, c.cOEi, exti);
                if (TB) iacc_ext = ((Ip >> p.LB) & 15) == PP_INC ? 1u : 0u;
                Iacc = (Ip & p.NPM) | p.PI4;
            } else {
                o.A[k] = Sdg + cc;   // row m: the diagonal candidate; finished per contig
            }
        }
        Sdg = Sup[k];
    }
    o.exit = Iacc;
    if (TB) o.exit_open = iacc_ext ? 0u : 1u;
}

// Pass 2 of one lane.  `cin` = chain value arriving at the strip's first row from earlier rows
// (PP_ICARRY), `cin_open` = it opens from the row just before (traceback variant).
// Outputs: S[k] clean keys of the ordinary rows, their running max in `colmax`, packed traceback
// bytes (TB), and for row m the insertion candidate arriving at it.
// `notq` (ordinary tiles of the bulk pass): accumulates (S ^ jump candidate) over the strip; the closed form of a quiet
// cell is the jump candidate with its priority cleared (PkQuiet::bk[s] == (Jc + cM|cX) & NPM == jp & NPM), so the strip's
// S keys are all in the closed form iff (notq & NPM) == 0.
template <bool SPECIAL, bool TB>
SHD void pk_pass2(const PK &p, const PCol &c, const PStrip &s, int32_t cin, uint32_t cin_open, int nv, bool has_m,
                  int32_t *S, int32_t &colmax, uint8_t *tb, int32_t *Iarr, int32_t &I_m, uint32_t &iext_m, int32_t *notq = nullptr) {
    int32_t cx = cin;
    STITCH_UNROLL
    for (int k = 0; k < STRIP; ++k) {
        const bool ordinary = !SPECIAL || k < nv;
        const bool rowm = SPECIAL && has_m && k == nv;
        if (ordinary) {
            const int32_t T1 = pk_max3(s.A[k], s.Inc[k], cx);
            const int32_t jT = s.jp[k] | (T1 & p.PB8);
            int32_t Sp = pk_max3(T1, jT, c.XC);
            if (SPECIAL) Sp = pk_max(Sp, s.YC[k]);
            const int32_t Sc = Sp & p.NPM;
            S[k] = Sc;
            colmax = pk_max(colmax, Sc);
            if (!SPECIAL && !TB && notq) *notq |= Sp ^ s.jp[k];
            if (TB) {
                const uint32_t pr = (uint32_t)(Sp >> p.LB) & 15u;
                uint32_t mv;
                if (pr == PP_DIAG) mv = (Sp == jT && jT != T1) ? ((s.fl[k] & 4u) ? MV_WRAP : MV_JUMP) : MV_DIAG;
                else if (pr == PP_D) mv = MV_DEL;
                else if (pr == PP_INC || pr == PP_ICARRY) mv = MV_INS;
                else if (pr == PP_JMP) mv = (s.fl[k] & 4u) ? MV_WRAP : MV_JUMP;
                else if (pr == PP_XC) mv = MV_XCLIP_PREFIX;
                else mv = MV_YCLIP_PREFIX;
                // the I pointer of this cell: the carried value wins ties against the strip's own chain
                const bool from_carry = (cx >> p.SH) >= (s.Inc[k] >> p.SH);
                const uint32_t iext = from_carry ? ((k == 0) ? (cin_open ? 0u : 1u) : 1u) : ((s.fl[k] & 2u) ? 1u : 0u);
                tb[k] = (uint8_t)(mv | ((s.fl[k] & 1u) ? TBB_DEXT : 0u) | (iext ? TBB_IEXT : 0u));
                Iarr[k] = from_carry ? cx : s.Inc[k];
            }
        } else if (rowm) {
            const bool from_carry = (cx >> p.SH) >= (s.Inc[k] >> p.SH);
            I_m = from_carry ? cx : s.Inc[k];
            iext_m = from_carry ? ((k == 0) ? (cin_open ? 0u : 1u) : 1u) : ((s.fl[k] & 2u) ? 1u : 0u);
        }
        cx = (TB || PK_FLOORS) ? pk_addmax(cx, c.cEi, p.NEGKEY + p.PI5) : pk_add(cx, c.cEi);
    }
}

// The wide records (dp_core.h) of one ordinary cell of the traceback variant: y-suffix tracker update
// (SCA:432-447) and the column-n record the end-of-read fix-up edits (SCA:453-555).
SHD void pk_cell_records(const PK &p, const PCol &c, const Scoring &sc, int32_t Skey, int32_t Ikey, uint32_t tbbyte, bool is_match,
                         uint32_t self_idx, uint32_t i, uint32_t m, const JumpInfo &J, uint32_t j, uint32_t n,
                         SnRec *sn_rec, LastCell *last_rec) {
    const uint32_t mv = tbbyte & 15u;
    const int32_t S = pk_abs(p, c.B, Skey);
    const uint32_t sl = pk_len(p, Skey);
    uint32_t idx, from;
    ptr_of_move(mv, self_idx, i, m, J, idx, from);
    if (sn_rec) sn_update(sc, *sn_rec, S, sl, idx, j, n);
    if (last_rec) {
        LastCell lc; lc.S = S; lc.I = pk_abs(p, c.B, Ikey); lc.sl = sl; lc.il = pk_len(p, Ikey); lc.idx = idx; lc.from = from;
        lc.s_tb = (uint8_t)tb_of_move(mv, is_match); lc.i_tb = 0;
        lc.flags = (uint8_t)(((tbbyte & TBB_IEXT) ? 1 : 0) | ((tbbyte & TBB_DEXT) ? 2 : 0)); lc.pad = 0; lc.pad2 = 0;
        *last_rec = lc;
    }
}

// Row m of a contig, finished from the packed candidates once the tracker over rows < m is known
// (the same finish_rowm as the wide path).  stash = {diagonal, D6, jump, I arriving} keys.
struct PkRowM { int32_t diag, D6, jp, I; uint32_t fl, iext; };   // fl: bit0 dext, bit2 wrap won
struct PkRowMOut { RowMOut ro; int32_t skey; uint32_t tbbyte; };
SHD PkRowMOut pk_finish_rowm(const PK &p, const PCol &c, const Scoring &sc, const PkRowM &st, const XsPart &tr, const Row0 &r0,
                             const JumpInfo &J, bool is_match, uint32_t self_idx, uint32_t m, uint32_t j) {
    RowM rm;
    rm.diag = pk_abs(p, c.B, st.diag); rm.dgl = pk_len(p, st.diag);
    rm.D = pk_abs(p, c.B, st.D6); rm.dl = pk_len(p, st.D6); rm.dext = st.fl & 1u;
    rm.I = pk_abs(p, c.B, st.I); rm.il = pk_len(p, st.I); rm.iext = st.iext;
    rm.jp.score = pk_abs(p, c.B, st.jp); rm.jp.len = pk_len(p, st.jp);
    if (st.fl & 4u) { rm.jp.idx = self_idx | 0x80000000u; rm.jp.from = m; }
    else { rm.jp.idx = J.idx; rm.jp.from = J.from; }
    { const int32_t dj = sc.o + sc.e * (int32_t)j; rm.xclip = sc.xp + (sc.yp > dj ? sc.yp : dj); }
    rm.xclip_len = r0.sl;
    rm.yclip = sc.yp + sc.o + sc.e * (int32_t)m; rm.yclip_len = 0;
    rm.is_match = is_match ? 1u : 0u;
    PkRowMOut o;
    o.ro = finish_rowm(sc, rm, tr, self_idx, m);
    o.skey = pk_from_wide(p, c.B, o.ro.c.S, o.ro.c.sl, 0);
    o.tbbyte = o.ro.c.mv | (rm.dext ? TBB_DEXT : 0u) | (rm.iext ? TBB_IEXT : 0u);
    return o;
}

// ---------------------------------------------------------------------------------------------
// Quiet tiles.  Away from the alignment paths every cell of a column takes the jump move
// (MCA:279-331 makes it available everywhere), so its S is J(c) + sub(x_i, y_j): one of TWO keys
// per (contig, column), decided by a byte compare.  A warp tile all of whose cells are in that
// closed form is "quiet"; while a handful of per-(contig, column) integer compares (pk_quiet_next)
// prove that a quiet tile stays quiet, the bulk fill neither loads, computes nor stores it.  The
// state of a quiet tile is re-materialised from the closed form when it is needed again (a column
// that fails the test, a checkpoint column, a neighbour that is not quiet).
//
//   S(r, j) = bk_j[sub(r, j)]
//   D(r, j) = clean6(max_k t_j[k][sub(r, j-1-k)], NEGKEY)   (k = 0 .. PKQ_L-1)   or any DEAD value
//     t[0] = open from S(r, j-1), t[k] = k-fold extension of the open from S(r, j-1-k); older deletion
//     runs are required to be dead.  On equal scores the dense fill keeps the OLDEST run (an extension
//     wins the tie against an open, SCA:329-338), so t[k] carries priority k while the max is taken.
//     A D value v (relative to B_j) is DEAD - can never win or tie again, now or after any number
//     of extensions - when v <= o + gmin + submin - 1: the best cell of column j-1 can itself be
//     extended by deletions, so G(j-1+k) >= G(j-1) + o + e k, and every cell of column j+k has the
//     jump candidate >= G(j-1+k) + gmin + submin (SCA:329-338, 373-382).  Dead values may differ
//     between the closed form and the dense fill; nothing observable does.
//
// sub index: 0 = the bases are equal (match score), 1 = not.
constexpr int PKQ_L = 3;
struct PkQuiet {            // per (contig, column)
    int32_t bk[2];          // clean S key of a quiet cell
    int32_t t[PKQ_L][2];    // raw D candidates (as pass 1 forms them, before clean6; priority field = k)
    int32_t stay;           // tiles quiet at column j-1 (with a quiet predecessor tile) are quiet at column j
    int32_t stay_first;     // ... and so is the FIRST tile of the contig (row 1: diagonal / chain from row 0, wrap, y-clip)
    int32_t why;            // which of C1..C5 failed (bit k-1), bit 5 = not allowed (diagnostics)
};
// What row 1 of a contig sees in column j beyond an ordinary row (SCA:188-239, 258-289, 391-399), as pk_tile forms it.
struct PkFirstIn {
    int32_t r0pkey;         // S(0, j-1): the diagonal source of row 1
    int32_t cr1key;         // insertion chain arriving at row 1 from row 0 (PP_ICARRY)
    int32_t wbase;          // circular wrap source pk_wbase(S(m, j-1)); only read when `wrap`
    int32_t yc1;            // y-prefix clip candidate of row 1 (PP_YC) or NEGKEY
    bool wrap;              // the contig is circular and the S code of (m, j-1) is not XCLIP_SUFFIX
};
SHD int32_t pk_deadrel(const Scoring &sc) {
    int32_t submin = sc.match < sc.mismatch ? sc.match : sc.mismatch;
    int32_t gmin = sc.g_same < sc.g_opp ? sc.g_same : sc.g_opp;
    gmin = gmin < sc.g_inter ? gmin : sc.g_inter;
    return sc.o + gmin + submin - 1;
}
SHD int32_t pk_clean6(const PK &p, int32_t v) { return (pk_max(v, p.NEGKEY) & p.NPM) | p.PD6; }
SHD PkQuiet pk_quiet_init(const PK &p) {
    PkQuiet q;
    q.bk[0] = q.bk[1] = p.NEGKEY; q.stay = 0; q.stay_first = 0; q.why = 0;
    for (int k = 0; k < PKQ_L; ++k) q.t[k][0] = q.t[k][1] = p.NEGKEY;
    return q;
}
// Closed form of column j from the one of column j-1.  `Jc` = pk_jc of the contig for column j.
// `allow` = this column may be skipped at all (not a checkpoint column, read base is one of ACGT, ...).
SHD PkQuiet pk_quiet_next(const PK &p, const Scoring &sc, const PCol &c, int32_t Jc, const PkQuiet &prev, bool allow,
                          const PkFirstIn *fi = nullptr) {
    PkQuiet q;
    const int32_t jM = Jc + c.cM, jX = Jc + c.cX;
    q.bk[0] = jM & p.NPM; q.bk[1] = jX & p.NPM;
    int32_t dcap = p.NEGKEY;
    STITCH_UNROLL
    for (int s = 0; s < 2; ++s) {
        q.t[0][s] = (int32_t)((uint32_t)prev.bk[s] + (uint32_t)c.cOE);
        STITCH_UNROLL
        for (int k = 1; k < PKQ_L; ++k)   // stored D (PP_D) + cE has priority 1; make it k
            q.t[k][s] = (int32_t)((uint32_t)pk_clean6(p, prev.t[k - 1][s]) + (uint32_t)c.cE + (uint32_t)((k - 1) * p.P1));
        STITCH_UNROLL
        for (int k = 0; k < PKQ_L; ++k) dcap = pk_max(dcap, q.t[k][s]);
    }
    const int32_t jlo = pk_max(jM, jX) == jM ? jX : jM;                 // the weaker of the two jump candidates
    const int32_t pbk = pk_max(prev.bk[0], prev.bk[1]);
    // C1 the diagonal from a quiet cell does not beat the jump (equal keys are the same cell value)
    const bool c1 = (int32_t)((uint32_t)pbk + (uint32_t)c.cM) <= (jM | p.PB8) && (int32_t)((uint32_t)pbk + (uint32_t)c.cX) <= (jX | p.PB8);
    // C2 no D of a quiet cell reaches the jump's score (D wins ties)
    const bool c2 = (dcap >> p.SH) < (jlo >> p.SH);
    // C3 x-prefix clip loses (lower priority than the jump: a plain key compare)
    const bool c3 = c.XC < jlo;
    // C4 no insertion chain out of a quiet cell of this column reaches the jump's score
    const bool c4 = (((int32_t)((uint32_t)pk_max(q.bk[0], q.bk[1]) + (uint32_t)c.cOEi)) >> p.SH) < (jlo >> p.SH);
    // C5 the deletion runs the closed form drops (extensions of the oldest term of column j-1) are dead
    const int32_t dropped = (int32_t)((uint32_t)pk_clean6(p, pk_max(prev.t[PKQ_L - 1][0], prev.t[PKQ_L - 1][1])) + (uint32_t)c.cE);
    const bool c5 = (dropped >> p.SH) <= pk_deadrel(sc);
    const bool ok = c1 && c2 && c3 && c4 && c5;
    q.why = (c1 ? 0 : 1) | (c2 ? 0 : 2) | (c3 ? 0 : 4) | (c4 ? 0 : 8) | (c5 ? 0 : 16) | (allow ? 0 : 32);
    q.stay = (ok && allow) ? 1 : 0;
    q.stay_first = 0;
    if (fi && q.stay) {
        // C6 the diagonal from row 0 does not beat the jump in row 1; C7 the chain from row 0 stays below the jump's score;
        // C8 the circular wrap (a second jump candidate of row 1) does not beat the jump; C9 the y-prefix clip loses
        const bool c6 = (int32_t)((uint32_t)fi->r0pkey + (uint32_t)c.cM) <= (jM | p.PB8) && (int32_t)((uint32_t)fi->r0pkey + (uint32_t)c.cX) <= (jX | p.PB8);
        const bool c7 = (fi->cr1key >> p.SH) < (jlo >> p.SH);
        const bool c8 = !fi->wrap || fi->wbase <= Jc;
        const bool c9 = fi->yc1 < jlo;
        q.stay_first = (c6 && c7 && c8 && c9) ? 1 : 0;
    }
    return q;
}
// The closed form of one cell of column j.  s[k] = sub index of the cell's contig base against y_{j-k}, k = 0 .. PKQ_L.
SHD int32_t pk_quiet_S(const PkQuiet &q, int s0) { return q.bk[s0]; }
SHD int32_t pk_quiet_D(const PK &p, const PkQuiet &q, const int *s) {
    int32_t d = q.t[0][s[1]];
    STITCH_UNROLL
    for (int k = 1; k < PKQ_L; ++k) d = pk_max(d, q.t[k][s[k + 1]]);
    return pk_clean6(p, d);
}
// Is a densely computed cell of column j in the closed form?
SHD bool pk_quiet_cell(const PK &p, const PkQuiet &q, int32_t deadrel, int32_t Skey, int32_t D6, const int *s) {
    return Skey == q.bk[s[0]] && (D6 == pk_quiet_D(p, q, s) || (D6 >> p.SH) <= deadrel);
}

// Row m of a contig whose LAST tile is skipped as quiet in column j: the candidates pk_tile would have stashed for
// pk_finish_rowm.  Row m itself is never in the closed form (it accumulates the x-suffix tracker, SCA:407-429), so its
// own S / D keys of column j-1 are kept per contig (SmKey / DmKey); the row above it is a quiet cell; the insertion chain
// arriving from quiet rows is dead (C4), which leaves the outcome of finish_rowm unchanged (a candidate below the jump's
// score loses to the jump or to whatever beats the jump).  Returns the stash; `*dm_new` = D key of (m, j) to keep.
SHD PkRowM pk_quiet_rowm(const PK &p, const PCol &c, const PkQuiet &qprev, int32_t Jc, int32_t SmKey_prev, int32_t DmKey_prev,
                         int s_above_prev /* sub index of row m-1 against y_{j-1} */, bool m_match /* x_m == y_j */, int32_t *dm_new) {
    const int32_t cc = m_match ? c.cM : c.cX;
    const int32_t ext = pk_addmax(DmKey_prev, c.cE, p.NEGKEY);
    const int32_t Dp = pk_addmax(SmKey_prev, c.cOE, ext);
    PkRowM rm;
    rm.D6 = (Dp & p.NPM) | p.PD6;
    rm.diag = (int32_t)((uint32_t)qprev.bk[s_above_prev] + (uint32_t)cc);
    rm.jp = Jc + cc;
    rm.I = p.NEGKEY + p.PI4;
    rm.fl = 0; rm.iext = 0;
    *dm_new = rm.D6;
    return rm;
}

// ---------------------------------------------------------------------------------------------
// Cone re-fill of the walk (TbUnit, dp_core.h).  The walk enters a (contig, block of columns) unit at cell (i_entry, j)
// and only moves up / left inside it.  A cell depends on the cell diagonally above-left, on D of the same row, and on the
// insertion chain of at most R = band / |e| + 1 rows above it in its own column, so whatever the rows above
// top(jj) = i_entry - slope * (j - jj), slope = R + 1, hold at column jj can only reach rows above top(jj + 1) at column
// jj + 1 (any IN-BAND value: stale keys of an earlier column are in band, so the chain's reach bound applies to them).
// The path itself stays inside the cone: one row per column on the diagonal, insertion runs no longer than R (a longer
// one scores below the jump that is available everywhere).  Row m (it accumulates the x-suffix tracker over the WHOLE
// column, SCA:407-429) and row 1 of a circular contig (it reads row m, SCA:258-289) need the full column.
// Checked on the CPU emulator against full re-fills (tests/emul, EMUL_CONE_AUDIT) and against the oracle end to end.
constexpr uint32_t PK_CONE_MAX_COLS = 384;   // cone units of up to this many columns get their per-column constants precomputed (shared memory)
struct PkCone { bool on; uint32_t slope, win_lo, win_n; };
SHD PkCone pk_cone_plan(const Scoring &sc, const ContigEntry &en, uint32_t i_entry, uint32_t j, uint32_t jb) {
    PkCone c; c.on = false; c.slope = 0; c.win_lo = 0; c.win_n = en.ntiles;
    if (sc.e >= 0 || i_entry < 1 || i_entry >= en.m || j <= jb) return c;
    c.slope = (uint32_t)(pk_band(sc) / -sc.e) + 2;
    const int64_t top_first = (int64_t)i_entry - (int64_t)c.slope * (int64_t)(j - jb - 1);   // top row of the cone at column jb + 1
    if (en.circular && top_first < 2) return c;
    // the window starts `slope` rows higher: column jb + 1 reads the checkpointed column jb that far above its top row
    const int64_t top_ck = top_first - (int64_t)c.slope;
    const uint32_t r_top = top_ck < 1 ? 1u : (uint32_t)top_ck;
    const uint32_t lo = (r_top - 1) / (uint32_t)TILE, hi = (i_entry - 1) / (uint32_t)TILE;
    c.on = true; c.win_lo = lo; c.win_n = hi - lo + 1;
    return c;
}
// First tile that is computed at column jj of a cone unit entered at (i_entry, j).
SHD uint32_t pk_cone_top_tile(const PkCone &c, uint32_t i_entry, uint32_t j, uint32_t jj) {
    const int64_t top = (int64_t)i_entry - (int64_t)c.slope * (int64_t)(j - jj);
    return top < 1 ? 0u : (uint32_t)((top - 1) / TILE);
}

// Carry into a lane from the previous lane's exit (PP_INC -> PP_ICARRY).
SHD int32_t pk_carry_from_exit(const PK &p, int32_t exit_key) { return exit_key + p.P1; }

// Carry arriving at row 1 of a contig from row 0 (SCA:317-326 with i = 1): always the open from S(0, j).
SHD int32_t pk_carry_row1(const PK &p, const PCol &c, const Scoring &sc, const Row0 &r0) {
    return pk_key(p, (int64_t)r0.S - c.B + sc.o + sc.e, PP_ICARRY, r0.sl + 1);
}

}  // namespace stitch
