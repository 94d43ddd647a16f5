// sweep_core.cuh -- per-lane arithmetic of the column sweep and the traceback walk.
//
// Shared verbatim by the sm_100a kernels (sweep_kernels.cu) and by the host emulator used in the
// CPU test-suite (emu.cpp, tests only), so the algebra can be checked without a GPU.
//
// The recurrence restated (reference: stringdecomposer/src/main.cpp:171-208):
//   H[i][k] = max( H[i][k-1]+del, H[i-1][k]+ins, H[i-1][k-1]+s, J[i]+k*del+s )      k>=1, i>=1
// Shift every cell by its position, Hh[i][k] = H[i][k] - i*ins - k*del, and the moves become
//   left: Hh[i][k-1]          up: Hh[i-1][k]          diag: Hh[i-1][k-1] + s''
//   jump: B[i] + s''          with s'' = s - ins - del and B[i] = J[i] - (i-1)*ins + del,
// i.e. the deletion chain is a pure prefix-max and the insertion move is free.  Cells are stored
// relative to the column base B[i] (so the jump operand is the constant 0) and multiplied by 4 with
// a 2-bit priority tag in the low bits (left/del=3 > up/ins=2 > diag=1 > jump=0), so that a plain
// integer max reproduces the reference's traceback priority del > ins > diag > jump
// (main.cpp:242-253) and the backpointer is (3 - tag) = reference code {0:del,1:ins,2:diag,3:jump}.
// State registers hold Z = 4*rel + 3 ("tag-3 form").
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SD_HD __host__ __device__ __forceinline__
#else
#define SD_HD inline
#endif

namespace sdb {

// --------------------------------------------------------------------------------------------
// Host-side emulation of the DPX instructions, with an optional overflow trap for the tests.
// --------------------------------------------------------------------------------------------
#ifndef __CUDA_ARCH__
struct EmuFlags { static inline bool &overflow() { static thread_local bool f = false; return f; } };
static inline int16_t emu_s16(int v) { if (v < -32768 || v > 32767) EmuFlags::overflow() = true; return (int16_t)v; }
static inline uint32_t emu_pack(int lo, int hi) { return (uint32_t)(uint16_t)emu_s16(lo) | ((uint32_t)(uint16_t)emu_s16(hi) << 16); }
static inline int emu_lo(uint32_t a) { return (int16_t)(a & 0xffffu); }
static inline int emu_hi(uint32_t a) { return (int16_t)(a >> 16); }
static inline int emu_max(int a, int b) { return a > b ? a : b; }
#endif

// Two DP rows per 32-bit register (lo half = forward monomer, hi half = its reverse complement).
struct Packed16 {
    static constexpr int ROWS = 2;             // DP rows per slot
    static constexpr int CELLS_PER_WORD = 8;   // backpointer cells per 32-bit word (per half: 8 x 2 bit)
    static constexpr uint32_t TAGMASK = 0x00030003u;
    static SD_HD uint32_t splat(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }
    static SD_HD int lo(uint32_t a) { return (int)(int16_t)(a & 0xffffu); }
    static SD_HD int hi(uint32_t a) { return (int)(int16_t)(a >> 16); }
    static SD_HD uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) {   // per half: max(a+b, c)   VIADDMNMX.S16x2
#ifdef __CUDA_ARCH__
        return __viaddmax_s16x2(a, b, c);
#else
        return emu_pack(emu_max(emu_lo(a) + emu_lo(b), emu_lo(c)), emu_max(emu_hi(a) + emu_hi(b), emu_hi(c)));
#endif
    }
    static SD_HD uint32_t max2(uint32_t a, uint32_t b) {                 // VIMNMX.S16x2
#ifdef __CUDA_ARCH__
        return __vmaxs2(a, b);
#else
        return emu_pack(emu_max(emu_lo(a), emu_lo(b)), emu_max(emu_hi(a), emu_hi(b)));
#endif
    }
    static SD_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) {     // VIMNMX3.S16x2
#ifdef __CUDA_ARCH__
        return __vimax3_s16x2(a, b, c);
#else
        return max2(max2(a, b), c);
#endif
    }
    static SD_HD uint32_t add(uint32_t a, uint32_t b) {                  // per half add
#ifdef __CUDA_ARCH__
        return __vadd2(a, b);
#else
        return emu_pack(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b));
#endif
    }
    static SD_HD uint32_t tag3(uint32_t a) { return a | TAGMASK; }
};

// One DP row per 32-bit register; used when the score range does not provably fit 14+2 bits.
struct Scalar32 {
    static constexpr int ROWS = 1;
    static constexpr int CELLS_PER_WORD = 16;
    static constexpr uint32_t TAGMASK = 3u;
    static SD_HD uint32_t splat(int v) { return (uint32_t)v; }
    static SD_HD int lo(uint32_t a) { return (int)a; }
    static SD_HD int hi(uint32_t a) { return (int)a; }
    static SD_HD uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) {   // VIADDMNMX
#ifdef __CUDA_ARCH__
        return (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);
#else
        int s = (int)a + (int)b; return (uint32_t)(s > (int)c ? s : (int)c);
#endif
    }
    static SD_HD uint32_t max2(uint32_t a, uint32_t b) { return (uint32_t)((int)a > (int)b ? (int)a : (int)b); }
    static SD_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
        return (uint32_t)__vimax3_s32((int)a, (int)b, (int)c);
#else
        return max2(max2(a, b), c);
#endif
    }
    static SD_HD uint32_t add(uint32_t a, uint32_t b) { return a + b; }
    static SD_HD uint32_t tag3(uint32_t a) { return a | 3u; }
};

// Per-column uniform operands of one segment.
struct ColumnConsts {
    uint32_t addc_dg;   // splat(-2 - 4*delta): tag-3 form -> diag candidate (tag 1) in this column's base
    uint32_t addc_up;   // splat(-1 - 4*delta): tag-3 form -> insertion candidate (tag 2)
    uint32_t jump0;     // the jump operand: 0 relative to the column base.  Passed as an opaque run-time zero so
                        // that ptxas keeps it in one register instead of materialising a packed zero per cell.
};

template <class P> SD_HD ColumnConsts make_column_consts(int delta, uint32_t zero)
{
    ColumnConsts c;
    c.jump0 = zero;
    c.addc_dg = P::splat(-2 - 4 * delta);
    c.addc_up = P::splat(-1 - 4 * delta);
    return c;
}

// Pass 1 of a column for one lane: the three chain-free candidates of its C cells.
//   X[kk] in : Z (tag-3 form) of the previous column          X[kk] out: m2 = max(up, diag, jump), tagged
//   prevZ    : previous-column Z of the cell left of X[0] (the left lane's X[C-1]; DEADZ for lane 0 of a slot)
//   prof     : C profile words 4*s'' (packed per half) for this column's read symbol, prof[kk*stride]
//   kill_first / kill_last: the cell is a k==0 cell -- no insertion candidate (main.cpp:194)
//   adj_first / adj_last  : added to the profile word of X[0] / X[C-1] (column 0 only, main.cpp:173-177 vs :180)
// Returns the lane's chain end max_k(m2[k]) in tag-3 form (what the lane hands to the lanes on its right).
template <class P, int C>
SD_HD uint32_t lane_pass1(uint32_t (&X)[C], uint32_t prevZ, const uint32_t *prof4, ColumnConsts cc, uint32_t deadz,
                          bool kill_first, bool kill_last, uint32_t adj_first, uint32_t adj_last)
{
    uint32_t prev = prevZ;
    uint32_t E = deadz;
#pragma unroll
    for (int kk = 0; kk < C; ++kk) {
        uint32_t s4 = prof4[kk];
        if (kk == 0) s4 = P::add(s4, adj_first);
        if (kk == C - 1) s4 = P::add(s4, adj_last);
        uint32_t m1 = P::addmax(prev, cc.addc_dg, cc.jump0);     // max(diag', jump'=0)
        uint32_t m1s = P::add(m1, s4);
        uint32_t up = X[kk];
        prev = up;
        if (kk == 0 && kill_first) up = deadz;
        if (kk == C - 1 && kill_last) up = deadz;
        uint32_t m2 = P::addmax(up, cc.addc_up, m1s);
        X[kk] = m2;
        E = P::max2(E, m2);
    }
    return P::tag3(E);
}

// Pass 2: run the deletion chain (prefix max) through the lane with the carry from the lanes on its left,
// leave the new Z in X and emit the 2-bit backpointers (3 - tag) of the C cells.
// Word layout: word wi holds cells [wi*CPW, wi*CPW + ncell); cell c of the word sits at bits 2*(ncell-1-c)
// (Packed16: forward row in bits 0..15, reverse-complement row in bits 16..31).
template <class P, int C>
SD_HD void lane_pass2(uint32_t (&X)[C], uint32_t carryZ, uint32_t *codes)
{
    constexpr int CPW = P::CELLS_PER_WORD;
    uint32_t zl = carryZ;
    uint32_t w = 0;
#pragma unroll
    for (int kk = 0; kk < C; ++kk) {
        uint32_t h = P::max2(zl, X[kk]);
        zl = P::tag3(h);
        w = w * 4u + (zl - h);
        X[kk] = zl;
        if ((kk + 1) % CPW == 0 || kk == C - 1) { codes[kk / CPW] = w; w = 0; }
    }
}

// Jump key of an end cell: (relative end score + (L-1)*del) * 4096 + (4095 - row): a plain integer max
// over the keys yields max score and, on ties, the lowest row (main.cpp:212 strict '<', :230-236 first match).
constexpr int SD_KEY_ROWS = 4096;
SD_HD int make_key(int z_end, int endadd, int row) { return ((z_end >> 2) + endadd) * SD_KEY_ROWS + (SD_KEY_ROWS - 1 - row); }
SD_HD int key_value(int key) { return key >> 12; }
SD_HD int key_row(int key) { return SD_KEY_ROWS - 1 - (key & (SD_KEY_ROWS - 1)); }

// --------------------------------------------------------------------------------------------
// Geometry of one launch: how DP rows map onto lanes and where backpointers live.
// --------------------------------------------------------------------------------------------
struct Geometry {
    int packed;      // 1: Packed16 (slot = forward row + its reverse complement), 0: Scalar32 (slot = one row)
    int C, T;        // cells per lane, lanes per slot; slot length SL = C*T >= longest row
    int nslots;      // slots per segment
    int M;           // forward monomers (rows = 2M)
    int NS;          // segments per CTA
    int NT;          // threads per CTA
    int CW;          // backpointer words per lane per column
};

struct Record { int32_t row, start, end, score; };

// Decode the backpointer of cell k (k>=1... any k inside the row) of `row` in column i of a segment.
// codes_col points at the first word of column i of the CTA that owns the segment.
SD_HD int fetch_code(const uint32_t *codes_col, const Geometry &g, int seg_local, int row, int rowlen, int k)
{
    int slot = g.packed ? (row < g.M ? row : row - g.M) : row;
    int pos = (rowlen == 1) ? (g.C * g.T - 1) : k;     // length-1 rows are right-aligned in their slot
    int t = pos / g.C, kk = pos - t * g.C;
    int cpw = g.packed ? 8 : 16;
    int wi = kk / cpw, c = kk - wi * cpw;
    int ncell = g.C - wi * cpw; if (ncell > cpw) ncell = cpw;
    int tid = (seg_local * g.nslots + slot) * g.T + t;
    uint32_t w = codes_col[(size_t)tid * g.CW + wi];
    int sh = 2 * (ncell - 1 - c) + ((g.packed && row >= g.M) ? 16 : 0);
    return (int)((w >> sh) & 3u);
}

// Traceback of one segment from the 2-bit backpointers (reference main.cpp:217-267).
//   jcol[i], i=1..n : J[i] = max_r H[i-1][r][last]  (jcol[n] is the final score), arow[i] its lowest row.
//   k == 0 cells are decided here from J and the two symbols involved (main.cpp:245 tests the insertion
//   equality at k==0 although the forward pass never takes that move, main.cpp:194).
// Records come out last-to-first (the caller reverses, main.cpp:268).  Returns the count or -1 on overflow.
template <class CodeAt>
SD_HD int traceback_segment(int n, const int *jcol, const int *arow, const uint8_t *seg, const uint8_t *rows,
                            const int *row_off, int ins, int del, int mismatch, int match,
                            CodeAt code_at, Record *out, int cap)
{
    (void)del;
    int cnt = 0;
    int i = n - 1;
    int r = arow[n];
    int last = row_off[r + 1] - row_off[r] - 1;
    int k = last;
    int end = i;
    int end_score = jcol[n];
    for (;;) {
        int code;
        if (k == 0) {
            code = 3;
            if (i != 0) {
                int s_here = rows[row_off[r]] == seg[i] ? match : mismatch;
                int s_prev = rows[row_off[r]] == seg[i - 1] ? match : mismatch;
                int h_here = jcol[i] + s_here;                               // H[i][r][0], main.cpp:191-193
                int h_prev = (i == 1) ? s_prev : jcol[i - 1] + s_prev;       // H[i-1][r][0] (row 0: main.cpp:173-177)
                if (h_here == h_prev + ins) code = 1;
            }
        } else {
            code = code_at(i, r, last + 1, k);
        }
        if (code == 0) { --k; continue; }
        if (code == 1) { --i; continue; }
        if (code == 2) { --i; --k; continue; }
        if (cnt >= cap) return -1;
        Record rec;
        rec.row = r; rec.start = i; rec.end = end;
        if (i == 0) { rec.score = end_score; out[cnt++] = rec; break; }      // main.cpp:258-262
        rec.score = end_score - jcol[i];                                      // main.cpp:253-257
        out[cnt++] = rec;
        end_score = jcol[i];
        r = arow[i];
        --i;
        last = row_off[r + 1] - row_off[r] - 1;
        k = last; end = i;
    }
    return cnt;
}

} // namespace sdb
