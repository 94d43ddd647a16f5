// emu.cpp -- host emulation of the sweep and traceback kernels.  CPU TEST-SUITE ONLY.
//
// Runs the very same per-lane functions as the sm_100a kernels (sweep_core.cuh) with the CTA-level
// choreography (shuffles, prefix-max scan across the lanes of a slot, key exchange, barrier) replaced by
// loops over an array of lane states.  It is linked into libsd_emu.so, which only tests/ load; the product
// library libsd_b200.so does not contain it and fails loudly without a CUDA device.
#include <algorithm>
#include <climits>
#include <cstring>

#include "common.h"

namespace sdb {

namespace {

template <class P, int C, int T>
void emu_cta(const Plan &p, const Batch &b, int seg_first, int nseg_cta, int nmax,
             uint32_t *codes, int *const *jcol, int *const *arow)
{
    const Geometry &g = p.g;
    const int NT = g.NT, NS = g.NS;
    std::vector<uint32_t> Xall((size_t)NT * C, (uint32_t)P::splat(p.deadz));
    std::vector<uint32_t> E(NT), incl(NT), carry(NT), prevZ(NT);
    std::vector<int> Bprev(NS, p.sc.ins), delta(NS, 0);
    std::vector<int> key(NS, INT_MIN), keynext(NS, INT_MIN);
    const uint32_t deadz = P::splat(p.deadz);
    uint32_t (*X)[C] = reinterpret_cast<uint32_t (*)[C]>(Xall.data());
    uint32_t prof4[C];
    uint32_t cw[8];
    for (int i = 0; i <= nmax; ++i) {
        for (int s = 0; s < nseg_cta; ++s) {
            if (i >= 1) {
                int vmax = key_value(key[s]);
                if (i <= b.len(seg_first + s)) {
                    jcol[s][i] = vmax + Bprev[s] + (i - 1) * p.sc.ins;
                    arow[s][i] = key_row(key[s]);
                }
                delta[s] = vmax + p.sc.del;
                Bprev[s] += delta[s];
            }
        }
        if (i == nmax) break;
        std::fill(keynext.begin(), keynext.end(), INT_MIN);
        // "shuffle": previous-column Z of the cell to the left of each lane's first cell
        for (int tid = 0; tid < NT; ++tid) {
            int t = tid % T;
            prevZ[tid] = (t == 0) ? deadz : X[tid - 1][C - 1];
        }
        for (int tid = 0; tid < NT; ++tid) {
            int ginst = tid / T, t = tid % T;
            int seg_local = ginst / g.nslots, slot = ginst % g.nslots;
            bool lane_ok = seg_local < NS;
            if (!lane_ok) { seg_local = 0; }
            int sl = slot * T + t;
            int n = (seg_local < nseg_cta) ? b.len(seg_first + seg_local) : 0;
            int sym = (i < n) ? b.bases[b.off[seg_first + seg_local] + i] : 0;
            for (int kk = 0; kk < C; ++kk)
                prof4[kk] = p.prof[(((size_t)sym * (C / 4) + kk / 4) * p.nsl + sl) * 4 + (kk % 4)];
            const int L = p.slot_len[slot];
            bool kill_first = (t == 0), kill_last = (t == T - 1 && L == 1);
            uint32_t adj_first = (i == 0 && t == 0 && L > 1) ? P::splat(4 * p.sc.del) : 0u;
            uint32_t adj_last = (i == 0 && t == T - 1 && L == 1) ? P::splat(4 * p.sc.del) : 0u;
            ColumnConsts cc = make_column_consts<P>(i == 0 ? 0 : delta[seg_local], 0u);
            E[tid] = lane_pass1<P, C>(X[tid], prevZ[tid], prof4, cc, deadz, kill_first, kill_last, adj_first, adj_last);
        }
        // inclusive prefix max inside each slot (Kogge-Stone on the device), then exclusive carry
        for (int tid = 0; tid < NT; ++tid) {
            int t = tid % T;
            incl[tid] = (t == 0) ? E[tid] : P::max2(incl[tid - 1], E[tid]);
        }
        for (int tid = 0; tid < NT; ++tid) carry[tid] = (tid % T == 0) ? deadz : incl[tid - 1];
        for (int tid = 0; tid < NT; ++tid) {
            int ginst = tid / T, t = tid % T;
            int seg_local = ginst / g.nslots, slot = ginst % g.nslots;
            lane_pass2<P, C>(X[tid], carry[tid], cw);
            bool active = seg_local < nseg_cta;
            int n = active ? b.len(seg_first + seg_local) : 0;
            if (active && i < n)
                for (int w = 0; w < g.CW; ++w) codes[((size_t)i * NT + tid) * g.CW + w] = cw[w];
            if (active && t == T - 1) {
                uint32_t z = X[tid][C - 1];
                int k0 = make_key(P::lo(z), p.slot_endadd[slot], slot);
                if (P::ROWS == 2) k0 = std::max(k0, make_key(P::hi(z), p.slot_endadd[slot], g.M + slot));
                keynext[seg_local] = std::max(keynext[seg_local], k0);
            }
        }
        key = keynext;
    }
}

template <class P> using CtaFn = void (*)(const Plan &, const Batch &, int, int, int, uint32_t *, int *const *, int *const *);

template <class P, int C> CtaFn<P> pick_t(int T)
{
    switch (T) {
    case 1: return emu_cta<P, C, 1>; case 2: return emu_cta<P, C, 2>; case 4: return emu_cta<P, C, 4>;
    case 8: return emu_cta<P, C, 8>; case 16: return emu_cta<P, C, 16>; case 32: return emu_cta<P, C, 32>;
    }
    return nullptr;
}
template <class P> CtaFn<P> pick(int C, int T)
{
    switch (C) {
    case 8: return pick_t<P, 8>(T); case 16: return pick_t<P, 16>(T); case 24: return pick_t<P, 24>(T);
    case 32: return pick_t<P, 32>(T); case 48: return pick_t<P, 48>(T);
    }
    return nullptr;
}

class EmuBackend : public Backend {
public:
    const char *name() const override { return "emu"; }
    void configure(const Plan &p, const MonomerSet &ms) override { plan_ = p; ms_ = ms; }
    int64_t wave_bytes(const Batch &b, int s0, int s1) const override
    {
        CtaLayout l = make_cta_layout(plan_, b, s0, s1);
        return l.cta_code_off.back() * 4 + l.seg_j_off.back() * 8 + l.seg_rec_off.back() * 16;
    }
    int64_t wave_budget() const override
    {
        if (const char *e = getenv("SD_WAVE_BYTES")) return atoll(e);
        return (int64_t)1 << 30;
    }
    void stage(const Batch &b, int s0, int s1) override { batch_ = &b; s0_ = s0; s1_ = s1; lay_ = make_cta_layout(plan_, b, s0, s1); }
    void execute() override
    {
        const Geometry &g = plan_.g;
        const Batch &b = *batch_;
        codes_.assign((size_t)lay_.cta_code_off.back(), 0u);
        jcol_.assign((size_t)lay_.seg_j_off.back(), 0); arow_.assign((size_t)lay_.seg_j_off.back(), 0);
        const int nseg = s1_ - s0_, nctas = (int)lay_.cta_nmax.size();
        EmuFlags::overflow() = false;
        for (int c = 0; c < nctas; ++c) {
            int first = c * g.NS, cnt = std::min(g.NS, nseg - first);
            std::vector<int *> jp(g.NS, nullptr), ap(g.NS, nullptr);
            for (int s = 0; s < cnt; ++s) { jp[s] = jcol_.data() + lay_.seg_j_off[first + s]; ap[s] = arow_.data() + lay_.seg_j_off[first + s]; }
            uint32_t *codes = codes_.data() + lay_.cta_code_off[c];
            if (g.packed) pick<Packed16>(g.C, g.T)(plan_, b, s0_ + first, cnt, lay_.cta_nmax[c], codes, jp.data(), ap.data());
            else pick<Scalar32>(g.C, g.T)(plan_, b, s0_ + first, cnt, lay_.cta_nmax[c], codes, jp.data(), ap.data());
        }
        overflowed_ = g.packed && EmuFlags::overflow();
        launches += 2;
        // traceback
        recs_.assign((size_t)lay_.seg_rec_off.back(), Record{});
        cnt_.assign(nseg, 0);
        for (int s = 0; s < nseg; ++s) {
            const int cta = s / g.NS, seg_local = s % g.NS, n = b.len(s0_ + s);
            if (n == 0) continue;
            const uint32_t *cbase = codes_.data() + lay_.cta_code_off[cta];
            auto code_at = [&](int i, int row, int rowlen, int k) {
                return fetch_code(cbase + (size_t)i * g.NT * g.CW, g, seg_local, row, rowlen, k);
            };
            cnt_[s] = traceback_segment(n, jcol_.data() + lay_.seg_j_off[s], arow_.data() + lay_.seg_j_off[s],
                                        b.bases.data() + b.off[s0_ + s], ms_.rows.data(), ms_.row_off.data(),
                                        plan_.sc.ins, plan_.sc.del, plan_.sc.mismatch, plan_.sc.match, code_at,
                                        recs_.data() + lay_.seg_rec_off[s], n);
        }
    }
    void fetch(BatchResult &out) override
    {
        if (overflowed_) throw PlanError{"emulator: 16-bit overflow in the packed sweep (range proof violated)"};
        const int nseg = s1_ - s0_;
        for (int s = 0; s < nseg; ++s) {
            if (cnt_[s] < 0) throw PlanError{"emulator: traceback record overflow"};
            const Record *r = recs_.data() + lay_.seg_rec_off[s];
            for (int x = cnt_[s] - 1; x >= 0; --x) out.recs.push_back(r[x]);
            out.rec_off.push_back((int64_t)out.recs.size());
        }
    }

private:
    Plan plan_; MonomerSet ms_;
    const Batch *batch_ = nullptr; int s0_ = 0, s1_ = 0;
    CtaLayout lay_;
    std::vector<uint32_t> codes_; std::vector<int> jcol_, arow_; std::vector<Record> recs_; std::vector<int> cnt_;
    bool overflowed_ = false;
};

} // namespace

Backend *make_emu_backend() { return new EmuBackend(); }

} // namespace sdb
