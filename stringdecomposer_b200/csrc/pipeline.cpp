// pipeline.cpp -- host driver around the device sweep: FASTA ingest, segmentation, per-device scheduling,
// overlap resolution and the raw TSV writer.  Reference behaviour restated from stringdecomposer/src/main.cpp
// (line numbers cited per function); the data structures and control flow are this project's own.
#include <algorithm>
#include <cerrno>
#include <cstring>
#include <fstream>
#include <chrono>
#include <condition_variable>
#include <future>
#include <mutex>
#include <thread>
#include <sys/stat.h>
#include <unistd.h>

#include "pipeline.h"

namespace sdb {

// ---------------------------------------------------------------------------------------------
// FASTA: name = first whitespace-delimited token of the header (main.cpp:321-325); sequence lines are
// appended verbatim (main.cpp:327) -- the toupper in Seq's constructor only ever sees "" (main.cpp:29,325),
// so lower case and '\r' are illegal symbols; alphabet check and messages main.cpp:330-344.
// ---------------------------------------------------------------------------------------------
int load_fasta(const std::string &path, FastaSet &out, std::string &diag)
{
    out.names.clear(); out.seqs.clear();
    // whole file in one read, lines found with memchr: at GPU speed the reference's getline loop would dominate
    std::string buf;
    if (FILE *f = fopen(path.c_str(), "rb")) {
        char chunk[1 << 16];
        struct stat sb;
        if (fstat(fileno(f), &sb) == 0 && sb.st_size > 0) buf.reserve((size_t)sb.st_size);
        size_t got;
        while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, got);
        fclose(f);
    }
    auto ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    bool has_n = false;
    int bad_seq = -1; char bad_char = 0;
    const char *p = buf.data(), *end = p + buf.size();
    while (p < end) {
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
        const char *le = nl ? nl : end;                     // the line is [p, le), like getline
        if (le > p && *p == '>') {                         // name = first token (main.cpp:321-325)
            const char *a = p + 1;
            while (a < le && ws(*a)) ++a;
            const char *b = a;
            while (b < le && !ws(*b)) ++b;
            out.names.emplace_back(a, b);
            out.seqs.emplace_back();
        } else if (!out.seqs.empty()) {                     // lines appended raw (main.cpp:327), checked at :330-341
            if (bad_seq < 0)
                for (const char *c = p; c < le; ++c) {
                    if (base_code(*c) < 0) { bad_seq = (int)out.seqs.size() - 1; bad_char = *c; break; }
                    has_n |= *c == 'N';
                }
            out.seqs.back().append(p, le);
        }
        p = nl ? nl + 1 : end;
    }
    if (bad_seq >= 0) {
        // the reference reports the first offending sequence in file order, and in it the first offending symbol
        diag += "ERROR: Sequence " + out.names[bad_seq] + " contains undefined symbol (not ACGT): " + std::string(1, bad_char) + "\n";
        return 255;
    }
    if (has_n) diag += "WARNING: sequences in " + path + " contain N symbol. It will be counted as a separate symbol in scoring!\n";
    return 0;
}

// main.cpp:73-79.  The test at :74 is evaluated in size_t (int operands converted), as written there.
int64_t segment_read(int64_t read_len, int part_size, int overlap, std::vector<std::pair<int, int>> *out)
{
    if (part_size <= 0) return -1;
    int64_t cnt = 0;
    const size_t len = (size_t)read_len;
    for (size_t i = 0; i < len; i += (size_t)part_size) {
        bool keep = ((size_t)(int)len - i >= (size_t)overlap) || (len < (size_t)overlap);
        if (!keep) continue;
        int rest = (int)(len - i), want = part_size + overlap;
        int take = std::min(want, rest);
        if (take < 0) take = rest;
        if (out) out->emplace_back((int)i, take);
        ++cnt;
    }
    return cnt;
}

// main.cpp:287-302
void postprocess(const std::vector<Record> &in, std::vector<Record> &out)
{
    out.clear();
    size_t i = 0; const size_t n = in.size();
    while (i < n) {
        for (size_t j = i + 1; j < std::min(i + 7, n); ++j) {
            if ((in[i].end - in[j].start) * 2 > (in[j].end - in[j].start)) {
                out.push_back(in[i]);
                i = j + 1;
                break;
            }
        }
        if (i < n) out.push_back(in[i]);
        ++i;
    }
}

// ---------------------------------------------------------------------------------------------
// One host thread per device, kept for the life of the engine: with kernels in the millisecond range the cost of
// spawning and joining a thread per device and call (tens of microseconds each) would show in the multi-GPU numbers.
class DevicePool {
public:
    explicit DevicePool(int n) : jobs_((size_t)n), state_((size_t)n, 0)
    {
        for (int d = 0; d < n; ++d) th_.emplace_back([this, d] { loop(d); });
    }
    ~DevicePool()
    {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // runs fn(d) for every device d on that device's thread and waits for all of them
    void run(const std::function<void(int)> &fn)
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            for (size_t d = 0; d < jobs_.size(); ++d) { jobs_[d] = &fn; state_[d] = 1; }
            pending_ = (int)jobs_.size();
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

private:
    void loop(int d)
    {
        for (;;) {
            const std::function<void(int)> *fn;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || state_[(size_t)d] == 1; });
                if (stop_) return;
                fn = jobs_[(size_t)d]; state_[(size_t)d] = 2;
            }
            (*fn)(d);
            {
                std::lock_guard<std::mutex> l(mu_);
                state_[(size_t)d] = 0;
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<std::thread> th_;
    std::vector<const std::function<void(int)> *> jobs_;
    std::vector<int> state_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    int pending_ = 0;
    bool stop_ = false;
};

Engine::Engine(const std::vector<std::string> &forward_monomers, const Scoring &sc, std::vector<std::unique_ptr<Backend>> devs)
    : sc_(sc), devs_(std::move(devs))
{
    build_monomer_set(forward_monomers, ms_);
    if (devs_.empty()) throw PlanError{"no device backend"};
    if (devs_.size() > 1) pool_.reset(new DevicePool((int)devs_.size()));
}

Engine::~Engine() {}

void Engine::on_devices(const std::function<void(int)> &fn)
{
    if (ndev() == 1) fn(0);
    else pool_->run(fn);
}

void Engine::set_ed_thr(int ed_thr)
{
    if (ed_thr > -1 && ms_.Lmax > 64 * SD_HW_BLOCKS) throw PlanError{"--ed_thr needs monomers of at most 1536 bp"};
    for (auto &d : devs_) d->set_filter(ed_thr < 0 ? -1 : ed_thr);
}

void Engine::plan_for(const Batch &b)
{
    int maxlen = 0;
    for (int s = 0; s < b.nseg(); ++s) maxlen = std::max(maxlen, b.len(s));
    const int64_t per_dev = (b.nseg() + ndev() - 1) / ndev();
    // re-plan only when the shape changes enough to matter (geometry depends on segment count and length)
    if (have_plan_ && maxlen <= plan_maxlen_ && maxlen * 2 > plan_maxlen_ && per_dev == plan_nseg_) return;
    plan_ = make_plan(ms_, sc_, maxlen, per_dev);
    plan_maxlen_ = maxlen; plan_nseg_ = per_dev; have_plan_ = true;
    for (auto &d : devs_) d->configure(plan_, ms_);
    stats.g = plan_.g;
}

void Engine::note_split(const Batch &b, const std::vector<int> &bounds)
{
    stats.sweep_store_bytes = 0;
    for (int d = 0; d < 8; ++d) stats.dev_segments[d] = 0;
    for (int d = 0; d < ndev(); ++d) {
        if (d < 8) stats.dev_segments[d] = bounds[d + 1] - bounds[d];
        if (bounds[d] == bounds[d + 1]) continue;
        const CtaLayout l = make_cta_layout(plan_, b, bounds[d], bounds[d + 1]);
        stats.sweep_store_bytes += l.cta_code_off.back() * 4 + l.seg_j_off.back() * (int64_t)sizeof(JR);
    }
}

// contiguous, column-balanced ranges, one per device (segments are independent, no collective)
void Engine::split(const Batch &b, std::vector<int> &bounds) const
{
    const int nd = ndev(), nseg = b.nseg();
    bounds.assign(nd + 1, nseg);
    bounds[0] = 0;
    const int64_t total = b.off[nseg] - b.off[0];
    int s = 0;
    for (int d = 1; d < nd; ++d) {
        const int64_t target = b.off[0] + total * d / nd;
        while (s < nseg && b.off[s] < target) ++s;
        int cut = s / plan_.g.NS * plan_.g.NS;           // keep CTAs whole
        bounds[d] = std::max(bounds[d - 1], std::min(cut, nseg));
    }
}

void Engine::decompose(const Batch &b, BatchResult &out)
{
    out.recs.clear(); out.rec_off.assign(1, 0);
    if (b.nseg() == 0) return;
    for (int s = 0; s < b.nseg(); ++s) if (b.len(s) <= 0) throw PlanError{"empty segment"};
    plan_for(b);
    std::vector<int> bounds; split(b, bounds);
    note_split(b, bounds);
    const int nd = ndev();
    std::vector<BatchResult> part(nd);
    std::vector<std::string> errs(nd);
    auto work = [&](int d) {
        try {
            Backend &dev = *devs_[d];
            dev.reset_stats();
            part[d].rec_off.assign(1, 0);
            const int s_end = bounds[d + 1];
            const int nsl = dev.wave_slots();
            // waves: as many segments as fit the budget of one wave slot (with two slots each gets half the memory, and
            // the copies of one wave overlap the kernels of the other)
            const int64_t budget = dev.wave_budget() / nsl;
            const bool prof = getenv("SD_PROFILE") != nullptr;
            std::vector<std::pair<int, int>> waves;
            for (int s0 = bounds[d]; s0 < s_end;) {
                int lo = std::min(s_end, s0 + plan_.g.NS), hi = s_end;
                if (dev.wave_bytes(b, s0, hi) > budget) {
                    while (hi - lo > plan_.g.NS) {
                        int mid = lo + (hi - lo) / 2 / plan_.g.NS * plan_.g.NS;
                        if (mid <= lo) break;
                        if (dev.wave_bytes(b, s0, mid) <= budget) lo = mid; else hi = mid;
                    }
                    hi = lo;
                }
                waves.emplace_back(s0, hi);
                s0 = hi;
            }
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            const auto t0 = now();
            for (size_t w = 0; w < waves.size(); ++w) {
                dev.submit((int)(w % (size_t)nsl), b, waves[w].first, waves[w].second);
                if (nsl == 1) dev.collect(0, part[d]);
                else if (w >= 1) dev.collect((int)((w - 1) % (size_t)nsl), part[d]);
            }
            if (nsl > 1 && !waves.empty()) dev.collect((int)((waves.size() - 1) % (size_t)nsl), part[d]);
            if (prof) fprintf(stderr, "[sd_b200 profile] dev %d: %zu wave(s), segments [%d,%d): %.3f ms host wall (sweep %.3f traceback %.3f h2d %.3f ms on the device)\n",
                              d, waves.size(), bounds[d], s_end, ms(t0, now()), dev.sweep_ms, dev.traceback_ms, dev.h2d_ms);
        } catch (PlanError &e) { errs[d] = e.msg.empty() ? "error" : e.msg; }
        catch (std::exception &e) { errs[d] = e.what(); }
    };
    on_devices(work);
    for (int d = 0; d < nd; ++d) if (!errs[d].empty()) throw PlanError{errs[d]};
    double sw = 0, tb = 0, h2d = 0, d2h = 0;
    for (int d = 0; d < nd; ++d) {
        const int64_t base = (int64_t)out.recs.size();
        out.recs.insert(out.recs.end(), part[d].recs.begin(), part[d].recs.end());
        for (size_t x = 1; x < part[d].rec_off.size(); ++x) out.rec_off.push_back(base + part[d].rec_off[x]);
        sw = std::max(sw, devs_[d]->sweep_ms); tb = std::max(tb, devs_[d]->traceback_ms);
        h2d = std::max(h2d, devs_[d]->h2d_ms); d2h = std::max(d2h, devs_[d]->d2h_ms);
        stats.h2d_bytes += devs_[d]->h2d_bytes; stats.d2h_bytes += devs_[d]->d2h_bytes; stats.launches += devs_[d]->launches;
    }
    stats.sweep_ms += sw; stats.traceback_ms += tb; stats.h2d_ms += h2d; stats.d2h_ms += d2h;
    const int64_t cols = b.off[b.nseg()] - b.off[0];
    stats.columns += cols; stats.segments += b.nseg(); stats.cells += cols * (int64_t)ms_.rows.size();
}

void Engine::stage(const Batch &b)
{
    if (b.nseg() == 0) throw PlanError{"nothing to stage"};
    staged_ = b;
    if (staged_.own.empty() && b.nseg()) staged_.own.assign(b.text, b.text + b.off[b.nseg()]);   // outlive the caller's buffer
    staged_.text = staged_.own.data();
    plan_for(staged_);
    split(staged_, staged_bounds_);
    note_split(staged_, staged_bounds_);
    for (int d = 0; d < ndev(); ++d) {
        if (staged_bounds_[d] == staged_bounds_[d + 1]) continue;
        if (devs_[d]->wave_bytes(staged_, staged_bounds_[d], staged_bounds_[d + 1]) > devs_[d]->wave_budget())
            throw PlanError{"staged batch does not fit one wave on the device"};
    }
    std::vector<std::string> errs(ndev());
    auto work = [&](int d) {
        try { if (staged_bounds_[d] < staged_bounds_[d + 1]) devs_[d]->stage(staged_, staged_bounds_[d], staged_bounds_[d + 1]); }
        catch (PlanError &e) { errs[d] = e.msg.empty() ? "error" : e.msg; }
        catch (std::exception &e) { errs[d] = e.what(); }
    };
    on_devices(work);
    for (auto &e : errs) if (!e.empty()) throw PlanError{e};
}

double Engine::run_staged()
{
    if (staged_.nseg() == 0) throw PlanError{"nothing staged"};
    std::vector<std::string> errs(ndev());
    auto work = [&](int d) {
        try {
            devs_[d]->reset_stats();
            if (staged_bounds_[d] < staged_bounds_[d + 1]) devs_[d]->execute();
        } catch (PlanError &e) { errs[d] = e.msg.empty() ? "error" : e.msg; }
        catch (std::exception &e) { errs[d] = e.what(); }
    };
    on_devices(work);
    for (auto &e : errs) if (!e.empty()) throw PlanError{e};
    double ms = 0, sw = 0, tb = 0;
    for (auto &d : devs_) { ms = std::max(ms, d->sweep_ms + d->traceback_ms); sw = std::max(sw, d->sweep_ms); tb = std::max(tb, d->traceback_ms); stats.launches += d->launches; }
    stats.sweep_ms += sw; stats.traceback_ms += tb;
    const int64_t cols = staged_.off[staged_.nseg()] - staged_.off[0];
    stats.columns += cols; stats.segments += staged_.nseg(); stats.cells += cols * (int64_t)ms_.rows.size();
    return ms;
}

void Engine::fetch_staged(BatchResult &out)
{
    out.recs.clear(); out.rec_off.assign(1, 0);
    for (int d = 0; d < ndev(); ++d) {
        if (staged_bounds_[d] == staged_bounds_[d + 1]) continue;
        devs_[d]->fetch(out);
    }
}

// ---------------------------------------------------------------------------------------------
namespace {

struct FdWriter {
    int fd; std::string buf;
    explicit FdWriter(int f) : fd(f) { buf.reserve(1 << 20); }
    bool failed = false;             // a write() error other than EINTR: the output is incomplete (ENOSPC, EPIPE, ...)
    bool flush()
    {
        if (fd < 0) return true;
        size_t o = 0;
        while (o < buf.size()) {
            ssize_t w = ::write(fd, buf.data() + o, buf.size() - o);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) { failed = true; break; }
            o += (size_t)w;
        }
        buf.clear();
        return !failed;
    }
    void add(const std::string &s) { buf += s; if (buf.size() > (1 << 20)) flush(); }
    void add_int(long v) { char t[24]; int n = 0; bool neg = v < 0; unsigned long u = neg ? 0ul - (unsigned long)v : (unsigned long)v;
        do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u); if (neg) t[n++] = '-'; while (n) buf.push_back(t[--n]); }
    ~FdWriter() { flush(); }
};

} // namespace

int run_files(const std::string &reads_path, const std::string &monomers_path, int threads, int part_size, int overlap,
              const Scoring &sc, int ed_thr, DeviceOpener open_devices, int out_fd, int err_fd, std::string &error)
{
    (void)threads;     // OpenMP width in the reference (main.cpp:85,88); the output does not depend on it
    FdWriter err(err_fd);
    std::vector<std::unique_ptr<Backend>> devs;
    std::string dev_error;
    std::future<int> dev_ready = std::async(std::launch::async, [&] { return open_devices(devs, dev_error); });
    err.add("Scores: insertion=" + std::to_string(sc.ins) + " deletion=" + std::to_string(sc.del) + " mismatch=" +
            std::to_string(sc.mismatch) + " match=" + std::to_string(sc.match) + "\n");                  // main.cpp:393
    err.flush();
    const bool prof = getenv("SD_PROFILE") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto tms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_begin = tnow();
    FastaSet reads, mons;
    std::string diag;
    int st = load_fasta(reads_path, reads, diag);
    err.add(diag); err.flush(); diag.clear();
    if (st) return st;
    st = load_fasta(monomers_path, mons, diag);
    err.add(diag); err.flush();
    if (st) return st;
    if (part_size <= 0) { error = "part-size must be positive"; err.add("ERROR: " + error + "\n"); return 1; }

    // segmentation of all reads, in read order (main.cpp:70-81)
    std::vector<int64_t> first(reads.seqs.size() + 1, 0);
    std::vector<std::pair<int, int>> segs;       // (offset in read, length)
    std::vector<int> seg_read;
    for (size_t p = 0; p < reads.seqs.size(); ++p) {
        first[p] = (int64_t)segs.size();
        size_t before = segs.size();
        segment_read((int64_t)reads.seqs[p].size(), part_size, overlap, &segs);
        for (size_t s = before; s < segs.size(); ++s) seg_read.push_back((int)p);
    }
    first[reads.seqs.size()] = (int64_t)segs.size();
    err.add("Prepared reads\n"); err.flush();                                                                    // main.cpp:82
    if (segs.empty()) return 0;
    if (mons.seqs.empty()) { error = "no monomers"; err.add("ERROR: " + error + "\n"); return 1; }

    const auto t_loaded = tnow();
    // pack the segments while the devices are still being opened
    Batch b;
    b.off.reserve(segs.size() + 1); b.off.push_back(0);
    {
        size_t total = 0;
        for (auto &s : segs) total += (size_t)s.second;
        b.own.resize(total);
        size_t o = 0;
        for (size_t s = 0; s < segs.size(); ++s) {
            const std::string &r = reads.seqs[seg_read[s]];
            memcpy(b.own.data() + o, r.data() + segs[s].first, (size_t)segs[s].second);
            o += (size_t)segs[s].second; b.off.push_back((int64_t)o);
        }
        b.text = b.own.data();
    }
    const auto t_batch = tnow();
    auto t_engine = t_batch, t_done = t_batch;
    if (int dst = dev_ready.get()) {                 // no usable device: fail loudly, there is no CPU path
        error = dev_error;
        err.add("ERROR: " + error + "\n");
        return dst;
    }
    BatchResult res;
    std::unique_ptr<Engine> eng;
    try {
        eng.reset(new Engine(mons.seqs, sc, std::move(devs)));
        t_engine = tnow();
        eng->set_ed_thr(ed_thr);          // FilterMonomersForRead (main.cpp:91-93,135-149) when ed_thr > -1
        eng->decompose(b, res);
        t_done = tnow();
        if (getenv("SD_VERBOSE")) {
            const EngineStats &s = eng->stats;
            char line[512];
            snprintf(line, sizeof line, "[sd_b200] devices=%d geometry packed=%d C=%d T=%d NS=%d NT=%d segments=%ld cells=%ld sweep=%.3f ms traceback=%.3f ms -> %.1f GCUPS (kernels)\n",
                     eng->ndev(), s.g.packed, s.g.C, s.g.T, s.g.NS, s.g.NT, (long)s.segments, (long)s.cells, s.sweep_ms, s.traceback_ms,
                     s.cells / ((s.sweep_ms + s.traceback_ms) * 1e6 + 1e-9));
            err.add(line);
        }
    } catch (PlanError &e) {
        error = e.msg;
        err.add("ERROR: " + error + "\n");
        return 3;
    }
    // The `dp` process exits right after this call: its device buffers need no orderly release (dp_main.cpp).
    if (getenv("SD_FAST_EXIT")) (void)eng.release();

    // per read: add segment offsets (main.cpp:110), PostProcessing (:116), SaveBatch (:117, :272-285); reads are
    // formatted in contiguous chunks on a few host threads and written in order
    const int M = (int)mons.seqs.size();
    const size_t nreads = reads.seqs.size();
    const int nthr = (int)std::max<size_t>(1, std::min<size_t>({8, std::thread::hardware_concurrency(), (res.recs.size() >> 14) + 1}));
    std::vector<std::string> out_part(nthr), err_part(nthr);
    auto format_reads = [&](int t) {
        std::string &ob = out_part[t], &eb = err_part[t];
        FdWriter w(-1);
        std::vector<Record> all, kept;
        // chunk borders balanced by segment count
        const int64_t s_lo = (int64_t)segs.size() * t / nthr, s_hi = (int64_t)segs.size() * (t + 1) / nthr;
        for (size_t p = 0; p < nreads; ++p) {
            if (first[p] < s_lo || first[p] >= s_hi) continue;
            all.clear();
            for (int64_t s = first[p]; s < first[p + 1]; ++s)
                for (int64_t x = res.rec_off[s]; x < res.rec_off[s + 1]; ++x) {
                    Record r = res.recs[x];
                    r.start += segs[s].first; r.end += segs[s].first;
                    all.push_back(r);
                }
            if (all.empty()) continue;     // the reference dereferences batch[0] of an empty vector here (UB)
            eb += std::to_string((p + 1) * 100 / nreads) + "%: Aligned " + reads.names[p] + "\n";            // main.cpp:115
            postprocess(all, kept);
            int prev_end = 0;
            for (const Record &r : kept) {
                w.buf += reads.names[p]; w.buf.push_back('\t');
                w.buf += mons.names[r.row < M ? r.row : r.row - M];
                if (r.row >= M) w.buf.push_back('\'');
                w.buf.push_back('\t'); w.add_int(r.start);
                w.buf.push_back('\t'); w.add_int(r.end);
                w.buf.push_back('\t'); w.add_int(r.score); w.buf += ".000000";      // to_string(float), main.cpp:279
                w.buf.push_back('\t'); w.add_int(r.start - prev_end);
                w.buf.push_back('\t'); w.add_int(r.end - r.start);
                w.buf.push_back('\n');
                prev_end = r.end;
            }
        }
        ob.swap(w.buf);
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < nthr; ++t) pool.emplace_back(format_reads, t);
        format_reads(0);
        for (auto &th : pool) th.join();
    }
    FdWriter outw(out_fd);
    for (int t = 0; t < nthr; ++t) {
        err.add(err_part[t]); err.flush();
        outw.buf.swap(out_part[t]);
        if (!outw.flush()) {                      // ENOSPC / EPIPE ...: a truncated raw TSV must not look like success
            error = std::string("writing the raw decomposition failed: ") + strerror(errno);
            err.add("ERROR: " + error + "\n");
            return 5;
        }
    }
    if (prof) {
        char line[256];
        snprintf(line, sizeof line, "[sd_b200 profile] run_files: fasta %.1f batch %.1f wait-for-device %.1f decompose %.1f output %.1f ms (%d threads)\n",
                 tms(t_begin, t_loaded), tms(t_loaded, t_batch), tms(t_batch, t_engine), tms(t_engine, t_done), tms(t_done, tnow()), nthr);
        err.add(line);
    }
    return 0;
}

} // namespace sdb
