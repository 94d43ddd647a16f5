// convert.cpp -- sd_convert: the reference's final-TSV stage (stringdecomposer/main.py:96-184: convert_tsv, print_read,
// convert_read, classify) for a whole raw decomposition in one call.  Same output bytes as the Python mirror in
// stringdecomposer_b200/convert.py (which keeps the reference's function-level API); this is the path the command
// line uses, because at GPU speed the per-line Python work (parsing, slicing, str.format) is what the stage costs.
// All alignments go through sd_identity (identity_kernels.cu); nothing here computes an alignment on the host.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <unistd.h>

#include "../../include/sd_b200.h"

namespace {

// stringdecomposer/models/ont_logreg_model.txt as read at main.py:22-26: intercept, identity, identity - second best
const double kLogReg[3] = {-31.48494996, 0.41784018, 0.69186882};
const int64_t kMaxPairsPerCall = int64_t(1) << 24;
const int64_t kMaxLinesPerCall = int64_t(1) << 21;      // bounds the packed intervals of one call (~0.4 GB for 171-bp monomers)

struct View { const char *p; size_t n; std::string str() const { return std::string(p, n); } };

bool py_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= '\x1c' && c <= '\x1f'); }

// `x.split()[0]` (main.py:175-176): the first blank-delimited word; false if there is none (IndexError there)
bool first_word(View v, View &out)
{
    size_t a = 0;
    while (a < v.n && py_space(v.p[a])) ++a;
    size_t b = a;
    while (b < v.n && !py_space(v.p[b])) ++b;
    out = View{v.p + a, b - a};
    return b > a;
}

// int(x) for the plain decimal text `dp` prints (optional sign, digits, blanks around)
bool py_int(View v, long long &out)
{
    size_t a = 0, b = v.n;
    while (a < b && py_space(v.p[a])) ++a;
    while (b > a && py_space(v.p[b - 1])) --b;
    bool neg = false;
    if (a < b && (v.p[a] == '+' || v.p[a] == '-')) { neg = v.p[a] == '-'; ++a; }
    if (a >= b || b - a > 18) return false;
    long long x = 0;
    for (; a < b; ++a) {
        if (v.p[a] < '0' || v.p[a] > '9') return false;
        x = x * 10 + (v.p[a] - '0');
    }
    out = neg ? -x : x;
    return true;
}

struct Writer {
    int fd; std::string buf;
    explicit Writer(int f) : fd(f) { buf.reserve(1 << 20); }
    bool flush()
    {
        size_t o = 0;
        while (o < buf.size()) { ssize_t w = ::write(fd, buf.data() + o, buf.size() - o); if (w <= 0) return false; o += (size_t)w; }
        buf.clear();
        return true;
    }
    void add(View v) { buf.append(v.p, v.n); }
    void add(const std::string &s) { buf += s; }
    void add_int(long long v) { char t[24]; int n = snprintf(t, sizeof t, "%lld", v); buf.append(t, (size_t)n); }
    void add_f2(double v) { char t[40]; int n = snprintf(t, sizeof t, "%.2f", v); buf.append(t, (size_t)n); }   // "{:.2f}".format
    bool maybe_flush() { return buf.size() < (1 << 20) || flush(); }
};

struct Line { View read, mono; long long start, end; int read_idx; int mono_idx; };

thread_local std::string g_convert_error;

int fail(int st, const std::string &msg) { g_convert_error = msg; return st; }

// x[lo:hi] with Python's slice rules on a sequence of length len
void py_slice(long long len, long long lo, long long hi, long long &a, long long &b)
{
    if (lo < 0) lo = std::max(0ll, lo + len);
    if (hi < 0) hi = std::max(0ll, hi + len);
    a = std::min(lo, len); b = std::min(hi, len);
    if (b < a) b = a;
}

// convert_to_homo (main.py:88-93) of every sequence of a packed batch
void collapse(const std::string &blob, const std::vector<int64_t> &off, std::string &oblob, std::vector<int64_t> &ooff)
{
    oblob.clear(); ooff.assign(1, 0);
    oblob.reserve(blob.size());
    for (size_t s = 0; s + 1 < off.size(); ++s) {
        for (int64_t x = off[s]; x < off[s + 1]; ++x)
            if (x == off[s] || blob[(size_t)x] != blob[(size_t)x - 1]) oblob.push_back(blob[(size_t)x]);
        ooff.push_back((int64_t)oblob.size());
    }
}

char classify(double score, double second)          // main.py:96-105
{
    const double z = 1.0 * kLogReg[0] + score * kLogReg[1] + (score - second) * kLogReg[2];
    return z > 0 ? '+' : '?';
}

} // namespace

extern "C" const char *sd_convert_error(void) { return g_convert_error.c_str(); }

extern "C" int sd_convert(const char *raw, int64_t raw_len,
                          const char *read_names, const int64_t *read_name_off, const char *reads, const int64_t *read_off, int64_t n_reads,
                          const char *mono_names, const int64_t *mono_name_off, const char *monos, const int64_t *mono_off, int32_t n_monomers,
                          int32_t min_identity, int32_t light, int32_t device, int out_fd, int alt_fd, sd_convert_stats *stats)
{
    if (stats) memset(stats, 0, sizeof *stats);
    if (raw_len < 0 || n_reads < 0 || n_monomers < 0 || (raw_len && !raw) || !read_name_off || !read_off || !mono_name_off || !mono_off)
        return fail(SD_ERR_ARG, "sd_convert: bad arguments");
    // reads: {id: sequence}; a repeated id is refused like Bio.SeqIO.to_dict does (main.py:65-68)
    std::unordered_map<std::string, int> read_idx;
    for (int64_t r = 0; r < n_reads; ++r) {
        std::string nm(read_names + read_name_off[r], (size_t)(read_name_off[r + 1] - read_name_off[r]));
        if (!read_idx.emplace(nm, (int)r).second) return fail(SD_ERR_ARG, "Duplicate key '" + nm + "'");
    }
    // monomers in add_rc_monomers() order; one trailing '*' of a sequence is dropped by aai() (main.py:38-41)
    const int nm = n_monomers;
    std::vector<std::string> mname((size_t)nm);
    std::string tblob; std::vector<int64_t> toff(1, 0);
    for (int m = 0; m < nm; ++m) {
        mname[(size_t)m].assign(mono_names + mono_name_off[m], (size_t)(mono_name_off[m + 1] - mono_name_off[m]));
        int64_t a = mono_off[m], b = mono_off[m + 1];
        if (b > a && monos[b - 1] == '*') --b;
        tblob.append(monos + a, (size_t)(b - a));
        toff.push_back((int64_t)tblob.size());
    }
    // `scores` of main.py:123-126 is a dict: a repeated name keeps its first position and its last value
    std::unordered_map<std::string, int> last_of, upos;
    std::vector<int> ucol; std::vector<const std::string *> uniq;
    for (int m = 0; m < nm; ++m) {
        auto it = upos.find(mname[(size_t)m]);
        if (it == upos.end()) { upos.emplace(mname[(size_t)m], (int)uniq.size()); uniq.push_back(&mname[(size_t)m]); ucol.push_back(m); }
        else ucol[(size_t)it->second] = m;
        last_of[mname[(size_t)m]] = m;
    }
    const int nu = (int)uniq.size();
    if (!light && nm < 2) return fail(SD_ERR_ARG, "sd_convert: --second-best needs at least two monomers");
    std::string tcblob; std::vector<int64_t> tcoff;
    if (!light) collapse(tblob, toff, tcblob, tcoff);

    // decomposition.split("\n")[:-1] (main.py:173): only newline-terminated lines count
    std::vector<Line> lines;
    {
        const char *p = raw, *end = raw + raw_len;
        std::unordered_map<std::string, int> rcache;
        while (p < end) {
            const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
            if (!nl) break;
            View f[4]; int nf = 0;
            const char *q = p;
            while (nf < 4) {
                const char *tab = static_cast<const char *>(memchr(q, '\t', (size_t)(nl - q)));
                const char *fe = tab ? tab : nl;
                f[nf++] = View{q, (size_t)(fe - q)};
                if (!tab) break;
                q = tab + 1;
            }
            if (nf < 4) return fail(SD_ERR_ARG, "raw decomposition: a line has fewer than 4 columns");
            Line ln;
            if (!first_word(f[0], ln.read) || !first_word(f[1], ln.mono)) return fail(SD_ERR_ARG, "raw decomposition: empty read or monomer name");
            if (!py_int(f[2], ln.start) || !py_int(f[3], ln.end)) return fail(SD_ERR_ARG, "raw decomposition: start/end is not an integer");
            auto it = read_idx.find(ln.read.str());
            if (it == read_idx.end()) return fail(SD_ERR_ARG, "KeyError: read '" + ln.read.str() + "' is not in the sequences file");
            ln.read_idx = it->second;
            if (light) {
                auto mt = last_of.find(ln.mono.str());
                if (mt == last_of.end()) return fail(SD_ERR_ARG, "KeyError: monomer '" + ln.mono.str() + "' is not in the monomers file");
                ln.mono_idx = mt->second;
            } else {
                auto mt = upos.find(ln.mono.str());
                if (mt == upos.end()) return fail(SD_ERR_ARG, "KeyError: monomer '" + ln.mono.str() + "' is not in the monomers file");
                ln.mono_idx = mt->second;
            }
            lines.push_back(ln);
            p = nl + 1;
        }
    }
    if (stats) stats->lines_in = (int64_t)lines.size();

    Writer out(out_fd), alt(alt_fd);
    const int64_t per_line = light ? 1 : 2 * (int64_t)std::max(1, nm);
    size_t chunk = (size_t)std::max<int64_t>(1, std::min<int64_t>(kMaxPairsPerCall / per_line, kMaxLinesPerCall));
    if (const char *e = getenv("SD_CONVERT_LINES")) if (atoll(e) > 0) chunk = (size_t)atoll(e);        // tests: several device calls
    std::string qblob, qcblob; std::vector<int64_t> qoff, qcoff;
    std::vector<int32_t> pq, pt, mt, col, mt2, col2;
    for (size_t lo = 0; lo < lines.size(); lo += chunk) {
        const size_t n = std::min(chunk, lines.size() - lo);
        // read.seq[start:end + 1] (main.py:115,125), minus one trailing '*'
        qblob.clear(); qoff.assign(1, 0);
        for (size_t k = 0; k < n; ++k) {
            const Line &ln = lines[lo + k];
            const int64_t base = read_off[ln.read_idx], len = read_off[ln.read_idx + 1] - base;
            long long a, b;
            py_slice(len, ln.start, ln.end + 1, a, b);
            if (b > a && reads[base + b - 1] == '*') --b;
            qblob.append(reads + base + a, (size_t)(b - a));
            qoff.push_back((int64_t)qblob.size());
        }
        int64_t hb = 0; double ms = 0;
        auto identity = [&](const std::string &qb, const std::vector<int64_t> &qo, const std::string &tb, const std::vector<int64_t> &to,
                            bool paired, std::vector<int32_t> &m, std::vector<int32_t> &c) {
            const int64_t np = paired ? (int64_t)n : (int64_t)n * nm;
            m.assign((size_t)np, 0); c.assign((size_t)np, 0);
            int64_t h = 0; double t = 0;
            const int st = sd_identity(qb.data(), qo.data(), (int64_t)n, tb.data(), to.data(), nm, paired ? pq.data() : nullptr,
                                       paired ? pt.data() : nullptr, np, m.data(), c.data(), nullptr, device, &h, &t);
            hb += h; ms += t;
            if (stats) stats->pairs += np;
            return st;
        };
        auto pct = [](int32_t m, int32_t c) { return c > 0 ? (double)m / (double)c * 100 : 0.0; };       // main.py:58-60
        if (light) {
            pq.resize(n); pt.resize(n);
            for (size_t k = 0; k < n; ++k) { pq[k] = (int32_t)k; pt[k] = lines[lo + k].mono_idx; }
            if (int st = identity(qblob, qoff, tblob, toff, true, mt, col)) return fail(st, sd_last_error(nullptr));
            for (size_t k = 0; k < n; ++k) {
                const Line &ln = lines[lo + k];
                const double v = pct(mt[k], col[k]);
                if (!(v >= min_identity)) continue;
                out.add(ln.read); out.buf.push_back('\t'); out.add(ln.mono); out.buf.push_back('\t');
                out.add_int(ln.start); out.buf.push_back('\t'); out.add_int(ln.end); out.buf.push_back('\t'); out.add_f2(v);
                out.buf += "\tNone\t-1.00\tNone\t-1.00\tNone\t-1.00\t";
                out.buf.push_back(classify(v, -1.0)); out.buf.push_back('\n');
                if (stats) ++stats->lines_out;
                if (!out.maybe_flush()) return fail(SD_ERR_INTERNAL, "sd_convert: write failed");
            }
        } else {
            collapse(qblob, qoff, qcblob, qcoff);
            if (int st = identity(qblob, qoff, tblob, toff, false, mt, col)) return fail(st, sd_last_error(nullptr));
            if (int st = identity(qcblob, qcoff, tcblob, tcoff, false, mt2, col2)) return fail(st, sd_last_error(nullptr));
            std::vector<double> row((size_t)nu);
            for (size_t k = 0; k < n; ++k) {
                const Line &ln = lines[lo + k];
                const size_t base = k * (size_t)nm;
                for (int u = 0; u < nu; ++u) row[(size_t)u] = pct(mt[base + (size_t)ucol[(size_t)u]], col[base + (size_t)ucol[(size_t)u]]);
                const double best = row[(size_t)ln.mono_idx];
                int second = -1; double second_score = -1;                       // first of the largest, main.py:131-135
                for (int u = 0; u < nu; ++u)
                    if (u != ln.mono_idx && (second < 0 || second_score < row[(size_t)u])) { second = u; second_score = row[(size_t)u]; }
                int h0 = -1, h1 = -1; double s0 = 0, s1 = 0;                      // two best of the stable sort, main.py:143
                for (int m = 0; m < nm; ++m) {
                    const double s = pct(mt2[base + (size_t)m], col2[base + (size_t)m]);
                    if (h0 < 0 || s > s0) { h1 = h0; s1 = s0; h0 = m; s0 = s; }
                    else if (h1 < 0 || s > s1) { h1 = m; s1 = s; }
                }
                if (!(best >= min_identity)) continue;
                const char q = classify(best, second < 0 ? -1.0 : second_score);
                out.add(ln.read); out.buf.push_back('\t'); out.add(ln.mono); out.buf.push_back('\t');
                out.add_int(ln.start); out.buf.push_back('\t'); out.add_int(ln.end); out.buf.push_back('\t'); out.add_f2(best);
                out.buf.push_back('\t');
                if (second < 0) out.buf += "None\t-1.00"; else { out.add(*uniq[(size_t)second]); out.buf.push_back('\t'); out.add_f2(second_score); }
                out.buf.push_back('\t'); out.add(mname[(size_t)h0]); out.buf.push_back('\t'); out.add_f2(s0);
                out.buf.push_back('\t'); out.add(mname[(size_t)h1]); out.buf.push_back('\t'); out.add_f2(s1);
                out.buf.push_back('\t'); out.buf.push_back(q); out.buf.push_back('\n');
                for (int u = 0; u < nu; ++u) {
                    alt.add(ln.read); alt.buf.push_back('\t'); alt.add(*uniq[(size_t)u]); alt.buf.push_back('\t');
                    alt.add_int(ln.start); alt.buf.push_back('\t'); alt.add_int(ln.end); alt.buf.push_back('\t'); alt.add_f2(row[(size_t)u]);
                    alt.buf.push_back('\t'); alt.buf.push_back(u == ln.mono_idx ? '*' : '-'); alt.buf.push_back('\n');
                }
                if (stats) ++stats->lines_out;
                if (!out.maybe_flush() || !alt.maybe_flush()) return fail(SD_ERR_INTERNAL, "sd_convert: write failed");
            }
        }
        if (stats) { stats->hirschberg_pairs += hb; stats->kernel_ms += ms; }
    }
    if (!out.flush() || !alt.flush()) return fail(SD_ERR_INTERNAL, "sd_convert: write failed");
    return SD_OK;
}
