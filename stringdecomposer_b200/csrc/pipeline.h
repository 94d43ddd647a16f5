// pipeline.h -- host driver: FASTA ingest, segmentation, device scheduling, overlap resolution, raw TSV.
#pragma once
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

namespace sdb {

struct FastaSet { std::vector<std::string> names, seqs; };

// load_fasta (main.cpp:314-346).  Returns 0, or 255 after writing the reference's error line to diag.
int load_fasta(const std::string &path, FastaSet &out, std::string &diag);

struct SegRef { int read; int off; int len; };
// AlignReadsSet segmentation (main.cpp:70-81).  Returns -1 for part_size <= 0 (the reference never terminates).
int64_t segment_read(int64_t read_len, int part_size, int overlap, std::vector<std::pair<int, int>> *out);

// PostProcessing (main.cpp:287-302)
void postprocess(const std::vector<Record> &in, std::vector<Record> &out);

struct EngineStats {
    double sweep_ms = 0, traceback_ms = 0, h2d_ms = 0, d2h_ms = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0, cells = 0, segments = 0, columns = 0, launches = 0;
    Geometry g{};
    int64_t sweep_store_bytes = 0;      // codes + J/argJ bytes of the last batch (all devices), from the layout
    int dev_segments[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// Owns the monomer set, the scoring and one backend per device; decomposes batches of segments.
class DevicePool;

class Engine {
public:
    Engine(const std::vector<std::string> &forward_monomers, const Scoring &sc, std::vector<std::unique_ptr<Backend>> devs);
    ~Engine();
    void decompose(const Batch &b, BatchResult &out);          // throws PlanError
    void stage(const Batch &b);                                 // whole batch resident, one wave per device
    double run_staged();                                        // returns kernel ms (max over devices)
    void fetch_staged(BatchResult &out);
    void set_ed_thr(int ed_thr);                                // --ed_thr monomer pre-filter, -1 = off (main.cpp:135-149)
    void set_plan_hint(int max_seg_len, int64_t nseg_total);    // shape of a job that arrives in chunks (run_files)
    const MonomerSet &monomers() const { return ms_; }
    EngineStats stats;
    int ndev() const { return (int)devs_.size(); }

private:
    void plan_for(const Batch &b);
    void split(const Batch &b, std::vector<int> &bounds) const;
    void note_split(const Batch &b, const std::vector<int> &bounds);
    void on_devices(const std::function<void(int)> &fn);      // fn(d) on every device's own host thread
    std::unique_ptr<DevicePool> pool_;
    MonomerSet ms_;
    Scoring sc_;
    std::vector<std::unique_ptr<Backend>> devs_;
    Plan plan_; bool have_plan_ = false; int plan_maxlen_ = -1; int64_t plan_nseg_ = -1;
    int hint_maxlen_ = 0; int64_t hint_nseg_ = 0;
    Batch staged_; std::vector<int> staged_bounds_;
};

// The whole `dp` run (main.cpp:374-402 after argv parsing).  Returns the exit status.  `open_devices` creates the
// backends (CUDA context creation, ~0.4 s); it runs on its own thread while the FASTA files are read and segmented.
using DeviceOpener = std::function<int(std::vector<std::unique_ptr<Backend>> &, std::string &)>;
int run_files(const std::string &reads_path, const std::string &monomers_path, int threads, int part_size, int overlap,
              const Scoring &sc, int ed_thr, DeviceOpener open_devices, int out_fd, int err_fd, std::string &error);

} // namespace sdb
