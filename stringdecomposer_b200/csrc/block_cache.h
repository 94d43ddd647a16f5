// block_cache.h -- process-wide cache of released device and page-locked blocks (policy only: no CUDA calls in here, the
// caller frees what flush() hands back; unit-tested on the CPU by tests/emu/block_cache_test.cpp).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace sdb {

// Process-wide cache of released device and page-locked blocks.  Creating and destroying an engine costs some thirty
// cudaMalloc / cudaFree / cudaHostAlloc calls; each is a round trip into the kernel driver (and cudaFree a device-wide
// synchronisation) that takes from a fraction of a millisecond to tens of milliseconds depending on the box -- far more
// than the kernels of a small batch.  A process that opens many handles (one per sd_run_files call, one per monomer
// set) gets its blocks back from here instead.  Bounded (4 GiB and 512 blocks per device, no block above 1 GiB); emptied
// and retried when an allocation fails; SD_NO_BUFFER_CACHE=1 turns it off.  Never destroyed: the driver reclaims the
// memory at process exit, and no CUDA call has to run from a static destructor.
class BlockCache {
public:
    enum Kind { Device = 0, Pinned = 1 };
    void *take(Kind k, int dev, size_t bytes, size_t *cap)
    {
        if (off()) return nullptr;
        std::lock_guard<std::mutex> l(mu_);
        int best = -1;
        for (size_t i = 0; i < blocks_.size(); ++i) {
            const Block &b = blocks_[i];
            if (b.kind != k || b.dev != dev || b.cap < bytes || b.cap > 4 * bytes + ((size_t)1 << 20)) continue;
            if (best < 0 || b.cap < blocks_[(size_t)best].cap) best = (int)i;
        }
        if (best < 0) return nullptr;
        Block b = blocks_[(size_t)best];
        blocks_.erase(blocks_.begin() + best);
        held_[k] -= b.cap;
        *cap = b.cap;
        return b.p;
    }
    bool give(Kind k, int dev, void *p, size_t cap)
    {
        if (off() || dev < 0 || cap > ((size_t)1 << 30)) return false;
        std::lock_guard<std::mutex> l(mu_);
        size_t n = 0, bytes = 0;
        for (const Block &b : blocks_) if (b.kind == k && b.dev == dev) { ++n; bytes += b.cap; }
        if (n >= 512 || bytes + cap > (k == Device ? (size_t)4 << 30 : (size_t)512 << 20)) return false;
        blocks_.push_back(Block{p, cap, dev, k});
        held_[k] += cap;
        return true;
    }
    std::vector<void *> flush(Kind k, int dev)      // an allocation failed: everything of that kind goes back to the driver (by the caller)
    {
        std::vector<void *> drop;
        std::lock_guard<std::mutex> l(mu_);
        for (size_t i = 0; i < blocks_.size();)
            if (blocks_[i].kind == k && blocks_[i].dev == dev) { drop.push_back(blocks_[i].p); held_[k] -= blocks_[i].cap; blocks_.erase(blocks_.begin() + (long)i); }
            else ++i;
        return drop;
    }
    size_t held(Kind k) { std::lock_guard<std::mutex> l(mu_); return held_[k]; }
    static BlockCache &get() { static BlockCache *c = new BlockCache; return *c; }

private:
    struct Block { void *p; size_t cap; int dev; Kind kind; };
    static bool off() { const char *e = getenv("SD_NO_BUFFER_CACHE"); return e && *e && *e != '0'; }
    std::mutex mu_;
    std::vector<Block> blocks_;
    size_t held_[2] = {0, 0};
};

} // namespace sdb
