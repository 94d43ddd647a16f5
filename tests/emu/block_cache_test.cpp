// CPU unit test of the block-cache policy (stringdecomposer_b200/csrc/block_cache.h): pointers are opaque to the cache,
// so plain integers stand in for device memory.  Exit status 0 = all checks passed.
#include <cstdint>
#include <cstdio>

#include "block_cache.h"

using sdb::BlockCache;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "block_cache_test: line %d: %s\n", __LINE__, #c); ++failures; } } while (0)
static void *ptr(uintptr_t v) { return reinterpret_cast<void *>(v); }

int main()
{
    BlockCache c;
    size_t cap = 0;
    const size_t MB = (size_t)1 << 20, GB = (size_t)1 << 30;
    CHECK(c.take(BlockCache::Device, 0, 100, &cap) == nullptr);                       // empty
    CHECK(c.give(BlockCache::Device, 0, ptr(1), 64 * MB));
    CHECK(c.give(BlockCache::Device, 0, ptr(2), 8 * MB));
    CHECK(c.give(BlockCache::Device, 0, ptr(3), 16 * MB));
    CHECK(c.give(BlockCache::Device, 1, ptr(4), 8 * MB));                             // another device
    CHECK(c.give(BlockCache::Pinned, 0, ptr(5), 8 * MB));                             // another kind
    CHECK(c.held(BlockCache::Device) == 96 * MB && c.held(BlockCache::Pinned) == 8 * MB);
    CHECK(c.take(BlockCache::Device, 0, 6 * MB, &cap) == ptr(2) && cap == 8 * MB);    // best fit, not first fit
    CHECK(c.take(BlockCache::Device, 0, 6 * MB, &cap) == ptr(3) && cap == 16 * MB);   // next best, within 4x + 1 MiB
    CHECK(c.take(BlockCache::Device, 0, 6 * MB, &cap) == nullptr);                    // 64 MiB is too wasteful for 6 MiB
    CHECK(c.take(BlockCache::Device, 0, 65 * MB, &cap) == nullptr);                   // too small
    CHECK(c.take(BlockCache::Device, 2, 1 * MB, &cap) == nullptr);                    // no block of that device
    CHECK(c.take(BlockCache::Pinned, 0, 8 * MB, &cap) == ptr(5));
    CHECK(c.take(BlockCache::Device, 1, 8 * MB, &cap) == ptr(4));
    CHECK(c.take(BlockCache::Device, 0, 32 * MB, &cap) == ptr(1) && cap == 64 * MB);
    CHECK(c.held(BlockCache::Device) == 0 && c.held(BlockCache::Pinned) == 0);
    // bounds: no block above 1 GiB, 4 GiB per device, 512 MiB of pinned memory per device, 512 blocks per device and kind
    CHECK(!c.give(BlockCache::Device, 0, ptr(10), GB + 1));
    CHECK(!c.give(BlockCache::Device, -1, ptr(10), MB));
    for (int i = 0; i < 4; ++i) CHECK(c.give(BlockCache::Device, 0, ptr(20 + (uintptr_t)i), GB));
    CHECK(!c.give(BlockCache::Device, 0, ptr(30), MB));                                // device 0 is full ...
    CHECK(c.give(BlockCache::Device, 1, ptr(31), MB));                                 // ... device 1 is not
    CHECK(c.give(BlockCache::Pinned, 0, ptr(32), 512 * MB) && !c.give(BlockCache::Pinned, 0, ptr(33), 1));
    std::vector<void *> dropped = c.flush(BlockCache::Device, 0);
    CHECK(dropped.size() == 4 && c.held(BlockCache::Device) == MB);
    CHECK(c.take(BlockCache::Device, 1, MB, &cap) == ptr(31));
    CHECK(c.flush(BlockCache::Pinned, 0).size() == 1 && c.flush(BlockCache::Pinned, 0).empty());
    int kept = 0;
    for (int i = 0; i < 600; ++i) kept += c.give(BlockCache::Device, 3, ptr(1000 + (uintptr_t)i), 256);
    CHECK(kept == 512);
    CHECK(c.flush(BlockCache::Device, 3).size() == 512);
    // the switch
    setenv("SD_NO_BUFFER_CACHE", "1", 1);
    CHECK(!c.give(BlockCache::Device, 0, ptr(40), MB) && c.take(BlockCache::Device, 0, 1, &cap) == nullptr);
    setenv("SD_NO_BUFFER_CACHE", "0", 1);
    CHECK(c.give(BlockCache::Device, 0, ptr(40), MB) && c.take(BlockCache::Device, 0, MB, &cap) == ptr(40));
    if (!failures) std::puts("block_cache_test: ok");
    return failures ? 1 : 0;
}
