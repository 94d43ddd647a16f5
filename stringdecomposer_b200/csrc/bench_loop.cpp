// bench_loop.cpp -- timing harness for bench.py's end-to-end leg: the C-ABI call sd_decompose() with host buffers, issued
// back to back from native code the way a C/C++ host application (such as the reference's own `dp` driver) would call
// it, so that the number does not include the ctypes/numpy cost of a Python caller.  Built as libsd_bench.so next to
// libsd_b200.so; it contains no kernels and no algorithm -- it only calls the public entry points of include/sd_b200.h.
#include <chrono>
#include <cstdint>

#include "../../include/sd_b200.h"

extern "C" int sd_bench_decompose(sd_handle *h, const char *segments, const int64_t *offsets, int64_t n_segments, int32_t warmup,
                                  int32_t steps, double *seconds, int64_t *n_records)
{
    sd_record *recs = nullptr;
    int64_t *roff = nullptr;
    for (int i = 0; i < warmup; ++i) {
        int st = sd_decompose(h, segments, offsets, n_segments, &recs, &roff);
        if (st) return st;
        sd_free(recs); sd_free(roff);
    }
    const auto t0 = std::chrono::steady_clock::now();
    int64_t total = 0;
    for (int i = 0; i < steps; ++i) {
        int st = sd_decompose(h, segments, offsets, n_segments, &recs, &roff);
        if (st) return st;
        total = roff[n_segments];
        sd_free(recs); sd_free(roff);
    }
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_records) *n_records = total;
    return 0;
}
