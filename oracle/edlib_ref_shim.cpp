// edlib_ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.  A three-line C entry point over the reference's own vendored
// edlib (compiled from /root/reference/stringdecomposer/src/edlib.cpp where it lies; nothing is copied), so the
// identity oracle can be pinned against the exact calls main.py:29-60 makes:
//   edlib.align(query, target, mode="NW", task="path")  ==  edlibAlign(q, qlen, t, tlen, {k=-1, NW, PATH}).
#include "edlib.h"
extern "C" int ref_nw_path_counts(const char *q, int qlen, const char *t, int tlen, int *matches, int *columns)
{
    EdlibAlignConfig cfg = edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_PATH, 0, 0);
    EdlibAlignResult r = edlibAlign(q, qlen, t, tlen, cfg);
    int m = 0;
    for (int i = 0; i < r.alignmentLength; ++i) m += r.alignment[i] == EDLIB_EDOP_MATCH;
    *matches = m; *columns = r.alignmentLength;
    int d = r.status == EDLIB_STATUS_OK ? r.editDistance : -2;
    edlibFreeAlignResult(r);
    return d;
}
extern "C" int ref_hw_distance(const char *q, int qlen, const char *t, int tlen)
{
    EdlibAlignResult r = edlibAlign(q, qlen, t, tlen, edlibNewAlignConfig(-1, EDLIB_MODE_HW, EDLIB_TASK_DISTANCE, 0, 0));
    int d = r.editDistance; edlibFreeAlignResult(r); return d;
}
