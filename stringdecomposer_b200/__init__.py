"""stringdecomposer_b200 -- B200-native string-decomposition DP behind the reference's `dp` boundary.

The product is ``libsd_b200.so`` (hand-written sm_100a CUDA + C++ host driver, C ABI in ``include/sd_b200.h``)
and the drop-in ``build/bin/dp`` binary.  This package is the thin Python host mirror used by the tests and
``bench.py``; it has no CPU fallback -- creating a :class:`Decomposer` without the CUDA library or without a
GPU raises.
"""
from ._lib import (Decomposer, SdError, load_library, segment_read, postprocess, run_files, int_peak,  # noqa: F401
                   library_path, device_count, hw_distance, nw_identity, RECORD_DTYPE)
from .hostpipe import decompose_reads, read_fasta, format_raw_tsv  # noqa: F401

__version__ = "0.1.0"
