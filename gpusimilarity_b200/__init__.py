"""gpusimilarity_b200 — B200-native brute-force Tanimoto scan + fused top-k behind
gpusimilarity's FingerprintDB API.  Compute lives in libgpusim_b200.so (hand-written sm_100a
CUDA, C ABI in include/gpusim_b200.h); this package is the host-side mirror of the reference
interface plus the .fsim reader.  Nothing here falls back to a CPU or oracle path."""
from .fingerprintdb import (FingerprintDB, GsbError, ScanInfo, fold_fingerprint, get_available_gpu_memory,
                            get_gpu_count, get_next_gpu, launch_count, merge_device,
                            top_results_bubble_sort)
from .fsim import FsimData, read_fsim, write_fsim

__all__ = ["FingerprintDB", "GsbError", "ScanInfo", "fold_fingerprint", "get_available_gpu_memory",
           "get_gpu_count", "get_next_gpu", "launch_count", "merge_device", "top_results_bubble_sort",
           "FsimData", "read_fsim", "write_fsim"]
