"""ntedit_b200 -- B200-native implementation of ntEdit's hot path (ntHash k-mer walk + Bloom-filter driven edit decision).

The product is the CUDA library ntedit_b200/_lib/libntedit_b200.so (sources in ntedit_b200/csrc, C ABI in
include/ntedit_b200.h); this package is the thin host-side mirror of the reference's interface on top of it.
"""
from . import lib, shard  # noqa: F401
from .api import (Batch, BloomFilter, PolishResult, default_params, kmerize_and_correct,  # noqa: F401
                  kmerize_and_correct_device, pack_contigs, polish, polish_per_contig, scan, write_edits)

__all__ = ["lib", "Batch", "BloomFilter", "PolishResult", "default_params", "kmerize_and_correct",
           "kmerize_and_correct_device", "pack_contigs", "polish", "polish_per_contig", "scan", "write_edits", "shard"]
