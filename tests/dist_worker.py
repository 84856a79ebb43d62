"""Worker of tests/test_shard_gloo.py: one of WORLD_SIZE processes (gloo, CPU).  Rank 0 builds the filter and
broadcasts its bytes (the design's single collective); every rank polishes its LPT shard of the contigs -- with the CPU
simulator of the device engine standing in for the GPU -- and the merged outputs must equal the single-process run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ntedit_b200 import shard, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.hostsim import pyhostsim as hs  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    k, h, fbytes = 25, 3, 1 << 16
    rng = np.random.default_rng(99)                       # same inputs on every rank
    truths = [synth.random_genome(n, rng) for n in (9000, 700, 4000, 60, 2500, 5200, 1300)]
    contigs = [(b"ctg%d x" % i, synth.mutate(t, rng, 2e-3, 5e-4, lower_frac=0.01, n_frac=0.004).tobytes())
               for i, t in enumerate(truths)]
    filt_t = torch.zeros(fbytes, dtype=torch.uint8)
    if rank == 0:
        f = po.OracleFilter.new(fbytes, k, h, False)
        for t in truths:
            f.insert_seq(t.tobytes())
        filt_t.copy_(torch.from_numpy(f.data().copy()))
        f.free()
    shard.broadcast_filter(dist, filt_t, src=0)
    filt = filt_t.numpy().tobytes()
    params = hs.default_params(mode=1)

    def polish_contigs(cs):
        out = []
        for c in cs:
            if len(c[1]) < params.min_contig_len:
                out.append(None)
                continue
            fa, tsv, vcf, _ = hs.polish([c], filt, k, h, False, params)
            out.append((fa, tsv.split(b"\n", 1)[1], vcf))   # drop the TSV header line
        return out

    fa, tsv, vcf = shard.polish_sharded(contigs, polish_contigs, dist, rank, world)
    owner = shard.assign_contigs([len(s) for _, s in contigs], world)
    if rank == 0:
        rfa, rtsv, rvcf, _ = hs.polish(contigs, filt, k, h, False, params)
        ok = fa == rfa and tsv == rtsv.split(b"\n", 1)[1] and vcf == rvcf
        loads = [sum(len(contigs[i][1]) for i in shard.my_contigs(owner, r)) for r in range(world)]
        print("SHARD_RESULT ok=%d loads=%s edits=%d" % (ok, loads, tsv.count(b"\n")), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
