"""Multi-GPU orchestration of the hot path: contigs are independent units (ntedit.cpp:2220-2245 -- the
reference hands one contig to each OpenMP thread), so they shard across the GPUs of a box with the filter replicated
on each GPU.  The only collective is the broadcast of the filter bytes at load time; results come back per contig and
are merged in input order (the reference's `-t 1` order).  Pure host logic -- no hashing, no filter probes.
"""
import heapq


def assign_contigs(lengths, world):
    """Longest-processing-time greedy: contig i -> rank owner[i], weight = contig length.  Deterministic."""
    owner = [0] * len(lengths)
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    for i in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + lengths[i], r))
    return owner


def my_contigs(owner, rank):
    return [i for i, r in enumerate(owner) if r == rank]


def broadcast_filter(dist, tensor, src=0):
    """The single collective of the design: replicate the filter bytes (a uint8 tensor, already allocated with the
    same size on every rank) from `src`.  NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests."""
    dist.broadcast(tensor, src=src)
    return tensor


def merge_in_input_order(n_contigs, gathered):
    """gathered: per rank, a dict {contig index: (fa, tsv_rows, vcf_rows)} -> three byte strings in input order
    (contigs below the -z cut-off are absent from every dict, as they are from the reference's outputs)."""
    fa, tsv, vcf = [], [], []
    for i in range(n_contigs):
        for part in gathered:
            if i in part:
                a, b, c = part[i]
                fa.append(a)
                tsv.append(b)
                vcf.append(c)
                break
    return b"".join(fa), b"".join(tsv), b"".join(vcf)


def polish_sharded(contigs, polish_contigs, dist, rank, world):
    """Polish `contigs` ([(header, seq)]) across `world` ranks.  `polish_contigs(list of (header, seq))` returns one
    (fa, tsv_rows, vcf_rows) per input contig (None for contigs dropped by -z).  Every rank gets the merged outputs."""
    owner = assign_contigs([len(s) for _, s in contigs], world)
    mine = my_contigs(owner, rank)
    outs = polish_contigs([contigs[i] for i in mine])
    part = {i: o for i, o in zip(mine, outs) if o is not None}
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    return merge_in_input_order(len(contigs), gathered)
