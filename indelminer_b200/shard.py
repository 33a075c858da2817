"""Region sharding of the realignment batch across GPUs and the order-preserving
merge of the per-rank results.

The path shards by independent units: a candidate read's result depends only on the read, its mate
position, its read group's range and the (replicated) reference.  Each rank owns contiguous genomic
regions, realigns the candidates anchored there on its own GPU and returns them to the host; the
host re-serialises by the global read index so that the unchanged consumer (fetch_func's evidence
lists, the READCHUNK flush counter, indelminer.c:617) sees the original BAM order.  There is no
collective on the data path; the only reduction is the per-read-group (min, max) insert range
(bamoperations.c:15-86), a few bytes.
"""
import numpy as np


def region_bounds(contig_lengths, world):
    """Split the concatenated genome into `world` contiguous regions of (almost) equal size.
    Returns int64[world + 1] boundaries in concatenated coordinates."""
    total = int(np.sum(contig_lengths, dtype=np.int64))
    return (np.arange(world + 1, dtype=np.int64) * total) // world


def owner_of(tid, position, contig_lengths, world):
    """Rank that owns each candidate: the region holding its anchor (mate position)."""
    starts = np.concatenate([[0], np.cumsum(contig_lengths, dtype=np.int64)[:-1]])
    g = starts[np.asarray(tid, dtype=np.int64)] + np.asarray(position, dtype=np.int64)
    b = region_bounds(contig_lengths, world)
    return (np.searchsorted(b, g, side="right") - 1).clip(0, world - 1).astype(np.int32)


def take_shard(batch, rank, owner):
    """The sub-batch of `rank` (dict with read_bases/read_off/tid/position/range1) plus the global
    indices of its reads, in their original order."""
    idx = np.nonzero(owner == rank)[0]
    off = batch["read_off"]
    lens = (off[1:] - off[:-1])[idx]
    new_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    bases = np.empty(int(new_off[-1]), dtype=np.uint8)
    for j, i in enumerate(idx):
        bases[new_off[j]:new_off[j + 1]] = batch["read_bases"][off[i]:off[i + 1]]
    return dict(read_bases=bases, read_off=new_off, tid=np.ascontiguousarray(batch["tid"][idx]),
                position=np.ascontiguousarray(batch["position"][idx]),
                range1=np.ascontiguousarray(batch["range1"][idx])), idx


def merge_results(n, parts):
    """parts: list of (global_idx, status, nseg, rstart, words_per_read list).  Returns the arrays in
    the original read order plus a compact segment array (seg_off authoritative)."""
    status = np.zeros(n, dtype=np.int32)
    nseg = np.zeros(n, dtype=np.int32)
    rstart = np.zeros(n, dtype=np.int32)
    words = [None] * n
    for idx, st, ns, rs, w in parts:
        status[idx] = st
        nseg[idx] = ns
        rstart[idx] = rs
        for j, i in enumerate(idx):
            words[i] = w[j]
    seg_off = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(nseg[:-1], out=seg_off[1:])
    segs = np.concatenate([np.asarray(w, dtype=np.uint32) for w in words if w is not None and len(w)] or
                          [np.zeros(0, dtype=np.uint32)])
    return status, nseg, rstart, seg_off, segs


def reduce_insert_ranges(local_ranges, dist=None):
    """Per-read-group (min, max) proper-insert range across ranks: element-wise min / max
    (bamoperations.c:62-86 computes them per BAM; sharded by region they are reduced here).
    local_ranges: int64[ngroups, 2].  With `dist` (torch.distributed, any backend) the reduction
    is two tiny all-reduces; without it the input is returned."""
    r = np.asarray(local_ranges, dtype=np.int64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return r
    import torch
    lo = torch.from_numpy(np.ascontiguousarray(r[:, 0]))
    hi = torch.from_numpy(np.ascontiguousarray(r[:, 1]))
    if dist.get_backend() == "nccl":
        lo, hi = lo.cuda(), hi.cuda()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return np.stack([lo.cpu().numpy(), hi.cpu().numpy()], axis=1)
