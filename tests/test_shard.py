"""Host logic of the region sharding, incl. a world-size-2 gloo run on CPU.
The per-rank 'realignment' in the gloo test is the CPU oracle standing in for a GPU: what is
under test is the partition, the insert-range reduction and the order-preserving merge."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from indelminer_b200 import shard, synth


def test_region_bounds_and_owner():
    lens = [1000, 500, 2500]
    b = shard.region_bounds(lens, 4)
    assert list(b) == [0, 1000, 2000, 3000, 4000]
    own = shard.owner_of([0, 0, 1, 2, 2, 2], [0, 999, 0, 499, 500, 2499], lens, 4)
    assert list(own) == [0, 0, 1, 1, 2, 3]


def test_take_shard_and_merge_roundtrip():
    ref = synth.make_reference(200_000, seed=3)
    w = synth.make_candidates(ref, 500, seed=5)
    own = shard.owner_of(w["tid"], w["position"], [len(ref)], 3)
    parts = []
    seen = 0
    for r in range(3):
        sub, idx = shard.take_shard(w, r, own)
        seen += len(idx)
        assert np.all(np.diff(idx) > 0)                       # original order kept inside a shard
        M = w["read_len"]
        for j, i in enumerate(idx):
            assert np.array_equal(sub["read_bases"][sub["read_off"][j]:sub["read_off"][j + 1]],
                                  w["read_bases"][i * M:(i + 1) * M])
        words = [[(int(i) << 4) | 7] * (1 + int(i) % 3) for i in idx]
        parts.append((idx, np.full(len(idx), r, np.int32), np.array([len(x) for x in words], np.int32),
                      sub["position"], words))
    assert seen == 500
    status, nseg, rstart, seg_off, segs = shard.merge_results(500, parts)
    assert np.array_equal(status, own)
    assert np.array_equal(rstart, w["position"])
    for i in range(500):
        assert list(segs[seg_off[i]:seg_off[i] + nseg[i]]) == [(i << 4) | 7] * (1 + i % 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    ref = synth.make_reference(120_000, seed=11)
    w = synth.make_candidates(ref, 240, seed=13)
    # every rank sees a different local estimate of the insert range; range[1] must end up global
    local = np.array([[100 + rank, 600 + 10 * rank]], dtype=np.int64)
    rng = shard.reduce_insert_ranges(local, dist)
    assert rng.tolist() == [[100, 600 + 10 * (world - 1)]]
    own = shard.owner_of(w["tid"], w["position"], [len(ref)], world)
    sub, idx = shard.take_shard(w, rank, own)
    p = O.default_params()
    cs = ref.tobytes()
    st, ns, rs, words = [], [], [], []
    for j in range(len(idx)):
        read = sub["read_bases"][sub["read_off"][j]:sub["read_off"][j + 1]].tobytes()
        o = O.realign_read(p, cs, int(sub["position"][j]), int(sub["range1"][j]), read)
        segs = o.segments()
        st.append(o.status); ns.append(len(segs)); rs.append(segs[0][2] if segs else 0)
        words.append([(ln << 4) | op for op, ln, _s, _e in segs])
    np.save(os.path.join(outdir, f"idx{rank}.npy"), idx)
    np.save(os.path.join(outdir, f"st{rank}.npy"), np.array(st, np.int32))
    np.save(os.path.join(outdir, f"ns{rank}.npy"), np.array(ns, np.int32))
    np.save(os.path.join(outdir, f"rs{rank}.npy"), np.array(rs, np.int32))
    np.save(os.path.join(outdir, f"w{rank}.npy"), np.array(words, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = []
    for r in range(world):
        ld = lambda n: np.load(os.path.join(tmp_path, f"{n}{r}.npy"), allow_pickle=True)
        parts.append((ld("idx"), ld("st"), ld("ns"), ld("rs"), list(ld("w"))))
    status, nseg, rstart, seg_off, segs = shard.merge_results(240, parts)
    from oracle import oracle as O
    ref = synth.make_reference(120_000, seed=11)
    w = synth.make_candidates(ref, 240, seed=13)
    p = O.default_params()
    cs = ref.tobytes()
    M = w["read_len"]
    for i in range(240):
        o = O.realign_read(p, cs, int(w["position"][i]), int(w["range1"][i]), w["read_bases"][i * M:(i + 1) * M].tobytes())
        segs_i = o.segments()
        assert status[i] == o.status
        assert list(segs[seg_off[i]:seg_off[i] + nseg[i]]) == [(ln << 4) | op for op, ln, _s, _e in segs_i]
        if segs_i:
            assert rstart[i] == segs_i[0][2]


def test_pack4_round_trip():
    """api.pack4 (ASCII -> the BAM's 4-bit form, optionally stored reverse-complemented and flagged) decoded the way
    unpack_reads4_kernel does (high nibble first, bit2char's table, complement + reversal when flagged) gives the
    reads back: equal-length fast path and ragged path"""
    import numpy as np
    from indelminer_b200 import api
    rng = np.random.default_rng(3)
    dec = {1: "A", 2: "C", 4: "G", 8: "T", 15: "N"}
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    for lens in ([151] * 40, list(rng.integers(1, 90, size=60))):
        reads = ["".join(rng.choice(list("ACGTN"), p=[.24, .24, .24, .24, .04]) for _ in range(L)) for L in lens]
        data, off = api.pack_sequences(reads)
        flags = rng.random(len(reads)) < 0.5
        seq4, boff, ln, fl = api.pack4(data, off, flags)
        assert list(ln) == list(lens) and boff[-1] == sum((L + 1) // 2 for L in lens)
        for i, r in enumerate(reads):
            b = seq4[boff[i]:boff[i + 1]]
            s = "".join(dec[(b[t // 2] >> 4) if t % 2 == 0 else (b[t // 2] & 15)] for t in range(lens[i]))
            if fl[i]:
                s = "".join(comp[ch] for ch in reversed(s))
            assert s == r, i
