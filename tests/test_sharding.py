"""Multi-GPU host logic on the CPU: the contig assignment, the position split, and the gather of
per-range results with torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from deepgrp_b200 import sharding


def test_assign_records_largest_first_balanced():
    lengths = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80,
               58, 64, 46, 50, 156, 57]          # human chromosomes, Mbp
    for world in (1, 2, 4, 8):
        owner = sharding.assign_records(lengths, world)
        assert len(owner) == len(lengths) and set(owner) <= set(range(world))
        load = [sum(l for l, o in zip(lengths, owner) if o == r) for r in range(world)]
        assert max(load) - min(load) <= max(lengths)
        assert max(load) <= 1.15 * sum(lengths) / world + (0 if world < 8 else 20)
    assert sharding.assign_records([5, 5, 5], 2) == [0, 1, 0]       # ties: stable order, lowest rank
    assert sharding.assign_records([], 4) == []


def test_split_positions_covers_exactly():
    for length in (0, 1, 7, 1000, 46_700_000):
        for parts in (1, 2, 3, 8):
            r = sharding.split_positions(length, parts)
            assert r[0][0] == 0 and r[-1][1] == length and len(r) == parts
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            assert all(a <= b for a, b in r)


def test_merge_record_texts_orders_by_record():
    pieces = [[(2, b"c\n"), (0, b"a\n")], [(1, b"b\n")]]
    assert sharding.merge_record_texts(pieces) == b"a\nb\nc\n"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, length, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # stand-in for the GPU range computation: a deterministic function of the position
        def compute_range(p0, p1):
            pos = np.arange(p0, p1, dtype=np.int64)
            return (pos % 5).astype(np.uint8), (np.sin(pos * 0.001) * 10).astype(np.float32)

        def finish(labels, scores):
            np.save(os.path.join(out_dir, "labels.npy"), labels)
            np.save(os.path.join(out_dir, "scores.npy"), scores)
            return True
        res = sharding.predict_record_sharded(compute_range, finish, length, rank, world, owner=0, dist=dist)
        assert (res is True) == (rank == 0)
    finally:
        dist.destroy_process_group()


def test_gather_ranges_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    length, world = 100_003, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, length, str(tmp_path)), nprocs=world, join=True)
    pos = np.arange(length, dtype=np.int64)
    assert np.array_equal(np.load(tmp_path / "labels.npy"), (pos % 5).astype(np.uint8))
    assert np.array_equal(np.load(tmp_path / "scores.npy"), (np.sin(pos * 0.001) * 10).astype(np.float32))


def test_fasta_slices_partition_the_text_at_headers():
    """dgrp_fasta_index (the streaming driver's host index, no GPU): slices are whole records of >= 8 MiB, they
    partition the text, and the largest-first assignment is balanced and the same for every rank."""
    from deepgrp_b200 import sharding
    rng = np.random.default_rng(4)
    sizes = [9_500_000, 300, 12_000_000, 40_000, 8_400_000, 20_000_000, 1_000]
    parts = []
    for k, n in enumerate(sizes):
        body = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=n)].tobytes()
        lines = b"\n".join(body[i:i + 60] for i in range(0, n, 60))
        parts.append(b">rec%d with > inside\n" % k + lines + b"\n")
    raw = b"".join(parts)
    cuts, owner = sharding.fasta_slices(raw, 3)
    assert cuts[0] == 0 and cuts[-1] == len(raw) and (np.diff(cuts) > 0).all()
    for c in cuts[1:-1]:
        assert raw[c:c + 4] == b">rec" and raw[c - 1:c] == b"\n"           # cut points are header lines
    assert (np.diff(cuts)[:-1] >= 8 << 20).all()                             # small records ride with a neighbour
    load = np.bincount(owner, weights=np.diff(cuts), minlength=3)
    assert load.max() <= 1.6 * load.mean()
    cuts1, owner1 = sharding.fasta_slices(raw, 1)
    assert np.array_equal(cuts1, cuts) and (owner1 == 0).all()
    assert sharding.fasta_slices(b"", 2)[0].tolist() == [0, 0]


def test_stream_slab_plan():
    """dgrp_fasta_stream_plan (host arithmetic of the early-rows route, no GPU): the slabs partition the record,
    every slab but the last is a whole number of forward waves minus the halo windows a position range recomputes
    (so the persistent kernel's CTAs get equal tile counts), they shrink, and short records get no plan."""
    from deepgrp_b200 import sharding
    T, step, unit = 342, 50, 148 * 128
    ends = sharding.stream_slabs(46_700_000, T, step)
    assert ends[-1] == 46_700_000 and len(ends) == 6 and (np.diff(ends) > 0).all()
    sizes = np.diff(np.concatenate([[0], ends]))
    assert (sizes[:-2] > sizes[1:-1]).all()                       # geometric: only the last slab's text is exposed
    assert sizes[-1] < 0.15 * 46_700_000 and sizes[0] < 0.30 * 46_700_000   # the copies start early, the tail is short
    halo = -(-T // step) + 8
    for n in sizes[:-1]:
        assert n % 64 == 0
        windows = n // step + halo                                # what the range recomputes, roughly
        assert windows % unit < 0.01 * unit or unit - windows % unit < 0.01 * unit
    assert len(sharding.stream_slabs(3_000_000, T, step)) == 0    # below ~4 waves: one call
    assert len(sharding.stream_slabs(5_000_000, T, step)) == 2
    for length in (4_736_000, 10_000_001, 248_000_000, 2_000_000_000):
        for slabs, ratio in ((0, 0), (2, 30), (6, 90), (3, 100)):
            e = sharding.stream_slabs(length, T, step, 0, slabs, ratio)
            assert len(e) == 0 or (e[-1] == length and (np.diff(e) > 0).all() and e[0] > 0 and len(e) <= max(slabs, 6))
    small = sharding.stream_slabs(60_000, 150, 50, 64, 6)          # the unit the GPU tests use
    assert 4 <= len(small) <= 6 and small[-1] == 60_000 and (np.diff(small) > 0).all()
