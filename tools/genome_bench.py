"""BASELINE.json configs[3] in shape: a synthetic multi-FASTA of 24 contigs with lengths proportional to
the human chromosomes (leading / trailing / internal N runs, ~50 % of the bases lower-cased in runs),
annotated end to end -- FASTA bytes in host memory -> GPU decode -> forward -> MSS -> TSV text in host
memory -- with the contigs sharded over the ranks (largest first, no collective on the data path).

    python tools/genome_bench.py --scale 0.25                         # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/genome_bench.py --scale 0.25

--scale 1.0 is the 3.1 Gbp genome.  With random-init weights almost every base lands in a TSV row
(~25 bytes of text per base); the text streams through 4 x 64 MiB of pinned host memory per rank
(dgrp_fasta_stream_*).  Rank 0 generates the file into /dev/shm, every rank maps it and uploads only its own
contigs.  `bench.py` runs the same measurement as its "genome" section (write_fasta is imported from here)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# GRCh38 chromosome lengths (Mbp): 1..22, X, Y
HUMAN_MBP = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28,
             114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]
NAMES = ["chr%d" % i for i in range(1, 23)] + ["chrX", "chrY"]


def contig_letters(n, k):
    """Record k: iid ACGT from default_rng([1, k]); 10 kb of N at either end, one
    internal 1-3 Mbp N run and several 50 kb N gaps (scaled down with the contig), soft-masked runs."""
    rng = np.random.default_rng([1, k])
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)].copy()
    edge = min(10_000, n // 100)
    letters[:edge] = ord("N")
    letters[n - edge:] = ord("N")
    big = int(min(rng.integers(1_000_000, 3_000_001), n // 20))
    p = int(rng.integers(n // 4, n // 2))
    letters[p:p + big] = ord("N")
    for _ in range(max(1, n // 20_000_000) + 2):
        g = int(rng.integers(edge, max(edge + 1, n - 60_000)))
        letters[g:g + min(50_000, n // 200)] = ord("N")
    # lower-case runs: alternating upper / lower stretches with geometric lengths (mean 300)
    n_runs = n // 300 + 16
    lens = rng.geometric(1.0 / 300.0, size=n_runs)
    mask = np.repeat(np.arange(n_runs, dtype=np.uint8) & 1, lens)[:n]
    if mask.size < n:
        mask = np.concatenate([mask, np.zeros(n - mask.size, np.uint8)])
    letters |= (mask << 5).astype(np.uint8)           # 'A' | 0x20 = 'a'
    return letters


def write_fasta(path, scale, width=60):
    lens = [int(m * 1e6 * scale) for m in HUMAN_MBP]
    sizes = []
    for name, n in zip(NAMES, lens):
        hdr = (">%s synthetic config-4 contig, %d bp\n" % (name, n)).encode()
        sizes.append(len(hdr) + n + (n + width - 1) // width)
    total = sum(sizes)
    mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.uint8, shape=(total,))
    off = 0
    for k, (name, n) in enumerate(zip(NAMES, lens)):
        hdr = (">%s synthetic config-4 contig, %d bp\n" % (name, n)).encode()
        mm[off:off + len(hdr)] = np.frombuffer(hdr, dtype=np.uint8)
        off += len(hdr)
        letters = contig_letters(n, k)
        full = n // width
        body = mm[off:off + full * (width + 1)].reshape(full, width + 1)
        body[:, :width] = letters[:full * width].reshape(full, width)
        body[:, width] = ord("\n")
        off += full * (width + 1)
        tail = n - full * width
        if tail:
            mm[off:off + tail] = letters[full * width:]
            mm[off + tail] = ord("\n")
            off += tail + 1
    assert off == total, (off, total)
    mm.flush()
    del mm
    return sum(lens)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--path", default="/dev/shm/dgrp_config4.npy")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from deepgrp_b200 import _lib, model, prediction
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t_gen = time.perf_counter()
    if rank == 0:
        bases = write_fasta(args.path, args.scale)
        json.dump({"bases": bases}, open(args.path + ".json", "w"))
    if world > 1:
        dist.barrier()
    t_gen = time.perf_counter() - t_gen
    bases = json.load(open(args.path + ".json"))["bases"]
    raw = np.load(args.path, mmap_mode="r")
    ctx = _lib.context(local_rank)
    weights = model.random_weights(342, 60, attention=True, seed=0)
    devnull = open(os.devnull, "wb")

    def step():
        st = prediction.predict_fasta_tsv_stream(weights, raw, "genome.fa", devnull, 50, 256, True, 50, 50,
                                                 rank=rank, world=world)
        return st["d2h_bytes"], st["rows"], st["records"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()                                    # warm-up: buffers, first-touch of the mapped file
    times = []
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        tsv_len, n_rows, n_rec = step()
        local = time.perf_counter() - t0
        barrier()
        times.append((time.perf_counter() - t0, local))
    wall = min(t[0] for t in times)
    stats = torch.tensor([float(tsv_len), float(n_rows), min(t[1] for t in times)], dtype=torch.float64, device="cuda")
    allstats = [torch.zeros_like(stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(allstats, stats)
    else:
        allstats = [stats]
    if rank == 0:
        per_rank = [[float(x) for x in s.cpu()] for s in allstats]
        line = {
            "metric": "bases classified/sec end-to-end", "unit": "Mbp/s", "value": bases / wall / 1e6,
            "n_gpus": world, "seconds": wall, "bases": bases, "records": n_rec, "scale": args.scale,
            "config": "BASELINE.json configs[3] in shape: 24 contigs proportional to the human chromosomes, "
                      "N runs, soft-masked lower case; defaults.toml architecture, random-init weights; "
                      "contig sharding, FASTA bytes in host memory -> TSV text in host memory",
            "tsv_bytes_total": sum(p[0] for p in per_rank), "rows_total": sum(p[1] for p in per_rank),
            "rank_seconds": [p[2] for p in per_rank], "file_bytes": int(raw.size), "generate_seconds": t_gen,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        for p in (args.path, args.path + ".json"):
            try:
                os.unlink(p)
            except OSError:
                pass


if __name__ == "__main__":
    main()
