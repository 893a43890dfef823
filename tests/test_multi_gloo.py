"""CPU, world_size 2, gloo: the sharded plan of localhgt_b200/multi.py (ordinal bases, the Q15 byte budget in
shard-local offsets, the count-table all-to-all + merge + all-gather, the verdict max-reduce) driven with the
oracle as the compute engine, against one oracle run over the concatenated sample."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fixtures
from localhgt_b200 import multi, synth
from oracle import orc

K, E, SEED = 20, 3, 1


class OracleEngine:
    """The operations multi.Shard needs, computed by the scalar oracle on CPU (test infrastructure)."""

    sharded_s2 = False

    def __init__(self, fq1, fq2, index_path, workdir):
        self.fq1, self.fq2, self.index_path, self.workdir = fq1, fq2, index_path, workdir
        self.o = None
        self.ratio = 100.0

    def reset(self):
        if self.o:
            self.o.close()
        self.o = orc.Oracle(K, E)
        assert self.o.load_coder(self.index_path) == 0

    def _records(self, path):
        n = 0
        with open(path, "rb") as f:
            for _ in f:
                n += 1
        return (n + 2) // 4

    def reads_records(self, mate): return self._records(self.fq2 if mate else self.fq1)
    def reads_bytes(self, mate): return os.path.getsize(self.fq2 if mate else self.fq1)

    def reads_seq_bases(self, mate):
        tot = 0
        with open(self.fq2 if mate else self.fq1, "rb") as f:
            for i, ln in enumerate(f):
                if i % 4 == 1:
                    tot += len(ln.rstrip(b"\n"))
        return tot

    def set_ordinal_base(self, base): self.base = base

    def set_sampling(self, ratio, seed, skip):
        assert ratio >= 100, "the scalar oracle has no ordinal offset; shard tests run unsampled"
        self.ratio = ratio

    def s1_count(self, mate, budget): return self.o.s1_count(self.fq2 if mate else self.fq1, budget, self.ratio)
    def table(self): return torch.from_numpy(self.o.count_table())

    def merge_into(self, byte_offset, other):
        tab = self.o.count_table()
        seg = tab[byte_offset:byte_offset + other.numel()]
        np.minimum(3, seg.astype(np.int16) + other.numpy().astype(np.int16), out=seg, casting="unsafe")

    def sync(self): pass
    def s2_peaks(self, hit, match, max_peak): return self.o.s2_peaks(self.index_path, hit, match, max_peak)
    def s3_pairs(self): return self.o.s3_pairs(self.fq1, self.fq2, self.ratio)
    def peak_filter(self): return torch.from_numpy(self.o.peak_filter())

    def intervals(self):
        p = os.path.join(self.workdir, "iv.txt")
        assert self.o.write_intervals(p) == 0
        return open(p, "rb").read()

    def stage_ms(self): return np.zeros(6, dtype=np.float32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shards, index_path, workdir, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fq1, fq2 = shards[rank]
        wd = os.path.join(workdir, f"r{rank}")
        os.makedirs(wd, exist_ok=True)
        eng = OracleEngine(fq1, fq2, index_path, wd)
        sh = multi.Shard(None, rank, world, dist, torch, engine=eng)
        text = sh.screen(size1=os.path.getsize(fq1), sample_arg=2e9, seed=SEED, rand_skip=0, hit=0.1, match=0.08, max_peak=1000000)
        with open(out + f".{rank}", "wb") as f:
            f.write(text)
        with open(out + f".{rank}.meta", "w") as f:
            f.write(repr((sh.last_peaks, sh.last_counts)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("data,world", [("base", 2), ("fq2long", 2), ("base", 3)])
def test_rank_shards_equal_one_run(data, world, tmp_path):
    work = str(tmp_path)
    fa, fq1, fq2 = fixtures.materialize(data, work)
    idx, lenp = os.path.join(work, "ref.index.dat"), os.path.join(work, "ref.len.txt")
    o = orc.Oracle(K, E); o.srand(SEED); o.random_coder()
    assert o.index_build(fa, idx, lenp) == 0
    # one run over the whole sample
    size1 = os.path.getsize(fq1)
    o.s1_count(fq1, size1, 100.0); o.s1_count(fq2, size1, 100.0)
    n_peaks = o.s2_peaks(idx, 0.1, 0.08, 1000000)
    o.s3_pairs(fq1, fq2, 100.0)
    whole = os.path.join(work, "whole.txt")
    o.write_intervals(whole)
    want = open(whole, "rb").read()
    assert n_peaks > 10
    # unequal shards, in record order
    n_rec = sum(1 for _ in open(fq1, "rb")) // 4
    cuts = [0] + [n_rec * (2 * r + 2) // (2 * world + 1) for r in range(world - 1)] + [n_rec]
    lines1, lines2 = open(fq1, "rb").readlines(), open(fq2, "rb").readlines()
    shards = []
    for r in range(world):
        a, b = os.path.join(work, f"s{r}.1.fq"), os.path.join(work, f"s{r}.2.fq")
        hi1 = len(lines1) if r == world - 1 else 4 * cuts[r + 1]
        hi2 = len(lines2) if r == world - 1 else 4 * cuts[r + 1]
        open(a, "wb").write(b"".join(lines1[4 * cuts[r]:hi1]))
        open(b, "wb").write(b"".join(lines2[4 * cuts[r]:hi2]))
        shards.append((a, b))
    out = os.path.join(work, "sharded.txt")
    mp.spawn(_worker, args=(world, _free_port(), shards, idx, work, out), nprocs=world, join=True)
    texts = [open(out + f".{r}", "rb").read() for r in range(world)]
    assert all(t == want for t in texts)
    metas = [eval(open(out + f".{r}.meta").read()) for r in range(world)]
    assert all(m[0] == n_peaks for m in metas)
    assert [m[1]["ordinal_base"] for m in metas] == cuts[:-1]


def test_split_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 78125):
        for parts in (1, 2, 3, 8):
            spans = [multi.split_range(n, parts, i) for i in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
