"""torchrun worker of tests/test_multi_gpu.py: one process per GPU runs multi.Shard with the REAL GpuEngine on its record
range of a fixture; rank 0 compares the sharded result with the scalar oracle's run over the whole files.

usage: torchrun --nproc-per-node N tests/multi_worker.py <workdir> <case> <p2p|nccl>
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import fixtures  # noqa: E402
from localhgt_b200 import api, multi  # noqa: E402
from oracle import orc  # noqa: E402


def record_starts(buf: bytes):
    """Byte offsets at which the 4-line records of a FASTQ image start."""
    starts, pos, line = [0] if buf else [], 0, 0
    while True:
        nl = buf.find(b"\n", pos)
        if nl < 0:
            break
        pos = nl + 1
        line += 1
        if line % 4 == 0 and pos < len(buf):
            starts.append(pos)
    return starts


def main():
    work, name, form = sys.argv[1], sys.argv[2], sys.argv[3]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = fixtures.BY_NAME[name]
    fa, fq1, fq2 = fixtures.materialize(case.data, os.path.join(work, f"r{rank}"))     # deterministic: every rank makes the same bytes
    b1, b2 = open(fq1, "rb").read(), open(fq2, "rb").read()
    st1, st2 = record_starts(b1), record_starts(b2)
    n1, n2 = len(st1), len(st2)
    # pairs are matched by record ordinal: rank r takes records [o_r, o_r+1) of both files; the last rank also takes
    # whatever fq2 holds beyond fq1's record count
    ords = [multi.split_range(n1, world, r)[0] for r in range(world)] + [n1]

    def at(starts, o, size):
        return starts[o] if o < len(starts) else size

    lo1, hi1 = at(st1, ords[rank], len(b1)), at(st1, ords[rank + 1], len(b1))
    lo2 = at(st2, ords[rank], len(b2))
    hi2 = len(b2) if rank == world - 1 else at(st2, ords[rank + 1], len(b2))
    m1, m2 = b1[lo1:hi1], b2[lo2:hi2]
    k, e = case.k, case.e
    cc, skip = api.random_coder(case.seed, k, e)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream), api.Screen(k, e, device=local) as s:
        s.set_stream(stream.cuda_stream)
        s.set_coder(cc)
        if os.environ.get("LHGT_TEST_BLOCKS"):                 # every rank keeps one block of the image only (row e-S)
            s.set_image_block(rank, world)
        s.index_build(open(fa, "rb").read())
        if os.environ.get("LHGT_TEST_BLOCKS"):
            off, nbytes, t0, t1 = s.index_block()
            assert world == 1 or nbytes < s.index_bytes()
        if os.environ.get("LHGT_TEST_S1_MODE"):
            s.set_s1_mode(int(os.environ["LHGT_TEST_S1_MODE"]))
        shard = multi.Shard(s, rank, world, dist, torch, same_stream=True, force_nccl=(form == "nccl"))
        assert form == "nccl" or shard.p2p, "peer-memory exchange unavailable on this box"
        for rep in range(2):                                   # twice: the second pass runs on re-used buffers
            s.reads_upload(0, m1); s.reads_upload(1, m2)
            text = shard.screen(size1=len(m1), sample_arg=case.sample, seed=case.seed, rand_skip=skip, hit=case.hit, match=case.match,
                                max_peak=case.max_peak)
        table = s.count_table()
        loci, filt = s.peaks()
        counts = shard.last_counts
        tot = torch.tensor([counts["s1"][0], counts["s1"][1], counts["s3"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        texts = [None] * world
        dist.all_gather_object(texts, text)
        assert all(t == text for t in texts), "ranks disagree on the interval text"
        if rank == 0:
            o = orc.Oracle(k, e)
            o.srand(case.seed); o.random_coder()
            idx, lenp = os.path.join(work, "whole.index.dat"), os.path.join(work, "whole.len.txt")
            assert o.index_build(fa, idx, lenp) == 0
            ratio = orc.sample_ratio(fq1, case.sample)
            assert abs(ratio - counts["ratio"]) < 1e-9 * max(1.0, ratio)
            if ratio < 100:
                o.fill_random(max(n1, n2) + 8)
            s1a, s1b = o.s1_count(fq1, len(b1), ratio), o.s1_count(fq2, len(b1), ratio)
            assert np.array_equal(o.count_table(), table), "count table differs from the oracle's"
            npk = o.s2_peaks(idx, case.hit, case.match, case.max_peak)
            assert npk == shard.last_peaks and np.array_equal(o.peak_loci(), loci)
            s3 = o.s3_pairs(fq1, fq2, ratio)
            assert [int(x) for x in tot.tolist()] == [s1a, s1b, s3], (tot.tolist(), s1a, s1b, s3)
            assert np.array_equal(o.peak_filter() >= 1, filt >= 1), "verdicts differ from the oracle's"
            out = os.path.join(work, "whole.interval.txt")
            o.write_intervals(out)
            assert open(out, "rb").read() == text, "interval text differs from the oracle's"
            o.close()
            print(f"OK {name} {form} world={world} peaks={npk} kept={int((filt >= 1).sum())} exchange={'p2p' if shard.p2p else 'nccl'}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
