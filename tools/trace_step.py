#!/usr/bin/env python
"""Host wall-clock per C-ABI call of one resident screening step (where does a step's time go between kernels?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from localhgt_b200 import api

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
fa, fq1, fq2, meta = bench.make_workload(name, 0)
b1 = np.fromfile(fq1, dtype=np.uint8); b2 = np.fromfile(fq2, dtype=np.uint8); fasta = np.fromfile(fa, dtype=np.uint8)
scr = api.Screen(bench.K, bench.E)
cc, _ = api.random_coder(bench.SEED, bench.K, bench.E)
scr.set_coder(cc)
scr.index_build(fasta)
d1 = torch.from_numpy(b1).cuda(); d2 = torch.from_numpy(b2).cuda()
torch.cuda.synchronize()


def T(label, fn, *a):
    t = time.perf_counter()
    r = fn(*a)
    scr.sync()
    print(f"  {label:28s} {1000 * (time.perf_counter() - t):9.3f} ms")
    return r


for it in range(3):
    print("step", it)
    t0 = time.perf_counter()
    T("reads_attach_device(0)", scr.reads_attach_device, 0, d1.data_ptr(), d1.numel())
    T("reads_attach_device(1)", scr.reads_attach_device, 1, d2.data_ptr(), d2.numel())
    T("reset", scr.reset)
    T("set_sampling", scr.set_sampling, 100.0, 1, 0)
    T("s1_count(0)", scr.s1_count, 0, d1.numel())
    T("s1_count(1)", scr.s1_count, 1, d1.numel())
    T("s2_peaks", scr.s2_peaks, bench.HIT, bench.MATCH, bench.MAX_PEAK)
    T("s3_pairs", scr.s3_pairs)
    T("intervals", scr.intervals)
    print(f"  {'total':28s} {1000 * (time.perf_counter() - t0):9.3f} ms   stage_ms {np.round(scr.stage_ms(), 3).tolist()}")
