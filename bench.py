#!/usr/bin/env python
"""bench.py — read pairs/s through the k-mer screen + peak extract (and index build Gbp/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg3|cfg2|mini]

A step = one pass of the hot path (LocalHGT extract_ref: FASTQ record location, S1 count of both mates,
S2 window/peak detection over the resident index, S3 pair confirmation, interval text) over one batch =
the whole synthetic sample of the workload.  The default workload is `cfg4`, the configuration BASELINE.json's
target is quoted on (configs[3]): a 5 Gbp synthetic reference (2 000 x 2.5 Mbp, half recipients, half absent
donors) and 30 M simulated 150 bp read pairs with planted transfers, k=32 e=3, LocalHGT's default arguments
(`--sample 2e9` => 22.2 % of the pairs are sampled).  It fits one B200 (index image 60 GB).  `cfg3` is configs[2]
(1 Gbp, 10 M pairs), `cfg2` configs[1] (80 Mbp, 5 M pairs, every pair sampled), `mini` a 1/100 cfg4 for quick checks.

  value      input pairs/s with the FASTQ bytes and the index already resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the C ABI with HOST (pinned) buffers, as a stream of samples against the resident index:
             both FASTQ images are copied host->device inside the timed region (sample i+1's while sample i is screened) and
             the interval text comes back device->host; `e2e_cold` carries nothing over: the reference FASTA crosses PCIe too
             and the 60 GB index image is re-built from it on the device inside the step
  roofline   the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference binary (oracle/_ref/extract_ref) on this box's host cores, on a bounded,
             self-consistent 1/f scale model of the workload (pairs and contigs), extrapolated linearly and said so
  --impl reference   times only that reference binary (no code of ours on its path)

N > 1 (torchrun): STRONG scaling -- the one sample is split N ways by record ranges (rank r screens pairs [lo_r, hi_r)),
the index is replicated, count tables are combined over NVLink peer memory, S2's table gather is sharded over reference
tiles, S3's verdicts are max-reduced.  The result is the N = 1 result bit for bit; rank 0 checks that once, outside the
timed region, by screening the whole sample alone.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K, E, SEED = 32, 3, 1
HIT, MATCH, MAX_PEAK, SAMPLE = 0.1, 0.08, 300000000, 2000000000.0      # scripts/localhgt.py:51-61
READ_LEN = 150
P = READ_LEN - K + 1
SECTOR = 32                                                             # bytes per random probe

# name: (n_genomes, genome_len, n_pairs, n_events).  cfg2 is made by the numpy generator (files on disk, so that the
# unmodified reference could be run on the very same bytes: KNOWN_ANSWERS); the others by the counter-based torch
# generator (localhgt_b200/synth_dev.py), on the GPU, sliceable by pair range.
WORKLOADS = {
    "cfg2": (40, 2_000_000, 5_000_000, 40),
    "cfg3": (500, 2_000_000, 10_000_000, 200),                          # BASELINE.json configs[2]
    "cfg4": (2000, 2_500_000, 30_000_000, 400),                         # BASELINE.json configs[3]: the headline
    "cfg5": (4000, 2_500_000, 10_000_000, 400),                         # BASELINE.json configs[4]: 10 Gbp, e = 1..4 sweep, image in blocks
    "mini": (20, 2_500_000, 300_000, 4),
    "small": (8, 500_000, 200_000, 6),
}
FILE_WORKLOADS = ("cfg2", "small")
# sha256 of the interval text the UNMODIFIED reference (oracle/_ref/extract_ref_z, -t 1) wrote for the workload's files
KNOWN_ANSWERS = {"cfg2": "1bd924af26dce59c383adc813ab9d123533b64d487ed010d33e17578d3efa8dd"}   # tests/golden/cfg2_reference.md


def _rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def split_range(n: int, parts: int, i: int):
    base, extra = divmod(n, parts)
    lo = i * base + min(i, extra)
    return lo, lo + base + (1 if i < extra else 0)


# ---------------------------------------------------------------------------------------------- workload
def workload_dir(name: str, shard: int = 0) -> str:
    base = os.environ.get("LHGT_BENCH_DIR", os.path.join(tempfile.gettempdir(), "lhgt_bench"))
    return os.path.join(base, f"{name}_s{shard}")


def make_workload(name: str, shard: int = 0):
    """File workloads (numpy generator): generates (or reuses) the synthetic files."""
    from localhgt_b200 import synth
    n_genomes, genome_len, n_pairs, n_events = WORKLOADS[name]
    d = workload_dir(name, shard)
    meta_path = os.path.join(d, "meta.json")
    fa, fq1, fq2 = (os.path.join(d, f"{name}.{x}") for x in ("fa", "1.fq", "2.fq"))
    if os.path.exists(meta_path):
        meta = json.load(open(meta_path))
        if all(os.path.exists(p) for p in (fa, fq1, fq2)) and meta.get("n_pairs") == n_pairs:
            return fa, fq1, fq2, meta
    os.makedirs(d, exist_ok=True)
    ref = synth.make_reference(1002, n_genomes, genome_len, n_rate=0.0001, short_contigs=(20,))
    real = [i for i, nm in enumerate(ref.names) if nm.startswith("g")]
    recipients, donors = real[: len(real) // 2], real[len(real) // 2:]
    sample, truth = synth.plant_hgt(3002, ref, recipients, donors, n_events, (1000, 50000))
    synth.write_fasta(fa, ref)
    sample_up = [np.where(s >= 97, s - 32, s).astype(np.uint8) for s in sample]
    n = synth.write_fastq_pair(fq1, fq2, synth.simulate_pairs(2002 + 7919 * shard, sample_up, n_pairs, read_len=READ_LEN,
                                                              sub_rate=0.01, indel_rate=0.001), read_len=READ_LEN)
    meta = {"n_pairs": n, "ref_bases": int(sum(len(s) for s in ref.seqs)), "n_contigs": len(ref.seqs),
            "truth": [[t.recipient, t.r_pos] for t in truth]}
    json.dump(meta, open(meta_path, "w"))
    return fa, fq1, fq2, meta


def spec_of(name: str, scale: int = 1):
    """synth_dev.Spec of a generated workload; scale = f gives its self-consistent 1/f model (pairs, contigs, events)."""
    from localhgt_b200 import synth_dev
    n_genomes, genome_len, n_pairs, n_events = WORKLOADS[name]
    return synth_dev.Spec(name if scale == 1 else f"{name}/{scale}", max(4, n_genomes // scale), genome_len, max(1000, n_pairs // scale),
                          max(2, n_events // scale), seed={"cfg3": 3, "cfg4": 4, "mini": 4, "cfg5": 5}.get(name, 9))


class Workload:
    """What one rank needs of a workload: the reference FASTA and ITS record range of the one sample, as device tensors."""

    def __init__(self, name: str, rank: int, world: int, torch, device):
        self.name, self.torch, self.device = name, torch, device
        if name in FILE_WORKLOADS:
            fa, fq1, fq2, meta = make_workload(name, 0)
            self.files = (fa, fq1, fq2)
            self.n_pairs, self.ref_bases, self.n_contigs = meta["n_pairs"], meta["ref_bases"], meta["n_contigs"]
            # meta's recipient index counts the 20-bp contig that follows g2; the interval file's ordinal does not (Q2)
            self.truth = [(rec + 1 if rec < 3 else rec, pos) for rec, pos in meta["truth"]]
            self.lo, self.hi = split_range(self.n_pairs, world, rank)
            self._fa = np.fromfile(fa, dtype=np.uint8)
            with open(fq1, "rb") as f:
                self.stride = len(f.readline() + f.readline() + f.readline() + f.readline())
            self.spec = None
        else:
            from localhgt_b200 import synth_dev
            self.spec = spec_of(name)
            self.n_pairs, self.ref_bases, self.n_contigs = self.spec.n_pairs, self.spec.ref_bases, self.spec.n_contigs
            self.truth = synth_dev.truth_positions(self.spec)
            self.lo, self.hi = split_range(self.n_pairs, world, rank)
            self.stride = self.spec.record_bytes
            self.files = None

    def _padded(self, n):
        return self.torch.empty(n + 64, dtype=self.torch.uint8, device=self.device)[:n]      # the kernels stage whole 16-byte granules

    def fasta_dev(self):
        if self.spec is None:
            t = self._padded(self._fa.size)
            t.copy_(self.torch.from_numpy(self._fa))
            return t
        from localhgt_b200 import synth_dev
        layout, total = synth_dev.fasta_layout(self.spec)
        t = self._padded(total)
        t.copy_(synth_dev.make_fasta(self.spec, self.device))
        return t

    def reads_dev(self, lo=None, hi=None):
        lo = self.lo if lo is None else lo
        hi = self.hi if hi is None else hi
        n = (hi - lo) * self.stride
        d1, d2 = self._padded(n), self._padded(n)
        if self.spec is None:
            for path, d in ((self.files[1], d1), (self.files[2], d2)):
                a = np.fromfile(path, dtype=np.uint8, count=n, offset=lo * self.stride)
                d.copy_(self.torch.from_numpy(a))
        else:
            from localhgt_b200 import synth_dev
            cat, offs = synth_dev.sample_genomes(self.spec, self.device)
            synth_dev.make_pairs(self.spec, cat, offs, lo, hi, d1, d2)
            del cat
        return d1, d2


def head_records(src: str, dst: str, n_records: int) -> None:
    """First n fixed-stride records of a generated FASTQ."""
    with open(src, "rb") as f:
        first = f.readline() + f.readline() + f.readline() + f.readline()
        stride = len(first)
        f.seek(0)
        with open(dst, "wb") as g:
            left = stride * n_records
            while left > 0:
                blk = f.read(min(left, 1 << 24))
                if not blk:
                    break
                g.write(blk)
                left -= len(blk)


def head_contigs(src: str, dst: str, n_contigs: int) -> int:
    bases, seen = 0, 0
    with open(src, "rb") as f, open(dst, "wb") as g:
        for ln in f:
            if ln.startswith(b">"):
                seen += 1
                if seen > n_contigs:
                    break
            else:
                bases += len(ln) - 1
            g.write(ln)
    return bases


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region, in-process through NVML (a forked `nvidia-smi -lms`
    loop takes driver locks every poll and visibly stretches the allocation-heavy host side of a step)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.samples, self.reasons = index, period_s, [], set()
        self.t, self.stop_flag, self.h, self.nv = None, threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            self.h, self.nv = pynvml.nvmlDeviceGetHandleByIndex(phys), pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None
            return
        self.t = threading.Thread(target=self._pump, daemon=True)
        self.t.start()

    def _pump(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def stop(self) -> dict:
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.t.join(timeout=2)
        sm = self.samples
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(self.reasons), "how": "NVML, %d ms period, inside the timed region" % int(1000 * self.period)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- reference arm
def ref_binary() -> str:
    from oracle import orc
    if not os.path.exists(orc.REF_BIN):
        if os.path.exists("/root/reference/src/extract_ref_normal_peak.cpp"):
            orc.build()
    if not os.path.exists(orc.REF_BIN):
        raise FileNotFoundError(orc.REF_BIN)
    return orc.REF_BIN


def run_reference_once(exe, fq1, fq2, fa, out, threads, sample=SAMPLE):
    argv = [exe, fq1, fq2, fa, out, repr(HIT), repr(MATCH), str(threads), str(K), str(MAX_PEAK), str(E), str(SEED),
            repr(float(sample))]
    t = time.perf_counter()
    r = subprocess.run(argv, capture_output=True, text=True)
    dt = time.perf_counter() - t
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stderr[-500:])
    return dt, r.stdout


def sampled_fraction(n_pairs: int) -> float:
    """E:1392-1398 with fixed-length reads: ratio = sample / (2 * bases of fq1)."""
    return min(1.0, SAMPLE / (2.0 * READ_LEN * n_pairs))


class CpuModel:
    """The bounded sample the reference binary is timed on, and the linear model that scales it to the full workload.

    CPU cost is linear in sampled pairs and in reference bases:
        T(run) = fixed + c_base * R + c_pair * sampled_pairs
    File workloads (cfg2): the first n pairs of the very files + the full reference (R is not scaled; it sits in `fixed`).
    Generated workloads (cfg3/cfg4): a self-consistent scale model from the same generator -- a `small` reference of 4
    contigs with `n` pairs drawn from ITS recipients (same read model, same sampled fraction, passed as a fraction so that
    the smaller file does not change it), and a `big` reference of 40 contigs that is only slid (S2) to measure c_base.
    """

    def __init__(self, name: str, n_pairs: int):
        self.name, self.n = name, n_pairs
        self.full_pairs = WORKLOADS[name][2]
        self.full_bases = WORKLOADS[name][0] * WORKLOADS[name][1]
        self.frac = sampled_fraction(self.full_pairs)
        self.dir = os.path.join(workload_dir(name, 0), f"cpu_{n_pairs}")
        os.makedirs(self.dir, exist_ok=True)
        j = lambda x: os.path.join(self.dir, x)
        self.fq = (j("s.1.fq"), j("s.2.fq"))
        self.tiny = (j("tiny.1.fq"), j("tiny.2.fq"))
        self.small_fa, self.big_fa = j("small.fa"), j("big.fa")
        self.generated = name not in FILE_WORKLOADS
        if self.generated:
            self._generate()
            self.sample_arg = self.frac if self.frac < 1 else SAMPLE
        else:
            fa, fq1, fq2, meta = make_workload(name, 0)
            self.n = min(self.n, meta["n_pairs"])
            if not os.path.exists(self.fq[1]):
                head_records(fq1, self.fq[0], self.n); head_records(fq2, self.fq[1], self.n)
            for link in (self.small_fa,):
                if not os.path.lexists(link):
                    os.symlink(fa, link)
            self.big_fa = None
            self.small_bases, self.big_bases = meta["ref_bases"], 0
            self.sample_arg = SAMPLE
        if not os.path.exists(self.tiny[1]):
            head_records(self.fq[0], self.tiny[0], 4); head_records(self.fq[1], self.tiny[1], 4)

    def _generate(self):
        import torch
        from localhgt_b200 import synth_dev
        g, L, _, _ = WORKLOADS[self.name]
        seed = spec_of(self.name).seed
        small = synth_dev.Spec(self.name + "/model", 4, L, self.n, 2, seed=seed)
        big = synth_dev.Spec(self.name + "/model-big", min(g, 40), L, 1000, 2, seed=seed)
        self.small_bases, self.big_bases = small.ref_bases, big.ref_bases
        if not os.path.exists(self.small_fa):
            synth_dev.make_fasta(small, "cpu").numpy().tofile(self.small_fa)
        if not os.path.exists(self.big_fa):
            synth_dev.make_fasta(big, "cpu").numpy().tofile(self.big_fa)
        if not os.path.exists(self.fq[1]):
            cat, offs = synth_dev.sample_genomes(small, "cpu")
            o1 = torch.empty(self.n * small.record_bytes, dtype=torch.uint8); o2 = torch.empty_like(o1)
            synth_dev.make_pairs(small, cat, offs, 0, self.n, o1, o2, chunk=1 << 16)
            o1.numpy().tofile(self.fq[0]); o2.numpy().tofile(self.fq[1])

    @staticmethod
    def index_of(fa):
        return f"{fa}.k{K}.h{E}.index.dat"

    def describe(self, cores):
        if self.generated:
            return (f"scale model of {self.name}: {self.n} pairs ({self.frac:.4f} of them sampled, as in the full workload) from the recipients of a "
                    f"{self.small_bases} bp / 4-contig reference made by the same generator, S2 cost per base from sliding a {self.big_bases} bp / "
                    f"40-contig reference; unmodified reference binary -t {cores}, index files present; value = full workload "
                    f"({self.full_pairs} pairs, {self.full_bases} bp) under T = fixed + c_base*R + c_pair*sampled_pairs fitted to these runs")
        return (f"first {self.n} of {self.full_pairs} pairs of the workload's own files, full {self.small_bases} bp reference, unmodified reference "
                f"binary -t {cores}, index file present; value = full workload under T = fixed + c_pair*pairs fitted to these runs")

    def estimate(self, t_tiny_small, t_pairs_small, t_tiny_big=None):
        """Returns (pairs/s of the full workload, model dict)."""
        sampled_model = self.n * (self.frac if self.generated else 1.0)
        c_pair = max(t_pairs_small - t_tiny_small, 1e-9) / sampled_model
        if self.generated and t_tiny_big is not None:
            c_base = max(t_tiny_big - t_tiny_small, 0.0) / max(self.big_bases - self.small_bases, 1)
            fixed = max(t_tiny_small - c_base * self.small_bases, 0.0)
            total = fixed + c_base * self.full_bases + c_pair * self.frac * self.full_pairs
        else:
            c_base, fixed = None, t_tiny_small
            total = fixed + c_pair * self.full_pairs
        return self.full_pairs / total, {"fixed_s": fixed, "c_base_s_per_base": c_base, "c_pair_s_per_sampled_pair": c_pair,
                                         "full_workload_seconds": total, "t_tiny_small": t_tiny_small, "t_pairs_small": t_pairs_small,
                                         "t_tiny_big": t_tiny_big}


def reference_arm(args) -> None:
    rank, _, world = _rank_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        exe = ref_binary()
    except Exception as ex:  # cannot happen where oracle/_ref travelled with the snapshot
        print(json.dumps({"impl": "reference", "unavailable": str(ex)[:200]}))
        return
    steps, warm = args.steps, args.warmup
    n = args.ref_pairs or (400_000 if steps + warm <= 3 else 200_000 if steps + warm <= 6 else 100_000)
    m = CpuModel(args.workload, n)
    out = os.path.join(m.dir, "ref.interval.txt")
    # the reference builds its own indexes (single-threaded by construction, E:1409): timed as its IB figure
    ib = None
    t_build = None
    if not os.path.exists(m.index_of(m.small_fa)):
        t_build, _ = run_reference_once(exe, *m.tiny, m.small_fa, out + ".ib", cores, m.sample_arg)
    t_tiny_small, _ = run_reference_once(exe, *m.tiny, m.small_fa, out + ".ib", cores, m.sample_arg)
    if t_build is not None:
        ib = {"gbp_per_s": m.small_bases / 1e9 / max(t_build - t_tiny_small, 1e-9), "seconds": t_build - t_tiny_small, "bases": m.small_bases,
              "threads": 1, "how": "wall(run that builds the index) - wall(run that reuses it), 4 read pairs"}
    t_tiny_big = None
    if m.generated:
        if not os.path.exists(m.index_of(m.big_fa)):
            run_reference_once(exe, *m.tiny, m.big_fa, out + ".ib", cores, m.sample_arg)
        t_tiny_big, _ = run_reference_once(exe, *m.tiny, m.big_fa, out + ".ib", cores, m.sample_arg)
    times = []
    for i in range(warm + steps):
        dt, _ = run_reference_once(exe, *m.fq, m.small_fa, out, cores, m.sample_arg)
        if i >= warm:
            times.append(dt)
    total = sum(times)
    value, model = m.estimate(t_tiny_small, total / steps, t_tiny_big)
    line = {
        "impl": "reference", "metric": "read pairs/sec through k-mer screen+peak extract", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000 * total / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_dict(args.workload),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": m.describe(cores), "model": model,
                         "measured_on_sample": {"pairs_per_s": m.n / (total / steps), "seconds_per_step": total / steps}},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if ib:
        line["index_build"] = ib
    print(json.dumps(line))


def config_dict(name):
    g, L, n_pairs, _ = WORKLOADS[name]
    frac = sampled_fraction(n_pairs)
    return {"workload": f"{name}: synthetic {g} x {L} bp reference ({g * L} bp) + ONE sample of {n_pairs} simulated {READ_LEN} bp read pairs "
                        f"with planted HGT breakpoints ({frac:.4f} of the pairs sampled by --sample 2e9); N GPUs split the sample N ways",
            "k": K, "e": E, "seed": SEED, "hit_ratio": HIT, "match_ratio": MATCH, "sample": SAMPLE, "max_peak": MAX_PEAK,
            "pairs": n_pairs, "ref_bases": g * L, "sampled_fraction": frac,
            "l2": "inputs larger than L2 (no flush needed): FASTQ images %.1f GB, index image %.1f GB, count table 1 GiB, peak table 16 GiB" %
                  (n_pairs * 2 * 318 / 1e9, g * L * 4 * E / 1e9)}


# ---------------------------------------------------------------------------------------------- our arm
def ours(args) -> None:
    import torch
    from localhgt_b200 import api, multi

    rank, local_rank, world = _rank_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (liblhgt has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    t_setup = time.perf_counter()
    if args.workload in FILE_WORKLOADS and dist:                          # one rank writes the files, the others wait for them
        if rank == 0:
            make_workload(args.workload, 0)
        dist.barrier()
    wl = Workload(args.workload, rank, world, torch, dev)
    d_fa = wl.fasta_dev()
    d1, d2 = wl.reads_dev()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t_setup

    blocks = world > 1 and (args.image_blocks or args.workload == "cfg5")    # every rank keeps one block of the index image (row e-S)
    sweep = index_sweep(args, api, torch, dist, d_fa, wl, rank, local_rank, world) if args.workload == "cfg5" else None
    image_gb = wl.ref_bases * 4 * E / 1e9 / (world if blocks else 1)
    if image_gb > 100:
        if rank == 0:
            print(json.dumps({"metric": "read pairs/sec through k-mer screen+peak extract", "value": None, "unit": "pairs/s", "n_gpus": world,
                              "config": config_dict(args.workload), "index_sweep": sweep,
                              "skipped": f"a {image_gb:.0f} GB index image per GPU does not fit beside the tables and the sample: run {args.workload} "
                                         f"with --gpus >= 2 (the image is then kept in blocks, one per GPU)"}))
        if dist:
            dist.destroy_process_group()
        return

    stream = torch.cuda.Stream()
    scr = api.Screen(K, E, device=local_rank)
    cc, skip = api.random_coder(SEED, K, E)
    scr.set_coder(cc)
    scr.set_s1_mode(args.s1_mode)
    if blocks:
        scr.set_image_block(rank, world)
    size1_total = wl.n_pairs * wl.stride                                   # E:1419: size(fq1) of the whole sample (Q15 budget)

    with torch.cuda.stream(stream):
        scr.set_stream(stream.cuda_stream)
        # ---- index build (the IB half of the metric): kernel time, and FASTA-on-device -> image-on-device time
        ib_ms_kernel, ib_ms_dev = [], []
        for i in range(3):
            torch.cuda.synchronize()
            t = time.perf_counter()
            scr.index_build_device(d_fa.data_ptr(), d_fa.numel())
            scr.sync()
            ib_ms_dev.append(1000 * (time.perf_counter() - t))
            ib_ms_kernel.append(float(scr.stage_ms()[5]))
        index_bases, index_bytes = scr.index_bases(), scr.index_bytes()
        index_build = {"gbp_per_s": index_bases / 1e6 / min(ib_ms_kernel), "kernel_ms": min(ib_ms_kernel),
                       "device_gbp_per_s": index_bases / 1e6 / min(ib_ms_dev), "device_ms": min(ib_ms_dev), "bases": index_bases,
                       "index_bytes": int(index_bytes),
                       "device_how": "FASTA text resident in HBM -> header scan, sequence compaction, hashing -> index image resident in HBM (wall clock of lhgt_index_build_device)",
                       "roofline": {"bound": "hbm", "achieved": index_bases * (1 + 4 * E) / 1e6 / min(ib_ms_kernel),
                                    "unit": "GB/s", "bytes_per_base": 1 + 4 * E}}

        shard = multi.Shard(scr, rank, world, dist, torch, same_stream=True)     # scr launches on `stream`, torch's current stream
        solo = multi.Shard(scr, 0, 1, None, torch, same_stream=True) if world > 1 else shard

        def step_resident():
            scr.reads_attach_device(0, d1.data_ptr(), d1.numel())
            scr.reads_attach_device(1, d2.data_ptr(), d2.numel())
            return shard.screen(size1=d1.numel(), sample_arg=SAMPLE, seed=SEED, rand_skip=0, hit=HIT, match=MATCH, max_peak=MAX_PEAK)

        # ---- N > 1: the answer of the sharded run must be the answer of the whole sample on one GPU
        text_whole = None
        if world > 1:
            text_sharded = step_resident()
            if rank == 0 and not args.no_whole_check and not blocks:
                w1, w2 = wl.reads_dev(0, wl.n_pairs)
                scr.reads_attach_device(0, w1.data_ptr(), w1.numel()); scr.reads_attach_device(1, w2.data_ptr(), w2.numel())
                text_whole = solo.screen(size1=w1.numel(), sample_arg=SAMPLE, seed=SEED, rand_skip=0, hit=HIT, match=MATCH, max_peak=MAX_PEAK)
                del w1, w2
                torch.cuda.empty_cache()
                assert text_whole == text_sharded, "the sharded screen and the single-GPU screen of the same sample disagree"
            dist.barrier()

        def timed(fn, steps, warm, sampler=None):
            for _ in range(warm):
                fn()
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            if sampler:
                sampler.start()
            l0 = scr.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stage = np.zeros(13)
            a.record(stream)
            res = None
            for _ in range(steps):
                res = fn()
                stage += shard.last_stage_ms
            b.record(stream)
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            clocks = sampler.stop() if sampler else None
            ms = a.elapsed_time(b)
            if dist:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, res, stage / steps, scr.launch_count() - l0, clocks

        sampler = ClockSampler(local_rank) if rank == 0 else None
        ms, res, stage, launches, clocks = timed(step_resident, args.steps, args.warmup, sampler)
        text_resident = res
        wall_resident = dict(shard.last_wall_ms)
        counts_resident = dict(shard.last_counts)

        # ---- e2e: host (pinned) FASTQ + FASTA -> result.  The resident copies go first (HBM budget at cfg4).
        h1 = torch.empty(d1.numel(), dtype=torch.uint8, pin_memory=True); h1.copy_(d1)
        h2 = torch.empty(d2.numel(), dtype=torch.uint8, pin_memory=True); h2.copy_(d2)
        hfa = torch.empty(d_fa.numel(), dtype=torch.uint8, pin_memory=True); hfa.copy_(d_fa)
        torch.cuda.synchronize()
        n1, n2, nfa = d1.numel(), d2.numel(), d_fa.numel()
        del d1, d2, d_fa
        torch.cuda.empty_cache()
        # N > 1: the FASTA crosses PCIe once per BOX, 1/N of it per rank (every rank holds the bytes in pinned memory), and
        # NVLink carries the slices to everybody (in-place all-gather of equal slices)
        fa_slice = -(-nfa // world)
        fa_bcast = torch.empty(fa_slice * world + 64, dtype=torch.uint8, device=dev) if world > 1 else None

        def build_index_e2e():
            if world == 1:
                scr.index_build_ptr(hfa.data_ptr(), nfa)                # adopts the prefetch issued at the top of the step
                return
            lo, hi = rank * fa_slice, min(nfa, (rank + 1) * fa_slice)
            if hi > lo:
                fa_bcast[lo:hi].copy_(hfa[lo:hi], non_blocking=True)
            dist.all_gather_into_tensor(fa_bcast[: fa_slice * world], fa_bcast[rank * fa_slice:(rank + 1) * fa_slice])
            scr.index_build_device(fa_bcast.data_ptr(), nfa)

        def step_e2e():
            """A sample from pinned host buffers against the RESIDENT index (the reference, too, builds its index once and reuses
            the file for every sample, E:1401-1417): the sample's FASTQ images are adopted -- the previous step started their
            host->device copies into the alternate buffers while it was screening -- and the next sample's copies are started."""
            scr.reads_upload_ptr(0, h1.data_ptr(), n1)
            scr.reads_upload_ptr(1, h2.data_ptr(), n2)
            scr.reads_prefetch_next_ptr(0, h1.data_ptr(), n1)
            scr.reads_prefetch_next_ptr(1, h2.data_ptr(), n2)
            return shard.screen(size1=n1, size2=n2, sample_arg=SAMPLE, seed=SEED, rand_skip=0, hit=HIT, match=MATCH, max_peak=MAX_PEAK)

        def step_e2e_cold():
            """Everything from host buffers, nothing carried over: FASTQ x2 and the reference FASTA cross PCIe in the order the
            stages need them (fq2 lands behind S1 of fq1, the FASTA behind S1 of fq2) and the index image is re-built on the device."""
            scr.reads_prefetch_ptr(0, h1.data_ptr(), n1)
            scr.reads_prefetch_ptr(1, h2.data_ptr(), n2)
            if world == 1:
                scr.fasta_prefetch_ptr(hfa.data_ptr(), nfa)
            scr.reads_upload_ptr(0, h1.data_ptr(), n1)
            return shard.screen(size1=n1, size2=n2, sample_arg=SAMPLE, seed=SEED, rand_skip=0, hit=HIT, match=MATCH, max_peak=MAX_PEAK,
                                before_mate2=lambda: scr.reads_upload_ptr(1, h2.data_ptr(), n2), before_s2=build_index_e2e)

        cold_steps = max(1, min(args.steps, 3))
        ms_cold, res_cold, stage_cold, _, _ = timed(step_e2e_cold, cold_steps, 1)
        assert res_cold == text_resident, "resident and cold host-buffer passes disagree"
        ms_e2e, res_e2e, stage_e2e, _, _ = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))
        assert res_e2e == text_resident, "resident and host-buffer passes disagree"

    if rank != 0:
        scr.close()
        if dist:
            dist.destroy_process_group()
        return

    sha = hashlib.sha256(text_resident).hexdigest()
    known = KNOWN_ANSWERS.get(args.workload)
    if known:
        assert sha == known, f"interval text of {args.workload} differs from the unmodified reference's ({sha} != {known})"
    value = wl.n_pairs * args.steps / (ms / 1000)
    e2e_value = wl.n_pairs * args.steps / (ms_e2e / 1000)
    peak, peak_src = measured_peak_gbs()
    names = ["fastq_record_scan", "s1_count", "s2_gather", "s2_finish", "s3_pairs", "index_build", "exchange", "host_setup",
             "s1_hash_streams", "s1_split_streams", "s1_apply_leaves", "s2_register", "s3_vote"]
    stage_ms = {nm: round(float(v), 3) for nm, v in zip(names, stage)}
    frac = sampled_fraction(wl.n_pairs)
    roofline = make_roofline(stage, wl, frac, world, peak, peak_src, n1 + n2)
    roofline["stage_ms_per_step"] = stage_ms
    h2d = n1 + n2
    h2d_cold = n1 + n2 + (nfa if world == 1 else -(-nfa // world))
    line = {
        "metric": "read pairs/sec through k-mer screen+peak extract", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_dict(args.workload),
        "parallelism": {"plan": f"one sample split {world} ways by record ranges, index replicated" if world > 1 else "1 GPU",
                        "count_exchange": ("one kernel over NVLink peer memory (CUDA IPC)" if shard.p2p else "NCCL all-to-all + merge + all-gather")
                        if world > 1 else "none",
                        "index_image": (f"in {world} blocks, one per GPU ({image_gb:.1f} GB each)" if blocks else "replicated" if world > 1 else "whole"),
                        "whole_sample_check": ("rank 0 screened the whole sample alone: identical text" if text_whole is not None else
                                               "skipped" if world > 1 else "n/a")},
        "sampled_pairs_per_s": {"value": frac * value, "e2e": frac * e2e_value, "sampled_fraction": frac},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "pairs/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": len(text_resident) + 64,
                "pcie_gbs": h2d / 1e6 / (ms_e2e / args.steps),
                "what": "a stream of samples against the resident index (the reference builds its index once and reuses the file): pinned host "
                        "FASTQ x2 -> HBM -> S1,S2,S3 -> interval text on host, through the C ABI; each sample's copies run on the copy stream "
                        "into alternate buffers while the previous sample is screened (lhgt_reads_prefetch_next), every byte of every step "
                        "crosses PCIe inside the timed region (rank 0's figure)",
                "stage_ms_per_step": {nm: round(float(v), 3) for nm, v in zip(names, stage_e2e)}},
        "e2e_cold": {"value": wl.n_pairs * cold_steps / (ms_cold / 1000), "unit": "pairs/s", "ms_per_step": ms_cold / cold_steps, "steps": cold_steps,
                     "h2d_bytes_per_step": int(h2d_cold), "pcie_gbs": h2d_cold / 1e6 / (ms_cold / cold_steps),
                     "what": "nothing carried over between steps: FASTQ x2 + the reference FASTA cross PCIe (copy stream, overlapping S1) and the "
                             "index image is re-built on the device inside the step (at N > 1 the FASTA crosses PCIe once per box, 1/N per rank, "
                             "and NVLink all-gathers the slices)",
                     "stage_ms_per_step": {nm: round(float(v), 3) for nm, v in zip(names, stage_cold)}},
        "gpu_launches": int(launches), "roofline": roofline, "index_build": index_build,
        "result": {"interval_lines": len(text_resident.splitlines()), "interval_sha256": sha,
                   "matches_unmodified_reference": bool(known) or None,
                   "planted_recovered": recovered(text_resident, wl.truth), "peaks": shard.last_peaks,
                   "sampled": counts_resident},
        "host_wall_ms_last_step": {k: round(v, 3) for k, v in shard.last_wall_ms.items()},
        "host_wall_ms_last_resident_step": {k: round(v, 3) for k, v in wall_resident.items()},
        "setup_seconds": {"generate_inputs": round(gen_s, 2)},
    }
    if sweep:
        line["index_sweep"] = sweep
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = cpu_baseline(args, scr)
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "failed: " + str(ex)[:200]}
    scr.close()
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def index_sweep(args, api, torch, dist, d_fa, wl, rank, local_rank, world):
    """BASELINE.json configs[4]: index build over the 10 Gbp reference for e = 1..4, the image kept in `world` blocks (one per
    GPU).  Per e: best of 3 builds, device time of the hashing kernel and wall time of lhgt_index_build_device, max over ranks."""
    out = []
    peak, _ = measured_peak_gbs()
    for e in (1, 2, 3, 4):
        gb = wl.ref_bases * 4 * e / 1e9 / world
        if gb > 120:
            out.append({"e": e, "skipped": f"{gb:.0f} GB of image per GPU: needs more GPUs"})
            continue
        scr = api.Screen(K, e, device=local_rank)
        cc, _ = api.random_coder(SEED, K, e)
        scr.set_coder(cc)
        if world > 1:
            scr.set_image_block(rank, world)
        kern, wall = [], []
        for _ in range(3):
            if dist:
                dist.barrier()
            torch.cuda.synchronize()
            t = time.perf_counter()
            scr.index_build_device(d_fa.data_ptr(), d_fa.numel())
            scr.sync()
            wall.append(1000 * (time.perf_counter() - t))
            kern.append(float(scr.stage_ms()[5]))
        k_ms, w_ms = min(kern), min(wall)
        if dist:
            tt = torch.tensor([k_ms, w_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            k_ms, w_ms = float(tt[0]), float(tt[1])
        bases, image_bytes = scr.index_bases(), scr.index_bytes()
        scr.close()
        torch.cuda.empty_cache()
        out.append({"e": e, "index_bytes": int(image_bytes), "image_gb_per_gpu": round(gb, 1), "kernel_ms": k_ms, "gbp_per_s": bases / 1e6 / k_ms,
                    "device_ms": w_ms, "device_gbp_per_s": bases / 1e6 / w_ms,
                    "roofline": {"bytes_per_base": 1 + 4 * e, "achieved_gbs_per_gpu": bases * (1 + 4 * e) / 1e6 / k_ms / world,
                                 "frac": bases * (1 + 4 * e) / 1e6 / k_ms / world / peak}})
    return out


def make_roofline(stage, wl, frac, world, peak, peak_src, fastq_bytes):
    """Roofline of the dominant kernel (by device time inside the timed region, rank 0), against the measured HBM copy bandwidth.

    Algorithmic bytes.  SURVEY §8(d) counts a table probe as one 32-byte DRAM sector:
      S1         : sampled reads x P x e probes x 32 B, one launch (set) per mate
      S3         : sampled pairs x 2 x P x e probes x 32 B
      S2 gather  : reference bases x (4e stored-hash bytes + 32e)
    With hash streams S1 is three kernels whose own DRAM bytes are different (that is the point of the design):
      s1_bin_kernel   : sampled FASTQ bytes read + 4 B per hash written
      s1_split_kernel : 4 B per hash read + 4 B per hash written
      s1_leaf_kernel  : 4 B per hash read + the 2^k x 2 bit table read and written back once
    For those the figure under the 32 B/probe convention is reported next to it as `probe_convention`; it can exceed the
    HBM peak because no probe goes to DRAM.  Units are this rank's share (pairs / world, tiles / world).
    """
    pairs_rank = wl.n_pairs / world
    probes_per_mate = pairs_rank * frac * P * E
    ref_share = wl.ref_bases / world
    traffic = load_traffic()
    kernels = {}
    streams = stage[8] > 0 or stage[9] > 0 or stage[10] > 0
    if streams:
        table_bytes = (1 << K) // 4
        kernels["s1_bin_kernel<3>"] = (stage[8] / 2, fastq_bytes * frac / 2 + 4 * probes_per_mate, probes_per_mate * SECTOR)
        kernels["s1_split_kernel"] = (stage[9] / 2, 8 * probes_per_mate, probes_per_mate * SECTOR)
        kernels["s1_leaf_kernel"] = (stage[10] / 2, 4 * probes_per_mate + 2 * table_bytes, probes_per_mate * SECTOR)
    else:
        kernels["s1_count_kernel<3>"] = (stage[1] / 2, probes_per_mate * SECTOR, probes_per_mate * SECTOR)
    s3_scan = max(float(stage[4] - stage[12]), 0.0)                     # stage[4] covers the scan launches and the vote launches
    kernels["s3_pairs_kernel<3>"] = (s3_scan, 2 * probes_per_mate * SECTOR, 2 * probes_per_mate * SECTOR)
    kernels["s2_gsemit+gsapply"] = (stage[2], ref_share * (E * SECTOR + 4 * E), ref_share * (E * SECTOR + 4 * E))
    share = {"s1_bin_kernel<3>": stage[8], "s1_split_kernel": stage[9], "s1_leaf_kernel": stage[10], "s1_count_kernel<3>": stage[1],
             "s3_pairs_kernel<3>": s3_scan, "s2_gsemit+gsapply": stage[2]}
    dom = max(kernels, key=lambda k: share[k])
    per_kernel = {}
    for k, (ms, nbytes, conv) in kernels.items():
        if ms <= 0:
            continue
        per_kernel[k] = {"ms_per_launch": round(ms, 4), "ms_per_step": round(float(share[k]), 3), "algorithmic_bytes_per_launch": int(nbytes),
                         "achieved": nbytes / 1e6 / ms, "frac": nbytes / 1e6 / ms / peak,
                         "probe_convention": {"bytes_per_launch": int(conv), "achieved": conv / 1e6 / ms, "frac": conv / 1e6 / ms / peak}}
    d = per_kernel[dom]
    device_ms = float(stage[0] + stage[1] + stage[2] + stage[3] + stage[4] + stage[6] + stage[11])
    overhead = {"fastq_record_scan": float(stage[0]), "s2_mark_windows_ids": float(stage[3]), "s2_register": float(stage[11]),
                "s3_vote": float(stage[12]), "exchange": float(stage[6])}
    whole_bytes = pairs_rank * frac * 2 * 2 * P * E * SECTOR + ref_share * (E * SECTOR + 4 * E)
    return {"bound": "hbm", "kernel": dom, "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"],
            "traffic": (traffic or {}).get(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_launch"],
            "ms_per_launch": d["ms_per_launch"], "kernels": per_kernel,
            "per_launch_note": "figures are per STEP for stages that run as several equal launches (S3 scan: one per arena batch, S2 gather: "
                               "an emit + apply pair per record chunk, S1: one set per mate -> halved); bytes and time scale together, so achieved "
                               "and frac are those of a single launch; traffic (profiles/traffic.json) is per step too",
            "stages_outside_the_byte_model_ms": {k: round(v, 3) for k, v in overhead.items()},
            "whole_step": {"algorithmic_bytes": int(whole_bytes), "device_ms": device_ms, "achieved": whole_bytes / 1e6 / max(device_ms, 1e-9),
                           "frac": whole_bytes / 1e6 / max(device_ms, 1e-9) / peak,
                           "what": "SURVEY 8(d): sampled pairs x 45 696 B + reference bases x 108 B over the summed device time of the step's stages"},
            "s1_mode": "hash streams (two-level split, table slices updated in shared memory)" if streams else "direct probes"}


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
    (profiles/traffic.json: {kernel name: bytes})."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def recovered(text: bytes, truth) -> str:
    """Planted recipient junctions strictly inside an emitted interval with 50 bp margin (paper_results/evaluation.py:64-76)."""
    by_contig = {}
    for ln in text.splitlines():
        c, a, b = map(int, ln.split(b"\t"))
        by_contig.setdefault(c, []).append((a, b))
    found = sum(any(a + 50 < pos < b - 50 for a, b in by_contig.get(ordinal, ())) for ordinal, pos in truth)
    return f"{found}/{len(truth)}"


def cpu_baseline(args, scr):
    """Times the unmodified reference binary on a bounded sample (rank 0, N=1) and scales it with CpuModel."""
    exe = ref_binary()
    cores = os.cpu_count() or 1
    m = CpuModel(args.workload, args.ref_pairs or 200_000)
    out = os.path.join(m.dir, "cpu.interval.txt")
    # index files: the bit-identical images our IB writes (tests/ prove the equality); the reference 'detects' and reuses them
    ib = None
    small_idx = m.index_of(m.small_fa)
    if m.generated and not os.path.exists(small_idx):                     # the small reference: let the reference build it once = its IB figure
        t_build, _ = run_reference_once(exe, *m.tiny, m.small_fa, out + ".ib", cores, m.sample_arg)
        t_reuse, _ = run_reference_once(exe, *m.tiny, m.small_fa, out + ".ib", cores, m.sample_arg)
        ib = {"gbp_per_s": m.small_bases / 1e9 / max(t_build - t_reuse, 1e-9), "bases": m.small_bases, "threads": 1,
              "how": "wall(run that builds the index) - wall(run that reuses it), 4 read pairs"}
    for fa in (m.small_fa, m.big_fa):
        if fa and not os.path.exists(m.index_of(fa)):
            scr.index_build_file(fa, m.index_of(fa), fa + ".genome.len.txt")
    t_tiny_small, _ = run_reference_once(exe, *m.tiny, m.small_fa, out, cores, m.sample_arg)
    t_pairs_small, _ = run_reference_once(exe, *m.fq, m.small_fa, out, cores, m.sample_arg)
    t_tiny_big = run_reference_once(exe, *m.tiny, m.big_fa, out, cores, m.sample_arg)[0] if m.generated else None
    value, model = m.estimate(t_tiny_small, t_pairs_small, t_tiny_big)
    res = {"value": value, "unit": "pairs/s", "cores": cores, "kind": "reference", "sample": m.describe(cores), "model": model,
           "measured_on_sample": {"pairs_per_s": m.n / t_pairs_small, "seconds": t_pairs_small}}
    if ib:
        res["index_build"] = ib
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs in the CPU reference's bounded sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-whole-check", action="store_true", help="N > 1: skip rank 0's single-GPU screen of the whole sample")
    ap.add_argument("--image-blocks", action="store_true", help="N > 1: every GPU keeps one block of the index image instead of a replica (cfg5 always does)")
    ap.add_argument("--s1-mode", type=int, default=0, help="0 auto (hash streams for tables > 64 MiB), 1 direct probes, 2 streams")
    args = ap.parse_args()
    _, _, world = _rank_env()
    if args.impl == "ours" and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29541"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
