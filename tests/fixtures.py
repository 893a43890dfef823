"""Named parity cases: how to build their inputs (deterministically) and which arguments to run.

tests/golden/make_golden.py runs the unmodified reference binary on every case and stores what it
produced in tests/golden/MANIFEST.json; the CPU tests hold the oracle port to that manifest and the GPU
tests hold the CUDA path to it (and, stage by stage, to the port).
"""
from __future__ import annotations

import dataclasses
import hashlib
import os
import sys
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from localhgt_b200 import synth  # noqa: E402


@dataclasses.dataclass
class Case:
    name: str
    data: str                      # which input set (see _DATA)
    k: int = 24
    e: int = 3
    seed: int = 1
    hit: float = 0.1
    match: float = 0.08
    sample: float = 2000000000.0
    max_peak: int = 1000000
    prebuilt_index: bool = False   # run once to build the index, then measure the second run (Q3)
    note: str = ""


CASES: List[Case] = [
    Case("base_k24", "base", note="6 x 50 kbp, 2 planted transfers, defaults"),
    Case("base_k20", "base", k=20, note="2^20-entry table: heavy hash collisions, thousands of peaks"),
    Case("base_k24_e4", "base", k=24, e=4, seed=5, hit=0.15, match=0.05, note="e not a multiple of 3: two draws per position"),
    Case("base_k31_e1", "base", k=31, e=1, seed=7),
    Case("base_k27_e5", "base", k=27, e=5, seed=11, hit=0.15, match=0.05),
    Case("half_build", "base", sample=0.5, note="ratio<100 with the index built in-run: sampling stream starts after 64 draws (Q3)"),
    Case("half_reuse", "base", sample=0.5, prebuilt_index=True, note="same with the index already on disk: stream starts at draw 0 (Q3)"),
    Case("noisy", "noisy", sample=0.7, note="short contig, N run, lower case, '/' and tab in headers, trimmed mate-1 reads (some < k), reads with N"),
    Case("shorts_bp", "shorts", seed=5, sample=1500000.0, note="100/400/k+1-bp contigs, fq2 shorter than fq1 (Q15), names with a space, bp-count sampling"),
    Case("fq2_longer", "fq2long", note="fq2 longer than fq1 in bytes: S1 ignores its tail (Q15); last line without newline"),
    Case("k32_default", "base", k=32, note="the default k; 2^32-entry tables"),
]

BY_NAME: Dict[str, Case] = {c.name: c for c in CASES}


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _pairs(seed, genomes, n, **kw):
    m1s, m2s = [], []
    for a, b in synth.simulate_pairs(seed, genomes, n, **kw):
        m1s.append(a); m2s.append(b)
    return np.concatenate(m1s), np.concatenate(m2s)


def _base(outdir: str):
    w = synth.make_workload(outdir, "base", seed=1, n_genomes=6, genome_len=50000, n_pairs=6000, n_events=2,
                            seg_len=(1000, 8000), sub_rate=0.005, read_n_rate=0.0005)
    return w.ref_fa, w.fq1, w.fq2


def _noisy(outdir: str):
    """Edge-heavy set (SURVEY Appendix B13 shape, smaller)."""
    rng = np.random.default_rng(np.random.PCG64(77))
    ref = synth.make_reference(1077, 6, 40000, jitter=0.5, short_contigs=(20,), lowercase_stretch=3000)
    at = {nm: i for i, nm in enumerate(ref.names)}
    ref.seqs[at["g2"]][15000:15037] = ord("N")
    ref.seqs[at["g4"]][5000:8000] = ref.seqs[at["g0"]][10000:13000]      # a 3 kbp inter-contig repeat
    ref.names[at["g1"]] = "g1/extra words"
    ref.names[at["g3"]] = "g3\ttabbed"
    ref.names[at["g4"]] = "g4 spaced name"
    real = [i for i, nm in enumerate(ref.names) if nm.startswith("g")]
    rec, don = real[:3], real[3:]
    sample, _ = synth.plant_hgt(3077, ref, rec, don, 3, (1000, 6000))
    fa = os.path.join(outdir, "noisy.fa")
    synth.write_fasta(fa, ref, width=70)
    up = [np.where(s >= 97, s - 32, s).astype(np.uint8) for s in sample]
    m1, m2 = _pairs(2077, up, 5000, sub_rate=0.004, indel_rate=0.001, n_rate=0.0005)
    names1, names2, s1, s2 = [], [], [], []
    for i in range(m1.shape[0]):
        a, b = m1[i].tobytes(), m2[i].tobytes()
        u = rng.random()
        if u < 0.05:
            a = a[: int(rng.integers(10, 150))]          # trimmed mate 1, some shorter than k
        elif u < 0.07:
            b = b[: int(rng.integers(10, 150))]
        names1.append(b"read%d/1" % i); names2.append(b"read%d/2" % i)
        s1.append(a); s2.append(b)
    fq1, fq2 = os.path.join(outdir, "noisy.1.fq"), os.path.join(outdir, "noisy.2.fq")
    synth.write_fastq_ragged(fq1, names1, s1)
    synth.write_fastq_ragged(fq2, names2, s2)
    return fa, fq1, fq2


def _shorts(outdir: str):
    rng = np.random.default_rng(np.random.PCG64(55))
    ref = synth.make_reference(1055, 6, 30000, jitter=0.3, short_contigs=(100, 400))
    ref.names.append("tiny_kplus1"); ref.seqs.append(synth.random_genome(rng, 25))     # k+1 at k=24
    real = [i for i, nm in enumerate(ref.names) if nm.startswith("g")]
    rec, don = real[:3], real[3:]
    sample, _ = synth.plant_hgt(3055, ref, rec, don, 2, (1000, 5000))
    # reads that saturate the 100- and 400-bp contigs too, so their windows turn good (Q11)
    sample = list(sample) + [np.tile(ref.seqs[i], 3) for i, nm in enumerate(ref.names) if nm.startswith("short")]
    fa = os.path.join(outdir, "shorts.fa")
    synth.write_fasta(fa, ref, width=60)
    m1, m2 = _pairs(2055, sample, 4000, sub_rate=0.003, indel_rate=0.0, frag_mean=260.0)
    names1, names2, s1, s2 = [], [], [], []
    for i in range(m1.shape[0]):
        a, b = m1[i].tobytes(), m2[i].tobytes()
        if rng.random() < 0.3:
            b = b[: int(rng.integers(40, 150))]          # trimmed mate 2 -> fq2 shorter than fq1
        names1.append(b"A00%d:7:H 1:N:0:ACGT" % i); names2.append(b"A00%d:7:H 2:N:0:ACGT" % i)
        s1.append(a); s2.append(b)
    fq1, fq2 = os.path.join(outdir, "shorts.1.fq"), os.path.join(outdir, "shorts.2.fq")
    synth.write_fastq_ragged(fq1, names1, s1)
    synth.write_fastq_ragged(fq2, names2, s2)
    return fa, fq1, fq2


def _fq2long(outdir: str):
    rng = np.random.default_rng(np.random.PCG64(66))
    ref = synth.make_reference(1066, 4, 40000)
    sample, _ = synth.plant_hgt(3066, ref, [0, 1], [2, 3], 2, (1000, 5000))
    fa = os.path.join(outdir, "fq2long.fa")
    synth.write_fasta(fa, ref)
    m1, m2 = _pairs(2066, sample, 4000, sub_rate=0.003, indel_rate=0.0)
    names1, names2, s1, s2 = [], [], [], []
    for i in range(m1.shape[0]):
        a, b = m1[i].tobytes(), m2[i].tobytes()
        if rng.random() < 0.25:
            a = a[: int(rng.integers(60, 150))]          # trimmed mate 1 -> fq2 longer than fq1
        names1.append(b"p%d/1" % i); names2.append(b"p%d/2" % i)
        s1.append(a); s2.append(b)
    fq1, fq2 = os.path.join(outdir, "fq2long.1.fq"), os.path.join(outdir, "fq2long.2.fq")
    synth.write_fastq_ragged(fq1, names1, s1, trailing_newline=False)
    synth.write_fastq_ragged(fq2, names2, s2, trailing_newline=False)
    return fa, fq1, fq2


_DATA = {"base": _base, "noisy": _noisy, "shorts": _shorts, "fq2long": _fq2long}
_made: Dict[str, tuple] = {}


def materialize(data: str, outdir: str):
    """Returns (ref.fa, fq1, fq2) for a data set, generating it once per outdir."""
    key = (data, outdir)
    if key not in _made:
        d = os.path.join(outdir, data)
        os.makedirs(d, exist_ok=True)
        _made[key] = _DATA[data](d)
    return _made[key]


def index_path(fa: str, k: int, e: int) -> str:
    return f"{fa}.k{k}.h{e}.index.dat"


def clean_outputs(fa: str) -> None:
    d = os.path.dirname(fa)
    for f in os.listdir(d):
        if f.startswith(os.path.basename(fa) + ".") and (f.endswith(".index.dat") or f.endswith(".genome.len.txt")):
            os.remove(os.path.join(d, f))
