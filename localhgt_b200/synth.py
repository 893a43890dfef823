"""Deterministic synthetic workloads for the k-mer screen (tests + bench.py).

Follows the recipe the reference's own evaluation uses (paper_results/simulation.py:201-306,
819-832): recipient genomes receive 1-50 kbp segments cut from donor genomes (half of them
reverse-complemented), the donor genomes themselves are ABSENT from the sequenced sample, reads are
150 bp pairs from ~N(350,10) fragments with substitutions and indels.  SURVEY.md §4 explains why the
donor must be absent (the screen keys on coverage edges).

Everything is numpy + a seeded PCG64 so the same (function, arguments) always gives the same bytes;
tests/golden/MANIFEST.json additionally pins the sha256 of every generated input so generator drift
is reported as such and never mistaken for a parity failure.

This module is data plumbing.  It contains no part of the screened path.
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


@dataclasses.dataclass
class Planted:
    """One planted transfer: donor[d_start:d_end] inserted into recipient at r_pos (reference coords)."""
    recipient: int      # 0-based contig ordinal in the reference FASTA
    r_pos: int          # insertion point on the recipient (0-based, reference coordinates)
    donor: int
    d_start: int
    d_end: int
    reverse: bool


@dataclasses.dataclass
class Reference:
    names: List[str]
    seqs: List[np.ndarray]          # uint8 ASCII per contig, as written to the FASTA


def random_genome(rng: np.random.Generator, n: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def make_reference(seed: int, n_contigs: int, contig_len: int, *, jitter: float = 0.0,
                   n_rate: float = 0.0, short_contigs: Sequence[int] = (),
                   lowercase_stretch: int = 0, name_fmt: str = "g{}") -> Reference:
    """n_contigs i.i.d. uniform-ACGT contigs.  Options keep the reference's edge paths hot:
    n_rate   - fraction of bases overwritten with 'N' (invalid k-mers, E:794-797,808-810)
    short_contigs - extra contigs of these lengths appended after every 3rd contig (len<=k are
                    skipped by the index but still counted)
    lowercase_stretch - lower-case the first that many bases of contig 1 (valid, E:1117-1151)
    """
    rng = np.random.default_rng(np.random.PCG64(seed))
    names, seqs = [], []
    shorts = list(short_contigs)
    for i in range(n_contigs):
        ln = contig_len if jitter == 0 else int(contig_len * (1 + jitter * (rng.random() - 0.5)))
        s = random_genome(rng, ln)
        if n_rate > 0:
            n_n = rng.binomial(ln, n_rate)
            if n_n:
                s[rng.integers(0, ln, size=n_n)] = ord("N")
        if lowercase_stretch and i == 1:
            s[:lowercase_stretch] |= 0x20
        names.append(name_fmt.format(i))
        seqs.append(s)
        if shorts and i % 3 == 2:
            ln_s = shorts.pop(0)
            names.append(f"short{ln_s}_{i}")
            seqs.append(random_genome(rng, ln_s))
    return Reference(names, seqs)


def write_fasta(path: str, ref: Reference, width: int = 80) -> None:
    with open(path, "wb") as f:
        for name, s in zip(ref.names, ref.seqs):
            f.write(b">" + name.encode() + b"\n")
            n = len(s)
            full = n // width
            if full:
                body = np.empty((full, width + 1), dtype=np.uint8)
                body[:, :width] = s[: full * width].reshape(full, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n % width:
                f.write(s[full * width:].tobytes() + b"\n")


def plant_hgt(seed: int, ref: Reference, recipients: Sequence[int], donors: Sequence[int],
              n_events: int, seg_len: Tuple[int, int] = (1000, 50000)) -> Tuple[List[np.ndarray], List[Planted]]:
    """Returns (sample genomes = recipients with donor segments inserted, truth list)."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    per_rec = {r: [] for r in recipients}
    truth: List[Planted] = []
    for _ in range(n_events):
        r = int(recipients[rng.integers(0, len(recipients))])
        d = int(donors[rng.integers(0, len(donors))])
        dl = len(ref.seqs[d])
        ln = int(rng.integers(seg_len[0], min(seg_len[1], dl // 2) + 1))
        ds = int(rng.integers(0, dl - ln))
        rp = int(rng.integers(2000, len(ref.seqs[r]) - 2000))
        ev = Planted(r, rp, d, ds, ds + ln, bool(rng.integers(0, 2)))
        per_rec[r].append(ev)
        truth.append(ev)
    sample = []
    for r in recipients:
        base = ref.seqs[r]
        evs = sorted(per_rec[r], key=lambda e: e.r_pos)
        parts, last = [], 0
        for ev in evs:
            parts.append(base[last:ev.r_pos])
            seg = ref.seqs[ev.donor][ev.d_start:ev.d_end]
            if ev.reverse:
                seg = _COMP[seg[::-1]]
            parts.append(seg)
            last = ev.r_pos
        parts.append(base[last:])
        sample.append(np.concatenate(parts))
    return sample, truth


def _names(first: int, n: int, mate: int, digits: int) -> np.ndarray:
    """(n, 2+digits+2) uint8 matrix of '@r000000123/1'."""
    out = np.empty((n, digits + 4), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    idx = np.arange(first, first + n, dtype=np.int64)
    for d in range(digits):
        out[:, 2 + d] = (idx // 10 ** (digits - 1 - d)) % 10 + 48
    out[:, 2 + digits] = ord("/")
    out[:, 3 + digits] = 48 + mate
    return out


def simulate_pairs(seed: int, genomes: Sequence[np.ndarray], n_pairs: int, *, read_len: int = 150,
                   frag_mean: float = 350.0, frag_sd: float = 10.0, sub_rate: float = 0.01,
                   indel_rate: float = 0.001, n_rate: float = 0.0, chunk: int = 200_000):
    """Yields (mate1, mate2) uint8 matrices of shape (m, read_len), ASCII, chunk by chunk."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    lens = np.array([len(g) for g in genomes], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)])
    cat = np.concatenate(genomes)
    p = lens / lens.sum()
    ar = np.arange(read_len, dtype=np.int64)
    done = 0
    while done < n_pairs:
        m = min(chunk, n_pairs - done)
        g = rng.choice(len(genomes), size=m, p=p)
        frag = np.clip(np.rint(rng.normal(frag_mean, frag_sd, size=m)).astype(np.int64),
                       read_len + 8, None)
        frag = np.minimum(frag, lens[g] - 16)
        start = (rng.random(m) * (lens[g] - frag - 8)).astype(np.int64) + offs[g]
        mates = []
        for mate in (0, 1):
            if mate == 0:
                idx = start[:, None] + ar[None, :]
            else:
                idx = (start + frag - 1)[:, None] - ar[None, :]
            # one indel per read at most (rate*len << 1): shift the gather index past position q
            if indel_rate > 0:
                has = rng.random(m) < indel_rate * read_len
                q = rng.integers(1, read_len - 1, size=m)
                is_del = rng.random(m) < 0.5
                step = 1 if mate == 0 else -1
                shift = (ar[None, :] >= q[:, None]) & has[:, None]
                idx = idx + np.where(is_del[:, None], step, -step) * shift
            seq = cat[idx]
            if mate == 1:
                seq = _COMP[seq]
            if sub_rate > 0:
                n_err = rng.binomial(m * read_len, sub_rate)
                pos = rng.integers(0, m * read_len, size=n_err)
                flat = seq.reshape(-1)
                flat[pos] = _ACGT[rng.integers(0, 4, size=n_err)]
            if n_rate > 0:
                n_n = rng.binomial(m * read_len, n_rate)
                seq.reshape(-1)[rng.integers(0, m * read_len, size=n_n)] = ord("N")
            mates.append(seq)
        yield mates[0], mates[1]
        done += m


def write_fastq_pair(path1: str, path2: str, pairs_iter, *, read_len: int = 150, digits: int = 9,
                     name_suffix=("/1", "/2")) -> int:
    """Fixed-stride FASTQ writer (vectorised).  Returns number of pairs written."""
    n = 0
    with open(path1, "wb") as f1, open(path2, "wb") as f2:
        for m1, m2 in pairs_iter:
            m = m1.shape[0]
            for mate, (f, seq) in enumerate(((f1, m1), (f2, m2))):
                nm = _names(n, m, mate + 1, digits)
                rec = np.empty((m, nm.shape[1] + 1 + read_len + 3 + read_len + 1), dtype=np.uint8)
                c = nm.shape[1]
                rec[:, :c] = nm
                rec[:, c] = 10
                rec[:, c + 1:c + 1 + read_len] = seq
                c += 1 + read_len
                rec[:, c] = 10
                rec[:, c + 1] = ord("+")
                rec[:, c + 2] = 10
                rec[:, c + 3:c + 3 + read_len] = ord("I")
                rec[:, c + 3 + read_len] = 10
                f.write(rec.tobytes())
            n += m
    return n


def write_fastq_ragged(path: str, names: Sequence[bytes], seqs: Sequence[bytes],
                       trailing_newline: bool = True) -> None:
    """Slow writer for edge-case fixtures (trimmed reads, odd names)."""
    with open(path, "wb") as f:
        for i, (nm, s) in enumerate(zip(names, seqs)):
            last = i == len(seqs) - 1
            f.write(b"@" + nm + b"\n" + s + b"\n+\n" + b"I" * len(s))
            if not last or trailing_newline:
                f.write(b"\n")


@dataclasses.dataclass
class Workload:
    ref_fa: str
    fq1: str
    fq2: str
    n_pairs: int
    ref_bases: int
    truth: List[Planted]


def make_workload(outdir: str, tag: str, *, seed: int, n_genomes: int, genome_len: int,
                  n_pairs: int, n_events: int, read_len: int = 150, sub_rate: float = 0.01,
                  indel_rate: float = 0.001, ref_n_rate: float = 0.0, read_n_rate: float = 0.0,
                  short_contigs: Sequence[int] = (), lowercase_stretch: int = 0,
                  seg_len: Tuple[int, int] = (1000, 50000), jitter: float = 0.0) -> Workload:
    """Half the genomes are recipients, half donors (absent from the sample), like BASELINE cfg 1/2."""
    os.makedirs(outdir, exist_ok=True)
    ref = make_reference(1000 + seed, n_genomes, genome_len, jitter=jitter, n_rate=ref_n_rate,
                         short_contigs=short_contigs, lowercase_stretch=lowercase_stretch)
    real = [i for i, nm in enumerate(ref.names) if nm.startswith("g")]
    recipients, donors = real[: len(real) // 2], real[len(real) // 2:]
    sample, truth = plant_hgt(3000 + seed, ref, recipients, donors, n_events, seg_len)
    fa = os.path.join(outdir, f"{tag}.fa")
    fq1 = os.path.join(outdir, f"{tag}.1.fq")
    fq2 = os.path.join(outdir, f"{tag}.2.fq")
    write_fasta(fa, ref)
    sample_up = [np.where(s >= 97, s - 32, s).astype(np.uint8) for s in sample]
    n = write_fastq_pair(fq1, fq2, simulate_pairs(2000 + seed, sample_up, n_pairs, read_len=read_len,
                                                  sub_rate=sub_rate, indel_rate=indel_rate,
                                                  n_rate=read_n_rate), read_len=read_len)
    return Workload(fa, fq1, fq2, n, int(sum(len(s) for s in ref.seqs)), truth)
