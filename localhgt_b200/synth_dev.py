"""Counter-based synthetic workloads generated with torch tensor ops (on the GPU for the big bench shapes).

Same recipe as synth.py (paper_results/simulation.py:201-306, 819-832: recipients carry donor segments, donors are
absent from the sample, 150 bp pairs from ~N(350,10) fragments with substitutions and indels), but every random
quantity is a pure function of (seed, global index): a splitmix64 hash evaluated with wrapping int64 arithmetic,
integer-only, so that

  * the bytes do not depend on the device (CPU torch == CUDA torch), on the chunk size or on which rank makes them:
    rank r of N generates pairs [lo, hi) of THE SAME sample (strong scaling: one sample split N ways), and the
    reference arm of bench.py can make the first n pairs / first c contigs of the same workload on the host;
  * a 5 Gbp reference + 30 M pairs (BASELINE.json configs[3]) take seconds on a B200 instead of tens of minutes of numpy.

This module is data plumbing.  It contains no part of the screened path.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    return (z >> s) & ((1 << (64 - s)) - 1)


def mix(x: torch.Tensor, salt: int) -> torch.Tensor:
    """splitmix64 finaliser of (x + salt * golden), int64 with wrap-around.  Top bit may be set: callers take bit fields."""
    z = x + _s64(salt * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)
    z = z + _C1
    z = (z ^ _lsr(z, 30)) * _C2
    z = (z ^ _lsr(z, 27)) * _C3
    return z ^ _lsr(z, 31)


def mix_py(x: int, salt: int) -> int:
    """The same function on a Python int (for host-side choices and tests)."""
    z = (x + salt * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & _M64
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _lut(vals: bytes, device) -> torch.Tensor:
    return torch.tensor(list(vals), dtype=torch.uint8, device=device)


def _comp_lut(device) -> torch.Tensor:
    t = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
        t[a] = b
    return torch.from_numpy(t).to(device)


@dataclasses.dataclass
class Planted:
    recipient: int      # contig ordinal among the `g` contigs (0-based)
    r_pos: int
    donor: int
    d_start: int
    d_end: int
    reverse: bool


@dataclasses.dataclass
class Spec:
    """A workload shape.  Contig i (i < n_genomes) is `g<i>`, contig_len bases; one 20-base contig `short20` follows g2
    (skipped by the index but counted).  The first half of the genomes are recipients, the rest donors."""
    name: str
    n_genomes: int
    genome_len: int
    n_pairs: int
    n_events: int
    seed: int
    read_len: int = 150
    n_rate_bits: int = 1678          # of 2^24: 1e-4 of the reference bases are 'N'
    sub_rate_bits: int = 167772      # of 2^24: 1 % substitutions (a quarter of them silent)
    indel_rate_bits: int = 157286    # of 2^20: 0.15 = 0.001 * 150 of the reads carry one indel
    seg_len: Tuple[int, int] = (1000, 50000)
    digits: int = 9

    @property
    def record_bytes(self) -> int:
        return (2 + self.digits + 2) + 1 + self.read_len + 3 + self.read_len + 1

    @property
    def ref_bases(self) -> int:
        return self.n_genomes * self.genome_len + (20 if self.n_genomes > 2 else 0)

    @property
    def n_contigs(self) -> int:
        return self.n_genomes + (1 if self.n_genomes > 2 else 0)


def contig_bases(spec: Spec, i: int, device, lo: int = 0, hi: int = -1) -> torch.Tensor:
    """ASCII bases [lo, hi) of genome i."""
    hi = spec.genome_len if hi < 0 else hi
    p = torch.arange(lo, hi, dtype=torch.int64, device=device) + (i << 32)
    z = mix(p, spec.seed * 4 + 1)
    out = _lut(b"ACGT", device)[(z & 3)]
    is_n = (_lsr(z, 8) & 0xFFFFFF) < spec.n_rate_bits
    return torch.where(is_n, torch.full_like(out, ord("N")), out)


def short_contig(spec: Spec, device) -> torch.Tensor:
    p = torch.arange(0, 20, dtype=torch.int64, device=device) + (0x7FFF << 32)
    return _lut(b"ACGT", device)[(mix(p, spec.seed * 4 + 1) & 3)]


def fasta_layout(spec: Spec, width: int = 80):
    """[(name, header offset, first sequence byte offset, bases)] and the file size."""
    out, at = [], 0
    order: List[Tuple[str, int]] = []
    for i in range(spec.n_genomes):
        order.append((f"g{i}", spec.genome_len))
        if i == 2 and spec.n_genomes > 2:
            order.append(("short20", 20))
    for name, n in order:
        h = len(name) + 2
        body = n + (n + width - 1) // width
        out.append((name, at, at + h, n))
        at += h + body
    return out, at


def make_fasta(spec: Spec, device, n_contigs: int = -1, width: int = 80) -> torch.Tensor:
    """The reference FASTA (80-column lines) as a uint8 tensor on `device`; n_contigs >= 0 keeps the first that many records."""
    layout, total = fasta_layout(spec, width)
    if n_contigs >= 0:
        layout = layout[:n_contigs]
        total = 0 if not layout else layout[-1][2] + layout[-1][3] + (layout[-1][3] + width - 1) // width
    buf = torch.empty(total, dtype=torch.uint8, device=device)
    gi = 0
    for name, at, seq_at, n in layout:
        hdr = (">" + name + "\n").encode()
        buf[at:seq_at] = _lut(hdr, device)
        if name.startswith("g"):
            s = contig_bases(spec, gi, device)
            gi += 1
        else:
            s = short_contig(spec, device)
        full = n // width
        if full:
            body = buf[seq_at:seq_at + full * (width + 1)].view(full, width + 1)
            body[:, :width] = s[:full * width].view(full, width)
            body[:, width] = 10
        if n % width:
            tail_at = seq_at + full * (width + 1)
            buf[tail_at:tail_at + n % width] = s[full * width:]
            buf[tail_at + n % width] = 10
    return buf


def plant(spec: Spec) -> List[Planted]:
    """Host-side choice of the planted transfers (a few hundred events): pure function of the spec."""
    rec_n = spec.n_genomes // 2
    don_n = spec.n_genomes - rec_n
    ev: List[Planted] = []
    L = spec.genome_len
    hi_seg = min(spec.seg_len[1], L // 2)
    for j in range(spec.n_events):
        h = [mix_py(j * 8 + q, spec.seed * 4 + 2) for q in range(6)]
        r = h[0] % rec_n
        d = rec_n + h[1] % don_n
        ln = spec.seg_len[0] + h[2] % (hi_seg - spec.seg_len[0] + 1)
        ds = h[3] % (L - ln)
        rp = 2000 + h[4] % (L - 4000)
        ev.append(Planted(r, rp, d, ds, ds + ln, bool(h[5] & 1)))
    return ev


def sample_genomes(spec: Spec, device, recipients: Sequence[int] = None) -> Tuple[torch.Tensor, np.ndarray]:
    """The sequenced sample: recipients with the donor segments inserted, concatenated (upper case ASCII, 'N' kept).
    Returns (cat, offsets[len(recipients) + 1])."""
    rec_n = spec.n_genomes // 2
    recipients = list(range(rec_n)) if recipients is None else list(recipients)
    events = plant(spec)
    comp = _comp_lut(device)
    parts, offs = [], [0]
    for r in recipients:
        base = contig_bases(spec, r, device)
        evs = sorted((e for e in events if e.recipient == r), key=lambda e: e.r_pos)
        last, total = 0, 0
        for e in evs:
            seg = contig_bases(spec, e.donor, device, e.d_start, e.d_end)
            if e.reverse:
                seg = comp[seg.flip(0).long()]
            parts += [base[last:e.r_pos], seg]
            total += e.r_pos - last + seg.numel()
            last = e.r_pos
        parts.append(base[last:])
        total += base.numel() - last
        offs.append(offs[-1] + total)
    return torch.cat(parts), np.asarray(offs, dtype=np.int64)


def _digits(idx: torch.Tensor, digits: int) -> torch.Tensor:
    cols = [(torch.div(idx, 10 ** (digits - 1 - d), rounding_mode="floor") % 10 + 48).to(torch.uint8) for d in range(digits)]
    return torch.stack(cols, dim=1)


def make_pairs(spec: Spec, cat: torch.Tensor, offs: np.ndarray, lo: int, hi: int, out1: torch.Tensor, out2: torch.Tensor,
               chunk: int = 1 << 20) -> None:
    """Writes the FASTQ records of pairs [lo, hi) of the sample into out1 / out2 (uint8, (hi - lo) * record_bytes each)."""
    dev = cat.device
    L, rb = spec.read_len, spec.record_bytes
    offs_t = torch.from_numpy(offs).to(dev)
    T = int(offs[-1])
    ar = torch.arange(L, dtype=torch.int64, device=dev)
    acgt, comp = _lut(b"ACGT", dev), _comp_lut(dev)
    salt = spec.seed * 4 + 3
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        m = c1 - c0
        n = torch.arange(c0, c1, dtype=torch.int64, device=dev)
        # fragment length ~ N(350, 10): Irwin-Hall sum of twelve 16-bit uniforms, integer arithmetic only
        s = torch.zeros(m, dtype=torch.int64, device=dev)
        for q in range(3):
            z = mix(n * 16 + q, salt)
            for f in range(4):
                s = s + (_lsr(z, 16 * f) & 0xFFFF)
        frag = 350 + ((10 * (s - 6 * 65536) + 32768) >> 16)
        frag = torch.clamp(frag, min=L + 8)
        z = mix(n * 16 + 3, salt)
        pos = ((_lsr(z, 11) & ((1 << 53) - 1)).double() * (float(T) / float(1 << 53))).long().clamp_(max=T - 1)
        g = torch.searchsorted(offs_t, pos, right=True) - 1
        g_lo, g_hi = offs_t[g], offs_t[g + 1]
        frag = torch.minimum(frag, g_hi - g_lo - 16)
        start = torch.maximum(torch.minimum(pos, g_hi - frag - 8), g_lo)
        for mate, out in ((0, out1), (1, out2)):
            if mate == 0:
                idx = start[:, None] + ar[None, :]
            else:
                idx = (start + frag - 1)[:, None] - ar[None, :]
            z = mix(n * 16 + 4 + mate, salt)
            has = (z & 0xFFFFF) < spec.indel_rate_bits
            q = 1 + (_lsr(z, 20) & 0xFFFF) % (L - 2)
            is_del = (_lsr(z, 40) & 1) == 1
            step = 1 if mate == 0 else -1
            shift = (ar[None, :] >= q[:, None]) & has[:, None]
            idx = idx + torch.where(is_del, step, -step)[:, None] * shift
            seq = cat[idx]
            if mate == 1:
                seq = comp[seq.long()]
            zz = mix(((n * 2 + mate) << 8)[:, None] + ar[None, :], salt + 1)
            sub = (_lsr(zz, 8) & 0xFFFFFF) < spec.sub_rate_bits
            seq = torch.where(sub, acgt[(zz & 3)], seq)
            rec = out[(c0 - lo) * rb:(c1 - lo) * rb].view(m, rb)
            c = 2 + spec.digits + 2
            rec[:, 0] = ord("@")
            rec[:, 1] = ord("r")
            rec[:, 2:2 + spec.digits] = _digits(n, spec.digits)
            rec[:, 2 + spec.digits] = ord("/")
            rec[:, 3 + spec.digits] = 49 + mate
            rec[:, c] = 10
            rec[:, c + 1:c + 1 + L] = seq
            rec[:, c + 1 + L] = 10
            rec[:, c + 2 + L] = ord("+")
            rec[:, c + 3 + L] = 10
            rec[:, c + 4 + L:c + 4 + 2 * L] = ord("I")
            rec[:, c + 4 + 2 * L] = 10


def truth_positions(spec: Spec):
    """[(contig ordinal in the interval file (1-based, indexed contigs only), recipient position)] of the planted junctions."""
    return [(e.recipient + 1, e.r_pos) for e in plant(spec)]
