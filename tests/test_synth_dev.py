"""The counter-based workload generator (localhgt_b200/synth_dev.py): its bytes must not depend on chunking, on which rank
makes which pair range, or on the device -- bench.py's strong-scaling split and its CPU scale model rely on that."""
import hashlib

import numpy as np
import pytest
import torch

from localhgt_b200 import synth_dev as sd

SPEC = sd.Spec("t", 8, 60000, 20000, 5, seed=5)


def _pairs(spec, device, lo, hi, chunk):
    cat, offs = sd.sample_genomes(spec, device)
    n = (hi - lo) * spec.record_bytes
    o1 = torch.empty(n, dtype=torch.uint8, device=device); o2 = torch.empty_like(o1)
    sd.make_pairs(spec, cat, offs, lo, hi, o1, o2, chunk=chunk)
    return o1.cpu().numpy(), o2.cpu().numpy()


def test_fasta_layout_and_contig_identity():
    fa = sd.make_fasta(SPEC, "cpu").numpy().tobytes()
    layout, total = sd.fasta_layout(SPEC)
    assert total == len(fa)
    lines = fa.split(b"\n")
    names = [ln[1:].decode() for ln in lines if ln.startswith(b">")]
    assert names == [f"g{i}" for i in range(3)] + ["short20"] + [f"g{i}" for i in range(3, 8)]
    assert all(len(ln) <= 80 for ln in lines if not ln.startswith(b">"))
    # the first contigs of a larger spec with the same seed are the same bytes (the CPU scale model relies on it)
    big = sd.Spec("big", 12, 60000, 10, 2, seed=5)
    assert torch.equal(sd.contig_bases(big, 2, "cpu"), sd.contig_bases(SPEC, 2, "cpu"))
    assert sd.mix_py(12345, 7) == int(sd.mix(torch.tensor([12345]), 7)[0]) & ((1 << 64) - 1)


def test_pairs_do_not_depend_on_chunking_or_range():
    a1, a2 = _pairs(SPEC, "cpu", 0, 5000, 1 << 20)
    b1, b2 = _pairs(SPEC, "cpu", 3000, 5000, 777)
    rb = SPEC.record_bytes
    assert np.array_equal(a1[3000 * rb:], b1) and np.array_equal(a2[3000 * rb:], b2)
    rec = a1[:rb].tobytes().split(b"\n")
    assert rec[0] == b"@r000000000/1" and len(rec[1]) == 150 and rec[2] == b"+" and rec[3] == b"I" * 150
    assert a2[:rb].tobytes().startswith(b"@r000000000/2\n")
    # reads are substrings of the sample genomes up to ~1 % substitutions
    cat, _ = sd.sample_genomes(SPEC, "cpu")
    hay = cat.numpy().tobytes()
    hits = sum(hay.find(a1[i * rb + 14:i * rb + 14 + 40].tobytes()) >= 0 for i in range(0, 5000, 50))
    assert hits >= 60


@pytest.mark.gpu
def test_device_bytes_equal_host_bytes():
    a = sd.make_fasta(SPEC, "cpu")
    b = sd.make_fasta(SPEC, "cuda").cpu()
    assert torch.equal(a, b)
    c1, c2 = _pairs(SPEC, "cpu", 100, 4100, 1 << 20)
    g1, g2 = _pairs(SPEC, "cuda", 100, 4100, 1500)
    assert hashlib.sha256(c1).digest() == hashlib.sha256(g1).digest()
    assert hashlib.sha256(c2).digest() == hashlib.sha256(g2).digest()
