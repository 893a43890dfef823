"""scripts/get_bed_file.py and `samtools faidx -r` (pipeline.sh:36-37) as host-side functions of
liblhgt.  The .bed text is pinned by tests/golden/BED.json (the unmodified reference script run on the golden interval
files); the FASTA cutter is checked against a plain-Python statement of faidx's documented format (samtools is not
available here: parity unpinned, as include/lhgt.h says)."""
import json
import os

import numpy as np
import pytest

from localhgt_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
BED = json.load(open(os.path.join(HERE, "golden", "BED.json")))
CASES = sorted(k for k in BED if not k.startswith("_"))


@pytest.mark.parametrize("name", CASES)
def test_bed_text_matches_reference_script(name):
    rec = BED[name]
    iv, ln = rec["interval_text"].encode(), rec["len_text"].encode()
    if rec["returncode"] != 0:                                   # the script died (KeyError on an unlisted ref_index)
        with pytest.raises(api.LhgtError):
            api.bed_text(iv, ln)
        return
    bed, total = api.bed_text(iv, ln)
    assert bed.decode() == rec["bed_text"]
    assert f"extracted ref length is: {total}\n" == rec["stdout"]


def _faidx_model(fasta: bytes, bed: bytes) -> bytes:
    seqs, name = {}, None
    for line in fasta.split(b"\n"):
        line = line.rstrip(b"\r")
        if line.startswith(b">"):
            name = line[1:].split()[0] if line[1:].split() else b""
            seqs[name] = []
        elif line and name is not None:
            seqs[name].append(line)
    seqs = {k: b"".join(v) for k, v in seqs.items()}
    out = []
    for region in bed.split(b"\n"):
        if not region.strip():
            continue
        nm, span = region.rsplit(b":", 1)
        a, b = span.split(b"-", 1)
        s = seqs[nm][max(int(a), 1) - 1:int(b)]
        out.append(b">" + region + b"\n")
        out.extend(s[i:i + 60] + b"\n" for i in range(0, len(s), 60))
    return b"".join(out)


@pytest.mark.parametrize("width", [80, 60, 7, 1000])
def test_regions_fasta_format(width):
    rng = np.random.default_rng(width)
    acgt = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    lens = {b"g0": 5000, b"g1 description here": 777, b"tiny": 3, b"g3": 12345}
    fasta = b""
    for nm, n in lens.items():
        s = acgt[rng.integers(0, len(acgt), n)].tobytes()
        fasta += b">" + nm + b"\n" + b"\n".join(s[i:i + width] for i in range(0, n, width)) + b"\n"
    bed = b"g0:1-60\ng0:2-61\ng0:4000-5500\ng1:1-777\ng1:700-9999\ntiny:1-3\ng3:61-120\ng3:12345-12345\ng3:1-12345\n"
    assert api.regions_fasta(fasta, bed) == _faidx_model(fasta, bed)
    with pytest.raises(api.LhgtError):
        api.regions_fasta(fasta, b"nope:1-5\n")


def test_files_round_trip(tmp_path):
    rec = BED["base_k24"]
    fa = tmp_path / "ref.fa"
    names = [ln.split("\t")[0] for ln in rec["len_text"].splitlines()]
    lens = [int(ln.split("\t")[2]) for ln in rec["len_text"].splitlines()]
    rng = np.random.default_rng(1)
    text = b""
    for nm, n in zip(names, lens):
        s = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, n)].tobytes()
        text += b">" + nm.encode() + b"\n" + b"\n".join(s[i:i + 80] for i in range(0, n, 80)) + b"\n"
    fa.write_bytes(text)
    (tmp_path / "ref.fa.genome.len.txt").write_text(rec["len_text"])
    iv = tmp_path / "s.interval.txt"
    iv.write_text(rec["interval_text"])
    total = api.extract_regions_files(str(fa), str(iv), str(tmp_path / "s.specific.ref.fasta"))
    assert (tmp_path / "s.interval.txt.bed").read_text() == rec["bed_text"]
    assert f"extracted ref length is: {total}\n" == rec["stdout"]
    assert (tmp_path / "s.specific.ref.fasta").read_bytes() == _faidx_model(text, rec["bed_text"].encode())
