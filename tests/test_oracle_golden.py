"""CPU: the oracle port (oracle/lhgt_oracle.c) against what the unmodified reference binary produced
(tests/golden/MANIFEST.json, made by tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import pytest

import fixtures
from oracle import orc


def run_port(case, work):
    fa, fq1, fq2 = fixtures.materialize(case.data, work)
    fixtures.clean_outputs(fa)
    out = os.path.join(work, case.name + ".port.interval.txt")
    kw = dict(hit=case.hit, match=case.match, k=case.k, e=case.e, seed=case.seed, sample=case.sample, max_peak=case.max_peak)
    if case.prebuilt_index:
        rc, _ = orc.extract_ref(fq1, fq2, fa, out + ".first", **kw)
        assert rc == 0
    rc, st = orc.extract_ref(fq1, fq2, fa, out, **kw)
    assert rc == 0
    return fa, fq1, fq2, out, st


@pytest.mark.parametrize("case", fixtures.CASES, ids=lambda c: c.name)
def test_port_matches_reference_binary(case, manifest, workdir):
    gold = manifest[case.name]
    fa, fq1, fq2, out, st = run_port(case, workdir)
    try:
        # generator drift would show up here, not as a parity failure
        assert fixtures.sha256(fa) == gold["inputs"]["fasta"]
        assert fixtures.sha256(fq1) == gold["inputs"]["fq1"]
        assert fixtures.sha256(fq2) == gold["inputs"]["fq2"]
        idx = fixtures.index_path(fa, case.k, case.e)
        assert os.path.getsize(idx) == gold["index_bytes"]
        assert fixtures.sha256(idx) == gold["index_sha256"]                 # bit-exact index
        assert open(fa + ".genome.len.txt").read() == gold["len_text"]      # bit-exact genome.len.txt
        assert open(out).read() == gold["interval_text"]                    # bit-exact intervals
        assert (st[0] + st[1]) // 2 == gold["ref_pairs_s1"]
        assert st[4] == gold["ref_pairs_s3"]
        assert st[3] == gold["ref_raw_peaks"]
    finally:
        fixtures.clean_outputs(fa)


def test_manifest_covers_every_case(manifest):
    assert {c.name for c in fixtures.CASES} <= set(manifest)
