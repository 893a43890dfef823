"""Regenerates tests/golden/MANIFEST.json by running the UNMODIFIED reference binary
(oracle/_ref/extract_ref_z = /root/reference/src/extract_ref_normal_peak.cpp + zero-filling operator
new[], built by oracle/Makefile) on every case in tests/fixtures.py at -t 1.

Run where /root/reference exists:   python tests/golden/make_golden.py
What is stored per case: sha256 of the three inputs (to detect generator drift), sha256 + size of the
index and of genome.len.txt, the genome.len.txt text, the interval file text, and the reference's own
"considered read pair" counters parsed from its stdout.
"""
import json
import os
import re
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import fixtures  # noqa: E402
from oracle import orc  # noqa: E402


def run_case(case, work):
    fa, fq1, fq2 = fixtures.materialize(case.data, work)
    fixtures.clean_outputs(fa)
    out = os.path.join(work, case.name + ".interval.txt")
    kw = dict(hit=case.hit, match=case.match, k=case.k, e=case.e, seed=case.seed, sample=case.sample, max_peak=case.max_peak)
    if case.prebuilt_index:
        orc.run_reference(fq1, fq2, fa, out + ".first", **kw)
    t = time.time()
    log = orc.run_reference(fq1, fq2, fa, out, **kw)
    dt = time.time() - t
    idx = fixtures.index_path(fa, case.k, case.e)
    lenf = fa + ".genome.len.txt"
    m1 = re.search(r"considered read pair num in kmer counting:(\d+)", log)
    m3 = re.search(r"considered read pair num in finding candidate HGT breakpoint:(\d+)", log)
    raw = re.findall(r"No\. of raw BKPs: (\d+)", log)
    rec = {
        "inputs": {"fasta": fixtures.sha256(fa), "fq1": fixtures.sha256(fq1), "fq2": fixtures.sha256(fq2)},
        "index_sha256": fixtures.sha256(idx), "index_bytes": os.path.getsize(idx),
        "len_sha256": fixtures.sha256(lenf), "len_text": open(lenf).read(),
        "interval_text": open(out).read(),
        "ref_pairs_s1": int(m1.group(1)) if m1 else None, "ref_pairs_s3": int(m3.group(1)) if m3 else None,
        "ref_raw_peaks": int(raw[-1]) if raw else None,
        "args": {"k": case.k, "e": case.e, "seed": case.seed, "hit": case.hit, "match": case.match,
                 "sample": case.sample, "max_peak": case.max_peak, "prebuilt_index": case.prebuilt_index},
        "note": case.note, "reference_seconds": round(dt, 1),
    }
    fixtures.clean_outputs(fa)
    return rec


def main():
    only = set(sys.argv[1:])
    path = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(path)) if os.path.exists(path) and only else {}
    with tempfile.TemporaryDirectory() as work:
        for case in fixtures.CASES:
            if only and case.name not in only:
                continue
            rec = run_case(case, work)
            manifest[case.name] = rec
            print(f"{case.name:14s} {rec['reference_seconds']:6.1f}s  raw_peaks={rec['ref_raw_peaks']}  "
                  f"intervals={len(rec['interval_text'].splitlines())}  index={rec['index_bytes']}B", flush=True)
    manifest["_generated_by"] = ("oracle/_ref/extract_ref_z (unmodified src/extract_ref_normal_peak.cpp, g++ -O2, "
                                 "zero-filling operator new[] shim), -t 1")
    json.dump(manifest, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
