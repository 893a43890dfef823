"""Regenerates tests/golden/BED.json: the UNMODIFIED reference script scripts/get_bed_file.py run (python3, here) on
the golden interval / genome.len.txt texts of MANIFEST.json, plus a few hand-made edge cases.

Run where /root/reference exists:   python tests/golden/make_bed_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/scripts/get_bed_file.py"

EDGE = {
    # name: (interval text, genome.len.txt text)
    "clamp_and_drop": ("1\t1\t1\n1\t-487\t513\n2\t0\t49\n2\t1\t51\n2\t10\t59\n2\t10\t60\n",
                       "a\t1\t5000\t5000\nb\t2\t7000\t12000\n"),
    "skipped_contig_shifts_names": ("1\t1\t1\n2\t100\t900\n3\t100\t900\n",            # quirk Q2
                                    "a\t1\t5000\t5000\nc\t3\t7000\t12020\nd\t4\t800\t12820\n"),
    "later_line_wins": ("1\t100\t900\n", "a\t1\t5000\t5000\nz\t1\t7000\t12000\n"),
    "only_the_initial_state": ("1\t1\t1\n", "a\t1\t5000\t5000\n"),
}


def run(interval_text: str, len_text: str):
    with tempfile.TemporaryDirectory() as d:
        ref = os.path.join(d, "ref.fa")
        iv = os.path.join(d, "s.interval.txt")
        open(ref + ".genome.len.txt", "w").write(len_text)
        open(iv, "w").write(interval_text)
        p = subprocess.run([sys.executable, SCRIPT, ref, iv], capture_output=True, text=True)
        bed = open(iv + ".bed").read() if os.path.exists(iv + ".bed") else ""
        return {"interval_text": interval_text, "len_text": len_text, "returncode": p.returncode,
                "bed_text": bed if p.returncode == 0 else None, "stdout": p.stdout if p.returncode == 0 else None}


def main():
    manifest = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    out = {"_generated_by": "python3 /root/reference/scripts/get_bed_file.py (unmodified) on the texts stored here"}
    for name, rec in manifest.items():
        if name.startswith("_"):
            continue
        out[name] = run(rec["interval_text"], rec["len_text"])
    for name, (iv, ln) in EDGE.items():
        out["edge_" + name] = run(iv, ln)
    json.dump(out, open(os.path.join(HERE, "BED.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        if not k.startswith("_"):
            print(f"{k:34s} rc={v['returncode']} lines={len((v['bed_text'] or '').splitlines())} {(v['stdout'] or '').strip()}")


if __name__ == "__main__":
    main()
