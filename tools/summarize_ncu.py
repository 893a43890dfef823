#!/usr/bin/env python
"""Turns what a GPU session left in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
    python tools/summarize_ncu.py kernel   gpurun_out/s1_count_kernel.ncu-rep > profiles/rNN_s1_count_kernel.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_lookup_miss.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[gi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none launch list: {len(rows) - 1} launches, {tot:.3f} ms of kernels")
    print("(per-launch times are cold-cache and serialised; compare shares, not absolutes)\n")
    print("| kernel | launches | total ms | share | grid of first launch |\n|---|---:|---:|---:|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[2]} |")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(f"### {d.get('Kernel Name', '?')}  grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | unit | value |\n|---|---|---:|")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or ("issue_stalled" in h and h.endswith("_per_warp_active.pct") and float(v or 0) >= 3):
                print(f"| {h} | {u} | {v} |")
        print()


def traffic(*paths):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed the way bench.py names kernels."""
    import json
    import os
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = {}
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
            name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").strip()
            tot = sum(float(d[m].replace(",", "")) * scale.get(u[m], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            out[name] = int(tot)
            if name.startswith("s1_split") or name.startswith("s1_leaf"):
                out[name.split("<")[0]] = int(tot)
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    cmd = {"launches": launches, "kernel": kernel, "traffic": traffic}[sys.argv[1]]
    cmd(*sys.argv[2:])
