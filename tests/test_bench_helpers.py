"""bench.py's host-side bookkeeping (no GPU): the strong-scaling split, the sampled fraction of E:1392-1398 and the linear model
that scales the reference binary's bounded sample to the full workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_split_range_tiles_the_sample():
    for n, parts in ((30_000_000, 8), (10, 3), (5, 8), (0, 2)):
        cuts = [bench.split_range(n, parts, i) for i in range(parts)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1


def test_sampled_fraction_follows_the_reference_rule():
    # ratio = sample / (2 * bases of fq1), capped at 1 (E:1258-1265, 1392-1398)
    assert bench.sampled_fraction(5_000_000) == 1.0
    assert abs(bench.sampled_fraction(30_000_000) - 2e9 / (2 * 150 * 30e6)) < 1e-15
    assert abs(bench.sampled_fraction(10_000_000) - 2 / 3) < 1e-12


def test_cpu_model_extrapolation_is_linear(tmp_path, monkeypatch):
    m = bench.CpuModel.__new__(bench.CpuModel)
    m.name, m.n, m.generated = "cfg4", 200_000, True
    m.full_pairs, m.full_bases = 30_000_000, 5_000_000_000
    m.frac = bench.sampled_fraction(m.full_pairs)
    m.small_bases, m.big_bases = 10_000_020, 100_000_020
    fixed, c_base, c_pair = 9.0, 4e-9, 2e-5
    t_tiny_small = fixed + c_base * m.small_bases
    t_tiny_big = fixed + c_base * m.big_bases
    t_pairs_small = t_tiny_small + c_pair * m.n * m.frac
    value, model = m.estimate(t_tiny_small, t_pairs_small, t_tiny_big)
    total = fixed + c_base * m.full_bases + c_pair * m.frac * m.full_pairs
    assert abs(model["full_workload_seconds"] - total) < 1e-6 * total
    assert abs(value - m.full_pairs / total) < 1e-6 * value
    assert set(bench.config_dict("cfg4")) >= {"workload", "pairs", "ref_bases", "sampled_fraction"}
