"""CPU: the algebra the CUDA kernels rely on, stated in numpy / plain Python and held against the oracle, so that a GPU parity
failure can be attributed to a kernel and not to the restructuring:

  * the closed forms of slide_window / add_peak (tests/model_np.py) reproduce the oracle's peak loci;
  * S2's marking: every flagged position lies in a tile within two tiles of a HOT tile (one holding a position whose
    500-wide sum of `trio` reaches three_min) -- so the window passes may skip every other tile;
  * S2's short-circuit gather: trio = AND of the e hits is decided by the first non-saturated hash;
  * S3's pre-check (s3_may_split): whenever check_split marks a peak, two contigs have >= 6 candidate positions, also
    when the count is taken per hashed slot with the most common contig counted exactly.
"""
import os
import random

import numpy as np
import pytest

import fixtures
import model_np
from oracle import orc

TILE = 1024


def _oracle_s2(case, workdir):
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    o = orc.Oracle(case.k, case.e)
    o.srand(case.seed); o.random_coder()
    idx, lenp = os.path.join(workdir, case.name + ".cf.index.dat"), os.path.join(workdir, case.name + ".cf.len.txt")
    assert o.index_build(fa, idx, lenp) == 0
    ratio = orc.sample_ratio(fq1, case.sample)
    if ratio < 100:
        o.fill_random(200000)
    size1 = os.path.getsize(fq1)
    o.s1_count(fq1, size1, ratio); o.s1_count(fq2, size1, ratio)
    o.s2_peaks(idx, case.hit, case.match, case.max_peak)
    return o, idx


def _contigs_of_index(idx, k, e):
    words = np.fromfile(idx, dtype=np.uint32)
    at, out = 300, []
    while at < len(words):
        ln = int(words[at])
        n = (ln - k + 1) * e
        out.append((ln, words[at + 1:at + 1 + n].reshape(-1, e)))
        at += 1 + n
    return out


@pytest.mark.parametrize("name", ["base_k24", "base_k20", "noisy"])
def test_closed_forms_and_marking(name, workdir):
    case = fixtures.BY_NAME[name]
    o, idx = _oracle_s2(case, workdir)
    table = o.count_table()
    one_min = int(np.float32(500) * np.float32(case.hit)); three_min = int(np.float32(500) * np.float32(case.match))
    flag_lists, n_flagged = [], 0
    for ln, hashes in _contigs_of_index(idx, case.k, case.e):
        hit = np.where(hashes != 0, table[hashes], 0).astype(np.int64)
        flags = model_np.contig_flags(hit, ln, case.k, one_min, three_min)
        flag_lists.append(flags)
        n_flagged += len(flags)
        # --- short-circuit AND: the first non-saturated hash decides trio; hash 0's hit is a lower bound of single
        sat = hit == 3
        trio_sc = sat.all(axis=1)
        alive = np.ones(len(sat), dtype=bool)                     # the kernel's loop: hash i+1 is only looked at where 0..i were saturated
        for i in range(case.e):
            alive &= sat[:, i]
        assert np.array_equal(alive, trio_sc)
        assert not np.any(sat[:, 0] & ~sat.any(axis=1))           # hash 0's hit never exceeds `single`
        # --- marking: hot tiles from the window sums of trio; flagged positions only within two tiles of one
        trio = np.zeros(ln, dtype=np.int64); trio[:len(sat)] = trio_sc
        ct = np.concatenate([[0], np.cumsum(trio)])
        j = np.arange(ln)
        three = ct[j + 1] - ct[np.maximum(0, j - 499)]
        ntiles = (ln + TILE - 1) // TILE
        hot = np.zeros(ntiles, dtype=bool)
        np.logical_or.at(hot, j // TILE, three >= three_min)
        need = np.zeros(ntiles, dtype=bool)
        for d in range(-2, 3):
            src = np.arange(ntiles) + d
            ok = (src >= 0) & (src < ntiles)
            need[ok] |= hot[src[ok]]
        assert need[flags // TILE].all(), "a flagged position outside the marked tiles"
    loci, _ = model_np.register_peaks(flag_lists, case.k)
    assert n_flagged == o.raw_positions()
    assert np.array_equal(loci, o.peak_loci())
    o.close()


def _vote(cands, e):
    """judge_base + check_split (E:118-202) on a list of positions, each a list of e (peak id, contig) candidates (peak 0 = none).
    Returns the set of marked peak ids."""
    votes, first = {}, {}
    for pos in cands:
        sel, sel_votes, seen = None, 0, False
        for pk, c in pos:
            if not pk:
                continue
            if c in votes:
                if votes[c] >= sel_votes:
                    sel, sel_votes, seen = (pk, c), votes[c], True
            elif sel is None:
                sel, sel_votes, seen = (pk, c), 0, False
        if sel is None:
            continue
        pk, c = sel
        if c in votes:
            votes[c] += 1
        else:
            votes[c] = 1
            first[c] = pk
    strong = sorted((v for v in votes.values() if v >= 6), reverse=True)
    if len(strong) < 2:
        return set()
    largest, second = strong[0], strong[1]
    return {first[c] for c, v in votes.items() if v >= 6 and v in (largest, second)}


def _may_split(cands, slots=1024):
    """The bound s3_may_split evaluates: [c* has >= 6 candidates] + sum over hashed slots of floor(count / 6) >= 2."""
    flat = [c for pos in cands for pk, c in pos if pk]
    if not flat:
        return False
    n = len([1 for pos in cands for _ in pos])
    entries = [c if pk else 0 for pos in cands for pk, c in pos]
    probes = [entries[(lane * n) >> 5] for lane in range(32)]
    probes = [p for p in probes if p]
    if not probes:
        return False
    cstar = max(sorted(set(probes)), key=probes.count)
    mine = sum(1 for c in flat if c == cstar)
    tally = {}
    for c in flat:
        if c != cstar:
            s = ((c * 2654435761) & 0xffffffff) >> (32 - slots.bit_length() + 1)
            tally[s] = tally.get(s, 0) + 1
    return (1 if mine >= 6 else 0) + sum(v // 6 for v in tally.values()) >= 2


def test_vote_precheck_is_a_necessary_condition():
    rng = random.Random(7)
    marked_cases = 0
    for trial in range(3000):
        e = rng.choice([1, 2, 3, 4])
        n_contigs = rng.choice([3, 8, 40, 2000])
        n_pos = rng.choice([6, 12, 40, 238])
        main = [rng.randrange(1, n_contigs + 1) for _ in range(rng.choice([1, 2, 3]))]
        cands = []
        for _ in range(n_pos):
            pos = []
            for _i in range(e):
                u = rng.random()
                if u < 0.25:
                    pos.append((0, 0))
                elif u < 0.75:
                    c = rng.choice(main)
                    pos.append((1000 * c + rng.randrange(1, 50), c))
                else:
                    c = rng.randrange(1, n_contigs + 1)
                    pos.append((1000 * c + rng.randrange(1, 50), c))
            if any(pk for pk, _ in pos):
                cands.append(pos)
        marked = _vote(cands, e)
        if marked:
            marked_cases += 1
            assert _may_split(cands), "the pre-check would have skipped a pair whose vote marks a peak"
            assert _may_split(cands, slots=8), "even with heavy slot sharing the bound must hold"
    assert marked_cases > 300
