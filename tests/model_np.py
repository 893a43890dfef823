"""TEST INFRASTRUCTURE — numpy statement of the closed forms the CUDA kernels implement.

The kernels do not evaluate the reference's sequential loops; they evaluate the equivalent data-parallel
forms below.  CPU tests check these forms against the oracle (tests/test_closed_forms.py) so a GPU parity
failure can be attributed to a kernel, not to the algebra.
"""
from __future__ import annotations

import numpy as np

WINDOW, PEAK_W, DIFF_MIN, BUCKET = 500, 5, 2, 50


# ---------------------------------------------------------------- hash
def plane_bits(seq: np.ndarray):
    """ASCII -> (p0,p1,p2,valid) 0/1 arrays.  coder0 A,T=1; coder1 A,C=1; coder2 A,G=1 (E:1109-1154)."""
    up = seq & 0xDF
    a, c, g, t = (up == 65), (up == 67), (up == 71), (up == 84)
    valid = a | c | g | t
    return (a | t).astype(np.uint8), (a | c).astype(np.uint8), (a | g).astype(np.uint8), valid.astype(np.uint8)


def masks(cc: np.ndarray, k: int, e: int):
    """M[c][i]: bit (k-1-z) set iff cc[z*e+i] == c."""
    M = np.zeros((3, e), dtype=np.uint64)
    for z in range(k):
        for i in range(e):
            M[int(cc[z * e + i]), i] |= np.uint64(1) << np.uint64(k - 1 - z)
    return M


def _rev_k(x: np.ndarray, k: int) -> np.ndarray:
    out = np.zeros_like(x)
    for b in range(k):
        out |= ((x >> np.uint64(b)) & np.uint64(1)) << np.uint64(k - 1 - b)
    return out


def hash_seq(seq: np.ndarray, k: int, e: int, cc: np.ndarray):
    """XOR-select form used on the device:
       F = W2 ^ ((W0^W2)&M0) ^ ((W1^W2)&M1)
       R = (~rev(W2) ^ (~rev(W0^W2) & M0) ^ (rev(W1^W2) & M1)) & kmask ; h = min(F, R)."""
    n = len(seq)
    npos = n - k + 1
    if npos <= 0:
        return np.zeros((0, e), np.uint32), np.zeros(0, np.uint8)
    p0, p1, p2, v = plane_bits(seq)
    W = []
    for p in (p0, p1, p2, v):
        w = np.zeros(npos, dtype=np.uint64)
        for z in range(k):
            w |= p[z:z + npos].astype(np.uint64) << np.uint64(k - 1 - z)
        W.append(w)
    W0, W1, W2, WV = W
    kmask = np.uint64((1 << k) - 1)
    M = masks(cc, k, e)
    x0, x1 = W0 ^ W2, W1 ^ W2
    rx0, rx1, rw2 = _rev_k(x0, k), _rev_k(x1, k), _rev_k(W2, k)
    out = np.zeros((npos, e), dtype=np.uint32)
    valid = (WV == kmask)
    for i in range(e):
        F = W2 ^ (x0 & M[0, i]) ^ (x1 & M[1, i])
        R = (~rw2 ^ (~rx0 & M[0, i]) ^ (rx1 & M[1, i])) & kmask
        out[:, i] = np.where(valid, np.minimum(F, R), 0).astype(np.uint32)
    return out, valid.astype(np.uint8)


# ---------------------------------------------------------------- S2 per contig
def contig_flags(hit: np.ndarray, length: int, k: int, one_min: int, three_min: int) -> np.ndarray:
    """hit: (len-k+1, e) counts 0..3 (already 0 where the stored hash is 0).  Returns the ascending
    positions the reference feeds to add_peak (E:688-712)."""
    e = hit.shape[1]
    npos = hit.shape[0]
    full = (hit == 3).sum(axis=1)
    single = np.zeros(length, dtype=np.int64)
    trio = np.zeros(length, dtype=np.int64)
    single[:npos] = full > 0
    trio[:npos] = full == e
    cs = np.concatenate([[0], np.cumsum(single)])     # cs[x] = sum single[0..x-1]
    ct = np.concatenate([[0], np.cumsum(trio)])
    j = np.arange(length)
    lo = np.maximum(0, j - WINDOW + 1)
    one = cs[j + 1] - cs[lo]
    three = ct[j + 1] - ct[lo]
    good = (one >= one_min) & (three >= three_min)
    cg = np.concatenate([[0], np.cumsum(good)])
    a = np.maximum(0, j - 2 * WINDOW)
    b = np.minimum(length - 1, j + 2 * WINDOW)
    in_iv = (cg[b + 1] - cg[a] > 0) & (j >= 1)
    # D(x) = sum single[x-4..x];  C(j) = D(j-5) - D(j-k-5) - D(j);  diff_t(j) = C(j) + D(j-k-5-t)
    def D(x):
        x = np.asarray(x)
        xx = np.clip(x, 0, length - 1)
        return cs[xx + 1] - cs[np.maximum(0, xx - PEAK_W + 1)]
    peak = np.zeros(length, dtype=bool)
    ok_j = (j > 2 * k + 2 * PEAK_W)
    C = np.where(ok_j, D(j - 5) - D(j - k - 5) - D(j), 0)
    for t in range(k):
        # rule (i): own position
        d_own = np.where(ok_j, C + D(j - k - 5 - t), 0)
        peak |= ok_j & (d_own <= -DIFF_MIN)
        # rule (ii): q = j-k-t-5 flagged from a later j
        jj = j + k + 5 + t                      # j here plays q
        okq = (jj < length)
        jjc = np.minimum(jj, length - 1)
        d_far = C[jjc] + D(j)
        peak |= okq & ok_j[jjc] & (d_far >= DIFF_MIN)
    return np.nonzero(peak & in_iv)[0]


def register_peaks(flag_lists, k: int):
    """flag_lists: list over contigs (record ordinal starting at 1) of ascending flagged positions.
    Returns loci (n,2) and, per flagged position, its peak id (first flagged position of each
    (contig, pos//50) bucket opens a new id, E:288-301)."""
    loci, ids = [], []
    for ordinal, pos in enumerate(flag_lists, start=1):
        pos = np.asarray(pos)
        if len(pos) == 0:
            ids.append(np.zeros(0, np.int64))
            continue
        bucket = pos // BUCKET
        new = np.ones(len(pos), dtype=bool)
        new[1:] = bucket[1:] != bucket[:-1]
        base = len(loci)
        ids.append(base + np.cumsum(new) - 1)
        for p in pos[new]:
            loci.append((ordinal, int(p)))
    return np.array(loci, dtype=np.int32).reshape(-1, 2), ids
