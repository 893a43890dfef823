"""TEST INFRASTRUCTURE ONLY — ctypes view of oracle/liblhgt_oracle.so (the scalar C restatement) and a
runner for oracle/_ref/extract_ref_z (the unmodified reference, zero-filling operator new[]).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblhgt_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "extract_ref")        # timing baseline
REF_BIN_Z = os.path.join(HERE, "_ref", "extract_ref_z")    # parity oracle (deterministic heap)

_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-s", "-C", HERE, "port"])
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl, u32, f32 = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_uint, C.c_float
        cs = C.c_char_p
        sig = {
            "orc_create": (vp, [i32, i32]), "orc_destroy": (None, [vp]),
            "orc_srand": (None, [vp, u32]), "orc_rand": (i32, [vp]),
            "orc_random_coder": (None, [vp]), "orc_set_coder": (None, [vp, vp]),
            "orc_coder": (vp, [vp]), "orc_load_coder_from_index": (i32, [vp, cs]),
            "orc_hash_seq": (i64, [vp, vp, i64, vp, vp]),
            "orc_index_build": (i32, [vp, cs, cs, cs]),
            "orc_sample_ratio": (dbl, [cs, dbl]),
            "orc_fill_random_array": (None, [vp, i64]),
            "orc_s1_count": (i64, [vp, cs, i64, dbl]),
            "orc_s2_peaks": (i64, [vp, cs, f32, f32, i64]),
            "orc_s3_pairs": (i64, [vp, cs, cs, dbl]),
            "orc_write_intervals": (i32, [vp, cs]),
            "orc_extract_ref": (i32, [cs, cs, cs, cs, dbl, dbl, i32, i64, i32, u32, dbl, vp]),
            "orc_count_table": (vp, [vp]), "orc_n_peaks": (i64, [vp]), "orc_peak_loci": (vp, [vp]),
            "orc_peak_filter": (vp, [vp]), "orc_peak_kmer": (vp, [vp]),
            "orc_raw_peak_positions": (i64, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _view(ptr: int, n: int, dtype) -> np.ndarray:
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class Oracle:
    """Stage-level handle (semantics of `extract_ref -t 1`)."""

    def __init__(self, k: int, e: int):
        self.k, self.e = k, e
        self.h = lib().orc_create(k, e)
        if not self.h:
            raise ValueError("bad k/e")

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    __del__ = close

    def srand(self, seed: int): lib().orc_srand(self.h, seed)
    def rand(self) -> int: return lib().orc_rand(self.h)
    def random_coder(self) -> np.ndarray:
        lib().orc_random_coder(self.h)
        return self.coder()
    def coder(self) -> np.ndarray:
        return _view(lib().orc_coder(self.h), 300, np.int16).copy()
    def set_coder(self, cc: np.ndarray):
        cc = np.ascontiguousarray(cc, dtype=np.int16)
        assert cc.size == 300
        lib().orc_set_coder(self.h, cc.ctypes.data)
    def load_coder(self, index_path: str) -> int:
        return lib().orc_load_coder_from_index(self.h, index_path.encode())

    def hash_seq(self, seq: bytes):
        s = np.frombuffer(seq, dtype=np.uint8)
        npos = max(0, len(s) - self.k + 1)
        out = np.zeros((npos, self.e), dtype=np.uint32)
        valid = np.zeros(npos, dtype=np.uint8)
        if npos:
            lib().orc_hash_seq(self.h, s.ctypes.data, len(s), out.ctypes.data, valid.ctypes.data)
        return out, valid

    def index_build(self, fasta, index_path, len_path) -> int:
        return lib().orc_index_build(self.h, fasta.encode(), index_path.encode(), len_path.encode())
    def fill_random(self, n: int): lib().orc_fill_random_array(self.h, n)
    def s1_count(self, fq, budget, ratio) -> int:
        return lib().orc_s1_count(self.h, fq.encode(), budget, ratio)
    def s2_peaks(self, index_path, hit, match, max_peak) -> int:
        return lib().orc_s2_peaks(self.h, index_path.encode(), hit, match, max_peak)
    def s3_pairs(self, fq1, fq2, ratio) -> int:
        return lib().orc_s3_pairs(self.h, fq1.encode(), fq2.encode(), ratio)
    def write_intervals(self, path) -> int:
        return lib().orc_write_intervals(self.h, path.encode())

    def count_table(self) -> np.ndarray:
        return _view(lib().orc_count_table(self.h), 1 << self.k, np.uint8)
    def n_peaks(self) -> int: return lib().orc_n_peaks(self.h)
    def raw_positions(self) -> int: return lib().orc_raw_peak_positions(self.h)
    def peak_loci(self) -> np.ndarray:
        return _view(lib().orc_peak_loci(self.h), 2 * self.n_peaks(), np.int32).reshape(-1, 2)
    def peak_filter(self) -> np.ndarray:
        return _view(lib().orc_peak_filter(self.h), self.n_peaks(), np.uint8)
    def peak_kmer(self) -> np.ndarray:
        return _view(lib().orc_peak_kmer(self.h), 1 << self.k, np.uint32)


def sample_ratio(fq1: str, sample_arg: float) -> float:
    return lib().orc_sample_ratio(fq1.encode(), sample_arg)


def extract_ref(fq1, fq2, fasta, interval, *, hit=0.1, match=0.08, k=32, max_peak=300000000, e=3,
                seed=1, sample=2000000000.0):
    """The port's main() at -t 1.  Returns (rc, stats[6])."""
    st = (C.c_long * 6)()
    rc = lib().orc_extract_ref(fq1.encode(), fq2.encode(), fasta.encode(), interval.encode(), hit, match,
                               k, max_peak, e, seed, sample, C.addressof(st))
    return rc, list(st)


def run_reference(fq1, fq2, fasta, interval, *, hit=0.1, match=0.08, threads=1, k=32,
                  max_peak=300000000, e=3, seed=1, sample=2000000000.0, zero_heap=True,
                  timeout: Optional[float] = None) -> str:
    """Runs the UNMODIFIED reference binary (same 12 positional args as scripts/pipeline.sh:35)."""
    exe = REF_BIN_Z if zero_heap else REF_BIN
    if not os.path.exists(exe):
        raise FileNotFoundError(exe + " (run `make -C oracle ref` where /root/reference exists)")
    argv = [exe, fq1, fq2, fasta, interval, repr(float(hit)), repr(float(match)), str(threads), str(k),
            str(max_peak), str(e), str(seed), repr(float(sample))]
    return subprocess.run(argv, check=True, capture_output=True, text=True, timeout=timeout).stdout
