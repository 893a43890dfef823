"""ctypes binding of liblhgt.so — the host-side mirror of include/lhgt.h.

Python is only the caller here: every byte of the screened path is computed by the CUDA kernels behind
the C ABI.  There is no CPU fallback; if the library or a GPU is missing the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from . import build as _build

CODER_SLOTS = 300

_ERR = {-1: "LHGT_E_ARG", -2: "LHGT_E_IO", -3: "LHGT_E_CUDA", -4: "LHGT_E_FORMAT", -5: "LHGT_E_TOO_MANY_PEAKS",
        -6: "LHGT_E_READ_TOO_LONG", -7: "LHGT_E_UNPAIRED", -8: "LHGT_E_STATE", -9: "LHGT_E_NOMEM"}


class LhgtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR.get(code, code)}: {msg}")
        self.code = code


class Args(C.Structure):
    _fields_ = [("fq1", C.c_char_p), ("fq2", C.c_char_p), ("fasta", C.c_char_p), ("interval", C.c_char_p),
                ("hit_ratio", C.c_double), ("match_ratio", C.c_double), ("threads", C.c_int), ("k", C.c_int),
                ("max_peak", C.c_long), ("e", C.c_int), ("seed", C.c_uint), ("sample", C.c_double),
                ("device", C.c_int), ("quiet", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("reads_s1", C.c_long * 2), ("flagged_positions", C.c_long), ("peaks", C.c_long),
                ("pairs_s3", C.c_long), ("kept_peaks", C.c_long), ("index_built", C.c_int),
                ("ratio_percent", C.c_double), ("seconds", C.c_double * 8)]


# every symbol include/lhgt.h declares: name -> (restype, argtypes)
_vp, _i, _l, _d, _u, _f, _s = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_uint, C.c_float, C.c_char_p
_u64, _sz = C.c_uint64, C.c_size_t
SYMBOLS = {
    "lhgt_abi_version": (_i, []),
    "lhgt_last_error": (_s, []),
    "lhgt_rand_stream": (_i, [_u, _l, _l, _vp]),
    "lhgt_random_coder": (_i, [_u, _i, _i, _vp]),
    "lhgt_coder_to_header": (_i, [_vp, _vp]),
    "lhgt_header_to_coder": (_i, [_vp, _vp]),
    "lhgt_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "lhgt_destroy": (None, [_vp]),
    "lhgt_set_coder": (_i, [_vp, _vp]),
    "lhgt_get_coder": (_i, [_vp, _vp]),
    "lhgt_set_stream": (_i, [_vp, C.c_size_t]),
    "lhgt_sync": (_i, [_vp]),
    "lhgt_hash_seq": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "lhgt_index_build": (_i, [_vp, _vp, _sz]),
    "lhgt_index_build_device": (_i, [_vp, _vp, _sz]),
    "lhgt_fasta_prefetch": (_i, [_vp, _vp, _sz]),
    "lhgt_index_bytes": (_u64, [_vp]),
    "lhgt_index_bases": (_u64, [_vp]),
    "lhgt_index_contigs": (_l, [_vp]),
    "lhgt_index_download": (_i, [_vp, _vp, _u64]),
    "lhgt_index_len_text": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "lhgt_index_record": (_l, [_vp, _l, _vp, _u64]),
    "lhgt_set_image_block": (_i, [_vp, _i, _i]),
    "lhgt_index_block": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_l), C.POINTER(_l)]),
    "lhgt_index_write_block": (_i, [_vp, _s]),
    "lhgt_index_upload": (_i, [_vp, _vp, _u64]),
    "lhgt_index_build_file": (_i, [_vp, _s, _s, _s]),
    "lhgt_index_load_file": (_i, [_vp, _s]),
    "lhgt_reads_upload": (_i, [_vp, _i, _vp, _u64]),
    "lhgt_reads_prefetch": (_i, [_vp, _i, _vp, _u64]),
    "lhgt_reads_upload_file": (_i, [_vp, _i, _s]),
    "lhgt_reads_prefetch_next": (_i, [_vp, _i, _vp, _u64]),
    "lhgt_index_prefetch": (_i, [_vp, _vp, _u64]),
    "lhgt_reads_attach_device": (_i, [_vp, _i, _vp, _u64]),
    "lhgt_reads_records": (_l, [_vp, _i]),
    "lhgt_reads_seq_bases": (_u64, [_vp, _i]),
    "lhgt_reads_bytes": (_u64, [_vp, _i]),
    "lhgt_sample_ratio": (_d, [_vp, _d]),
    "lhgt_set_sampling": (_i, [_vp, _d, _u, _l]),
    "lhgt_set_ordinal_base": (_i, [_vp, _u64]),
    "lhgt_s1_count": (_l, [_vp, _i, _u64]),
    "lhgt_set_s1_mode": (_i, [_vp, _i]),
    "lhgt_s2_peaks": (_l, [_vp, _f, _f, _l]),
    "lhgt_s3_pairs": (_l, [_vp, _l, _l]),
    "lhgt_intervals": (_i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    "lhgt_reset": (_i, [_vp]),
    "lhgt_count_table_copy": (_i, [_vp, _vp]),
    "lhgt_peaks_copy": (_l, [_vp, _vp, _vp, _l]),
    "lhgt_flagged_positions": (_l, [_vp]),
    "lhgt_peak_kmer_copy": (_i, [_vp, _vp]),
    "lhgt_s2_tiles": (_l, [_vp]),
    "lhgt_s2_gather": (_i, [_vp, _l, _l]),
    "lhgt_s2_finish": (_i, [_vp, _f, _f, _l, C.POINTER(_l)]),
    "lhgt_s2_need_range": (_i, [_vp, _i, _i, C.POINTER(_l), C.POINTER(_l)]),
    "lhgt_s2_windows": (_i, [_vp, _f, _f, _l, _l]),
    "lhgt_s2_flagged_in_range": (_l, [_vp]),
    "lhgt_s2_ids": (_i, [_vp, _l, _l, C.POINTER(_l)]),
    "lhgt_s2_register": (_i, [_vp, _l, _l]),
    "lhgt_s2_dense": (_i, [_vp]),
    "lhgt_dev_tile_new": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_dev_flagged": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_dev_peak_table": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_dev_loci": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_s2_mark": (_i, [_vp, _f]),
    "lhgt_s2_complete": (_i, [_vp, _l, _l]),
    "lhgt_s2_needed_tiles": (_l, [_vp]),
    "lhgt_dev_count_table": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_dev_hit_bits": (_vp, [_vp, _i, C.POINTER(_u64)]),
    "lhgt_dev_peak_filter": (_vp, [_vp, C.POINTER(_u64)]),
    "lhgt_count_merge": (_i, [_vp, _vp, _u64, _u64]),
    "lhgt_count_table_histogram": (_i, [_vp, _vp]),
    "lhgt_count_table_ipc": (_i, [_vp, _vp]),
    "lhgt_peers_open": (_i, [_vp, _i, _i, _vp]),
    "lhgt_count_exchange_p2p": (_i, [_vp]),
    "lhgt_set_deferred": (_i, [_vp, _i]),
    "lhgt_deferred_counts": (_i, [_vp, _vp]),
    "lhgt_stage_ms": (_i, [_vp, _vp]),
    "lhgt_stage_ms_ex": (_i, [_vp, _vp, _i]),
    "lhgt_launch_count": (_l, [_vp]),
    "lhgt_extract_ref": (_i, [C.POINTER(Args), C.POINTER(Stats)]),
    "lhgt_main": (_i, [_i, C.POINTER(_s)]),
    "lhgt_bed_text": (_i, [_s, _sz, _s, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(_l)]),
    "lhgt_regions_fasta": (_i, [_vp, _sz, _s, _sz, _vp, _sz, C.POINTER(_sz)]),
    "lhgt_extract_regions_files": (_i, [_s, _s, _s, C.POINTER(_l)]),
}

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Loads liblhgt.so and types every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is None:
        path = os.environ.get("LHGT_LIB") or _build.LIB          # LHGT_LIB: a variant build (tools/sweep.sh)
        if not os.path.exists(path):
            if not build_if_missing or path != _build.LIB:
                raise FileNotFoundError(path + " is not built (python -m localhgt_b200.build)")
            _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc: int) -> int:
    if rc < 0:
        raise LhgtError(rc, load().lhgt_last_error().decode(errors="replace"))
    return rc


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def rand_stream(seed: int, skip: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    _check(load().lhgt_rand_stream(seed, skip, n, _ptr(out)))
    return out


def random_coder(seed: int, k: int, e: int) -> Tuple[np.ndarray, int]:
    cc = np.zeros(CODER_SLOTS, dtype=np.int16)
    draws = _check(load().lhgt_random_coder(seed, k, e, _ptr(cc)))
    return cc, draws


def coder_to_header(cc: np.ndarray) -> np.ndarray:
    cc = np.ascontiguousarray(cc, dtype=np.int16)
    w = np.zeros(CODER_SLOTS, dtype=np.uint32)
    _check(load().lhgt_coder_to_header(_ptr(cc), _ptr(w)))
    return w


def header_to_coder(words: np.ndarray) -> np.ndarray:
    words = np.ascontiguousarray(words, dtype=np.uint32)
    cc = np.zeros(CODER_SLOTS, dtype=np.int16)
    _check(load().lhgt_header_to_coder(_ptr(words), _ptr(cc)))
    return cc


class Screen:
    """One GPU context: index resident in HBM, FASTQ images resident in HBM, stages S1..OUT."""

    def __init__(self, k: int = 32, e: int = 3, device: int = 0):
        self.k, self.e, self.device = k, e, device
        self._h = _vp()
        self._L = load()
        _check(self._L.lhgt_create(C.byref(self._h), device, k, e))
        self._keep = []

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.lhgt_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- coder / hashing
    def set_coder(self, cc: np.ndarray) -> None:
        cc = np.ascontiguousarray(cc, dtype=np.int16)
        assert cc.size == CODER_SLOTS
        _check(self._L.lhgt_set_coder(self._h, _ptr(cc)))

    def coder(self) -> np.ndarray:
        cc = np.zeros(CODER_SLOTS, dtype=np.int16)
        _check(self._L.lhgt_get_coder(self._h, _ptr(cc)))
        return cc

    def set_stream(self, cuda_stream: int) -> None:
        _check(self._L.lhgt_set_stream(self._h, cuda_stream))

    def sync(self) -> None:
        _check(self._L.lhgt_sync(self._h))

    def hash_seq(self, seq: bytes) -> Tuple[np.ndarray, np.ndarray]:
        s = np.frombuffer(seq, dtype=np.uint8)
        npos = max(0, len(s) - self.k + 1)
        out = np.zeros((npos, self.e), dtype=np.uint32)
        valid = np.zeros(npos, dtype=np.uint8)
        _check(self._L.lhgt_hash_seq(self._h, _ptr(s) if len(s) else None, len(s), _ptr(out), _ptr(valid)))
        return out, valid

    # ---- index
    def index_build(self, fasta: bytes) -> None:
        buf = np.frombuffer(fasta, dtype=np.uint8)
        _check(self._L.lhgt_index_build(self._h, _ptr(buf) if len(buf) else None, len(buf)))

    def index_build_ptr(self, host_ptr: int, n: int) -> None:
        _check(self._L.lhgt_index_build(self._h, host_ptr, n))

    def index_build_device(self, dev_ptr: int, n: int) -> None:
        """FASTA text already resident on this device (16-byte aligned)."""
        _check(self._L.lhgt_index_build_device(self._h, dev_ptr, n))

    def fasta_prefetch_ptr(self, host_ptr: int, n: int) -> None:
        _check(self._L.lhgt_fasta_prefetch(self._h, host_ptr, n))

    def index_bytes(self) -> int: return int(self._L.lhgt_index_bytes(self._h))
    def index_bases(self) -> int: return int(self._L.lhgt_index_bases(self._h))
    def index_contigs(self) -> int: return int(self._L.lhgt_index_contigs(self._h))

    def set_image_block(self, part: int, parts: int) -> None:
        """Keep only block `part` of `parts` of the image resident from the next index build on (images larger than one GPU)."""
        _check(self._L.lhgt_set_image_block(self._h, part, parts))

    def index_block(self):
        """(byte offset in the file, bytes, first tile, end tile) of the resident part of the image."""
        off, n, t0, t1 = _u64(0), _u64(0), _l(0), _l(0)
        _check(self._L.lhgt_index_block(self._h, C.byref(off), C.byref(n), C.byref(t0), C.byref(t1)))
        return int(off.value), int(n.value), int(t0.value), int(t1.value)

    def index_write_block(self, path: str) -> None:
        _check(self._L.lhgt_index_write_block(self._h, path.encode()))

    def index_download(self) -> np.ndarray:
        out = np.zeros(self.index_block()[1], dtype=np.uint8)
        _check(self._L.lhgt_index_download(self._h, _ptr(out), out.size))
        return out

    def index_download_ptr(self, host_ptr: int, cap: int) -> None:
        """Into caller-owned host memory (pinned memory makes the copy run at PCIe speed)."""
        _check(self._L.lhgt_index_download(self._h, host_ptr, cap))

    def index_len_text(self) -> bytes:
        n = _sz(0)
        _check(self._L.lhgt_index_len_text(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        _check(self._L.lhgt_index_len_text(self._h, buf, n.value, C.byref(n)))
        return buf.raw[: n.value]

    def index_record(self, record: int) -> np.ndarray:
        """[len, hashes...] of one indexed contig, from the resident image."""
        n = _check(self._L.lhgt_index_record(self._h, record, None, 0))
        out = np.zeros(n, dtype=np.uint32)
        _check(self._L.lhgt_index_record(self._h, record, _ptr(out), n))
        return out

    def index_upload(self, image) -> None:
        buf = np.frombuffer(image, dtype=np.uint8)
        _check(self._L.lhgt_index_upload(self._h, _ptr(buf), len(buf)))

    def index_upload_ptr(self, host_ptr: int, n: int) -> None:
        _check(self._L.lhgt_index_upload(self._h, host_ptr, n))

    def index_build_file(self, fasta: str, index_path: str, len_path: str) -> None:
        _check(self._L.lhgt_index_build_file(self._h, fasta.encode(), index_path.encode(), len_path.encode()))

    def index_load_file(self, index_path: str) -> None:
        _check(self._L.lhgt_index_load_file(self._h, index_path.encode()))

    # ---- reads
    def reads_upload(self, mate: int, fq) -> None:
        """fq: bytes / numpy uint8 / anything exposing a host pointer via numpy."""
        buf = fq if isinstance(fq, np.ndarray) else np.frombuffer(fq, dtype=np.uint8)
        _check(self._L.lhgt_reads_upload(self._h, mate, _ptr(buf) if len(buf) else None, len(buf)))

    def reads_prefetch_ptr(self, mate: int, host_ptr: int, n: int) -> None:
        """Starts the H2D copy on the copy stream; reads_upload_ptr(mate, same ptr, same n) adopts it later."""
        _check(self._L.lhgt_reads_prefetch(self._h, mate, host_ptr, n))

    def reads_prefetch_next_ptr(self, mate: int, host_ptr: int, n: int) -> None:
        """The NEXT sample's image starts its H2D copy into the alternate buffer; its reads_upload_ptr adopts it."""
        _check(self._L.lhgt_reads_prefetch_next(self._h, mate, host_ptr, n))

    def index_prefetch_ptr(self, host_ptr: int, n: int) -> None:
        _check(self._L.lhgt_index_prefetch(self._h, host_ptr, n))

    def reads_upload_ptr(self, mate: int, host_ptr: int, n: int) -> None:
        _check(self._L.lhgt_reads_upload(self._h, mate, host_ptr, n))

    def reads_upload_file(self, mate: int, path: str) -> None:
        _check(self._L.lhgt_reads_upload_file(self._h, mate, path.encode()))

    def reads_attach_device(self, mate: int, dev_ptr: int, n: int) -> None:
        _check(self._L.lhgt_reads_attach_device(self._h, mate, dev_ptr, n))

    def reads_records(self, mate: int) -> int: return int(self._L.lhgt_reads_records(self._h, mate))
    def reads_seq_bases(self, mate: int) -> int: return int(self._L.lhgt_reads_seq_bases(self._h, mate))
    def reads_bytes(self, mate: int) -> int: return int(self._L.lhgt_reads_bytes(self._h, mate))

    def sample_ratio(self, sample_arg: float) -> float:
        r = self._L.lhgt_sample_ratio(self._h, sample_arg)
        if r < 0 and sample_arg > 1:
            raise LhgtError(-8, self._L.lhgt_last_error().decode())
        return r

    def set_sampling(self, ratio_percent: float, seed: int = 1, rand_skip: int = 0) -> None:
        _check(self._L.lhgt_set_sampling(self._h, ratio_percent, seed, rand_skip))

    def set_ordinal_base(self, base: int) -> None:
        _check(self._L.lhgt_set_ordinal_base(self._h, base))

    # ---- stages
    def s1_count(self, mate: int, byte_budget: int) -> int:
        return _check(self._L.lhgt_s1_count(self._h, mate, byte_budget))

    def set_s1_mode(self, mode: int) -> None:
        """0 auto, 1 direct probes, 2 hash streams (identical counts; see include/lhgt.h)."""
        _check(self._L.lhgt_set_s1_mode(self._h, mode))

    def s2_peaks(self, hit_ratio: float = 0.1, match_ratio: float = 0.08, max_peak: int = 300000000) -> int:
        return _check(self._L.lhgt_s2_peaks(self._h, hit_ratio, match_ratio, max_peak))

    def s2_tiles(self) -> int: return int(self._L.lhgt_s2_tiles(self._h))

    def s2_gather(self, tile_begin: int = 0, tile_end: int = -1) -> None:
        _check(self._L.lhgt_s2_gather(self._h, tile_begin, tile_end))

    def s2_mark(self, match_ratio: float = 0.08) -> None:
        _check(self._L.lhgt_s2_mark(self._h, match_ratio))

    def s2_complete(self, tile_begin: int = 0, tile_end: int = -1) -> None:
        _check(self._L.lhgt_s2_complete(self._h, tile_begin, tile_end))

    def s2_needed_tiles(self) -> int: return int(self._L.lhgt_s2_needed_tiles(self._h))

    def s2_need_range(self, part: int, parts: int):
        lo, hi = _l(0), _l(0)
        _check(self._L.lhgt_s2_need_range(self._h, part, parts, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def s2_windows(self, hit_ratio: float, match_ratio: float, tile_begin: int = 0, tile_end: int = -1) -> None:
        _check(self._L.lhgt_s2_windows(self._h, hit_ratio, match_ratio, tile_begin, tile_end))

    def s2_flagged_in_range(self) -> int:
        return _check(self._L.lhgt_s2_flagged_in_range(self._h))

    def s2_ids(self, max_peak: int = 300000000, flagged_total: int = -1) -> int:
        n = _l(0)
        _check(self._L.lhgt_s2_ids(self._h, max_peak, flagged_total, C.byref(n)))
        return n.value

    def s2_register(self, tile_begin: int = 0, tile_end: int = -1) -> None:
        _check(self._L.lhgt_s2_register(self._h, tile_begin, tile_end))

    def s2_dense(self) -> bool: return bool(self._L.lhgt_s2_dense(self._h))

    def _dev(self, fn):
        n = _u64(0)
        p = fn(self._h, C.byref(n))
        return int(p or 0), int(n.value)

    def dev_tile_new(self): return self._dev(self._L.lhgt_dev_tile_new)
    def dev_flagged(self): return self._dev(self._L.lhgt_dev_flagged)
    def dev_peak_table(self): return self._dev(self._L.lhgt_dev_peak_table)
    def dev_loci(self): return self._dev(self._L.lhgt_dev_loci)

    def s2_finish(self, hit_ratio: float = 0.1, match_ratio: float = 0.08, max_peak: int = 300000000) -> int:
        n = _l(0)
        _check(self._L.lhgt_s2_finish(self._h, hit_ratio, match_ratio, max_peak, C.byref(n)))
        return n.value

    def s3_pairs(self, first: int = 0, count: int = -1) -> int:
        return _check(self._L.lhgt_s3_pairs(self._h, first, count))

    def intervals(self) -> bytes:
        n = _sz(0)
        _check(self._L.lhgt_intervals(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(max(1, n.value))
        _check(self._L.lhgt_intervals(self._h, buf, n.value, C.byref(n)))
        return buf.raw[: n.value]

    def reset(self) -> None:
        _check(self._L.lhgt_reset(self._h))

    # ---- state
    def count_table(self) -> np.ndarray:
        out = np.zeros(1 << self.k, dtype=np.uint8)
        _check(self._L.lhgt_count_table_copy(self._h, _ptr(out)))
        return out

    def peaks(self) -> Tuple[np.ndarray, np.ndarray]:
        n = _check(self._L.lhgt_peaks_copy(self._h, None, None, 0))
        loci = np.zeros((n, 2), dtype=np.int32)
        filt = np.zeros(n, dtype=np.uint8)
        if n:
            _check(self._L.lhgt_peaks_copy(self._h, _ptr(loci), _ptr(filt), n))
        return loci, filt

    def flagged_positions(self) -> int: return int(self._L.lhgt_flagged_positions(self._h))

    def peak_kmer(self) -> np.ndarray:
        out = np.zeros(1 << self.k, dtype=np.uint32)
        _check(self._L.lhgt_peak_kmer_copy(self._h, _ptr(out)))
        return out

    def dev_count_table(self) -> Tuple[int, int]:
        n = _u64(0)
        p = self._L.lhgt_dev_count_table(self._h, C.byref(n))
        return int(p or 0), int(n.value)

    def dev_hit_bits(self, which: int) -> Tuple[int, int]:
        n = _u64(0)
        p = self._L.lhgt_dev_hit_bits(self._h, which, C.byref(n))
        return int(p or 0), int(n.value)

    def dev_peak_filter(self) -> Tuple[int, int]:
        n = _u64(0)
        p = self._L.lhgt_dev_peak_filter(self._h, C.byref(n))
        return int(p or 0), int(n.value)

    def count_merge(self, dev_other: int, nbytes: int, word_offset: int = 0) -> None:
        _check(self._L.lhgt_count_merge(self._h, dev_other, nbytes, word_offset))

    def count_table_histogram(self) -> np.ndarray:
        """[v] = number of counters holding v (0..3); empty rate = [0] / 2^k (count_diff_kmer.cpp:26-50)."""
        out = np.zeros(4, dtype=np.uint64)
        _check(self._L.lhgt_count_table_histogram(self._h, _ptr(out)))
        return out

    def count_table_ipc(self) -> bytes:
        buf = C.create_string_buffer(64)
        _check(self._L.lhgt_count_table_ipc(self._h, buf))
        return buf.raw

    def peers_open(self, rank: int, world: int, handles: bytes) -> None:
        assert len(handles) == 64 * world
        _check(self._L.lhgt_peers_open(self._h, rank, world, handles))

    def count_exchange_p2p(self) -> None:
        _check(self._L.lhgt_count_exchange_p2p(self._h))

    def set_deferred(self, on: bool) -> None:
        _check(self._L.lhgt_set_deferred(self._h, int(on)))

    def deferred_counts(self):
        out = np.zeros(3, dtype=np.int64)
        _check(self._L.lhgt_deferred_counts(self._h, _ptr(out)))
        return int(out[0]), int(out[1]), int(out[2])

    def stage_ms(self) -> np.ndarray:
        """Device ms per stage since the last call: [0] FASTQ record scan [1] S1 [2] S2 gather [3] S2 finish [4] S3
        [5] IB [6] S1 hash-stream kernel [7] S1 stream-split kernel [8] S1 leaf-apply kernel [9] peer-memory count exchange
        [10] S2 peak registration [11] S3 vote."""
        ms = np.zeros(12, dtype=np.float32)
        _check(self._L.lhgt_stage_ms_ex(self._h, _ptr(ms), 12))
        return ms

    def launch_count(self) -> int: return int(self._L.lhgt_launch_count(self._h))


def extract_ref(fq1: str, fq2: str, fasta: str, interval: str, *, hit_ratio: float = 0.1, match_ratio: float = 0.08,
                threads: int = 1, k: int = 32, max_peak: int = 300000000, e: int = 3, seed: int = 1,
                sample: float = 2000000000.0, device: int = 0, quiet: bool = True) -> Stats:
    """The whole program (reference main(), E:1342-1519) through the C ABI."""
    a = Args(fq1.encode(), fq2.encode(), fasta.encode(), interval.encode(), float(np.float32(hit_ratio)),
             float(np.float32(match_ratio)), threads, k, max_peak, e, seed, sample, device, int(quiet))
    st = Stats()
    _check(load().lhgt_extract_ref(C.byref(a), C.byref(st)))
    return st


# ---------------------------------------------------------------------------------------------- post-screen glue
def bed_text(interval_text: bytes, len_text: bytes):
    """scripts/get_bed_file.py on buffers: returns (bed text, extracted length)."""
    L = load()
    n, total = _sz(0), _l(0)
    _check(L.lhgt_bed_text(interval_text, len(interval_text), len_text, len(len_text), None, 0, C.byref(n), C.byref(total)))
    buf = C.create_string_buffer(max(1, n.value))
    _check(L.lhgt_bed_text(interval_text, len(interval_text), len_text, len(len_text), buf, n.value, C.byref(n), C.byref(total)))
    return buf.raw[:n.value], int(total.value)


def regions_fasta(fasta: bytes, bed: bytes) -> bytes:
    """`samtools faidx -r` on buffers (format per its documentation; see include/lhgt.h)."""
    L = load()
    n = _sz(0)
    src = (C.c_ubyte * max(1, len(fasta))).from_buffer_copy(fasta or b"\0")
    _check(L.lhgt_regions_fasta(src, len(fasta), bed, len(bed), None, 0, C.byref(n)))
    buf = C.create_string_buffer(max(1, n.value))
    _check(L.lhgt_regions_fasta(src, len(fasta), bed, len(bed), buf, n.value, C.byref(n)))
    return buf.raw[:n.value]


def extract_regions_files(fasta: str, interval: str, out_fasta: Optional[str] = None) -> int:
    total = _l(0)
    _check(load().lhgt_extract_regions_files(fasta.encode(), interval.encode(), out_fasta.encode() if out_fasta else None,
                                             C.byref(total)))
    return int(total.value)
