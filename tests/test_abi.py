"""CPU: liblhgt.so loads, exports every symbol include/lhgt.h declares, its host-only helpers agree with
the oracle / glibc, and the GPU entries fail loudly (never fall back) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from localhgt_b200 import api, build as lhgt_build
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "lhgt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lhgt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    names = _declared()
    assert len(names) >= 45
    lhgt_build.build()
    lib = C.CDLL(lhgt_build.LIB)
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/lhgt.h but not exported by liblhgt.so"
    assert set(names) == set(api.SYMBOLS), set(names) ^ set(api.SYMBOLS)
    assert api.load().lhgt_abi_version() == 1
    assert os.path.exists(lhgt_build.EXE)


@pytest.mark.parametrize("seed", [1, 5, 12345, 0])
def test_rand_stream_is_glibc(seed):
    libc = C.CDLL("libc.so.6")
    libc.srand(seed)
    want = [libc.rand() for _ in range(400)]
    assert api.rand_stream(seed, 0, 400).tolist() == want
    assert api.rand_stream(seed, 64, 100).tolist() == want[64:164]


@pytest.mark.parametrize("k,e,seed", [(32, 3, 1), (24, 4, 5), (31, 1, 7), (27, 5, 11), (30, 10, 9)])
def test_random_coder_and_header_round_trip(k, e, seed):
    o = orc.Oracle(k, e); o.srand(seed); want = o.random_coder()
    cc, draws = api.random_coder(seed, k, e)
    assert np.array_equal(cc, want) and draws == k * (e // 3 + 1)
    words = api.coder_to_header(cc)
    assert words[0] == (int(cc[0]) | (int(cc[1]) << 16))               # quirk Q1 (E:755-757)
    assert words[299] >> 16 == 0 and (words[299] & 0xffff) == (int(cc[299]) & 0xffff)   # last high half is zero
    assert np.array_equal(api.header_to_coder(words), cc)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for boxes without one")
    with pytest.raises(api.LhgtError) as ei:
        api.Screen(20, 3)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)
    with pytest.raises(api.LhgtError):
        api.extract_ref("a.fq", "b.fq", "r.fa", "out.txt", k=20)
