"""GPU: the CUDA path, called through the C ABI (ctypes), against
  (1) the oracle port stage by stage on the same inputs,
  (2) the committed outputs of the unmodified reference binary (tests/golden/MANIFEST.json),
  (3) size-independent properties on a workload the oracle could not finish in seconds.
Bit-exact everywhere: this path is integer/byte work.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import fixtures
import model_np
from localhgt_b200 import api, build as lhgt_build, synth
from oracle import orc

pytestmark = pytest.mark.gpu


def _read(path):
    with open(path, "rb") as f:
        return f.read()


# ------------------------------------------------------------------ hashing
@pytest.mark.parametrize("k,e,seed", [(32, 3, 1), (24, 3, 1), (20, 3, 2), (31, 1, 7), (24, 4, 5), (27, 5, 11), (13, 7, 3), (30, 10, 9), (2, 1, 1)])
def test_hash_seq_matches_oracle(k, e, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnRY\r-", dtype=np.uint8)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 5000)].copy()
    noise = rng.integers(0, 5000, 40)
    seq[noise] = alphabet[rng.integers(0, len(alphabet), 40)]
    seq[3000:3100] |= 0x20                                   # lower-case stretch is valid
    seq[1020:1030] = ord("N")                                # straddles the 1024-position tile boundary
    o = orc.Oracle(k, e); o.srand(seed); cc = o.random_coder()
    cc2, draws = api.random_coder(seed, k, e)
    assert np.array_equal(cc, cc2) and draws == k * (e // 3 + 1)
    with api.Screen(k, e) as s:
        s.set_coder(cc)
        for piece in (seq.tobytes(), seq[:k].tobytes(), seq[: k - 1].tobytes(), seq[:1025].tobytes(), seq[7:2055].tobytes()):
            h0, v0 = o.hash_seq(piece)
            h1, v1 = s.hash_seq(piece)
            assert np.array_equal(v0, v1)
            assert np.array_equal(h0, h1)
    # and the numpy statement of the closed form agrees too
    h2, v2 = model_np.hash_seq(seq[:600], k, e, cc)
    h0, v0 = o.hash_seq(seq[:600].tobytes())
    assert np.array_equal(h0, h2) and np.array_equal(v0, v2)


# ------------------------------------------------------------------ stage by stage
def _stage_run(case, work):
    """Runs port and GPU side by side with identical decisions; returns both states."""
    fa, fq1, fq2 = fixtures.materialize(case.data, work)
    k, e = case.k, case.e
    o = orc.Oracle(k, e)
    s = api.Screen(k, e)
    skip = 0
    o.srand(case.seed)
    if not case.prebuilt_index:
        cc = o.random_coder()
        cc2, skip = api.random_coder(case.seed, k, e)
        assert np.array_equal(cc, cc2)
    else:                                     # coder comes from an index built earlier with the same seed
        tmp = orc.Oracle(k, e); tmp.srand(case.seed); cc = tmp.random_coder(); o.set_coder(cc)
    s.set_coder(cc)
    idx = os.path.join(work, case.name + ".stage.index.dat")
    lenp = os.path.join(work, case.name + ".stage.len.txt")
    assert o.index_build(fa, idx, lenp) == 0
    s.index_build(_read(fa))
    return fa, fq1, fq2, o, s, idx, lenp, skip


@pytest.mark.parametrize("case", [c for c in fixtures.CASES if c.k <= 27], ids=lambda c: c.name)
def test_stages_match_oracle(case, manifest, workdir):
    gold = manifest[case.name]
    fa, fq1, fq2, o, s, idx, lenp, skip = _stage_run(case, workdir)
    with s:
        # IB: bit-exact index image and genome.len.txt
        image = s.index_download()
        assert image.tobytes() == _read(idx)
        assert s.index_len_text() == _read(lenp)
        assert s.index_len_text().decode() == gold["len_text"]
        # reads
        b1, b2 = _read(fq1), _read(fq2)
        s.reads_upload(0, b1); s.reads_upload(1, b2)
        ratio_o = orc.sample_ratio(fq1, case.sample)
        ratio_g = s.sample_ratio(case.sample)
        assert ratio_o == ratio_g
        if ratio_o < 100:
            o.fill_random(max(s.reads_records(0), s.reads_records(1)) + 8)
        s.set_sampling(ratio_g, case.seed, skip)
        # S1
        n1o, n2o = o.s1_count(fq1, len(b1), ratio_o), o.s1_count(fq2, len(b1), ratio_o)
        n1g, n2g = s.s1_count(0, len(b1)), s.s1_count(1, len(b1))
        assert (n1o, n2o) == (n1g, n2g)
        assert np.array_equal(o.count_table(), s.count_table())
        # S2
        npo = o.s2_peaks(idx, case.hit, case.match, case.max_peak)
        npg = s.s2_peaks(case.hit, case.match, case.max_peak)
        assert npo == npg == gold["ref_raw_peaks"]
        assert o.raw_positions() == s.flagged_positions()
        loci_g, _ = s.peaks()
        assert np.array_equal(o.peak_loci(), loci_g)
        assert np.array_equal(o.peak_kmer(), s.peak_kmer())
        # S3
        so, sg = o.s3_pairs(fq1, fq2, ratio_o), s.s3_pairs()
        assert so == sg == gold["ref_pairs_s3"]
        _, filt_g = s.peaks()
        assert np.array_equal(o.peak_filter() >= 1, filt_g >= 1)
        # OUT
        assert s.intervals().decode() == gold["interval_text"]
        # a second sample through the same context after reset gives the same answer
        s.reset()
        s.set_sampling(ratio_g, case.seed, skip)
        s.s1_count(0, len(b1)); s.s1_count(1, len(b1))
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == npg
        s.s3_pairs()
        assert s.intervals().decode() == gold["interval_text"]
    o.close()


# ------------------------------------------------------------------ S1: direct probes vs hash streams
def _low_complexity_fastq(path, rng, n_random=3000):
    """Reads that defeat uniform hashing: homopolymers and short tandem repeats put every k-mer of a read
    (and of thousands of reads) into ONE stream, overflowing the shared-memory buckets and a small stream region."""
    seqs = []
    for i in range(1500):
        seqs.append(b"A" * 150)
        seqs.append((b"AC" * 80)[: 100 + i % 50])
        seqs.append((b"GATTACA" * 30)[:150])
        seqs.append(b"T" * 40 + b"N" + b"T" * 60)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(n_random):
        seqs.append(acgt[rng.integers(0, 4, int(rng.integers(20, 250)))].tobytes())
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]
    synth.write_fastq_ragged(path, [b"r%d/1" % i for i in range(len(seqs))], seqs)
    return len(seqs)


@pytest.mark.parametrize("k,e", [(24, 3), (16, 5), (27, 10)])
def test_s1_streams_equal_direct_and_oracle(k, e, workdir, monkeypatch):
    rng = np.random.default_rng(k * 100 + e)
    fq = os.path.join(workdir, f"lowc_{k}_{e}.fq")
    n = _low_complexity_fastq(fq, rng)
    raw = _read(fq)
    o = orc.Oracle(k, e); o.srand(3); cc = o.random_coder()
    o.fill_random(n + 8)
    ratio = 80.0
    want_n = o.s1_count(fq, len(raw), ratio)
    want = o.count_table().copy()
    o.close()
    # modes: direct probes; hash streams with the default leaves (2^18 counters: few leaves at these k); streams with
    # small leaves (the k=32 shape: 2^6 streams x 2^8 leaves each); the same with tiny pools (many chunks, full regions)
    for mode, pool_mb, leaf_log2 in ((1, None, None), (1, None, "10"), (2, None, None), (2, None, "10"), (2, "1", "10")):
        for name, val in (("LHGT_BIN_POOL_MB", pool_mb), ("LHGT_LEAF_LOG2", leaf_log2)):
            if val:
                monkeypatch.setenv(name, val)
            else:
                monkeypatch.delenv(name, raising=False)
        if mode == 2 and leaf_log2 is None and k <= 18:
            continue                                               # a single leaf: nothing to stream
        with api.Screen(k, e) as s:
            s.set_coder(cc)
            s.set_s1_mode(mode)
            s.reads_upload(0, raw)
            s.set_sampling(ratio, 3, k * (e // 3 + 1))
            assert s.s1_count(0, len(raw)) == want_n
            got = s.count_table()
            assert np.array_equal(got, want), (mode, pool_mb, leaf_log2, int((got != want).sum()))
            s.s1_count(0, len(raw))                               # counting again only saturates further
            assert np.array_equal(s.count_table(), np.minimum(3, 2 * want.astype(np.int32)))


@pytest.mark.parametrize("case", [fixtures.BY_NAME[n] for n in ("base_k24", "noisy", "shorts_bp", "fq2_longer")], ids=lambda c: c.name)
def test_streamed_s1_whole_run(case, manifest, workdir):
    """The full pass with S1 forced through the stream path reproduces the reference's interval text."""
    gold = manifest[case.name]
    fa, fq1, fq2, o, s, idx, lenp, skip = _stage_run(case, workdir)
    o.close()
    with s:
        s.set_s1_mode(2)
        b1, b2 = _read(fq1), _read(fq2)
        s.reads_upload(0, b1); s.reads_upload(1, b2)
        s.set_sampling(s.sample_ratio(case.sample), case.seed, skip)
        n1, n2 = s.s1_count(0, len(b1)), s.s1_count(1, len(b1))
        assert (n1 + n2) // 2 == gold["ref_pairs_s1"]
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == gold["ref_raw_peaks"]
        assert s.s3_pairs() == gold["ref_pairs_s3"]
        assert s.intervals().decode() == gold["interval_text"]


# ------------------------------------------------------------------ the whole program, file level
@pytest.mark.parametrize("case", fixtures.CASES, ids=lambda c: c.name)
def test_extract_ref_matches_reference_binary(case, manifest, workdir):
    gold = manifest[case.name]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    fixtures.clean_outputs(fa)
    out = os.path.join(workdir, case.name + ".gpu.interval.txt")
    kw = dict(hit_ratio=case.hit, match_ratio=case.match, k=case.k, e=case.e, seed=case.seed, sample=case.sample,
              max_peak=case.max_peak)
    try:
        if case.prebuilt_index:
            api.extract_ref(fq1, fq2, fa, out + ".first", **kw)
        st = api.extract_ref(fq1, fq2, fa, out, **kw)
        idx = fixtures.index_path(fa, case.k, case.e)
        assert os.path.getsize(idx) == gold["index_bytes"]
        assert fixtures.sha256(idx) == gold["index_sha256"]
        assert _read(fa + ".genome.len.txt").decode() == gold["len_text"]
        assert _read(out).decode() == gold["interval_text"]
        assert (st.reads_s1[0] + st.reads_s1[1]) // 2 == gold["ref_pairs_s1"]
        assert st.pairs_s3 == gold["ref_pairs_s3"]
        assert st.peaks == gold["ref_raw_peaks"]
        assert bool(st.index_built) == (not case.prebuilt_index)
    finally:
        fixtures.clean_outputs(fa)


def test_cli_binary_is_a_drop_in(manifest, workdir):
    """The `extract_ref` executable with the 12 positional arguments of scripts/pipeline.sh:35."""
    case = fixtures.BY_NAME["noisy"]
    gold = manifest[case.name]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    fixtures.clean_outputs(fa)
    out = os.path.join(workdir, "cli.interval.txt")
    exe = lhgt_build.EXE
    assert os.path.exists(exe)
    # Python's f-string spelling of the numbers, as infer_HGT_breakpoint.py:29 passes them
    argv = [exe, fq1, fq2, fa, out, f"{case.hit}", f"{case.match}", "10", f"{case.k}", f"{float(case.max_peak)}",
            f"{case.e}", f"{case.seed}", f"{case.sample}"]
    try:
        r = subprocess.run(argv, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert _read(out).decode() == gold["interval_text"]
        assert fixtures.sha256(fixtures.index_path(fa, case.k, case.e)) == gold["index_sha256"]
        # second run reuses the index file (E:1403-1413) and, at ratio<100, samples differently (Q3)
        r2 = subprocess.run(argv, capture_output=True, text=True, timeout=600)
        assert r2.returncode == 0 and "Reference index is detected." in r2.stdout
    finally:
        fixtures.clean_outputs(fa)
    bad = subprocess.run([exe, "a", "b"], capture_output=True, text=True)
    assert bad.returncode != 0 and "usage" in bad.stderr


# ------------------------------------------------------------------ error behaviour
def test_errors_are_loud(workdir):
    with pytest.raises(api.LhgtError):
        api.Screen(33, 3)
    with pytest.raises(api.LhgtError):
        api.Screen(32, 11)
    with api.Screen(16, 3) as s:
        with pytest.raises(api.LhgtError):
            s.s1_count(0, 10)                                   # nothing uploaded
        with pytest.raises(api.LhgtError):
            s.index_upload(b"\0" * 100)                        # shorter than the header
        long_read = b"@r\n" + b"A" * 600 + b"\n+\n" + b"I" * 600 + b"\n"
        s.reads_upload(0, long_read); s.reads_upload(1, long_read)
        s.set_sampling(100.0)
        with pytest.raises(api.LhgtError) as ei:
            s.s1_count(0, len(long_read))
        assert ei.value.code == -6
    with pytest.raises(api.LhgtError):
        api.extract_ref("/nonexistent.1.fq", "/nonexistent.2.fq", "/nonexistent.fa", os.path.join(workdir, "x.txt"), k=16)


def test_empty_and_degenerate_inputs(workdir):
    fa = os.path.join(workdir, "tiny.fa")
    with open(fa, "wb") as f:
        f.write(b">only_short\nACGTACGT\n>also_short\nAC\n")          # no contig longer than k
    fq = os.path.join(workdir, "empty.fq")
    open(fq, "wb").close()
    out = os.path.join(workdir, "tiny.interval.txt")
    fixtures.clean_outputs(fa)
    try:
        st = api.extract_ref(fq, fq, fa, out, k=16, e=3)
        assert _read(out) == b"1\t1\t1\n"                                # Q9: the initial state is always printed
        assert st.peaks == 0 and st.pairs_s3 == 0
        assert os.path.getsize(fixtures.index_path(fa, 16, 3)) == 1200
        assert _read(fa + ".genome.len.txt") == b""
        # the oracle port says the same
        fixtures.clean_outputs(fa)
        rc, _ = orc.extract_ref(fq, fq, fa, out + ".port", k=16, e=3)
        assert rc == 0 and _read(out + ".port") == b"1\t1\t1\n"
    finally:
        fixtures.clean_outputs(fa)


@pytest.mark.parametrize("k,n_genomes", [(14, 90), (16, 48)])
def test_s3_vote_with_many_contigs(k, n_genomes, workdir):
    """A tiny hash space fills the peak table, so nearly every read position holds a peak k-mer of some random contig:
    pairs vote for far more than 32 contigs, which takes S3 off the in-register tally onto the direct-addressed one."""
    e, seed = 3, 2
    w = synth.make_workload(workdir, f"many_{k}", seed=k, n_genomes=n_genomes, genome_len=6000, n_pairs=4000, n_events=6,
                            seg_len=(300, 900))
    o = orc.Oracle(k, e); o.srand(seed); cc = o.random_coder()
    idx, lenp = w.ref_fa + ".many.index.dat", w.ref_fa + ".many.len.txt"
    assert o.index_build(w.ref_fa, idx, lenp) == 0
    b1, b2 = _read(w.fq1), _read(w.fq2)
    with api.Screen(k, e) as s:
        s.set_coder(cc)
        s.index_build(_read(w.ref_fa))
        s.reads_upload(0, b1); s.reads_upload(1, b2)
        s.set_sampling(100.0, seed, 0)
        assert (o.s1_count(w.fq1, len(b1), 100.0), o.s1_count(w.fq2, len(b1), 100.0)) == (s.s1_count(0, len(b1)), s.s1_count(1, len(b1)))
        npo, npg = o.s2_peaks(idx, 0.1, 0.08, 1000000), s.s2_peaks(0.1, 0.08, 1000000)
        assert npo == npg and npg >= 40
        loci_g, _ = s.peaks()
        assert len(set(loci_g[:, 0].tolist())) > 32, "the case must spread its peaks over more than 32 contigs"
        assert np.array_equal(o.peak_kmer(), s.peak_kmer())
        assert o.s3_pairs(w.fq1, w.fq2, 100.0) == s.s3_pairs()
        _, filt_g = s.peaks()
        assert np.array_equal(o.peak_filter() >= 1, filt_g >= 1)
        assert int((filt_g >= 1).sum()) > 0
        out = w.ref_fa + ".many.interval.txt"
        o.write_intervals(out)
        assert s.intervals() == _read(out)
    o.close()


@pytest.mark.parametrize("shape", ["leading_sequence", "no_final_newline", "gt_inside_line", "blank_lines_crlf", "long_lines", "headers_only"])
def test_fasta_shapes_match_oracle_index(shape, workdir):
    """The FASTA is parsed on the device (header lines found, sequence compacted there); the oracle walks it line by line."""
    rng = np.random.default_rng(len(shape))
    acgt = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    def seq(n):
        return acgt[rng.integers(0, len(acgt), n)].tobytes()
    def wrap(s, w):
        return b"\n".join(s[i:i + w] for i in range(0, len(s), w))
    k, e = 21, 3
    if shape == "leading_sequence":          # bytes before the first header form a contig named "start" (E:747)
        text = wrap(seq(700), 60) + b"\n>c1 desc\n" + wrap(seq(20000), 70) + b"\n>c2/1\n" + wrap(seq(16384 * 2 + 5), 80) + b"\n"
    elif shape == "no_final_newline":
        text = b">a\n" + wrap(seq(5000), 61) + b"\n>b\tx\n" + wrap(seq(3000), 61)
    elif shape == "gt_inside_line":          # a '>' that does not start a line is just a (bad) base
        text = b">a\n" + seq(400) + b">not_a_header" + seq(300) + b"\n" + seq(50) + b"\n>b\n" + wrap(seq(4000), 100) + b"\n"
    elif shape == "blank_lines_crlf":
        text = b">a\n\n" + wrap(seq(3000), 50).replace(b"\n", b"\r\n") + b"\r\n\n\n>b\n" + seq(2500) + b"\n\n"
    elif shape == "long_lines":              # one unwrapped line longer than several tiles, header spans across a tile edge
        text = b">" + b"x" * 16380 + b" tail\n" + seq(70000) + b"\n>s\n" + seq(10) + b"\n>t\n" + seq(40000) + b"\n"
    else:
        text = b">a\n>b\n>c\n"
    fa = os.path.join(workdir, f"shape_{shape}.fa")
    open(fa, "wb").write(text)
    o = orc.Oracle(k, e); o.srand(5); cc = o.random_coder()
    idx, lenp = fa + ".oracle.index.dat", fa + ".oracle.len.txt"
    assert o.index_build(fa, idx, lenp) >= 0
    o.close()
    with api.Screen(k, e) as s:
        s.set_coder(cc)
        s.index_build(text)
        assert s.index_len_text() == _read(lenp)
        assert bytes(s.index_download()) == _read(idx)
        # the same from text that already lives on the device
        import torch
        t = torch.empty(len(text) + 64, dtype=torch.uint8, device="cuda")
        t[:len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
        s.index_build_device(t.data_ptr(), len(text))
        assert s.index_len_text() == _read(lenp)
        assert bytes(s.index_download()) == _read(idx)


@pytest.mark.slow
def test_fasta_beyond_4_gib_keeps_64_bit_offsets():
    """A 4.4 GB FASTA (64-bit file offsets and sequence counts in the device scan): 2 200 contigs whose
    bases are a function of their ordinal; lengths, names and a sample of index records -- including the last contig's,
    which sits beyond byte 2^32 of the file and of the compacted sequence -- are checked against the oracle's hashes of
    the same bases.  k = 20, e = 1 keeps the image at 17 GB."""
    import torch
    from localhgt_b200 import synth_dev
    spec = synth_dev.Spec("wide", 2200, 2_000_000, 10, 2, seed=6)
    fa = synth_dev.make_fasta(spec, "cuda")
    assert fa.numel() > (1 << 32)
    k, e = 20, 1
    o = orc.Oracle(k, e); o.srand(3); cc = o.random_coder()
    buf = torch.empty(fa.numel() + 64, dtype=torch.uint8, device="cuda")
    buf[:fa.numel()] = fa
    del fa
    with api.Screen(k, e) as s:
        s.set_coder(cc)
        s.index_build_device(buf.data_ptr(), buf.numel() - 64)
        lens = s.index_len_text().decode().splitlines()
        assert len(lens) == 2200 and s.index_bases() == 2200 * 2_000_000
        assert lens[0] == "g0\t1\t2000000\t2000000" and lens[3] == "g3\t5\t2000000\t8000020"
        assert lens[-1] == f"g2199\t2201\t2000000\t{2200 * 2000000 + 20}"
        per = 1 + (2_000_000 - k + 1) * e
        assert s.index_bytes() == 1200 + 4 * per * 2200
        for gi in (0, 1100, 2199):
            seq = synth_dev.contig_bases(spec, gi, "cpu").numpy().tobytes()
            want, _ = o.hash_seq(seq)
            got = s.index_record(gi)
            assert got[0] == 2_000_000
            assert np.array_equal(got[1:], want.reshape(-1))
    o.close()


# ------------------------------------------------------------------ properties at a size the oracle cannot do in seconds
@pytest.fixture(scope="module")
def big(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("big"))
    w = synth.make_workload(d, "big", seed=9, n_genomes=20, genome_len=400000, n_pairs=400000, n_events=8,
                            ref_n_rate=0.0001, short_contigs=(20, 31))
    return w


def test_large_run_properties(big):
    k, e = 32, 3
    fa, b1, b2 = _read(big.ref_fa), _read(big.fq1), _read(big.fq2)
    cc, _ = api.random_coder(1, k, e)
    with api.Screen(k, e) as s:
        s.set_coder(cc)
        s.index_build(fa)
        image = s.index_download()
        # index: header round trip and record structure
        assert np.array_equal(api.header_to_coder(image[:1200].view(np.uint32)), cc)
        assert s.index_bytes() == 1200 + 4 * sum(1 + (len(q) - k + 1) * e for q in _contigs(fa) if len(q) > k)
        # spot-check 200 random positions of the image against the scalar oracle
        o = orc.Oracle(k, e); o.set_coder(cc)
        words = image.view(np.uint32)
        rng = np.random.default_rng(3)
        at = 300
        for q in _contigs(fa):
            if len(q) <= k:
                continue
            assert words[at] == len(q)
            for j in rng.integers(0, len(q) - k + 1, 10):
                h, _ = o.hash_seq(q[j:j + k])
                assert np.array_equal(words[at + 1 + j * e: at + 1 + (j + 1) * e], h[0])
            at += 1 + (len(q) - k + 1) * e
        s.reads_upload(0, b1); s.reads_upload(1, b2)
        assert s.reads_records(0) == s.reads_records(1) == big.n_pairs
        assert s.reads_seq_bases(0) == big.n_pairs * 150
        s.set_sampling(100.0)
        assert s.s1_count(0, len(b1)) == big.n_pairs
        assert s.s1_count(1, len(b1)) == big.n_pairs
        n_peaks = s.s2_peaks()
        s.s3_pairs()
        text = s.intervals()
        loci, filt = s.peaks()
        # peaks are strictly ordered by (contig, position) and one per 50-bp bucket (E:288-301)
        key = loci[:, 0].astype(np.int64) * (1 << 32) + loci[:, 1]
        assert np.all(np.diff(key) > 0)
        bucket = loci[:, 0].astype(np.int64) * (1 << 32) + loci[:, 1] // 50
        assert len(np.unique(bucket)) == n_peaks
        # every planted junction on a recipient lies strictly inside an emitted interval (evaluation.py:64-76)
        ivs = [tuple(map(int, ln.split("\t"))) for ln in text.decode().splitlines()]
        names = [ln[1:].split()[0] for ln in fa.decode().splitlines() if ln.startswith(">")]
        long_names = [n for n, q in zip(names, _contigs(fa)) if len(q) > k]
        found = 0
        for ev in big.truth:
            ordinal = long_names.index(names[ev.recipient]) + 1
            found += any(c == ordinal and a + 50 < ev.r_pos < b - 50 for c, a, b in ivs)
        assert found >= 0.75 * len(big.truth), (found, len(big.truth))
        # idempotence: the same sample again after reset -> identical text; uploading the image instead of building -> identical
        s.reset()
        s.index_upload(image)
        s.set_sampling(100.0)
        s.s1_count(0, len(b1)); s.s1_count(1, len(b1))
        assert s.s2_peaks() == n_peaks
        s.s3_pairs()
        assert s.intervals() == text
        # counting is commutative and saturating: mates in the other order give the same table checksum
        tab = s.count_table()
        s.reset(); s.set_sampling(100.0)
        s.s1_count(1, len(b1)); s.s1_count(0, len(b1))
        assert np.array_equal(np.bincount(tab, minlength=4), np.bincount(s.count_table(), minlength=4))


def _contigs(fa: bytes):
    out, cur = [], []
    for ln in fa.split(b"\n"):
        if ln.startswith(b">"):
            out.append(b"".join(cur))         # first append is the (empty) text before the first header
            cur = []
        else:
            cur.append(ln)
    out.append(b"".join(cur))
    return out[1:]


@pytest.mark.slow
def test_large_run_matches_reference_binary_on_this_box(big, tmp_path):
    """k=32 end to end against the real reference executable (needs ~21 GB host RAM, ~2-3 min)."""
    if not os.path.exists(orc.REF_BIN_Z):
        pytest.skip("oracle/_ref/extract_ref_z not present")
    n = 40000
    lines = 4 * n
    def head(src, dst):
        with open(src, "rb") as f, open(dst, "wb") as g:
            for _ in range(lines):
                g.write(f.readline())
    d = str(tmp_path)
    fq1, fq2 = os.path.join(d, "h.1.fq"), os.path.join(d, "h.2.fq")
    head(big.fq1, fq1); head(big.fq2, fq2)
    fa_r, fa_g = os.path.join(d, "r", "ref.fa"), os.path.join(d, "g", "ref.fa")
    os.makedirs(os.path.dirname(fa_r)); os.makedirs(os.path.dirname(fa_g))
    shutil.copy(big.ref_fa, fa_r); shutil.copy(big.ref_fa, fa_g)
    orc.run_reference(fq1, fq2, fa_r, os.path.join(d, "r.txt"), k=32, e=3, seed=1, sample=0.9, max_peak=2000000, timeout=1500)
    api.extract_ref(fq1, fq2, fa_g, os.path.join(d, "g.txt"), k=32, e=3, seed=1, sample=0.9, max_peak=2000000)
    assert fixtures.sha256(fa_r + ".k32.h3.index.dat") == fixtures.sha256(fa_g + ".k32.h3.index.dat")
    assert _read(fa_r + ".genome.len.txt") == _read(fa_g + ".genome.len.txt")
    assert _read(os.path.join(d, "r.txt")) == _read(os.path.join(d, "g.txt"))


@pytest.mark.parametrize("k", [3, 20, 24])
def test_count_table_histogram(k, workdir):
    """the occupancy diagnostic of count_diff_kmer.cpp as a reduction over the packed table."""
    rng = np.random.default_rng(k)
    fq = os.path.join(workdir, f"hist_{k}.fq")
    n = _low_complexity_fastq(fq, rng, n_random=1500)
    raw = _read(fq)
    with api.Screen(k, 3) as s:
        h0 = s.count_table_histogram()
        assert h0.tolist() == [1 << k, 0, 0, 0]
        s.reads_upload(0, raw)
        s.set_sampling(100.0, 1, 0)
        assert s.s1_count(0, len(raw)) == n
        want = np.bincount(s.count_table(), minlength=4)
        assert s.count_table_histogram().tolist() == want.tolist()
        assert want[3] > 0


def test_ordinal_base_makes_shards_equal_the_whole_sample(workdir):
    """The multi-GPU plan gives every rank a record range and the ordinal of its first record (multi.Shard.screen);
    with a sampling ratio below 100 % the sampled subset must be the whole run's.  One GPU, two shards fed one after
    the other into the same table: counts, sampled-read totals and the S3 verdicts must match the unsharded run."""
    case = fixtures.BY_NAME["half_build"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    k, e, seed, ratio = 20, 3, 7, 37.5
    b1, b2 = _read(fq1), _read(fq2)
    cc, skip = api.random_coder(seed, k, e)

    def cut(buf, n_rec):
        pos = 0
        for _ in range(4 * n_rec):
            pos = buf.index(b"\n", pos) + 1
        return buf[:pos], buf[pos:]

    n_rec = b1.count(b"\n") // 4
    first = n_rec * 3 // 7
    a1, z1 = cut(b1, first)
    a2, z2 = cut(b2, first)
    with api.Screen(k, e) as whole, api.Screen(k, e) as parts:
        for s in (whole, parts):
            s.set_coder(cc)
            s.index_build(_read(fa))
        whole.reads_upload(0, b1); whole.reads_upload(1, b2)
        whole.set_sampling(ratio, seed, skip)
        n_whole = (whole.s1_count(0, len(b1)), whole.s1_count(1, len(b1)))
        assert 0 < n_whole[0] < n_rec
        got = [0, 0]
        for base, (m1, m2) in ((0, (a1, a2)), (first, (z1, z2))):
            parts.reads_upload(0, m1); parts.reads_upload(1, m2)
            parts.set_ordinal_base(base)
            parts.set_sampling(ratio, seed, skip)
            got[0] += parts.s1_count(0, len(b1)); got[1] += parts.s1_count(1, len(b1))
        assert tuple(got) == n_whole
        assert np.array_equal(whole.count_table(), parts.count_table())
        assert whole.s2_peaks(case.hit, case.match, case.max_peak) == parts.s2_peaks(case.hit, case.match, case.max_peak)
        # S3 on the second shard only sees its own pairs; OR-ing the verdicts of both shards gives the whole run's
        n3_whole = whole.s3_pairs()
        _, f_whole = whole.peaks()
        n3 = parts.s3_pairs()                                    # second shard is resident, base = first
        _, f_b = parts.peaks()
        parts.reads_upload(0, a1); parts.reads_upload(1, a2)
        parts.set_ordinal_base(0); parts.set_sampling(ratio, seed, skip)
        n3 += parts.s3_pairs()
        _, f_ab = parts.peaks()                                  # the filter accumulates across s3 calls
        assert n3 == n3_whole
        assert np.array_equal(f_whole >= 1, f_ab >= 1) and not np.any((f_b >= 1) & ~(f_ab >= 1))


# ------------------------------------------------------------------ S2 in its steps, count merge, S3 vote paths
def _screened(case, workdir, **env):
    """One GPU context with `case` loaded and S1 done; returns (screen, oracle with S1 + S2 done, ratio, skip, paths)."""
    fa, fq1, fq2, o, s, idx, lenp, skip = _stage_run(case, workdir)
    b1, b2 = _read(fq1), _read(fq2)
    s.reads_upload(0, b1); s.reads_upload(1, b2)
    ratio = s.sample_ratio(case.sample)
    if ratio < 100:
        o.fill_random(max(s.reads_records(0), s.reads_records(1)) + 8)
    s.set_sampling(ratio, case.seed, skip)
    o.s1_count(fq1, len(b1), ratio); o.s1_count(fq2, len(b1), ratio)
    s.s1_count(0, len(b1)); s.s1_count(1, len(b1))
    return s, o, ratio, (fq1, fq2, idx)


@pytest.mark.parametrize("sliced", [None, "big", "64"], ids=["direct", "sliced", "sliced_small_regions"])
@pytest.mark.parametrize("name", ["base_k24", "base_k20", "noisy", "shorts_bp", "base_k24_e4", "base_k31_e1"])
def test_s2_in_steps_over_tile_ranges_equals_the_oracle(name, sliced, workdir, monkeypatch):
    """gather / complete on tile ranges (what each rank of the multi-GPU plan runs) + mark + finish == one s2_peaks ==
    the oracle: the short-circuit trio gather, the hot/needed tile marking and the needed-tile passes drop nothing.
    `sliced` runs the gather through table slices instead (records bucketed by slice, answered from L2; forced here at
    small k), once with record regions so small that every call takes several chunks and overflows."""
    case = fixtures.BY_NAME[name]
    if sliced:
        monkeypatch.setenv("LHGT_S2_SLICED", "1")
        if sliced != "big":
            monkeypatch.setenv("LHGT_S2_POOL_KB", sliced)
    else:
        monkeypatch.setenv("LHGT_S2_SLICED", "0")
    s, o, ratio, (fq1, fq2, idx) = _screened(case, workdir)
    with s:
        npo = o.s2_peaks(idx, case.hit, case.match, case.max_peak)
        nt = s.s2_tiles()
        cuts = sorted({0, nt // 3, nt // 3 + 1, (2 * nt) // 3, nt})
        for a, b in zip(cuts, cuts[1:]):
            s.s2_gather(a, b)
        s.s2_mark(case.match)
        for a, b in zip(reversed(cuts[:-1]), reversed(cuts[1:])):
            s.s2_complete(a, b)
        assert s.s2_finish(case.hit, case.match, case.max_peak) == npo
        assert 0 <= s.s2_needed_tiles() <= nt
        assert o.raw_positions() == s.flagged_positions()
        loci, _ = s.peaks()
        assert np.array_equal(o.peak_loci(), loci)
        assert np.array_equal(o.peak_kmer(), s.peak_kmer())
        so, sg = o.s3_pairs(fq1, fq2, ratio), s.s3_pairs()
        assert so == sg
        assert np.array_equal(o.peak_filter() >= 1, s.peaks()[1] >= 1)
    o.close()


def test_s2_marks_few_tiles_when_nothing_is_saturated(workdir):
    """A sample that shares nothing with the reference: no hot tile, no needed tile, no peak -- and S2 says so without
    visiting the windows."""
    case = fixtures.BY_NAME["base_k24"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    rng = np.random.default_rng(5)
    other = synth.Reference(["x"], [synth.random_genome(rng, 300000)])
    with api.Screen(case.k, case.e) as s:
        cc, _ = api.random_coder(case.seed, case.k, case.e)
        s.set_coder(cc)
        s.index_build(b">x\n" + other.seqs[0].tobytes() + b"\n")
        b1, b2 = _read(fq1), _read(fq2)
        s.reads_upload(0, b1); s.reads_upload(1, b2)
        s.set_sampling(100.0, case.seed, 0)
        s.s1_count(0, len(b1)); s.s1_count(1, len(b1))
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == 0
        assert s.s2_needed_tiles() == 0 and s.flagged_positions() == 0
        s.s3_pairs()
        assert s.intervals() == b"1\t1\t1\n"


@pytest.mark.parametrize("k", [16, 24])
def test_count_merge_is_a_saturating_add(k, workdir):
    """lhgt_count_merge on the packed 2-bit tables (the NCCL form of the multi-GPU count exchange) == min(3, a + b)."""
    case = fixtures.BY_NAME["base_k24"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    b1, b2 = _read(fq1), _read(fq2)
    cc, _ = api.random_coder(3, k, 3)
    import torch
    with api.Screen(k, 3) as a, api.Screen(k, 3) as b:
        for s, buf in ((a, b1), (b, b2)):
            s.set_coder(cc)
            s.reads_upload(0, buf)
            s.set_sampling(100.0, 1, 0)
            s.s1_count(0, len(buf))
            if s is b:
                s.s1_count(0, len(buf))                    # twice: plenty of 2s and 3s on this side
        ta, tb = a.count_table(), b.count_table()
        assert ta.max() == 3 or tb.max() == 3
        ptr, n = b.dev_count_table()
        half = (n // 8) * 4                                # a word-aligned split: two merges of sub-ranges
        a.count_merge(ptr, half, 0)
        a.count_merge(ptr + half, n - half, half // 4)
        a.sync()
        assert np.array_equal(a.count_table(), np.minimum(3, ta.astype(np.int32) + tb).astype(np.uint8))


@pytest.mark.parametrize("name", ["base_k20", "noisy"])
def test_s3_vote_paths_agree(name, workdir, monkeypatch):
    """S3's order-dependent vote runs in s3_vote_kernel (one thread per pair, candidates handed over through an arena);
    a 1 MiB arena forces several batches and overflow into the in-warp vote, 0 disables the hand-over: same verdicts."""
    case = fixtures.BY_NAME[name]
    s, o, ratio, (fq1, fq2, idx) = _screened(case, workdir)
    with s:
        assert o.s2_peaks(idx, case.hit, case.match, case.max_peak) == s.s2_peaks(case.hit, case.match, case.max_peak)
        o.s3_pairs(fq1, fq2, ratio)
        want = o.peak_filter() >= 1
        assert want.any()
        for arena_mb in (None, "1", "0"):
            if arena_mb is None:
                monkeypatch.delenv("LHGT_S3_ARENA_MB", raising=False)
            else:
                monkeypatch.setenv("LHGT_S3_ARENA_MB", arena_mb)
            s.s2_finish(case.hit, case.match, case.max_peak)      # clears the verdicts (same peaks)
            s.s3_pairs()
            assert np.array_equal(want, s.peaks()[1] >= 1), arena_mb
    o.close()


# ------------------------------------------------------------------ streamed ingest, fq2 re-synchronisation
def test_streamed_ingest_through_a_small_ring(manifest, workdir, monkeypatch):
    """Files are read through the pinned staging ring; 8 KiB chunks make every file wrap the ring many times, with the
    chunk edges falling inside lines and records."""
    case = fixtures.BY_NAME["fq2_longer"]                      # also the case whose last line has no newline
    gold = manifest[case.name]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    fixtures.clean_outputs(fa)
    monkeypatch.setenv("LHGT_RING_KB", "8")
    out = os.path.join(workdir, "ring.interval.txt")
    try:
        for _ in range(2):                                     # builds the index, then loads the file it wrote
            api.extract_ref(fq1, fq2, fa, out, hit_ratio=case.hit, match_ratio=case.match, k=case.k, max_peak=case.max_peak, e=case.e,
                            seed=case.seed, sample=case.sample)
            assert _read(out).decode() == gold["interval_text"]
        assert fixtures.sha256(fixtures.index_path(fa, case.k, case.e)) == gold["index_sha256"]
    finally:
        fixtures.clean_outputs(fa)


def test_fq2_with_leading_records_is_resynchronised_like_the_reference(workdir):
    """E:368-399: when the first read ids differ the reference scans fq2 for fq1's first id and pairs from there.  Three
    stray records in front of fq2: same intervals as the unmodified reference binary (run here, k = 20)."""
    if not os.path.exists(orc.REF_BIN_Z):
        pytest.skip("oracle/_ref not built")
    case = fixtures.BY_NAME["base_k20"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    stray = b"".join(b"@stray%d/2\n" % i + b"ACGT" * 30 + b"\n+\n" + b"I" * 120 + b"\n" for i in range(3))
    fq2s = os.path.join(workdir, "stray.2.fq")
    open(fq2s, "wb").write(stray + _read(fq2))
    d = os.path.join(workdir, "stray_ref"); os.makedirs(d, exist_ok=True)
    fa_ref = os.path.join(d, "ref.fa"); shutil.copy(fa, fa_ref)
    out_ref, out_gpu = os.path.join(d, "ref.txt"), os.path.join(d, "gpu.txt")
    orc.run_reference(fq1, fq2s, fa_ref, out_ref, hit=case.hit, match=case.match, k=case.k, max_peak=case.max_peak, e=case.e, seed=case.seed,
                      sample=case.sample, timeout=600)
    fixtures.clean_outputs(fa)
    try:
        api.extract_ref(fq1, fq2s, fa, out_gpu, hit_ratio=case.hit, match_ratio=case.match, k=case.k, max_peak=case.max_peak, e=case.e,
                        seed=case.seed, sample=case.sample)
        assert _read(out_gpu) == _read(out_ref)
        assert len(_read(out_gpu).splitlines()) > 1
        with pytest.raises(api.LhgtError) as ei:               # no record of fq2 is named like fq1's first
            other = os.path.join(workdir, "other.2.fq")
            open(other, "wb").write(_read(fq2).replace(b"@r", b"@q"))
            api.extract_ref(fq1, other, fa, out_gpu, k=case.k, e=case.e, max_peak=case.max_peak)
        assert ei.value.code == -7
    finally:
        fixtures.clean_outputs(fa)


@pytest.mark.parametrize("name,pool_kb", [("base_k20", None), ("base_k20", "64"), ("noisy", "16"), ("base_k24_e4", None)])
def test_bucketed_registration_equals_direct(name, pool_kb, workdir, monkeypatch):
    """Peak registration through hash buckets (records staged in shared memory, appended in runs, applied bucket by bucket
    against L2-resident table slices) == the direct scatter-max == the oracle's peak_kmer; small record regions force
    several chunks and the overflow path."""
    case = fixtures.BY_NAME[name]
    s, o, ratio, (fq1, fq2, idx) = _screened(case, workdir)
    with s:
        npo = o.s2_peaks(idx, case.hit, case.match, case.max_peak)
        want = o.peak_kmer().copy()
        monkeypatch.setenv("LHGT_REG_BUCKETED", "1")
        if pool_kb:
            monkeypatch.setenv("LHGT_REG_POOL_KB", pool_kb)
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == npo
        assert np.array_equal(want, s.peak_kmer())
        assert np.array_equal(o.peak_loci(), s.peaks()[0])
        monkeypatch.setenv("LHGT_REG_BUCKETED", "0")
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == npo      # un-writes / clears, registers again directly
        assert np.array_equal(want, s.peak_kmer())
        monkeypatch.setenv("LHGT_REG_BUCKETED", "1")
        assert s.s2_peaks(case.hit, case.match, case.max_peak) == npo
        so, sg = o.s3_pairs(fq1, fq2, ratio), s.s3_pairs()
        assert so == sg and np.array_equal(o.peak_filter() >= 1, s.peaks()[1] >= 1)
    o.close()


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_image_blocks_tile_the_image(parts, workdir):
    """lhgt_set_image_block: the blocks the ranks of a box build, laid side by side at the offsets they report, are the
    index file byte for byte (header in block 0, every contig's length word with its first tile); written with
    lhgt_index_write_block they make the same file."""
    case = fixtures.BY_NAME["noisy"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    k, e = 22, 3
    cc, _ = api.random_coder(4, k, e)
    with api.Screen(k, e) as whole:
        whole.set_coder(cc)
        whole.index_build(_read(fa))
        image = bytes(whole.index_download())
        ntiles = whole.s2_tiles()
    path = os.path.join(workdir, f"blocks{parts}.index.dat")
    if os.path.exists(path):
        os.remove(path)
    got = bytearray(len(image))
    covered = 0
    for part in range(parts):
        with api.Screen(k, e) as s:
            s.set_coder(cc)
            s.set_image_block(part, parts)
            s.index_build(_read(fa))
            off, n, t0, t1 = s.index_block()
            assert off == covered and 0 <= t0 <= t1 <= ntiles
            got[off:off + n] = bytes(s.index_download())
            covered += n
            s.index_write_block(path)
            if t1 > t0:
                with pytest.raises(api.LhgtError):
                    s.s2_gather(0, ntiles) if parts > 1 and (t0, t1) != (0, ntiles) else (_ for _ in ()).throw(api.LhgtError(-8, "x"))
    assert covered == len(image) and bytes(got) == image
    assert _read(path) == image


def test_next_sample_prefetch_swaps_the_right_images(workdir):
    """lhgt_reads_prefetch_next: while sample A is screened, sample B's FASTQ images cross PCIe into the alternate buffers and
    B's upload adopts them (and vice versa, several rounds): every sample gets exactly the answer it gets on its own."""
    import torch
    case = fixtures.BY_NAME["base_k24"]
    fa, fq1, fq2 = fixtures.materialize(case.data, workdir)
    a1, a2 = _read(fq1), _read(fq2)

    def head(buf, n_rec):
        pos = 0
        for _ in range(4 * n_rec):
            pos = buf.index(b"\n", pos) + 1
        return buf[:pos]

    b1, b2 = head(a1, 2500), head(a2, 2500)
    cc, skip = api.random_coder(case.seed, case.k, case.e)

    def alone(m1, m2):
        with api.Screen(case.k, case.e) as s:
            s.set_coder(cc); s.index_build(_read(fa))
            s.reads_upload(0, m1); s.reads_upload(1, m2)
            s.set_sampling(100.0, case.seed, skip)
            s.s1_count(0, len(m1)); s.s1_count(1, len(m1))
            s.s2_peaks(case.hit, case.match, case.max_peak); s.s3_pairs()
            return s.intervals()

    want = {"A": alone(a1, a2), "B": alone(b1, b2)}
    assert want["A"] != want["B"]
    pin = {k: torch.frombuffer(bytearray(v), dtype=torch.uint8).pin_memory() for k, v in (("A1", a1), ("A2", a2), ("B1", b1), ("B2", b2))}
    with api.Screen(case.k, case.e) as s:
        s.set_coder(cc); s.index_build(_read(fa))
        order = ["A", "B", "B", "A", "B"]
        for i, name in enumerate(order):
            m1, m2 = pin[name + "1"], pin[name + "2"]
            s.reset()
            s.reads_upload_ptr(0, m1.data_ptr(), m1.numel()); s.reads_upload_ptr(1, m2.data_ptr(), m2.numel())
            if i + 1 < len(order):
                n1, n2 = pin[order[i + 1] + "1"], pin[order[i + 1] + "2"]
                s.reads_prefetch_next_ptr(0, n1.data_ptr(), n1.numel()); s.reads_prefetch_next_ptr(1, n2.data_ptr(), n2.numel())
            s.set_sampling(100.0, case.seed, skip)
            s.s1_count(0, m1.numel()); s.s1_count(1, m1.numel())
            s.s2_peaks(case.hit, case.match, case.max_peak); s.s3_pairs()
            assert s.intervals() == want[name], (i, name)
