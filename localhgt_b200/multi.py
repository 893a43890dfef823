"""One process per GPU: the screen of one sample whose read pairs are partitioned across ranks.

What is sharded and what is exchanged:

  S1   each rank counts the k-mers of ITS record range into its own 2^k table; the tables are then
       combined field-wise with min(3, a+b) — exact because min(3, sum) == min(3, sum of min(3, .)).
       Exchange, preferred form: the ranks map each other's tables (CUDA IPC) and rank r runs ONE kernel that
       reads slice r of every table over NVLink, merges, and stores the result into every table
       (lhgt_count_exchange_p2p) — no staging buffer, no separate merge pass.  NCCL form (when IPC is not
       available): all-to-all of table slices, local merge, all-gather of the merged slices.  Either way
       2 x table bytes cross each rank's links instead of (N-1) x.
  S2   the table gather runs on equal blocks of reference tiles and the per-position hit bits (2 bits per
       reference base) are all-gathered in place; the tiles where anything can happen are marked (replicated,
       cheap); windows and peak registration run on equal SHARES OF THE MARKED TILES (they cluster where the
       sample's genomes are), with the per-tile new-peak counts summed so that every rank derives the same peak
       ids.  Dense results: every rank registers its share and the peak tables are MAX-combined (ids grow with
       position, so MAX = the last writer of the reference's sequential loop); sparse results: the flagged bits
       are combined and everybody registers everything.  With the image kept in blocks (set_image_block, indexes
       larger than one GPU) registration follows the image blocks instead.
  S3   each rank confirms peaks with its own pairs; only `peak_filter >= 1` is consumed (E:526), so the
       verdict bytes are max-reduced.
  OUT  identical on every rank; rank 0's text is the result.

Result = the reference's `-t 1` run over the concatenation of the shards in rank order: sampling
ordinals are offset by the records of the preceding shards and the fq2 byte budget (quirk Q15) is
translated into shard-local offsets.

The collectives are torch.distributed (NCCL on GPUs; gloo in the CPU test).  The engine underneath is
liblhgt through api.Screen; tests/test_multi_gloo.py substitutes the oracle to exercise this file's
logic on CPU.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np


class _DevMem:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def split_range(n: int, parts: int, i: int):
    """Contiguous near-equal split of range(n): part i of `parts`."""
    base, extra = divmod(n, parts)
    lo = i * base + min(i, extra)
    return lo, lo + base + (1 if i < extra else 0)


class GpuEngine:
    """api.Screen seen through the few operations the sharded plan needs."""

    sharded_s2 = True

    def __init__(self, scr, torch):
        self.scr, self.torch = scr, torch

    def _wrap(self, ptr_n):
        ptr, n = ptr_n
        return self.torch.as_tensor(_DevMem(ptr, n), device=f"cuda:{self.scr.device}") if ptr and n else None

    def table(self):
        return self._wrap(self.scr.dev_count_table())

    def merge_into(self, byte_offset: int, other) -> None:
        self.scr.count_merge(other.data_ptr(), other.numel(), byte_offset // 4)

    def hit_bits(self, which: int):
        return self._wrap(self.scr.dev_hit_bits(which))

    def tile_bytes(self) -> int:
        return 128                                   # 1024 positions, one bit each

    def peak_filter(self):
        return self._wrap(self.scr.dev_peak_filter())

    def tile_new(self):
        return self._wrap(self.scr.dev_tile_new())

    def flagged(self):
        return self._wrap(self.scr.dev_flagged())

    def peak_table(self):
        return self._wrap(self.scr.dev_peak_table()).view(self.torch.int32)      # ids stay below 2^31 (max_peak)

    def loci(self):
        t = self._wrap(self.scr.dev_loci())
        return None if t is None else t.view(self.torch.int32)

    def sync(self):
        self.scr.sync()

    def __getattr__(self, name):                      # everything else is the Screen's own method
        return getattr(self.scr, name)


class Shard:
    """same_stream=True: the engine launches on torch's CURRENT stream (the caller did scr.set_stream(that stream)), so the
    engine's kernels and the collectives are ordered by the stream itself and a step needs no host synchronisation
    between them: a barrier is a one-element all-reduce in stream order.  Otherwise every hand-over is fenced on the host."""

    def __init__(self, scr, rank: int = 0, world: int = 1, dist=None, torch=None, engine=None, same_stream: bool = False,
                 force_nccl: bool = False):
        self.rank, self.world, self.dist, self.torch = rank, world, dist, torch
        self.eng = engine if engine is not None else GpuEngine(scr, torch)
        self.same_stream = same_stream and engine is None
        self.last_stage_ms = np.zeros(13)
        self.last_peaks = 0
        self.last_counts = {}
        self.last_wall_ms = {}
        self.p2p = False if force_nccl else self._open_peers()
        self._tiny = None
        self._spans = []

    def _open_peers(self) -> bool:
        """Maps every rank's count table into this process (CUDA IPC) so that the count exchange can run as one kernel
        over NVLink peer memory.  All ranks agree on the outcome; when any of them cannot (no IPC in this container, a
        non-GPU engine) everybody uses the NCCL form."""
        if self.world == 1 or not hasattr(self.eng, "count_table_ipc") or self.dist is None:
            return False
        ok, handles = 1, [None] * self.world
        try:
            mine = self.eng.count_table_ipc()
        except Exception:
            ok, mine = 0, b"\0" * 64
        self.dist.all_gather_object(handles, mine)
        if ok:
            try:
                self.eng.peers_open(self.rank, self.world, b"".join(handles))
            except Exception:
                ok = 0
        t = self.torch.tensor([ok], dtype=self.torch.int32, device=self._dev())
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))

    # ---- collectives on small host scalars
    def _all_gather_ints(self, vals):
        if self.world == 1:
            return [list(vals)]
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self._dev())
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    def _dev(self):
        tab = self.eng.table()
        return tab.device

    def _fence(self, t) -> None:
        """The engine's kernels run on its own stream: make the host wait for the collective (which torch
        orders after the current stream) before the next engine call touches the buffer."""
        if self.same_stream:
            return
        if t is not None and t.is_cuda:
            self.torch.cuda.current_stream(t.device).synchronize()

    def _engine_done(self) -> None:
        if not self.same_stream:
            self.eng.sync()

    def _barrier(self) -> None:
        if not self.same_stream:
            self.dist.barrier()
            return
        if self._tiny is None:
            self._tiny = self.torch.zeros(1, dtype=self.torch.int32, device=self._dev())
        self.dist.all_reduce(self._tiny)                       # in stream order: completes once every rank has reached it

    def _span(self, name):
        """Device time of an exchange, from CUDA events on the shared stream (read at the end of the step)."""
        if not self.same_stream:
            return None
        ev = (name, self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True))
        ev[1].record()
        self._spans.append(ev)
        return ev

    @staticmethod
    def _end(ev):
        if ev is not None:
            ev[2].record()

    # ---- exchanges
    def exchange_counts(self) -> None:
        """count := min(3, sum over ranks), on every rank."""
        if self.p2p:                                           # one kernel per rank over peer memory, between two barriers
            self._engine_done()
            self._barrier()                                    # every rank's S1 has finished: all tables are final
            self.eng.count_exchange_p2p()
            self._engine_done()
            self._barrier()                                    # every slice has been written into every table
            return
        tab = self.eng.table()
        n = tab.numel()
        w = self.world
        if n % 4:
            raise ValueError("count table is not a whole number of 32-bit words")
        self._engine_done()
        # slice j (whole words; the slices differ by at most one word when w does not divide the table) belongs to rank j
        spans = [tuple(4 * x for x in split_range(n // 4, w, j)) for j in range(w)]
        sizes = [hi - lo for lo, hi in spans]
        lo, mine_n = spans[self.rank][0], sizes[self.rank]
        recv = self.torch.empty(w * mine_n, dtype=tab.dtype, device=tab.device)
        self.dist.all_to_all_single(recv, tab, output_split_sizes=[mine_n] * w, input_split_sizes=sizes)   # recv[j] = rank j's slice `rank`
        self._fence(recv)
        for j in range(w):
            if j != self.rank and mine_n:
                self.eng.merge_into(lo, recv[j * mine_n:(j + 1) * mine_n])
        self._engine_done()
        if len(set(sizes)) == 1:
            mine = tab[lo:lo + mine_n].clone()
            self.dist.all_gather_into_tensor(tab, mine)
        else:                                                  # uneven slices: one broadcast per owner
            for j, (a, b) in enumerate(spans):
                if b > a:
                    self.dist.broadcast(tab[a:b], src=j)
        self._fence(tab)

    def tile_block(self, ntiles: int):
        """Equal blocks of tiles per rank (the last ones may be short or empty): rank r owns [r * B, (r + 1) * B) cut at ntiles."""
        B = -(-ntiles // self.world)
        lo = min(ntiles, self.rank * B)
        return B, lo, min(ntiles, lo + B)

    def _gather_blocks(self, t, ntiles: int, bytes_per_tile: int) -> None:
        """All-gather, in place, of the per-rank tile blocks of a tile-major device array (padded to world * B tiles)."""
        B, _, _ = self.tile_block(ntiles)
        n = B * bytes_per_tile
        if t.numel() < self.world * n:
            raise ValueError("tile-major array lacks the padding for equal tile blocks")
        self.dist.all_gather_into_tensor(t[: self.world * n], t[self.rank * n:(self.rank + 1) * n])
        self._fence(t)

    def exchange_hit_bits(self, ntiles: int, which=(0, 1)) -> None:
        self._engine_done()
        for w_ in which:
            self._gather_blocks(self.eng.hit_bits(w_), ntiles, self.eng.tile_bytes())

    # ---- the pass
    def screen(self, *, size1: int, sample_arg: float, seed: int, rand_skip: int, hit: float, match: float,
               max_peak: int, size2: Optional[int] = None, before_mate2=None, before_s2=None) -> bytes:
        """Reads must already be resident (reads_upload / reads_attach_device) and the index loaded -- or arrive through
        the hooks: before_mate2() runs after S1 of fq1 was launched and must make fq2 resident (size2 = its byte size,
        needed up front for the Q15 budget), before_s2() must make the index resident.  That is the stage order of the
        reference's main() (E:1426-1507), which lets host->device copies hide behind S1."""
        eng, w = self.eng, self.world
        ms = np.zeros(13)
        wall = {}
        t0 = time.perf_counter()
        self._spans = []
        deferred = hasattr(eng, "set_deferred")

        def lap(name, since):
            now = time.perf_counter()
            wall[name] = wall.get(name, 0.0) + 1000 * (now - since)
            return now

        eng.reset()
        if deferred:
            eng.set_deferred(True)
        nrec1 = eng.reads_records(0)
        t = lap("reset", t0)
        info = self._all_gather_ints([nrec1, eng.reads_seq_bases(0), size1, eng.reads_bytes(1) if size2 is None else size2])
        t = lap("gather_sizes", t)
        base = sum(r[0] for r in info[: self.rank])
        if sample_arg <= 1:
            ratio = 100 * sample_arg                                   # E:1392-1394
        else:
            ratio = 100 * sample_arg / float(2 * sum(r[1] for r in info))   # E:1258-1265 over the whole fq1
        size1_total = sum(r[2] for r in info)
        fq2_before = sum(r[3] for r in info[: self.rank])
        budget2 = size1_total - fq2_before if w > 1 else size1         # Q15 in shard-local byte offsets
        eng.set_ordinal_base(base)
        eng.set_sampling(ratio, seed, rand_skip)
        t = lap("set_sampling", t)
        ms[7] = 1000 * (time.perf_counter() - t0)
        n1 = eng.s1_count(0, size1_total)                              # fq1 records never start beyond size(fq1)
        if before_mate2 is not None:
            before_mate2()
            t = lap("s1", t)
            eng.set_sampling(ratio, seed, rand_skip)                   # fq2 may hold more records than fq1: cover them
        n2 = eng.s1_count(1, budget2) if budget2 >= 0 else 0
        t = lap("s1", t)
        if w > 1:
            ev = self._span("exchange_counts")
            self.exchange_counts()
            self._end(ev)
            t = lap("exchange_counts", t)
        if before_s2 is not None:
            before_s2()
            t = lap("index_upload", t)
        if w > 1 and eng.sharded_s2:
            nt = eng.s2_tiles()
            _, lo, hi = self.tile_block(nt)
            blk = eng.index_block() if hasattr(eng, "index_block") else None
            partial = blk is not None and (blk[2], blk[3]) != (0, nt)      # this rank holds one block of the image only (row e-S)
            if partial and (blk[2], blk[3]) != (lo, hi):
                raise ValueError("image block and tile block differ: build the index with set_image_block(rank, world)")
            eng.s2_gather(lo, hi)
            t = lap("s2_gather", t)
            ev = self._span("exchange_hit_bits")
            self.exchange_hit_bits(nt)
            self._end(ev)
            t = lap("exchange_hit_bits", t)
            eng.s2_mark(match)
            eng.s2_complete(lo, hi)
            t = lap("s2_complete", t)
            ev = self._span("exchange_single")
            self.exchange_hit_bits(nt, which=(0,))
            self._end(ev)
            t = lap("exchange_hit_bits", t)
            # windows / peak opening / registration by equal shares of the MARKED tiles (they cluster where the sample's genomes
            # are); ids need every tile's new-peak count
            wlo, whi = eng.s2_need_range(self.rank, w)
            eng.s2_windows(hit, match, wlo, whi)
            ev = self._span("exchange_tile_new")
            self._engine_done()
            tn = eng.tile_new().view(self.torch.int32)
            self.dist.all_reduce(tn)                                   # every tile is counted by exactly one rank
            self._fence(tn)
            self._end(ev)
            flagged_total = sum(r[0] for r in self._all_gather_ints([eng.s2_flagged_in_range()]))
            n_peaks = eng.s2_ids(max_peak, flagged_total)
            t = lap("s2_windows", t)
            if n_peaks > 0 and partial:
                # the hashes of a tile live on one rank only: flagged bits go to everybody, every rank registers its image block,
                # the tables are combined with MAX
                ev = self._span("exchange_flagged")
                self._engine_done()
                fl = eng.flagged()
                self.dist.all_reduce(fl, op=self.dist.ReduceOp.MAX)
                self._fence(fl)
                self._end(ev)
                eng.s2_register(lo, hi)
                ev = self._span("reduce_peak_table")
                self._engine_done()
                for buf in (eng.peak_table(), eng.loci()):
                    self.dist.all_reduce(buf, op=self.dist.ReduceOp.MAX)
                    self._fence(buf)
                self._end(ev)
            elif n_peaks > 0 and eng.s2_dense():
                # many registered k-mers: each rank registers its share, the tables are combined with MAX (= the last writer of
                # the sequential loop, since ids grow with position)
                eng.s2_register(wlo, whi)
                ev = self._span("reduce_peak_table")
                self._engine_done()
                for buf in (eng.peak_table(), eng.loci()):
                    self.dist.all_reduce(buf, op=self.dist.ReduceOp.MAX)
                    self._fence(buf)
                self._end(ev)
            elif n_peaks > 0:
                # few: everybody registers everything, from the combined flagged bits (a word is written by its owner and, for
                # the halo tile, identically by the next share's owner: MAX of equal or zero words)
                ev = self._span("exchange_flagged")
                self._engine_done()
                fl = eng.flagged()                                     # as bytes: unsigned, so MAX never drops a word with its top bit set
                self.dist.all_reduce(fl, op=self.dist.ReduceOp.MAX)
                self._fence(fl)
                self._end(ev)
                eng.s2_register(0, nt)
            t = lap("s2_register", t)
        else:
            n_peaks = eng.s2_peaks(hit, match, max_peak)
            t = lap("s2", t)
        n3 = eng.s3_pairs()
        t = lap("s3", t)
        if w > 1 and n_peaks > 0:
            self._engine_done()
            filt = eng.peak_filter()
            ev = self._span("reduce_filter")
            self.dist.all_reduce(filt, op=self.dist.ReduceOp.MAX)
            self._end(ev)
            self._fence(filt)
            t = lap("reduce_filter", t)
        text = eng.intervals()
        t = lap("intervals", t)
        if deferred:
            n1, n2, n3 = eng.deferred_counts()
            eng.set_deferred(False)
        self.last_wall_ms = wall
        st = eng.stage_ms()
        ms[:6] = st[:6]
        if len(st) >= 9:
            ms[8:11] = st[6:9]                                          # S1 in streams: hash, split, leaf-apply kernels
        if len(st) >= 12:
            ms[11:13] = st[10:12]                                       # S2 peak registration, S3 vote
        ms[6] = sum(a.elapsed_time(b) for _, a, b in self._spans) if self._spans else 0.0
        self.last_exchange_ms = {}
        for name, a, b in self._spans:
            self.last_exchange_ms[name] = self.last_exchange_ms.get(name, 0.0) + a.elapsed_time(b)
        self.last_stage_ms = ms
        self.last_peaks = int(n_peaks)
        self.last_counts = {"s1": (int(n1), int(n2)), "s3": int(n3), "ratio": ratio, "ordinal_base": base}
        return text
