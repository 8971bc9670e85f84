"""Region read fetch + trim on the GPU: the host-side mirror of
``BAM_handler.get_reads(chromosome, start, stop, include_supplementary, min_mapq, min_baseq)``
(pepper/modules/src/dataio/bam_handler.cpp:115-451, pybind_api.h) for a batch of region queries, plus the reservoir
down-sampling the callers apply to its result (pepper_variant AlignmentSummarizer.py:109-125, 191-208;
pepper AlignmentSummarizer.py:313-325).

The records of one contig live in HBM (`DeviceRecords`); `ReadTrimmer.get_reads` returns a pb_reads_t view in HBM that
the encoders consume directly, and per-query read ranges for the region table.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from .abi import (HostRecords, PbRecords, PbReads, PbGetReadsOptions, PbInterval, intervals_array)
from .synth import ReadBatch, RecordBatch

RANDOM_SEED = 2719747673            # AlingerOptions.RANDOM_SEED (Options.py:99 / :29)


def _bind(L):
    if getattr(L, "_reads_bound", False):
        return
    L.pb_read_trimmer_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.pb_read_trimmer_destroy.argtypes = [C.c_void_p]
    L.pb_get_reads_plan_device.argtypes = [C.c_void_p, C.POINTER(PbRecords), C.c_void_p, C.c_int64,
                                           C.POINTER(PbGetReadsOptions), C.c_void_p, C.c_void_p]
    L.pb_get_reads_plan_host.argtypes = L.pb_get_reads_plan_device.argtypes
    L.pb_get_reads_emit_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PbReads), C.c_void_p, C.c_void_p,
                                           C.c_void_p]
    L.pb_get_reads_sizes.argtypes = [C.c_void_p, C.c_void_p]
    L.pb_get_reads_fetch.argtypes = [C.c_void_p] * 10
    L._reads_bound = True


def reservoir_select(total_reads: int, total_allowed: int, seed: int = RANDOM_SEED) -> np.ndarray | None:
    """Indices (into the get_reads result, in the order the reference's `sample` list ends up) kept by the reservoir
    sampler of AlignmentSummarizer.py:113-125; None when nothing is dropped.  Same RandomState stream as the reference."""
    if total_reads <= total_allowed:
        return None
    rng = np.random.RandomState(seed)
    sample = list(range(total_allowed))
    for i in range(total_allowed, total_reads):
        j = rng.randint(0, i + 1)
        if j < total_allowed:
            sample[j] = i
    return np.asarray(sample, dtype=np.int32)


class DeviceRecords:
    """A RecordBatch resident in HBM (torch is only the allocator)."""

    def __init__(self, rec: RecordBatch, device: int = 0):
        import torch
        dev = torch.device("cuda", device)
        hr = HostRecords(rec)
        self._keep = []

        def up(a):
            t = torch.from_numpy(a).to(dev)
            self._keep.append(t)
            return t.data_ptr()
        self.struct = PbRecords(rec.n_records, up(hr.pos), up(hr.seq_off), up(hr.cigar_off), up(hr.flag.view(np.int16)), up(hr.mapq),
                                up(hr.seq), up(hr.qual), up(hr.cigar.view(np.int32)))
        self.nbytes = hr.nbytes


class TrimmedReads:
    """Result of a batched get_reads: a pb_reads_t whose pointers are device pointers owned by the trimmer (valid until
    its next call), and the read range of each query."""

    def __init__(self, struct: PbReads, read_begin: np.ndarray, read_end: np.ndarray, total_reads: np.ndarray, trimmer):
        self.struct = struct
        self.read_begin = read_begin
        self.read_end = read_end
        self.total_reads = total_reads      # len(get_reads(...)) per query, before down-sampling
        self._trimmer = trimmer

    def to_host(self) -> ReadBatch:
        return self._trimmer.fetch()


class ReadTrimmer:
    def __init__(self, device: int = 0):
        _lib.require_gpu()
        self.L = _lib.lib()
        _bind(self.L)
        self.h = C.c_void_p()
        _lib.check(self.L.pb_read_trimmer_create(C.byref(self.h), device), "pb_read_trimmer_create")
        self.device = device

    def close(self):
        if self.h:
            self.L.pb_read_trimmer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass

    def get_reads(self, records, intervals, include_supplementary: bool = False, min_mapq: int = 0, min_baseq: int = 0,
                  max_reads: int | None = None, downsample_rate: float = 1.0, stream: int = 0) -> TrimmedReads:
        """`records`: DeviceRecords (HBM), RecordBatch (host, uploaded) or a BamReader.fetch view (host, uploaded).  `intervals`: [(start, stop), ...] exactly as the
        reference passes them to get_reads.  `max_reads` / `downsample_rate`: the caller-side reservoir sample
        (total_allowed = int(min(max_reads, downsample_rate * total_reads)), AlignmentSummarizer.py:110)."""
        iv, _keep = intervals_array(intervals)
        n = len(iv)
        opt = PbGetReadsOptions(int(bool(include_supplementary)), int(min_mapq), int(min_baseq), 0)
        counts = np.zeros(n, dtype=np.int64)
        if isinstance(records, RecordBatch) or getattr(records, "on_host", False):
            hr = HostRecords(records) if isinstance(records, RecordBatch) else records
            rc = self.L.pb_get_reads_plan_host(self.h, C.byref(hr.struct), C.cast(iv, C.c_void_p), n, C.byref(opt),
                                               counts.ctypes.data, C.c_void_p(stream))
        else:
            rc = self.L.pb_get_reads_plan_device(self.h, C.byref(records.struct), C.cast(iv, C.c_void_p), n, C.byref(opt),
                                                 counts.ctypes.data, C.c_void_p(stream))
        _lib.check(rc, "pb_get_reads_plan")
        sel_off = sel = None
        if max_reads is not None:
            picks, any_cut = [], False
            for i in range(n):
                total = int(counts[i])
                allowed = int(min(max_reads, downsample_rate * total))
                s = reservoir_select(total, allowed)
                any_cut |= s is not None
                picks.append(np.arange(total, dtype=np.int32) if s is None else s)
            if any_cut:
                sel_off = np.zeros(n + 1, dtype=np.int64)
                np.cumsum([p.shape[0] for p in picks], out=sel_off[1:])
                sel = np.ascontiguousarray(np.concatenate(picks) if picks else np.zeros(0, np.int32), dtype=np.int32)
        out = PbReads()
        rb, re_ = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        _lib.check(self.L.pb_get_reads_emit_device(self.h, sel_off.ctypes.data if sel_off is not None else None,
                                                   sel.ctypes.data if sel is not None else None, C.byref(out),
                                                   rb.ctypes.data, re_.ctypes.data, C.c_void_p(stream)), "pb_get_reads_emit_device")
        return TrimmedReads(out, rb, re_, counts, self)

    def fetch(self, stream: int = 0) -> ReadBatch:
        """Copy the last result to the host (parity tests, the one-region-per-call mirror classes)."""
        sizes = np.zeros(3, dtype=np.int64)
        _lib.check(self.L.pb_get_reads_sizes(self.h, sizes.ctypes.data), "pb_get_reads_sizes")
        n, nb, nc = (int(x) for x in sizes)
        pos = np.zeros(n, dtype=np.int64)
        seq_off = np.zeros(n + 1, dtype=np.int64)
        cigar_off = np.zeros(n + 1, dtype=np.int64)
        flags = np.zeros(n, dtype=np.uint8)
        mapq = np.zeros(n, dtype=np.uint8)
        seq = np.zeros((nb + 1) // 2, dtype=np.uint8)
        qual = np.zeros(nb, dtype=np.uint8)
        cigar = np.zeros(nc, dtype=np.uint32)
        _lib.check(self.L.pb_get_reads_fetch(self.h, pos.ctypes.data, seq_off.ctypes.data, cigar_off.ctypes.data, flags.ctypes.data,
                                             mapq.ctypes.data, seq.ctypes.data, qual.ctypes.data, cigar.ctypes.data,
                                             C.c_void_p(stream)), "pb_get_reads_fetch")
        return ReadBatch(pos, seq_off, cigar_off, flags, mapq, seq, qual, cigar)
