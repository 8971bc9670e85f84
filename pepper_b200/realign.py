"""Read -> reference realignment on the GPU: host-side mirror of
``ReadAligner(ref_start, ref_end, ref_seq).align_reads_to_reference(reads)``
(pepper/modules/src/local_reassembly/simple_aligner.cpp:58-106; caller pepper AlignmentSummarizer.py:159-177) for a batch
of regions.  Bit-identical to the SSW library the reference vendors (see pepper_b200/csrc/realign.cu)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from .abi import HostReads, PbReads, PbRegion, regions_array
from .synth import ReadBatch, RegionTable

ALIGNMENT_SAFE_BASES = 20            # AlingerOptions.ALIGNMENT_SAFE_BASES (pepper Options.py:25)


def _bind(L):
    if getattr(L, "_realign_bound", False):
        return
    L.pb_realigner_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.pb_realigner_destroy.argtypes = [C.c_void_p]
    L.pb_realign_device.argtypes = [C.c_void_p, C.POINTER(PbReads), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                    C.POINTER(PbReads), C.c_void_p]
    L.pb_realign_host.argtypes = [C.c_void_p, C.POINTER(PbReads), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]
    L.pb_realign_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L._realign_bound = True


def realign_regions(regions: RegionTable, genome: np.ndarray, contig_start: int = 0) -> RegionTable:
    """Region table whose reference strings are get_reference_sequence(chrom, region_start, region_end + 20)
    (AlignmentSummarizer.py:164-170), cut from an in-memory contig."""
    tab = regions.table.copy()
    refs, off = [], 0
    for r in range(tab.shape[0]):
        s = int(tab[r, 0]) - contig_start
        e = min(int(tab[r, 1]) + ALIGNMENT_SAFE_BASES - contig_start, genome.shape[0])
        refs.append(genome[s:e])
        tab[r, 4], tab[r, 5] = off, e - s
        off += e - s
    return RegionTable(tab, np.concatenate(refs) if refs else np.zeros(1, np.uint8))


class Realigner:
    def __init__(self, device: int = 0):
        _lib.require_gpu()
        self.L = _lib.lib()
        _bind(self.L)
        self.h = C.c_void_p()
        _lib.check(self.L.pb_realigner_create(C.byref(self.h), device), "pb_realigner_create")

    def close(self):
        if self.h:
            self.L.pb_realigner_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001
            pass

    def realign(self, reads: ReadBatch, regions: RegionTable, stream: int = 0) -> ReadBatch:
        """Host arrays in, a new ReadBatch out (sequence / qualities shared with the input)."""
        hr = HostReads(reads)
        regs, keep = regions_array(regions)
        ref = np.ascontiguousarray(regions.ref, dtype=np.uint8)
        n = reads.n_reads
        pos = np.zeros(n, dtype=np.int64)
        cigar_off = np.zeros(n + 1, dtype=np.int64)
        cap = int(reads.cigar_off[-1]) + reads.n_bases // 4 + 64
        while True:
            cigar = np.zeros(cap, dtype=np.uint32)
            nc = C.c_int64(0)
            rc = self.L.pb_realign_host(self.h, C.byref(hr.struct), C.cast(regs, C.c_void_p), regions.n_regions, ref.ctypes.data, ref.shape[0],
                                        pos.ctypes.data, cigar_off.ctypes.data, cigar.ctypes.data, cap, C.byref(nc), C.c_void_p(stream))
            if rc == -3:
                cap = int(nc.value) + 16
                continue
            _lib.check(rc, "pb_realign_host")
            break
        return ReadBatch(pos, reads.seq_off, cigar_off, reads.flags, reads.mapq, reads.seq, reads.qual, cigar[:int(nc.value)].copy())

    def realign_device(self, dreads, stream: int = 0):
        """`dreads`: DeviceReads / FetchedReads (reads, region table and reference strings in HBM).  Returns a pb_reads_t view
        (device pointers owned by the realigner) that replaces dreads.struct."""
        out = PbReads()
        _lib.check(self.L.pb_realign_device(self.h, C.byref(dreads.struct), dreads.d_regions, C.cast(dreads.h_regions, C.c_void_p),
                                            dreads.n_regions, dreads.d_ref, dreads.ref_bytes, C.byref(out), C.c_void_p(stream)),
                   "pb_realign_device")
        return out

    def stats(self) -> dict:
        a, b, m0, m1 = C.c_int64(0), C.c_int64(0), C.c_float(0), C.c_float(0)
        self.L.pb_realign_stats(self.h, C.byref(a), C.byref(b), C.byref(m0), C.byref(m1))
        return dict(aligned=int(a.value), realigned=int(b.value), sw_ms=float(m0.value), cigar_ms=float(m1.value))
