"""TEST INFRASTRUCTURE ONLY (oracle) — never imported by the product package.

ctypes front-end over
  * ``oracle/liboracle_port.so``  — plain-C restatement of the two encoders (port_encoders.c)
  * ``oracle/_ref/libref_*.so``   — the UNMODIFIED reference C++ encoders compiled from /root/reference
and the CPU PyTorch restatement of the two networks (nets.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from pepper_b200.abi import (HostReads, HostRecords, PbReads, PbRecords, PbRegion, PbVariantParams, regions_array,
                             variant_params, WINDOW, FEATURES, ALLELE_STRIDE, POLISH_FEATURES)
from pepper_b200.synth import ReadBatch, RecordBatch, RegionTable, NT16, CODE_OF, pack_codes

HERE = os.path.dirname(os.path.abspath(__file__))


def build(quiet: bool = True) -> None:
    """make -C oracle (compiles the port; compiles _ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def _load(path: str):
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_libs: dict[str, C.CDLL] = {}


def lib(kind: str) -> C.CDLL:
    """kind in {'port', 'ref_variant', 'ref_polish', 'ref_getreads', 'ref_realign'}"""
    if kind not in _libs:
        path = {"port": os.path.join(HERE, "liboracle_port.so"),
                "ref_variant": os.path.join(HERE, "_ref", "libref_variant.so"),
                "ref_polish": os.path.join(HERE, "_ref", "libref_polish.so"),
                "ref_getreads": os.path.join(HERE, "_ref", "libref_getreads.so"),
                "ref_realign": os.path.join(HERE, "_ref", "libref_realign.so")}[kind]
        L = _load(path)
        if kind in ("port", "ref_variant"):
            pre = "port" if kind == "port" else "ref"
            f = getattr(L, pre + "_variant_run")
            f.restype = C.c_int64
            f.argtypes = [C.POINTER(PbReads), C.POINTER(PbRegion), C.c_void_p, C.POINTER(PbVariantParams)]
            g = getattr(L, pre + "_variant_fetch")
            g.restype = None
            g.argtypes = [C.c_void_p] * 5
        if kind in ("port", "ref_polish"):
            pre = "port" if kind == "port" else "ref"
            f = getattr(L, pre + "_polish_run")
            f.restype = C.c_int64
            f.argtypes = [C.POINTER(PbReads), C.POINTER(PbRegion)]
            g = getattr(L, pre + "_polish_fetch")
            g.restype = None
            g.argtypes = [C.c_void_p] * 3
        if kind == "port":
            L.port_get_reads.restype = C.c_int64
            L.port_get_reads.argtypes = [C.POINTER(PbRecords), C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11
        if kind == "port":
            L.port_ssw_align.restype = C.c_int
            L.port_ssw_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
            L.port_realign.restype = C.c_int64
            L.port_realign.argtypes = [C.POINTER(PbReads), C.c_int64, C.c_int64, C.c_int64, C.c_char_p, C.c_int64] + [C.c_void_p] * 6
        if kind == "ref_realign":
            L.ref_realign_run.restype = None
            L.ref_realign_run.argtypes = [C.POINTER(PbReads), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p]
            L.ref_realign_fetch.restype = None
            L.ref_realign_fetch.argtypes = [C.c_void_p] * 5
            L.ref_ssw_align.restype = None
            L.ref_ssw_align.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_char_p, C.c_int32]
        if kind == "ref_getreads":
            L.ref_getreads_load.restype = None
            L.ref_getreads_load.argtypes = [C.POINTER(PbRecords)]
            L.ref_getreads_query.restype = None
            L.ref_getreads_query.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
            L.ref_getreads_fetch.restype = None
            L.ref_getreads_fetch.argtypes = [C.c_void_p] * 11
        if kind == "port":
            L.port_variant_debug.restype = None
            L.port_variant_debug.argtypes = [C.c_void_p] * 5
        _libs[kind] = L
    return _libs[kind]


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_variant.so")) and \
        os.path.exists(os.path.join(HERE, "_ref", "libref_polish.so"))


def have_ref_getreads() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_getreads.so"))


_loaded_records = [None]


def get_reads(records: RecordBatch, start: int, stop: int, include_supplementary: bool = False, min_mapq: int = 0,
              min_baseq: int = 0, impl: str = "port"):
    """BAM_handler.get_reads for one query.  Returns (ReadBatch, pos_end int64 [n], n_bad int64 [n]).
    impl='port': oracle/port_getreads.c;  impl='ref': the unmodified reference function (oracle/_ref/libref_getreads.so)."""
    hr = HostRecords(records)
    nb, nc, n = int(records.seq_off[-1]), int(records.cigar_off[-1]), records.n_records
    sizes = np.zeros(3, dtype=np.int64)
    if impl == "port":
        L = lib("port")
        pos, pos_end, n_bad = (np.zeros(n + 1, dtype=np.int64) for _ in range(3))
        seq_off, cigar_off = np.zeros(n + 2, dtype=np.int64), np.zeros(n + 2, dtype=np.int64)
        flags, mapq = np.zeros(n + 1, dtype=np.uint8), np.zeros(n + 1, dtype=np.uint8)
        seq, qual, cigar = np.zeros(nb // 2 + 2, dtype=np.uint8), np.zeros(nb + 1, dtype=np.uint8), np.zeros(nc + 1, dtype=np.uint32)
        L.port_get_reads(C.byref(hr.struct), start, stop, int(include_supplementary), min_mapq, min_baseq,
                         pos.ctypes.data, pos_end.ctypes.data, seq_off.ctypes.data, cigar_off.ctypes.data, flags.ctypes.data,
                         mapq.ctypes.data, seq.ctypes.data, qual.ctypes.data, cigar.ctypes.data, n_bad.ctypes.data, sizes.ctypes.data)
        m, mb, mc = (int(x) for x in sizes)
        return (ReadBatch(pos[:m].copy(), seq_off[:m + 1].copy(), cigar_off[:m + 1].copy(), flags[:m].copy(), mapq[:m].copy(),
                          seq[:(mb + 1) // 2].copy(), qual[:mb].copy(), cigar[:mc].copy()), pos_end[:m].copy(), n_bad[:m].copy())
    L = lib("ref_getreads")
    if _loaded_records[0] is not records:
        L.ref_getreads_load(C.byref(hr.struct))
        _loaded_records[0] = records
    L.ref_getreads_query(start, stop, int(include_supplementary), min_mapq, min_baseq, sizes.ctypes.data)
    m, mb, mc = (int(x) for x in sizes)
    pos, pos_end, n_bad = (np.zeros(m + 1, dtype=np.int64) for _ in range(3))
    seq_off, cigar_off = np.zeros(m + 1, dtype=np.int64), np.zeros(m + 1, dtype=np.int64)
    flags, mapq = np.zeros(m + 1, dtype=np.uint8), np.zeros(m + 1, dtype=np.uint8)
    ascii_ = np.zeros(mb + 1, dtype=np.uint8)
    qual = np.zeros(mb + 1, dtype=np.uint8)
    op, ln = np.zeros(mc + 1, dtype=np.int32), np.zeros(mc + 1, dtype=np.int32)
    L.ref_getreads_fetch(pos.ctypes.data, pos_end.ctypes.data, seq_off.ctypes.data, cigar_off.ctypes.data, flags.ctypes.data,
                         mapq.ctypes.data, ascii_.ctypes.data, qual.ctypes.data, op.ctypes.data, ln.ctypes.data, n_bad.ctypes.data)
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, code in CODE_OF.items():
        lut[ord(ch)] = code
    codes = lut[ascii_[:mb]]
    assert (codes != 255).all()
    cigar = ((ln[:mc].astype(np.uint32) << 4) | op[:mc].astype(np.uint32)).astype(np.uint32)
    return (ReadBatch(pos[:m].copy(), seq_off.copy(), cigar_off.copy(), flags[:m].copy(), mapq[:m].copy(), pack_codes(codes),
                      qual[:mb].copy(), cigar), pos_end[:m].copy(), n_bad[:m].copy())


def variant_encode(reads: ReadBatch, regions: RegionTable, params: dict, impl: str = "port",
                   debug: bool = False):
    """Run the variant encoder region by region (as the reference does).  Returns a dict of
    images int32 [N,33,26], positions int64, depths, freqs int32, keys list[str], region_of int32,
    and (debug, port only) per-region (matrix, cov, snp, ins, del)."""
    L = lib("port" if impl == "port" else "ref_variant")
    pre = "port" if impl == "port" else "ref"
    run, fetch = getattr(L, pre + "_variant_run"), getattr(L, pre + "_variant_fetch")
    hr = HostReads(reads)
    regs, _keep = regions_array(regions)
    ref = np.ascontiguousarray(regions.ref, dtype=np.uint8)
    p = variant_params(**params)
    out = dict(images=[], positions=[], depths=[], freqs=[], keys=[], region_of=[], debug=[])
    for r in range(regions.n_regions):
        n = run(C.byref(hr.struct), C.byref(regs[r]), ref.ctypes.data, C.byref(p))
        img = np.zeros((n, WINDOW, FEATURES), dtype=np.int32)
        pos = np.zeros(n, dtype=np.int64)
        dep = np.zeros(n, dtype=np.int32)
        frq = np.zeros(n, dtype=np.int32)
        keys = np.zeros((n, ALLELE_STRIDE), dtype=np.uint8)
        if n:
            fetch(img.ctypes.data, pos.ctypes.data, dep.ctypes.data, frq.ctypes.data, keys.ctypes.data)
        out["images"].append(img)
        out["positions"].append(pos)
        out["depths"].append(dep)
        out["freqs"].append(frq)
        out["keys"].extend(bytes(k).split(b"\0", 1)[0].decode() for k in keys)
        out["region_of"].append(np.full(n, r, dtype=np.int32))
        if debug and impl == "port":
            L1 = int(regions.table[r, 1] - regions.table[r, 0] + 1)
            m = np.zeros((L1, FEATURES), dtype=np.int32)
            v = [np.zeros(L1, dtype=np.int32) for _ in range(4)]
            L.port_variant_debug(m.ctypes.data, *[x.ctypes.data for x in v])
            out["debug"].append((m, *v))
    for k in ("images", "positions", "depths", "freqs", "region_of"):
        out[k] = np.concatenate(out[k]) if out[k] else np.zeros(0)
    return out


def polish_encode(reads: ReadBatch, regions: RegionTable, impl: str = "port"):
    """Run the polish encoder region by region.  Returns image uint8 [cols,10], pos int64, idx int32,
    col_off int64 [n_regions+1]."""
    L = lib("port" if impl == "port" else "ref_polish")
    pre = "port" if impl == "port" else "ref"
    run, fetch = getattr(L, pre + "_polish_run"), getattr(L, pre + "_polish_fetch")
    hr = HostReads(reads)
    regs, _keep = regions_array(regions)
    imgs, poss, idxs, off = [], [], [], [0]
    for r in range(regions.n_regions):
        n = run(C.byref(hr.struct), C.byref(regs[r]))
        img = np.zeros((n, POLISH_FEATURES), dtype=np.uint8)
        pos = np.zeros(n, dtype=np.int64)
        idx = np.zeros(n, dtype=np.int32)
        if n:
            fetch(img.ctypes.data, pos.ctypes.data, idx.ctypes.data)
        if impl == "port":
            pos += int(regions.table[r, 0])
        imgs.append(img)
        poss.append(pos)
        idxs.append(idx)
        off.append(off[-1] + n)
    return dict(image=np.concatenate(imgs), pos=np.concatenate(poss), idx=np.concatenate(idxs),
                col_off=np.array(off, dtype=np.int64))


def images_to_int8(images_i32: np.ndarray) -> np.ndarray:
    """DataStore.write_summary stores int8 (pepper_variant DataStore.py:68): values wrap mod 256."""
    return images_i32.astype(np.int64).astype(np.uint8).view(np.int8) if images_i32.size else \
        np.zeros(images_i32.shape, dtype=np.int8)


def have_ref_realign() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_realign.so"))


_OPCH = "MIDNSHP=X"


def ssw_align(query: str, ref: str, impl: str = "port"):
    """LibSSWPairwiseAligner.set_reference(ref); .align(query) (simple_aligner.cpp:12-29).
    Returns (score, ref_begin, ref_end, query_begin, query_end, mismatches, cigar_string)."""
    out = np.zeros(6, dtype=np.int32)
    if impl == "port":
        cap = 2 * len(query) + len(ref) + 16
        cig = np.zeros(cap, dtype=np.uint32)
        n = lib("port").port_ssw_align(query.encode(), len(query), ref.encode(), len(ref), out.ctypes.data, cig.ctypes.data, cap)
        s = "".join(f"{int(c >> 4)}{_OPCH[int(c & 15)]}" for c in cig[:n])
        return tuple(int(x) for x in out) + (s,)
    buf = C.create_string_buffer(4 * (len(query) + len(ref)) + 64)
    lib("ref_realign").ref_ssw_align(query.encode(), ref.encode(), len(ref), out.ctypes.data, buf, len(buf))
    return tuple(int(x) for x in out) + (buf.value.decode(),)


def realign(reads: ReadBatch, rb: int, re_: int, region_start: int, region_end: int, ref_seq: str, impl: str = "port"):
    """ReadAligner(region_start, region_end, ref_seq).align_reads_to_reference(reads[rb:re]).
    Returns (pos, pos_end, cigar_off, cigar uint32) of the output reads (sequence / qualities are unchanged)."""
    hr = HostReads(reads)
    if impl == "port":
        n_in = re_ - rb
        nb = int(reads.seq_off[re_] - reads.seq_off[rb])
        cap = 2 * nb + (len(ref_seq) + 16) * n_in + int(reads.cigar_off[-1]) + 64
        pos, pos_end, kept = (np.zeros(n_in + 1, dtype=np.int64) for _ in range(3))
        cigar_off = np.zeros(n_in + 2, dtype=np.int64)
        cigar = np.zeros(cap, dtype=np.uint32)
        score = np.zeros(n_in + 1, dtype=np.int32)
        n = lib("port").port_realign(C.byref(hr.struct), rb, re_, region_start, ref_seq.encode(), len(ref_seq), pos.ctypes.data,
                                     pos_end.ctypes.data, cigar_off.ctypes.data, cigar.ctypes.data, kept.ctypes.data, score.ctypes.data)
        n = int(n)
        return pos[:n].copy(), pos_end[:n].copy(), cigar_off[:n + 1].copy(), cigar[:int(cigar_off[n])].copy()
    L = lib("ref_realign")
    sizes = np.zeros(2, dtype=np.int64)
    L.ref_realign_run(C.byref(hr.struct), rb, re_, region_start, region_end, ref_seq.encode(), len(ref_seq), sizes.ctypes.data)
    n, nc = int(sizes[0]), int(sizes[1])
    pos, pos_end = np.zeros(n + 1, dtype=np.int64), np.zeros(n + 1, dtype=np.int64)
    cigar_off = np.zeros(n + 1, dtype=np.int64)
    op, ln = np.zeros(nc + 1, dtype=np.int32), np.zeros(nc + 1, dtype=np.int32)
    L.ref_realign_fetch(pos.ctypes.data, pos_end.ctypes.data, cigar_off.ctypes.data, op.ctypes.data, ln.ctypes.data)
    return pos[:n].copy(), pos_end[:n].copy(), cigar_off.copy(), ((ln[:nc].astype(np.uint32) << 4) | op[:nc].astype(np.uint32))
