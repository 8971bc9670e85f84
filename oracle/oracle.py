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

from pepper_b200.abi import (HostReads, PbReads, PbRegion, PbVariantParams, regions_array, variant_params,
                             WINDOW, FEATURES, ALLELE_STRIDE, POLISH_FEATURES)
from pepper_b200.synth import ReadBatch, RegionTable

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
    """kind in {'port', 'ref_variant', 'ref_polish'}"""
    if kind not in _libs:
        path = {"port": os.path.join(HERE, "liboracle_port.so"),
                "ref_variant": os.path.join(HERE, "_ref", "libref_variant.so"),
                "ref_polish": os.path.join(HERE, "_ref", "libref_polish.so")}[kind]
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
            L.port_variant_debug.restype = None
            L.port_variant_debug.argtypes = [C.c_void_p] * 5
        _libs[kind] = L
    return _libs[kind]


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_variant.so")) and \
        os.path.exists(os.path.join(HERE, "_ref", "libref_polish.so"))


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
