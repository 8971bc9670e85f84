"""Pins oracle/stitch.py against the UNMODIFIED reference `small_chunk_stitch` / `create_consensus_sequence`
(pepper/modules/python/Stitch.py) run in this container: the module is imported from /root/reference with an npz-backed
stand-in for h5py (h5py/libhdf5 are not installed) and `np.int = int` (alias removed from numpy >= 1.24).
    python tests/golden/make_golden_stitch.py"""
import os
import sys
import types
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from pepper_b200 import datastore as ds  # noqa: E402
from oracle import stitch as ostitch  # noqa: E402


class _Node:
    def __init__(self, store, prefix):
        self.s, self.p = store, prefix

    def keys(self):
        return self.s.keys(self.p)

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        path = (self.p + "/" + k).strip("/")
        if path in self.s.data:
            return _Leaf(self.s.data[path])
        return _Node(self.s, path)


class _Leaf:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, k):
        return self.a[k] if k != () else (self.a[()] if self.a.shape == () else self.a)


class _File(_Node):
    def __init__(self, name, mode="r"):
        super().__init__(ds._Store(name, mode="r", backend="npz"), "")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def close(self):
        pass


fake = types.ModuleType("h5py")
fake.File = _File
sys.modules["h5py"] = fake
np.int = int
sys.path.insert(0, "/root/reference")
from pepper.modules.python import Stitch as RefStitch  # noqa: E402

rng = np.random.default_rng(11)
# 3 regions tiled like ImageGenerationUI.py:269-272 with insert columns; 2-3 chunks of 1000 columns each, 50 overlap
regions = [(0, 1100), (900, 2100), (1900, 3000)]
images = []          # (region, chunk_id, position[1000], index[1000], bases[1000])
for r, (s, e) in enumerate(regions):
    cols = []
    for p in range(s, e + 1):
        cols.append((p, 0))
        for k in range(int(rng.random() < 0.25) * int(rng.integers(1, 4))):
            cols.append((p, k + 1))
    n = len(cols)
    start, end, cid = 0, min(n, 1000), 0
    while True:
        pos = np.full(1000, -1, np.int64); idx = np.full(1000, -1, np.int64)
        m = end - start
        pos[:m] = [c[0] for c in cols[start:end]]; idx[:m] = [c[1] for c in cols[start:end]]
        images.append((r, cid, pos, idx, rng.integers(0, 5, 1000).astype(np.uint8)))
        cid += 1
        if end == n:
            break
        start = end - 50
        end = min(n, start + 1000)

tmp = tempfile.mkdtemp()
fname = os.path.join(tmp, "pred.hdf")
with ds.PolishPredictionStore(fname, backend="npz") as st:
    for r, cid, pos, idx, b in images:
        st.write_prediction("ctg1", regions[r][0], regions[r][1], cid, pos, idx, b, np.zeros(1000, np.uint8))
keys = [(fname, "ctg1-%d-%d" % (s, e), s, e) for s, e in regions]
# the reference's per-thread worker, one call per region group as create_consensus_sequence does with threads=1
want_chunks = []
for grp in RefStitch.chunks([(fname, "ctg1", s, e) for s, e in sorted(regions)], max(2, int(len(regions) / 1) + 1)):
    first, last, seq = RefStitch.small_chunk_stitch("ctg1", grp)
    want_chunks.append((first, last, seq))
want = "".join(s for _, _, s in sorted(want_chunks))
got = ostitch.stitch(np.stack([i[4] for i in images]), np.stack([i[2] for i in images]), np.stack([i[3] for i in images]),
                     np.array([i[0] for i in images]), np.array([i[1] for i in images]), [r[0] for r in regions], [r[1] for r in regions])
assert got == want, (len(got), len(want))
np.savez_compressed(os.path.join(HERE, "stitch_seed11.npz"), bases=np.stack([i[4] for i in images]), position=np.stack([i[2] for i in images]),
                    index=np.stack([i[3] for i in images]), image_region=np.array([i[0] for i in images], np.int32),
                    chunk_id=np.array([i[1] for i in images], np.int32), region_starts=np.array([r[0] for r in regions]),
                    region_ends=np.array([r[1] for r in regions]), consensus=np.array(want))
print("stitch pinned: %d images, consensus length %d" % (len(images), len(want)))
