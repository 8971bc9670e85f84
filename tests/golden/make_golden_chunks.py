"""Pins oracle/chunk_images.py against the UNMODIFIED reference `AlignmentSummarizer.chunk_images`
(pepper/modules/python/AlignmentSummarizer.py:19-56) run in this container.  The reference module imports its compiled
extension (`from pepper.build import PEPPER`), which cannot be built here (htslib download); chunk_images itself is pure
Python, so an empty stand-in module is registered under that name before the import.
    python tests/golden/make_golden_chunks.py"""
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import chunk_images as och  # noqa: E402

sys.path.insert(0, "/root/reference")
import pepper  # noqa: E402  (namespace package rooted at /root/reference/pepper)
build = types.ModuleType("pepper.build")
build.PEPPER = types.ModuleType("pepper.build.PEPPER")
sys.modules["pepper.build"] = build
sys.modules["pepper.build.PEPPER"] = build.PEPPER
from pepper.modules.python.AlignmentSummarizer import AlignmentSummarizer  # noqa: E402
from pepper.modules.python.Options import ImageSizeOptions  # noqa: E402


class Summary:          # the attributes chunk_images reads from the pybind SummaryGenerator object
    def __init__(self, image, genomic_pos):
        self.image, self.genomic_pos = image, genomic_pos


rng = np.random.default_rng(13)
# column counts per region: below / at / just above one chunk, exactly two chunks' worth (1000 + 950), ragged multi-chunk, tiny
counts = [700, 1000, 1001, 1950, 1951, 2417, 1, 50, 51, 3850]
image, pos, idx, col_off = [], [], [], [0]
want_img, want_pos, want_cid, want_reg = [], [], [], []
for r, n in enumerate(counts):
    img = rng.integers(0, 255, size=(n, 10)).astype(np.uint8)
    p = np.sort(rng.integers(1000 * r, 1000 * r + max(2, n // 2), size=n)).astype(np.int64)
    k = np.zeros(n, np.int64)
    for i in range(1, n):
        k[i] = k[i - 1] + 1 if p[i] == p[i - 1] else 0
    s = Summary([list(map(int, row)) for row in img], [(int(a), int(b)) for a, b in zip(p, k)])
    images, labels, positions, chunk_ids = AlignmentSummarizer.chunk_images(s, ImageSizeOptions.SEQ_LENGTH, ImageSizeOptions.SEQ_OVERLAP)
    want_img.append(np.array(images, dtype=np.float64).astype(np.uint8))
    want_pos.append(np.array(positions, dtype=np.int64))
    want_cid.append(np.array(chunk_ids, dtype=np.int32))
    want_reg.append(np.full(len(chunk_ids), r, np.int32))
    assert all(len(l) == 1000 and not any(l) for l in labels)
    image.append(img); pos.append(p); idx.append(k); col_off.append(col_off[-1] + n)
image, pos, idx = np.concatenate(image), np.concatenate(pos), np.concatenate(idx)
W = np.concatenate(want_pos)
gi, gp, gk, gc, gr = och.chunk_images(image, pos, idx, np.array(col_off))
assert np.array_equal(gi, np.concatenate(want_img)) and np.array_equal(gp, W[:, :, 0]) and np.array_equal(gk, W[:, :, 1])
assert np.array_equal(gc, np.concatenate(want_cid)) and np.array_equal(gr, np.concatenate(want_reg))
# the fixture keeps the inputs and the reference's chunk table (start column, valid columns) + a checksum of the contents
starts, nvalid = [], []
for r, n in enumerate(counts):
    for c in range(len(want_cid[r])):
        v = int((want_pos[r][c][:, 0] >= 0).sum())
        nvalid.append(v)
        starts.append(col_off[r] + int(np.flatnonzero((pos[col_off[r]:col_off[r + 1]] == want_pos[r][c][0, 0]) & (idx[col_off[r]:col_off[r + 1]] == want_pos[r][c][0, 1]))[0]))
np.savez_compressed(os.path.join(HERE, "chunks_seed13.npz"), image=image, pos=pos, idx=idx, col_off=np.array(col_off), chunk_id=gc, region=gr,
                    start=np.array(starts), nvalid=np.array(nvalid), position=gp, index=gk,
                    image_sum=gi.astype(np.int64).sum(axis=(1, 2)))
print("chunk_images pinned: %d regions -> %d chunks" % (len(counts), gc.shape[0]))
