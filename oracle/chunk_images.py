"""TEST INFRASTRUCTURE ONLY (oracle).  Restatement of the polish chunking step
(pepper/modules/python/AlignmentSummarizer.py:19-56, `AlignmentSummarizer.chunk_images`) on the column arrays of one
encoded region.  Pinned against the UNMODIFIED reference function by tests/golden/make_golden_chunks.py (the reference
module is imported from /root/reference with a stand-in for its compiled `pepper.build.PEPPER` import)."""
from __future__ import annotations

import numpy as np

SEQ_LENGTH = 1000        # Options.py: ImageSizeOptions.SEQ_LENGTH
SEQ_OVERLAP = 50         # ImageSizeOptions.SEQ_OVERLAP
IMAGE_HEIGHT = 10        # ImageSizeOptions.IMAGE_HEIGHT


def chunk_region(image, pos, idx, chunk_size: int = SEQ_LENGTH, chunk_overlap: int = SEQ_OVERLAP):
    """image uint8 [n,10], pos int64 [n], idx [n] of ONE region -> (images [k,1000,10], positions [k,1000,2], chunk_ids [k])."""
    n = int(pos.shape[0])
    chunk_start, chunk_id = 0, 0                                   # :20-21
    chunk_end = min(n, chunk_size)                                 # :22
    images, positions, chunk_ids = [], [], []
    while True:                                                    # :28
        img = np.zeros((chunk_size, IMAGE_HEIGHT), dtype=np.uint8)              # padding rows are [0.0] * 10  (:38-40)
        p = np.full((chunk_size, 2), -1, dtype=np.int64)                        # padding positions are (-1, -1)  (:39)
        m = chunk_end - chunk_start
        img[:m] = image[chunk_start:chunk_end]                     # :29
        p[:m, 0] = pos[chunk_start:chunk_end]                      # :30
        p[:m, 1] = idx[chunk_start:chunk_end]
        images.append(img); positions.append(p); chunk_ids.append(chunk_id)     # :44-47
        chunk_id += 1                                              # :48
        if chunk_end == n:                                         # :50
            break
        chunk_start = chunk_end - chunk_overlap                    # :53
        chunk_end = min(n, chunk_start + chunk_size)               # :54
    return np.stack(images), np.stack(positions), np.array(chunk_ids, dtype=np.int32)


def chunk_images(image, pos, idx, col_off):
    """All regions of a batch (columns of region r = [col_off[r], col_off[r+1])), in region order, as the reference's caller
    loops over regions (AlignmentSummarizer.py:334-345).  Returns images, position [n,1000], index [n,1000], chunk_id, region."""
    imgs, poss, cids, regs = [], [], [], []
    for r in range(len(col_off) - 1):
        a, b = int(col_off[r]), int(col_off[r + 1])
        i, p, c = chunk_region(image[a:b], pos[a:b], idx[a:b])
        imgs.append(i); poss.append(p); cids.append(c); regs.append(np.full(c.shape[0], r, dtype=np.int32))
    if not imgs:
        return (np.zeros((0, SEQ_LENGTH, IMAGE_HEIGHT), np.uint8), np.zeros((0, SEQ_LENGTH), np.int64), np.zeros((0, SEQ_LENGTH), np.int64),
                np.zeros(0, np.int32), np.zeros(0, np.int32))
    P = np.concatenate(poss)
    return np.concatenate(imgs), P[:, :, 0].copy(), P[:, :, 1].copy(), np.concatenate(cids), np.concatenate(regs)
