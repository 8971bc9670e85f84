"""TEST INFRASTRUCTURE ONLY (oracle).  Restatement of the polish stitch step
(pepper/modules/python/Stitch.py:36-128: small_chunk_stitch + create_consensus_sequence) on in-memory prediction
arrays.  Pinned against the UNMODIFIED reference function by tests/golden/make_golden_stitch.py (the reference module is
imported from /root/reference with an npz-backed stand-in for h5py)."""
from __future__ import annotations

from collections import defaultdict
import numpy as np

LABEL_DECODER = {1: "A", 2: "C", 3: "G", 4: "T", 0: ""}          # Stitch.py:13
BUFFER = 200                                                       # 2 * MIN_IMAGE_OVERLAP, Stitch.py:42


def stitch(bases, position, index, image_region, chunk_id, region_starts, region_ends):
    """bases uint8 [n,1000], position int64 [n,1000], index [n,1000], image_region [n], chunk_id [n];
    region r spans [region_starts[r], region_ends[r]].  Returns the consensus string of the contig."""
    order = sorted(range(len(region_starts)), key=lambda r: (int(region_starts[r]), int(region_ends[r])))   # Stitch.py:104
    by_region = defaultdict(list)
    for i in range(len(image_region)):
        by_region[int(image_region[i])].append(i)
    table = {}
    for r in order:
        st = int(region_starts[r])
        imgs = sorted(by_region.get(r, []), key=lambda i: str(int(chunk_id[i])))      # sorted(smaller_chunks): string order
        for i in imgs:
            for pos, idx, b in zip(position[i].tolist(), index[i].tolist(), bases[i].tolist()):
                if st > 0 and pos <= st + BUFFER:                                    # Stitch.py:66
                    continue
                if idx < 0 or pos < 0:                                               # :69
                    continue
                table[(pos, idx)] = b                                               # last writer wins (:72)
    keys = sorted(table.keys())
    return "".join(LABEL_DECODER[int(table[k])] for k in keys)
