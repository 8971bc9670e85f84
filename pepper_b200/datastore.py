"""Image / prediction stores with the reference's HDF5 layouts.

    variant images       pepper_variant/modules/python/DataStore.py:54-71         VariantImageStore.write_summary
    variant predictions  pepper_variant/modules/python/DataStorePredict.py:49-66  VariantPredictionStore.write_prediction
    polish images        pepper/modules/python/DataStore.py:53-67                 PolishImageStore.write_summary
    polish predictions   pepper/modules/python/DataStorePredict.py:49-76          PolishPredictionStore.write_prediction

Group / dataset names, dtypes and shapes are the reference's.  h5py / libhdf5 are not part of the build image (nor of the GPU
box), so the backend is chosen at run time: h5py when importable, otherwise a single-file ``.npz`` container whose keys are the
HDF5 paths.  String datasets follow the reference exactly: `candidates` is variable-length `str` (h5py.special_dtype(vlen=str),
DataStore.py:60 / DataStorePredict.py:60 — its readers parse ``str(candidates[i])``, CandidateFinder.py:374-381), the polish
`contig` is a `str` scalar (pepper DataStore.py:62); in the npz container those are unicode arrays, which read back as `str`
objects just like h5py 2.10 (requirements.txt:1) returned them.  tests/test_datastore_reference_readers.py runs the
reference's OWN reader code (both dataloader_predict.py, CandidateFinder.small_chunk_stitch, Stitch.small_chunk_stitch) over
stores written here, through an h5py stand-in backed by the npz container.
Deviations forced by modern numpy, all value-preserving: ``np.float`` -> ``np.float64`` (removed alias,
DataStorePredict.py:62); int8 images are passed as int8 already (the reference casts a list of Python ints).
"""
from __future__ import annotations

import os
import numpy as np

try:                                     # pragma: no cover - not installed in the build image
    import h5py                          # type: ignore
    HAVE_H5PY = True
except Exception:                        # noqa: BLE001
    h5py = None
    HAVE_H5PY = False


class _Store:
    """Minimal h5py-like writer: create_dataset(path, data) / attrs-free / close()."""

    def __init__(self, filename: str, mode: str = "w", backend: str | None = None):
        self.filename = filename
        self.backend = backend or ("h5py" if HAVE_H5PY else "npz")
        if self.backend == "h5py":
            self.f = h5py.File(filename, mode)
        else:
            self.data = {}
            if mode != "w" and os.path.exists(self._npz_name()):
                with np.load(self._npz_name(), allow_pickle=False) as z:
                    self.data = {k: z[k] for k in z.files}

    def _npz_name(self):
        return self.filename if self.filename.endswith(".npz") else self.filename + ".npz"

    def put(self, path: str, data):
        if self.backend == "h5py":
            if path in self.f:
                del self.f[path]
            self.f.create_dataset(path, data=data)
        else:
            self.data[path] = np.asarray(data)

    def put_vlen_str(self, path: str, rows):
        """`rows`: list of lists of str (one list per candidate) -> variable-length string dataset [n][len(row)]."""
        if self.backend == "h5py":
            if path in self.f:
                del self.f[path]
            self.f.create_dataset(path, data=np.array(rows, dtype=h5py.special_dtype(vlen=str)))      # DataStore.py:60,67
        else:
            self.data[path] = np.array(rows, dtype="U")          # reads back as str objects, like a vlen-str dataset under h5py 2.10

    def put_str(self, path: str, value: str):
        if self.backend == "h5py":
            if path in self.f:
                del self.f[path]
            self.f[path] = value                                  # pepper DataStore.py:62 assigns the Python str
        else:
            self.data[path] = np.array(value, dtype="U")

    def keys(self, prefix: str):
        if self.backend == "h5py":
            return list(self.f[prefix].keys()) if prefix in self.f else []
        pre = prefix.rstrip("/") + "/"
        return sorted({k[len(pre):].split("/", 1)[0] for k in self.data if k.startswith(pre)})

    def get(self, path: str):
        return self.f[path][()] if self.backend == "h5py" else self.data[path]

    def close(self):
        if self.backend == "h5py":
            self.f.close()
        else:
            np.savez(self._npz_name(), **self.data)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class VariantImageStore(_Store):
    """summaries/<region_name>/{contigs, positions, depths, candidates, candidate_frequency, images}"""

    def write_summary(self, region_name: str, contig: str, positions, depths, keys, freqs, images):
        n = len(positions)
        g = f"summaries/{region_name}/"
        self.put(g + "contigs", np.array([contig.encode()] * n, dtype="S"))
        self.put(g + "positions", np.asarray(positions, dtype=np.int32))                 # DataStore.py:64
        self.put(g + "depths", np.asarray(depths, dtype=np.uint8))
        self.put_vlen_str(g + "candidates", [[str(k)] for k in keys])                   # DataStore.py:60,67: vlen str [n][1]
        self.put(g + "candidate_frequency", np.asarray(freqs, dtype=np.uint8).reshape(n, 1))
        self.put(g + "images", np.asarray(images, dtype=np.int8).reshape(n, 33, 26))     # DataStore.py:68


class VariantPredictionStore(_Store):
    """predictions/batch_<n>/{contigs, positions, depths, candidates, candidate_frequency, base_prediction}"""

    def write_prediction(self, batch: int, contigs, positions, depths, keys, freqs, probs):
        n = len(positions)
        g = f"predictions/batch_{batch}/"
        self.put(g + "contigs", np.array([c.encode() if isinstance(c, str) else c for c in contigs], dtype="S"))
        self.put(g + "positions", np.asarray(positions, dtype=np.int32))
        self.put(g + "depths", np.asarray(depths, dtype=np.uint8))
        self.put_vlen_str(g + "candidates", [[str(k)] for k in keys])                   # DataStorePredict.py:60,65
        self.put(g + "candidate_frequency", np.asarray(freqs, dtype=np.uint8).reshape(n, 1))
        self.put(g + "base_prediction", np.asarray(probs, dtype=np.float64).reshape(n, 3))  # np.float in the reference


class PolishImageStore(_Store):
    """summaries/<contig>_<start>_<end>_<chunk>/{image, label, position, index, contig, region_start, region_end, chunk_id}"""

    def write_summary(self, contig: str, region_start: int, region_end: int, chunk_id: int, image, position, index, label=None):
        g = f"summaries/{contig}_{region_start}_{region_end}_{chunk_id}/"
        self.put(g + "image", np.asarray(image, dtype=np.uint8).reshape(1000, 10))
        self.put(g + "label", np.zeros(1000, dtype=np.uint8) if label is None else np.asarray(label, dtype=np.uint8))
        self.put(g + "position", np.asarray(position, dtype=np.int64))
        self.put(g + "index", np.asarray(index, dtype=np.int64))
        self.put_str(g + "contig", contig)
        self.put(g + "region_start", np.int64(region_start))
        self.put(g + "region_end", np.int64(region_end))
        self.put(g + "chunk_id", np.int64(chunk_id))


class PolishPredictionStore(_Store):
    """predictions/<contig>/<contig>-<rs>-<re>/{contig_start, contig_end}, .../<chunk>/{position, index, bases, phred_score}"""

    def write_prediction(self, contig: str, contig_start: int, contig_end: int, chunk_id: int, position, index, bases, phred):
        name = f"{contig}-{contig_start}-{contig_end}"
        g = f"predictions/{contig}/{name}/"
        self.put(g + "contig_start", np.int64(contig_start))
        self.put(g + "contig_end", np.int64(contig_end))
        c = g + f"{chunk_id}/"
        self.put(c + "position", np.asarray(position, dtype=np.int64))
        self.put(c + "index", np.asarray(index, dtype=np.int64))
        self.put(c + "bases", np.asarray(bases, dtype=np.uint8))
        self.put(c + "phred_score", np.asarray(phred, dtype=np.uint8))
