"""CPU (needs /root/reference, i.e. runs in the build container): the four store layouts written by pepper_b200/datastore.py
are read back by the reference's OWN reader code — pepper_variant dataloader_predict.SequenceDataset (a9),
pepper dataloader_predict.SequenceDataset (a12), pepper Stitch.small_chunk_stitch (a14) and
pepper_variant CandidateFinder.small_chunk_stitch (a16, which parses ``str(candidates[i])``) — through an `h5py` stand-in
backed by the npz container (h5py / libhdf5 are in neither image).  The stand-in hands datasets back unchanged, so a wrong
dtype (e.g. fixed bytes instead of str) breaks the reference's parsing exactly as it would under h5py 2.10."""
import os
import sys
import types
import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


class _Leaf:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, k):
        if k != ():
            return self.a[k]
        return self.a[()] if self.a.shape == () else self.a


class _Node:
    def __init__(self, store, prefix):
        self.s, self.p = store, prefix

    def keys(self):
        return self.s.keys(self.p) if self.p else sorted({k.split("/", 1)[0] for k in self.s.data})

    def __contains__(self, k):
        return k in self.keys()

    def __getitem__(self, k):
        path = (self.p + "/" + k).strip("/")
        return _Leaf(self.s.data[path]) if path in self.s.data else _Node(self.s, path)


@pytest.fixture()
def ref_modules(monkeypatch):
    from pepper_b200 import datastore as ds

    class _File(_Node):
        def __init__(self, name, mode="r"):
            super().__init__(ds._Store(name, mode="r", backend="npz"), "")

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def close(self):
            pass
    h5 = types.ModuleType("h5py"); h5.File = _File
    tv = types.ModuleType("torchvision"); tvt = types.ModuleType("torchvision.transforms")
    tvt.Compose = lambda x: x; tvt.ToTensor = lambda: None; tv.transforms = tvt
    pv = types.ModuleType("pepper_variant.build.PEPPER_VARIANT")

    class CandidateImagePrediction:
        def __init__(self, contig, position, depth, candidates, candidate_frequency, prediction_base, prediction_type):
            self.contig, self.position, self.depth = contig, position, depth
            self.candidates, self.candidate_frequency = candidates, candidate_frequency
            self.prediction_base, self.prediction_type = prediction_base, prediction_type
    pv.CandidateImagePrediction = CandidateImagePrediction
    bv = types.ModuleType("pepper_variant.build"); bv.PEPPER_VARIANT = pv
    bp = types.ModuleType("pepper.build"); bp.PEPPER = types.ModuleType("pepper.build.PEPPER")
    for name, mod in {"h5py": h5, "torchvision": tv, "torchvision.transforms": tvt, "pepper_variant.build": bv,
                      "pepper_variant.build.PEPPER_VARIANT": pv, "pepper.build": bp, "pepper.build.PEPPER": bp.PEPPER}.items():
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.setattr(np, "int", int, raising=False)
    for m in [k for k in sys.modules if k.startswith(("pepper.modules", "pepper_variant.modules"))]:
        monkeypatch.delitem(sys.modules, m)
    return pv


def test_variant_image_store_read_by_reference_dataloader(tmp_path, ref_modules):
    from pepper_b200 import datastore as ds
    from pepper_variant.modules.python.models.dataloader_predict import SequenceDataset
    f = str(tmp_path / "img.hdf5")
    rng = np.random.default_rng(1)
    imgs = rng.integers(-128, 128, size=(3, 33, 26)).astype(np.int8)
    keys = ["1T", "2ACG", "3" + "ACGT" * 15]
    with ds.VariantImageStore(f, backend="npz") as s:
        s.write_summary("chr20_1000_2000", "chr20", [1001, 1500, 1999], [30, 125, 7], keys, [5, 12, 3], imgs)
    data = SequenceDataset(None, input_file=f)
    assert len(data) == 3
    for i in range(3):
        contig, position, depth, candidate, freq, image = data[i]
        assert contig == "chr20" and position == [1001, 1500, 1999][i] and depth == [30, 125, 7][i]
        assert [str(c) for c in candidate] == [keys[i]] and isinstance(candidate[0], str)
        assert list(freq) == [[5], [12], [3]][i] and np.array_equal(image, imgs[i]) and image.dtype == np.int8
    batch = SequenceDataset.my_collate([data[0], data[1]])
    assert tuple(batch[5].shape) == (2, 33, 26)


def test_variant_prediction_store_read_by_reference_candidate_finder(tmp_path, ref_modules):
    from pepper_b200 import datastore as ds
    from oracle import find_candidates as ofc
    rng = np.random.default_rng(2)
    L = 600
    genome = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
    n = 120
    positions = np.sort(rng.integers(0, L - 8, n))
    keys = []
    for p in positions:
        t = rng.integers(1, 4)
        keys.append("1" + "ACGT"[rng.integers(0, 4)] if t == 1 else ("2" + genome[p] + "AC" if t == 2 else "3" + genome[p:p + 3]))
    depths = rng.integers(4, 60, n); freqs = np.minimum(depths, rng.integers(1, 30, n))
    probs = rng.dirichlet([0.6, 0.5, 0.4], n).astype(np.float32)

    class FASTA_handler:
        def __init__(self, path):
            pass

        def get_reference_sequence(self, contig, a, b):
            return genome[max(0, a):max(0, b)]
    ref_modules.FASTA_handler = FASTA_handler
    from pepper_variant.modules.python import CandidateFinder as RefCF
    f = str(tmp_path / "pred.hdf")
    with ds.VariantPredictionStore(f, backend="npz") as s:
        s.write_prediction(0, ["ctg"] * n, positions, depths, keys, freqs, probs)
    options = dict(snp_p_value=0.1, insert_p_value=0.1, delete_p_value=0.1, snp_p_value_in_lc=0.3, insert_p_value_in_lc=0.35,
                   delete_p_value_in_lc=0.25, report_snp_above_freq=0.0, report_indel_above_freq=0.6)
    got_m, got_d = RefCF.small_chunk_stitch(types.SimpleNamespace(fasta="x", **options), [(f, "batch_0")])
    want_m, want_d = ofc.select(options, "ctg", positions, depths, keys, freqs, probs.astype(np.float64), lambda c, a, b: genome[max(0, a):max(0, b)])
    assert len(got_d) == len(want_d) > 20 and len(got_m) == len(want_m) > 5       # fixed-bytes candidates would drop every record
    for a, b in zip(got_d, want_d):
        assert (a[0], int(a[1]), int(a[2]), a[3], a[4], a[5], int(a[6]), [int(x) for x in a[7]], bool(a[11])) == \
               (b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[11])
        assert abs(float(a[8]) - float(b[8])) < 1e-12


def test_polish_image_store_read_by_reference_dataloader(tmp_path, ref_modules):
    from pepper_b200 import datastore as ds
    from pepper.modules.python.models.dataloader_predict import SequenceDataset
    f = str(tmp_path / "pimg.hdf")
    img = (np.arange(10000) % 255).astype(np.uint8).reshape(1000, 10)
    with ds.PolishImageStore(f, backend="npz") as s:
        s.write_summary("ctg1", 0, 1100, 1, img, np.arange(1000), np.zeros(1000, np.int64))
    data = SequenceDataset(None, file_list=[f])
    assert len(data) == 1
    contig, cs, ce, cid, image, position, index = data[0]
    assert isinstance(contig, str) and contig == "ctg1" and "b'" not in (contig + "-" + str(cs))
    assert (int(cs), int(ce), int(cid)) == (0, 1100, 1) and np.array_equal(image, img) and position.dtype == np.int64


def test_polish_prediction_store_read_by_reference_stitch(tmp_path, ref_modules):
    from pepper_b200 import datastore as ds
    from oracle import stitch as ostitch
    from pepper.modules.python import Stitch as RefStitch
    rng = np.random.default_rng(3)
    regions = [(0, 1100), (900, 2000)]
    f = str(tmp_path / "ppred.hdf")
    imgs = []
    with ds.PolishPredictionStore(f, backend="npz") as s:
        for r, (a, b) in enumerate(regions):
            n = b - a + 1
            start, end, cid = 0, min(n, 1000), 0
            while True:
                pos = np.full(1000, -1, np.int64); idx = np.full(1000, -1, np.int64)
                pos[:end - start] = np.arange(a + start, a + end); idx[:end - start] = 0
                bases = rng.integers(0, 5, 1000).astype(np.uint8)
                s.write_prediction("ctg1", a, b, cid, pos, idx, bases, np.zeros(1000, np.uint8))
                imgs.append((r, cid, pos, idx, bases)); cid += 1
                if end == n:
                    break
                start = end - 50; end = min(n, start + 1000)
    first, last, seq = RefStitch.small_chunk_stitch("ctg1", [(f, "ctg1", a, b) for a, b in regions])
    want = ostitch.stitch(np.stack([i[4] for i in imgs]), np.stack([i[2] for i in imgs]), np.stack([i[3] for i in imgs]),
                          np.array([i[0] for i in imgs]), np.array([i[1] for i in imgs]), [r[0] for r in regions], [r[1] for r in regions])
    assert seq == want and len(seq) > 1000
