"""Pins oracle/nets.py against the reference's own nn.Module classes imported from /root/reference, and writes
small golden vectors.  Run in the build container:  python tests/golden/make_golden_nets.py"""
import os
import sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle import nets  # noqa: E402
from pepper.modules.python.models.simple_model import TransducerGRU as RefPolish  # noqa: E402
from pepper_variant.modules.python.models.simple_model import TransducerGRU as RefVariant  # noqa: E402

torch.set_num_threads(1)
rng = np.random.default_rng(5)

# ---- variant
sv = nets.make_variant_weights(0)
ref = RefVariant(26, 1, 256, 28, 3, bidirectional=True)
ref.load_state_dict(sv)
ref.eval()
imgs = rng.integers(-30, 31, size=(24, 33, 26)).astype(np.int8)
imgs[:, :, 0] = rng.integers(1, 6, size=(24, 33))
with torch.no_grad():
    want = ref(torch.from_numpy(imgs.astype(np.float32))).numpy()
got, hid = nets.variant_predict(sv, imgs, return_hidden=True)
assert np.array_equal(want, got), np.abs(want - got).max()
np.savez_compressed(os.path.join(HERE, "variant_net_seed0.npz"), images=imgs, probs=want, hidden=hid[:4],
                    torch_version=np.array(torch.__version__))
print("variant net pinned: max prob", want.max(), "argmax hist", np.bincount(want.argmax(1), minlength=3))

# ---- polish
sp = nets.make_polish_weights(0)
refp = RefPolish(1, 10, 1, 128, 5, bidirectional=True)
refp.load_state_dict(sp)
refp.eval()
pim = np.zeros((3, 1000, 10), np.uint8)
cov = rng.integers(0, 255, size=(3, 1000, 1))
pim[:] = (rng.random((3, 1000, 10)) < 0.25) * cov
pim[2, 700:] = 0
# the reference loop (predict.py:47-93 / predict_distributed_cpu.py:50-90) around the REFERENCE module
with torch.no_grad():
    images = torch.from_numpy(pim).type(torch.FloatTensor)
    hidden = torch.zeros(3, 2, 128)
    acc = torch.zeros(3, 1000, 5)
    hids = []
    for i in range(0, 1000, 50):
        if i + 100 > 1000:
            break
        out, hidden = refp(images[:, i:i + 100], hidden)
        hids.append(hidden.numpy().copy())
        acc = acc + torch.nn.ZeroPad2d((0, 0, i, 900 - i))(torch.softmax(out, dim=2))
    vals, labels = torch.max(acc, 2)
b, ph, h, a = nets.polish_predict(sp, pim)
assert np.array_equal(labels.numpy().astype(np.uint8), b)
assert np.array_equal(np.stack(hids), h), np.abs(np.stack(hids) - h).max()
assert np.array_equal(acc.numpy(), a)
np.savez_compressed(os.path.join(HERE, "polish_net_seed0.npz"), images=pim, bases=b, phred=ph, hidden=h[:, :, :, ::8],
                    acc=a[:, ::10], torch_version=np.array(torch.__version__))
print("polish net pinned: base hist", np.bincount(b.ravel(), minlength=5))
