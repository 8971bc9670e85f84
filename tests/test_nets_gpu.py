"""GPU parity of the two networks against the CPU PyTorch oracle (oracle/nets.py, itself pinned bit-exact against
the reference's nn.Module classes by tests/golden/make_golden_nets.py).

Tolerances (BASELINE.json north_star / BASELINE.md §3): hidden states within 1e-3 absolute (fp32); class indices
bit-exact wherever the oracle's top-2 margin exceeds 1e-4; softmax probabilities within 1e-3."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-3
MARGIN = 1e-4


def _variant_images(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(-40, 41, size=(n, 33, 26)).astype(np.int8)
    x[:, :, 0] = rng.integers(1, 6, size=(n, 33))
    x[rng.random((n, 33, 26)) < 0.5] = 0
    return x


def test_variant_net_golden():
    from oracle import nets
    from pepper_b200.variant import VariantNet
    g = np.load(os.path.join(GOLD, "variant_net_seed0.npz"))
    net = VariantNet(nets.make_variant_weights(0))
    net.set_mode(0)
    probs, hid = net.predict(g["images"], return_hidden=True)
    assert np.abs(hid[:4] - g["hidden"]).max() < TOL
    assert np.abs(probs - g["probs"]).max() < TOL
    srt = np.sort(g["probs"], axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > MARGIN
    assert np.array_equal(probs.argmax(1)[clear], g["probs"].argmax(1)[clear])
    assert net.launches() == 33 * 2 + 5 + 1


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,seed", [(1, 1), (130, 2), (700, 3)])
def test_variant_net_vs_oracle(n, seed, mode):
    from oracle import nets
    from pepper_b200.variant import VariantNet
    state = nets.make_variant_weights(seed)
    x = _variant_images(n, seed)
    want, whid = nets.variant_predict(state, x, threads=8, return_hidden=True)
    net = VariantNet(state)
    net.set_mode(mode)
    got, hid = net.predict(x, return_hidden=True)
    assert np.abs(hid - whid).max() < TOL, np.abs(hid - whid).max()
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()
    srt = np.sort(want, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > MARGIN
    assert np.array_equal(got.argmax(1)[clear], want.argmax(1)[clear])
    assert clear.mean() > 0.9
    assert np.allclose(got.sum(1), 1.0, atol=1e-5)


def _polish_images(n, seed):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 1000, 10), np.uint8)
    cov = rng.integers(0, 255, size=(n, 1000, 1))
    x[:] = (rng.random((n, 1000, 10)) < 0.25) * cov
    if n > 1:
        x[-1, 600:] = 0           # zero padded tail chunk
    return x


def test_polish_net_golden():
    from oracle import nets
    from pepper_b200.polish import PolishNet
    g = np.load(os.path.join(GOLD, "polish_net_seed0.npz"))
    net = PolishNet(nets.make_polish_weights(0))
    net.set_mode(0)
    bases, phred, hid, acc = net.predict(g["images"], debug=True)
    assert np.abs(hid[:, :, :, ::8] - g["hidden"]).max() < TOL
    assert np.abs(acc[:, ::10] - g["acc"]).max() < TOL
    _check_bases(bases, g["bases"], None, phred, g["phred"], acc_full=None)


def _check_bases(bases, want_bases, want_acc, phred, want_phred, acc_full):
    if want_acc is not None:
        srt = np.sort(want_acc, axis=2)
        clear = (srt[:, :, -1] - srt[:, :, -2]) > MARGIN
        assert np.array_equal(bases[clear], want_bases[clear])
        assert clear.mean() > 0.99
    else:
        assert (bases != want_bases).mean() < 1e-3
    # phred: uint8 truncation of an fp32 log10 -> may differ by 1 at integer boundaries
    d = np.abs(phred.astype(np.int32) - want_phred.astype(np.int32))
    assert (d <= 1).mean() > 0.999


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,seed", [(1, 4), (5, 5), (140, 6)])
def test_polish_net_vs_oracle(n, seed, mode):
    from oracle import nets
    from pepper_b200.polish import PolishNet
    state = nets.make_polish_weights(seed)
    x = _polish_images(n, seed)
    wb, wp, wh, wa = nets.polish_predict(state, x, threads=8)
    net = PolishNet(state)
    net.set_mode(mode)
    bases, phred, hid, acc = net.predict(x, debug=True)
    assert np.abs(hid - wh).max() < TOL, np.abs(hid - wh).max()        # hidden carried through all 19 windows
    assert np.abs(acc - wa).max() < TOL, np.abs(acc - wa).max()
    _check_bases(bases, wb, wa, phred, wp, acc)
    b2, p2 = net.predict(x)
    assert np.array_equal(b2, bases) and np.array_equal(p2, phred)
    assert net.launches() == (19 * 201 + 1 if mode == 0 else 19 * 4 + 1)     # tcgen05: pack, encoder window, decoder window, dense+softmax


def test_variant_net_chunk_invariance():
    """Full-size property: a candidate's probabilities do not depend on what else is in the batch or on where the
    9,472-candidate chunk boundaries fall (20,000 candidates = 3 chunks vs the same rows predicted in small batches)."""
    from pepper_b200 import weights
    from pepper_b200.variant import VariantNet
    net = VariantNet(weights.random_variant_state(3))
    x = _variant_images(20000, 9)
    full = net.predict(x)
    assert np.isfinite(full).all() and np.allclose(full.sum(1), 1.0, atol=1e-5)
    for lo, hi in ((0, 100), (9400, 9600), (18900, 20000)):
        assert np.array_equal(net.predict(x[lo:hi]), full[lo:hi])
    assert np.array_equal(net.predict(x), full)                      # idempotent


def test_polish_net_batch_invariance():
    from pepper_b200 import weights
    from pepper_b200.polish import PolishNet
    net = PolishNet(weights.random_polish_state(4))
    x = _polish_images(300, 8)
    b, p = net.predict(x)
    b2, p2 = net.predict(x[120:140])
    assert np.array_equal(b2, b[120:140]) and np.array_equal(p2, p[120:140])
