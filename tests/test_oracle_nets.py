"""CPU: the network oracle (oracle/nets.py) reproduces the golden vectors that were generated from the
reference's own nn.Module classes (tests/golden/make_golden_nets.py)."""
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_variant_oracle_golden():
    from oracle import nets
    g = np.load(os.path.join(GOLD, "variant_net_seed0.npz"))
    p, h = nets.variant_predict(nets.make_variant_weights(0), g["images"], return_hidden=True)
    assert np.abs(p - g["probs"]).max() < 1e-5 and np.abs(h[:4] - g["hidden"]).max() < 1e-5


def test_polish_oracle_golden():
    from oracle import nets
    g = np.load(os.path.join(GOLD, "polish_net_seed0.npz"))
    b, ph, h, a = nets.polish_predict(nets.make_polish_weights(0), g["images"])
    assert np.abs(h[:, :, :, ::8] - g["hidden"]).max() < 1e-5
    assert (b != g["bases"]).mean() < 1e-3
    # hidden-state plumbing quirk: window k+1 starts from the decoder state of window k
    assert np.abs(h[1] - h[0]).max() > 1e-3
