"""CPU: the accounting constants of bench.py agree with the model they describe (161.4 MFLOP per variant candidate;
pepper_variant simple_model.py:23-46 shapes) and with the shipped product mask of the network kernels (handles.cuh)."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("pb_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_variant_flop_split_matches_the_model_shapes():
    b = _bench()
    steps, dirs, feat, hidden, gates = 33, 2, 26, 256, 4 * 256
    want = {"encoder_x": 2 * feat * gates, "encoder_h": 2 * hidden * gates, "decoder_x": 2 * 2 * hidden * gates, "decoder_h": 2 * hidden * gates}
    for k, per_step in want.items():
        assert abs(b.VARIANT_GEMMS[k][0] - per_step * steps * dirs) / (per_step * steps * dirs) < 0.01, k
    head = 2 * (2 * hidden * steps) * 512 + 4 * 2 * 512 * 512 + 2 * 512 * 3
    assert abs(b.VARIANT_GEMMS["head"][0] - head) / head < 0.01
    total = sum(f for f, _ in b.VARIANT_GEMMS.values())
    assert abs(total - b.FLOP_PER_CAND) / b.FLOP_PER_CAND < 0.005
    assert 2.0 <= b.PRODUCTS_VARIANT <= 3.0


def test_products_per_gemm_follow_the_shipped_mask():
    """bit set in TcVariant::lo_mask = that GEMM keeps its third product (bit 0 encoder h, 1 decoder x, 2 decoder h, 3 linear_1,
    4 linear_2-5); the encoder's x-part has no lo operand at all (int8 images)."""
    b = _bench()
    src = open(os.path.join(ROOT, "pepper_b200", "csrc", "handles.cuh")).read()
    mask = int(re.search(r"struct TcVariant.*?int lo_mask = (0x[0-9a-fA-F]+);", src, re.S).group(1), 16)
    products = {k: p for k, (_, p) in b.VARIANT_GEMMS.items()}
    assert products["encoder_x"] == 2
    assert products["encoder_h"] == 2 + (mask & 1)
    assert products["decoder_x"] == 2 + ((mask >> 1) & 1)
    assert products["decoder_h"] == 2 + ((mask >> 2) & 1)
    assert products["head"] == 2 + ((mask >> 3) & 1) == 2 + ((mask >> 4) & 1)
