"""GPU: the tcgen05 GEMM (bf16 hi/lo split, three products) against a float64 numpy GEMM, and both networks in
tensor-core mode against the CPU oracle with the same tolerances as the fp32 path."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-3
MARGIN = 1e-4


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 96), (100, 256, 288), (300, 512, 768), (257, 256, 2048)])
def test_tc_gemm_vs_numpy(M, N, K):
    from pepper_b200 import _lib
    L = _lib.lib()
    L.pb_test_tc_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    out = np.zeros((M, N), np.float32)
    _lib.check(L.pb_test_tc_gemm(M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, out.ctypes.data), "pb_test_tc_gemm")
    want = A.astype(np.float64) @ W.astype(np.float64).T + b
    err = np.abs(out - want).max()
    assert err < 2e-4, err        # ~2^-16 relative per product; |terms| ~ 1/sqrt(K)
    # structured check: identity-like weights pick single columns exactly (catches any tile / descriptor permutation)
    W2 = np.zeros((N, K), np.float32)
    for n in range(N):
        W2[n, (7 * n + 3) % K] = 1.0
    out2 = np.zeros((M, N), np.float32)
    zb = np.zeros(N, np.float32)
    _lib.check(L.pb_test_tc_gemm(M, N, K, A.ctypes.data, W2.ctypes.data, zb.ctypes.data, out2.ctypes.data), "gemm")
    want2 = A[:, [(7 * n + 3) % K for n in range(N)]]
    assert np.abs(out2 - want2).max() < 1e-4


def test_variant_net_tc_mode():
    from oracle import nets
    from pepper_b200.variant import VariantNet
    from tests.test_nets_gpu import _variant_images
    state = nets.make_variant_weights(3)
    x = _variant_images(700, 3)
    want, whid = nets.variant_predict(state, x, threads=8, return_hidden=True)
    net = VariantNet(state)
    net.set_mode(1)
    got, hid = net.predict(x, return_hidden=True)
    assert np.abs(hid - whid).max() < TOL, np.abs(hid - whid).max()
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()
    srt = np.sort(want, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > MARGIN
    assert np.array_equal(got.argmax(1)[clear], want.argmax(1)[clear])
    net.set_mode(0)
    ref = net.predict(x)
    assert np.abs(ref - got).max() < TOL


def test_polish_net_tc_mode():
    from oracle import nets
    from pepper_b200.polish import PolishNet
    from tests.test_nets_gpu import _polish_images, _check_bases
    state = nets.make_polish_weights(6)
    x = _polish_images(140, 6)
    wb, wp, wh, wa = nets.polish_predict(state, x, threads=8)
    net = PolishNet(state)
    net.set_mode(1)
    bases, phred, hid, acc = net.predict(x, debug=True)
    assert np.abs(hid - wh).max() < TOL, np.abs(hid - wh).max()
    assert np.abs(acc - wa).max() < TOL, np.abs(acc - wa).max()
    _check_bases(bases, wb, wa, phred, wp, acc)
