import ctypes as C, numpy as np, sys
sys.path.insert(0, '.')
from pepper_b200 import _lib
L = _lib.lib()
L.pb_test_tc_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
def run(M,N,K,kind,rep=1):
    rng = np.random.default_rng(1)
    A = rng.standard_normal((M, K)).astype(np.float32)
    if kind == 'rand':
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    else:
        W = np.zeros((N, K), np.float32)
        for n in range(N): W[n, (7*n+3) % K] = 1.0
    b = np.zeros(N, np.float32)
    want = A.astype(np.float64) @ W.astype(np.float64).T
    for r in range(rep):
        out = np.full((M, N), 7.0, np.float32)
        _lib.check(L.pb_test_tc_gemm(M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, out.ctypes.data), "gemm")
        bad = np.argwhere(np.abs(out - want) > 1e-3)
        print(M,N,K,kind,'rep',r,'bad',len(bad), 'rows', sorted(set(bad[:,0]))[:10], 'cols', sorted(set(bad[:,1]))[:10], 'vals', out[tuple(bad[0])] if len(bad) else None, flush=True)
for shape in [(128,128,32),(300,512,768),(300,512,768),(257,128,2048),(8192,1024,288)]:
    for kind in ('rand','sel'):
        run(*shape, kind, rep=3)
