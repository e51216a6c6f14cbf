"""Where does the tensor-core GEMM lose accuracy?  Inputs exactly representable in tf32 (lo = 0) isolate the accumulator."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib
lib = _lib.load_library()
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))

def tf32(x):
    b = x.view(np.uint32)
    return ((b + 0x1000) & 0xFFFFE000).view(np.float32)

for positive in (0, 1):
    for K in (16, 64, 256, 1024):
        for exact in (1, 0):
            rng = np.random.RandomState(K)
            A = rng.randn(256, K).astype(np.float32); B = rng.randn(256, K).astype(np.float32)
            if positive: A = np.abs(A); B = np.abs(B)
            if exact: A = tf32(A); B = tf32(B)
            ref = A.astype(np.float64) @ B.astype(np.float64).T
            out = np.zeros((256, 256), np.float32)
            _lib.check(lib.spp_umma_gemm_selftest(1, 0, 256, K, 1, f(A), f(np.ascontiguousarray(B.T)), f(out), None))
            seq = np.zeros((256, 256), np.float32)
            err = out - ref
            scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
            print("positive %d K %4d tf32-exact inputs %d: relnorm %.2e  mean(err/scale) %+.2e  rms(err/scale) %.2e   numpy fp32 relnorm %.2e"
                  % (positive, K, exact, np.linalg.norm(err) / np.linalg.norm(ref), (err / scale).mean(), np.sqrt(((err / scale) ** 2).mean()),
                     np.linalg.norm((A @ B.T) - ref) / np.linalg.norm(ref)))
