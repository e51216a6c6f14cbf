"""One launch of the GEMM self-test kernel (for ncu): args a_km b_km M K reps mode"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib
lib = _lib.load_library()
a_km, b_km, M, K, reps, mode = [int(x) for x in sys.argv[1:7]]
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
rng = np.random.RandomState(1)
A = rng.randn(M, K).astype(np.float32); B = rng.randn(256, K).astype(np.float32)
a_s = np.ascontiguousarray(A if a_km else A.T); b_s = np.ascontiguousarray(B if b_km else B.T)
out = np.zeros((M, 256), np.float32)
ms = C.c_float(0)
_lib.check(lib.spp_umma_gemm_selftest(a_km, b_km, M, K, reps | (mode << 16), f(a_s), f(b_s), f(out), C.byref(ms)))
print("ms", ms.value)
