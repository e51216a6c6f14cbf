"""Timing experiment on one CTA: full GEMM vs no-MMA vs no-staging vs no-epilogue (reps>>16 carries the debug mask)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib
lib = _lib.load_library()
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
for a_km, b_km in [(1, 0), (0, 0)]:
    for M, K in [(256, 256), (128, 256)]:
        rng = np.random.RandomState(1)
        A = rng.randn(M, K).astype(np.float32); B = rng.randn(256, K).astype(np.float32)
        a_s = np.ascontiguousarray(A if a_km else A.T); b_s = np.ascontiguousarray(B if b_km else B.T)
        out = np.zeros((M, 256), np.float32)
        for mode, name in [(0, "full"), (4, "no-epilogue"), (5, "staging only"), (6, "mma only"), (7, "sync skeleton"), (3, "epilogue only")]:
            ts = []
            for reps in (1, 201):
                ms = C.c_float(0)
                _lib.check(lib.spp_umma_gemm_selftest(a_km, b_km, M, K, reps | (mode << 16), f(a_s), f(b_s), f(out), C.byref(ms)))
                ts.append(ms.value)
            print("a_km %d b_km %d M %3d K %3d %-14s %.2f us" % (a_km, b_km, M, K, name, (ts[1] - ts[0]) / 200 * 1e3))
