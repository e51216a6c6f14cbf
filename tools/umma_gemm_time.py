"""Time the pipelined tensor-core GEMM (one CTA, 256 x 256 x K, repeated inside the kernel) and report its accuracy."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib

lib = _lib.load_library()
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
for a_km, b_km in [(1, 0), (0, 0), (1, 1)]:
    for M, K in [(256, 256), (256, 16), (128, 256)]:
        rng = np.random.RandomState(1)
        A = rng.randn(M, K).astype(np.float32); B = rng.randn(256, K).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        a_s = np.ascontiguousarray(A if a_km else A.T); b_s = np.ascontiguousarray(B if b_km else B.T)
        out = np.zeros((M, 256), np.float32)
        ts = []
        for reps in (1, 201):
            ms = C.c_float(0)
            _lib.check(lib.spp_umma_gemm_selftest(a_km, b_km, M, K, reps, f(a_s), f(b_s), f(out), C.byref(ms)))
            ts.append(ms.value)
        per = (ts[1] - ts[0]) / 200 * 1e3
        rel = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        ref32 = A @ B.T
        rel32 = np.linalg.norm(ref32 - ref) / np.linalg.norm(ref)
        print("a_km %d b_km %d M %3d K %3d: %.2f us per GEMM (%.1f TFLOP/s algorithmic on one SM x148 = %.0f), relnorm vs fp64 %.2e (numpy fp32: %.2e)"
              % (a_km, b_km, M, K, per, 2 * M * 256 * K / per / 1e6, 2 * M * 256 * K / per / 1e6 * 148, rel, rel32))
