"""Hardware probes for the SPP-PPO critic's tensor-core kernel (spp_umma_probe): see umma_selftest.cu.  One mode per process
(python tools/umma_probe.py MODE): an unsupported layout raises a device exception that poisons the context."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib
lib = _lib.load_library()
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
rng = np.random.RandomState(0)
tf = lambda a: (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
A = tf(rng.randn(128, 32).astype(np.float32)); B = tf(rng.randn(128, 32).astype(np.float32))
mode = int(sys.argv[1])
out = np.zeros((128, 128), np.float32)
if mode >= 4:
    Bk = np.ascontiguousarray(B.T)          # [32 k][128 n]
    _lib.check(lib.spp_umma_probe(mode, f(A), f(Bk), f(out)))
else:
    _lib.check(lib.spp_umma_probe(mode, f(A), f(B), f(out)))
ref = A.astype(np.float64) @ B.astype(np.float64).T
if mode != 2:
    err = np.abs(out - ref).max() / np.abs(ref).max()
    print("mode", mode, "max rel err", err, "OK" if err < 1e-5 else "MISMATCH")
else:
    nan_rows = np.isnan(out).all(axis=1)
    print("mode 2 (M=64): lanes untouched:", np.nonzero(nan_rows)[0].tolist())
    for lane in range(128):
        if nan_rows[lane]:
            continue
        d = np.abs(ref[:64] - out[lane][None, :]).max(axis=1)
        r = int(np.argmin(d))
        print("lane %3d <- row %2d (err %.1e)" % (lane, r, d[r] / np.abs(ref).max()), end=";  " if lane % 4 != 3 else "\n")
