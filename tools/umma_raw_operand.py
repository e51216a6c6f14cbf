"""What does kind::tf32 do with the 13 low mantissa bits of a raw fp32 operand -- truncate or round?  (spp_umma_selftest, split = 0:
the raw words are the operands.)  A = x on the diagonal, B = identity: C[i][i] shows the tf32 value the tensor core used."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import _lib
lib = _lib.load_library()
f = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
K = 128
xs = np.array([1 + 0.75 * 2 ** -10, 1 + 0.5 * 2 ** -10, 1 + 0.25 * 2 ** -10, 1 + 2 ** -10 + 0.99 * 2 ** -10, -(1 + 0.75 * 2 ** -10), 3.1415927,
               1 + (2 ** 13 - 1) * 2.0 ** -23], np.float32)
A = np.zeros((128, K), np.float32); B = np.zeros((128, K), np.float32)
for i, x in enumerate(xs):
    A[i, i] = x
for i in range(128):
    B[i, i] = 1.0
out = np.zeros((128, 128), np.float32)
_lib.check(lib.spp_umma_selftest(0, 0, K, 0, f(A), f(B), f(out)))
trunc = (xs.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
rnd = ((xs.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
for i, x in enumerate(xs):
    got = out[i, i]
    print("x=%.9g  got=%.9g  trunc=%.9g  round=%.9g  -> %s" % (x, got, trunc[i], rnd[i], "TRUNC" if got == trunc[i] else ("ROUND" if got == rnd[i] else "OTHER")))
# same for the B operand
A2 = np.zeros((128, K), np.float32); B2 = np.zeros((128, K), np.float32)
for i in range(128):
    A2[i, i] = 1.0
for i, x in enumerate(xs):
    B2[i, i] = x
_lib.check(lib.spp_umma_selftest(0, 0, K, 0, f(A2), f(B2), f(out)))
for i, x in enumerate(xs):
    got = out[i, i]
    print("B: x=%.9g  got=%.9g -> %s" % (x, got, "TRUNC" if got == trunc[i] else ("ROUND" if got == rnd[i] else "OTHER")))
