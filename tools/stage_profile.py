"""Stage timeline of the fused update burst (agent 0 of a full population): microseconds per stage, from %globaltimer stamps
inside the kernel (spp_update_stage_profile).  python tools/stage_profile.py [P] [algo] [ob] [ac]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

NAMES = ["bookkeeping+gather", "A actor fwd (fc1,fc2,heads)", "A sample", "A acm fwd", "A target critics hidden", "A q target",
         "B critics hidden", "B critic head bwd + Adam fc3/b3/b2", "B dz1 (dX fc2)", "B dW2,dW1 + Adam",
         "C actor fwd", "C sample", "C acm fwd", "C critics hidden", "C policy head bwd", "C dz1 (dX fc2)", "C dxc (dX fc1)",
         "C acm bwd dx", "C actor head bwd + alpha", "C dza2 (dX heads)", "C dza1 (dX fc2)", "C actor dW x3 + Adam", "losses out"]


def main():
    import bench
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    pop = bench.build_population(0, P, 100_000)
    for _ in range(2):
        pop.update_ring_device(10, seed=1)
    pop.sync()
    out = np.zeros(32, np.float64)
    n = C.c_int()
    from spp_rl_b200._lib import check
    check(pop.lib.spp_update_stage_profile(pop.h, 20, 5, out.ctypes.data_as(C.POINTER(C.c_double)), 32, C.byref(n)))
    tot = out[: n.value].sum()
    print("P=%d  total %.1f us per update (agent 0)" % (P, tot))
    for name, v in zip(NAMES, out[: n.value]):
        print("  %-40s %8.1f us  %5.1f %%" % (name, v, 100 * v / tot))
    pop.close()


if __name__ == "__main__":
    main()
