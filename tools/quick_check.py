"""One-process check of an update-kernel build: (1) oracle parity of a spread of off-policy cases, (2) a dump of every loss /
weight / moment of three cases for a BITWISE comparison with another build (SPP_RL_B200_LIB=<other .so>), (3) the headline timing
(148 agents x 50 updates, 1 M rings, bench.py's population).  Prints one flushed line per result, so a cut-off run still tells.
    python tools/quick_check.py --parity --dump /tmp/a.npz --time
    SPP_RL_B200_LIB=variants/libspp_rl_b200_old.so python tools/quick_check.py --dump /tmp/b.npz --time --compare /tmp/a.npz"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--dump", default="")
    ap.add_argument("--compare", default="")
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    t00 = time.perf_counter()
    import torch

    from tests.parity_util import run_offpolicy_parity_case
    from tests.test_gpu_update_parity import DDPG_CASES, SAC_CASES
    print("lib", os.environ.get("SPP_RL_B200_LIB", "default"), "import %.1fs" % (time.perf_counter() - t00), flush=True)
    dump_cases = [("sac", SAC_CASES[0]), ("sac", SAC_CASES[1]), ("ddpg", DDPG_CASES[0])]
    if a.dump:
        out = {}
        for i, (algo, case) in enumerate(dump_cases):
            err, dump = run_offpolicy_parity_case(algo=algo, verbose=False, per_agent=True, oracle_agents=(), **case)
            for name, x in dump:
                out["%d.%s" % (i, name)] = np.asarray(x)
        np.savez(a.dump, **out)
        print("dumped %d arrays -> %s (%.1fs)" % (len(out), a.dump, time.perf_counter() - t00), flush=True)
        if a.compare:
            other = np.load(a.compare)
            same = set(other.files) == set(out) and all(np.array_equal(other[k], out[k]) for k in out)
            worst = max(float(np.max(np.abs(other[k].astype(np.float64) - out[k].astype(np.float64)))) for k in out) if set(other.files) == set(out) else -1
            print("BITWISE %s vs %s (max abs diff %.3e over %d arrays)" % ("EQUAL" if same else "DIFFERENT", a.compare, worst, len(out)), flush=True)
    if a.parity:
        worst_all = 0.0
        for algo, case in [("sac", SAC_CASES[0]), ("sac", SAC_CASES[1]), ("sac", SAC_CASES[7]), ("sac", SAC_CASES[8]), ("sac", SAC_CASES[9]),
                           ("sac", SAC_CASES[10]), ("sac", SAC_CASES[6]), ("ddpg", DDPG_CASES[0]), ("ddpg", DDPG_CASES[1]), ("ddpg", DDPG_CASES[4])]:
            w = run_offpolicy_parity_case(algo=algo, verbose=False, **case)
            worst_all = max(worst_all, w)
            print("parity %s ob%d B%d: %.3e %s" % (algo, case["ob"], case["batch"], w, "ok" if w < 1e-5 else "FAIL"), flush=True)
        print("PARITY worst %.3e -> %s (%.1fs)" % (worst_all, "GREEN" if worst_all < 1e-5 else "RED", time.perf_counter() - t00), flush=True)
    if a.time:
        import bench
        pop = bench.build_population(0, 148)
        st = torch.cuda.Stream()
        for i in range(2):
            pop.update_ring_device(50, seed=100 + i, stream=st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for k in range(3):
                pop.update_ring_device(50, seed=1000 + k, stream=st.cuda_stream)
            e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("TIMING 148 agents x 50 updates: %.2f ms per burst = %.0f updates/s (%.1fs)" % (ms, 148 * 50 / (ms * 1e-3), time.perf_counter() - t00), flush=True)
        pop.close()


if __name__ == "__main__":
    main()
