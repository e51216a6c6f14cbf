"""Updates/s of the fused SPP-SAC burst (bench.py's workload, device-resident) on each GEMM path of spp_set_gemm_path:
1 = tcgen05 3-pass tf32 split (fp32-accurate, the bench headline), 2 = tcgen05 single tf32 pass (reduced-precision variant,
stated tolerance 1e-2), 0 = FFMA tiles.  Prints one JSON line per path; alternates the paths to cancel clock drift."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import build_population  # noqa: E402
from spp_rl_b200 import _lib  # noqa: E402


def main():
    P, G, steps = 148, 50, 4
    lib = _lib.load_library()
    pop = build_population(0, P)
    ts = torch.cuda.Stream()      # a non-default stream: a null handle would make the library launch on its own stream
    stream = ts.cuda_stream
    res = {0: [], 1: [], 2: []}
    for rnd in range(3):
        for path in (1, 2, 0):
            lib.spp_set_gemm_path(path)
            for w in range(2):
                pop.update_ring_device(G, seed=10 + w, stream=stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            for k in range(steps):
                pop.update_ring_device(G, seed=100 + 10 * rnd + k, stream=stream)
            e1.record(ts); torch.cuda.synchronize()
            res[path].append(P * G * steps / (e0.elapsed_time(e1) * 1e-3))
    lib.spp_set_gemm_path(1)
    names = {1: "tcgen05 tf32 x3 split (fp32-accurate, default)", 2: "tcgen05 single tf32 pass (reduced precision, tol 1e-2)", 0: "FFMA tiles"}
    for path in (1, 2, 0):
        print(json.dumps({"gemm_path": path, "name": names[path], "updates_per_s": float(np.median(res[path])), "runs": res[path]}))


if __name__ == "__main__":
    main()
