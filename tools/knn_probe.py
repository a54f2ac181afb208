"""Developer probe (GPU): scf_knn_l2 method 1 (tcgen05) against method 0 (FP64 brute force) with timings and
guard-failure counts.  usage: python tools/knn_probe.py [n dim k]...   (KNN_PROBE_NQ=<rows>: only the first <rows>
cells are queries -- the shape one rank of a sharded run sees)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import ops  # noqa: E402


def embedding(n, dim, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    scale = torch.sqrt(80.0 * 0.93 ** torch.arange(dim, device="cuda", dtype=torch.float32) + 4.0)
    y = torch.randn((n, dim), generator=g, device="cuda") * scale
    ld = ops.round_up(dim, 32)
    out = torch.zeros((n, ld), device="cuda")
    out[:, :dim] = y
    return out


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    cases = [(5000, 25, 11), (20000, 50, 11), (100000, 50, 11)]
    if len(sys.argv) > 3:
        a = list(map(int, sys.argv[1:]))
        cases = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)]
    for n, dim, k in cases:
        y = embedding(n, dim)
        nq = min(n, int(os.environ.get("KNN_PROBE_NQ", n)))
        yq = y[:nq]
        st = {}
        kev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        kev[0].record(), kev[1].record()
        t1, (i1, d1) = timed(lambda: ops.knn_l2(yq, y, dim, k, self_offset=0, method=1, stats=st, kernel_events=kev))
        torch.cuda.synchronize()
        tk = kev[0].elapsed_time(kev[1])
        fails = int(st["guard_fail_rows"].item()) if st.get("guard_fail_rows") is not None else -1
        flop = 2.0 * nq * n * dim
        line = (f"nq={nq} n={n} dim={dim} k={k}: tc {t1:.3f} ms ({flop / t1 / 1e9:.1f} TFLOP/s alg), tensor kernel "
                f"{tk:.3f} ms ({flop / tk / 1e9:.1f} TFLOP/s alg), guard fails {fails}")
        if nq * n <= 4e10 and not os.environ.get("KNN_PROBE_NO_EXACT"):
            t0, (i0, d0) = timed(lambda: ops.knn_l2(yq, y, dim, k, self_offset=0, method=0), reps=1)
            same_i = bool(torch.equal(i0, i1))
            same_d = bool(torch.equal(d0, d1))
            line += f" | exact {t0:.1f} ms | idx equal {same_i} dist equal {same_d}"
            if not same_i:
                bad = (i0 != i1).any(dim=1).nonzero().flatten()
                line += f" | {bad.numel()} rows differ, first {bad[:5].tolist()}"
                r = int(bad[0])
                line += f"\n   row {r}: exact {i0[r].tolist()} {d0[r].tolist()}\n   tc    {i1[r].tolist()} {d1[r].tolist()}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
