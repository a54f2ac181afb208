"""Developer probe (GPU): scf_eig_topk on the covariance of a C2 step (tools/build/c2_cov.npy, dumped in round 1) --
time, rounds, residual, agreement with torch.linalg.eigh.  usage: python tools/eig_probe.py [dims ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import lib, ops  # noqa: E402


def main():
    path = os.environ.get("EIG_PROBE_COV", os.path.join("tools", "build", "c2_cov.npy"))
    cov = torch.from_numpy(np.load(path)).cuda()
    h = cov.shape[0]
    n = int(os.environ.get("EIG_PROBE_N", "100000"))
    ld = (h + 31) // 32 * 32
    g = torch.zeros((ld, ld), dtype=torch.int64, device="cuda")
    g[:h, :h] = torch.round(cov * (n - 1) * 2.0 ** lib.GRAM_SHIFT).to(torch.int64)
    scale = 2.0 ** -lib.GRAM_SHIFT / (n - 1)
    wf, vf = torch.linalg.eigh(g[:h, :h].double() * scale)
    for dims in [int(x) for x in sys.argv[1:]] or [50, 100]:
        st = {}
        ops.eig_topk(g, h, dims, scale, stats=st)
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 5
        for _ in range(reps):
            ev, load = ops.eig_topk(g, h, dims, scale, stats=st)
        torch.cuda.synchronize()
        ms = (time.time() - t0) / reps * 1e3
        err = float((ev - torch.flip(wf[-dims:], [0])).abs().max() / wf[-1])
        ref = torch.flip(vf[:, -dims:], [1])
        ang = float(torch.acos((ref * load).sum(0).abs().clamp(max=1.0)).max())
        print(f"dims={dims}: {ms:.3f} ms, {st}, eigenvalue error {err:.2e}, max angle {ang:.2e} rad", flush=True)


if __name__ == "__main__":
    main()
