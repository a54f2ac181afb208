"""Developer probe (GPU): scf_gram_accumulate tensor-core modes against float64, with layout diagnostics."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scarf_b200 import lib, ops  # noqa: E402

torch.set_printoptions(linewidth=200, precision=4)


def run(n_rows, n_cols, ldz, pattern="rand"):
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.zeros((n_rows, ldz), dtype=torch.float32, device="cuda")
    if pattern == "rand":
        z[:, :n_cols] = torch.randn((n_rows, n_cols), generator=g, device="cuda")
    elif pattern == "onehot":  # row r has a single 1 at column r % n_cols, value r+1: G diagonal = sum of squares
        for r in range(n_rows):
            z[r, r % n_cols] = float(r + 1)
    elif pattern == "colid":   # z[r, c] = c + 1 for row 0 only: G[m, n] = (m+1)(n+1)
        z[0, :n_cols] = torch.arange(1, n_cols + 1, device="cuda", dtype=torch.float32)
    z_lo = z - (z.view(torch.int32) & -8192).view(torch.float32)
    ref = (z[:, :n_cols].double().T @ z[:, :n_cols].double()).cpu().numpy()
    for mode in (1, 3):
        gfx = ops.gram_accumulate(z, n_rows, n_cols, mode=mode, z_lo=z_lo)
        torch.cuda.synchronize()
        gm = gfx[:n_cols, :n_cols].cpu().numpy().astype(np.float64) * 2.0 ** -lib.GRAM_SHIFT
        up = np.triu(np.ones_like(ref, dtype=bool))
        err = np.abs(gm - ref)[up]
        print(f"rows {n_rows} cols {n_cols} ldz {ldz} {pattern} mode {mode}: max|err| {err.max():.4g} "
              f"max|ref| {np.abs(ref).max():.4g} nonzero {np.count_nonzero(gm)} of {up.sum()} (upper)")
        if err.max() > 1e-2 * np.abs(ref).max():
            k = min(n_cols, 12)
            print(" gm :\n", np.array2string(gm[:k, :k], precision=2, max_line_width=200))
            print(" ref:\n", np.array2string(ref[:k, :k], precision=2, max_line_width=200))


if __name__ == "__main__":
    run(8, 40, 64, "colid")
    run(8, 40, 64, "onehot")
    run(3000, 500, 512)
