"""Developer probe (GPU): time of graph.fit_kmeans at the C2 shape (100k cells, 50 dims, 1,000 centres)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from scarf_b200 import graph, ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
n, d, nc = 100_000, 50, 1000
y = torch.zeros((n, ops.round_up(d, 32)), device="cuda")
y[:, :d] = torch.randn((n, d), device="cuda", generator=g) * torch.linspace(8, 1, d, device="cuda")
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    c, l = graph.fit_kmeans(y, d, nc, 4466)
    torch.cuda.synchronize()
    print(f"fit_kmeans {n} x {d}, {nc} centres: {1e3 * (time.time() - t0):.1f} ms", flush=True)
