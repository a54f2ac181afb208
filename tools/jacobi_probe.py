"""Developer probe (GPU): sweeps and time of the one-CTA Jacobi eigensolver on the Rayleigh-Ritz matrices of eig_topk."""
import sys
import torch
sys.path.insert(0, ".")
from scarf_b200 import graph, ops  # noqa: E402

h = 2000
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn((20000, h), device="cuda", dtype=torch.float64, generator=g)
f = torch.randn((20000, 65), device="cuda", dtype=torch.float64, generator=g)
w = torch.randn((65, h), device="cuda", dtype=torch.float64, generator=g) * (1.2 * 0.93 ** torch.arange(65, device="cuda"))[:, None]
z = z + f @ (w * (torch.rand((65, h), device="cuda", dtype=torch.float64, generator=g) < 0.04))
z = (z - z.mean(0)) / z.std(0)
cov = (z.T @ z) / (z.shape[0] - 1)
orig = ops.sym_eig_small
log = []


def spy(a, info=None, max_n=ops.JACOBI_MAX_N):
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(a, info, max_n)
    e1.record()
    torch.cuda.synchronize()
    off = (a - torch.diag(torch.diagonal(a))).abs().max() / torch.diagonal(a).abs().max()
    log.append((int(a.shape[0]), int(info.item()), e0.elapsed_time(e1), float(off)))
    return out


ops.sym_eig_small = spy
for dims in (25, 50):
    log.clear()
    st = {}
    graph.eig_topk(cov, dims, stats=st)
    graph.eig_topk(cov, dims, stats=st)
    print(dims, st, [(n, sw, round(ms, 3), f"{off:.1e}") for n, sw, ms, off in log])
