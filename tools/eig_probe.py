"""Developer probe (GPU): time symmetric eigensolvers on an H x H covariance-like matrix."""
import sys
import time

import torch

sys.path.insert(0, ".")
from scarf_b200 import graph  # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn((20000, h), device="cuda", dtype=torch.float64, generator=g)
f = torch.randn((20000, 65), device="cuda", dtype=torch.float64, generator=g)
w = torch.randn((65, h), device="cuda", dtype=torch.float64, generator=g) * (1.2 * 0.93 ** torch.arange(65, device="cuda"))[:, None]
z = z + f @ w
z = (z - z.mean(0)) / z.std(0)
cov = (z.T @ z) / (z.shape[0] - 1)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3, out


for name, fn in (("eigh f64", lambda: torch.linalg.eigh(cov)), ("eigh f32", lambda: torch.linalg.eigh(cov.float())),
                 ("eigvalsh f64", lambda: torch.linalg.eigvalsh(cov)),
                 ("gemm f64 HxHx164", lambda: cov @ cov[:, :164]), ("gemm f64 HxHxH", lambda: cov @ cov),
                 ("eigh f64 164", lambda: torch.linalg.eigh(cov[:164, :164])),
                 ("eigh f64 512", lambda: torch.linalg.eigh(cov[:512, :512])),
                 ("qr f64 Hx164", lambda: torch.linalg.qr(cov[:, :164]))):
    ms, _ = timed(fn)
    print(f"{name:20s} {ms:8.3f} ms")
y164 = cov[:, :164].contiguous()
for name, fn in (("cholqr2 Hx164", lambda: graph._cholqr2(y164)), ("cholqr2 Hx82", lambda: graph._cholqr2(y164[:, :82].contiguous())),
                 ("cheb filter deg 3 Hx164", lambda: graph._cheb_filter(cov, y164, 3, 1.0, 200.0))):
    ms, _ = timed(fn, reps=10)
    print(f"{name:28s} {ms:8.3f} ms")
for dims in (25, 50, 100):
    st = {}
    ms, _ = timed(lambda: graph.eig_topk(cov, dims, stats=st), reps=2)
    print(f"eig_topk dims={dims}: {ms:.2f} ms {st}")
