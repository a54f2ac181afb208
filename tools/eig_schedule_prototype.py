"""Developer prototype (CPU, numpy/torch): block width / start schedule of the Chebyshev-filtered subspace iteration on the
covariance of a C2 step (tools/build/c2_cov.npy, dumped on the GPU in round 1) -- rounds, column products, dense sizes.
The native solver (scarf_b200/csrc/eig_topk.cu) uses the (dims + 32, four start steps of degree 3) line."""
import numpy as np, torch, math, sys, time
torch.set_num_threads(8)
cov = torch.from_numpy(np.load('tools/build/c2_cov.npy'))
h = cov.shape[0]
wfull, vfull = torch.linalg.eigh(cov)

def cheb(cov, x, degree, cut, top):
    e = 0.5*cut; t = (top-e)/e
    r = 1.0/(t+math.sqrt(t*t-1))
    j = torch.arange(1, degree+1, dtype=torch.float64)
    sig = r*(1+r**(2*(j-1)))/(1+r**(2*j))
    a = 2*sig[1:]/e; d = sig[:-1]*sig[1:]
    cs = cov - e*torch.eye(h, dtype=torch.float64)
    y = (cs@x)*(sig[0]/e)
    for i in range(degree-1):
        yn = (cs@y)*a[i] - x*d[i]
        x, y = y, yn
    return y

def cholqr2(y):
    y = y/y.norm(dim=0, keepdim=True)
    for _ in range(2):
        l = torch.linalg.cholesky(y.T@y)
        y = torch.linalg.solve_triangular(l, y.T, upper=False).T
    return y

def growth(x): return x+math.sqrt(max(x*x-1,0))

def run(dims, b0, keep, deg0_steps, tol=1e-8, max_rounds=8, verbose=True):
    g = torch.Generator().manual_seed(4466)
    q = torch.randn((h,b0), dtype=torch.float64, generator=g)
    trace = torch.diagonal(cov).sum().item(); top0 = cov.abs().sum(0).max().item()
    mv = 0
    for d in deg0_steps:
        q = cholqr2(cheb(cov, q, d, trace/h, top0)); mv += d*q.shape[1]
    eigs = []
    for rnd in range(1, max_rounds+1):
        aq = cov@q; mv += q.shape[1]
        t = q.T@aq
        w, s = torch.linalg.eigh(t); eigs.append(t.shape[0])
        topv, wt = s[:, -dims:], w[-dims:]
        v = q@topv
        res = ((aq@topv - v*wt).norm(dim=0).max()/w[-1]).item()
        width = q.shape[1]; kp = min(width, keep)
        th_min, th_max, th_dims = w[0].item(), w[-1].item(), wt[0].item()
        bulk = (trace - w.sum().item())/(h-width)
        th_keep = w[-kp].item(); bulk_keep = (trace - w[-kp:].sum().item())/(h-kp)
        if verbose: print(f"  round {rnd}: width {width} res {res:.2e}")
        if res <= tol:
            ang = torch.acos(((vfull[:, -dims:].flip(1))*v.flip(1)).sum(0).abs().clamp(max=1)).max().item()
            return rnd, mv, eigs, res, ang
        if kp < width and th_keep < 0.8*th_dims:
            s = s[:, -kp:]; th_min, bulk = th_keep, bulk_keep
        cut = max(th_min, min(bulk, 0.5*(th_min+th_dims)))
        cut = min(max(cut, 1e-3*th_dims), 0.9*th_dims)
        e = 0.5*cut
        rho = growth((th_max-e)/e)/growth((th_dims-e)/e)
        m = int(max(2, min(32, math.floor(math.log(1e20)/math.log(max(rho, 1+1e-9))))))
        y = cheb(cov, q@s.flip(1), m, cut, th_max); mv += m*s.shape[1]
        q = cholqr2(y)
        if verbose: print(f"     filter m={m} cut={cut:.3f} width->{s.shape[1]}")
    return -1, mv, eigs, res, None

if __name__ == "__main__":
    dims = 50
    for (b0, keep, steps) in [(2*dims+64, dims+32, [3,3]), (dims+32, dims+32, [3,3]), (dims+32, dims+32, [3,3,3]), (dims+46, dims+46, [3,3]), (dims+46, dims+46, [3,3,3]),(dims+32, dims+32, [3,3,3,3])]:
        t=time.time()
        print(b0, keep, steps, run(dims, b0, keep, steps), f"{time.time()-t:.1f}s")
