"""Times the dense -> CSR kernels on a block of raw counts (rows x genes uint32, ~7 % non-zero): CUDA events, block > L2."""
import sys

import torch

from scarf_b200 import lib, ops

rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8000, 30000)
ld = ops.round_up(cols, 4)
g = torch.Generator(device="cuda").manual_seed(1)
block = ((torch.rand((rows, ld), device="cuda", generator=g) < 0.07)
         * torch.randint(1, 50, (rows, ld), device="cuda", generator=g)).to(torch.int32)
cnt = torch.empty(rows, dtype=torch.int64, device="cuda")
ptr = None
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(4):
    s = torch.cuda.current_stream().cuda_stream
    ev[0].record()
    lib.call("scf_dense_row_nnz", block.data_ptr(), rows, cols, ld, cnt.data_ptr(), s)
    ev[1].record()
    ptr = torch.cumsum(cnt, 0) - cnt
    nnz = int(cnt.sum().item())
    idx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz, dtype=torch.int32, device="cuda")
    ev[2].record()
    lib.call("scf_dense_to_csr", block.data_ptr(), rows, cols, ld, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), s)
    ev[3].record()
    torch.cuda.synchronize()
t1, t2 = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
gb = rows * ld * 4 / 1e9
print(f"dense {rows} x {cols} ({gb:.2f} GB, nnz {nnz}): count {t1:.3f} ms = {gb / t1 * 1e3:.0f} GB/s, "
      f"fill {t2:.3f} ms = {(gb + nnz * 8e-9) / t2 * 1e3:.0f} GB/s")
