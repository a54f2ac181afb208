"""Developer probe (GPU): what the FP64 paths of this GPU deliver -- cuBLAS DGEMM on a large square and on the
eigensolver's tall x skinny shape (the yardstick for eig_dgemm_mma_kernel), measured with CUDA events."""
import torch


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for m, k, n in [(8192, 8192, 8192), (2000, 2000, 128), (2000, 2000, 96), (2000, 128, 128)]:
        a = torch.randn((m, k), device="cuda", dtype=torch.float64, generator=g)
        b = torch.randn((k, n), device="cuda", dtype=torch.float64, generator=g)
        ms = timed(lambda: torch.mm(a, b), reps=5 if m > 4000 else 50)
        print(f"cuBLAS DGEMM {m} x {k} x {n}: {ms * 1e3:.1f} us, {2.0 * m * k * n / ms / 1e9:.2f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
