"""HBM probe: pure-write (fill) and copy bandwidth on this GPU, to judge write-heavy kernels against."""
import torch

dev = "cuda:0"
n = 1 << 29                      # 2 GiB of fp32
a = torch.empty(n, device=dev)
b = torch.empty(n, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t = timed(lambda: a.fill_(1.0))
print(f"fill  : {n * 4 / t / 1e9:8.1f} GB/s written")
t = timed(lambda: b.copy_(a))
print(f"copy  : {2 * n * 4 / t / 1e9:8.1f} GB/s read+written")
t = timed(lambda: torch.sum(a))
print(f"reduce: {n * 4 / t / 1e9:8.1f} GB/s read")
