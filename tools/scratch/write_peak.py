"""Scratch: pure-write / pure-read / copy HBM rates on this box (torch kernels), for context."""
import torch
n = 1_728_000_000  # bytes, about one tick's observation output for 2M envs
x = torch.empty(n, dtype=torch.int8, device="cuda")
y = torch.empty(n, dtype=torch.int8, device="cuda")
def t(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: x.zero_()); print("memset      %.3f ms  %.0f GB/s written" % (ms, n / ms / 1e6))
ms = t(lambda: x.fill_(1)); print("fill kernel %.3f ms  %.0f GB/s written" % (ms, n / ms / 1e6))
x32 = x.view(torch.int32)
ms = t(lambda: x32.fill_(7)); print("fill i32    %.3f ms  %.0f GB/s written" % (ms, n / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy        %.3f ms  %.0f GB/s read+written" % (ms, 2 * n / ms / 1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print("sum (read)  %.3f ms  %.0f GB/s read" % (ms, n / ms / 1e6))
