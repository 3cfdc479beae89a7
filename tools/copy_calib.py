"""Calibration: what plain streaming kernels achieve at the tensor sizes of the normalisation kernels (CUDA-graph timed,
12 back-to-back launches rotating over 4 buffer pairs).  memcpy = cudaMemcpyAsync D2D (copy_), add = ATen vectorised
elementwise kernel (read + write), sum = ATen reduction (read only)."""
import torch


def timed(fn, reps=12):
    for r in range(4):
        fn(r)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=side):
        for r in range(reps):
            fn(r % 4)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for mb in (10.5, 21, 42, 84, 168, 336, 1024):
    n = int(mb * 1e6 / 2)
    src = [torch.randn(n, device="cuda").bfloat16() for _ in range(4)]
    dst = [torch.empty_like(s) for s in src]
    out = [torch.empty((), device="cuda", dtype=torch.float32) for _ in range(4)]
    t_cp = timed(lambda i: dst[i].copy_(src[i]))
    t_add = timed(lambda i: torch.add(src[i], 1.0, out=dst[i]))
    t_sum = timed(lambda i: torch.sum(src[i], dim=(0,), dtype=torch.float32, out=out[i]))
    print("%7.1f MB tensor: memcpy %7.1f us %5.0f GB/s | add (r+w) %7.1f us %5.0f GB/s | sum (r) %7.1f us %5.0f GB/s" %
          (mb, t_cp, 2 * mb * 1e3 / t_cp, t_add, 2 * mb * 1e3 / t_add, t_sum, mb * 1e3 / t_sum))
    del src, dst
