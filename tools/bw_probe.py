import torch, time
dev = torch.device("cuda:0")
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for gb in (1, 4):
    x = torch.empty(gb * (1 << 30) // 2, dtype=torch.bfloat16, device=dev).normal_()
    y = torch.empty_like(x)
    ms = t(lambda: y.copy_(x)); print(f"copy {gb} GiB: {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s (read+write)")
    ms = t(lambda: x.float().sum() if False else torch.sum(x, dtype=torch.float32)); print(f"sum  {gb} GiB: {x.numel() * 2 / ms / 1e6:.0f} GB/s (read only)")
    xi = x.view(torch.int32)
    ms = t(lambda: torch.max(xi)); print(f"max(int32) {gb} GiB: {x.numel() * 2 / ms / 1e6:.0f} GB/s (read only)")
    ms = t(lambda: y.zero_()); print(f"memset {gb} GiB: {x.numel() * 2 / ms / 1e6:.0f} GB/s (write only)")
    del x, y, xi
