import torch
x = torch.empty(4_300_000_000, dtype=torch.uint8, device='cuda')
y = torch.empty_like(x)
for name, fn, nbytes in [('memset', lambda: x.zero_(), x.numel()), ('copy', lambda: y.copy_(x), 2 * x.numel())]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('%s: %.3f ms  %.2f TB/s' % (name, ms, nbytes / ms / 1e9))
