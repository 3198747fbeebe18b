import torch, time
n = 4455188001 // 4
x = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in [("fill_", lambda: x.fill_(1.5)), ("zero_", lambda: x.zero_()), ("copy(2x bytes)", None)]:
    if fn is None:
        y = torch.empty_like(x)
        fn = lambda: y.copy_(x)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    b = n * 4 * (2 if "copy" in name else 1)
    print(name, "%.3f ms" % ms, "%.0f GB/s" % (b / ms / 1e6))
