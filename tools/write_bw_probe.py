import torch
x = torch.empty(2**32, dtype=torch.float32, device="cuda")  # 17.2 GB
for name, fn in [("fill_", lambda: x.fill_(1.0)), ("zero_", lambda: x.zero_())]:
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, ms, "ms", x.numel() * 4 / ms / 1e6, "GB/s")
y = torch.empty(2**31, dtype=torch.float32, device="cuda"); z = torch.empty_like(y)
for _ in range(2): z.copy_(y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): z.copy_(y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("copy", ms, "ms", 2 * y.numel() * 4 / ms / 1e6, "GB/s (read+write)")
