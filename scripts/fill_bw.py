import torch, time
x = torch.empty(4_246_732_800, dtype=torch.uint8, device="cuda")
y = torch.empty_like(x)
for name, fn in [("zero_", lambda: x.zero_()), ("copy_", lambda: y.copy_(x))]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, f"{ms:.3f} ms", f"{x.numel() * (2 if name == 'copy_' else 1) / ms / 1e6:.0f} GB/s")
