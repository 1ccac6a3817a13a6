import torch, time
for mb in (4, 16, 64):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    s = torch.cuda.current_stream()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    t0=time.perf_counter()
    for _ in range(10):
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    wall=(time.perf_counter()-t0)/10*1e3
    print(f"D2H {mb} MiB: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s; wall incl sync {wall:.3f} ms")
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"H2D {mb} MiB: {ms:.3f} ms = {n/ms/1e6:.1f} GB/s")
