import torch, time
dev = torch.device("cuda:0")
for nbytes in (524288, 1310720, 5111808, 20*1024*1024):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50): fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 50
        print(name, nbytes, "%.1f us  %.1f GB/s" % (dt * 1e6, nbytes / dt / 1e9))
# sync latency: tiny kernel + sync
x = torch.zeros(1, device=dev)
for _ in range(10): x.add_(1); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): x.add_(1); torch.cuda.synchronize()
print("launch+sync %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
