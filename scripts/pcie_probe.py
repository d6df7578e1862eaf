import torch, time
x = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(x, device="cuda")
for n in (1 << 20, 4 << 20, 64 << 20):
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(20):
            if direction == "h2d": d[:n].copy_(x[:n], non_blocking=True)
            else: x[:n].copy_(d[:n], non_blocking=True)
        torch.cuda.synchronize(); el = time.perf_counter() - t
        print(direction, n >> 20, "MiB:", 20 * n / el / 1e9, "GB/s")
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); y = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); e = torch.empty_like(y, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(e, non_blocking=True)
torch.cuda.synchronize(); el = time.perf_counter() - t
print("duplex 64 MiB each way:", 20 * (64 << 20) / el / 1e9, "GB/s per direction")
