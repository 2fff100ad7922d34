import torch, time
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def T(f, k=10):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k
one = T(lambda: d.copy_(h, non_blocking=True))
s = [torch.cuda.Stream() for _ in range(4)]
def multi(m):
    c = n // m
    for i in range(m):
        with torch.cuda.stream(s[i]):
            d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
print("1 stream : %.2f ms  %.1f GB/s" % (one * 1e3, n / one / 1e9))
for m in (2, 4):
    t = T(lambda: multi(m))
    print("%d streams: %.2f ms  %.1f GB/s" % (m, t * 1e3, n / t / 1e9))
# chunked 64 MiB sequential on one stream
def chunked():
    c = 64 << 20
    for i in range(n // c):
        d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
t = T(chunked); print("8 x 64MiB : %.2f ms  %.1f GB/s" % (t * 1e3, n / t / 1e9))
h2 = torch.empty(19 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(19 << 20, dtype=torch.uint8, device="cuda")
t = T(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 19 MiB: %.3f ms" % (t * 1e3))
