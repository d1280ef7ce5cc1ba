import torch, time
x=torch.empty(288_000_000, dtype=torch.uint8, pin_memory=True)
d=torch.empty_like(x, device="cuda")
o=torch.empty(72_000_000, dtype=torch.uint8, device="cuda"); oh=torch.empty(72_000_000, dtype=torch.uint8, pin_memory=True)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(n, both):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): oh.copy_(o, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n
run(3,True)
a=run(20,False); b=run(20,True)
print(f"H2D alone: {a*1e3:.3f} ms = {288/a/1e3:.1f} GB/s; with concurrent D2H 72MB: {b*1e3:.3f} ms = {288/b/1e3:.1f} GB/s")
# two halves on two streams
h=144_000_000
def run2(n):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n):
        with torch.cuda.stream(s1): d[:h].copy_(x[:h], non_blocking=True)
        with torch.cuda.stream(s2): d[h:].copy_(x[h:], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n
c=run2(20); print(f"H2D split over 2 streams: {c*1e3:.3f} ms = {288/c/1e3:.1f} GB/s")
