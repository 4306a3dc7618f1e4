import time, torch
n=40_000_000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device="cuda")
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def one():
    d.copy_(h,non_blocking=True)
def two():
    with torch.cuda.stream(s1): d[:n//2].copy_(h[:n//2],non_blocking=True)
    with torch.cuda.stream(s2): d[n//2:].copy_(h[n//2:],non_blocking=True)
def five():   # five separate arrays like the upload
    k=n//5
    for i in range(5): d[i*k:(i+1)*k].copy_(h[i*k:(i+1)*k],non_blocking=True)
for nm,f in (("one stream",one),("two streams",two),("five copies one stream",five)):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(20): f()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
    print(f"H2D 40 MB {nm}: {dt*1e3:.3f} ms ({40/dt/1e3:.1f} GB/s)")
h2=torch.empty(21_000_000,dtype=torch.uint8).pin_memory()
def bidir():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d[:21_000_000],non_blocking=True)
for _ in range(3): bidir()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20): bidir()
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/20
print(f"H2D 40 MB + D2H 21 MB concurrently: {dt*1e3:.3f} ms")
