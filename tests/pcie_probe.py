import torch, time
n = 16*1024*1024
h_in = torch.empty(n//4, dtype=torch.float32).pin_memory(); h_out = torch.empty(n//4, dtype=torch.float32).pin_memory()
d_in = torch.empty(n//4, dtype=torch.float32, device='cuda'); d_out = torch.empty(n//4, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a,b,c = t(h2d), t(d2h), t(both)
print(f"16 MiB: H2D {a*1e3:.3f} ms ({n/a/1e9:.1f} GB/s)  D2H {b*1e3:.3f} ms ({n/b/1e9:.1f} GB/s)  both concurrently {c*1e3:.3f} ms ({2*n/c/1e9:.1f} GB/s total)")
