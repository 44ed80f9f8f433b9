import torch, time
n = 1<<31
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device='cuda'); d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
for f in (h2d, d2h, both): f()
print("h2d GB/s", n/t(h2d)/1e9); print("d2h GB/s", n/t(d2h)/1e9); print("both each GB/s", n/t(both)/1e9)
import os; print(os.cpu_count()); os.system("nvidia-smi topo -m | head -5; free -g | head -2; lscpu | grep -E 'Model name|Socket|NUMA node' ")
