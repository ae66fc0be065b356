import torch, time
n = 45*2*65536  # one ciphertext = 45 MiB
h = torch.empty(2*n, dtype=torch.int64).pin_memory(); d = torch.empty(2*n, dtype=torch.int64, device="cuda")
h2 = torch.empty(n, dtype=torch.int64).pin_memory(); d2 = torch.empty(n, dtype=torch.int64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
a = t(lambda: d.copy_(h, non_blocking=True)); print("H2D 90MiB: %.2f ms  %.1f GB/s" % (a*1e3, h.numel()*8/a/1e9))
b = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H 45MiB: %.2f ms  %.1f GB/s" % (b*1e3, h2.numel()*8/b/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print("both: %.2f ms  agg %.1f GB/s" % (c*1e3, (h.numel()+h2.numel())*8/c/1e9))
import subprocess; print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv; lscpu | grep -E 'Model name|NUMA|Socket' | head; numactl -H 2>/dev/null | head -5", shell=True, capture_output=True, text=True).stdout)
