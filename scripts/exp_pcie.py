"""PCIe ceiling of the host-buffer entry points: pinned H2D / D2H alone and concurrently (what bounds bench.py's e2e)."""
import torch, time
n_in, n_out = 512 << 20, 320 << 20
hin = torch.empty(n_in, dtype=torch.uint8).pin_memory(); hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
din = torch.empty(n_in, dtype=torch.uint8, device="cuda"); dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D 512 MiB alone {a*1e3:.2f} ms = {n_in/a/1e9:.1f} GB/s; D2H 320 MiB alone {b*1e3:.2f} ms = {n_out/b/1e9:.1f} GB/s; both concurrently {c*1e3:.2f} ms "
      f"(H2D-equivalent {n_in/c/1e9:.1f} GB/s) -> e2e ceiling {(1<<24)/c/1e6:.0f} Mrays/s")
