"""PCIe probe (not product): pinned host<->device copy bandwidth alone and both ways at once, 256 MiB transfers."""
import time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, it=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * it / dt / 1e9
run(True, True, 3)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s each" % (run(True, False), run(False, True), run(True, True)))
