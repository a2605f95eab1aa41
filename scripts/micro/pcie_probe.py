"""Host<->device copy rates of the box (pinned memory) next to the end-to-end rate of gpc_match_batch at two batch sizes."""
import json, subprocess, sys, time
import torch
n = 229 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / 10
run(False); run(True)
t = run(False); print(f"H2D alone: {n / t / 1e9:.1f} GB/s")
t = run(True); print(f"H2D with concurrent D2H of half the bytes: {n / t / 1e9:.1f} GB/s (+ {n / 2 / t / 1e9:.1f} GB/s D2H)")
for b in (256, 1024):
    out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--batch", str(b), "--steps", "10"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    j = json.loads(out)
    print(f"batch {b}: value {j['value']:.0f} e2e {j['e2e']['value']:.0f} pairs/s, H2D {j['e2e']['h2d_bytes_per_step'] * j['e2e']['value'] / b / 1e9:.1f} GB/s")
