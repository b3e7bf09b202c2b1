"""PCIe probe for the e2e leg of bench.py: H2D alone, D2H alone, both at once, for the byte volume of one config-3 step."""
import torch, time
n = 1_386_231_680 // 4
h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for _ in range(2): run(True, True, 1)
a, b, c = run(True, False), run(False, True), run(True, True)
gb = n * 4 / 1e9
print(f"H2D {a:.1f} ms ({gb / a * 1e3:.1f} GB/s)  D2H {b:.1f} ms ({gb / b * 1e3:.1f} GB/s)  both {c:.1f} ms ({2 * gb / c * 1e3:.1f} GB/s total)")
