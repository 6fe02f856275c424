"""What the PCIe link of this box gives the host-buffer step: copy-engine bandwidth per direction for the step's byte counts
(2.1 MB in, 5.6 MB out), alone and both directions at once.
    python scripts/pcie_probe.py"""
import torch

dev = torch.device("cuda", 0)
h_in = torch.empty(2097152, dtype=torch.uint8).pin_memory()
h_out = torch.empty(5636096, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    cur = torch.cuda.current_stream(dev)
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)


t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
print(f"H2D 2.1 MB: {t_in:6.1f} us = {2097152 / t_in / 1e3:5.1f} GB/s")
print(f"D2H 5.6 MB: {t_out:6.1f} us = {5636096 / t_out / 1e3:5.1f} GB/s")
print(f"both at once: {t_both:6.1f} us (sum of the two alone {t_in + t_out:6.1f} us)")
