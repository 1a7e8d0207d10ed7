"""Ad-hoc probe: pinned-host <-> device copy rates of the box (the bound of bench.py's e2e figure), one direction at
a time and both at once, for a few transfer sizes."""
import json
import torch

dev = torch.device("cuda", 0)
res = {}
for mb in (64, 192, 1024):
    n = mb << 20
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in = torch.empty(n // 6, dtype=torch.uint8).pin_memory()
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    d_in = torch.empty(n // 6, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=8):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def d2h():
        with torch.cuda.stream(s1):
            h_out.copy_(d_out, non_blocking=True)

    def h2d():
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)

    def both():
        d2h(); h2d()

    import time
    def wall(fn, reps=8):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    res[mb] = {"d2h_gbs": n / wall(d2h) / 1e6, "h2d_gbs": (n // 6) / wall(h2d) / 1e6, "both_d2h_gbs": n / wall(both) / 1e6}
print(json.dumps(res))
