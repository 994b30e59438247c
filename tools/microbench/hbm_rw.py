"""HBM bandwidth by direction on this GPU (dev tool): pure write (fill_), pure read (sum), copy (read + write).
The roofline denominator MEASURED_PEAKS.json:hbm_gbs is the COPY figure; write-only kernels (the activation stash, the dZ
images) are bounded by the write-only one."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)
res = {}
for name, fn, nbytes in (("write_fill", lambda: a.fill_(1.0), 4 * n), ("read_sum", lambda: a.sum(), 4 * n),
                         ("copy", lambda: b.copy_(a), 8 * n), ("write_zero", lambda: a.zero_(), 4 * n)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    res[name] = round(nbytes / best / 1e6, 1)
print(json.dumps({"GB/s": res}))
