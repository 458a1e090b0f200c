"""How much faster is L2 than HBM for streaming traffic?  Device-to-device copies (read + write) and reductions (read only) of
buffers that fit the 126 MB L2 against ones that do not.  Decides whether keeping hand-overs L2-resident can beat the HBM roofline."""
import json, torch
torch.cuda.init()
def t(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
for mb in (4, 8, 16, 32, 48, 64, 128, 512, 2048):
    n = mb * (1 << 20) // 4
    a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
    b = torch.empty_like(a)
    reps = max(10, 4096 // mb)
    tc = t(lambda: b.copy_(a), reps)
    tr = t(lambda: a.sum(), reps)
    tw = t(lambda: b.fill_(1.0), reps)
    print(json.dumps({"buffer_MB": mb, "copy_GBps_read_plus_write": 2 * n * 4 / tc / 1e9, "read_GBps": n * 4 / tr / 1e9, "write_GBps": n * 4 / tw / 1e9}), flush=True)
