"""How much of a buffer read by one kernel is still in L2 for the next kernel? (B200: 126 MB L2)"""
import torch
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def t(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3
for mb in (8, 16, 24, 32, 48, 64, 80, 96, 128):
    a = torch.randn(mb << 18, device="cuda")
    b = torch.empty_like(a)
    res = []
    for rep in range(3):
        flush.fill_(rep); torch.cuda.synchronize()
        cold = t(lambda: a.sum())
        warm = t(lambda: a.sum())
        flush.fill_(rep); torch.cuda.synchronize()
        t(lambda: a.sum())
        # second pass in reverse halves: read, then copy (read + write)
        warm_copy = t(lambda: b.copy_(a))
        flush.fill_(rep); torch.cuda.synchronize()
        cold_copy = t(lambda: b.copy_(a))
        res.append((cold, warm, cold_copy, warm_copy))
    c, w, cc, wc = [min(r[i] for r in res) for i in range(4)]
    print("%4d MB: sum cold %.1f us (%.2f TB/s)  warm %.1f us (%.2f TB/s) | copy cold %.1f us  after a read %.1f us"
          % (mb, c, mb * 1.048576 / c, w, mb * 1.048576 / w, cc, wc))
