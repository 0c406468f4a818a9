"""HBM bandwidth of pure write / pure read / copy / 1:2 read:write mixes on this GPU (torch kernels),
to put the write-heavy stage kernels into perspective."""
import torch, json
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n = 1 << 29   # 4 GiB of doubles
x = torch.empty(n, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
x.fill_(1.0)
out = {}
out['write_GBps'] = 8 * n / t(lambda: y.fill_(2.0)) / 1e6
out['read_GBps'] = 8 * n / t(lambda: x.sum()) / 1e6
out['copy_GBps'] = 16 * n / t(lambda: y.copy_(x)) / 1e6
# one read : two writes
z = torch.empty_like(x)
out['r1w2_GBps'] = 24 * n / t(lambda: (y.copy_(x), z.copy_(x))) / 1e6   # 2 reads + 2 writes actually
out['axpy_r2w1_GBps'] = 24 * n / t(lambda: torch.add(x, y, out=z)) / 1e6
print(json.dumps(out))
