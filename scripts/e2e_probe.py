import sys, time, gc, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
hosts = bench.host_batches(512, 3, 20240607)
devs = [bench.to_device(h, dev) for h in hosts]
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
def run(steps=60, do_gc=None):
    out = []
    for i in range(steps):
        flush.fill_(float(i))
        torch.cuda.synchronize()
        g0 = gc.get_count()
        t0 = time.perf_counter()
        loss = tr.train_step(devs[i % 3])
        t2 = time.perf_counter()
        v = loss.item()
        t3 = time.perf_counter()
        out.append((t2 - t0, t3 - t2, t3 - t0, g0))
    return out
for tag in ('gc on', 'gc off'):
    if tag == 'gc off':
        gc.collect(); gc.disable()
    r = run()[5:]
    tot = sorted(x[2] for x in r)
    print(tag, 'median %.2f ms  max %.2f ms' % (1e3 * tot[len(tot) // 2], 1e3 * tot[-1]))
    for x in r:
        if x[2] > 1.3 * tot[len(tot) // 2]:
            print('   outlier: enqueue %.2f wait %.2f total %.2f gc_count %s' % (1e3 * x[0], 1e3 * x[1], 1e3 * x[2], x[3]))
