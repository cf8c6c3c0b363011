import sys, subprocess, os
CODE = r'''
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from graphnet_b200 import ops
from helpers import tie_heavy_events
ops.set_precision('tf32')
seq = [int(v) for v in sys.argv[1].split(',')]
sizes = [1, 2, 5, 9, 10, 64, 130, 12, 300]
x, batch, _ = tie_heavy_events(sizes, 5, seed=3)
ptr = ops.batch_to_ptr(batch.cuda(), len(sizes))
graph = ops.knn_table(x.cuda(), [0, 1, 2], ptr, 8)
n = x.shape[0]
ops.set_edgeconv_variant(2)
for hdim in seq:
    g = torch.Generator().manual_seed(1)
    pq = torch.randint(-2, 3, (n, 2 * hdim), generator=g).float().cuda()
    w2 = torch.randint(-1, 2, (256, hdim), generator=g).float().cuda()
    b2 = torch.randint(-3, 4, (256,), generator=g).float().cuda()
    y = ops.edgeconv_fused_forward(pq, w2, b2, graph, 'add')
    torch.cuda.synchronize()
    ops.set_edgeconv_variant(1)
    y1 = ops.edgeconv_fused_forward(pq, w2, b2, graph, 'add')
    torch.cuda.synchronize()
    ops.set_edgeconv_variant(2)
    print('hdim', hdim, 'ok, equal to single-CTA:', bool(torch.equal(y, y1)), flush=True)
'''
for seq in ("128", "336,336,336", "336,128", "64", "96", "160", "224", "352"):
    r = subprocess.run([sys.executable, "-c", CODE, seq], capture_output=True, text=True, timeout=120)
    print("SEQ", seq, "rc", r.returncode, "|", r.stdout.strip().replace("\n", " ; "), "|", r.stderr.strip().split("\n")[-1][:150] if r.returncode else "")
