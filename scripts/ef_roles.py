"""Per-role wait/total cycle counters of the CTA-pair fused EdgeConv kernel (cluster 0)."""
import sys, ctypes, torch
sys.path.insert(0, '.')
import bench
from graphnet_b200 import ops, _lib
ops.set_precision('tf32')
dev = torch.device('cuda', 0)
tr = bench.Trainer(dev, 1)
db = bench.to_device(bench.host_batches(1024, 1, 777)[0], dev)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
import os
lib = _lib.load()
lib.gnb_edgeconv_set_variant(int(os.environ.get('EF_DEBUG', '0')) << 8)
for _ in range(2):
    tr.infer_step(db)
_lib.load().gnb_edgeconv_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
tr.infer_step(db)
torch.cuda.synchronize()
v = buf.cpu().tolist()
names = ["meta: wait epi_done", "meta: -", "meta: total", "mma: wait b_full", "mma: wait tmem_empty", "mma: total",
         "epi: wait tmem_full", "epi: arrives", "epi: total", "builder: wait meta_full", "builder: wait b_empty", "builder: total",
         "signal: wait named bar", "signal: -", "signal: total"]
for n, x in zip(names, v):
    print(f"{n:28s} {x:12d} cycles")
