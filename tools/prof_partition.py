"""Single-GPU exercise of the shuffle's local kernels for ncu: the stable 8-way partition plan of 300 M int64 keys (what a
rank runs on its filtered lineitem shard before an exchange), the local scatter of two payload columns, the 8 000-way
plan the partition-based as-of path uses, and a hash aggregate of 30 M rows.  Run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quokka_b200 import _lib as L, ops

torch.cuda.set_device(0)
n = 300_000_000
g = torch.Generator(device="cuda").manual_seed(1)
key = torch.randint(0, 1 << 40, (n,), device="cuda", dtype=torch.int64, generator=g)
a = torch.rand(n, device="cuda", dtype=torch.float64, generator=g)
for rep in range(2):
    dest, offs = ops.partition_plan(key, 8, L.PART_MOD)
    outs = ops.scatter([key, a], dest)
    torch.cuda.synchronize()
    del outs, dest
sym = (key % 8000).to(torch.int32)
dest, offs = ops.partition_plan(sym, 8000, L.PART_CODE)
torch.cuda.synchronize()
del dest
m = 30_000_000
ha = ops.HashAggState([torch.int64], [L.AGG_SUM], 2 * m, "cuda")
ha.update([key[:m] % 3_000_000], [a[:m]])
ha.finalize()
torch.cuda.synchronize()
print("ok")
