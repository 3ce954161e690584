import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import bjxa_b200, batchgen
from bjxa_b200.api import PLAN_ENCODE_SEARCH, PLAN_ENCODE
from oracle import binding
lib = bjxa_b200.load()
o = binding.Oracle()
specs = [dict(bits=8, channels=1, frames=64, key=1), dict(bits=4, channels=2, frames=40, key=2)]
descs, arena, xa_bytes, pcms = batchgen.encode_batch(specs)
for kind in (PLAN_ENCODE, PLAN_ENCODE_SEARCH):
    d_src = lib.gpu_alloc(max(arena.size, 16)); d_dst = lib.gpu_alloc(xa_bytes + 64)
    lib.upload(d_src, arena); lib.upload(d_dst, np.full(xa_bytes + 64, 0xCD, dtype=np.uint8))
    plan = lib.plan_create(kind, descs)
    print("launches", lib.plan_launches(plan))
    lib.plan_run(plan, d_dst, xa_bytes + 64, d_src, arena.size)
    print("sync", lib._bjxa_gpu_sync(0))
    out = lib.plan_fetch(plan, descs.size)
    xa = lib.download(d_dst, xa_bytes + 64)
    print(kind, xa[:40], out["prev"].tolist())
want, st = o.encode_search_blocks(8, 1, [[0,0],[0,0]], pcms[0])
print("want", want[:40], st)
