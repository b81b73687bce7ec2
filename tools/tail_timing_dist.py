"""Per-CTA timeline of the one-launch apply on several ranks (library built with SEMB_EXTRA_FLAGS=-DSEMB_TAIL_TIMING):
    torchrun --nproc-per-node 2 tools/tail_timing_dist.py 9 1112 139     (elements per rank in y)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import spectralelements_jl_b200 as sem

nr, Ex, Ey = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pcg = len(sys.argv) > 4 and sys.argv[4] == "pcg"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sem.init(local)
ctx.comm_init_torch()
m = sem.Mesh(nr, nr, Ex, Ey * world, (False, False), "wavy", ctx=ctx)
u, out, x = m.field().fill_random(1), m.field(), m.field()
pl = m.plan()
ncta = pl["nstrips"] * pl["ngroups"]
if pcg:
    m.pcg_begin(u, x, nu=1.0, k=0.0, bc="DDDD", tol=0.0, maxiter=10 ** 9)
fn = (lambda: m.pcg_iterate(1)) if pcg else (lambda: m.oplhs_device(u, out, nu=1.0, k=0.0, bc="DDDD"))
for _ in range(20):
    fn()
ctx.sync()
dist.barrier()
ctx.timer_start()
for _ in range(100):
    fn()
ms = ctx.timer_stop() / 100
buf = np.zeros(8 * ncta, dtype=np.int64)
sem._lib.check(ctx.lib.semb_mesh_debug_read(m.h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), ncta))
d = buf.reshape(ncta, 8).astype(np.float64)
t0 = d[:, 0].min()
us = lambda c: (d[:, c] - t0) / 1e3
ns, ng = pl["nstrips"], pl["ngroups"]
lines = ["rank %d: %.2f us per %s, plan %s" % (rank, ms * 1e3, "pcg iteration" if pcg else "apply", pl)]
for by in range(ng):
    sl = slice(by * ns, (by + 1) * ns)
    two = d[sl, 3].max() > 0
    rows_done = us(3)[sl] if two else us(1)[sl]
    ann = us(4)[sl] if two else us(2)[sl]
    lines.append("  CTA row %2d%s: start %5.1f  rows done med %6.1f max %6.1f  announced max %6.1f  prologue max %6.1f  tasks done max %6.1f  (micro med %d)%s"
                 % (by, " (2 chunks)" if two else "", us(0)[sl].max(), np.median(rows_done), rows_done.max(), ann.max(), us(5)[sl].max(),
                    us(6)[sl].max(), np.median(d[sl, 7]), ("  first chunk done med %.1f announced %.1f" % (np.median(us(1)[sl]), np.median(us(2)[sl]))) if two else ""))
for r in range(world):
    dist.barrier()
    if r == rank:
        print("\n".join(lines), flush=True)
m.free()
sem.finalize()
dist.destroy_process_group()
