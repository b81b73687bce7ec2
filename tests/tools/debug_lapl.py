import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import sem_oracle as so
import spectralelements_jl_b200 as sem
ctx = sem.init(0)
nr, Ex, Ey = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
om = so.make_mesh(nr, nr, Ex, Ey, (False, False), so.wavy)
gm = sem.Mesh.from_arrays(nr, nr, Ex, Ey, (False, False), om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
print(gm.plan())
if len(sys.argv) > 4: gm.set_chunks(int(sys.argv[4]))
u = so.splitmix_uniform(gm.shape)
for trial in range(3):
    a = sem.lapl(u, gm); b = so.lapl(u, om)
    err = np.abs(a - b) / np.max(np.abs(b))
    bad = np.argwhere(err > 1e-12)
    print("trial", trial, "max err", err.max(), "nbad", len(bad))
    if len(bad):
        print("bad x range", bad[:, 0].min(), bad[:, 0].max(), "y range", bad[:, 1].min(), bad[:, 1].max())
        ys = np.unique(bad[:, 1]); print("bad rows", ys[:40])
        xs = np.unique(bad[:, 0]); print("bad cols", xs[:40])
