"""Diagnostic: per-iteration norm(r,Inf) of the device PCG against the CPU oracle (same mesh, same rhs).
Shows how rounding-level differences (reduction order, FMA contraction) grow along the CG trajectory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import sem_oracle as so
import spectralelements_jl_b200 as sem

ctx = sem.init(0)
nr, E, deform = int(sys.argv[1]) if len(sys.argv) > 1 else 9, int(sys.argv[2]) if len(sys.argv) > 2 else 8, so.wavy
om = so.make_mesh(nr, nr, E, E, (False, False), deform)
gm = sem.Mesh.from_arrays(nr, nr, E, E, (False, False), om.Dr, om.Ds, om.G11, om.G12, om.G22, om.B, ctx=ctx)
M = so.generateMask(list("DDDD"), om).astype(np.float64)
b = so.gatherScatter(so.mask(so.mass(np.ones(gm.shape), om), M), om)
io = {}
xo = so.pcg(b, lambda v: so.opLHS(v, 1.0, 0.0, M, om), mult=om.mult, info=io)
# a second oracle run whose operator is evaluated in a mathematically equivalent but differently rounded way
def op2(v):
    Au = so.laplace(v, om.Dr, om.Ds, om.G11, om.G12, om.G22)
    Au = Au + 0.0
    return so.mask(so.gatherScatter_index(Au * (1.0 + 2.3e-16), om.nr, om.ns, om.Ex, om.Ey, om.ifperiodic), M)
io2 = {}
so.pcg(b, op2, mult=om.mult, info=io2)
fb, fx = gm.field(b), gm.field()
gm.pcg_begin(fb, fx, bc="DDDD", tol=1e-8)
hist = [gm.pcg_status()[1]]
while True:
    gm.pcg_iterate(1)
    it, res, done = gm.pcg_status()
    if it + 1 > len(hist):
        hist.append(res)
    if done:
        break
print("iters: oracle %d, oracle(1-ulp perturbed op) %d, gpu %d" % (io["iters"], io2["iters"], len(hist) - 1))
ho, h2 = np.array(io["hist"]), np.array(io2["hist"])
n = min(len(ho), len(hist), len(h2))
for k in list(range(0, min(n, 12))) + list(range(12, n, max(1, n // 25))):
    print("k=%4d  r_oracle=%.6e  rel.diff gpu=%.2e  rel.diff oracle2=%.2e" % (k, ho[k], abs(hist[k] - ho[k]) / ho[k], abs(h2[k] - ho[k]) / ho[k]))
