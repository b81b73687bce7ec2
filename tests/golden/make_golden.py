"""Generate tests/golden/*.npz from the CPU oracle (oracle/sem_oracle.py).

The reference has no golden vectors and cannot be executed here (Julia missing), so these fixtures pin
the ORACLE (drift detector) and give the GPU tests committed input/output pairs.  tools/ref_dump.jl
writes the same arrays from the real Julia reference for anyone who has it.
    python tests/golden/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sem_oracle as so

CASES = {
    "p2d_annulus_5x5_nr8": dict(nr=8, Ex=5, Ey=5, per=(False, True), deform="annulus", bc="DDNN", nu=1.0, k=0.0),
    "cfg1_wavy_8x8_nr9": dict(nr=9, Ex=8, Ey=8, per=(False, False), deform="wavy", bc="DDDD", nu=1.0, k=0.0),
    "helmholtz_wavy_6x4_nr9": dict(nr=9, Ex=6, Ey=4, per=(False, False), deform="wavy", bc="DDDD", nu=0.7, k=1.3),
    "order12_wavy_3x3_nr13": dict(nr=13, Ex=3, Ey=3, per=(True, False), deform="wavy", bc="NNDD", nu=1.0, k=0.5),
}
DEF = {"annulus": so.annulus, "wavy": so.wavy, "box": so.fixU}


def build(c):
    m = so.make_mesh(c["nr"], c["nr"], c["Ex"], c["Ey"], c["per"], DEF[c["deform"]])
    M = so.generateMask(list(c["bc"]), m).astype(np.float64)
    u = so.splitmix_uniform(m.x.shape, seed=0x5EED)
    f = np.ones(m.x.shape)
    b = so.gatherScatter(so.mask(so.mass(f, m), M), m)
    info = {}
    x = so.pcg(b, lambda v: so.opLHS(v, c["nu"], c["k"], M, m), mult=m.mult, tol=1e-8, info=info)
    x12 = so.pcg(b, lambda v: so.opLHS(v, c["nu"], c["k"], M, m), mult=m.mult, tol=1e-12)
    # FDM preconditioner: h = opM(r) for a continuous masked r, and the preconditioned iteration count
    P = so.fdm_schwarz(m, c["bc"], c["nu"], c["k"])
    fr = so.mask(so.gatherScatter(u * m.mult, m), M)
    finfo = {}
    so.pcg(b, lambda v: so.opLHS(v, c["nu"], c["k"], M, m), opM=P, mult=m.mult, tol=1e-8, info=finfo)
    return dict(fdm_r=fr, fdm_h=P(fr), pcg_fdm_iters=np.array(finfo["iters"]),
                G11=m.G11, G12=m.G12, G22=m.G22, B=m.B, mult=m.mult, Dr=m.Dr, x=m.x, y=m.y, M=M, u=u,
                lapl=so.lapl(u, m), hlmz=so.hlmz(u, c["nu"], c["k"], m), gs=so.gatherScatter(u, m),
                oplhs=so.opLHS(u, c["nu"], c["k"], M, m), rhs=b, pcg_x=x, pcg_x_tol12=x12,
                pcg_iters=np.array(info["iters"]), pcg_hist=np.array(info["hist"][:13]))


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, c in CASES.items():
        np.savez_compressed(os.path.join(out, name + ".npz"), **build(c))
        print("wrote", name)
