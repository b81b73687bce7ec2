"""Small run through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import spectralelements_jl_b200 as sem

ctx = sem.init(0)
rng = np.random.default_rng(0)
for nr, Ex, Ey, per in ((9, 30, 3, (False, False)), (8, 33, 2, (True, True)), (13, 15, 2, (False, True)), (4, 70, 3, (True, False)),
                        (17, 3, 2, (False, False)), (2, 5, 4, (False, False))):
    m = sem.Mesh(nr, nr, Ex, Ey, per, sem.wavy, ctx=ctx)
    u = np.asfortranarray(rng.standard_normal(m.shape))
    nu = 1.0 + 0.1 * u ** 2
    sem.lapl(u, m); sem.hlmz(u, nu, 0.5, m); sem.mass(u, m); sem.gatherScatter(u, m)
    M = sem.generateMask(list("DDNN"), m)
    sem.mask(u, M, m)
    for nch in (1, Ey):
        m.set_chunks(nch)
        sem.OpLHS(m, 1.0, 0.0, bc="DDNN")(u)
        sem.OpLHS(m, nu, 1.0, M=M.astype(float))(u)
    b = sem.gatherScatter(sem.mask(sem.mass(np.ones(m.shape), m), M, m), m)
    info = {}
    sem.pcg(b, sem.OpLHS(m, 1.0, 0.3, bc="DDNN"), maxiter=20, info=info)
    sem.pcg(b, sem.OpLHS(m, 0.01, 300.0, bc="DDNN"), opM=sem.DiagPrecond(m, 300.0), maxiter=20, info=info)
    fa = m.field(u); m.dot_mult(fa, fa); m.norm_inf(fa)
    if nr >= 3 and not (per[0] and Ex < 2) and not (per[1] and Ey < 2):   # FDM preconditioner: stand-alone and inside pcg
        P = sem.FdmPrecond(m, "DDNN", 1.0, 0.3)
        P(sem.mask(sem.gatherScatter(u * m.mult, m), M, m))
        sem.pcg(b, sem.OpLHS(m, 1.0, 0.3, bc="DDNN"), opM=P, mult=m.mult, maxiter=10, info=info)
    sem.grad(u, m)
    m.free()
# generic path, ABu, drivers
g = sem.Mesh(5, 7, 3, 2, (False, False), sem.wavy, ctx=ctx)
ug = np.asfortranarray(rng.standard_normal(g.shape))
sem.OpLHS(g, 1.0, 0.2, bc="DDDD")(ug); sem.gatherScatter(ug, g)
sem.pcg(sem.gatherScatter(ug, g), sem.OpLHS(g, 1.0, 1.0, bc="DDDD"), maxiter=5)
g.free()
sem.ABu(rng.standard_normal((4, 6)), rng.standard_normal((5, 8)), np.asfortranarray(rng.standard_normal((24, 18))), ctx=ctx)
mV, mD = sem.Mesh(6, 6, 4, 3, (True, False), ctx=ctx), sem.Mesh(9, 9, 4, 3, (True, False), ctx=ctx)
cd = sem.ConvectionDiffusion("ps", list("NNDD"), mV, mD, 0 * mV.x + 1.0, 0 * mV.x, Tf=1.0, dt=1e-2,
                             setNu=lambda x, y, t: 1e-3 + 0 * x)
cd.u = np.sin(np.pi * mV.x) * np.sin(np.pi * mV.y)
for _ in range(3):
    sem.step_b(cd)
d = sem.Diffusion(list("DDDD"), mV, Tf=1.0, dt=0.1)
d.nu = 1.0 + 0 * mV.x; d.f = 1.0 + 0 * mV.x
sem.evolve_b(d)
cd.free(); d.free(); mV.free(); mD.free()
# register-tiled advection (ragged batches), Stokes element kernels, one-pass gatherScatter, explicit-argument forms
for nr, nrd, Ex in ((9, 14, 11), (8, 12, 13), (5, 8, 17)):
    aV, aD = sem.Mesh(nr, nr, Ex, 2, (False, False), sem.wavy, ctx=ctx), sem.Mesh(nrd, nrd, Ex, 2, (False, False), sem.wavy, ctx=ctx)
    T = np.asfortranarray(rng.standard_normal(aV.shape))
    sem.advect(T, 1.0 + 0.3 * aV.x, np.cos(aV.y), aV, aD)
    aV.free(); aD.free()
for nr, Ex in ((11, 13), (7, 19), (4, 33)):
    sV, sP = sem.Mesh(nr, nr, Ex, 2, (False, False), sem.wavy, ctx=ctx), sem.Mesh(nr - 2, nr - 2, Ex, 2, (False, False), sem.wavy, ctx=ctx)
    sk = sem.Stokes("DDDD", "DDNN", sV, sP, 1.0)
    q = np.asfortranarray(rng.standard_normal(sP.shape))
    v = np.asfortranarray(rng.standard_normal(sV.shape))
    sem.opStokesLHS(q, sk); sem.diver(v, v, sk); sem.diverT(q, sk); sem.gradT(v, sV)
    sem.pressureProject(sem.mask(v, sem.generateMask(list("DDDD"), sV), sV), 0 * v, 0 * q, sk, tol=1e-6, maxiter=5)
    sk.free(); sV.free(); sP.free()
eV = sem.Mesh(5, 5, 3, 2, (False, False), sem.wavy, ctx=ctx)
ue = np.asfortranarray(rng.standard_normal(eV.shape))
Dr = np.asfortranarray(eV.Dr)
sem.laplace(ue, Dr, Dr, eV.G11, eV.G12, eV.G22)
sem.mass(ue, [], eV.B, [], [], [], [], eV.mult)
eV.free()
sem.finalize()
print("sanitize smoke done")
