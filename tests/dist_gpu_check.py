"""Multi-GPU parity check, one process per GPU (run under torchrun by test_gpu_multi.py):
y-slab partition + NCCL halo exchange + gathered PCG scalars against the single-domain CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import torch.distributed as dist

import sem_oracle as so
import spectralelements_jl_b200 as sem


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sem.init(local)
    ctx.comm_init_torch()
    fails = []
    cases = [(9, 8, 8, (False, False), so.wavy, "DDDD"), (8, 5, 6, (False, True), so.annulus, "DDNN"),
             (6, 40, 5, (True, True), so.wavy, "NNNN"), (5, 7, 9, (False, False), so.wavy, "DDDD"),
             (9, 60, 3 * world + 1, (False, False), so.wavy, "DDDD")]   # three strips, several chunks per slab
    if world > 5:
        cases = [c for c in cases if c[2] >= world] + [(9, 8, 2 * world, (False, True), so.wavy, "DDNN")]
    # one element row per rank (first row == last row of the slab), with and without the periodic wrap over the ranks
    cases += [(9, 8, world, (False, False), so.wavy, "DDDD"), (7, 30, world, (False, True), so.wavy, "DDNN")]
    only = os.environ.get("SEMB_DIST_CASES")   # e.g. "5,6": run a subset (debugging)
    if only:
        cases = [cases[int(i)] for i in only.split(",")]
    for nr, Ex, Ey, per, deform, bc in cases:
        if Ey < world:
            continue
        om = so.make_mesh(nr, nr, Ex, Ey, per, deform)
        e0, ne = sem.partition(Ey, world, rank)
        sl = slice(e0 * nr, (e0 + ne) * nr)
        loc = lambda a: np.asfortranarray(a[:, sl])
        gm = sem.Mesh.from_arrays(nr, nr, Ex, Ey, per, om.Dr, om.Ds, loc(om.G11), loc(om.G12), loc(om.G22), loc(om.B),
                                  ctx=ctx)
        tag = "nr=%d %dx%d per=%s bc=%s" % (nr, Ex, Ey, per, bc)
        if rank == 0 and os.environ.get("SEMB_DIST_VERBOSE"):
            print("case", tag, flush=True)
        u = so.splitmix_uniform(om.x.shape, seed=21)
        M = so.generateMask(list(bc), om).astype(np.float64)
        if not np.array_equal(gm.mult, loc(om.mult)):
            fails.append(tag + " mult")
        if not np.array_equal(sem.gatherScatter(loc(u), gm), loc(so.gatherScatter(u, om))):
            fails.append(tag + " gatherScatter not bit-exact")
        if not np.array_equal(sem.generateMask(list(bc), gm), loc(so.generateMask(list(bc), om))):
            fails.append(tag + " mask")
        e = relerr(sem.OpLHS(gm, 1.0, 0.7, bc=bc)(loc(u)), loc(so.opLHS(u, 1.0, 0.7, M, om)))
        if e > 1e-12:
            fails.append(tag + " opLHS %g" % e)
        if Ey // world >= 2:   # every chunking of the slab gives the same bits (2-term interface sums); an apply is
            # collective (halo exchange), so the condition must not depend on the rank's own slab
            ref_bits = sem.OpLHS(gm, 1.0, 0.0, bc=bc)(loc(u))
            for nch in (1, gm.ney):
                gm.set_chunks(nch)
                if not np.array_equal(sem.OpLHS(gm, 1.0, 0.0, bc=bc)(loc(u)), ref_bits):
                    fails.append(tag + " opLHS bits change with %d chunks" % nch)
            gm.set_chunks(max(1, gm.ney // 2))
        try:
            gm.peer_status()   # a timed-out peer wait inside the apply kernels shows up here, not three calls later
        except Exception as ex:
            fails.append(tag + " " + str(ex)[:120])
        # device random fill uses the GLOBAL index: the slabs tile the single-domain stream
        if not np.array_equal(gm.field().fill_random(5).download(), loc(so.splitmix_uniform(om.x.shape, seed=5))):
            fails.append(tag + " fill_random")
        fa, fb = gm.field(loc(u)), gm.field(loc(M * u))
        ref = float(np.sum(u * (M * u) * om.mult))
        if abs(gm.dot_mult(fa, fb) - ref) > 1e-12 * abs(ref) or gm.norm_inf(fa) != float(np.max(np.abs(u))):
            fails.append(tag + " reductions")
        b = so.gatherScatter(so.mask(so.mass(np.ones(om.x.shape), om), M), om)
        io, ig = {}, {}
        kk = 0.7 if bc == "NNNN" else 0.0
        xo = so.pcg(b, lambda v: so.opLHS(v, 1.0, kk, M, om), mult=om.mult, tol=1e-12, info=io)
        xg = sem.pcg(loc(b), sem.OpLHS(gm, 1.0, kk, bc=bc), tol=1e-12, info=ig)
        e = relerr(xg, loc(xo)) if np.max(np.abs(loc(xo))) > 0 else 0.0
        if e > 1e-9 or abs(ig["iters"] - io["iters"]) > max(3, int(0.02 * io["iters"])):
            fails.append(tag + " pcg err %g iters %d vs %d" % (e, ig["iters"], io["iters"]))
        # FDM preconditioner on slabs: the halo tiles read the neighbour rank's rows of r through peer memory
        if not os.environ.get("SEMB_NO_P2P") and nr >= 4 and Ey // world >= 2:
            Po, Pg = so.fdm_schwarz(om, bc, 1.0, kk), sem.FdmPrecond(gm, bc, 1.0, kk)
            r = so.mask(so.gatherScatter(so.splitmix_uniform(om.x.shape, seed=8) * om.mult, om), M)
            ho = Po(r)
            hg = Pg(loc(r))
            e = float(np.max(np.abs(hg - loc(ho))) / np.max(np.abs(ho)))
            if e > 1e-12:
                fails.append(tag + " fdm apply %g" % e)
            if not np.array_equal(Pg(loc(r)), hg):
                fails.append(tag + " fdm apply not deterministic")
            io2, ig2 = {}, {}
            xo2 = so.pcg(b, lambda v: so.opLHS(v, 1.0, kk, M, om), opM=Po, mult=om.mult, tol=1e-10, info=io2)
            xg2 = sem.pcg(loc(b), sem.OpLHS(gm, 1.0, kk, bc=bc), opM=Pg, mult=gm.mult, tol=1e-10, info=ig2)
            e = relerr(xg2, loc(xo2)) if np.max(np.abs(loc(xo2))) > 0 else 0.0
            # (same bar as the unpreconditioned solves: exact below the rounding horizon, the oracle's own 2 % spread above)
            if e > 1e-8 or abs(ig2["iters"] - io2["iters"]) > (0 if io2["iters"] <= 60 else max(3, int(0.02 * io2["iters"]))):
                fails.append(tag + " pcg+fdm err %g iters %d vs %d" % (e, ig2["iters"], io2["iters"]))
        gm.free()
    # pipelined host twin on slabs (semb_oplhs_host: upload | compute | download by slabs, the two lines shared with the
    # neighbour ranks exchanged and downloaded last): same bits as the device-resident one-launch apply
    if not only:
        gm = sem.Mesh(9, 9, 256, 204 * world, (False, False), "wavy", ctx=ctx)
        uu = np.asfortranarray(np.random.default_rng(100 + rank).standard_normal(gm.shape))
        for kk in (0.0, 0.4):
            hh = sem.OpLHS(gm, 1.3, kk, bc="DNDD")(uu)
            fu, fo = gm.field(uu), gm.field()
            gm.oplhs_device(fu, fo, nu=1.3, k=kk, bc="DNDD")
            if not np.array_equal(hh, fo.download()):
                fails.append("pipelined host twin differs from the device apply (k=%g)" % kk)
            fu.free(); fo.free()
        try:
            gm.peer_status()
        except Exception as ex:
            fails.append("host twin: " + str(ex)[:120])
        gm.free()
    # Stokes split on slabs: element-local kernels + gatherScatter on both meshes + all-reduced PCG scalars
    for nr, Ex, Ey in [(7, 3, 8), (9, 4, 2 * world)]:
        if Ey < world:
            continue
        oV, oP = so.make_mesh(nr, nr, Ex, Ey, (False, False), so.wavy), so.make_mesh(nr - 2, nr - 2, Ex, Ey, (False, False), so.wavy)
        gV = sem.Mesh(nr, nr, Ex, Ey, (False, False), "wavy", deform_params=(0.1,), ctx=ctx)
        gP = sem.Mesh(nr - 2, nr - 2, Ex, Ey, (False, False), "wavy", deform_params=(0.1,), ctx=ctx)
        osk, gsk = so.make_stokes(list("DDDD"), list("DDNN"), oV, oP, 1.0), sem.Stokes("DDDD", "DDNN", gV, gP, 1.0)
        e0, ne = sem.partition(Ey, world, rank)
        locV = lambda a: np.asfortranarray(a[:, e0 * nr:(e0 + ne) * nr])
        locP = lambda a: np.asfortranarray(a[:, e0 * (nr - 2):(e0 + ne) * (nr - 2)])
        tag = "stokes nr=%d %dx%d" % (nr, Ex, Ey)
        q = so.splitmix_uniform(oP.x.shape, seed=3)
        e = relerr(sem.opStokesLHS(locP(q), gsk), locP(so.opStokesLHS(q, osk)))
        if e > 1e-11:
            fails.append(tag + " opStokesLHS %g" % e)
        vx = so.mask(so.gatherScatter(so.splitmix_uniform(oV.x.shape, seed=5) * oV.mult, oV), osk.Mvx)
        vy = so.mask(so.gatherScatter(so.splitmix_uniform(oV.x.shape, seed=6) * oV.mult, oV), osk.Mvy)
        ox, oy, op = so.pressureProject(vx, vy, np.zeros(oP.x.shape), osk, tol=1e-9)
        gx, gy, gp = sem.pressureProject(locV(vx), locV(vy), locP(np.zeros(oP.x.shape)), gsk, tol=1e-9)
        e = max(np.max(np.abs(gx - locV(ox))), np.max(np.abs(gy - locV(oy))))
        if e > 1e-6 or abs(gsk.pcg_iters[-1] - osk.pcg_iters[-1]) > max(3, int(0.02 * osk.pcg_iters[-1])):
            fails.append(tag + " project err %g iters %d vs %d" % (e, gsk.pcg_iters[-1], osk.pcg_iters[-1]))
        gsk.free(); gV.free(); gP.free()
    flag = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(flag)
    for f in fails:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("DIST_CHECK", "OK" if int(flag.item()) == 0 else "FAILED", "world", world, flush=True)
    sem.finalize()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
