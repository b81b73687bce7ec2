"""world_size-N CPU (gloo) execution of the multi-GPU protocol of libsemb (semb_api.cu run_operator /
halo_exchange / gather_scalars), with the oracle as the per-slab arithmetic: slab partition and halo
plan come from the product's C ABI (semb_partition, semb_halo_plan); each rank computes its slab's
local operator, forms x pairs, exchanges ONE boundary row per neighbour in the product's message order,
forms y pairs, masks; PCG scalars are all-gathered and combined in rank order.  Must reproduce the
single-domain oracle BIT FOR BIT (operator) / to rounding (PCG)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import torch.distributed as dist

import sem_oracle as so
import spectralelements_jl_b200 as sem


def slab_oplhs(u_loc, loc, nu, k, om, nr, ns, Ex, per, plan, Mloc):
    """One fused apply on this rank's slab, following semb_api.cu:run_operator step by step."""
    halo_lo, halo_hi, rank_lo, rank_hi = plan
    Au = nu * so.laplace(u_loc, om.Dr, om.Ds, loc(om.G11), loc(om.G12), loc(om.G22)) + k * (loc(om.B) * u_loc)
    ney = u_loc.shape[1] // ns
    v = so.gatherScatter_index(Au, nr, ns, Ex, 1, (per[0], False))  # x pairs only (strip + x-seam kernels)
    # in-slab y pairs (register carry + chunk seams); single rank: the periodic wrap is local
    v = np.asfortranarray(v)
    if ney > 1:
        a = np.arange(1, ney) * ns - 1
        s = v[:, a] + v[:, a + 1]
        v[:, a] = s
        v[:, a + 1] = s
    if per[1] and dist.get_world_size() == 1:
        s = v[:, -1] + v[:, 0]
        v[:, 0] = s
        v[:, -1] = s
    # halo exchange: send(last row -> hi), send(first row -> lo), recv(lo), recv(hi)  [product order]
    first, last = torch.from_numpy(np.ascontiguousarray(v[:, 0])), torch.from_numpy(np.ascontiguousarray(v[:, -1]))
    rlo, rhi = torch.empty_like(first), torch.empty_like(first)
    ops = []
    if halo_hi:
        ops.append(dist.P2POp(dist.isend, last, rank_hi))
    if halo_lo:
        ops.append(dist.P2POp(dist.isend, first, rank_lo))
    if halo_lo:
        ops.append(dist.P2POp(dist.irecv, rlo, rank_lo))
    if halo_hi:
        ops.append(dist.P2POp(dist.irecv, rhi, rank_hi))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if halo_lo:
        v[:, 0] = v[:, 0] + rlo.numpy()
    if halo_hi:
        v[:, -1] = v[:, -1] + rhi.numpy()
    return Mloc * v


def gathered_sum(x):
    """all-gather one double per rank, combine in rank order (identical bits on every rank)"""
    t = [torch.zeros(1, dtype=torch.float64) for _ in range(dist.get_world_size())]
    dist.all_gather(t, torch.tensor([x], dtype=torch.float64))
    s = 0.0
    for v in t:
        s += float(v.item())
    return s


def gathered_max(x):
    t = [torch.zeros(1, dtype=torch.float64) for _ in range(dist.get_world_size())]
    dist.all_gather(t, torch.tensor([x], dtype=torch.float64))
    return max(float(v.item()) for v in t)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fails = []
    for nr, Ex, Ey, per, deform, bc in [(5, 3, 4, (False, False), so.wavy, "DDDD"), (4, 2, 2, (True, True), so.wavy, "NNNN"),
                                        (6, 2, 5, (False, True), so.annulus, "DDNN")]:
        if Ey < world:
            continue
        om = so.make_mesh(nr, nr, Ex, Ey, per, deform)
        e0, ne = sem.partition(Ey, world, rank)          # product C ABI
        plan = sem.halo_plan(world, rank, per[1])          # product C ABI
        sl = slice(e0 * nr, (e0 + ne) * nr)
        loc = lambda a: np.asfortranarray(a[:, sl])
        M = so.generateMask(list(bc), om).astype(np.float64)
        u = so.splitmix_uniform(om.x.shape, seed=3)
        ref = so.opLHS(u, 0.9, 0.4, M, om)
        out = slab_oplhs(loc(u), loc, 0.9, 0.4, om, nr, nr, Ex, per, plan, loc(M))
        if not np.array_equal(out, loc(ref)):
            fails.append("opLHS slab != single domain (nr=%d %dx%d per=%s): %g" % (nr, Ex, Ey, per, np.max(np.abs(out - loc(ref)))))
        # distributed PCG with gathered scalars (pcg.jl:16-60 on slabs)
        b = loc(so.gatherScatter(so.mask(so.mass(np.ones(om.x.shape), om), M), om))
        kk = 0.4 if bc == "NNNN" else 0.0
        x, r, p = np.zeros_like(b), b.copy(), np.zeros_like(b)
        mult = loc(om.mult)
        t_prev, it = 0.0, 0
        t = gathered_sum(float(np.sum(r * r * mult)))
        while gathered_max(float(np.max(np.abs(r)))) > 1e-10 and it < 500:
            beta = 0.0 if it == 0 else t / t_prev
            p = r + beta * p
            Ap = slab_oplhs(p, loc, 1.0, kk, om, nr, nr, Ex, per, plan, loc(M))
            alpha = t / gathered_sum(float(np.sum(p * Ap * mult)))
            x = x + alpha * p
            r = r - alpha * Ap
            t_prev, t = t, gathered_sum(float(np.sum(r * r * mult)))
            it += 1
        io = {}
        xo = so.pcg(so.gatherScatter(so.mask(so.mass(np.ones(om.x.shape), om), M), om),
                    lambda v: so.opLHS(v, 1.0, kk, M, om), mult=om.mult, tol=1e-10, info=io)
        err = np.max(np.abs(x - loc(xo))) / max(np.max(np.abs(xo)), 1e-300)
        if err > 1e-8 or abs(it - io["iters"]) > max(3, 0.05 * io["iters"]):
            fails.append("pcg slabs vs single domain: err %g iters %d vs %d" % (err, it, io["iters"]))
    flag = torch.tensor([len(fails)])
    dist.all_reduce(flag)
    for f in fails:
        print("[rank %d] FAIL %s" % (rank, f), flush=True)
    if rank == 0:
        print("DIST_CPU", "OK" if int(flag.item()) == 0 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
