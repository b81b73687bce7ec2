"""Fixed cost of one semb_pcg call (graph capture / instantiate, allocations) vs its per-iteration cost:
wall-clock of whole solves with maxiter = 16, 32, 64, 128 (tol = 0) on a cfg4-size mesh (diag preconditioner, k != 0).
    python tools/pcg_fixed_cost.py [nr E]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spectralelements_jl_b200 as sem

nr = int(sys.argv[1]) if len(sys.argv) > 1 else 9
E = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ctx = sem.init(0)
msh = sem.Mesh(nr, nr, E, E, (True, False), "identity", ctx=ctx)
b, x = msh.field().fill_random(1), msh.field()
for rep in range(2):
    for its in (1, 16, 32, 64, 128):
        ctx.sync()
        t0 = time.perf_counter()
        info = msh.pcg_device(b, x, nu=1e-3, k=366.7, bc="NNDD", precond=True, prec_b0=366.7, tol=0.0, maxiter=its)
        ctx.sync()
        print("graph=%s maxiter %4d: %.3f ms  %s" % (os.environ.get("SEMB_NO_GRAPH", "0") != "1", its, (time.perf_counter() - t0) * 1e3, info))
