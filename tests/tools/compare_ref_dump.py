"""Compare a true-reference dump (tools/ref_dump.jl, run where Julia exists) with the CPU oracle, and -- when a GPU
and libsemb.so are available -- with the CUDA path.  This is how "parity unpinned" gets pinned.
    python tests/tools/compare_ref_dump.py ref_dump_dir [--gpu]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")]
import numpy as np
import sem_oracle as so
from make_golden import CASES, build


def load(d, name, key, shape):
    return np.fromfile(os.path.join(d, "%s.%s.f64" % (name, key)), dtype="<f8").reshape(shape, order="F")


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    d = sys.argv[1]
    use_gpu = "--gpu" in sys.argv
    for name, c in CASES.items():
        if not os.path.exists(os.path.join(d, name + ".u.f64")):
            continue
        shape = (c["nr"] * c["Ex"], c["nr"] * c["Ey"])
        o = build(c)
        print("==", name)
        for key in ("G11", "G12", "G22", "B", "mult", "lapl", "hlmz", "gs", "oplhs", "rhs", "pcg_x_tol12"):
            ref = load(d, name, key, shape)
            scale = np.max(np.abs(load(d, name, "G11", shape))) if key == "G12" else None
            e = np.max(np.abs(o[key] - ref)) / scale if scale else rel(o[key], ref)
            print("  oracle vs Julia  %-12s %.3e" % (key, e))
        if use_gpu:
            import spectralelements_jl_b200 as sem
            G = [load(d, name, k, shape) for k in ("G11", "G12", "G22", "B")]
            m = so.make_mesh(c["nr"], c["nr"], c["Ex"], c["Ey"], c["per"])
            gm = sem.Mesh.from_arrays(c["nr"], c["nr"], c["Ex"], c["Ey"], c["per"], m.Dr, m.Ds, *G)
            u = load(d, name, "u", shape)
            print("  CUDA  vs Julia  lapl         %.3e" % rel(sem.lapl(u, gm), load(d, name, "lapl", shape)))
            print("  CUDA  vs Julia  oplhs        %.3e" % rel(sem.OpLHS(gm, c["nu"], c["k"], bc=c["bc"])(u), load(d, name, "oplhs", shape)))
            x = sem.pcg(load(d, name, "rhs", shape), sem.OpLHS(gm, c["nu"], c["k"], bc=c["bc"]), tol=1e-12)
            print("  CUDA  vs Julia  pcg_x_tol12  %.3e" % rel(x, load(d, name, "pcg_x_tol12", shape)))
            gm.free()


if __name__ == "__main__":
    main()
