"""CPU-side checks of the product: libsemb.so loads and exports every symbol include/semb.h declares,
its host (no-GPU) entry points agree with the oracle, it fails loudly without a GPU, and nothing
in the product imports the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import sem_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sem):
    lib = sem._lib.load()
    hdr = open(os.path.join(ROOT, "include", "semb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(semb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 60
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    # ... and the ctypes binding covers the whole header
    unbound = sorted(names - set(sem._lib._SIGS))
    assert not unbound, unbound
    assert lib.semb_version() >= 100


def test_host_setup_helpers_match_oracle(sem):
    for n in (2, 3, 5, 8, 9, 13, 17):
        z, w = sem.gausslobatto(n)
        zo, wo = so.gausslobatto(n)
        assert np.max(np.abs(z - zo)) < 1e-15 and np.max(np.abs(w - wo)) < 1e-15
        D, Do = sem.derivMat(z), so.derivMat(zo)
        assert np.max(np.abs(D - Do)) < 1e-13 * np.max(np.abs(Do))
        J = sem.interpMat(so.gausslobatto(n + 4)[0], zo)
        assert np.max(np.abs(J - so.interpMat(so.gausslobatto(n + 4)[0], zo))) < 1e-14
    for E, n in ((1, 4), (7, 9), (1112, 9)):
        z, w = sem.semmesh(E, n)
        zo, wo = so.semmesh(E, n)
        assert np.max(np.abs(z - zo)) < 3e-16 and np.max(np.abs(w - wo)) < 1e-16
    for t in ([0.0] * 4, [0.01, 0, 0, 0], [0.02, 0.01, 0, 0], [0.03, 0.02, 0.01, 0.0], [0.5, 0.3, 0.2, 0.1]):
        a, b = sem.bdfExtK(t)
        ao, bo = so.bdfExtK(np.array(t))
        assert np.allclose(a, ao, rtol=1e-13, atol=1e-13) and np.allclose(b, bo, rtol=1e-13, atol=1e-11)


def test_partition_is_a_contiguous_tiling(sem):
    for Ey in (1, 2, 7, 8, 1112, 1113):
        for P in (1, 2, 3, 4, 8):
            if P > Ey:
                with pytest.raises(sem.SembError):
                    sem.partition(Ey, P, 0)
                continue
            nxt = 0
            for r in range(P):
                e0, ne = sem.partition(Ey, P, r)
                assert e0 == nxt and ne >= Ey // P
                nxt = e0 + ne
            assert nxt == Ey


def test_no_gpu_means_loud_failure(sem):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(sem.SembError) as ei:
        sem.Context(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "spectralelements.jl_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, fn), errors="replace").read()
                assert "sem_oracle" not in src and "oracle/" not in src, fn
    shim = open(os.path.join(ROOT, "spectralelements_jl_b200.py")).read()
    assert "oracle" not in shim


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--elements", "64", "--cpu-rows", "8"], capture_output=True, text=True, timeout=300)
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


def test_plan_chunks_rule(sem):
    """Launch-plan host logic (semb_plan_chunks, used by mesh_build_plan): the counts the chunk sweeps on the B200 found
    best (profiles/r01_sweep_chunks_r1l.txt, 296 CTA slots = 148 SMs x 2), and the invariants of any plan."""
    import ctypes as C
    lib = sem._lib.load()

    def pick(nstrips, ney, slots=296):
        n = C.c_int()
        assert lib.semb_plan_chunks(nstrips, ney, slots, C.byref(n)) == 0
        return n.value

    measured = {  # (strips, element rows) -> best measured chunk count
        (40, 1112): 22,   # order 8, 1112x1112 (headline): 736 us; 7/14/29/36/44 chunks: 784-795 us
        (19, 512): 31,    # order 8, 512x512 (cfg4)
        (10, 256): 29,    # order 8, 256x256 (cfg2): 59.4 us vs 66.5 us at 59 chunks
        (12, 256): 24,    # order 10, 256x256 (cfg5 velocity mesh): 84.5 us vs 92.8 us at 74
        (56, 776): 21,    # order 12, 776x776 (cfg3)
        (40, 139): 7,     # 1/8 strong-scaling slab of the headline mesh: 113.4 us vs 118.3 us at 22
        (5, 128): 59,
        (3, 64): 64,      # fewer rows than slots per strip: one chunk per element row
    }
    for (nstrips, ney), want in measured.items():
        assert pick(nstrips, ney) == want, (nstrips, ney)
    rng = np.random.default_rng(3)
    for _ in range(300):
        nstrips, ney, slots = int(rng.integers(1, 200)), int(rng.integers(1, 5000)), int(rng.integers(1, 9)) * 148
        nc = pick(nstrips, ney, slots)
        assert 1 <= nc <= ney
        lo = max(1, slots // nstrips)
        if ney > 2 * lo:  # large slab: at least one wave's worth of CTAs, at most four, >= 2 element rows per chunk
            assert lo <= nc <= max(lo, 4 * slots // nstrips) and nc <= ney // 2
    assert lib.semb_plan_chunks(0, 10, 296, C.byref(C.c_int())) < 0   # bad arguments are errors, not crashes
    assert b"semb_plan_chunks" in lib.semb_last_error()
