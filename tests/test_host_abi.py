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
