"""CPU-side checks of the drop-in boundary: libjexrhs.so loads, exports every symbol that
include/jexrhs.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest

from jexpresso_b200 import capi


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(capi.LIB_PATH)


def test_header_symbols_exported(built_lib):
    names = capi.declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in include/jexrhs.h but not exported"


def test_exports_are_extern_c(built_lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for n in capi.declared_symbols():
        assert n in exported


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.JexError) as e:
        capi.Context(device=0)
    assert e.value.code == capi.JX_ENODEV


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under jexpresso_b200/ may import, link or load it."""
    root = os.path.join(os.path.dirname(__file__), "..", "jexpresso_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import oracle", "from oracle", "libjexref", "jexref.c\"", "oracle/_build", "oracle/_ref"):
                    assert needle not in txt, (f, needle)


def test_params_setup_refuses_a_rank_local_effective_delta():
    """mesh.Δeffective_l is a maximum over ALL ranks (mesh.jl:5629-5632); a closure run on several ranks must be handed that value,
    never this rank's own -- checked before anything touches the GPU."""
    import pytest
    from jexpresso_b200 import rhs as jrhs
    from jexpresso_b200.sem.problems import box3d, euler_case
    sems, qns, qes, us = euler_case(box3d((4, 2, 2), 2), 2, lpert=False)
    inputs = {"SOL_VARS_TYPE": "TOTAL", "lsource": True, "lvisc": True, "mu": [0.0, 1.0, 1.0, 1.0, 2.0], "visc_model": "VREM"}

    class _Ctx:                      # stands in for capi.Context up to the point of the check
        def __init__(self, *a, **k): pass
        def set_option(self, *a): pass
        def set_problem(self, *a): pass
        def close(self): pass

    real = jrhs.capi.Context
    jrhs.capi.Context = _Ctx
    try:
        with pytest.raises(ValueError, match="GLOBAL"):
            jrhs.params_setup(sems[0], qes[0], inputs, rank=0, nranks=2, nccl_uid=b"\0" * 128)
    finally:
        jrhs.capi.Context = real
