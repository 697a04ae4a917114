"""The reference-side binding julia/rhs_b200.jl cannot run in this image (no Julia).  What can be checked: every field of
`params`, `mesh`, `metrics`, the AssemblerCache and PhysicalConst, and every `inputs[:key]` the shim READS exists in the
reference (names extracted from /root/reference by tests/golden/make_reference_names.py into tests/golden/reference_names.json;
the GPU box has no reference tree), and every C symbol it `ccall`s is declared in include/jexrhs.h with the same arity."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = open(os.path.join(ROOT, "julia", "rhs_b200.jl"), encoding="utf-8").read()
NAMES = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_names.json"), encoding="utf-8"))
CODE = "\n".join(line.split("#")[0] for line in SHIM.split("\n"))      # comments stripped
ID = r"[A-Za-zξηζψωγμκενρλδΔ_][\wξηζψωγμκενρλδΔ]*"


def test_params_fields_exist_in_reference():
    used = set(re.findall(r"\bparams\.(" + ID + ")", CODE))
    assert {"mesh", "metrics", "basis", "neqs", "ω", "Minv", "qp", "visc_coeff", "g_dss_cache"} <= used
    missing = used - set(NAMES["params"])
    assert not missing, f"julia/rhs_b200.jl reads params fields the reference does not have: {sorted(missing)}"


def test_inputs_keys_exist_in_reference_or_are_ours():
    used = set(re.findall(r"inputs\[:(" + ID + r")\]", CODE)) | set(re.findall(r"get\(inputs,\s*:(" + ID + ")", CODE)) \
        | set(re.findall(r"haskey\(inputs,\s*:(" + ID + ")", CODE))
    ours = {k for k in used if k.startswith("b200_")} | {"backend"}      # keys this binding introduces (documented in INTEGRATION.md)
    missing = used - ours - set(NAMES["inputs"])
    assert not missing, f"inputs keys neither set by the reference nor introduced by the binding: {sorted(missing)}"
    assert {"_parsed_equations", "energy_equation", "SOL_VARS_TYPE", "lsource", "lvisc", "ode_solver"} <= used


def test_struct_fields_exist_in_reference():
    for var, struct in (("mesh", "St_mesh"), ("metrics", "St_metrics"), ("cache", "AssemblerCache"), ("PC", "PhysicalConst")):
        used = set(re.findall(r"\b" + var + r"\.(" + ID + ")", CODE))
        assert used, var
        missing = used - set(NAMES[struct])
        assert not missing, f"{var}.* fields missing from the reference's {struct}: {sorted(missing)}"
    assert NAMES["rhs_jl_has_PHYS_CONST"]
    # the SGS binding reads the closure constants and the mesh's effective resolution / AMR levels
    assert {"Pr_t", "Sc_t", "μ_mol", "κ_mol", "Ri_crit", "C_s"} <= set(re.findall(r"\bPC\.(" + ID + ")", CODE))
    assert {"Δeffective_l", "ad_lvl"} <= set(re.findall(r"\bmesh\.(" + ID + ")", CODE))


def test_dispatch_tags_exist_in_reference():
    """`vm isa J.SMAG` etc.: the tag types the shim dispatches on are structs of src/kernel/abstractTypes.jl."""
    used = set(re.findall(r"\bisa\s+J\.(\w+)", CODE)) | set(re.findall(r"==\s*J\.(\w+)\(\)", CODE))
    assert {"SMAG", "VREM", "PERT"} <= used
    missing = used - set(NAMES["abstract_types"])
    assert not missing, sorted(missing)


def test_ccalls_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "jexrhs.h")).read()
    decl = {}
    for m in re.finditer(r"\b(?:int|void|int64_t)\s+(jx_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        decl[m.group(1)] = len(args)
    calls = re.findall(r"ccall\(\(:(jx_\w+),\s*LIB\),\s*\w+,\s*\(([^)]*)\)", CODE)
    assert calls
    for name, types in calls:
        assert name in decl, f"{name} is not declared in include/jexrhs.h"
        n = len([t for t in types.split(",") if t.strip()])
        assert n == decl[name], f"{name}: ccall passes {n} arguments, the header declares {decl[name]}"
