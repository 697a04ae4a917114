"""Extracts, from the reference tree, the names the Julia shim (julia/rhs_b200.jl) relies on, into
tests/golden/reference_names.json (the GPU box has no /root/reference):
  * the fields of the `params` NamedTuple of the non-Laguerre branch (params_setup.jl:442-479),
  * every `inputs[:key]` the reference itself assigns or defaults (mod_inputs.jl, run.jl),
  * the fields of St_mesh / St_metrics / the Lagrange basis / St_SolutionVars / AssemblerCache / PhysicalConst the shim reads."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def read(p):
    return open(os.path.join(REF, p), encoding="utf-8").read()


def params_fields():
    src = read("src/kernel/infrastructure/params_setup.jl")
    i = src.rindex("params = (backend,")
    j = src.index("coupling = coupling)", i)
    body = src[i + len("params = ("):j + len("coupling = coupling")]
    names = set()
    for item in body.replace("\n", " ").split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            names.add(item.split("=")[0].strip())
        else:
            names.add(item.split(".")[-1].strip())
    return sorted(names)


def inputs_keys():
    keys = set()
    for p in ("src/io/mod_inputs.jl", "src/run.jl"):
        keys |= set(re.findall(r"inputs\[:([A-Za-z_Δμδ][A-Za-z0-9_Δμδ]*)\]\s*=", read(p)))
        keys |= set(re.findall(r"haskey\(inputs,\s*:([A-Za-z_Δμδ][A-Za-z0-9_Δμδ]*)\)", read(p)))
    return sorted(keys)


def struct_fields(path, struct):
    src = read(path)
    m = re.search(r"struct\s+" + re.escape(struct) + r"\b.*?\n(.*?)\nend", src, re.S)
    names = set()
    for line in m.group(1).split("\n"):
        line = line.split("#")[0].strip()
        mm = re.match(r"([A-Za-zξηζψωγμκενρλδΔ_][\wξηζψωγμκενρλδΔ]*)\s*(::|=)", line)
        if mm:
            names.add(mm.group(1))
    return sorted(names)


out = {
    "params": params_fields(),
    "inputs": inputs_keys(),
    "St_mesh": struct_fields("src/kernel/mesh/meshStructs.jl", "St_mesh"),
    "St_metrics": struct_fields("src/kernel/mesh/metric_terms.jl", "St_metrics"),
    "AssemblerCache": struct_fields("src/kernel/mpi/mpi_communications.jl", "AssemblerCache"),
    "PhysicalConst": struct_fields("src/kernel/physics/globalConstantsPhysics.jl", "PhysicalConst"),
    "abstract_types": sorted(set(re.findall(r"^struct\s+(\w+)", read("src/kernel/abstractTypes.jl"), re.M))),
    "rhs_jl_has_PHYS_CONST": "const PHYS_CONST = PhysicalConst{Float64}()" in read("src/kernel/operators/rhs.jl"),
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_names.json"), "w"), indent=1, ensure_ascii=False, sort_keys=True)
print({k: (len(v) if isinstance(v, list) else v) for k, v in out.items()})
