"""The Julia shim (rayuela.jl_b200/julia/RayuelaB200.jl) cannot be executed in this image (no Julia), so its `ccall`s
are checked statically: every symbol it binds must be declared in include/rayuela_b200.h with the same number of
arguments and, argument by argument, the same ABI class (pointer vs integer, width), and the same return class."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "rayuela.jl_b200", "julia", "RayuelaB200.jl")
HEADER = os.path.join(ROOT, "include", "rayuela_b200.h")

JL = {"Cdouble": ("float", 8), "Cint": ("int", 4), "Cuint": ("int", 4), "Int64": ("int", 8), "UInt64": ("int", 8), "Cfloat": ("float", 4),
      "Cstring": ("ptr", 8), "Nothing": ("void", 0)}


def jl_class(t):
    t = t.strip()
    if t.startswith("Ptr{"):
        return ("ptr", 8)
    return JL[t]


def c_class(t):
    t = t.strip()
    if "*" in t:
        return ("ptr", 8)
    base = t.replace("const", "").replace("unsigned", "").strip().split()[0] if t.replace("const", "").replace(
        "unsigned", "").strip() else "int"
    return {"int": ("int", 4), "int32_t": ("int", 4), "uint32_t": ("int", 4), "int64_t": ("int", 8),
            "uint64_t": ("int", 8), "float": ("float", 4), "double": ("float", 8), "void": ("void", 0),
            "char": ("int", 1), "uint8_t": ("int", 1)}[base]


def header_protos():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))      # preprocessor lines
    src = src.replace('extern "C" {', "").replace("typedef struct rayuela_index rayuela_index;", "")
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w \*]*?)\s*\b(\w+)\s*\(([^;{}]*?)\)\s*;", src):
        if name in ("defined",):
            continue
        args = [a.strip() for a in args.replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        # drop the parameter name: everything up to the last identifier
        types = [re.sub(r"\b\w+$", "", a).strip() or a for a in args]
        protos[name] = (ret.strip(), types)
    return protos


def shim_ccalls():
    src = open(SHIM).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*librayuela_b200\),\s*(\w+),\s*\(([^)]*)\)", src):
        name, ret, args = m.group(1), m.group(2), m.group(3)
        depth, cur, parts = 0, "", []
        for ch in args:                                  # split on top-level commas (Ptr{...} has none, but be safe)
            if ch == "{":
                depth += 1
            if ch == "}":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        out.append((name, ret, [p.strip() for p in parts if p.strip()]))
    return out


def test_every_ccall_matches_the_header():
    protos = header_protos()
    calls = shim_ccalls()
    assert len(calls) >= 11
    for name, ret, args in calls:
        assert name in protos, "%s is not declared in rayuela_b200.h" % name
        cret, cargs = protos[name]
        assert len(args) == len(cargs), "%s: shim passes %d arguments, header declares %d" % (name, len(args), len(cargs))
        assert jl_class(ret)[0] == c_class(cret)[0] or (jl_class(ret) == ("ptr", 8) and "*" in cret), name
        for i, (a, c) in enumerate(zip(args, cargs)):
            assert jl_class(a) == c_class(c), "%s argument %d: Julia %s vs C '%s'" % (name, i + 1, a, c)


def test_shim_defines_the_reference_api_names():
    src = open(SHIM).read()
    for fn in ("encoding_icm", "encode_icm_cuda", "veccost", "qerror", "quantize_pq", "quantize_opq", "linscan_pq",
               "linscan_opq", "linscan_lsq", "linscan_cq", "quantize_norms", "quantize_chainq", "fast_bin_matmul",
               "update_codebooks_fast_bin"):
        assert re.search(r"^(function\s+)?%s\(" % fn, src, flags=re.M), fn
