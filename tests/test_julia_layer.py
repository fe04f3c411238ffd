"""The Julia host layer (julia/OceanosticsB200.jl) against the C header, without Julia: the struct mirrors must list the
header's fields in order with matching types and array lengths, the diagnostic table must be the header's enum in order,
every constant must agree, every exported entry point must be bound by a `ccall`, and every `ccall` must name a declared
symbol -- so the file a maintainer drops next to Oceanostics cannot drift from include/oceanostics_b200.h."""
import os
import re

from helpers import ROOT, ocb

JL = open(os.path.join(ROOT, "julia", "OceanosticsB200.jl"), encoding="utf-8").read()
HDR_RAW = open(os.path.join(ROOT, "include", "oceanostics_b200.h"), encoding="utf-8").read()
HDR = re.sub(r"/\*.*?\*/", "", HDR_RAW, flags=re.S)

MIRRORS = {"ocb_grid": "OcbGrid", "ocb_field": "OcbField", "ocb_closure": "OcbClosure", "ocb_request": "OcbRequest",
           "ocb_sweep_inputs": "OcbSweepInputs", "ocb_filter_pass": "OcbFilterPass", "ocb_ring": "OcbRing"}
DEFINES = {m.group(1): m.group(2) for m in re.finditer(r"#define\s+(OCB_[A-Z_0-9]+)\s+(\S+)", HDR)}


def _c_int(expr):
    expr = expr.strip()
    for name, val in DEFINES.items():
        expr = re.sub(r"\b%s\b" % name, val, expr)
    return int(eval(expr, {"__builtins__": {}}))


def c_struct(name):
    """[(field, ctype, array length or None)] of `typedef struct { ... } name;`"""
    body = re.search(r"typedef struct\s*\{([^{}]*)\}\s*%s\s*;" % name, HDR).group(1)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const\s+)?([A-Za-z_0-9]+)\s*(\*?)\s*(.*)$", decl)
        ctype = m.group(2) + ("*" if m.group(3) else "")
        for item in m.group(4).split(","):
            item = item.strip()
            a = re.match(r"(\*?)\s*([A-Za-z_0-9]+)\s*(?:\[(.+)\])?$", item)
            ptr, fname, length = a.group(1), a.group(2), a.group(3)
            out.append((fname, ctype + ("*" if ptr else ""), _c_int(length) if length else None))
    return out


def jl_struct(name):
    body = re.search(r"^struct %s\n(.*?)^end" % name, JL, flags=re.S | re.M).group(1)
    out = []
    for line in body.strip().splitlines():
        fname, jtype = [x.strip() for x in line.split("::")]
        m = re.match(r"NTuple\{(\d+),\s*([\w{}]+)\}$", jtype)
        out.append((fname, m.group(2), int(m.group(1))) if m else (fname, jtype, None))
    return out


C2JL = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "void*": "CuPtr{Cvoid}"}
C2JL.update(MIRRORS)


def test_struct_mirrors_match_the_header_field_by_field():
    for cname, jname in MIRRORS.items():
        c, j = c_struct(cname), jl_struct(jname)
        assert [f[0] for f in c] == [f[0] for f in j], (cname, [f[0] for f in c], [f[0] for f in j])
        for (fname, ctype, clen), (_, jtype, jlen) in zip(c, j):
            assert C2JL[ctype] == jtype, (cname, fname, ctype, jtype)
            assert clen == jlen, (cname, fname, clen, jlen)


def test_diagnostic_table_is_the_header_enum():
    body = re.search(r"enum\s*\{\s*(OCB_DIAG_KE\b.*?)OCB_DIAG_COUNT", HDR, flags=re.S).group(1)
    enum = [x.strip().replace("OCB_DIAG_", "").split("=")[0].strip() for x in body.split(",") if x.strip()]
    names = re.search(r"const DIAG_NAMES = \((.*?)\)\n", JL, flags=re.S).group(1)
    jl = [x.strip().lstrip(":") for x in names.replace("\n", " ").split(",") if x.strip()]
    assert jl == enum == ocb._lib._DIAG_NAMES
    # every row of the kernel-function table names a diagnostic of the enum, and every diagnostic a Julia caller can reach
    # through an Oceanostics constructor has a row
    rows = set(re.findall(r"=>\s*\(:([A-Z_0-9]+),", JL)) | set(re.findall(r"\] = \(:([A-Z_0-9]+),", JL))
    rows |= set(re.findall(r"\(:([A-Z_0-9]+), \d\)", JL))          # rows filled in a loop: for (row, key) in ((:U_TERM, 1), ...)
    assert rows <= set(enum)
    reached_otherwise = {"NU_SMAG"}        # compute_smagorinsky_viscosity! (no Oceanostics kernel function: Oceananigans' closure step)
    assert set(enum) - rows == reached_otherwise, sorted(set(enum) - rows)


def test_constants_agree_with_the_header():
    for name, val in DEFINES.items():
        m = re.search(r"const %s = (\d+)" % name, JL)
        if name in ("OCB_ABI_VERSION", "OCB_MAX_CLOSURE_FIELDS", "OCB_N_AUX", "OCB_MAX_FILTER_WIDTH", "OCB_MAX_RANKS", "OCB_RING_SYNC_WORDS"):
            assert m and int(m.group(1)) == _c_int(val), name
    # enumerators written as `A = 0, B = 1` or `A = 1 << 4` in the header
    for name, expr in re.findall(r"\b(OCB_[A-Z_0-9]+)\s*=\s*([0-9]+(?:\s*<<\s*[0-9]+)?)", HDR):
        if name.startswith("OCB_DIAG_") or name.startswith("OCB_ERR") or name == "OCB_OK":
            continue
        m = re.search(r"\b%s\b[^\n]*" % name, JL)
        assert m, f"{name} is missing from the Julia layer"
        # positional tuple constants: find this name's position in the `const A, B, C = x, y, z` line
        line = m.group(0) if "const" in m.group(0) else re.search(r"const [^\n]*\b%s\b[^\n]*" % name, JL).group(0)
        lhs, rhs = line.replace("const ", "").split(" = ", 1)
        names = [x.strip() for x in lhs.split(",")]
        vals = re.findall(r"Int32\(([^)]*)\)", rhs)
        got = eval(vals[names.index(name)], {"__builtins__": {}})
        assert got == eval(expr, {"__builtins__": {}}), (name, got, expr)


def test_every_entry_point_is_bound_and_every_ccall_is_declared():
    declared = sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", HDR)))
    bound = set(re.findall(r"ccall\(\(\s*:([a-z0-9_]+)\s*,\s*lib\)", JL)) | set(re.findall(r":(ocb_[a-z0-9_]+)", JL))
    assert bound <= set(declared), sorted(bound - set(declared))
    assert set(declared) <= bound, sorted(set(declared) - bound)
    assert sorted(ocb._lib.EXPORTED_SYMBOLS) == declared


def test_overrides_sit_at_the_reference_seams():
    """The methods the layer adds are the ones SURVEY.md section 8(b) names: compute! for Field{..,<:CustomKFO}, for reduced
    Fields over Integral / Average, for the staged filter aliases and for the sorted reference height."""
    for needle in ("const ComputedKFOField = Field{<:Any,<:Any,<:Any,<:CustomKFO}", "function compute!(comp::ComputedKFOField",
                   "function compute!(comp::ReducedKFOField", "SF._BoxFilter3D", "SF._GaussianFilter3D",
                   "function compute!(z✶::BPEq.SortedReferenceHeightField", "function fuse!(", "fill_halo_regions!(comp)",
                   "set_status!(comp.status, time)", "compute_at!(a, time)"):
        assert needle in JL, needle
