"""Static check of the Julia host shim (julia/SpectralElementsB200.jl) against the C ABI (include/semb.h).

Julia is not installed in the build container, so the shim cannot be executed here.  What can be verified
without it: every `ccall((:sym, libsemb), Ret, (ArgTypes...), args...)` names an exported symbol, declares the
prototype's return type, the prototype's number of parameters with ABI-compatible Julia types, and passes as
many values as it declares types.  A wrong count or type in a ccall is silent memory corruption at run time,
so this is the host-logic test for the reference-language side of the boundary."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "semb.h")
SHIM = os.path.join(ROOT, "julia", "SpectralElementsB200.jl")

# C parameter type (normalised) -> Julia ccall types that are ABI-compatible with it
HANDLE = {"Ptr{Cvoid}"}
HANDLE_OUT = {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"}
COMPAT = {
    "int": {"Cint"},
    "double": {"Cdouble", "Float64"},
    "long long": {"Clonglong", "Int64"},
    "size_t": {"Csize_t"},
    "uint64_t": {"UInt64", "Culonglong"},
    "double*": {"Ptr{Float64}", "Ref{Float64}", "Ptr{Cdouble}", "Ref{Cdouble}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "long long*": {"Ptr{Clonglong}", "Ref{Clonglong}", "Ptr{Int64}", "Ref{Int64}"},
    "char*": {"Cstring", "Ptr{UInt8}", "Ptr{Cchar}"},
    "void*": {"Ptr{Cvoid}"},
    "void**": HANDLE_OUT,
}
RET = {"int": {"Cint"}, "const char*": {"Cstring", "Ptr{UInt8}", "Ptr{Cchar}"}, "char*": {"Cstring", "Ptr{UInt8}"}}


def _strip_c_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def _norm_ctype(param):
    """'const semb_field* nu_arr' -> 'semb_field*'; 'const char bc[4]' -> 'char*'; 'int n' -> 'int'."""
    p = param.strip()
    arr = "[" in p
    p = re.sub(r"\[[^\]]*\]", "", p)
    p = re.sub(r"\bconst\b", " ", p)
    stars = p.count("*")
    p = p.replace("*", " ")
    toks = p.split()
    base = {"int", "double", "char", "void", "size_t", "uint64_t", "long", "unsigned"}
    if len(toks) > 1 and not (toks[-1] in base or toks[-1].startswith("semb_")):
        toks = toks[:-1]  # drop the parameter name
    elif len(toks) > 1 and toks[-1].startswith("semb_") and toks[-2].startswith("semb_"):
        toks = toks[:-1]
    t = " ".join(toks)
    return t + "*" * (stars + (1 if arr else 0))


def header_prototypes():
    txt = _strip_c_comments(open(HEADER).read())
    txt = re.sub(r"\s+", " ", txt)
    protos = {}
    for m in re.finditer(r"\b((?:const )?(?:int|char) ?\*?) *(semb_\w+) *\(([^()]*)\) *;", txt):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [_norm_ctype(p) for p in params.split(",")]
        protos[name] = (ret.replace(" *", "*"), plist)
    return protos


def _split_top(s):
    """split on commas that are not inside (), [], {}"""
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        out.append(tail)
    return out


def julia_ccalls():
    src = open(SHIM).read()
    src = re.sub(r"#[^\n]*", "", src)  # the shim has no '#' inside strings
    calls = []
    for m in re.finditer(r"ccall\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        body = src[m.end():i - 1]
        parts = _split_top(body)
        sym = re.match(r"\(\s*:(\w+)\s*,\s*libsemb\s*\)", parts[0])
        assert sym, "unrecognised ccall target: %r" % parts[0]
        argt = parts[2].strip()
        assert argt.startswith("(") and argt.endswith(")"), parts[2]
        types = _split_top(argt[1:-1])
        calls.append((sym.group(1), parts[1].strip(), types, parts[3:], src.count("\n", 0, m.start()) + 1))
    return calls


def _compatible(ctype, jtype):
    if ctype.startswith("semb_"):
        if ctype.endswith("**"):
            return jtype in HANDLE_OUT
        if ctype == "semb_pcg_opts*":
            return jtype.startswith("Ref{") or jtype.startswith("Ptr{")
        return jtype in HANDLE
    return jtype in COMPAT.get(ctype, set())


def test_header_parses_every_export():
    protos = header_prototypes()
    n_decl = len(re.findall(r"\bsemb_\w+ *\(", _strip_c_comments(open(HEADER).read())))
    assert len(protos) == n_decl and n_decl > 80, (len(protos), n_decl)
    assert protos["semb_init"] == ("int", ["int", "semb_ctx**"])
    assert protos["semb_last_error"] == ("const char*", [])
    assert protos["semb_mask_bc"][1] == ["semb_mesh*", "semb_field*", "char*", "semb_field*"]


def test_every_ccall_matches_its_prototype():
    protos, calls = header_prototypes(), julia_ccalls()
    assert len(calls) >= 35
    bad = []
    for sym, ret, types, args, line in calls:
        where = "%s (shim line ~%d)" % (sym, line)
        if sym not in protos:
            bad.append(where + ": not declared in include/semb.h")
            continue
        cret, cparams = protos[sym]
        if ret not in RET.get(cret, set()):
            bad.append(where + ": return type %s for C %s" % (ret, cret))
        if len(types) != len(cparams):
            bad.append(where + ": %d argument types, prototype has %d" % (len(types), len(cparams)))
            continue
        if len(args) != len(types):
            bad.append(where + ": %d values for %d declared types" % (len(args), len(types)))
        for k, (ct, jt) in enumerate(zip(cparams, types)):
            if not _compatible(ct, jt):
                bad.append(where + ": parameter %d is C %s, ccall says %s" % (k + 1, ct, jt))
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("path", [SHIM, os.path.join(ROOT, "tools", "ref_dump.jl"), os.path.join(ROOT, "bench", "ref_cpu.jl"),
                                  os.path.join(ROOT, "julia", "test", "runtests.jl")])
def test_julia_blocks_balance(path):
    """Every block opener has its `end` (a cheap guard for files that cannot be parsed by Julia here)."""
    src = re.sub(r"#[^\n]*", "", open(path).read())
    src = re.sub(r'"""(?:.|\n)*?"""', '""', src)
    src = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', src)
    src = re.sub(r"\[[^\[\]\n]*\bend\b[^\[\]\n]*\]", "[]", src)  # a[end] indexing
    src = re.sub(r"\[[^\[\]\n]*\bfor\b[^\[\]\n]*\]", "[]", src)  # comprehensions
    opens = len(re.findall(r"(?<![\w.:])(?:function|if|for|while|let|try|begin|do|struct|module|quote|macro)\b(?!\s*=)", src))
    opens -= len(re.findall(r"\bmutable\s+struct\b", src)) * 0
    ends = len(re.findall(r"(?<![\w.:])end\b", src))
    assert opens == ends, (opens, ends)


@pytest.mark.parametrize("name", ["ABu", "lapl", "hlmz", "mass", "gatherScatter", "mask", "pcg", "pcg!", "opLHS", "solve!",
                                  "grad", "advect", "laplace", "evolve!", "step!"])
def test_shim_defines_the_reference_methods(name):
    """The reference's signatures for the path are extended, not renamed."""
    src = open(SHIM).read()
    imported = re.search(r"^import SpectralElements:((?:[^\n]*,\n)*[^\n]*)", src, flags=re.M).group(1)
    assert name in [x.strip() for x in imported.split(",")], name
    assert re.search(r"^(?:function\s+)?%s\(" % re.escape(name), src, flags=re.M), name


def test_pcg_opts_struct_layout_matches():
    """`struct PcgOpts` in the shim is passed by reference as `semb_pcg_opts`: same fields, order and C types."""
    hdr = _strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct semb_pcg_opts \{(.*?)\} semb_pcg_opts;", hdr, flags=re.S).group(1)
    cfields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            name = re.search(r"(\w+)$", decl).group(1)
            cfields.append((name, _norm_ctype(decl)))
    src = re.sub(r"#[^\n]*", "", open(SHIM).read())
    jbody = re.search(r"struct PcgOpts\n(.*?)\nend", src, flags=re.S).group(1)
    jfields = [tuple(x.strip() for x in f.split("::")) for f in re.split(r"[;\n]", jbody) if f.strip()]
    assert [n for n, _ in cfields] == [n for n, _ in jfields]
    want = {"double": {"Cdouble"}, "int": {"Cint"}, "long long": {"Clonglong"}, "semb_field*": {"Ptr{Cvoid}"},
            "char*": {"Ptr{UInt8}", "Cstring"}}
    for (cn, ct), (jn, jt) in zip(cfields, jfields):
        assert jt in want[ct], (cn, ct, jt)
    # the positional constructor call in pcg() passes one value per field
    ctor = re.search(r"PcgOpts\(([^\n]*)\)\)\n", src).group(1)
    assert len(_split_top(ctor)) == len(cfields)


def _c_enum(name):
    hdr = _strip_c_comments(open(HEADER).read())
    body = re.search(r"enum %s \{(.*?)\}" % name, hdr, flags=re.S).group(1)
    vals, nxt = {}, 0
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = [x.strip() for x in item.split("=")]
            nxt = int(v)
        else:
            k = item
        vals[k] = nxt
        nxt += 1
    return vals


def test_shim_enum_constants_match_the_header():
    """Integer selectors hard-coded in the shim (mesh arrays for semb_mesh_set, driver fields) equal the C enums."""
    src = open(SHIM).read()
    ma = _c_enum("semb_mesh_array")
    pairs = re.search(r"for \(which, a\) in \(\((\d+), msh\.Jac\), \((\d+), msh\.Jaci\), \((\d+), msh\.rx\), \((\d+), msh\.ry\), "
                      r"\((\d+), msh\.sx\), \((\d+), msh\.sy\), \((\d+), msh\.Bi\)\)", src)
    assert [int(g) for g in pairs.groups()] == [ma["SEMB_JAC"], ma["SEMB_JACI"], ma["SEMB_RX"], ma["SEMB_RY"], ma["SEMB_SX"],
                                                ma["SEMB_SY"], ma["SEMB_BI"]]   # (Bi: the Stokes split needs it)
    df = _c_enum("semb_diffusion_field_id")
    names = re.search(r"const (DFN_\w+(?:, DFN_\w+)*) = ([\d, ]+)\n", src)
    got = dict(zip([n.strip() for n in names.group(1).split(",")], [int(v) for v in names.group(2).split(",")]))
    want = {"DFN_U": df["SEMB_DFN_U"], "DFN_UB": df["SEMB_DFN_UB"], "DFN_NU": df["SEMB_DFN_NU"], "DFN_F": df["SEMB_DFN_F"],
            "DFN_RHS": df["SEMB_DFN_RHS"], "DFN_VX": df["SEMB_DFN_VX"], "DFN_VY": df["SEMB_DFN_VY"], "DFN_UH0": df["SEMB_DFN_UH0"]}
    assert got == want
