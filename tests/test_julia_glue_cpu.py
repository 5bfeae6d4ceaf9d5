"""The Julia glue (shiftedproximaloperators.jl_b200/julia/ShiftedProxB200.jl) cannot be executed in this image (no
`julia`), so it is checked statically: it is GENERATED (tools/gen_julia_glue.py) and must be up to date, every `ccall`
in it must name a function include/shiftedprox.h declares, with the same number of arguments and compatible types,
and the eleven shifted types of the path must each have their struct, their `shifted` constructors, `prox!` and ψ(y)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "shiftedproximaloperators.jl_b200", "julia", "ShiftedProxB200.jl")
HDR = os.path.join(ROOT, "include", "shiftedprox.h")


def header_decls():
    """{name: (return type, [arg types])} from the preprocessed header."""
    src = subprocess.run(["gcc", "-E", "-P", HDR], capture_output=True, text=True, check=True).stdout
    src = re.sub(r"\s+", " ", src)
    decls = {}
    for m in re.finditer(r"(const char\*|int32_t|double|float) (spx_\w+)\(([^()]*(?:\([^()]*\)[^()]*)*)\);", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        if args in ("void", ""):
            decls[name] = (ret, [])
            continue
        types = []
        for a in args.split(","):
            a = a.strip()
            a = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a  # drop the parameter name
            types.append(re.sub(r"\s+", " ", a))
        decls[name] = (ret, types)
    return decls


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    text = open(JL).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(spx_\w+), libshiftedprox\), (\w+),\s*\(", text):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        calls.append((m.group(1), m.group(2), split_top(text[m.end():i - 1])))
    return calls


SCALAR = {"int32_t": {"Int32"}, "int64_t": {"Int64"}, "size_t": {"Csize_t"}, "double": {"Cdouble", "Float64"},
          "float": {"Cfloat", "Float32"}, "uint64_t": {"UInt64"}}


def compatible(c, j):
    c = c.replace("const ", "").strip()
    if c in SCALAR:
        return j in SCALAR[c]
    if not c.endswith("*"):
        return False
    if not (j.startswith("Ptr{") or j.startswith("Ref{")):
        return False
    inner = j[4:-1]
    pointee = c[:-1].strip()
    table = {"spx_ctx": {"Cvoid"}, "spx_ctx*": {"Ptr{Cvoid}"}, "void": {"Cvoid", "Float64", "Float32", "Int64", "UInt32"},
             "void*": {"Ptr{Cvoid}"}, "double": {"Cdouble", "Float64"}, "float": {"Cfloat", "Float32"},
             "int64_t": {"Int64"}, "int32_t": {"Int32"}, "uint32_t": {"UInt32"}, "uint64_t": {"UInt64"},
             "spx_bound": {"SpxBound"}, "spx_sel": {"SpxSel"}}
    return inner in table.get(pointee, set())


def test_generated_file_is_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_julia_glue.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_ccall_matches_the_header():
    decls = header_decls()
    assert len(decls) > 90  # the preprocessed header really was parsed
    calls = julia_ccalls()
    assert len(calls) > 80
    for name, ret, jtypes in calls:
        assert name in decls, f"{name} is not declared in include/shiftedprox.h"
        cret, ctypes_ = decls[name]
        assert (cret, ret) in (("int32_t", "Int32"), ("const char*", "Cstring")), (name, cret, ret)
        assert len(ctypes_) == len(jtypes), (name, ctypes_, jtypes)
        for k, (c, j) in enumerate(zip(ctypes_, jtypes)):
            assert compatible(c, j), (name, k, c, j)


def test_all_eleven_types_are_wired():
    text = open(JL).read()
    types = ["ShiftedNormL1", "ShiftedNormL0", "ShiftedRootNormLhalf", "ShiftedNormL1Box", "ShiftedNormL0Box",
             "ShiftedRootNormLhalfBox", "ShiftedNormL1B2", "ShiftedIndBallL0", "ShiftedIndBallL0BInf", "ShiftedGroupNormL2",
             "ShiftedGroupNormL2Binf"]
    for t in types:
        assert re.search(rf"mutable struct {t}\{{", text), t
        assert re.search(rf"shifted\(ψ::{t}\{{", text), f"second shift of {t}"
        for R in ("Float64", "Float32"):
            assert re.search(rf"function prox!\(y::DeviceVector\{{{R}\}}, ψ::{t}\{{[^}}]*{R}\}}", text), (t, R, "prox!")
            assert re.search(rf"function \(ψ::{t}\{{[^}}]*{R}\}}\)\(y::DeviceVector\{{{R}\}}\)", text), (t, R, "ψ(y)")
    for t in ("ShiftedNormL1", "ShiftedNormL0", "ShiftedNormL1Box", "ShiftedNormL0Box"):
        assert re.search(rf"function iprox!\(y::DeviceVector\{{Float64\}}, ψ::{t}\{{Float64\}}", text), (t, "iprox!")
    for verb in ("shift!", "set_bounds!", "set_radius!", "prox_zero", "iprox_zero", "Base.getproperty"):
        assert verb in text
    # balanced `function` / `end` pairs as a cheap syntax sanity check
    opens = len(re.findall(r"^\s*(?:mutable struct|struct|function|module|for|if)\b", text, re.M))
    opens += len(re.findall(r"\bbegin\b", text))
    ends = len(re.findall(r"^\s*end\b", text, re.M))
    assert opens == ends, (opens, ends)
