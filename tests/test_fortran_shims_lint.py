"""The Fortran drop-in modules (fortran/*.f90) cannot be compiled in this image (no Fortran compiler).  This is a structural lint
that catches the typos a blind edit produces: every block opener has its closer, every C function bound in ipic_b200_c.f90 exists in
include/ipic1d_b200.h with the same number of arguments, and every ipic_* call in the shims has an interface."""
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(f for f in os.listdir(os.path.join(ROOT, "fortran")) if f.endswith(".f90"))


def code_lines(path):
    out = []
    for ln in open(path):
        ln = ln.split("!")[0].rstrip()                      # comments never contain code here (no '!' inside strings that matter)
        if ln.strip() and not ln.lstrip().startswith("#"):
            out.append(ln.strip().lower())
    # join continuation lines
    joined, cur = [], ""
    for ln in out:
        if cur:
            ln = ln.lstrip("&").strip()
        cur += ln
        if cur.endswith("&"):
            cur = cur[:-1]
            continue
        joined.append(cur)
        cur = ""
    return joined


@pytest.mark.parametrize("name", FILES)
def test_blocks_are_balanced(name):
    lines = code_lines(os.path.join(ROOT, "fortran", name))
    def count(pat):
        return sum(1 for ln in lines if re.match(pat, ln))
    assert count(r"(pure\s+)?subroutine\s+\w+") == count(r"end\s+subroutine")
    assert count(r"(\w+(\([^)]*\))?\s+)?function\s+\w+\s*\(") == count(r"end\s+function")
    assert count(r"module\s+(?!procedure)\w+") == count(r"end\s+module")
    assert count(r"interface\b") == count(r"end\s+interface")
    assert count(r"type\s*(,\s*bind\(c\))?\s*::\s*\w+$") == count(r"end\s+type")
    assert count(r"(\w+\s*:\s*)?do\b") == count(r"end\s*do")
    # block IFs end with THEN; one-line IFs do not
    assert count(r"(else\s*)?if\s*\(.*\)\s*then$") - count(r"else\s*if\s*\(.*\)\s*then$") == count(r"end\s*if")
    assert count(r"select\s+case") == count(r"end\s+select")


def header_functions():
    txt = open(os.path.join(ROOT, "include", "ipic1d_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    fns = {}
    for m in re.finditer(r"\b(?:int|const char \*)\s*(ipic_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        fns[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return fns


def test_bound_functions_exist_in_the_header_with_the_same_arity():
    hdr = header_functions()
    lines = code_lines(os.path.join(ROOT, "fortran", "ipic_b200_c.f90"))
    bound = {}
    for ln in lines:
        m = re.match(r"function\s+(\w+)\s*\(([^)]*)\).*bind\(c,\s*name=\"(\w+)\"\)", ln)
        if m:
            args = [a for a in m.group(2).split(",") if a.strip()]
            bound[m.group(3).lower()] = (m.group(1), len(args))
    assert len(bound) > 30
    lower = {k.lower(): v for k, v in hdr.items()}
    for cname, (fname, nargs) in bound.items():
        assert cname in lower, cname + " is not declared in include/ipic1d_b200.h"
        assert lower[cname] == nargs, "%s: %d arguments in the Fortran interface, %d in the header" % (cname, nargs, lower[cname])
        assert fname == cname
    # every ipic_* function the shims call has an interface (helpers defined in the shims themselves excepted)
    defined = set()
    for name in FILES:
        for ln in code_lines(os.path.join(ROOT, "fortran", name)):
            m = re.match(r"(?:pure\s+)?(?:subroutine|function)\s+(ipic_\w+)", ln)
            if m:
                defined.add(m.group(1))
    for name in FILES:
        if name == "ipic_b200_c.f90":
            continue
        for ln in code_lines(os.path.join(ROOT, "fortran", name)):
            for m in re.finditer(r"\b(ipic_[a-z0-9_]+)\s*\(", ln):
                f = m.group(1)
                assert f in bound or f in defined or f in ("ipic_check",), "%s: %s has no interface" % (name, f)


def test_every_abi_entry_is_mapped_to_the_reference_in_INTEGRATION_md():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(f for f in header_functions() if f not in doc)
    assert not missing, "INTEGRATION.md (section 6) does not map: %s" % missing


def test_bind_c_types_match_the_c_structs():
    """field order, kinds and array lengths of the bind(c) derived types against the ctypes mirrors of the header's structs (which the GPU
    tests exercise against the library)"""
    import ctypes as C
    from implicitpic1d_b200 import _lib
    lines = code_lines(os.path.join(ROOT, "fortran", "ipic_b200_c.f90"))
    params = {"ipic_max_collisions": _lib.MAX_COLLISIONS}
    kinds = {"c_int32_t": C.c_int32, "c_int64_t": C.c_int64, "c_double": C.c_double, "c_int": C.c_int}
    types, cur = {}, None
    for ln in lines:
        m = re.match(r"type\s*,\s*bind\(c\)\s*::\s*(\w+)", ln)
        if m:
            cur = m.group(1); types[cur] = []
            continue
        if cur and re.match(r"end\s+type", ln):
            cur = None
            continue
        if cur:
            m = re.match(r"(?:integer|real)\((\w+)\)\s*::\s*(.*)", ln)
            assert m, ln
            for item in re.findall(r"(\w+)(?:\(([^)]*)\))?", m.group(2)):
                n = item[1].strip()
                count = 1 if not n else (int(n) if n.isdigit() else params[n])
                types[cur].append((item[0], kinds[m.group(1)], count))
    for fname, struct in (("ipic_move_stats", _lib.MoveStats), ("ipic_solver_stats", _lib.SolverStats), ("ipic_collision_stats", _lib.CollisionStats)):
        got = types[fname]
        want = []
        for name, ct in struct._fields_:
            if hasattr(ct, "_length_"):
                want.append((name.lower(), ct._type_, ct._length_))
            else:
                want.append((name.lower(), ct, 1))
        assert [(a, b, c) for a, b, c in got] == want, fname
