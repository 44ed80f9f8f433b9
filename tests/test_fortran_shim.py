"""Consistency checks of the ISO_C_BINDING shim (fortran/surface_physics_b200.f90).  No Fortran compiler exists in this image,
so the shim cannot be compiled here; these tests check what can be checked without one: every bind(C) name is a symbol the
library exports and the header declares, the bind(C) struct twins have the member order of the C structs, the field ids are the
header's, blocks are balanced, lines are within the free-form limit, and the reference module's public names are all there
(reference: surface_physics.f90:28-130).  CPU only."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "fortran", "surface_physics_b200.f90")
HEADER = os.path.join(ROOT, "include", "semic_b200.h")
LIB = os.path.join(ROOT, "semic_b200", "lib", "libsemic_b200.so")


def _logical_lines():
    """free-form source -> statements: comments stripped, continuation lines joined, lower case (strings kept)"""
    out, cur = [], ""
    for raw in open(SHIM):
        line = raw.rstrip("\n")
        assert len(line) <= 132, "line longer than the free-form limit: %r" % line[:60]
        code, in_str, q = "", False, ""
        for ch in line:                                     # strip the comment, not a '!' inside a string
            if in_str:
                code += ch
                if ch == q:
                    in_str = False
            elif ch in "'\"":
                in_str, q = True, ch
                code += ch
            elif ch == "!":
                break
            else:
                code += ch
        code = code.strip()
        if not code:
            continue
        if code.startswith("&"):
            code = code[1:].lstrip()
        if code.endswith("&"):
            cur += code[:-1] + " "
            continue
        out.append((cur + code).strip())
        cur = ""
    assert cur == "", "dangling continuation at the end of the file"
    return out


def _struct_members(name):
    """member names of `struct name { ... }` in the header, in order"""
    text = open(HEADER).read()
    m = re.search(r"(?:typedef\s+)?struct\s+%s\s*\{(.*?)\}" % name, text, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(unsigned\s+)?(char|int|double|float|long)\s+", "", decl)
        for part in decl.split(","):
            names.append(re.sub(r"\[.*", "", part).strip())
    return names


def _type_members(lines, tname):
    """component names of `type[, bind(C)] :: tname` in the shim, in order"""
    names, inside = [], False
    for s in lines:
        low = s.lower()
        if re.match(r"type\s*(,\s*bind\s*\(\s*c\s*\))?\s*(::)?\s*%s$" % tname, low):
            inside = True
            continue
        if inside and low.startswith("end type"):
            return names
        if inside:
            rhs = s.split("::", 1)[1]
            for part in re.split(r",(?![^()]*\))", rhs):
                names.append(re.sub(r"(\(.*|=>.*|=.*)", "", part).strip().lower())
    raise AssertionError("type %s not found" % tname)


def test_bind_c_names_are_exported_and_declared():
    text = open(SHIM).read()
    names = sorted(set(re.findall(r'bind\s*\(\s*C\s*,\s*name\s*=\s*"(\w+)"\s*\)', text, re.I)))
    assert len(names) >= 15
    header = open(HEADER).read()
    lib = C.CDLL(LIB)
    for n in names:
        assert re.search(r"\b%s\s*\(" % n, header), "%s is not declared in include/semic_b200.h" % n
        assert hasattr(lib, n), "%s is not exported by libsemic_b200.so" % n


def test_struct_twins_follow_the_c_structs():
    lines = _logical_lines()
    assert _type_members(lines, "surface_param_c") == _struct_members("semic_par")
    assert _type_members(lines, "boundary_opt_c") == _struct_members("semic_bnd")
    # the user-facing types carry the same components in the same order (what the reference declares, :28-54 and :92-105)
    assert _type_members(lines, "surface_param_class") == _struct_members("semic_par")
    assert _type_members(lines, "boundary_opt_class") == _struct_members("semic_bnd")


def test_field_ids_are_the_headers():
    header = open(HEADER).read()
    ids, k = {}, 0                                           # enum: explicit values where given, else declaration order
    m = re.search(r"enum[^{]*\{([^}]*SEMIC_F_T2M[^}]*)\}", header, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    for item in body.split(","):
        mm = re.match(r"SEMIC_(\w+)(?:\s*=\s*(\d+))?$", item.strip())
        if not mm:
            continue
        if mm.group(2) is not None:
            k = int(mm.group(2))
        ids[mm.group(1)] = k
        k += 1
    stmt = [s for s in _logical_lines() if "F_T2M" in s and "parameter" in s.lower()][0]
    shim = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"(F_\w+)\s*=\s*(\d+)", stmt))
    assert len(shim) == 31
    for k, v in shim.items():
        assert ids[k] == v, k
    # components of surface_state_class: the 31 fields in id order, then mask and the handle
    comps = _type_members(_logical_lines(), "surface_state_class")
    order = [k[2:].lower() for k, _ in sorted(shim.items(), key=lambda kv: kv[1])]
    assert comps[:31] == order and comps[31:] == ["mask", "handle"]


OPENERS = [("module", r"module\s+(?!procedure)\w+$"), ("type", r"type\s*(,[^:]*)?(::)?\s*\w+$"), ("interface", r"interface(\s+\w+)?$"),
           ("subroutine", r"((pure|elemental|recursive)\s+)*subroutine\s+\w+"), ("function", r"((pure|elemental|recursive|integer\s*\([^)]*\)|integer|type\s*\([^)]*\)|real\s*\([^)]*\)|double precision|logical)\s+)*function\s+\w+"),
           ("if", r"(\w+\s*:\s*)?if\s*\(.*\)\s*then$"), ("do", r"(\w+\s*:\s*)?do(\s+.*)?$"), ("select", r"select\s+case\s*\(")]


def test_blocks_are_balanced():
    stack = []
    for s in _logical_lines():
        low = re.sub(r"'[^']*'|\"[^\"]*\"", "''", s.lower())
        m = re.match(r"end\s*(module|type|interface|subroutine|function|if|do|select)\b", low)
        if m:
            assert stack, "unmatched %r" % s
            kind = stack.pop()
            assert kind == m.group(1), "%r closes a %s block" % (s, kind)
            continue
        if low == "end":
            assert stack
            stack.pop()
            continue
        for kind, pat in OPENERS:
            if re.match(pat, low):
                stack.append(kind)
                break
    assert stack == [], "unclosed blocks: %s" % stack


def test_public_names_of_the_reference_module_exist():
    low = "\n".join(_logical_lines()).lower()
    for sub in ("surface_alloc", "surface_dealloc", "surface_physics_par_load", "surface_boundary_define",
                "surface_energy_and_mass_balance", "energy_balance", "mass_balance", "surface_physics_average",
                "print_param", "print_boundary_opt"):
        assert re.search(r"subroutine\s+%s\s*\(" % sub, low), sub
    for gen in ("sensible_heat_flux", "latent_heat_flux", "longwave_upward"):
        assert re.search(r"interface\s+%s\b" % gen, low), gen
    for t in ("surface_param_class", "surface_state_class", "boundary_opt_class", "surface_physics_class"):
        assert re.search(r"type\s+%s\b" % t, low), t
    for const in ("pi", "t0", "sigm", "eps", "cls", "clm", "clv", "cap", "rhow", "hsmax", "epsil"):
        assert re.search(r"parameter\s*::\s*%s\s*=" % const, low), const
    assert re.search(r"^module\s+surface_physics$", low, re.M) and re.search(r"^end module surface_physics$", low, re.M)


@pytest.mark.parametrize("sub,args", [("surface_alloc", ["now", "npts"]), ("surface_dealloc", ["now"]),
                                      ("surface_physics_par_load", ["par", "filename"]),
                                      ("surface_boundary_define", ["bnd", "boundary"]),
                                      ("surface_energy_and_mass_balance", ["now", "par", "bnd", "day", "year"]),
                                      ("surface_physics_average", ["ave", "now", "step", "nt"])])
def test_dummy_argument_lists_match_the_reference(sub, args):
    """surface_physics.f90:138, :565, :633, :681, :729, :819"""
    low = "\n".join(_logical_lines()).lower()
    m = re.search(r"subroutine\s+%s\s*\(([^)]*)\)" % sub, low)
    assert m and [a.strip() for a in m.group(1).split(",")] == args


def test_bind_c_interfaces_have_the_argument_counts_of_the_prototypes():
    """each `function f(a, b, ...) bind(C, name="semic_x")` of the shim against `semic_x(...)` of the header"""
    header = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    proto = {}
    for m in re.finditer(r"\b(semic_\w+)\s*\(([^;{]*?)\)\s*;", header, re.S):
        args = m.group(2).strip()
        proto[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    checked = 0
    for s in _logical_lines():
        m = re.search(r"(?:function|subroutine)\s+\w+\s*\(([^)]*)\)\s*(?:result\s*\(\w+\)\s*)?bind\s*\(\s*c\s*,\s*name\s*=\s*\"(\w+)\"\s*\)", s, re.I)
        if not m:
            continue
        nargs = 0 if not m.group(1).strip() else m.group(1).count(",") + 1
        assert m.group(2) in proto, m.group(2)
        assert nargs == proto[m.group(2)], "%s: %d dummy arguments, the prototype has %d" % (m.group(2), nargs, proto[m.group(2)])
        checked += 1
    assert checked >= 15
