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


# ---- implicit none: every identifier of every procedure is a dummy, a local, a module-level name, a keyword or an intrinsic ----
_KEYWORDS = set("""if then else elseif end endif do while call return stop error exit cycle select case default where elsewhere
 subroutine function module contains use only implicit none interface procedure type integer real double precision character
 logical complex intent in out inout optional pointer contiguous dimension target allocatable parameter value import bind
 c name result pure elemental recursive write print read open close unit file status iostat fmt len kind intrinsic
 and or not eq ne lt le gt ge true false eqv neqv nullify allocate deallocate stat associate save public private""".split())
_INTRINSICS = set("""trim adjustl adjustr int real dble merge size present associated len_trim achar char ichar iachar min max abs
 epsilon null c_f_pointer c_loc c_associated c_funloc ieee_value ieee_quiet_nan shape reshape transfer huge tiny sqrt exp
 c_int c_double c_char c_ptr c_null_char c_null_ptr c_int64_t c_long c_size_t c_float c_bool iso_c_binding ieee_arithmetic
 lbound ubound index scan verify repeat any all count sum product maxval minval mod nint floor ceiling sign""".split())
_DECL = re.compile(r"^(integer|real|double precision|character|logical|type\s*\(|class\s*\()", re.I)
_PROC = re.compile(r"(?:(?:pure|elemental|recursive|integer\s*\([^)]*\)|integer|type\s*\([^)]*\)|real\s*\([^)]*\)|double precision|logical)\s+)*"
                   r"(subroutine|function)\s+(\w+)\s*(\(([^)]*)\))?")


def _idents(s):
    s = re.sub(r"z'[0-9a-fA-F]+'", " ", s)                                    # BOZ literals
    s = re.sub(r"'[^']*'|\"[^\"]*\"", " ", s)                                  # strings
    s = re.sub(r"(?<![\w])\d+\.?\d*([ed][+-]?\d+)?(_\w+)?", " ", s, flags=re.I)  # numbers with a kind suffix
    s = re.sub(r"%\s*\w+", " ", s)                                             # components
    s = re.sub(r"\b\w+\s*=(?!=)", lambda m: m.group(0) if not re.match(r".*[(,]\s*$", s[:m.start()]) else " ", s)   # keyword arguments
    s = re.sub(r"\.\w+\.", " ", s)                                             # .and. .eq. ...
    return [t.lower() for t in re.findall(r"[A-Za-z_]\w*", s)]


def _decl_names(s):
    return [re.sub(r"(\(.*|=>.*|=.*)", "", part).strip().lower() for part in re.split(r",(?![^()]*\))", s.split("::", 1)[1])]


def _undeclared(lines):
    """list of (procedure, identifier, statement) for identifiers that nothing declares"""
    modnames, procs, cur, in_iface = set(), [], None, 0
    for s in lines:
        low = s.lower()
        if re.match(r"interface\b", low):
            in_iface += 1
            m = re.match(r"interface\s+(\w+)", low)
            if m:
                modnames.add(m.group(1))
            continue
        if re.match(r"end\s*interface", low):
            in_iface -= 1
            continue
        m = _PROC.match(low)
        if m:
            modnames.add(m.group(2))
            if not in_iface:
                cur = dict(name=m.group(2), args=[a.strip() for a in (m.group(4) or "").split(",") if a.strip()], body=[], res=None)
                r = re.search(r"result\s*\(\s*(\w+)\s*\)", low)
                if r:
                    cur["res"] = r.group(1)
                procs.append(cur)
            continue
        if re.match(r"end\s*(subroutine|function)", low):
            if not in_iface:
                cur = None
            continue
        if in_iface:
            continue
        if cur is None:
            if _DECL.match(low) and "::" in low:
                modnames.update(_decl_names(s))
            m = re.match(r"type\s*(,[^:]*)?(::)?\s*(\w+)$", low)
            if m:
                modnames.add(m.group(3))
        else:
            cur["body"].append(s)
    bad = []
    for p in procs:
        local = set(p["args"]) | {p["name"]}
        if p["res"]:
            local.add(p["res"])
        for s in p["body"]:
            if _DECL.match(s.lower()) and "::" in s:
                local.update(_decl_names(s))
        for s in p["body"]:
            if _DECL.match(s.lower()) and "::" in s:
                toks = _idents(s.split("::", 1)[0])
                for part in re.split(r",(?![^()]*\))", s.split("::", 1)[1]):
                    toks += _idents(re.sub(r"^[^=(]*", "", part))
            else:
                toks = _idents(s)
            bad += [(p["name"], t, s) for t in toks if not (t in _KEYWORDS or t in _INTRINSICS or t in local or t in modnames)]
        bad += [(p["name"], a, "dummy argument without a declaration") for a in p["args"]
                if not any(_DECL.match(s.lower()) and "::" in s and a in _decl_names(s) for s in p["body"])]
    return bad, len(procs)


def test_every_identifier_is_declared():
    """the module is `implicit none`: an undeclared name is a compile error"""
    bad, nproc = _undeclared(_logical_lines())
    assert nproc >= 20 and bad == []


def test_the_identifier_check_sees_a_seeded_mistake():
    lines = [s.replace("integer :: q, k", "integer :: q") for s in _logical_lines()]      # par_to_c loses the declaration of k
    bad, _ = _undeclared(lines)
    assert {(p, t) for p, t, _ in bad} >= {("par_to_c", "k"), ("par_from_c", "k")}
    lines = [s.replace("call par_to_c(par, pc)", "call par_to_c(par, pcc)") for s in _logical_lines()]
    assert any(t == "pcc" for _, t, _ in _undeclared(lines)[0])


def _call_sites(stmt, name):
    """argument counts of every reference `name(...)` in a statement (balanced parentheses, top-level commas)"""
    s = re.sub(r"'[^']*'|\"[^\"]*\"", "''", stmt)
    out = []
    for m in re.finditer(r"(?<![\w%%])%s\s*\(" % re.escape(name), s, re.I):
        depth, i, commas, empty = 1, m.end(), 0, True
        while i < len(s) and depth:
            ch = s[i]
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            elif ch == "," and depth == 1:
                commas += 1
            if depth and not ch.isspace():
                empty = False
            i += 1
        out.append(0 if empty else commas + 1)
    return out


def test_references_pass_as_many_arguments_as_the_procedures_take():
    lines = _logical_lines()
    sig, optional, in_proc = {}, set(), None
    for s in lines:
        low = s.lower()
        m = _PROC.match(low)
        if m:
            in_proc = m.group(2)
            sig[in_proc] = len([a for a in (m.group(4) or "").split(",") if a.strip()])
        elif in_proc and "optional" in low.split("::")[0]:
            optional.add(in_proc)
    checked = 0
    for s in lines:
        low = s.lower()
        if _PROC.match(low) or re.match(r"(end\b|module procedure)", low):
            continue
        for name, n in sig.items():
            if name in optional or name not in low:
                continue
            for got in _call_sites(s, name):
                assert got == n, "%s takes %d arguments, %d given in: %s" % (name, n, got, s[:120])
                checked += 1
    assert checked >= 40


def test_elemental_procedures_only_reference_pure_interfaces_and_generics_name_real_procedures():
    lines = _logical_lines()
    pure, defined, iface_funcs = set(), set(), set()
    in_iface = 0
    for s in lines:
        low = s.lower()
        if re.match(r"interface\b", low):
            in_iface += 1
        elif re.match(r"end\s*interface", low):
            in_iface -= 1
        m = _PROC.match(low)
        if m:
            (iface_funcs if in_iface else defined).add(m.group(2))
            if re.match(r"(pure|elemental)\b", low):
                pure.add(m.group(2))
    for s in lines:
        m = re.match(r"module procedure\s+(.*)$", s.lower())
        if m:
            for name in m.group(1).split(","):
                assert name.strip() in defined, name
    cur = None
    for s in lines:
        low = s.lower()
        m = _PROC.match(low)
        if m:
            cur = m.group(2) if re.match(r"elemental\b", low) else None
            continue
        if re.match(r"end\s*(subroutine|function)", low):
            cur = None
        if cur:
            for f in iface_funcs | defined:
                if f != cur and _call_sites(s, f):
                    assert f in pure, "%s (elemental) references %s, which is not pure" % (cur, f)
            assert not re.match(r"(call check|error stop|stop|write|print)\b", low), "%s: %s" % (cur, s)


REF = "/root/reference/surface_physics.f90"


def _ref_logical_lines():
    out, cur = [], ""
    for raw in open(REF):
        code = raw.split("!")[0].strip()
        if not code:
            continue
        if code.startswith("&"):
            code = code[1:].lstrip()
        if code.endswith("&"):
            cur += code[:-1] + " "
            continue
        out.append(cur + code)
        cur = ""
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is not on this machine")
def test_types_and_public_procedures_match_the_reference_module():
    """drop-in: same type names, same component names (order included), same procedure names and dummy-argument lists as
    surface_physics.f90 (the reference is read where it lies; nothing is copied)"""
    ref, shim = _ref_logical_lines(), _logical_lines()
    for t in ("surface_param_class", "boundary_opt_class", "surface_physics_class"):
        assert _type_members(shim, t) == _type_members(ref, t), t
    # surface_state_class: the reference's components, then the shim's handle
    want = _type_members(ref, "surface_state_class")
    got = _type_members(shim, "surface_state_class")
    assert got[:len(want)] == want and got[len(want):] == ["handle"]
    low_ref, low_shim = "\n".join(ref).lower(), "\n".join(shim).lower()
    for m in re.finditer(r"^\s*(?:elemental\s+)?subroutine\s+(\w+)\s*\(([^)]*)\)", low_ref, re.M):
        name, args = m.group(1), [a.strip() for a in m.group(2).split(",")]
        if name in ("field_average", "diurnal_cycle", "albedo_isba", "albedo_slater", "albedo_denby"):
            continue                                                    # private helpers of the reference (not in its public list, :121-130)
        mm = re.search(r"subroutine\s+%s(?:_v)?\s*\(([^)]*)\)" % name, low_shim)
        assert mm, "%s is missing from the shim" % name
        assert [a.strip() for a in mm.group(1).split(",")] == args, name


def _public_names(lines):
    names = []
    for s in lines:
        m = re.match(r"public\s*::\s*(.*)$", s.lower())
        if m:
            names += [n.strip() for n in m.group(1).split(",")]
    return sorted(names)


def test_public_list_is_the_reference_modules():
    """everything else (helpers, bind(C) twins, field ids, pi/t0/...) is private, so no helper name can clash in a user program"""
    lines = _logical_lines()
    assert any(s.lower() == "private" for s in lines)
    want = sorted("""mass_balance energy_balance surface_energy_and_mass_balance surface_physics_class surface_state_class
                     surface_param_class boundary_opt_class surface_physics_par_load surface_alloc surface_dealloc
                     surface_boundary_define surface_physics_average print_param print_boundary_opt sigm cap eps cls clv
                     longwave_upward latent_heat_flux sensible_heat_flux""".split())
    assert _public_names(lines) == want
    if os.path.exists(REF):
        assert _public_names(_ref_logical_lines()) == want
