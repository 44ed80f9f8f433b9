#!/usr/bin/env python3
"""
f90_to_c.py -- a small, physics-free Fortran-90 -> C source-to-source translator.

TEST INFRASTRUCTURE ONLY (see oracle/semic_oracle.h for who may use oracle/).

Why it exists: the reference (mkrapp/semic) is Fortran and this image has no Fortran compiler, so
the reference cannot be compiled as it stands.  This script translates the reference's OWN source
files, where they lie under /root/reference, into C at build time (`make -C oracle ref`); the
output goes to oracle/_ref/ (git-ignored, never committed, never hand-edited) and is compiled by
the same gcc middle/back end gfortran uses.  The hand-written restatement (semic_oracle.c) is then
checked bit for bit against that mechanically derived build (tests/test_ref_pin.py).

The translator knows NOTHING about snow, albedo or energy balances: it is a lexer, a recursive-
descent parser and a type-directed code generator for the Fortran subset the reference uses:

  * modules with parameters, derived types, module variables, `contains`, programs, `use`
  * subroutines/functions, `elemental`, `result(...)`, `optional` + `present()`, intent
  * double precision / real / integer / logical / character(len=N|*) entities; allocatable,
    automatic, explicit-shape, assumed-shape and fixed-size arrays of rank 1-2; array sections
  * whole-array assignment (one C loop per statement, like gfortran's scalarizer), `where /
    elsewhere / end where` (mask evaluated first into a temporary, one loop per body statement),
    elemental calls on arrays (one loop per call), `sum()` over a rank-1 expression (sequential)
  * if / else if / else, one-line if, do loops, select case on strings, exit, return, stop
  * allocate / deallocate, namelist groups + open / read(nml=) / read(u,*) / write / close /
    inquire through the tiny runtime in f90_rt.h
  * Fortran expression grammar: `**` right-associative and above unary minus, left-to-right
    `* /` and `+ -`, parentheses preserved, mixed-mode arithmetic with C's identical promotions,
    default-real literals emitted as C float literals (so `0.8` is 0.800000011920929, as in
    Fortran), `x**y` with a real exponent -> pow(x, y) and with an integer constant exponent ->
    repeated multiplication (what gfortran emits), dmax1/dmin1 -> gfortran's compare-and-select.

Every generated statement carries a `/* file:line */` tag of the statement it came from; no
reference source text is reproduced in the output.

usage: f90_to_c.py [-D NAME] [--prefix ref_] [--skip proc,proc] -o out.c in1.f90 [in2.f90 ...]
"""
import re
import sys
import os
from collections import OrderedDict

# ------------------------------------------------------------------------------------------ lexer

TOK_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[edED][+-]?\d+)?(?:_\w+)?)
  | (?P<dotop>\.(?:not|and|or|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
  | (?P<name>[A-Za-z_]\w*)
  | (?P<op>\*\*|==|/=|<=|>=|=>|::|//|[-+*/()=,:%<>])
""", re.X | re.I)


class Tok:
    __slots__ = ("k", "v")

    def __init__(self, k, v):
        self.k, self.v = k, v

    def __repr__(self):
        return "%s:%s" % (self.k, self.v)


class F90Error(Exception):
    pass


def strip_comment(line):
    """cut a trailing `! comment` that is outside a string literal"""
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def logical_lines(text, defines):
    """yield (lineno, text) of logical statements: cpp #if/#else/#endif resolved, comments cut,
    `&` continuations joined"""
    out, cur, cur_no = [], "", None
    active = [True]
    for no, raw in enumerate(text.split("\n"), 1):
        s = raw.strip()
        if s.startswith("#"):
            m = re.match(r"#\s*(if|ifdef|else|endif)\b\s*(\w+)?", s)
            if not m:
                continue
            if m.group(1) in ("if", "ifdef"):
                active.append(active[-1] and bool(defines.get(m.group(2), 0)))
            elif m.group(1) == "else":
                top = active.pop()
                active.append(active[-1] and not top)
            else:
                active.pop()
            continue
        if not active[-1]:
            continue
        line = strip_comment(raw).rstrip()
        if not line.strip():
            continue
        body = line.strip()
        if cur:
            if body.startswith("&"):
                body = body[1:]
        else:
            cur_no = no
        if body.endswith("&"):
            cur += body[:-1] + " "
            continue
        cur += body
        out.append((cur_no, cur))
        cur = ""
    if cur:
        out.append((cur_no, cur))
    return out


def tokenize(s, where):
    toks, i, n = [], 0, len(s)
    while i < n:
        ch = s[i]
        if ch in "'\"":
            j, buf = i + 1, []
            while True:
                if j >= n:
                    raise F90Error("%s: unterminated string" % where)
                if s[j] == ch:
                    if j + 1 < n and s[j + 1] == ch:
                        buf.append(ch)
                        j += 2
                        continue
                    break
                buf.append(s[j])
                j += 1
            toks.append(Tok("str", "".join(buf)))
            i = j + 1
            continue
        m = TOK_RE.match(s, i)
        if not m:
            raise F90Error("%s: cannot tokenize at %r" % (where, s[i:i + 20]))
        i = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        if k == "num":
            # `1.eq.` style: a trailing '.' followed by a dot-operator belongs to the operator
            if v.endswith(".") and re.match(r"(?:not|and|or|eq|ne|lt|le|gt|ge)\.", s[i:], re.I):
                v = v[:-1]
                i -= 1
            toks.append(Tok("num", v))
        elif k == "dotop":
            toks.append(Tok("op", v.lower()))
        elif k == "name":
            toks.append(Tok("name", v.lower()))
        else:
            toks.append(Tok("op", v))
    return toks


# -------------------------------------------------------------------------------------------- AST

class Node:
    pass


class Num(Node):
    def __init__(self, text):
        self.text = text


class Str(Node):
    def __init__(self, s):
        self.s = s


class Logical(Node):
    def __init__(self, v):
        self.v = v


class Name(Node):
    def __init__(self, name):
        self.name = name


class Comp(Node):        # base % name
    def __init__(self, base, name):
        self.base, self.name = base, name


class Apply(Node):       # base(args): array element / section / function call
    def __init__(self, base, args, kw):
        self.base, self.args, self.kw = base, args, kw


class Range(Node):       # lo:hi (either may be None)
    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi


class Un(Node):
    def __init__(self, op, x):
        self.op, self.x = op, x


class Bin(Node):
    def __init__(self, op, l, r):
        self.op, self.l, self.r = op, l, r


class Paren(Node):
    def __init__(self, x):
        self.x = x


# ----------------------------------------------------------------------------------------- parser

REL = {"==": "==", "/=": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=",
       ".eq.": "==", ".ne.": "!=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


class ExprParser:
    def __init__(self, toks, where):
        self.t, self.i, self.where = toks, 0, where

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else Tok("eof", "")

    def at(self, v):
        p = self.peek()
        return p.k in ("op", "name") and p.v == v

    def at_op(self, v):
        p = self.peek()
        return p.k == "op" and p.v == v

    def take(self, v=None):
        p = self.peek()
        if v is not None and p.v != v:
            raise F90Error("%s: expected %r, got %r" % (self.where, v, p.v))
        self.i += 1
        return p

    def done(self):
        return self.i >= len(self.t)

    # Fortran 90 expression grammar (R723 ff.), lowest to highest precedence
    def expr(self):
        l = self.and_operand_chain()
        while self.at_op(".or."):
            self.take()
            l = Bin("||", l, self.and_operand_chain())
        return l

    def and_operand_chain(self):
        l = self.not_operand()
        while self.at_op(".and."):
            self.take()
            l = Bin("&&", l, self.not_operand())
        return l

    def not_operand(self):
        if self.at_op(".not."):
            self.take()
            return Un("!", self.not_operand())
        return self.rel()

    def rel(self):
        l = self.concat()
        p = self.peek()
        if p.k == "op" and p.v in REL:
            self.take()
            return Bin(REL[p.v], l, self.concat())
        return l

    def concat(self):
        l = self.additive()
        while self.at_op("//"):
            self.take()
            l = Bin("//", l, self.additive())
        return l

    def additive(self):
        if self.at_op("-") or self.at_op("+"):
            op = self.take().v
            l = self.term()
            if op == "-":
                l = Un("-", l)      # the sign applies to the whole first term: -a*b == -(a*b)
        else:
            l = self.term()
        while self.at_op("+") or self.at_op("-"):
            op = self.take().v
            l = Bin(op, l, self.term())
        return l

    def term(self):
        l = self.factor()
        while self.at_op("*") or self.at_op("/"):
            op = self.take().v
            l = Bin(op, l, self.factor())
        return l

    def factor(self):
        l = self.primary()
        if self.at_op("**"):
            self.take()
            if self.at_op("-"):
                self.take()
                r = Un("-", self.factor())
            else:
                r = self.factor()      # right-associative
            return Bin("**", l, r)
        return l

    def primary(self):
        p = self.take()
        if p.k == "num":
            return Num(p.v)
        if p.k == "str":
            return Str(p.v)
        if p.k == "op" and p.v == ".true.":
            return Logical(1)
        if p.k == "op" and p.v == ".false.":
            return Logical(0)
        if p.k == "op" and p.v == "(":
            e = self.expr()
            self.take(")")
            return Paren(e)
        if p.k == "name":
            node = Name(p.v)
            return self.postfix(node)
        raise F90Error("%s: unexpected token %r in expression" % (self.where, p.v))

    def postfix(self, node):
        while True:
            if self.at_op("("):
                self.take()
                args, kw = self.arglist()
                self.take(")")
                node = Apply(node, args, kw)
            elif self.at_op("%"):
                self.take()
                node = Comp(node, self.take().v)
            else:
                return node

    def arglist(self):
        args, kw = [], OrderedDict()
        if self.at_op(")"):
            return args, kw
        while True:
            if self.peek().k == "name" and self.peek(1).k == "op" and self.peek(1).v == "=":
                k = self.take().v
                self.take("=")
                kw[k] = self.expr_or_star()
            else:
                args.append(self.section_or_expr())
            if self.at_op(","):
                self.take()
                continue
            return args, kw

    def expr_or_star(self):
        if self.at_op("*"):
            self.take()
            return Name("*")
        return self.expr()

    def section_or_expr(self):
        if self.at_op("*"):
            nxt = self.peek(1)
            if nxt.k == "op" and nxt.v in (",", ")"):
                self.take()
                return Name("*")
        lo = None
        if not self.at_op(":"):
            lo = self.expr()
            if not self.at_op(":"):
                return lo
        self.take(":")
        hi = None
        if not (self.at_op(",") or self.at_op(")")):
            hi = self.expr()
        return Range(lo, hi)


# ------------------------------------------------------------------------------------------ types

class Ty:
    """base in {r8, r4, int, logical, char, derived}; rank 0..2"""

    def __init__(self, base, rank=0, charlen=None, derived=None):
        self.base, self.rank, self.charlen, self.derived = base, rank, charlen, derived

    def scalar(self):
        return Ty(self.base, 0, self.charlen, self.derived)

    def ctype(self, prefix):
        return {"r8": "double", "r4": "float", "int": "int", "logical": "int"}.get(self.base) or \
            ("%s%s" % (prefix, self.derived) if self.base == "derived" else "char")


class Sym:
    """a data entity.  storage: 'scalar' | 'desc' (descriptor struct) | 'explicit' (pointer + dims)
    | 'fixed' (C array).  is_ptr: the C name is a pointer to the object (dummy arguments)."""

    def __init__(self, name, ty, storage="scalar", is_ptr=False, dims=None, intent=None,
                 optional=False, param=None, cname=None, allocatable=False, automatic=False):
        self.name, self.ty, self.storage, self.is_ptr = name, ty, storage, is_ptr
        self.dims, self.intent, self.optional, self.param = dims, intent, optional, param
        self.cname = cname or name
        self.allocatable, self.automatic = allocatable, automatic


class Proc:
    def __init__(self, kind, name, args, result, elemental, line, src):
        self.kind, self.name, self.args, self.result = kind, name, args, result
        self.elemental, self.line, self.src = elemental, line, src
        self.decls = []      # (lineno, toks)
        self.body = []       # (lineno, toks)
        self.syms = OrderedDict()
        self.namelists = OrderedDict()


class Module:
    def __init__(self, name, is_program=False):
        self.name, self.is_program = name, is_program
        self.types = OrderedDict()      # name -> OrderedDict(comp -> Sym)
        self.syms = OrderedDict()       # parameters + module variables
        self.procs = OrderedDict()
        self.uses = []
        self.main = None                # Proc for a program


C_KEYWORDS = {"double", "float", "int", "char", "long", "short", "void", "if", "else", "for", "do",
              "while", "return", "struct", "union", "enum", "static", "const", "signed", "unsigned",
              "switch", "case", "default", "break", "continue", "goto", "sizeof", "typedef",
              "register", "auto", "extern", "volatile", "inline", "restrict", "exit", "abs", "exp",
              "pow", "sqrt", "acos", "tanh", "fabs", "free", "malloc", "main", "time", "y0", "y1",
              "j0", "j1", "jn", "yn", "index", "remove", "div", "open", "close", "read", "write"}


def cident(name):
    return name + "_" if name in C_KEYWORDS else name


# ------------------------------------------------------------------------------- front end (units)

DECL_HEADS = ("integer", "double", "real", "logical", "character", "type")


class FrontEnd:
    """splits the logical lines of one file into modules / programs / procedures and parses
    declarations; executable statements are parsed lazily by the generator"""

    def __init__(self, path, defines):
        self.path = path
        self.base = os.path.basename(path)
        with open(path, "r", encoding="latin-1") as f:
            text = f.read()
        self.lines = [(no, tokenize(s, "%s:%d" % (self.base, no))) for no, s in logical_lines(text, defines)]
        self.units = []
        self.parse_units()

    def where(self, no):
        return "%s:%d" % (self.base, no)

    def parse_units(self):
        cur, proc, tdef = None, None, None
        for no, t in self.lines:
            h = t[0].v if t[0].k == "name" else ""
            h2 = t[1].v if len(t) > 1 and t[1].k == "name" else ""
            # ---- end of a unit / procedure / type
            if h.startswith("end") and (h[3:] or h2) in ("module", "program", "subroutine", "function", "type"):
                what = h[3:] or h2
                if what == "type":
                    tdef = None
                elif what in ("module", "program"):
                    cur, proc = None, None
                else:
                    proc = None
                continue
            if h == "end" and len(t) == 1:
                if proc is not None and proc.kind != "program":
                    proc = None
                else:
                    cur, proc = None, None
                continue
            # ---- start of a unit
            if h == "module" and h2 and cur is None:
                cur = Module(h2)
                self.units.append(cur)
                continue
            if h == "program":
                cur = Module(h2, is_program=True)
                cur.main = proc = Proc("program", h2, [], None, False, no, self.base)
                self.units.append(cur)
                continue
            if cur is None:
                raise F90Error("%s: statement outside a program unit" % self.where(no))
            if h == "contains":
                continue
            # ---- derived-type definition: `type name` (a declaration is `type(name) ...`)
            if h == "type" and h2 and tdef is None:
                tdef = OrderedDict()
                cur.types[h2] = tdef
                continue
            if tdef is not None:
                for s in self.parse_decl(no, t, component=True):
                    tdef[s.name] = s
                continue
            # ---- procedure header
            names = [x.v for x in t[:3] if x.k == "name"]
            kind = "subroutine" if "subroutine" in names else ("function" if "function" in names else None)
            if kind and h in ("subroutine", "function", "elemental", "pure", "recursive") and proc is None:
                k = [x.v for x in t].index(kind)
                p = ExprParser(t[k + 2:], self.where(no))
                args = []
                if p.at_op("("):
                    p.take()
                    while not p.at_op(")"):
                        args.append(p.take().v)
                        if p.at_op(","):
                            p.take()
                    p.take(")")
                result = None
                if p.at("result"):
                    p.take()
                    p.take("(")
                    result = p.take().v
                    p.take(")")
                name = t[k + 1].v
                if kind == "function" and result is None:
                    result = name
                proc = Proc(kind, name, args, result, "elemental" in names, no, self.base)
                cur.procs[name] = proc
                continue
            # ---- specification statements
            if h in ("implicit", "private", "public", "save"):
                continue
            if h == "use":
                cur.uses.append(t[1].v)
                continue
            if h == "namelist":
                proc.namelists[t[2].v] = [x.v for x in t[4:] if x.k == "name"]
                continue
            if self.is_decl(t) and (proc is None or not proc.body):
                if proc is None:
                    for s in self.parse_decl(no, t, component=False):
                        cur.syms[s.name] = s
                else:
                    proc.decls.append((no, t))
                continue
            if proc is None:
                raise F90Error("%s: executable statement outside a procedure" % self.where(no))
            proc.body.append((no, t))

    @staticmethod
    def is_decl(t):
        h = t[0].v if t[0].k == "name" else ""
        if h not in DECL_HEADS:
            return False
        if h == "type":
            return len(t) > 1 and t[1].k == "op" and t[1].v == "("
        if h == "real" and len(t) > 1 and t[1].k == "op" and t[1].v == "=":
            return False
        return True

    def parse_decl(self, no, t, component):
        """type-spec [, attrs] [::] entity-list  ->  [Sym] (storage decided by the caller scope)"""
        p = ExprParser(t, self.where(no))
        h = p.take().v
        charlen, derived = None, None
        if h == "double":
            p.take("precision")
            base = "r8"
        elif h == "real":
            base = "r4"
            if p.at_op("("):
                p.take()
                args, kw = p.arglist()
                p.take(")")
                k = kw.get("kind", args[0] if args else None)
                if isinstance(k, Name) and k.name == "dp" or isinstance(k, Num) and k.text == "8":
                    base = "r8"
        elif h == "integer":
            base = "int"
        elif h == "logical":
            base = "logical"
        elif h == "character":
            base = "char"
            charlen = 1
            if p.at_op("("):
                p.take()
                args, kw = p.arglist()
                p.take(")")
                ln = kw.get("len", args[0] if args else None)
                charlen = "*" if isinstance(ln, Name) and ln.name == "*" else int(ln.text)
        elif h == "type":
            p.take("(")
            derived = p.take().v
            p.take(")")
            base = "derived"
        else:
            raise F90Error("%s: unknown type %r" % (self.where(no), h))
        attrs = {}
        while p.at_op(","):
            p.take()
            a = p.take().v
            if a == "dimension":
                p.take("(")
                dims, _ = p.arglist()
                p.take(")")
                attrs["dimension"] = dims
            elif a == "intent":
                p.take("(")
                w = p.take().v
                if w == "in" and (p.at("out")):
                    p.take()
                    w = "inout"
                p.take(")")
                attrs["intent"] = w
            else:
                attrs[a] = True
        if p.at_op("::"):
            p.take()
        out = []
        while not p.done():
            name = p.take().v
            dims = attrs.get("dimension")
            if p.at_op("("):
                p.take()
                dims, _ = p.arglist()
                p.take(")")
            init = None
            if p.at_op("="):
                p.take()
                init = p.expr()
            rank = len(dims) if dims else 0
            s = Sym(name, Ty(base, rank, charlen, derived), dims=dims, intent=attrs.get("intent"),
                    optional="optional" in attrs, param=init if "parameter" in attrs else None,
                    cname=cident(name), allocatable="allocatable" in attrs)
            s.init = init
            s.line = no
            out.append(s)
            if p.at_op(","):
                p.take()
        return out


# ------------------------------------------------------------------------------------- generator

INTRINSIC_R8 = {"dexp": "exp", "exp": "exp", "dacos": "acos", "acos": "acos", "dsqrt": "sqrt",
                "sqrt": "sqrt", "tanh": "tanh", "dtanh": "tanh", "dlog": "log", "log": "log",
                "dcos": "cos", "cos": "cos", "dsin": "sin", "sin": "sin"}


class Gen:
    def __init__(self, files, prefix, skip):
        self.files, self.prefix, self.skip = files, prefix, set(skip)
        self.modules = OrderedDict()
        for fe in files:
            for u in fe.units:
                self.modules[u.name] = u
        self.out = []
        self.kinds = {}          # integer parameters that name a real kind, e.g. dp
        self.tmp = 0

    # ---- symbol / type lookup
    def visible_modules(self, mod):
        seen, order = set(), []

        def rec(m):
            if m.name in seen:
                return
            seen.add(m.name)
            for u in m.uses:
                if u in self.modules:
                    rec(self.modules[u])
            order.append(m)
        rec(mod)
        return order

    def find_type(self, name):
        for m in self.modules.values():
            if name in m.types:
                return m.types[name]
        raise F90Error("unknown derived type %s" % name)

    def find_proc(self, name):
        for m in self.visible_modules(self.cur_mod):
            if name in m.procs:
                return m.procs[name]
        return None

    def lookup(self, name):
        if name in self.cur_proc.syms:
            return self.cur_proc.syms[name]
        for m in self.visible_modules(self.cur_mod):
            if name in m.syms:
                return m.syms[name]
        return None

    def newtmp(self, stem="t"):
        self.tmp += 1
        return "%s%d_" % (stem, self.tmp)

    # ---- emit helpers
    def emit(self, s, ind=None):
        self.out.append(("    " * (self.ind if ind is None else ind)) + s)

    def tag(self, no):
        return "/* %s:%d */" % (self.cur_src, no)

    # ------------------------------------------------------------------ top level
    def generate(self):
        self.ind = 0
        e = self.emit
        e("/* GENERATED by oracle/f90_to_c.py from: %s" % ", ".join(fe.base for fe in self.files))
        e(" * Do not edit, do not commit: rebuilt by `make -C oracle ref`.  TEST INFRASTRUCTURE ONLY. */")
        e('#include "f90_rt.h"')
        e("")
        mods = list(self.modules.values())
        # kinds
        for m in mods:
            for s in m.syms.values():
                if s.param is not None and s.ty.base == "int" and isinstance(s.param, Apply) \
                        and isinstance(s.param.base, Name) and s.param.base.name == "kind":
                    self.kinds[s.name] = "r8" if re.search(r"[dD]", s.param.args[0].text) else "r4"
        # derived types
        for m in mods:
            for tname, comps in m.types.items():
                self.cur_mod, self.cur_src = m, ""
                e("typedef struct %s%s {" % (self.prefix, tname))
                for c in comps.values():
                    self.fix_storage(c, component=True)
                    e("    %s;" % self.c_decl(c, component=True))
                e("} %s%s;" % (self.prefix, tname))
                e("")
        # reflection tables (for the ctypes binding: component name -> kind, offset)
        for m in mods:
            for tname, comps in m.types.items():
                e("static const f_compinfo %scomps_%s[] = {" % (self.prefix, tname))
                for c in comps.values():
                    e('    {"%s", %s, offsetof(%s%s, %s), %d},' % (
                        c.name, self.comp_kind(c), self.prefix, tname, c.cname,
                        int(c.ty.charlen) if c.ty.base == "char" else 0))
                e("    {0, 0, 0, 0}")
                e("};")
        e("const f_typeinfo %stypes[] = {" % self.prefix)
        for m in mods:
            for tname in m.types:
                e('    {"%s", sizeof(%s%s), %scomps_%s},' % (tname, self.prefix, tname, self.prefix, tname))
        e("    {0, 0, 0}")
        e("};")
        e("")
        # module parameters and variables
        for m in mods:
            self.cur_mod = m
            self.cur_proc = Proc("none", "", [], None, False, 0, "")
            for s in m.syms.values():
                if s.name in self.kinds:
                    continue
                self.fix_storage(s, component=False)
                src = [fe.base for fe in self.files if m in fe.units][0]
                if s.param is not None:
                    txt, _ = self.expr(s.param, None)
                    e("static const %s %s = %s; /* %s:%d */" % (s.ty.ctype(self.prefix), s.cname, txt, src, s.line))
                else:
                    e("%s; /* %s:%d */" % (self.c_decl(s, component=True), src, s.line))
        e("")
        # prototypes, then bodies
        for m in mods:
            self.cur_mod = m
            for p in m.procs.values():
                if p.name in self.skip:
                    continue
                self.prepare_proc(p)
                e(self.c_signature(p) + ";")
        e("")
        for m in mods:
            self.cur_mod = m
            for p in m.procs.values():
                if p.name in self.skip:
                    continue
                self.gen_proc(p)
            if m.main is not None:
                self.prepare_proc(m.main)
                self.gen_proc(m.main)
        return "\n".join(self.out) + "\n"

    def comp_kind(self, c):
        b = {"r8": "D", "r4": "F", "int": "I", "logical": "L", "char": "S", "derived": "T"}[c.ty.base]
        if c.storage == "desc":
            return "F_K_%s_DESC" % b
        if c.storage == "fixed":
            return "F_K_%s_FIXED" % b
        return "F_K_%s" % b

    def fix_storage(self, s, component):
        if s.ty.rank == 0:
            s.storage = "scalar"
        elif s.allocatable or (s.dims and all(isinstance(d, Range) and d.lo is None and d.hi is None for d in s.dims)):
            s.storage = "desc"
        elif all(isinstance(d, Num) for d in s.dims):
            s.storage = "fixed"
        else:
            s.storage = "explicit"

    def c_decl(self, s, component):
        ct = s.ty.ctype(self.prefix)
        if s.ty.base == "char":
            ln = int(s.ty.charlen)
            if s.storage == "fixed":
                return "char %s[%s][%d]" % (s.cname, "][".join(d.text for d in s.dims), ln)
            return "char %s[%d]" % (s.cname, ln)
        if s.storage == "scalar":
            return "%s %s" % (ct, s.cname)
        if s.storage == "desc":
            return "f_arr_%s %s" % ({"double": "d", "int": "i"}[ct], s.cname)
        if s.storage == "fixed":
            return "%s %s[%s]" % (ct, s.cname, "*".join(d.text for d in s.dims))
        raise F90Error("cannot declare %s" % s.name)

    # ------------------------------------------------------------------ procedures
    def prepare_proc(self, p):
        if p.syms:
            return
        fe = [f for f in self.files if f.base == p.src][0]
        for no, t in p.decls:
            for s in fe.parse_decl(no, t, component=False):
                self.fix_storage(s, component=False)
                if s.name in p.args:
                    s.is_dummy = True
                    if s.storage == "scalar" and s.ty.base != "char":
                        s.is_ptr = True
                    elif s.ty.base == "derived":
                        s.is_ptr = True
                    if s.storage == "fixed":
                        s.storage = "explicit"
                else:
                    s.is_dummy = False
                    if s.storage == "explicit":
                        s.storage, s.automatic = "desc", True
                p.syms[s.name] = s
        for a in p.args:
            if a not in p.syms:
                raise F90Error("%s: dummy %s of %s not declared" % (p.src, a, p.name))
        if p.kind == "function":
            r = p.syms[p.result]
            r.is_dummy = True
            r.is_result = True
            if r.storage == "scalar":
                r.is_ptr = False
            elif r.storage == "desc" and r.automatic:
                r.storage, r.automatic = "explicit", False

    def c_param(self, s):
        ct = s.ty.ctype(self.prefix)
        if s.ty.base == "char":
            if s.ty.rank == 0:
                return "char *%s, long %s_len" % (s.cname, s.cname)
            return "char *%s, long %s_n" % (s.cname, s.cname)
        if s.storage == "scalar":
            return "%s *%s" % (ct, s.cname)
        if s.storage == "desc":
            return "f_arr_%s %s" % ({"double": "d", "int": "i"}[ct], s.cname)
        if s.storage == "explicit":
            return "%s *%s" % (ct, s.cname)
        raise F90Error("param %s" % s.name)

    def c_signature(self, p):
        params = [self.c_param(p.syms[a]) for a in p.args]
        ret = "void"
        if p.kind == "function":
            r = p.syms[p.result]
            if r.ty.rank == 0:
                ret = r.ty.ctype(self.prefix)
            else:
                params.append("%s *%s" % (r.ty.ctype(self.prefix), r.cname))
        return "%s %s%s(%s)" % (ret, self.prefix, p.name, ", ".join(params) or "void")

    def gen_proc(self, p):
        self.cur_proc, self.cur_src = p, p.src
        self.ind = 0
        e = self.emit
        e("")
        e("/* %s %s, %s:%d */" % (p.kind, p.name, p.src, p.line))
        if p.kind == "program":
            e("int main(int argc, char **argv)")
        else:
            e(self.c_signature(p))
        e("{")
        self.ind = 1
        if p.kind == "program":
            e("f_set_args(argc, argv);")
        # locals
        autos = []
        for s in p.syms.values():
            if getattr(s, "is_dummy", False) and not (getattr(s, "is_result", False) and s.ty.rank == 0):
                continue
            if s.param is not None:
                txt, _ = self.expr(s.param, None)
                e("const %s %s = %s;" % (s.ty.ctype(self.prefix), s.cname, txt))
            elif s.ty.base == "derived":
                e("%s%s %s; memset(&%s, 0, sizeof %s);" % (self.prefix, s.ty.derived, s.cname, s.cname, s.cname))
            elif s.storage == "desc" and s.automatic:
                autos.append(s)
                e("%s = {0, 0, 0};" % self.c_decl(s, False))
            elif s.storage == "desc":
                e("%s = {0, 0, 0};" % self.c_decl(s, False))
            elif s.ty.base == "char":
                e("%s; memset(%s, ' ', sizeof %s);" % (self.c_decl(s, False), s.cname, s.cname))
            else:
                e("%s;" % self.c_decl(s, False))
        e("long i_; (void)i_;")
        for s in autos:
            dims = [self.expr(d, None)[0] for d in s.dims]
            while len(dims) < 2:
                dims.append("1")
            e("f_alloc_%s(&%s, %s, %s);" % ("d" if s.ty.base == "r8" else "i", s.cname, dims[0], dims[1]))
        # namelist tables
        for grp, names in p.namelists.items():
            e("f_nml_item nml_%s[] = {" % grp)
            for n in names:
                s = self.lookup(n)
                kind = {"r8": "F_K_D", "int": "F_K_I", "char": "F_K_S", "logical": "F_K_L", "r4": "F_K_F"}[s.ty.base]
                nelem = "1"
                if s.ty.rank:
                    nelem = "*".join(d.text for d in s.dims)
                addr = s.cname if (s.ty.rank or s.ty.base == "char") else "&" + s.cname
                e('    {"%s", %s, (void *)%s, %s, %d},' % (n, kind, addr, nelem,
                                                         int(s.ty.charlen) if s.ty.base == "char" else 0))
            e("};")
        self.has_return = False
        self.block(p.body, 0, len(p.body))
        if self.has_return:
            e("f_exit_: ;", ind=0)
        for s in autos:
            e("free(%s.a);" % s.cname)
        if p.kind == "function" and p.syms[p.result].ty.rank == 0:
            e("return %s;" % p.syms[p.result].cname)
        if p.kind == "program":
            e("return 0;")
        self.ind = 0
        e("}")

    # ------------------------------------------------------------------ statements
    def head(self, t):
        return t[0].v if t[0].k == "name" else ""

    def block(self, L, i, j):
        """generate statements L[i:j]"""
        while i < j:
            i = self.stmt(L, i, j)

    def find_end(self, L, i, j, opens, closes):
        """index of the statement closing the construct opened at L[i] (same nesting level)"""
        depth = 0
        k = i
        while k < j:
            t = L[k][1]
            if opens(t):
                depth += 1
            elif closes(t):
                depth -= 1
                if depth == 0:
                    return k
            k += 1
        raise F90Error("%s:%d: unterminated construct" % (self.cur_src, L[i][0]))

    @staticmethod
    def is_if_then(t):
        return t[0].k == "name" and t[0].v == "if" and t[-1].k == "name" and t[-1].v == "then"

    @staticmethod
    def is_end(t, what):
        if t[0].k != "name":
            return False
        if t[0].v == "end" + what:
            return True
        return t[0].v == "end" and len(t) > 1 and t[1].v == what

    def stmt(self, L, i, j):
        no, t = L[i]
        h = self.head(t)
        e = self.emit
        W = "%s:%d" % (self.cur_src, no)
        # ---- block if
        if self.is_if_then(t):
            end = self.find_end(L, i, j, self.is_if_then, lambda x: self.is_end(x, "if"))
            # split into branches at depth 1
            branches, depth, k, start = [], 0, i, i
            conds = []
            while k <= end:
                tk = L[k][1]
                if self.is_if_then(tk):
                    depth += 1
                    if depth == 1:
                        conds.append(("if", tk[1:-1], L[k][0]))
                        start = k + 1
                elif self.is_end(tk, "if"):
                    if depth == 1:
                        branches.append((start, k))
                    depth -= 1
                elif depth == 1 and (self.head(tk) == "elseif" or (self.head(tk) == "else" and
                                     (len(tk) == 1 or tk[1].v == "if"))):
                    branches.append((start, k))
                    start = k + 1
                    if (self.head(tk) == "else" and len(tk) > 1 and tk[1].v == "if") or self.head(tk) == "elseif":
                        off = 2 if self.head(tk) == "else" else 1
                        conds.append(("else if", tk[off:-1], L[k][0]))
                    else:
                        conds.append(("else", None, L[k][0]))
                k += 1
            for (kw, ct, cno), (a, b) in zip(conds, branches):
                if ct is not None:
                    c, _ = self.expr(ExprParser(ct, W).expr(), None)
                    e("%s%s (%s) { %s" % ("} " if kw != "if" else "", kw, c, self.tag(cno)))
                else:
                    e("} else { %s" % self.tag(cno))
                self.ind += 1
                self.block(L, a, b)
                self.ind -= 1
            e("}")
            return end + 1
        # ---- one-line if
        if h == "if":
            p = ExprParser(t[1:], W)
            p.take("(")
            c = p.expr()
            p.take(")")
            ctext, _ = self.expr(c, None)
            e("if (%s) { %s" % (ctext, self.tag(no)))
            self.ind += 1
            self.stmt([(no, t[1 + p.i:])], 0, 1)
            self.ind -= 1
            e("}")
            return i + 1
        # ---- do loop
        if h == "do":
            end = self.find_end(L, i, j, lambda x: self.head(x) == "do", lambda x: self.is_end(x, "do"))
            p = ExprParser(t[1:], W)
            var = p.take().v
            p.take("=")
            lo = p.expr()
            p.take(",")
            hi = p.expr()
            step = None
            if p.at_op(","):
                p.take()
                step = p.expr()
            v, _ = self.expr(Name(var), None)
            lo_c, hi_c = self.expr(lo, None)[0], self.expr(hi, None)[0]
            hv = self.newtmp("hi")
            if step is None:
                e("{ const int %s = %s; for (%s = %s; %s <= %s; ++%s) { %s" % (hv, hi_c, v, lo_c, v, hv, v, self.tag(no)))
            else:
                st = self.expr(step, None)[0]
                e("{ const int %s = %s; for (%s = %s; %s <= %s; %s += %s) { %s" % (hv, hi_c, v, lo_c, v, hv, v, st, self.tag(no)))
            self.ind += 1
            self.block(L, i + 1, end)
            self.ind -= 1
            e("} }")
            return end + 1
        # ---- where construct
        if h == "where":
            end = self.find_end(L, i, j, lambda x: self.head(x) == "where" and self.where_is_construct(x),
                                lambda x: self.is_end(x, "where"))
            self.gen_where(L, i, end)
            return end + 1
        # ---- select case
        if h == "select":
            end = self.find_end(L, i, j, lambda x: self.head(x) == "select", lambda x: self.is_end(x, "select"))
            p = ExprParser(t[2:], W)
            p.take("(")
            sel = p.expr()
            p.take(")")
            first, k = True, i + 1
            seltxt, selty = self.expr(sel, None)
            cases = []
            while k < end:
                if self.head(L[k][1]) == "case":
                    cases.append(k)
                k += 1
            cases.append(end)
            for a, b in zip(cases[:-1], cases[1:]):
                ct = L[a][1]
                if ct[1].k == "name" and ct[1].v == "default":
                    e("%s{ %s" % ("" if first else "} else ", self.tag(L[a][0])))
                else:
                    q = ExprParser(ct[1:], W)
                    q.take("(")
                    vals, _ = q.arglist()
                    q.take(")")
                    tests = []
                    for v in vals:
                        vt, vty = self.expr(v, None)
                        tests.append("f_str_eq(%s, %s)" % (seltxt, vt) if selty.base == "char"
                                     else "(%s) == (%s)" % (seltxt, vt))
                    e("%sif (%s) { %s" % ("" if first else "} else ", " || ".join(tests), self.tag(L[a][0])))
                first = False
                self.ind += 1
                self.block(L, a + 1, b)
                self.ind -= 1
            e("}")
            return end + 1
        # ---- simple statements
        if h == "call":
            self.gen_call(no, t)
        elif h == "return" and len(t) == 1:
            self.has_return = True
            e("goto f_exit_; %s" % self.tag(no))
        elif h == "stop" and len(t) == 1:
            e("f_stop(); %s" % self.tag(no))
        elif h == "exit" and len(t) == 1:
            e("break; %s" % self.tag(no))
        elif h == "allocate" and t[1].v == "(":
            p = ExprParser(t[1:], W)
            p.take("(")
            a = p.primary()
            dims = [self.expr(d, None)[0] for d in a.args]
            while len(dims) < 2:
                dims.append("1")
            txt, ty = self.designator(a.base)
            e("f_alloc_%s(&%s, %s, %s); %s" % ("d" if ty.base == "r8" else "i", txt, dims[0], dims[1], self.tag(no)))
        elif h == "deallocate" and t[1].v == "(":
            p = ExprParser(t[1:], W)
            p.take("(")
            a = p.primary()
            txt, ty = self.designator(a)
            e("f_dealloc_%s(&%s); %s" % ("d" if ty.base == "r8" else "i", txt, self.tag(no)))
        elif h in ("open", "close", "read", "write", "inquire") and t[1].v == "(":
            self.gen_io(no, t)
        else:
            self.gen_assign(no, t)
        return i + 1

    def where_is_construct(self, t):
        # `where (mask)` alone on the line is a construct; `where (mask) a = b` is a statement
        depth = 0
        for k, x in enumerate(t[1:], 1):
            if x.k == "op" and x.v == "(":
                depth += 1
            elif x.k == "op" and x.v == ")":
                depth -= 1
                if depth == 0:
                    return k == len(t) - 1
        return True

    def gen_where(self, L, i, end):
        e = self.emit
        no, t = L[i]
        W = "%s:%d" % (self.cur_src, no)
        p = ExprParser(t[1:], W)
        p.take("(")
        mask = p.expr()
        p.take(")")
        n = self.extent(mask)
        m = self.newtmp("m")
        ctext, _ = self.expr(mask, "i_")
        e("{ %s" % self.tag(no))
        self.ind += 1
        e("const long n_ = %s;" % n)
        e("unsigned char *%s = (unsigned char *)f_malloc((size_t)n_);" % m)
        e("for (i_ = 0; i_ < n_; ++i_) %s[i_] = (unsigned char)(%s);" % (m, ctext))
        sense = "%s[i_]" % m
        for k in range(i + 1, end):
            kno, kt = L[k]
            hk = self.head(kt)
            if hk in ("elsewhere", "else"):
                if len(kt) > (1 if hk == "elsewhere" else 2):
                    raise F90Error("%s:%d: masked elsewhere not supported" % (self.cur_src, kno))
                sense = "!%s[i_]" % m
                continue
            self.gen_assign(kno, kt, mask=sense, n_override="n_")
        e("free(%s);" % m)
        self.ind -= 1
        e("}")

    # ---- assignment
    def split_assign(self, t, W):
        depth = 0
        for k, x in enumerate(t):
            if x.k == "op":
                if x.v == "(":
                    depth += 1
                elif x.v == ")":
                    depth -= 1
                elif x.v == "=" and depth == 0:
                    return ExprParser(t[:k], W).expr(), ExprParser(t[k + 1:], W).expr()
        raise F90Error("%s: not an assignment: %r" % (W, t))

    def gen_assign(self, no, t, mask=None, n_override=None):
        e = self.emit
        W = "%s:%d" % (self.cur_src, no)
        lhs, rhs = self.split_assign(t, W)
        lty = self.typeof(lhs)
        if lty.base == "char":
            if lty.rank == 0:
                l, _ = self.expr(lhs, None)
                r, _ = self.expr(rhs, None)
                e("f_str_assign(%s, %s); %s" % (l, r, self.tag(no)))
            else:
                n = self.extent(lhs)
                l, _ = self.expr(lhs, "i_")
                r, _ = self.expr(rhs, "i_")
                e("for (i_ = 0; i_ < %s; ++i_) f_str_assign(%s, %s); %s" % (n, l, r, self.tag(no)))
            return
        if lty.rank == 0:
            pre = []
            self.pre = pre
            r, _ = self.expr(rhs, None)
            l, _ = self.expr(lhs, None)
            self.pre = None
            for s in pre:
                e(s)
            e("%s = %s; %s" % (l, r, self.tag(no)))
            return
        # array assignment: one loop per statement
        rty = self.typeof(rhs)
        if rty.rank and isinstance(rhs, Apply) and isinstance(rhs.base, Name) and \
                self.find_proc(rhs.base.name) is not None and not self.find_proc(rhs.base.name).elemental:
            # array-valued function: result written through the trailing pointer argument
            callee = self.find_proc(rhs.base.name)
            self.prepare_proc(callee)
            args = self.actuals(callee, rhs.args, None)
            ltxt, _ = self.designator(lhs)
            e("%s%s(%s, %s.a); %s" % (self.prefix, callee.name, ", ".join(args), ltxt, self.tag(no)))
            return
        n = n_override or self.extent(lhs)
        l, _ = self.expr(lhs, "i_")
        r, _ = self.expr(rhs, "i_")
        if mask:
            e("for (i_ = 0; i_ < %s; ++i_) if (%s) %s = %s; %s" % (n, mask, l, r, self.tag(no)))
        else:
            e("_Pragma(\"GCC ivdep\") for (i_ = 0; i_ < %s; ++i_) %s = %s; %s" % (n, l, r, self.tag(no)))

    # ---- call
    def gen_call(self, no, t):
        e = self.emit
        W = "%s:%d" % (self.cur_src, no)
        p = ExprParser(t[1:], W)
        name = p.take().v
        args = []
        if p.at_op("("):
            p.take()
            args, _ = p.arglist()
            p.take(")")
        callee = self.find_proc(name)
        if callee is None:
            # runtime-provided intrinsic subroutines
            if name in ("getarg", "get_command_argument"):
                a0, _ = self.expr(args[0], None)
                a1, _ = self.expr(args[1], None)
                e("f_getarg(%s, %s); %s" % (a0, a1, self.tag(no)))
            elif name == "cpu_time":
                a0, _ = self.expr(args[0], None)
                e("%s = f_cpu_time(); %s" % (a0, self.tag(no)))
            elif name == "exit":
                e("exit(0); %s" % self.tag(no))
            else:
                raise F90Error("%s: unknown procedure %s" % (W, name))
            return
        if callee.name in self.skip:
            e("/* call %s skipped (not translated) */ %s" % (callee.name, self.tag(no)))
            return
        self.prepare_proc(callee)
        ranks = [self.typeof(a).rank for a in args]
        if callee.elemental and any(ranks):
            first = args[[k for k, r in enumerate(ranks) if r][0]]
            n = self.extent(first)
            actual = self.actuals(callee, args, "i_")
            e("for (i_ = 0; i_ < %s; ++i_) %s%s(%s); %s" % (n, self.prefix, name, ", ".join(actual), self.tag(no)))
        else:
            pre = []
            self.pre = pre
            actual = self.actuals(callee, args, None)
            self.pre = None
            for s in pre:
                e(s)
            e("%s%s(%s); %s" % (self.prefix, name, ", ".join(actual), self.tag(no)))

    def actuals(self, callee, args, ix):
        out = []
        for k, dn in enumerate(callee.args):
            d = callee.syms[dn]
            if k >= len(args):
                if not d.optional:
                    raise F90Error("missing actual for %s of %s" % (dn, callee.name))
                out.append("0")
                continue
            a = args[k]
            aty = self.typeof(a)
            if d.ty.base == "char":
                if d.ty.rank == 0:
                    txt, _ = self.expr(a, ix)
                    out.append(txt)
                else:
                    txt, ty = self.designator(a)
                    out.append("(char *)%s, %s" % (txt, self.extent(a)))
            elif d.ty.base == "derived":
                txt, _ = self.designator(a)
                sym_ptr = txt.startswith("(*")
                out.append(txt[2:-1] if sym_ptr else "&" + txt)
            elif d.storage == "scalar":
                if self.is_designator(a) and aty.base == d.ty.base and not self.is_const(a):
                    txt, _ = self.expr(a, ix)
                    if txt.startswith("(*") and txt.endswith(")") and txt[2:-1].isidentifier():
                        out.append(txt[2:-1])
                    else:
                        out.append("&" + txt)
                else:
                    txt, _ = self.expr(a, ix)
                    out.append("&(%s){%s}" % (d.ty.ctype(self.prefix), txt))
            elif d.storage == "desc":
                txt, _ = self.designator(a)
                out.append(txt)
            elif d.storage == "explicit":
                txt, _ = self.designator(a)
                s = self.sym_of(a)
                out.append(txt if (s is not None and s.storage in ("explicit", "fixed")) else txt + ".a")
            else:
                raise F90Error("actual for %s" % dn)
        return out

    def is_const(self, node):
        if isinstance(node, Name):
            s = self.lookup(node.name)
            return s is not None and s.param is not None
        return False

    # ---- I/O (through the runtime)
    def gen_io(self, no, t):
        e = self.emit
        W = "%s:%d" % (self.cur_src, no)
        h = t[0].v
        p = ExprParser(t[1:], W)
        p.take("(")
        args, kw = p.arglist()
        p.take(")")
        items = []
        if not p.done():
            while True:
                items.append(p.section_or_expr())
                if p.at_op(","):
                    p.take()
                    continue
                break
        unit = kw.get("unit", args[0] if args else None)

        def unit_c():
            if isinstance(unit, Name) and unit.name == "*":
                return "-1"
            return self.expr(unit, None)[0]

        def strarg(node):
            return self.expr(node, None)[0] if node is not None else "0, 0"
        if h == "open":
            e("f_open(%s, %s, %s); %s" % (unit_c(), strarg(kw.get("file")), strarg(kw.get("action")), self.tag(no)))
            return
        if h == "close":
            e("f_close(%s); %s" % (unit_c(), self.tag(no)))
            return
        if h == "inquire":
            ex, _ = self.expr(kw["exist"], None)
            e("%s = f_inquire_exist(%s); %s" % (ex, strarg(kw.get("file")), self.tag(no)))
            return
        fmt = kw.get("fmt", args[1] if len(args) > 1 else None)
        if "nml" in kw:
            grp = kw["nml"].name
            e("f_read_nml(%s, \"%s\", nml_%s, (int)(sizeof nml_%s / sizeof nml_%s[0])); %s"
              % (unit_c(), grp, grp, grp, grp, self.tag(no)))
            return
        # internal file (character variable as unit)?
        internal = unit is not None and not (isinstance(unit, Name) and unit.name == "*") \
            and self.typeof(unit).base == "char"
        fmt_c = "0, 0"
        if fmt is not None and not (isinstance(fmt, Name) and fmt.name == "*"):
            fmt_c = self.expr(fmt, None)[0]
        e("{ %s" % self.tag(no))
        self.ind += 1
        if internal:
            e("f_io_begin_internal(%d, %s, %s);" % (1 if h == "write" else 0, self.expr(unit, None)[0], fmt_c))
        else:
            e("f_io_begin(%d, %s, %s);" % (1 if h == "write" else 0, unit_c(), fmt_c))
        for it in items:
            ty = self.typeof(it)
            rw = "put" if h == "write" else "get"
            if ty.rank == 0:
                txt, _ = self.expr(it, None)
                if ty.base == "char":
                    e("f_io_%s_s(%s);" % (rw, txt))
                elif h == "write":
                    e("f_io_put_%s(%s);" % ({"r8": "d", "r4": "d", "int": "i", "logical": "l"}[ty.base], txt))
                else:
                    e("f_io_get_%s(&%s);" % ({"r8": "d", "int": "i"}[ty.base], txt))
            else:
                n = self.extent(it)
                txt, _ = self.expr(it, "i_")
                if h == "write":
                    e("for (i_ = 0; i_ < %s; ++i_) f_io_put_%s(%s);" % (n, "d" if ty.base in ("r8", "r4") else "i", txt))
                else:
                    e("for (i_ = 0; i_ < %s; ++i_) f_io_get_%s(&%s);" % (n, "d" if ty.base == "r8" else "i", txt))
        e("f_io_end();")
        self.ind -= 1
        e("}")

    # ------------------------------------------------------------------ expressions
    def is_designator(self, node):
        if isinstance(node, Name):
            return self.lookup(node.name) is not None
        if isinstance(node, Comp):
            return True
        if isinstance(node, Apply):
            return self.sym_of(node) is not None and not any(isinstance(a, Range) for a in node.args)
        return False

    def sym_of(self, node):
        """the data Sym a designator ends in (component Sym for a%b), or None for calls/exprs"""
        if isinstance(node, Name):
            return self.lookup(node.name)
        if isinstance(node, Comp):
            bty = self.typeof(node.base)
            if bty.base != "derived":
                raise F90Error("%% on non-derived in %s" % self.cur_src)
            comps = self.find_type(bty.derived)
            return comps[node.name]
        if isinstance(node, Apply):
            if isinstance(node.base, Name) and self.lookup(node.base.name) is None:
                return None
            return self.sym_of(node.base)
        return None

    def typeof(self, node):
        if isinstance(node, Num):
            return Ty(self.num_kind(node.text))
        if isinstance(node, Str):
            return Ty("char", 0, len(node.s))
        if isinstance(node, Logical):
            return Ty("logical")
        if isinstance(node, Paren):
            return self.typeof(node.x)
        if isinstance(node, Un):
            return self.typeof(node.x)
        if isinstance(node, Name):
            s = self.lookup(node.name)
            if s is None:
                raise F90Error("%s: unknown name %s in %s" % (self.cur_src, node.name, self.cur_proc.name))
            return s.ty
        if isinstance(node, Comp):
            return self.sym_of(node).ty
        if isinstance(node, Apply):
            s = self.sym_of(node)
            if s is not None:           # array element or section
                rank = sum(1 for a in node.args if isinstance(a, Range))
                return Ty(s.ty.base, rank, s.ty.charlen, s.ty.derived)
            name = node.base.name
            argt = [self.typeof(a) for a in node.args]
            rank = max([a.rank for a in argt] or [0])
            if name in INTRINSIC_R8:
                return Ty("r8" if argt[0].base != "r4" else "r4", rank)
            if name in ("abs",):
                return Ty(argt[0].base, rank)
            if name in ("dmax1", "dmin1", "dble"):
                return Ty("r8", rank)
            if name in ("max", "min"):
                return Ty(self.promote([a.base for a in argt]), rank)
            if name in ("mod",):
                return Ty(self.promote([a.base for a in argt]), rank)
            if name == "sum":
                return Ty(argt[0].base, 0)
            if name in ("size", "len_trim", "len", "int", "nint"):
                return Ty("int")
            if name == "present":
                return Ty("logical")
            if name == "trim":
                return Ty("char", 0, "*")
            if name == "epsilon":
                return Ty(argt[0].base)
            if name == "real":
                return Ty("r4", rank)
            callee = self.find_proc(name)
            if callee is not None and callee.kind == "function":
                self.prepare_proc(callee)
                r = callee.syms[callee.result].ty
                return Ty(r.base, max(r.rank, rank if callee.elemental else 0))
            raise F90Error("%s: unknown function %s" % (self.cur_src, name))
        if isinstance(node, Bin):
            lt, rt = self.typeof(node.l), self.typeof(node.r)
            rank = max(lt.rank, rt.rank)
            if node.op in ("||", "&&", "==", "!=", "<", "<=", ">", ">="):
                return Ty("logical", rank)
            if node.op == "//":
                return Ty("char", 0, "*")
            if node.op == "**":
                return Ty(self.promote([lt.base, rt.base]) if rt.base != "int" else lt.base, rank)
            return Ty(self.promote([lt.base, rt.base]), rank)
        if isinstance(node, Range):
            return Ty("int", 1)
        raise F90Error("typeof %r" % node)

    @staticmethod
    def promote(bases):
        for b in ("r8", "r4", "int", "logical"):
            if b in bases:
                return b
        return bases[0]

    def num_kind(self, text):
        m = re.match(r"^(.*?)(?:_(\w+))?$", text)
        body, kind = m.group(1), m.group(2)
        if kind:
            k = self.kinds.get(kind.lower())
            if k:
                return k
            return "r8" if kind == "8" else "r4"
        if re.search(r"[dD]", body):
            return "r8"
        if re.search(r"[.eE]", body):
            return "r4"
        return "int"

    def num_c(self, text):
        m = re.match(r"^(.*?)(?:_(\w+))?$", text)
        body = m.group(1)
        k = self.num_kind(text)
        if k == "int":
            return body
        body = re.sub(r"[dD]", "e", body)
        if not re.search(r"[.e]", body):
            body += ".0"
        if body.endswith("."):
            body += "0"
        body = re.sub(r"\.e", ".0e", body)
        return body + ("f" if k == "r4" else "")

    def extent(self, node):
        """C text of the extent of the first array-valued operand of an array expression"""
        if isinstance(node, (Paren, Un)):
            return self.extent(node.x)
        if isinstance(node, Bin):
            return self.extent(node.l if self.typeof(node.l).rank else node.r)
        if isinstance(node, (Name, Comp)):
            s = self.sym_of(node)
            if s.ty.rank == 0:
                raise F90Error("extent of scalar %s" % s.name)
            return self.dim_extent(node, s, 0) if s.ty.rank == 1 else \
                "(%s * %s)" % (self.dim_extent(node, s, 0), self.dim_extent(node, s, 1))
        if isinstance(node, Apply):
            s = self.sym_of(node)
            if s is None:
                for a in node.args:
                    if self.typeof(a).rank:
                        return self.extent(a)
                raise F90Error("extent of call")
            for d, a in enumerate(node.args):
                if isinstance(a, Range):
                    lo = self.expr(a.lo, None)[0] if a.lo is not None else "1"
                    hi = self.expr(a.hi, None)[0] if a.hi is not None else self.dim_extent(node.base, s, d)
                    return "((long)(%s) - (%s) + 1)" % (hi, lo)
        raise F90Error("extent of %r" % node)

    def dim_extent(self, desig, s, d):
        if s.ty.base == "char" and getattr(s, "is_dummy", False):
            return "%s_n" % s.cname
        if s.storage == "desc":
            txt, _ = self.designator(desig)
            return "%s.n%d" % (txt, d)
        if s.storage == "fixed" and s.ty.base == "char":
            return s.dims[d].text
        dim = s.dims[d]
        if isinstance(dim, Range):
            raise F90Error("assumed-shape extent of %s" % s.name)
        return "(long)(%s)" % self.expr(dim, None)[0]

    def designator(self, node):
        """C lvalue text of a whole object (no subscripts) + its type"""
        if isinstance(node, Name):
            s = self.lookup(node.name)
            if s is None:
                raise F90Error("unknown name %s" % node.name)
            if s.is_ptr:
                return "(*%s)" % s.cname, s.ty
            return s.cname, s.ty
        if isinstance(node, Comp):
            btxt, bty = self.designator(node.base)
            comps = self.find_type(bty.derived)
            c = comps[node.name]
            if btxt.startswith("(*") and btxt.endswith(")") and btxt[2:-1].isidentifier():
                return "%s->%s" % (btxt[2:-1], c.cname), c.ty
            return "%s.%s" % (btxt, c.cname), c.ty
        if isinstance(node, Apply) and self.sym_of(node) is not None and self.typeof(node).base == "derived":
            raise F90Error("arrays of derived type not supported")
        if isinstance(node, Apply) and all(isinstance(a, Range) and a.lo is None and a.hi is None for a in node.args):
            return self.designator(node.base)
        raise F90Error("not a designator: %r" % node)

    def element(self, node, s, subs):
        """C lvalue of element `subs` (0-based C index texts) of array designator `node`"""
        txt, _ = self.designator(node)
        if s.ty.base == "char":
            ln = int(s.ty.charlen)
            if getattr(s, "is_dummy", False):
                return "(%s + (%s) * %d), %dL" % (txt, subs[0], ln, ln)
            return "%s[%s], %dL" % (txt, subs[0], ln)
        if s.storage == "desc":
            if len(subs) == 1:
                return "%s.a[%s]" % (txt, subs[0])
            return "%s.a[(%s) + (%s) * %s.n0]" % (txt, subs[0], subs[1], txt)
        if len(subs) == 1:
            return "%s[%s]" % (txt, subs[0])
        return "%s[(%s) + (%s) * %s]" % (txt, subs[0], subs[1], self.dim_extent(node, s, 0))

    def expr(self, node, ix):
        """-> (C text, Ty).  ix: None in scalar context, else the C loop index (0-based) used for
        every array-valued operand.  Character values are `ptr, len` pairs."""
        if isinstance(node, Num):
            return self.num_c(node.text), Ty(self.num_kind(node.text))
        if isinstance(node, Str):
            esc = node.s.replace("\\", "\\\\").replace('"', '\\"')
            return '"%s", %dL' % (esc, len(node.s)), Ty("char", 0, len(node.s))
        if isinstance(node, Logical):
            return str(node.v), Ty("logical")
        if isinstance(node, Paren):
            t, ty = self.expr(node.x, ix)
            if ty.base == "char":
                return t, ty
            return "(%s)" % t, ty
        if isinstance(node, Un):
            t, ty = self.expr(node.x, ix)
            return "(%s(%s))" % (node.op, t), (Ty("logical", ty.rank) if node.op == "!" else ty)
        if isinstance(node, (Name, Comp)):
            s = self.sym_of(node)
            if s is None:
                raise F90Error("%s: unknown name %s (in %s)" % (self.cur_src, getattr(node, "name", "?"), self.cur_proc.name))
            if isinstance(node, Name) and node.name in self.kinds:
                raise F90Error("kind parameter used as value")
            if s.ty.rank == 0:
                txt, ty = self.designator(node)
                if ty.base == "char":
                    if isinstance(node, Name) and getattr(s, "is_dummy", False):
                        return "%s, %s_len" % (s.cname, s.cname), ty
                    return "%s, %dL" % (txt, int(ty.charlen)), ty
                return txt, ty
            if ix is None:
                raise F90Error("%s: whole array %s in scalar context (%s)" % (self.cur_src, s.name, self.cur_proc.name))
            if s.ty.rank != 1:
                raise F90Error("whole rank-2 array in elementwise context: %s" % s.name)
            return self.element(node, s, [ix]), s.ty.scalar()
        if isinstance(node, Apply):
            s = self.sym_of(node)
            if s is not None:
                subs, used = [], False
                for d, a in enumerate(node.args):
                    if isinstance(a, Range):
                        if ix is None or used:
                            raise F90Error("%s: array section in scalar context / rank-2 section" % self.cur_src)
                        lo = self.expr(a.lo, None)[0] if a.lo is not None else "1"
                        subs.append("(%s) - 1 + %s" % (lo, ix))
                        used = True
                    else:
                        subs.append("(%s) - 1" % self.expr(a, None)[0])
                return self.element(node.base, s, subs), s.ty.scalar()
            return self.call_expr(node, ix)
        if isinstance(node, Bin):
            lt, lty = self.expr(node.l, ix)
            rt, rty = self.expr(node.r, ix)
            if node.op in ("==", "!=") and lty.base == "char":
                return "%sf_str_eq(%s, %s)" % ("!" if node.op == "!=" else "", lt, rt), Ty("logical")
            if node.op == "**":
                return self.power(node, lt, lty, rt, rty)
            if node.op == "//":
                raise F90Error("string concatenation not supported")
            rank = max(lty.rank, rty.rank)
            if node.op in ("||", "&&", "==", "!=", "<", "<=", ">", ">="):
                return "(%s %s %s)" % (lt, node.op, rt), Ty("logical", rank)
            return "(%s %s %s)" % (lt, node.op, rt), Ty(self.promote([lty.base, rty.base]), rank)
        raise F90Error("expr %r" % node)

    def power(self, node, lt, lty, rt, rty):
        r = node.r
        while isinstance(r, Paren):
            r = r.x
        if rty.base == "int":
            if isinstance(r, Num):
                n = int(r.text)
                if 1 <= n <= 4:
                    # integer constant exponent: gfortran expands to multiplications (x*x, (x*x)*x ...)
                    v = self.newtmp("p")
                    ct = lty.ctype(self.prefix)
                    mul = " * ".join([v] * n) if n != 4 else "(%s * %s) * (%s * %s)" % (v, v, v, v)
                    return "__extension__({ const %s %s = %s; %s; })" % (ct, v, lt, mul), lty
            return "f_powi(%s, %s)" % (lt, rt), Ty("r8")
        return "pow(%s, %s)" % (lt, rt), Ty("r8")

    def call_expr(self, node, ix):
        name = node.base.name
        args = node.args
        if name in INTRINSIC_R8:
            t, ty = self.expr(args[0], ix)
            f = INTRINSIC_R8[name]
            if ty.base == "r4":
                return "%sf(%s)" % (f, t), ty
            return "%s(%s)" % (f, t), Ty("r8")
        if name == "abs":
            t, ty = self.expr(args[0], ix)
            return ("abs(%s)" if ty.base == "int" else "fabs(%s)") % t, ty
        if name in ("dmax1", "dmin1", "max", "min"):
            parts = [self.expr(a, ix) for a in args]
            base = self.promote([p[1].base for p in parts])
            fn = ("f_max" if "max" in name else "f_min") + ("_i" if base == "int" else "")
            txt = parts[0][0]
            for p in parts[1:]:
                txt = "%s(%s, %s)" % (fn, txt, p[0])
            return txt, Ty(base)
        if name == "dble":
            t, _ = self.expr(args[0], ix)
            return "((double)(%s))" % t, Ty("r8")
        if name == "real":
            t, _ = self.expr(args[0], ix)
            return "((float)(%s))" % t, Ty("r4")
        if name == "int":
            t, _ = self.expr(args[0], ix)
            return "((int)(%s))" % t, Ty("int")
        if name == "mod":
            a, aty = self.expr(args[0], ix)
            b, bty = self.expr(args[1], ix)
            if aty.base == "int" and bty.base == "int":
                return "((%s) %% (%s))" % (a, b), Ty("int")
            return "fmod(%s, %s)" % (a, b), Ty("r8")
        if name == "epsilon":
            ty = self.typeof(args[0])
            return ("DBL_EPSILON" if ty.base == "r8" else "FLT_EPSILON"), ty
        if name == "trim":
            t, ty = self.expr(args[0], ix)
            ptr = self.split_charpair(t)[0]
            return "%s, f_len_trim(%s)" % (ptr, t), Ty("char", 0, "*")
        if name == "len_trim":
            t, ty = self.expr(args[0], ix)
            return "((int)f_len_trim(%s))" % t, Ty("int")
        if name == "size":
            return "((int)%s)" % self.extent(args[0]), Ty("int")
        if name == "present":
            s = self.lookup(args[0].name)
            return "(%s != 0)" % s.cname, Ty("logical")
        if name == "sum":
            if ix is not None and self.typeof(args[0]).rank == 0:
                raise F90Error("sum of scalar")
            n = self.extent(args[0])
            v, j = self.newtmp("s"), self.newtmp("j")
            t, ty = self.expr(args[0], j)
            ct = ty.ctype(self.prefix)
            stmt = "%s %s = 0; { long %s; for (%s = 0; %s < %s; ++%s) %s = %s + %s; }" % (ct, v, j, j, j, n, j, v, v, t)
            if self.pre is None:
                raise F90Error("%s: sum() outside a scalar assignment" % self.cur_src)
            self.pre.append(stmt)
            return v, ty
        callee = self.find_proc(name)
        if callee is None or callee.kind != "function":
            raise F90Error("%s: unknown function %s" % (self.cur_src, name))
        self.prepare_proc(callee)
        actual = self.actuals(callee, args, ix)
        r = callee.syms[callee.result].ty
        return "%s%s(%s)" % (self.prefix, name, ", ".join(actual)), r.scalar()

    @staticmethod
    def split_charpair(t):
        depth = 0
        for k in range(len(t) - 1, -1, -1):
            ch = t[k]
            if ch == ")":
                depth += 1
            elif ch == "(":
                depth -= 1
            elif ch == "," and depth == 0:
                return t[:k].strip(), t[k + 1:].strip()
        raise F90Error("not a char pair: %s" % t)

    pre = None


def main(argv):
    defines, prefix, skip, out, files = {}, "ref_", [], None, []
    i = 1
    while i < len(argv):
        a = argv[i]
        if a == "-D":
            defines[argv[i + 1]] = 1
            i += 2
        elif a.startswith("-D"):
            defines[a[2:]] = 1
            i += 1
        elif a == "--prefix":
            prefix = argv[i + 1]
            i += 2
        elif a == "--skip":
            skip = [x for x in argv[i + 1].split(",") if x]
            i += 2
        elif a == "-o":
            out = argv[i + 1]
            i += 2
        else:
            files.append(a)
            i += 1
    if not files or out is None:
        sys.stderr.write(__doc__)
        return 2
    fes = [FrontEnd(f, defines) for f in files]
    text = Gen(fes, prefix, skip).generate()
    with open(out, "w") as f:
        f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
