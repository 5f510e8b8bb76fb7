#!/usr/bin/env python3
"""f90toc.py -- mechanical translator of the Fortran 90 subset the reference is written in into C++.

TEST INFRASTRUCTURE ONLY (oracle/).  No Fortran compiler exists in this image, so the reference
cannot be built.  This script is the next best pin for the hand-written oracle: it reads the reference's
own source files where they lie (/root/reference/*.f90, never copied into the repo), translates them
statement by statement into C++ with the same operation order, and writes the result into
oracle/_ref/ (git-ignored).  oracle/Makefile compiles that into oracle/_ref/libpigs_ref.so with
g++ -O2 -ffp-contract=off -fwrapv, and tests/test_ref_pin.py demands BIT EQUALITY between it and
oracle/pigs_oracle.cpp on the same inputs and the same MT19937 stream.

What is covered: free-form source with '&' continuations; modules, module variables, allocatable and
explicit-shape arrays with arbitrary lower bounds (column-major), automatic arrays, optional dummies
(present()), subroutines/functions (by-reference semantics for every dummy the callee or its callees
may assign, by value otherwise), do / do-forever / exit / cycle, block and one-line if, statement
functions, COMMON/DATA/SAVE/PARAMETER/IMPLICIT, complex(kind=8), the intrinsics the reference uses,
default-kind real() (single precision -- the reference's float32 ratios survive), integer powers
expanded exactly as gfortran/GCC expand them (the powi table), real**real through pow().
I/O: print / namelist / rewind / open / close are dropped; `write(unit,...)` of numbers is captured
through f90rt::write_rec (so e_vpi.out etc. can be compared as numbers); `read` goes to f90rt::read_rec,
served by the glue.  Anything outside the subset raises an error instead of guessing.
"""
from __future__ import annotations

import re
import sys
from dataclasses import dataclass, field

# --------------------------------------------------------------------------------------- source reader


def logical_lines(path):
    """[(lineno, text)] with comments stripped, continuations joined, code lower-cased outside strings"""
    out = []
    buf, start = "", None
    for no, raw in enumerate(open(path, encoding="latin-1"), 1):
        line = raw.rstrip("\n")
        # strip comments (outside strings) and lower-case code
        res, q = [], None
        for ch in line:
            if q:
                res.append(ch)
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
                res.append(ch)
                continue
            if ch == "!":
                break
            res.append(ch.lower())
        text = "".join(res).strip()
        if not text:
            continue
        if buf:
            if text.startswith("&"):
                text = text[1:].lstrip()
            buf += text
        else:
            buf, start = text, no
        if buf.endswith("&"):
            buf = buf[:-1].rstrip() + " "
            continue
        out.append((start, buf))
        buf = ""
    return out


# --------------------------------------------------------------------------------------- expressions
TOK = re.compile(r"""
    (?P<real>(\d+\.\d*|\.\d+|\d+)([ed][+-]?\d+)|\d+\.\d*|\.\d+(?![a-z]))
  | (?P<int>\d+)
  | (?P<dot>\.(and|or|not|eqv|neqv|true|false|eq|ne|lt|le|gt|ge)\.)
  | (?P<name>[a-z_][a-z0-9_]*)
  | (?P<str>'[^']*'|"[^"]*")
  | (?P<op>\*\*|==|/=|<=|>=|//|[-+*/(),:<>=%])
""", re.X)

DOTMAP = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def tokenize(s):
    toks, i = [], 0
    s = s.strip()
    while i < len(s):
        if s[i].isspace():
            i += 1
            continue
        m = TOK.match(s, i)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[i:]!r} in {s!r}")
        kind = m.lastgroup
        text = m.group(kind)
        if kind == "dot":
            if text in DOTMAP:
                kind, text = "op", DOTMAP[text]
            elif text in (".true.", ".false."):
                kind = "logical"
            else:
                kind = "op"
        toks.append((kind, text))
        i = m.end()
    return toks


@dataclass
class Node:
    k: str                 # num, str, logical, name, call, un, bin, colon, implied
    v: object = None
    a: list = field(default_factory=list)


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, text=None):
        k, v = self.peek()
        if text is not None and v != text:
            raise SyntaxError(f"expected {text!r}, got {v!r} in {self.t}")
        self.i += 1
        return k, v

    # precedence climbing
    def expr(self):
        return self.eqv()

    def eqv(self):
        l = self.or_()
        while self.peek()[1] in (".eqv.", ".neqv."):
            op = self.take()[1]
            l = Node("bin", op, [l, self.or_()])
        return l

    def or_(self):
        l = self.and_()
        while self.peek()[1] == ".or.":
            self.take()
            l = Node("bin", ".or.", [l, self.and_()])
        return l

    def and_(self):
        l = self.not_()
        while self.peek()[1] == ".and.":
            self.take()
            l = Node("bin", ".and.", [l, self.not_()])
        return l

    def not_(self):
        if self.peek()[1] == ".not.":
            self.take()
            return Node("un", ".not.", [self.not_()])
        return self.rel()

    def rel(self):
        l = self.add()
        if self.peek()[1] in ("==", "/=", "<", "<=", ">", ">="):
            op = self.take()[1]
            l = Node("bin", op, [l, self.add()])
        return l

    def add(self):
        if self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            l = Node("un", op, [self.mul()])
        else:
            l = self.mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            l = Node("bin", op, [l, self.mul()])
        return l

    def mul(self):
        l = self.pow_()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            l = Node("bin", op, [l, self.pow_()])
        return l

    def pow_(self):
        b = self.primary()
        if self.peek()[1] == "**":
            self.take()
            # right associative; the exponent may carry a sign
            if self.peek()[1] in ("+", "-"):
                op = self.take()[1]
                e = Node("un", op, [self.pow_()])
            else:
                e = self.pow_()
            return Node("bin", "**", [b, e])
        return b

    def primary(self):
        k, v = self.take()
        if k == "real":
            return Node("num", v)
        if k == "int":
            return Node("num", v)
        if k == "str":
            return Node("str", v[1:-1])
        if k == "logical":
            return Node("logical", v == ".true.")
        if v == "(":
            e = self.expr()
            if self.peek()[1] == ",":          # implied do in an I/O list: (items, var=a,b)
                items = [e]
                while self.peek()[1] == ",":
                    self.take()
                    items.append(self.arg())
                self.take(")")
                return Node("implied", None, items)
            self.take(")")
            return Node("paren", None, [e])
        if k == "name":
            if self.peek()[1] == "(":
                self.take()
                args = []
                if self.peek()[1] != ")":
                    args.append(self.arg())
                    while self.peek()[1] == ",":
                        self.take()
                        args.append(self.arg())
                self.take(")")
                return Node("call", v, args)
            return Node("name", v)
        raise SyntaxError(f"unexpected token {v!r} in {self.t}")

    def arg(self):
        # a subscript, a section bound pair, a keyword argument or an implied-do control (k=1,dim)
        if self.peek()[1] == ":":
            self.take()
            if self.peek()[1] in (",", ")"):
                return Node("colon", None, [None, None])
            return Node("colon", None, [None, self.expr()])
        e = self.expr()
        if self.peek()[1] == ":":
            self.take()
            hi = None if self.peek()[1] in (",", ")") else self.expr()
            return Node("colon", None, [e, hi])
        if self.peek()[1] == "=" and e.k == "name":
            self.take()
            return Node("kw", e.v, [self.expr()])
        return e


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {s!r}: {p.t[p.i:]}")
    return e


def split_top(s, sep=","):
    """split at top-level separators (outside parentheses and strings)"""
    parts, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
            continue
        cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def matching_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError(f"unbalanced parentheses in {s!r}")


# --------------------------------------------------------------------------------------- symbols
CTYPE = {"int": "int", "real4": "float", "real8": "double", "logical": "bool", "complex8": "std::complex<double>",
         "char": "std::string"}


@dataclass
class Sym:
    name: str
    ty: str
    dims: list | None = None        # list of (lo_expr_str, hi_expr_str) or None; deferred shape: [(None, None), ...]
    allocatable: bool = False
    optional: bool = False
    dummy: bool = False
    save: bool = False
    init: str | None = None
    parameter: bool = False
    scope: str = "local"            # local | module | common
    is_func_result: bool = False
    charlen: int = 0


@dataclass
class Proc:
    name: str
    kind: str                       # subroutine | function | program
    args: list
    module: str | None
    body: list = field(default_factory=list)      # [(lineno, text)]
    syms: dict = field(default_factory=dict)
    result_ty: str | None = None
    implicit_int: bool = False
    stmt_funcs: dict = field(default_factory=dict)
    modified: set = field(default_factory=set)    # dummies that may be assigned (directly or through callees)
    file: str = ""
    line: int = 0


TYPE_RE = re.compile(r"^(real|integer|logical|character|complex|double\s+precision)\b")


def base_type(spec):
    spec = spec.replace(" ", "")
    if spec.startswith("doubleprecision"):
        return "real8", 0
    if spec.startswith("real"):
        return ("real8" if ("kind=8" in spec or "(8)" in spec) else "real4"), 0
    if spec.startswith("integer"):
        return "int", 0
    if spec.startswith("logical"):
        return "logical", 0
    if spec.startswith("complex"):
        return "complex8", 0
    if spec.startswith("character"):
        m = re.search(r"len=(\d+)", spec)
        return "char", int(m.group(1)) if m else 1
    raise SyntaxError(spec)


def parse_dims(s):
    """'dim,Np,0:2*Nb' -> [(lo,hi),...]; ':' -> (None,None)"""
    dims = []
    for part in split_top(s):
        if part == ":":
            dims.append((None, None))
        elif ":" in split_top(part, ":")[0:0] or len(split_top(part, ":")) == 2:
            lo, hi = split_top(part, ":")
            dims.append((lo, hi))
        else:
            dims.append(("1", part))
    return dims


class Translator:
    def __init__(self, skip=(), count=()):
        self.count = set(count)          # procedures whose calls are counted (f90rt-free global `calls_<name>_`)
        self.modvars: dict[str, Sym] = {}
        self.procs: dict[str, Proc] = {}
        self.order: list[str] = []
        self.skip = set(skip)
        self.out: list[str] = []

    # ---------------------------------------------------------------- pass 1: structure
    def load(self, path):
        lines = logical_lines(path)
        i, module = 0, None
        fname = path.split("/")[-1]
        in_contains = False
        while i < len(lines):
            no, t = lines[i]
            m = re.match(r"^module\s+(\w+)$", t)
            if m:
                module, in_contains = m.group(1), False
                i += 1
                continue
            if re.match(r"^end\s*module", t):
                module = None
                i += 1
                continue
            if t == "contains":
                in_contains = True
                i += 1
                continue
            m = re.match(r"^(?:(double\s+precision|real\s*\(kind=8\)|integer(?:\s*\(kind=4\))?|logical)\s+)?(subroutine|function|program)\s+(\w+)\s*(\((.*)\))?\s*$", t)
            if m:
                pre, kind, name, _, args = m.groups()
                args = [a.strip() for a in args.split(",")] if args and args.strip() else []
                p = Proc(name, kind, args, module, file=fname, line=no)
                if pre:
                    p.result_ty = base_type(pre)[0]
                i += 1
                while not re.match(rf"^end\s*({kind})?(\s+{name})?\s*$", lines[i][1]):
                    p.body.append(lines[i])
                    i += 1
                i += 1
                self.procs[name] = p
                self.order.append(name)
                continue
            if module and not in_contains:
                self.declaration(t, self.modvars, scope="module", where=f"{fname}:{no}")
                i += 1
                continue
            raise SyntaxError(f"{fname}:{no}: unexpected top-level statement {t!r}")

    # ---------------------------------------------------------------- declarations
    def declaration(self, t, table, scope, where, proc=None):
        """returns True if t was a declaration-like statement"""
        if t.startswith("use ") or t == "implicit none" or t.startswith("namelist") or t.startswith("external"):
            return True
        if t.startswith("implicit integer"):
            if proc:
                proc.implicit_int = True
            return True
        if t.startswith("save"):
            return True
        m = re.match(r"^dimension\s+(.*)$", t)
        if m:
            for ent in split_top(m.group(1)):
                mm = re.match(r"^(\w+)\((.*)\)$", ent)
                name = mm.group(1)
                s = table.get(name) or Sym(name, "int", scope=scope)
                s.dims = parse_dims(mm.group(2))
                table[name] = s
            return True
        m = re.match(r"^common\s*/(\w+)/\s*(.*)$", t)
        if m:
            for name in split_top(m.group(2)):
                s = table.pop(name, None) or Sym(name, "int")
                s.scope = "common"
                if name in self.modvars:
                    # a later procedure re-declares the block: keep the first definition (and its DATA)
                    if s.init and not self.modvars[name].init:
                        self.modvars[name].init = s.init
                else:
                    self.modvars[name] = s
                if proc:
                    proc.syms.pop(name, None)
                    proc.commons = getattr(proc, "commons", set()) | {name}
            return True
        m = re.match(r"^data\s+(\w+)\s*/(.*)/$", t)
        if m:
            name, vals = m.group(1), m.group(2)
            s = table.get(name) or self.modvars.get(name)
            if s is None:
                s = Sym(name, "int", scope=scope)
                table[name] = s
            s.init = vals
            s.save = True
            return True
        if TYPE_RE.match(t):
            mt = re.match(r"^(real|integer|logical|complex|character|double\s+precision)\s*(\([^)]*\))?", t)
            spec = re.sub(r"\s+", "", mt.group(0))
            rest = t[mt.end():].strip()
            if re.match(r"^function\b", rest):
                return False
            if "::" in rest:
                left, right = rest.split("::", 1)
                attrs = [a.replace(" ", "") for a in split_top(left) if a.strip()]
            else:
                attrs, right = [], rest
            ty, clen = base_type(spec)
            dims = None
            for a in attrs:
                if a.startswith("dimension"):
                    dims = parse_dims(a[a.index("(") + 1:-1])
            for ent in split_top(right):
                init = None
                if "=" in ent and not re.search(r"\([^)]*=[^)]*\)", ent):
                    ent, init = [x.strip() for x in ent.split("=", 1)]
                mm = re.match(r"^(\w+)\s*(\((.*)\))?$", ent)
                name = mm.group(1)
                edims = parse_dims(mm.group(3)) if mm.group(3) else dims
                s = table.get(name)
                if s is None:
                    s = Sym(name, ty, scope=scope)
                    table[name] = s
                s.ty, s.charlen = ty, clen
                if edims:
                    s.dims = edims
                s.allocatable = s.allocatable or "allocatable" in attrs
                s.optional = s.optional or "optional" in attrs
                s.save = s.save or "save" in attrs or init is not None
                s.parameter = s.parameter or "parameter" in attrs
                if init is not None:
                    init = init.strip()
                    if init.startswith("(/"):
                        init = init[2:-2]
                    s.init = init
            return True
        return False

    # ---------------------------------------------------------------- pass 2: per-procedure symbol tables
    def analyse(self):
        for p in self.procs.values():
            for a in p.args:
                p.syms[a] = Sym(a, "int" if p.implicit_int else "?", dummy=True)
            rest = []
            for no, t in p.body:
                if re.match(r"^\d+\s+format", t):
                    continue
                t2 = re.sub(r"^\d+\s+", "", t)      # statement labels
                # statement function:  name(args) = expr   with name undeclared as array
                if self.declaration(t2, p.syms, "local", f"{p.file}:{no}", proc=p):
                    continue
                m = re.match(r"^(\w+)\(([\w,\s]*)\)\s*=\s*(.*)$", t2)
                if m and p.implicit_int and m.group(1) not in p.syms and m.group(1) not in self.modvars and not rest:
                    p.stmt_funcs[m.group(1)] = ([a.strip() for a in m.group(2).split(",")], m.group(3))
                    continue
                rest.append((no, t2))
            p.body = rest
            for a in p.args:
                p.syms[a].dummy = True
                if p.syms[a].ty == "?" and p.implicit_int:
                    p.syms[a].ty = "int"
                if p.syms[a].ty == "?":
                    raise SyntaxError(f"{p.file}:{p.line}: dummy {a} of {p.name} has no type")
            if p.kind == "function":
                s = p.syms.get(p.name)
                if s is None:
                    s = Sym(p.name, p.result_ty or ("int" if p.implicit_int else "?"))
                    p.syms[p.name] = s
                if p.result_ty is None:
                    p.result_ty = s.ty
                s.is_func_result = True
        # names declared like variables but naming external functions (real(kind=8) :: Interpolate)
        for p in self.procs.values():
            for n in list(p.syms):
                if n in self.procs and n != p.name and not p.syms[n].dummy:
                    del p.syms[n]
        # which dummies may be modified: fixed point over the call graph
        for p in self.procs.values():
            p.calls = []
            for no, t in p.body:
                self.scan_modified(p, t)
        for p in self.procs.values():
            if p.name in self.skip:          # supplied by the glue, which may write every scalar dummy
                p.modified = {a for a in p.args if p.syms[a].dims is None and (p.syms[a].ty != "char" or p.name == "readparameters")}
        changed = True
        while changed:
            changed = False
            for p in self.procs.values():
                for callee, actuals in p.calls:
                    q = self.procs.get(callee)
                    if q is None:
                        continue
                    for pos, act in enumerate(actuals):
                        if pos < len(q.args) and q.args[pos] in q.modified and act in p.syms and p.syms[act].dummy and act not in p.modified:
                            p.modified.add(act)
                            changed = True

    def scan_modified(self, p, t):
        """record direct assignments to dummies and the calls (for the fixed point)"""
        m = re.match(r"^if\s*\(", t)
        if m:
            j = matching_paren(t, t.index("("))
            rest = t[j + 1:].strip()
            self.scan_calls(p, t[t.index("("):j + 1])
            if rest and rest != "then":
                self.scan_modified(p, rest)
            return
        m = re.match(r"^else\s*if\s*\(", t)
        if m:
            self.scan_calls(p, t)
            return
        m = re.match(r"^do\s+(\w+)\s*=", t)
        if m:
            if m.group(1) in p.syms and p.syms[m.group(1)].dummy:
                p.modified.add(m.group(1))
            self.scan_calls(p, t.split("=", 1)[1])
            return
        m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", t)
        if m:
            args = split_top(m.group(3)) if m.group(3) else []
            names = []
            for a in args:
                mm = re.match(r"^(\w+)", a)
                names.append(mm.group(1) if mm and re.match(r"^\w+(\(.*\))?$", a) else None)
            p.calls.append((m.group(1), names))
            for a in args:
                self.scan_calls(p, a)
            return
        m = re.match(r"^read\s*\(", t)
        if m:
            j = matching_paren(t, t.index("("))
            for a in split_top(t[j + 1:].strip()):
                mm = re.match(r"^\(?(\w+)", a)
                if mm and mm.group(1) in p.syms and p.syms[mm.group(1)].dummy:
                    p.modified.add(mm.group(1))
            return
        # assignment
        eq = self.find_assign(t)
        if eq is not None:
            lhs = t[:eq].strip()
            mm = re.match(r"^(\w+)", lhs)
            if mm and mm.group(1) in p.syms and p.syms[mm.group(1)].dummy:
                p.modified.add(mm.group(1))
            self.scan_calls(p, t[eq + 1:])

    def scan_calls(self, p, text):
        """function references inside an expression: their actual arguments may be modified too"""
        for m in re.finditer(r"\b(\w+)\s*\(", text):
            name = m.group(1)
            if name in self.procs and self.procs[name].kind == "function":
                j = matching_paren(text, m.end() - 1)
                args = split_top(text[m.end():j])
                names = []
                for a in args:
                    mm = re.match(r"^(\w+)(\(.*\))?$", a)
                    names.append(mm.group(1) if mm else None)
                p.calls.append((name, names))

    @staticmethod
    def find_assign(t):
        depth, q = 0, None
        for i, ch in enumerate(t):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0:
                if t[i + 1:i + 2] == "=" or t[i - 1] in "<>/=":
                    continue
                return i
        return None

    # ---------------------------------------------------------------- types of expressions
    def lookup(self, p, name):
        if p and name in p.syms:
            return p.syms[name]
        return self.modvars.get(name)

    INTRINSIC_TY = {"sqrt": "arg", "exp": "arg", "log": "arg", "cos": "arg", "sin": "arg", "acos": "arg", "abs": "arg",
                    "int": "int", "dble": "real8", "mod": "arg", "min": "arg", "max": "arg", "iand": "int", "ior": "int",
                    "ieor": "int", "ishft": "int", "aimag": "real8", "minval": "real8", "present": "logical", "sign": "arg",
                    "nint": "int"}

    def typeof(self, p, e):
        k = e.k
        if k == "num":
            v = e.v
            if re.fullmatch(r"\d+", v):
                return "int"
            return "real8" if "d" in v else "real4"
        if k == "logical":
            return "logical"
        if k == "str":
            return "char"
        if k == "paren":
            return self.typeof(p, e.a[0])
        if k == "name":
            s = self.lookup(p, e.v)
            if s is None:
                if p and p.implicit_int:
                    return "int"
                raise SyntaxError(f"{p.name if p else ''}: unknown name {e.v}")
            return s.ty
        if k == "call":
            s = self.lookup(p, e.v)
            if s is not None and s.dims is not None:
                return s.ty
            if p and e.v in p.stmt_funcs:
                return "int"
            if e.v in self.procs and (s is None or not s.dummy):
                return self.procs[e.v].result_ty
            if e.v in ("aint", "huge"):
                return "real8"
            if e.v == "real":
                if len(e.a) == 2:
                    return "real8"
                at = self.typeof(p, e.a[0])
                return "real8" if at == "complex8" else "real4"
            if e.v == "cmplx":
                return "complex8"
            if e.v in self.INTRINSIC_TY:
                r = self.INTRINSIC_TY[e.v]
                if r == "arg":
                    tys = [self.typeof(p, a) for a in e.a]
                    return self.promote(tys)
                return r
            if e.v == "r8_gamma":
                return "real8"
            raise SyntaxError(f"{p.name}: unknown function/array {e.v}")
        if k == "un":
            return "logical" if e.v == ".not." else self.typeof(p, e.a[0])
        if k == "bin":
            if e.v in ("==", "/=", "<", "<=", ">", ">=", ".and.", ".or.", ".eqv.", ".neqv."):
                return "logical"
            lt, rt = self.typeof(p, e.a[0]), self.typeof(p, e.a[1])
            if e.v == "**":
                return lt if rt == "int" else self.promote([lt, rt])
            return self.promote([lt, rt])
        raise SyntaxError(f"typeof {e}")

    @staticmethod
    def promote(tys):
        for t in ("complex8", "real8", "real4", "int", "logical", "char"):
            if t in tys:
                return t
        raise SyntaxError(str(tys))

    # ---------------------------------------------------------------- expression emission
    def cname(self, name):
        return name + "_"

    def lit(self, v):
        if re.fullmatch(r"\d+", v):
            return v
        if "d" in v:
            m, x = v.split("d")
            if "." not in m:
                m += ".0"
            if m.endswith("."):
                m += "0"
            if m.startswith("."):
                m = "0" + m
            return f"{m}e{x}" if x not in ("0", "+0", "-0") else m
        m = v
        if "e" in m:
            return m + "f"
        if m.endswith("."):
            m += "0"
        if m.startswith("."):
            m = "0" + m
        return m + "f"

    def emit(self, p, e):
        k = e.k
        if k == "num":
            return self.lit(e.v)
        if k == "logical":
            return "true" if e.v else "false"
        if k == "str":
            return 'std::string("' + e.v.replace('"', '\\"') + '")'
        if k == "paren":
            return "(" + self.emit(p, e.a[0]) + ")"
        if k == "name":
            s = self.lookup(p, e.v)
            if s is None and p and p.implicit_int:
                p.syms[e.v] = Sym(e.v, "int")
                s = p.syms[e.v]
            if s is None:
                raise SyntaxError(f"{p.name}: unknown name {e.v}")
            if s.dims is not None:
                return self.cname(e.v) + "_p"           # whole array -> pointer to its first element
            if s.dummy and s.optional:
                return "(*" + self.cname(e.v) + ")"
            return self.cname(e.v)
        if k == "call":
            return self.emit_call(p, e)
        if k == "un":
            if e.v == ".not.":
                return "(!" + self.emit(p, e.a[0]) + ")"
            return "(" + e.v + self.emit(p, e.a[0]) + ")"
        if k == "bin":
            op = e.v
            l, r = e.a
            if op == "**":
                lt, rt = self.typeof(p, l), self.typeof(p, r)
                if rt == "int":
                    # constant exponent: the multiplication tree gfortran/GCC emit; variable: the library loop
                    rr = r
                    while rr.k == "paren":
                        rr = rr.a[0]
                    if rr.k == "num":
                        if lt == "int":
                            return f"f90rt::ipow_int({self.emit(p, l)}, {rr.v})"
                        return f"f90rt::powi_c<{rr.v}>({self.emit(p, l)})"
                    if lt == "int":
                        return f"f90rt::ipow_int({self.emit(p, l)}, {self.emit(p, r)})"
                    return f"f90rt::powi_v({self.emit(p, l)}, {self.emit(p, r)})"
                return f"std::pow((double)({self.emit(p, l)}), (double)({self.emit(p, r)}))"
            cop = {"/=": "!=", ".and.": "&&", ".or.": "||", ".eqv.": "==", ".neqv.": "!="}.get(op, op)
            if op in ("==", "/=") and self.typeof(p, l) == "char":
                return f"(f90rt::streq({self.emit(p, l)}, {self.emit(p, r)}) {'==' if op == '==' else '!='} true)"
            return "(" + self.emit(p, l) + " " + cop + " " + self.emit(p, r) + ")"
        raise SyntaxError(f"emit {e}")

    def emit_call(self, p, e):
        name, args = e.v, e.a
        s = self.lookup(p, name)
        if s is not None and s.dims is not None:
            if any(a.k == "colon" for a in args):
                # a contiguous leading section used as an actual argument: pointer to its first element
                idx = []
                for d, a in enumerate(args):
                    if a.k == "colon":
                        if a.a[0] is not None or a.a[1] is not None:
                            raise SyntaxError(f"{p.name}: only full ':' sections are supported ({name})")
                        idx.append(f"{self.cname(name)}_l{d + 1}")
                    else:
                        idx.append(self.emit(p, a))
                return f"(&{self.cname(name)}({', '.join(idx)}))"
            return f"{self.cname(name)}({', '.join(self.emit(p, a) for a in args)})"
        if p and name in p.stmt_funcs:
            return f"{self.cname(name)}sf({', '.join(self.emit(p, a) for a in args)})"
        if name in self.procs and (s is None or not s.dummy):
            return self.emit_proc_call(p, name, args)
        A = [self.emit(p, a) for a in args if a.k != "kw"]
        if name == "aint":
            return f"std::trunc({A[0]})"
        if name == "huge":
            return "1.7976931348623157e308"
        if name == "real":
            if len(args) == 2:
                return f"((double)({A[0]}))"
            if self.typeof(p, args[0]) == "complex8":
                return f"(({A[0]}).real())"
            return f"((float)({A[0]}))"
        if name == "dble":
            return f"((double)({A[0]}))"
        if name == "int":
            return f"((int)({A[0]}))"
        if name == "nint":
            return f"((int)std::lround({A[0]}))"
        if name in ("sqrt", "exp", "log", "cos", "sin", "acos"):
            t = self.typeof(p, args[0])
            if t == "real4":
                return f"std::{name}((float)({A[0]}))"
            return f"std::{name}((double)({A[0]}))"
        if name == "abs":
            return f"std::abs({A[0]})"
        if name == "mod":
            if self.typeof(p, e) == "int":
                return f"(({A[0]}) % ({A[1]}))"
            return f"std::fmod({A[0]}, {A[1]})"
        if name in ("min", "max"):
            t = CTYPE[self.typeof(p, e)]
            r = f"({t})({A[0]})"
            for x in A[1:]:
                r = f"std::{name}<{t}>({r}, ({t})({x}))"
            return r
        if name == "sign":
            return f"f90rt::sign({A[0]}, {A[1]})"
        if name in ("iand", "ior", "ieor"):
            return f"(({A[0]}) {dict(iand='&', ior='|', ieor='^')[name]} ({A[1]}))"
        if name == "ishft":
            return f"f90rt::ishft({A[0]}, {A[1]})"
        if name == "cmplx":
            return f"std::complex<double>({A[0]}, {A[1]})"
        if name == "aimag":
            return f"(({A[0]}).imag())"
        if name == "present":
            return f"({self.cname(args[0].v)} != nullptr)"
        if name == "minval":
            sy = self.lookup(p, args[0].v)
            return f"f90rt::minval({self.cname(args[0].v)}_p, {self.cname(args[0].v)}_n1)"
        if name == "r8_gamma":
            return f"r8_gamma_({A[0]})"
        raise SyntaxError(f"{p.name}: unknown intrinsic {name}")

    def emit_proc_call(self, p, name, args):
        q = self.procs[name]
        outs = []
        for pos, dn in enumerate(q.args):
            ds = q.syms[dn]
            if pos >= len(args):
                if not ds.optional:
                    raise SyntaxError(f"{p.name}: too few arguments in call of {name}")
                outs.append("nullptr")
                continue
            a = args[pos]
            if a.k == "kw":
                raise SyntaxError("keyword arguments are not supported")
            if ds.dims is not None:
                outs.append(self.emit(p, a))              # whole array or section -> pointer
                if a.k == "call" and not any(x.k == "colon" for x in a.a):
                    outs[-1] = "&" + outs[-1]             # array element as the start of a sequence
                continue
            if ds.optional:
                # pass the address of an lvalue (forwarding an optional dummy keeps nullptr)
                src = self.lookup(p, a.v) if a.k == "name" else None
                if src is not None and src.dummy and src.optional:
                    outs.append(self.cname(a.v))
                else:
                    outs.append("&" + self.emit(p, a))
                continue
            if dn in q.modified:
                if a.k not in ("name", "call"):
                    raise SyntaxError(f"{p.name}: expression passed to modified dummy {dn} of {name}")
                outs.append(self.emit(p, a))
            else:
                outs.append(self.emit(p, a))
        return f"{self.cname(name)}({', '.join(outs)})"

    # ---------------------------------------------------------------- statements
    def array_macros(self, p, s, lines, pointer_decl=None):
        """bounds + index macro of array s (dims known at entry)"""
        n = self.cname(s.name)
        nd = len(s.dims)
        if not s.allocatable:
            for d, (lo, hi) in enumerate(s.dims, 1):
                lines.append(f"  const int {n}_l{d} = {self.emit(p, parse_expr(lo))}, {n}_n{d} = ({self.emit(p, parse_expr(hi))}) - {n}_l{d} + 1; (void){n}_n{d};")
        params = ", ".join(f"i{d}" for d in range(1, nd + 1))
        expr = ""
        for d in range(nd, 0, -1):
            term = f"((i{d}) - {n}_l{d})"
            expr = term if not expr else f"({term} + {n}_n{d} * {expr})" if False else expr
        # column-major: idx = (i1-l1) + n1*((i2-l2) + n2*((i3-l3)))
        expr = f"((i{nd}) - {n}_l{nd})"
        for d in range(nd - 1, 0, -1):
            expr = f"(((i{d}) - {n}_l{d}) + {n}_n{d} * {expr})"
        lines.append(f"#define {n}({params}) {n}_p[{expr}]")

    def proto(self, q):
        ps = []
        for a in q.args:
            s = q.syms[a]
            ct = CTYPE[s.ty]
            if s.dims is not None:
                ps.append(f"{ct}* {self.cname(a)}_p")
            elif s.optional:
                ps.append(f"{ct}* {self.cname(a)}")
            elif a in q.modified:
                ps.append(f"{ct}& {self.cname(a)}")
            else:
                ps.append(f"{ct} {self.cname(a)}")
        ret = "void" if q.kind != "function" else CTYPE[q.result_ty]
        return f"{ret} {self.cname(q.name)}({', '.join(ps)})"

    def translate_proc(self, q):
        L = [f"// {q.file}:{q.line}  {q.kind} {q.name}", self.proto(q) + " {"]
        if q.name in self.count:
            L.append(f"  ++calls_{self.cname(q.name)};")
        undef = []
        # statement functions
        for sf, (sargs, sexpr) in q.stmt_funcs.items():
            for a in sargs:
                q.syms.setdefault(a, Sym(a, "int"))
            L.append(f"  auto {self.cname(sf)}sf = [&]({', '.join('int ' + self.cname(a) for a in sargs)}) {{ return {self.emit(q, parse_expr(sexpr))}; }};")
        # implicit-int names appear on first use: pre-scan the body for undeclared names
        if q.implicit_int:
            for no, t in q.body:
                for m in re.finditer(r"\b([a-z_]\w*)\b", re.sub(r"'[^']*'", "", t)):
                    n = m.group(1)
                    if n in q.syms or n in self.modvars or n in self.procs or n in q.stmt_funcs:
                        continue
                    if n in ("if", "then", "else", "endif", "end", "do", "call", "return", "ge", "eq", "lt", "iand", "ior", "ieor", "ishft", "dble", "enddo"):
                        continue
                    q.syms[n] = Sym(n, "int")
        # locals
        for s in q.syms.values():
            if s.dummy:
                if s.dims is not None:
                    self.array_macros(q, s, L)
                    undef.append(self.cname(s.name))
                continue
            if s.scope == "common":
                continue
            ct = CTYPE[s.ty]
            n = self.cname(s.name)
            if s.parameter:
                L.append(f"  const {ct} {n} = {self.emit(q, parse_expr(s.init))};")
                continue
            if s.dims is not None:
                if s.allocatable:
                    nd = len(s.dims)
                    L.append(f"  {ct}* {n}_p = nullptr; " + " ".join(f"int {n}_l{d} = 1, {n}_n{d} = 0; (void){n}_l{d}; (void){n}_n{d};" for d in range(1, nd + 1)))
                    self.array_macros(q, s, L)
                else:
                    tmp = []
                    self.array_macros(q, s, tmp)
                    L.extend(tmp[:-1])
                    size = " * ".join(f"{n}_n{d}" for d in range(1, len(s.dims) + 1))
                    if s.save and s.init:
                        vals = ", ".join(self.emit(q, parse_expr(v)) for v in split_top(s.init))
                        L.append(f"  static {ct} {n}_p[] = {{{vals}}};")
                    else:
                        L.append(f"  {ct} {n}_p[({size}) > 0 ? ({size}) : 1];")
                    L.append(tmp[-1])
                undef.append(n)
                continue
            if s.is_func_result:
                L.append(f"  {ct} {n} = {ct}();")
                continue
            if s.save:
                init = f" = {self.emit(q, parse_expr(s.init))}" if s.init else ""
                L.append(f"  static {ct} {n}{init};")
            else:
                L.append(f"  {ct} {n} = {ct}(); (void){n};")
        # body
        ind = 1
        blocks = []
        for no, t in q.body:
            for line, delta_before, delta_after in self.statement(q, t, no, blocks):
                ind += delta_before
                L.append("  " * ind + line)
                ind += delta_after
        if q.kind == "function":
            L.append(f"  return {self.cname(q.name)};")
        L.append("}")
        for u in undef:
            L.append(f"#undef {u}")
        return L

    def ret_stmt(self, q):
        return f"return {self.cname(q.name)};" if q.kind == "function" else "return;"

    def statement(self, q, t, no, blocks):
        """yields (text, indent_before, indent_after)"""
        tag = f"  // {q.file}:{no}"
        if t in ("return",):
            yield self.ret_stmt(q) + tag, 0, 0
            return
        if t == "stop":
            yield "f90rt::stop();" + tag, 0, 0
            return
        if t == "exit":
            yield "break;" + tag, 0, 0
            return
        if t == "cycle":
            yield "continue;" + tag, 0, 0
            return
        if t in ("else",):
            yield "} else {", -1, 1
            return
        if re.match(r"^end\s*if$", t) or re.match(r"^end\s*do$", t):
            yield "}", -1, 0
            return
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", t)
        if m:
            yield f"}} else if ({self.emit(q, parse_expr(m.group(1)))}) {{" + tag, -1, 1
            return
        if t.startswith("if"):
            m = re.match(r"^if\s*\(", t)
            if m:
                j = matching_paren(t, t.index("("))
                cond, rest = t[t.index("(") + 1:j], t[j + 1:].strip()
                c = self.emit(q, parse_expr(cond))
                if rest == "then":
                    yield f"if ({c}) {{" + tag, 0, 1
                else:
                    yield f"if ({c}) {{" + tag, 0, 1
                    for x in self.statement(q, rest, no, blocks):
                        yield x
                    yield "}", -1, 0
                return
        if t == "do":
            yield "for (;;) {" + tag, 0, 1
            return
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", t)
        if m:
            var, ctl = m.group(1), split_top(m.group(2))
            v = self.emit(q, Node("name", var))
            lo, hi = self.emit(q, parse_expr(ctl[0])), self.emit(q, parse_expr(ctl[1]))
            if len(ctl) == 3:
                raise SyntaxError(f"{q.file}:{no}: do with a stride is not in the subset")
            # Fortran evaluates the bounds once and leaves var = hi+1 after a complete loop
            yield f"for (int hi__{no} = {hi}, {v}__{no} = ({v} = {lo}, 0); {v} <= hi__{no}; ++{v}) {{ (void){v}__{no};" + tag, 0, 1
            return
        m = re.match(r"^call\s+(\w+)\s*(\((.*)\))?$", t)
        if m:
            name = m.group(1)
            if name == "cpu_time":
                yield f"/* cpu_time */", 0, 0
                return
            if name not in self.procs:
                raise SyntaxError(f"{q.file}:{no}: call of unknown procedure {name}")
            args = [Parser(tokenize(a)).arg() for a in split_top(m.group(3))] if m.group(3) else []
            yield self.emit_proc_call(q, name, args) + ";" + tag, 0, 0
            return
        m = re.match(r"^open\s*\((.*)\)$", t)
        if m:
            kv = dict(x.split("=", 1) for x in split_top(m.group(1)) if "=" in x)
            yield f"f90rt::open_unit({kv['unit'].strip()}, {kv['file'].strip().replace(chr(39), chr(34))});" + tag, 0, 0
            return
        if re.match(r"^(print|rewind|close)\b", t) or re.match(r"^write\s*\(\*", t) or "nml=" in t:
            yield f"/* i/o dropped: {t[:60]} */", 0, 0
            return
        m = re.match(r"^(write|read)\s*\(", t)
        if m:
            j = matching_paren(t, t.index("("))
            ctl, items = split_top(t[t.index("(") + 1:j]), t[j + 1:].strip()
            unit = ctl[0].replace("unit=", "")
            yield "{ f90rt::Rec rec__;" + tag, 0, 1
            its = [Parser(tokenize(a)).arg() for a in split_top(items)] if items else []
            for x in self.io_items(q, m.group(1), its):
                yield x
            yield f"f90rt::{m.group(1)}_rec({unit}, rec__); }}", 0, -1
            if m.group(1) == "read":
                for x in self.io_items(q, "readback", its):
                    yield x
            return
        m = re.match(r"^allocate\s*\((.*)\)$", t)
        if m:
            for ent in split_top(m.group(1)):
                mm = re.match(r"^(\w+)\((.*)\)$", ent)
                s = self.lookup(q, mm.group(1))
                n = self.cname(s.name)
                dims = parse_dims(mm.group(2))
                parts = []
                for d, (lo, hi) in enumerate(dims, 1):
                    parts.append(f"{n}_l{d} = {self.emit(q, parse_expr(lo))}; {n}_n{d} = ({self.emit(q, parse_expr(hi))}) - {n}_l{d} + 1;")
                size = " * ".join(f"(size_t){n}_n{d}" for d in range(1, len(dims) + 1))
                yield " ".join(parts) + f" {n}_p = ({CTYPE[s.ty]}*)f90rt::alloc(({size}) * sizeof({CTYPE[s.ty]}));" + tag, 0, 0
            return
        m = re.match(r"^deallocate\s*\((.*)\)$", t)
        if m:
            for ent in split_top(m.group(1)):
                n = self.cname(ent.strip())
                yield f"f90rt::dealloc({n}_p); {n}_p = nullptr;", 0, 0
            return
        eq = self.find_assign(t)
        if eq is not None:
            lhs, rhs = t[:eq].strip(), t[eq + 1:].strip()
            le = Parser(tokenize(lhs)).arg() if False else parse_expr(lhs)
            re_ = parse_expr(rhs)
            ls = self.lookup(q, le.v) if le.k in ("name", "call") else None
            if ls is None and q.implicit_int and le.k == "name":
                q.syms[le.v] = Sym(le.v, "int")
                ls = q.syms[le.v]
            if ls is None:
                raise SyntaxError(f"{q.file}:{no}: assignment to unknown {lhs}")
            whole = ls.dims is not None and (le.k == "name" or (le.k == "call" and all(a.k == "colon" and a.a == [None, None] for a in le.a)))
            if whole:
                n = self.cname(ls.name)
                size = " * ".join(f"{n}_n{d}" for d in range(1, len(ls.dims) + 1))
                yield f"for (int i__ = 0; i__ < {size}; ++i__) {n}_p[i__] = {self.emit(q, re_)};" + tag, 0, 0
                return
            lt, rt = ls.ty, self.typeof(q, re_)
            r = self.emit(q, re_)
            if lt != rt and lt in CTYPE and lt != "char":
                r = f"({CTYPE[lt]})({r})"
            yield f"{self.emit(q, le)} = {r};" + tag, 0, 0
            return
        raise SyntaxError(f"{q.file}:{no}: statement outside the subset: {t!r}")

    def io_items(self, q, mode, items):
        for it in items:
            if it.k == "implied":
                *vals, ctl, hi = it.a
                var, lo = ctl.v, ctl.a[0]
                v = self.emit(q, Node("name", var))
                yield f"for ({v} = {self.emit(q, lo)}; {v} <= {self.emit(q, hi)}; ++{v}) {{", 0, 1
                for x in self.io_items(q, mode, vals):
                    yield x
                yield "}", -1, 0
                continue
            ty = self.typeof(q, it)
            if mode == "write":
                if ty == "char":
                    yield f"rec__.s.push_back({self.emit(q, it)});", 0, 0
                else:
                    yield f"rec__.v.push_back((double)({self.emit(q, it)}));", 0, 0
            elif mode == "read":
                yield "rec__.n += 1;", 0, 0
            else:
                yield f"{self.emit(q, it)} = ({CTYPE[ty]})f90rt::read_next();", 0, 0

    # ---------------------------------------------------------------- whole program
    def generate(self):
        self.analyse()
        out = ["// GENERATED by oracle/f90toc/f90toc.py from the reference's Fortran sources -- do not edit, do not commit.",
               '#include "f90rt.h"', ""]
        for n in sorted(self.count):
            out.append(f"long long calls_{self.cname(n)} = 0;      // instrumentation: calls of {n}")
        # module / common variables
        for s in self.modvars.values():
            ct, n = CTYPE[s.ty], self.cname(s.name)
            if s.dims is not None and not s.allocatable:
                fake = Proc("", "module", [], None)
                if s.init:
                    vals = ", ".join(self.emit(None, parse_expr(v)) for v in split_top(s.init))
                    out.append(f"{ct} {n}_p[] = {{{vals}}};")
                    out.append(f"const int {n}_l1 = {self.const_eval(s.dims[0][0])};")
                else:
                    lo, hi = s.dims[0]
                    out.append(f"const int {n}_l1 = {self.const_eval(lo)}, {n}_n1 = ({self.const_eval(hi)}) - ({self.const_eval(lo)}) + 1;")
                    out.append(f"{ct} {n}_p[{n}_n1];")
                out.append(f"#define {n}(i1) {n}_p[(i1) - {n}_l1]")
            elif s.dims is not None:
                nd = len(s.dims)
                out.append(f"{ct}* {n}_p = nullptr; " + " ".join(f"int {n}_l{d} = 1, {n}_n{d} = 0;" for d in range(1, nd + 1)))
                tmp = []
                self.array_macros(None, s, tmp)
                out.append(tmp[-1])
            else:
                init = f" = {self.const_eval(s.init)}" if s.init else f" = {ct}()"
                out.append(f"{ct} {n}{init};")
        out.append("")
        todo = [q for q in (self.procs[n] for n in self.order)]
        for q in todo:
            out.append(self.proto(q) + ";")
        out.append("")
        for q in todo:
            if q.name in self.skip:
                out.append(f"// {q.name}: supplied by the glue (I/O procedure)")
                continue
            out.extend(self.translate_proc(q))
            out.append("")
        return "\n".join(out) + "\n"

    def const_eval(self, s):
        """integer constant expressions of COMMON/DATA declarations (N-1, N1, ...)"""
        consts = {"n": 624, "n1": 625}
        e = parse_expr(s)

        def ev(x):
            if x.k == "num":
                return int(x.v)
            if x.k == "name":
                return consts[x.v]
            if x.k == "paren":
                return ev(x.a[0])
            if x.k == "un":
                return -ev(x.a[0]) if x.v == "-" else ev(x.a[0])
            if x.k == "bin":
                a, b = ev(x.a[0]), ev(x.a[1])
                return {"+": a + b, "-": a - b, "*": a * b}[x.v]
            raise SyntaxError(s)
        return str(ev(e))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", required=True)
    ap.add_argument("--files", default="global_mod.f90,bessel_skip,pbc_mod.f90,interpolate.f90,r8_gamma.f90,random_mod.f90,system_mod.f90,sample_mod.f90,vpi_mod.f90,vpi.f90")
    ap.add_argument("--skip", default="readparameters,readsystemparameters,mtsavef,mtgetf,checkpoint")
    ap.add_argument("--count", default="updateaction", help="procedures whose calls are counted (the metric's unit)")
    a = ap.parse_args()
    tr = Translator(skip=a.skip.split(","), count=[c for c in a.count.split(",") if c])
    for f in a.files.split(","):
        if f.endswith("_skip"):
            continue
        tr.load(f"{a.ref}/{f}")
    open(a.out, "w").write(tr.generate())
    print(f"wrote {a.out}: {len(tr.procs)} procedures", file=sys.stderr)


if __name__ == "__main__":
    main()
