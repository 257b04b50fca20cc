"""f77_interp.py -- a small interpreter for the fixed-form Fortran 77 subset the reference's list bookkeeping is written in.

TEST INFRASTRUCTURE (oracle/): it exists to PIN oracle/regcor_oracle.c.  This image has no Fortran compiler, so the
reference's own source text is executed instead: ``oracle/make_regcor_golden.py`` reads
``/root/reference/src/Main/regcor_gpu.F`` (lines 263-459) and ``util_gpu.F`` (lines 102-111) where they lie -- nothing is
copied into this repository -- runs them statement by statement on seeded inputs through this interpreter, and commits the
outputs as golden vectors under ``tests/golden/``.  No hand transcription sits between the reference's text and the vectors.

Subset: assignment (scalars, array elements, ``A(lo:hi) = expr`` sections), logical IF, block IF / ELSE IF / ELSE / END IF,
GO TO, labelled and unlabelled DO (trip count fixed at entry, as the standard says) with CONTINUE / END DO, CONTINUE;
comment lines (``*``, ``C``, ``c``, ``!`` in column 1, which also covers ``!$omp`` directives), blank lines, continuation
lines (column 6).  Anything else (CALL, WRITE, ...) raises: a region that needs it is outside what this tool can certify.

Arithmetic: INTEGER is a Python int (division truncates toward zero), REAL*8 a Python float -- every operation is the
single IEEE-754 double operation the statement names, evaluated left to right with Fortran's precedence, no contraction.
``x**2`` is ``x*x`` (what every Fortran compiler emits).  The reference declares IMPLICIT REAL*8 (A-H,O-Z)
(include/common6.h); default-kind REAL literals such as ``4.0`` are accepted only when they are exactly representable in
single precision, so that their promotion to REAL*8 is the same number.
"""
from __future__ import annotations

import math
import re
import struct


class F77Error(Exception):
    pass


# ------------------------------------------------------------------------------------------------ source -> statements
def read_statements(path, first, last):
    """Statements of lines first..last (1-based, inclusive) of a fixed-form source file: [(label or None, text, line)]."""
    with open(path, "r", errors="replace") as fh:
        lines = fh.read().split("\n")
    out = []
    for no in range(first, last + 1):
        raw = lines[no - 1].rstrip("\r").expandtabs(8)
        if not raw.strip():
            continue
        if raw[0] in "*Cc!#":
            continue
        raw = raw[:72]
        head, cont, body = raw[:5], raw[5:6], raw[6:]
        if cont.strip() and cont != "0" and not head.strip():
            if not out:
                raise F77Error("line %d: continuation without a statement" % no)
            lab, txt, l0 = out[-1]
            out[-1] = (lab, txt + body.strip(), l0)
            continue
        label = int(head) if head.strip() else None
        if "!" in body:                                   # trailing comment (no character constants in this subset)
            body = body[:body.index("!")]
        out.append((label, body.strip(), no))
    return [(lab, re.sub(r"\s+", " ", txt.upper()), no) for lab, txt, no in out if txt.strip()]


# ------------------------------------------------------------------------------------------------------- expressions
_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[ED][+-]?\d+)?|\.\d+(?:[ED][+-]?\d+)?|\d+[ED][+-]?\d+|\d+)"
                    r"|(\.(?:EQ|NE|LT|LE|GT|GE|AND|OR|NOT|TRUE|FALSE)\.)|([A-Z][A-Z0-9_]*)|(\*\*|[-+*/(),:=]))")


def tokenize(s):
    toks, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m:
            raise F77Error("cannot tokenize %r at %d" % (s, pos))
        num, dot, name, op = m.groups()
        if num is not None:
            # "1.EQ." must not swallow the dot of a relational operator: re-lex "digits." followed by a dotted keyword
            if num.endswith(".") and re.match(r"(?:EQ|NE|LT|LE|GT|GE|AND|OR|NOT)\.", s[m.end():]):
                num = num[:-1]
                pos = m.end() - 1
            else:
                pos = m.end()
            toks.append(("num", num))
            continue
        pos = m.end()
        if dot is not None:
            toks.append(("dot", dot))
        elif name is not None:
            toks.append(("name", name))
        else:
            toks.append(("op", op))
    return toks


def _number(text):
    if re.fullmatch(r"\d+", text):
        return int(text)
    if "D" in text:
        return float(text.replace("D", "E"))
    v = float(text)                                       # default-kind REAL literal: REAL*4 promoted to REAL*8
    if struct.unpack("f", struct.pack("f", v))[0] != v:
        raise F77Error("single-precision literal %s is not exactly representable: its promotion differs from %r" % (text, v))
    return v


class Parser:
    """Recursive descent with Fortran's precedence; produces nested tuples."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if (kind is not None and k != kind) or (val is not None and v != val):
            raise F77Error("expected %s %s, found %s %s" % (kind, val, k, v))
        self.i += 1
        return v

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek() == ("dot", ".OR."):
            self.take(); a = ("or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek() == ("dot", ".AND."):
            self.take(); a = ("and", a, self.p_not())
        return a

    def p_not(self):
        if self.peek() == ("dot", ".NOT."):
            self.take(); return ("not", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        k, v = self.peek()
        if k == "dot" and v in (".EQ.", ".NE.", ".LT.", ".LE.", ".GT.", ".GE."):
            self.take(); return ("rel", v, a, self.p_add())
        return a

    def p_add(self):
        k, v = self.peek()
        if k == "op" and v in "+-":
            self.take()
            a = self.p_mul()
            if v == "-":
                a = ("neg", a)
        else:
            a = self.p_mul()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("+", "-"):
                self.take(); a = ("bin", v, a, self.p_mul())
            else:
                return a

    def p_mul(self):
        a = self.p_pow()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("*", "/"):
                self.take(); a = ("bin", v, a, self.p_pow())
            else:
                return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek() == ("op", "**"):
            self.take(); return ("pow", a, self.p_pow())          # right associative
        return a

    def p_primary(self):
        k, v = self.peek()
        if k == "num":
            self.take(); return ("num", _number(v))
        if k == "dot" and v in (".TRUE.", ".FALSE."):
            self.take(); return ("num", v == ".TRUE.")
        if k == "name":
            self.take()
            if self.peek() == ("op", "("):
                self.take()
                args = [self.expr()]
                while self.peek() == ("op", ","):
                    self.take(); args.append(self.expr())
                self.take("op", ")")
                return ("ref", v, args)
            return ("var", v)
        if (k, v) == ("op", "("):
            self.take(); a = self.expr(); self.take("op", ")"); return a
        raise F77Error("unexpected token %s %s" % (k, v))


def parse_expr(text):
    p = Parser(tokenize(text))
    e = p.expr()
    if p.i != len(p.t):
        raise F77Error("trailing tokens in expression %r" % text)
    return e


class FArray:
    """A Fortran array seen through get(idx tuple) / set(idx tuple, value); indices are the Fortran ones."""

    def __init__(self, get, set_=None):
        self.get, self.set = get, set_


def farray_numpy(a, lower=1, integer=False):
    """View of a numpy array whose FIRST Fortran index is the LAST numpy axis reversed (X(3,N) <-> x[N][3]); 1-D arrays as is.
    lower: lower bound of the last Fortran index (the particle number of element 0) or of the only index."""
    conv = int if integer else float
    if a.ndim == 1:
        return FArray(lambda ix: conv(a[ix[0] - lower]), lambda ix, v: a.__setitem__(ix[0] - lower, v))
    if a.ndim == 2:
        return FArray(lambda ix: conv(a[ix[1] - lower, ix[0] - 1]), lambda ix, v: a.__setitem__((ix[1] - lower, ix[0] - 1), v))
    raise F77Error("unsupported rank")


_INTRINSICS = {
    "SQRT": lambda x: math.sqrt(x), "DSQRT": lambda x: math.sqrt(x), "FLOAT": lambda x: float(x), "DBLE": lambda x: float(x),
    "ABS": lambda x: abs(x), "MIN": min, "MAX": max, "INT": lambda x: int(x),
}


class Machine:
    def __init__(self, statements, scalars, arrays, max_steps=50_000_000):
        self.st = statements
        self.env = dict(scalars)
        self.arr = dict(arrays)
        self.max_steps = max_steps
        self.labels = {lab: k for k, (lab, _, _) in enumerate(statements) if lab is not None}
        self.code = [self._compile(k, txt, no) for k, (_, txt, no) in enumerate(statements)]
        self._link_blocks()

    # --- compile ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _split_if(txt):
        """'IF (cond) rest' -> (cond, rest)."""
        assert txt.startswith("IF")
        s = txt[2:].lstrip()
        if not s.startswith("("):
            raise F77Error("bad IF: %r" % txt)
        depth = 0
        for k, ch in enumerate(s):
            depth += ch == "("
            depth -= ch == ")"
            if depth == 0:
                return s[1:k], s[k + 1:].strip()
        raise F77Error("unbalanced IF: %r" % txt)

    def _compile(self, k, txt, no):
        try:
            return self._compile1(txt)
        except F77Error as e:
            raise F77Error("line %d: %s   [%s]" % (no, e, txt))

    def _compile1(self, txt):
        if txt in ("CONTINUE",):
            return ("nop",)
        if txt in ("END IF", "ENDIF"):
            return ["endif"]
        if txt in ("END DO", "ENDDO"):
            return ["enddo"]
        if txt == "ELSE":
            return ["else", None]
        m = re.fullmatch(r"GO ?TO (\d+)", txt)
        if m:
            return ("goto", int(m.group(1)))
        if txt.startswith("ELSE IF") or txt.startswith("ELSEIF"):
            cond, rest = self._split_if(txt[4:].lstrip())
            if rest != "THEN":
                raise F77Error("bad ELSE IF")
            return ["elseif", parse_expr(cond), None, None]
        if re.match(r"IF ?\(", txt):
            cond, rest = self._split_if(txt)
            if rest == "THEN":
                return ["ifthen", parse_expr(cond), None, None]
            return ("if", parse_expr(cond), self._compile1(rest))
        m = re.fullmatch(r"DO (?:(\d+) )?([A-Z][A-Z0-9_]*) ?= ?(.+)", txt)
        if m and "," in m.group(3):
            parts = self._split_commas(m.group(3))
            if len(parts) not in (2, 3):
                raise F77Error("bad DO")
            return ["do", int(m.group(1)) if m.group(1) else None, m.group(2)] + [[parse_expr(p) for p in parts]] + [None]
        if re.match(r"(CALL|WRITE|READ|PRINT|RETURN|STOP|END)\b", txt):
            raise F77Error("statement outside the supported subset")
        # assignment
        toks = tokenize(txt)
        depth, eq = 0, None
        for i, (kk, v) in enumerate(toks):
            if kk == "op":
                depth += v == "("
                depth -= v == ")"
                if v == "=" and depth == 0:
                    eq = i
                    break
        if eq is None:
            raise F77Error("not an assignment")
        lhs, rhs = toks[:eq], toks[eq + 1:]
        p = Parser(rhs); value = p.expr()
        if p.i != len(rhs):
            raise F77Error("trailing tokens")
        name = lhs[0][1]
        if len(lhs) == 1:
            return ("set", name, value)
        # A(i[,j]); any index may be a section lo:hi (the value is a scalar broadcast over it)
        inner = lhs[2:-1]
        idx, cur, depth = [], [], 0
        for t in inner:
            if t == ("op", ",") and depth == 0:
                idx.append(cur); cur = []
                continue
            depth += t == ("op", "(")
            depth -= t == ("op", ")")
            cur.append(t)
        idx.append(cur)

        def one(ix):
            if ("op", ":") in ix:
                c = ix.index(("op", ":"))
                return ("range", Parser(ix[:c]).expr(), Parser(ix[c + 1:]).expr())
            return ("at", Parser(ix).expr())
        parsed = [one(ix) for ix in idx]
        if any(q[0] == "range" for q in parsed):
            return ("setsec", name, parsed, value)
        return ("setel", name, [q[1] for q in parsed], value)

    @staticmethod
    def _split_commas(s):
        parts, depth, cur = [], 0, ""
        for ch in s:
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
                continue
            depth += ch == "("
            depth -= ch == ")"
            cur += ch
        parts.append(cur)
        return parts

    def _link_blocks(self):
        stack = []
        for k, c in enumerate(self.code):
            op = c[0]
            if op == "ifthen":
                stack.append(("if", [k]))
            elif op in ("elseif", "else"):
                if not stack or stack[-1][0] != "if":
                    raise F77Error("ELSE without IF (the region must hold whole blocks)")
                stack[-1][1].append(k)
            elif op == "endif":
                kind, chain = stack.pop()
                if kind != "if":
                    raise F77Error("END IF closes a DO")
                chain.append(k)
                for a, b in zip(chain[:-1], chain[1:]):
                    cc = self.code[a]
                    if cc[0] in ("ifthen", "elseif"):
                        cc[2], cc[3] = b, k                # next clause, end of block
                    else:
                        cc[1] = k
            elif op == "do":
                stack.append(("do", k))
                if c[1] is not None and c[1] not in self.labels:
                    raise F77Error("DO terminator %d outside the region" % c[1])
            elif op == "enddo":
                kind, d = stack.pop()
                if kind != "do" or self.code[d][1] is not None:
                    raise F77Error("END DO without an unlabelled DO")
                self.code[d][4] = k
                c.append(d)
            # labelled DO terminators
            lab = self.st[k][0]
            while stack and stack[-1][0] == "do" and self.code[stack[-1][1]][1] is not None and self.code[stack[-1][1]][1] == lab:
                d = stack.pop()[1]
                self.code[d][4] = k
        if stack:
            raise F77Error("unterminated block in the region")

    # --- evaluate ---------------------------------------------------------------------------------------------------
    def ev(self, e):
        op = e[0]
        if op == "num":
            return e[1]
        if op == "var":
            try:
                return self.env[e[1]]
            except KeyError:
                raise F77Error("variable %s used before it is defined" % e[1])
        if op == "ref":
            name, args = e[1], [self.ev(a) for a in e[2]]
            if name in self.arr:
                return self.arr[name].get(tuple(args))
            if name in _INTRINSICS:
                return _INTRINSICS[name](*args)
            raise F77Error("unknown array or function %s" % name)
        if op == "bin":
            a, b = self.ev(e[2]), self.ev(e[3])
            o = e[1]
            if o == "+": return a + b
            if o == "-": return a - b
            if o == "*": return a * b
            if isinstance(a, int) and isinstance(b, int):
                q = abs(a) // abs(b)
                return q if (a >= 0) == (b >= 0) else -q
            return a / b
        if op == "neg":
            return -self.ev(e[1])
        if op == "pow":
            a, b = self.ev(e[1]), self.ev(e[2])
            if isinstance(b, int) and 0 <= b <= 4:
                r = 1 if isinstance(a, int) else 1.0
                for _ in range(b):
                    r = r * a
                return r
            return a ** b
        if op == "rel":
            a, b = self.ev(e[2]), self.ev(e[3])
            return {".EQ.": a == b, ".NE.": a != b, ".LT.": a < b, ".LE.": a <= b, ".GT.": a > b, ".GE.": a >= b}[e[1]]
        if op == "and":
            a = self.ev(e[1]); b = self.ev(e[2])          # Fortran may evaluate both operands: so do we
            return bool(a) and bool(b)
        if op == "or":
            a = self.ev(e[1]); b = self.ev(e[2])
            return bool(a) or bool(b)
        if op == "not":
            return not self.ev(e[1])
        raise F77Error("bad expression node %r" % (op,))

    @staticmethod
    def _is_int_name(name):
        return name[0] in "IJKLMN"

    def _store(self, name, v):
        if self._is_int_name(name):
            v = int(v)
        elif not isinstance(v, bool):
            v = float(v)
        self.env[name] = v

    def run(self, start_label=None):
        pc = self.labels[start_label] if start_label is not None else 0
        loops = []                                        # [do index, var, remaining trips, step]
        steps = 0
        n = len(self.code)

        def simple(c):
            nonlocal pc, loops
            op = c[0]
            if op == "nop":
                return
            if op == "set":
                self._store(c[1], self.ev(c[2])); return
            if op == "setel":
                self.arr[c[1]].set(tuple(self.ev(a) for a in c[2]), self.ev(c[3])); return
            if op == "setsec":
                import itertools
                v = self.ev(c[3])
                axes = [range(self.ev(q[1]), self.ev(q[2]) + 1) if q[0] == "range" else [self.ev(q[1])] for q in c[2]]
                for ix in itertools.product(*axes):
                    self.arr[c[1]].set(tuple(ix), v)
                return
            if op == "goto":
                tgt = self.labels.get(c[1])
                if tgt is None:
                    raise F77Error("GO TO %d leaves the region" % c[1])
                loops = [l for l in loops if l[0] < tgt <= self.code[l[0]][4]]
                pc = tgt - 1                              # the main loop adds 1
                return
            raise F77Error("statement %r not allowed here" % (op,))

        while pc < n:
            steps += 1
            if steps > self.max_steps:
                raise F77Error("step limit exceeded")
            c = self.code[pc]
            op = c[0]
            if op == "if":
                if self.ev(c[1]):
                    simple(c[2])
            elif op == "ifthen":
                if not self.ev(c[1]):
                    nxt = c[2]
                    while True:                           # walk the clauses until one is taken
                        cc = self.code[nxt]
                        if cc[0] == "elseif":
                            if self.ev(cc[1]):
                                pc = nxt; break
                            nxt = cc[2]
                        elif cc[0] == "else":
                            pc = nxt; break
                        else:                             # endif
                            pc = nxt; break
            elif op == "elseif":
                pc = c[3]                                 # reached by falling out of the previous clause
            elif op == "else":
                pc = c[1]
            elif op == "endif":
                pass
            elif op == "do":
                parts = [self.ev(p) for p in c[3]]
                lo, hi, stp = parts[0], parts[1], (parts[2] if len(parts) > 2 else 1)
                trips = max((hi - lo + stp) // stp, 0)
                self._store(c[2], lo)
                if trips == 0:
                    pc = c[4]                             # skip the body; the terminator itself is not executed
                    if self.code[pc][0] not in ("enddo", "nop"):
                        raise F77Error("DO terminator must be CONTINUE or END DO in this subset")
                else:
                    loops.append([pc, c[2], trips, stp])
            else:
                if op != "enddo":
                    simple(c)
            # loop terminators (a labelled CONTINUE or END DO): the innermost active DO ending here
            while loops and self.code[loops[-1][0]][4] == pc:
                l = loops[-1]
                l[2] -= 1
                self._store(l[1], self.env[l[1]] + l[3])
                if l[2] > 0:
                    pc = l[0]                             # back to the first statement of the body
                    break
                loops.pop()
            pc += 1
        return self.env
