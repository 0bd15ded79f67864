#!/usr/bin/env python3
"""Cycle-accurate evaluator for the MATLAB "Filter Design HDL Coder" VHDL subset the reference's FPGA
filters are written in - it executes the reference's OWN source text.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): used by tests/ and tools/gen_golden_hdl.py to pin the
golden C model (oracle/ddc_golden.c, duc_golden.c) and, through it, the CUDA kernels to the HDL itself.  No
HDL simulator exists in the image (no ghdl / nvc / iverilog / verilator), so this file is one: it reads
`/root/reference/FPGA/{rx_cic,rx_ciccomp,rx_hilb,tx_cic,tx_ciccomp}.vhd` where they lie, parses entity,
constants, signals, concurrent (conditional) signal assignments and clocked processes, types every expression
by the IEEE numeric_std rules and then either
  * interprets the design on Python integers (`Design.instance()`: literal bit-vector semantics, slow), or
  * emits a C translation (`Design.emit_c()`), which oracle/hdl/Makefile compiles into oracle/_ref/ (a build
    product derived from the reference's sources: git-ignored, never copied into the repository).
The two back ends share the parser and the type checker but not the arithmetic, and the tests run them
against each other.  The type checker is itself a check: VHDL requires both sides of every assignment to
have the same width, so a wrong numeric_std rule (result width of "+", "*", "&", resize ...) fails loudly
on the reference's own 2 900 lines.

numeric_std semantics implemented (IEEE 1076.3):
  signed +/- signed   -> width max(L,R), wraps;  with an integer -> width of the vector operand
  signed * signed     -> width L+R
  - signed            -> same width, wraps (-(-2^(n-1)) = -2^(n-1))
  a & b               -> concatenation; '0'/'1' and "0101" literals take the type of the other operand
  resize(signed, n)   -> sign-extends, or TRUNCATES KEEPING THE SIGN BIT plus the n-1 low bits
  resize(unsigned, n) -> zero-extends or keeps the n low bits
  shift_right(signed) -> arithmetic;  x(h DOWNTO l) slices keep the kind of x
  relational operators compare numerically; a string literal compared with signed is signed
  operators of equal precedence ("&", "+", "-") associate to the left
Process semantics: every signal read inside a process sees the value before the clock edge, all registers of
a module update together, concurrent assignments settle (in dependency order) after the edge.
"""
import os
import re
import sys

# ----------------------------------------------------------------------------------------------
# lexer
# ----------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<chr>'[01]')
  | (?P<str>"[01]*")
  | (?P<num>\d+)
  | (?P<id>[A-Za-z][A-Za-z0-9_]*)
  | (?P<sym><=|>=|/=|:=|=>|<>|[()\[\];:,&+\-*=<>'.])
""", re.X)


def lex(text):
    out, i = [], 0
    while i < len(text):
        m = _TOKEN.match(text, i)
        if not m:
            raise SyntaxError("vhdl_eval: cannot tokenise at %r" % text[i:i + 30])
        i = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        if k == "id":
            v = v.lower()
        out.append((k, v, text.count("\n", 0, m.start()) + 1))
    out.append(("eof", "", 0))
    return out


# ----------------------------------------------------------------------------------------------
# types
# ----------------------------------------------------------------------------------------------
class T:
    """kind: signed | unsigned | slv | vec (untyped bit-string) | bit | bool | int | array"""
    __slots__ = ("kind", "w", "lo", "hi", "elem")

    def __init__(self, kind, w=0, lo=0, hi=0, elem=None):
        self.kind, self.w, self.lo, self.hi, self.elem = kind, w, lo, hi, elem

    def __eq__(self, o):
        return (self.kind, self.w, self.lo, self.hi, self.elem) == (o.kind, o.w, o.lo, o.hi, o.elem)

    def __repr__(self):
        if self.kind == "array":
            return "array(%d..%d of %r)" % (self.lo, self.hi, self.elem)
        return "%s%s" % (self.kind, "(%d)" % self.w if self.kind in ("signed", "unsigned", "slv", "vec") else "")

    @property
    def is_vec(self):
        return self.kind in ("signed", "unsigned", "slv", "vec")


BIT, BOOL, INT = T("bit", 1), T("bool", 1), T("int")


class N:
    """AST node"""
    __slots__ = ("op", "a", "t", "line")

    def __init__(self, op, *a, line=0):
        self.op, self.a, self.t, self.line = op, list(a), None, line

    def __repr__(self):
        return "%s%r" % (self.op, tuple(self.a))


FUNCS = {"resize", "shift_right", "shift_left", "to_signed", "to_unsigned", "signed", "unsigned", "std_logic_vector"}


# ----------------------------------------------------------------------------------------------
# parser
# ----------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, text):
        self.tk = lex(text)
        self.i = 0

    def peek(self, k=0):
        return self.tk[self.i + k]

    def next(self):
        t = self.tk[self.i]
        self.i += 1
        return t

    def accept(self, v):
        if self.peek()[1] == v and self.peek()[0] in ("id", "sym"):
            return self.next()
        return None

    def expect(self, v):
        t = self.next()
        if t[1] != v:
            raise SyntaxError("vhdl_eval: line %d: expected %r, found %r" % (t[2], v, t[1]))
        return t

    def ident(self):
        t = self.next()
        if t[0] != "id":
            raise SyntaxError("vhdl_eval: line %d: identifier expected, found %r" % (t[2], t[1]))
        return t[1]

    # ---- design units
    def parse_file(self):
        d = {"ports": [], "types": {}, "consts": [], "signals": [], "conc": [], "procs": []}
        while self.peek()[1] in ("library", "use"):
            while self.next()[1] != ";":
                pass
        self.expect("entity")
        d["name"] = self.ident()
        self.expect("is")
        self.expect("port")
        self.expect("(")
        while True:
            name = self.ident()
            self.expect(":")
            direction = self.ident()
            d["ports"].append((name, direction, self.parse_type(d)))
            if not self.accept(";"):
                break
        self.expect(")")
        self.expect(";")
        self.expect("end")
        self.ident()
        self.expect(";")
        self.expect("architecture")
        self.ident()
        self.expect("of")
        self.ident()
        self.expect("is")
        while not self.accept("begin"):
            kw = self.ident()
            if kw == "type":
                name = self.ident()
                for w in ("is", "array", "(", "natural", "range", "<>", ")", "of"):
                    self.expect(w)
                d["types"][name] = self.parse_type(d)
                self.expect(";")
            elif kw in ("constant", "signal"):
                name = self.ident()
                self.expect(":")
                ty = self.parse_type(d)
                init = None
                if self.accept(":="):
                    init = self.expr()
                self.expect(";")
                d["consts" if kw == "constant" else "signals"].append((name, ty, init))
            else:
                raise SyntaxError("vhdl_eval: unsupported declaration %r" % kw)
        while not self.accept("end"):
            self.concurrent(d)
        self.ident()
        self.expect(";")
        return d

    def const_int(self, e):
        if e.op == "int":
            return e.a[0]
        raise SyntaxError("vhdl_eval: constant integer expected in a range")

    def parse_type(self, d):
        name = self.ident()
        if name == "std_logic":
            return BIT
        self.expect("(")
        a = self.const_int(self.expr())
        direction = self.ident()
        b = self.const_int(self.expr())
        self.expect(")")
        if name in ("signed", "unsigned", "std_logic_vector") and direction == "downto":
            assert b == 0, "vector ranges are N DOWNTO 0 in this subset"
            return T({"std_logic_vector": "slv"}.get(name, name), a + 1)
        if name == "std_logic_vector" and direction == "to":        # ascending bit pipe = array of bits
            return T("array", 0, a, b, BIT)
        if name in d["types"] and direction == "to":
            return T("array", 0, a, b, d["types"][name])
        raise SyntaxError("vhdl_eval: unsupported type %s(%d %s %d)" % (name, a, direction, b))

    # ---- statements
    def target(self):
        line = self.peek()[2]
        n = N("name", self.ident(), line=line)
        while self.peek()[1] == "(":
            n = self.postfix(n)
        return n

    def concurrent(self, d):
        if self.peek()[0] == "id" and self.peek(1)[1] == ":":          # label : PROCESS
            self.ident()
            self.expect(":")
        if self.accept("process"):
            self.expect("(")
            while self.next()[1] != ")":
                pass
            self.expect("begin")
            body = self.seq_list(("end",))
            self.expect("end")
            self.expect("process")
            if self.peek()[0] == "id":
                self.ident()
            self.expect(";")
            d["procs"].append(body)
            return
        tgt = self.target()
        self.expect("<=")
        arms = []
        e = self.expr()
        while self.accept("when"):
            c = self.expr()
            self.expect("else")
            arms.append((c, e))
            e = self.expr()
        self.expect(";")
        d["conc"].append(N("cassign", tgt, arms, e, line=tgt.line))

    def seq_list(self, stops):
        out = []
        while self.peek()[1] not in stops:
            out.append(self.seq())
        return out

    def seq(self):
        if self.accept("if"):
            arms = []
            c = self.expr()
            self.expect("then")
            arms.append((c, self.seq_list(("elsif", "else", "end"))))
            els = []
            while True:
                if self.accept("elsif"):
                    c = self.expr()
                    self.expect("then")
                    arms.append((c, self.seq_list(("elsif", "else", "end"))))
                elif self.accept("else"):
                    els = self.seq_list(("end",))
                else:
                    break
            self.expect("end")
            self.expect("if")
            self.expect(";")
            return N("if", arms, els)
        tgt = self.target()
        self.expect("<=")
        e = self.expr()
        self.expect(";")
        return N("assign", tgt, e, line=tgt.line)

    # ---- expressions (VHDL precedence: logical < relational < adding(+,-,&) < sign < multiplying < not)
    def expr(self):
        l = self.relation()
        while self.peek()[1] in ("and", "or", "xor"):
            op = self.next()[1]
            l = N(op, l, self.relation())
        return l

    def relation(self):
        l = self.simple()
        if self.peek()[1] in ("=", "/=", "<", "<=", ">", ">="):
            op = self.next()[1]
            l = N("cmp", op, l, self.simple())
        return l

    def simple(self):
        line = self.peek()[2]
        neg = False
        if self.peek()[1] in ("+", "-"):
            neg = self.next()[1] == "-"
        l = self.term()
        if neg:
            l = N("neg", l, line=line)
        while self.peek()[1] in ("+", "-", "&"):
            op = self.next()[1]
            l = N({"+": "add", "-": "sub", "&": "cat"}[op], l, self.term(), line=line)
        return l

    def term(self):
        l = self.factor()
        while self.peek()[1] == "*":
            self.next()
            l = N("mul", l, self.factor())
        return l

    def factor(self):
        if self.accept("not"):
            return N("not", self.primary())
        return self.primary()

    def postfix(self, base):
        self.expect("(")
        first = self.expr()
        if self.peek()[1] in ("downto", "to"):
            direction = self.next()[1]
            second = self.expr()
            self.expect(")")
            return N("slice", base, self.const_int(first), direction, self.const_int(second), line=base.line)
        args = [first]
        while self.accept(","):
            args.append(self.expr())
        self.expect(")")
        if base.op == "name" and base.a[0] in FUNCS:
            return N("call", base.a[0], args, line=base.line)
        assert len(args) == 1
        return N("index", base, args[0], line=base.line)

    def primary(self):
        k, v, line = self.next()
        if k == "num":
            return N("int", int(v), line=line)
        if k == "chr":
            return N("bitlit", int(v[1]), line=line)
        if k == "str":
            return N("strlit", v[1:-1], line=line)
        if v == "(":
            if self.peek()[1] == "others":
                self.next()
                self.expect("=>")
                e = self.expr()
                self.expect(")")
                return N("others", e, line=line)
            e = self.expr()
            self.expect(")")
            return e
        if k == "id":
            n = N("name", v, line=line)
            if self.peek()[1] == "'" and self.peek(1)[1] == "event":
                self.next()
                self.next()
                return N("event", v, line=line)
            while self.peek()[1] == "(":
                n = self.postfix(n)
            return n
        raise SyntaxError("vhdl_eval: line %d: unexpected %r" % (line, v))


# ----------------------------------------------------------------------------------------------
# type checker (numeric_std result types)
# ----------------------------------------------------------------------------------------------
def _num_kind(a, b, line):
    ks = {a.kind, b.kind} - {"vec", "int"}
    if len(ks) != 1 or next(iter(ks)) not in ("signed", "unsigned"):
        raise TypeError("vhdl_eval: line %d: arithmetic on %r and %r" % (line, a, b))
    return next(iter(ks))


class Design:
    def __init__(self, path):
        self.path = path
        with open(path, "r") as f:
            self.text = f.read()
        d = Parser(self.text).parse_file()
        self.name = d["name"]
        self.ports = d["ports"]
        self.sym = {}
        for name, direction, ty in d["ports"]:
            self.sym[name] = ty
        self.consts = {}
        for name, ty, init in d["consts"]:
            self.sym[name] = ty
        for name, ty, init in d["signals"]:
            assert init is None
            self.sym[name] = ty
        self.signals = [s[0] for s in d["signals"]]
        for name, ty, init in d["consts"]:
            self.typ(init, want=ty)
            self.consts[name] = (ty, init)
        self.conc = d["conc"]
        self.procs = d["procs"]
        self.regs = set()
        for st in self.conc:
            self.check_assign(st.a[0], [e for _, e in st.a[1]] + [st.a[2]])
            for c, _ in st.a[1]:
                self.want_bool(c)
        for body in self.procs:
            self.check_seq(body)
        self.order_concurrent()

    # ---- typing
    def want_bool(self, c):
        t = self.typ(c)
        if t.kind != "bool":
            raise TypeError("vhdl_eval: %s line %d: condition is %r" % (self.name, c.line, t))

    def check_seq(self, body):
        for st in body:
            if st.op == "if":
                for c, b in st.a[0]:
                    self.want_bool(c)
                    self.check_seq(b)
                self.check_seq(st.a[1])
            else:
                self.check_assign(st.a[0], [st.a[1]])
                self.regs.add(self.base_name(st.a[0]))

    @staticmethod
    def base_name(n):
        while n.op != "name":
            n = n.a[0]
        return n.a[0]

    def check_assign(self, tgt, exprs):
        tt = self.typ(tgt)
        for e in exprs:
            et = self.typ(e, want=tt)
            ok = et == tt or (et.kind == "vec" and tt.is_vec and et.w == tt.w) or \
                (tt.is_vec and et.is_vec and et.kind == tt.kind and et.w == tt.w) or \
                (tt.kind == et.kind == "array" and tt.hi - tt.lo == et.hi - et.lo and tt.elem == et.elem)
            if not ok:
                raise TypeError("vhdl_eval: %s line %d: %s is %r but the expression is %r"
                                % (self.name, tgt.line, self.base_name(tgt), tt, et))

    def typ(self, n, want=None):
        t = self._typ(n, want)
        n.t = t
        return t

    def _typ(self, n, want):
        op = n.op
        if op == "name":
            if n.a[0] not in self.sym:
                raise NameError("vhdl_eval: %s line %d: unknown name %s" % (self.name, n.line, n.a[0]))
            return self.sym[n.a[0]]
        if op == "int":
            return INT
        if op == "bitlit":
            return BIT
        if op == "strlit":
            return T("vec", len(n.a[0]))
        if op == "others":
            if want is None:
                raise TypeError("vhdl_eval: aggregate without a target type")
            self.typ(n.a[0], want=want.elem if want.kind == "array" else BIT)
            return want
        if op == "event":
            return BOOL
        if op == "index":
            bt = self.typ(n.a[0])
            it = self.typ(n.a[1])
            assert it.kind == "int"
            i = n.a[1].a[0]
            if bt.kind == "array":
                assert bt.lo <= i <= bt.hi
                return bt.elem
            assert bt.is_vec and 0 <= i < bt.w, (self.name, n.line)
            return BIT
        if op == "slice":
            bt = self.typ(n.a[0])
            x, direction, y = n.a[1], n.a[2], n.a[3]
            if bt.kind == "array":
                assert direction == "to" and bt.lo <= x <= y <= bt.hi
                return T("array", 0, x, y, bt.elem)
            assert bt.is_vec and direction == "downto" and bt.w > x >= y >= 0, (self.name, n.line)
            return T(bt.kind, x - y + 1)
        if op == "call":
            fn, args = n.a
            if fn in ("to_signed", "to_unsigned"):
                self.typ(args[0]); self.typ(args[1])
                return T(fn[3:], args[1].a[0])
            if fn in ("signed", "unsigned", "std_logic_vector"):
                at = self.typ(args[0])
                assert at.is_vec
                return T({"std_logic_vector": "slv"}.get(fn, fn), at.w)
            if fn == "resize":
                at = self.typ(args[0]); self.typ(args[1])
                assert at.kind in ("signed", "unsigned"), (self.name, n.line, at)
                return T(at.kind, args[1].a[0])
            if fn in ("shift_right", "shift_left"):
                at = self.typ(args[0]); self.typ(args[1])
                assert at.kind in ("signed", "unsigned")
                return T(at.kind, at.w)
        if op in ("add", "sub"):
            a, b = self.typ(n.a[0]), self.typ(n.a[1])
            kind = _num_kind(a, b, n.line)
            return T(kind, max(a.w, b.w))       # an integer operand has w = 0
        if op == "mul":
            a, b = self.typ(n.a[0]), self.typ(n.a[1])
            kind = _num_kind(a, b, n.line)
            assert a.kind == b.kind == kind
            return T(kind, a.w + b.w)
        if op == "neg":
            a = self.typ(n.a[0])
            if a.kind == "int":                   # -18 in to_signed(-18, 16): fold the literal
                n.op, n.a = "int", [-n.a[0].a[0]]
                return INT
            assert a.kind == "signed"
            return T("signed", a.w)
        if op == "cat":
            a, b = self.typ(n.a[0]), self.typ(n.a[1])
            kinds = {a.kind, b.kind} - {"bit", "vec"}
            assert len(kinds) <= 1, (self.name, n.line, a, b)
            return T(next(iter(kinds)) if kinds else "vec", a.w + b.w)
        if op == "not":
            a = self.typ(n.a[0])
            assert a.kind in ("bit", "bool")
            return a
        if op in ("and", "or", "xor"):
            a, b = self.typ(n.a[0]), self.typ(n.a[1])
            assert a.kind == b.kind and a.kind in ("bit", "bool"), (self.name, n.line, a, b)
            return a
        if op == "cmp":
            a, b = self.typ(n.a[1]), self.typ(n.a[2])
            if a.kind == "bit" or b.kind == "bit":
                assert a.kind == b.kind == "bit" and n.a[0] in ("=", "/=")
            else:
                _num_kind(a, b, n.line)
            return BOOL
        raise TypeError("vhdl_eval: cannot type %r" % n)

    # ---- dependency order of the concurrent assignments
    def reads(self, n, acc):
        if isinstance(n, N):
            if n.op == "name":
                acc.add(n.a[0])
            for x in n.a:
                self.reads(x, acc)
        elif isinstance(n, (list, tuple)):
            for x in n:
                self.reads(x, acc)
        return acc

    def order_concurrent(self):
        by_target = {}
        for st in self.conc:
            name = self.base_name(st.a[0])
            assert st.a[0].op == "name" and name not in by_target and name not in self.regs, name
            by_target[name] = st
        deps = {k: {r for r in self.reads([st.a[1], st.a[2]], set()) if r in by_target} for k, st in by_target.items()}
        order, done = [], set()

        def visit(k, stack):
            if k in done:
                return
            if k in stack:
                raise ValueError("vhdl_eval: combinational loop through " + k)
            for r in sorted(deps[k]):
                visit(r, stack | {k})
            done.add(k)
            order.append(by_target[k])

        old = sys.getrecursionlimit()
        sys.setrecursionlimit(10000)
        for k in by_target:
            visit(k, frozenset())
        sys.setrecursionlimit(old)
        self.conc = order
        self.comb = [self.base_name(st.a[0]) for st in order]
        in_ports = {p[0] for p in self.ports if p[1] == "in"}
        for name in self.sym:
            assert name in in_ports or name in self.consts or name in self.regs or name in by_target, \
                "vhdl_eval: %s: signal %s is never driven" % (self.name, name)

    def instance(self):
        return Instance(self)

    def emit_c(self):
        return CEmitter(self).emit()


# ----------------------------------------------------------------------------------------------
# back end 1: interpreter on Python integers.  A vector value is its raw bit pattern (0 <= raw < 2^w);
# signedness is applied where numeric_std applies it.
# ----------------------------------------------------------------------------------------------
def _mask(w):
    return (1 << w) - 1


def _val(raw, t):
    """numeric value of raw bits under type t"""
    if t.kind == "signed" and raw >> (t.w - 1):
        return raw - (1 << t.w)
    return raw


class Instance:
    def __init__(self, design):
        self.d = design
        self.v = {}
        for name, (ty, init) in design.consts.items():
            self.v[name] = self.ev(init, ty)
        for name, ty in design.sym.items():
            if name not in self.v:
                self.v[name] = self.zero(ty)
        self.edge = False

    @staticmethod
    def zero(ty):
        return [Instance.zero(ty.elem) for _ in range(ty.hi - ty.lo + 1)] if ty.kind == "array" else 0

    # ---- public driver interface (values are Python ints; signed ports are passed/returned as raw bits)
    def set(self, port, raw):
        self.v[port] = raw & _mask(self.d.sym[port].w)

    def get(self, name):
        return self.v[name]

    def get_signed(self, name):
        w = self.d.sym[name].w
        raw = self.v[name]
        return raw - (1 << w) if raw >> (w - 1) else raw

    def settle(self):
        for st in self.d.conc:
            self.v[st.a[0].a[0]] = self.ev_cassign(st)

    def clock(self, clk="clk"):
        """one rising edge of clk with the current input ports; outputs are settled afterwards"""
        self.settle()
        self.v[clk] = 1
        self.edge = True
        nxt = {}
        for body in self.d.procs:
            self.run(body, nxt)
        self.edge = False
        self.v.update(nxt)
        self.settle()

    def apply_reset(self, reset="reset"):
        """asynchronous reset held: run the processes without a clock event"""
        self.v[reset] = 1
        nxt = {}
        for body in self.d.procs:
            self.run(body, nxt)
        self.v.update(nxt)
        self.v[reset] = 0
        self.settle()

    # ---- evaluation
    def ev_cassign(self, st):
        tt = st.a[0].t
        for c, e in st.a[1]:
            if self.ev(c):
                return self.ev(e, tt)
        return self.ev(st.a[2], tt)

    def run(self, body, nxt):
        for st in body:
            if st.op == "if":
                for c, b in st.a[0]:
                    if self.ev(c):
                        self.run(b, nxt)
                        break
                else:
                    self.run(st.a[1], nxt)
            else:
                tgt, e = st.a
                val = self.ev(e, tgt.t)
                if tgt.op == "name":
                    nxt[tgt.a[0]] = val
                else:
                    name = tgt.a[0].a[0]
                    bt = self.d.sym[name]
                    assert tgt.a[0].op == "name" and bt.kind == "array"
                    arr = nxt.get(name)
                    if arr is None:
                        arr = nxt[name] = list(self.v[name])
                    if tgt.op == "index":
                        arr[tgt.a[1].a[0] - bt.lo] = val
                    else:
                        x, y = tgt.a[1], tgt.a[3]
                        arr[x - bt.lo:y - bt.lo + 1] = val

    def ev(self, n, want=None):
        op = n.op
        if op == "name":
            return self.v[n.a[0]]
        if op == "int":
            return n.a[0]
        if op == "bitlit":
            return n.a[0]
        if op == "strlit":
            return int(n.a[0], 2) if n.a[0] else 0
        if op == "others":
            if want.kind == "array":
                return [self.ev(n.a[0], want.elem) for _ in range(want.hi - want.lo + 1)]
            return _mask(want.w) if self.ev(n.a[0], BIT) else 0
        if op == "event":
            return self.edge
        if op == "index":
            b = self.ev(n.a[0])
            bt = n.a[0].t
            i = n.a[1].a[0]
            return b[i - bt.lo] if bt.kind == "array" else (b >> i) & 1
        if op == "slice":
            b = self.ev(n.a[0])
            bt = n.a[0].t
            if bt.kind == "array":
                return list(b[n.a[1] - bt.lo:n.a[3] - bt.lo + 1])
            return (b >> n.a[3]) & _mask(n.a[1] - n.a[3] + 1)
        if op == "call":
            fn, args = n.a
            if fn in ("to_signed", "to_unsigned"):
                return args[0].a[0] & _mask(n.t.w) if args[0].op == "int" else self.ev(args[0]) & _mask(n.t.w)
            if fn in ("signed", "unsigned", "std_logic_vector"):
                return self.ev(args[0])
            a = self.ev(args[0])
            at = args[0].t
            if fn == "resize":
                nw = n.t.w
                if at.kind == "signed":
                    if nw >= at.w:
                        return _val(a, at) & _mask(nw)
                    sign = a >> (at.w - 1)
                    return (sign << (nw - 1)) | (a & _mask(nw - 1))
                return a & _mask(nw)
            k = args[1].a[0]
            if fn == "shift_right":
                return (_val(a, at) >> k) & _mask(at.w)
            if fn == "shift_left":
                return (a << k) & _mask(at.w)
        if op in ("add", "sub", "mul"):
            ta, tb = n.a[0].t, n.a[1].t
            kind = n.t.kind
            va = self.num(n.a[0], kind)
            vb = self.num(n.a[1], kind)
            r = va + vb if op == "add" else (va - vb if op == "sub" else va * vb)
            return r & _mask(n.t.w)
        if op == "neg":
            return (-_val(self.ev(n.a[0]), n.a[0].t)) & _mask(n.t.w)
        if op == "cat":
            return (self.ev(n.a[0]) << n.a[1].t.w) | self.ev(n.a[1])
        if op == "not":
            return 1 - int(self.ev(n.a[0]))
        if op == "and":
            return int(self.ev(n.a[0])) & int(self.ev(n.a[1]))
        if op == "or":
            return int(self.ev(n.a[0])) | int(self.ev(n.a[1]))
        if op == "xor":
            return int(self.ev(n.a[0])) ^ int(self.ev(n.a[1]))
        if op == "cmp":
            rel, l, r = n.a
            if l.t.kind == "bit":
                a, b = self.ev(l), self.ev(r)
            else:
                kind = _num_kind(l.t, r.t, n.line)
                a, b = self.num(l, kind), self.num(r, kind)
            return {"=": a == b, "/=": a != b, "<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[rel]
        raise ValueError("vhdl_eval: cannot evaluate %r" % n)

    def num(self, n, kind):
        """numeric value of an operand of an arithmetic/relational operator whose vector kind is `kind`"""
        raw = self.ev(n)
        if n.t.kind == "int":
            return raw
        return _val(raw, T(kind, n.t.w))


# ----------------------------------------------------------------------------------------------
# back end 2: C translation.  Every vector lives in an int64_t holding its NUMERIC value (signed kinds
# sign-extended, the others zero-extended); widths are limited to 62 bits, which the reference's 61-bit
# adders and 57-bit accumulators respect.
# ----------------------------------------------------------------------------------------------
class CEmitter:
    def __init__(self, design):
        self.d = design
        self.p = design.name

    def cname(self, name):
        return "v_" + name

    def norm(self, expr, t):
        """wrap an int64 expression to the numeric value of type t"""
        assert t.w <= 62, "vhdl_eval: vector wider than 62 bits"
        if t.kind == "signed":
            return "SX(%s, %d)" % (expr, t.w)
        return "ZX(%s, %d)" % (expr, t.w)

    def raw(self, n):
        """C expression for the raw (zero-extended) bit pattern of vector/bit expression n"""
        e = self.ex(n)
        if n.t.kind == "signed":
            return "ZX(%s, %d)" % (e, n.t.w)
        return e

    def num(self, n, kind):
        if n.t.kind == "int":
            return "INT64_C(%d)" % n.a[0]
        if n.t.kind == "vec" and kind == "signed":
            return "SX(%s, %d)" % (self.ex(n), n.t.w)
        assert n.t.kind in (kind, "vec"), (n, kind)
        return self.ex(n)

    def ex(self, n, want=None):
        op = n.op
        if op == "name":
            name = n.a[0]
            if name in self.d.consts:
                ty, init = self.d.consts[name]
                return self.ex(init, ty)
            return "s->" + self.cname(name)
        if op == "int":
            return "INT64_C(%d)" % n.a[0]
        if op == "bitlit":
            return "INT64_C(%d)" % n.a[0]
        if op == "strlit":
            return "INT64_C(%d)" % (int(n.a[0], 2) if n.a[0] else 0)
        if op == "others":
            assert want.kind != "array"
            bit = self.ex(n.a[0], BIT)
            full = -1 if want.kind == "signed" else _mask(want.w)
            return "((%s) ? INT64_C(%d) : INT64_C(0))" % (bit, full)
        if op == "event":
            return "1"
        if op == "index":
            bt = n.a[0].t
            i = n.a[1].a[0]
            if bt.kind == "array":
                assert n.a[0].op == "name"
                return "s->%s[%d]" % (self.cname(n.a[0].a[0]), i - bt.lo)
            return "((%s >> %d) & 1)" % (self.ex(n.a[0]), i)
        if op == "slice":
            bt = n.a[0].t
            assert bt.kind != "array"
            return self.norm("(%s >> %d)" % (self.ex(n.a[0]), n.a[3]), n.t)
        if op == "call":
            fn, args = n.a
            if fn in ("to_signed", "to_unsigned"):
                v = args[0].a[0] & _mask(n.t.w)
                return "INT64_C(%d)" % _val(v, n.t)
            if fn in ("signed", "unsigned", "std_logic_vector"):
                return self.norm(self.ex(args[0]), n.t)
            a = self.ex(args[0])
            at = args[0].t
            if fn == "resize":
                nw = n.t.w
                if at.kind == "signed":
                    if nw >= at.w:
                        return a
                    return "RESIZE_S(%s, %d)" % (a, nw)
                return a if nw >= at.w else "ZX(%s, %d)" % (a, nw)
            k = args[1].a[0]
            if fn == "shift_right":
                return "(%s >> %d)" % (a, k)
            if fn == "shift_left":
                return self.norm("(int64_t)((uint64_t)%s << %d)" % (a, k), n.t)
        if op in ("add", "sub", "mul"):
            kind = n.t.kind
            a, b = self.num(n.a[0], kind), self.num(n.a[1], kind)
            c = {"add": "+", "sub": "-", "mul": "*"}[op]
            return self.norm("(%s %s %s)" % (a, c, b), n.t)
        if op == "neg":
            return self.norm("(-%s)" % self.ex(n.a[0]), n.t)
        if op == "cat":
            e = "((int64_t)((uint64_t)%s << %d) | %s)" % (self.raw(n.a[0]), n.a[1].t.w, self.raw(n.a[1]))
            return self.norm(e, n.t) if n.t.kind == "signed" else e
        if op == "not":
            return "(1 ^ %s)" % self.ex(n.a[0])
        if op in ("and", "or", "xor"):
            return "(%s %s %s)" % (self.ex(n.a[0]), {"and": "&", "or": "|", "xor": "^"}[op], self.ex(n.a[1]))
        if op == "cmp":
            rel, l, r = n.a
            if l.t.kind == "bit":
                a, b = self.ex(l), self.ex(r)
            else:
                kind = _num_kind(l.t, r.t, n.line)
                a, b = self.num(l, kind), self.num(r, kind)
            return "(int64_t)(%s %s %s)" % (a, {"=": "==", "/=": "!="}.get(rel, rel), b)
        raise ValueError("vhdl_eval: cannot translate %r" % n)

    def assign(self, tgt, e, dst, ind):
        """statements storing expression e into target tgt of struct `dst`"""
        tt = tgt.t
        if tgt.op == "name" and tt.kind != "array":
            return ["%s%s->%s = %s;" % (ind, dst, self.cname(tgt.a[0]), self.ex(e, tt))]
        name = self.d.base_name(tgt)
        bt = self.d.sym[name]
        assert bt.kind == "array"
        if tgt.op == "index":
            return ["%s%s->%s[%d] = %s;" % (ind, dst, self.cname(name), tgt.a[1].a[0] - bt.lo, self.ex(e, bt.elem))]
        lo, hi = (tt.lo, tt.hi) if tgt.op != "name" else (bt.lo, bt.hi)
        if e.op == "others":
            inner = e.a[0]
            val = self.ex(inner, bt.elem)
            return ["%sfor (int k = %d; k <= %d; k++) %s->%s[k] = %s;" % (ind, lo - bt.lo, hi - bt.lo, dst, self.cname(name), val)]
        assert e.op == "slice" and e.a[0].op == "name" and e.t.kind == "array"
        src = e.a[0].a[0]
        st = self.d.sym[src]
        return ["%sfor (int k = 0; k < %d; k++) %s->%s[%d + k] = s->%s[%d + k];"
                % (ind, hi - lo + 1, dst, self.cname(name), lo - bt.lo, self.cname(src), e.a[1] - st.lo)]

    def seq(self, body, ind):
        out = []
        for st in body:
            if st.op == "if":
                first = True
                for c, b in st.a[0]:
                    out.append("%s%s (%s) {" % (ind, "if" if first else "} else if", self.ex(c)))
                    out += self.seq(b, ind + "    ")
                    first = False
                if st.a[1]:
                    out.append(ind + "} else {")
                    out += self.seq(st.a[1], ind + "    ")
                out.append(ind + "}")
            else:
                out += self.assign(st.a[0], st.a[1], "n", ind)
        return out

    def emit(self):
        d, p = self.d, self.p
        o = ["/* GENERATED by tools/vhdl_eval.py from %s - a build product derived from the reference's own" % d.path,
             " * source text, written under oracle/_ref/ only (git-ignored).  Do not edit, do not commit. */",
             "#include <stdint.h>", "#include <string.h>",
             "#define ZX(x, w) ((int64_t)((uint64_t)(x) & ((UINT64_C(1) << (w)) - 1)))",
             "#define SX(x, w) ((int64_t)((uint64_t)(x) << (64 - (w))) >> (64 - (w)))",
             "/* numeric_std resize of a signed value to fewer bits: sign bit kept, w-1 low bits kept */",
             "#define RESIZE_S(x, w) (ZX((x), (w) - 1) - ((int64_t)((x) < 0) << ((w) - 1)))",
             "typedef struct {"]
        for name, ty in d.sym.items():
            if name in d.consts:
                continue
            if ty.kind == "array":
                o.append("    int64_t %s[%d];" % (self.cname(name), ty.hi - ty.lo + 1))
            else:
                o.append("    int64_t %s;" % self.cname(name))
        o.append("} %s_t;" % p)
        o.append("void %s_settle(%s_t *s)\n{" % (p, p))
        for st in d.conc:
            tt = st.a[0].t
            e = self.ex(st.a[2], tt)
            for c, a in reversed(st.a[1]):
                e = "(%s) ? %s :\n        %s" % (self.ex(c), self.ex(a, tt), e)
            o.append("    s->%s = %s;" % (self.cname(st.a[0].a[0]), e))
        o.append("}")
        o.append("/* one rising clock edge: every process reads pre-edge values (s) and writes the next state (n) */")
        o.append("void %s_clock(%s_t *s)\n{\n    %s_settle(s);\n    %s_t nx = *s, *n = &nx;" % (p, p, p, p))
        for body in d.procs:
            o += self.seq(body, "    ")
        o.append("    *s = nx;\n    %s_settle(s);\n}" % p)
        o.append("size_t %s_sizeof(void) { return sizeof(%s_t); }" % (p, p))
        for name, direction, ty in d.ports:
            if direction == "in":
                o.append("void %s_set_%s(%s_t *s, int64_t x) { s->%s = %s; }" % (p, name, p, self.cname(name), "ZX(x, %d)" % ty.w))
            else:
                o.append("int64_t %s_get_%s(const %s_t *s) { return s->%s; }" % (p, name, p, self.cname(name)))
        o.insert(4, "#include <stddef.h>")
        return "\n".join(o) + "\n"


# ----------------------------------------------------------------------------------------------
# helpers for drivers
# ----------------------------------------------------------------------------------------------
def reset_instance(inst, clk_enable=1):
    """hold the asynchronous reset, release it with clk_enable set (UA3REO.bdf: reset = RX_N, clk_enable = RX)"""
    inst.v["clk_enable"] = 0
    inst.apply_reset()
    inst.v["clk_enable"] = clk_enable
    inst.settle()


def to_signed(raw, w):
    return raw - (1 << w) if raw >> (w - 1) else raw


REF_FPGA = "/root/reference/FPGA"
MODULES = ("rx_cic", "rx_ciccomp", "rx_hilb", "tx_cic", "tx_ciccomp")


def main(argv):
    if len(argv) >= 3 and argv[1] == "--emit-c":
        outdir = argv[2]
        os.makedirs(outdir, exist_ok=True)
        src = argv[3] if len(argv) > 3 else REF_FPGA
        for m in MODULES:
            d = Design(os.path.join(src, m + ".vhd"))
            with open(os.path.join(outdir, m + "_hdl.c"), "w") as f:
                f.write(d.emit_c())
            print("vhdl_eval: %s: %d signals, %d concurrent assignments, %d processes -> %s_hdl.c"
                  % (m, len(d.signals), len(d.conc), len(d.procs), m))
        return 0
    print(__doc__)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
