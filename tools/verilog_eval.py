#!/usr/bin/env python3
"""tools/verilog_eval.py - executes the reference's hand-written Verilog.  TEST INFRASTRUCTURE ONLY.

tools/vhdl_eval.py runs the five machine-generated VHDL filters; what it left restated were the small Verilog units around
them: `mixer.v` / `tx_mixer.v` / `tx_summator.v` (Quartus megafunction wrappers around lpm_mult / lpm_add_sub),
`rx_mixer_shift.v`, `nco_shift.v`, `data_delay.v`, `DAC_corrector.v` and the bus state machine `stm32_interface.v`.  This
file parses those sources WHERE THEY LIE under /root/reference/FPGA and translates them to C (one struct of signals, one
function per clock, one for the continuous assignments), with the IEEE 1364-2001 expression rules applied on every node:

  * self-determined width of every operand, context width = max over the context-determined operands and the target;
  * an expression is signed only if ALL its operands are (part selects, concatenations, comparisons and based literals
    without `s` are unsigned; plain decimal literals and `integer` are signed 32 bit), and that one signedness decides how
    every operand is extended to the context width;
  * comparison / logical / concatenation / index operands are self-determined contexts of their own;
  * blocking assignments act in statement order, non-blocking ones read the values from before the clock edge;
  * registers start from their declared initial value, otherwise 0 (what a Cyclone IV register powers up with; simulation X
    is not modelled);
  * an `inout` driven by `assign bus = oe ? value : 'bZ` becomes: value seen on the bus = oe ? value : <bus>__ext, where
    <bus>__ext is what the other side drives, and <bus>__oe is exposed.
  * lpm_mult / lpm_add_sub instances (Altera LPM, semantics from the published LPM standard: signed or unsigned product of
    widtha x widthb bits, the widthp MOST significant bits when narrower; two's complement sum with overflow flag; `lpm_pipeline`
    registers on `clock` gated by `clken`) are built in; their parameters come from the wrapper's `defparam`.

Not supported (none of the eight files needs it): case, functions/tasks, generate, negedge / level sensitivity, delays,
division, X/Z arithmetic, hierarchical names other than defparam.

  python tools/verilog_eval.py c <out.c> <file.v>[:param=value,...][@cname] ...     emit C for the listed modules

`Interp` executes the same modules in Python over the same typed tree - a second evaluation path that the tests compare with
the compiled translation step by step.
"""
import os
import re
import sys

REF = "/root/reference/FPGA"


class VError(Exception):
    pass


# ----------------------------------------------------------------------------------------------------------------------
# lexer
# ----------------------------------------------------------------------------------------------------------------------
TOKEN = re.compile(r"""
    (?P<based>\d*\s*'\s*[sS]?[bBdDhHoO]\s*[0-9a-fA-FxXzZ_?]+)
  | (?P<dec>\d[\d_]*)
  | (?P<id>[A-Za-z_][A-Za-z0-9_$]*)
  | (?P<str>"[^"]*")
  | (?P<op><<<|>>>|<=|>=|==|!=|&&|\|\||<<|>>|[-+*/%&|^~!<>=?:;,.()\[\]{}@\#])
  | (?P<ws>\s+)
""", re.X)


def lex(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"`\w+[^\n]*", "", text)                 # `timescale and friends
    pos, out = 0, []
    while pos < len(text):
        m = TOKEN.match(text, pos)
        if not m:
            raise VError("lex error at %r" % text[pos:pos + 20])
        pos = m.end()
        k = m.lastgroup
        if k != "ws":
            out.append((k, m.group(k)))
    out.append(("eof", ""))
    return out


def parse_based(tok):
    m = re.match(r"(\d*)\s*'\s*([sS]?)([bBdDhHoO])\s*([0-9a-fA-FxXzZ_?]+)", tok)
    width = int(m.group(1)) if m.group(1) else None
    signed = bool(m.group(2))
    base = {"b": 2, "d": 10, "h": 16, "o": 8}[m.group(3).lower()]
    digits = m.group(4).replace("_", "")
    if re.fullmatch(r"[zZ?]+", digits):
        return ("num", 0, width, signed, True)
    if re.search(r"[xXzZ?]", digits):
        raise VError("X/Z digits inside a number are not supported: %s" % tok)
    return ("num", int(digits, base), width, signed, False)


# ----------------------------------------------------------------------------------------------------------------------
# parser
# ----------------------------------------------------------------------------------------------------------------------
class Sig:
    def __init__(self, name):
        self.name, self.kind = name, None            # kind: input / output / inout / None (internal)
        self.is_reg = False
        self.integer = False
        self.signed = False
        self.rng = None                              # (msb_expr, lsb_expr) or None = 1 bit
        self.arr = None                              # (hi_expr, lo_expr) for memories
        self.init = None                             # expr
        self.width = 1                               # elaborated
        self.length = 0                              # elaborated: 0 = scalar


BINARY_PREC = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", "<=", ">", ">="), ("<<", ">>", "<<<", ">>>"),
               ("+", "-"), ("*", "/", "%")]


class Parser:
    def __init__(self, text):
        self.t = lex(text)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise VError("expected %r, found %r (token %d)" % (val, self.peek()[1], self.i))

    def ident(self):
        k, v = self.next()
        if k != "id":
            raise VError("identifier expected, found %r" % v)
        return v

    # -- expressions --
    def expr(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("tern", c, a, b)
        return c

    def binary(self, level):
        if level == len(BINARY_PREC):
            return self.unary()
        a = self.binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in BINARY_PREC[level]:
            op = self.next()[1]
            b = self.binary(level + 1)
            a = ("bin", op, a, b)
        return a

    def unary(self):
        k, v = self.peek()
        if k == "op" and v in ("!", "~", "-", "+"):
            self.next()
            return ("un", v, self.unary())
        return self.primary()

    def primary(self):
        k, v = self.next()
        if k == "based":
            return parse_based(v)
        if k == "dec":
            if self.peek()[0] == "based" and self.peek()[1].lstrip().startswith("'"):   # "8 'b0" split by the lexer
                b = parse_based(self.next()[1])
                return ("num", b[1], int(v), b[3], b[4])
            return ("num", int(v.replace("_", "")), None, True, False)
        if k == "str":
            return ("str", v[1:-1])
        if k == "op" and v == "(":
            e = self.expr()
            self.expect(")")
            return e
        if k == "op" and v == "{":
            items = [self.expr()]
            while self.accept(","):
                items.append(self.expr())
            self.expect("}")
            return ("cat", items)
        if k == "id":
            if self.accept("["):
                a = self.expr()
                if self.accept(":"):
                    b = self.expr()
                    self.expect("]")
                    return ("range", v, a, b)
                self.expect("]")
                if self.accept("["):                 # memory word then bit/part select: not needed here
                    raise VError("select of a memory word is not supported")
                return ("idx", v, a)
            return ("id", v)
        raise VError("unexpected token %r in expression" % v)

    # -- statements --
    def lvalue(self):
        e = self.primary()
        if e[0] not in ("id", "idx", "range"):
            raise VError("bad assignment target %r" % (e,))
        return e

    def statement(self):
        if self.accept("begin"):
            body = []
            while not self.accept("end"):
                body.append(self.statement())
            return ("block", body)
        if self.accept("if"):
            self.expect("(")
            c = self.expr()
            self.expect(")")
            a = self.statement()
            b = self.statement() if self.accept("else") else None
            return ("if", c, a, b)
        if self.accept("for"):
            self.expect("(")
            init = self.assignment()
            self.expect(";")
            cond = self.expr()
            self.expect(";")
            step = self.assignment()
            self.expect(")")
            return ("for", init, cond, step, self.statement())
        if self.accept(";"):
            return ("block", [])
        a = self.assignment()
        self.expect(";")
        return a

    def assignment(self):
        lv = self.lvalue()
        if self.accept("="):
            return ("assign", lv, self.expr(), False)
        if self.accept("<="):
            return ("assign", lv, self.expr(), True)
        raise VError("assignment expected after %r" % (lv,))

    # -- module --
    def module(self):
        while not self.accept("module"):
            if self.peek()[0] == "eof":
                raise VError("no module found")
            self.next()
        m = {"name": self.ident(), "ports": [], "sigs": {}, "order": [], "params": {}, "assigns": [], "always": [], "inst": [],
             "defparam": {}}
        if self.accept("("):
            while not self.accept(")"):
                m["ports"].append(self.ident())
                self.accept(",")
        self.expect(";")

        def sig(name):
            if name not in m["sigs"]:
                m["sigs"][name] = Sig(name)
                m["order"].append(name)
            return m["sigs"][name]

        while not self.accept("endmodule"):
            k, v = self.peek()
            if v in ("input", "output", "inout", "reg", "wire", "integer"):
                self.next()
                kind = v if v in ("input", "output", "inout") else None
                is_reg = v == "reg"
                integer = v == "integer"
                if kind and self.peek()[1] in ("reg", "wire"):
                    is_reg = self.next()[1] == "reg"
                signed = False
                if self.peek()[1] in ("signed", "unsigned"):
                    signed = self.next()[1] == "signed"
                rng = None
                if self.accept("["):
                    a = self.expr()
                    self.expect(":")
                    b = self.expr()
                    self.expect("]")
                    rng = (a, b)
                while True:
                    s = sig(self.ident())
                    if kind:
                        s.kind = kind
                    s.is_reg = s.is_reg or is_reg or integer
                    if integer:
                        s.integer, s.signed = True, True
                    else:
                        s.signed = s.signed or signed
                        if rng is not None:
                            s.rng = rng
                    if self.accept("["):
                        a = self.expr()
                        self.expect(":")
                        b = self.expr()
                        self.expect("]")
                        s.arr = (a, b)
                    if self.accept("="):
                        e = self.expr()
                        if v == "wire" or (kind and not s.is_reg):
                            m["assigns"].append((("id", s.name), e))
                        else:
                            s.init = e
                    if not self.accept(","):
                        break
                self.expect(";")
            elif v == "parameter":
                self.next()
                while True:
                    name = self.ident()
                    self.expect("=")
                    m["params"][name] = self.expr()
                    if not self.accept(","):
                        break
                self.expect(";")
            elif v == "assign":
                self.next()
                lv = self.lvalue()
                self.expect("=")
                m["assigns"].append((lv, self.expr()))
                self.expect(";")
            elif v == "always":
                self.next()
                self.expect("@")
                self.expect("(")
                self.expect("posedge")
                clk = self.ident()
                self.expect(")")
                m["always"].append((clk, self.statement()))
            elif v == "defparam":
                self.next()
                while True:
                    inst = self.ident()
                    self.expect(".")
                    p = self.ident()
                    self.expect("=")
                    m["defparam"].setdefault(inst, {})[p] = self.expr()
                    if not self.accept(","):
                        break
                self.expect(";")
            elif k == "id":                           # module instance: type name ( .port(expr), ... );
                mtype, iname = self.ident(), self.ident()
                self.expect("(")
                conns = {}
                while not self.accept(")"):
                    self.expect(".")
                    port = self.ident()
                    self.expect("(")
                    conns[port] = None if self.peek()[1] == ")" else self.expr()
                    self.expect(")")
                    self.accept(",")
                self.expect(";")
                m["inst"].append((mtype, iname, conns))
            else:
                raise VError("unexpected %r at module level of %s" % (v, m["name"]))
        return m


def parse_file(path):
    return Parser(open(path).read()).module()


# ----------------------------------------------------------------------------------------------------------------------
# elaboration + C emission
# ----------------------------------------------------------------------------------------------------------------------
def const_eval(e, params):
    k = e[0]
    if k == "num":
        return e[1]
    if k == "id":
        if e[1] not in params:
            raise VError("not a constant: %s" % e[1])
        return params[e[1]]
    if k == "un":
        a = const_eval(e[2], params)
        return {"-": -a, "+": a, "~": ~a, "!": int(not a)}[e[1]]
    if k == "bin":
        a, b = const_eval(e[2], params), const_eval(e[3], params)
        return {"+": a + b, "-": a - b, "*": a * b, "<<": a << b, ">>": a >> b}[e[1]]
    raise VError("not a constant expression: %r" % (e,))


def mask(w):
    return "0x%xULL" % ((1 << w) - 1)


class Emitter:
    """One module, elaborated with its parameter values, as C."""

    def __init__(self, mod, overrides=None, cname=None):
        self.m = mod
        self.cname = cname or mod["name"]
        self.params = {}
        for k, e in mod["params"].items():
            self.params[k] = const_eval(e, self.params)
        for k, v in (overrides or {}).items():
            if k not in self.params:
                raise VError("%s has no parameter %s" % (mod["name"], k))
            self.params[k] = int(v)
        self.sigs = mod["sigs"]
        for s in self.sigs.values():
            if s.integer:
                s.width = 32
            elif s.rng:
                msb, lsb = const_eval(s.rng[0], self.params), const_eval(s.rng[1], self.params)
                if lsb != 0 or msb < 0:
                    raise VError("%s: only [msb:0] vectors are supported" % s.name)
                s.width = msb + 1
            if s.width > 64:
                raise VError("%s is wider than 64 bits" % s.name)
            if s.arr:
                hi, lo = const_eval(s.arr[0], self.params), const_eval(s.arr[1], self.params)
                if lo != 0:
                    raise VError("%s: only [n:0] memories are supported" % s.name)
                s.length = hi + 1
        self.extra = []                              # (name, width, length) fields added by the translation
        self.nba = set()
        self.loopvars = set()

    # -- typing --
    def sig(self, name):
        if name not in self.sigs:
            raise VError("%s: undeclared signal %s" % (self.m["name"], name))
        return self.sigs[name]

    def size(self, e):
        k = e[0]
        if k == "num":
            return e[2] if e[2] else 32
        if k == "id":
            if e[1] in self.params:
                return 32
            s = self.sig(e[1])
            if s.length:
                raise VError("memory %s used without an index" % s.name)
            return s.width
        if k == "idx":
            s = self.sig(e[1])
            return s.width if s.length else 1
        if k == "range":
            return const_eval(e[2], self.params) - const_eval(e[3], self.params) + 1
        if k == "cat":
            return sum(self.size(x) for x in e[1])
        if k == "un":
            return 1 if e[1] == "!" else self.size(e[2])
        if k == "bin":
            if e[1] in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return 1
            if e[1] in ("<<", ">>", "<<<", ">>>"):
                return self.size(e[2])
            return max(self.size(e[2]), self.size(e[3]))
        if k == "tern":
            return max(self.size(e[2]), self.size(e[3]))
        raise VError("size of %r" % (e,))

    def signed(self, e):
        k = e[0]
        if k == "num":
            return e[3]
        if k == "id":
            return True if e[1] in self.params else self.sig(e[1]).signed
        if k == "idx":
            s = self.sig(e[1])
            return s.signed if s.length else False
        if k in ("range", "cat"):
            return False
        if k == "un":
            return False if e[1] == "!" else self.signed(e[2])
        if k == "bin":
            if e[1] in ("==", "!=", "<", "<=", ">", ">=", "&&", "||"):
                return False
            if e[1] in ("<<", ">>", "<<<", ">>>"):
                return self.signed(e[2])
            return self.signed(e[2]) and self.signed(e[3])
        if k == "tern":
            return self.signed(e[2]) and self.signed(e[3])
        raise VError("signedness of %r" % (e,))

    # -- expressions: the returned C expression is the W-bit result, zero in the bits above W --
    def ref(self, name):
        if name in self.loopvars:
            return "((uint64_t)(uint32_t)%s)" % name
        return "s->%s" % name

    def ext(self, c, w, W, S):
        """operand of self-determined width w into a (W, S) context"""
        if w > W:
            return "((%s) & %s)" % (c, mask(W))
        if S and w < W:
            return "(vl_sx(%s, %d) & %s)" % (c, w, mask(W))
        return c

    def self_ctx(self, e):
        """a self-determined operand: (C expression, width)"""
        w, s = self.size(e), self.signed(e)
        return self.emit(e, w, s), w

    def boolean(self, e):
        c, _ = self.self_ctx(e)
        return "((%s) != 0)" % c

    def emit(self, e, W, S):
        k = e[0]
        if k == "num":
            if e[4]:
                raise VError("high impedance is only supported as `assign bus = oe ? value : 'bZ`")
            w = e[2] if e[2] else 32
            return self.ext("0x%xULL" % (e[1] & ((1 << w) - 1)), w, W, S)
        if k == "id":
            if e[1] in self.params:
                return self.ext("0x%xULL" % (self.params[e[1]] & 0xFFFFFFFF), 32, W, S)
            return self.ext(self.ref(e[1]), self.size(e), W, S)
        if k == "idx":
            s = self.sig(e[1])
            i, _ = self.self_ctx(e[2])
            if s.length:
                return self.ext("s->%s[vl_index(%s, %d)]" % (s.name, i, s.length), s.width, W, S)
            return "((%s >> vl_index(%s, %d)) & 1ULL)" % (self.ref(s.name), i, s.width)
        if k == "range":
            s = self.sig(e[1])
            msb, lsb = const_eval(e[2], self.params), const_eval(e[3], self.params)
            if not (s.width > msb >= lsb >= 0) or s.length:
                raise VError("part select %s[%d:%d] out of range" % (s.name, msb, lsb))
            return "((%s >> %d) & %s)" % (self.ref(s.name), lsb, mask(msb - lsb + 1))          # unsigned: zero extension is the mask
        if k == "cat":
            parts, sh = [], 0
            for x in reversed(e[1]):
                if x[0] == "num" and x[2] is None:
                    raise VError("unsized literal in a concatenation")
                c, w = self.self_ctx(x)
                parts.append("((%s) << %d)" % (c, sh) if sh else "(%s)" % c)
                sh += w
            if sh > 64:
                raise VError("concatenation wider than 64 bits")
            return self.ext("(" + " | ".join(parts) + ")", sh, W, False)
        if k == "un":
            if e[1] == "!":
                return "((uint64_t)!%s)" % self.boolean(e[2])
            a = self.emit(e[2], W, S)
            if e[1] == "+":
                return a
            return "((%s(%s)) & %s)" % ("0ULL - " if e[1] == "-" else "~", a, mask(W))
        if k == "bin":
            op = e[1]
            if op in ("&&", "||"):
                return "((uint64_t)(%s %s %s))" % (self.boolean(e[2]), op, self.boolean(e[3]))
            if op in ("==", "!=", "<", "<=", ">", ">="):
                w = max(self.size(e[2]), self.size(e[3]))
                sg = self.signed(e[2]) and self.signed(e[3])
                a, b = self.emit(e[2], w, sg), self.emit(e[3], w, sg)
                if sg and op not in ("==", "!="):
                    return "((uint64_t)((int64_t)vl_sx(%s, %d) %s (int64_t)vl_sx(%s, %d)))" % (a, w, op, b, w)
                return "((uint64_t)((%s) %s (%s)))" % (a, op, b)
            if op in ("<<", ">>", "<<<", ">>>"):
                a = self.emit(e[2], W, S)
                n, _ = self.self_ctx(e[3])
                if op in ("<<", "<<<"):
                    return "(vl_shl(%s, %s) & %s)" % (a, n, mask(W))
                if op == ">>>" and S:
                    return "(vl_sar(%s, %d, %s) & %s)" % (a, W, n, mask(W))
                return "vl_shr(%s, %s)" % (a, n)
            if op in ("/", "%"):
                raise VError("division is not supported")
            a, b = self.emit(e[2], W, S), self.emit(e[3], W, S)
            return "(((%s) %s (%s)) & %s)" % (a, op, b, mask(W))
        if k == "tern":
            return "(%s ? (%s) : (%s))" % (self.boolean(e[1]), self.emit(e[2], W, S), self.emit(e[3], W, S))
        raise VError("cannot emit %r" % (e,))

    # -- assignments --
    def target(self, lv, nba):
        """(C template with %s for the new value, width)"""
        s = self.sig(lv[1]) if lv[1] not in self.loopvars else None
        if s is None:
            if lv[0] != "id":
                raise VError("select of loop variable")
            return "%s = (int32_t)(%%s);" % lv[1], 32
        base = ("n_%s" % s.name) if nba else ("s->%s" % s.name)
        if nba:
            self.nba.add(s.name)
        if lv[0] == "id":
            if s.length:
                raise VError("assignment to a whole memory")
            return "%s = (%%s) & %s;" % (base, mask(s.width)), s.width
        if lv[0] == "idx":
            i, _ = self.self_ctx(lv[2])
            if s.length:
                return "%s[vl_index(%s, %d)] = (%%s) & %s;" % (base, i, s.length, mask(s.width)), s.width
            return "{ const unsigned i_ = vl_index(%s, %d); %s = (%s & ~(1ULL << i_)) | (((%%s) & 1ULL) << i_); }" % \
                   (i, s.width, base, base), 1
        msb, lsb = const_eval(lv[2], self.params), const_eval(lv[3], self.params)
        if not (s.width > msb >= lsb >= 0) or s.length:
            raise VError("part select %s[%d:%d] out of range" % (s.name, msb, lsb))
        n = msb - lsb + 1
        return "%s = (%s & ~(%s << %d)) | (((%s) & %s) << %d);" % (base, base, mask(n), lsb, "%s", mask(n), lsb), n

    def assign(self, lv, rhs, nba):
        tmpl, wl = self.target(lv, nba)
        W = max(wl, self.size(rhs))
        return tmpl % self.emit(rhs, W, self.signed(rhs))

    def stmt(self, st, ind):
        pad = "    " * ind
        k = st[0]
        if k == "block":
            return "".join(self.stmt(x, ind) for x in st[1])
        if k == "assign":
            return pad + self.assign(st[1], st[2], st[3]) + "\n"
        if k == "if":
            out = pad + "if (%s) {\n" % self.boolean(st[1]) + self.stmt(st[2], ind + 1) + pad + "}"
            if st[3] is not None:
                out += " else {\n" + self.stmt(st[3], ind + 1) + pad + "}"
            return out + "\n"
        if k == "for":
            init, cond, step, body = st[1:]
            var = init[1][1]
            if init[1][0] != "id" or not self.sig(var).integer:
                raise VError("for loop variable must be an integer")
            self.loopvars.add(var)
            head = pad + "for (%s %s; %s) {\n" % (self.assign(init[1], init[2], False), self.boolean(cond),
                                                  self.assign(step[1], step[2], False).rstrip(";"))
            out = head + self.stmt(body, ind + 1) + pad + "}\n"
            self.loopvars.discard(var)
            return out
        raise VError("statement %r" % (st,))

    # -- LPM instances --
    def lpm(self, mtype, iname, conns):
        p = {k: v for k, v in self.m["defparam"].get(iname, {}).items()}

        def par(name, default=None):
            if name not in p:
                if default is None:
                    raise VError("%s: lpm parameter %s missing" % (iname, name))
                return default
            e = p[name]
            return e[1] if e[0] in ("num", "str") else const_eval(e, self.params)

        for port in ("aclr", "sclr", "sum", "cin"):
            e = conns.get(port)
            if e is not None and not (e[0] == "num" and e[1] == 0):
                raise VError("%s.%s must be constant 0 or open" % (iname, port))
        for port in ("add_sub", "cout"):
            if conns.get(port) is not None:
                raise VError("%s.%s is not supported" % (iname, port))
        pipe = int(par("lpm_pipeline", 0))
        sgn = str(par("lpm_representation", "UNSIGNED")).upper() == "SIGNED"
        clock = conns.get("clock")
        if pipe and (clock is None or clock[0] != "id"):
            raise VError("%s: pipelined lpm instance without a clock" % iname)
        en = self.boolean(conns["clken"]) if conns.get("clken") is not None else "1"

        def operand(port, w):
            e = conns[port]
            c, we = self.self_ctx(e)
            if we != w:
                raise VError("%s.%s: connected width %d, lpm width %d" % (iname, port, we, w))
            return "(int64_t)vl_sx(%s, %d)" % (c, w) if sgn else "(int64_t)(%s)" % c

        comb = []
        outs = []                                                     # (wire lvalue, pipe field base, width)
        if mtype == "lpm_mult":
            wa, wb, wp = int(par("lpm_widtha")), int(par("lpm_widthb")), int(par("lpm_widthp"))
            if wa + wb > 63:
                raise VError("lpm_mult wider than 63 bits")
            prod = "((uint64_t)(%s * %s))" % (operand("dataa", wa), operand("datab", wb))
            drop = max(0, wa + wb - wp)                                # LPM: the widthp most significant bits
            comb.append(("result", "(((%s & %s) >> %d) & %s)" % (prod, mask(wa + wb), drop, mask(wp)), wp))
        elif mtype == "lpm_add_sub":
            w = int(par("lpm_width"))
            direction = str(par("lpm_direction", "ADD")).upper()
            if direction not in ("ADD", "SUB"):
                raise VError("lpm_add_sub direction %s" % direction)
            op = "+" if direction == "ADD" else "-"
            full = "(%s %s %s)" % (operand("dataa", w), op, operand("datab", w))
            comb.append(("result", "((uint64_t)%s & %s)" % (full, mask(w)), w))
            if sgn:
                ov = "((uint64_t)((int64_t)vl_sx((uint64_t)%s & %s, %d) != %s))" % (full, mask(w), w, full)
            else:
                ov = "((uint64_t)(((uint64_t)%s >> %d) & 1ULL))" % (full, w)
            comb.append(("overflow", ov, 1))
        else:
            raise VError("module %s (instance %s) is not known" % (mtype, iname))
        clocked = []
        for port, cexpr, w in comb:
            wire = conns.get(port)
            if wire is None:
                continue
            if wire[0] != "id":
                raise VError("%s.%s must connect to a plain wire" % (iname, port))
            if pipe:
                fld = "%s__%s" % (iname, port)
                self.extra.append((fld, w, pipe))
                body = "".join("        s->%s[%d] = s->%s[%d];\n" % (fld, j, fld, j - 1) for j in range(pipe - 1, 0, -1))
                body += "        s->%s[0] = %s;\n" % (fld, cexpr)
                clocked.append(body)
                outs.append((wire, "s->%s[%d]" % (fld, pipe - 1)))
            else:
                outs.append((wire, cexpr))
        return (clock[1] if pipe else None), en, clocked, outs

    # -- module --
    def emit_module(self):
        n = self.cname
        clocks = {}                                   # clock name -> C body
        settle_items = []                             # (target name, C statement, names read)
        for mtype, iname, conns in self.m["inst"]:
            clock, en, clocked, outs = self.lpm(mtype, iname, conns)
            if clock:
                clocks.setdefault(clock, "")
                clocks[clock] += "    if (%s) {\n%s    }\n" % (en, "".join(clocked))
            for wire, cexpr in outs:
                s = self.sig(wire[1])
                settle_items.append((s.name, "    s->%s = (%s) & %s;\n" % (s.name, cexpr, mask(s.width)), cexpr))
        for clk, body in self.m["always"]:
            self.nba = set()
            code = self.stmt(body, 1)
            pre = post = ""
            for name in sorted(self.nba):
                s = self.sig(name)
                if s.length:
                    pre += "    uint64_t n_%s[%d]; memcpy(n_%s, s->%s, sizeof n_%s);\n" % (name, s.length, name, name, name)
                    post += "    memcpy(s->%s, n_%s, sizeof n_%s);\n" % (name, name, name)
                else:
                    pre += "    uint64_t n_%s = s->%s;\n" % (name, name)
                    post += "    s->%s = n_%s;\n" % (name, name)
            loopdecl = "".join("    int32_t %s = 0; (void)%s;\n" % (v, v) for v in sorted(self._loopvars_in(body)))
            clocks[clk] = clocks.get(clk, "") + loopdecl + pre + code + post
        for lv, e in self.m["assigns"]:
            s = self.sig(lv[1])
            if e[0] == "tern" and e[3][0] == "num" and e[3][4]:      # assign bus = oe ? value : 'bZ
                if s.kind != "inout" or lv[0] != "id":
                    raise VError("tri-state assign to something that is not an inout")
                self.extra.append((s.name + "__ext", s.width, 0))
                self.extra.append((s.name + "__oe", 1, 0))
                oe = self.boolean(e[1])
                val = self.emit(e[2], max(s.width, self.size(e[2])), self.signed(e[2]))
                c = "    s->%s__oe = %s;\n    s->%s = (s->%s__oe ? (%s) : s->%s__ext) & %s;\n" % (s.name, oe, s.name, s.name, val, s.name, mask(s.width))
                settle_items.append((s.name, c, c))
            else:
                c = "    " + self.assign(lv, e, False) + "\n"
                settle_items.append((s.name, c, c))
        # continuous assignments in dependency order (a wire is computed after the wires it reads)
        done, ordered, pending = set(), [], list(settle_items)
        targets = {t for t, _, _ in settle_items}
        while pending:
            progress = False
            for item in list(pending):
                reads = {t for t in targets if t != item[0] and re.search(r"s->%s\b" % re.escape(t), item[2])}
                if reads <= done:
                    ordered.append(item)
                    pending.remove(item)
                    if not any(p[0] == item[0] for p in pending):
                        done.add(item[0])
                    progress = True
            if not progress:
                raise VError("%s: combinational loop in the continuous assignments" % n)
        fields = [(s.name, s.width, s.length, s.signed, s.kind or "") for s in (self.sigs[k] for k in self.m["order"])]
        fields += [(nm, w, ln, False, "") for nm, w, ln in self.extra]
        out = ["/* module %s (%s) - generated by tools/verilog_eval.py from the reference's Verilog; do not edit */\n" %
               (self.m["name"], ", ".join("%s=%d" % kv for kv in sorted(self.params.items())) or "no parameters")]
        out.append("typedef struct {\n" + "".join("    uint64_t %s%s;\n" % (f[0], "[%d]" % f[2] if f[2] else "") for f in fields) + "} %s_t;\n" % n)
        out.append("void %s_settle(%s_t *s) {\n    (void)s;\n%s}\n" % (n, n, "".join(c for _, c, _ in ordered)))
        init = "    memset(s, 0, sizeof *s);\n"
        for s in self.sigs.values():
            if s.init is not None:
                v = const_eval(s.init, self.params) & ((1 << s.width) - 1)
                init += "    s->%s = 0x%xULL;\n" % (s.name, v)
        out.append("void %s_init(%s_t *s) {\n%s    %s_settle(s);\n}\n" % (n, n, init, n))
        for clk, body in clocks.items():
            out.append("void %s_posedge_%s(%s_t *s) {\n%s    %s_settle(s);\n}\n" % (n, clk, n, body, n))
        out.append("static const vl_field %s_fields[] = {\n%s};\n" % (n, "".join(
            '    {"%s", offsetof(%s_t, %s), %d, %d, %d, "%s"},\n' % (f[0], n, f[0], f[1], f[2], int(f[3]), f[4]) for f in fields)))
        clk_tab = "".join('    {"%s", (void (*)(void *))%s_posedge_%s},\n' % (c, n, c) for c in clocks)
        out.append("static const vl_clock %s_clocks[] = {\n%s    {0, 0}\n};\n" % (n, clk_tab))
        self.entry = '    {"%s", sizeof(%s_t), (void (*)(void *))%s_init, (void (*)(void *))%s_settle, %s_fields, %d, %s_clocks},\n' % \
                     (n, n, n, n, n, len(fields), n)
        return "".join(out)

    def _loopvars_in(self, st):
        if st[0] == "block":
            r = set()
            for x in st[1]:
                r |= self._loopvars_in(x)
            return r
        if st[0] == "if":
            return self._loopvars_in(st[2]) | (self._loopvars_in(st[3]) if st[3] else set())
        if st[0] == "for":
            return {st[1][1][1]} | self._loopvars_in(st[4])
        return set()


class Interp:
    """The same module executed by a Python interpreter over the same typed tree - an independent evaluation path for the
    C emitter above (tests/test_verilog_pin.py drives both with the same stimulus).  Values are Python integers, reduced to
    the context width after every operation; LPM instances are evaluated from their parameters directly."""

    def __init__(self, mod, overrides=None):
        self.e = Emitter(mod, overrides)              # elaboration: parameters, widths, typing rules
        self.m = mod
        self.v = {}
        for s in self.e.sigs.values():
            self.v[s.name] = [0] * s.length if s.length else 0
            if s.init is not None:
                self.v[s.name] = const_eval(s.init, self.e.params) & ((1 << s.width) - 1)
        self.ext_in, self.pipes = {}, {}
        for lv, ex in mod["assigns"]:
            if ex[0] == "tern" and ex[3][0] == "num" and ex[3][4]:
                self.ext_in[lv[1]] = 0
        for mtype, iname, conns in mod["inst"]:
            n = int(self._par(iname, "lpm_pipeline", 0))
            self.pipes[iname] = {"result": [0] * n, "overflow": [0] * n}
        self.settle()

    # -- values --
    @staticmethod
    def _sx(v, w):
        v &= (1 << w) - 1
        return v - (1 << w) if v >> (w - 1) else v

    def _ext(self, v, w, W, S):
        v &= (1 << w) - 1
        if S and w < W:
            v = self._sx(v, w)
        return v & ((1 << W) - 1)

    def _self(self, e):
        w, s = self.e.size(e), self.e.signed(e)
        return self.ev(e, w, s), w

    def ev(self, e, W, S):
        k, M = e[0], (1 << W) - 1
        E = self.e
        if k == "num":
            return self._ext(e[1], e[2] if e[2] else 32, W, S)
        if k == "id":
            if e[1] in E.params:
                return self._ext(E.params[e[1]], 32, W, S)
            return self._ext(self.v[e[1]], E.size(e), W, S)
        if k == "idx":
            s = E.sig(e[1])
            i, _ = self._self(e[2])
            if s.length:
                return self._ext(self.v[s.name][i if i < s.length else 0], s.width, W, S)
            return (self.v[s.name] >> (i if i < s.width else 0)) & 1
        if k == "range":
            msb, lsb = const_eval(e[2], E.params), const_eval(e[3], E.params)
            return (self.v[e[1]] >> lsb) & ((1 << (msb - lsb + 1)) - 1)
        if k == "cat":
            acc = 0
            for x in e[1]:
                c, w = self._self(x)
                acc = (acc << w) | c
            return acc & M
        if k == "un":
            if e[1] == "!":
                return int(self._self(e[2])[0] == 0)
            a = self.ev(e[2], W, S)
            return a if e[1] == "+" else ((-a) & M if e[1] == "-" else (~a) & M)
        if k == "bin":
            op = e[1]
            if op == "&&":
                return int(self._self(e[2])[0] != 0 and self._self(e[3])[0] != 0)
            if op == "||":
                return int(self._self(e[2])[0] != 0 or self._self(e[3])[0] != 0)
            if op in ("==", "!=", "<", "<=", ">", ">="):
                w = max(E.size(e[2]), E.size(e[3]))
                sg = E.signed(e[2]) and E.signed(e[3])
                a, b = self.ev(e[2], w, sg), self.ev(e[3], w, sg)
                if sg:
                    a, b = self._sx(a, w), self._sx(b, w)
                return int({"==": a == b, "!=": a != b, "<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b}[op])
            if op in ("<<", "<<<", ">>", ">>>"):
                a, n = self.ev(e[2], W, S), self._self(e[3])[0]
                if op in ("<<", "<<<"):
                    return (a << n) & M if n < 64 else 0
                if op == ">>>" and S:
                    return (self._sx(a, W) >> min(n, 63)) & M
                return a >> n if n < 64 else 0
            a, b = self.ev(e[2], W, S), self.ev(e[3], W, S)
            return {"+": a + b, "-": a - b, "*": a * b, "&": a & b, "|": a | b, "^": a ^ b}[op] & M
        if k == "tern":
            return self.ev(e[2], W, S) if self._self(e[1])[0] != 0 else self.ev(e[3], W, S)
        raise VError("cannot interpret %r" % (e,))

    # -- statements --
    def _store(self, lv, val, target):
        s = self.e.sig(lv[1])
        if lv[0] == "id":
            target[s.name] = val & ((1 << s.width) - 1)
        elif lv[0] == "idx":
            i, _ = self._self(lv[2])
            if s.length:
                arr = list(target[s.name])
                arr[i if i < s.length else 0] = val & ((1 << s.width) - 1)
                target[s.name] = arr
            else:
                i = i if i < s.width else 0
                target[s.name] = (target[s.name] & ~(1 << i)) | ((val & 1) << i)
        else:
            msb, lsb = const_eval(lv[2], self.e.params), const_eval(lv[3], self.e.params)
            m = ((1 << (msb - lsb + 1)) - 1)
            target[s.name] = (target[s.name] & ~(m << lsb)) | ((val & m) << lsb)

    def _lwidth(self, lv):
        s = self.e.sig(lv[1])
        if lv[0] == "id":
            return s.width
        if lv[0] == "idx":
            return s.width if s.length else 1
        return const_eval(lv[2], self.e.params) - const_eval(lv[3], self.e.params) + 1

    def _assign(self, lv, rhs, target):
        W = max(self._lwidth(lv), self.e.size(rhs))
        self._store(lv, self.ev(rhs, W, self.e.signed(rhs)), target)

    def _stmt(self, st, nba):
        k = st[0]
        if k == "block":
            for x in st[1]:
                self._stmt(x, nba)
        elif k == "assign":
            self._assign(st[1], st[2], nba if st[3] else self.v)
        elif k == "if":
            if self._self(st[1])[0] != 0:
                self._stmt(st[2], nba)
            elif st[3] is not None:
                self._stmt(st[3], nba)
        elif k == "for":
            self._assign(st[1][1], st[1][2], self.v)
            while self._self(st[2])[0] != 0:
                self._stmt(st[4], nba)
                self._assign(st[3][1], st[3][2], self.v)
        else:
            raise VError("statement %r" % (st,))

    # -- LPM --
    def _par(self, iname, name, default=None):
        p = self.m["defparam"].get(iname, {})
        if name not in p:
            return default
        ex = p[name]
        return ex[1] if ex[0] in ("num", "str") else const_eval(ex, self.e.params)

    def _lpm_now(self, mtype, iname, conns):
        sgn = str(self._par(iname, "lpm_representation", "UNSIGNED")).upper() == "SIGNED"

        def operand(port, w):
            c, _ = self._self(conns[port])
            return self._sx(c, w) if sgn else c
        if mtype == "lpm_mult":
            wa, wb, wp = (int(self._par(iname, k)) for k in ("lpm_widtha", "lpm_widthb", "lpm_widthp"))
            prod = (operand("dataa", wa) * operand("datab", wb)) & ((1 << (wa + wb)) - 1)
            return {"result": (prod >> max(0, wa + wb - wp)) & ((1 << wp) - 1)}
        w = int(self._par(iname, "lpm_width"))
        a, b = operand("dataa", w), operand("datab", w)
        full = a + b if str(self._par(iname, "lpm_direction", "ADD")).upper() == "ADD" else a - b
        res = full & ((1 << w) - 1)
        ov = int(self._sx(res, w) != full) if sgn else (full >> w) & 1
        return {"result": res, "overflow": ov}

    # -- module --
    def settle(self):
        for _ in range(4):                            # a few passes reach the fixed point of these small netlists
            for mtype, iname, conns in self.m["inst"]:
                now = self._lpm_now(mtype, iname, conns)
                for port, val in now.items():
                    wire = conns.get(port)
                    if wire is not None:
                        pipe = self.pipes[iname][port]
                        self.v[wire[1]] = (pipe[-1] if pipe else val) & ((1 << self.e.sig(wire[1]).width) - 1)
            for lv, ex in self.m["assigns"]:
                if lv[1] in self.ext_in and ex[0] == "tern" and ex[3][0] == "num" and ex[3][4]:
                    s = self.e.sig(lv[1])
                    oe = self._self(ex[1])[0] != 0
                    self.v[lv[1] + "__oe"] = int(oe)
                    val = self.ev(ex[2], max(s.width, self.e.size(ex[2])), self.e.signed(ex[2])) if oe else self.ext_in[lv[1]]
                    self.v[lv[1]] = val & ((1 << s.width) - 1)
                else:
                    self._assign(lv, ex, self.v)

    def clock(self, clk):
        hit = False
        for mtype, iname, conns in self.m["inst"]:
            ck = conns.get("clock")
            if ck is not None and ck[1] == clk and int(self._par(iname, "lpm_pipeline", 0)):
                hit = True
                if conns.get("clken") is None or self._self(conns["clken"])[0] != 0:
                    now = self._lpm_now(mtype, iname, conns)
                    for port, val in now.items():
                        pipe = self.pipes[iname][port]
                        pipe.insert(0, val)
                        pipe.pop()
        for c, body in self.m["always"]:
            if c == clk:
                hit = True
                nba = {}
                shadow = _NbaView(self.v, nba)
                self._stmt(body, shadow)
                for name, val in nba.items():
                    self.v[name] = val
        if not hit:
            raise KeyError("no process on posedge %s" % clk)
        self.settle()


class _NbaView(dict):
    """target of non-blocking assignments: reads of a not yet written name see the CURRENT value, writes go to the side"""

    def __init__(self, cur, side):
        super().__init__()
        self.cur, self.side = cur, side

    def __getitem__(self, k):
        return self.side[k] if k in self.side else self.cur[k]

    def __setitem__(self, k, v):
        self.side[k] = v


PRELUDE = r"""/* generated by tools/verilog_eval.py - the reference's Verilog translated to C.  TEST INFRASTRUCTURE ONLY. */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
static inline uint64_t vl_sx(uint64_t v, int w) { if (w >= 64) return v; const uint64_t m = 1ULL << (w - 1); v &= (m << 1) - 1; return (v ^ m) - m; }
static inline uint64_t vl_shl(uint64_t v, uint64_t n) { return n >= 64 ? 0 : v << n; }
static inline uint64_t vl_shr(uint64_t v, uint64_t n) { return n >= 64 ? 0 : v >> n; }
static inline uint64_t vl_sar(uint64_t v, int w, uint64_t n) { const int64_t x = (int64_t)vl_sx(v, w); return (uint64_t)(x >> (n >= 63 ? 63 : n)); }
static inline unsigned vl_index(uint64_t i, unsigned n) { return i < n ? (unsigned)i : 0u; }   /* out-of-range selects read X in simulation; none occurs */
typedef struct { const char *name; size_t offset; int width, length, is_signed; const char *dir; } vl_field;
typedef struct { const char *name; void (*fn)(void *); } vl_clock;
typedef struct { const char *name; size_t size; void (*init)(void *); void (*settle)(void *); const vl_field *fields; int n_fields; const vl_clock *clocks; } vl_module;
"""

POSTLUDE = r"""
const vl_module *vl_find(const char *name) { for (int i = 0; vl_modules[i].name; i++) if (!strcmp(vl_modules[i].name, name)) return &vl_modules[i]; return 0; }
size_t vl_size(const vl_module *m) { return m->size; }
void vl_init(const vl_module *m, void *s) { m->init(s); }
void vl_settle(const vl_module *m, void *s) { m->settle(s); }
int vl_n_fields(const vl_module *m) { return m->n_fields; }
const vl_field *vl_field_at(const vl_module *m, int i) { return &m->fields[i]; }
const char *vl_field_name(const vl_module *m, int i) { return m->fields[i].name; }
size_t vl_field_offset(const vl_module *m, int i) { return m->fields[i].offset; }
int vl_field_width(const vl_module *m, int i) { return m->fields[i].width; }
int vl_field_length(const vl_module *m, int i) { return m->fields[i].length; }
int vl_field_signed(const vl_module *m, int i) { return m->fields[i].is_signed; }
/* n steps of one module: inputs in[j * n + t] -> scalar field in_idx[j], continuous assignments, one rising edge of `clock` (none
 * if clock is null: a combinational module), then out[k * n + t] = raw bits of scalar field out_idx[k] */
int vl_run(const vl_module *m, void *s, const char *clock, size_t n, int n_in, const int *in_idx, const uint64_t *in,
           int n_out, const int *out_idx, uint64_t *out) {
    void (*edge)(void *) = 0;
    if (clock) {
        for (int i = 0; m->clocks[i].name; i++) if (!strcmp(m->clocks[i].name, clock)) edge = m->clocks[i].fn;
        if (!edge) return -1;
    }
    for (int j = 0; j < n_in; j++) if (in_idx[j] < 0 || in_idx[j] >= m->n_fields || m->fields[in_idx[j]].length) return -2;
    for (int k = 0; k < n_out; k++) if (out_idx[k] < 0 || out_idx[k] >= m->n_fields || m->fields[out_idx[k]].length) return -2;
    for (size_t t = 0; t < n; t++) {
        for (int j = 0; j < n_in; j++) {
            const vl_field *f = &m->fields[in_idx[j]];
            *(uint64_t *)((char *)s + f->offset) = in[(size_t)j * n + t] & (f->width >= 64 ? ~0ULL : ((1ULL << f->width) - 1ULL));
        }
        m->settle(s);
        if (edge) edge(s);
        for (int k = 0; k < n_out; k++) out[(size_t)k * n + t] = *(const uint64_t *)((const char *)s + m->fields[out_idx[k]].offset);
    }
    return 0;
}
int vl_clock_edge(const vl_module *m, void *s, const char *clock) {
    for (int i = 0; m->clocks[i].name; i++) if (!strcmp(m->clocks[i].name, clock)) { m->clocks[i].fn(s); return 0; }
    return -1;
}
"""


def translate(specs):
    """specs: list of (path, {param: value}, cname or None) -> C source text"""
    out, entries = [PRELUDE], []
    for path, overrides, cname in specs:
        em = Emitter(parse_file(path), overrides, cname)
        out.append(em.emit_module())
        entries.append(em.entry)
    out.append("static const vl_module vl_modules[] = {\n" + "".join(entries) + "    {0, 0, 0, 0, 0, 0, 0}\n};\n")
    out.append(POSTLUDE)
    return "\n".join(out)


def parse_spec(arg):
    cname = None
    if "@" in arg:
        arg, cname = arg.rsplit("@", 1)
    overrides = {}
    if ":" in arg:
        arg, ps = arg.split(":", 1)
        overrides = {k: int(v) for k, v in (kv.split("=") for kv in ps.split(","))}
    path = arg if os.path.isabs(arg) else os.path.join(REF, arg)
    return path, overrides, cname


def main(argv):
    if len(argv) >= 4 and argv[1] == "c":
        src = translate([parse_spec(a) for a in argv[3:]])
        with open(argv[2], "w") as f:
            f.write(src)
        return 0
    sys.stderr.write(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv))
