"""TEST INFRASTRUCTURE ONLY -- a small interpreter for the GLSL ES 1.00 subset the reference's
shaders use, so that the reference's OWN SHADER SOURCE TEXT can be executed here.

kcdodd/fusion-sim runs all of its arithmetic in WebGL-1 shaders embedded as string arrays in
public/javascripts/empic.js; there is no JS engine, no GL and no browser in this image, so the
reference cannot run as a whole.  What CAN run is each shader: this module
parses the GLSL text (extracted from the reference tree by tests/golden/make_reference_vectors.py,
never copied into this repository) and evaluates it over a batch of fragments with NumPy.  The
outputs are committed as golden vectors (tests/golden/reference_glsl_*.npz) and the CPU oracle is
held to them bit for bit (tests/test_reference_glsl.py): that pins the oracle's restatement of
every per-fragment formula -- operands, signs, constants, operation order -- to the reference's
source, instead of to a reading of it.

Semantics where GLSL ES 1.00 leaves freedom (the same choices the oracle documents):
  * arithmetic is IEEE round-to-nearest in the chosen dtype (float64 or float32), evaluated
    strictly left to right as the expression is written, no fused multiply-add;
  * dot(a,b) = ((a.x*b.x + a.y*b.y) + a.z*b.z) + ..., length(v) = sqrt(dot(v,v)),
    cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x);
  * texture2D samples NEAREST with CLAMP_TO_EDGE (utilities.js:528-531): texel
    clamp(floor(u*W), 0, W-1); a NaN coordinate samples texel 0;
  * both branches of ?: are evaluated and selected per fragment; if/else and for run under a
    per-fragment execution mask.

Supported: precision/uniform/attribute/varying declarations, float/vec2/vec3/vec4 locals, swizzles
(xyzw/rgba), constructors, + - * / unary -, comparisons, || &&, ?:, = += -= *= /=, ++ in for
headers, if/else, for, and the builtins abs cos cross dot floor length max min mod sign sqrt
texture2D.
"""
from __future__ import annotations

import math
import re

import numpy as np

_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|\d+)"
                    r"|([A-Za-z_][A-Za-z0-9_]*)|(\+\+|--|\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||[-+*/<>=!?:;,.(){}\[\]]))")
_TYPES = {"float": 1, "vec2": 2, "vec3": 3, "vec4": 4, "int": 1}
_SWZ = {c: i for i, c in enumerate("xyzw")}
_SWZ.update({c: i for i, c in enumerate("rgba")})
_SWZ.update({c: i for i, c in enumerate("stpq")})


def tokenize(src: str):
    out, pos = [], 0
    src = src.rstrip()
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"GLSL: cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            out.append(("num", m.group(2)))
        elif m.group(3) is not None:
            out.append(("id", m.group(3)))
        else:
            out.append(("op", m.group(4)))
    return out


class Parser:
    """Recursive descent -> nested tuples."""

    def __init__(self, src: str):
        self.t = tokenize(src)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"GLSL: expected {val!r}, got {self.peek()!r}")

    # -- translation unit -------------------------------------------------------------------
    def parse(self):
        decls, main = [], None
        while self.peek()[0] != "eof":
            kind, val = self.peek()
            if val == "precision":
                while self.next()[1] != ";":
                    pass
            elif val in ("uniform", "attribute", "varying", "const"):
                self.next()
                if self.peek()[1] in ("highp", "mediump", "lowp"):
                    self.next()
                typ = self.next()[1]
                name = self.next()[1]
                init = None
                if self.accept("="):
                    init = self.expr()
                self.expect(";")
                decls.append((val, typ, name, init))
            elif val == "void":
                self.next()
                if self.next()[1] != "main":
                    raise SyntaxError("GLSL: only main() is supported")
                self.expect("(")
                self.accept("void")
                self.expect(")")
                main = self.block()
            else:
                raise SyntaxError(f"GLSL: unexpected {self.peek()!r} at global scope")
        return decls, main

    def block(self):
        self.expect("{")
        body = []
        while not self.accept("}"):
            body.append(self.statement())
        return ("block", body)

    def statement(self):
        kind, val = self.peek()
        if val == "{":
            return self.block()
        if val == "if":
            self.next()
            self.expect("(")
            cond = self.expr()
            self.expect(")")
            then = self.statement()
            other = self.statement() if self.accept("else") else None
            return ("if", cond, then, other)
        if val == "for":
            self.next()
            self.expect("(")
            init = self.simple()
            self.expect(";")
            cond = self.expr()
            self.expect(";")
            step = self.simple()
            self.expect(")")
            return ("for", init, cond, step, self.statement())
        st = self.simple()
        self.expect(";")
        return st

    def simple(self):
        kind, val = self.peek()
        if kind == "id" and val in _TYPES and self.peek(1)[0] == "id":
            self.next()
            name = self.next()[1]
            init = self.expr() if self.accept("=") else None
            return ("decl", val, name, init)
        target = self.postfix()
        kind, val = self.next()
        if val in ("=", "+=", "-=", "*=", "/="):
            return ("assign", val, target, self.expr())
        if val in ("++", "--"):
            return ("assign", "+=" if val == "++" else "-=", target, ("num", "1.0"))
        raise SyntaxError(f"GLSL: expected an assignment, got {val!r}")

    # -- expressions (precedence climbing) ----------------------------------------------------
    def expr(self):
        c = self.lor()
        if self.accept("?"):
            a = self.expr()
            self.expect(":")
            b = self.expr()
            return ("sel", c, a, b)
        return c

    def _binary(self, sub, ops):
        left = sub()
        while self.peek()[0] == "op" and self.peek()[1] in ops:
            op = self.next()[1]
            left = ("bin", op, left, sub())
        return left

    def lor(self):
        return self._binary(self.land, ("||",))

    def land(self):
        return self._binary(self.cmp, ("&&",))

    def cmp(self):
        return self._binary(self.add, ("<", ">", "<=", ">=", "==", "!="))

    def add(self):
        return self._binary(self.mul, ("+", "-"))

    def mul(self):
        return self._binary(self.unary, ("*", "/"))

    def unary(self):
        if self.accept("-"):
            return ("neg", self.unary())
        if self.accept("+"):
            return self.unary()
        if self.accept("!"):
            return ("not", self.unary())
        return self.postfix()

    def postfix(self):
        kind, val = self.next()
        if kind == "num":
            node = ("num", val)
        elif kind == "id":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    args.append(self.expr())
                    while self.accept(","):
                        args.append(self.expr())
                    self.expect(")")
                node = ("call", val, args)
            else:
                node = ("var", val)
        elif val == "(":
            node = self.expr()
            self.expect(")")
        else:
            raise SyntaxError(f"GLSL: unexpected {val!r} in expression")
        while self.accept("."):
            node = ("swz", node, self.next()[1])
        return node


class Texture:
    """RGBA texel array [H][W][4]; texel (i, j) of the reference = data[j, i] (index i + j*W)."""

    def __init__(self, data, width=None, height=None):
        a = np.asarray(data)
        if a.ndim == 2:
            a = a.reshape(height, width, a.shape[1])
        self.data = a
        self.h, self.w = a.shape[0], a.shape[1]


def _idx(u, n):
    with np.errstate(invalid="ignore"):
        t = u * u.dtype.type(n)
        i = np.floor(t)
        i = np.where(t > 0, i, 0)            # <= 0 and NaN -> texel 0
        i = np.where(t >= n, n - 1, i)
    return i.astype(np.int64)


class Shader:
    def __init__(self, src: str, dtype=np.float64):
        self.decls, self.main = Parser(src).parse()
        self.dtype = np.dtype(dtype)
        if self.main is None:
            raise SyntaxError("GLSL: no main()")

    # -- helpers ---------------------------------------------------------------------------
    def _lit(self, text):
        return np.full(self.n, self.dtype.type(float(text)), self.dtype)

    @staticmethod
    def _align(a, b):
        if a.ndim == b.ndim:
            return a, b
        if a.ndim == 1:
            return a[:, None], b
        return a, b[:, None]

    def _bin(self, op, a, b):
        if op in ("||", "&&"):
            return (a | b) if op == "||" else (a & b)
        a, b = self._align(a, b)
        with np.errstate(all="ignore"):
            if op == "+": return a + b
            if op == "-": return a - b
            if op == "*": return a * b
            if op == "/": return a / b
            if op == "<": return a < b
            if op == ">": return a > b
            if op == "<=": return a <= b
            if op == ">=": return a >= b
            if op == "==": return a == b
            if op == "!=": return a != b
        raise NotImplementedError(op)

    def _dot(self, a, b):
        with np.errstate(all="ignore"):
            acc = a[:, 0] * b[:, 0]
            for k in range(1, a.shape[1]):
                acc = acc + a[:, k] * b[:, k]
        return acc

    def _call(self, name, args):
        T = self.dtype.type
        if name in ("vec2", "vec3", "vec4", "float"):
            want = _TYPES[name]
            cols = []
            for a in args:
                if a.ndim == 1:
                    cols.append(a[:, None])
                else:
                    cols.append(a)
            v = np.concatenate(cols, 1).astype(self.dtype)
            if v.shape[1] == 1 and want > 1:
                v = np.repeat(v, want, 1)
            if v.shape[1] < want:
                raise TypeError(f"GLSL: {name}() needs {want} components, got {v.shape[1]}")
            v = v[:, :want]
            return v[:, 0] if want == 1 else v
        if name == "texture2D":
            tex, uv = args
            i, j = _idx(uv[:, 0], tex.w), _idx(uv[:, 1], tex.h)
            return tex.data[j, i].astype(self.dtype)
        with np.errstate(all="ignore"):
            if name == "sqrt": return np.sqrt(args[0])
            if name == "abs": return np.abs(args[0])
            if name == "floor": return np.floor(args[0])
            if name == "mod":  # GLSL ES 1.00 section 8.3: x - y * floor(x / y)
                a, b = self._align(*args)
                return a - b * np.floor(a / b)
            if name == "sign": return np.sign(args[0]).astype(self.dtype)
            if name == "cos":  # host libm, element by element, like the oracle's table (fsim_oracle.c: orc_cos_table)
                a = args[0]
                return np.array([math.cos(float(x)) for x in a.reshape(-1)], np.float64).reshape(a.shape).astype(self.dtype)
            if name in ("max", "min"):
                a, b = self._align(*args)
                return np.maximum(a, b) if name == "max" else np.minimum(a, b)
            if name == "dot": return self._dot(*args)
            if name == "length": return np.sqrt(self._dot(args[0], args[0]))
            if name == "cross":
                a, b = args
                return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                                 a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)
        raise NotImplementedError("GLSL builtin " + name)

    def _eval(self, node):
        kind = node[0]
        if kind == "num":
            return self._lit(node[1])
        if kind == "var":
            if node[1] not in self.env:
                raise NameError("GLSL: undefined " + node[1])
            return self.env[node[1]]
        if kind == "neg":
            return -self._eval(node[1])
        if kind == "not":
            return ~self._eval(node[1])
        if kind == "bin":
            a = self._eval(node[2])      # left operand first: left-to-right evaluation
            b = self._eval(node[3])
            return self._bin(node[1], a, b)
        if kind == "sel":
            c = self._eval(node[1])
            a, b = self._eval(node[2]), self._eval(node[3])
            a, b = self._align(a, b)
            return np.where(c[:, None] if a.ndim == 2 else c, a, b)
        if kind == "call":
            return self._call(node[1], [self._eval(a) for a in node[2]])
        if kind == "swz":
            v = self._eval(node[1])
            idx = [_SWZ[c] for c in node[2]]
            return v[:, idx[0]] if len(idx) == 1 else v[:, idx]
        raise NotImplementedError(kind)

    def _store(self, target, value, mask):
        if target[0] == "var":
            name = target[1]
            old = self.env.get(name)
            if old is None:
                self.env[name] = value
                return
            m = mask if value.ndim == 1 else mask[:, None]
            if old.ndim == 2 and value.ndim == 1:
                value = np.repeat(value[:, None], old.shape[1], 1)
            self.env[name] = np.where(m, value, old)
        elif target[0] == "swz" and target[1][0] == "var":
            name = target[1][1]
            old = self.env[name].copy()
            idx = [_SWZ[c] for c in target[2]]
            val = value if value.ndim == 2 else value[:, None]
            for k, ix in enumerate(idx):
                old[:, ix] = np.where(mask, val[:, k if val.shape[1] > 1 else 0], old[:, ix])
            self.env[name] = old
        else:
            raise NotImplementedError("GLSL: assignment target")

    def _exec(self, node, mask):
        kind = node[0]
        if kind == "block":
            for st in node[1]:
                self._exec(st, mask)
        elif kind == "decl":
            _, typ, name, init = node
            w = _TYPES[typ]
            if init is None:
                val = np.zeros(self.n if w == 1 else (self.n, w), self.dtype)
            else:
                val = self._eval(init)
                if w > 1 and val.ndim == 1:
                    val = np.repeat(val[:, None], w, 1)
            self.env[name] = val.astype(self.dtype)
        elif kind == "assign":
            _, op, target, rhs = node
            val = self._eval(rhs)
            if op != "=":
                val = self._bin(op[0], self._eval(target), val)
            self._store(target, val.astype(self.dtype), mask)
        elif kind == "if":
            c = self._eval(node[1])
            self._exec(node[2], mask & c)
            if node[3] is not None:
                self._exec(node[3], mask & ~c)
        elif kind == "for":
            _, init, cond, step, body = node
            self._exec(init, mask)
            guard = 0
            while True:
                live = mask & self._eval(cond)
                if not live.any():
                    break
                self._exec(body, live)
                self._exec(step, live)
                guard += 1
                if guard > 100000:
                    raise RuntimeError("GLSL: runaway loop")
        else:
            raise NotImplementedError(kind)

    # -- entry point ---------------------------------------------------------------------------
    def run(self, n: int, inputs: dict, outputs=("gl_FragColor",)):
        """Execute main() for n fragments/vertices.  `inputs` maps uniform/attribute/varying names to
        Texture objects, python floats (uniform scalars) or arrays of shape (n,) / (n, k)."""
        self.n = n
        self.env = {}
        for qual, typ, name, init in self.decls:
            if qual == "const":
                self.env[name] = self._eval(init)
                continue
            if name not in inputs:
                if qual == "varying" or typ == "sampler2D":
                    continue  # a varying this stage writes / a sampler that main() never reads
                if qual == "uniform" and typ in _TYPES:
                    w = _TYPES[typ]  # GL: a uniform that was never set holds zero (programMoments01's unused u_weight)
                    self.env[name] = np.zeros(n if w == 1 else (n, w), self.dtype)
                    continue
                raise KeyError(f"GLSL: no value for {qual} {name}")
            v = inputs[name]
            if isinstance(v, Texture):
                self.env[name] = v
            elif np.isscalar(v):
                self.env[name] = np.full(n, self.dtype.type(v), self.dtype)
            else:
                self.env[name] = np.asarray(v).astype(self.dtype)
        for k, v in inputs.items():  # built-in inputs such as gl_PointCoord
            if k.startswith("gl_"):
                self.env[k] = np.asarray(v).astype(self.dtype)
        for name in outputs:
            if name not in self.env:
                w = 1 if name == "gl_PointSize" else 4
                self.env[name] = np.zeros(n if w == 1 else (n, w), self.dtype)
        self._exec(self.main, np.ones(n, bool))
        return {name: self.env[name] for name in outputs}
