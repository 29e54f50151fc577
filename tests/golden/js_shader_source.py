"""TEST TOOLING: reads GLSL shader source out of the reference's JavaScript at run time.

The reference keeps every shader as a JS array of string pieces joined with '\n'
(`var src_arr = [ "...", "..." + N(expr) + "...", ... ]`, N(x) = x.toFixed(20)).  This module finds
those arrays in a .js file of /root/reference, evaluates the few JS expression forms they use
(string literals, +, parentheses, N(...), numeric sub-expressions, `cond ? a : b`, !==) and returns
the assembled GLSL text.  Nothing is copied into the repository; only generators that run where
the reference tree exists import this."""
import re


def js_array_elements(text, start):
    """Elements of the JS array literal whose '[' is at text[start]; returns (elements, end)."""
    i, depth, cur, out = start + 1, 0, [], []
    while True:
        c = text[i]
        if c == '"' or c == "'":
            j = i + 1
            while text[j] != c:
                j += 2 if text[j] == "\\" else 1
            cur.append(text[i:j + 1])
            i = j + 1
        elif text.startswith("//", i):
            i = text.index("\n", i)
        elif text.startswith("/*", i):
            i = text.index("*/", i) + 2
        elif c in "([":
            depth += 1; cur.append(c); i += 1
        elif c == ")" or (c == "]" and depth > 0):
            depth -= 1; cur.append(c); i += 1
        elif c == "]":
            if "".join(cur).strip():
                out.append("".join(cur).strip())
            return out, i
        elif c == "," and depth == 0:
            out.append("".join(cur).strip()); cur = []; i += 1
        else:
            cur.append(c); i += 1


_JS_TOKEN = re.compile(r"""\s*(?:("(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')|(\d+\.?\d*(?:[eE][-+]?\d+)?)|([A-Za-z_][A-Za-z0-9_.]*)|(!==|===|[-+*/?:(),]))""")


class JsExpr:
    """Evaluator for the JS expressions found inside the shader string arrays."""

    def __init__(self, text, env):
        self.toks, pos = [], 0
        text = text.strip()
        while pos < len(text):
            m = _JS_TOKEN.match(text, pos)
            if not m:
                raise SyntaxError("unexpected JS in a shader string array: " + text[pos:pos + 40])
            pos = m.end()
            self.toks.append(m.group(1) and ("str", m.group(1)) or m.group(2) and ("num", m.group(2)) or
                             m.group(3) and ("id", m.group(3)) or ("op", m.group(4)))
        self.i, self.env = 0, env

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else ("eof", "")

    def take(self, v=None):
        t = self.peek()
        if v is not None and t[1] != v:
            raise SyntaxError(f"JS: expected {v!r}, got {t!r}")
        self.i += 1
        return t

    def value(self):
        v = self.ternary()
        if self.peek()[0] != "eof":
            raise SyntaxError(f"JS: trailing {self.peek()!r}")
        return v

    def ternary(self):
        c = self.compare()
        if self.peek()[1] == "?":
            self.take()
            a = self.ternary()
            self.take(":")
            b = self.ternary()
            return a if c else b
        return c

    def compare(self):
        a = self.additive()
        while self.peek()[1] in ("!==", "==="):
            op = self.take()[1]
            b = self.additive()
            a = (a != b) if op == "!==" else (a == b)
        return a

    def additive(self):
        a = self.multiplicative()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            b = self.multiplicative()
            if op == "+" and (isinstance(a, str) or isinstance(b, str)):
                a = str(a) + str(b)  # only strings and N(...) results are ever concatenated
            else:
                a = a + b if op == "+" else a - b
        return a

    def multiplicative(self):
        a = self.unary()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            b = self.unary()
            a = a * b if op == "*" else a / b
        return a

    def unary(self):
        if self.peek()[1] == "-":
            self.take()
            return -self.unary()
        kind, v = self.take()
        if kind == "str":
            return bytes(v[1:-1], "utf-8").decode("unicode_escape")
        if kind == "num":
            return float(v)
        if v == "(":
            x = self.ternary()
            self.take(")")
            return x
        if kind == "id":
            if self.peek()[1] == "(":
                self.take()
                args = []
                if self.peek()[1] != ")":
                    args.append(self.ternary())
                    while self.peek()[1] == ",":
                        self.take()
                        args.append(self.ternary())
                self.take(")")
                if v == "N":
                    return "%.20f" % args[0]  # Number.prototype.toFixed(20)
                if v == "Math.pow":
                    return float(args[0]) ** float(args[1])
                raise NameError("JS function " + v)
            return float(self.env[v])
        raise SyntaxError(f"JS: unexpected {v!r}")


def shader_sources(path, env, owner_env=None):
    """{name: [GLSL source, ...]} -- name = the `var NAME =` that owns each `src_arr` array.
    owner_env: {name: extra variables} for arrays built inside a function with parameters."""
    text = open(path).read()
    owners = [(m.start(), m.group(1)) for m in
              re.finditer(r"var\s+(\w+)\s*=\s*(?:function\s*\(|webgl\.linkProgram\s*\()", text)]
    out = {}
    for m in re.finditer(r"var\s+src_arr\s*=\s*\[", text):
        name = [n for pos, n in owners if pos < m.start()][-1]
        elems, _ = js_array_elements(text, m.end() - 1)
        e = dict(env, **(owner_env or {}).get(name, {}))
        try:
            out.setdefault(name, []).append("\n".join(JsExpr(x, e).value() for x in elems))
        except KeyError:
            out.setdefault(name, []).append(None)  # needs variables the caller did not give (not on this path)
    return out


# ---- which texture feeds which uniform, and in which order the draws happen -------------------------
def _match(text, i):
    """Index of the bracket matching the opening bracket at text[i] (strings and comments skipped)."""
    pairs = {"(": ")", "{": "}", "[": "]"}
    stack = [pairs[text[i]]]
    i += 1
    while stack:
        c = text[i]
        if c in "\"'":
            j = i + 1
            while text[j] != c:
                j += 2 if text[j] == "\\" else 1
            i = j
        elif text.startswith("//", i):
            i = text.index("\n", i)
        elif text.startswith("/*", i):
            i = text.index("*/", i) + 1
        elif c in pairs:
            stack.append(pairs[c])
        elif c == stack[-1]:
            stack.pop()
        i += 1
    return i - 1


def _object_literal(text):
    """{'key': 'value text'} of a flat JS object literal body (keys quoted or bare)."""
    out, i, depth, cur = {}, 0, 0, []
    parts = []
    while i < len(text):
        c = text[i]
        if text.startswith("//", i):
            i = text.index("\n", i) if "\n" in text[i:] else len(text)
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append("".join(cur)); cur = []
        else:
            cur.append(c)
        i += 1
    parts.append("".join(cur))
    for p in parts:
        if ":" in p:
            k, v = p.split(":", 1)
            out[k.strip().strip("\"'")] = v.strip()
    return out


def _strip_comments(text):
    out, i = [], 0
    while i < len(text):
        c = text[i]
        if c in "\"'":
            j = i + 1
            while text[j] != c:
                j += 2 if text[j] == "\\" else 1
            out.append(text[i:j + 1]); i = j + 1
        elif text.startswith("//", i):
            i = text.index("\n", i) if "\n" in text[i:] else len(text)
        elif text.startswith("/*", i):
            i = text.index("*/", i) + 2
        else:
            out.append(c); i += 1
    return "".join(out)


def program_bindings(path):
    """{program: {"fragment": owner of its fragment source, "uniforms": {name: JS value text}}} from
    `var P = webgl.linkProgram({...}).set({...})`."""
    text = open(path).read()
    out = {}
    for m in re.finditer(r"var\s+(\w+)\s*=\s*webgl\.linkProgram\s*\(", text):
        name = m.group(1)
        end = _match(text, m.end() - 1)
        args = text[m.end():end]
        f = re.search(r"fragmentShaderSource\s*:\s*(\w+)\s*\(\s*\)", args)
        entry = {"fragment": f.group(1) if f else name, "uniforms": {}}
        rest = text[end + 1:]
        s = re.match(r"\s*\.set\s*\(\s*\{", rest)
        if s:
            close = _match(rest, s.end() - 1)
            entry["uniforms"] = _object_literal(rest[s.end():close])
        out[name] = entry
    return out


def draw_sequence(path, method):
    """[(program, {draw options}, {uniforms set in the same statement})] of `out.METHOD = function...`."""
    text = open(path).read()
    m = re.search(r"out\." + method + r"\s*=\s*function\s*\([^)]*\)\s*\{", text)
    body = _strip_comments(text[m.end():_match(text, m.end() - 1)])
    seq = []
    for d in re.finditer(r"(\w+)((?:\s*\.set\s*\(\s*\{[^}]*\}\s*\))?)\s*\.draw\s*\(\s*\{", body):
        close = _match(body, d.end() - 1)
        opts = _object_literal(body[d.end():close])
        sets = {}
        s = re.search(r"\{([^}]*)\}", d.group(2)) if d.group(2) else None
        if s:
            sets = _object_literal(s.group(1))
        seq.append((d.group(1), opts, sets))
    return seq
