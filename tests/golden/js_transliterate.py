"""TEST TOOLING: mechanical transliteration of the small imperative JavaScript blocks of the reference
(the host-side table building inside set(), empic.js:1263-1339) into Python, so that THAT code -- read
out of /root/reference at run time, never copied -- can be executed here and the oracle's restatement
of it compared with its results.  Line-oriented, for the regular formatting the reference uses:

    for(i = a; i < N; i++) {        ->  for i in range(int(a), int(N)):
    while(cond) { / if (cond) {     ->  while cond: / if cond:
    var f = function(a, b) {        ->  def f(a, b):
    var x = e; / x++; / return e;   ->  x = e / x += 1 / return e
    throw new Error("m");           ->  raise RuntimeError("m")
    a.length, ===, !==, Math.min, Math.floor, [] (auto-growing array; out-of-range reads are
    `undefined`, which compares false and poisons arithmetic like NaN)

JS numbers are IEEE doubles: the caller passes numpy float64 scalars so that 0/0 is NaN, not an
exception."""
import math
import re
import types


class JsArray(list):
    """JS array: assignment past the end grows it, reading past the end gives undefined (NaN here)."""

    def __getitem__(self, k):
        if isinstance(k, float):
            if k != k or k != int(k):
                return float("nan")
            k = int(k)
        return list.__getitem__(self, k) if 0 <= k < len(self) else float("nan")

    def __setitem__(self, k, v):
        k = int(k)
        while len(self) <= k:
            self.append(float("nan"))
        list.__setitem__(self, k, v)


def js_floor(x):
    return math.floor(x) if x == x and abs(x) != float("inf") else float("nan")


def js_min(*a):
    return float("nan") if any(v != v for v in a) else min(a)


def _expr(e):
    e = e.strip().rstrip(";").strip()
    e = e.replace("===", "==").replace("!==", "!=").replace("||", " or ").replace("&&", " and ")
    e = re.sub(r"([A-Za-z_][\w\.]*(?:\[[^\]]*\])*)\.length", r"len(\1)", e)
    e = e.replace("Math.min", "js_min").replace("Math.floor", "js_floor").replace("Math.max", "max")
    e = e.replace("Math.random", "js_random")
    e = re.sub(r"\[\s*\]", "JsArray()", e)
    return e


def transliterate(js: str) -> str:
    out, depth = [], 0
    emit = lambda s: out.append("    " * depth + s)
    for raw in js.splitlines():
        line = raw.split("//")[0].strip()
        if not line:
            continue
        if line in ("}", "};"):
            depth -= 1
            continue
        m = re.match(r"^}\s*else\s*{$", line)
        if m:
            depth -= 1; emit("else:"); depth += 1
            continue
        m = re.match(r"^for\s*\(\s*(\w+)\s*=\s*(.+?)\s*;\s*\1\s*<\s*(.+?)\s*;\s*\1\+\+\s*\)\s*{$", line)
        if m:
            emit(f"for {m.group(1)} in range(int({_expr(m.group(2))}), int({_expr(m.group(3))})):"); depth += 1
            continue
        m = re.match(r"^(while|if)\s*\((.*)\)\s*{$", line)
        if m:
            emit(f"{m.group(1)} {_expr(m.group(2))}:"); depth += 1
            continue
        m = re.match(r"^var\s+(\w+)\s*=\s*function\s*\(([^)]*)\)\s*{$", line)
        if m:
            emit(f"def {m.group(1)}({m.group(2)}):"); depth += 1
            continue
        m = re.match(r'^throw new Error\((.*)\);$', line)
        if m:
            emit(f"raise RuntimeError({m.group(1)})")
            continue
        m = re.match(r"^return\s+(.*);$", line)
        if m:
            emit("return " + _expr(m.group(1)))
            continue
        m = re.match(r"^(\w+)\+\+;$", line)
        if m:
            emit(f"{m.group(1)} += 1")
            continue
        m = re.match(r"^var\s+(.*)$", line)
        if m:
            emit(_expr(m.group(1)))
            continue
        if line.endswith(";") and "{" not in line:
            emit(_expr(line))
            continue
        raise SyntaxError("JS form not handled by the transliterator: " + line)
    if depth != 0:
        raise SyntaxError("unbalanced braces in the JS block")
    return "\n".join(out)


def run(js: str, env: dict) -> dict:
    scope = dict(JsArray=JsArray, js_floor=js_floor, js_min=js_min, math=math, types=types)
    scope.update(env)
    exec(compile(transliterate(js), "<reference js>", "exec"), scope)
    return scope
