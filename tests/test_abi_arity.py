"""Every ctypes call site passes exactly as many arguments as the prototype in include/dggb.h declares
(ctypes does not check arity for cdecl functions: a drifted call would corrupt the stack silently)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "learning-adaptive-neighborhoods-for-gnns_b200")


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _matching_paren(text, start):
    depth = 0
    for i in range(start, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise AssertionError("unbalanced parentheses")


def header_arity():
    text = open(os.path.join(ROOT, "include", "dggb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(dggb_[a-z0-9_]+)\s*\(", text):
        end = _matching_paren(text, m.end() - 1)
        args = text[m.end():end].strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top_level(args))
    return protos


def call_sites():
    files = [os.path.join(PKG, f) for f in os.listdir(PKG) if f.endswith(".py")]
    files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for path in files:
        text = open(path).read()
        for m in re.finditer(r"(?:lib\(\)|\bL|\b_lib)\.(dggb_[a-z0-9_]+)\s*\(", text):
            end = _matching_paren(text, m.end() - 1)
            args = text[m.end():end].strip()
            yield os.path.basename(path), m.group(1), (0 if args == "" else len(_split_top_level(args)))


def test_ctypes_calls_match_header_arity():
    protos = header_arity()
    assert len(protos) >= 25
    seen = 0
    for fname, sym, n_args in call_sites():
        assert sym in protos, f"{fname}: {sym} is not declared in include/dggb.h"
        assert n_args == protos[sym], f"{fname}: {sym} called with {n_args} args, header declares {protos[sym]}"
        seen += 1
    assert seen >= 20
