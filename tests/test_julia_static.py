"""The Julia files of this repository (the `ccall` glue and the golden-vector generators) cannot be executed here —
there is no Julia in the image — so the least a CPU test can do is read them the way Julia's parser would at the
block level: strings, comments and character literals set aside, every `function / if / for / while / begin / let /
do / struct / module / try / quote / macro` closed by its `end`, brackets balanced, no `end` left over.  It is a
block-structure check, not a parser; tests/test_abi.py checks the `ccall` signatures against the header."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = sorted(glob.glob(os.path.join(ROOT, "julia", "*.jl")))
OPENERS = {"function", "if", "for", "while", "begin", "let", "do", "struct", "module", "baremodule", "try", "quote", "macro"}


def strip_julia(src):
    """comments, strings (with $(...) interpolation, whose code is kept) and char literals blanked out"""
    out, i, n = [], 0, len(src)
    stack = []                      # ')' entries for interpolations we are inside of
    prev_sig = ""                   # last significant character kept (to tell 'x' from the transpose x')

    def skip_string(i, triple):
        q = '"""' if triple else '"'
        i += len(q)
        while i < n:
            if src[i] == "\\":
                i += 2
                continue
            if src.startswith("$(", i):            # interpolation: hand the code back to the main loop
                return i + 2, True
            if src.startswith(q, i):
                return i + len(q), False
            i += 1
        raise AssertionError("unterminated string")

    in_string = []                  # stack of (triple) for strings suspended by an interpolation
    depth_at_interp = []
    depth = 0
    while i < n:
        c = src[i]
        if src.startswith("#=", i):
            j = src.index("=#", i) + 2
            out.append(" " * (j - i - src.count("\n", i, j)) + "\n" * src.count("\n", i, j))
            i = j
        elif c == "#":
            j = src.find("\n", i)
            j = n if j < 0 else j
            i = j
        elif c == '"':
            triple = src.startswith('"""', i)
            i, interp = skip_string(i, triple)
            out.append(' "" ')
            prev_sig = '"'
            if interp:
                in_string.append(triple)
                depth_at_interp.append(depth)
                depth += 1
                out.append("(")
        elif c == "'" and not (prev_sig.isalnum() or prev_sig in "_)]}'"):
            j = i + 1
            if src[j] == "\\":
                j += 1
            j = src.index("'", j + 1)
            out.append(" 'c' ")
            i = j + 1
            prev_sig = "'"
        else:
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
                if in_string and depth == depth_at_interp[-1]:
                    out.append(")")
                    triple = in_string.pop()
                    depth_at_interp.pop()
                    # resume the suspended string right after the interpolation
                    q = '"""' if triple else '"'
                    i += 1
                    while i < n:
                        if src[i] == "\\":
                            i += 2
                            continue
                        if src.startswith("$(", i):
                            in_string.append(triple)
                            depth_at_interp.append(depth)
                            depth += 1
                            out.append("(")
                            i += 2
                            break
                        if src.startswith(q, i):
                            i += len(q)
                            break
                        i += 1
                    continue
            out.append(c)
            if not c.isspace():
                prev_sig = c
            i += 1
    assert not in_string, "string with an unterminated interpolation"
    return "".join(out)


def block_balance(code):
    """walk the tokens: openers push, `end` pops; inside [...] `for`/`if` belong to comprehensions and `end` is an
    index; `abstract type` / `primitive type` open a block too; returns the list of problems"""
    problems, blocks, brackets = [], [], []
    toks = re.finditer(r"[A-Za-z_ -￿][A-Za-z_0-9! -￿]*|[()\[\]{}]|\n|:(?=[A-Za-z_])|\.(?=[A-Za-z_])|\S", code)
    line, prev = 1, ""
    for m in toks:
        t = m.group(0)
        if t == "\n":
            line += 1
            continue
        if t in "([{":
            brackets.append((t, line))
        elif t in ")]}":
            if not brackets or "([{".index(brackets[-1][0]) != ")]}".index(t):
                problems.append(f"line {line}: unmatched {t}")
            else:
                brackets.pop()
        elif prev in (":", "."):
            pass                                     # a symbol (:end) or a field (x.begin): not a keyword
        elif t in OPENERS:
            in_square = any(b[0] == "[" for b in brackets)
            in_generator = bool(brackets) and t in ("for", "if") and prev not in ("", "(", ";") and not prev == "begin"
            if t in ("for", "if") and (in_square or in_generator):
                pass
            elif t == "struct" and prev == "mutable":
                blocks.append((t, line))
            else:
                blocks.append((t, line))
        elif t == "type" and prev in ("abstract", "primitive"):
            blocks.append((t, line))
        elif t == "end":
            if any(b[0] == "[" for b in brackets) and (not blocks or blocks[-1][1] < brackets[-1][1]):
                pass                                 # a[end]
            elif not blocks:
                problems.append(f"line {line}: `end` without a block")
            else:
                blocks.pop()
        prev = t
    problems += [f"line {ln}: `{b}` never closed" for b, ln in blocks]
    problems += [f"line {ln}: `{b}` never closed" for b, ln in brackets]
    return problems


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_julia_file_is_block_balanced(path):
    src = open(path, encoding="utf-8").read()
    assert block_balance(strip_julia(src)) == []


def test_the_checker_sees_what_it_should():
    ok = '''
    module M
    f(x) = x[end] + 1          # one-liner, `end` as an index
    function g(v; k=[i for i in 1:3 if i > 1])
        s = "a $(v[end]) \\" b) end"   # keywords and brackets inside a string
        for i in v
            if i > 0 && (c = 'e'; true)
                s *= string(i)
            elseif i == 0
                continue
            end
        end
        map(x -> begin x + 1 end, v)
        y = v'                   # transpose, not a character
        return (s, :end, M.begin)
    end
    abstract type A end
    mutable struct B <: A
        x::Int
    end
    end
    '''
    assert block_balance(strip_julia(ok)) == []
    for broken in (ok.replace("            end\n        end\n", "            end\n", 1),    # a `for` left open
                   ok.replace("abstract type A end", "abstract type A"),
                   ok.replace("map(x -> begin x + 1 end, v)", "map(x -> begin x + 1 end, v))"),
                   ok + "\nend\n"):
        assert block_balance(strip_julia(broken)) != []
