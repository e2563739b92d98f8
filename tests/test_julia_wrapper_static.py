"""The Julia wrapper cannot be executed here (no Julia in the image), so its ccall signatures are checked statically
against include/snake_b200.h: every bound symbol exists, takes the same number of arguments, and each Julia argument type is
one that matches the C parameter type (Cint for int, Int64 for int64_t, Ptr/Ref for pointers and handles, ...)."""
import os
import re

from tests.util import ROOT

HEADER = os.path.join(ROOT, "include", "snake_b200.h")
WRAPPER = os.path.join(ROOT, "laplace-dqn-snake-game_b200", "julia", "SnakeB200.jl")


def _split_top(s):
    """split on commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


rets = {}       # symbol -> C return type, filled by c_declarations()


def c_declarations():
    src = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    handles = set(re.findall(r"typedef\s+struct\s+\w+\s*\*\s*(\w+)\s*;", src))
    decls = {}
    rets.clear()
    for ret, name, params in re.findall(r"SNK_API\s+([\w\s\*]+?)\s*\**\s*(snk_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ps = [] if params.strip() in ("", "void") else _split_top(" ".join(params.split()))
        decls[name] = ps
        rets[name] = " ".join(ret.split())
    return decls, handles


def julia_ccalls():
    src = open(WRAPPER).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(snk_\w+),\s*lib\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                                   # the argument-type tuple, balanced
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = _split_top(src[m.end():i - 1])
        calls.append((m.group(1), m.group(2), types, src.count("\n", 0, m.start()) + 1))
    return calls


def _compatible(cparam, jtype, handles):
    c = cparam.replace("const ", "").strip()
    base = c.rsplit(" ", 1)[0].strip() if " " in c else c      # drop the parameter name
    if "*" in c:
        return jtype.startswith(("Ptr{", "Ref{")) or jtype == "Cstring"
    if base in handles:
        return jtype in ("Ptr{Cvoid}",)
    table = {
        "int": {"Cint", "Int32"}, "int32_t": {"Cint", "Int32"}, "unsigned": {"Cuint", "UInt32"}, "unsigned int": {"Cuint", "UInt32"},
        "uint32_t": {"Cuint", "UInt32"}, "int64_t": {"Int64", "Clonglong"}, "long long": {"Int64", "Clonglong"},
        "uint64_t": {"UInt64", "Culonglong"}, "float": {"Cfloat", "Float32"}, "double": {"Cdouble", "Float64"},
        "size_t": {"Csize_t", "UInt64"}, "uint8_t": {"UInt8"},
    }
    return jtype in table.get(base, set())


def test_every_ccall_matches_a_header_declaration():
    decls, handles = c_declarations()
    assert len(decls) >= 50 and "snk_step_fused_host" in decls and handles
    calls = julia_ccalls()
    assert len(calls) >= 25
    problems = []
    for name, ret, types, line in calls:
        if name not in decls:
            problems.append("%s (line %d): not declared in the header" % (name, line))
            continue
        params = decls[name]
        if len(params) != len(types):
            problems.append("%s (line %d): %d Julia argument types for %d C parameters" % (name, line, len(types), len(params)))
            continue
        for k, (cp, jt) in enumerate(zip(params, types)):
            if not _compatible(cp, jt, handles):
                problems.append("%s (line %d): argument %d is `%s` in C but `%s` in Julia" % (name, line, k + 1, cp, jt))
        want = {"int": "Cint", "int64_t": "Int64", "const char": "Cstring"}[rets[name]]
        if ret != want:
            problems.append("%s (line %d): returns %s, bound as %s" % (name, line, rets[name], ret))
    assert not problems, "\n".join(problems)


def _julia_definitions(src):
    """name -> concatenated source text of every method / constructor of that name (block `function` forms up to the matching
    column-0 `end`, one-line `name(args) = ...` forms up to the next blank line, `mutable struct` blocks incl. inner constructors)"""
    defs = {}
    lines = src.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"(?:mutable\s+)?struct\s+(\w+)", ln) or re.match(r"function\s+(?:Base\.)?(\w+!?)\s*\(", ln)
        if m:
            j = i + 1
            while j < len(lines) and not re.match(r"end\b", lines[j]):
                j += 1
            defs[m.group(1)] = defs.get(m.group(1), "") + "\n".join(lines[i:j + 1]) + "\n"
            i = j + 1
            continue
        m = re.match(r"(?:Base\.)?(\w+!?)\(", ln)           # at column 0 a `name(` line is always a short-form definition here
        if m:
            j = i
            while j + 1 < len(lines) and lines[j + 1].startswith((" ", "\t")):
                j += 1
            defs[m.group(1)] = defs.get(m.group(1), "") + "\n".join(lines[i:j + 1]) + "\n"
            i = j + 1
            continue
        i += 1
    return defs


def test_every_exported_julia_name_reaches_a_ccall():
    """no exported function answers from Julia-side bookkeeping alone (round 1's assemble_state! fabricated the board):
    each exported name has a definition whose body contains a ccall, or calls a function of the module that does"""
    src = open(WRAPPER).read()
    exported = re.search(r"^export\s+(.*?)\n\n", src, flags=re.S | re.M).group(1)
    names = [n.strip() for n in exported.replace("\n", " ").split(",") if n.strip()]
    assert len(names) >= 40 and "assemble_state!" in names and "virtual_step" in names and "GramShard" in names
    defs = _julia_definitions(src)
    missing = [n for n in names if n not in defs]
    assert not missing, "exported but not defined: %s" % missing
    reaches = {n for n, body in defs.items() if "ccall(" in body}
    changed = True
    while changed:
        changed = False
        for n, body in defs.items():
            if n not in reaches and any(re.search(r"(?<![\w!])%s\(" % re.escape(r), body) for r in reaches):
                reaches.add(n)
                changed = True
    dead = [n for n in names if n not in reaches]
    assert not dead, "exported names that never reach the library: %s" % dead
    # the host getters must be the library's, not Julia-side mirrors
    for fn, sym in (("assemble_state!", "snk_state_host"), ("virtual_step", "snk_losing_mask_host"), ("lost", "snk_get_done_host"),
                    ("score", "snk_get_score_host"), ("available_actions", "snk_available_actions_host")):
        assert sym in defs[fn], (fn, sym)
