"""The Rust FFI declarations (bindings/rust/b2z-sys/src/lib.rs) against the C ABI (include/b2z.h): same functions, same
arity, same pointer / size classes for every parameter and return value, same struct fields -- parsed independently
of the generator (tools/gen_rust_sys.py), so INTEGRATION.md cannot claim a binding for a symbol that does not exist.
(The build image has no Rust toolchain: the crates are committed uncompiled; this is the check that can run here.)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_protos():
    txt = open(os.path.join(ROOT, "include", "b2z.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"#.*", "", txt)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(b2z_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        ps = [] if params in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
        out[name] = (ret, ps)
    return out


def _c_class(decl: str) -> str:
    """'const uint32_t* in' -> a coarse class: pointer levels, constness of the outermost pointee, base kind.
    Every declaration carries a name as its last identifier (return types get a dummy one)."""
    toks = re.findall(r"[A-Za-z_]\w*|\*|\[\d*\]", decl)
    idents = [k for k, t in enumerate(toks) if re.match(r"[A-Za-z_]", t)]
    toks = toks[: idents[-1]] + toks[idents[-1] + 1:]  # drop the name
    levels = sum(1 for t in toks if t == "*" or t.startswith("["))
    const0 = toks[0] == "const"
    words = [t for t in toks if re.match(r"[A-Za-z_]", t) and t not in ("const", "struct")]
    b = words[0]
    if b == "b2z_err":
        return "p1c:char"
    kind = {"char": "char", "void": "void", "int": "int", "unsigned": "uint", "size_t": "usize", "uint32_t": "u32", "uint64_t": "u64",
            "uint8_t": "u8", "double": "f64"}.get(b, b)
    return f"p{levels}{'c' if (levels and const0) else ''}:{kind}"


def _rust_class(t: str) -> str:
    t = t.strip()
    levels = t.count("*")
    const0 = False
    parts = re.findall(r"\*(const|mut)", t)
    if parts:
        const0 = parts[-1] == "const"  # the innermost pointer written last
    base = re.sub(r"\*(const|mut)\s*", "", t).strip()
    kind = {"c_char": "char", "c_void": "void", "c_int": "int", "c_uint": "uint", "usize": "usize", "u32": "u32", "u64": "u64", "u8": "u8",
            "f64": "f64"}.get(base, base)
    return f"p{levels}{'c' if (levels and const0) else ''}:{kind}"


def _rust_protos():
    txt = open(os.path.join(ROOT, "bindings", "rust", "b2z-sys", "src", "lib.rs")).read()
    blk = txt[txt.index('extern "C" {'):]
    out = {}
    for m in re.finditer(r"pub fn (b2z_\w+)\((.*?)\)(?:\s*->\s*([^;]+))?;", blk):
        name, params, ret = m.group(1), m.group(2).strip(), (m.group(3) or "()").strip()
        ps = [] if not params else [p.split(":", 1)[1].strip() for p in params.split(", ")]
        out[name] = (ret, ps)
    return out


def test_rust_extern_block_matches_the_header_one_to_one():
    c, r = _c_protos(), _rust_protos()
    assert len(c) >= 65
    assert set(c) == set(r), f"only in b2z.h: {sorted(set(c) - set(r))}; only in lib.rs: {sorted(set(r) - set(c))}"
    for name, (cret, cps) in c.items():
        rret, rps = r[name]
        assert len(cps) == len(rps), f"{name}: arity {len(cps)} (C) vs {len(rps)} (Rust)"
        want_ret = "p0:()" if cret == "void" else _c_class(cret + " x")
        got_ret = "p0:()" if rret == "()" else _rust_class(rret)
        assert want_ret == got_ret, f"{name}: return {cret!r} vs {rret!r}"
        for i, (cp, rp) in enumerate(zip(cps, rps)):
            assert _c_class(cp) == _rust_class(rp), f"{name} arg {i}: {cp!r} ({_c_class(cp)}) vs {rp!r} ({_rust_class(rp)})"


def test_rust_structs_match_the_header():
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "b2z.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "bindings", "rust", "b2z-sys", "src", "lib.rs")).read()
    for name in ("b2z_segment_input", "b2z_segment_receipt", "b2z_receipt"):
        cbody = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), h, flags=re.S).group(1)
        cfields = [re.match(r"^(.*?)([A-Za-z_]\w*)(\[\d+\])?$", " ".join(f.split())).group(2) for f in cbody.split(";") if f.strip()]
        rbody = re.search(r"pub struct %s \{(.*?)\n\}" % name, rs, flags=re.S).group(1)
        rfields = re.findall(r"pub (\w+):", rbody)
        assert cfields == rfields, f"{name}: {cfields} vs {rfields}"


def test_safe_wrapper_only_calls_declared_symbols():
    sysrs = open(os.path.join(ROOT, "bindings", "rust", "b2z-sys", "src", "lib.rs")).read()
    hal = open(os.path.join(ROOT, "bindings", "rust", "b2z-hal", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (b2z_\w+)\(", sysrs))
    used = set(re.findall(r"sys::(b2z_\w+)\(", hal))
    assert used and used <= declared, f"b2z-hal calls undeclared symbols: {sorted(used - declared)}"


def test_safe_wrapper_calls_every_symbol_with_the_declared_number_of_arguments():
    """cargo is not available here, so the cheapest class of compile errors is caught by hand: each `sys::b2z_*(...)` call in
    b2z-hal passes as many arguments as the `extern "C"` declaration of b2z-sys takes"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sysrs = open(os.path.join(root, "bindings", "rust", "b2z-sys", "src", "lib.rs")).read()
    hal = open(os.path.join(root, "bindings", "rust", "b2z-hal", "src", "lib.rs")).read()
    decl = {}
    for m in re.finditer(r"pub fn (b2z_\w+)\((.*?)\)\s*(->\s*[^;]+)?;", sysrs, re.S):
        decl[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])

    def split_top(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            depth += ch in "([{"
            depth -= ch in ")]}"
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        if cur.strip():
            out.append(cur)
        return out

    calls = 0
    for m in re.finditer(r"sys::(b2z_\w+)\(", hal):
        i = j = m.end()
        depth = 1
        while depth:
            depth += hal[j] == "("
            depth -= hal[j] == ")"
            j += 1
        assert decl.get(m.group(1)) == len(split_top(hal[i:j - 1])), m.group(1)
        calls += 1
    assert calls >= 40
    # a type that implements Drop cannot be rebuilt with struct-update syntax (E0509)
    assert not re.search(r"\.\.\s*b\s*\}", hal)
