"""The Rust binding (rust-shim/src/ffi.rs) cannot be compiled in this image, so it is checked mechanically against the C ABI:
every function include/hmgpu.h declares must be bound, with the same arity, parameter order and types (C -> Rust FFI mapping
restated here independently of the generator), every status / op constant must be present with its value, and the committed
file must be what tools/gen_rust_ffi.py generates from the current header."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FFI = os.path.join(ROOT, "rust-shim", "src", "ffi.rs")
HDR = os.path.join(ROOT, "include", "hmgpu.h")

C2RUST_BASE = {"int": "c_int", "long": "c_long", "double": "f64", "size_t": "usize", "char": "c_char", "void": "c_void",
               "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64",
               "hm_context": "hm_context", "hm_batch": "hm_batch", "hm_group": "hm_group", "hm_group_batch": "hm_group_batch"}


def c_to_rust(ctype: str) -> str:
    """`const uint8_t *const *` -> `*const *const u8`, by walking the C declarator right to left."""
    s = ctype.replace("*", " * ")
    toks = s.split()
    base = [t for t in toks if t not in ("*", "const")]
    assert len(base) == 1, ctype
    first_star = toks.index("*") if "*" in toks else len(toks)
    const_base = "const" in toks[:first_star]
    out = C2RUST_BASE[base[0]]
    quals = toks[first_star:]  # e.g. ['*', 'const', '*']
    pointee_is_const = const_base
    j = 0
    while j < len(quals):
        assert quals[j] == "*"
        out = ("*const " if pointee_is_const else "*mut ") + out
        pointee_is_const = j + 1 < len(quals) and quals[j + 1] == "const"
        j += 2 if pointee_is_const else 1
    return out


def header_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    text = re.sub(r"typedef\s+enum.*?;", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(hm_\w+)\s*\(([^;{}]*?)\)\s*;", text):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        params = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = a.strip()
                arr = a.endswith("]")
                a = re.sub(r"\[\d*\]$", "", a)
                ctype = re.match(r"^(.*?)(\w+)$", a).group(1).strip() + (" *" if arr else "")
                params.append(c_to_rust(ctype))
        out[name] = (None if ret == "void" else c_to_rust(ret), params)
    return out


def rust_functions():
    src = open(FFI).read()
    out = {}
    for m in re.finditer(r"pub fn (hm_\w+)\((.*?)\)(?:\s*->\s*([^;]+))?;", src):
        params = [p.split(":", 1)[1].strip() for p in m.group(2).split(",") if p.strip()]
        out[m.group(1)] = (m.group(3).strip() if m.group(3) else None, params)
    return out


def test_every_header_function_is_bound_with_the_same_signature():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 90
    assert sorted(h) == sorted(r), f"missing in ffi.rs: {sorted(set(h) - set(r))}; not in the header: {sorted(set(r) - set(h))}"
    for name in h:
        assert h[name] == r[name], f"{name}: header {h[name]} vs ffi.rs {r[name]}"


def test_bound_names_are_what_the_library_exports():
    sys.path.insert(0, ROOT)
    from homomorph_rust_b200 import _native as N

    assert sorted(rust_functions()) == N.exported_symbols()
    lib = N.lib()
    for name in rust_functions():
        assert hasattr(lib, name)


def test_constants_match_the_header():
    text = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    consts = dict((k, int(v)) for k, v in re.findall(r"\b(HM_(?:OK|ERR_\w+|OP_\w+))\s*=\s*(-?\d+)", text))
    assert len(consts) == 11 + 6
    rs = dict((k, int(v)) for k, v in re.findall(r"pub const (HM_\w+): c_int = (-?\d+);", open(FFI).read()))
    assert rs == consts


def test_committed_binding_is_up_to_date():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode == 0, \
        "include/hmgpu.h changed: run python tools/gen_rust_ffi.py"


def test_safe_layer_only_calls_bound_functions():
    bound = set(rust_functions())
    for fn in ("lib.rs", "mask_rng.rs"):
        src = open(os.path.join(ROOT, "rust-shim", "src", fn)).read()
        used = set(re.findall(r"\bffi::(hm_\w+)\s*\(", src))
        assert used, fn
        assert used <= bound, f"{fn} calls unbound functions: {sorted(used - bound)}"
