#!/usr/bin/env python
"""Generate the Rust FFI stub (`src/gpu/sys.rs`) a raft-rs maintainer would add, from include/raftgpu.h.

    python scripts/gen_rust_stub.py            # print the stub
    python scripts/gen_rust_stub.py --update   # rewrite the block between the markers in INTEGRATION.md

This image has no Rust toolchain (no bindgen), so the stub is produced by this small parser of the header's
regular subset (numeric #defines, `typedef struct {..} name;`, `int32_t raftgpu_*(..);` prototypes).
tests/test_abi.py checks that INTEGRATION.md carries exactly this output and that every struct's size and
field offsets agree with the ctypes mirror (binding.py) -- the 24-byte `raftgpu_step_result` that round 1's
hand-written stub declared against the header's 48 bytes cannot happen again."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "raftgpu.h")
DOC = os.path.join(ROOT, "INTEGRATION.md")
BEGIN, END = "<!-- BEGIN generated: scripts/gen_rust_stub.py -->", "<!-- END generated -->"

SCALAR = {"uint8_t": ("u8", 1), "uint16_t": ("u16", 2), "uint32_t": ("u32", 4), "int32_t": ("i32", 4),
          "uint64_t": ("u64", 8), "int64_t": ("i64", 8), "char": ("c_char", 1)}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def parse_header(path=HEADER):
    raw = open(path, encoding="utf-8").read()
    text = strip_comments(raw)
    consts = []
    for m in re.finditer(r"^#define\s+(RAFTGPU_[A-Z0-9_]+)\s+(.+?)\s*$", text, flags=re.M):
        name, val = m.group(1), m.group(2).strip()
        mm = re.fullmatch(r"\(?(-?\s*(?:0x[0-9a-fA-F]+|\d+))(u|ull|ULL|U)?\)?", val.replace(" ", ""))
        if mm:
            consts.append((name, int(mm.group(1), 0), mm.group(2) or ""))
        elif val == "UINT64_MAX":
            consts.append((name, (1 << 64) - 1, "ull"))
    structs = []
    for m in re.finditer(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            fm = re.fullmatch(r"(const\s+)?(\w+)\s*(\*?)\s*(.+)", decl)
            const, ctype, first_ptr, rest = fm.groups()
            for item in rest.split(","):
                item = item.strip()
                ptr = bool(first_ptr) or item.startswith("*")
                item = item.lstrip("* ")
                am = re.fullmatch(r"(\w+)\[(\d+)\]", item)
                fields.append({"name": am.group(1) if am else item, "ctype": ctype, "ptr": ptr, "const": bool(const),
                               "array": int(am.group(2)) if am else 0})
                first_ptr = ""   # `uint64_t *a, *b` : the star belongs to each declarator
        structs.append((m.group(2), fields))
    opaque = re.findall(r"typedef\s+struct\s+(\w+)\s+\1\s*;", text)
    funcs = []
    for m in re.finditer(r"^\s*(const\s+char\s*\*|int32_t|uint32_t|uint64_t)\s*(raftgpu_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        funcs.append((name, ret.replace(" ", ""), [] if args in ("void", "") else [a.strip() for a in args.split(",")]))
    return consts, opaque, structs, funcs


def rust_type(ctype, ptr_depth, const):
    if ctype == "void":
        base = "c_void"
    elif ctype in SCALAR:
        base = SCALAR[ctype][0]
    else:
        base = ctype
    t = base
    for d in range(ptr_depth):
        t = ("*const " if (const and d == 0) else "*mut ") + t
    return t


def arg_to_rust(arg):
    m = re.fullmatch(r"(const\s+)?(?:struct\s+)?(\w+)\s*((?:\*\s*(?:const\s*)?)*)\s*(\w+)?", arg)
    const, ctype, stars, name = m.groups()
    return (name or "_"), rust_type(ctype, stars.count("*"), bool(const))


def layout(fields, sizes):
    """C layout: (size, align, [(name, offset, size)])"""
    off, align, out = 0, 1, []
    for f in fields:
        if f["ptr"]:
            sz = al = 8
        elif f["ctype"] in SCALAR:
            sz = al = SCALAR[f["ctype"]][1]
        else:
            sz, al = sizes[f["ctype"]]
        n = max(1, f["array"])
        off = (off + al - 1) // al * al
        out.append((f["name"], off, sz * n))
        off += sz * n
        align = max(align, al)
    return (off + align - 1) // align * align, align, out


def struct_layouts():
    _, _, structs, _ = parse_header()
    sizes, res = {}, {}
    for name, fields in structs:
        size, align, offs = layout(fields, sizes)
        sizes[name] = (size, align)
        res[name] = (size, offs)
    return res


def generate():
    consts, opaque, structs, funcs = parse_header()
    lay = struct_layouts()
    out = ["// src/gpu/sys.rs -- generated from include/raftgpu.h by scripts/gen_rust_stub.py; do not edit by hand.",
           "#![allow(non_camel_case_types, dead_code)]", "use std::os::raw::{c_char, c_void};", ""]
    for name, val, suf in consts:
        ty = "i32" if (val < 0 or name == "RAFTGPU_OK") else ("u64" if (suf.lower().startswith("ull") or val > 0xFFFFFFFF) else "u32")
        out.append(f"pub const {name}: {ty} = {val if val < 0 else hex(val) if val > 255 else val};")
    out.append("")
    for name in opaque:
        out.append(f"#[repr(C)] pub struct {name} {{ _private: [u8; 0] }}   // opaque")
    for name, fields in structs:
        out.append(f"#[repr(C)] #[derive(Clone, Copy)]")
        out.append(f"pub struct {name} {{   // {lay[name][0]} bytes")
        for f in fields:
            t = rust_type(f["ctype"], 1 if f["ptr"] else 0, f["const"])
            if f["array"]:
                t = f"[{t}; {f['array']}]"
            out.append(f"    pub {f['name']}: {t},")
        out.append("}")
        out.append(f"const _: () = assert!(std::mem::size_of::<{name}>() == {lay[name][0]});")
    out.append("")
    out.append('extern "C" {')
    for name, ret, args in funcs:
        rargs = ", ".join(f"{n}: {t}" for n, t in map(arg_to_rust, args))
        rret = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "constchar*": "*const c_char"}[ret]
        out.append(f"    pub fn {name}({rargs}) -> {rret};")
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    stub = generate()
    if "--update" in sys.argv:
        doc = open(DOC, encoding="utf-8").read()
        a, b = doc.index(BEGIN), doc.index(END)
        doc = doc[:a] + BEGIN + "\n```rust\n" + stub + "```\n" + doc[b:]
        open(DOC, "w", encoding="utf-8").write(doc)
        print(f"updated {DOC}")
    else:
        sys.stdout.write(stub)


if __name__ == "__main__":
    main()
