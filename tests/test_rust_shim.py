"""The Rust `-sys` crate (rust/zelll-b200-sys/src/lib.rs) against include/zelll_b200.h: same symbol set,
same argument counts and types, same struct fields in the same order, same enum values.  The crates
cannot be compiled here (no cargo/rustc in the image), so this is what keeps them from drifting."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "zelll_b200.h")
SYS = os.path.join(ROOT, "rust", "zelll-b200-sys", "src", "lib.rs")
WRAP = os.path.join(ROOT, "rust", "zelll-b200", "src", "lib.rs")

C2RUST = {
    "int": "c_int", "void": None, "uint64_t": "u64", "int64_t": "i64", "uint32_t": "u32", "int32_t": "i32",
    "uint8_t": "u8", "double": "f64", "char": "c_char", "zb_grid": "zb_grid", "zb_info": "zb_info",
    "zb_slab_info": "zb_slab_info",
}


def _strip_c(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def _c_type_to_rust(ctype):
    ctype = ctype.strip()
    const = "const " in ctype + " " or ctype.startswith("const")
    base = ctype.replace("const", "").replace("struct", "").strip()
    stars = base.count("*")
    base = base.replace("*", "").strip()
    rust = C2RUST[base]
    if stars == 0:
        return rust
    inner = "c_void" if rust is None else rust
    out = inner
    for level in range(stars):
        innermost = level == 0
        out = ("*const " if (const and innermost) else "*mut ") + out
    return out


def c_functions():
    text = _strip_c(open(HEADER).read())
    fns = {}
    for m in re.finditer(r"\b((?:const\s+)?[a-z_0-9]+\s*\**)\s*(zb_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                pm = re.match(r"(.*?)([A-Za-z_0-9]+)$", a)
                params.append((pm.group(2), _c_type_to_rust(pm.group(1))))
        fns[name] = (_c_type_to_rust(ret), params)
    return fns


def rust_functions():
    text = re.sub(r"//[^\n]*", " ", open(SYS).read())
    block = re.search(r'extern "C" \{(.*)\}', text, flags=re.S).group(1)
    fns = {}
    for m in re.finditer(r"pub fn (zb_[a-z_0-9]+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        name, args, ret = m.group(1), m.group(2).strip(), m.group(3)
        params = []
        if args:
            for a in args.split(","):
                pname, ptype = a.split(":", 1)
                params.append((pname.strip(), ptype.strip()))
        fns[name] = (ret.strip() if ret else None, params)
    return fns


def c_structs():
    text = _strip_c(open(HEADER).read())
    out = {}
    for m in re.finditer(r"typedef struct (zb_[a-z_]+) \{(.*?)\} \1;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ctype, rest = decl.split(None, 1)
            for item in rest.split(","):
                fm = re.match(r"\s*([A-Za-z_0-9]+)(?:\[(\d+)\])?", item)
                rust = C2RUST[ctype]
                fields.append((fm.group(1), f"[{rust}; {fm.group(2)}]" if fm.group(2) else rust))
        out[m.group(1)] = fields
    return out


def rust_structs():
    text = re.sub(r"//[^\n]*", " ", open(SYS).read())
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{]*?pub struct (zb_[a-z_]+) \{(.*?)\}", text, flags=re.S):
        fields = [(f.group(1), f.group(2).strip()) for f in re.finditer(r"pub ([a-z_0-9]+):\s*([^,\n]+),", m.group(2))]
        out[m.group(1)] = fields
    return out


def test_extern_block_matches_header():
    c, r = c_functions(), rust_functions()
    assert len(c) >= 30
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name in c:
        cret, cparams = c[name]
        rret, rparams = r[name]
        assert cret == rret, (name, cret, rret)
        assert [p[0] for p in cparams] == [p[0] for p in rparams], name
        assert [p[1] for p in cparams] == [p[1] for p in rparams], (name, cparams, rparams)


def test_repr_c_structs_match_header():
    c, r = c_structs(), rust_structs()
    assert set(c) == {"zb_info", "zb_slab_info"}
    for name, fields in c.items():
        assert r[name] == fields, (name, r[name], fields)


def test_enum_constants_match_header():
    header = _strip_c(open(HEADER).read())
    rust = open(SYS).read()
    consts = dict(re.findall(r"\b(ZB_[A-Z0-9_]+)\s*=\s*(\d+)", header))
    consts["ZB_ABI_VERSION"] = re.search(r"#define ZB_ABI_VERSION (\d+)", header).group(1)
    assert len(consts) > 20
    for name, value in consts.items():
        m = re.search(rf"pub const {name}: [a-z_]+ = (\d+);", rust)
        assert m and m.group(1) == value, name


def test_wrapper_mirrors_the_reference_surface():
    """CellGrid::{new, rebuild, rebuild_mut, particle_pairs, par_particle_pairs, info, query_neighbors}
    (src/cellgrid.rs:166-451) exist in the wrapper crate and call only symbols the -sys crate declares."""
    text = open(WRAP).read()
    for fn in ("pub fn new<", "pub fn rebuild<", "pub fn rebuild_mut<", "pub fn particle_pairs(", "pub fn par_particle_pairs(",
               "pub fn info(", "pub fn query_neighbors<", "pub fn cell_storage("):
        assert fn in text, fn
    used = set(re.findall(r"sys::(zb_[a-z_0-9]+)\(", text))
    assert used and used <= set(rust_functions()), used - set(rust_functions())
