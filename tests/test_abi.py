"""The C-ABI library loads, exports every symbol include/dfb200.h declares, and the ctypes
signatures in dformer_b200/_lib.py agree with the header (no GPU needed, no compute calls)."""
import ctypes
import os
import re

from dformer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    h = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    h = re.sub(r"typedef struct .*?\}\s*\w+;", "", h, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long|const char\*)\s+(dfb200_\w+)\s*\((.*?)\)\s*;", h, flags=re.S):
        out[m.group(1)] = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
    return out


def _ctype_of(decl):
    if "*" in decl:
        return "P"
    t = decl.split()
    if "long" in t:
        return ctypes.c_long
    if "double" in t:
        return ctypes.c_double
    if "float" in t:
        return ctypes.c_float
    if "int" in t:
        return ctypes.c_int
    raise AssertionError(decl)


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.lib()
    assert lib.version() >= 100
    decls = _header_decls()
    assert len(decls) >= 38
    for name in decls:
        assert hasattr(lib.cdll, name), name


def test_ctypes_signatures_match_header():
    decls = _header_decls()
    for name, args in L.SIGNATURES.items():
        hargs = decls[name]
        assert len(hargs) == len(args), (name, len(hargs), len(args))
        for i, (h, a) in enumerate(zip(hargs, args)):
            want = _ctype_of(h)
            if want == "P":
                assert a is ctypes.c_void_p or issubclass(a, ctypes._Pointer), (name, i, h)
            else:
                assert a is want, (name, i, h, a)
    assert set(decls) - {"dfb200_last_error", "dfb200_version", "dfb200_launch_count"} == set(L.SIGNATURES)


def test_struct_layout_matches_header():
    h = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    body = re.search(r"typedef struct dfb200_gemm_args \{(.*?)\} dfb200_gemm_args;", h, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            names.append(part.replace("*", " ").split()[-1])
    assert names == [f[0] for f in L.GemmArgs._fields_]


def test_only_the_checkers_touch_the_oracle_and_nothing_shipped_reads_the_reference_tree():
    """Task rule: only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline / reference legs may import `oracle/`; nothing
    that runs on the GPU box (package, tools, bench, smoke) may read /root/reference."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    product = glob.glob(os.path.join(root, "dformer_b200", "**", "*.py"), recursive=True) + glob.glob(os.path.join(root, "tools", "*.py"))
    assert len(product) > 15
    imp = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for f in product:
        src = open(f).read()
        assert not imp.search(src), f"{f} imports the oracle"
        assert "/root/reference" not in src, f"{f} names the reference tree"
    for f in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(root, f)).read(), f
