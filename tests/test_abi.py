"""The C ABI: header <-> Python layout constants, and liblls.so exports every declared symbol (no GPU calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lls.h")


def _header():
    with open(HEADER) as fh:
        return fh.read()


def _enum_values(text):
    vals = {}
    for m in re.finditer(r"\b(LLS_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        vals[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"#define\s+(LLS_[A-Z0-9_]+)\s+(\d+)\b", text):
        vals[m.group(1)] = int(m.group(2))
    return vals


def test_layout_matches_header():
    from machupx_b200 import layout as L
    vals = _enum_values(_header())
    prefixes = {"F_": "LLS_F_", "I_": "LLS_I_", "AF_": "LLS_AF_", "AC_": "LLS_AC_", "O_": "LLS_O_",
                "FLAG_": "LLS_FLAG_", "STATUS_": "LLS_STATUS_"}
    checked = 0
    for name in dir(L):
        for short, full in prefixes.items():
            if name.startswith(short) and name != "FLAG_DEFAULT":
                key = full + name[len(short):]
                assert key in vals, key
                assert vals[key] == getattr(L, name), (key, vals[key], getattr(L, name))
                checked += 1
    assert checked > 80
    assert vals["LLS_NF"] == L.NF and vals["LLS_NI"] == L.NI and vals["LLS_NO"] == L.NO


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", _header(), flags=re.S)
    return sorted(set(re.findall(r"\b(lls_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from machupx_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "liblls.so does not export " + n
    # and the ctypes binding covers exactly the header
    assert sorted(_native.SIGNATURES) == names
    lib2 = _native.load()
    assert lib2.lls_abi_version() == _native.ABI_VERSION


def test_sizes_query_is_pure_host():
    import ctypes
    from machupx_b200 import _native, layout as L
    lib = _native.load()
    s = _native.Sizes()
    _native.check(lib.lls_query_sizes(4000, 5, ctypes.byref(s)))
    assert s.ld == L.padded_ld(4000) == 4032
    assert s.vji_bytes == 4000 * 3 * 4032 * 8 and s.mat_bytes == 4032 * 4032 * 8
    assert s.tab_bytes == L.NF * 4032 * 8 and s.out_bytes == L.NO * 4032 * 8
    with pytest.raises(_native.LlsError):
        _native.check(lib.lls_query_sizes(0, 1, ctypes.byref(s)))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under machupx_b200/ may reference it."""
    pkg = os.path.join(ROOT, "machupx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/lls.h must compile as C99 (no C++ in the declarations) and a C program must link
    against liblls.so and call it -- here only lls_abi_version(), which touches no device."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "lls.h"\n'
                   'int main(void) { lls_scene* h = 0; (void)h; return lls_abi_version() > 0 ? 0 : 1; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "machupx_b200")
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                           "-L", lib_dir, "-l:liblls.so", "-Wl,-rpath," + lib_dir, "-o", str(exe)])
    assert subprocess.run([str(exe)]).returncode == 0
