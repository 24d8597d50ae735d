"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every
symbol include/rama_b200.h declares; error paths do not need a device."""
import ctypes as C
import os
import re

import pytest

from rama_b200 import _lib
from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "rama_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rama_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_binding_uses():
    assert _declared() == _lib.EXPORTS


def test_library_loads_and_exports_every_symbol():
    assert os.path.exists(_lib.SO_PATH), "run __graft_entry__.build() first"
    L = C.CDLL(_lib.SO_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    assert _lib.lib().rama_abi_version() == 1


def test_header_is_plain_c_and_the_static_archive_links_from_c(tmp_path):
    """The boundary is a C ABI: include/rama_b200.h must compile as C99 (no C++ in the signatures — what a cgo / Rust-FFI /
    ctypes binding relies on), and a C program must link against the static archive the Rust build.rs would use."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "rama_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "use.c"
    src.write_text('#include "rama_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { int n = -1; int rc = rama_device_count(&n);\n'
                   '  printf("abi %d rc %d n %d err %s\\n", rama_abi_version(), rc, n, rc ? rama_last_error() : "-"); return 0; }\n')
    exe = tmp_path / "use"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L", os.path.join(ROOT, "rama_b200"), "-lrama_b200", "-Wl,-rpath," + os.path.join(ROOT, "rama_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("abi 1 rc ")   # without a GPU: rc = RAMA_E_CUDA and a message, no crash


def test_static_archive_present():
    assert os.path.exists(os.path.join(ROOT, "rama_b200", "librama_b200.a"))


def test_null_arguments_are_errors_not_crashes():
    L = _lib.lib()
    assert L.rama_ctx_create(0, None, None) == -1
    assert b"NULL" in L.rama_last_error()
    assert L.rama_session_create(None, None) == -1
    assert L.rama_forward(None, 0, 0) == -1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = _lib.lib().rama_ctx_create(0, None, C.byref(h))
    assert rc == -2  # RAMA_E_CUDA: fails loudly, nothing runs on the CPU
    with pytest.raises(_lib.RamaError):
        _lib.check(rc)
