"""CPU checks of the drop-in boundary: librender.so loads, exports every symbol
include/vtrace_abi.h declares (and nothing of the oracle), struct layouts match the Rust side,
and the product fails loudly — not silently on a CPU path — when no CUDA device is present.
No compute call is made without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from vtrace_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vtrace_abi.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_\s\*]*?\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return [n for n in names if n not in ("defined",)]


def test_header_declares_the_seven_reference_symbols():
    names = declared_functions()
    for ref in ("entry", "render_tick", "get_input_data_pointer", "add_texture", "start_update_instances",
                "end_update_instances", "cleanup"):  # src/render.rs:110-128
        assert ref in names
    assert len(names) == len(set(names)) >= 20


def test_library_exports_every_declared_symbol():
    lib = abi.load()
    names = declared_functions()
    assert set(names) == set(abi.SYMBOLS), set(names) ^ set(abi.SYMBOLS)
    for n in names:
        assert getattr(lib, n) is not None


def test_library_is_not_linked_to_the_oracle():
    out = subprocess.run(["nm", "-D", abi.library_path()], capture_output=True, text=True, check=True).stdout
    assert " vo_" not in out and "vtrace_oracle" not in out
    ldd = subprocess.run(["ldd", abi.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    # and the product package never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vtrace_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f   # comments may cite it; code may not use it
                assert not re.search(r"^\s*(import|from)\s+\S*oracle", text, flags=re.M), f
                assert "oracle_lib" not in text and "CDLL(" not in text.replace("C.CDLL(path)", ""), f


def test_struct_layouts_match_the_rust_side():
    assert C.sizeof(abi.UserInput) == 40            # src/render.rs:37-51
    assert abi.UserInput.mouse_x.offset == 8 and abi.UserInput.last_mouse_y.offset == 32
    assert C.sizeof(abi.RenderTickInfo) == 2 * C.sizeof(C.c_void_p)  # src/render.rs:177-181
    assert abi.HIT_DTYPE.itemsize == 16
    assert C.sizeof(abi.VtConfig) == 48 and C.sizeof(abi.VtStats) == 88


def test_sass_is_sm100a_with_tma_bulk_copy():
    out = subprocess.run(["cuobjdump", "-sass", abi.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "UBLKCP" in out.stdout, "TMA bulk copy (cp.async.bulk) missing from the trace kernels"
    assert "FMNMX3" in out.stdout


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    code = lib.entry()
    assert code != 0 and "CUDA" in abi.last_error()
    assert lib.add_texture(None, 1, 1, 1) == -1
    assert not lib.start_update_instances(1)
    assert lib.render_tick(None, None, None) == -1
    from vtrace_b200.renderer import Renderer
    with pytest.raises(RuntimeError):
        Renderer()


def test_static_library_defines_the_seven_symbols_and_links_into_a_c_host(tmp_path):
    """librender.a is what the reference's build script links (`cargo:rustc-link-lib=static=render`,
    build.rs:96-97): a plain C host must link against it with only the CUDA and C++ runtimes added."""
    abi.load()
    from vtrace_b200 import build
    assert os.path.exists(build.STATIC_LIB)
    defined = subprocess.run(["nm", "-g", "--defined-only", build.STATIC_LIB], capture_output=True, text=True, check=True).stdout
    refs = ("entry", "render_tick", "get_input_data_pointer", "add_texture", "start_update_instances", "end_update_instances", "cleanup")
    for ref in refs:
        assert re.search(rf"\bT {ref}\b", defined), ref
    src = tmp_path / "host.c"
    src.write_text('''#include "vtrace_abi.h"
#include <stdio.h>
int main(void) {
    uint64_t rc = entry();
    if (rc) { printf("entry failed: %s\\n", vt_last_error()); return 3; }
    float P[16] = {0}, V[16] = {0}; render_tick_info info = {P, V}; int32_t w = 0, h = 0;
    (void)get_input_data_pointer();
    float* m = start_update_instances(1); if (!m || end_update_instances(1)) return 4;
    (void)add_texture(0, 0, 0, 0); (void)render_tick(&w, &h, &info); cleanup(); return 0;
}
''')
    exe = tmp_path / "host_static"
    cuda_lib = os.environ.get("CUDA_LIB", "/usr/local/cuda/lib64")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", os.path.dirname(build.STATIC_LIB),
                    "-l:librender.a", "-L", cuda_lib, "-lcudart_static", "-lstdc++", "-lm", "-ldl", "-lpthread", "-lrt"],
                   check=True, capture_output=True)
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "librender" not in ldd  # nothing of the renderer is left to the dynamic loader
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert out.returncode == 0, out.stdout + out.stderr
    else:  # fails loudly, through the statically linked entry()
        assert out.returncode == 3 and "CUDA" in out.stdout


def test_compiled_host_builds_and_links_the_seven_symbols():
    """host/vtrace_headless.cpp (C++ mirror of the Rust engine's host side) links against librender.so."""
    abi.load()  # makes sure librender.so exists
    subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
    exe = os.path.join(ROOT, "host", "vtrace_headless")
    undefined = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    for ref in ("entry", "render_tick", "get_input_data_pointer", "add_texture", "start_update_instances",
                "end_update_instances", "cleanup"):
        assert re.search(rf"\bU {ref}\b", undefined), ref
