"""The C-ABI library loads without a GPU and exports every symbol include/fsr1_b200.h declares; argument
validation is exercised through paths that return before any CUDA call."""
import ctypes
import os
import re

import fsr1_b200 as F
from fsr1_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fsr1_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsr1_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert sorted(_lib.SYMBOLS) == names
    assert L.fsr1_abi_version() == 3


def test_error_strings_and_validation_without_gpu():
    L = _lib.lib()
    assert L.fsr1_error_string(0) == b"ok"
    assert b"invalid" in L.fsr1_error_string(-1)
    img = _lib.Image(0, 0, 0, 0, 0, 0, 1, 0)  # null data
    con = (ctypes.c_uint32 * 16)()
    assert L.fsr1_easu(ctypes.byref(img), ctypes.byref(img), con, 0, 0, 0, None) == -1
    assert L.fsr1_rcas(ctypes.byref(img), ctypes.byref(img), con, 0, 0, 0, None) == -1
    assert L.fsr1_upscale(None, None, None, con, con, 0, 0, 0, None) == -1
    buf = (ctypes.c_uint8 * 4096)()
    addr = ctypes.addressof(buf)
    addr += (-addr) % 16
    ok_in = _lib.Image(addr, 64, 8, 4, 0, 4, 1, 0)
    bad_fmt = _lib.Image(addr, 64, 8, 4, 0, 4, 9, 0)
    assert L.fsr1_easu(ctypes.byref(ok_in), ctypes.byref(bad_fmt), con, 0, 0, 0, None) == -1
    f32_out = _lib.Image(addr, 256, 16, 8, 0, 8, 2, 0)
    assert L.fsr1_easu(ctypes.byref(ok_in), ctypes.byref(f32_out), con, 0, 0, 0, None) == -2  # mixed formats
    assert L.fsr1_easu(ctypes.byref(ok_in), ctypes.byref(ok_in), con, 0, 0, 1 << 20, None) == -1  # unknown flag
    assert L.fsr1_rcas(ctypes.byref(ok_in), ctypes.byref(ok_in), con, 0, 0, 1 << 20, None) == -1
    assert L.fsr1_rcas(ctypes.byref(ok_in), ctypes.byref(ok_in), con, 0, 0, 0, None) == -1      # in place: RCAS reads neighbours it would overwrite
    # sharding: argument validation happens before any CUDA call
    h = ctypes.c_void_p()
    assert L.fsr1_shard_create(ctypes.byref(h), 64, 8, 128, 16, 1, 16, 0, 1, ctypes.c_float(0.25), 0) == -1   # more ranks than rows
    assert L.fsr1_shard_create(ctypes.byref(h), 64, 64, 128, 128, 1, 4, 4, 1, ctypes.c_float(0.25), 0) == -1  # rank >= world
    assert L.fsr1_shard_create(ctypes.byref(h), 64, 64, 128, 128, 9, 4, 0, 1, ctypes.c_float(0.25), 0) == -1  # unknown format
    assert L.fsr1_shard_submit(None, 0, None) == -1 and L.fsr1_shard_status(None) == -1
    assert b"neighbour" in L.fsr1_error_string(-6)
    small_pitch = _lib.Image(addr, 32, 8, 4, 0, 4, 1, 0)
    assert L.fsr1_easu(ctypes.byref(small_pitch), ctypes.byref(ok_in), con, 0, 0, 0, None) == -1
    # window that does not hold the rows EASU would read -> FSR1_ERR_WINDOW, before any launch
    econ = (ctypes.c_uint32 * 16)(*F.api.easu_con(8, 64, 8, 64, 16, 128))
    win_in = _lib.Image(addr, 64, 8, 64, 10, 4, 1, 0)
    out = _lib.Image(addr, 128, 16, 128, 0, 128, 1, 0)
    assert L.fsr1_easu(ctypes.byref(win_in), ctypes.byref(out), econ, 0, 0, 0, None) == -3
    a, b = ctypes.c_uint32(), ctypes.c_uint32()
    assert L.fsr1_easu_input_rows(econ, 64, 20, 40, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == (8, 21)  # rows floor((20+.5)/2-.5)-1 .. floor((39+.5)/2-.5)+2
    # pointwise companions: same validation rules, all decided before any launch
    h_img = _lib.Image(addr, 64, 8, 4, 0, 4, 1, 0)
    u8_img = _lib.Image(addr, 32, 8, 4, 0, 4, 3, 0)
    u10_img = _lib.Image(addr, 32, 8, 4, 0, 4, 4, 0)
    other = _lib.Image(addr, 64, 8, 5, 0, 5, 1, 0)
    assert L.fsr1_srtm(None, ctypes.byref(h_img), 0, 0, 0, None) == -1
    assert L.fsr1_srtm(ctypes.byref(h_img), ctypes.byref(other), 0, 0, 0, None) == -1        # sizes differ
    assert L.fsr1_srtm(ctypes.byref(h_img), ctypes.byref(u8_img), 0, 0, 0, None) == -2       # SRTM does not convert formats
    assert L.fsr1_srtm(ctypes.byref(h_img), ctypes.byref(h_img), 1, 3, 2, None) == -1        # empty row range
    assert L.fsr1_lfga(ctypes.byref(h_img), None, ctypes.byref(h_img), 0.5, 0, 0, None) == -1
    assert L.fsr1_lfga(ctypes.byref(h_img), ctypes.byref(u8_img), ctypes.byref(h_img), 0.5, 0, 0, None) == -2  # grain is signed
    assert L.fsr1_tepd(ctypes.byref(h_img), None, ctypes.byref(h_img), 9, 0, 0, 0, None) == -1   # bits must be 8 or 10
    assert L.fsr1_tepd(ctypes.byref(h_img), None, ctypes.byref(u10_img), 8, 0, 0, 0, None) == -2  # 8-bit codes into RGB10A2
    win = _lib.Image(addr, 64, 8, 64, 10, 4, 1, 0)
    assert L.fsr1_srtm(ctypes.byref(win), ctypes.byref(win), 0, 0, 0, None) == -3            # window lacks rows [0,64)
    # the half-precision (H / Hx2) forms: RGBA16F everywhere, the same rules otherwise
    f32_img = _lib.Image(addr, 128, 8, 4, 0, 4, 2, 0)
    assert L.fsr1_srtm_h(None, ctypes.byref(h_img), 0, 0, 0, None) == -1
    assert L.fsr1_srtm_h(ctypes.byref(f32_img), ctypes.byref(f32_img), 0, 0, 0, None) == -2   # half images only
    assert L.fsr1_srtm_h(ctypes.byref(h_img), ctypes.byref(u8_img), 0, 0, 0, None) == -2
    assert L.fsr1_srtm_h(ctypes.byref(h_img), ctypes.byref(other), 0, 0, 0, None) == -1
    assert L.fsr1_lfga_h(ctypes.byref(h_img), None, ctypes.byref(h_img), 0.5, 0, 0, None) == -1
    assert L.fsr1_lfga_h(ctypes.byref(h_img), ctypes.byref(f32_img), ctypes.byref(h_img), 0.5, 0, 0, None) == -2   # the grain tile too
    assert L.fsr1_tepd_h(ctypes.byref(h_img), None, ctypes.byref(h_img), 9, 0, 0, 0, None) == -1
    assert L.fsr1_tepd_h(ctypes.byref(h_img), None, ctypes.byref(u8_img), 8, 0, 0, 0, None) == -2   # writes RGBA16F (fsr1_tepd writes code values)
    assert L.fsr1_srtm_h(ctypes.byref(win), ctypes.byref(win), 0, 0, 0, None) == -3
    # FSR1_FLAG_RCAS_HX2 (1 << 10): a known flag, half images only
    rc4 = (ctypes.c_uint32 * 4)()
    f32_a, f32_b = _lib.Image(addr, 128, 8, 4, 0, 4, 2, 0), _lib.Image(addr + 2048, 128, 8, 4, 0, 4, 2, 0)
    assert L.fsr1_rcas(ctypes.byref(f32_a), ctypes.byref(f32_b), rc4, 0, 0, 1 << 10, None) == -2
    assert L.fsr1_launch_count() == 0


def test_cpp_filter_mirror_compiles_and_links(tmp_path):
    """fidelityfx-fsr_b200/fsr_filter.hpp: the C++ twin of the reference's FSR_Filter over the C ABI."""
    import subprocess
    src = tmp_path / "f.cpp"
    src.write_text('#include "fidelityfx-fsr_b200/fsr_filter.hpp"\n'
                   'int main(){ fsr1::FSR_Filter f; f.OnCreate(); fsr1::State s; s.renderWidth = 8;\n'
                   '  AU1 c[16]; FsrEasuCon(c, c+4, c+8, c+12, 8.f, 8.f, 8.f, 8.f, 16.f, 16.f);\n'
                   '  return (fsr1_abi_version() == 3 && c[0] == 0x3f000000u) ? 0 : 1; }\n')
    exe = tmp_path / "f"
    libdir = os.path.join(ROOT, "fidelityfx-fsr_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", ROOT, str(src), "-o", str(exe), "-L", libdir, "-lfsr1_b200",
                           "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0


def test_easu_input_rows_against_brute_force():
    """fsr1_easu_input_rows (host logic, no GPU): first/last input row of an output row range = min/max tap row over
    its pixels, in the kernels' float arithmetic, for the quality presets and random dynamic-resolution viewports."""
    import numpy as np
    L = _lib.lib()
    rng = np.random.default_rng(11)
    cases = [(1080, 2160), (1440, 2160), (1661, 2160), (1270, 2160), (2160, 4320), (17, 31), (5, 40), (64, 64)]
    cases += [(int(rng.integers(4, 3000)), 0) for _ in range(40)]
    for in_h, out_h in cases:
        if out_h == 0:
            out_h = int(in_h * rng.uniform(1.0, 2.0))
        vp_h = in_h if rng.random() < 0.5 else int(in_h * rng.uniform(0.5, 1.0)) or 1
        con = (ctypes.c_uint32 * 16)(*F.api.easu_con(64, vp_h, 64, in_h, 128, out_h))
        scale = np.array([con[1]], np.uint32).view(np.float32)[0]
        off = np.array([con[3]], np.uint32).view(np.float32)[0]
        for _ in range(6):
            y0 = int(rng.integers(0, out_h))
            y1 = int(rng.integers(y0 + 1, out_h + 1))
            cells = np.floor((np.arange(y0, y1, dtype=np.float32) * scale).astype(np.float32) + off).astype(np.int64)
            lo = int(np.clip(cells.min() - 1, 0, in_h - 1))
            hi = int(np.clip(cells.max() + 2, 0, in_h - 1))
            a, b = ctypes.c_uint32(), ctypes.c_uint32()
            assert L.fsr1_easu_input_rows(con, in_h, y0, y1, ctypes.byref(a), ctypes.byref(b)) == 0
            assert (a.value, b.value) == (lo, hi), (in_h, out_h, vp_h, y0, y1)


def _build_c_demo(tmp_path):
    import subprocess
    exe = tmp_path / "fsr1_demo"
    libdir = os.path.join(ROOT, "fidelityfx-fsr_b200", "lib")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(ROOT, "examples", "fsr1_demo.c"), "-o", str(exe), "-L", libdir, "-lfsr1_b200",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_plain_c_host_program_builds_against_the_abi(tmp_path):
    """examples/fsr1_demo.c: C99, the reference's include lines (served by include/compat), one fsr1_upscale call."""
    import subprocess
    import torch
    exe = _build_c_demo(tmp_path)
    if not torch.cuda.is_available():
        assert subprocess.call([str(exe)], stderr=subprocess.DEVNULL) == 2      # "no CUDA device": loud, no fallback
