"""CPU-only: the C-ABI library builds, loads and exports every symbol include/gseg.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def declared_symbols(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gseg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(gseg):
    L = gseg.load()
    syms = declared_symbols(gseg.HEADER)
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), "libgseg.so does not export %s" % s


def test_version_and_strerror(gseg):
    L = gseg.load()
    assert L.gseg_version() == 200
    assert L.gseg_strerror(0) == b"ok"
    assert b"no CPU fallback" in L.gseg_strerror(-2)


def test_struct_layout_matches_header(gseg):
    assert C.sizeof(gseg.Params) == 32
    assert C.sizeof(gseg.RoundStat) == 56
    L = gseg.load()
    out = (C.c_int32 * 8)()
    L.gseg_abi_sizes.argtypes = [C.POINTER(C.c_int32), C.c_int]
    assert L.gseg_abi_sizes(out, 8) == 5
    mirrors = [gseg.Params, gseg.RoundStat, gseg.KernelTime, gseg.PoolJob, gseg.PoolResult]
    assert [out[i] for i in range(5)] == [C.sizeof(m) for m in mirrors]


def test_pool_and_new_entry_points_fail_cleanly_without_gpu(gseg):
    """Argument errors need no device; creation without a device is GSEG_E_CUDA, never a fallback."""
    import torch
    L = gseg.load()
    h = C.c_void_p()
    assert L.gseg_pool_create(C.byref(h), 0, 64, 64, 8, 0, 0) == -1          # no contexts
    assert L.gseg_pool_create(None, 0, 64, 64, 8, 4, 0) == -1
    assert L.gseg_pool_next(None, None) == -1 and L.gseg_pool_pending(None) == -1
    assert L.gseg_labels_ex(None, 0, None, 4, 0) == -1 and L.gseg_label_bytes(None, 0) < 0
    assert L.gseg_strip_record(None, 1, None, 0, None) == -1
    assert L.gseg_join_segment(None, None, 1, 0, 0, None, None, 4, 0, None, None) == -1
    assert L.gseg_set_dedup(None, 1, 0, 0, 0) == -1 and L.gseg_reserve(None, 1) == -1
    assert L.gseg_strerror(-9).decode().startswith("label type too narrow")
    if not torch.cuda.is_available():
        assert L.gseg_pool_create(C.byref(h), 0, 64, 64, 8, 2, 0) == -2


def test_no_gpu_fails_loudly(gseg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gseg.GsegError) as e:
        gseg.Segmenter(64, 64)
    assert "no CPU fallback" in str(e.value)


def test_jpeg_entry_points_without_gpu(gseg):
    """nvJPEG is loaded with dlopen (libgseg.so must not link it); gseg_jpeg_info is a header parse on the host and
    needs no device; the decode itself has no host path behind it (no context without a device); argument errors
    do not need a device."""
    import subprocess
    import torch
    L = gseg.load()
    w, h = C.c_int32(0), C.c_int32(0)
    assert L.gseg_jpeg_info(None, 0, C.byref(w), C.byref(h)) == -1
    assert L.gseg_segment_jpeg(None, b"x", 1, None, None, None) == -1
    assert L.gseg_jpeg_decode_async(None, b"x", 1, None, 0, None, None, None) == -1
    assert L.gseg_set_jpeg_backend(None, 0) == -1 and L.gseg_jpeg_backend_used(None) == -1
    assert L.gseg_input_rgb(None, None, 0) == -1
    assert L.gseg_strerror(-8).decode().startswith("optional dependency")
    deps = subprocess.run(["ldd", gseg.LIB_PATH], capture_output=True, text=True).stdout
    assert "nvjpeg" not in deps
    with pytest.raises(gseg.GsegError):  # SOI + a segment that runs past the end: not a JPEG
        gseg.jpeg_info(b"\xff\xd8\xff\xe0" + bytes(32))
    cv2 = pytest.importorskip("cv2")
    ok, enc = cv2.imencode(".jpg", np.zeros((37, 53, 3), np.uint8))
    assert gseg.jpeg_info(enc.tobytes()) == (53, 37)
    if not torch.cuda.is_available():
        with pytest.raises(gseg.GsegError) as e:
            gseg.Segmenter(64, 64)
        assert "no CPU fallback" in str(e.value)


def test_bad_create_args(gseg):
    L = gseg.load()
    h = C.c_void_p()
    assert L.gseg_create(C.byref(h), 0, 0, 10) == -1
    assert L.gseg_create(None, 0, 10, 10) == -1


def test_product_does_not_reference_oracle(gseg):
    """The product path must never route through the oracle."""
    pkg = os.path.dirname(gseg.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "liboracle" not in txt and "gseg_oracle" not in txt and "import oracle" not in txt, f


def test_cli_builds_and_fails_loudly_without_gpu(gseg, tmp_path):
    """The C++ host program is built with the library; without a CUDA device it must refuse, not fall back."""
    import subprocess
    assert os.path.exists(gseg.CLI_PATH)
    r = subprocess.run([gseg.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: gseg" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([gseg.CLI_PATH, "--synth", "64x48:1", "0.8", "300", "20", "-", str(tmp_path / "o.ppm")],
                           capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
