"""CPU tests (no GPU): the C-ABI shared library loads and exports every symbol include/*.h declares, the boundary types
have the reference's layout, and the model loader reproduces the reference's error contract (the negative matrix of
test/validation_suite/run_tests_avxout.sh:109-165 + create_wrong_files.sh)."""
import ctypes as C
import os
import re
import shutil

import pytest

import raisr_testlib as T

LIB = T.product_lib_path()
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="libraisr.so not built (python -c 'import __graft_entry__ as g; g.build()')")


def declared_symbols():
    names = set()
    inc = os.path.join(T.ROOT, "include")
    for path in (os.path.join(inc, "raisr_cuda.h"), os.path.join(inc, "raisr", "RaisrHandler.h")):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names |= set(re.findall(r"\b(raisr_cuda_\w+|RNLHandler_\w+)\s*\(", text))
    return sorted(names)


def test_every_declared_symbol_is_exported():
    L = C.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 14, syms
    for s in syms:
        assert hasattr(L, s), s


def test_cxx_api_symbols_exported():
    """Raisr.h's C++ functions (RNLInit, RNLSetRes, RNLProcess, RNLSetOpenCLContext, RNLDeinit) are exported mangled."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", LIB], capture_output=True, text=True).stdout
    for s in ("RNLInit(", "RNLSetRes(", "RNLProcess(", "RNLSetOpenCLContext(", "RNLDeinit()"):
        assert s in out, s


def test_boundary_types():
    # RaisrDefaults.h:13-20: pointer + 4 x unsigned
    assert C.sizeof(T.VideoDataType) == 24
    assert T.VideoDataType.step.offset == 16 and T.VideoDataType.bitShift.offset == 20
    assert T.RNLErrorBadParameter & 0xffffffff == 0x80001002 and T.RNLErrorUndefined & 0xffffffff == 0x80001001


def test_version_string():
    L = C.CDLL(LIB)
    L.raisr_cuda_version.restype = C.c_char_p
    assert b"23.11" in L.raisr_cuda_version()


needs_filters = pytest.mark.skipif(not T.have_filters(), reason="trained filter folders not staged")


def init(folder, ratio=2.0, bits=8, passes=1, mode=1, capfd=None):
    L = T.handler_lib(LIB)
    rc = L.RNLHandler_Init(folder.encode(), ratio, bits, T.VideoRange, 20, T.AVX512, passes, mode)
    L.RNLHandler_Deinit()
    return rc


def no_gpu():
    import torch
    return not torch.cuda.is_available()


@needs_filters
def test_good_model_reaches_the_device_check(capfd):
    """A valid model passes every loader check; without a GPU the engine then refuses to run (no CPU path)."""
    rc = init(T.filter_folder("filters_2x/filters_lowres"))
    out = capfd.readouterr().out
    if no_gpu():
        assert rc == T.RNLErrorUndefined and "no CUDA device" in out
    else:
        assert rc == 0


@needs_filters
@pytest.mark.parametrize("kw,needle", [
    (dict(bits=9), "bits is NOT supported"),                       # run_tests_avxout.sh:112
    (dict(passes=3), "Only support passes 1 or 2"),                # :127
    (dict(ratio=1.5), "number of pixel types"),                    # 2x table at ratio 1.5 (:116-118)
    (dict(bits=16), "Unable to load model"),                       # no 16-bit table is shipped
    (dict(passes=2, bits=10), "Unable to load model"),             # filters_lowres ships no 10-bit second-pass table
])
def test_bad_parameters(kw, needle, capfd):
    rc = init(T.filter_folder("filters_2x/filters_lowres"), **kw)
    out = capfd.readouterr().out
    assert rc != 0 and needle in out, out


@needs_filters
def test_one_pass_mode2_warns(capfd):
    init(T.filter_folder("filters_2x/filters_lowres"), passes=1, mode=2)
    assert "[RAISR WARNING] 1 pass with upscale in 2d pass, mode = 2 ignored" in capfd.readouterr().out   # Raisr.cpp:1434-1435


@needs_filters
@pytest.mark.parametrize("line,needle", [
    ("12 3 3 11", "number of hash keys"),      # create_wrong_files.sh: wrong angle count
    ("24 3 3", "configFile corrupted"),        # three tokens
    ("24 3 3 6", "configFile corrupted"),      # even patch
    ("24 3 3 9", "configFile corrupted"),      # patch != 11
    ("24 x 3 11", "configFile corrupted"),     # not a number
])
def test_corrupted_config(tmp_path, line, needle, capfd):
    src = T.filter_folder("filters_2x/filters_highres")
    dst = tmp_path / "model"
    shutil.copytree(src, dst)
    (dst / "config").write_text(line + "\n")
    rc = init(str(dst))
    assert rc == T.RNLErrorBadParameter and needle in capfd.readouterr().out


@needs_filters
@pytest.mark.parametrize("victim,needle", [
    ("filterbin_2_8", "Unable to load model"),
    ("Qfactor_strbin_2_8", "Unable to load model"),
    ("Qfactor_cohbin_2_8", "Unable to load model"),
    ("config", "Unable to open config file"),
])
def test_missing_files(tmp_path, victim, needle, capfd):
    dst = tmp_path / "model"
    shutil.copytree(T.filter_folder("filters_2x/filters_highres"), dst)
    os.remove(dst / victim)
    rc = init(str(dst))
    assert rc == T.RNLErrorBadParameter and needle in capfd.readouterr().out


@needs_filters
def test_corrupted_tables(tmp_path, capfd):
    dst = tmp_path / "model"
    shutil.copytree(T.filter_folder("filters_2x/filters_highres"), dst)
    raw = (dst / "filterbin_2_8").read_bytes()
    (dst / "filterbin_2_8").write_bytes(raw[:-8])                       # truncated table (size check, Raisr.cpp:294)
    assert init(str(dst)) == T.RNLErrorBadParameter and "hashtable corrupted" in capfd.readouterr().out
    (dst / "filterbin_2_8").write_bytes(b"fp64" + raw[4:])              # unknown tag (Raisr.cpp:279-282)
    assert init(str(dst)) == T.RNLErrorBadParameter and "hashtable corrupted" in capfd.readouterr().out
    (dst / "filterbin_2_8").write_bytes(raw)
    (dst / "Qfactor_strbin_2_8").write_text("0.5\n0.7\n0.9\n")          # three thresholds (Raisr.cpp:389-393)
    assert init(str(dst)) == T.RNLErrorBadParameter and "StrFile corrupted" in capfd.readouterr().out
    (dst / "Qfactor_strbin_2_8").write_text("0.5\n1e-3\n")              # exponent char rejected by the whitelist (Raisr.cpp:187-211)
    assert init(str(dst)) == T.RNLErrorBadParameter and "StrFile corrupted" in capfd.readouterr().out


def test_process_null_planes():
    """Null plane pointers -> RNLErrorBadParameter (Raisr.cpp:1297-1299,1358)."""
    L = T.handler_lib(LIB)
    v = T.VideoDataType()
    P = C.byref(v)
    assert L.RNLHandler_Process(P, P, P, P, P, P, T.CountOfBitsChanged) == T.RNLErrorBadParameter
    assert L.RNLHandler_Process(None, None, None, None, None, None, T.CountOfBitsChanged) == T.RNLErrorBadParameter


def test_sliding_window_schedule_and_tree_table():
    """Stage D of the pipelined kernel (csrc/raisr_pipe_kernel.cuh) walks down pixel columns with rotating chain ownership.  Its
    schedule and folded lane tree are verified symbolically against the reference's chain / tree order (DotProdPatch_AVX512_32f,
    Raisr_AVX512.cpp:134-149, sumitup_ps_512 :37-44) by tools/slide_model.py; the shuffle-source table in the kernel header must be the
    one that model produces."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("slide_model", os.path.join(root, "tools", "slide_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.check_macs(48)
    words = m.tree_tables()
    m.check_tree(words)
    src = open(os.path.join(root, "video-super-resolution-library_b200", "csrc", "raisr_pipe_kernel.cuh")).read()
    tbl = re.search(r"c_slide_tbl\[8\] = \{([^}]*)\}", src).group(1)
    assert [int(x.strip().rstrip("u"), 16) for x in tbl.split(",")] == words


def test_host_copy_rows_of_the_pageable_path_equals_memcpy(tmp_path):
    """csrc/raisr_hostcopy.cpp (streaming-store row copies between the caller's pageable planes and the staging planes): 400 random
    sizes / alignments / strides against memcpy(), with guard bytes, in whichever mode this CPU selects and in the memcpy mode."""
    import subprocess
    src = os.path.join(T.PKG_DIR, "csrc")
    exe = str(tmp_path / "hostcopy_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + src, os.path.join(T.ROOT, "tests", "harness", "hostcopy_check.cpp"),
                           os.path.join(src, "raisr_hostcopy.cpp"), "-o", exe])
    for env in ({}, {"RAISR_CUDA_NT_COPY": "0"}):
        r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
        if env:
            assert r.stdout.startswith("mode 0")
