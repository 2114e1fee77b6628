"""Shared test helpers: synthetic frames, model-file reader, ctypes bindings to the CHECKERS
(oracle/_build/liboracle.so = C restatement, oracle/_ref/libraisr_ref*.so = the compiled reference)
and to the product's C ABI (libraisr.so).  Only tests/, bench.py's cpu_baseline/reference legs and
__graft_entry__.smoke() import this module.
"""
import ctypes as C
import os
import struct
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
PKG_DIR = os.path.join(ROOT, "video-super-resolution-library_b200")

RNLErrorNone = 0
RNLErrorBadParameter = C.c_int32(0x80001002).value
RNLErrorUndefined = C.c_int32(0x80001001).value
VideoRange, FullRange = 1, 2
AVX2, AVX512, AVX512_FP16 = 1, 2, 5
Randomness, CountOfBitsChanged = 1, 2


class VideoDataType(C.Structure):  # RaisrDefaults.h:13-20
    _fields_ = [("pData", C.c_void_p), ("width", C.c_uint), ("height", C.c_uint),
                ("step", C.c_uint), ("bitShift", C.c_uint)]


DATA_DIR = os.path.join(ROOT, "data")


def filter_folder(name):
    """'filters_2x/filters_lowres' -> absolute path of the staged copy (data/: the reference's trained tables, consumed
    read-only like any caller-supplied model folder) or the reference tree."""
    for base in (DATA_DIR, "/root/reference"):
        p = os.path.join(base, name)
        if os.path.isdir(p):
            return p
    raise FileNotFoundError(name)


def have_filters():
    try:
        filter_folder("filters_2x/filters_lowres")
        return True
    except FileNotFoundError:
        return False


def color_range(bits, rng=VideoRange):
    if bits == 8:
        return (16, 235) if rng == VideoRange else (0, 255)
    if bits == 10:
        return (64, 940) if rng == VideoRange else (0, 1023)
    return (0, 65535)


def read_model(folder, bits, second=False):
    """Reads filterbin_2_<bits>[_2] + Qfactor files (format: Raisr.cpp:246-433).
    Returns (filters[216][P][121] float32, qstr[2], qcoh[2])."""
    sfx = "_2_%d" % bits + ("_2" if second else "")
    with open(os.path.join(folder, "filterbin" + sfx), "rb") as f:
        raw = f.read()
    assert raw[:4] == b"fp32", raw[:4]
    hk, pt, rows = struct.unpack("<III", raw[4:16])
    w = np.frombuffer(raw, dtype="<f4", offset=16).reshape(hk, pt, rows).copy()
    qs = np.array(open(os.path.join(folder, "Qfactor_strbin" + sfx)).read().split(), dtype=np.float64).astype(np.float32)
    qc = np.array(open(os.path.join(folder, "Qfactor_cohbin" + sfx)).read().split(), dtype=np.float64).astype(np.float32)
    return w, qs, qc


def synth_frame(w, h, bits=8, seed=1234, kind="mix"):
    """Deterministic synthetic luma plane: sinusoid + checker + uniform noise."""
    mx = (1 << bits) - 1
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    mid = 128.0 * (1 << (bits - 8))
    if kind == "noise":
        v = rs.randint(0, mx + 1, size=(h, w)).astype(np.float64)
    elif kind == "flat":
        v = np.full((h, w), mid)
    elif kind == "edges":
        v = mid + 0.4 * mx * np.sign(np.sin(0.11 * x + 0.07 * y) * np.cos(0.05 * x - 0.13 * y))
    else:
        checker = (((x // 37).astype(np.int64) + (y // 23).astype(np.int64)) & 1) * 2.0 - 1.0
        noise = rs.uniform(-16.0, 16.0, size=(h, w)) * (1 << (bits - 8))
        v = mid + 0.35 * mx * np.sin(0.05 * x) * np.cos(0.07 * y) + 0.15 * mx * checker + noise
    v = np.clip(np.rint(v), 0, mx)
    return v.astype(np.uint8 if bits == 8 else np.uint16)


def synth_chroma(w, h, bits=8, seed=99):
    rs = np.random.RandomState(seed)
    return rs.randint(0, 1 << bits, size=(h, w)).astype(np.uint8 if bits == 8 else np.uint16)


# ------------------------------------------------------------------------------------------------
# oracle (C restatement)
# ------------------------------------------------------------------------------------------------
class OraclePassParams(C.Structure):
    _fields_ = [("bits", C.c_int), ("lo", C.c_int), ("hi", C.c_int), ("nptypes", C.c_int),
                ("filters", C.POINTER(C.c_float)), ("qstr", C.c_float * 2), ("qcoh", C.c_float * 2),
                ("sqrt_mode", C.c_int), ("blending", C.c_int)]


_oracle = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
        src = [os.path.join(ORACLE_DIR, f) for f in ("raisr_oracle.c", "raisr_oracle.h", "x86_approx.c")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            build_oracle()
        L = C.CDLL(path)
        L.oracle_pass.restype = C.c_int
        L.oracle_process_y.restype = C.c_int
        L.oracle_process_y_ratio.restype = C.c_int
        for n in ("oracle_x86_rcp14", "oracle_x86_rsqrt14", "oracle_x86_rcpps", "oracle_x86_rsqrtps"):
            getattr(L, n).restype = C.c_float
            getattr(L, n).argtypes = [C.c_float]
        _oracle = L
    return _oracle


class OracleModel:
    """Holds the tables of one pass alive for the C side."""

    def __init__(self, folder, bits, second=False, rng=VideoRange, sqrt_mode=0, blending=2):
        self.w, self.qs, self.qc = read_model(folder, bits, second)
        self.w = np.ascontiguousarray(self.w, dtype=np.float32)
        lo, hi = color_range(bits, rng)
        p = OraclePassParams()
        p.bits, p.lo, p.hi, p.nptypes = bits, lo, hi, self.w.shape[1]
        p.filters = self.w.ctypes.data_as(C.POINTER(C.c_float))
        p.qstr[0], p.qstr[1] = float(self.qs[0]), float(self.qs[1])
        p.qcoh[0], p.qcoh[1] = float(self.qc[0]), float(self.qc[1])
        p.sqrt_mode, p.blending = sqrt_mode, blending
        self.p = p


def oracle_pass(S, model, want=("out",)):
    """One pass over an integer plane S (H x W).  Returns dict with out/hash/gtwg/hr as requested."""
    L = oracle_lib()
    S16 = np.ascontiguousarray(S, dtype=np.uint16)
    H, W = S16.shape
    out = np.zeros((H, W), np.uint16)
    hash_ = np.zeros((H, W), np.int32) if "hash" in want else None
    gtwg = np.zeros((H, W, 3), np.float32) if "gtwg" in want else None
    hr = np.zeros((H, W), np.float32) if "hr" in want else None
    ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rc = L.oracle_pass(ptr(S16), W, H, C.byref(model.p), ptr(hash_), ptr(gtwg), ptr(hr), ptr(out))
    assert rc == 0
    return {"out": out, "hash": hash_, "gtwg": gtwg, "hr": hr}


def oracle_resize(img, outW, outH):
    L = oracle_lib()
    a = np.ascontiguousarray(img, dtype=np.uint16)
    out = np.zeros((outH, outW), np.uint16)
    L.oracle_resize(a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], out.ctypes.data_as(C.c_void_p), outW, outH)
    return out.astype(img.dtype)


def oracle_process_y(img, outW, outH, m1, m2=None, passes=1, mode=1, want_hash=False, ratio=None):
    """ratio: the reference's gRatio (1.5 or 2.0); default = the nearest of those to outW / inW.  It matters only when the output
    size was truncated (odd input at 1.5x, evenoutput): the luma resize then reads (int)(outH / ratio) source rows."""
    L = oracle_lib()
    if ratio is None:
        ratio = round(2.0 * outW / img.shape[1]) / 2.0
    a = np.ascontiguousarray(img, dtype=np.uint16)
    inH, inW = a.shape
    out = np.zeros((outH, outW), np.uint16)
    h1 = h2 = None
    if want_hash:
        h1 = np.zeros((inH, inW) if (passes == 2 and mode == 2) else (outH, outW), np.int32)
        h2 = np.zeros((outH, outW), np.int32)
    ptr = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
    rc = L.oracle_process_y_ratio(ptr(a), inW, inH, ptr(out), outW, outH, C.c_float(ratio), passes, mode, C.byref(m1.p),
                                  C.byref(m2.p) if m2 is not None else None, ptr(h1), ptr(h2))
    assert rc == 0
    out = out.astype(img.dtype)
    return (out, h1, h2) if want_hash else out


# ------------------------------------------------------------------------------------------------
# RNLHandler_* bindings: the compiled reference (oracle/_ref) and the product expose the same C ABI
# ------------------------------------------------------------------------------------------------
def _bind_handler(L):
    L.RNLHandler_Init.restype = C.c_int32
    L.RNLHandler_Init.argtypes = [C.c_char_p, C.c_float, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint, C.c_uint]
    P = C.POINTER(VideoDataType)
    L.RNLHandler_SetRes.restype = C.c_int32
    L.RNLHandler_SetRes.argtypes = [P] * 6
    L.RNLHandler_Process.restype = C.c_int32
    L.RNLHandler_Process.argtypes = [P] * 6 + [C.c_int]
    L.RNLHandler_Deinit.restype = C.c_int32
    return L


def ref_lib_path(dbg=False):
    if dbg:
        return os.path.join(REF_DIR, "libraisr_ref_dbg.so")
    flags = open("/proc/cpuinfo").read()
    name = "libraisr_ref.so" if "avx512_fp16" in flags else "libraisr_ref_v4.so"
    return os.path.join(REF_DIR, name)


def have_ref(dbg=False):
    return os.path.exists(ref_lib_path(dbg)) and "avx512f" in open("/proc/cpuinfo").read()


_libs = {}


def handler_lib(path):
    if path not in _libs:
        _libs[path] = _bind_handler(C.CDLL(path))
    return _libs[path]


def product_lib_path():
    return os.path.join(PKG_DIR, "libraisr.so")


def vdt(arr, width=None, height=None):
    """VideoDataType over a numpy plane (2-D, possibly a column-sliced view with padded step)."""
    v = VideoDataType()
    v.pData = arr.ctypes.data
    v.width = arr.shape[1] if width is None else width
    v.height = arr.shape[0] if height is None else height
    v.step = arr.strides[0]
    v.bitShift = 0
    return v


def run_handler(L, folder, inY, ratio=2.0, bits=8, rng=VideoRange, threads=1, asm=AVX512, passes=1, mode=1,
                blending=CountOfBitsChanged, inU=None, inV=None, frames=1, out_pad=0, chroma_shift=(1, 1)):
    """Init -> SetRes -> Process x frames -> Deinit, exactly vf_raisr.c's call sequence (vf_raisr.c:226-332).
    Returns (outY, outU, outV) of the last frame."""
    dt = np.uint8 if bits == 8 else np.uint16
    H, W = inY.shape
    oH, oW = int(H * ratio), int(W * ratio)
    cw, ch = W >> chroma_shift[0], H >> chroma_shift[1]
    if inU is None:
        inU = synth_chroma(cw, ch, bits, 7)
    if inV is None:
        inV = synth_chroma(cw, ch, bits, 8)
    ocw, och = int(inU.shape[1] * ratio), int(inU.shape[0] * ratio)
    outYb = np.zeros((oH, oW + out_pad), dt)
    outY = outYb[:, :oW]
    outU = np.zeros((och, ocw), dt)
    outV = np.zeros((och, ocw), dt)
    inY = np.ascontiguousarray(inY, dtype=dt) if inY.strides[1] != inY.itemsize else inY
    vs = [vdt(inY), vdt(inU), vdt(inV), vdt(outY), vdt(outU), vdt(outV)]
    rc = L.RNLHandler_Init(folder.encode(), ratio, bits, rng, threads, asm, passes, mode)
    if rc != 0:
        return rc
    try:
        rc = L.RNLHandler_SetRes(*[C.byref(v) for v in vs])
        assert rc == 0, hex(rc & 0xffffffff)
        for _ in range(frames):
            rc = L.RNLHandler_Process(*[C.byref(v) for v in vs], blending)
            assert rc == 0, hex(rc & 0xffffffff)
    finally:
        L.RNLHandler_Deinit()
    return outY.copy(), outU, outV


def run_ref_subprocess(folder_name, inY, ratio=2.0, bits=8, rng=VideoRange, threads=1, asm=AVX512, passes=1, mode=1,
                       want_hash=False):
    """Runs the compiled reference in a FRESH process (its configuration lives in process globals that RNLInit does not
    reset, Raisr_globals.h:140-203).  Returns (outY, hash planes or None)."""
    import pickle
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "in.npy"), inY)
        code = (
            "import sys, pickle, ctypes as C, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "import raisr_testlib as T\n"
            "a = pickle.load(open(%r, 'rb'))\n"
            "img = np.load(%r)\n"
            "L = T.handler_lib(T.ref_lib_path(dbg=a['want_hash']))\n"
            "H, W = img.shape; oW, oH = int(W * a['ratio']), int(H * a['ratio'])\n"
            "planes = []\n"
            "if a['want_hash']:\n"
            "    hp = (C.c_void_p * 2).in_dll(L, 'g_raisr_dbg_hash')\n"
            "    for i in range(a['passes']):\n"
            "        lr = a['passes'] == 2 and a['mode'] == 2 and i == 0\n"
            "        planes.append(np.full((H, W) if lr else (oH, oW), -1, np.int32)); hp[i] = planes[-1].ctypes.data\n"
            "out = T.run_handler(L, T.filter_folder(a['folder']), img, a['ratio'], a['bits'], a['rng'], a['threads'], a['asm'], a['passes'], a['mode'])\n"
            "np.savez(%r, out_y=out[0], **{'hash%%d' %% i: p for i, p in enumerate(planes)})\n"
        ) % (os.path.dirname(os.path.abspath(__file__)), os.path.join(td, "args.pkl"), os.path.join(td, "in.npy"),
             os.path.join(td, "out.npz"))
        pickle.dump(dict(folder=folder_name, ratio=ratio, bits=bits, rng=rng, threads=threads, asm=asm, passes=passes,
                         mode=mode, want_hash=want_hash), open(os.path.join(td, "args.pkl"), "wb"))
        subprocess.check_call([sys.executable, "-c", code], stdout=subprocess.DEVNULL)
        z = np.load(os.path.join(td, "out.npz"))
        hashes = [z["hash%d" % i] for i in range(passes)] if want_hash else None
        return z["out_y"], hashes


def load_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    ratio, bits, passes, mode, rng, seed = z["meta"][:6]
    blending = int(z["meta"][6]) if len(z["meta"]) > 6 else CountOfBitsChanged
    return dict(in_y=z["in_y"], in_u=z["in_u"], in_v=z["in_v"], out_y=z["out_y"], out_u=z["out_u"], out_v=z["out_v"],
                hash=[z["hash%d" % i].astype(np.int32) for i in range(int(passes))], ratio=float(ratio), bits=int(bits),
                passes=int(passes), mode=int(mode), rng=int(rng), folder=str(z["folder"]), kind=str(z["kind"]), blending=blending)


def golden_names():
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz") and not f.startswith("ipp_"))


def ipp_golden_names():
    """outputs of REAL Intel IPP for the golden inputs (tools/ipp_pin/; absent until someone with IPP runs it)"""
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(f[4:-4] for f in os.listdir(d) if f.endswith(".npz") and f.startswith("ipp_"))


def have_avx512():
    return "avx512f" in open("/proc/cpuinfo").read()
