"""ctypes binding of libraisr.so's thin C ABI (include/raisr_cuda.h) for the tests and the bench.
The product itself is the shared library; this file only marshals pointers."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libraisr.so")

NUMERICS_IEEE, NUMERICS_X86, NUMERICS_AUTO, NUMERICS_FP16_FILTER, NUMERICS_FAST_HASH = 0, 1, 2, 3, 4


class Config(C.Structure):
    _fields_ = [("model_path", C.c_char_p), ("ratio", C.c_float), ("bit_depth", C.c_uint), ("range_type", C.c_int),
                ("passes", C.c_uint), ("two_pass_mode", C.c_uint), ("device", C.c_int), ("numerics", C.c_int),
                ("keep_hash", C.c_int)]


_lib = None


def lib():
    """Loads libraisr.so; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libraisr.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        vp, sz = C.c_void_p, C.c_size_t
        L.raisr_cuda_create.restype = C.c_int32
        L.raisr_cuda_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.raisr_cuda_set_res.restype = C.c_int32
        L.raisr_cuda_set_res.argtypes = [vp] + [C.c_uint] * 8
        L.raisr_cuda_process_host.restype = C.c_int32
        L.raisr_cuda_process_host.argtypes = [vp] + [vp, sz] * 6 + [C.c_int]
        L.raisr_cuda_process_device.restype = C.c_int32
        L.raisr_cuda_process_device.argtypes = [vp] + [vp, sz] * 6 + [C.c_int, vp]
        L.raisr_cuda_process_device_semiplanar.restype = C.c_int32
        L.raisr_cuda_process_device_semiplanar.argtypes = [vp] + [vp, sz] * 4 + [C.c_int, C.c_int, vp]
        L.raisr_cuda_process_device_rows.restype = C.c_int32
        L.raisr_cuda_process_device_rows.argtypes = [vp, vp, sz, vp, sz, C.c_int, C.c_uint, C.c_uint, vp]
        L.raisr_cuda_read_hash.restype = C.c_int32
        L.raisr_cuda_read_hash.argtypes = [vp, C.c_int, vp, sz]
        L.raisr_cuda_launch_count.restype = C.c_ulonglong
        L.raisr_cuda_launch_count.argtypes = [vp]
        L.raisr_cuda_numerics.restype = C.c_int
        L.raisr_cuda_numerics.argtypes = [vp]
        L.raisr_cuda_destroy.restype = None
        L.raisr_cuda_destroy.argtypes = [vp]
        L.raisr_cuda_version.restype = C.c_char_p
        for n in ("raisr_cuda_ipc_export", "raisr_cuda_ipc_open", "raisr_cuda_ipc_close"):
            getattr(L, n).restype = C.c_int32
        L.raisr_cuda_ipc_export.argtypes = [vp, C.c_char_p, C.POINTER(sz)]
        L.raisr_cuda_ipc_open.argtypes = [C.c_char_p, sz, C.POINTER(vp)]
        L.raisr_cuda_ipc_close.argtypes = [vp]
        _lib = L
    return _lib


def ipc_export(device_ptr):
    """64-byte CUDA IPC handle of a device allocation (peer-store row bands: the gathering rank exports its frame buffer)"""
    buf = C.create_string_buffer(64)
    off = C.c_size_t(0)
    rc = lib().raisr_cuda_ipc_export(device_ptr, buf, C.byref(off))
    if rc != 0:
        raise RuntimeError("raisr_cuda_ipc_export failed: 0x%08x" % (rc & 0xffffffff))
    return buf.raw, int(off.value)


def ipc_open(handle_and_offset):
    """device pointer (int) in THIS process for another process's exported allocation (+ the exporter's interior offset)"""
    handle, off = handle_and_offset
    p = C.c_void_p()
    rc = lib().raisr_cuda_ipc_open(handle, off, C.byref(p))
    if rc != 0:
        raise RuntimeError("raisr_cuda_ipc_open failed: 0x%08x" % (rc & 0xffffffff))
    return p.value


def ipc_close(ptr, handle_and_offset):
    lib().raisr_cuda_ipc_close(ptr - handle_and_offset[1])


class Engine:
    """Thin RAII wrapper over raisr_cuda_engine*."""

    def __init__(self, model_path, ratio=2.0, bits=8, range_type=1, passes=1, mode=1, device=-1,
                 numerics=NUMERICS_AUTO, keep_hash=False):
        self.L = lib()
        self.cfg = Config(model_path.encode(), ratio, bits, range_type, passes, mode, device, numerics, int(keep_hash))
        self.h = C.c_void_p()
        rc = self.L.raisr_cuda_create(C.byref(self.cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise RuntimeError("raisr_cuda_create failed: 0x%08x" % (rc & 0xffffffff))
        self.bits, self.ratio, self.passes, self.mode = bits, ratio, passes, mode

    def set_res(self, in_w, in_h, out_w, out_h, in_cw=0, in_ch=0, out_cw=0, out_ch=0):
        rc = self.L.raisr_cuda_set_res(self.h, in_w, in_h, out_w, out_h, in_cw, in_ch, out_cw, out_ch)
        if rc != 0:
            raise RuntimeError("raisr_cuda_set_res failed: 0x%08x" % (rc & 0xffffffff))
        self.dims = (in_w, in_h, out_w, out_h)

    def process_host(self, in_y, out_y, in_u=None, in_v=None, out_u=None, out_v=None, blending=2):
        """numpy planes (2-D, row stride = .strides[0])"""
        def ps(a):
            return (a.ctypes.data, a.strides[0]) if a is not None else (None, 0)
        args = [*ps(in_y), *ps(in_u), *ps(in_v), *ps(out_y), *ps(out_u), *ps(out_v)]
        return self.L.raisr_cuda_process_host(self.h, *args, blending)

    def process_device(self, in_y, in_y_step, out_y, out_y_step, in_u=None, in_u_step=0, in_v=None, in_v_step=0,
                       out_u=None, out_u_step=0, out_v=None, out_v_step=0, blending=2, stream=None):
        """raw device pointers (ints)"""
        return self.L.raisr_cuda_process_device(self.h, in_y, in_y_step, in_u, in_u_step, in_v, in_v_step, out_y,
                                                out_y_step, out_u, out_u_step, out_v, out_v_step, blending, stream)

    def process_device_semiplanar(self, in_y, in_y_step, in_uv, in_uv_step, out_y, out_y_step, out_uv, out_uv_step, shift=0, blending=2, stream=None):
        """NV12 (shift 0) / P010 (shift 6) device frames: raw device pointers (ints)"""
        return self.L.raisr_cuda_process_device_semiplanar(self.h, in_y, in_y_step, in_uv, in_uv_step, out_y, out_y_step, out_uv, out_uv_step,
                                                           shift, blending, stream)

    def process_device_rows(self, in_y, in_y_step, out_y, out_y_step, row0, row1, blending=2, stream=None):
        return self.L.raisr_cuda_process_device_rows(self.h, in_y, in_y_step, out_y, out_y_step, blending, row0, row1, stream)

    def read_hash(self, pass_idx, w, h):
        import numpy as np
        out = np.empty((h, w), np.int32)
        rc = self.L.raisr_cuda_read_hash(self.h, pass_idx, out.ctypes.data, out.size)
        if rc != 0:
            raise RuntimeError("raisr_cuda_read_hash failed: 0x%08x" % (rc & 0xffffffff))
        return out

    def numerics(self):
        return int(self.L.raisr_cuda_numerics(self.h))

    def launch_count(self):
        return int(self.L.raisr_cuda_launch_count(self.h))

    def close(self):
        if self.h:
            self.L.raisr_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
