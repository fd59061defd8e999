"""gseg -- B200-native graph-based image segmentation (host-side Python mirror of include/gseg.h).

The product is `libgseg.so` (hand-written sm_100a CUDA kernels behind a C-ABI).  This module only
builds it, loads it with ctypes and wraps the handle; it contains no compute and no fallback: if
the library is missing or no CUDA device is present every compute call raises.

Import it with importlib (the directory name is not a Python identifier):
    gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("GSEG_LIB") or os.path.join(_HERE, "libgseg.so")
CLI_PATH = os.path.join(_HERE, "gseg")
HEADER = os.path.join(_ROOT, "include", "gseg.h")

FELZ, HIER, SUPERPIX = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_HOST_LOOP = 1

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    """Compile libgseg.so (and the gseg CLI) in-tree for sm_100a with nvcc.  Works without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER]
    if force or _newer(LIB_PATH, srcs):
        cmd = ["nvcc"] + NVCC_FLAGS + ["-shared", "-o", LIB_PATH, os.path.join(CSRC, "gseg_api.cu")]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    cli_src = os.path.join(CSRC, "gseg_cli.cpp")
    if os.path.exists(cli_src) and (force or _newer(CLI_PATH, [cli_src, os.path.join(CSRC, "gseg_imageio.hpp"), LIB_PATH, HEADER])):
        cmd = ["g++", "-O2", "-std=c++17", "-o", CLI_PATH, cli_src, "-I", os.path.join(_ROOT, "include"),
               "-L", _HERE, "-lgseg", "-lz", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB_PATH


class Params(C.Structure):
    _fields_ = [("sigma", C.c_float), ("k", C.c_float), ("min_size", C.c_int32), ("connectivity", C.c_int32),
                ("variant", C.c_int32), ("max_levels", C.c_int32), ("max_rounds", C.c_int32), ("flags", C.c_uint32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("round", C.c_int32), ("ms", C.c_float), ("algo_bytes", C.c_double)]


class RoundStat(C.Structure):
    _fields_ = [("n_components", C.c_int64), ("n_edges", C.c_int64), ("n_merged", C.c_int64),
                ("phase", C.c_int32), ("in_tail", C.c_int32), ("us_end", C.c_float), ("us_S", C.c_float),
                ("us_R", C.c_float), ("us_E", C.c_float), ("n_pages", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def load():
    """Load libgseg.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libgseg.so is not built (run __graft_entry__.build()); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.gseg_version.restype = i32
    L.gseg_strerror.restype = C.c_char_p
    L.gseg_strerror.argtypes = [i32]
    L.gseg_last_error.restype = C.c_char_p
    L.gseg_last_error.argtypes = [vp]
    L.gseg_create.argtypes = [C.POINTER(vp), i32, i32, i32]
    L.gseg_create_ex.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    L.gseg_destroy.argtypes = [vp]
    L.gseg_destroy.restype = None
    L.gseg_set_stream.argtypes = [vp, vp]
    L.gseg_set_tail.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.gseg_set_blocks_per_sm.argtypes = [vp, i32]
    L.gseg_segment.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(Params)]
    L.gseg_segment_async.argtypes = [vp, vp, i32, i32, i32, i32, C.POINTER(Params)]
    L.gseg_wait.argtypes = [vp]
    L.gseg_jpeg_info.argtypes = [vp, C.c_size_t, C.POINTER(i32), C.POINTER(i32)]
    L.gseg_segment_jpeg.argtypes = [vp, vp, C.c_size_t, C.POINTER(Params), C.POINTER(i32), C.POINTER(i32)]
    L.gseg_segment_jpeg_async.argtypes = [vp, vp, C.c_size_t, C.POINTER(Params), C.POINTER(i32), C.POINTER(i32)]
    L.gseg_input_rgb.argtypes = [vp, vp, i32]
    L.gseg_num_levels.argtypes = [vp]
    L.gseg_num_components.argtypes = [vp, i32]
    L.gseg_labels.argtypes = [vp, i32, vp, i32]
    L.gseg_labels_async.argtypes = [vp, i32, vp, i32]
    L.gseg_sync.argtypes = [vp]
    L.gseg_labels_all.argtypes = [vp, vp, i32, i32]
    L.gseg_colorize.argtypes = [vp, i32, u64, vp, i32]
    L.gseg_weights.argtypes = [vp, vp, i32]
    L.gseg_blurred.argtypes = [vp, vp, i32]
    L.gseg_stats.argtypes = [vp, C.POINTER(RoundStat), i32]
    i64 = C.c_int64
    L.gseg_export_graph.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64), vp, vp, vp, vp, vp, i64, i64]
    L.gseg_blurred_rows.argtypes = [vp, i32, i32, vp, i32]
    L.gseg_segment_graph.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp, C.POINTER(Params), vp]
    L.gseg_synth.argtypes = [vp, vp, i32, i32, u64, i32]
    L.gseg_set_profiling.argtypes = [vp, i32]
    L.gseg_profile_read.argtypes = [vp, C.POINTER(KernelTime), i32]
    L.gseg_launch_count.argtypes = [vp]
    L.gseg_launch_count.restype = C.c_longlong
    L.gseg_sort_pairs_u64.argtypes = [vp, vp, vp, C.c_int64, i32, i32]
    _lib = L
    return L


class GsegError(RuntimeError):
    pass


def jpeg_info(data):
    """(w, h) of a JPEG (gseg_jpeg_info; needs a CUDA device and libnvjpeg like the decode itself)."""
    L = load()
    w, h = C.c_int32(0), C.c_int32(0)
    rc = L.gseg_jpeg_info(data, len(data), C.byref(w), C.byref(h))
    if rc:
        raise GsegError("gseg_jpeg_info: %s" % L.gseg_strerror(rc).decode())
    return int(w.value), int(h.value)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if _is_torch(x):
        return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    return x.ctypes.data, MEM_HOST


class Segmenter:
    """One gseg context (one GPU, one stream).  Mirrors the reference executables' parameter list:
    image, sigma, k, min_size, connectivity, variant / hierarchy level (BASELINE.json north_star)."""

    def __init__(self, max_w, max_h, device=0, max_connectivity=8):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.gseg_create_ex(C.byref(self.h), device, max_w, max_h, max_connectivity)
        if rc != 0:
            raise GsegError("gseg_create: %s" % self.L.gseg_strerror(rc).decode())
        self.device = device
        self.w = self.hh = 0
        self.D = 2

    def close(self):
        if self.h:
            self.L.gseg_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc < 0:
            raise GsegError("%s: %s (%s)" % (what, self.L.gseg_strerror(rc).decode(),
                                             self.L.gseg_last_error(self.h).decode()))
        return rc

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.gseg_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "gseg_set_stream")

    def set_tail(self, max_edges, max_components):
        self._ck(self.L.gseg_set_tail(self.h, max_edges, max_components), "gseg_set_tail")

    def set_blocks_per_sm(self, blocks):
        self._ck(self.L.gseg_set_blocks_per_sm(self.h, blocks), "gseg_set_blocks_per_sm")

    def params(self, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=FELZ, max_levels=0, max_rounds=0,
               flags=0):
        return Params(sigma, k, min_size, connectivity, variant, max_levels, max_rounds, flags)

    def segment(self, img, params=None, wait=True, **kw):
        """img: (h, w, 3) uint8 numpy array / torch CPU tensor (host) or torch CUDA tensor (device)."""
        p = params if params is not None else self.params(**kw)
        hh, w = int(img.shape[0]), int(img.shape[1])
        if _is_torch(img):
            assert img.dtype.__str__() == "torch.uint8" and img.is_contiguous()
            stride = img.stride(0)
        else:
            assert img.dtype == np.uint8 and img.flags["C_CONTIGUOUS"]
            stride = img.strides[0]
        ptr, kind = _ptr(img)
        self.w, self.hh, self.D = w, hh, (4 if p.connectivity == 8 else 2)
        self._keep = img
        fn = self.L.gseg_segment if wait else self.L.gseg_segment_async
        self._ck(fn(self.h, C.c_void_p(ptr), w, hh, stride, kind, C.byref(p)), "gseg_segment")
        return self

    def wait(self):
        self._ck(self.L.gseg_wait(self.h), "gseg_wait")
        return self

    def segment_jpeg(self, data, params=None, wait=True, **kw):
        """data: the bytes of a JPEG file.  nvJPEG decodes them on the GPU into the context's staged RGB
        buffer and the usual path runs on it (gseg_segment_jpeg); returns (w, h)."""
        p = params if params is not None else self.params(**kw)
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        w, h = C.c_int32(0), C.c_int32(0)
        fn = self.L.gseg_segment_jpeg if wait else self.L.gseg_segment_jpeg_async
        self._keep = buf
        self._ck(fn(self.h, C.cast(buf, C.c_void_p), len(data), C.byref(p), C.byref(w), C.byref(h)), "gseg_segment_jpeg")
        self.w, self.hh, self.D = int(w.value), int(h.value), (4 if p.connectivity == 8 else 2)
        return self.w, self.hh

    def input_rgb(self):
        """(h, w, 3) uint8: the image the last run read, when the context staged it (host input or JPEG)."""
        out = np.empty((self.hh, self.w, 3), np.uint8)
        self._ck(self.L.gseg_input_rgb(self.h, out.ctypes.data, 0), "gseg_input_rgb")
        return out

    def num_levels(self):
        return self._ck(self.L.gseg_num_levels(self.h), "gseg_num_levels")

    def num_components(self, level=-1):
        return self._ck(self.L.gseg_num_components(self.h, level), "gseg_num_components")

    def labels(self, level=-1, out=None, wait=True):
        """Label image of a level.  wait=False only enqueues it (out must be a CUDA tensor or pinned
        host memory); sync() or any later synchronous call completes it."""
        if out is None:
            out = np.empty((self.hh, self.w), np.int32)
        ptr, kind = _ptr(out)
        fn = self.L.gseg_labels if wait else self.L.gseg_labels_async
        self._ck(fn(self.h, level, C.c_void_p(ptr), kind), "gseg_labels")
        return out

    def sync(self):
        self._ck(self.L.gseg_sync(self.h), "gseg_sync")

    def labels_all(self, max_levels=64, out=None):
        n = min(max_levels, self.num_levels())
        if out is None:
            out = np.empty((max(n, 1), self.hh, self.w), np.int32)
        ptr, kind = _ptr(out)
        got = self._ck(self.L.gseg_labels_all(self.h, C.c_void_p(ptr), int(out.shape[0]), kind), "gseg_labels_all")
        return out[:got]

    def colorize(self, level=-1, seed=1, out=None):
        if out is None:
            out = np.empty((self.hh, self.w, 3), np.uint8)
        ptr, kind = _ptr(out)
        self._ck(self.L.gseg_colorize(self.h, level, seed, C.c_void_p(ptr), kind), "gseg_colorize")
        return out

    def weights(self):
        out = np.empty(self.hh * self.w * self.D, np.float32)
        self._ck(self.L.gseg_weights(self.h, C.c_void_p(out.ctypes.data), MEM_HOST), "gseg_weights")
        return out

    def blurred(self):
        out = np.empty((3, self.hh, self.w), np.float32)
        self._ck(self.L.gseg_blurred(self.h, C.c_void_p(out.ctypes.data), MEM_HOST), "gseg_blurred")
        return out

    def export_graph(self, dedup=True):
        """Final component graph of the last FELZ/HIER run: dict(size, Int, ea, eb, w); ids = labels(-1).
        dedup: parallel edges reduced to their minimum (weight, position) with the onesweep radix sort."""
        nv, ne = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.gseg_export_graph(self.h, int(dedup), C.byref(nv), C.byref(ne), None, None, None, None, None, 0, 0),
                 "gseg_export_graph")
        size = np.empty(nv.value, np.uint32)
        Int = np.empty(nv.value, np.float32)
        ea = np.empty(max(ne.value, 1), np.uint32)
        eb = np.empty(max(ne.value, 1), np.uint32)
        w = np.empty(max(ne.value, 1), np.float32)
        self._ck(self.L.gseg_export_graph(self.h, int(dedup), C.byref(nv), C.byref(ne), size.ctypes.data, Int.ctypes.data,
                                          ea.ctypes.data, eb.ctypes.data, w.ctypes.data, len(size), len(ea)),
                 "gseg_export_graph")
        return dict(size=size, Int=Int, ea=ea[:ne.value], eb=eb[:ne.value], w=w[:ne.value])

    def blurred_rows(self, y0, nrows=1):
        out = np.empty((3, nrows, self.w), np.float32)
        self._ck(self.L.gseg_blurred_rows(self.h, y0, nrows, C.c_void_p(out.ctypes.data), MEM_HOST), "gseg_blurred_rows")
        return out

    def segment_graph(self, size, Int, ea, eb, w, params=None, **kw):
        """Boruvka rounds on an explicit graph; returns (dense final label per input component, #components)."""
        p = params if params is not None else self.params(**kw)
        size = np.ascontiguousarray(size, np.uint32)
        Int = np.ascontiguousarray(Int, np.float32)
        ea = np.ascontiguousarray(ea, np.uint32)
        eb = np.ascontiguousarray(eb, np.uint32)
        w = np.ascontiguousarray(w, np.float32)
        out = np.empty(len(size), np.int32)
        n = self._ck(self.L.gseg_segment_graph(self.h, len(size), size.ctypes.data, Int.ctypes.data, len(ea),
                                               ea.ctypes.data, eb.ctypes.data, w.ctypes.data, C.byref(p),
                                               out.ctypes.data), "gseg_segment_graph")
        return out, n

    def stats(self):
        arr = (RoundStat * 64)()
        n = self._ck(self.L.gseg_stats(self.h, arr, 64), "gseg_stats")
        return [(arr[i].n_components, arr[i].n_edges, arr[i].n_merged, arr[i].phase) for i in range(n)]

    def timeline(self):
        """Device timeline of the last run: [(round, in_tail, us_end, us_S, us_R, us_E)]."""
        arr = (RoundStat * 64)()
        n = self._ck(self.L.gseg_stats(self.h, arr, 64), "gseg_stats")
        return [(i, arr[i].in_tail, arr[i].us_end, arr[i].us_S, arr[i].us_R, arr[i].us_E, arr[i].n_pages) for i in range(n)]

    def set_profiling(self, on):
        self._ck(self.L.gseg_set_profiling(self.h, int(on)), "gseg_set_profiling")

    def profile(self):
        """[(kernel name, round, ms, algorithmic bytes)] of the last host-driven run with profiling on."""
        arr = (KernelTime * 1024)()
        n = self._ck(self.L.gseg_profile_read(self.h, arr, 1024), "gseg_profile_read")
        return [(arr[i].name.decode(), arr[i].round, arr[i].ms, arr[i].algo_bytes) for i in range(n)]

    def launch_count(self):
        return int(self.L.gseg_launch_count(self.h))

    def synth(self, w, h, seed, out=None):
        if out is None:
            out = np.empty((h, w, 3), np.uint8)
        ptr, kind = _ptr(out)
        self._ck(self.L.gseg_synth(self.h, C.c_void_p(ptr), w, h, seed, kind), "gseg_synth")
        return out

    def sort_pairs(self, keys_dev_ptr, vals_dev_ptr, n, begin_bit=0, end_bit=64):
        self._ck(self.L.gseg_sort_pairs_u64(self.h, C.c_void_p(keys_dev_ptr), C.c_void_p(vals_dev_ptr), n,
                                            begin_bit, end_bit), "gseg_sort_pairs_u64")
