"""ctypes front-end of the CPU oracle (oracle/gseg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Parity status: *** parity unpinned *** -- the mounted reference contains no code, tests or golden
vectors; see the header of gseg_oracle.c for what is restated from where.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

FELZ, HIER, SUPERPIX, KRUSKAL = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "gseg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, f32p, i32p, i64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int64))
        L.orc_synth.argtypes = [u8p, C.c_int, C.c_int, C.c_uint64]
        L.orc_synth.restype = None
        L.orc_gauss_mask.argtypes = [C.c_float, f32p]
        L.orc_gauss_mask.restype = C.c_int
        L.orc_blur.argtypes = [u8p, C.c_int, C.c_int, C.c_float, f32p]
        L.orc_blur.restype = None
        L.orc_sobel.argtypes = [f32p, C.c_int, C.c_int, f32p]
        L.orc_sobel.restype = None
        L.orc_edges.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_edges.restype = C.c_int64
        L.orc_strength.argtypes = [f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_strength.restype = None
        L.orc_felz_kruskal.argtypes = [C.c_int, C.c_int, C.c_int, f32p, C.c_float, C.c_int, i32p]
        L.orc_felz_kruskal.restype = C.c_int
        L.orc_boruvka.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_float, C.c_int, C.c_int, f32p,
                                  i32p, i32p, C.c_int, i32p, i64p, C.c_int, C.POINTER(C.c_int)]
        L.orc_boruvka.restype = C.c_int
        L.orc_set_int_out.argtypes = [f32p]
        L.orc_set_int_out.restype = None
        L.orc_boruvka_graph.argtypes = [C.c_int, i32p, f32p, C.c_int64, i32p, i32p, f32p, C.c_int, C.c_float, C.c_int,
                                        C.c_int, i32p, i64p, C.c_int]
        L.orc_boruvka_graph.restype = C.c_int
        L.orc_canon.argtypes = [i32p, C.c_int64]
        L.orc_canon.restype = C.c_int
        L.orc_segment.argtypes = [u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                  i32p]
        L.orc_segment.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def synth(w, h, seed):
    img = np.empty((h, w, 3), np.uint8)
    lib().orc_synth(_p(img, C.c_uint8), w, h, seed)
    return img


def synth_numpy(w, h, seed):
    """Independent numpy restatement of the generator (checks orc_synth and the CUDA generator)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)

    def sm64(x):
        with np.errstate(over="ignore"):
            z = (x + np.uint64(0x9E3779B97F4A7C15)) & M
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
            return z ^ (z >> np.uint64(31))

    def hash2(a, b):
        with np.errstate(over="ignore"):
            return sm64(sm64(np.uint64(seed) ^ (a * np.uint64(0xD6E8FEB86659FD93))) + np.uint64(b))

    ys, xs = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    cx, cy = xs >> 6, ys >> 6
    bestd = np.full((h, w), np.iinfo(np.int64).max, np.int64)
    besth = np.zeros((h, w), np.uint64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            ccx, ccy = cx + dx, cy + dy
            cell = ((ccy + 1).astype(np.uint64) << np.uint64(20)) | (ccx + 1).astype(np.uint64)
            hs = hash2(cell, 1)
            sx = ccx * 64 + (hs & np.uint64(63)).astype(np.int64)
            sy = ccy * 64 + ((hs >> np.uint64(6)) & np.uint64(63)).astype(np.int64)
            d = (xs - sx) ** 2 + (ys - sy) ** 2
            m = d < bestd
            bestd[m] = d[m]
            besth[m] = hs[m]
    hn = hash2((ys * w + xs).astype(np.uint64), 2)
    out = np.empty((h, w, 3), np.uint8)
    for c in range(3):
        base = ((besth >> np.uint64(16 + 8 * c)) & np.uint64(255)).astype(np.int64)
        n = (((hn >> np.uint64(16 * c)) & np.uint64(0xFFFF)) % np.uint64(17)).astype(np.int64) - 8
        out[..., c] = np.clip(base + n, 0, 255).astype(np.uint8)
    return out


def gauss_mask(sigma):
    m = np.zeros(64, np.float32)
    n = lib().orc_gauss_mask(sigma, _p(m, C.c_float))
    return m[:n].copy()


def blur(img, sigma):
    h, w, _ = img.shape
    img = np.ascontiguousarray(img)
    out = np.empty((3, h, w), np.float32)
    lib().orc_blur(_p(img, C.c_uint8), w, h, sigma, _p(out, C.c_float))
    return out


def sobel(planes):
    _, h, w = planes.shape
    G = np.empty((h, w), np.float32)
    lib().orc_sobel(_p(planes, C.c_float), w, h, _p(G, C.c_float))
    return G


def edges(planes, conn):
    """Edge weights in edge-index order idx = d*V + p; +inf where the edge does not exist."""
    _, h, w = planes.shape
    D = 4 if conn == 8 else 2
    wts = np.empty(h * w * D, np.float32)
    n = lib().orc_edges(_p(planes, C.c_float), w, h, conn, _p(wts, C.c_float))
    return wts, int(n)


def strength(G, conn):
    h, w = G.shape
    D = 4 if conn == 8 else 2
    s = np.empty(h * w * D, np.float32)
    lib().orc_strength(_p(G, C.c_float), w, h, conn, _p(s, C.c_float))
    return s


def canon(labels):
    lab = np.ascontiguousarray(labels, np.int32).copy().reshape(-1)
    n = lib().orc_canon(_p(lab, C.c_int32), lab.size)
    return lab.reshape(np.shape(labels)), n


def felz_kruskal(wts, w, h, conn, k, min_size):
    lab = np.empty(h * w, np.int32)
    n = lib().orc_felz_kruskal(w, h, conn, _p(wts, C.c_float), k, min_size, _p(lab, C.c_int32))
    return lab.reshape(h, w), n


def boruvka(wts, w, h, conn, variant, k=0.0, min_size=0, max_rounds=64, planes=None, max_levels=0, want_int=False):
    """Returns dict(labels, levels[list of label images], ncomp[list], stats[rounds x 4], n).
    max_levels = 0: run until one component and keep no per-level images.
    want_int: also "int" = final Int(C) indexed by representative pixel (the values labels take)."""
    V = h * w
    int_out = np.zeros(V, np.float32) if want_int else None
    lib().orc_set_int_out(_p(int_out, C.c_float) if want_int else None)
    lab = np.empty(V, np.int32)
    lev = np.empty((max(max_levels, 1), V), np.int32)
    nco = np.zeros(max(max_levels, 1), np.int32)
    stats = np.zeros((4 * max_rounds + 8, 4), np.int64)
    fin = C.c_int(0)
    pl = _p(np.ascontiguousarray(planes, np.float32), C.c_float) if planes is not None else None
    nl = lib().orc_boruvka(w, h, conn, variant, _p(wts, C.c_float), k, min_size, max_rounds, pl,
                           _p(lab, C.c_int32), _p(lev, C.c_int32) if max_levels > 0 else None,
                           max_levels if max_levels > 0 else 1 << 30,
                           _p(nco, C.c_int32) if max_levels > 0 else None, _p(stats, C.c_int64), stats.shape[0],
                           C.byref(fin))
    lib().orc_set_int_out(None)
    nst = int(np.count_nonzero(stats[:, 0]))
    nkeep = min(nl, max_levels)
    return dict(int=int_out, labels=lab.reshape(h, w), levels=[lev[i].reshape(h, w) for i in range(nkeep)],
                ncomp=[int(x) for x in nco[:nkeep]], stats=stats[:nst].copy(), n=fin.value, nlevels=nl)


def boruvka_graph(size, Int, ea, eb, w, variant, k=0.0, min_size=0, max_rounds=64):
    """Rounds on an explicit graph (second phase of the tiled schedule).  Returns (labels per input
    component = representative ids, number of final components, stats[rounds x 4])."""
    size = np.ascontiguousarray(size, np.int32)
    Int = np.ascontiguousarray(Int, np.float32)
    ea = np.ascontiguousarray(ea, np.int32)
    eb = np.ascontiguousarray(eb, np.int32)
    w = np.ascontiguousarray(w, np.float32)
    nv, ne = len(size), len(ea)
    lab = np.empty(max(nv, 1), np.int32)
    stats = np.zeros((4 * max_rounds + 8, 4), np.int64)
    n = lib().orc_boruvka_graph(nv, _p(size, C.c_int32), _p(Int, C.c_float), ne, _p(ea, C.c_int32), _p(eb, C.c_int32),
                                _p(w, C.c_float), variant, k, min_size, max_rounds, _p(lab, C.c_int32),
                                _p(stats, C.c_int64), stats.shape[0])
    nst = int(np.count_nonzero(stats[:, 0]))
    return lab[:nv], n, stats[:nst].copy()


def segment(img, sigma, k, min_size, conn, variant, max_rounds=64):
    h, w, _ = img.shape
    img = np.ascontiguousarray(img)
    lab = np.empty(h * w, np.int32)
    n = lib().orc_segment(_p(img, C.c_uint8), w, h, sigma, k, min_size, conn, variant, max_rounds, _p(lab, C.c_int32))
    return lab.reshape(h, w), n


def pipeline(img, sigma, k, min_size, conn, variant, max_rounds=64, max_levels=0):
    """Stage-by-stage oracle run keeping every intermediate (used by the parity tests)."""
    h, w, _ = img.shape
    planes = blur(img, sigma)
    if variant == SUPERPIX:
        G = sobel(planes)
        wts = strength(G, conn)
    else:
        wts, _ = edges(planes, conn)
    if variant == KRUSKAL:
        lab, n = felz_kruskal(wts, w, h, conn, k, min_size)
        return dict(planes=planes, wts=wts, labels=lab, n=n)
    r = boruvka(wts, w, h, conn, variant, k, min_size, max_rounds, planes, max_levels)
    r.update(planes=planes, wts=wts)
    return r
