"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libtlc_oracle.so (the CPU restatement).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtlc_oracle.so")

MODE_EDGE, MODE_NODE, MODE_EDGE_FORCED, MODE_EDGE_UNION, MODE_EDGE_REMOVEINTER = 0, 1, 2, 3, 4
DESC = {"min": 0, "max": 1, "sum": 2}
F_NORM, F_EXTENDED, F_KEEP_ZERO, F_NORM_EPS, F_SUM_PLAIN, F_ASIS_FV = 1, 2, 4, 8, 16, 32
F_FILT_DEGREE, F_FILT_CENTRALITY, F_FILT_CLUSTERING = 512, 1024, 2048


def hks_signature(n, elo, ehi, time=0.1):
    """Knowledge_Distillation/data_utils_NC.py:87-93,115-117 on a vicinity given by its canonical edge list: the reference's
    own lines (nx.adjacency_matrix -> csgraph.laplacian(normed=True) -> eigh -> sum of squares), then / (max + 1e-10)"""
    import scipy.sparse as sp
    from scipy.linalg import eigh
    from scipy.sparse import csgraph
    a = np.asarray(elo, dtype=np.int64)
    b = np.asarray(ehi, dtype=np.int64)
    A = sp.coo_matrix((np.ones(2 * len(a)), (np.concatenate([a, b]), np.concatenate([b, a]))), shape=(n, n)).tocsr()
    Lm = csgraph.laplacian(A, normed=True)
    egvals, egvectors = eigh(Lm.toarray())
    f = np.square(egvectors).dot(np.diag(np.exp(-time * egvals))).sum(axis=1)
    return f / (max(f) + 1e-10)
K_UP, K_ESS, K_DOWN, K_ESS_REV, K_ONE = 0, 1, 2, 3, 4
ST_NAMES = ["OK", "TRIVIAL", "EMPTY", "DISCONNECTED", "DEGENERATE", "UNKNOWN_NODE", "BAD_DESCRIPTOR", "NO_TREE_EDGES"]


def build(force=False):
    src = os.path.join(_HERE, "tlc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Graph(C.Structure):
    _fields_ = [("N", C.c_int32), ("rowptr", C.c_void_p), ("col", C.c_void_p), ("kappa", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("hop", C.c_int32), ("mode", C.c_int32), ("descriptor", C.c_int32), ("resolution", C.c_int32),
                ("flags", C.c_uint32), ("img_mask", C.c_uint32)]


_DET_I = ["vert", "elo", "ehi"]
_DET_FIELDS = [("n", C.c_int32), ("m", C.c_int32), ("npairs", C.c_int32), ("npos", C.c_int32), ("nneg", C.c_int32),
               ("lu", C.c_int32), ("lv", C.c_int32),
               ("vert", C.c_void_p), ("elo", C.c_void_p), ("ehi", C.c_void_p), ("ew", C.c_void_p),
               ("d1", C.c_void_p), ("d2", C.c_void_p), ("fval", C.c_void_p),
               ("ord_asc", C.c_void_p), ("ord_desc", C.c_void_p),
               ("pkind", C.c_void_p), ("pbv", C.c_void_p), ("pdv", C.c_void_p),
               ("pbirth", C.c_void_p), ("pdeath", C.c_void_p), ("pos", C.c_void_p), ("neg", C.c_void_p)]


class _Detail(C.Structure):
    _fields_ = _DET_FIELDS


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.tlo_run_one.restype = C.c_int
        _lib.tlo_run_one.argtypes = [C.POINTER(_Graph), C.c_int32, C.c_int32, C.POINTER(_Params), C.c_void_p,
                                     C.POINTER(_Detail)]
        _lib.tlo_run_one_fval.restype = C.c_int
        _lib.tlo_run_one_fval.argtypes = [C.POINTER(_Graph), C.c_int32, C.c_int32, C.POINTER(_Params), C.c_void_p, C.c_void_p,
                                          C.POINTER(_Detail)]
        _lib.tlo_run_batch.restype = C.c_int64
        _lib.tlo_run_batch.argtypes = [C.POINTER(_Graph), C.c_void_p, C.c_int64, C.POINTER(_Params), C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.tlo_pimg_transform.restype = None
        _lib.tlo_pimg_transform.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        _lib.tlo_max_threads.restype = C.c_int
    return _lib


def default_img_mask(extended, kd=False):
    if kd:
        return (1 << K_UP) | (1 << K_ONE)  # data_utils_NC.py:172-180: Ord0 u Ext1
    return (1 << K_UP) | (1 << K_ESS) | (1 << K_DOWN) | (1 << K_ESS_REV) | ((1 << K_ONE) if extended else 0)


class OracleGraph:
    """CSR in the reference's node numbering (graph2pi.__init__, riccidist2dgm.py:216-226)."""

    def __init__(self, rowptr, col, kappa):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        self.N = self.rowptr.size - 1
        self.M = self.col.size // 2
        self._g = _Graph(self.N, self.rowptr.ctypes.data, self.col.ctypes.data, self.kappa.ctypes.data)

    def _params(self, hop, mode, descriptor, resolution, flags, img_mask):
        d = DESC[descriptor] if isinstance(descriptor, str) else int(descriptor)
        return _Params(hop, mode, d, resolution, flags, img_mask)

    def run_batch(self, targets, hop=2, mode=MODE_EDGE, descriptor="sum", resolution=5,
                  flags=F_NORM, img_mask=None, nthreads=1):
        t = np.ascontiguousarray(targets, dtype=np.int32).reshape(-1, 2)
        E = t.shape[0]
        if img_mask is None:
            img_mask = default_img_mask(bool(flags & F_EXTENDED), bool(flags & F_KEEP_ZERO))
        p = self._params(hop, mode, descriptor, resolution, flags, img_mask)
        pi = np.zeros((E, resolution * resolution), dtype=np.float64)
        status = np.zeros(E, dtype=np.uint8)
        n = np.zeros(E, dtype=np.int32)
        m = np.zeros(E, dtype=np.int32)
        npairs = np.zeros(E, dtype=np.int32)
        cnt = lib().tlo_run_batch(C.byref(self._g), t.ctypes.data, E, C.byref(p), nthreads, pi.ctypes.data,
                                  status.ctypes.data, n.ctypes.data, m.ctypes.data, npairs.ctypes.data)
        return dict(pi=pi, status=status, n=n, m=m, npairs=npairs, cnt_compute=int(cnt))

    def run_one(self, u, v, hop=2, mode=MODE_EDGE, descriptor="sum", resolution=5, flags=F_NORM, img_mask=None, fval=None):
        """every intermediate of one target (stage-level parity checks).  fval: filtration values to use instead of the
        computed ones (canonical vertex order), e.g. a heat kernel signature evaluated with the reference's numpy lines."""
        if img_mask is None:
            img_mask = default_img_mask(bool(flags & F_EXTENDED), bool(flags & F_KEEP_ZERO))
        p = self._params(hop, mode, descriptor, resolution, flags, img_mask)
        ncap, mcap = self.N + 1, self.M + 1
        pc = ncap + mcap + 4
        bufs = dict(vert=np.zeros(ncap, np.int32), elo=np.zeros(mcap, np.int32), ehi=np.zeros(mcap, np.int32),
                    ew=np.zeros(mcap), d1=np.zeros(ncap), d2=np.zeros(ncap), fval=np.zeros(ncap),
                    ord_asc=np.zeros(mcap, np.int32), ord_desc=np.zeros(mcap, np.int32),
                    pkind=np.zeros(pc, np.int32), pbv=np.zeros(pc, np.int32), pdv=np.zeros(pc, np.int32),
                    pbirth=np.zeros(pc), pdeath=np.zeros(pc), pos=np.zeros(mcap, np.int32), neg=np.zeros(mcap, np.int32))
        det = _Detail()
        for k, a in bufs.items():
            setattr(det, k, a.ctypes.data)
        img = np.zeros(resolution * resolution)
        if fval is not None:
            fv = np.ascontiguousarray(fval, dtype=np.float64)
            st = lib().tlo_run_one_fval(C.byref(self._g), int(u), int(v), C.byref(p), fv.ctypes.data, img.ctypes.data, C.byref(det))
        else:
            st = lib().tlo_run_one(C.byref(self._g), int(u), int(v), C.byref(p), img.ctypes.data, C.byref(det))
        n, m, npairs = det.n, det.m, det.npairs
        out = dict(status=st, n=n, m=m, lu=det.lu, lv=det.lv, img=img, npos=det.npos, nneg=det.nneg)
        for k in ("vert", "d1", "d2", "fval"):
            out[k] = bufs[k][:n].copy()
        for k in ("elo", "ehi", "ew", "ord_asc", "ord_desc"):
            out[k] = bufs[k][:m].copy()
        for k in ("pkind", "pbv", "pdv", "pbirth", "pdeath"):
            out[k] = bufs[k][:npairs].copy()
        out["pos"] = bufs["pos"][:det.npos].copy()
        out["neg"] = bufs["neg"][:det.nneg].copy()
        return out


def pimg_transform(dgm, resolution=5):
    d = np.ascontiguousarray(dgm, dtype=np.float64).reshape(-1, 2)
    img = np.zeros(resolution * resolution)
    lib().tlo_pimg_transform(d.ctypes.data, d.shape[0], resolution, img.ctypes.data)
    return img.reshape(resolution, resolution)


def max_threads():
    return int(lib().tlo_max_threads())
