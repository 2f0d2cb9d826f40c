"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  Oracle     -> oracle/libfme_oracle.so   plain-C restatement (always available; built on demand)
  Reference  -> oracle/_ref/libhmref.so   the reference's own compiled objects (prebuilt in the dev
                                          container where /root/reference exists; travels to the
                                          GPU box as a built .so; may be absent -> None)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_short_p = C.POINTER(C.c_short)
c_uint_p = C.POINTER(C.c_uint)

PU_DTYPE = np.dtype(
    [("x", "<i2"), ("y", "<i2"), ("w", "u1"), ("h", "u1"), ("refSlot", "u1"), ("flags", "u1"),
     ("mvIntX", "<i2"), ("mvIntY", "<i2"), ("mvPredX", "<i2"), ("mvPredY", "<i2"), ("err", "<u4", (9,))],
    align=True)
RESULT_DTYPE = np.dtype(
    [("halfX", "i1"), ("halfY", "i1"), ("qterX", "i1"), ("qterY", "i1"), ("cost", "<u4"),
     ("nnHalfX", "i1"), ("nnHalfY", "i1"), ("nnQterX", "i1"), ("nnQterY", "i1"), ("nnClass", "u1"),
     ("pad", "u1", (3,))],
    align=True)
assert PU_DTYPE.itemsize == 52 and RESULT_DTYPE.itemsize == 16


def _ptr(a, off_elems=0):
    """short* into an int16 ndarray at a flat element offset (may be negative-index safe via base addr)."""
    assert a.dtype == np.int16
    return C.cast(a.ctypes.data + 2 * off_elems, c_short_p)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libfme_oracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("fme_oracle.c", "fme_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libfme_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_reference():
    """Build oracle/_ref/libhmref.so when the reference tree is present; returns path or None."""
    so = os.path.join(ORACLE_DIR, "_ref", "libhmref.so")
    if os.path.isdir("/root/reference/source/Lib"):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-j8", "ref"], stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


class Oracle:
    """Plain-C restatement (oracle/fme_oracle.c)."""

    def __init__(self):
        L = self.L = C.CDLL(build_oracle())
        L.orc_motion_lambda.restype = C.c_double
        L.orc_motion_lambda.argtypes = [C.c_double]
        L.orc_slice_lambda.restype = C.c_double
        L.orc_slice_lambda.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_mv_cost.restype = C.c_uint
        L.orc_mv_cost.argtypes = [C.c_double] + [C.c_int] * 5
        L.orc_exp_golomb_bits.restype = C.c_uint
        L.orc_dist.restype = C.c_uint
        L.orc_dist.argtypes = [C.c_int, c_short_p, C.c_int, c_short_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_frac_dif.argtypes = [c_short_p, C.c_int, C.c_int, C.c_int, c_short_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_short_p, c_short_p, c_uint_p]
        L.orc_filtered_block.restype = c_short_p
        L.orc_nn_pred.restype = C.c_int
        L.orc_nn_pred.argtypes = [C.c_void_p, c_uint_p, C.c_int, C.c_int, C.POINTER(C.c_float), c_short_p, c_short_p]
        L.orc_nn_pred_f64.restype = C.c_int
        L.orc_nn_pred_f64.argtypes = [C.c_void_p, C.POINTER(C.c_double), c_uint_p, C.POINTER(C.c_double)]
        L.orc_run_pu_list.argtypes = [c_short_p, C.c_int, C.POINTER(c_short_p), C.c_int, C.c_void_p, C.c_int,
                                      C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_subpel_plane.argtypes = [c_short_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       c_short_p, C.c_int]
        L.orc_fill_surface.argtypes = [c_short_p, C.c_int, C.POINTER(c_short_p), C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_int_surface.argtypes = [c_short_p, C.c_int, C.c_int, C.c_int, c_short_p, C.c_int, C.c_int, c_uint_p]

    def filter_hor(self, is_luma, src, src_off, sstride, w, h, frac, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self.L.orc_filter_hor(int(is_luma), _ptr(src, src_off), sstride, _ptr(dst), w, w, h, frac, int(is_last),
                              bit_depth)
        return dst

    def filter_ver(self, is_luma, src, src_off, sstride, w, h, frac, is_first, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self.L.orc_filter_ver(int(is_luma), _ptr(src, src_off), sstride, _ptr(dst), w, w, h, frac, int(is_first),
                              int(is_last), bit_depth)
        return dst

    def bi_pattern(self, org, org_off, ostride, ref, ref_off, rstride, w, h, mvx, mvy):
        """2*org - xPredInterBlk(ref at quarter-pel mv): the search pattern of a bi-predictive refinement call."""
        pred = np.zeros((h, w), np.int16)
        self.L.orc_mc_luma(_ptr(ref, ref_off + (mvy >> 2) * rstride + (mvx >> 2)), rstride, _ptr(pred), w, w, h,
                           mvx & 3, mvy & 3)
        pat = np.zeros((h, w), np.int16)
        self.L.orc_bi_pattern(_ptr(org, org_off), ostride, _ptr(pred), w, _ptr(pat), w, w, h)
        return pat

    def add_avg(self, a, a_off, astride, b, b_off, bstride, w, h, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self.L.orc_add_avg(_ptr(a, a_off), astride, _ptr(b, b_off), bstride, _ptr(dst), w, w, h, bit_depth)
        return dst

    def dist(self, kind, org, org_off, ostride, cur, cur_off, cstride, w, h, bit_depth=8, sub_shift=0):
        return int(self.L.orc_dist(kind, _ptr(org, org_off), ostride, _ptr(cur, cur_off), cstride, w, h, bit_depth,
                                   sub_shift))

    def mv_cost(self, lam, x, y, scale, px, py):
        return int(self.L.orc_mv_cost(self.L.orc_motion_lambda(lam), x, y, scale, px, py))

    def frac_dif(self, org, org_off, ostride, w, h, ref, ref_off, rstride, mvx, mvy, px, py, lam, use_had=1,
                 lossless=0):
        hxy = (C.c_short * 2)()
        qxy = (C.c_short * 2)()
        cost = C.c_uint()
        self.L.orc_frac_dif(_ptr(org, org_off), ostride, w, h, _ptr(ref, ref_off), rstride, mvx, mvy, px, py, lam,
                            use_had, lossless, hxy, qxy, C.byref(cost))
        return (hxy[0], hxy[1]), (qxy[0], qxy[1]), cost.value

    def filtered_block(self, v, h, w, hgt):
        p = self.L.orc_filtered_block(v, h)
        a = np.ctypeslib.as_array(p, shape=(66, 80))
        return a[:hgt, :w].copy()

    def subpel_plane(self, ref, ref_off, rstride, x0, y0, w, h, fy, fx):
        out = np.zeros((h, w), np.int16)
        self.L.orc_subpel_plane(_ptr(ref, ref_off), rstride, x0, y0, w, h, fy, fx, _ptr(out), w)
        return out

    def int_surface(self, org, org_off, ostride, w, h, ref, ref_off, rstride, fen=1):
        e = (C.c_uint * 9)()
        self.L.orc_int_surface(_ptr(org, org_off), ostride, w, h, _ptr(ref, ref_off), rstride, fen, e)
        return np.array(list(e), np.uint32)

    def nn_pred(self, blob, err9, h, w):
        e = (C.c_uint * 9)(*[int(v) for v in err9])
        logits = (C.c_float * 64)()
        hxy = (C.c_short * 2)()
        qxy = (C.c_short * 2)()
        buf = C.create_string_buffer(bytes(blob), len(blob))
        cls = self.L.orc_nn_pred(buf, e, h, w, logits, hxy, qxy)
        return cls, np.array(logits[:49], np.float32), (hxy[0], hxy[1]), (qxy[0], qxy[1])

    def nn_pred_f64(self, blob, payload, err9):
        """Double-precision forward of the reference's 3-layer backup network: (class, outputs[49])."""
        e = (C.c_uint * 9)(*[int(v) for v in err9])
        outs = (C.c_double * 64)()
        buf = C.create_string_buffer(bytes(blob[:64]), 64)
        pl = np.ascontiguousarray(payload, np.float64)
        cls = self.L.orc_nn_pred_f64(buf, pl.ctypes.data_as(C.POINTER(C.c_double)), e, outs)
        return cls, np.array(outs[:49])

    def fill_surface(self, org, ostride, refs, ref_offs, rstride, pus, fen=1):
        arr = (c_short_p * len(refs))(*[_ptr(r, o) for r, o in zip(refs, ref_offs)])
        assert pus.flags.c_contiguous and pus.dtype == PU_DTYPE
        self.L.orc_fill_surface(_ptr(org), ostride, arr, rstride, C.c_void_p(pus.ctypes.data), len(pus), fen)
        return pus

    def run_pu_list(self, org, ostride, refs, ref_offs, rstride, pus, mode, lam, use_had, blob):
        """org: int16 plane with picture (0,0) at flat offset 0; refs: list of padded int16 planes."""
        n = len(pus)
        out = np.zeros(n, RESULT_DTYPE)
        arr = (c_short_p * len(refs))(*[_ptr(r, o) for r, o in zip(refs, ref_offs)])
        buf = C.create_string_buffer(bytes(blob), len(blob)) if blob is not None else None
        pus = np.ascontiguousarray(pus)
        self.L.orc_run_pu_list(_ptr(org), ostride, arr, rstride, pus.ctypes.data, n, mode, lam, use_had, buf,
                               out.ctypes.data)
        return out


class Reference:
    """The reference's own compiled code (oracle/_ref/libhmref.so)."""

    def __init__(self, path):
        L = self.L = C.CDLL(path)
        L.hmref_set_lambda.argtypes = [C.c_double]
        L.hmref_dist.restype = C.c_uint
        L.hmref_mv_cost.restype = C.c_uint
        L.hmref_run_pu_list.argtypes = [c_short_p, C.c_int, C.POINTER(c_short_p), C.c_int, C.c_void_p, C.c_int,
                                        C.c_int, C.c_void_p]
        self._state = None

    def init(self, qp=22, use_had=1, fen=1):
        if self._state != (qp, use_had, fen):
            self.L.hmref_init(qp, use_had, fen)
            self._state = (qp, use_had, fen)

    def set_lambda(self, lam):
        self.L.hmref_set_lambda(float(lam))

    def filter_hor(self, is_luma, src, src_off, sstride, w, h, frac, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self.L.hmref_filter_hor(0 if is_luma else 1, _ptr(src, src_off), sstride, _ptr(dst), w, w, h, frac,
                                int(is_last), bit_depth)
        return dst

    def filter_ver(self, is_luma, src, src_off, sstride, w, h, frac, is_first, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self.L.hmref_filter_ver(0 if is_luma else 1, _ptr(src, src_off), sstride, _ptr(dst), w, w, h, frac,
                                int(is_first), int(is_last), bit_depth)
        return dst

    def add_avg(self, a, a_off, astride, b, b_off, bstride, w, h):
        """TComYuv::addAvg at the bit depth of the last init() (8)."""
        dst = np.zeros((h, w), np.int16)
        self.L.hmref_add_avg(_ptr(a, a_off), astride, _ptr(b, b_off), bstride, _ptr(dst), w, w, h)
        return dst

    def dist(self, kind, org, org_off, ostride, cur, cur_off, cstride, w, h, bit_depth=8, sub_shift=0):
        return int(self.L.hmref_dist(kind, _ptr(org, org_off), ostride, _ptr(cur, cur_off), cstride, w, h, bit_depth,
                                     sub_shift))

    def mv_cost(self, x, y, scale, px, py):
        return int(self.L.hmref_mv_cost(x, y, scale, px, py))

    def frac_dif(self, org, org_off, ostride, w, h, ref, ref_off, rstride, mvx, mvy, px, py, lossless=0):
        hxy = (C.c_short * 2)()
        qxy = (C.c_short * 2)()
        cost = C.c_uint()
        self.L.hmref_frac_dif(_ptr(org, org_off), ostride, w, h, _ptr(ref, ref_off), rstride, mvx, mvy, px, py,
                              lossless, hxy, qxy, C.byref(cost))
        return (hxy[0], hxy[1]), (qxy[0], qxy[1]), cost.value

    def filtered_block(self, v, h, w, hgt):
        dst = np.zeros((hgt, w), np.int16)
        self.L.hmref_get_filtered_block(v, h, _ptr(dst), w, w, hgt)
        return dst

    def int_surface(self, org, org_off, ostride, w, h, ref, ref_off, rstride):
        e = (C.c_uint * 9)()
        self.L.hmref_int_surface(_ptr(org, org_off), ostride, w, h, _ptr(ref, ref_off), rstride, e)
        return np.array(list(e), np.uint32)

    def nn_pred(self, err9, h, w):
        e = (C.c_uint * 9)(*[int(v) for v in err9])
        cls = C.c_int()
        hxy = (C.c_short * 2)()
        qxy = (C.c_short * 2)()
        self.L.hmref_nn_pred(e, h, w, C.byref(cls), hxy, qxy)
        return cls.value, (hxy[0], hxy[1]), (qxy[0], qxy[1])

    def run_pu_list(self, org, ostride, refs, ref_offs, rstride, pus, mode):
        n = len(pus)
        out = np.zeros(n, RESULT_DTYPE)
        arr = (c_short_p * len(refs))(*[_ptr(r, o) for r, o in zip(refs, ref_offs)])
        pus = np.ascontiguousarray(pus)
        self.L.hmref_run_pu_list(_ptr(org), ostride, arr, rstride, pus.ctypes.data, n, mode, out.ctypes.data)
        return out


_ORACLE = None
_REFERENCE = False


def oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def reference():
    """Reference instance or None (no prebuilt libhmref.so and no /root/reference)."""
    global _REFERENCE
    if _REFERENCE is False:
        p = build_reference()
        _REFERENCE = Reference(p) if p else None
    return _REFERENCE


def pad_plane(pic, margin):
    """TComPicYuv-style padded Pel plane: picture + `margin` edge-replicated samples (TComPicYuv.cpp:229-276)."""
    return np.ascontiguousarray(np.pad(np.asarray(pic).astype(np.int16), margin, mode="edge"))


class CpuFrame:
    """Host-side frame set shared by the oracle / reference runners in the tests and the CPU baseline."""

    def __init__(self, org_u8, refs_u8, margin=80):
        self.margin = margin
        self.h, self.w = org_u8.shape
        self.org = np.ascontiguousarray(org_u8.astype(np.int16))
        self.refs = [pad_plane(r, margin) for r in refs_u8]
        self.rstride = self.w + 2 * margin
        self.ref_offs = [margin * self.rstride + margin] * len(self.refs)

    def oracle_fill_surface(self, pus, fen=1):
        return oracle().fill_surface(self.org, self.w, self.refs, self.ref_offs, self.rstride, pus, fen)

    def oracle_run(self, pus, mode, lam, use_had, blob):
        return oracle().run_pu_list(self.org, self.w, self.refs, self.ref_offs, self.rstride, pus, mode, lam,
                                    int(use_had), blob)

    def reference_run(self, pus, mode):
        return reference().run_pu_list(self.org, self.w, self.refs, self.ref_offs, self.rstride, pus, mode)
