"""fme_b200 -- Python host-side mirror of the C ABI in include/fme_b200.h.

The product is libfme_b200.so (hand-written sm_100a CUDA kernels behind a C ABI); this module is the thin
ctypes binding the tests and bench use, mirroring the reference's operator surface for the fractional-ME
path (names follow HM: frac_dif = TEncSearch::xPatternSearchFracDIF, nn_pred = NN_pred, filter_hor/ver =
TComInterpolationFilter::filterHor/filterVer, dist = DistParam::DistFunc).

There is NO CPU fallback: importing works anywhere (so host-side helpers can be tested), but creating an
`Fme` context raises unless the CUDA library loads and an sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

from . import nn_weights  # noqa: F401  (re-export)
from . import formats  # noqa: F401
from .pu_list import PU_DTYPE, HEAD_DTYPE, GRID_DTYPE, COMPACT_DTYPE, RESULT_DTYPE, MC_PU_DTYPE, MC_BI_PU_DTYPE, CAND_DTYPE  # noqa: F401

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# FME_B200_LIB selects another build of the same library (A/B builds for profiling, e.g. variants/libfme_swar8.so)
LIB_PATH = os.environ.get("FME_B200_LIB") or os.path.join(PKG_DIR, "libfme_b200.so")

MODE_STD, MODE_NN, MODE_BOTH = 1, 2, 3
K2_PATH_AUTO, K2_PATH_SWAR, K2_PATH_MMA_PACK, K2_PATH_MMA_GROUP, K2_PATH_UMMA = 0, 1, 2, 3, 4
K1_PATH_AUTO, K1_PATH_DP4A, K1_PATH_MMA, K1_PATH_UMMA = 0, 1, 2, 3
PU_LOSSLESS, PU_ERR_ON_GPU, PU_BI = 0x01, 0x02, 0x04
CAND_SAD = 0x08

# every symbol include/fme_b200.h declares (checked by tests/test_abi.py against the header)
EXPORTS = [
    "fme_create", "fme_destroy", "fme_last_error", "fme_version", "fme_set_stream", "fme_synchronize",
    "fme_set_nn_weights", "fme_load_nn_csv_dir", "fme_set_slice", "fme_upload_ref", "fme_upload_ref_u8",
    "fme_upload_org", "fme_upload_org_u8", "fme_submit", "fme_submit_async", "fme_submit_heads", "fme_submit_heads_async", "fme_submit_heads_grids", "fme_submit_heads_grids_async", "fme_submit_compact", "fme_submit_compact_async", "fme_wait_oldest", "fme_submit_device", "fme_interp_slot",
    "fme_upload_ref_device_u8", "fme_upload_ref_device_u8_rows", "fme_upload_org_device_u8", "fme_upload_org_device_u8_rows", "fme_int_surface_device", "fme_filter_hor",
    "fme_filter_ver", "fme_dist", "fme_mv_cost", "fme_upload_ref_chroma", "fme_upload_ref_chroma_u8", "fme_upload_ref_yuv420_u8", "fme_upload_org_yuv420_u8", "fme_mc", "fme_mc_bi", "fme_pred_error", "fme_cand_cost", "fme_cand_cost_device", "fme_mc_luma_compact",
    "fme_mc_luma_compact_device", "fme_download_plane",
    "fme_last_kernel_ms", "fme_launch_count", "fme_set_profiling",
]


MODE_RESULT8 = 0x10   # FME_MODE_RESULT8: 8-byte results (fme_result8)
RESULT8_DTYPE = np.dtype([("cost", "<u4"), ("mv", "<u4")])


def unpack_result8(r8):
    """fme_result8 records -> fme_result records (fme_result8_unpack of include/fme_b200.h)."""
    out = np.zeros(len(r8), RESULT_DTYPE)
    m = r8["mv"].astype(np.int64)
    out["cost"] = r8["cost"]
    for k, f in enumerate(("halfX", "halfY", "qterX", "qterY", "nnHalfX", "nnHalfY", "nnQterX", "nnQterY")):
        out[f] = ((m >> (2 * k)) & 3) - 1
    out["nnClass"] = (m >> 16) & 63
    return out


class FmeConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("margin", C.c_int32),
                ("bitDepth", C.c_int32), ("numRefSlots", C.c_int32), ("maxPUs", C.c_int32), ("useHadME", C.c_int32),
                ("fen", C.c_int32), ("nnFma", C.c_int32), ("biPred", C.c_int32), ("k2Path", C.c_int32),
                ("k1Path", C.c_int32), ("k3Fuse", C.c_int32), ("reserved", C.c_int32 * 2)]


class FmeError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libfme_b200.so; fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FmeError("libfme_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                       "this engine has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, sp, u8p = C.c_void_p, C.c_int, C.POINTER(C.c_short), C.POINTER(C.c_uint8)
    lib.fme_last_error.restype = C.c_char_p
    lib.fme_version.restype = C.c_char_p
    lib.fme_launch_count.restype = C.c_int64
    lib.fme_launch_count.argtypes = [vp]
    lib.fme_create.argtypes = [C.POINTER(FmeConfig), C.POINTER(vp)]
    lib.fme_destroy.argtypes = [vp]
    lib.fme_destroy.restype = None
    lib.fme_set_stream.argtypes = [vp, vp]
    lib.fme_synchronize.argtypes = [vp]
    lib.fme_set_nn_weights.argtypes = [vp, vp, C.c_size_t]
    lib.fme_load_nn_csv_dir.argtypes = [vp, C.c_char_p]
    lib.fme_set_slice.argtypes = [vp, C.c_double]
    lib.fme_upload_ref.argtypes = [vp, i32, vp, i32]
    lib.fme_upload_ref_u8.argtypes = [vp, i32, vp, i32]
    lib.fme_upload_ref_yuv420_u8.argtypes = [vp, i32, vp, i32]
    lib.fme_upload_org_yuv420_u8.argtypes = [vp, vp]
    lib.fme_upload_org.argtypes = [vp, vp, i32]
    lib.fme_upload_org_u8.argtypes = [vp, vp, i32]
    lib.fme_submit.argtypes = [vp, vp, i32, vp, i32]
    lib.fme_submit_async.argtypes = [vp, vp, i32, vp, i32]
    lib.fme_submit_heads.argtypes = [vp, vp, i32, vp, i32]
    lib.fme_submit_heads_async.argtypes = [vp, vp, i32, vp, i32]
    lib.fme_submit_heads_grids.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.fme_submit_heads_grids_async.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.fme_submit_compact.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.fme_submit_compact_async.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.fme_submit_device.argtypes = [vp, vp, i32, vp, i32]
    lib.fme_cand_cost.argtypes = [vp, vp, i32, vp, vp]
    lib.fme_cand_cost_device.argtypes = [vp, vp, i32, vp, vp]
    lib.fme_mc_luma_compact.argtypes = [vp, vp, i32, vp, vp, C.c_size_t]
    lib.fme_mc_luma_compact_device.argtypes = [vp, vp, i32, vp, vp]
    lib.fme_wait_oldest.argtypes = [vp]
    lib.fme_interp_slot.argtypes = [vp, i32]
    lib.fme_upload_ref_device_u8.argtypes = [vp, i32, vp, i32]
    lib.fme_upload_ref_device_u8_rows.argtypes = [vp, i32, vp, i32, i32, i32]
    lib.fme_upload_org_device_u8.argtypes = [vp, vp, i32]
    lib.fme_upload_org_device_u8_rows.argtypes = [vp, vp, i32, i32, i32]
    lib.fme_int_surface_device.argtypes = [vp, vp, i32]
    lib.fme_filter_hor.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32]
    lib.fme_filter_ver.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32]
    lib.fme_dist.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.fme_mv_cost.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_uint32)]
    lib.fme_upload_ref_chroma.argtypes = [vp, i32, vp, vp, i32]
    lib.fme_mc.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.fme_mc_bi.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.fme_pred_error.argtypes = [vp, vp, i32, vp]
    lib.fme_download_plane.argtypes = [vp, i32, i32, i32, vp, i32]
    lib.fme_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
    lib.fme_set_profiling.argtypes = [vp, i32]
    _lib = lib
    return lib


def _addr(a, off_elems=0):
    return C.c_void_p(a.ctypes.data + off_elems * a.itemsize)


class Fme:
    """One engine context (= one encoder instance's TEncSearch for the fractional-ME path)."""

    def __init__(self, width, height, num_ref_slots=4, max_pus=1 << 20, margin=80, use_had=True, fen=True, device=0,
                 nn_fma=False, bi_pred=False, k2_path=None, k1_path=None, k3_fuse=None):
        self.lib = load_library()
        if k2_path is None:   # FME_K2_PATH=1|2|3 runs a whole test / bench session on one K2 path (all are bit-identical)
            k2_path = int(os.environ.get("FME_K2_PATH", "0"))
        if k1_path is None:   # FME_K1_PATH=1|2 likewise for the plane builder
            k1_path = int(os.environ.get("FME_K1_PATH", "0"))
        if k3_fuse is None:   # FME_K3_FUSE=1 folds K3's work items into the K2 kernel (fme_config.k3Fuse, experimental)
            k3_fuse = os.environ.get("FME_K3_FUSE", "0") == "1"
        self.cfg = FmeConfig(device=device, width=width, height=height, margin=margin, bitDepth=8, k3Fuse=1 if k3_fuse else 0,
                             numRefSlots=num_ref_slots, maxPUs=max_pus, useHadME=int(use_had), fen=int(fen),
                             nnFma=int(nn_fma), biPred=int(bi_pred), k2Path=int(k2_path), k1Path=int(k1_path))
        self.h = C.c_void_p()
        self._check(self.lib.fme_create(C.byref(self.cfg), C.byref(self.h)))
        self.width, self.height, self.margin = width, height, margin

    def _check(self, rc):
        if rc != 0:
            raise FmeError("fme error %d: %s" % (rc, self.lib.fme_last_error().decode()))

    def close(self):
        if self.h:
            self.lib.fme_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----
    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.fme_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        self._check(self.lib.fme_synchronize(self.h))

    def set_nn_weights(self, blob):
        buf = C.create_string_buffer(bytes(blob), len(blob))
        self._check(self.lib.fme_set_nn_weights(self.h, C.cast(buf, C.c_void_p), len(blob)))

    def load_nn_csv_dir(self, path):
        self._check(self.lib.fme_load_nn_csv_dir(self.h, path.encode()))

    def set_slice(self, lam):
        self._check(self.lib.fme_set_slice(self.h, float(lam)))

    def set_profiling(self, on):
        self._check(self.lib.fme_set_profiling(self.h, int(on)))

    # ---- frames ----
    def upload_ref(self, slot, pic):
        """pic: (H, W) int16 (Pel) or uint8 array, picture area only."""
        pic = np.ascontiguousarray(pic)
        assert pic.shape == (self.height, self.width)
        if pic.dtype == np.uint8:
            self._check(self.lib.fme_upload_ref_u8(self.h, slot, _addr(pic), pic.shape[1]))
        else:
            assert pic.dtype == np.int16
            self._check(self.lib.fme_upload_ref(self.h, slot, _addr(pic), pic.shape[1]))

    def upload_ref_padded(self, slot, padded, margin):
        """padded: TComPicYuv-style int16 plane with `margin` replicated samples around the picture."""
        assert padded.dtype == np.int16 and padded.flags.c_contiguous
        stride = padded.shape[1]
        self._check(self.lib.fme_upload_ref(self.h, slot, _addr(padded, margin * stride + margin), stride))

    def upload_org(self, pic):
        pic = np.ascontiguousarray(pic)
        assert pic.shape == (self.height, self.width)
        if pic.dtype == np.uint8:
            self._check(self.lib.fme_upload_org_u8(self.h, _addr(pic), pic.shape[1]))
        else:
            assert pic.dtype == np.int16
            self._check(self.lib.fme_upload_org(self.h, _addr(pic), pic.shape[1]))

    def upload_ref_device_u8(self, slot, dev_ptr, pitch):
        self._check(self.lib.fme_upload_ref_device_u8(self.h, slot, C.c_void_p(dev_ptr), pitch))

    def upload_org_device_u8(self, dev_ptr, pitch):
        self._check(self.lib.fme_upload_org_device_u8(self.h, C.c_void_p(dev_ptr), pitch))

    def upload_org_device_u8_rows(self, dev_ptr, pitch, row_begin, row_end):
        self._check(self.lib.fme_upload_org_device_u8_rows(self.h, C.c_void_p(dev_ptr), pitch, int(row_begin), int(row_end)))

    def upload_ref_device_u8_rows(self, slot, d_ptr, pitch, row_begin, row_end):
        self._check(self.lib.fme_upload_ref_device_u8_rows(self.h, slot, C.c_void_p(d_ptr), pitch, int(row_begin), int(row_end)))

    def interp_slot(self, slot):
        self._check(self.lib.fme_interp_slot(self.h, slot))

    def upload_ref_chroma(self, slot, cb, cr):
        cb, cr = np.ascontiguousarray(cb, np.int16), np.ascontiguousarray(cr, np.int16)
        assert cb.shape == cr.shape == (self.height // 2, self.width // 2)
        self._check(self.lib.fme_upload_ref_chroma(self.h, slot, _addr(cb), _addr(cr), cb.shape[1]))

    def upload_ref_yuv420(self, slot, frame, with_chroma=True):
        """frame: one raw 8-bit 4:2:0 frame (uint8, W*H*3/2 bytes) exactly as read from a .yuv file."""
        frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
        assert frame.size == self.width * self.height * 3 // 2
        self._check(self.lib.fme_upload_ref_yuv420_u8(self.h, slot, _addr(frame), int(with_chroma)))

    def upload_org_yuv420(self, frame):
        frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
        assert frame.size >= self.width * self.height
        self._check(self.lib.fme_upload_org_yuv420_u8(self.h, _addr(frame)))

    def download_plane(self, slot, fy, fx):
        wp, hp = self.width + 2 * self.margin, self.height + 2 * self.margin
        out = np.zeros((hp, wp), np.uint8)
        self._check(self.lib.fme_download_plane(self.h, slot, fy, fx, _addr(out), wp))
        return out

    # ---- search ----
    def submit(self, pus, mode=MODE_BOTH):
        pus = np.ascontiguousarray(pus, dtype=PU_DTYPE)
        if mode & MODE_RESULT8:
            out8 = np.zeros(len(pus), RESULT8_DTYPE)
            self._check(self.lib.fme_submit(self.h, _addr(pus), len(pus), _addr(out8), mode))
            return unpack_result8(out8)
        out = np.zeros(len(pus), RESULT_DTYPE)
        self._check(self.lib.fme_submit(self.h, _addr(pus), len(pus), _addr(out), mode))
        return out

    def submit_heads(self, heads, mode=MODE_BOTH):
        """Records without err[]: the engine computes the 3x3 integer error surface itself (K0)."""
        heads = np.ascontiguousarray(heads, dtype=HEAD_DTYPE)
        out = np.zeros(len(heads), RESULT_DTYPE)
        self._check(self.lib.fme_submit_heads(self.h, _addr(heads), len(heads), _addr(out), mode))
        return out

    def submit_heads_grids(self, heads, grids, mode=MODE_BOTH):
        """Heads plus the caller's own error grids for some of them (GRID_DTYPE: pu index, err[9]); K0 fills the rest."""
        heads = np.ascontiguousarray(heads, dtype=HEAD_DTYPE)
        grids = np.ascontiguousarray(grids, dtype=GRID_DTYPE)
        out = np.zeros(len(heads), RESULT_DTYPE)
        self._check(self.lib.fme_submit_heads_grids(self.h, _addr(heads), len(heads), _addr(grids), len(grids), _addr(out), mode))
        return out

    def submit_compact(self, recs, big=None, mode=MODE_BOTH):
        """44-byte records (COMPACT_DTYPE: head + nine 24-bit grid values) plus full grids for the PUs that need 32 bits."""
        recs = np.ascontiguousarray(recs, dtype=COMPACT_DTYPE)
        big = np.ascontiguousarray(big if big is not None else np.zeros(0, GRID_DTYPE), dtype=GRID_DTYPE)
        out = np.zeros(len(recs), RESULT_DTYPE)
        self._check(self.lib.fme_submit_compact(self.h, _addr(recs), len(recs), _addr(big) if len(big) else None, len(big),
                                                _addr(out), mode))
        return out

    def submit_compact_async(self, recs_ptr, n, big_ptr, n_big, out_ptr, mode=MODE_BOTH):
        self._check(self.lib.fme_submit_compact_async(self.h, C.c_void_p(recs_ptr), n, C.c_void_p(big_ptr) if n_big else None, n_big,
                                                      C.c_void_p(out_ptr), mode))

    def submit_heads_grids_async(self, heads_ptr, n, grids_ptr, n_grids, out_ptr, mode=MODE_BOTH):
        self._check(self.lib.fme_submit_heads_grids_async(self.h, C.c_void_p(heads_ptr), n, C.c_void_p(grids_ptr), n_grids,
                                                          C.c_void_p(out_ptr), mode))

    def submit_heads_async(self, heads_ptr, n, out_ptr, mode=MODE_BOTH):
        self._check(self.lib.fme_submit_heads_async(self.h, C.c_void_p(heads_ptr), n, C.c_void_p(out_ptr), mode))

    def submit_async(self, pus_ptr, n, out_ptr, mode=MODE_BOTH):
        self._check(self.lib.fme_submit_async(self.h, C.c_void_p(pus_ptr), n, C.c_void_p(out_ptr), mode))

    def wait_oldest(self):
        self._check(self.lib.fme_wait_oldest(self.h))

    def submit_device(self, d_pus_ptr, n, d_out_ptr, mode=MODE_BOTH):
        self._check(self.lib.fme_submit_device(self.h, C.c_void_p(d_pus_ptr), n, C.c_void_p(d_out_ptr), mode))

    def int_surface_device(self, d_pus_ptr, n):
        self._check(self.lib.fme_int_surface_device(self.h, C.c_void_p(d_pus_ptr), n))

    # ---- block-level (reference operator names) ----
    def filter_hor(self, comp, src, src_off, src_stride, w, h, frac, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self._check(self.lib.fme_filter_hor(self.h, comp, _addr(src, src_off), src_stride, _addr(dst), w, w, h, frac,
                                            int(is_last), bit_depth))
        return dst

    def filter_ver(self, comp, src, src_off, src_stride, w, h, frac, is_first, is_last, bit_depth=8):
        dst = np.zeros((h, w), np.int16)
        self._check(self.lib.fme_filter_ver(self.h, comp, _addr(src, src_off), src_stride, _addr(dst), w, w, h, frac,
                                            int(is_first), int(is_last), bit_depth))
        return dst

    def dist(self, kind, org_blocks, cur_blocks, w, h, bit_depth=8, sub_shift=0):
        """org_blocks/cur_blocks: (n, h, stride) int16 arrays."""
        org_blocks = np.ascontiguousarray(org_blocks, np.int16)
        cur_blocks = np.ascontiguousarray(cur_blocks, np.int16)
        n = org_blocks.shape[0]
        out = np.zeros(n, np.uint32)
        self._check(self.lib.fme_dist(self.h, kind, _addr(org_blocks), org_blocks.shape[2], _addr(cur_blocks),
                                      cur_blocks.shape[2], w, h, bit_depth, sub_shift, n, _addr(out)))
        return out

    def mv_cost(self, x, y, scale, px, py):
        v = C.c_uint32()
        self._check(self.lib.fme_mv_cost(self.h, x, y, scale, px, py, C.byref(v)))
        return v.value

    def mc(self, pus, chroma=True):
        pus = np.ascontiguousarray(pus, dtype=MC_PU_DTYPE)
        n = len(pus)
        y = np.zeros((n, 64, 64), np.int16)
        cb = np.zeros((n, 32, 32), np.int16)
        cr = np.zeros((n, 32, 32), np.int16)
        self._check(self.lib.fme_mc(self.h, _addr(pus), n, _addr(y), _addr(cb) if chroma else None,
                                    _addr(cr) if chroma else None))
        return y, cb, cr

    def mc_bi(self, pus, chroma=True):
        """xPredInterBi with both lists valid: two 14-bit uni predictions averaged by addAvg."""
        pus = np.ascontiguousarray(pus, dtype=MC_BI_PU_DTYPE)
        n = len(pus)
        y = np.zeros((n, 64, 64), np.int16)
        cb = np.zeros((n, 32, 32), np.int16)
        cr = np.zeros((n, 32, 32), np.int16)
        self._check(self.lib.fme_mc_bi(self.h, _addr(pus), n, _addr(y), _addr(cb) if chroma else None,
                                       _addr(cr) if chroma else None))
        return y, cb, cr

    def pred_error(self, pus):
        pus = np.ascontiguousarray(pus, dtype=MC_PU_DTYPE)
        out = np.zeros(len(pus), np.uint32)
        self._check(self.lib.fme_pred_error(self.h, _addr(pus), len(pus), _addr(out)))
        return out

    def cand_cost(self, cands, want_best=True):
        """fme_cand_cost: (cost[n], bestIndex[n] or None) for AMVP-template / merge candidates."""
        cands = np.ascontiguousarray(cands, dtype=CAND_DTYPE)
        cost = np.zeros(len(cands), np.uint32)
        best = np.zeros(len(cands), np.int32) if want_best else None
        self._check(self.lib.fme_cand_cost(self.h, _addr(cands), len(cands), _addr(cost), _addr(best) if want_best else None))
        return cost, best

    def cand_cost_device(self, d_cands_ptr, n, d_cost_ptr, d_best_ptr=0):
        self._check(self.lib.fme_cand_cost_device(self.h, C.c_void_p(d_cands_ptr), n, C.c_void_p(d_cost_ptr),
                                                  C.c_void_p(d_best_ptr) if d_best_ptr else None))

    def mc_luma_compact(self, pus):
        """fme_mc_luma_compact: (flat uint8 output, offsets); block i = out[offsets[i] : offsets[i] + w*h].reshape(h, w)."""
        pus = np.ascontiguousarray(pus, dtype=MC_PU_DTYPE)
        sizes = pus["w"].astype(np.int64) * pus["h"]
        offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint32)
        out = np.zeros(int(sizes.sum()), np.uint8)
        self._check(self.lib.fme_mc_luma_compact(self.h, _addr(pus), len(pus), _addr(offsets), _addr(out), C.c_size_t(out.size)))
        return out, offsets

    def mc_luma_compact_device(self, d_pus_ptr, n, d_offsets_ptr, d_out_ptr):
        self._check(self.lib.fme_mc_luma_compact_device(self.h, C.c_void_p(d_pus_ptr), n, C.c_void_p(d_offsets_ptr),
                                                        C.c_void_p(d_out_ptr)))

    # ---- introspection ----
    def last_kernel_ms(self):
        k = (C.c_float * 4)()
        self._check(self.lib.fme_last_kernel_ms(self.h, C.byref(k)))
        return dict(k1_interp=k[0], k2_refine=k[1], k3_nn=k[2], k0_surface=k[3])

    def launch_count(self):
        return int(self.lib.fme_launch_count(self.h))
