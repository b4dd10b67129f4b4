"""mmo_b200 -- host-side mirror (ctypes) of the reference's scoring interface over libmmo_b200.so.

The product is the C-ABI library `mmo_b200/csrc/libmmo_b200.so` (hand-written sm_100a CUDA).  This
package only binds it for the tests and for bench.py, under the reference's own names:

    Mol.ene_inter_UFF_shifted_brute / _global_brute / _shifted_bst_components   src/mol.ml:796-960
    Mol.ene_intra_UFFNB_brute                                                    src/mol.ml:881-903
    Mol.ene_inter_UFF_interp, G3D.trilin                                         src/mol.ml:1012-1020, src/G3D.ml:97-157
    Lds.pre_calculate_FF_components_grid, Lds.vdW_volume                         src/lds.ml:452-469, 187-196
    Lds.exhaustive_rigid_ligand_docking                                          src/lds.ml:1040-1114
    SO3.rotations, Rot.r_xyz, Rot.decompose, Grid.from_box                       src/SO3.ml, src/rot.ml, src/grid.ml

There is NO CPU fallback: if the shared library is missing or no B200 is visible, calls raise.
No torch import here; numpy arrays in, numpy arrays out.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import pqrs  # noqa: F401  (host-side file formats)

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMO_B200_LIB: kernel-tuning experiments load an alternative build of the same library (tools/variants.sh)
LIB_PATH = os.environ.get("MMO_B200_LIB") or os.path.join(_HERE, "csrc", "libmmo_b200.so")

VARIANT_GLOBAL, VARIANT_SHIFTED = 0, 1      # -ff BrG | BrL/Bst
PREC_FP32, PREC_FP64 = 0, 1

_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class MmoError(RuntimeError):
    pass


class ScanParams(C.Structure):
    _fields_ = [("rec", _vp), ("grid", _vp), ("lig", _vp), ("vdw_mask", _vp),
                ("variant", C.c_int32), ("prec", C.c_int32),
                ("roi_c", C.c_double * 3), ("roi_r", C.c_double), ("trans_step", C.c_double),
                ("n_rot", C.c_int32), ("rot9", _dp), ("e_intra_const", C.c_double), ("topk", C.c_int32),
                ("first_point", C.c_int64), ("n_points", C.c_int64)]


class ScanResult(C.Structure):
    _fields_ = [("n_candidates", C.c_int64), ("n_scored", C.c_int64), ("best_score", C.c_double),
                ("best_frame", C.c_int64), ("best_pos", C.c_double * 3), ("best_rot_i", C.c_int32),
                ("n_top", C.c_int32), ("lattice_dims", C.c_int32 * 3),
                ("pairs_evaluated", C.c_int64), ("pairs_inside", C.c_int64), ("device_ms", C.c_float)]


class McParams(C.Structure):
    _fields_ = [("roi_c", C.c_double * 3), ("roi_r", C.c_double), ("temperature_K", C.c_double),
                ("n_steps", C.c_int32), ("tweak_rbonds", C.c_int32), ("hard_roi", C.c_int32),
                ("no_flip", C.c_int32), ("intra_nb", C.c_int32)]


class McResult(C.Structure):
    _fields_ = [("best_E", C.c_double), ("prev_E", C.c_double), ("best_rot", C.c_double * 9),
                ("best_pos", C.c_double * 3), ("max_rot", C.c_double), ("max_trans", C.c_double),
                ("n_accept_rigid", C.c_int64), ("n_reject_rigid", C.c_int64), ("n_accept_conf", C.c_int64),
                ("n_reject_conf", C.c_int64), ("n_ooroi", C.c_int64), ("n_ezero", C.c_int64),
                ("too_long", C.c_int32), ("frames_done", C.c_int32)]


def lib():
    """The loaded C-ABI library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmoError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                           " (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.mmo_last_error.restype = C.c_char_p
        L.mmo_build_info.restype = C.c_char_p
        L.mmo_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise MmoError(f"libmmo_b200 error {rc}: {lib().mmo_last_error().decode()}")


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


_inited = None


def init(device: int = 0):
    global _inited
    _ck(lib().mmo_init(C.c_int(device)))
    _inited = device


def _need_init():
    if _inited is None:
        init(int(os.environ.get("LOCAL_RANK", os.environ.get("MMO_DEVICE", "0"))))


def launch_count() -> int:
    return int(lib().mmo_launch_count())


class Receptor:
    """Protein atoms in HBM (Mol.t of the receptor, src/mol.ml:17-35)."""

    def __init__(self, xs, ys, zs, q, anum):
        _need_init()
        self.n = len(xs)
        self._keep = [_d(xs), _d(ys), _d(zs), _d(q), _i(anum)]
        self.h = _vp()
        k = self._keep
        _ck(lib().mmo_receptor_create(C.c_int32(self.n), k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], C.byref(self.h)))

    @classmethod
    def from_mol(cls, m):
        return cls(m.xs, m.ys, m.zs, m.q, m.anum)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_receptor_destroy(self.h)
            self.h = None


class Ligand:
    """Ligand template in HBM.  Coordinates must be centred for the *_poses entry points."""

    def __init__(self, xs, ys, zs, q, anum, r=None, typ=None, dists=None, rb_left=None, rb_right=None,
                 rg_off=None, rg_idx=None):
        _need_init()
        self.n = len(xs)
        nrb = 0 if rb_left is None else len(rb_left)
        null_d, null_i = C.cast(None, _dp), C.cast(None, _ip)
        keep = [_d(xs), _d(ys), _d(zs), _d(q), _i(anum)]
        pr = _d(r) if r is not None else (None, null_d)
        pt = _i(typ) if typ is not None else (None, null_i)
        pd = _i(dists) if dists is not None else (None, null_i)
        pl = _i(rb_left) if nrb else (None, null_i)
        prr = _i(rb_right) if nrb else (None, null_i)
        po = _i(rg_off) if nrb else (None, null_i)
        pi = _i(rg_idx) if nrb else (None, null_i)
        self._keep = keep + [pr, pt, pd, pl, prr, po, pi]
        self.h = _vp()
        _ck(lib().mmo_ligand_create(C.c_int32(self.n), keep[0][1], keep[1][1], keep[2][1], keep[3][1], pr[1],
                                    keep[4][1], pt[1], pd[1], C.c_int32(nrb), pl[1], prr[1], po[1], pi[1],
                                    C.byref(self.h)))

    @classmethod
    def from_mol(cls, m, centered=True):
        """Ligand of a pqrs.Mol; `centered` translates it to the origin as lds.ml:44-52 does
        (Mol.translate_to lig V3.origin with the Kahan-averaged centre, mol.ml:353-356)."""
        xs, ys, zs = m.xs, m.ys, m.zs
        if centered:
            c = [favg(m.xs), favg(m.ys), favg(m.zs)]
            xs, ys, zs = m.xs + (0.0 - c[0]), m.ys + (0.0 - c[1]), m.zs + (0.0 - c[2])
        off, idx = m.rgroup_csr()
        o = cls(xs, ys, zs, m.q, m.anum, r=m.r, typ=m.typ, dists=m.dists, rb_left=m.rb_left, rb_right=m.rb_right,
                rg_off=off, rg_idx=idx)
        o.xs, o.ys, o.zs = np.array(xs), np.array(ys), np.array(zs)
        return o

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_ligand_destroy(self.h)
            self.h = None


def favg(a) -> float:
    """Batteries' A.favg as restated in the oracle: Kahan-compensated sum / n (unpinned library)."""
    s = 0.0
    c = 0.0
    for v in np.asarray(a, np.float64):
        y = float(v) - c
        t = s + y
        c = (t - s) - y
        s = t
    return s / len(a)


class EnergyGrid:
    """T float32 energy maps, type-major, resident in HBM (G3D.t array)."""

    def __init__(self, handle, step, dims, T):
        self.h, self.step, self.dims, self.T = handle, float(step), tuple(int(d) for d in dims), int(T)

    @property
    def nvox(self):
        return self.dims[0] * self.dims[1] * self.dims[2]

    def download(self):
        out = np.empty((self.T, self.nvox), np.float32)
        _ck(lib().mmo_grid_download(self.h, out.ctypes.data_as(_fp)))
        return out

    def write_ba1(self, t, path):
        _ck(lib().mmo_grid_write_ba1(self.h, C.c_int32(t), path.encode()))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_grid_destroy(self.h)
            self.h = None


class VdwMask:
    def __init__(self, handle, step, dims, bits=None):
        self.h, self.step, self.dims, self.bits = handle, float(step), tuple(int(d) for d in dims), bits

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_mask_destroy(self.h)
            self.h = None


class Desolv:
    """Lds.protein_desolv's per-voxel contributions, device resident (N4); keeps the shell mask alive"""

    def __init__(self, handle, shell):
        self.h, self.shell = handle, shell

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_desolv_destroy(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------
class Mol:
    """src/mol.ml energy functions, batched: coordinates are arrays [n_poses, L]."""

    @staticmethod
    def _score(rec, lig, variant, prec, xs, ys, zs):
        xs = np.atleast_2d(np.asarray(xs, np.float64))
        ys = np.atleast_2d(np.asarray(ys, np.float64))
        zs = np.atleast_2d(np.asarray(zs, np.float64))
        n = xs.shape[0]
        assert xs.shape == (n, lig.n) == ys.shape == zs.shape
        out = np.empty(n, np.float64)
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        _ck(lib().mmo_score_coords(rec.h, lig.h, C.c_int(variant), C.c_int(prec), C.c_int64(n), pa, pb, pc,
                                   out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def ene_inter_UFF_shifted_brute(rec, lig, xs, ys, zs, prec=PREC_FP32):
        return Mol._score(rec, lig, VARIANT_SHIFTED, prec, xs, ys, zs)

    @staticmethod
    def ene_inter_UFF_global_brute(rec, lig, xs, ys, zs, prec=PREC_FP32):
        return Mol._score(rec, lig, VARIANT_GLOBAL, prec, xs, ys, zs)

    @staticmethod
    def ene_inter_UFF_shifted_bst_components(rec, lig, xs, ys, zs):
        xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
        zs = np.atleast_2d(np.asarray(zs, np.float64))
        n = xs.shape[0]
        e = np.empty(n); v = np.empty(n)
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        _ck(lib().mmo_score_coords_components(rec.h, lig.h, C.c_int64(n), pa, pb, pc, e.ctypes.data_as(_dp),
                                              v.ctypes.data_as(_dp)))
        return e, v

    @staticmethod
    def score_poses(rec, lig, rot9, trans3, variant=VARIANT_SHIFTED, prec=PREC_FP32):
        """rotate_then_translate_copy (mol.ml:669-672) + ene_inter for n poses."""
        rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        trans3 = np.ascontiguousarray(trans3, np.float64).reshape(-1, 3)
        n = rot9.shape[0]
        assert trans3.shape[0] == n
        out = np.empty(n)
        _ck(lib().mmo_score_poses(rec.h, lig.h, C.c_int(variant), C.c_int(prec), C.c_int64(n),
                                  rot9.ctypes.data_as(_dp), trans3.ctypes.data_as(_dp), out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def ene_intra_UFFNB_brute(lig, xs, ys, zs):
        xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
        zs = np.atleast_2d(np.asarray(zs, np.float64))
        n = xs.shape[0]
        out = np.empty(n)
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        _ck(lib().mmo_intra_nb(lig.h, C.c_int64(n), pa, pb, pc, out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def ene_inter_UFF_interp(grid, lig, xs, ys, zs):
        xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
        zs = np.atleast_2d(np.asarray(zs, np.float64))
        n = xs.shape[0]
        out = np.empty(n)
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        _ck(lib().mmo_score_interp_coords(grid.h, lig.h, C.c_int64(n), pa, pb, pc, out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def interp_poses(grid, lig, rot9, trans3):
        rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        trans3 = np.ascontiguousarray(trans3, np.float64).reshape(-1, 3)
        n = rot9.shape[0]
        out = np.empty(n)
        _ck(lib().mmo_score_interp_poses(grid.h, lig.h, C.c_int64(n), rot9.ctypes.data_as(_dp),
                                         trans3.ctypes.data_as(_dp), out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def protein_ligand_clash(mask, lig, rot9, trans3):
        rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        trans3 = np.ascontiguousarray(trans3, np.float64).reshape(-1, 3)
        n = rot9.shape[0]
        out = np.empty(n, np.uint8)
        _ck(lib().mmo_clash_poses(mask.h, lig.h, C.c_int64(n), rot9.ctypes.data_as(_dp), trans3.ctypes.data_as(_dp),
                                  out.ctypes.data_as(_bp)))
        return out.astype(bool)

    @staticmethod
    def last_pair_stats():
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _ck(lib().mmo_last_pair_stats(C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


class Grid:
    @staticmethod
    def from_box(step, bx, by, bz):
        dims = (C.c_int32 * 3)()
        _ck(lib().mmo_grid_from_box(C.c_double(step), C.c_double(bx), C.c_double(by), C.c_double(bz), dims))
        return tuple(dims)


class G3D:
    @staticmethod
    def trilin(grid, t, xs, ys, zs):
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        out = np.empty(len(a))
        _ck(lib().mmo_trilin(grid.h, C.c_int32(t), C.c_int64(len(a)), pa, pb, pc, out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def upload(step, dims, maps):
        maps = np.ascontiguousarray(maps, np.float32)
        T = maps.shape[0]
        h = _vp()
        d = (C.c_int32 * 3)(*dims)
        _ck(lib().mmo_grid_upload(C.c_double(step), d, C.c_int32(T), maps.ctypes.data_as(_fp), C.byref(h)))
        return EnergyGrid(h, step, dims, T)

    @staticmethod
    def of_ba1_files(paths):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        h = _vp()
        _ck(lib().mmo_grid_read_ba1(arr, C.c_int32(len(paths)), C.byref(h)))
        first = paths[0][:-4] if paths[0].endswith(".zst") else paths[0]      # the .dims side-car is never compressed
        step, dims = _parse_dims(first + ".dims")
        return EnergyGrid(h, step, dims, len(paths))


def _parse_dims(fn):
    v = [l.split(":")[1] for l in open(fn).read().strip().split("\n")]
    return float(v[0]), (int(v[1]), int(v[2]), int(v[3]))


class Lds:
    @staticmethod
    def pre_calculate_FF_components_grid(rec, step, dims, type_anum, type_q, mask_bits=None, want_host=True):
        """src/lds.ml:452-469.  Returns (EnergyGrid, host maps or None)."""
        _need_init()
        T = len(type_anum)
        (ta, pta), (tq, ptq) = _i(type_anum), _d(type_q)
        nvox = int(dims[0]) * int(dims[1]) * int(dims[2])
        host = np.empty((T, nvox), np.float32) if want_host else None
        pm = C.cast(None, _bp)
        if mask_bits is not None:
            mask_bits = np.ascontiguousarray(mask_bits, np.uint8)
            assert mask_bits.size >= (nvox + 7) // 8
            pm = mask_bits.ctypes.data_as(_bp)
        h = _vp()
        d = (C.c_int32 * 3)(*[int(v) for v in dims])
        _ck(lib().mmo_grid_build(rec.h, C.c_double(step), d, pm, C.c_int32(T), pta, ptq,
                                 host.ctypes.data_as(_fp) if want_host else C.cast(None, _fp), C.byref(h)))
        return EnergyGrid(h, step, dims, T), host

    @staticmethod
    def vdW_volume(xs, ys, zs, radii, step, dims):
        """src/lds.ml:187-196 -> (VdwMask, bits as uint8 array, LSB-first)."""
        _need_init()
        (a, pa), (b, pb), (c, pc), (r, pr) = _d(xs), _d(ys), _d(zs), _d(radii)
        nvox = int(dims[0]) * int(dims[1]) * int(dims[2])
        bits = np.zeros((nvox + 7) // 8, np.uint8)
        h = _vp()
        d = (C.c_int32 * 3)(*[int(v) for v in dims])
        _ck(lib().mmo_vdw_mask_build(C.c_int32(len(a)), pa, pb, pc, pr, C.c_double(step), d,
                                     bits.ctypes.data_as(_bp), C.byref(h)))
        return VdwMask(h, step, dims, bits)

    @staticmethod
    def _mask_call(fn, dims, step, *args):
        _need_init()
        nvox = int(dims[0]) * int(dims[1]) * int(dims[2])
        bits = np.zeros((nvox + 7) // 8, np.uint8)
        h = _vp()
        dd = (C.c_int32 * 3)(*[int(v) for v in dims])
        _ck(fn(*args, C.c_double(step), dd, bits.ctypes.data_as(_bp), C.byref(h)))
        return VdwMask(h, step, dims, bits)

    @staticmethod
    def first_solvent_shell(xs, ys, zs, radii, step, dims):
        """src/lds.ml:172-184"""
        (a, pa), (b, pb), (c, pc), (r, pr) = _d(xs), _d(ys), _d(zs), _d(radii)
        return Lds._mask_call(lib().mmo_mask_first_solvent_shell, dims, step, C.c_int32(len(a)), pa, pb, pc, pr)

    @staticmethod
    def bitmask_whole_protein(xs, ys, zs, step, dims):
        """src/lds.ml:97-145"""
        (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
        return Lds._mask_call(lib().mmo_mask_whole_protein, dims, step, C.c_int32(len(a)), pa, pb, pc)

    @staticmethod
    def bitmask_ROI_only(roi, step, dims):
        """src/lds.ml:269-305; roi = (cx, cy, cz, r)"""
        c = (C.c_double * 3)(*roi[:3])
        return Lds._mask_call(lib().mmo_mask_roi_only, dims, step, c, C.c_double(roi[3]))

    @staticmethod
    def protein_desolv(roi, rec, prot_solvent_shell, want_host=True):
        """src/lds.ml:204-236; roi = (cx, cy, cz, r), prot_solvent_shell = Lds.first_solvent_shell of the receptor.
        Returns (Desolv handle, per-voxel contributions or None)."""
        _need_init()
        m = prot_solvent_shell
        nvox = m.dims[0] * m.dims[1] * m.dims[2]
        host = np.empty(nvox) if want_host else None
        h = _vp()
        c = (C.c_double * 3)(*roi[:3])
        _ck(lib().mmo_desolv_protein(rec.h, m.h, c, C.c_double(roi[3]),
                                     host.ctypes.data_as(_dp) if want_host else C.cast(None, _dp), C.byref(h)))
        return Desolv(h, m), host

    @staticmethod
    def desolvation_penalty(desolv, lig, xs=None, ys=None, zs=None, rot9=None, trans3=None):
        """src/lds.ml:239-267 for many poses (explicit coordinates [n, L] or rot9/trans3): (prot[n], lig[n])"""
        if xs is not None:
            xs, ys, zs = (np.atleast_2d(np.asarray(a, np.float64)) for a in (xs, ys, zs))
            n = xs.shape[0]
            assert xs.shape == (n, lig.n) == ys.shape == zs.shape
            op, ol = np.empty(n), np.empty(n)
            (a, pa), (b, pb), (c, pc) = _d(xs), _d(ys), _d(zs)
            _ck(lib().mmo_desolv_penalty_coords(desolv.h, lig.h, C.c_int64(n), pa, pb, pc, op.ctypes.data_as(_dp),
                                                ol.ctypes.data_as(_dp)))
            return op, ol
        rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        trans3 = np.ascontiguousarray(trans3, np.float64).reshape(-1, 3)
        n = rot9.shape[0]
        op, ol = np.empty(n), np.empty(n)
        _ck(lib().mmo_desolv_penalty_poses(desolv.h, lig.h, C.c_int64(n), rot9.ctypes.data_as(_dp), trans3.ctypes.data_as(_dp),
                                           op.ctypes.data_as(_dp), ol.ctypes.data_as(_dp)))
        return op, ol

    @staticmethod
    def exhaustive_rigid_ligand_docking(topk, roi, trans_step, rotations, lig, rec=None, grid=None, vdw_mask=None,
                                        e_intra_const=0.0, variant=VARIANT_SHIFTED, prec=PREC_FP32,
                                        first_point=0, n_points=-1):
        """src/lds.ml:1040-1114.  roi = (cx, cy, cz, r) in simulation-box coordinates.
        Returns dict(top_scores, top_frames, best_score, best_frame, best_pos, best_rot_i, ...)."""
        _need_init()
        rot = np.ascontiguousarray(rotations, np.float64).reshape(-1, 9)
        P = ScanParams()
        P.rec = rec.h if rec is not None else None
        P.grid = grid.h if grid is not None else None
        P.lig = lig.h
        P.vdw_mask = vdw_mask.h if vdw_mask is not None else None
        P.variant, P.prec = variant, prec
        P.roi_c = (C.c_double * 3)(*roi[:3])
        P.roi_r = roi[3]
        P.trans_step = trans_step
        P.n_rot = rot.shape[0]
        P.rot9 = rot.ctypes.data_as(_dp)
        P.e_intra_const = e_intra_const
        P.topk = topk
        P.first_point, P.n_points = first_point, n_points
        k = max(topk, 1)
        ts = np.empty(k); tf = np.empty(k, np.int64)
        R = ScanResult()
        _ck(lib().mmo_scan(C.byref(P), ts.ctypes.data_as(_dp), tf.ctypes.data_as(_lp), C.byref(R)))
        return dict(top_scores=ts[:R.n_top].copy(), top_frames=tf[:R.n_top].copy(), best_score=R.best_score,
                    best_frame=R.best_frame, best_pos=tuple(R.best_pos), best_rot_i=R.best_rot_i,
                    n_candidates=R.n_candidates, n_scored=R.n_scored, lattice_dims=tuple(R.lattice_dims),
                    pairs_evaluated=R.pairs_evaluated, pairs_inside=R.pairs_inside, device_ms=R.device_ms)


    @staticmethod
    def simulate_lig(grid, lig, roi, n_steps, seeds, start_rot9, start_pos3, tweak_rbonds=True, hard_roi=True,
                     no_flip=False, intra_nb=True, temperature_K=293.15, want_xyz=False, want_trace=False, rec=None):
        """src/lds.ml:741-1000 frame loop for len(seeds) independent chains in one launch."""
        _need_init()
        seeds = np.ascontiguousarray(seeds, np.uint64)
        n = len(seeds)
        rot = np.ascontiguousarray(start_rot9, np.float64).reshape(n, 9)
        pos = np.ascontiguousarray(start_pos3, np.float64).reshape(n, 3)
        P = McParams()
        P.roi_c = (C.c_double * 3)(*roi[:3]); P.roi_r = roi[3]; P.temperature_K = temperature_K
        P.n_steps = n_steps; P.tweak_rbonds = int(tweak_rbonds); P.hard_roi = int(hard_roi)
        P.no_flip = int(no_flip); P.intra_nb = int(intra_nb)
        res = (McResult * n)()
        xyz = np.empty((n, 3, lig.n)) if want_xyz else None
        trace = np.empty((n_steps, 4)) if want_trace else None
        _ck(lib().mmo_mc_run(rec.h if rec is not None else None, grid.h if grid is not None else None, lig.h, C.byref(P), C.c_int64(n), seeds.ctypes.data_as(C.POINTER(C.c_uint64)),
                             rot.ctypes.data_as(_dp), pos.ctypes.data_as(_dp), res,
                             xyz.ctypes.data_as(_dp) if want_xyz else C.cast(None, _dp),
                             trace.ctypes.data_as(_dp) if want_trace else C.cast(None, _dp)))
        out = []
        for r in res:
            out.append(dict(best_E=r.best_E, prev_E=r.prev_E, best_rot=np.array(r.best_rot), best_pos=np.array(r.best_pos),
                            max_rot=r.max_rot, max_trans=r.max_trans, n_accept_rigid=r.n_accept_rigid,
                            n_reject_rigid=r.n_reject_rigid, n_accept_conf=r.n_accept_conf, n_reject_conf=r.n_reject_conf,
                            n_ooroi=r.n_ooroi, n_ezero=r.n_ezero, too_long=r.too_long, frames_done=r.frames_done))
        return out, xyz, trace


def place_ligand_in_ROI(rec_mol, lig, roi, seed, n_starts, clash_check=True):
    """Lds.place_ligand_in_ROI (src/lds.ml:308-345): (rot9[n], pos3[n], trials); rec_mol = ALL receptor atoms"""
    (a, pa), (b, pb), (c, pc) = _d(rec_mol.xs), _d(rec_mol.ys), _d(rec_mol.zs)
    an, pan = _i(rec_mol.anum)
    rot = np.empty((n_starts, 9)); pos = np.empty((n_starts, 3))
    tr = C.c_int32()
    rc = (C.c_double * 3)(*roi[:3])
    _ck(lib().mmo_place_ligand_in_roi(C.c_int32(len(a)), pa, pb, pc, pan, lig.h, rc, C.c_double(roi[3]), C.c_uint64(seed),
                                      C.c_int32(n_starts), C.c_int32(1 if clash_check else 0), rot.ctypes.data_as(_dp),
                                      pos.ctypes.data_as(_dp), C.byref(tr)))
    return rot, pos, tr.value


class Optim:
    @staticmethod
    def apply_config(lig, config):
        """src/optim.ml:64-80 (place_ligand): x y z alpha beta gamma [rbond angles] -> coordinates, too_long"""
        cfg = np.ascontiguousarray(config, np.float64)
        x = np.empty(lig.n); y = np.empty(lig.n); z = np.empty(lig.n)
        tl = C.c_int32()
        _ck(lib().mmo_apply_config(lig.h, cfg.ctypes.data_as(_dp), C.c_int32(len(cfg)), x.ctypes.data_as(_dp),
                                   y.ctypes.data_as(_dp), z.ctypes.data_as(_dp), C.byref(tl)))
        return x, y, z, bool(tl.value)

    @staticmethod
    def rotated_copies(lig, center, rot9):
        """src/lig_rot_sample.ml:23-45: Mol.center_rotate_translate_copy mol rot center for every rotation"""
        rot = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        n = rot.shape[0]
        X = np.empty((n, lig.n)); Y = np.empty((n, lig.n)); Z = np.empty((n, lig.n))
        c = (C.c_double * 3)(*center)
        _ck(lib().mmo_rotated_copies(lig.h, c, C.c_int32(n), rot.ctypes.data_as(_dp), X.ctypes.data_as(_dp),
                                     Y.ctypes.data_as(_dp), Z.ctypes.data_as(_dp)))
        return X, Y, Z


class SO3:
    @staticmethod
    def rotations(n):
        out = np.empty((n, 9))
        _ck(lib().mmo_so3_rotations(C.c_int32(n), out.ctypes.data_as(_dp)))
        return out


class Rot:
    @staticmethod
    def r_xyz(a, b, g):
        out = np.empty(9)
        _ck(lib().mmo_rot_r_xyz(C.c_double(a), C.c_double(b), C.c_double(g), out.ctypes.data_as(_dp)))
        return out

    @staticmethod
    def decompose(r):
        r = np.ascontiguousarray(r, np.float64).reshape(9)
        out = np.empty(3)
        _ck(lib().mmo_rot_decompose(r.ctypes.data_as(_dp), out.ctypes.data_as(_dp)))
        return out


class MolFile:
    """N2: a mol2 / pqrs file parsed by the library's own C++ reader (mmo_molfile_*): mol2pqrs without the
    subprocess (src/mol2pqrs.ml, src/mol_graph.ml, src/pqrs.ml).  Host only; `ligand(k)` needs the GPU."""

    def __init__(self, path, kind="mol2"):
        self.h = C.c_void_p()
        p = os.fsencode(path)
        if kind == "mol2":
            _ck(lib().mmo_molfile_read_mol2(p, C.byref(self.h)))
        else:
            _ck(lib().mmo_molfile_read_pqrs(p, C.c_int(1 if kind == "receptor_pqrs" else 0), C.byref(self.h)))
        n, sk = C.c_int32(), C.c_int32()
        _ck(lib().mmo_molfile_count(self.h, C.byref(n), C.byref(sk)))
        self.n_mols, self.n_skipped = n.value, sk.value

    def mol(self, k):
        """molecule k as a pqrs.Mol (same layout as the Python reader's)"""
        from . import pqrs
        na, nrb, tot = C.c_int32(), C.c_int32(), C.c_int32()
        name = C.create_string_buffer(512)
        _ck(lib().mmo_molfile_shape(self.h, C.c_int32(k), C.byref(na), C.byref(nrb), C.byref(tot), name, C.c_int32(512)))
        n = na.value
        xs, ys, zs, q, r = (np.empty(n) for _ in range(5))
        anum, typ = np.empty(n, np.int32), np.empty(n, np.int32)
        dists = np.zeros(n * n, np.int32)
        left, right = np.empty(nrb.value, np.int32), np.empty(nrb.value, np.int32)
        off, idx = np.zeros(nrb.value + 1, np.int32), np.empty(max(1, tot.value), np.int32)
        _ck(lib().mmo_molfile_get(self.h, C.c_int32(k), xs.ctypes.data_as(_dp), ys.ctypes.data_as(_dp), zs.ctypes.data_as(_dp),
                                  q.ctypes.data_as(_dp), r.ctypes.data_as(_dp), anum.ctypes.data_as(_ip), typ.ctypes.data_as(_ip),
                                  dists.ctypes.data_as(_ip), left.ctypes.data_as(_ip), right.ctypes.data_as(_ip),
                                  off.ctypes.data_as(_ip), idx.ctypes.data_as(_ip)))
        groups = [idx[off[b]:off[b + 1]].copy() for b in range(nrb.value)]
        return pqrs.Mol(name.value.decode(), xs, ys, zs, q, r, anum, dists, left, right, groups, typ)

    def types(self):
        n = C.c_int32()
        _ck(lib().mmo_molfile_types(self.h, C.byref(n), None, None))
        ta, tq = np.empty(n.value, np.int32), np.empty(n.value)
        _ck(lib().mmo_molfile_types(self.h, C.byref(n), ta.ctypes.data_as(_ip), tq.ctypes.data_as(_dp)))
        return ta, tq

    def reduce_charges(self):
        """lds --less-charges (lds.ml:1887-1894): charges to two decimals, FF types re-assigned"""
        _ck(lib().mmo_molfile_reduce_charges(self.h))

    def write_pqrs(self, path):
        _ck(lib().mmo_molfile_write_pqrs(self.h, os.fsencode(path)))

    def write_mol2(self, path, k=0, xs=None, ys=None, zs=None, append=False):
        """Mol2.output_one of Mol.update_mol2 (mol2.ml:326-343, mol.ml:544-552): molecule k with the coordinates of
        every row of xs/ys/zs ([n_copies][n_atoms]); without coordinates the file's own, once"""
        if xs is None:
            n, px, py, pz = 1, None, None, None
        else:
            xs, ys, zs = (np.ascontiguousarray(np.atleast_2d(a), np.float64) for a in (xs, ys, zs))
            n, px, py, pz = xs.shape[0], xs.ctypes.data_as(_dp), ys.ctypes.data_as(_dp), zs.ctypes.data_as(_dp)
        _ck(lib().mmo_molfile_write_mol2(self.h, C.c_int32(k), C.c_int32(n), px, py, pz, os.fsencode(path),
                                         C.c_int(1 if append else 0)))

    def rotated_copies(self, k, rot9):
        """lig_rot_sample's body on the host (lig_rot_sample.ml:23-45): [n][n_atoms] coordinates"""
        rot = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
        n, L = rot.shape[0], self.mol(k).n
        X, Y, Z = (np.empty((n, L)) for _ in range(3))
        _ck(lib().mmo_molfile_rotated_copies(self.h, C.c_int32(k), C.c_int32(n), rot.ctypes.data_as(_dp),
                                             X.ctypes.data_as(_dp), Y.ctypes.data_as(_dp), Z.ctypes.data_as(_dp)))
        return X, Y, Z

    def apply_config(self, k, config):
        """place_ligand's body on the host (place_ligand.ml:36-59): Mol.center, then Optim.apply_config"""
        cfg = np.ascontiguousarray(config, np.float64)
        L = self.mol(k).n
        x, y, z = (np.empty(L) for _ in range(3))
        tl = C.c_int32()
        _ck(lib().mmo_molfile_apply_config(self.h, C.c_int32(k), cfg.ctypes.data_as(_dp), C.c_int32(len(cfg)),
                                           x.ctypes.data_as(_dp), y.ctypes.data_as(_dp), z.ctypes.data_as(_dp), C.byref(tl)))
        return x, y, z, bool(tl.value)

    def ligand(self, k, centered=True):
        _need_init()
        m = self.mol(k)
        o = Ligand.__new__(Ligand)
        o.h = C.c_void_p()
        _ck(lib().mmo_molfile_ligand(self.h, C.c_int32(k), C.c_int(1 if centered else 0), C.byref(o.h)))
        o.n = m.n
        if centered:
            c = [favg(m.xs), favg(m.ys), favg(m.zs)]
            o.xs, o.ys, o.zs = m.xs + (0.0 - c[0]), m.ys + (0.0 - c[1]), m.zs + (0.0 - c[2])
        else:
            o.xs, o.ys, o.zs = m.xs, m.ys, m.zs
        return o

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.mmo_molfile_destroy(self.h)
            self.h = None
