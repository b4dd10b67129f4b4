"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.  Never the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)
_lib = None


class ScanArgs(C.Structure):
    _fields_ = [("P", C.c_int), ("px", _dp), ("py", _dp), ("pz", _dp), ("pq", _dp), ("panum", _ip),
                ("L", C.c_int), ("lx", _dp), ("ly", _dp), ("lz", _dp), ("lq", _dp), ("lr", _dp),
                ("lanum", _ip), ("ltyp", _ip),
                ("scorer", C.c_int),
                ("g_step", C.c_double), ("g_dims", C.c_int * 3), ("maps", _fp),
                ("vdw_mask", _bp), ("m_step", C.c_double), ("m_dims", C.c_int * 3),
                ("roi_c", C.c_double * 3), ("roi_r", C.c_double),
                ("trans_step", C.c_double),
                ("n_rot", C.c_int), ("rot9", _dp),
                ("e_intra_const", C.c_double),
                ("topk", C.c_int),
                ("first_point", C.c_int64), ("n_points", C.c_int64)]


class ScanResult(C.Structure):
    _fields_ = [("n_scored", C.c_int64), ("n_candidates", C.c_int64), ("best_score", C.c_double),
                ("best_frame", C.c_int64), ("n_top", C.c_int), ("lattice_dims", C.c_int * 3)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        for name in ("orc_pow6", "orc_shift_12A", "orc_geo_mean", "orc_non_zero_dist", "orc_elec_weight",
                     "orc_beta", "orc_vdw_radius", "orc_ene_inter_global_brute", "orc_ene_inter_shifted_brute",
                     "orc_ene_intra_uffnb_brute", "orc_grid_node", "orc_trilin", "orc_ene_inter_interp",
                     "orc_favg", "orc_radius", "orc_rng_uniform"):
            if hasattr(L, name):
                getattr(L, name).restype = C.c_double
        _lib = L
    return _lib


def d(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(_dp)


def i32(a):
    a = np.ascontiguousarray(a, np.int32)
    return a, a.ctypes.data_as(_ip)


def _mol(m):
    return [d(m.xs), d(m.ys), d(m.zs), d(m.q), i32(m.anum)]


# ---- scalars ---------------------------------------------------------------------------------
def shift_12A(x):
    return lib().orc_shift_12A(C.c_double(x))


def vdw_xidi(a1, a2):
    out = (C.c_double * 2)()
    lib().orc_vdw_xidi(C.c_int(a1), C.c_int(a2), out)
    return out[0], out[1]


def pair_energy(anum1, q1, anum2, q2, r, shifted):
    """energy of a two-atom system through the full brute-force routines"""
    fn = lib().orc_ene_inter_shifted_brute if shifted else lib().orc_ene_inter_global_brute
    z = np.zeros(1)
    return fn(C.c_int(1), d(z)[1], d(z)[1], d(z)[1], d([q1])[1], i32([anum1])[1],
              C.c_int(1), d([r])[1], d(z)[1], d(z)[1], d([q2])[1], i32([anum2])[1])


# ---- energies -----------------------------------------------------------------------------------
def ene_inter(rec, lig_q, lig_anum, xs, ys, zs, shifted=True):
    """xs/ys/zs: [n_poses, L]"""
    fn = lib().orc_ene_inter_shifted_brute if shifted else lib().orc_ene_inter_global_brute
    xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
    zs = np.atleast_2d(np.asarray(zs, np.float64))
    r = _mol(rec)
    (lq, plq), (la, pla) = d(lig_q), i32(lig_anum)
    out = np.empty(xs.shape[0])
    L = xs.shape[1]
    for p in range(xs.shape[0]):
        out[p] = fn(C.c_int(rec.n), r[0][1], r[1][1], r[2][1], r[3][1], r[4][1], C.c_int(L),
                    d(xs[p])[1], d(ys[p])[1], d(zs[p])[1], plq, pla)
    return out


def ene_inter_components(rec, lig_q, lig_anum, xs, ys, zs):
    xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
    zs = np.atleast_2d(np.asarray(zs, np.float64))
    r = _mol(rec)
    (lq, plq), (la, pla) = d(lig_q), i32(lig_anum)
    e = np.empty(xs.shape[0]); v = np.empty(xs.shape[0])
    out = (C.c_double * 2)()
    for p in range(xs.shape[0]):
        lib().orc_ene_inter_shifted_components(C.c_int(rec.n), r[0][1], r[1][1], r[2][1], r[3][1], r[4][1],
                                               C.c_int(xs.shape[1]), d(xs[p])[1], d(ys[p])[1], d(zs[p])[1],
                                               plq, pla, out)
        e[p], v[p] = out[0], out[1]
    return e, v


def ene_intra(lig, xs, ys, zs):
    xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
    zs = np.atleast_2d(np.asarray(zs, np.float64))
    (lq, plq), (la, pla), (ld, pld) = d(lig.q), i32(lig.anum), i32(lig.dists)
    out = np.empty(xs.shape[0])
    for p in range(xs.shape[0]):
        out[p] = lib().orc_ene_intra_uffnb_brute(C.c_int(lig.n), d(xs[p])[1], d(ys[p])[1], d(zs[p])[1], plq, pla, pld)
    return out


def score_poses_mt(rec, cx, cy, cz, lig_q, lig_anum, rot9, trans3, nthreads=0):
    r = _mol(rec)
    rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
    trans3 = np.ascontiguousarray(trans3, np.float64).reshape(-1, 3)
    n = rot9.shape[0]
    out = np.empty(n)
    lib().orc_score_poses_shifted_mt(C.c_int(rec.n), r[0][1], r[1][1], r[2][1], r[3][1], r[4][1],
                                     C.c_int(len(cx)), d(cx)[1], d(cy)[1], d(cz)[1], d(lig_q)[1], i32(lig_anum)[1],
                                     C.c_int64(n), rot9.ctypes.data_as(_dp), trans3.ctypes.data_as(_dp),
                                     out.ctypes.data_as(_dp), C.c_int(nthreads))
    return out


def num_threads():
    return int(lib().orc_num_threads())


# ---- poses ----------------------------------------------------------------------------------------
def rotate_then_translate(cx, cy, cz, rot9, t3):
    L = len(cx)
    ox = np.empty(L); oy = np.empty(L); oz = np.empty(L)
    lib().orc_rotate_then_translate(C.c_int(L), d(cx)[1], d(cy)[1], d(cz)[1], d(rot9)[1], d(t3)[1],
                                    ox.ctypes.data_as(_dp), oy.ctypes.data_as(_dp), oz.ctypes.data_as(_dp))
    return ox, oy, oz


def pose_coords(cx, cy, cz, rot9, trans3):
    rot9 = np.asarray(rot9, np.float64).reshape(-1, 9)
    trans3 = np.asarray(trans3, np.float64).reshape(-1, 3)
    n, L = rot9.shape[0], len(cx)
    X = np.empty((n, L)); Y = np.empty((n, L)); Z = np.empty((n, L))
    for p in range(n):
        X[p], Y[p], Z[p] = rotate_then_translate(cx, cy, cz, rot9[p], trans3[p])
    return X, Y, Z


def so3_rotations(n):
    out = np.empty((n, 9))
    lib().orc_so3_rotations(C.c_int(n), out.ctypes.data_as(_dp))
    return out


def so3_quat(n, i):
    q = (C.c_double * 4)()
    lib().orc_so3_quat(C.c_int(n), C.c_int(i), q)
    return tuple(q)


def rot_r_xyz(a, b, g):
    out = np.empty(9)
    lib().orc_rot_r_xyz(C.c_double(a), C.c_double(b), C.c_double(g), out.ctypes.data_as(_dp))
    return out


def rot_decompose(r):
    out = np.empty(3)
    lib().orc_rot_decompose(d(r)[1], out.ctypes.data_as(_dp))
    return out


def rot_axis(axis, theta):
    out = np.empty(9)
    fn = {0: lib().orc_rot_rx, 1: lib().orc_rot_ry, 2: lib().orc_rot_rz}[axis]
    fn(C.c_double(theta), out.ctypes.data_as(_dp))
    return out


def rot_mult(a, b):
    out = np.empty(9)
    lib().orc_rot_mult(d(a)[1], d(b)[1], out.ctypes.data_as(_dp))
    return out


def favg(a):
    return lib().orc_favg(C.c_int(len(a)), d(a)[1])


# ---- grids --------------------------------------------------------------------------------------
def grid_from_box(step, bx, by, bz):
    dims = (C.c_int * 3)()
    lib().orc_grid_from_box(C.c_double(step), C.c_double(bx), C.c_double(by), C.c_double(bz), dims)
    return tuple(dims)


def grid_node(step, dim, i):
    return lib().orc_grid_node(C.c_double(step), C.c_int(dim), C.c_int(i))


def grid_build(rec, step, dims, type_anum, type_q, mask=None):
    T = len(type_anum)
    nvox = dims[0] * dims[1] * dims[2]
    maps = np.zeros((T, nvox), np.float32)
    r = _mol(rec)
    pm = C.cast(None, _bp)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
        pm = mask.ctypes.data_as(_bp)
    lib().orc_grid_build(C.c_int(rec.n), r[0][1], r[1][1], r[2][1], r[3][1], r[4][1], C.c_double(step),
                         (C.c_int * 3)(*dims), pm, C.c_int(T), i32(type_anum)[1], d(type_q)[1],
                         maps.ctypes.data_as(_fp))
    return maps


def trilin(step, dims, arr, x, y, z):
    arr = np.ascontiguousarray(arr, np.float32)
    return lib().orc_trilin(C.c_double(step), (C.c_int * 3)(*dims), arr.ctypes.data_as(_fp),
                            C.c_double(x), C.c_double(y), C.c_double(z))


def ene_inter_interp(step, dims, maps, ltyp, xs, ys, zs):
    maps = np.ascontiguousarray(maps, np.float32)
    xs = np.atleast_2d(np.asarray(xs, np.float64)); ys = np.atleast_2d(np.asarray(ys, np.float64))
    zs = np.atleast_2d(np.asarray(zs, np.float64))
    out = np.empty(xs.shape[0])
    (lt, plt) = i32(ltyp)
    for p in range(xs.shape[0]):
        out[p] = lib().orc_ene_inter_interp(C.c_double(step), (C.c_int * 3)(*dims), maps.ctypes.data_as(_fp),
                                            C.c_int(xs.shape[1]), d(xs[p])[1], d(ys[p])[1], d(zs[p])[1], plt)
    return out


def bitmask_sphere(step, dims, c, r):
    nvox = dims[0] * dims[1] * dims[2]
    mask = np.zeros((nvox + 7) // 8, np.uint8)
    lib().orc_bitmask_sphere(C.c_double(step), (C.c_int * 3)(*dims), C.c_double(c[0]), C.c_double(c[1]),
                             C.c_double(c[2]), C.c_double(r), mask.ctypes.data_as(_bp))
    return mask


def first_solvent_shell(xs, ys, zs, radii, step, dims):
    nvox = dims[0] * dims[1] * dims[2]
    mask = np.zeros((nvox + 7) // 8 + 8, np.uint8)
    lib().orc_first_solvent_shell(C.c_int(len(xs)), d(xs)[1], d(ys)[1], d(zs)[1], d(radii)[1], C.c_double(step),
                                  (C.c_int * 3)(*dims), mask.ctypes.data_as(_bp))
    return mask[:(nvox + 7) // 8]


def bitmask_whole_protein(xs, ys, zs, step, dims):
    nvox = dims[0] * dims[1] * dims[2]
    mask = np.zeros((nvox + 7) // 8 + 8, np.uint8)
    lib().orc_bitmask_whole_protein(C.c_int(len(xs)), d(xs)[1], d(ys)[1], d(zs)[1], C.c_double(step),
                                    (C.c_int * 3)(*dims), mask.ctypes.data_as(_bp))
    return mask[:(nvox + 7) // 8]


def protein_desolv(rec, step, dims, shell, roi):
    """lds.ml:204-236: one double per voxel"""
    nvox = dims[0] * dims[1] * dims[2]
    res = np.empty(nvox)
    shell = np.ascontiguousarray(shell, np.uint8)
    lib().orc_protein_desolv(C.c_int(rec.n), d(rec.xs)[1], d(rec.ys)[1], d(rec.zs)[1], d(rec.q)[1], C.c_double(step),
                             (C.c_int * 3)(*dims), shell.ctypes.data_as(_bp), (C.c_double * 4)(*roi), res.ctypes.data_as(_dp))
    return res


def desolvation_penalty(step, dims, prot_shell, contribs, xs, ys, zs, q, radii):
    """lds.ml:239-267: (prot, lig) of one ligand pose"""
    prot_shell = np.ascontiguousarray(prot_shell, np.uint8)
    contribs = np.ascontiguousarray(contribs, np.float64)
    op, ol = C.c_double(), C.c_double()
    lib().orc_desolvation_penalty(C.c_double(step), (C.c_int * 3)(*dims), prot_shell.ctypes.data_as(_bp),
                                  contribs.ctypes.data_as(_dp), C.c_int(len(xs)), d(xs)[1], d(ys)[1], d(zs)[1], d(q)[1],
                                  d(radii)[1], C.byref(op), C.byref(ol))
    return op.value, ol.value


def vdw_volume(xs, ys, zs, radii, step, dims):
    nvox = dims[0] * dims[1] * dims[2]
    mask = np.zeros((nvox + 7) // 8 + 8, np.uint8)
    lib().orc_vdw_volume(C.c_int(len(xs)), d(xs)[1], d(ys)[1], d(zs)[1], d(radii)[1], C.c_double(step),
                         (C.c_int * 3)(*dims), mask.ctypes.data_as(_bp))
    return mask


def protein_ligand_clash(step, dims, mask, xs, ys, zs):
    mask = np.ascontiguousarray(mask, np.uint8)
    return bool(lib().orc_protein_ligand_clash(C.c_double(step), (C.c_int * 3)(*dims), mask.ctypes.data_as(_bp),
                                               C.c_int(len(xs)), d(xs)[1], d(ys)[1], d(zs)[1]))


# ---- scan ------------------------------------------------------------------------------------------
def scan(rec, lig, cx, cy, cz, roi, trans_step, rot9, topk, scorer=0, e_intra_const=0.0, vdw_mask=None,
         m_step=0.5, m_dims=(0, 0, 0), maps=None, g_step=0.0, g_dims=(0, 0, 0), first_point=0, n_points=-1,
         score_frames=None):
    keep = []

    def K(x):
        keep.append(x)
        return x[1]
    A = ScanArgs()
    A.P = rec.n if rec is not None else 0
    if rec is not None:
        A.px, A.py, A.pz, A.pq, A.panum = K(d(rec.xs)), K(d(rec.ys)), K(d(rec.zs)), K(d(rec.q)), K(i32(rec.anum))
    A.L = len(cx)
    A.lx, A.ly, A.lz, A.lq, A.lr = K(d(cx)), K(d(cy)), K(d(cz)), K(d(lig.q)), K(d(lig.r))
    A.lanum = K(i32(lig.anum))
    A.ltyp = K(i32(lig.typ if lig.typ is not None else np.zeros(lig.n, np.int32)))
    A.scorer = scorer
    if maps is not None:
        maps = np.ascontiguousarray(maps, np.float32)
        keep.append(maps)
        A.maps = maps.ctypes.data_as(_fp)
        A.g_step = g_step
        A.g_dims = (C.c_int * 3)(*g_dims)
    if vdw_mask is not None:
        vdw_mask = np.ascontiguousarray(vdw_mask, np.uint8)
        keep.append(vdw_mask)
        A.vdw_mask = vdw_mask.ctypes.data_as(_bp)
        A.m_step = m_step
        A.m_dims = (C.c_int * 3)(*m_dims)
    A.roi_c = (C.c_double * 3)(*roi[:3])
    A.roi_r = roi[3]
    A.trans_step = trans_step
    rot9 = np.ascontiguousarray(rot9, np.float64).reshape(-1, 9)
    A.n_rot = rot9.shape[0]
    A.rot9 = rot9.ctypes.data_as(_dp)
    A.e_intra_const = e_intra_const
    A.topk = topk
    A.first_point, A.n_points = first_point, n_points
    if score_frames is not None:
        fr = np.ascontiguousarray(score_frames, np.int64)
        out = np.empty(len(fr))
        lib().orc_scan_score_frames(C.byref(A), C.c_int(len(fr)), fr.ctypes.data_as(_lp), out.ctypes.data_as(_dp))
        return out
    k = max(topk, 1)
    ts = np.empty(k); tf = np.empty(k, np.int64)
    R = ScanResult()
    lib().orc_scan(C.byref(A), ts.ctypes.data_as(_dp), tf.ctypes.data_as(_lp), C.byref(R))
    return dict(top_scores=ts[:R.n_top].copy(), top_frames=tf[:R.n_top].copy(), best_score=R.best_score,
                best_frame=R.best_frame, n_scored=R.n_scored, n_candidates=R.n_candidates,
                lattice_dims=tuple(R.lattice_dims))


# ---- Monte-Carlo chain ---------------------------------------------------------------------------------
class McArgs(C.Structure):
    _fields_ = [("L", C.c_int), ("lx", _dp), ("ly", _dp), ("lz", _dp), ("lq", _dp), ("lanum", _ip), ("ltyp", _ip),
                ("dists", _ip),
                ("n_rbonds", C.c_int), ("rb_left", _ip), ("rb_right", _ip), ("rg_off", _ip), ("rg_idx", _ip),
                ("scorer", C.c_int),
                ("P", C.c_int), ("px", _dp), ("py", _dp), ("pz", _dp), ("pq", _dp), ("panum", _ip),
                ("g_step", C.c_double), ("g_dims", C.c_int * 3), ("maps", _fp),
                ("roi_c", C.c_double * 3), ("roi_r", C.c_double),
                ("tweak_rbonds", C.c_int), ("hard_roi", C.c_int), ("no_flip", C.c_int), ("intra_nb", C.c_int),
                ("beta", C.c_double), ("n_steps", C.c_int), ("seed", C.c_uint64),
                ("rot0", C.c_double * 9), ("pos0", C.c_double * 3)]


class McResult(C.Structure):
    _fields_ = [("best_E", C.c_double), ("prev_E", C.c_double), ("best_rot", C.c_double * 9), ("best_pos", C.c_double * 3),
                ("n_accept_rigid", C.c_int64), ("n_reject_rigid", C.c_int64), ("n_accept_conf", C.c_int64),
                ("n_reject_conf", C.c_int64), ("n_ooroi", C.c_int64), ("n_ezero", C.c_int64),
                ("too_long", C.c_int), ("frames_done", C.c_int), ("max_rot", C.c_double), ("max_trans", C.c_double)]


def mc_run(lig, cx, cy, cz, roi, n_steps, seed, rot0, pos0, maps=None, g_step=0.0, g_dims=(0, 0, 0), rec=None,
           tweak_rbonds=True, hard_roi=True, no_flip=False, intra_nb=True, temperature_K=293.15):
    keep = []

    def K(x):
        keep.append(x)
        return x[1]
    A = McArgs()
    A.L = lig.n
    A.lx, A.ly, A.lz, A.lq = K(d(cx)), K(d(cy)), K(d(cz)), K(d(lig.q))
    A.lanum, A.ltyp, A.dists = K(i32(lig.anum)), K(i32(lig.typ)), K(i32(lig.dists))
    off, idx = lig.rgroup_csr()
    A.n_rbonds = lig.n_rbonds
    A.rb_left, A.rb_right, A.rg_off, A.rg_idx = K(i32(lig.rb_left)), K(i32(lig.rb_right)), K(i32(off)), K(i32(idx))
    if maps is not None:
        A.scorer = 2
        maps = np.ascontiguousarray(maps, np.float32); keep.append(maps)
        A.maps = maps.ctypes.data_as(_fp); A.g_step = g_step; A.g_dims = (C.c_int * 3)(*g_dims)
    else:
        A.scorer = 0
        A.P = rec.n
        A.px, A.py, A.pz, A.pq, A.panum = K(d(rec.xs)), K(d(rec.ys)), K(d(rec.zs)), K(d(rec.q)), K(i32(rec.anum))
    A.roi_c = (C.c_double * 3)(*roi[:3]); A.roi_r = roi[3]
    A.tweak_rbonds, A.hard_roi, A.no_flip, A.intra_nb = int(tweak_rbonds), int(hard_roi), int(no_flip), int(intra_nb)
    A.beta = lib().orc_beta(C.c_double(temperature_K))
    A.n_steps = n_steps; A.seed = int(seed)
    A.rot0 = (C.c_double * 9)(*np.asarray(rot0, np.float64).reshape(9)); A.pos0 = (C.c_double * 3)(*pos0)
    R = McResult()
    xyz = np.empty((3, lig.n)); trace = np.empty((n_steps, 4))
    lib().orc_mc_run(C.byref(A), C.byref(R), xyz.ctypes.data_as(_dp), trace.ctypes.data_as(_dp))
    res = {f: getattr(R, f) for f, _ in McResult._fields_ if f not in ("best_rot", "best_pos")}
    res["best_rot"] = np.array(R.best_rot); res["best_pos"] = np.array(R.best_pos)
    return res, xyz, trace[:R.frames_done]


def apply_config(lig, cx, cy, cz, config):
    off, idx = lig.rgroup_csr()
    L = lig.n
    ox = np.empty(L); oy = np.empty(L); oz = np.empty(L)
    cfg = np.ascontiguousarray(config, np.float64)
    tl = lib().orc_apply_config(C.c_int(L), d(cx)[1], d(cy)[1], d(cz)[1], C.c_int(lig.n_rbonds), i32(lig.rb_left)[1],
                                i32(lig.rb_right)[1], i32(off)[1], i32(idx)[1], cfg.ctypes.data_as(_dp), C.c_int(len(cfg)),
                                ox.ctypes.data_as(_dp), oy.ctypes.data_as(_dp), oz.ctypes.data_as(_dp))
    return ox, oy, oz, bool(tl)
