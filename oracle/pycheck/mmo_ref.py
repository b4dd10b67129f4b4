"""mmo_ref.py -- a SECOND, independent restatement of the reference's scoring path, in pure Python.

TEST INFRASTRUCTURE ONLY: imported by tests/ (never by the package, bench.py's product arm or any kernel).
PARITY UNPINNED, like oracle/mmo_oracle.c: the OCaml reference cannot be built here and ships no golden
energies (F4).  What this file adds: it was written from the OCaml text alone (src/FF.ml, UFF.ml, math.ml, V3.ml,
mol.ml, G3D.ml, grid.ml, rot.ml, move.ml, SW.ml, lds.ml -- file:line on every function), in another language, with the
reference's own data structures (records, refs, molecule copies) instead of the flat arrays of the C oracle and the
kernels.  tests/test_second_checker.py requires the two restatements to agree BIT FOR BIT on whole-pose energies, map
voxels, trilinear look-ups and 200 Monte-Carlo frames, which removes transcription slips of either one; a misreading
of the OCaml text shared by both would survive (the readings that remain single-source are listed in DESIGN.md).

Python floats are IEEE binary64 and +, -, *, /, math.sqrt round to nearest, as OCaml's unboxed floats do; Python never
contracts a*b+c.  The random stream and sin/cos/exp of the Monte-Carlo loop are those of include/mmo_detmath.h (the
repository's replacement for OCaml's Random.State and libm, see there), restated at the end of this file.
"""
import math
import struct

# ---------------------------------------------------------------------------------------------- FF.ml, UFF.ml


def square(x):                      # FF.ml:5-6
    return x * x


def pow3(x):                        # FF.ml:8-9   x *. x *. x  = (x *. x) *. x
    return x * x * x


def pow6(x):                        # FF.ml:11-12
    return pow3(x * x)


def geo_mean(x, y):                 # FF.ml:14-15
    return math.sqrt(x * y)


def shift_12A(d):                   # FF.ml:17-20
    if d < 12.0:
        return square(1.0 - (square(d / 12.0)))
    return 0.0


ANUM_XI_DI = [(0, (0.0, 0.0)), (1, (2.886, 0.044)), (6, (3.851, 0.105)), (7, (3.660, 0.069)), (8, (3.500, 0.060)),
              (9, (3.364, 0.050)), (12, (3.021, 0.111)), (15, (4.147, 0.305)), (16, (4.035, 0.274)),
              (17, (3.947, 0.227)), (35, (4.189, 0.251)), (53, (4.500, 0.339))]        # UFF.ml:10-22
EPSILON_PROT = 4.0                  # const.ml:18
ELEC_WEIGHT_PROTEIN = 332.0637 / EPSILON_PROT      # UFF.ml:25
ANUMS = 119                         # UFF.ml:32
XIDI = {}                           # UFF.ml:37-51; missing keys are the {nan; nan} of UFF.ml:37
for _a1, (_x1, _d1) in ANUM_XI_DI:
    for _a2, (_x2, _d2) in ANUM_XI_DI:
        XIDI[_a1 * ANUMS + _a2] = (geo_mean(_x1, _x2), geo_mean(_d1, _d2))


def vdW_xiDi(anu1, anu2):           # UFF.ml:39-40 -> (x_ij, d_ij)
    return XIDI.get(anu1 * ANUMS + anu2, (math.nan, math.nan))


def non_zero_dist(x):               # math.ml:58-62
    return 0.01 if x < 0.01 else x


def dist2(u, v):                    # V3.ml:23-28
    dx = u[0] - v[0]
    dy = u[1] - v[1]
    dz = u[2] - v[2]
    return dx * dx + dy * dy + dz * dz


def dist(u, v):                     # V3.ml:31-33
    return math.sqrt(dist2(u, v))


# ---------------------------------------------------------------------------------------------- Mol.t (mol.ml:17-35)
class SW:
    """SW.ml: sliding window of accept/reject events"""

    def __init__(self, n):          # SW.ml:14-18
        self.size, self.accepts, self.rejects, self.events = n, 0, 0, []

    def process(self, evt):         # SW.ml:21-34
        self.events.append(evt)
        if evt:
            self.accepts += 1
        else:
            self.rejects += 1
        if len(self.events) > self.size:
            too_old = self.events.pop(0)
            if too_old:
                self.accepts -= 1
            else:
                self.rejects -= 1

    def get_ratio(self):            # SW.ml:36-37 (0/0 = nan, as float division in OCaml)
        den = self.accepts + self.rejects
        return float(self.accepts) / float(den) if den else math.nan

    def reset(self):                # SW.ml:40-43
        self.accepts, self.rejects, self.events = 0, 0, []


class Mol:
    def __init__(self, xs, ys, zs, q_a, elt_a, t_a=None, dists_a=None, rbonds=(), rgroups=(), block_size=100,
                 max_rbond_rot=None, center=None):
        n = len(xs)
        self.xs, self.ys, self.zs = [float(v) for v in xs], [float(v) for v in ys], [float(v) for v in zs]
        self.q_a, self.elt_a = [float(v) for v in q_a], [int(v) for v in elt_a]
        self.t_a = [int(v) for v in t_a] if t_a is not None else [-1] * n
        self.n = n
        # interacting_a.(i).(j) = dists_a.(i + j*n) >= 3 (mol.ml:151-152, 203-208)
        self.interacting = None
        if dists_a is not None:
            self.interacting = [[int(dists_a[i + j * n]) >= 3 for j in range(n)] for i in range(n)]
        self.rbonds = [(int(l), int(r)) for l, r in rbonds]                    # {left; right}
        self.rgroups = [[int(v) for v in g] for g in rgroups]
        self.rbonds_dr = [max_rbond_rot] * len(self.rbonds)                    # Params.max_rbond_rot each
        self.rbonds_sw = [SW(block_size) for _ in self.rbonds]
        # a ligand that went through Mol.center (lds.ml:44-52) has center = V3.origin exactly, not a recomputed mean
        self.center = (0.0, 0.0, 0.0)
        if center is None:
            self.update_center()
        else:
            self.center = tuple(float(v) for v in center)

    def get_xyz(self, i):           # mol.ml:113-117
        return (self.xs[i], self.ys[i], self.zs[i])

    def set_xyz(self, i, p):
        self.xs[i], self.ys[i], self.zs[i] = p

    def copy(self):                 # mol.ml:57-62: coordinates and step sizes fresh, the window records SHARED
        m = Mol.__new__(Mol)
        m.__dict__.update(self.__dict__)
        m.xs, m.ys, m.zs = list(self.xs), list(self.ys), list(self.zs)
        m.rbonds_dr = list(self.rbonds_dr)
        m.rbonds_sw = list(self.rbonds_sw)          # A.copy of an array of records: same SW objects
        return m

    def update_center(self):        # mol.ml:353-356, A.favg (Batteries: Kahan-summed fsum / n -- unpinned library)
        self.center = (favg(self.xs), favg(self.ys), favg(self.zs))

    def radius(self):               # mol.ml:576-583
        maxi = 0.0
        for i in range(self.n):
            maxi = max(maxi, 0.01 + dist(self.center, self.get_xyz(i)))
        return maxi

    def translate_by(self, v):      # mol.ml:593-600
        for i in range(self.n):
            self.xs[i] = self.xs[i] + v[0]
            self.ys[i] = self.ys[i] + v[1]
            self.zs[i] = self.zs[i] + v[2]
        self.center = (self.center[0] + v[0], self.center[1] + v[1], self.center[2] + v[2])

    def centered_rotate(self, r):   # mol.ml:603-607
        for i in range(self.n):
            self.set_xyz(i, rot_rotate(r, self.get_xyz(i)))

    def rotate_bond(self, i, alpha, cos_sin):       # mol.ml:610-631
        left, right = self.rbonds[i]
        center = self.get_xyz(right)
        d = v_diff(center, self.get_xyz(left))
        axis = v_normalize(d)
        rot = rot_of_axis_angle(axis, alpha, cos_sin)
        for k in self.rgroups[i]:
            self.set_xyz(k, v_add(rot_rotate(rot, v_diff(self.get_xyz(k), center)), center))
        self.update_center()

    def center_(self):              # mol.ml:697-701 (Mol.center)
        mean = self.center
        self.translate_by((-mean[0], -mean[1], -mean[2]))
        self.center = (0.0, 0.0, 0.0)


def favg(a):
    """Batteries' A.favg = fsum a /. float n, fsum being Kahan's compensated sum (library not vendored: the same
    reading as the C oracle's; for already-centred inputs it is what defines conf'.center)"""
    s = 0.0
    c = 0.0
    for v in a:
        y = v - c
        t = s + y
        c = (t - s) - y
        s = t
    return s / float(len(a))


def v_add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def v_diff(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def v_normalize(v):                 # Vector3.normalize (library): v / |v|, |v| = sqrt (x*x + y*y + z*z)
    m = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return (v[0] / m, v[1] / m, v[2] / m)


# ---------------------------------------------------------------------------------------------- rot.ml
ROT_ID = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)      # rot.ml:14-17


def rot_rx(theta, cos_sin):         # rot.ml:22-28
    c, s = cos_sin(theta)
    return (1.0, 0.0, 0.0, 0.0, c, s, 0.0, -s, c)


def rot_ry(theta, cos_sin):         # rot.ml:31-37
    c, s = cos_sin(theta)
    return (c, 0.0, -s, 0.0, 1.0, 0.0, s, 0.0, c)


def rot_rz(theta, cos_sin):         # rot.ml:40-46
    c, s = cos_sin(theta)
    return (c, s, 0.0, -s, c, 0.0, 0.0, 0.0, 1.0)


def rot_mult(r1, r2):               # rot.ml:77-94
    a1, b1, c1, d1, e1, f1, g1, h1, i1 = r1
    a2, b2, c2, d2, e2, f2, g2, h2, i2 = r2
    return (a1 * a2 + b1 * d2 + c1 * g2, a1 * b2 + b1 * e2 + c1 * h2, a1 * c2 + b1 * f2 + c1 * i2,
            d1 * a2 + e1 * d2 + f1 * g2, d1 * b2 + e1 * e2 + f1 * h2, d1 * c2 + e1 * f2 + f1 * i2,
            g1 * a2 + h1 * d2 + i1 * g2, g1 * b2 + h1 * e2 + i1 * h2, g1 * c2 + h1 * f2 + i1 * i2)


def rot_rotate(r, v):               # rot.ml:97-100
    return (r[0] * v[0] + r[1] * v[1] + r[2] * v[2],
            r[3] * v[0] + r[4] * v[1] + r[5] * v[2],
            r[6] * v[0] + r[7] * v[1] + r[8] * v[2])


def rot_of_axis_angle(axis, theta, cos_sin):       # rot.ml:136-146
    x, y, z = axis
    c, s = cos_sin(theta)
    omc = 1.0 - c
    return (c + x * x * omc, x * y * omc - z * s, x * z * omc + y * s,
            x * y * omc + z * s, c + y * y * omc, y * z * omc - x * s,
            x * z * omc - y * s, y * z * omc + x * s, c + z * z * omc)


# ---------------------------------------------------------------------------------------------- mol.ml energies
def ene_inter_UFF_global_brute(prot, lig):          # mol.ml:796-818
    sum_elec = 0.0
    sum_vdW = 0.0
    for i in range(prot.n):
        q_i, r_i, prot_anum = prot.q_a[i], prot.get_xyz(i), prot.elt_a[i]
        for j in range(lig.n):
            q_j, r_j, lig_anum = lig.q_a[j], lig.get_xyz(j), lig.elt_a[j]
            r_ij = non_zero_dist(dist(r_i, r_j))
            x_ij, d_ij = vdW_xiDi(prot_anum, lig_anum)
            p6 = pow6(x_ij / r_ij)
            sum_elec = sum_elec + ((q_i * q_j) / r_ij)
            sum_vdW = sum_vdW + (d_ij * ((-2.0 * p6) + (p6 * p6)))
    return (ELEC_WEIGHT_PROTEIN * sum_elec) + sum_vdW


def ene_inter_UFF_shifted_brute(prot, lig):         # mol.ml:822-849
    sum_elec = 0.0
    sum_vdW = 0.0
    for i in range(prot.n):
        q_i, r_i, prot_anum = prot.q_a[i], prot.get_xyz(i), prot.elt_a[i]
        for j in range(lig.n):
            r_ij2 = dist2(r_i, lig.get_xyz(j))
            if r_ij2 < 144.0:
                q_j, lig_anum = lig.q_a[j], lig.elt_a[j]
                r_ij = non_zero_dist(math.sqrt(r_ij2))
                w = shift_12A(r_ij)
                x_ij, d_ij = vdW_xiDi(prot_anum, lig_anum)
                p6 = pow6(x_ij / r_ij)
                sum_elec = sum_elec + w * ((q_i * q_j) / r_ij)
                sum_vdW = sum_vdW + w * (d_ij * ((-2.0 * p6) + (p6 * p6)))
    return (ELEC_WEIGHT_PROTEIN * sum_elec) + sum_vdW


def ene_intra_UFFNB_brute(lig):                     # mol.ml:881-903
    n = lig.n
    sum_elec = 0.0
    sum_vdW = 0.0
    for i in range(n - 1):
        q_i, r_i, anum_i = lig.q_a[i], lig.get_xyz(i), lig.elt_a[i]
        interacting = lig.interacting[i]
        for j in range(i + 1, n):
            if interacting[j]:
                r_ij = non_zero_dist(dist(r_i, lig.get_xyz(j)))
                x_ij, d_ij = vdW_xiDi(anum_i, lig.elt_a[j])
                p6 = pow6(x_ij / r_ij)
                sum_elec = sum_elec + (q_i * lig.q_a[j]) / r_ij
                sum_vdW = sum_vdW + d_ij * ((-2.0 * p6) + (p6 * p6))
    return (ELEC_WEIGHT_PROTEIN * sum_elec) + sum_vdW


def ene_inter_UFF_shifted_grid(prot, r_j, probes):  # mol.ml:964-989; probes = [(anum, q)], BST.neighbors in index order
    n = len(probes)
    sum_elec = [0.0] * n
    sum_vdW = [0.0] * n
    for i in range(prot.n):
        r_i = prot.get_xyz(i)
        if not (dist(r_i, r_j) <= 12.0):            # BST.neighbors query 12.0 (w(12) = 0: boundary convention irrelevant)
            continue
        q_i, prot_anum = prot.q_a[i], prot.elt_a[i]
        r_ij = non_zero_dist(dist(r_i, r_j))
        w = shift_12A(r_ij)
        for l, (anum, q) in enumerate(probes):
            x_ij, d_ij = vdW_xiDi(anum, prot_anum)
            p6 = pow6(x_ij / r_ij)
            sum_elec[l] = sum_elec[l] + w * ((q_i * q) / r_ij)
            sum_vdW[l] = sum_vdW[l] + w * (d_ij * ((-2.0 * p6) + (p6 * p6)))
    return [ELEC_WEIGHT_PROTEIN * e + v for e, v in zip(sum_elec, sum_vdW)]


def f32(x):
    """Bigarray float32 store (BA1.ml:7): round to nearest single, read back as double"""
    return struct.unpack("f", struct.pack("f", x))[0]


def map_value(e):                   # lds.ml:463-467: G3D.init_idx maps.(i) idx (min Params.max_E e)
    max_E = 100000.0                # params.ml:26
    return f32(max_E if max_E <= e else e)         # OCaml min: if a <= b then a else b


class Grid:                         # grid.ml:11-52
    def __init__(self, step, bx, by, bz):
        self.step = step
        self.one_div_step = 1.0 / step
        n = [int(math.ceil(l / step)) for l in (bx, by, bz)]           # num_steps, grid.ml:37-38
        self.x_dim, self.y_dim, self.z_dim = n[0] + 1, n[1] + 1, n[2] + 1
        self.xy_dim = self.x_dim * self.y_dim
        # L.frange 0.0 `To (step *. float n') (n' + 1)  (Batteries, not vendored): start + i * ((stop - start) / (n - 1))
        self.xs = frange(0.0, step * float(n[0]), self.x_dim)
        self.ys = frange(0.0, step * float(n[1]), self.y_dim)
        self.zs = frange(0.0, step * float(n[2]), self.z_dim)


def frange(start, stop, n):
    if n == 1:
        return [start]
    inc = (stop - start) / float(n - 1)
    return [start + float(i) * inc for i in range(n)]


def trilin(grid, arr, p):           # G3D.ml:97-157 (arr: sequence of float32 values, index i + j*x_dim + k*xy_dim)
    i0 = int(p[0] * grid.one_div_step)
    j0 = int(p[1] * grid.one_div_step)
    k0 = int(p[2] * grid.one_div_step)
    i1, j1, k1 = i0 + 1, j0 + 1, k0 + 1
    j0x, j1x = j0 * grid.x_dim, j1 * grid.x_dim
    k0xy, k1xy = k0 * grid.xy_dim, k1 * grid.xy_dim
    lx, ly, lz = grid.xs[i0], grid.ys[j0], grid.zs[k0]
    wlx = (p[0] - lx) * grid.one_div_step
    wly = (p[1] - ly) * grid.one_div_step
    wlz = (p[2] - lz) * grid.one_div_step
    whx, why, whz = 1.0 - wlx, 1.0 - wly, 1.0 - wlz
    g = lambda idx: float(arr[idx])     # noqa: E731
    return (g(i0 + j0x + k0xy) * (whx * why * whz) +
            g(i1 + j0x + k0xy) * (wlx * why * whz) +
            g(i1 + j1x + k0xy) * (wlx * wly * whz) +
            g(i0 + j1x + k0xy) * (whx * wly * whz) +
            g(i0 + j0x + k1xy) * (whx * why * wlz) +
            g(i1 + j0x + k1xy) * (wlx * why * wlz) +
            g(i1 + j1x + k1xy) * (wlx * wly * wlz) +
            g(i0 + j1x + k1xy) * (whx * wly * wlz))


def ene_inter_UFF_interp(grid, ff_comps, lig):      # mol.ml:1012-1020
    res = 0.0
    for j in range(lig.n):
        res = res + trilin(grid, ff_comps[lig.t_a[j]], lig.get_xyz(j))
    return res


# ---------------------------------------------------------------------------------------------- pose builders
def rotate_then_translate_copy(lig, rot, pos):      # mol.ml:664-672
    m = lig.copy()
    m.centered_rotate(rot)
    m.translate_by(pos)
    return m


def center_rotate_translate_copy(lig, rot, pos):    # mol.ml:705-710
    m = lig.copy()
    m.center_()
    m.centered_rotate(rot)
    m.translate_by(pos)
    return m


# ---------------------------------------------------------------------------------------------- lds.ml Monte Carlo
class TooLong(Exception):           # mol.ml:587
    pass


class Rng:
    """Random.State stand-in: include/mmo_detmath.h's counter-based stream (NOT OCaml's generator)"""

    def __init__(self, seed):
        self.seed, self.ctr = seed, 0

    def float(self, scale):         # Random.State.float rng scale
        u = det_uniform(self.seed, self.ctr)
        self.ctr += 1
        return scale * u

    def int(self, n):               # Random.State.int rng n
        u = det_uniform(self.seed, self.ctr)
        self.ctr += 1
        v = int(u * float(n))
        return n - 1 if v >= n else v


def rand_angle(rng, max_rot):       # move.ml:14-15
    return rng.float(2.0 * max_rot) - max_rot


def rand_rot(rng, dr, rot):         # move.ml:20-31
    theta = rand_angle(rng, dr)
    axis = rng.int(3)
    rotate_by = (rot_rx, rot_ry, rot_rz)[axis](theta, det_cos_sin)
    return rot_mult(rotate_by, rot)


def rand_trans(rng, dt, pos):       # move.ml:46-54; V3.make's arguments are evaluated right to left: z, y, x
    z = rng.float(2.0) - 1.0
    y = rng.float(2.0) - 1.0
    x = rng.float(2.0) - 1.0
    return (pos[0] + x * dt, pos[1] + y * dt, pos[2] + z * dt)


def simulate_lig(centered_lig, ene_inter, nsteps, seed, rot0, pos0, roi, tweak_rbonds=True, enforce_ROI=True,
                 intra_nb=True, no_flip=False, temperature_K=293.15):
    """Lds.simulate_lig (lds.ml:741-1000) without its file output.  Returns (trace rows [curr_E, E_inter, E_intra,
    accepted(-1 = no test)], best_E, prev_E, counters dict)."""
    pi = 4.0 * math.atan(1.0)                        # math.ml:13
    p_max_rot = 15.0 * (pi / 180.0)                  # params.ml:11, math.ml to_radian
    p_max_trans = 0.15                               # params.ml:14
    block_size = 100                                 # params.ml:29
    max_rbond_flip = pi                              # params.ml:20
    target_low, target_high = 0.5 - 0.05, 0.5 + 0.05     # lds.ml:651-652
    beta = 1.0 / (0.0019872041 * temperature_K)      # lds.ml:66-67, const.ml:24
    rng = Rng(seed)
    max_rot, max_trans = p_max_rot, p_max_trans
    rot, pos = rot0, pos0
    start_conf = rotate_then_translate_copy(centered_lig, rot0, pos0)
    prev_lig = start_conf
    conf = centered_lig.copy()
    accepts_rejects = SW(block_size)
    roi_center, out_radius = roi[:3], roi[3]
    num_rbonds = len(centered_lig.rbonds)
    rbonds_block_size = block_size * num_rbonds
    flexible = tweak_rbonds and num_rbonds > 0

    def rotate_bond(i, lig0):                        # lds.ml:781-798
        if not flexible:
            return -1, centered_lig
        rbf = (1 << 62) if no_flip else block_size   # Params.rbond_flip_block (--no-flip: never)
        lig = lig0.copy()
        if i > 0 and i % rbf == 0:                   # Mol.flip_rbond, mol.ml:644-647
            b = rng.int(num_rbonds)
            lig.rotate_bond(b, rand_angle(rng, max_rbond_flip), det_cos_sin)
        else:                                        # Mol.tweak_rbond, mol.ml:635-638
            b = rng.int(num_rbonds)
            lig.rotate_bond(b, rand_angle(rng, lig.rbonds_dr[b]), det_cos_sin)
        if lig.radius() > 12.0:                      # Mol.check_elongation_exn lig Const.charged_cutoff
            raise TooLong()
        return b, lig

    if not intra_nb:
        ene_intra = lambda m: 0.0                    # noqa: E731   --no-E-intra
    elif flexible:
        ene_intra = ene_intra_UFFNB_brute
    else:                                            # lds.ml:706-712: constant of the centred ligand
        const = ene_intra_UFFNB_brute(centered_lig)
        ene_intra = lambda m: const                  # noqa: E731
    prev_E_intra = ene_intra(prev_lig)
    prev_E_inter = ene_inter(prev_lig)
    prev_E = prev_E_inter + prev_E_intra
    best_E = prev_E
    cnt = dict(acc_rigid=0, rej_rigid=0, acc_conf=0, rej_conf=0, ooroi=0, ezero=0, too_long=0, frames=0)
    trace = []
    rigid_step = conf_step = 0
    try:
        for frame in range(nsteps):
            xRIGID = (frame % 2 == 0)
            xCONF = not xRIGID
            just_rotated, conf_p = rotate_bond(conf_step, conf) if xCONF else (-1, conf)
            if xRIGID:                               # tuple evaluated right to left: rand_trans first
                pos_p = rand_trans(rng, max_trans, pos)
                rot_p = rand_rot(rng, max_rot, rot)
            else:
                rot_p, pos_p = rot, pos
            lig_p = center_rotate_translate_copy(conf_p, rot_p, pos_p)
            if xCONF:
                prev_E_intra = ene_intra(lig_p)
            prev_E_inter = ene_inter(lig_p)
            curr_E = prev_E_inter + prev_E_intra
            dist_roi = dist(roi_center, lig_p.center)
            accepted = -1
            def reset_run_params():                  # lds.ml:632-648 (conf, the accumulated conformer, is NOT reset)
                nonlocal max_rot, max_trans, rot, pos, prev_E, best_E, prev_lig
                max_rot, max_trans = p_max_rot, p_max_trans
                rot, pos = rot0, pos0
                prev_E = math.inf
                best_E = math.inf
                prev_lig = rotate_then_translate_copy(centered_lig, rot, pos)
                accepts_rejects.reset()

            # lds.ml:908-990:   if enforce_ROI then if dist_roi > out_radius then (reset) else (E = 0 test, Metropolis)
            # OCaml attaches the `else` to the nearest `if`: without --hard-ROI nothing below runs at all
            if enforce_ROI:
                if dist_roi > out_radius:
                    reset_run_params()
                    cnt["ooroi"] += 1
                elif prev_E_inter == 0.0:
                    reset_run_params()
                    cnt["ezero"] += 1
                else:
                    if curr_E <= prev_E or rng.float(1.0) < det_exp((-(curr_E - prev_E)) * beta):
                        accepted = 1
                        if xRIGID:
                            accepts_rejects.process(True)
                            cnt["acc_rigid"] += 1
                        else:
                            if just_rotated > -1:
                                conf_p.rbonds_sw[just_rotated].process(True)
                            cnt["acc_conf"] += 1
                        rot, pos, prev_E, prev_lig, conf = rot_p, pos_p, curr_E, lig_p, conf_p
                    else:
                        accepted = 0
                        if xRIGID:
                            accepts_rejects.process(False)
                            cnt["rej_rigid"] += 1
                        else:
                            if just_rotated > -1:
                                conf_p.rbonds_sw[just_rotated].process(False)
                            cnt["rej_conf"] += 1
                    if curr_E < best_E:
                        best_E = curr_E
                    if xRIGID and rigid_step > 0 and rigid_step % block_size == 0:       # lds.ml:586-600
                        ar = accepts_rejects.get_ratio()
                        if ar <= target_low:
                            max_trans = 0.95 * max_trans
                            max_rot = 0.95 * max_rot
                        elif ar >= target_high:
                            max_trans = 1.05 * max_trans
                            m = 1.05 * max_rot
                            max_rot = pi if pi <= m else m
                    if flexible and xCONF and conf_step > 0 and conf_step % rbonds_block_size == 0:
                        for b in range(num_rbonds):                                       # lds.ml:603-621, on conf'
                            ar = conf_p.rbonds_sw[b].get_ratio()
                            if ar <= target_low:
                                conf_p.rbonds_dr[b] = 0.95 * conf_p.rbonds_dr[b]
                            elif ar >= target_high:
                                m = 1.05 * conf_p.rbonds_dr[b]
                                conf_p.rbonds_dr[b] = pi if pi <= m else m
            trace.append((curr_E, prev_E_inter, prev_E_intra, float(accepted)))
            if xRIGID:
                rigid_step += 1
            else:
                conf_step += 1
            cnt["frames"] = frame + 1
    except TooLong:
        cnt["too_long"] = 1
    return trace, best_E, prev_E, cnt


# ---------------------------------------------------------------------------------------------- include/mmo_detmath.h
_M64 = (1 << 64) - 1


def det_u64(seed, counter):
    z = (seed + 0x9E3779B97F4A7C15 * (counter + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def det_uniform(seed, counter):
    return float(det_u64(seed, counter) >> 11) * (1.0 / 9007199254740992.0)


def _det_floor(x):
    t = float(int(x))
    return t - 1.0 if t > x else t


def det_cos_sin(x):
    """(cos, sin) as Math.cos_sin returns them (math.ml:5-11), from the deterministic kernels"""
    two_over_pi = 6.36619772367581382433e-01
    pio2_1 = 1.57079632673412561417e+00
    pio2_1t = 6.07710050650619224932e-11
    kf = _det_floor(x * two_over_pi + 0.5)
    k = int(kf)
    r = (x - kf * pio2_1) - kf * pio2_1t
    z = r * r
    S1, S2, S3 = -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04
    S4, S5, S6 = 2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10
    ps = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)))
    sr = r + (r * z) * (S1 + z * ps)
    C1, C2, C3 = 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05
    C4, C5, C6 = -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11
    pc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))))
    cr = (1.0 - 0.5 * z) + z * pc
    q = k & 3
    if q == 0:
        s, c = sr, cr
    elif q == 1:
        s, c = cr, -sr
    elif q == 2:
        s, c = -sr, -cr
    else:
        s, c = -cr, sr
    return c, s


def det_exp(x):
    if x > 0.0:
        x = 0.0
    if x < -700.0:
        return 0.0
    inv_ln2 = 1.44269504088896338700e+00
    ln2_hi, ln2_lo = 6.93147180369123816490e-01, 1.90821492927058770002e-10
    kf = _det_floor(x * inv_ln2 + 0.5)
    r = (x - kf * ln2_hi) - kf * ln2_lo
    p = 1.0 / 6227020800.0
    for d in (479001600.0, 39916800.0, 3628800.0, 362880.0, 40320.0, 5040.0, 720.0, 120.0, 24.0, 6.0):
        p = 1.0 / d + r * p
    p = 0.5 + r * p
    p = 1.0 + r * p
    p = 1.0 + r * p
    k = int(kf)
    scale = struct.unpack("d", struct.pack("Q", (k + 1023) << 52))[0]
    return p * scale


# ---------------------------------------------------------------------------------------------- SO3.ml, quat.ml
def libm_cos_sin(theta):            # math.ml:9-11 with the platform's libm (host-side code: SO3.rotations runs once per lds run)
    return math.cos(theta), math.sin(theta)


MATH_PI = 4.0 * math.atan(1.0)      # math.ml:13
MATH_TWO_PI = 2.0 * MATH_PI         # math.ml:15
SO3_PHI = math.sqrt(2.0)            # SO3.ml:13
SO3_PSI = 1.533751168755204288118041     # SO3.ml:14


def super_fibonacci(n, i):          # SO3.ml:18-30, n already a float; returns the quaternion (w, x, y, z)
    s = float(i) + 0.5
    t = s / n
    d = MATH_TWO_PI * s
    c_r = math.sqrt(t)
    c_R = math.sqrt(1.0 - t)
    alpha = d / SO3_PHI
    beta = d / SO3_PSI
    return (c_r * math.sin(alpha), c_r * math.cos(alpha), c_R * math.sin(beta), c_R * math.cos(beta))


def quat_to_axis_angle(q):          # quat.ml:33-37
    w, x, y, z = q
    mag = math.sqrt(x * x + y * y + z * z)
    axis = (x / mag, y / mag, z / mag)
    theta = 2.0 * math.atan2(mag, w)
    return axis, theta


def so3_rotations(n):               # SO3.ml:32-39
    out = []
    for i in range(n):
        axis, angle = quat_to_axis_angle(super_fibonacci(float(n), i))
        out.append(rot_of_axis_angle(axis, angle, libm_cos_sin))
    return out


# ---------------------------------------------------------------------------------------------- lds.ml bitmasks, G3D.ml clash test
R_H2O = 1.4                         # const.ml:10


def coord_of_point(grid, p):        # grid.ml:87-90; int_of_float truncates toward zero
    return (int((p[0] - grid.xs[0]) / grid.step), int((p[1] - grid.ys[0]) / grid.step), int((p[2] - grid.zs[0]) / grid.step))


def atom_bitmask_set(grid, mask, xyz, radius, b):       # lds.ml:148-169; mask: a Python list of bools, Bitv index order
    i0, j0, k0 = coord_of_point(grid, xyz)
    r_steps = int(math.ceil(radius / grid.step))
    r2 = radius * radius
    for i in range(i0 - r_steps, i0 + r_steps + 1):
        x = grid.xs[i]              # OCaml raises Invalid_argument outside the array; the fixtures keep atoms inside
        for j in range(j0 - r_steps, j0 + r_steps + 1):
            y = grid.ys[j]
            for k in range(k0 - r_steps, k0 + r_steps + 1):
                if dist2(xyz, (x, y, grid.zs[k])) < r2:
                    mask[i + j * grid.x_dim + k * grid.xy_dim] = b


def vdW_volume(grid, atoms, radii):             # lds.ml:187-196
    mask = [False] * (grid.x_dim * grid.y_dim * grid.z_dim)
    for xyz, r in zip(atoms, radii):
        atom_bitmask_set(grid, mask, xyz, r, True)
    return mask


def first_solvent_shell(grid, atoms, radii):    # lds.ml:172-184
    mask = [False] * (grid.x_dim * grid.y_dim * grid.z_dim)
    for xyz, r in zip(atoms, radii):
        atom_bitmask_set(grid, mask, xyz, r + R_H2O, True)
    for xyz, r in zip(atoms, radii):
        atom_bitmask_set(grid, mask, xyz, r, False)
    return mask


def vdW_clash_OR(grid, bitmask, p):             # G3D.ml:162-186
    i0 = int(p[0] * grid.one_div_step)
    j0 = int(p[1] * grid.one_div_step)
    k0 = int(p[2] * grid.one_div_step)
    i1, j1, k1 = i0 + 1, j0 + 1, k0 + 1
    j0x, j1x = j0 * grid.x_dim, j1 * grid.x_dim
    k0xy, k1xy = k0 * grid.xy_dim, k1 * grid.xy_dim
    return (bitmask[i0 + j0x + k0xy] or bitmask[i1 + j0x + k0xy] or bitmask[i1 + j1x + k0xy] or bitmask[i0 + j1x + k0xy] or
            bitmask[i0 + j0x + k1xy] or bitmask[i1 + j0x + k1xy] or bitmask[i1 + j1x + k1xy] or bitmask[i0 + j1x + k1xy])


def protein_ligand_clash(grid, bitmask, lig_atoms):     # mol.ml:1195-1203
    for p in lig_atoms:
        if vdW_clash_OR(grid, bitmask, p):
            return True
    return False


# ---------------------------------------------------------------------------------------------- lds.ml exhaustive scan
def centered_rotate_copy(lig, r):               # mol.ml:664-667
    m = lig.copy()
    m.centered_rotate(r)
    return m


def translate_copy_to(m, p):                    # mol.ml:689-696
    c = m.copy()
    c.translate_by(v_diff(p, c.center))
    c.center = p
    return c


def roi_is_inside(roi, p):                      # ROI.ml:65-66: V3.dist2 s.c p < s.out_r2, out_r2 = out_r * out_r
    return dist2(roi[:3], p) < roi[3] * roi[3]


def exhaustive_rigid_ligand_docking(topk, roi, trans_step, rotations, centered_lig, score, clash=None):
    """lds.ml:1040-1114 without the two vdW prefilters unless `clash` (Mol.protein_ligand_clash on the pose) is given:
    returns (top scores, best score, best frame, poses scored).  TopK keeps the k highest of -score (Cpm.TopKeeper,
    not vendored: only the kept scores are output, lds.ml:1110-1113); the best pose is the FIRST one with the lowest
    score in loop order z, y, x, rotation (strict <, lds.ml:1099)."""
    x0, x1 = roi[0] - roi[3], roi[0] + roi[3]   # ROI.get_bounds (ROI.ml:76-82)
    y0, y1 = roi[1] - roi[3], roi[1] + roi[3]
    z0, z1 = roi[2] - roi[3], roi[2] + roi[3]
    g = Grid(trans_step, x1 - x0, y1 - y0, z1 - z0)           # Bbox.create_6f: dims = high - low; Grid.from_box
    xs = [x0 + v for v in g.xs]                 # A.map ((+.) x_min) g.xs
    ys = [y0 + v for v in g.ys]
    zs = [z0 + v for v in g.zs]
    n_rot = len(rotations)
    best_score, best_frame = float("inf"), -1
    kept = []
    n_scored = 0
    rotated = [centered_rotate_copy(centered_lig, r) for r in rotations]
    for k, z in enumerate(zs):
        z_dim = k * g.xy_dim
        for j, y in enumerate(ys):
            jz_dim = j * g.x_dim + z_dim
            for i, x in enumerate(xs):
                pos = (x, y, z)
                if not roi_is_inside(roi, pos):
                    continue
                for rot_i, rot_lig in enumerate(rotated):
                    lig2 = translate_copy_to(rot_lig, pos)
                    if clash is not None and clash(lig2):
                        continue
                    curr = score(lig2)
                    n_scored += 1
                    kept.append(curr)
                    if curr < best_score:
                        best_score = curr
                        best_frame = rot_i + n_rot * (i + jz_dim)
    kept.sort()
    return kept[:topk] if topk > 0 else [], best_score, best_frame, n_scored


# ---------------------------------------------------------------------------------------------- lds.ml desolvation (N4)
CHARGED_CUTOFF = 12.0                                               # const.ml:12
CHARGED_CUTOFF_SQUARED = CHARGED_CUTOFF * CHARGED_CUTOFF            # const.ml:14
DESOLVATION = (1.0 / 4.0 - 1.0 / 78.5) / (8.0 * MATH_PI)            # const.ml:18-31


def ijk_of_idx(grid, idx):          # grid.ml:101-105
    k = idx // grid.xy_dim
    j = (idx - k * grid.xy_dim) // grid.x_dim
    i = idx - (k * grid.xy_dim + j * grid.x_dim)
    return i, j, k


def protein_desolv(roi, grid, shell, prot):     # lds.ml:204-236; BST.neighbors restated as index order with dist <= cutoff
    voxel_vol = grid.step * grid.step * grid.step
    n = grid.x_dim * grid.y_dim * grid.z_dim
    assert n == len(shell)
    res = [0.0] * n
    for idx in range(n):                        # Bitv.iteri_true: ascending index
        if not shell[idx]:
            continue
        i, j, k = ijk_of_idx(grid, idx)
        x_p = (grid.xs[i], grid.ys[j], grid.zs[k])
        if roi_is_inside(roi, x_p):
            for a in range(prot.n):
                x_j = prot.get_xyz(a)
                if dist(x_p, x_j) <= CHARGED_CUTOFF:
                    d2 = dist2(x_p, x_j)
                    x = prot.q_a[a] / d2
                    res[idx] = res[idx] + (x * x)
            res[idx] = DESOLVATION * (voxel_vol * res[idx])
    return res


def desolvation_penalty(grid, contribs, prot_shell, lig_atoms, lig_charges, lig_radii):     # lds.ml:239-267
    voxel_vol = grid.step * grid.step * grid.step
    lig_shell = first_solvent_shell(grid, lig_atoms, lig_radii)
    lig_desolv = 0.0
    prot_desolv = 0.0
    for idx in range(len(prot_shell)):
        if not (prot_shell[idx] and lig_shell[idx]):        # Bitv.bw_and, then iteri_true
            continue
        i, j, k = ijk_of_idx(grid, idx)
        x_p = (grid.xs[i], grid.ys[j], grid.zs[k])
        for x_j, q_j in zip(lig_atoms, lig_charges):
            d2 = dist2(x_p, x_j)
            if d2 < CHARGED_CUTOFF_SQUARED:
                x = q_j / d2
                lig_desolv = lig_desolv + (x * x)
        prot_desolv = prot_desolv + contribs[idx]
    return prot_desolv, DESOLVATION * (voxel_vol * lig_desolv)


# ---------------------------------------------------------------------------------------------- lds.ml N3 bitmasks
def bitmask_whole_protein(grid, atoms):         # lds.ml:97-145: nearest protein atom closer than Const.charged_cutoff
    mask = [False] * (grid.x_dim * grid.y_dim * grid.z_dim)
    for i in range(grid.x_dim):
        x = grid.xs[i]
        for j in range(grid.y_dim):
            y = grid.ys[j]
            for k in range(grid.z_dim):
                p = (x, y, grid.zs[k])
                # BST.nearest_neighbor returns the distance to the closest atom: min over atoms of V3.dist
                if min(dist(p, a) for a in atoms) < CHARGED_CUTOFF:
                    mask[i + j * grid.x_dim + k * grid.xy_dim] = True
    return mask


def bitmask_ROI_only(roi, grid):                # lds.ml:269-305
    c = roi[:3]
    r = roi[3] + (CHARGED_CUTOFF * 2.0)
    r2 = r * r
    mask = [False] * (grid.x_dim * grid.y_dim * grid.z_dim)
    for i in range(grid.x_dim):
        x = grid.xs[i]
        for j in range(grid.y_dim):
            y = grid.ys[j]
            for k in range(grid.z_dim):
                if dist2(c, (x, y, grid.zs[k])) < r2:
                    mask[i + j * grid.x_dim + k * grid.xy_dim] = True
    return mask
