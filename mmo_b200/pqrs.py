"""Host-side readers/writers for the text formats that feed the scoring path.

Mirrors the reference's input layer so that real `lds` inputs can be fed to the C-ABI:
  * `.pqrs`   -- src/pqrs.ml:19-87, src/mol.ml:368-402 (ligand: `N:R:name`, receptor: `N:name`)
  * ROI .bild -- src/ROI.ml:22-32 (exactly one `.sphere x y z r` line)
  * mol2 -> pqrs preparation (rotatable bonds, rotatable groups, topological distances)
              -- src/mol2.ml:139-228, src/mol_graph.ml:41-200, src/mol2pqrs.ml:10-43

Pure host logic (numpy only); no energies are computed here.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from collections import deque

import numpy as np

# src/ptable.ml:86-100 (sym2anu) and 41-54 (vdW_radii)
SYM2ANUM = {"C": 6, "H": 1, "N": 7, "O": 8, "P": 15, "S": 16, "F": 9, "Cl": 17, "Br": 35, "I": 53, "Mg": 12}
ANUM2SYM = {v: k for k, v in SYM2ANUM.items()}
VDW_RADII = {1: 1.2, 6: 1.7, 7: 1.6, 8: 1.55, 9: 1.5, 12: 2.2, 15: 1.95, 16: 1.8, 17: 1.8, 35: 1.9, 53: 2.1}

# src/ptable.ml:168-207 anum_of_mol2_type (only the types whose element the FF supports matter)
_MOL2_PREFIX = {"H": 1, "C": 6, "N": 7, "O": 8, "F": 9, "Mg": 12, "P": 15, "S": 16, "Cl": 17, "Br": 35, "I": 53}


def anum_of_mol2_type(typ: str) -> int:
    head = typ.split(".")[0]
    if head not in _MOL2_PREFIX:
        raise ValueError(f"unsupported mol2 atom type: {typ}")
    return _MOL2_PREFIX[head]


@dataclass
class Mol:
    """SoA molecule, the layout of `Mol.t` (src/mol.ml:17-35)."""
    name: str
    xs: np.ndarray
    ys: np.ndarray
    zs: np.ndarray
    q: np.ndarray
    r: np.ndarray
    anum: np.ndarray                      # int32
    # ligand-only parts
    dists: np.ndarray | None = None       # int32 N*N, element (i,j) at i + j*N (mol.ml:151-152)
    rb_left: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rb_right: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rgroups: list = field(default_factory=list)   # list of int32 arrays (axis tip excluded)
    typ: np.ndarray | None = None         # FF atom types (mol.ml:456-462)

    @property
    def n(self) -> int:
        return len(self.xs)

    @property
    def n_rbonds(self) -> int:
        return len(self.rb_left)

    def rgroup_csr(self):
        off = np.zeros(self.n_rbonds + 1, np.int32)
        for i, g in enumerate(self.rgroups):
            off[i + 1] = off[i] + len(g)
        idx = np.concatenate(self.rgroups).astype(np.int32) if self.rgroups else np.zeros(0, np.int32)
        return off, idx

    def copy(self) -> "Mol":
        return Mol(self.name, self.xs.copy(), self.ys.copy(), self.zs.copy(), self.q.copy(), self.r.copy(),
                   self.anum.copy(), None if self.dists is None else self.dists.copy(),
                   self.rb_left.copy(), self.rb_right.copy(), [g.copy() for g in self.rgroups],
                   None if self.typ is None else self.typ.copy())


def _parse_atoms(lines, n):
    xs = np.empty(n); ys = np.empty(n); zs = np.empty(n); q = np.empty(n); r = np.empty(n)
    anum = np.empty(n, np.int32)
    for i in range(n):
        t = lines[i].split()
        xs[i], ys[i], zs[i], q[i], r[i] = map(float, t[:5])
        anum[i] = SYM2ANUM[t[5]]
    return xs, ys, zs, q, r, anum


def read_receptor_pqrs(fn: str) -> Mol:
    """src/mol.ml:413-417 receptor_of_pqrs_file."""
    with open(fn) as f:
        lines = f.read().split("\n")
    head = lines[0].split(":")
    n = int(head[0])
    name = head[-1]
    xs, ys, zs, q, r, anum = _parse_atoms(lines[1:], n)
    return Mol(name, xs, ys, zs, q, r, anum)


def read_ligands_pqrs(fn: str) -> list[Mol]:
    """src/mol.ml:420-440 ligands_of_pqrs_file / 378-402 ligand_pqrs_read_one."""
    with open(fn) as f:
        lines = [l for l in f.read().split("\n")]
    res = []
    p = 0
    while p < len(lines) and lines[p].strip():
        head = lines[p].split(":")
        n, nrb, name = int(head[0]), int(head[1]), ":".join(head[2:])
        p += 1
        xs, ys, zs, q, r, anum = _parse_atoms(lines[p:], n)
        p += n
        left, right, groups = [], [], []
        for _ in range(nrb):                       # src/pqrs.ml:39-61 parse_rot_bond
            axis, bits = lines[p].split("=")
            b, e = map(int, axis.split("-"))
            flags = [t == "1" for t in bits.split()]
            assert len(flags) == n and flags[b] != flags[e]
            l, rr = (b, e) if flags[e] else (e, b)  # axis goes fixed -> movable
            left.append(l); right.append(rr)
            groups.append(np.array([i for i in range(n) if flags[i] and i != rr], np.int32))  # pqrs.ml:80-87
            p += 1
        dists = np.zeros(n * n, np.int32)          # src/pqrs.ml:63-77 parse_dist_matrix
        for i in range(n):
            row = lines[p].split(" ")
            assert len(row) == n
            for j, d in enumerate(row):
                dists[i + j * n] = int(d)
            p += 1
        res.append(Mol(name, xs, ys, zs, q, r, anum, dists, np.array(left, np.int32),
                       np.array(right, np.int32), groups))
    names = [m.name for m in res]
    if len(set(names)) != len(names):              # mol.ml:428-438
        raise ValueError(f"duplicate molecule names in {fn}")
    return res


def read_roi_bild(fn: str):
    """src/ROI.ml:22-32 from_bild -> (cx, cy, cz, r)."""
    ok = [l for l in open(fn).read().split("\n") if l.startswith(".sphere ")]
    if len(ok) != 1:
        raise ValueError(f"ROI.from_bild: several sphere lines in: {fn}")
    t = ok[0].split()
    return tuple(float(v) for v in t[1:5])


# ---------------------------------------------------------------------------------------------
# mol2 -> pqrs (src/mol2pqrs.ml:10-43)

def parse_mol2(fn: str):
    """First molecule of a mol2 file -> (name, atoms[(x,y,z,q,anum)], bonds[(src,dst,type)]).
    src/mol2.ml:184-228 (atom line), 169-182 + 139-149 (bond line / bond order)."""
    lines = open(fn).read().split("\n")
    i = lines.index("@<TRIPOS>MOLECULE")
    name = lines[i + 1].strip()
    n_atoms, n_bonds = map(int, lines[i + 2].split()[:2])
    a0 = lines.index("@<TRIPOS>ATOM") + 1
    atoms = []
    for l in lines[a0:a0 + n_atoms]:
        t = l.split()
        atoms.append((float(t[2]), float(t[3]), float(t[4]), float(t[8]), anum_of_mol2_type(t[5])))
    b0 = lines.index("@<TRIPOS>BOND") + 1
    bonds = []
    for l in lines[b0:b0 + n_bonds]:
        t = l.split()
        bonds.append((int(t[1]) - 1, int(t[2]) - 1, t[3]))
    return name, atoms, bonds


_BOND_ORDER = {"1": 1.0, "2": 2.0, "3": 3.0, "ar": 1.5, "am": 1.0}   # src/mol2.ml:139-149


def _components_without(n, bonds, skip):
    adj = [[] for _ in range(n)]
    for bi, (s, d, _) in enumerate(bonds):
        if bi != skip:
            adj[s].append(d); adj[d].append(s)
    comp = [-1] * n
    for s in range(n):                    # label = smallest atom index of the component,
        if comp[s] >= 0:                  # as the min-label propagation of mol_graph.ml:160-186 yields
            continue
        comp[s] = s
        dq = deque([s])
        while dq:
            u = dq.popleft()
            for v in adj[u]:
                if comp[v] < 0:
                    comp[v] = s; dq.append(v)
    return comp


def mol2_to_ligand(fn: str) -> Mol:
    name, atoms, bonds = parse_mol2(fn)
    n = len(atoms)
    adj = [[] for _ in range(n)]
    deg = [0] * n
    for s, d, _ in bonds:
        adj[s].append(d); adj[d].append(s); deg[s] += 1; deg[d] += 1
    # all-pairs topological distances (mol_graph.ml:45-63; unit edge weights -> BFS)
    dists = np.zeros(n * n, np.int32)
    for s in range(n):
        dist = [-1] * n
        dist[s] = 0
        dq = deque([s])
        while dq:
            u = dq.popleft()
            for v in adj[u]:
                if dist[v] < 0:
                    dist[v] = dist[u] + 1; dq.append(v)
        if min(dist) < 0:
            raise ValueError("disconnected atom")     # Mol_graph.Disconnected_atom
        for j in range(n):
            dists[s + j * n] = dist[j]
    left, right, groups = [], [], []
    for bi, (s, d, typ) in enumerate(bonds):      # mol_graph.ml:128-138 list_rotatable_bonds
        if _BOND_ORDER[typ] != 1.0 or deg[s] <= 1 or deg[d] <= 1:
            continue
        comp = _components_without(n, bonds, bi)
        if comp[s] == comp[d]:                        # ring bond (mol_graph.ml:108-117)
            continue
        # mol_graph.ml:141-200: movable side = the smaller component (ties: Hashtbl order, unpinned;
        # resolved here towards the component holding the lower atom index)
        g1, g2 = sorted(set(comp))
        c1, c2 = comp.count(g1), comp.count(g2)
        small = g1 if c1 <= c2 else g2
        flags = [c == small for c in comp]
        l, r = (s, d) if flags[d] else (d, s)         # pqrs.ml:49-56
        left.append(l); right.append(r)
        groups.append(np.array([i for i in range(n) if flags[i] and i != r], np.int32))
    a = np.array([(x, y, z, q) for x, y, z, q, _ in atoms])
    anum = np.array([t[4] for t in atoms], np.int32)
    rad = np.array([VDW_RADII[int(z)] for z in anum])
    return Mol(name, a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy(), a[:, 3].copy(), rad, anum, dists,
               np.array(left, np.int32), np.array(right, np.int32), groups)


def _g(x: float) -> str:
    return "%g" % x


def write_ligand_pqrs(fn: str, m: Mol, mode: str = "w") -> None:
    """Same text as mol2pqrs.ml:10-43 emits (values through `%g`, mol2.ml:67-69)."""
    n = m.n
    out = [f"{n}:{m.n_rbonds}:{m.name}"]
    for i in range(n):
        out.append(" ".join([_g(m.xs[i]), _g(m.ys[i]), _g(m.zs[i]), _g(m.q[i]), _g(m.r[i]), ANUM2SYM[int(m.anum[i])]]))
    for l, r, g in zip(m.rb_left, m.rb_right, m.rgroups):
        member = set(int(v) for v in g) | {int(r)}
        out.append(f"{l}-{r}=" + "".join(" 1" if i in member else " 0" for i in range(n)))
    for i in range(n):
        out.append(" ".join(str(int(m.dists[i + j * n])) for j in range(n)))
    with open(fn, mode) as f:
        f.write("\n".join(out) + "\n")


def write_receptor_pqrs(fn: str, m: Mol) -> None:
    out = [f"{m.n}:{m.name}"]
    for i in range(m.n):
        out.append(" ".join([_g(m.xs[i]), _g(m.ys[i]), _g(m.zs[i]), _g(m.q[i]), _g(m.r[i]), ANUM2SYM[int(m.anum[i])]]))
    with open(fn, "w") as f:
        f.write("\n".join(out) + "\n")


def assign_ff_types(ligs: list[Mol]):
    """(anum, exact charge) -> dense id in first-seen order (src/mol.ml:280-293, 456-469).
    Returns (type_anum int32[T], type_q float64[T]) and fills `typ` of every ligand."""
    ids: dict = {}
    for m in ligs:
        for a, q in zip(m.anum, m.q):
            key = (int(a), float(q))
            if key not in ids:
                ids[key] = len(ids)
    for m in ligs:
        m.typ = np.array([ids[(int(a), float(q))] for a, q in zip(m.anum, m.q)], np.int32)
    keys = sorted(ids, key=ids.get)
    return np.array([k[0] for k in keys], np.int32), np.array([k[1] for k in keys], np.float64)
