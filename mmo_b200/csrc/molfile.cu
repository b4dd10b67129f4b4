// molfile.cu -- N2: ligand / receptor preparation on the host, in C++ inside the library (no subprocess, no Python).
// Restates the reference's input layer:
//   mol2 reader            src/mol2.ml:139-149 (bond order), 184-228 (atom / bond lines), 243-320 (one molecule),
//                          lone pairs dropped (mol2.ml:54-56), element from the mol2 type (src/ptable.ml:168-207)
//   molecular graph        src/mol_graph.ml:45-63 (all-pairs topological distances), 93-102 (degrees),
//                          108-138 (rotatable bond = single, not in a ring, no terminal atom),
//                          141-200 (rotatable group = the smaller side once the bond is cut)
//   .pqrs text             src/mol2pqrs.ml:10-58 (writer, values through %g, mol2.ml:67-69),
//                          src/pqrs.ml:19-87 + src/mol.ml:368-440 (reader: rotatable bond lines, distance matrix)
//   FF atom types          src/mol.ml:280-293, 456-469 ((anum, exact charge) -> id in first-seen order over the file)
// Pure host code: everything but mmo_molfile_ligand works without a GPU.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <map>
#include <queue>
#include <sstream>

namespace {

struct Molecule {
    std::string name;
    std::vector<double> x, y, z, q, r;
    std::vector<int32_t> anum;
    bool is_ligand = false;
    std::vector<int32_t> dists;                  // n*n, element (i, j) at i + j*n (mol.ml:151-152)
    std::vector<int32_t> rb_src, rb_dst;         // as written in the file (mol2 bond order)
    std::vector<int32_t> rb_left, rb_right;      // fixed -> movable (pqrs.ml:49-56)
    std::vector<std::vector<uint8_t>> rb_flags;  // movable side, axis tip included (the pqrs text)
    std::vector<int32_t> typ;                    // FF types, assigned over the whole file
    // mol2 text kept for the writer (mol2.ml:40-49, 71-74): atom names / types and bonds, lone pairs removed
    std::vector<std::string> aname, atype;
    std::vector<int32_t> b_src, b_dst;
    std::vector<std::string> b_typ;
    int n() const { return (int)x.size(); }
};

// src/ptable.ml:41-54, 86-100
struct Elt { const char *sym; int anum; double radius; };
const Elt kElts[] = {{"H", 1, 1.2}, {"C", 6, 1.7}, {"N", 7, 1.6}, {"O", 8, 1.55}, {"F", 9, 1.5}, {"Mg", 12, 2.2},
                     {"P", 15, 1.95}, {"S", 16, 1.8}, {"Cl", 17, 1.8}, {"Br", 35, 1.9}, {"I", 53, 2.1}};
const Elt *elt_by_sym(const std::string &s) {
    for (const Elt &e : kElts) if (s == e.sym) return &e;
    return nullptr;
}
const Elt *elt_by_anum(int a) {
    for (const Elt &e : kElts) if (a == e.anum) return &e;
    return nullptr;
}

std::vector<std::string> split_ws(const std::string &l) {
    std::vector<std::string> t;
    std::istringstream is(l);
    std::string w;
    while (is >> w) t.push_back(w);
    return t;
}
std::string rstrip(std::string s) {
    while (!s.empty() && (s.back() == '\r' || s.back() == '\n' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
    return s;
}
std::string strip(std::string s) {
    s = rstrip(s);
    size_t k = 0;
    while (k < s.size() && (s[k] == ' ' || s[k] == '\t')) k++;
    return s.substr(k);
}

// graph part of mol2pqrs: distances, rotatable bonds, rotatable groups.  false = disconnected atom
// (Mol_graph.Disconnected_atom: the reference logs an error and emits nothing for that molecule)
bool analyse_graph(Molecule &m, const std::vector<int> &bs, const std::vector<int> &bd, const std::vector<double> &bo) {
    const int n = m.n(), nbonds = (int)bs.size();
    std::vector<std::vector<int>> adj(n);
    std::vector<int> deg(n, 0);
    for (int b = 0; b < nbonds; b++) { adj[bs[b]].push_back(bd[b]); adj[bd[b]].push_back(bs[b]); deg[bs[b]]++; deg[bd[b]]++; }
    m.dists.assign((size_t)n * n, 0);
    std::vector<int> dist(n);
    for (int s = 0; s < n; s++) {                // unit edge weights: BFS == the reference's shortest paths
        std::fill(dist.begin(), dist.end(), -1);
        dist[s] = 0;
        std::queue<int> qu;
        qu.push(s);
        while (!qu.empty()) {
            int u = qu.front(); qu.pop();
            for (int v : adj[u]) if (dist[v] < 0) { dist[v] = dist[u] + 1; qu.push(v); }
        }
        for (int j = 0; j < n; j++) {
            if (dist[j] < 0) return false;
            m.dists[s + (size_t)j * n] = dist[j];
        }
    }
    std::vector<int> comp(n);
    for (int b = 0; b < nbonds; b++) {
        if (bo[b] != 1.0 || deg[bs[b]] <= 1 || deg[bd[b]] <= 1) continue;
        // components once bond b is cut; label = smallest atom index, as the min-label propagation yields
        std::fill(comp.begin(), comp.end(), -1);
        for (int s = 0; s < n; s++) {
            if (comp[s] >= 0) continue;
            comp[s] = s;
            std::queue<int> qu;
            qu.push(s);
            while (!qu.empty()) {
                int u = qu.front(); qu.pop();
                for (int v : adj[u]) {
                    if ((u == bs[b] && v == bd[b]) || (u == bd[b] && v == bs[b])) continue;
                    if (comp[v] < 0) { comp[v] = s; qu.push(v); }
                }
            }
        }
        if (comp[bs[b]] == comp[bd[b]]) continue;            // ring bond
        const int g1 = std::min(comp[bs[b]], comp[bd[b]]), g2 = std::max(comp[bs[b]], comp[bd[b]]);
        int c1 = 0, c2 = 0;
        for (int i = 0; i < n; i++) { c1 += comp[i] == g1; c2 += comp[i] == g2; }
        // movable = the smaller side; ties go to the group holding the lower atom index (the reference
        // takes whichever binding its hash table lists first: unpinned)
        const int small = c1 <= c2 ? g1 : g2;
        std::vector<uint8_t> flags(n);
        for (int i = 0; i < n; i++) flags[i] = comp[i] == small;
        m.rb_src.push_back(bs[b]); m.rb_dst.push_back(bd[b]);
        const bool dst_moves = flags[bd[b]];
        m.rb_left.push_back(dst_moves ? bs[b] : bd[b]);
        m.rb_right.push_back(dst_moves ? bd[b] : bs[b]);
        m.rb_flags.push_back(flags);
    }
    return true;
}

int parse_int(const std::string &s, bool &ok) {
    char *e = nullptr;
    long v = strtol(s.c_str(), &e, 10);
    if (e == s.c_str() || *e != 0) ok = false;
    return (int)v;
}
double parse_dbl(const std::string &s, bool &ok) {
    char *e = nullptr;
    double v = strtod(s.c_str(), &e);
    if (e == s.c_str() || *e != 0) ok = false;
    return v;
}

}  // namespace

struct mmo_molfile {
    std::vector<Molecule> mols;
    int n_skipped = 0;
    std::vector<int32_t> type_anum;
    std::vector<double> type_q;
    void assign_types() {                         // mol.ml:280-293: first-seen order over ligands, then atoms
        std::map<std::pair<int32_t, double>, int32_t> ids;
        type_anum.clear(); type_q.clear();
        for (Molecule &m : mols) {
            m.typ.resize(m.n());
            for (int i = 0; i < m.n(); i++) {
                auto key = std::make_pair(m.anum[i], m.q[i]);
                auto it = ids.find(key);
                if (it == ids.end()) {
                    it = ids.emplace(key, (int32_t)ids.size()).first;
                    type_anum.push_back(m.anum[i]); type_q.push_back(m.q[i]);
                }
                m.typ[i] = it->second;
            }
        }
    }
};

using namespace mmo;

extern "C" {

// analyse = false: Mol2.read_one_from_file as scissors uses it (src/scissors.ml:48-55): the first molecule only, atoms and
// bonds as written, no graph analysis (a protein is not one connected molecule), never skipped for its topology
static int read_mol2(const char *path, bool analyse, mmo_molfile **out) {
    MMO_REQUIRE(path && out, "mmo_molfile_read_mol2: null pointer");
    *out = nullptr;
    std::ifstream in(path);
    MMO_REQUIRE(in.good(), "mmo_molfile_read_mol2: cannot open %s", path);
    std::vector<std::string> lines;
    for (std::string l; std::getline(in, l);) lines.push_back(rstrip(l));
    mmo_molfile *f = new mmo_molfile();
    size_t p = 0;
    while (p < lines.size()) {
        if (lines[p] != "@<TRIPOS>MOLECULE") { p++; continue; }
        size_t end = p + 1;
        while (end < lines.size() && lines[end] != "@<TRIPOS>MOLECULE") end++;
        // one molecule block [p, end)
        Molecule m;
        m.is_ligand = true;
        bool ok = p + 2 < end;
        int n_atoms = 0, n_bonds = 0;
        if (ok) {
            m.name = strip(lines[p + 1]);
            auto t = split_ws(lines[p + 2]);
            ok = t.size() >= 2;
            if (ok) { n_atoms = parse_int(t[0], ok); n_bonds = parse_int(t[1], ok); }
        }
        size_t a0 = p + 3;
        while (ok && a0 < end && lines[a0] != "@<TRIPOS>ATOM") a0++;
        ok = ok && a0 + 1 + (size_t)n_atoms <= end;
        std::vector<int> keep;                    // atom id (0-based, file order) -> index after lone-pair removal
        if (ok) {
            keep.assign(n_atoms, -1);
            for (int i = 0; i < n_atoms && ok; i++) {
                auto t = split_ws(lines[a0 + 1 + i]);      // id name x y z type resnum resname charge
                if (t.size() < 9) { ok = false; break; }
                if (t[5] == "LP") continue;                // lone pair: dropped together with its bonds
                const std::string head = t[5].substr(0, t[5].find('.'));
                const Elt *e = elt_by_sym(head);
                if (!e) { set_error("mmo_molfile_read_mol2: unsupported mol2 atom type %s in %s", t[5].c_str(), m.name.c_str()); ok = false; break; }
                keep[i] = m.n();
                m.x.push_back(parse_dbl(t[2], ok)); m.y.push_back(parse_dbl(t[3], ok)); m.z.push_back(parse_dbl(t[4], ok));
                m.q.push_back(parse_dbl(t[8], ok)); m.r.push_back(e->radius); m.anum.push_back(e->anum);
                m.aname.push_back(t[1]); m.atype.push_back(t[5]);
            }
        }
        size_t b0 = a0 + 1 + (size_t)n_atoms;
        while (ok && b0 < end && lines[b0] != "@<TRIPOS>BOND") b0++;
        ok = ok && b0 + 1 + (size_t)n_bonds <= end;
        std::vector<int> bs, bd;
        std::vector<double> bo;
        for (int b = 0; b < n_bonds && ok; b++) {
            auto t = split_ws(lines[b0 + 1 + b]);          // id src dst type
            if (t.size() < 4) { ok = false; break; }
            int s = parse_int(t[1], ok) - 1, d = parse_int(t[2], ok) - 1;
            if (!ok || s < 0 || d < 0 || s >= n_atoms || d >= n_atoms) { ok = false; break; }
            double order;
            if (t[3] == "1" || t[3] == "am") order = 1.0;
            else if (t[3] == "2") order = 2.0;
            else if (t[3] == "3") order = 3.0;
            else if (t[3] == "ar") order = 1.5;
            else { set_error("mmo_molfile_read_mol2: bond type %s in %s", t[3].c_str(), m.name.c_str()); ok = false; break; }
            if (keep[s] < 0 || keep[d] < 0) continue;
            bs.push_back(keep[s]); bd.push_back(keep[d]); bo.push_back(order);
            m.b_src.push_back(keep[s]); m.b_dst.push_back(keep[d]); m.b_typ.push_back(t[3]);
        }
        if (ok && m.n() > 0 && (!analyse || analyse_graph(m, bs, bd, bo))) f->mols.push_back(std::move(m));
        else f->n_skipped++;
        p = end;
        if (!analyse) break;
    }
    f->assign_types();
    *out = f;
    return MMO_OK;
}

int mmo_molfile_read_mol2(const char *path, mmo_molfile **out) try { return read_mol2(path, true, out); } MMO_CATCH_ALL
int mmo_molfile_read_mol2_atoms(const char *path, mmo_molfile **out) try { return read_mol2(path, false, out); } MMO_CATCH_ALL

int mmo_molfile_read_pqrs(const char *path, int is_receptor, mmo_molfile **out) try {
    MMO_REQUIRE(path && out, "mmo_molfile_read_pqrs: null pointer");
    *out = nullptr;
    std::ifstream in(path);
    MMO_REQUIRE(in.good(), "mmo_molfile_read_pqrs: cannot open %s", path);
    std::vector<std::string> lines;
    for (std::string l; std::getline(in, l);) lines.push_back(rstrip(l));
    mmo_molfile *f = new mmo_molfile();
    size_t p = 0;
    auto fail = [&](const char *what) { set_error("mmo_molfile_read_pqrs: %s (%s line %zu)", what, path, p + 1); delete f; return MMO_EINVAL; };
    while (p < lines.size() && !strip(lines[p]).empty()) {
        Molecule m;
        m.is_ligand = !is_receptor;
        // header  N:R:name  (ligand, the name may hold ':')  or  N:name  (receptor)
        const std::string &h = lines[p];
        bool ok = true;
        size_t c1 = h.find(':');
        if (c1 == std::string::npos) return fail("bad header");
        const int n = parse_int(h.substr(0, c1), ok);
        int nrb = 0;
        if (is_receptor) {
            m.name = h.substr(c1 + 1);
        } else {
            size_t c2 = h.find(':', c1 + 1);
            if (c2 == std::string::npos) return fail("bad ligand header");
            nrb = parse_int(h.substr(c1 + 1, c2 - c1 - 1), ok);
            m.name = h.substr(c2 + 1);
        }
        if (!ok || n <= 0 || nrb < 0 || p + 1 + (size_t)n > lines.size()) return fail("bad header counts");
        p++;
        for (int i = 0; i < n; i++, p++) {
            auto t = split_ws(lines[p]);
            if (t.size() < 6) return fail("bad atom line");
            const Elt *e = elt_by_sym(t[5]);
            if (!e) return fail("unknown element symbol");
            m.x.push_back(parse_dbl(t[0], ok)); m.y.push_back(parse_dbl(t[1], ok)); m.z.push_back(parse_dbl(t[2], ok));
            m.q.push_back(parse_dbl(t[3], ok)); m.r.push_back(parse_dbl(t[4], ok)); m.anum.push_back(e->anum);
            if (!ok) return fail("bad number in atom line");
        }
        if (!is_receptor) {
            if (p + (size_t)nrb + (size_t)n > lines.size()) return fail("truncated ligand block");
            for (int b = 0; b < nrb; b++, p++) {            // src/pqrs.ml:39-61
                const std::string &l = lines[p];
                size_t eq = l.find('='), dash = l.find('-');
                if (eq == std::string::npos || dash == std::string::npos || dash > eq) return fail("bad rotatable bond line");
                const int s = parse_int(l.substr(0, dash), ok), d = parse_int(l.substr(dash + 1, eq - dash - 1), ok);
                auto t = split_ws(l.substr(eq + 1));
                if (!ok || (int)t.size() != n || s < 0 || d < 0 || s >= n || d >= n) return fail("bad rotatable bond line");
                std::vector<uint8_t> flags(n);
                for (int i = 0; i < n; i++) flags[i] = t[i] == "1";
                if (flags[s] == flags[d]) return fail("rotatable bond with both ends on one side");
                m.rb_src.push_back(s); m.rb_dst.push_back(d);
                m.rb_left.push_back(flags[d] ? s : d);
                m.rb_right.push_back(flags[d] ? d : s);
                m.rb_flags.push_back(flags);
            }
            m.dists.assign((size_t)n * n, 0);
            for (int i = 0; i < n; i++, p++) {              // src/pqrs.ml:63-77
                auto t = split_ws(lines[p]);
                if ((int)t.size() != n) return fail("bad distance matrix row");
                for (int j = 0; j < n; j++) m.dists[i + (size_t)j * n] = parse_int(t[j], ok);
                if (!ok) return fail("bad distance matrix entry");
            }
        }
        f->mols.push_back(std::move(m));
        if (is_receptor) break;
    }
    for (size_t a = 0; a < f->mols.size(); a++)              // mol.ml:428-438
        for (size_t b = a + 1; b < f->mols.size(); b++)
            if (f->mols[a].name == f->mols[b].name) { set_error("mmo_molfile_read_pqrs: duplicate molecule name %s", f->mols[a].name.c_str()); delete f; return MMO_EINVAL; }
    f->assign_types();
    *out = f;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_molfile_count(const mmo_molfile *f, int32_t *n_mols, int32_t *n_skipped) try {
    MMO_REQUIRE(f && n_mols, "mmo_molfile_count: null pointer");
    *n_mols = (int32_t)f->mols.size();
    if (n_skipped) *n_skipped = f->n_skipped;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_molfile_shape(const mmo_molfile *f, int32_t k, int32_t *n_atoms, int32_t *n_rbonds, int32_t *rg_total,
                      char *name, int32_t name_cap) try {
    MMO_REQUIRE(f && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_shape: molecule index out of range");
    const Molecule &m = f->mols[k];
    if (n_atoms) *n_atoms = m.n();
    if (n_rbonds) *n_rbonds = (int32_t)m.rb_left.size();
    if (rg_total) {
        int tot = 0;
        for (size_t b = 0; b < m.rb_flags.size(); b++)
            for (int i = 0; i < m.n(); i++) tot += m.rb_flags[b][i] && i != m.rb_right[b];
        *rg_total = tot;
    }
    if (name && name_cap > 0) { strncpy(name, m.name.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_molfile_get(const mmo_molfile *f, int32_t k, double *xs, double *ys, double *zs, double *q, double *r,
                    int32_t *anum, int32_t *typ, int32_t *dists, int32_t *rb_left, int32_t *rb_right,
                    int32_t *rg_off, int32_t *rg_idx) try {
    MMO_REQUIRE(f && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_get: molecule index out of range");
    const Molecule &m = f->mols[k];
    const int n = m.n();
    if (xs) memcpy(xs, m.x.data(), n * sizeof(double));
    if (ys) memcpy(ys, m.y.data(), n * sizeof(double));
    if (zs) memcpy(zs, m.z.data(), n * sizeof(double));
    if (q) memcpy(q, m.q.data(), n * sizeof(double));
    if (r) memcpy(r, m.r.data(), n * sizeof(double));
    if (anum) memcpy(anum, m.anum.data(), n * sizeof(int32_t));
    if (typ) memcpy(typ, m.typ.data(), n * sizeof(int32_t));
    if (dists && !m.dists.empty()) memcpy(dists, m.dists.data(), (size_t)n * n * sizeof(int32_t));
    const int nrb = (int)m.rb_left.size();
    if (rb_left) memcpy(rb_left, m.rb_left.data(), nrb * sizeof(int32_t));
    if (rb_right) memcpy(rb_right, m.rb_right.data(), nrb * sizeof(int32_t));
    if (rg_off) {
        int tot = 0;
        for (int b = 0; b < nrb; b++) {              // movable atoms, axis tip excluded (pqrs.ml:80-87)
            rg_off[b] = tot;
            for (int i = 0; i < n; i++)
                if (m.rb_flags[b][i] && i != m.rb_right[b]) { if (rg_idx) rg_idx[tot] = i; tot++; }
        }
        rg_off[nrb] = tot;
    }
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_molfile_types(const mmo_molfile *f, int32_t *n_types, int32_t *type_anum, double *type_q) try {
    MMO_REQUIRE(f && n_types, "mmo_molfile_types: null pointer");
    *n_types = (int32_t)f->type_anum.size();
    if (type_anum) memcpy(type_anum, f->type_anum.data(), f->type_anum.size() * sizeof(int32_t));
    if (type_q) memcpy(type_q, f->type_q.data(), f->type_q.size() * sizeof(double));
    return MMO_OK;
} MMO_CATCH_ALL

// lds --less-charges (src/lds.ml:1887-1894): Mol.reduce_partial_charges_precision on every ligand (src/mol.ml:256-260,
// Utls.reduce_precision src/utls.ml:127-132: two decimals, rounded away from zero), BEFORE the FF types are assigned
int mmo_molfile_reduce_charges(mmo_molfile *f) try {
    MMO_REQUIRE(f != nullptr, "mmo_molfile_reduce_charges: null pointer");
    for (Molecule &m : f->mols)
        for (double &q : m.q) q = (double)(long long)(q * 100.0 + (q >= 0.0 ? 0.5 : -0.5)) / 100.0;
    f->assign_types();
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_molfile_write_pqrs(const mmo_molfile *f, const char *path) try {
    MMO_REQUIRE(f && path, "mmo_molfile_write_pqrs: null pointer");
    FILE *o = fopen(path, "w");
    MMO_REQUIRE(o != nullptr, "mmo_molfile_write_pqrs: cannot create %s", path);
    for (const Molecule &m : f->mols) {
        const int n = m.n();
        if (m.is_ligand) fprintf(o, "%d:%d:%s\n", n, (int)m.rb_left.size(), m.name.c_str());
        else fprintf(o, "%d:%s\n", n, m.name.c_str());
        for (int i = 0; i < n; i++) {
            const Elt *e = elt_by_anum(m.anum[i]);
            fprintf(o, "%g %g %g %g %g %s\n", m.x[i], m.y[i], m.z[i], m.q[i], m.r[i], e ? e->sym : "X");
        }
        if (!m.is_ligand) continue;
        for (size_t b = 0; b < m.rb_flags.size(); b++) {
            fprintf(o, "%d-%d=", m.rb_src[b], m.rb_dst[b]);
            for (int i = 0; i < n; i++) fputs(m.rb_flags[b][i] ? " 1" : " 0", o);
            fputc('\n', o);
        }
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < n; j++) fprintf(o, j ? " %d" : "%d", m.dists[i + (size_t)j * n]);
            fputc('\n', o);
        }
    }
    fclose(o);
    return MMO_OK;
} MMO_CATCH_ALL

// straight to the device handle: Mol.translate_to lig V3.origin (lds.ml:44-52) when `centered`, with the
// Kahan-averaged centre (Batteries A.favg as restated in the oracle)
int mmo_molfile_ligand(const mmo_molfile *f, int32_t k, int centered, mmo_ligand **out) try {
    MMO_REQUIRE(f && out && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_ligand: molecule index out of range");
    const Molecule &m = f->mols[k];
    const int n = m.n(), nrb = (int)m.rb_left.size();
    std::vector<double> x = m.x, y = m.y, z = m.z;
    if (centered) {
        auto favg = [n](const std::vector<double> &a) {
            double s = 0.0, c = 0.0;
            for (int i = 0; i < n; i++) { double yy = a[i] - c, t = s + yy; c = (t - s) - yy; s = t; }
            return s / n;
        };
        const double cx = favg(m.x), cy = favg(m.y), cz = favg(m.z);
        for (int i = 0; i < n; i++) { x[i] = m.x[i] + (0.0 - cx); y[i] = m.y[i] + (0.0 - cy); z[i] = m.z[i] + (0.0 - cz); }
    }
    std::vector<int32_t> off(nrb + 1), idx;
    {
        int32_t tot = 0;
        MMO_TRY(mmo_molfile_shape(f, k, nullptr, nullptr, &tot, nullptr, 0));
        idx.resize(std::max(1, tot));
        MMO_TRY(mmo_molfile_get(f, k, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, off.data(), idx.data()));
    }
    return mmo_ligand_create(n, x.data(), y.data(), z.data(), m.q.data(), m.r.data(), m.anum.data(), m.typ.data(),
                             m.dists.empty() ? nullptr : m.dists.data(), nrb, m.rb_left.data(), m.rb_right.data(),
                             off.data(), idx.data(), out);
} MMO_CATCH_ALL

// Mol2.output_one (src/mol2.ml:326-343, line formats 184-190, 209-210) of Mol.update_mol2 mol2 m (src/mol.ml:544-552):
// molecule k written n_copies times with its coordinates replaced by copy c of xs/ys/zs ([n_copies][n_atoms]; NULL =
// the file's own).  What lig_rot_sample (one block per rotation) and place_ligand (one block) emit.
int mmo_molfile_write_mol2(const mmo_molfile *f, int32_t k, int32_t n_copies, const double *xs, const double *ys,
                           const double *zs, const char *path, int append) try {
    MMO_REQUIRE(f && path && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_write_mol2: molecule index out of range");
    const Molecule &m = f->mols[k];
    const int n = m.n();
    MMO_REQUIRE((int)m.aname.size() == n, "mmo_molfile_write_mol2: molecule %s was not read from a mol2 file", m.name.c_str());
    MMO_REQUIRE(n_copies >= 0 && ((xs && ys && zs) || (!xs && !ys && !zs)), "mmo_molfile_write_mol2: bad arguments");
    FILE *o = fopen(path, append ? "a" : "w");
    MMO_REQUIRE(o != nullptr, "mmo_molfile_write_mol2: cannot create %s", path);
    for (int c = 0; c < n_copies; c++) {
        fprintf(o, "@<TRIPOS>MOLECULE\n%s\n%5d%6d%6d%6d%6d\nSMALL\nUSER_CHARGES\n\n@<TRIPOS>ATOM\n", m.name.c_str(), n,
                (int)m.b_src.size(), 0, 0, 0);
        for (int i = 0; i < n; i++) {
            const double x = xs ? xs[(size_t)c * n + i] : m.x[i], y = ys ? ys[(size_t)c * n + i] : m.y[i];
            const double z = zs ? zs[(size_t)c * n + i] : m.z[i];
            fprintf(o, "%7d %-8s%10.4f%10.4f%10.4f %-8s  1 <0>     %10.4f\n", i + 1, m.aname[i].c_str(), x, y, z,
                    m.atype[i].c_str(), m.q[i]);
        }
        fputs("@<TRIPOS>BOND\n", o);
        for (size_t b = 0; b < m.b_src.size(); b++)
            fprintf(o, "%6d%5d%5d %s\n", (int)b + 1, m.b_src[b] + 1, m.b_dst[b] + 1, m.b_typ[b].c_str());
    }
    const bool bad = ferror(o) != 0;
    MMO_REQUIRE(fclose(o) == 0 && !bad, "mmo_molfile_write_mol2: write error on %s", path);
    return MMO_OK;
} MMO_CATCH_ALL

namespace {
struct HostView {
    std::vector<int32_t> off, idx;
    mmo::HostLig h;
};
int host_view(const mmo_molfile *f, int32_t k, const std::vector<double> &x, const std::vector<double> &y,
              const std::vector<double> &z, HostView &v) {
    const Molecule &m = f->mols[k];
    const int nrb = (int)m.rb_left.size();
    int32_t tot = 0;
    MMO_TRY(mmo_molfile_shape(f, k, nullptr, nullptr, &tot, nullptr, 0));
    v.off.resize(nrb + 1);
    v.idx.resize(std::max(1, tot));
    MMO_TRY(mmo_molfile_get(f, k, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                            v.off.data(), v.idx.data()));
    v.h = {m.n(), x.data(), y.data(), z.data(), nrb, m.rb_left.data(), m.rb_right.data(), v.off.data(), v.idx.data()};
    return MMO_OK;
}
}  // namespace

// the body of the lig_rot_sample tool (src/lig_rot_sample.ml:23-45) on molecule k of the file, host only:
// Mol.center_rotate_translate_copy mol rot (Mol.get_center mol) for every rotation
int mmo_molfile_rotated_copies(const mmo_molfile *f, int32_t k, int32_t n, const double *rot9, double *out_xs,
                               double *out_ys, double *out_zs) try {
    MMO_REQUIRE(f && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_rotated_copies: molecule index out of range");
    MMO_REQUIRE(n >= 0 && (n == 0 || (rot9 && out_xs && out_ys && out_zs)), "mmo_molfile_rotated_copies: bad arguments");
    const Molecule &m = f->mols[k];
    const double center[3] = {favg_host(m.x.data(), m.n()), favg_host(m.y.data(), m.n()), favg_host(m.z.data(), m.n())};
    const mmo::HostLig h = {m.n(), m.x.data(), m.y.data(), m.z.data(), 0, nullptr, nullptr, nullptr, nullptr};
    rotated_copies_host(h, center, n, rot9, out_xs, out_ys, out_zs);
    return MMO_OK;
} MMO_CATCH_ALL

// the body of the place_ligand tool (src/place_ligand.ml:36-59) on molecule k, host only: Mol.center, then
// Optim.apply_config centered_lig (x y z a b g [rbond angles]) (src/optim.ml:64-80)
int mmo_molfile_apply_config(const mmo_molfile *f, int32_t k, const double *config, int32_t n_config, double *out_xs,
                             double *out_ys, double *out_zs, int32_t *too_long) try {
    MMO_REQUIRE(f && k >= 0 && k < (int32_t)f->mols.size(), "mmo_molfile_apply_config: molecule index out of range");
    MMO_REQUIRE(config && out_xs && out_ys && out_zs, "mmo_molfile_apply_config: null argument");
    const Molecule &m = f->mols[k];
    const int n = m.n();
    const double cx = favg_host(m.x.data(), n), cy = favg_host(m.y.data(), n), cz = favg_host(m.z.data(), n);
    std::vector<double> x(n), y(n), z(n);
    for (int i = 0; i < n; i++) { x[i] = m.x[i] + (0.0 - cx); y[i] = m.y[i] + (0.0 - cy); z[i] = m.z[i] + (0.0 - cz); }
    HostView v;
    MMO_TRY(host_view(f, k, x, y, z, v));
    return apply_config_host(v.h, config, n_config, out_xs, out_ys, out_zs, too_long);
} MMO_CATCH_ALL

int mmo_molfile_destroy(mmo_molfile *f) try {
    delete f;
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
