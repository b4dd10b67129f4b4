// molecules.cu -- receptor / ligand handles: host-side preparation of the HBM layouts.
//   receptor: original-order fp64 SoA (strict kernels) + fp32 k-d groups of kBlob atoms
//             with bounding boxes (fast kernel) + close-contact voxel lists (fp64 correction pass)
//   ligand  : template conformer, fp32 vdW factors, interacting-pair list, rotatable bonds
// Reference data model: src/mol.ml:17-35 (Mol.t), src/UFF.ml:10-51, src/ptable.ml:41-54.
#include "common.cuh"
#include <math.h>
#include <algorithm>
#include <numeric>

namespace mmo {

// src/UFF.ml:10-22
const int kEltAnum[kNumElt] = {0, 1, 6, 7, 8, 9, 12, 15, 16, 17, 35, 53};
const double kEltXi[kNumElt] = {0.0, 2.886, 3.851, 3.660, 3.500, 3.364, 3.021, 4.147, 4.035, 3.947, 4.189, 4.500};
const double kEltDi[kNumElt] = {0.0, 0.044, 0.105, 0.069, 0.060, 0.050, 0.111, 0.305, 0.274, 0.227, 0.251, 0.339};

int elt_index(int anum) {
    for (int i = 0; i < kNumElt; i++)
        if (kEltAnum[i] == anum) return i;
    return kEltUnsupported;
}
// src/ptable.ml:41-54
double vdw_radius(int anum) {
    switch (anum) {
    case 1: return 1.2; case 6: return 1.7; case 7: return 1.6; case 8: return 1.55;
    case 9: return 1.5; case 12: return 2.2; case 15: return 1.95; case 16: return 1.8;
    case 17: return 1.8; case 35: return 1.9; case 53: return 2.1;
    default: return NAN;
    }
}

static void kd_rec(const double *const c[3], int *idx, int n, int leaf) {
    if (n <= leaf) return;
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
        lo[d] = hi[d] = c[d][idx[0]];
        for (int k = 1; k < n; k++) { lo[d] = std::min(lo[d], c[d][idx[k]]); hi[d] = std::max(hi[d], c[d][idx[k]]); }
    }
    int ax = 0;
    if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
    if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
    // left part: a multiple of `leaf`, as close to n/2 as possible -> every leaf but the last is full
    int nl = ((n / 2 + leaf - 1) / leaf) * leaf;
    if (nl >= n) nl = n - leaf;
    const double *v = c[ax];
    std::nth_element(idx, idx + nl, idx + n, [v](int a, int b) { return v[a] < v[b] || (v[a] == v[b] && a < b); });
    kd_rec(c, idx, nl, leaf);
    kd_rec(c, idx + nl, n - nl, leaf);
}

void kd_order(int n, const double *x, const double *y, const double *z, int leaf, std::vector<int> &order) {
    order.resize(n);
    std::iota(order.begin(), order.end(), 0);
    const double *c[3] = {x, y, z};
    if (n > 0) kd_rec(c, order.data(), n, leaf);
}

// vdW factors of the A/B form: d_ij*(p6^2 - 2 p6) = (A_i A_j) s^6 - (B_i B_j) s^3 with s = 1/r^2,
// A = sqrt(D) x^6, B = sqrt(2 D) x^3 (x_ij = sqrt(x_i x_j), d_ij = sqrt(D_i D_j), src/UFF.ml:44-50)
void vdw_factors(int elt, float *A, float *B) {
    if (elt >= kNumElt) { *A = NAN; *B = NAN; return; }
    double x = kEltXi[elt], d = kEltDi[elt];
    double x3 = x * x * x;
    *A = (float)(sqrt(d) * x3 * x3);
    *B = (float)(sqrt(2.0 * d) * x3);
}

}  // namespace mmo

using namespace mmo;

extern "C" {

int mmo_receptor_create(int32_t n, const double *xs, const double *ys, const double *zs,
                        const double *q, const int32_t *anum, mmo_receptor **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(out != nullptr, "mmo_receptor_create: null output pointer");
    *out = nullptr;
    MMO_REQUIRE(n >= 0, "mmo_receptor_create: negative atom count");
    MMO_REQUIRE(n == 0 || (xs && ys && zs && q && anum), "mmo_receptor_create: null input array");
    mmo_receptor *r = new mmo_receptor();
    r->n = n;
    r->hx.assign(xs, xs + n); r->hy.assign(ys, ys + n); r->hz.assign(zs, zs + n);
    r->hq.assign(q, q + n); r->hanum.assign(anum, anum + n);
    std::vector<int32_t> elt(n);
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    r->x_max = 0.0;
    for (int i = 0; i < n; i++) {
        elt[i] = elt_index(anum[i]);
        double p[3] = {xs[i], ys[i], zs[i]};
        for (int d = 0; d < 3; d++) {
            if (i == 0 || p[d] < lo[d]) lo[d] = p[d];
            if (i == 0 || p[d] > hi[d]) hi[d] = p[d];
        }
        if (elt[i] < kNumElt && kEltXi[elt[i]] > r->x_max) r->x_max = kEltXi[elt[i]];
    }
    for (int d = 0; d < 3; d++) { r->origin[d] = 0.5 * (lo[d] + hi[d]); r->bb_lo[d] = lo[d]; r->bb_hi[d] = hi[d]; }

    // ---- k-d leaves -> groups of kBlob spatially close atoms (first level of distance culling); the
    //      element of every slot is kept beside it, its vdW factors come from a 13-entry table
    std::vector<int> order;
    kd_order(n, xs, ys, zs, kBlob, order);
    r->n_blobs = (n + kBlob - 1) / kBlob;
    r->n_pad = r->n_blobs * kBlob;
    // one more group than there are leaves: group n_blobs holds far-away, charge-free atoms (padding of the near lists)
    std::vector<float4> xyzq((size_t)r->n_pad + kBlob, make_float4(kFarAway, kFarAway, kFarAway, 0.f));
    std::vector<uint8_t> gelt((size_t)r->n_pad + kBlob, 0);
    std::vector<float4> box((size_t)std::max(1, r->n_blobs) * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int b = 0; b < r->n_blobs; b++) {
        float blo[3] = {3e38f, 3e38f, 3e38f}, bhi[3] = {-3e38f, -3e38f, -3e38f};
        for (int s = 0; s < kBlob; s++) {
            int k = b * kBlob + s;
            if (k >= n) break;        // padding slots stay far away and charge-free: never listed
            int i = order[k];
            float4 v;
            v.x = (float)(xs[i] - r->origin[0]);
            v.y = (float)(ys[i] - r->origin[1]);
            v.z = (float)(zs[i] - r->origin[2]);
            v.w = (float)(kElecWeight * q[i]);
            xyzq[k] = v;
            gelt[k] = (uint8_t)elt[i];
            float p[3] = {v.x, v.y, v.z};
            for (int d = 0; d < 3; d++) { blo[d] = fminf(blo[d], p[d]); bhi[d] = fmaxf(bhi[d], p[d]); }
        }
        box[(size_t)b * 2] = make_float4(blo[0], blo[1], blo[2], 0.f);
        box[(size_t)b * 2 + 1] = make_float4(bhi[0], bhi[1], bhi[2], 0.f);
    }

    // super-groups: 32 consecutive leaves of the balanced k-d order are one subtree (or two neighbouring ones), i.e.
    // spatially compact: first stage of the item kernel's group culling
    r->n_sup = (r->n_blobs + 31) / 32;
    std::vector<float4> sup((size_t)std::max(1, r->n_sup) * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int sg = 0; sg < r->n_sup; sg++) {
        float4 lo4 = make_float4(3e38f, 3e38f, 3e38f, 0.f), hi4 = make_float4(-3e38f, -3e38f, -3e38f, 0.f);
        for (int b = sg * 32; b < std::min(r->n_blobs, sg * 32 + 32); b++) {
            const float4 bl = box[(size_t)b * 2], bh = box[(size_t)b * 2 + 1];
            lo4.x = fminf(lo4.x, bl.x); lo4.y = fminf(lo4.y, bl.y); lo4.z = fminf(lo4.z, bl.z);
            hi4.x = fmaxf(hi4.x, bh.x); hi4.y = fmaxf(hi4.y, bh.y); hi4.z = fmaxf(hi4.z, bh.z);
        }
        sup[(size_t)sg * 2] = lo4; sup[(size_t)sg * 2 + 1] = hi4;
    }

    // ---- close-contact voxel lists: atoms within r_list of any point of the voxel (conservative)
    // 1 A voxels (fewer candidates per lookup) while the table stays small, 2 A otherwise
    {
        const double rl = sqrt(std::max(r->x_max, 1.0) * 4.5 / kTau);
        double vol = 1.0;
        for (int d = 0; d < 3; d++) vol *= (hi[d] - lo[d] + 2.0 * rl);
        r->vox_edge = (n > 0 && vol <= 2.0e6) ? 1.0 : 2.0;
    }
    const double g = r->vox_edge;
    // 4.5 = largest x_i of src/UFF.ml:22; the slack covers the fp32 position the lookup is made with
    const double r_list = sqrt(std::max(r->x_max, 1.0) * 4.5 / kTau) + 1e-3;
    const double reach = r_list + 0.5 * g * sqrt(3.0);
    for (int d = 0; d < 3; d++) {
        r->vox_lo[d] = lo[d] - r_list - 1e-6;
        r->vox_dim[d] = std::max(1, (int)ceil((hi[d] + r_list + 1e-6 - r->vox_lo[d]) / g));
    }
    size_t nvox = (size_t)r->vox_dim[0] * r->vox_dim[1] * r->vox_dim[2];
    if (n == 0) nvox = 1;
    std::vector<int32_t> cnt(nvox + 1, 0);
    auto for_each_voxel = [&](int i, auto &&fn) {
        double p[3] = {xs[i], ys[i], zs[i]};
        int a0[3], a1[3];
        for (int d = 0; d < 3; d++) {
            a0[d] = std::max(0, (int)floor((p[d] - reach - r->vox_lo[d]) / g));
            a1[d] = std::min(r->vox_dim[d] - 1, (int)floor((p[d] + reach - r->vox_lo[d]) / g));
        }
        for (int k = a0[2]; k <= a1[2]; k++)
            for (int j = a0[1]; j <= a1[1]; j++)
                for (int ii = a0[0]; ii <= a1[0]; ii++) {
                    double c[3] = {r->vox_lo[0] + (ii + 0.5) * g, r->vox_lo[1] + (j + 0.5) * g, r->vox_lo[2] + (k + 0.5) * g};
                    double d2 = (c[0] - p[0]) * (c[0] - p[0]) + (c[1] - p[1]) * (c[1] - p[1]) + (c[2] - p[2]) * (c[2] - p[2]);
                    if (d2 <= reach * reach) fn((size_t)ii + (size_t)j * r->vox_dim[0] + (size_t)k * r->vox_dim[0] * r->vox_dim[1]);
                }
    };
    for (int i = 0; i < n; i++) for_each_voxel(i, [&](size_t v) { cnt[v + 1]++; });
    for (size_t v = 0; v < nvox; v++) cnt[v + 1] += cnt[v];
    std::vector<int32_t> idx(std::max<size_t>(1, (size_t)cnt[nvox]));
    std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
    // ascending atom index per voxel; the atom's compact element index rides in the top byte (one load less per pair in
    // the close-contact pass)
    MMO_REQUIRE(n < (1 << 24), "mmo_receptor_create: more than 16 M atoms");
    for (int i = 0; i < n; i++) for_each_voxel(i, [&](size_t v) { idx[fill[v]++] = i | (elt[i] << 24); });

    std::vector<double4> xyzq64(std::max(1, n));
    for (int i = 0; i < n; i++) xyzq64[i] = make_double4(xs[i], ys[i], zs[i], q[i]);
    int rc = MMO_OK;
    do {
        if ((rc = r->xyzq64.upload(xyzq64))) break;
        {
            std::vector<float4> v32(std::max(1, n));
            for (int i = 0; i < n; i++) v32[i] = make_float4((float)(xs[i] - r->vox_lo[0]), (float)(ys[i] - r->vox_lo[1]), (float)(zs[i] - r->vox_lo[2]), 0.f);
            if ((rc = r->xyz32v.upload(v32))) break;
        }
        if ((rc = r->x.upload(r->hx)) || (rc = r->y.upload(r->hy)) || (rc = r->z.upload(r->hz)) ||
            (rc = r->q.upload(r->hq)) || (rc = r->elt.upload(elt)) || (rc = r->xyzq.upload(xyzq)) || (rc = r->gelt.upload(gelt)) ||
            (rc = r->blob_box.upload(box)) || (rc = r->sup_box.upload(sup)) || (rc = r->vox_off.upload(cnt)) ||
            (rc = r->vox_idx.upload(idx)))
            break;
    } while (0);
    if (rc != MMO_OK) { delete r; return rc; }
    *out = r;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_receptor_destroy(mmo_receptor *rec) try {
    delete rec;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_ligand_create(int32_t n, const double *xs, const double *ys, const double *zs,
                      const double *q, const double *r, const int32_t *anum, const int32_t *typ,
                      const int32_t *dists,
                      int32_t n_rbonds, const int32_t *rb_left, const int32_t *rb_right,
                      const int32_t *rg_off, const int32_t *rg_idx, mmo_ligand **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(out != nullptr, "mmo_ligand_create: null output pointer");
    *out = nullptr;
    MMO_REQUIRE(n > 0, "mmo_ligand_create: a ligand needs at least one atom");
    MMO_REQUIRE(xs && ys && zs && q && anum, "mmo_ligand_create: null input array");
    MMO_REQUIRE(n_rbonds >= 0, "mmo_ligand_create: negative rotatable-bond count");
    MMO_REQUIRE(n_rbonds == 0 || (rb_left && rb_right && rg_off && rg_idx),
                "mmo_ligand_create: rotatable bonds announced but arrays are null");
    // the CSR offsets are lengths of caller memory: validated before anything is copied with them
    for (int b = 0; b < n_rbonds; b++)
        MMO_REQUIRE(rg_off[0] == 0 && rg_off[b] >= 0 && rg_off[b] <= rg_off[b + 1] && rg_off[b + 1] <= n_rbonds * n,
                    "mmo_ligand_create: rotatable group offsets are not a monotonic CSR (bond %d)", b);
    mmo_ligand *l = new mmo_ligand();
    l->n = n;
    l->hx.assign(xs, xs + n); l->hy.assign(ys, ys + n); l->hz.assign(zs, zs + n);
    l->hq.assign(q, q + n); l->hanum.assign(anum, anum + n);
    if (r) { l->hr.assign(r, r + n); l->has_r = true; }
    if (typ) { l->htyp.assign(typ, typ + n); l->has_typ = true; }
    else l->htyp.assign(n, 0);
    std::vector<int32_t> pi, pj;
    if (dists) {
        l->hdists.assign(dists, dists + (size_t)n * n);
        l->has_dists = true;
        for (int i = 0; i < n - 1; i++)          // src/mol.ml:885-891 loop order
            for (int j = i + 1; j < n; j++)
                if (dists[i + (size_t)j * n] >= 3) { pi.push_back(i); pj.push_back(j); }   // mol.ml:203-208
    }
    l->n_pairs = (int)pi.size();
    l->h_pair_i = pi; l->h_pair_j = pj;
    l->n_rbonds = n_rbonds;
    if (n_rbonds > 0) {
        l->rb_left.assign(rb_left, rb_left + n_rbonds);
        l->rb_right.assign(rb_right, rb_right + n_rbonds);
        l->rg_off.assign(rg_off, rg_off + n_rbonds + 1);
        l->rg_idx.assign(rg_idx, rg_idx + rg_off[n_rbonds]);
        for (int b = 0; b < n_rbonds; b++) {
            bool ok = rb_left[b] >= 0 && rb_left[b] < n && rb_right[b] >= 0 && rb_right[b] < n &&
                      rg_off[b] <= rg_off[b + 1];
            if (!ok) { delete l; set_error("mmo_ligand_create: rotatable bond %d is malformed", b); return MMO_EINVAL; }
        }
        for (int32_t v : l->rg_idx)
            if (v < 0 || v >= n) { delete l; set_error("mmo_ligand_create: rotatable group index out of range"); return MMO_EINVAL; }
    } else {
        l->rg_off.assign(1, 0);
    }
    std::vector<int32_t> elt(n);
    l->x_max = 0.0;
    for (int j = 0; j < n; j++) {
        elt[j] = elt_index(anum[j]);
        if (elt[j] < kNumElt && kEltXi[elt[j]] > l->x_max) l->x_max = kEltXi[elt[j]];
    }
    // fast-path order: k-d leaves of 8 atoms = the register chunks of the fp32 kernel
    std::vector<int> forder;
    kd_order(n, xs, ys, zs, 8, forder);
    l->n_fast = ((n + 7) / 8) * 8;
    std::vector<float4> fp(l->n_fast, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<double> fx(l->n_fast, 0.0), fy(l->n_fast, 0.0), fz(l->n_fast, 0.0);
    for (int k = 0; k < n; k++) {
        int j = forder[k];
        float A, B;
        vdw_factors(elt[j], &A, &B);
        // .w = the atom's UFF distance x_j (>= 1: also marks a real atom; 0 = padding): the close-contact threshold of
        // the fp32 kernels is per ligand atom, H_j = hscale * x_j with hscale = x_max(receptor) / kTau
        fp[k] = make_float4(A, B, (float)q[j], (float)std::max(elt[j] < kNumElt ? kEltXi[elt[j]] : 1.0, 1.0));
        fx[k] = xs[j]; fy[k] = ys[j]; fz[k] = zs[j];
    }
    int rc = MMO_OK;
    do {
        if ((rc = l->x.upload(l->hx)) || (rc = l->y.upload(l->hy)) || (rc = l->z.upload(l->hz)) ||
            (rc = l->q.upload(l->hq)) || (rc = l->elt.upload(elt)) || (rc = l->typ.upload(l->htyp)) ||
            (rc = l->fparam.upload(fp)) || (rc = l->fx.upload(fx)) || (rc = l->fy.upload(fy)) || (rc = l->fz.upload(fz)) || (rc = l->forder.upload(std::vector<int32_t>(forder.begin(), forder.end()))) || (rc = l->pair_i.upload(pi)) || (rc = l->pair_j.upload(pj)) ||
            (rc = l->d_rb_left.upload(l->rb_left)) || (rc = l->d_rb_right.upload(l->rb_right)) ||
            (rc = l->d_rg_off.upload(l->rg_off)) || (rc = l->d_rg_idx.upload(l->rg_idx)))
            break;
    } while (0);
    if (rc != MMO_OK) { delete l; return rc; }
    *out = l;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_ligand_destroy(mmo_ligand *lig) try {
    delete lig;
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
