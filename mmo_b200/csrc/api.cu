// api.cu -- extern "C" entry points that move host buffers to HBM, launch, and copy results back.
// Each function names the reference function it replaces in include/mmo_b200.h.
#include "common.cuh"
#include <ctype.h>
#include <stdlib.h>
#include <algorithm>
#include <math.h>
#include <string.h>
#include <string>

using namespace mmo;

namespace {

int check_variant_prec(int variant, int prec) {
    MMO_REQUIRE(variant == MMO_VARIANT_GLOBAL || variant == MMO_VARIANT_SHIFTED, "bad variant %d", variant);
    MMO_REQUIRE(prec == MMO_PREC_FP32 || prec == MMO_PREC_FP64, "bad precision %d", prec);
    return MMO_OK;
}

int run_direct(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec, const PoseSrc &src,
               int64_t n, double *d_out) {
    if (prec == MMO_PREC_FP64) return launch_direct_fp64(rec, lig, variant, src, n, d_out);
    return launch_direct_fp32(rec, lig, variant, src, n, d_out, rt().collect_stats);
}

int d2h_sync(void *host, const void *dev, size_t bytes) {
    MMO_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, rt().stream));
    MMO_CUDA(cudaStreamSynchronize(rt().stream));
    return MMO_OK;
}

// Host arrays of one call -> ONE device buffer (dev[i] = start of array i), and no host synchronisation in between: the
// call's closing d2h_sync orders everything.  Small inputs (single-pose calls: the reference's closures are
// Mol.t -> float) are packed into a pinned staging buffer and travel as one copy.
constexpr size_t kStageBytes = kStageHalf;
int upload_parts(DevBuf<double> &buf, int n, const double *const *host, const size_t *count, const double **dev) {
    Runtime &R = rt();
    size_t total = 0;
    for (int i = 0; i < n; i++) total += count[i];
    MMO_TRY(buf.alloc(total));
    size_t off = 0;
    for (int i = 0; i < n; i++) { dev[i] = buf.p + off; off += count[i]; }
    if (total * sizeof(double) <= kStageBytes) {
        void *stage = nullptr;
        MMO_TRY(stage_buffer(&stage));
        // the previous call that used the staging buffer ended with a stream synchronisation
        off = 0;
        for (int i = 0; i < n; i++) { memcpy((double *)R.stage + off, host[i], count[i] * sizeof(double)); off += count[i]; }
        if (total) MMO_CUDA(cudaMemcpyAsync(buf.p, R.stage, total * sizeof(double), cudaMemcpyHostToDevice, R.stream));
    } else {
        for (int i = 0; i < n; i++)
            if (count[i]) MMO_CUDA(cudaMemcpyAsync((double *)dev[i], host[i], count[i] * sizeof(double), cudaMemcpyHostToDevice, R.stream));
        MMO_CUDA(cudaStreamSynchronize(R.stream));      // pageable sources: the caller may reuse them after we return
    }
    return MMO_OK;
}
int upload_xyz(DevBuf<double> &buf, const double *xs, const double *ys, const double *zs, size_t m, const double **dev) {
    const double *h[3] = {xs, ys, zs};
    const size_t c[3] = {m, m, m};
    return upload_parts(buf, 3, h, c, dev);
}
int upload_rt(DevBuf<double> &buf, const double *rot9, const double *trans3, size_t n_poses, const double **dev) {
    const double *h[2] = {rot9, trans3};
    const size_t c[2] = {n_poses * 9, n_poses * 3};
    return upload_parts(buf, 2, h, c, dev);
}

PoseSrc coords_src(const double *x, const double *y, const double *z) {
    PoseSrc s = {};
    s.kind = 1;
    s.xs = x; s.ys = y; s.zs = z;
    return s;
}
PoseSrc rt_src(const double *rot9, const double *trans3) {
    PoseSrc s = {};
    s.kind = 0;
    s.rot9 = rot9; s.trans3 = trans3;
    return s;
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ direct pair path
int mmo_score_coords(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                     int64_t n_poses, const double *xs, const double *ys, const double *zs, double *out_E) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && lig, "mmo_score_coords: null handle");
    MMO_TRY(check_variant_prec(variant, prec));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_coords: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out_E, "mmo_score_coords: null buffer");
    const size_t m = (size_t)n_poses * lig->n;
    DevBuf<double> dxyz, dE;
    const double *d[3];
    MMO_TRY(upload_xyz(dxyz, xs, ys, zs, m, d));
    MMO_TRY(dE.alloc((size_t)n_poses));
    MMO_TRY(run_direct(rec, lig, variant, prec, coords_src(d[0], d[1], d[2]), n_poses, dE.p));
    return d2h_sync(out_E, dE.p, (size_t)n_poses * sizeof(double));
} MMO_CATCH_ALL

int mmo_score_poses(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                    int64_t n_poses, const double *rot9, const double *trans3, double *out_E) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && lig, "mmo_score_poses: null handle");
    MMO_TRY(check_variant_prec(variant, prec));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_poses: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(rot9 && trans3 && out_E, "mmo_score_poses: null buffer");
    DevBuf<double> drt, dE;
    const double *d[2];
    MMO_TRY(upload_rt(drt, rot9, trans3, (size_t)n_poses, d));
    MMO_TRY(dE.alloc((size_t)n_poses));
    MMO_TRY(run_direct(rec, lig, variant, prec, rt_src(d[0], d[1]), n_poses, dE.p));
    return d2h_sync(out_E, dE.p, (size_t)n_poses * sizeof(double));
} MMO_CATCH_ALL

int mmo_score_poses_dev(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                        int64_t n_poses, const double *d_rot9, const double *d_trans3, double *d_out_E) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && lig, "mmo_score_poses_dev: null handle");
    MMO_TRY(check_variant_prec(variant, prec));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_poses_dev: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(d_rot9 && d_trans3 && d_out_E, "mmo_score_poses_dev: null buffer");
    if (prec == MMO_PREC_FP64) return launch_direct_fp64(rec, lig, variant, rt_src(d_rot9, d_trans3), n_poses, d_out_E);
    return launch_direct_fp32(rec, lig, variant, rt_src(d_rot9, d_trans3), n_poses, d_out_E, rt().collect_stats);
} MMO_CATCH_ALL

int mmo_score_coords_dev(const mmo_receptor *rec, const mmo_ligand *lig, int variant, int prec,
                         int64_t n_poses, const double *d_xs, const double *d_ys, const double *d_zs, double *d_out_E) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && lig, "mmo_score_coords_dev: null handle");
    MMO_TRY(check_variant_prec(variant, prec));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_coords_dev: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(d_xs && d_ys && d_zs && d_out_E, "mmo_score_coords_dev: null buffer");
    if (prec == MMO_PREC_FP64) return launch_direct_fp64(rec, lig, variant, coords_src(d_xs, d_ys, d_zs), n_poses, d_out_E);
    return launch_direct_fp32(rec, lig, variant, coords_src(d_xs, d_ys, d_zs), n_poses, d_out_E, rt().collect_stats);
} MMO_CATCH_ALL

int mmo_score_coords_components(const mmo_receptor *rec, const mmo_ligand *lig, int64_t n_poses,
                                const double *xs, const double *ys, const double *zs,
                                double *out_elec, double *out_vdw) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && lig, "mmo_score_coords_components: null handle");
    MMO_REQUIRE(n_poses >= 0, "mmo_score_coords_components: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out_elec && out_vdw, "mmo_score_coords_components: null buffer");
    const size_t m = (size_t)n_poses * lig->n;
    DevBuf<double> dx, dy, dz, de, dv;
    MMO_TRY(dx.upload(xs, m)); MMO_TRY(dy.upload(ys, m)); MMO_TRY(dz.upload(zs, m));
    MMO_TRY(de.alloc((size_t)n_poses)); MMO_TRY(dv.alloc((size_t)n_poses));
    MMO_TRY(launch_components_fp64(rec, lig, coords_src(dx.p, dy.p, dz.p), n_poses, de.p, dv.p));
    MMO_TRY(d2h_sync(out_elec, de.p, (size_t)n_poses * sizeof(double)));
    return d2h_sync(out_vdw, dv.p, (size_t)n_poses * sizeof(double));
} MMO_CATCH_ALL

int mmo_intra_nb(const mmo_ligand *lig, int64_t n_confs, const double *xs, const double *ys,
                 const double *zs, double *out_E) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(lig, "mmo_intra_nb: null handle");
    MMO_REQUIRE(lig->has_dists, "mmo_intra_nb: the ligand was created without topological distances");
    MMO_REQUIRE(n_confs >= 0, "mmo_intra_nb: negative conformer count");
    if (n_confs == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out_E, "mmo_intra_nb: null buffer");
    const size_t m = (size_t)n_confs * lig->n;
    DevBuf<double> dxyz, dE;
    const double *d[3];
    MMO_TRY(upload_xyz(dxyz, xs, ys, zs, m, d));
    MMO_TRY(dE.alloc((size_t)n_confs));
    MMO_TRY(launch_intra_fp64(lig, n_confs, d[0], d[1], d[2], dE.p));
    return d2h_sync(out_E, dE.p, (size_t)n_confs * sizeof(double));
} MMO_CATCH_ALL

int mmo_set_collect_stats(int on) try {
    rt().collect_stats = on != 0;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_last_pair_stats(int64_t *pairs_evaluated, int64_t *pairs_inside, int64_t *pairs_fp64) try {
    if (pairs_evaluated) *pairs_evaluated = rt().stat_pairs;
    if (pairs_inside) *pairs_inside = rt().stat_inside;
    if (pairs_fp64) *pairs_fp64 = rt().stat_fp64;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_selftest_division(uint64_t seed, int64_t n, int64_t *mismatches) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(n > 0 && mismatches, "mmo_selftest_division: bad arguments");
    return division_selftest(seed, n, mismatches);
} MMO_CATCH_ALL

int mmo_direct_set_mode(int mode) try {
    MMO_REQUIRE(mode >= 0 && mode <= 2, "mmo_direct_set_mode: mode must be 0 (auto), 1 (pose kernel) or 2 (item kernel)");
    direct_set_mode(mode);
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_last_fix_stats(int64_t *atoms_flagged) try {
    if (atoms_flagged) *atoms_flagged = rt().stat_flagged;
    return MMO_OK;
} MMO_CATCH_ALL

// ------------------------------------------------------------------ energy grids
static int grid_alloc(double step, const int32_t dims[3], int32_t T, mmo_grid **out) {
    MMO_REQUIRE(out != nullptr, "null grid output pointer");
    MMO_REQUIRE(step > 0.0 && dims && dims[0] > 1 && dims[1] > 1 && dims[2] > 1 && T > 0, "bad grid geometry");
    MMO_REQUIRE((double)dims[0] * dims[1] * dims[2] < 2.0e9, "grid too large for 32-bit voxel indices");
    mmo_grid *g = new mmo_grid();
    g->step = step;
    for (int d = 0; d < 3; d++) g->dims[d] = dims[d];
    g->T = T;
    g->nvox = (size_t)dims[0] * dims[1] * dims[2];
    int rc = g->maps.alloc(g->nvox * (size_t)T);
    if (rc != MMO_OK) { delete g; return rc; }
    *out = g;
    return MMO_OK;
}

int mmo_grid_build(const mmo_receptor *rec, double step, const int32_t dims[3],
                   const uint8_t *mask_bits, int32_t T, const int32_t *type_anum,
                   const double *type_q, float *out_maps, mmo_grid **out_grid) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && type_anum && type_q, "mmo_grid_build: null argument");
    mmo_grid *g = nullptr;
    MMO_TRY(grid_alloc(step, dims, T, &g));
    int rc = MMO_OK;
    do {
        // G3D.create: BA1.fill arr 0.0 (G3D.ml:47-51)
        cudaError_t e = cudaMemsetAsync(g->maps.p, 0, g->nvox * (size_t)T * sizeof(float), rt().stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "memset maps", __FILE__, __LINE__); break; }
        DevBuf<uint32_t> dmask;
        if (mask_bits) {
            size_t nwords = (g->nvox + 31) / 32;
            std::vector<uint32_t> w(nwords, 0u);
            memcpy(w.data(), mask_bits, (g->nvox + 7) / 8);
            if ((rc = dmask.upload(w))) break;
        }
        // The kernel visits the types sorted by element (stable): types of one element are neighbours inside a
        // thread's group, so the vdW term -- a function of the two elements only -- is formed once per element and
        // reused; the maps keep the caller's type order (tidx) and every value is the same double as before.
        std::vector<int32_t> ord(T), telt(T), tidx(T);
        for (int t = 0; t < T; t++) ord[t] = t;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return elt_index(type_anum[a]) < elt_index(type_anum[b]); });
        std::vector<double> tqs(T);
        for (int t = 0; t < T; t++) { telt[t] = elt_index(type_anum[ord[t]]); tqs[t] = type_q[ord[t]]; tidx[t] = ord[t]; }
        DevBuf<int32_t> dte, dti;
        DevBuf<double> dtq;
        if ((rc = dte.upload(telt)) || (rc = dtq.upload(tqs)) || (rc = dti.upload(tidx))) break;
        if ((rc = launch_grid_build(rec, g, mask_bits ? dmask.p : nullptr, dte.p, dtq.p, dti.p))) break;
        if (out_maps) rc = d2h_sync(out_maps, g->maps.p, g->nvox * (size_t)T * sizeof(float));
        else { cudaError_t e2 = cudaStreamSynchronize(rt().stream); if (e2 != cudaSuccess) rc = cuda_fail(e2, "sync", __FILE__, __LINE__); }
    } while (0);
    if (rc != MMO_OK || !out_grid) { delete g; g = nullptr; }
    if (out_grid) *out_grid = g;
    return rc;
} MMO_CATCH_ALL

int mmo_grid_upload(double step, const int32_t dims[3], int32_t T, const float *maps, mmo_grid **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(maps != nullptr, "mmo_grid_upload: null maps");
    mmo_grid *g = nullptr;
    MMO_TRY(grid_alloc(step, dims, T, &g));
    cudaError_t e = cudaMemcpyAsync(g->maps.p, maps, g->nvox * (size_t)T * sizeof(float), cudaMemcpyHostToDevice, rt().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "upload maps", __FILE__, __LINE__); }
    *out = g;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_grid_download(const mmo_grid *grid, float *maps) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(grid && maps, "mmo_grid_download: null argument");
    return d2h_sync(maps, grid->maps.p, grid->nvox * (size_t)grid->T * sizeof(float));
} MMO_CATCH_ALL

int mmo_grid_destroy(mmo_grid *grid) try {
    delete grid;
    return MMO_OK;
} MMO_CATCH_ALL

// G3D.to_ba1_file (G3D.ml:14-32): raw f32 image + `.dims` side-car with step/x_dim/y_dim/z_dim
int mmo_grid_write_ba1(const mmo_grid *grid, int32_t type, const char *path) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(grid && path, "mmo_grid_write_ba1: null argument");
    MMO_REQUIRE(type >= 0 && type < grid->T, "mmo_grid_write_ba1: type %d out of range", type);
    std::vector<float> h(grid->nvox);
    MMO_TRY(d2h_sync(h.data(), grid->maps.p + (size_t)type * grid->nvox, grid->nvox * sizeof(float)));
    FILE *f = fopen(path, "wb");
    MMO_REQUIRE(f != nullptr, "mmo_grid_write_ba1: cannot create %s", path);
    size_t w = fwrite(h.data(), sizeof(float), h.size(), f);
    fclose(f);
    MMO_REQUIRE(w == h.size(), "mmo_grid_write_ba1: short write to %s", path);
    std::string dn = std::string(path) + ".dims";
    f = fopen(dn.c_str(), "w");
    MMO_REQUIRE(f != nullptr, "mmo_grid_write_ba1: cannot create %s", dn.c_str());
    fprintf(f, "step: %g\nx_dim: %d\ny_dim: %d\nz_dim: %d\n", grid->step, grid->dims[0], grid->dims[1], grid->dims[2]);
    fclose(f);
    return MMO_OK;
} MMO_CATCH_ALL

// G3D.parse_dims_file + of_ba1_file (G3D.ml:37-64) for T maps of identical geometry
int mmo_grid_read_ba1(const char *const *paths, int32_t T, mmo_grid **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(paths && T > 0 && out, "mmo_grid_read_ba1: bad arguments");
    double step = 0.0;
    int32_t dims[3] = {0, 0, 0};
    std::vector<float> all;
    size_t nvox = 0;
    for (int t = 0; t < T; t++) {
        // a compressed cache file (lds.ml:540-553): uncompressed next to it, read, and the transient copy removed
        std::string fn = paths[t];
        const bool zst = fn.size() > 4 && fn.compare(fn.size() - 4, 4, ".zst") == 0;
        if (zst) {
            MMO_TRY(mmo_zstd_uncompress_file(fn.c_str()));
            fn.resize(fn.size() - 4);
        }
        struct Cleanup { std::string f; bool on; ~Cleanup() { if (on) remove(f.c_str()); } } cleanup{fn, zst};
        std::string dn = fn + ".dims";
        FILE *f = fopen(dn.c_str(), "r");
        MMO_REQUIRE(f != nullptr, "mmo_grid_read_ba1: cannot open %s", dn.c_str());
        double s; int a, b, c;
        int ok = fscanf(f, "step: %lf x_dim: %d y_dim: %d z_dim: %d", &s, &a, &b, &c);
        fclose(f);
        MMO_REQUIRE(ok == 4, "mmo_grid_read_ba1: cannot parse %s", dn.c_str());
        if (t == 0) {
            // the .dims side-car is untrusted text: check it before it sizes an allocation
            MMO_REQUIRE(s > 0.0 && a > 1 && b > 1 && c > 1 && (double)a * b * c * T < 2.0e9, "mmo_grid_read_ba1: %s: bad geometry (step %g, %d x %d x %d, %d maps)", dn.c_str(), s, a, b, c, T);
            step = s; dims[0] = a; dims[1] = b; dims[2] = c; nvox = (size_t)a * b * c; all.resize(nvox * (size_t)T);
        }
        MMO_REQUIRE(s == step && a == dims[0] && b == dims[1] && c == dims[2], "mmo_grid_read_ba1: %s has another geometry", dn.c_str());
        f = fopen(fn.c_str(), "rb");
        MMO_REQUIRE(f != nullptr, "mmo_grid_read_ba1: cannot open %s", fn.c_str());
        size_t r = fread(all.data() + (size_t)t * nvox, sizeof(float), nvox, f);
        int extra = fgetc(f);
        fclose(f);
        MMO_REQUIRE(r == nvox && extra == EOF, "mmo_grid_read_ba1: %s does not hold %zu floats", fn.c_str(), nvox);   // assert(BA1.dim ba1 = n)
    }
    return mmo_grid_upload(step, dims, T, all.data(), out);
} MMO_CATCH_ALL

int mmo_trilin(const mmo_grid *grid, int32_t type, int64_t n, const double *xs, const double *ys,
               const double *zs, double *out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(grid, "mmo_trilin: null grid");
    MMO_REQUIRE(type >= 0 && type < grid->T, "mmo_trilin: type %d out of range", type);
    MMO_REQUIRE(n >= 0, "mmo_trilin: negative count");
    if (n == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out, "mmo_trilin: null buffer");
    DevBuf<double> dx, dy, dz, dE;
    MMO_TRY(dx.upload(xs, (size_t)n)); MMO_TRY(dy.upload(ys, (size_t)n)); MMO_TRY(dz.upload(zs, (size_t)n));
    MMO_TRY(dE.alloc((size_t)n));
    MMO_TRY(launch_trilin(grid, type, n, dx.p, dy.p, dz.p, dE.p));
    return d2h_sync(out, dE.p, (size_t)n * sizeof(double));
} MMO_CATCH_ALL

static int check_interp(const mmo_grid *grid, const mmo_ligand *lig) {
    MMO_REQUIRE(grid && lig, "interp: null handle");
    MMO_REQUIRE(lig->has_typ, "interp: the ligand was created without FF atom types");
    for (int j = 0; j < lig->n; j++)
        MMO_REQUIRE(lig->htyp[j] >= 0 && lig->htyp[j] < grid->T, "interp: atom %d has type %d but the grid holds %d maps", j, lig->htyp[j], grid->T);
    return MMO_OK;
}

int mmo_score_interp_coords(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                            const double *xs, const double *ys, const double *zs, double *out_E) try {
    MMO_TRY(require_ready());
    MMO_TRY(check_interp(grid, lig));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_interp_coords: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out_E, "mmo_score_interp_coords: null buffer");
    const size_t m = (size_t)n_poses * lig->n;
    DevBuf<double> dxyz, dE;
    const double *d[3];
    MMO_TRY(upload_xyz(dxyz, xs, ys, zs, m, d));
    MMO_TRY(dE.alloc((size_t)n_poses));
    MMO_TRY(launch_interp(grid, lig, coords_src(d[0], d[1], d[2]), n_poses, dE.p));
    return d2h_sync(out_E, dE.p, (size_t)n_poses * sizeof(double));
} MMO_CATCH_ALL

int mmo_score_interp_poses(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                           const double *rot9, const double *trans3, double *out_E) try {
    MMO_TRY(require_ready());
    MMO_TRY(check_interp(grid, lig));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_interp_poses: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(rot9 && trans3 && out_E, "mmo_score_interp_poses: null buffer");
    DevBuf<double> drt, dE;
    const double *d[2];
    MMO_TRY(upload_rt(drt, rot9, trans3, (size_t)n_poses, d));
    MMO_TRY(dE.alloc((size_t)n_poses));
    MMO_TRY(launch_interp(grid, lig, rt_src(d[0], d[1]), n_poses, dE.p));
    return d2h_sync(out_E, dE.p, (size_t)n_poses * sizeof(double));
} MMO_CATCH_ALL

int mmo_score_interp_poses_dev(const mmo_grid *grid, const mmo_ligand *lig, int64_t n_poses,
                               const double *d_rot9, const double *d_trans3, double *d_out_E) try {
    MMO_TRY(require_ready());
    MMO_TRY(check_interp(grid, lig));
    MMO_REQUIRE(n_poses >= 0, "mmo_score_interp_poses_dev: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(d_rot9 && d_trans3 && d_out_E, "mmo_score_interp_poses_dev: null buffer");
    return launch_interp(grid, lig, rt_src(d_rot9, d_trans3), n_poses, d_out_E);
} MMO_CATCH_ALL

// ------------------------------------------------------------------ vdW occupancy mask
static int mask_alloc(double step, const int32_t dims[3], mmo_mask **out) {
    MMO_REQUIRE(step > 0.0 && dims && dims[0] > 1 && dims[1] > 1 && dims[2] > 1, "bad mask geometry");
    MMO_REQUIRE((double)dims[0] * dims[1] * dims[2] < 2.0e9, "mask too large for 32-bit voxel indices");
    mmo_mask *m = new mmo_mask();
    m->step = step;
    for (int d = 0; d < 3; d++) m->dims[d] = dims[d];
    m->nbits = (size_t)dims[0] * dims[1] * dims[2];
    size_t nwords = (m->nbits + 31) / 32 + 1;
    m->hwords.assign(nwords, 0u);
    int rc = m->words.alloc(nwords);
    if (rc != MMO_OK) { delete m; return rc; }
    *out = m;
    return MMO_OK;
}

int mmo_vdw_mask_build(int32_t n, const double *xs, const double *ys, const double *zs,
                       const double *radii, double step, const int32_t dims[3], uint8_t *out_bits,
                       mmo_mask **out_mask) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(n >= 0 && (n == 0 || (xs && ys && zs && radii)), "mmo_vdw_mask_build: bad atom arrays");
    mmo_mask *m = nullptr;
    MMO_TRY(mask_alloc(step, dims, &m));
    int rc = MMO_OK;
    do {
        cudaError_t e = cudaMemsetAsync(m->words.p, 0, m->hwords.size() * sizeof(uint32_t), rt().stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "memset mask", __FILE__, __LINE__); break; }
        DevBuf<double> dx, dy, dz, dr;
        if ((rc = dx.upload(xs, (size_t)n)) || (rc = dy.upload(ys, (size_t)n)) || (rc = dz.upload(zs, (size_t)n)) ||
            (rc = dr.upload(radii, (size_t)n)))
            break;
        if ((rc = launch_vdw_mask(n, dx.p, dy.p, dz.p, dr.p, m))) break;
        rc = d2h_sync(m->hwords.data(), m->words.p, m->hwords.size() * sizeof(uint32_t));
    } while (0);
    if (rc == MMO_OK && out_bits) memcpy(out_bits, m->hwords.data(), (m->nbits + 7) / 8);
    if (rc != MMO_OK || !out_mask) { delete m; m = nullptr; }
    if (out_mask) *out_mask = m;
    return rc;
} MMO_CATCH_ALL

// N3 masks.  mode 0: vdW volume (radii), 1: first solvent shell (radii + 1.4 set, radii unset), 2: whole protein (12 A)
static int atom_mask_build(int mode, int32_t n, const double *xs, const double *ys, const double *zs, const double *radii,
                           double step, const int32_t dims[3], uint8_t *out_bits, mmo_mask **out_mask) {
    MMO_TRY(require_ready());
    MMO_REQUIRE(n >= 0 && (n == 0 || (xs && ys && zs && (radii || mode == 2))), "mask build: bad atom arrays");
    mmo_mask *m = nullptr;
    MMO_TRY(mask_alloc(step, dims, &m));
    int rc = MMO_OK;
    do {
        cudaError_t e = cudaMemsetAsync(m->words.p, 0, m->hwords.size() * sizeof(uint32_t), rt().stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "memset mask", __FILE__, __LINE__); break; }
        std::vector<double> r1((size_t)n);
        for (int i = 0; i < n; i++) r1[i] = mode == 2 ? 12.0 : (mode == 1 ? radii[i] + 1.4 : radii[i]);   // const.ml:10,12
        DevBuf<double> dx, dy, dz, dr, dr0;
        if ((rc = dx.upload(xs, (size_t)n)) || (rc = dy.upload(ys, (size_t)n)) || (rc = dz.upload(zs, (size_t)n)) ||
            (rc = dr.upload(r1)))
            break;
        if ((rc = launch_vdw_mask(n, dx.p, dy.p, dz.p, dr.p, m, true))) break;
        if (mode == 1) {
            if ((rc = dr0.upload(radii, (size_t)n))) break;
            if ((rc = launch_vdw_mask(n, dx.p, dy.p, dz.p, dr0.p, m, false))) break;
        }
        rc = d2h_sync(m->hwords.data(), m->words.p, m->hwords.size() * sizeof(uint32_t));
    } while (0);
    if (rc == MMO_OK && out_bits) memcpy(out_bits, m->hwords.data(), (m->nbits + 7) / 8);
    if (rc != MMO_OK || !out_mask) { delete m; m = nullptr; }
    if (out_mask) *out_mask = m;
    return rc;
}

int mmo_mask_first_solvent_shell(int32_t n, const double *xs, const double *ys, const double *zs, const double *radii,
                                 double step, const int32_t dims[3], uint8_t *out_bits, mmo_mask **out_mask) try {
    return atom_mask_build(1, n, xs, ys, zs, radii, step, dims, out_bits, out_mask);
} MMO_CATCH_ALL

int mmo_mask_whole_protein(int32_t n, const double *xs, const double *ys, const double *zs,
                           double step, const int32_t dims[3], uint8_t *out_bits, mmo_mask **out_mask) try {
    return atom_mask_build(2, n, xs, ys, zs, nullptr, step, dims, out_bits, out_mask);
} MMO_CATCH_ALL

int mmo_mask_roi_only(const double roi_c[3], double roi_r, double step, const int32_t dims[3], uint8_t *out_bits,
                      mmo_mask **out_mask) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(roi_c && roi_r >= 0.0, "mmo_mask_roi_only: bad ROI");
    mmo_mask *m = nullptr;
    MMO_TRY(mask_alloc(step, dims, &m));
    const double r = roi_r + (12.0 * 2.0);               // lds.ml:272-275: E_inter must be able to vanish outside the ROI
    int rc = launch_sphere_mask(roi_c[0], roi_c[1], roi_c[2], r, m);
    if (rc == MMO_OK) rc = d2h_sync(m->hwords.data(), m->words.p, m->hwords.size() * sizeof(uint32_t));
    if (rc == MMO_OK && out_bits) memcpy(out_bits, m->hwords.data(), (m->nbits + 7) / 8);
    if (rc != MMO_OK || !out_mask) { delete m; m = nullptr; }
    if (out_mask) *out_mask = m;
    return rc;
} MMO_CATCH_ALL

int mmo_mask_upload(double step, const int32_t dims[3], const uint8_t *bits, mmo_mask **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(bits && out, "mmo_mask_upload: null argument");
    mmo_mask *m = nullptr;
    MMO_TRY(mask_alloc(step, dims, &m));
    memcpy(m->hwords.data(), bits, (m->nbits + 7) / 8);
    cudaError_t e = cudaMemcpyAsync(m->words.p, m->hwords.data(), m->hwords.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, rt().stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(rt().stream);
    if (e != cudaSuccess) { delete m; return cuda_fail(e, "upload mask", __FILE__, __LINE__); }
    *out = m;
    return MMO_OK;
} MMO_CATCH_ALL

// Utls.bitmask_to_file / bitmask_from_file (src/utls.ml:12-20): one line of '0'/'1' characters, Bitv.M.to_string mask.
// Bitv.M is the "most significant bit first" flavour of the bitv library (not vendored: the order is taken from its
// documentation, UNPINNED): the first character is bit n - 1, the last one bit 0.  msb_first = 0 gives Bitv.L's order.
int mmo_mask_write_bitmask(const mmo_mask *mask, const char *path, int msb_first) try {
    MMO_REQUIRE(mask && path, "mmo_mask_write_bitmask: null argument");
    const size_t n = mask->nbits;
    std::string line(n + 1, '0');
    for (size_t i = 0; i < n; i++) {
        const bool b = (mask->hwords[i >> 5] >> (i & 31)) & 1u;
        line[msb_first ? n - 1 - i : i] = b ? '1' : '0';
    }
    line[n] = '\n';
    FILE *f = fopen(path, "w");
    MMO_REQUIRE(f != nullptr, "mmo_mask_write_bitmask: cannot create %s", path);
    const size_t w = fwrite(line.data(), 1, line.size(), f);
    fclose(f);
    MMO_REQUIRE(w == line.size(), "mmo_mask_write_bitmask: short write to %s", path);
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_mask_read_bitmask(const char *path, double step, const int32_t dims[3], int msb_first, mmo_mask **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(path && dims && out, "mmo_mask_read_bitmask: null argument");
    *out = nullptr;
    MMO_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && (double)dims[0] * dims[1] * dims[2] < 4.0e9, "mmo_mask_read_bitmask: bad dims");
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    FILE *f = fopen(path, "r");
    MMO_REQUIRE(f != nullptr, "mmo_mask_read_bitmask: cannot open %s", path);
    std::string line(n + 2, '\0');
    const size_t r = fread(&line[0], 1, n + 2, f);
    fclose(f);
    // input_line: everything up to the first newline; Bitv.M.of_string raises on another length or character
    size_t len = 0;
    while (len < r && line[len] != '\n') len++;
    MMO_REQUIRE(len == n, "mmo_mask_read_bitmask: %s holds %zu bits, the grid has %zu voxels", path, len, n);
    std::vector<uint8_t> bits((n + 7) / 8 + 4, 0);
    for (size_t c = 0; c < n; c++) {
        const char ch = line[c];
        MMO_REQUIRE(ch == '0' || ch == '1', "mmo_mask_read_bitmask: %s: character %d at column %zu", path, (int)ch, c);
        const size_t i = msb_first ? n - 1 - c : c;
        if (ch == '1') bits[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
    return mmo_mask_upload(step, dims, bits.data(), out);
} MMO_CATCH_ALL

int mmo_mask_download(const mmo_mask *mask, uint8_t *out_bits) try {
    MMO_REQUIRE(mask && out_bits, "mmo_mask_download: null argument");
    memcpy(out_bits, mask->hwords.data(), (mask->nbits + 7) / 8);
    return MMO_OK;
} MMO_CATCH_ALL

// Utls.zstd_compress_file / zstd_uncompress_file (src/utls.ml:22-45): the reference shells out to the zstd program, and
// so does this (same command lines); an image without zstd gets MMO_EINVAL and keeps the uncompressed cache file
static bool shell_safe(const char *p) {
    for (; *p; p++)
        if (!(isalnum((unsigned char)*p) || *p == '/' || *p == '.' || *p == '_' || *p == '-' || *p == '+' || *p == ',' || *p == '=' || *p == '@')) return false;
    return true;
}
int mmo_zstd_compress_file(const char *path) try {
    MMO_REQUIRE(path && *path && shell_safe(path), "mmo_zstd_compress_file: path must be made of [A-Za-z0-9/._+,=@-]");
    const std::string cmd = std::string("zstd --rm -qf ") + path;        // remove quiet force
    const int ret = system(cmd.c_str());
    MMO_REQUIRE(ret == 0, "mmo_zstd_compress_file: command failed (is zstd installed?): %s", cmd.c_str());
    return MMO_OK;
} MMO_CATCH_ALL
int mmo_zstd_uncompress_file(const char *path_zst) try {
    MMO_REQUIRE(path_zst && shell_safe(path_zst), "mmo_zstd_uncompress_file: path must be made of [A-Za-z0-9/._+,=@-]");
    const size_t n = strlen(path_zst);
    MMO_REQUIRE(n > 4 && strcmp(path_zst + n - 4, ".zst") == 0, "mmo_zstd_uncompress_file: %s does not end in .zst", path_zst);
    const std::string cmd = std::string("zstd -dqfk ") + path_zst;       // decompress quiet force keep
    const int ret = system(cmd.c_str());
    MMO_REQUIRE(ret == 0, "mmo_zstd_uncompress_file: command failed (is zstd installed?): %s", cmd.c_str());
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_mask_destroy(mmo_mask *mask) try {
    delete mask;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_clash_poses(const mmo_mask *mask, const mmo_ligand *lig, int64_t n_poses,
                    const double *rot9, const double *trans3, uint8_t *out_flags) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(mask && lig, "mmo_clash_poses: null handle");
    MMO_REQUIRE(n_poses >= 0, "mmo_clash_poses: negative pose count");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(rot9 && trans3 && out_flags, "mmo_clash_poses: null buffer");
    DevBuf<double> dr, dt;
    DevBuf<uint8_t> df;
    MMO_TRY(dr.upload(rot9, (size_t)n_poses * 9));
    MMO_TRY(dt.upload(trans3, (size_t)n_poses * 3));
    MMO_TRY(df.alloc((size_t)n_poses));
    MMO_TRY(launch_clash(mask, lig, rt_src(dr.p, dt.p), n_poses, df.p));
    return d2h_sync(out_flags, df.p, (size_t)n_poses);
} MMO_CATCH_ALL

// ------------------------------------------------------------------ N3: ligand-defined binding site (scissors)
int mmo_carve_near_ligand(int32_t n_rec, const double *px, const double *py, const double *pz, int32_t n_lig,
                          const double *lx, const double *ly, const double *lz, double cutoff, uint8_t *out_keep,
                          int32_t *n_kept) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(n_rec >= 0 && n_lig >= 0 && cutoff >= 0.0, "mmo_carve_near_ligand: bad sizes");
    if (n_kept) *n_kept = 0;
    if (n_rec == 0) return MMO_OK;
    MMO_REQUIRE(px && py && pz && out_keep && (n_lig == 0 || (lx && ly && lz)), "mmo_carve_near_ligand: null buffer");
    DevBuf<double> dpx, dpy, dpz, dlx, dly, dlz;
    DevBuf<uint8_t> dk;
    MMO_TRY(dpx.upload(px, (size_t)n_rec)); MMO_TRY(dpy.upload(py, (size_t)n_rec)); MMO_TRY(dpz.upload(pz, (size_t)n_rec));
    MMO_TRY(dlx.upload(lx, (size_t)n_lig)); MMO_TRY(dly.upload(ly, (size_t)n_lig)); MMO_TRY(dlz.upload(lz, (size_t)n_lig));
    MMO_TRY(dk.alloc((size_t)n_rec));
    MMO_TRY(launch_carve(n_rec, dpx.p, dpy.p, dpz.p, n_lig, dlx.p, dly.p, dlz.p, cutoff, dk.p));
    MMO_TRY(d2h_sync(out_keep, dk.p, (size_t)n_rec));
    if (n_kept) for (int32_t i = 0; i < n_rec; i++) *n_kept += out_keep[i];
    return MMO_OK;
} MMO_CATCH_ALL

// ------------------------------------------------------------------ N4: desolvation sums
int mmo_desolv_protein(const mmo_receptor *rec, const mmo_mask *prot_shell, const double roi_c[3], double roi_r,
                       double *out_contribs, mmo_desolv **out) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(rec && prot_shell && roi_c && roi_r >= 0.0, "mmo_desolv_protein: bad arguments");
    if (out) *out = nullptr;
    mmo_desolv *d = new mmo_desolv();
    {
        mmo_mask *own = nullptr;
        int rc0 = mmo_mask_upload(prot_shell->step, prot_shell->dims, (const uint8_t *)prot_shell->hwords.data(), &own);
        if (rc0 != MMO_OK) { delete d; return rc0; }
        d->shell = own;
    }
    const double roi[4] = {roi_c[0], roi_c[1], roi_c[2], roi_r};
    int rc = d->contribs.alloc(prot_shell->nbits);
    if (rc == MMO_OK) rc = launch_desolv_protein(rec, prot_shell, roi, d->contribs.p);
    if (rc == MMO_OK && out_contribs) rc = d2h_sync(out_contribs, d->contribs.p, prot_shell->nbits * sizeof(double));
    if (rc == MMO_OK && !out_contribs) { cudaError_t e = cudaStreamSynchronize(rt().stream); if (e != cudaSuccess) rc = cuda_fail(e, "sync", __FILE__, __LINE__); }
    if (rc != MMO_OK || !out) { delete d->shell; delete d; d = nullptr; }
    if (out) *out = d;
    return rc;
} MMO_CATCH_ALL

static int desolv_penalty(const mmo_desolv *d, const mmo_ligand *lig, const PoseSrc &src, int64_t n_poses,
                          double *out_prot, double *out_lig) {
    MMO_REQUIRE(lig->has_r, "mmo_desolv_penalty: the ligand was created without radii");
    DevBuf<double> dr, dp, dl;
    MMO_TRY(dr.upload(lig->hr));
    MMO_TRY(dp.alloc((size_t)n_poses));
    MMO_TRY(dl.alloc((size_t)n_poses));
    MMO_TRY(launch_desolv_penalty(d->shell, d->contribs.p, lig, dr.p, src, n_poses, dp.p, dl.p));
    MMO_TRY(d2h_sync(out_prot, dp.p, (size_t)n_poses * sizeof(double)));
    return d2h_sync(out_lig, dl.p, (size_t)n_poses * sizeof(double));
}

int mmo_desolv_penalty_coords(const mmo_desolv *d, const mmo_ligand *lig, int64_t n_poses, const double *xs,
                              const double *ys, const double *zs, double *out_prot, double *out_lig) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(d && lig && n_poses >= 0, "mmo_desolv_penalty_coords: bad arguments");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(xs && ys && zs && out_prot && out_lig, "mmo_desolv_penalty_coords: null buffer");
    DevBuf<double> dx, dy, dz;
    const size_t n = (size_t)n_poses * lig->n;
    MMO_TRY(dx.upload(xs, n)); MMO_TRY(dy.upload(ys, n)); MMO_TRY(dz.upload(zs, n));
    return desolv_penalty(d, lig, coords_src(dx.p, dy.p, dz.p), n_poses, out_prot, out_lig);
} MMO_CATCH_ALL

int mmo_desolv_penalty_poses(const mmo_desolv *d, const mmo_ligand *lig, int64_t n_poses, const double *rot9,
                             const double *trans3, double *out_prot, double *out_lig) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(d && lig && n_poses >= 0, "mmo_desolv_penalty_poses: bad arguments");
    if (n_poses == 0) return MMO_OK;
    MMO_REQUIRE(rot9 && trans3 && out_prot && out_lig, "mmo_desolv_penalty_poses: null buffer");
    DevBuf<double> dr, dt;
    MMO_TRY(dr.upload(rot9, (size_t)n_poses * 9));
    MMO_TRY(dt.upload(trans3, (size_t)n_poses * 3));
    return desolv_penalty(d, lig, rt_src(dr.p, dt.p), n_poses, out_prot, out_lig);
} MMO_CATCH_ALL

int mmo_desolv_destroy(mmo_desolv *d) try {
    if (d) delete d->shell;
    delete d;
    return MMO_OK;
} MMO_CATCH_ALL

}  // extern "C"
