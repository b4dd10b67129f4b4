// common.cuh -- internal declarations shared by the translation units of libmmo_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/mmo_b200.h"

#define MMO_STR2(x) #x
#define MMO_STR(x) MMO_STR2(x)

namespace mmo {

// ---- error plumbing: no exception crosses the C ABI -------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define MMO_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return mmo::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define MMO_REQUIRE(cond, ...)                                                \
    do {                                                                      \
        if (!(cond)) { mmo::set_error(__VA_ARGS__); return MMO_EINVAL; }      \
    } while (0)
// every extern "C" entry point is a function-try-block closed by this macro: std::bad_alloc / std::length_error from a
// host-side container never unwinds into the C (or OCaml) caller
int on_exception();
#define MMO_CATCH_ALL catch (...) { return mmo::on_exception(); }
#define MMO_TRY(call)                                                         \
    do { int rc__ = (call); if (rc__ != MMO_OK) return rc__; } while (0)

struct Runtime {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // one-shot scans check the caller's rotation set behind the kernels (scan.cu)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void *l2_scratch = nullptr;
    size_t l2_scratch_bytes = 0;
    int64_t launches = 0;
    int64_t stat_pairs = 0, stat_inside = 0, stat_fp64 = 0, stat_flagged = 0;
    bool collect_stats = false;
    void *stage = nullptr;           // pinned staging buffer (stage_buffer()): [0, 64 KB) small host->device copies of api.cu,
                                     // [64 KB, 256 KB) per-slab read-backs of the scan driver
    int epoch = 0;                   // bumped by every mmo_init: device-resident caches of an older epoch are stale
};
Runtime &rt();
int require_ready();
constexpr size_t kStageHalf = 64 * 1024;        // uploads: [0, kStageHalf)
constexpr size_t kStageReadback = 192 * 1024;   // read-backs: [kStageHalf, kStageHalf + kStageReadback)
int stage_buffer(void **p);          // the pinned staging buffer, allocated on first use

// per-kernel device timing (CUDA events on the library stream), off unless mmo_kernel_timing(1)
enum KernelId { K_DIRECT_FP32 = 0, K_HARD_FIX, K_DIRECT_FP64, K_INTRA, K_GRID_BUILD, K_INTERP, K_PREFILTER,
                K_REDUCE, K_VDW_MASK, K_MC, K_ITEM_PREP, K_COUNT };
struct KernelScope {
    int id;
    bool on;
    explicit KernelScope(int kid);
    ~KernelScope();
};
inline void count_launch(int n = 1) { rt().launches += n; }
#define MMO_LAUNCH_CHECK()                                                    \
    do { mmo::count_launch(); MMO_CUDA(cudaGetLastError()); } while (0)

// Device allocations go through a small caching pool (runtime.cu): the one-shot entry points
// (mmo_scan, mmo_score_*) allocate and free the same buffer sizes on every call, and cudaMalloc/cudaFree
// cost more than the kernels of a small batch.
int pool_alloc(void **p, size_t bytes);
void pool_free(void *p);
void pool_trim();
void scan_drop_caches();   // scan.cu: device-resident rotation set
void direct_drop_caches(); // direct_fp32.cu: tables, work counters, item-mode scratch
void mc_drop_caches();     // mc.cu: UFF tables

// simple owning device buffer
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) pool_free(p); p = nullptr; n = 0; }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        MMO_TRY(pool_alloc((void **)&p, count * sizeof(T)));
        n = count;
        return MMO_OK;
    }
    int upload(const T *host, size_t count) {
        MMO_TRY(alloc(count));
        if (count) MMO_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, rt().stream));
        MMO_CUDA(cudaStreamSynchronize(rt().stream));
        return MMO_OK;
    }
    int upload(const std::vector<T> &v) { return upload(v.data(), v.size()); }
};

// ---- UFF parameters (src/UFF.ml:10-22) in a compact element index ---------------------------------
constexpr int kNumElt = 12;      // supported elements, index 0 = virtual element without vdW
constexpr int kEltUnsupported = 12;  // row/column of NaNs (src/UFF.ml:32-37)
constexpr int kEltTab = 13;
int elt_index(int anum);         // anum -> compact index, kEltUnsupported if not in the table
extern const int kEltAnum[kNumElt];
extern const double kEltXi[kNumElt];
extern const double kEltDi[kNumElt];
double vdw_radius(int anum);     // src/ptable.ml:41-54, NaN if unsupported
constexpr double kElecWeight = 332.0637 / 4.0;   // src/UFF.ml:25, src/const.ml:18
constexpr double kMaxE = 100000.0;               // src/params.ml:26

// ---- geometry of the fp32 direct kernel -----------------------------------------------------------
constexpr int kBlob = 16;        // receptor atoms per k-d leaf ("group"): first level of distance culling
// close-contact threshold: pairs with x_i*x_j / r^2 > kTau are re-evaluated in fp64
constexpr double kTau = 1.5;
constexpr float kFarAway = 1e6f;  // coordinate of padding atoms: beyond every cut-off, squares stay finite in fp32

struct Aabb { float lo[3], hi[3]; };

// host pose builders shared by the handle-based entry points (host_math.cu) and the file-based ones (molfile.cu)
struct HostLig {
    int n;
    const double *x, *y, *z;
    int n_rbonds;
    const int32_t *rb_left, *rb_right, *rg_off, *rg_idx;
};
double favg_host(const double *a, int n);          // Batteries A.favg as restated in the oracle (Kahan sum / n)
int apply_config_host(const HostLig &lig, const double *config, int32_t n_config, double *out_xs, double *out_ys,
                      double *out_zs, int32_t *too_long);
void rotated_copies_host(const HostLig &lig, const double center[3], int32_t n, const double *rot9, double *out_xs,
                         double *out_ys, double *out_zs);

}  // namespace mmo

// ---- opaque handles -----------------------------------------------------------------------------------
struct mmo_receptor {
    int n = 0;                       // real atoms
    int n_pad = 0;                   // n_blobs * kBlob
    int n_blobs = 0;
    double origin[3] = {0, 0, 0};    // fp32 coordinates are relative to this point
    double bb_lo[3] = {0, 0, 0}, bb_hi[3] = {0, 0, 0};   // bounding box of the atoms
    // original order, double (strict fp64 kernels)
    mmo::DevBuf<double> x, y, z, q;
    mmo::DevBuf<int32_t> elt;        // compact element index
    mmo::DevBuf<double4> xyzq64;     // {x, y, z, q} packed for the close-contact pass
    mmo::DevBuf<float4> xyz32v;      // positions relative to vox_lo, fp32 (pre-test of the close-contact pass)
    // group order, fp32 (fast kernel): xyzq = {x-ox, y-oy, z-oz, EW*q}; gelt = compact element index
    mmo::DevBuf<float4> xyzq;
    mmo::DevBuf<uint8_t> gelt;
    mmo::DevBuf<float4> blob_box;    // n_blobs x 2 : {lo xyz, 0}, {hi xyz, 0} (relative coordinates)
    mmo::DevBuf<float4> sup_box;     // n_sup x 2 : boxes of 32 consecutive groups
    int n_sup = 0;
    // close-contact voxel lists (fp64 correction pass)
    double vox_lo[3] = {0, 0, 0};
    double vox_edge = 2.0;
    int vox_dim[3] = {0, 0, 0};
    mmo::DevBuf<int32_t> vox_off;    // nvox + 1
    mmo::DevBuf<int32_t> vox_idx;    // atom indices (original order) | compact element index << 24
    double x_max = 0.0;              // largest UFF x_i among receptor atoms
    std::vector<double> hx, hy, hz, hq;   // host copies (original order)
    std::vector<int32_t> hanum;
};

struct mmo_ligand {
    int n = 0;
    std::vector<double> hx, hy, hz, hq, hr;
    std::vector<int32_t> hanum, htyp, hdists;
    bool has_r = false, has_typ = false, has_dists = false;
    int n_rbonds = 0;
    std::vector<int32_t> rb_left, rb_right, rg_off, rg_idx;
    double x_max = 0.0;
    // device copies
    mmo::DevBuf<double> x, y, z, q;          // template conformer
    mmo::DevBuf<int32_t> elt, typ;
    mmo::DevBuf<float4> fparam;              // {A_j, B_j, q_j, 0} in fast-path (k-d) order
    mmo::DevBuf<double> fx, fy, fz;          // template conformer in fast-path order
    int n_fast = 0;                          // atoms in fast-path order, padded to a multiple of 8
    mmo::DevBuf<int32_t> forder;             // fast-path position -> original atom index
    mmo::DevBuf<int32_t> pair_i, pair_j;     // interacting pairs (i<j, dists>=3) in reference order
    std::vector<int32_t> h_pair_i, h_pair_j; // host copy
    mmo::DevBuf<double4> mc_pair_tab;        // mc.cu, built on first use: {x_ij, d_ij, q_i q_j, bits(i | j << 16)} per pair
    int n_pairs = 0;
    mmo::DevBuf<int32_t> d_rb_left, d_rb_right, d_rg_off, d_rg_idx;
};

struct mmo_grid {
    double step = 0.0;
    int dims[3] = {0, 0, 0};
    int T = 0;
    size_t nvox = 0;
    mmo::DevBuf<float> maps;   // type-major: map t at maps + t*nvox
    // look-up copy, built on the first look-up (strict_fp64.cu: grid_zpairs): zpair[t][k][j][i] = {map[k][j][i], map[k+1][j][i]}
    // for k < z_dim - 1.  The eight corners of a trilinear cell are then four 8-byte loads in two rows of x instead of
    // eight 4-byte loads in four: half the L2 sectors per look-up, same floats, same arithmetic.
    mutable mmo::DevBuf<float2> zpair;
    mutable bool zpair_ready = false;
};

struct mmo_mask {
    double step = 0.0;
    int dims[3] = {0, 0, 0};
    size_t nbits = 0;
    mmo::DevBuf<uint32_t> words;   // bit idx at (words[idx>>5] >> (idx&31)) & 1
    std::vector<uint32_t> hwords;  // host copy (lattice-point AND test of the scan driver)
};

struct mmo_desolv {
    const mmo_mask *shell = nullptr;   // the protein's first solvent shell: the handle's own copy (2 MB for the 0.5 A grid),
                                       // so that the caller's mask may be destroyed (or garbage collected) first
    mmo::DevBuf<double> contribs;      // Lds.protein_desolv: one double per voxel, 0.0 outside shell AND ROI
};

namespace mmo {
// pose sources understood by the kernels
struct PoseSrc {
    int kind;                    // 0 = rot9/trans3 arrays, 1 = explicit coordinates, 2 = scan frames
    const double *rot9;          // kind 0: per pose; kind 2: per rotation
    const double *trans3;        // kind 0
    const double *xs, *ys, *zs;  // kind 1 (pose-major, stride L)
    const int64_t *frames;       // kind 2: frame ids
    int n_rot;                   // kind 2
    int lat_dims[3];             // kind 2
    double lat_min[3];           // kind 2
    double lat_q[3];             // kind 2: node i on axis d at lat_min[d] + i*lat_q[d] (grid.ml:49-51)
};

// kernels' host entry points (defined in the .cu files)
int launch_direct_fp32(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out, bool collect_stats);
int division_selftest(uint64_t seed, int64_t n, int64_t *mismatches);   // strict_fp64.cu: div_by vs '/'
void direct_set_mode(int mode);          // 0 auto, 1 pose-mode kernel always, 2 item-mode kernel for every pose list
int launch_direct_fp64(const mmo_receptor *rec, const mmo_ligand *lig, int variant, const PoseSrc &src,
                       int64_t n_poses, double *d_out);
int launch_components_fp64(const mmo_receptor *rec, const mmo_ligand *lig, const PoseSrc &src,
                           int64_t n_poses, double *d_elec, double *d_vdw);
int launch_intra_fp64(const mmo_ligand *lig, int64_t n_confs, const double *d_xs, const double *d_ys,
                      const double *d_zs, double *d_out);
int launch_grid_build(const mmo_receptor *rec, const mmo_grid *g, const uint32_t *d_mask_words,
                      const int32_t *d_type_elt, const double *d_type_q, const int32_t *d_type_idx);
int launch_interp(const mmo_grid *g, const mmo_ligand *lig, const PoseSrc &src, int64_t n_poses, double *d_out);
int grid_zpairs(const mmo_grid *g, const float2 **out);
int launch_trilin(const mmo_grid *g, int type, int64_t n, const double *d_x, const double *d_y,
                  const double *d_z, double *d_out);
int launch_vdw_mask(int n, const double *d_x, const double *d_y, const double *d_z, const double *d_r,
                    const mmo_mask *m, bool set_bits = true);
int launch_sphere_mask(double cx, double cy, double cz, double r, const mmo_mask *m);
int launch_clash(const mmo_mask *m, const mmo_ligand *lig, const PoseSrc &src, int64_t n_poses, uint8_t *d_flags);
int launch_carve(int n_rec, const double *d_px, const double *d_py, const double *d_pz, int n_lig, const double *d_lx,
                 const double *d_ly, const double *d_lz, double cutoff, uint8_t *d_keep);
// desolv.cu (N4): Lds.protein_desolv / Lds.desolvation_penalty
int launch_desolv_protein(const mmo_receptor *rec, const mmo_mask *shell, const double roi[4], double *d_contribs);
int launch_desolv_penalty(const mmo_mask *shell, const double *d_contribs, const mmo_ligand *lig, const double *d_radii,
                          const PoseSrc &src, int64_t n_poses, double *d_prot, double *d_lig);
int launch_scan_prefilter(const mmo_mask *m, const mmo_ligand *lig, const PoseSrc &src, const int64_t *d_points,
                          const int32_t *d_rot_perm, int64_t n_cand, int64_t *d_frames, unsigned long long *d_counter);

// balanced k-d ordering: permutation of n points such that consecutive groups of `leaf` points are
// spatially compact (only the last group may be short)
void vdw_factors(int elt, float *A, float *B);   // A = sqrt(D) x^6, B = sqrt(2D) x^3 (NaN if unsupported)
void kd_order(int n, const double *x, const double *y, const double *z, int leaf, std::vector<int> &order);

// host-side mirrors (host_math.cu)
void so3_rotations(int n, double *rot9);
void rot_r_xyz(double a, double b, double g, double r[9]);
void rot_decompose(const double r[9], double abg[3]);
int grid_num_steps(double dx, double length);
double grid_node(double step, int dim, int i);
}  // namespace mmo
