// C ABI of pyiga_b200 (see include/pyiga_b200.h): host-side table construction, kernel launches.
#include "backend.cuh"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/pyiga_b200.h"
#include "basis.cuh"
#include "entries.cuh"
#include "geo_fields.cuh"
#include "mlb.cuh"
#include "csr.cuh"
#include "plans.cuh"
#include "walk1.cuh"
#include "walk_geo.cuh"
#include "fused23.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static long long g_launches = 0;       // kernels launched by this library (reported by the benchmark)

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        pbError e_ = (call);                                                                       \
        if (e_ != pbSuccess)                                                                       \
            return fail(PB200_ECUDA, "%s failed: %s (%s:%d)", #call, pbErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// kernel launchers (CUDA) / sequential emulation (tests only, see backend.cuh)
// ------------------------------------------------------------------------------------------------
static void k_basis(const double* kv, int nk, int p, const double* nodes, int m, int nd, int* first, double* values,
                    pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(m, [&](long long g) { pb_basis_node(kv, nk, p, nodes, nd, first, values, (int)g); });
#else
    pb_basis_kernel<<<(m + 127) / 128, 128, 0, st>>>(kv, nk, p, nodes, m, nd, first, values);
#endif
}

struct BasisJobs {
    PbBasisBatch b;
    BasisJobs() { memset(&b, 0, sizeof b); }
    void add(const double* kv, int nk, int p, const double* nodes, int m, int nd, int* first, double* values) {
        const int j = b.njobs++;
        b.kv[j] = kv; b.nk[j] = nk; b.p[j] = p; b.nodes[j] = nodes; b.m[j] = m; b.nd[j] = nd; b.first[j] = first; b.values[j] = values;
    }
};
static void k_basis_batch(const BasisJobs& J, pbStream st) {
    if (J.b.njobs == 0) return;
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    for (int j = 0; j < J.b.njobs; ++j)
        pb_emu_for(J.b.m[j], [&](long long g) {
            pb_basis_node(J.b.kv[j], J.b.nk[j], J.b.p[j], J.b.nodes[j], J.b.nd[j], J.b.first[j], J.b.values[j], (int)g);
        });
#else
    int mmax = 0;
    for (int j = 0; j < J.b.njobs; ++j) mmax = std::max(mmax, J.b.m[j]);
    dim3 grid((unsigned)((mmax + 127) / 128), (unsigned)J.b.njobs);
    pb_basis_batch_kernel<<<grid, 128, 0, st>>>(J.b);
#endif
}

template <int DIM, class Prog>
static void k_fields(const PbFieldParams& prm, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(prm.pt_end - prm.pt_begin, [&](long long i) { pb_fields_point<DIM, Prog>(prm, prm.pt_begin + i); });
#else
    const long long blocks = std::min<long long>((prm.pt_end - prm.pt_begin + 255) / 256, 148LL * 32);
    pb_fields_kernel<DIM, Prog><<<(unsigned)blocks, 256, 0, st>>>(prm);
#endif
}

// row-wise K2 for spline geometries: rows [row_begin, row_end) of the grid (row = g0 or (g0,g1))
template <int DIM, int NC, class Prog>
static int k_fields_rows(const PbFieldParams& prm, long long row_begin, long long row_end, pbStream st) {
    const size_t ybytes = (size_t)prm.geo.Ng[DIM - 1] * NC * DIM * sizeof(double);
    g_launches += 1;
#ifdef PB_EMULATE
    std::vector<double> Y(ybytes / sizeof(double));
    pb_emu_for(row_end - row_begin, [&](long long r) { pb_fields_row_seq<DIM, NC, Prog>(prm, row_begin + r, Y.data()); });
    (void)st;
    return 0;
#else
    auto kern = pb_fields_row_kernel<DIM, NC, Prog>;
    const size_t gbytes = ybytes * PB_K2_ROWS;
    if (gbytes > 48 * 1024) {
        if (gbytes > 200 * 1024) return fail(PB200_EUNSUPPORTED, "geometry control net too long on the last axis");
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gbytes));
    }
    const long long groups = (row_end - row_begin + PB_K2_ROWS - 1) / PB_K2_ROWS;
    if (groups > 0) kern<<<(unsigned)groups, 128, gbytes, st>>>(prm, row_begin, row_end);
    return 0;
#endif
}

// values [pts][dim] and Jacobians [pts][dim][sdim] on a tensor grid
template <int DIM>
PB_HD void pb_geo_grid_point(const PbGeoDev& geo, const int* G, double* values, double* jac, long long idx) {
    int g[3];
    long long r = idx;
    for (int k = DIM - 1; k >= 0; --k) { g[k] = (int)(r % G[k]); r /= G[k]; }
    PbPoint pt;
    pb_geo_eval<DIM>(geo, g, pt);
    if (values)
        for (int i = 0; i < geo.dim; ++i) values[idx * geo.dim + i] = pt.x[i];
    if (jac)
        for (int i = 0; i < geo.dim; ++i)
            for (int j = 0; j < DIM; ++j) jac[(idx * geo.dim + i) * DIM + j] = pt.J[i][j];
}
#ifndef PB_EMULATE
template <int DIM>
__global__ void pb_geo_grid_kernel(PbGeoDev geo, int G0, int G1, int G2, long long npts, double* values, double* jac) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int G[3] = {G0, G1, G2};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < npts; idx += stride)
        pb_geo_grid_point<DIM>(geo, G, values, jac, idx);
}
#endif
template <int DIM>
static void k_geo_grid(const PbGeoDev& geo, const int* G, long long npts, double* values, double* jac, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(npts, [&](long long i) { pb_geo_grid_point<DIM>(geo, G, values, jac, i); });
#else
    const long long blocks = std::min<long long>((npts + 255) / 256, 148LL * 32);
    pb_geo_grid_kernel<DIM><<<(unsigned)blocks, 256, 0, st>>>(geo, G[0], G[1], DIM > 2 ? G[2] : 1, npts, values, jac);
#endif
}

template <int DIM>
static void k_entries(const PbEntryParams& prm, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(prm.n, [&](long long e) { prm.out[e] = pb_entry<DIM>(prm, prm.ij[2 * e], prm.ij[2 * e + 1]); });
#else
    const long long blocks = std::min<long long>((prm.n + 127) / 128, 148LL * 64);
    pb_entries_kernel<DIM><<<(unsigned)blocks, 128, 0, st>>>(prm);
#endif
}
template <int DIM>
static void k_entries_mlb(const PbEntryParams& prm, long long mu0_begin, long long count, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(count, [&](long long e) { prm.out[e] = pb_entry_mlb<DIM>(prm, mu0_begin, e); });
#else
    const long long blocks = std::min<long long>((count + 127) / 128, 148LL * 64);
    pb_entries_mlb_kernel<DIM><<<(unsigned)blocks, 128, 0, st>>>(prm, mu0_begin, count);
#endif
}

template <class IdxT>
static void k_csr(const PbMlbParams& p, long long nrows, long long count, IdxT* indptr, IdxT* indices, double* values,
                  pbStream st) {
    g_launches += 2;
#ifdef PB_EMULATE
    if (indptr) pb_emu_for(nrows + 1, [&](long long r) { pb_csr_indptr_row<IdxT>(p, nrows, indptr, r); });
    pb_emu_for(nrows, [&](long long r) {
        int i[3] = {0, 0, 0}, rs[3] = {0, 0, 0}, nb[3] = {1, 1, 1}, jm[3] = {0, 0, 0};
        pb_csr_row_tables(p, r, i, rs, nb, jm);
        const long long rowoff = pb_csr_row_offset(p, i);
        for (int e = 0; e < nb[0] * nb[1] * nb[2]; ++e) pb_csr_fill_row_entry<IdxT>(p, i, rs, nb, jm, rowoff, e, indices, values);
    });
    (void)count;
#else
    const unsigned b1 = (unsigned)((nrows + 1 + 255) / 256);
    const unsigned b2 = (unsigned)std::min<long long>((nrows + 7) / 8, 148LL * 32);
    if (indptr) pb_csr_indptr_kernel<IdxT><<<b1, 256, 0, st>>>(p, nrows, indptr);
    pb_csr_fill_rows_kernel<IdxT><<<b2, 256, 0, st>>>(p, nrows, indices, values);
    (void)count;
#endif
}

static void k_matvec(const PbMlbParams& p, long long nrows, const double* x, int x_j0, double* y, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    pb_emu_for(nrows, [&](long long r) { pb_mlb_matvec_row(p, x, x_j0, y, r); });
#else
    pb_mlb_matvec_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(p, nrows, x, x_j0, y);
#endif
}

static void k_modek(const double* A, int m, int n, const double* x, long long outer, long long inner, double* y,
                    pbStream st) {
    ++g_launches;
    const long long total = outer * m * inner;
#ifdef PB_EMULATE
    pb_emu_for(total, [&](long long e) { pb_modek_elem(A, m, n, x, inner, y, e); });
#else
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 64);
    pb_modek_kernel<<<blocks, 256, 0, st>>>(A, m, n, x, outer, inner, y);
#endif
}

extern "C" const char* pb200_last_error(void) { return g_err.c_str(); }
extern "C" int pb200_version(void) { return 100; }
extern "C" long long pb200_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------
// walk-kernel registry
// ------------------------------------------------------------------------------------------------
typedef std::map<std::tuple<int, int, int>, PbWalkLaunch> WalkRegistry;
static WalkRegistry& registry() {
    static WalkRegistry r;      // function-local: safe to use from other objects' static initialisers
    return r;
}
extern "C" void pb200_register_walk(int plan_id, int P, int Q, PbWalkLaunch fn) {
    registry()[std::make_tuple(plan_id, P, Q)] = fn;
}
PbWalkLaunch pb_find_walk(int plan_id, int P, int Q) {
    auto it = registry().find(std::make_tuple(plan_id, P, Q));
    return it == registry().end() ? nullptr : it->second;
}

typedef std::map<std::pair<int, int>, PbWalk1Launch> Walk1Registry;
static Walk1Registry& registry1() {
    static Walk1Registry r;
    return r;
}
extern "C" void pb200_register_walk1(int P, int Q, PbWalk1Launch fn) { registry1()[std::make_pair(P, Q)] = fn; }
typedef std::map<std::tuple<int, int, int>, PbS32Launch> S32Registry;
static S32Registry& registry_s32() {
    static S32Registry r;
    return r;
}
extern "C" void pb200_register_s32(int form, int P, int Q, PbS32Launch fn) { registry_s32()[std::make_tuple(form, P, Q)] = fn; }
PbS32Launch pb_find_s32(int form, int P, int Q) {
    auto it = registry_s32().find(std::make_tuple(form, P, Q));
    return it == registry_s32().end() ? nullptr : it->second;
}
PbWalk1Launch pb_find_walk1(int P, int Q) {
    auto it = registry1().find(std::make_pair(P, Q));
    return it == registry1().end() ? nullptr : it->second;
}

// ------------------------------------------------------------------------------------------------
// host-side axis tables
// ------------------------------------------------------------------------------------------------
struct KvHost {
    int p = 0;
    std::vector<double> kv;
    std::vector<double> mesh;       // distinct knots
    std::vector<int> k2m;           // knot index -> mesh index
    std::vector<int> first;         // per span: first active function
    std::vector<int> supp;          // [N][2] span range of each function
    int N() const { return (int)kv.size() - p - 1; }
    int nspans() const { return (int)mesh.size() - 1; }
};

static int build_kv(const double* knots, int nk, int p, KvHost& K) {
    if (!knots || nk < 2 * (p + 1) || p < 0) return fail(PB200_EINVAL, "invalid knot vector (nknots=%d, p=%d)", nk, p);
    if (p > PB_MAXP) return fail(PB200_EUNSUPPORTED, "spline degree %d above the supported maximum %d", p, PB_MAXP);
    K.p = p;
    K.kv.assign(knots, knots + nk);
    for (int i = 1; i < nk; ++i)
        if (K.kv[i] < K.kv[i - 1]) return fail(PB200_EINVAL, "knots should be increasing");
    K.mesh.clear();
    K.k2m.resize(nk);
    for (int i = 0; i < nk; ++i) {
        if (i == 0 || K.kv[i] != K.kv[i - 1]) K.mesh.push_back(K.kv[i]);
        K.k2m[i] = (int)K.mesh.size() - 1;
    }
    const int n = K.nspans(), N = K.N();
    if (n < 1 || N < 1) return fail(PB200_EINVAL, "knot vector has no spans");
    K.first.assign(n, 0);
    for (int i = 0; i + 1 < nk; ++i)
        if (K.kv[i] != K.kv[i + 1]) K.first[K.k2m[i]] = i - p;   // knot span i is mesh span k2m[i]
    K.supp.resize(2 * N);
    for (int j = 0; j < N; ++j) {
        K.supp[2 * j] = K.k2m[j];
        K.supp[2 * j + 1] = K.k2m[j + p + 1];
    }
    for (int s = 0; s < n; ++s)
        if (K.first[s] < 0 || K.first[s] + p >= N) return fail(PB200_EINVAL, "knot vector is not open (span %d)", s);
    return 0;
}

// band structure: (i,j), i over test functions, j over trial functions with joint support,
// sorted by i then j  (pyiga/mlmatrix.py:420-440)
static void build_band(const KvHost& U, const KvHost& V, std::vector<int>& row_start, std::vector<int>& jmin,
                       std::vector<int>& pi, std::vector<int>& pj) {
    const int Nv = V.N(), Nu = U.N();
    row_start.assign(Nv + 1, 0);
    jmin.assign(Nv, 0);
    pi.clear();
    pj.clear();
    int j0 = 0;
    for (int i = 0; i < Nv; ++i) {
        const int a = V.supp[2 * i], b = V.supp[2 * i + 1];
        while (j0 < Nu && U.supp[2 * j0 + 1] <= a) ++j0;     // supports end in non-decreasing order
        jmin[i] = j0;
        row_start[i] = (int)pi.size();
        for (int j = j0; j < Nu && U.supp[2 * j] < b; ++j) {
            if (std::min(b, U.supp[2 * j + 1]) > std::max(a, U.supp[2 * j])) {
                pi.push_back(i);
                pj.push_back(j);
            }
        }
    }
    row_start[Nv] = (int)pi.size();
}

extern "C" int pb200_band_structure(const double* ku, int nku, int pu, const double* kvv, int nkv, int pv,
                                    uint32_t* h_bidx, int* nband) {
    KvHost U, V;
    int rc = build_kv(ku, nku, pu, U);
    if (rc) return rc;
    rc = kvv ? build_kv(kvv, nkv, pv, V) : build_kv(ku, nku, pu, V);
    if (rc) return rc;
    std::vector<int> rs, jm, pi, pj;
    build_band(U, V, rs, jm, pi, pj);
    if (nband) *nband = (int)pi.size();
    if (h_bidx)
        for (size_t m = 0; m < pi.size(); ++m) {
            h_bidx[2 * m] = (uint32_t)pi[m];
            h_bidx[2 * m + 1] = (uint32_t)pj[m];
        }
    return 0;
}

struct AxisHost {
    KvHost U, V;
    bool same = true;
    int n = 0, q = 0, G = 0, M = 0;
    std::vector<double> nodes, weights;
    std::vector<int> row_start, jmin, pair_i, pair_j, tr, ret_mu;
};

// ------------------------------------------------------------------------------------------------
// device pool: one allocation for all small tables, 256-byte aligned pieces
// ------------------------------------------------------------------------------------------------
struct Pool {
    std::vector<char> host;
    char* dev = nullptr;
    size_t put(const void* p, size_t bytes) {
        size_t off = (host.size() + 255) & ~size_t(255);
        host.resize(off + bytes);
        if (p) memcpy(host.data() + off, p, bytes);
        return off;
    }
    template <class T> size_t put(const std::vector<T>& v) { return put(v.data(), v.size() * sizeof(T)); }
};

// structure-only view used by the MLB utilities (CSR export, matvec)
struct pb200_mlstruct {
    int device = 0;
    int dim = 0;
    int Nv[PB_MAXDIM] = {1, 1, 1}, Nu[PB_MAXDIM] = {1, 1, 1}, M[PB_MAXDIM] = {1, 1, 1};
    std::vector<int> row_start0;            // host copy of the axis-0 row offsets (slab sizes)
    const int* d_row_start[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    const int* d_jmin[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    const int* d_pair_i[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    const int* d_pair_j[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    void* mem = nullptr;                    // owned device memory (stand-alone structures)
};

struct pb200_assembler {
    pb200_mlstruct ml;
    int device = 0;
    int dim = 0;
    int form = 0;
    int nfields = 0;
    bool symmetric = false;
    bool same_space = true;
    int arity = 2;
    AxisHost hax[PB_MAXDIM];
    PbAxis dax[PB_MAXDIM];
    size_t off_knots_u[PB_MAXDIM], off_knots_v[PB_MAXDIM];
    Pool pool;
    double* d_fields = nullptr;
    std::vector<PbTerm> terms;
    std::vector<std::pair<int, int>> field_slots;       // custom forms: (test slot, trial slot) of field f
    void* geo_scratch = nullptr;
    size_t geo_scratch_bytes = 0;
    PbGeoDev geo_dev;                                   // device tables of the bound spline geometry (in geo_scratch)
    bool geo_valid = false;                             // geo_dev describes the current geometry
    bool fields_valid = false;                          // d_fields holds the fields of the current geometry
    bool fuse = true;                                   // 3D mass / stiffness: fused stage 1 (geometry + fields + axis 0, walk_geo.cuh)
    bool fuse23 = true;                                 // 3D mass / stiffness: fused stages 2 + 3 (fused23.cuh), no X2 in HBM
    long long npts = 0, nnz = 0;
    int fast = 0;
    bool ext_slots = false;                             // some term has a second or mixed derivative slot (PB_SLOT_EXT)
    const double* walk_table = nullptr;                 // run_stage: table override of the current launch
    // fused stages 2 + 3: device list of the band entries of axis 0 a launch computes (cached per slab)
    std::vector<int> s32_keep_host[2];
    int s32_keep_flip = 0;
    int* s32_keep_dev = nullptr;
    size_t s32_keep_cap = 0;
    long long s32_keep_key[5] = {-1, -1, -1, -1, -1};
    int s32_nkeep = 0;
    bool pack_tails = true;                             // several short last batches share a block
    int s32_whole = -1;                                 // tests: with walk_split > 1, this many tasks stay unsplit
    const double* pair_tab[PB_MAXDIM][3] = {};          // two-row tables of derivative orders (0,1) (0,2) (1,2) for the walks
    bool lane_ok[PB_MAXDIM] = {false, false, false};   // single interior knots on the axis
    bool force_walk = false;                            // debugging / tests: never use the lane-span kernels
    bool lane_v1 = false;                               // use the register-prefetch version of the lane-span kernel
    int lane_lines = 64;                                // lines per warp of the lane-span kernel (<= PbLaneCfg::DQ)
    bool walk_rot = true;                               // rotating-window walk on regular axes (no register shifts)
    int walk_split = 0;                                 // pieces of the walk axis: 0 = choose by occupancy, 1 = never, K = always K
    int sm_count = 0;                                   // multiprocessors of the device
    bool mirror_opt = true;                             // symmetric forms: compute the upper half of the final stage, mirror the rest
    bool fused_plans = true;                            // multi-output stage kernels (S1A/S1B/S2B); false: one launch per output
    // optional per-kernel timing of the last assemble call (CUDA events on the launch stream)
    bool timing = false;
    std::vector<std::string> stage_names;
#ifndef PB_EMULATE
    std::vector<cudaEvent_t> stage_events;
#endif
};

static void mark_stage(pb200_assembler* a, const char* name, pbStream st) {
    if (!a->timing) return;
#ifndef PB_EMULATE
    const size_t i = a->stage_names.size();
    if (a->stage_events.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        a->stage_events.push_back(e);
    }
    cudaEventRecord(a->stage_events[i], st);
#else
    (void)st;
#endif
    a->stage_names.push_back(name);
}

extern "C" int pb200_asm_set_option(pb200_assembler* a, const char* name, int value) {
    if (!a || !name) return fail(PB200_EINVAL, "null argument");
    if (!strcmp(name, "force_walk")) { a->force_walk = value != 0; return 0; }
    if (!strcmp(name, "lane_v1")) { a->lane_v1 = value != 0; return 0; }
    if (!strcmp(name, "walk_rot")) { a->walk_rot = value != 0; return 0; }
    if (!strcmp(name, "walk_split")) { if (value < 0 || value > PB_WALK_MAXSPLIT) return fail(PB200_EINVAL, "walk_split must be in 0..%d", PB_WALK_MAXSPLIT); a->walk_split = value; return 0; }
    if (!strcmp(name, "lane_lines")) { if (value < 1 || value > 64) return fail(PB200_EINVAL, "lane_lines must be in 1..64"); a->lane_lines = value; return 0; }
    if (!strcmp(name, "fused_plans")) { a->fused_plans = value != 0; return 0; }
    if (!strcmp(name, "mirror")) { a->mirror_opt = value != 0; return 0; }
    if (!strcmp(name, "fuse")) { a->fuse = value != 0; return 0; }
    if (!strcmp(name, "fuse23")) { a->fuse23 = value != 0; return 0; }
    if (!strcmp(name, "pack_tails")) { a->pack_tails = value != 0; return 0; }
    if (!strcmp(name, "s32_whole")) { a->s32_whole = value; return 0; }
    return fail(PB200_EINVAL, "unknown option '%s'", name);
}

extern "C" int pb200_asm_set_timing(pb200_assembler* a, int enable) {
    if (!a) return fail(PB200_EINVAL, "null handle");
    a->timing = enable != 0;
    a->stage_names.clear();
    return 0;
}

// durations (ms) between consecutive marks of the last assemble call; names joined by ';'
extern "C" int pb200_asm_get_timing(pb200_assembler* a, int max_stages, float* ms, char* names, int names_len, int* nstages) {
    if (!a || !ms || !nstages) return fail(PB200_EINVAL, "null argument");
    const int n = (int)a->stage_names.size() - 1;
    *nstages = n > 0 ? n : 0;
    std::string joined;
    for (int i = 0; i < n && i < max_stages; ++i) {
#ifndef PB_EMULATE
        CK(cudaEventSynchronize(a->stage_events[i + 1]));
        CK(cudaEventElapsedTime(&ms[i], a->stage_events[i], a->stage_events[i + 1]));
#else
        ms[i] = 0.f;
#endif
        joined += a->stage_names[i];
        joined += ';';
    }
    if (names && names_len > 0) {
        strncpy(names, joined.c_str(), names_len - 1);
        names[names_len - 1] = 0;
    }
    return 0;
}

static bool have_plan(int plan, int P, int Q) { return pb_find_walk(plan, P, Q) != nullptr; }

static bool ext_plan_ok(const pb200_assembler* a);
static int detect_fast_path(const pb200_assembler* a) {
    if (!a->same_space) return 0;
    if (a->form == PB200_FORM_CUSTOM && a->arity == 1) {    // every axis is contracted on its own
        for (int k = 0; k < a->dim; ++k)
            if (!pb_find_walk1(a->hax[k].U.p, a->hax[k].q)) return 0;
        return 1;
    }
    const int Q = a->hax[0].q;
    for (int k = 0; k < a->dim; ++k)
        if (a->hax[k].q != Q) return 0;
    auto P = [&](int k) { return a->hax[k].U.p; };
    if (a->form == PB200_FORM_MASS) {
        for (int k = 0; k < a->dim; ++k)
            if (!have_plan(PB_PLAN_COPY, P(k), Q)) return 0;
        return 1;
    }
    if (a->form == PB200_FORM_CUSTOM && a->arity == 1) {
        for (int k = 0; k < a->dim; ++k)
            if (!pb_find_walk1(P(k), Q)) return 0;
        return 1;
    }
    if (a->form == PB200_FORM_CUSTOM) {
        for (int k = 0; k < a->dim; ++k)
            if (!have_plan(PB_PLAN_GEN4, P(k), Q) || !have_plan(PB_PLAN_COPY, P(k), Q)) return 0;
        return 1;
    }
    if (a->form == PB200_FORM_STIFFNESS) {
        if (a->dim == 2)
            return have_plan(PB_PLAN_S1_2D, P(0), Q) && have_plan(PB_PLAN_FINAL4, P(1), Q);
        return have_plan(PB_PLAN_S1A, P(0), Q) && have_plan(PB_PLAN_S1B, P(0), Q) && have_plan(PB_PLAN_FINAL4, P(1), Q)
               && have_plan(PB_PLAN_S2B, P(1), Q) && have_plan(PB_PLAN_FINAL4, P(2), Q) && have_plan(PB_PLAN_ONE11, P(0), Q)
               && have_plan(PB_PLAN_PAIRT, P(1), Q);
    }
    return 0;
}

static int run_basis(pb200_assembler* a, pbStream st) {
    BasisJobs J;        // all axes (and both spaces) in one launch
    for (int k = 0; k < a->dim; ++k) {
        AxisHost& H = a->hax[k];
        PbAxis& D = a->dax[k];
        const double* dku = reinterpret_cast<const double*>(a->pool.dev + a->off_knots_u[k]);
        J.add(dku, (int)H.U.kv.size(), H.U.p, D.nodes, H.G, D.nd, nullptr, const_cast<double*>(D.Vu));
        if (!H.same) {
            const double* dkv = reinterpret_cast<const double*>(a->pool.dev + a->off_knots_v[k]);
            J.add(dkv, (int)H.V.kv.size(), H.V.p, D.nodes, H.G, D.nd, nullptr, const_cast<double*>(D.Vv));
        }
    }
    k_basis_batch(J, st);
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_asm_create(const pb200_desc* desc, int device, void* stream, pb200_assembler** out) {
    if (!desc || !out) return fail(PB200_EINVAL, "null argument");
    if (desc->dim < 2 || desc->dim > 3) return fail(PB200_EINVAL, "Assembler requires 2 or 3 knot vectors (dim=%d)", desc->dim);
    if (desc->form != PB200_FORM_MASS && desc->form != PB200_FORM_STIFFNESS && desc->form != PB200_FORM_CUSTOM)
        return fail(PB200_EINVAL, "unknown form id %d", desc->form);
    std::unique_ptr<pb200_assembler> a(new pb200_assembler);
    a->device = device;
#ifndef PB_EMULATE
    cudaDeviceGetAttribute(&a->sm_count, cudaDevAttrMultiProcessorCount, device);
#endif
    a->dim = desc->dim;
    a->form = desc->form;
    const int dim = desc->dim;

    a->npts = 1;
    a->nnz = 1;
    for (int k = 0; k < dim; ++k) {
        const pb200_axis_desc& X = desc->axis[k];
        AxisHost& H = a->hax[k];
        int rc = build_kv(X.h_knots_trial, X.nknots_trial, X.p_trial, H.U);
        if (rc) return rc;
        H.same = (X.h_knots_test == nullptr);
        rc = H.same ? build_kv(X.h_knots_trial, X.nknots_trial, X.p_trial, H.V)
                    : build_kv(X.h_knots_test, X.nknots_test, X.p_test, H.V);
        if (rc) return rc;
        if (H.U.mesh != H.V.mesh) return fail(PB200_EINVAL, "trial and test space must share the mesh (axis %d)", k);
        if (!H.same) a->same_space = false;
        H.n = H.U.nspans();
        H.q = X.nq;
        if (H.q < 1 || H.q > 16) return fail(PB200_EINVAL, "invalid number of Gauss nodes per span: %d", H.q);
        H.G = H.n * H.q;
        if (!X.h_nodes || !X.h_weights) return fail(PB200_EINVAL, "Gauss rule missing (axis %d)", k);
        H.nodes.assign(X.h_nodes, X.h_nodes + H.G);
        H.weights.assign(X.h_weights, X.h_weights + H.G);
        for (int s = 0; s < H.n; ++s)
            for (int g = 0; g < H.q; ++g) {
                const double t = H.nodes[s * H.q + g];
                if (!(t >= H.U.mesh[s] && t <= H.U.mesh[s + 1]))
                    return fail(PB200_EINVAL, "Gauss node %d of axis %d lies outside its span", s * H.q + g, k);
            }
        build_band(H.U, H.V, H.row_start, H.jmin, H.pair_i, H.pair_j);
        H.M = (int)H.pair_i.size();
        // transposed pairs and the retire table of the walk kernels (test == trial only)
        H.tr.assign(H.M, -1);
        const int P = H.U.p, N = H.V.N();
        H.ret_mu.assign((size_t)N * (2 * P + 1), -1);
        if (H.same) {
            auto mu_of = [&](int i, int j) -> int {
                if (i < 0 || j < 0 || i >= N || j >= N) return -1;
                const int off = j - H.jmin[i];
                if (off < 0 || off >= H.row_start[i + 1] - H.row_start[i]) return -1;
                return H.row_start[i] + off;
            };
            for (int m = 0; m < H.M; ++m) H.tr[m] = mu_of(H.pair_j[m], H.pair_i[m]);
            for (int f = 0; f < N; ++f)
                for (int kk = 0; kk <= 2 * P; ++kk) {
                    const int i = (kk <= P) ? f : f + (kk - P);
                    const int j = (kk <= P) ? f + kk : f;
                    H.ret_mu[(size_t)f * (2 * P + 1) + kk] = mu_of(i, j);
                }
        }
        a->npts *= H.G;
        a->nnz *= H.M;
        a->lane_ok[k] = H.same;
        for (int s = 1; s < H.n; ++s)
            if (H.U.first[s] != H.U.first[s - 1] + 1) a->lane_ok[k] = false;
    }

    if (desc->form == PB200_FORM_MASS) {
        a->nfields = 1;
        a->symmetric = a->same_space;
        a->terms.push_back(PbTerm{0, 0, 0});
    } else if (desc->form == PB200_FORM_STIFFNESS) {
        a->nfields = dim * (dim + 1) / 2;
        a->symmetric = a->same_space;
        // sum_{a,b} B[b][a] d_a u d_b v with x,y,z indices a,b <-> tensor axes dim-1-a, dim-1-b
        for (int b = 0; b < dim; ++b)
            for (int c = 0; c < dim; ++c) {
                const int lo = std::min(b, c), hi = std::max(b, c);
                const int sym = lo * dim - lo * (lo - 1) / 2 + (hi - lo);
                a->terms.push_back(PbTerm{sym, 1 + (dim - 1 - b), 1 + (dim - 1 - c)});
            }
    } else {
        if (desc->nfields < 1 || desc->nfields > PB_MAXFIELDS) return fail(PB200_EINVAL, "invalid number of fields %d", desc->nfields);
        if (desc->nterms < 1 || desc->nterms > PB_MAXTERMS || !desc->terms) return fail(PB200_EINVAL, "invalid number of terms %d", desc->nterms);
        a->nfields = desc->nfields;
        a->symmetric = desc->symmetric != 0;
        int nlinear = 0;
        for (int t = 0; t < desc->nterms; ++t) {
            const pb200_term& T = desc->terms[t];
            auto slot_ok = [&](int sl) {
                if (sl >= 0 && sl <= dim) return true;
                return sl >= PB_SLOT_EXT && sl < PB_SLOT_EXT + (dim == 3 ? 27 : 9);
            };
            if (T.field < 0 || T.field >= a->nfields || !slot_ok(T.slot_test) || (T.slot_trial != -1 && !slot_ok(T.slot_trial)))
                return fail(PB200_EINVAL, "invalid term %d", t);
            nlinear += T.slot_trial < 0;
            const int bt = pb_slot_contract(T.slot_test, -1), bu = T.slot_trial < 0 ? -1 : pb_slot_contract(T.slot_trial, -1);
            a->ext_slots = a->ext_slots || bt >= PB_SLOT_EXT || bu >= PB_SLOT_EXT;
            a->terms.push_back(PbTerm{T.field, bt, bu});
        }
        if (nlinear != 0 && nlinear != desc->nterms) return fail(PB200_EINVAL, "terms mix linear and bilinear slots");
        a->arity = nlinear ? 1 : 2;
        a->field_slots.assign(a->nfields, std::make_pair(-1, -1));
        for (const PbTerm& T : a->terms) {
            if (a->field_slots[T.field].first >= 0) return fail(PB200_EINVAL, "field %d is used by more than one term", T.field);
            a->field_slots[T.field] = std::make_pair(T.bt, T.bu);
        }
        for (int f = 0; f < a->nfields; ++f)
            if (a->field_slots[f].first < 0) return fail(PB200_EINVAL, "field %d is not used by any term", f);
        for (size_t i = 0; i < a->terms.size(); ++i)
            for (size_t j = i + 1; j < a->terms.size(); ++j)
                if (a->terms[i].bt == a->terms[j].bt && a->terms[i].bu == a->terms[j].bu)
                    return fail(PB200_EINVAL, "duplicate term (slots %d,%d): merge the coefficients", a->terms[i].bt, a->terms[i].bu);
    }

    // ---- upload ---------------------------------------------------------------------------------
    CK(pbSetDevice(device));
    struct Offs { size_t nodes, weights, fu, fv, Vu, Vv, V3u, V3v, pair02, pair12, rs, jm, su, sv, pi, pj, tr, ret; } offs[PB_MAXDIM];
    Pool& pool = a->pool;
    const int nd = 2;
    for (int k = 0; k < dim; ++k) {
        AxisHost& H = a->hax[k];
        Offs& o = offs[k];
        a->off_knots_u[k] = pool.put(H.U.kv);
        a->off_knots_v[k] = pool.put(H.V.kv);
        o.nodes = pool.put(H.nodes);
        o.weights = pool.put(H.weights);
        o.fu = pool.put(H.U.first);
        o.fv = pool.put(H.V.first);
        o.Vu = pool.put(nullptr, (size_t)H.G * nd * (H.U.p + 1) * sizeof(double));
        o.Vv = H.same ? o.Vu : pool.put(nullptr, (size_t)H.G * nd * (H.V.p + 1) * sizeof(double));
        o.V3u = o.V3v = o.pair02 = o.pair12 = 0;
        if (a->ext_slots) {
            // second / mixed derivatives: three-row tables for the per-entry path and two-row tables of the
            // order pairs (0,2), (1,2) for the walks; small (G rows), computed on the host
            auto three_rows = [&](const KvHost& Sp) {
                std::vector<double> T((size_t)H.G * 3 * (Sp.p + 1));
                for (int g = 0; g < H.G; ++g)
                    pb_basis_node(Sp.kv.data(), (int)Sp.kv.size(), Sp.p, H.nodes.data(), 3, nullptr, T.data(), g);
                return T;
            };
            const std::vector<double> Tu = three_rows(H.U);
            o.V3u = pool.put(Tu);
            o.V3v = H.same ? o.V3u : pool.put(three_rows(H.V));
            const int w = H.U.p + 1;
            std::vector<double> P02((size_t)H.G * 2 * w), P12((size_t)H.G * 2 * w);
            for (int g = 0; g < H.G; ++g)
                for (int c = 0; c < w; ++c) {
                    const double* r = &Tu[(size_t)g * 3 * w];
                    P02[(size_t)g * 2 * w + c] = r[c];     P02[(size_t)g * 2 * w + w + c] = r[2 * w + c];
                    P12[(size_t)g * 2 * w + c] = r[w + c]; P12[(size_t)g * 2 * w + w + c] = r[2 * w + c];
                }
            o.pair02 = pool.put(P02);
            o.pair12 = pool.put(P12);
        }
        o.rs = pool.put(H.row_start);
        o.jm = pool.put(H.jmin);
        o.su = pool.put(H.U.supp);
        o.sv = pool.put(H.V.supp);
        o.pi = pool.put(H.pair_i);
        o.pj = pool.put(H.pair_j);
        o.tr = pool.put(H.tr);
        o.ret = pool.put(H.ret_mu);
    }
    CK(pbMalloc((void**)&pool.dev, pool.host.size() + 256));
    CK(pbMemcpyH2D(pool.dev, pool.host.data(), pool.host.size(), (pbStream)stream));
    for (int k = 0; k < dim; ++k) {
        AxisHost& H = a->hax[k];
        PbAxis& D = a->dax[k];
        const Offs& o = offs[k];
        char* b = pool.dev;
        D.n = H.n; D.q = H.q; D.G = H.G;
        D.pu = H.U.p; D.pv = H.V.p; D.Nu = H.U.N(); D.Nv = H.V.N(); D.nd = nd; D.M = H.M;
        D.nodes = (const double*)(b + o.nodes);
        D.weights = (const double*)(b + o.weights);
        D.first_u = (const int*)(b + o.fu);
        D.first_v = (const int*)(b + o.fv);
        D.Vu = (const double*)(b + o.Vu);
        D.Vv = (const double*)(b + o.Vv);
        D.V3u = a->ext_slots ? (const double*)(b + o.V3u) : nullptr;
        D.V3v = a->ext_slots ? (const double*)(b + o.V3v) : nullptr;
        a->pair_tab[k][0] = D.Vu;
        a->pair_tab[k][1] = a->ext_slots ? (const double*)(b + o.pair02) : nullptr;
        a->pair_tab[k][2] = a->ext_slots ? (const double*)(b + o.pair12) : nullptr;
        D.row_start = (const int*)(b + o.rs);
        D.jmin = (const int*)(b + o.jm);
        D.supp_u = (const int*)(b + o.su);
        D.supp_v = (const int*)(b + o.sv);
        D.pair_i = (const int*)(b + o.pi);
        D.pair_j = (const int*)(b + o.pj);
        D.tr = (const int*)(b + o.tr);
        D.ret_mu = (const int*)(b + o.ret);
    }
    int rc = run_basis(a.get(), (pbStream)stream);
    if (rc) return rc;
    CK(pbStreamSync((pbStream)stream));   // pool.host may be released by the caller's thread later
    a->fast = detect_fast_path(a.get());
    if (a->fast && a->ext_slots && !ext_plan_ok(a.get())) a->fast = 0;      // the per-entry path serves it
    a->ml.device = device;
    a->ml.dim = dim;
    a->ml.row_start0 = a->hax[0].row_start;
    for (int k = 0; k < dim; ++k) {
        a->ml.Nv[k] = a->hax[k].V.N();
        a->ml.Nu[k] = a->hax[k].U.N();
        a->ml.M[k] = a->hax[k].M;
        a->ml.d_row_start[k] = a->dax[k].row_start;
        a->ml.d_jmin[k] = a->dax[k].jmin;
        a->ml.d_pair_i[k] = a->dax[k].pair_i;
        a->ml.d_pair_j[k] = a->dax[k].pair_j;
    }
    *out = a.release();
    return 0;
}

extern "C" const pb200_mlstruct* pb200_asm_mlstruct(const pb200_assembler* a) { return a ? &a->ml : nullptr; }

extern "C" int pb200_mlstruct_create(int nlevels, const int* rows, const int* cols, const int* nband,
                                     const uint32_t* const* h_bidx, int device, pb200_mlstruct** out) {
    if (!rows || !cols || !nband || !h_bidx || !out) return fail(PB200_EINVAL, "null argument");
    if (nlevels < 2 || nlevels > PB_MAXDIM) return fail(PB200_EUNSUPPORTED, "MLB kernels support 2 or 3 levels (got %d)", nlevels);
    std::unique_ptr<pb200_mlstruct> s(new pb200_mlstruct);
    s->device = device;
    s->dim = nlevels;
    Pool pool;
    size_t off[PB_MAXDIM][4];
    for (int k = 0; k < nlevels; ++k) {
        const int m = rows[k], M = nband[k];
        if (m < 1 || cols[k] < 1 || M < 1) return fail(PB200_EINVAL, "empty level %d", k);
        std::vector<int> rs(m + 1, 0), jm(m, 0), pi(M), pj(M);
        for (int e = 0; e < M; ++e) {
            pi[e] = (int)h_bidx[k][2 * e];
            pj[e] = (int)h_bidx[k][2 * e + 1];
            if (pi[e] >= m || pj[e] >= cols[k]) return fail(PB200_EINVAL, "index out of range in level %d", k);
            if (e > 0 && (pi[e] < pi[e - 1] || (pi[e] == pi[e - 1] && pj[e] != pj[e - 1] + 1)))
                return fail(PB200_EUNSUPPORTED, "level %d: rows must be sorted with contiguous column ranges", k);
            rs[pi[e] + 1]++;
        }
        for (int i = 0; i < m; ++i) rs[i + 1] += rs[i];
        for (int i = 0; i < m; ++i) jm[i] = rs[i + 1] > rs[i] ? pj[rs[i]] : 0;
        s->Nv[k] = m; s->Nu[k] = cols[k]; s->M[k] = M;
        if (k == 0) s->row_start0 = rs;
        off[k][0] = pool.put(rs); off[k][1] = pool.put(jm); off[k][2] = pool.put(pi); off[k][3] = pool.put(pj);
    }
    CK(pbSetDevice(device));
    CK(pbMalloc(&s->mem, pool.host.size() + 256));
    CK(pbMemcpyH2D(s->mem, pool.host.data(), pool.host.size(), (pbStream)0));
    CK(pbStreamSync((pbStream)0));
    for (int k = 0; k < nlevels; ++k) {
        char* b = (char*)s->mem;
        s->d_row_start[k] = (const int*)(b + off[k][0]);
        s->d_jmin[k] = (const int*)(b + off[k][1]);
        s->d_pair_i[k] = (const int*)(b + off[k][2]);
        s->d_pair_j[k] = (const int*)(b + off[k][3]);
    }
    *out = s.release();
    return 0;
}

extern "C" int pb200_mlstruct_destroy(pb200_mlstruct* s) {
    if (!s) return 0;
    if (s->mem) { pbSetDevice(s->device); pbFree(s->mem); }
    delete s;
    return 0;
}

extern "C" int pb200_asm_destroy(pb200_assembler* a) {
    if (!a) return 0;
    pbSetDevice(a->device);
    if (a->pool.dev) pbFree(a->pool.dev);
    if (a->geo_scratch) pbFree(a->geo_scratch);
    if (a->s32_keep_dev) pbFree(a->s32_keep_dev);
    delete a;
    return 0;
}

extern "C" int pb200_asm_tabulate(pb200_assembler* a, void* stream) {
    if (!a) return fail(PB200_EINVAL, "null handle");
    CK(pbSetDevice(a->device));
    return run_basis(a, (pbStream)stream);
}

extern "C" int pb200_asm_info(const pb200_assembler* a, pb200_info* info) {
    if (!a || !info) return fail(PB200_EINVAL, "null argument");
    memset(info, 0, sizeof *info);
    info->dim = a->dim;
    for (int k = 0; k < a->dim; ++k) {
        info->ndofs_test[k] = a->hax[k].V.N();
        info->ndofs_trial[k] = a->hax[k].U.N();
        info->nnodes[k] = a->hax[k].G;
        info->nband[k] = a->hax[k].M;
    }
    info->nfields = a->nfields;
    info->fast_path = a->fast;
    info->nnz = a->nnz;
    info->npoints = a->npts;
    return 0;
}

extern "C" int pb200_asm_structure(const pb200_assembler* a, int axis, uint32_t* h_bidx) {
    if (!a || !h_bidx || axis < 0 || axis >= a->dim) return fail(PB200_EINVAL, "invalid argument");
    const AxisHost& H = a->hax[axis];
    for (int m = 0; m < H.M; ++m) {
        h_bidx[2 * m] = (uint32_t)H.pair_i[m];
        h_bidx[2 * m + 1] = (uint32_t)H.pair_j[m];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K2: geometry tables + fields
// ------------------------------------------------------------------------------------------------
struct GeoTables {
    PbGeoDev dev;
    void* mem = nullptr;
};

// uploads knots + control net and evaluates the 1D geometry tables at the given device nodes
// `reuse` / `reuse_bytes`: an earlier allocation of the caller that is used again when it is large
// enough (work already queued on `st` that reads it is ordered before the new copies); otherwise it
// is released after a stream synchronisation and replaced.
static int build_geo_tables(const pb200_geo_desc* geo, int dim, const int* G, const double* const* d_nodes,
                            pbStream st, GeoTables& T, bool square = true, void** reuse = nullptr,
                            size_t* reuse_bytes = nullptr) {
    if (!geo) return fail(PB200_EINVAL, "geometry missing");
    if (geo->sdim != dim) return fail(PB200_EINVAL, "Geometry has wrong source dimension");
    if (square && geo->dim != dim) return fail(PB200_EINVAL, "Geometry has wrong dimension");
    if (geo->dim < 1 || geo->dim > 3) return fail(PB200_EUNSUPPORTED, "geometry with %d components not supported", geo->dim);
    if (!geo->h_coeffs) return fail(PB200_EINVAL, "geometry coefficients missing");
    const int nc = geo->dim + (geo->rational ? 1 : 0);
    size_t bytes = 0;
    auto reserve = [&](size_t b) { size_t o = (bytes + 255) & ~size_t(255); bytes = o + b; return o; };
    size_t off_k[PB_MAXDIM], off_f[PB_MAXDIM], off_v[PB_MAXDIM], off_c;
    long long ncoef = 1;
    // host data (knot vectors, control net) first and contiguous: ONE host->device copy per call
    for (int k = 0; k < dim; ++k) {
        if (geo->p[k] < 0 || geo->p[k] > PB_MAXP) return fail(PB200_EUNSUPPORTED, "geometry degree %d not supported", geo->p[k]);
        if (!geo->h_knots[k] || geo->nknots[k] < 2 * (geo->p[k] + 1)) return fail(PB200_EINVAL, "invalid geometry knot vector");
        off_k[k] = reserve(sizeof(double) * geo->nknots[k]);
        ncoef *= geo->nknots[k] - geo->p[k] - 1;
    }
    off_c = reserve(sizeof(double) * ncoef * nc);
    const size_t host_bytes = bytes;
    for (int k = 0; k < dim; ++k) {
        off_f[k] = reserve(sizeof(int) * G[k]);
        off_v[k] = reserve(sizeof(double) * G[k] * 2 * (geo->p[k] + 1));
    }
    if (reuse && *reuse && *reuse_bytes >= bytes + 256) {
        T.mem = *reuse;
    } else {
        if (reuse && *reuse) { CK(pbStreamSync(st)); CK(pbFree(*reuse)); *reuse = nullptr; *reuse_bytes = 0; }
        CK(pbMalloc(&T.mem, bytes + 256));
        if (reuse) { *reuse = T.mem; *reuse_bytes = bytes + 256; }
    }
    char* b = (char*)T.mem;
    PbGeoDev& D = T.dev;
    D.sdim = dim; D.dim = geo->dim; D.nc = nc; D.rational = geo->rational ? 1 : 0;
    {
        std::vector<char> stage(host_bytes, 0);
        for (int k = 0; k < dim; ++k) memcpy(stage.data() + off_k[k], geo->h_knots[k], sizeof(double) * geo->nknots[k]);
        memcpy(stage.data() + off_c, geo->h_coeffs, sizeof(double) * ncoef * nc);
        CK(pbMemcpyH2D(b, stage.data(), host_bytes, st));       // pageable source: staged by the driver before it returns
    }
    BasisJobs J;
    for (int k = 0; k < dim; ++k) {
        D.pg[k] = geo->p[k];
        D.Ng[k] = geo->nknots[k] - geo->p[k] - 1;
        D.gfirst[k] = (const int*)(b + off_f[k]);
        D.GV[k] = (const double*)(b + off_v[k]);
        J.add((const double*)(b + off_k[k]), geo->nknots[k], geo->p[k], d_nodes[k], G[k], 2, (int*)(b + off_f[k]),
              (double*)(b + off_v[k]));
    }
    k_basis_batch(J, st);
    for (int k = dim; k < PB_MAXDIM; ++k) { D.pg[k] = 0; D.Ng[k] = 1; D.gfirst[k] = nullptr; D.GV[k] = nullptr; }
    D.coeffs = (const double*)(b + off_c);
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_asm_bind_fields(pb200_assembler* a, double* d_fields) {
    if (!a) return fail(PB200_EINVAL, "null handle");
    a->d_fields = d_fields;
    a->fields_valid = d_fields != nullptr;      // the caller's buffer is taken as the fields until a new geometry is bound
    return 0;
}

// upload the spline geometry and tabulate its 1D basis functions at the Gauss nodes; the tables stay
// in a scratch buffer of the assembler that is kept between calls (no allocation, release or
// synchronisation in the steady state)
static int bind_geometry(pb200_assembler* a, const pb200_geo_desc* geo, pbStream st) {
    const double* d_nodes[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    int G[PB_MAXDIM] = {1, 1, 1};
    for (int k = 0; k < a->dim; ++k) { G[k] = a->hax[k].G; d_nodes[k] = a->dax[k].nodes; }
    GeoTables T;
    a->geo_valid = false;
    int rc = build_geo_tables(geo, a->dim, G, d_nodes, st, T, true, &a->geo_scratch, &a->geo_scratch_bytes);
    if (rc) return rc;
    a->geo_dev = T.dev;
    a->geo_valid = true;
    a->fields_valid = false;
    return 0;
}

extern "C" int pb200_asm_set_geometry(pb200_assembler* a, const pb200_geo_desc* geo, void* stream) {
    if (!a || !geo) return fail(PB200_EINVAL, "null argument");
    CK(pbSetDevice(a->device));
    return bind_geometry(a, geo, (pbStream)stream);
}

template <int DIM, class Prog>
static int launch_fields(const PbFieldParams& prm, pbStream st) {
    k_fields<DIM, Prog>(prm, st);
    CK(pbLastError());
    return 0;
}

struct GeneralSpec {
    int nphys = 0;
    const pb200_phys_term* phys = nullptr;
    int ninputs = 0;
    const double* const* d_inputs = nullptr;
};

static int compute_fields_impl(pb200_assembler* a, const pb200_geo_desc* geo, const double* d_jac, pbStream st,
                               int row0_begin = -1, int row0_end = -1, const GeneralSpec* gen = nullptr) {
    if (!a) return fail(PB200_EINVAL, "null handle");
    if (!a->d_fields) return fail(PB200_EINVAL, "no field buffer bound (pb200_asm_bind_fields)");
    if ((a->form == PB200_FORM_CUSTOM) != (gen != nullptr))
        return fail(PB200_EINVAL, a->form == PB200_FORM_CUSTOM ? "custom forms need pb200_asm_compute_fields_general"
                                                                : "built-in forms compute their own fields");
    CK(pbSetDevice(a->device));
    PbFieldParams prm;
    memset(&prm, 0, sizeof prm);
    if (gen) {
        if (gen->nphys < 1 || gen->nphys > PB_MAXPHYS || !gen->phys) return fail(PB200_EINVAL, "invalid number of coefficient terms %d", gen->nphys);
        if (gen->ninputs < 0 || gen->ninputs > PB_MAXPHYS) return fail(PB200_EINVAL, "invalid number of input arrays %d", gen->ninputs);
        prm.nphys = gen->nphys;
        for (int t = 0; t < gen->nphys; ++t) {
            const pb200_phys_term& T = gen->phys[t];
            const int lo = a->arity == 1 ? -1 : 0, hi = a->arity == 1 ? -1 : a->dim;
            if (T.slot_test < 0 || T.slot_test > a->dim || T.slot_trial < lo || T.slot_trial > hi || T.input >= gen->ninputs)
                return fail(PB200_EINVAL, "invalid coefficient term %d", t);
            prm.phys[t].bt = T.slot_test; prm.phys[t].bu = T.slot_trial;
            prm.phys[t].input = T.input < 0 ? -1 : T.input;
            prm.phys[t].scale = T.scale;
        }
        for (int i = 0; i < gen->ninputs; ++i) prm.inputs[i] = gen->d_inputs[i];
        for (int c = 0; c < a->nfields; ++c) { prm.outmap[c].bp = a->field_slots[c].first; prm.outmap[c].ap = a->field_slots[c].second; }
    }
    prm.dim = a->dim;
    const double* d_nodes[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    int G[PB_MAXDIM] = {1, 1, 1};
    for (int k = 0; k < a->dim; ++k) {
        prm.G[k] = G[k] = a->hax[k].G;
        prm.gw[k] = a->dax[k].weights;
        d_nodes[k] = a->dax[k].nodes;
    }
    prm.fields = a->d_fields;
    prm.npts = a->npts;
    prm.pt_begin = 0;
    prm.pt_end = a->npts;
    if (row0_begin >= 0) {      // only the Gauss planes in the support of the rows of the slab
        const AxisHost& H0 = a->hax[0];
        if (row0_end > H0.V.N() || row0_begin >= row0_end) return fail(PB200_EINVAL, "invalid row slab [%d,%d)", row0_begin, row0_end);
        const long long plane = a->npts / H0.G;
        prm.pt_begin = (long long)H0.V.supp[2 * row0_begin] * H0.q * plane;
        prm.pt_end = (long long)H0.V.supp[2 * (row0_end - 1) + 1] * H0.q * plane;
    }
    prm.nf = a->nfields;
    prm.jac_in = d_jac;
    if (!d_jac) {
        if (geo) {
            int rc = bind_geometry(a, geo, st);
            if (rc) return rc;
        } else if (!a->geo_valid) {
            return fail(PB200_EINVAL, "geometry missing");
        }
        prm.geo = a->geo_dev;
    } else {
        a->geo_valid = false;
    }
    (void)d_nodes;
    a->fields_valid = true;     // the launches below are ordered before any consumer on the stream
    const bool mass = a->form == PB200_FORM_MASS;
    if (gen) {
        if (!d_jac) {
            const long long lastG = a->hax[a->dim - 1].G;
            const long long r0 = prm.pt_begin / lastG, r1 = prm.pt_end / lastG;
            const bool rat = prm.geo.rational != 0;
            int rc;
            if (a->dim == 2) rc = rat ? k_fields_rows<2, 3, PbProgGeneral<2>>(prm, r0, r1, st) : k_fields_rows<2, 2, PbProgGeneral<2>>(prm, r0, r1, st);
            else rc = rat ? k_fields_rows<3, 4, PbProgGeneral<3>>(prm, r0, r1, st) : k_fields_rows<3, 3, PbProgGeneral<3>>(prm, r0, r1, st);
            if (rc) return rc;
            CK(pbLastError());
            return 0;
        }
        return a->dim == 2 ? launch_fields<2, PbProgGeneral<2>>(prm, st) : launch_fields<3, PbProgGeneral<3>>(prm, st);
    }
    if (!d_jac) {
        // spline geometry: row-wise sum-factorised evaluation
        const long long lastG = a->hax[a->dim - 1].G;
        const long long r0 = prm.pt_begin / lastG, r1 = prm.pt_end / lastG;
        const bool rat = prm.geo.rational != 0;
        int rc;
        if (a->dim == 2) {
            if (mass) rc = rat ? k_fields_rows<2, 3, PbProgMass<2>>(prm, r0, r1, st) : k_fields_rows<2, 2, PbProgMass<2>>(prm, r0, r1, st);
            else rc = rat ? k_fields_rows<2, 3, PbProgStiffness<2>>(prm, r0, r1, st) : k_fields_rows<2, 2, PbProgStiffness<2>>(prm, r0, r1, st);
        } else {
            if (mass) rc = rat ? k_fields_rows<3, 4, PbProgMass<3>>(prm, r0, r1, st) : k_fields_rows<3, 3, PbProgMass<3>>(prm, r0, r1, st);
            else rc = rat ? k_fields_rows<3, 4, PbProgStiffness<3>>(prm, r0, r1, st) : k_fields_rows<3, 3, PbProgStiffness<3>>(prm, r0, r1, st);
        }
        if (rc) return rc;
        CK(pbLastError());
        return 0;
    }
    if (a->dim == 2)
        return mass ? launch_fields<2, PbProgMass<2>>(prm, st) : launch_fields<2, PbProgStiffness<2>>(prm, st);
    return mass ? launch_fields<3, PbProgMass<3>>(prm, st) : launch_fields<3, PbProgStiffness<3>>(prm, st);
}

// `geo` may be null in the two calls below when a geometry has been bound with pb200_asm_set_geometry
extern "C" int pb200_asm_compute_fields(pb200_assembler* a, const pb200_geo_desc* geo, void* stream) {
    return compute_fields_impl(a, geo, nullptr, (pbStream)stream);
}
extern "C" int pb200_asm_compute_fields_slab(pb200_assembler* a, const pb200_geo_desc* geo, int row0_begin, int row0_end,
                                             void* stream) {
    return compute_fields_impl(a, geo, nullptr, (pbStream)stream, row0_begin, row0_end);
}
extern "C" int pb200_asm_compute_fields_general(pb200_assembler* a, const pb200_geo_desc* geo, const double* d_jac,
                                                int nphys, const pb200_phys_term* phys, int ninputs,
                                                const double* const* d_inputs, int row0_begin, int row0_end, void* stream) {
    GeneralSpec g;
    g.nphys = nphys; g.phys = phys; g.ninputs = ninputs; g.d_inputs = d_inputs;
    if (!geo && !d_jac) return fail(PB200_EINVAL, "geometry missing");
    return compute_fields_impl(a, d_jac ? nullptr : geo, d_jac, (pbStream)stream, row0_begin, row0_end, &g);
}
extern "C" int pb200_asm_compute_fields_from_jacobian(pb200_assembler* a, const double* d_jac, void* stream) {
    if (!d_jac) return fail(PB200_EINVAL, "null Jacobian array");
    return compute_fields_impl(a, nullptr, d_jac, (pbStream)stream);
}

extern "C" int pb200_geo_eval_grid(const pb200_geo_desc* geo, const int* npts, const double* const* h_grid,
                                   double* d_values, double* d_jac, int device, void* stream) {
    if (!geo || !npts || !h_grid) return fail(PB200_EINVAL, "null argument");
    const int dim = geo->sdim;
    if (dim < 2 || dim > 3) return fail(PB200_EUNSUPPORTED, "geometry evaluation implemented for 2 and 3 parameters");
    if (geo->dim > 3) return fail(PB200_EUNSUPPORTED, "at most 3 output components");
    CK(pbSetDevice(device));
    pbStream st = (pbStream)stream;
    double* d_nodes_mem = nullptr;
    size_t tot = 0;
    for (int k = 0; k < dim; ++k) tot += npts[k];
    CK(pbMalloc((void**)&d_nodes_mem, sizeof(double) * tot));
    const double* d_nodes[PB_MAXDIM] = {nullptr, nullptr, nullptr};
    size_t o = 0;
    long long total = 1;
    for (int k = 0; k < dim; ++k) {
        CK(pbMemcpyH2D(d_nodes_mem + o, h_grid[k], sizeof(double) * npts[k], st));
        d_nodes[k] = d_nodes_mem + o;
        o += npts[k];
        total *= npts[k];
    }
    GeoTables T;
    int rc = build_geo_tables(geo, dim, npts, d_nodes, st, T, /*square=*/false);
    if (rc == 0) {
        if (dim == 2) k_geo_grid<2>(T.dev, npts, total, d_values, d_jac, st);
        else k_geo_grid<3>(T.dev, npts, total, d_values, d_jac, st);
        pbError e = pbLastError();
        if (e != pbSuccess) rc = fail(PB200_ECUDA, "geometry kernel launch failed: %s", pbErrorString(e));
    }
    pbStreamSync(st);
    if (T.mem) pbFree(T.mem);
    pbFree(d_nodes_mem);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// K3: the sum-factorised pipeline
// ------------------------------------------------------------------------------------------------
struct Slab {
    int ra, rb;             // rows of axis 0
    int ea, eb;             // extended rows (transposed partners)
    int sa, sb;             // spans of axis 0 covered
    int mu_lo, mu_hi;       // band range of the rows
    int ext_lo, ext_hi;     // band range of the extended rows
};

static int make_slab(const pb200_assembler* a, int ra, int rb, bool sym, Slab& S) {
    const AxisHost& H = a->hax[0];
    const int N = H.V.N();
    if (ra < 0 || rb > N || ra >= rb) return fail(PB200_EINVAL, "invalid row slab [%d,%d) of %d rows", ra, rb, N);
    S.ra = ra; S.rb = rb;
    const int P = H.U.p;
    S.ea = sym ? std::max(0, ra - P) : ra;
    S.eb = sym ? std::min(N, rb + P) : rb;
    S.sa = H.V.supp[2 * ra];
    S.sb = H.V.supp[2 * (rb - 1) + 1];
    S.mu_lo = H.row_start[ra]; S.mu_hi = H.row_start[rb];
    S.ext_lo = H.row_start[S.ea]; S.ext_hi = H.row_start[S.eb];
    return 0;
}

static bool uses_transposes(const pb200_assembler* a) { return a->form == PB200_FORM_STIFFNESS; }

// 3D mass / stiffness on a bound spline geometry: stage 1 evaluates geometry and fields itself
// (walk_geo.cuh) — no K2 launch, no field buffer
// Shared memory of the fused stage 1: the 1D tables of the walked spans are staged per block next to the
// thread-private columns, and two blocks must fit an SM.  Long axes are therefore walked in pieces (grid.y; the
// pieces overlap by p spans): the smallest number of pieces for `nspans` spans, 0 if even PB_WALK_MAXSPLIT do not fit.
static int s1f_pieces(const pb200_assembler* a, int nspans) {
    const PbGeoDev& g = a->geo_dev;
    const AxisHost& H = a->hax[0];
    const int P = H.U.p, Q = H.q;
    for (int K = 1; K <= PB_WALK_MAXSPLIT; ++K) {
        const size_t sp = (size_t)std::min(nspans, (nspans + K - 1) / K + (K > 1 ? P + 1 : 0));
        const size_t nodes = sp * Q;
        const size_t smem = nodes * 2 * (P + 1) * 8 + (sp + 4 + (sp + P) * (2 * P + 1)) * 4
                            + nodes * (2 * (g.pg[0] + 1) + 1) * 8 + nodes * 4
                            + ((size_t)g.Ng[0] * ((g.nc * 3 + 1) & ~1) + (size_t)std::max(Q * 6, (P + 1) * 6) + 2) * 128 * 8 + 1024;
        if (smem <= 113 * 1024) return K;
    }
    return 0;
}

static bool fused_stage1(const pb200_assembler* a) {
    if (!a->fuse || a->dim != 3 || !a->geo_valid || a->arity != 2 || !a->fast || a->force_walk) return false;
    if (a->form != PB200_FORM_STIFFNESS && a->form != PB200_FORM_MASS) return false;
    if (!a->lane_ok[0] || !a->walk_rot) return false;           // rotating-window walk on axis 0
    const PbGeoDev& g = a->geo_dev;
    if (g.dim != 3 || g.sdim != 3 || g.Ng[0] * ((g.nc * 3 + 1) & ~1) > PB_GEO_ZMAX) return false;
    const AxisHost& H = a->hax[0];
    const int P = H.U.p, Q = H.q;
    // (degree 4: the six (p+1)^2 windows of the stiffness form do not fit the register file — ptxas spills; the single
    // symmetric window of the mass form does)
    if (P > 4 || (P > 3 && a->form != PB200_FORM_MASS)) return false;
    if (!have_plan(a->form == PB200_FORM_STIFFNESS ? PB_PLAN_S1F : PB_PLAN_S1F_MASS, P, Q)) return false;
    return s1f_pieces(a, H.n) > 0;
}

extern "C" int pb200_asm_uses_fused_fields(const pb200_assembler* a) { return a && fused_stage1(a) ? 1 : 0; }

// 3D mass / stiffness with equal degrees and single interior knots on axes 1 and 2: stages 2 and 3 run
// as one kernel (fused23.cuh); X2 is never written
static void fill_s32_axes(const pb200_assembler* a, PbS32Params& p) {
    const AxisHost &H1 = a->hax[1], &H2 = a->hax[2];
    const PbAxis &D0 = a->dax[0], &D1 = a->dax[1], &D2 = a->dax[2];
    p.pair_i0 = D0.pair_i; p.pair_j0 = D0.pair_j; p.tr0 = D0.tr;
    p.G1 = H1.G; p.G2 = H2.G; p.n1 = H1.n; p.n2 = H2.n;
    p.N1 = H1.V.N(); p.N2 = H2.V.N(); p.M1 = H1.M; p.M2 = H2.M;
    p.first1 = D1.first_u; p.V1 = D1.Vu; p.ret_mu1 = D1.ret_mu; p.tr1 = D1.tr; p.pair_i1 = D1.pair_i;
    p.first2 = D2.first_u; p.V2 = D2.Vu; p.ret_mu2 = D2.ret_mu; p.tr2 = D2.tr;
}

// nodes of axis 1 a block of the fused stages 2 + 3 walks when the axis is cut into K pieces (upper bound)
static int s32_piece_rows(const pb200_assembler* a, int K) {
    const AxisHost& H1 = a->hax[1];
    if (K <= 1) return H1.G;
    return std::min(H1.n, (H1.n + K - 1) / K + H1.U.p + 2) * H1.q;
}

// Tail effect of the fused stages 2 + 3: one block per SM and equal tasks — a slab whose tasks fill the GPU 2.01
// times pays for 3 waves.  Two remedies, whichever gives the shorter sum of waves (in span steps of axis 1):
//   * cut axis 1 into K pieces for ALL tasks (K x as many, shorter blocks, p spans of overlap each);
//   * run the full waves of whole tasks first and cut only the remainder, so that it fits one more, short wave.
static void choose_s32_pieces(const pb200_assembler* a, PbS32Params& q) {
    const AxisHost& H1 = a->hax[1];
    const long long tasks = pb_s32_tasks(q);
    const int P = H1.U.p, nsp = H1.n, N1 = H1.V.N(), S = a->sm_count;
    int K = 1;
    long long n_whole = 0;
    if (a->walk_split > 1) {
        K = std::min(a->walk_split, PB_S32_MAXPIECE);
        if (a->s32_whole >= 0) n_whole = std::min<long long>(tasks, a->s32_whole);
    } else if (a->walk_split == 0 && S > 0 && tasks > 0) {
        double best = (double)((tasks + S - 1) / S) * nsp;
        for (int k = 2; k <= 4; ++k) {
            if (nsp / k < 4 * (P + 1)) break;
            const double cost = (double)((k * tasks + S - 1) / S) * ((double)nsp / k + P + 2);
            if (cost < 0.95 * best) { best = cost; K = k; }
        }
        const long long w = tasks / S, r = tasks - w * S;
        if (w >= 1 && r > 0) {
            int k = (int)std::min<long long>(PB_S32_MAXPIECE, S / r);
            while (k > 1 && nsp / k < 4 * (P + 1)) --k;
            if (k > 1) {
                const double cost = (double)w * nsp + ((double)nsp / k + P + 2);
                if (cost < 0.97 * best) { best = cost; K = k; n_whole = w * S; }
            }
        }
    }
    K = std::min(K, N1);
    // long axes: the staged axis-1 table must fit shared memory — cut ALL tasks into at least Kfit pieces
    {
        PbS32Launch fn = pb_find_s32(a->form, H1.U.p, H1.q);
        int Kfit = 1;
        for (; fn && Kfit < PB_S32_MAXPIECE; ++Kfit) {
            PbS32Params t = q;
            t.out = nullptr;
            t.v1_rows = s32_piece_rows(a, Kfit);
            if (fn(&t, nullptr) == 0) break;
        }
        if (Kfit > 1) { K = std::max(K, Kfit); n_whole = 0; }
    }
    q.npiece = 0;
    q.n_whole = 0;
    q.v1_rows = H1.G;
    if (K > 1) {
        q.npiece = K;
        q.n_whole = (int)n_whole;
        for (int y = 0; y < K; ++y) {
            const int lo = (int)((long long)N1 * y / K), hi = (int)((long long)N1 * (y + 1) / K);
            q.pw_lo[y] = lo; q.pw_hi[y] = hi;
            q.ps_begin[y] = H1.V.supp[2 * lo];
            q.ps_end[y] = H1.V.supp[2 * (hi - 1) + 1];
        }
        if (q.n_whole == 0) {       // no task walks the whole axis: stage only the longest piece
            int longest = 0;
            for (int y = 0; y < K; ++y) longest = std::max(longest, q.ps_end[y] - q.ps_begin[y]);
            q.v1_rows = longest * H1.q;
        }
    }
    static const bool debug_split = getenv("PB200_DEBUG_SPLIT") != nullptr;
    if (debug_split) fprintf(stderr, "[pb200] fused stage 2+3: %lld tasks, %d pieces, %d tasks unsplit\n", tasks, q.npiece, q.n_whole);
}

// tasks of a fused stage-2+3 launch: the kept band entries of axis 0 (uploaded once per slab) and the packing of
// the short last batches (fused23.cuh)
static int fill_s32_tasks(pb200_assembler* a, const Slab& S, PbS32Params& q, pbStream st) {
    const AxisHost &H0 = a->hax[0], &H2 = a->hax[2];
    const long long key[5] = {S.mu_lo, S.mu_hi, S.ra, S.rb, q.symmetric};
    if (memcmp(key, a->s32_keep_key, sizeof key) != 0 || !a->s32_keep_dev) {
        // (the host copy is a member: an asynchronous copy from pageable memory may still read it after this call)
        std::vector<int>& keep = a->s32_keep_host[a->s32_keep_flip ^= 1];
        keep.clear();
        for (int mu = S.mu_lo; mu < S.mu_hi; ++mu) {
            const int i0 = H0.pair_i[mu], j0 = H0.pair_j[mu];
            if (!(q.symmetric && j0 >= S.ra && j0 < S.rb && j0 < i0)) keep.push_back(mu);
        }
        if (keep.size() > a->s32_keep_cap) {
            if (a->s32_keep_dev) pbFree(a->s32_keep_dev);
            a->s32_keep_dev = nullptr;
            a->s32_keep_cap = std::max(keep.size(), (size_t)H0.M) + 64;       // every slab of this assembler fits
            CK(pbMalloc((void**)&a->s32_keep_dev, a->s32_keep_cap * sizeof(int)));
        }
        if (!keep.empty()) CK(pbMemcpyH2D(a->s32_keep_dev, keep.data(), keep.size() * sizeof(int), st));
        a->s32_nkeep = (int)keep.size();
        memcpy(a->s32_keep_key, key, sizeof key);
    }
    q.keep = a->s32_keep_dev;
    q.nkeep = a->s32_nkeep;
    const int P = H2.U.p;
    const int sb_tail = (q.nbatch - 1) * (32 - P);
    const int wt = std::min(32, H2.n + P - sb_tail);       // lanes the last batch needs
    q.tail_w = wt;
    q.tail_k = (a->pack_tails && q.nbatch >= 2 && wt > 0) ? std::min(4, 32 / wt) : 1;
    if (q.tail_k < 2) q.tail_k = 1;
    return 0;
}
// generic forms: a term travels through the stages as (test slot, trial slot, buffer slot)
struct GenTerm { int bt, bu, slot; };
struct GenStage {
    std::vector<GenTerm> out;                    // outputs (slot = index in the stage's output buffer)
    std::vector<std::array<int, 9>> ops9;        // per output: input slot for derivative orders (test, trial) = 3*ft + fu, or -1
    std::vector<std::array<int, 4>> ops;         // the same by ROWS (rt, ru) = 2*rt + ru of the two-row table `tab`
    std::vector<int> tab;                        // per output: table of the orders (0,1) / (0,2) / (1,2) on this axis
    bool ok = true;                              // false: an output needs all three orders of this axis (no walk for that)
};
// rows of the two-row tables: orders {0,1} -> table 0, {0,2} -> table 1, {1,2} -> table 2
static bool pair_table(unsigned mask, int& tab, int row[3]) {
    row[0] = row[1] = row[2] = -1;
    if (!(mask & 4u)) { tab = 0; row[0] = 0; row[1] = 1; return true; }
    if (!(mask & 2u)) { tab = 1; row[0] = 0; row[2] = 1; return true; }
    if (!(mask & 1u)) { tab = 2; row[1] = 0; row[2] = 1; return true; }
    return false;
}
// contract `axis`: the derivative order along it leaves the slot
static GenStage plan_generic_stage(const std::vector<GenTerm>& in, int axis) {
    GenStage S;
    for (const GenTerm& t : in) {
        const int ft = pb_slot_order(t.bt, axis), fu = pb_slot_order(t.bu, axis);
        const int bt = pb_slot_contract(t.bt, axis), bu = pb_slot_contract(t.bu, axis);
        size_t o = 0;
        for (; o < S.out.size(); ++o)
            if (S.out[o].bt == bt && S.out[o].bu == bu) break;
        if (o == S.out.size()) {
            S.out.push_back(GenTerm{bt, bu, (int)o});
            S.ops9.push_back({-1, -1, -1, -1, -1, -1, -1, -1, -1});
        }
        S.ops9[o][ft * 3 + fu] = t.slot;
    }
    for (size_t o = 0; o < S.out.size(); ++o) {
        unsigned mask = 0;
        for (int c = 0; c < 9; ++c)
            if (S.ops9[o][c] >= 0) mask |= (1u << (c / 3)) | (1u << (c % 3));
        int tab = 0, row[3];
        S.ok = pair_table(mask, tab, row) && S.ok;
        std::array<int, 4> r4 = {-1, -1, -1, -1};
        for (int c = 0; c < 9; ++c)
            if (S.ops9[o][c] >= 0 && row[c / 3] >= 0 && row[c % 3] >= 0) r4[row[c / 3] * 2 + row[c % 3]] = S.ops9[o][c];
        S.ops.push_back(r4);
        S.tab.push_back(tab);
    }
    return S;
}
static bool generic_plan(const pb200_assembler* a, std::vector<GenStage>& stages) {
    std::vector<GenTerm> cur;
    bool ok = true;
    for (const PbTerm& t : a->terms) cur.push_back(GenTerm{t.bt, t.bu, t.field});
    for (int k = 0; k < a->dim; ++k) {
        stages.push_back(plan_generic_stage(cur, k));
        ok = ok && stages.back().ok;
        cur = stages.back().out;
    }
    return ok;
}

static bool fused_stage23(const pb200_assembler* a) {
    if (!a->fuse23 || a->dim != 3 || a->arity != 2 || !a->fast || a->force_walk) return false;
    if (a->form == PB200_FORM_CUSTOM) {
        if (!a->same_space) return false;       // general forms: no symmetry is used, every band entry is computed
        if (a->ext_slots) {                     // second / mixed derivatives on axes 1, 2: the unfused walks
            std::vector<GenStage> stages;
            generic_plan(a, stages);
            for (const GenTerm& t : stages[0].out)
                if (t.bt >= PB_SLOT_EXT || t.bu >= PB_SLOT_EXT) return false;
        }
    } else {
        if (a->form != PB200_FORM_STIFFNESS && a->form != PB200_FORM_MASS) return false;
        if (!(a->mirror_opt && a->symmetric && a->same_space)) return false;
    }
    if (!a->lane_ok[1] || !a->lane_ok[2] || !a->walk_rot) return false;
    const AxisHost &H1 = a->hax[1], &H2 = a->hax[2];
    if (H1.U.p != H2.U.p || H1.q != H2.q) return false;
    if ((long long)H1.M * H2.M >= (1LL << 31)) return false;
    PbS32Launch fn = pb_find_s32(a->form, H1.U.p, H1.q);
    if (!fn) return false;
    PbS32Params q;
    memset(&q, 0, sizeof q);
    fill_s32_axes(a, q);
    q.v1_rows = s32_piece_rows(a, PB_S32_MAXPIECE);
    return fn(&q, nullptr) == 0;        // query: shared memory fits, if need be with axis 1 cut into pieces
}

static void stage_sizes(const pb200_assembler* a, const Slab& S, size_t& x1_terms, size_t& x1_stride, size_t& x2_terms,
                        size_t& x2_stride) {
    const size_t Mext = (size_t)(S.ext_hi - S.ext_lo);
    const bool st = a->form == PB200_FORM_STIFFNESS;
    if (a->form == PB200_FORM_CUSTOM) {
        std::vector<GenStage> stages;
        generic_plan(a, stages);
        x1_terms = stages[0].out.size();
        x1_stride = Mext * a->hax[1].G * (a->dim == 3 ? a->hax[2].G : 1);
        x2_terms = a->dim == 3 ? stages[1].out.size() : 0;
        x2_stride = a->dim == 3 ? Mext * a->hax[1].M * a->hax[2].G : 0;
        if (a->dim == 3 && fused_stage23(a)) { x2_terms = 0; x2_stride = 0; }
        return;
    }
    if (a->dim == 2) {
        x1_terms = st ? 3 : 1;
        x1_stride = Mext * a->hax[1].G;
        x2_terms = 0; x2_stride = 0;
    } else {
        x1_terms = st ? 6 : 1;
        x1_stride = Mext * a->hax[1].G * a->hax[2].G;
        x2_terms = st ? 3 : 1;
        x2_stride = Mext * a->hax[1].M * a->hax[2].G;
        if (fused_stage23(a)) { x2_terms = 0; x2_stride = 0; }
    }
}

extern "C" int pb200_asm_workspace_bytes(const pb200_assembler* a, int row0_begin, int row0_end, size_t* bytes) {
    if (!a || !bytes) return fail(PB200_EINVAL, "null argument");
    if (!a->fast || a->arity != 2) { *bytes = 0; return 0; }
    Slab S;
    int rc = make_slab(a, row0_begin, row0_end, uses_transposes(a), S);
    if (rc) return rc;
    size_t t1, s1, t2, s2;
    stage_sizes(a, S, t1, s1, t2, s2);
    *bytes = (t1 * s1 + t2 * s2) * sizeof(double) + 512;
    return 0;
}

static int run_stage(int plan, pb200_assembler* a, int axis, PbWalkParams& prm, pbStream st, const char* name) {
    mark_stage(a, name, st);
    ++g_launches;
    const AxisHost& H = a->hax[axis];
    const PbAxis& D = a->dax[axis];
    const int P = H.U.p, Q = H.q;
    PbWalkLaunch fn = pb_find_walk(plan, P, Q);
    if (!fn) return fail(PB200_EUNSUPPORTED, "no walk kernel for plan %d, p=%d, q=%d", plan, P, Q);
    prm.N = H.V.N();
    prm.f_lo = H.U.first[prm.s_begin];
    prm.f_hi = std::min(H.V.N(), H.U.first[prm.s_end - 1] + P + 1);
    prm.first = D.first_u;
    prm.V2 = a->walk_table ? a->walk_table : D.Vu;      // generic forms choose the table of the derivative orders they need
    prm.ret_mu = D.ret_mu;
    prm.regular = (a->lane_ok[axis] && a->walk_rot) ? 1 : 0;
    // final stages (node axis contiguous, one output, single interior knots): warp-per-line kernel
    bool nofilter = true;
    for (int o = 0; o < PB_WALK_MAXOUT; ++o) nofilter = nofilter && prm.w_mode[o] == 0;
    if (prm.in_sc == 1 && prm.out_smu == 1 && nofilter && a->lane_ok[axis] && !a->force_walk) {
        PbWalkLaunch lane = pb_find_walk(PB_PLAN_LANE_BASE + plan, P, Q);
        if (lane) {
            int e = lane(&prm, a->lane_v1 ? -16 : a->lane_lines, 0, st);
            if (e) return fail(PB200_ECUDA, "lane-span kernel launch failed (plan %d, p=%d, q=%d): %s", plan, P, Q, pbErrorString((pbError)e));
            return 0;
        }
    }
    const size_t smem = (size_t)(prm.s_end - prm.s_begin) * Q * 2 * (P + 1) * sizeof(double);
    const size_t ismem = ((size_t)(prm.s_end - prm.s_begin + 4) + (size_t)(prm.f_hi - prm.f_lo) * (2 * P + 1)) * sizeof(int);
    const int use_smem = smem + ismem <= 64 * 1024;     // longer axes read the table through L1 (keeps occupancy)
    // Tail effect: a stage whose blocks fill the GPU 1.x times spends the second wave almost idle.
    // Cut the walk axis into K pieces (grid.y) when that shortens the sum of the waves; the pieces
    // overlap by p spans.  Needs a stage without a walk-axis filter of its own.
    prm.nsplit = 0;
    // rows the pieces partition: all rows of the axis, or the (extended) slab of a filtered stage
    const int r_lo = nofilter ? 0 : prm.w_ext_lo, r_hi = nofilter ? H.V.N() : prm.w_ext_hi;
    // the fused stage 1 needs its staged tables to fit shared memory: a lower bound on the pieces
    const int Kmin = (plan == PB_PLAN_S1F || plan == PB_PLAN_S1F_MASS) ? std::max(1, s1f_pieces(a, prm.s_end - prm.s_begin)) : 1;
    if (r_hi > r_lo && (a->walk_split != 1 || Kmin > 1)) {
        const int nsp = prm.s_end - prm.s_begin;
        int K = 1;
        if (a->walk_split == 1) {
            K = 1;
        } else if (a->walk_split > 1) {
            K = a->walk_split;
        } else {
            const int occ = fn(&prm, -1, use_smem ? smem + ismem + 256 : 0, st);
            if (occ > 0 && a->sm_count > 0) {
                const long long B = (prm.nthreads + 127) / 128, slots = (long long)a->sm_count * occ;
                double best = (double)((B + slots - 1) / slots) * nsp;
                for (int k = 2; k <= PB_WALK_MAXSPLIT; ++k) {
                    if (nsp / k < 4 * (P + 1)) break;                  // pieces too short to pay for the overlap
                    // measured on B200 (3D p=3 n=128, stage 2): two pieces are ~10 % faster than the wave
                    // count alone predicts (shorter-lived blocks balance better), hence the 0.9
                    const double cost = 0.9 * (double)((k * B + slots - 1) / slots) * ((double)nsp / k + P);
                    if (cost < 0.93 * best) { best = cost; K = k; }
                }
            }
        }
        K = std::min(std::max(K, Kmin), std::min(PB_WALK_MAXSPLIT, r_hi - r_lo));
        static const bool debug_split = getenv("PB200_DEBUG_SPLIT") != nullptr;
        if (K > 1 && debug_split) fprintf(stderr, "[pb200] stage %s: walk axis cut into %d pieces\n", name, K);
        if (K > 1) {
            prm.nsplit = K;
            for (int y = 0; y < K; ++y) {
                const int lo = r_lo + (int)((long long)(r_hi - r_lo) * y / K), hi = r_lo + (int)((long long)(r_hi - r_lo) * (y + 1) / K);
                prm.sp_w_lo[y] = lo; prm.sp_w_hi[y] = hi;
                // spans the rows of the piece see, inside the span range of the stage
                int sb = std::max(prm.s_begin, H.V.supp[2 * lo]), se = std::min(prm.s_end, H.V.supp[2 * (hi - 1) + 1]);
                if (se <= sb) { sb = prm.s_begin; se = prm.s_begin + 1; }      // nothing to do: walk one span, retire nothing new
                prm.sp_s_begin[y] = sb; prm.sp_s_end[y] = se;
                prm.sp_f_lo[y] = H.U.first[sb];
                prm.sp_f_hi[y] = std::min(H.V.N(), H.U.first[se - 1] + P + 1);
            }
        }
    }
    int e = fn(&prm, use_smem, use_smem ? smem : 0, st);
    if (e) return fail(PB200_ECUDA, "walk kernel launch failed (plan %d, p=%d, q=%d): %s", plan, P, Q, pbErrorString((pbError)e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// linear forms: load vector by three single-function walks
// ------------------------------------------------------------------------------------------------
struct Gen1Stage {
    std::vector<int> out_slot;
    std::vector<std::array<int, 3>> ops3;       // per output: input for test derivative order 0 / 1 / 2 on this axis
    std::vector<std::array<int, 2>> ops;        // the same by rows of the two-row table `tab`
    std::vector<int> tab;
    bool ok = true;
};

static bool plan_linear(const pb200_assembler* a, std::vector<Gen1Stage>& stages) {
    std::vector<std::pair<int, int>> cur;       // (test slot, buffer slot)
    bool ok = true;
    for (const PbTerm& t : a->terms) cur.push_back(std::make_pair(t.bt, t.field));
    for (int k = 0; k < a->dim; ++k) {
        Gen1Stage S;
        std::vector<std::pair<int, int>> next;
        for (auto& t : cur) {
            const int ft = pb_slot_order(t.first, k);
            const int bt = pb_slot_contract(t.first, k);
            size_t o = 0;
            for (; o < S.out_slot.size(); ++o)
                if (S.out_slot[o] == bt) break;
            if (o == S.out_slot.size()) { S.out_slot.push_back(bt); S.ops3.push_back({-1, -1, -1}); next.push_back(std::make_pair(bt, (int)o)); }
            S.ops3[o][ft] = t.second;
        }
        for (size_t o = 0; o < S.out_slot.size(); ++o) {
            unsigned mask = 0;
            for (int c = 0; c < 3; ++c)
                if (S.ops3[o][c] >= 0) mask |= 1u << c;
            int tab = 0, row[3];
            S.ok = pair_table(mask, tab, row) && S.ok;
            std::array<int, 2> r2 = {-1, -1};
            for (int c = 0; c < 3; ++c)
                if (S.ops3[o][c] >= 0 && row[c] >= 0) r2[row[c]] = S.ops3[o][c];
            S.ops.push_back(r2);
            S.tab.push_back(tab);
        }
        ok = ok && S.ok;
        stages.push_back(S);
        cur = next;
    }
    return ok;
}

// forms with second / mixed derivatives: every stage output must get by with two derivative orders per axis
static bool ext_plan_ok(const pb200_assembler* a) {
    if (a->arity == 1) {
        std::vector<Gen1Stage> st;
        return plan_linear(a, st);
    }
    std::vector<GenStage> st;
    return generic_plan(a, st);
}

extern "C" int pb200_asm_vector_workspace_bytes(const pb200_assembler* a, size_t* bytes) {
    if (!a || !bytes) return fail(PB200_EINVAL, "null argument");
    if (a->arity != 1) return fail(PB200_EINVAL, "not a linear form");
    size_t y1 = 1, y2 = 1;
    y1 = (size_t)a->hax[0].V.N() * a->hax[1].G * (a->dim == 3 ? a->hax[2].G : 1);
    y2 = a->dim == 3 ? (size_t)a->hax[0].V.N() * a->hax[1].V.N() * a->hax[2].G : 0;
    *bytes = (y1 * (a->dim + 1) + y2 * a->dim) * sizeof(double) + 512;
    return 0;
}

extern "C" int pb200_asm_assemble_vector(pb200_assembler* a, double* d_out, void* d_work, size_t work_bytes, void* stream) {
    if (!a || !d_out) return fail(PB200_EINVAL, "null argument");
    if (a->arity != 1) return fail(PB200_EINVAL, "assemble_vector needs a linear form (arity 1)");
    if (!a->d_fields) return fail(PB200_EINVAL, "fields have not been computed");
    if (!a->fast) return fail(PB200_EUNSUPPORTED, "no vector kernels for these degrees");
    CK(pbSetDevice(a->device));
    pbStream st = (pbStream)stream;
    size_t need = 0;
    int rc = pb200_asm_vector_workspace_bytes(a, &need);
    if (rc) return rc;
    if (!d_work || work_bytes < need) return fail(PB200_ENOMEM, "workspace too small: need %zu bytes, have %zu", need, work_bytes);
    const int dim = a->dim;
    const long long N0 = a->hax[0].V.N(), N1 = a->hax[1].V.N(), G1 = a->hax[1].G;
    const long long G2 = dim == 3 ? a->hax[2].G : 1, N2 = dim == 3 ? a->hax[2].V.N() : 1;
    const size_t y1 = (size_t)N0 * G1 * G2, y2 = dim == 3 ? (size_t)N0 * N1 * G2 : 0;
    double* Y1 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(d_work) + 255) & ~uintptr_t(255));
    double* Y2 = Y1 + y1 * (dim + 1);
    std::vector<Gen1Stage> stages;
    plan_linear(a, stages);
    for (int k = 0; k < dim; ++k) {
        const AxisHost& H = a->hax[k];
        PbWalk1Launch fn = pb_find_walk1(H.U.p, H.q);
        if (!fn) return fail(PB200_EUNSUPPORTED, "no vector kernel for p=%d, q=%d", H.U.p, H.q);
        const bool last = k == dim - 1;
        const double* inbase = k == 0 ? a->d_fields : (k == 1 ? Y1 : Y2);
        const long long instride = k == 0 ? a->npts : (k == 1 ? (long long)y1 : (long long)y2);
        double* outbase = last ? d_out : (k == 0 ? Y1 : Y2);
        const long long outstride = last ? 0 : (k == 0 ? (long long)y1 : (long long)y2);
        for (size_t o = 0; o < stages[k].out_slot.size(); ++o) {
            PbWalk1Params p;
            memset(&p, 0, sizeof p);
            p.in0 = stages[k].ops[o][0] >= 0 ? inbase + (long long)stages[k].ops[o][0] * instride : nullptr;
            p.in1 = stages[k].ops[o][1] >= 0 ? inbase + (long long)stages[k].ops[o][1] * instride : nullptr;
            p.out = outbase + (long long)o * outstride;
            p.n = H.n; p.N = H.V.N();
            p.first = a->dax[k].first_u; p.V2 = a->pair_tab[k][stages[k].tab[o]];
            if (k == 0) {                           // F[g0][g1][g2] -> Y1[i0][g1][g2]
                p.X = (int)(G1 * G2); p.nthreads = G1 * G2;
                p.in_sx = 1; p.in_sc = G1 * G2; p.out_sx = 1; p.out_sf = G1 * G2;
            } else if (k == 1 && dim == 3) {        // Y1[i0][g1][g2] -> Y2[i0][i1][g2]
                p.X = (int)G2; p.nthreads = N0 * G2;
                p.in_su = G1 * G2; p.in_sx = 1; p.in_sc = G2;
                p.out_su = N1 * G2; p.out_sx = 1; p.out_sf = G2;
            } else if (dim == 2) {                  // Y1[i0][g1] -> out[i0][i1]
                p.X = 1; p.nthreads = N0;
                p.in_su = G1; p.in_sc = 1; p.out_su = N1; p.out_sf = 1;
            } else {                                // Y2[i0][i1][g2] -> out[i0][i1][i2]
                p.X = 1; p.nthreads = N0 * N1;
                p.in_su = G2; p.in_sc = 1; p.out_su = N2; p.out_sf = 1;
            }
            ++g_launches;
            int e = fn(&p, st);
            if (e) return fail(PB200_ECUDA, "vector kernel launch failed: %s", pbErrorString((pbError)e));
        }
    }
    return 0;
}

extern "C" int pb200_asm_assemble_mlb(pb200_assembler* a, int row0_begin, int row0_end, double* d_out, void* d_work,
                                      size_t work_bytes, void* stream) {
    if (!a || !d_out) return fail(PB200_EINVAL, "null argument");
    if (a->arity != 2) return fail(PB200_EINVAL, "matrix assembly needs a bilinear form (arity 2)");
    const bool fuse1 = fused_stage1(a);
    const bool fuse23 = fused_stage23(a);
    if (!fuse1 && (!a->d_fields || !a->fields_valid)) return fail(PB200_EINVAL, "fields have not been computed");
    if (!a->fast) return pb200_asm_assemble_mlb_entrywise(a, row0_begin, row0_end, d_out, stream);
    CK(pbSetDevice(a->device));
    pbStream st = (pbStream)stream;
    const bool stiff = a->form == PB200_FORM_STIFFNESS;
    a->stage_names.clear();
    Slab S;
    int rc = make_slab(a, row0_begin, row0_end, uses_transposes(a), S);
    if (rc) return rc;
    size_t t1, s1, t2, s2;
    stage_sizes(a, S, t1, s1, t2, s2);
    const size_t need = (t1 * s1 + t2 * s2) * sizeof(double);
    if (!d_work || work_bytes < need) return fail(PB200_ENOMEM, "workspace too small: need %zu bytes, have %zu", need, work_bytes);
    double* X1 = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(d_work) + 255) & ~uintptr_t(255));
    if ((size_t)((char*)X1 - (char*)d_work) + need > work_bytes) X1 = (double*)d_work;
    double* X2 = X1 + t1 * s1;
    const long long npts = a->npts;
    const double* F = a->d_fields;
    // slab filters (see walk.cuh): terms that are read through the transposed band index must exist
    // for every pair with i or j in the slab (2); terms that are only read directly are needed for
    // the lines the final stage computes: rows of the slab (1), or - when the final stage mirrors
    // the symmetric half - only their "upper" part (3)
    const int last_axis = a->dim - 1;
    const int final_plan = stiff ? PB_PLAN_FINAL4 : PB_PLAN_COPY;
    const bool mirror = a->mirror_opt && a->symmetric && a->same_space && a->form != PB200_FORM_CUSTOM
                        && a->lane_ok[last_axis] && !a->force_walk
                        && pb_find_walk(PB_PLAN_LANE_BASE + final_plan, a->hax[last_axis].U.p, a->hax[last_axis].q) != nullptr;
    const int m_tr = uses_transposes(a) ? 2 : 1;        // terms read directly and transposed
    const int m_dir = (mirror || fuse23) ? 3 : 1;       // terms read directly only
    auto set_modes = [](int* dst, std::initializer_list<int> m) { int k = 0; for (int v : m) dst[k++] = v; };

    const AxisHost &H0 = a->hax[0], &H1 = a->hax[1];
    const PbAxis &D0 = a->dax[0], &D1 = a->dax[1];
    const long long G1 = H1.G, M1 = H1.M;
    const int Mrows = S.mu_hi - S.mu_lo, Mext = S.ext_hi - S.ext_lo;

    auto base_params = [&]() {
        PbWalkParams p;
        memset(&p, 0, sizeof p);
        p.V = 1; p.X = 1;
        return p;
    };

    if (a->form == PB200_FORM_CUSTOM) {
        // ---- generic scalar form: one GEN4 / COPY launch per output term of every stage -------
        std::vector<GenStage> stages;
        generic_plan(a, stages);
        const int dim = a->dim;
        const long long Glast = dim == 3 ? a->hax[2].G : 1;
        for (int k = 0; k < dim; ++k) {
            if (k == 1 && fuse23) {
                // stages 2 + 3 in one kernel (fused23.cuh, PbS32Generic): input stream 3 * test + trial reads
                // the X1 term whose remaining slots are (test, trial) in {v, d1, d2}
                PbS32Params q;
                memset(&q, 0, sizeof q);
                fill_s32_axes(a, q);
                for (int i = 0; i < 9; ++i) q.in_slot[i] = -1;
                auto code = [](int slot) { return slot == 0 ? 0 : slot - 1; };         // v -> 0, d1 (slot 2) -> 1, d2 (slot 3) -> 2
                for (const GenTerm& t : stages[0].out) q.in_slot[3 * code(t.bt) + code(t.bu)] = t.slot;
                q.X1 = X1; q.x1_stride = (long long)s1; q.x1_mu_base = S.mu_lo;
                q.mu0_begin = S.mu_lo; q.mu0_count = Mrows;
                q.u_lo = S.ra; q.u_hi = S.rb;
                q.symmetric = 0;
                q.out = d_out; q.out_mu_base = S.mu_lo;
                q.nbatch = pb_lane_batches(a->hax[2].n, a->hax[2].U.p);
                rc = fill_s32_tasks(a, S, q, st);
                if (rc) return rc;
                choose_s32_pieces(a, q);
                mark_stage(a, "g23", st);
                ++g_launches;
                int e = pb_find_s32(PB200_FORM_CUSTOM, H1.U.p, H1.q)(&q, st);
                if (e) return fail(PB200_ECUDA, "fused stage 2+3 launch failed: %s", pbErrorString((pbError)e));
                break;
            }
            const GenStage& G = stages[k];
            const bool last = (k == dim - 1);
            for (size_t o = 0; o < G.out.size(); ++o) {
                PbWalkParams p = base_params();
                const double* inbase = (k == 0) ? F : (k == 1 ? X1 : X2);
                const long long instride = (k == 0) ? npts : (k == 1 ? (long long)s1 : (long long)s2);
                double* outbase = last ? d_out : (k == 0 ? X1 : X2);
                const long long outstride = last ? 0 : (k == 0 ? (long long)s1 : (long long)s2);
                for (int c = 0; c < 4; ++c) p.in[c] = G.ops[o][c] >= 0 ? inbase + (long long)G.ops[o][c] * instride : nullptr;
                p.out[0] = outbase + (long long)o * outstride;
                if (k == 0) {               // axis 0: lines are the points of the trailing grid axes
                    p.X = (int)(G1 * Glast); p.nthreads = G1 * Glast;
                    p.in_sx = 1; p.in_sc = G1 * Glast;
                    p.out_sx = 1; p.out_smu = G1 * Glast; p.mu_base = S.mu_lo;
                    p.s_begin = S.sa; p.s_end = S.sb;
                    p.w_mode[0] = 1; p.w_lo = S.ra; p.w_hi = S.rb; p.w_ext_lo = S.ra; p.w_ext_hi = S.rb;
                } else if (k == 1 && dim == 3) {
                    p.X = (int)Glast; p.nthreads = (long long)Mrows * Glast;
                    p.u_begin = S.mu_lo; p.u_base_in = S.mu_lo; p.u_base_out = S.mu_lo;
                    p.in_su = G1 * Glast; p.in_sx = 1; p.in_sc = Glast;
                    p.out_su = M1 * Glast; p.out_sx = 1; p.out_smu = Glast; p.mu_base = 0;
                    p.s_begin = 0; p.s_end = H1.n;
                } else if (dim == 2) {      // final stage in 2D: axis 1
                    p.nthreads = Mrows;
                    p.u_begin = S.mu_lo; p.u_base_in = S.mu_lo; p.u_base_out = S.mu_lo;
                    p.in_su = G1; p.in_sc = 1;
                    p.out_su = M1; p.out_smu = 1; p.mu_base = 0;
                    p.s_begin = 0; p.s_end = H1.n;
                } else {                    // final stage in 3D: axis 2
                    const long long M2g = a->hax[2].M;
                    p.V = (int)M1; p.nthreads = (long long)Mrows * M1;
                    p.u_begin = S.mu_lo; p.u_base_in = S.mu_lo; p.u_base_out = S.mu_lo;
                    p.in_su = M1 * Glast; p.in_sv = Glast; p.in_sc = 1;
                    p.out_su = M1 * M2g; p.out_sv = M2g; p.out_smu = 1; p.mu_base = 0;
                    p.s_begin = 0; p.s_end = a->hax[2].n;
                }
                const bool copy_only = G.ops[o][1] < 0 && G.ops[o][2] < 0 && G.ops[o][3] < 0;
                char nm[32];
                snprintf(nm, sizeof nm, "g%d_%s%d", k + 1, copy_only ? "copy" : "gen", (int)o);
                a->walk_table = a->pair_tab[k][G.tab[o]];
                rc = run_stage(copy_only ? PB_PLAN_COPY : PB_PLAN_GEN4, a, k, p, st, nm);
                a->walk_table = nullptr;
                if (rc) return rc;
            }
        }
        mark_stage(a, "end", st);
        return 0;
    }

    if (a->dim == 2) {
        // stage 1: axis 0,  fields[f][g0][g1] -> X1[t][mu0 - ext_lo][g1]
        {
            PbWalkParams p = base_params();
            p.X = (int)G1; p.nthreads = G1;
            p.in_sx = 1; p.in_sc = G1;
            p.out_sx = 1; p.out_smu = G1; p.mu_base = S.ext_lo;
            p.s_begin = S.sa; p.s_end = S.sb;
            p.w_lo = S.ra; p.w_hi = S.rb; p.w_ext_lo = S.ea; p.w_ext_hi = S.eb;
            const int modes2d[3] = {m_dir, m_tr, m_dir};                        // (v,v), (v,d1), (d1,d1)
            if (stiff && a->fused_plans) {
                p.in[0] = F + 2 * npts; p.in[1] = F + 1 * npts; p.in[2] = F;     // B11, B01, B00
                for (int t = 0; t < 3; ++t) { p.out[t] = X1 + t * s1; p.w_mode[t] = modes2d[t]; }
                rc = run_stage(PB_PLAN_S1_2D, a, 0, p, st, "s1_2d");
            } else if (stiff) {
                // one launch per output: every field is read by exactly one of them
                const int plan[3] = {PB_PLAN_ONE11, PB_PLAN_ONE10, PB_PLAN_COPY};
                const int field[3] = {2, 1, 0};                                 // B11, B01, B00
                const char* nm[3] = {"s1_one11", "s1_one10", "s1_copy"};
                for (int t = 0; t < 3 && !rc; ++t) {
                    PbWalkParams q = p;
                    q.in[0] = F + field[t] * npts; q.out[0] = X1 + t * s1; q.w_mode[0] = modes2d[t];
                    rc = run_stage(plan[t], a, 0, q, st, nm[t]);
                }
            } else {
                p.in[0] = F; p.out[0] = X1; p.w_mode[0] = m_dir;
                rc = run_stage(PB_PLAN_COPY, a, 0, p, st, "s1_copy");
            }
            if (rc) return rc;
        }
        // stage 2: axis 1,  X1[t][mu0][g1] -> out[mu0 - mu_lo][mu1]
        {
            PbWalkParams p = base_params();
            p.nthreads = Mrows;
            p.u_begin = S.mu_lo; p.u_base_in = S.ext_lo; p.u_base_out = S.mu_lo;
            p.tr_u = D0.tr;
            p.in_su = G1; p.in_sc = 1;
            p.out_su = M1; p.out_smu = 1; p.mu_base = 0;
            p.s_begin = 0; p.s_end = H1.n;
            p.out[0] = d_out;
            if (mirror) {
                p.u_pair_i = D0.pair_i; p.u_pair_j = D0.pair_j;
                p.u_mode[0] = 1; p.u_lo = S.ra; p.u_hi = S.rb;
                p.mirror = 1;
            }
            if (stiff) {
                p.in[0] = X1; p.in[1] = X1 + s1; p.in[2] = X1 + s1; p.in[3] = X1 + 2 * s1;
                rc = run_stage(PB_PLAN_FINAL4, a, 1, p, st, "s2_final4");
            } else {
                p.in[0] = X1;
                rc = run_stage(PB_PLAN_COPY, a, 1, p, st, "s2_copy");
            }
            mark_stage(a, "end", st);
            return rc;
        }
    }

    const AxisHost& H2 = a->hax[2];
    const PbAxis& D2 = a->dax[2];
    (void)D2;
    const long long G2 = H2.G, M2 = H2.M;
    // stage 1: axis 0,  fields[f][g0][g1][g2] -> X1[t][mu0 - ext_lo][g1][g2]
    {
        PbWalkParams p = base_params();
        p.X = (int)(G1 * G2); p.nthreads = G1 * G2;
        p.in_sx = 1; p.in_sc = G1 * G2;
        p.out_sx = 1; p.out_smu = G1 * G2; p.mu_base = S.ext_lo;
        p.s_begin = S.sa; p.s_end = S.sb;
        p.w_lo = S.ra; p.w_hi = S.rb; p.w_ext_lo = S.ea; p.w_ext_hi = S.eb;
        // X1 terms (v,v) (v,d1) (v,d2) (d1,d1) (d1,d2) (d2,d2): which are read through a transposed index?
        // (with the fused stages 2 + 3 the term (d1,d2) is only read directly)
        const int modes1[6] = {m_dir, m_tr, m_tr, m_dir, fuse23 ? m_dir : m_tr, m_dir};
        PbGeoLineParams gl;
        if (fuse1) {
            gl.geo = a->geo_dev;
            for (int k = 0; k < 3; ++k) gl.gw[k] = a->dax[k].weights;
            gl.G1 = (int)G1; gl.G2 = (int)G2;
            p.geo_line = &gl;
            if (stiff) {
                for (int t = 0; t < 6; ++t) { p.out[t] = X1 + t * s1; p.w_mode[t] = modes1[t]; }
                rc = run_stage(PB_PLAN_S1F, a, 0, p, st, "s1f");
            } else {
                p.out[0] = X1; p.w_mode[0] = m_dir;
                rc = run_stage(PB_PLAN_S1F_MASS, a, 0, p, st, "s1f_mass");
            }
        } else if (stiff && !a->fused_plans) {
            const int plan[6] = {PB_PLAN_ONE11, PB_PLAN_ONE10, PB_PLAN_ONE10, PB_PLAN_COPY, PB_PLAN_COPY, PB_PLAN_COPY};
            const int field[6] = {5, 4, 2, 3, 1, 0};                            // B22, B12, B02, B11, B01, B00
            const char* nm[6] = {"s1_one11", "s1_one10a", "s1_one10b", "s1_copya", "s1_copyb", "s1_copyc"};
            for (int t = 0; t < 6 && !rc; ++t) {
                PbWalkParams q = p;
                q.in[0] = F + field[t] * npts; q.out[0] = X1 + t * s1; q.w_mode[0] = modes1[t];
                rc = run_stage(plan[t], a, 0, q, st, nm[t]);
            }
        } else if (stiff) {
            PbWalkParams pa = p, pb = p;
            for (int t = 0; t < 3; ++t) { pa.w_mode[t] = modes1[t]; pb.w_mode[t] = modes1[3 + t]; }
            pa.in[0] = F + 5 * npts; pa.in[1] = F + 4 * npts; pa.in[2] = F + 2 * npts;   // B22, B12, B02
            for (int t = 0; t < 3; ++t) pa.out[t] = X1 + t * s1;
            pb.in[0] = F + 3 * npts; pb.in[1] = F + 1 * npts; pb.in[2] = F;              // B11, B01, B00
            for (int t = 0; t < 3; ++t) pb.out[t] = X1 + (3 + t) * s1;
            rc = run_stage(PB_PLAN_S1A, a, 0, pa, st, "s1a");
            if (rc) return rc;
            rc = run_stage(PB_PLAN_S1B, a, 0, pb, st, "s1b");
        } else {
            p.in[0] = F; p.out[0] = X1; p.w_mode[0] = m_dir;
            rc = run_stage(PB_PLAN_COPY, a, 0, p, st, "s1_copy");
        }
        if (rc) return rc;
    }
    if (fuse23) {
        // stages 2 + 3 in one kernel: X1[t][mu0][g1][g2] -> out[mu0 - mu_lo][mu1][mu2]
        PbS32Params q;
        memset(&q, 0, sizeof q);
        fill_s32_axes(a, q);
        q.X1 = X1; q.x1_stride = (long long)s1; q.x1_mu_base = S.ext_lo;
        q.mu0_begin = S.mu_lo; q.mu0_count = Mrows;
        q.u_lo = S.ra; q.u_hi = S.rb;
        q.symmetric = 1;
        q.out = d_out; q.out_mu_base = S.mu_lo;
        q.nbatch = pb_lane_batches(H2.n, H2.U.p);
        rc = fill_s32_tasks(a, S, q, st);
        if (rc) return rc;
        choose_s32_pieces(a, q);
        mark_stage(a, stiff ? "s23" : "s23_mass", st);
        ++g_launches;
        int e = pb_find_s32(a->form, H1.U.p, H1.q)(&q, st);
        if (e) return fail(PB200_ECUDA, "fused stage 2+3 launch failed: %s", pbErrorString((pbError)e));
        mark_stage(a, "end", st);
        return 0;
    }
    // stage 2: axis 1,  X1[t][mu0][g1][g2] -> X2[t][mu0][mu1][g2]
    {
        PbWalkParams p = base_params();
        p.X = (int)G2; p.nthreads = (long long)Mext * G2;
        p.u_begin = S.ext_lo; p.u_base_in = S.ext_lo; p.u_base_out = S.ext_lo;
        p.tr_u = D0.tr;
        p.u_pair_i = D0.pair_i; p.u_pair_j = D0.pair_j;
        p.u_lo = S.ra; p.u_hi = S.rb;
        // X2 terms (v,v) (v,d2) (d2,d2): only (v,d2) is read through the transposed index
        const int modes2[3] = {m_dir, m_tr, m_dir};
        p.in_su = G1 * G2; p.in_sx = 1; p.in_sc = G2;
        p.out_su = M1 * G2; p.out_sx = 1; p.out_smu = G2; p.mu_base = 0;
        p.s_begin = 0; p.s_end = H1.n;
        if (stiff) {
            PbWalkParams pa = p, pb = p;
            pa.in[0] = X1; pa.in[1] = X1 + s1; pa.in[2] = X1 + s1; pa.in[3] = X1 + 3 * s1;
            pa.out[0] = X2; pa.u_mode[0] = modes2[0];
            pb.in[0] = X1 + 2 * s1; pb.in[1] = X1 + 4 * s1; pb.in[2] = X1 + 5 * s1;
            pb.out[0] = X2 + s2; pb.out[1] = X2 + 2 * s2;
            pb.u_mode[0] = modes2[1]; pb.u_mode[1] = modes2[2];
            rc = run_stage(PB_PLAN_FINAL4, a, 1, pa, st, "s2a_final4");
            if (rc) return rc;
            if (a->fused_plans) {
                rc = run_stage(PB_PLAN_S2B, a, 1, pb, st, "s2b");
            } else {
                PbWalkParams pc = p, pd = p;
                pc.in[0] = X1 + 2 * s1; pc.in[1] = X1 + 4 * s1; pc.out[0] = X2 + s2;    // (v,d2)[0,0] + (d1,d2)[1,0]
                pd.in[0] = X1 + 5 * s1; pd.out[0] = X2 + 2 * s2;                          // (d2,d2)[0,0]
                pc.u_mode[0] = modes2[1]; pd.u_mode[0] = modes2[2];
                rc = run_stage(PB_PLAN_PAIRT, a, 1, pc, st, "s2_pairt");
                if (rc) return rc;
                rc = run_stage(PB_PLAN_COPY, a, 1, pd, st, "s2_copy");
            }
        } else {
            p.in[0] = X1; p.out[0] = X2; p.u_mode[0] = m_dir;
            rc = run_stage(PB_PLAN_COPY, a, 1, p, st, "s2_copy");
        }
        if (rc) return rc;
    }
    // stage 3: axis 2,  X2[t][mu0][mu1][g2] -> out[mu0 - mu_lo][mu1][mu2]
    {
        PbWalkParams p = base_params();
        p.V = (int)M1; p.nthreads = (long long)Mrows * M1;
        p.u_begin = S.mu_lo; p.u_base_in = S.ext_lo; p.u_base_out = S.mu_lo;
        p.tr_u = D0.tr; p.tr_v = D1.tr;
        p.in_su = M1 * G2; p.in_sv = G2; p.in_sc = 1;
        p.out_su = M1 * M2; p.out_sv = M2; p.out_smu = 1; p.mu_base = 0;
        p.s_begin = 0; p.s_end = H2.n;
        p.out[0] = d_out;
        if (mirror) {
            p.u_pair_i = D0.pair_i; p.u_pair_j = D0.pair_j;
            p.v_pair_i = D1.pair_i; p.v_pair_j = D1.pair_j;
            p.u_mode[0] = 1; p.u_lo = S.ra; p.u_hi = S.rb;
            p.tr_u = D0.tr; p.tr_v = D1.tr;
            p.mirror = 1;
        }
        if (stiff) {
            p.in[0] = X2; p.in[1] = X2 + s2; p.in[2] = X2 + s2; p.in[3] = X2 + 2 * s2;
            rc = run_stage(PB_PLAN_FINAL4, a, 2, p, st, "s3_final4");
        } else {
            p.in[0] = X2;
            rc = run_stage(PB_PLAN_COPY, a, 2, p, st, "s3_copy");
        }
    }
    (void)H0;
    mark_stage(a, "end", st);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// per-entry path
// ------------------------------------------------------------------------------------------------
static void fill_entry_params(const pb200_assembler* a, PbEntryParams& prm) {
    memset(&prm, 0, sizeof prm);
    prm.dim = a->dim;
    for (int k = 0; k < a->dim; ++k) prm.ax[k] = a->dax[k];
    prm.fields = a->d_fields;
    prm.npts = a->npts;
    prm.nterms = (int)a->terms.size();
    prm.ext = a->ext_slots ? 1 : 0;
    for (int t = 0; t < prm.nterms; ++t) {
        prm.terms[t] = a->terms[t];
        for (int k = 0; k < a->dim; ++k) {
            prm.ot[t][k] = (unsigned char)pb_slot_order(a->terms[t].bt, k);
            prm.ou[t][k] = (unsigned char)(a->terms[t].bu < 0 ? 0 : pb_slot_order(a->terms[t].bu, k));
        }
    }
}

extern "C" int pb200_asm_multi_entries(pb200_assembler* a, const uint64_t* d_ij, size_t n, double* d_out, void* stream) {
    if (!a || (n && (!d_ij || !d_out))) return fail(PB200_EINVAL, "null argument");
    if (a->arity != 2) return fail(PB200_EINVAL, "multi_entries needs a bilinear form (arity 2)");
    if (!a->d_fields) return fail(PB200_EINVAL, "fields have not been computed");
    if (n == 0) return 0;
    CK(pbSetDevice(a->device));
    PbEntryParams prm;
    fill_entry_params(a, prm);
    prm.ij = reinterpret_cast<const unsigned long long*>(d_ij);
    prm.n = (long long)n;
    prm.out = d_out;
    if (a->dim == 2) k_entries<2>(prm, (pbStream)stream);
    else k_entries<3>(prm, (pbStream)stream);
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_asm_rows_count(const pb200_assembler* a, const int64_t* h_rows, long long n, int64_t* h_indptr) {
    if (!a || n < 0 || !h_indptr || (n > 0 && !h_rows)) return fail(PB200_EINVAL, "null argument");
    if (a->arity != 2) return fail(PB200_EINVAL, "partial-row assembly needs a bilinear form (arity 2)");
    long long nrows_total = 1;
    for (int k = 0; k < a->dim; ++k) nrows_total *= a->ml.Nv[k];
    h_indptr[0] = 0;
    for (long long r = 0; r < n; ++r) {
        long long I = h_rows[r];
        if (I < 0 || I >= nrows_total) return fail(PB200_EINVAL, "row index %lld out of range [0,%lld)", I, nrows_total);
        long long cnt = 1;
        for (int k = a->dim - 1; k >= 0; --k) {
            const int i = (int)(I % a->ml.Nv[k]);
            I /= a->ml.Nv[k];
            cnt *= a->hax[k].row_start[i + 1] - a->hax[k].row_start[i];
        }
        h_indptr[r + 1] = h_indptr[r] + cnt;
    }
    return 0;
}

template <int DIM, class IdxT>
static void k_rows_fill(const PbEntryParams& prm, const long long* rows, long long n, const IdxT* indptr, IdxT* indices,
                        double* values, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    for (long long r = 0; r < n; ++r)
        for (long long e = 0; e < (long long)(indptr[r + 1] - indptr[r]); ++e)
            pb_row_entry<DIM, IdxT>(prm, (unsigned long long)rows[r], (long long)indptr[r], (int)e, indices, values);
#else
    const unsigned blocks = (unsigned)std::min<long long>(n, 148LL * 64);
    pb_rows_fill_kernel<DIM, IdxT><<<blocks, 128, 0, st>>>(prm, rows, indptr, n, indices, values);
#endif
}

extern "C" int pb200_asm_rows_fill(pb200_assembler* a, const int64_t* d_rows, long long n, const void* d_indptr,
                                   void* d_indices, double* d_values, int idx_bytes, void* stream) {
    if (!a || n < 0 || (n > 0 && (!d_rows || !d_indptr))) return fail(PB200_EINVAL, "null argument");
    if (idx_bytes != 4 && idx_bytes != 8) return fail(PB200_EINVAL, "idx_bytes must be 4 or 8");
    if (a->arity != 2) return fail(PB200_EINVAL, "partial-row assembly needs a bilinear form (arity 2)");
    if (!a->d_fields) return fail(PB200_EINVAL, "fields have not been computed");
    if (n == 0) return 0;
    CK(pbSetDevice(a->device));
    PbEntryParams prm;
    fill_entry_params(a, prm);
    pbStream st = (pbStream)stream;
    const long long* rows = reinterpret_cast<const long long*>(d_rows);
    if (a->dim == 2) {
        if (idx_bytes == 4) k_rows_fill<2, int>(prm, rows, n, (const int*)d_indptr, (int*)d_indices, d_values, st);
        else k_rows_fill<2, long long>(prm, rows, n, (const long long*)d_indptr, (long long*)d_indices, d_values, st);
    } else {
        if (idx_bytes == 4) k_rows_fill<3, int>(prm, rows, n, (const int*)d_indptr, (int*)d_indices, d_values, st);
        else k_rows_fill<3, long long>(prm, rows, n, (const long long*)d_indptr, (long long*)d_indices, d_values, st);
    }
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_asm_assemble_mlb_entrywise(pb200_assembler* a, int row0_begin, int row0_end, double* d_out, void* stream) {
    if (!a || !d_out) return fail(PB200_EINVAL, "null argument");
    if (!a->d_fields) return fail(PB200_EINVAL, "fields have not been computed");
    CK(pbSetDevice(a->device));
    Slab S;
    int rc = make_slab(a, row0_begin, row0_end, false, S);
    if (rc) return rc;
    PbEntryParams prm;
    fill_entry_params(a, prm);
    prm.out = d_out;
    long long count = S.mu_hi - S.mu_lo;
    for (int k = 1; k < a->dim; ++k) count *= a->hax[k].M;
    if (a->dim == 2) k_entries_mlb<2>(prm, S.mu_lo, count, (pbStream)stream);
    else k_entries_mlb<3>(prm, S.mu_lo, count, (pbStream)stream);
    CK(pbLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// MLB utilities
// ------------------------------------------------------------------------------------------------
static int fill_mlb_params(const pb200_mlstruct* a, int ra, int rb, const double* d_mlb, PbMlbParams& p) {
    const int N = a->Nv[0];
    if (ra < 0 || rb > N || ra >= rb) return fail(PB200_EINVAL, "invalid row slab [%d,%d) of %d rows", ra, rb, N);
    memset(&p, 0, sizeof p);
    p.dim = a->dim;
    for (int k = 0; k < a->dim; ++k) {
        p.Nv[k] = a->Nv[k];
        p.Nu[k] = a->Nu[k];
        p.M[k] = a->M[k];
        p.row_start[k] = a->d_row_start[k];
        p.jmin[k] = a->d_jmin[k];
        p.pair_i[k] = a->d_pair_i[k];
        p.pair_j[k] = a->d_pair_j[k];
    }
    p.row0_begin = ra; p.row0_end = rb;
    p.data = d_mlb;
    return 0;
}

extern "C" int pb200_mlb_to_csr(const pb200_mlstruct* a, int ra, int rb, const double* d_mlb, void* d_indptr,
                                void* d_indices, double* d_values, int idx_bytes, void* stream) {
    // d_indptr / d_indices may be null: values only (the pattern can be produced on the host by
    // pb200_csr_pattern_host while the values are still being computed and copied)
    if (!a || !d_mlb || !d_values) return fail(PB200_EINVAL, "null argument");
    if (idx_bytes != 4 && idx_bytes != 8) return fail(PB200_EINVAL, "idx_bytes must be 4 or 8");
    CK(pbSetDevice(a->device));
    PbMlbParams p;
    int rc = fill_mlb_params(a, ra, rb, d_mlb, p);
    if (rc) return rc;
    long long nrows = rb - ra, count = a->row_start0[rb] - a->row_start0[ra];
    for (int k = 1; k < a->dim; ++k) { nrows *= p.Nv[k]; count *= p.M[k]; }
    if (idx_bytes == 4 && count > 2147483647LL) return fail(PB200_EINVAL, "slab has %lld entries: needs 64-bit CSR indices", count);
    pbStream st = (pbStream)stream;
    if (idx_bytes == 4) k_csr<int>(p, nrows, count, (int*)d_indptr, (int*)d_indices, d_values, st);
    else k_csr<long long>(p, nrows, count, (long long*)d_indptr, (long long*)d_indices, d_values, st);
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_mlb_matvec(const pb200_mlstruct* a, int ra, int rb, const double* d_mlb, const double* d_x,
                                int x_j0_begin, double* d_y, void* stream) {
    if (!a || !d_mlb || !d_x || !d_y) return fail(PB200_EINVAL, "null argument");
    CK(pbSetDevice(a->device));
    PbMlbParams p;
    int rc = fill_mlb_params(a, ra, rb, d_mlb, p);
    if (rc) return rc;
    long long nrows = rb - ra;
    for (int k = 1; k < a->dim; ++k) nrows *= p.Nv[k];
    k_matvec(p, nrows, d_x, x_j0_begin, d_y, (pbStream)stream);
    CK(pbLastError());
    return 0;
}

// host-only: CSR pattern of the rows [row0_begin, row0_end) of axis 0 of a multi-level banded structure.
// Every thread writes one contiguous piece of `indices`.  Rows are built in an L1-resident staging buffer (an
// interior row is the previous row shifted by one column) and leave it in whole cache lines with non-temporal
// stores: no read-for-ownership of the result arrays, which the GPU's D2H copies of the values are filling at
// the same time (1.4-2.3x the rate of plain stores on 1-8 threads).
#include <cstdint>
#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
#define PB_NT_STORES 1
#endif

template <class IT>
struct PbRowStream {
    static constexpr int CAP = 4096;                    // staging entries (16 / 32 KB)
    static constexpr int LINE = 64 / (int)sizeof(IT);   // entries per cache line
    IT* dst;                                            // next output position
    int fill = 0;
    bool aligned = false;                               // dst is on a cache-line boundary
    alignas(64) IT buf[CAP];
    explicit PbRowStream(IT* d) : dst(d) {}
    static bool misaligned(const IT* q) { return ((uintptr_t)q & 63) != 0; }
    // room for a row of `len` entries.  Earlier rows may leave; the row at buf[keep] (if keep >= 0) stays
    // addressable and `keep` follows it.
    IT* reserve(int len, int& keep) {
        if (fill + len > CAP) flush(keep);
        return buf + fill;
    }
    void flush(int& keep) {
        const int upto = keep >= 0 ? keep : fill;       // entries [0, upto) may leave
        int done = 0;
        if (!aligned) {                                 // head of the piece: plain stores up to the first boundary
            while (done < upto && misaligned(dst + done)) { dst[done] = buf[done]; ++done; }
            aligned = !misaligned(dst + done);
        }
        if (aligned) {
            const int lines = (upto - done) / LINE;
#ifdef PB_NT_STORES
            const __m128i* s = (const __m128i*)(buf + done);
            __m128i* d = (__m128i*)(dst + done);
            for (int l = 0; l < lines; ++l, s += 4, d += 4) {
                const __m128i a = _mm_loadu_si128(s), b = _mm_loadu_si128(s + 1), c = _mm_loadu_si128(s + 2), e = _mm_loadu_si128(s + 3);
                _mm_stream_si128(d, a); _mm_stream_si128(d + 1, b); _mm_stream_si128(d + 2, c); _mm_stream_si128(d + 3, e);
            }
#else
            memcpy(dst + done, buf + done, (size_t)lines * 64);
#endif
            done += lines * LINE;
        }
        dst += done;
        const int rest = fill - done;
        if (rest > 0 && done > 0) memmove(buf, buf + done, (size_t)rest * sizeof(IT));
        fill = rest;
        if (keep >= 0) keep -= done;
    }
    void finish() {                                     // everything out (the tail of less than a line: plain stores)
        int none = -1;
        flush(none);
        for (int e = 0; e < fill; ++e) dst[e] = buf[e];
        dst += fill;
        fill = 0;
#ifdef PB_NT_STORES
        _mm_sfence();
#endif
    }
};

template <class IT>
static void pb_pattern_row(int L, const int* Nu, const int* nb, const int* j0, IT* o) {
    if (L == 2) {
        for (int k0 = 0; k0 < nb[0]; ++k0) {
            const long long base = (long long)(j0[0] + k0) * Nu[1] + j0[1];
            for (int k1 = 0; k1 < nb[1]; ++k1) *o++ = (IT)(base + k1);
        }
    } else {
        for (int k0 = 0; k0 < nb[0]; ++k0)
            for (int k1 = 0; k1 < nb[1]; ++k1) {
                const long long base = ((long long)(j0[0] + k0) * Nu[1] + j0[1] + k1) * Nu[2] + j0[2];
                for (int k2 = 0; k2 < nb[2]; ++k2) *o++ = (IT)(base + k2);
            }
    }
}

template <class IT>
static void csr_pattern_rows(int L, const int* Nv, const int* Nu, const int* const* rs, const int* const* jm, const int* M,
                             int ra, long long r_begin, long long r_end, IT* indptr, IT* indices, long long off0) {
    if (r_begin >= r_end) return;
    const long long m1 = L > 1 ? M[1] : 1, m2 = L > 2 ? M[2] : 1;
    std::unique_ptr<PbRowStream<IT>> st(new PbRowStream<IT>(indices));
    int prev = -1;                                      // the previous row in the staging buffer
    int prev_nb = 0, prev_j0 = 0, prev_len = 0;
    int i[3] = {0, 0, 0};                               // per-axis row indices, advanced like an odometer
    {
        long long t = r_begin;
        for (int k = L - 1; k >= 1; --k) { i[k] = (int)(t % Nv[k]); t /= Nv[k]; }
        i[0] = (int)t + ra;
    }
    for (long long r = r_begin; r < r_end; ++r) {
        int nb[3] = {1, 1, 1}, j0[3] = {0, 0, 0};
        for (int k = 0; k < L; ++k) { nb[k] = rs[k][i[k] + 1] - rs[k][i[k]]; j0[k] = jm[k][i[k]]; }
        long long off = (long long)(rs[0][i[0]] - rs[0][ra]) * m1 * m2;
        if (L == 2) off += (long long)nb[0] * rs[1][i[1]];
        if (L == 3) off += (long long)nb[0] * ((long long)rs[1][i[1]] * m2 + (long long)nb[1] * rs[2][i[2]]);
        indptr[r] = (IT)(off0 + off);
        const int len = nb[0] * nb[1] * nb[2];
        if (st->dst + st->fill != indices + off) {      // first row of the piece (rows of a piece are contiguous in CSR order)
            st->finish();
            st->dst = indices + off;
            st->aligned = false;
            prev = -1;
        }
        if (len > PbRowStream<IT>::CAP / 2) {           // very wide rows are not staged
            st->finish();
            pb_pattern_row<IT>(L, Nu, nb, j0, indices + off);
            st->dst = indices + off + len;
            st->aligned = false;
            prev = -1;
        } else {
            IT* __restrict cur = st->reserve(len, prev);
            // interior rows: the pattern of (.., i_last + 1) is the pattern of (.., i_last) shifted by one column
            if (prev >= 0 && i[L - 1] > 0 && nb[L - 1] == prev_nb && j0[L - 1] == prev_j0 + 1 && len == prev_len) {
                const IT* src = st->buf + prev;
                for (int e = 0; e < len; ++e) cur[e] = src[e] + 1;
            } else {
                pb_pattern_row<IT>(L, Nu, nb, j0, cur);
            }
            prev = (int)(cur - st->buf);
            st->fill += len;
        }
        prev_nb = nb[L - 1]; prev_j0 = j0[L - 1]; prev_len = len;
        int k = L - 1;
        while (k >= 1 && ++i[k] == Nv[k]) { i[k] = 0; --k; }
        if (k == 0) ++i[0];
    }
    st->finish();
}

extern "C" int pb200_csr_pattern_host(int nlevels, const int* rows, const int* cols, const int* nband,
                                      const int* const* h_row_start, const int* const* h_jmin, int row0_begin,
                                      int row0_end, void* h_indptr, void* h_indices, int idx_bytes,
                                      long long indptr_offset, int nthreads) {
    if (!rows || !cols || !nband || !h_row_start || !h_jmin || !h_indptr || !h_indices) return fail(PB200_EINVAL, "null argument");
    if (nlevels < 2 || nlevels > 3) return fail(PB200_EUNSUPPORTED, "2 or 3 levels (got %d)", nlevels);
    if (idx_bytes != 4 && idx_bytes != 8) return fail(PB200_EINVAL, "idx_bytes must be 4 or 8");
    if (row0_begin < 0 || row0_end > rows[0] || row0_begin > row0_end) return fail(PB200_EINVAL, "invalid row slab");
    long long nrows = row0_end - row0_begin;
    for (int k = 1; k < nlevels; ++k) nrows *= rows[k];
    long long count = (long long)(h_row_start[0][row0_end] - h_row_start[0][row0_begin]);
    for (int k = 1; k < nlevels; ++k) count *= nband[k];
    nthreads = std::max(1, std::min(nthreads, 256));
    auto work = [&](long long a, long long b) {
        if (idx_bytes == 4)
            csr_pattern_rows<int>(nlevels, rows, cols, h_row_start, h_jmin, nband, row0_begin, a, b, (int*)h_indptr,
                                  (int*)h_indices, indptr_offset);
        else
            csr_pattern_rows<long long>(nlevels, rows, cols, h_row_start, h_jmin, nband, row0_begin, a, b,
                                        (long long*)h_indptr, (long long*)h_indices, indptr_offset);
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        const long long a = nrows * t / nthreads, b = nrows * (t + 1) / nthreads;
        if (b > a) pool.emplace_back(work, a, b);
    }
    for (auto& th : pool) th.join();
    if (idx_bytes == 4) ((int*)h_indptr)[nrows] = (int)(indptr_offset + count);
    else ((long long*)h_indptr)[nrows] = indptr_offset + count;
    return 0;
}

// ---- CSR utilities (csr.cuh) ---------------------------------------------------------------------
// Entry points that take only device pointers (no handle that knows its device): launch on the device
// the data lives on, whatever the caller's current device is.
static int use_device_of(const void* d_ptr) {
#ifndef PB_EMULATE
    if (!d_ptr) return 0;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, d_ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) CK(cudaSetDevice(at.device));
    else cudaGetLastError();
#else
    (void)d_ptr;
#endif
    return 0;
}

extern "C" int pb200_csr_restrict_workspace(long long nrows_new, int idx_bytes, size_t* bytes) {
    if (!bytes || nrows_new < 0 || (idx_bytes != 4 && idx_bytes != 8)) return fail(PB200_EINVAL, "invalid argument");
#ifdef PB_EMULATE
    *bytes = 8;
#else
    *bytes = (size_t)((nrows_new + PB_SCAN_BLOCK - 1) / PB_SCAN_BLOCK + 1) * (size_t)idx_bytes;
#endif
    return 0;
}

template <class IT>
static int csr_restrict_count(long long n, const int* rows, const IT* indptr, const IT* indices, const int* colmap,
                              IT* indptr_new, IT* work, pbStream st) {
#ifdef PB_EMULATE
    (void)work; (void)st;
    indptr_new[0] = 0;
    for (long long r = 0; r < n; ++r) indptr_new[r + 1] = indptr_new[r] + (IT)pb_csr_count_row<IT>(indptr, indices, colmap, rows[r]);
    g_launches += 1;
#else
    CK(cudaMemsetAsync(indptr_new, 0, sizeof(IT), st));
    if (n > 0) {
        const unsigned b = (unsigned)std::min<long long>((n + 7) / 8, 148LL * 32);
        const unsigned nb = (unsigned)((n + PB_SCAN_BLOCK - 1) / PB_SCAN_BLOCK);
        pb_csr_restrict_count_kernel<IT><<<b, 256, 0, st>>>(n, rows, indptr, indices, colmap, indptr_new + 1);
        pb_scan_sums_kernel<IT><<<nb, 256, 0, st>>>(indptr_new + 1, n, work);
        pb_scan_top_kernel<IT><<<1, 256, 0, st>>>(work, nb);
        pb_scan_apply_kernel<IT><<<nb, 256, 0, st>>>(indptr_new + 1, n, work);
        g_launches += 4;
    }
    CK(pbLastError());
#endif
    return 0;
}

extern "C" int pb200_csr_restrict_count(long long nrows_new, const int32_t* d_rows, const void* d_indptr,
                                        const void* d_indices, int idx_bytes, const int32_t* d_colmap,
                                        void* d_indptr_new, void* d_work, size_t work_bytes, void* stream) {
    if (nrows_new < 0 || !d_indptr || !d_colmap || !d_indptr_new || (nrows_new > 0 && !d_rows))
        return fail(PB200_EINVAL, "null argument");
    if (int e = use_device_of(d_indptr)) return e;
    size_t need = 0;
    int rc = pb200_csr_restrict_workspace(nrows_new, idx_bytes, &need);
    if (rc) return rc;
    if (!d_work || work_bytes < need) return fail(PB200_EINVAL, "workspace too small: %zu < %zu bytes", work_bytes, need);
    pbStream st = (pbStream)stream;
    if (idx_bytes == 4)
        return csr_restrict_count<int>(nrows_new, d_rows, (const int*)d_indptr, (const int*)d_indices, d_colmap,
                                       (int*)d_indptr_new, (int*)d_work, st);
    return csr_restrict_count<long long>(nrows_new, d_rows, (const long long*)d_indptr, (const long long*)d_indices,
                                         d_colmap, (long long*)d_indptr_new, (long long*)d_work, st);
}

template <class IT>
static int csr_restrict_fill(long long n, const int* rows, const IT* indptr, const IT* indices, const double* values,
                             const int* colmap, const IT* indptr_new, IT* indices_new, double* values_new, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    for (long long r = 0; r < n; ++r)
        pb_csr_fill_row<IT>(indptr, indices, values, colmap, rows[r], indptr_new[r], indices_new, values_new);
#else
    if (n > 0) {
        const unsigned b = (unsigned)std::min<long long>((n + 7) / 8, 148LL * 32);
        pb_csr_restrict_fill_kernel<IT><<<b, 256, 0, st>>>(n, rows, indptr, indices, values, colmap, indptr_new, indices_new,
                                                          values_new);
    }
    CK(pbLastError());
#endif
    return 0;
}

extern "C" int pb200_csr_restrict_fill(long long nrows_new, const int32_t* d_rows, const void* d_indptr,
                                       const void* d_indices, const double* d_values, int idx_bytes,
                                       const int32_t* d_colmap, const void* d_indptr_new, void* d_indices_new,
                                       double* d_values_new, void* stream) {
    if (nrows_new < 0 || !d_indptr || !d_colmap || !d_indptr_new) return fail(PB200_EINVAL, "null argument");
    if (idx_bytes != 4 && idx_bytes != 8) return fail(PB200_EINVAL, "idx_bytes must be 4 or 8");
    if (int e = use_device_of(d_indptr)) return e;
    pbStream st = (pbStream)stream;
    if (idx_bytes == 4)
        return csr_restrict_fill<int>(nrows_new, d_rows, (const int*)d_indptr, (const int*)d_indices, d_values, d_colmap,
                                      (const int*)d_indptr_new, (int*)d_indices_new, d_values_new, st);
    return csr_restrict_fill<long long>(nrows_new, d_rows, (const long long*)d_indptr, (const long long*)d_indices, d_values,
                                        d_colmap, (const long long*)d_indptr_new, (long long*)d_indices_new, d_values_new, st);
}

template <class IT>
static int csr_matvec(long long n, const IT* indptr, const IT* indices, const double* values, const double* x,
                      const double* y_in, double alpha, double* y_out, pbStream st) {
    ++g_launches;
#ifdef PB_EMULATE
    (void)st;
    for (long long r = 0; r < n; ++r)
        y_out[r] = (y_in ? y_in[r] : 0.0) + alpha * pb_csr_row_dot<IT>(indptr, indices, values, x, r);
#else
    if (n > 0) {
        const unsigned b = (unsigned)std::min<long long>((n + 7) / 8, 148LL * 32);
        pb_csr_matvec_kernel<IT><<<b, 256, 0, st>>>(n, indptr, indices, values, x, y_in, alpha, y_out);
    }
    CK(pbLastError());
#endif
    return 0;
}

extern "C" int pb200_csr_matvec(long long nrows, const void* d_indptr, const void* d_indices, const double* d_values,
                                int idx_bytes, const double* d_x, const double* d_y_in, double alpha, double* d_y_out,
                                void* stream) {
    if (nrows < 0 || !d_indptr || !d_x || !d_y_out) return fail(PB200_EINVAL, "null argument");
    if (idx_bytes != 4 && idx_bytes != 8) return fail(PB200_EINVAL, "idx_bytes must be 4 or 8");
    if (int e = use_device_of(d_indptr)) return e;
    pbStream st = (pbStream)stream;
    if (idx_bytes == 4)
        return csr_matvec<int>(nrows, (const int*)d_indptr, (const int*)d_indices, d_values, d_x, d_y_in, alpha, d_y_out, st);
    return csr_matvec<long long>(nrows, (const long long*)d_indptr, (const long long*)d_indices, d_values, d_x, d_y_in,
                                 alpha, d_y_out, st);
}

extern "C" int pb200_vec_gather(long long n, const int32_t* d_idx, const double* d_in, double* d_out, void* stream) {
    if (n < 0 || (n > 0 && (!d_idx || !d_in || !d_out))) return fail(PB200_EINVAL, "null argument");
    if (int e = use_device_of(d_out)) return e;
    ++g_launches;
#ifdef PB_EMULATE
    (void)stream;
    for (long long k = 0; k < n; ++k) d_out[k] = d_in[d_idx[k]];
#else
    if (n > 0) {
        const unsigned b = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32);
        pb_gather_kernel<<<b, 256, 0, (pbStream)stream>>>(n, d_idx, d_in, d_out);
    }
    CK(pbLastError());
#endif
    return 0;
}

extern "C" int pb200_vec_scatter(long long n, const int32_t* d_idx, const double* d_in, double* d_out, void* stream) {
    if (n < 0 || (n > 0 && (!d_idx || !d_in || !d_out))) return fail(PB200_EINVAL, "null argument");
    if (int e = use_device_of(d_out)) return e;
    ++g_launches;
#ifdef PB_EMULATE
    (void)stream;
    for (long long k = 0; k < n; ++k) d_out[d_idx[k]] = d_in[k];
#else
    if (n > 0) {
        const unsigned b = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32);
        pb_scatter_kernel<<<b, 256, 0, (pbStream)stream>>>(n, d_idx, d_in, d_out);
    }
    CK(pbLastError());
#endif
    return 0;
}

extern "C" int pb200_kron_matvec(int d, const double* const* d_factors, const int* rows, const int* cols,
                                 const double* d_x, double* d_y, double* d_tmp, void* stream) {
    if (d < 1 || d > 8 || !d_factors || !rows || !cols || !d_x || !d_y) return fail(PB200_EINVAL, "invalid argument");
    if (int e = use_device_of(d_y)) return e;
    // apply the factors from the last axis to the first; intermediate shape: (rows[0..k-1] ; cols[k..])
    long long maxsz = 1;
    {
        long long sz = 1;
        for (int k = 0; k < d; ++k) sz *= cols[k];
        maxsz = sz;
        for (int k = d - 1; k >= 0; --k) { sz = sz / cols[k] * rows[k]; maxsz = std::max(maxsz, sz); }
    }
    if (d > 1 && !d_tmp) return fail(PB200_EINVAL, "temporary buffer missing");
    const double* cur = d_x;
    pbStream st = (pbStream)stream;
    for (int k = d - 1; k >= 0; --k) {
        long long outer = 1, inner = 1;
        for (int l = 0; l < k; ++l) outer *= cols[l];
        for (int l = k + 1; l < d; ++l) inner *= rows[l];
        double* dst = (k == 0) ? d_y : (d_tmp + (((d - 1 - k) & 1) ? maxsz : 0));
        k_modek(d_factors[k], rows[k], cols[k], cur, outer, inner, dst, st);
        cur = dst;
    }
    CK(pbLastError());
    return 0;
}

extern "C" int pb200_basis_eval(const double* d_knots, int nknots, int p, const double* d_nodes, int m, int nderiv,
                                int32_t* d_first, double* d_values, void* stream) {
    if (!d_knots || !d_nodes || !d_values) return fail(PB200_EINVAL, "null argument");
    if (p < 0 || p > PB_MAXP) return fail(PB200_EUNSUPPORTED, "spline degree %d above the supported maximum %d", p, PB_MAXP);
    if (nderiv < 0 || nderiv > 2) return fail(PB200_EUNSUPPORTED, "derivatives up to order 2 are supported");
    if (m <= 0) return 0;
    if (int e = use_device_of(d_values)) return e;
    k_basis(d_knots, nknots, p, d_nodes, m, nderiv + 1, d_first, d_values, (pbStream)stream);
    CK(pbLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// FP64 throughput probe
// ------------------------------------------------------------------------------------------------
#ifndef PB_EMULATE
__global__ void __launch_bounds__(256) pb_dfma_kernel(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;
}

#endif

extern "C" int pb200_probe_fp64(int device, int iters, double* gflops) {
    if (!gflops || iters < 1) return fail(PB200_EINVAL, "invalid argument");
#ifdef PB_EMULATE
    (void)device;
    return fail(PB200_EUNSUPPORTED, "no device in the emulation build");
#else
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    CK(cudaMalloc((void**)&sink, 8));
    const int blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    pb_dfma_kernel<<<blocks, 256>>>(iters / 8 + 1, 1.0, sink);      // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        pb_dfma_kernel<<<blocks, 256>>>(iters, 1.0, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * 256.0 * blocks;
        best = std::max(best, fl / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *gflops = best;
    return 0;
#endif
}

#include "distcg_api.cuh"
