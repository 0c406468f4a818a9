/* pyiga_b200 — C ABI of the B200 tensor-product IgA assembly path.
 *
 * This is the drop-in boundary: the entry points below are what a binding of the reference
 * (c-f-h/pyiga, a Python/Cython package) attaches to in place of its Cython assembler objects.
 * Every function cites the reference interface it replaces (paths relative to the pyiga source
 * tree).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - all functions return 0 on success and a negative PB200_E* code otherwise; they never throw.
 *     The message of the last failure on the calling thread is returned by pb200_last_error().
 *   - "h_" arguments are host pointers, "d_" arguments are device pointers on the device the handle
 *     was created on; both are caller-owned.  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Calls are asynchronous with respect to the host unless they
 *     write to host memory.
 *   - tensor axes are in z,y,x order as in the reference (pyiga/bspline.py:812): axis 0 is the
 *     slowest axis and the last physical coordinate.
 *   - a handle is bound to one device, is not thread-safe, and different handles may be used from
 *     different threads.
 */
#ifndef PYIGA_B200_H
#define PYIGA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_MAXDIM 3

#if defined(__GNUC__)
#define PB200_API __attribute__((visibility("default")))
#else
#define PB200_API
#endif

enum {
    PB200_OK = 0,
    PB200_EINVAL = -1,      /* bad argument (message says which) */
    PB200_ECUDA = -2,       /* CUDA runtime error */
    PB200_ENOMEM = -3,      /* workspace too small / allocation failed */
    PB200_EUNSUPPORTED = -4 /* configuration has no device implementation */
};

/* built-in bilinear forms (pyiga/assemblers.pyx: MassAssembler{2,3}D :26,:1158;
 * StiffnessAssembler{2,3}D :174,:1324) */
enum { PB200_FORM_MASS = 1, PB200_FORM_STIFFNESS = 2, PB200_FORM_CUSTOM = 100 };

typedef struct pb200_assembler pb200_assembler;
typedef struct pb200_mlstruct pb200_mlstruct;   /* multi-level band structure on the device */

/* One tensor axis of the discretisation: trial space (matrix columns, `kvs0` of the reference)
 * and test space (rows, `kvs1`); both must share the mesh (pyiga/assemblers.pyx:1343).
 * `nodes`/`weights` are the iterated Gauss rule of pyiga/quadrature.py:3-21, nq nodes per span. */
typedef struct {
    int p_trial;
    int nknots_trial;
    const double* h_knots_trial;
    int p_test;                 /* ignored when h_knots_test == NULL (test space == trial space) */
    int nknots_test;
    const double* h_knots_test;
    int nq;                     /* Gauss nodes per span (reference: max_k p_k + 1) */
    const double* h_nodes;      /* [numspans * nq] */
    const double* h_weights;    /* [numspans * nq] */
} pb200_axis_desc;

/* Spline geometry map (pyiga/bspline.py:827 BSplineFunc, pyiga/geometry.py:27 NurbsFunc):
 * knot vectors per axis and the control net in C order [N0][N1][N2][dim (+1)].  For rational
 * maps the coordinates are premultiplied by the weight, which is the last component — exactly
 * the `coeffs` attribute of the reference's NurbsFunc. */
typedef struct {
    int sdim, dim;
    int rational;
    int p[PB200_MAXDIM];
    int nknots[PB200_MAXDIM];
    const double* h_knots[PB200_MAXDIM];
    const double* h_coeffs;
} pb200_geo_desc;

/* One term of a custom form:  integral of  C(field) * d^{slot_test} v * d^{slot_trial} u.
 * slot 0 = function value, slot 1+k = derivative along tensor axis k (parametric).
 * Second and mixed derivatives (the reference's numderiv = 2 slots, pyiga/vform.py:1766-1772):
 * slot = PB200_SLOT_EXT + d0 + 3*d1 + 9*d2 with d_k in 0..2 the derivative order along tensor axis k.
 * Linear forms (arity 1, load vectors) set slot_trial = -1 in every term. */
#define PB200_SLOT_EXT 16
typedef struct { int field, slot_test, slot_trial; } pb200_term;

/* One physical coefficient term of a general first-order scalar form (see
 * pb200_asm_compute_fields_general): slots 0 = value, 1+a = derivative w.r.t. physical coordinate a
 * (x,y,z order); input = index of a coefficient array on the Gauss grid, or -1 for the constant 1;
 * the coefficient is scale * input. */
typedef struct { int slot_test, slot_trial, input; double scale; } pb200_phys_term;

typedef struct {
    int dim;
    pb200_axis_desc axis[PB200_MAXDIM];
    int form;                   /* PB200_FORM_* */
    int nfields;                /* CUSTOM only: number of coefficient fields (caller uploads them) */
    int nterms;                 /* CUSTOM only */
    const pb200_term* terms;    /* CUSTOM only */
    int symmetric;              /* CUSTOM only: form is symmetric and test == trial */
} pb200_desc;

typedef struct {
    int dim;
    int ndofs_test[PB200_MAXDIM], ndofs_trial[PB200_MAXDIM];
    int nnodes[PB200_MAXDIM];   /* Gauss nodes per axis */
    int nband[PB200_MAXDIM];    /* band entries per axis (len(bidx[k])) */
    int nfields;
    int fast_path;              /* 1 if the sum-factorised pipeline has an instantiation */
    long long nnz;              /* prod nband */
    long long npoints;          /* prod nnodes */
} pb200_info;

PB200_API int pb200_version(void);
PB200_API const char* pb200_last_error(void);
/* number of kernels this library has launched so far (benchmark bookkeeping) */
PB200_API long long pb200_launch_count(void);

/* ---- assembler object -------------------------------------------------------------------------
 * Replaces the constructor of the assembler classes (pyiga/assemblers.pyx:1336-1383): builds the
 * span/support/band tables on the host, uploads them, and evaluates the 1D basis tables on the
 * device (K1; replaces compute_values_derivs, pyiga/assemble_tools.py:7-12). */
PB200_API int pb200_asm_create(const pb200_desc* desc, int device, void* stream, pb200_assembler** out);
PB200_API int pb200_asm_destroy(pb200_assembler* a);
PB200_API int pb200_asm_info(const pb200_assembler* a, pb200_info* info);

/* re-run K1 (basis tables) — lets a benchmark keep the table evaluation inside the timed region */
PB200_API int pb200_asm_tabulate(pb200_assembler* a, void* stream);

/* Band structure of one axis: h_bidx receives nband x 2 uint32 (i,j), identical to
 * MLStructure.from_kvs(...).bidx[axis] (pyiga/mlmatrix.py:59-65, :420-440). */
PB200_API int pb200_asm_structure(const pb200_assembler* a, int axis, uint32_t* h_bidx);

/* Geometry + coefficient fields (K2).  `d_fields` is caller-allocated, nfields*npoints doubles,
 * layout [field][g0][g1][g2]; the handle keeps the pointer (no copy).  Replaces
 * geo.grid_jacobian + precompute_fields (pyiga/bspline.py:897-921, geometry.py:116-123,
 * assemblers.pyx:1389-1449). */
PB200_API int pb200_asm_bind_fields(pb200_assembler* a, double* d_fields);
/* Bind a spline geometry WITHOUT evaluating the fields: control net and knots are uploaded and the
 * 1D geometry basis tables are evaluated at the Gauss nodes (K1).  For 3D mass / stiffness the
 * fused stage 1 of pb200_asm_assemble_mlb then evaluates grid_jacobian + precompute_fields
 * (same reference code as above) in registers, point by point, while it contracts the first axis
 * — the field array is never materialised.  pb200_asm_uses_fused_fields tells whether the next
 * pb200_asm_assemble_mlb takes that path (1) or needs pb200_asm_compute_fields* first (0);
 * after pb200_asm_set_geometry the latter may be called with geo = NULL. */
PB200_API int pb200_asm_set_geometry(pb200_assembler* a, const pb200_geo_desc* geo, void* stream);
PB200_API int pb200_asm_uses_fused_fields(const pb200_assembler* a);
PB200_API int pb200_asm_compute_fields(pb200_assembler* a, const pb200_geo_desc* geo, void* stream);
/* same, restricted to the Gauss planes that the rows [row0_begin,row0_end) of the first axis see
 * (slab-sharded assembly: every rank evaluates only its planes plus the p-span overlap) */
PB200_API int pb200_asm_compute_fields_slab(pb200_assembler* a, const pb200_geo_desc* geo, int row0_begin,
                                            int row0_end, void* stream);
/* Fields of a PB200_FORM_CUSTOM assembler for a general scalar form with at most first derivatives,
 *     a(u,v) = int sum_t c_t(x) d^{slot_test} v d^{slot_trial} u dx   (physical derivatives),
 * pulled back with J^-1 and weighted with GaussWeight*|det J| on the device; field f receives the
 * coefficient of the parametric slot pair of the (unique) term that uses field f.  Replaces the
 * generated precompute_fields of compiled vforms (pyiga/codegen/cython.py:673-701).  d_inputs are
 * coefficient arrays on the full Gauss grid (user callables are evaluated on the host by the
 * caller, exactly as pyiga does, pyiga/codegen/cython.py:465-484).  Give either `geo` or `d_jac`;
 * row0_begin < 0 selects the whole grid.
 * This call is also the ABI form of the generated `update(name=func)` / `update_params(name=value)`
 * (pyiga/codegen/cython.py:703-744): an updated input is a new coefficient array in `d_inputs`, an updated
 * parameter a new `scale`; the name -> array bookkeeping stays in the host layer (Assembler.update,
 * pyiga/assemble.py:984-994), nothing else of the assembler is rebuilt. */
PB200_API int pb200_asm_compute_fields_general(pb200_assembler* a, const pb200_geo_desc* geo, const double* d_jac,
                                               int nphys, const pb200_phys_term* phys, int ninputs,
                                               const double* const* d_inputs, int row0_begin, int row0_end,
                                               void* stream);
/* same, from Jacobians the caller evaluated on the Gauss grid (geometry objects that are not
 * splines): d_jac is [npoints][dim][dim] */
PB200_API int pb200_asm_compute_fields_from_jacobian(pb200_assembler* a, const double* d_jac, void* stream);

/* Geometry evaluation on the assembler's Gauss grid or on an arbitrary tensor grid: writes
 * [npoints][dim][dim] Jacobians and/or [npoints][dim] values (either pointer may be NULL).
 * Replaces BSplineFunc/NurbsFunc.grid_eval / grid_jacobian. */
PB200_API int pb200_geo_eval_grid(const pb200_geo_desc* geo, const int* npts, const double* const* h_grid,
                        double* d_values, double* d_jac, int device, void* stream);

/* ---- assembly ----------------------------------------------------------------------------------
 * Rows of the first tensor axis in [row0_begin,row0_end) form a slab; its MLB values are the
 * contiguous block data[bidx0 offset of row0_begin .. of row0_end][:][:].
 * Replaces assemble_entries' S.nonzero() + asm.multi_entries(IJ) for the whole pattern
 * (pyiga/assemble.py:741-745) with the sum-factorised pipeline; output layout is MLMatrix.data
 * (pyiga/mlmatrix.py:201-269). */
PB200_API int pb200_asm_workspace_bytes(const pb200_assembler* a, int row0_begin, int row0_end, size_t* bytes);
PB200_API int pb200_asm_assemble_mlb(pb200_assembler* a, int row0_begin, int row0_end, double* d_out,
                           void* d_work, size_t work_bytes, void* stream);
/* same result through the per-entry quadrature kernel (reference algorithm, for cross-checks and
 * configurations without a fast path) */
PB200_API int pb200_asm_assemble_mlb_entrywise(pb200_assembler* a, int row0_begin, int row0_end, double* d_out,
                                     void* stream);

/* Per-kernel timing of the pipeline (for the roofline report): when enabled, CUDA events are
 * recorded on the launch stream around every stage of the next assemble call;
 * pb200_asm_get_timing waits for them and returns the durations in ms and the ';'-joined names. */
PB200_API int pb200_asm_set_timing(pb200_assembler* a, int enable);
/* tuning / test switches: "force_walk" = 1 disables the warp-per-line kernels of the final stage */
PB200_API int pb200_asm_set_option(pb200_assembler* a, const char* name, int value);
PB200_API int pb200_asm_get_timing(pb200_assembler* a, int max_stages, float* ms, char* names, int names_len,
                                   int* nstages);

/* Linear forms (arity 1): the load vector, ndofs doubles in C order of the test space.  Replaces
 * BaseAssembler*.assemble_vector (pyiga/genericasm.pxi:129-145, 762-778). */
PB200_API int pb200_asm_vector_workspace_bytes(const pb200_assembler* a, size_t* bytes);
PB200_API int pb200_asm_assemble_vector(pb200_assembler* a, double* d_out, void* d_work, size_t work_bytes,
                                        void* stream);

/* multi_entries(indices) (pyiga/genericasm.pxi:722-758): d_ij is n x 2 uint64 (row, column),
 * d_out n doubles; pairs outside the pattern give 0.0. */
PB200_API int pb200_asm_multi_entries(pb200_assembler* a, const uint64_t* d_ij, size_t n, double* d_out, void* stream);

/* ---- MLB matrices ------------------------------------------------------------------------------
 * A structure handle holds the per-level row tables of an MLStructure on the device.  The one of
 * an assembler is borrowed (owned by the assembler); stand-alone ones are built from the bidx
 * arrays of pyiga's MLStructure (pyiga/mlmatrix.py:15-58; levels must be sorted by row with
 * contiguous column ranges, which holds for every spline pattern) and must be destroyed. */
PB200_API const pb200_mlstruct* pb200_asm_mlstruct(const pb200_assembler* a);
PB200_API int pb200_mlstruct_create(int nlevels, const int* rows, const int* cols, const int* nband,
                                    const uint32_t* const* h_bidx, int device, pb200_mlstruct** out);
PB200_API int pb200_mlstruct_destroy(pb200_mlstruct* s);
/* CSR export of a slab (replaces ml_nonzero_* + COO->CSR, pyiga/mlmatrix_cy.pyx:189-289,
 * pyiga/assemble.py:745): idx_bytes is 4 or 8; d_indptr has nrows+1 entries relative to the slab.
 * d_indptr and d_indices may both be NULL: only the values are permuted. */
PB200_API int pb200_mlb_to_csr(const pb200_mlstruct* s, int row0_begin, int row0_end, const double* d_mlb,
                     void* d_indptr, void* d_indices, double* d_values, int idx_bytes, void* stream);
/* host-only helper (no CUDA call, `nthreads` host threads): the CSR pattern (indptr with nrows+1
 * entries, first entry = indptr_offset; sorted column indices) of the same slab, from the per-level
 * row tables (row_start[m+1], jmin[m]).  It is the closed form of ml_nonzero_* + COO->CSR
 * (pyiga/mlmatrix_cy.pyx:189-289, pyiga/assemble.py:745) and lets a caller who wants the matrix in
 * host memory produce the integer arrays there while the values (pb200_mlb_to_csr with null
 * d_indptr / d_indices) are still being computed and copied: 8 instead of 12 B/nnz cross PCIe. */
PB200_API int pb200_csr_pattern_host(int nlevels, const int* rows, const int* cols, const int* nband,
                           const int* const* h_row_start, const int* const* h_jmin, int row0_begin,
                           int row0_end, void* h_indptr, void* h_indices, int idx_bytes,
                           long long indptr_offset, int nthreads);
/* y = A x for the slab rows (replaces ml_matvec_2d/3d, pyiga/mlmatrix_cy.pyx:224-325).  d_x starts
 * at trial index x_j0_begin on axis 0 (use 0 for a full vector). */
PB200_API int pb200_mlb_matvec(const pb200_mlstruct* s, int row0_begin, int row0_end, const double* d_mlb,
                     const double* d_x, int x_j0_begin, double* d_y, void* stream);
/* y = (A_0 (x) ... (x) A_{d-1}) x with dense row-major factors (pyiga/kronecker.py:15-34);
 * d_tmp holds two buffers of max intermediate size. */
PB200_API int pb200_kron_matvec(int d, const double* const* d_factors, const int* rows, const int* cols,
                      const double* d_x, double* d_y, double* d_tmp, void* stream);

/* ---- partial-row assembly --------------------------------------------------------------------------
 * Replaces _assemble_partial_rows (pyiga/_hdiscr.py:5-12): MLStructure.nonzeros_for_rows
 * (pyiga/mlmatrix.py:150-185) + multi_entries + COO->CSR for a list of matrix rows (hierarchical /
 * adaptive callers).  Phase 1 (host, integer closed form): h_indptr[0..n] = prefix sums of the
 * pattern sizes of the rows.  Phase 2 (device): sorted column indices and per-entry quadrature
 * values of every row; d_indptr is the uploaded phase-1 result in the chosen integer width. */
PB200_API int pb200_asm_rows_count(const pb200_assembler* a, const int64_t* h_rows, long long n, int64_t* h_indptr);
PB200_API int pb200_asm_rows_fill(pb200_assembler* a, const int64_t* d_rows, long long n, const void* d_indptr,
                        void* d_indices, double* d_values, int idx_bytes, void* stream);

/* ---- elimination of constrained dofs on device CSR arrays ----------------------------------------
 * Replaces the selection-matrix products of RestrictedLinearSystem (pyiga/assemble.py:575-652):
 *   A_r = R_free_v A R_free^T  (restrict_matrix, :632-637),  b_r = R_free_v (b - A R_elim^T values) (:616).
 * d_rows: old row index of every kept row (nrows_new, increasing); d_colmap: new column index of every
 * old column or -1 if eliminated.  idx_bytes (4 or 8) is the integer width of all CSR index arrays.
 * Phase 1 writes the new indptr (nrows_new+1 entries; the last one is the new nnz, read it back to
 * size the outputs); phase 2 compacts indices and values.  Order inside rows is preserved. */
PB200_API int pb200_csr_restrict_workspace(long long nrows_new, int idx_bytes, size_t* bytes);
PB200_API int pb200_csr_restrict_count(long long nrows_new, const int32_t* d_rows, const void* d_indptr,
                             const void* d_indices, int idx_bytes, const int32_t* d_colmap,
                             void* d_indptr_new, void* d_work, size_t work_bytes, void* stream);
PB200_API int pb200_csr_restrict_fill(long long nrows_new, const int32_t* d_rows, const void* d_indptr,
                            const void* d_indices, const double* d_values, int idx_bytes,
                            const int32_t* d_colmap, const void* d_indptr_new, void* d_indices_new,
                            double* d_values_new, void* stream);
/* y_out = (d_y_in ? y_in : 0) + alpha * A x for device CSR arrays (the rhs update above; scipy's
 * csr_matrix.dot on the reference side) */
PB200_API int pb200_csr_matvec(long long nrows, const void* d_indptr, const void* d_indices, const double* d_values,
                     int idx_bytes, const double* d_x, const double* d_y_in, double alpha, double* d_y_out,
                     void* stream);
/* out[k] = in[idx[k]] / out[idx[k]] = in[k]  (restrict, extend, complete; pyiga/assemble.py:618-652) */
PB200_API int pb200_vec_gather(long long n, const int32_t* d_idx, const double* d_in, double* d_out, void* stream);
PB200_API int pb200_vec_scatter(long long n, const int32_t* d_idx, const double* d_in, double* d_out, void* stream);

/* K1 stand-alone (replaces bspline.active_deriv / collocation_derivs_info,
 * pyiga/bspline_cy.pyx:126-145, pyiga/bspline.py:648-660): d_first m int32, d_values [m][nderiv+1][p+1] */
PB200_API int pb200_basis_eval(const double* d_knots, int nknots, int p, const double* d_nodes, int m, int nderiv,
                     int32_t* d_first, double* d_values, void* stream);

/* FP64 FMA throughput probe used for the roofline denominator: runs `iters` dependent-chain DFMA
 * rounds on every SM and returns the measured GFLOP/s through *gflops. */
PB200_API int pb200_probe_fp64(int device, int iters, double* gflops);

/* host-only helper (no CUDA call): band structure of a pair of knot vectors; h_bidx may be NULL to
 * query the count.  Used by MLStructure.from_kvs on the host side. */
PB200_API int pb200_band_structure(const double* h_knots_trial, int nknots_trial, int p_trial,
                         const double* h_knots_test, int nknots_test, int p_test,
                         uint32_t* h_bidx, int* nband);

/* ---- slab-distributed preconditioned CG (csrc/distcg.cuh) ------------------------------------------------
 * Replaces scipy.sparse.linalg.cg(M, b, M=KroneckerOperator(*Minvs)) of pyiga/approx.py:82-96 with the operator
 * of pyiga/mlmatrix_cy.pyx:295-325 and the preconditioner of pyiga/kronecker.py:15-34, for a matrix whose rows
 * of the first tensor axis are sharded over the GPUs of one node.  All ranks of the solve share WINDOWS of
 * device memory (CUDA IPC: peer loads / stores over NVLink); halo planes, dot products and the gathered
 * preconditioner input travel through them inside the kernels — no NCCL call, no host in the iteration. */
typedef struct pb200_comm pb200_comm;
typedef struct pb200_cg pb200_cg;
PB200_API int pb200_comm_create(int device, int rank, int world, size_t bytes, pb200_comm** out);
PB200_API int pb200_comm_handle(pb200_comm* c, void* handle64);              /* 64 bytes, to be sent to the peers */
PB200_API int pb200_comm_open_peers(pb200_comm* c, const void* handles);     /* world x 64 bytes in rank order */
PB200_API int pb200_comm_destroy(pb200_comm* c);
/* window size of a solver: cuts[world+1] = row slabs of the first axis, halo = planes the band reaches into a
 * neighbouring slab (the degree for spline spaces) */
PB200_API int pb200_cg_window_bytes(const pb200_mlstruct* S, int world, const int* cuts, int halo, size_t* bytes);
/* d_mlb: value tensor of the local slab; d_Ainv[3]: dense row-major inverses of the Kronecker factors (device);
 * comm: window with the peers opened (NULL with world == 1) */
PB200_API int pb200_cg_create(const pb200_mlstruct* S, int rank, int world, const int* cuts, int halo,
                              const double* d_mlb, const double* const* d_Ainv, pb200_comm* comm, pb200_cg** out);
PB200_API int pb200_cg_destroy(pb200_cg* g);
/* x0 = 0; stops when ||r|| <= rtol ||b||; the host looks at the device-side flag every check_every iterations
 * (one CUDA graph launch per batch) */
PB200_API int pb200_cg_solve(pb200_cg* g, const double* d_b_local, double* d_x_local, double rtol, int maxiter,
                             int check_every, int* iters, double* relres, void* stream);
/* y_local = A_slab p: the halo-exchanging matvec alone (ml_matvec_3d of pyiga/mlmatrix_cy.pyx:295-325) */
PB200_API int pb200_cg_matvec(pb200_cg* g, const double* d_p_local, double* d_y_local, void* stream);

#ifdef __cplusplus
}
#endif
#endif
