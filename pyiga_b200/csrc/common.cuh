// Shared definitions for the pyiga_b200 device code (sm_100a).
//
// Everything in here is plain data: per-axis tables of a tensor-product spline space on its
// Gauss grid, in the layout the kernels consume.  The tables restate what the reference keeps in
// `S*_meshsupp*`, `S*_C*` and `gaussweights*` (pyiga/genericasm.pxi:631-659), but in *compact*
// span-local form: for every Gauss node only the p+1 active functions are stored.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#if !defined(__CUDACC__)
#include <cmath>
using std::fma;
using std::fabs;
struct alignas(16) double2 { double x, y; };     // host build of the shared kernel bodies
#endif

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#define PB_D __device__ __forceinline__
#else
#define PB_HD inline
#define PB_D inline
#endif

#define PB_MAXDIM 3
#define PB_MAXP 8           // highest spline degree the device tables support
#define PB_MAXFIELDS 96      // coefficient fields / terms of a custom form (9 x 9 slot pairs of a 3D fourth-order form fit)
#define PB_MAXTERMS 96
#define PB_MAXPHYS 16        // physical terms / input arrays of the general coefficient program (geo_fields.cuh)

// One tensor axis of a (test, trial) space pair on a common mesh.
// Node index g = s*q + gq  (span s, local Gauss node gq).
struct PbAxis {
    int n;          // non-empty spans
    int q;          // Gauss nodes per span
    int G;          // n*q
    // trial space (matrix columns, "u") and test space (matrix rows, "v")
    int pu, pv;     // degrees
    int Nu, Nv;     // number of B-splines
    int nd;         // number of derivative rows stored (value + nd-1 derivatives)
    int M;          // number of band entries (i,j) with joint support, sorted by i then j
    const double* nodes;    // [G]
    const double* weights;  // [G]
    const int* first_u;     // [n] first active trial function on span s
    const int* first_v;     // [n] first active test function on span s
    const double* Vu;       // [G][nd][pu+1]  values / derivatives of the active trial functions
    const double* Vv;       // [G][nd][pv+1]
    const int* row_start;   // [Nv+1] offset of row i inside the band list
    const int* jmin;        // [Nv]   first trial function interacting with test function i
    const int* supp_u;      // [Nu][2] first span / one-past-last span of the support
    const int* supp_v;      // [Nv][2]
    const int* pair_i;      // [M] row (test) index of band entry mu
    const int* pair_j;      // [M] column (trial) index of band entry mu
    const int* tr;          // [M] band index of the transposed pair (j,i); only when test == trial
    const int* ret_mu;      // [Nv + 2*pv+2][2*pv+1] retire table of the walk kernels (see walk.cuh)
    // forms with second or mixed derivatives only (null otherwise): value, 1st and 2nd derivative
    const double* V3u;      // [G][3][pu+1]
    const double* V3v;      // [G][3][pv+1]
};

// Derivative slots of a basis function.  Slot 0 is the value and slot 1+k the first derivative along
// tensor axis k; a slot >= PB_SLOT_EXT carries one derivative order (0..2) per tensor axis in base 3,
// PB_SLOT_EXT + d0 + 3*d1 + 9*d2  (second and mixed derivatives, e.g. the reference's space-time wave
// form, pyiga/vform.py:1766-1772).
#define PB_SLOT_EXT 16
PB_HD int pb_slot_order(int slot, int axis) {
    if (slot < PB_SLOT_EXT) return slot == 1 + axis ? 1 : 0;
    int c = slot - PB_SLOT_EXT;
    for (int k = 0; k < axis; ++k) c /= 3;
    return c % 3;
}
// the slot with the derivative along `axis` removed, in its shortest encoding
PB_HD int pb_slot_contract(int slot, int axis) {
    int d[3] = {pb_slot_order(slot, 0), pb_slot_order(slot, 1), pb_slot_order(slot, 2)};
    if (axis >= 0) d[axis] = 0;        // axis < 0: only re-encode
    const int total = d[0] + d[1] + d[2];
    if (total == 0) return 0;
    if (total == 1) return d[0] ? 1 : (d[1] ? 2 : 3);
    return PB_SLOT_EXT + d[0] + 3 * d[1] + 9 * d[2];
}

PB_HD int pb_min(int a, int b) { return a < b ? a : b; }
PB_HD int pb_max(int a, int b) { return a > b ? a : b; }
