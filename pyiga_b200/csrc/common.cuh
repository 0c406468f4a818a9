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
#define PB_MAXFIELDS 16
#define PB_MAXTERMS 16

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
};

PB_HD int pb_min(int a, int b) { return a < b ? a : b; }
PB_HD int pb_max(int a, int b) { return a > b ? a : b; }
