// Stage plans of the built-in forms for the walk kernels (see walk.cuh).
//
// A plan lists, for one contraction stage, which input terms are combined with which derivative
// flags into which output terms.  The plans below encode mass and stiffness; they exploit that
// for these forms test space == trial space and the coefficient tensor is symmetric, so terms that
// are transposes of each other are computed once and read back through the transposed band index.
//
// Derivative bookkeeping (stiffness, 3D, tensor axes 0,1,2 = z,y,x): the integrand is
//   sum_{k,l} C[k][l] d_k v d_l u,   C[k][l] = B[2-k][2-l],  B = W J^-1 J^-T
// with B stored symmetric-packed [B00,B01,B02,B11,B12,B22] (pyiga/assemblers.pyx:1443-1449).
// A term is named (test slot, trial slot) with slots v = value, dk = derivative on axis k; once an
// axis is contracted its derivative slot turns into v.
//
//  stage 1 (axis 0):  fields -> X1 = { (v,v), (v,d1), (v,d2), (d1,d1), (d1,d2), (d2,d2) }
//      S1A:  B22 [1,1]->(v,v)   B12 [1,0]->(v,d1)   B02 [1,0]->(v,d2)
//      S1B:  B11 [0,0]->(d1,d1) B01 [0,0]->(d1,d2)  B00 [0,0]->(d2,d2)
//      derived:  (d1,v)[mu0] = (v,d1)[mu0^T],  (d2,v)[mu0] = (v,d2)[mu0^T],  (d2,d1) = (d1,d2)
//  stage 2 (axis 1):  X1 -> X2 = { (v,v), (v,d2), (d2,d2) }
//      FINAL4: (v,v)[0,0] + (v,d1)[0,1] + (v,d1)^T[1,0] + (d1,d1)[1,1]  -> (v,v)
//      S2B:    (v,d2)[0,0] + (d1,d2)[1,0] -> (v,d2) ;   (d2,d2)[0,0] -> (d2,d2)
//      derived:  (d2,v)[mu0,mu1] = (v,d2)[mu0^T,mu1^T]
//  stage 3 (axis 2):  FINAL4: (v,v)[0,0] + (v,d2)[0,1] + (v,d2)^T[1,0] + (d2,d2)[1,1] -> K
#pragma once
#include "walk.cuh"

enum PbPlanId {
    PB_PLAN_COPY = 0,   // one term, values only (mass, every stage)
    PB_PLAN_FINAL4 = 1, // [0,0] + [0,1] + [1,0]^T + [1,1] -> one output
    PB_PLAN_S1A = 2,
    PB_PLAN_S1B = 3,
    PB_PLAN_S2B = 4,
    PB_PLAN_S1_2D = 5,  // 2D stiffness stage 1:  B11 [1,1]->(v,v)  B01 [1,0]->(v,d1)  B00 [0,0]->(d1,d1)
    PB_PLAN_GEN4 = 6,   // generic forms: [0,0] + [0,1] + [1,0] + [1,1] -> one output, no transposes, null = absent
    PB_PLAN_ONE11 = 7,  // one term, test and trial derivative on this axis
    PB_PLAN_ONE10 = 8,  // one term, test derivative on this axis
    PB_PLAN_PAIRT = 9,  // [0,0] + [1,0] -> one output
    PB_PLAN_S1F = 10,       // fused stage 1 of 3D stiffness: geometry + fields in registers -> all six X1 terms (walk_geo.cuh)
    PB_PLAN_S1F_MASS = 11,  // fused stage 1 of 3D mass
    PB_PLAN_COUNT = 12,
    PB_PLAN_LANE_BASE = 1000    // + plan id: the lane-per-span variant (second argument = lines per warp)
};

struct PbPlanCopy {
    static constexpr int NOPS = 1, NOUT = 1, MINB = 4, NPF = 4;
    static constexpr int MINB4 = 4;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int) { return PbOp{0, 0, 0, 0, 0}; }
};
struct PbPlanOne11 {
    static constexpr int NOPS = 1, NOUT = 1, MINB = 4, NPF = 4;
    static constexpr int MINB4 = 4;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int) { return PbOp{0, 0, 1, 1, 0}; }
};
struct PbPlanOne10 {
    static constexpr int NOPS = 1, NOUT = 1, MINB = 4, NPF = 4;
    static constexpr int MINB4 = 4;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int) { return PbOp{0, 0, 1, 0, 0}; }
};
struct PbPlanPairT {
    static constexpr int NOPS = 2, NOUT = 1, MINB = 4, NPF = 3;
    static constexpr int MINB4 = 4;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[2] = {{0, 0, 0, 0, 0}, {1, 0, 1, 0, 0}};
        return t[i];
    }
};
struct PbPlanFinal4 {
    static constexpr int NOPS = 4, NOUT = 1, MINB = 3, NPF = 2;
    static constexpr int MINB4 = 3;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = true;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[4] = {{0, 0, 0, 0, 0}, {1, 0, 0, 1, 0}, {1, 1, 1, 0, 0}, {2, 0, 1, 1, 0}};
        return t[i];
    }
};
struct PbPlanGen4 {
    static constexpr int NOPS = 4, NOUT = 1, MINB = 3, NPF = 2;
    static constexpr int MINB4 = 3;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[4] = {{0, 0, 0, 0, 0}, {1, 0, 0, 1, 0}, {2, 0, 1, 0, 0}, {3, 0, 1, 1, 0}};
        return t[i];
    }
};
struct PbPlanS1A {
    static constexpr int NOPS = 3, NOUT = 3, MINB = 3, NPF = 3;
    static constexpr int MINB4 = 2;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[3] = {{0, 0, 1, 1, 0}, {1, 0, 1, 0, 1}, {2, 0, 1, 0, 2}};
        return t[i];
    }
};
struct PbPlanS1B {
    static constexpr int NOPS = 3, NOUT = 3, MINB = 3, NPF = 3;
    static constexpr int MINB4 = 2;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[3] = {{0, 0, 0, 0, 0}, {1, 0, 0, 0, 1}, {2, 0, 0, 0, 2}};
        return t[i];
    }
};
struct PbPlanS2B {
    static constexpr int NOPS = 3, NOUT = 2, MINB = 3, NPF = 3;
    static constexpr int MINB4 = 2;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[3] = {{0, 0, 0, 0, 0}, {1, 0, 1, 0, 0}, {2, 0, 0, 0, 1}};
        return t[i];
    }
};
struct PbPlanS1_2D {
    static constexpr int NOPS = 3, NOUT = 3, MINB = 2, NPF = 3;
    static constexpr int MINB4 = 2;   // resident blocks asked for when P >= 4 (bigger windows)
    static constexpr bool HAS_TR = false;
    static constexpr bool sym(int) { return false; }
    static constexpr PbOp op(int i) {
        constexpr PbOp t[3] = {{0, 0, 1, 1, 0}, {1, 0, 1, 0, 1}, {2, 0, 0, 0, 2}};
        return t[i];
    }
};

// Fused stage 1 (walk_geo.cuh): the inputs are not read but evaluated from the geometry; `field(i)`
// is the index of op i's input in the output of the field program.  Outputs whose op has ft == fu
// have a symmetric window (10 instead of 16 accumulators for p = 3).
struct PbPlanS1F {
    static constexpr int NOPS = 6, NOUT = 6, MINB = 2, NPF = 1;
    static constexpr int MINB4 = 1;
    static constexpr bool HAS_TR = false;
    static constexpr PbOp op(int i) {
        constexpr PbOp t[6] = {{0, 0, 1, 1, 0}, {1, 0, 1, 0, 1}, {2, 0, 1, 0, 2}, {3, 0, 0, 0, 3}, {4, 0, 0, 0, 4}, {5, 0, 0, 0, 5}};
        return t[i];
    }
    static constexpr int field(int i) {     // B22, B12, B02, B11, B01, B00 of the symmetric-packed B
        constexpr int f[6] = {5, 4, 2, 3, 1, 0};
        return f[i];
    }
    static constexpr bool sym(int o) { return o == 0 || o >= 3; }
};
struct PbPlanS1FMass {
    static constexpr int NOPS = 1, NOUT = 1, MINB = 3, NPF = 1;
    static constexpr int MINB4 = 2;
    static constexpr bool HAS_TR = false;
    static constexpr PbOp op(int) { return PbOp{0, 0, 0, 0, 0}; }
    static constexpr int field(int) { return 0; }
    static constexpr bool sym(int) { return true; }
};

// launcher registry (filled by the per-(P,Q) translation units and by JIT-compiled form modules)
typedef int (*PbWalkLaunch)(const PbWalkParams* prm, int use_smem, size_t smem_bytes, void* stream);
extern "C" __attribute__((visibility("default"))) void pb200_register_walk(int plan_id, int P, int Q, PbWalkLaunch fn);
PbWalkLaunch pb_find_walk(int plan_id, int P, int Q);
