// Instantiates the walk kernels of the built-in plans for one (degree, nodes-per-span) pair.
// Compiled once per pair with -DPB_P=<p> -DPB_Q=<q>; each object registers its launchers.
#include <algorithm>
#include <cstring>
#include <vector>
#include "backend.cuh"
#include "plans.cuh"
#include "walk1.cuh"
#include "walk_geo.cuh"
#include "fused23.cuh"

#ifndef PB_P
#error "compile with -DPB_P=<degree> -DPB_Q=<nodes per span>"
#endif

#ifndef PB_LANE_NST
#define PB_LANE_NST 4     // ring depth of the lane-span kernel (3 stages = 3 resident blocks per SM measured the same)
#endif

namespace {

template <class Plan>
int launch(const PbWalkParams* prm, int use_smem, size_t smem_bytes, void* stream) {
#ifdef PB_EMULATE
    (void)use_smem; (void)smem_bytes; (void)stream;
    for (int y = 0; y < std::max(1, prm->nsplit); ++y)
        pb_emu_for(prm->nthreads, [&](long long tid) { pb_walk_line<Plan, PB_P, PB_Q>(*prm, tid, prm->V2, y); });
    return 0;
#else
    auto kern = pb_walk_kernel<Plan, PB_P, PB_Q, (PB_P >= 4 ? Plan::MINB4 : Plan::MINB), Plan::NPF>;
    // the opt-in to large dynamic shared memory is a per-device function attribute
    static unsigned long long configured = 0;       // bit d: done on device d
    int devno = 0;
    cudaGetDevice(&devno);
    if (devno >= 64 || !((configured >> devno) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        if (devno < 64) configured |= 1ull << devno;
    }
    // the asynchronous loader's ring (plans with NPF >= 2) sits behind the table slice
    const size_t ring = Plan::NPF >= 2 ? (size_t)Plan::NPF * PB_Q * Plan::NOPS * 128 * sizeof(double) : 0;
    if (use_smem < 0) {
        // query: resident blocks per SM with `smem_bytes` of table staging (tables + integer tables)
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, smem_bytes + ring);
        return e == cudaSuccess ? nb : 0;
    }
    const long long blocks = (prm->nthreads + 127) / 128;
    if (blocks <= 0) return 0;
    if (prm->out_smu >= (1LL << 31)) return (int)cudaErrorInvalidValue;   // band stride is used as a 32-bit factor
    // table slice and integer tables (spans padded to 4 + retired functions x (2P+1)) of the largest piece
    const int ny = std::max(1, prm->nsplit);
    size_t vpad = 0, ipad = 0;
    for (int y = 0; y < ny; ++y) {
        const PbWalkRange rg = pb_walk_range(*prm, y);
        const int nsp = rg.s_end - rg.s_begin;
        vpad = std::max(vpad, ((size_t)nsp * PB_Q * 2 * (PB_P + 1) * sizeof(double) + 127) & ~size_t(127));
        ipad = std::max(ipad, (((size_t)((nsp + 3) & ~3) + (size_t)(rg.f_hi - rg.f_lo) * (2 * PB_P + 1)) * sizeof(int) + 127) & ~size_t(127));
    }
    if (!use_smem) vpad = ipad = 0;
    (void)smem_bytes;
    dim3 grid((unsigned)blocks, (unsigned)ny);
    kern<<<grid, 128, vpad + ipad + ring, (cudaStream_t)stream>>>(*prm, use_smem);
    return (int)cudaGetLastError();
#endif
}

template <class Plan>
int launch_lane(const PbWalkParams* prm, int lines_per_warp, size_t, void* stream) {
    const int nb = pb_lane_batches(prm->s_end - prm->s_begin, PB_P);
#ifdef PB_EMULATE
    (void)stream; (void)lines_per_warp;
    for (int b = 0; b < nb; ++b)
        pb_emu_for(prm->nthreads, [&](long long line) { pb_lane_span_seq<Plan, PB_P, PB_Q>(*prm, line, b); });
    return 0;
#else
    const bool v1 = lines_per_warp < 0;             // negative: the register-prefetch version (v1)
    if (v1) lines_per_warp = -lines_per_warp;
    const long long warps = (prm->nthreads + lines_per_warp - 1) / lines_per_warp;
    const long long blocks = (warps + 3) / 4;
    if (blocks <= 0) return 0;
    dim3 grid((unsigned)blocks, (unsigned)nb);
    if (v1) {
        pb_lane_span_kernel<Plan, PB_P, PB_Q><<<grid, 128, 0, (cudaStream_t)stream>>>(*prm, lines_per_warp);
        return (int)cudaGetLastError();
    }
    constexpr int NST = PB_LANE_NST;
    if (lines_per_warp > PbLaneCfg<PB_P, PB_Q>::DQ) return (int)cudaErrorInvalidValue;
    using Cfg = PbLaneCfg<PB_P, PB_Q>;
    const size_t smem = 4 * (size_t)(NST * Plan::NOPS * Cfg::SEG + Cfg::OUTPAD + Cfg::LOSLOTS) * sizeof(double);
    auto kern = pb_lane_span_kernel_v2<Plan, PB_P, PB_Q, NST>;
    static unsigned long long configured = 0;       // bit d: done on device d (per-device attribute)
    int devno = 0;
    cudaGetDevice(&devno);
    if (devno >= 64 || !((configured >> devno) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (devno < 64) configured |= 1ull << devno;
    }
    kern<<<grid, 128, smem, (cudaStream_t)stream>>>(*prm, lines_per_warp);
    return (int)cudaGetLastError();
#endif
}

// fused stage 1 (walk_geo.cuh): geometry and fields are evaluated by the walk itself
template <class Plan, class Prog, int NC>
int launch_geo_nc(const PbWalkParams* prm, const PbGeoLineParams& gp, int use_smem, void* stream) {
#ifdef PB_EMULATE
    (void)use_smem; (void)stream;
    for (int y = 0; y < std::max(1, prm->nsplit); ++y)
        pb_emu_for(prm->nthreads, [&](long long tid) { pb_walk_geo_line<Plan, PB_P, PB_Q, NC, Prog>(*prm, gp, tid, y); });
    return 0;
#else
    auto kern = pb_walk_geo_kernel<Plan, PB_P, PB_Q, NC, Prog, (PB_P >= 4 ? Plan::MINB4 : Plan::MINB)>;
    static unsigned long long configured = 0;       // bit d: done on device d
    int devno = 0;
    cudaGetDevice(&devno);
    if (devno >= 64 || !((configured >> devno) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        if (devno < 64) configured |= 1ull << devno;
    }
    const int ny = std::max(1, prm->nsplit);
    size_t smem = 0;
    for (int y = 0; y < ny; ++y) {
        size_t vb, ib, gb, zb;
        pb_walk_geo_smem<PB_P, PB_Q>(pb_walk_range(*prm, y), gp.geo.pg[0], gp.geo.Ng[0] * PbGeoLoader<Plan, PB_Q, NC, Prog>::ZI + PbGeoLoader<Plan, PB_Q, NC, Prog>::stage_doubles(PB_P), vb, ib, gb, zb);
        smem = std::max(smem, vb + ib + gb + zb);
    }
    if (use_smem < 0) {     // query: resident blocks per SM
        int nb = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 128, smem);
        return e == cudaSuccess ? nb : 0;
    }
    const long long blocks = (prm->nthreads + 127) / 128;
    if (blocks <= 0) return 0;
    if (prm->out_smu >= (1LL << 31)) return (int)cudaErrorInvalidValue;
    dim3 grid((unsigned)blocks, (unsigned)ny);
    kern<<<grid, 128, smem, (cudaStream_t)stream>>>(*prm, gp);
    return (int)cudaGetLastError();
#endif
}

template <class Plan, class Prog>
int launch_geo(const PbWalkParams* prm, int use_smem, size_t, void* stream) {
    const PbGeoLineParams* gp = static_cast<const PbGeoLineParams*>(prm->geo_line);
    if (!gp || gp->geo.Ng[0] * ((gp->geo.nc * 3 + 1) & ~1) > PB_GEO_ZMAX) return 1;      // cudaErrorInvalidValue
    return gp->geo.nc == 4 ? launch_geo_nc<Plan, Prog, 4>(prm, *gp, use_smem, stream)
                           : launch_geo_nc<Plan, Prog, 3>(prm, *gp, use_smem, stream);
}

// fused stages 2 + 3 (fused23.cuh); prm->out == nullptr: query, 0 if the configuration can be launched
template <class Form>
int launch_s32(const PbS32Params* prm, void* stream) {
#ifdef PB_EMULATE
    (void)stream;
    if (!prm->out) return 0;
    std::vector<double> T((size_t)prm->G1 * Form::NT * PbS32Cfg<PB_P>::TPAD);
    for (int u = 0; u < prm->mu0_count; ++u)
        for (int b = 0; b < prm->nbatch; ++b)
            for (int y = 0; y < std::max(1, prm->npiece); ++y) {
                std::fill(T.begin(), T.end(), 0.0);
                pb_s32_seq<Form, PB_P, PB_Q>(*prm, prm->mu0_begin + u, b, y, T.data());
            }
    return 0;
#else
    const PbS32Smem<Form, PB_P, PB_Q> lay(prm->v1_rows > 0 ? prm->v1_rows : prm->G1, prm->N1);
    if (lay.total > 227 * 1024) return 1;
    if (!prm->out) return 0;
    auto kern = pb_s32_kernel<Form, PB_P, PB_Q>;
    static unsigned long long configured = 0;       // bit d: done on device d
    int devno = 0;
    cudaGetDevice(&devno);
    if (devno >= 64 || !((configured >> devno) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        if (devno < 64) configured |= 1ull << devno;
    }
    const long long blocks = prm->n_whole + (pb_s32_tasks(*prm) - prm->n_whole) * std::max(1, prm->npiece);
    if (blocks <= 0) return 0;
    kern<<<(unsigned)blocks, (PB_Q * PbS32Split<Form, PB_P, PB_Q>::NH + PbS32Cfg<PB_P>::NCW) * 32, lay.total, (cudaStream_t)stream>>>(*prm);
    return (int)cudaGetLastError();
#endif
}

int launch_walk1(const PbWalk1Params* prm, void* stream) {
#ifdef PB_EMULATE
    (void)stream;
    pb_emu_for(prm->nthreads, [&](long long tid) { pb_walk1_line<PB_P, PB_Q>(*prm, tid); });
    return 0;
#else
    const long long blocks = (prm->nthreads + 127) / 128;
    if (blocks <= 0) return 0;
    pb_walk1_kernel<PB_P, PB_Q><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*prm);
    return (int)cudaGetLastError();
#endif
}

struct Registrar {
    Registrar() {
        pb200_register_walk1(PB_P, PB_Q, &launch_walk1);
#ifndef PB_WALK1_ONLY       // pairs that only occur in linear forms (the one-node axis of boundary integrals)
        pb200_register_walk(PB_PLAN_LANE_BASE + PB_PLAN_COPY, PB_P, PB_Q, &launch_lane<PbPlanCopy>);
        pb200_register_walk(PB_PLAN_LANE_BASE + PB_PLAN_FINAL4, PB_P, PB_Q, &launch_lane<PbPlanFinal4>);
        pb200_register_walk(PB_PLAN_LANE_BASE + PB_PLAN_GEN4, PB_P, PB_Q, &launch_lane<PbPlanGen4>);
        pb200_register_walk(PB_PLAN_GEN4, PB_P, PB_Q, &launch<PbPlanGen4>);
        pb200_register_walk(PB_PLAN_ONE11, PB_P, PB_Q, &launch<PbPlanOne11>);
        pb200_register_walk(PB_PLAN_ONE10, PB_P, PB_Q, &launch<PbPlanOne10>);
        pb200_register_walk(PB_PLAN_PAIRT, PB_P, PB_Q, &launch<PbPlanPairT>);
        pb200_register_walk(PB_PLAN_COPY, PB_P, PB_Q, &launch<PbPlanCopy>);
        pb200_register_walk(PB_PLAN_FINAL4, PB_P, PB_Q, &launch<PbPlanFinal4>);
        pb200_register_walk(PB_PLAN_S1A, PB_P, PB_Q, &launch<PbPlanS1A>);
        pb200_register_walk(PB_PLAN_S1B, PB_P, PB_Q, &launch<PbPlanS1B>);
        pb200_register_walk(PB_PLAN_S2B, PB_P, PB_Q, &launch<PbPlanS2B>);
        pb200_register_walk(PB_PLAN_S1_2D, PB_P, PB_Q, &launch<PbPlanS1_2D>);
#if PB_Q == PB_P + 1 && PB_P <= 3
        pb200_register_s32(2 /* PB200_FORM_STIFFNESS */, PB_P, PB_Q, &launch_s32<PbS32Stiffness>);
        pb200_register_s32(1 /* PB200_FORM_MASS */, PB_P, PB_Q, &launch_s32<PbS32Mass>);
        pb200_register_s32(100 /* PB200_FORM_CUSTOM */, PB_P, PB_Q, &launch_s32<PbS32Generic>);
#endif
        pb200_register_walk(PB_PLAN_S1F, PB_P, PB_Q, &launch_geo<PbPlanS1F, PbProgStiffness<3>>);
        pb200_register_walk(PB_PLAN_S1F_MASS, PB_P, PB_Q, &launch_geo<PbPlanS1FMass, PbProgMass<3>>);
#endif
    }
};
Registrar registrar;

}  // namespace
