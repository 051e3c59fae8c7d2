// extern "C" boundary of libp5b200.so — see include/p5_b200.h for the contract of every entry point.
#include "common.cuh"
#include "engine.h"
#include "../../include/p5_b200.h"
#include <string>

using namespace p5;

static thread_local std::string g_last_error;
namespace p5 { int g_launches = 0; }

#define P5_API_BEGIN try {
#define P5_API_END                                        \
    }                                                     \
    catch (const P5Error& e) {                            \
        g_last_error = e.what();                          \
        return e.code ? e.code : 1;                       \
    }                                                     \
    catch (const std::exception& e) {                     \
        g_last_error = std::string("exception: ") + e.what(); \
        return 99;                                        \
    }                                                     \
    return 0;

extern "C" {

const char* p5_last_error(void) { return g_last_error.c_str(); }
int p5_version(void) { return 100; }
int p5_launch_count(void) { return p5::g_launches + gemm_tc_launch_count(); }

int p5_op_gemm(const P5GemmDesc* d, void* cuda_stream) {
    P5_API_BEGIN
    GemmProblem p;
    p.M = d->M; p.N = d->N; p.K = d->K; p.nb1 = d->nb1 > 0 ? d->nb1 : 1; p.nb2 = d->nb2 > 0 ? d->nb2 : 1;
    p.A.ptr = d->A; p.A.dtype = d->a_dtype; p.A.major = d->a_major; p.A.ld = d->lda; p.A.bs1 = d->a_bs1; p.A.bs2 = d->a_bs2;
    p.B.ptr = d->B; p.B.dtype = d->b_dtype; p.B.major = d->b_major; p.B.ld = d->ldb; p.B.bs1 = d->b_bs1; p.B.bs2 = d->b_bs2;
    p.epi.C = d->C; p.epi.c_dtype = d->c_dtype; p.epi.ldc = d->ldc; p.epi.cs1 = d->c_bs1; p.epi.cs2 = d->c_bs2;
    p.epi.alpha = d->alpha; p.epi.flags = d->flags; p.epi.aux = d->aux; p.epi.aux_dtype = d->aux_dtype;
    p.epi.resid = d->resid; p.epi.seed = d->seed; p.epi.site = d->site;
    p.epi.drop_thr = drop_threshold(d->drop_p);
    p.epi.inv_keep = d->drop_p < 1.f ? 1.f / (1.f - d->drop_p) : 0.f;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d->backend == 0) {
        gemm_simt(p, st);
    } else if (d->backend == 1) {
        gemm_tc_force_block_n(d->force_block_n);
        gemm_tc(p, st);
        gemm_tc_force_block_n(0);
    } else {
        gemm_auto(p, st, true);
    }
    P5_API_END
}

}  // extern "C"
