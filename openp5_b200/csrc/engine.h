// engine.h — host-side declarations shared between the C-ABI and the engine implementation.
#pragma once
#include "common.cuh"

namespace p5 {
extern int g_launches;
void gemm_tc_force_block_n(int bn);
// tcgen05 when the problem qualifies (bf16 operands, aligned), SIMT otherwise
void gemm_auto(const GemmProblem& p, cudaStream_t stream, bool allow_mn_major);
}  // namespace p5
