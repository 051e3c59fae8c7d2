// fp32 SIMT GEMM: the exact-arithmetic ("parity") form of every contraction and the fallback for shapes the
// tcgen05 kernel does not take (unaligned leading dimensions, fp32 operands).  Same GemmProblem contract and
// the same fused epilogue as gemm_tc.cu.  64x64 tiles, 16x16 threads, 4x4 micro-tiles, fp32 accumulate.
#include "common.cuh"

namespace p5 {

static constexpr int ST = 64;   // tile M, N
static constexpr int SK = 16;   // tile K

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, int nb1, const TA* __restrict__ A, int64_t a_rs, int64_t a_ks, int64_t a_bs1,
                 int64_t a_bs2, const TB* __restrict__ B, int64_t b_rs, int64_t b_ks, int64_t b_bs1, int64_t b_bs2,
                 GemmEpilogue epi) {
    pdl_wait();   // programmatic dependent launch: everything above the wait overlaps the previous kernel
    pdl_launch_dependents();
    __shared__ float sA[SK][ST + 1];
    __shared__ float sB[SK][ST + 1];
    const int b = blockIdx.z;
    const int b1 = b % nb1, b2 = b / nb1;
    A += b1 * a_bs1 + b2 * a_bs2;
    B += b1 * b_bs1 + b2 * b_bs2;
    const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SK) {
        // 64x16 elements per operand, 256 threads -> 4 each.  Pick the thread->element map that is coalesced
        // for the operand's contiguous index.
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = threadIdx.x + i * 256;
            int r, k;
            if (a_ks == 1) { k = e & 15; r = e >> 4; } else { r = e & 63; k = e >> 6; }
            int gm = m0 + r, gk = k0 + k;
            sA[k][r] = (gm < M && gk < K) ? to_f32(A[gm * a_rs + gk * a_ks]) : 0.f;
            if (b_ks == 1) { k = e & 15; r = e >> 4; } else { r = e & 63; k = e >> 6; }
            int gn = n0 + r;
            gk = k0 + k;
            sB[k][r] = (gn < N && gk < K) ? to_f32(B[gn * b_rs + gk * b_ks]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SK; ++k) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int64_t boff = (int64_t)b1 * epi.cs1 + (int64_t)b2 * epi.cs2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            int64_t idx = boff + (int64_t)gm * epi.ldc + gn;
            float v = epilogue_apply(epi, acc[i][j], idx);
            if (epi.flags & EPI_ATOMIC) atomicAdd((float*)epi.C + idx, v);
            else st_from_f32(epi.C, epi.c_dtype, idx, v);
        }
    }
}

template <typename TA, typename TB>
static void launch_simt(const GemmProblem& p, cudaStream_t stream) {
    dim3 grid((unsigned)cdiv(p.N, ST), (unsigned)cdiv(p.M, ST), (unsigned)(p.nb1 * p.nb2));
    int64_t a_rs = p.A.major == MAJOR_K ? p.A.ld : 1, a_ks = p.A.major == MAJOR_K ? 1 : p.A.ld;
    int64_t b_rs = p.B.major == MAJOR_K ? p.B.ld : 1, b_ks = p.B.major == MAJOR_K ? 1 : p.B.ld;
    launch_k(gemm_simt_kernel<TA, TB>, grid, 256, 0, stream, p.M, p.N, p.K, p.nb1, (const TA*)p.A.ptr, a_rs, a_ks, p.A.bcast1 ? 0 : p.A.bs1,
                                                        p.A.bcast2 ? 0 : p.A.bs2, (const TB*)p.B.ptr, b_rs, b_ks, p.B.bcast1 ? 0 : p.B.bs1, p.B.bcast2 ? 0 : p.B.bs2,
                                                        p.epi);
    P5_CUDA(cudaGetLastError());
}

void gemm_simt(const GemmProblem& p, cudaStream_t stream) {
    P5_CHECK(p.M > 0 && p.N > 0 && p.K > 0, "gemm_simt: empty problem");
    P5_CHECK(cdiv(p.M, ST) <= 65535 && p.nb1 * p.nb2 <= 65535, "gemm_simt: grid too large");
    if (p.A.dtype == DT_F32 && p.B.dtype == DT_F32) launch_simt<float, float>(p, stream);
    else if (p.A.dtype == DT_BF16 && p.B.dtype == DT_BF16) launch_simt<bf16, bf16>(p, stream);
    else if (p.A.dtype == DT_F32 && p.B.dtype == DT_BF16) launch_simt<float, bf16>(p, stream);
    else launch_simt<bf16, float>(p, stream);
}

}  // namespace p5

namespace p5 {
extern int g_launches;
void gemm_auto(const GemmProblem& p, cudaStream_t stream, bool allow_mn_major) {
    if (gemm_tc_supported(p, allow_mn_major)) {
        gemm_tc(p, stream);
    } else {
        gemm_simt(p, stream);
        ++g_launches;
    }
}
}  // namespace p5
