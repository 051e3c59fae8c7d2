// Collaborative item indexing, the O(sum_u len_u^2) part on the GPU.
//   ref src/src_t5/utils/indexing.py:149-180  : item co-occurrence ("adjacency") matrix over the training prefix of every
//                                               user sequence — a Python double loop over itertools.combinations
//   ref utils/indexing.py:220-231              : the same matrix restricted to the items of one cluster (Python O(n^2) loop)
// The counts are small integers, so fp32 / fp64 atomics give the reference's matrix bit for bit; the spectral clustering
// that consumes it stays scikit-learn on the host (openp5_b200/indexing.py), as in the reference.
#include "common.cuh"

namespace p5 {
extern int g_launches;

template <typename T>
__global__ void __launch_bounds__(256)
cooccurrence_kernel(const int* __restrict__ items, const long long* __restrict__ offs, int n_items, T* __restrict__ adj) {
    pdl_wait();
    pdl_launch_dependents();
    const int u = blockIdx.x;
    const long long o = offs[u];
    const int n = (int)(offs[u + 1] - o);
    const long long pairs = (long long)n * (n - 1) / 2;
    for (long long p = threadIdx.x; p < pairs; p += blockDim.x) {
        // p -> (i, j), i < j, in combinations() order (the order does not matter for the sums)
        int i = (int)((2.0 * n - 1.0 - sqrt((2.0 * n - 1.0) * (2.0 * n - 1.0) - 8.0 * (double)p)) * 0.5);
        long long base = (long long)i * (2 * n - i - 1) / 2;
        while (base > p) { --i; base = (long long)i * (2 * n - i - 1) / 2; }
        while (base + (n - i - 1) <= p) { base += n - i - 1; ++i; }
        const int j = i + 1 + (int)(p - base);
        const int a = items[o + i], b = items[o + j];
        if (a < 0 || b < 0 || a >= n_items || b >= n_items) continue;
        atomicAdd(adj + (long long)a * n_items + b, (T)1);
        atomicAdd(adj + (long long)b * n_items + a, (T)1);
    }
}

template <typename T>
__global__ void submatrix_kernel(const T* __restrict__ adj, int n_items, const int* __restrict__ idx, int m, T* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const long long total = (long long)m * m;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int i = (int)(e / m), j = (int)(e - (long long)i * m);
        out[e] = (i == j) ? (T)0 : adj[(long long)idx[i] * n_items + idx[j]];     // the reference fills i != j only
    }
}

void cooccurrence(const int* items, const long long* offs, int n_users, int n_items, int f64, void* adj, cudaStream_t st) {
    P5_CHECK(n_users >= 0 && n_items >= 1, "cooccurrence: empty problem");
    const size_t bytes = (size_t)n_items * n_items * (f64 ? 8 : 4);
    P5_CUDA(cudaMemsetAsync(adj, 0, bytes, st));
    if (n_users == 0) return;
    if (f64) launch_k(cooccurrence_kernel<double>, (unsigned)n_users, 256, 0, st, items, offs, n_items, (double*)adj);
    else launch_k(cooccurrence_kernel<float>, (unsigned)n_users, 256, 0, st, items, offs, n_items, (float*)adj);
    P5_CUDA(cudaGetLastError());
    ++g_launches;
}
void submatrix(const void* adj, int n_items, int f64, const int* idx, int m, void* out, cudaStream_t st) {
    if (m <= 0) return;
    const unsigned grid = (unsigned)std::min<long long>(((long long)m * m + 255) / 256, 148 * 16);
    if (f64) launch_k(submatrix_kernel<double>, grid, 256, 0, st, (const double*)adj, n_items, idx, m, (double*)out);
    else launch_k(submatrix_kernel<float>, grid, 256, 0, st, (const float*)adj, n_items, idx, m, (float*)out);
    P5_CUDA(cudaGetLastError());
    ++g_launches;
}

}  // namespace p5
