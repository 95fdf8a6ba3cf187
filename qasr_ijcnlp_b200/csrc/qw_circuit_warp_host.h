// Host prototypes of the thread-group cooperative circuit path (qw_circuit_warp.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace qw {
namespace wc {
int wcirc_grid(long long W, int q);
size_t wcirc_workspace_bytes(long long W, int q, int Lq, int elem_size);
template <typename T>
int wcirc_forward(const T* pre, const T* qwts, T* out, long long W, int q, int Lq, int emb, cudaStream_t st);
template <typename T>
int wcirc_backward(const T* pre, const T* qwts, const T* gout, T* gpre, T* gqw, void* ws, long long W, int q, int Lq, int emb,
                   cudaStream_t st);
}  // namespace wc
}  // namespace qw
