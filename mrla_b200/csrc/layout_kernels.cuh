// NCHW -> NHWC repack used when a dense NCHW activation is promoted to the TMA (channels_last) kernels.
// Classic shared-memory tile transpose of the [C, HW] matrix of every sample: coalesced reads along HW,
// coalesced 2-element stores along C.
#pragma once
#include "common.cuh"

namespace mrla {

template <typename T>
__global__ void __launch_bounds__(256) k_nchw_to_nhwc(const T* __restrict__ src, T* __restrict__ dst, int C, int HW,
                                                      int64_t bs_src, int64_t bs_dst) {
  constexpr int TILE = 64;
  __shared__ T tile[TILE][TILE + 2];
  const int b = blockIdx.z;
  const int hw0 = blockIdx.x * TILE, c0 = blockIdx.y * TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* s = src + (int64_t)b * bs_src;
  T* d = dst + (int64_t)b * bs_dst;
#pragma unroll
  for (int i = 0; i < TILE / 8; ++i) {
    const int c = c0 + warp + i * 8;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int hw = hw0 + lane + 32 * k;
      if (c < C && hw < HW) tile[warp + i * 8][lane + 32 * k] = s[(int64_t)c * HW + hw];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < TILE / 8; ++i) {
    const int hw = hw0 + warp + i * 8;
    const int c = c0 + 2 * lane;
    if (hw < HW && c + 1 < C) {
      Pack<T, 2> pk;
      pk.v[0] = tile[2 * lane][warp + i * 8];
      pk.v[1] = tile[2 * lane + 1][warp + i * 8];
      *reinterpret_cast<Pack<T, 2>*>(d + (int64_t)hw * C + c) = pk;
    } else if (hw < HW && c < C) {
      d[(int64_t)hw * C + c] = tile[2 * lane][warp + i * 8];
    }
  }
}

// x = relu(z + id): the residual add + ReLU in front of the MRLA tail (resnet_mrla_light.py:113-114), one
// vectorised pass (R2 W1) instead of the reference's in-place add (R2 W1) followed by relu_ (R1 W1).
template <typename T>
__global__ void __launch_bounds__(256) k_add_relu(const T* __restrict__ z, const T* __restrict__ idt, T* __restrict__ x,
                                                  int64_t n_vec) {
  constexpr int V = 16 / sizeof(T);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    float a[V], b[V];
    ld_vec<T, V>(z + i * V, a);
    ld_vec<T, V>(idt + i * V, b);
#pragma unroll
    for (int k = 0; k < V; ++k) a[k] = fmaxf(a[k] + b[k], 0.f);
    st_vec<T, V>(x + i * V, a);
  }
}

}  // namespace mrla
