// 3x3 / stride 2 / pad 1 max pooling on channels-last activations: the ResNet stem's nn.MaxPool2d between bn1+relu and
// layer1 (reference resnet/models/resnet_mrla_light.py `self.maxpool`, resnet_mrla_base.py likewise).  It sits on the
// whole-step path of bench.py only: PyTorch's max_pool_backward_nhwc alone took 1.7 ms of a
// 35 ms resnet50_mrlal step (0.5 TB/s).  Forward keeps the winning tap (0..8) per element in one byte, so backward is
// a gather over the <= 4 windows that cover an input pixel: no atomics, deterministic, R(dy, idx) W(dx) once.
// Tie / NaN rules follow at::max_pool2d: taps are scanned row-major, a later tap wins only if strictly greater or NaN.
#include "../../include/mrla_b200.h"
#include "common.cuh"

namespace mrla {

extern thread_local int g_launch_count;

struct PoolShape { int B, C, H, W, OH, OW, CV; };

// Thread = (output row, 8-channel vector of one output pixel): blockIdx.x walks the B*OH output rows, blockIdx.y the
// chunks of OW*CV vectors inside a row, so the index arithmetic is one 32-bit division per thread.
template <typename T>
__global__ void __launch_bounds__(256) k_maxpool3x3s2_fwd(const T* __restrict__ x, T* __restrict__ y,
                                                          uint8_t* __restrict__ idx, PoolShape s) {
  const int e = blockIdx.y * blockDim.x + threadIdx.x;   // (ow, cv) inside the row
  if (e >= s.OW * s.CV) return;
  const int ow = e / s.CV, cv = e - ow * s.CV;
  const int b = blockIdx.x / s.OH, oh = blockIdx.x - b * s.OH;
  const int h0 = 2 * oh - 1, w0 = 2 * ow - 1;
  const T* xb = x + (int64_t)b * s.H * s.W * s.C + cv * 8;
  float m[8];
  int mi[8];
  bool first = true;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int h = h0 + kh;
    if (h < 0 || h >= s.H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int w = w0 + kw;
      if (w < 0 || w >= s.W) continue;
      float a[8];
      ld_vec<T, 8>(xb + ((int64_t)h * s.W + w) * s.C, a);
      if (first) {
        // at::max_pool2d starts at -inf with the index of the first in-bounds tap
#pragma unroll
        for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; mi[i] = kh * 3 + kw; }
        first = false;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (a[i] > m[i] || a[i] != a[i]) { m[i] = a[i]; mi[i] = kh * 3 + kw; }
    }
  }
  const int64_t o = ((int64_t)blockIdx.x * s.OW + ow) * s.C + cv * 8;
  st_vec<T, 8>(y + o, m);
  Pack<uint8_t, 8> pk;
#pragma unroll
  for (int i = 0; i < 8; ++i) pk.v[i] = (uint8_t)mi[i];
  *reinterpret_cast<Pack<uint8_t, 8>*>(idx + o) = pk;
}

// Thread = (input row, 8-channel vector of one input pixel): blockIdx.x walks the B*H input rows.
template <typename T>
__global__ void __launch_bounds__(256) k_maxpool3x3s2_bwd(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                          T* __restrict__ dx, PoolShape s) {
  const int e = blockIdx.y * blockDim.x + threadIdx.x;   // (w, cv) inside the row
  if (e >= s.W * s.CV) return;
  const int w = e / s.CV, cv = e - w * s.CV;
  const int b = blockIdx.x / s.H, h = blockIdx.x - b * s.H;
  // windows covering (h, w): 2*o - 1 <= h <= 2*o + 1
  const int oh_lo = h / 2, oh_hi = min((h + 1) / 2, s.OH - 1);
  const int ow_lo = w / 2, ow_hi = min((w + 1) / 2, s.OW - 1);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int oh = oh_lo; oh <= oh_hi; ++oh) {
    const int kh = h - (2 * oh - 1);
    for (int ow = ow_lo; ow <= ow_hi; ++ow) {
      const int k = kh * 3 + (w - (2 * ow - 1));
      const int64_t o = (((int64_t)b * s.OH + oh) * s.OW + ow) * s.C + cv * 8;
      const Pack<uint8_t, 8> pk = *reinterpret_cast<const Pack<uint8_t, 8>*>(idx + o);
      float g[8];
      ld_vec<T, 8>(dy + o, g);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += (pk.v[i] == k) ? g[i] : 0.f;
    }
  }
  st_vec<T, 8>(dx + ((int64_t)blockIdx.x * s.W + w) * s.C + cv * 8, acc);
}

static int pool_shape(int B, int C, int H, int W, int dtype, PoolShape* s) {
  if (B < 1 || C < 8 || H < 1 || W < 1) return MRLA_ERR_SHAPE;
  if (C % 8) return MRLA_ERR_ALIGN;
  if (dtype < MRLA_F32 || dtype > MRLA_F16) return MRLA_ERR_UNSUPPORTED;
  s->B = B; s->C = C; s->H = H; s->W = W;
  s->OH = (H - 1) / 2 + 1;
  s->OW = (W - 1) / 2 + 1;
  s->CV = C / 8;
  return MRLA_OK;
}

// grid: x = rows (B*OH or B*H), y = 256-thread chunks of one row's (column, channel-vector) pairs
static inline bool pool_grid(int64_t rows, int64_t per_row, dim3* g) {
  const int64_t gy = (per_row + 255) / 256;
  if (rows < 1 || rows > 0x7fffffffLL || gy > 65535) return false;
  *g = dim3((unsigned)rows, (unsigned)gy, 1);
  return true;
}

}  // namespace mrla

using namespace mrla;

extern "C" {

int mrla_maxpool3x3s2_forward(const void* x, void* y, unsigned char* idx, int B, int C, int H, int W, int dtype,
                              void* stream) {
  NvtxRange nvtx_("mrla_maxpool3x3s2_forward");
  g_launch_count = 0;
  PoolShape s;
  int rc = pool_shape(B, C, H, W, dtype, &s);
  if (rc) return rc;
  if (!x || !y || !idx) return MRLA_ERR_NULL;
  const uintptr_t al = dtype == MRLA_F32 ? 32 : 16;
  if (((uintptr_t)x % al) || ((uintptr_t)y % al) || ((uintptr_t)idx % 8)) return MRLA_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid;
  if (!pool_grid((int64_t)B * s.OH, (int64_t)s.OW * s.CV, &grid)) return MRLA_ERR_SHAPE;
  if (dtype == MRLA_F32)
    k_maxpool3x3s2_fwd<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), static_cast<float*>(y), idx, s);
  else if (dtype == MRLA_BF16)
    k_maxpool3x3s2_fwd<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, s);
  else
    k_maxpool3x3s2_fwd<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(y), idx, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++g_launch_count;
  return MRLA_OK;
}

int mrla_maxpool3x3s2_backward(const void* dy, const unsigned char* idx, void* dx, int B, int C, int H, int W, int dtype,
                               void* stream) {
  NvtxRange nvtx_("mrla_maxpool3x3s2_backward");
  g_launch_count = 0;
  PoolShape s;
  int rc = pool_shape(B, C, H, W, dtype, &s);
  if (rc) return rc;
  if (!dy || !dx || !idx) return MRLA_ERR_NULL;
  const uintptr_t al = dtype == MRLA_F32 ? 32 : 16;
  if (((uintptr_t)dy % al) || ((uintptr_t)dx % al) || ((uintptr_t)idx % 8)) return MRLA_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid;
  if (!pool_grid((int64_t)B * H, (int64_t)W * s.CV, &grid)) return MRLA_ERR_SHAPE;
  if (dtype == MRLA_F32)
    k_maxpool3x3s2_bwd<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dy), idx, static_cast<float*>(dx), s);
  else if (dtype == MRLA_BF16)
    k_maxpool3x3s2_bwd<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), idx, static_cast<__nv_bfloat16*>(dx), s);
  else
    k_maxpool3x3s2_bwd<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(dy), idx, static_cast<__half*>(dx), s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++g_launch_count;
  return MRLA_OK;
}

}  // extern "C"
