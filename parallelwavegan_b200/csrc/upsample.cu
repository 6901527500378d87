// PWG conditioning upsampler stage: nearest repeat x s + (2s+1)-tap FIR (layers/upsample.py:112-128),
// evaluated poly-phase: output o only touches input frames (o-s)/s .. (o+s)/s, no stretched tensor.
#include "common.cuh"

namespace pwgb {

__global__ void upsample_fir_kernel(int rows, int rows_per_batch, int t_in, int s, const float* __restrict__ x,
                                    const float* __restrict__ fir, float* __restrict__ y, long long ybs) {
  extern __shared__ float f[];
  for (int i = threadIdx.x; i < 2 * s + 1; i += blockDim.x) f[i] = fir[i];
  __syncthreads();
  const int r = blockIdx.y;
  const int t_out = t_in * s;
  const float* xr = x + (long long)r * t_in;
  float* yr = y + (long long)(r / rows_per_batch) * ybs + (long long)(r % rows_per_batch) * t_out;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < t_out; o += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k <= 2 * s; ++k) {
      const int i = o + k - s;
      if (i >= 0 && i < t_out) acc = fmaf(f[k], __ldg(xr + i / s), acc);
    }
    yr[o] = acc;
  }
}

}  // namespace pwgb

using namespace pwgb;

extern "C" int pwgb_upsample_fir_forward(int rows, int rows_per_batch, int t_in, int scale, const float* x,
                                         const float* fir, float* y, long long y_batch_stride, void* stream) {
  PWGB_CHECK_ARG(x && fir && y, "upsample_fir: null argument");
  PWGB_CHECK_ARG(rows >= 0 && rows_per_batch > 0 && t_in > 0 && scale > 0 && rows % rows_per_batch == 0,
                 "upsample_fir: bad sizes");
  PWGB_UNSUPPORTED_IF(rows > 65535, "upsample_fir: too many rows");
  if (rows == 0) return PWGB_OK;
  const int t_out = t_in * scale;
  dim3 grid(ceil_div(t_out, 256) < 64 ? ceil_div(t_out, 256) : 64, rows);
  upsample_fir_kernel<<<grid, 256, (2 * scale + 1) * sizeof(float), (cudaStream_t)stream>>>(
      rows, rows_per_batch, t_in, scale, x, fir, y, y_batch_stride ? y_batch_stride : (long long)rows_per_batch * t_out);
  return check_launch("upsample_fir_kernel");
}
