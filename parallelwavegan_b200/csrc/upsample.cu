// PWG conditioning upsampler stage: nearest repeat x s + (2s+1)-tap FIR (layers/upsample.py:112-128),
// evaluated poly-phase: output o only touches input frames (o-s)/s .. (o+s)/s, no stretched tensor.
#include "common.cuh"

namespace pwgb {

// Output o = j*s + ph only touches the input frames j-1, j, j+1: y[o] = A[ph] x[j-1] + B[ph] x[j] + C[ph] x[j+1] with the
// taps pre-summed per phase (A: taps that land in frame j-1, ...).  One thread per input frame writes its s outputs
// (one 16-byte store per 4): 3 FMAs per output instead of 2s+1 taps with an integer division each (the tap loop ran
// at 1/12 of the HBM rate: 13 % of the Parallel WaveGAN forward).
__global__ void upsample_fir_kernel(int rows, int rows_per_batch, int t_in, int s, const float* __restrict__ x,
                                    const float* __restrict__ fir, float* __restrict__ y, long long ybs) {
  extern __shared__ float coef[];  // [3][s]
  for (int ph = threadIdx.x; ph < s; ph += blockDim.x) {
    float a = 0.f, bsum = 0.f, c = 0.f;
    for (int k = 0; k <= 2 * s; ++k) {
      const int d = ph + k - s;
      const float f = fir[k];
      if (d < 0)
        a += f;
      else if (d < s)
        bsum += f;
      else
        c += f;
    }
    coef[ph] = a;
    coef[s + ph] = bsum;
    coef[2 * s + ph] = c;
  }
  __syncthreads();
  const int r = blockIdx.y;
  const int t_out = t_in * s;
  const float* xr = x + (long long)r * t_in;
  float* yr = y + (long long)(r / rows_per_batch) * ybs + (long long)(r % rows_per_batch) * t_out;
  const bool vec = (s % 4 == 0) && ((reinterpret_cast<uintptr_t>(yr) & 15) == 0);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < t_in; j += gridDim.x * blockDim.x) {
    const float xm = j > 0 ? __ldg(xr + j - 1) : 0.f, x0 = __ldg(xr + j), xp = j + 1 < t_in ? __ldg(xr + j + 1) : 0.f;
    float* dst = yr + (long long)j * s;
    if (vec) {
      for (int ph = 0; ph < s; ph += 4) {
        float4 v;
        v.x = fmaf(coef[2 * s + ph], xp, fmaf(coef[s + ph], x0, coef[ph] * xm));
        v.y = fmaf(coef[2 * s + ph + 1], xp, fmaf(coef[s + ph + 1], x0, coef[ph + 1] * xm));
        v.z = fmaf(coef[2 * s + ph + 2], xp, fmaf(coef[s + ph + 2], x0, coef[ph + 2] * xm));
        v.w = fmaf(coef[2 * s + ph + 3], xp, fmaf(coef[s + ph + 3], x0, coef[ph + 3] * xm));
        *reinterpret_cast<float4*>(dst + ph) = v;
      }
    } else {
      for (int ph = 0; ph < s; ++ph) dst[ph] = fmaf(coef[2 * s + ph], xp, fmaf(coef[s + ph], x0, coef[ph] * xm));
    }
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
  dim3 grid(ceil_div(t_in, 256) < 64 ? ceil_div(t_in, 256) : 64, rows);
  upsample_fir_kernel<<<grid, 256, 3 * scale * sizeof(float), (cudaStream_t)stream>>>(
      rows, rows_per_batch, t_in, scale, x, fir, y, y_batch_stride ? y_batch_stride : (long long)rows_per_batch * t_out);
  return check_launch("upsample_fir_kernel");
}
