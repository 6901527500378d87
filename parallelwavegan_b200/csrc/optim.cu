// Multi-tensor optimizer step with fused global-norm gradient clipping:
//   torch.nn.utils.clip_grad_norm_ (bin/train.py:289-293, 329-333) + Adam (torch.optim.Adam semantics, the
//   HiFi-GAN recipes) / RAdam (optimizers/radam.py:27-99, the Parallel WaveGAN recipes)
// for ALL parameters of a model in three launches (chunk partial norms, ordered reduction, update) instead of
// several tiny launches per parameter.  Deterministic: fixed chunking, fixed reduction order, no atomics.
#include "common.cuh"

namespace pwgb {

struct MtEntry {  // one row of the device-side tensor table (5 x int64)
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};

constexpr int MT_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = l < MT_THREADS / 32 ? sh[l] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(MT_THREADS)
    mt_sqnorm_kernel(const MtEntry* __restrict__ table, const int2* __restrict__ chunks, int chunk_elems, float* __restrict__ partial) {
  __shared__ float sh[MT_THREADS / 32];
  const int2 ch = chunks[blockIdx.x];
  const MtEntry e = table[ch.x];
  const long long lo = (long long)ch.y * chunk_elems;
  const long long hi = lo + chunk_elems < e.n ? lo + chunk_elems : e.n;
  float acc = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += MT_THREADS) {
    const float g = e.g[i];
    acc = fmaf(g, g, acc);
  }
  const float t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = total norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (1 when max_norm <= 0)
__global__ void __launch_bounds__(1024) mt_clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += (double)partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sh[0]);
    out[0] = norm;
    float c = 1.f;
    if (max_norm > 0.f) {
      c = max_norm / (norm + 1e-6f);
      c = c > 1.f ? 1.f : c;
    }
    out[1] = c;
  }
}

// mode 0: Adam   (torch.optim.Adam, amsgrad=False):  g += wd*p; m,v update; p -= c1 * m / (sqrt(v) * c2 + eps)
//                c1 = lr / (1 - beta1^t), c2 = 1 / sqrt(1 - beta2^t)
// mode 1: RAdam rectified   (radam.py:89-91):  m,v update; p -= wd*lr*p; p -= c1 * m / (sqrt(v) + eps), c1 = step_size * lr
// mode 2: RAdam unrectified (radam.py:92-93):  m,v update; p -= wd*lr*p; p -= c1 * m
__global__ void __launch_bounds__(MT_THREADS)
    mt_adam_kernel(const MtEntry* __restrict__ table, const int2* __restrict__ chunks, int chunk_elems, int mode, float lr, float beta1,
                   float beta2, float eps, float wd, float c1, float c2, const float* __restrict__ coef, int write_grad) {
  const int2 ch = chunks[blockIdx.x];
  const MtEntry e = table[ch.x];
  const long long lo = (long long)ch.y * chunk_elems;
  const long long hi = lo + chunk_elems < e.n ? lo + chunk_elems : e.n;
  const float cs = coef ? coef[1] : 1.f;
  for (long long i = lo + threadIdx.x; i < hi; i += MT_THREADS) {
    float g = e.g[i] * cs;
    float p = e.p[i];
    if (write_grad && coef) const_cast<float*>(e.g)[i] = g;
    if (mode == 0 && wd != 0.f) g = fmaf(wd, p, g);
    const float m = beta1 * e.m[i] + (1.f - beta1) * g;
    const float v = beta2 * e.v[i] + (1.f - beta2) * g * g;
    e.m[i] = m;
    e.v[i] = v;
    if (mode == 0) {
      p -= c1 * (m / (sqrtf(v) * c2 + eps));
    } else {
      if (wd != 0.f) p -= wd * lr * p;
      p -= mode == 1 ? c1 * (m / (sqrtf(v) + eps)) : c1 * m;
    }
    e.p[i] = p;
  }
}

}  // namespace pwgb

using namespace pwgb;

extern "C" int pwgb_mt_clip_coef(const void* table, const void* chunks, int n_chunks, int chunk_elems, float max_norm, float* partial,
                                 float* out2, void* stream) {
  PWGB_CHECK_ARG(table && chunks && partial && out2 && n_chunks >= 0 && chunk_elems > 0, "mt_clip_coef: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_chunks > 0) {
    mt_sqnorm_kernel<<<n_chunks, MT_THREADS, 0, st>>>((const MtEntry*)table, (const int2*)chunks, chunk_elems, partial);
    int rc = check_launch("mt_sqnorm_kernel");
    if (rc) return rc;
  }
  mt_clip_coef_kernel<<<1, 1024, 0, st>>>(partial, n_chunks, max_norm, out2);
  return check_launch("mt_clip_coef_kernel");
}

extern "C" int pwgb_mt_adam_step(const void* table, const void* chunks, int n_chunks, int chunk_elems, int mode, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, float c1, float c2, const float* coef2, int write_clipped_grad,
                                 void* stream) {
  PWGB_CHECK_ARG(table && chunks && n_chunks >= 0 && chunk_elems > 0 && mode >= 0 && mode <= 2, "mt_adam_step: bad argument");
  if (n_chunks == 0) return PWGB_OK;
  mt_adam_kernel<<<n_chunks, MT_THREADS, 0, (cudaStream_t)stream>>>((const MtEntry*)table, (const int2*)chunks, chunk_elems, mode, lr, beta1,
                                                                  beta2, eps, weight_decay, c1, c2, coef2, write_clipped_grad);
  return check_launch("mt_adam_kernel");
}
