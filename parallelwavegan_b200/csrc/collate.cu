// Training collater on the GPU (Collater.__call__, bin/train.py:711-798): the random-crop gather
// of a batch from a device-resident corpus in ONE launch.  The crop positions are drawn on the host with the
// reference's own RNG call sequence (np.random.randint per item), so the same seed gives the same batch; the data
// never leaves HBM (a 180 GB part holds LJSpeech-size corpora whole: 24 h of 22.05 kHz fp32 audio = 7.6 GB).
#include "common.cuh"

namespace pwgb {

// y[b, t]      = audio[x_off[b] + t]                         t < T
// c[b, ch, f]  = feats[(c_off[b] + f) * C + ch]              f < F   (rows of the (frames, C) feature matrix, transposed)
__global__ void collate_crop_kernel(const float* __restrict__ audio, const long long* __restrict__ x_off, const float* __restrict__ feats,
                                    const long long* __restrict__ c_off, float* __restrict__ y, float* __restrict__ c, int T, int C,
                                    int F, int y_blocks) {
  const int b = blockIdx.y;
  if ((int)blockIdx.x < y_blocks) {
    const float* src = audio + x_off[b];
    float* dst = y + (long long)b * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += y_blocks * blockDim.x) dst[t] = __ldg(src + t);
    return;
  }
  if (!feats) return;
  __shared__ float tile[32][33];
  const int tiles_c = (C + 31) / 32;
  const int tix = blockIdx.x - y_blocks;
  const int f0 = (tix / tiles_c) * 32, c0 = (tix % tiles_c) * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 256 threads = 32 x 8
  const float* src = feats + c_off[b] * C;
  for (int i = ly; i < 32; i += 8) {
    const int f = f0 + i, ch = c0 + lx;
    tile[i][lx] = (f < F && ch < C) ? __ldg(src + (long long)f * C + ch) : 0.f;
  }
  __syncthreads();
  for (int i = ly; i < 32; i += 8) {
    const int ch = c0 + i, f = f0 + lx;
    if (ch < C && f < F) c[((long long)b * C + ch) * F + f] = tile[lx][i];
  }
}

}  // namespace pwgb

using namespace pwgb;

extern "C" int pwgb_collate_crop(const float* audio, const long long* x_offsets, const float* feats, const long long* c_offsets, float* y,
                                 float* c, int batch, int t, int channels, int frames, void* stream) {
  PWGB_CHECK_ARG(audio && x_offsets && y && batch >= 0 && t > 0, "collate_crop: bad argument");
  PWGB_CHECK_ARG(!feats || (c_offsets && c && channels > 0 && frames > 0), "collate_crop: bad feature arguments");
  if (batch == 0) return PWGB_OK;
  const int y_blocks = ceil_div(t, 256 * 8) < 1 ? 1 : ceil_div(t, 256 * 8);
  const int c_blocks = feats ? ceil_div(frames, 32) * ceil_div(channels, 32) : 0;
  dim3 grid((unsigned)(y_blocks + c_blocks), (unsigned)batch);
  collate_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(audio, x_offsets, feats, c_offsets, y, c, t, channels, frames, y_blocks);
  return check_launch("collate_crop_kernel");
}
