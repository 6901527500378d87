// C-ABI plumbing: error strings, launch counter, argument validation and dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pwgb {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int conv1d_forward_simt(const pwgb_conv1d_desc* d, const float* x, const float* w, const float* bias,
                        const float* residual, float* y, cudaStream_t st);

}  // namespace pwgb

using namespace pwgb;

extern "C" const char* pwgb_last_error(void) { return g_err; }
extern "C" int pwgb_version(void) { return 100; }
extern "C" int pwgb_compiled_arch(void) { return 100; }
extern "C" long long pwgb_launch_count(void) { return g_launches; }
extern "C" void pwgb_reset_launch_count(void) { g_launches = 0; }

extern "C" int pwgb_conv1d_forward(const pwgb_conv1d_desc* d, const float* x, const float* w, const float* bias,
                                   const float* residual, float* y, void* stream) {
  PWGB_CHECK_ARG(d && x && w && y, "conv1d: null argument");
  PWGB_CHECK_ARG(d->batch >= 0 && d->cin > 0 && d->cout > 0 && d->t_in > 0 && d->t_out >= 0, "conv1d: bad sizes");
  PWGB_CHECK_ARG(d->kernel > 0 && d->stride > 0 && d->dilation > 0 && d->groups > 0, "conv1d: bad kernel geometry");
  PWGB_CHECK_ARG(d->cin % d->groups == 0 && d->cout % d->groups == 0, "conv1d: channels not divisible by groups");
  PWGB_CHECK_ARG(d->pad_mode >= 0 && d->pad_mode <= 2, "conv1d: bad pad_mode");
  const int P = d->period < 1 ? 1 : d->period;
  PWGB_CHECK_ARG(P == 1 || d->pad_mode == PWGB_PAD_ZERO, "conv1d: period > 1 needs zero padding");
  PWGB_CHECK_ARG(d->pad_mode != PWGB_PAD_REFLECT || d->pad_left < d->t_in, "conv1d: reflect pad >= length");
  // the last output row must not read past the right padding implied by the caller
  const long long need = (long long)(d->t_out - 1) * d->stride + (long long)(d->kernel - 1) * d->dilation + 1 - d->pad_left;
  PWGB_CHECK_ARG(d->t_out == 0 || need <= (long long)d->t_in + (d->pad_mode == PWGB_PAD_ZERO ? (1LL << 30) : d->t_in - 1),
                 "conv1d: t_out too large for t_in");
  PWGB_CHECK_ARG(d->shuffle <= 1 || (d->cout % d->shuffle == 0 && !residual), "conv1d: bad shuffle");
  return conv1d_forward_simt(d, x, w, bias, residual, y, (cudaStream_t)stream);
}
