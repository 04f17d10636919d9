// fir.cuh — FirEngine: device-resident streaming / stateless complex FIR over [channels][samples].
// Stands in for ComplexFIRFilter (MS/Models/FIRFilter.cs:8-232); used directly by the qpsk_fir_*
// entry points and as the matched filter inside the demodulator chain.
#pragma once
#include "common.cuh"

namespace qpsk {

struct FirEngine {
  int n_taps = 0;          // complex taps N
  bool real_taps = false;  // every imaginary part is +-0 -> 1 FFMA2 per tap per sample
  int channels = 1;
  int device = 0;              // ordinal the state and streams live on (fixed at init)
  int mode = QPSK_FIR_FAST;
  std::vector<float> taps_iq;  // h[j], interleaved, as given
  int HL = 0;                  // history length kept per channel (even, >= N-1)
  DevBuf<float2> hist[2];      // ping-pong: kernel reads hist[cur], writes hist[cur^1]
  int cur = 0;
  DevBuf<float> d_taps;        // [2][N] planar h (generic kernel)
  cudaStream_t stream = nullptr;  // owned
  const char* last_kernel = "";  // the kernel the last filter call launched (measurement: bench.py names it in `roofline`)

  ~FirEngine();
  int init(const float* taps_iq_in, int n_floats, int channels_in);
  int reset(cudaStream_t s);
  // streaming Filter (:80-91): history consumed and advanced
  int filter_dev(const float2* x, float2* y, int64_t L, int64_t ldx, int64_t ldy, cudaStream_t s);
  // stateless fftFilter alignment (:96-141)
  int fft_filter_dev(const float2* x, float2* y, int64_t L, int64_t ldx, int64_t ldy, cudaStream_t s);
  // decimate-by-`dec` streaming filter: the Filter() output kept at stream indices 0, dec, 2*dec, ... (counted across
  // calls: `dec_skip` input samples are still to pass before the next kept one).  *n_out = outputs per channel.
  int decimate_dev(const float2* x, int64_t L, int64_t ldx, int dec, float2* y, int64_t cap, int64_t ldy, int64_t* n_out,
                   cudaStream_t s);
  int64_t dec_skip = 0;
  int dec_last = 0;            // decimation factor of the previous call (a change restarts the phase)
  DevBuf<float2> dec_tmp;      // full-rate scratch of the fallback path
  int get_state(float* hist_iq, int64_t cap_floats);
  int set_state(const float* hist_iq, int64_t n_floats);

 private:
  int run(const float2* x, float2* y, int64_t L, int64_t ldx, int64_t ldy, bool stateless, cudaStream_t s);
};

}  // namespace qpsk
