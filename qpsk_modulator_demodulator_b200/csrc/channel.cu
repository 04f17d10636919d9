// channel.cu — K7 synthetic impairments (two unstable LOs + AWGN + static multipath), K6 per-channel
// BER counters, and the payload generator.  Compiled with --fmad=false: the arithmetic restates the
// reference's C# expressions operation by operation.
//
//   NCO            TB/Simulated/LocalOscilator.cs:5-194  (ppm static error + bounded random-walk drift
//                  every 1 ms, NextSample() = e^{j phase})
//   NoiseGenerator TB/HelperModels.cs:17-45              (Box-Muller, dBFS rms, clamp to +-1)
//   mixing         TB/Simulated/testAtDataLevel.cs:39-42 (mode 0), testFullDemodChain.cs:73 (mode 1)
//   multipath      not in the reference (README.md:2 names it, no code implements it); defined in
//                  DESIGN.md "impairments": y[n] = sum_k g[k] * x[n - d[k]] in fp32, zero initial state
// System.Random is replaced by the counter RNG (common.cuh), streams 4c / 4c+1 / 4c+2 / 4c+3 of
// channel c for the tx NCO, rx NCO, noise and payload.
#include <math.h>

#include "common.cuh"

namespace qpsk {

constexpr int kMaxPaths = 4;
constexpr int kPathHist = 64;   // longest multipath delay + 1, in samples

struct NcoState {
  double phase, static_ppm, drift_ppm, total_ppm, cur_freq;
  unsigned long long counter;
  int drift_counter;
  int pad;
};

struct NcoParams {
  double base_freq, fs, max_ppm, phase0;
  int drift_interval;
};

struct ChanArgs {
  NcoParams tx, rx;
  float noise_rms;        // linear; <= 0 disables the noise term
  int mode;
  int n_paths;
  float gain[2 * kMaxPaths];
  int delay[kMaxPaths];
  unsigned long long seed;
  int first_channel;
  int C;
};

__device__ __forceinline__ void nco_update_freq(const NcoParams& P, NcoState& S) {   // :181-186
  S.cur_freq = P.base_freq * (1.0 + S.total_ppm * 1e-6);
}
// fmod(x, 2 pi) for the phases an NCO produces.  fmod is exact — its result x - k*y (k = trunc(x/y)) is representable — so
// one FMA with the right k returns it bit for bit; k from a product with 1/y can be off by one either way, which the sign
// / range of the remainder shows.  Anything outside [0, 1e9) goes to the library routine.
__device__ __forceinline__ double fmod_two_pi(double x) {
  const double y = 2.0 * 3.14159265358979323846;
  if (x >= 0.0 && x < 1e9) {
    // branch-free on the recurrence: the three candidate remainders (k - 1, k, k + 1) are independent FMAs, the right one
    // is the one in [0, y)
    const double k = floor(x * (1.0 / y));
    const double r0 = __fma_rn(-k, y, x);
    const double rm = __fma_rn(-(k - 1.0), y, x);
    const double rp = __fma_rn(-(k + 1.0), y, x);
    return (r0 < 0.0) ? rm : ((r0 >= y) ? rp : r0);
  }
  return fmod(x, y);
}
__device__ __forceinline__ void nco_wrap(NcoState& S) {                              // :188-193
  const double two_pi = 2.0 * 3.14159265358979323846;
  S.phase = fmod_two_pi(S.phase);
  if (S.phase < 0) S.phase += two_pi;
}
__device__ __forceinline__ void nco_init(const NcoParams& P, NcoState& S, unsigned long long seed, unsigned long long stream) {
  S.phase = P.phase0;
  S.counter = 0;
  S.drift_counter = 0;
  S.pad = 0;
  if (P.max_ppm > 0.0) S.static_ppm = (rng_double(seed, stream, S.counter++) * 2.0 - 1.0) * P.max_ppm;   // :124-128
  else S.static_ppm = 0.0;
  S.drift_ppm = 0.0;
  S.total_ppm = S.static_ppm;
  nco_update_freq(P, S);
  nco_wrap(S);
}
// NextSample() :69-79 up to the new phase; the caller takes e^{j phase} (:77).  `inc` is the phase increment of :73,
// 2 pi f / fs: the reference recomputes it for every sample, but its operands only change when the drift step below fires
// (every drift_interval samples), so the quotient is cached between steps — same operands, same value.
__device__ __forceinline__ double nco_increment(const NcoParams& P, const NcoState& S) {
  return 2.0 * 3.14159265358979323846 * S.cur_freq / P.fs;                           // :73
}
__device__ __forceinline__ void nco_advance(const NcoParams& P, NcoState& S, unsigned long long seed, unsigned long long stream,
                                            double& inc) {
  if (P.max_ppm > 0.0) {                                                             // :144-149 (else: cur_freq = base_freq)
    S.drift_counter++;
    if (S.drift_counter >= P.drift_interval) {
      S.drift_counter = 0;
      const double step_std = P.max_ppm * 0.001;
      const double step = (rng_double(seed, stream, S.counter++) * 2.0 - 1.0) * step_std;
      S.drift_ppm += step;
      S.total_ppm = S.static_ppm + S.drift_ppm;
      if (S.total_ppm > P.max_ppm) { S.total_ppm = P.max_ppm; S.drift_ppm = S.total_ppm - S.static_ppm; }
      else if (S.total_ppm < -P.max_ppm) { S.total_ppm = -P.max_ppm; S.drift_ppm = S.total_ppm - S.static_ppm; }
      nco_update_freq(P, S);
      inc = nco_increment(P, S);
    }
  }
  S.phase += inc;
  nco_wrap(S);
}

struct ChanState {
  NcoState tx, rx;
  unsigned long long noise_pos;   // samples generated so far (noise counter = 2*pos, 2*pos+1)
  unsigned long long pad;
};

__global__ void chan_init_kernel(const ChanArgs a, ChanState* st) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const unsigned long long ch = (unsigned long long)(a.first_channel + c);
  ChanState S;
  nco_init(a.tx, S.tx, a.seed, 4 * ch + 0);
  nco_init(a.rx, S.rx, a.seed, 4 * ch + 1);
  S.noise_pos = 0;
  S.pad = 0;
  st[c] = S;
}

// The simulator in three passes.  Only the NCO phase is a recurrence (phase += inc, wrap, a drift step every millisecond
// of samples: LocalOscilator.cs:69-79,144-186); everything else — e^{j phase} of both oscillators, the Box-Muller noise
// pair, the multipath sum, the two complex mixes — depends on the sample index alone.
//   chan_phase_kernel   one thread per (channel, oscillator): the phase sequence, to scratch [C][ldp] doubles
//   chan_mix_kernel     one thread per sample: the rest, operation by operation as before
//   chan_hist_kernel    one thread per channel: slides the multipath history
// (The one-thread-per-channel version did all of it in the serial loop: 6.1 ms for 2048 channels x 4196 samples.)
__global__ void __launch_bounds__(64)
    chan_phase_kernel(const ChanArgs a, ChanState* st, long long L, double* __restrict__ ph_tx, double* __restrict__ ph_rx,
                      long long ldp) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = idx >> 1, which = idx & 1;
  if (c >= a.C) return;
  const unsigned long long ch = (unsigned long long)(a.first_channel + c);
  const NcoParams P = which ? a.rx : a.tx;
  NcoState S = which ? st[c].rx : st[c].tx;
  double* out = (which ? ph_rx : ph_tx) + (long long)c * ldp;
  if (P.max_ppm <= 0.0) S.cur_freq = P.base_freq;                                    // :144-149, every sample in the reference
  double inc = nco_increment(P, S);
  for (long long n = 0; n < L; ++n) {
    nco_advance(P, S, a.seed, 4 * ch + (unsigned long long)which, inc);
    out[n] = S.phase;
  }
  if (which) {
    st[c].rx = S;
  } else {
    st[c].tx = S;
    st[c].noise_pos += (unsigned long long)L;        // the mixing pass of this slab counts back from here
  }
}

__global__ void __launch_bounds__(256)
    chan_mix_kernel(const ChanArgs a, const ChanState* __restrict__ st, const float2* __restrict__ hist, const float2* __restrict__ x,
                    long long n0, long long Ls, long long ldx, float2* __restrict__ y, long long ldy,
                    const double* __restrict__ ph_tx, const double* __restrict__ ph_rx, long long ldp) {
  for (int c = blockIdx.y; c < a.C; c += gridDim.y) {
    const unsigned long long ch = (unsigned long long)(a.first_channel + c);
    const float2* xc = x + (long long)c * ldx;
    float2* yc = y + (long long)c * ldy;
    const float2* hc = hist + (long long)c * kPathHist;   // last kPathHist inputs of earlier calls, oldest first
    const unsigned long long pos0 = st[c].noise_pos - (unsigned long long)Ls;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Ls; i += (long long)gridDim.x * blockDim.x) {
      const long long n = n0 + i;
      float xr, xi;
      if (a.n_paths > 0) {
        float accR = 0.f, accI = 0.f;
        for (int k = 0; k < a.n_paths; ++k) {
          const long long m = n - a.delay[k];
          const float2 v = (m >= 0) ? xc[m] : hc[kPathHist + m];
          const float gr = a.gain[2 * k], gi = a.gain[2 * k + 1];
          const float p1 = gr * v.x, p2 = gi * v.y, p3 = gr * v.y, p4 = gi * v.x;
          accR = accR + (p1 - p2);
          accI = accI + (p3 + p4);
        }
        xr = accR; xi = accI;
      } else {
        const float2 v = xc[n];
        xr = v.x; xi = v.y;
      }
      double tr, ti, rr, ri;
      sincos(ph_tx[(long long)c * ldp + i], &ti, &tr);             // NextSample :77
      sincos(ph_rx[(long long)c * ldp + i], &ri, &rr);
      ri = -ri;                                                    // Complex.Conjugate
      double dr = (double)xr, di = (double)xi, yr, yi;
      if (a.mode == 0) {                                           // x * (tx * conj(rx))
        const double pr = tr * rr - ti * ri, pi = tr * ri + ti * rr;
        yr = dr * pr - di * pi; yi = dr * pi + di * pr;
      } else {                                                     // ((x + noise) * tx) * conj(rx)
        if (a.noise_rms > 0.f) {
          const unsigned long long k = 2ULL * (pos0 + (unsigned long long)i);
          const double u1 = 1.0 - rng_double(a.seed, 4 * ch + 2, k);
          const double u2 = 1.0 - rng_double(a.seed, 4 * ch + 2, k + 1);
          const double mag = sqrt(-2.0 * log(u1)) * (double)a.noise_rms;
          const double ph = 2.0 * 3.14159265358979323846 * u2;
          double sn, cs;
          sincos(ph, &sn, &cs);
          float ni = (float)(mag * cs), nq = (float)(mag * sn);
          ni = fminf(fmaxf(ni, -1.f), 1.f);
          nq = fminf(fmaxf(nq, -1.f), 1.f);
          dr = dr + (double)ni; di = di + (double)nq;
        }
        const double ar = dr * tr - di * ti, ai = dr * ti + di * tr;
        yr = ar * rr - ai * ri; yi = ar * ri + ai * rr;
      }
      yc[n] = make_float2((float)yr, (float)yi);
    }
  }
}

// slide the multipath history: keep the newest kPathHist inputs of (history ++ x)
__global__ void chan_hist_kernel(const ChanArgs a, float2* hist, const float2* __restrict__ x, long long L, long long ldx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float2* xc = x + (long long)c * ldx;
  float2* hc = hist + (long long)c * kPathHist;
  if (L >= kPathHist) {
    for (int i = 0; i < kPathHist; ++i) hc[i] = xc[L - kPathHist + i];
  } else {
    const int keep = kPathHist - (int)L;
    for (int i = 0; i < keep; ++i) hc[i] = hc[i + (int)L];
    for (int i = 0; i < (int)L; ++i) hc[keep + i] = xc[i];
  }
}

__global__ void fill_bytes_kernel(unsigned long long seed, int first_channel, int C, long long n, uint8_t* out) {
  const long long total = (long long)C * n;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / n, k = idx - c * n;
    out[idx] = (uint8_t)(rng_u64(seed, 4ULL * (unsigned long long)(first_channel + c) + 3, (unsigned long long)k) >> 56);
  }
}

// bits[c][8k+j] = bit (7-j) of bytes[c][k]  (BitPacker.BytesToBitString, HelperFunctions.cs:14-29), as bytes 0/1
__global__ void unpack_bits_kernel(const uint8_t* __restrict__ bytes, long long n_bytes, long long bytes_stride, int C,
                                   uint8_t* __restrict__ bits, long long bits_stride) {
  const long long total = (long long)C * n_bytes;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / n_bytes, k = idx - c * n_bytes;
    const unsigned v = bytes[c * bytes_stride + k];
    uint8_t* o = bits + c * bits_stride + 8 * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (uint8_t)((v >> (7 - j)) & 1u);
  }
}

// packed[c][k] = bits 8k..8k+7 of channel c, MSB first (BitPacker.BitsToBytes with bitOffset 0,
// HelperFunctions.cs:32-57), except that a trailing incomplete byte is kept, zero-padded on the right — the bit
// count travels separately.  One thread per output byte, one 8-byte load when the row is 8-byte aligned.
__global__ void pack_bits_kernel(const uint8_t* __restrict__ bits, long long bits_stride, const long long* __restrict__ n_bits,
                                 int C, uint8_t* __restrict__ packed, long long packed_stride, long long max_bytes) {
  const long long total = (long long)C * max_bytes;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / max_bytes, k = idx - c * max_bytes;
    const long long nb = n_bits[c];
    if (8 * k >= nb) continue;
    const uint8_t* b = bits + c * bits_stride + 8 * k;
    unsigned v = 0;
    if (8 * k + 8 <= nb && ((reinterpret_cast<uintptr_t>(b) & 7) == 0)) {
      const unsigned long long w = *reinterpret_cast<const unsigned long long*>(b);
#pragma unroll
      for (int j = 0; j < 8; ++j) v |= (unsigned)((w >> (8 * j)) & 1ull) << (7 - j);
    } else {
      const int m = (int)((nb - 8 * k) < 8 ? (nb - 8 * k) : 8);
      for (int j = 0; j < m; ++j) v |= (unsigned)(b[j] & 1u) << (7 - j);
    }
    packed[c * packed_stride + k] = (uint8_t)v;
  }
}

// one CTA of four warps per channel (one warp per channel left every lane a chain of ~70 dependent load batches on an
// 8400-bit burst: 15 us for 2048 channels, against ~5 for the bytes it reads)
constexpr int kBerThreads = 128;
__global__ void __launch_bounds__(kBerThreads)
    ber_kernel(const uint8_t* __restrict__ rx, long long rx_stride, const long long* __restrict__ n_rx,
               const uint8_t* __restrict__ ref, long long ref_stride, long long n_ref, int C, uint32_t* counters) {
  __shared__ unsigned warp_err[kBerThreads / 32];
  const int c = blockIdx.x;
  if (c >= C) return;
  const int lane = threadIdx.x;                       // position inside the channel's CTA
  const uint8_t* r = rx + (long long)c * rx_stride;
  const uint8_t* f = ref + (long long)c * ref_stride;
  const long long have = n_rx[c];
  const long long n = have < n_ref ? have : n_ref;
  unsigned err = 0;
  // head bytes up to a 4-byte boundary of r, then one aligned word of r against two aligned words of f funnel-shifted
  // to the same bytes (byte loads made this a 128-iteration latency chain per lane), then the tail bytes
  const long long head = (long long)((4 - (reinterpret_cast<uintptr_t>(r) & 3)) & 3) < n
                             ? (long long)((4 - (reinterpret_cast<uintptr_t>(r) & 3)) & 3) : n;
  if (lane < head) err += (r[lane] != f[lane]) ? 1u : 0u;
  const long long nw = (n - head) >> 2;
  if (nw > 0) {
    const uint32_t* rw = reinterpret_cast<const uint32_t*>(r + head);
    const uint8_t* fb = f + head;
    const int foff = (int)(reinterpret_cast<uintptr_t>(fb) & 3);
    const uint32_t* fw = reinterpret_cast<const uint32_t*>(fb - foff);
    const int sh = 8 * foff;
#pragma unroll 4
    for (long long i = lane; i < nw; i += kBerThreads) {
      const uint32_t a = rw[i];
      const uint32_t lo = fw[i];
      const uint32_t hi = sh ? fw[i + 1] : 0u;               // holds bytes of this word whenever sh != 0
      const uint32_t b = __funnelshift_r(lo, hi, sh);
      err += (unsigned)__popc(__vcmpne4(a, b)) >> 3;
    }
  }
  for (long long i = head + 4 * nw + lane; i < n; i += kBerThreads) err += (r[i] != f[i]) ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
  if ((lane & 31) == 0) warp_err[lane >> 5] = err;
  __syncthreads();
  if (lane == 0) {
    unsigned tot = 0;
#pragma unroll
    for (int w = 0; w < kBerThreads / 32; ++w) tot += warp_err[w];
    counters[2 * c] = tot + (uint32_t)(n_ref - n);   // bits that never arrived count as errors
    counters[2 * c + 1] = (uint32_t)n_ref;
  }
}

}  // namespace qpsk

using namespace qpsk;

struct qpsk_chan {
  ChanArgs args;
  int channels = 0;
  int device = 0;
  DevBuf<ChanState> d_state;
  DevBuf<float2> d_hist, d_in, d_out;
  DevBuf<double> d_phase;          // [2][channels][slab] oscillator phases of the slab being mixed
  cudaStream_t stream = nullptr;
  ~qpsk_chan() {
    if (stream) cudaStreamDestroy(stream);
  }
};

extern "C" {

int qpsk_chan_create(const qpsk_chan_params* p, int channels, int first_channel, qpsk_chan** out) {
  if (!p || !out) return QPSK_ERR_NULL;
  *out = nullptr;
  if (channels <= 0 || first_channel < 0) return QPSK_ERR_RANGE;
  if (!(p->sample_rate_hz > 0)) return QPSK_ERR_RANGE;               // LocalOscilator.cs:48-49
  if (p->n_paths < 0 || p->n_paths > kMaxPaths) return QPSK_ERR_RANGE;
  for (int k = 0; k < p->n_paths; ++k)
    if (p->path_delay[k] < 0 || p->path_delay[k] >= kPathHist) return QPSK_ERR_RANGE;
  if (p->mode != 0 && p->mode != 1) return QPSK_ERR_RANGE;
  QPSK_TRY(ensure_device());
  qpsk_chan* c = new (std::nothrow) qpsk_chan();
  if (!c) return QPSK_ERR_NOMEM;
  c->device = current_device();
  ChanArgs& a = c->args;
  auto nco = [&](double f, double ppm, double ph) {
    NcoParams n;
    n.base_freq = f; n.fs = p->sample_rate_hz; n.max_ppm = fabs(ppm); n.phase0 = ph;             // :56
    const double iv = p->sample_rate_hz * 1e-3;                                                    // :59
    n.drift_interval = (int)(iv > 1.0 ? iv : 1.0);
    return n;
  };
  a.tx = nco(p->tx_freq_hz, p->tx_ppm, p->tx_phase0);
  a.rx = nco(p->rx_freq_hz, p->rx_ppm, p->rx_phase0);
  a.noise_rms = (p->noise_dbfs <= -300.0f) ? 0.0f : (float)pow(10.0, p->noise_dbfs / 20.0);        // HelperModels.cs:23
  a.mode = p->mode;
  a.n_paths = p->n_paths;
  for (int k = 0; k < kMaxPaths; ++k) {
    a.gain[2 * k] = p->path_gain_iq[2 * k];
    a.gain[2 * k + 1] = p->path_gain_iq[2 * k + 1];
    a.delay[k] = p->path_delay[k];
  }
  a.seed = p->seed;
  a.first_channel = first_channel;
  a.C = channels;
  c->channels = channels;
  int st = QPSK_OK;
  do {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { st = QPSK_ERR_CUDA; break; }
    if ((st = c->d_state.alloc((size_t)channels)) != QPSK_OK) break;
    if ((st = c->d_hist.alloc((size_t)channels * kPathHist)) != QPSK_OK) break;
    if ((st = c->d_hist.zero(c->stream)) != QPSK_OK) break;
    chan_init_kernel<<<(channels + 127) / 128, 128, 0, c->stream>>>(a, c->d_state.p);
    count_launch();
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { st = QPSK_ERR_CUDA; break; }
  } while (0);
  if (st != QPSK_OK) { delete c; return st; }
  *out = c;
  return QPSK_OK;
}

int qpsk_chan_destroy(qpsk_chan* c) {
  if (c) {
    if (c->stream) cudaStreamSynchronize(c->stream);
    delete c;
  }
  return QPSK_OK;
}

int qpsk_chan_apply_dev(qpsk_chan* c, const float* d_x, int64_t n_floats, int64_t x_stride_floats, float* d_y,
                        int64_t y_stride_floats, void* stream) {
  if (!c) return QPSK_ERR_NULL;
  if (n_floats < 0) return QPSK_ERR_RANGE;
  if ((n_floats & 1) || (x_stride_floats & 1) || (y_stride_floats & 1)) return QPSK_ERR_ARG;
  if (n_floats == 0) return QPSK_OK;
  if (!d_x || !d_y) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(c->device));
  cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
  // time slabs bound the phase scratch (2 x 128 MB); an in-place call is allowed without multipath only, as before
  const long long L = n_floats >> 1;
  const int C = c->channels;
  long long slab = (16LL << 20) / C;
  if (slab < 256) slab = 256;
  if (slab > L) slab = L;
  QPSK_TRY(c->d_phase.ensure((size_t)2 * C * slab));
  double* ph_tx = c->d_phase.p;
  double* ph_rx = c->d_phase.p + (size_t)C * slab;
  for (long long n0 = 0; n0 < L; n0 += slab) {
    const long long Ls = (L - n0 < slab) ? (L - n0) : slab;
    chan_phase_kernel<<<(2 * C + 63) / 64, 64, 0, s>>>(c->args, c->d_state.p, Ls, ph_tx, ph_rx, slab);
    QPSK_LAUNCH_CHECK();
    long long bx = (Ls + 255) / 256;
    if (bx > 64) bx = 64;
    chan_mix_kernel<<<dim3((unsigned)bx, (unsigned)(C < 65535 ? C : 65535)), 256, 0, s>>>(
        c->args, c->d_state.p, c->d_hist.p, (const float2*)d_x, n0, Ls, x_stride_floats >> 1, (float2*)d_y, y_stride_floats >> 1, ph_tx,
        ph_rx, slab);
    QPSK_LAUNCH_CHECK();
  }
  if (c->args.n_paths > 0) {
    chan_hist_kernel<<<(C + 127) / 128, 128, 0, s>>>(c->args, c->d_hist.p, (const float2*)d_x, L, x_stride_floats >> 1);
    QPSK_LAUNCH_CHECK();
  }
  return QPSK_OK;
}

int qpsk_chan_apply(qpsk_chan* c, const float* x, int64_t n_floats, float* y) {
  if (!c) return QPSK_ERR_NULL;
  if (n_floats < 0) return QPSK_ERR_RANGE;
  if (n_floats & 1) return QPSK_ERR_ARG;
  if (n_floats == 0) return QPSK_OK;
  if (!x || !y) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(c->device));
  const size_t tot = (size_t)(n_floats >> 1) * c->channels;
  QPSK_TRY(c->d_in.ensure(tot));
  QPSK_TRY(c->d_out.ensure(tot));
  QPSK_CUDA_TRY(cudaMemcpyAsync(c->d_in.p, x, tot * 8, cudaMemcpyHostToDevice, c->stream));
  QPSK_TRY(qpsk_chan_apply_dev(c, (const float*)c->d_in.p, n_floats, n_floats, (float*)c->d_out.p, n_floats, c->stream));
  QPSK_CUDA_TRY(cudaMemcpyAsync(y, c->d_out.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
  QPSK_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return QPSK_OK;
}

int qpsk_fill_bytes_dev(uint64_t seed, int first_channel, int channels, int64_t n_bytes, uint8_t* d_out, void* stream) {
  if (channels < 0 || n_bytes < 0 || first_channel < 0) return QPSK_ERR_RANGE;
  if (channels == 0 || n_bytes == 0) return QPSK_OK;
  if (!d_out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  const long long total = (long long)channels * n_bytes;
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  fill_bytes_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seed, first_channel, channels, n_bytes, d_out);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_unpack_bits_dev(const uint8_t* d_bytes, int64_t n_bytes, int64_t bytes_stride, int channels, uint8_t* d_bits,
                         int64_t bits_stride, void* stream) {
  if (channels < 0 || n_bytes < 0) return QPSK_ERR_RANGE;
  if (channels == 0 || n_bytes == 0) return QPSK_OK;
  if (!d_bytes || !d_bits) return QPSK_ERR_NULL;
  if (channels > 1 && bits_stride < 8 * n_bytes) return QPSK_ERR_ARG;
  QPSK_TRY(ensure_device());
  const long long total = (long long)channels * n_bytes;
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  unpack_bits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_bytes, n_bytes, bytes_stride, channels, d_bits, bits_stride);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_pack_bits_dev(const uint8_t* d_bits, int64_t bits_stride, const int64_t* d_n_bits, int64_t max_bits, int channels,
                       uint8_t* d_packed, int64_t packed_stride, void* stream) {
  if (channels < 0 || max_bits < 0) return QPSK_ERR_RANGE;
  if (channels == 0 || max_bits == 0) return QPSK_OK;
  if (!d_bits || !d_n_bits || !d_packed) return QPSK_ERR_NULL;
  const long long max_bytes = (max_bits + 7) / 8;
  if (channels > 1 && packed_stride < max_bytes) return QPSK_ERR_ARG;
  QPSK_TRY(ensure_device());
  const long long total = (long long)channels * max_bytes;
  long long blocks = (total + 255) / 256;
  const long long cap = 16LL * device_sm_count();
  if (blocks > cap) blocks = cap;
  pack_bits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_bits, bits_stride, (const long long*)d_n_bits, channels,
                                                                 d_packed, packed_stride, max_bytes);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_ber_count_dev(const uint8_t* d_rx_bits, int64_t rx_stride, const int64_t* d_n_rx, const uint8_t* d_ref_bits,
                       int64_t ref_stride, int64_t n_ref, int channels, uint32_t* d_counters, void* stream) {
  if (channels < 0 || n_ref < 0) return QPSK_ERR_RANGE;
  if (channels == 0) return QPSK_OK;
  if (!d_rx_bits || !d_n_rx || !d_counters) return QPSK_ERR_NULL;
  if (n_ref > 0 && !d_ref_bits) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  ber_kernel<<<channels, kBerThreads, 0, (cudaStream_t)stream>>>(d_rx_bits, rx_stride, (const long long*)d_n_rx, d_ref_bits,
                                                                  ref_stride, n_ref, channels, d_counters);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

}  // extern "C"
