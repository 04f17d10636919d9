// chain.cu — the full receive chain as wired by hand in TB/Simulated/testFullDemodChain.cs:22-108
// (README.md:16): band-edge FLL -> RRC matched filter -> Mueller-Muller -> Costas, batched over
// independent channels, with the three streams that test publishes as ZMQ topics as outputs:
//   "baseband"                          the FLL output, one cf32 per input sample        (:73-76, :88-95)
//   "baseband_PostSymbolSync"           the Mueller-Muller symbols                        (:84, :97-103)
//   "baseband_PostSymbolSyncPostCostas" those symbols after the Costas loop               (:106-110)
// The reference interleaves FLL and matched filter sample by sample and runs MM / Costas once per
// 4096-sample frame; neither block feeds back into an earlier one, so block-wise FLL -> MF over the whole
// call followed by MM and Costas gives the same three streams (the blocks themselves are chunk-invariant:
// state is carried in device memory across calls).  Upstream the test no longer constructs (SURVEY §4:
// FLLBandEdgeFilter gained a bandwidth argument, MuellerMuller a required gain pair); this is the repaired
// wiring with every block parameter explicit.
// Compiled with --fmad=false like the loop kernels it drives.
#include <math.h>

#include "fir.cuh"
#include "loops.cuh"

namespace qpsk {

struct ChainEngine {
  int channels = 1;
  int device = 0;
  FllEngine fll;
  FirEngine mf;
  MmEngine mm;
  CostasEngine costas;
  double mm_sps = 0.0;
  DevBuf<float2> t_mf, h_in, h_bb, h_sym, h_cos;
  DevBuf<int> d_nsym;
  cudaStream_t stream = nullptr;

  ~ChainEngine() {
    if (stream) cudaStreamDestroy(stream);
  }

  int init(const qpsk_chain_params& p, int channels_in) {
    if (channels_in <= 0) return QPSK_ERR_RANGE;
    if (p.symbol_rate == 0) return QPSK_ERR_RANGE;
    // the reference ctor (MuellerMuller.cs:38-50) validates nothing; with sps <= 0.1 its loop (:62-120) never advances
    if (!(p.mm_sps > 0.1)) return QPSK_ERR_UNSUPPORTED;
    QPSK_TRY(ensure_device());
    device = current_device();
    channels = channels_in;
    mm_sps = p.mm_sps;
    QPSK_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    const std::vector<double> h = design_rrc(p.rrc_span, p.rrc_alpha, p.sample_rate, p.symbol_rate);
    const std::vector<float> iq = real_taps_as_iq(h);
    QPSK_TRY(mf.init(iq.data(), (int)iq.size(), channels));
    mf.mode = QPSK_FIR_FMA;
    QPSK_TRY(fll.init(p.fll_sps, p.fll_rolloff, p.fll_size, p.fll_bw, channels));
    QPSK_TRY(mm.init(p.mm_sps, p.mm_kp, p.mm_ki, channels));
    QPSK_TRY(costas.init(p.costas_sample_rate, p.costas_bw_hz, p.costas_damping, channels));
    QPSK_TRY(d_nsym.alloc((size_t)channels));
    return QPSK_OK;
  }

  // every symbol advances time by >= sps - 0.1 samples (MuellerMuller.cs:87-89)
  long long symbols_bound(int64_t L) const {
    if (mm_sps - 0.1 <= 1.0) return L;
    long long b = (long long)((double)(L + 4) / (mm_sps - 0.1)) + 2;
    return b < L ? b : L;
  }

  int process_dev(const float2* x, int64_t L, int64_t ldx, float2* bb, int64_t ld_bb, float2* sym, int64_t ld_sym,
                  float2* cos_out, int64_t ld_cos, int* n_sym, cudaStream_t s) {
    if (!s) s = stream;
    if (L == 0) {
      QPSK_CUDA_TRY(cudaMemsetAsync(n_sym, 0, sizeof(int) * channels, s));
      return QPSK_OK;
    }
    const long long need = symbols_bound(L);
    if (ld_sym < need || ld_cos < need) return QPSK_ERR_CAPACITY;
    const int64_t ld = L + (L & 1);
    QPSK_TRY(t_mf.ensure((size_t)ld * channels));
    QPSK_TRY(fll.process_dev(x, bb, L, ldx, ld_bb, s));                               // :73
    QPSK_TRY(mf.filter_dev(bb, t_mf.p, L, ld_bb, ld, s));                             // :79-80
    QPSK_TRY(mm.process_dev(t_mf.p, L, ld, sym, 2 * ld_sym, ld_sym, n_sym, s));       // :84
    QPSK_TRY(costas.process_dev(sym, cos_out, need, ld_sym, ld_cos, n_sym, s));       // :106
    return QPSK_OK;
  }
};

}  // namespace qpsk

using namespace qpsk;

struct qpsk_chain {
  ChainEngine eng;
};

extern "C" {

int qpsk_chain_default_params(qpsk_chain_params* p) {
  if (!p) return QPSK_ERR_NULL;
  // the literal values of testFullDemodChain.cs:18-45
  const int fs = 10000000, rs = fs / 30;
  p->sample_rate = fs;
  p->symbol_rate = rs;
  p->rrc_span = 11.0;
  p->rrc_alpha = 0.9;
  p->fll_sps = (float)(fs / rs);
  p->fll_rolloff = 0.9f;
  p->fll_size = 10;
  p->fll_bw = 0.1f;
  const double bn = 0.000000002, zeta = 1.0 / sqrt(2.0), kd = 1.0;   // :25-27 (IEEE division of the rounded root)
  const double omega = 2.0 * 3.14159265358979323846 * bn;
  p->mm_sps = (double)(fs / rs);
  p->mm_kp = 2.0 * zeta * omega / kd;
  p->mm_ki = omega * omega / kd;
  p->costas_sample_rate = (double)rs;
  p->costas_bw_hz = (double)(rs / 10);
  p->costas_damping = 0.707;                                  // CostasLoopQpsk.cs:29 default
  return QPSK_OK;
}

int qpsk_chain_create(const qpsk_chain_params* p, int channels, qpsk_chain** out) {
  if (!p || !out) return QPSK_ERR_NULL;
  *out = nullptr;
  qpsk_chain* c = new (std::nothrow) qpsk_chain();
  if (!c) return QPSK_ERR_NOMEM;
  int st = c->eng.init(*p, channels);
  if (st != QPSK_OK) { delete c; return st; }
  *out = c;
  return QPSK_OK;
}

int qpsk_chain_destroy(qpsk_chain* c) {
  if (c) {
    cudaSetDevice(c->eng.device);
    if (c->eng.stream) cudaStreamSynchronize(c->eng.stream);
    delete c;
  }
  return QPSK_OK;
}

int qpsk_chain_set_fir_mode(qpsk_chain* c, int mode) {
  if (!c) return QPSK_ERR_NULL;
  if (mode != QPSK_FIR_FAST && mode != QPSK_FIR_EXACT) return QPSK_ERR_RANGE;
  c->eng.mf.mode = mode == QPSK_FIR_FAST ? QPSK_FIR_FMA : mode;   // chunk-invariant bit for bit (the split kernel is not)
  return QPSK_OK;
}

int qpsk_chain_symbols_bound(const qpsk_chain* c, int64_t n_floats, int64_t* cap_floats) {
  if (!c || !cap_floats) return QPSK_ERR_NULL;
  if (n_floats < 0) return QPSK_ERR_RANGE;
  *cap_floats = 2 * c->eng.symbols_bound(n_floats >> 1);
  return QPSK_OK;
}

int qpsk_chain_process_dev(qpsk_chain* c, const float* d_in, int64_t n_floats, int64_t in_stride, float* d_baseband,
                           int64_t bb_stride, float* d_sync, int64_t sync_stride, float* d_costas, int64_t costas_stride,
                           int* d_n_sym, void* stream) {
  if (!c || !d_n_sym) return QPSK_ERR_NULL;
  if (n_floats < 0) return QPSK_ERR_RANGE;
  if ((n_floats & 1) || (in_stride & 1) || (bb_stride & 1) || (sync_stride & 1) || (costas_stride & 1)) return QPSK_ERR_ARG;
  if (n_floats > 0 && (!d_in || !d_baseband || !d_sync || !d_costas)) return QPSK_ERR_NULL;
  if (c->eng.channels > 1 && (in_stride < n_floats || bb_stride < n_floats)) return QPSK_ERR_ARG;
  QPSK_TRY(ensure_device(c->eng.device));
  return c->eng.process_dev((const float2*)d_in, n_floats >> 1, in_stride >> 1, (float2*)d_baseband, bb_stride >> 1,
                            (float2*)d_sync, sync_stride >> 1, (float2*)d_costas, costas_stride >> 1, d_n_sym,
                            (cudaStream_t)stream);
}

int qpsk_chain_process(qpsk_chain* c, const float* iq_in, int64_t n_floats, float* baseband_out, float* sync_out,
                       float* costas_out, int64_t sym_cap_floats, int* n_sym) {
  if (!c || !n_sym) return QPSK_ERR_NULL;
  if (n_floats < 0 || sym_cap_floats < 0) return QPSK_ERR_RANGE;
  if ((n_floats & 1) != 0) return QPSK_ERR_ARG;
  ChainEngine& e = c->eng;
  if (n_floats == 0) {
    for (int i = 0; i < e.channels; ++i) n_sym[i] = 0;
    return QPSK_OK;
  }
  if (!iq_in || !baseband_out || !sync_out || !costas_out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(c->eng.device));
  const int64_t L = n_floats >> 1;
  const int64_t ld = L + (L & 1);
  const int64_t lds = e.symbols_bound(L);
  if ((sym_cap_floats >> 1) < lds) return QPSK_ERR_CAPACITY;
  const int C = e.channels;
  QPSK_TRY(e.h_in.ensure((size_t)ld * C));
  QPSK_TRY(e.h_bb.ensure((size_t)ld * C));
  QPSK_TRY(e.h_sym.ensure((size_t)lds * C));
  QPSK_TRY(e.h_cos.ensure((size_t)lds * C));
  cudaStream_t s = e.stream;
  QPSK_CUDA_TRY(cudaMemcpy2DAsync(e.h_in.p, (size_t)ld * 8, iq_in, (size_t)L * 8, (size_t)L * 8, (size_t)C, cudaMemcpyHostToDevice, s));
  QPSK_TRY(e.process_dev(e.h_in.p, L, ld, e.h_bb.p, ld, e.h_sym.p, lds, e.h_cos.p, lds, e.d_nsym.p, s));
  QPSK_CUDA_TRY(cudaMemcpy2DAsync(baseband_out, (size_t)L * 8, e.h_bb.p, (size_t)ld * 8, (size_t)L * 8, (size_t)C, cudaMemcpyDeviceToHost, s));
  QPSK_CUDA_TRY(cudaMemcpyAsync(n_sym, e.d_nsym.p, sizeof(int) * C, cudaMemcpyDeviceToHost, s));
  // rows of the symbol outputs are sym_cap_floats apart on the host, lds complex apart on the device
  const size_t w = (size_t)lds * 8;
  QPSK_CUDA_TRY(cudaMemcpy2DAsync(sync_out, (size_t)sym_cap_floats * 4, e.h_sym.p, w, w, (size_t)C, cudaMemcpyDeviceToHost, s));
  QPSK_CUDA_TRY(cudaMemcpy2DAsync(costas_out, (size_t)sym_cap_floats * 4, e.h_cos.p, w, w, (size_t)C, cudaMemcpyDeviceToHost, s));
  QPSK_CUDA_TRY(cudaStreamSynchronize(s));
  return QPSK_OK;
}

int qpsk_chain_loop_state(qpsk_chain* c, float* fll_phase, float* fll_freq, double* mm_mu, double* costas_theta,
                          double* costas_freq) {
  if (!c) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(c->eng.device));
  ChainEngine& e = c->eng;
  QPSK_CUDA_TRY(cudaStreamSynchronize(e.stream));
  QPSK_CUDA_TRY(cudaDeviceSynchronize());
  const size_t C = (size_t)e.channels;
  if (fll_phase || fll_freq) {
    std::vector<float2> pf(C);
    QPSK_CUDA_TRY(cudaMemcpy(pf.data(), e.fll.d_pf.p, C * sizeof(float2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < C; ++i) {
      if (fll_phase) fll_phase[i] = pf[i].x;
      if (fll_freq) fll_freq[i] = pf[i].y;
    }
  }
  if (mm_mu) {
    std::vector<MmState> h(C);
    QPSK_CUDA_TRY(cudaMemcpy(h.data(), e.mm.d_state.p, C * sizeof(MmState), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < C; ++i) mm_mu[i] = h[i].mu;
  }
  if (costas_theta || costas_freq) {
    std::vector<CostasState> h(C);
    QPSK_CUDA_TRY(cudaMemcpy(h.data(), e.costas.d_state.p, C * sizeof(CostasState), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < C; ++i) {
      if (costas_theta) costas_theta[i] = h[i].theta;
      if (costas_freq) costas_freq[i] = h[i].freq;
    }
  }
  return QPSK_OK;
}

}  // extern "C"
