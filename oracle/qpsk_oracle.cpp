// qpsk_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A line-by-line CPU restatement of the hot path of NustyFrozen/QPSK-Modulator-Demodulator
// (C#, net9.0).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; libqpskcuda.so never links or calls it.
//
// PARITY UNPINNED: the reference ships no golden vectors, no known-answer tests and no
// seeded RNG, and no .NET runtime exists in this image, so this
// restatement cannot be checked against outputs of the reference itself.  It is pinned
// only against (i) an independently written numpy twin (oracle/np_twin.py) and (ii) the
// behavioural checklist of SURVEY.md §4 items 1-12 (tests/test_oracle_*.py).
//
// Conventions restated from the C# source ("MS/" = Modulation-Simulation/, "TB/" = TestBench/):
//   * float arithmetic is IEEE binary32 with every operation rounded separately
//     (RyuJIT never contracts mul+add into FMA) -> build with -ffp-contract=off -mno-fma.
//   * System.Numerics.Vector<float>.Count == 8 (AVX2; stays 256-bit on AVX-512 hosts).
//   * Math.Round is round-half-to-even -> std::nearbyint under FE_TONEAREST.
//   * System.Random (unseeded in the reference) is replaced by the counter-based generator
//     documented in DESIGN.md ("counter RNG"), so that host and device regenerate the same
//     impairments from (seed, stream, counter).
//
// Build: see oracle/Makefile (g++ -O2 -std=c++17 -ffp-contract=off -mavx2 -mno-fma).

#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

// Status codes: the same numbering as include/qpskcuda.h so tests compare error behaviour.
enum {
  ORC_OK = 0,
  ORC_ERR_NULL = -1,      // ArgumentNullException
  ORC_ERR_ARG = -2,       // ArgumentException
  ORC_ERR_RANGE = -3,     // ArgumentOutOfRangeException
  ORC_ERR_CAPACITY = -6,  // caller buffer too small
};

namespace {

constexpr int kLanes = 8;  // Vector<float>.Count on AVX2

// ---------------------------------------------------------------------------------------------
// Counter RNG (replaces System.Random; spec in DESIGN.md).  splitmix64 finaliser.
// ---------------------------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint64_t rng_u64(uint64_t seed, uint64_t stream, uint64_t counter) {
  uint64_t a = mix64(seed + 0x9E3779B97F4A7C15ULL * (stream + 1));
  return mix64(a + 0xD1B54A32D192ED03ULL * (counter + 1));
}
// Random.NextDouble() stand-in: uniform in [0,1) with 53 bits.
inline double rng_double(uint64_t seed, uint64_t stream, uint64_t counter) {
  return (double)(rng_u64(seed, stream, counter) >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------
// a1. RRCFilter.generateCoefficents — MS/Models/RRC-filter.cs:16-75
// ---------------------------------------------------------------------------------------------
std::vector<double> rrc_taps(double spanSymbols, double beta, int sampleRate, int symbolRate) {
  double spsExact = (double)sampleRate / symbolRate;     // :23
  int sps = (int)std::nearbyint(spsExact);               // :24 Math.Round (banker's)
  int spanSymInt = (int)std::nearbyint(spanSymbols);     // :26
  int taps = spanSymInt * sps + 1;                       // :27
  std::vector<double> h((size_t)std::max(taps, 0));
  int mid = (taps - 1) / 2;                              // :31
  const double pi = M_PI;
  const double eps = 1e-8;                               // :33
  for (int n = 0; n < taps; n++) {
    double t = (n - mid) / (double)sps;                  // :38
    double val;
    if (std::fabs(t) < eps) {                            // :41
      val = 1.0 + beta * (4.0 / pi - 1.0);               // :44
    } else if (std::fabs(std::fabs(t) - 1.0 / (4.0 * beta)) < eps) {  // :46
      val = (beta / std::sqrt(2.0)) *
            ((1.0 + 2.0 / pi) * std::sin(pi / (4.0 * beta)) +
             (1.0 - 2.0 / pi) * std::cos(pi / (4.0 * beta)));  // :49-51
    } else {
      double num = std::sin(pi * t * (1.0 - beta)) +
                   4.0 * beta * t * std::cos(pi * t * (1.0 + beta));   // :56-57
      // :58  Math.Pow(x, 2.0) is modelled as the correctly rounded square x*x (glibc's pow(x, 2.0) can be
      // one ulp off, and GCC folds pow(x, 2.0) into x*x anyway); DESIGN.md "transcendentals"
      double fbt = 4.0 * beta * t;
      double den = pi * t * (1.0 - fbt * fbt);
      val = num / den;
    }
    h[n] = val;
  }
  double energy = 0.0;                                   // :66-68
  for (int i = 0; i < taps; i++) energy += h[i] * h[i];
  double norm = std::sqrt(energy);
  for (int i = 0; i < taps; i++) h[i] /= norm;           // :71-72
  return h;
}

// ToInterleavedIQRealTaps — MS/QPSKDeModulator.cs:278-288, MS/QPSKModulator.cs:43-53
std::vector<float> real_taps_to_iq(const std::vector<double>& h) {
  std::vector<float> t(h.size() * 2);
  for (size_t i = 0; i < h.size(); i++) { t[2 * i] = (float)h[i]; t[2 * i + 1] = 0.0f; }
  return t;
}

// ---------------------------------------------------------------------------------------------
// a2-a5. ComplexFIRFilter — MS/Models/FIRFilter.cs:8-232
// ---------------------------------------------------------------------------------------------
struct Fir {
  std::vector<float> taps;              // interleaved IQ clone (:35)
  std::vector<float> tapsIRev, tapsQRev;  // reversed planar (:43-48)
  std::vector<float> delayI2N, delayQ2N;  // double-length planar delay lines (:50-51)
  int nTaps = 0;
  int pos = 0;

  explicit Fir(const float* t, int nFloats) {
    taps.assign(t, t + nFloats);
    nTaps = nFloats >> 1;
    tapsIRev.resize(nTaps); tapsQRev.resize(nTaps);
    for (int k = 0; k < nTaps; k++) {
      int src = (nTaps - 1 - k) << 1;
      tapsIRev[k] = taps[src];
      tapsQRev[k] = taps[src + 1];
    }
    // +kLanes slack so the AVX path may load a full vector at the tail without faulting
    delayI2N.assign((size_t)nTaps * 2 + kLanes, 0.0f);
    delayQ2N.assign((size_t)nTaps * 2 + kLanes, 0.0f);
    pos = 0;
  }
  void reset() {
    std::fill(delayI2N.begin(), delayI2N.end(), 0.0f);
    std::fill(delayQ2N.begin(), delayQ2N.end(), 0.0f);
    pos = 0;
  }

  // ComplexDotWindow — FIRFilter.cs:144-211 (hardware-accelerated branch, w = 8)
  inline void dot(int start, float& outI, float& outQ) const {
    const int N = nTaps;
    const float* xI = delayI2N.data() + start;
    const float* xQ = delayQ2N.data() + start;
    const float* hI = tapsIRev.data();
    const float* hQ = tapsQRev.data();
    const int nVec = N - (N % kLanes);
    float accI = 0.0f, accQ = 0.0f;
    alignas(32) float lI[kLanes], lQ[kLanes];
#if defined(__AVX2__)
    __m256 vAccI = _mm256_setzero_ps(), vAccQ = _mm256_setzero_ps();
    for (int i = 0; i < nVec; i += kLanes) {
      __m256 vXI = _mm256_loadu_ps(xI + i), vXQ = _mm256_loadu_ps(xQ + i);
      __m256 vHI = _mm256_loadu_ps(hI + i), vHQ = _mm256_loadu_ps(hQ + i);
      // :172-173  vAccI += (vHI*vXI) - (vHQ*vXQ);  vAccQ += (vHI*vXQ) + (vHQ*vXI)
      vAccI = _mm256_add_ps(vAccI, _mm256_sub_ps(_mm256_mul_ps(vHI, vXI), _mm256_mul_ps(vHQ, vXQ)));
      vAccQ = _mm256_add_ps(vAccQ, _mm256_add_ps(_mm256_mul_ps(vHI, vXQ), _mm256_mul_ps(vHQ, vXI)));
    }
    _mm256_store_ps(lI, vAccI); _mm256_store_ps(lQ, vAccQ);
#else
    for (int l = 0; l < kLanes; l++) { lI[l] = 0.0f; lQ[l] = 0.0f; }
    for (int i = 0; i < nVec; i += kLanes)
      for (int l = 0; l < kLanes; l++) {
        float a = hI[i + l] * xI[i + l], b = hQ[i + l] * xQ[i + l];
        float c = hI[i + l] * xQ[i + l], d = hQ[i + l] * xI[i + l];
        float e = a - b, f = c + d;
        lI[l] = lI[l] + e; lQ[l] = lQ[l] + f;
      }
#endif
    for (int l = 0; l < kLanes; l++) { accI += lI[l]; accQ += lQ[l]; }   // :176-180
    for (int i = nVec; i < N; i++) {                                      // :183-192
      float xi = xI[i], xq = xQ[i], hi = hI[i], hq = hQ[i];
      float a = hi * xi, b = hq * xq, c = hi * xq, d = hq * xi;
      float e = a - b, f = c + d;
      accI += e; accQ += f;
    }
    outI = accI; outQ = accQ;
  }

  // Filter(float,float,out,out) — FIRFilter.cs:59-77
  inline void filter1(float inI, float inQ, float& outI, float& outQ) {
    int p = pos, pN = p + nTaps;
    delayI2N[p] = inI; delayQ2N[p] = inQ;
    delayI2N[pN] = inI; delayQ2N[pN] = inQ;
    int start = p + 1;
    if (start >= nTaps) start -= nTaps;
    dot(start, outI, outQ);
    p++;
    if (p == nTaps) p = 0;
    pos = p;
  }

  // Filter(ReadOnlySpan<float>, Span<float>) — FIRFilter.cs:80-91
  void filter(const float* in, float* out, int64_t nFloats) {
    for (int64_t s = 0; s < nFloats; s += 2) {
      float yI, yQ;
      filter1(in[s], in[s + 1], yI, yQ);
      out[s] = yI; out[s + 1] = yQ;
    }
  }

  // fftFilter — FIRFilter.cs:96-141.  MathNet's FFT computes the linear convolution in
  // fp64 (Complex) and the result is rounded to fp32; restated as a direct fp64 convolution
  // with the same alignment: y[i] = conv(x,h)[i + nTaps - 1], i < nData (:130-138).
  void fft_filter(const float* in, float* out, int64_t nFloats) const {
    int64_t nData = nFloats >> 1;
    for (int64_t i = 0; i < nData; i++) {
      int64_t n = i + nTaps - 1;  // index into the full convolution
      double accR = 0.0, accI = 0.0;
      // conv[n] = sum_k h[k] * x[n-k], 0<=k<nTaps, 0<=n-k<nData
      int64_t kmin = std::max<int64_t>(0, n - (nData - 1));
      int64_t kmax = std::min<int64_t>(n, nTaps - 1);
      for (int64_t k = kmin; k <= kmax; k++) {
        double hr = taps[2 * k], hi = taps[2 * k + 1];
        double xr = in[2 * (n - k)], xi = in[2 * (n - k) + 1];
        accR += xr * hr - xi * hi;
        accI += xr * hi + xi * hr;
      }
      out[2 * i] = (float)accR;
      out[2 * i + 1] = (float)accI;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// a7-a8. FLLBandEdgeFilter — MS/Models/Band-Edge Filter.cs:14-203
// ---------------------------------------------------------------------------------------------
constexpr float kPiF = 3.14159274101257324f;     // MathF.PI
constexpr float kTwoPiF = 2.0f * kPiF;           // :16

// MathF.Sin / MathF.Cos model: the correctly rounded fp32 result, obtained by evaluating in fp64 and
// rounding once.  (glibc's sinf/cosf are NOT correctly rounded — about 1.3 % of arguments differ by
// one ulp — and neither the Windows CRT behind .NET's MathF nor CUDA's sinf is bit-specified, so the
// oracle and the CUDA path both use this definition; DESIGN.md "transcendentals".)
inline float cr_sinf(float x) { return (float)std::sin((double)x); }
inline float cr_cosf(float x) { return (float)std::cos((double)x); }

inline float sincf_(float x) {                   // :197-202
  if (x == 0.0f) return 1.0f;
  float arg = kPiF * x;
  return cr_sinf(arg) / arg;
}

struct Fll {
  float sps, rolloff, bandwidth;
  int filterSize;
  float phase = 0.0f, freq = 0.0f;
  float alpha, beta, maxFreq, minFreq;
  std::vector<float> tapsLowerIQ, tapsUpperIQ;
  Fir* lower = nullptr;
  Fir* upper = nullptr;

  Fll(float sps_, float rolloff_, int size_, float bw_)
      : sps(sps_), rolloff(rolloff_), bandwidth(bw_), filterSize(size_) {
    alpha = 0.0f;                              // :55
    beta = 4.0f * bandwidth / sps;             // :56
    maxFreq = kTwoPiF * (2.0f / sps);          // :58
    minFreq = -maxFreq;
    design();
  }
  ~Fll() { delete lower; delete upper; }

  void design() {                              // :132-183
    int numTaps = filterSize;
    int mid = (numTaps - 1) / 2;
    std::vector<float> bb(numTaps);
    float sum = 0.0f;
    for (int i = 0; i < numTaps; i++) {
      float k = (float)(i - mid) / (2.0f * sps);
      float pos = rolloff * k;
      float tap = sincf_(pos - 0.5f) + sincf_(pos + 0.5f);
      sum += tap;
      bb[i] = tap;
    }
    for (int i = 0; i < numTaps; i++) bb[i] /= sum;
    tapsLowerIQ.resize((size_t)numTaps << 1);
    tapsUpperIQ.resize((size_t)numTaps << 1);
    for (int i = 0; i < numTaps; i++) {
      float k = (float)(i - mid) / (2.0f * sps);
      float angle = -kTwoPiF * (1.0f + rolloff) * k;
      float wc = cr_cosf(angle), ws = cr_sinf(angle);
      float li = bb[i] * wc, lq = bb[i] * ws;
      int t = i << 1;
      tapsLowerIQ[t] = li; tapsLowerIQ[t + 1] = lq;
      tapsUpperIQ[t] = li; tapsUpperIQ[t + 1] = -lq;
    }
    lower = new Fir(tapsLowerIQ.data(), (int)tapsLowerIQ.size());
    upper = new Fir(tapsUpperIQ.data(), (int)tapsUpperIQ.size());
  }

  inline void process1(float inI, float inQ, float& outI, float& outQ) {  // :102-129
    float c = cr_cosf(phase), s = cr_sinf(phase);
    float a = inI * c, b = inQ * s; outI = a - b;
    float d = inI * s, e = inQ * c; outQ = d + e;
    float upI, upQ, loI, loQ;
    upper->filter1(outI, outQ, upI, upQ);
    lower->filter1(outI, outQ, loI, loQ);
    float pu1 = upI * upI, pu2 = upQ * upQ; float powUpper = pu1 + pu2;
    float pl1 = loI * loI, pl2 = loQ * loQ; float powLower = pl1 + pl2;
    float error = powLower - powUpper;
    float be = beta * error;
    freq += be;                                          // :124
    float ae = alpha * error;
    float inc = freq + ae;
    phase += inc;                                        // :125
    if (phase > kTwoPiF || phase < -kTwoPiF)             // :185-189
      phase = remainderf(phase, kTwoPiF);                // MathF.IEEERemainder
    if (freq > maxFreq) freq = maxFreq;                  // :191-195
    else if (freq < minFreq) freq = minFreq;
  }
};

// ---------------------------------------------------------------------------------------------
// a9. MuellerMuller — MS/Models/MuellerMuller.cs:17-250
// ---------------------------------------------------------------------------------------------
struct Mm {
  double samplesPerSymbol, kp, ki;
  int baseIndex = 1;         // :44
  double mu = 0.0;           // :45
  double ncoIntegral = 0.0;
  float prevSampleI = 0, prevSampleQ = 0, prevDecisionI = 0, prevDecisionQ = 0;
  bool hasPrev = false;
  std::vector<float> buf;    // logical queue [_bufStart .. _bufStart+_bufCount) flattened

  Mm(double sps, double kp_, double ki_) : samplesPerSymbol(sps), kp(kp_), ki(ki_) {}

  inline void lagrange(int n, double mu_, float& outI, float& outQ) const {  // :160-190
    const float* b = buf.data();
    int idxm1 = (n - 1) << 1, idx0 = n << 1, idx1 = idx0 + 2, idx2 = idx0 + 4;
    float xm1I = b[idxm1], xm1Q = b[idxm1 + 1];
    float x0I = b[idx0], x0Q = b[idx0 + 1];
    float x1I = b[idx1], x1Q = b[idx1 + 1];
    float x2I = b[idx2], x2Q = b[idx2 + 1];
    float t = (float)mu_;                    // :177
    float tm1 = t - 1.0f, tm2 = t - 2.0f, tp1 = t + 1.0f;
    const float sixth = 1.0f / 6.0f, half = 1.0f / 2.0f;
    float p;
    p = t * tm1;   p = p * tm2; float c_m1 = (-p) * sixth;   // :183
    p = tp1 * tm1; p = p * tm2; float c_0 = p * half;        // :184
    p = tp1 * t;   p = p * tm2; float c_1 = (-p) * half;     // :185
    p = tp1 * t;   p = p * tm1; float c_2 = p * sixth;       // :186
    float a0 = c_m1 * xm1I, a1 = c_0 * x0I, a2 = c_1 * x1I, a3 = c_2 * x2I;
    float s = a0 + a1; s = s + a2; s = s + a3; outI = s;     // :188
    float b0 = c_m1 * xm1Q, b1 = c_0 * x0Q, b2 = c_1 * x1Q, b3 = c_2 * x2Q;
    s = b0 + b1; s = s + b2; s = s + b3; outQ = s;           // :189
  }

  int process(const float* in, int64_t nFloatsIn, float* out, int64_t capFloats) {  // :52-136
    buf.insert(buf.end(), in, in + nFloatsIn);               // Append :58,200-212
    int bufCount = (int)(buf.size() >> 1);
    int outSymbols = 0;
    while (baseIndex + 2 < bufCount) {                       // :62
      float currI, currQ;
      lagrange(baseIndex, mu, currI, currQ);
      float decI = (currI >= 0.0f) ? 1.0f : -1.0f;           // :194-198
      float decQ = (currQ >= 0.0f) ? 1.0f : -1.0f;
      double advance;
      if (hasPrev) {
        double term1Real = (double)prevDecisionI * currI + (double)prevDecisionQ * currQ;  // :78
        double term2Real = (double)decI * prevSampleI + (double)decQ * prevSampleQ;        // :79
        double e = term1Real - term2Real;
        ncoIntegral += ki * e;                               // :83
        double correction = kp * e + ncoIntegral;            // :84
        const double maxStep = 0.1;
        if (correction > maxStep) correction = maxStep;
        if (correction < -maxStep) correction = -maxStep;
        advance = samplesPerSymbol + correction;
      } else {
        hasPrev = true;
        advance = samplesPerSymbol;
      }
      int64_t o = (int64_t)outSymbols << 1;
      if (o + 1 >= capFloats) break;                         // :101-102 (state already touched)
      out[o] = currI; out[o + 1] = currQ;
      outSymbols++;
      prevSampleI = currI; prevSampleQ = currQ;
      prevDecisionI = decI; prevDecisionQ = decQ;
      double newTime = baseIndex + mu + advance;             // :113
      baseIndex = (int)std::floor(newTime);
      mu = newTime - baseIndex;
      if (baseIndex + 1 >= bufCount) break;                  // :118-119
    }
    int consumed = std::min(std::max(0, baseIndex - 1), std::max(0, bufCount - 3));  // :123
    if (consumed > 0) {
      buf.erase(buf.begin(), buf.begin() + (size_t)consumed * 2);
      baseIndex -= consumed;
    }
    return outSymbols;
  }
};

// ---------------------------------------------------------------------------------------------
// a10. CostasLoopQpsk — MS/Models/CostasLoopQpsk.cs:19-131
// ---------------------------------------------------------------------------------------------
struct Costas {
  double alpha, beta, theta = 0.0, freq = 0.0;
  Costas(double sampleRate, double loopBandwidthHz, double damping) {
    double bw = 2.0 * M_PI * loopBandwidthHz / sampleRate;   // :39
    double d = 1.0 + 2.0 * damping * bw + bw * bw;           // :42
    alpha = (4.0 * damping * bw) / d;
    beta = (4.0 * bw * bw) / d;
  }
  inline void process1(float inI, float inQ, float& outI, float& outQ) {  // :63-92
    double c = std::cos(theta), s = std::sin(theta);
    double mi = (double)inI * c + (double)inQ * s;
    double mq = (double)inQ * c - (double)inI * s;
    outI = (float)mi; outQ = (float)mq;
    float estI = (outI >= 0.0f) ? 1.0f : -1.0f;
    float estQ = (outQ >= 0.0f) ? 1.0f : -1.0f;
    double phaseError = (double)estI * mq - (double)estQ * mi;
    freq += beta * phaseError;
    theta += freq + alpha * phaseError;
    const double TwoPi = 2.0 * M_PI;
    if (theta > M_PI) theta -= TwoPi;
    else if (theta < -M_PI) theta += TwoPi;
  }
};

// ---------------------------------------------------------------------------------------------
// BitPacker — MS/Models/HelperFunctions.cs:11-71
// ---------------------------------------------------------------------------------------------
std::string bytes_to_bits(const uint8_t* d, size_t n) {            // :14-29
  std::string s; s.resize(n * 8);
  size_t k = 0;
  for (size_t i = 0; i < n; i++)
    for (int bit = 7; bit >= 0; bit--) s[k++] = ((d[i] >> bit) & 1) == 0 ? '0' : '1';
  return s;
}
std::vector<uint8_t> bits_to_bytes(const std::string& bits, int bitOffset) {  // :32-57
  int64_t usable = (int64_t)bits.size() - bitOffset;
  if (usable < 8) return {};
  int64_t nBytes = usable / 8;
  std::vector<uint8_t> out((size_t)nBytes);
  size_t p = (size_t)bitOffset;
  for (int64_t i = 0; i < nBytes; i++) {
    uint8_t v = 0;
    for (int j = 0; j < 8; j++) { v <<= 1; if (bits[p++] == '1') v |= 1; }
    out[(size_t)i] = v;
  }
  return out;
}
int64_t index_of(const uint8_t* hay, int64_t nh, const uint8_t* needle, int64_t nn) {  // :59-70
  if (nn == 0) return 0;
  if (nn > nh) return -1;
  for (int64_t i = 0; i <= nh - nn; i++)
    if (std::memcmp(hay + i, needle, (size_t)nn) == 0) return i;
  return -1;
}

inline bool is_null_or_whitespace(const char* s) {
  if (!s) return true;
  for (; *s; ++s) if (!(*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\v' || *s == '\f')) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// a6. QPSKModulator — MS/QPSKModulator.cs:18-168
// ---------------------------------------------------------------------------------------------
struct Mod {
  int sampleRate, symbolRate;
  bool diff;
  bool hasTsc;
  std::string tsc;
  std::vector<double> rrcCoeff;
  Fir* rrcTx;
  Mod(int fs, int rs, double alpha, int span, bool diff_, const char* tsc_)
      : sampleRate(fs), symbolRate(rs), diff(diff_) {
    hasTsc = !is_null_or_whitespace(tsc_);                 // :27
    if (hasTsc) tsc = tsc_;
    rrcCoeff = rrc_taps(span, alpha, fs, rs);              // :29-30
    auto iq = real_taps_to_iq(rrcCoeff);                   // :39-41
    rrcTx = new Fir(iq.data(), (int)iq.size());
  }
  ~Mod() { delete rrcTx; }

  // Modulate(string, bool) — :104-167.  Returns status; fills `out`.
  int modulate(const std::string& bitsIn, bool pulseShaping, std::vector<float>& out) const {
    std::string data = hasTsc ? (tsc + bitsIn) : bitsIn;   // :109
    int64_t nDibits = (int64_t)data.size() >> 1;           // :112
    out.clear();
    if (nDibits == 0) return ORC_OK;                       // :113
    int sps = sampleRate / symbolRate;                     // :115 integer division
    if (sps <= 0) return ORC_ERR_RANGE;                    // :116-117
    int delay = ((int)rrcCoeff.size() - 1) / 2;            // :119
    int64_t baseComplex = delay + nDibits * sps;           // :120
    int64_t totalComplex = pulseShaping ? (baseComplex + delay) : baseComplex;  // :121
    std::vector<float> up((size_t)totalComplex << 1, 0.0f);
    const float InvSqrt2 = 0.7071067811865475f;            // :36
    float prevI = InvSqrt2, prevQ = InvSqrt2;              // :126
    int64_t writeComplex = delay;
    for (int64_t d = 0; d < nDibits; d++) {
      int bi = data[(size_t)(d << 1)] - '0';
      int bq = data[(size_t)(d << 1) + 1] - '0';
      float symI, symQ;
      if (diff) {
        float dI, dQ;                                      // DibitToDelta :92-102
        if (bi == 0 && bq == 0) { dI = 1.0f; dQ = 0.0f; }
        else if (bi == 0 && bq == 1) { dI = 0.0f; dQ = 1.0f; }
        else if (bi == 1 && bq == 1) { dI = -1.0f; dQ = 0.0f; }
        else { dI = 0.0f; dQ = -1.0f; }
        float a = prevI * dI, b = prevQ * dQ; symI = a - b;     // :142
        float c = prevI * dQ, e = prevQ * dI; symQ = c + e;     // :143
        prevI = symI; prevQ = symQ;
      } else {
        symI = (bi == 0 ? -InvSqrt2 : InvSqrt2);           // :150-151
        symQ = (bq == 0 ? -InvSqrt2 : InvSqrt2);
      }
      size_t w = (size_t)writeComplex << 1;
      up[w] = symI; up[w + 1] = symQ;
      writeComplex += sps;
    }
    if (!pulseShaping) { out.swap(up); return ORC_OK; }    // :162-163
    out.resize(up.size());
    rrcTx->fft_filter(up.data(), out.data(), (int64_t)up.size());   // :166
    return ORC_OK;
  }
};

// ---------------------------------------------------------------------------------------------
// a11-a12. QPSKDeModulator — MS/QPSKDeModulator.cs:11-456
// ---------------------------------------------------------------------------------------------
struct Demod {
  bool diff, hasTsc, useFll;
  std::string tsc;
  Fir* rrc;
  Fll* fll;
  Mm* mm;
  Costas* costas;
  // framer state (:58-73)
  std::vector<uint8_t> ring;      // payload bytes of the current frame (logical ring content)
  int64_t ringCapacity;           // 300_000_000 in the reference (:58); configurable for tests
  bool inFrame = false;
  std::string searchCarryBits;
  int lockedBitOffset = -1;
  uint8_t packByte = 0;
  int packBits = 0;
  bool diffHavePrev = false;
  float prevDecI = 0.0f, prevDecQ = 0.0f;
  std::vector<float> tmpFll, tmpRrc, tmpSym;

  Demod(int fs, int rs, float alpha, int span, double symBw, double costasBw, double cfoBw,
        bool diff_, const char* tsc_, bool useFll_, int64_t ringCap)
      : diff(diff_), useFll(useFll_), ringCapacity(ringCap) {
    hasTsc = !is_null_or_whitespace(tsc_);                 // :21
    if (hasTsc) tsc = tsc_;
    auto h = rrc_taps(span, (double)alpha, fs, rs);        // :28-32 (float alpha widened)
    auto iq = real_taps_to_iq(h);
    rrc = new Fir(iq.data(), (int)iq.size());
    fll = new Fll((float)(fs / rs), alpha, 40, (float)cfoBw);   // :35 integer division
    // setupSymbolSync :39-55
    double zeta = 1.0 / std::sqrt(2.0);
    double Bn = symBw;
    double wn = ((2.0 * M_PI * Bn) / (zeta + 0.25) / zeta);
    double denom = 1.0 + 2.0 * zeta * wn + wn * wn;
    double kp = (4.0 * zeta * wn) / denom;
    double ki = (4.0 * wn * wn) / denom;
    mm = new Mm((double)fs / (double)rs, kp, ki);
    costas = new Costas((double)rs, (double)rs / costasBw, 0.707);   // :56
  }
  ~Demod() { delete rrc; delete fll; delete mm; delete costas; }

  // front end shared by DeModulate / deModulateConstellation (:355-367, :433-443)
  int front(const float* in, int64_t nFloats) {
    tmpRrc.resize((size_t)nFloats); tmpSym.resize((size_t)nFloats);
    const float* src = in;
    if (useFll) {   // the call at :359 / :435, commented out in the shipped source
      tmpFll.resize((size_t)nFloats);
      for (int64_t s = 0; s < nFloats; s += 2)
        fll->process1(in[s], in[s + 1], tmpFll[(size_t)s], tmpFll[(size_t)s + 1]);
      src = tmpFll.data();
    }
    rrc->filter(src, tmpRrc.data(), nFloats);
    return mm->process(tmpRrc.data(), nFloats, tmpSym.data(), nFloats);
  }

  static void append_decision_bits(std::string& sb, float di, float dq) {   // :304-318
    if (di < 0.0f) { if (dq < 0.0f) sb += "00"; else sb += "01"; }
    else { if (dq >= 0.0f) sb += "11"; else sb += "10"; }
  }
  static void append_delta_bits(std::string& sb, float deltaI, float deltaQ) {  // :320-337
    float ar = std::fabs(deltaI), aq = std::fabs(deltaQ);
    if (ar >= aq) { if (deltaI >= 0.0f) sb += "00"; else sb += "11"; }
    else { if (deltaQ >= 0.0f) sb += "01"; else sb += "10"; }
  }

  // DeModulate(ReadOnlySpan<float>) — :345-425
  std::string demodulate(const float* in, int64_t nFloats) {
    if (nFloats == 0) return "";
    int nSymbols = front(in, nFloats);
    std::string bits; bits.reserve((size_t)nSymbols * 2);
    for (int k = 0; k < nSymbols; k++) {
      float rotI, rotQ;
      costas->process1(tmpSym[(size_t)k * 2], tmpSym[(size_t)k * 2 + 1], rotI, rotQ);
      float decI = (rotI >= 0.0f) ? 1.0f : -1.0f, decQ = (rotQ >= 0.0f) ? 1.0f : -1.0f;
      if (diff) {
        if (!diffHavePrev) { prevDecI = decI; prevDecQ = decQ; diffHavePrev = true; continue; }
        float a = decI * prevDecI, b = decQ * prevDecQ; float deltaI = a + b;   // :397
        float c = decQ * prevDecI, d = decI * prevDecQ; float deltaQ = c - d;   // :398
        prevDecI = decI; prevDecQ = decQ;
        append_delta_bits(bits, deltaI, deltaQ);
      } else {
        append_decision_bits(bits, decI, decQ);
      }
    }
    if (hasTsc) {                                          // :413-422
      size_t idx = bits.find(tsc);
      if (idx == std::string::npos) return "";
      size_t start = idx + tsc.size();
      if (start > bits.size()) return "";
      return bits.substr(start);
    }
    return bits;
  }

  // deModulateConstellation — :427-455
  int constellation(const float* in, int64_t nFloats, std::vector<float>& y) {
    int nSymbols = front(in, nFloats);
    y.resize((size_t)nSymbols * 2);
    for (int k = 0; k < nSymbols; k++)
      costas->process1(tmpSym[(size_t)k * 2], tmpSym[(size_t)k * 2 + 1], y[(size_t)k * 2], y[(size_t)k * 2 + 1]);
    return nSymbols;
  }

  void reset_framer() {                                    // :159-167
    inFrame = false; lockedBitOffset = -1; searchCarryBits.clear();
    ring.clear(); packByte = 0; packBits = 0;
  }
  int64_t append_bits_to_ring(const std::string& bits) {   // :108-129
    int64_t produced = 0;
    for (char c : bits) {
      packByte = (uint8_t)((packByte << 1) | (c == '1' ? 1 : 0));
      packBits++;
      if (packBits == 8) {
        if ((int64_t)ring.size() >= ringCapacity) return -1;   // RingTryWriteByte :96-104
        ring.push_back(packByte);
        produced++; packBits = 0; packByte = 0;
      }
    }
    return produced;
  }
  int64_t ring_index_of(const uint8_t* pat, int64_t np, int64_t from) const {   // :133-149
    if (np == 0) return 0;
    int64_t cnt = (int64_t)ring.size();
    if (cnt < np) return -1;
    int64_t last = cnt - np;
    for (int64_t i = std::max<int64_t>(0, from); i <= last; i++)
      if (std::memcmp(ring.data() + i, pat, (size_t)np) == 0) return i;
    return -1;
  }

  // DeModulateBytes — :169-259.  Returns status; payload in `payload`.
  int demodulate_bytes(const float* in, int64_t nFloats, const uint8_t* sm, int64_t ns,
                       const uint8_t* em, int64_t ne, std::vector<uint8_t>& payload) {
    payload.clear();
    if (ns == 0 || ne == 0) return ORC_ERR_ARG;            // :174-175
    std::string rxBits = demodulate(in, nFloats);
    return frame_bits(rxBits, sm, ns, em, ne, payload);
  }

  // the rest of DeModulateBytes after its DeModulate call — :179-259 (test hook: the framer on given bits)
  int frame_bits(const std::string& rxBits, const uint8_t* sm, int64_t ns, const uint8_t* em, int64_t ne,
                 std::vector<uint8_t>& payload) {
    payload.clear();
    if (ns == 0 || ne == 0) return ORC_ERR_ARG;            // :174-175
    if (rxBits.empty()) return ORC_OK;                     // :179-180
    if (!inFrame) {
      std::string cand = searchCarryBits + rxBits;         // :185
      for (int bitOffset = 0; bitOffset < 8; bitOffset++) {
        auto bytes = bits_to_bytes(cand, bitOffset);
        if (bytes.empty()) continue;
        int64_t s = index_of(bytes.data(), (int64_t)bytes.size(), sm, ns);
        if (s < 0) continue;
        int64_t markerEndBitPos = bitOffset + 8 * (s + ns);           // :198
        if ((uint64_t)markerEndBitPos > (uint64_t)cand.size()) continue;
        inFrame = true; lockedBitOffset = bitOffset;
        ring.clear(); packByte = 0; packBits = 0;
        std::string after = cand.substr((size_t)markerEndBitPos);
        int64_t appended = append_bits_to_ring(after);
        if (appended < 0) { reset_framer(); return ORC_OK; }
        int64_t endAt = ring_index_of(em, ne, std::max<int64_t>(0, (int64_t)ring.size() - (appended + ne)));
        if (endAt >= 0) { payload.assign(ring.begin(), ring.begin() + endAt); reset_framer(); return ORC_OK; }
        return ORC_OK;
      }
      size_t keepBits = std::min(cand.size(), (size_t)(ns * 8 + 7));   // :233
      searchCarryBits = keepBits == 0 ? "" : cand.substr(cand.size() - keepBits, keepBits);
      return ORC_OK;
    }
    {
      int64_t appended = append_bits_to_ring(rxBits);      // :240
      if (appended < 0) { reset_framer(); return ORC_OK; }
      int64_t scanFrom = std::max<int64_t>(0, (int64_t)ring.size() - (appended + ne));
      int64_t endAt = ring_index_of(em, ne, scanFrom);
      if (endAt >= 0) { payload.assign(ring.begin(), ring.begin() + endAt); reset_framer(); return ORC_OK; }
      return ORC_OK;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// NCO — TB/Simulated/LocalOscilator.cs:5-194 (System.Random -> counter RNG)
// ---------------------------------------------------------------------------------------------
struct Nco {
  double baseFrequencyHz, sampleRateHz, ppm;
  double currentFrequencyHz = 0.0, phase;
  double maxPpmError, staticPpmError = 0.0, driftPpm = 0.0, currentTotalPpm = 0.0;
  int driftUpdateIntervalSamples, driftCounter = 0;
  uint64_t seed, stream, counter = 0;
  double next_double() { return rng_double(seed, stream, counter++); }

  Nco(double f, double fs, double ppm_, double phase0, uint64_t seed_, uint64_t stream_)
      : baseFrequencyHz(f), sampleRateHz(fs), ppm(ppm_), phase(phase0), seed(seed_), stream(stream_) {
    maxPpmError = std::fabs(ppm);                                              // :56
    driftUpdateIntervalSamples = (int)std::max(1.0, sampleRateHz * 1e-3);      // :59
    if (maxPpmError > 0.0) staticPpmError = (next_double() * 2.0 - 1.0) * maxPpmError;  // :124-128
    else staticPpmError = 0.0;
    driftPpm = 0.0; currentTotalPpm = staticPpmError;
    update_current();
    wrap();
  }
  void update_current() { currentFrequencyHz = baseFrequencyHz * (1.0 + currentTotalPpm * 1e-6); }  // :181-186
  void wrap() {                                                                // :188-193
    const double twoPi = 2.0 * M_PI;
    phase = std::fmod(phase, twoPi);
    if (phase < 0) phase += twoPi;
  }
  void next(double& re, double& im) {                                          // :69-79
    if (maxPpmError <= 0.0) {                                                  // :144-149
      currentFrequencyHz = baseFrequencyHz;
    } else {
      driftCounter++;
      if (driftCounter >= driftUpdateIntervalSamples) {
        driftCounter = 0;
        double stepStdPpm = maxPpmError * 0.001;
        double stepPpm = (next_double() * 2.0 - 1.0) * stepStdPpm;
        driftPpm += stepPpm;
        currentTotalPpm = staticPpmError + driftPpm;
        if (currentTotalPpm > maxPpmError) { currentTotalPpm = maxPpmError; driftPpm = currentTotalPpm - staticPpmError; }
        else if (currentTotalPpm < -maxPpmError) { currentTotalPpm = -maxPpmError; driftPpm = currentTotalPpm - staticPpmError; }
        update_current();
      }
    }
    double inc = 2.0 * M_PI * currentFrequencyHz / sampleRateHz;               // :73
    phase += inc;
    wrap();
    re = std::cos(phase); im = std::sin(phase);
  }
};

}  // namespace

// =============================================================================================
// C interface (ctypes)
// =============================================================================================

ORC_API int orc_simd_lanes() { return kLanes; }
ORC_API int orc_built_with_avx2() {
#if defined(__AVX2__)
  return 1;
#else
  return 0;
#endif
}

ORC_API uint64_t orc_rng_u64(uint64_t seed, uint64_t stream, uint64_t counter) { return rng_u64(seed, stream, counter); }
ORC_API double orc_rng_double(uint64_t seed, uint64_t stream, uint64_t counter) { return rng_double(seed, stream, counter); }
// uniform(-1,1) float fill used for the FIR sweep inputs: value k = (float)(2*u-1), counter = k
ORC_API void orc_fill_uniform(uint64_t seed, uint64_t stream, int64_t first, int64_t n, float* out) {
  for (int64_t k = 0; k < n; k++) out[k] = (float)(rng_double(seed, stream, (uint64_t)(first + k)) * 2.0 - 1.0);
}

// payload generator: byte k of stream s = top byte of rng(seed, s, k)
ORC_API void orc_fill_bytes(uint64_t seed, uint64_t stream, int64_t first, int64_t n, uint8_t* out) {
  for (int64_t k = 0; k < n; k++) out[k] = (uint8_t)(rng_u64(seed, stream, (uint64_t)(first + k)) >> 56);
}

ORC_API int orc_rrc_taps(double span, double beta, int fs, int rs, double* out, int cap, int* n) {
  if (!n) return ORC_ERR_NULL;
  auto h = rrc_taps(span, beta, fs, rs);
  *n = (int)h.size();
  if (!out) return ORC_OK;
  if (cap < (int)h.size()) return ORC_ERR_CAPACITY;
  std::memcpy(out, h.data(), h.size() * sizeof(double));
  return ORC_OK;
}

// ---- FIR ----
ORC_API int orc_fir_create(const float* taps_iq, int n_floats, void** out) {
  if (!taps_iq || !out) return ORC_ERR_NULL;                 // FIRFilter.cs:31
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // :32
  if (n_floats == 0) return ORC_ERR_ARG;                     // :33
  *out = new Fir(taps_iq, n_floats);
  return ORC_OK;
}
ORC_API void orc_fir_destroy(void* h) { delete (Fir*)h; }
ORC_API void orc_fir_reset(void* h) { ((Fir*)h)->reset(); }
ORC_API int orc_fir_filter(void* h, const float* in, float* out, int64_t n_floats, int64_t out_cap_floats) {
  if (!h || (!in && n_floats) || (!out && n_floats)) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // :82
  if (out_cap_floats < n_floats) return ORC_ERR_ARG;         // :83
  ((Fir*)h)->filter(in, out, n_floats);
  return ORC_OK;
}
ORC_API int orc_fir_fft_filter(void* h, const float* in, float* out, int64_t n_floats) {
  if (!h || !in) return ORC_ERR_NULL;                        // :98
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // :99
  if (n_floats == 0) return ORC_OK;                          // :100
  if (!out) return ORC_ERR_NULL;
  ((Fir*)h)->fft_filter(in, out, n_floats);
  return ORC_OK;
}
// HelperFunctions.Convolve (MS/Models/HelperFunctions.cs:113-132): complex fp64 (*) real fp64,
// full length na+nb-1.  The "reference Complex path" yardstick.
ORC_API void orc_convolve_f64(const double* a_iq, int64_t na, const double* b, int64_t nb, double* out_iq) {
  int64_t M = na + nb - 1;
  for (int64_t n = 0; n < M; n++) {
    double re = 0.0, im = 0.0;
    int64_t kmin = std::max<int64_t>(0, n - (na - 1));
    int64_t kmax = std::min<int64_t>(n, nb - 1);
    for (int64_t k = kmin; k <= kmax; k++) { re += a_iq[2 * (n - k)] * b[k]; im += a_iq[2 * (n - k) + 1] * b[k]; }
    out_iq[2 * n] = re; out_iq[2 * n + 1] = im;
  }
}
// fp64 streaming FIR with complex fp32 taps (zero initial state): error yardstick for a4.
ORC_API void orc_fir_filter_f64(const float* taps_iq, int n_taps, const float* in, double* out, int64_t n_complex) {
  for (int64_t n = 0; n < n_complex; n++) {
    double re = 0.0, im = 0.0;
    int64_t kmax = std::min<int64_t>(n, n_taps - 1);
    for (int64_t k = 0; k <= kmax; k++) {
      double hr = taps_iq[2 * k], hi = taps_iq[2 * k + 1];
      double xr = in[2 * (n - k)], xi = in[2 * (n - k) + 1];
      re += hr * xr - hi * xi; im += hr * xi + hi * xr;
    }
    out[2 * n] = re; out[2 * n + 1] = im;
  }
}

// CPU baseline: `threads` independent streams, each an identical-taps Fir over its own slice of
// [in, in + threads*n_floats_each); one std::thread per stream (SURVEY §8d "(ii)").
ORC_API int orc_fir_filter_mt(const float* taps_iq, int n_taps_floats, const float* in, float* out,
                              int64_t n_floats_each, int threads) {
  if (!taps_iq || !in || !out) return ORC_ERR_NULL;
  if ((n_taps_floats & 1) || n_taps_floats == 0 || (n_floats_each & 1)) return ORC_ERR_ARG;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([=]() {
      Fir f(taps_iq, n_taps_floats);
      f.filter(in + (size_t)t * n_floats_each, out + (size_t)t * n_floats_each, n_floats_each);
    });
  for (auto& x : th) x.join();
  return ORC_OK;
}

// ---- FLL ----
ORC_API int orc_fll_create(float sps, float rolloff, int filter_size, float bw, void** out) {
  if (!out) return ORC_ERR_NULL;
  if (!(sps > 0.0f)) return ORC_ERR_RANGE;                   // Band-Edge Filter.cs:42
  if (rolloff < 0 || rolloff > 1.0f) return ORC_ERR_RANGE;   // :43
  if (filter_size <= 0) return ORC_ERR_RANGE;                // :44
  if (!(bw > 0.0f)) return ORC_ERR_RANGE;                    // :45
  *out = new Fll(sps, rolloff, filter_size, bw);
  return ORC_OK;
}
ORC_API void orc_fll_destroy(void* h) { delete (Fll*)h; }
ORC_API int orc_fll_taps(void* h, float* lower_iq, float* upper_iq) {
  Fll* f = (Fll*)h;
  std::memcpy(lower_iq, f->tapsLowerIQ.data(), f->tapsLowerIQ.size() * sizeof(float));
  std::memcpy(upper_iq, f->tapsUpperIQ.data(), f->tapsUpperIQ.size() * sizeof(float));
  return ORC_OK;
}
ORC_API int orc_fll_process(void* h, const float* in, float* out, int64_t n_floats, int64_t out_cap_floats) {
  if (!h) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // :66-67
  if (out_cap_floats < n_floats) return ORC_ERR_ARG;         // :68-69
  Fll* f = (Fll*)h;
  for (int64_t s = 0; s < n_floats; s += 2) f->process1(in[s], in[s + 1], out[s], out[s + 1]);
  return ORC_OK;
}
ORC_API void orc_fll_get_state(void* h, float* phase, float* freq) { *phase = ((Fll*)h)->phase; *freq = ((Fll*)h)->freq; }
ORC_API void orc_fll_set_state(void* h, float phase, float freq) { ((Fll*)h)->phase = phase; ((Fll*)h)->freq = freq; }

// ---- Mueller-Muller ----
ORC_API int orc_mm_create(double sps, double kp, double ki, void** out) {
  if (!out) return ORC_ERR_NULL;
  *out = new Mm(sps, kp, ki);
  return ORC_OK;
}
ORC_API void orc_mm_destroy(void* h) { delete (Mm*)h; }
ORC_API int orc_mm_process(void* h, const float* in, int64_t n_floats, float* out, int64_t cap_floats, int* n_sym) {
  if (!h || !n_sym) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // MuellerMuller.cs:54-55
  *n_sym = ((Mm*)h)->process(in, n_floats, out, cap_floats);
  return ORC_OK;
}
ORC_API void orc_mm_get_state(void* h, int* baseIndex, double* mu, double* integ, int* queued) {
  Mm* m = (Mm*)h;
  *baseIndex = m->baseIndex; *mu = m->mu; *integ = m->ncoIntegral; *queued = (int)(m->buf.size() >> 1);
}
// setupSymbolSync gains — MS/QPSKDeModulator.cs:39-55
ORC_API void orc_mm_gains_from_bw(double symBw, double* kp, double* ki) {
  double zeta = 1.0 / std::sqrt(2.0);
  double wn = ((2.0 * M_PI * symBw) / (zeta + 0.25) / zeta);
  double denom = 1.0 + 2.0 * zeta * wn + wn * wn;
  *kp = (4.0 * zeta * wn) / denom;
  *ki = (4.0 * wn * wn) / denom;
}

// ---- Costas ----
ORC_API int orc_costas_create(double fs, double bw_hz, double damping, void** out) {
  if (!out) return ORC_ERR_NULL;
  *out = new Costas(fs, bw_hz, damping);
  return ORC_OK;
}
ORC_API void orc_costas_destroy(void* h) { delete (Costas*)h; }
ORC_API int orc_costas_process(void* h, const float* in, float* out, int64_t n_floats, int64_t out_cap_floats) {
  if (!h) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;               // CostasLoopQpsk.cs:100-101
  if (out_cap_floats < n_floats) return ORC_ERR_ARG;         // :102-103
  Costas* c = (Costas*)h;
  for (int64_t s = 0; s < n_floats; s += 2) c->process1(in[s], in[s + 1], out[s], out[s + 1]);
  return ORC_OK;
}
ORC_API void orc_costas_get_state(void* h, double* theta, double* freq) { *theta = ((Costas*)h)->theta; *freq = ((Costas*)h)->freq; }
ORC_API void orc_costas_gains(void* h, double* alpha, double* beta) { *alpha = ((Costas*)h)->alpha; *beta = ((Costas*)h)->beta; }

// ---- BitPacker ----
ORC_API void orc_bytes_to_bits(const uint8_t* d, int64_t n, char* out) {
  auto s = bytes_to_bits(d, (size_t)n);
  std::memcpy(out, s.data(), s.size());
}
ORC_API int64_t orc_bits_to_bytes(const char* bits, int64_t nbits, int bit_offset, uint8_t* out, int64_t cap) {
  if (!bits) return ORC_ERR_NULL;
  if ((unsigned)bit_offset > 7u) return ORC_ERR_RANGE;
  auto v = bits_to_bytes(std::string(bits, (size_t)nbits), bit_offset);
  if ((int64_t)v.size() > cap) return ORC_ERR_CAPACITY;
  if (!v.empty()) std::memcpy(out, v.data(), v.size());
  return (int64_t)v.size();
}
ORC_API int64_t orc_index_of(const uint8_t* hay, int64_t nh, const uint8_t* needle, int64_t nn) {
  return index_of(hay, nh, needle, nn);
}

// ---- Modulator ----
ORC_API int orc_mod_create(int fs, int rs, double alpha, int span, int diff, const char* tsc, void** out) {
  if (!out) return ORC_ERR_NULL;
  *out = new Mod(fs, rs, alpha, span, diff != 0, tsc);
  return ORC_OK;
}
ORC_API void orc_mod_destroy(void* h) { delete (Mod*)h; }
ORC_API int orc_mod_taps(void* h, double* out, int cap, int* n) {
  Mod* m = (Mod*)h;
  *n = (int)m->rrcCoeff.size();
  if (out) { if (cap < *n) return ORC_ERR_CAPACITY; std::memcpy(out, m->rrcCoeff.data(), sizeof(double) * (size_t)*n); }
  return ORC_OK;
}
ORC_API int orc_mod_modulate_bits(void* h, const char* bits, int64_t nbits, int pulse, float* out, int64_t cap_floats, int64_t* n_floats) {
  if (!h || !bits || !n_floats) return ORC_ERR_NULL;          // QPSKModulator.cs:106
  std::vector<float> y;
  int st = ((Mod*)h)->modulate(std::string(bits, (size_t)nbits), pulse != 0, y);
  if (st != ORC_OK) return st;
  *n_floats = (int64_t)y.size();
  if (!out) return ORC_OK;
  if (cap_floats < (int64_t)y.size()) return ORC_ERR_CAPACITY;
  if (!y.empty()) std::memcpy(out, y.data(), y.size() * sizeof(float));
  return ORC_OK;
}
ORC_API int orc_mod_modulate_bytes(void* h, const uint8_t* payload, int64_t np, const uint8_t* sm, int64_t ns,
                                   const uint8_t* em, int64_t ne, int pulse, float* out, int64_t cap_floats, int64_t* n_floats) {
  if (!h || !n_floats) return ORC_ERR_NULL;
  if (ns == 0 || ne == 0) return ORC_ERR_ARG;                 // :60-61
  std::vector<uint8_t> framed;                                // :64-67
  framed.insert(framed.end(), sm, sm + ns);
  framed.insert(framed.end(), payload, payload + np);
  framed.insert(framed.end(), em, em + ne);
  std::string bits = bytes_to_bits(framed.data(), framed.size());   // :70
  return orc_mod_modulate_bits(h, bits.data(), (int64_t)bits.size(), pulse, out, cap_floats, n_floats);
}

// ---- Demodulator ----
ORC_API int orc_demod_create(int fs, int rs, float alpha, int span, double sym_bw, double costas_bw, double cfo_bw,
                             int diff, const char* tsc, int use_fll, int64_t ring_capacity, void** out) {
  if (!out) return ORC_ERR_NULL;
  if (ring_capacity <= 0) ring_capacity = 300000000;          // QPSKDeModulator.cs:58
  if (rs == 0) return ORC_ERR_RANGE;                          // DivideByZeroException in the field initialisers
  {  // the FLLBandEdgeFilter field initialiser (:35) throws for bad parameters (Band-Edge Filter.cs:42-45)
    float fsps = (float)(fs / rs), fbw = (float)cfo_bw;
    if (!(fsps > 0.0f) || alpha < 0 || alpha > 1.0f || !(fbw > 0.0f)) return ORC_ERR_RANGE;
  }
  *out = new Demod(fs, rs, alpha, span, sym_bw, costas_bw, cfo_bw, diff != 0, tsc, use_fll != 0, ring_capacity);
  return ORC_OK;
}
ORC_API void orc_demod_destroy(void* h) { delete (Demod*)h; }
ORC_API int orc_demod_bits(void* h, const float* in, int64_t n_floats, char* out, int64_t cap, int64_t* n_bits) {
  if (!h || !n_bits) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;                // :347-348
  std::string s = ((Demod*)h)->demodulate(in, n_floats);
  *n_bits = (int64_t)s.size();
  if ((int64_t)s.size() > cap) return ORC_ERR_CAPACITY;
  if (!s.empty()) std::memcpy(out, s.data(), s.size());
  return ORC_OK;
}
ORC_API int orc_demod_bytes(void* h, const float* in, int64_t n_floats, const uint8_t* sm, int64_t ns,
                            const uint8_t* em, int64_t ne, uint8_t* out, int64_t cap, int64_t* n_bytes) {
  if (!h || !n_bytes) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;
  std::vector<uint8_t> p;
  int st = ((Demod*)h)->demodulate_bytes(in, n_floats, sm, ns, em, ne, p);
  if (st != ORC_OK) return st;
  *n_bytes = (int64_t)p.size();
  if ((int64_t)p.size() > cap) return ORC_ERR_CAPACITY;
  if (!p.empty()) std::memcpy(out, p.data(), p.size());
  return ORC_OK;
}
ORC_API int orc_demod_frame_bits(void* h, const char* bits, int64_t n_bits, const uint8_t* sm, int64_t ns,
                                 const uint8_t* em, int64_t ne, uint8_t* out, int64_t cap, int64_t* n_bytes) {
  if (!h || !n_bytes) return ORC_ERR_NULL;
  std::vector<uint8_t> p;
  int st = ((Demod*)h)->frame_bits(std::string(bits ? bits : "", (size_t)n_bits), sm, ns, em, ne, p);
  if (st != ORC_OK) return st;
  *n_bytes = (int64_t)p.size();
  if ((int64_t)p.size() > cap) return ORC_ERR_CAPACITY;
  if (!p.empty()) std::memcpy(out, p.data(), p.size());
  return ORC_OK;
}
ORC_API int orc_demod_constellation(void* h, const float* in, int64_t n_floats, float* out, int64_t cap_floats, int64_t* n_sym) {
  if (!h || !n_sym) return ORC_ERR_NULL;
  if ((n_floats & 1) != 0) return ORC_ERR_ARG;                // :430
  std::vector<float> y;
  *n_sym = ((Demod*)h)->constellation(in, n_floats, y);
  if ((int64_t)y.size() > cap_floats) return ORC_ERR_CAPACITY;
  if (!y.empty()) std::memcpy(out, y.data(), y.size() * sizeof(float));
  return ORC_OK;
}
ORC_API void orc_demod_loop_state(void* h, double* costas_theta, double* costas_freq, double* mm_mu, double* mm_integ,
                                  float* fll_phase, float* fll_freq) {
  Demod* d = (Demod*)h;
  *costas_theta = d->costas->theta; *costas_freq = d->costas->freq;
  *mm_mu = d->mm->mu; *mm_integ = d->mm->ncoIntegral;
  *fll_phase = d->fll->phase; *fll_freq = d->fll->freq;
}
ORC_API int orc_demod_in_frame(void* h) { return ((Demod*)h)->inFrame ? 1 : 0; }

// ---- NCO / noise / channel ----
ORC_API int orc_nco_create(double f, double fs, double ppm, double phase0, uint64_t seed, uint64_t stream, void** out) {
  if (!out) return ORC_ERR_NULL;
  if (!(fs > 0)) return ORC_ERR_RANGE;                        // LocalOscilator.cs:48-49
  *out = new Nco(f, fs, ppm, phase0, seed, stream);
  return ORC_OK;
}
ORC_API void orc_nco_destroy(void* h) { delete (Nco*)h; }
ORC_API void orc_nco_generate(void* h, double* out_iq, int64_t n) {
  Nco* o = (Nco*)h;
  for (int64_t i = 0; i < n; i++) o->next(out_iq[2 * i], out_iq[2 * i + 1]);
}
// NoiseGenerator.GenerateIqNoise — TB/HelperModels.cs:17-45; u1 = rng(counter 2n), u2 = rng(counter 2n+1)
ORC_API void orc_noise_iq(float dbfs, int64_t count, uint64_t seed, uint64_t stream, uint64_t first_sample, float* out_iq) {
  float linearRms = (float)std::pow(10.0, dbfs / 20.0);
  for (int64_t n = 0; n < count; n++) {
    uint64_t c = (first_sample + (uint64_t)n) * 2;
    double u1 = 1.0 - rng_double(seed, stream, c);
    double u2 = 1.0 - rng_double(seed, stream, c + 1);
    double mag = std::sqrt(-2.0 * std::log(u1)) * linearRms;
    double ph = 2.0 * M_PI * u2;
    float i = (float)(mag * std::cos(ph));
    float q = (float)(mag * std::sin(ph));
    i = std::min(std::max(i, -1.0f), 1.0f);
    q = std::min(std::max(q, -1.0f), 1.0f);
    out_iq[2 * n] = i; out_iq[2 * n + 1] = q;
  }
}
// Two-unstable-LO channel.
//   mode 0 (TB/Simulated/testAtDataLevel.cs:37-44): y = x * (tx.Next() * conj(rx.Next()))
//   mode 1 (TB/Simulated/testFullDemodChain.cs:73):  y = ((x + noise) * tx.Next()) * conj(rx.Next())
// x is fp32 interleaved, arithmetic in Complex (fp64), result cast to fp32.  noise may be NULL.
ORC_API void orc_channel_apply(void* tx_nco, void* rx_nco, int mode, const float* x_iq, const float* noise_iq,
                               int64_t n, float* y_iq) {
  Nco* tx = (Nco*)tx_nco; Nco* rx = (Nco*)rx_nco;
  for (int64_t i = 0; i < n; i++) {
    double tr, ti, rr, ri;
    tx->next(tr, ti);
    rx->next(rr, ri);
    ri = -ri;  // Conjugate()
    double xr = x_iq[2 * i], xi = x_iq[2 * i + 1];
    double yr, yi;
    if (mode == 0) {
      double pr = tr * rr - ti * ri, pi = tr * ri + ti * rr;
      yr = xr * pr - xi * pi; yi = xr * pi + xi * pr;
    } else {
      if (noise_iq) { xr = xr + (double)noise_iq[2 * i]; xi = xi + (double)noise_iq[2 * i + 1]; }
      double ar = xr * tr - xi * ti, ai = xr * ti + xi * tr;
      yr = ar * rr - ai * ri; yi = ar * ri + ai * rr;
    }
    y_iq[2 * i] = (float)yr; y_iq[2 * i + 1] = (float)yi;
  }
}

// Static multipath (NOT in the reference — README.md:2 mentions it, no code implements it; the
// generator defines it, DESIGN.md "impairments"): y[n] = sum_k g[k] * x[n - d[k]], complex fp32
// gains, integer delays, fp32 arithmetic with separate roundings, zero initial state.
ORC_API void orc_multipath(const float* x_iq, int64_t n, const float* gains_iq, const int* delays, int n_paths, float* y_iq) {
  for (int64_t i = 0; i < n; i++) {
    float accR = 0.0f, accI = 0.0f;
    for (int k = 0; k < n_paths; k++) {
      int64_t m = i - delays[k];
      if (m < 0) continue;
      float gr = gains_iq[2 * k], gi = gains_iq[2 * k + 1];
      float xr = x_iq[2 * m], xi = x_iq[2 * m + 1];
      float a = gr * xr, b = gi * xi, c = gr * xi, d = gi * xr;
      float e = a - b, f = c + d;
      accR = accR + e; accI = accI + f;
    }
    y_iq[2 * i] = accR; y_iq[2 * i + 1] = accI;
  }
}
