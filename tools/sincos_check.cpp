// sincos_check.cpp — host brute force: the FLL's shortened fp64 sincos (sc<2>, csrc/common.cuh sincos_f32arg_k) is
// bit-identical in fp64 to sincos_fast_f64 (sc<0>) for every fp32 argument 2^-31 <= |x| < 64.
// build: g++ -O2 -ffp-contract=off -mfma -fopenmp tools/sincos_check.cpp -o /tmp/sincos_check
// host check of the fp64 sincos variants for fp32 arguments: compare (float)variant vs (float)glibc
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <omp.h>
static const double S[10] = {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0,
    -1.0 / 1307674368000.0, 1.0 / 355687428096000.0, -1.0 / 121645100408832000.0, 1.0 / 51090942171709440000.0};
static const double Cc_[10] = {1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0, -1.0 / 87178291200.0,
    1.0 / 20922789888000.0, -1.0 / 6402373705728000.0, 1.0 / 2432902008176640000.0, -1.0 / 1124000727777607680000.0};
template <int V>
static inline void sc(double x, double* sn, double* cs) {
  double k; int q;
  if (V == 0) { k = rint(x * 0.63661977236758134308); q = (int)k; }
  else { const double M = 6755399441055744.0; double t = fma(x, 0.63661977236758134308, M); k = t - M; int64_t b; memcpy(&b, &t, 8); q = (int)(uint32_t)b; }
  double r = fma(-k, 1.5707963267948966e+00, x);
  r = fma(-k, 6.123233995736766e-17, r);
  if (V == 0) r = fma(-k, -1.4973849048591698e-33, r);
  const double z = r * r, z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  double a01 = fma(S[1], z, S[0]), a23 = fma(S[3], z, S[2]), a45 = fma(S[5], z, S[4]), a67 = fma(S[7], z, S[6]), a89 = fma(S[9], z, S[8]);
  double d01 = fma(Cc_[1], z, Cc_[0]), d23 = fma(Cc_[3], z, Cc_[2]), d45 = fma(Cc_[5], z, Cc_[4]), d67 = fma(Cc_[7], z, Cc_[6]), d89 = fma(Cc_[9], z, Cc_[8]);
  double Sv, Cv;
  if (V < 2) {
    Sv = fma(a89, z8, fma(fma(a67, z2, a45), z4, fma(a23, z2, a01)));
    Cv = fma(d89, z8, fma(fma(d67, z2, d45), z4, fma(d23, z2, d01)));
  } else {
    // fewer terms: |r| <= pi/4 -> z <= 0.617; term z^8/(19!)... keep 8 coefficients (degree 7 in z)
    Sv = fma(fma(a67, z2, a45), z4, fma(a23, z2, a01));
    Cv = fma(fma(d67, z2, d45), z4, fma(d23, z2, d01));
  }
  const double sr = fma(r * z, Sv, r);
  const double hz = 0.5 * z, w = 1.0 - hz;
  double cr;
  if (V == 0) cr = w + (((1.0 - w) - hz) + z2 * Cv);
  else cr = w + fma(z2, Cv, (1.0 - w) - hz);
  const double s0 = (q & 1) ? cr : sr, c0 = (q & 1) ? sr : cr;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
template <int V> void run(const char* name) {
  long long bad_s = 0, bad_c = 0, tot = 0, bad64 = 0;
  #pragma omp parallel for reduction(+:bad_s,bad_c,tot,bad64) schedule(dynamic,1)
  for (int e = 0; e < 260; ++e) {   // exponent blocks: floats with |x| in [2^-126, 8): bits 0x00800000 .. 0x41000000
    uint32_t lo = 0x00800000u + (uint32_t)((0x41000000ull - 0x00800000ull) * e / 260), hi = 0x00800000u + (uint32_t)((0x41000000ull - 0x00800000ull) * (e + 1) / 260);
    for (uint32_t b = lo; b < hi; b += 1) {
      for (int sg = 0; sg < 2; ++sg) {
        uint32_t bb = b | (sg ? 0x80000000u : 0); float x; memcpy(&x, &bb, 4);
        double s, c; sc<V>((double)x, &s, &c);
        double gs = sin((double)x), gc = cos((double)x);
        bad_s += ((float)s != (float)gs); bad_c += ((float)c != (float)gc); bad64 += (s != gs) + (c != gc); tot++;
      }
    }
  }
  printf("%s: total %lld, fp32 mismatches sin %lld cos %lld, fp64 diffs %lld\n", name, tot, bad_s, bad_c, bad64);
}
// csrc/fll_duo.cu duo_wrap_phase vs remainderf, every fp32 with 2*pi < |x| < 1e5
static void check_wrap() {
  const float cf = 6.2831854820251464844f;
  const double c = cf;
  long long bad = 0, n = 0;
  #pragma omp parallel for reduction(+:bad,n) schedule(static)
  for (uint32_t b = 0x40C90FDBu; b < 0x47C35000u; ++b)
    for (int sg = 0; sg < 2; ++sg) {
      uint32_t bb = b | (sg ? 0x80000000u : 0);
      float x; memcpy(&x, &bb, 4);
      const float want = remainderf(x, cf);
      const double pd = x, nn = rint(pd * (1.0 / c)), r = fma(-nn, c, pd);
      const float got = (r == 0.0) ? copysignf(0.f, x) : (float)r;
      bad += memcmp(&want, &got, 4) != 0; n++;
    }
  printf("wrap: %lld args, mismatches vs remainderf %lld\n", n, bad);
}
int main() {
  check_wrap();
  long long ds=0,dc=0,n=0;
  #pragma omp parallel for reduction(+:ds,dc,n) schedule(dynamic,1)
  for (int blk = 0; blk < 512; ++blk) {
    const uint64_t lo0 = 0x30000000ull, hi0 = 0x42800000ull;   // 2^-31 .. 64
    uint32_t lo = (uint32_t)(lo0 + (hi0 - lo0) * blk / 512), hi = (uint32_t)(lo0 + (hi0 - lo0) * (blk + 1) / 512);
    for (uint32_t b = lo; b < hi; ++b) for (int sg = 0; sg < 2; ++sg) { uint32_t bb = b | (sg ? 0x80000000u : 0); float x; memcpy(&x,&bb,4);
      double s0,c0,s2,c2; sc<0>(x,&s0,&c0); sc<2>(x,&s2,&c2); ds += memcmp(&s0,&s2,8)!=0; dc += memcmp(&c0,&c2,8)!=0; n++; } }
  printf("n %lld V0!=V2 (bitwise fp64) sin %lld cos %lld\n", n, ds, dc);
}
