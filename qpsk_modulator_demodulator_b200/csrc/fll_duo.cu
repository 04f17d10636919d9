// fll_duo.cu — K4, the band-edge FLL (MS/Models/Band-Edge Filter.cs:102-129,185-195) for N = 8*K taps:
// two warps per four streams, the per-sample recurrence alone on one of them.
//
// Why two warps.  With a few thousand streams there is about one FLL warp per warp scheduler, and a lone warp
// issues in order: every instruction that is not on the loop-carried chain
//      phase -> sin/cos -> rotate -> newest tap term -> |lo|^2-|up|^2 -> freq -> phase
// still delays it.  So the work is split by what it depends on:
//   * CHAIN warp (8 lanes per stream = the 8 SIMD lanes of the reference's Vector<float> dot product,
//     FIRFilter.cs:165-180): sin/cos, rotation, and only the terms that involve the newest output out[n];
//   * SIDE warp (same lane <-> (stream, SIMD lane) mapping): everything that depends only on outputs at least 8
//     samples old — the first K-1 elements of every lane partial — plus all global I/O (input prefetch, output
//     flush, state load/store).  It runs one batch (4 samples) ahead and hands results over through shared memory
//     with mbarriers; its latencies never touch the chain.
//
// Order of additions (bit-identical to ComplexDotWindow, FIRFilter.cs:165-192, for N % 8 == 0):
//   window n, SIMD lane l:   L_l(n) = (((0 + e_{l}) + e_{l+8}) + ...) + e_{l+8(K-1)},  e_i = tap_rev[i] (x) out[n-(N-1)+i]
//   horizontal sum:          acc(n) = ((((((0 + L_0) + L_1) + ...) + L_6) + L_7
// The last element of lane l in window m is out[m-7+l], so at step n (out[n] just computed) GPU lane g finishes
// L_g of window m = n+7-g and extends that window's prefix  P_g(m) = P_{g-1}(m) + L_g(m)  with the value lane g-1
// produced one step earlier (one SHFL, issued before the rotation is known): the horizontal sum is a systolic
// pipeline across the 8 lanes and across time, and nothing but register arithmetic follows the rotation.  Every
// lane also plays lane 7's role for the current window (acc(n) = P_6(n) + L_7(n)), redundantly, so the loop state
// (phase, freq) stays uniform in the group without a broadcast on the chain.
#include "loops.cuh"

#include <type_traits>

namespace qpsk {

namespace {

constexpr int kDuoStreams = 4;            // streams per warp pair
constexpr int kDuoRing = 64;              // ring slots per stream (power of two, >= N + 16)
constexpr int kDuoRingStride = kDuoRing + 2;   // float2; 528 B: 16-byte aligned rows, streams 4 banks apart
constexpr int kDuoBatch = 4;              // samples per hand-over
constexpr int kDuoSuper = 16;             // samples per global-memory transaction and stream
constexpr int kDuoXStride = kDuoSuper + 2;

__device__ __forceinline__ uint32_t duo_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void duo_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(duo_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void duo_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(duo_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void duo_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(duo_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

struct DuoSmem {
  float2 ring[kDuoStreams * kDuoRingStride];            // past outputs, slot = sample index & 63
  float4 lp[2][kDuoBatch][32];                          // per batch parity: stale lane partials (loI, loQ, upI, upQ)
  float2 xq[2][kDuoStreams * kDuoXStride];              // input samples, one super-batch per slot
  uint64_t lp_full[2];                                  // side -> chain: batch's partials (and inputs) are in place
  uint64_t out_full[2];                                 // chain -> side: batch's outputs are in the ring
};

// the four sums of one window element: (loI, loQ, upI, upQ) += tap (x) v, the reference's products and order
// (Band-Edge Filter.cs:115-116 through FIRFilter.cs:165-192; upper = conj(lower) shares the four products)
__device__ __forceinline__ void duo_acc(float4& a, float ta, float tb, float vx, float vy) {
  const float p1 = ta * vx, p2 = tb * vy, p3 = ta * vy, p4 = tb * vx;
  a.x = a.x + (p1 - p2);
  a.y = a.y + (p3 + p4);
  a.z = a.z + (p1 + p2);
  a.w = a.w + (p3 - p4);
}

template <int K, int PAIRS>
__global__ void __launch_bounds__(64 * PAIRS)
    fll_duo_kernel(const FllParams P, const float* __restrict__ taps, float2* ring_g, int* head_g, float2* pf_g, int C,
                   const float2* __restrict__ x, float2* __restrict__ y, long long L, long long ldx, long long ldy) {
  constexpr int N = 8 * K;
  static_assert(N + 16 <= kDuoRing, "ring too small");
  __shared__ __align__(16) DuoSmem smem[PAIRS];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = warp >> 1;
  const bool is_chain = (warp & 1) == 0;
  DuoSmem& S = smem[pair];
  const int sl = lane >> 3;                                 // stream slot in the pair
  const int g = lane & 7;                                   // reference SIMD lane
  const int c_raw = (blockIdx.x * PAIRS + pair) * kDuoStreams + sl;
  const bool live = c_raw < C;
  const int c = live ? c_raw : C - 1;                       // idle groups shadow the last stream, no stores
  float2* myring = S.ring + sl * kDuoRingStride;
  const long long nB = (L + kDuoBatch - 1) / kDuoBatch;     // batches of real samples; batches -2, -1 are the warm-up

  if (lane == 0 && is_chain) {
    duo_mbar_init(&S.lp_full[0], 1);
    duo_mbar_init(&S.lp_full[1], 1);
    duo_mbar_init(&S.out_full[0], 1);
    duo_mbar_init(&S.out_full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (!is_chain) {
    // =========================================== SIDE warp ===========================================
    float ta[K > 1 ? K - 1 : 1], tb[K > 1 ? K - 1 : 1];
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
      ta[k] = taps[g + 8 * k];
      tb[k] = taps[N + g + 8 * k];
    }
    // history: chronological element j (oldest first) of the saved ring goes to sample index j - N
    {
      const int head = head_g[c];
      for (int j = g; j < N; j += 8) {
        int slot = head + j;
        if (slot >= N) slot -= N;
        myring[(j - N) & (kDuoRing - 1)] = ring_g[(long long)slot * C + c];
      }
    }
    __syncwarp();
    const float2* xc = x + (long long)c * ldx;
    float2* yc = y + (long long)c * ldy;
    // input prefetch registers: super-batch sb holds samples [16 sb, 16 sb + 16); lane g loads g and g + 8
    float2 xr0 = make_float2(0.f, 0.f), xr1 = make_float2(0.f, 0.f);
    if (g < L) xr0 = xc[g];
    if (g + 8 < L) xr1 = xc[g + 8];
    long long flushed = 0;
    for (long long b = -2; b < nB; ++b) {
      if (b >= 0) duo_mbar_wait(&S.out_full[b & 1], (uint32_t)((b >> 1) & 1));   // chain finished batch b-2
      if (b >= 0 && (b & 3) == 0) {
        const long long sb = b >> 2;
        float2* q = S.xq[sb & 1] + sl * kDuoXStride;
        q[g] = xr0;
        q[g + 8] = xr1;
        const long long i0 = (sb + 1) * kDuoSuper + g;
        if (i0 < L) xr0 = xc[i0];
        if (i0 + 8 < L) xr1 = xc[i0 + 8];
      }
      if ((b & 3) == 1 && b >= 5) {
        // outputs through sample 4(b-2)+3 = 16 sb - 1 are final: flush [16(sb-1), 16 sb)
        const long long f0 = ((b >> 2) - 1) * kDuoSuper;
        if (live) {
          yc[f0 + g] = myring[(int)(f0 + g) & (kDuoRing - 1)];
          yc[f0 + g + 8] = myring[(int)(f0 + g + 8) & (kDuoRing - 1)];
        }
        flushed = f0 + kDuoSuper;
      }
      // stale partials for the steps n = 4b + j: elements k = 0..K-2 of this lane's window, out[n - 8(K-1-k)]
      float4 acc[kDuoBatch];
#pragma unroll
      for (int j = 0; j < kDuoBatch; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < K - 1; ++k) {
        const int s0 = (int)(4 * b - 8 * (K - 1 - k)) & (kDuoRing - 1);        // multiple of 4
        const float4 v01 = *reinterpret_cast<const float4*>(myring + s0);
        const float4 v23 = *reinterpret_cast<const float4*>(myring + s0 + 2);
        duo_acc(acc[0], ta[k], tb[k], v01.x, v01.y);
        duo_acc(acc[1], ta[k], tb[k], v01.z, v01.w);
        duo_acc(acc[2], ta[k], tb[k], v23.x, v23.y);
        duo_acc(acc[3], ta[k], tb[k], v23.z, v23.w);
      }
#pragma unroll
      for (int j = 0; j < kDuoBatch; ++j) S.lp[b & 1][j][lane] = acc[j];
      __syncwarp();
      if (lane == 0) duo_mbar_arrive(&S.lp_full[b & 1]);
    }
    // the chain's last batch
    {
      const long long bl = nB - 1;
      duo_mbar_wait(&S.out_full[bl & 1], (uint32_t)(((bl + 2) >> 1) & 1));
    }
    if (live) {
      for (long long i = flushed + g; i < L; i += 8) yc[i] = myring[(int)i & (kDuoRing - 1)];
      // state: chronological ring, oldest at slot 0 = where the next output goes
      for (int j = g; j < N; j += 8) ring_g[(long long)j * C + c] = myring[(int)(L - N + j) & (kDuoRing - 1)];
      if (g == 0) head_g[c] = 0;
    }
    return;
  }

  // =========================================== CHAIN warp ===========================================
  const float tA = taps[g + 8 * (K - 1)], tB = taps[N + g + 8 * (K - 1)];   // tap of this lane's newest element
  const float nA = taps[N - 1], nBq = taps[2 * N - 1];                        // lane 7's: tap of out[n] in window n
  const SinCosF SK = sincos_f_load_consts();
  const float2 pf = pf_g[c];
  float phase = pf.x, freq = pf.y;
  float4 Pp = make_float4(0.f, 0.f, 0.f, 0.f);              // this lane's prefix from the previous step
  const float4* lp7_base = &S.lp[0][0][sl * 8 + 7];
  // one sample step; WARM: warm-up (the output is read back from the ring, no loop update)
  auto step = [&](auto warm_tag, long long b, int j) {
    constexpr bool WARM = decltype(warm_tag)::value;
    const int n = (int)(4 * b + j) & (kDuoRing - 1);
    // known before the rotation: stale partials, previous-step prefixes of the neighbours
    const float4 lpo = S.lp[b & 1][j][lane];
    const float4 lp7 = lp7_base[((int)(b & 1) * kDuoBatch + j) * 32];
    float4 Pin, P6;
    Pin.x = __shfl_up_sync(0xffffffffu, Pp.x, 1, 8);
    Pin.y = __shfl_up_sync(0xffffffffu, Pp.y, 1, 8);
    Pin.z = __shfl_up_sync(0xffffffffu, Pp.z, 1, 8);
    Pin.w = __shfl_up_sync(0xffffffffu, Pp.w, 1, 8);
    if (!WARM) {
      P6.x = __shfl_sync(0xffffffffu, Pp.x, 6, 8);
      P6.y = __shfl_sync(0xffffffffu, Pp.y, 6, 8);
      P6.z = __shfl_sync(0xffffffffu, Pp.z, 6, 8);
      P6.w = __shfl_sync(0xffffffffu, Pp.w, 6, 8);
    }
    if (g == 0) Pin = make_float4(0.f, 0.f, 0.f, 0.f);       // aLo = 0; aLo += lane 0 (:176-180)
    float oI, oQ;
    if (WARM) {
      const float2 o = myring[n];                            // outputs of earlier calls
      oI = o.x;
      oQ = o.y;
    } else {
      float s, co;
      sincos_f32arg_k(phase, SK, &s, &co);                   // MathF.Cos/Sin(phase) :108-109
      const float2 in = (S.xq[(b >> 2) & 1] + sl * kDuoXStride + (int)(b & 3) * kDuoBatch)[j];
      oI = in.x * co - in.y * s;                             // :111
      oQ = in.x * s + in.y * co;                             // :112
    }
    // this lane's newest element completes L_g of window n+7-g; extend that window's prefix
    float4 Lg = lpo;
    duo_acc(Lg, tA, tB, oI, oQ);
    Pp.x = Pin.x + Lg.x;
    Pp.y = Pin.y + Lg.y;
    Pp.z = Pin.z + Lg.z;
    Pp.w = Pin.w + Lg.w;
    if (!WARM) {
      // lane 7's role for the current window, on every lane: acc = P_6(n) + L_7(n)
      float4 L7 = lp7;
      duo_acc(L7, nA, nBq, oI, oQ);
      const float aLoI = P6.x + L7.x, aLoQ = P6.y + L7.y, aUpI = P6.z + L7.z, aUpQ = P6.w + L7.w;
      const float powUpper = aUpI * aUpI + aUpQ * aUpQ;      // :118
      const float powLower = aLoI * aLoI + aLoQ * aLoQ;      // :119
      const float error = powLower - powUpper;               // :121
      freq += P.beta * error;                                // :124
      phase += freq + P.alpha * error;                       // :125
      if (g == 7) myring[n] = make_float2(oI, oQ);
      if (phase > kTwoPiF || phase < -kTwoPiF) phase = remainderf(phase, kTwoPiF);   // :185-189
      if (freq > P.max_freq) freq = P.max_freq;              // :191-195
      else if (freq < P.min_freq) freq = P.min_freq;
    }
  };
  auto batch_done = [&](long long b) {
    __syncwarp();
    if (lane == 0) duo_mbar_arrive(&S.out_full[b & 1]);
  };
  auto batch_wait = [&](long long b) { duo_mbar_wait(&S.lp_full[b & 1], (uint32_t)(((b + 2) >> 1) & 1)); };
  // warm-up: fills the prefix pipeline from the outputs of earlier calls (steps -8..-1)
  for (long long b = -2; b < 0; ++b) {
    batch_wait(b);
#pragma unroll
    for (int j = 0; j < kDuoBatch; ++j) step(std::true_type{}, b, j);
    batch_done(b);
  }
  const long long nFull = L / kDuoBatch;
  for (long long b = 0; b < nFull; ++b) {
    batch_wait(b);
#pragma unroll
    for (int j = 0; j < kDuoBatch; ++j) step(std::false_type{}, b, j);
    batch_done(b);
  }
  if (nFull < nB) {
    batch_wait(nFull);
    const int ns = (int)(L - 4 * nFull);
    for (int j = 0; j < ns; ++j) step(std::false_type{}, nFull, j);
    batch_done(nFull);
  }
  if (live && g == 0) pf_g[c] = make_float2(phase, freq);
}

template <int PAIRS>
int launch_duo(int K, const FllParams& P, const float* taps, float2* ring, int* head, float2* pf, int C, const float2* x,
               float2* y, long long L, long long ldx, long long ldy, cudaStream_t s) {
  const int per_cta = PAIRS * kDuoStreams;
  const int blocks = (C + per_cta - 1) / per_cta;
  switch (K) {
#define QPSK_DUO_CASE(KK) \
  case KK: fll_duo_kernel<KK, PAIRS><<<blocks, 64 * PAIRS, 0, s>>>(P, taps, ring, head, pf, C, x, y, L, ldx, ldy); break;
    QPSK_DUO_CASE(1) QPSK_DUO_CASE(2) QPSK_DUO_CASE(3) QPSK_DUO_CASE(4) QPSK_DUO_CASE(5) QPSK_DUO_CASE(6)
#undef QPSK_DUO_CASE
    default: return QPSK_ERR_UNSUPPORTED;
  }
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

}  // namespace

bool fll_duo_supported(int n_taps) { return n_taps >= 8 && n_taps <= 48 && (n_taps & 7) == 0; }

int fll_duo_launch(const FllParams& P, const float* taps, float2* ring, int* head, float2* pf, int C, const float2* x,
                   float2* y, long long L, long long ldx, long long ldy, cudaStream_t s) {
  return launch_duo<1>(P.n_taps / 8, P, taps, ring, head, pf, C, x, y, L, ldx, ldy, s);
}

}  // namespace qpsk
