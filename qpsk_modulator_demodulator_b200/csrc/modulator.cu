// modulator.cu — K3: QPSKModulator on the GPU (MS/QPSKModulator.cs:18-168).
//
// The reference builds a zero-stuffed symbol train `up` (one symbol every sps samples, `delay` zeros
// in front and behind, :120-159) and pushes it through fftFilter (:166, FIRFilter.cs:96-141), i.e.
//     y[i] = sum_d sym[d] * h[i + (N-1-delay) - d*sps],   0 <= i < 2*delay + nDibits*sps
// (N-1-delay == delay for the usual odd N).
// Here that is computed directly as a polyphase interpolator: no zero-stuffed buffer, no FFT.
//
//   mod_tile_rot_kernel    per tile: sum of the differential rotations mod 4 (16 payload bits per thread)
//   mod_tile_scan_kernel   one warp per frame: exclusive scan of the tile sums (the differential encoder
//                          sym[d] = sym[d-1]*delta[d] (:137-146) is a prefix sum of quadrants)
//   mod_shape_kernel<MT>   per (frame, tile): every thread unpacks 8 dibits (two framed bytes) -> block
//                          scan -> 2048 symbols (tile + halo) in shared memory -> polyphase FIR.
//                          A thread owns one phase p and walks groups of R = 8 consecutive symbols: the
//                          MT = ceil(N/sps) taps of its phase stay in registers, the symbols slide through
//                          a circular register window (one LDS.128 per two taps), 8 packed FFMA2 per tap.
//                          Lanes run over consecutive phases, so a store instruction writes runs of sps
//                          consecutive samples (whole 32-byte sectors for sps >= 4).
//   Algorithmic traffic: 8 B written per output sample, 0.25/sps B read twice.
//
// Symbols are exactly (+-1/sqrt2, +-1/sqrt2) in fp32 (prev*delta with delta on the axes is exact), so
// the quadrant index carries all the information: q = 0:(+,+) 1:(-,+) 2:(-,-) 3:(+,-), and the
// rotations are 00->+0, 01->+1, 11->+2, 10->+3 quadrants (DibitToDelta :92-102).
#include <string>

#include "common.cuh"
#include "design.h"

namespace qpsk {

constexpr int kModThreads = 256;
constexpr int kModR = 8;                         // symbols per group / register window
constexpr int kModRegion = kModThreads * 8;      // symbols (halo + tile) staged per CTA: 8 per thread
constexpr int kModMaxMT = 96;                    // taps per phase supported by the register-resident kernel
constexpr float kInvSqrt2 = 0.7071067811865475f; // QPSKModulator.cs:36

struct ModArgs {
  // source: mode 0 = framed bytes (tsc | start | payload[f] | end), mode 1 = one code byte per dibit
  const uint8_t* payload;   // mode 0: [frames][n_payload]; mode 1: codes [n_dibits]
  const uint8_t* meta;      // tsc values (0,1,2=other) | start bytes | end bytes
  long long n_payload;
  int n_tsc, n_start, n_end;
  int mode, diff;
  long long n_dibits;       // per frame
  long long total;          // complex output samples per frame
  int tiles;                // per frame
  int frames;
  int sps, MT, HP, TD, delay;  // MT taps per phase (padded), HP halo symbols (multiple of 8), TD = 2048 - HP
  const float* poly;        // [sps][MT]: poly[p*MT + m] = h[p + m*sps]
  uint8_t* tile_sum;        // [frames][tiles]
  uint8_t* tile_pre;        // [frames][tiles]
  float2* out;
  long long out_stride;     // complex samples between frames
};

__device__ __forceinline__ int code_from(int b0, int b1, int diff) {
  if (diff) {                                   // DibitToDelta :92-102 (anything else -> -j)
    if (b0 == 0 && b1 == 0) return 0;
    if (b0 == 0 && b1 == 1) return 1;
    if (b0 == 1 && b1 == 1) return 2;
    return 3;
  }
  return ((b0 != 0) << 1) | (b1 != 0);          // :150-151  (bit==0 ? -1/sqrt2 : +1/sqrt2)
}

__device__ __forceinline__ int frame_byte(const ModArgs& a, int f, long long b) {
  if (b < a.n_start) return a.meta[a.n_tsc + b];
  if (b < a.n_start + a.n_payload) return a.payload[(long long)f * a.n_payload + (b - a.n_start)];
  return a.meta[a.n_tsc + a.n_start + (b - a.n_start - a.n_payload)];
}

__device__ __forceinline__ int frame_bit(const ModArgs& a, int f, long long k) {
  if (k < a.n_tsc) return a.meta[k];
  k -= a.n_tsc;
  return (frame_byte(a, f, k >> 3) >> (7 - (int)(k & 7))) & 1;   // MSB first (HelperFunctions.cs:14-29)
}

__device__ __forceinline__ int dibit_code(const ModArgs& a, int f, long long d) {
  if (d < 0 || d >= a.n_dibits) return 0;
  if (a.mode == 1) return a.payload[d];
  return code_from(frame_bit(a, f, 2 * d), frame_bit(a, f, 2 * d + 1), a.diff);
}

// Codes of the 8 symbols d0 .. d0+7 (d0 a multiple of 8), packed MSB first: symbol j in bits [15-2j, 14-2j].
// Symbols outside [0, n_dibits) get code 0.
__device__ __forceinline__ unsigned chunk_codes(const ModArgs& a, int f, long long d0) {
  if (d0 + 8 <= 0 || d0 >= a.n_dibits) return 0u;
  if (a.mode == 0 && d0 >= 0 && d0 + 8 <= a.n_dibits && (a.n_tsc & 7) == 0) {
    const long long k0 = 2 * d0 - a.n_tsc;       // first bit of the chunk inside the framed bytes
    if (k0 >= 0) {
      // 16 framed bits, MSB first; per 2-bit field b0b1: 00->0 01->1 11->2 10->3 is two ^ (two >> 1)
      const long long b = k0 >> 3;
      unsigned v;
      const long long pb = b - a.n_start;
      if (pb >= 0 && pb + 1 < a.n_payload) {
        const uint8_t* q = a.payload + (long long)f * a.n_payload + pb;
        v = ((unsigned)q[0] << 8) | q[1];
      } else {
        v = ((unsigned)frame_byte(a, f, b) << 8) | (unsigned)frame_byte(a, f, b + 1);
      }
      return a.diff ? (v ^ ((v >> 1) & 0x5555u)) : v;
    }
  }
  unsigned c = 0;                                 // TSC characters, odd TSC length, frame edges, code-byte mode
#pragma unroll 1
  for (int j = 0; j < 8; ++j) c = (c << 2) | (unsigned)dibit_code(a, f, d0 + j);
  return c;
}

// one warp per tile (8 tiles per CTA): every lane sums the rotations of its 8-symbol chunks lane, lane+32, ...
// (the first version spent a 256-thread CTA on a tile, 16 payload bits per thread: 2.1 M warps for 134 MB)
__global__ void __launch_bounds__(kModThreads) mod_tile_rot_kernel(const ModArgs a) {
  const int f = blockIdx.y;
  const int t = blockIdx.x * (kModThreads / 32) + (threadIdx.x >> 5);
  if (t >= a.tiles) return;
  const int lane = threadIdx.x & 31;
  const long long D0 = (long long)t * a.TD;
  unsigned s = 0;
  // Interior tiles (all TD dibits inside the payload bytes, dibits byte-aligned): the sum of the 2-bit rotation codes
  // does not depend on the order of the dibits, so the bytes are summed 16 per lane as 32-bit words (aligned loads +
  // funnel shifts), code = v ^ ((v >> 1) & 0x5555...) per field exactly as in chunk_codes.
  if (a.mode == 0 && a.diff && (a.n_tsc & 7) == 0 && (a.TD & 3) == 0) {
    const long long k0 = 2 * D0 - a.n_tsc;                   // first framed bit of the tile
    const long long pb = (k0 >> 3) - a.n_start;              // its payload byte
    const int nbytes = a.TD >> 2;
    if (k0 >= 0 && (k0 & 7) == 0 && pb >= 0 && pb + nbytes <= a.n_payload) {
      const uint8_t* q = a.payload + (long long)f * a.n_payload + pb;
      const int off = (int)(reinterpret_cast<uintptr_t>(q) & 3u);
      const int sh = 8 * off;
      const uint32_t* qa = reinterpret_cast<const uint32_t*>(q - off);         // aligned word holding q[0]
      for (int b0 = 16 * lane; b0 < nbytes; b0 += 512) {
        const int nb = (nbytes - b0 < 16) ? (nbytes - b0) : 16;               // bytes of this lane's piece
        const uint32_t* w = qa + (b0 >> 2);
        uint32_t cur = w[0];                                                  // holds q[b0]
        for (int i = 0; 4 * i < nb; ++i) {
          const int left = nb - 4 * i;
          const int take = left < 4 ? left : 4;
          // w[i+1] is read only when it holds a byte of the piece (an aligned word that overlaps the row is readable)
          const uint32_t nxt = (off + take > 4 || 4 * (i + 1) < nb) ? w[i + 1] : 0u;
          uint32_t v = __funnelshift_r(cur, nxt, sh);                         // bytes q[b0+4i .. +3], little endian
          if (take < 4) v &= (1u << (8 * take)) - 1u;                         // zero bytes add nothing (00 -> +0)
          uint32_t c = v ^ ((v >> 1) & 0x55555555u);
          c = (c & 0x33333333u) + ((c >> 2) & 0x33333333u);
          c = (c & 0x0F0F0F0Fu) + ((c >> 4) & 0x0F0F0F0Fu);
          s += (c * 0x01010101u) >> 24;
          cur = nxt;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) a.tile_sum[(long long)f * a.tiles + t] = (uint8_t)(s & 3u);
      return;
    }
  }
  for (int k = lane; 8 * k < a.TD; k += 32) {
    unsigned c = chunk_codes(a, f, D0 + 8 * k);
    c = (c & 0x3333u) + ((c >> 2) & 0x3333u);   // sum of the eight 2-bit fields
    c = (c & 0x0F0Fu) + ((c >> 4) & 0x0F0Fu);
    s += (c & 0xFFu) + (c >> 8);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) a.tile_sum[(long long)f * a.tiles + t] = (uint8_t)(s & 3u);
}

__global__ void mod_tile_scan_kernel(const ModArgs a) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= a.frames) return;
  const int lane = threadIdx.x & 31;
  const uint8_t* ts = a.tile_sum + (long long)f * a.tiles;
  uint8_t* tp = a.tile_pre + (long long)f * a.tiles;
  int carry = 0;
  for (int t0 = 0; t0 < a.tiles; t0 += 32) {
    const int t = t0 + lane;
    const int v = (t < a.tiles) ? ts[t] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (t < a.tiles) tp[t] = (uint8_t)((carry + inc - v) & 3);
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
}

__device__ __forceinline__ float2 quadrant_symbol(int q) {
  const float i = (q == 0 || q == 3) ? kInvSqrt2 : -kInvSqrt2;
  const float v = (q == 0 || q == 1) ? kInvSqrt2 : -kInvSqrt2;
  return make_float2(i, v);
}

// SPS > 0: samples per symbol known at compile time (2, 4, 8 — the store offsets r*sps become immediates and tid / sps a
// shift); SPS == 0: read from the arguments.
template <int MT, int SPS = 0>
__global__ void __launch_bounds__(kModThreads) mod_shape_kernel(const ModArgs a) {
  constexpr int R = kModR;
  // symbol D0 - HP + k lives at sym[k + 2*(k/8)]: two pad slots after every 8 symbols make the per-thread and
  // per-group stride 80 B (5 x 16 B, odd), so the 128-bit accesses of a quarter-warp fall in distinct banks
  __shared__ __align__(16) float2 sym[kModRegion + kModRegion / 4];
  __shared__ __align__(16) int warp_tot[kModThreads / 32];
  __shared__ int halo_sum_s;
  // SPS 2 / 4 / 8: the 32 lanes of a warp own 32 / SPS consecutive groups = 256 consecutive output samples (2 KiB) per round,
  // but lane (group, phase) holds every SPS-th sample of its group: stored straight from the registers, one instruction
  // touches 32 / SPS different 128-byte lines with 8 * SPS bytes each (ncu round 1: l1tex 89 % busy, the kernel's limiter).
  // The round goes through a warp-private staging row instead — groups (R + 1) * SPS float2 apart, so that the 64-bit writes of
  // a half-warp fall in distinct banks — and leaves as whole lines.
  constexpr bool kStage = SPS == 2 || SPS == 4 || SPS == 8;
  constexpr int kStageRow = 32 * (R + 1);
  __shared__ __align__(16) float2 stage[kStage ? (kModThreads / 32) * kStageRow : 1];

  const int f = blockIdx.y, t = blockIdx.x;
  const long long D0 = (long long)t * a.TD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // ---- 8 dibits per thread -> quadrants (block-wide scan) -> symbols ----
  {
    const long long d0 = D0 - a.HP + 8 * tid;
    const unsigned c = chunk_codes(a, f, d0);
    // The eight 2-bit codes spread to one nibble each (symbol j in nibble 7-j), so that the running rotation of all eight
    // symbols is three shift-add steps on the packed word (every sum is only needed mod 4: the mask keeps each nibble
    // below 4, no carry ever crosses a nibble) instead of eight serial adds.
    unsigned x = c & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    unsigned sI, sQ;                                // sign of I / Q of symbol j in bit 4*(7-j)
    if (a.diff) {
      unsigned pre = x;                             // nibble 7-j: sum of the codes of symbols 0..j, mod 4
      pre = (pre + (pre >> 4)) & 0x33333333u;
      pre = (pre + (pre >> 8)) & 0x33333333u;
      pre = (pre + (pre >> 16)) & 0x33333333u;
      const int acc = (int)(pre & 3u);              // the chunk's total rotation
      int inc = acc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (lane == 31) warp_tot[warp] = inc;
      // sum over the halo = exclusive prefix of chunk HP/8 (warp 0): the tile prefix refers to symbol D0
      if (tid == (a.HP >> 3)) halo_sum_s = inc - acc;
      __syncthreads();
      int base = inc - acc;
      {
        const int4 t0 = *reinterpret_cast<const int4*>(warp_tot), t1 = *reinterpret_cast<const int4*>(warp_tot + 4);
        base += (warp > 0 ? t0.x : 0) + (warp > 1 ? t0.y : 0) + (warp > 2 ? t0.z : 0) + (warp > 3 ? t0.w : 0) +
                (warp > 4 ? t1.x : 0) + (warp > 5 ? t1.y : 0) + (warp > 6 ? t1.z : 0);
      }
      const unsigned q0 = (unsigned)((int)a.tile_pre[(long long)f * a.tiles + t] - halo_sum_s + base) & 3u;
      const unsigned q = (pre + q0 * 0x11111111u) & 0x33333333u;      // quadrant of every symbol
      // quadrant_symbol: q = 0:(+,+) 1:(-,+) 2:(-,-) 3:(+,-)  ->  I negative iff bit0 ^ bit1, Q negative iff bit1
      sI = (q ^ (q >> 1)) & 0x11111111u;
      sQ = (q >> 1) & 0x11111111u;
    } else {
      // :150-151  bit == 0 ? -1/sqrt2 : +1/sqrt2, I from the first bit of the dibit, Q from the second
      sI = (~x >> 1) & 0x11111111u;
      sQ = ~x & 0x11111111u;
    }
    const unsigned kMag = __float_as_uint(kInvSqrt2);
    float4* dst = reinterpret_cast<float4*>(sym + 10 * tid);
    if (d0 >= 0 && d0 + 8 <= a.n_dibits) {                    // all eight symbols exist (every chunk but the frame's edges)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float2 v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int sh = 31 - 4 * (7 - (j + u));    // moves the symbol's sign bit to bit 31
          v[u] = make_float2(__uint_as_float(kMag | ((sI << sh) & 0x80000000u)), __uint_as_float(kMag | ((sQ << sh) & 0x80000000u)));
        }
        dst[j >> 1] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
      }
    } else {                                                  // kept out of the common path: 64-bit range tests per symbol
      float2* d2 = sym + 10 * tid;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        const int sh = 3 + 4 * j;
        const long long d = d0 + j;
        float2 sv = make_float2(__uint_as_float(kMag | ((sI << sh) & 0x80000000u)), __uint_as_float(kMag | ((sQ << sh) & 0x80000000u)));
        if (d < 0 || d >= a.n_dibits) sv = make_float2(0.f, 0.f);
        d2[j] = sv;
      }
    }
  }
  __syncthreads();

  // ---- polyphase FIR: this thread's phase p, groups of R consecutive symbols ----
  const int sps = SPS > 0 ? SPS : a.sps;
  const int gl = tid / sps;                       // group slot within a round
  const int p = tid - gl * sps;
  const int G = kModThreads / sps;                // whole groups per round
  if (gl >= G) return;
  float tap[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) tap[m] = a.poly[p * MT + m];
  const int n_groups = a.TD / R;
  float2* ob = a.out + (long long)f * a.out_stride + (D0 * sps - a.delay);   // output of (symbol D0, phase 0)
  const long long i_tile = D0 * sps - a.delay;    // its index in the frame
  const bool tile_in = i_tile >= 0 && i_tile + (long long)n_groups * (R * sps) <= a.total;   // every sample of the tile exists
  for (int g = gl; g < n_groups; g += G) {
    const float2* gs = sym + 10 * ((a.HP >> 3) + g);   // the group's first symbol (padded layout)
    float2 w[R], acc[R];
    const float4* s4 = reinterpret_cast<const float4*>(gs);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
      const float4 v = s4[r >> 1];
      w[r] = make_float2(v.x, v.y);
      w[r + 1] = make_float2(v.z, v.w);
      acc[r] = make_float2(0.f, 0.f);
      acc[r + 1] = make_float2(0.f, 0.f);
    }
    // tap m multiplies sym[first + r - m]; slot of symbol offset j is (j mod R); one new (older) symbol per tap,
    // fetched two at a time (16-byte aligned for even m).  MT = span + 1 is odd for the usual N = span*sps + 1: the last
    // tap then stands alone and needs no fetch (padding it to a pair cost R FFMA2 and one LDS.128 per group).
#pragma unroll
    for (int m = 0; m < MT; m += 2) {
      constexpr int kLast = MT - 1;
      const bool pair = m < kLast;                 // compile-time after unrolling
      float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
      // symbols (first - m - 2, first - m - 1): block ceil((m+2)/8) back, slot (8 - (m+2)%8) % 8
      if (pair) nx = *reinterpret_cast<const float4*>(gs - 10 * ((m + 2 + 7) / 8) + ((8 - ((m + 2) & 7)) & 7));
      {
        const float2 tt = make_float2(tap[m], tap[m]);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = ffma2(w[(r - m + 16 * R) % R], tt, acc[r]);
        if (pair) w[(R - 1 - m + 16 * R) % R] = make_float2(nx.z, nx.w);      // symbol first - m - 1
      }
      if (pair) {
        const float2 tt = make_float2(tap[m + 1], tap[m + 1]);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = ffma2(w[(r - m - 1 + 16 * R) % R], tt, acc[r]);
        w[(R - 2 - m + 16 * R) % R] = make_float2(nx.x, nx.y);      // symbol first - m - 2
      }
    }
    if constexpr (kStage) {
      const int glw = lane / SPS;                   // group slot within the warp
      const int g_w = g - glw;                      // the warp's first group of this round
      const long long iw = i_tile + (long long)g_w * (R * SPS);
      // warp-uniform: all 32 / SPS groups exist and every sample lies inside the frame
      if (g_w + 32 / SPS <= n_groups && (tile_in || (iw >= 0 && iw + 32 * R <= a.total))) {
        float2* sw = stage + warp * kStageRow;
        float2* mine = sw + glw * ((R + 1) * SPS) + p;
#pragma unroll
        for (int r = 0; r < R; ++r) mine[r * SPS] = acc[r];
        __syncwarp();
        float2* og = ob + g_w * (R * SPS);
        if ((reinterpret_cast<unsigned long long>(og) & 15ull) == 0) {
#pragma unroll
          for (int k = 0; k < R / 2; ++k) {
            const int l2 = 2 * (k * 32 + lane);     // sample index inside the round
            const float4 v = *reinterpret_cast<const float4*>(sw + (l2 / (R * SPS)) * ((R + 1) * SPS) + l2 % (R * SPS));
            __stcs(reinterpret_cast<float4*>(og + l2), v);
          }
        } else {                                    // odd filter delay or odd frame stride: 8-byte stores, still contiguous
#pragma unroll
          for (int k = 0; k < R; ++k) {
            const int l1 = k * 32 + lane;
            __stcs(og + l1, sw[(l1 / (R * SPS)) * ((R + 1) * SPS) + l1 % (R * SPS)]);
          }
        }
        __syncwarp();
        continue;
      }
    }
    const int off = g * R * sps + p;              // relative to ob, r = 0
    const long long i0 = i_tile + off;
    if (tile_in || (i0 >= 0 && i0 + (long long)(R - 1) * sps < a.total)) {
#pragma unroll
      for (int r = 0; r < R; ++r) ob[off + r * sps] = acc[r];
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long i = i0 + (long long)r * sps;
        if (i >= 0 && i < a.total) ob[off + r * sps] = acc[r];
      }
    }
  }
}

typedef void (*ModShapeFn)(const ModArgs);
template <int SPS>
static ModShapeFn mod_shape_pick_sps(int mt) {
  switch (mt) {
    case 4: return mod_shape_kernel<4, SPS>;
    case 5: return mod_shape_kernel<5, SPS>;
    case 7: return mod_shape_kernel<7, SPS>;
    case 8: return mod_shape_kernel<8, SPS>;
    case 9: return mod_shape_kernel<9, SPS>;
    case 11: return mod_shape_kernel<11, SPS>;
    case 12: return mod_shape_kernel<12, SPS>;
    case 13: return mod_shape_kernel<13, SPS>;
    case 16: return mod_shape_kernel<16, SPS>;
    case 17: return mod_shape_kernel<17, SPS>;
    case 20: return mod_shape_kernel<20, SPS>;
    case 24: return mod_shape_kernel<24, SPS>;
    default: return nullptr;
  }
}
static ModShapeFn mod_shape_pick(int mt, int sps) {
  ModShapeFn f = nullptr;
  if (sps == 2) f = mod_shape_pick_sps<2>(mt);
  else if (sps == 4) f = mod_shape_pick_sps<4>(mt);
  else if (sps == 8) f = mod_shape_pick_sps<8>(mt);
  if (f) return f;
  switch (mt) {
    case 4: return mod_shape_kernel<4>;
    case 5: return mod_shape_kernel<5>;
    case 7: return mod_shape_kernel<7>;
    case 8: return mod_shape_kernel<8>;
    case 9: return mod_shape_kernel<9>;
    case 11: return mod_shape_kernel<11>;
    case 13: return mod_shape_kernel<13>;
    case 17: return mod_shape_kernel<17>;
    case 12: return mod_shape_kernel<12>;
    case 16: return mod_shape_kernel<16>;
    case 20: return mod_shape_kernel<20>;
    case 24: return mod_shape_kernel<24>;
    case 32: return mod_shape_kernel<32>;
    case 48: return mod_shape_kernel<48>;
    case 64: return mod_shape_kernel<64>;
    case 96: return mod_shape_kernel<96>;
    default: return nullptr;
  }
}
static int mod_pad_taps(int m) {
  static const int sizes[] = {4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 20, 24, 32, 48, 64, 96};   // span + 1 for the usual spans, exactly
  for (int v : sizes)
    if (m <= v) return v;
  return 0;
}

// ---------------------------------------------------------------------------------------------
struct ModEngine {
  int fs = 0, rs = 0;
  int device = 0;
  bool diff = true, has_tsc = false;
  std::string tsc;
  std::vector<double> taps_d;
  std::vector<float> taps_f;
  int sps = 0, delay = 0;
  // two polyphase banks: [0] pulse shaping with the RRC taps, [1] the "no pulse shaping" delta
  DevBuf<float> d_poly[2];
  int MT[2] = {0, 0}, bank_delay[2] = {0, 0};
  DevBuf<uint8_t> d_meta, d_tile_sum, d_tile_pre, d_src;
  DevBuf<float2> d_out;
  std::vector<uint8_t> meta_host;
  cudaStream_t stream = nullptr;

  ~ModEngine() {
    if (stream) cudaStreamDestroy(stream);
  }

  int make_bank(int which, const std::vector<float>& h, int bank_delay_in) {
    const int n = (int)h.size();
    const int m = (n + sps - 1) / sps;            // taps per phase
    const int mt = mod_pad_taps(m);
    MT[which] = mt; bank_delay[which] = bank_delay_in;
    if (mt == 0) return QPSK_OK;                  // too long for the register-resident kernel: reported at run()
    std::vector<float> poly((size_t)sps * mt, 0.0f);
    for (int p = 0; p < sps; ++p)
      for (int k = 0; k < m; ++k) {
        const long long j = p + (long long)k * sps;
        if (j < n) poly[(size_t)p * mt + k] = h[(size_t)j];
      }
    QPSK_TRY(d_poly[which].alloc(poly.size()));
    QPSK_CUDA_TRY(cudaMemcpyAsync(d_poly[which].p, poly.data(), poly.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    QPSK_CUDA_TRY(cudaStreamSynchronize(stream));
    return QPSK_OK;
  }

  int init(int fs_in, int rs_in, double alpha, int span, int diff_in, const char* tsc_in) {
    if (rs_in == 0) return QPSK_ERR_RANGE;
    QPSK_TRY(ensure_device());
    device = current_device();
    fs = fs_in; rs = rs_in; diff = diff_in != 0;
    has_tsc = !blank_or_null(tsc_in);              // :27
    if (has_tsc) tsc = tsc_in;
    taps_d = design_rrc((double)span, alpha, fs, rs);   // :29-30
    taps_f.resize(taps_d.size());
    for (size_t i = 0; i < taps_d.size(); ++i) taps_f[i] = (float)taps_d[i];   // :43-53
    sps = fs / rs;                                  // :115 integer division
    delay = ((int)taps_d.size() - 1) / 2;           // :119
    QPSK_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (sps > 0 && sps <= kModThreads && !taps_f.empty()) {
      // fftFilter advances by N-1 (FIRFilter.cs:130-138) while the symbols sit at delay + d*sps (:128,154):
      // tap index = i + (N-1-delay) - d*sps.  N-1-delay == delay only for odd N.
      QPSK_TRY(make_bank(0, taps_f, (int)taps_f.size() - 1 - delay));
      // up[delay + d*sps] = sym[d]  ==  polyphase with h = delta[j - delay] and zero filter delay
      std::vector<float> delta((size_t)delay + 1, 0.0f);
      delta[(size_t)delay] = 1.0f;
      QPSK_TRY(make_bank(1, delta, 0));
    }
    return QPSK_OK;
  }

  // sizes (:112-121): returns complex samples per frame for n_bits data bits (TSC added here)
  int64_t frame_complex(int64_t n_bits, bool pulse, int64_t* n_dibits) const {
    const int64_t all = n_bits + (has_tsc ? (int64_t)tsc.size() : 0);
    const int64_t nd = all >> 1;
    if (n_dibits) *n_dibits = nd;
    if (nd == 0) return 0;
    const int64_t base = delay + nd * sps;
    return pulse ? base + delay : base;
  }

  int upload_meta(const uint8_t* sm, int64_t ns, const uint8_t* em, int64_t ne, cudaStream_t s) {
    std::vector<uint8_t> m;
    if (has_tsc)
      for (char c : tsc) m.push_back((uint8_t)(c == '0' ? 0 : (c == '1' ? 1 : 2)));
    m.insert(m.end(), sm, sm + ns);
    m.insert(m.end(), em, em + ne);
    if (m.empty()) m.push_back(0);
    if (m != meta_host || !d_meta.p) {
      QPSK_CUDA_TRY(cudaStreamSynchronize(s));      // the previous launch may still read the old meta
      QPSK_TRY(d_meta.ensure(m.size()));
      meta_host = m;
      QPSK_CUDA_TRY(cudaMemcpyAsync(d_meta.p, meta_host.data(), meta_host.size(), cudaMemcpyHostToDevice, s));
    }
    return QPSK_OK;
  }

  // launches the kernels; `a` has source fields filled in
  int run(ModArgs a, bool pulse, int64_t n_dibits, int frames, float2* out, int64_t out_stride, cudaStream_t s) {
    const int bank = pulse ? 0 : 1;
    if (sps > kModThreads) return QPSK_ERR_UNSUPPORTED;
    ModShapeFn shape = mod_shape_pick(MT[bank], sps);
    if (!shape) return QPSK_ERR_UNSUPPORTED;        // more than kModMaxMT taps per phase (DESIGN.md limits)
    a.diff = diff ? 1 : 0;
    a.n_dibits = n_dibits;
    a.frames = frames;
    a.sps = sps; a.MT = MT[bank]; a.delay = bank_delay[bank];
    a.HP = (a.MT + 7) & ~7;
    a.TD = kModRegion - a.HP;
    a.poly = d_poly[bank].p;
    a.total = pulse ? (2LL * delay + n_dibits * sps) : ((long long)delay + n_dibits * sps);
    // last output index + filter delay, in symbols (+1): symbols past n_dibits are zero
    const long long n_virtual = (a.total - 1 + a.delay) / sps + 1;
    const long long tiles = (n_virtual + a.TD - 1) / a.TD;
    if (tiles > 0x7fffffffLL) return QPSK_ERR_UNSUPPORTED;
    a.tiles = (int)tiles;
    a.out_stride = out_stride;
    // frames ride on gridDim.y (<= 65535): larger batches go in slices, each with its own view of the frame arrays
    constexpr int kFrameSlice = 65535;
    const int slice = frames < kFrameSlice ? frames : kFrameSlice;
    QPSK_TRY(d_tile_sum.ensure((size_t)tiles * slice));
    QPSK_TRY(d_tile_pre.ensure((size_t)tiles * slice));
    a.tile_sum = d_tile_sum.p; a.tile_pre = d_tile_pre.p;
    const uint8_t* payload0 = a.payload;
    for (int f0 = 0; f0 < frames; f0 += kFrameSlice) {
      const int nf = (frames - f0 < kFrameSlice) ? (frames - f0) : kFrameSlice;
      a.frames = nf;
      if (a.mode == 0 && payload0) a.payload = payload0 + (long long)f0 * a.n_payload;
      a.out = out + (long long)f0 * out_stride;
      const dim3 grid((unsigned)tiles, (unsigned)nf);
      if (diff) {
        const dim3 grid_rot((unsigned)((tiles + kModThreads / 32 - 1) / (kModThreads / 32)), (unsigned)nf);
        mod_tile_rot_kernel<<<grid_rot, kModThreads, 0, s>>>(a);
        QPSK_LAUNCH_CHECK();
        mod_tile_scan_kernel<<<(nf + 3) / 4, 128, 0, s>>>(a);
        QPSK_LAUNCH_CHECK();
      }
      shape<<<grid, kModThreads, 0, s>>>(a);
      QPSK_LAUNCH_CHECK();
    }
    return QPSK_OK;
  }
};

}  // namespace qpsk

using namespace qpsk;

struct qpsk_mod {
  ModEngine eng;
  // host batch pipeline (qpsk_mod_modulate_frames): payload H2D + shaping kernel on eng.stream, samples D2H on s_out,
  // kSlots groups of frames in flight
  static constexpr int kSlots = 3;
  cudaStream_t s_out = nullptr;
  cudaEvent_t ev_k[kSlots] = {}, ev_d2h[kSlots] = {};
  DevBuf<uint8_t> d_pay[kSlots];
  DevBuf<float2> d_grp[kSlots];
  // page-locked landing slots for a pageable output block (common.cuh: host copy pool), allocated on first use
  float2* h_grp[kSlots] = {};
  size_t h_grp_elems = 0;
  int landing(size_t elems) {
    if (elems <= h_grp_elems) return QPSK_OK;
    for (int i = 0; i < kSlots; ++i) {
      if (h_grp[i]) cudaFreeHost(h_grp[i]);
      h_grp[i] = nullptr;
    }
    h_grp_elems = 0;
    for (int i = 0; i < kSlots; ++i) QPSK_CUDA_TRY(cudaHostAlloc((void**)&h_grp[i], elems * sizeof(float2), cudaHostAllocPortable));
    h_grp_elems = elems;
    return QPSK_OK;
  }
  ~qpsk_mod() {
    for (int i = 0; i < kSlots; ++i) {
      if (ev_k[i]) cudaEventDestroy(ev_k[i]);
      if (ev_d2h[i]) cudaEventDestroy(ev_d2h[i]);
      if (h_grp[i]) cudaFreeHost(h_grp[i]);
    }
    if (s_out) cudaStreamDestroy(s_out);
  }
};

// Modulate(string bits, bool pulseShaping) :104-167 for a bit source `val(k)` in {0, 1, 2 = any other character}
template <typename BitAt>
static int mod_bits_common(qpsk_mod* m, BitAt data_bit, int64_t n_bits, int pulse_shaping, float* iq_out, int64_t cap_floats,
                           int64_t* n_floats) {
  ModEngine& e = m->eng;
  int64_t nd = 0;
  const bool pulse = pulse_shaping != 0;
  const int64_t total = e.frame_complex(n_bits, pulse, &nd);
  *n_floats = 2 * total;
  if (nd == 0) return QPSK_OK;                              // :113
  if (e.sps <= 0) return QPSK_ERR_RANGE;                    // :116-117
  if (!iq_out) return QPSK_OK;                              // size query
  if (cap_floats < 2 * total) return QPSK_ERR_CAPACITY;
  QPSK_TRY(ensure_device(m->eng.device));
  // one code per dibit, with the reference's `c - '0'` semantics for any character
  std::vector<uint8_t> codes((size_t)nd);
  const int64_t nt = e.has_tsc ? (int64_t)e.tsc.size() : 0;
  auto val = [&](int64_t k) -> int {
    if (k < nt) {
      const char c = e.tsc[(size_t)k];
      return c == '0' ? 0 : (c == '1' ? 1 : 2);
    }
    return data_bit(k - nt);
  };
  for (int64_t d = 0; d < nd; ++d) {
    const int b0 = val(2 * d), b1 = val(2 * d + 1);
    int code;
    if (e.diff) code = (b0 == 0 && b1 == 0) ? 0 : (b0 == 0 && b1 == 1) ? 1 : (b0 == 1 && b1 == 1) ? 2 : 3;
    else code = ((b0 != 0) << 1) | (b1 != 0);
    codes[(size_t)d] = (uint8_t)code;
  }
  QPSK_TRY(e.d_src.ensure((size_t)nd));
  QPSK_TRY(e.d_out.ensure((size_t)total));
  cudaStream_t s = e.stream;
  QPSK_CUDA_TRY(cudaMemcpyAsync(e.d_src.p, codes.data(), (size_t)nd, cudaMemcpyHostToDevice, s));
  ModArgs a{};
  a.mode = 1; a.payload = e.d_src.p; a.meta = nullptr; a.n_payload = 0; a.n_tsc = a.n_start = a.n_end = 0;
  QPSK_TRY(e.run(a, pulse, nd, 1, e.d_out.p, total, s));
  QPSK_CUDA_TRY(cudaMemcpyAsync(iq_out, e.d_out.p, (size_t)total * 8, cudaMemcpyDeviceToHost, s));
  QPSK_CUDA_TRY(cudaStreamSynchronize(s));
  return QPSK_OK;
}

extern "C" {

int qpsk_mod_create(int sample_rate, int symbol_rate, double rrc_alpha, int rrc_span, int differential,
                    const char* tsc_bits, qpsk_mod** out) {
  if (!out) return QPSK_ERR_NULL;
  *out = nullptr;
  qpsk_mod* m = new (std::nothrow) qpsk_mod();
  if (!m) return QPSK_ERR_NOMEM;
  int st = m->eng.init(sample_rate, symbol_rate, rrc_alpha, rrc_span, differential, tsc_bits);
  if (st != QPSK_OK) { delete m; return st; }
  *out = m;
  return QPSK_OK;
}

int qpsk_mod_destroy(qpsk_mod* m) {
  if (m) {
    if (m->eng.stream) cudaStreamSynchronize(m->eng.stream);
    delete m;
  }
  return QPSK_OK;
}

int qpsk_mod_taps(const qpsk_mod* m, double* out, int cap, int* n) {
  if (!m || !n) return QPSK_ERR_NULL;
  *n = (int)m->eng.taps_d.size();
  if (!out) return QPSK_OK;
  if (cap < *n) return QPSK_ERR_CAPACITY;
  if (*n) memcpy(out, m->eng.taps_d.data(), sizeof(double) * (size_t)*n);
  return QPSK_OK;
}

int qpsk_mod_modulate_bits(qpsk_mod* m, const char* bits, int64_t n_bits, int pulse_shaping, float* iq_out,
                           int64_t cap_floats, int64_t* n_floats) {
  if (!m || !n_floats) return QPSK_ERR_NULL;
  if (!bits && n_bits > 0) return QPSK_ERR_NULL;            // :106
  if (n_bits < 0) return QPSK_ERR_RANGE;
  auto at = [&](int64_t k) -> int {
    const char c = bits[k];
    return c == '0' ? 0 : (c == '1' ? 1 : 2);
  };
  return mod_bits_common(m, at, n_bits, pulse_shaping, iq_out, cap_floats, n_floats);
}

// the same call with the bit string packed MSB-first, 8 bits per byte (what BitPacker.BytesToBitString :14-29 expands)
int qpsk_mod_modulate_packed(qpsk_mod* m, const uint8_t* packed_bits, int64_t n_bits, int pulse_shaping, float* iq_out,
                             int64_t cap_floats, int64_t* n_floats) {
  if (!m || !n_floats) return QPSK_ERR_NULL;
  if (!packed_bits && n_bits > 0) return QPSK_ERR_NULL;
  if (n_bits < 0) return QPSK_ERR_RANGE;
  auto at = [&](int64_t k) -> int { return (packed_bits[k >> 3] >> (7 - (int)(k & 7))) & 1; };
  return mod_bits_common(m, at, n_bits, pulse_shaping, iq_out, cap_floats, n_floats);
}

static int mod_frames_common(qpsk_mod* m, const uint8_t* d_payloads, int64_t n_payload, int frames,
                             const uint8_t* sm, int64_t ns, const uint8_t* em, int64_t ne, int pulse_shaping,
                             float2* d_out, int64_t out_stride_c, int64_t* frame_floats, cudaStream_t s) {
  ModEngine& e = m->eng;
  if (ns == 0 || ne == 0) return QPSK_ERR_ARG;               // :60-61
  if (!sm || !em) return QPSK_ERR_NULL;
  if (n_payload < 0 || ns < 0 || ne < 0 || frames < 0) return QPSK_ERR_RANGE;
  int64_t nd = 0;
  const bool pulse = pulse_shaping != 0;
  const int64_t total = e.frame_complex(8 * (ns + n_payload + ne), pulse, &nd);
  if (frame_floats) *frame_floats = 2 * total;
  if (nd == 0 || frames == 0) return QPSK_OK;
  if (e.sps <= 0) return QPSK_ERR_RANGE;
  if (!d_out) return QPSK_OK;                                // size query
  if (n_payload > 0 && !d_payloads) return QPSK_ERR_NULL;
  if (frames > 1 && out_stride_c < total) return QPSK_ERR_ARG;
  QPSK_TRY(ensure_device(m->eng.device));
  QPSK_TRY(e.upload_meta(sm, ns, em, ne, s));
  ModArgs a{};
  a.mode = 0; a.payload = d_payloads; a.meta = e.d_meta.p; a.n_payload = n_payload;
  a.n_tsc = e.has_tsc ? (int)e.tsc.size() : 0; a.n_start = (int)ns; a.n_end = (int)ne;
  return e.run(a, pulse, nd, frames, d_out, out_stride_c, s);
}

int qpsk_mod_modulate_bytes(qpsk_mod* m, const uint8_t* payload, int64_t n_payload, const uint8_t* start_marker,
                            int64_t n_start, const uint8_t* end_marker, int64_t n_end, int pulse_shaping,
                            float* iq_out, int64_t cap_floats, int64_t* n_floats) {
  if (!m || !n_floats) return QPSK_ERR_NULL;
  if (n_start == 0 || n_end == 0) return QPSK_ERR_ARG;       // :60-61
  if (n_payload > 0 && !payload) return QPSK_ERR_NULL;
  ModEngine& e = m->eng;
  int64_t ff = 0;
  QPSK_TRY(mod_frames_common(m, nullptr, n_payload, 1, start_marker, n_start, end_marker, n_end, pulse_shaping, nullptr, 0, &ff,
                             nullptr));
  *n_floats = ff;
  if (ff == 0 || !iq_out) return QPSK_OK;
  if (cap_floats < ff) return QPSK_ERR_CAPACITY;
  QPSK_TRY(ensure_device(m->eng.device));
  cudaStream_t s = e.stream;
  QPSK_TRY(e.d_src.ensure((size_t)(n_payload > 0 ? n_payload : 1)));
  QPSK_TRY(e.d_out.ensure((size_t)(ff >> 1)));
  if (n_payload > 0) QPSK_CUDA_TRY(cudaMemcpyAsync(e.d_src.p, payload, (size_t)n_payload, cudaMemcpyHostToDevice, s));
  QPSK_TRY(mod_frames_common(m, e.d_src.p, n_payload, 1, start_marker, n_start, end_marker, n_end, pulse_shaping, e.d_out.p,
                             ff >> 1, &ff, s));
  QPSK_CUDA_TRY(cudaMemcpyAsync(iq_out, e.d_out.p, (size_t)ff * 4, cudaMemcpyDeviceToHost, s));
  QPSK_CUDA_TRY(cudaStreamSynchronize(s));
  return QPSK_OK;
}

// ModulateBytes over a batch of frames, host memory in and out: payloads [frames][n_payload] -> samples
// [frames][out_stride_floats].  The output is 32*sps times the input (8 B per sample, 4*sps samples per payload byte), so
// the call is bound by the device-to-host copy: groups of frames (~32 MiB of samples) go through a 3-slot pipeline, the
// shaping kernel of group g+1 running while group g is copied out.
int qpsk_mod_modulate_frames(qpsk_mod* m, const uint8_t* payloads, int64_t n_payload, int frames, const uint8_t* start_marker,
                             int64_t n_start, const uint8_t* end_marker, int64_t n_end, float* iq_out, int64_t out_stride_floats,
                             int64_t* frame_floats) {
  if (!m) return QPSK_ERR_NULL;
  if (out_stride_floats & 1) return QPSK_ERR_ARG;
  int64_t ff = 0;
  QPSK_TRY(mod_frames_common(m, nullptr, n_payload, frames, start_marker, n_start, end_marker, n_end, 1, nullptr, 0, &ff, nullptr));
  if (frame_floats) *frame_floats = ff;
  if (ff == 0 || frames == 0 || !iq_out) return QPSK_OK;
  if (n_payload > 0 && !payloads) return QPSK_ERR_NULL;
  if (frames > 1 && out_stride_floats < ff) return QPSK_ERR_ARG;
  ModEngine& e = m->eng;
  QPSK_TRY(ensure_device(e.device));
  if (!m->s_out) {
    QPSK_CUDA_TRY(cudaStreamCreateWithFlags(&m->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < qpsk_mod::kSlots; ++i) {
      QPSK_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_k[i], cudaEventDisableTiming));
      QPSK_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  const int64_t fc = ff >> 1;                                 // complex samples per frame
  const int64_t stride_c = fc + (fc & 1);                     // device rows stay 16-byte aligned
  int64_t G = (32LL << 20) / (stride_c * 8);
  if (G < 1) G = 1;
  if (G > frames) G = frames;
  for (int b = 0; b < qpsk_mod::kSlots; ++b) {
    QPSK_TRY(m->d_grp[b].ensure((size_t)(G * stride_c)));
    QPSK_TRY(m->d_pay[b].ensure((size_t)(G * (n_payload > 0 ? n_payload : 1))));
  }
  cudaStream_t s = e.stream;
  // a pageable output block: an asynchronous copy into it blocks the calling thread, so the groups land in page-locked slots
  // and the host copy pool moves them on while the next groups are shaped and copied
  const bool bounce = host_ptr_is_pageable(iq_out);
  if (bounce) QPSK_TRY(m->landing((size_t)(G * stride_c)));
  auto drain = [&](int gi) -> int {
    const int b = gi % qpsk_mod::kSlots;
    const int64_t f0 = (int64_t)gi * G;
    const int nf = (int)((frames - f0 < G) ? (frames - f0) : G);
    QPSK_CUDA_TRY(cudaEventSynchronize(m->ev_d2h[b]));
    host_parallel_copy_rows(iq_out + (size_t)f0 * out_stride_floats, (size_t)out_stride_floats * 4, m->h_grp[b], (size_t)stride_c * 8,
                            (size_t)fc * 8, (size_t)nf);
    return QPSK_OK;
  };
  int g = 0;
  for (int64_t f0 = 0; f0 < frames; f0 += G, ++g) {
    const int b = g % qpsk_mod::kSlots;
    const int nf = (int)((frames - f0 < G) ? (frames - f0) : G);
    if (g >= qpsk_mod::kSlots) QPSK_CUDA_TRY(cudaStreamWaitEvent(s, m->ev_d2h[b], 0));   // the slot's previous copy-out is done
    if (n_payload > 0)
      QPSK_CUDA_TRY(cudaMemcpyAsync(m->d_pay[b].p, payloads + (size_t)f0 * n_payload, (size_t)nf * n_payload, cudaMemcpyHostToDevice, s));
    int64_t ff2 = 0;
    QPSK_TRY(mod_frames_common(m, m->d_pay[b].p, n_payload, nf, start_marker, n_start, end_marker, n_end, 1, m->d_grp[b].p, stride_c,
                               &ff2, s));
    QPSK_CUDA_TRY(cudaEventRecord(m->ev_k[b], s));
    QPSK_CUDA_TRY(cudaStreamWaitEvent(m->s_out, m->ev_k[b], 0));
    if (bounce) {
      QPSK_CUDA_TRY(cudaMemcpyAsync(m->h_grp[b], m->d_grp[b].p, (size_t)nf * stride_c * 8, cudaMemcpyDeviceToHost, m->s_out));
    } else {
      QPSK_CUDA_TRY(cudaMemcpy2DAsync(iq_out + (size_t)f0 * out_stride_floats, (size_t)out_stride_floats * 4, m->d_grp[b].p,
                                      (size_t)stride_c * 8, (size_t)fc * 8, (size_t)nf, cudaMemcpyDeviceToHost, m->s_out));
    }
    QPSK_CUDA_TRY(cudaEventRecord(m->ev_d2h[b], m->s_out));
    if (bounce && g >= 1) QPSK_TRY(drain(g - 1));             // group g - 1 leaves its slot two iterations before it is reused
  }
  if (bounce && g >= 1) QPSK_TRY(drain(g - 1));
  QPSK_CUDA_TRY(cudaStreamSynchronize(m->s_out));
  QPSK_CUDA_TRY(cudaStreamSynchronize(s));
  return QPSK_OK;
}

int qpsk_mod_modulate_frames_dev(qpsk_mod* m, const uint8_t* d_payloads, int64_t n_payload, int frames,
                                 const uint8_t* start_marker, int64_t n_start, const uint8_t* end_marker, int64_t n_end,
                                 float* d_iq_out, int64_t out_stride_floats, int64_t* frame_floats, void* stream) {
  if (!m) return QPSK_ERR_NULL;
  if (out_stride_floats & 1) return QPSK_ERR_ARG;
  cudaStream_t s = stream ? (cudaStream_t)stream : m->eng.stream;
  return mod_frames_common(m, d_payloads, n_payload, frames, start_marker, n_start, end_marker, n_end, 1, (float2*)d_iq_out,
                           out_stride_floats >> 1, frame_floats, s);
}

}  // extern "C"
