// stream.cu — streaming front-end: the receive loop of TB/SDR/ModDemodOverSDR.cs:116-183
// reads one MTU of cf32 samples from the radio into a caller-owned buffer, calls
// DeModulateTextUtf8(span, start, end) on it (:136) and sleeps until the next block is due (:157-176).
// qpsk_stream keeps that call order and its results — block k's payload is what the k-th DeModulateBytes call
// returns — but never makes the caller wait for the GPU:
//   push   copies the block into a pinned staging slot (the radio buffer is free again when push returns),
//          enqueues H2D on a copy stream, then [CS16 -> cf32,] the demodulator chain + framer and the D2H of
//          the payload on the compute stream, and returns;
//   poll   hands out finished blocks' payloads in push order.
// `depth` slots rotate, so the PCIe copy of block k+1 overlaps the kernels of block k, and the loop state lives
// in the demodulator handle exactly as in the per-call path.
// CS16 (interleaved int16 I,Q — the format SaveAsCs16 writes, MS/Models/HelperFunctions.cs:75-106, and SDRs
// deliver natively) halves the PCIe bytes per sample; it is widened to cf32 on the device.
#include <deque>

#include "common.cuh"

namespace qpsk {

// out[k] = in[k] * scale (int16 -> float), 8 values per thread
__global__ void cs16_to_cf32_kernel(const int16_t* __restrict__ in, long long n, float scale, float* __restrict__ out) {
  const long long n8 = n >> 3;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (aligned) {
    const int4* in4 = reinterpret_cast<const int4*>(in);
    float4* out4 = reinterpret_cast<float4*>(out);
    for (long long i = tid; i < n8; i += stride) {
      const int4 v = in4[i];
      const int w[4] = {v.x, v.y, v.z, v.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f[2 * j] = (float)(short)(w[j] & 0xffff) * scale;
        f[2 * j + 1] = (float)(short)(w[j] >> 16) * scale;
      }
      out4[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
      out4[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (long long k = 8 * n8 + tid; k < n; k += stride) out[k] = (float)in[k] * scale;
  } else {
    for (long long k = tid; k < n; k += stride) out[k] = (float)in[k] * scale;
  }
}

// SaveAsCs16 pass 1 (:83-90): maxVal = max(|re|, |im|) over the buffer; non-negative floats order like their bit
// patterns, so one atomicMax on the int view per block suffices.
__global__ void absmax_kernel(const float* __restrict__ x, long long n, int* __restrict__ max_bits) {
  float m = 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(x[k]);
    if (a > m) m = a;                     // NaN compares false, like Math.Abs(NaN) > maxVal
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, m, o);
    if (t > m) m = t;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(max_bits, __float_as_int(m));
}

// SaveAsCs16 pass 2 (:97-106): (short) max(short.MinValue, min(short.MaxValue, v / maxVal * short.MaxValue)) in fp64;
// the C# (short) cast of a double truncates toward zero.
__global__ void cf32_to_cs16_kernel(const float* __restrict__ x, long long n, const int* __restrict__ max_bits,
                                    int16_t* __restrict__ out) {
  double maxv = (double)__int_as_float(*max_bits);
  if (maxv < 1e-12) maxv = 1.0;           // :92
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
    double v = (double)x[k] / maxv * 32767.0;
    v = fmin(32767.0, v);
    v = fmax(-32768.0, v);
    out[k] = (int16_t)(int)v;             // cvt.rzi
  }
}

// rows of `L` complex samples: out[c][n] = (float)in[c][n] * scale (one int16 pair -> one float2 per thread and pass)
__global__ void cs16_rows_to_cf32_kernel(const int16_t* __restrict__ in, long long ld_in, float scale, float2* __restrict__ out,
                                         long long ld_out, long long L, int C) {
  const long long total = (long long)C * L;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / L);
    const long long n = idx - (long long)c * L;
    const short2 v = reinterpret_cast<const short2*>(in + 2 * (long long)c * ld_in)[n];
    out[(long long)c * ld_out + n] = make_float2((float)v.x * scale, (float)v.y * scale);
  }
}

inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = 16LL * device_sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int cs16_to_cf32_launch(const int16_t* d_in, long long ld_in, float scale, float2* d_out, long long ld_out, long long L, int channels,
                        cudaStream_t s) {
  if (L <= 0 || channels <= 0) return QPSK_OK;
  cs16_rows_to_cf32_kernel<<<grid_for((long long)channels * L, 256), 256, 0, s>>>(d_in, ld_in, scale, d_out, ld_out, L, channels);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

}  // namespace qpsk

using namespace qpsk;

struct qpsk_stream {
  qpsk_demod* demod = nullptr;
  int device = 0;
  int depth = 0;
  int64_t max_block_floats = 0, max_payload = 0;
  std::vector<uint8_t> start, end;
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  struct Slot {
    void* h_in = nullptr;          // pinned, max_block_floats * 4 bytes (cf32) — CS16 uses the first half
    float* d_in = nullptr;         // device cf32 block
    int16_t* d_raw = nullptr;      // device CS16 block
    uint8_t* d_payload = nullptr;
    long long* d_n = nullptr;
    uint8_t* h_payload = nullptr;  // pinned
    long long* h_n = nullptr;      // pinned
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr;
    bool in_flight = false;
  };
  std::vector<Slot> slots;
  int64_t pushed = 0, polled = 0;                      // block sequence numbers
  std::deque<std::vector<uint8_t>> spill;              // results of slots that had to be reused before they were polled
  std::deque<long long> spill_n;
  int64_t spill_first = 0;                             // sequence number of spill.front()

  ~qpsk_stream() {
    for (auto& s : slots) {
      if (s.ev_done) cudaEventSynchronize(s.ev_done);
      if (s.h_in) cudaFreeHost(s.h_in);
      if (s.d_in) cudaFree(s.d_in);
      if (s.d_raw) cudaFree(s.d_raw);
      if (s.d_payload) cudaFree(s.d_payload);
      if (s.d_n) cudaFree(s.d_n);
      if (s.h_payload) cudaFreeHost(s.h_payload);
      if (s.h_n) cudaFreeHost(s.h_n);
      if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
      if (s.ev_done) cudaEventDestroy(s.ev_done);
    }
    if (s_copy) cudaStreamDestroy(s_copy);
    if (s_comp) cudaStreamDestroy(s_comp);
  }
};

namespace {

// make slot `i` reusable: wait for its block and park the result if nobody polled it yet
int retire_slot(qpsk_stream* st, int i, int64_t seq_of_slot) {
  qpsk_stream::Slot& s = st->slots[(size_t)i];
  if (!s.in_flight) return QPSK_OK;
  QPSK_CUDA_TRY(cudaEventSynchronize(s.ev_done));
  if (seq_of_slot >= st->polled) {
    // results leave in order: everything older is already in the spill queue (slots are reused in order)
    if (st->spill.empty()) st->spill_first = seq_of_slot;
    const long long n = *s.h_n;
    const long long keep = n < st->max_payload ? n : st->max_payload;
    st->spill.emplace_back(s.h_payload, s.h_payload + (keep > 0 ? keep : 0));
    st->spill_n.push_back(n);
  }
  s.in_flight = false;
  return QPSK_OK;
}

int stream_push_common(qpsk_stream* st, const void* data, int64_t n_items, bool cs16, float scale) {
  if (!st) return QPSK_ERR_NULL;
  if (n_items < 0) return QPSK_ERR_RANGE;
  if ((n_items & 1) != 0) return QPSK_ERR_ARG;               // interleaved IQ (QPSKDeModulator.cs:347-348)
  if (n_items > st->max_block_floats) return QPSK_ERR_CAPACITY;
  if (n_items > 0 && !data) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(st->device));
  const int i = (int)(st->pushed % st->depth);
  qpsk_stream::Slot& s = st->slots[(size_t)i];
  QPSK_TRY(retire_slot(st, i, st->pushed - st->depth));
  const size_t bytes = (size_t)n_items * (cs16 ? 2 : 4);
  if (bytes) memcpy(s.h_in, data, bytes);                     // the caller's radio buffer is free again after this
  if (bytes) {
    QPSK_CUDA_TRY(cudaMemcpyAsync(cs16 ? (void*)s.d_raw : (void*)s.d_in, s.h_in, bytes, cudaMemcpyHostToDevice, st->s_copy));
  }
  QPSK_CUDA_TRY(cudaEventRecord(s.ev_h2d, st->s_copy));
  QPSK_CUDA_TRY(cudaStreamWaitEvent(st->s_comp, s.ev_h2d, 0));
  if (cs16 && n_items > 0) {
    cs16_to_cf32_kernel<<<grid_for(n_items / 8 + 1, 256), 256, 0, st->s_comp>>>(s.d_raw, n_items, scale, s.d_in);
    QPSK_LAUNCH_CHECK();
  }
  QPSK_TRY(qpsk_demod_bytes_dev(st->demod, s.d_in, n_items, n_items, st->start.data(), (int64_t)st->start.size(), st->end.data(),
                                (int64_t)st->end.size(), s.d_payload, st->max_payload, (int64_t*)s.d_n, st->s_comp));
  QPSK_CUDA_TRY(cudaMemcpyAsync(s.h_n, s.d_n, sizeof(long long), cudaMemcpyDeviceToHost, st->s_comp));
  QPSK_CUDA_TRY(cudaMemcpyAsync(s.h_payload, s.d_payload, (size_t)st->max_payload, cudaMemcpyDeviceToHost, st->s_comp));
  QPSK_CUDA_TRY(cudaEventRecord(s.ev_done, st->s_comp));
  s.in_flight = true;
  ++st->pushed;
  return QPSK_OK;
}

}  // namespace

extern "C" {

int qpsk_stream_create(qpsk_demod* d, int64_t max_block_floats, int64_t max_payload_bytes, int depth, const uint8_t* start_marker,
                       int64_t n_start, const uint8_t* end_marker, int64_t n_end, qpsk_stream** out) {
  if (!d || !out) return QPSK_ERR_NULL;
  *out = nullptr;
  if (n_start == 0 || n_end == 0) return QPSK_ERR_ARG;        // QPSKDeModulator.cs:174-175
  if (!start_marker || !end_marker) return QPSK_ERR_NULL;
  if (max_block_floats < 2 || (max_block_floats & 1) || max_payload_bytes <= 0 || depth < 1 || depth > 64 || n_start < 0 || n_end < 0)
    return QPSK_ERR_RANGE;
  int ch = 0;
  QPSK_TRY(qpsk_demod_channels(d, &ch));
  if (ch != 1) return QPSK_ERR_UNSUPPORTED;                   // one radio stream per front-end
  int dev = 0;
  QPSK_TRY(qpsk_demod_device(d, &dev));                       // the front-end lives on the demodulator's device
  QPSK_TRY(ensure_device(dev));
  qpsk_stream* st = new (std::nothrow) qpsk_stream();
  if (!st) return QPSK_ERR_NOMEM;
  st->device = dev;
  st->demod = d;
  st->depth = depth;
  st->max_block_floats = max_block_floats;
  st->max_payload = max_payload_bytes;
  st->start.assign(start_marker, start_marker + n_start);
  st->end.assign(end_marker, end_marker + n_end);
  st->slots.resize((size_t)depth);
  auto fail = [&](int code) { delete st; return code; };
#define QPSK_S_TRY(expr)                                          \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) {                                      \
      ::qpsk::set_cuda_error(_e, #expr, __FILE__, __LINE__);      \
      return fail(_e == cudaErrorMemoryAllocation ? QPSK_ERR_NOMEM : QPSK_ERR_CUDA); \
    }                                                             \
  } while (0)
  QPSK_S_TRY(cudaStreamCreateWithFlags(&st->s_copy, cudaStreamNonBlocking));
  QPSK_S_TRY(cudaStreamCreateWithFlags(&st->s_comp, cudaStreamNonBlocking));
  for (auto& s : st->slots) {
    QPSK_S_TRY(cudaHostAlloc(&s.h_in, (size_t)max_block_floats * 4, cudaHostAllocPortable));
    QPSK_S_TRY(cudaMalloc((void**)&s.d_in, (size_t)max_block_floats * 4));
    QPSK_S_TRY(cudaMalloc((void**)&s.d_raw, (size_t)max_block_floats * 2));
    QPSK_S_TRY(cudaMalloc((void**)&s.d_payload, (size_t)max_payload_bytes));
    QPSK_S_TRY(cudaMalloc((void**)&s.d_n, sizeof(long long)));
    QPSK_S_TRY(cudaHostAlloc((void**)&s.h_payload, (size_t)max_payload_bytes, cudaHostAllocPortable));
    QPSK_S_TRY(cudaHostAlloc((void**)&s.h_n, sizeof(long long), cudaHostAllocPortable));
    QPSK_S_TRY(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
    QPSK_S_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
  }
#undef QPSK_S_TRY
  *out = st;
  return QPSK_OK;
}

int qpsk_stream_destroy(qpsk_stream* st) {
  if (st) {
    cudaSetDevice(st->device);
    delete st;
  }
  return QPSK_OK;
}

int qpsk_stream_push(qpsk_stream* st, const float* iq, int64_t n_floats) {
  return stream_push_common(st, iq, n_floats, false, 1.0f);
}

int qpsk_stream_push_cs16(qpsk_stream* st, const int16_t* iq, int64_t n_int16, float scale) {
  return stream_push_common(st, iq, n_int16, true, scale);
}

int qpsk_stream_pending(qpsk_stream* st, int64_t* pushed, int64_t* polled) {
  if (!st) return QPSK_ERR_NULL;
  if (pushed) *pushed = st->pushed;
  if (polled) *polled = st->polled;
  return QPSK_OK;
}

int qpsk_stream_poll(qpsk_stream* st, int wait, uint8_t* payload_out, int64_t cap, int64_t* n_bytes, int* have_block) {
  if (!st || !n_bytes || !have_block) return QPSK_ERR_NULL;
  *n_bytes = 0;
  *have_block = 0;
  if (st->polled >= st->pushed) return QPSK_OK;              // nothing outstanding
  QPSK_TRY(ensure_device(st->device));
  const uint8_t* src = nullptr;
  long long n = 0;
  bool from_spill = false;
  if (!st->spill.empty() && st->spill_first == st->polled) {
    src = st->spill.front().data();
    n = st->spill_n.front();
    from_spill = true;
  } else {
    qpsk_stream::Slot& s = st->slots[(size_t)(st->polled % st->depth)];
    if (wait) {
      QPSK_CUDA_TRY(cudaEventSynchronize(s.ev_done));
    } else {
      const cudaError_t q = cudaEventQuery(s.ev_done);
      if (q == cudaErrorNotReady) return QPSK_OK;            // block still in flight
      QPSK_CUDA_TRY(q);
    }
    src = s.h_payload;
    n = *s.h_n;
  }
  *n_bytes = n;
  *have_block = 1;
  int status = QPSK_OK;
  if (n > st->max_payload || n > cap) status = QPSK_ERR_CAPACITY;   // the block is consumed either way, like the per-call path
  else if (n > 0) {
    if (!payload_out) return QPSK_ERR_NULL;
    memcpy(payload_out, src, (size_t)n);
  }
  if (from_spill) {
    st->spill.pop_front();
    st->spill_n.pop_front();
    ++st->spill_first;
  } else {
    st->slots[(size_t)(st->polled % st->depth)].in_flight = false;
  }
  ++st->polled;
  return status;
}

int qpsk_stream_flush(qpsk_stream* st) {
  if (!st) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device(st->device));
  QPSK_CUDA_TRY(cudaStreamSynchronize(st->s_copy));
  QPSK_CUDA_TRY(cudaStreamSynchronize(st->s_comp));
  return QPSK_OK;
}

// ---- CS16 <-> cf32 --------------------------------------------------------------------------------
int qpsk_cs16_to_cf32_dev(const int16_t* d_in, int64_t n_int16, float scale, float* d_out, void* stream) {
  if (n_int16 < 0) return QPSK_ERR_RANGE;
  if ((n_int16 & 1) != 0) return QPSK_ERR_ARG;
  if (n_int16 == 0) return QPSK_OK;
  if (!d_in || !d_out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  cs16_to_cf32_kernel<<<grid_for(n_int16 / 8 + 1, 256), 256, 0, (cudaStream_t)stream>>>(d_in, n_int16, scale, d_out);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_cf32_to_cs16_dev(const float* d_in, int64_t n_floats, int16_t* d_out, float* d_max_abs, void* stream) {
  if (n_floats < 0) return QPSK_ERR_RANGE;
  if ((n_floats & 1) != 0) return QPSK_ERR_ARG;
  if (n_floats == 0) return QPSK_ERR_ARG;                     // "IQ array is empty." :79-80
  if (!d_in || !d_out || !d_max_abs) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  cudaStream_t s = (cudaStream_t)stream;
  QPSK_CUDA_TRY(cudaMemsetAsync(d_max_abs, 0, sizeof(float), s));
  absmax_kernel<<<grid_for(n_floats, 256), 256, 0, s>>>(d_in, n_floats, reinterpret_cast<int*>(d_max_abs));
  QPSK_LAUNCH_CHECK();
  cf32_to_cs16_kernel<<<grid_for(n_floats, 256), 256, 0, s>>>(d_in, n_floats, reinterpret_cast<const int*>(d_max_abs), d_out);
  QPSK_LAUNCH_CHECK();
  return QPSK_OK;
}

int qpsk_cs16_to_cf32(const int16_t* in, int64_t n_int16, float scale, float* out) {
  if (n_int16 < 0) return QPSK_ERR_RANGE;
  if ((n_int16 & 1) != 0) return QPSK_ERR_ARG;
  if (n_int16 == 0) return QPSK_OK;
  if (!in || !out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  DevBuf<int16_t> di;
  DevBuf<float> dout;
  QPSK_TRY(di.alloc((size_t)n_int16));
  QPSK_TRY(dout.alloc((size_t)n_int16));
  QPSK_CUDA_TRY(cudaMemcpy(di.p, in, (size_t)n_int16 * 2, cudaMemcpyHostToDevice));
  QPSK_TRY(qpsk_cs16_to_cf32_dev(di.p, n_int16, scale, dout.p, nullptr));
  QPSK_CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)n_int16 * 4, cudaMemcpyDeviceToHost));
  return QPSK_OK;
}

int qpsk_cf32_to_cs16(const float* iq, int64_t n_floats, int16_t* out, float* max_abs) {
  if (n_floats < 0) return QPSK_ERR_RANGE;
  if ((n_floats & 1) != 0 || n_floats == 0) return QPSK_ERR_ARG;
  if (!iq || !out) return QPSK_ERR_NULL;
  QPSK_TRY(ensure_device());
  DevBuf<float> di, dm;
  DevBuf<int16_t> dout;
  QPSK_TRY(di.alloc((size_t)n_floats));
  QPSK_TRY(dm.alloc(1));
  QPSK_TRY(dout.alloc((size_t)n_floats));
  QPSK_CUDA_TRY(cudaMemcpy(di.p, iq, (size_t)n_floats * 4, cudaMemcpyHostToDevice));
  QPSK_TRY(qpsk_cf32_to_cs16_dev(di.p, n_floats, dout.p, dm.p, nullptr));
  QPSK_CUDA_TRY(cudaMemcpy(out, dout.p, (size_t)n_floats * 2, cudaMemcpyDeviceToHost));
  if (max_abs) QPSK_CUDA_TRY(cudaMemcpy(max_abs, dm.p, sizeof(float), cudaMemcpyDeviceToHost));
  return QPSK_OK;
}

}  // extern "C"
