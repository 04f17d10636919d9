// stubs.cu — TEMPORARY: entry points not implemented yet return QPSK_ERR_UNSUPPORTED.
#include "common.cuh"
#define S(...) { return QPSK_ERR_UNSUPPORTED; }
extern "C" {
int qpsk_mod_create(int, int, double, int, int, const char*, qpsk_mod**) S()
int qpsk_mod_destroy(qpsk_mod*) S()
int qpsk_mod_taps(const qpsk_mod*, double*, int, int*) S()
int qpsk_mod_modulate_bits(qpsk_mod*, const char*, int64_t, int, float*, int64_t, int64_t*) S()
int qpsk_mod_modulate_bytes(qpsk_mod*, const uint8_t*, int64_t, const uint8_t*, int64_t, const uint8_t*, int64_t, int, float*, int64_t, int64_t*) S()
int qpsk_mod_modulate_frames_dev(qpsk_mod*, const uint8_t*, int64_t, int, const uint8_t*, int64_t, const uint8_t*, int64_t, float*, int64_t, int64_t*, void*) S()
int qpsk_demod_create(int, int, float, int, double, double, double, int, const char*, int, int64_t, qpsk_demod**) S()
int qpsk_demod_create_batch(int, int, float, int, double, double, double, int, const char*, int, int64_t, int, qpsk_demod**) S()
int qpsk_demod_destroy(qpsk_demod*) S()
int qpsk_demod_set_fir_mode(qpsk_demod*, int) S()
int qpsk_demod_bits(qpsk_demod*, const float*, int64_t, char*, int64_t, int64_t*) S()
int qpsk_demod_bytes(qpsk_demod*, const float*, int64_t, const uint8_t*, int64_t, const uint8_t*, int64_t, uint8_t*, int64_t, int64_t*) S()
int qpsk_demod_constellation(qpsk_demod*, const float*, int64_t, float*, int64_t, int64_t*) S()
int qpsk_demod_bits_dev(qpsk_demod*, const float*, int64_t, int64_t, uint8_t*, int64_t, int64_t*, void*) S()
int qpsk_demod_loop_state(qpsk_demod*, double*, double*, double*, double*, float*, float*) S()
int qpsk_chan_create(const qpsk_chan_params*, int, int, qpsk_chan**) S()
int qpsk_chan_destroy(qpsk_chan*) S()
int qpsk_chan_apply_dev(qpsk_chan*, const float*, int64_t, int64_t, float*, int64_t, void*) S()
int qpsk_chan_apply(qpsk_chan*, const float*, int64_t, float*) S()
int qpsk_fill_bytes_dev(uint64_t, int, int, int64_t, uint8_t*, void*) S()
int qpsk_ber_count_dev(const uint8_t*, int64_t, const int64_t*, const uint8_t*, int64_t, int64_t, int, uint32_t*, void*) S()
}
