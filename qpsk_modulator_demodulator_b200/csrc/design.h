// design.h — host-side designers shared by the handle implementations.
#pragma once
#include <vector>

namespace qpsk {

constexpr float kPiF = 3.14159274101257324f;  // MathF.PI
constexpr float kTwoPiF = 2.0f * kPiF;        // Band-Edge Filter.cs:16

std::vector<double> design_rrc(double span_symbols, double beta, int sample_rate, int symbol_rate);
std::vector<float> real_taps_as_iq(const std::vector<double>& h);
void design_band_edge(float sps, float rolloff, int size, std::vector<float>& lower, std::vector<float>& upper);
void mm_gains(double bn, double* kp, double* ki);
void costas_gains(double fs, double bw_hz, double damping, double* alpha, double* beta);
bool blank_or_null(const char* s);

}  // namespace qpsk
