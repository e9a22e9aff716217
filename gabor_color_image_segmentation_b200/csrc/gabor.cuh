// gabor.cuh — host-side description of a Gabor bank and its device tables.
#pragma once
#include <vector>

#include "common.cuh"

namespace gcis {

constexpr int GB_MAX_JOBS = 16;   // row-filter/column-filter jobs per scale
constexpr int GB_MAX_SCALES = 16;
constexpr int GB_TAP_PAD = 16;    // zero taps on each side of every 1-D filter

// One job = one complex row filter followed by one complex column filter.  A job emits the
// orientation `out0`, and — when theta' = pi - theta is also in the bank — its conjugate
// partner `out1` from the same four real convolution sums.
struct GaborJob {
    int h;                        // half-width (same for the row and column factor)
    int row_re, row_im;           // offsets of the padded tap arrays in the tap table; -1 = all zero
    int col_re, col_im;
    int out0, out1;               // orientation indices; out1 = -1 when unpaired
};

struct GaborScale {
    int n_jobs;
    int hmax;
    GaborJob jobs[GB_MAX_JOBS];
};

struct GaborBankHost {
    int S = 0, O = 0;
    int hmax = 0;                 // widest half-width in the bank
    std::vector<GaborScale> scales;
    std::vector<float> taps;      // padded tap arrays; entry [off + GB_TAP_PAD + t] is tap t (offset t-h)
    double flops_per_pixel_channel = 0;  // useful FMAs*2 per pixel per channel (for the roofline)
};

double gabor_sigma(double frequency, double bandwidth);
int gabor_half_width(double frequency, double theta, double bandwidth, double n_stds);
// gx, gy: 2h+1 complex taps each; returns h.
int gabor_separable(double frequency, double theta, double bandwidth, double n_stds, std::vector<double> &gx_re,
                    std::vector<double> &gx_im, std::vector<double> &gy_re, std::vector<double> &gy_im);
int build_bank(const double *freqs, int S, const double *thetas, int O, double bandwidth, double n_stds,
               GaborBankHost &out);

}  // namespace gcis
