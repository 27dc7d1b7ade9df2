// bmfr_run — the reference's driver program rebuilt on the C ABI.
//
// Plays the role of tasks() in /root/reference/opencl/bmfr.cpp:179-556: "Initialize", load the input
// frames (here: the synth-v1 generator instead of .exr files, bmfr.cpp:259-307), run and profile the
// kernels frame by frame (bmfr.cpp:417-485), then print the per-kernel tables in the layout of
// clutils::ProfilingInfo::print (CLUtils.hpp:313-332) with the reference's frame-0 exclusion rules
// (bmfr.cpp:392-397,488-506): K1 / K4 / K5 / total over frames 1..N-1, K2 / K3 over frames 0..N-1.
//
//   bmfr_run [width height frames staged|fused]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/bmfr_b200.h"

struct Table {  // same numbers and layout as ProfilingInfo (Mean / Min / Max / Total, 3 decimals)
    std::string label;
    std::vector<double> t;
    void print() const {
        const int width = 4 + (int)log10((double)std::max<size_t>(t.size(), 1));
        const double total = std::accumulate(t.begin(), t.end(), 0.0);
        printf("\n %s\n %s\n", label.c_str(), std::string(label.size(), '-').c_str());
        printf("   Mean   : %*.3f ms\n", width, total / (double)t.size());
        printf("   Min    : %*.3f ms\n", width, *std::min_element(t.begin(), t.end()));
        printf("   Max    : %*.3f ms\n", width, *std::max_element(t.begin(), t.end()));
        printf("   Total  : %*.3f ms\n\n", width, total);
    }
};

#define CHECK(call)                                                   \
    do {                                                              \
        int _st = (call);                                             \
        if (_st != BMFR_OK) {                                         \
            printf("Error %d: %s\n", _st, bmfr_last_error());         \
            return _st;                                               \
        }                                                             \
    } while (0)

int main(int argc, char** argv) {
    const int W = argc > 2 ? atoi(argv[1]) : 1280, H = argc > 2 ? atoi(argv[2]) : 720;  // bmfr.cpp:39-40
    const int frames = argc > 3 ? atoi(argv[3]) : 60;                                     // bmfr.cpp:42
    const bool staged = !(argc > 4 && strcmp(argv[4], "fused") == 0);

    printf("Initialize.\n");
    bmfr_params prm;
    bmfr_default_params(&prm, W, H);
    prm.mode = staged ? BMFR_MODE_STAGED : BMFR_MODE_FUSED;
    prm.profile = 1;
    bmfr_ctx* ctx = nullptr;
    CHECK(bmfr_create(&prm, &ctx));

    printf("Loading input data.\n");
    const size_t n = (size_t)W * H * 3;
    std::vector<std::vector<float>> albedo(frames), normal(frames), position(frames), noisy(frames), out(frames);
    for (int f = 0; f < frames; ++f) {
        albedo[f].resize(n); normal[f].resize(n); position[f].resize(n); noisy[f].resize(n); out[f].resize(n);
        CHECK(bmfr_synth_frame_host(W, H, 0, H, f, 0x424D4652u, albedo[f].data(), normal[f].data(), position[f].data(),
                                    noisy[f].data()));
    }

    printf("Run and profile kernels.\n");
    for (int f = 0; f < frames; ++f) {
        float cam[16], off[2], unused[16];
        bmfr_synth_camera(f == 0 ? 0 : f - 1, W, H, 0, cam, unused);  // camera_matrices[matrix_index], bmfr.cpp:440-442
        bmfr_synth_camera(f, W, H, 0, unused, off);                   // pixel_offsets[frame], bmfr.cpp:443-444
        CHECK(bmfr_denoise_frame_host(ctx, f, albedo[f].data(), normal[f].data(), position[f].data(), noisy[f].data(), cam,
                                      off, out[f].data()));
    }
    CHECK(bmfr_sync(ctx));

    Table t[6] = {{"Accumulation of noisy data", {}}, {"Fitting feature buffers to noisy data", {}}, {"Weighted sum", {}},
                  {"Accumulation of filtered data", {}}, {"TAA", {}},
                  {"Total time in all kernels (including intermediate launch overheads)", {}}};
    for (int f = 0; f < frames; ++f) {
        float ms[BMFR_STAGE_COUNT];
        CHECK(bmfr_get_stage_ms(ctx, f, ms));
        if (f > 0) {  // bmfr.cpp:491-503
            t[0].t.push_back(ms[BMFR_STAGE_ACCUM_NOISY]);
            t[3].t.push_back(ms[BMFR_STAGE_ACCUM_FILTERED]);
            t[4].t.push_back(ms[BMFR_STAGE_TAA]);
            t[5].t.push_back(ms[BMFR_STAGE_TOTAL]);
        }
        t[1].t.push_back(ms[BMFR_STAGE_FITTER]);  // bmfr.cpp:504-505
        t[2].t.push_back(ms[BMFR_STAGE_WEIGHTED_SUM]);
    }
    if (!staged) {
        t[1].label = "Accumulation of noisy data + fitting (fit_kernel)";
        t[4].label = "Weighted sum + accumulation of filtered data + TAA (post_kernel)";
    }
    if (frames > 1 && staged) t[0].print();  // bmfr.cpp:508-517
    t[1].print();
    if (staged) t[2].print();
    if (frames > 1) {
        if (staged) t[3].print();
        t[4].print();
        t[5].print();
    }
    double cs = 0;
    for (float v : out[frames - 1]) cs += v;
    printf("checksum of the last frame: %.6f (%d x %d, %d frames, %lld kernel launches)\n", cs, W, H, frames,
           bmfr_kernel_launches(ctx));
    bmfr_destroy(ctx);
    return 0;
}
