// bmfr_run — the reference's driver program rebuilt on the C ABI.
//
// Plays the role of tasks() in /root/reference/opencl/bmfr.cpp:179-556: "Initialize", load the input
// frames (here: the synth-v1 generator instead of .exr files, bmfr.cpp:259-307), run and profile the
// kernels frame by frame (bmfr.cpp:417-485), then print the per-kernel tables in the layout of
// clutils::ProfilingInfo::print (CLUtils.hpp:313-332) with the reference's frame-0 exclusion rules
// (bmfr.cpp:392-397,488-506): K1 / K4 / K5 / total over frames 1..N-1, K2 / K3 over frames 0..N-1.
//
//   bmfr_run [width height frames staged|fused]                      synthetic sequence (synth-v1)
//   bmfr_run --data DIR [--frames N] [--out DIR] [staged|fused]     a dataset in the reference's format:
//       DIR/{color,shading_normal,world_position,albedo}N.exr + DIR/camera_matrices.h (bmfr.cpp:43-52),
//       image size taken from color0.exr (the reference's TODO at bmfr.cpp:37), position / normal
//       limits and the per-frame matrices / pixel offsets parsed from the header; with --out the
//       frames are written as DIR/outputN.png (bmfr.cpp:52,520-539); with --truth PREFIX every output
//       frame is compared with the linear ground-truth image PREFIX + N + ".exr" after the display
//       transform of bmfr.cl:852-856 (PSNR and SSIM per frame and their means).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/bmfr_b200.h"
#include "../../include/bmfr_io.h"

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

#define CHECK_IO(call)                                                \
    do {                                                              \
        int _st = (call);                                             \
        if (_st != BMFR_IO_OK) {                                      \
            printf("Error %d: %s\n", _st, bmfr_io_last_error());      \
            return 1;                                                 \
        }                                                             \
    } while (0)

int main(int argc, char** argv) {
    int W = 1280, H = 720, frames = 60;  // bmfr.cpp:39-42
    bool staged = true;
    std::string data_dir, out_dir, truth_prefix;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--data" && i + 1 < argc) data_dir = argv[++i];
        else if (a == "--out" && i + 1 < argc) out_dir = argv[++i];
        else if (a == "--truth" && i + 1 < argc) truth_prefix = argv[++i];
        else if (a == "--frames" && i + 1 < argc) frames = atoi(argv[++i]);
        else if (a == "fused") staged = false;
        else if (a == "staged") staged = true;
        else pos.push_back(a);
    }
    if (data_dir.empty() && pos.size() >= 2) {
        W = atoi(pos[0].c_str());
        H = atoi(pos[1].c_str());
        if (pos.size() >= 3) frames = atoi(pos[2].c_str());
    }
    if (frames < 1) {
        printf("Error: no frames\n");
        return 1;
    }

    printf("Initialize.\n");
    bmfr_params prm;
    std::vector<float> matrices((size_t)frames * 16), offsets((size_t)frames * 2);
    if (!data_dir.empty()) {
        int channels = 0;
        CHECK_IO(bmfr_io_exr_info((data_dir + "/color0.exr").c_str(), &W, &H, &channels));
        bmfr_default_params(&prm, W, H);
        int nm = 0, no = 0;
        CHECK_IO(bmfr_io_parse_camera_header((data_dir + "/camera_matrices.h").c_str(), frames, matrices.data(), offsets.data(), &nm,
                                             &no, &prm.position_limit_squared, &prm.normal_limit_squared));
        if (nm < frames || no < frames) {
            printf("Error: camera_matrices.h holds %d matrices and %d pixel offsets, %d frames were asked for\n", nm, no, frames);
            return 1;
        }
    } else {
        bmfr_default_params(&prm, W, H);
    }
    prm.mode = staged ? BMFR_MODE_STAGED : BMFR_MODE_FUSED;
    prm.profile = 1;
    bmfr_ctx* ctx = nullptr;
    CHECK(bmfr_create(&prm, &ctx));

    printf("Loading input data.\n");
    const size_t n = (size_t)W * H * 3;
    std::vector<std::vector<float>> albedo(frames), normal(frames), position(frames), noisy(frames), out(frames);
    if (!data_dir.empty()) {  // bmfr.cpp:259-312
        bool error = false;
#pragma omp parallel for
        for (int f = 0; f < frames; ++f) {
            if (error) continue;
            albedo[f].resize(n); normal[f].resize(n); position[f].resize(n); noisy[f].resize(n); out[f].resize(n);
            const struct { const char* stem; float* dst; const char* what; } files[4] = {
                {"albedo", albedo[f].data(), "Albedo"}, {"shading_normal", normal[f].data(), "Normal"},
                {"world_position", position[f].data(), "Position"}, {"color", noisy[f].data(), "Noisy"}};
            for (const auto& file : files) {
                const std::string path = data_dir + "/" + file.stem + std::to_string(f) + ".exr";
                if (bmfr_io_read_exr_rgb(path.c_str(), W, H, file.dst) != BMFR_IO_OK) {
                    error = true;
                    printf("%s buffer loading failed, reason: %s\n", file.what, bmfr_io_last_error());
                    break;
                }
            }
        }
        if (error) {
            printf("One or more errors occurred during buffer loading\n");
            return 1;
        }
    } else {
        for (int f = 0; f < frames; ++f) {
            albedo[f].resize(n); normal[f].resize(n); position[f].resize(n); noisy[f].resize(n); out[f].resize(n);
            CHECK(bmfr_synth_frame_host(W, H, 0, H, f, 0x424D4652u, albedo[f].data(), normal[f].data(), position[f].data(),
                                        noisy[f].data()));
            bmfr_synth_camera(f, W, H, 0, &matrices[(size_t)f * 16], &offsets[(size_t)f * 2]);
        }
    }

    printf("Run and profile kernels.\n");
    for (int f = 0; f < frames; ++f) {
        const int matrix_index = f == 0 ? 0 : f - 1;  // camera_matrices[matrix_index], bmfr.cpp:440-442
        CHECK(bmfr_denoise_frame_host(ctx, f, albedo[f].data(), normal[f].data(), position[f].data(), noisy[f].data(),
                                      &matrices[(size_t)matrix_index * 16], &offsets[(size_t)f * 2],  // pixel_offsets[frame], :443-444
                                      out[f].data()));
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

    if (!truth_prefix.empty()) {
        std::vector<double> psnr(frames, 0.0), ssim(frames, 0.0);
        bool error = false;
#pragma omp parallel for
        for (int f = 0; f < frames; ++f) {
            if (error) continue;
            std::vector<float> truth(n);
            const std::string path = truth_prefix + std::to_string(f) + ".exr";
            if (bmfr_io_read_exr_rgb(path.c_str(), W, H, truth.data()) != BMFR_IO_OK) {
                printf("Ground truth loading failed, reason: %s\n", bmfr_io_last_error());
                error = true;
                continue;
            }
            bmfr_io_tone_map(truth.data(), n);
            if (bmfr_io_psnr(out[f].data(), truth.data(), n, 1.f, &psnr[f]) != BMFR_IO_OK ||
                bmfr_io_ssim_rgb(out[f].data(), truth.data(), W, H, 1.f, &ssim[f]) != BMFR_IO_OK) {
                printf("Quality metrics failed for frame %d, reason: %s\n", f, bmfr_io_last_error());
                error = true;
            }
        }
        if (error) return 1;
        double mp = 0, ms = 0;
        printf("\n Quality against %sN.exr (tone-mapped)\n", truth_prefix.c_str());
        for (int f = 0; f < frames; ++f) {
            printf("   frame %3d : PSNR %7.3f dB   SSIM %.5f\n", f, psnr[f], ssim[f]);
            mp += psnr[f];
            ms += ssim[f];
        }
        printf("   mean      : PSNR %7.3f dB   SSIM %.5f\n\n", mp / frames, ms / frames);
    }

    if (!out_dir.empty()) {  // "Store results", bmfr.cpp:519-553
        bool error = false;
#pragma omp parallel for
        for (int f = 0; f < frames; ++f) {
            if (error) continue;
            const std::string path = out_dir + "/output" + std::to_string(f) + ".png";
            if (bmfr_io_write_png_rgb(path.c_str(), W, H, out[f].data(), (size_t)W * 3) != BMFR_IO_OK) {
                printf("Can't create image file on disk to location %s\n", path.c_str());
                error = true;
            }
        }
        if (error) {
            printf("One or more errors occurred during image saving\n");
            return 1;
        }
    }
    return 0;
}
