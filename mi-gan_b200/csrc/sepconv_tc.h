// tcgen05 fused SeparableConv2d kernel: host-side plan + launcher (sepconv_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace migan {

// One resolved launch: tensor maps, tiling and pipeline depths are computed once per
// (batch size, workspace) at plan time and replayed by every forward.  The kernel parameter
// block is opaque here (defined in sepconv_tc.cu).
struct SepconvTcArgs {
    unsigned grid = 0;
    unsigned smem_bytes = 0;
    int num_tiles = 0;
    alignas(64) unsigned char params_blob[1152];
};

// Optional fused torgb + image path in the epilogue (cout <= 128 only).
struct SepconvTcRgb {
    const float* w;       // [3][cout]
    const float* b;       // [3]
    const float* fir;     // [16][3] taps of the image up-sampling (unused if img_lo == null)
    const float* img_lo;  // [n][3][res/2][res/2] planar or null
    float* img_out;       // [n][3][res][res] planar (may be overridden at launch: the caller's y)
    int store_out;        // 0: do not write the feature map (last block)
};

// in_f32 != null : A operand = act(dw3x3(in_f32) + bias), produced in the kernel prologue
// in_f32 == null : A operand = pre-split fp16 (a_hi, a_lo) [n*res*res][cin] loaded by TMA
// passes: 3 = fp16 hi/lo split (fp32-faithful), 1 = single fp16 pass.
// Returns nullptr on success, else an error string.
const char* sepconv_tc_plan(SepconvTcArgs* a, int passes, const float* in_f32, const __half* a_hi, const __half* a_lo,
                            const float* w9, const float* bias, const __half* w_hi, const __half* w_lo,
                            float inv_scale, const float* noise, float* out, int n, int res, int cin, int cout, int act,
                            const SepconvTcRgb* rgb = nullptr);
cudaError_t sepconv_tc_read_trace(unsigned long long* host_4096);
cudaError_t launch_sepconv_tc(const SepconvTcArgs& a, cudaStream_t s, float* img_out_override = nullptr);

}  // namespace migan
