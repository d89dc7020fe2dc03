// tcgen05 fused SeparableConv2d kernel: host-side plan + launcher (sepconv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace migan {

// Everything one launch needs, resolved at plan time (tensor maps are encoded once per
// (batch size, workspace) and reused by every forward).
struct SepconvTcArgs {
    CUtensorMap map_in;    // fp32 NHWC activation  [n][H][W][cin]   (A_DW mode)
    CUtensorMap map_a_hi;  // fp16 [P][cin] pre-split A operand      (A_TMA mode)
    CUtensorMap map_a_lo;
    CUtensorMap map_w_hi;  // fp16 [cout][cin] K-major weights
    CUtensorMap map_w_lo;
    CUtensorMap map_out;   // fp32 NHWC output [n][H][W][cout]
    const float* w9;       // [9][cin] depthwise taps
    const float* bias;     // [cin]
    const float* noise;    // [H*W] or null
    float* out;
    float inv_scale;
    int n, res, cin, cout, act, passes;
    int a_mode;            // 0 = depthwise prologue from fp32 NHWC, 1 = pre-split fp16 via TMA
    int tile_h, tile_w, tile_n;  // spatial tile: tile_n images x tile_h x tile_w = 128 pixels
    int num_m_tiles, num_n_tiles;
    int variant;           // kernel template instance
    unsigned grid;
    unsigned smem_bytes;
};

// Returns nullptr on success, else a static error string.
const char* sepconv_tc_plan(SepconvTcArgs* a, int passes, const float* in_f32, const __half* a_hi, const __half* a_lo,
                            const float* w9, const float* bias, const __half* w_hi, const __half* w_lo,
                            float inv_scale, const float* noise, float* out, int n, int res, int cin, int cout, int act);
cudaError_t launch_sepconv_tc(const SepconvTcArgs& a, cudaStream_t s);

}  // namespace migan
