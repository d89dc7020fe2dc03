// Host build of the shared-memory stage functions of the fused SeparableConv2d kernel (TEST INFRASTRUCTURE ONLY).
// mi-gan_b200/csrc/sepconv_stages.cuh is compiled as plain C++ (-DMIGAN_EMULATE); each entry point loops the 128
// workers of one prologue group over host buffers laid out exactly like the kernel's pipeline stage.  The package never
// loads this library.
#include "sepconv_stages.cuh"

using namespace migan::stages;

extern "C" {

// depthwise 3x3 + act -> A operand half (tile shape selected like the kernel does)
void emul_prologue_chunk(int tile_w, const float* in_stage, uint8_t* a_hi, uint8_t* a_lo, const float* w9, const float* bias,
                         int cin, int cg0, int g) {
    for (int tg = 0; tg < 128; ++tg) {
        const f4* sin = reinterpret_cast<const f4*>(in_stage);
        if (tile_w == 16) prologue_chunk<1, 8, 16>(sin, a_hi, a_lo, w9, bias, cin, cg0, g, tg);
        else if (tile_w == 8) prologue_chunk<2, 8, 8>(sin, a_hi, a_lo, w9, bias, cin, cg0, g, tg);
        else prologue_chunk<8, 4, 4>(sin, a_hi, a_lo, w9, bias, cin, cg0, g, tg);
    }
}

void emul_prestage_up(float* in_stage, const float* t_area, const float* nz_area, const float* taps16, int y0, int x0, int R,
                      int has_noise) {
    UpTaps taps;
    for (int i = 0; i < 16; ++i) taps.f[i] = taps16[i];
    for (int tg = 0; tg < 128; ++tg)
        prestage_up(reinterpret_cast<f4*>(in_stage), reinterpret_cast<const f4*>(t_area), nz_area, taps, y0, x0, R, has_noise, tg);
}

void emul_prestage_stem(float* in_stage, const float* x_area, const float* ws, const float* bs, int cg0, int y0, int x0, int R) {
    for (int tg = 0; tg < 128; ++tg) prestage_stem(reinterpret_cast<f4*>(in_stage), x_area, ws, bs, cg0, y0, x0, R, tg);
}

}  // extern "C"
