// uint8 pre/post-processing around the generator (scripts/demo.py:56-66 and :135-142):
//   pre : img u8 [n][R][R][3] + mask u8 [n][R][R] (255 = known)  ->  x fp32 NCHW [n][4][R][R] = cat([mask-0.5, img*mask])
//   post: y fp32 NCHW [n][3][R][R] -> (y*0.5+0.5).clamp(0,1)*255 -> u8 (truncation) -> composite with the known pixels,
//         u8 HWC [n][R][R][3].
// With these the host<->device traffic of a request is 4 + 3 bytes per pixel instead of 16 + 12.
// Item kernels (one pixel per thread) in the comod_kernels.cuh style, so the emulation build checks them on the CPU.
// Every product of the reference is rounded separately (no FMA contraction) to stay bit-exact with torch.
#include "comod_kernels.cuh"

#ifdef MIGAN_EMULATE
#define PP_MUL(a, b) ((a) * (b))
#define PP_ADD(a, b) ((a) + (b))
#define PP_DIV(a, b) ((a) / (b))
#else
#include "kernels.h"
#define PP_MUL(a, b) __fmul_rn((a), (b))
#define PP_ADD(a, b) __fadd_rn((a), (b))
#define PP_DIV(a, b) __fdiv_rn((a), (b))
#endif

namespace comod {

struct PreU8K {        // items = n*R*R
    const uint8_t* img; const uint8_t* mask; float* x; int64_t HW;
#ifdef MIGAN_EMULATE
    inline
#else
    __device__ __forceinline__
#endif
    void operator()(int64_t i) const {
        const int64_t n = i / HW, p = i - n * HW;
        const float m = mask[i] == 255 ? 1.f : 0.f;                        // mask // 255            (demo.py:60)
        float* xo = x + n * 4 * HW + p;
        xo[0] = PP_ADD(m, -0.5f);                                          // mask - 0.5             (:65)
        for (int c = 0; c < 3; ++c) {
            const float v = PP_ADD(PP_DIV(PP_MUL((float)img[i * 3 + c], 2.f), 255.f), -1.f);   // img * 2 / 255 - 1  (:61)
            xo[(c + 1) * HW] = PP_MUL(v, m);                               // img * mask             (:65)
        }
    }
};

struct PostU8K {       // items = n*R*R
    const float* y; const uint8_t* img; const uint8_t* mask; uint8_t* out; int64_t HW;
#ifdef MIGAN_EMULATE
    inline
#else
    __device__ __forceinline__
#endif
    void operator()(int64_t i) const {
        const int64_t n = i / HW, p = i - n * HW;
        const bool known = mask[i] == 255;
        for (int c = 0; c < 3; ++c) {
            float t = PP_ADD(PP_MUL(y[(n * 3 + c) * HW + p], 0.5f), 0.5f);  // y * 0.5 + 0.5         (demo.py:135)
            t = fminf(fmaxf(t, 0.f), 1.f);                                  // .clamp(0, 1)
            t = PP_MUL(t, 255.f);
            out[i * 3 + c] = known ? img[i * 3 + c] : (uint8_t)t;           // .to(uint8); img*mask + result*(1-mask)  (:136,140)
        }
    }
};

}  // namespace comod

namespace migan {
int launch_preprocess_u8(const uint8_t* img, const uint8_t* mask, float* x, int n, int r, ck_stream_t s) {
    comod::PreU8K k{img, mask, x, (int64_t)r * r};
    return (int)comod::ck_launch(k, (int64_t)n * r * r, s);
}
int launch_postprocess_u8(const float* y, const uint8_t* img, const uint8_t* mask, uint8_t* out, int n, int r, ck_stream_t s) {
    comod::PostU8K k{y, img, mask, out, (int64_t)r * r};
    return (int)comod::ck_launch(k, (int64_t)n * r * r, s);
}
}  // namespace migan

#ifdef MIGAN_EMULATE   // the product's extern "C" wrappers (with error strings) live in migan_abi.cu
extern "C" int b200_preprocess_u8(const uint8_t* img, const uint8_t* mask, float* x, int n, int r, void* s) {
    return migan::launch_preprocess_u8(img, mask, x, n, r, s);
}
extern "C" int b200_postprocess_u8(const float* y, const uint8_t* img, const uint8_t* mask, uint8_t* out, int n, int r, void* s) {
    return migan::launch_postprocess_u8(y, img, mask, out, n, r, s);
}
#endif
