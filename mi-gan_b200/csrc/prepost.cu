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

// Feathered composite of the deployed (ONNX) pipeline, scripts/create_onnx_pipeline.py:233-245, for a crop at the model
// resolution: blend weight = 5x5 smoothing (reflect border, kernel of :66-88 with kernel_size 5 / sigma 1) of the 3x3
// max-pooled mask, / 255;  out = clamp(image * w + ((y * 0.5 + 0.5) * 255).clamp(0, 255) * (1 - w), 0, 255) -> uint8.
// NCHW uint8 in / out like the pipeline.  The 25-tap sum of integer-valued mask levels is accumulated in fp64 and
// rounded once: an fp32 running sum gives 254.99998 instead of 255 on a fully known neighbourhood, which would turn every
// known pixel v into v - 1 after the uint8 truncation (torch's convolution returns exactly 255 there).
struct FeatherK {      // items = n*H*W
    const float* y; const uint8_t* img; const uint8_t* mask; uint8_t* out; int H, W; float k[25];
#ifdef MIGAN_EMULATE
    inline
#else
    __device__ __forceinline__
#endif
    void operator()(int64_t i) const {
        const int64_t HW = (int64_t)H * W;
        const int64_t n = i / HW, p = i - n * HW;
        const int h = (int)(p / W), w = (int)(p - (int64_t)h * W);
        const uint8_t* m = mask + n * HW;
        double acc = 0.0;
        for (int dy = -2; dy <= 2; ++dy) {
            int yy = h + dy;
            yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);            // reflect (no edge repeat), F.pad mode='reflect'
            for (int dx = -2; dx <= 2; ++dx) {
                int xx = w + dx;
                xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
                int mx = 0;                                                  // max_pool2d(3, stride 1, padding 1): window clipped to the image
                for (int a = -1; a <= 1; ++a) {
                    const int y2 = yy + a;
                    if (y2 < 0 || y2 >= H) continue;
                    for (int b = -1; b <= 1; ++b) {
                        const int x2 = xx + b;
                        if (x2 < 0 || x2 >= W) continue;
                        const int v = m[(int64_t)y2 * W + x2];
                        mx = v > mx ? v : mx;
                    }
                }
                acc += (double)k[(dy + 2) * 5 + (dx + 2)] * (double)mx;
            }
        }
        const float wgt = PP_DIV((float)acc, 255.f);                        // mask / 255                         (:240)
        const float inv = PP_ADD(1.f, -wgt);
        for (int c = 0; c < 3; ++c) {
            const int64_t o = (n * 3 + c) * HW + p;
            float g = PP_MUL(PP_ADD(PP_MUL(y[o], 0.5f), 0.5f), 255.f);       // ((y * 0.5 + 0.5) * 255)             (:234)
            g = fminf(fmaxf(g, 0.f), 255.f);
            const float v = PP_ADD(PP_MUL((float)img[o], wgt), PP_MUL(g, inv));   // image * mask + out * (1 - mask)   (:241)
            out[o] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);                   // .clamp(0, 255).to(uint8)             (:242)
        }
    }
};

}  // namespace comod

namespace migan {
int launch_feather_composite(const float* y, const uint8_t* img, const uint8_t* mask, uint8_t* out, int n, int H, int W,
                             const float* k25_host, ck_stream_t s) {
    comod::FeatherK k{y, img, mask, out, H, W, {0}};
    for (int i = 0; i < 25; ++i) k.k[i] = k25_host[i];
    return (int)comod::ck_launch(k, (int64_t)n * H * W, s);
}
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
extern "C" int b200_feather_composite(const float* y, const uint8_t* img, const uint8_t* mask, uint8_t* out, int n, int H, int W,
                                      const float* k25, void* s) {
    if (!k25) return 1;
    return migan::launch_feather_composite(y, img, mask, out, n, H, W, k25, s);
}
extern "C" int b200_preprocess_u8(const uint8_t* img, const uint8_t* mask, float* x, int n, int r, void* s) {
    return migan::launch_preprocess_u8(img, mask, x, n, r, s);
}
extern "C" int b200_postprocess_u8(const float* y, const uint8_t* img, const uint8_t* mask, uint8_t* out, int n, int r, void* s) {
    return migan::launch_postprocess_u8(y, img, mask, out, n, r, s);
}
#endif
