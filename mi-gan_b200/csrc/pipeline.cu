// Arbitrary-resolution crop pipeline around the generator: the deployed (ONNX) form of the reference,
// scripts/create_onnx_pipeline.py:121-264 (MIGAN_Pipeline), SURVEY.md 8(f) row f4.
//
//   mask (any size) --nearest--> image size                                                   (:255)
//   columns / rows that contain a hole (value < 255) -> flags -> [host] crop window            (:133-227, migan_crop_box)
//   crop --anti-aliased bilinear, round--> res x res uint8 -> x = cat([m/255 - 0.5, (v*2/255-1) * m/255])   (:229-236)
//   [generator forward]
//   y -> ((y*0.5+0.5)*255).clamp(0,255) --anti-aliased bilinear--> crop size -> feathered blend -> image[crop]  (:238-262)
//
// The resize is what torchvision's tensor `resize` lowers to: ATen's separable anti-aliased bilinear filter
// (aten/src/ATen/native/cpu/UpSampleKernel.cpp: HelperInterpBase::_compute_indices_min_size_weights_aa and
// basic_loop_aa_*): width pass, then height pass, each output = sum_j src[xmin + j] * w[j] accumulated left to
// right with fused multiply-adds in fp32; the weight table is evaluated exactly like the C++ (float scale / support /
// centre, the filter argument through double, w / total).  A pass whose size does not change is the identity there;
// here it runs with the table {xmin = i, 1 tap, w = 1}, which is the same value bit for bit.
// Item kernels (comod_kernels.cuh style): the emulation build runs the very same functors on the CPU against the oracle.
#include "comod_kernels.cuh"

#ifdef MIGAN_EMULATE
#include <cmath>
#define PL_MUL(a, b) ((a) * (b))
#define PL_ADD(a, b) ((a) + (b))
#define PL_DIV(a, b) ((a) / (b))
#define PL_FMA(a, b, c) fmaf((a), (b), (c))
#define PL_RINT(a) rintf(a)
#define PL_D inline
#else
#include "kernels.h"
#define PL_MUL(a, b) __fmul_rn((a), (b))
#define PL_ADD(a, b) __fadd_rn((a), (b))
#define PL_DIV(a, b) __fdiv_rn((a), (b))
#define PL_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define PL_RINT(a) rintf(a)
#define PL_D __device__ __forceinline__
#endif

namespace comod {

// torch's mode='nearest' source index (ATen/native/UpSample.h nearest_neighbor_compute_source_index)
PL_D int nearest_src(int dst, int in_size, int out_size) {
    if (in_size == out_size) return dst;
    const float scale = PL_DIV((float)in_size, (float)out_size);
    const int s = (int)floorf(PL_MUL((float)dst, scale));
    return s < in_size - 1 ? s : in_size - 1;
}

struct NearestU8K {    // items = oh*ow: one uint8 plane [H][W] -> [oh][ow]
    const uint8_t* in; uint8_t* out; int H, W, oh, ow;
    PL_D void operator()(int64_t i) const {
        const int y = (int)(i / ow), x = (int)(i - (int64_t)y * ow);
        out[i] = in[(int64_t)nearest_src(y, H, oh) * W + nearest_src(x, W, ow)];
    }
};

struct HoleFlagsK {    // items = W + H: flags[x] = column x contains a value < 255, flags[W + y] likewise for row y  (:144-149)
    const uint8_t* mask; uint8_t* flags; int H, W;
    PL_D void operator()(int64_t i) const {
        uint8_t f = 0;
        if (i < W) { for (int y = 0; y < H && !f; ++y) f = mask[(int64_t)y * W + i] < 255; }
        else { const uint8_t* r = mask + (i - W) * (int64_t)W; for (int x = 0; x < W && !f; ++x) f = r[x] < 255; }
        flags[i] = f;
    }
};

// One axis of the anti-aliased bilinear filter: for output index i the first source index, the tap count and the taps.
struct AaTable { int* xmin; int* xsize; float* w; int maxk; };

struct AaWeightsK {    // items = out_size
    AaTable t; int in_size, out_size;
    PL_D void operator()(int64_t i) const {
        float* w = t.w + i * t.maxk;
        if (in_size == out_size) {                          // the reference skips this pass: identity table
            t.xmin[i] = (int)i; t.xsize[i] = 1; w[0] = 1.f;
            for (int j = 1; j < t.maxk; ++j) w[j] = 0.f;
            return;
        }
        const float scale = PL_DIV((float)in_size, (float)out_size);           // area_pixel_compute_scale<float>
        const bool down = scale >= 1.0f;
        const float support = down ? scale : 1.f;                              // (interp_size * 0.5) * scale, interp_size = 2
        const float invscale = down ? (float)(1.0 / (double)scale) : 1.f;
        const float center = (float)((double)scale * ((double)i + 0.5));
        int lo = (int)((double)PL_ADD(center, -support) + 0.5);
        if (lo < 0) lo = 0;
        int hi = (int)((double)PL_ADD(center, support) + 0.5);
        if (hi > in_size) hi = in_size;
        int n = hi - lo;
        n = n < 0 ? 0 : (n > t.maxk ? t.maxk : n);
        float total = 0.f;
        for (int j = 0; j < n; ++j) {
            float a = (float)(((double)PL_ADD((float)(j + lo), -center) + 0.5) * (double)invscale);
            a = fabsf(a);
            const float v = a < 1.0f ? PL_ADD(1.f, -a) : 0.f;                  // aa_filter (triangle)
            w[j] = v;
            total = PL_ADD(total, v);
        }
        if (total != 0.f)
            for (int j = 0; j < n; ++j) w[j] = PL_DIV(w[j], total);
        for (int j = n; j < t.maxk; ++j) w[j] = 0.f;
        t.xmin[i] = lo; t.xsize[i] = n;
    }
};

// Width pass.  SRC = 0: uint8 planes of the image crop; SRC = 1: generator output y mapped to [0, 255] on load (:239).
template <int SRC>
struct ResizeWidthK {  // items = C * rows * ow -> tmp [C][rows][ow]
    const void* src; int64_t plane_stride; int row_stride; AaTable t; float* out; int rows, ow;
    PL_D float load(int64_t o) const {
        if (SRC == 0) return (float)static_cast<const uint8_t*>(src)[o];
        float g = PL_MUL(PL_ADD(PL_MUL(static_cast<const float*>(src)[o], 0.5f), 0.5f), 255.f);   // ((y * 0.5 + 0.5) * 255)
        return fminf(fmaxf(g, 0.f), 255.f);                                                       // .clamp(0, 255)
    }
    PL_D void operator()(int64_t i) const {
        const int x = (int)(i % ow); int64_t r = i / ow;
        const int y = (int)(r % rows); const int c = (int)(r / rows);
        const int64_t base = c * plane_stride + (int64_t)y * row_stride + t.xmin[x];
        const float* w = t.w + (int64_t)x * t.maxk;
        float acc = PL_MUL(load(base), w[0]);
        for (int j = 1; j < t.xsize[x]; ++j) acc = PL_FMA(load(base + j), w[j], acc);
        out[i] = acc;
    }
};

struct ResizeHeightK { // items = C * oh * ow: tmp [C][ih][ow] -> out [C][oh][ow]
    const float* in; AaTable t; float* out; int ih, oh, ow;
    PL_D void operator()(int64_t i) const {
        const int x = (int)(i % ow); int64_t r = i / ow;
        const int y = (int)(r % oh); const int c = (int)(r / oh);
        const float* p = in + ((int64_t)c * ih + t.xmin[y]) * ow + x;
        const float* w = t.w + (int64_t)y * t.maxk;
        float acc = PL_MUL(p[0], w[0]);
        for (int j = 1; j < t.xsize[y]; ++j) acc = PL_FMA(p[(int64_t)j * ow], w[j], acc);
        out[i] = acc;
    }
};

// Height pass of the image crop fused with the rest of MIGAN_Pipeline.preprocess (:229-236): round to uint8, nearest
// mask, normalise, x = cat([mask - 0.5, image * mask]).
struct PreprocessK {   // items = res * res
    const float* in; AaTable t; const uint8_t* mask; int mask_row_stride; int hc, wc, res; float* x;
    PL_D void operator()(int64_t i) const {
        const int ox = (int)(i % res), oy = (int)(i / res);
        const int64_t plane = (int64_t)res * res;
        const float m = PL_DIV((float)mask[(int64_t)nearest_src(oy, hc, res) * mask_row_stride + nearest_src(ox, wc, res)], 255.f);
        x[i] = PL_ADD(m, -0.5f);
        const float* w = t.w + (int64_t)oy * t.maxk;
        for (int c = 0; c < 3; ++c) {
            const float* p = in + ((int64_t)c * hc + t.xmin[oy]) * res + ox;
            float acc = PL_MUL(p[0], w[0]);
            for (int j = 1; j < t.xsize[oy]; ++j) acc = PL_FMA(p[(int64_t)j * res], w[j], acc);
            const float v = (float)(uint8_t)PL_RINT(acc);                                        // round(), .to(uint8)
            const float n = PL_ADD(PL_DIV(PL_MUL(v, 2.f), 255.f), -1.f);                         // image * 2 / 255 - 1
            x[(c + 1) * plane + i] = PL_MUL(n, m);
        }
    }
};

// Feathered blend of the resized generator output into the crop, in place (:241-248, :262).  Same arithmetic as FeatherK
// (prepost.cu) on a window of the full image: max-pool and reflect padding act at the CROP border, like the reference,
// which slices the crop before post-processing.
struct FeatherCropK {  // items = hc * wc
    const float* g; uint8_t* img; const uint8_t* mask; int64_t img_plane; int W; int hc, wc; float k[25];
    PL_D void operator()(int64_t i) const {
        const int h = (int)(i / wc), w = (int)(i - (int64_t)h * wc);
        double acc = 0.0;
        for (int dy = -2; dy <= 2; ++dy) {
            int yy = h + dy;
            yy = yy < 0 ? -yy : (yy >= hc ? 2 * hc - 2 - yy : yy);
            for (int dx = -2; dx <= 2; ++dx) {
                int xx = w + dx;
                xx = xx < 0 ? -xx : (xx >= wc ? 2 * wc - 2 - xx : xx);
                int mx = 0;
                for (int a = -1; a <= 1; ++a) {
                    const int y2 = yy + a;
                    if (y2 < 0 || y2 >= hc) continue;
                    for (int b = -1; b <= 1; ++b) {
                        const int x2 = xx + b;
                        if (x2 < 0 || x2 >= wc) continue;
                        const int v = mask[(int64_t)y2 * W + x2];
                        mx = v > mx ? v : mx;
                    }
                }
                acc += (double)k[(dy + 2) * 5 + (dx + 2)] * (double)mx;
            }
        }
        const float wgt = PL_DIV((float)acc, 255.f);
        const float inv = PL_ADD(1.f, -wgt);
        for (int c = 0; c < 3; ++c) {
            uint8_t* px = img + c * img_plane + (int64_t)h * W + w;
            const float v = PL_ADD(PL_MUL((float)*px, wgt), PL_MUL(g[(int64_t)c * hc * wc + i], inv));
            *px = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
        }
    }
};

}  // namespace comod

namespace migan {

namespace {

inline int aa_maxk(int in_size, int out_size) {    // max_interp_size of _compute_index_ranges_weights (antialias branch)
    if (in_size == out_size) return 1;
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.0f ? scale : 1.f;
    return (int)ceilf(support) * 2 + 1;
}
inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
inline size_t table_bytes(int in_size, int out_size) {
    return align256((size_t)out_size * 2 * sizeof(int)) + align256((size_t)out_size * aa_maxk(in_size, out_size) * sizeof(float));
}
// carve one axis table out of the scratch area
inline comod::AaTable take_table(unsigned char*& p, int in_size, int out_size) {
    comod::AaTable t;
    t.maxk = aa_maxk(in_size, out_size);
    t.xmin = reinterpret_cast<int*>(p);
    t.xsize = t.xmin + out_size;
    p += align256((size_t)out_size * 2 * sizeof(int));
    t.w = reinterpret_cast<float*>(p);
    p += align256((size_t)out_size * t.maxk * sizeof(float));
    return t;
}

}  // namespace

size_t pipeline_scratch_bytes(int H, int W, int res) {
    // the crop is at most the whole image: pre = tables (H -> res, W -> res) + [3][H][res]; post = tables (res -> H, res -> W)
    // + [3][res][W] + [3][H][W]; the two phases reuse the same area
    const size_t pre = table_bytes(H, res) + table_bytes(W, res) + align256((size_t)3 * H * res * sizeof(float));
    const size_t post = table_bytes(res, H) + table_bytes(res, W) + align256((size_t)3 * res * W * sizeof(float)) +
                        align256((size_t)3 * H * W * sizeof(float));
    return (pre > post ? pre : post) + 1024;
}

int launch_resize_nearest_u8(const uint8_t* in, int H, int W, uint8_t* out, int oh, int ow, ck_stream_t s) {
    comod::NearestU8K k{in, out, H, W, oh, ow};
    return (int)comod::ck_launch(k, (int64_t)oh * ow, s);
}

int launch_hole_flags(const uint8_t* mask, int H, int W, uint8_t* flags, ck_stream_t s) {
    comod::HoleFlagsK k{mask, flags, H, W};
    return (int)comod::ck_launch(k, (int64_t)W + H, s);
}

// Host arithmetic of get_masked_bbox (:151-227) on the hole flags (flags[x] for columns, flags[W + y] for rows).
void crop_box_from_flags(const uint8_t* flags, int H, int W, int res, int padding, int* box) {
    auto imin = [](int a, int b) { return a < b ? a : b; };
    auto imax = [](int a, int b) { return a > b ? a : b; };
    int x_min = W, x_max = 0, y_min = H, y_max = 0;
    for (int x = 0; x < W; ++x) if (flags[x]) { x_min = imin(x_min, x); x_max = imax(x_max, x); }
    for (int y = 0; y < H; ++y) if (flags[W + y]) { y_min = imin(y_min, y); y_max = imax(y_max, y); }
    x_min = imin(x_min, x_max); x_max = imax(x_min, x_max);
    y_min = imin(y_min, y_max); y_max = imax(y_min, y_max);
    const int cnt_x = (x_min + x_max) / 2, cnt_y = (y_min + y_max) / 2;      // non-negative: floor division == truncation
    int crop = imax(x_max - x_min, y_max - y_min) + padding * 2;
    crop = imax(crop, res);
    const int off = crop / 2;
    x_min = imax(cnt_x - off, 0); x_max = imin(cnt_x + off, W);
    y_min = imax(cnt_y - off, 0); y_max = imin(cnt_y + off, H);
    const int x_ex = imax(crop - (x_max - x_min), 0), y_ex = imax(crop - (y_max - y_min), 0);
    x_min = imax(x_min - x_ex, 0); x_max = imin(x_max + x_ex, W);
    y_min = imax(y_min - y_ex, 0); y_max = imin(y_max + y_ex, H);
    box[0] = x_min; box[1] = x_max; box[2] = y_min; box[3] = y_max;
}

int launch_pipeline_preprocess(const uint8_t* image, const uint8_t* mask, int H, int W, const int* box, int res, float* x,
                               void* scratch, ck_stream_t s) {
    const int x0 = box[0], x1 = box[1], y0 = box[2], y1 = box[3];
    const int wc = x1 - x0, hc = y1 - y0;
    unsigned char* p = static_cast<unsigned char*>(scratch);
    comod::AaTable tw = take_table(p, wc, res), th = take_table(p, hc, res);
    float* tmp = reinterpret_cast<float*>(p);                                   // [3][hc][res]
    int e;
    if ((e = (int)comod::ck_launch(comod::AaWeightsK{tw, wc, res}, res, s))) return e;
    if ((e = (int)comod::ck_launch(comod::AaWeightsK{th, hc, res}, res, s))) return e;
    comod::ResizeWidthK<0> kw{image + (int64_t)y0 * W + x0, (int64_t)H * W, W, tw, tmp, hc, res};
    if ((e = (int)comod::ck_launch(kw, (int64_t)3 * hc * res, s))) return e;
    comod::PreprocessK kp{tmp, th, mask + (int64_t)y0 * W + x0, W, hc, wc, res, x};
    return (int)comod::ck_launch(kp, (int64_t)res * res, s);
}

int launch_pipeline_postprocess(const float* y, uint8_t* image, const uint8_t* mask, int H, int W, const int* box, int res,
                                const float* k25_host, void* scratch, ck_stream_t s) {
    const int x0 = box[0], x1 = box[1], y0 = box[2], y1 = box[3];
    const int wc = x1 - x0, hc = y1 - y0;
    unsigned char* p = static_cast<unsigned char*>(scratch);
    comod::AaTable tw = take_table(p, res, wc), th = take_table(p, res, hc);
    float* tmp = reinterpret_cast<float*>(p);                                   // [3][res][wc]
    p += align256((size_t)3 * res * wc * sizeof(float));
    float* g = reinterpret_cast<float*>(p);                                     // [3][hc][wc]
    int e;
    if ((e = (int)comod::ck_launch(comod::AaWeightsK{tw, res, wc}, wc, s))) return e;
    if ((e = (int)comod::ck_launch(comod::AaWeightsK{th, res, hc}, hc, s))) return e;
    comod::ResizeWidthK<1> kw{y, (int64_t)res * res, res, tw, tmp, res, wc};
    if ((e = (int)comod::ck_launch(kw, (int64_t)3 * res * wc, s))) return e;
    comod::ResizeHeightK kh{tmp, th, g, res, hc, wc};
    if ((e = (int)comod::ck_launch(kh, (int64_t)3 * hc * wc, s))) return e;
    comod::FeatherCropK kf{g, image + (int64_t)y0 * W + x0, mask + (int64_t)y0 * W + x0, (int64_t)H * W, W, hc, wc, {0}};
    for (int i = 0; i < 25; ++i) kf.k[i] = k25_host[i];
    return (int)comod::ck_launch(kf, (int64_t)hc * wc, s);
}

}  // namespace migan

#ifdef MIGAN_EMULATE   // the product's extern "C" wrappers (argument checks, error strings) live in migan_abi.cu
namespace migan {
size_t pipeline_scratch_bytes(int H, int W, int res);
}
extern "C" size_t b200_pipeline_scratch_bytes(int H, int W, int res) { return migan::pipeline_scratch_bytes(H, W, res); }
extern "C" int b200_resize_nearest_u8(const uint8_t* in, int H, int W, uint8_t* out, int oh, int ow, void* s) {
    return migan::launch_resize_nearest_u8(in, H, W, out, oh, ow, s);
}
extern "C" int b200_hole_flags(const uint8_t* mask, int H, int W, uint8_t* flags, void* s) { return migan::launch_hole_flags(mask, H, W, flags, s); }
extern "C" int migan_crop_box(const uint8_t* flags, int H, int W, int res, int padding, int* box) {
    migan::crop_box_from_flags(flags, H, W, res, padding, box);
    return 0;
}
extern "C" int b200_pipeline_preprocess(const uint8_t* image, const uint8_t* mask, int H, int W, const int* box, int res, float* x,
                                        void* scratch, size_t, void* s) {
    return migan::launch_pipeline_preprocess(image, mask, H, W, box, res, x, scratch, s);
}
extern "C" int b200_pipeline_postprocess(const float* y, uint8_t* image, const uint8_t* mask, int H, int W, const int* box, int res,
                                         const float* k25, void* scratch, size_t, void* s) {
    return migan::launch_pipeline_postprocess(y, image, mask, H, W, box, res, k25, scratch, s);
}
#endif
