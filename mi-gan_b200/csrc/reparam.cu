// Training snapshot -> inference weights on the GPU: the filter transformation of the reference's export script,
// scripts/export_inference_model.py:18-27 (`get_source_w`), SURVEY.md 8(f) row f3:
//   w = (w0 + w1 + ... + w{k-1}) / sqrt(k)      re-parameterised layers (k = num_reparam_tensors), else w = weight
//   w = w * rsqrt(sum over (cin, kh, kw) of w^2 + 1e-8)     per output filter
// One item per output filter (at most 512 filters of at most 4608 taps per layer, ~6 M values per model: a one-off).
// The fp32 operations are the reference's, in its order (sequential adds, true division by (float)sqrt(k), 1 / sqrt(s));
// the sum of squares is accumulated in fp64 and rounded once, where torch's CPU reduction uses a vectorised fp32 tree --
// so the result agrees with the reference to a few ulp, not bit for bit (tests: relative error < 1e-6).
// Item kernel (comod_kernels.cuh style): the emulation build runs the same functor on the CPU.
#include "comod_kernels.cuh"

#ifdef MIGAN_EMULATE
#include <cmath>
#define RP_ADD(a, b) ((a) + (b))
#define RP_MUL(a, b) ((a) * (b))
#define RP_DIV(a, b) ((a) / (b))
#define RP_SQRT(a) sqrtf(a)
#define RP_D inline
#else
#include "kernels.h"
#define RP_ADD(a, b) __fadd_rn((a), (b))
#define RP_MUL(a, b) __fmul_rn((a), (b))
#define RP_DIV(a, b) __fdiv_rn((a), (b))
#define RP_SQRT(a) __fsqrt_rn(a)
#define RP_D __device__ __forceinline__
#endif

namespace comod {

constexpr int kMaxReparam = 16;

struct ReparamFilterK {   // items = cout
    const float* w[kMaxReparam]; int k; int64_t fan; float div; float* out;
    RP_D float merged(int64_t idx) const {
        float v = w[0][idx];
        for (int j = 1; j < k; ++j) v = RP_ADD(v, w[j][idx]);
        return k > 1 ? RP_DIV(v, div) : v;
    }
    RP_D void operator()(int64_t o) const {
        double acc = 0.0;
        for (int64_t i = 0; i < fan; ++i) { const float v = merged(o * fan + i); acc += (double)RP_MUL(v, v); }
        const float s = RP_ADD((float)acc, 1e-8f);
        const float r = RP_DIV(1.f, RP_SQRT(s));
        for (int64_t i = 0; i < fan; ++i) out[o * fan + i] = RP_MUL(merged(o * fan + i), r);
    }
};

}  // namespace comod

namespace migan {
int launch_reparam_filter(const float* const* w_dev, int k, int cout, int64_t fan, float* out, ck_stream_t s) {
    comod::ReparamFilterK f;
    for (int j = 0; j < comod::kMaxReparam; ++j) f.w[j] = j < k ? w_dev[j] : nullptr;
    f.k = k; f.fan = fan; f.div = (float)sqrt((double)k); f.out = out;
    return (int)comod::ck_launch(f, (int64_t)cout, s);
}
}  // namespace migan

#ifdef MIGAN_EMULATE   // the product's extern "C" wrapper (argument checks, error strings) lives in migan_abi.cu
extern "C" int b200_reparam_filter(const float* const* w, int k, int cout, int64_t fan, float* out, void* s) {
    if (k < 1 || k > comod::kMaxReparam) return 1;
    return migan::launch_reparam_filter(w, k, cout, fan, out, s);
}
#endif
