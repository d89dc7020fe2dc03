// Fused SeparableConv2d on Blackwell tensor cores (sm_100a): the hot kernel of the generator.
//
//   out = epi( PW( act( DW3x3( src ) + b ) ) )          lib/model_zoo/migan_inference.py:154-170
//
//   src                        where the depthwise conv's input comes from (SepconvSource):
//                                NHWC   a TMA-staged tile + 1-pixel halo of an NHWC fp32 tensor
//                                UP     rebuilt in shared memory from the PREVIOUS layer's raw low-resolution 1x1 output:
//                                       lrelu_agc(Upsample2d(t) + noise) + skip (:98-103, :165-169, :304-305) -- the
//                                       up-sampled tensor never exists in HBM
//                                STEM   fromrgb 1x1 (4 -> C0) + activation (:193-196) recomputed on the halo from the
//                                       generator input x -- the stem tensor never exists in HBM
//                                SPLIT  no depthwise stage: a pre-split fp16 hi/lo A operand is loaded by TMA
//   DW3x3 + bias + lrelu_agc   CUDA cores (packed FFMA2), fp32 (sepconv_stages.cuh)
//   PW (1x1 conv, Cin -> Cout) tcgen05.mma kind::f16, M = 128 pixels x N <= 128 channels per accumulator, accumulators
//                              in TMEM.  fp32-faithful mode: both operands are split into fp16 (hi, lo) pairs;
//                              [main | corr] = Ah * [Bh ; Bl]^T (one N = 2*n_tile instruction), corr += Al * Bh,
//                              main + corr summed in the epilogue (separate accumulators: the tensor core truncates
//                              on accumulate).  Fast mode: Ah * Bh only.
//   epi                        TMEM -> registers: * 2^-k, + noise, lrelu_agc [-> torgb + image] -> NHWC fp32, either
//                              through swizzled staging + per-warp TMA store or straight from registers (256-bit stores)
//
// One persistent CTA per SM, 14 warps x 128 registers, warp-specialised:
//   warps 0-7    prologue      two groups of 4: pre-stage (UP / STEM) + depthwise conv -> fp16 hi/lo A operand
//   warps 8-11   epilogue      one TMEM lane quarter each
//   warp 12      TMA producer  input chunks, weight K blocks, L2 prefetch
//   warp 13      MMA issuer    one elected lane issues tcgen05.mma / tcgen05.commit; owns TMEM
// The two 256-column accumulator regions are either the double buffer of consecutive M tiles (nt_share = 1) or the two
// N halves of ONE M tile (nt_share = 2, Cout >= 256: the A operand -- and the whole prologue -- is produced once and
// multiplied against both weight halves; the epilogue drains half 0 while the MMAs of half 1 finish).
// All hand-offs are mbarrier pipelines (input ring, A ring, B ring, accumulator regions).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "sepconv_stages.cuh"
#include "sepconv_tc.h"

namespace migan {

namespace {

using namespace stages;

constexpr int kThreads = 448;      // 14 warps; registers are allocated per 4 warps: 16 x 32 x 128 = the whole file
constexpr int kNumProWarps = 8;    // warps 0-7: the SM's issue arbiter favours higher warp ids, so the latency-critical
constexpr int kEpiWarp0 = 8;       // single-warp roles get the highest ids and the throughput-oriented prologue the lowest
constexpr int kProducerWarp = 12;
constexpr int kMmaWarp = 13;
constexpr int kTileM = 128;
constexpr int kKBlock = 64;        // channels per A/B stage (128 bytes of fp16: one SW128 row)
constexpr int kChunkC = 32;        // channels per input chunk (128 bytes of fp32)
constexpr uint32_t kABytes = kTileM * kKBlock * 2;   // 16 KB per hi or lo
constexpr uint32_t kAStage = 2 * kABytes;            // hi + lo
constexpr uint32_t kEpiWarpBuf = 32 * 32 * 4;        // 4 KB: 32 pixel rows x 32 fp32 of one epilogue warp
constexpr uint32_t kSmemLimit = 232448;              // 227 KB
constexpr int kRgbBias = 3 * 256, kRgbFir = 3 * 256 + 4;   // offsets inside the torgb table s_rgb
constexpr uint32_t kStaticSmem = 4096;               // barriers + torgb table (static __shared__), rounded up

struct Params {
    CUtensorMap map_in, map_t, map_aux, map_a_hi, map_a_lo, map_w_hi, map_w_lo, map_out;   // map_out box = 32 pixels x 32 channels
    const float* w9;                          // [9][cin] depthwise taps * (kActSplitScale * sqrt 2)
    const float* bias;                        // [cin]              * (kActSplitScale * sqrt 2)
    const float* noise;
    float* out;                               // NHWC output (direct-store epilogue)
    float inv_scale;
    int n, H, W, cin, cout;
    int act, passes;
    int source;                               // SepconvSource
    int nt_share;                             // accumulator regions per M tile (1 or 2)
    int tile_n, tile_h, tile_w, n_tile;      // n_tile = N per accumulator region (64 or 128)
    int tiles_x, tiles_y, tiles_n, num_n_tiles, num_tiles;   // num_n_tiles = cout / (n_tile * nt_share)
    int l_tx, l_ty, l_nt;                     // log2 of tiles_x, tiles_y, num_n_tiles (all powers of two: no runtime division)
    int num_kb;                               // cin / 64
    int in_stages, a_stages, b_stages, b_resident, epi_bufs, epi_direct;
    uint32_t in_tx_bytes;                     // bytes the TMA unit delivers per input stage
    uint32_t in_stage_stride;                 // rounded to 1024
    uint32_t off_t, off_aux;                  // T / NZ|XA areas inside an input stage
    uint32_t off_in, off_a, off_b, off_epi;   // smem offsets from the 1024-aligned base
    uint32_t off_dw;                          // [10][cin] depthwise taps + bias staged in smem (0xFFFFFFFF: read from global)
    uint32_t off_stem;                        // STEM: [cin][4] + [cin] fromrgb table
    int prefetch;                             // L2 prefetch distance of the producer, in chunks / K-blocks
    // UP pre-stage
    UpTaps up_taps;
    int up_has_noise;
    // STEM pre-stage
    const float* stem_w;
    const float* stem_b;
    // fused torgb + image path (SynthesisBlock.forward, migan_inference.py:308-313); needs all of cout in one region
    int torgb, store_out;
    const float* rgb_w;                       // [3][cout]
    const float* rgb_b;                       // [3]
    const float* rgb_fir;                     // [16][3] up-sampling taps of the image path
    const float* img_lo;                      // [n][3][H/2][W/2] planar, or null (first block)
    float* img_out;                           // [n][3][H][W] planar
    int* error_flag;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware (NANOSLEEP.SYNCS) until the phase completes or the hint
// elapses, instead of burning issue slots.  (The round-2 ncu source view attributed 25-32 % of all executed instructions
// to the previous wait loop, which read %globaltimer on every iteration.)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(200000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Bounded wait: a protocol bug must never hang the GPU -- record which wait gave up and trap instead.  The bound is
// wall-clock time (seconds), so profiler replay, sanitizers or a co-tenant cannot turn a legitimately long wait into a
// false positive.  The record goes to a host-mapped word (it survives the trap).  The clock is only looked at every
// 1024 failed probes: the common path is the probe loop alone.
__device__ __noinline__ void mbar_timeout(int code, uint32_t parity, int* error_flag) {
    if (error_flag) {
        *reinterpret_cast<volatile int*>(error_flag) = code | ((int)parity << 12) | ((int)(blockIdx.x & 0xFFF) << 16);
        __threadfence_system();
    }
    __trap();
}
// kRelaxed: a waiter with slack (the consumers behind a double buffer) backs off between probes, so that it does not
// compete for issue slots with the warps doing the arithmetic (round-2 ncu: a quarter of all executed instructions were probes).
template <bool kRelaxed = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code, int* error_flag) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0 = 0;
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 1024; ++i) {
            if (kRelaxed) asm volatile("nanosleep.u32 256;" ::: "memory");
            if (mbar_try_wait(bar, parity)) return;
        }
        const unsigned long long t = globaltimer_ns();
        if (t0 == 0) t0 = t;
        else if (t - t0 > 4000000000ull) mbar_timeout(code, parity, error_flag);
    }
}
__device__ __forceinline__ void bar_sync_named(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// The loaded registers are tied to the wait as in/out operands so the compiler cannot hoist
// their uses above it (tcgen05.ld is asynchronous until wait::ld).
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}
__device__ __forceinline__ void stg_v8(float* p, const float (&o)[32], int i) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(o[i]), "f"(o[i + 1]), "f"(o[i + 2]),
                 "f"(o[i + 3]), "f"(o[i + 4]), "f"(o[i + 5]), "f"(o[i + 6]), "f"(o[i + 7]) : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 64 fp16 (128 B),
// 8-row groups 1024 B apart (SBO).  Bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // SBO = 1024 B   [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
    return d;
}
// Instruction descriptor (InstrDescriptor in the same header): D=f32, A=B=f16, K-major both, M=128.
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct TileCoord {
    int n0, y0, x0, nt;   // first image / row / column of the M tile, N-tile (group) index
};
__device__ __forceinline__ TileCoord decode_tile(const Params& p, int tile) {
    TileCoord c;
    c.nt = tile & (p.num_n_tiles - 1);
    int m = tile >> p.l_nt;
    c.x0 = (m & (p.tiles_x - 1)) * p.tile_w;
    m >>= p.l_tx;
    c.y0 = (m & (p.tiles_y - 1)) * p.tile_h;
    c.n0 = (m >> p.l_ty) * p.tile_n;
    return c;
}

// Ring-buffer cursor: stage index + phase parity, advanced incrementally (no div/mod on the hot path).
struct Ring {
    int stage, phase, n, step;
    __device__ __forceinline__ Ring(int n_, int start = 0, int step_ = 1) : stage(start), phase(0), n(n_), step(step_) {}
    __device__ __forceinline__ void advance() {
        stage += step;
        if (stage >= n) { stage -= n; phase ^= 1; }
    }
};

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
sepconv_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    // mbarriers: [0,8) full_in  [8,16) empty_in  [16,20) full_a  [20,24) empty_a
    //            [24,32) full_b  [32,40) empty_b  [40,42) full_acc  [42,44) empty_acc
    __shared__ __align__(8) uint64_t bars[44];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_rgb[3 * 256 + 4 + 48];   // torgb weights [3][cout<=256], bias[3(+1)], fir[16][3]

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = smem_u32(bars);
    auto full_in = [&](int s) { return bar0 + 8u * (0 + s); };
    auto empty_in = [&](int s) { return bar0 + 8u * (8 + s); };
    auto full_a = [&](int s) { return bar0 + 8u * (16 + s); };
    auto empty_a = [&](int s) { return bar0 + 8u * (20 + s); };
    auto full_b = [&](int s) { return bar0 + 8u * (24 + s); };
    auto empty_b2 = [&](int s) { return bar0 + 8u * (32 + s); };
    auto full_acc = [&](int s) { return bar0 + 8u * (40 + s); };
    auto empty_acc = [&](int s) { return bar0 + 8u * (42 + s); };

    const int a_dw = (p.source != SEPCONV_SRC_SPLIT);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.in_stages; ++s) { mbar_init(full_in(s), 1); mbar_init(empty_in(s), kNumProWarps / 2); }
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(full_a(s), a_dw ? kNumProWarps : 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < p.b_stages; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b2(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(full_acc(s), 1); mbar_init(empty_acc(s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarp && lane == 0) {
        if (p.source == SEPCONV_SRC_SPLIT) { prefetch_tensormap(&p.map_a_hi); prefetch_tensormap(&p.map_a_lo); }
        else if (p.source == SEPCONV_SRC_STEM) prefetch_tensormap(&p.map_aux);
        else prefetch_tensormap(&p.map_in);
        if (p.source == SEPCONV_SRC_UP) { prefetch_tensormap(&p.map_t); if (p.up_has_noise) prefetch_tensormap(&p.map_aux); }
        prefetch_tensormap(&p.map_w_hi); prefetch_tensormap(&p.map_w_lo); prefetch_tensormap(&p.map_out);
    }
    if (p.torgb) {
        for (int i = threadIdx.x; i < 3 * p.cout; i += kThreads) s_rgb[i] = __ldg(p.rgb_w + i);
        if (threadIdx.x < 3) s_rgb[kRgbBias + threadIdx.x] = __ldg(p.rgb_b + threadIdx.x);
        if (threadIdx.x < 48 && p.img_lo) s_rgb[kRgbFir + threadIdx.x] = __ldg(p.rgb_fir + threadIdx.x);
    }
    if (a_dw && p.off_dw != 0xFFFFFFFFu) {   // depthwise taps + bias for all channels: read per chunk by every prologue thread
        float* sdw = reinterpret_cast<float*>(smem_gen + p.off_dw);
        for (int i = threadIdx.x; i < 9 * p.cin; i += kThreads) sdw[i] = __ldg(p.w9 + i);
        for (int i = threadIdx.x; i < p.cin; i += kThreads) sdw[9 * p.cin + i] = __ldg(p.bias + i);
    }
    if (p.source == SEPCONV_SRC_STEM) {
        float* st = reinterpret_cast<float*>(smem_gen + p.off_stem);
        // pair-interleaved: [channel pair][plane][2] so that the pre-stage reads (w[ch][i], w[ch+1][i]) as one register pair
        for (int i = threadIdx.x; i < 4 * p.cin; i += kThreads) st[(i >> 3) * 8 + (i & 3) * 2 + ((i >> 2) & 1)] = __ldg(p.stem_w + i);
        for (int i = threadIdx.x; i < p.cin; i += kThreads) st[4 * p.cin + i] = __ldg(p.stem_b + i);
    }
    pdl_trigger();           // the successor's blocks may start their own setup as soon as SMs free up
    if (warp == kMmaWarp) {  // TMEM: all 512 columns (one CTA per SM by construction)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();              // barriers, TMEM and the weight tables above are set up under the predecessor's tail; its output is read below
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    const int num_kb = p.num_kb;
    const int my_tiles = (p.num_tiles > (int)blockIdx.x) ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nts = p.nt_share;
    const uint32_t b_stage_bytes = (uint32_t)p.n_tile * kKBlock * 2 * 2;   // [Bh ; Bl] of one accumulator region

    if (warp == kProducerWarp) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            // L2 prefetch cursor: the smem rings are too shallow to cover HBM latency at full bandwidth, so the
            // producer also asks the L2 for the data `prefetch` loads ahead (possibly in the next tile).
            int pf_it = 0, pf_kb = 0, pf_g = 0;
            auto prefetch_next = [&]() {
                if (pf_it >= my_tiles) return;
                if (a_dw) {
                    const TileCoord t2 = decode_tile(p, blockIdx.x + pf_it * gridDim.x);
                    if (p.source != SEPCONV_SRC_STEM)
                        tma_prefetch_4d(&p.map_in, pf_kb * kKBlock + pf_g * kChunkC, t2.x0 - 1, t2.y0 - 1, t2.n0);
                    if (++pf_g < 2) return;
                    pf_g = 0;
                } else {
                    const int row0 = ((blockIdx.x + pf_it * gridDim.x) >> p.l_nt) * kTileM;
                    tma_prefetch_2d(&p.map_a_hi, pf_kb * kKBlock, row0);
                    tma_prefetch_2d(&p.map_a_lo, pf_kb * kKBlock, row0);
                }
                if (++pf_kb == num_kb) { pf_kb = 0; ++pf_it; }
            };
            for (int c = 0; c < p.prefetch; ++c) prefetch_next();
            Ring rin(p.in_stages), ra(p.a_stages), rb(p.b_stages);
            for (int it = 0; it < my_tiles; ++it) {
                const TileCoord tc = decode_tile(p, blockIdx.x + it * gridDim.x);
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (a_dw) {
                        for (int g = 0; g < 2; ++g) {
                            const int s = rin.stage;
                            if (p.prefetch) prefetch_next();
                            mbar_wait<true>(empty_in(s), rin.phase ^ 1, 100 + s, p.error_flag);
                            rin.advance();
                            mbar_expect_tx(full_in(s), p.in_tx_bytes);
                            const uint32_t dst = smem_base + p.off_in + s * p.in_stage_stride;
                            const int c0 = kb * kKBlock + g * kChunkC;
                            if (p.source == SEPCONV_SRC_STEM) {
                                tma_load_4d(dst + p.off_aux, &p.map_aux, full_in(s), tc.x0 - kAuxLeft, tc.y0 - 1, 0, tc.n0);
                            } else {
                                tma_load_4d(dst, &p.map_in, full_in(s), c0, tc.x0 - 1, tc.y0 - 1, tc.n0);
                                if (p.source == SEPCONV_SRC_UP) {
                                    tma_load_4d(dst + p.off_t, &p.map_t, full_in(s), c0, (tc.x0 >> 1) - 1, (tc.y0 >> 1) - 1, tc.n0);
                                    if (p.up_has_noise) tma_load_2d(dst + p.off_aux, &p.map_aux, full_in(s), tc.x0 - kAuxLeft, tc.y0 - 1);
                                }
                            }
                        }
                    } else {
                        const int s = ra.stage;
                        if (p.prefetch) prefetch_next();
                        mbar_wait(empty_a(s), ra.phase ^ 1, 110 + s, p.error_flag);
                        ra.advance();
                        mbar_expect_tx(full_a(s), kAStage);
                        const int row0 = ((blockIdx.x + it * gridDim.x) >> p.l_nt) * kTileM;
                        tma_load_2d(smem_base + p.off_a + s * kAStage, &p.map_a_hi, full_a(s), kb * kKBlock, row0);
                        tma_load_2d(smem_base + p.off_a + s * kAStage + kABytes, &p.map_a_lo, full_a(s), kb * kKBlock, row0);
                    }
                    if (!p.b_resident || it == 0) {
                        for (int sub = 0; sub < nts; ++sub) {
                            const int s = p.b_resident ? kb : rb.stage;
                            if (!p.b_resident) { mbar_wait(empty_b2(s), rb.phase ^ 1, 120 + s, p.error_flag); rb.advance(); }
                            mbar_expect_tx(full_b(s), b_stage_bytes);
                            const uint32_t dst = smem_base + p.off_b + s * b_stage_bytes;
                            const int row = (tc.nt * nts + sub) * p.n_tile;
                            tma_load_2d(dst, &p.map_w_hi, full_b(s), kb * kKBlock, row);
                            tma_load_2d(dst + b_stage_bytes / 2, &p.map_w_lo, full_b(s), kb * kKBlock, row);
                        }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer ==================================
        const uint32_t idesc = umma_idesc_f16(p.n_tile), idesc2 = umma_idesc_f16(2 * p.n_tile);
        Ring ra(p.a_stages), rb(p.b_stages);
        for (int it = 0; it < my_tiles; ++it) {
            // Accumulator region = [main | correction] column blocks.  The 2^-11-sized correction products
            // (Al*Bh + Ah*Bl) go to their own accumulator: the tensor core truncates on every accumulate,
            // and adding them into the large main sum would cost ~0.5 ulp of the MAIN sum per MMA.
            if (nts == 1) {
                mbar_wait(empty_acc(it & 1), ((it >> 1) & 1) ^ 1, 200 + (it & 1), p.error_flag);
            } else {
                mbar_wait(empty_acc(0), (it & 1) ^ 1, 200, p.error_flag);
                mbar_wait(empty_acc(1), (it & 1) ^ 1, 201, p.error_flag);
            }
            tc_fence_after();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int sa = ra.stage;
                mbar_wait(full_a(sa), ra.phase, 210 + sa, p.error_flag);
                ra.advance();
                for (int sub = 0; sub < nts; ++sub) {
                    const int sb = p.b_resident ? kb : rb.stage;
                    mbar_wait(full_b(sb), p.b_resident ? 0 : rb.phase, 220 + sb, p.error_flag);
                    if (!p.b_resident) rb.advance();
                    tc_fence_after();
                    if (lane == 0) {
                        const int region = (nts == 1) ? (it & 1) : sub;
                        const uint32_t tmem_d = tmem_base + (uint32_t)(region * 2 * p.n_tile);
                        const uint32_t tmem_c = tmem_d + (uint32_t)p.n_tile;
                        const uint32_t a_hi = smem_base + p.off_a + sa * kAStage, a_lo = a_hi + kABytes;
                        const uint32_t b_hi = smem_base + p.off_b + sb * b_stage_bytes;
                        const uint64_t dah = umma_desc_sw128(a_hi), dal = umma_desc_sw128(a_lo);
                        const uint64_t dbh = umma_desc_sw128(b_hi);   // Bl sits right behind Bh: [Bh ; Bl] is one K-major operand of 2*n_tile rows
#pragma unroll
                        for (int k = 0; k < kKBlock / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);   // 16 fp16 = 32 bytes along K inside the SW128 row
                            if (p.passes == 3) {
                                // [main | corr] = Ah * [Bh ; Bl]^T in ONE instruction (N = 2*n_tile): Ah is read from shared
                                // memory once for both products; then corr += Al * Bh.
                                tc_mma_f16(tmem_d, dah + adv, dbh + adv, idesc2, (kb | k) != 0);
                                tc_mma_f16(tmem_c, dal + adv, dbh + adv, idesc, 1u);
                            } else {
                                tc_mma_f16(tmem_d, dah + adv, dbh + adv, idesc, (kb | k) != 0);
                            }
                        }
                        if (!p.b_resident) tc_commit(empty_b2(sb));
                        if (sub == nts - 1) tc_commit(empty_a(sa));      // A slot reusable once these MMAs retire
                        if (kb == num_kb - 1) tc_commit(full_acc(region));   // accumulator region complete
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ================================ epilogue ====================================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;               // pixel row of the M tile
        const int hw = p.tile_h * p.tile_w;
        const int img_l = row / hw, yl = (row / p.tile_w) % p.tile_h, xl = row % p.tile_w;   // once per kernel
        const int chunks = p.n_tile / 32;
        int buf = 0;
        // each epilogue warp stages and TMA-stores its own 32 pixel rows: no cross-warp barrier in the loop
        const int q_img = (q * 32) / hw, q_rem = (q * 32) % hw, q_y = q_rem / p.tile_w, q_x = q_rem % p.tile_w;
        const uint32_t stage_base = smem_base + p.off_epi + (uint32_t)(q * p.epi_bufs) * kEpiWarpBuf;
        for (int it = 0; it < my_tiles; ++it) {
            const TileCoord tc = decode_tile(p, blockIdx.x + it * gridDim.x);
          // torgb partial sums live across the N halves of the tile (nts = 2: the same thread drains both regions of its pixel)
          P2 rgb0 = p2zero(), rgb1 = p2zero(), rgb2 = p2zero();   // (even, odd) partial sums of the three torgb outputs
          float lo_tap[12];
          for (int sub = 0; sub < nts; ++sub) {
            // accumulator region + the parity of its current use: consecutive tiles alternate regions (nts = 1), or
            // every tile uses both regions for its two N halves (nts = 2)
            const int e = (nts == 1) ? (it & 1) : sub;
            const uint32_t use_parity = (nts == 1) ? ((it >> 1) & 1) : (it & 1);
            const uint32_t acc_col = (uint32_t)(e * 2 * p.n_tile);
            const int chan0 = (tc.nt * nts + sub) * p.n_tile;
            // operands of the tile's tail (noise, low-res image taps) are fetched BEFORE waiting for the
            // accumulator so that their L2 latency hides behind the MMA
            float nz = 0.f;
            const int pimg = tc.n0 + img_l, poy = tc.y0 + yl, pox = tc.x0 + xl;
            if (p.noise) nz = __ldg(p.noise + poy * p.W + pox);
            if (p.torgb && p.img_lo && sub == 0) {
                const int h = p.H >> 1, w = p.W >> 1;
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int bb = 0; bb < 2; ++bb) {
                        const int iy = (poy + (poy & 1) + 2 * a - 2) >> 1, ix = (pox + (pox & 1) + 2 * bb - 2) >> 1;
                        const bool ok = pimg < p.n && iy >= 0 && iy < h && ix >= 0 && ix < w;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            lo_tap[(a * 2 + bb) * 3 + k] = ok ? __ldg(p.img_lo + (((size_t)pimg * 3 + k) * h + iy) * w + ix) : 0.f;
                    }
            }
            mbar_wait<true>(full_acc(e), use_parity, 300 + e, p.error_flag);
            tc_fence_after();
            const float scale_g = p.inv_scale * kActGain, nz_g = nz * kActGain;
            float* out_px = p.out + (((size_t)pimg * p.H + poy) * p.W + pox) * p.cout + chan0;
            for (int j = 0; j < chunks; ++j) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)(j * 32);
                tc_ld32(taddr, v);
                if (p.passes == 3) {
                    uint32_t c[32];
                    tc_ld32(taddr + (uint32_t)p.n_tile, c);   // both loads in flight before the wait
                    tc_wait_ld(v);
                    tc_wait_ld(c);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const f2 t = unpk(fadd2(pk(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                                pk(__uint_as_float(c[i]), __uint_as_float(c[i + 1]))));
                        v[i] = __float_as_uint(t.x); v[i + 1] = __float_as_uint(t.y);
                    }
                } else {
                    tc_wait_ld(v);
                }
                if (j == chunks - 1) {               // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_acc(e));
                }
                float o[32];
                if (p.act) {                         // clamp(lrelu(f) * sqrt2) with the gain folded into the scale
                    const P2 sc2 = pk(scale_g, scale_g), nz2 = pk(nz_g, nz_g), al2 = pk(kLreluAlpha, kLreluAlpha);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const P2 a = ffma2(pk(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nz2);
                        const P2 b = fmul2(a, al2);
                        o[i] = fminf(fmax3(a.x, b.x, -kActClamp), kActClamp);
                        o[i + 1] = fminf(fmax3(a.y, b.y, -kActClamp), kActClamp);
                    }
                } else {
                    const P2 sc2 = pk(p.inv_scale, p.inv_scale), nz2 = pk(nz, nz);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const f2 a = unpk(ffma2(pk(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), sc2, nz2));
                        o[i] = a.x; o[i + 1] = a.y;
                    }
                }
                if (p.torgb) {   // 1x1 conv Cout -> 3 on the activated row: even/odd partial sums, packed FMAs, weights broadcast from smem
                    const float4* w0 = reinterpret_cast<const float4*>(s_rgb + chan0 + j * 32);
                    const float4* w1 = reinterpret_cast<const float4*>(s_rgb + p.cout + chan0 + j * 32);
                    const float4* w2 = reinterpret_cast<const float4*>(s_rgb + 2 * p.cout + chan0 + j * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 a = w0[i], b = w1[i], c = w2[i];
                        const P2 o01 = pk(o[4 * i], o[4 * i + 1]), o23 = pk(o[4 * i + 2], o[4 * i + 3]);
                        rgb0 = ffma2(o01, pk(a.x, a.y), rgb0); rgb0 = ffma2(o23, pk(a.z, a.w), rgb0);
                        rgb1 = ffma2(o01, pk(b.x, b.y), rgb1); rgb1 = ffma2(o23, pk(b.z, b.w), rgb1);
                        rgb2 = ffma2(o01, pk(c.x, c.y), rgb2); rgb2 = ffma2(o23, pk(c.z, c.w), rgb2);
                    }
                }
                if (!p.store_out) continue;          // last block: the feature map is consumed by torgb only
                if (p.epi_direct) {                  // 128 contiguous bytes per thread, four 256-bit stores
                    if (pimg < p.n) {
                        float* dstp = out_px + j * 32;
                        stg_v8(dstp, o, 0); stg_v8(dstp + 8, o, 8); stg_v8(dstp + 16, o, 16); stg_v8(dstp + 24, o, 24);
                    }
                    continue;
                }
                // this warp's staging buffer `buf` must have been drained by its own TMA store issued epi_bufs chunks ago
                if (lane == 0) { if (p.epi_bufs == 2) tma_wait_group_read<1>(); else tma_wait_group_read<0>(); }
                __syncwarp();
                const uint32_t dst = stage_base + buf * kEpiWarpBuf + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t a = dst + (((uint32_t)c ^ (uint32_t)(lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(o[4 * c]), "f"(o[4 * c + 1]),
                                 "f"(o[4 * c + 2]), "f"(o[4 * c + 3]) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&p.map_out, stage_base + buf * kEpiWarpBuf, chan0 + j * 32, tc.x0 + q_x, tc.y0 + q_y, tc.n0 + q_img);
                    tma_commit_group();
                }
                buf = (buf + 1 == p.epi_bufs) ? 0 : buf + 1;
            }
            if (p.torgb && sub == nts - 1) {
                // img = upsample(img_lo) + (torgb(x) + b)   (migan_inference.py:308-313); planar NCHW
                if (pimg < p.n) {
                    float r[3] = {rgb0.x + rgb0.y + s_rgb[kRgbBias], rgb1.x + rgb1.y + s_rgb[kRgbBias + 1], rgb2.x + rgb2.y + s_rgb[kRgbBias + 2]};
                    if (p.img_lo) {
                        float up[3] = {0.f, 0.f, 0.f};
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int bb = 0; bb < 2; ++bb) {
                                const int ty = (poy & 1) + 2 * a, tx = (pox & 1) + 2 * bb;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    up[k] = fmaf(s_rgb[kRgbFir + (ty * 4 + tx) * 3 + k], lo_tap[(a * 2 + bb) * 3 + k], up[k]);
                            }
#pragma unroll
                        for (int k = 0; k < 3; ++k) r[k] = up[k] + r[k];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) p.img_out[(((size_t)pimg * 3 + k) * p.H + poy) * p.W + pox] = r[k];
                }
            }
          }
        }
        if (lane == 0) tma_wait_group_all();
    } else if (a_dw) {
        // ================================ prologue (pre-stage + depthwise) ========================
        const int g = warp >> 2;                          // channel half of the K-block this group produces
        const int tg = threadIdx.x - g * 128;
        Ring rin(p.in_stages, g, 2), ra(p.a_stages);     // this group owns input stages g, g+2, ...
        const float* w9p = (p.off_dw != 0xFFFFFFFFu) ? reinterpret_cast<const float*>(smem_gen + p.off_dw) : p.w9;
        const float* bp = (p.off_dw != 0xFFFFFFFFu) ? w9p + 9 * p.cin : p.bias;
        const float* stem_tab = reinterpret_cast<const float*>(smem_gen + p.off_stem);
        for (int it = 0; it < my_tiles; ++it) {
            const TileCoord tc = decode_tile(p, blockIdx.x + it * gridDim.x);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = rin.stage;
                const int sa = ra.stage;
                mbar_wait(full_in(s), rin.phase, 400 + s, p.error_flag);
                rin.advance();
                uint8_t* stage = smem_gen + p.off_in + s * p.in_stage_stride;
                f4* sin = reinterpret_cast<f4*>(stage);
                const int cg0 = kb * kKBlock + g * kChunkC;
                if (p.source == SEPCONV_SRC_UP) {
                    prestage_up(sin, reinterpret_cast<const f4*>(stage + p.off_t), reinterpret_cast<const float*>(stage + p.off_aux),
                                p.up_taps, tc.y0, tc.x0, p.H, p.up_has_noise, tg);
                    bar_sync_named(1 + g, 128);
                } else if (p.source == SEPCONV_SRC_STEM) {
                    prestage_stem(sin, reinterpret_cast<const float*>(stage + p.off_aux), stem_tab, stem_tab + 4 * p.cin,
                                  cg0, tc.y0, tc.x0, p.H, tg);
                    bar_sync_named(1 + g, 128);
                }
                mbar_wait(empty_a(sa), ra.phase ^ 1, 410 + sa, p.error_flag);
                ra.advance();
                uint8_t* a_hi = smem_gen + p.off_a + sa * kAStage;
                uint8_t* a_lo = a_hi + kABytes;
                if (p.tile_w == 16) prologue_chunk<1, 8, 16>(sin, a_hi, a_lo, w9p, bp, p.cin, cg0, g, tg);
                else if (p.tile_w == 8) prologue_chunk<2, 8, 8>(sin, a_hi, a_lo, w9p, bp, p.cin, cg0, g, tg);
                else prologue_chunk<8, 4, 4>(sin, a_hi, a_lo, w9p, bp, p.cin, cg0, g, tg);
                fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core / TMA (async proxy)
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(full_a(sa));
                    mbar_arrive(empty_in(s));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

const char* encode_map(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* addr, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return "cuTensorMapEncodeTiled entry point not available";
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(addr), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        static thread_local char msg[96];
        snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return msg;
    }
    return nullptr;
}

const char* encode_x_map(CUtensorMap* m, const void* x, int n, int res) {
    const uint64_t dims[4] = {(uint64_t)res, (uint64_t)res, 4, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)res * 4, (uint64_t)res * res * 4, (uint64_t)res * res * 16};
    const uint32_t box[4] = {(uint32_t)kAuxW, 10, 4, 1};
    return encode_map(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

constexpr int kMaxDevices = 64;
int* g_error_flag[kMaxDevices] = {nullptr};        // device alias of the host-mapped timeout record, per device
int* g_error_flag_host[kMaxDevices] = {nullptr};

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// Params is kept opaque to the ABI layer: SepconvTcArgs carries it as bytes.
static_assert(sizeof(Params) <= sizeof(((SepconvTcArgs*)0)->params_blob), "params blob too small");

// Tensor map of an NHWC fp32 tensor for the TMA-staged down-sampling kernel (elementwise.cu): box = 64 channels x 36 columns x 1 row.
const char* make_down_tensor_map(DownTensorMap* desc, const float* in, int n, int H, int W, int C) {
    static_assert(sizeof(CUtensorMap) <= sizeof(DownTensorMap), "tensor map size");
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)n};
    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
    const uint32_t box[4] = {64u, 36u, 1u, 1u};
    return encode_map(reinterpret_cast<CUtensorMap*>(desc), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

cudaError_t configure_sepconv_tc() {
    cudaError_t e = cudaFuncSetAttribute(sepconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemLimit - kStaticSmem));
    if (e != cudaSuccess) return e;
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices && !g_error_flag[dev]) {
        // host-mapped so that the record of a timed-out wait survives the trap that follows it
        e = cudaHostAlloc(reinterpret_cast<void**>(&g_error_flag_host[dev]), sizeof(int), cudaHostAllocMapped);
        if (e != cudaSuccess) return e;
        *g_error_flag_host[dev] = 0;
        e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_error_flag[dev]), g_error_flag_host[dev], 0);
    }
    return e;
}

int sepconv_tc_timeout_record(int device) {
    return (device >= 0 && device < kMaxDevices && g_error_flag_host[device]) ? *g_error_flag_host[device] : 0;
}

const char* sepconv_tc_plan_ex(SepconvTcArgs* args, const SepconvTcDesc& d) {
    Params p;
    memset(&p, 0, sizeof(p));
    const int n = d.n, res = d.res, cin = d.cin, cout = d.cout;
    if (cin % kKBlock != 0 || cout % 64 != 0) return "cin must be a multiple of 64 and cout of 64";
    if (res < 4 || (res & (res - 1))) return "resolution must be a power of two >= 4";
    p.source = d.source;
    const bool dw = (d.source != SEPCONV_SRC_SPLIT);
    if (d.source == SEPCONV_SRC_SPLIT && !(d.a_hi && d.a_lo)) return "no A operand";
    if ((d.source == SEPCONV_SRC_NHWC || d.source == SEPCONV_SRC_UP) && !d.in_f32) return "no input tensor";
    if (d.source == SEPCONV_SRC_UP && !d.up_t) return "no low-resolution tensor for the up-sampling prologue";
    if (d.source == SEPCONV_SRC_STEM && !(d.stem_w && d.stem_b)) return "no fromrgb weights for the stem prologue";
    if ((d.source == SEPCONV_SRC_UP || d.source == SEPCONV_SRC_STEM) && res < 16) return "fused prologues need 8 x 16 tiles (res >= 16)";
    p.w9 = d.w9; p.bias = d.bias; p.noise = d.noise; p.inv_scale = d.inv_scale; p.out = d.out;
    p.n = n; p.H = res; p.W = res; p.cin = cin; p.cout = cout; p.act = d.act; p.passes = d.passes;
    p.n_tile = (cout == 64) ? 64 : 128;
    // Two 128-column accumulator regions can hold the two N halves of one M tile: the A operand (and for the fused
    // sources the whole prologue) is then produced once per pixel tile instead of once per N tile.
    p.nt_share = (cout >= 256 && (env_int("MIGAN_TC_NT_SHARE", 1) != 0 || (d.rgb && cout == 256))) ? 2 : 1;   // fused torgb at 256 needs both halves in one CTA
    p.num_kb = cin / kKBlock;
    if (dw) {   // spatial tiles with halo
        p.tile_w = res >= 16 ? 16 : res;
        p.tile_h = res >= 8 ? 8 : res;
    } else {    // linear tiles (rows of the [P][cin] operand) expressed as a 4-D box for the store
        p.tile_w = res >= 128 ? 128 : res;
        p.tile_h = std::min(res, kTileM / p.tile_w);
    }
    p.tile_n = kTileM / (p.tile_w * p.tile_h);
    p.tiles_x = res / p.tile_w; p.tiles_y = res / p.tile_h; p.tiles_n = (n + p.tile_n - 1) / p.tile_n;
    int dev0 = 0, sms0 = 148;
    cudaGetDevice(&dev0);
    cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
    // Few pixel tiles (the 4 x 4 ... 16 x 16 levels at small batches): a CTA that owns 256 output channels streams the whole
    // weight matrix and issues every MMA of its tile while most SMs idle.  Narrow N tiles spread the layer over 4 x the CTAs.
    if (!d.rgb && cout >= 256 && p.tiles_x * p.tiles_y * p.tiles_n * (cout / (p.n_tile * p.nt_share)) * 2 < sms0 &&
        env_int("MIGAN_TC_SMALL_N", 1) != 0) {
        p.n_tile = 64; p.nt_share = 1;
    }
    p.num_n_tiles = cout / (p.n_tile * p.nt_share);
    p.num_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.num_n_tiles;
    auto ilog2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
    p.l_tx = ilog2(p.tiles_x); p.l_ty = ilog2(p.tiles_y); p.l_nt = ilog2(p.num_n_tiles);

    // ---- input stage: IN area (+ T, NZ | XA) ----
    const uint32_t in_chunk_bytes = (uint32_t)(p.tile_n * (p.tile_h + 2) * (p.tile_w + 2) * kChunkC * 4);
    uint32_t stage_bytes = in_chunk_bytes;
    p.in_tx_bytes = in_chunk_bytes;
    p.up_has_noise = (d.source == SEPCONV_SRC_UP && d.up_noise) ? 1 : 0;
    if (d.source == SEPCONV_SRC_UP) {
        p.off_t = in_chunk_bytes;                 // 23040: 128-byte aligned
        p.off_aux = p.off_t + 6 * 10 * kChunkC * 4;
        stage_bytes = p.off_aux + 10 * kAuxW * 4;
        p.in_tx_bytes = in_chunk_bytes + 6 * 10 * kChunkC * 4 + (p.up_has_noise ? 10 * kAuxW * 4 : 0);
        memcpy(p.up_taps.f, d.up_taps, sizeof(p.up_taps.f));
    } else if (d.source == SEPCONV_SRC_STEM) {
        p.off_aux = in_chunk_bytes;
        stage_bytes = p.off_aux + 4 * 10 * kAuxW * 4;
        p.in_tx_bytes = 4 * 10 * kAuxW * 4;
        p.stem_w = d.stem_w; p.stem_b = d.stem_b;
    }
    p.in_stage_stride = (stage_bytes + 1023u) & ~1023u;

    p.store_out = 1;
    if (d.rgb) {
        if (p.num_n_tiles != 1 || cout > 256) return "fused torgb needs all output channels of a pixel tile in one CTA (cout <= 256)";
        p.torgb = 1; p.store_out = d.rgb->store_out;
        p.rgb_w = d.rgb->w; p.rgb_b = d.rgb->b; p.rgb_fir = d.rgb->fir; p.img_lo = d.rgb->img_lo; p.img_out = d.rgb->img_out;
    }

    // ---- shared-memory budget -> stage counts ----
    const uint32_t b_stage = (uint32_t)p.n_tile * kKBlock * 2 * 2;
    const uint32_t budget = kSmemLimit - kStaticSmem - 1024;   // static smem + alignment slack
    // The input ring is shared by the two prologue groups (even chunks -> group 0, odd -> group 1):
    // the stage count must be EVEN so that each stage always belongs to one group and every waiter
    // observes every phase of its barriers (parity waits alias otherwise).
    p.b_resident = (p.num_n_tiles == 1 && p.nt_share == 1 && (uint32_t)p.num_kb * b_stage <= 65536) ? 1 : 0;
    const uint32_t tables = dw ? 40u * cin + (d.source == SEPCONV_SRC_STEM ? 20u * cin : 0u) : 0u;
    const int epi_mode = env_int("MIGAN_TC_EPI", 0);   // 0 auto, 1 staged only, 2 direct
    bool found = false;
    // candidate configurations in order of preference: deeper input ring first, then more staging, then the
    // staging-free epilogue (256-bit stores straight from registers)
    struct Cand { int in_stages, epi_bufs, direct, b_stages, a_stages; };
    std::vector<Cand> cands;
    {
        const int a_st = dw ? 2 : 3;
        const int b_res = p.b_resident ? p.num_kb : 0;
        // input ring depths to try, deepest first (even: each prologue group owns every other stage).  Round-2 ncu: with two
        // stages per group the prologue warps of a 128 -> 128 layer wait for TMA data 29 % of the time.
        const int max_in = env_int("MIGAN_TC_MAX_IN", 6);
        std::vector<int> ins;
        if (dw) { for (int v = 6; v >= 2; v -= 2) if (v <= max_in) ins.push_back(v); } else ins.push_back(0);
        const int bopts[2] = {b_res ? b_res : 3, b_res ? b_res : 2};
        for (int bi = 0; bi < 2; ++bi) {
            if (bi == 1 && bopts[1] == bopts[0]) break;
            const int bs = bopts[bi];
            for (int in_st : ins) {
                if (!p.store_out) { cands.push_back({in_st, 0, 0, bs, a_st}); continue; }
                if (epi_mode != 2) { cands.push_back({in_st, 2, 0, bs, a_st}); cands.push_back({in_st, 1, 0, bs, a_st}); }
                if (epi_mode != 1) cands.push_back({in_st, 0, 1, bs, a_st});
            }
        }
    }
    uint32_t smem_bytes = 0;
    for (const Cand& c : cands) {
        const uint32_t epi = (uint32_t)c.epi_bufs * 4u * kEpiWarpBuf;
        const uint32_t total = (uint32_t)c.in_stages * p.in_stage_stride + (uint32_t)c.a_stages * kAStage + (uint32_t)c.b_stages * b_stage + epi + tables;
        if (total > budget) continue;
        p.in_stages = c.in_stages; p.a_stages = c.a_stages; p.b_stages = c.b_stages; p.epi_bufs = std::max(c.epi_bufs, 1);
        p.epi_direct = c.direct;
        p.off_in = 0;
        p.off_a = (uint32_t)c.in_stages * p.in_stage_stride;
        p.off_b = p.off_a + (uint32_t)c.a_stages * kAStage;
        p.off_epi = p.off_b + (uint32_t)c.b_stages * b_stage;
        p.off_dw = dw ? p.off_epi + epi : 0xFFFFFFFFu;
        p.off_stem = p.off_dw + 40u * cin;
        smem_bytes = total + 1024;
        found = true;
        break;
    }
    if (!found) return "shared memory budget exceeded";
    if (p.in_stages > 8 || p.a_stages > 4 || p.b_stages > 8) return "internal: ring too deep for the barrier table";
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.error_flag = (dev >= 0 && dev < kMaxDevices) ? g_error_flag[dev] : nullptr;
    p.prefetch = env_int("MIGAN_TC_PREFETCH", dw ? 4 : 2);

    // ---- tensor maps ----
    const char* err = nullptr;
    const uint64_t P = (uint64_t)n * res * res;
    if (d.source == SEPCONV_SRC_NHWC || d.source == SEPCONV_SRC_UP) {
        const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)res, (uint64_t)res, (uint64_t)n};
        const uint64_t str[3] = {(uint64_t)cin * 4, (uint64_t)res * cin * 4, (uint64_t)res * res * cin * 4};
        const uint32_t box[4] = {(uint32_t)kChunkC, (uint32_t)p.tile_w + 2, (uint32_t)p.tile_h + 2, (uint32_t)p.tile_n};
        if ((err = encode_map(&p.map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.in_f32, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return err;
    }
    if (d.source == SEPCONV_SRC_UP) {
        const int h = res / 2;
        const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)h, (uint64_t)h, (uint64_t)n};
        const uint64_t str[3] = {(uint64_t)cin * 4, (uint64_t)h * cin * 4, (uint64_t)h * h * cin * 4};
        const uint32_t box[4] = {(uint32_t)kChunkC, 10, 6, 1};
        if ((err = encode_map(&p.map_t, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.up_t, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return err;
        if (p.up_has_noise) {
            const uint64_t nd[2] = {(uint64_t)res, (uint64_t)res};
            const uint64_t ns[1] = {(uint64_t)res * 4};
            const uint32_t nb[2] = {(uint32_t)kAuxW, 10};
            if ((err = encode_map(&p.map_aux, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.up_noise, nd, ns, nb, CU_TENSOR_MAP_SWIZZLE_NONE))) return err;
        }
    }
    if (d.source == SEPCONV_SRC_SPLIT) {
        const uint64_t dims[2] = {(uint64_t)cin, P};
        const uint64_t str[1] = {(uint64_t)cin * 2};
        const uint32_t box[2] = {(uint32_t)kKBlock, (uint32_t)kTileM};
        if ((err = encode_map(&p.map_a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.a_hi, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return err;
        if ((err = encode_map(&p.map_a_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.a_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return err;
    }
    {
        const uint64_t dims[2] = {(uint64_t)cin, (uint64_t)cout};
        const uint64_t str[1] = {(uint64_t)cin * 2};
        const uint32_t box[2] = {(uint32_t)kKBlock, (uint32_t)p.n_tile};
        if ((err = encode_map(&p.map_w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.w_hi, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return err;
        if ((err = encode_map(&p.map_w_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d.w_lo, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return err;
    }
    if (d.out) {
        const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)res, (uint64_t)res, (uint64_t)n};
        const uint64_t str[3] = {(uint64_t)cout * 4, (uint64_t)res * cout * 4, (uint64_t)res * res * cout * 4};
        // one epilogue warp stores 32 consecutive pixel rows of the tile: quarter box
        const uint32_t qw = (uint32_t)std::min(p.tile_w, 32), qh = (uint32_t)std::min(p.tile_h, 32 / (int)qw), qn = 32u / (qw * qh);
        const uint32_t box[4] = {32u, qw, qh, qn};
        if ((err = encode_map(&p.map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.out, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B))) return err;
    } else if (p.store_out) {
        return "no output tensor";
    }
    // Persistent grid = one CTA per SM.  MIGAN_TC_RESERVE_SMS leaves SMs free for a concurrently running collective.
    const int reserve = std::max(0, std::min(env_int("MIGAN_TC_RESERVE_SMS", 0), sms / 2));
    args->grid = (unsigned)std::min(p.num_tiles, sms - reserve);
    args->smem_bytes = smem_bytes;
    args->num_tiles = p.num_tiles;
    args->bound_x = nullptr;
    memcpy(args->params_blob, &p, sizeof(p));
    return nullptr;
}

const char* sepconv_tc_plan(SepconvTcArgs* args, int passes, const float* in_f32, const __half* a_hi, const __half* a_lo,
                            const float* w9, const float* bias, const __half* w_hi, const __half* w_lo,
                            float inv_scale, const float* noise, float* out, int n, int res, int cin, int cout, int act,
                            const SepconvTcRgb* rgb) {
    SepconvTcDesc d;
    d.passes = passes;
    d.source = in_f32 ? SEPCONV_SRC_NHWC : SEPCONV_SRC_SPLIT;
    d.in_f32 = in_f32; d.a_hi = a_hi; d.a_lo = a_lo; d.w9 = w9; d.bias = bias; d.w_hi = w_hi; d.w_lo = w_lo;
    d.inv_scale = inv_scale; d.noise = noise; d.out = out; d.n = n; d.res = res; d.cin = cin; d.cout = cout; d.act = act; d.rgb = rgb;
    return sepconv_tc_plan_ex(args, d);
}

cudaError_t launch_sepconv_tc(SepconvTcArgs& a, cudaStream_t s, float* img_out_override, const float* x_nchw) {
    Params* pp = reinterpret_cast<Params*>(a.params_blob);
    if (pp->source == SEPCONV_SRC_STEM) {
        if (!x_nchw) return cudaErrorInvalidValue;
        if (a.bound_x != x_nchw) {    // the caller's x changes between calls: re-point the input tensor map
            if (encode_x_map(&pp->map_aux, x_nchw, pp->n, pp->H)) return cudaErrorInvalidValue;
            a.bound_x = x_nchw;
        }
    }
    Params p;
    memcpy(&p, a.params_blob, sizeof(p));
    if (img_out_override) p.img_out = img_out_override;
    return launch_pdl(sepconv_tc_kernel, dim3(a.grid), dim3(kThreads), a.smem_bytes, s, p);
}

}  // namespace migan
