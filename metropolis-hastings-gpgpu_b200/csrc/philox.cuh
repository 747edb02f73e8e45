// philox.cuh -- counter-based Philox4x32-10 in registers, and the uniform / Box-Muller
// transforms of the chain's random stream.
//
// Replaces the reference's per-thread XORWOW state array and its init kernel
// (Kernel.cu:152-160, 939-943): no RNG state in memory, no set-up launch, and a chain's
// stream depends only on (seed, global chain id, iteration), never on the launch shape.
//
// Stream spec (the test oracle implements the same spec independently):
//   key     = (seed lo, seed hi)
//   counter = (iteration lo, draw block + (iteration hi << 16), chain lo, chain hi)
//   block 0 : w0 -> move type, w1 -> first object, w2/w3 -> Box-Muller pair (translate,
//             rotate) or w2 -> second object (swap)
//   block 1 : w0 -> acceptance uniform
//   block 2+t: t-th re-draw while a picked object is frozen (w0 first, w1 second object)
//   block 0xFFFF: w0 -> replica-exchange uniform of the pair whose lower chain this is
#pragma once
#include <stdint.h>

namespace mh {

struct Philox4 {
    uint32_t x, y, z, w;
};

// UNROLL = 10: straight-line code (the full-scan kernels, whose n^2 loop dwarfs it); UNROLL = 2: a short
// loop for the delta kernel, whose per-iteration code path must stay inside the 32 KB instruction cache.
template <int UNROLL>
__device__ __noinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll UNROLL
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

template <int UNROLL = 10>
__device__ __forceinline__ Philox4 draw_block(uint64_t seed, uint64_t chain, uint64_t it, uint32_t block)
{
    return philox4x32_10<UNROLL>((uint32_t)it, block + ((uint32_t)(it >> 32) << 16), (uint32_t)chain, (uint32_t)(chain >> 32),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// (0, 1], the transform of curand_uniform.h:69-72.  The multiply is by 2^-32 and therefore
// exact, so the value does not depend on FMA contraction.
__device__ __forceinline__ float uniform01(uint32_t x) { return __fmaf_rn((float)x, 0x1p-32f, 0x1p-33f); }

// curand_normal.h:70-92: first normal = s sin v, second = s cos v, v = 2 pi (y + 0.5) / 2^32.
// MH_BOX_MULLER_PI (default): sin and cos of v through sincospif((y + 0.5) / 2^31) -- the argument is exact and needs
// no reduction by pi, which halves the code of this per-iteration helper (the memo kernel is sensitive to the length of
// its per-iteration path); the normals differ from cuRAND's float-rounded v by <= 4e-7 s.
#ifndef MH_BOX_MULLER_PI
#define MH_BOX_MULLER_PI 1
#endif
__device__ __noinline__ float2 box_muller_pair(uint32_t x, uint32_t y)
{
    const float u = uniform01(x);
    const float s = sqrtf(-2.0f * logf(u));
    float sn, cs;
#if MH_BOX_MULLER_PI
    sincospif(__fmaf_rn((float)y, 0x1p-31f, 0x1p-32f), &sn, &cs);
#else
    const float k = 1.46291807e-09f; // 2 pi / 2^32
    const float v = __fmaf_rn((float)y, k, k / 2.0f);
    sincosf(v, &sn, &cs);
#endif
    return make_float2(s * sn, s * cs);
}
__device__ __forceinline__ void box_muller(uint32_t x, uint32_t y, float &n0, float &n1)
{
    const float2 p = box_muller_pair(x, y);
    n0 = p.x;
    n1 = p.y;
}

// Kernel.cu:566-574 with u supplied: trunc(u * (max - min + 0.999999) + min), the scale in
// double like the reference; quirk Q13 (index max+1 when u == 1) is clamped away.
__device__ __forceinline__ int random_int(float u, int maxv)
{
    float p = (float)((double)u * ((double)maxv + 0.999999));
    int v = (int)truncf(p);
    return v > maxv ? maxv : v;
}

} // namespace mh
