// Shared device helpers for the b200chan kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rcb {

constexpr int kNumSMsB200 = 148;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

// atan2f with <= ~1.5e-7 rad absolute error: octant fold, MUFU.RCP division, degree-6 (in z^2)
// minimax polynomial (coefficients fitted and float32-verified by the script recorded in
// DESIGN.md; max |err| of the polynomial evaluated in float32 is 1.1e-7 on [0,1]).
// Replaces gnuradio-runtime/lib/math/fast_atan2f.cc (table lookup, 1.25e-6 rad error) as called by
// gr-analog quadrature_demod_cf (reference: moto_control_demod.py:105, p25_control_demod.py:121,
// edacs_control_demod.py:82-84, logging_receiver.py:234).  (0,0) -> 0 like the reference.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float z = __fdividef(mn, mx);
    z = (mx == 0.0f) ? 0.0f : z;
    const float s = z * z;
    float r = -0.004370174370706081f;
    r = fmaf(r, s, 0.023092154413461685f);
    r = fmaf(r, s, -0.05784549191594124f);
    r = fmaf(r, s, 0.0979914739727974f);
    r = fmaf(r, s, -0.13978290557861328f);
    r = fmaf(r, s, 0.1996297985315323f);
    r = fmaf(r, s, -0.33331674337387085f);
    r = r * s;
    r = fmaf(r, z, z);                          // atan(z), z in [0,1]
    r = (ay > ax) ? (1.57079632679489661923f - r) : r;
    r = (x < 0.0f) ? (3.14159265358979323846f - r) : r;
    return copysignf(r, y);
}

// 256-bit global store (sm_100+: st.global.v8.f32) - one full 32 B sector per lane.
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ void st_global_v8_hint(float* p, const float (&v)[8], uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ float2 ldg_f2_hint(const float2* p, uint64_t policy) {
    float2 r;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// streaming (read-once) 64-bit load: bypass L1 allocation
__device__ __forceinline__ float2 ld_stream_f2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

}  // namespace rcb
