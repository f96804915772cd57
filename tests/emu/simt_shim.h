// simt_shim.h -- TEST INFRASTRUCTURE: runs the __device__ code of surtr_b200/csrc (clip_sub.cuh, the small tier of K3
// and the K4 moments) on the HOST, one warp at a time, so that the kernel source itself can be checked against the
// oracle without a GPU (tests/test_k3_emulation.py).
//
// The 32 lanes of a warp are 32 ucontext coroutines on one OS thread.  A lane runs until it reaches a warp collective
// (__ballot_sync, __shfl_*_sync, __any_sync, __reduce_max_sync, __syncwarp -- the kernels only ever use the full member
// mask), deposits its operand and yields; when all 32 lanes have arrived at the SAME collective they are resumed and
// each reads what it needs from the deposited operands.  Between two collectives the lanes therefore run one after the
// other (lane 0 first) -- one of the interleavings the GPU may produce, and the only kind of ordering the kernels rely on
// (they separate conflicting shared-memory accesses by __syncwarp).  A lane that returns while others wait at a
// collective, or lanes meeting at different collectives, abort the run: that would be a deadlock or undefined
// behaviour on the device.
//
// The arithmetic intrinsics map to single IEEE operations (build with -ffp-contract=off on x86-64 SSE2, as the oracle).
#pragma once

#include <cuda_runtime.h>   // float4 / make_float4 / __device__ (empty) for the host compiler; no device code is generated
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
// (__noinline__ is defined by the including file around the device headers only: libstdc++ spells the GNU attribute
// __attribute__((__noinline__)) itself, so the macro must not be visible to standard headers)

namespace simt
{
constexpr int WARP = 32;
enum Kind : int { K_NONE, K_SYNC, K_BALLOT, K_ANY, K_SHFL, K_SHFL_UP, K_SHFL_XOR, K_REDUCE_MAX };

struct Warp
{
    ucontext_t sched;
    ucontext_t lane[WARP];
    std::vector<char> stack[WARP];
    bool done[WARP];
    int cur = 0;
    unsigned long seq[WARP];      // collectives this lane has arrived at
    uint64_t buf[2][WARP];        // operands, double buffered by collective parity
    int kind[2][WARP];
    unsigned long n_collectives = 0;
    std::function<void(int)> body;
};

inline Warp*& current()
{
    static thread_local Warp* w = nullptr;
    return w;
}

inline int lane_id() { return current()->cur; }

[[noreturn]] inline void die(const char* msg)
{
    std::fprintf(stderr, "simt emulation: %s\n", msg);
    std::abort();
}

// deposit an operand, wait for the other 31 lanes, return the slot parity to read from
inline int arrive(int kind, uint64_t operand)
{
    Warp& w = *current();
    const int l = w.cur;
    const int par = (int)(w.seq[l] & 1ul);
    w.buf[par][l] = operand;
    w.kind[par][l] = kind;
    w.seq[l]++;
    swapcontext(&w.lane[l], &w.sched);
    return par;
}

inline void trampoline()
{
    Warp& w = *current();
    const int l = w.cur;
    w.body(l);
    w.done[l] = true;
    swapcontext(&w.lane[l], &w.sched);
}

// Runs body(lane) for the 32 lanes of one warp in lock step at the collectives.  Returns the number of collectives.
inline unsigned long run_warp(const std::function<void(int)>& body)
{
    Warp w;
    w.body = body;
    Warp* saved = current();
    current() = &w;
    for (int l = 0; l < WARP; l++)
    {
        w.stack[l].resize(256 * 1024);
        w.done[l] = false;
        w.seq[l] = 0;
        getcontext(&w.lane[l]);
        w.lane[l].uc_stack.ss_sp = w.stack[l].data();
        w.lane[l].uc_stack.ss_size = w.stack[l].size();
        w.lane[l].uc_link = nullptr;
        makecontext(&w.lane[l], (void (*)())trampoline, 0);
    }
    while (true)
    {
        int n_done = 0;
        for (int l = 0; l < WARP; l++)
        {
            if (w.done[l]) { n_done++; continue; }
            w.cur = l;
            swapcontext(&w.sched, &w.lane[l]);
            if (w.done[l]) n_done++;
        }
        if (n_done == WARP) break;
        if (n_done != 0) die("some lanes returned while others wait at a warp collective");
        const int par = (int)((w.seq[0] - 1) & 1ul);
        for (int l = 1; l < WARP; l++)
            if (w.seq[l] != w.seq[0] || w.kind[par][l] != w.kind[par][0]) die("lanes met at different warp collectives");
        w.n_collectives++;
    }
    current() = saved;
    return w.n_collectives;
}

template <class T> inline uint64_t pack(T v)
{
    static_assert(sizeof(T) <= 8, "operand too wide");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T> inline T unpack(uint64_t u)
{
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
} // namespace simt

// ---- warp collectives (full member mask only) ----
inline void require_full(unsigned mask)
{
    if (mask != 0xffffffffu) simt::die("only full-mask collectives are emulated");
}
inline void __syncwarp(unsigned mask = 0xffffffffu)
{
    require_full(mask);
    simt::arrive(simt::K_SYNC, 0);
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_BALLOT, pred ? 1u : 0u);
    unsigned b = 0;
    for (int l = 0; l < simt::WARP; l++) b |= (unsigned)(simt::current()->buf[par][l] & 1u) << l;
    return b;
}
inline int __any_sync(unsigned mask, int pred)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_ANY, pred ? 1u : 0u);
    for (int l = 0; l < simt::WARP; l++)
        if (simt::current()->buf[par][l]) return 1;
    return 0;
}
inline int __reduce_max_sync(unsigned mask, int v)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_REDUCE_MAX, simt::pack(v));
    int m = simt::unpack<int>(simt::current()->buf[par][0]);
    for (int l = 1; l < simt::WARP; l++) m = std::max(m, simt::unpack<int>(simt::current()->buf[par][l]));
    return m;
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_SHFL, simt::pack(v));
    const int me = simt::lane_id();
    const int from = (me & ~(width - 1)) | (src & (width - 1));
    return simt::unpack<T>(simt::current()->buf[par][from]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_SHFL_UP, simt::pack(v));
    const int me = simt::lane_id();
    const int from = me - (int)delta;
    if (from < (me & ~(width - 1))) return v;   // below the segment: the lane keeps its own value
    return simt::unpack<T>(simt::current()->buf[par][from]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32)
{
    require_full(mask);
    const int par = simt::arrive(simt::K_SHFL_XOR, simt::pack(v));
    const int me = simt::lane_id();
    const int from = me ^ lane_mask;
    if ((from & ~(width - 1)) != (me & ~(width - 1))) return v;
    return simt::unpack<T>(simt::current()->buf[par][from]);
}

// ---- integer / bit intrinsics ----
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s)
{
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
    {
        const unsigned sel = (s >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;   // replicate the sign bit
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __vcmpeq4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}

// ---- float intrinsics: one IEEE-754 binary32 operation each, round to nearest even ----
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fsqrt_rn(float a) { volatile float r = std::sqrt(a); return r; }
inline float __uint_as_float(unsigned u) { return simt::unpack<float>(u); }
inline unsigned __float_as_uint(float f) { return (unsigned)simt::pack(f); }
inline int __float_as_int(float f) { return (int)(unsigned)simt::pack(f); }
inline float __int_as_float(int i) { return simt::unpack<float>((unsigned)i); }
template <class T> inline T __ldg(const T* p) { return *p; }

using std::max;
using std::min;
