// simt_shim.h -- TEST INFRASTRUCTURE: runs the __device__ code of surtr_b200/csrc (clip_sub.cuh, the small tier of K3
// and the K4 moments) on the HOST, one warp at a time, so that the kernel source itself can be checked against the
// oracle without a GPU (tests/test_k3_emulation.py).
//
// The threads of a block (one warp, or several for the block-per-pair tiers) are ucontext coroutines on one OS thread.  A lane runs until it reaches a warp collective
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
enum Kind : int { K_NONE, K_SYNC, K_BALLOT, K_ANY, K_SHFL, K_SHFL_UP, K_SHFL_XOR, K_REDUCE_MAX, K_REDUCE_OR, K_SYNCTHREADS, K_SYNCTHREADS_OR };
enum Scope : int { S_WARP, S_BLOCK };

struct Block
{
    int n = 0;                                  // threads (a multiple of 32)
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stack;
    std::vector<char> done, waiting;
    std::vector<int> kind, scope;
    std::vector<uint64_t> op;                   // operand deposited by a waiting thread
    std::vector<uint64_t> wsnap;                // [warp][32] operands of the warp collective just released
    std::vector<uint64_t> bsnap;                // [n] operands of the block collective just released
    int cur = 0;
    unsigned long n_collectives = 0;
    std::function<void(int)> body;
};

inline Block*& current()
{
    static thread_local Block* b = nullptr;
    return b;
}

// Order in which the runnable threads are resumed between two collectives: 0 = ascending thread id, 1 = descending,
// any other value = a fresh pseudo-random permutation per pass seeded with it.  Code that is free of races between
// its barriers gives the same result under every order; a missing __syncwarp / __syncthreads shows up as a difference.
inline unsigned& schedule()
{
    static thread_local unsigned mode = 0;
    return mode;
}

inline int thread_id() { return current()->cur; }
inline int lane_id() { return current()->cur & 31; }

[[noreturn]] inline void die(const char* msg)
{
    std::fprintf(stderr, "simt emulation: %s\n", msg);
    std::abort();
}

// deposit an operand and wait until every thread of the scope (the caller's warp, or the block) has arrived
inline void arrive(int scope, int kind, uint64_t operand)
{
    Block& b = *current();
    const int t = b.cur;
    b.op[t] = operand;
    b.kind[t] = kind;
    b.scope[t] = scope;
    b.waiting[t] = 1;
    swapcontext(&b.ctx[t], &b.sched);
}
inline const uint64_t* warp_operands() { return current()->wsnap.data() + (size_t)(current()->cur >> 5) * WARP; }
inline const uint64_t* block_operands() { return current()->bsnap.data(); }

inline void trampoline()
{
    Block& b = *current();
    const int t = b.cur;
    b.body(t);
    b.done[t] = 1;
    swapcontext(&b.ctx[t], &b.sched);
}

// Runs body(tid) for the n threads of one block.  Threads run one after the other until each waits at a collective
// (or returns); a warp is released when its 32 lanes wait at the same warp collective, the block when all its threads
// wait at the same block collective.  Returns the number of collectives released.
inline unsigned long run_block(int n, const std::function<void(int)>& body)
{
    if (n <= 0 || n % WARP) die("block size must be a positive multiple of 32");
    Block b;
    b.n = n;
    b.body = body;
    b.ctx.resize(n); b.stack.resize(n);
    b.done.assign(n, 0); b.waiting.assign(n, 0); b.kind.assign(n, K_NONE); b.scope.assign(n, S_WARP); b.op.assign(n, 0);
    b.wsnap.assign(n, 0); b.bsnap.assign(n, 0);
    Block* saved = current();
    current() = &b;
    for (int t = 0; t < n; t++)
    {
        b.stack[t].resize(192 * 1024);
        getcontext(&b.ctx[t]);
        b.ctx[t].uc_stack.ss_sp = b.stack[t].data();
        b.ctx[t].uc_stack.ss_size = b.stack[t].size();
        b.ctx[t].uc_link = nullptr;
        makecontext(&b.ctx[t], (void (*)())trampoline, 0);
    }
    std::vector<int> order(n);
    for (int t = 0; t < n; t++) order[t] = schedule() == 1 ? n - 1 - t : t;
    unsigned rng = schedule() * 2654435761u + 12345u;
    while (true)
    {
        bool progressed = false;
        int n_done = 0;
        if (schedule() > 1)
            for (int i = n - 1; i > 0; i--)   // Fisher-Yates with an xorshift generator
            {
                rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5;
                std::swap(order[i], order[rng % (unsigned)(i + 1)]);
            }
        for (int k = 0; k < n; k++)
        {
            const int t = order[k];
            if (b.done[t]) { n_done++; continue; }
            if (b.waiting[t]) continue;
            b.cur = t;
            swapcontext(&b.sched, &b.ctx[t]);
            progressed = true;
            if (b.done[t]) n_done++;
        }
        if (n_done == n) break;
        // warp collectives
        for (int w = 0; w < n / WARP; w++)
        {
            int waiting = 0, finished = 0;
            for (int l = 0; l < WARP; l++)
            {
                const int t = w * WARP + l;
                finished += b.done[t];
                waiting += b.waiting[t] && b.scope[t] == S_WARP;
            }
            if (!waiting) continue;
            if (finished) die("some lanes of a warp returned while others wait at a warp collective");
            if (waiting != WARP) continue;   // the rest of the warp waits at a block collective or has not run yet
            for (int l = 1; l < WARP; l++)
                if (b.kind[w * WARP + l] != b.kind[w * WARP]) die("the lanes of a warp met at different collectives");
            for (int l = 0; l < WARP; l++)
            {
                b.wsnap[(size_t)w * WARP + l] = b.op[w * WARP + l];
                b.waiting[w * WARP + l] = 0;
            }
            b.n_collectives++;
            progressed = true;
        }
        // block collectives
        {
            int waiting = 0;
            for (int t = 0; t < n; t++) waiting += b.waiting[t] && b.scope[t] == S_BLOCK;
            if (waiting == n)
            {
                for (int t = 1; t < n; t++)
                    if (b.kind[t] != b.kind[0]) die("the threads of a block met at different barriers");
                for (int t = 0; t < n; t++) { b.bsnap[t] = b.op[t]; b.waiting[t] = 0; }
                b.n_collectives++;
                progressed = true;
            }
            else if (waiting && n_done) die("some threads returned while others wait at __syncthreads");
        }
        if (!progressed) die("deadlock: threads wait at collectives that can never complete");
    }
    current() = saved;
    return b.n_collectives;
}
inline unsigned long run_warp(const std::function<void(int)>& body) { return run_block(WARP, body); }

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
    simt::arrive(simt::S_WARP, simt::K_SYNC, 0);
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_BALLOT, pred ? 1u : 0u);
    const uint64_t* in = simt::warp_operands();
    unsigned b = 0;
    for (int l = 0; l < simt::WARP; l++) b |= (unsigned)(in[l] & 1u) << l;
    return b;
}
inline int __any_sync(unsigned mask, int pred)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_ANY, pred ? 1u : 0u);
    const uint64_t* in = simt::warp_operands();
    for (int l = 0; l < simt::WARP; l++)
        if (in[l]) return 1;
    return 0;
}
inline int __reduce_max_sync(unsigned mask, int v)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_REDUCE_MAX, simt::pack(v));
    const uint64_t* in = simt::warp_operands();
    int m = simt::unpack<int>(in[0]);
    for (int l = 1; l < simt::WARP; l++) m = std::max(m, simt::unpack<int>(in[l]));
    return m;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_REDUCE_MAX, simt::pack(v));
    const uint64_t* in = simt::warp_operands();
    unsigned m = simt::unpack<unsigned>(in[0]);
    for (int l = 1; l < simt::WARP; l++) m = std::max(m, simt::unpack<unsigned>(in[l]));
    return m;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_REDUCE_OR, simt::pack(v));
    const uint64_t* in = simt::warp_operands();
    unsigned m = 0;
    for (int l = 0; l < simt::WARP; l++) m |= simt::unpack<unsigned>(in[l]);
    return m;
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_SHFL, simt::pack(v));
    const int me = simt::lane_id();
    const int from = (me & ~(width - 1)) | (src & (width - 1));
    return simt::unpack<T>(simt::warp_operands()[from]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_SHFL_UP, simt::pack(v));
    const int me = simt::lane_id();
    const int from = me - (int)delta;
    if (from < (me & ~(width - 1))) return v;   // below the segment: the lane keeps its own value
    return simt::unpack<T>(simt::warp_operands()[from]);
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32)
{
    require_full(mask);
    simt::arrive(simt::S_WARP, simt::K_SHFL_XOR, simt::pack(v));
    const int me = simt::lane_id();
    const int from = me ^ lane_mask;
    if ((from & ~(width - 1)) != (me & ~(width - 1))) return v;
    return simt::unpack<T>(simt::warp_operands()[from]);
}
// ---- block barriers ----
inline void __syncthreads() { simt::arrive(simt::S_BLOCK, simt::K_SYNCTHREADS, 0); }
inline int __syncthreads_or(int pred)
{
    simt::arrive(simt::S_BLOCK, simt::K_SYNCTHREADS_OR, pred ? 1u : 0u);
    const uint64_t* in = simt::block_operands();
    for (int t = 0; t < simt::current()->n; t++)
        if (in[t]) return 1;
    return 0;
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
