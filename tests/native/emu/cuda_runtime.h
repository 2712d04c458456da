// TEST INFRASTRUCTURE — a small CPU emulation of the CUDA execution model, just large enough to run this repository's
// kernels (rmqtt_b200/csrc/*.cuh, compiled with -DGM_CPU_EMU) under g++ and the sanitizers on a box without a GPU.
// It shadows <cuda_runtime.h> for the test build only (tests/native/emu comes first on the include path); nothing under
// rmqtt_b200/ ever includes or links it, and libgpumqtt.so is built from the same kernel sources WITHOUT GM_CPU_EMU
// (the PTX helpers then are the real ones; `cuobjdump -sass` of the library is byte-identical with and without the
// #ifdef blocks).
//
// Model: one CTA at a time; every thread of the CTA is a fiber (ucontext) on ONE OS thread.  A fiber runs until it reaches
// a barrier (__syncthreads) or a warp collective (__shfl*_sync, __ballot_sync, __reduce*_sync, __syncwarp), where it waits
// for the other threads of its CTA / lanes of its warp — exactly the points where CUDA code may assume nothing about the
// others' progress.  Lanes that have exited count as arrived.  Atomics are plain operations (one OS thread), so logic,
// indexing and memory safety are checked here; memory-ORDERING bugs are not (compute-sanitizer racecheck on the GPU is).
// A CTA that cannot make progress (a deadlock: e.g. a collective inside divergent code that not all lanes reach) aborts
// with a message instead of hanging.
#pragma once
#ifndef GM_CPU_EMU
#error "tests/native/emu/cuda_runtime.h is for the -DGM_CPU_EMU test build only"
#endif
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define EMU_ASAN 1
#endif

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() = default; dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
typedef void* cudaStream_t;

namespace emu {

constexpr unsigned MAX_THREADS = 1024;
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool started = false, finished = false;
    dim3 tid;
    const void* asan_bottom = nullptr; size_t asan_size = 0;
};
struct WarpState { uint64_t val[32], snap[32]; uint32_t arrived = 0, exited = 0, gen = 0; };
struct State {
    dim3 block_idx, block_dim, grid_dim;
    Fiber fibers[MAX_THREADS];
    WarpState warps[MAX_THREADS / 32];
    unsigned nthreads = 0, n_finished = 0, bar_arrived = 0, bar_gen = 0;
    Fiber* cur = nullptr;
    ucontext_t sched;
    uint64_t progress = 0;
    std::function<void()> body;
    alignas(128) unsigned char dyn[232 * 1024];
    size_t stack_bytes = 512 * 1024;
};
inline State& S() { static State* s = new State(); return *s; }
inline unsigned char* dyn_smem() { return S().dyn; }

inline void switch_to_sched() {
    State& s = S();
    Fiber* f = s.cur;
#ifdef EMU_ASAN
    void* fake = nullptr;
    __sanitizer_start_switch_fiber(f->finished ? nullptr : &fake, nullptr, 0);      // (the scheduler runs on the thread's own stack)
#endif
    swapcontext(&f->ctx, &s.sched);
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
}
inline void yield() { switch_to_sched(); }

inline void complete_warp_if_ready(WarpState& w) {
    if (w.arrived && (w.arrived | w.exited) == 0xFFFFFFFFu) { std::memcpy(w.snap, w.val, sizeof(w.val)); w.arrived = 0; w.gen++; S().progress++; }
}
inline void complete_barrier_if_ready() {
    State& s = S();
    if (s.bar_arrived && s.bar_arrived + s.n_finished == s.nthreads) { s.bar_arrived = 0; s.bar_gen++; s.progress++; }
}

// every lane of the warp contributes `v`; returns the 32 contributions (valid until this lane's next collective)
inline const uint64_t* collective(unsigned mask, uint64_t v) {
    if (mask != 0xFFFFFFFFu) { fprintf(stderr, "emu: only full-warp collectives are modelled (mask %08x)\n", mask); std::abort(); }
    State& s = S();
    const unsigned t = s.cur->tid.x, lane = t & 31u;
    WarpState& w = s.warps[t >> 5];
    const uint32_t gen = w.gen;
    w.val[lane] = v;
    w.arrived |= 1u << lane;
    s.progress++;
    complete_warp_if_ready(w);
    while (w.gen == gen) yield();
    return w.snap;
}
inline void barrier() {
    State& s = S();
    const unsigned gen = s.bar_gen;
    s.bar_arrived++;
    s.progress++;
    complete_barrier_if_ready();
    while (s.bar_gen == gen) yield();
}

inline void trampoline() {
    State& s = S();
#ifdef EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, nullptr, nullptr);
#endif
    s.body();
    Fiber* f = s.cur;
    f->finished = true;
    s.n_finished++;
    s.progress++;
    WarpState& w = s.warps[f->tid.x >> 5];
    w.exited |= 1u << (f->tid.x & 31u);
    complete_warp_if_ready(w);
    complete_barrier_if_ready();
    switch_to_sched();
    std::abort();      // a finished fiber is never resumed
}

// grid.x CTAs of block.x threads (1-D launches only: all this repository uses); `body` calls the kernel with its arguments
template <class F>
inline void launch(dim3 grid, dim3 block, F&& body) {
    State& s = S();
    if (block.x == 0 || block.x > MAX_THREADS || block.x % 32 != 0 && block.x != 1) { fprintf(stderr, "emu: unsupported block size %u\n", block.x); std::abort(); }
    s.block_dim = block; s.grid_dim = grid; s.nthreads = block.x;
    s.body = std::function<void()>(body);
    for (unsigned t = 0; t < block.x; ++t)
        if (!s.fibers[t].stack) { s.fibers[t].stack = static_cast<char*>(std::malloc(s.stack_bytes)); if (!s.fibers[t].stack) std::abort(); }
    for (unsigned b = 0; b < grid.x; ++b) {
        s.block_idx = dim3(b);
        s.n_finished = 0; s.bar_arrived = 0;
        for (unsigned wi = 0; wi < (block.x + 31) / 32; ++wi) { s.warps[wi].arrived = 0; s.warps[wi].exited = block.x == 1 ? 0xFFFFFFFEu : 0u; }
        for (unsigned t = 0; t < block.x; ++t) {
            Fiber& f = s.fibers[t];
            f.started = false; f.finished = false; f.tid = dim3(t);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = s.stack_bytes; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, reinterpret_cast<void (*)()>(trampoline), 0);
        }
        while (s.n_finished < block.x) {
            const uint64_t before = s.progress;
            for (unsigned t = 0; t < block.x; ++t) {
                Fiber& f = s.fibers[t];
                if (f.finished) continue;
                s.cur = &f;
#ifdef EMU_ASAN
                void* fake = nullptr;
                __sanitizer_start_switch_fiber(&fake, f.stack, s.stack_bytes);
#endif
                swapcontext(&s.sched, &f.ctx);
#ifdef EMU_ASAN
                __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
            }
            if (s.progress == before) {
                fprintf(stderr, "emu: CTA %u cannot make progress (deadlock): %u of %u threads finished, %u at the barrier;", b, s.n_finished, block.x, s.bar_arrived);
                for (unsigned wi = 0; wi < (block.x + 31) / 32; ++wi) fprintf(stderr, " warp %u arrived %08x exited %08x", wi, s.warps[wi].arrived, s.warps[wi].exited);
                fprintf(stderr, "\n");
                std::abort();
            }
        }
    }
    s.cur = nullptr;
}

}  // namespace emu

// ---- the CUDA surface the kernels use -----------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__          /* (libstdc++ spells the GNU attribute __noinline__: it must expand to nothing there) */
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().block_idx)
#define blockDim (emu::S().block_dim)
#define gridDim (emu::S().grid_dim)

template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) { using T = typename std::common_type<A, B>::type; return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b); }
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) { using T = typename std::common_type<A, B>::type; return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(b) : static_cast<T>(a); }

inline void __syncthreads() { emu::barrier(); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::collective(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { emu::S().progress++; emu::yield(); }      // (a polling loop: let the others run)
inline long long clock64() { static long long c = 0; return c += 20000000; }   // (bounded device-side waits end after a few hundred polls)

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of up to 64 bits");
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    const uint64_t* all = emu::collective(mask, raw);
    T out; std::memcpy(&out, &all[static_cast<unsigned>(src) & 31u], sizeof(T));
    return out;
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int = 32) {
    const unsigned lane = threadIdx.x & 31u;
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    const uint64_t* all = emu::collective(mask, raw);
    T out; std::memcpy(&out, &all[lane >= delta ? lane - delta : lane], sizeof(T));
    return out;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int = 32) {
    const unsigned lane = threadIdx.x & 31u;
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    const uint64_t* all = emu::collective(mask, raw);
    T out; std::memcpy(&out, &all[(lane ^ static_cast<unsigned>(lanemask)) & 31u], sizeof(T));
    return out;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t* all = emu::collective(mask, pred ? 1u : 0u);
    const emu::WarpState& w = emu::S().warps[threadIdx.x >> 5];
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if (!(w.exited >> l & 1u) && all[l]) r |= 1u << l;
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::collective(mask, v);
    const emu::WarpState& w = emu::S().warps[threadIdx.x >> 5];
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if (!(w.exited >> l & 1u)) r = std::max<unsigned>(r, static_cast<unsigned>(all[l]));
    return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    const uint64_t* all = emu::collective(mask, v);
    const emu::WarpState& w = emu::S().warps[threadIdx.x >> 5];
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if (!(w.exited >> l & 1u)) r += static_cast<unsigned>(all[l]);
    return r;
}

template <class T, class U> inline T atomicAdd(T* p, U v) { T o = *p; *p = static_cast<T>(o + static_cast<T>(v)); return o; }
template <class T, class U> inline T atomicOr(T* p, U v) { T o = *p; *p = static_cast<T>(o | static_cast<T>(v)); return o; }
template <class T, class U> inline T atomicExch(T* p, U v) { T o = *p; *p = static_cast<T>(v); return o; }
template <class T, class U> inline T atomicMax(T* p, U v) { T o = *p; if (static_cast<T>(v) > o) *p = static_cast<T>(v); return o; }
template <class T, class U, class V> inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == static_cast<T>(cmp)) *p = static_cast<T>(v); return o; }

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return static_cast<unsigned>(((static_cast<uint64_t>(hi) << 32) | lo) >> (shift & 31u)); }
