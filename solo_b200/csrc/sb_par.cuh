// solo_b200 -- "one stream = one warp" execution model of the analysis stage (sb_coop.cuh is written against it).
//
// Two builds of the same source:
//   SB_COOP (device, sb_analysis.cu) : the 32 lanes of a warp cooperate on one stream.  SB_PARFOR spreads independent loop
//                                      iterations over the lanes, SB_SYNC() is __syncwarp(), the w* / g* collectives are
//                                      shuffle trees; arrays touched by more than one lane live in shared memory.
//   SB_EMU (host, tests/hostsim)     : 32 fibres per stream, run round-robin between barriers; shuffles and ballots go through a
//                                      shared scratch line.  A lane that reads what another lane has not written yet, or
//                                      that takes a different number of collectives, shows up deterministically on a machine
//                                      without a GPU (test infrastructure only).
// Without either macro the file only provides inert definitions, so the scalar headers can include it.
//
// Rules the cooperative routines follow:
//   * scalars are computed redundantly by every lane from shared data ("uniform" code: no divergence, no broadcast);
//   * state is read into registers before the SB_SYNC() that precedes any lane's write to it;
//   * a lane writes shared data either inside SB_PARFOR (its own iterations) or under an explicit lane test;
//   * integer sums that the reference accumulates with wrap-around (no saturation, no intermediate shift) may be split
//     across lanes and combined with wsum(): addition modulo 2^32 / 2^64 is associative, so the result is bit-identical;
//   * every lane of the warp executes every collective (loops around collectives have warp-uniform trip counts).
#pragma once
#include "sb_common.cuh"

#if defined(__CUDACC__) && defined(SB_COOP)
#define SB_COOP_ACTIVE 1
#define SB_NLANES 32
#define SB_LANE ((int)(threadIdx.x & 31))
#define SB_SYNC() __syncwarp()
#elif defined(SB_EMU) && !defined(__CUDA_ARCH__)
#define SB_COOP_ACTIVE 1
#define SB_NLANES 32
namespace sb { namespace emu {
extern int lane;
void barrier();
long long* scratch();   // 32 x 8-byte slots shared by the lanes of the current stream
} }
#define SB_LANE (::sb::emu::lane)
#define SB_SYNC() ::sb::emu::barrier()
#else
#define SB_COOP_ACTIVE 0
#define SB_NLANES 1
#define SB_LANE 0
#define SB_SYNC() ((void)0)
#endif

#define SB_LANE0 (SB_LANE == 0)
// Phase alignment: the warps of a block (one stream each) pass the analysis phases together, so that the instruction lines a
// phase needs are fetched once per SM and not once per warp (the per-stream code path is long and mostly straight-line).
#if defined(__CUDACC__) && defined(SB_PHASE_TIMING)
// development aid: cycle stamps of one block at every phase boundary (tools/phase_times.py)
__device__ long long sb_phase_stamp[512];
__device__ int sb_phase_line[512];
__device__ int sb_phase_count;
__device__ __forceinline__ void sb_mark(int line) {
    if (threadIdx.x == 0 && blockIdx.x == 300) { const int i = sb_phase_count; if (i < 512) { sb_phase_stamp[i] = clock64(); sb_phase_line[i] = line; sb_phase_count = i + 1; } }
}
#define SB_MARK() sb_mark(__LINE__)
#else
#define SB_MARK() ((void)0)
#endif
#if defined(__CUDA_ARCH__) && defined(SB_COOP) && defined(SB_PHASE_ALIGN)
#define SB_PHASE() do { __syncthreads(); SB_MARK(); } while (0)
#else
#define SB_PHASE() ((void)0)
#endif
// cooperative routines: device-only in the CUDA build, plain functions under the host emulation
#if defined(__CUDACC__)
#define SB_CFN __device__ inline
#else
#define SB_CFN inline
#endif
// iterations lo <= i < hi, spread over the lanes (all of them on the single lane of a serial build)
#define SB_PARFOR(i, lo, hi) for (int i = (lo) + SB_LANE; i < (hi); i += SB_NLANES)
// run a statement block on lane 0 only, then make its effects visible to the other lanes
#define SB_SERIAL(...) do { if (SB_LANE0) { __VA_ARGS__; } SB_SYNC(); } while (0)

namespace sb {

// ---- reductions / broadcasts over the lanes of a stream (identity in serial builds) --------------------------------
#if defined(__CUDACC__) && defined(SB_COOP)
__device__ __forceinline__ i32 wsum(i32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = (i32)((u32)v + (u32)__shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ i64 wsum64(i64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u32 lo = __shfl_xor_sync(0xffffffffu, (u32)v, o), hi = __shfl_xor_sync(0xffffffffu, (u32)((u64)v >> 32), o);
        v = (i64)((u64)v + (((u64)hi << 32) | lo));
    }
    return v;
}
__device__ __forceinline__ i32 wmax(i32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { i32 t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
__device__ __forceinline__ i32 wmin(i32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { i32 t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ i32 wbcast(i32 v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// general exchanges: value of lane `src` (per-lane source), of the lane d below / above, of lane ^ m; ballot
__device__ __forceinline__ i32 wshfl(i32 v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ i32 wshfl_up(i32 v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }      // lanes < d keep their own value
__device__ __forceinline__ i32 wshfl_down(i32 v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }  // lanes >= 32 - d keep their own value
__device__ __forceinline__ i32 wshfl_xor(i32 v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ i64 wshfl64(i64 v, int src) {
    u32 lo = __shfl_sync(0xffffffffu, (u32)v, src), hi = __shfl_sync(0xffffffffu, (u32)((u64)v >> 32), src);
    return (i64)(((u64)hi << 32) | lo);
}
__device__ __forceinline__ u32 wballot(bool p) { return __ballot_sync(0xffffffffu, p); }
// sums / extrema inside aligned groups of G lanes (G = 2, 4, 8, 16)
template <int G> __device__ __forceinline__ i32 gsum(i32 v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = (i32)((u32)v + (u32)__shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int G> __device__ __forceinline__ i64 gsum64(i64 v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        u32 lo = __shfl_xor_sync(0xffffffffu, (u32)v, o), hi = __shfl_xor_sync(0xffffffffu, (u32)((u64)v >> 32), o);
        v = (i64)((u64)v + (((u64)hi << 32) | lo));
    }
    return v;
}
template <int G> __device__ __forceinline__ i32 gmax(i32 v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) { i32 t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}
// minimum and, among equals, the smallest index, inside aligned groups of G lanes
template <int G> __device__ __forceinline__ void gargmin(i32& v, i32& idx) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        i32 tv = __shfl_xor_sync(0xffffffffu, v, o), ti = __shfl_xor_sync(0xffffffffu, idx, o);
        if (tv < v || (tv == v && ti < idx)) { v = tv; idx = ti; }
    }
}
// minimum value and, among equal values, the smallest index (what a first-minimum-wins scan returns)
__device__ __forceinline__ void wargmin(i32& v, i32& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        i32 tv = __shfl_xor_sync(0xffffffffu, v, o), ti = __shfl_xor_sync(0xffffffffu, idx, o);
        if (tv < v || (tv == v && ti < idx)) { v = tv; idx = ti; }
    }
}
__device__ __forceinline__ void wargmax(i32& v, i32& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        i32 tv = __shfl_xor_sync(0xffffffffu, v, o), ti = __shfl_xor_sync(0xffffffffu, idx, o);
        if (tv > v || (tv == v && ti < idx)) { v = tv; idx = ti; }
    }
}
#elif defined(SB_EMU) && !defined(__CUDA_ARCH__)
template <class T, class F> inline T emu_reduce(T v, F f) {
    long long* s = emu::scratch();
    s[emu::lane] = (long long)v;
    emu::barrier();
    T r = (T)s[0];
    for (int i = 1; i < 32; i++) r = f(r, (T)s[i]);
    emu::barrier();
    return r;
}
inline i32 wsum(i32 v) { return emu_reduce<i32>(v, [](i32 a, i32 b) { return (i32)((u32)a + (u32)b); }); }
inline i64 wsum64(i64 v) { return emu_reduce<i64>(v, [](i64 a, i64 b) { return (i64)((u64)a + (u64)b); }); }
inline i32 wmax(i32 v) { return emu_reduce<i32>(v, [](i32 a, i32 b) { return a > b ? a : b; }); }
inline i32 wmin(i32 v) { return emu_reduce<i32>(v, [](i32 a, i32 b) { return a < b ? a : b; }); }
inline i32 wbcast(i32 v, int src) {
    long long* s = emu::scratch();
    s[emu::lane] = v;
    emu::barrier();
    i32 r = (i32)s[src];
    emu::barrier();
    return r;
}
inline i32 wshfl(i32 v, int src) {
    long long* s = emu::scratch();
    s[emu::lane] = v;
    emu::barrier();
    i32 r = (i32)s[src & 31];
    emu::barrier();
    return r;
}
inline i32 wshfl_up(i32 v, int d) { return wshfl(v, emu::lane >= d ? emu::lane - d : emu::lane); }
inline i32 wshfl_down(i32 v, int d) { return wshfl(v, emu::lane + d < 32 ? emu::lane + d : emu::lane); }
inline i32 wshfl_xor(i32 v, int m) { return wshfl(v, emu::lane ^ m); }
inline i64 wshfl64(i64 v, int src) {
    long long* s = emu::scratch();
    s[emu::lane] = v;
    emu::barrier();
    i64 r = s[src & 31];
    emu::barrier();
    return r;
}
inline u32 wballot(bool p) {
    long long* s = emu::scratch();
    s[emu::lane] = p ? 1 : 0;
    emu::barrier();
    u32 r = 0;
    for (int i = 0; i < 32; i++) r |= (u32)s[i] << i;
    emu::barrier();
    return r;
}
template <class T, int G, class F> inline T emu_greduce(T v, F f) {
    long long* s = emu::scratch();
    s[emu::lane] = (long long)v;
    emu::barrier();
    const int b = emu::lane & ~(G - 1);
    T r = (T)s[b];
    for (int i = 1; i < G; i++) r = f(r, (T)s[b + i]);
    emu::barrier();
    return r;
}
template <int G> inline i32 gsum(i32 v) { return emu_greduce<i32, G>(v, [](i32 a, i32 b) { return (i32)((u32)a + (u32)b); }); }
template <int G> inline i64 gsum64(i64 v) { return emu_greduce<i64, G>(v, [](i64 a, i64 b) { return (i64)((u64)a + (u64)b); }); }
template <int G> inline i32 gmax(i32 v) { return emu_greduce<i32, G>(v, [](i32 a, i32 b) { return a > b ? a : b; }); }
template <int G> inline void gargmin(i32& v, i32& idx) {
    long long* s = emu::scratch();
    s[emu::lane] = ((long long)v << 32) | (u32)idx;
    emu::barrier();
    const int b = emu::lane & ~(G - 1);
    i32 bv = (i32)(s[b] >> 32), bi = (i32)(u32)s[b];
    for (int i = 1; i < G; i++) { i32 tv = (i32)(s[b + i] >> 32), ti = (i32)(u32)s[b + i]; if (tv < bv || (tv == bv && ti < bi)) { bv = tv; bi = ti; } }
    emu::barrier();
    v = bv; idx = bi;
}
inline void wargmin(i32& v, i32& idx) {
    long long* s = emu::scratch();
    s[emu::lane] = ((long long)v << 32) | (u32)idx;
    emu::barrier();
    i32 bv = (i32)(s[0] >> 32), bi = (i32)(u32)s[0];
    for (int i = 1; i < 32; i++) { i32 tv = (i32)(s[i] >> 32), ti = (i32)(u32)s[i]; if (tv < bv || (tv == bv && ti < bi)) { bv = tv; bi = ti; } }
    emu::barrier();
    v = bv; idx = bi;
}
inline void wargmax(i32& v, i32& idx) {
    long long* s = emu::scratch();
    s[emu::lane] = ((long long)v << 32) | (u32)idx;
    emu::barrier();
    i32 bv = (i32)(s[0] >> 32), bi = (i32)(u32)s[0];
    for (int i = 1; i < 32; i++) { i32 tv = (i32)(s[i] >> 32), ti = (i32)(u32)s[i]; if (tv > bv || (tv == bv && ti < bi)) { bv = tv; bi = ti; } }
    emu::barrier();
    v = bv; idx = bi;
}
#else
SB_HD i32 wsum(i32 v) { return v; }
SB_HD i64 wsum64(i64 v) { return v; }
SB_HD i32 wmax(i32 v) { return v; }
SB_HD i32 wmin(i32 v) { return v; }
SB_HD i32 wbcast(i32 v, int) { return v; }
SB_HD void wargmin(i32&, i32&) {}
SB_HD void wargmax(i32&, i32&) {}
SB_HD i32 wshfl(i32 v, int) { return v; }
SB_HD i64 wshfl64(i64 v, int) { return v; }
SB_HD u32 wballot(bool p) { return p ? 1u : 0u; }
#endif

// ---- per-stream serial work, transposed over the block ---------------------------------------------------------------
// A block of the analysis kernel holds SB_BLOCK_STREAMS streams, one warp and one shared-memory slot each.  Work that exists
// K times per stream as a short scalar recursion (K = 1: the signal-rate recurrences; K = 4: shaping windows, LTP
// sub-frames, interpolation candidates ...) would occupy K lanes of every warp; c_instances<K> instead hands all
// SB_BLOCK_STREAMS * K instances of the block to consecutive threads (stream = id / K), so that the issue slots it costs do
// not grow with the number of streams -- while another block of the SM is in a wide phase.  f(d, k): d = byte offset from
// the caller's own slot to the instance's slot (apply with xoff), k = instance number inside the stream.
#if defined(__CUDA_ARCH__) && defined(SB_COOP) && defined(SB_XPOSE)
#define SB_XPOSE_ACTIVE 1
#else
#define SB_XPOSE_ACTIVE 0
#endif
#if defined(__CUDACC__) && defined(SB_PHASE_TIMING)
__device__ __forceinline__ void sb_mark_end(int k) { sb_mark(-k); }
#else
SB_HD void sb_mark_end(int) {}
#endif
#if SB_COOP_ACTIVE
SB_CFN int sb_slot_bytes();      // stride of the per-stream slots (defined with the slot layout)
// (the slots live in shared memory; telling the compiler so keeps the accesses of the instance code on the shared-memory
//  path instead of generic loads / stores)
#if defined(__CUDA_ARCH__) && !defined(SB_NO_SHARED_HINT)
#define SB_ASSUME_SHARED(q) __builtin_assume(__isShared(q))
#else
#define SB_ASSUME_SHARED(q) ((void)0)
#endif
template <class T> SB_HD T* xoff(T* p, int d) { T* q = reinterpret_cast<T*>(reinterpret_cast<char*>(p) + d); SB_ASSUME_SHARED(q); return q; }
template <class T> SB_HD const T* xoff(const T* p, int d) { const T* q = reinterpret_cast<const T*>(reinterpret_cast<const char*>(p) + d); SB_ASSUME_SHARED(q); return q; }
template <int K, class F> SB_CFN void c_instances(F f) {
#if SB_XPOSE_ACTIVE
    __syncthreads();
    SB_MARK();
    const int id = (int)threadIdx.x;
    if (id < SB_BLOCK_STREAMS * K) { const int sj = id / K; f((sj - (id >> 5)) * sb_slot_bytes(), id - sj * K); }
    __syncthreads();
    sb_mark_end(K);
#else
    SB_SYNC();
    if (SB_LANE < K) f(0, SB_LANE);
    SB_SYNC();
#endif
}
#endif

}  // namespace sb
