// solo_b200 -- encoder stage A as a warp-per-stream kernel (sm_100a).
//
// This translation unit compiles the analysis routines in their cooperative form (SB_COOP, see sb_par.cuh): the 32 lanes
// of a warp work on ONE stream, whose persistent analysis state (EncCore, 3.2 KB) and per-packet working set live in
// shared memory for the duration of the packet.  State is moved between HBM and shared memory with 128-bit accesses,
// 512 contiguous bytes per warp instruction; the PCM row is read the same way.
#define SB_COOP 1
#include <cuda_runtime.h>
#include "sb_enc.cuh"

using namespace sb;

#ifndef SB_ANA_WARPS
#define SB_ANA_WARPS 2
#endif

namespace {

struct AnaSmem {
    EncCore st;
    EncAnalysisWork W;
    alignas(16) i16 pcm[PACKET];
};
static_assert(sizeof(EncCore) % 16 == 0, "EncCore is moved with 128-bit accesses");
static_assert(sizeof(EncState) % 16 == 0, "stream stride keeps EncCore 16-byte aligned");

__device__ __forceinline__ void copy16(void* dst, const void* src, int bytes, int lane) {
    const int4* s = reinterpret_cast<const int4*>(src);
    int4* d = reinterpret_cast<int4*>(dst);
    for (int i = lane; i < bytes / 16; i += 32) d[i] = s[i];
}

__global__ void __launch_bounds__(SB_ANA_WARPS * 32) sb_enc_analysis_warp_kernel(EncState* states, EncScratch* scratch, const i16* __restrict__ pcm, int spp, int n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * SB_ANA_WARPS + w;
    if (s >= n) return;
    AnaSmem* S = reinterpret_cast<AnaSmem*>(smem_raw) + w;
    copy16(&S->st, static_cast<EncCore*>(&states[s]), (int)sizeof(EncCore), lane);
    copy16(S->pcm, pcm + (size_t)s * spp, spp * 2, lane);
    __syncwarp();
    if (lane == 0) S->W.nlsf_fast = nullptr;
    __syncwarp();
    enc_packet_analysis(&S->st, &S->W, S->pcm, &scratch[s]);
    __syncwarp();
    copy16(static_cast<EncCore*>(&states[s]), &S->st, (int)sizeof(EncCore), lane);
}

}  // namespace

// called from the host code in solo_b200.cu; returns a CUDA error code
extern "C" int sb_launch_enc_analysis_warp(void* states, void* scratch, const void* pcm, int spp, int n, void* stream) {
    static bool configured = false;
    const int smem = SB_ANA_WARPS * (int)sizeof(AnaSmem);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sb_enc_analysis_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    sb_enc_analysis_warp_kernel<<<(n + SB_ANA_WARPS - 1) / SB_ANA_WARPS, SB_ANA_WARPS * 32, smem, (cudaStream_t)stream>>>(
        (EncState*)states, (EncScratch*)scratch, (const i16*)pcm, spp, n);
    return (int)cudaGetLastError();
}
