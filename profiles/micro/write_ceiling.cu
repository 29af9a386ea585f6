// Write-only ceiling for the geometry of craft_rollout_kernel: `grid` CTAs of `block` threads, of
// which the last `store_warps` warps write; CTA b writes `cta_bytes` contiguous bytes at
// frame + b*cta_bytes for each of `frames` frames (frame f at base + (f % ring) * frame_bytes),
// 16 bytes per lane per store (st.global.cs.v4), straight from registers.  No shared memory, no
// compute: what the memory system delivers for this store pattern.
#include <cstdint>
#include <cuda_runtime.h>

// interleaved != 0: chunk c of storing warp sw of CTA b is global chunk c*(grid*store_warps) +
// b*store_warps + sw, i.e. at every step the whole grid writes ONE contiguous region.
__global__ void write_ceiling_kernel(float4 *base, long frame_f4, int cta_f4, int frames, int ring,
                                     int store_warps, int interleaved) {
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sw = warp - (warps - store_warps);
    if (sw < 0) return;
    const float4 v = make_float4(1.f, 0.f, 0.f, (float)blockIdx.x);
    for (int f = 0; f < frames; f++) {
        float4 *dst = base + (long)(f % ring) * frame_f4 + (long)blockIdx.x * cta_f4;
        // warp sw takes the chunks sw, sw + store_warps, ... of 404 float4 (6,464 bytes) each
        if (interleaved == 2) {
            // frames innermost: chunk c goes to every frame before chunk c + store_warps is touched
            if (f > 0) break;
            for (int c = sw * 404; c < cta_f4; c += store_warps * 404)
                for (int ff = 0; ff < frames; ff++) {
                    float4 *d = base + (long)(ff % ring) * frame_f4 + (long)blockIdx.x * cta_f4 + c;
                    for (int i = lane; i < 404 && c + i < cta_f4; i += 32) __stcs(d + i, v);
                }
            continue;
        }
        if (interleaved) {
            const int per_warp = cta_f4 / 404 / store_warps;
            float4 *fr = base + (long)(f % ring) * frame_f4;
            for (int c = 0; c < per_warp; c++) {
                float4 *d = fr + ((long)c * gridDim.x * store_warps + blockIdx.x * store_warps + sw) * 404;
                for (int i = lane; i < 404; i += 32) __stcs(d + i, v);
            }
            continue;
        }
        for (int c = sw * 404; c < cta_f4; c += store_warps * 404)
            for (int i = lane; i < 404 && c + i < cta_f4; i += 32) __stcs(dst + c + i, v);
    }
}

extern "C" int write_ceiling(void *base, long frame_bytes, int grid, int block, int cta_bytes,
                             int frames, int ring, int store_warps, int interleaved, void *stream) {
    write_ceiling_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        (float4 *)base, frame_bytes / 16, cta_bytes / 16, frames, ring, store_warps, interleaved);
    return (int)cudaGetLastError();
}
