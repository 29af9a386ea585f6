/* How fast can the GPU box's host cores widen a u8 feature frame to f32?  (Decides whether the
 * host-facing tick should ship bytes over PCIe and widen on the host: it pays only if this beats the
 * ~56 GB/s the f32 frame gets over PCIe.)   gcc -O3 -fopenmp -mavx2 host_widen_probe.c -o probe */
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void widen(const uint8_t *src, float *dst, size_t n, int nt_store) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m128i a = _mm_loadu_si128((const __m128i *)(src + i));
        __m128i b = _mm_loadu_si128((const __m128i *)(src + i + 16));
        __m256 f0 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(a));
        __m256 f1 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(a, 8)));
        __m256 f2 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(b));
        __m256 f3 = _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(b, 8)));
        if (nt_store) {
            _mm256_stream_ps(dst + i, f0); _mm256_stream_ps(dst + i + 8, f1);
            _mm256_stream_ps(dst + i + 16, f2); _mm256_stream_ps(dst + i + 24, f3);
        } else {
            _mm256_storeu_ps(dst + i, f0); _mm256_storeu_ps(dst + i + 8, f1);
            _mm256_storeu_ps(dst + i + 16, f2); _mm256_storeu_ps(dst + i + 24, f3);
        }
    }
    for (; i < n; i++) dst[i] = (float)src[i];
}

int main(void) {
    size_t n = (size_t)65536 * 404;
    uint8_t *src = aligned_alloc(64, n);
    float *dst = aligned_alloc(64, n * 4 + 64);
    for (size_t i = 0; i < n; i++) src[i] = (uint8_t)(i * 2654435761u >> 24);
    memset(dst, 0, n * 4);
    int maxt = omp_get_max_threads();
    printf("{\"max_threads\": %d, \"runs\": [", maxt);
    int first = 1;
    for (int nt_store = 0; nt_store < 2; nt_store++)
        for (int th = 1; th <= maxt; th = (th * 2 <= maxt || th == maxt) ? th * 2 : maxt) {
            double best = 1e9, sum = 0; int reps = 12;
            for (int r = 0; r < reps + 2; r++) {
                double t0 = now();
#pragma omp parallel num_threads(th)
                {
                    int k = omp_get_thread_num(), K = omp_get_num_threads();
                    size_t per = ((n / K) + 63) & ~(size_t)63, lo = per * k, hi = lo + per > n ? n : lo + per;
                    if (lo < n) widen(src + lo, dst + lo, hi - lo, nt_store);
                }
                double dt = now() - t0;
                if (r >= 2) { if (dt < best) best = dt; sum += dt; }
            }
            printf("%s{\"threads\": %d, \"nt_store\": %d, \"best_ms\": %.3f, \"mean_ms\": %.3f, \"out_GBps_best\": %.1f}",
                   first ? "" : ", ", th, nt_store, best * 1e3, sum / reps * 1e3, n * 4 / best / 1e9);
            first = 0;
            if (th == maxt) break;
        }
    double chk = 0; for (size_t i = 0; i < n; i += 4097) chk += dst[i];
    printf("], \"check\": %.0f}\n", chk);
    return 0;
}
