// Host-side half of the "bytes on the wire" feature format of psk_craft_host_tick_resident
// (PSK_FEATURES_F32_WIRE_U8): the device writes the u8 frame (every feature is an exact integer
// <= 255), PCIe carries a quarter of the bytes, and a small pool of host threads widens each chunk
// to the caller's f32 buffer while the next chunk is still in flight.  The frame the caller sees is
// the f32 frame of state.features() (reference worlds/craft.py:142-181), bit for bit.
//
// Measured on the GPU box (profiles/probes/host_widen_probe.c, 16 cores): 8 threads with
// non-temporal stores widen a 65,536 x 404 frame in 0.54 ms (196 GB/s of f32 written) against
// 1.9 ms for the same f32 frame over PCIe.
#include "psk_hostcpu.h"

#include <immintrin.h>

#include <cstdlib>
#include <cstring>

namespace {

__attribute__((target("avx2"))) void widen_avx2(const uint8_t *src, float *dst, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) {      // align the stores
        dst[i] = static_cast<float>(src[i]);
        i++;
    }
    for (; i + 32 <= n; i += 32) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
        _mm256_stream_ps(dst + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(a)));
        _mm256_stream_ps(dst + i + 8, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(a, 8))));
        _mm256_stream_ps(dst + i + 16, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(b)));
        _mm256_stream_ps(dst + i + 24, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(b, 8))));
    }
    for (; i < n; i++) dst[i] = static_cast<float>(src[i]);
    _mm_sfence();
}

void widen_scalar(const uint8_t *src, float *dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = static_cast<float>(src[i]);
}

}  // namespace

void psk_widen_u8_f32(const uint8_t *src, float *dst, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) widen_avx2(src, dst, n);
    else widen_scalar(src, dst, n);
}

// ---------------------------------------------------------------------------------------------------

PskWidenPool::PskWidenPool(int n_threads, size_t capacity) : ring_(capacity) {
    for (int i = 0; i < n_threads; i++) {
        try {
            workers_.emplace_back([this] { worker(); });
        } catch (...) {         // thread limit reached: go on with the workers that started (the
            break;              // calling thread widens too, so zero workers still works)
        }
    }
}

PskWidenPool::~PskWidenPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_.store(true);
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
}

int PskWidenPool::default_threads() {
    if (const char *e = getenv("PSK_HOST_THREADS")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 256) return v;
    }
    const unsigned hw = std::thread::hardware_concurrency();
    // half the cores, at most 8: DRAM write bandwidth saturates there (probe: 8 threads 196 GB/s,
    // 16 threads 146 GB/s), and the caller's own threads keep the other half
    int n = static_cast<int>(hw / 2);
    if (n > 8) n = 8;
    if (n < 1) n = 1;
    return n - 1;               // the calling thread widens too
}

bool PskWidenPool::run_one() {
    uint64_t h = head_.load(std::memory_order_relaxed);
    while (h < tail_.load(std::memory_order_acquire)) {
        if (head_.compare_exchange_weak(h, h + 1, std::memory_order_acq_rel)) {
            const Block b = ring_[h % ring_.size()];
            psk_widen_u8_f32(b.src, b.dst, b.n);
            done_.fetch_add(1, std::memory_order_release);
            return true;
        }
    }
    return false;
}

void PskWidenPool::worker() {
    for (;;) {
        if (run_one()) continue;
        bool found = false;
        for (int spin = 0; spin < 20000 && !found; spin++) {        // ~100-200 us: one tick's gap
            _mm_pause();
            found = head_.load(std::memory_order_relaxed) < tail_.load(std::memory_order_acquire);
            if (stop_.load(std::memory_order_relaxed)) return;
        }
        if (found) continue;
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_.fetch_add(1);
        cv_.wait(lk, [this] {
            return stop_.load() || head_.load() < tail_.load();
        });
        sleepers_.fetch_sub(1);
        if (stop_.load()) return;
    }
}

void PskWidenPool::submit(const uint8_t *src, float *dst, size_t n, size_t block) {
    // the caller waits (finish) before more than `capacity` blocks are outstanding: a tick submits
    // ceil(frame / block) + one per chunk, and the ring is sized for that by the host context
    for (size_t off = 0; off < n; off += block) {
        const uint64_t t = tail_.load(std::memory_order_relaxed);
        while (t - done_.load(std::memory_order_acquire) >= ring_.size()) run_one();
        ring_[t % ring_.size()] = Block{src + off, dst + off, n - off < block ? n - off : block};
        tail_.store(t + 1, std::memory_order_release);
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);     // tail_ visible before sleepers_ is read
    if (sleepers_.load()) {
        { std::lock_guard<std::mutex> lk(mu_); }
        cv_.notify_all();
    }
}

void PskWidenPool::help() {
    if (!run_one()) _mm_pause();
}

void PskWidenPool::finish() {
    while (run_one()) {}
    const uint64_t t = tail_.load(std::memory_order_relaxed);
    while (done_.load(std::memory_order_acquire) < t) _mm_pause();
}

int psk_wire_split_next(int chunks, int d, double t_pcie_us, double t_widen_us, double *p_us, double *w_us) {
    if (d < 0) d = 0;
    if (chunks < 2 || d >= chunks || t_pcie_us <= 0.0 || t_widen_us <= 0.0)
        return d < chunks ? d : (chunks > 0 ? chunks - 1 : 0);          // nothing to balance
    const double p = t_pcie_us / (chunks + 3.0 * d), w = t_widen_us / (chunks - d);
    *p_us = *p_us > 0 ? 0.5 * (*p_us + p) : p;
    *w_us = *w_us > 0 ? 0.5 * (*w_us + w) : w;
    const double ps = *p_us, ws = *w_us;
    // 0.8: an f32 chunk costs the host more than its wire time — its DMA writes compete with the threads'
    // stores for the same DRAM (measured: where the plain balance said d = 1 the step was 4 % slower than
    // with d = 0).  Rounded down, and half a chunk of hysteresis on the way back.
    const double best = ws > ps ? 0.8 * chunks * (ws - ps) / (ws + 3.0 * ps) : 0.0;
    int nd = static_cast<int>(best);
    if (nd > chunks - 1) nd = chunks - 1;
    if (nd > d) return nd;
    if (best < d - 0.5) return nd;
    return d;
}

extern "C" int psk_debug_wire_split_next(int chunks, int d, double t_pcie_us, double t_widen_us, double *p_us,
                                         double *w_us) {
    if (!p_us || !w_us) return -1;
    return psk_wire_split_next(chunks, d, t_pcie_us, t_widen_us, p_us, w_us);
}

// ---------------------------------------------------------------------------------------------------
// C ABI (include/psk_craft.h): the widening on its own, for callers that take PSK_FEATURES_U8 frames
// and want f32 later.  No CUDA call: usable (and tested) on a box without a GPU.
extern "C" int psk_host_widen_u8_f32(const uint8_t *src, float *dst, size_t n, int threads) {
    if ((!src || !dst) && n) return 2;              // PSK_ERR_BADARG
    if (threads < 0 || threads > 256) return 2;
    if (threads <= 1 || n < (1u << 20)) {
        psk_widen_u8_f32(src, dst, n);
        return 0;
    }
    const size_t block = 64u << 10;
    PskWidenPool pool(threads - 1, n / block + 2);
    pool.submit(src, dst, n, block);
    pool.finish();
    return 0;
}
