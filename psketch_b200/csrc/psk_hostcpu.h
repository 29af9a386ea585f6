// Internal (not part of the C ABI): host threads that widen u8 feature frames to f32.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

// dst[i] = (float)src[i]; AVX2 + non-temporal stores when the CPU has them.
void psk_widen_u8_f32(const uint8_t *src, float *dst, size_t n);

// One producer (the thread inside psk_craft_host_tick_resident), n_threads consumers + the producer
// itself in finish().  Counters only grow; the ring holds at most `capacity` outstanding blocks.
class PskWidenPool {
  public:
    PskWidenPool(int n_threads, size_t capacity);
    ~PskWidenPool();
    static int default_threads();       // env PSK_HOST_THREADS, else min(8, cores / 2) - 1
    void submit(const uint8_t *src, float *dst, size_t n, size_t block);
    void finish();                      // help, then wait until everything submitted is written
    // the same in pieces, for a caller that watches something else while it helps
    bool idle() const { return done_.load(std::memory_order_acquire) == tail_.load(std::memory_order_relaxed); }
    void help();                        // widen one block if one is waiting, else pause
    int threads() const { return static_cast<int>(workers_.size()); }

  private:
    struct Block {
        const uint8_t *src;
        float *dst;
        size_t n;
    };
    bool run_one();
    void worker();
    std::vector<Block> ring_;
    std::vector<std::thread> workers_;
    std::atomic<uint64_t> head_{0}, tail_{0}, done_{0};
    std::atomic<int> sleepers_{0};
    std::atomic<bool> stop_{false};
    std::mutex mu_;
    std::condition_variable cv_;
};
