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

// Split frame of PSK_FEATURES_F32_WIRE_U8 (include/psk_craft.h): how many of the next call's `chunks`
// chunks should cross PCIe as f32, given that the call just made sent `d` that way, its last byte landed
// t_pcie_us and its widening finished t_widen_us after it started.  One byte chunk costs p on the wire
// (an f32 chunk 4 p) and w on the host threads; both sides finish together at
// d* = chunks (w - p) / (w + 3 p); the rule takes floor(0.8 d*) (DMA writes and the threads' stores share
// the host's DRAM).  *p_us / *w_us carry the smoothed rates from call to call (0 = none yet).  Half a
// chunk of hysteresis on the way down.  Pure arithmetic: exported as psk_debug_wire_split_next for the CPU test.
int psk_wire_split_next(int chunks, int d, double t_pcie_us, double t_widen_us, double *p_us, double *w_us);

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
