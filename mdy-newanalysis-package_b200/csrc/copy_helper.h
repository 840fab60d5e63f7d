// copy_helper.h - host copies into the pinned upload ring (no CUDA in here; tests/stress_copy_helper.cpp runs it alone).
//
// copy_to_ring: streaming stores.  The only reader of a ring slot is the DMA engine, so the lines need not be fetched
// for ownership nor kept in the caches: one third less host-memory traffic per frame, and the caller's arrays stay
// cached.  Measured with one rank per GPU on the 16-core box (cfg3, 100 MB of frames per 64-frame step and rank):
// end to end 10,064 -> 11,494 x 10^9 pair evaluations/s at 8 GPUs.
//
// CopyHelper: an optional second host thread that takes the second half of every large frame.  A frame of BASELINE cfg3
// is 1.6 MB, which one core of the box moves in ~0.3 ms - about what the GPU needs for the frame - so the host copy, not
// the kernel, bounds the end-to-end rate of one rank.  The helper spins for a short while after a job (frames arrive
// back to back) and sleeps on a condition variable otherwise.
#pragma once

#include <emmintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace gfn {

inline void copy_to_ring(void* dst_, const void* src_, size_t n)
{
    unsigned char* dst = static_cast<unsigned char*>(dst_);
    const unsigned char* src = static_cast<const unsigned char*>(src_);
    if (n < 4096) { memcpy(dst, src, n); return; }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
    size_t k = 0;
    for (; k + 64 <= n; k += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k + 32));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k + 48), d);
    }
    if (k < n) memcpy(dst + k, src + k, n - k);
    _mm_sfence();
}

class CopyHelper {
public:
    struct Seg { unsigned char* d; const unsigned char* s; size_t n; };
    static constexpr int kMaxSegs = 8;
    static constexpr size_t kMinSplitBytes = 512u << 10;

    CopyHelper() : th_(&CopyHelper::run, this) {}
    ~CopyHelper()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_.store(true);
        }
        cv_.notify_all();
        th_.join();
    }
    CopyHelper(const CopyHelper&) = delete;
    CopyHelper& operator=(const CopyHelper&) = delete;

    // Copies the segments of one frame; returns when ALL bytes are in place.  The caller keeps bytes [0, half), the
    // helper gets [half, total), cut at a 64-byte multiple inside a segment.
    void copy(const Seg* segs, int nseg)
    {
        size_t total = 0;
        for (int k = 0; k < nseg; ++k) total += segs[k].n;
        const size_t half = total / 2;
        Seg mine[kMaxSegs];
        int nm = 0;
        nseg_ = 0;
        size_t at = 0;
        for (int k = 0; k < nseg; ++k) {
            const size_t n = segs[k].n;
            if (at + n <= half) mine[nm++] = segs[k];
            else if (at >= half) segs_[nseg_++] = segs[k];
            else {
                const size_t cut = ((half - at) + 63) & ~(size_t)63;
                const size_t m = cut < n ? cut : n;
                mine[nm++] = Seg{segs[k].d, segs[k].s, m};
                if (m < n) segs_[nseg_++] = Seg{segs[k].d + m, segs[k].s + m, n - m};
            }
            at += n;
        }
        const unsigned ticket = posted_.load(std::memory_order_relaxed) + 1;
        // seq_cst on purpose: this store and the load of sleeping_ below form a Dekker pair with the helper's
        // "sleeping_ = true; re-check posted_".  With a release store the load may pass it (x86 store buffer): the
        // caller sees "not sleeping", the helper sees "nothing posted" and sleeps for good.
        posted_.store(ticket, std::memory_order_seq_cst);
        if (sleeping_.load(std::memory_order_seq_cst)) {
            std::lock_guard<std::mutex> lk(mu_);
            cv_.notify_one();
        }
        for (int k = 0; k < nm; ++k) copy_to_ring(mine[k].d, mine[k].s, mine[k].n);
        // the helper started at the same time on the same amount of bytes: normally done by now.  If it was not
        // scheduled (oversubscribed box), give the core away instead of spinning against it.
        for (int spins = 0; done_.load(std::memory_order_acquire) != ticket; ++spins) {
            if (spins < 2000) _mm_pause();
            else std::this_thread::yield();
        }
    }

private:
    void run()
    {
        unsigned seen = 0;
        for (;;) {
            int spins = 0;
            while (posted_.load(std::memory_order_acquire) == seen) {
                if (stop_.load(std::memory_order_relaxed)) return;
                if (++spins > 4000) {
                    std::unique_lock<std::mutex> lk(mu_);
                    sleeping_.store(true, std::memory_order_seq_cst);
                    cv_.wait(lk, [&] { return stop_.load() || posted_.load(std::memory_order_seq_cst) != seen; });
                    sleeping_.store(false, std::memory_order_seq_cst);
                    spins = 0;
                } else {
                    _mm_pause();
                }
            }
            seen = posted_.load(std::memory_order_acquire);
            for (int k = 0; k < nseg_; ++k) copy_to_ring(segs_[k].d, segs_[k].s, segs_[k].n);
            done_.store(seen, std::memory_order_release);
        }
    }

    std::mutex mu_;
    std::condition_variable cv_;
    Seg segs_[kMaxSegs];
    int nseg_ = 0;
    std::atomic<unsigned> posted_{0}, done_{0};
    std::atomic<bool> sleeping_{false}, stop_{false};
    std::thread th_;      // last member: starts when everything above is constructed
};


// CopyPool: worker threads for BLOCK calls (gfn_enqueue_frames with many frames, above all one process feeding several
// devices): the frames that go into one batch are copied into their ring slots in parallel, one frame per task.  One
// host core moves a cfg3 frame (1.6 MB) in ~0.2-0.3 ms, a B200 consumes one in 0.33 ms: one process that feeds 8 devices
// from pageable arrays needs several copying cores.  run(n, fn) returns when fn(0..n-1) have all finished; the caller
// works too.
class CopyPool {
public:
    explicit CopyPool(int nthreads)
    {
        for (int t = 0; t < nthreads; ++t) th_.emplace_back(&CopyPool::worker, this);
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread& t : th_) t.join();
    }
    CopyPool(const CopyPool&) = delete;
    CopyPool& operator=(const CopyPool&) = delete;
    int threads() const { return (int)th_.size(); }

    void run(int n, const std::function<void(int)>& fn)
    {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_.store(0); left_.store(n); ++gen_;
        }
        cv_.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return left_.load() == 0 && busy_ == 0; });
        fn_ = nullptr;
    }

private:
    void drain()
    {
        for (;;) {
            const int i = next_.fetch_add(1);
            if (i >= n_) return;
            (*fn_)(i);
            left_.fetch_sub(1);
        }
    }
    void worker()
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                ++busy_;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu_);
                --busy_;
            }
            done_cv_.notify_all();
        }
    }

    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, busy_ = 0;
    unsigned long long gen_ = 0;
    std::atomic<int> next_{0}, left_{0};
    bool stop_ = false;
    std::vector<std::thread> th_;
};

}  // namespace gfn
