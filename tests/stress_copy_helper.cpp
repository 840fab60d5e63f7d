// Stress test of csrc/copy_helper.h (no GPU): several independent caller/helper pairs, more threads than cores, copies
// of random size and alignment, pauses that let the helpers fall asleep between jobs.  Every copy is verified.
// usage: stress_copy_helper <pairs> <iterations> <sleep_every>    exit code 0 = all copies correct (a stuck pair never returns:
// run it under a timeout).  With the caller's store of the ticket relaxed to memory_order_release the 16-pair run hangs
// within seconds on an 8-core host (lost wake-up), which is how the 8-rank bench hang was found.
#include "../mdy-newanalysis-package_b200/csrc/copy_helper.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <thread>
#include <vector>

// "pool" mode: stress_copy_helper pool <pools> <rounds> - several CopyPool instances (more threads than cores), rounds of
// parallel frame copies with 1 .. 40 tasks each, pauses in between so that the workers sleep on their condition variable;
// every task must run exactly once per round and every copy is verified.
static int pool_mode(int pools, int rounds)
{
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int p = 0; p < pools; ++p) th.emplace_back([&, p] {
        std::mt19937_64 rng(99 + p);
        const size_t frame = 200u << 10;
        const int maxn = 40;
        std::vector<unsigned char> src(frame * maxn), dst(frame * maxn);
        for (size_t k = 0; k < src.size(); ++k) src[k] = (unsigned char)(rng() >> 56);
        gfn::CopyPool pool(1 + (int)(rng() % 5));
        std::vector<std::atomic<int>> hits(maxn);
        for (int r = 0; r < rounds; ++r) {
            const int n = 1 + (int)(rng() % maxn);
            for (auto& h : hits) h.store(0);
            memset(dst.data(), 0, frame * n);
            pool.run(n, [&](int i) {
                hits[i].fetch_add(1);
                gfn::copy_to_ring(dst.data() + frame * i, src.data() + frame * i, frame - (size_t)(i % 3) * 17);
            });
            for (int i = 0; i < n; ++i) {
                if (hits[i].load() != 1 || memcmp(dst.data() + frame * i, src.data() + frame * i, frame - (size_t)(i % 3) * 17) != 0) { bad.fetch_add(1); return; }
            }
            if (r % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(100 + rng() % 400));
        }
    });
    for (auto& t : th) t.join();
    printf("pools %d rounds %d bad %d\n", pools, rounds, bad.load());
    return bad.load() ? 1 : 0;
}

int main(int argc, char** argv)
{
    if (argc > 1 && std::string(argv[1]) == "pool") return pool_mode(argc > 2 ? atoi(argv[2]) : 4, argc > 3 ? atoi(argv[3]) : 2000);
    const int pairs = argc > 1 ? atoi(argv[1]) : 8;
    const int iters = argc > 2 ? atoi(argv[2]) : 20000;
    const int sleep_every = argc > 3 ? atoi(argv[3]) : 3;
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int p = 0; p < pairs; ++p) th.emplace_back([&, p] {
        std::mt19937_64 rng(1234 + p);
        const size_t cap = 3u << 20;
        std::vector<unsigned char> src(cap + 64), dst(cap + 64);
        for (size_t k = 0; k < src.size(); ++k) src[k] = (unsigned char)(rng() >> 56);
        gfn::CopyHelper helper;
        for (int it = 0; it < iters; ++it) {
            const int nseg = 1 + (int)(rng() % 4);
            gfn::CopyHelper::Seg segs[4];
            size_t off = rng() % 16;
            const size_t soff = rng() % 16;
            size_t total = 0;
            for (int k = 0; k < nseg; ++k) {
                const size_t n = (it % 7 == 0 ? 64 : 4096) + rng() % (cap / 4 - 4096);
                segs[k] = {dst.data() + off, src.data() + soff + off, n};
                off += n; total += n;
            }
            memset(dst.data(), 0, off + 16);
            helper.copy(segs, nseg);
            if (memcmp(dst.data() + (segs[0].d - dst.data()), segs[0].s, total) != 0) { bad.fetch_add(1); break; }
            if (it % sleep_every == 0) std::this_thread::sleep_for(std::chrono::microseconds(150 + rng() % 300));   // helper goes to sleep
        }
    });
    for (auto& t : th) t.join();
    printf("pairs %d iterations %d bad %d\n", pairs, iters, bad.load());
    return bad.load() ? 1 : 0;
}
