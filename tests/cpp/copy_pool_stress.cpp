#include "copy_pool.hpp"
#include <cstdio>
#include <cstdlib>
#include <atomic>
int main() {
    aero::host::CopyPool &pool = *new aero::host::CopyPool(7);  // never destroyed (its threads sleep on its condition variable)
    const size_t N = 64u << 20;
    std::vector<uint8_t> a(N), b(N);
    for (int it = 0; it < 50; it++) {
        for (size_t i = 0; i < N; i += 4099) a[i] = (uint8_t)(i * 31 + it);
        std::vector<aero::host::CopyPool::Chunk> ch;
        const size_t chunk = (it % 3 == 0) ? (1u << 20) : (it % 3 == 1 ? 333333 : (8u << 20));
        for (size_t x = 0; x < N; x += chunk) ch.push_back({b.data() + x, a.data() + x, std::min(chunk, N - x)});
        pool.run(std::move(ch));
        if (memcmp(a.data(), b.data(), N)) { printf("MISMATCH at iteration %d\n", it); return 1; }
    }
    // concurrent callers
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++) th.emplace_back([&, t] {
        std::vector<uint8_t> x(8u << 20, (uint8_t)t), y(8u << 20);
        for (int it = 0; it < 20; it++) {
            std::vector<aero::host::CopyPool::Chunk> ch;
            for (size_t o = 0; o < x.size(); o += (1u << 20)) ch.push_back({y.data() + o, x.data() + o, (size_t)(1u << 20)});
            pool.run(std::move(ch));
            if (memcmp(x.data(), y.data(), x.size())) bad++;
        }
    });
    for (auto &t : th) t.join();
    printf(bad ? "CONCURRENT MISMATCH\n" : "copy pool ok\n");
    return bad ? 1 : 0;
}
