// Host check of csrc/sincos_core.cuh against the C library: every one of the 2^32 float arguments (or a band of them)
// must give the bits of sinf / cosf. Build and run: see scripts/verify_sincos_all_floats.sh.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <thread>
#include <vector>

#include "../gnuradio4_b200/csrc/sincos_core.cuh"

int main(int argc, char** argv) {
    const unsigned long long first = argc > 1 ? std::strtoull(argv[1], nullptr, 0) : 0ull;
    const unsigned long long last  = argc > 2 ? std::strtoull(argv[2], nullptr, 0) : (1ull << 32);
    const unsigned           nThreads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<unsigned long long> mismatches{0}, nanMismatches{0};
    std::vector<std::thread>        pool;
    for (unsigned t = 0; t < nThreads; ++t) {
        pool.emplace_back([&, t] {
            unsigned long long bad = 0, badNan = 0;
            for (unsigned long long bits = first + t; bits < last; bits += nThreads) {
                const unsigned u = static_cast<unsigned>(bits);
                float          y;
                std::memcpy(&y, &u, 4);
                volatile float yv = y; // keep the compiler from folding the library calls
                const float    sl = sinf(yv), cl = cosf(yv);
                float          s, c;
                gr4b200::sinCosGlibc(y, &s, &c);
                if (((u >> 20) & 0x7ffu) < 0x42fu) { // |y| < 120: the library's own operation sequence must agree as well
                    float sr, cr;
                    gr4b200::sinCosGlibcReference(y, &sr, &cr);
                    if (std::memcmp(&s, &sr, 4) != 0 || std::memcmp(&c, &cr, 4) != 0) {
                        ++bad;
                        continue;
                    }
                }
                if (std::isnan(sl) || std::isnan(cl)) {
                    badNan += !(std::isnan(s) && std::isnan(c));
                    continue;
                }
                if (std::memcmp(&s, &sl, 4) != 0 || std::memcmp(&c, &cl, 4) != 0) {
                    if (bad < 4 && t == 0) {
                        std::printf("mismatch at %a: sin %a vs %a, cos %a vs %a\n", y, s, sl, c, cl);
                    }
                    ++bad;
                }
            }
            mismatches += bad;
            nanMismatches += badNan;
        });
    }
    for (auto& th : pool) {
        th.join();
    }
    std::printf("{\"first\": %llu, \"last\": %llu, \"mismatches\": %llu, \"nan_mismatches\": %llu, \"threads\": %u}\n", first, last, mismatches.load(), nanMismatches.load(), nThreads);
    return mismatches.load() == 0 && nanMismatches.load() == 0 ? 0 : 1;
}
