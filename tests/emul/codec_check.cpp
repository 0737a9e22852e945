// Test-only: the "fast" R11G11B10 codecs of image_view.h against the contract's codecs - every code (decode) and every binary32 value (encode).
#include <cstdio>
#include <cstdint>
#include <thread>
#include <vector>
#include <atomic>
#include "image_view.h"
using namespace pv;
int main() {
    long bad = 0;
    for (int mbits : {5, 6})
        for (uint32_t v = 0; v < (32u << mbits); v++) {
            float a = pv::decodeSmallFloat(v, mbits), b = pv::decodeSmallFloatFast(v, mbits);
            if (dm::f2u(a) != dm::f2u(b)) { if (bad < 10) printf("decode %d %u: %08x %08x\n", mbits, v, dm::f2u(a), dm::f2u(b)); bad++; }
        }
    printf("decode mismatches: %ld\n", bad);
    std::atomic<long> ebad{0};
    std::vector<std::thread> th;
    const int T = 16;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (uint64_t u = (uint64_t)t; u < (1ull << 32); u += T) {
                float f = dm::u2f((uint32_t)u);
                for (int mbits : {5, 6})
                    if (pv::encodeSmallFloat(f, mbits) != pv::encodeSmallFloatFast(f, mbits)) {
                        if (ebad++ < 10) printf("encode %d %08x: %u %u\n", mbits, (uint32_t)u, pv::encodeSmallFloat(f, mbits), pv::encodeSmallFloatFast(f, mbits));
                    }
            }
        });
    for (auto& x : th) x.join();
    printf("encode mismatches over all 2^32 binary32 values x 2 formats: %ld\n", (long)ebad);
    return (bad || ebad) ? 1 : 0;
}
