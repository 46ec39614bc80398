// TEST INFRASTRUCTURE.  Streams of the reference's OWN generator classes (src/sst/core/rng/mersenne.cc, marsaglia.cc,
// xorshift.cc), linked from the reference objects oracle/build_ref.py compiled: position-weighted checksums
// sum_i (i+1) * draw_i (mod 2^64) per block of 2^20 draws, plus the first and last draw.  tests/golden/make_golden.py
// turns its output into fixtures that pin the oracle's restatement (and through it the device generators) on streams
// far longer than the 100 000-tick golden files of the reference.
//   rng_ref_driver <kind 0 mersenne|1 marsaglia|2 xorshift> <s1> <s2> <n> [u64]
#include "sst_config.h"

#include "sst/core/rng/marsaglia.h"
#include "sst/core/rng/mersenne.h"
#include "sst/core/rng/xorshift.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>

int
main(int argc, char** argv)
{
    if ( argc < 5 ) return 2;
    const int      kind = atoi(argv[1]);
    const uint64_t s1 = strtoull(argv[2], nullptr, 10), s2 = strtoull(argv[3], nullptr, 10), n = strtoull(argv[4], nullptr, 10);
    const bool     u64  = argc > 5;
    SST::RNG::Random* g = nullptr;
    if ( kind == 0 )
        g = new SST::RNG::MersenneRNG((unsigned int)s1);
    else if ( kind == 1 )
        g = new SST::RNG::MarsagliaRNG((unsigned int)s1, (unsigned int)s2); // (z, w)
    else
        g = new SST::RNG::XORShiftRNG((unsigned int)s1);
    uint64_t total = 0, block = 0, first = 0, last = 0;
    for ( uint64_t i = 0; i < n; ++i ) {
        const uint64_t v = u64 ? g->generateNextUInt64() : (uint64_t)g->generateNextUInt32();
        if ( i == 0 ) first = v;
        last = v;
        total += (i + 1) * v;
        block += (i + 1) * v;
        if ( ((i + 1) & ((1u << 20) - 1)) == 0 ) {
            printf("block %llu %llu\n", (unsigned long long)(i >> 20), (unsigned long long)block);
            block = 0;
        }
    }
    printf("total %llu first %llu last %llu\n", (unsigned long long)total, (unsigned long long)first, (unsigned long long)last);
    delete g;
    return 0;
}
