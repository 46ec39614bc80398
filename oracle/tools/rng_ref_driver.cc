// TEST INFRASTRUCTURE.  Streams of the reference's OWN generator classes (src/sst/core/rng/mersenne.cc, marsaglia.cc,
// xorshift.cc), linked from the reference objects oracle/build_ref.py compiled: position-weighted checksums
// sum_i (i+1) * draw_i (mod 2^64) per block of 2^20 draws, plus the first and last draw.  tests/golden/make_golden.py
// turns its output into fixtures that pin the oracle's restatement (and through it the device generators) on streams
// far longer than the 100 000-tick golden files of the reference.
//   rng_ref_driver <kind 0 mersenne|1 marsaglia|2 xorshift> <s1> <s2> <n> [u64]
//   rng_ref_driver dist <d 0 discrete|1 expon|2 gaussian|3 poisson|4 uniform|5 constant> <kind> <s1> <s2> <n> <params...>
//       the reference's distribution classes (rng/discrete.h, expon.h, gaussian.h, poisson.h, uniform.h, constant.h)
//       over one of its generators: n values, printed as the bit patterns of the doubles
#include "sst_config.h"

#include "sst/core/rng/constant.h"
#include "sst/core/rng/discrete.h"
#include "sst/core/rng/expon.h"
#include "sst/core/rng/gaussian.h"
#include "sst/core/rng/marsaglia.h"
#include "sst/core/rng/mersenne.h"
#include "sst/core/rng/poisson.h"
#include "sst/core/rng/uniform.h"
#include "sst/core/rng/xorshift.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static SST::RNG::Random*
make(int kind, uint64_t s1, uint64_t s2)
{
    if ( kind == 0 ) return new SST::RNG::MersenneRNG((unsigned int)s1);
    if ( kind == 1 ) return new SST::RNG::MarsagliaRNG((unsigned int)s1, (unsigned int)s2); // (z, w)
    return new SST::RNG::XORShiftRNG((unsigned int)s1);
}

static int
distributions(int argc, char** argv)
{
    if ( argc < 7 ) return 2;
    const int      d = atoi(argv[2]), kind = atoi(argv[3]);
    const uint64_t s1 = strtoull(argv[4], nullptr, 10), s2 = strtoull(argv[5], nullptr, 10), n = strtoull(argv[6], nullptr, 10);
    std::vector<double> p;
    for ( int i = 7; i < argc; ++i )
        p.push_back(atof(argv[i]));
    SST::RNG::Random*             g = make(kind, s1, s2); // owned (and deleted) by the distribution
    SST::RNG::RandomDistribution* dist = nullptr;
    switch ( d ) {
    case 0: dist = new SST::RNG::DiscreteDistribution(p.data(), (uint32_t)p.size(), g); break;
    case 1: dist = new SST::RNG::ExponentialDistribution(p[0], g); break;
    case 2: dist = new SST::RNG::GaussianDistribution(p[0], p[1], g); break;
    case 3: dist = new SST::RNG::PoissonDistribution(p[0], g); break;
    case 4: dist = new SST::RNG::UniformDistribution((uint32_t)p[0], g); break;
    default: dist = new SST::RNG::ConstantDistribution(p[0]); delete g; break;
    }
    for ( uint64_t i = 0; i < n; ++i ) {
        const double v = dist->getNextDouble();
        uint64_t     bits;
        memcpy(&bits, &v, 8);
        printf("%llu\n", (unsigned long long)bits);
    }
    delete dist;
    return 0;
}

int
main(int argc, char** argv)
{
    if ( argc > 1 && strcmp(argv[1], "dist") == 0 ) return distributions(argc, argv);
    if ( argc < 5 ) return 2;
    const int      kind = atoi(argv[1]);
    const uint64_t s1 = strtoull(argv[2], nullptr, 10), s2 = strtoull(argv[3], nullptr, 10), n = strtoull(argv[4], nullptr, 10);
    const bool     u64  = argc > 5;
    SST::RNG::Random* g = make(kind, s1, s2);
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
