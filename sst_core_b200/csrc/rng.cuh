// Device re-implementation of SST::RNG (reference: src/sst/core/rng/*).  Streams are bit-for-bit those of the
// reference classes; see tests/test_gpu_rng.py (vs oracle + tests/golden RNG vectors).
//
//   MersenneDev   rng/mersenne.cc:49-53,56-68,88-102,149-159  MT19937.  The reference regenerates all 624 words in
//                 a batch when index wraps; here the SAME recurrence is evaluated one word per draw, in place
//                 ("incremental twist"): word i of the next generation depends on old[i], old[i+1] and
//                 (i<227 ? old[i+397] : new[i-227]) -- exactly the values the in-order batch loop reads -- so the
//                 output stream is identical while a draw costs 3 word reads + 1 word write and never stalls a warp
//                 on a 624-step regeneration.  State = 624 words in HBM (per-component contiguous, 78 sectors) +
//                 a 32-bit index kept in the component's POD state.
//   MarsagliaDev  rng/marsaglia.cc:29-34,64-87,143-148
//   XorShiftDev   rng/xorshift.cc:44-50,58-77,131-140
//   i32/i64/u64   composed exactly as the reference does (lower half drawn first; u64 is the bit pattern of i64)
//   nextUniform   mersenne/xorshift: u32 / 4294967295.0 retried while >= 1.0; marsaglia: (u32+1 wrapping) * 2^-32
//                 constant 2.328306435454494e-10, retried while >= 1.0.  IEEE double division/multiplication are
//                 correctly rounded on the device, so these are bit-exact too.
//   distributions rng/discrete.h:96-109, uniform.h, poisson.h:73-85 (L = exp(-lambda) supplied by the host so that
//                 no device libm call is involved), expon.h:73-77 and gaussian.h:81-110 (device log/sqrt: agree with
//                 glibc within 1e-9 relative, not bit-exact -- SURVEY.md section 7 "libm parity").
#pragma once
#include <cstdint>
#if !defined(__CUDACC__)
struct uint2
{
    uint32_t x, y;
};
#else
#include <vector_types.h>
#endif

namespace sstb200 {

#if defined(__CUDACC__)
#define SSTB_HD __host__ __device__ __forceinline__
#else
#define SSTB_HD inline
#endif

constexpr int MT_N = 624;
constexpr int MT_M = 397;
// a window's draws may be staged ahead of time only while no pair reads a word an earlier pair of the same window
// wrote: pair k reads mt[i+2k+398] = the second word pair k-114 wrote.  Stay well below 114.
constexpr uint32_t MT_STAGE_MAX_PAIRS = 100;

SSTB_HD uint32_t
mt_temper(uint32_t t)
{
    t ^= t >> 11;
    t ^= (t << 7) & 2636928640u;
    t ^= (t << 15) & 4022730752u;
    t ^= t >> 18;
    return t;
}

// seeds 624 words (mersenne.cc:149-159); the index starts at 0 so the first draw twists word 0.
SSTB_HD void
mt_seed(uint32_t* mt, uint32_t seed)
{
    uint32_t prev = seed;
    mt[0]         = prev;
    for ( int i = 1; i < MT_N; i++ ) {
        prev  = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
        mt[i] = prev;
    }
}

struct MersenneDev
{
    uint32_t* mt;  // 624 words
    uint32_t  idx; // next word to twist+emit
    SSTB_HD uint32_t u32()
    {
        const uint32_t i  = idx;
        const uint32_t i1 = (i + 1 == MT_N) ? 0u : i + 1;
        const uint32_t im = (i + MT_M >= MT_N) ? i + MT_M - MT_N : i + MT_M;
        const uint32_t y  = (mt[i] & 0x80000000u) + (mt[i1] & 0x7fffffffu);
        uint32_t       v  = mt[im] ^ (y >> 1);
        if ( y & 1u ) v ^= 2567483615u;
        mt[i] = v;
        idx   = i1;
        return mt_temper(v);
    }
    // Two consecutive draws with all five word reads issued before either write: the second draw never depends on
    // the first one's store (word i+1 is still old, (i+1+397)%624 != i), so the loads overlap instead of forming two
    // dependent memory round trips.  Same stream as two u32() calls.
    SSTB_HD void u32x2(uint32_t& r0, uint32_t& r1)
    {
        const uint32_t i   = idx;
        const uint32_t i1  = (i + 1 == MT_N) ? 0u : i + 1;
        const uint32_t i2  = (i1 + 1 == MT_N) ? 0u : i1 + 1;
        const uint32_t im  = (i + MT_M >= MT_N) ? i + MT_M - MT_N : i + MT_M;
        const uint32_t im1 = (im + 1 == MT_N) ? 0u : im + 1;
        const uint32_t a0 = mt[i], a1 = mt[i1], a2 = mt[i2], b0 = mt[im], b1 = mt[im1];
        const uint32_t y0 = (a0 & 0x80000000u) + (a1 & 0x7fffffffu);
        uint32_t       v0 = b0 ^ (y0 >> 1);
        if ( y0 & 1u ) v0 ^= 2567483615u;
        // word i2 is new-generation only when the pair wraps (i1 == 623 -> i2 == 0, written 623 draws ago) and
        // word im1 == i only if 398 == 0 mod 624: never.  But im1 == i is impossible while im == i1 is too
        // (397 != 1 mod 624), so a2/b1 are exactly what a second u32() would have read -- except a2 when i2 == i
        // (MT_N == 2: impossible).
        const uint32_t y1 = (a1 & 0x80000000u) + (a2 & 0x7fffffffu);
        uint32_t       v1 = b1 ^ (y1 >> 1);
        if ( y1 & 1u ) v1 ^= 2567483615u;
        mt[i]  = v0;
        mt[i1] = v1;
        idx    = i2;
        r0     = mt_temper(v0);
        r1     = mt_temper(v1);
    }
    SSTB_HD double uniform()
    {
        double d;
        do {
            d = (double)u32() / 4294967295.0;
        } while ( d >= 1.0 );
        return d;
    }
};

// Register-resident read-ahead for a run of paired draws at EVEN index (624 is even, so parity never changes while a
// component only draws in pairs).  The words a pair needs are mt[i], mt[i+1], mt[i+2] and mt[i+397], mt[i+398]; with
// 8-byte-aligned chunks C_j = (mt[2j], mt[2j+1]) that is C_j + first word of C_j+1, and second word of the chunk
// holding mt[i+397] + first word of E_j = (mt[i+398], mt[i+399]).  The cache holds C_j, C_j+1, that second word and
// E_j, all loaded during EARLIER pairs: a pair is computed from registers only, then the two 8-byte loads for the
// pair after next are issued and are not waited for until they are needed -- one iteration of the handler loop later.
// Per pair: two 8-byte loads + one 8-byte store (instead of five 4-byte loads + two 4-byte stores), off the
// critical path.  Reading up to two pairs early is safe: pairs j and j+1 write only mt[i..i+3], and the words read
// early (mt[i+4], mt[i+5], mt[i+400], mt[i+401]) are never congruent to those mod 624.
struct MtPairCache
{
    uint32_t a0, a1, a2, a3; // C_j, C_j+1
    uint32_t b0, b1, b2;     // mt[i+397]; E_j = (mt[i+398], mt[i+399])
    uint32_t valid;
};

SSTB_HD uint32_t
mt_wrap(uint32_t i)
{
    return i >= MT_N ? i - MT_N : i;
}

// One pair at even index i from words already read: a0,a1,a2 = mt[i], mt[i+1], mt[i+2]; b0,b1 = mt[i+397], mt[i+398]
// (all mod 624).  Writes the two new-generation words back (one 8-byte store) and advances the index.
SSTB_HD void
mt_pair_from_words(uint32_t* mt, uint32_t& idx, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1,
    uint32_t& r0, uint32_t& r1)
{
    const uint32_t i  = idx;
    const uint32_t y0 = (a0 & 0x80000000u) + (a1 & 0x7fffffffu);
    uint32_t       v0 = b0 ^ (y0 >> 1);
    if ( y0 & 1u ) v0 ^= 2567483615u;
    const uint32_t y1 = (a1 & 0x80000000u) + (a2 & 0x7fffffffu);
    uint32_t       v1 = b1 ^ (y1 >> 1);
    if ( y1 & 1u ) v1 ^= 2567483615u;
    *reinterpret_cast<uint2*>(mt + i) = uint2 { v0, v1 };
    idx = mt_wrap(i + 2);
    r0  = mt_temper(v0);
    r1  = mt_temper(v1);
}

// loads the cache for a pair at even index i (C_j, C_j+1, mt[i+397], E_j)
SSTB_HD void
mt_cache_fill(const uint32_t* mt, uint32_t i, MtPairCache& c)
{
    const uint2 ca = *reinterpret_cast<const uint2*>(mt + i);
    const uint2 cb = *reinterpret_cast<const uint2*>(mt + mt_wrap(i + 2));
    const uint2 ce = *reinterpret_cast<const uint2*>(mt + mt_wrap(i + MT_M + 1));
    c.b0           = mt[mt_wrap(i + MT_M)];
    c.a0           = ca.x;
    c.a1           = ca.y;
    c.a2           = cb.x;
    c.a3           = cb.y;
    c.b1           = ce.x;
    c.b2           = ce.y;
    c.valid        = 1;
}

SSTB_HD void
mt_u32x2_cached(uint32_t* mt, uint32_t& idx, MtPairCache& c, uint32_t& r0, uint32_t& r1)
{
    const uint32_t i = idx;
    if ( (i & 1u) != 0 ) { // odd position: plain path
        c.valid = 0;
        MersenneDev g { mt, i };
        g.u32x2(r0, r1);
        idx = g.idx;
        return;
    }
    if ( !c.valid ) mt_cache_fill(mt, i, c);
    const uint32_t y0 = (c.a0 & 0x80000000u) + (c.a1 & 0x7fffffffu);
    uint32_t       v0 = c.b0 ^ (y0 >> 1);
    if ( y0 & 1u ) v0 ^= 2567483615u;
    const uint32_t y1 = (c.a1 & 0x80000000u) + (c.a2 & 0x7fffffffu);
    uint32_t       v1 = c.b1 ^ (y1 >> 1);
    if ( y1 & 1u ) v1 ^= 2567483615u;
    *reinterpret_cast<uint2*>(mt + i) = uint2 { v0, v1 };
    // read-ahead for the pair after next
    const uint2 na = *reinterpret_cast<const uint2*>(mt + mt_wrap(i + 4));        // C_j+2
    const uint2 ne = *reinterpret_cast<const uint2*>(mt + mt_wrap(i + MT_M + 3)); // E_j+1
    c.a0    = c.a2;
    c.a1    = c.a3;
    c.b0    = c.b2;
    c.a2    = na.x;
    c.a3    = na.y;
    c.b1    = ne.x;
    c.b2    = ne.y;
    c.valid = 1;
    idx     = mt_wrap(i + 2);
    r0      = mt_temper(v0);
    r1      = mt_temper(v1);
}

struct MarsagliaDev
{
    uint32_t z, w;
    SSTB_HD uint32_t u32()
    {
        z = 36969u * (z & 65535u) + (z >> 16);
        w = 18000u * (w & 65535u) + (w >> 16);
        return (z << 16) + w;
    }
    SSTB_HD double uniform()
    {
        double d;
        do {
            d = (double)(uint32_t)(u32() + 1u) * 2.328306435454494e-10;
        } while ( d >= 1.0 );
        return d;
    }
};

struct XorShiftDev
{
    uint32_t x, y, z, w;
    SSTB_HD void seed(uint32_t s)
    {
        x = s;
        y = 0;
        z = 0;
        w = 0;
    }
    SSTB_HD uint32_t u32()
    {
        uint32_t t = x ^ (x << 11);
        x          = y;
        y          = z;
        z          = w;
        w          = w ^ (w >> 19) ^ t ^ (t >> 8);
        return w;
    }
    SSTB_HD double uniform()
    {
        double d;
        do {
            d = (double)u32() / 4294967295.0;
        } while ( d >= 1.0 );
        return d;
    }
};

// Composite draws shared by all generators (Random interface, rng/rng.h:29-65).
template <class G>
SSTB_HD int32_t
rng_i32(G& g)
{
    return (int32_t)g.u32();
}
template <class G>
SSTB_HD int64_t
rng_i64(G& g)
{
    uint32_t lo = g.u32();
    uint32_t hi = g.u32();
    return (int64_t)((uint64_t)lo | ((uint64_t)hi << 32));
}
template <class G>
SSTB_HD uint64_t
rng_u64(G& g)
{
    return (uint64_t)rng_i64(g);
}

// ---- distributions (RandomDistribution::getNextDouble, rng/distrib.h:23-46)
// discrete.h:40-55,96-109: prefix[] holds EXCLUSIVE prefix sums; result = first index with prefix[index] >= u
template <class G>
SSTB_HD double
dist_discrete(G& g, const double* prefix, uint32_t count)
{
    const double u     = g.uniform();
    uint32_t     index = 0;
    for ( ; index < count; index++ )
        if ( prefix[index] >= u ) break;
    return (double)index;
}
// uniform.h getNextDouble
template <class G>
SSTB_HD double
dist_uniform_bins(G& g, uint32_t bins)
{
    const double probPerBin = 1.0 / (double)bins;
    const double u          = g.uniform();
    uint32_t     bin        = 1;
    while ( u > ((double)bin * probPerBin) )
        bin++;
    return (double)(bin - 1);
}
// poisson.h:73-85 with L = exp(-lambda) computed on the host
template <class G>
SSTB_HD double
dist_poisson(G& g, double L)
{
    double p = 1.0;
    int    k = 0;
    do {
        k++;
        p *= g.uniform();
    } while ( p > L );
    return (double)(k - 1);
}
#if defined(__CUDACC__)
// expon.h:73-77
template <class G>
__device__ __forceinline__ double
dist_exponential(G& g, double lambda)
{
    const double u = g.uniform();
    return log(1 - u) / (-1 * lambda);
}
// gaussian.h:81-110 (polar method; second value cached)
struct GaussianDev
{
    double mean, stddev, unused;
    bool   usePair;
    template <class G>
    __device__ __forceinline__ double next(G& g)
    {
        if ( usePair ) {
            usePair = false;
            return unused;
        }
        double gu, gv, sq;
        do {
            gu = g.uniform();
            gv = g.uniform();
            sq = (gu * gu) + (gv * gv);
        } while ( sq >= 1 || sq == 0 );
        if ( g.uniform() < 0.5 ) gu *= -1.0;
        if ( g.uniform() < 0.5 ) gv *= -1.0;
        double mult = sqrt(-2.0 * log(sq) / sq);
        unused      = mean + stddev * gv * mult;
        usePair     = true;
        return mean + stddev * gu * mult;
    }
};
#endif

} // namespace sstb200
