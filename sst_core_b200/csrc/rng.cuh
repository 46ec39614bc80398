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

// ---- MT19937 kept as WHOLE generations: the reference's own representation (mersenne.cc:56-68 generateNextBatch
// rebuilds all 624 words when the index wraps; a draw is temper(numbers[index++])).  A generator owns two 624-word
// buffers: buf[par] is the current generation, the other one receives the NEXT generation ahead of time -- as soon as
// the index passes MT_REGEN_AT -- so that a lookahead window whose draws run over the end of the current generation
// finds the following one in place, and so that a draw is ONE read and no write.  What element code keeps in its POD
// state is `pos` = index | par << 12.
constexpr uint32_t MT_REGEN_AT  = 424;  // 624 - 2 * (most pairs one window may draw in the event-parallel form)
constexpr uint32_t MT_POS_SHIFT = 12;

SSTB_HD uint32_t
mt_twist(uint32_t a, uint32_t b, uint32_t m)
{
    const uint32_t y = (a & 0x80000000u) + (b & 0x7fffffffu);
    uint32_t       v = m ^ (y >> 1);
    if ( y & 1u ) v ^= 2567483615u;
    return v;
}

// next[] = the generation after cur[], exactly as the reference's in-place loop produces it (word i reads the NEW
// word i-227 for i >= 227, and the last word reads the NEW word 0).  One thread; the warp-cooperative form is
// mt_warp_regen in parallel_form.cuh.
SSTB_HD void
mt_regen_serial(const uint32_t* cur, uint32_t* next)
{
    for ( int i = 0; i < MT_N; ++i ) {
        const uint32_t b = (i + 1 == MT_N) ? next[0] : cur[i + 1];
        const uint32_t m = (i < MT_N - MT_M) ? cur[i + MT_M] : next[i - (MT_N - MT_M)];
        next[i]          = mt_twist(cur[i], b, m);
    }
}

// ---- route digest (MessageMesh RouteMessage): the port choice of the pair of draws that starts at word 2j of a
// generation, temper(gen[2j]) % ng, 2 bits per pair, 16 pairs per word, 312 pairs = 20 words (the last half empty)
constexpr uint32_t MT_PAIRS             = MT_N / 2;
constexpr uint32_t ROUTE_DIGEST_WORDS   = 20;
constexpr uint32_t ROUTE_DIGEST_SPLIT   = 13; // words [0,13) = pairs [0,208) are refreshed when the next generation is
                                              // built, words [13,20) = pairs [208,312) when it becomes the current one
SSTB_HD uint32_t
route_digest_word(const uint32_t* gen, uint32_t wd, uint32_t ng)
{
    uint32_t out = 0;
    for ( uint32_t e = 0; e < 16; ++e ) {
        const uint32_t j = 16 * wd + e;
        if ( j < MT_PAIRS ) out |= (mt_temper(gen[2 * j]) % ng) << (2 * e);
    }
    return out;
}

// the digest entries of 16 consecutive pairs starting at pair h.  The entries form a bit stream: words 0..18, the low
// half of word 19, then the NEXT generation from word 0 on (in place once the index has passed MT_REGEN_AT).
SSTB_HD uint32_t
route_digest_window16(const uint32_t* dig, uint32_t h)
{
    const uint32_t wi = h >> 4;
    uint64_t       v;
    if ( wi == ROUTE_DIGEST_WORDS - 1 )
        v = (uint64_t)(dig[wi] & 0xFFFFu) | ((uint64_t)dig[0] << 16) | ((uint64_t)dig[1] << 48);
    else if ( wi == ROUTE_DIGEST_WORDS - 2 )
        v = (uint64_t)dig[wi] | ((uint64_t)(dig[wi + 1] & 0xFFFFu) << 32) | ((uint64_t)dig[0] << 48);
    else
        v = (uint64_t)dig[wi] | ((uint64_t)dig[wi + 1] << 32);
    return (uint32_t)(v >> ((h & 15u) * 2));
}
// entry of pair p, p in [0, 2 * 312): pairs past the end of the current generation belong to the next one
SSTB_HD uint32_t
route_digest_entry(const uint32_t* dig, uint32_t p)
{
    if ( p >= MT_PAIRS ) p -= MT_PAIRS;
    return (dig[p >> 4] >> ((p & 15u) * 2)) & 3u;
}
// n pairs drawn at once (n <= (624 - MT_REGEN_AT) / 2, even index): new position, and what has to be done before the
// next window: MT_OP_REGEN (the index passed MT_REGEN_AT: build the next generation and digest entries [0,208) of it)
// or MT_OP_TAIL (the next generation became the current one: digest entries [208,312) of it)
enum : uint32_t { MT_OP_NONE = 0, MT_OP_INIT = 1, MT_OP_REGEN = 2, MT_OP_TAIL = 3 };
constexpr uint32_t MT_MAX_PAIRS_PER_WINDOW = (MT_N - MT_REGEN_AT) / 2;
static_assert(MT_PAIRS + MT_MAX_PAIRS_PER_WINDOW - MT_PAIRS <= ROUTE_DIGEST_SPLIT * 16,
    "pairs read from the next generation in one window must lie in the digest words refreshed with it");
static_assert(MT_REGEN_AT / 2 >= ROUTE_DIGEST_SPLIT * 16, "digest entries refreshed at MT_REGEN_AT must be consumed");
SSTB_HD uint32_t
mt_pos_advance_pairs(uint32_t& pos, uint32_t n)
{
    uint32_t       idx = pos & ((1u << MT_POS_SHIFT) - 1), par = pos >> MT_POS_SHIFT, op = MT_OP_NONE;
    const uint32_t ni  = idx + 2 * n;
    if ( idx < MT_REGEN_AT && ni >= MT_REGEN_AT ) op = MT_OP_REGEN;
    idx = ni;
    if ( ni >= (uint32_t)MT_N ) { // cannot also have passed MT_REGEN_AT in this window: 2n <= 624 - MT_REGEN_AT
        idx = ni - MT_N;
        par ^= 1u;
        op = MT_OP_TAIL;
    }
    pos = idx | (par << MT_POS_SHIFT);
    return op;
}

// Sequential draws from the two-generation storage (the element's slow path; the event-parallel form reads the
// digest).  `digest` (may be null) is kept in step: it is what later windows in the event-parallel form will read.
// The two rare steps are kept out of line so that the handler code around a draw stays small.
#if defined(__CUDACC__)
#define SSTB_RARE __host__ __device__ __noinline__
#else
#define SSTB_RARE inline
#endif
SSTB_RARE void
mt_gen_build_next(uint32_t* buf, uint32_t par, uint32_t* digest, uint32_t ng)
{
    mt_regen_serial(buf + par * MT_N, buf + (par ^ 1u) * MT_N);
    if ( digest )
        for ( uint32_t wd = 0; wd < ROUTE_DIGEST_SPLIT; ++wd )
            digest[wd] = route_digest_word(buf + (par ^ 1u) * MT_N, wd, ng);
}
SSTB_RARE void
mt_gen_digest_tail(const uint32_t* gen, uint32_t* digest, uint32_t ng)
{
    for ( uint32_t wd = ROUTE_DIGEST_SPLIT; wd < ROUTE_DIGEST_WORDS; ++wd )
        digest[wd] = route_digest_word(gen, wd, ng);
}

struct MtGenDev
{
    uint32_t* buf;    // [2][624]
    uint32_t  pos;    // index | par << 12
    uint32_t* digest; // [ROUTE_DIGEST_WORDS] or null
    uint32_t  ng;
    SSTB_HD uint32_t u32()
    {
        uint32_t       idx = pos & ((1u << MT_POS_SHIFT) - 1), par = pos >> MT_POS_SHIFT;
        const uint32_t v   = buf[par * MT_N + idx];
        ++idx;
        if ( idx == MT_REGEN_AT ) mt_gen_build_next(buf, par, digest, ng);
        if ( idx == (uint32_t)MT_N ) {
            idx = 0;
            par ^= 1u;
            if ( digest ) mt_gen_digest_tail(buf + par * MT_N, digest, ng);
        }
        pos = idx | (par << MT_POS_SHIFT);
        return mt_temper(v);
    }
};

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

// x % d for 32-bit x and a divisor fixed at construction: M = 2^64 / d (rounded up) is kept with the component's
// constants and the remainder is the high half of (M * x mod 2^64) * d -- two multiplications instead of a division
// sequence (D. Lemire, "Faster remainder by direct computation", 2019; exact for all 32-bit x and d >= 1)
SSTB_HD uint64_t
fastmod_magic(uint32_t d)
{
    return 0xFFFFFFFFFFFFFFFFull / d + 1;
}
SSTB_HD uint32_t
fastmod_u32(uint32_t x, uint64_t M, uint32_t d)
{
    const uint64_t low = M * x;
#if defined(__CUDA_ARCH__)
    return (uint32_t)__umul64hi(low, d);
#else
    return (uint32_t)(((unsigned __int128)low * d) >> 64);
#endif
}

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
