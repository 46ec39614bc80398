// Host compile of the DEVICE RNG header (sst_core_b200/csrc/rng.cuh is host/device code): lets the CPU-only test
// tier check the incremental MT19937 twist and the paired draw against the oracle without a GPU.
#include "../sst_core_b200/csrc/rng.cuh"

#include <cstdint>
using namespace sstb200;

extern "C" {
// mode 0: u32() only; mode 1: u32x2() pairs; mode 2: `lead` single draws then pairs (exercises every alignment of
// the pair against the 624-word wrap)
void
dev_mt_stream(uint32_t seed, int mode, uint32_t lead, uint64_t n, uint32_t* out)
{
    static uint32_t mt[MT_N];
    mt_seed(mt, seed);
    MersenneDev g { mt, 0 };
    uint64_t    i = 0;
    if ( mode == 2 )
        for ( ; i < lead && i < n; ++i )
            out[i] = g.u32();
    if ( mode == 3 ) { // `lead` single draws, then cached pairs; the cache is dropped every `lead+3` pairs (window end)
        for ( ; i < lead && i < n; ++i )
            out[i] = g.u32();
        MtPairCache c {};
        uint32_t    idx = g.idx, pairs = 0;
        for ( ; i + 1 < n; i += 2 ) {
            mt_u32x2_cached(mt, idx, c, out[i], out[i + 1]);
            if ( ++pairs % (lead + 3) == 0 ) c.valid = 0;
        }
        g.idx = idx;
        for ( ; i < n; ++i )
            out[i] = g.u32();
        return;
    }
    if ( mode == 4 ) { // windows of `lead`+1, 1, 7, 100 ... pairs, every window staged ahead like MeshElement::beginWindow
        uint32_t       idx     = 0, wsel = 0;
        const uint32_t sizes[] = { lead + 1, 1, 7, MT_STAGE_MAX_PAIRS, 3, 113 };
        static uint32_t scr[4 * 128 + 8];
        while ( i + 1 < n ) {
            uint32_t np = sizes[wsel++ % 6];
            if ( np > MT_STAGE_MAX_PAIRS ) np = MT_STAGE_MAX_PAIRS;
            for ( uint32_t c = 0; c <= np; ++c ) {
                scr[2 * c]     = mt[mt_wrap(mt_wrap(idx + 2 * c))];
                scr[2 * c + 1] = mt[mt_wrap(mt_wrap(idx + 2 * c)) + 1];
            }
            uint32_t* scb = scr + 2 * np + 2;
            for ( uint32_t j = 0; j <= 2 * np; ++j )
                scb[j] = mt[mt_wrap(mt_wrap(idx + MT_M + j))];
            for ( uint32_t kq = 0; kq < np && i + 1 < n; ++kq, i += 2 )
                mt_pair_from_words(mt, idx, scr[2 * kq], scr[2 * kq + 1], scr[2 * kq + 2], scb[2 * kq], scb[2 * kq + 1],
                    out[i], out[i + 1]);
        }
        g.idx = idx;
        for ( ; i < n; ++i )
            out[i] = g.u32();
        return;
    }
    if ( mode == 0 ) {
        for ( ; i < n; ++i )
            out[i] = g.u32();
    }
    else {
        for ( ; i + 1 < n; i += 2 )
            g.u32x2(out[i], out[i + 1]);
        for ( ; i < n; ++i )
            out[i] = g.u32();
    }
}
void
dev_marsaglia_stream(uint32_t z, uint32_t w, uint64_t n, uint32_t* out, double* uni)
{
    MarsagliaDev g { z, w };
    for ( uint64_t i = 0; i < n; ++i )
        out[i] = g.u32();
    MarsagliaDev h { z, w };
    for ( uint64_t i = 0; i < n; ++i )
        uni[i] = h.uniform();
}
void
dev_xorshift_stream(uint32_t seed, uint64_t n, uint32_t* out)
{
    XorShiftDev g;
    g.seed(seed);
    for ( uint64_t i = 0; i < n; ++i )
        out[i] = g.u32();
}
}
