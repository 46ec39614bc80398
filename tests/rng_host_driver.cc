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
// MT19937 kept as whole generations (rng.cuh MtGenDev, what MeshElement uses): seed words in the second buffer, first
// generation built from them, then n sequential draws with the route digest maintained on the way
void
dev_mtgen_stream(uint32_t seed, uint64_t n, uint32_t* out)
{
    static uint32_t buf[2 * MT_N], dig[ROUTE_DIGEST_WORDS];
    mt_seed(buf + MT_N, seed);
    mt_regen_serial(buf + MT_N, buf);
    MtGenDev g { buf, 0, dig, 4 };
    for ( uint64_t i = 0; i < n; ++i )
        out[i] = g.u32();
}
// The event-parallel form of MessageMesh on the host: `lead` single draws (setup), then windows of n pairs whose
// ports are read from the route digest (first 16 from the window word, the rest entry by entry), the position
// advanced for all n at once and the maintenance operation it asks for applied after the window -- exactly what
// window_parallel_kernel + maintenance_kernel do.  Every 5th window is run by the sequential form instead (a block
// that fell back).  out[j] = port of the j-th pair; returns the number of pairs.
uint64_t
dev_mesh_routes(uint32_t seed, uint32_t lead, uint32_t ng, uint64_t max_pairs, const uint32_t* sizes, uint32_t n_sizes,
    uint32_t* out)
{
    static uint32_t buf[2 * MT_N], dig[ROUTE_DIGEST_WORDS];
    mt_seed(buf + MT_N, seed);
    mt_regen_serial(buf + MT_N, buf);
    for ( uint32_t wd = 0; wd < ROUTE_DIGEST_WORDS; ++wd )
        dig[wd] = route_digest_word(buf, wd, ng);
    MtGenDev g { buf, 0, dig, ng };
    for ( uint32_t i = 0; i < lead; ++i )
        g.u32();
    uint32_t pos = g.pos;
    uint64_t j   = 0;
    for ( uint32_t wsel = 0; j < max_pairs; ++wsel ) {
        uint32_t n = sizes[wsel % n_sizes];
        if ( j + n > max_pairs ) n = (uint32_t)(max_pairs - j);
        if ( wsel % 5 == 4 ) { // sequential form
            MtGenDev q { buf, pos, dig, ng };
            for ( uint32_t kq = 0; kq < n; ++kq ) {
                out[j++] = q.u32() % ng;
                q.u32();
            }
            pos = q.pos;
            continue;
        }
        const uint32_t h   = (pos & ((1u << MT_POS_SHIFT) - 1)) >> 1;
        const uint32_t hdr = route_digest_window16(dig, h);
        for ( uint32_t kq = 0; kq < n; ++kq )
            out[j++] = kq < 16 ? (hdr >> (2 * kq)) & 3u : route_digest_entry(dig, h + kq);
        const uint32_t op = mt_pos_advance_pairs(pos, n);
        if ( op == MT_OP_REGEN ) mt_gen_build_next(buf, pos >> MT_POS_SHIFT, dig, ng);
        if ( op == MT_OP_TAIL ) mt_gen_digest_tail(buf + (pos >> MT_POS_SHIFT) * MT_N, dig, ng);
    }
    return j;
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

extern "C" int
dev_fastmod_check(uint32_t d, const uint32_t* xs, uint64_t n)
{
    const uint64_t M = fastmod_magic(d);
    for ( uint64_t i = 0; i < n; ++i )
        if ( fastmod_u32(xs[i], M, d) != xs[i] % d ) return 0;
    return 1;
}
