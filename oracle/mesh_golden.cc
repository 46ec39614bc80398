// TEST INFRASTRUCTURE -- count-only golden model of the MessageMesh at sizes the event-by-event oracle
// (pdes_oracle.cc) and the reference cannot hold.  Only tests/, tests/golden/make_bench_digests.py and bench.py's checker
// leg may load it; the product never does.
//
// SURVEY.md section 8c ("restatement already validated"): per component std::mt19937(id + 100)
// (enclosingComponent.cc:164-170 MersenneRNG(my_id + 100); the stream equals std::mt19937's); setup(): one draw per
// port, and with mod = 1 one event to each of the 4 neighbours (enclosingComponent.cc:188-199); every 1 ns step a
// component that received n events adds n to its count and, n times, draws p = u32 % 4, draws and discards one more
// u32, and forwards the event to neighbour p (enclosingComponent.cc:179-186; ports 0: x+1, 1: x-1, 2: y+1, 3: y-1 with
// torus wrap, tests/test_MessageMesh.py:26-121).  Mesh events carry no payload and all events of a step share one
// timestamp, so per-component results do not depend on the order inside a step.
//
// Memory: the route decisions of a component's first PAIRS pairs of draws are kept as 2 bits each, not the 2.5 KB
// generator.  PAIRS is sized from the number of windows (4 events per component and window on average, plus 10 standard
// deviations and a margin); a component that needs more makes the function fail rather than wrap.
//
// Pinned by tests/test_oracle.py: equal to pdes_oracle.cc (itself pinned to the reference's golden files) on meshes
// up to 256 x 256, and to tests/refFiles/test_MessageMesh.out's run through the 6 x 6 fixture for its first windows.
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace {
uint32_t
pairs_for(uint32_t windows)
{
    const double mean = 4.0 * windows;
    uint32_t     p    = (uint32_t)(mean + 10.0 * __builtin_sqrt(mean) + 64.0);
    return p < 310 ? 310 : p;
}

template <class F>
void
parallel_for(uint64_t n, unsigned threads, F&& body)
{
    if ( threads <= 1 || n < 4096 ) {
        body((uint64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    for ( unsigned t = 0; t < threads; ++t )
        th.emplace_back([&, t] { body(n * t / threads, n * (t + 1) / threads); });
    for ( auto& x : th )
        x.join();
}
} // namespace

extern "C" {

// counts_out[x*y]: events received by every component in windows 1 .. windows-1 (window 0 has nothing due; the
// engine's "W windows" run delivers at times 1 ns .. (W-1) ns).  Returns the number of events delivered, or
// (uint64_t)-1 when a component ran out of first-generation pairs.
uint64_t
mesh_golden_counts(uint32_t X, uint32_t Y, uint32_t windows, unsigned threads, int32_t* counts_out)
{
    const uint64_t n = (uint64_t)X * Y;
    const uint32_t PAIRS = pairs_for(windows), WORDS = (PAIRS + 15) / 16;
    std::vector<uint32_t> route(n * WORDS);
    parallel_for(n, threads, [&](uint64_t lo, uint64_t hi) {
        for ( uint64_t c = lo; c < hi; ++c ) {
            std::mt19937 g((uint32_t)(c + 100));
            for ( int i = 0; i < 4; ++i )
                (void)g(); // setup(): one draw per port link; mod = 1 sends on every one of them
            uint32_t* r = &route[c * WORDS];
            for ( uint32_t j = 0; j < PAIRS; ++j ) {
                const uint32_t p = g() % 4;
                (void)g();
                r[j >> 4] |= p << (2 * (j & 15));
            }
        }
    });
    std::vector<uint16_t> used(n, 0);
    std::vector<uint16_t> in(n, 4), nxt(n); // after setup every component has 4 events in flight to it
    std::memset(counts_out, 0, n * sizeof(int32_t));
    uint64_t             total = 0;
    volatile char        fail  = 0;
    std::vector<uint8_t> out4(n * 4);
    for ( uint32_t w = 1; w < windows; ++w ) {
        // gather form (no write conflicts between threads): what component c receives next = the events its four
        // neighbours forward towards it this step.  First every component turns its n events into per-port counts.
        std::memset(out4.data(), 0, out4.size());
        parallel_for(n, threads, [&](uint64_t lo, uint64_t hi) {
            for ( uint64_t c = lo; c < hi; ++c ) {
                const uint32_t k = in[c];
                if ( !k ) continue;
                counts_out[c] += (int32_t)k;
                uint32_t        u = used[c];
                if ( u + k > PAIRS ) {
                    fail = 1;
                    continue;
                }
                const uint32_t* r = &route[c * WORDS];
                for ( uint32_t e = 0; e < k; ++e, ++u )
                    out4[c * 4 + ((r[u >> 4] >> (2 * (u & 15))) & 3u)]++;
                used[c] = (uint16_t)u;
            }
        });
        if ( fail ) return ~0ull;
        parallel_for(n, threads, [&](uint64_t lo, uint64_t hi) {
            for ( uint64_t c = lo; c < hi; ++c ) {
                const uint64_t mx = c % X, my = c / X;
                const uint64_t xp = ((mx + 1) % X) + my * X, xn = ((mx + X - 1) % X) + my * X;
                const uint64_t yp = mx + ((my + 1) % Y) * X, yn = mx + ((my + Y - 1) % Y) * X;
                // c receives what x-1 sent on its port 0 (x+), x+1 on port 1, y-1 on port 2, y+1 on port 3
                nxt[c] = (uint16_t)(out4[xn * 4 + 0] + out4[xp * 4 + 1] + out4[yn * 4 + 2] + out4[yp * 4 + 3]);
            }
        });
        for ( uint64_t c = 0; c < n; ++c )
            total += in[c];
        in.swap(nxt);
    }
    return total;
}
}
