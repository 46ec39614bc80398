// An element library from OUTSIDE the engine's tree, compiled in through csrc/elements.def
// (python -m sst_core_b200.build --with relay_elements.cuh:relay_elements.def --out libsstb200_relay.so).
// The reference loads such a library with dlopen and its SST_ELI_REGISTER_COMPONENT static initialisers
// (elemLoader.cc:159-177, eli/elementinfo.h:426-435); device code has to be part of the kernels, so here the library is
// a header with the element types + a list file, and its host half is registered from Python (relay.py).
//
// demo.relay: two ports.  setup(): a component with `start` != 0 sends a token (payload 1) out of port 0.  Every
// delivery counts the hop, adds the token's value to `sum`, and forwards the token + 1 out of the OTHER port.
#pragma once
#include "components.cuh"

namespace sstb200 {
#if defined(__CUDACC__)
struct RelayElement
{
    static constexpr uint32_t    TYPE            = 16;
    static constexpr const char* ELI_NAME        = "demo.relay";
    static constexpr const char* DESCRIPTION     = "example element from outside the tree: forwards a counting token";
    static constexpr uint32_t    AUX_BYTES       = 0;
    static constexpr bool        AUX_MT_SEED     = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool HAS_PAYLOAD = true;
    struct State
    {
        uint64_t hops, sum;
        int64_t  start, pad;
    };
    static constexpr uint32_t STORE_END = 16, CONST_END = 32, STATS_END = 32;
    static constexpr bool     TIES_INTERCHANGEABLE = false;
    static constexpr bool     STATS_EVERY_EVENT    = false;
    static constexpr bool     PARALLEL_EVENTS      = false;
    // no clock, one send per delivery, nothing but State / the event / ctx.send used: the ordered event-parallel
    // window kernel can run a rank of relays -- one declaration, the handler stays as it is
    static constexpr bool     ORDERED_EVENTS       = true;
    static constexpr uint32_t ORD_MAX_EVENTS       = 16;
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx&, State& s, const int64_t* p)
    {
        s.hops = s.sum = 0;
        s.start        = p[0];
        s.pad          = 0;
    }
    template <class Ctx>
    __device__ static void setup(Ctx& ctx, State& s)
    {
        if ( s.start ) ctx.send(0, 0, 1);
    }
    template <class Ctx>
    __device__ static void handleEvent(Ctx& ctx, State& s, int port, uint64_t& payload)
    {
        s.hops++;
        s.sum += payload;
        ctx.send(port ^ 1, 0, payload + 1);
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx&, State&, uint64_t)
    {
        return true;
    }
    __device__ static void initStats(State&) {}
    __device__ static void mergeStats(State&, const State&) {}
    __device__ static uint64_t statsVersion(const State&) { return 0; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = s.hops;
        r[1] = s.sum;
    }
};
#endif
} // namespace sstb200
