// GPU-eligible element types: POD state + device handler functors, registered in a compile-time table.
//
// This is the device half of the reference's component plugin API:
//   SST_ELI_REGISTER_COMPONENT(cls, "lib", "name", ...)   component.h:93-95  ->  SSTB200_ELI_DEVICE_COMPONENT below:
//                                                         the type appears in DeviceElements with its ELI name
//   Component(ComponentId_t, Params&)                     ->  static construct(Ctx&, State&, const int64_t params[8])
//                                                         (the host half, sst_core_b200/elements.py, has already
//                                                         parsed Params and fixed the configureLink order)
//   setup()                                               ->  static setup(Ctx&, State&)
//   Event::Handler<cls,&cls::fn,int>(this, port)          ->  static handleEvent(Ctx&, State&, int port, uint64_t& payload)
//   Clock::Handler<cls,&cls::fn>  (bool = unregister)     ->  static clockTick(Ctx&, State&, uint64_t cycle)
//   Link::send(delay, ev) / send(ev)   link.h:202-220     ->  ctx.send(port, delay, payload)
//   Statistic<T>::addData              statbase.h:393-401 ->  Accumulator<T>::add   (stataccumulator.h:89-95)
//   primaryComponentOKToEndSim         baseComponent.cc:869-920 -> ctx.primaryComponentOKToEndSim()
//   getCurrentSimCycle                                    ->  ctx.now
//   (none: engine-side latency hiding)                    ->  static prefetch(state*, aux*): issue L2 prefetches for what
//                                                         the handlers will read this window
//   finish()                                              ->  static results(const State&, uint64_t r[32]); the host
//                                                         half formats the console line / statistic rows from r.
//   several clocks / handlers per component, ordered handlers, activate() / deactivate() / reregisterClock()
//   (clock.h:60-260, clock.cc:25-155,193-250; baseComponent.h:488-620)
//                                                         ->  MULTI_CLOCK = true + a ClockSet<> in the State:
//                                                         static clockFire(Ctx&, State&, time) runs every handler due at
//                                                         `time`, static nextClock(State&) is the next such time
//   configureSelfLink (baseComponent.cc:374-378)          ->  SELF_LINKS = true; the self-link ports are the LAST ports of
//                                                         the component, ctx.send() on them may arrive inside the window
//   getSimulationOutput().output(fmt, ...)                ->  ctx.output(code, a, b): one record of the engine's output
//                                                         log; the host half owns the format strings
//   statistics with an output rate / startat / stopat / resetOnOutput (statengine.cc:107-204,346-505)
//                                                         ->  STAT_CLOCK = true + EngineStat<T> members: static
//                                                         nextStatTime(State&) / statFire(Ctx&, State&, time) run the
//                                                         statistics engine's activities of the component (priority 85:
//                                                         after the clocks and events of the same time)
//
// A handler runs on exactly one thread and is never concurrent with another handler of the same component -- the
// reference's threading contract.  Components touch no memory but their own State, their aux
// block and what ctx hands them.
#pragma once
#include "rng.cuh"

#if defined(__CUDACC__)
#include <cuda_pipeline.h>
#endif

#include <cstdint>
#include <limits>

namespace sstb200 {

enum : uint32_t { TYPE_MESH = 1, TYPE_PHOLD = 2, TYPE_CLOCKNODE = 3, TYPE_RNGCOMP = 4, TYPE_CORETEST = 5, TYPE_CLOCKER = 6,
                  TYPE_STATS_INT = 7, TYPE_STATS_FLOAT = 8,
                  TYPE_MAX = 32 }; // type codes are bits of 32-bit masks; 16..31 are free for elements compiled in from outside
constexpr int NPARAM  = 8;
constexpr int NRESULT = 32;

// statapi/stataccumulator.h:44-209 -- fields Sum, SumSQ, Count, Min, Max
template <typename T>
struct Accumulator
{
    T        sum, sumsq, mn, mx;
    uint64_t count;
    SSTB_HD void init()
    {
        sum   = 0;
        sumsq = 0;
        mn    = std::numeric_limits<T>::max();
        mx    = std::numeric_limits<T>::min();
        count = 0;
    }
    SSTB_HD void add(T v)
    {
        sum += v;
        sumsq += v * v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        count++;
    }
    // addData_impl_Ntimes (stataccumulator.h:97-103) + the collection count
    SSTB_HD void addN(T v, uint64_t n)
    {
        sum += (T)n * v;
        sumsq += (T)n * v * v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        count += n;
    }
    // fold the data another accumulator collected into this one (same result as presenting its values here)
    SSTB_HD void merge(const Accumulator& o)
    {
        sum += o.sum;
        sumsq += o.sumsq;
        mn = o.mn < mn ? o.mn : mn;
        mx = o.mx > mx ? o.mx : mx;
        count += o.count;
    }
    SSTB_HD void out(uint64_t* r) const
    {
        r[0] = (uint64_t)(int64_t)sum;
        r[1] = (uint64_t)(int64_t)sumsq;
        r[2] = count;
        r[3] = count ? (uint64_t)(int64_t)mn : 0;
        r[4] = count ? (uint64_t)(int64_t)mx : 0;
    }
};

// oracle/probe/pdes_elements.cc mixDelivery (order-sensitive per-component delivery hash)
SSTB_HD uint64_t
mixDelivery(uint64_t h, uint64_t time, uint64_t port, uint64_t payload)
{
    uint64_t v = time * 0x9E3779B97F4A7C15ull + (port + 1) * 0xC2B2AE3D27D4EB4Full + payload;
    h ^= v;
    h *= 0x100000001B3ull;
    h ^= h >> 29;
    return h;
}

// addData(1) on one of four accumulators chosen at run time, written as a static selection so that the accumulators
// (and with them the element's whole State) stay in registers instead of a dynamically indexed local-memory array
template <typename T>
SSTB_HD void
add_one_of4(Accumulator<T>* a, uint32_t which)
{
    if ( which == 0 )
        a[0].add(1);
    else if ( which == 1 )
        a[1].add(1);
    else if ( which == 2 )
        a[2].add(1);
    else
        a[3].add(1);
}

// (x % f) == 0 for a signed draw x and a positive divisor f fixed at construction (the C++ remainder is zero exactly
// when |x| is a multiple of f)
SSTB_HD bool
divisible_i32(int32_t x, uint64_t magic, int32_t f)
{
    const uint32_t a = x < 0 ? 0u - (uint32_t)x : (uint32_t)x;
    return fastmod_u32(a, magic, (uint32_t)f) == 0;
}

// ------------------------------------------------------------------------------------------------------------------
// The clocks of ONE component (clock.cc / clock.h restated per component).  In the reference a Clock object per period
// is shared by all components of a partition (simulation.cc:1403-1427); what a component can observe of it is:
//   * its own unordered handlers are called in the order they were (re)registered / activated (Clock::registerHandler
//     appends, clock.cc:193-231; removal keeps the order of the others, :233-250, :324-367);
//   * its ordered handlers -- its Clock::Group on that clock (component.cc:34-45) -- are called after them, in
//     registration order (Group::activateHandler re-inserts by order_, clock.cc:48-88; Clock::execute runs the groups
//     after the handler list, :369-390);
//   * a handler that returns true is removed; the cycle number passed is time / period;
//   * a handler (re)activated at time t from an event handler first fires at the next multiple of the period after t
//     (Clock::schedule / getNextCycle, clock.cc:305-311,400-420: events run at priority 50, after the clocks of t).
// Between Clock objects of different periods due at the same time the reference's order is the TimeVortex insertion
// order -- the longer period was re-inserted earlier and runs first in steady state.  That is the order used here; a
// handler activated by a clock handler of a longer period therefore still fires at the same time when that time is a
// multiple of its own period.  (The reference leaves this case to global state no component controls, which is why
// coreTestClockerComponent routes its operations through a self link, coreTest_ClockerComponent.h:154-160.)
// Handlers are numbered 0..NH-1 in registration order; classes (periods) 0..NC-1.
template <int NC, int NH>
struct ClockSet
{
    uint64_t period[NC];
    uint64_t next[NC];     // next time class c fires, meaningful while it has an active handler
    uint8_t  list[NC][NH]; // active unordered handlers of the class, call order
    uint8_t  n[NC];
    uint8_t  cls[NH], ordered[NH], active[NH];
    uint8_t  nh, nclass, phase, pad0; // phase: class being executed by fire(), NC outside
    static constexpr uint64_t NEVER = ~0ull;

    SSTB_HD void init()
    {
        for ( int c = 0; c < NC; ++c ) {
            period[c] = 0;
            next[c]   = NEVER;
            n[c]      = 0;
            for ( int h = 0; h < NH; ++h )
                list[c][h] = 0;
        }
        for ( int h = 0; h < NH; ++h )
            cls[h] = ordered[h] = active[h] = 0;
        nh = nclass = 0;
        phase       = NC;
        pad0        = 0;
    }
    SSTB_HD int classFor(uint64_t per)
    {
        for ( int c = 0; c < (int)nclass; ++c )
            if ( period[c] == per ) return c;
        period[nclass] = per;
        return (int)nclass++;
    }
    SSTB_HD bool anyActive(int c) const
    {
        if ( n[c] ) return true;
        for ( int h = 0; h < (int)nh; ++h )
            if ( cls[h] == c && ordered[h] && active[h] ) return true;
        return false;
    }
    // registerClock / registerOrderedClock at time `now` (0 in a constructor): returns the handler id
    SSTB_HD int add(uint64_t per, bool is_ordered, uint64_t now)
    {
        const int h = (int)nh++;
        cls[h]      = (uint8_t)classFor(per);
        ordered[h]  = is_ordered ? 1 : 0;
        active[h]   = 0;
        activate(h, now);
        return h;
    }
    // HandlerBase::activate / reregisterClock: returns the cycle at which the clock fires next (getNextCycle)
    SSTB_HD uint64_t activate(int h, uint64_t now)
    {
        const int      c     = cls[h];
        const uint64_t per   = period[c];
        // has this time's execute of the handler's clock still to come? (only inside fire(), for a shorter period)
        const bool     later = phase < NC && per < period[phase] && now != 0 && now % per == 0;
        if ( !active[h] ) {
            const bool was_empty = !anyActive(c);
            active[h]            = 1;
            if ( !ordered[h] ) list[c][n[c]++] = (uint8_t)h;
            if ( was_empty ) next[c] = later ? now : (now / per + 1) * per;
        }
        return later ? now / per : now / per + 1;
    }
    // HandlerBase::deactivate / unregisterClock
    SSTB_HD void deactivate(int h)
    {
        if ( !active[h] ) return;
        active[h]   = 0;
        const int c = cls[h];
        if ( ordered[h] ) return;
        int at = -1;
        for ( int j = 0; j < (int)n[c]; ++j )
            if ( list[c][j] == h ) at = j;
        if ( at < 0 ) return;
        for ( int j = at; j + 1 < (int)n[c]; ++j )
            list[c][j] = list[c][j + 1];
        n[c]--;
    }
    SSTB_HD uint64_t nextTime() const
    {
        uint64_t t = NEVER;
        for ( int c = 0; c < (int)nclass; ++c )
            if ( next[c] < t && anyActive(c) ) t = next[c];
        return t;
    }
    // Clock::execute of every clock of this component due at time t; call(h, cycle) -> true = remove the handler.
    // Returns the number of handler calls.
    template <class F>
    SSTB_HD uint32_t fire(uint64_t t, F&& call)
    {
        uint32_t calls = 0, done = 0;
        for ( int pass = 0; pass < (int)nclass; ++pass ) {
            int c = -1;
            for ( int k = 0; k < (int)nclass; ++k )
                if ( !(done >> k & 1u) && next[k] == t && anyActive(k) && (c < 0 || period[k] > period[c]) ) c = k;
            if ( c < 0 ) break;
            done |= 1u << c;
            phase                = (uint8_t)c;
            const uint64_t cycle = t / period[c];
            for ( int i = 0; i < (int)n[c]; ) {
                const int  h    = list[c][i];
                const bool stop = call(h, cycle);
                calls++;
                int at = -1; // the callee may have changed the list
                for ( int j = 0; j < (int)n[c]; ++j )
                    if ( list[c][j] == h ) at = j;
                if ( at < 0 ) continue;
                if ( stop ) {
                    deactivate(h);
                    i = at;
                }
                else
                    i = at + 1;
            }
            for ( int h = 0; h < (int)nh; ++h )
                if ( cls[h] == c && ordered[h] && active[h] ) {
                    if ( call(h, cycle) ) active[h] = 0;
                    calls++;
                }
            next[c] = t + period[c];
        }
        phase = NC;
        return calls;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// A statistic under the statistics engine (statapi/statbase.h:393-401 addData; statbase.cc:84-142 collection counts
// and the event-based output trigger; statengine.cc:107-204 "rate" / "startat" / "stopat" / "resetOnOutput", :346-376
// and :468-505 periodic output and what follows an output), restated per statistic.  Configuration: six int64 from
// the component's extra parameters (sst_core_b200/elements.py _stat_slot_params): flags (1 exists, 2 resetOnOutput,
// 4 event-based), rate (cycles or events), startat, stopat, index of the object the slot adds to, unused.
// An output is one record of the output log: code = slot | end_of_simulation << 8, values Sum, SumSQ, Count, Min, Max
// as raw 64-bit patterns (Min / Max 0 while Count is 0, stataccumulator.h:171-194).
enum : uint32_t { ST_EXISTS = 1, ST_CLEAR = 2, ST_EVENT = 4 };
#if defined(__CUDACC__)
template <typename T>
__device__ __forceinline__ uint64_t
stat_bits(T v)
{
    return (uint64_t)(int64_t)v;
}
template <>
__device__ __forceinline__ uint64_t
stat_bits<float>(float v)
{
    return (uint64_t)__float_as_uint(v);
}
template <>
__device__ __forceinline__ uint64_t
stat_bits<double>(double v)
{
    return (uint64_t)__double_as_longlong(v);
}
// sum += v, sumsq += v * v exactly as the host compiler evaluates them: separate, correctly rounded operations (never
// a fused multiply-add)
template <typename T>
__device__ __forceinline__ void
stat_accumulate(T& sum, T& sumsq, T v)
{
    sum += v;
    sumsq += v * v;
}
template <>
__device__ __forceinline__ void
stat_accumulate<float>(float& sum, float& sumsq, float v)
{
    sum   = __fadd_rn(sum, v);
    sumsq = __fadd_rn(sumsq, __fmul_rn(v, v));
}
template <>
__device__ __forceinline__ void
stat_accumulate<double>(double& sum, double& sumsq, double v)
{
    sum   = __dadd_rn(sum, v);
    sumsq = __dadd_rn(sumsq, __dmul_rn(v, v));
}

struct EngineStatCtl
{
    uint64_t rate, start, stop; // configuration
    uint64_t next_out;          // next periodic output, ~0 = none
    uint64_t count, out_count;  // current_collection_count_, output_collection_count_
    uint32_t flags;
    uint8_t  registered, enabled, start_pending, stop_pending;
    static constexpr uint64_t NEVER = ~0ull;
    __device__ void configure(const int64_t* x)
    {
        flags = x ? (uint32_t)x[0] : 0;
        rate  = x ? (uint64_t)x[1] : 0;
        start = x ? (uint64_t)x[2] : 0;
        stop  = x ? (uint64_t)x[3] : 0;
        next_out = NEVER;
        count = out_count = 0;
        registered = start_pending = stop_pending = 0;
        enabled                                   = 1;
    }
    // registerStatisticWithEngine at time `now` (0 in a constructor)
    __device__ void registerNow(uint64_t now)
    {
        if ( !(flags & ST_EXISTS) || registered ) return;
        registered = 1;
        if ( !(flags & ST_EVENT) && rate > 0 ) next_out = (now / rate + 1) * rate; // the rate's statistic clock
        if ( start > 0 && start > now ) { // setStatisticStartTime: disabled until then
            enabled       = 0;
            start_pending = 1;
        }
        if ( stop > 0 && stop > now ) stop_pending = 1;
    }
    __device__ uint64_t nextTime() const
    {
        uint64_t t = next_out;
        if ( start_pending && start < t ) t = start;
        if ( stop_pending && stop < t ) t = stop;
        return t;
    }
};
template <typename T>
struct EngineStat
{
    EngineStatCtl c;
    T             sum, sumsq, mn, mx;
    __device__ void clearData()
    {
        sum   = 0;
        sumsq = 0;
        mn    = std::numeric_limits<T>::max();
        mx    = std::numeric_limits<T>::min();
    }
    __device__ void configure(const int64_t* x)
    {
        c.configure(x);
        clearData();
    }
    __device__ void fields(uint64_t* v) const
    {
        v[0] = stat_bits(sum);
        v[1] = stat_bits(sumsq);
        v[2] = c.count;
        v[3] = c.count ? stat_bits(mn) : 0;
        v[4] = c.count ? stat_bits(mx) : 0;
    }
    template <class Ctx>
    __device__ void output(Ctx& ctx, uint32_t slot)
    {
        uint64_t v[5];
        fields(v);
        ctx.output(slot, v[0], v[1], v[2], v[3], v[4]);
        // performStatisticOutputImpl: event-based statistics restart their count, resetOnOutput clears the data
        if ( c.flags & ST_EVENT ) c.out_count = 0;
        if ( c.flags & ST_CLEAR ) {
            clearData();
            c.count = c.out_count = 0;
        }
    }
    template <class Ctx>
    __device__ void addData(Ctx& ctx, uint32_t slot, T v)
    {
        if ( !c.registered || !c.enabled ) return; // NullStatistic / disabled statistic
        stat_accumulate(sum, sumsq, v);
        mn = (v < mn) ? v : mn;
        mx = (v > mx) ? v : mx;
        c.count++;
        c.out_count++;
        if ( (c.flags & ST_EVENT) && c.rate >= 1 && c.out_count >= c.rate ) output(ctx, slot);
    }
    // the statistics engine's activities of this statistic at time t
    template <class Ctx>
    __device__ void fire(Ctx& ctx, uint32_t slot, uint64_t t)
    {
        if ( c.start_pending && c.start == t ) {
            c.enabled       = 1;
            c.start_pending = 0;
        }
        if ( c.stop_pending && c.stop == t ) {
            c.enabled      = 0;
            c.stop_pending = 0;
        }
        if ( c.next_out == t ) {
            output(ctx, slot);
            c.next_out = t + c.rate;
        }
    }
};
#endif

// events a component sent to itself over a self link that are due inside the current window: a small queue in the
// owning thread, ordered by (time, port, sequence)
struct SelfQueue
{
    static constexpr int CAP = 8;
    uint32_t n;
    uint32_t port[CAP], seq[CAP];
    uint64_t time[CAP], pay[CAP];
};

#if defined(__CUDACC__)

__device__ __forceinline__ void
prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ------------------------------------------------------------------------------------------------------------------
// coreTestElement.message_mesh.enclosing_component (+ message_port / port_slot / route_message subcomponents,
// flattened).  Reference: testElements/message_mesh/enclosingComponent.cc:29-69 (ctor), :72-81 (setup), :92-99
// (handleEvent), :164-199 (RouteMessage).
//
// Component record (RECORD_BYTES = 112; a block's 256 records are moved to shared memory by one bulk copy):
//   [0,16)    mutable core: message_count, MT19937 position (index | buffer << 12), msg_count statistic
//   [16,32)   construction constants
//   [32,112)  route digest of the generator's current generation (rng.cuh): the port RouteMessage::send picks for the
//             pair of draws starting at word 2j, 2 bits per pair.  The handler of an event is then a table look-up.
// aux block: two 624-word MT19937 generations (rng.cuh MtGenDev).
enum : uint32_t { MAINT_INIT = MT_OP_INIT, MAINT_REGEN = MT_OP_REGEN, MAINT_TAIL = MT_OP_TAIL };

struct MeshElement
{
    static constexpr uint32_t    TYPE         = TYPE_MESH;
    static constexpr const char* ELI_NAME     = "coreTestElement.message_mesh.enclosing_component";
    static constexpr const char* DESCRIPTION = "message_mesh enclosing component with message_port/port_slot/route_message";
    static constexpr bool OWN_KERNEL = true; // a window kernel compiled for this element alone when a rank holds no other
    static constexpr uint32_t    AUX_BYTES    = 2 * MT_N * 4;
    static constexpr bool        HAS_PAYLOAD  = false;
    // the aux block is a MersenneRNG state the engine seeds warp-cooperatively before construct() (seed_mt_kernel):
    // RouteMessage ctor, MersenneRNG(my_id + 100) (enclosingComponent.cc:164-170).  The seed words go to the second
    // buffer; maintain(MAINT_INIT) builds the first generation from them.
    static constexpr bool     AUX_MT_SEED     = true;
    static constexpr uint32_t AUX_SEED_OFFSET = MT_N * 4;
    __device__ static uint32_t auxSeed(const int64_t* p) { return (uint32_t)(p[0] + 100); }
    struct State
    {
        // [0,16) mutable core
        int32_t  message_count;
        uint32_t mt_pos;
        // RouteMessage's "msg_count" statistic only ever sees addData(1) (enclosingComponent.cc:179-186), so its
        // Accumulator after n calls is (Sum, SumSQ, Count, Min, Max) = (n, n, n, 1, 1): it is kept as n alone
        uint64_t stat_n;
        // [16,32) construction constants
        uint16_t ngroups;
        uint16_t digest_ok; // the route digest is maintained: one link per port group, 2 or 4 ports
        uint32_t mod;
        uint64_t counts; // links per port group, 8 bits each
    };
    static constexpr uint32_t STORE_END = 16, CONST_END = 32, STATS_END = 32, PERSIST_BYTES = 32;
    static constexpr uint32_t RECORD_BYTES = 112, DIGEST_OFF = 32; // 28 words: a warp's 16-byte accesses to 32 records are bank-conflict free
    static constexpr bool     TIES_INTERCHANGEABLE = true;
    static constexpr bool     STATS_EVERY_EVENT    = false;
    // EnclosingComponent::handleEvent / RouteMessage::send look at neither the port an event came in on nor at its
    // (empty) payload: events that are due at the same time are interchangeable, whatever order they are run in the
    // component ends in the same state and sends the same events (enclosingComponent.cc:92-99,179-186)

    // ---- event-parallel form (window_parallel_kernel).  RouteMessage::send's effect for the k-th event of a window
    // is a pure function of the window-start record and k: its pair of draws starts at word index + 2k, the port is
    // that pair's digest entry, and the state update (message_count, MT position, msg_count statistic) is a sum.
    static constexpr bool     PARALLEL_EVENTS     = true;
    static constexpr uint32_t PARALLEL_MAX_EVENTS = MT_MAX_PAIRS_PER_WINDOW; // 100
    // component thread, start of the window: may this component's events run in the event-parallel form (even MT19937
    // index, route digest maintained)?  hdr = the digest entries of the window's first 16 pairs (2 bits each)
    __device__ static bool parHeader(const uint8_t* rec, uint32_t& hdr)
    {
        const uint32_t* r32 = reinterpret_cast<const uint32_t*>(rec);
        const uint32_t  idx = r32[1] & ((1u << MT_POS_SHIFT) - 1);
        hdr                 = 0;
        if ( (r32[4] >> 16) == 0 || (idx & 1u) != 0 ) return false;
        hdr = route_digest_window16(r32 + DIGEST_OFF / 4, idx >> 1);
        return true;
    }
    // the port the k-th event of the window leaves on
    __device__ static uint32_t parRoute(uint32_t hdr, const uint8_t* rec, uint32_t k)
    {
        if ( k < 16 ) return (hdr >> (2 * k)) & 3u;
        const uint32_t* r32 = reinterpret_cast<const uint32_t*>(rec);
        return route_digest_entry(r32 + DIGEST_OFF / 4, ((r32[1] & ((1u << MT_POS_SHIFT) - 1)) >> 1) + k);
    }
    // all n events of the window at once on the mutable core; returns the maintenance the component needs afterwards
    __device__ static uint32_t parFinish(uint4& core, uint32_t n, bool stats_on)
    {
        core.x += n;
        const uint32_t op = mt_pos_advance_pairs(core.y, n);
        if ( stats_on ) {
            const uint64_t sn = (((uint64_t)core.w << 32) | core.z) + n; // n x addData(1)
            core.z            = (uint32_t)sn;
            core.w            = (uint32_t)(sn >> 32);
        }
        return op;
    }
    // warp-cooperative maintenance (maintenance_kernel): gen = 624 words of shared memory owned by the warp
    __device__ static void maintain(uint32_t op, uint8_t* rec, uint8_t* aux, const int64_t* param, uint32_t* gen,
        uint32_t lane);

    template <class Ctx>
    __device__ static MtGenDev generator(Ctx& ctx, const State& s)
    {
        return MtGenDev { (uint32_t*)ctx.aux, s.mt_pos,
            s.digest_ok ? reinterpret_cast<uint32_t*>(ctx.record() + DIGEST_OFF) : nullptr, s.ngroups };
    }
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t* aux)
    {
        const uint32_t pos = __ldg(reinterpret_cast<const uint32_t*>(state) + 1);
        prefetch_l2(aux + ((pos >> MT_POS_SHIFT) * MT_N + (pos & ((1u << MT_POS_SHIFT) - 1))) * 4);
    }
    template <class Ctx>
    __device__ static void construct(Ctx& ctx, State& s, const int64_t* p)
    {
        s.message_count = 0;
        s.mt_pos        = 0;
        s.stat_n        = 0;
        s.mod           = (uint32_t)p[1];
        s.ngroups       = (uint16_t)p[4];
        s.counts        = (uint64_t)p[5];
        s.digest_ok     = digestOk(p) ? 1 : 0;
    }
    __device__ static bool digestOk(const int64_t* p)
    {
        const uint64_t ng = (uint64_t)p[4];
        return (ng == 2 || ng == 4) && (uint64_t)p[5] == (0x0101010101010101ull >> (8 * (8 - ng)));
    }
    template <class Ctx>
    __device__ static void setup(Ctx& ctx, State& s)
    {
        // RouteMessage::sendInitialEvents :188-199
        MtGenDev rng  = generator(ctx, s);
        uint32_t flat = 0;
        for ( uint32_t i = 0; i < s.ngroups; i++ ) {
            uint32_t cnt = (uint32_t)((s.counts >> (8 * i)) & 0xFF);
            for ( uint32_t j = 0; j < cnt; j++, flat++ )
                if ( (rng.u32() % s.mod) == 0 ) ctx.send((int)flat, 0, 0);
        }
        s.mt_pos = rng.pos;
    }
    template <class Ctx>
    __device__ static void handleEvent(Ctx& ctx, State& s, int /*port*/, uint64_t& payload)
    {
        s.message_count++;
        MtGenDev       rng = generator(ctx, s);
        const uint32_t r0 = rng.u32(), r1 = rng.u32(); // the two draws of RouteMessage::send
        s.mt_pos = rng.pos;
        // nextport = r0 % ports.size(); portnum = r1 % counts[nextport] (both draws are always made)
        const uint32_t ng       = s.ngroups;
        const uint32_t nextport = r0 % ng;
        uint32_t       portnum = 0, base = nextport;
        if ( s.counts != (0x0101010101010101ull >> (8 * (8 - ng))) ) { // not the one-link-per-port-group case
            const uint32_t cnt = (uint32_t)((s.counts >> (8 * nextport)) & 0xFF);
            portnum            = r1 % cnt;
            base               = 0;
            for ( uint32_t g = 0; g < nextport; ++g )
                base += (uint32_t)((s.counts >> (8 * g)) & 0xFF);
        }
        ctx.send((int)(base + portnum), 0, payload);
        if ( ctx.stats_on ) s.stat_n++; // mcnt_->addData(1)
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx&, State&, uint64_t)
    {
        return true;
    }
    // the statistic lives in the mutable core (stat_n): there is no statistics region to fold
    __device__ static void initStats(State&) {}
    __device__ static void mergeStats(State&, const State&) {}
    __device__ static uint64_t statsVersion(const State&) { return 0; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = (uint64_t)(int64_t)s.message_count;
        Accumulator<uint64_t> mcnt;
        mcnt.init();
        if ( s.stat_n ) mcnt.addN(1, s.stat_n); // addData_impl_Ntimes, stataccumulator.h:97-103
        mcnt.out(r + 1);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// pdes.phold (oracle/probe/pdes_elements.cc PholdNode)
struct PholdElement
{
    static constexpr uint32_t    TYPE        = TYPE_PHOLD;
    static constexpr const char* ELI_NAME    = "pdes.phold";
    static constexpr const char* DESCRIPTION = "PHOLD-style node, integer delays, 8-byte payload";
    static constexpr bool OWN_KERNEL = true; // a window kernel compiled for this element alone when a rank holds no other
    static constexpr uint32_t    AUX_BYTES   = 0;
    static constexpr bool        AUX_MT_SEED = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool        HAS_PAYLOAD = true;
    struct State
    {
        // [0,32) mutable core
        uint64_t    recv, hash;
        XorShiftDev rng;
        // [32,48) construction constants
        uint32_t dmod, init;
        uint64_t dmod_magic; // fastmod_magic(dmod)
        // [48,96) statistics
        Accumulator<uint64_t> delay;
        uint64_t              pad1;
    };
    static constexpr uint32_t STORE_END = 32, CONST_END = 48, STATS_END = 96;
    static constexpr bool     TIES_INTERCHANGEABLE = false; // the delivery hash / handlers see port and order
    // every delivery adds to the "delay" statistic: a window with events loads the statistics region together
    // with the rest of the state and stores it back, instead of folding a per-window accumulator afterwards
    static constexpr bool     STATS_EVERY_EVENT = true;
    static constexpr bool     PARALLEL_EVENTS = false; // handlers chain through the state (hash, RNG): sequential
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx&, State& s, const int64_t* p)
    {
        s.recv = 0;
        s.hash = 0;
        s.rng.seed((uint32_t)(p[0] + 1));
        s.init = (uint32_t)p[1];
        s.dmod = (uint32_t)p[2];
        s.dmod_magic = fastmod_magic(s.dmod);
        s.delay.init();
    }
    template <class Ctx>
    __device__ static void setup(Ctx& ctx, State& s)
    {
        for ( uint32_t i = 0; i < s.init; ++i ) {
            uint32_t pp = s.rng.u32() % ctx.nports;
            uint64_t d  = fastmod_u32(s.rng.u32(), s.dmod_magic, s.dmod);
            ctx.send((int)pp, d, d);
        }
    }
    template <class Ctx>
    __device__ static void handleEvent(Ctx& ctx, State& s, int port, uint64_t& payload)
    {
        s.recv++;
        if ( ctx.stats_on ) s.delay.add(payload);
        s.hash            = mixDelivery(s.hash, ctx.now, (uint64_t)port, payload);
        const uint32_t np = ctx.nports, r0 = s.rng.u32();
        const uint32_t pp = (np & (np - 1)) == 0 ? (r0 & (np - 1)) : r0 % np;
        const uint64_t d  = fastmod_u32(s.rng.u32(), s.dmod_magic, s.dmod);
        ctx.send((int)pp, d, d);
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx&, State&, uint64_t)
    {
        return true;
    }
    // ordered event-parallel form (ordered_form.cuh): no clock, at most one send per delivery, the handler touches
    // nothing but its State, the event and the context's now / nports / stats_on / send / connected
    static constexpr bool     ORDERED_EVENTS = true;
    static constexpr uint32_t ORD_MAX_EVENTS = 32; // deliveries of one component per window the form takes
    // statistics are collected per window into an empty accumulator and folded into memory only if the window
    // added data: the statistics region is not even read by windows that leave it alone
    __device__ static void initStats(State& s) { s.delay.init(); }
    __device__ static void mergeStats(State& mem, const State& d) { mem.delay.merge(d.delay); }
    // number of addData calls in the accumulators: tells the engine whether a window touched the statistics
    __device__ static uint64_t statsVersion(const State& s) { return s.delay.count; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = s.recv;
        r[1] = s.hash;
        s.delay.out(r + 2);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// pdes.clocknode (oracle/probe/pdes_elements.cc ClockNode; testElements/coreTest_Component.cc:25-173 style)
struct ClockNodeElement
{
    static constexpr uint32_t    TYPE        = TYPE_CLOCKNODE;
    static constexpr const char* ELI_NAME    = "pdes.clocknode";
    static constexpr const char* DESCRIPTION = "clock-driven node (coreTestComponent style)";
    static constexpr bool OWN_KERNEL = true; // a window kernel compiled for this element alone when a rank holds no other
    static constexpr uint32_t    AUX_BYTES   = 96; // the statistics (STATS_IN_AUX): the record itself is the hot 64 bytes
    static constexpr bool        AUX_MT_SEED = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool        HAS_PAYLOAD = false;
    struct State
    {
        // [0,48) mutable core; [0,32) is all a clock tick changes (TICK_STORE_END)
        MarsagliaDev rng;
        uint32_t     neighbor;
        uint32_t     pad0;
        uint64_t     ticks, last_cycle;
        uint64_t     recv, hash;
        // [48,64) construction constants
        int32_t  commFreq;
        uint32_t pad1;
        uint64_t comm_magic; // fastmod_magic(commFreq)
        // [64,160) statistics
        Accumulator<int32_t> count[4];
    };
    static constexpr uint32_t STORE_END = 48, CONST_END = 64, STATS_END = 160;
    // clock form (clock_form.cuh): one clock handler, no self links, no console output, nothing but State / ctx used
    static constexpr bool     CLOCK_FORM = true;
    static constexpr uint32_t TICK_STORE_END = 32; // a window without deliveries leaves [32, STORE_END) as it was
    static constexpr bool     TIES_INTERCHANGEABLE = false; // the delivery hash / handlers see port and order
    static constexpr bool     STATS_EVERY_EVENT = false;
    static constexpr bool     PARALLEL_EVENTS = false; // handlers chain through the state (hash, RNG): sequential
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = CONST_END;
    static constexpr bool     STATS_IN_AUX = true; // [CONST_END, STATS_END) lives at the start of the aux block
    static_assert(AUX_BYTES == STATS_END - CONST_END, "aux block = the statistics");
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx&, State& s, const int64_t* p)
    {
        s.ticks = s.last_cycle = s.recv = s.hash = 0;
        s.rng.z    = (uint32_t)(11 + p[0]);
        s.rng.w    = 272727u;
        s.commFreq   = (int32_t)p[1];
        s.comm_magic = fastmod_magic((uint32_t)s.commFreq);
        s.neighbor   = s.rng.u32() % 4;
        for ( int i = 0; i < 4; ++i )
            s.count[i].init();
    }
    template <class Ctx>
    __device__ static void setup(Ctx&, State&)
    {}
    template <class Ctx>
    __device__ static void handleEvent(Ctx& ctx, State& s, int port, uint64_t&)
    {
        s.recv++;
        s.hash = mixDelivery(s.hash, ctx.now, (uint64_t)port, s.ticks);
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx& ctx, State& s, uint64_t cycle)
    {
        s.ticks++;
        s.last_cycle = cycle;
        if ( divisible_i32(rng_i32(s.rng), s.comm_magic, s.commFreq) ) {
            s.neighbor = (s.neighbor + 1) % 4;
            if ( s.neighbor < ctx.nports && ctx.connected((int)s.neighbor) ) {
                ctx.send((int)s.neighbor, 0, 0);
                if ( ctx.stats_on ) add_one_of4(s.count, s.neighbor);
            }
        }
        return false;
    }
    // statistics are collected per window into an empty accumulator and folded into memory only if the window
    // added data: the statistics region is not even read by windows that leave it alone
    __device__ static void initStats(State& s) { for ( int i = 0; i < 4; ++i ) s.count[i].init(); }
    __device__ static void mergeStats(State& mem, const State& d) { for ( int i = 0; i < 4; ++i ) mem.count[i].merge(d.count[i]); }
    // number of addData calls in the accumulators: tells the engine whether a window touched the statistics
    __device__ static uint64_t statsVersion(const State& s) { return s.count[0].count + s.count[1].count + s.count[2].count + s.count[3].count; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = s.ticks;
        r[1] = s.last_cycle;
        r[2] = s.recv;
        r[3] = s.hash;
        for ( int i = 0; i < 4; ++i )
            s.count[i].out(r + 4 + 5 * i);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// coreTestElement.coreTestComponent (testElements/coreTest_Component.cc:25-173; tests/test_Component.py): one clock
// handler; every tick draws generateNextInt32() and, with probability 1/commFreq, sends an event to the next of its
// four neighbours (N, S, E, W in turn) and counts it in that direction's statistic.  All components seed
// MarsagliaRNG(11, 272727).  `neighbor` is a C++ int: the constructor's generateNextInt32() % 4 may be negative, and
// then the first sends fall into the switch's default ("bad neighbor", nothing sent) until it has counted up to 0.
// The event's payload (commSize bytes of 1) is only summed into itself by the receiver and never observed, so the
// device event carries none; workPerCycle is a busy loop without effect.
struct CoreTestElement
{
    static constexpr uint32_t    TYPE        = TYPE_CORETEST;
    static constexpr const char* ELI_NAME    = "coreTestElement.coreTestComponent";
    static constexpr const char* DESCRIPTION = "coreTestComponent: clock handler sending to its N/S/E/W neighbours in turn";
    static constexpr bool OWN_KERNEL = true; // a window kernel compiled for this element alone when a rank holds no other
    static constexpr uint32_t    AUX_BYTES   = 96; // the statistics (STATS_IN_AUX): the record itself is the hot 32 bytes
    static constexpr bool        AUX_MT_SEED = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool        HAS_PAYLOAD = false;
    struct State
    {
        // [0,16) mutable core
        MarsagliaDev rng;
        int32_t      neighbor;
        uint32_t     bad; // "bad neighbor" lines the reference would have printed
        // [16,32) construction constants
        int32_t  commFreq;
        uint32_t pad0;
        uint64_t comm_magic; // fastmod_magic(commFreq)
        // [32,128) statistics N, S, E, W
        Accumulator<int32_t> count[4];
    };
    static constexpr uint32_t STORE_END = 16, CONST_END = 32, STATS_END = 128;
    static constexpr bool     CLOCK_FORM = true; // clock form (clock_form.cuh)
    static constexpr bool     TIES_INTERCHANGEABLE = true; // handleEvent looks at nothing but the payload it deletes
    static constexpr bool     STATS_EVERY_EVENT = false;
    static constexpr bool     PARALLEL_EVENTS      = false;
    static constexpr uint32_t PERSIST_BYTES        = sizeof(State), RECORD_BYTES = CONST_END;
    static constexpr bool     STATS_IN_AUX = true; // [CONST_END, STATS_END) lives at the start of the aux block
    static_assert(AUX_BYTES == STATS_END - CONST_END, "aux block = the statistics");
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx&, State& s, const int64_t* p)
    {
        s.rng.z    = 11u;
        s.rng.w    = 272727u;
        s.commFreq   = (int32_t)p[1];
        s.comm_magic = fastmod_magic((uint32_t)s.commFreq);
        s.neighbor   = rng_i32(s.rng) % 4;
        s.bad        = 0;
        s.pad0       = 0;
        for ( int i = 0; i < 4; ++i )
            s.count[i].init();
    }
    template <class Ctx>
    __device__ static void setup(Ctx&, State&)
    {}
    template <class Ctx>
    __device__ static void handleEvent(Ctx&, State&, int, uint64_t&)
    {}
    template <class Ctx>
    __device__ static bool clockTick(Ctx& ctx, State& s, uint64_t)
    {
        if ( divisible_i32(rng_i32(s.rng), s.comm_magic, s.commFreq) ) {
            s.neighbor = (s.neighbor + 1) % 4; // C++ remainder: stays negative until it reaches 0
            if ( s.neighbor >= 0 ) {
                ctx.send(s.neighbor, 0, 0);
                if ( ctx.stats_on ) add_one_of4(s.count, (uint32_t)s.neighbor);
            }
            else
                s.bad++;
        }
        return false;
    }
    __device__ static void initStats(State& s) { for ( int i = 0; i < 4; ++i ) s.count[i].init(); }
    __device__ static void mergeStats(State& mem, const State& d) { for ( int i = 0; i < 4; ++i ) mem.count[i].merge(d.count[i]); }
    __device__ static uint64_t statsVersion(const State& s) { return s.count[0].count + s.count[1].count + s.count[2].count + s.count[3].count; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = s.bad;
        for ( int i = 0; i < 4; ++i )
            s.count[i].out(r + 4 + 5 * i);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// coreTestElement.coreTestRNGComponent (testElements/coreTest_RNGComponent.cc:26-74 ctor, :87-109 tick)
struct RngElement
{
    static constexpr uint32_t    TYPE        = TYPE_RNGCOMP;
    static constexpr const char* ELI_NAME    = "coreTestElement.coreTestRNGComponent";
    static constexpr const char* DESCRIPTION = "coreTestRNGComponent: 1 GHz clock drawing from an SST::RNG generator";
    static constexpr uint32_t    AUX_BYTES   = MT_N * 4;
    static constexpr bool        AUX_MT_SEED = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool        HAS_PAYLOAD = false;
    struct State
    {
        int64_t  rng_count, max_count;
        uint64_t last[5];
        uint64_t hash;
        uint32_t kind, a, b, c, d; // generator state: mersenne idx in a; marsaglia z,w in a,b; xorshift x,y,z,w
        uint32_t pad0;
        uint64_t pad1;
    };
    static constexpr uint32_t STORE_END = 96, CONST_END = 96, STATS_END = 96;
    static constexpr bool     TIES_INTERCHANGEABLE = false; // the delivery hash / handlers see port and order
    static constexpr bool     STATS_EVERY_EVENT = false;
    static constexpr bool     PARALLEL_EVENTS = false; // handlers chain through the state (hash, RNG): sequential
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx& ctx, State& s, const int64_t* p)
    {
        s.rng_count = 0;
        s.max_count = p[3];
        s.hash      = 0;
        for ( int i = 0; i < 5; ++i )
            s.last[i] = 0;
        s.kind = (uint32_t)p[0];
        s.a = s.b = s.c = s.d = 0;
        if ( s.kind == 0 ) {
            mt_seed((uint32_t*)ctx.aux, (uint32_t)p[1]);
        }
        else if ( s.kind == 1 ) {
            s.a = (uint32_t)p[1];
            s.b = (uint32_t)p[2];
        }
        else {
            s.a = (uint32_t)p[1];
        }
    }
    template <class Ctx>
    __device__ static void setup(Ctx&, State&)
    {}
    template <class Ctx>
    __device__ static void handleEvent(Ctx&, State&, int, uint64_t&)
    {}
    template <class G>
    __device__ static void draw(G& g, State& s)
    {
        double   nU  = g.uniform();
        uint32_t U32 = g.u32();
        uint64_t U64 = rng_u64(g);
        int32_t  I32 = rng_i32(g);
        int64_t  I64 = rng_i64(g);
        s.last[0]    = (uint64_t)__double_as_longlong(nU);
        s.last[1]    = U32;
        s.last[2]    = U64;
        s.last[3]    = (uint64_t)(int64_t)I32;
        s.last[4]    = (uint64_t)I64;
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx& ctx, State& s, uint64_t)
    {
        if ( s.kind == 0 ) {
            MersenneDev g { (uint32_t*)ctx.aux, s.a };
            draw(g, s);
            s.a = g.idx;
        }
        else if ( s.kind == 1 ) {
            MarsagliaDev g { s.a, s.b };
            draw(g, s);
            s.a = g.z;
            s.b = g.w;
        }
        else {
            XorShiftDev g { s.a, s.b, s.c, s.d };
            draw(g, s);
            s.a = g.x;
            s.b = g.y;
            s.c = g.z;
            s.d = g.w;
        }
        s.rng_count++;
        for ( int i = 0; i < 5; ++i )
            s.hash = mixDelivery(s.hash, (uint64_t)s.rng_count, (uint64_t)i, s.last[i]);
        if ( s.rng_count == s.max_count ) {
            ctx.primaryComponentOKToEndSim();
            return true;
        }
        return false;
    }
    // statistics are collected per window into an empty accumulator and folded into memory only if the window
    // added data: the statistics region is not even read by windows that leave it alone
    __device__ static void initStats(State& s) {  }
    __device__ static void mergeStats(State& mem, const State& d) {  }
    // number of addData calls in the accumulators: tells the engine whether a window touched the statistics
    __device__ static uint64_t statsVersion(const State& s) { return 0; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = (uint64_t)s.rng_count;
        for ( int i = 0; i < 5; ++i )
            r[1 + i] = s.last[i];
        r[6] = s.hash;
    }
};

// ------------------------------------------------------------------------------------------------------------------
// coreTestElement.coreTestClockerComponent (testElements/coreTest_ClockerComponent.{h,cc}; tests/test_Clocks.py,
// tests/refFiles/test_Clocks_basic.out).  Ports: 0 "left", 1 "right", 2 "inst_link" (self link).  Handlers: 0..3 the
// test clocks registered with registerClock (0, 1 old style: unregisterClock / reregisterClock; 2, 3: activate /
// deactivate), 4..8 the ordered handlers (6 and 7 registered by the anonymous and the user ClockSubComponent, into the
// parent's group: baseComponent.h:1426), 9 the master clock.  Output codes: sst_core_b200/elements.py _CLOCKER_FMT.
struct ClockerElement
{
    static constexpr uint32_t    TYPE            = TYPE_CLOCKER;
    static constexpr const char* ELI_NAME        = "coreTestElement.coreTestClockerComponent";
    static constexpr const char* DESCRIPTION = "coreTestClockerComponent: ten clock handlers (ordered and unordered) started and stopped over a self link";
    static constexpr uint32_t    AUX_BYTES       = 0;
    static constexpr bool        AUX_MT_SEED     = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool HAS_PAYLOAD = true; // IntEvent / OpEvent
    static constexpr bool MULTI_CLOCK = true;
    static constexpr bool SELF_LINKS  = true;
    static constexpr bool PRINTS      = true; // calls ctx.output
    static constexpr uint64_t MASTER_PERIOD = 5000, TEST_PERIOD = 1000; // "5ns", "1ns" at the 1 ps time base
    static constexpr int      TOTAL_COUNT   = 3;
    enum { OP_NOP, OP_START, OP_STOP, OP_TERM };
    struct State
    {
        int32_t         id_, inst_count, ordered_var;
        uint8_t         done, ordered_remove, pad0, pad1;
        int32_t         counter[4];
        ClockSet<2, 10> clk;
        uint64_t        pad2;
    };
    static_assert(sizeof(State) % 16 == 0, "State is moved in 16-byte pieces");
    static constexpr uint32_t STORE_END = sizeof(State), CONST_END = sizeof(State), STATS_END = sizeof(State);
    static constexpr bool     TIES_INTERCHANGEABLE = false;
    static constexpr bool     STATS_EVERY_EVENT    = false;
    static constexpr bool     PARALLEL_EVENTS      = false;
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    // coreTest_ClockerComponent.h:265-276: op << 4 | clock, 0 ends an instruction
    __device__ static uint8_t program(int inst, int k)
    {
        constexpr uint8_t S = OP_START << 4, P = OP_STOP << 4, T = OP_TERM << 4;
        static constexpr uint8_t prog[13][6] = {
            { S | 1, S | 5, S | 6, S | 7, S | 8, 0 }, { S | 0, P | 6, 0 }, { S | 3, P | 7, 0 }, { P | 1, S | 6, 0 },
            { S | 2, S | 4, 0 }, { P | 3, S | 5, S | 6, S | 8, 0 }, { S | 1, S | 6, S | 5, S | 8, S | 7, 0 }, { S | 0, 0 },
            { S | 3, 0 }, { P | 1, 0 }, { S | 2, 0 }, { P | 3, 0 }, { T | 15, 0 } };
        return inst < 13 ? prog[inst][k] : 0;
    }
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    __device__ static uint64_t tag(const State& s, int second)
    {
        return (uint64_t)(uint32_t)s.id_ | ((uint64_t)(uint32_t)second << 32);
    }
    template <class Ctx>
    __device__ static void construct(Ctx& ctx, State& s, const int64_t*)
    {
        s.id_ = -1;
        s.inst_count = 0;
        s.ordered_var = -1;
        s.done = s.ordered_remove = s.pad0 = s.pad1 = 0;
        for ( int i = 0; i < 4; ++i )
            s.counter[i] = TOTAL_COUNT;
        s.clk.init();
        s.pad2 = 0;
        if ( !ctx.connected(0) ) s.id_ = 0; // nullptr == left_
        // handler ids follow the registration order: the test clocks first here (ids 0..8), the master last (9); what
        // the reference's order between the master and the test handlers would be is not observable (different clocks)
        for ( int i = 0; i < 4; ++i )
            s.clk.deactivate(s.clk.add(TEST_PERIOD, false, 0)); // registered, then unregistered / deactivated
        for ( int i = 4; i <= 8; ++i )
            s.clk.deactivate(s.clk.add(TEST_PERIOD, true, 0));
        const int m = s.clk.add(MASTER_PERIOD, false, 0);
        if ( s.id_ != 0 ) s.clk.deactivate(m);
    }
    template <class Ctx>
    __device__ static void setup(Ctx& ctx, State& s)
    {
        if ( s.id_ != 0 ) return;
        ctx.output(0, tag(s, 0), ctx.now);
        ctx.send(1, 1 * MASTER_PERIOD, 1); // IntEvent(1), delayed by one master period
    }
    template <class Ctx>
    __device__ static void handleEvent(Ctx& ctx, State& s, int port, uint64_t& payload)
    {
        if ( port == 0 ) { // handle_left
            s.id_ = (int32_t)payload;
            payload++;
            ctx.send(1, 0, payload);
            s.clk.activate(9, ctx.now);
            ctx.output(1, tag(s, 0), ctx.now);
            return;
        }
        // inst_handler
        const int op = (int)(payload & 0xFF), clock = (int)((payload >> 8) & 0xFF);
        if ( op == OP_START ) {
            const uint64_t next = s.clk.activate(clock, ctx.now);
            ctx.output(2, tag(s, clock), next);
        }
        else if ( op == OP_STOP ) {
            s.clk.deactivate(clock);
            ctx.output(3, tag(s, clock), ctx.now);
        }
        else if ( op == OP_TERM ) {
            ctx.primaryComponentOKToEndSim();
            s.done = 1;
            ctx.output(4, tag(s, 0), ctx.now);
        }
    }
    template <class Ctx>
    __device__ static bool handler(Ctx& ctx, State& s, int hid, uint64_t cycle)
    {
        if ( hid == 9 ) { // master_handler
            if ( s.done ) return true;
            const int inst = s.inst_count++;
            for ( int k = 0; k < 6; ++k ) {
                const uint8_t b = program(inst, k);
                if ( !b ) break;
                ctx.send(2, 0, (uint64_t)(b >> 4) | ((uint64_t)(b & 15) << 8));
            }
            return false;
        }
        if ( hid <= 3 ) { // test_handler
            ctx.output(5, tag(s, hid), cycle);
            if ( hid == 0 || hid == 2 ) {
                if ( --s.counter[hid] == 0 ) {
                    ctx.output(6, tag(s, hid), ctx.now);
                    s.counter[hid] = TOTAL_COUNT;
                    return true;
                }
            }
            return false;
        }
        // test_ordered_handler
        if ( hid == 4 )
            s.ordered_remove = 1;
        else if ( hid == 5 )
            s.ordered_var = 36;
        else if ( hid == 6 )
            s.ordered_var /= 2;
        else if ( hid == 7 )
            s.ordered_var -= 6;
        else
            ctx.output(7, tag(s, s.ordered_var), cycle);
        return s.ordered_remove != 0;
    }
    template <class Ctx>
    __device__ static uint32_t clockFire(Ctx& ctx, State& s, uint64_t t)
    {
        return s.clk.fire(t, [&](int h, uint64_t cycle) { return handler(ctx, s, h, cycle); });
    }
    __device__ static uint64_t nextClock(const State& s) { return s.clk.nextTime(); }
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
        r[0] = (uint64_t)(int64_t)s.id_;
        r[1] = (uint64_t)s.inst_count;
        r[2] = s.done;
        r[3] = (uint64_t)(int64_t)s.ordered_var;
        for ( int i = 0; i < 9; ++i )
            r[4 + i] = (uint64_t)(int64_t)(i < 4 ? s.counter[i] : TOTAL_COUNT);
    }
};

// ------------------------------------------------------------------------------------------------------------------
// coreTestElement.StatisticsComponent.int (testElements/coreTest_StatisticsComponent.cc:26-140; tests/
// test_StatisticsComponent_basic.py): a 1 ns clock draws u32, u64, i32, i64, scales them and adds them to stat1_U32 ..
// stat4_I64, and to stat5_dyn once it has been registered (at tick `register_dynamic`).  Result record: [0] rng_count,
// [1 + 5 i ..] the five fields of statistic i.
struct StatsIntElement
{
    static constexpr uint32_t    TYPE            = TYPE_STATS_INT;
    static constexpr const char* ELI_NAME        = "coreTestElement.StatisticsComponent.int";
    static constexpr const char* DESCRIPTION = "StatisticsComponent.int: 1 ns clock feeding five integer accumulators under the statistics engine";
    static constexpr uint32_t    AUX_BYTES       = MT_N * 4;
    static constexpr bool        AUX_MT_SEED     = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool HAS_PAYLOAD = false;
    static constexpr bool STAT_CLOCK  = true;
    static constexpr bool PRINTS      = true; // statistic outputs are records of the output log
    struct State
    {
        int64_t              rng_count, max_count, dynamic_reg;
        uint32_t             kind, a, b, pad0; // generator: mersenne index in a; marsaglia z, w in a, b
        EngineStat<uint32_t> s1;
        EngineStat<uint64_t> s2;
        EngineStat<int32_t>  s3;
        EngineStat<int64_t>  s4, s5;
    };
    static_assert(sizeof(State) % 16 == 0, "State is moved in 16-byte pieces");
    static constexpr uint32_t STORE_END = sizeof(State), CONST_END = sizeof(State), STATS_END = sizeof(State);
    static constexpr bool     TIES_INTERCHANGEABLE = true; // the ports are configured and never used
    static constexpr bool     STATS_EVERY_EVENT    = false;
    static constexpr bool     PARALLEL_EVENTS      = false;
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx& ctx, State& s, const int64_t* p)
    {
        s.rng_count   = 0;
        s.max_count   = p[3];
        s.dynamic_reg = p[4];
        s.kind        = (uint32_t)p[0];
        s.a = s.b = s.pad0 = 0;
        if ( s.kind == 0 )
            mt_seed((uint32_t*)ctx.aux, (uint32_t)p[1]);
        else {
            s.a = (uint32_t)p[1];
            s.b = (uint32_t)p[2];
        }
        const int64_t* x = ctx.xparam;
        s.s1.configure(ctx.nxparam >= 6 ? x : nullptr);
        s.s2.configure(ctx.nxparam >= 12 ? x + 6 : nullptr);
        s.s3.configure(ctx.nxparam >= 18 ? x + 12 : nullptr);
        s.s4.configure(ctx.nxparam >= 24 ? x + 18 : nullptr);
        s.s5.configure(ctx.nxparam >= 30 ? x + 24 : nullptr);
        s.s1.c.registerNow(0);
        s.s2.c.registerNow(0);
        s.s3.c.registerNow(0);
        s.s4.c.registerNow(0);
    }
    template <class Ctx>
    __device__ static void setup(Ctx&, State&)
    {}
    template <class Ctx>
    __device__ static void handleEvent(Ctx&, State&, int, uint64_t&)
    {}
    template <class Ctx, class G>
    __device__ static void draw(Ctx& ctx, State& s, G& g)
    {
        const uint32_t U32 = g.u32();
        const uint64_t U64 = rng_u64(g);
        const int32_t  I32 = rng_i32(g);
        const int64_t  I64 = rng_i64(g);
        s.rng_count++;
        s.s1.addData(ctx, 0, U32 / 10000000u);
        s.s2.addData(ctx, 1, U64 / 1000000000000000ull);
        s.s3.addData(ctx, 2, I32 / 10000000);
        s.s4.addData(ctx, 3, I64 / 1000000000000000ll);
        s.s5.addData(ctx, 4, I64 / 1000000000000000ll);
        if ( s.rng_count == s.dynamic_reg ) s.s5.c.registerNow(ctx.now);
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx& ctx, State& s, uint64_t)
    {
        if ( s.kind == 0 ) {
            MersenneDev g { (uint32_t*)ctx.aux, s.a };
            draw(ctx, s, g);
            s.a = g.idx;
        }
        else {
            MarsagliaDev g { s.a, s.b };
            draw(ctx, s, g);
            s.a = g.z;
            s.b = g.w;
        }
        if ( s.rng_count >= s.max_count ) {
            ctx.primaryComponentOKToEndSim();
            return true;
        }
        return false;
    }
    __device__ static uint64_t nextStatTime(const State& s)
    {
        uint64_t t = s.s1.c.nextTime();
        uint64_t u = s.s2.c.nextTime();
        t          = u < t ? u : t;
        u          = s.s3.c.nextTime();
        t          = u < t ? u : t;
        u          = s.s4.c.nextTime();
        t          = u < t ? u : t;
        u          = s.s5.c.nextTime();
        return u < t ? u : t;
    }
    template <class Ctx>
    __device__ static void statFire(Ctx& ctx, State& s, uint64_t t)
    {
        s.s1.fire(ctx, 0, t);
        s.s2.fire(ctx, 1, t);
        s.s3.fire(ctx, 2, t);
        s.s4.fire(ctx, 3, t);
        s.s5.fire(ctx, 4, t);
    }
    __device__ static void initStats(State&) {}
    __device__ static void mergeStats(State&, const State&) {}
    __device__ static uint64_t statsVersion(const State&) { return 0; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = (uint64_t)s.rng_count;
        s.s1.fields(r + 1);
        s.s2.fields(r + 6);
        s.s3.fields(r + 11);
        s.s4.fields(r + 16);
        s.s5.fields(r + 21);
    }
};

// coreTestElement.StatisticsComponent.float (coreTest_StatisticsComponent.cc:160-260): two nextUniform() per tick;
// stat1_F32 gets (float)(a * 1000), stat2_F64 b * 1000, stat3_F64 b * 1000 + 10.  A slot may add to another slot's
// statistic object (createStatistic + setStatistic: extra parameter [4] of the slot).
struct StatsFloatElement
{
    static constexpr uint32_t    TYPE            = TYPE_STATS_FLOAT;
    static constexpr const char* ELI_NAME        = "coreTestElement.StatisticsComponent.float";
    static constexpr const char* DESCRIPTION = "StatisticsComponent.float: 1 ns clock feeding three floating-point accumulators under the statistics engine";
    static constexpr uint32_t    AUX_BYTES       = MT_N * 4;
    static constexpr bool        AUX_MT_SEED     = false;
    static constexpr uint32_t    AUX_SEED_OFFSET = 0;
    __device__ static uint32_t auxSeed(const int64_t*) { return 0; }
    static constexpr bool HAS_PAYLOAD = false;
    static constexpr bool STAT_CLOCK  = true;
    static constexpr bool PRINTS      = true; // statistic outputs are records of the output log
    struct State
    {
        int64_t            rng_count, max_count;
        uint32_t           kind, a, b, target2, target3, pad0, pad1, pad2;
        EngineStat<float>  s1;
        EngineStat<double> s2, s3;
        uint64_t           pad3;
    };
    static_assert(sizeof(State) % 16 == 0, "State is moved in 16-byte pieces");
    static constexpr uint32_t STORE_END = sizeof(State), CONST_END = sizeof(State), STATS_END = sizeof(State);
    static constexpr bool     TIES_INTERCHANGEABLE = true;
    static constexpr bool     STATS_EVERY_EVENT    = false;
    static constexpr bool     PARALLEL_EVENTS      = false;
    static constexpr uint32_t PERSIST_BYTES = sizeof(State), RECORD_BYTES = sizeof(State);
    template <class Ctx>
    __device__ static void beginWindow(Ctx&, State&, uint32_t)
    {}
    __device__ static void prefetch(const uint8_t* state, const uint8_t*) { prefetch_l2(state); }
    template <class Ctx>
    __device__ static void construct(Ctx& ctx, State& s, const int64_t* p)
    {
        s.rng_count = 0;
        s.max_count = p[3];
        s.kind      = (uint32_t)p[0];
        s.a = s.b = s.pad0 = s.pad1 = s.pad2 = 0;
        s.pad3                               = 0;
        if ( s.kind == 0 )
            mt_seed((uint32_t*)ctx.aux, (uint32_t)p[1]);
        else {
            s.a = (uint32_t)p[1];
            s.b = (uint32_t)p[2];
        }
        const int64_t* x = ctx.xparam;
        s.s1.configure(ctx.nxparam >= 6 ? x : nullptr);
        s.s2.configure(ctx.nxparam >= 12 ? x + 6 : nullptr);
        s.s3.configure(ctx.nxparam >= 18 ? x + 12 : nullptr);
        s.target2 = (ctx.nxparam >= 12 && (x[6] & ST_EXISTS)) ? (uint32_t)x[10] : 1u;
        s.target3 = (ctx.nxparam >= 18 && (x[12] & ST_EXISTS)) ? (uint32_t)x[16] : 2u;
        s.s1.c.registerNow(0);
        if ( s.target2 == 1 ) s.s2.c.registerNow(0);
        if ( s.target3 == 2 ) s.s3.c.registerNow(0);
    }
    template <class Ctx>
    __device__ static void setup(Ctx&, State&)
    {}
    template <class Ctx>
    __device__ static void handleEvent(Ctx&, State&, int, uint64_t&)
    {}
    template <class Ctx>
    __device__ static void addD(Ctx& ctx, State& s, uint32_t target, double v)
    {
        if ( target == 1 )
            s.s2.addData(ctx, 1, v);
        else if ( target == 2 )
            s.s3.addData(ctx, 2, v);
    }
    template <class Ctx, class G>
    __device__ static void draw(Ctx& ctx, State& s, G& g)
    {
        const double a = g.uniform();
        const double b = g.uniform();
        s.rng_count++;
        s.s1.addData(ctx, 0, __double2float_rn(__dmul_rn(a, 1000.0)));
        const double sb = __dmul_rn(b, 1000.0);
        addD(ctx, s, s.target2, sb);
        addD(ctx, s, s.target3, __dadd_rn(sb, 10.0));
    }
    template <class Ctx>
    __device__ static bool clockTick(Ctx& ctx, State& s, uint64_t)
    {
        if ( s.kind == 0 ) {
            MersenneDev g { (uint32_t*)ctx.aux, s.a };
            draw(ctx, s, g);
            s.a = g.idx;
        }
        else {
            MarsagliaDev g { s.a, s.b };
            draw(ctx, s, g);
            s.a = g.z;
            s.b = g.w;
        }
        if ( s.rng_count >= s.max_count ) {
            ctx.primaryComponentOKToEndSim();
            return true;
        }
        return false;
    }
    __device__ static uint64_t nextStatTime(const State& s)
    {
        uint64_t t = s.s1.c.nextTime();
        uint64_t u = s.s2.c.nextTime();
        t          = u < t ? u : t;
        u          = s.s3.c.nextTime();
        return u < t ? u : t;
    }
    template <class Ctx>
    __device__ static void statFire(Ctx& ctx, State& s, uint64_t t)
    {
        s.s1.fire(ctx, 0, t);
        s.s2.fire(ctx, 1, t);
        s.s3.fire(ctx, 2, t);
    }
    __device__ static void initStats(State&) {}
    __device__ static void mergeStats(State&, const State&) {}
    __device__ static uint64_t statsVersion(const State&) { return 0; }
    __device__ static void results(const State& s, uint64_t* r)
    {
        r[0] = (uint64_t)s.rng_count;
        s.s1.fields(r + 1);
        s.s2.fields(r + 6);
        s.s3.fields(r + 11);
    }
};

#endif // __CUDACC__

// ---- the registry: what SST_ELI_REGISTER_COMPONENT's static initialisers build in the reference
// (eli/elibase.h:124-147 LoadedLibraries); here a constexpr table, queried through sstb200_component_type_*.
struct ElementInfo
{
    uint32_t    type;
    const char* eli_name;
    uint32_t    state_bytes;
    uint32_t    aux_bytes;
    int         has_payload;
    bool        parallel_events;                  // has an event-parallel handler
    bool        ties_interchangeable;             // events of equal delivery time are interchangeable for the element
    const char* description;
    bool        prints = false;                   // calls ctx.output(): the engine keeps an output log
    bool        self_links = false;               // configures self links
    bool        stat_clock = false;               // has statistics under the statistics engine (STAT_CLOCK)
    bool        ordered_events = false;           // has the ordered event-parallel form (ORDERED_EVENTS)
    bool        clock_form = false;               // has the clock form (CLOCK_FORM)
    bool        stats_in_aux = false;             // keeps its statistics in its aux block (STATS_IN_AUX)
};

} // namespace sstb200
