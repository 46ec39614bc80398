// TEST INFRASTRUCTURE ONLY.  CPU restatement ("golden model") of SST-core's event-delivery hot path.
//
// This is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load it.  It restates, event by event and serially, what the reference does, so that the CUDA engine
// can be checked at sizes the reference itself cannot hold.
//
// Parity of THIS file is pinned (tests/test_oracle.py) against
//   * the reference's golden vectors: tests/refFiles/test_MessageMesh.out, test_RNGComponent_{mersenne,marsaglia,
//     xorshift}.out (copied as fixtures under tests/golden/ by tests/golden/make_golden.py), and
//   * outputs of the reference binary itself (oracle/_ref/libexec/sstsim.x) for PHOLD, clock populations and
//     per-component event-order traces (pdes.trace probe).
//
// Reference lines followed (all under /root/reference/src/sst/core/):
//   activity.h:64-73,230-238   total order (delivery_time, priority<<32|order_tag, queue_order); priorities :29-40
//   impl/timevortex/timeVortexPQ.cc:62-85   insert stamps queue_order = insertOrder++, pop = heap min
//   link.cc:623-658            send: t = now + delay*timebase + latency; stamp (tag, delivery_info); insert
//   baseComponent.cc:541-642   order tag = 1-based sequence number of configureLink calls (component id order)
//   link.cc:469-479            self links keep tag 0xFFFFFFFF (configureSelfLink: not used by in-scope models)
//   clock.cc:314-420           one Clock per period; handlers in registration order; first tick at `period`;
//                              handler returning true is unregistered; re-insert at now+period
//   simulation.cc:1099-1153    run loop: pop, set time, execute
//   stopAction.h:27-57         StopAction priority 30: activities AT stop time with priority > 30 do not run
//   exit.cc:53-79              primary-component refcount; serial: Exit action at now+1 (priority 99)
//   rng/mersenne.cc:49-159, rng/marsaglia.cc:29-148, rng/xorshift.cc:44-140   generators
//   rng/discrete.h:40-109, expon.h:73-77, gaussian.h:81-110, poisson.h:73-85, uniform.h   distributions
//   statapi/stataccumulator.h:89-103,171-194   accumulator update and output fields
//   testElements/message_mesh/enclosingComponent.cc:29-199, testElements/coreTest_RNGComponent.cc:26-109
//   and oracle/probe/pdes_elements.cc for the PHOLD / clocknode models (ours, run by the reference core).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------ RNG
struct Random
{
    virtual ~Random() {}
    virtual uint32_t u32()     = 0;
    virtual double   uniform() = 0;
    int32_t          i32()
    {
        uint32_t v = u32();
        int32_t  r;
        memcpy(&r, &v, 4);
        return r;
    }
    int64_t i64()
    {
        uint32_t lo = u32();
        uint32_t hi = u32();
        uint64_t v  = (uint64_t)lo | ((uint64_t)hi << 32);
        int64_t  r;
        memcpy(&r, &v, 8);
        return r;
    }
    uint64_t u64()
    {
        int64_t  v = i64();
        uint64_t r;
        memcpy(&r, &v, 8);
        return r;
    }
};

// rng/mersenne.cc:49-53 (ctor), :56-68 (batch), :76-102 (draw), :149-159 (seed)
struct Mersenne : Random
{
    uint32_t numbers[624];
    int      index;
    explicit Mersenne(uint32_t seed)
    {
        numbers[0] = seed;
        index      = 0;
        for ( int i = 1; i < 624; i++ )
            numbers[i] = 1812433253u * (numbers[i - 1] ^ (numbers[i - 1] >> 30)) + (uint32_t)i;
    }
    void batch()
    {
        index = 0;
        for ( int i = 0; i < 624; ++i ) {
            uint32_t temp = (numbers[i] & 0x80000000u) + (numbers[(i + 1) % 624] & 0x7fffffffu);
            numbers[i]    = numbers[(i + 397) % 624] ^ (temp >> 1);
            if ( temp % 2 != 0 ) numbers[i] ^= 2567483615u;
        }
    }
    uint32_t u32() override
    {
        if ( index == 0 ) batch();
        uint32_t t = numbers[index];
        t ^= t >> 11;
        t ^= (t << 7) & 2636928640u;
        t ^= (t << 15) & 4022730752u;
        t ^= t >> 18;
        index = (index + 1) % 624;
        return t;
    }
    double uniform() override
    {
        double d;
        do {
            d = (double)u32() / 4294967295.0;
        } while ( d >= 1.0 );
        return d;
    }
};

// rng/marsaglia.cc:29-34 (ctor z,w), :64-87 (generateNext, nextUniform)
struct Marsaglia : Random
{
    uint32_t z, w;
    Marsaglia(uint32_t z, uint32_t w) : z(z), w(w) {}
    uint32_t u32() override
    {
        z = 36969u * (z & 65535u) + (z >> 16);
        w = 18000u * (w & 65535u) + (w >> 16);
        return (z << 16) + w;
    }
    double uniform() override
    {
        double d;
        do {
            d = (double)(uint32_t)(u32() + 1u) * 2.328306435454494e-10;
        } while ( d >= 1.0 );
        return d;
    }
};

// rng/xorshift.cc:44-50 (ctor), :58-77 (draw), :131-140 (seed)
struct XorShift : Random
{
    uint32_t x, y, z, w;
    explicit XorShift(uint32_t seed) : x(seed), y(0), z(0), w(0) {}
    uint32_t u32() override
    {
        uint32_t t = x ^ (x << 11);
        x          = y;
        y          = z;
        z          = w;
        return w   = w ^ (w >> 19) ^ t ^ (t >> 8);
    }
    double uniform() override
    {
        double d;
        do {
            d = (double)u32() / 4294967295.0;
        } while ( d >= 1.0 );
        return d;
    }
};

Random*
makeRng(int kind, uint64_t s1, uint64_t s2)
{
    switch ( kind ) {
    case 0:
        return new Mersenne((uint32_t)s1);
    case 1:
        return new Marsaglia((uint32_t)s1, (uint32_t)s2);
    default:
        return new XorShift((uint32_t)s1);
    }
}

// ------------------------------------------------------------------------------------------------ statistics
// statapi/stataccumulator.h:89-95 (addData_impl), statbase.h:393-401 (count++), :171-194 (output fields)
template <typename T>
struct Accumulator
{
    T        sum = 0, sumsq = 0;
    T        mn = std::numeric_limits<T>::max();
    T        mx = std::numeric_limits<T>::min();
    uint64_t count = 0;
    void     add(T v)
    {
        sum += v;
        sumsq += v * v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        count++;
    }
    // Output order Sum, SumSQ, Count, Min, Max; Min/Max print 0 when nothing was collected.
    void out(uint64_t* r) const
    {
        r[0] = (uint64_t)(int64_t)sum;
        r[1] = (uint64_t)(int64_t)sumsq;
        r[2] = count;
        r[3] = count ? (uint64_t)(int64_t)mn : 0;
        r[4] = count ? (uint64_t)(int64_t)mx : 0;
    }
};

// ------------------------------------------------------------------------------------------------ engine
enum : uint64_t { PRIO_STOP = 30, PRIO_CLOCK = 40, PRIO_EVENT = 50, PRIO_EXIT = 99 };
enum { KIND_EVENT, KIND_CLOCK, KIND_STOP, KIND_EXIT, KIND_STAT };
constexpr uint64_t PRIO_STATCLOCK = 85; // STATISTICCLOCKPRIORITY activity.h:38
enum { TYPE_MESH = 1, TYPE_PHOLD = 2, TYPE_CLOCKNODE = 3, TYPE_RNGCOMP = 4, TYPE_CORETEST = 5, TYPE_CLOCKER = 6, TYPE_STATS_INT = 7,
       TYPE_STATS_FLOAT = 8 };
constexpr int      NPARAM   = 8;
constexpr int      NRESULT  = 32;
constexpr uint32_t NO_PEER  = 0xFFFFFFFFu;

struct Activity
{
    uint64_t time, prio_order, queue_order;
    int      kind;
    uint32_t target; // event: receiving directed port; clock: clock index
    uint64_t payload;
};
struct ActGreater
{
    bool operator()(const Activity& a, const Activity& b) const
    {
        if ( a.time != b.time ) return a.time > b.time;
        if ( a.prio_order != b.prio_order ) return a.prio_order > b.prio_order;
        return a.queue_order > b.queue_order;
    }
};

inline uint64_t
mixDelivery(uint64_t h, uint64_t time, uint64_t port, uint64_t payload)
{
    uint64_t v = time * 0x9E3779B97F4A7C15ull + (port + 1) * 0xC2B2AE3D27D4EB4Full + payload;
    h ^= v;
    h *= 0x100000001B3ull;
    h ^= h >> 29;
    return h;
}

struct Sim;
struct Component
{
    Sim*     sim  = nullptr;
    uint32_t id   = 0; // global component id (creation order)
    uint32_t port0 = 0, nports = 0;
    virtual ~Component() {}
    virtual void setup() {}
    virtual void handleEvent(int /*port*/, uint64_t& /*payload*/) {}
    virtual bool tick(uint64_t /*cycle*/) { return true; }
    // Clock::Handler<cls, &cls::fn, int>(this, hid): several handlers per component (coreTest_ClockerComponent.cc);
    // a component with one handler only overrides tick()
    virtual bool clockHandler(int /*hid*/, uint64_t cycle) { return tick(cycle); }
    // statistics engine activities of this component (priority 85): what = slot | kind << 8
    virtual void statEvent(uint32_t /*what*/) {}
    virtual void endOfSimulation() {} // StatisticProcessingEngine::endOfSimulation statengine.cc:296-311
    virtual void results(uint64_t* /*r*/) const {}
    void         send(int port, uint64_t delay, uint64_t payload);
    void         output(uint32_t code, uint64_t a, uint64_t b); // getSimulationOutput().output(...): one log record
};

// clock.h:60-260 Clock::HandlerBase as seen by the core: which clock (or ordered group) it belongs to, whether it is
// in the call list, its position among the ordered handlers of its group (order_ < 0: unordered)
struct ClockObj;
struct ClockGroup;
struct ClockHandler
{
    uint32_t    comp   = 0;
    int         hid    = 0;
    bool        active = false;
    int         order  = -1;
    ClockObj*   clock  = nullptr;
    ClockGroup* group  = nullptr;
};
// clock.h Clock::Group / clock.cc:25-155: the ordered handlers of ONE component on one clock, called in registration
// order after the clock's unordered handlers
struct ClockGroup
{
    ClockObj*                  clock = nullptr;
    uint32_t                   comp  = 0;
    std::vector<ClockHandler*> all; // by order_
    bool                       in_list = false;
    bool anyActive() const
    {
        for ( auto* h : all )
            if ( h->active ) return true;
        return false;
    }
};
struct ClockObj
{
    uint64_t                   period;
    std::vector<ClockHandler*> handlers; // active unordered handlers, call order = (re)registration order
    std::vector<ClockGroup*>   groups;   // active groups, activation order
    uint64_t                   cycle     = 0;
    bool                       scheduled = false;
    uint32_t                   index     = 0;
};

struct OutRec
{
    uint64_t time;
    uint32_t comp, code;
    uint64_t a, b;
};

struct TraceRec
{
    uint64_t time;
    uint32_t comp, port;
    uint64_t payload;
};

struct Sim
{
    uint32_t                 ncomp = 0;
    std::vector<Component*>  comps;
    std::vector<uint32_t>    port_off, peer_port, port_comp, port_tag;
    std::vector<uint64_t>    latency;
    std::vector<ClockObj*>   clocks;
    std::vector<ClockHandler*> all_handlers;
    std::vector<ClockGroup*>   all_groups;
    std::vector<OutRec>      out_log;
    struct OutExtra
    {
        uint64_t c, d, e;
    };
    std::vector<OutExtra>    out_extra; // third to fifth value of the records that carry five (statistic outputs)
    uint64_t                 cur_prio = 0; // priority of the activity being executed (Simulation::getCurrentPriority)
    std::priority_queue<Activity, std::vector<Activity>, ActGreater> pq;
    uint64_t                 insert_order = 0;
    uint64_t                 now = 0, stop_at = 0, end_time = 0;
    uint64_t                 events_delivered = 0, clock_ticks = 0;
    int64_t                  primary = 0;
    bool                     end_sim = false, in_run = false, stats_on = true, tracing = false;
    std::vector<TraceRec>    trace;
    // partitioned stepping (RankSyncSerialSkip contract): this Sim owns components [own_lo, own_hi); events for other
    // components are collected in `outbox` as {time, dst_port<<32|seq, payload, dst_comp} and handed to the owner
    // at the next sync, which re-inserts them (rankSyncSerialSkip.cc:291-295)
    uint32_t                 own_lo = 0, own_hi = 0xFFFFFFFFu;
    std::vector<uint64_t>    outbox;
    std::vector<uint32_t>    send_seq;

    ~Sim()
    {
        for ( auto c : comps )
            delete c;
        for ( auto c : clocks )
            delete c;
        for ( auto h : all_handlers )
            delete h;
        for ( auto g : all_groups )
            delete g;
    }
    void insert(Activity a)
    {
        a.queue_order = insert_order++;
        pq.push(a);
    }
    // ---- clocks (simulation.cc:1403-1516, clock.cc).  One Clock object per period (all handlers here run at
    // CLOCKPRIORITY); it is created and scheduled the first time anybody asks for the period.
    ClockObj* clockFor(uint64_t period)
    {
        for ( auto* c : clocks )
            if ( c->period == period ) return c;
        ClockObj* c = new ClockObj();
        c->period   = period;
        c->index    = (uint32_t)clocks.size();
        clocks.push_back(c);
        schedule(c);
        return c;
    }
    // Clock::schedule clock.cc:400-420
    void schedule(ClockObj* c)
    {
        c->cycle      = now / c->period;
        uint64_t next = c->cycle * c->period + c->period;
        if ( cur_prio < PRIO_CLOCK && now != 0 && !end_sim ) {
            if ( now % c->period == 0 ) {
                next = now;
                c->cycle--;
            }
        }
        insert({ next, PRIO_CLOCK << 32, 0, KIND_CLOCK, c->index, 0 });
        c->scheduled = true;
    }
    uint64_t nextCycle(ClockObj* c) // Clock::getNextCycle clock.cc:305-311
    {
        if ( !c->scheduled ) c->cycle = now / c->period;
        return c->cycle + 1;
    }
    void attach(ClockHandler* h) // Clock::registerHandler clock.cc:193-231
    {
        if ( h->active ) return;
        h->active = true;
        h->clock->handlers.push_back(h);
        if ( !h->clock->scheduled ) schedule(h->clock);
    }
    void activateGroup(ClockGroup* g) // Clock::activateGroup clock.cc:289-303
    {
        if ( g->in_list ) return;
        g->in_list = true;
        g->clock->groups.push_back(g);
        if ( !g->clock->scheduled ) schedule(g->clock);
    }
    ClockHandler* newHandler(uint32_t comp, uint64_t period, int hid)
    {
        ClockHandler* h = new ClockHandler();
        h->comp         = comp;
        h->hid          = hid;
        h->clock        = clockFor(period);
        all_handlers.push_back(h);
        return h;
    }
    // BaseComponent::registerClock -> Simulation::registerClock
    ClockHandler* registerClock(uint32_t comp, uint64_t period, int hid = 0)
    {
        ClockHandler* h = newHandler(comp, period, hid);
        attach(h);
        return h;
    }
    // BaseComponent::registerOrderedClock -> Component::getClockGroup (component.cc:34-45) + Group::registerHandler
    // (clock.cc:25-46)
    ClockHandler* registerOrderedClock(uint32_t comp, uint64_t period, int hid)
    {
        ClockHandler* h = newHandler(comp, period, hid);
        ClockGroup*   g = nullptr;
        for ( auto* x : all_groups )
            if ( x->comp == comp && x->clock == h->clock ) g = x;
        if ( !g ) {
            g        = new ClockGroup();
            g->clock = h->clock;
            g->comp  = comp;
            all_groups.push_back(g);
        }
        h->group  = g;
        h->order  = (int)g->all.size();
        h->active = true;
        g->all.push_back(h);
        activateGroup(g);
        return h;
    }
    // HandlerBase::activate clock.h:180-190 (= reregisterClock for unordered handlers, simulation.cc:1441-1451)
    uint64_t activate(ClockHandler* h)
    {
        if ( h->order < 0 ) {
            attach(h);
        }
        else if ( !h->active ) {
            h->active = true;
            activateGroup(h->group);
        }
        return nextCycle(h->clock);
    }
    // HandlerBase::deactivate clock.h:195-204 (= unregisterClock, clock.cc:233-250)
    void deactivate(ClockHandler* h)
    {
        if ( !h->active ) return;
        h->active = false;
        if ( h->order < 0 ) {
            auto& v = h->clock->handlers;
            for ( size_t i = 0; i < v.size(); ++i )
                if ( v[i] == h ) {
                    v.erase(v.begin() + (long)i);
                    break;
                }
        }
        // (a group notices at its next execute that nobody is active, clock.cc:104-108)
    }
    // Clock::execute clock.cc:314-397 + Clock::Group::execute clock.cc:110-155
    void executeClock(ClockObj& ck)
    {
        if ( ck.handlers.empty() && ck.groups.empty() ) {
            ck.scheduled = false;
            return;
        }
        ck.cycle++;
        for ( size_t i = 0; i < ck.handlers.size(); ) {
            ClockHandler* h = ck.handlers[i];
            clock_ticks++;
            const bool stop = comps[h->comp]->clockHandler(h->hid, ck.cycle);
            // the callee may have changed the list: find the handler again
            size_t at = ck.handlers.size();
            for ( size_t j = 0; j < ck.handlers.size(); ++j )
                if ( ck.handlers[j] == h ) {
                    at = j;
                    break;
                }
            if ( at == ck.handlers.size() ) continue; // it removed itself: the next one has moved into slot i
            if ( stop ) {
                h->active = false;
                ck.handlers.erase(ck.handlers.begin() + (long)at);
                i = at;
            }
            else
                i = at + 1;
        }
        for ( size_t gi = 0; gi < ck.groups.size(); ) {
            ClockGroup* g = ck.groups[gi];
            for ( size_t j = 0; j < g->all.size(); ++j ) {
                ClockHandler* h = g->all[j];
                if ( !h->active ) continue;
                clock_ticks++;
                if ( comps[h->comp]->clockHandler(h->hid, ck.cycle) ) h->active = false;
            }
            if ( !g->anyActive() ) {
                g->in_list = false;
                ck.groups.erase(ck.groups.begin() + (long)gi);
            }
            else
                ++gi;
        }
        insert({ now + ck.period, PRIO_CLOCK << 32, 0, KIND_CLOCK, ck.index, 0 });
    }
    void primaryOK()
    {
        // exit.cc:53-68
        if ( --primary == 0 ) {
            end_time = now;
            insert({ now + 1, PRIO_EXIT << 32, 0, KIND_EXIT, 0, 0 });
        }
    }
    void run()
    {
        if ( stop_at ) insert({ stop_at, PRIO_STOP << 32, 0, KIND_STOP, 0, 0 });
        in_run = true;
        for ( auto c : comps )
            c->setup();
        while ( !end_sim && !pq.empty() ) {
            Activity a = pq.top();
            pq.pop();
            now      = a.time;
            cur_prio = a.prio_order >> 32;
            switch ( a.kind ) {
            case KIND_EVENT:
            {
                uint32_t   c = port_comp[a.target];
                Component* k = comps[c];
                events_delivered++;
                if ( tracing ) trace.push_back({ now, c, a.target - k->port0, a.payload });
                k->handleEvent((int)(a.target - k->port0), a.payload);
                break;
            }
            case KIND_CLOCK:
                executeClock(*clocks[a.target]);
                break;
            case KIND_STAT:
                comps[a.target]->statEvent((uint32_t)a.payload);
                break;
            case KIND_STOP:
                end_sim  = true;
                end_time = now;
                break;
            case KIND_EXIT:
                end_sim = true;
                break;
            }
        }
        now = end_time;
        for ( auto c : comps )
            c->endOfSimulation();
    }
};

void
Component::send(int port, uint64_t delay, uint64_t payload)
{
    uint32_t p    = port0 + (uint32_t)port;
    uint32_t peer = sim->peer_port[p];
    if ( peer == NO_PEER ) return;
    uint64_t t  = sim->now + delay + sim->latency[p];
    uint32_t dc = sim->port_comp[peer];
    uint32_t sq = sim->send_seq[id]++;
    if ( peer != p && (dc < sim->own_lo || dc >= sim->own_hi) ) {
        sim->outbox.insert(sim->outbox.end(), { t, ((uint64_t)peer << 32) | sq, payload, (uint64_t)dc });
        return;
    }
    sim->insert({ t, (PRIO_EVENT << 32) | sim->port_tag[peer], 0, KIND_EVENT, peer, payload });
}

void
Component::output(uint32_t code, uint64_t a, uint64_t b)
{
    sim->out_log.push_back({ sim->now, id, code, a, b });
}

// ------------------------------------------------------------------------------------------------ models
// testElements/message_mesh/enclosingComponent.cc
struct MeshComp : Component
{
    int                   my_id, mod;
    int32_t               message_count = 0;
    Mersenne              rng;
    Accumulator<uint64_t> mcnt;
    std::vector<uint32_t> counts, base; // links per MessagePort, flattened base index
    MeshComp(const int64_t* p) : my_id((int)p[0]), mod((int)p[1]), rng((uint32_t)(p[0] + 100))
    {
        int      ng = (int)p[4];
        uint32_t b  = 0;
        for ( int g = 0; g < ng; ++g ) {
            uint32_t c = (uint32_t)((p[5] >> (8 * g)) & 0xFF);
            counts.push_back(c);
            base.push_back(b);
            b += c;
        }
    }
    void setup() override
    {
        // RouteMessage::sendInitialEvents :188-199
        for ( size_t i = 0; i < counts.size(); i++ )
            for ( uint32_t j = 0; j < counts[i]; j++ )
                if ( (rng.u32() % (uint32_t)mod) == 0 ) send((int)(base[i] + j), 0, 0);
    }
    void handleEvent(int, uint64_t& payload) override
    {
        message_count++; // EnclosingComponent::handleEvent :92-99
        // RouteMessage::send :179-186
        int      nextport = (int)(rng.u32() % (uint32_t)counts.size());
        uint32_t portnum  = rng.u32() % counts[nextport];
        send((int)(base[nextport] + portnum), 0, payload);
        if ( sim->stats_on ) mcnt.add(1);
    }
    void results(uint64_t* r) const override
    {
        r[0] = (uint64_t)(int64_t)message_count;
        mcnt.out(r + 1);
    }
};

// oracle/probe/pdes_elements.cc PholdNode
struct PholdComp : Component
{
    int64_t               id_, init_;
    uint32_t              dmod;
    uint64_t              recv = 0, hash = 0;
    XorShift              rng;
    Accumulator<uint64_t> delay;
    PholdComp(const int64_t* p) : id_(p[0]), init_(p[1]), dmod((uint32_t)p[2]), rng((uint32_t)(p[0] + 1)) {}
    void setup() override
    {
        for ( int64_t i = 0; i < init_; ++i ) {
            uint32_t pp = rng.u32() % nports;
            uint64_t d  = rng.u32() % dmod;
            send((int)pp, d, d);
        }
    }
    void handleEvent(int port, uint64_t& payload) override
    {
        recv++;
        if ( sim->stats_on ) delay.add(payload);
        hash        = mixDelivery(hash, sim->now, (uint64_t)port, payload);
        uint32_t pp = rng.u32() % nports;
        uint64_t d  = rng.u32() % dmod;
        send((int)pp, d, d);
    }
    void results(uint64_t* r) const override
    {
        r[0] = recv;
        r[1] = hash;
        delay.out(r + 2);
    }
};

// oracle/probe/pdes_elements.cc ClockNode
struct ClockNodeComp : Component
{
    int32_t          commFreq;
    uint32_t         neighbor;
    uint64_t         ticks = 0, last_cycle = 0, recv = 0, hash = 0;
    Marsaglia        rng;
    Accumulator<int> count[4];
    ClockNodeComp(const int64_t* p) : commFreq((int32_t)p[1]), rng((uint32_t)(11 + p[0]), 272727u)
    {
        neighbor = rng.u32() % 4;
    }
    bool tick(uint64_t cycle) override
    {
        ticks++;
        last_cycle = cycle;
        if ( (rng.i32() % commFreq) == 0 ) {
            neighbor = (neighbor + 1) % 4;
            if ( neighbor < nports && sim->peer_port[port0 + neighbor] != NO_PEER ) {
                send((int)neighbor, 0, 0);
                if ( sim->stats_on ) count[neighbor].add(1);
            }
        }
        return false;
    }
    void handleEvent(int port, uint64_t&) override
    {
        recv++;
        hash = mixDelivery(hash, sim->now, (uint64_t)port, ticks);
    }
    void results(uint64_t* r) const override
    {
        r[0] = ticks;
        r[1] = last_cycle;
        r[2] = recv;
        r[3] = hash;
        for ( int i = 0; i < 4; ++i )
            count[i].out(r + 4 + 5 * i);
    }
};

// testElements/coreTest_Component.cc:25-173 (tests/test_Component.py): MarsagliaRNG(11, 272727) in every component;
// ctor :52-53 neighbor = generateNextInt32() % 4 (a C++ int: may start negative); clockTic :114-165 draws
// generateNextInt32() % commFreq, on 0 advances neighbor = (neighbor + 1) % 4 and sends on N/S/E/W = 0/1/2/3 with
// countX->addData(1), any other value prints "bad neighbor" and sends nothing; handleEvent :87-101 only touches the
// payload it then deletes.
struct CoreTestComp : Component
{
    int32_t          commFreq;
    int32_t          neighbor;
    uint64_t         bad = 0;
    Marsaglia        rng;
    Accumulator<int> count[4];
    CoreTestComp(const int64_t* p) : commFreq((int32_t)p[1]), rng(11u, 272727u) { neighbor = rng.i32() % 4; }
    bool tick(uint64_t) override
    {
        if ( (rng.i32() % commFreq) == 0 ) {
            neighbor = (neighbor + 1) % 4;
            if ( neighbor >= 0 ) {
                send(neighbor, 0, 0);
                if ( sim->stats_on ) count[neighbor].add(1);
            }
            else
                bad++;
        }
        return false;
    }
    void handleEvent(int, uint64_t&) override {}
    void results(uint64_t* r) const override
    {
        r[0] = bad;
        for ( int i = 0; i < 4; ++i )
            count[i].out(r + 4 + 5 * i);
    }
};

// testElements/coreTest_RNGComponent.cc:26-109
struct RngComp : Component
{
    Random*  rng;
    int64_t  max_count;
    int64_t  rng_count = 0;
    uint64_t last[5]   = { 0, 0, 0, 0, 0 };
    uint64_t hash      = 0;
    RngComp(const int64_t* p) : rng(makeRng((int)p[0], (uint64_t)p[1], (uint64_t)p[2])), max_count(p[3]) {}
    ~RngComp() { delete rng; }
    bool tick(uint64_t) override
    {
        double   nU  = rng->uniform();
        uint32_t U32 = rng->u32();
        uint64_t U64 = rng->u64();
        int32_t  I32 = rng->i32();
        int64_t  I64 = rng->i64();
        rng_count++;
        memcpy(&last[0], &nU, 8);
        last[1] = U32;
        last[2] = U64;
        last[3] = (uint64_t)(int64_t)I32;
        last[4] = (uint64_t)I64;
        for ( int i = 0; i < 5; ++i )
            hash = mixDelivery(hash, (uint64_t)rng_count, (uint64_t)i, last[i]);
        if ( rng_count == max_count ) {
            sim->primaryOK();
            return true;
        }
        return false;
    }
    void results(uint64_t* r) const override
    {
        r[0] = (uint64_t)rng_count;
        for ( int i = 0; i < 5; ++i )
            r[1 + i] = last[i];
        r[6] = hash;
    }
};

// testElements/coreTest_ClockerComponent.{h,cc} (tests/test_Clocks.py, tests/refFiles/test_Clocks_basic.out): a chain
// of components, each running the same schedule of start / stop operations on nine test clocks: old-style
// register/unregister/reregisterClock (0, 1), handler activate()/deactivate() (2, 3) and five ordered handlers (4..8, two
// of them registered by subcomponents) that operate on one variable in registration order.  The master clock (5 ns)
// issues the operations as events on a self link, so that they run after every clock of that time.
// Ports: 0 "left", 1 "right", 2 "inst_link" (configureSelfLink).  Handler ids: 0..8 the test clocks, 9 the master.
// Output record codes (a = id | second integer << 32, b = the 64-bit value):
//   0 "%d: starting test sequence at %llu"      1 "%d: Starting test sequence at %llu"
//   2 "%d: Clock %d will restart at cycle %llu" 3 "%d: Stopping Clock %d at time %llu"
//   4 "%d: Terminating test sequence at %llu"   5 "%d: Clock %d at cycle %llu"
//   6 "%d: Self stopping Clock %d at time %llu" 7 "%d: ordered clock result = %d (cycle = %llu)"
struct ClockerComp : Component
{
    static constexpr int      total_count   = 3;
    static constexpr uint64_t master_period = 5000, test_period = 1000; // "5ns", "1ns" at a 1 ps time base
    enum { OP_NOP, OP_START, OP_STOP, OP_TERM };
    struct OpBundle
    {
        int op, clock;
    };
    int           id_ = -1, inst_count_ = 0;
    bool          done_ = false;
    ClockHandler* master_ = nullptr;
    struct ClockInfo
    {
        ClockHandler* handler;
        int64_t       counter;
        bool          new_style;
    };
    std::vector<ClockInfo> clocks_;
    int                    ordered_var_test_       = -1;
    bool                   ordered_handler_remove_ = false;

    static const std::vector<std::vector<OpBundle>>& program()
    {
        // coreTest_ClockerComponent.h:265-276
        static const std::vector<std::vector<OpBundle>> p = {
            { { OP_START, 1 }, { OP_START, 5 }, { OP_START, 6 }, { OP_START, 7 }, { OP_START, 8 } },
            { { OP_START, 0 }, { OP_STOP, 6 } }, { { OP_START, 3 }, { OP_STOP, 7 } },
            { { OP_STOP, 1 }, { OP_START, 6 } },
            { { OP_START, 2 }, { OP_START, 4 } },
            { { OP_STOP, 3 }, { OP_START, 5 }, { OP_START, 6 }, { OP_START, 8 } },
            { { OP_START, 1 }, { OP_START, 6 }, { OP_START, 5 }, { OP_START, 8 }, { OP_START, 7 } },
            { { OP_START, 0 } }, { { OP_START, 3 } }, { { OP_STOP, 1 } }, { { OP_START, 2 } }, { { OP_STOP, 3 } },
            { { OP_TERM, -1 } }
        };
        return p;
    }
    // the constructor's clock registrations (coreTest_ClockerComponent.cc:75-165); called once ports are known
    void construct()
    {
        const bool left_connected = nports > 0 && sim->peer_port[port0 + 0] != NO_PEER;
        if ( !left_connected ) id_ = 0;
        master_ = sim->registerClock(id, master_period, 9);
        if ( id_ != 0 ) sim->deactivate(master_);
        for ( int i = 0; i < 4; ++i )
            clocks_.push_back({ sim->registerClock(id, test_period, i), total_count, i >= 2 });
        for ( int i = 0; i < 4; ++i )
            sim->deactivate(clocks_[i].handler); // unregisterClock x2, deactivate x2
        for ( int i = 4; i <= 8; ++i ) { // 6 and 7 come from the anonymous and the user subcomponent
            ClockHandler* h = sim->registerOrderedClock(id, test_period, i);
            sim->deactivate(h);
            clocks_.push_back({ h, total_count, true });
        }
    }
    uint64_t tag(int second) const { return (uint64_t)(uint32_t)id_ | ((uint64_t)(uint32_t)second << 32); }
    void     setup() override
    {
        if ( id_ != 0 ) return;
        output(0, tag(0), sim->now);
        send(1, 1 * master_period, 1); // IntEvent(1) on "right", delayed by one master period
    }
    void handleEvent(int port, uint64_t& payload) override
    {
        if ( port == 0 ) { // handle_left
            id_ = (int)payload;
            payload++;
            send(1, 0, payload);
            sim->activate(master_);
            output(1, tag(0), sim->now);
            return;
        }
        // inst_handler
        const int  op = (int)(payload & 0xFF), clock = (int)(int8_t)((payload >> 8) & 0xFF);
        switch ( op ) {
        case OP_START:
        {
            ClockInfo& info = clocks_[clock];
            uint64_t   next = sim->activate(info.handler); // activate() / reregisterClock()
            output(2, tag(clock), next);
            break;
        }
        case OP_STOP:
            sim->deactivate(clocks_[clock].handler); // deactivate() / unregisterClock()
            output(3, tag(clock), sim->now);
            break;
        case OP_TERM:
            sim->primaryOK();
            done_ = true;
            output(4, tag(0), sim->now);
            break;
        default:
            break;
        }
    }
    bool clockHandler(int hid, uint64_t cycle) override
    {
        if ( hid == 9 ) { // master_handler
            if ( done_ ) return true;
            for ( const OpBundle& x : program()[inst_count_++] )
                send(2, 0, (uint64_t)x.op | ((uint64_t)(uint8_t)x.clock << 8));
            return false;
        }
        if ( hid <= 3 ) { // test_handler
            output(5, tag(hid), cycle);
            if ( hid == 0 || hid == 2 ) {
                if ( --clocks_[hid].counter == 0 ) {
                    output(6, tag(hid), sim->now);
                    clocks_[hid].counter = total_count;
                    return true;
                }
            }
            return false;
        }
        // test_ordered_handler
        switch ( hid ) {
        case 4: ordered_handler_remove_ = true; break;
        case 5: ordered_var_test_ = 36; break;
        case 6: ordered_var_test_ /= 2; break;
        case 7: ordered_var_test_ -= 6; break;
        case 8: output(7, tag(ordered_var_test_), cycle); break;
        }
        return ordered_handler_remove_;
    }
    void results(uint64_t* r) const override
    {
        r[0] = (uint64_t)(int64_t)id_;
        r[1] = (uint64_t)inst_count_;
        r[2] = done_ ? 1 : 0;
        r[3] = (uint64_t)(int64_t)ordered_var_test_;
        for ( int i = 0; i < 9; ++i )
            r[4 + i] = (uint64_t)clocks_[i].counter;
    }
};

} // namespace

// ---- statistics with their engine (statapi/statbase.h:393-401 addData, statbase.cc:84-142 collection counts,
// statengine.cc:107-204 registration: "rate" in time or events, "startat", "stopat", "resetOnOutput"; :346-376,468-505
// periodic output and what follows an output; :296-311 the output at the end), restated per statistic.  A statistic
// slot is described by six int64 (the element's extra parameters, see sst_core_b200/elements.py):
//   [0] flags: 1 exists (not a NullStatistic), 2 resetOnOutput, 4 event-based rate   [1] rate: cycles, or events
//   [2] startat cycles   [3] stopat cycles   [4] index of the statistic object the slot adds to (shared objects)
//   [5] unused here (the host's output routing)
// Output record: code = slot | end_of_sim << 8, then Sum, SumSQ, Count, Min, Max as raw 64-bit patterns.
enum { ST_EXISTS = 1, ST_CLEAR = 2, ST_EVENT = 4 };
enum { SE_OUTPUT = 0, SE_START = 1, SE_STOP = 2 };
template <typename T>
uint64_t
bitsOf(T v)
{
    if constexpr ( std::is_same<T, float>::value ) {
        uint32_t u;
        memcpy(&u, &v, 4);
        return u;
    }
    else if constexpr ( std::is_same<T, double>::value ) {
        uint64_t u;
        memcpy(&u, &v, 8);
        return u;
    }
    else
        return (uint64_t)(int64_t)v;
}
struct StatSlotBase
{
    Component* comp = nullptr;
    uint32_t   slot = 0;
    int64_t    flags = 0, rate = 0, start = 0, stop = 0;
    bool       registered = false, enabled = true;
    uint64_t   count = 0, out_count = 0; // current_collection_count_, output_collection_count_
    virtual ~StatSlotBase() {}
    virtual void clearData()               = 0;
    virtual void fields(uint64_t* v) const = 0;
    void         configure(Component* c, uint32_t s, const int64_t* x)
    {
        comp = c, slot = s, flags = x[0], rate = x[1], start = x[2], stop = x[3];
    }
    // registerStatisticWithEngine at time `now` (0 in a constructor, later for a statistic registered in a handler)
    void registerNow()
    {
        if ( !(flags & ST_EXISTS) || registered ) return;
        registered   = true;
        Sim*     sim = comp->sim;
        uint64_t now = sim->now;
        if ( !(flags & ST_EVENT) && rate > 0 ) // the rate's statistic clock: next multiple of the rate (clock.cc:400-420)
            sim->insert({ (now / (uint64_t)rate + 1) * (uint64_t)rate, PRIO_STATCLOCK << 32, 0, KIND_STAT, comp->id,
                slot | (SE_OUTPUT << 8) });
        if ( start > 0 && (uint64_t)start > now ) { // setStatisticStartTime statengine.cc:395-422
            enabled = false;
            sim->insert({ (uint64_t)start, PRIO_STATCLOCK << 32, 0, KIND_STAT, comp->id, slot | (SE_START << 8) });
        }
        if ( stop > 0 && (uint64_t)stop > now ) // setStatisticStopTime :436-460
            sim->insert({ (uint64_t)stop, PRIO_STATCLOCK << 32, 0, KIND_STAT, comp->id, slot | (SE_STOP << 8) });
    }
    void output(bool end)
    {
        uint64_t v[5];
        fields(v);
        comp->output(slot | (end ? 0x100u : 0u), 0, 0); // placeholder replaced below
        OutRec& r = comp->sim->out_log.back();
        r.a = v[0], r.b = v[1];
        comp->sim->out_extra.push_back({ v[2], v[3], v[4] });
        if ( !end ) { // performStatisticOutputImpl statengine.cc:478-491
            if ( flags & ST_EVENT ) out_count = 0;
            if ( flags & ST_CLEAR ) {
                clearData();
                count = out_count = 0;
            }
        }
    }
    void event(uint32_t kind)
    {
        if ( kind == SE_START )
            enabled = true;
        else if ( kind == SE_STOP )
            enabled = false;
        else {
            output(false);
            comp->sim->insert({ comp->sim->now + (uint64_t)rate, PRIO_STATCLOCK << 32, 0, KIND_STAT, comp->id,
                slot | (SE_OUTPUT << 8) });
        }
    }
    void counted() // incrementCollectionCount + checkEventForOutput statbase.cc:84-90,134-142
    {
        count++;
        out_count++;
        if ( (flags & ST_EVENT) && rate >= 1 && out_count >= (uint64_t)rate ) output(false);
    }
};
template <typename T>
struct StatSlot : StatSlotBase
{
    T sum = 0, sumsq = 0, mn = std::numeric_limits<T>::max(), mx = std::numeric_limits<T>::min();
    void clearData() override
    {
        sum = 0, sumsq = 0, mn = std::numeric_limits<T>::max(), mx = std::numeric_limits<T>::min();
    }
    void addData(T v)
    {
        if ( !registered || !enabled ) return; // NullStatistic / disabled statistic
        sum += v;
        sumsq += (v * v);
        mn = (v < mn) ? v : mn;
        mx = (v > mx) ? v : mx;
        counted();
    }
    void fields(uint64_t* v) const override
    {
        v[0] = bitsOf(sum), v[1] = bitsOf(sumsq), v[2] = count;
        v[3] = count ? bitsOf(mn) : 0, v[4] = count ? bitsOf(mx) : 0;
    }
};

// testElements/coreTest_StatisticsComponent.cc:26-140 StatisticsComponentInt: a 1 ns clock draws u32, u64, i32, i64,
// scales them and adds them to stat1_U32 .. stat4_I64 (and to stat5_dyn, registered at tick `register_dynamic`)
struct StatsIntComp : Component
{
    Random*           rng;
    int64_t           rng_count = 0, max_count, dynamic_reg;
    StatSlot<uint32_t> s1;
    StatSlot<uint64_t> s2;
    StatSlot<int32_t>  s3;
    StatSlot<int64_t>  s4, s5;
    StatSlotBase*      slots[5];
    StatsIntComp(const int64_t* p) :
        rng(makeRng((int)p[0], (uint64_t)p[1], (uint64_t)p[2])), max_count(p[3]), dynamic_reg(p[4])
    {
        slots[0] = &s1, slots[1] = &s2, slots[2] = &s3, slots[3] = &s4, slots[4] = &s5;
    }
    ~StatsIntComp() { delete rng; }
    void construct(const int64_t* x, uint32_t nx)
    {
        for ( uint32_t i = 0; i < 5; ++i ) {
            static const int64_t none[6] = { 0, 0, 0, 0, 0, 0 };
            slots[i]->configure(this, i, i * 6 + 6 <= nx ? x + i * 6 : none);
            if ( i < 4 ) slots[i]->registerNow();
        }
        // (the 1 ns clock comes with the model's clock_period array)
    }
    bool tick(uint64_t) override
    {
        uint32_t U32 = rng->u32();
        uint64_t U64 = rng->u64();
        int32_t  I32 = rng->i32();
        int64_t  I64 = rng->i64();
        rng_count++;
        s1.addData(U32 / 10000000);
        s2.addData(U64 / 1000000000000000ull);
        s3.addData(I32 / 10000000);
        s4.addData(I64 / 1000000000000000ll);
        s5.addData(I64 / 1000000000000000ll);
        if ( rng_count == dynamic_reg ) s5.registerNow();
        if ( rng_count >= max_count ) {
            sim->primaryOK();
            return true;
        }
        return false;
    }
    void statEvent(uint32_t what) override { slots[what & 0xFF]->event(what >> 8); }
    void endOfSimulation() override
    {
        for ( auto* s : slots )
            if ( s->registered ) s->output(true);
    }
    void results(uint64_t* r) const override
    {
        r[0] = (uint64_t)rng_count;
        for ( int i = 0; i < 5; ++i )
            slots[i]->fields(r + 1 + 5 * i);
    }
};

// coreTest_StatisticsComponent.cc:160-260 StatisticsComponentFloat: two nextUniform() per tick; stat1_F32 gets
// (float)(a * 1000), stat2_F64 gets b * 1000, stat3_F64 gets b * 1000 + 10.  Slots may share a statistic object
// (createStatistic + setStatistic): slot parameter [4] names the object a slot adds to.
struct StatsFloatComp : Component
{
    Random*          rng;
    int64_t          rng_count = 0, max_count;
    StatSlot<float>  s1;
    StatSlot<double> s2, s3;
    StatSlotBase*    slots[3];
    int              target[3] = { 0, 1, 2 };
    StatsFloatComp(const int64_t* p) : rng(makeRng((int)p[0], (uint64_t)p[1], (uint64_t)p[2])), max_count(p[3])
    {
        slots[0] = &s1, slots[1] = &s2, slots[2] = &s3;
    }
    ~StatsFloatComp() { delete rng; }
    void construct(const int64_t* x, uint32_t nx)
    {
        for ( uint32_t i = 0; i < 3; ++i ) {
            static const int64_t none[6] = { 0, 0, 0, 0, 0, 0 };
            const int64_t*       c       = i * 6 + 6 <= nx ? x + i * 6 : none;
            target[i]                    = (c[0] & ST_EXISTS) ? (int)c[4] : (int)i;
            slots[i]->configure(this, i, c);
            if ( target[i] == (int)i ) slots[i]->registerNow();
        }
    }
    void addD(int slot, double v)
    {
        if ( target[slot] == 1 )
            s2.addData(v);
        else if ( target[slot] == 2 )
            s3.addData(v);
    }
    bool tick(uint64_t) override
    {
        double a = rng->uniform();
        double b = rng->uniform();
        rng_count++;
        s1.addData((float)(a * 1000));
        addD(1, b * 1000);
        addD(2, b * 1000 + 10.0);
        if ( rng_count >= max_count ) {
            sim->primaryOK();
            return true;
        }
        return false;
    }
    void statEvent(uint32_t what) override { slots[what & 0xFF]->event(what >> 8); }
    void endOfSimulation() override
    {
        for ( int i = 0; i < 3; ++i )
            if ( slots[i]->registered && target[i] == i ) slots[i]->output(true);
    }
    void results(uint64_t* r) const override
    {
        r[0] = (uint64_t)rng_count;
        for ( int i = 0; i < 3; ++i )
            slots[i]->fields(r + 1 + 5 * i);
    }
};

// ================================================================================================ C interface
extern "C" {

// Build a serial simulation from the raw wired-model arrays (same layout the product's C-ABI takes; see
// include/sstb200.h): component types + 8 int64 params each, directed-port CSR in configureLink order, peer port and
// send latency per directed port, clock period per component (0 = none), stop time (0 = none).
// port_self (may be NULL): 1 = a configureSelfLink port (baseComponent.cc:374-378, link.h:512-521): its peer is itself,
// its latency may be 0 and it keeps order tag 0xFFFFFFFF (Link::setTag link.h:309-319 leaves a SelfLink's tag alone),
// although the configureLink call still takes the next tag number.
// xparam_off / xparam (may be NULL): CSR of further int64 construction parameters per component.
void*
oracle_create3(uint32_t ncomp, const uint32_t* comp_type, const int64_t* comp_param, const uint32_t* port_off,
    const uint32_t* peer_port, const uint64_t* latency, const uint64_t* clock_period, const uint8_t* port_self,
    const uint32_t* xparam_off, const int64_t* xparam, uint64_t stop_at, int stats_on, int tracing)
{
    Sim* s      = new Sim();
    s->ncomp    = ncomp;
    s->stop_at  = stop_at;
    s->stats_on = stats_on != 0;
    s->tracing  = tracing != 0;
    s->port_off.assign(port_off, port_off + ncomp + 1);
    uint32_t nport = port_off[ncomp];
    s->peer_port.assign(peer_port, peer_port + nport);
    s->latency.assign(latency, latency + nport);
    s->port_comp.resize(nport);
    s->port_tag.resize(nport);
    s->send_seq.assign(ncomp, 0);
    uint32_t next_tag = 1; // simulation.h:566-572
    for ( uint32_t c = 0; c < ncomp; ++c ) {
        const int64_t* p = comp_param + (size_t)c * NPARAM;
        Component*     k = nullptr;
        switch ( comp_type[c] ) {
        case TYPE_MESH:
            k = new MeshComp(p);
            break;
        case TYPE_PHOLD:
            k = new PholdComp(p);
            break;
        case TYPE_CLOCKNODE:
            k = new ClockNodeComp(p);
            break;
        case TYPE_RNGCOMP:
            k = new RngComp(p);
            break;
        case TYPE_CORETEST:
            k = new CoreTestComp(p);
            break;
        case TYPE_CLOCKER:
            k = new ClockerComp();
            break;
        case TYPE_STATS_INT:
            k = new StatsIntComp(p);
            break;
        case TYPE_STATS_FLOAT:
            k = new StatsFloatComp(p);
            break;
        default:
            delete s;
            return nullptr;
        }
        k->sim    = s;
        k->id     = c;
        k->port0  = port_off[c];
        k->nports = port_off[c + 1] - port_off[c];
        s->comps.push_back(k);
        for ( uint32_t q = port_off[c]; q < port_off[c + 1]; ++q ) {
            s->port_comp[q] = c;
            // connected ports get the next tag in configureLink order.  (configureSelfLink links, which keep tag
            // 0xFFFFFFFF (link.cc:469-479), are not used by any in-scope model and are not representable here; a
            // config-graph loopback, peer == own port, is configured like any other link.)
            s->port_tag[q] = (peer_port[q] == NO_PEER) ? 0 : next_tag++;
            if ( port_self && port_self[q] ) s->port_tag[q] = 0xFFFFFFFFu;
        }
        if ( comp_type[c] == TYPE_CLOCKER ) static_cast<ClockerComp*>(k)->construct();
        const int64_t* xp = xparam_off ? xparam + xparam_off[c] : nullptr;
        const uint32_t nx = xparam_off ? xparam_off[c + 1] - xparam_off[c] : 0;
        if ( comp_type[c] == TYPE_STATS_INT ) static_cast<StatsIntComp*>(k)->construct(xp, nx);
        if ( comp_type[c] == TYPE_STATS_FLOAT ) static_cast<StatsFloatComp*>(k)->construct(xp, nx);
        // every in-scope model registers as primary and says DoNotEndSim in its constructor
        s->primary++;
        if ( clock_period[c] ) s->registerClock(c, clock_period[c]);
    }
    return s;
}

void*
oracle_create2(uint32_t ncomp, const uint32_t* comp_type, const int64_t* comp_param, const uint32_t* port_off,
    const uint32_t* peer_port, const uint64_t* latency, const uint64_t* clock_period, const uint8_t* port_self,
    uint64_t stop_at, int stats_on, int tracing)
{
    return oracle_create3(ncomp, comp_type, comp_param, port_off, peer_port, latency, clock_period, port_self, nullptr,
        nullptr, stop_at, stats_on, tracing);
}
void*
oracle_create(uint32_t ncomp, const uint32_t* comp_type, const int64_t* comp_param, const uint32_t* port_off,
    const uint32_t* peer_port, const uint64_t* latency, const uint64_t* clock_period, uint64_t stop_at, int stats_on,
    int tracing)
{
    return oracle_create2(ncomp, comp_type, comp_param, port_off, peer_port, latency, clock_period, nullptr, stop_at,
        stats_on, tracing);
}
// the third to fifth value of the records that have them, in record order (records with code >= ... carry five values
// when their component is a statistics element): extra[n_extra][3]
uint64_t
oracle_output_extra_count(void* h)
{
    return ((Sim*)h)->out_extra.size();
}
void
oracle_output_extra(void* h, uint64_t* extra)
{
    Sim* s = (Sim*)h;
    for ( size_t i = 0; i < s->out_extra.size(); ++i ) {
        extra[3 * i]     = s->out_extra[i].c;
        extra[3 * i + 1] = s->out_extra[i].d;
        extra[3 * i + 2] = s->out_extra[i].e;
    }
}

// console-output records (Component::output) in the order the serial run produced them
uint64_t
oracle_output_count(void* h)
{
    return ((Sim*)h)->out_log.size();
}
void
oracle_output(void* h, uint64_t* time, uint32_t* comp, uint32_t* code, uint64_t* a, uint64_t* b)
{
    Sim* s = (Sim*)h;
    for ( size_t i = 0; i < s->out_log.size(); ++i ) {
        time[i] = s->out_log[i].time;
        comp[i] = s->out_log[i].comp;
        code[i] = s->out_log[i].code;
        a[i]    = s->out_log[i].a;
        b[i]    = s->out_log[i].b;
    }
}

// ---- partitioned stepping: the reference's conservative window protocol (SyncManager::execute +
// RankSyncSerialSkip::exchange) restated for one partition.  Components outside [own_lo, own_hi) are constructed but
// never run (their setup() is skipped and nothing is delivered to them).
void
oracle_set_ownership(void* h, uint32_t own_lo, uint32_t own_hi)
{
    Sim* s    = (Sim*)h;
    s->own_lo = own_lo;
    s->own_hi = own_hi;
}

void
oracle_setup_owned(void* h)
{
    Sim* s = (Sim*)h;
    // non-owned components: remove their clock registrations and primary counts
    for ( auto* ck : s->clocks ) {
        std::vector<ClockHandler*> keep;
        for ( auto* h : ck->handlers )
            if ( h->comp >= s->own_lo && h->comp < s->own_hi ) keep.push_back(h);
        ck->handlers = keep;
        std::vector<ClockGroup*> keepg;
        for ( auto* g : ck->groups )
            if ( g->comp >= s->own_lo && g->comp < s->own_hi ) keepg.push_back(g);
        ck->groups = keepg;
    }
    s->primary = 0;
    for ( uint32_t c = s->own_lo; c < s->own_hi && c < s->ncomp; ++c )
        s->primary++;
    s->in_run = true;
    for ( uint32_t c = s->own_lo; c < s->own_hi && c < s->ncomp; ++c )
        s->comps[c]->setup();
}

// deliver every local activity with time < t_end (and, like StopAction, nothing at or after stop_at); returns the
// number of 4-word records now in the outbox
uint64_t
oracle_run_until(void* h, uint64_t t_end)
{
    Sim* s = (Sim*)h;
    while ( !s->pq.empty() ) {
        Activity a = s->pq.top();
        if ( a.time >= t_end ) break;
        s->pq.pop();
        s->now      = a.time;
        s->cur_prio = a.prio_order >> 32;
        if ( a.kind == KIND_EVENT ) {
            uint32_t   c = s->port_comp[a.target];
            Component* k = s->comps[c];
            s->events_delivered++;
            if ( s->tracing ) s->trace.push_back({ s->now, c, a.target - k->port0, a.payload });
            k->handleEvent((int)(a.target - k->port0), a.payload);
        }
        else if ( a.kind == KIND_CLOCK ) {
            s->executeClock(*s->clocks[a.target]);
        }
        else if ( a.kind == KIND_STAT ) {
            s->comps[a.target]->statEvent((uint32_t)a.payload);
        }
    }
    return s->outbox.size() / 4;
}

void
oracle_take_outbox(void* h, uint64_t* out)
{
    Sim* s = (Sim*)h;
    memcpy(out, s->outbox.data(), s->outbox.size() * 8);
    s->outbox.clear();
}

// re-insert events received from other partitions.  The tie-break among events with equal (time, tag) must be the
// sender's order: the records carry the sender's per-component sequence number, so insert them sorted by it.
void
oracle_ingest(void* h, const uint64_t* rec, uint64_t n)
{
    Sim*                  s = (Sim*)h;
    std::vector<uint64_t> idx(n);
    for ( uint64_t i = 0; i < n; ++i )
        idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) {
        return (int32_t)((uint32_t)rec[4 * a + 1] - (uint32_t)rec[4 * b + 1]) < 0;
    });
    for ( uint64_t j = 0; j < n; ++j ) {
        const uint64_t* r    = rec + 4 * idx[j];
        uint32_t        peer = (uint32_t)(r[1] >> 32);
        s->insert({ r[0], (PRIO_EVENT << 32) | s->port_tag[peer], 0, KIND_EVENT, peer, r[2] });
    }
}

// control words of this partition: [0] pending events, [1] primary components still running, [2] active clock
// handlers, [3] error, [4] exit time
void
oracle_control_words(void* h, int64_t* out)
{
    Sim*    s       = (Sim*)h;
    int64_t pending = 0, clocks = 0;
    // count queued events / active handlers
    auto copy = s->pq;
    while ( !copy.empty() ) {
        if ( copy.top().kind == KIND_EVENT ) pending++;
        copy.pop();
    }
    for ( auto* ck : s->clocks ) {
        clocks += (int64_t)ck->handlers.size();
        for ( auto* g : ck->groups )
            for ( auto* h : g->all )
                clocks += h->active ? 1 : 0;
    }
    out[0] = pending + (int64_t)(s->outbox.size() / 4);
    out[1] = s->primary;
    out[2] = clocks;
    out[3] = 0;
    out[4] = (int64_t)s->end_time;
    out[5] = out[6] = out[7] = 0;
}

void
oracle_destroy(void* h)
{
    delete (Sim*)h;
}

void
oracle_run(void* h)
{
    ((Sim*)h)->run();
}

// results[ncomp][32] u64, type-specific meaning (see include/sstb200.h "result record")
void
oracle_results(void* h, uint64_t* out)
{
    Sim* s = (Sim*)h;
    memset(out, 0, sizeof(uint64_t) * NRESULT * s->ncomp);
    for ( uint32_t c = 0; c < s->ncomp; ++c )
        s->comps[c]->results(out + (size_t)c * NRESULT);
}

// counters: [0] events delivered, [1] clock handler invocations, [2] end time, [3] trace length
void
oracle_counters(void* h, uint64_t* out)
{
    Sim* s = (Sim*)h;
    out[0] = s->events_delivered;
    out[1] = s->clock_ticks;
    out[2] = s->end_time;
    out[3] = s->trace.size();
}

// trace in global delivery order: time[], comp[], port[], payload[]
void
oracle_trace(void* h, uint64_t* time, uint32_t* comp, uint32_t* port, uint64_t* payload)
{
    Sim* s = (Sim*)h;
    for ( size_t i = 0; i < s->trace.size(); ++i ) {
        time[i]    = s->trace[i].time;
        comp[i]    = s->trace[i].comp;
        port[i]    = s->trace[i].port;
        payload[i] = s->trace[i].payload;
    }
}

// Raw generator stream: n draws of generateNextUInt32().  kind: 0 mersenne(seed=s1), 1 marsaglia(z=s1,w=s2),
// 2 xorshift(seed=s1).  `skip` draws are discarded first.  out may be NULL (checksum only).
// Returns the 64-bit checksum sum_i (i+1+skip)*draw_i mod 2^64 (position-sensitive).
uint64_t
oracle_rng_u32_stream(int kind, uint64_t s1, uint64_t s2, uint64_t skip, uint64_t n, uint32_t* out)
{
    Random* r = makeRng(kind, s1, s2);
    for ( uint64_t i = 0; i < skip; ++i )
        r->u32();
    uint64_t cs = 0;
    for ( uint64_t i = 0; i < n; ++i ) {
        uint32_t v = r->u32();
        if ( out ) out[i] = v;
        cs += (i + 1 + skip) * (uint64_t)v;
    }
    delete r;
    return cs;
}

// The coreTestRNGComponent tick sequence: per tick (nextUniform, u32, u64, i32, i64) -> 5 u64 words
// (double bit pattern, zero-extended u32, u64, sign-extended i32, i64).  Writes ticks [first, first+n).
void
oracle_rng_ticks(int kind, uint64_t s1, uint64_t s2, uint64_t first, uint64_t n, uint64_t* out)
{
    Random* r = makeRng(kind, s1, s2);
    for ( uint64_t t = 0; t < first + n; ++t ) {
        double   nU  = r->uniform();
        uint32_t U32 = r->u32();
        uint64_t U64 = r->u64();
        int32_t  I32 = r->i32();
        int64_t  I64 = r->i64();
        if ( t >= first ) {
            uint64_t* o = out + (t - first) * 5;
            memcpy(&o[0], &nU, 8);
            o[1] = U32;
            o[2] = U64;
            o[3] = (uint64_t)(int64_t)I32;
            o[4] = (uint64_t)I64;
        }
    }
    delete r;
}

// ------------------------------------------------------------------------------------------------ distributions
// kind of distribution: 0 discrete (rng/discrete.h:40-109), 1 exponential (expon.h:73-77), 2 gaussian
// (gaussian.h:81-110), 3 poisson (poisson.h:73-85), 4 uniform (uniform.h), 5 constant (constant.h).
// params: discrete: probs[np]; expon: lambda; gaussian: mean, stddev; poisson: lambda; uniform: bins; constant: v.
// The base generator is given by (rng kind, s1, s2).
void
oracle_distribution(int dist, const double* params, int np, int rng_kind, uint64_t s1, uint64_t s2, uint64_t n,
    double* out)
{
    Random* r = makeRng(rng_kind, s1, s2);
    switch ( dist ) {
    case 0:
    {
        // discrete.h:40-55: probabilities[i] = exclusive prefix sum; :96-109: first index with prob >= nextD wins,
        // scanning from index 0 upward; result is the index (as double)
        std::vector<double> prefix(np);
        double              sum = 0;
        for ( int i = 0; i < np; ++i ) {
            prefix[i] = sum;
            sum += params[i];
        }
        for ( uint64_t k = 0; k < n; ++k ) {
            const double nextD = r->uniform();
            uint32_t     index = 0;
            for ( ; index < (uint32_t)np; index++ )
                if ( prefix[index] >= nextD ) break;
            out[k] = (double)index;
        }
        break;
    }
    case 1:
        for ( uint64_t k = 0; k < n; ++k ) {
            const double next = r->uniform();
            out[k]            = log(1 - next) / (-1 * params[0]);
        }
        break;
    case 2:
    {
        bool   usePair    = false; // gaussian.h ctor: no cached value at start
        double unusedPair = 0;
        for ( uint64_t k = 0; k < n; ++k ) {
            if ( usePair ) {
                usePair = false;
                out[k]  = unusedPair;
            }
            else {
                double gauss_u, gauss_v, sq_sum;
                do {
                    gauss_u = r->uniform();
                    gauss_v = r->uniform();
                    sq_sum  = (gauss_u * gauss_u) + (gauss_v * gauss_v);
                } while ( sq_sum >= 1 || sq_sum == 0 );
                if ( r->uniform() < 0.5 ) gauss_u *= -1.0;
                if ( r->uniform() < 0.5 ) gauss_v *= -1.0;
                double multiplier = sqrt(-2.0 * log(sq_sum) / sq_sum);
                unusedPair        = params[0] + params[1] * gauss_v * multiplier;
                usePair           = true;
                out[k]            = params[0] + params[1] * gauss_u * multiplier;
            }
        }
        break;
    }
    case 3:
        for ( uint64_t k = 0; k < n; ++k ) {
            const double L = exp(-params[0]);
            double       p = 1.0;
            int          kk = 0;
            do {
                kk++;
                p *= r->uniform();
            } while ( p > L );
            out[k] = (double)(kk - 1);
        }
        break;
    case 4:
        for ( uint64_t k = 0; k < n; ++k ) {
            // uniform.h: walks bins: probPerBin = 1/bins; first bin whose cumulative probability exceeds the draw
            const uint32_t bins       = (uint32_t)params[0];
            const double   probPerBin = 1.0 / (double)bins;
            const double   nextD      = r->uniform();
            uint32_t       current_bin = 1;
            while ( nextD > ((double)current_bin * probPerBin) )
                current_bin++;
            out[k] = (double)(current_bin - 1);
        }
        break;
    default:
        for ( uint64_t k = 0; k < n; ++k )
            out[k] = params[0];
    }
    delete r;
}

} // extern "C"
