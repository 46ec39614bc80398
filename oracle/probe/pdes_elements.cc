// TEST INFRASTRUCTURE ONLY -- out-of-tree SST element library "pdes" (libpdes.so), built against the
// UNMODIFIED reference headers by oracle/build_ref.py and loaded by the reference core (sstsim.x) so that the
// reference itself is the oracle for models it does not ship:
//
//   pdes.phold       PHOLD-style node: XORShiftRNG(id+1), integer-only delays,
//                    8-byte payload, order-sensitive delivery hash.
//   pdes.clocknode   clock-driven node in the style of coreTestComponent (core/testElements/coreTest_Component.cc:25-173):
//                    one clock handler, MarsagliaRNG(11+id, 272727), 1/commFreq chance per tick of sending a
//                    payload-less event to the next neighbour, 4 integer accumulators, delivery hash.
//   pdes.trace       EventHandlerProfileTool (core/profile/eventHandlerProfileTool.h:34-66) that logs one line per
//                    delivered event: "<time> <priority> <order tag> <queue order> <component:port>".
//
// The SAME models are restated in oracle/pdes_oracle.cc (CPU golden model) and implemented as device functors in
// sst_core_b200/csrc/components/*.cuh; parity tests compare all three.
#include "sst/core/sst_config.h"

#include "sst/core/component.h"
#include "sst/core/event.h"
#include "sst/core/link.h"
#include "sst/core/output.h"
#include "sst/core/params.h"
#include "sst/core/profile/eventHandlerProfileTool.h"
#include "sst/core/rng/marsaglia.h"
#include "sst/core/rng/xorshift.h"
#include "sst/core/statapi/statbase.h"

#include <cinttypes>
#include <cstdio>
#include <string>
#include <vector>

namespace PDESProbe {

// Order-sensitive mix of one delivery into a per-component 64-bit hash.  Any change in the per-component delivery
// sequence (time, port, payload) changes the final value.
static inline uint64_t
mixDelivery(uint64_t h, uint64_t time, uint64_t port, uint64_t payload)
{
    uint64_t v = time * 0x9E3779B97F4A7C15ull + (port + 1) * 0xC2B2AE3D27D4EB4Full + payload;
    h ^= v;
    h *= 0x100000001B3ull;
    h ^= h >> 29;
    return h;
}

class PholdEvent : public SST::Event
{
public:
    PholdEvent() : SST::Event(), d(0) {}
    explicit PholdEvent(uint64_t d) : SST::Event(), d(d) {}
    uint64_t d;
    void     serialize_order(SST::Core::Serialization::serializer& ser) override
    {
        Event::serialize_order(ser);
        SST_SER(d);
    }
    ImplementSerializable(PDESProbe::PholdEvent);
};

class PholdNode : public SST::Component
{
public:
    SST_ELI_REGISTER_COMPONENT(PholdNode, "pdes", "phold", SST_ELI_ELEMENT_VERSION(1, 0, 0),
        "PHOLD-style node with integer delays", COMPONENT_CATEGORY_UNCATEGORIZED)
    SST_ELI_DOCUMENT_PARAMS(
        { "id", "node id (seeds XORShiftRNG(id+1))", NULL },
        { "init_events", "events injected by setup()", "16" },
        { "delay_mod", "extra delay is u32 % delay_mod core cycles", "8000" },
        { "verbose", "print per-node summary in finish()", "1" })
    SST_ELI_DOCUMENT_PORTS({ "port%d", "links to neighbours", {} })
    SST_ELI_DOCUMENT_STATISTICS({ "delay", "delay carried by received events", "cycles", 1 })

    PholdNode(SST::ComponentId_t cid, SST::Params& params) : SST::Component(cid)
    {
        id_       = params.find<int64_t>("id", -1);
        init_     = params.find<int64_t>("init_events", 16);
        dmod_     = (uint32_t)params.find<int64_t>("delay_mod", 8000);
        verbose_  = params.find<int64_t>("verbose", 1);
        rng_      = new SST::RNG::XORShiftRNG((unsigned int)(id_ + 1));
        std::string name = "port0";
        while ( isPortConnected(name) ) {
            int idx = (int)links_.size();
            links_.push_back(configureLink(
                name, "1ps", new SST::Event::Handler<PholdNode, &PholdNode::handleEvent, int>(this, idx)));
            name = "port" + std::to_string(links_.size());
        }
        delay_ = registerStatistic<uint64_t>("delay");
        registerAsPrimaryComponent();
        primaryComponentDoNotEndSim();
    }

    void setup() override
    {
        for ( int64_t i = 0; i < init_; ++i ) {
            uint32_t p = rng_->generateNextUInt32() % (uint32_t)links_.size();
            uint64_t d = rng_->generateNextUInt32() % dmod_;
            links_[p]->send(d, new PholdEvent(d));
        }
    }

    void handleEvent(SST::Event* ev, int port)
    {
        // an event of another element (pdes.clocknode sends payload-less events) counts as delay 0 and is replaced by
        // a PHOLD event on its way on -- the device engine's events carry one 8-byte payload, 0 when the sender has none
        PholdEvent* pe = dynamic_cast<PholdEvent*>(ev);
        if ( !pe ) {
            delete ev;
            pe = new PholdEvent(0);
        }
        recv_++;
        delay_->addData(pe->d);
        hash_      = mixDelivery(hash_, getCurrentSimCycle(), (uint64_t)port, pe->d);
        uint32_t p = rng_->generateNextUInt32() % (uint32_t)links_.size();
        uint64_t d = rng_->generateNextUInt32() % dmod_;
        pe->d      = d;
        links_[p]->send(d, pe);
    }

    void finish() override
    {
        if ( verbose_ )
            SST::Output::getDefaultObject().output(
                "phold %" PRId64 " recv %" PRIu64 " hash %016" PRIx64 "\n", id_, recv_, hash_);
    }

private:
    int64_t                         id_, init_, verbose_;
    uint32_t                        dmod_;
    uint64_t                        recv_ = 0, hash_ = 0;
    SST::RNG::XORShiftRNG*          rng_;
    std::vector<SST::Link*>         links_;
    SST::Statistics::Statistic<uint64_t>*       delay_;
};

class ClockEvent : public SST::Event
{
public:
    void serialize_order(SST::Core::Serialization::serializer& ser) override { Event::serialize_order(ser); }
    ImplementSerializable(PDESProbe::ClockEvent);
};

class ClockNode : public SST::Component
{
public:
    SST_ELI_REGISTER_COMPONENT(ClockNode, "pdes", "clocknode", SST_ELI_ELEMENT_VERSION(1, 0, 0),
        "clock-driven node (coreTestComponent style)", COMPONENT_CATEGORY_UNCATEGORIZED)
    SST_ELI_DOCUMENT_PARAMS(
        { "id", "node id (seeds MarsagliaRNG(11+id, 272727))", NULL },
        { "clock", "clock frequency or period", "1GHz" },
        { "commFreq", "1/commFreq chance per tick of sending", "1000" },
        { "verbose", "print per-node summary in finish()", "1" })
    SST_ELI_DOCUMENT_PORTS({ "port%d", "links to neighbours", {} })
    SST_ELI_DOCUMENT_STATISTICS(
        { "N", "events sent on port0", "counts", 1 }, { "S", "events sent on port1", "counts", 1 },
        { "E", "events sent on port2", "counts", 1 }, { "W", "events sent on port3", "counts", 1 })

    ClockNode(SST::ComponentId_t cid, SST::Params& params) : SST::Component(cid)
    {
        id_       = params.find<int64_t>("id", -1);
        commFreq_ = (int32_t)params.find<int64_t>("commFreq", 1000);
        verbose_  = params.find<int64_t>("verbose", 1);
        rng_      = new SST::RNG::MarsagliaRNG((unsigned int)(11 + id_), 272727);
        neighbor_ = rng_->generateNextUInt32() % 4;
        const char* names[4] = { "port0", "port1", "port2", "port3" };
        const char* stats[4] = { "N", "S", "E", "W" };
        for ( int i = 0; i < 4; ++i ) {
            links_[i] = isPortConnected(names[i])
                            ? configureLink(names[i],
                                  new SST::Event::Handler<ClockNode, &ClockNode::handleEvent, int>(this, i))
                            : nullptr;
        }
        for ( int i = 0; i < 4; ++i )
            count_[i] = registerStatistic<int>(stats[i]);
        registerAsPrimaryComponent();
        primaryComponentDoNotEndSim();
        registerClock(params.find<std::string>("clock", "1GHz"),
            new SST::Clock::Handler<ClockNode, &ClockNode::tick>(this));
    }

    bool tick(SST::Cycle_t cycle)
    {
        ticks_++;
        last_cycle_ = cycle;
        if ( (rng_->generateNextInt32() % commFreq_) == 0 ) {
            neighbor_ = (neighbor_ + 1) % 4;
            if ( links_[neighbor_] ) {
                links_[neighbor_]->send(new ClockEvent());
                count_[neighbor_]->addData(1);
            }
        }
        return false;
    }

    void handleEvent(SST::Event* ev, int port)
    {
        recv_++;
        hash_ = mixDelivery(hash_, getCurrentSimCycle(), (uint64_t)port, ticks_);
        delete ev;
    }

    void finish() override
    {
        if ( verbose_ )
            SST::Output::getDefaultObject().output("clocknode %" PRId64 " ticks %" PRIu64 " last_cycle %" PRIu64
                                                   " recv %" PRIu64 " hash %016" PRIx64 "\n",
                id_, ticks_, last_cycle_, recv_, hash_);
    }

private:
    int64_t                   id_, verbose_;
    int32_t                   commFreq_;
    uint32_t                  neighbor_;
    uint64_t                  ticks_ = 0, last_cycle_ = 0, recv_ = 0, hash_ = 0;
    SST::RNG::MarsagliaRNG*   rng_;
    SST::Link*                links_[4];
    SST::Statistics::Statistic<int>*      count_[4];
};

// One line per delivered event, in the order the reference core delivers them.
class TraceTool : public SST::Profile::EventHandlerProfileTool
{
public:
    SST_ELI_REGISTER_PROFILETOOL(TraceTool, SST::Profile::EventHandlerProfileTool, "pdes", "trace",
        SST_ELI_ELEMENT_VERSION(0, 1, 0), "logs (time, priority, order tag, queue order, handler key) per delivery")
    SST_ELI_DOCUMENT_PARAMS({ "file", "output file", "pdes_trace.txt" })

    TraceTool(const std::string& name, SST::Params& params) : EventHandlerProfileTool(name, params)
    {
        std::string file = params.find<std::string>("file", "pdes_trace.txt");
        fp_              = fopen(file.c_str(), "w");
    }
    ~TraceTool()
    {
        if ( fp_ ) fclose(fp_);
    }
    uintptr_t registerHandler(const SST::AttachPointMetaData& mdata) override
    {
        keys_.push_back(getKeyForHandler(mdata));
        return (uintptr_t)(keys_.size() - 1);
    }
    uintptr_t registerLinkAttachTool(const SST::AttachPointMetaData& mdata) override
    {
        keys_.push_back(getKeyForHandler(mdata));
        return (uintptr_t)(keys_.size() - 1);
    }
    void beforeHandler(uintptr_t key, const SST::Event* ev) override
    {
        if ( !fp_ ) return;
        if ( ev )
            fprintf(fp_, "%" PRIu64 " %d %" PRIu32 " %" PRIu64 " %s\n", ev->getDeliveryTime(), ev->getPriority(),
                ev->getOrderTag(), ev->getQueueOrder(), keys_[key].c_str());
        else
            fprintf(fp_, "null 0 0 0 %s\n", keys_[key].c_str());
    }
    void outputData(SST::Util::DataRecord*, SST::RankInfo) override
    {
        if ( fp_ ) fflush(fp_);
    }

private:
    FILE*                    fp_ = nullptr;
    std::vector<std::string> keys_;
};

} // namespace PDESProbe
