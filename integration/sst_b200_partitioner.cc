// Reference-side binding of libsstb200's partitioner: an ELI partitioner plugin for stock SST-core.
//
//     sst --partitioner=b200.linear ...        (library libb200.so on --lib-path)
//
// SSTPartitioner is the reference's plugin interface for partitioning (src/sst/core/sstpart.h:31-66, registered with
// SST_ELI_REGISTER_PARTITIONER :69-72 and created through Factory::CreatePartitioner, factory.cc).  The core hands the
// plugin the PartitionGraph -- the ConfigGraph with every group of components joined by no-cut links collapsed into
// one PartitionComponent (ConfigGraph::getCollapsedPartitionGraph, model/configGraph.cc:764-850) -- and expects a
// RankInfo{rank, thread} on every PartitionComponent.  This plugin gets the assignment from the C ABI
// (sstb200_partition_linear, include/sstb200.h), the same entry point the device engine's host driver uses, so that
// stock SST and the engine cut a model identically; oracle/build_ref.py builds it against the reference headers and
// tests/test_integration_shim.py runs the unmodified reference binary with it.
#include "sst_config.h"

#include "sst/core/sstpart.h"

#include "sst/core/eli/elementinfo.h"
#include "sst/core/model/configGraph.h"
#include "sst/core/output.h"
#include "sst/core/rankInfo.h"

#include "sstb200.h"

#include <cstdint>
#include <vector>

namespace B200 {

class LinearPartition : public SST::Partition::SSTPartitioner
{
public:
    SST_ELI_REGISTER_PARTITIONER(
        LinearPartition,
        "b200",
        "linear",
        SST_ELI_ELEMENT_VERSION(1,0,0),
        "sst.linear as computed by libsstb200 (sstb200_partition_linear): contiguous blocks of no-cut groups, "
        "thread-major within a rank")

    LinearPartition(SST::RankInfo total, SST::RankInfo /*my_rank*/, int /*verbosity*/) : total_(total) {}

    void performPartition(SST::PartitionGraph* graph) override
    {
        SST::PartitionComponentMap_t& comps = graph->getComponentMap();
        const uint32_t                n     = (uint32_t)graph->getNumComponents();
        const uint32_t                parts = total_.rank * total_.thread;
        std::vector<uint32_t>         part(n ? n : 1);
        // the groups are already collapsed: no pairs to pass
        if ( sstb200_partition_linear(n, nullptr, 0, parts, part.data()) != 0 )
            SST::Output::getDefaultObject().fatal(CALL_INFO, 1, "b200.linear: %s\n", sstb200_last_error());
        uint32_t i = 0;
        for ( auto it = comps.begin(); it != comps.end(); ++it, ++i )
            (*it)->rank = SST::RankInfo(part[i] / total_.thread, part[i] % total_.thread);
    }

    bool requiresConfigGraph() override { return false; }
    bool spawnOnAllRanks() override { return false; }

private:
    SST::RankInfo total_;
};

} // namespace B200
