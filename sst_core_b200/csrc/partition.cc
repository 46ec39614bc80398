// Host C++: the `sst.linear` partitioner over no-cut-collapsed groups.
// Reference behaviour restated (not copied): ConfigGraph::getCollapsedPartitionGraph model/configGraph.cc:764-850
// (components joined by no-cut links form one partition vertex; vertices are numbered in order of their first
// component id) and SSTLinearPartition::performPartition impl/partitioners/linpart.cc:30-82 (vertices are dealt out in
// contiguous runs of floor(G/P), the first G mod P partitions taking one extra; with fewer vertices than partitions
// each vertex gets its own partition).
#include "../../include/sstb200.h"

#include <cstdint>
#include <numeric>
#include <string>
#include <vector>

namespace sstb200 {
void set_error(const std::string& s);
}

namespace {
uint32_t
find_root(std::vector<uint32_t>& parent, uint32_t v)
{
    uint32_t r = v;
    while ( parent[r] != r )
        r = parent[r];
    while ( parent[v] != r ) {
        uint32_t n = parent[v];
        parent[v]  = r;
        v          = n;
    }
    return r;
}
} // namespace

extern "C" int
sstb200_partition_linear(uint32_t ncomp, const uint32_t* nocut_pairs, uint32_t n_pairs, uint32_t n_parts,
    uint32_t* rank_out)
{
    if ( n_parts == 0 || (!rank_out && ncomp) ) {
        sstb200::set_error("partition_linear: bad arguments");
        return -1;
    }
    std::vector<uint32_t> parent(ncomp);
    std::iota(parent.begin(), parent.end(), 0u);
    for ( uint32_t i = 0; i < n_pairs; ++i ) {
        uint32_t a = nocut_pairs[2 * i], b = nocut_pairs[2 * i + 1];
        if ( a >= ncomp || b >= ncomp ) {
            sstb200::set_error("partition_linear: no-cut pair references a component out of range");
            return -1;
        }
        uint32_t ra = find_root(parent, a), rb = find_root(parent, b);
        if ( ra != rb ) parent[ra > rb ? ra : rb] = (ra > rb ? rb : ra); // root = smallest id of the group
    }
    // group index in order of first component id
    std::vector<uint32_t> group_of_root(ncomp, 0xFFFFFFFFu), comp_group(ncomp);
    uint32_t              ngroups = 0;
    for ( uint32_t c = 0; c < ncomp; ++c ) {
        uint32_t r = find_root(parent, c);
        if ( group_of_root[r] == 0xFFFFFFFFu ) group_of_root[r] = ngroups++;
        comp_group[c] = group_of_root[r];
    }
    // deal groups out linearly
    std::vector<uint32_t> group_rank(ngroups);
    const uint64_t        per   = ngroups / n_parts;
    uint64_t              rem   = (per == 0) ? 0 : ngroups % n_parts;
    uint32_t              cur   = 0;
    uint64_t              onCur = 0;
    for ( uint32_t g = 0; g < ngroups; ++g ) {
        group_rank[g] = cur;
        onCur++;
        if ( onCur >= per ) {
            if ( rem > 0 && g + 1 < ngroups ) {
                --rem;
                ++g;
                group_rank[g] = cur;
            }
            cur++;
            onCur = 0;
        }
    }
    for ( uint32_t c = 0; c < ncomp; ++c )
        rank_out[c] = group_rank[comp_group[c]];
    return 0;
}
