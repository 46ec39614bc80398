/* A host driver written against include/sstb200.h alone (plain C, no Python, no SST headers): builds the MessageMesh
 * torus the way tests/models/mesh.py wires it, runs it on one GPU and prints what the reference's finish() prints.
 *
 *   gcc -O2 -Iinclude integration/c_abi_example.c -Lsst_core_b200 -lsstb200 -Wl,-rpath,$PWD/sst_core_b200 -o mesh_c
 *   ./mesh_c 6 6 10000000        ->  "0 received 39451 messages" ... (tests/refFiles/test_MessageMesh.out)
 *   ./mesh_c --layout            ->  sizes and field offsets of the ABI structs (tests/test_host.py checks them
 *                                    against the ctypes mirror in sst_core_b200/engine.py; needs no GPU)
 */
#include "sstb200.h"

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int
print_layout(void)
{
    printf("model %zu ncomp %zu comp_type %zu comp_param %zu port_off %zu peer_port %zu latency %zu clock_period %zu "
           "stop_at %zu stats_enabled %zu port_self %zu xparam_off %zu xparam %zu\n",
        sizeof(sstb200_model), offsetof(sstb200_model, ncomp), offsetof(sstb200_model, comp_type),
        offsetof(sstb200_model, comp_param), offsetof(sstb200_model, port_off), offsetof(sstb200_model, peer_port),
        offsetof(sstb200_model, latency), offsetof(sstb200_model, clock_period), offsetof(sstb200_model, stop_at),
        offsetof(sstb200_model, stats_enabled), offsetof(sstb200_model, port_self), offsetof(sstb200_model, xparam_off),
        offsetof(sstb200_model, xparam));
    printf("options %zu device %zu rank %zu n_ranks %zu comp_begin %zu comp_end %zu rank_comp_begin %zu calendar_slots %zu "
           "bin_capacity %zu smem_events %zu outbox_capacity %zu window %zu trace_capacity %zu stream %zu "
           "parallel_max_events %zu handler_counts %zu model_type_mask %zu model_uniform_ports %zu output_capacity %zu\n",
        sizeof(sstb200_options), offsetof(sstb200_options, device), offsetof(sstb200_options, rank),
        offsetof(sstb200_options, n_ranks), offsetof(sstb200_options, comp_begin), offsetof(sstb200_options, comp_end),
        offsetof(sstb200_options, rank_comp_begin), offsetof(sstb200_options, calendar_slots),
        offsetof(sstb200_options, bin_capacity), offsetof(sstb200_options, smem_events),
        offsetof(sstb200_options, outbox_capacity), offsetof(sstb200_options, window),
        offsetof(sstb200_options, trace_capacity), offsetof(sstb200_options, stream),
        offsetof(sstb200_options, parallel_max_events), offsetof(sstb200_options, handler_counts),
        offsetof(sstb200_options, model_type_mask), offsetof(sstb200_options, model_uniform_ports),
        offsetof(sstb200_options, output_capacity));
    printf("status %zu now %zu end_time %zu windows %zu events_delivered %zu clock_ticks %zu events_sent %zu "
           "max_bin_fill %zu max_due %zu finished %zu error %zu error_info %zu launches %zu\n",
        sizeof(sstb200_status), offsetof(sstb200_status, now), offsetof(sstb200_status, end_time),
        offsetof(sstb200_status, windows), offsetof(sstb200_status, events_delivered),
        offsetof(sstb200_status, clock_ticks), offsetof(sstb200_status, events_sent),
        offsetof(sstb200_status, max_bin_fill), offsetof(sstb200_status, max_due), offsetof(sstb200_status, finished),
        offsetof(sstb200_status, error), offsetof(sstb200_status, error_info), offsetof(sstb200_status, launches));
    return 0;
}

int
main(int argc, char** argv)
{
    if ( argc > 1 && strcmp(argv[1], "--layout") == 0 ) return print_layout();
    const uint32_t X = argc > 1 ? (uint32_t)atoi(argv[1]) : 6, Y = argc > 2 ? (uint32_t)atoi(argv[2]) : 6;
    const uint64_t stop = argc > 3 ? strtoull(argv[3], NULL, 10) : 10000000ull; /* cycles of 1 ps: 10 us */
    const uint32_t n = X * Y;
    const int      type = sstb200_component_type_lookup("coreTestElement.message_mesh.enclosing_component");
    if ( type < 0 ) {
        fprintf(stderr, "the library does not know the mesh element\n");
        return 1;
    }
    uint32_t* comp_type = malloc(n * sizeof *comp_type);
    int64_t*  param     = calloc((size_t)n * SSTB200_NPARAM, sizeof *param);
    uint32_t* port_off  = malloc((n + 1) * sizeof *port_off);
    uint32_t* peer      = malloc((size_t)n * 4 * sizeof *peer);
    uint64_t* latency   = malloc((size_t)n * 4 * sizeof *latency);
    uint64_t* clock     = calloc(n, sizeof *clock);
    for ( uint32_t i = 0; i < n; ++i ) {
        const uint32_t x = i % X, y = i / X;
        comp_type[i]     = (uint32_t)type;
        int64_t* p       = param + (size_t)i * SSTB200_NPARAM; /* id, mod, verbose, stats, port groups, links per group */
        p[0] = i, p[1] = 1, p[2] = 1, p[3] = 0, p[4] = 4, p[5] = 0x01010101;
        port_off[i] = 4 * i;
        /* ports in configureLink order x+, x-, y+, y-; a link joins port 0 with the neighbour's port 1, 2 with 3 */
        peer[4 * i + 0] = 4 * (((x + 1) % X) + y * X) + 1;
        peer[4 * i + 1] = 4 * (((x + X - 1) % X) + y * X) + 0;
        peer[4 * i + 2] = 4 * (x + ((y + 1) % Y) * X) + 3;
        peer[4 * i + 3] = 4 * (x + ((y + Y - 1) % Y) * X) + 2;
        for ( int k = 0; k < 4; ++k )
            latency[4 * i + k] = 1000; /* "1ns" */
    }
    port_off[n] = 4 * n;
    sstb200_model   m = { n, comp_type, param, port_off, peer, latency, clock, stop, 0, NULL, NULL, NULL };
    sstb200_options o;
    memset(&o, 0, sizeof o);
    o.n_ranks        = 1;
    o.calendar_slots = 2;
    sstb200_engine* e = NULL;
    int             finished = 0;
    if ( sstb200_engine_create(&m, &o, &e) || sstb200_engine_setup(e) || sstb200_engine_run(e, 0, 64, &finished) ) {
        fprintf(stderr, "FATAL: %s\n", sstb200_last_error());
        return 1;
    }
    uint64_t* res = malloc((size_t)n * sizeof *res);
    sstb200_status st;
    if ( sstb200_engine_results(e, res, 1) || sstb200_engine_status(e, &st) || st.error ) {
        fprintf(stderr, "FATAL: %s (engine error %u)\n", sstb200_last_error(), st.error);
        return 1;
    }
    for ( uint32_t i = 0; i < n; ++i )
        printf("%u received %d messages\n", i, (int)(int32_t)res[i]);
    printf("Simulation is complete, simulated time: %llu ps (%llu events, %llu windows)\n",
        (unsigned long long)st.end_time, (unsigned long long)st.events_delivered, (unsigned long long)st.windows);
    sstb200_engine_destroy(e);
    return 0;
}
