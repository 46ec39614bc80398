/* sstb200.h -- C ABI of libsstb200.so, the B200-native event-delivery engine behind SST-core's plugin surface.
 *
 * The reference (sstsimulator/sst-core) has no C boundary for this path: the run loop, the TimeVortex, Link::send and
 * the rank/thread sync are C++ objects inside one binary.  Each entry point below names the reference interface it
 * replaces (paths relative to src/sst/core/).  A host driver -- sst_core_b200/engine.py here, or a C++ shim inside
 * stock SST -- builds the wired-model arrays from its ConfigGraph and calls these.
 * Plain pointers and sizes only; all arrays are caller-owned host memory unless a parameter is documented as a
 * device pointer.  Every function returns 0 on success or a negative error code; sstb200_last_error() gives the text
 * (the reference reports such errors through Output::fatal, output.cc:156-215).
 * There is no CPU fallback: creating an engine without a CUDA device fails.
 */
#ifndef SSTB200_H
#define SSTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSTB200_ABI_VERSION 2
#define SSTB200_NPARAM 8    /* int64 construction params per component                        */
#define SSTB200_NRESULT 32  /* u64 words in a component's result record                       */
#define SSTB200_NO_PEER 0xFFFFFFFFu

/* Result record (u64[32]) per component type, what finish() / the statistics engine would report
 * (Accumulator words are Sum, SumSQ, Count, Min, Max -- statapi/stataccumulator.h:171-194):
 *   1 coreTestElement.message_mesh.enclosing_component: [0] message_count (i32), [1..5] route "msg_count"
 *   2 pdes.phold:      [0] recv, [1] delivery hash, [2..6] "delay"
 *   3 pdes.clocknode:  [0] ticks, [1] last cycle, [2] recv, [3] delivery hash, [4+5i..] "N","S","E","W" (i32)
 *   4 coreTestElement.coreTestRNGComponent: [0] rng_count, [1..5] last tick (double bits,u32,u64,i32,i64), [6] hash
 *   6 coreTestElement.coreTestClockerComponent: [0] id, [1] instructions issued, [2] done, [3] ordered variable,
 *     [4..12] the nine clocks' counters
 */

typedef struct sstb200_engine sstb200_engine;

/* ---- library / ELI registry (eli/elibase.h:124-147 LoadedLibraries, factory.h:168-260; what sst-info prints) ---- */
int         sstb200_abi_version(void);
const char* sstb200_last_error(void);
int         sstb200_num_component_types(void);
/* idx in [0, num): ELI "lib.name", type code, POD state bytes, aux bytes per component, payload flag */
int         sstb200_component_type_info(int idx, const char** eli_name, uint32_t* type_code, uint32_t* state_bytes,
                                        uint32_t* aux_bytes, int* has_payload);
int         sstb200_component_type_lookup(const char* eli_name); /* type code or -1 */

/* ---- partitioning (impl/partitioners/linpart.cc:30-82 after ConfigGraph no-cut collapse, configGraph.cc:764-850)
 * nocut_pairs: n_pairs x 2 component ids joined by no-cut links.  rank_out[ncomp] receives the owning partition. */
int sstb200_partition_linear(uint32_t ncomp, const uint32_t* nocut_pairs, uint32_t n_pairs, uint32_t n_parts,
                             uint32_t* rank_out);

/* ---- engine (replaces Simulation::run simulation.cc:1072-1193, TimeVortexPQ impl/timevortex/timeVortexPQ.cc:62-95,
 *      Link::send_impl link.cc:623-658, Clock::execute clock.cc:314-397, SyncManager::execute
 *      sync/syncManager.cc:547-732 and RankSyncSerialSkip::exchange sync/rankSyncSerialSkip.cc:209-343) ---- */
typedef struct sstb200_model {
    uint32_t        ncomp;        /* GLOBAL number of components                                              */
    const uint32_t* comp_type;    /* [ncomp]   type codes                                                     */
    const int64_t*  comp_param;   /* [ncomp*8] construction params (host half of the element parses Params)   */
    const uint32_t* port_off;     /* [ncomp+1] CSR over directed ports, in configureLink order                */
    const uint32_t* peer_port;    /* [nport]   receiving directed port, SSTB200_NO_PEER if unconnected        */
    const uint64_t* latency;      /* [nport]   cycles added when sending FROM this port (own-end latency,
                                               simulation.cc:721-722)                                         */
    const uint64_t* clock_period; /* [ncomp]   cycles, 0 = no clock (registerClock, clock.cc:400-420)         */
    uint64_t        stop_at;      /* --stop-at in cycles, 0 = none (StopAction priority 30)                   */
    int             stats_enabled;
    const uint8_t*  port_self;    /* [nport] or NULL: 1 = a configureSelfLink port (baseComponent.cc:374-378,
                                               link.h:512-521): peer_port[p] == p, latency may be 0, its events are
                                               delivered after those of the component's other ports at equal time (the
                                               link keeps order tag 0xFFFFFFFF, link.h:309-319) -- self-link ports are
                                               the LAST ports of their component -- and it never limits the lookahead */
    const uint32_t* xparam_off;   /* [ncomp+1] or NULL: CSR over `xparam`, further int64 construction parameters of
                                               components whose Params do not fit the 8 fixed ones (the host half of
                                               the element defines their meaning)                               */
    const int64_t*  xparam;
} sstb200_model;

typedef struct sstb200_options {
    int      device;          /* CUDA device ordinal                                                         */
    uint32_t rank, n_ranks;   /* this engine's partition and the number of partitions (GPUs)                  */
    uint32_t comp_begin, comp_end; /* components [begin,end) are owned by this rank (contiguous after wire-up)  */
    const uint32_t* rank_comp_begin; /* [n_ranks+1] first component of every rank (NULL when n_ranks == 1)       */
    uint32_t calendar_slots;  /* W: ring of lookahead windows events are binned into (>= 2)                   */
    uint32_t bin_capacity;    /* events per (window slot, block of 256 components); 0 = default               */
    uint32_t smem_events;     /* due events per block staged in shared memory; 0 = default                    */
    uint32_t outbox_capacity; /* events per destination rank per window; 0 = default                          */
    uint64_t window;          /* lookahead window in cycles; 0 = min link latency (configGraph.cc:711-730)    */
    uint64_t trace_capacity;  /* delivery trace records to keep (0 = tracing off)                             */
    void*    stream;          /* cudaStream_t to run on (NULL = engine-owned stream)                          */
    uint32_t parallel_max_events; /* event-parallel form: most events of one component per window (0 = default 100;
                                 blocks with a busier component run the sequential form; lower it to test that) */
    uint32_t handler_counts;  /* non-zero: count deliveries per receiving port and clock-handler calls per component
                                 (sst.profile.handler.event.count / .clock.count, profile/eventHandlerProfileTool.cc,
                                 clockHandlerProfileTool.cc); read them with sstb200_engine_handler_counts          */
    uint32_t model_type_mask; /* multi-rank: bit t set = some component of the WHOLE model has type code t (the partitioner
                                 knows; the reference gets the same facts from its graph before Simulation exists).
                                 0 = unknown: engine_create scans all ncomp components itself                      */
    uint32_t model_uniform_ports; /* with model_type_mask: 1 = every component of the model has the same number of
                                 ports, 2 = not                                                                    */
    uint64_t output_capacity; /* console-output records to keep (Output::output calls of device components, one record
                                 each; 0 = default 65536 when an element of the model prints, else none)           */
} sstb200_options;

int  sstb200_engine_create(const sstb200_model* model, const sstb200_options* opt, sstb200_engine** out);
void sstb200_engine_destroy(sstb200_engine* e);

/* construct every component and run setup() (performWireUp simulation.cc:824-859, setup :969-987) */
int sstb200_engine_setup(sstb200_engine* e);

/* Run whole lookahead windows until the simulation ends or `max_windows` have been executed (0 = until the end).
 * Single-rank only (an n_ranks > 1 engine is stepped with the three calls below).  Asynchronous on the engine's
 * stream except for a completion check every `check_every` windows.  *finished is set to 1 when the simulation has
 * ended (stop time reached, primary components done, or nothing left to run). */
int sstb200_engine_run(sstb200_engine* e, uint64_t max_windows, uint32_t check_every, int* finished);
/* optional: capture the CUDA graph of a chunk of `check_every` windows now (nothing is launched), so that the first
 * sstb200_engine_run with the same check_every does not pay for the capture */
int sstb200_engine_prepare_run(sstb200_engine* e, uint32_t check_every);

/* Multi-rank stepping (one lookahead window = SyncManager::execute):
 *   dispatch  : deliver this window's events and clock ticks; cross-rank sends land in per-destination outboxes, and
 *               this rank's control words are written into every outbox header
 *   exchange  : the DRIVER moves outbox[r] of every rank to inbox[sender] of rank r (ONE all-to-all over NVLink/NCCL,
 *               CUDA-aware MPI, or device copies); buffers are DEVICE memory owned by the engine
 *   ingest    : bin received events into the local calendar; collect every rank's control words from the inbox headers
 *   advance   : fold the control words (termination, errors) and open the next window                          */
int sstb200_engine_dispatch(sstb200_engine* e);
int sstb200_engine_ingest(sstb200_engine* e);
int sstb200_engine_advance(sstb200_engine* e);
/* ---- peer mode: the exchange over NVLink without a collective (replaces RankSyncSerialSkip::exchange's MPI_Isend/
 * Irecv of serialized event vectors, sync/rankSyncSerialSkip.cc:239-300, and its two MPI_Allreduce, :318-335).
 * Every rank maps the others' calendars; a send to a remote component reserves its slot in the OWNER's bin with a
 * system-scope atomic and stores the record there (peer stores over NVLink); `dispatch` then ends with one mailbox
 * write per rank (control words + window sequence number), `ingest` has nothing to do, and `advance` waits for all
 * ranks' mailbox entries of the window before it folds them.  With the ranks connected, sstb200_engine_run drives a
 * multi-rank engine natively (CUDA-graph replay of whole windows) -- every rank calls it with the same arguments.
 *   peer_export : after engine_create; handle64 = 64-byte CUDA IPC handle of this rank's peer-visible region (may be
 *                 NULL), *nbins = its number of component blocks, *region_dev = the region's device pointer
 *   peer_connect: before engine_setup, on every rank, after ALL ranks have been created: handles = n_ranks x 64 bytes
 *                 and nbins[n_ranks] gathered from peer_export in rank order (the entry of this rank is ignored)
 *   peer arena  : cudaIpcOpenMemHandle costs 10-25 ms per peer, per allocation.  A process that runs many simulations
 *                 (or wants its first one to start fast) creates ONE arena per device up front -- like a communicator --
 *                 exchanges its handle once (peer_arena_create on every rank, all-gather, peer_arena_open), and engines
 *                 created afterwards place their peer-visible region inside it when it fits; peer_connect_arena then
 *                 connects the ranks without mapping anything.  *in_arena of peer_export says whether the engine did.
 *   peer_connect_local: the same for ranks living in ONE process on one GPU (regions = their device pointers);
 *                 the driver steps the ranks itself: dispatch on all, then advance on all */
int sstb200_engine_peer_export(sstb200_engine* e, void* handle64, uint32_t* nbins, void** region_dev);
int sstb200_engine_peer_connect(sstb200_engine* e, const void* handles, const uint32_t* nbins);
int sstb200_engine_peer_connect_local(sstb200_engine* e, void* const* regions, const uint32_t* nbins);
/* handler-time profile tools (profile/eventHandlerProfileTool.h:76-120 sst.profile.handler.event.time.*,
 * clockHandlerProfileTool.h sst.profile.handler.clock.time.*): nanoseconds of SM time spent inside the event handlers of
 * every receiving port [nport_local] and inside the clock handlers of every component [n_local] of this rank.  Needs
 * options.handler_counts & 2 (the counts above are collected too).  Either pointer may be NULL. */
int sstb200_engine_handler_times(sstb200_engine* e, double* event_recv_ns, double* clock_ns);
/* sync profile tools (profile/syncProfileTool.h:60-140 sst.profile.sync.count / .time.*): out4 = {rank syncs (windows of a
 * multi-rank run), events sent to other ranks, bytes of those records, nanoseconds this rank waited for its neighbour at the
 * window barrier (peer mode)} */
int sstb200_engine_sync_profile(sstb200_engine* e, uint64_t* out4);
/* console output of device components (Output::output, output.cc:156-215): records of 8 u64 words
 * {time, component | code << 32, v0, v1, v2, v3, v4, 0}; per component in the order it printed them.  *n_inout: capacity of
 * `records` in records on entry, number written on return (records == NULL: only the count).  The host half of the
 * element owns the format strings (sst_core_b200/elements.py). */
int sstb200_engine_output(sstb200_engine* e, uint64_t* n_inout, uint64_t* records);
/* bytes of the peer-visible region of an engine with n_local components (0 options = the engine's defaults) */
uint64_t sstb200_peer_region_bytes(uint32_t n_local, uint32_t calendar_slots, uint32_t bin_capacity, int has_payload);
int sstb200_peer_arena_create(int device, uint64_t bytes, void* handle64_out);
int sstb200_peer_arena_open(int device, uint32_t rank, uint32_t n_ranks, const void* handles);
int sstb200_peer_arena_destroy(int device);
int sstb200_engine_in_arena(sstb200_engine* e, int* in_arena);
int sstb200_engine_peer_connect_arena(sstb200_engine* e, const uint32_t* nbins);
/* the cudaStream_t the engine launches on (options.stream, or the one it created) */
int sstb200_engine_stream(sstb200_engine* e, void** stream_out);
/* which window kernel runs this rank's windows, decided at create from what the rank holds -- the analogue of asking
 * the reference which TimeVortex it loaded (simulation.cc:303-315): 0 the sequential form (dispatch_kernel), 1 the
 * event-parallel form (MessageMesh), 2 the ordered event-parallel form (PHOLD), 3 the clock form (clock-driven
 * populations).  Blocks a form cannot take in some window run in the sequential form in that window. */
int sstb200_engine_form(sstb200_engine* e, int* form_out);

/* device pointers + geometry of the exchange buffers.  Slot layout (u64 words), one slot per rank:
 *   [0] record count, [1..8] the sender's control words, [9] reserved, then capacity x {time, dst_seq, payload,
 *   dst_comp}; words_per_rank = 10 + capacity*4 */
int sstb200_engine_exchange_buffers(sstb200_engine* e, void** outbox_dev, void** inbox_dev, uint64_t* words_per_rank);
/* device pointers to this rank's 8 x i64 control words and to the n_ranks x 8 table `ingest` fills from the inbox
 * headers: [0] pending events (sent - delivered, may be negative on one rank), [1] primary components still running,
 * [2] active clocks, [3] error flag, [4] exit time, [5..7] reserved.  advance folds them: sums of 0..3, max of 4.
 * (A driver that cannot piggy-back may instead all-gather `local` into `global` itself after ingest.) */
int sstb200_engine_control_words(sstb200_engine* e, void** local_dev, void** global_dev);

/* ---- results (finish() simulation.cc:1286-1311, statistics dump statapi/statengine.cc:476-548) ---- */
typedef struct sstb200_status {
    uint64_t now;              /* start of the current window (cycles)                                        */
    uint64_t end_time;         /* simulated end time once finished                                            */
    uint64_t windows;          /* windows executed                                                            */
    uint64_t events_delivered; /* committed events on this rank                                               */
    uint64_t clock_ticks;      /* clock handler invocations on this rank                                      */
    uint64_t events_sent;
    uint64_t max_bin_fill;     /* high-water mark of any calendar bin (capacity tuning)                       */
    uint64_t max_due;          /* high-water mark of due events in one block                                  */
    uint32_t finished;
    uint32_t error;            /* 0 ok; 1 calendar bin overflow; 2 shared-memory staging overflow; 3 time fault;
                                  4 outbox overflow; 5 trace overflow; 6 a rank's mailbox entry did not arrive  */
    uint64_t error_info[4];
    uint64_t launches;         /* kernels launched by the engine so far                                       */
} sstb200_status;
int sstb200_engine_status(sstb200_engine* e, sstb200_status* out);
/* results[(comp_end-comp_begin) * words] u64 for the components this rank owns: the first `words_per_component` words
 * of each result record (0 = all 32) */
int sstb200_engine_results(sstb200_engine* e, uint64_t* results, uint32_t words_per_component);
/* delivery trace of this rank: per delivered event (time, component, port index within component, payload), in
 * per-component delivery order; rows of one component are contiguous in `comp`-sorted order after this call.   */
int sstb200_engine_trace(sstb200_engine* e, uint64_t* n_inout, uint64_t* time, uint32_t* comp, uint32_t* port,
                         uint64_t* payload);

/* ---- periodic statistic output (StatisticProcessingEngine::performStatisticOutput at a statistic clock,
 * statapi/statengine.cc:346-376; the clock runs at priority 85, i.e. AFTER the clock handlers and events of its time) ----
 * run_to_report: run whole windows up to and including the one that contains report_time and return every local
 *   component's result record as of the END of time report_time (layout as sstb200_engine_results).  *reported = 0 and
 *   nothing is run past the end when the simulation ends first or report_time >= stop_at (StopAction, priority 30, wins).
 * report_begin / report_end: the same for a driver that steps windows itself (multi-rank): bracket the window that
 *   contains report_time -- begin, dispatch [exchange, ingest], advance, end. */
int sstb200_engine_run_to_report(sstb200_engine* e, uint64_t report_time, uint64_t* results, uint32_t words_per_component,
                                 int* reported, int* finished);
int sstb200_engine_report_begin(sstb200_engine* e, uint64_t report_time);
int sstb200_engine_report_end(sstb200_engine* e, uint64_t* results, uint32_t words_per_component);

/* ---- checkpoint / restart (Simulation::checkpoint simulation.cc:2015-2071, restart :2073-2200; CheckpointAction) ----
 * The blob holds this rank's complete simulation state between two windows: engine control block, per-component
 * control blocks and element state (RNGs, statistics), the event calendar, exchange buffers, handler counters.
 * save:  after engine_setup, between windows (after engine_run returned / after engine_advance).
 * load:  into an engine created from the SAME model slice with the same calendar_slots, bin_capacity, window, ranks
 *        and handler_counts setting, INSTEAD of engine_setup; the run then continues bit-identically.
 * A restart on a different number of ranks needs the records re-binned and is not provided. */
int sstb200_engine_checkpoint_size(sstb200_engine* e, uint64_t* bytes);
int sstb200_engine_checkpoint_save(sstb200_engine* e, void* blob_host, uint64_t bytes);
int sstb200_engine_checkpoint_load(sstb200_engine* e, const void* blob_host, uint64_t bytes);

/* handler-count profile data of this rank (options.handler_counts): event_recv[local directed ports] = events
 * delivered to the handler of each port (EventHandlerProfileToolCount::beforeHandler, eventHandlerProfileTool.cc:96-100),
 * clock_calls[local components] = clock handler invocations (ClockHandlerProfileToolCount, clockHandlerProfileTool.cc:
 * 76-86).  Either pointer may be NULL. */
int sstb200_engine_handler_counts(sstb200_engine* e, uint64_t* event_recv, uint64_t* clock_calls);

/* ---- SST::RNG streams on the device (rng/mersenne.cc, marsaglia.cc, xorshift.cc; BASELINE config 2) ----
 * kind: 0 mersenne(seed=s1), 1 marsaglia(z=s1,w=s2), 2 xorshift(seed=s1).
 * n_streams independent generators, stream j seeded with (s1 + j*stride1, s2 + j*stride2), each producing
 * draws_per_stream u32 draws.  out_dev (DEVICE pointer, may be NULL) receives them stream-major
 * [n_streams][draws_per_stream]; checksum_dev (DEVICE, u64[n_streams]) receives sum_i (i+1)*draw_i per stream.   */
int sstb200_rng_streams(int device, void* stream, int kind, uint64_t s1, uint64_t s2, uint64_t stride1,
                        uint64_t stride2, uint64_t n_streams, uint64_t draws_per_stream, void* out_dev,
                        void* checksum_dev);
/* coreTestRNGComponent tick tuples: ticks [first, first+n) of one generator -> out_host[n*5] u64 */
int sstb200_rng_ticks(int device, int kind, uint64_t s1, uint64_t s2, uint64_t first, uint64_t n, uint64_t* out_host);
/* distributions (rng/discrete.h, expon.h, gaussian.h, poisson.h, uniform.h, constant.h): dist 0..5 as in
 * oracle_distribution; n values of one generator -> out_host[n] doubles */
int sstb200_rng_distribution(int device, int dist, const double* params, int n_params, int rng_kind, uint64_t s1,
                             uint64_t s2, uint64_t n, double* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SSTB200_H */
