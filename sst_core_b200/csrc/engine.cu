// libsstb200: B200-native conservative parallel discrete-event engine (sm_100a) for SST-core's event-delivery path.
// See DESIGN.md for the data layout and the roofline of each kernel; include/sstb200.h for the C ABI.
//
// What replaces what (reference paths relative to src/sst/core/):
//   TimeVortexPQ (impl/timevortex/timeVortexPQ.cc:62-95) + Link::send_impl (link.cc:623-658)
//       -> a calendar of W lookahead-window slots x blocks of 256 components, structure-of-arrays in HBM.  A send is
//          the most-significant-digit pass of a radix sort on the 64-bit delivery key
//          (destination block | window slot) fused into the producing handler: one reservation + one record write.
//   Simulation::run pop/execute loop (simulation.cc:1099-1153) + Clock::execute (clock.cc:314-397)
//       -> dispatch_kernel: one CTA per block of 256 components loads the block's bin for the current window,
//          finishes the sort in shared memory (counting sort by component with warp-level scans, then the
//          (time, port order, sequence) order inside each component's segment -- the reference's
//          (delivery_time, priority<<32|order_tag, queue_order) order, activity.h:64-73), and every thread walks its
//          component's segment in order, merged with that component's clock ticks (clock priority 40 before event
//          priority 50 at equal time, activity.h:29-40), calling the device handler functors (components.cuh).
//   SyncManager::execute / RankSyncSerialSkip::exchange (sync/syncManager.cc:547-732, sync/rankSyncSerialSkip.cc:209-343)
//       -> per window: cross-rank sends go to per-destination outboxes, the driver moves them over NVLink (NCCL),
//          ingest_kernel bins them, advance_kernel folds the all-gathered control words (pending events, primary
//          components, active clocks, error, exit time) and opens the next window.
//   Exit / StopAction (exit.cc:53-132, stopAction.h:27-57) -> advance_kernel; the window is clipped at stop_at so that
//          nothing AT the stop time runs (StopAction priority 30 < clock 40 < event 50).
#include "../../include/sstb200.h"
#include "components.cuh"
#ifdef SSTB200_EXTRA_ELEMENTS_HEADER
#include SSTB200_EXTRA_ELEMENTS_HEADER // an element library compiled in (elements.def)
#endif

#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <type_traits>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace sstb200 {

static thread_local std::string g_error;
void
set_error(const std::string& s)
{
    g_error = s;
}

#define CUDA_OK(expr)                                                                                      \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if ( _e != cudaSuccess ) {                                                                         \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +             \
                      std::to_string(__LINE__) + ")");                                                     \
            return -2;                                                                                     \
        }                                                                                                  \
    } while ( 0 )

#ifndef SSTB200_BLOCK
#define SSTB200_BLOCK 256
#endif
constexpr int      BLOCK      = SSTB200_BLOCK; // components per CTA == threads per CTA
constexpr uint64_t TIME_NEVER = ~0ull;
constexpr uint32_t PAR_MAX_EVENTS = MeshElement::PARALLEL_MAX_EVENTS; // events of one component per window in the event-parallel form

enum : uint32_t { ERR_BIN_OVERFLOW = 1, ERR_SMEM_OVERFLOW = 2, ERR_TIME_FAULT = 3, ERR_OUTBOX_OVERFLOW = 4, ERR_TRACE_OVERFLOW = 5,
                  ERR_PEER_SYNC = 6, ERR_OUTPUT_OVERFLOW = 7, ERR_SELF_QUEUE = 8 };
enum { CW_PENDING = 0, CW_PRIMARY = 1, CW_CLOCKS = 2, CW_ERROR = 3, CW_EXIT_TIME = 4, CW_WORDS = 8 };
// exchange slot layout (u64 words): [0] record count, [1..8] the sender's control words (piggy-backed so that one
// all-to-all per window carries everything), [9] reserved, then capacity x {time, dst_seq, payload, dst_comp}
constexpr uint32_t XHDR = 10;
// peer mode (NVLink): every rank maps the calendars of the others and a send to a remote component is a reservation
// + record store in the OWNER's calendar; what is left of the window's exchange is one small mailbox write per rank.
// Mailbox slot (u64 words): [0..7] the sender's control words, [8] sequence number (windows completed + 1), [9] pad
constexpr uint32_t MAX_RANKS = 16;
constexpr uint32_t MAILW     = 10;

struct Ctrl
{
    uint64_t           k;     // window index; the window is [k*w, min((k+1)*w, stop_at))
    uint64_t           T, T_end;
    uint64_t           stop_at, w;
    unsigned long long events_delivered, clock_ticks, events_sent;
    unsigned long long max_bin_fill, max_due;
    uint64_t           end_time;
    uint32_t           done, error;
    unsigned long long error_info[4];
    long long          local[CW_WORDS]; // control words folded across ranks by advance_kernel
    unsigned long long trace_n;
    uint64_t           windows;
    uint32_t           had_primary, slot; // slot = k % W, kept by advance (a 64-bit remainder per thread otherwise)
    // periodic statistic output (statengine.cc:346-376): when report_at lies inside the current window every component
    // leaves its result record as of the END of time report_at (StatisticOutput clocks run at priority 85, after the
    // clock handlers and events of their time) in Dev::snapshot; TIME_NEVER otherwise
    uint64_t           report_at;
    // event-parallel form (parallel_form.cuh): blocks of this window left to the sequential form, components whose
    // generator needs maintenance after this window; both reset by advance_kernel
    uint32_t           fb_n, maint_n;
    unsigned long long out_n; // console-output records written (Dev::out_log)
    // sync profile tools (profile/syncProfileTool.h:60-140): events this rank sent to other ranks, nanoseconds this rank
    // waited for the others at the window barrier (peer mode)
    unsigned long long remote_events, sync_wait_ns;
    uint32_t           tail_done; // CTAs of maintenance_kernel that have finished (the last one ends the window)
    uint32_t           pf_next;   // blocks handed out by window_parallel_kernel beyond each CTA's first (reset per window)
    uint32_t           fb_taken, pad_fb; // entries of fb_list taken by CTAs of window_parallel_kernel that ran out of blocks
};
static_assert(sizeof(Ctrl) == 272, "Ctrl changed: keep _CTRL in sst_core_b200/checkpoint.py (repartitioned restart) in step");

// Per-component engine control block: everything the dispatch kernel needs about a component besides its element
// state, in one 32-byte record (two 16-byte loads; only the first half is ever written back).
struct CompCtl
{
    uint64_t next_tick; // CYCLE number of the next clock tick (its time = cycle x period: no division on the hot path),
                        // TIME_NEVER when the component has no (active) clock.  Elements with several clocks
                        // (MULTI_CLOCK) keep the TIME of their next clock activity here, with period 1
    uint32_t send_seq;  // per-component send sequence (link-FIFO tie-break)
    uint32_t deliv_seq; // deliveries so far (trace index)
    uint64_t period;    // clock period in cycles, 0 = none          -- constant
    uint32_t type;      // element type code                        -- constant
    uint32_t pad;
};
static_assert(sizeof(CompCtl) == 32, "CompCtl must be 32 bytes");

struct Dev
{
    Ctrl*           ctrl;
    uint32_t        n_local, c0, nbins, W, cap, emax;
    // divisions by the window length and remainders by the slot count without the divide unit: floor(x / w) =
    // mulhi64(x, w_magic) for x < 2^32 (w_magic = floor((2^64 - 1) / w) + 1); W_mask = W - 1 when W is a power of two
    uint64_t        w_magic;
    uint32_t        W_mask, pad_div;
    uint32_t        lmax, uniform_nports; // link rows staged per block; ports per component when uniform (else 0)
    uint32_t        uniform_shift;        // log2(uniform_nports) when it is a power of two, else 32
    uint32_t        uniform_type;         // the element type when every local component has the same one, else 0
    uint32_t        ties_free;            // every local element type declares events of equal time interchangeable
    uint32_t        par_max;              // most events of one component per window the event-parallel form takes
    // the per-component control blocks are neither read nor written by dispatch_kernel when nothing in them can
    // matter: no local component has a clock, one local element type, no tracing, and EVERY element type of the model
    // (any rank) treats events of equal time as interchangeable, so nobody ever looks at a send sequence number
    uint32_t        ctl_free;
    uint32_t        P0, nport_local;
    uint32_t        rank, n_ranks;
    const uint32_t* rank_comp_begin; // [n_ranks+1] (device)
    const uint32_t* comp_type; // [n_local] (setup / hooks; the hot path reads CompCtl)
    const int64_t*  comp_param;
    const uint32_t* port_off; // [n_local+1] GLOBAL directed-port ids
    const uint4*    link;     // [nport_local] {dst_port, dst_comp, latency lo, latency hi}
    const uint64_t* clock_period; // [n_local] (setup only)
    const uint32_t* xparam_off;   // [n_local+1] into xparam (GLOBAL offsets), null when the model has none (setup only)
    const int64_t*  xparam;       // this rank's slice, starting at xparam_off[0]
    CompCtl*        ctl;          // [n_local] per-component engine control block, 32 B
    uint8_t*        state;
    uint32_t        state_stride;
    uint8_t*        aux;
    uint32_t        aux_stride;
    ulonglong2*     ev;         // calendar records {time, dst_port<<32 | seq}, 16 B each
    uint64_t*       ev_payload; // only when an element of the model carries a payload
    uint32_t*       bin_count; // [W*nbins]
    unsigned long long* outbox;    // [n_ranks][1 + outbox_cap*4]
    unsigned long long* inbox;     // same shape, filled by the driver's exchange
    uint64_t        xwords;    // words per rank slot
    uint32_t        outbox_cap;
    long long*      gathered;  // [n_ranks][CW_WORDS]
    uint64_t*       tr_time;
    uint32_t*       tr_comp;
    uint32_t*       tr_port;
    uint32_t*       tr_idx;
    uint64_t*       tr_payload;
    uint64_t        trace_cap;
    // console output of device components (Output::output, output.cc): records of 8 words
    // {time, component | code << 32, v0..v4, 0}, in the order the owning threads wrote them
    unsigned long long* out_log;
    uint64_t        out_cap;
    int             stats_on, has_payload;
    // handler-count profile tools (profile/eventHandlerProfileTool.cc:108-116, clockHandlerProfileTool.cc:89-97):
    // deliveries per receiving port and clock-handler calls per component; null unless the option is set
    unsigned long long* prof_recv;  // [nport_local]
    unsigned long long* prof_clock; // [n_local]
    // handler-time profile tools (eventHandlerProfileTool.h:76-120, clockHandlerProfileTool.h): SM cycles spent inside
    // the event handlers of every receiving port / the clock handlers of every component; null unless asked for
    unsigned long long* prof_recv_cyc;  // [nport_local]
    unsigned long long* prof_clock_cyc; // [n_local]
    uint8_t*            snapshot;   // copy of `state` holding every component's state as of Ctrl::report_at
    uint8_t*            snapshot_aux; // ... and of `aux`, when a local element keeps its statistics there (STATS_IN_AUX)
    // event-parallel form: staging capacity for a bin's records, the two lists (see Ctrl)
    uint32_t            pf_rcap;
    uint32_t*           fb_list;    // [nbins]
    uint32_t*           maint_list; // [n_local] component | operation << 30
    // peer mode: the other ranks' calendars and mailboxes as mapped in this process (own entries = own buffers)
    uint32_t            peer_on;
    ulonglong2*         peer_ev[MAX_RANKS];
    uint64_t*           peer_pay[MAX_RANKS];
    uint32_t*           peer_bin_count[MAX_RANKS];
    unsigned long long* peer_mail[MAX_RANKS]; // [2][n_ranks][MAILW] of the owner
    uint32_t            peer_nbins[MAX_RANKS];
    unsigned long long* mail;                 // this rank's mailbox
};

// ================================================================================================ device side
__device__ __forceinline__ void
raise_error(const Dev& d, uint32_t code, uint64_t a, uint64_t b, uint64_t c)
{
    if ( atomicCAS(&d.ctrl->error, 0u, code) == 0u ) {
        d.ctrl->error_info[0] = a;
        d.ctrl->error_info[1] = b;
        d.ctrl->error_info[2] = c;
        d.ctrl->error_info[3] = d.ctrl->k;
    }
}

__device__ __forceinline__ uint32_t
div_w(const Dev& d, uint32_t x)
{
    return d.w_magic ? (uint32_t)__umul64hi((uint64_t)x, d.w_magic) : x; // w_magic 0: w == 1
}
__device__ __forceinline__ uint32_t
wrap_slot(const Dev& d, uint32_t x) // x mod W for x < 2W - 1 + W
{
    return d.W_mask ? (x & d.W_mask) : x % d.W;
}

// The fused "send": most-significant-digit scatter of the delivery key.  Local destination -> calendar bin of the
// (window slot, destination block); remote destination -> outbox of the owning rank.
__device__ __forceinline__ void
emit(const Dev& d, uint64_t k, uint64_t time, uint32_t dst_port, uint32_t dst_comp, uint32_t seq, uint64_t payload)
{
    uint32_t*   counts = d.bin_count;
    ulonglong2* evs    = d.ev;
    uint64_t*   pays   = d.ev_payload;
    uint32_t    nb = d.nbins, lc = dst_comp - d.c0;
    bool        remote = false;
    if ( d.n_ranks > 1 && (dst_comp < d.c0 || dst_comp >= d.c0 + d.n_local) ) {
        uint32_t r = 0;
        while ( r + 1 < d.n_ranks && dst_comp >= d.rank_comp_begin[r + 1] )
            ++r;
        atomicAdd(&d.ctrl->remote_events, 1ull);
        if ( !d.peer_on ) {
            unsigned long long* box = d.outbox + (uint64_t)r * d.xwords;
            unsigned long long  pos = atomicAdd(&box[0], 1ull);
            if ( pos >= d.outbox_cap ) {
                raise_error(d, ERR_OUTBOX_OVERFLOW, r, pos, d.outbox_cap);
                return;
            }
            unsigned long long* rec = box + XHDR + pos * 4;
            rec[0]                  = time;
            rec[1]                  = ((unsigned long long)dst_port << 32) | seq;
            rec[2]                  = payload;
            rec[3]                  = dst_comp;
            return;
        }
        // peer mode: straight into the owner's calendar (all ranks run the same window k, same slots, same capacity)
        counts = d.peer_bin_count[r];
        evs    = d.peer_ev[r];
        pays   = d.peer_pay[r];
        nb     = d.peer_nbins[r];
        lc     = dst_comp - d.rank_comp_begin[r];
        remote = true;
    }
    const uint32_t bin  = lc / BLOCK;
    const uint64_t base = k * d.ctrl->w;
    uint64_t       dk;
    const uint64_t dt = time - base;
    if ( time < base + d.ctrl->w ) {
        // would have to be delivered in the window that is being drained: violates the lookahead
        raise_error(d, ERR_TIME_FAULT, time, base, dst_comp);
        return;
    }
    if ( (dt >> 32) == 0 )
        dk = div_w(d, (uint32_t)dt);
    else
        dk = dt / d.ctrl->w;
    if ( dk > d.W - 1 ) dk = d.W - 1; // beyond the calendar horizon: parked in the farthest slot, re-binned there
    // (k == the window being run when called from a window kernel: its slot is in Ctrl; setup / ingest pass other k)
    const uint32_t slot = wrap_slot(d, (k == d.ctrl->k ? d.ctrl->slot : (uint32_t)(k % d.W)) + (uint32_t)dk);
    const uint32_t idx  = slot * nb + bin;
    const uint32_t pos  = remote ? atomicAdd_system(&counts[idx], 1u) : atomicAdd(&counts[idx], 1u);
    if ( pos >= d.cap ) {
        raise_error(d, ERR_BIN_OVERFLOW, slot, bin, pos);
        return;
    }
    const uint64_t off = (uint64_t)idx * d.cap + pos;
    evs[off] = make_ulonglong2(time, ((unsigned long long)dst_port << 32) | seq);
    if ( d.has_payload ) pays[off] = payload;
}

// ---- remote destinations in the block-wide flush of both window kernels.  Key of a remote destination bin:
// outbox mode 0x80000000 | rank; peer mode 0x80000000 | rank << 27 | index of the bin in the owner's calendar
__device__ __forceinline__ uint32_t
remote_key(const Dev& d, uint32_t dcomp, uint32_t sl)
{
    uint32_t r = 0;
    while ( r + 1 < d.n_ranks && dcomp >= d.rank_comp_begin[r + 1] )
        ++r;
    if ( !d.peer_on ) return 0x80000000u | r;
    return 0x80000000u | (r << 27) | (sl * d.peer_nbins[r] + (dcomp - d.rank_comp_begin[r]) / BLOCK);
}
__device__ __forceinline__ uint32_t
remote_reserve(const Dev& d, uint32_t ky, uint32_t cnt)
{
    if ( d.peer_on ) return atomicAdd_system(&d.peer_bin_count[(ky >> 27) & 15u][ky & 0x7FFFFFFu], cnt);
    return (uint32_t)atomicAdd(&d.outbox[(uint64_t)(ky & 0x7FFFFFFFu) * d.xwords], (unsigned long long)cnt);
}
__device__ __forceinline__ void
remote_write(const Dev& d, uint32_t ky, uint32_t pos, uint64_t tm, uint64_t ds, uint64_t payload, uint32_t dcomp)
{
    if ( d.peer_on ) {
        const uint32_t r = (ky >> 27) & 15u, idx = ky & 0x7FFFFFFu;
        if ( pos >= d.cap ) {
            raise_error(d, ERR_BIN_OVERFLOW, idx / d.peer_nbins[r], idx % d.peer_nbins[r], pos);
            return;
        }
        const uint64_t off = (uint64_t)idx * d.cap + pos;
        d.peer_ev[r][off]  = make_ulonglong2(tm, ds);
        if ( d.has_payload ) d.peer_pay[r][off] = payload;
        return;
    }
    const uint32_t r = ky & 0x7FFFFFFFu;
    if ( pos >= d.outbox_cap ) {
        raise_error(d, ERR_OUTBOX_OVERFLOW, r, pos, d.outbox_cap);
        return;
    }
    ulonglong2* rec = reinterpret_cast<ulonglong2*>(d.outbox + (uint64_t)r * d.xwords + XHDR + (uint64_t)pos * 4);
    rec[0]          = make_ulonglong2(tm, ds);
    rec[1]          = make_ulonglong2(payload, dcomp);
}

// Shared-memory view of one block's staged events.  Due events are staged here by phase 1; as a component consumes
// its events the slots are recycled for the events it sends (out-staging), and the owning warp flushes them with
// warp-aggregated reservations.  Record layout per slot (4 x u32 [+ u64 payload]):
//   staged-in : t = time - T, a = receiving port, sc = {sequence, comp<<16 | rank (sorting scratch)}
//   staged-out: t = arrival - T, a = receiving port, sc = {sequence, destination component}
struct Stage
{
    uint32_t* t;
    uint32_t* a;
    uint2*    sc; // {sequence, destination component}
    uint64_t* pay;
};

// What a handler sees of the core (BaseComponent / Link API subset)
template <bool STAGED>
struct CtxT
{
    const Dev&   d;
    uint64_t     now;   // getCurrentSimCycle()
    uint64_t     k;     // current window
    uint64_t     T;     // start of the current window
    uint32_t     comp;  // global component id
    uint32_t     port0; // global id of this component's first directed port
    uint32_t     nports;
    uint8_t*     aux;
    uint32_t     seq;   // per-component send sequence: the FIFO tie-break the reference gets from queue_order
    uint32_t     sent;
    int          stats_on;
    const uint4* links; // this component's link row (shared memory when the block's rows were staged, else global)
    // out-staging (STAGED only): this component's consumed slots are seg[free_head .. free_tail)
    Stage           st;
    const uint16_t* seg;
    uint32_t        free_head, free_tail;
    // elements with self links (SELF_LINKS): end of the window being run and the owning thread's queue of self events
    // that are due before it; null for every other element
    uint64_t        T_end  = 0;
    SelfQueue*      selfq  = nullptr;
    // constructors only: the component's further construction parameters (sstb200_model.xparam)
    const int64_t*  xparam  = nullptr;
    uint32_t        nxparam = 0;

    __device__ __forceinline__ bool connected(int port) const { return links[port].x != SSTB200_NO_PEER; }
    // getSimulationOutput().output(...): one record of the output log
    __device__ __forceinline__ void output(uint32_t code, uint64_t v0, uint64_t v1 = 0, uint64_t v2 = 0, uint64_t v3 = 0,
        uint64_t v4 = 0)
    {
        if ( !d.out_cap ) return;
        const unsigned long long pos = atomicAdd(&d.ctrl->out_n, 1ull);
        if ( pos >= d.out_cap ) {
            raise_error(d, ERR_OUTPUT_OVERFLOW, pos, d.out_cap, comp);
            return;
        }
        ulonglong2* r = reinterpret_cast<ulonglong2*>(d.out_log + pos * 8);
        r[0]          = make_ulonglong2(now, (unsigned long long)comp | ((unsigned long long)code << 32));
        r[1]          = make_ulonglong2(v0, v1);
        r[2]          = make_ulonglong2(v2, v3);
        r[3]          = make_ulonglong2(v4, 0);
    }
    // the component's record in memory (what of it is not part of State: MessageMesh route digest)
    __device__ __forceinline__ uint8_t* record() const { return d.state + (uint64_t)(comp - d.c0) * d.state_stride; }
    // Link::send(delay, ev): delivery time = now + delay + latency of the sending end (link.cc:636)
    __device__ __forceinline__ void send(int port, uint64_t delay, uint64_t payload)
    {
        const uint4 l = links[port];
        if ( l.x == SSTB200_NO_PEER ) return;
        const uint64_t lat  = ((uint64_t)l.w << 32) | l.z;
        const uint64_t when = now + delay + lat;
        const uint32_t sq   = seq++;
        sent++;
        if ( selfq != nullptr && l.x == port0 + (uint32_t)port && when < T + d.ctrl->w ) {
            // a self link, due inside this window: stays with the owning thread, ordered by (time, port, sequence)
            if ( when >= T_end ) return; // at or after the stop time: never delivered
            SelfQueue& q = *selfq;
            if ( q.n >= (uint32_t)SelfQueue::CAP ) {
                raise_error(d, ERR_SELF_QUEUE, comp, q.n, when);
                return;
            }
            int i = (int)q.n++;
            while ( i > 0 && (q.time[i - 1] > when || (q.time[i - 1] == when && q.port[i - 1] > l.x)) ) {
                q.time[i] = q.time[i - 1];
                q.port[i] = q.port[i - 1];
                q.seq[i]  = q.seq[i - 1];
                q.pay[i]  = q.pay[i - 1];
                --i;
            }
            q.time[i] = when;
            q.port[i] = l.x;
            q.seq[i]  = sq;
            q.pay[i]  = payload;
            return;
        }
        if ( STAGED ) {
            const uint64_t off = when - T;
            if ( free_head < free_tail && (off >> 32) == 0 ) {
                const uint32_t slot = seg[free_head++];
                st.t[slot]          = (uint32_t)off;
                st.a[slot]          = l.x;
                st.sc[slot]         = make_uint2(sq, l.y);
                if ( st.pay ) st.pay[slot] = payload;
                return;
            }
        }
        emit(d, k, when, l.x, l.y, sq, payload);
    }
    __device__ __forceinline__ void primaryComponentOKToEndSim()
    {
        atomicAdd((unsigned long long*)&d.ctrl->local[CW_PRIMARY], (unsigned long long)(-1ll));
        atomicMax((unsigned long long*)&d.ctrl->local[CW_EXIT_TIME], (unsigned long long)now);
    }
};

template <class F>
__device__ __forceinline__ void
with_element(uint32_t type, F&& f)
{
    switch ( type ) {
#define SSTB200_ELEMENT(E) \
    case E::TYPE:          \
        f(E {});           \
        break;
#include "elements.def"
#undef SSTB200_ELEMENT
    default:
        break;
    }
}

// optional element traits (absent = false)
template <class E, class = void>
struct multi_clock_of : std::false_type
{};
template <class E>
struct multi_clock_of<E, std::void_t<decltype(E::MULTI_CLOCK)>> : std::bool_constant<E::MULTI_CLOCK>
{};
template <class E, class = void>
struct own_kernel_of : std::false_type
{};
template <class E>
struct own_kernel_of<E, std::void_t<decltype(E::OWN_KERNEL)>> : std::bool_constant<E::OWN_KERNEL>
{};
template <class E, class = void>
struct prints_of : std::false_type
{};
template <class E>
struct prints_of<E, std::void_t<decltype(E::PRINTS)>> : std::bool_constant<E::PRINTS>
{};
template <class E, class = void>
struct stat_clock_of : std::false_type
{};
template <class E>
struct stat_clock_of<E, std::void_t<decltype(E::STAT_CLOCK)>> : std::bool_constant<E::STAT_CLOCK>
{};
template <class E, class = void>
struct self_links_of : std::false_type
{};
template <class E>
struct self_links_of<E, std::void_t<decltype(E::SELF_LINKS)>> : std::bool_constant<E::SELF_LINKS>
{};

// EU = the rank's only element type when it has just one (the kernel is then compiled for that element alone: no
// type switch, registers allocated for one handler), void otherwise
template <class EU, class F>
__device__ __forceinline__ void
with_element_of(uint32_t type, F&& f)
{
    if constexpr ( std::is_void<EU>::value )
        with_element(type, static_cast<F&&>(f));
    else
        f(EU {});
}

// Element state moves between HBM and registers in 16-byte pieces (state_stride is a multiple of 16).  `bytes` may be
// State layout convention: [0, STORE_END) mutable core (always loaded and stored), [STORE_END, CONST_END) construction
// constants (loaded, never stored), [CONST_END, STATS_END) statistics (loaded and stored only when statistics are
// enabled), then register-only scratch (neither).  All boundaries are multiples of 16.
template <class S>
__device__ __forceinline__ void
load_state(S& s, const uint8_t* p, uint32_t from, uint32_t to)
{
    static_assert(sizeof(S) % 8 == 0, "element State must be a multiple of 8 bytes");
    const uint4* src = reinterpret_cast<const uint4*>(p);
    // (only the pieces asked for are touched: a State is filled by one call per place its pieces live in)
#pragma unroll
    for ( int i = 0; i < (int)((sizeof(S) + 15) / 16); ++i )
        if ( (uint32_t)i * 16 >= from && (uint32_t)i * 16 < to ) {
            const uint4 t = src[i];
            memcpy(reinterpret_cast<uint8_t*>(&s) + (size_t)i * 16, &t, sizeof(S) - (size_t)i * 16 < 16 ? sizeof(S) - (size_t)i * 16 : 16);
        }
}
template <class S>
__device__ __forceinline__ void
store_state(const S& s, uint8_t* p, uint32_t from, uint32_t to)
{
    uint4* dst = reinterpret_cast<uint4*>(p);
    uint4  tmp[(sizeof(S) + 15) / 16];
    memcpy(tmp, &s, sizeof(S));
#pragma unroll
    for ( int i = 0; i < (int)((sizeof(S) + 15) / 16); ++i )
        if ( (uint32_t)i * 16 >= from && (uint32_t)i * 16 < to ) dst[i] = tmp[i];
}

// Where the statistics pieces [CONST_END, STATS_END) of component lc's State live: behind the rest of the record, or --
// elements that declare STATS_IN_AUX -- at the start of the component's aux block, so that the records of a
// clock-driven population are nothing but their hot heads, side by side (a window streams them instead of picking 64
// bytes out of every 160).  The pointer returned is the one load_state / store_state index with the State's own byte
// offsets: piece `off` is at base + off.
template <class E, class = void>
struct stats_in_aux_of : std::false_type
{};
template <class E>
struct stats_in_aux_of<E, std::void_t<decltype(E::STATS_IN_AUX)>> : std::bool_constant<E::STATS_IN_AUX>
{};
template <class E>
__device__ __forceinline__ uint8_t*
stats_base(uint8_t* state, uint32_t state_stride, uint8_t* aux, uint32_t aux_stride, uint32_t lc)
{
    if constexpr ( stats_in_aux_of<E>::value )
        return aux + (uint64_t)lc * aux_stride - E::CONST_END;
    else
        return state + (uint64_t)lc * state_stride;
}

// ---- MT19937 seeding (mersenne.cc:149-159) of the aux blocks of elements that declare AUX_MT_SEED.  The recurrence is
// serial per generator, so every lane runs its own component's; 64 words at a time go through a shared-memory tile and
// the warp writes them out one component per store instruction: 256 contiguous bytes (two full DRAM lines) each.
// (A thread per component storing its words one by one costs one 4-byte L2 transaction per word at a 2.5 KB stride --
// 88 ms for 16.7 M components; 64-byte pieces per component still left the 42 GB of stores at 1.2 TB/s.)
constexpr uint32_t SEED_TILE = 64;
__global__ void __launch_bounds__(128)
seed_mt_kernel(Dev d, uint32_t c_lo, uint32_t c_hi)
{
    __shared__ uint32_t tile[4][32][SEED_TILE + 1];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lc   = c_lo + blockIdx.x * 128 + threadIdx.x;
    bool           want = false;
    uint32_t       prev = 0, seed_off = 0;
    if ( lc < c_hi )
        with_element(d.comp_type[lc], [&](auto el) {
            using E = decltype(el);
            if constexpr ( E::AUX_MT_SEED ) {
                want     = true;
                prev     = E::auxSeed(d.comp_param + (uint64_t)lc * NPARAM);
                seed_off = E::AUX_SEED_OFFSET;
            }
        });
    const uint32_t wantmask = __ballot_sync(0xffffffffu, want);
    if ( !wantmask ) return;
    uint8_t* const warp_aux = d.aux + (uint64_t)(lc - lane) * d.aux_stride;
    for ( uint32_t chunk = 0; chunk < (MT_N + SEED_TILE - 1) / SEED_TILE; ++chunk ) {
#pragma unroll 8
        for ( uint32_t j = 0; j < SEED_TILE; ++j ) {
            const uint32_t i = chunk * SEED_TILE + j;
            if ( i ) prev = 1812433253u * (prev ^ (prev >> 30)) + i;
            tile[warp][lane][j] = prev;
        }
        __syncwarp();
        // words [chunk*64, chunk*64+64) of one component per iteration: lane l writes words 2l, 2l+1
        const uint32_t w0 = chunk * SEED_TILE + 2 * lane;
        for ( uint32_t comp = 0; comp < 32; ++comp ) {
            const uint32_t off = __shfl_sync(0xffffffffu, seed_off, comp);
            if ( (wantmask >> comp & 1u) && w0 < (uint32_t)MT_N )
                *(uint2*)(warp_aux + (uint64_t)comp * d.aux_stride + off + (uint64_t)w0 * 4) =
                    make_uint2(tile[warp][comp][2 * lane], tile[warp][comp][2 * lane + 1]);
        }
        __syncwarp();
    }
}

// ---- construct + setup(): performWireUp (simulation.cc:824-859) and setup() (simulation.cc:969-987)
__global__ void __launch_bounds__(BLOCK)
setup_kernel(Dev d)
{
    const uint32_t lc = blockIdx.x * BLOCK + threadIdx.x;
    if ( lc >= d.n_local ) return;
    const uint32_t type = d.comp_type[lc];
    const uint32_t p0   = d.port_off[lc];
    CtxT<false>    ctx { d, 0, 0, 0, d.c0 + lc, p0, d.port_off[lc + 1] - p0, d.aux + (uint64_t)lc * d.aux_stride,
                         0, 0, d.stats_on, d.link + (p0 - d.P0), Stage {}, nullptr, 0, 0 };
    const uint64_t period = d.clock_period[lc];
    uint64_t       first  = period ? 1 : TIME_NEVER; // first tick at cycle 1 = time `period`, never at 0 (clock.cc:400-420)
    uint64_t       period_ctl = period;
    uint32_t       always = 0;
    if ( d.xparam_off ) {
        ctx.xparam  = d.xparam + (d.xparam_off[lc] - d.xparam_off[0]);
        ctx.nxparam = d.xparam_off[lc + 1] - d.xparam_off[lc];
    }
    with_element(type, [&](auto el) {
        using E = decltype(el);
        typename E::State s;
        if constexpr ( stat_clock_of<E>::value ) always = 1; // visited every window: its statistics may be due
        E::construct(ctx, s, d.comp_param + (uint64_t)lc * NPARAM);
        E::setup(ctx, s);
        if constexpr ( multi_clock_of<E>::value ) { // the constructor registered its clocks
            first      = E::nextClock(s);
            period_ctl = 1;
        }
        store_state(s, d.state + (uint64_t)lc * d.state_stride, 0, E::CONST_END);
        store_state(s, stats_base<E>(d.state, d.state_stride, d.aux, d.aux_stride, lc), E::CONST_END, E::PERSIST_BYTES);
    });
    d.ctl[lc] = CompCtl { first, ctx.seq, 0, period_ctl, type, always };
    // counters: summed over the lanes that are here, ONE set of atomics per warp.  (One set per thread was 50 M atomics
    // on one cache line for a 16 M-component model -- and every send's loads of Ctrl::w / k / slot queued behind them:
    // 28 ms for what a window kernel does in one.)
    const uint32_t here  = __activemask();
    const uint32_t sent  = __reduce_add_sync(here, ctx.sent);
    const uint32_t clks  = __reduce_add_sync(here, first != TIME_NEVER ? 1u : 0u);
    const uint32_t prims = __popc(here); // every in-scope element is a primary component
    if ( (threadIdx.x & 31u) == (uint32_t)__ffs(here) - 1u ) {
        if ( sent ) {
            atomicAdd(&d.ctrl->events_sent, (unsigned long long)sent);
            atomicAdd((unsigned long long*)&d.ctrl->local[CW_PENDING], (unsigned long long)sent);
        }
        if ( clks ) atomicAdd((unsigned long long*)&d.ctrl->local[CW_CLOCKS], (unsigned long long)clks);
        atomicAdd((unsigned long long*)&d.ctrl->local[CW_PRIMARY], (unsigned long long)prims);
    }
}

// Exclusive scan of one u32 per thread over the CTA (warp shuffles + one shared round).
__device__ __forceinline__ uint32_t
block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t       inc = v;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ( lane >= (uint32_t)o ) inc += t;
    }
    if ( lane == 31 ) s_warp[warp] = inc;
    __syncthreads();
    if ( warp == 0 ) {
        uint32_t wv = (lane < BLOCK / 32) ? s_warp[lane] : 0;
        uint32_t wi = wv;
#pragma unroll
        for ( int o = 1; o < BLOCK / 32; o <<= 1 ) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ( lane >= (uint32_t)o ) wi += t;
        }
        if ( lane < BLOCK / 32 ) s_warp[lane] = wi - wv; // exclusive warp offsets
        if ( lane == BLOCK / 32 - 1 ) s_warp[BLOCK / 32] = wi;
    }
    __syncthreads();
    total = s_warp[BLOCK / 32];
    return s_warp[warp] + inc - v;
}

// wrap-safe "key of a < key of b" for two staged events of one component: (time, port, sequence)
__device__ __forceinline__ bool
key_less(uint32_t ta, uint32_t da, uint32_t qa, uint32_t tb, uint32_t db, uint32_t qb)
{
    return (ta < tb) || (ta == tb && (da < db || (da == db && (int32_t)(qa - qb) < 0)));
}

#ifdef SSTB200_TIMING
// phase timestamps (globaltimer, ns) of every CTA of the LAST window: [nbins][8]; profiles/phase_times.py reads them
__device__ unsigned long long g_phase_times[1 << 20];
__device__ __forceinline__ unsigned long long
gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define PHASE_MARK(i)                                                                    \
    do {                                                                                 \
        if ( threadIdx.x == 0 && blockIdx.x < (1 << 17) ) g_phase_times[blockIdx.x * 8 + (i)] = gtimer(); \
    } while ( 0 )
#else
#define PHASE_MARK(i)
#endif

// (time, port, sequence) order of one component's segment: insertion sort by the owning thread.  Used when no local
// element has an event-parallel form (the lanes of a warp own components with nearly equal event counts, and the sort
// overlaps the element's first loads); otherwise the block orders its events with the event-parallel rank sort.
__device__ __forceinline__ void
sort_segment(uint16_t* seg, uint32_t cnt, const uint32_t* s_time, const uint32_t* s_dst, const uint2* s_sc)
{
    for ( uint32_t a = 1; a < cnt; ++a ) {
        const uint16_t ea = seg[a];
        const uint32_t ta = s_time[ea], da = s_dst[ea], qa = s_sc[ea].x;
        uint32_t       b  = a;
        while ( b > 0 ) {
            const uint16_t eb = seg[b - 1];
            if ( !key_less(ta, da, qa, s_time[eb], s_dst[eb], s_sc[eb].x) ) break;
            seg[b] = eb;
            --b;
        }
        seg[b] = ea;
    }
}

// ---- one lookahead window for one block of 256 components
// (the sequential form: one thread per component walks the component's events and clock ticks in delivery order).
// REPORT: the launch of a window that contains a periodic-statistics report time (Ctrl::report_at).  A separate
// instantiation so that the ordinary launches carry none of it.
template <bool PAYLOAD, bool REPORT, class EU>
__device__ __forceinline__ void
dispatch_block(const Dev& d, const uint32_t bin, uint8_t* smem_raw)
{
    Ctrl* const    ctrl  = d.ctrl;
    const uint64_t k     = ctrl->k;
    const uint64_t T     = ctrl->T;
    const uint64_t T_end = ctrl->T_end;
    const uint32_t w     = (uint32_t)ctrl->w; // < 2^31 (checked by the host)
    const uint32_t slot  = ctrl->slot;
    const uint32_t slot1 = slot + 1 == d.W ? 0 : slot + 1; // where events due in the next window go
    const uint32_t tid   = threadIdx.x;
    const uint32_t lane  = tid & 31;

    PHASE_MARK(0);
    // shared-memory carve-up (host: engine_create computes the same sizes)
    uint4*    s_link     = reinterpret_cast<uint4*>(smem_raw);                          // [lmax]
    uint64_t* s_pay      = reinterpret_cast<uint64_t*>(smem_raw + (size_t)d.lmax * 16); // [emax] (PAYLOAD only)
    uint8_t*  p          = reinterpret_cast<uint8_t*>(s_pay) + (PAYLOAD ? (size_t)d.emax * 8 : 0);
    uint32_t* s_time     = reinterpret_cast<uint32_t*>(p);
    uint32_t* s_dst      = s_time + d.emax;
    uint2*    s_sc       = reinterpret_cast<uint2*>(s_dst + d.emax); // [emax] {sequence, comp<<16|rank / dst comp}
    uint32_t* s_port_off = reinterpret_cast<uint32_t*>(s_sc + d.emax); // [BLOCK+1]
    uint32_t* s_cnt      = s_port_off + BLOCK + 1;  // [BLOCK]
    uint32_t* s_off      = s_cnt + BLOCK;           // [BLOCK]
    uint32_t* s_warp     = s_off + BLOCK;           // [BLOCK/32+1]
    uint32_t* s_misc     = s_warp + BLOCK / 32 + 1; // [4]: 0 due events
    uint32_t* s_hist     = s_misc + 4;              // [32] components per event-count bucket, then bucket fill
    uint32_t* s_hbase    = s_hist + 32;             // [32]
    uint16_t* s_perm     = reinterpret_cast<uint16_t*>(s_hbase + 32); // [emax] per-component segments
    uint16_t* s_order    = s_perm + d.emax;                           // [BLOCK] thread -> component of the block

    // ---- prologue: the bin's records (as many as its fill count says: clock-driven populations have near-empty bins,
    // and loading 1024 slots blindly cost them 16 KB of DRAM reads per block) and everything else a block needs from
    // memory, issued back to back.
    const uint32_t n_in     = d.bin_count[slot * d.nbins + bin];
#ifndef SSTB200_SPEC
#define SSTB200_SPEC 4
#endif
    constexpr int SPEC = SSTB200_SPEC;
    ulonglong2    ev_spec[SPEC];
    uint64_t      pl_spec[SPEC];
#pragma unroll
    for ( int b = 0; b < SPEC; ++b ) {
        const uint32_t i = tid + b * BLOCK;
        pl_spec[b]       = 0;
        if ( i < n_in ) {
            ev_spec[b] = d.ev[(uint64_t)(slot * d.nbins + bin) * d.cap + i];
            if ( PAYLOAD ) pl_spec[b] = d.ev_payload[(uint64_t)(slot * d.nbins + bin) * d.cap + i];
        }
    }
    {
        const uint32_t lc0 = bin * BLOCK + tid;
        s_port_off[tid]    = d.port_off[lc0 <= d.n_local ? lc0 : d.n_local];
        if ( tid == 0 ) {
            const uint32_t e  = (bin + 1) * BLOCK;
            s_port_off[BLOCK] = d.port_off[e <= d.n_local ? e : d.n_local];
            s_misc[0]         = 0;
            // periodic statistic output falls into this window when T <= report_at < T_end: its offset, else all ones
            // (kept in shared memory: two more registers live across the whole kernel would spill)
            if constexpr ( REPORT ) {
                const uint64_t rep = ctrl->report_at;
                s_misc[1]          = (rep >= T && rep < T_end) ? (uint32_t)(rep - T) : 0xFFFFFFFFu;
            }
        }
        s_cnt[tid] = 0;
        if ( tid < 32 ) s_hist[tid] = 0;
        // element hook: start pulling what the handlers will touch towards L2
        if ( lc0 < d.n_local )
            with_element_of<EU>(d.uniform_type ? d.uniform_type : d.comp_type[lc0], [&](auto el) {
                using E = decltype(el);
                E::prefetch(d.state + (uint64_t)lc0 * d.state_stride, d.aux + (uint64_t)lc0 * d.aux_stride);
            });
    }
    __syncthreads();
    PHASE_MARK(1);
    const uint32_t blk_port0     = s_port_off[0];
    const uint32_t blk_nports    = s_port_off[BLOCK] - blk_port0;
    // stage the block's link rows in shared memory when the block has a fair number of events to forward; a block that
    // only ticks (clock populations send once in a thousand ticks) reads the odd row it needs from memory instead
    const bool     links_in_smem = blk_nports <= d.lmax && n_in >= BLOCK / 8;
    if ( links_in_smem ) { // asynchronous copy (LDGSTS): nobody waits for the rows until phase 3
        const uint4* src = d.link + (blk_port0 - d.P0);
        for ( uint32_t i = tid; i < blk_nports; i += BLOCK )
            __pipeline_memcpy_async(&s_link[i], &src[i], 16);
    }
    __pipeline_commit();

    // ---- phase 1: load the bin; stage due events in shared memory, re-bin the ones that are not due yet
    uint32_t       dropped  = 0;
    const uint64_t bin_base = (uint64_t)(slot * d.nbins + bin) * d.cap;
    auto           classify = [&](uint64_t t, uint64_t ds, uint64_t pl) {
        const uint32_t dst = (uint32_t)(ds >> 32);
        // component of the receiving port: largest j with s_port_off[j] <= dst
        uint32_t lo;
        if ( d.uniform_shift < 32 ) {
            lo = (dst - blk_port0) >> d.uniform_shift;
        }
        else if ( d.uniform_nports ) {
            lo = (dst - blk_port0) / d.uniform_nports;
        }
        else {
            lo          = 0;
            uint32_t hi = BLOCK;
            while ( hi - lo > 1 ) {
                const uint32_t mid = (lo + hi) >> 1;
                if ( s_port_off[mid] <= dst )
                    lo = mid;
                else
                    hi = mid;
            }
        }
        if ( t < T ) {
            raise_error(d, ERR_TIME_FAULT, t, T, bin * BLOCK + lo + d.c0);
        }
        else if ( t >= T_end ) {
            if ( t < T + w ) {
                dropped++; // at or after a stop time inside this window: never delivered (StopAction priority 30)
            }
            else {
                emit(d, k, t, dst, d.c0 + bin * BLOCK + lo, (uint32_t)ds, pl);
            }
        }
        else {
            const uint32_t e = atomicAdd(&s_misc[0], 1u);
            if ( d.prof_recv ) atomicAdd(&d.prof_recv[dst - d.P0], 1ull);
            if ( e < d.emax ) {
                const uint32_t r = atomicAdd(&s_cnt[lo], 1u);
                s_time[e]        = (uint32_t)(t - T);
                s_dst[e]         = dst;
                s_sc[e]          = make_uint2((uint32_t)ds, (lo << 16) | (r & 0xFFFFu));
                if ( PAYLOAD ) s_pay[e] = pl;
            }
        }
    };
#pragma unroll
    for ( int b = 0; b < SPEC; ++b )
        if ( tid + b * BLOCK < n_in ) classify(ev_spec[b].x, ev_spec[b].y, pl_spec[b]);
    for ( uint32_t i = tid + SPEC * BLOCK; i < n_in; i += BLOCK ) {
        const ulonglong2 ev = d.ev[bin_base + i];
        uint64_t         pl = 0;
        if ( PAYLOAD ) pl = d.ev_payload[bin_base + i];
        classify(ev.x, ev.y, pl);
    }
    __pipeline_wait_prior(0);
    __syncthreads();
    PHASE_MARK(2);
    const uint32_t n_due = s_misc[0];
    if ( n_due > d.emax ) {
        if ( tid == 0 ) raise_error(d, ERR_SMEM_OVERFLOW, bin, n_due, d.emax);
        return; // uniform across the CTA
    }

    // ---- phase 2: finish the sort in shared memory (skipped by a block without due events: a whole-population clock
    // tick needs none of it, and its sends go straight to the calendar)
    uint32_t c = tid;
    if ( n_due ) {
    // (a) counting sort by component: exclusive scan of the per-component counts (warp shuffles), scatter
    {
        const uint32_t cnt_c = s_cnt[tid];
        uint32_t       total;
        s_off[tid] = block_exclusive_scan(cnt_c, s_warp, total);
        // (c, first half) bucket the block's components by how many events they got, busiest first, so that the
        // 32 components a warp runs in phase 3 have (nearly) the same trip count
        atomicAdd(&s_hist[31 - (cnt_c < 31 ? cnt_c : 31)], 1u);
    }
    __syncthreads();
    for ( uint32_t e = tid; e < n_due; e += BLOCK ) {
        const uint32_t cr  = s_sc[e].y;
        const uint32_t pos = s_off[cr >> 16] + (cr & 0xFFFFu);
        s_perm[pos]        = (uint16_t)e;
    }
    if ( tid < 32 ) { // exclusive scan of the 32 buckets by one warp
        const uint32_t v   = s_hist[tid];
        uint32_t       inc = v;
#pragma unroll
        for ( int o = 1; o < 32; o <<= 1 ) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if ( lane >= (uint32_t)o ) inc += t;
        }
        s_hbase[tid] = inc - v;
        s_hist[tid]  = 0; // reused as the bucket fill counter
    }
    __syncthreads();
    // (b) the order inside each component's segment, (time, port, sequence), is made by the owning thread in phase 3
    // (insertion sort: the lanes of a warp own components with nearly equal event counts, and the sort overlaps the
    // element's first loads)
    // (c, second half) thread -> component assignment
    {
        const uint32_t cnt_c  = s_cnt[tid];
        const uint32_t bucket = 31 - (cnt_c < 31 ? cnt_c : 31);
        s_order[s_hbase[bucket] + atomicAdd(&s_hist[bucket], 1u)] = (uint16_t)tid;
    }
    __syncthreads();
    c = s_order[tid];
    }
    const uint16_t* s_ordered = s_perm;

    PHASE_MARK(3);
    // ---- phase 3: run the handlers.  Thread t owns component s_order[t] (the t-th busiest of the block).
    const uint32_t  lc     = bin * BLOCK + c;
    const bool      valid  = lc < d.n_local;
    const uint32_t  my_cnt = n_due ? s_cnt[c] : 0;
    const uint16_t* seg    = s_ordered + (n_due ? s_off[c] : 0);
    uint32_t        delivered = 0, ticks = 0, sent = 0;
    uint64_t        next_store = TIME_NEVER; // what goes back into CompCtl::next_tick
    int32_t         clock_delta = 0; // components whose clocks went quiet (-1) / came back (+1) in this window
    uint4           c_lo = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0), c_hi = make_uint4(0, 0, 0, 0);
    if ( valid && !d.ctl_free ) {
        const uint4* ctlp = reinterpret_cast<const uint4*>(&d.ctl[lc]);
        c_lo              = ctlp[0];
        c_hi              = ctlp[1];
    }
    const uint64_t period    = ((uint64_t)c_hi.y << 32) | c_hi.x;
    const uint64_t next_cyc  = ((uint64_t)c_lo.y << 32) | c_lo.x;
    uint64_t       next_tick = next_cyc == TIME_NEVER ? TIME_NEVER : next_cyc * period;
    const uint32_t my_type   = d.uniform_type ? d.uniform_type : c_hi.z;
    const bool     has_work  = valid && (my_cnt > 0 || next_tick < T_end || (c_hi.w & 1u));
    const uint32_t p0        = s_port_off[c];
    const uint4*   my_links  = links_in_smem ? (const uint4*)(s_link + (p0 - blk_port0)) : d.link + (p0 - d.P0);

    if ( has_work ) {
        // ---- phase 3 (sequential form): walks the component's segment in delivery order, merged with its clock
        // ticks; sends are staged in the slots the component has already consumed
        CtxT<true> ctx { d, T, k, T, d.c0 + lc, p0, s_port_off[c + 1] - p0,
                         d.aux + (uint64_t)lc * d.aux_stride, c_lo.z, 0, d.stats_on, my_links,
                         Stage { s_time, s_dst, s_sc, PAYLOAD ? s_pay : nullptr }, seg, 0, 0 };
        uint32_t   dseq = c_lo.w;
        with_element_of<EU>(my_type, [&](auto el) {
            using E = decltype(el);
            typename E::State s;
            // statistics: elements that add data with every delivery get their statistics region in the same round
            // trip as the rest of the state; for the others this window's statistics start empty and are folded into
            // memory at the end, if the window added any
            const bool stats_direct = E::STATS_EVERY_EVENT && d.stats_on && my_cnt > 0;
            uint8_t* const st_base = stats_base<E>(d.state, d.state_stride, d.aux, d.aux_stride, lc);
            load_state(s, d.state + (uint64_t)lc * d.state_stride, 0, E::CONST_END);
            if ( stats_direct )
                load_state(s, st_base, E::CONST_END, E::STATS_END);
            else
                E::initStats(s);
            E::beginWindow(ctx, s, my_cnt); // issues the element's first loads (MT19937 read-ahead)
            sort_segment(s_perm + s_off[c], my_cnt, s_time, s_dst, s_sc);
            uint64_t cycle = next_cyc;
            uint32_t i     = 0;
            constexpr bool MC = multi_clock_of<E>::value, SL = self_links_of<E>::value;
            const bool     had_clock = next_tick != TIME_NEVER;
            SelfQueue      selfq;
            if constexpr ( SL ) {
                selfq.n   = 0;
                ctx.selfq = &selfq;
                ctx.T_end = T_end;
            }
            const uint32_t r_off   = REPORT ? s_misc[1] : 0xFFFFFFFFu;
            const uint64_t R       = T + r_off;
            bool           snapped = r_off == 0xFFFFFFFFu;
            for ( ;; ) {
                uint64_t ev_t      = (i < my_cnt) ? T + s_time[seg[i]] : TIME_NEVER;
                bool     from_self = false;
                if constexpr ( SL ) {
                    // the next event is the earlier of the calendar's and the self queue's, by (time, port, sequence)
                    if ( selfq.n ) {
                        const uint64_t st = selfq.time[0];
                        if ( st < ev_t )
                            from_self = true;
                        else if ( st == ev_t ) {
                            const uint32_t e0 = seg[i];
                            from_self = selfq.port[0] < s_dst[e0] ||
                                        (selfq.port[0] == s_dst[e0] && (int32_t)(selfq.seq[0] - s_sc[e0].x) < 0);
                        }
                        if ( from_self ) ev_t = st;
                    }
                }
                if constexpr ( stat_clock_of<E>::value ) {
                    // the statistics engine's activities (priority 85): after the clocks and events of the same time
                    const uint64_t st = E::nextStatTime(s);
                    if ( st < T_end && st < next_tick && st < ev_t ) {
                        ctx.now = st;
                        E::statFire(ctx, s, st);
                        continue;
                    }
                }
                if ( REPORT && !snapped && ev_t > R && !(next_tick < T_end && next_tick <= R) ) {
                    // everything up to the report time has run: leave the component's state as of now in the snapshot
                    // copy (results_kernel turns it into result records afterwards).  Statistics collected per window
                    // are folded with what is in memory first.
                    uint8_t* const snap = d.snapshot + (uint64_t)lc * d.state_stride;
                    uint8_t* const snap_st = stats_base<E>(d.snapshot, d.state_stride, d.snapshot_aux, d.aux_stride, lc);
                    store_state(s, snap, 0, E::CONST_END);
                    if ( stats_direct ) {
                        store_state(s, snap_st, E::CONST_END, E::STATS_END);
                    }
                    else if ( d.stats_on && E::statsVersion(s) != 0 ) {
                        typename E::State mem;
                        load_state(mem, st_base, E::CONST_END, E::STATS_END);
                        E::mergeStats(mem, s);
                        store_state(mem, snap_st, E::CONST_END, E::STATS_END);
                    }
                    snapped = true;
                }
                if ( next_tick < T_end && next_tick <= ev_t ) {
                    ctx.now = next_tick;
                    const long long tc0 = d.prof_clock_cyc ? clock64() : 0;
                    if constexpr ( MC ) {
                        // every clock of the component due now, each with all its handlers (ClockSet::fire)
                        ticks += E::clockFire(ctx, s, next_tick);
                        next_tick = E::nextClock(s);
                    }
                    else {
                        const bool stop = E::clockTick(ctx, s, cycle);
                        ticks++;
                        cycle++;
                        if ( stop )
                            next_tick = TIME_NEVER; // handler unregistered itself (clock.cc:324-367)
                        else
                            next_tick += period;
                    }
                    if ( d.prof_clock_cyc ) d.prof_clock_cyc[lc] += (unsigned long long)(clock64() - tc0);
                    continue;
                }
                uint32_t port;
                uint64_t pl;
                if ( SL && from_self ) {
                    port = selfq.port[0] - ctx.port0;
                    pl   = selfq.pay[0];
                    selfq.n--;
                    for ( uint32_t j = 0; j < selfq.n; ++j ) {
                        selfq.time[j] = selfq.time[j + 1];
                        selfq.port[j] = selfq.port[j + 1];
                        selfq.seq[j]  = selfq.seq[j + 1];
                        selfq.pay[j]  = selfq.pay[j + 1];
                    }
                }
                else {
                    if ( i >= my_cnt ) break;
                    const uint32_t e = seg[i++];
                    port             = s_dst[e] - ctx.port0;
                    pl               = PAYLOAD ? s_pay[e] : 0;
                    s_dst[e]         = SSTB200_NO_PEER; // slot e is consumed: free for out-staging
                    ctx.free_tail    = i;
                }
                ctx.now = ev_t;
                if ( d.trace_cap ) {
                    const unsigned long long tp = atomicAdd(&ctrl->trace_n, 1ull);
                    if ( tp < d.trace_cap ) {
                        d.tr_time[tp]    = ev_t;
                        d.tr_comp[tp]    = ctx.comp;
                        d.tr_port[tp]    = port;
                        d.tr_idx[tp]     = dseq;
                        d.tr_payload[tp] = pl;
                    }
                    dseq++;
                }
                if ( d.prof_recv_cyc ) {
                    const long long te0 = clock64();
                    E::handleEvent(ctx, s, (int)port, pl);
                    atomicAdd(&d.prof_recv_cyc[ctx.port0 + port - d.P0], (unsigned long long)(clock64() - te0));
                }
                else
                    E::handleEvent(ctx, s, (int)port, pl);
                delivered++;
                if constexpr ( MC ) next_tick = E::nextClock(s); // the handler may have started or stopped a clock
            }
            clock_delta = (int32_t)(next_tick != TIME_NEVER) - (int32_t)had_clock;
            next_store  = (MC || next_tick == TIME_NEVER) ? next_tick : cycle;
            store_state(s, d.state + (uint64_t)lc * d.state_stride, 0, E::STORE_END);
            // statistics: read-modify-write only when the window added data (a clock-driven element may tick a thousand
            // times between two addData calls)
            if ( stats_direct ) {
                store_state(s, st_base, E::CONST_END, E::STATS_END);
            }
            else if ( d.stats_on && E::statsVersion(s) != 0 ) {
                typename E::State mem;
                load_state(mem, st_base, E::CONST_END, E::STATS_END);
                E::mergeStats(mem, s);
                store_state(mem, st_base, E::CONST_END, E::STATS_END);
            }
        });
        if ( !d.ctl_free )
            *reinterpret_cast<uint4*>(&d.ctl[lc]) =
                make_uint4((uint32_t)next_store, (uint32_t)(next_store >> 32), ctx.seq, dseq);
        sent = ctx.sent;
        if ( d.prof_clock && ticks ) d.prof_clock[lc] += ticks;
    }
    __syncthreads();
    PHASE_MARK(4);

    // ---- phase 4: flush the staged sends, the whole CTA over all slots (balanced: ~4 records per thread however the
    // events were distributed over components).  Destination bins are counted in a small shared hash table
    // (key = window slot x destination block, or the destination rank): a record's rank inside its bin comes from a
    // shared-memory atomic, ONE global atomic per distinct bin reserves the calendar slots, then the records are
    // written side by side.  Records that find no table entry (more than HT distinct bins per CTA: irregular graphs)
    // reserve their slot directly.  (Nothing is staged in a block without due events.)
    if ( n_due ) {
        constexpr uint32_t HT = 128;
        uint32_t* ht_key  = s_cnt;      // [HT] phase-2 scratch is dead by now
        uint32_t* ht_cnt  = s_cnt + HT; // [HT]
        uint32_t* ht_base = s_off;      // [HT] (low 32 bits of the reserved base)
        if ( tid < HT ) {
            ht_key[tid] = 0xFFFFFFFFu;
            ht_cnt[tid] = 0;
        }
        __syncthreads();
#ifndef SSTB200_MAXR
#define SSTB200_MAXR 5
#endif
        constexpr int MAXR = SSTB200_MAXR; // slots tid, tid+256, ... per thread in registers between the passes
        uint32_t      r_key[MAXR], r_rank[MAXR], r_ent[MAXR];
        auto          classify_out = [&](uint32_t e, uint32_t& ky_out, uint32_t& ent_out, uint32_t& rank_out) {
            // called by all 32 lanes of a warp together (lanes without a record carry the key 0xFFFFFFFF)
            ent_out      = 0xFFFFFFFFu;
            uint32_t ky  = 0xFFFFFFFFu, dt = 0, dcomp = 0;
            if ( e < n_due && s_dst[e] != SSTB200_NO_PEER ) {
                dcomp = s_sc[e].y;
                dt    = s_time[e];
                const bool remote = d.n_ranks > 1 && (dcomp < d.c0 || dcomp >= d.c0 + d.n_local);
                if ( dt < w && (!remote || d.peer_on) ) {
                    raise_error(d, ERR_TIME_FAULT, T + dt, T, dcomp); // a send violated the lookahead
                }
                else {
                    uint32_t sl = slot1;
                    if ( dt >= 2 * w ) {
                        uint32_t dk = div_w(d, dt);
                        if ( dk > d.W - 1 ) dk = d.W - 1; // beyond the horizon: parked in the farthest slot
                        sl = wrap_slot(d, slot + dk);
                    }
                    ky = remote ? remote_key(d, dcomp, sl) : sl * d.nbins + (dcomp - d.c0) / BLOCK;
                }
            }
            // lanes of the warp with the same bin go through the table as ONE entry look-up + ONE counter update
#ifdef SSTB200_FLUSH_MATCH
            const uint32_t peers  = __match_any_sync(0xffffffffu, ky); // lanes with the same bin share one look-up
#else
            const uint32_t peers  = 1u << lane; // every record looks its bin up itself (measured: same speed, less code)
#endif
            const uint32_t leader = __ffs(peers) - 1;
            uint32_t       ent = 0xFFFFFFFFu, base = 0;
            if ( ky != 0xFFFFFFFFu && lane == leader ) {
                uint32_t h = (ky * 2654435761u) >> 25;
#pragma unroll
                for ( int pr = 0; pr < 4; ++pr ) { // open addressing, 4 probes
                    uint32_t old = ht_key[h]; // usually already claimed for this bin: no atomic needed
                    if ( old != ky ) old = atomicCAS(&ht_key[h], 0xFFFFFFFFu, ky);
                    if ( old == 0xFFFFFFFFu || old == ky ) {
                        ent = h;
                        break;
                    }
                    h = (h + 1) & (HT - 1);
                }
                if ( ent != 0xFFFFFFFFu ) base = atomicAdd(&ht_cnt[ent], (uint32_t)__popc(peers));
            }
            ent  = __shfl_sync(0xffffffffu, ent, leader);
            base = __shfl_sync(0xffffffffu, base, leader);
            if ( ky == 0xFFFFFFFFu ) return;
            if ( ent != 0xFFFFFFFFu ) {
                ky_out   = ky;
                ent_out  = ent;
                rank_out = base + __popc(peers & ((1u << lane) - 1));
            }
            else { // table full: reserve directly
                emit(d, k, T + dt, s_dst[e], dcomp, s_sc[e].x, PAYLOAD ? s_pay[e] : 0);
            }
        };
#pragma unroll
        for ( int j = 0; j < MAXR; ++j )
            classify_out(tid + j * BLOCK, r_key[j], r_ent[j], r_rank[j]);
        for ( uint32_t e = tid + MAXR * BLOCK; e < n_due; e += BLOCK ) // beyond the register window: direct
            if ( s_dst[e] != SSTB200_NO_PEER )
                emit(d, k, T + s_time[e], s_dst[e], s_sc[e].y, s_sc[e].x, PAYLOAD ? s_pay[e] : 0);
        __syncthreads();
        if ( tid < HT && ht_key[tid] != 0xFFFFFFFFu && ht_cnt[tid] ) {
            const uint32_t ky = ht_key[tid];
            if ( ky & 0x80000000u ) {
                ht_base[tid] = remote_reserve(d, ky, ht_cnt[tid]);
                atomicAdd(&ctrl->remote_events, (unsigned long long)ht_cnt[tid]);
            }
            else
                ht_base[tid] = atomicAdd(&d.bin_count[ky], ht_cnt[tid]);
        }
        __syncthreads();
#pragma unroll
        for ( int j = 0; j < MAXR; ++j ) {
            if ( r_ent[j] == 0xFFFFFFFFu ) continue;
            const uint32_t e   = tid + j * BLOCK;
            const uint32_t ky  = r_key[j];
            const uint32_t pos = ht_base[r_ent[j]] + r_rank[j];
            const uint64_t tm  = T + s_time[e];
            const uint64_t ds  = ((uint64_t)s_dst[e] << 32) | s_sc[e].x;
            if ( ky & 0x80000000u ) {
                remote_write(d, ky, pos, tm, ds, PAYLOAD ? s_pay[e] : 0, s_sc[e].y);
            }
            else if ( pos >= d.cap ) {
                raise_error(d, ERR_BIN_OVERFLOW, ky / d.nbins, ky % d.nbins, pos);
            }
            else {
                const uint64_t off = (uint64_t)ky * d.cap + pos;
                d.ev[off]          = make_ulonglong2(tm, ds);
                if ( PAYLOAD ) d.ev_payload[off] = s_pay[e];
            }
        }
    }

    PHASE_MARK(5);
    // ---- phase 5: fold counters (warp shuffles, then one value per warp through shared memory: ONE atomic per counter
    // and CTA -- a clock population runs 65 536 CTAs per window, and every same-address atomic is serialised in L2),
    // recycle the bin
    {
        uint32_t v0 = delivered, v1 = ticks, v2 = sent, v3 = (uint32_t)clock_delta, v4 = dropped;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
            v2 += __shfl_down_sync(0xffffffffu, v2, o);
            v3 += __shfl_down_sync(0xffffffffu, v3, o);
            v4 += __shfl_down_sync(0xffffffffu, v4, o);
        }
        uint32_t* s_red = s_hist; // [5][BLOCK/32]: phase-2 scratch (s_hist + s_hbase, 64 words) is dead by now
        if ( lane == 0 ) {
            const uint32_t wp = tid >> 5;
            s_red[0 * (BLOCK / 32) + wp] = v0;
            s_red[1 * (BLOCK / 32) + wp] = v1;
            s_red[2 * (BLOCK / 32) + wp] = v2;
            s_red[3 * (BLOCK / 32) + wp] = v3;
            s_red[4 * (BLOCK / 32) + wp] = v4;
        }
        __syncthreads();
        if ( tid == 0 ) {
            v0 = v1 = v2 = v3 = v4 = 0;
#pragma unroll
            for ( int wp = 0; wp < BLOCK / 32; ++wp ) {
                v0 += s_red[0 * (BLOCK / 32) + wp];
                v1 += s_red[1 * (BLOCK / 32) + wp];
                v2 += s_red[2 * (BLOCK / 32) + wp];
                v3 += s_red[3 * (BLOCK / 32) + wp];
                v4 += s_red[4 * (BLOCK / 32) + wp];
            }
            if ( v0 ) atomicAdd(&ctrl->events_delivered, (unsigned long long)v0);
            if ( v1 ) atomicAdd(&ctrl->clock_ticks, (unsigned long long)v1);
            if ( v2 ) atomicAdd(&ctrl->events_sent, (unsigned long long)v2);
            const long long dp = (long long)v2 - (long long)v0 - (long long)v4;
            if ( dp ) atomicAdd((unsigned long long*)&ctrl->local[CW_PENDING], (unsigned long long)dp);
            if ( v3 ) atomicAdd((unsigned long long*)&ctrl->local[CW_CLOCKS], (unsigned long long)(long long)(int32_t)v3);
        }
    }
    if ( tid == 0 ) {
        // every thread of the CTA finished reading this bin before the phase-1 barrier
        d.bin_count[slot * d.nbins + bin] = 0;
        if ( n_in > ctrl->max_bin_fill ) atomicMax(&ctrl->max_bin_fill, (unsigned long long)n_in);
        if ( n_due > ctrl->max_due ) atomicMax(&ctrl->max_due, (unsigned long long)n_due);
    }
}

#ifndef SSTB200_MIN_BLOCKS
#define SSTB200_MIN_BLOCKS 3 // 80 registers, no spills, 3 CTAs/SM: measured 27 % faster than 4 CTAs/SM at 64 registers with spills
#endif
// LIST false: one CTA per block of the rank.  LIST true: the blocks the event-parallel kernel left to this form
// (Dev::fb_list), walked by a small grid.
template <bool PAYLOAD, bool REPORT = false, class EU = void, bool LIST = false>
__global__ void __launch_bounds__(BLOCK, SSTB200_MIN_BLOCKS)
dispatch_kernel(Dev d)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    if ( d.ctrl->done ) return;
    if constexpr ( !LIST ) {
        dispatch_block<PAYLOAD, REPORT, EU>(d, blockIdx.x, smem_raw);
    }
    else {
        const uint32_t n = min(d.ctrl->fb_n, d.nbins);
        for ( uint32_t i = blockIdx.x; i < n; i += gridDim.x ) {
            dispatch_block<PAYLOAD, REPORT, EU>(d, d.fb_list[i], smem_raw);
            __syncthreads();
        }
    }
}

} // namespace sstb200
#include "parallel_form.cuh"
#include "ordered_form.cuh"
#include "clock_form.cuh"
namespace sstb200 {

// ---- wire-up on the device: link row of every local directed port = {peer port, component owning the peer (binary
// search in the GLOBAL port CSR), send latency}.  Replaces the per-Link object setup of Simulation::prepareLinks
// (simulation.cc:668-820) for 67 M ports without a host loop.
__global__ void __launch_bounds__(BLOCK)
build_links_kernel(uint4* link, const uint32_t* peer, const uint64_t* latency, const uint32_t* port_off_global,
    uint32_t ncomp_global, uint32_t nport_local, uint32_t global_shift)
{
    for ( uint32_t p = blockIdx.x * BLOCK + threadIdx.x; p < nport_local; p += gridDim.x * BLOCK ) {
        const uint32_t pr = peer[p];
        uint32_t       lo = 0;
        if ( pr != SSTB200_NO_PEER && global_shift < 32 ) {
            lo = pr >> global_shift; // every component of the model has 2^shift ports: no table needed
        }
        else if ( pr != SSTB200_NO_PEER ) {
            uint32_t hi = ncomp_global; // largest c with port_off[c] <= pr
            while ( hi - lo > 1 ) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if ( port_off_global[mid] <= pr )
                    lo = mid;
                else
                    hi = mid;
            }
        }
        const uint64_t lat = latency[p];
        link[p]            = make_uint4(pr, lo, (uint32_t)(lat & 0xFFFFFFFFull), (uint32_t)(lat >> 32));
    }
}

// ---- bin the events received from other ranks (RankSyncSerialSkip::exchange re-send, rankSyncSerialSkip.cc:291-295)
__global__ void __launch_bounds__(BLOCK)
ingest_kernel(Dev d)
{
    if ( d.ctrl->done ) return;
    const uint64_t k = d.ctrl->k;
    for ( uint32_t r = 0; r < d.n_ranks; ++r ) {
        if ( r == d.rank ) continue;
        const unsigned long long* box = d.inbox + (uint64_t)r * d.xwords;
        const uint64_t            n   = box[0] < d.outbox_cap ? box[0] : d.outbox_cap;
        for ( uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK ) {
            const unsigned long long* rec = box + XHDR + i * 4;
            emit(d, k, rec[0], (uint32_t)(rec[1] >> 32), (uint32_t)rec[3], (uint32_t)rec[1], rec[2]);
        }
    }
}

// ---- open the next window (SyncManager::computeNextInsert syncManager.cc:772-798; Exit::check exit.cc:111-132)
__device__ void
advance_fold(const Dev& d)
{
    Ctrl* c = d.ctrl;
    long long pending = 0, primary = 0, clocks = 0, err = 0;
    uint64_t  exit_time = 0;
    if ( d.n_ranks > 1 ) {
        for ( uint32_t r = 0; r < d.n_ranks; ++r ) {
            const long long* g = d.gathered + (uint64_t)r * CW_WORDS;
            pending += g[CW_PENDING];
            primary += g[CW_PRIMARY];
            clocks += g[CW_CLOCKS];
            err += g[CW_ERROR];
            if ( (uint64_t)g[CW_EXIT_TIME] > exit_time ) exit_time = (uint64_t)g[CW_EXIT_TIME];
        }
        if ( !d.peer_on )
            for ( uint32_t r = 0; r < d.n_ranks; ++r )
                d.outbox[(uint64_t)r * d.xwords] = 0; // outboxes have been shipped
    }
    else {
        pending   = c->local[CW_PENDING];
        primary   = c->local[CW_PRIMARY];
        clocks    = c->local[CW_CLOCKS];
        exit_time = (uint64_t)c->local[CW_EXIT_TIME];
    }
    c->windows++;
    c->fb_n     = 0;
    c->fb_taken = 0;
    c->maint_n  = 0;
    c->pf_next  = 0;
    if ( c->error || err ) {
        if ( !c->error ) c->error = (uint32_t)0x80000000u; // another rank failed
        c->done = 1;
        return;
    }
    if ( c->had_primary && primary <= 0 ) { // Exit: last primary component said OKToEndSim (exit.cc:53-68)
        c->done     = 1;
        c->end_time = exit_time;
        return;
    }
    const uint64_t next_T = (c->k + 1) * c->w;
    if ( c->stop_at && next_T >= c->stop_at ) { // StopAction
        c->done     = 1;
        c->end_time = c->stop_at;
        return;
    }
    if ( pending <= 0 && clocks <= 0 ) { // nothing left in the (virtual) TimeVortex
        c->done     = 1;
        c->end_time = c->T_end;
        return;
    }
    c->k     = c->k + 1;
    c->slot  = (uint32_t)(c->k % d.W);
    c->T     = next_T;
    c->T_end = (c->stop_at && next_T + c->w > c->stop_at) ? c->stop_at : next_T + c->w;
}

__global__ void
advance_kernel(Dev d)
{
    if ( threadIdx.x != 0 || blockIdx.x != 0 || d.ctrl->done ) return;
    advance_fold(d);
}

// ---- peer mode: the window's exchange.  The boundary events are already in their owners' calendars (emit /
// remote_write); what RankSyncSerialSkip::exchange still needs (rankSyncSerialSkip.cc:209-343: "everybody is done with
// this window" + the MIN/flag all-reduce) is one mailbox write per rank: this rank's control words and the sequence
// number of the window, stored over NVLink into every rank's mailbox, after a system-scope fence.
__device__ __forceinline__ void
peer_post(const Dev& d)
{
    Ctrl* const              c   = d.ctrl;
    const unsigned long long seq = c->windows + 1;
    const uint32_t           par = (uint32_t)(seq & 1ull);
    if ( threadIdx.x == 0 ) c->local[CW_ERROR] = c->error ? 1 : 0;
    __syncthreads();
    __threadfence_system(); // this rank's records are in the peers' calendars before anybody sees the sequence number
    const uint32_t r = threadIdx.x;
    if ( r < d.n_ranks ) {
        volatile unsigned long long* m = d.peer_mail[r] + ((uint64_t)par * d.n_ranks + d.rank) * MAILW;
        for ( uint32_t i = 0; i < (uint32_t)CW_WORDS; ++i )
            m[i] = (unsigned long long)c->local[i];
        __threadfence_system();
        m[CW_WORDS] = seq;
    }
}
__global__ void
peer_post_kernel(Dev d)
{
    if ( d.ctrl->done || blockIdx.x != 0 ) return;
    peer_post(d);
}

// wait for every rank's mailbox entry of this window (spin: the peers are other GPUs running their own kernels; with
// all ranks stepped by one process on one GPU the entries are already there and a missing one is an error), then fold
// the control words and open the next window exactly like advance_kernel
__device__ __forceinline__ void
peer_wait_and_fold(const Dev& d, int spin) // called by a whole CTA (>= n_ranks threads)
{
    Ctrl* c = d.ctrl;
    const unsigned long long seq = c->windows + 1;
    const uint32_t           par = (uint32_t)(seq & 1ull);
    const uint32_t           r   = threadIdx.x;
    if ( r < d.n_ranks ) {
        volatile unsigned long long* m = d.mail + ((uint64_t)par * d.n_ranks + r) * MAILW;
        if ( spin ) {
            const long long t0 = clock64();
            unsigned long long g0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
            while ( m[CW_WORDS] != seq ) {
                __nanosleep(200);
                if ( clock64() - t0 > 20000000000ll ) break; // ~10 s: a peer is gone
            }
            if ( r == (d.rank + 1) % d.n_ranks ) { // one sample per window: how long this rank waited for its neighbour
                unsigned long long g1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
                c->sync_wait_ns += g1 - g0;
            }
        }
        if ( m[CW_WORDS] != seq ) raise_error(d, ERR_PEER_SYNC, r, seq, m[CW_WORDS]);
        __threadfence_system();
        for ( uint32_t i = 0; i < (uint32_t)CW_WORDS; ++i )
            d.gathered[(uint64_t)r * CW_WORDS + i] = (long long)m[i];
    }
    __syncthreads();
    if ( threadIdx.x == 0 ) advance_fold(d);
}

__global__ void
peer_advance_kernel(Dev d, int spin, int post_first)
{
    Ctrl* c = d.ctrl;
    if ( c->done || blockIdx.x != 0 ) return;
    if ( post_first ) peer_post(d); // ranks on their own GPUs: the mailbox write and the wait in one launch
    peer_wait_and_fold(d, spin);
}

__global__ void
publish_control_kernel(Dev d)
{
    // after dispatch: snapshot this rank's control words (with the error flag) into the header of every outbox slot,
    // so the all-to-all that ships the boundary events also ships them (RankSyncSerialSkip needs a second collective,
    // MPI_Allreduce, for this: rankSyncSerialSkip.cc:318-335)
    if ( blockIdx.x != 0 ) return;
    if ( threadIdx.x == 0 ) d.ctrl->local[CW_ERROR] = d.ctrl->error ? 1 : 0;
    __syncthreads();
    if ( d.n_ranks > 1 )
        for ( uint32_t i = threadIdx.x; i < d.n_ranks * CW_WORDS; i += blockDim.x )
            d.outbox[(uint64_t)(i / CW_WORDS) * d.xwords + 1 + (i % CW_WORDS)] =
                (unsigned long long)d.ctrl->local[i % CW_WORDS];
}

// after the exchange: every rank's control words, as received in the inbox headers (own words from ctrl)
__global__ void
collect_control_kernel(Dev d)
{
    if ( blockIdx.x != 0 ) return;
    for ( uint32_t i = threadIdx.x; i < d.n_ranks * CW_WORDS; i += blockDim.x ) {
        const uint32_t r = i / CW_WORDS, wd = i % CW_WORDS;
        d.gathered[i]    = (r == d.rank) ? d.ctrl->local[wd] : (long long)d.inbox[(uint64_t)r * d.xwords + 1 + wd];
    }
}

__global__ void __launch_bounds__(BLOCK)
results_kernel(Dev d, uint64_t* out, uint32_t words)
{
    const uint32_t lc = blockIdx.x * BLOCK + threadIdx.x;
    if ( lc >= d.n_local ) return;
    uint64_t r[NRESULT];
#pragma unroll
    for ( int i = 0; i < NRESULT; ++i )
        r[i] = 0;
    with_element(d.ctl[lc].type, [&](auto el) {
        using E = decltype(el);
        typename E::State s;
        load_state(s, d.state + (uint64_t)lc * d.state_stride, 0, E::CONST_END);
        load_state(s, stats_base<E>(d.state, d.state_stride, d.aux, d.aux_stride, lc), E::CONST_END, E::PERSIST_BYTES);
        E::results(s, r);
    });
#pragma unroll
    for ( int i = 0; i < NRESULT; ++i )
        if ( (uint32_t)i < words ) out[(uint64_t)lc * words + i] = r[i];
}

// ================================================================================================ RNG stream kernels
// BASELINE config 2: n_streams independent generators x `draws` u32 each, written stream-major, plus the position-
// weighted checksum sum (i+1) * draw_i per stream.
// Marsaglia / xorshift (8 / 16 bytes of state): a lane per generator, 64 draws at a time through a shared-memory tile;
// the warp then writes one generator's 64 draws per store instruction (256 contiguous bytes), so the stores are full
// DRAM lines instead of one 4-byte word per lane at a stride of `draws` words.
constexpr uint32_t RNG_TILE = 64, RNG_ROW = RNG_TILE + 4; // row stride 68 words: rows 16-byte aligned, 16-byte
                                                          // accesses of 8 lanes (a quarter warp) on distinct banks
template <class G>
__device__ __forceinline__ void
small_stream_body(G& g, bool valid, uint64_t draws, uint32_t* warp_out, uint64_t stream_stride, uint32_t lane,
    uint32_t (*tile)[RNG_ROW], unsigned long long* checksum)
{
    unsigned long long cs = 0;
    const bool         wide = warp_out && (stream_stride & 3u) == 0 && (((uintptr_t)warp_out) & 15u) == 0;
    for ( uint64_t base = 0; base < draws; base += RNG_TILE ) {
        const uint32_t n = (uint32_t)min((uint64_t)RNG_TILE, draws - base);
        if ( valid ) {
            if ( n == RNG_TILE && (base >> 32) == 0 ) {
                const uint32_t b32 = (uint32_t)base;
#pragma unroll 4
                for ( uint32_t j = 0; j < RNG_TILE; j += 4 ) { // four draws per 16-byte shared store
                    uint4 v;
                    v.x = g.u32();
                    v.y = g.u32();
                    v.z = g.u32();
                    v.w = g.u32();
                    cs += (unsigned long long)(b32 + j + 1) * v.x;
                    cs += (unsigned long long)(b32 + j + 2) * v.y;
                    cs += (unsigned long long)(b32 + j + 3) * v.z;
                    cs += (unsigned long long)(b32 + j + 4) * v.w;
                    if ( warp_out ) *reinterpret_cast<uint4*>(&tile[lane][j]) = v;
                }
            }
            else
                for ( uint32_t j = 0; j < n; ++j ) {
                    const uint32_t v = g.u32();
                    cs += (base + j + 1) * (unsigned long long)v;
                    if ( warp_out ) tile[lane][j] = v;
                }
        }
        if ( warp_out ) {
            __syncwarp();
            const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
            if ( wide && n == RNG_TILE && valid_mask == 0xffffffffu ) {
                // two generators' rows per store instruction: a half warp writes 256 contiguous bytes of each
                const uint32_t half = lane >> 4, q = lane & 15u;
#pragma unroll 4
                for ( uint32_t sidx = 0; sidx < 32; sidx += 2 ) {
                    const uint32_t r = sidx + half;
                    reinterpret_cast<uint4*>(warp_out + (uint64_t)r * stream_stride + base)[q] =
                        *reinterpret_cast<const uint4*>(&tile[r][4 * q]);
                }
            }
            else
                for ( uint32_t sidx = 0; sidx < 32; ++sidx ) {
                    if ( !(valid_mask >> sidx & 1u) ) continue;
                    uint32_t* o = warp_out + (uint64_t)sidx * stream_stride + base;
                    for ( uint32_t j = lane; j < n; j += 32 )
                        o[j] = tile[sidx][j];
                }
            __syncwarp();
        }
    }
    if ( valid ) *checksum = cs;
}

__global__ void __launch_bounds__(128)
rng_small_stream_kernel(int kind, uint64_t s1, uint64_t s2, uint64_t stride1, uint64_t stride2, uint64_t n_streams,
    uint64_t draws, uint32_t* out, unsigned long long* checksum)
{
    __shared__ __align__(16) uint32_t tile[4][32][RNG_ROW];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t j     = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool     valid = j < n_streams;
    uint32_t*      wout  = out ? out + (j - lane) * draws : nullptr;
    const uint64_t a = s1 + j * stride1, b = s2 + j * stride2;
    if ( kind == 1 ) {
        MarsagliaDev g { (uint32_t)a, (uint32_t)b };
        small_stream_body(g, valid, draws, wout, draws, lane, tile[warp], checksum + j);
    }
    else {
        XorShiftDev g;
        g.seed((uint32_t)a);
        small_stream_body(g, valid, draws, wout, draws, lane, tile[warp], checksum + j);
    }
}

// MersenneRNG seeding (mersenne.cc:149-159) for the stream kernel: the recurrence is serial per generator, so a LANE
// runs each generator's; 64 words at a time go through a shared-memory tile (16-byte accesses on both sides, as in
// small_stream_body) and leave as 256-byte rows of the generator's state.  One lane of a warp doing this inside the
// stream kernel cost it a fifth of its issue slots.
__global__ void __launch_bounds__(128)
rng_mt_seed_kernel(uint64_t s1, uint64_t stride1, uint64_t j0, uint64_t n, uint32_t* state /* [n][MT_N] */)
{
    __shared__ __align__(16) uint32_t tile[4][32][RNG_ROW];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t g     = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; // generator j0 + g
    const uint64_t g0    = g - lane;
    uint32_t       prev  = (uint32_t)(s1 + (j0 + g) * stride1);
    for ( uint32_t base = 0; base < (uint32_t)MT_N; base += RNG_TILE ) {
        const uint32_t nw = min((uint32_t)RNG_TILE, (uint32_t)MT_N - base); // 64, the last time 48
#pragma unroll 4
        for ( uint32_t j = 0; j < nw; j += 4 ) {
            uint32_t v[4];
#pragma unroll
            for ( uint32_t u = 0; u < 4; ++u ) {
                const uint32_t i = base + j + u;
                if ( i ) prev = 1812433253u * (prev ^ (prev >> 30)) + i;
                v[u] = prev;
            }
            *reinterpret_cast<uint4*>(&tile[warp][lane][j]) = make_uint4(v[0], v[1], v[2], v[3]);
        }
        __syncwarp();
        const uint32_t half = lane >> 4, q = lane & 15u;
        for ( uint32_t r0 = 0; r0 < 32; r0 += 2 ) {
            const uint32_t r = r0 + half;
            if ( g0 + r < n && 4 * q < nw )
                reinterpret_cast<uint4*>(state + (g0 + r) * MT_N + base)[q] = *reinterpret_cast<const uint4*>(&tile[warp][r][4 * q]);
        }
        __syncwarp();
    }
}

// MersenneRNG (2.5 KB of state): a WARP per generator, the state in shared memory, seeded by rng_mt_seed_kernel (or,
// without a seeded state, by one lane); every generation is then built by the 32 lanes together and its 624 draws are
// tempered and written by the warp as contiguous 128-byte lines.
__global__ void __launch_bounds__(256)
rng_mt_stream_kernel(uint64_t s1, uint64_t stride1, uint64_t j0, uint64_t n_streams, uint64_t draws, uint32_t* out,
    unsigned long long* checksum, const uint32_t* seeded /* [n_streams - j0 ...][MT_N] or null */)
{
    __shared__ __align__(16) uint32_t s_gen[8][MT_N];
    const uint32_t lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    uint32_t* const gen = s_gen[wic];
    const uint64_t  nw = (uint64_t)gridDim.x * 8;
    for ( uint64_t j = j0 + (uint64_t)blockIdx.x * 8 + wic; j < n_streams; j += nw ) {
        if ( seeded )
            mt_warp_load(gen, seeded + (j - j0) * MT_N, lane);
        else if ( lane == 0 )
            mt_seed(gen, (uint32_t)(s1 + j * stride1));
        __syncwarp();
        unsigned long long cs = 0;
        uint32_t*          o  = out ? out + j * draws : nullptr;
        for ( uint64_t base = 0; base < draws; base += MT_N ) {
            const uint32_t n = (uint32_t)min((uint64_t)MT_N, draws - base);
            // the twist (mt_warp_twist: chunks of 32 words in increasing order, reads of a chunk before its writes) with
            // the tempering, the checksum and the store of every new word folded into the same pass.  Four chunks go
            // through together: none of them reads a word another one of the four writes (word i reads i, i+1 and
            // i+397 -- 12.4 chunks ahead or 7.1 behind), so their twelve loads are in flight at once.
            constexpr uint32_t NCH = (MT_N + 31) / 32, GRP = 4;
            for ( uint32_t c0 = 0; c0 < NCH; c0 += GRP ) {
                uint32_t v[GRP];
#pragma unroll
                for ( uint32_t u = 0; u < GRP; ++u ) {
                    const uint32_t i = (c0 + u) * 32 + lane;
                    v[u]             = 0;
                    if ( i < (uint32_t)MT_N )
                        v[u] = mt_twist(gen[i], gen[i + 1 == (uint32_t)MT_N ? 0 : i + 1],
                            gen[i + MT_M >= (uint32_t)MT_N ? i + MT_M - MT_N : i + MT_M]);
                }
                __syncwarp();
#pragma unroll
                for ( uint32_t u = 0; u < GRP; ++u ) {
                    const uint32_t i = (c0 + u) * 32 + lane;
                    if ( i < (uint32_t)MT_N ) {
                        gen[i] = v[u];
                        if ( i < n ) {
                            const uint32_t t   = mt_temper(v[u]);
                            const uint64_t pos = base + i + 1;
                            cs += (pos >> 32) == 0 ? (unsigned long long)(uint32_t)pos * t : pos * (unsigned long long)t;
                            if ( o ) o[base + i] = t;
                        }
                    }
                }
                __syncwarp();
            }
        }
#pragma unroll
        for ( int off = 16; off > 0; off >>= 1 )
            cs += __shfl_down_sync(0xffffffffu, cs, off);
        if ( lane == 0 ) checksum[j] = cs;
    }
}

template <class G>
__device__ void
ticks_body(G& g, uint64_t first, uint64_t n, uint64_t* out)
{
    for ( uint64_t t = 0; t < first + n; ++t ) {
        double   nU  = g.uniform();
        uint32_t U32 = g.u32();
        uint64_t U64 = rng_u64(g);
        int32_t  I32 = rng_i32(g);
        int64_t  I64 = rng_i64(g);
        if ( t >= first ) {
            uint64_t* o = out + (t - first) * 5;
            o[0]        = (uint64_t)__double_as_longlong(nU);
            o[1]        = U32;
            o[2]        = U64;
            o[3]        = (uint64_t)(int64_t)I32;
            o[4]        = (uint64_t)I64;
        }
    }
}

__global__ void
rng_ticks_kernel(int kind, uint64_t s1, uint64_t s2, uint64_t first, uint64_t n, uint64_t* out, uint32_t* mt)
{
    if ( threadIdx.x || blockIdx.x ) return;
    if ( kind == 0 ) {
        mt_seed(mt, (uint32_t)s1);
        MersenneDev g { mt, 0 };
        ticks_body(g, first, n, out);
    }
    else if ( kind == 1 ) {
        MarsagliaDev g { (uint32_t)s1, (uint32_t)s2 };
        ticks_body(g, first, n, out);
    }
    else {
        XorShiftDev g;
        g.seed((uint32_t)s1);
        ticks_body(g, first, n, out);
    }
}

template <class G>
__device__ void
dist_body(G& g, int dist, const double* params, int np, uint64_t n, double* out)
{
    GaussianDev gs { np > 0 ? params[0] : 0.0, np > 1 ? params[1] : 1.0, 0.0, false };
    for ( uint64_t i = 0; i < n; ++i ) {
        double v;
        switch ( dist ) {
        case 0:
            v = dist_discrete(g, params, (uint32_t)np); // params = exclusive prefix sums (host-prepared)
            break;
        case 1:
            v = dist_exponential(g, params[0]);
            break;
        case 2:
            v = gs.next(g);
            break;
        case 3:
            v = dist_poisson(g, params[1]); // params[1] = exp(-lambda) from the host
            break;
        case 4:
            v = dist_uniform_bins(g, (uint32_t)params[0]);
            break;
        default:
            v = params[0];
        }
        out[i] = v;
    }
}

__global__ void
rng_dist_kernel(int dist, const double* params, int np, int kind, uint64_t s1, uint64_t s2, uint64_t n, double* out,
    uint32_t* mt)
{
    if ( threadIdx.x || blockIdx.x ) return;
    if ( kind == 0 ) {
        mt_seed(mt, (uint32_t)s1);
        MersenneDev g { mt, 0 };
        dist_body(g, dist, params, np, n, out);
    }
    else if ( kind == 1 ) {
        MarsagliaDev g { (uint32_t)s1, (uint32_t)s2 };
        dist_body(g, dist, params, np, n, out);
    }
    else {
        XorShiftDev g;
        g.seed((uint32_t)s1);
        dist_body(g, dist, params, np, n, out);
    }
}

// ================================================================================================ host side
template <class E>
constexpr ElementInfo
element_info()
{
    static_assert(E::TYPE > 0 && E::TYPE < TYPE_MAX, "element type codes are 1..31");
    return ElementInfo { E::TYPE, E::ELI_NAME, E::RECORD_BYTES, E::AUX_BYTES, E::HAS_PAYLOAD, E::PARALLEL_EVENTS,
        E::TIES_INTERCHANGEABLE, E::DESCRIPTION, prints_of<E>::value, self_links_of<E>::value, stat_clock_of<E>::value,
        ordered_of<E>::value, clock_form_of<E>::value, stats_in_aux_of<E>::value };
}
static const ElementInfo kElements[] = {
#define SSTB200_ELEMENT(E) element_info<E>(),
#include "elements.def"
#undef SSTB200_ELEMENT
};
constexpr int kNumElements = (int)(sizeof(kElements) / sizeof(kElements[0]));

} // namespace sstb200

using namespace sstb200;

// Peer arena (include/sstb200.h "peer arena"): one allocation per device of this process that other ranks map ONCE;
// engines place their peer-visible region at its start while they live.
struct PeerArena
{
    uint8_t* base     = nullptr;
    size_t   bytes    = 0;
    bool     in_use   = false;
    uint32_t rank = 0, n_ranks = 0;          // set by peer_arena_open
    uint8_t* mapped[MAX_RANKS] = {};         // the other ranks' arenas as mapped here (own entry = base)
};
static PeerArena g_arena[64];

struct sstb200_engine
{
    Dev                d {};
    int                device      = 0;
    cudaStream_t       stream      = nullptr;
    cudaStream_t       up_stream   = nullptr; // engine_create's second upload stream
    bool               own_stream  = false;
    std::vector<void*> allocs;
    Ctrl*              h_ctrl      = nullptr; // pinned mirror
    uint64_t*          d_results   = nullptr;
    size_t             smem_bytes  = 0;
    uint64_t           launches    = 0;
    bool               is_setup    = false;
    bool               report_pending = false; // between report_begin and report_end: dispatch uses the REPORT kernel
    bool               stats_in_aux   = false; // a local element type keeps its statistics in its aux block
    uint32_t           n_ranks     = 1;
    // device buffers that make up the simulation state (checkpoint sections, in a fixed order)
    struct Section
    {
        void*  ptr;
        size_t bytes;
    };
    std::vector<Section> sections;
    template <class T>
    void section(T* p, size_t count)
    {
        if ( p ) sections.push_back(Section { (void*)p, count * sizeof(T) });
    }

    template <class T>
    int alloc(T** p, size_t count, bool zero = true)
    {
        void*  q     = nullptr;
        size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        CUDA_OK(cudaMalloc(&q, bytes));
        if ( zero ) CUDA_OK(cudaMemsetAsync(q, 0, bytes, stream));
        allocs.push_back(q);
        *p = (T*)q;
        return 0;
    }
    template <class T>
    int upload(const T** p, const T* host, size_t count)
    {
        T* q = nullptr;
        if ( int rc = alloc(&q, count, false) ) return rc;
        if ( count ) CUDA_OK(cudaMemcpyAsync(q, host, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        *p = q;
        return 0;
    }
    int fetch_ctrl()
    {
        CUDA_OK(cudaMemcpyAsync(h_ctrl, d.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
        CUDA_OK(cudaStreamSynchronize(stream));
        return 0;
    }
    // event-parallel form (parallel_form.cuh): on when every local component is of one element type with
    // PARALLEL_EVENTS and nothing of the model needs what the form leaves out (see engine_create)
    bool     pf_on = false;
    uint32_t pf_ordered = 0; // element type whose ordered form (ordered_form.cuh) runs the windows, 0 = MessageMesh form
    uint32_t pf_clock   = 0; // element type whose clock form (clock_form.cuh) runs the windows
    size_t   pf_smem = 0;
    uint32_t pf_grid = 0, fb_grid = 0, maint_grid = 0;
    // one lookahead window: the event-parallel kernel, the sequential kernel over the blocks it left (or over all
    // blocks when the form is off, or for a window with a statistics report), generator maintenance
    int launch_dispatch(bool report = false)
    {
        if ( report ) {
            if ( d.has_payload )
                dispatch_kernel<true, true><<<d.nbins, BLOCK, smem_bytes, stream>>>(d);
            else
                dispatch_kernel<false, true><<<d.nbins, BLOCK, smem_bytes, stream>>>(d);
            launches++;
        }
        else if ( pf_on ) {
            launch_form();
            maintenance_kernel<<<maint_grid, 256, 0, stream>>>(d, 0, 0, 0, 0, 0); // the window is ended by the driver
            launches += 2;
        }
        else {
            // one element type on the rank and the model's payload setting is that element's own: the kernel compiled
            // for that element alone; else the kernel that switches on the type of every component
            const uint32_t ut       = d.uniform_type;
            bool           launched = false;
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( own_kernel_of<E>::value ) {                                                                   \
        if ( !launched && ut == E::TYPE && d.has_payload == (int)E::HAS_PAYLOAD ) {                               \
            dispatch_kernel<E::HAS_PAYLOAD, false, E><<<d.nbins, BLOCK, smem_bytes, stream>>>(d);                \
            launched = true;                                                                                     \
        }                                                                                                        \
    }
#include "elements.def"
#undef SSTB200_ELEMENT
            if ( launched ) {
            }
            else if ( d.has_payload )
                dispatch_kernel<true><<<d.nbins, BLOCK, smem_bytes, stream>>>(d);
            else
                dispatch_kernel<false><<<d.nbins, BLOCK, smem_bytes, stream>>>(d);
            launches++;
        }
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    // the native run loop with the event-parallel form: the window kernel (which also re-runs the blocks it could not
    // take, sequentially) + the maintenance kernel, whose last CTA ends the window -- two launches per window
    bool     fused_tail() const { return pf_on && (n_ranks == 1 || (d.peer_on && peer_spin)); }
    uint32_t launches_per_window() const { return fused_tail() ? 2u : (pf_on ? 3u : 2u) + (n_ranks > 1 && !(d.peer_on && peer_spin) ? 1u : 0u); }
    // the window kernel of the rank's form: MessageMesh form (pf_ordered 0) or the ordered form of element type pf_ordered
    void launch_form()
    {
        if ( pf_clock ) {
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( clock_form_of<E>::value ) {                                                                   \
        if ( pf_clock == E::TYPE ) window_clock_kernel<E><<<pf_grid, BLOCK, pf_smem, stream>>>(d);               \
    }
#include "elements.def"
#undef SSTB200_ELEMENT
            return;
        }
        if ( !pf_ordered ) {
            window_parallel_kernel<MeshElement><<<pf_grid, PF_THREADS, pf_smem, stream>>>(d);
            return;
        }
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( ordered_of<E>::value ) {                                                                      \
        if ( pf_ordered == E::TYPE ) window_ordered_kernel<E><<<pf_grid, OF_THREADS, pf_smem, stream>>>(d);      \
    }
#include "elements.def"
#undef SSTB200_ELEMENT
    }
    int      launch_window_fused()
    {
        launch_form();
        maintenance_kernel<<<maint_grid, 256, 0, stream>>>(d, 0, 0, 0, n_ranks > 1 ? 2 : 1, peer_spin ? 1 : 0);
        launches += 2;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    // `n` windows (dispatch + advance each).  The kernels take the same arguments every window -- the window state
    // lives in device memory -- so a chunk is captured once into a CUDA graph and replayed: small models are bound by
    // launch overhead, not by the kernels.
    cudaGraphExec_t graph_exec  = nullptr;
    uint32_t        graph_chunk = 0;
    bool            graph_failed = false;
    int launch_windows(uint32_t n, bool capture_only = false)
    {
        if ( !graph_failed && n >= 8 && (graph_exec == nullptr || graph_chunk == n) ) {
            if ( graph_exec == nullptr ) {
                cudaGraph_t g = nullptr;
                if ( cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed) == cudaSuccess ) {
                    const uint64_t before = launches;
                    int            rc     = 0;
                    for ( uint32_t i = 0; i < n && !rc; ++i ) {
                        if ( fused_tail() ) {
                            rc = launch_window_fused();
                            continue;
                        }
                        rc = launch_dispatch();
                        if ( !rc && n_ranks > 1 ) rc = launch_publish();
                        if ( !rc ) rc = launch_advance();
                    }
                    launches = before; // counted per replay below
                    cudaError_t ce = cudaStreamEndCapture(stream, &g);
                    if ( rc == 0 && ce == cudaSuccess && cudaGraphInstantiate(&graph_exec, g, 0) == cudaSuccess )
                        graph_chunk = n;
                    else
                        graph_failed = true;
                    if ( g ) cudaGraphDestroy(g);
                }
                else
                    graph_failed = true;
                (void)cudaGetLastError();
            }
            if ( capture_only ) return 0;
            if ( graph_exec && graph_chunk == n ) {
                CUDA_OK(cudaGraphLaunch(graph_exec, stream));
                launches += (uint64_t)launches_per_window() * n;
                return 0;
            }
        }
        if ( capture_only ) return 0;
        for ( uint32_t i = 0; i < n; ++i ) {
            if ( fused_tail() ) {
                if ( int rc = launch_window_fused() ) return rc;
                continue;
            }
            if ( int rc = launch_dispatch() ) return rc;
            if ( n_ranks > 1 )
                if ( int rc = launch_publish() ) return rc;
            if ( int rc = launch_advance() ) return rc;
        }
        return 0;
    }
    int launch_advance()
    {
        if ( d.peer_on )
            peer_advance_kernel<<<1, 32, 0, stream>>>(d, peer_spin ? 1 : 0, peer_spin ? 1 : 0);
        else
            advance_kernel<<<1, 32, 0, stream>>>(d);
        launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    // what follows this rank's window kernels in a multi-rank run: outbox mode -- the control words into the outbox
    // headers (the driver then exchanges the outboxes); peer mode -- the mailbox write to every rank
    int launch_publish()
    {
        if ( d.peer_on && peer_spin ) return 0; // posted by peer_advance_kernel itself
        if ( d.peer_on )
            peer_post_kernel<<<1, 32, 0, stream>>>(d);
        else
            publish_control_kernel<<<1, 64, 0, stream>>>(d);
        launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    // peer mode (sstb200_engine_peer_connect*): one device allocation holds everything other ranks write into --
    // calendar records, payloads, bin fill counts, mailbox -- so that one IPC handle per rank maps it all
    uint8_t*           region       = nullptr;
    size_t             region_bytes = 0;
    bool               region_in_arena = false;
    bool               peer_spin    = false; // peers are other processes / GPUs: peer_advance_kernel waits for them
    std::vector<void*> ipc_opened;
};

#ifdef SSTB200_TIMING
extern "C" int
sstb200_debug_phase_times(unsigned long long* out, uint32_t n_words)
{
    return cudaMemcpyFromSymbol(out, g_phase_times, (size_t)n_words * 8) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" {

int
sstb200_abi_version(void)
{
    return SSTB200_ABI_VERSION;
}

const char*
sstb200_last_error(void)
{
    return g_error.c_str();
}

int
sstb200_num_component_types(void)
{
    return kNumElements;
}

int
sstb200_component_type_info(int idx, const char** eli_name, uint32_t* type_code, uint32_t* state_bytes,
    uint32_t* aux_bytes, int* has_payload)
{
    if ( idx < 0 || idx >= kNumElements ) {
        set_error("component type index out of range");
        return -1;
    }
    if ( eli_name ) *eli_name = kElements[idx].eli_name;
    if ( type_code ) *type_code = kElements[idx].type;
    if ( state_bytes ) *state_bytes = kElements[idx].state_bytes;
    if ( aux_bytes ) *aux_bytes = kElements[idx].aux_bytes;
    if ( has_payload ) *has_payload = kElements[idx].has_payload;
    return 0;
}

int
sstb200_component_type_lookup(const char* eli_name)
{
    for ( int i = 0; i < kNumElements; ++i )
        if ( eli_name && !strcmp(eli_name, kElements[i].eli_name) ) return (int)kElements[i].type;
    return -1;
}

int
sstb200_engine_create(const sstb200_model* m, const sstb200_options* o, sstb200_engine** out)
{
    if ( !m || !o || !out ) {
        set_error("engine_create: null argument");
        return -1;
    }
    int ndev = 0;
    if ( cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 ) {
        set_error("engine_create: no CUDA device available (libsstb200 has no CPU fallback)");
        return -3;
    }
    if ( o->device < 0 || o->device >= ndev ) {
        set_error("engine_create: bad device ordinal");
        return -1;
    }
    const uint32_t n_ranks = o->n_ranks ? o->n_ranks : 1;
    const uint32_t c0      = n_ranks > 1 ? o->comp_begin : 0;
    const uint32_t c1      = n_ranks > 1 ? o->comp_end : m->ncomp;
    if ( c1 < c0 || c1 > m->ncomp || (n_ranks > 1 && !o->rank_comp_begin) ) {
        set_error("engine_create: bad component range");
        return -1;
    }
    CUDA_OK(cudaSetDevice(o->device));
    // SSTB200_CREATE_TIMING=1 prints where engine_create spends its time (host scans, uploads, allocations)
    const bool create_timing = getenv("SSTB200_CREATE_TIMING") != nullptr;
    auto       t_prev        = std::chrono::steady_clock::now();
    auto       lap           = [&](const char* what, cudaStream_t s) {
        if ( !create_timing ) return;
        if ( s ) cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[sstb200 create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    if ( const char* g = getenv("SSTB200_L2_FETCH") ) { // experiment knob: the L2's DRAM fetch granularity hint
        size_t            before = 0, after = 0;
        cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
        const cudaError_t rc = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[sstb200] L2 fetch granularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(rc));
    }
    sstb200_engine* e = new sstb200_engine();
    e->device         = o->device;
    e->n_ranks        = n_ranks;
    if ( o->stream ) {
        e->stream = (cudaStream_t)o->stream;
    }
    else {
        CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        e->own_stream = true;
    }
    Dev& d    = e->d;
    d.n_local = c1 - c0;
    d.c0      = c0;
    d.rank    = o->rank;
    d.n_ranks = n_ranks;
    d.nbins   = (d.n_local + BLOCK - 1) / BLOCK;
    if ( d.nbins == 0 ) d.nbins = 1;
    d.W       = o->calendar_slots >= 2 ? o->calendar_slots : 2;
    d.cap     = o->bin_capacity ? o->bin_capacity : 8 * BLOCK;
    if ( d.cap < BLOCK ) d.cap = BLOCK; // the dispatch prologue reads the first BLOCK slots of a bin unconditionally
    d.emax    = o->smem_events ? o->smem_events : 5 * BLOCK; // 1280: three CTAs per SM also when events carry payloads
    if ( d.emax > 32768 ) d.emax = 32768;
    d.emax = (d.emax + 3u) & ~3u;
    d.stats_on = m->stats_enabled;

    // Host scans run on a few threads.  Order of engine_create: (1) the component scan (element types, strides);
    // (2) the uploads start: component types and parameters on the engine's stream, followed there by the MT19937
    // seeding and first-generation kernels (they need nothing else), the port arrays and the link-row kernel on a second
    // stream; (3) WHILE those run, the port scan for the lookahead window on the host; (4) the remaining allocations.
    // engine_create returns when the caller's arrays have been consumed; the seeding kernels may still be running
    // (engine_setup and everything else is ordered behind them on the engine's stream).
    auto scan = [](size_t n, auto&& body) { // body(lo, hi, slot) on up to 16 threads
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned T  = n < (1u << 20) ? 1u : std::min(16u, hw);
        if ( T == 1 ) {
            body((size_t)0, n, 0u);
            return 1u;
        }
        std::vector<std::thread> th;
        for ( unsigned t = 0; t < T; ++t )
            th.emplace_back([&, t] { body(n * t / T, n * (t + 1) / T, t); });
        for ( auto& x : th )
            x.join();
        return T;
    };
    // per-rank slices
    d.P0          = m->port_off[c0];
    d.nport_local = m->port_off[c1] - d.P0;
    // one pass over the components: element types present (whole model: payload arrays are needed if ANY component of
    // any rank carries payloads; this rank: state/aux strides, forms), port-count uniformity, ports per block (link
    // rows of a block are staged in shared memory when they fit, 2048 rows = 32 KB; a uniform port count lets the
    // kernel find an event's component with one division instead of a binary search)
    const ElementInfo* by_type[TYPE_MAX] = {};
    for ( int i = 0; i < kNumElements; ++i )
        if ( kElements[i].type < TYPE_MAX ) by_type[kElements[i].type] = &kElements[i];
    struct Part
    {
        uint32_t seen_all = 0, seen_local = 0, bad = 0xFFFFFFFFu, max_blk = 0;
        bool     uni_ok = true, uni_all = true, clocks = false;
    } part[16];
    const uint32_t uni0 = d.n_local ? m->port_off[c0 + 1] - m->port_off[c0] : 0;
    // (a multi-rank driver that knows the model-wide facts passes them -- options.model_type_mask -- and every rank
    // scans its own slice only, like the reference's per-rank wire-up touches only local links, simulation.cc:668-859)
    const bool   hinted = n_ranks > 1 && o->model_type_mask != 0;
    const size_t sc_lo  = hinted ? c0 : 0, sc_hi = hinted ? c1 : m->ncomp;
    scan(sc_hi - sc_lo, [&](size_t lo_, size_t hi_, unsigned t) {
        Part         q;
        const size_t lo = sc_lo + lo_, hi = sc_lo + hi_;
        for ( size_t c = lo; c < hi; ++c ) {
            const uint32_t ty = m->comp_type[c];
            if ( m->port_off[c + 1] - m->port_off[c] != uni0 ) q.uni_all = false;
            if ( ty >= TYPE_MAX || !by_type[ty] ) {
                if ( c >= c0 && c < c1 && q.bad == 0xFFFFFFFFu ) q.bad = (uint32_t)c;
                continue;
            }
            q.seen_all |= 1u << ty;
            if ( c >= c0 && c < c1 ) {
                q.seen_local |= 1u << ty;
                if ( m->port_off[c + 1] - m->port_off[c] != uni0 ) q.uni_ok = false;
                if ( m->clock_period[c] ) q.clocks = true;
            }
        }
        part[t] = q;
    });
    scan(d.nbins, [&](size_t lo, size_t hi, unsigned t) {
        uint32_t mx = 0;
        for ( size_t b = lo; b < hi; ++b ) {
            const uint32_t blo = c0 + (uint32_t)b * BLOCK, bhi = std::min<uint32_t>(c1, blo + BLOCK);
            if ( blo < bhi ) mx = std::max(mx, m->port_off[bhi] - m->port_off[blo]);
        }
        part[t].max_blk = std::max(part[t].max_blk, mx);
    });
    uint32_t seen_all = 0, seen_local = 0, bad = 0xFFFFFFFFu, max_blk = 0;
    bool     uni_ok = true, uni_all = true, any_clock = false;
    for ( const Part& q : part ) {
        any_clock = any_clock || q.clocks;
        seen_all |= q.seen_all;
        seen_local |= q.seen_local;
        bad     = std::min(bad, q.bad);
        max_blk = std::max(max_blk, q.max_blk);
        uni_ok  = uni_ok && q.uni_ok;
        uni_all = uni_all && q.uni_all;
    }
    if ( hinted ) {
        seen_all = o->model_type_mask | seen_local;
        uni_all  = uni_ok && o->model_uniform_ports == 1;
    }
    if ( bad != 0xFFFFFFFFu ) {
        set_error("engine_create: component " + std::to_string(bad) + " has unknown type code " +
                  std::to_string(m->comp_type[bad]));
        delete e;
        return -1;
    }
    uint32_t state_stride = 8, aux_stride = 0;
    int      has_payload  = 0;
    bool any_parallel     = false;
    d.ties_free           = 1;
    d.uniform_type        = 0;
    for ( uint32_t ty = 0; ty < TYPE_MAX; ++ty ) {
        const ElementInfo* ei = by_type[ty];
        if ( !ei ) continue;
        if ( (seen_all >> ty & 1u) && ei->has_payload ) has_payload = 1;
        if ( !(seen_local >> ty & 1u) ) continue;
        state_stride = std::max(state_stride, ei->state_bytes);
        aux_stride   = std::max(aux_stride, ei->aux_bytes);
        if ( ei->parallel_events ) any_parallel = true;
        if ( ei->stats_in_aux ) e->stats_in_aux = true;
        if ( !ei->ties_interchangeable ) d.ties_free = 0;
        if ( (seen_local & (seen_local - 1)) == 0 ) d.uniform_type = ty; // the only type on this rank
    }
    bool ties_free_model = true;
    for ( uint32_t ty = 0; ty < TYPE_MAX; ++ty )
        if ( (seen_all >> ty & 1u) && by_type[ty] && !by_type[ty]->ties_interchangeable ) ties_free_model = false;
    d.ctl_free = (!any_clock && d.uniform_type && ties_free_model && o->trace_capacity == 0) ? 1u : 0u;
    d.par_max = (o->parallel_max_events && o->parallel_max_events < PAR_MAX_EVENTS) ? o->parallel_max_events : PAR_MAX_EVENTS;
    d.pf_rcap = PF_ROUNDS * PF_THREADS; // 1280 records: a 4096^2 mesh bin holds 1024 +- 32
    if ( const char* rc_env = getenv("SSTB200_PF_RCAP") ) d.pf_rcap = (uint32_t)atoi(rc_env);
    if ( d.pf_rcap > (uint32_t)PF_ROUNDS * PF_THREADS ) d.pf_rcap = PF_ROUNDS * PF_THREADS;
    if ( d.pf_rcap < 16 ) d.pf_rcap = 16;
    uint32_t global_shift = 32; // log2(ports per component) when EVERY component of the model has that many
    {
        const uint32_t uni = uni_ok ? uni0 : 0;
        d.lmax             = max_blk <= 2048 ? max_blk : 0;
        d.uniform_nports   = uni;
        d.uniform_shift    = 32;
        if ( uni && (uni & (uni - 1)) == 0 ) {
            d.uniform_shift = 0;
            while ( (1u << d.uniform_shift) < uni )
                ++d.uniform_shift;
            if ( uni_all ) global_shift = d.uniform_shift;
        }
    }
    d.state_stride = (state_stride + 15u) & ~15u;
    // elements whose hot part is the head of a longer record (statistics behind it): records start on 64-byte
    // boundaries, so that the 32/64 hot bytes of a component are one DRAM burst and not the tails of two.  The forms
    // that stage whole records with one bulk copy keep them packed.
    {
        bool packed = false;
        for ( uint32_t ty = 0; ty < TYPE_MAX; ++ty )
            if ( (seen_local >> ty & 1u) && by_type[ty] && (by_type[ty]->parallel_events || by_type[ty]->ordered_events) )
                packed = true;
        if ( !packed && d.state_stride > 64 ) d.state_stride = (d.state_stride + 63u) & ~63u;
    }
    d.aux_stride   = (aux_stride + 31u) & ~31u;
    d.has_payload  = has_payload;
    lap("element scans", nullptr);

    int rc = 0;
    CUDA_OK(cudaStreamCreateWithFlags(&e->up_stream, cudaStreamNonBlocking));
    cudaEvent_t ev_up = nullptr;
    CUDA_OK(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
    // (2) Every host->device copy goes through the second stream, back to back, so that PCIe never idles: component
    // types, the parameters in a few chunks, then the port arrays (+ the link-row kernel).  The engine's stream clears
    // the element state and, chunk by chunk as the parameters arrive, seeds the generators and builds their first
    // generation: the generators of the first chunk are being built while the rest of the model is still crossing PCIe.
    {
        cudaStream_t main_stream = e->stream;
        e->stream                = e->up_stream; // alloc()/upload() below work on the second stream
        rc |= e->upload(&d.comp_type, m->comp_type + c0, d.n_local);
        int64_t* param_dev = nullptr;
        rc |= e->alloc(&param_dev, (size_t)d.n_local * NPARAM, false);
        d.comp_param = param_dev;
        struct Chunk
        {
            uint32_t    lo, hi;
            cudaEvent_t ev;
        };
        std::vector<Chunk> chunks;
        const uint32_t     n_chunks = d.n_local >= (1u << 20) ? 8u : 1u;
        for ( uint32_t ch = 0; ch < n_chunks && !rc; ++ch ) {
            const uint32_t lo = (uint32_t)((uint64_t)d.n_local * ch / n_chunks) & ~127u;
            const uint32_t hi = ch + 1 == n_chunks ? d.n_local : (uint32_t)((uint64_t)d.n_local * (ch + 1) / n_chunks) & ~127u;
            if ( hi <= lo ) continue;
            CUDA_OK(cudaMemcpyAsync(param_dev + (size_t)lo * NPARAM, m->comp_param + ((size_t)c0 + lo) * NPARAM,
                (size_t)(hi - lo) * NPARAM * sizeof(int64_t), cudaMemcpyHostToDevice, e->up_stream));
            Chunk c { lo, hi, nullptr };
            CUDA_OK(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
            CUDA_OK(cudaEventRecord(c.ev, e->up_stream));
            chunks.push_back(c);
        }
        // the port arrays; link rows are built on the device from the rank's slice of peer_port / latency (and the
        // global port CSR when port counts differ between components)
        rc |= e->upload(&d.port_off, m->port_off + c0, (size_t)d.n_local + 1);
        uint4*          link_dev = nullptr;
        const uint32_t* peer_dev = nullptr;
        const uint64_t* lat_dev  = nullptr;
        const uint32_t* poff_dev = nullptr;
        rc |= e->alloc(&link_dev, d.nport_local, false);
        rc |= e->upload(&peer_dev, m->peer_port + d.P0, d.nport_local);
        rc |= e->upload(&lat_dev, m->latency + d.P0, d.nport_local);
        if ( global_shift >= 32 ) rc |= e->upload(&poff_dev, m->port_off, (size_t)m->ncomp + 1);
        if ( !rc && d.nport_local ) {
            build_links_kernel<<<148 * 8, BLOCK, 0, e->stream>>>(link_dev, peer_dev, lat_dev, poff_dev, m->ncomp,
                d.nport_local, global_shift);
            e->launches++;
        }
        d.link = link_dev;
        rc |= e->upload(&d.clock_period, m->clock_period + c0, d.n_local);
        if ( m->xparam_off && m->xparam ) {
            rc |= e->upload(&d.xparam_off, m->xparam_off + c0, (size_t)d.n_local + 1);
            rc |= e->upload(&d.xparam, m->xparam + m->xparam_off[c0], (size_t)(m->xparam_off[c1] - m->xparam_off[c0]));
        }
        if ( n_ranks > 1 ) rc |= e->upload(&d.rank_comp_begin, o->rank_comp_begin, (size_t)n_ranks + 1);
        CUDA_OK(cudaEventRecord(ev_up, e->stream));
        e->stream = main_stream;
        // every copy is queued: PCIe is busy from here on.  Now the big allocations (tens of milliseconds for the 80 GB
        // of generator state of a 16.7 M-component mesh), then the generators, chunk by chunk as their parameters arrive
        rc |= e->alloc(&d.state, (size_t)d.n_local * d.state_stride);
        rc |= e->alloc(&d.aux, (size_t)d.n_local * d.aux_stride, false);
        int n_sm0 = 148;
        cudaDeviceGetAttribute(&n_sm0, cudaDevAttrMultiProcessorCount, o->device);
        for ( const Chunk& c : chunks ) {
            if ( !rc ) CUDA_OK(cudaStreamWaitEvent(main_stream, c.ev, 0));
            cudaEventDestroy(c.ev); // released once the wait has been satisfied
            if ( !rc && d.aux_stride ) {
                seed_mt_kernel<<<(c.hi - c.lo + 127) / 128, 128, 0, main_stream>>>(d, c.lo, c.hi);
                // first MT19937 generation (and route digest) of every generator seeded above
                maintenance_kernel<<<(uint32_t)n_sm0 * 8, 256, 0, main_stream>>>(d, 1, c.lo, c.hi, 0, 0);
                e->launches += 2;
            }
        }
    }
    lap("uploads issued", nullptr);

    // (3) lookahead window = minimum latency over every connected link (cut or not: a window is processed as one batch
    // per component, so nothing sent inside it may be due inside it).  A multi-rank driver that passes the window (the
    // minimum over ALL ranks, which it can get from one MIN all-reduce) lets every rank scan its own ports only.
    const uint32_t nport   = m->port_off[m->ncomp];
    const size_t   scan_lo = o->window ? d.P0 : 0, scan_hi = o->window ? (size_t)d.P0 + d.nport_local : nport;
    uint64_t       mn_lat[16], mx_lat[16];
    for ( auto& v : mn_lat )
        v = TIME_NEVER;
    for ( auto& v : mx_lat )
        v = 0;
    scan(scan_hi - scan_lo, [&](size_t lo, size_t hi, unsigned t) {
        uint64_t mn = TIME_NEVER, mx = 0;
        for ( size_t p = scan_lo + lo; p < scan_lo + hi; ++p )
            if ( m->peer_port[p] != SSTB200_NO_PEER && !(m->port_self && m->port_self[p]) ) {
                if ( m->latency[p] < mn ) mn = m->latency[p];
                if ( m->latency[p] > mx ) mx = m->latency[p];
            }
        mn_lat[t] = mn;
        mx_lat[t] = mx;
    });
    const uint64_t mn         = *std::min_element(mn_lat, mn_lat + 16);
    const uint64_t mx_lat_all = *std::max_element(mx_lat, mx_lat + 16);
    uint64_t       w          = o->window ? o->window : (mn == TIME_NEVER ? 1000000ull : mn);
    if ( w == 0 || w > 0x7fffffffull ) {
        set_error("engine_create: lookahead window must be in [1, 2^31) cycles (zero-latency links are not allowed)");
        sstb200_engine_destroy(e);
        return -1;
    }
    if ( mn < w ) {
        set_error("engine_create: window larger than a link latency");
        sstb200_engine_destroy(e);
        return -1;
    }
    lap("window scan", nullptr);
    // The event-parallel form runs the windows of this rank when: one local element type, and it has the form; nobody
    // in the model looks at send sequence numbers, clocks or traces (ctl_free); a power-of-two port count per
    // component whose link rows fit the staging buffer; no payloads; no handler counting; link latencies below 2^31.
    // Blocks it cannot take in some window (see parallel_form.cuh) are run by the sequential form in that window.
    e->pf_on = any_parallel && d.uniform_type == TYPE_MESH && d.ctl_free && uni_ok && uni0 && (uni0 & (uni0 - 1)) == 0 &&
               (size_t)uni0 * BLOCK <= 2048 && !has_payload && !o->handler_counts && mx_lat_all < 0x7fffffffull &&
               getenv("SSTB200_NO_PARALLEL_FORM") == nullptr;
    // The ordered event-parallel form (ordered_form.cuh): one local element type that declares ORDERED_EVENTS and whose
    // record is the rank's state stride; no local clocks; the model's payload setting is the element's; no tracing, no
    // handler counting; a power-of-two port count per component; an even bin capacity (8-byte payloads move in 16-byte
    // pieces); link latencies below 2^31.
    if ( !e->pf_on && d.uniform_type && by_type[d.uniform_type] && by_type[d.uniform_type]->ordered_events && !any_clock &&
         has_payload == by_type[d.uniform_type]->has_payload && has_payload && o->trace_capacity == 0 && !o->handler_counts &&
         uni_ok && uni0 && (uni0 & (uni0 - 1)) == 0 && (size_t)uni0 * BLOCK <= 2048 && mx_lat_all < 0x7fffffffull &&
         d.state_stride == by_type[d.uniform_type]->state_bytes && (d.cap & 1u) == 0 && (w * uni0) < (1ull << 32) &&
         getenv("SSTB200_NO_PARALLEL_FORM") == nullptr ) {
        e->pf_on      = true;
        e->pf_ordered = d.uniform_type;
        d.pf_rcap &= ~1u;
    }
    // The clock form (clock_form.cuh): one local element type that declares CLOCK_FORM; the model's payload setting is
    // the element's; no tracing, no handler counting; a power-of-two port count per component.
    if ( !e->pf_on && d.uniform_type && by_type[d.uniform_type] && by_type[d.uniform_type]->clock_form &&
         has_payload == by_type[d.uniform_type]->has_payload && o->trace_capacity == 0 && !o->handler_counts && uni_ok &&
         uni0 && (uni0 & (uni0 - 1)) == 0 && getenv("SSTB200_NO_PARALLEL_FORM") == nullptr ) {
        e->pf_on    = true;
        e->pf_clock = d.uniform_type;
    }

    // (4) the remaining allocations (engine stream)
    rc |= e->alloc(&d.ctl, d.n_local, false); // every record is written by setup_kernel
    const size_t nev = (size_t)d.W * d.nbins * d.cap;
    {
        // one allocation for what other ranks may write into (see sstb200_engine::region): records | payloads | bin
        // fill counts | mailbox; only the counts and the mailbox are cleared
        const size_t ev_b  = nev * sizeof(ulonglong2), pay_b = has_payload ? nev * 8 : 0;
        const size_t cnt_b = (((size_t)d.W * d.nbins * 4) + 255) & ~(size_t)255;
        const size_t mail_b = (size_t)2 * MAX_RANKS * MAILW * 8;
        e->region_bytes     = ev_b + pay_b + cnt_b + mail_b;
        PeerArena& ar       = g_arena[o->device & 63];
        if ( n_ranks > 1 && ar.base && !ar.in_use && ar.n_ranks == n_ranks && ar.rank == o->rank &&
             e->region_bytes <= ar.bytes ) {
            e->region          = ar.base;
            e->region_in_arena = true;
            ar.in_use          = true;
        }
        else
            rc |= e->alloc(&e->region, e->region_bytes, false);
        if ( !rc ) {
            d.ev         = reinterpret_cast<ulonglong2*>(e->region);
            d.ev_payload = has_payload ? reinterpret_cast<uint64_t*>(e->region + ev_b) : nullptr;
            d.bin_count  = reinterpret_cast<uint32_t*>(e->region + ev_b + pay_b);
            d.mail       = reinterpret_cast<unsigned long long*>(e->region + ev_b + pay_b + cnt_b);
            CUDA_OK(cudaMemsetAsync(d.bin_count, 0, cnt_b + mail_b, e->stream));
        }
    }
    rc |= e->alloc(&d.ctrl, 1);
    d.outbox_cap = o->outbox_capacity ? o->outbox_capacity : 65536;
    d.xwords     = XHDR + (uint64_t)d.outbox_cap * 4;
    if ( n_ranks > 1 ) {
        rc |= e->alloc(&d.outbox, (size_t)n_ranks * d.xwords);
        rc |= e->alloc(&d.inbox, (size_t)n_ranks * d.xwords);
        rc |= e->alloc(&d.gathered, (size_t)n_ranks * CW_WORDS);
    }
    d.trace_cap = o->trace_capacity;
    if ( d.trace_cap ) {
        rc |= e->alloc(&d.tr_time, d.trace_cap);
        rc |= e->alloc(&d.tr_comp, d.trace_cap);
        rc |= e->alloc(&d.tr_port, d.trace_cap);
        rc |= e->alloc(&d.tr_idx, d.trace_cap);
        rc |= e->alloc(&d.tr_payload, d.trace_cap);
    }
    {
        bool prints = false;
        for ( uint32_t ty = 0; ty < TYPE_MAX; ++ty )
            if ( (seen_local >> ty & 1u) && by_type[ty] && by_type[ty]->prints ) prints = true;
        d.out_cap = o->output_capacity ? o->output_capacity : (prints ? 65536 : 0);
        if ( d.out_cap ) rc |= e->alloc(&d.out_log, (size_t)d.out_cap * 8, false);
    }
    if ( o->handler_counts ) {
        rc |= e->alloc(&d.prof_recv, d.nport_local);
        rc |= e->alloc(&d.prof_clock, d.n_local);
    }
    if ( o->handler_counts & 2u ) { // ... and the time inside the handlers
        rc |= e->alloc(&d.prof_recv_cyc, d.nport_local);
        rc |= e->alloc(&d.prof_clock_cyc, d.n_local);
    }
    if ( e->pf_on ) {
        rc |= e->alloc(&d.fb_list, d.nbins); // zeroed: an entry is block + 1, 0 = not written yet
        rc |= e->alloc(&d.maint_list, d.n_local, false);
    }
    rc |= e->alloc(&e->d_results, (size_t)d.n_local * NRESULT, false); // results_kernel writes every word it exports
    lap("allocations", nullptr);
    // the caller's arrays have been consumed once both upload batches are through; whatever follows on the engine's
    // stream is ordered behind the second stream's work
    if ( !rc ) {
        CUDA_OK(cudaStreamWaitEvent(e->stream, ev_up, 0));
        CUDA_OK(cudaEventSynchronize(ev_up));
    }
    cudaEventDestroy(ev_up);
    lap("uploads done", nullptr);
    if ( rc ) {
        sstb200_engine_destroy(e);
        return -2;
    }
    // everything a checkpoint has to carry (Simulation::checkpoint simulation.cc:2015-2071 serialises the same things:
    // component state, the TimeVortex contents, clocks, statistics, sync queues)
    e->section(d.ctrl, 1);
    e->section(d.ctl, d.n_local);
    e->section(d.state, (size_t)d.n_local * d.state_stride);
    e->section(d.aux, (size_t)d.n_local * d.aux_stride);
    e->section(d.bin_count, (size_t)d.W * d.nbins);
    e->section(d.ev, nev);
    e->section(d.ev_payload, has_payload ? nev : 0);
    if ( n_ranks > 1 ) {
        e->section(d.outbox, (size_t)n_ranks * d.xwords);
        e->section(d.inbox, (size_t)n_ranks * d.xwords);
        e->section(d.gathered, (size_t)n_ranks * CW_WORDS);
    }
    e->section(d.prof_recv, d.nport_local);
    e->section(d.prof_clock, d.n_local);
    CUDA_OK(cudaMallocHost((void**)&e->h_ctrl, sizeof(Ctrl)));
    memset(e->h_ctrl, 0, sizeof(Ctrl));
    e->h_ctrl->w       = w;
    d.w_magic          = w == 1 ? 0 : ~0ull / w + 1; // ceil(2^64 / w): exact quotients for 32-bit dividends
    d.W_mask           = (d.W & (d.W - 1)) == 0 ? d.W - 1 : 0;
    e->h_ctrl->stop_at = m->stop_at;
    e->h_ctrl->k       = 0;
    e->h_ctrl->T       = 0;
    e->h_ctrl->T_end   = (m->stop_at && w > m->stop_at) ? m->stop_at : w;
    e->h_ctrl->had_primary = 1;
    e->h_ctrl->report_at   = TIME_NEVER;
    CUDA_OK(cudaMemcpyAsync(d.ctrl, e->h_ctrl, sizeof(Ctrl), cudaMemcpyHostToDevice, e->stream));

    e->smem_bytes = (size_t)d.lmax * 16 + (size_t)d.emax * (has_payload ? 8 : 0) + (size_t)d.emax * 16 +
                    (size_t)(BLOCK + 1 + BLOCK + BLOCK + BLOCK / 32 + 1 + 4 + 32 + 32) * 4 + (size_t)d.emax * 2 +
                    (size_t)BLOCK * 2 + 16;
    if ( e->smem_bytes > 226 * 1024 ) {
        set_error("engine_create: smem_events too large for 227 KB of shared memory");
        sstb200_engine_destroy(e);
        return -1;
    }
    CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( own_kernel_of<E>::value )                                                                     \
        CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<E::HAS_PAYLOAD, false, E>,                                  \
            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#include "elements.def"
#undef SSTB200_ELEMENT
    CUDA_OK(cudaFuncSetAttribute(dispatch_kernel<false, false, MeshElement, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int n_sm = 148;
    CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, o->device));
    e->maint_grid = (uint32_t)n_sm * 8;
    if ( e->pf_on ) {
        // persistent grid: as many CTAs as fit, each walks blocks b, b + grid, ...
        int per_sm = 0;
        if ( e->pf_clock ) {
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( clock_form_of<E>::value ) {                                                                   \
        if ( e->pf_clock == E::TYPE ) {                                                                          \
            e->pf_smem = e->smem_bytes;                                                                          \
            CUDA_OK(cudaFuncSetAttribute(window_clock_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                226 * 1024)); /* process-wide: engines of several sizes share it; 1 KB for its static arrays */ \
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_clock_kernel<E>, BLOCK,        \
                e->pf_smem));                                                                                    \
        }                                                                                                        \
    }
#include "elements.def"
#undef SSTB200_ELEMENT
        }
        else if ( e->pf_ordered ) {
#define SSTB200_ELEMENT(E)                                                                                       \
    if constexpr ( ordered_of<E>::value ) {                                                                      \
        if ( e->pf_ordered == E::TYPE ) {                                                                        \
            e->pf_smem = std::max(of_smem_bytes<E>(d.pf_rcap, d.uniform_shift), e->smem_bytes);                  \
            CUDA_OK(cudaFuncSetAttribute(window_ordered_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                227 * 1024));                                                                                    \
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_ordered_kernel<E>, OF_THREADS, \
                e->pf_smem));                                                                                    \
        }                                                                                                        \
    }
#include "elements.def"
#undef SSTB200_ELEMENT
        }
        else {
        e->pf_smem = (size_t)BLOCK * MeshElement::RECORD_BYTES + ((size_t)BLOCK * uni0) * 16 + (size_t)d.pf_rcap * 16 +
                     (size_t)(2 * BLOCK + 3 * PF_HT + 4) * 4 + 3 * 8 + 16;
        CUDA_OK(cudaFuncSetAttribute(window_parallel_kernel<MeshElement>, cudaFuncAttributeMaxDynamicSharedMemorySize,
            227 * 1024));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, window_parallel_kernel<MeshElement>, PF_THREADS,
            e->pf_smem));
        }
        if ( per_sm < 1 ) per_sm = 1;
        if ( const char* g = getenv("SSTB200_PF_CTAS") ) per_sm = std::max(1, atoi(g));
        e->pf_grid = std::min<uint32_t>(d.nbins, (uint32_t)(per_sm * n_sm));
        e->fb_grid = std::min<uint32_t>(d.nbins, (uint32_t)(2 * n_sm));
        if ( e->pf_smem < e->smem_bytes ) e->pf_smem = e->smem_bytes; // the kernel re-runs blocks in the sequential form
    }
    // (the Ctrl mirror is pinned memory of the engine: later copies on the same stream are ordered behind this one)
    if ( create_timing ) CUDA_OK(cudaStreamSynchronize(e->stream));
    lap("generators seeded (waited for only when timing)", nullptr);
    *out = e;
    return 0;
}

void
sstb200_engine_destroy(sstb200_engine* e)
{
    if ( !e ) return;
    cudaSetDevice(e->device);
    if ( e->stream ) cudaStreamSynchronize(e->stream);
    for ( void* p : e->ipc_opened )
        cudaIpcCloseMemHandle(p);
    if ( e->region_in_arena ) g_arena[e->device & 63].in_use = false;
    for ( void* p : e->allocs )
        cudaFree(p);
    if ( e->graph_exec ) cudaGraphExecDestroy(e->graph_exec);
    if ( e->h_ctrl ) cudaFreeHost(e->h_ctrl);
    if ( e->up_stream ) {
        cudaStreamSynchronize(e->up_stream);
        cudaStreamDestroy(e->up_stream);
    }
    if ( e->own_stream && e->stream ) cudaStreamDestroy(e->stream);
    delete e;
}

int
sstb200_engine_setup(sstb200_engine* e)
{
    if ( !e ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( e->is_setup ) {
        set_error("engine_setup called twice");
        return -1;
    }
    // (the generators were seeded by engine_create, on this stream)
    setup_kernel<<<e->d.nbins, BLOCK, 0, e->stream>>>(e->d);
    e->launches++;
    CUDA_OK(cudaGetLastError());
    e->is_setup = true;
    return 0;
}

int
sstb200_engine_prepare_run(sstb200_engine* e, uint32_t check_every)
{
    if ( !e ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( !e->is_setup ) {
        set_error("engine_prepare_run before engine_setup");
        return -1;
    }
    if ( e->n_ranks > 1 && !(e->d.peer_on && e->peer_spin) ) return 0; // stepped by the driver: nothing to capture
    return e->launch_windows(check_every ? check_every : 64, true);
}

int
sstb200_engine_run(sstb200_engine* e, uint64_t max_windows, uint32_t check_every, int* finished)
{
    if ( !e ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( !e->is_setup ) {
        set_error("engine_run before engine_setup");
        return -1;
    }
    if ( e->n_ranks > 1 && !(e->d.peer_on && e->peer_spin) ) {
        set_error("engine_run needs a single rank or ranks connected in peer mode (sstb200_engine_peer_connect); "
                  "otherwise step the ranks with dispatch / exchange / ingest / advance");
        return -1;
    }
    if ( !check_every ) check_every = 64;
    uint64_t done_windows = 0;
    if ( finished ) *finished = 0;
    for ( ;; ) {
        uint32_t chunk = check_every;
        if ( max_windows && max_windows - done_windows < chunk ) chunk = (uint32_t)(max_windows - done_windows);
        if ( int rc = e->launch_windows(chunk) ) return rc;
        done_windows += chunk;
        if ( int rc = e->fetch_ctrl() ) return rc;
        if ( e->h_ctrl->error ) {
            char buf[256];
            snprintf(buf, sizeof buf, "engine error %u at window %llu (info %llu %llu %llu)", e->h_ctrl->error,
                (unsigned long long)e->h_ctrl->error_info[3], (unsigned long long)e->h_ctrl->error_info[0],
                (unsigned long long)e->h_ctrl->error_info[1], (unsigned long long)e->h_ctrl->error_info[2]);
            set_error(buf);
            if ( finished ) *finished = 1;
            return -4;
        }
        if ( e->h_ctrl->done ) {
            if ( finished ) *finished = 1;
            return 0;
        }
        if ( max_windows && done_windows >= max_windows ) return 0;
    }
}

int
sstb200_engine_dispatch(sstb200_engine* e)
{
    if ( !e || !e->is_setup ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->launch_dispatch(e->report_pending) ) return rc;
    return e->launch_publish();
}

int
sstb200_engine_ingest(sstb200_engine* e)
{
    if ( !e || !e->is_setup ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( e->n_ranks > 1 && !e->d.peer_on ) { // peer mode: the records are already in the calendar
        ingest_kernel<<<64, BLOCK, 0, e->stream>>>(e->d);
        e->launches++;
        CUDA_OK(cudaGetLastError());
        collect_control_kernel<<<1, 64, 0, e->stream>>>(e->d);
        e->launches++;
        CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int
sstb200_engine_advance(sstb200_engine* e)
{
    if ( !e || !e->is_setup ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    return e->launch_advance();
}

// ---- peer mode ----------------------------------------------------------------------------------------------------
namespace {
// where rank r's buffers lie inside its region, from the geometry every rank can compute (W, cap, payload setting are
// the same on all ranks; the number of component blocks differs)
void
peer_layout(const Dev& d, uint32_t nbins_r, uint8_t* base, uint32_t r, Dev& out)
{
    const size_t nev   = (size_t)d.W * nbins_r * d.cap;
    const size_t ev_b  = nev * sizeof(ulonglong2), pay_b = d.has_payload ? nev * 8 : 0;
    const size_t cnt_b = (((size_t)d.W * nbins_r * 4) + 255) & ~(size_t)255;
    out.peer_ev[r]        = reinterpret_cast<ulonglong2*>(base);
    out.peer_pay[r]       = d.has_payload ? reinterpret_cast<uint64_t*>(base + ev_b) : nullptr;
    out.peer_bin_count[r] = reinterpret_cast<uint32_t*>(base + ev_b + pay_b);
    out.peer_mail[r]      = reinterpret_cast<unsigned long long*>(base + ev_b + pay_b + cnt_b);
    out.peer_nbins[r]     = nbins_r;
}
int
peer_check(sstb200_engine* e, const char* what)
{
    if ( !e ) return -1;
    if ( e->n_ranks < 2 || e->n_ranks > MAX_RANKS ) {
        set_error(std::string(what) + ": peer mode needs 2.." + std::to_string(MAX_RANKS) + " ranks");
        return -1;
    }
    if ( e->is_setup ) {
        set_error(std::string(what) + ": connect the ranks before engine_setup (setup() already sends events)");
        return -1;
    }
    return 0;
}
} // namespace

int
sstb200_engine_peer_export(sstb200_engine* e, void* handle64, uint32_t* nbins, void** region_dev)
{
    if ( !e ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    {
        const auto t0 = std::chrono::steady_clock::now();
        CUDA_OK(cudaStreamSynchronize(e->stream)); // the region has been cleared before anybody may write into it
        if ( getenv("SSTB200_CREATE_TIMING") )
            fprintf(stderr, "[sstb200 peer] rank %u export: waited %.3f ms for the engine's stream\n", e->d.rank,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    if ( handle64 ) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, e->region));
        memcpy(handle64, &h, 64);
    }
    if ( nbins ) *nbins = e->d.nbins;
    if ( region_dev ) *region_dev = e->region;
    return 0;
}

int
sstb200_engine_peer_connect(sstb200_engine* e, const void* handles, const uint32_t* nbins)
{
    if ( int rc = peer_check(e, "engine_peer_connect") ) return rc;
    if ( !handles || !nbins ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    Dev&       d      = e->d;
    const bool timing = getenv("SSTB200_CREATE_TIMING") != nullptr;
    for ( uint32_t r = 0; r < e->n_ranks; ++r ) {
        uint8_t* base = e->region;
        if ( r != d.rank ) {
            const auto         t0 = std::chrono::steady_clock::now();
            cudaIpcMemHandle_t h;
            memcpy(&h, (const uint8_t*)handles + (size_t)r * 64, 64);
            void* p = nullptr;
            CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            e->ipc_opened.push_back(p);
            base = (uint8_t*)p;
            if ( timing )
                fprintf(stderr, "[sstb200 peer] rank %u opens rank %u's region: %.3f ms\n", d.rank, r,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
        peer_layout(d, nbins[r], base, r, d);
    }
    d.peer_on    = 1;
    e->peer_spin = true;
    return 0;
}

uint64_t
sstb200_peer_region_bytes(uint32_t n_local, uint32_t calendar_slots, uint32_t bin_capacity, int has_payload)
{
    const size_t W = calendar_slots >= 2 ? calendar_slots : 2, nbins = std::max<size_t>(1, ((size_t)n_local + BLOCK - 1) / BLOCK);
    const size_t cap = std::max<size_t>(bin_capacity ? bin_capacity : 8 * BLOCK, BLOCK);
    const size_t nev = W * nbins * cap;
    return nev * sizeof(ulonglong2) + (has_payload ? nev * 8 : 0) + (((W * nbins * 4) + 255) & ~(size_t)255) +
           (size_t)2 * MAX_RANKS * MAILW * 8;
}

int
sstb200_peer_arena_create(int device, uint64_t bytes, void* handle64_out)
{
    if ( device < 0 || device >= 64 || !bytes ) return -1;
    CUDA_OK(cudaSetDevice(device));
    PeerArena& ar = g_arena[device];
    if ( ar.base ) {
        set_error("peer_arena_create: this device already has an arena (destroy it first)");
        return -1;
    }
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, bytes));
    ar       = PeerArena {};
    ar.base  = (uint8_t*)p;
    ar.bytes = bytes;
    if ( handle64_out ) {
        cudaIpcMemHandle_t h;
        CUDA_OK(cudaIpcGetMemHandle(&h, p));
        memcpy(handle64_out, &h, 64);
    }
    return 0;
}

int
sstb200_peer_arena_open(int device, uint32_t rank, uint32_t n_ranks, const void* handles)
{
    if ( device < 0 || device >= 64 || !handles || n_ranks < 2 || n_ranks > MAX_RANKS || rank >= n_ranks ) return -1;
    CUDA_OK(cudaSetDevice(device));
    PeerArena& ar = g_arena[device];
    if ( !ar.base || ar.n_ranks ) {
        set_error("peer_arena_open: create the arena first, open it once");
        return -1;
    }
    for ( uint32_t r = 0; r < n_ranks; ++r ) {
        if ( r == rank ) {
            ar.mapped[r] = ar.base;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t*)handles + (size_t)r * 64, 64);
        void* p = nullptr;
        CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ar.mapped[r] = (uint8_t*)p;
    }
    ar.rank    = rank;
    ar.n_ranks = n_ranks;
    return 0;
}

int
sstb200_peer_arena_destroy(int device)
{
    if ( device < 0 || device >= 64 ) return -1;
    PeerArena& ar = g_arena[device];
    if ( !ar.base ) return 0;
    if ( ar.in_use ) {
        set_error("peer_arena_destroy: an engine still lives in the arena");
        return -1;
    }
    CUDA_OK(cudaSetDevice(device));
    for ( uint32_t r = 0; r < ar.n_ranks; ++r )
        if ( ar.mapped[r] && ar.mapped[r] != ar.base ) cudaIpcCloseMemHandle(ar.mapped[r]);
    cudaFree(ar.base);
    ar = PeerArena {};
    return 0;
}

int
sstb200_engine_in_arena(sstb200_engine* e, int* in_arena)
{
    if ( !e || !in_arena ) return -1;
    *in_arena = e->region_in_arena ? 1 : 0;
    return 0;
}

// every rank's engine lives at the start of its arena (sstb200_engine_in_arena said so on ALL ranks): nothing to map
int
sstb200_engine_peer_connect_arena(sstb200_engine* e, const uint32_t* nbins)
{
    if ( int rc = peer_check(e, "engine_peer_connect_arena") ) return rc;
    PeerArena& ar = g_arena[e->device & 63];
    if ( !nbins || !e->region_in_arena || ar.n_ranks != e->n_ranks ) {
        set_error("engine_peer_connect_arena: this engine's region is not in an opened arena");
        return -1;
    }
    Dev& d = e->d;
    for ( uint32_t r = 0; r < e->n_ranks; ++r )
        peer_layout(d, nbins[r], ar.mapped[r], r, d);
    d.peer_on    = 1;
    e->peer_spin = true;
    return 0;
}

int
sstb200_engine_peer_connect_local(sstb200_engine* e, void* const* regions, const uint32_t* nbins)
{
    if ( int rc = peer_check(e, "engine_peer_connect_local") ) return rc;
    if ( !regions || !nbins ) return -1;
    Dev& d = e->d;
    for ( uint32_t r = 0; r < e->n_ranks; ++r )
        peer_layout(d, nbins[r], (uint8_t*)regions[r], r, d);
    d.peer_on    = 1;
    e->peer_spin = false; // the ranks are stepped one after the other on one GPU: nothing to wait for
    return 0;
}

int
sstb200_engine_output(sstb200_engine* e, uint64_t* n_inout, uint64_t* records)
{
    if ( !e || !n_inout ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->fetch_ctrl() ) return rc;
    const uint64_t n = e->h_ctrl->out_n < e->d.out_cap ? e->h_ctrl->out_n : e->d.out_cap;
    if ( !records ) {
        *n_inout = n;
        return 0;
    }
    if ( *n_inout < n ) {
        set_error("engine_output: buffer too small");
        return -1;
    }
    if ( n ) {
        CUDA_OK(cudaMemcpyAsync(records, e->d.out_log, (size_t)n * 64, cudaMemcpyDeviceToHost, e->stream));
        CUDA_OK(cudaStreamSynchronize(e->stream));
    }
    *n_inout = n;
    return 0;
}

int
sstb200_engine_stream(sstb200_engine* e, void** stream_out)
{
    if ( !e || !stream_out ) return -1;
    *stream_out = (void*)e->stream;
    return 0;
}

int
sstb200_engine_form(sstb200_engine* e, int* form_out)
{
    if ( !e || !form_out ) return -1;
    *form_out = !e->pf_on ? 0 : e->pf_clock ? 3 : e->pf_ordered ? 2 : 1;
    return 0;
}

int
sstb200_engine_exchange_buffers(sstb200_engine* e, void** outbox_dev, void** inbox_dev, uint64_t* words_per_rank)
{
    if ( !e ) return -1;
    if ( outbox_dev ) *outbox_dev = e->d.outbox;
    if ( inbox_dev ) *inbox_dev = e->d.inbox;
    if ( words_per_rank ) *words_per_rank = e->d.xwords;
    return 0;
}

int
sstb200_engine_control_words(sstb200_engine* e, void** local_dev, void** global_dev)
{
    if ( !e ) return -1;
    if ( local_dev ) *local_dev = (void*)&e->d.ctrl->local[0];
    if ( global_dev ) *global_dev = (void*)e->d.gathered;
    return 0;
}

int
sstb200_engine_status(sstb200_engine* e, sstb200_status* s)
{
    if ( !e || !s ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->fetch_ctrl() ) return rc;
    const Ctrl& c       = *e->h_ctrl;
    s->now              = c.T;
    s->end_time         = c.end_time;
    s->windows          = c.windows;
    s->events_delivered = c.events_delivered;
    s->clock_ticks      = c.clock_ticks;
    s->events_sent      = c.events_sent;
    s->max_bin_fill     = c.max_bin_fill;
    s->max_due          = c.max_due;
    s->finished         = c.done;
    s->error            = c.error;
    for ( int i = 0; i < 4; ++i )
        s->error_info[i] = c.error_info[i];
    s->launches = e->launches;
    return 0;
}

int
sstb200_engine_results(sstb200_engine* e, uint64_t* results, uint32_t words_per_component)
{
    if ( !e || !results ) return -1;
    const uint32_t words = (words_per_component == 0 || words_per_component > NRESULT) ? NRESULT : words_per_component;
    CUDA_OK(cudaSetDevice(e->device));
    results_kernel<<<e->d.nbins, BLOCK, 0, e->stream>>>(e->d, e->d_results, words);
    e->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(results, e->d_results, (size_t)e->d.n_local * words * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

// ---- checkpoint / restart of the device state.  Blob layout (u64 words unless noted):
//   [0] magic "SSTB200C"  [1] ABI version  [2] number of sections  [3..14] geometry: n_local, nport_local, c0, P0, W,
//   cap, state_stride, aux_stride, has_payload, n_ranks, rank, window;  then per section: byte count, bytes (padded to 8)
namespace {
constexpr uint64_t CKPT_MAGIC = 0x4330303242545353ull; // "SSTB200C" little-endian
constexpr int      CKPT_HDR   = 15;
void
ckpt_geometry(const sstb200_engine* e, uint64_t* g)
{
    const Dev& d = e->d;
    const uint64_t v[12] = { d.n_local, d.nport_local, d.c0, d.P0, d.W, d.cap, d.state_stride, d.aux_stride,
                             (uint64_t)d.has_payload, d.n_ranks, d.rank, e->h_ctrl->w };
    memcpy(g, v, sizeof v);
}
} // namespace

int
sstb200_engine_checkpoint_size(sstb200_engine* e, uint64_t* bytes)
{
    if ( !e || !bytes ) return -1;
    uint64_t n = CKPT_HDR * 8;
    for ( const auto& sec : e->sections )
        n += 8 + ((sec.bytes + 7) & ~(size_t)7);
    *bytes = n;
    return 0;
}

int
sstb200_engine_checkpoint_save(sstb200_engine* e, void* blob, uint64_t bytes)
{
    if ( !e || !blob ) return -1;
    uint64_t need = 0;
    sstb200_engine_checkpoint_size(e, &need);
    if ( bytes < need ) {
        set_error("checkpoint_save: buffer too small (" + std::to_string(bytes) + " < " + std::to_string(need) + ")");
        return -1;
    }
    if ( !e->is_setup ) {
        set_error("checkpoint_save before engine_setup");
        return -1;
    }
    if ( e->d.trace_cap ) {
        set_error("checkpoint_save: delivery tracing and checkpoints cannot be combined");
        return -1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    uint64_t* h = (uint64_t*)blob;
    h[0]        = CKPT_MAGIC;
    h[1]        = SSTB200_ABI_VERSION;
    h[2]        = e->sections.size();
    ckpt_geometry(e, h + 3);
    uint8_t* p = (uint8_t*)blob + CKPT_HDR * 8;
    for ( const auto& sec : e->sections ) {
        *(uint64_t*)p = sec.bytes;
        p += 8;
        if ( sec.bytes ) CUDA_OK(cudaMemcpyAsync(p, sec.ptr, sec.bytes, cudaMemcpyDeviceToHost, e->stream));
        p += (sec.bytes + 7) & ~(size_t)7;
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

int
sstb200_engine_checkpoint_load(sstb200_engine* e, const void* blob, uint64_t bytes)
{
    if ( !e || !blob ) return -1;
    uint64_t need = 0;
    sstb200_engine_checkpoint_size(e, &need);
    const uint64_t* h = (const uint64_t*)blob;
    if ( bytes < CKPT_HDR * 8 || h[0] != CKPT_MAGIC ) {
        set_error("checkpoint_load: not a checkpoint of this engine");
        return -1;
    }
    uint64_t g[12];
    ckpt_geometry(e, g);
    if ( h[1] != SSTB200_ABI_VERSION || h[2] != e->sections.size() || memcmp(h + 3, g, sizeof g) != 0 || bytes < need ) {
        set_error("checkpoint_load: the checkpoint was written by an engine with a different model slice or calendar "
                  "geometry (components, ports, calendar_slots, bin_capacity, window, ranks and options must match)");
        return -1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    const uint8_t* p = (const uint8_t*)blob + CKPT_HDR * 8;
    for ( const auto& sec : e->sections ) {
        if ( *(const uint64_t*)p != sec.bytes ) {
            set_error("checkpoint_load: section size mismatch");
            return -1;
        }
        p += 8;
        if ( sec.bytes ) CUDA_OK(cudaMemcpyAsync(sec.ptr, p, sec.bytes, cudaMemcpyHostToDevice, e->stream));
        p += (sec.bytes + 7) & ~(size_t)7;
    }
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->is_setup = true; // the restored state replaces construct + setup()
    return 0;
}

// ---- periodic statistic output -----------------------------------------------------------------------------------
int
sstb200_engine_report_begin(sstb200_engine* e, uint64_t report_time)
{
    if ( !e || !e->is_setup ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    const size_t bytes = (size_t)e->d.n_local * e->d.state_stride;
    if ( !e->d.snapshot ) {
        if ( int rc = e->alloc(&e->d.snapshot, bytes, false) ) return rc;
        e->graph_failed = true; // the captured graph holds the old kernel arguments: plain launches from now on
    }
    // components that do nothing in the window up to the report time keep the state they have now
    if ( bytes ) CUDA_OK(cudaMemcpyAsync(e->d.snapshot, e->d.state, bytes, cudaMemcpyDeviceToDevice, e->stream));
    if ( e->stats_in_aux ) { // ... and the statistics they keep in their aux blocks
        const size_t abytes = (size_t)e->d.n_local * e->d.aux_stride;
        if ( !e->d.snapshot_aux )
            if ( int rc = e->alloc(&e->d.snapshot_aux, abytes, false) ) return rc;
        if ( abytes ) CUDA_OK(cudaMemcpyAsync(e->d.snapshot_aux, e->d.aux, abytes, cudaMemcpyDeviceToDevice, e->stream));
    }
    CUDA_OK(cudaMemcpyAsync(&e->d.ctrl->report_at, &report_time, 8, cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream)); // report_time is a stack variable
    e->report_pending = true;
    return 0;
}

int
sstb200_engine_report_end(sstb200_engine* e, uint64_t* results, uint32_t words_per_component)
{
    if ( !e || !results || !e->d.snapshot ) return -1;
    const uint32_t words = (words_per_component == 0 || words_per_component > NRESULT) ? NRESULT : words_per_component;
    CUDA_OK(cudaSetDevice(e->device));
    const uint64_t never = TIME_NEVER;
    e->report_pending    = false;
    CUDA_OK(cudaMemcpyAsync(&e->d.ctrl->report_at, &never, 8, cudaMemcpyHostToDevice, e->stream));
    Dev snap   = e->d;
    snap.state = e->d.snapshot; // result records of the snapshot copy
    if ( e->d.snapshot_aux ) snap.aux = e->d.snapshot_aux;
    results_kernel<<<e->d.nbins, BLOCK, 0, e->stream>>>(snap, e->d_results, words);
    e->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(results, e->d_results, (size_t)e->d.n_local * words * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

int
sstb200_engine_run_to_report(sstb200_engine* e, uint64_t report_time, uint64_t* results, uint32_t words_per_component,
    int* reported, int* finished)
{
    if ( !e || !results ) return -1;
    if ( reported ) *reported = 0;
    if ( finished ) *finished = 0;
    if ( e->n_ranks > 1 ) {
        set_error("engine_run_to_report is single-rank; a multi-rank driver brackets the window with report_begin/end");
        return -1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->fetch_ctrl() ) return rc;
    const Ctrl& c = *e->h_ctrl;
    if ( c.done ) {
        if ( finished ) *finished = 1;
        return 0;
    }
    if ( report_time < c.T ) {
        set_error("engine_run_to_report: the report time lies before the current window");
        return -1;
    }
    if ( c.stop_at && report_time >= c.stop_at ) return 0; // StopAction (priority 30) ends the run before that output
    const uint64_t before = report_time / c.w - c.k;
    if ( before ) {
        int fin = 0;
        if ( int rc = sstb200_engine_run(e, before, 64, &fin) ) return rc;
        if ( fin ) {
            if ( finished ) *finished = 1;
            return 0;
        }
    }
    if ( int rc = sstb200_engine_report_begin(e, report_time) ) return rc;
    if ( int rc = e->launch_dispatch(true) ) return rc;
    if ( int rc = e->launch_advance() ) return rc;
    if ( int rc = sstb200_engine_report_end(e, results, words_per_component) ) return rc;
    if ( int rc = e->fetch_ctrl() ) return rc;
    if ( e->h_ctrl->error ) {
        set_error("engine error " + std::to_string(e->h_ctrl->error) + " in the report window");
        return -4;
    }
    if ( reported ) *reported = 1;
    if ( finished ) *finished = e->h_ctrl->done ? 1 : 0;
    return 0;
}

int
sstb200_engine_handler_counts(sstb200_engine* e, uint64_t* event_recv, uint64_t* clock_calls)
{
    if ( !e ) return -1;
    if ( !e->d.prof_recv ) {
        set_error("engine_handler_counts: the engine was created without options.handler_counts");
        return -1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    if ( event_recv && e->d.nport_local )
        CUDA_OK(cudaMemcpyAsync(event_recv, e->d.prof_recv, (size_t)e->d.nport_local * 8, cudaMemcpyDeviceToHost, e->stream));
    if ( clock_calls && e->d.n_local )
        CUDA_OK(cudaMemcpyAsync(clock_calls, e->d.prof_clock, (size_t)e->d.n_local * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return 0;
}

int
sstb200_engine_handler_times(sstb200_engine* e, double* event_recv_ns, double* clock_ns)
{
    if ( !e ) return -1;
    if ( !e->d.prof_recv_cyc ) {
        set_error("engine_handler_times: the engine was created without options.handler_counts & 2");
        return -1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    int khz = 0;
    CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device));
    const double ns_per_cycle = khz > 0 ? 1.0e6 / (double)khz : 0.5;
    std::vector<unsigned long long> a(e->d.nport_local), b(e->d.n_local);
    if ( !a.empty() )
        CUDA_OK(cudaMemcpyAsync(a.data(), e->d.prof_recv_cyc, a.size() * 8, cudaMemcpyDeviceToHost, e->stream));
    if ( !b.empty() )
        CUDA_OK(cudaMemcpyAsync(b.data(), e->d.prof_clock_cyc, b.size() * 8, cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    if ( event_recv_ns )
        for ( size_t i = 0; i < a.size(); ++i )
            event_recv_ns[i] = (double)a[i] * ns_per_cycle;
    if ( clock_ns )
        for ( size_t i = 0; i < b.size(); ++i )
            clock_ns[i] = (double)b[i] * ns_per_cycle;
    return 0;
}

int
sstb200_engine_sync_profile(sstb200_engine* e, uint64_t* out4)
{
    if ( !e || !out4 ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->fetch_ctrl() ) return rc;
    const uint64_t rec_bytes = e->d.has_payload ? 24 : 16;
    out4[0] = e->n_ranks > 1 ? e->h_ctrl->windows : 0;
    out4[1] = e->h_ctrl->remote_events;
    out4[2] = e->h_ctrl->remote_events * rec_bytes;
    out4[3] = e->h_ctrl->sync_wait_ns;
    return 0;
}

int
sstb200_engine_trace(sstb200_engine* e, uint64_t* n_inout, uint64_t* time, uint32_t* comp, uint32_t* port,
    uint64_t* payload)
{
    if ( !e || !n_inout ) return -1;
    CUDA_OK(cudaSetDevice(e->device));
    if ( int rc = e->fetch_ctrl() ) return rc;
    uint64_t n = e->h_ctrl->trace_n;
    if ( n > e->d.trace_cap ) {
        set_error("trace overflow: " + std::to_string(n) + " deliveries, capacity " + std::to_string(e->d.trace_cap));
        return -4;
    }
    if ( !time ) {
        *n_inout = n;
        return 0;
    }
    if ( *n_inout < n ) {
        set_error("trace buffer too small");
        return -1;
    }
    std::vector<uint64_t> t(n), pl(n);
    std::vector<uint32_t> c(n), p(n), ix(n);
    if ( n ) {
        CUDA_OK(cudaMemcpy(t.data(), e->d.tr_time, n * 8, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(pl.data(), e->d.tr_payload, n * 8, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(c.data(), e->d.tr_comp, n * 4, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(p.data(), e->d.tr_port, n * 4, cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(ix.data(), e->d.tr_idx, n * 4, cudaMemcpyDeviceToHost));
    }
    std::vector<uint64_t> order(n);
    for ( uint64_t i = 0; i < n; ++i )
        order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
        if ( c[a] != c[b] ) return c[a] < c[b];
        return ix[a] < ix[b];
    });
    for ( uint64_t i = 0; i < n; ++i ) {
        time[i]    = t[order[i]];
        comp[i]    = c[order[i]];
        port[i]    = p[order[i]];
        payload[i] = pl[order[i]];
    }
    *n_inout = n;
    return 0;
}

// ------------------------------------------------------------------------------------------------ RNG entry points
int
sstb200_rng_streams(int device, void* stream, int kind, uint64_t s1, uint64_t s2, uint64_t stride1, uint64_t stride2,
    uint64_t n_streams, uint64_t draws, void* out_dev, void* checksum_dev)
{
    int ndev = 0;
    if ( cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 ) {
        set_error("rng_streams: no CUDA device available (no CPU fallback)");
        return -3;
    }
    if ( kind < 0 || kind > 2 || !checksum_dev || n_streams == 0 ) {
        set_error("rng_streams: bad arguments");
        return -1;
    }
    CUDA_OK(cudaSetDevice(device));
    if ( kind == 0 ) {
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        // groups of 2^16 generators: seeded a lane per generator into a scratch array (164 MB), then streamed
        const uint64_t group = 1ull << 16;
        static uint32_t* scratch[64]; // per device, kept (an allocation per call costs more than the seeding saves)
        static size_t    scratch_gens[64];
        const size_t     want = (size_t)std::min(group, n_streams);
        if ( scratch_gens[device & 63] < want ) {
            if ( scratch[device & 63] ) {
                CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
                cudaFree(scratch[device & 63]);
                scratch[device & 63] = nullptr;
                scratch_gens[device & 63] = 0;
            }
            CUDA_OK(cudaMalloc((void**)&scratch[device & 63], group * MT_N * 4));
            scratch_gens[device & 63] = group;
        }
        uint32_t* const seeded = scratch[device & 63];
        for ( uint64_t j0 = 0; j0 < n_streams; j0 += group ) {
            const uint64_t n = std::min(group, n_streams - j0);
            rng_mt_seed_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(s1, stride1, j0, n, seeded);
            const uint64_t blocks = std::min<uint64_t>((n + 7) / 8, (uint64_t)n_sm * 8);
            rng_mt_stream_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(s1, stride1, j0, j0 + n, draws,
                (uint32_t*)out_dev, (unsigned long long*)checksum_dev, seeded);
        }

    }
    else {
        const uint64_t blocks = (n_streams + 127) / 128;
        rng_small_stream_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(kind, s1, s2, stride1, stride2,
            n_streams, draws, (uint32_t*)out_dev, (unsigned long long*)checksum_dev);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int
sstb200_rng_ticks(int device, int kind, uint64_t s1, uint64_t s2, uint64_t first, uint64_t n, uint64_t* out_host)
{
    int ndev = 0;
    if ( cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 ) {
        set_error("rng_ticks: no CUDA device available (no CPU fallback)");
        return -3;
    }
    CUDA_OK(cudaSetDevice(device));
    uint64_t* out = nullptr;
    uint32_t* mt  = nullptr;
    CUDA_OK(cudaMalloc((void**)&out, std::max<uint64_t>(n, 1) * 5 * 8));
    CUDA_OK(cudaMalloc((void**)&mt, MT_N * 4));
    rng_ticks_kernel<<<1, 32>>>(kind, s1, s2, first, n, out, mt);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(out_host, out, n * 5 * 8, cudaMemcpyDeviceToHost));
    cudaFree(out);
    cudaFree(mt);
    return 0;
}

int
sstb200_rng_distribution(int device, int dist, const double* params, int n_params, int rng_kind, uint64_t s1,
    uint64_t s2, uint64_t n, double* out_host)
{
    int ndev = 0;
    if ( cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 ) {
        set_error("rng_distribution: no CUDA device available (no CPU fallback)");
        return -3;
    }
    if ( dist < 0 || dist > 5 || n_params < 1 || !params ) {
        set_error("rng_distribution: bad arguments");
        return -1;
    }
    CUDA_OK(cudaSetDevice(device));
    // host-side preparation mirrors the reference constructors: discrete.h:40-55 exclusive prefix sums;
    // poisson L = exp(-lambda) (poisson.h:75) computed with the host libm so the device does no transcendental.
    std::vector<double> hp(params, params + n_params);
    if ( dist == 0 ) {
        double sum = 0;
        for ( int i = 0; i < n_params; ++i ) {
            hp[i] = sum;
            sum += params[i];
        }
    }
    if ( dist == 3 ) {
        hp.resize(2);
        hp[1] = exp(-params[0]);
    }
    double*   dp  = nullptr;
    double*   out = nullptr;
    uint32_t* mt  = nullptr;
    CUDA_OK(cudaMalloc((void**)&dp, hp.size() * 8));
    CUDA_OK(cudaMalloc((void**)&out, std::max<uint64_t>(n, 1) * 8));
    CUDA_OK(cudaMalloc((void**)&mt, MT_N * 4));
    CUDA_OK(cudaMemcpy(dp, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
    rng_dist_kernel<<<1, 32>>>(dist, dp, (int)hp.size(), rng_kind, s1, s2, n, out, mt);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(out_host, out, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(dp);
    cudaFree(out);
    cudaFree(mt);
    return 0;
}

} // extern "C"
