// Clock form of a lookahead window (sm_100a) -- included by engine.cu after ordered_form.cuh.
//
// window_clock_kernel<E>: the window kernel for ranks whose components are all of ONE clock-driven element type E that
// declares CLOCK_FORM (pdes.clocknode, coreTestElement.coreTestComponent): one clock handler per component, events
// are rare (a component sends once in ~1000 ticks).  Replaces Clock::execute walking its handler list
// (clock.cc:314-397) -- one TimeVortex pop + one virtual call per handler per cycle in the reference -- by a
// whole-population pass: one thread per component, NO shared memory and NO barrier on the common path.
//
//   * a block (256 components) whose calendar bin is empty: every thread loads its 32-byte control block; a component
//     whose next tick lies beyond the window is done (its state is not even read: slow clocks of a mixed-period
//     population cost 32 bytes per window); the others load the hot part of their state (mutable core + construction
//     constants), run their ticks of this window, store the core (only its tick part, E::TICK_STORE_END, when no event
//     was delivered) and the first half of the control block.  A two-deep software pipeline in registers: the control
//     blocks of the CTA's block after next and the hot state of its next block are being loaded while a block runs;
//   * a block with a few events in its bin (at most CF_MAX_IN): every thread reads all of them (broadcast loads) and
//     keeps the one addressed to its component; the component runs its ticks and the event merged in time order
//     (clock priority 40 before event priority 50 at equal times, activity.h:64-73);
//   * sends go straight to the calendar (emit: one reservation + one record), statistics are collected per window and
//     folded into memory only when the window added data -- as in the sequential form; elements of this form keep
//     their statistics in the aux block (STATS_IN_AUX), so that the records are nothing but hot heads side by side;
//   * anything else -- more events than CF_MAX_IN, two events for one component, an event that is not due or lies at /
//     after a stop time -- sends the block to the fallback list BEFORE anything was written (every thread reaches the
//     verdict on its own from the same records: no vote, no barrier); the CTAs re-run those blocks in the sequential
//     form once their own blocks are done.
// Counters: ticks are kept per thread over all blocks of the (persistent) CTA, ONE atomic per CTA at the end (the
// sequential form's one-atomic-per-block costs 65 536 same-address atomics per counter and window on a 16 M
// population); deliveries, sends and clocks that stop are rare and counted where they happen.
#pragma once

namespace sstb200 {

template <class E, class = void>
struct clock_form_of : std::false_type
{};
template <class E>
struct clock_form_of<E, std::void_t<decltype(E::CLOCK_FORM)>> : std::bool_constant<E::CLOCK_FORM>
{};
template <class E, class = void>
struct tick_store_end_of : std::integral_constant<uint32_t, E::STORE_END>
{};
template <class E>
struct tick_store_end_of<E, std::void_t<decltype(E::TICK_STORE_END)>> : std::integral_constant<uint32_t, E::TICK_STORE_END>
{};

constexpr uint32_t CF_MAX_IN = 8; // events in a block's bin the form takes (their components fit one 64-bit register)

// 16-byte load that asks DRAM for the 64-byte half line only (SASS LDG.E.LTC64B): a plain load of a line that is not
// in L2 brings all 128 bytes (profiles/r01_notes.md, sector accounting), and a component's hot state is 64 bytes of a
// record whose rest is cold
__device__ __forceinline__ uint4
ldg_half_line(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <class E>
__global__ void __launch_bounds__(BLOCK, 2)
window_clock_kernel(const __grid_constant__ Dev d)
{
    extern __shared__ __align__(128) uint8_t pf_smem[];
    __shared__ uint32_t s_red[BLOCK / 32];
    __shared__ uint32_t s_take;
    // which of this CTA's blocks went to the fallback list (bit i = its i-th block).  The bins of the others are
    // cleared after the last barrier of the kernel: with no barrier inside the loop the warps of a CTA drift apart, and
    // a warp that fetched the fill of a bin AFTER another one had cleared it would miss the bin's events.
    constexpr uint32_t CF_MASK_WORDS = 128; // 4096 blocks per CTA = 1.2 M blocks per rank at 296 CTAs
    __shared__ uint32_t s_fbmask[CF_MASK_WORDS];
    Ctrl* const ctrl = d.ctrl;
    if ( ctrl->done ) return;
    const uint64_t k     = ctrl->k;
    const uint64_t T     = ctrl->T;
    const uint64_t T_end = ctrl->T_end;
    const uint32_t slot  = ctrl->slot;
    const uint32_t tid   = threadIdx.x;
    const uint32_t lane  = tid & 31;
    const uint32_t sh    = d.uniform_shift;
    using State          = typename E::State;

    uint32_t ticks = 0; // (deliveries, sends and clocks that stop are rare: counted where they happen)
    for ( uint32_t i = tid; i < CF_MASK_WORDS; i += BLOCK )
        s_fbmask[i] = 0;
    __syncthreads();
    // (a rank with more blocks per CTA than the mask has bits keeps the warps of a CTA together with a barrier instead)
    const bool lockstep = (d.nbins + gridDim.x - 1) / gridDim.x > CF_MASK_WORDS * 32;

    // Two blocks ahead of the one being run: the control block (first half + clock period) and the bin fill are loaded
    // into registers; one block ahead: the hot part of the state of the components that will tick.  Two blocks' loads
    // are in flight per thread while it runs a third -- nothing on the common path waits for memory.
    constexpr int NHOT = E::CONST_END / 16;
    struct Pre
    {
        uint4    lo;     // {next tick cycle lo, hi, send sequence, delivery sequence}
        uint64_t period;
        uint32_t n_in;
    };
    auto fetch_ctl = [&](uint32_t b) -> Pre {
        Pre p { make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0), 0, 0 };
        if ( b < d.nbins ) {
            p.n_in = d.bin_count[slot * d.nbins + b];
            if ( b * BLOCK + tid < d.n_local ) {
                p.lo     = *reinterpret_cast<const uint4*>(&d.ctl[b * BLOCK + tid]);
                p.period = d.ctl[b * BLOCK + tid].period;
            }
        }
        return p;
    };
    auto ticks_in_window = [&](const Pre& p) -> bool {
        const uint64_t cyc = ((uint64_t)p.lo.y << 32) | p.lo.x;
        return cyc != TIME_NEVER && cyc * p.period < T_end;
    };
    uint4 hot[NHOT], hot_next[NHOT];
    auto  fetch_hot = [&](uint32_t b, uint4* h) {
        const uint4* g = reinterpret_cast<const uint4*>(d.state + (uint64_t)(b * BLOCK + tid) * d.state_stride);
#pragma unroll
        for ( int i = 0; i < NHOT; ++i )
            h[i] = ldg_half_line(g + i);
    };
    Pre  cur      = fetch_ctl(blockIdx.x);
    Pre  nxt      = fetch_ctl(blockIdx.x + gridDim.x);
    bool hot_have = ticks_in_window(cur);
#pragma unroll
    for ( int i = 0; i < NHOT; ++i )
        hot[i] = hot_next[i] = make_uint4(0, 0, 0, 0);
    if ( hot_have ) fetch_hot(blockIdx.x, hot);

    for ( uint32_t bin = blockIdx.x, it = 0; bin < d.nbins; bin += gridDim.x, ++it ) {
        const uint32_t lc     = bin * BLOCK + tid;
        const bool     valid  = lc < d.n_local;
        const uint4    c_lo   = cur.lo;
        const uint64_t period = cur.period;
        const uint32_t n_in   = cur.n_in;
        const Pre      nn     = fetch_ctl(bin + 2 * gridDim.x);
        const bool     hot_next_have = ticks_in_window(nxt);
        if ( hot_next_have ) fetch_hot(bin + gridDim.x, hot_next);
        // ---- the (rare) events of the block: every thread looks at all of them and keeps its own
        bool       have_ev = false;
        ulonglong2 evr     = make_ulonglong2(0, 0);
        uint64_t   ev_pl   = 0;
        // A block this CTA has handed to the fallback list may already have been run by another CTA (bin cleared, state
        // advanced) when a warp that lags behind fetches its fill: the mask, set before the hand-over, tells it.
        const bool handed_over = !lockstep && ((*(volatile uint32_t*)&s_fbmask[it >> 5] >> (it & 31u)) & 1u);
        if ( n_in || handed_over ) { // (n_in: uniform over the CTA up to such laggards)
            bool           bad       = handed_over || n_in > CF_MAX_IN;
            const uint64_t base      = (uint64_t)(slot * d.nbins + bin) * d.cap;
            const uint32_t blk_port0 = d.P0 + ((bin * BLOCK) << sh);
            // Every thread sees the same records and comes to the same verdict -- also on "two events for one
            // component" (the sequential form orders those): the components seen so far ride along in one register,
            // a byte each.  No vote and no barrier: the warps of a block with events stay as independent as the others.
            uint64_t seen = 0;
            if ( !bad )
                for ( uint32_t i = 0; i < n_in; ++i ) {
                    const ulonglong2 r = d.ev[base + i];
                    const uint32_t   c = ((uint32_t)(r.y >> 32) - blk_port0) >> sh;
                    if ( r.x < T || r.x >= T_end || c >= (uint32_t)BLOCK ) bad = true; // fault / not due / stop time
                    for ( uint32_t j = 0; j < i; ++j )
                        if ( (uint32_t)(seen >> (8 * j) & 0xFFu) == c ) bad = true;
                    seen |= (uint64_t)(c & 0xFFu) << (8 * i);
                    if ( c == tid ) {
                        have_ev = true;
                        evr     = r;
                        if ( E::HAS_PAYLOAD ) ev_pl = d.ev_payload[base + i];
                    }
                }
            if ( lockstep ) __syncthreads();
            if ( bad ) {
                if ( tid == 0 ) {
                    if ( !lockstep ) {
                        s_fbmask[it >> 5] |= 1u << (it & 31u);
                        __threadfence_block();
                    }
                    d.fb_list[atomicAdd(&ctrl->fb_n, 1u)] = bin + 1; // 0 = entry not written yet
                }
                cur      = nxt;
                nxt      = nn;
                hot_have = hot_next_have;
#pragma unroll
                for ( int i = 0; i < NHOT; ++i )
                    hot[i] = hot_next[i];
                continue;
            }
        }
        const uint64_t next_cyc  = ((uint64_t)c_lo.y << 32) | c_lo.x;
        uint64_t       next_tick = next_cyc == TIME_NEVER ? TIME_NEVER : next_cyc * period;
        if ( valid && (have_ev || next_tick < T_end) ) {
            const uint32_t p0 = d.P0 + (lc << sh);
            CtxT<false>    ctx { d, T, k, T, d.c0 + lc, p0, 1u << sh, d.aux + (uint64_t)lc * d.aux_stride, c_lo.z, 0,
                              d.stats_on, d.link + (p0 - d.P0), Stage {}, nullptr, 0, 0 };
            State          s;
            uint8_t* const g = d.state + (uint64_t)lc * d.state_stride;
            uint8_t* const g_st = stats_base<E>(d.state, d.state_stride, d.aux, d.aux_stride, lc);
            const bool stats_direct = E::STATS_EVERY_EVENT && d.stats_on && have_ev;
            if ( !hot_have ) fetch_hot(bin, hot); // woken by an event only
            memcpy(&s, hot, E::CONST_END);
            if ( stats_direct )
                load_state(s, g_st, E::CONST_END, E::STATS_END);
            else
                E::initStats(s);
            const bool had_clock = next_tick != TIME_NEVER;
            uint64_t   cycle     = next_cyc;
            uint64_t   ev_t      = have_ev ? evr.x : TIME_NEVER;
            for ( ;; ) {
                if ( next_tick < T_end && next_tick <= ev_t ) {
                    ctx.now         = next_tick;
                    const bool stop = E::clockTick(ctx, s, cycle);
                    ticks++;
                    cycle++;
                    if ( stop )
                        next_tick = TIME_NEVER; // handler unregistered itself (clock.cc:324-367)
                    else
                        next_tick += period;
                    continue;
                }
                if ( ev_t == TIME_NEVER ) break;
                ctx.now = ev_t;
                E::handleEvent(ctx, s, (int)((uint32_t)(evr.y >> 32) - p0), ev_pl);
                atomicAdd(&ctrl->events_delivered, 1ull);
                atomicAdd((unsigned long long*)&ctrl->local[CW_PENDING], (unsigned long long)(-1ll));
                ev_t = TIME_NEVER;
            }
            if ( (next_tick != TIME_NEVER) != had_clock )
                atomicAdd((unsigned long long*)&ctrl->local[CW_CLOCKS], (unsigned long long)(had_clock ? -1ll : 1ll));
            const uint64_t next_store = next_tick == TIME_NEVER ? TIME_NEVER : cycle;
            store_state(s, g, 0, have_ev ? E::STORE_END : tick_store_end_of<E>::value);
            if ( stats_direct ) {
                store_state(s, g_st, E::CONST_END, E::STATS_END);
            }
            else if ( d.stats_on && E::statsVersion(s) != 0 ) {
                State mem;
                load_state(mem, g_st, E::CONST_END, E::STATS_END);
                E::mergeStats(mem, s);
                store_state(mem, g_st, E::CONST_END, E::STATS_END);
            }
            *reinterpret_cast<uint4*>(&d.ctl[lc]) = make_uint4((uint32_t)next_store, (uint32_t)(next_store >> 32), ctx.seq, c_lo.w);
            if ( ctx.sent ) {
                atomicAdd(&ctrl->events_sent, (unsigned long long)ctx.sent);
                atomicAdd((unsigned long long*)&ctrl->local[CW_PENDING], (unsigned long long)ctx.sent);
            }
        }
        if ( n_in && tid == 0 ) {
            if ( lockstep ) d.bin_count[slot * d.nbins + bin] = 0; // (every warp read the fill before the barrier above)
            if ( n_in > ctrl->max_bin_fill ) atomicMax(&ctrl->max_bin_fill, (unsigned long long)n_in);
            if ( n_in > ctrl->max_due ) atomicMax(&ctrl->max_due, (unsigned long long)n_in);
        }
        cur      = nxt;
        nxt      = nn;
        hot_have = hot_next_have;
#pragma unroll
        for ( int i = 0; i < NHOT; ++i )
            hot[i] = hot_next[i];
    }

    // ---- the tick counter: one atomic per CTA
    {
        uint32_t v1 = ticks;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 )
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
        if ( lane == 0 ) s_red[tid >> 5] = v1;
        __syncthreads();
        if ( tid == 0 ) {
            v1 = 0;
#pragma unroll
            for ( int wp = 0; wp < BLOCK / 32; ++wp )
                v1 += s_red[wp];
            if ( v1 ) atomicAdd(&ctrl->clock_ticks, (unsigned long long)v1);
        }
    }
    // ---- every warp of the CTA is through its blocks: the bins this CTA consumed are cleared now
    if ( !lockstep )
        for ( uint32_t it = tid, bin = blockIdx.x + tid * gridDim.x; bin < d.nbins; it += BLOCK, bin += BLOCK * gridDim.x )
            if ( !(s_fbmask[it >> 5] >> (it & 31u) & 1u) && d.bin_count[slot * d.nbins + bin] ) d.bin_count[slot * d.nbins + bin] = 0;

    // ---- the blocks left to the sequential form (see window_parallel_kernel: same protocol)
    for ( ;; ) {
        __syncthreads();
        if ( tid == 0 ) {
            uint32_t take = PF_NONE;
            uint32_t i    = *(volatile uint32_t*)&ctrl->fb_taken;
            while ( i < *(volatile uint32_t*)&ctrl->fb_n ) {
                const uint32_t old = atomicCAS(&ctrl->fb_taken, i, i + 1);
                if ( old == i ) {
                    volatile uint32_t* ent = d.fb_list + i;
                    uint32_t           v;
                    while ( (v = *ent) == 0 ) {}
                    *ent = 0;
                    take = v - 1;
                    break;
                }
                i = old;
            }
            s_take = take;
        }
        __syncthreads();
        const uint32_t fb = s_take;
        if ( fb == PF_NONE ) break;
        run_block_sequential<E>(d, fb, pf_smem);
    }
}

} // namespace sstb200
