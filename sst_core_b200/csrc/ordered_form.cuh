// Ordered event-parallel form of a lookahead window (sm_100a) -- included by engine.cu after parallel_form.cuh.
//
// window_ordered_kernel<E>: the window kernel for ranks whose components are all of ONE element type E that declares
// ORDERED_EVENTS (PHOLD): no clock, at most one send per delivery, a handler that touches nothing but its State, the
// event and the context's now / nports / stats_on / send / connected.  Unlike the MessageMesh form (parallel_form.cuh)
// the events of a window have different delivery times, ties are NOT interchangeable, events carry an 8-byte payload
// and sends carry the sender's sequence number.  The delivery order (time, receiving port, sequence) the reference
// gets from TimeVortexPQ (impl/timevortex/timeVortexPQ.cc:62-95, activity.h:64-73) is made per component by RANKING
// every event among the events of its component (one thread per event), not by sorting; the element's own
// handleEvent then runs unchanged, per component in that order, against a context that CAPTURES the send
// (Link::send_impl, link.cc:623-658) instead of performing it; the sends are performed event-parallel afterwards.
//
// Per block of 256 components (persistent CTAs of 448 threads, blocks handed out dynamically, TMA bulk copies of the
// NEXT block's records / payloads / link rows / component records in flight under the current block's sends):
//   1. one thread per EVENT: component c and arrival index r from one shared atomic;
//   2. components ordered busiest first (counting sort on the event count: the 32 components a warp owns later have
//      nearly equal trip counts), their segments laid out in that order by one CTA scan, events scattered into them;
//   3. one thread per EVENT, in segment order (the lanes of a warp work on the same few segments): rank k = number of
//      events of the component with a smaller key, the event index goes to slot k of the segment;
//   4. one thread per COMPONENT: E::handleEvent over the segment; the captured send becomes the outgoing calendar
//      record {arrival time, receiving port | sequence}, payload and destination IN PLACE of the consumed one; the
//      state goes back into the staged copy of the component record;
//   5. all threads: staged records -> HBM (mutable core and statistics only), outgoing records into registers, then
//      the same MSD scatter as the other forms (match.any + shared hash table + ONE global atomic per destination bin
//      per CTA), record and payload written side by side.
// Anything the form does not cover -- events not due / at or after a stop time, a time fault, a component with more
// than ORD_MAX_EVENTS deliveries, a bin longer than the staging buffer, a handler that sends twice or further than
// 2^31 cycles ahead -- is decided BEFORE the block has written anything to memory; the block is appended to the
// fallback list and re-run in the sequential form by the CTAs that have run out of blocks.
#pragma once

namespace sstb200 {

template <class E, class = void>
struct ordered_of : std::false_type
{};
template <class E>
struct ordered_of<E, std::void_t<decltype(E::ORDERED_EVENTS)>> : std::bool_constant<E::ORDERED_EVENTS>
{};

#ifndef SSTB200_OF_THREADS
#define SSTB200_OF_THREADS 448
#endif
constexpr int OF_THREADS = SSTB200_OF_THREADS; // 256 component threads + event threads
constexpr int OF_ROUNDS  = 1536 / OF_THREADS;  // events tid, tid + OF_THREADS, ...: up to 1536 staged records per block
constexpr int OF_BUCKETS = 64; // components by event count, busiest first (ORD_MAX_EVENTS < OF_BUCKETS)

// shared memory of one CTA of window_ordered_kernel<E> (host and device compute the same layout)
template <class E>
__host__ __device__ constexpr size_t
of_smem_bytes(uint32_t rcap, uint32_t sh)
{
    return (size_t)BLOCK * E::RECORD_BYTES + ((size_t)BLOCK << sh) * 16 + (size_t)rcap * 16 + (size_t)rcap * 8 +
           (size_t)rcap * 8 + (size_t)rcap * 4 + (size_t)rcap * 2 * 2 + (size_t)BLOCK * 2 +
           (size_t)(3 * BLOCK + 3 * PF_HT + 2 * OF_BUCKETS + 20 + 6) * 4 + 3 * 8 + 16;
}

// What a handler sees in this form: the CtxT subset named above; send() keeps the (one) send for the engine.
struct OrdCtx
{
    uint64_t     now;
    uint32_t     comp, nports;
    int          stats_on;
    const uint4* links; // the component's link rows (shared memory)
    uint32_t     n_sent;
    uint4        out_link;
    uint64_t     out_delay, out_payload;
    __device__ __forceinline__ bool connected(int port) const { return links[port].x != SSTB200_NO_PEER; }
    __device__ __forceinline__ void send(int port, uint64_t delay, uint64_t payload)
    {
        const uint4 l = links[port];
        if ( l.x == SSTB200_NO_PEER ) return; // (takes no sequence number: CtxT::send)
        n_sent++;
        out_link    = l;
        out_delay   = delay;
        out_payload = payload;
    }
};

// exclusive scan of one u32 per thread over the 512 threads of the CTA
__device__ __forceinline__ uint32_t
of_scan(uint32_t v, uint32_t* s_warp /* [OF_THREADS / 32 + 1] */)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t       inc  = v;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if ( lane >= (uint32_t)o ) inc += t;
    }
    if ( lane == 31 ) s_warp[warp] = inc;
    __syncthreads();
    if ( warp == 0 ) {
        const uint32_t wv = lane < OF_THREADS / 32 ? s_warp[lane] : 0;
        uint32_t       wi = wv;
#pragma unroll
        for ( int o = 1; o < OF_THREADS / 32; o <<= 1 ) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if ( lane >= (uint32_t)o ) wi += t;
        }
        if ( lane < OF_THREADS / 32 ) s_warp[lane] = wi - wv;
    }
    __syncthreads();
    return s_warp[warp] + inc - v;
}

template <class E>
__global__ void __launch_bounds__(OF_THREADS, 2)
window_ordered_kernel(const __grid_constant__ Dev d)
{
    extern __shared__ __align__(128) uint8_t pf_smem[];
    static_assert(E::ORD_MAX_EVENTS < OF_BUCKETS, "event-count buckets");
    Ctrl* const ctrl = d.ctrl;
    if ( ctrl->done ) return;
    const uint64_t     k     = ctrl->k;
    const uint64_t     T     = ctrl->T;
    const uint64_t     T_end = ctrl->T_end;
    const uint32_t     w     = (uint32_t)ctrl->w;
    const uint32_t     slot  = ctrl->slot;
    const uint32_t     slot1 = slot + 1 == d.W ? 0 : slot + 1;
    const uint32_t     tid   = threadIdx.x;
    const uint32_t     lane  = tid & 31;
    const uint32_t     sh    = d.uniform_shift;
    const uint32_t     rcap  = d.pf_rcap; // even, <= OF_ROUNDS * OF_THREADS (host)
    constexpr uint32_t REC   = E::RECORD_BYTES;
    using State              = typename E::State;

    uint8_t*    s_hot    = pf_smem;                                                          // [BLOCK][REC]
    uint4*      s_link   = reinterpret_cast<uint4*>(s_hot + (size_t)BLOCK * REC);            // [BLOCK << sh]
    ulonglong2* s_rec    = reinterpret_cast<ulonglong2*>(s_link + ((size_t)BLOCK << sh));    // [rcap] in, then out records
    uint64_t*   s_pay    = reinterpret_cast<uint64_t*>(s_rec + rcap);                        // [rcap] in, then out payloads
    uint2*      s_key    = reinterpret_cast<uint2*>(s_pay + rcap);                           // [rcap] {time offset | port, sequence}
    uint32_t*   s_dcp    = reinterpret_cast<uint32_t*>(s_key + rcap);                        // [rcap] destination component
    uint16_t*   s_perm   = reinterpret_cast<uint16_t*>(s_dcp + rcap);                        // [rcap] segments, arrival order
    uint16_t*   s_sorted = s_perm + rcap;                                                    // [rcap] segments, delivery order
    uint16_t*   s_order  = s_sorted + rcap;                                                  // [BLOCK] thread -> component
    uint32_t*   s_cnt    = reinterpret_cast<uint32_t*>(s_order + BLOCK);                     // [BLOCK] events per component
    uint32_t*   s_off    = s_cnt + BLOCK;                                                    // [BLOCK] segment start
    uint32_t*   s_seq    = s_off + BLOCK;                                                    // [BLOCK] send sequence
    uint32_t*   ht_key   = s_seq + BLOCK;                                                    // [PF_HT]
    uint32_t*   ht_cnt   = ht_key + PF_HT;
    uint32_t*   ht_base  = ht_cnt + PF_HT;
    uint32_t*   s_hist   = ht_base + PF_HT;                                                  // [OF_BUCKETS]
    uint32_t*   s_hbase  = s_hist + OF_BUCKETS;                                              // [OF_BUCKETS]
    uint32_t*   s_warp   = s_hbase + OF_BUCKETS;                                             // [20]
    uint32_t*   s_nin    = s_warp + 20;                                                      // [2]
    uint32_t*   s_nbin   = s_nin + 2;                                                        // [2]
    uint32_t*   s_flag   = s_nbin + 2;                                                       // [2]
    uint64_t*   s_bar    = reinterpret_cast<uint64_t*>(s_flag + 2);                          // [3] hot, link, rec

    // thread 0: the staging copies of block b (bin fill n) -- component records | link rows, records and payloads
    auto issue_hot = [&](uint32_t b) {
        const uint32_t nc = min((uint32_t)BLOCK, d.n_local - b * BLOCK);
        mbar_expect_tx(&s_bar[0], nc * REC);
        bulk_copy_g2s(s_hot, d.state + (uint64_t)b * BLOCK * REC, nc * REC, &s_bar[0]);
    };
    auto issue_rest = [&](uint32_t b, uint32_t n) {
        const uint32_t nc = min((uint32_t)BLOCK, d.n_local - b * BLOCK);
        mbar_expect_tx(&s_bar[1], (nc << sh) * 16u);
        bulk_copy_g2s(s_link, d.link + (((uint64_t)b * BLOCK) << sh), (nc << sh) * 16u, &s_bar[1]);
        if ( n > rcap ) n = rcap; // the block will fall back; nothing past the buffer is looked at
        const uint32_t np = (n + 1u) & ~1u; // payloads are 8 bytes: an even number keeps the copy a multiple of 16
        mbar_expect_tx(&s_bar[2], n * 16u + np * 8u);
        if ( n ) {
            const uint64_t base = (uint64_t)(slot * d.nbins + b) * d.cap;
            bulk_copy_g2s(s_rec, d.ev + base, n * 16u, &s_bar[2]);
            bulk_copy_g2s(s_pay, d.ev_payload + base, np * 8u, &s_bar[2]);
        }
    };
    // the send sequence numbers (CompCtl) of block b's components, fetched one block ahead
    auto fetch_seq = [&](uint32_t b) -> uint32_t {
        if ( b >= d.nbins ) return 0;
        const uint32_t lc = b * BLOCK + tid;
        return (tid < (uint32_t)BLOCK && lc < d.n_local) ? d.ctl[lc].send_seq : 0;
    };

    uint32_t bin = blockIdx.x;
    if ( tid == 0 ) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ( tid < (uint32_t)BLOCK ) s_cnt[tid] = 0;
    __syncthreads();
    if ( tid == 0 && bin < d.nbins ) {
        const uint32_t n0 = d.bin_count[slot * d.nbins + bin];
        s_nin[0]          = n0;
        issue_hot(bin);
        issue_rest(bin, n0);
    }
    uint32_t seq_next = fetch_seq(bin);
    __syncthreads();

    uint32_t delivered = 0, sent = 0, max_fill = 0;
    for ( uint32_t it = 0; bin < d.nbins; ++it, bin = s_nbin[it & 1u] ) {
        const uint32_t ph  = it & 1u;
        uint32_t       nxt = 0xFFFFFFFFu, n_next = 0; // thread 0 only
        if ( tid == 0 ) {
            nxt       = gridDim.x + atomicAdd(&ctrl->pf_next, 1u); // (its bin fill is read after the first barrier)
            s_flag[0] = 0;
        }
        if ( tid < PF_HT ) {
            ht_key[tid] = PF_NONE;
            ht_cnt[tid] = 0;
        }
        if ( tid < (uint32_t)OF_BUCKETS ) s_hist[tid] = 0;
        const uint32_t n_in      = s_nin[ph];
        const uint32_t lc0       = bin * BLOCK;
        const uint32_t blk_port0 = d.P0 + (lc0 << sh);
        const uint32_t n_comp    = min((uint32_t)BLOCK, d.n_local - lc0);
        if ( tid < n_comp ) s_seq[tid] = seq_next;

        // the end of a block that cannot be taken: on the fallback list, next block's copies, nothing written
        auto give_up = [&]() {
            if ( tid == 0 ) {
                d.fb_list[atomicAdd(&ctrl->fb_n, 1u)] = bin + 1; // 0 = entry not written yet
                s_nbin[ph ^ 1u] = nxt;
                if ( nxt < d.nbins ) {
                    s_nin[ph ^ 1u] = n_next;
                    issue_hot(nxt);
                    issue_rest(nxt, n_next);
                }
            }
            if ( tid < (uint32_t)BLOCK ) s_cnt[tid] = 0;
            __syncthreads();
            seq_next = fetch_seq(s_nbin[ph ^ 1u]);
        };

        // ---- 1. records: component and arrival index of every event
        mbar_wait(&s_bar[2], ph);
        bool           bad   = n_in > rcap;
        const uint32_t n_rec = min(n_in, rcap);
        uint32_t       ck[OF_ROUNDS];
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j ) {
            const uint32_t e = tid + j * OF_THREADS;
            ck[j]            = PF_NONE;
            if ( e < n_rec ) {
                const ulonglong2 r = s_rec[e];
                const uint32_t   c = ((uint32_t)(r.y >> 32) - blk_port0) >> sh;
                if ( r.x < T || r.x >= T_end || c >= n_comp )
                    bad = true; // time fault / not due / at or after a stop time: the sequential form
                else {
                    ck[j] = c | (atomicAdd(&s_cnt[c], 1u) << 8);
                    // the delivery key in 8 bytes: (time in the window, port of the component) | sequence
                    s_key[e] = make_uint2(((uint32_t)(r.x - T) << sh) | (((uint32_t)(r.y >> 32) - blk_port0) & ((1u << sh) - 1u)),
                        (uint32_t)r.y);
                }
            }
        }
        mbar_wait(&s_bar[0], ph);
        __syncthreads();
        if ( tid == 0 && nxt < d.nbins ) n_next = d.bin_count[slot * d.nbins + nxt];
        // ---- 2. components busiest first
        uint32_t bucket = OF_BUCKETS - 1;
        if ( tid < (uint32_t)BLOCK ) {
            const uint32_t n = s_cnt[tid];
            if ( n > E::ORD_MAX_EVENTS ) bad = true;
            bucket = n > E::ORD_MAX_EVENTS ? 0u : E::ORD_MAX_EVENTS - n;
            atomicAdd(&s_hist[bucket], 1u);
        }
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        mbar_wait(&s_bar[1], ph); // everybody, also a block that falls back: the buffer is refilled below
        if ( any_bad ) {
            give_up();
            continue;
        }
        if ( tid < 32 ) { // exclusive scan of the buckets by one warp, two per lane
            const uint32_t v0 = s_hist[2 * tid], v1 = s_hist[2 * tid + 1];
            uint32_t       inc = v0 + v1;
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if ( lane >= (uint32_t)o ) inc += t;
            }
            s_hbase[2 * tid]     = inc - v0 - v1;
            s_hbase[2 * tid + 1] = inc - v1;
            s_hist[2 * tid]      = 0; // reused as the bucket fill counters
            s_hist[2 * tid + 1]  = 0;
        }
        __syncthreads();
        if ( tid < (uint32_t)BLOCK ) s_order[s_hbase[bucket] + atomicAdd(&s_hist[bucket], 1u)] = (uint16_t)tid;
        __syncthreads();
        // thread t (< 256) owns component s_order[t] from here on; segments in that order
        uint32_t my_c = 0, my_n = 0;
        if ( tid < (uint32_t)BLOCK ) {
            my_c = s_order[tid];
            my_n = s_cnt[my_c];
        }
        const uint32_t my_off = of_scan(my_n, s_warp);
        if ( tid < (uint32_t)BLOCK ) s_off[my_c] = my_off;
        __syncthreads();
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j )
            if ( ck[j] != PF_NONE ) s_perm[s_off[ck[j] & 0xFFu] + (ck[j] >> 8)] = (uint16_t)(tid + j * OF_THREADS);
        __syncthreads();
        // ---- 3. rank of every event in its component's delivery order (threads in segment order)
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j ) {
            const uint32_t p = tid + j * OF_THREADS;
            if ( p >= n_rec ) break;
            const uint32_t e    = s_perm[p];
            const uint2    me   = s_key[e];
            const uint32_t c    = (reinterpret_cast<const uint32_t*>(&s_rec[e])[3] - blk_port0) >> sh;
            const uint32_t base = s_off[c], nc = s_cnt[c];
            uint32_t       kk   = 0;
            for ( uint32_t i = 0; i < nc; ++i ) {
                const uint32_t e2 = s_perm[base + i];
                const uint2    o  = s_key[e2];
                // (time, receiving port) then the sequence, compared wrap-safe; identical keys by staging index
                const int32_t  dq = (int32_t)(o.y - me.y);
                kk += ((o.x < me.x) | ((o.x == me.x) & ((dq < 0) | ((dq == 0) & (e2 < e))))) ? 1u : 0u;
            }
            s_sorted[base + kk] = (uint16_t)e;
        }
        __syncthreads();
        // ---- 4. the handlers, one thread per component in delivery order; sends captured in place of the records
        if ( my_n ) {
            State s;
            load_state(s, s_hot + (size_t)my_c * REC, 0, E::STATS_END);
            const uint32_t port0 = blk_port0 + (my_c << sh);
            uint32_t       seq   = s_seq[my_c];
            OrdCtx         ctx { 0, d.c0 + lc0 + my_c, 1u << sh, d.stats_on, s_link + (my_c << sh), 0, make_uint4(0, 0, 0, 0), 0, 0 };
            bool           far = false;
            for ( uint32_t i = 0; i < my_n; ++i ) {
                const uint32_t   e  = s_sorted[my_off + i];
                const ulonglong2 r  = s_rec[e];
                uint64_t         pl = s_pay[e];
                ctx.now             = r.x;
                ctx.n_sent          = 0;
                E::handleEvent(ctx, s, (int)((uint32_t)(r.y >> 32) - port0), pl);
                uint32_t dcp = SSTB200_NO_PEER;
                if ( ctx.n_sent == 1 ) {
                    const uint4    l   = ctx.out_link;
                    const uint64_t dt6 = (r.x - T) + (((uint64_t)l.w << 32) | l.z) + ctx.out_delay;
                    if ( dt6 >= 0x7FFFFFFFull )
                        far = true; // needs 64-bit times: the sequential form
                    else {
                        s_rec[e] = make_ulonglong2(T + dt6, ((uint64_t)l.x << 32) | seq);
                        s_pay[e] = ctx.out_payload;
                        dcp      = l.y;
                        seq++;
                    }
                }
                else if ( ctx.n_sent > 1 )
                    far = true;
                s_dcp[e] = dcp;
            }
            store_state(s, s_hot + (size_t)my_c * REC, 0, E::STORE_END);
            if ( d.stats_on ) store_state(s, s_hot + (size_t)my_c * REC, E::CONST_END, E::STATS_END);
            s_seq[my_c] = seq;
            if ( far ) s_flag[0] = 1;
        }
        __syncthreads();
        if ( s_flag[0] ) {
            give_up();
            continue;
        }
        // ---- 5. nothing can stop the block any more: the sends ranked inside their destination bins
        uint32_t key[OF_ROUNDS], dtm[OF_ROUNDS], dcp[OF_ROUNDS], er[OF_ROUNDS];
        uint64_t ds[OF_ROUNDS], pay[OF_ROUNDS];
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j ) {
            const uint32_t e = tid + j * OF_THREADS;
            key[j]           = PF_NONE;
            dtm[j] = dcp[j] = 0;
            ds[j] = pay[j] = 0;
            if ( e >= n_rec ) continue;
            delivered++;
            const uint32_t dc = s_dcp[e];
            if ( dc == SSTB200_NO_PEER ) continue;
            sent++;
            const ulonglong2 r  = s_rec[e];
            const uint32_t   dt = (uint32_t)(r.x - T);
            dcp[j]              = dc;
            dtm[j]              = dt;
            ds[j]               = r.y;
            pay[j]              = s_pay[e];
            uint32_t sl         = slot1;
            if ( dt >= 2 * w ) {
                uint32_t dk = div_w(d, dt);
                if ( dk > d.W - 1 ) dk = d.W - 1; // beyond the horizon: parked in the farthest slot, re-binned there
                sl = wrap_slot(d, slot + dk);
            }
            if ( d.n_ranks > 1 && (dc < d.c0 || dc >= d.c0 + d.n_local) )
                key[j] = remote_key(d, dc, sl);
            else
                key[j] = sl * d.nbins + (dc - d.c0) / BLOCK;
        }
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j ) {
            er[j] = PF_NONE;
            if ( (tid & ~31u) + j * OF_THREADS >= n_rec ) continue; // the whole warp is past the last record
            const uint32_t ky     = key[j];
            const uint32_t peers  = __match_any_sync(0xffffffffu, ky);
            const uint32_t leader = __ffs(peers) - 1;
            uint32_t       ent = PF_NONE, base = 0;
            if ( ky != PF_NONE && lane == leader ) {
                uint32_t h = (ky * 2654435761u) >> 25;
#pragma unroll
                for ( int pr = 0; pr < 4; ++pr ) { // open addressing, 4 probes
                    uint32_t old = ht_key[h];
                    if ( old != ky ) old = atomicCAS(&ht_key[h], PF_NONE, ky);
                    if ( old == PF_NONE || old == ky ) {
                        ent = h;
                        break;
                    }
                    h = (h + 1) & (PF_HT - 1);
                }
                if ( ent != PF_NONE ) base = atomicAdd(&ht_cnt[ent], (uint32_t)__popc(peers));
            }
            ent   = __shfl_sync(0xffffffffu, ent, leader);
            base  = __shfl_sync(0xffffffffu, base, leader);
            er[j] = PF_NONE;
            if ( ky == PF_NONE ) continue;
            if ( ent != PF_NONE )
                er[j] = ent | ((base + __popc(peers & ((1u << lane) - 1))) << 8);
            else // more than PF_HT distinct destination bins in one block: reserve directly
                emit(d, k, T + dtm[j], (uint32_t)(ds[j] >> 32), dcp[j], (uint32_t)ds[j], pay[j]);
        }
        __syncthreads();
        // records, payloads and link rows are free: the next block's copies fly from here on
        if ( tid == 0 ) {
            s_nbin[ph ^ 1u] = nxt;
            if ( nxt < d.nbins ) {
                s_nin[ph ^ 1u] = n_next;
                issue_rest(nxt, n_next);
            }
        }
        if ( tid < PF_HT && ht_cnt[tid] ) { // ONE global atomic per destination bin (its latency runs under the copy below)
            const uint32_t ky = ht_key[tid];
            if ( ky & 0x80000000u ) {
                ht_base[tid] = remote_reserve(d, ky, ht_cnt[tid]);
                atomicAdd(&ctrl->remote_events, (unsigned long long)ht_cnt[tid]); // sync profile tools
            }
            else
                ht_base[tid] = atomicAdd(&d.bin_count[ky], ht_cnt[tid]);
        }
        // state and sequence numbers to memory: the staged records' mutable core and statistics, 16 bytes a thread
        {
            constexpr uint32_t NA = E::STORE_END / 16, NB = (E::STATS_END - E::CONST_END) / 16;
            uint4* const       g  = reinterpret_cast<uint4*>(d.state + (uint64_t)lc0 * REC);
            const uint4* const sm = reinterpret_cast<const uint4*>(s_hot);
            for ( uint32_t i = tid; i < n_comp * NA; i += OF_THREADS ) {
                const uint32_t c = i / NA, q = (REC / 16) * c + (i - c * NA);
                g[q]             = sm[q];
            }
            if constexpr ( NB > 0 ) {
                if ( d.stats_on )
                    for ( uint32_t i = tid; i < n_comp * NB; i += OF_THREADS ) {
                        const uint32_t c = i / NB, q = (REC / 16) * c + E::CONST_END / 16 + (i - c * NB);
                        g[q]             = sm[q];
                    }
            }
            if ( tid < (uint32_t)BLOCK ) {
                if ( tid < n_comp && s_cnt[tid] ) d.ctl[lc0 + tid].send_seq = s_seq[tid];
                s_cnt[tid] = 0;
            }
        }
        __syncthreads();
        if ( tid == 0 && nxt < d.nbins ) issue_hot(nxt); // the component records are free too
        seq_next = fetch_seq(s_nbin[ph ^ 1u]);
#pragma unroll
        for ( int j = 0; j < OF_ROUNDS; ++j ) {
            if ( er[j] == PF_NONE ) continue;
            const uint32_t ky  = key[j];
            const uint32_t pos = ht_base[er[j] & 0xFFu] + (er[j] >> 8);
            const uint64_t tm  = T + dtm[j];
            if ( ky & 0x80000000u ) {
                remote_write(d, ky, pos, tm, ds[j], pay[j], dcp[j]);
            }
            else if ( pos >= d.cap ) {
                raise_error(d, ERR_BIN_OVERFLOW, ky / d.nbins, ky % d.nbins, pos);
            }
            else {
                const uint64_t off = (uint64_t)ky * d.cap + pos;
                d.ev[off]          = make_ulonglong2(tm, ds[j]);
                d.ev_payload[off]  = pay[j];
            }
        }
        if ( tid == 0 ) {
            d.bin_count[slot * d.nbins + bin] = 0; // every thread read its records before the barriers above
            max_fill                          = max(max_fill, n_in);
        }
    }

    // ---- counters: one atomic per warp
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) {
        delivered += __shfl_down_sync(0xffffffffu, delivered, o);
        sent += __shfl_down_sync(0xffffffffu, sent, o);
    }
    if ( lane == 0 && delivered ) {
        atomicAdd(&ctrl->events_delivered, (unsigned long long)delivered);
        if ( sent ) atomicAdd(&ctrl->events_sent, (unsigned long long)sent);
        if ( sent != delivered )
            atomicAdd((unsigned long long*)&ctrl->local[CW_PENDING], (unsigned long long)((long long)sent - (long long)delivered));
    }
    if ( tid == 0 && max_fill ) {
        if ( max_fill > ctrl->max_bin_fill ) atomicMax(&ctrl->max_bin_fill, (unsigned long long)max_fill);
        if ( max_fill > ctrl->max_due ) atomicMax(&ctrl->max_due, (unsigned long long)max_fill);
    }

    // ---- the blocks left to the sequential form, by the CTAs that have no block of their own left (see
    // window_parallel_kernel: same protocol on Ctrl::fb_n / fb_taken / Dev::fb_list).  The sequential form is a
    // 256-thread routine: the second half of the CTA only keeps the barriers company.
    if ( tid >= (uint32_t)BLOCK ) return; // (whole warps: the barriers below count the warps that are left)
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
            s_nbin[0] = take;
        }
        __syncthreads();
        const uint32_t fb = s_nbin[0];
        if ( fb == PF_NONE ) break;
        __syncthreads(); // s_nbin lies inside the area dispatch_block lays out afresh
        run_block_sequential<E>(d, fb, pf_smem);
    }
}

} // namespace sstb200
