// Event-parallel form of a lookahead window (sm_100a) -- included by engine.cu after Dev / Ctrl / emit / with_element.
//
// window_parallel_kernel<E>: the window kernel for ranks whose components are all of ONE element type E that declares
// PARALLEL_EVENTS and TIES_INTERCHANGEABLE (MessageMesh): the handler of the k-th event a component receives in a
// window is a pure function of the component's record at the start of the window and k, and events of equal delivery
// time may run in any order.  Replaces, for such models, TimeVortexPQ pop (impl/timevortex/timeVortexPQ.cc:62-95) +
// Event::execute (event.cc:27-30) + Link::send_impl (link.cc:623-658) per event by, per block of 256 components:
//
//   * three 1-D bulk copies (cp.async.bulk, completion on an mbarrier) issued by one thread: the block's calendar bin
//     (16-byte records), its link rows, its component records.  The kernel is persistent: a CTA walks blocks
//     b, b+G, b+2G ... and every buffer is refilled for the NEXT block as soon as the current block has made its last
//     read of it, so the copies of block b+G fly under the handlers of block b and nothing in the block's critical
//     path waits for DRAM;
//   * one thread per EVENT: its component and its rank k among the component's events of the window come from ONE
//     shared-memory atomic; the element turns (record, k) into the port the event leaves on (MessageMesh: a look-up in
//     the route digest of its MT19937 generation, components.cuh); the link row gives the receiving port, the
//     destination component and the latency;
//   * one thread per COMPONENT: the state update for all n events at once (E::parFinish), written back as one
//     16-byte store; components whose generator needs work (next generation, digest tail) are appended to the rank's
//     maintenance list, which maintenance_kernel processes after the window, a warp per component;
//   * the send = MSD radix scatter on (window slot | destination block): lanes of a warp with the same destination
//     bin are found with match.any, one shared-memory atomic per distinct bin per warp gives the records their rank,
//     ONE global atomic per distinct bin per CTA reserves the calendar slots, the records are written side by side.
//
// Anything the form does not cover -- a block with events of several delivery times, events that are not due or lie
// at/after a stop time, a component with more than PARALLEL_MAX_EVENTS events, a bin longer than the staging buffer,
// a component the element refuses (parHeader) -- is decided BEFORE the block has written anything; the block is then
// appended to the fallback list and run by dispatch_kernel (sequential form) right after this kernel.
#pragma once

namespace sstb200 {

constexpr int      PF_THREADS = 256;
constexpr int      PF_ROUNDS  = 5; // records tid, tid+256, ... : up to 1280 staged records per block
constexpr uint32_t PF_HT      = 128;
constexpr uint32_t PF_NONE    = 0xFFFFFFFFu;

// the end of a window (engine.cu): maintenance_kernel's last CTA runs it in the native run loop
__device__ void advance_fold(const Dev& d);
__device__ __forceinline__ void peer_post(const Dev& d);
__device__ __forceinline__ void peer_wait_and_fold(const Dev& d, int spin);

__device__ __forceinline__ uint32_t
smem_addr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void
mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void
mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP); dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void
bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void
mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while ( !done );
}

// window slot of a send that arrives `dt` cycles after the start of the current window, dt >= 2w (rare path)
__device__ __noinline__ uint32_t
pf_far_slot(uint32_t dt, uint32_t w, uint32_t slot, uint32_t W)
{
    uint32_t dk = dt / w;
    if ( dk > W - 1 ) dk = W - 1; // beyond the calendar horizon: parked in the farthest slot, re-binned there
    return (slot + dk) % W;
}

// the sequential form of one block, kept out of line: the register allocation of the event-parallel loop stays its own
template <class E>
__device__ __noinline__ void
run_block_sequential(const Dev& d, uint32_t bin, uint8_t* smem)
{
    dispatch_block<E::HAS_PAYLOAD, false, E>(d, bin, smem);
}

template <class E>
__global__ void __launch_bounds__(PF_THREADS, 3)
window_parallel_kernel(const __grid_constant__ Dev d) // (grid constant: run_block_sequential takes it by reference)
{
    extern __shared__ __align__(128) uint8_t pf_smem[];
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
    constexpr uint32_t REC   = E::RECORD_BYTES;

    uint8_t*    s_hot   = pf_smem;                                            // [BLOCK][REC] component records
    uint4*      s_link  = reinterpret_cast<uint4*>(s_hot + (size_t)BLOCK * REC); // [BLOCK << sh] link rows
    ulonglong2* s_rec   = reinterpret_cast<ulonglong2*>(s_link + ((size_t)BLOCK << sh)); // [pf_rcap] calendar records
    uint32_t*   s_cnt   = reinterpret_cast<uint32_t*>(s_rec + d.pf_rcap);     // [BLOCK] events per component
    uint32_t*   s_hdr   = s_cnt + BLOCK;                                      // [BLOCK] element header words
    uint32_t*   ht_key  = s_hdr + BLOCK;                                      // [PF_HT] destination bins of the block
    uint32_t*   ht_cnt  = ht_key + PF_HT;
    uint32_t*   ht_base = ht_cnt + PF_HT;
    uint32_t*   s_nin   = ht_base + PF_HT;                                    // [2] fill of the bin in s_rec, by parity
    uint32_t*   s_nbin  = s_nin + 2;                                          // [2] the block that bin belongs to
    uint64_t*   s_bar   = reinterpret_cast<uint64_t*>(s_nbin + 2);            // [3] hot, link, rec

    auto issue_hot = [&](uint32_t b) {
        const uint32_t nc    = min((uint32_t)BLOCK, d.n_local - b * BLOCK);
        const uint32_t bytes = nc * REC;
        mbar_expect_tx(&s_bar[0], bytes);
        bulk_copy_g2s(s_hot, d.state + (uint64_t)b * BLOCK * REC, bytes, &s_bar[0]);
    };
    auto issue_link = [&](uint32_t b) {
        const uint32_t nc    = min((uint32_t)BLOCK, d.n_local - b * BLOCK);
        const uint32_t bytes = (nc << sh) * 16u;
        mbar_expect_tx(&s_bar[1], bytes);
        bulk_copy_g2s(s_link, d.link + (((uint64_t)b * BLOCK) << sh), bytes, &s_bar[1]);
    };
    auto issue_rec = [&](uint32_t b, uint32_t n) {
        if ( n > d.pf_rcap ) n = d.pf_rcap; // the block will fall back; nothing past the buffer is looked at
        mbar_expect_tx(&s_bar[2], n * 16u);
        if ( n ) bulk_copy_g2s(s_rec, d.ev + (uint64_t)(slot * d.nbins + b) * d.cap, n * 16u, &s_bar[2]);
    };

    uint32_t bin = blockIdx.x;
    if ( tid == 0 ) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init(&s_bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    s_cnt[tid] = 0;
    __syncthreads();
    if ( tid == 0 && bin < d.nbins ) {
        const uint32_t n0 = d.bin_count[slot * d.nbins + bin];
        s_nin[0]          = n0;
        issue_rec(bin, n0);
        issue_hot(bin);
        issue_link(bin);
    }
    __syncthreads();

    uint32_t delivered = 0, unsent = 0, max_fill = 0;
    // Blocks are handed out dynamically (Ctrl::pf_next, reset when a window opens): the CTAs of a rank
    // that holds 8192 blocks make ~18 rounds each, and with a fixed stride the last round would leave most SMs idle
    // while the slowest CTAs finish.
    for ( uint32_t it = 0; bin < d.nbins; ++it, bin = s_nbin[it & 1u] ) {
        const uint32_t ph  = it & 1u;
        uint32_t       nxt = 0xFFFFFFFFu, n_next = 0; // thread 0 only
        if ( tid == 0 ) {
            nxt = gridDim.x + atomicAdd(&ctrl->pf_next, 1u); // the first gridDim.x blocks are the CTAs' own
            if ( nxt < d.nbins ) n_next = d.bin_count[slot * d.nbins + nxt]; // used after the first barrier
        }
        if ( tid < PF_HT ) {
            ht_key[tid] = PF_NONE;
            ht_cnt[tid] = 0;
        }
        const uint32_t n_in      = s_nin[ph];
        const uint32_t lc0       = bin * BLOCK;
        const uint32_t blk_port0 = d.P0 + (lc0 << sh);

        // ---- records: component and rank of every event (one shared atomic each)
        mbar_wait(&s_bar[2], ph);
        bool     bad = n_in > d.pf_rcap;
        uint64_t t0  = T;
        if ( n_in ) t0 = s_rec[0].x;
        if ( t0 < T || t0 >= T_end ) bad = true; // time fault / not due / at or after a stop time: the sequential form
        const uint32_t n_rec = min(n_in, d.pf_rcap);
        uint32_t       ck[PF_ROUNDS];
#pragma unroll
        for ( int j = 0; j < PF_ROUNDS; ++j ) {
            const uint32_t e = tid + j * PF_THREADS;
            ck[j]            = PF_NONE;
            if ( e < n_rec ) {
                const ulonglong2 r = s_rec[e];
                const uint32_t   c = ((uint32_t)(r.y >> 32) - blk_port0) >> sh;
                if ( r.x != t0 || c >= (uint32_t)BLOCK )
                    bad = true;
                else
                    ck[j] = c | (atomicAdd(&s_cnt[c], 1u) << 8);
            }
        }
        // ---- component records: header word of every component
        mbar_wait(&s_bar[0], ph);
        uint32_t   my_hdr = 0;
        const bool my_ok  = (lc0 + tid < d.n_local) && E::parHeader(s_hot + (size_t)tid * REC, my_hdr);
        s_hdr[tid]        = my_hdr;
        __syncthreads();
        // the record buffer is free: the next block's bin is copied while this block runs
        if ( tid == 0 ) {
            s_nbin[ph ^ 1u] = nxt;
            if ( nxt < d.nbins ) {
                s_nin[ph ^ 1u] = n_next;
                issue_rec(nxt, n_next);
            }
        }
        const uint32_t n = s_cnt[tid];
        s_cnt[tid]       = 0;
        if ( n && (!my_ok || n > d.par_max) ) bad = true;
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        mbar_wait(&s_bar[1], ph); // link rows (everybody, also a block that falls back: the buffer is refilled below)

        if ( any_bad ) {
            // nothing has been written: the sequential form runs this block after this kernel
            if ( tid == 0 ) d.fb_list[atomicAdd(&ctrl->fb_n, 1u)] = bin + 1; // 0 = entry not written yet
            __syncthreads();
            if ( tid == 0 && nxt < d.nbins ) {
                issue_hot(nxt);
                issue_link(nxt);
            }
            continue;
        }

        // ---- component thread: all n events at once
        if ( n ) {
            uint4          core = *reinterpret_cast<const uint4*>(s_hot + (size_t)tid * REC);
            const uint32_t op   = E::parFinish(core, n, d.stats_on != 0);
            *reinterpret_cast<uint4*>(d.state + (uint64_t)(lc0 + tid) * REC) = core;
            if ( op ) d.maint_list[atomicAdd(&ctrl->maint_n, 1u)] = (lc0 + tid) | (op << 30);
        }
        // ---- event threads: the handler (a look-up) and the send
        uint32_t key[PF_ROUNDS], dpt[PF_ROUNDS], dtm[PF_ROUNDS], dcp[PF_ROUNDS], er[PF_ROUNDS];
        const uint32_t t_off = (uint32_t)(t0 - T);
#pragma unroll
        for ( int j = 0; j < PF_ROUNDS; ++j ) {
            key[j] = PF_NONE;
            dpt[j] = dtm[j] = dcp[j] = 0;
            if ( ck[j] == PF_NONE ) continue;
            delivered++;
            const uint32_t c    = ck[j] & 0xFFu;
            const uint32_t port = E::parRoute(s_hdr[c], s_hot + (size_t)c * REC, ck[j] >> 8);
            const uint4    l    = s_link[(c << sh) + port];
            if ( l.x == SSTB200_NO_PEER ) {
                unsent++;
                continue;
            }
            const uint32_t dt = t_off + l.z; // link latencies are below 2^31 (checked by the host)
            dpt[j]            = l.x;
            dcp[j]            = l.y;
            dtm[j]            = dt;
            const uint32_t sl = dt < 2 * w ? slot1 : pf_far_slot(dt, w, slot, d.W);
            if ( d.n_ranks > 1 && (l.y < d.c0 || l.y >= d.c0 + d.n_local) )
                key[j] = remote_key(d, l.y, sl);
            else
                key[j] = sl * d.nbins + (l.y - d.c0) / BLOCK;
        }
        // ---- rank of every record inside its destination bin
#pragma unroll
        for ( int j = 0; j < PF_ROUNDS; ++j ) {
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
            else // more than PF_HT distinct destination bins in one block (irregular graphs): reserve directly
                emit(d, k, T + dtm[j], dpt[j], dcp[j], 0, 0);
        }
        __syncthreads();
        // link rows and component records are free: the next block's are copied while this block writes its sends
        if ( tid == 0 && nxt < d.nbins ) {
            issue_hot(nxt);
            issue_link(nxt);
        }
        if ( tid < PF_HT && ht_cnt[tid] ) {
            const uint32_t ky = ht_key[tid];
            if ( ky & 0x80000000u ) {
                ht_base[tid] = remote_reserve(d, ky, ht_cnt[tid]);
                atomicAdd(&ctrl->remote_events, (unsigned long long)ht_cnt[tid]); // sync profile tools
            }
            else
                ht_base[tid] = atomicAdd(&d.bin_count[ky], ht_cnt[tid]);
        }
        __syncthreads();
#pragma unroll
        for ( int j = 0; j < PF_ROUNDS; ++j ) {
            if ( er[j] == PF_NONE ) continue;
            const uint32_t ky  = key[j];
            const uint32_t pos = ht_base[er[j] & 0xFFu] + (er[j] >> 8);
            const uint64_t tm  = T + dtm[j];
            const uint64_t ds  = (uint64_t)dpt[j] << 32; // nobody reads a sequence number in these models (ctl_free)
            if ( ky & 0x80000000u ) {
                remote_write(d, ky, pos, tm, ds, 0, dcp[j]);
            }
            else if ( pos >= d.cap ) {
                raise_error(d, ERR_BIN_OVERFLOW, ky / d.nbins, ky % d.nbins, pos);
            }
            else {
                d.ev[(uint64_t)ky * d.cap + pos] = make_ulonglong2(tm, ds);
            }
        }
        if ( tid == 0 ) {
            d.bin_count[slot * d.nbins + bin] = 0; // every thread read its records before the first barrier
            max_fill                          = max(max_fill, n_in);
        }
    }

    // ---- counters: one atomic per warp
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) {
        delivered += __shfl_down_sync(0xffffffffu, delivered, o);
        unsent += __shfl_down_sync(0xffffffffu, unsent, o);
    }
    if ( lane == 0 ) {
        if ( delivered ) {
            atomicAdd(&ctrl->events_delivered, (unsigned long long)delivered);
            atomicAdd(&ctrl->events_sent, (unsigned long long)(delivered - unsent));
        }
        if ( unsent ) atomicAdd((unsigned long long*)&ctrl->local[CW_PENDING], (unsigned long long)(-(long long)unsent));
    }
    if ( tid == 0 && max_fill ) {
        if ( max_fill > ctrl->max_bin_fill ) atomicMax(&ctrl->max_bin_fill, (unsigned long long)max_fill);
        if ( max_fill > ctrl->max_due ) atomicMax(&ctrl->max_due, (unsigned long long)max_fill);
    }

    // ---- the blocks left to the sequential form (Ctrl::fb_n entries of fb_list) are run here, by the CTAs that have
    // no block of their own left: every entry is appended by a CTA that comes through here afterwards, so none is
    // missed, and a block re-run by one CTA while others still work on theirs is no different from two blocks of the
    // event-parallel form running side by side.  An index is taken only when an entry exists (CAS), and the entry
    // itself is waited for: its writer bumps the count first.
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

// ---- MT19937: next generation of one generator by a warp.  gen = 624 words of shared memory owned by the warp.
// Chunks of 32 words in increasing order, reads of a chunk before its writes: word i reads the OLD words i, i+1 and
// (i < 227 ? the OLD word i+397 : the NEW word i-227), the last word the NEW word 0 -- what the reference's in-place
// loop reads (mersenne.cc:56-68).  src/dst 16-byte aligned.
__device__ __forceinline__ void
mt_warp_load(uint32_t* gen, const uint32_t* src, uint32_t lane)
{
    for ( uint32_t i = lane; i < MT_N / 4; i += 32 )
        reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(src)[i];
    __syncwarp();
}
// gen (shared memory) <- the generation after gen, in place
__device__ __forceinline__ void
mt_warp_twist(uint32_t* gen, uint32_t lane)
{
    for ( uint32_t c = 0; c < (MT_N + 31) / 32; ++c ) {
        const uint32_t i   = c * 32 + lane;
        const bool     act = i < (uint32_t)MT_N;
        uint32_t       v   = 0;
        if ( act ) v = mt_twist(gen[i], gen[i + 1 == (uint32_t)MT_N ? 0 : i + 1], gen[i + MT_M >= (uint32_t)MT_N ? i + MT_M - MT_N : i + MT_M]);
        __syncwarp();
        if ( act ) gen[i] = v;
        __syncwarp();
    }
}
__device__ __forceinline__ void
mt_warp_regen(uint32_t* gen, const uint32_t* src, uint32_t* dst, uint32_t lane)
{
    mt_warp_load(gen, src, lane);
    mt_warp_twist(gen, lane);
    for ( uint32_t i = lane; i < MT_N / 4; i += 32 )
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(gen)[i];
}
__device__ inline void
MeshElement::maintain(uint32_t op, uint8_t* rec, uint8_t* aux, const int64_t* param, uint32_t* gen, uint32_t lane)
{
    uint32_t* const buf = reinterpret_cast<uint32_t*>(aux);
    uint32_t* const dig = reinterpret_cast<uint32_t*>(rec + DIGEST_OFF);
    uint32_t        ng, ok, par;
    if ( op == MAINT_INIT ) { // the seed words are in the second buffer; the record has not been constructed yet
        ng  = (uint32_t)param[4];
        ok  = digestOk(param) ? 1u : 0u;
        par = 1;
    }
    else {
        const uint32_t pos = reinterpret_cast<const uint32_t*>(rec)[1];
        const uint32_t cst = reinterpret_cast<const uint32_t*>(rec)[4];
        ng                 = cst & 0xFFFFu;
        ok                 = cst >> 16;
        par                = pos >> MT_POS_SHIFT;
    }
    if ( op == MAINT_TAIL ) { // a generation has just become the current one: digest entries of its pairs [208,312)
        if ( !ok ) return;
        // only the words those entries are made of: the draws of pairs [208,312) are words [416,624) -- a third of the
        // generation (this read is 1 of the 3 x 2.5 KB a generator costs per 624 draws)
        constexpr uint32_t Q0 = ROUTE_DIGEST_SPLIT * 32 / 4; // first 16-byte piece
        static_assert(ROUTE_DIGEST_SPLIT * 32 % 4 == 0, "tail words start on a 16-byte piece");
        for ( uint32_t i = Q0 + lane; i < MT_N / 4; i += 32 )
            reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(buf + par * MT_N)[i];
        __syncwarp();
        if ( lane >= ROUTE_DIGEST_SPLIT && lane < ROUTE_DIGEST_WORDS ) dig[lane] = route_digest_word(gen, lane, ng);
        return;
    }
    mt_warp_regen(gen, buf + par * MT_N, buf + (par ^ 1u) * MT_N, lane);
    __syncwarp();
    if ( ok && lane < (op == MAINT_INIT ? ROUTE_DIGEST_WORDS : ROUTE_DIGEST_SPLIT) ) dig[lane] = route_digest_word(gen, lane, ng);
}

// after a window (list mode) or before setup() (init_all: every local component in [c_lo, c_hi)): a warp per entry.
// advance != 0 (list mode, the native run loop): the last CTA to finish ends the window -- 1 single rank
// (advance_fold), 2 peer mode (mailbox post, wait for the other ranks, fold): no separate launch for it.
__global__ void __launch_bounds__(256)
maintenance_kernel(Dev d, int init_all, uint32_t c_lo, uint32_t c_hi, int advance, int spin)
{
    __shared__ __align__(16) uint32_t s_gen[8][MT_N];
    __shared__ uint32_t               s_last;
    const uint32_t lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const uint32_t nw = gridDim.x * 8, gw = blockIdx.x * 8 + wic;
    uint32_t       n  = c_hi - c_lo;
    Ctrl* const    c  = d.ctrl;
    if ( !init_all ) {
        if ( c->done ) return;
        n = min(c->maint_n, d.n_local);
    }
    for ( uint32_t i = gw; i < n; i += nw ) {
        const uint32_t ent = init_all ? ((c_lo + i) | ((uint32_t)MAINT_INIT << 30)) : d.maint_list[i];
        const uint32_t lc = ent & 0x3FFFFFFFu, op = ent >> 30;
        with_element(d.uniform_type ? d.uniform_type : d.comp_type[lc], [&](auto el) {
            using E = decltype(el);
            if constexpr ( E::PARALLEL_EVENTS )
                E::maintain(op, d.state + (uint64_t)lc * d.state_stride, d.aux + (uint64_t)lc * d.aux_stride,
                    d.comp_param + (uint64_t)lc * NPARAM, s_gen[wic], lane);
        });
        __syncwarp();
    }
    if ( !advance ) return;
    __threadfence();
    __syncthreads();
    if ( threadIdx.x == 0 ) s_last = atomicAdd(&c->tail_done, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if ( !s_last ) return;
    __threadfence();
    if ( threadIdx.x == 0 ) c->tail_done = 0;
    if ( advance == 2 ) {
        peer_post(d);
        peer_wait_and_fold(d, spin);
    }
    else if ( threadIdx.x == 0 )
        advance_fold(d);
}

} // namespace sstb200
