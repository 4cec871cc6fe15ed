// kernel_sequencer.cuh -- the in-order placement sequencer (one CTA).
//
// Placement is inherently ordered: every reservation changes the node state the
// next decision reads (SURVEY.md hard part A).  This kernel walks a range of
// task groups in canonical order and, for each one, reproduces
//
//   nodeSet.tree (k best feasible nodes)       manager/scheduler/nodeset.go:50-124
//   decisionTree.orderedNodes (best -> worst)  manager/scheduler/decision_tree.go:24-52
//   scheduleNTasksOnNodes (fill + round-robin) manager/scheduler/scheduler.go:844-924
//   NodeInfo.addTask (the reservation)         manager/scheduler/nodeinfo.go:108-154
//   Pipeline failure counters / Explain        manager/scheduler/pipeline.go:56-103
//
// Two paths per group:
//  * fast path (k == 1, scan result available): the batched scan kernel already
//    evaluated this task against the node table as it stood when the batch began
//    and left the bitmap of its best rank class.  Within a batch node state only
//    gets worse, so the first class member not yet touched in this batch is the
//    sequentially-correct argmin; it is found with one masked find-first over the
//    bitmap.  Tasks are taken in chunks: descriptors and the first 32k-node
//    window of each bitmap are staged in shared memory together, so the ordered
//    part touches shared memory only.
//  * generic path (any k; also the fall-back when a class was consumed): full
//    table evaluation against the live state; k == 1: block arg-min; k > 1:
//    radix-select of the k smallest rank keys, bitonic sort, sequential fill on
//    staged rows, parallel write-back.
#pragma once
#include "kernels_common.cuh"

namespace pe {

struct CandKey {
    unsigned long long pref;
    uint32_t tie;
    uint32_t node;
};
__device__ __forceinline__ bool key_less(const CandKey &a, const CandKey &b) {
    return a.pref < b.pref || (a.pref == b.pref && a.tie < b.tie);
}

#define PE_SEQ_THREADS 512     // 128 registers per thread: the ordered fast path must not spill
#define PE_SEQ_KS 2048          // candidates staged in shared memory
#define PE_SEQ_RING 16          // fast-mode ring slots
#define PE_SEQ_NPW 8            // producer warps (warps 1..NPW)
#define PE_SEQ_WIN 1024         // bitmap words staged per task (32k nodes)
#define PE_MAX_GEN_WANTS 8
#define PE_CTX_MAXC 16
#define PE_ST_FAILED 1u
#define PE_ST_BLOCKED 2u


__device__ __forceinline__ uint32_t seq_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sq_mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(seq_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sq_mbar_inval(unsigned long long *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(seq_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sq_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(seq_smem_u32(bar)) : "memory");
}
// Arrive without release ordering: the consumer's empty-slot signal must not wait for its
// outstanding global reductions / stores to be acknowledged (that costs ~1 us per task).
__device__ __forceinline__ void sq_mbar_arrive_relaxed(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(seq_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool sq_mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(seq_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void sq_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(seq_smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void sq_tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(seq_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(seq_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sq_mbar_wait(unsigned long long *bar, uint32_t parity) {
    while (!sq_mbar_try_wait(bar, parity)) {}
}
// Same, but gives up after ~1 s and raises a device error instead of hanging the GPU.
__device__ __forceinline__ bool sq_mbar_wait_wd(unsigned long long *bar, uint32_t parity, DevCounters *ctr, uint32_t code) {
    const long long t0 = clock64();
    while (!sq_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) { atomicOr(&ctr->error, code); return false; }
    }
    return true;
}

struct SeqParams {
    DevTable T;
    TickDev K;
    uint32_t g_begin, g_end;
    const ScanResult *scan;  // [row] or nullptr
    const uint32_t *task_row; // [g_end - g_begin] task -> scan row (tasks with identical descriptors share one)
    const uint32_t *E;       // class bitmaps, two rows of e_stride words per scan row (best class, second class)
    uint32_t e_stride;
    const uint32_t *L;       // class member lists, [row][2][PE_LIST_CAP]
    uint8_t *ff8;                 // [cap] first failing filter (0 pass, 0xFF not in set)
    unsigned long long *pref64;   // [cap]
    CandKey *cand_g;              // [st_cap]
    int64_t *st_cpu_g, *st_mem_g; // [st_cap]
    uint32_t *st_svc_g, *st_tot_g, *st_placed_g;
    uint8_t *st_flags_g;
    int64_t *st_gen_g;            // [PE_MAX_GEN_WANTS][st_cap]
    uint32_t st_cap;
    uint32_t *touched_g;          // [touched_words] global fallback
    uint32_t touched_words;
    uint32_t touched_in_smem;
    DevCounters *ctr;
};

// Per-group evaluation context resolved once into shared memory so that a node
// evaluation is one level of independent global loads.
struct GroupCtx {
    uint32_t *svccol;
    const uint32_t *con_col[PE_CTX_MAXC];
    uint32_t con_val[PE_CTX_MAXC];
    uint32_t con_neq[PE_CTX_MAXC];
    uint32_t usable;   // 1: every constraint is staged here (con_cnt <= PE_CTX_MAXC)
};

struct FastTask {      // staged descriptor of one k=1 task
    unsigned long long c0;
    long long cpu_res, mem_res;
    uint32_t *svccol;
    uint32_t tie_start, task_off, simple, counts, ws, row;
    uint32_t n_list, n_class;   // list mode (tie_start == 0): listed / total members of the best class; n_list == 0: bitmap mode
};

struct SeqShared {
    pe_group G;
    GroupCtx C;
    FastTask ft[PE_SEQ_RING];
    unsigned long long full_bar[PE_SEQ_RING], empty_bar[PE_SEQ_RING];
    uint32_t stop, resume, stop_reason, bars_live, consumed;
    uint32_t armed[PE_SEQ_RING];
    uint32_t red32[40];
    unsigned long long red64[40];
    uint32_t bins[256];
    uint32_t cnt8[8];
    uint32_t bestv[3];   // rotating slots: see block_min_pos
    uint32_t sel_bin, sel_before, ncand, neutral, any_pass, done, dead;
    unsigned long long best_pref;
    uint32_t best_tie;
};

__device__ __forceinline__ uint32_t block_sum(uint32_t v, SeqShared &S) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane == 0) S.red32[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = lane < (blockDim.x >> 5) ? S.red32[lane] : 0;
        t = __reduce_add_sync(0xFFFFFFFFu, t);
        if (lane == 0) S.red32[32] = t;
    }
    __syncthreads();
    uint32_t r = S.red32[32];
    __syncthreads();
    return r;
}

// OR / AND of 64-bit and 32-bit words over the block (for digit skipping)
__device__ __forceinline__ void block_or_and(unsigned long long &o64, unsigned long long &a64, uint32_t &o32,
                                             uint32_t &a32, SeqShared &S) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t olo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)o64), ohi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(o64 >> 32));
    uint32_t alo = __reduce_and_sync(0xFFFFFFFFu, (uint32_t)a64), ahi = __reduce_and_sync(0xFFFFFFFFu, (uint32_t)(a64 >> 32));
    uint32_t ot = __reduce_or_sync(0xFFFFFFFFu, o32), at = __reduce_and_sync(0xFFFFFFFFu, a32);
    __syncthreads();
    if (lane == 0) {
        S.red64[warp] = ((unsigned long long)ohi << 32) | olo;
        S.red32[warp] = ot;
    }
    __syncthreads();
    unsigned long long O = 0; uint32_t Ot = 0;
    for (uint32_t w = 0; w < nw; w++) { O |= S.red64[w]; Ot |= S.red32[w]; }
    __syncthreads();
    if (lane == 0) {
        S.red64[warp] = ((unsigned long long)ahi << 32) | alo;
        S.red32[warp] = at;
    }
    __syncthreads();
    unsigned long long A = ~0ull; uint32_t At = ~0u;
    for (uint32_t w = 0; w < nw; w++) { A &= S.red64[w]; At &= S.red32[w]; }
    __syncthreads();
    o64 = O; a64 = A; o32 = Ot; a32 = At;
}

// Block-wide minimum of a candidate position; every thread gets the result.
// Three rotating slots make it safe with ONE barrier per call: the slot used
// by call i is re-armed during call i+1 (after that call's barrier) and used
// again by call i+3, so a re-arm can never race with an atomicMin or a read.
__device__ __forceinline__ uint32_t block_min_pos(uint32_t c, SeqShared &S, uint32_t &slot) {
    c = __reduce_min_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && c != PE_NONE) atomicMin(&S.bestv[slot], c);
    __syncthreads();
    const uint32_t r = S.bestv[slot];
    const uint32_t prev = slot == 0 ? 2u : slot - 1u;
    if (threadIdx.x == 0) S.bestv[prev] = PE_NONE;
    slot = slot == 2 ? 0u : slot + 1u;
    return r;
}

// First set bit of (E & ~touched) in tie order, restricted to words >= w0 (global-memory walk).
__device__ __forceinline__ uint32_t find_first(const uint32_t *Erow, uint32_t w0, uint32_t N, uint32_t ts,
                                               const uint32_t *touched, SeqShared &S, uint32_t &slot) {
    const uint32_t base_bit = w0 * 32u;
    for (int seg = 0; seg < 2; seg++) {
        uint32_t lo_bit, hi_bit;
        if (seg == 0) { lo_bit = ts > base_bit ? ts : base_bit; hi_bit = N; }
        else { lo_bit = base_bit; hi_bit = ts; }
        if (lo_bit >= hi_bit) continue;
        const uint32_t wlo = lo_bit >> 5, whi = (hi_bit + 31u) >> 5;
        for (uint32_t base = wlo; base < whi; base += blockDim.x) {
            const uint32_t w = base + threadIdx.x;
            uint32_t v = 0;
            if (w < whi) {
                v = Erow[w] & ~touched[w];
                if (w == wlo) v &= 0xFFFFFFFFu << (lo_bit & 31u);
                if (w == whi - 1 && (hi_bit & 31u)) v &= (1u << (hi_bit & 31u)) - 1u;
            }
            const uint32_t b = block_min_pos(v ? w * 32u + (uint32_t)__ffs((int)v) - 1u : PE_NONE, S, slot);
            if (b != PE_NONE) return b;
        }
    }
    return PE_NONE;
}

// Pipeline.Process for one node with the constraint columns pre-resolved.
__device__ __forceinline__ uint32_t eval_ctx(const DevTable &T, const TickDev &K, const pe_group &g, const GroupCtx &C, uint32_t n,
                                             uint32_t meta, uint32_t svc_n) {
    if (!C.usable || (g.filter_mask & ~((1u << PE_F_READY) | (1u << PE_F_CONSTRAINT) | (1u << PE_F_PLATFORM))) || g.ip_cnt ||
        (g.flags & PE_G_CONSTRAINT_NEVER))
        return eval_ff(T, K, g, n, meta, svc_n);
    // common case: Ready + attribute constraints + platform, all loads independent
    if ((g.filter_mask & (1u << PE_F_READY)) && !(meta & PE_NODE_READY)) return 1 + PE_F_READY;
    if (g.filter_mask & (1u << PE_F_CONSTRAINT)) {
        uint32_t v[PE_CTX_MAXC];
        const uint32_t cc = g.con_cnt;
#pragma unroll
        for (int e = 0; e < PE_CTX_MAXC; e++) v[e] = (uint32_t)e < cc ? C.con_col[e][n] : 0u;
        bool ok = true;
#pragma unroll
        for (int e = 0; e < PE_CTX_MAXC; e++)
            if ((uint32_t)e < cc) ok = ok && ((v[e] == C.con_val[e]) != (C.con_neq[e] != 0u));
        if (!ok) return 1 + PE_F_CONSTRAINT;
    }
    if ((g.filter_mask & (1u << PE_F_PLATFORM)) && g.plat_cnt) {
        bool ok = false;
        if (meta & PE_NODE_HAS_PLATFORM) {
            const uint32_t os = (meta >> 8) & 0xFF, arch = (meta >> 16) & 0xFF;
            for (uint32_t i = 0; i < g.plat_cnt && !ok; i++) {
                const pe_platform p = K.plats[g.plat_off + i];
                ok = (p.arch_id == 0 || p.arch_id == arch) && (p.os_id == 0 || p.os_id == os);
            }
        }
        if (!ok) return 1 + PE_F_PLATFORM;
    }
    return 0;
}


// ---- fast mode, consumer warp: tasks strictly in order, shared memory only on the common path.
// One warp executes a dependent chain (~7 cycles per instruction), so this loop is kept as short
// as possible: no per-task global loads, no block barriers, reservations as fire-and-forget reductions.
struct SeqDebug { unsigned long long n_fast, n_placed, iters, stops[5]; long long cyc_wait, cyc_work; };

__device__ __forceinline__ void fast_commit(const SeqParams &P, const FastTask &f, uint32_t gq, uint32_t n, uint32_t *touched) {
    const DevTable &T = P.T;
    P.K.out_node[f.task_off] = n;
    if (f.simple) {   // NodeInfo.addTask, nodeinfo.go:125-153, as fire-and-forget reductions
        if (f.cpu_res) atomicAdd(reinterpret_cast<unsigned long long *>(&T.cpu[n]), (unsigned long long)(-f.cpu_res));
        if (f.mem_res) atomicAdd(reinterpret_cast<unsigned long long *>(&T.mem[n]), (unsigned long long)(-f.mem_res));
        if (f.counts) { atomicAdd(&T.total[n], 1u); atomicAdd(&f.svccol[n], 1u); }
    } else {
        add_task_global(T, P.K, P.K.groups[gq], n, f.counts != 0, P.ctr);
    }
    touched[n >> 5] |= 1u << (n & 31u);
}

__device__ __forceinline__ void fast_consumer(const SeqParams &P, SeqShared &S, uint32_t *touched, const uint32_t *ring, uint32_t start,
                                              SeqDebug &dbg) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t N = P.T.n_nodes, nwords = (N + 31u) >> 5;
    uint32_t i = 0, reason = 0;
    for (; start + i < P.g_end; i++) {
        const uint32_t slot = i % PE_SEQ_RING, round = i / PE_SEQ_RING;
        if (!sq_mbar_wait_wd(&S.full_bar[slot], round & 1u, P.ctr, PE_DEV_ERR_WD_CONSUMER)) { reason = 1; break; }
        const FastTask &f = S.ft[slot];
        if (f.c0 == PE_PREF_NONE) { reason = 1; break; }      // nothing feasible when the batch began, or k != 1
        uint32_t n = PE_NONE;
        const uint32_t n_list = f.n_list;
        if (n_list) {
            // ---- list mode (canonical tie order): the first members of the class in node order
            const uint4 *lst4 = reinterpret_cast<const uint4 *>(ring + slot * PE_SEQ_WIN);
            for (uint32_t j = 0; j < n_list && n == PE_NONE; j += 128u) {
                const uint32_t il = j + lane * 4u;
                uint32_t first = PE_NONE;
                if (il < n_list) {
                    const uint4 e = lst4[il >> 2];
                    const uint32_t left = n_list - il;
                    const uint32_t tx = touched[e.x >> 5], ty = left > 1u ? touched[e.y >> 5] : ~0u;
                    const uint32_t tz = left > 2u ? touched[e.z >> 5] : ~0u, tw = left > 3u ? touched[e.w >> 5] : ~0u;
                    if (!((tw >> (e.w & 31u)) & 1u)) first = e.w;
                    if (!((tz >> (e.z & 31u)) & 1u)) first = e.z;
                    if (!((ty >> (e.y & 31u)) & 1u)) first = e.y;
                    if (!((tx >> (e.x & 31u)) & 1u)) first = e.x;
                }
                const uint32_t b = __ballot_sync(0xFFFFFFFFu, first != PE_NONE);
                if (b) n = __shfl_sync(0xFFFFFFFFu, first, __ffs((int)b) - 1);
            }
            if (n == PE_NONE && f.n_class > n_list) {
                // the class goes on past the listed members: continue on its bitmap (L2), 32 words a round
                const uint32_t last = ring[slot * PE_SEQ_WIN + n_list - 1u];
                const uint32_t *Erow = P.E + (size_t)f.row * 2u * P.e_stride;
                for (uint32_t wb = last >> 5; wb < nwords && n == PE_NONE; wb += 32u) {
                    const uint32_t w = wb + lane;
                    uint32_t v = 0;
                    if (w < nwords) {
                        v = Erow[w] & ~touched[w];
                        if (w == (last >> 5)) v &= (last & 31u) == 31u ? 0u : (0xFFFFFFFFu << ((last & 31u) + 1u));
                        if (w == nwords - 1 && (N & 31u)) v &= (1u << (N & 31u)) - 1u;
                    }
                    const uint32_t b = __ballot_sync(0xFFFFFFFFu, v != 0u);
                    if (b) {
                        const int src = __ffs((int)b) - 1;
                        const uint32_t vv = __shfl_sync(0xFFFFFFFFu, v, src);
                        n = (wb + (uint32_t)src) * 32u + (uint32_t)__ffs((int)vv) - 1u;
                    }
                }
            }
            if (n == PE_NONE) { reason = 4; break; }           // the whole class is consumed
        } else {
            // ---- bitmap mode (rotated tie order): the staged window of the class bitmap
            const uint32_t lo_bit = f.tie_start;
            const uint32_t lo_word = lo_bit >> 5;
            const uint32_t nwin = min((uint32_t)PE_SEQ_WIN, nwords - f.ws);    // f.ws: window start, multiple of 4 words
            const uint32_t *buf = ring + slot * PE_SEQ_WIN;
            for (uint32_t j = 0; j < nwin && n == PE_NONE; j += 32u) {
                const uint32_t w = f.ws + j + lane;
                uint32_t v = 0;
                if (j + lane < nwin && w >= lo_word) {
                    v = buf[j + lane] & ~touched[w];
                    if (w == lo_word) v &= 0xFFFFFFFFu << (lo_bit & 31u);
                    if (w == nwords - 1 && (N & 31u)) v &= (1u << (N & 31u)) - 1u;
                }
                const uint32_t b = __ballot_sync(0xFFFFFFFFu, v != 0u);
                if (b) {
                    const int src = __ffs((int)b) - 1;
                    const uint32_t vv = __shfl_sync(0xFFFFFFFFu, v, src);
                    n = (f.ws + j + (uint32_t)src) * 32u + (uint32_t)__ffs((int)vv) - 1u;
                }
            }
            if (n == PE_NONE) { reason = 2; break; }           // not inside the staged window (may wrap): block-wide walk
        }
        const bool counts = f.counts != 0u;
        if (lane == 0) {
            // out_fail rows were zeroed when the tick started.  For the common reservation (no generic
            // resources / host ports) only the choice is recorded here; the column updates (four
            // reductions per task) are applied by the whole block when fast mode ends -- nothing reads
            // those columns in between, and global atomics in this loop would serialise on their latency.
            if (f.simple) { P.K.out_node[f.task_off] = n; touched[n >> 5] |= 1u << (n & 31u); }
            else fast_commit(P, f, start + i, n, touched);
        }
        __syncwarp();
        if (lane == 0) sq_mbar_arrive_relaxed(&S.empty_bar[slot]);   // slot reads are done (their values were used above)
        if (!counts) { reason = 3; i++; break; }               // rank did not move: later class bitmaps may hide this node
    }
    const uint32_t committed = (reason == 3) ? i : i;          // tasks [start, start+i) were placed here
    dbg.n_fast += committed; dbg.n_placed += committed;
    dbg.stops[reason < 5 ? reason : 0]++;
    if (lane == 0) {
        S.resume = start + i;
        S.consumed = i;
        S.stop_reason = reason;
        if (reason == 3) S.neutral = 1;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&S.stop) = 1;
    }
}

// ---- fast mode, producer warps (warps 1..PE_SEQ_NPW): warp p prefetches the tasks i == p-1 (mod NPW);
// one converged lane per warp issues the TMA copy (the CUTLASS elect-one idiom).
__device__ __forceinline__ void fast_producer(const SeqParams &P, SeqShared &S, uint32_t *ring, uint32_t start) {
    const uint32_t lane = threadIdx.x & 31u, pw = (threadIdx.x >> 5) - 1u;
    const uint32_t nwords = (P.T.n_nodes + 31u) >> 5;
    const ScanResult *scan = P.scan; const uint32_t *E = P.E, *L = P.L;
    const uint32_t g_begin = P.g_begin, g_end = P.g_end, e_stride = P.e_stride;
    volatile uint32_t *vstop = &S.stop;
    for (uint32_t i = pw; start + i < g_end; i += PE_SEQ_NPW) {
        const uint32_t slot = i % PE_SEQ_RING, round = i / PE_SEQ_RING;
        const uint32_t gq = start + i;
        uint32_t go = 1;
        if (lane == 0) {
            const uint32_t row = P.task_row[gq - g_begin];
            const ScanResult sr = scan[row];                   // issued before the wait: overlaps it
            const uint32_t task_off = P.K.groups[gq].task_off;
            while (!sq_mbar_try_wait(&S.empty_bar[slot], (round & 1u) ^ 1u)) {
                if (*vstop) { go = 0; break; }
                __nanosleep(32);
            }
            if (*vstop) go = 0;
            if (go) {
                FastTask f;
                f.c0 = (sr.flags & PE_SR_K1) ? sr.c0 : PE_PREF_NONE;
                f.row = row;
                f.tie_start = sr.tie_start;
                f.task_off = task_off;
                f.cpu_res = sr.cpu_res; f.mem_res = sr.mem_res;
                f.simple = (sr.flags & PE_SR_SIMPLE) ? 1u : 0u;
                f.counts = (sr.flags & PE_SR_COUNTS) ? 1u : 0u;
                f.svccol = sr.svccol;
                f.ws = (sr.tie_start >> 5) & ~3u;                       // 16-byte aligned window start
                f.n_class = sr.n0;
                f.n_list = (sr.tie_start == 0u && f.c0 != PE_PREF_NONE) ? min(sr.n0, (uint32_t)PE_LIST_CAP) : 0u;
                S.ft[slot] = f;
                S.armed[slot] = i + 1u;
                if (f.n_list) {
                    const uint32_t bytes = ((f.n_list + 3u) & ~3u) * 4u;
                    sq_mbar_expect_tx(&S.full_bar[slot], bytes);
                    sq_tma_bulk_g2s(ring + slot * PE_SEQ_WIN, L + (size_t)row * 2u * PE_LIST_CAP, bytes, &S.full_bar[slot]);
                } else if (f.c0 != PE_PREF_NONE && f.ws < nwords) {
                    // rows are padded to e_stride (a multiple of 32 words), so the copy may run past nwords
                    const uint32_t nw = min((uint32_t)PE_SEQ_WIN, e_stride - f.ws);
                    const uint32_t *src = E + (size_t)row * 2u * e_stride + f.ws;
                    sq_mbar_expect_tx(&S.full_bar[slot], nw * 4u);
                    sq_tma_bulk_g2s(ring + slot * PE_SEQ_WIN, src, nw * 4u, &S.full_bar[slot]);
                } else {
                    sq_mbar_arrive(&S.full_bar[slot]);
                }
            }
        }
        go = __shfl_sync(0xFFFFFFFFu, go, 0);    // the whole warp leaves together (and reaches the block barrier converged)
        if (!go) break;
    }
}

__global__ void __launch_bounds__(PE_SEQ_THREADS, 1) k_sequencer(const __grid_constant__ SeqParams P) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ SeqShared S;
    __shared__ uint32_t hist[8][256];

    const DevTable &T = P.T;
    const TickDev &K = P.K;
    const uint32_t tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t N = T.n_nodes;

    // ---- carve dynamic shared memory
    CandKey *cand_s = reinterpret_cast<CandKey *>(dyn_smem);
    int64_t *st_cpu_s = reinterpret_cast<int64_t *>(cand_s + PE_SEQ_KS);
    int64_t *st_mem_s = st_cpu_s + PE_SEQ_KS;
    uint32_t *st_svc_s = reinterpret_cast<uint32_t *>(st_mem_s + PE_SEQ_KS);
    uint32_t *st_tot_s = st_svc_s + PE_SEQ_KS;
    uint32_t *st_placed_s = st_tot_s + PE_SEQ_KS;
    uint8_t *st_flags_s = reinterpret_cast<uint8_t *>(st_placed_s + PE_SEQ_KS);
    uint32_t *touched_s = reinterpret_cast<uint32_t *>(st_flags_s + PE_SEQ_KS);
    uint32_t *touched = P.touched_in_smem ? touched_s : P.touched_g;
    uint32_t *ring = touched_s + (P.touched_in_smem ? ((P.touched_words + 3u) & ~3u) : 0u);   // [PE_SEQ_RING][PE_SEQ_WIN]

    for (uint32_t w = tid; w < P.touched_words; w += nth) touched[w] = 0;
    if (tid == 0) { S.neutral = 0; S.bars_live = 0; S.bestv[0] = S.bestv[1] = S.bestv[2] = PE_NONE; }
    uint32_t slot = 0;   // uniform across the block
    unsigned long long n_fast = 0, n_medium = 0, n_slow = 0, n_placed = 0, n_evalg = 0;
    long long cyc_fast = 0, cyc_medium = 0, cyc_generic = 0, t_mark = clock64();
    SeqDebug dbg{};   // thread 0's private tallies
    __syncthreads();

    uint32_t gi = P.g_begin;
    while (gi < P.g_end) {
        // ================= fast mode: warp-specialised pipeline over k == 1 tasks ====
        // Producer warps prefetch, per task, the scan record and the first PE_SEQ_WIN
        // words of its best-class bitmap (one TMA bulk copy, completion on the slot's
        // mbarrier) into a ring of shared-memory slots; warp 0 consumes the slots IN
        // ORDER, so the ordered part touches shared memory only.  The mode ends at the
        // first task the bitmaps cannot resolve; that task takes the block-wide path.
        if (P.scan != nullptr && !S.neutral) {
            __syncthreads();
            { const long long t1 = clock64(); cyc_generic += t1 - t_mark; t_mark = t1; }
            if (tid == 0) {
                for (int r = 0; r < PE_SEQ_RING; r++) {
                    if (S.bars_live) { sq_mbar_inval(&S.full_bar[r]); sq_mbar_inval(&S.empty_bar[r]); }
                    sq_mbar_init(&S.full_bar[r], 1);
                    sq_mbar_init(&S.empty_bar[r], 1);
                    S.armed[r] = 0;
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                S.bars_live = 1;
                S.stop = 0;
                S.resume = P.g_end;
                S.stop_reason = 0;
            }
            __syncthreads();
            const uint32_t start = gi;
            const uint32_t nwords = (N + 31u) >> 5;
            volatile uint32_t *vstop = &S.stop;
            if (warp == 0) fast_consumer(P, S, touched, ring, start, dbg);
            else if (warp <= PE_SEQ_NPW) fast_producer(P, S, ring, start);
            __syncthreads();
            // bulk copies armed for slots the consumer never took must land before the ring is reused
            if (warp == 0) {
                for (uint32_t r = 0; r < PE_SEQ_RING; r++) {     // converged: every lane waits on the same slot
                    const uint32_t a = S.armed[r];
                    if (a != 0u && a - 1u >= S.consumed) sq_mbar_wait_wd(&S.full_bar[r], ((a - 1u) / PE_SEQ_RING) & 1u, P.ctr, PE_DEV_ERR_WD_DRAIN);
                }
            }
            __syncthreads();
            // apply the deferred reservations of the tasks fast mode placed: [start, start + consumed)
            for (uint32_t q = start + tid; q < start + S.consumed; q += nth) {
                const ScanResult sr = P.scan[P.task_row[q - P.g_begin]];
                if (!(sr.flags & PE_SR_SIMPLE)) continue;              // already applied in place
                const uint32_t n = K.out_node[K.groups[q].task_off];
                if (sr.cpu_res) atomicAdd(reinterpret_cast<unsigned long long *>(&T.cpu[n]), (unsigned long long)(-sr.cpu_res));
                if (sr.mem_res) atomicAdd(reinterpret_cast<unsigned long long *>(&T.mem[n]), (unsigned long long)(-sr.mem_res));
                if (sr.flags & PE_SR_COUNTS) { atomicAdd(&T.total[n], 1u); atomicAdd(&sr.svccol[n], 1u); }
            }
            __syncthreads();
            { const long long t1 = clock64(); cyc_fast += t1 - t_mark; t_mark = t1; }
            gi = S.resume;
            if (gi >= P.g_end) break;
            // group gi takes the block-wide path
        }

        // ================= generic path (one group) ===============================
        __syncthreads();
        if (tid < sizeof(pe_group) / 4) reinterpret_cast<uint32_t *>(&S.G)[tid] = reinterpret_cast<const uint32_t *>(&K.groups[gi])[tid];
        if (tid < 8) S.cnt8[tid] = 0;
        __syncthreads();
        const pe_group &G = S.G;
        const uint32_t k = G.n_tasks;
        uint32_t *ofail = K.out_fail + (size_t)gi * PE_NUM_FILTERS;
        const uint32_t this_gi = gi;
        gi++;
        if (k == 0) { if (tid < 8) ofail[tid] = 0; continue; }
        // resolve the evaluation context
        if (tid < PE_CTX_MAXC && tid < G.con_cnt) {
            const pe_constraint c = K.cons[G.con_off + tid];
            S.C.con_col[tid] = T.attr[c.col];
            S.C.con_val[tid] = c.value;
            S.C.con_neq[tid] = c.neq;
        }
        if (tid == 32) { S.C.svccol = T.svc[G.svc_id]; S.C.usable = G.con_cnt <= PE_CTX_MAXC ? 1u : 0u; }
        __syncthreads();
        uint32_t *svccol = S.C.svccol;

        if (k == 1 && P.scan != nullptr && !S.neutral) {
            // ---- medium path: the best class of this task was consumed earlier in the
            // batch.  Every node outside the two recorded classes ranked strictly worse
            // than the second class when the batch began and ranks can only have grown,
            // so the arg-min is either the first untouched member of the second class or
            // a touched member of the best class re-evaluated against the live state.
            const uint32_t srow = P.task_row[this_gi - P.g_begin];
            const ScanResult sr = P.scan[srow];
            const bool exhausted = S.stop_reason == 4u && S.resume == this_gi;   // the pipeline walked the complete member list
            if (sr.c0 != PE_PREF_NONE && !exhausted) {
                // (i) the best class may continue beyond the window / list the pipeline staged
                const uint32_t *E1w = P.E + (size_t)srow * 2u * P.e_stride;
                const uint32_t n1 = find_first(E1w, 0u, N, G.tie_start, touched, S, slot);
                if (n1 != PE_NONE) {
                    if (tid == 0) {
                        const bool counts = (K.task_flags[G.task_off] & PE_T_COUNTS) != 0;
                        K.out_node[G.task_off] = n1;
                        add_task_global(T, K, G, n1, counts, P.ctr);
                        touched[n1 >> 5] |= 1u << (n1 & 31u);
                        if (!counts) S.neutral = 1;
                        n_fast++; n_placed++;
                    }
                    if (tid < 8) ofail[tid] = 0;
                    __syncthreads();
                    { const long long t1 = clock64(); cyc_medium += t1 - t_mark; t_mark = t1; }
                    continue;
                }
            }
            // (ii) best class consumed: second class + touched members of the best class
            if (sr.c0 != PE_PREF_NONE && sr.c1 != PE_PREF_NONE) {
                const uint32_t *E1 = P.E + (size_t)srow * 2u * P.e_stride;
                const uint32_t *E2 = E1 + P.e_stride;
                uint32_t n2 = PE_NONE;
                const bool use_lists = G.tie_start == 0u;
                const uint32_t *L1 = P.L + (size_t)srow * 2u * PE_LIST_CAP;
                const uint32_t *L2 = L1 + PE_LIST_CAP;
                if (use_lists && sr.n1 > 0u) {
                    const uint32_t nl = min(sr.n1, (uint32_t)PE_LIST_CAP);
                    uint32_t c = PE_NONE;
                    for (uint32_t t = tid; t < nl && c == PE_NONE; t += nth) { const uint32_t p2 = L2[t]; if (!((touched[p2 >> 5] >> (p2 & 31u)) & 1u)) c = t; }
                    const uint32_t at = block_min_pos(c, S, slot);
                    if (at != PE_NONE) n2 = L2[at];
                    else if (sr.n1 > nl) n2 = find_first(E2, 0u, N, G.tie_start, touched, S, slot);
                } else {
                    n2 = find_first(E2, 0u, N, G.tie_start, touched, S, slot);
                }
                if (n2 != PE_NONE) {
                    unsigned long long bp = sr.c1;
                    uint32_t bt = tie_pos(n2, G.tie_start, N), bn = n2;
                    auto consider = [&](uint32_t n) {
                        const uint32_t meta = T.meta[n];
                        const uint32_t sv = svccol[n];
                        if (eval_ctx(T, K, G, S.C, n, meta, sv) != 0) return;
                        const uint32_t fails = G.fail_cnt ? fail_count(K, G, n) : 0u;
                        const unsigned long long pref = make_pref(fails, sv, T.total[n]);
                        const uint32_t tp = tie_pos(n, G.tie_start, N);
                        if (pref < bp || (pref == bp && tp < bt)) { bp = pref; bt = tp; bn = n; }
                    };
                    if (use_lists && sr.n0 <= (uint32_t)PE_LIST_CAP) {
                        // the complete best class is listed: one member per thread
                        for (uint32_t t = tid; t < sr.n0; t += nth) {
                            const uint32_t n = L1[t];
                            if ((touched[n >> 5] >> (n & 31u)) & 1u) consider(n);
                        }
                    } else {
                        const uint32_t nw = (N + 31u) >> 5;
                        for (uint32_t w = tid; w < nw; w += nth) {
                            uint32_t v = E1[w] & touched[w];
                            if (w == nw - 1 && (N & 31u)) v &= (1u << (N & 31u)) - 1u;
                            while (v) {
                                const uint32_t n = w * 32u + (uint32_t)__ffs((int)v) - 1u;
                                v &= v - 1u;
                                consider(n);
                            }
                        }
                    }
                    const uint32_t hi = (uint32_t)(bp >> 32), lo = (uint32_t)bp;
                    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
                    const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
                    const uint32_t mt = __reduce_min_sync(0xFFFFFFFFu, (hi == mh && lo == ml) ? bt : 0xFFFFFFFFu);
                    __syncthreads();
                    if (lane == 0) { S.red64[warp] = ((unsigned long long)mh << 32) | ml; S.red32[warp] = mt; }
                    __syncthreads();
                    unsigned long long gp = ~0ull; uint32_t gt = ~0u;
                    for (uint32_t w = 0; w < (nth >> 5); w++) {
                        const unsigned long long p = S.red64[w]; const uint32_t t = S.red32[w];
                        if (p < gp || (p == gp && t < gt)) { gp = p; gt = t; }
                    }
                    // every thread starts from (c1, n2), so several may hold the winner: lowest thread commits
                    const uint32_t who = __reduce_min_sync(0xFFFFFFFFu, (bp == gp && bt == gt) ? tid : 0xFFFFFFFFu);
                    if (lane == 0) S.red32[warp] = who;
                    __syncthreads();
                    uint32_t winner = 0xFFFFFFFFu;
                    for (uint32_t w = 0; w < (nth >> 5); w++) winner = min(winner, S.red32[w]);
                    if (tid == winner) {
                        const bool counts = (K.task_flags[G.task_off] & PE_T_COUNTS) != 0;
                        K.out_node[G.task_off] = bn;
                        add_task_global(T, K, G, bn, counts, P.ctr);
                        touched[bn >> 5] |= 1u << (bn & 31u);
                        if (!counts) S.neutral = 1;
                    }
                    if (tid == 0) { n_medium++; n_placed++; }
                    if (tid < 8) ofail[tid] = 0;
                    __syncthreads();
                    { const long long t1 = clock64(); cyc_medium += t1 - t_mark; t_mark = t1; }
                    continue;
                }
            }
        }
        if (k == 1) {
            // ---- one task: block arg-min over the live table (nodeset.go:107-120 with a heap of one)
            unsigned long long bp = ~0ull;
            uint32_t bt = ~0u, bn = PE_NONE, myF = 0;
            uint32_t c[PE_NUM_FILTERS];
            for (int f = 0; f < PE_NUM_FILTERS; f++) c[f] = 0;
#pragma unroll 2
            for (uint32_t n = tid; n < N; n += nth) {
                const uint32_t meta = T.meta[n];
                if (!(meta & PE_NODE_VALID)) continue;
                const uint32_t sv = svccol[n];
                const uint32_t tot = T.total[n];
                const uint32_t ff = eval_ctx(T, K, G, S.C, n, meta, sv);
                for (int f = 0; f < PE_NUM_FILTERS; f++) c[f] += (ff == (uint32_t)(f + 1));
                if (ff == 0) {
                    const uint32_t fails = G.fail_cnt ? fail_count(K, G, n) : 0u;
                    const unsigned long long pref = make_pref(fails, sv, tot);
                    const uint32_t tp = tie_pos(n, G.tie_start, N);
                    myF++;
                    if (pref < bp || (pref == bp && tp < bt)) { bp = pref; bt = tp; bn = n; }
                }
            }
            // block arg-min on (pref, tie)
            const uint32_t hi = (uint32_t)(bp >> 32), lo = (uint32_t)bp;
            uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
            uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
            uint32_t mt = __reduce_min_sync(0xFFFFFFFFu, (hi == mh && lo == ml) ? bt : 0xFFFFFFFFu);
            if (lane == 0) { S.red64[warp] = ((unsigned long long)mh << 32) | ml; S.red32[warp] = mt; }
            for (int f = 0; f < PE_NUM_FILTERS; f++) {
                const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, c[f]);
                if (lane == 0 && s) atomicAdd(&S.cnt8[f], s);
            }
            __syncthreads();
            unsigned long long gp = ~0ull; uint32_t gt = ~0u;
            for (uint32_t w = 0; w < (nth >> 5); w++) {
                const unsigned long long p = S.red64[w]; const uint32_t t = S.red32[w];
                if (p < gp || (p == gp && t < gt)) { gp = p; gt = t; }
            }
            const bool mine = bn != PE_NONE && bp == gp && bt == gt;   // unique: tie positions are distinct
            if (gp == ~0ull) {
                if (tid == 0) { K.out_node[G.task_off] = PE_NONE; n_slow++; n_evalg += N; }
                __syncthreads();
                if (tid < 8) ofail[tid] = S.cnt8[tid];
            } else {
                if (mine) {
                    const bool counts = (K.task_flags[G.task_off] & PE_T_COUNTS) != 0;
                    K.out_node[G.task_off] = bn;
                    add_task_global(T, K, G, bn, counts, P.ctr);
                    if (P.touched_words) atomicOr(&touched[bn >> 5], 1u << (bn & 31u));
                    if (!counts) S.neutral = 1;
                }
                if (tid == 0) { n_slow++; n_placed++; n_evalg += N; }
                if (tid < 8) ofail[tid] = 0;
            }
            (void)myF; (void)this_gi;
            __syncthreads();
            continue;
        }

        // ---- 1. evaluate every node against the live state (nodeset.go:57-121)
        uint32_t myF = 0;
        unsigned long long o64 = 0, a64 = ~0ull;
        uint32_t o32 = 0, a32 = ~0u;
        for (uint32_t n = tid; n < N; n += nth) {
            const uint32_t meta = T.meta[n];
            if (!(meta & PE_NODE_VALID)) { P.ff8[n] = 0xFF; continue; }
            const uint32_t sv = svccol[n];
            const uint32_t ff = eval_ctx(T, K, G, S.C, n, meta, sv);
            const uint32_t fails = G.fail_cnt ? fail_count(K, G, n) : 0u;
            const unsigned long long pref = make_pref(fails, sv, T.total[n]);
            P.ff8[n] = (uint8_t)ff;
            P.pref64[n] = pref;
            if (ff == 0) {
                myF++;
                const uint32_t tp = tie_pos(n, G.tie_start, N);
                o64 |= pref; a64 &= pref; o32 |= tp; a32 &= tp;
            }
        }
        const uint32_t F = block_sum(myF, S);
        if (tid == 0) { n_evalg += N; n_slow++; }
        const uint32_t m = F < k ? F : k;

        // ---- 2. radix-select the k-th smallest (pref, tie) key among feasible nodes
        unsigned long long thr_pref = ~0ull;
        uint32_t thr_tie = ~0u;
        if (F > k) {
            block_or_and(o64, a64, o32, a32, S);
            unsigned long long dec_p = 0, msk_p = 0;
            uint32_t dec_t = 0, msk_t = 0;
            uint32_t rank = k;
            for (int b = 11; b >= 0; b--) {
                const bool inp = b >= 4;
                const int sh = inp ? (b - 4) * 8 : b * 8;
                const uint32_t vary = inp ? (uint32_t)(((o64 ^ a64) >> sh) & 0xFF) : ((o32 ^ a32) >> sh) & 0xFF;
                if (vary == 0) {  // every feasible key has the same byte here
                    if (inp) { dec_p |= a64 & (0xFFull << sh); msk_p |= 0xFFull << sh; }
                    else { dec_t |= a32 & (0xFFu << sh); msk_t |= 0xFFu << sh; }
                    continue;
                }
                for (uint32_t i = tid; i < 8 * 256; i += nth) (&hist[0][0])[i] = 0;
                __syncthreads();
                for (uint32_t n = tid; n < N; n += nth) {
                    if (P.ff8[n] != 0) continue;
                    const unsigned long long pr = P.pref64[n];
                    const uint32_t tp = tie_pos(n, G.tie_start, N);
                    if ((pr & msk_p) != dec_p || (tp & msk_t) != dec_t) continue;
                    const uint32_t d = inp ? (uint32_t)((pr >> sh) & 0xFF) : (tp >> sh) & 0xFF;
                    atomicAdd(&hist[warp & 7][d], 1u);
                }
                __syncthreads();
                if (tid < 256) {
                    uint32_t s = 0;
                    for (int h = 0; h < 8; h++) s += hist[h][tid];
                    S.bins[tid] = s;
                }
                __syncthreads();
                if (warp == 0) {
                    uint32_t loc[8], sum = 0;
                    for (int j = 0; j < 8; j++) { loc[j] = S.bins[lane * 8 + j]; sum += loc[j]; }
                    uint32_t incl = sum;
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if ((int)lane >= d) incl += t;
                    }
                    const uint32_t excl = incl - sum;
                    if (rank > excl && rank <= incl) {
                        uint32_t cc = excl;
                        for (int j = 0; j < 8; j++) {
                            if (rank <= cc + loc[j]) { S.sel_bin = lane * 8 + j; S.sel_before = cc; break; }
                            cc += loc[j];
                        }
                    }
                }
                __syncthreads();
                rank -= S.sel_before;
                if (inp) { dec_p |= (unsigned long long)S.sel_bin << sh; msk_p |= 0xFFull << sh; }
                else { dec_t |= S.sel_bin << sh; msk_t |= 0xFFu << sh; }
                __syncthreads();
            }
            thr_pref = dec_p;
            thr_tie = dec_t;
        }

        // ---- 3. collect + 4. sort the candidates (decision_tree.go:24-52)
        uint32_t m2 = 1;
        while (m2 < m) m2 <<= 1;
        const bool in_smem = m2 <= PE_SEQ_KS;
        CandKey *cand = in_smem ? cand_s : P.cand_g;
        int64_t *st_cpu = in_smem ? st_cpu_s : P.st_cpu_g;
        int64_t *st_mem = in_smem ? st_mem_s : P.st_mem_g;
        uint32_t *st_svc = in_smem ? st_svc_s : P.st_svc_g;
        uint32_t *st_tot = in_smem ? st_tot_s : P.st_tot_g;
        uint32_t *st_placed = in_smem ? st_placed_s : P.st_placed_g;
        uint8_t *st_flags = in_smem ? st_flags_s : P.st_flags_g;
        if (tid == 0) { S.ncand = 0; S.done = 0; S.any_pass = 0; S.dead = 0; }
        __syncthreads();
        if (m > 0) {
            for (uint32_t n = tid; n < N; n += nth) {
                if (P.ff8[n] != 0) continue;
                const unsigned long long pr = P.pref64[n];
                const uint32_t tp = tie_pos(n, G.tie_start, N);
                if (pr < thr_pref || (pr == thr_pref && tp <= thr_tie)) {
                    const uint32_t slot = atomicAdd(&S.ncand, 1u);
                    if (slot < m2) { cand[slot].pref = pr; cand[slot].tie = tp; cand[slot].node = n; }
                }
            }
            for (uint32_t i = m + tid; i < m2; i += nth) { cand[i].pref = ~0ull; cand[i].tie = ~0u; cand[i].node = PE_NONE; }
            __syncthreads();
            for (uint32_t size = 2; size <= m2; size <<= 1) {
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t i = tid; i < (m2 >> 1); i += nth) {
                        const uint32_t lo = 2 * i - (i & (stride - 1));
                        const uint32_t hi = lo + stride;
                        const bool up = (lo & size) == 0;
                        CandKey a = cand[lo], b = cand[hi];
                        if (key_less(b, a) == up) { cand[lo] = b; cand[hi] = a; }
                    }
                    __syncthreads();
                }
            }
            // ---- stage the candidates' dynamic state
            for (uint32_t i = tid; i < m; i += nth) {
                const uint32_t n = cand[i].node;
                st_cpu[i] = T.cpu[n];
                st_mem[i] = T.mem[n];
                st_svc[i] = svccol[n];
                st_tot[i] = T.total[n];
                st_placed[i] = 0;
                st_flags[i] = 0;
                for (uint32_t w = 0; w < G.gen_cnt; w++)
                    if (gen_first_occurrence(K, G, w))
                        P.st_gen_g[(size_t)w * P.st_cap + i] = T.gen[K.gens[G.gen_off + w].kind][n];
            }
        }
        __syncthreads();

        // ---- 5. scheduleNTasksOnNodes (scheduler.go:844-924), one thread, staged rows
        if (tid == 0 && m > 0) {
            const uint32_t fm = G.filter_mask;
            const bool f_res = (fm >> PE_F_RESOURCE) & 1u, f_port = ((fm >> PE_F_HOSTPORT) & 1u) && G.port_cnt > 0;
            const bool f_max = (fm >> PE_F_MAXREPLICAS) & 1u;
            uint32_t cnt[PE_NUM_FILTERS];
            for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] = 0;
            uint32_t done = 0, any_pass = 0, neutral = 0;
            unsigned long long it = 0;
            for (uint32_t ti = 0; ti < k; ti++) {
                const uint32_t i = (uint32_t)(it % m);
                K.out_node[G.task_off + ti] = cand[i].node;
                const bool counts = (K.task_flags[G.task_off + ti] & PE_T_COUNTS) != 0;
                // NodeInfo.addTask on the staged row (nodeinfo.go:125-153)
                st_mem[i] -= G.mem_res;
                st_cpu[i] -= G.cpu_res;
                for (uint32_t w = 0; w < G.gen_cnt; w++)
                    if (gen_first_occurrence(K, G, w)) {
                        int64_t *cell = &P.st_gen_g[(size_t)w * P.st_cap + i];
                        *cell = claim_cell(K, G, w, *cell);
                    }
                if (G.port_cnt) st_flags[i] |= PE_ST_BLOCKED;
                if (counts) { st_svc[i]++; st_tot[i]++; } else neutral = 1;
                st_placed[i]++;
                done++;
                if (done == k) break;
                if (it + 1 < m) {  // :899-905 first pass: move on only if the next node is now strictly better
                    const uint32_t j = (uint32_t)((it + 1) % m);
                    const uint32_t fa = (uint32_t)(cand[j].pref >> 56), fb = (uint32_t)(cand[i].pref >> 56);
                    bool less = fa != fb ? fa < fb : (st_svc[j] != st_svc[i] ? st_svc[j] < st_svc[i] : st_tot[j] < st_tot[i]);
                    if (less) it++;
                } else {
                    it++;          // :906-910 later passes: round-robin
                }
                const unsigned long long start = it;
                bool dead = false;
                for (;;) {         // :912-920
                    const uint32_t j = (uint32_t)(it % m);
                    bool ok = !(st_flags[j] & PE_ST_FAILED);
                    if (ok) {      // Pipeline.Process on the staged row: only the dynamic filters can have changed
                        int ff = -1;
                        if (f_res) {
                            if (G.cpu_res > st_cpu[j] || G.mem_res > st_mem[j]) ff = PE_F_RESOURCE;
                            for (uint32_t w = 0; w < G.gen_cnt && ff < 0; w++) {
                                uint32_t fw = w;
                                const uint32_t kind = K.gens[G.gen_off + w].kind;
                                for (uint32_t x = 0; x < w; x++)
                                    if (K.gens[G.gen_off + x].kind == kind) { fw = x; break; }
                                if (!gen_enough(P.st_gen_g[(size_t)fw * P.st_cap + j], K.gens[G.gen_off + w].value)) ff = PE_F_RESOURCE;
                            }
                        }
                        if (ff < 0 && f_port && (st_flags[j] & PE_ST_BLOCKED)) ff = PE_F_HOSTPORT;
                        if (ff < 0 && f_max && !((unsigned long long)st_svc[j] < G.max_replicas)) ff = PE_F_MAXREPLICAS;
                        if (ff >= 0) { cnt[ff]++; ok = false; }
                        else { for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] = 0; any_pass = 1; }
                    }
                    if (ok) break;
                    st_flags[j] |= PE_ST_FAILED;
                    it++;
                    if (it - start == m) { dead = true; break; }
                }
                if (dead) break;
            }
            S.done = done;
            S.any_pass = any_pass;
            if (neutral) S.neutral = 1;
            for (int f = 0; f < PE_NUM_FILTERS; f++) S.cnt8[f] = cnt[f];
            n_placed += done;
        }
        __syncthreads();
        const uint32_t done = S.done;

        // ---- 6. write the staged rows back
        for (uint32_t i = tid; i < m; i += nth) {
            if (!st_placed[i]) continue;
            const uint32_t n = cand[i].node;
            T.cpu[n] = st_cpu[i];
            T.mem[n] = st_mem[i];
            T.total[n] = st_tot[i];
            svccol[n] = st_svc[i];
            if (st_svc[i] >= 0xFFFFF0u) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
            for (uint32_t w = 0; w < G.gen_cnt; w++)
                if (gen_first_occurrence(K, G, w))
                    T.gen[K.gens[G.gen_off + w].kind][n] = P.st_gen_g[(size_t)w * P.st_cap + i];
            for (uint32_t p = 0; p < G.port_cnt; p++) {
                const uint32_t s = K.ports[G.port_off + p];
                T.ports[s >> 5][n] |= 1u << (s & 31u);
            }
            if (P.touched_words) atomicOr(&touched[n >> 5], 1u << (n & 31u));
        }
        for (uint32_t ti = done + tid; ti < k; ti += nth) K.out_node[G.task_off + ti] = PE_NONE;

        // ---- 7. Explain counters for unplaced tasks (pipeline.go:56-68).  The tree
        // build leaves "first failing filter" counts of the nodes visited after the
        // last one that entered the heap; a node is visited while the heap is not
        // full or when it ranks below the heap's worst member (nodeset.go:111-120).
        if (done < k) {
            if (!S.any_pass) {
                uint32_t posL = 0;
                bool haveL = false;
                unsigned long long Mp = ~0ull;
                uint32_t Mt = ~0u;
                if (m > 0) {
                    // L = heap member visited last; M = the heap's worst key
                    uint32_t mx = 0;
                    for (uint32_t i = tid; i < m; i += nth) mx = max(mx, cand[i].tie);
                    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
                    if (lane == 0) S.red32[warp] = mx;
                    __syncthreads();
                    for (uint32_t w = 0; w < (nth >> 5); w++) posL = max(posL, S.red32[w]);
                    __syncthreads();
                    haveL = true;
                    if (m == k) { Mp = cand[m - 1].pref; Mt = cand[m - 1].tie; }
                }
                uint32_t c[PE_NUM_FILTERS];
                for (int f = 0; f < PE_NUM_FILTERS; f++) c[f] = 0;
                for (uint32_t n = tid; n < N; n += nth) {
                    const uint32_t ff = P.ff8[n];
                    if (ff == 0 || ff == 0xFF) continue;
                    const uint32_t tp = tie_pos(n, G.tie_start, N);
                    if (haveL && tp <= posL) continue;
                    if (m == k) {
                        const unsigned long long pr = P.pref64[n];
                        if (!(pr < Mp || (pr == Mp && tp < Mt))) continue;
                    }
                    for (int f = 0; f < PE_NUM_FILTERS; f++) c[f] += (ff == (uint32_t)(f + 1));
                }
                for (int f = 0; f < PE_NUM_FILTERS; f++) {
                    const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, c[f]);
                    if (lane == 0 && s) atomicAdd(&S.cnt8[f], s);
                }
                __syncthreads();
            }
            if (tid < 8) ofail[tid] = S.cnt8[tid];
        } else {
            if (tid < 8) ofail[tid] = 0;
        }
    }
    if (tid == 0) {
        n_fast += dbg.n_fast; n_placed += dbg.n_placed;
        P.ctr->fast_path += n_fast;
        P.ctr->medium_path += n_medium;
        cyc_generic += clock64() - t_mark;   // whatever is left is the generic path (and loop overhead)
        P.ctr->cyc_fast += (unsigned long long)cyc_fast;
        P.ctr->cyc_medium += (unsigned long long)cyc_medium;
        P.ctr->cyc_generic += (unsigned long long)cyc_generic;
        P.ctr->cyc_cons_wait += (unsigned long long)dbg.cyc_wait; P.ctr->cyc_cons_work += (unsigned long long)dbg.cyc_work; P.ctr->iters += dbg.iters;
        for (int r = 0; r < 5; r++) P.ctr->stops[r] += dbg.stops[r];
        P.ctr->slow_path += n_slow;
        P.ctr->placements += n_placed;
        P.ctr->evals_generic += n_evalg;
    }
}

static inline size_t seq_dyn_smem_bytes(uint32_t touched_words_in_smem) {
    size_t a = (size_t)PE_SEQ_KS * (sizeof(CandKey) + 8 + 8 + 4 + 4 + 4 + 1);
    size_t b = (size_t)PE_SEQ_RING * PE_SEQ_WIN * 4;    // fast-mode ring
    return a + (((size_t)touched_words_in_smem + 3) & ~(size_t)3) * 4 + b + 16;
}

}  // namespace pe
