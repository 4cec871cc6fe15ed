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
// Paths per group (all exact):
//  * fast mode (k == 1, scan result available): the batched scan kernel already
//    evaluated this task's descriptor against the node table as it stood when the
//    batch began and left its two best rank classes (bitmaps + member lists).
//    Within a batch node state only gets worse, so the first class member not yet
//    touched in this batch is the sequentially-correct argmin.  Warp-specialised:
//    producer warps stage each task's candidates (TMA copy of the list window at
//    the row's cursor, walk against the touched bitmap), warp 0 consumes tasks in
//    order, eight per iteration, a committer warp applies the reservations.
//  * best class consumed: re-rank its members at their live state against the
//    first untouched member of the second class -- by the ordered warp itself
//    (inline_medium) or, for tasks with generic resources / host ports /
//    recent-failure counts, by the whole block.
//  * generic path (any k; also the fall-back): full table evaluation against the
//    live state; k == 1: block arg-min; k > 1: radix-select of the k smallest rank
//    keys, bitonic sort, sequential fill on staged rows, parallel write-back.
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

#ifndef PE_SEQ_PROFILE
#define PE_SEQ_PROFILE 0        // 1: the consumer warp also tallies wait / work cycles (diagnostic builds)
#endif
#define PE_SEQ_THREADS 512     // 128 registers per thread: the ordered fast path must not spill
#define PE_SEQ_LISTED 1024      // members of a class list this kernel looks at (the lists hold PE_LIST_CAP; beyond -> class bitmap)
#define PE_SEQ_KS 2048          // candidates staged in shared memory
#define PE_SEQ_RING 64          // fast-mode ring slots (copies are started this far ahead of the ordered warp)
#define PE_SEQ_NPW 11           // producer warps: 1-3, 5-7, 9-11, 13-14.  Warp 15 commits; warps 4, 8, 12 sit fast mode out so
                                // that the ordered warp (warp 0) has its SM sub-partition's issue slots to itself
#define PE_SEQ_WIN 256          // words staged per task: list entries, or bitmap words (8k nodes)
#define PE_SEQ_DEPTH 3          // copies a producer warp keeps in flight
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
__device__ __forceinline__ void st_release_smem(uint32_t *p, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(seq_smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_smem(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(seq_smem_u32(p)) : "memory");
    return v;
}
// Same, but gives up after ~1 s and raises a device error instead of hanging the GPU.
__device__ __forceinline__ bool sq_mbar_wait_wd(unsigned long long *bar, uint32_t parity, DevCounters *ctr, uint32_t code) {
    if (sq_mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!sq_mbar_try_wait(bar, parity)) {
        __nanosleep(20);     // a spinning warp would take issue slots from the warps that do the work
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
    // continuation of a batch the chunked parallel placement step (kernel_place.cuh) began: start at task *resume of the
    // batch (nothing to do if that is the batch's end) with the touched bitmap it left in touched_g
    const uint32_t *resume;
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

// Staged descriptor of one k=1 task: two 16-byte records the consumer reads with two loads.
struct FastTask {
    // a
    uint32_t n_cand;     // candidates the producer left in S.cands[slot] (list mode)
    uint32_t last;       // last staged member (where the class bitmap takes over)
    uint32_t n_list;     // staged members (list mode); 0 in bitmap mode
    uint32_t task_off;
    // b
    uint32_t n_class;    // members of the best class
    uint32_t row;        // scan row
    uint32_t flags;      // PE_FT_*
    uint32_t tie_start;
};
#define PE_FT_VALID 1u    // the scan found a feasible node and the group has exactly one task
#define PE_FT_SIMPLE 2u   // reservation = counters (+ cpu / mem): logged by the ordered warp, applied by the committer warp
#define PE_FT_COUNTS 4u   // DesiredState <= COMPLETED
#define PE_FT_PLAIN (PE_FT_VALID | PE_FT_SIMPLE | PE_FT_COUNTS)
#define PE_FT_INLINE 8u   // only cpu / memory / max-replicas can change feasibility: a consumed best class is resolved by the ordered warp itself
#define PE_SEQ_NCAND 32   // candidates per task = how far ahead of the ordered warp a list is walked (see fast_consumer)
#define PE_SEQ_GROUP 8    // tasks the ordered warp resolves per iteration
#define PE_SEQ_LOGN 512   // placements the ordered warp may run ahead of the committer warp
#define PE_SEQ_ROWCUR 6656 // scan rows that get list cursors (the ring and the cursors alias the k > 1 staging area)
#define PE_SEQ_STAGING_BYTES (PE_SEQ_KS * 45)   // CandKey 16 + cpu 8 + mem 8 + svc 4 + tot 4 + placed 4 + flags 1 per staged candidate
#define PE_SEQ_FAST_BYTES (PE_SEQ_RING * PE_SEQ_WIN * 4 + PE_SEQ_ROWCUR * 8)
#define PE_SEQ_REGION0 (((PE_SEQ_STAGING_BYTES > PE_SEQ_FAST_BYTES ? PE_SEQ_STAGING_BYTES : PE_SEQ_FAST_BYTES) + 15) & ~15)

struct SeqShared {
    pe_group G;
    GroupCtx C;
    __align__(16) FastTask ft[PE_SEQ_RING];
    uint32_t cands[PE_SEQ_RING][PE_SEQ_NCAND];
    uint32_t quick[PE_SEQ_RING];  // candidates of a task the ordered warp may resolve in a group (plain reservation, list mode), else PE_NONE
    uint32_t log[PE_SEQ_LOGN];   // node chosen for fast task q of the session, at q % PE_SEQ_LOGN
    uint32_t pub, applied;       // tasks placed by the ordered warp / reservations applied by the committer warp
    uint32_t tma_ph[PE_SEQ_RING];   // phase of each slot's copy barrier (one producer owns a slot at a time)
    unsigned long long tma_bar[PE_SEQ_RING];
    uint32_t ready[PE_SEQ_RING];  // session index + 1 of the task staged in the slot (release-stored by its producer)
    uint32_t stop, resume, stop_reason, bars_live, consumed;
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


struct SeqDebug { unsigned long long n_fast, n_placed, n_medium, iters, stops[5]; long long cyc_wait, cyc_work; unsigned long long prof[16]; };
// prof (PE_SEQ_PROFILE builds): 0 single-task calls, 1 their cycles, 2 bitmap continuations, 3 their cycles,
// 4 inline-medium calls, 5 their cycles, 6 cycles waiting for the committer inside them, 7 best-class members re-ranked

// ---- fast mode -------------------------------------------------------------------------------
// Warp 0 consumes tasks strictly in order; warps 1..PE_SEQ_NPW stay ahead of it.  Producer warp p
// stages the tasks i == p-1 (mod NPW): the scan row's member list comes into the task's ring slot
// by TMA bulk copy, the producer walks it against the `touched` bitmap and leaves the first
// PE_SEQ_NCAND untouched members as the task's candidates.  Touched bits never clear, so every
// member before a candidate stays touched and the first candidate that is still untouched at the
// task's turn is exactly the first untouched member; at most PE_SEQ_RING - 1 placements happen in
// between, which almost never takes all the candidates (then the ordered warp walks the window
// itself).  The ordered warp's dependent chain per task is: one candidate per lane, its touched
// bit, a warp min-reduction.  PE_SEQ_GROUP consecutive tasks are resolved per iteration: their loads
// are issued together and the choices of the earlier ones are applied to the later ones in registers.

__device__ __forceinline__ void fast_commit(const SeqParams &P, uint32_t gq, uint32_t task_off, bool counts, uint32_t n, uint32_t *touched) {
    P.K.out_node[task_off] = n;
    add_task_global(P.T, P.K, P.K.groups[gq], n, counts, P.ctr);   // generic resources / host ports: in place
    touched[n >> 5] |= 1u << (n & 31u);
}

// One task, every case.  Returns 0 to go on, else the stop reason (+ 16 if the task itself was placed).
// All reservations of the session's tasks [0, upto) are in the global columns (the caller published `pub` >= upto).
__device__ __forceinline__ bool wait_applied(const SeqParams &P, SeqShared &S, uint32_t upto) {
    const long long t0 = clock64();
    while (ld_acquire_smem(&S.applied) < upto) {
        __nanosleep(20);
        if (clock64() - t0 > 2000000000LL) { atomicOr(&P.ctr->error, PE_DEV_ERR_WD_CONSUMER); return false; }
    }
    return true;
}

// The best class of a task is consumed (every member was taken earlier in the batch).  Every node outside
// the two recorded classes ranked strictly worse than the second class -- or was infeasible -- when the
// batch began; inside a batch ranks only grow and feasibility only shrinks (resources are only reserved,
// counts only rise).  So the arg-min is the first untouched member of the second class (untouched: still
// feasible, still at rank c1) or a member of the best class that is STILL feasible, at its LIVE rank.
// PE_NONE: not resolvable here (no second class / no untouched member of it / best class not fully listed).
__device__ __forceinline__ uint32_t inline_medium(const SeqParams &P, SeqShared &S, const uint32_t *touched, uint32_t *rowcur2, uint32_t row,
                                                  uint32_t gq, uint32_t i, uint32_t lane, SeqDebug &dbg) {
    const uint32_t N = P.T.n_nodes, nwords = (N + 31u) >> 5;
    const ScanResult *sr = &P.scan[row];
    const uint32_t *L1 = P.L + (size_t)row * 2u * PE_LIST_CAP, *L2 = L1 + PE_LIST_CAP;
    // one round of independent global loads: the row record, the head of the best class's list and 128
    // members of the second class starting at the row's cursor (members before it are touched for good)
    const uint32_t cur2 = row < (uint32_t)PE_SEQ_ROWCUR ? (rowcur2[row] & ~31u) : 0u;
    uint32_t v2[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t idx = cur2 + (uint32_t)k * 32u + lane; v2[k] = idx < (uint32_t)PE_SEQ_LISTED ? L2[idx] : 0u; }
    const unsigned long long c1 = sr->c1;
    const uint4 meta = *reinterpret_cast<const uint4 *>(&sr->n0);   // n0, n1, tie_start, flags
    const uint32_t *svccol = sr->svccol;
    const long long cpu_res = sr->cpu_res, mem_res = sr->mem_res;
    const unsigned long long max_replicas = sr->max_replicas;   // (the row record is L2-hot; the 112-byte group descriptors are not)
    (void)gq;
    const uint32_t head1 = L1[lane];                                 // (rows are PE_LIST_CAP long: in bounds)
    const uint32_t n0 = meta.x, n1 = meta.y;
    if (c1 == PE_PREF_NONE || n0 > (uint32_t)PE_SEQ_LISTED) return PE_NONE;
    const uint32_t nl = min(n1, (uint32_t)PE_SEQ_LISTED);
    uint32_t n2 = PE_NONE;
    for (uint32_t j = cur2; j < nl && n2 == PE_NONE; j += 128u) {
        if (j != cur2) {
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint32_t idx = j + (uint32_t)k * 32u + lane; v2[k] = idx < nl ? L2[idx] : 0u; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t idx = j + (uint32_t)k * 32u + lane;
            const bool u = idx < nl && !((touched[v2[k] >> 5] >> (v2[k] & 31u)) & 1u);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, u);
            if (b && n2 == PE_NONE) {
                const int src = __ffs((int)b) - 1;
                n2 = __shfl_sync(0xFFFFFFFFu, v2[k], src);
                if (lane == 0 && row < (uint32_t)PE_SEQ_ROWCUR) rowcur2[row] = j + (uint32_t)k * 32u + (uint32_t)src;
            }
        }
    }
    if (n2 == PE_NONE && n1 > nl) {
        const uint32_t last = L2[nl - 1u];
        const uint32_t *Erow = P.E + ((size_t)row * 2u + 1u) * P.e_stride;
        for (uint32_t wb = last >> 5; wb < nwords && n2 == PE_NONE; wb += 32u) {
            const uint32_t w = wb + lane;
            uint32_t v = 0;
            if (w < nwords) {
                v = Erow[w] & ~touched[w];
                if (w == (last >> 5)) v &= (last & 31u) == 31u ? 0u : (0xFFFFFFFFu << ((last & 31u) + 1u));
                if (w == nwords - 1 && (N & 31u)) v &= (1u << (N & 31u)) - 1u;
            }
            const uint32_t bb = __ballot_sync(0xFFFFFFFFu, v != 0u);
            if (bb) {
                const int src = __ffs((int)bb) - 1;
                const uint32_t vv = __shfl_sync(0xFFFFFFFFu, v, src);
                n2 = (wb + (uint32_t)src) * 32u + (uint32_t)__ffs((int)vv) - 1u;
            }
        }
    }
    if (n2 == PE_NONE) return PE_NONE;
#if PE_SEQ_PROFILE
    const long long tw0 = clock64();
#endif
    if (!wait_applied(P, S, i)) return PE_NONE;
#if PE_SEQ_PROFILE
    dbg.prof[6] += (unsigned long long)(clock64() - tw0);
    dbg.prof[7] += n0;
#endif
    unsigned long long bp = c1;
    uint32_t bn = n2;
    for (uint32_t j = lane; j < n0; j += 32u) {
        const uint32_t n = j < 32u ? head1 : L1[j];   // (j < 32 only in the first round, where j == lane)
        const uint32_t sv = __ldcg(svccol + n), tot = __ldcg(P.T.total + n);    // live: the committer's reductions land in L2
        bool ok = true;
        if (meta.w & PE_SR_RES)               // ResourceFilter.Check on the live amounts, filter.go:76-84
            ok = cpu_res <= __ldcg(reinterpret_cast<const long long *>(P.T.cpu + n)) && mem_res <= __ldcg(reinterpret_cast<const long long *>(P.T.mem + n));
        if (meta.w & PE_SR_MAXREP) ok = ok && (unsigned long long)sv < max_replicas;   // filter.go:379-381
        const unsigned long long pref = make_pref(0u, sv, tot);
        if (ok && (pref < bp || (pref == bp && n < bn))) { bp = pref; bn = n; }
    }
    const uint32_t hi = (uint32_t)(bp >> 32), lo = (uint32_t)bp;
    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
    const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
    return __reduce_min_sync(0xFFFFFFFFu, (hi == mh && lo == ml) ? bn : 0xFFFFFFFFu);
}

__device__ __forceinline__ uint32_t consume_one(const SeqParams &P, SeqShared &S, uint32_t *touched, const uint32_t *ring, uint32_t *rowcur2,
                                                uint32_t gq, uint32_t slot, uint32_t lane, uint32_t i, SeqDebug &dbg) {
    const uint32_t N = P.T.n_nodes, nwords = (N + 31u) >> 5;
#if PE_SEQ_PROFILE
    const long long tq0 = clock64();
    dbg.prof[0]++;
#endif
    const uint4 fa = *reinterpret_cast<const uint4 *>(&S.ft[slot].n_cand);
    const uint4 fb = *reinterpret_cast<const uint4 *>(&S.ft[slot].n_class);
    const uint32_t n_cand = fa.x, last = fa.y, n_list = fa.z, task_off = fa.w;
    const uint32_t n_class = fb.x, row = fb.y, flags = fb.z, tie_start = fb.w;
    if (!(flags & PE_FT_VALID)) return 1;                  // nothing feasible when the batch began, or k != 1
    uint32_t n = PE_NONE;
    if (n_list) {
        // ---- list mode (canonical tie order)
        const uint32_t c = lane < n_cand ? S.cands[slot][lane] : 0u;
        const bool u = lane < n_cand && !((touched[c >> 5] >> (c & 31u)) & 1u);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, u);
        if (b) n = __shfl_sync(0xFFFFFFFFu, c, __ffs((int)b) - 1);
        if (n == PE_NONE && n_cand >= PE_SEQ_NCAND) {
            // every candidate went to the tasks in between (the list was walked up to PE_SEQ_RING tasks ago): walk the
            // staged window again; members before it are touched, so what it finds is the first untouched member
            const uint32_t *lst = ring + slot * PE_SEQ_WIN;
            for (uint32_t j = 0; j < tie_start && n == PE_NONE; j += 32u) {           // (list mode keeps the staged count in this field)
                const uint32_t idx = j + lane;
                const uint32_t v = idx < tie_start ? lst[idx] : 0u;
                const bool uu = idx < tie_start && !((touched[v >> 5] >> (v & 31u)) & 1u);
                const uint32_t bb = __ballot_sync(0xFFFFFFFFu, uu);
                if (bb) n = __shfl_sync(0xFFFFFFFFu, v, __ffs((int)bb) - 1);
            }
        }
        if (n == PE_NONE && n_class > n_list) {
            // every staged member is taken and the class goes on: continue on its bitmap (L2), 32 words a round
#if PE_SEQ_PROFILE
            const long long tb0 = clock64();
            dbg.prof[2]++;
#endif
            const uint32_t *Erow = P.E + (size_t)row * 2u * P.e_stride;
            for (uint32_t wb = last >> 5; wb < nwords && n == PE_NONE; wb += 32u) {
                const uint32_t w = wb + lane;
                uint32_t v = 0;
                if (w < nwords) {
                    v = Erow[w] & ~touched[w];
                    if (w == (last >> 5)) v &= (last & 31u) == 31u ? 0u : (0xFFFFFFFFu << ((last & 31u) + 1u));
                    if (w == nwords - 1 && (N & 31u)) v &= (1u << (N & 31u)) - 1u;
                }
                const uint32_t bb = __ballot_sync(0xFFFFFFFFu, v != 0u);
                if (bb) {
                    const int src = __ffs((int)bb) - 1;
                    const uint32_t vv = __shfl_sync(0xFFFFFFFFu, v, src);
                    n = (wb + (uint32_t)src) * 32u + (uint32_t)__ffs((int)vv) - 1u;
                }
            }
#if PE_SEQ_PROFILE
            dbg.prof[3] += (unsigned long long)(clock64() - tb0);
#endif
        }
        if (n == PE_NONE && (flags & PE_FT_INLINE) && (flags & PE_FT_PLAIN) == PE_FT_PLAIN) {
#if PE_SEQ_PROFILE
            const long long tm0 = clock64();
            dbg.prof[4]++;
#endif
            n = inline_medium(P, S, touched, rowcur2, row, gq, i, lane, dbg);
            if (n != PE_NONE) dbg.n_medium++;
#if PE_SEQ_PROFILE
            dbg.prof[5] += (unsigned long long)(clock64() - tm0);
#endif
        }
        if (n == PE_NONE) return 4;                        // the whole class is consumed
    } else {
        // ---- bitmap mode (rotated tie order): the staged window of the class bitmap
        const uint32_t lo_bit = tie_start;
        const uint32_t lo_word = lo_bit >> 5;
        const uint32_t ws = lo_word & ~3u;                 // window start, multiple of 4 words
        const uint32_t nwin = min((uint32_t)PE_SEQ_WIN, nwords - ws);
        const uint32_t *buf = ring + slot * PE_SEQ_WIN;
        for (uint32_t j = 0; j < nwin && n == PE_NONE; j += 32u) {
            const uint32_t w = ws + j + lane;
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
                n = (ws + j + (uint32_t)src) * 32u + (uint32_t)__ffs((int)vv) - 1u;
            }
        }
        if (n == PE_NONE) return 2;                        // not inside the staged window (may wrap): block-wide walk
    }
    const bool counts = (flags & PE_FT_COUNTS) != 0u;
    if (lane == 0) {
        // out_fail rows were zeroed when the tick started.  For the common reservation (no generic
        // resources / host ports) only the choice is logged here; the committer warp writes out_node
        // and applies the column updates (four reductions per task) -- global traffic in this loop
        // would serialise on its latency.
        if (flags & PE_FT_SIMPLE) { atomicOr(&touched[n >> 5], 1u << (n & 31u)); S.log[i % PE_SEQ_LOGN] = n; }
    }
    if (!(flags & PE_FT_SIMPLE)) {
        // generic resources / host ports are read-modify-write in place: the earlier reservations must have landed
        if (!wait_applied(P, S, i)) return 1;
        if (lane == 0) fast_commit(P, gq, task_off, counts, n, touched);
    }
    __syncwarp();   // (the slot is handed back when the caller publishes `pub`)
#if PE_SEQ_PROFILE
    dbg.prof[1] += (unsigned long long)(clock64() - tq0);
#endif
    return counts ? 0u : 16u + 3u;                         // rank did not move: later class bitmaps may hide this node
}

// The ordered warp's accesses to the touched bitmap, spelled out: an unconditional load (the compiler would
// otherwise wrap each one in a divergent branch and serialise the eight loads of a group) and a single
// reduction without the warp-aggregation code nvcc puts around atomicOr.
template <bool TS>
__device__ __forceinline__ uint32_t ld_touched(const uint32_t *touched, uint32_t word) {
    uint32_t v;
    if (TS) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(seq_smem_u32(touched + word)));
    else asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(touched + word));
    return v;
}
template <bool TS>
__device__ __forceinline__ void mark_touched(uint32_t *touched, uint32_t n) {
    if (TS) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(seq_smem_u32(touched + (n >> 5))), "r"(1u << (n & 31u)) : "memory");
    else asm volatile("red.global.or.b32 [%0], %1;" ::"l"(touched + (n >> 5)), "r"(1u << (n & 31u)) : "memory");
}

// TS: the touched bitmap lives in shared memory.
template <bool TS>
__device__ __forceinline__ void fast_consumer(const SeqParams &P, SeqShared &S, uint32_t *touched_s, const uint32_t *ring, uint32_t *rowcur2,
                                              uint32_t start, SeqDebug &dbg) {
    uint32_t *touched = TS ? touched_s : P.touched_g;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t total = P.g_end - start;
    uint32_t i = 0, reason = 0, applied_seen = 0;
    while (i < total) {
        if (i + (uint32_t)PE_SEQ_GROUP - applied_seen > (uint32_t)PE_SEQ_LOGN) {     // the placement log is a ring: never lap the committer
            applied_seen = ld_acquire_smem(&S.applied);
            if (i + (uint32_t)PE_SEQ_GROUP - applied_seen > (uint32_t)PE_SEQ_LOGN) {
                if (!wait_applied(P, S, i + (uint32_t)PE_SEQ_GROUP - (uint32_t)PE_SEQ_LOGN)) { reason = 1; break; }
                applied_seen = ld_acquire_smem(&S.applied);
            }
        }
        // ---- PE_SEQ_GROUP plain list-mode tasks at once.  Groups are aligned, so the eight slots are
        // consecutive, share the barrier phase, and every shared-memory address is base + constant.
        if ((i & (uint32_t)(PE_SEQ_GROUP - 1)) == 0u && i + (uint32_t)PE_SEQ_GROUP <= total) {
#if PE_SEQ_PROFILE
            const long long tp0 = clock64();
#endif
            const uint32_t slot0 = i % PE_SEQ_RING, par = (i / PE_SEQ_RING) & 1u;
            // eight plain loads of the slots' ready words (an mbarrier wait costs ~150 cycles each), then one acquire fence
            const volatile uint32_t *rd = &S.ready[slot0];
            bool ready = true;
            {
                uint32_t rv[PE_SEQ_GROUP];
#pragma unroll
                for (int q = 0; q < PE_SEQ_GROUP; q++) rv[q] = rd[q];
#pragma unroll
                for (int q = 0; q < PE_SEQ_GROUP; q++) ready = ready && rv[q] == i + (uint32_t)q + 1u;
            }
            if (!ready) {
                const long long t0 = clock64();
#pragma unroll 1
                for (int q = 0; q < PE_SEQ_GROUP; q++) {
                    while (rd[q] != i + (uint32_t)q + 1u) {
                        __nanosleep(20);
                        if (clock64() - t0 > 2000000000LL) { atomicOr(&P.ctr->error, PE_DEV_ERR_WD_CONSUMER); reason = 1; break; }
                    }
                    if (reason) break;
                }
                if (reason) break;
            }
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            (void)par;
#if PE_SEQ_PROFILE
            const long long tp1 = clock64();
            dbg.cyc_wait += tp1 - tp0;
#endif
            // k[q]: this lane's candidate for task q if it is still untouched, else PE_NONE.  Lists are in node
            // order, so the first untouched candidate is the smallest key: one warp reduction per task.
            uint32_t nc[PE_SEQ_GROUP], k[PE_SEQ_GROUP];
            const uint32_t *qk = &S.quick[slot0];
            const uint32_t *cd = &S.cands[slot0][lane];
#pragma unroll
            for (int q = 0; q < PE_SEQ_GROUP; q++) { nc[q] = qk[q]; k[q] = cd[q * PE_SEQ_NCAND]; }
#pragma unroll
            for (int q = 0; q < PE_SEQ_GROUP; q++) {
                const bool in = nc[q] != PE_NONE && lane < nc[q];      // (slots hold stale words past the candidate count)
                const uint32_t c = in ? k[q] : 0u;
                const uint32_t tw = ld_touched<TS>(touched, c >> 5);
                k[q] = (in && !((tw >> (c & 31u)) & 1u)) ? c : PE_NONE;
            }
#if PE_SEQ_PROFILE
            { uint32_t any = 0;
#pragma unroll
              for (int q = 0; q < PE_SEQ_GROUP; q++) any |= k[q];
              if (__any_sync(0xFFFFFFFFu, any == 77u)) dbg.prof[7]++; }
            const long long tp2 = clock64();
            dbg.prof[8] += (unsigned long long)(tp2 - tp1);       // loads + touched bits
#endif
            uint32_t done = 0, mine = 0;
#pragma unroll
            for (int q = 0; q < PE_SEQ_GROUP; q++) {
                if (done != (uint32_t)q) continue;
                const uint32_t n = __reduce_min_sync(0xFFFFFFFFu, k[q]);
                if (n == PE_NONE) continue;                                      // the general routine takes it
#pragma unroll
                for (int r = q + 1; r < PE_SEQ_GROUP; r++) k[r] = (k[r] == n) ? PE_NONE : k[r];   // what this choice takes from the later ones
                mine = lane == (uint32_t)q ? n : mine;
                done = (uint32_t)q + 1u;
            }
            if (lane < done) {
                mark_touched<TS>(touched, mine);
                S.log[(i % PE_SEQ_LOGN) + lane] = mine;                          // out_node + reservation: committer warp
            }
#if PE_SEQ_PROFILE
            const long long tp3 = clock64();
            dbg.prof[9] += (unsigned long long)(tp3 - tp2);       // resolving the eight tasks
#endif
            __syncwarp();
            i += done;
            if (lane == 0 && done) st_release_smem(&S.pub, i);                   // committer may apply them, producers may refill the slots
#if PE_SEQ_PROFILE
            dbg.prof[10] += (unsigned long long)(clock64() - tp3);  // publish
            dbg.prof[11]++;
            dbg.cyc_work += clock64() - tp1;
            if (done != (uint32_t)PE_SEQ_GROUP) dbg.iters++;
#endif
            if (done == (uint32_t)PE_SEQ_GROUP) continue;
            if (i >= total) break;
        }
        // ---- one task, every case
        const uint32_t slot = i % PE_SEQ_RING;
        {
            const long long t0 = clock64();
            while (ld_acquire_smem(&S.ready[slot]) != i + 1u) {
                __nanosleep(20);
                if (clock64() - t0 > 2000000000LL) { atomicOr(&P.ctr->error, PE_DEV_ERR_WD_CONSUMER); reason = 1; break; }
            }
            if (reason) break;
        }
        const uint32_t r = consume_one(P, S, touched, ring, rowcur2, start + i, slot, lane, i, dbg);
        if (r == 0u || r >= 16u) {
            i++;
            if (lane == 0) st_release_smem(&S.pub, i);
        }
        if (r == 0u) continue;
        reason = r >= 16u ? r - 16u : r;
        break;
    }
    dbg.n_fast += i; dbg.n_placed += i;                        // tasks [start, start+i) were placed here (n_medium of them by inline_medium)
    dbg.stops[reason < 5 ? reason : 0]++;
    if (lane == 0) {
        S.resume = start + i;
        S.consumed = i;
        S.stop_reason = reason;
        if (reason == 3) S.neutral = 1;
        st_release_smem(&S.pub, i);
        st_release_smem(&S.stop, 1u);
    }
}

// Committer warp: applies the reservations of the tasks the ordered warp placed (NodeInfo.addTask,
// nodeinfo.go:125-153, as reductions), 32 tasks at a time, and publishes how far it got.
__device__ __forceinline__ void fast_committer(const SeqParams &P, SeqShared &S, uint32_t start) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t done = 0;
    for (;;) {
        const uint32_t stop = ld_acquire_smem(&S.stop);
        const uint32_t pub = ld_acquire_smem(&S.pub);
        if (pub == done) {
            if (stop) break;
            __nanosleep(20);
            continue;
        }
        const uint32_t q = done + lane;
        if (q < pub) {
            const uint32_t gq = start + q;
            const ScanResult *sr = &P.scan[P.task_row[gq - P.g_begin]];
            const uint32_t fl = sr->flags;
            if (fl & PE_SR_SIMPLE) {               // others were applied in place by the ordered warp
                const uint32_t n = S.log[q % PE_SEQ_LOGN];
                P.K.out_node[P.K.groups[gq].task_off] = n;
                const long long cpu_res = sr->cpu_res, mem_res = sr->mem_res;
                if (cpu_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.cpu[n]), (unsigned long long)(-cpu_res));
                if (mem_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.mem[n]), (unsigned long long)(-mem_res));
                if (fl & PE_SR_COUNTS) {
                    atomicAdd(&P.T.total[n], 1u);
                    if (atomicAdd(&sr->svccol[n], 1u) + 1u >= 0xFFFFFFu) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
                }
            }
        }
        __threadfence();
        __syncwarp();
        done = min(pub, done + 32u);
        if (lane == 0) st_release_smem(&S.applied, done);
    }
}

// Producer side of one task, split in two so that a warp keeps two tasks in flight: `stage_issue`
// (lane 0) waits for the slot, writes the descriptor and starts the copy; `stage_finish` (whole
// warp) waits for the copy, picks the candidates and hands the slot to the consumer.
struct ProdDesc { uint32_t row, off; unsigned long long c0; uint4 meta; };   // lane 0's prefetched view of a task (meta = n0, n1, tie_start, flags)
struct ProdTask { uint32_t go, n_list, base, n_listed, row, copy; };   // n_list members staged from list offset base; n_listed in the row's list

__device__ __forceinline__ ProdTask stage_issue(const SeqParams &P, SeqShared &S, uint32_t *ring, uint32_t *rowcur, uint32_t start, uint32_t i,
                                                uint32_t lane, const ProdDesc &d) {
    const uint32_t row = d.row, task_off = d.off;
    ProdTask t; t.go = 1; t.n_list = 0; t.base = 0; t.n_listed = 0; t.row = row; t.copy = 0;
    const uint32_t slot = i % PE_SEQ_RING;
    if (lane == 0) {
        const uint32_t nwords = (P.T.n_nodes + 31u) >> 5;
        volatile uint32_t *vstop = &S.stop;
#if PE_SEQ_PROFILE
        const long long ta = clock64();
#endif
        const unsigned long long c0 = d.c0;
        const uint4 meta = d.meta;
#if PE_SEQ_PROFILE
        const long long tb = clock64() + (long long)(meta.x & 0u) + (long long)(c0 & 0ull);   // (depends on the loads)
#endif
        while (ld_acquire_smem(&S.pub) + (uint32_t)PE_SEQ_RING <= i) {      // the slot's previous task (i - RING) is not consumed yet
            if (*vstop) { t.go = 0; break; }
            __nanosleep(32);
        }
#if PE_SEQ_PROFILE
        const long long tc = clock64();
        if (threadIdx.x == 32) { atomicAdd(&P.ctr->prof[12], (unsigned long long)(tb - ta)); atomicAdd(&P.ctr->prof[13], (unsigned long long)(tc - tb)); }
#endif
        if (*vstop) t.go = 0;
        if (t.go) {
            const bool valid = (meta.w & PE_SR_K1) && c0 != PE_PREF_NONE;
            FastTask &f = S.ft[slot];
            // list mode: stage the window of the member list that starts at the row's cursor (every member
            // before the cursor is known to be touched: tasks that share the row, and windows already walked)
            if (meta.z == 0u && valid) {
                t.n_listed = min(meta.x, (uint32_t)PE_SEQ_LISTED);
                t.base = row < (uint32_t)PE_SEQ_ROWCUR ? (min(rowcur[row], t.n_listed - 1u) & ~3u) : 0u;
                t.n_list = min(t.n_listed - t.base, (uint32_t)PE_SEQ_WIN);
            }
            f.n_cand = 0; f.last = 0;
            S.quick[slot] = PE_NONE;
            f.n_list = t.n_list;
            f.task_off = task_off;
            f.n_class = meta.x;
            f.row = row;
            f.flags = (valid ? PE_FT_VALID : 0u) | ((meta.w & PE_SR_SIMPLE) ? PE_FT_SIMPLE : 0u) | ((meta.w & PE_SR_COUNTS) ? PE_FT_COUNTS : 0u) |
                      ((meta.w & PE_SR_INLINE) ? PE_FT_INLINE : 0u);
            f.tie_start = meta.z;
            bool has_copy = false;
            if (t.n_list) {
                const uint32_t bytes = ((t.n_list + 3u) & ~3u) * 4u;
                sq_mbar_expect_tx(&S.tma_bar[slot], bytes);
                sq_tma_bulk_g2s(ring + slot * PE_SEQ_WIN, P.L + (size_t)row * 2u * PE_LIST_CAP + t.base, bytes, &S.tma_bar[slot]);
                has_copy = true;
            } else if (valid) {
                const uint32_t ws = (meta.z >> 5) & ~3u;                       // 16-byte aligned window start
                if (ws < nwords) {
                    // rows are padded to e_stride (a multiple of 32 words), so the copy may run past nwords
                    const uint32_t nw = min((uint32_t)PE_SEQ_WIN, P.e_stride - ws);
                    sq_mbar_expect_tx(&S.tma_bar[slot], nw * 4u);
                    sq_tma_bulk_g2s(ring + slot * PE_SEQ_WIN, P.E + (size_t)row * 2u * P.e_stride + ws, nw * 4u, &S.tma_bar[slot]);
                    has_copy = true;
                }
            }
            t.copy = has_copy ? 1u : 0u;
        }
    }
    // lane 0 -> warp: (go, copy, n_list <= 256, base < 1024, n_listed <= 1024) in one word, the row in another
    uint32_t pk = t.go | (t.copy << 1) | (t.n_list << 2) | (t.base << 11) | (t.n_listed << 21);
    pk = __shfl_sync(0xFFFFFFFFu, pk, 0);
    t.row = __shfl_sync(0xFFFFFFFFu, t.row, 0);
    t.go = pk & 1u; t.copy = (pk >> 1) & 1u; t.n_list = (pk >> 2) & 0x1FFu; t.base = (pk >> 11) & 0x3FFu; t.n_listed = pk >> 21;
    return t;
}

__device__ __forceinline__ bool stage_finish(const SeqParams &P, SeqShared &S, uint32_t *ring, uint32_t *rowcur, const uint32_t *touched,
                                             uint32_t i, uint32_t lane, ProdTask t) {
    const uint32_t slot = i % PE_SEQ_RING;
    uint32_t *buf = ring + slot * PE_SEQ_WIN;
    auto wait_copy = [&]() -> bool {
        const uint32_t ph = S.tma_ph[slot];
        if (!sq_mbar_wait_wd(&S.tma_bar[slot], ph, P.ctr, PE_DEV_ERR_WD_DRAIN)) return false;
        __syncwarp();
        if (lane == 0) S.tma_ph[slot] = ph ^ 1u;
        __syncwarp();
        return true;
    };
#if PE_SEQ_PROFILE
    const long long td = clock64();
#endif
    if (t.copy && !wait_copy()) return false;
#if PE_SEQ_PROFILE
    const long long te = clock64();
#endif
    if (t.n_list) {
        const uint32_t lane_lt = (1u << lane) - 1u;
        uint32_t found = 0, first_at = 0;
        for (;;) {
            // the first PE_SEQ_NCAND untouched members of the window, in list order (lane-major: 4 consecutive entries per lane)
            const uint4 *lst4 = reinterpret_cast<const uint4 *>(buf);
            for (uint32_t j = 0; j < t.n_list && found < PE_SEQ_NCAND; j += 128u) {
                const uint32_t il = j + lane * 4u;
                uint32_t m = 0;
                uint4 e = make_uint4(0, 0, 0, 0);
                if (il < t.n_list) {
                    e = lst4[il >> 2];
                    const uint32_t left = t.n_list - il;
                    const uint32_t tx = touched[e.x >> 5], ty = left > 1u ? touched[e.y >> 5] : ~0u;
                    const uint32_t tz = left > 2u ? touched[e.z >> 5] : ~0u, tw = left > 3u ? touched[e.w >> 5] : ~0u;
                    m = (((tx >> (e.x & 31u)) & 1u) ^ 1u) | ((((ty >> (e.y & 31u)) & 1u) ^ 1u) << 1) | ((((tz >> (e.z & 31u)) & 1u) ^ 1u) << 2) |
                        ((((tw >> (e.w & 31u)) & 1u) ^ 1u) << 3);
                }
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, m & 1u), b1 = __ballot_sync(0xFFFFFFFFu, m & 2u);
                const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, m & 4u), b3 = __ballot_sync(0xFFFFFFFFu, m & 8u);
                const uint32_t any = b0 | b1 | b2 | b3;
                if (found == 0u && any) first_at = j + (uint32_t)(__ffs((int)any) - 1) * 4u;   // (rounded down to the lane's 4 entries)
                uint32_t pos = found + __popc(b0 & lane_lt) + __popc(b1 & lane_lt) + __popc(b2 & lane_lt) + __popc(b3 & lane_lt);
                if ((m & 1u) && pos < PE_SEQ_NCAND) S.cands[slot][pos] = e.x;
                pos += m & 1u;
                if ((m & 2u) && pos < PE_SEQ_NCAND) S.cands[slot][pos] = e.y;
                pos += (m >> 1) & 1u;
                if ((m & 4u) && pos < PE_SEQ_NCAND) S.cands[slot][pos] = e.z;
                pos += (m >> 2) & 1u;
                if ((m & 8u) && pos < PE_SEQ_NCAND) S.cands[slot][pos] = e.w;
                found += __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            }
            // every member before this one is touched for good: later tasks of the row start here
            const uint32_t cur = t.base + (found ? first_at : t.n_list);
            if (lane == 0 && t.row < (uint32_t)PE_SEQ_ROWCUR) atomicMax(&rowcur[t.row], cur);
            if (found || t.base + t.n_list >= t.n_listed) break;
            // the whole window is taken and the list goes on: bring the next window into the same slot
            t.base += t.n_list;
            t.n_list = min(t.n_listed - t.base, (uint32_t)PE_SEQ_WIN);
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t bytes = ((t.n_list + 3u) & ~3u) * 4u;
                sq_mbar_expect_tx(&S.tma_bar[slot], bytes);
                sq_tma_bulk_g2s(buf, P.L + (size_t)t.row * 2u * PE_LIST_CAP + t.base, bytes, &S.tma_bar[slot]);
            }
            if (!wait_copy()) return false;
        }
        if (lane == 0) {
            const uint32_t fl = S.ft[slot].flags;
            S.quick[slot] = (fl & PE_FT_PLAIN) == PE_FT_PLAIN ? min(found, (uint32_t)PE_SEQ_NCAND) : PE_NONE;
            S.ft[slot].n_cand = min(found, (uint32_t)PE_SEQ_NCAND);
            S.ft[slot].n_list = t.base + t.n_list;     // list members known to the pipeline (all touched, or candidates)
            S.ft[slot].tie_start = t.n_list;           // list mode (tie_start == 0): the field carries the staged count
            S.ft[slot].last = buf[t.n_list - 1u];
        }
    }
    __syncwarp();
    if (lane == 0) st_release_smem(&S.ready[slot], i + 1u);   // release: descriptor, candidates and the copied words are visible
#if PE_SEQ_PROFILE
    if (threadIdx.x == 32) { atomicAdd(&P.ctr->prof[14], (unsigned long long)(te - td)); atomicAdd(&P.ctr->prof[15], (unsigned long long)(clock64() - te)); }
#endif
    return true;
}

// One converged lane per warp issues the TMA copies (the CUTLASS elect-one idiom).  A producer
// never leaves with a copy in flight: whatever it issued it also waits for.
__device__ __forceinline__ void fast_producer(const SeqParams &P, SeqShared &S, uint32_t *ring, uint32_t *rowcur, const uint32_t *touched,
                                              uint32_t start) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t pw = warp - 1u - (warp >> 2);     // 1,2,3,5,6,7,9,10,11,13,14 -> 0..10
    const uint32_t total = P.g_end - start;
    if (pw >= total) return;
    // lane 0's global loads run ahead of their use: the task -> row map two tasks ahead, the row record one ahead
    auto load_row = [&](uint32_t i, ProdDesc &d) {
        if (lane == 0 && i < total) { d.row = P.task_row[start + i - P.g_begin]; d.off = P.K.groups[start + i].task_off; }
    };
    auto load_scan = [&](uint32_t i, ProdDesc &d) {
        if (lane == 0 && i < total) { d.c0 = P.scan[d.row].c0; d.meta = *reinterpret_cast<const uint4 *>(&P.scan[d.row].n0); }
    };
    // software pipeline over this warp's tasks t_m = pw + m * NPW.  In iteration m: the row of t_{m+DEPTH+1} is
    // requested, then the row record of t_{m+DEPTH} (its row arrived an iteration ago), then the copy of
    // t_{m+DEPTH-1} is started (its record arrived an iteration ago), then t_m is walked and handed over --
    // no global load is waited for in the iteration that issued it.
    ProdDesc dI{}, dS{};
    ProdTask pt[PE_SEQ_DEPTH];
    {
        ProdDesc d[PE_SEQ_DEPTH + 1];
#pragma unroll
        for (int k = 0; k <= PE_SEQ_DEPTH; k++) { d[k] = ProdDesc{}; load_row(pw + (uint32_t)k * PE_SEQ_NPW, d[k]); }
#pragma unroll
        for (int k = 0; k < PE_SEQ_DEPTH; k++) load_scan(pw + (uint32_t)k * PE_SEQ_NPW, d[k]);
#pragma unroll
        for (int k = 0; k < PE_SEQ_DEPTH - 1; k++) {
            pt[k].go = 0; pt[k].n_list = 0; pt[k].base = 0; pt[k].n_listed = 0; pt[k].row = 0; pt[k].copy = 0;
            const uint32_t ti = pw + (uint32_t)k * PE_SEQ_NPW;
            if (ti < total && (k == 0 || pt[k - 1].go)) pt[k] = stage_issue(P, S, ring, rowcur, start, ti, lane, d[k]);
        }
        dI = d[PE_SEQ_DEPTH - 1];   // row + record
        dS = d[PE_SEQ_DEPTH];       // row only
    }
    for (uint32_t i = pw;; i += PE_SEQ_NPW) {
        if (!pt[0].go) return;                       // (stage_issue starts no copy when it returns go == 0)
        const uint32_t ahead = i + (uint32_t)(PE_SEQ_DEPTH - 1) * PE_SEQ_NPW;
        ProdDesc dR{};
        load_row(ahead + 2u * PE_SEQ_NPW, dR);
        load_scan(ahead + PE_SEQ_NPW, dS);
        ProdTask nx; nx.go = 0; nx.n_list = 0; nx.base = 0; nx.n_listed = 0; nx.row = 0; nx.copy = 0;
        if (ahead < total && pt[PE_SEQ_DEPTH - 2].go)
            nx = stage_issue(P, S, ring, rowcur, start, ahead, lane, dI);   // this copy flies while earlier lists are walked
        pt[PE_SEQ_DEPTH - 1] = nx;
        const bool ok = stage_finish(P, S, ring, rowcur, touched, i, lane, pt[0]);
        if (!ok) {
            // leaving: whatever was issued must land first
#pragma unroll
            for (int k = 1; k < PE_SEQ_DEPTH; k++) {
                const uint32_t slot = (i + (uint32_t)k * PE_SEQ_NPW) % PE_SEQ_RING;
                if (pt[k].go && pt[k].copy) sq_mbar_wait_wd(&S.tma_bar[slot], S.tma_ph[slot], P.ctr, PE_DEV_ERR_WD_DRAIN);
            }
            return;
        }
        if (i + PE_SEQ_NPW >= total) return;
#pragma unroll
        for (int k = 0; k < PE_SEQ_DEPTH - 1; k++) pt[k] = pt[k + 1];
        dI = dS; dS = dR;
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
    uint32_t *touched_s = reinterpret_cast<uint32_t *>(dyn_smem + PE_SEQ_REGION0);   // after the staging area / the ring + cursors aliasing it
    uint32_t *touched = P.touched_in_smem ? touched_s : P.touched_g;
    // fast mode (scan batches hold k == 1 groups only) re-uses the k > 1 staging area: ring slots + per-row list cursors
    uint32_t *ring = reinterpret_cast<uint32_t *>(dyn_smem);   // [PE_SEQ_RING][PE_SEQ_WIN]
    uint32_t *rowcur = ring + PE_SEQ_RING * PE_SEQ_WIN;        // [PE_SEQ_ROWCUR]
    uint32_t *rowcur2 = rowcur + PE_SEQ_ROWCUR;                // second-class cursors (inline_medium)
    uint32_t gi = P.g_begin;
    if (P.resume != nullptr) {
        const uint32_t r = *P.resume;
        if (r >= P.g_end - P.g_begin) return;        // the parallel step placed the whole batch
        gi += r;
    }
    if (P.scan != nullptr) for (uint32_t r = tid; r < 2u * PE_SEQ_ROWCUR; r += nth) rowcur[r] = 0;

    if (P.resume == nullptr) { for (uint32_t w = tid; w < P.touched_words; w += nth) touched[w] = 0; }
    else if (P.touched_in_smem) { for (uint32_t w = tid; w < P.touched_words; w += nth) touched_s[w] = P.touched_g[w]; }
    if (tid == 0) { S.neutral = 0; S.bars_live = 0; S.bestv[0] = S.bestv[1] = S.bestv[2] = PE_NONE; }
    uint32_t slot = 0;   // uniform across the block
    unsigned long long n_fast = 0, n_medium = 0, n_slow = 0, n_placed = 0, n_evalg = 0;
    long long cyc_fast = 0, cyc_medium = 0, cyc_generic = 0, t_mark = clock64();
    SeqDebug dbg{};   // thread 0's private tallies
    __syncthreads();

    while (gi < P.g_end) {
        // ================= fast mode: warp-specialised pipeline over k == 1 tasks ====
        // (see the comment above fast_consumer).  The mode ends at the first task the
        // ordered warp cannot resolve from the scan's classes; that task takes the
        // block-wide path below and fast mode starts again after it.
        if (P.scan != nullptr && !S.neutral) {
            __syncthreads();
            { const long long t1 = clock64(); cyc_generic += t1 - t_mark; t_mark = t1; }
            if (tid == 0) {
                for (int r = 0; r < PE_SEQ_RING; r++) {
                    if (S.bars_live) sq_mbar_inval(&S.tma_bar[r]);
                    S.ready[r] = 0;
                    sq_mbar_init(&S.tma_bar[r], 1);
                    S.tma_ph[r] = 0;
                }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                S.bars_live = 1;
                S.stop = 0; S.pub = 0; S.applied = 0;
                S.resume = P.g_end;
                S.stop_reason = 0;
            }
            __syncthreads();
            const uint32_t start = gi;
            if (warp == 0) {
                if (P.touched_in_smem) fast_consumer<true>(P, S, touched_s, ring, rowcur2, start, dbg);
                else fast_consumer<false>(P, S, touched_s, ring, rowcur2, start, dbg);
            }
            else if (warp == 15) fast_committer(P, S, start);
            else if ((warp & 3u) != 0u) fast_producer(P, S, ring, rowcur, touched, start);
            __syncthreads();   // no copy is in flight (every producer waits for the one it issued); every reservation is applied
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
                    const uint32_t nl = min(sr.n1, (uint32_t)PE_SEQ_LISTED);
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
                    if (use_lists && sr.n0 <= (uint32_t)PE_SEQ_LISTED) {
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
                if (!(meta & PE_NODE_VALID) || (G.leaf_cnt && !in_leaf(T, K, G, n))) continue;
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
            if (!(meta & PE_NODE_VALID) || (G.leaf_cnt && !in_leaf(T, K, G, n))) { P.ff8[n] = 0xFF; continue; }   // not in the node set / under another leaf
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
        n_fast += dbg.n_fast - dbg.n_medium; n_medium += dbg.n_medium; n_placed += dbg.n_placed;
        P.ctr->fast_path += n_fast;
        P.ctr->medium_path += n_medium;
        cyc_generic += clock64() - t_mark;   // whatever is left is the generic path (and loop overhead)
        P.ctr->cyc_fast += (unsigned long long)cyc_fast;
        P.ctr->cyc_medium += (unsigned long long)cyc_medium;
        P.ctr->cyc_generic += (unsigned long long)cyc_generic;
        P.ctr->cyc_cons_wait += (unsigned long long)dbg.cyc_wait; P.ctr->cyc_cons_work += (unsigned long long)dbg.cyc_work; P.ctr->iters += dbg.iters;
        for (int r = 0; r < 5; r++) P.ctr->stops[r] += dbg.stops[r];
        for (int r = 0; r < 12; r++) P.ctr->prof[r] += dbg.prof[r];   // 12..15: producer warp 1 (load wait, slot wait, copy wait, walk)
        P.ctr->slow_path += n_slow;
        P.ctr->placements += n_placed;
        P.ctr->evals_generic += n_evalg;
    }
}

static inline size_t seq_dyn_smem_bytes(uint32_t touched_words_in_smem) {
    return (size_t)PE_SEQ_REGION0 + (((size_t)touched_words_in_smem + 3) & ~(size_t)3) * 4 + 16;
}

}  // namespace pe
