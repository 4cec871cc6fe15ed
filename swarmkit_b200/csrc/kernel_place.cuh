// kernel_place.cuh -- the chunked PARALLEL placement step for batches of one-off
// tasks: what turns the scan's per-row classes into the sequentially exact
// placements of
//
//   Scheduler.tick one-off loop            manager/scheduler/scheduler.go:467-469
//   scheduleTaskGroup (k = 1)              manager/scheduler/scheduler.go:694-748
//   nodeSet.tree with a heap of one        manager/scheduler/nodeset.go:107-120
//   scheduleNTasksOnNodes (one task)       manager/scheduler/scheduler.go:844-924
//   NodeInfo.addTask (the reservation)     manager/scheduler/nodeinfo.go:108-154
//
// The reference places one task after the other; every reservation changes what
// the next decision reads.  Here a batch is cut into chunks of PE_PL_CHUNK tasks
// and every chunk runs three bulk-synchronous phases on a thread-block cluster
// (8 CTAs x 16 warps, distributed shared memory, cluster barriers):
//
//   stage   (parallel, one warp per task, all 8 CTAs): against the state at the
//           chunk's start, the task's candidates in the order the reference would
//           prefer them -- the untouched members of its row's best class, then
//           (when those run short) the touched members of that class at their
//           LIVE rank merged with the untouched members of the second class.
//           Task i of a chunk gets i + 1 candidates: the i tasks before it take
//           one node each, so the list cannot run out.  Written straight into
//           CTA 0's shared memory (DSMEM).
//   resolve (ordered, ONE warp of CTA 0): the chunk's tasks one after the other,
//           each taking its first candidate that no earlier task holds -- the 32
//           lanes test 32 candidates at once against the batch's touched bitmap in
//           shared memory, the loads of four tasks in flight together.  A task that
//           skipped a candidate ranked strictly better than its choice (the node was
//           taken inside the chunk: its rank moved) recomputes those ranks from the
//           chunk's log -- exactly.
//   commit  (parallel, CTA 0): NodeInfo.addTask for the chunk's placements as
//           reductions on the global columns, the batch's touched bitmap.
//
// Nothing here is approximate: a task the chunk logic cannot place exactly (no
// feasible node when the batch began, rotated tie order, a non-counting task,
// both recorded classes consumed, ...) ends the kernel at that task and the
// ordered sequencer (kernel_sequencer.cuh) takes the rest of the batch.
// oracle/place_model.cpp restates this algorithm on the CPU and
// tests/test_place_model_cpu.py pins it to the sequential oracle.
#pragma once
#include <cooperative_groups.h>

#include "kernel_scan.cuh"

namespace pe {
namespace cg = cooperative_groups;

#define PE_PL_CLUSTER 8                                        // CTAs per cluster: 8 (portable) or 16 (opt-in), chosen at launch
#define PE_PL_CLUSTER_MAX 16
#define PE_PL_WARPS 16
#define PE_PL_THREADS (PE_PL_WARPS * 32)
#define PE_PL_CHUNK (PE_PL_CLUSTER_MAX * PE_PL_WARPS)          // most tasks per chunk (a chunk = the warps in the cluster)
#define PE_PL_SCR 128                                          // words per scratch buffer of a staging warp (>= 2 * PE_PL_K)
#define PE_PL_K 64                                             // candidates per task (task i of a chunk asks for min(i + 1, K))
#define PE_PL_KS (PE_PL_K + 1)                                 // row stride of the candidate table in words: odd, so that the 32 lanes
                                                               // of the resolve warp reading candidate j of 32 consecutive rows hit 32 banks
#define PE_PL_HASH 2048                                        // the chunk's set of nodes given a task inside this chunk (at most PE_PL_CHUNK)
#define PE_PL_EMPTY 0xFFFFFFFFu
#define PE_PL_TOUCHED 0x80000000u    // candidate flag: the node already carried a task of this batch when the chunk began (a touched
                                     // member of the best class in a tail).  Such a node can only be "taken in this chunk" by a task
                                     // that JOINED it: the in-chunk hash set answers; for every other candidate the bitmap does.
#define PE_PL_NODE(x) ((x) & 0x7FFFFFFFu)
#define PE_PL_NIL 0xFFFFu

struct PlaceParams {
    DevTable T;
    TickDev K;
    uint32_t b0, B;               // the batch: groups [b0, b0 + B), all k == 1
    const ScanResult *scan;       // [row]
    const uint32_t *task_row;     // [B] task -> row
    const uint32_t *L;            // class member lists [row][2][PE_LIST_CAP]
    uint32_t list_cap;            // members of a list that are there (PE_LIST_CAP; fewer when the lists came over the wire)
    uint32_t *touched;            // [(N + 31) / 32] nodes given a task in this batch (zero when the batch begins)
    uint32_t *cursors;            // [rows][2] list positions before which every member is touched (zero when the batch begins)
    uint32_t *resume;             // out: first task of the batch (relative) NOT placed here; B = all of them
    uint32_t tk_words;            // words of the shared-memory copy of the touched bitmap kept by CTA 0 (0: too many nodes,
                                  // the resolve phase asks the in-chunk hash set instead)
    DevCounters *ctr;
};

// Candidate list of one task of the chunk.  Candidates are sorted by (rank key at the chunk's start, node); rank
// group 0 = untouched members of the best class (key = the row's c0), groups 1..3 start at rs1..rs3.
struct PlSlot {
    uint32_t kind;                // 0: candidates follow; 1: not placeable here (the ordered sequencer takes over)
    uint32_t n_cand;
    uint32_t rs1, rs2, rs3;       // first candidate of rank group 1 / 2 / 3 (= n_cand when the group is absent)
    uint32_t row;
    uint32_t task_off, flags;     // the group's task slot; the row's PE_SR_* flags
    unsigned long long c0, kv1, kv2, kv3;   // rank keys of groups 0..3
    // what the commit phase and the exact re-ranking need of the row, so that neither waits on global memory
    uint32_t *svccol;
    long long cpu_res, mem_res;
    unsigned long long max_replicas;
};

struct PlShared {
    PlSlot slot[PE_PL_CHUNK];
    uint32_t cand[PE_PL_CHUNK * PE_PL_KS];
    // what the resolve warp reads of every slot, as arrays (a 96-byte struct stride would put 32 lanes on 4 banks)
    uint16_t s_ncand[PE_PL_CHUNK], s_rs1[PE_PL_CHUNK], s_rs2[PE_PL_CHUNK], s_rs3[PE_PL_CHUNK];
    uint8_t s_kind[PE_PL_CHUNK];
    uint32_t hkey[PE_PL_HASH];                 // open addressing: nodes given a task inside this chunk
    uint32_t log_node[PE_PL_CHUNK];
    uint16_t log_task[PE_PL_CHUNK];            // task (chunk-relative = slot index)
    uint8_t log_tail[PE_PL_CHUNK];             // 1: the choice came from rank group >= 1
    uint32_t n_log, stop, cut, done;            // done: tasks of the chunk the resolve phase settled (the next chunk starts after them)
    uint32_t n_pub, res_done;                   // resolve -> commit hand-over: log entries published so far; the resolve warp is done
    uint32_t wd;                                // watchdog: a loop of this chunk ran away (bit per loop); the chunk is handed to the ordered sequencer untouched
    uint32_t scratch[PE_PL_WARPS][3][PE_PL_SCR];   // per warp: tail gather buffers, the candidate list under construction
};

__device__ __forceinline__ uint32_t pl_hash(uint32_t node) { return (node * 2654435761u) >> 21; }   // 11 bits


__device__ __forceinline__ unsigned long long pl_wmin64(unsigned long long v) {
    uint32_t h = (uint32_t)(v >> 32), l = (uint32_t)v;
    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, h);
    l = __reduce_min_sync(0xFFFFFFFFu, h == mh ? l : 0xFFFFFFFFu);
    return ((unsigned long long)mh << 32) | l;
}

// number of elements of the sorted array a[0, n) that are < x
__device__ __forceinline__ uint32_t pl_lower(const uint32_t *a, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1u; else hi = mid; }
    return lo;
}

// The first `want` untouched members of a class list from its cursor, in list (= node) order, written to out[0..).
// Returns how many were found (<= want); moves the cursor to the first untouched member seen (or the list's end).
__device__ __forceinline__ uint32_t pl_walk(const uint32_t *list, uint32_t n_listed, uint32_t *cursor, const uint32_t *touched,
                                            uint32_t want, uint32_t *out, uint32_t lane) {
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t found = 0, first_at = n_listed;
    bool have_first = false;
    for (uint32_t p = __ldcg(cursor); p < n_listed && found < want; p += 128u) {
        uint32_t m[4], tw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = p + (uint32_t)k * 32u + lane; m[k] = idx < n_listed ? list[idx] : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = p + (uint32_t)k * 32u + lane; tw[k] = idx < n_listed ? __ldcg(touched + (m[k] >> 5)) : 0xFFFFFFFFu; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool u = !((tw[k] >> (m[k] & 31u)) & 1u);
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, u);
            if (b) {
                if (!have_first) { have_first = true; first_at = p + (uint32_t)k * 32u + (uint32_t)__ffs((int)b) - 1u; }
                const uint32_t pos = found + __popc(b & lane_lt);
                if (u && pos < want) out[pos] = m[k];
                found += __popc(b);
            }
        }
    }
    if (lane == 0) atomicMax(cursor, first_at);
    return found < want ? found : want;
}

// ---- stage: candidates of task i of the chunk (one warp) ---------------------------------------------
struct PlRowView {                 // the row record, in registers
    unsigned long long c0, c1;
    uint32_t n0, n1, flags;
    long long cpu_res, mem_res;
    const uint32_t *svccol;
    unsigned long long max_replicas;
};

// live rank of a TOUCHED member of the best class (state as of the chunk's start); false = no longer feasible.
// Only cpu / memory / max-replicas can have changed for the rows that get here (PE_SR_INLINE).
__device__ __forceinline__ bool pl_live(const PlaceParams &P, const PlRowView &R, uint32_t m, unsigned long long &key) {
    const uint32_t sv = __ldcg(R.svccol + m), tot = __ldcg(P.T.total + m);
    bool ok = true;
    if (R.flags & PE_SR_RES)       // ResourceFilter.Check on the live amounts, filter.go:76-84
        ok = R.cpu_res <= __ldcg(reinterpret_cast<const long long *>(P.T.cpu + m)) &&
             R.mem_res <= __ldcg(reinterpret_cast<const long long *>(P.T.mem + m));
    if (R.flags & PE_SR_MAXREP) ok = ok && (unsigned long long)sv < R.max_replicas;   // filter.go:379-381
    key = make_pref(0u, sv, tot);
    return ok;
}

// The touched members of the best class whose live rank is exactly `kcur`, node order, at most `room` of them -> out.
// 128 members per round, their loads issued together.
__device__ __forceinline__ uint32_t pl_gather(const PlaceParams &P, const PlRowView &R, const uint32_t *L0, uint32_t nl0,
                                              unsigned long long kcur, uint32_t room, uint32_t *out, uint32_t lane) {
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t cnt = 0;
    for (uint32_t q = 0; q < nl0 && cnt < room; q += 128u) {
        uint32_t m[4], tw[4];
        bool f[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = q + (uint32_t)k * 32u + lane; m[k] = idx < nl0 ? L0[idx] : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = q + (uint32_t)k * 32u + lane; tw[k] = idx < nl0 ? __ldcg(P.touched + (m[k] >> 5)) : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            f[k] = false;
            if ((tw[k] >> (m[k] & 31u)) & 1u) { unsigned long long key; f[k] = pl_live(P, R, m[k], key) && key == kcur; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, f[k]);
            const uint32_t pos = cnt + __popc(b & lane_lt);
            if (f[k] && pos < room) out[pos] = m[k];
            cnt += __popc(b);
        }
    }
    return cnt < room ? cnt : room;
}

// smallest live rank above kprev among the touched members of the best class
__device__ __forceinline__ unsigned long long pl_kmin(const PlaceParams &P, const PlRowView &R, const uint32_t *L0, uint32_t nl0,
                                                      unsigned long long kprev, uint32_t lane) {
    unsigned long long kmin = PE_PREF_NONE;
    for (uint32_t q = 0; q < nl0; q += 128u) {
        uint32_t m[4], tw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = q + (uint32_t)k * 32u + lane; m[k] = idx < nl0 ? L0[idx] : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint32_t idx = q + (uint32_t)k * 32u + lane; tw[k] = idx < nl0 ? __ldcg(P.touched + (m[k] >> 5)) : 0u; }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if ((tw[k] >> (m[k] & 31u)) & 1u) {
                unsigned long long key;
                if (pl_live(P, R, m[k], key) && key > kprev && key < kmin) kmin = key;
            }
    }
    return pl_wmin64(kmin);
}

// What a task's staging needs that does not depend on the placements of the batch: loaded one chunk ahead, while the
// resolve phase of the current chunk runs.
struct PlPre { uint32_t row, task_off, tie_start; PlRowView R; };

__device__ __forceinline__ PlPre pl_prefetch(const PlaceParams &P, uint32_t task) {
    PlPre q;
    q.row = P.task_row[task];
    q.task_off = P.K.groups[P.b0 + task].task_off;     // (independent of the row record: both loads fly together)
    const ScanResult *sr = &P.scan[q.row];
    q.R.c0 = sr->c0; q.R.c1 = sr->c1;
    const uint4 meta = *reinterpret_cast<const uint4 *>(&sr->n0);   // n0, n1, tie_start, flags
    q.R.n0 = meta.x; q.R.n1 = meta.y; q.tie_start = meta.z; q.R.flags = meta.w;
    q.R.cpu_res = sr->cpu_res; q.R.mem_res = sr->mem_res; q.R.svccol = sr->svccol; q.R.max_replicas = sr->max_replicas;
    return q;
}

// The set of nodes given a task inside the current chunk (CTA 0's shared memory; one warp reads and writes it).
__device__ __forceinline__ bool pl_taken(const PlShared &S, uint32_t node) {
    for (uint32_t h = pl_hash(node);; h = (h + 1u) & (PE_PL_HASH - 1u)) {
        const uint32_t k = S.hkey[h];
        if (k == node) return true;
        if (k == PE_PL_EMPTY) return false;      // (at most PE_PL_CHUNK of PE_PL_HASH entries are ever used)
    }
}
__device__ __forceinline__ void pl_take(PlShared &S, uint32_t node) {   // (lanes of one warp may insert together)
    for (uint32_t h = pl_hash(node);; h = (h + 1u) & (PE_PL_HASH - 1u)) {
        const uint32_t old = atomicCAS(&S.hkey[h], PE_PL_EMPTY, node);
        if (old == PE_PL_EMPTY || old == node) return;
    }
}

__device__ __forceinline__ void pl_stage(const PlaceParams &P, const PlPre &pre, uint32_t *scr_a, uint32_t *scr_b, uint32_t *cand, PlShared *S0,
                                         uint32_t i, uint32_t lane, uint32_t &n_tail) {
    PlSlot *slot = &S0->slot[i];      // (cand: this warp's own shared memory; published to CTA 0 at the end)
    const uint32_t row = pre.row, task_off = pre.task_off;
    const PlRowView &R = pre.R;
    const bool eligible = (R.flags & PE_SR_K1) && pre.tie_start == 0u && (R.flags & PE_SR_COUNTS) && R.c0 != PE_PREF_NONE;
    uint32_t kind = 0, n_cand = 0, rs1 = 0, rs2 = 0, rs3 = 0;
    unsigned long long kv1 = PE_PREF_NONE, kv2 = PE_PREF_NONE, kv3 = PE_PREF_NONE;
    if (!eligible) {
        kind = 1;
    } else {
        const uint32_t want = min(i + 1u, (uint32_t)PE_PL_K);
        const uint32_t *L0 = P.L + (size_t)row * 2u * PE_LIST_CAP, *L1 = L0 + PE_LIST_CAP;
        const uint32_t nl0 = min(R.n0, P.list_cap);
        n_cand = pl_walk(L0, nl0, P.cursors + (size_t)row * 2u, P.touched, want, cand, lane);
        rs1 = rs2 = rs3 = n_cand;
        if (n_cand < want) {
            if (R.n0 > P.list_cap || !(R.flags & PE_SR_INLINE)) {
                // the class goes on beyond its list, or the row's feasibility can change in ways the tail does not
                // re-check (generic resources, host ports, recent failures): what was found is all there is here
                if (n_cand == 0u) kind = 1;
            } else {
                // ---- tail: the touched members of the best class at their LIVE rank, merged with the untouched
                // members of the second class.  Every other node ranked above c1 when the batch began and ranks only
                // grow inside a tick, so nothing else can rank <= c1.
                n_tail++;
                // a touched member carries at least one more task than when the batch began: its rank is >= kfloor
                const unsigned long long kfloor = make_pref(0u, (uint32_t)(R.c0 >> 32) & 0xFFFFFFu, (uint32_t)R.c0 + 1u);
                unsigned long long kprev = R.c0;
                for (int rank = 1; rank <= 3; rank++) {
                    if (n_cand >= want) break;
                    const uint32_t room = want - n_cand;
                    unsigned long long kcur = PE_PREF_NONE;
                    uint32_t cnt_a = 0;
                    bool have = false;
                    if (rank == 1) {         // the common case, without a pass over the whole class: members at the floor rank
                        kcur = kfloor;
                        cnt_a = pl_gather(P, R, L0, nl0, kcur, room, scr_a, lane);
                        have = cnt_a > 0u || kfloor == R.c1;
                    }
                    if (!have) {
                        kcur = pl_kmin(P, R, L0, nl0, kprev, lane);
                        if (R.c1 != PE_PREF_NONE && R.c1 > kprev && R.c1 < kcur) kcur = R.c1;
                        if (kcur == PE_PREF_NONE) break;
                        if (R.c1 != PE_PREF_NONE && kcur > R.c1) break;       // beyond the second class nothing is known
                        cnt_a = pl_gather(P, R, L0, nl0, kcur, room, scr_a, lane);
                    }
                    __syncwarp();
                    const uint32_t start = n_cand;
                    uint32_t added = 0;
                    if (kcur == R.c1) {
                        // merge with the first untouched members of the second class (both in node order)
                        const uint32_t nl1 = min(R.n1, P.list_cap);
                        const uint32_t cnt_b = pl_walk(L1, nl1, P.cursors + (size_t)row * 2u + 1u, P.touched, room, scr_b, lane);
                        __syncwarp();
                        uint32_t lim = 0xFFFFFFFFu;       // only what precedes the last listed untouched member is certain
                        if (cnt_b < room && R.n1 > nl1) lim = cnt_b ? scr_b[cnt_b - 1u] + 1u : 0u;
                        const uint32_t na = lim == 0xFFFFFFFFu ? cnt_a : pl_lower(scr_a, cnt_a, lim);
                        for (uint32_t e = lane; e < na; e += 32u) {
                            const uint32_t a = scr_a[e], pos = e + pl_lower(scr_b, cnt_b, a);
                            if (pos < room) cand[n_cand + pos] = a | PE_PL_TOUCHED;
                        }
                        for (uint32_t e = lane; e < cnt_b; e += 32u) {
                            const uint32_t b = scr_b[e], pos = e + pl_lower(scr_a, na, b);
                            if (pos < room) cand[n_cand + pos] = b;
                        }
                        added = min(room, na + cnt_b);
                    } else {
                        for (uint32_t e = lane; e < cnt_a; e += 32u) cand[n_cand + e] = scr_a[e] | PE_PL_TOUCHED;
                        added = cnt_a;
                    }
                    __syncwarp();
                    n_cand += added;
                    if (rank == 1) { rs1 = start; kv1 = kcur; rs2 = rs3 = n_cand; }
                    else if (rank == 2) { rs2 = start; kv2 = kcur; rs3 = n_cand; }
                    else { rs3 = start; kv3 = kcur; }
                    if (kcur == R.c1) break;
                    kprev = kcur;
                }
                if (n_cand == 0u) kind = 1;
            }
        }
    }
    if (kind) { n_cand = 0; rs1 = rs2 = rs3 = 0; }
    __syncwarp();
    for (uint32_t e = lane; e < n_cand; e += 32u) S0->cand[(size_t)i * PE_PL_KS + e] = cand[e];
    if (lane == 0) {
        S0->s_kind[i] = (uint8_t)kind; S0->s_ncand[i] = (uint16_t)n_cand;
        S0->s_rs1[i] = (uint16_t)rs1; S0->s_rs2[i] = (uint16_t)rs2; S0->s_rs3[i] = (uint16_t)rs3;
        slot->kind = kind; slot->n_cand = n_cand; slot->rs1 = rs1; slot->rs2 = rs2; slot->rs3 = rs3; slot->row = row;
        slot->task_off = task_off; slot->flags = R.flags;
        slot->c0 = R.c0; slot->kv1 = kv1; slot->kv2 = kv2; slot->kv3 = kv3;
        slot->svccol = const_cast<uint32_t *>(R.svccol); slot->cpu_res = R.cpu_res; slot->mem_res = R.mem_res; slot->max_replicas = R.max_replicas;
    }
}

__device__ __forceinline__ uint32_t pl_rank_of(const PlSlot &s, uint32_t j) {
    return (j >= s.rs1 ? 1u : 0u) + (j >= s.rs2 ? 1u : 0u) + (j >= s.rs3 ? 1u : 0u);
}
__device__ __forceinline__ unsigned long long pl_key_of(const PlSlot &s, uint32_t rank) {
    return rank == 0u ? s.c0 : rank == 1u ? s.kv1 : rank == 2u ? s.kv2 : s.kv3;
}

// ---- resolve: warp 0 of CTA 0, the chunk's tasks ONE AFTER THE OTHER, lanes = candidates ---------------------------------
// The reference's order, literally: task after task, each taking its first candidate that no earlier task holds.  What
// makes it fast is that nothing is left to look up: a task's candidates sit in shared memory in preference order, so the
// 32 lanes test 32 of them at once (one load of the candidate, one of its word of the touched bitmap, one ballot) and the
// loads of four tasks are in flight together; a task's choice is kept from the following three in registers.
// Final tasks are known by: the batch's touched bitmap kept in CTA 0's shared memory (tk, BM = true) for candidates that
// were untouched when the chunk began; the chunk's own set of joined nodes (hkey) for candidates flagged PE_PL_TOUCHED
// (their bit was set before the chunk began) and for every candidate when the node table is too big for tk.
// (Two earlier versions resolved the tasks in parallel -- lanes = tasks, deferred acceptance with the task order as the
// nodes' preference: ~8 rounds of ~4200 cycles per chunk with block barriers, ~23 rounds of ~1100 cycles with match.any
// in one warp.  The picks chase a frontier of low node indices, so about one round per four tasks is inherent; a lone
// warp pays ~8 cycles per dependent instruction, so the round count, not the work, set the time.)
struct PlProf { unsigned long long slow, rewin, retries; };

#define PE_PL_UNROLL 4

template <bool BM>
__device__ __forceinline__ void pl_resolve(const PlaceParams &P, PlShared &S, uint32_t *tk, uint32_t c0_task, uint32_t nc, uint32_t lane,
                                           uint32_t &n_amb, PlProf &prof) {
    uint32_t n_log = 0;
    bool leave = false;
    uint32_t cut = PE_NONE, done = nc;
    for (uint32_t t0 = 0; t0 < nc && !leave; t0 += PE_PL_UNROLL) {
        uint32_t c[PE_PL_UNROLL], w[PE_PL_UNROLL], ncand[PE_PL_UNROLL], pick[PE_PL_UNROLL];
#pragma unroll
        for (int u = 0; u < PE_PL_UNROLL; u++) {
            const uint32_t i = min(t0 + (uint32_t)u, nc - 1u);
            ncand[u] = t0 + (uint32_t)u < nc ? S.s_ncand[i] : 0u;
            c[u] = lane < ncand[u] ? S.cand[(size_t)i * PE_PL_KS + lane] : 0u;
            pick[u] = PE_NONE;
        }
        if (BM) {
#pragma unroll
            for (int u = 0; u < PE_PL_UNROLL; u++) w[u] = tk[PE_PL_NODE(c[u]) >> 5];
        }
#pragma unroll
        for (int u = 0; u < PE_PL_UNROLL; u++) {
            const uint32_t i = t0 + (uint32_t)u;
            if (i >= nc || leave) break;
            if (S.s_kind[i] != 0u) { cut = c0_task + i; done = i; leave = true; break; }      // not placeable here: the ordered sequencer takes over
            const uint32_t *cd = S.cand + (size_t)i * PE_PL_KS;
            const uint32_t rs1 = S.s_rs1[i], rs2 = S.s_rs2[i], rs3 = S.s_rs3[i];
            auto rank_of = [&](uint32_t j) -> uint32_t { return (j >= rs1 ? 1u : 0u) + (j >= rs2 ? 1u : 0u) + (j >= rs3 ? 1u : 0u); };
            // is candidate x (of this task) held by an earlier task?  `word`: its word of tk as loaded before this iteration
            auto held = [&](uint32_t x, uint32_t word) -> bool {
                const uint32_t n = PE_PL_NODE(x);
                bool h;
                if (BM && !(x & PE_PL_TOUCHED)) h = (word >> (n & 31u)) & 1u;
                else h = pl_taken(S, n);
#pragma unroll
                for (int v = 0; v < PE_PL_UNROLL; v++) if (v < u) h = h || n == pick[v];   // (chosen after the loads of this iteration)
                return h;
            };
            uint32_t jsel = PE_NONE, node = PE_NONE;
            {
                const bool free = lane < ncand[u] && !held(c[u], BM ? w[u] : 0u);
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, free);
                if (bal) { jsel = (uint32_t)__ffs((int)bal) - 1u; node = __shfl_sync(0xFFFFFFFFu, PE_PL_NODE(c[u]), (int)jsel); }
            }
            for (uint32_t jb = 32u; jsel == PE_NONE && jb < ncand[u]; jb += 32u) {      // beyond the first 32 candidates (rare)
                prof.slow += lane == 0 ? 1u : 0u;
                const uint32_t x = jb + lane < ncand[u] ? cd[jb + lane] : 0u;
                const uint32_t word = BM ? tk[PE_PL_NODE(x) >> 5] : 0u;
                const bool free = jb + lane < ncand[u] && !held(x, word);
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, free);
                if (bal) { jsel = jb + (uint32_t)__ffs((int)bal) - 1u; node = __shfl_sync(0xFFFFFFFFu, PE_PL_NODE(x), __ffs((int)bal) - 1); }
            }
            if (jsel == PE_NONE) {
                // out of candidates.  A list that was cut at PE_PL_K is staged again with what follows it (the chunk ends here,
                // the next one starts with this task); otherwise the task is not placeable here
                if (ncand[u] == (uint32_t)PE_PL_K && i + 1u > (uint32_t)PE_PL_K) { prof.retries += lane == 0 ? 1u : 0u; }
                else cut = c0_task + i;
                done = i; leave = true;
                break;
            }
            uint32_t jfin = jsel;
            if (jsel > 0u && rank_of(0u) < rank_of(jsel)) {
                // ---- it skipped a candidate that ranked strictly better than its choice when the chunk began.  That node
                // was taken inside the chunk, so its rank moved: recompute it from the chunk's log (one lane; rare).
                n_amb += lane == 0 ? 1u : 0u;
                uint32_t bn = node, bj = jsel;
                if (lane == 0) {
                    const PlSlot &sl = S.slot[i];
                    unsigned long long bk = pl_key_of(sl, rank_of(jsel));
                    for (uint32_t q = 0; q < jsel; q++) {
                        const uint32_t n = PE_PL_NODE(cd[q]);
                        const unsigned long long k0 = pl_key_of(sl, rank_of(q));
                        uint32_t svc = (uint32_t)(k0 >> 32) & 0xFFFFFFu, tot = (uint32_t)k0;
                        long long dcpu = 0, dmem = 0;
                        for (uint32_t e = 0; e < n_log; e++) {          // every placement of this chunk on that node
                            if (S.log_node[e] != n) continue;
                            const PlSlot &o = S.slot[S.log_task[e]];
                            tot++;
                            if (o.svccol == sl.svccol) svc++;
                            dcpu += o.cpu_res; dmem += o.mem_res;
                        }
                        bool ok = true;
                        if (sl.flags & PE_SR_RES)      // (the columns still hold the chunk-start amounts: reductions come after this phase)
                            ok = sl.cpu_res <= __ldcg(reinterpret_cast<const long long *>(P.T.cpu + n)) - dcpu &&
                                 sl.mem_res <= __ldcg(reinterpret_cast<const long long *>(P.T.mem + n)) - dmem;
                        if (sl.flags & PE_SR_MAXREP) ok = ok && (unsigned long long)svc < sl.max_replicas;
                        const unsigned long long k = make_pref(0u, svc, tot);
                        if (ok && (k < bk || (k == bk && n < bn))) { bk = k; bn = n; bj = q; }
                    }
                }
                node = __shfl_sync(0xFFFFFFFFu, bn, 0); jfin = __shfl_sync(0xFFFFFFFFu, bj, 0);
            }
            pick[u] = node;
            if (lane == 0) {
                S.log_node[n_log] = node; S.log_task[n_log] = (uint16_t)i; S.log_tail[n_log] = rank_of(jfin) != 0u ? 1 : 0;
                const bool was_touched = (cd[jfin] & PE_PL_TOUCHED) != 0u;
                if (BM) tk[node >> 5] |= 1u << (node & 31u);      // (one warp owns tk during the phase: plain read-modify-write)
                if (!BM || was_touched) pl_take(S, node);          // the set answers for what the bitmap cannot
            }
            n_log++;
            __syncwarp();
        }
        if (lane == 0) {        // the committer warps take the iteration's placements from here
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t *>(&S.n_pub) = n_log;
        }
    }
    __syncwarp();
    if (lane == 0) {
        S.n_log = n_log; S.cut = cut; S.done = done; S.stop = cut == PE_NONE ? 0u : 1u;
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&S.res_done) = 1u;
    }
}

static inline size_t place_smem_bytes(uint32_t tk_words) { return sizeof(PlShared) + (size_t)tk_words * 4 + 16; }
static_assert(sizeof(PlShared) + 18432u * 4 + 16 <= 232448, "k_place: shared memory budget");
static_assert(PE_PL_SCR >= 2 * PE_PL_K, "k_place: scratch buffers hold a candidate list");
static_assert(PE_PL_CHUNK <= 288, "k_place: one committer thread per log entry");
static_assert(4 * PE_PL_CHUNK <= PE_PL_HASH, "k_place: the set of taken nodes must stay sparse");
#define PE_PL_TK_MAX_WORDS 18432u     // 72 KB of touched bitmap next to the slots and the node table: up to ~590 k nodes

// (launched with a cluster dimension attribute: gridDim.x = the cluster size = 8 or 16)
__global__ void __launch_bounds__(PE_PL_THREADS, 1) k_place(const __grid_constant__ PlaceParams P) {
    extern __shared__ __align__(16) unsigned char pl_smem[];
    PlShared &S = *reinterpret_cast<PlShared *>(pl_smem);
    uint32_t *tk = reinterpret_cast<uint32_t *>(pl_smem + ((sizeof(PlShared) + 15) & ~(size_t)15));
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t crank = cluster.block_rank();
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    PlShared *S0 = cluster.map_shared_rank(&S, 0);     // CTA 0's copy: the slots every CTA stages into

    for (uint32_t h = tid; h < PE_PL_HASH; h += blockDim.x) S.hkey[h] = PE_PL_EMPTY;
    if (crank == 0) for (uint32_t w = tid; w < P.tk_words; w += blockDim.x) tk[w] = 0u;
    if (tid == 0) { S.n_log = 0; S.stop = 0; S.cut = PE_NONE; S.wd = 0; }
    uint32_t n_tail = 0, n_amb = 0, n_fast = 0, n_medium = 0, n_chunks = 0;
    long long cyc_stage = 0, cyc_resolve = 0, cyc_commit = 0;
    PlProf prof{};
    cluster.sync();

    const uint32_t chunk_cap = cluster.num_blocks() * PE_PL_WARPS;     // tasks per chunk = warps in the cluster
    const uint32_t slot_i = crank * PE_PL_WARPS + warp;          // the chunk task this warp stages
    PlPre pre{};
    uint32_t pre_task = PE_NONE;
    if (slot_i < P.B) { pre = pl_prefetch(P, slot_i); pre_task = slot_i; }
    for (uint32_t c0 = 0; c0 < P.B;) {
        const uint32_t nc = min(chunk_cap, P.B - c0);
        const long long t0 = clock64();
        if (crank == 0 && tid == 0) { S.n_pub = 0; S.res_done = 0; }           // (the cluster barrier below comes before anyone polls them)
        if (slot_i < nc) {
            if (pre_task != c0 + slot_i) pre = pl_prefetch(P, c0 + slot_i);      // (the previous chunk ended early)
            pl_stage(P, pre, S.scratch[warp][0], S.scratch[warp][1], S.scratch[warp][2], S0, slot_i, lane, n_tail);
        }
        cluster.sync();                                         // slots are in CTA 0's shared memory
        const long long t1 = clock64();
        pre_task = c0 + nc + slot_i;                            // the next chunk's task, if this one runs to its end: loads fly during resolve
        if (pre_task < P.B) pre = pl_prefetch(P, pre_task);
        if (crank == 0) {
            if (warp == 0) {
                if (P.tk_words) pl_resolve<true>(P, S, tk, c0, nc, lane, n_amb, prof);
                else pl_resolve<false>(P, S, tk, c0, nc, lane, n_amb, prof);
            } else if ((warp & 3u) != 0u && warp < 12u) {
                // ---- commit, WHILE the resolve warp works: NodeInfo.addTask (nodeinfo.go:125-153) for the chunk's placements,
                // from the slots alone.  One thread per log entry, on the three schedulers the resolve warp does not use; an
                // entry is taken as soon as the resolve warp has published it.
                const uint32_t e = ((warp >> 2) * 3u + (warp & 3u) - 1u) * 32u + lane;        // 0 .. 287 >= PE_PL_CHUNK
                bool mine = false;
                for (;;) {
                    if (e < *reinterpret_cast<volatile uint32_t *>(&S.n_pub)) { mine = true; break; }
                    if (*reinterpret_cast<volatile uint32_t *>(&S.res_done)) { mine = e < *reinterpret_cast<volatile uint32_t *>(&S.n_pub); break; }
                    __nanosleep(100);
                }
                if (mine) {
                    __threadfence_block();
                    const uint32_t n = S.log_node[e];
                    const PlSlot &sl = S.slot[S.log_task[e]];
                    P.K.out_node[sl.task_off] = n;
                    const long long cpu_res = sl.cpu_res, mem_res = sl.mem_res;
                    atomicAdd(&P.T.total[n], 1u);                   // (every task placed here counts: PE_SR_COUNTS)
                    if (atomicAdd(&sl.svccol[n], 1u) + 1u >= 0xFFFFFFu) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
                    if (cpu_res | mem_res) {
                        // the resolve warp's exact re-ranking reads these two columns as they were when the chunk began
                        // (and takes the chunk's own placements from the log): they move when it is done
                        while (!*reinterpret_cast<volatile uint32_t *>(&S.res_done)) __nanosleep(100);
                        if (cpu_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.cpu[n]), (unsigned long long)(-cpu_res));
                        if (mem_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.mem[n]), (unsigned long long)(-mem_res));
                    }
                    if (!(sl.flags & PE_SR_SIMPLE)) {
                        // generic resources / host ports: such a task is the FIRST on its node in this batch and tasks that
                        // join a touched node never reserve these, so one thread owns the cells
                        const pe_group *g = &P.K.groups[P.b0 + c0 + S.log_task[e]];
                        for (uint32_t w = 0; w < g->gen_cnt; w++) {
                            if (!gen_first_occurrence(P.K, *g, w)) continue;
                            int64_t *col = P.T.gen[P.K.gens[g->gen_off + w].kind];
                            col[n] = claim_cell(P.K, *g, w, col[n]);
                        }
                        for (uint32_t w = 0; w < g->port_cnt; w++) {
                            const uint32_t s = P.K.ports[g->port_off + w];
                            P.T.ports[s >> 5][n] |= 1u << (s & 31u);
                        }
                    }
                    atomicOr(&P.touched[n >> 5], 1u << (n & 31u));
                    if (S.log_tail[e]) n_medium++; else n_fast++;
                }
            }
            const long long t2 = clock64();
            __syncthreads();
            for (uint32_t h = tid; h < PE_PL_HASH; h += blockDim.x) S.hkey[h] = PE_PL_EMPTY;     // the set of taken nodes is per chunk
            __threadfence();        // the chunk's reductions are performed before anyone passes the barrier
            cyc_resolve += t2 - t1; cyc_commit += clock64() - t2;
        }
        cyc_stage += t1 - t0;
        n_chunks++;
        cluster.sync();
        if (S0->stop) break;
        c0 += S0->done;
        if (n_chunks > P.B + 16u) { if (crank == 0 && tid == 0) { S.wd |= 8u; S.cut = c0; } break; }   // (uniform)
    }
    cluster.sync();                                             // CTA 0's shared memory stays alive until every CTA is done with it
    // tallies: per-thread counters -> one atomic per warp
    n_tail = __reduce_add_sync(0xFFFFFFFFu, lane == 0 ? n_tail : 0u);
    n_amb = __reduce_add_sync(0xFFFFFFFFu, n_amb);                 // (counted by whichever resolve lane of CTA 0 re-ranked)
    const uint32_t n_retry = (uint32_t)prof.retries, n_slow = (uint32_t)prof.slow;      // (lane 0 of the resolve warp counts)
    n_fast = __reduce_add_sync(0xFFFFFFFFu, n_fast);
    n_medium = __reduce_add_sync(0xFFFFFFFFu, n_medium);
    if (lane == 0) {
        if (n_tail) atomicAdd(&P.ctr->place_tails, (unsigned long long)n_tail);
        if (n_fast) { atomicAdd(&P.ctr->fast_path, (unsigned long long)n_fast); }
        if (n_medium) atomicAdd(&P.ctr->medium_path, (unsigned long long)n_medium);
        if (n_fast + n_medium) atomicAdd(&P.ctr->placements, (unsigned long long)(n_fast + n_medium));
        if (n_amb) atomicAdd(&P.ctr->place_amb, (unsigned long long)n_amb);
    }
    if (crank == 0 && tid == 0) {
        const uint32_t cut = S.cut;
        *P.resume = cut == PE_NONE ? P.B : cut;
        if (cut != PE_NONE) atomicAdd(&P.ctr->place_cuts, 1ull);
        if (S.wd) atomicAdd(&P.ctr->prof[7], (unsigned long long)S.wd | 0x100ull);       // watchdog bits (+ 0x100 per event)
        atomicAdd(&P.ctr->place_chunks, (unsigned long long)n_chunks);
        atomicAdd(&P.ctr->place_tasks, (unsigned long long)(cut == PE_NONE ? P.B : cut));
        atomicAdd(&P.ctr->place_cyc[0], (unsigned long long)cyc_stage);
        atomicAdd(&P.ctr->place_cyc[1], (unsigned long long)cyc_resolve);
        atomicAdd(&P.ctr->place_cyc[2], (unsigned long long)cyc_commit);
        // resolve-phase diagnostics (bench.py "place.resolve"): passes over a 32-task group, proposal rounds, chunks ended early
        atomicAdd(&P.ctr->prof[0], (unsigned long long)n_slow); atomicAdd(&P.ctr->prof[2], (unsigned long long)n_retry);
    }
}

}  // namespace pe
