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
//   resolve (ordered, warp 0 of CTA 0, lanes = tasks, 32 tasks per group, the
//           chunk's groups one after the other): every lane proposes its first
//           candidate that no final task took; lanes that propose the same node find
//           each other with match.any and all but the lowest move on.  Lanes only
//           move forward and a node given up by one lane is held by a lower one, so
//           the fixed point is the sequential result.  A lane that skipped a
//           candidate ranked strictly better than its choice (the node was taken
//           inside the chunk: its rank moved) recomputes those ranks from the
//           chunk's log -- exactly.  (Which candidates are still free is worked
//           out by all sixteen warps of CTA 0 before each group's rounds.)
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
    uint32_t ctl_fa, ctl_leave, ctl_gdone;      // resolve: what warp 0 decided at the end of a pass, for the other warps
    unsigned long long avail[PE_PL_CHUNK];      // resolve: per task, which of its (up to 64) candidates no final task holds
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

// ---- resolve: CTA 0, 32 tasks (one group) at a time ------------------------------------------------------------------
// Deferred acceptance with the tasks' order as every node's preference, by WARP 0 with lanes = tasks: a lane proposes its
// first candidate that no FINAL task holds; lanes proposing the same node find each other with match.any and every lane
// but the lowest moves on to its next candidate.  A lane only gives a node up to a lower lane, and lower GROUPS are final
// before a group starts, so what the lanes hold when nobody moves any more is what the reference's one-task-at-a-time loop
// produces.  A round is a few dozen warp-synchronous instructions and one shared-memory load.
// Which candidates a final task holds is settled before the rounds by ALL SIXTEEN WARPS, two tasks each, lanes =
// candidates (two coalesced loads of the candidates, their words of the bitmap, two ballots -> a 64-bit mask per task):
// a lone warp pays ~8 cycles for every dependent instruction, so work that does not need the tasks' order is spread out.
// Final tasks are known by: the batch's touched bitmap kept in CTA 0's shared memory (tk, BM = true) for candidates that
// were untouched when the chunk began; the chunk's own set of joined nodes (hkey) for candidates flagged PE_PL_TOUCHED
// (their bit was set before the chunk began) and for every candidate when the node table is too big for tk.
// Measured alternatives, all exact (SM cycles of the phase per 128 tasks, cfg3 one-off, B200):
//   35k  thread = task for the whole chunk, proposals by atomicMin on a claim table in shared memory, a block barrier per
//        round, per-thread candidate windows re-read inside the rounds (the first version)
//   36k  the same with the masks below and no re-reads: ~8 rounds per chunk but ~3400 cycles per round -- the tasks at the
//        frontier propose the same few nodes, and their atomics on one shared-memory word are served one after the other
//   25k  one warp, a group at a time, per-lane candidate windows re-read inside the rounds
//   51k  the same with the masks computed by that one warp
//   87k  the tasks strictly one after the other, lanes = candidates
//   23.5k this version (5.5 rounds of ~560 cycles per group; one ballot per bit of the node index instead of match.any: 28k)
//   33k  this version with the NEXT group's masks worked out by the other warps while warp 0 runs the rounds (and every
//        popped candidate looked up once more): anything else that runs on the SM slows the lone warp down -- its rounds
//        took 2.1x as long -- which is also why the commit phase does not run next to the resolve phase
struct PlProf { unsigned long long passes, rounds, retries; long long cyc_filter, cyc_rounds, cyc_final; };   // (cycles: thread 0's view)

template <bool BM>
__device__ __forceinline__ void pl_resolve(const PlaceParams &P, PlShared &S, uint32_t *tk, uint32_t c0_task, uint32_t nc, uint32_t tid,
                                           uint32_t &n_amb, PlProf &prof) {
    const uint32_t lane = tid & 31u, warp = tid >> 5;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t n_log = 0, cut = PE_NONE, done = nc;      // (warp 0's)
    bool leave = false;
    for (uint32_t g0 = 0; g0 < nc && !leave; g0 += 32u) {
        const uint32_t g_end = min(g0 + 32u, nc);
        const uint32_t i = g0 + lane;
        const bool present = i < nc;
        const uint32_t is = present ? i : 0u;
        const uint32_t *cd = S.cand + (size_t)is * PE_PL_KS;
        const uint32_t n_cand = present ? S.s_ncand[is] : 0u;
        const uint32_t rs1 = S.s_rs1[is], rs2 = S.s_rs2[is], rs3 = S.s_rs3[is];
        const bool unusable = present && S.s_kind[is] != 0u;
        auto rank_of = [&](uint32_t j) -> uint32_t { return (j >= rs1 ? 1u : 0u) + (j >= rs2 ? 1u : 0u) + (j >= rs3 ? 1u : 0u); };
        uint32_t fa = g0;                           // tasks of the group below fa are final
        for (uint32_t pass = 0;; pass++) {
            // ---- every warp: which candidates of tasks fa + warp, fa + warp + 16 can still be proposed (lanes = candidates)
            const long long tq0 = clock64();
            for (uint32_t t = fa + warp; t < g_end; t += PE_PL_WARPS) {
                const uint32_t nn = S.s_kind[t] ? 0u : S.s_ncand[t];
                const uint32_t *row = S.cand + (size_t)t * PE_PL_KS;
                const uint32_t x0 = lane < nn ? row[lane] : 0u, x1 = 32u + lane < nn ? row[32u + lane] : 0u;
                bool f0 = lane < nn, f1 = 32u + lane < nn;
                if (BM) {
                    const uint32_t w0 = tk[PE_PL_NODE(x0) >> 5], w1 = tk[PE_PL_NODE(x1) >> 5];
                    if (f0) f0 = (x0 & PE_PL_TOUCHED) ? !pl_taken(S, PE_PL_NODE(x0)) : !((w0 >> (x0 & 31u)) & 1u);
                    if (f1) f1 = (x1 & PE_PL_TOUCHED) ? !pl_taken(S, PE_PL_NODE(x1)) : !((w1 >> (x1 & 31u)) & 1u);
                } else {
                    if (f0) f0 = !pl_taken(S, PE_PL_NODE(x0));
                    if (f1) f1 = !pl_taken(S, PE_PL_NODE(x1));
                }
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, f0), b1 = __ballot_sync(0xFFFFFFFFu, f1);
                if (lane == 0) S.avail[t] = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
            }
            __syncthreads();
            const long long tq1 = clock64();
            long long tq2 = tq1;
            if (warp == 0) {
                uint32_t bad = 32u;
                if (pass > 34u) {       // (cannot happen: every pass makes at least one more task final)
                    if (lane == 0) S.wd |= 2u;
                    cut = c0_task + fa; done = fa; leave = true;
                } else {
                    // ---- the rounds
                    const bool act = present && i >= fa && !unusable;
                    unsigned long long avail = act ? S.avail[is] : 0ull;
                    uint32_t j = 0, prop = 0;
                    bool moving = act, dead = false, have = false;
                    prof.passes += lane == 0 ? 1u : 0u;
                    uint32_t n_rounds = 0;
                    for (uint32_t r = 0; r < 64u * PE_PL_K; r++) {
                        if (moving) {
                            if (avail) { j = (uint32_t)__ffsll((long long)avail) - 1u; avail &= avail - 1ull; prop = PE_PL_NODE(cd[j]); have = true; }
                            else dead = true;                  // out of candidates
                            moving = false;
                        }
                        // lanes that propose the same node: all but the lowest move on (a higher lane that held it before finds out here)
                        // (one ballot per bit of the node index instead of match.any, unrolled so that the votes overlap, measured
                        // 608 against 560 cycles per round)
                        const uint32_t same = __match_any_sync(0xFFFFFFFFu, have ? prop : (0x80000000u | lane));
                        const bool lose = have && (same & lane_lt) != 0u;
                        n_rounds++;
                        if (!__any_sync(0xFFFFFFFFu, lose)) break;
                        if (lose) { have = false; moving = true; }
                    }
                    tq2 = clock64();
                    prof.rounds += lane == 0 ? n_rounds : 0u;
                    const bool valid = act && have;
                    const bool isbad = present && i >= fa && !valid;
                    // a task that ran out of a list that was cut at PE_PL_K candidates is staged again with what follows it
                    const bool isretry = isbad && !unusable && dead && n_cand == (uint32_t)PE_PL_K && i + 1u > (uint32_t)PE_PL_K;
                    const bool isamb = valid && j > 0u && rank_of(0u) < rank_of(j);
                    bad = __reduce_min_sync(0xFFFFFFFFu, (isbad || isamb) ? lane : 32u);     // lowest special lane of the group
                    // tasks below the first special task are final: log them (task order), mark their nodes taken
                    const bool fin = valid && lane < bad;
                    const uint32_t finm = __ballot_sync(0xFFFFFFFFu, fin);
                    if (fin) {
                        const uint32_t e = n_log + __popc(finm & lane_lt);
                        S.log_node[e] = prop; S.log_task[e] = (uint16_t)i; S.log_tail[e] = rank_of(j) != 0u ? 1 : 0;
                        if (BM) atomicOr(&tk[prop >> 5], 1u << (prop & 31u));
                        if (!BM || (cd[j] & PE_PL_TOUCHED)) pl_take(S, prop);      // the set answers for what the bitmap cannot
                    }
                    n_log += __popc(finm);
                    __syncwarp();
                    if (bad != 32u) {
                        const uint32_t ib = g0 + bad;
                        const bool b_retry = __shfl_sync(0xFFFFFFFFu, isretry ? 1u : 0u, (int)bad) != 0u;
                        const bool b_bad = __shfl_sync(0xFFFFFFFFu, isbad ? 1u : 0u, (int)bad) != 0u;
                        if (b_retry) { done = ib; leave = true; prof.retries += lane == 0 ? 1u : 0u; }     // the chunk ends here; the next one starts with this task
                        else if (b_bad) { cut = c0_task + ib; done = ib; leave = true; }                      // not placeable here: the ordered sequencer takes over
                        else {
                            // ---- it skipped a candidate that ranked strictly better than its choice when the chunk began.  That
                            // node was taken inside the chunk, so its rank moved: recompute the skipped candidates' ranks from the
                            // chunk's log -- the warp's lanes take one skipped candidate each (rare).
                            n_amb += lane == 0 ? 1u : 0u;
                            const uint32_t jb = __shfl_sync(0xFFFFFFFFu, j, (int)bad), pb = __shfl_sync(0xFFFFFFFFu, prop, (int)bad);
                            const PlSlot &sl = S.slot[ib];
                            const uint32_t *cdb = S.cand + (size_t)ib * PE_PL_KS;
                            const uint32_t q1 = S.s_rs1[ib], q2 = S.s_rs2[ib], q3 = S.s_rs3[ib];
                            auto rank_b = [&](uint32_t x) -> uint32_t { return (x >= q1 ? 1u : 0u) + (x >= q2 ? 1u : 0u) + (x >= q3 ? 1u : 0u); };
                            unsigned long long bk = pl_key_of(sl, rank_b(jb));     // the choice itself: every lane starts from it
                            uint32_t bn = pb, bj = jb;
                            for (uint32_t q = lane; q < jb; q += 32u) {
                                const uint32_t n = PE_PL_NODE(cdb[q]);
                                const unsigned long long k0 = pl_key_of(sl, rank_b(q));
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
                            // the warp's best: smallest (rank key, node)
                            const unsigned long long kmin = pl_wmin64(bk);
                            const uint32_t nmin = __reduce_min_sync(0xFFFFFFFFu, bk == kmin ? bn : 0xFFFFFFFFu);
                            const uint32_t who = (uint32_t)__ffs((int)__ballot_sync(0xFFFFFFFFu, bk == kmin && bn == nmin)) - 1u;
                            bn = nmin; bj = __shfl_sync(0xFFFFFFFFu, bj, (int)who);
                            if (lane == 0) {
                                // (a winner other than its own proposal was taken before: the bitmap has it already)
                                S.log_node[n_log] = bn; S.log_task[n_log] = (uint16_t)ib; S.log_tail[n_log] = rank_b(bj) != 0u ? 1 : 0;
                                if (BM) atomicOr(&tk[bn >> 5], 1u << (bn & 31u));
                                if (!BM || (cdb[bj] & PE_PL_TOUCHED)) pl_take(S, bn);
                            }
                            n_log++;
                            __syncwarp();
                        }
                        fa = ib + 1u;                   // the tasks above the special one propose again, from their first candidate
                    }
                }
                if (lane == 0) { S.ctl_fa = fa; S.ctl_leave = leave ? 1u : 0u; S.ctl_gdone = (bad == 32u || leave) ? 1u : 0u; }
            }
            __syncthreads();
            if (tid == 0) { prof.cyc_filter += tq1 - tq0; prof.cyc_rounds += tq2 - tq1; prof.cyc_final += clock64() - tq2; }
            fa = S.ctl_fa; leave = S.ctl_leave != 0u;
            if (S.ctl_gdone) break;
        }
    }
    if (tid == 0) { S.n_log = n_log; S.cut = cut; S.done = done; S.stop = cut == PE_NONE ? 0u : 1u; }
}

// commit of one log entry: NodeInfo.addTask (nodeinfo.go:125-153) from the slot alone
__device__ __forceinline__ void pl_commit(const PlaceParams &P, PlShared &S, uint32_t c0, uint32_t e, uint32_t &n_fast, uint32_t &n_medium) {
    const uint32_t n = S.log_node[e];
    const PlSlot &sl = S.slot[S.log_task[e]];
    P.K.out_node[sl.task_off] = n;
    const long long cpu_res = sl.cpu_res, mem_res = sl.mem_res;
    atomicAdd(&P.T.total[n], 1u);                   // (every task placed here counts: PE_SR_COUNTS)
    if (atomicAdd(&sl.svccol[n], 1u) + 1u >= 0xFFFFFFu) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
    if (cpu_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.cpu[n]), (unsigned long long)(-cpu_res));
    if (mem_res) atomicAdd(reinterpret_cast<unsigned long long *>(&P.T.mem[n]), (unsigned long long)(-mem_res));
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

static inline size_t place_smem_bytes(uint32_t tk_words) { return sizeof(PlShared) + (size_t)tk_words * 4 + 16; }
static_assert(sizeof(PlShared) + 18432u * 4 + 16 <= 232448, "k_place: shared memory budget");
static_assert(PE_PL_K <= 64, "k_place: a task's candidates fit a 64-bit mask");
static_assert(PE_PL_SCR >= 2 * PE_PL_K, "k_place: scratch buffers hold a candidate list");
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
        if (slot_i < nc) {
            if (pre_task != c0 + slot_i) pre = pl_prefetch(P, c0 + slot_i);      // (the previous chunk ended early)
            pl_stage(P, pre, S.scratch[warp][0], S.scratch[warp][1], S.scratch[warp][2], S0, slot_i, lane, n_tail);
        }
        cluster.sync();                                         // slots are in CTA 0's shared memory
        const long long t1 = clock64();
        pre_task = c0 + nc + slot_i;                            // the next chunk's task, if this one runs to its end: loads fly during resolve
        if (pre_task < P.B) pre = pl_prefetch(P, pre_task);
        if (crank == 0) {
            if (P.tk_words) pl_resolve<true>(P, S, tk, c0, nc, tid, n_amb, prof);
            else pl_resolve<false>(P, S, tk, c0, nc, tid, n_amb, prof);
            __syncthreads();
            const long long t2 = clock64();
            // ---- commit: NodeInfo.addTask (nodeinfo.go:125-153) for the chunk's placements, from the slots alone
            {
                const uint32_t n_log = S.n_log;
                for (uint32_t e = tid; e < n_log; e += blockDim.x) pl_commit(P, S, c0, e, n_fast, n_medium);
                __syncthreads();
            }
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
    const uint32_t n_retry = (uint32_t)prof.retries;      // (lane 0 of the resolve warp counts)
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
        atomicAdd(&P.ctr->prof[0], prof.passes); atomicAdd(&P.ctr->prof[1], prof.rounds); atomicAdd(&P.ctr->prof[2], (unsigned long long)n_retry);
        atomicAdd(&P.ctr->prof[3], (unsigned long long)prof.cyc_filter); atomicAdd(&P.ctr->prof[4], (unsigned long long)prof.cyc_rounds);
        atomicAdd(&P.ctr->prof[5], (unsigned long long)prof.cyc_final);
    }
}

}  // namespace pe
