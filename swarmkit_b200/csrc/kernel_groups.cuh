// kernel_groups.cuh -- task GROUPS (k > 1 replicas of one spec) on the whole machine.
//
// The production-normal shape: every service created through the API carries a
// SpecVersion, so its pending tasks arrive as one group (scheduler.go:446-466) and
// scheduleTaskGroup (scheduler.go:694-748) runs once per group:
//
//   nodeSet.tree            nodeset.go:50-124    the k best feasible nodes (bounded max-heap)
//   orderedNodes            decision_tree.go:24-52  best -> worst
//   scheduleNTasksOnNodes   scheduler.go:844-924  the ordered fill
//   NodeInfo.addTask        nodeinfo.go:108-154   the reservations
//   Pipeline.Explain        pipeline.go:84-103    failure counters for what is left over
//
// Groups depend on each other through the node state, so they run one after the
// other -- but everything inside a group that touches all N nodes runs on every SM.
// One persistent cooperative kernel (one CTA per SM, grid barriers between phases)
// loops over the groups of a run:
//
//   1 evaluate   every CTA: Pipeline.Process + rank prefix for its slice of the tie order;
//                the distinct rank prefixes ("classes": a handful in practice) are counted
//                in a shared-memory table per CTA, then in a global one
//   2 threshold  CTA 0: sort the classes, find the class the k-th best node falls in
//   3 count      every CTA: members of each selected class inside its slice
//   4 offsets    CTA 0: exclusive scan over the CTAs, per class
//   5 compact    every CTA: its members written at their final position of the SORTED candidate
//                list -- classes ascend, the tie order ascends inside a class, so the list needs
//                no sort (the reference heap-sorts; the order is the same total order)
//   6 fill       CTA 0: the literal scheduleNTasksOnNodes loop on staged rows, write-back,
//                Explain counters
//
// A group with more than PE_GR_MAXCLS distinct rank prefixes among its feasible nodes
// is not handled here: the kernel stops in front of it and the one-CTA path of
// kernel_sequencer.cuh (exact for any key distribution) takes the rest of the run.
#pragma once
#include <cooperative_groups.h>

#include "kernel_sequencer.cuh"

namespace pe {

#define PE_GR_THREADS 256
#define PE_GR_WARPS (PE_GR_THREADS / 32)
#define PE_GR_MAXCLS 256          // distinct rank prefixes of one group that this path sorts by counting
#define PE_GR_LOCAL 128           // slots of the per-CTA class table
#define PE_GR_TABLE 1024          // slots of the global class table

struct GroupSel {                 // what CTA 0 publishes after the threshold phase
    uint32_t n_cls;               // selected classes (those the k best nodes fall in)
    uint32_t m;                   // candidates = min(feasible, k)
    uint32_t r_last;              // members taken from the last selected class
    uint32_t feasible;
    uint32_t fallback;            // 1: too many classes -- stop in front of this group
    uint32_t pad[3];
    unsigned long long pref[PE_GR_MAXCLS];   // ascending
    uint32_t base[PE_GR_MAXCLS];             // position of the class's first member in the candidate list
};

struct GroupsParams {
    DevTable T;
    TickDev K;
    uint32_t g_begin, g_end;
    uint8_t *ff8;                    // [cap] first failing filter (0 pass, 0xFF not in set)
    unsigned long long *pref64;      // [cap]
    CandKey *cand_g;                 // [st_cap] sorted candidates
    int64_t *st_cpu_g, *st_mem_g, *st_gen_g;
    uint32_t *st_svc_g, *st_tot_g, *st_placed_g;
    uint8_t *st_flags_g;
    uint32_t st_cap;
    unsigned long long *cls_key;     // [PE_GR_TABLE] global class table, all ones = empty
    uint32_t *cls_cnt;               // [PE_GR_TABLE]
    GroupSel *sel;
    uint32_t *cntmat;                // [gridDim.x][PE_GR_MAXCLS]: members of a class inside one CTA's slice of the tie order
    uint32_t *resume;                // out: groups of the run handled here (g_end - g_begin = all)
    DevCounters *ctr;
};

struct GroupsShared {
    pe_group G;
    GroupCtx C;
    unsigned long long lkey[PE_GR_LOCAL];
    uint32_t lcnt[PE_GR_LOCAL];
    uint32_t ccnt[PE_GR_WARPS][PE_GR_MAXCLS];
    unsigned long long skey[PE_GR_MAXCLS];
    uint32_t scnt[PE_GR_MAXCLS];
    uint32_t n_found, overflow;
    uint32_t cnt8[8], red32[8];
    uint32_t done, any_pass;
    // staged candidate rows for the fill (m <= PE_SEQ_KS; larger groups stage in global memory)
    int64_t st_cpu[PE_SEQ_KS], st_mem[PE_SEQ_KS];
    uint32_t st_svc[PE_SEQ_KS], st_tot[PE_SEQ_KS], st_placed[PE_SEQ_KS];
    uint8_t st_flags[PE_SEQ_KS];
    uint32_t st_node[PE_SEQ_KS];      // the candidates' node and failure band, next to the fill loop (global loads would stall it)
    uint8_t st_f5[PE_SEQ_KS];
    uint32_t all_count;               // every task of the group has DesiredState <= COMPLETED (the normal case)
};

__device__ __forceinline__ uint32_t gr_hash(unsigned long long k) {
    k ^= k >> 29; k *= 0xBF58476D1CE4E5B9ull; k ^= k >> 32;
    return (uint32_t)k;
}

// index of pref in the ascending array a[0, n), or n if absent / beyond
__device__ __forceinline__ uint32_t gr_class_of(const unsigned long long *a, uint32_t n, unsigned long long pref) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < pref) lo = mid + 1u; else hi = mid; }
    return (lo < n && a[lo] == pref) ? lo : n;
}

static inline size_t groups_smem_bytes() { return sizeof(GroupsShared) + 16; }

__global__ void __launch_bounds__(PE_GR_THREADS, 1) k_groups(const __grid_constant__ GroupsParams P) {
    extern __shared__ __align__(16) unsigned char gr_smem[];
    GroupsShared &S = *reinterpret_cast<GroupsShared *>(gr_smem);
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const DevTable &T = P.T;
    const TickDev &K = P.K;
    const uint32_t tid = threadIdx.x, nth = blockDim.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t N = T.n_nodes, nb = gridDim.x, b = blockIdx.x;
    // this CTA's slice of the TIE ORDER (positions, not node indices), cut into one share per warp: multiples of 32 positions
    const uint32_t per_w = (((N + nb - 1u) / nb + PE_GR_WARPS - 1u) / PE_GR_WARPS + 31u) & ~31u;
    const uint32_t per = per_w * PE_GR_WARPS;
    const uint32_t p_lo = min(b * per, N), p_hi = min(p_lo + per, N);
    const uint32_t w_lo = min(p_lo + warp * per_w, p_hi), w_hi = min(w_lo + per_w, p_hi);   // this warp's share
    unsigned long long n_placed = 0, n_slow = 0, n_evalg = 0;
    long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // block 0: cycles per phase (evaluate, threshold, count, offsets, compact, stage rows, fill, write-back + explain)
    long long tm = clock64();
#define GR_MARK(slot) do { const long long t_ = clock64(); cyc[slot] += t_ - tm; tm = t_; } while (0)

    uint32_t gi = P.g_begin;
    for (; gi < P.g_end; gi++) {
        // ---- the group and its evaluation context (every CTA keeps its own copy)
        __syncthreads();
        if (tid < sizeof(pe_group) / 4) reinterpret_cast<uint32_t *>(&S.G)[tid] = reinterpret_cast<const uint32_t *>(&K.groups[gi])[tid];
        for (uint32_t i = tid; i < PE_GR_LOCAL; i += nth) { S.lkey[i] = ~0ull; S.lcnt[i] = 0; }
        for (uint32_t i = tid; i < PE_GR_WARPS * PE_GR_MAXCLS; i += nth) (&S.ccnt[0][0])[i] = 0;
        if (tid < 8) S.cnt8[tid] = 0;
        if (tid == 0) { S.overflow = 0; S.n_found = 0; }
        __syncthreads();
        const pe_group &G = S.G;
        const uint32_t k = G.n_tasks;
        uint32_t *ofail = K.out_fail + (size_t)gi * PE_NUM_FILTERS;
        if (k == 0) { if (b == 0 && tid < 8) ofail[tid] = 0; continue; }
        if (tid < PE_CTX_MAXC && tid < G.con_cnt) {
            const pe_constraint c = K.cons[G.con_off + tid];
            S.C.con_col[tid] = T.attr[c.col]; S.C.con_val[tid] = c.value; S.C.con_neq[tid] = c.neq;
        }
        if (tid == 32) { S.C.svccol = T.svc[G.svc_id]; S.C.usable = G.con_cnt <= PE_CTX_MAXC ? 1u : 0u; }
        __syncthreads();
        uint32_t *svccol = S.C.svccol;
        const uint32_t ts = G.tie_start;

        // ================= 1 evaluate (nodeset.go:57-121: every node, live state) =================
        for (uint32_t p = p_lo + tid; p < p_hi; p += nth) {
            const uint32_t n = p + ts >= N ? p + ts - N : p + ts;
            const uint32_t meta = T.meta[n];
            if (!(meta & PE_NODE_VALID) || (G.leaf_cnt && !in_leaf(T, K, G, n))) { P.ff8[n] = 0xFF; continue; }   // not in the node set / under another leaf
            const uint32_t sv = svccol[n];
            const uint32_t ff = eval_ctx(T, K, G, S.C, n, meta, sv);
            const uint32_t fails = G.fail_cnt ? fail_count(K, G, n) : 0u;
            const unsigned long long pref = make_pref(fails, sv, T.total[n]);
            P.ff8[n] = (uint8_t)ff;
            P.pref64[n] = pref;
            if (ff == 0) {
                // count the class in this CTA's table
                bool placed_in = false;
                for (uint32_t probe = 0, h = gr_hash(pref) & (PE_GR_LOCAL - 1u); probe < PE_GR_LOCAL; probe++, h = (h + 1u) & (PE_GR_LOCAL - 1u)) {
                    unsigned long long cur = S.lkey[h];
                    if (cur == ~0ull) cur = atomicCAS(&S.lkey[h], ~0ull, pref), cur = cur == ~0ull ? pref : cur;
                    if (cur == pref) { atomicAdd(&S.lcnt[h], 1u); placed_in = true; break; }
                }
                if (!placed_in) S.overflow = 1;
            }
        }
        __syncthreads();
        // ... and in the global one
        for (uint32_t i = tid; i < PE_GR_LOCAL; i += nth) {
            const uint32_t c = S.lcnt[i];
            if (!c) continue;
            const unsigned long long pref = S.lkey[i];
            bool ok = false;
            for (uint32_t probe = 0, h = gr_hash(pref) & (PE_GR_TABLE - 1u); probe < PE_GR_TABLE; probe++, h = (h + 1u) & (PE_GR_TABLE - 1u)) {
                unsigned long long cur = P.cls_key[h];
                if (cur == ~0ull) cur = atomicCAS(&P.cls_key[h], ~0ull, pref), cur = cur == ~0ull ? pref : cur;
                if (cur == pref) { atomicAdd(&P.cls_cnt[h], c); ok = true; break; }
            }
            if (!ok) S.overflow = 1;
        }
        __syncthreads();
        if (tid == 0 && S.overflow) atomicExch(&P.sel->fallback, 1u);
        __threadfence();
        grid.sync();
        GR_MARK(0);

        // ================= 2 threshold (CTA 0) =================
        if (b == 0) {
            // gather the classes (and clear the table for the next group)
            for (uint32_t h = tid; h < PE_GR_TABLE; h += nth) {
                const unsigned long long pref = P.cls_key[h];
                if (pref != ~0ull) {
                    const uint32_t at = atomicAdd(&S.n_found, 1u);
                    if (at < PE_GR_MAXCLS) { S.skey[at] = pref; S.scnt[at] = P.cls_cnt[h]; }
                    P.cls_key[h] = ~0ull; P.cls_cnt[h] = 0;
                }
            }
            __syncthreads();
            const uint32_t nf = S.n_found;
            const bool too_many = nf > PE_GR_MAXCLS || P.sel->fallback != 0u;
            if (!too_many) {
                // rank sort (distinct keys): class i goes to position #{j : key_j < key_i}
                unsigned long long mykey = 0; uint32_t mycnt = 0, rank = 0;
                if (tid < nf) {
                    mykey = S.skey[tid]; mycnt = S.scnt[tid];
                    for (uint32_t j = 0; j < nf; j++) rank += S.skey[j] < mykey ? 1u : 0u;
                }
                __syncthreads();
                if (tid < nf) { P.sel->pref[rank] = mykey; S.scnt[rank] = mycnt; S.skey[rank] = mykey; }
                __syncthreads();
                if (tid == 0) {
                    uint32_t cum = 0, ncs = 0, r_last = 0;
                    for (uint32_t c = 0; c < nf; c++) {
                        P.sel->base[c] = cum;
                        const uint32_t cnt = S.scnt[c];
                        if (cum + cnt >= k) { ncs = c + 1u; r_last = k - cum; cum += cnt; for (uint32_t d = c + 1u; d < nf; d++) cum += S.scnt[d]; break; }
                        cum += cnt; ncs = c + 1u; r_last = cnt;
                    }
                    P.sel->n_cls = ncs; P.sel->r_last = r_last; P.sel->feasible = cum; P.sel->m = cum < k ? cum : k;
                }
            } else if (tid == 0) {
                P.sel->fallback = 1u;
            }
            __threadfence();
        }
        grid.sync();
        GR_MARK(1);
        if (P.sel->fallback) break;                      // (uniform: every CTA reads the same flag after the barrier)
        const uint32_t ncs = P.sel->n_cls, m = P.sel->m, r_last = P.sel->r_last, F = P.sel->feasible;

        // ================= 3 count the selected classes inside every warp's share =================
        for (uint32_t c = tid; c < ncs; c += nth) S.skey[c] = P.sel->pref[c];        // (the class keys, next to the lanes)
        __syncthreads();
        for (uint32_t p = w_lo + lane; p < w_hi; p += 32u) {
            const uint32_t n = p + ts >= N ? p + ts - N : p + ts;
            if (P.ff8[n] != 0) continue;
            const uint32_t c = gr_class_of(S.skey, ncs, P.pref64[n]);
            if (c < ncs) atomicAdd(&S.ccnt[warp][c], 1u);
        }
        __syncthreads();
        // (the per-warp counts stay in shared memory for phase 5; the grid only scans the CTAs' totals)
        for (uint32_t c = tid; c < ncs; c += nth) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < PE_GR_WARPS; w++) t += S.ccnt[w][c];
            P.cntmat[(size_t)b * PE_GR_MAXCLS + c] = t;
        }
        __threadfence();
        grid.sync();
        GR_MARK(2);

        // ================= 4 offsets: exclusive scan over the CTAs, per class (CTA 0, one warp per class) =================
        if (b == 0) {
            for (uint32_t c = warp; c < ncs; c += PE_GR_WARPS) {
                uint32_t run = P.sel->base[c];
                for (uint32_t q0 = 0; q0 < nb; q0 += 32u) {
                    const uint32_t q = q0 + lane;
                    const uint32_t v = q < nb ? P.cntmat[(size_t)q * PE_GR_MAXCLS + c] : 0u;
                    uint32_t incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += x; }
                    if (q < nb) P.cntmat[(size_t)q * PE_GR_MAXCLS + c] = run + incl - v;
                    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
            }
            __threadfence();
        }
        grid.sync();
        GR_MARK(3);

        // ================= 5 compact: members at their position of the sorted candidate list =================
        {
            // every warp walks its share in tie order; S.ccnt[warp][c] becomes the next free position of class c:
            // the CTA's base (grid scan) + what the lower warps of this CTA hold (their counts are still in shared memory)
            for (uint32_t c = tid; c < ncs; c += nth) {
                uint32_t run = P.cntmat[(size_t)b * PE_GR_MAXCLS + c];
#pragma unroll
                for (int w = 0; w < PE_GR_WARPS; w++) { const uint32_t t = S.ccnt[w][c]; S.ccnt[w][c] = run; run += t; }
            }
            __syncthreads();
            const uint32_t last_end = ncs ? P.sel->base[ncs - 1u] + r_last : 0u;
            for (uint32_t p0 = w_lo; p0 < w_hi; p0 += 32u) {
                const uint32_t p = p0 + lane;
                uint32_t c = ncs, n = 0;
                unsigned long long pref = 0;
                if (p < w_hi) {
                    n = p + ts >= N ? p + ts - N : p + ts;
                    if (P.ff8[n] == 0) { pref = P.pref64[n]; c = gr_class_of(S.skey, ncs, pref); }
                }
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, c);
                if (c < ncs) {
                    const uint32_t at = S.ccnt[warp][c] + __popc(peers & ((1u << lane) - 1u));
                    if (c + 1u < ncs || at < last_end) { P.cand_g[at].pref = pref; P.cand_g[at].tie = p; P.cand_g[at].node = n; }
                }
                __syncwarp();
                if (c < ncs && (uint32_t)(__ffs((int)peers) - 1) == lane) S.ccnt[warp][c] += __popc(peers);
                __syncwarp();
            }
            __threadfence();
        }
        grid.sync();
        GR_MARK(4);

        // ================= 6 fill, write-back, Explain (CTA 0) =================
        if (b == 0) {
            const CandKey *cand = P.cand_g;
            const bool in_smem = m <= PE_SEQ_KS;
            int64_t *st_cpu = in_smem ? S.st_cpu : P.st_cpu_g;
            int64_t *st_mem = in_smem ? S.st_mem : P.st_mem_g;
            uint32_t *st_svc = in_smem ? S.st_svc : P.st_svc_g;
            uint32_t *st_tot = in_smem ? S.st_tot : P.st_tot_g;
            uint32_t *st_placed = in_smem ? S.st_placed : P.st_placed_g;
            uint8_t *st_flags = in_smem ? S.st_flags : P.st_flags_g;
            if (tid == 0) { S.done = 0; S.any_pass = 0; S.all_count = 1; n_evalg += N; n_slow++; }
            __syncthreads();
            for (uint32_t i = tid; i < m; i += nth) {
                const uint32_t n = cand[i].node;
                st_cpu[i] = T.cpu[n]; st_mem[i] = T.mem[n]; st_svc[i] = svccol[n]; st_tot[i] = T.total[n];
                st_placed[i] = 0; st_flags[i] = 0;
                if (in_smem) { S.st_node[i] = n; S.st_f5[i] = (uint8_t)(cand[i].pref >> 56); }
                for (uint32_t w = 0; w < G.gen_cnt; w++)
                    if (gen_first_occurrence(K, G, w)) P.st_gen_g[(size_t)w * P.st_cap + i] = T.gen[K.gens[G.gen_off + w].kind][n];
            }
            {   // do all tasks move the spread counters?  (one pass in parallel instead of a global load per fill step)
                bool all = true;
                for (uint32_t ti = tid; ti < k; ti += nth) all = all && (K.task_flags[G.task_off + ti] & PE_T_COUNTS) != 0;
                if (!__all_sync(0xFFFFFFFFu, all) && lane == 0) S.all_count = 0;
            }
            __syncthreads();
            GR_MARK(5);
            // ---- The normal shape needs no loop.  If the first k candidates share one (failure band, service count) and every
            // task counts, scheduleNTasksOnNodes (scheduler.go:844-924) gives task t to candidate t: after its task a node
            // holds one more of the service than the next candidate, so the walk moves on every time (:899-905), the next
            // candidate is untouched (it passes Process as it did when it was staged), and with k <= m the first lap never ends.
            bool uniform = k <= m && S.all_count != 0u;
            if (uniform) {
                const uint32_t head = (uint32_t)(cand[0].pref >> 32);
                bool same = true;
                for (uint32_t i = tid; i < k; i += nth) same = same && (uint32_t)(cand[i].pref >> 32) == head;
                uniform = __syncthreads_and(same ? 1 : 0) != 0;
            }
            if (uniform) {
                for (uint32_t i = tid; i < k; i += nth) {
                    K.out_node[G.task_off + i] = cand[i].node;
                    st_mem[i] -= G.mem_res; st_cpu[i] -= G.cpu_res;
                    for (uint32_t w = 0; w < G.gen_cnt; w++)
                        if (gen_first_occurrence(K, G, w)) {
                            int64_t *cell = &P.st_gen_g[(size_t)w * P.st_cap + i];
                            *cell = claim_cell(K, G, w, *cell);
                        }
                    st_svc[i]++; st_tot[i]++; st_placed[i] = 1;
                }
                if (tid == 0) { S.done = k; S.any_pass = 1; n_placed += k; }
            }
            // ---- scheduleNTasksOnNodes, the literal loop: one thread, staged rows
            if (!uniform && tid == 0 && m > 0) {
                const uint32_t fm = G.filter_mask;
                const bool f_res = (fm >> PE_F_RESOURCE) & 1u, f_port = ((fm >> PE_F_HOSTPORT) & 1u) && G.port_cnt > 0;
                const bool f_max = (fm >> PE_F_MAXREPLICAS) & 1u;
                const bool all_count = S.all_count != 0u;
                uint32_t cnt[PE_NUM_FILTERS];
                for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] = 0;
                uint32_t done = 0, any_pass = 0;
                // `it` of the reference only ever matters modulo m and through the test `it + 1 < m`: keep the position and
                // whether the first lap is over (no 64-bit division on the one thread everybody waits for)
                uint32_t pos = 0;
                bool first_lap = true;
                auto node_of = [&](uint32_t i) -> uint32_t { return in_smem ? S.st_node[i] : cand[i].node; };
                auto f5_of = [&](uint32_t i) -> uint32_t { return in_smem ? (uint32_t)S.st_f5[i] : (uint32_t)(cand[i].pref >> 56); };
                for (uint32_t ti = 0; ti < k; ti++) {
                    const uint32_t i = pos;
                    K.out_node[G.task_off + ti] = node_of(i);
                    const bool counts = all_count || (K.task_flags[G.task_off + ti] & PE_T_COUNTS) != 0;
                    st_mem[i] -= G.mem_res;                  // NodeInfo.addTask on the staged row (nodeinfo.go:125-153)
                    st_cpu[i] -= G.cpu_res;
                    for (uint32_t w = 0; w < G.gen_cnt; w++)
                        if (gen_first_occurrence(K, G, w)) {
                            int64_t *cell = &P.st_gen_g[(size_t)w * P.st_cap + i];
                            *cell = claim_cell(K, G, w, *cell);
                        }
                    if (G.port_cnt) st_flags[i] |= PE_ST_BLOCKED;
                    if (counts) { st_svc[i]++; st_tot[i]++; }
                    st_placed[i]++;
                    done++;
                    if (done == k) break;
                    if (first_lap && pos + 1u < m) {  // :899-905 first pass: move on only if the next node is now strictly better
                        const uint32_t j = pos + 1u;
                        const uint32_t fa = f5_of(j), fb = f5_of(i);
                        const bool less = fa != fb ? fa < fb : (st_svc[j] != st_svc[i] ? st_svc[j] < st_svc[i] : st_tot[j] < st_tot[i]);
                        if (less) pos++;
                    } else {
                        first_lap = false;     // :906-910 later passes: round-robin
                        pos = pos + 1u == m ? 0u : pos + 1u;
                    }
                    uint32_t walked = 0;
                    bool dead = false;
                    for (;;) {         // :912-920
                        const uint32_t j = pos;
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
                        // (the reference's `it` keeps counting: once it has passed m - 1 the first lap is over for good)
                        if (pos + 1u == m) { pos = 0; first_lap = false; } else pos++;
                        if (++walked == m) { dead = true; break; }
                    }
                    if (dead) break;
                }
                S.done = done; S.any_pass = any_pass;
                for (int f = 0; f < PE_NUM_FILTERS; f++) S.cnt8[f] = cnt[f];
                n_placed += done;
            }
            __syncthreads();
            GR_MARK(6);
            const uint32_t done = S.done;
            // ---- write the staged rows back
            for (uint32_t i = tid; i < m; i += nth) {
                if (!st_placed[i]) continue;
                const uint32_t n = cand[i].node;
                T.cpu[n] = st_cpu[i]; T.mem[n] = st_mem[i]; T.total[n] = st_tot[i]; svccol[n] = st_svc[i];
                if (st_svc[i] >= 0xFFFFF0u) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
                for (uint32_t w = 0; w < G.gen_cnt; w++)
                    if (gen_first_occurrence(K, G, w)) T.gen[K.gens[G.gen_off + w].kind][n] = P.st_gen_g[(size_t)w * P.st_cap + i];
                for (uint32_t q = 0; q < G.port_cnt; q++) {
                    const uint32_t s = K.ports[G.port_off + q];
                    T.ports[s >> 5][n] |= 1u << (s & 31u);
                }
            }
            for (uint32_t ti = done + tid; ti < k; ti += nth) K.out_node[G.task_off + ti] = PE_NONE;
            // ---- Explain counters for unplaced tasks (pipeline.go:56-68).  The tree build leaves "first failing filter"
            // counts of the nodes visited after the last one that entered the heap; a node is visited while the heap is not
            // full or when it ranks below the heap's worst member (nodeset.go:111-120).
            if (done < k) {
                if (!S.any_pass) {
                    uint32_t posL = 0;
                    const bool haveL = m > 0;
                    unsigned long long Mp = ~0ull;
                    uint32_t Mt = ~0u;
                    if (m > 0) {
                        uint32_t mx = 0;
                        for (uint32_t i = tid; i < m; i += nth) mx = max(mx, cand[i].tie);
                        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
                        if (lane == 0) S.red32[warp] = mx;
                        __syncthreads();
                        for (uint32_t w = 0; w < (nth >> 5); w++) posL = max(posL, S.red32[w]);
                        __syncthreads();
                        if (m == k) { Mp = cand[m - 1].pref; Mt = cand[m - 1].tie; }
                    }
                    uint32_t c[PE_NUM_FILTERS];
                    for (int f = 0; f < PE_NUM_FILTERS; f++) c[f] = 0;
                    for (uint32_t n = tid; n < N; n += nth) {
                        const uint32_t ff = P.ff8[n];
                        if (ff == 0 || ff == 0xFF) continue;
                        const uint32_t tp = tie_pos(n, ts, N);
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
            __threadfence();
        }
        (void)F;
        grid.sync();
        GR_MARK(7);      // the reservations of this group are in the columns before the next group reads them
    }
    if (b == 0 && tid == 0) {
        *P.resume = gi - P.g_begin;
        P.sel->fallback = 0;
        P.ctr->slow_path += n_slow;
        P.ctr->placements += n_placed;
        P.ctr->evals_generic += n_evalg;
        for (int q = 0; q < 8; q++) P.ctr->prof[8 + q] += (unsigned long long)cyc[q];
    }
}

}  // namespace pe
