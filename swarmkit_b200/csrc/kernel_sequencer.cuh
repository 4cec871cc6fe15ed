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
//    bitmap.
//  * generic path (any k; also the fall-back when the class was consumed): full
//    table evaluation against the live state, radix-select of the k smallest rank
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

#define PE_SEQ_THREADS 1024
#define PE_SEQ_KS 2048          // candidates staged in shared memory
#define PE_MAX_GEN_WANTS 8
#define PE_ST_FAILED 1u
#define PE_ST_BLOCKED 2u

struct SeqParams {
    DevTable T;
    TickDev K;
    uint32_t g_begin, g_end;
    const ScanResult *scan;  // [g_end - g_begin] or nullptr
    const uint32_t *E;       // class bitmaps, e_stride words per task
    uint32_t e_stride;
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

struct SeqShared {
    pe_group G;
    uint32_t red32[40];
    unsigned long long red64[40];
    uint32_t bins[256];
    uint32_t cnt8[8];
    uint32_t best, sel_bin, sel_before, ncand, neutral, any_pass, done, dead;
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

// First set bit of (E & ~touched) in tie order, restricted to words >= w0.
__device__ __forceinline__ uint32_t find_first(const uint32_t *Erow, uint32_t w0, uint32_t N, uint32_t ts,
                                               const uint32_t *touched, SeqShared &S) {
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x == 0) S.best = PE_NONE;
    __syncthreads();
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
            uint32_t c = v ? w * 32u + (uint32_t)__ffs((int)v) - 1u : PE_NONE;
            c = __reduce_min_sync(0xFFFFFFFFu, c);
            if (lane == 0 && c != PE_NONE) atomicMin(&S.best, c);
            __syncthreads();
            const uint32_t b = S.best;
            if (b != PE_NONE) return b;
        }
    }
    return PE_NONE;
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

    for (uint32_t w = tid; w < P.touched_words; w += nth) touched[w] = 0;
    if (tid == 0) S.neutral = 0;
    __syncthreads();

    for (uint32_t gi = P.g_begin; gi < P.g_end; gi++) {
        __syncthreads();
        if (tid < sizeof(pe_group) / 4) reinterpret_cast<uint32_t *>(&S.G)[tid] = reinterpret_cast<const uint32_t *>(&K.groups[gi])[tid];
        if (tid < 8) S.cnt8[tid] = 0;
        __syncthreads();
        const pe_group &G = S.G;
        const uint32_t k = G.n_tasks;
        uint32_t *ofail = K.out_fail + (size_t)gi * PE_NUM_FILTERS;
        if (k == 0) { if (tid < 8) ofail[tid] = 0; continue; }
        uint32_t *svccol = T.svc[G.svc_id];

        // ================= fast path ==========================================
        if (P.scan != nullptr && k == 1 && !S.neutral) {
            const ScanResult sr = P.scan[gi - P.g_begin];
            if (sr.c0 != PE_PREF_NONE) {
                const uint32_t *Erow = P.E + (size_t)(gi - P.g_begin) * P.e_stride;
                const uint32_t n = find_first(Erow, sr.w0, N, G.tie_start, touched, S);
                if (n != PE_NONE) {
                    if (tid == 0) {
                        const bool counts = (K.task_flags[G.task_off] & PE_T_COUNTS) != 0;
                        K.out_node[G.task_off] = n;
                        add_task_global(T, K, G, n, counts, P.ctr);
                        touched[n >> 5] |= 1u << (n & 31u);
                        if (!counts) S.neutral = 1;  // rank did not move: later class bitmaps may hide this node
                        P.ctr->fast_path++;
                        P.ctr->placements++;
                    }
                    if (tid < 8) ofail[tid] = 0;
                    continue;
                }
            }
        }

        // ================= generic path =======================================
        // ---- 1. evaluate every node against the live state (nodeset.go:57-121)
        uint32_t myF = 0;
        unsigned long long o64 = 0, a64 = ~0ull;
        uint32_t o32 = 0, a32 = ~0u;
        for (uint32_t n = tid; n < N; n += nth) {
            const uint32_t meta = T.meta[n];
            if (!(meta & PE_NODE_VALID)) { P.ff8[n] = 0xFF; continue; }
            const uint32_t sv = svccol[n];
            const uint32_t ff = eval_ff(T, K, G, n, meta, sv);
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
        if (tid == 0) { P.ctr->evals_generic += N; P.ctr->slow_path++; }
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
                        uint32_t c = excl;
                        for (int j = 0; j < 8; j++) {
                            if (rank <= c + loc[j]) { S.sel_bin = lane * 8 + j; S.sel_before = c; break; }
                            c += loc[j];
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
            P.ctr->placements += done;
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
            if (st_svc[i] >= 0xFFFFFFu) atomicOr(&P.ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
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
}

static inline size_t seq_dyn_smem_bytes(uint32_t touched_words_in_smem) {
    return (size_t)PE_SEQ_KS * (sizeof(CandKey) + 8 + 8 + 4 + 4 + 4 + 1) + (size_t)touched_words_in_smem * 4 + 16;
}

}  // namespace pe
