// kernel_classify.cuh -- what makes batching pay: work shared between pending
// one-off tasks that the reference evaluates one at a time.
//
// The reference walks every node for every k=1 group (scheduler.go:467-469 ->
// scheduleTaskGroup -> nodeSet.tree, nodeset.go:57-121).  Inside one tick the
// node attributes the Ready / Plugin / Constraint / Platform filters read do
// not change (only reservations do), and many one-off tasks carry the same
// descriptor.  So, per run of k=1 groups:
//
//   k_classify  gives every group the representative of (a) its full descriptor
//               class (everything but task_off) and (b) its static signature
//               (the inputs of the four attribute filters), by content, through
//               two open-addressing hash tables;
//   k_static    evaluates the attribute filters once per signature and node and
//               keeps the result as a bitmap row (the "pre-evaluated predicate
//               mask" encoding of SURVEY 8d);
//   k_rows      per batch, collapses tasks of the same descriptor class into
//               one scan ROW (they would get bit-identical scan results: the
//               scan reads batch-start state only).
//
// The scan kernel then evaluates only the state-dependent part per (row, node).
// Placements stay exactly those of the sequential reference: the sequencer
// still visits tasks one by one (kernel_sequencer.cuh).
#pragma once
#include "kernels_common.cuh"

namespace pe {

#define PE_STATIC_FM ((1u << PE_F_READY) | (1u << PE_F_PLUGIN) | (1u << PE_F_CONSTRAINT) | (1u << PE_F_PLATFORM))

__device__ __forceinline__ unsigned long long hmix(unsigned long long h, unsigned long long v) {
    h ^= v;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    return h;
}
__device__ __forceinline__ unsigned long long hwords(unsigned long long h, const void *p, uint32_t nwords) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    for (uint32_t i = 0; i < nwords; i++) h = hmix(h, w[i]);
    return hmix(h, nwords);
}
__device__ __forceinline__ bool words_eq(const void *a, const void *b, uint32_t nwords) {
    if (a == b) return true;
    const uint32_t *x = reinterpret_cast<const uint32_t *>(a), *y = reinterpret_cast<const uint32_t *>(b);
    for (uint32_t i = 0; i < nwords; i++) if (x[i] != y[i]) return false;
    return true;
}

// inputs of ReadyFilter / PluginFilter / ConstraintFilter / PlatformFilter
__device__ __forceinline__ unsigned long long static_hash(const TickDev &K, const pe_group &g) {
    unsigned long long h = 0x243F6A8885A308D3ull;
    h = hmix(h, (g.filter_mask & PE_STATIC_FM) | ((unsigned long long)g.flags << 32));
    h = hmix(h, (g.flags & PE_G_LOG_DRIVER) ? g.log_plugin : PE_NONE);
    h = hwords(h, K.cons + g.con_off, g.con_cnt * (uint32_t)(sizeof(pe_constraint) / 4));
    h = hwords(h, K.ips + g.ip_off, g.ip_cnt * (uint32_t)(sizeof(pe_ip_constraint) / 4));
    h = hwords(h, K.plats + g.plat_off, g.plat_cnt * (uint32_t)(sizeof(pe_platform) / 4));
    h = hwords(h, K.plugs + g.plug_off, g.plug_cnt);
    return h;
}
__device__ __forceinline__ bool static_eq(const TickDev &K, const pe_group &a, const pe_group &b) {
    if ((a.filter_mask & PE_STATIC_FM) != (b.filter_mask & PE_STATIC_FM) || a.flags != b.flags) return false;
    if ((a.flags & PE_G_LOG_DRIVER) && a.log_plugin != b.log_plugin) return false;
    if (a.con_cnt != b.con_cnt || a.ip_cnt != b.ip_cnt || a.plat_cnt != b.plat_cnt || a.plug_cnt != b.plug_cnt) return false;
    return words_eq(K.cons + a.con_off, K.cons + b.con_off, a.con_cnt * (uint32_t)(sizeof(pe_constraint) / 4)) &&
           words_eq(K.ips + a.ip_off, K.ips + b.ip_off, a.ip_cnt * (uint32_t)(sizeof(pe_ip_constraint) / 4)) &&
           words_eq(K.plats + a.plat_off, K.plats + b.plat_off, a.plat_cnt * (uint32_t)(sizeof(pe_platform) / 4)) &&
           words_eq(K.plugs + a.plug_off, K.plugs + b.plug_off, a.plug_cnt);
}
// the rest of the descriptor (everything but task_off; n_tasks is 1 on this path)
__device__ __forceinline__ unsigned long long dyn_hash(const TickDev &K, const pe_group &g, unsigned long long h) {
    h = hmix(h, g.svc_id | ((unsigned long long)g.filter_mask << 32));
    h = hmix(h, (unsigned long long)g.cpu_res);
    h = hmix(h, (unsigned long long)g.mem_res);
    h = hmix(h, g.max_replicas);
    h = hmix(h, g.tie_start | ((unsigned long long)K.task_flags[g.task_off] << 32));
    h = hwords(h, K.gens + g.gen_off, g.gen_cnt * (uint32_t)(sizeof(pe_generic_want) / 4));
    h = hwords(h, K.ports + g.port_off, g.port_cnt);
    h = hwords(h, K.fails + g.fail_off, g.fail_cnt * (uint32_t)(sizeof(pe_node_fail) / 4));
    return h;
}
__device__ __forceinline__ bool dyn_eq(const TickDev &K, const pe_group &a, const pe_group &b) {
    if (a.svc_id != b.svc_id || a.filter_mask != b.filter_mask || a.cpu_res != b.cpu_res || a.mem_res != b.mem_res ||
        a.max_replicas != b.max_replicas || a.tie_start != b.tie_start || a.n_tasks != b.n_tasks ||
        K.task_flags[a.task_off] != K.task_flags[b.task_off])
        return false;
    if (a.gen_cnt != b.gen_cnt || a.port_cnt != b.port_cnt || a.fail_cnt != b.fail_cnt) return false;
    return words_eq(K.gens + a.gen_off, K.gens + b.gen_off, a.gen_cnt * (uint32_t)(sizeof(pe_generic_want) / 4)) &&
           words_eq(K.ports + a.port_off, K.ports + b.port_off, a.port_cnt) &&
           words_eq(K.fails + a.fail_off, K.fails + b.fail_off, a.fail_cnt * (uint32_t)(sizeof(pe_node_fail) / 4));
}

struct ClassifyParams {
    TickDev K;
    uint32_t g0, n;              // the run of k=1 groups [g0, g0 + n)
    uint32_t *ht_full, *ht_static;   // open addressing, entries = absolute group index or PE_NONE
    uint32_t ht_mask;
    uint32_t *dcls;              // [n] representative group of the full descriptor class
    uint32_t *scls;              // [n] representative group of the static signature
    uint32_t *srow;              // [n] indexed by (representative - g0): compact signature id
    uint32_t *static_reps;       // [n] signature id -> representative group
    uint32_t *counters;          // [0] signatures, [1] descriptor classes
};

__global__ void __launch_bounds__(256) k_classify(const ClassifyParams P) {
    const TickDev &K = P.K;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = P.g0 + i;
        const pe_group g = K.groups[gi];
        const unsigned long long hs = static_hash(K, g);
        uint32_t rep = PE_NONE;
        for (uint32_t s = (uint32_t)hs & P.ht_mask;; s = (s + 1u) & P.ht_mask) {
            uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&P.ht_static[s]);   // most groups find their class already there
            if (cur == PE_NONE) cur = atomicCAS(&P.ht_static[s], PE_NONE, gi);
            if (cur == PE_NONE) {
                rep = gi;
                const uint32_t sid = atomicAdd(&P.counters[0], 1u);
                P.srow[i] = sid;
                P.static_reps[sid] = gi;
                break;
            }
            if (cur == gi || static_eq(K, g, K.groups[cur])) { rep = cur; break; }
        }
        P.scls[i] = rep;
        const unsigned long long hf = dyn_hash(K, g, hs);
        for (uint32_t s = (uint32_t)(hf >> 7) & P.ht_mask;; s = (s + 1u) & P.ht_mask) {
            uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&P.ht_full[s]);
            if (cur == PE_NONE) cur = atomicCAS(&P.ht_full[s], PE_NONE, gi);
            if (cur == PE_NONE) { rep = gi; atomicAdd(&P.counters[1], 1u); break; }
            if (cur == gi) { rep = cur; break; }
            const pe_group o = K.groups[cur];
            if (dyn_eq(K, g, o) && static_eq(K, g, o)) { rep = cur; break; }
        }
        P.dcls[i] = rep;
    }
}

// ---- the attribute filters, once per (signature, node) ----------------------
// ReadyFilter filter.go:40-43, PluginFilter :141-183, ConstraintFilter :241-243
// (constraint.NodeMatches constraint.go:107-207), PlatformFilter :272-312.
__device__ __forceinline__ bool static_ok(const DevTable &T, const TickDev &K, const pe_group &g, uint32_t n, uint32_t meta) {
    const uint32_t fm = g.filter_mask;
    bool ok = (meta & PE_NODE_VALID) != 0 && (g.leaf_cnt == 0u || in_leaf(T, K, g, n));   // (leaf visits never take the batched path: engine.cu)
    if (fm & (1u << PE_F_READY)) ok = ok && (meta & PE_NODE_READY);
    if ((fm & (1u << PE_F_PLUGIN)) && (meta & PE_NODE_HAS_ENGINE)) {
        for (uint32_t i = 0; i < g.plug_cnt; i++) {
            const uint32_t s = K.plugs[g.plug_off + i];
            ok = ok && ((T.plug[s >> 5][n] >> (s & 31u)) & 1u);
        }
        if (g.flags & PE_G_LOG_DRIVER) {
            const uint32_t s = g.log_plugin;
            const bool exists = (T.plug[s >> 5][n] >> (s & 31u)) & 1u;
            ok = ok && (exists || !(meta & PE_NODE_HAS_LOGPLUGIN));
        }
    }
    if (fm & (1u << PE_F_CONSTRAINT)) {
        if (g.flags & PE_G_CONSTRAINT_NEVER) ok = false;
        for (uint32_t i = 0; i < g.con_cnt; i++) {
            const pe_constraint c = K.cons[g.con_off + i];
            ok = ok && ((T.attr[c.col][n] == c.value) != (c.neq != 0u));
        }
        for (uint32_t i = 0; i < g.ip_cnt; i++) {
            const pe_ip_constraint c = K.ips[g.ip_off + i];
            bool hit = (meta & PE_NODE_IP_VALID) != 0;
            if (hit && c.is_cidr) hit = ((meta & PE_NODE_IP_V4) != 0) == (c.is_v4 != 0);
            if (hit) {
                const uint4 a = T.ip[n];
                hit = (a.x & c.mask[0]) == c.net[0] && (a.y & c.mask[1]) == c.net[1] && (a.z & c.mask[2]) == c.net[2] &&
                      (a.w & c.mask[3]) == c.net[3];
            }
            ok = ok && (hit != (c.neq != 0u));
        }
    }
    if ((fm & (1u << PE_F_PLATFORM)) && g.plat_cnt) {
        bool any = false;
        if (meta & PE_NODE_HAS_PLATFORM) {
            const uint32_t os = (meta >> 8) & 0xFFu, arch = (meta >> 16) & 0xFFu;
            for (uint32_t i = 0; i < g.plat_cnt; i++) {
                const pe_platform p = K.plats[g.plat_off + i];
                any = any || ((p.arch_id == 0 || p.arch_id == arch) && (p.os_id == 0 || p.os_id == os));
            }
        }
        ok = ok && any;
    }
    return ok;
}

struct StaticParams {
    DevTable T;
    TickDev K;
    const uint32_t *reps;      // [*n_reps] representative group per bitmap row
    const uint32_t *n_reps;    // device count
    uint32_t *S;               // [rows][s_stride] feasibility bitmaps
    uint32_t s_stride;         // words per row (multiple of 32, >= padded rows / 32)
    uint32_t lo, hi;           // node range this rank evaluates (bits outside stay 0)
    DevCounters *ctr;
};

// One warp per (row, chunk of 32 words = 1024 nodes); lane = node inside a step.
__global__ void __launch_bounds__(256) k_static(const StaticParams P) {
    const uint32_t lane = threadIdx.x & 31u, wpb = blockDim.x >> 5;
    const uint32_t n_rows = *P.n_reps;
    const uint32_t chunks = P.s_stride >> 5;
    const unsigned long long units = (unsigned long long)n_rows * chunks;
    for (unsigned long long u = (unsigned long long)blockIdx.x * wpb + (threadIdx.x >> 5); u < units; u += (unsigned long long)gridDim.x * wpb) {
        const uint32_t row = (uint32_t)(u / chunks), chunk = (uint32_t)(u % chunks);
        const pe_group g = P.K.groups[P.reps[row]];
        uint32_t mine = 0;
        const uint32_t base = chunk * 1024u;
        if (base < P.hi && base + 1024u > P.lo) {
            for (uint32_t s = 0; s < 32u; s++) {
                const uint32_t n = base + s * 32u + lane;
                bool ok = false;
                if (n >= P.lo && n < P.hi) ok = static_ok(P.T, P.K, g, n, P.T.meta[n]);
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, ok);
                if (lane == s) mine = w;
            }
        }
        P.S[(size_t)row * P.s_stride + chunk * 32u + lane] = mine;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.ctr->static_evals, (unsigned long long)n_rows * (P.hi - P.lo));
}

// ---- per batch: tasks -> rows -------------------------------------------------
struct RowsParams {
    const uint32_t *dcls, *scls, *srow;   // run-relative arrays of k_classify
    uint32_t g0;               // run start (absolute group index)
    uint32_t b0, B;            // batch [b0, b0 + B), absolute
    uint32_t *mark;            // [run length] last batch stamp that saw the class (zeroed per run)
    uint32_t *firstof;         // [run length] first task of the class in that batch
    uint32_t *rowof;           // [run length] row of the class in that batch
    uint32_t stamp;            // batch number + 1
    uint32_t static_cached;    // 1: bitmap rows are per signature for the whole run; 0: one per row of this batch
    uint32_t *row_group;       // [B] row -> representative group (absolute)
    uint32_t *row_srow;        // [B] row -> bitmap row
    uint32_t *task_row;        // [B] task (batch-relative) -> row
    uint32_t *n_rows;          // device count
};

// Rows are numbered in order of first appearance in the batch, so every rank of a node-sharded engine
// builds the same table (the rows index the partial results the ranks exchange).  One CTA.
__global__ void __launch_bounds__(1024) k_rows(const RowsParams P) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t total;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t t = tid; t < P.B; t += blockDim.x) {
        const uint32_t c = P.dcls[P.b0 + t - P.g0] - P.g0;
        if (atomicExch(&P.mark[c], P.stamp) != P.stamp) P.firstof[c] = 0xFFFFFFFFu;   // first visit of the class in this batch
    }
    __syncthreads();
    for (uint32_t t = tid; t < P.B; t += blockDim.x) atomicMin(&P.firstof[P.dcls[P.b0 + t - P.g0] - P.g0], t);
    __syncthreads();
    // thread k owns the contiguous tasks [k * per, (k + 1) * per): count its representatives, scan, number them
    const uint32_t per = (P.B + blockDim.x - 1u) / blockDim.x;
    const uint32_t lo = min(tid * per, P.B), hi = min(lo + per, P.B);
    uint32_t cnt = 0;
    for (uint32_t t = lo; t < hi; t++) cnt += P.firstof[P.dcls[P.b0 + t - P.g0] - P.g0] == t ? 1u : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += x; }
    if (lane == 31u) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < (blockDim.x >> 5) ? wsum[lane] : 0u, w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, w, o); if (lane >= (uint32_t)o) w += x; }
        wsum[lane] = w - v;                 // exclusive
        if (lane == 31u) total = w;
    }
    __syncthreads();
    uint32_t r = wsum[warp] + incl - cnt;
    for (uint32_t t = lo; t < hi; t++) {
        const uint32_t i = P.b0 + t - P.g0;
        const uint32_t c = P.dcls[i] - P.g0;
        if (P.firstof[c] == t) {
            P.rowof[c] = r;
            P.row_group[r] = P.b0 + t;
            P.row_srow[r] = P.static_cached ? P.srow[P.scls[i] - P.g0] : r;
            r++;
        }
    }
    __syncthreads();
    for (uint32_t t = tid; t < P.B; t += blockDim.x) P.task_row[t] = P.rowof[P.dcls[P.b0 + t - P.g0] - P.g0];
    if (tid == 0) *P.n_rows = total;
}

}  // namespace pe
