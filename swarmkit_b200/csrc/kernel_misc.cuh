// kernel_misc.cuh -- node-mirror maintenance and taskFitNode.
//
//   k_upsert  <- Scheduler.createOrUpdateNode / buildNodeSet   scheduler.go:368-396,973-990
//   k_remove  <- nodeSet.remove                                nodeset.go:46-48
//   k_delta   <- NodeInfo.addTask / removeTask outside a tick  nodeinfo.go:66-154
//   k_fit     <- Scheduler.taskFitNode                         scheduler.go:646-690
#pragma once
#include "kernels_common.cuh"

namespace pe {

struct UpsertParams {
    DevTable T;
    const pe_node_row *rows;
    uint32_t n_rows;
    const pe_kv32 *attrs;
    const pe_kv64 *gens;
    const pe_kv32 *svcs;
    const uint32_t *ports;
    const uint32_t *plugs;
};

// One thread per row.  Rows of one call must name distinct nodes.
__global__ void k_upsert(const __grid_constant__ UpsertParams P) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_rows) return;
    const DevTable &T = P.T;
    const pe_node_row row = P.rows[r];
    const uint32_t n = row.node_idx;
    T.meta[n] = (row.flags & PE_META_FLAGS_MASK) | ((row.os_id & 0xFFu) << 8) | ((row.arch_id & 0xFFu) << 16);
    T.cpu[n] = row.cpu_avail;
    T.mem[n] = row.mem_avail;
    T.total[n] = row.total_tasks;
    T.ip[n] = make_uint4(row.ip[0], row.ip[1], row.ip[2], row.ip[3]);
    for (uint32_t c = 0; c < T.n_attr; c++) if (T.attr[c]) T.attr[c][n] = 0;
    for (uint32_t c = 0; c < T.n_gen; c++) if (T.gen[c]) T.gen[c][n] = 0;
    for (uint32_t c = 0; c < T.n_portw; c++) if (T.ports[c]) T.ports[c][n] = 0;
    for (uint32_t c = 0; c < T.n_plugw; c++) if (T.plug[c]) T.plug[c][n] = 0;
    for (uint32_t i = 0; i < row.attr_cnt; i++) { const pe_kv32 kv = P.attrs[row.attr_off + i]; T.attr[kv.key][n] = kv.value; }
    for (uint32_t i = 0; i < row.gen_cnt; i++) { const pe_kv64 kv = P.gens[row.gen_off + i]; T.gen[kv.key][n] = kv.value; }
    // the host prepends {svc, 0} entries for services this node used to have
    for (uint32_t i = 0; i < row.svc_cnt; i++) { const pe_kv32 kv = P.svcs[row.svc_off + i]; T.svc[kv.key][n] = kv.value; }
    for (uint32_t i = 0; i < row.port_cnt; i++) { const uint32_t s = P.ports[row.port_off + i]; T.ports[s >> 5][n] |= 1u << (s & 31u); }
    for (uint32_t i = 0; i < row.plug_cnt; i++) { const uint32_t s = P.plugs[row.plug_off + i]; T.plug[s >> 5][n] |= 1u << (s & 31u); }
}

// A row upsert replaces the node's whole ActiveTasksCountByService map.
__global__ void k_zero_svc_rows(DevTable T, const pe_node_row *rows, uint32_t n_rows) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const uint32_t n = rows[r].node_idx;
    for (uint32_t c = 0; c < T.n_svc; c++) if (T.svc[c]) T.svc[c][n] = 0;
}

__global__ void k_remove(DevTable T, const uint32_t *idx, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) T.meta[idx[i]] = 0;
}

// ---- placement preferences: the leaves of nodeSet.tree (nodeset.go:59-101) ------------------------------------
// Every node of the set hangs under the leaf named by its values of the preference labels; a leaf's `tasks` is the sum
// of ActiveTasksCountByService[service] over its nodes.  Leaves are found through an open-addressing table keyed by a
// 64-bit fingerprint of the value tuple; k_pref_emit then checks every node's tuple against its leaf's representative
// node, so a fingerprint collision is reported, never silently merged.
struct PrefParams {
    DevTable T;
    const uint32_t *svccol;                 // the service's counter column
    uint32_t cols[PE_MAX_PREF_LEVELS], n_levels;
    unsigned long long *keys;               // [mask + 1] 0 = empty
    uint32_t *sums, *reps;                  // [mask + 1] task sums; lowest node of the leaf (all ones at start)
    uint32_t mask;
    uint32_t *out_vals, *out_tasks, *out_n, *out_err;   // device staging of the result; out_err: 1 = collision
    uint32_t cap;
};
__device__ __forceinline__ unsigned long long pref_key(const PrefParams &P, uint32_t n) {
    unsigned long long h = 0x9E3779B97F4A7C15ull;
    for (uint32_t l = 0; l < P.n_levels; l++) {
        h ^= P.T.attr[P.cols[l]][n] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
        h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33;
    }
    return h ? h : 1ull;
}
__global__ void __launch_bounds__(256) k_pref_leaves(const PrefParams P) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < P.T.n_nodes; n += gridDim.x * blockDim.x) {
        if (!(P.T.meta[n] & PE_NODE_VALID)) continue;
        const unsigned long long key = pref_key(P, n);
        for (uint32_t h = (uint32_t)(key >> 20) & P.mask;; h = (h + 1u) & P.mask) {
            unsigned long long cur = P.keys[h];
            if (cur == 0ull) cur = atomicCAS(&P.keys[h], 0ull, key), cur = cur == 0ull ? key : cur;
            if (cur == key) { atomicAdd(&P.sums[h], P.svccol[n]); atomicMin(&P.reps[h], n); break; }
        }
    }
}
__global__ void __launch_bounds__(256) k_pref_emit(const PrefParams P) {
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t n = t0; n < P.T.n_nodes; n += stride) {          // exactness: the node's tuple IS its leaf's tuple
        if (!(P.T.meta[n] & PE_NODE_VALID)) continue;
        const unsigned long long key = pref_key(P, n);
        uint32_t h = (uint32_t)(key >> 20) & P.mask;
        while (P.keys[h] != key) h = (h + 1u) & P.mask;
        const uint32_t r = P.reps[h];
        for (uint32_t l = 0; l < P.n_levels; l++) if (P.T.attr[P.cols[l]][n] != P.T.attr[P.cols[l]][r]) atomicOr(P.out_err, 1u);
    }
    for (uint32_t h = t0; h <= P.mask; h += stride) {
        if (P.keys[h] == 0ull) continue;
        const uint32_t i = atomicAdd(P.out_n, 1u);
        if (i >= P.cap) continue;
        const uint32_t r = P.reps[h];
        for (uint32_t l = 0; l < P.n_levels; l++) P.out_vals[(size_t)i * P.n_levels + l] = P.T.attr[P.cols[l]][r];
        P.out_tasks[i] = P.sums[h];
    }
}

// Commutative part of the deltas: one thread each, atomics.
__global__ void k_delta_add(DevTable T, const pe_task_delta *d, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const pe_task_delta x = d[i];
    const long long sg = x.sign >= 0 ? 1 : -1;
    atomicAdd(reinterpret_cast<unsigned long long *>(&T.cpu[x.node_idx]), (unsigned long long)(-sg * x.cpu));
    atomicAdd(reinterpret_cast<unsigned long long *>(&T.mem[x.node_idx]), (unsigned long long)(-sg * x.mem));
    if (x.counts) {
        atomicAdd(&T.total[x.node_idx], (uint32_t)sg);
        atomicAdd(&T.svc[x.svc_id][x.node_idx], (uint32_t)sg);
    }
}
// Order-dependent part (cell overwrite, port set/clear): one thread, in call order.
__global__ void k_delta_seq(DevTable T, const pe_task_delta *d, uint32_t n, const pe_kv64 *gens, const uint32_t *ports) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (uint32_t i = 0; i < n; i++) {
        const pe_task_delta x = d[i];
        for (uint32_t j = 0; j < x.gen_cnt; j++) T.gen[gens[x.gen_off + j].key][x.node_idx] = gens[x.gen_off + j].value;
        for (uint32_t j = 0; j < x.port_cnt; j++) {
            const uint32_t s = ports[x.port_off + j];
            if (x.sign >= 0) T.ports[s >> 5][x.node_idx] |= 1u << (s & 31u);
            else T.ports[s >> 5][x.node_idx] &= ~(1u << (s & 31u));
        }
    }
}

// taskFitNode: pipeline on the one named node, then reserve.  Requests are
// dependent when they name the same node, so one thread walks them in order.
__global__ void k_fit(DevTable T, TickDev K, uint32_t n_groups, const uint32_t *node_idx, uint8_t *out_ok, DevCounters *ctr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (uint32_t gi = 0; gi < n_groups; gi++) {
        const pe_group g = K.groups[gi];
        uint32_t *of = K.out_fail + (size_t)gi * PE_NUM_FILTERS;
        for (int f = 0; f < PE_NUM_FILTERS; f++) of[f] = 0;
        const uint32_t n = node_idx[gi];
        if (n >= T.n_nodes || !(T.meta[n] & PE_NODE_VALID)) { out_ok[gi] = 2; continue; }  // scheduler.go:648-651
        const uint32_t ff = eval_ff(T, K, g, n, T.meta[n], T.svc[g.svc_id][n]);
        if (ff) { of[ff - 1] = 1; out_ok[gi] = 0; continue; }                              // :654-660
        const bool counts = g.n_tasks ? (K.task_flags[g.task_off] & PE_T_COUNTS) != 0 : true;
        add_task_global(T, K, g, n, counts, ctr);                                          // :686-688
        ctr->placements++;
        out_ok[gi] = 1;
    }
}

}  // namespace pe
