// kernels_common.cuh -- per-(group,node) predicates and the reservation, on
// the global-memory columns.  Used by the sequencer, the fit kernel and (for
// the rarely used filters) the scan kernel.  Every function cites the
// reference Check / addTask it restates.
#pragma once
#include "dev_types.h"

namespace pe {

__device__ __forceinline__ uint32_t tie_pos(uint32_t n, uint32_t tie_start, uint32_t N) {
    return n >= tie_start ? n - tie_start : n + N - tie_start;
}

// countRecentFailures (nodeinfo.go:206-221); the shim pre-counts, sorted by node.
__device__ __forceinline__ uint32_t fail_count(const TickDev &K, const pe_group &g, uint32_t n) {
    uint32_t lo = 0, hi = g.fail_cnt;
    const pe_node_fail *f = K.fails + g.fail_off;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        uint32_t v = f[mid].node_idx;
        if (v < n) lo = mid + 1; else hi = mid;
    }
    return (lo < g.fail_cnt && f[lo].node_idx == n) ? f[lo].count : 0u;
}

// Rank prefix of nodeLess (scheduler.go:708-735): (f >= 5 ? f : 0, svc, total).
__device__ __forceinline__ unsigned long long make_pref(uint32_t fails, uint32_t svc, uint32_t total) {
    uint32_t f5 = fails >= 5u ? (fails > 255u ? 255u : fails) : 0u;
    return ((unsigned long long)((f5 << 24) | (svc & 0xFFFFFFu)) << 32) | total;
}

// ResourceFilter.Check generic part, filter.go:85-90 + validate.go:24-51
__device__ __forceinline__ bool gen_enough(int64_t cell, int64_t want) {
    return (cell & 3) != PE_GEN_ABSENT && want <= (cell >> 2);
}

// nodeSet.tree's branch walk (nodeset.go:59-101) seen from one leaf of the preference tree: is the node under it?
// (pe_group.leaf_cnt terms follow the group's constraints; a group without preferences has none.)
__device__ __forceinline__ bool in_leaf(const DevTable &T, const TickDev &K, const pe_group &g, uint32_t n) {
    for (uint32_t i = 0; i < g.leaf_cnt; i++) {
        const pe_constraint c = K.cons[g.con_off + g.con_cnt + i];
        if (T.attr[c.col][n] != c.value) return false;
    }
    return true;
}

// Pipeline.Process (pipeline.go:56-68) for one node: 0 = every enabled filter
// passed, else 1 + index of the first failing filter.
__device__ __noinline__ uint32_t eval_ff(const DevTable &T, const TickDev &K, const pe_group &g, uint32_t n,
                                         uint32_t meta, uint32_t svc_n) {
    const uint32_t fm = g.filter_mask;
    // ReadyFilter.Check, filter.go:40-43
    if ((fm & (1u << PE_F_READY)) && !(meta & PE_NODE_READY)) return 1 + PE_F_READY;
    // ResourceFilter.Check, filter.go:76-93
    if (fm & (1u << PE_F_RESOURCE)) {
        if (g.cpu_res > T.cpu[n]) return 1 + PE_F_RESOURCE;
        if (g.mem_res > T.mem[n]) return 1 + PE_F_RESOURCE;
        for (uint32_t i = 0; i < g.gen_cnt; i++) {
            const pe_generic_want w = K.gens[g.gen_off + i];
            if (!gen_enough(T.gen[w.kind][n], w.value)) return 1 + PE_F_RESOURCE;
        }
    }
    // PluginFilter.Check, filter.go:141-183
    if ((fm & (1u << PE_F_PLUGIN)) && (meta & PE_NODE_HAS_ENGINE)) {
        for (uint32_t i = 0; i < g.plug_cnt; i++) {
            uint32_t s = K.plugs[g.plug_off + i];
            if (!((T.plug[s >> 5][n] >> (s & 31)) & 1u)) return 1 + PE_F_PLUGIN;
        }
        if (g.flags & PE_G_LOG_DRIVER) {
            uint32_t s = g.log_plugin;
            bool exists = (T.plug[s >> 5][n] >> (s & 31)) & 1u;
            if (!exists && (meta & PE_NODE_HAS_LOGPLUGIN)) return 1 + PE_F_PLUGIN;
        }
    }
    // ConstraintFilter.Check, filter.go:241-243 -> constraint.NodeMatches, constraint.go:107-207
    if (fm & (1u << PE_F_CONSTRAINT)) {
        if (g.flags & PE_G_CONSTRAINT_NEVER) return 1 + PE_F_CONSTRAINT;
        for (uint32_t i = 0; i < g.con_cnt; i++) {
            const pe_constraint c = K.cons[g.con_off + i];
            bool match = T.attr[c.col][n] == c.value;
            if (match == (c.neq != 0)) return 1 + PE_F_CONSTRAINT;
        }
        for (uint32_t i = 0; i < g.ip_cnt; i++) {
            const pe_ip_constraint c = K.ips[g.ip_off + i];
            bool hit = (meta & PE_NODE_IP_VALID) != 0;
            if (hit && c.is_cidr) hit = ((meta & PE_NODE_IP_V4) != 0) == (c.is_v4 != 0);
            if (hit) {
                uint4 a = T.ip[n];
                hit = (a.x & c.mask[0]) == c.net[0] && (a.y & c.mask[1]) == c.net[1] &&
                      (a.z & c.mask[2]) == c.net[2] && (a.w & c.mask[3]) == c.net[3];
            }
            if (hit == (c.neq != 0)) return 1 + PE_F_CONSTRAINT;
        }
    }
    // PlatformFilter.Check, filter.go:272-312
    if ((fm & (1u << PE_F_PLATFORM)) && g.plat_cnt) {
        bool ok = false;
        if (meta & PE_NODE_HAS_PLATFORM) {
            uint32_t os = (meta >> 8) & 0xFF, arch = (meta >> 16) & 0xFF;
            for (uint32_t i = 0; i < g.plat_cnt && !ok; i++) {
                const pe_platform p = K.plats[g.plat_off + i];
                ok = (p.arch_id == 0 || p.arch_id == arch) && (p.os_id == 0 || p.os_id == os);
            }
        }
        if (!ok) return 1 + PE_F_PLATFORM;
    }
    // HostPortFilter.Check, filter.go:342-353
    if (fm & (1u << PE_F_HOSTPORT)) {
        for (uint32_t i = 0; i < g.port_cnt; i++) {
            uint32_t s = K.ports[g.port_off + i];
            if ((T.ports[s >> 5][n] >> (s & 31)) & 1u) return 1 + PE_F_HOSTPORT;
        }
    }
    // MaxReplicasFilter.Check, filter.go:379-381
    if ((fm & (1u << PE_F_MAXREPLICAS)) && !((unsigned long long)svc_n < g.max_replicas)) return 1 + PE_F_MAXREPLICAS;
    return 0;
}

// genericresource.Claim reduced to counts (resource_management.go:11-72,
// helpers.go:58-111): new cell for the kind of want #i given the ORIGINAL cell.
// Handles a kind listed more than once (discrete amounts add up; named members
// are the same list prefix, so the longest selection wins).
__device__ __forceinline__ int64_t claim_cell(const TickDev &K, const pe_group &g, uint32_t i, int64_t cell) {
    const uint32_t kind = K.gens[g.gen_off + i].kind;
    const int type = (int)(cell & 3);
    const int64_t count = cell >> 2;
    if (type == PE_GEN_ABSENT) return cell;
    int64_t take = 0;
    for (uint32_t j = i; j < g.gen_cnt; j++) {
        const pe_generic_want x = K.gens[g.gen_off + j];
        if (x.kind != kind) continue;
        if (type == PE_GEN_DISCRETE) {
            if (count >= x.value && x.value != 0) take += x.value;   // selectNodeResources :52-57
        } else {
            int64_t sel = x.value == 0 ? count : (x.value < count ? x.value : count);  // :58-63
            take = take > sel ? take : sel;
        }
    }
    const int64_t left = count - take;
    return left <= 0 ? 0 : PE_GEN_ENCODE(left, type);                // dropped at <= 0, helpers.go:94-97
}

__device__ __forceinline__ bool gen_first_occurrence(const TickDev &K, const pe_group &g, uint32_t i) {
    const uint32_t kind = K.gens[g.gen_off + i].kind;
    for (uint32_t j = 0; j < i; j++)
        if (K.gens[g.gen_off + j].kind == kind) return false;
    return true;
}

// NodeInfo.addTask (new-task branch, nodeinfo.go:125-153) applied straight to
// the global columns by ONE thread.
__device__ __noinline__ void add_task_global(const DevTable &T, const TickDev &K, const pe_group &g, uint32_t n,
                                             bool counts, DevCounters *ctr) {
    T.mem[n] -= g.mem_res;
    T.cpu[n] -= g.cpu_res;
    for (uint32_t i = 0; i < g.gen_cnt; i++) {
        if (!gen_first_occurrence(K, g, i)) continue;
        int64_t *col = T.gen[K.gens[g.gen_off + i].kind];
        col[n] = claim_cell(K, g, i, col[n]);
    }
    for (uint32_t i = 0; i < g.port_cnt; i++) {
        uint32_t s = K.ports[g.port_off + i];
        T.ports[s >> 5][n] |= 1u << (s & 31);
    }
    if (counts) {
        T.total[n] += 1;
        uint32_t v = T.svc[g.svc_id][n] + 1;
        T.svc[g.svc_id][n] = v;
        if (v >= 0xFFFFFFu) atomicOr(&ctr->error, PE_DEV_ERR_SVC_OVERFLOW);
    }
}

}  // namespace pe
