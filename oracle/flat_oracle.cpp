// flat_oracle.cpp -- CPU oracle for the placement hot path, encoded (integer) form.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference arm may load this library.  The
// product (swarmkit_b200/csrc) never links, loads or calls it.
//
// What it is: a single-threaded restatement of the reference scheduler's group
// loop on the same dictionary-encoded inputs the CUDA engine receives, behind
// the same C ABI (include/placement_engine.h) with the prefix ope_ instead of
// pe_.  It follows the reference control flow literally, with the canonical
// orders of SURVEY.md 8(c) / Appendix A substituted for Go map order:
//
//   schedule_group   <- Scheduler.scheduleTaskGroup      manager/scheduler/scheduler.go:694-748
//   tree build       <- nodeSet.tree                      manager/scheduler/nodeset.go:50-124
//   node_less        <- nodeLess closure                  manager/scheduler/scheduler.go:708-735
//   process          <- Pipeline.Process                  manager/scheduler/pipeline.go:56-68
//   check_*          <- the Check methods                 manager/scheduler/filter.go:40,76,141,241,272,342,379
//   ordered_nodes    <- decisionTree.orderedNodes         manager/scheduler/decision_tree.go:24-52
//   place_on_nodes   <- Scheduler.scheduleNTasksOnNodes   manager/scheduler/scheduler.go:844-924
//   add_task         <- NodeInfo.addTask + genericresource.Claim
//                                                         manager/scheduler/nodeinfo.go:108-154,
//                                                         api/genericresource/resource_management.go:11-72
//   fit              <- Scheduler.taskFitNode             manager/scheduler/scheduler.go:646-690
//
// Pinning: this encoded oracle is cross-checked against the object-level oracle
// (oracle/sched_oracle.cpp, which is pinned by the reference's own known-answer
// tests) in tests/test_shim_cpu.py; see DESIGN.md "Oracle".
//
// Placement preferences: the flat ABI carries ONE LEAF VISIT of the decision tree per group (pe_group.leaf_cnt: the
// node set is cut down to the leaf, nodeset.go:59-101) plus pref_leaves (the tree's per-leaf task sums); the tree walk
// of scheduleNTasksOnSubtree (scheduler.go:772-825) is host code above the ABI.  The object-level oracle implements
// the whole of it literally and pins this decomposition (tests/test_shim_cpu.py).

#include "../include/placement_engine.h"

#include <algorithm>
#include <map>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct Node {
    uint32_t flags = 0, os = 0, arch = 0, total = 0;
    int64_t cpu = 0, mem = 0;
    uint32_t ip[4] = {0, 0, 0, 0};
};

struct Oracle {
    std::vector<Node> nodes;
    uint32_t n_nodes = 0;
    std::vector<std::vector<uint32_t>> attr;   // [col][node]
    std::vector<std::vector<int64_t>> gen;     // [kind][node]
    std::vector<std::vector<uint32_t>> svc;    // [service][node]
    std::vector<std::vector<uint32_t>> ports;  // [word][node]
    std::vector<std::vector<uint32_t>> plug;   // [word][node]
    std::string err;
    pe_stats stats{};
    // staged tick
    std::vector<pe_group> t_groups;
    std::vector<uint8_t> t_flags;
    std::vector<pe_generic_want> t_gens;
    std::vector<pe_constraint> t_cons;
    std::vector<pe_ip_constraint> t_ips;
    std::vector<pe_platform> t_plats;
    std::vector<uint32_t> t_ports, t_plugs;
    std::vector<pe_node_fail> t_fails;
    std::vector<uint32_t> r_node, r_fail;

    void ensure_rows(uint32_t n) {
        if (n <= nodes.size()) return;
        nodes.resize(n);
        for (auto &c : attr) c.resize(n, 0);
        for (auto &c : gen) c.resize(n, 0);
        for (auto &c : svc) c.resize(n, 0);
        for (auto &c : ports) c.resize(n, 0);
        for (auto &c : plug) c.resize(n, 0);
    }
    template <class T> std::vector<T> &col(std::vector<std::vector<T>> &tab, uint32_t i) {
        if (i >= tab.size()) tab.resize(i + 1);
        if (tab[i].size() < nodes.size()) tab[i].resize(nodes.size(), 0);
        return tab[i];
    }
};

static std::string g_create_err;

// View of one tick's side arrays.
struct Tick {
    const pe_group *groups; uint32_t n_groups;
    const uint8_t *task_flags; uint32_t n_tasks;
    const pe_generic_want *gens;
    const pe_constraint *cons;
    const pe_ip_constraint *ips;
    const pe_platform *plats;
    const uint32_t *ports;
    const uint32_t *plugs;
    const pe_node_fail *fails;
};

struct Sched {
    Oracle &o;
    const Tick &t;
    const pe_group &g;
    uint32_t failcnt[PE_NUM_FILTERS];
    uint32_t N;
    const uint32_t *svc_col;

    Sched(Oracle &o_, const Tick &t_, const pe_group &g_) : o(o_), t(t_), g(g_) {
        std::memset(failcnt, 0, sizeof failcnt);  // pipeline.go:76-81 SetTask
        N = o.n_nodes;
        svc_col = o.col(o.svc, g.svc_id).data();
    }

    uint32_t pos(uint32_t n) const { return n >= g.tie_start ? n - g.tie_start : n + N - g.tie_start; }

    // nodeSet.tree's branch walk, nodeset.go:59-101, seen from one leaf: is the node under it?
    bool in_leaf(uint32_t n) const {
        for (uint32_t i = 0; i < g.leaf_cnt; i++) {
            const pe_constraint &c = t.cons[g.con_off + g.con_cnt + i];
            uint32_t v = c.col < o.attr.size() && n < o.attr[c.col].size() ? o.attr[c.col][n] : 0;
            if (v != c.value) return false;
        }
        return true;
    }

    // countRecentFailures, nodeinfo.go:206-221 (the shim pre-counts per node)
    uint32_t failures(uint32_t n) const {
        if (g.fail_cnt == 0) return 0;
        const pe_node_fail *b = t.fails + g.fail_off, *e = b + g.fail_cnt;
        const pe_node_fail *it = std::lower_bound(b, e, n, [](const pe_node_fail &f, uint32_t v) { return f.node_idx < v; });
        return (it != e && it->node_idx == n) ? it->count : 0;
    }

    // nodeLess, scheduler.go:708-735.  strict=true is the reference comparator;
    // strict=false appends the canonical tie-break.
    bool node_less(uint32_t a, uint32_t b, bool strict) const {
        uint32_t fa = failures(a), fb = failures(b);
        if (fa >= 5 || fb >= 5) {           // maxFailures, scheduler.go:23
            if (fa > fb) return false;
            if (fb > fa) return true;
        }
        uint32_t sa = svc_col[a], sb = svc_col[b];
        if (sa < sb) return true;
        if (sa > sb) return false;
        uint32_t ta = o.nodes[a].total, tb = o.nodes[b].total;
        if (strict) return ta < tb;
        if (ta != tb) return ta < tb;
        return pos(a) < pos(b);
    }

    // ---- the Check methods ------------------------------------------------
    bool check_ready(uint32_t n) const { return o.nodes[n].flags & PE_NODE_READY; }  // filter.go:40-43

    bool check_resource(uint32_t n) const {  // filter.go:76-93 + validate.go:24-51
        const Node &nd = o.nodes[n];
        if (g.cpu_res > nd.cpu) return false;
        if (g.mem_res > nd.mem) return false;
        for (uint32_t i = 0; i < g.gen_cnt; i++) {
            const pe_generic_want &w = t.gens[g.gen_off + i];
            int64_t cell = w.kind < o.gen.size() && n < o.gen[w.kind].size() ? o.gen[w.kind][n] : 0;
            if ((cell & 3) == PE_GEN_ABSENT) return false;   // len(nrs) == 0
            if (w.value > (cell >> 2)) return false;         // discrete value / named member count
        }
        return true;
    }

    bool has_plug(uint32_t n, uint32_t slot) const {
        uint32_t w = slot >> 5;
        return w < o.plug.size() && n < o.plug[w].size() && ((o.plug[w][n] >> (slot & 31)) & 1);
    }
    bool check_plugin(uint32_t n) const {  // filter.go:141-183
        const Node &nd = o.nodes[n];
        if (!(nd.flags & PE_NODE_HAS_ENGINE)) return true;
        for (uint32_t i = 0; i < g.plug_cnt; i++)
            if (!has_plug(n, t.plugs[g.plug_off + i])) return false;
        if (g.flags & PE_G_LOG_DRIVER)
            if (!has_plug(n, g.log_plugin) && (nd.flags & PE_NODE_HAS_LOGPLUGIN)) return false;
        return true;
    }

    bool check_constraint(uint32_t n) const {  // filter.go:241-243 + constraint.go:107-207
        if (g.flags & PE_G_CONSTRAINT_NEVER) return false;
        for (uint32_t i = 0; i < g.con_cnt; i++) {
            const pe_constraint &c = t.cons[g.con_off + i];
            uint32_t v = c.col < o.attr.size() && n < o.attr[c.col].size() ? o.attr[c.col][n] : 0;
            bool match = v == c.value;            // strings.EqualFold on folded ids
            if (match == (c.neq != 0)) return false;
        }
        const Node &nd = o.nodes[n];
        for (uint32_t i = 0; i < g.ip_cnt; i++) {  // constraint.go:127-146
            const pe_ip_constraint &c = t.ips[g.ip_off + i];
            bool hit = (nd.flags & PE_NODE_IP_VALID) != 0;
            if (hit && c.is_cidr) hit = ((nd.flags & PE_NODE_IP_V4) != 0) == (c.is_v4 != 0);
            for (int k = 0; k < 4 && hit; k++) hit = (nd.ip[k] & c.mask[k]) == c.net[k];
            if (hit == (c.neq != 0)) return false;
        }
        return true;
    }

    bool check_platform(uint32_t n) const {  // filter.go:272-312
        if (g.plat_cnt == 0) return true;
        const Node &nd = o.nodes[n];
        if (!(nd.flags & PE_NODE_HAS_PLATFORM)) return false;
        for (uint32_t i = 0; i < g.plat_cnt; i++) {
            const pe_platform &p = t.plats[g.plat_off + i];
            if ((p.arch_id == 0 || p.arch_id == nd.arch) && (p.os_id == 0 || p.os_id == nd.os)) return true;
        }
        return false;
    }

    bool port_used(uint32_t n, uint32_t slot) const {
        uint32_t w = slot >> 5;
        return w < o.ports.size() && n < o.ports[w].size() && ((o.ports[w][n] >> (slot & 31)) & 1);
    }
    bool check_hostport(uint32_t n) const {  // filter.go:342-353
        for (uint32_t i = 0; i < g.port_cnt; i++)
            if (port_used(n, t.ports[g.port_off + i])) return false;
        return true;
    }

    bool check_maxreplicas(uint32_t n) const {  // filter.go:379-381
        return (uint64_t)svc_col[n] < g.max_replicas;
    }

    // Pipeline.Process, pipeline.go:56-68
    bool process(uint32_t n) {
        for (int f = 0; f < PE_NUM_FILTERS; f++) {
            if (!((g.filter_mask >> f) & 1)) continue;
            bool ok = true;
            switch (f) {
                case PE_F_READY: ok = check_ready(n); break;
                case PE_F_RESOURCE: ok = check_resource(n); break;
                case PE_F_PLUGIN: ok = check_plugin(n); break;
                case PE_F_CONSTRAINT: ok = check_constraint(n); break;
                case PE_F_PLATFORM: ok = check_platform(n); break;
                case PE_F_HOSTPORT: ok = check_hostport(n); break;
                case PE_F_MAXREPLICAS: ok = check_maxreplicas(n); break;
                default: ok = true; break;  // Volumes: evaluated by the host shim
            }
            if (!ok) { failcnt[f]++; return false; }
        }
        std::memset(failcnt, 0, sizeof failcnt);
        return true;
    }

    // NodeInfo.addTask (new task branch), nodeinfo.go:125-153, with
    // genericresource.Claim reduced to per-kind counts (SURVEY hard part D).
    void add_task(uint32_t n, bool counts) {
        Node &nd = o.nodes[n];
        nd.mem -= g.mem_res;
        nd.cpu -= g.cpu_res;
        for (uint32_t i = 0; i < g.gen_cnt; i++) {
            const pe_generic_want &w = t.gens[g.gen_off + i];
            bool seen = false;
            for (uint32_t j = 0; j < i; j++) seen |= t.gens[g.gen_off + j].kind == w.kind;
            if (seen) continue;  // aggregated with the first occurrence of the kind
            int64_t &cell = o.col(o.gen, w.kind)[n];
            int type = (int)(cell & 3);
            int64_t count = cell >> 2;
            if (type == PE_GEN_ABSENT) continue;
            int64_t take = 0;
            for (uint32_t j = i; j < g.gen_cnt; j++) {
                const pe_generic_want &x = t.gens[g.gen_off + j];
                if (x.kind != w.kind) continue;
                if (type == PE_GEN_DISCRETE) {
                    // selectNodeResources, resource_management.go:52-57
                    if (count >= x.value && x.value != 0) take += x.value;
                } else {
                    // named: first x.value members; x.value == 0 never hits the
                    // `len(nrs) == tr.Value` exit and selects them all (:58-63)
                    int64_t sel = x.value == 0 ? count : std::min(x.value, count);
                    take = std::max(take, sel);
                }
            }
            int64_t left = count - take;
            // ConsumeNodeResources/remove, helpers.go:58-111: an entry at <= 0 is dropped
            cell = left <= 0 ? 0 : PE_GEN_ENCODE(left, type);
        }
        for (uint32_t i = 0; i < g.port_cnt; i++) {  // nodeinfo.go:139-146
            uint32_t slot = t.ports[g.port_off + i];
            o.col(o.ports, slot >> 5)[n] |= 1u << (slot & 31);
        }
        if (counts) {  // nodeinfo.go:148-151
            nd.total++;
            o.col(o.svc, g.svc_id)[n]++;
            svc_col = o.svc[g.svc_id].data();
        }
    }
};

// Scheduler.scheduleTaskGroup, scheduler.go:694-748 (no preferences => single leaf)
static void schedule_group(Oracle &o, const Tick &t, uint32_t gi, uint32_t *out_node, uint32_t *out_fail) {
    const pe_group &g = t.groups[gi];
    Sched s(o, t, g);
    uint32_t k = g.n_tasks, N = o.n_nodes;
    for (uint32_t i = 0; i < k; i++) out_node[g.task_off + i] = PE_NONE;
    if (k == 0) { std::memset(out_fail + (size_t)gi * PE_NUM_FILTERS, 0, PE_NUM_FILTERS * 4); return; }

    // ---- nodeSet.tree, nodeset.go:107-120: bounded max-heap of the k best feasible nodes
    std::vector<uint32_t> heap;
    heap.reserve(std::min<uint64_t>(k, N));
    auto worse = [&](uint32_t a, uint32_t b) { return s.node_less(a, b, false); };  // max-heap on less
    for (uint32_t p = 0; p < N; p++) {
        uint32_t n = p + g.tie_start; if (n >= N) n -= N;   // canonical evaluation order
        if (!(o.nodes[n].flags & PE_NODE_VALID)) continue;  // not in nodeSet.nodes
        if (g.leaf_cnt && !s.in_leaf(n)) continue;          // under another leaf of the preference tree
        o.stats.evals_generic++;
        if (heap.size() < k) {
            if (s.process(n)) { heap.push_back(n); std::push_heap(heap.begin(), heap.end(), worse); }
        } else if (s.node_less(n, heap.front(), false)) {
            if (s.process(n)) {
                std::pop_heap(heap.begin(), heap.end(), worse);
                heap.back() = n;
                std::push_heap(heap.begin(), heap.end(), worse);
            }
        }
    }
    // ---- decisionTree.orderedNodes, decision_tree.go:24-52: best -> worst
    std::sort_heap(heap.begin(), heap.end(), worse);
    std::vector<uint32_t> &nodes = heap;

    // ---- scheduleNTasksOnNodes, scheduler.go:844-924
    uint32_t done = 0;
    if (!nodes.empty()) {
        uint32_t m = (uint32_t)nodes.size();
        std::vector<uint8_t> failed(m, 0);
        uint64_t it = 0;
        for (uint32_t ti = 0; ti < k; ti++) {
            uint32_t i = (uint32_t)(it % m);
            uint32_t n = nodes[i];
            out_node[g.task_off + ti] = n;
            s.add_task(n, (t.task_flags[g.task_off + ti] & PE_T_COUNTS) != 0);
            o.stats.placements++;
            done++;
            if (done == k) break;
            if (it + 1 < m) {
                if (s.node_less(nodes[(it + 1) % m], n, true)) it++;   // :899-905, strict nodeLess
            } else {
                it++;                                                  // :906-910
            }
            uint64_t start = it;
            bool dead = false;
            while (failed[it % m] || !s.process(nodes[it % m])) {       // :912-920
                failed[it % m] = 1;
                it++;
                if (it - start == m) { dead = true; break; }
            }
            if (dead) break;
        }
    }
    std::memcpy(out_fail + (size_t)gi * PE_NUM_FILTERS, s.failcnt, sizeof s.failcnt);
    (void)done;
}

static Tick view(const pe_tick *tk) {
    Tick t{};
    t.groups = tk->groups; t.n_groups = tk->n_groups;
    t.task_flags = tk->task_flags; t.n_tasks = tk->n_tasks;
    t.gens = tk->gens; t.cons = tk->cons; t.ips = tk->ips; t.plats = tk->plats;
    t.ports = tk->ports; t.plugs = tk->plugs; t.fails = tk->fails;
    return t;
}

static int32_t validate(Oracle &o, const pe_tick *tk) {
    if (!tk) { o.err = "null tick"; return PE_ERR_INVALID; }
    uint64_t tasks = 0;
    for (uint32_t i = 0; i < tk->n_groups; i++) {
        const pe_group &g = tk->groups[i];
        if ((uint64_t)g.task_off + g.n_tasks > tk->n_tasks) { o.err = "group task range out of bounds"; return PE_ERR_INVALID; }
        if ((uint64_t)g.gen_off + g.gen_cnt > tk->n_gens || (uint64_t)g.con_off + g.con_cnt + g.leaf_cnt > tk->n_cons ||
            (uint64_t)g.ip_off + g.ip_cnt > tk->n_ips || (uint64_t)g.plat_off + g.plat_cnt > tk->n_plats ||
            (uint64_t)g.port_off + g.port_cnt > tk->n_ports || (uint64_t)g.plug_off + g.plug_cnt > tk->n_plugs ||
            (uint64_t)g.fail_off + g.fail_cnt > tk->n_fails) { o.err = "group side-array range out of bounds"; return PE_ERR_INVALID; }
        if (o.n_nodes && g.tie_start >= o.n_nodes) { o.err = "tie_start >= node count"; return PE_ERR_INVALID; }
        tasks += g.n_tasks;
    }
    (void)tasks;
    return PE_OK;
}

}  // namespace

struct pe_engine { Oracle o; };

extern "C" {

uint32_t ope_abi_version(void) { return PE_ABI_VERSION; }

int32_t ope_create(const pe_config *cfg, pe_engine **out) {
    if (!cfg || !out) { g_create_err = "null argument"; return PE_ERR_INVALID; }
    if (cfg->abi_version != PE_ABI_VERSION) { g_create_err = "ABI version mismatch"; return PE_ERR_INVALID; }
    pe_engine *h = new pe_engine();
    h->o.ensure_rows(cfg->node_capacity);
    *out = h;
    return PE_OK;
}

void ope_destroy(pe_engine *h) { delete h; }

const char *ope_last_error(const pe_engine *h) { return h ? h->o.err.c_str() : g_create_err.c_str(); }

int32_t ope_set_node_count(pe_engine *h, uint32_t n) {
    h->o.ensure_rows(n);
    h->o.n_nodes = n;
    return PE_OK;
}

// createOrUpdateNode, scheduler.go:368-396: the shim recomputes the row, we store it
int32_t ope_node_upsert(pe_engine *h, const pe_node_row *rows, uint32_t n_rows, const pe_kv32 *attrs,
                        const pe_kv64 *gens, const pe_kv32 *svcs, const uint32_t *ports, const uint32_t *plugs) {
    Oracle &o = h->o;
    for (uint32_t r = 0; r < n_rows; r++) {
        const pe_node_row &row = rows[r];
        uint32_t n = row.node_idx;
        if (row.os_id > 255 || row.arch_id > 255) { o.err = "os/arch id > 255"; return PE_ERR_INVALID; }
        o.ensure_rows(n + 1);
        if (n >= o.n_nodes) o.n_nodes = n + 1;
        Node &nd = o.nodes[n];
        nd.flags = row.flags; nd.os = row.os_id; nd.arch = row.arch_id;
        nd.cpu = row.cpu_avail; nd.mem = row.mem_avail; nd.total = row.total_tasks;
        std::memcpy(nd.ip, row.ip, sizeof nd.ip);
        for (auto &c : o.attr) if (n < c.size()) c[n] = 0;
        for (auto &c : o.gen) if (n < c.size()) c[n] = 0;
        for (auto &c : o.svc) if (n < c.size()) c[n] = 0;
        for (auto &c : o.ports) if (n < c.size()) c[n] = 0;
        for (auto &c : o.plug) if (n < c.size()) c[n] = 0;
        for (uint32_t i = 0; i < row.attr_cnt; i++) o.col(o.attr, attrs[row.attr_off + i].key)[n] = attrs[row.attr_off + i].value;
        for (uint32_t i = 0; i < row.gen_cnt; i++) o.col(o.gen, gens[row.gen_off + i].key)[n] = gens[row.gen_off + i].value;
        for (uint32_t i = 0; i < row.svc_cnt; i++) o.col(o.svc, svcs[row.svc_off + i].key)[n] = svcs[row.svc_off + i].value;
        for (uint32_t i = 0; i < row.port_cnt; i++) { uint32_t s = ports[row.port_off + i]; o.col(o.ports, s >> 5)[n] |= 1u << (s & 31); }
        for (uint32_t i = 0; i < row.plug_cnt; i++) { uint32_t s = plugs[row.plug_off + i]; o.col(o.plug, s >> 5)[n] |= 1u << (s & 31); }
    }
    return PE_OK;
}

// nodeSet.remove, nodeset.go:46-48
int32_t ope_node_remove(pe_engine *h, const uint32_t *idx, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        if (idx[i] >= h->o.nodes.size()) { h->o.err = "node index out of range"; return PE_ERR_INVALID; }
        h->o.nodes[idx[i]].flags = 0;
    }
    return PE_OK;
}

// addTask / removeTask outside a tick, nodeinfo.go:66-154
int32_t ope_node_task_delta(pe_engine *h, const pe_task_delta *d, uint32_t n, const pe_kv64 *gens, const uint32_t *ports) {
    Oracle &o = h->o;
    for (uint32_t i = 0; i < n; i++) {
        const pe_task_delta &x = d[i];
        if (x.node_idx >= o.nodes.size()) { o.err = "node index out of range"; return PE_ERR_INVALID; }
        Node &nd = o.nodes[x.node_idx];
        int64_t sg = x.sign >= 0 ? 1 : -1;
        nd.cpu -= sg * x.cpu;
        nd.mem -= sg * x.mem;
        if (x.counts) {
            nd.total += (uint32_t)sg;
            o.col(o.svc, x.svc_id)[x.node_idx] += (uint32_t)sg;
        }
        for (uint32_t j = 0; j < x.gen_cnt; j++) o.col(o.gen, gens[x.gen_off + j].key)[x.node_idx] = gens[x.gen_off + j].value;
        for (uint32_t j = 0; j < x.port_cnt; j++) {
            uint32_t s = ports[x.port_off + j];
            uint32_t &w = o.col(o.ports, s >> 5)[x.node_idx];
            if (sg > 0) w |= 1u << (s & 31); else w &= ~(1u << (s & 31));
        }
    }
    return PE_OK;
}

// nodeSet.tree, nodeset.go:59-101: the leaves and their `tasks` sums (every node of the set, feasible or not)
int32_t ope_pref_leaves(pe_engine *h, uint32_t svc_id, const uint32_t *cols, uint32_t n_levels, uint32_t cap,
                        uint32_t *out_vals, uint32_t *out_tasks, uint32_t *out_n_leaves) {
    Oracle &o = h->o;
    if (n_levels > PE_MAX_PREF_LEVELS || !out_n_leaves || (n_levels && !cols)) { o.err = "pref_leaves: bad arguments"; return PE_ERR_INVALID; }
    std::map<std::vector<uint32_t>, uint32_t> leaves;
    const std::vector<uint32_t> &sc = o.col(o.svc, svc_id);
    for (uint32_t n = 0; n < o.n_nodes; n++) {
        if (!(o.nodes[n].flags & PE_NODE_VALID)) continue;
        std::vector<uint32_t> key(n_levels);
        for (uint32_t l = 0; l < n_levels; l++) key[l] = cols[l] < o.attr.size() && n < o.attr[cols[l]].size() ? o.attr[cols[l]][n] : 0;
        leaves[key] += sc[n];
    }
    if (leaves.size() > cap) { o.err = "pref_leaves: more leaves than the caller's buffers hold"; return PE_ERR_OVERFLOW; }
    uint32_t i = 0;
    for (auto &kv : leaves) {
        for (uint32_t l = 0; l < n_levels; l++) out_vals[(size_t)i * n_levels + l] = kv.first[l];
        out_tasks[i++] = kv.second;
    }
    *out_n_leaves = i;
    return PE_OK;
}

// The node-attribute filters of every group against every node (constraint.NodeMatches per (service, node) when only the
// constraint filter is enabled): constraint_enforcer.go:65-226, global.go:306,440,513
int32_t ope_match_matrix(pe_engine *h, const pe_tick *tk, uint32_t *out_bits) {
    Oracle &o = h->o;
    int32_t rc = validate(o, tk);
    if (rc) return rc;
    Tick t = view(tk);
    const size_t words = (o.n_nodes + 31) / 32;
    std::memset(out_bits, 0, words * 4 * tk->n_groups);
    for (uint32_t g = 0; g < tk->n_groups; g++) {
        Sched s(o, t, t.groups[g]);
        const uint32_t fm = t.groups[g].filter_mask;
        for (uint32_t n = 0; n < o.n_nodes; n++) {
            if (!(o.nodes[n].flags & PE_NODE_VALID)) continue;
            if (t.groups[g].leaf_cnt && !s.in_leaf(n)) continue;
            bool ok = true;
            if ((fm >> PE_F_READY) & 1) ok = ok && s.check_ready(n);
            if ((fm >> PE_F_PLUGIN) & 1) ok = ok && s.check_plugin(n);
            if ((fm >> PE_F_CONSTRAINT) & 1) ok = ok && s.check_constraint(n);
            if ((fm >> PE_F_PLATFORM) & 1) ok = ok && s.check_platform(n);
            if (ok) out_bits[g * words + (n >> 5)] |= 1u << (n & 31);
        }
    }
    return PE_OK;
}

int32_t ope_schedule(pe_engine *h, const pe_tick *tk, uint32_t *out_node, uint32_t *out_fail) {
    Oracle &o = h->o;
    int32_t rc = validate(o, tk);
    if (rc) return rc;
    Tick t = view(tk);
    for (uint32_t g = 0; g < tk->n_groups; g++) schedule_group(o, t, g, out_node, out_fail);  // scheduler.go:464-469
    return PE_OK;
}

int32_t ope_tick_upload(pe_engine *h, const pe_tick *tk) {
    Oracle &o = h->o;
    int32_t rc = validate(o, tk);
    if (rc) return rc;
    o.t_groups.assign(tk->groups, tk->groups + tk->n_groups);
    o.t_flags.assign(tk->task_flags, tk->task_flags + tk->n_tasks);
    o.t_gens.assign(tk->gens, tk->gens + tk->n_gens);
    o.t_cons.assign(tk->cons, tk->cons + tk->n_cons);
    o.t_ips.assign(tk->ips, tk->ips + tk->n_ips);
    o.t_plats.assign(tk->plats, tk->plats + tk->n_plats);
    o.t_ports.assign(tk->ports, tk->ports + tk->n_ports);
    o.t_plugs.assign(tk->plugs, tk->plugs + tk->n_plugs);
    o.t_fails.assign(tk->fails, tk->fails + tk->n_fails);
    return PE_OK;
}

int32_t ope_tick_run(pe_engine *h) {
    Oracle &o = h->o;
    pe_tick tk{};
    tk.groups = o.t_groups.data(); tk.n_groups = (uint32_t)o.t_groups.size();
    tk.task_flags = o.t_flags.data(); tk.n_tasks = (uint32_t)o.t_flags.size();
    tk.gens = o.t_gens.data(); tk.n_gens = (uint32_t)o.t_gens.size();
    tk.cons = o.t_cons.data(); tk.n_cons = (uint32_t)o.t_cons.size();
    tk.ips = o.t_ips.data(); tk.n_ips = (uint32_t)o.t_ips.size();
    tk.plats = o.t_plats.data(); tk.n_plats = (uint32_t)o.t_plats.size();
    tk.ports = o.t_ports.data(); tk.n_ports = (uint32_t)o.t_ports.size();
    tk.plugs = o.t_plugs.data(); tk.n_plugs = (uint32_t)o.t_plugs.size();
    tk.fails = o.t_fails.data(); tk.n_fails = (uint32_t)o.t_fails.size();
    o.r_node.assign(tk.n_tasks, PE_NONE);
    o.r_fail.assign((size_t)tk.n_groups * PE_NUM_FILTERS, 0);
    return ope_schedule(h, &tk, o.r_node.data(), o.r_fail.data());
}

int32_t ope_tick_download(pe_engine *h, uint32_t *out_node, uint32_t *out_fail) {
    Oracle &o = h->o;
    if (out_node) std::memcpy(out_node, o.r_node.data(), o.r_node.size() * 4);
    if (out_fail) std::memcpy(out_fail, o.r_fail.data(), o.r_fail.size() * 4);
    return PE_OK;
}

// taskFitNode, scheduler.go:646-690
int32_t ope_fit(pe_engine *h, const pe_tick *tk, const uint32_t *node_idx, uint8_t *out_ok, uint32_t *out_fail) {
    Oracle &o = h->o;
    int32_t rc = validate(o, tk);
    if (rc) return rc;
    Tick t = view(tk);
    for (uint32_t gi = 0; gi < tk->n_groups; gi++) {
        const pe_group &g = t.groups[gi];
        uint32_t n = node_idx[gi];
        std::memset(out_fail + (size_t)gi * PE_NUM_FILTERS, 0, PE_NUM_FILTERS * 4);
        if (n >= o.n_nodes || !(o.nodes[n].flags & PE_NODE_VALID)) { out_ok[gi] = 2; continue; }  // :648-651
        Sched s(o, t, g);
        if (!s.process(n)) {                                                                       // :654-660
            out_ok[gi] = 0;
            std::memcpy(out_fail + (size_t)gi * PE_NUM_FILTERS, s.failcnt, sizeof s.failcnt);
            continue;
        }
        bool counts = g.n_tasks ? (t.task_flags[g.task_off] & PE_T_COUNTS) != 0 : true;
        s.add_task(n, counts);                                                                     // :686-688
        out_ok[gi] = 1;
    }
    return PE_OK;
}

int32_t ope_snapshot(pe_engine *h, uint32_t first, uint32_t n, pe_node_state *out) {
    Oracle &o = h->o;
    if ((uint64_t)first + n > o.nodes.size()) { o.err = "snapshot range"; return PE_ERR_INVALID; }
    for (uint32_t i = 0; i < n; i++) {
        const Node &nd = o.nodes[first + i];
        out[i].flags = nd.flags; out[i].total_tasks = nd.total; out[i].cpu_avail = nd.cpu; out[i].mem_avail = nd.mem;
    }
    return PE_OK;
}
int32_t ope_snapshot_service(pe_engine *h, uint32_t svc, uint32_t first, uint32_t n, uint32_t *out) {
    Oracle &o = h->o;
    if ((uint64_t)first + n > o.nodes.size()) { o.err = "snapshot range"; return PE_ERR_INVALID; }
    auto &c = o.col(o.svc, svc);
    for (uint32_t i = 0; i < n; i++) out[i] = c[first + i];
    return PE_OK;
}
int32_t ope_snapshot_generic(pe_engine *h, uint32_t kind, uint32_t first, uint32_t n, int64_t *out) {
    Oracle &o = h->o;
    if ((uint64_t)first + n > o.nodes.size()) { o.err = "snapshot range"; return PE_ERR_INVALID; }
    auto &c = o.col(o.gen, kind);
    for (uint32_t i = 0; i < n; i++) out[i] = c[first + i];
    return PE_OK;
}
int32_t ope_snapshot_ports(pe_engine *h, uint32_t slot, uint32_t first, uint32_t n, uint8_t *out) {
    Oracle &o = h->o;
    if ((uint64_t)first + n > o.nodes.size()) { o.err = "snapshot range"; return PE_ERR_INVALID; }
    auto &c = o.col(o.ports, slot >> 5);
    for (uint32_t i = 0; i < n; i++) out[i] = (c[first + i] >> (slot & 31)) & 1;
    return PE_OK;
}

int32_t ope_nccl_unique_id(void *) { return PE_ERR_UNSUPPORTED; }   // the oracle is one process, one table
int32_t ope_get_stats(pe_engine *h, pe_stats *out) { *out = h->o.stats; return PE_OK; }
int32_t ope_stats_reset(pe_engine *h) { h->o.stats = pe_stats{}; return PE_OK; }

// strings.EqualFold folding rule for valuePattern-admissible operands (constraint.go:26,90)
int32_t ope_fold_value(const char *in, uint32_t len, char *out, uint32_t cap) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < len;) {
        unsigned char c = (unsigned char)in[i];
        char r; uint32_t adv = 1;
        if (c >= 'A' && c <= 'Z') r = (char)(c + 32);
        else if (c == 0xE2 && i + 2 < len && (unsigned char)in[i + 1] == 0x84 && (unsigned char)in[i + 2] == 0xAA) { r = 'k'; adv = 3; }
        else if (c == 0xC5 && i + 1 < len && (unsigned char)in[i + 1] == 0xBF) { r = 's'; adv = 2; }
        else r = (char)c;
        if (w >= cap) return -1;
        out[w++] = r;
        i += adv;
    }
    return (int32_t)w;
}

}  // extern "C"
