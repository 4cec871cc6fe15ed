// place_model.cpp -- HOST MODEL of the engine's chunked parallel placement step
// (swarmkit_b200/csrc/kernel_place.cuh).  TEST INFRASTRUCTURE, NOT PRODUCT: it
// lives in oracle/, is built into oracle/_build/ and only tests/ load it.
//
// Why it exists: the CUDA kernel replaces the strictly ordered placement loop of
// the reference (Scheduler.tick's one-off loop, manager/scheduler/scheduler.go:
// 467-469 -> scheduleTaskGroup :694-748 -> nodeSet.tree nodeset.go:107-120 ->
// scheduleNTasksOnNodes scheduler.go:844-924 with a single task) by bulk-
// synchronous chunks: a parallel "stage" phase (every task of the chunk gets a
// sorted candidate list against the state at the chunk's start) and an ordered
// "resolve" phase (32 tasks at a time, lanes = tasks).  The exactness argument
//  is subtle, so the same algorithm is restated here on the CPU,
// phase by phase, with small chunk / candidate / list sizes, and compared with
// the sequential oracle (flat_oracle.cpp) on random workloads in
// tests/test_place_model_cpu.py.  The exported ABI is the oracle's with the
// prefix mpe_; everything but the k == 1 batch path is the oracle's own code.
//
// Knobs (environment, read per schedule call): PM_BATCH (tasks per scan batch),
// PM_CHUNK (tasks per chunk), PM_K (cap on the candidates per task: task i of a chunk gets min(i + 1, K); 0 = no cap), PM_LISTCAP (members
// listed per class), PM_GROUP (lanes per resolve group).
#define ope_abi_version mpe_abi_version
#define ope_create mpe_create
#define ope_destroy mpe_destroy
#define ope_last_error mpe_last_error
#define ope_node_upsert mpe_node_upsert
#define ope_node_remove mpe_node_remove
#define ope_set_node_count mpe_set_node_count
#define ope_node_task_delta mpe_node_task_delta
#define ope_schedule mpe_seq_schedule
#define ope_tick_upload mpe_tick_upload
#define ope_tick_run mpe_seq_tick_run
#define ope_tick_download mpe_tick_download
#define ope_fit mpe_fit
#define ope_snapshot mpe_snapshot
#define ope_snapshot_service mpe_snapshot_service
#define ope_snapshot_generic mpe_snapshot_generic
#define ope_snapshot_ports mpe_snapshot_ports
#define ope_get_stats mpe_get_stats
#define ope_stats_reset mpe_stats_reset
#define ope_fold_value mpe_fold_value
#define ope_nccl_unique_id mpe_nccl_unique_id
#define ope_pref_leaves mpe_pref_leaves
#define ope_match_matrix mpe_match_matrix
#include "flat_oracle.cpp"

#include <cstdlib>
#include <map>

namespace {

constexpr uint64_t KEY_NONE = ~0ull;
constexpr int MAX_RANKS = 4;   // rank 0 = untouched members of the best class; 1..3 = tail groups

uint32_t env_u32(const char *name, uint32_t dflt) {
    const char *s = std::getenv(name);
    return s && *s ? (uint32_t)std::strtoul(s, nullptr, 10) : dflt;
}

// rank prefix of nodeLess (scheduler.go:708-735), same packing as the engine (kernels_common.cuh make_pref)
uint64_t make_pref(uint32_t fails, uint32_t svc, uint32_t total) {
    uint32_t f5 = fails >= 5 ? (fails > 255 ? 255 : fails) : 0;
    return ((uint64_t)((f5 << 24) | (svc & 0xFFFFFF)) << 32) | total;
}

struct Row {
    uint32_t group;              // representative group
    uint64_t c0 = KEY_NONE, c1 = KEY_NONE;
    uint32_t n0 = 0, n1 = 0;
    std::vector<uint32_t> L0, L1;   // first LISTCAP members, node order
    bool eligible = false;       // the parallel path may place this row's tasks
    bool inline_ok = false;      // only cpu / memory / max-replicas can change feasibility (PE_SR_INLINE)
    bool f_res = false, f_max = false;
    uint32_t cur0 = 0, cur1 = 0; // list cursors: every member before is touched for good
};

struct Slot {
    bool cut = false;            // the parallel path cannot place this task: the ordered fall-back takes over here
    int why = 0;                 // diagnostic: 1 not eligible, 2 class beyond the list / no tail for the row, 3 empty tail
    std::vector<uint32_t> cand;  // sorted by (rank key at the chunk's start, node)
    std::vector<uint8_t> rank;   // index into keyval
    uint64_t keyval[MAX_RANKS];
};

struct LogEntry { uint32_t node, row, task; };

struct Model {
    Oracle &o;
    const Tick &t;
    uint32_t *out_node, *out_fail;
    uint32_t Bmax, C, K, LISTCAP, GROUP;
    uint64_t n_par = 0, n_cut = 0, n_amb = 0, n_tail = 0, n_retry = 0;
    uint64_t n_rounds = 0, n_skips = 0, n_groups_run = 0, n_maxskip = 0;   // diagnostics of the resolve loop

    bool desc_equal(const pe_group &a, const pe_group &b) const {
        if (a.svc_id != b.svc_id || a.filter_mask != b.filter_mask || a.cpu_res != b.cpu_res || a.mem_res != b.mem_res ||
            a.max_replicas != b.max_replicas || a.tie_start != b.tie_start || a.flags != b.flags || a.n_tasks != b.n_tasks ||
            t.task_flags[a.task_off] != t.task_flags[b.task_off] || a.log_plugin != b.log_plugin)
            return false;
        if (a.gen_cnt != b.gen_cnt || a.con_cnt != b.con_cnt || a.ip_cnt != b.ip_cnt || a.plat_cnt != b.plat_cnt ||
            a.port_cnt != b.port_cnt || a.plug_cnt != b.plug_cnt || a.fail_cnt != b.fail_cnt)
            return false;
        return !std::memcmp(t.gens + a.gen_off, t.gens + b.gen_off, a.gen_cnt * sizeof(pe_generic_want)) &&
               !std::memcmp(t.cons + a.con_off, t.cons + b.con_off, a.con_cnt * sizeof(pe_constraint)) &&
               !std::memcmp(t.ips + a.ip_off, t.ips + b.ip_off, a.ip_cnt * sizeof(pe_ip_constraint)) &&
               !std::memcmp(t.plats + a.plat_off, t.plats + b.plat_off, a.plat_cnt * sizeof(pe_platform)) &&
               !std::memcmp(t.ports + a.port_off, t.ports + b.port_off, a.port_cnt * 4) &&
               !std::memcmp(t.plugs + a.plug_off, t.plugs + b.plug_off, a.plug_cnt * 4) &&
               !std::memcmp(t.fails + a.fail_off, t.fails + b.fail_off, a.fail_cnt * sizeof(pe_node_fail));
    }

    // Pipeline.Process without the counters
    bool feasible(Sched &s, uint32_t n) const {
        uint32_t save[PE_NUM_FILTERS];
        std::memcpy(save, s.failcnt, sizeof save);
        bool ok = (o.nodes[n].flags & PE_NODE_VALID) && s.process(n);
        std::memcpy(s.failcnt, save, sizeof save);
        return ok;
    }

    // the batched scan (kernel_scan.cuh): two smallest rank prefixes + their first members, batch-start state
    void scan_row(Row &r) {
        const pe_group &g = t.groups[r.group];
        Sched s(o, t, g);
        const uint32_t N = o.n_nodes;
        std::vector<uint64_t> key(N, KEY_NONE);
        for (uint32_t n = 0; n < N; n++)
            if (feasible(s, n)) key[n] = make_pref(s.failures(n), s.svc_col[n], o.nodes[n].total);
        for (uint32_t n = 0; n < N; n++) if (key[n] < r.c0) r.c0 = key[n];
        for (uint32_t n = 0; n < N; n++) if (key[n] != r.c0 && key[n] < r.c1) r.c1 = key[n];
        for (uint32_t n = 0; n < N; n++) {
            if (r.c0 != KEY_NONE && key[n] == r.c0) { if (r.L0.size() < LISTCAP) r.L0.push_back(n); r.n0++; }
            if (r.c1 != KEY_NONE && key[n] == r.c1) { if (r.L1.size() < LISTCAP) r.L1.push_back(n); r.n1++; }
        }
        const uint32_t fm = g.filter_mask;
        const bool counts = (t.task_flags[g.task_off] & PE_T_COUNTS) != 0;
        r.inline_ok = g.gen_cnt == 0 && g.port_cnt == 0 && !(fm & (1u << PE_F_HOSTPORT)) && g.fail_cnt == 0;
        r.f_res = (fm >> PE_F_RESOURCE) & 1u;
        r.f_max = (fm >> PE_F_MAXREPLICAS) & 1u;
        r.eligible = g.n_tasks == 1 && g.tie_start == 0 && counts && r.c0 != KEY_NONE;
    }

    // ---- stage phase (parallel on the device: one warp per task), state as of the chunk's start ----
    // K: candidates wanted.  Lane i of a chunk can lose at most i candidates to the lower lanes (each task takes
    // one node), so i + 1 candidates can never run out: PM_K = 0 asks for exactly that.
    Slot stage(Row &r, const std::vector<uint8_t> &touched, uint32_t K) {
        Slot sl;
        for (int i = 0; i < MAX_RANKS; i++) sl.keyval[i] = KEY_NONE;
        if (!r.eligible) { sl.cut = true; sl.why = 1; return sl; }
        sl.keyval[0] = r.c0;
        // rank 0: the first K untouched members of the best class, from the row's cursor
        uint32_t p = r.cur0, first_untouched = (uint32_t)r.L0.size();
        for (; p < r.L0.size() && sl.cand.size() < K; p++)
            if (!touched[r.L0[p]]) {
                if (sl.cand.empty()) first_untouched = p;
                sl.cand.push_back(r.L0[p]); sl.rank.push_back(0);
            }
        r.cur0 = std::max(r.cur0, sl.cand.empty() ? p : first_untouched);
        if (sl.cand.size() == K) return sl;
        // fewer than K untouched members are listed
        if (r.n0 > r.L0.size() || !r.inline_ok) {      // the class goes on beyond the list / no tail for this kind of row
            if (sl.cand.empty()) { sl.cut = true; sl.why = 2; }
            return sl;                                  // (exhausting these candidates in the resolve phase is a cut)
        }
        // ---- tail: touched members of the best class at their LIVE rank, merged with the untouched members of
        // the second class.  Every other node ranked above c1 when the batch began (ranks only grow), so nothing
        // outside these two sources can rank <= c1.
        n_tail++;
        const pe_group &g = t.groups[r.group];
        Sched s(o, t, g);
        struct TC { uint64_t key; uint32_t node; };
        std::vector<TC> tc;
        for (uint32_t m : r.L0) {
            if (!touched[m]) continue;
            if (!feasible(s, m)) continue;
            tc.push_back({make_pref(0, s.svc_col[m], o.nodes[m].total), m});
        }
        std::vector<uint32_t> c1list;
        bool c1_incomplete = false;
        if (r.c1 != KEY_NONE) {
            uint32_t q = r.cur1, fu = (uint32_t)r.L1.size();
            for (; q < r.L1.size() && c1list.size() < K; q++)
                if (!touched[r.L1[q]]) { if (c1list.empty()) fu = q; c1list.push_back(r.L1[q]); }
            r.cur1 = std::max(r.cur1, c1list.empty() ? q : fu);
            if (c1list.size() < K && r.n1 > r.L1.size()) c1_incomplete = true;   // unlisted members may follow
        }
        uint64_t kprev = r.c0;
        for (int rank = 1; rank < MAX_RANKS && sl.cand.size() < K; rank++) {
            uint64_t kcur = KEY_NONE;
            for (auto &x : tc) if (x.key > kprev && x.key < kcur) kcur = x.key;
            if (r.c1 != KEY_NONE && r.c1 > kprev && r.c1 < kcur) kcur = r.c1;
            if (kcur == KEY_NONE) break;
            if (r.c1 != KEY_NONE && kcur > r.c1) break;   // beyond the second class nothing is known
            std::vector<uint32_t> grp;
            for (auto &x : tc) if (x.key == kcur) grp.push_back(x.node);   // node order (the list's order)
            std::vector<uint32_t> merged;
            if (kcur == r.c1) {
                merged.resize(grp.size() + c1list.size());
                std::merge(grp.begin(), grp.end(), c1list.begin(), c1list.end(), merged.begin());
                if (c1_incomplete) {   // only what precedes the last listed untouched member of the second class is certain
                    uint32_t lim = c1list.empty() ? 0u : c1list.back() + 1u;
                    while (!merged.empty() && merged.back() >= lim) merged.pop_back();
                }
            } else {
                merged = grp;
            }
            sl.keyval[rank] = kcur;
            for (uint32_t n : merged) {
                if (sl.cand.size() >= K) break;
                sl.cand.push_back(n); sl.rank.push_back((uint8_t)rank);
            }
            if (kcur == r.c1) break;
            kprev = kcur;
        }
        if (sl.cand.empty()) { sl.cut = true; sl.why = 3; }
        return sl;
    }

    // one batch [b0, b0 + B) of k == 1 groups
    void run_batch(uint32_t b0, uint32_t B) {
        const uint32_t N = o.n_nodes;
        std::vector<Row> rows;
        std::vector<uint32_t> task_row(B);
        for (uint32_t i = 0; i < B; i++) {
            const pe_group &g = t.groups[b0 + i];
            uint32_t r = 0;
            for (; r < rows.size(); r++) if (desc_equal(g, t.groups[rows[r].group])) break;
            if (r == rows.size()) { rows.emplace_back(); rows.back().group = b0 + i; scan_row(rows.back()); }
            task_row[i] = r;
        }
        std::vector<uint8_t> touched(N, 0);
        uint32_t cut = B;
        uint32_t done = 0;
        for (uint32_t c0 = 0; c0 < B && cut == B; c0 += done) {
            const uint32_t nc = std::min(C, B - c0);
            done = nc;      // tasks this chunk settles; fewer when a lane runs out of a capped candidate list
            // ---- stage: any order (parallel on the device); here reversed to expose order dependence
            std::vector<Slot> slots(nc);
            for (uint32_t i = nc; i-- > 0;) slots[i] = stage(rows[task_row[c0 + i]], touched, K ? std::min(K, i + 1) : i + 1);
            // ---- resolve: groups of GROUP lanes in task order
            std::map<uint32_t, std::vector<uint32_t>> H;   // node -> in-chunk placements (log indices)
            std::vector<LogEntry> log;
            for (uint32_t gb = 0; gb < nc && cut == B && done == nc; gb += GROUP) {
                const uint32_t ng = std::min(GROUP, nc - gb);
                uint32_t first_active = 0;
                while (first_active < ng && cut == B && done == nc) {
                    // the monotone loop: every active lane proposes its first candidate that is neither taken in
                    // this chunk nor proposed by a lower lane; a lane only ever moves forward
                    std::vector<uint32_t> j(ng, 0), prop(ng, PE_NONE);
                    n_groups_run++;
                    for (;;) {
                        n_rounds++;
                        uint64_t mx = 0;
                        for (uint32_t l = first_active; l < ng; l++) {
                            const Slot &sl = slots[gb + l];
                            uint64_t sk = 0;
                            while (j[l] < sl.cand.size() && H.count(sl.cand[j[l]])) { j[l]++; n_skips++; sk++; }
                            mx = std::max(mx, sk);
                            prop[l] = (!sl.cut && j[l] < sl.cand.size()) ? sl.cand[j[l]] : PE_NONE;
                        }
                        n_maxskip += mx;
                        bool any = false;
                        std::vector<uint8_t> kicked(ng, 0);
                        for (uint32_t l = first_active; l < ng; l++) {
                            if (prop[l] == PE_NONE) continue;
                            for (uint32_t m = first_active; m < l; m++) if (prop[m] == prop[l]) { kicked[l] = 1; any = true; break; }
                        }
                        if (!any) break;
                        for (uint32_t l = first_active; l < ng; l++) if (kicked[l]) j[l]++;
                    }
                    // lowest lane that is not a plain placement
                    uint32_t bad = ng; int why = 0;   // 1 cut / exhausted, 2 ambiguous
                    for (uint32_t l = first_active; l < ng; l++) {
                        const Slot &sl = slots[gb + l];
                        if (sl.cut || j[l] >= sl.cand.size()) { bad = l; why = 1; break; }
                        if (j[l] > 0 && sl.rank[0] < sl.rank[j[l]]) { bad = l; why = 2; break; }
                    }
                    auto commit_lane = [&](uint32_t l, uint32_t node) {
                        H[node].push_back((uint32_t)log.size());
                        log.push_back({node, task_row[c0 + gb + l], c0 + gb + l});
                    };
                    for (uint32_t l = first_active; l < bad; l++) commit_lane(l, prop[l]);
                    if (bad == ng) break;
                    if (why == 1 && K && !slots[gb + bad].cut && slots[gb + bad].cand.size() == K && gb + bad + 1 > K) {
                        done = gb + bad; n_retry++; break;     // its list was cut at K candidates: the next chunk starts with this task
                    }
                    if (why == 1) { cut = c0 + gb + bad; if (std::getenv("PM_DEBUG")) { const Slot &sl = slots[gb + bad]; const Row &r = rows[task_row[cut]]; fprintf(stderr, "cut task %u: slotcut=%d why=%d ncand=%zu c0=%llx c1=%llx n0=%u n1=%u inline=%d\n", cut, (int)sl.cut, sl.why, sl.cand.size(), (unsigned long long)r.c0, (unsigned long long)r.c1, r.n0, r.n1, (int)r.inline_ok); } break; }
                    // ---- ambiguous lane: a skipped candidate ranked strictly better than the chosen one when the
                    // chunk began; it was taken inside the chunk, so its rank moved -- recompute it exactly
                    n_amb++;
                    {
                        const Slot &sl = slots[gb + bad];
                        const Row &r = rows[task_row[c0 + gb + bad]];
                        const pe_group &g = t.groups[r.group];
                        uint64_t bk = sl.keyval[sl.rank[j[bad]]];
                        uint32_t bn = sl.cand[j[bad]];
                        for (uint32_t i = 0; i < j[bad]; i++) {
                            const uint32_t n = sl.cand[i];
                            const uint64_t k0 = sl.keyval[sl.rank[i]];
                            uint32_t svc = (uint32_t)(k0 >> 32) & 0xFFFFFF, tot = (uint32_t)k0;
                            int64_t dcpu = 0, dmem = 0;
                            for (uint32_t e : H.at(n)) {
                                const pe_group &h = t.groups[rows[log[e].row].group];
                                tot++;
                                if (h.svc_id == g.svc_id) svc++;
                                dcpu += h.cpu_res; dmem += h.mem_res;
                            }
                            bool ok = true;
                            if (r.f_res) ok = g.cpu_res <= o.nodes[n].cpu - dcpu && g.mem_res <= o.nodes[n].mem - dmem;
                            if (r.f_max) ok = ok && (uint64_t)svc < g.max_replicas;
                            const uint64_t k = make_pref(0, svc, tot);
                            if (ok && (k < bk || (k == bk && n < bn))) { bk = k; bn = n; }
                        }
                        commit_lane(bad, bn);
                    }
                    first_active = bad + 1;
                }
            }
            // ---- commit: NodeInfo.addTask for the chunk's placements (parallel reductions on the device)
            for (const LogEntry &e : log) {
                const pe_group &g = t.groups[b0 + e.task];
                Sched s(o, t, g);
                s.add_task(e.node, true);
                o.stats.placements++;
                out_node[g.task_off] = e.node;
                std::memset(out_fail + (size_t)(b0 + e.task) * PE_NUM_FILTERS, 0, PE_NUM_FILTERS * 4);
                touched[e.node] = 1;
                n_par++;
            }
        }
        // ---- the ordered fall-back takes the rest of the batch (on the device: k_sequencer from the cut on)
        if (cut < B) n_cut++;
        for (uint32_t i = cut; i < B; i++) schedule_group(o, t, b0 + i, out_node, out_fail);
    }

    void run() {
        uint32_t gi = 0;
        while (gi < t.n_groups) {
            uint32_t e = gi;
            const bool one = t.groups[gi].n_tasks == 1;
            while (e < t.n_groups && (t.groups[e].n_tasks == 1) == one) e++;
            if (one && e - gi >= 4 && o.n_nodes > 0) {
                for (uint32_t b0 = gi; b0 < e; b0 += Bmax) run_batch(b0, std::min(Bmax, e - b0));
            } else {
                for (uint32_t g = gi; g < e; g++) schedule_group(o, t, g, out_node, out_fail);
            }
            gi = e;
        }
    }
};

uint64_t g_model_counters[4];

}  // namespace

extern "C" {

int32_t mpe_schedule(pe_engine *h, const pe_tick *tk, uint32_t *out_node, uint32_t *out_fail) {
    Oracle &o = h->o;
    int32_t rc = validate(o, tk);
    if (rc) return rc;
    Tick t = view(tk);
    Model m{o, t, out_node, out_fail, env_u32("PM_BATCH", 512), env_u32("PM_CHUNK", 128), env_u32("PM_K", 32),
            env_u32("PM_LISTCAP", 1024), env_u32("PM_GROUP", 32)};
    m.run();
    g_model_counters[0] += m.n_par; g_model_counters[1] += m.n_cut; g_model_counters[2] += m.n_amb; g_model_counters[3] += m.n_tail;
    if (std::getenv("PM_DEBUG")) fprintf(stderr, "resolve: passes %llu rounds %llu skips %llu sum-of-max-skips-per-round %llu\n", (unsigned long long)m.n_groups_run, (unsigned long long)m.n_rounds, (unsigned long long)m.n_skips, (unsigned long long)m.n_maxskip);
    return PE_OK;
}

// tasks placed by the parallel path / batches cut short / ambiguous lanes / tails built (since the last call)
void mpe_model_counters(uint64_t *out4) {
    for (int i = 0; i < 4; i++) { out4[i] = g_model_counters[i]; g_model_counters[i] = 0; }
}

int32_t mpe_tick_run(pe_engine *h) {
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
    return mpe_schedule(h, &tk, o.r_node.data(), o.r_fail.data());
}

}  // extern "C"
