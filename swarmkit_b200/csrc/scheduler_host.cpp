// scheduler_host.cpp -- host side of the drop-in: a C++ mirror of the Go
// scheduler (manager/scheduler/scheduler.go) that keeps the reference's
// function names, argument meaning and error behaviour, but hands the hot path
// -- filter pipeline, spread ranking, reservation -- to the CUDA engine through
// the C ABI in include/placement_engine.h.
//
// The reference is Go and there is no Go toolchain in this image, so the host
// side is written in C++ (the cgo binding a maintainer would add is shown in
// INTEGRATION.md).  What stays on the host is exactly what the Go shim would
// keep: the store/watch plumbing (here: a JSON event feed used by the tests),
// NodeInfo bookkeeping incl. the named generic-resource member lists
// (SURVEY hard part D), string interning / case folding / IP parsing (hard part
// C), Explain strings, and the commit/rollback of decisions.
//
//   Scheduler::setupTasksList        scheduler.go:68-125
//   Scheduler::createTask/updateTask/deleteTask            :254-366
//   Scheduler::createOrUpdateNode    :368-396
//   Scheduler::processPreassignedTasks / taskFitNode       :398-426, :646-690
//   Scheduler::tick / scheduleTaskGroup / noSuitableNode   :429-488, :694-748, :928-971
//   NodeInfo::addTask/removeTask/taskFailed/countRecentFailures   nodeinfo.go:66-221
//   Encoder (rows, group descriptors) <- the SetTask methods      filter.go:35,60,118,224,259,328,369
//
// CSI cluster volumes (VolumesFilter): the volume bookkeeping and the filter stay on the host as in the
// reference; the shim hands the engine the node set the filter allows (scheduleVolumeGroup).  Groups whose volume
// availability changes with every placement walk the reference's fill loop with one engine question per step
// (scheduleVolumeGroupStepwise); every filter evaluation and every ranking of nodes is the engine's.
// Placement.Preferences (scheduler.go:772-825): the branch walk is host code here as in the
// reference; the engine supplies the tree's leaves and fills one leaf per group.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/placement_engine.h"
#include "minijson.h"

namespace sk {

// ------------------------------------------------------------------ api objects
enum { TaskStatePending = 64, TaskStateAssigned = 192, TaskStateRunning = 512, TaskStateCompleted = 576, TaskStateShutdown = 640,
       TaskStateFailed = 704, TaskStateRejected = 768 };
enum { NodeReady = 2, AvailActive = 0, PublishHost = 1, MountVolume = 1, MountCluster = 4 };

struct Generic { bool named = false; std::string kind, value; int64_t amount = 0; };
struct Resources { int64_t cpu = 0, mem = 0; std::vector<Generic> generic; };
struct Plugin { std::string type, name; };
struct CsiInfo { std::string plugin, node_id; bool has_top = false; std::map<std::string, std::string> segments; };   // api.NodeCSIInfo
struct Node {
    std::vector<CsiInfo> csi;
    std::string id, hostname, addr, os, arch;
    bool has_desc = false, has_platform = false, has_resources = false, has_engine = false, has_labels = false, has_elabels = false;
    Resources resources;
    std::map<std::string, std::string> labels, elabels;
    std::vector<Plugin> plugins;
    int state = 0, availability = 0, role = 0;
    uint64_t version = 0;
};
struct Port { int protocol = 0; uint32_t port = 0; int mode = 0; };
struct Mount { int type = 0; bool has_driver = false; std::string driver; std::string source, target; bool read_only = false; };
struct Attachment { std::string id, source, target; };                     // api.VolumeAttachment
struct Volume {                                                             // api.Volume, what the scheduler reads of it
    std::string id, name, group, driver;
    int availability = 0, scope = 0, sharing = 0;                           // ACTIVE / SINGLE_NODE / NONE = 0
    bool has_info = false; std::string volume_id;
    std::vector<std::map<std::string, std::string>> accessible;
};
enum { VolScopeSingle = 0, VolScopeMulti = 1, VolShareNone = 0, VolShareReadOnly = 1, VolShareOneWriter = 2, VolShareAll = 3 };
struct Task {
    std::string id, service, node_id, err, message;
    int desired = TaskStateRunning, state = 0;
    bool has_spec_version = false; uint64_t spec_version = 0;
    bool has_res = false; Resources res;
    bool has_placement = false;
    std::vector<std::string> constraints, preferences;
    std::vector<std::pair<std::string, std::string>> platforms;  // (arch, os)
    uint64_t max_replicas = 0;
    bool has_container = false; std::vector<Mount> mounts;
    bool has_log = false; std::string log_driver;
    std::vector<std::pair<bool, std::string>> networks;
    bool has_endpoint = false; std::vector<Port> ports;
    std::vector<Generic> assigned;
    std::vector<Attachment> volumes;
    mj::Value raw;  // echoed back in snapshots of decisions
};
using TaskP = std::shared_ptr<Task>;
using NodeP = std::shared_ptr<Node>;

// ------------------------------------------------------------------ JSON -> objects
static int enum_of(const mj::Value &v, std::initializer_list<std::pair<const char *, int>> names, int def) {
    if (v.type == mj::Value::Int) return (int)v.i;
    if (v.type == mj::Value::Str) for (auto &p : names) if (v.s == p.first) return p.second;
    return def;
}
static int task_state(const mj::Value &v, int def) {
    return enum_of(v, {{"NEW", 0}, {"PENDING", 64}, {"ASSIGNED", 192}, {"ACCEPTED", 256}, {"PREPARING", 320}, {"READY", 384}, {"STARTING", 448},
                       {"RUNNING", 512}, {"COMPLETE", 576}, {"COMPLETED", 576}, {"SHUTDOWN", 640}, {"FAILED", 704}, {"REJECTED", 768},
                       {"REMOVE", 800}, {"ORPHANED", 832}}, def);
}
static const char *task_state_name(int s) {
    switch (s) { case 0: return "NEW"; case 64: return "PENDING"; case 192: return "ASSIGNED"; case 256: return "ACCEPTED"; case 320: return "PREPARING";
                 case 384: return "READY"; case 448: return "STARTING"; case 512: return "RUNNING"; case 576: return "COMPLETE"; case 640: return "SHUTDOWN";
                 case 704: return "FAILED"; case 768: return "REJECTED"; case 800: return "REMOVE"; case 832: return "ORPHANED"; }
    return "?";
}
static Resources parse_resources(const mj::Value &v) {
    Resources r;
    r.cpu = v.at("nano_cpus").as_int(); r.mem = v.at("memory_bytes").as_int();
    for (auto &g : v.at("generic").a) {
        Generic x; x.kind = g.at("kind").as_str();
        if (g.find("named")) { x.named = true; x.value = g.at("named").as_str(); } else x.amount = g.at("value").as_int();
        r.generic.push_back(x);
    }
    return r;
}
static mj::Value generic_json(const std::vector<Generic> &gs) {
    mj::Value a = mj::Value::array();
    for (auto &x : gs) {
        mj::Value e = mj::Value::object(); e.set("kind", mj::Value::string(x.kind));
        if (x.named) e.set("named", mj::Value::string(x.value)); else e.set("value", mj::Value::integer(x.amount));
        a.push(e);
    }
    return a;
}
static NodeP parse_node(const mj::Value &v) {
    NodeP n(new Node());
    n->id = v.at("id").as_str();
    n->role = enum_of(v.at("role"), {{"WORKER", 0}, {"MANAGER", 1}}, 0);
    n->version = (uint64_t)v.at("version").as_int();
    n->availability = enum_of(v.at("spec").at("availability"), {{"ACTIVE", 0}, {"PAUSE", 1}, {"DRAIN", 2}}, 0);
    const mj::Value &lb = v.at("spec").at("labels");
    if (!lb.is_null()) { n->has_labels = true; for (auto &kv : lb.o) n->labels[kv.first] = kv.second.as_str(); }
    n->state = enum_of(v.at("status").at("state"), {{"UNKNOWN", 0}, {"DOWN", 1}, {"READY", 2}, {"DISCONNECTED", 3}}, 0);
    n->addr = v.at("status").at("addr").as_str();
    const mj::Value &d = v.at("description");
    if (!d.is_null()) {
        n->has_desc = true;
        n->hostname = d.at("hostname").as_str();
        if (!d.at("platform").is_null()) { n->has_platform = true; n->os = d.at("platform").at("os").as_str(); n->arch = d.at("platform").at("arch").as_str(); }
        if (!d.at("resources").is_null()) { n->has_resources = true; n->resources = parse_resources(d.at("resources")); }
        const mj::Value &e = d.at("engine");
        if (!e.is_null()) {
            n->has_engine = true;
            if (!e.at("labels").is_null()) { n->has_elabels = true; for (auto &kv : e.at("labels").o) n->elabels[kv.first] = kv.second.as_str(); }
            for (auto &p : e.at("plugins").a) n->plugins.push_back({p.at("type").as_str(), p.at("name").as_str()});
        }
        if (!d.at("csi_info").is_null())
            for (auto &c : d.at("csi_info").a) {
                CsiInfo ci; ci.plugin = c.at("plugin").as_str(); ci.node_id = c.at("node_id").as_str();
                if (!c.at("topology").is_null()) { ci.has_top = true; for (auto &kv : c.at("topology").o) ci.segments[kv.first] = kv.second.as_str(); }
                n->csi.push_back(ci);
            }
    }
    return n;
}
static Volume parse_volume(const mj::Value &v) {
    Volume x;
    x.id = v.at("id").as_str(); x.name = v.at("name").as_str(); x.group = v.at("group").as_str(); x.driver = v.at("driver").as_str();
    x.availability = enum_of(v.at("availability"), {{"ACTIVE", 0}, {"PAUSE", 1}, {"DRAIN", 2}}, 0);
    x.scope = enum_of(v.at("scope"), {{"SINGLE_NODE", 0}, {"MULTI_NODE", 1}}, 0);
    x.sharing = enum_of(v.at("sharing"), {{"NONE", 0}, {"READ_ONLY", 1}, {"ONE_WRITER", 2}, {"ALL", 3}}, 0);
    if (!v.at("volume_info").is_null()) {
        x.has_info = true; x.volume_id = v.at("volume_info").at("volume_id").as_str();
        if (!v.at("volume_info").at("accessible_topology").is_null())
            for (auto &t : v.at("volume_info").at("accessible_topology").a) { std::map<std::string, std::string> sg; for (auto &kv : t.o) sg[kv.first] = kv.second.as_str(); x.accessible.push_back(sg); }
    }
    return x;
}
static TaskP parse_task(const mj::Value &v) {
    TaskP t(new Task());
    t->id = v.at("id").as_str(); t->service = v.at("service_id").as_str(); t->node_id = v.at("node_id").as_str();
    t->desired = v.find("desired_state") ? task_state(v.at("desired_state"), TaskStateRunning) : TaskStateRunning;
    t->state = task_state(v.at("status").at("state"), 0);
    t->err = v.at("status").at("err").as_str(); t->message = v.at("status").at("message").as_str();
    if (!v.at("spec_version").is_null()) { t->has_spec_version = true; t->spec_version = (uint64_t)v.at("spec_version").as_int(); }
    const mj::Value &spec = v.at("spec");
    if (!spec.at("resources").is_null() && !spec.at("resources").at("reservations").is_null()) { t->has_res = true; t->res = parse_resources(spec.at("resources").at("reservations")); }
    const mj::Value &pl = spec.at("placement");
    if (!pl.is_null()) {
        t->has_placement = true;
        for (auto &c : pl.at("constraints").a) t->constraints.push_back(c.as_str());
        for (auto &c : pl.at("preferences").a) t->preferences.push_back(c.as_str());
        for (auto &p : pl.at("platforms").a) t->platforms.push_back({p.at("arch").as_str(), p.at("os").as_str()});
        t->max_replicas = (uint64_t)pl.at("max_replicas").as_int();
    }
    if (!spec.at("container").is_null()) {
        t->has_container = true;
        for (auto &m : spec.at("container").at("mounts").a) {
            Mount mm; mm.type = enum_of(m.at("type"), {{"BIND", 0}, {"VOLUME", 1}, {"TMPFS", 2}, {"NPIPE", 3}, {"CLUSTER", 4}}, 0);
            if (!m.at("driver").is_null()) { mm.has_driver = true; mm.driver = m.at("driver").as_str(); }
            mm.source = m.at("source").as_str(); mm.target = m.at("target").as_str(); mm.read_only = !m.at("read_only").is_null() && m.at("read_only").as_bool();
            t->mounts.push_back(mm);
        }
    }
    if (!spec.at("log_driver").is_null()) { t->has_log = true; t->log_driver = spec.at("log_driver").at("name").as_str(); }
    for (auto &nw : v.at("networks").a) t->networks.push_back({!nw.at("driver").is_null(), nw.at("driver").as_str()});
    if (!v.at("endpoint").is_null()) {
        t->has_endpoint = true;
        for (auto &p : v.at("endpoint").at("ports").a) {
            Port pc; pc.protocol = enum_of(p.at("protocol"), {{"TCP", 0}, {"UDP", 1}, {"SCTP", 2}}, 0);
            pc.port = (uint32_t)p.at("published_port").as_int();
            pc.mode = enum_of(p.at("publish_mode"), {{"INGRESS", 0}, {"HOST", 1}}, 0);
            t->ports.push_back(pc);
        }
    }
    for (auto &g : v.at("assigned_generic").a) {
        Generic x; x.kind = g.at("kind").as_str();
        if (g.find("named")) { x.named = true; x.value = g.at("named").as_str(); } else x.amount = g.at("value").as_int();
        t->assigned.push_back(x);
    }
    if (!v.at("volumes").is_null()) for (auto &a : v.at("volumes").a) t->volumes.push_back({a.at("id").as_str(), a.at("source").as_str(), a.at("target").as_str()});
    return t;
}

// ------------------------------------------------------------------ generic resources (host keeps the member lists)
namespace gr {
static std::vector<Generic *> of_kind(std::vector<Generic> &l, const std::string &k) { std::vector<Generic *> o; for (auto &g : l) if (g.kind == k) o.push_back(&g); return o; }
// Claim = select then consume (api/genericresource/resource_management.go:11-72, helpers.go:58-111)
static void claim(std::vector<Generic> &avail, std::vector<Generic> &assigned, const std::vector<Generic> &wants) {
    std::vector<Generic> sel;
    for (auto &w : wants) {
        if (w.named) return;
        std::vector<Generic> picked; bool any = false, done = false;
        for (auto &r : avail) {
            if (r.kind != w.kind) continue;
            any = true;
            if (!r.named) { if (r.amount >= w.amount && w.amount != 0) { Generic d; d.kind = w.kind; d.amount = w.amount; picked.push_back(d); } done = true; break; }
            picked.push_back(r);
            if ((int64_t)picked.size() == w.amount) { done = true; break; }
        }
        if (!done && picked.empty()) return;  // "not enough resources": nothing at all is claimed
        (void)any;
        sel.insert(sel.end(), picked.begin(), picked.end());
    }
    assigned.insert(assigned.end(), sel.begin(), sel.end());
    std::vector<Generic> keep;
    for (auto na : avail) {
        bool gone = false;
        for (auto &r : sel) {
            if (na.kind != r.kind) continue;
            if (!r.named) { if (na.named) continue; na.amount -= r.amount; if (na.amount <= 0) { gone = true; break; } }
            else if (na.named && na.value == r.value) { gone = true; break; }
        }
        if (!gone) keep.push_back(na);
    }
    avail.swap(keep);
}
static void consume(std::vector<Generic> &avail, const std::vector<Generic> &res) {  // ConsumeNodeResources
    std::vector<Generic> keep;
    for (auto na : avail) {
        bool gone = false;
        for (auto &r : res) {
            if (na.kind != r.kind) continue;
            if (!r.named) { if (na.named) continue; na.amount -= r.amount; if (na.amount <= 0) { gone = true; break; } }
            else if (na.named && na.value == r.value) { gone = true; break; }
        }
        if (!gone) keep.push_back(na);
    }
    avail.swap(keep);
}
// Reclaim = reclaimResources + sanitize (resource_management.go:75-205)
static void reclaim(std::vector<Generic> &avail, const std::vector<Generic> &assigned, const std::vector<Generic> &spec) {
    for (auto &r : assigned) {
        if (r.named) { avail.push_back(r); continue; }
        auto cur = of_kind(avail, r.kind);
        if (cur.empty()) { avail.push_back(r); continue; }
        if (cur.size() != 1 || cur[0]->named) continue;
        cur[0]->amount += r.amount;
    }
    std::vector<Generic> kept, reset; std::set<std::string> done;
    for (auto &na : avail) {
        std::vector<Generic> sp; for (auto &s : spec) if (s.kind == na.kind) sp.push_back(s);
        bool ok;
        if (!na.named) ok = sp.size() == 1 && !sp[0].named && na.amount <= sp[0].amount;
        else {
            ok = false; bool type_change = sp.empty();
            for (auto &s : sp) { if (!s.named) { type_change = true; break; } if (s.value == na.value) { ok = true; break; } }
            if (!ok && !type_change) sp.clear();  // member removed from the spec: drop it, add nothing back
        }
        if (ok) { kept.push_back(na); continue; }
        if (done.count(na.kind)) continue;
        done.insert(na.kind);
        reset.insert(reset.end(), sp.begin(), sp.end());
    }
    kept.insert(kept.end(), reset.begin(), reset.end());
    avail.swap(kept);
}
}  // namespace gr

// ------------------------------------------------------------------ NodeInfo (host mirror, nodeinfo.go:28-44)
static const int64_t kMonitorFailures = 300LL * 1000000000LL;  // scheduler.go:19
static const int kMaxFailures = 5;                              // scheduler.go:28
struct SvcVer { std::string svc; uint64_t ver; bool operator<(const SvcVer &o) const { return svc != o.svc ? svc < o.svc : ver < o.ver; } };
struct NodeInfo {
    NodeP node;
    std::map<std::string, TaskP> tasks;
    int active = 0;
    std::map<std::string, int> by_service;
    Resources avail;
    std::set<std::pair<int, uint32_t>> ports;
    std::map<SvcVer, std::vector<int64_t>> failures;
    int64_t last_cleanup = 0;

    static Resources reservations(const Task &t) { return t.has_res ? t.res : Resources(); }
    bool addTask(const TaskP &t) {  // nodeinfo.go:108-154
        auto it = tasks.find(t->id);
        if (it != tasks.end()) {
            bool was = it->second->desired <= TaskStateCompleted, is = t->desired <= TaskStateCompleted;
            if (is && !was) { it->second = t; active++; by_service[t->service]++; return true; }
            if (!is && was) { it->second = t; active--; by_service[t->service]--; return true; }
            return false;
        }
        tasks[t->id] = t;
        Resources r = reservations(*t);
        avail.mem -= r.mem; avail.cpu -= r.cpu;
        t->assigned.clear();
        gr::claim(avail.generic, t->assigned, r.generic);
        if (t->has_endpoint) for (auto &p : t->ports) if (p.mode == PublishHost && p.port) ports.insert({p.protocol, p.port});
        if (t->desired <= TaskStateCompleted) { active++; by_service[t->service]++; }
        return true;
    }
    bool removeTask(const Task &t) {  // nodeinfo.go:66-104
        auto it = tasks.find(t.id);
        if (it == tasks.end()) return false;
        bool was = it->second->desired <= TaskStateCompleted;
        tasks.erase(it);
        if (was) { active--; by_service[t.service]--; }
        if (t.has_endpoint) for (auto &p : t.ports) if (p.mode == PublishHost && p.port) ports.erase({p.protocol, p.port});
        Resources r = reservations(t);
        avail.mem += r.mem; avail.cpu += r.cpu;
        if (!node->has_desc || !node->has_resources) return true;
        gr::reclaim(avail.generic, t.assigned, node->resources.generic);
        return true;
    }
    void taskFailed(const Task &t, int64_t now) {  // nodeinfo.go:163-202
        if (now - last_cleanup >= kMonitorFailures) {
            for (auto it = failures.begin(); it != failures.end();) {
                bool recent = false;
                for (int64_t ts : it->second) if (now - ts < kMonitorFailures) { recent = true; break; }
                it = recent ? std::next(it) : failures.erase(it);
            }
            last_cleanup = now;
        }
        auto &l = failures[{t.service, t.has_spec_version ? t.spec_version : 0}];
        size_t expired = 0;
        for (int64_t ts : l) { if (now - ts < kMonitorFailures) break; expired++; }
        l.erase(l.begin(), l.begin() + expired);
        l.push_back(now);
    }
    int countRecentFailures(int64_t now, const SvcVer &k) const {  // nodeinfo.go:206-221
        auto it = failures.find(k);
        if (it == failures.end()) return 0;
        int n = (int)it->second.size();
        for (int i = n - 1; i >= 0; i--) if (now - it->second[i] > kMonitorFailures) { n -= i + 1; break; }
        return n;
    }
};

// ------------------------------------------------------------------ strings / addresses
static std::string fold(const std::string &s) {
    std::string o(s.size(), 0);
    int32_t n = pe_fold_value(s.data(), (uint32_t)s.size(), &o[0], (uint32_t)o.size());
    o.resize(n < 0 ? 0 : n);
    return o;
}
static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    auto sp = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (a < b && sp((unsigned char)s[a])) a++;
    while (b > a && sp((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
struct Addr { bool ok = false, v4 = false; uint32_t w[4] = {0, 0, 0, 0}; };
static bool parse_v4(const std::string &s, uint8_t out[4]) {
    size_t p = 0;
    for (int f = 0; f < 4; f++) {
        if (f) { if (p >= s.size() || s[p] != '.') return false; p++; }
        size_t st = p; int v = 0;
        while (p < s.size() && s[p] >= '0' && s[p] <= '9') { v = v * 10 + (s[p] - '0'); if (v > 255) return false; p++; }
        if (p == st || (p - st > 1 && s[st] == '0')) return false;
        out[f] = (uint8_t)v;
    }
    return p == s.size();
}
static Addr parse_ip(const std::string &s) {  // net.ParseIP
    Addr a; uint8_t b[16] = {0}, v4[4];
    if (s.find(':') == std::string::npos) {
        if (!parse_v4(s, v4)) return a;
        b[10] = b[11] = 0xff; memcpy(b + 12, v4, 4);
    } else {
        std::vector<uint16_t> head, tail; bool ell = false; size_t p = 0;
        if (s.size() >= 2 && s[0] == ':' && s[1] == ':') { ell = true; p = 2; }
        auto *cur = ell ? &tail : &head;
        while (p < s.size()) {
            size_t st = p; uint32_t v = 0; int nd = 0;
            while (p < s.size() && isxdigit((unsigned char)s[p]) && nd < 5) { char c = (char)tolower(s[p]); v = v * 16 + (uint32_t)(c <= '9' ? c - '0' : c - 'a' + 10); p++; nd++; }
            if (nd == 0 || nd > 4) return a;
            if (p < s.size() && s[p] == '.') { if (!parse_v4(s.substr(st), v4)) return a; cur->push_back((uint16_t)(v4[0] << 8 | v4[1])); cur->push_back((uint16_t)(v4[2] << 8 | v4[3])); p = s.size(); break; }
            cur->push_back((uint16_t)v);
            if (p == s.size()) break;
            if (s[p] != ':') return a;
            p++;
            if (p < s.size() && s[p] == ':') { if (ell) return a; ell = true; cur = &tail; p++; }
            else if (p == s.size()) return a;
        }
        size_t n = head.size() + tail.size();
        if ((!ell && n != 8) || (ell && n > 7)) return a;
        head.resize(8 - tail.size(), 0); head.insert(head.end(), tail.begin(), tail.end());
        for (int i = 0; i < 8; i++) { b[2 * i] = (uint8_t)(head[i] >> 8); b[2 * i + 1] = (uint8_t)head[i]; }
    }
    a.ok = true;
    static const uint8_t pre[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff};
    a.v4 = memcmp(b, pre, 12) == 0;
    for (int i = 0; i < 4; i++) a.w[i] = (uint32_t)b[4 * i] << 24 | (uint32_t)b[4 * i + 1] << 16 | (uint32_t)b[4 * i + 2] << 8 | b[4 * i + 3];
    return a;
}

// ------------------------------------------------------------------ the scheduler
// ------------------------------------------------------------------ CSI cluster volumes (host side)
// The reference keeps this bookkeeping in the scheduler and evaluates VolumesFilter per node (filter.go:388-447,
// volumes.go:45-327, topology.go:22-47); it "stays on host" here too: the shim decides which nodes can satisfy a task's
// cluster mounts and hands the engine that node set (see Scheduler::scheduleVolumeGroup).
struct VolumeBook {
    struct Use { std::string node; bool read_only = false; };
    struct Entry { Volume v; std::map<std::string, Use> tasks; std::map<std::string, int> nodes; };
    std::map<std::string, Entry> vols;
    std::map<std::string, std::set<std::string>> by_group;     // a group's volumes in ascending ID (canonical order)
    std::map<std::string, std::string> by_name;
    uint64_t epoch = 0;      // moves whenever an answer of check() may have: the shim keeps VolumesFilter's last node set by it

    // does a use of this volume (by a reader / a writer) change what check() says about it for anyone else?
    static bool counts_users(const Volume &v, bool ro) { return v.scope == VolScopeSingle || v.sharing == VolShareNone || (v.sharing == VolShareOneWriter && !ro); }
    void addOrUpdate(const Volume &v) {                         // volumes.go:61-81 (an update keeps the stored spec: :69-70)
        epoch++;
        if (!vols.count(v.id)) { Entry e; e.v = v; vols[v.id] = e; }
        by_group[v.group].insert(v.id);
        by_name[v.name] = v.id;
    }
    void remove(const std::string &id) {                        // :83-96
        auto it = vols.find(id);
        if (it == vols.end()) return;
        epoch++;
        by_group[it->second.v.group].erase(id);
        by_name.erase(it->second.v.name);
        vols.erase(it);
    }
    void reserve(const std::string &vid, const std::string &task, const std::string &node, bool ro) {   // :150-160
        auto it = vols.find(vid);
        if (it == vols.end()) return;
        if (counts_users(it->second.v, ro)) epoch++;
        it->second.tasks[task] = Use{node, ro};
        it->second.nodes[node]++;
    }
    void release(const std::string &vid, const std::string &task) {                                     // :162-184
        auto it = vols.find(vid);
        if (it == vols.end()) return;
        auto u = it->second.tasks.find(task);
        if (u == it->second.tasks.end()) return;
        if (counts_users(it->second.v, u->second.read_only)) epoch++;
        int &c = it->second.nodes[u->second.node];
        if (c > 0) c--;
        it->second.tasks.erase(u);
    }
    static bool inTopology(const CsiInfo *ci, const std::vector<std::map<std::string, std::string>> &accessible) {   // topology.go:22-47
        if (!ci || !ci->has_top || accessible.empty()) return true;
        for (auto &want : accessible) {
            bool all = true;
            for (auto &kv : want) { auto f = ci->segments.find(kv.first); if (f == ci->segments.end() ? !kv.second.empty() : f->second != kv.second) { all = false; break; } }
            if (all) return true;
        }
        return false;
    }
    bool check(const std::string &id, const Node &n, bool ro) const {                                    // :257-318
        auto it = vols.find(id);
        if (it == vols.end()) return false;
        const Entry &e = it->second;
        if (e.v.availability != 0) return false;
        const CsiInfo *ci = nullptr;
        if (n.has_desc) for (auto &c : n.csi) if (c.plugin == e.v.driver) { ci = &c; break; }
        if (e.v.scope == VolScopeSingle) for (auto &u : e.tasks) if (u.second.node != n.id) return false;
        if (e.v.sharing == VolShareNone && !e.tasks.empty()) return false;
        if (e.v.sharing == VolShareReadOnly && !ro) return false;
        if (e.v.sharing == VolShareOneWriter && !ro) for (auto &u : e.tasks) if (!u.second.read_only) return false;
        return inTopology(ci, e.v.has_info ? e.v.accessible : std::vector<std::map<std::string, std::string>>());
    }
    std::string availableOn(const Mount &m, const Node &n) const {                                        // :223-255
        if (m.source.compare(0, 6, "group:") == 0) {
            auto g = by_group.find(m.source.substr(6));
            if (g == by_group.end()) return "";
            for (auto &id : g->second) if (check(id, n, m.read_only)) return id;
            return "";
        }
        auto f = by_name.find(m.source);
        return (f != by_name.end() && check(f->second, n, m.read_only)) ? f->second : std::string();
    }
    // VolumesFilter.Check, filter.go:432-440: ONE satisfiable cluster mount is enough
    bool filterCheck(const Task &t, const Node &n) const {
        for (auto &m : t.mounts) if (m.type == MountCluster && !availableOn(m, n).empty()) return true;
        return false;
    }
    // chooseTaskVolumes, :98-136
    bool choose(const Task &t, const Node &n, std::vector<Attachment> &out, std::string &err) {
        std::vector<Attachment> got; bool ok = true;
        if (t.has_container)
            for (auto &m : t.mounts) {
                if (m.type != MountCluster) continue;
                std::string id = availableOn(m, n);
                if (id.empty()) { err = "cannot find volume to satisfy mount with source " + m.source; ok = false; break; }
                reserve(id, t.id, n.id, m.read_only);
                got.push_back({id, m.source, m.target});
            }
        for (auto &a : got) release(a.id, t.id);
        out = ok ? got : std::vector<Attachment>();
        return ok;
    }
    void reserveTask(const Task &t) {                                                                     // :138-148
        for (auto &va : t.volumes) for (auto &m : t.mounts) if (m.source == va.source && m.target == va.target) reserve(va.id, t.id, t.node_id, m.read_only);
    }
    // Can a placement of one task of this spec change what VolumesFilter answers for the next task of the same group?
    // (scope SINGLE_NODE pins the volume to the first node that uses it; sharing NONE / ONE_WRITER count users.)
    bool staticFor(const Task &t) const {
        for (auto &m : t.mounts) {
            if (m.type != MountCluster) continue;
            std::vector<std::string> ids;
            if (m.source.compare(0, 6, "group:") == 0) { auto g = by_group.find(m.source.substr(6)); if (g != by_group.end()) ids.assign(g->second.begin(), g->second.end()); }
            else { auto f = by_name.find(m.source); if (f != by_name.end()) ids.push_back(f->second); }
            for (auto &id : ids) {
                const Volume &v = vols.at(id).v;
                if (v.scope != VolScopeMulti) return false;
                if (v.sharing == VolShareNone) return false;
                if (v.sharing == VolShareOneWriter && !m.read_only) return false;
            }
        }
        return true;
    }
};
static bool has_cluster_mounts(const Task &t) { if (!t.has_container) return false; for (auto &m : t.mounts) if (m.type == MountCluster) return true; return false; }

struct Decision { TaskP old_, new_; };

struct Scheduler {
    pe_engine *eng = nullptr;
    std::string fatal;   // engine error, surfaced to the caller (the call failed; nothing was lost: see scheduleTaskGroups)
    std::string unsupported;   // first group of the last tick the engine could not take (its tasks stayed pending)
    std::map<std::string, TaskP> unassignedTasks, pendingPreassignedTasks, allTasks;
    std::set<std::string> preassignedTasks;
    std::map<std::string, NodeInfo> nodeSet;   // ordered: row index = rank of the node ID
    std::map<std::string, std::pair<bool, uint64_t>> services;
    VolumeBook volumes;                  // CSI cluster volumes: host-side bookkeeping (volumes.go)
    uint32_t vol_col = PE_NONE;          // attribute column that names the node set of a group with cluster mounts
    static const uint32_t vol_gen = 1;   // the column's value on the rows of the set (0 = "" on every other row)
    std::set<std::string> vol_in;        // the nodes that carry it, written into the row by encode_row
    // VolumesFilter's answer for the mounts asked about last: the tasks of one service, or a tick's one-off tasks of one
    // spec, ask the same question again and again; it stands until a volume, a counted use or a node changes
    struct VolAnswer { bool valid = false; std::string mounts; uint64_t vol_epoch = 0, node_epoch = 0, id = 0; std::vector<std::string> allowed, excluded; } vol_answer;
    uint64_t node_epoch = 0;             // moves with every node event
    uint64_t vol_in_answer = 0;          // the answer vol_in holds in full (0 = none: a stepwise walk took nodes out)
    bool leaf_also_in_volume_set = false; // a preference group with cluster mounts: every leaf visit carries the volume term too
    bool leaf_stepwise = false;           // ... and, when its volumes count their users, walks the fill loop itself (fillLeafStepwise)
    int64_t now = 0;

    // ---- dictionaries (exact interning)
    std::unordered_map<std::string, uint32_t> val_ids{{"", 0}}, exact_ids{{"", 0}}, svc_ids, kind_ids, label_cols, plugin_slots;
    std::map<std::pair<int, uint32_t>, uint32_t> port_slots;
    // placement preferences (nodeset.go:59-101): one attribute column per spread label, holding EXACT value ids (a branch
    // is a label value as written, not case-folded); pref_strings[id] gives the value back for the canonical branch order
    std::unordered_map<std::string, uint32_t> pref_cols, pref_ids{{"", 0}};
    std::vector<std::string> pref_strings{""};
    uint32_t next_label_col = PE_ATTR_FIRST_LABEL;
    uint64_t rows_uploaded = 0, full_uploads = 0;   // node rows sent with pe_node_upsert; how often the whole table went
    double host_encode_ms = 0, host_engine_ms = 0, host_apply_ms = 0;   // where a tick's host time went (last tick)
    bool layout_dirty = true;            // membership / dictionary change: re-upload every row
    std::set<std::string> dirty_nodes;   // rows whose NodeInfo changed on the host side
    std::vector<std::string> idx_to_id;

    uint32_t intern(std::unordered_map<std::string, uint32_t> &d, const std::string &s) {
        auto it = d.find(s);
        if (it != d.end()) return it->second;
        uint32_t id = (uint32_t)d.size();
        d[s] = id;
        return id;
    }
    uint32_t value_id(const std::string &s) { return intern(val_ids, fold(s)); }
    uint32_t exact_id(const std::string &s) { return intern(exact_ids, s); }
    // Service ids name the engine's per-service counter columns (one dense column per id, never freed by the engine): ids of
    // services nobody counts or references any more are handed out again, so the columns are bounded by the services alive
    // at one time, not by the services ever seen.  (A recycled id's column is all zeros: the device counts mirror the host's.)
    std::vector<uint32_t> free_svc_ids;
    uint32_t next_svc_id = 0;
    size_t svc_recycle_min = 4096, svc_recycle_at = 4096;   // look for dead services once this many ids are in use
    uint32_t svc_id(const std::string &s) {
        auto it = svc_ids.find(s);
        if (it != svc_ids.end()) return it->second;
        uint32_t id;
        if (!free_svc_ids.empty()) { id = free_svc_ids.back(); free_svc_ids.pop_back(); }
        else id = next_svc_id++;
        svc_ids[s] = id;
        return id;
    }
    void recycle_service_ids() {
        if (svc_ids.size() < svc_recycle_at) return;
        std::set<std::string> live;
        for (auto &kv : nodeSet) for (auto &c : kv.second.by_service) if (c.second != 0) live.insert(c.first);
        for (auto &kv : allTasks) if (kv.second) live.insert(kv.second->service);
        for (auto &kv : unassignedTasks) if (kv.second) live.insert(kv.second->service);
        for (auto &kv : pendingPreassignedTasks) if (kv.second) live.insert(kv.second->service);
        for (auto it = svc_ids.begin(); it != svc_ids.end();) {
            if (live.count(it->first)) { ++it; continue; }
            free_svc_ids.push_back(it->second);
            it = svc_ids.erase(it);
        }
        svc_recycle_at = std::max<size_t>(svc_recycle_min, 2 * svc_ids.size());      // (amortised: the sweep is O(nodes + tasks))
    }
    uint32_t kind_id(const std::string &s) { return intern(kind_ids, s); }
    uint32_t label_col(const std::string &prefixed) {   // "n:<key>" node label, "e:<key>" engine label
        auto it = label_cols.find(prefixed);
        if (it != label_cols.end()) return it->second;
        label_cols[prefixed] = next_label_col;
        layout_dirty = true;             // a new column has to be filled for every node
        return next_label_col++;
    }
    uint32_t pref_col(const std::string &prefixed) {    // "n:<key>" node label, "e:<key>" engine label (exact values)
        auto it = pref_cols.find(prefixed);
        if (it != pref_cols.end()) return it->second;
        pref_cols[prefixed] = next_label_col;
        layout_dirty = true;
        return next_label_col++;
    }
    uint32_t pref_id(const std::string &v) {
        auto it = pref_ids.find(v);
        if (it != pref_ids.end()) return it->second;
        uint32_t id = (uint32_t)pref_strings.size();
        pref_ids[v] = id; pref_strings.push_back(v);
        return id;
    }
    uint32_t plugin_slot(const std::string &type, const std::string &name) {
        std::string k = type + "\x1f" + name;
        auto it = plugin_slots.find(k);
        if (it != plugin_slots.end()) return it->second;
        uint32_t s = (uint32_t)plugin_slots.size();
        plugin_slots[k] = s;
        layout_dirty = true;
        return s;
    }
    uint32_t port_slot(int proto, uint32_t port) {
        auto k = std::make_pair(proto, port);
        auto it = port_slots.find(k);
        if (it != port_slots.end()) return it->second;
        uint32_t s = (uint32_t)port_slots.size();
        port_slots[k] = s;
        return s;
    }
    static std::string norm_arch(const std::string &a) { return a == "x86_64" ? "amd64" : a == "aarch64" ? "arm64" : a; }  // filter.go:291-306

    bool check(int32_t rc, const char *what) {
        if (rc == PE_OK) return true;
        fatal = std::string(what) + ": " + pe_last_error(eng);
        return false;
    }

    // ---- node rows ------------------------------------------------------------
    uint32_t node_index(const std::string &id) const {
        auto it = std::lower_bound(idx_to_id.begin(), idx_to_id.end(), id);
        return (it != idx_to_id.end() && *it == id) ? (uint32_t)(it - idx_to_id.begin()) : PE_NONE;
    }
    struct RowBatch {
        std::vector<pe_node_row> rows; std::vector<pe_kv32> attrs, svcs; std::vector<pe_kv64> gens; std::vector<uint32_t> ports, plugs;
    };
    void encode_row(uint32_t idx, const NodeInfo &ni, RowBatch &b) {
        const Node &n = *ni.node;
        pe_node_row r{};
        r.node_idx = idx;
        r.flags = PE_NODE_VALID;
        if (n.state == NodeReady && n.availability == AvailActive) r.flags |= PE_NODE_READY;       // filter.go:40-43
        if (n.has_desc && n.has_platform) { r.flags |= PE_NODE_HAS_PLATFORM; r.os_id = exact_id(n.os); r.arch_id = exact_id(norm_arch(n.arch)); }
        if (n.has_desc && n.has_engine) r.flags |= PE_NODE_HAS_ENGINE;
        Addr a = parse_ip(n.addr);
        if (a.ok) { r.flags |= PE_NODE_IP_VALID | (a.v4 ? PE_NODE_IP_V4 : 0); memcpy(r.ip, a.w, sizeof r.ip); }
        r.cpu_avail = ni.avail.cpu; r.mem_avail = ni.avail.mem;
        r.total_tasks = (uint32_t)ni.active;
        r.attr_off = (uint32_t)b.attrs.size();
        auto attr = [&](uint32_t col, const std::string &s) { uint32_t v = value_id(s); if (v) b.attrs.push_back({col, v}); };
        attr(PE_ATTR_NODE_ID, n.id);                                                               // constraint.go:110
        attr(PE_ATTR_HOSTNAME, n.has_desc ? n.hostname : "");                                      // :114-125
        attr(PE_ATTR_ROLE, n.role == 1 ? "MANAGER" : "WORKER");                                    // :147-150 (observed role)
        attr(PE_ATTR_OS, (n.has_desc && n.has_platform) ? n.os : "");                              // :151-160
        attr(PE_ATTR_ARCH, (n.has_desc && n.has_platform) ? n.arch : "");                          // :161-170 (not normalised)
        for (auto &lc : label_cols) {
            const std::string key = lc.first.substr(2);
            const std::map<std::string, std::string> *m = nullptr;
            if (lc.first[0] == 'n') { if (n.has_labels) m = &n.labels; }
            else if (n.has_desc && n.has_engine && n.has_elabels) m = &n.elabels;
            if (!m) continue;
            auto f = m->find(key);
            if (f != m->end()) attr(lc.second, f->second);
        }
        if (vol_col != PE_NONE && vol_in.count(n.id)) b.attrs.push_back({vol_col, vol_gen});
        for (auto &pc : pref_cols) {                                                               // nodeset.go:69-82
            const std::string key = pc.first.substr(2);
            const std::map<std::string, std::string> *m = nullptr;
            if (pc.first[0] == 'n') { if (n.has_labels) m = &n.labels; }
            else if (n.has_desc && n.has_engine && n.has_elabels) m = &n.elabels;
            if (!m) continue;
            auto f = m->find(key);
            if (f != m->end()) { uint32_t v = pref_id(f->second); if (v) b.attrs.push_back({pc.second, v}); }
        }
        r.attr_cnt = (uint32_t)b.attrs.size() - r.attr_off;
        r.gen_off = (uint32_t)b.gens.size();
        std::set<std::string> seen;
        for (auto &g : ni.avail.generic) {
            if (!seen.insert(g.kind).second) continue;
            int64_t cnt = 0;
            if (!g.named) cnt = g.amount; else for (auto &x : ni.avail.generic) if (x.kind == g.kind) cnt++;   // validate.go:36-48
            b.gens.push_back({kind_id(g.kind), 0, PE_GEN_ENCODE(cnt, g.named ? PE_GEN_NAMED : PE_GEN_DISCRETE)});
        }
        r.gen_cnt = (uint32_t)b.gens.size() - r.gen_off;
        r.svc_off = (uint32_t)b.svcs.size();
        for (auto &kv : ni.by_service) if (kv.second) b.svcs.push_back({svc_id(kv.first), (uint32_t)kv.second});
        r.svc_cnt = (uint32_t)b.svcs.size() - r.svc_off;
        r.port_off = (uint32_t)b.ports.size();
        for (auto &p : ni.ports) b.ports.push_back(port_slot(p.first, p.second));
        r.port_cnt = (uint32_t)b.ports.size() - r.port_off;
        r.plug_off = (uint32_t)b.plugs.size();
        if (n.has_desc && n.has_engine) {
            for (auto &ps : plugin_slots) {
                size_t sep = ps.first.find('\x1f');
                const std::string type = ps.first.substr(0, sep), name = ps.first.substr(sep + 1);
                for (auto &np : n.plugins) {                                                       // filter.go:186-205
                    if (np.type != type) continue;
                    if (np.name == name || (np.name.size() == name.size() + 7 && np.name.compare(0, name.size(), name) == 0 && np.name.substr(name.size()) == ":latest")) { b.plugs.push_back(ps.second); break; }
                }
            }
            for (auto &np : n.plugins) if (np.type == "Log") { r.flags |= PE_NODE_HAS_LOGPLUGIN; break; }
        }
        r.plug_cnt = (uint32_t)b.plugs.size() - r.plug_off;
        b.rows.push_back(r);
    }
    bool flush_rows() {
        if (!layout_dirty && dirty_nodes.empty()) return true;
        RowBatch b;
        if (layout_dirty) {
            // dictionaries may grow while encoding (label columns / plugin slots are fixed by now; value ids are append-only)
            idx_to_id.clear();
            for (auto &kv : nodeSet) idx_to_id.push_back(kv.first);
            uint32_t i = 0;
            for (auto &kv : nodeSet) encode_row(i++, kv.second, b);
            if (!check(pe_set_node_count(eng, (uint32_t)nodeSet.size()), "pe_set_node_count")) return false;
        } else {
            for (auto &id : dirty_nodes) {
                uint32_t idx = node_index(id);
                auto it = nodeSet.find(id);
                if (idx != PE_NONE && it != nodeSet.end()) encode_row(idx, it->second, b);
            }
        }
        if (layout_dirty) full_uploads++;
        rows_uploaded += b.rows.size();
        layout_dirty = false;
        dirty_nodes.clear();
        if (b.rows.empty()) return true;
        return check(pe_node_upsert(eng, b.rows.data(), (uint32_t)b.rows.size(), b.attrs.data(), b.gens.data(), b.svcs.data(), b.ports.data(), b.plugs.data()),
                     "pe_node_upsert");
    }
    void touch(const std::string &id) { dirty_nodes.insert(id); }

    // ---- store events -----------------------------------------------------------
    void enqueue(const TaskP &t) { unassignedTasks[t->id] = t; }
    void setupTasksList(const std::vector<NodeP> &nodes, const std::vector<TaskP> &tasks, const std::vector<Volume> &vols = {}) {   // scheduler.go:68-125 + buildNodeSet :973-990
        for (auto &v : vols) if (v.has_info && !v.volume_id.empty()) volumes.addOrUpdate(v);      // only volumes created with their plugin (:75-81)
        std::map<std::string, std::vector<TaskP>> byNode;
        for (auto &t : tasks) {
            if (t->state < TaskStatePending || t->state > TaskStateRunning) continue;
            if (t->state == TaskStatePending && t->desired > TaskStateCompleted) continue;
            allTasks[t->id] = t;
            if (t->node_id.empty()) { enqueue(t); continue; }
            if (t->state == TaskStatePending) { preassignedTasks.insert(t->id); pendingPreassignedTasks[t->id] = t; continue; }
            volumes.reserveTask(*t);                 // track the volumes in use by the task (:115-116)
            byNode[t->node_id].push_back(t);
        }
        for (auto &n : nodes) {
            NodeInfo ni; ni.node = n; ni.last_cleanup = now;
            if (n->has_desc && n->has_resources) ni.avail = n->resources;
            for (auto &t : byNode[n->id]) ni.addTask(t);
            nodeSet[n->id] = ni;
        }
        layout_dirty = true;
    }
    void createTask(const TaskP &t) {   // scheduler.go:254-281
        if (t->state < TaskStatePending || t->state > TaskStateRunning) return;
        allTasks[t->id] = t;
        if (t->node_id.empty()) { enqueue(t); return; }
        if (t->state == TaskStatePending) { preassignedTasks.insert(t->id); pendingPreassignedTasks[t->id] = t; return; }
        auto it = nodeSet.find(t->node_id);
        if (it != nodeSet.end() && it->second.addTask(t)) touch(t->node_id);
    }
    void updateVolume(const Volume &v) { if (v.has_info && !v.volume_id.empty()) volumes.addOrUpdate(v); }   // scheduler.go:205-217
    void deleteTask(const Task &t) {   // scheduler.go:350-366
        allTasks.erase(t.id); preassignedTasks.erase(t.id); pendingPreassignedTasks.erase(t.id);
        for (auto &va : t.volumes) volumes.release(va.id, t.id);
        auto it = nodeSet.find(t.node_id);
        if (it != nodeSet.end() && it->second.removeTask(t)) touch(t.node_id);
    }
    void updateTask(const TaskP &t) {   // scheduler.go:283-348
        if (t->state < TaskStatePending) return;
        TaskP old; auto o = allTasks.find(t->id); if (o != allTasks.end()) old = o->second;
        if (t->state > TaskStateRunning) {
            if (!old) return;
            if (t->state != old->state && (t->state == TaskStateFailed || t->state == TaskStateRejected) && !preassignedTasks.count(t->id)) {
                auto it = nodeSet.find(t->node_id);
                if (it != nodeSet.end()) it->second.taskFailed(*t, now);
            }
            deleteTask(*old);
            return;
        }
        if (t->node_id.empty()) { if (old) deleteTask(*old); allTasks[t->id] = t; enqueue(t); return; }
        if (t->state == TaskStatePending) {
            if (old) deleteTask(*old);
            preassignedTasks.insert(t->id); allTasks[t->id] = t; pendingPreassignedTasks[t->id] = t; return;
        }
        allTasks[t->id] = t;
        auto it = nodeSet.find(t->node_id);
        if (it != nodeSet.end() && it->second.addTask(t)) touch(t->node_id);
    }
    void createOrUpdateNode(const NodeP &n) {   // scheduler.go:368-396
        auto it = nodeSet.find(n->id);
        Resources res;
        if (n->has_desc && n->has_resources) {
            res = n->resources;
            if (it != nodeSet.end())
                for (auto &kv : it->second.tasks) {
                    Resources r = NodeInfo::reservations(*kv.second);
                    res.mem -= r.mem; res.cpu -= r.cpu;
                    gr::consume(res.generic, kv.second->assigned);
                }
        }
        node_epoch++;
        if (it == nodeSet.end()) { NodeInfo ni; ni.node = n; ni.avail = res; ni.last_cleanup = now; nodeSet[n->id] = ni; layout_dirty = true; }
        else { it->second.node = n; it->second.avail = res; touch(n->id); }
    }
    void removeNode(const std::string &id) { node_epoch++; if (nodeSet.erase(id)) layout_dirty = true; vol_in.erase(id); }   // nodeset.go:46-48

    // ---- group descriptors <- the filters' SetTask (filter.go) --------------------
    struct TickBuf {
        std::vector<pe_group> groups; std::vector<uint8_t> flags; std::vector<pe_generic_want> gens; std::vector<pe_constraint> cons;
        std::vector<pe_ip_constraint> ips; std::vector<pe_platform> plats; std::vector<uint32_t> ports, plugs; std::vector<pe_node_fail> fails;
        // countRecentFailures(now) per node, built ONCE per tick from the nodes that hold failure records at all and
        // shared by every group of the same (service, spec version): (offset, count) into `fails`
        bool fails_built = false;
        std::map<SvcVer, std::pair<uint32_t, uint32_t>> fail_range;
        // an encoder that gives up half-way must leave no side arrays behind: remember the sizes, cut back to them
        // (the failure lists are built once per tick and stay)
        struct Mark { size_t groups, flags, gens, cons, ips, plats, ports, plugs; };
        Mark mark() const { return {groups.size(), flags.size(), gens.size(), cons.size(), ips.size(), plats.size(), ports.size(), plugs.size()}; }
        void rollback(const Mark &m) {
            groups.resize(m.groups); flags.resize(m.flags); gens.resize(m.gens); cons.resize(m.cons); ips.resize(m.ips);
            plats.resize(m.plats); ports.resize(m.ports); plugs.resize(m.plugs);
        }
        pe_tick view() const {
            pe_tick t{};
            t.groups = groups.data(); t.n_groups = (uint32_t)groups.size();
            t.task_flags = flags.data(); t.n_tasks = (uint32_t)flags.size();
            t.gens = gens.data(); t.n_gens = (uint32_t)gens.size();
            t.cons = cons.data(); t.n_cons = (uint32_t)cons.size();
            t.ips = ips.data(); t.n_ips = (uint32_t)ips.size();
            t.plats = plats.data(); t.n_plats = (uint32_t)plats.size();
            t.ports = ports.data(); t.n_ports = (uint32_t)ports.size();
            t.plugs = plugs.data(); t.n_plugs = (uint32_t)plugs.size();
            t.fails = fails.data(); t.n_fails = (uint32_t)fails.size();
            return t;
        }
    };
    // constraint.Parse (constraint.go:40-81) + key dispatch of NodeMatches (:107-207), compiled to integer programs
    // (one-off tasks of one service carry the same expressions: a tick of 1M tasks parses ~1000 distinct lists, not 1M.
    // The compiled form only holds ids that never change once handed out -- columns, value ids -- so it stays valid.)
    struct Compiled { bool ok = false, never = false; std::vector<pe_constraint> cons; std::vector<pe_ip_constraint> ips; };
    std::unordered_map<std::string, Compiled> compiled_cache;
    bool compile_constraints(const std::vector<std::string> &env, pe_group &g, TickBuf &b) {
        std::string key;
        for (auto &e : env) { key += e; key += '\x1f'; }
        auto hit = compiled_cache.find(key);
        if (hit == compiled_cache.end()) {
            if (compiled_cache.size() > (1u << 16)) compiled_cache.clear();      // (bounded; a miss only costs the parse)
            Compiled c;
            c.ok = parse_constraints(env, c);
            hit = compiled_cache.emplace(std::move(key), std::move(c)).first;
        }
        const Compiled &c = hit->second;
        if (!c.ok) return false;
        g.con_off = (uint32_t)b.cons.size(); g.con_cnt = (uint32_t)c.cons.size();
        b.cons.insert(b.cons.end(), c.cons.begin(), c.cons.end());
        g.ip_off = (uint32_t)b.ips.size(); g.ip_cnt = (uint32_t)c.ips.size();
        b.ips.insert(b.ips.end(), c.ips.begin(), c.ips.end());
        if (c.never) g.flags |= PE_G_CONSTRAINT_NEVER;
        return true;
    }
    bool parse_constraints(const std::vector<std::string> &env, Compiled &out) {
        auto alpha = [](unsigned char c) { return c >= 'a' && c <= 'z'; };
        std::vector<pe_constraint> cons; std::vector<pe_ip_constraint> ips; bool never = false;
        for (auto &e : env) {
            size_t at = e.find("=="); int op = 0;
            if (at == std::string::npos) { at = e.find("!="); op = 1; }
            if (at == std::string::npos) return false;
            std::string key = trim(e.substr(0, at)), val = trim(e.substr(at + 2));
            std::string fk = fold(key), fv = fold(val);
            if (fk.size() < 2 || !(alpha((unsigned char)fk[0]) || fk[0] == '_')) return false;          // alphaNumeric
            for (size_t i = 1; i < fk.size(); i++) { unsigned char c = (unsigned char)fk[i]; if (!(alpha(c) || (c >= '0' && c <= '9') || c == '-' || c == '_' || c == '.')) return false; }
            if (fv.empty()) return false;                                                                // valuePattern
            for (unsigned char c : fv) if (!(alpha(c) || (c >= '0' && c <= '9') || (c && strchr(":-_.*()?+[]\\^$|/", c)) || c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r')) return false;
            uint32_t col = PE_NONE;
            if (fk == "node.id") col = PE_ATTR_NODE_ID;
            else if (fk == "node.hostname") col = PE_ATTR_HOSTNAME;
            else if (fk == "node.role") col = PE_ATTR_ROLE;
            else if (fk == "node.platform.os") col = PE_ATTR_OS;
            else if (fk == "node.platform.arch") col = PE_ATTR_ARCH;
            else if (fk == "node.ip") {
                pe_ip_constraint ic{}; ic.neq = (uint32_t)op;
                Addr a = parse_ip(val);
                if (a.ok) { memcpy(ic.net, a.w, 16); for (auto &m : ic.mask) m = 0xFFFFFFFFu; ips.push_back(ic); continue; }
                size_t sl = val.find('/');
                bool good = false;
                if (sl != std::string::npos) {
                    std::string ad = val.substr(0, sl), bits = val.substr(sl + 1);
                    Addr n = parse_ip(ad);
                    bool v4 = ad.find(':') == std::string::npos;
                    int nb = -1;
                    if (!bits.empty() && bits.size() <= 3 && !(bits.size() > 1 && bits[0] == '0')) { nb = 0; for (char c : bits) { if (c < '0' || c > '9') { nb = -1; break; } nb = nb * 10 + (c - '0'); } }
                    int len = v4 ? 32 : 128;
                    if (n.ok && nb >= 0 && nb <= len) {
                        int total = v4 ? 96 + nb : nb;
                        for (int w = 0; w < 4; w++) { int r = total - 32 * w; ic.mask[w] = r >= 32 ? 0xFFFFFFFFu : r <= 0 ? 0u : 0xFFFFFFFFu << (32 - r); ic.net[w] = n.w[w] & ic.mask[w]; }
                        ic.is_cidr = 1; ic.is_v4 = v4 ? 1 : 0; good = true;
                    }
                }
                if (good) ips.push_back(ic); else never = true;                                         // constraint.go:144-146
                continue;
            }
            else if (fk.size() > 12 && fk.compare(0, 12, "node.labels.") == 0) col = label_col("n:" + key.substr(key.size() - (fk.size() - 12)));
            else if (fk.size() > 14 && fk.compare(0, 14, "engine.labels.") == 0) col = label_col("e:" + key.substr(key.size() - (fk.size() - 14)));
            else { never = true; continue; }                                                            // constraint.go:200-203
            cons.push_back({col, value_id(val), (uint32_t)op});
        }
        out.cons = std::move(cons); out.ips = std::move(ips); out.never = never;
        return true;
    }
    bool encode_group(const std::vector<TaskP> &tasks, TickBuf &b, bool for_fit = false) {
        const Task &t = *tasks[0];
        // (placement preferences: the caller cuts the group into leaf visits, see schedulePreferenceGroup; taskFitNode
        // checks one named node and preferences play no part there, scheduler.go:646-690)
        (void)for_fit;
        // (cluster mounts: VolumesFilter is evaluated by the shim, which hands the engine the node set: scheduleVolumeGroup)
        pe_group g{};
        g.log_plugin = PE_NONE;
        g.svc_id = svc_id(t.service);
        g.n_tasks = (uint32_t)tasks.size();
        g.task_off = (uint32_t)b.flags.size();
        for (auto &x : tasks) b.flags.push_back(x->desired <= TaskStateCompleted ? PE_T_COUNTS : 0);
        g.filter_mask = 1u << PE_F_READY;                                                               // filter.go:35
        Resources r = NodeInfo::reservations(t);
        g.cpu_res = r.cpu; g.mem_res = r.mem;
        g.gen_off = (uint32_t)b.gens.size();
        for (auto &w : r.generic) {
            if (w.named) { fatal = "task " + t.id + " reserves a named generic resource"; return false; }
            b.gens.push_back({kind_id(w.kind), 0, w.amount});
        }
        g.gen_cnt = (uint32_t)b.gens.size() - g.gen_off;
        if (t.has_res && (r.cpu || r.mem || !r.generic.empty())) g.filter_mask |= 1u << PE_F_RESOURCE;    // filter.go:60-73
        // PluginFilter.SetTask, filter.go:118-139
        bool volT = false;
        for (auto &m : t.mounts) if (m.type == MountVolume && m.has_driver && !m.driver.empty() && m.driver != "local") volT = true;
        if ((t.has_container && volT) || !t.networks.empty() || t.has_log) {
            g.filter_mask |= 1u << PE_F_PLUGIN;
            g.plug_off = (uint32_t)b.plugs.size();
            for (auto &m : t.mounts) if (m.type == MountVolume && m.has_driver && !m.driver.empty() && m.driver != "local") b.plugs.push_back(plugin_slot("Volume", m.driver));
            for (auto &nw : t.networks) if (nw.first && !nw.second.empty()) b.plugs.push_back(plugin_slot("Network", nw.second));
            g.plug_cnt = (uint32_t)b.plugs.size() - g.plug_off;
            if (t.has_log && t.log_driver != "none" && !t.log_driver.empty()) { g.flags |= PE_G_LOG_DRIVER; g.log_plugin = plugin_slot("Log", t.log_driver); }
        }
        if (t.has_placement && !t.constraints.empty()) {                                                // filter.go:224-238
            pe_group probe = g;
            if (compile_constraints(t.constraints, probe, b)) { g = probe; g.filter_mask |= 1u << PE_F_CONSTRAINT; }
        }
        if (t.has_placement && !t.platforms.empty()) {                                                  // filter.go:259-269
            g.filter_mask |= 1u << PE_F_PLATFORM;
            g.plat_off = (uint32_t)b.plats.size();
            for (auto &p : t.platforms) b.plats.push_back({p.second.empty() ? 0u : exact_id(p.second), p.first.empty() ? 0u : exact_id(norm_arch(p.first))});
            g.plat_cnt = (uint32_t)b.plats.size() - g.plat_off;
        }
        g.port_off = (uint32_t)b.ports.size();
        if (t.has_endpoint) for (auto &p : t.ports) if (p.mode == PublishHost && p.port) b.ports.push_back(port_slot(p.protocol, p.port));
        g.port_cnt = (uint32_t)b.ports.size() - g.port_off;
        if (g.port_cnt) g.filter_mask |= 1u << PE_F_HOSTPORT;                                           // filter.go:328-339
        if (t.has_placement && t.max_replicas > 0) { g.filter_mask |= 1u << PE_F_MAXREPLICAS; g.max_replicas = t.max_replicas; }   // filter.go:369-376
        // countRecentFailures per node (scheduler.go:711-712); only nodes that have an entry
        if (!b.fails_built) build_fail_lists(b);
        auto fr = b.fail_range.find(SvcVer{t.service, t.has_spec_version ? t.spec_version : 0});
        if (fr != b.fail_range.end()) { g.fail_off = fr->second.first; g.fail_cnt = fr->second.second; }
        b.groups.push_back(g);
        return true;
    }
    // One walk over the node set per tick (not per group): nodes are visited in row order, so every key's list comes out
    // sorted by node index, which is what pe_node_fail asks for.
    void build_fail_lists(TickBuf &b) {
        std::map<SvcVer, std::vector<pe_node_fail>> per_key;
        uint32_t idx = 0;
        for (auto &kv : nodeSet) {
            for (auto &f : kv.second.failures) {
                int c = kv.second.countRecentFailures(now, f.first);
                if (c > 0) per_key[f.first].push_back({idx, (uint32_t)c});
            }
            idx++;
        }
        for (auto &kv : per_key) {
            b.fail_range[kv.first] = {(uint32_t)b.fails.size(), (uint32_t)kv.second.size()};
            b.fails.insert(b.fails.end(), kv.second.begin(), kv.second.end());
        }
        b.fails_built = true;
    }

    // Pipeline.Explain, pipeline.go:84-103 + filter.go Explain methods
    static std::string explain(const uint32_t *cnt) {
        static const char *one[8] = {"1 node not available for new tasks", "insufficient resources on 1 node", "missing plugin on 1 node",
                                     "scheduling constraints not satisfied on 1 node", "unsupported platform on 1 node",
                                     "host-mode port already in use on 1 node", "max replicas per node limit exceed",
                                     "cannot fulfill requested CSI volume mounts on 1 node"};
        static const char *many[8] = {"%u nodes not available for new tasks", "insufficient resources on %u nodes", "missing plugin on %u nodes",
                                      "scheduling constraints not satisfied on %u nodes", "unsupported platform on %u nodes",
                                      "host-mode port already in use on %u nodes", "max replicas per node limit exceed",
                                      "cannot fulfill requested CSI volume mounts on %u nodes"};
        int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        for (int i = 1; i < 8; i++) for (int j = i; j > 0 && cnt[order[j]] > cnt[order[j - 1]]; j--) std::swap(order[j], order[j - 1]);
        std::string out;
        for (int f : order) {
            if (!cnt[f]) continue;
            if (!out.empty()) out += "; ";
            char buf[96]; snprintf(buf, sizeof buf, cnt[f] == 1 ? one[f] : many[f], cnt[f]);
            out += buf;
        }
        return out;
    }

    // scheduleTaskGroup for every group of the tick in ONE engine call (scheduler.go:464-469,694-748)
    // A group the engine cannot take (see encode_group) does not stop the tick: its tasks stay pending with the reason in
    // Status.Err, exactly like tasks without a suitable node (scheduler.go:958-962), and every other group is scheduled.
    // If the engine itself fails, every task goes back to the queue and the device mirror is rebuilt from the host's
    // NodeInfo on the next call (the device rows may have moved part-way).
    // The spread levels of a group's placement preferences (nodeset.go:59-82): attribute columns of exact label values.
    // Descriptors that are neither node.labels.* nor engine.labels.* contribute no level.
    std::vector<uint32_t> preference_levels(const Task &t) {
        std::vector<uint32_t> cols;
        if (!t.has_placement) return cols;
        for (auto &d : t.preferences) {
            if (d.size() > 12 && fold(d.substr(0, 12)) == "node.labels.") cols.push_back(pref_col("n:" + d.substr(12)));
            else if (d.size() > 14 && fold(d.substr(0, 14)) == "engine.labels.") cols.push_back(pref_col("e:" + d.substr(14)));
        }
        return cols;
    }

    // what the engine decided for one encoded group: decisions, host mirror; tasks without a node come back in `left`
    void apply_group(const std::vector<TaskP> &grp, const uint32_t *out_node, std::vector<TaskP> &left, std::map<std::string, Decision> &decisions) {
        const bool csi = !grp.empty() && has_cluster_mounts(*grp[0]);
        for (size_t i = 0; i < grp.size(); i++) {
            const TaskP &t = grp[i];
            uint32_t idx = out_node[i];
            if (idx == PE_NONE || idx >= idx_to_id.size()) { left.push_back(t); continue; }
            if (csi) { assign_with_volumes(t, idx_to_id[idx], decisions); continue; }  // (attachments chosen and reserved)
            TaskP nt(new Task(*t));                                                   // scheduler.go:871-880
            nt->node_id = idx_to_id[idx];
            nt->state = TaskStateAssigned; nt->err.clear(); nt->message = "scheduler assigned task to node";
            allTasks[t->id] = nt;
            nodeSet[nt->node_id].addTask(nt);   // host mirror: counters, named generic members (the device row already moved)
            decisions[t->id] = {t, nt};
        }
    }

    // scheduleTaskGroup for every group of the tick (scheduler.go:464-469,694-748): maximal runs of groups without
    // placement preferences go to the engine in ONE call; a group with preferences is walked leaf by leaf in between.
    // A group the engine cannot take (see encode_group) does not stop the tick: its tasks stay pending with the reason in
    // Status.Err, exactly like tasks without a suitable node (scheduler.go:958-962), and every other group is scheduled.
    // If the engine itself fails, every task not yet decided goes back to the queue and the device mirror is rebuilt from
    // the host's NodeInfo on the next call (the device rows may have moved part-way).
    bool scheduleTaskGroups(std::vector<std::vector<TaskP>> &all_groups, std::map<std::string, Decision> &decisions) {
        unsupported.clear();
        std::vector<std::vector<TaskP>> run;
        for (size_t gi = 0; gi < all_groups.size(); gi++) {
            auto &g = all_groups[gi];
            const bool pref = !g.empty() && !preference_levels(*g[0]).empty();
            const bool csi = !g.empty() && has_cluster_mounts(*g[0]);
            if (!pref && !csi) { run.push_back(g); continue; }
            bool ok = scheduleRun(run, decisions);
            run.clear();
            if (ok) ok = csi ? scheduleVolumeGroup(g, decisions) : schedulePreferenceGroup(g, decisions);
            if (!ok) {
                for (size_t r = gi + 1; r < all_groups.size(); r++) for (auto &t : all_groups[r]) enqueue(t);
                return false;
            }
        }
        return scheduleRun(run, decisions);
    }

    bool scheduleRun(std::vector<std::vector<TaskP>> &all_groups, std::map<std::string, Decision> &decisions) {
        if (all_groups.empty()) return true;
        const auto tp0 = std::chrono::steady_clock::now();
        auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
        TickBuf b;
        std::vector<std::vector<TaskP>> groups;
        for (auto &g : all_groups) {
            const TickBuf::Mark m = b.mark();       // (copying the buffer per group, as round 1 did, is quadratic in the tick)
            fatal.clear();
            if (encode_group(g, b)) { groups.push_back(g); continue; }
            b.rollback(m);
            if (unsupported.empty()) unsupported = fatal;
            noSuitableNode(g, "unsupported by the placement engine: " + fatal, decisions);
        }
        fatal.clear();
        if (groups.empty()) return true;
        auto give_back = [&]() {
            for (auto &g : groups) for (auto &t : g) enqueue(t);
            layout_dirty = true;
        };
        if (!flush_rows()) { give_back(); return false; }     // after encoding: new label columns / plugin slots force a re-upload
        std::vector<uint32_t> out_node(b.flags.size(), PE_NONE), out_fail(b.groups.size() * PE_NUM_FILTERS, 0);
        pe_tick tk = b.view();
        host_encode_ms += since(tp0);
        const auto tp1 = std::chrono::steady_clock::now();
        if (!check(pe_schedule(eng, &tk, out_node.data(), out_fail.data()), "pe_schedule")) { give_back(); return false; }
        host_engine_ms += since(tp1);
        const auto tp2 = std::chrono::steady_clock::now();
        for (size_t gi = 0; gi < groups.size(); gi++) {
            std::vector<TaskP> left;
            apply_group(groups[gi], &out_node[b.groups[gi].task_off], left, decisions);
            if (!left.empty()) noSuitableNode(left, explain(&out_fail[gi * PE_NUM_FILTERS]), decisions);
        }
        host_apply_ms += since(tp2);
        return true;
    }

    // ---- cluster (CSI) volumes: VolumesFilter (filter.go:388-447) is evaluated HERE, on the host, against the shim's own
    // volume bookkeeping; the engine gets the answer as the group's node set -- the rows marked in one attribute column,
    // named by a leaf term (the mechanism of the preference leaves) -- and does everything else.  Exact for a single task and
    // as ONE engine group for groups whose volumes cannot change availability while the group is placed
    // (VolumeBook::staticFor); a group whose volume availability moves with every placement (scope SINGLE_NODE, sharing
    // NONE / ONE_WRITER writers) goes through scheduleVolumeGroupStepwise, because the reference re-evaluates the filter
    // inside its fill loop (scheduler.go:912-920).  Left-over tasks get the reference's exact explanation.
    void mark_volume_nodes(const std::vector<std::string> &ids) {       // ids ascending (built in nodeSet order)
        if (vol_col == PE_NONE) { vol_col = next_label_col++; layout_dirty = true; }
        // only the rows whose membership changes are uploaded: consecutive groups with the same mounts (the tasks of one
        // service, a tick's one-off tasks of one spec) mostly ask for the set that is already marked
        std::set<std::string> next(ids.begin(), ids.end());
        for (auto it = vol_in.begin(); it != vol_in.end();) { if (!next.count(*it)) { touch(*it); it = vol_in.erase(it); } else ++it; }
        for (auto &id : ids) if (vol_in.insert(id).second) touch(id);
    }
    // taskFitNode's question for one task of this spec on each of `ids` (batches of pe_fit requests): ok[i] as pe_fit gives
    // it, fail[i*8..] the first failing filter.  A success RESERVES on the device; with `undo` those reservations are
    // taken back (the rows are uploaded again from NodeInfo before the next engine call).
    bool fit_many(const TaskP &t, const std::vector<std::string> &ids, std::vector<uint8_t> &ok, std::vector<uint32_t> &fail, bool undo) {
        const size_t kBatch = 4096;
        ok.assign(ids.size(), 0); fail.assign(ids.size() * PE_NUM_FILTERS, 0);
        for (size_t lo = 0; lo < ids.size(); lo += kBatch) {
            const size_t n = std::min(kBatch, ids.size() - lo);
            TickBuf b;
            std::vector<TaskP> one{t};
            for (size_t i = 0; i < n; i++) if (!encode_group(one, b, true)) return false;
            if (!flush_rows()) return false;
            std::vector<uint32_t> idx(n);
            for (size_t i = 0; i < n; i++) idx[i] = node_index(ids[lo + i]);
            pe_tick tk = b.view();
            if (!check(pe_fit(eng, &tk, idx.data(), ok.data() + lo, fail.data() + lo * PE_NUM_FILTERS), "pe_fit")) return false;
            if (undo) for (size_t i = 0; i < n; i++) if (ok[lo + i] == 1) touch(ids[lo + i]);
        }
        return true;
    }
    // first failing filter of every node of `ids`, none of which passes VolumesFilter: a device filter if one fails,
    // VolumesFilter (the last of the pipeline) otherwise
    bool count_excluded(const TaskP &t, const std::vector<std::string> &ids, std::vector<uint32_t> &cnt) {
        std::vector<uint8_t> ok; std::vector<uint32_t> fail;
        if (!fit_many(t, ids, ok, fail, true)) return false;
        for (size_t i = 0; i < ids.size(); i++) {
            if (ok[i] == 1) cnt[PE_F_VOLUMES]++;
            else if (ok[i] == 0) for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] += fail[i * PE_NUM_FILTERS + f];
        }
        return true;
    }
    // scheduler.go:857-893 for one task: attachments chosen and reserved, NodeInfo moved (the device row already has)
    void assign_with_volumes(const TaskP &ti, const std::string &node_id, std::map<std::string, Decision> &decisions) {
        NodeInfo &ni = nodeSet[node_id];
        TaskP nt(new Task(*ti));
        std::string verr;
        volumes.choose(*ti, *ni.node, nt->volumes, verr);                         // (a failure is only logged there)
        nt->node_id = node_id;
        volumes.reserveTask(*nt);
        nt->state = TaskStateAssigned; nt->err.clear(); nt->message = "scheduler assigned task to node";
        allTasks[ti->id] = nt;
        ni.addTask(nt);
        decisions[ti->id] = {ti, nt};
    }
    // A group whose volume availability moves with every placement (scope SINGLE_NODE, sharing NONE, writers on ONE_WRITER):
    // the reference re-runs the whole pipeline, VolumesFilter included, on the next node before every further placement
    // (scheduler.go:912-920).  The shim walks that loop itself and asks the engine one question at a time:
    //   * the heap of the k best feasible nodes in rank order (scheduler.go:722-731, nodeset.go:103-121) = k successive
    //     one-task groups on the allowed set, each answer taken out of the set and its reservation taken back;
    //   * "does this node still pass, and if so place the next task there" = pe_fit (device filters + reservation), then
    //     VolumesFilter here -- the pipeline's order, so the failure counters come out as the reference's;
    //   * nodeLess between two nodes of the heap (scheduler.go:708-734) is three integer compares on NodeInfo.
    // O(k) engine calls: for the handful of replicas a service with its own volumes has, not a throughput path.
    // nodeLess (scheduler.go:708-734) + the canonical tie-break, as a key: recent failures count only from maxFailures up
    struct RankKey {
        int f, s, a; std::string id;
        bool less_no_tie(const RankKey &o) const { if (f != o.f) return f < o.f; if (s != o.s) return s < o.s; return a < o.a; }
        bool operator<(const RankKey &o) const { if (less_no_tie(o)) return true; if (o.less_no_tie(*this)) return false; return id < o.id; }
    };
    RankKey rank_key(const NodeInfo &n, const Task &t) const {
        const int f = n.countRecentFailures(now, SvcVer{t.service, t.has_spec_version ? t.spec_version : 0});
        auto sv = n.by_service.find(t.service);
        return RankKey{f >= kMaxFailures ? f : 0, sv == n.by_service.end() ? 0 : sv->second, n.active, n.node->id};
    }
    // The `want` best feasible nodes of (leaf terms AND the volume set), in rank order: successive one-task groups, each
    // answer taken out of the volume set and its reservation taken back (the row goes up again before the next call).
    // 0 = engine failure, 1 = ok, 2 = the group cannot be encoded (reason in `fatal`).  first_fail: the counters of the
    // first question when it found no node.
    int probe_best(const TaskP &t, const std::vector<pe_constraint> &leaf_terms, size_t want, std::vector<std::string> &cand, std::vector<uint32_t> *first_fail) {
        TickBuf b;
        fatal.clear();
        std::vector<TaskP> one{t};
        if (!encode_group(one, b)) return 2;
        pe_group &g = b.groups.back();
        if (g.con_cnt == 0) g.con_off = (uint32_t)b.cons.size();
        b.cons.insert(b.cons.end(), leaf_terms.begin(), leaf_terms.end());
        b.cons.push_back({vol_col, vol_gen, 0});
        g.leaf_cnt = (uint32_t)leaf_terms.size() + 1;
        for (size_t i = 0; i < want; i++) {
            if (!flush_rows()) return 0;
            uint32_t pn = PE_NONE; std::vector<uint32_t> pf(PE_NUM_FILTERS, 0);
            pe_tick tk = b.view();
            if (!check(pe_schedule(eng, &tk, &pn, pf.data()), "pe_schedule")) return 0;
            if (pn == PE_NONE || pn >= idx_to_id.size()) { if (i == 0 && first_fail) *first_fail = pf; break; }
            cand.push_back(idx_to_id[pn]);
            vol_in.erase(idx_to_id[pn]);
            vol_in_answer = 0;
            touch(idx_to_id[pn]);
        }
        return 1;
    }
    // scheduleNTasksOnNodes (scheduler.go:844-924) over an ordered node list with one engine question per step: "does this
    // node still pass, and if so the next task goes there" = pe_fit (device filters + reservation), then VolumesFilter
    // here -- the pipeline's order, so `cnt` moves as the reference's failure counters do (cleared by every pass,
    // pipeline.go:64-66; `passed` says whether one happened).  Returns the tasks placed (a prefix of `tasks`), -1 on an
    // engine failure.
    int fill_stepwise(const std::vector<TaskP> &tasks, const std::vector<std::string> &cand, std::vector<uint32_t> &cnt, bool &passed,
                      std::map<std::string, Decision> &decisions) {
        const size_t k = tasks.size(), m = cand.size();
        if (!k || !m) return 0;
        std::vector<char> failed(m, 0);
        std::vector<uint8_t> ok; std::vector<uint32_t> fail;
        size_t it = 0, placed = 0;
        // (the first node is taken without a question: it passed when the list was made; pe_fit reserves there)
        if (!fit_many(tasks[0], {cand[0]}, ok, fail, false)) return -1;
        if (ok[0] != 1) { fatal = "placement engine: the best node of a group does not fit its first task"; return -1; }
        for (;;) {
            const std::string &nid = cand[it % m];
            assign_with_volumes(tasks[placed], nid, decisions);
            placed++;
            if (placed == k) return (int)placed;
            if (it + 1 < m) {                                                                         // first pass: level the nodes
                if (rank_key(nodeSet[cand[(it + 1) % m]], *tasks[0]).less_no_tie(rank_key(nodeSet[nid], *tasks[0]))) it++;
            } else it++;                                                                              // later passes: one task per node
            const size_t orig = it;
            bool found = false;
            while (!found) {
                const size_t i = it % m;
                if (!failed[i]) {
                    if (!fit_many(tasks[placed], {cand[i]}, ok, fail, false)) return -1;
                    if (ok[0] == 1 && volumes.filterCheck(*tasks[placed], *nodeSet[cand[i]].node)) {
                        std::fill(cnt.begin(), cnt.end(), 0u);
                        passed = true;
                        found = true;
                        break;
                    }
                    if (ok[0] == 1) { cnt[PE_F_VOLUMES]++; touch(cand[i]); }                        // (the reservation is taken back)
                    else for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] += fail[f];
                    failed[i] = 1;
                }
                it++;
                if (it - orig == m) break;                     // none of the nodes meets the constraints any more
            }
            if (!found) return (int)placed;
        }
    }
    bool scheduleVolumeGroupStepwise(std::vector<TaskP> &grp, const std::vector<std::string> &excluded, std::map<std::string, Decision> &decisions) {
        const TaskP t = grp[0];
        const size_t k = grp.size();
        auto give_back = [&](size_t from) { for (size_t i = from; i < grp.size(); i++) enqueue(grp[i]); layout_dirty = true; return false; };
        const std::set<std::string> allowed = vol_in;         // VolumesFilter's answer at tree-building time
        std::vector<uint32_t> cnt(PE_NUM_FILTERS, 0);
        // ---- the heap, in rank order
        std::vector<std::string> cand;
        const int pr = probe_best(t, {}, k, cand, &cnt);
        if (pr == 0) return give_back(0);
        if (pr == 2) { if (unsupported.empty()) unsupported = fatal; noSuitableNode(grp, "unsupported by the placement engine: " + fatal, decisions); fatal.clear(); return true; }
        if (cand.empty()) {                                   // nothing passed: every node is counted by its first failing filter
            if (!count_excluded(t, excluded, cnt)) return give_back(0);
            noSuitableNode(grp, explain(cnt.data()), decisions);
            return true;
        }
        const RankKey first_key = rank_key(nodeSet[cand[0]], *t);
        bool passed = false;                                  // has a Process() passed since the tree was built?
        const int res = fill_stepwise(grp, cand, cnt, passed, decisions);
        if (res < 0) { size_t done = 0; for (auto &x : grp) if (decisions.count(x->id)) done++; return give_back(done); }
        const size_t placed = (size_t)res;
        if (placed == k) return true;
        std::vector<TaskP> left(grp.begin() + (long)placed, grp.end());
        if (!passed) {
            // No re-check passed, so what the tree building left in the counters still stands (pipeline.go:55-68): the
            // failures after the last node that passed, among the nodes the tree building RAN the pipeline on -- every node
            // while the heap had room, afterwards only nodes that rank ahead of the heap's worst (nodeset.go:103-121).
            // Replay that walk (ascending node ID): pe_fit says which nodes pass the device filters -- every node but cand[0]
            // is as it was then (failed re-checks were taken back) -- `allowed` is VolumesFilter's answer of then, and the
            // rank keys are three integers of NodeInfo (cand[0]'s were saved before it took the task).
            std::vector<std::string> ids;
            for (auto &kv : nodeSet) if (kv.first != cand[0]) ids.push_back(kv.first);
            std::vector<uint8_t> ok; std::vector<uint32_t> fail;
            if (!fit_many(left[0], ids, ok, fail, true)) return give_back(placed);
            std::vector<uint32_t> tree(PE_NUM_FILTERS, 0);
            auto worse = [](const RankKey &a, const RankKey &b) { return a < b; };   // max-heap: the worst on top
            std::vector<RankKey> heap;
            size_t j = 0;
            for (auto &kv : nodeSet) {
                const bool is_first = kv.first == cand[0];
                const RankKey rk = is_first ? first_key : rank_key(kv.second, *t);
                const size_t q = is_first ? 0 : j++;
                if (heap.size() >= k && !(rk < heap.front())) continue;              // the pipeline was not run on it
                const bool passes = is_first || (ok[q] == 1 && allowed.count(kv.first));
                if (!passes) {
                    if (ok[q] == 1) tree[PE_F_VOLUMES]++;
                    else if (ok[q] == 0) for (int f = 0; f < PE_NUM_FILTERS; f++) tree[f] += fail[q * PE_NUM_FILTERS + f];
                    continue;
                }
                std::fill(tree.begin(), tree.end(), 0u);
                if (heap.size() >= k) { std::pop_heap(heap.begin(), heap.end(), worse); heap.pop_back(); }
                heap.push_back(rk); std::push_heap(heap.begin(), heap.end(), worse);
            }
            for (int f = 0; f < PE_NUM_FILTERS; f++) cnt[f] += tree[f];
        }
        noSuitableNode(left, explain(cnt.data()), decisions);
        return true;
    }
    bool scheduleVolumeGroup(std::vector<TaskP> &grp, std::map<std::string, Decision> &decisions) {
        const Task &t = *grp[0];
        auto refuse = [&](const std::string &why) {
            if (unsupported.empty()) unsupported = why;
            noSuitableNode(grp, "unsupported by the placement engine: " + why, decisions);
            return true;
        };
        const bool prefs = !preference_levels(t).empty();
        auto give_back = [&](const std::vector<TaskP> &ts) { for (auto &x : ts) enqueue(x); layout_dirty = true; };
        std::string mounts;
        for (auto &m : t.mounts) if (m.type == MountCluster) { mounts += m.source; mounts += m.read_only ? "\x01" : "\x02"; }
        if (!(vol_answer.valid && vol_answer.mounts == mounts && vol_answer.vol_epoch == volumes.epoch && vol_answer.node_epoch == node_epoch)) {
            vol_answer.allowed.clear(); vol_answer.excluded.clear();
            for (auto &kv : nodeSet) (volumes.filterCheck(t, *kv.second.node) ? vol_answer.allowed : vol_answer.excluded).push_back(kv.first);
            vol_answer.valid = true; vol_answer.mounts = mounts; vol_answer.vol_epoch = volumes.epoch; vol_answer.node_epoch = node_epoch;
            vol_answer.id++;
        }
        const std::vector<std::string> &allowed = vol_answer.allowed, &excluded = vol_answer.excluded;   // (stand until the next question)
        if (vol_in_answer != vol_answer.id || vol_col == PE_NONE) { mark_volume_nodes(allowed); vol_in_answer = vol_answer.id; }
        if (prefs) {
            // the tree's branches and task sums do not depend on the pipeline (nodeset.go:59-101 counts every node); each
            // leaf visit is the leaf's engine group cut down to the volume set as well when the answer of VolumesFilter
            // cannot move while this group is placed, and a walk of the leaf's heap step by step when it can
            leaf_also_in_volume_set = true;
            leaf_stepwise = grp.size() > 1 && !volumes.staticFor(t);
            const bool ok = schedulePreferenceGroup(grp, decisions);
            leaf_also_in_volume_set = leaf_stepwise = false;
            return ok;
        }
        if (grp.size() > 1 && !volumes.staticFor(t)) return scheduleVolumeGroupStepwise(grp, excluded, decisions);
        TickBuf b;
        fatal.clear();
        if (!encode_group(grp, b)) { std::string why = fatal; fatal.clear(); return refuse(why); }
        const TickBuf::Mark after_group = b.mark();
        auto restrict_to = [&](uint32_t gen) {
            b.rollback(after_group);
            pe_group &g = b.groups.back();
            if (g.con_cnt == 0) g.con_off = (uint32_t)b.cons.size();
            b.cons.resize(g.con_off + g.con_cnt);
            b.cons.push_back({vol_col, gen, 0});
            g.leaf_cnt = 1;
        };
        restrict_to(vol_gen);
        if (!flush_rows()) { give_back(grp); return false; }
        std::vector<uint32_t> out_node(grp.size(), PE_NONE), out_fail(PE_NUM_FILTERS, 0);
        pe_tick tk = b.view();
        if (!check(pe_schedule(eng, &tk, out_node.data(), out_fail.data()), "pe_schedule")) { give_back(grp); return false; }
        std::vector<TaskP> left;
        std::string first_placed;
        for (size_t i = 0; i < grp.size(); i++) {
            const TaskP &ti = grp[i];
            const uint32_t idx = out_node[i];
            if (idx == PE_NONE || idx >= idx_to_id.size()) { left.push_back(ti); continue; }
            if (first_placed.empty()) first_placed = idx_to_id[idx];
            assign_with_volumes(ti, idx_to_id[idx], decisions);
        }
        if (left.empty()) return true;
        std::vector<uint32_t> cnt(out_fail);
        // The reference ran its pipeline over EVERY node while it built the tree (scheduler.go:722-731), in this model's
        // canonical order (ascending node ID); a node outside the volume set fails there on its first failing filter -- a
        // device filter if one fails, VolumesFilter (the last of the pipeline) otherwise -- and is in no heap afterwards.
        // Process() clears the counters whenever a node passes (pipeline.go:55-68), so what Explain() reports is what failed
        // AFTER the last pass: with no task placed nothing ever passed and every excluded node counts; with one task placed
        // exactly one node passed the tree building (two would both be in the heap, and the second, untouched, would have
        // passed its re-check and taken the second task) and the excluded nodes after it count; with two or more placed a
        // re-check inside the fill loop passed (scheduler.go:912-920) and wiped everything from the tree building -- the
        // engine's counters, which cover the nodes of the set, are complete.  The group placed nothing on excluded nodes,
        // so asking about them now gives the answer of that moment.
        const size_t placed = grp.size() - left.size();
        std::vector<std::string> counted;
        if (placed == 0) counted = excluded;
        else if (placed == 1) for (auto &id : excluded) if (id > first_placed) counted.push_back(id);
        if (!counted.empty() && !count_excluded(left[0], counted, cnt)) { give_back(left); return false; }
        noSuitableNode(left, explain(cnt.data()), decisions);
        return true;
    }

    // ---- placement preferences: nodeSet.tree's branches (nodeset.go:59-101) + scheduleNTasksOnSubtree (scheduler.go:772-825).
    // The engine supplies the leaves with their task sums (pe_pref_leaves) and places n tasks on one leaf per call (a group
    // with leaf_cnt > 0: the leaf's k best feasible nodes and scheduleNTasksOnNodes over them); the walk between the
    // branches -- integer bookkeeping on a handful of branches -- is this host code, as in the reference.
    // Why one leaf visit = one engine group is exact: a visit assigns n' <= (tasks left) tasks and, before the fill
    // wraps, advances at most one node per task, so only the n' best live nodes of the leaf matter; the reference's leaf
    // heap holds the k best at tree-building time of which at least k - (tasks already placed there) >= n' are untouched and
    // still rank ahead of every node outside it, and feasibility only shrinks inside a group: the two candidate lists agree.
    struct PrefTree {
        int tasks = 0;
        std::map<std::string, std::unique_ptr<PrefTree>> next;   // canonical branch order: ascending label value, "" first
        std::vector<pe_constraint> leaf;                          // the (column == value) path that names a leaf
        std::vector<std::string> cand;                            // fillLeafStepwise: the leaf's heap (decision_tree.go), best first
        bool cand_built = false;
    };
    struct PrefWalk {
        std::vector<TaskP> pending;            // ascending task ID
        std::vector<uint32_t> last_fail;       // failure counters of the last visit that came back short
        bool have_fail = false, engine_failed = false;
        size_t group_size = 0;                 // maxAssignments of the tree building: the size of every leaf's heap
    };
    // A leaf visit of a group whose cluster volumes count their users: the leaf's heap is made once -- the k best nodes of
    // (leaf AND volume set) as the tree building saw them: nothing has touched this leaf's nodes since, other leaves hold
    // other nodes -- and every visit walks scheduleNTasksOnNodes over it with one engine question per step.  A visit after
    // the first finds the heap collapsed (decision_tree.go:30-45): every node is put to the pipeline again, the ones that
    // fail leave the heap for good, the rest are ordered by their rank of now.
    int fillLeafStepwise(int n, PrefTree &leaf, PrefWalk &w, std::map<std::string, Decision> &decisions) {
        const size_t m = std::min<size_t>((size_t)std::max(n, 0), w.pending.size());
        if (m == 0 || w.engine_failed) return 0;
        std::vector<TaskP> sub(w.pending.begin(), w.pending.begin() + (long)m);
        const TaskP t = sub[0];
        if (!leaf.cand_built) {
            const int pr = probe_best(t, leaf.leaf, w.group_size, leaf.cand, nullptr);
            if (pr == 0) { w.engine_failed = true; return 0; }
            if (pr == 2) { if (unsupported.empty()) unsupported = fatal; fatal.clear(); return 0; }
            leaf.cand_built = true;
        } else {
            std::vector<uint8_t> ok; std::vector<uint32_t> fail;
            if (!fit_many(t, leaf.cand, ok, fail, true)) { w.engine_failed = true; return 0; }
            std::vector<RankKey> kept;
            for (size_t i = 0; i < leaf.cand.size(); i++)
                if (ok[i] == 1 && volumes.filterCheck(*t, *nodeSet[leaf.cand[i]].node)) kept.push_back(rank_key(nodeSet[leaf.cand[i]], *t));
            std::sort(kept.begin(), kept.end());
            leaf.cand.clear();
            for (auto &rk : kept) leaf.cand.push_back(rk.id);
        }
        if (leaf.cand.empty()) return 0;
        std::vector<uint32_t> cnt(PE_NUM_FILTERS, 0);
        bool passed = false;
        const int res = fill_stepwise(sub, leaf.cand, cnt, passed, decisions);
        size_t placed = 0;
        if (res < 0) { w.engine_failed = true; for (auto &x : sub) if (decisions.count(x->id)) placed++; }
        else placed = (size_t)res;
        w.pending.erase(w.pending.begin(), w.pending.begin() + (long)placed);
        if (placed < m) { w.last_fail = cnt; w.have_fail = true; }
        return (int)placed;
    }
    int fillLeaf(int n, PrefTree &leaf, PrefWalk &w, std::map<std::string, Decision> &decisions) {
        if (leaf_stepwise) return fillLeafStepwise(n, leaf, w, decisions);
        const size_t m = std::min<size_t>((size_t)std::max(n, 0), w.pending.size());
        if (m == 0 || w.engine_failed) return 0;
        std::vector<TaskP> sub(w.pending.begin(), w.pending.begin() + (long)m);
        TickBuf b;
        fatal.clear();
        if (!encode_group(sub, b)) {                                   // unsupported reservation etc.: nothing is placed
            if (unsupported.empty()) unsupported = fatal;
            fatal.clear();
            return 0;
        }
        pe_group &g = b.groups.back();
        if (g.con_cnt == 0) g.con_off = (uint32_t)b.cons.size();       // the leaf terms follow the group's constraints
        b.cons.insert(b.cons.end(), leaf.leaf.begin(), leaf.leaf.end());
        g.leaf_cnt = (uint32_t)leaf.leaf.size();
        if (leaf_also_in_volume_set) { b.cons.push_back({vol_col, vol_gen, 0}); g.leaf_cnt++; }   // (scheduleVolumeGroup)
        if (!flush_rows()) { w.engine_failed = true; return 0; }
        std::vector<uint32_t> out_node(m, PE_NONE), out_fail(PE_NUM_FILTERS, 0);
        pe_tick tk = b.view();
        if (!check(pe_schedule(eng, &tk, out_node.data(), out_fail.data()), "pe_schedule")) { w.engine_failed = true; return 0; }
        std::vector<TaskP> left;
        apply_group(sub, out_node.data(), left, decisions);
        const size_t placed = m - left.size();
        // (scheduleNTasksOnNodes hands out the group's tasks in order: the ones it placed are a prefix)
        std::vector<TaskP> rest(left);
        rest.insert(rest.end(), w.pending.begin() + (long)m, w.pending.end());
        w.pending.swap(rest);
        if (placed < m) { w.last_fail = out_fail; w.have_fail = true; }
        return (int)placed;
    }
    int scheduleNTasksOnSubtree(int n, PrefTree &tree, PrefWalk &w, std::map<std::string, Decision> &decisions) {   // scheduler.go:772-825
        if (tree.next.empty()) return fillLeaf(n, tree, w, decisions);
        int tasksScheduled = 0, tasksInUsableBranches = tree.tasks;
        std::set<PrefTree *> noRoom;
        bool converging = true;
        while (tasksScheduled != n && noRoom.size() != tree.next.size() && converging) {
            const int usable = (int)(tree.next.size() - noRoom.size());
            const int desiredTasksPerBranch = (tasksInUsableBranches + n - tasksScheduled) / usable;
            int remainder = (tasksInUsableBranches + n - tasksScheduled) % usable;
            converging = false;
            for (auto &kv : tree.next) {
                PrefTree *subtree = kv.second.get();
                if (noRoom.count(subtree)) continue;
                const int subtreeTasks = subtree->tasks;
                if (subtreeTasks < desiredTasksPerBranch || (subtreeTasks == desiredTasksPerBranch && remainder > 0)) {
                    converging = true;
                    int tasksToAssign = desiredTasksPerBranch - subtreeTasks;
                    if (remainder > 0) tasksToAssign++;
                    const int res = scheduleNTasksOnSubtree(tasksToAssign, *subtree, w, decisions);
                    if (res < tasksToAssign) { noRoom.insert(subtree); tasksInUsableBranches -= subtreeTasks; }
                    else if (remainder > 0) remainder--;
                    tasksScheduled += res;
                }
            }
        }
        return tasksScheduled;
    }
    bool schedulePreferenceGroup(std::vector<TaskP> &grp, std::map<std::string, Decision> &decisions) {
        const Task &t = *grp[0];
        const std::vector<uint32_t> cols = preference_levels(t);
        auto give_back = [&](const std::vector<TaskP> &ts) { for (auto &x : ts) enqueue(x); layout_dirty = true; };
        if (!flush_rows()) { give_back(grp); return false; }            // (a new preference column is filled for every node)
        // ---- the leaves and their task sums, from the device mirror
        const uint32_t cap = (uint32_t)nodeSet.size() + 1u, L = (uint32_t)cols.size();
        std::vector<uint32_t> vals((size_t)cap * L), tasks(cap);
        uint32_t n_leaves = 0;
        if (!check(pe_pref_leaves(eng, svc_id(t.service), cols.data(), L, cap, vals.data(), tasks.data(), &n_leaves), "pe_pref_leaves")) { give_back(grp); return false; }
        PrefTree root;
        for (uint32_t i = 0; i < n_leaves; i++) {                        // nodeset.go:59-101, one leaf's worth of nodes at a time
            PrefTree *tr = &root;
            std::vector<pe_constraint> path;
            for (uint32_t l = 0; l < L; l++) {
                tr->tasks += (int)tasks[i];
                const uint32_t v = vals[(size_t)i * L + l];
                std::unique_ptr<PrefTree> &nx = tr->next[v < pref_strings.size() ? pref_strings[v] : std::string()];
                if (!nx) nx.reset(new PrefTree());
                tr = nx.get();
                path.push_back({cols[l], v, 0});
            }
            tr->tasks += (int)tasks[i];
            tr->leaf = path;
        }
        PrefWalk w;
        w.pending = grp;
        w.group_size = grp.size();
        scheduleNTasksOnSubtree((int)grp.size(), root, w, decisions);
        if (w.engine_failed) { give_back(w.pending); return false; }
        if (!w.pending.empty()) noSuitableNode(w.pending, w.have_fail ? explain(w.last_fail.data()) : std::string(), decisions);
        return true;
    }
    void noSuitableNode(const std::vector<TaskP> &left, const std::string &explanation, std::map<std::string, Decision> &decisions) {   // scheduler.go:928-971
        for (auto &t : left) {
            auto sv = services.find(t->service);
            if (sv == services.end()) continue;
            TaskP nt(new Task(*t));
            if (sv->second.first && t->has_spec_version && sv->second.second > t->spec_version) {
                if (t->state == TaskStatePending && t->desired >= TaskStateShutdown) { nt->state = TaskStateShutdown; nt->err.clear(); }
            } else {
                nt->err = explanation.empty() ? "no suitable node" : "no suitable node (" + explanation + ")";
                enqueue(nt);
            }
            allTasks[t->id] = nt;
            decisions[t->id] = {t, nt};
        }
    }
    void rollback(const Decision &d, bool requeue) {   // scheduler.go:472-487 / :416-424
        allTasks[d.old_->id] = d.old_;
        auto it = nodeSet.find(d.new_->node_id);
        if (it != nodeSet.end() && it->second.removeTask(*d.new_)) touch(d.new_->node_id);
        for (auto &va : d.new_->volumes) volumes.release(va.id, d.new_->id);        // release the volumes we tried to use (:422-424, :480-483)
        if (requeue) enqueue(d.old_);
    }
    // tick, scheduler.go:429-488
    bool tick(const std::set<std::string> &failCommit, std::map<std::string, Decision> &decisions) {
        host_encode_ms = host_engine_ms = host_apply_ms = 0;
        recycle_service_ids();
        std::map<std::pair<std::string, uint64_t>, std::vector<TaskP>> bySpec;
        std::vector<TaskP> oneOff;
        for (auto it = unassignedTasks.begin(); it != unassignedTasks.end(); it = unassignedTasks.erase(it)) {
            TaskP t = it->second;
            if (!t || !t->node_id.empty()) continue;
            if (t->has_spec_version) bySpec[{t->service, t->spec_version}].push_back(t);   // map order: ascending task ID inside a group
            else oneOff.push_back(t);
        }
        std::vector<std::vector<TaskP>> groups;
        for (auto &kv : bySpec) groups.push_back(kv.second);
        for (auto &t : oneOff) groups.push_back({t});
        if (!scheduleTaskGroups(groups, decisions)) return false;
        for (auto &kv : decisions) if (failCommit.count(kv.first) && kv.second.new_->state == TaskStateAssigned) rollback(kv.second, true);
        return true;
    }
    // processPreassignedTasks + taskFitNode, scheduler.go:398-426,646-690
    // one pe_fit call for a run of preassigned tasks.  vol_fail: the (single) task of the run has cluster mounts that this
    // node cannot satisfy -- VolumesFilter is the LAST filter of the pipeline, so a device filter that fails still names
    // the error; if none does, the engine's reservation is taken back (the row is uploaded again from NodeInfo).
    bool fit_run(const std::vector<TaskP> &pend, bool vol_fail, std::map<std::string, Decision> &decisions) {
        if (pend.empty()) return true;
        TickBuf b;
        for (auto &t : pend) { std::vector<TaskP> one{t}; if (!encode_group(one, b, true)) return false; }
        if (!flush_rows()) return false;
        std::vector<uint32_t> idx; for (auto &t : pend) idx.push_back(node_index(t->node_id));
        std::vector<uint8_t> ok(pend.size(), 0); std::vector<uint32_t> fail(pend.size() * PE_NUM_FILTERS, 0);
        pe_tick tk = b.view();
        if (!check(pe_fit(eng, &tk, idx.data(), ok.data(), fail.data()), "pe_fit")) return false;
        for (size_t i = 0; i < pend.size(); i++) {
            const TaskP &t = pend[i];
            if (ok[i] == 2) continue;
            TaskP nt(new Task(*t));
            if (ok[i] == 0) { nt->err = explain(&fail[i * PE_NUM_FILTERS]); allTasks[t->id] = nt; decisions[t->id] = {t, nt}; continue; }
            if (vol_fail) {
                touch(t->node_id);
                nt->err = "cannot fulfill requested CSI volume mounts on 1 node";                 // filter.go:442-447
                allTasks[t->id] = nt; decisions[t->id] = {t, nt};
                continue;
            }
            if (has_cluster_mounts(*t)) {      // scheduler.go:664-675: the attachments are chosen (and, on this path, not reserved)
                std::string verr;
                if (!volumes.choose(*t, *nodeSet[t->node_id].node, nt->volumes, verr)) { touch(t->node_id); nt->err = verr; allTasks[t->id] = nt; decisions[t->id] = {t, nt}; continue; }
            }
            nt->state = TaskStateAssigned; nt->err.clear(); nt->message = "scheduler confirmed task can run on preassigned node";
            allTasks[t->id] = nt;
            nodeSet[t->node_id].addTask(nt);
            decisions[t->id] = {t, nt};
        }
        return true;
    }
    bool processPreassignedTasks(const std::set<std::string> &failCommit, std::map<std::string, Decision> &decisions) {
        std::vector<TaskP> run;
        for (auto &kv : pendingPreassignedTasks) {
            const TaskP &t = kv.second;
            auto ns = nodeSet.find(t->node_id);
            if (ns == nodeSet.end()) continue;
            if (has_cluster_mounts(*t) && !volumes.filterCheck(*t, *ns->second.node)) {
                // (its own call, so that a reservation the engine makes for it is undone before the next task is checked)
                if (!fit_run(run, false, decisions)) return false;
                run.clear();
                if (!fit_run({t}, true, decisions)) return false;
                continue;
            }
            run.push_back(t);
        }
        if (!fit_run(run, false, decisions)) return false;
        for (auto &kv : decisions) {
            if (failCommit.count(kv.first)) { if (kv.second.new_->state == TaskStateAssigned) rollback(kv.second, false); }
            else if (kv.second.new_->state == TaskStateAssigned) pendingPreassignedTasks.erase(kv.first);
        }
        return true;
    }
};

// ------------------------------------------------------------------ JSON driver (same protocol as the oracle's)
static std::set<std::string> strset(const mj::Value &v) { std::set<std::string> s; for (auto &x : v.a) s.insert(x.as_str()); return s; }
static mj::Value decisions_json(const std::map<std::string, Decision> &ds) {
    mj::Value arr = mj::Value::array();
    for (auto &kv : ds) {
        mj::Value d = mj::Value::object();
        d.set("id", mj::Value::string(kv.first));
        d.set("node_id", mj::Value::string(kv.second.new_->node_id));
        d.set("state", mj::Value::string(task_state_name(kv.second.new_->state)));
        d.set("err", mj::Value::string(kv.second.new_->err));
        d.set("message", mj::Value::string(kv.second.new_->message));
        d.set("assigned_generic", generic_json(kv.second.new_->assigned));
        if (!kv.second.new_->volumes.empty()) {      // (only tasks with cluster mounts carry the key)
            mj::Value vols = mj::Value::array();
            for (auto &a : kv.second.new_->volumes) { mj::Value e = mj::Value::object(); e.set("id", mj::Value::string(a.id)); e.set("source", mj::Value::string(a.source)); e.set("target", mj::Value::string(a.target)); vols.push(e); }
            d.set("volumes", vols);
        }
        arr.push(d);
    }
    return arr;
}
static mj::Value apply(Scheduler &S, const mj::Value &ev) {
    mj::Value out = mj::Value::object();
    const std::string op = ev.at("op").as_str();
    if (ev.find("now_ns")) S.now = ev.at("now_ns").as_int();
    S.fatal.clear();
    if (op == "init") {
        std::vector<NodeP> nodes; std::vector<TaskP> tasks;
        for (auto &n : ev.at("nodes").a) nodes.push_back(parse_node(n));
        for (auto &t : ev.at("tasks").a) tasks.push_back(parse_task(t));
        for (auto &s : ev.at("services").a) S.services[s.at("id").as_str()] = {!s.at("spec_version").is_null(), (uint64_t)s.at("spec_version").as_int()};
        if (!ev.at("svc_recycle_at").is_null()) S.svc_recycle_min = S.svc_recycle_at = (size_t)ev.at("svc_recycle_at").as_int();
        std::vector<Volume> vols;
        if (!ev.at("volumes").is_null()) for (auto &v : ev.at("volumes").a) vols.push_back(parse_volume(v));
        S.setupTasksList(nodes, tasks, vols);
    } else if (op == "update_volume") S.updateVolume(parse_volume(ev.at("volume")));
    else if (op == "delete_volume") S.volumes.remove(ev.at("id").as_str());
    else if (op == "volume_usage") {
        mj::Value vols = mj::Value::object();
        for (auto &kv : S.volumes.vols) {
            mj::Value tasks = mj::Value::object();
            for (auto &t : kv.second.tasks) { mj::Value u = mj::Value::object(); u.set("node", mj::Value::string(t.second.node)); u.set("read_only", mj::Value::boolean(t.second.read_only)); tasks.set(t.first, u); }
            vols.set(kv.first, tasks);
        }
        out.set("volumes", vols);
    } else if (op == "set_service") S.services[ev.at("id").as_str()] = {!ev.at("spec_version").is_null(), (uint64_t)ev.at("spec_version").as_int()};
    else if (op == "delete_service") S.services.erase(ev.at("id").as_str());
    else if (op == "create_node" || op == "update_node") S.createOrUpdateNode(parse_node(ev.at("node")));
    else if (op == "delete_node") S.removeNode(ev.at("id").as_str());
    else if (op == "create_task") S.createTask(parse_task(ev.at("task")));
    else if (op == "update_task") S.updateTask(parse_task(ev.at("task")));
    else if (op == "delete_task") {
        Task t; t.id = ev.at("id").as_str();
        auto it = S.allTasks.find(t.id);
        if (ev.find("task")) t = *parse_task(ev.at("task")); else if (it != S.allTasks.end()) t = *it->second;
        S.deleteTask(t);
    } else if (op == "tick" || op == "preassigned") {
        std::map<std::string, Decision> ds;
        bool ok = op == "tick" ? S.tick(strset(ev.at("fail_commit")), ds) : S.processPreassignedTasks(strset(ev.at("fail_commit")), ds);
        if (!ok) { out.set("error", mj::Value::string(S.fatal)); return out; }
        out.set("decisions", decisions_json(ds));
        if (op == "tick") {      // where the host time of the tick went (encode groups / engine call / apply decisions)
            mj::Value hm = mj::Value::object();
            hm.set("encode", mj::Value::integer((int64_t)(S.host_encode_ms * 1000))); hm.set("engine", mj::Value::integer((int64_t)(S.host_engine_ms * 1000)));
            hm.set("apply", mj::Value::integer((int64_t)(S.host_apply_ms * 1000)));
            out.set("host_us", hm);
        }
        if (!S.unsupported.empty()) out.set("unsupported", mj::Value::string(S.unsupported));   // those tasks stayed pending
    } else if (op == "snapshot" || op == "device_check") {
        if (!S.flush_rows()) { out.set("error", mj::Value::string(S.fatal)); return out; }
        mj::Value arr = mj::Value::array(), bad = mj::Value::array();
        std::vector<pe_node_state> dev(S.nodeSet.size());
        if (op == "device_check" && !dev.empty() && !S.check(pe_snapshot(S.eng, 0, (uint32_t)dev.size(), dev.data()), "pe_snapshot")) { out.set("error", mj::Value::string(S.fatal)); return out; }
        uint32_t idx = 0;
        for (auto &kv : S.nodeSet) {
            const NodeInfo &ni = kv.second;
            if (op == "device_check") {
                // the device mirror must equal the host NodeInfo after every event / tick
                bool same = dev[idx].cpu_avail == ni.avail.cpu && dev[idx].mem_avail == ni.avail.mem && dev[idx].total_tasks == (uint32_t)ni.active;
                for (auto &s : ni.by_service) {
                    uint32_t v = 0;
                    auto known = S.svc_ids.find(s.first);        // (a service whose id was recycled counts nowhere)
                    if (known == S.svc_ids.end()) { same = same && s.second == 0; continue; }
                    if (!S.check(pe_snapshot_service(S.eng, known->second, idx, 1, &v), "pe_snapshot_service")) { out.set("error", mj::Value::string(S.fatal)); return out; }
                    same = same && v == (uint32_t)s.second;
                }
                // ... and no column counts anything for this node beyond what NodeInfo knows
                for (auto &kv : S.svc_ids) {
                    if (ni.by_service.count(kv.first)) continue;
                    uint32_t v = 0;
                    if (!S.check(pe_snapshot_service(S.eng, kv.second, idx, 1, &v), "pe_snapshot_service")) { out.set("error", mj::Value::string(S.fatal)); return out; }
                    same = same && v == 0u;
                }
                std::set<std::string> kinds; for (auto &g : ni.avail.generic) kinds.insert(g.kind);
                for (auto &kd : S.kind_ids) {
                    int64_t cell = 0, want = 0;
                    if (!S.check(pe_snapshot_generic(S.eng, kd.second, idx, 1, &cell), "pe_snapshot_generic")) { out.set("error", mj::Value::string(S.fatal)); return out; }
                    for (auto &g : ni.avail.generic) if (g.kind == kd.first) { int64_t c = 0; if (!g.named) c = g.amount; else for (auto &x : ni.avail.generic) if (x.kind == g.kind) c++; want = PE_GEN_ENCODE(c, g.named ? PE_GEN_NAMED : PE_GEN_DISCRETE); break; }
                    same = same && cell == want;
                }
                for (auto &ps : S.port_slots) {
                    uint8_t used = 0;
                    if (!S.check(pe_snapshot_ports(S.eng, ps.second, idx, 1, &used), "pe_snapshot_ports")) { out.set("error", mj::Value::string(S.fatal)); return out; }
                    same = same && (used != 0) == (ni.ports.count(ps.first) != 0);
                }
                if (!same) bad.push(mj::Value::string(kv.first));
            }
            mj::Value n = mj::Value::object();
            n.set("id", mj::Value::string(kv.first));
            n.set("active_tasks", mj::Value::integer(ni.active));
            mj::Value bs = mj::Value::object();
            for (auto &s : ni.by_service) bs.set(s.first, mj::Value::integer(s.second));
            n.set("by_service", bs);
            mj::Value av = mj::Value::object();
            av.set("nano_cpus", mj::Value::integer(ni.avail.cpu)); av.set("memory_bytes", mj::Value::integer(ni.avail.mem)); av.set("generic", generic_json(ni.avail.generic));
            n.set("available", av);
            mj::Value ports = mj::Value::array();
            for (auto &p : ni.ports) { mj::Value e = mj::Value::array(); e.push(mj::Value::integer(p.first)); e.push(mj::Value::integer(p.second)); ports.push(e); }
            n.set("ports", ports);
            mj::Value fl = mj::Value::object();
            for (auto &f : ni.failures) fl.set(f.first.svc + "@" + std::to_string(f.first.ver), mj::Value::integer((int64_t)f.second.size()));
            n.set("failures", fl);
            mj::Value tk = mj::Value::array();
            for (auto &t : ni.tasks) tk.push(mj::Value::string(t.first));
            n.set("tasks", tk);
            arr.push(n);
            idx++;
        }
        out.set("nodes", arr);
        if (op == "device_check") out.set("mismatch", bad);
        out.set("rows_uploaded", mj::Value::integer((int64_t)S.rows_uploaded));      // event ingestion: only rows
        out.set("full_uploads", mj::Value::integer((int64_t)S.full_uploads));        // store events touched cross the ABI
        out.set("service_ids", mj::Value::integer((int64_t)S.svc_ids.size()));       // ids in use / ever handed out (recycling)
        out.set("service_id_high_water", mj::Value::integer((int64_t)S.next_svc_id));
        mj::Value un = mj::Value::array();
        for (auto &kv : S.unassignedTasks) un.push(mj::Value::string(kv.first));
        out.set("unassigned", un);
        mj::Value pp = mj::Value::array();
        for (auto &kv : S.pendingPreassignedTasks) pp.push(mj::Value::string(kv.first));
        out.set("pending_preassigned", pp);
    } else if (op == "stats") {
        pe_stats st{};
        pe_get_stats(S.eng, &st);
        out.set("kernel_launches", mj::Value::integer((int64_t)st.kernel_launches));
        out.set("placements", mj::Value::integer((int64_t)st.placements));
    } else out.set("error", mj::Value::string("unknown op " + op));
    return out;
}

}  // namespace sk

// C entry points of libswarmsched.so.  ss_create fails (returns NULL) when the
// CUDA engine cannot be created: there is no CPU fallback.
extern "C" {
static std::string g_ss_err;
const char *ss_last_error() { return g_ss_err.c_str(); }
void *ss_create() {
    pe_config cfg{};
    cfg.abi_version = PE_ABI_VERSION; cfg.device = -1; cfg.node_capacity = 1024; cfg.world_size = 1;
    pe_engine *e = nullptr;
    int32_t rc = pe_create(&cfg, &e);
    if (rc != PE_OK) { g_ss_err = pe_last_error(nullptr); return nullptr; }
    sk::Scheduler *s = new sk::Scheduler();
    s->eng = e;
    return s;
}
void ss_destroy(void *h) {
    sk::Scheduler *s = reinterpret_cast<sk::Scheduler *>(h);
    if (!s) return;
    pe_destroy(s->eng);
    delete s;
}
char *ss_apply(void *h, const char *json) {
    std::string out;
    try { out = mj::dump(sk::apply(*reinterpret_cast<sk::Scheduler *>(h), mj::parse(json))); }
    catch (const std::exception &e) { mj::Value o = mj::Value::object(); o.set("error", mj::Value::string(e.what())); out = mj::dump(o); }
    char *r = (char *)malloc(out.size() + 1);
    memcpy(r, out.c_str(), out.size() + 1);
    return r;
}
void ss_free(char *p) { free(p); }
}
