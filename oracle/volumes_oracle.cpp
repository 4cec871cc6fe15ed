// volumes_oracle.cpp -- TEST INFRASTRUCTURE: CPU restatement of the reference's CSI cluster-volume
// bookkeeping and of the filter that reads it, so that a host-side VolumesFilter -- the
// product refuses such groups today, DESIGN.md section 8 -- has its checker ready and pinned:
//
//   volumeSet.addOrUpdateVolume / removeVolume        manager/scheduler/volumes.go:61-96
//   volumeSet.chooseTaskVolumes                       manager/scheduler/volumes.go:98-136
//   volumeSet.reserveTaskVolumes / reserveVolume      manager/scheduler/volumes.go:138-160
//   volumeSet.releaseVolume                           manager/scheduler/volumes.go:162-184
//   volumeSet.isVolumeAvailableOnNode                 manager/scheduler/volumes.go:223-255
//   volumeSet.checkVolume / hasWriter                 manager/scheduler/volumes.go:257-327
//   IsInTopology                                      manager/scheduler/topology.go:22-47
//   VolumesFilter.SetTask / Check                     manager/scheduler/filter.go:399-440
//
// Pinned by the reference's own tests, ported in tests/test_oracle_volumes.py: topology_test.go:9-177
// (TestIsInTopology, 7 cases), volumes_test.go:47-160 (add / remove / track / reserve), :164-343 (the
// checkVolume table, 9 cases), :344-471 (volume or group availability), :473-528 (chooseTaskVolumes).
// Canonicalisation: the reference ranges over Go maps (the volumes of a group, volumes.go:233); here
// a group's volumes are visited in ascending volume ID.
//
// Nothing under swarmkit_b200/ or bench.py's timed region links, loads or calls this file.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "../swarmkit_b200/csrc/minijson.h"   // (a JSON reader for the test-driver protocol: utility, not algorithm)

namespace {

enum Scope { ScopeSingleNode = 0, ScopeMultiNode = 1 };                         // api.VolumeAccessMode_Scope
enum Sharing { SharingNone = 0, SharingReadOnly = 1, SharingOneWriter = 2, SharingAll = 3 };   // api.VolumeAccessMode_SharingMode
enum { AvailabilityActive = 0, AvailabilityPause = 1, AvailabilityDrain = 2 };  // api.VolumeSpec_VolumeAvailability

using Segments = std::map<std::string, std::string>;
struct Topology { bool present = false; Segments segments; };   // *api.Topology (nil = not present)

struct Volume {
    std::string id, name, group, driver;
    int availability = AvailabilityActive, scope = ScopeSingleNode, sharing = SharingNone;
    bool has_info = false;                       // VolumeInfo != nil
    std::vector<Segments> accessible;            // VolumeInfo.AccessibleTopology
};
struct Usage { std::string node; bool read_only = false; };
struct VolumeInfo {
    Volume volume;
    std::map<std::string, Usage> tasks;          // task ID -> usage
    std::map<std::string, int> nodes;            // node ID -> active tasks using the volume there
};
struct CSIInfo { std::string plugin; Topology top; };
struct Node { std::string id; bool has_desc = false; std::vector<CSIInfo> csi; };
struct Mount { std::string type, source, target; bool read_only = false; };
struct Attachment { std::string id, source, target; };

// IsInTopology, topology.go:22-47
bool is_in_topology(const Topology &top, const std::vector<Segments> &accessible) {
    if (!top.present || accessible.empty()) return true;          // any part of the equation missing: it does fit
    for (const Segments &topology : accessible) {
        bool all = true;
        for (auto &kv : topology) {
            auto f = top.segments.find(kv.first);
            if ((f == top.segments.end() ? std::string() : f->second) != kv.second) { all = false; break; }
        }
        if (all) return true;
    }
    return false;
}

struct VolumeSet {
    std::map<std::string, VolumeInfo> volumes;
    std::map<std::string, std::set<std::string>> by_group;       // canonical: ascending volume ID inside a group
    std::map<std::string, std::string> by_name;

    // addOrUpdateVolume, volumes.go:61-81
    void add_or_update(const Volume &v) {
        auto it = volumes.find(v.id);
        if (it == volumes.end()) { VolumeInfo vi; vi.volume = v; volumes[v.id] = vi; }
        // (an update assigns to a COPY of the map's value in the reference, volumes.go:69-70: the stored volume keeps
        // its old spec.  Restated as is.)
        by_group[v.group].insert(v.id);
        by_name[v.name] = v.id;
    }
    // removeVolume, volumes.go:83-96
    void remove(const std::string &id) {
        auto it = volumes.find(id);
        if (it == volumes.end()) return;
        by_group[it->second.volume.group].erase(id);
        by_name.erase(it->second.volume.name);
        volumes.erase(it);
    }
    // reserveVolume, volumes.go:150-160
    void reserve(const std::string &vid, const std::string &task, const std::string &node, bool read_only) {
        auto it = volumes.find(vid);
        if (it == volumes.end()) return;
        it->second.tasks[task] = Usage{node, read_only};
        it->second.nodes[node] += 1;
    }
    // releaseVolume, volumes.go:162-184
    void release(const std::string &vid, const std::string &task) {
        auto it = volumes.find(vid);
        if (it == volumes.end()) return;
        auto u = it->second.tasks.find(task);
        if (u == it->second.tasks.end()) return;
        int &c = it->second.nodes[u->second.node];
        if (c > 0) c -= 1;
        it->second.tasks.erase(u);
    }
    static bool has_writer(const VolumeInfo &vi) {               // volumes.go:320-327
        for (auto &kv : vi.tasks) if (!kv.second.read_only) return true;
        return false;
    }
    // checkVolume, volumes.go:257-318
    bool check(const std::string &id, const Node &node, bool read_only) const {
        auto it = volumes.find(id);
        if (it == volumes.end()) return false;
        const VolumeInfo &vi = it->second;
        if (vi.volume.availability != AvailabilityActive) return false;
        Topology top;                                             // the node's topology for the volume's plugin
        if (node.has_desc)
            for (auto &info : node.csi) if (info.plugin == vi.volume.driver) { top = info.top; break; }
        if (vi.volume.scope == ScopeSingleNode)
            for (auto &kv : vi.tasks) if (kv.second.node != node.id) return false;
        switch (vi.volume.sharing) {
            case SharingNone: if (!vi.tasks.empty()) return false; break;
            case SharingOneWriter: if (!read_only && has_writer(vi)) return false; break;
            case SharingReadOnly: if (!read_only) return false; break;
            default: break;
        }
        return is_in_topology(top, vi.volume.has_info ? vi.volume.accessible : std::vector<Segments>());
    }
    // isVolumeAvailableOnNode, volumes.go:223-255: the volume ID that satisfies the mount on this node, or ""
    std::string available_on_node(const Mount &m, const Node &node) const {
        const std::string &source = m.source;
        if (source.compare(0, 6, "group:") == 0) {
            auto g = by_group.find(source.substr(6));
            if (g == by_group.end()) return "";
            for (auto &id : g->second) if (check(id, node, m.read_only)) return id;
            return "";
        }
        auto n = by_name.find(source);
        if (n == by_name.end() || !check(n->second, node, m.read_only)) return "";
        return n->second;
    }
    // chooseTaskVolumes, volumes.go:98-136: reservations made while choosing are released again before returning
    bool choose(const std::string &task_id, const std::vector<Mount> &mounts, const Node &node, std::vector<Attachment> &out, std::string &err) {
        std::vector<Attachment> chosen;
        bool ok = true;
        for (auto &m : mounts) {
            if (m.type != "CLUSTER") continue;
            std::string cand = available_on_node(m, node);
            if (cand.empty()) { err = "cannot find volume to satisfy mount with source " + m.source; ok = false; break; }
            reserve(cand, task_id, node.id, m.read_only);
            chosen.push_back({cand, m.source, m.target});
        }
        for (auto &a : chosen) release(a.id, task_id);
        if (ok) out = chosen;
        return ok;
    }
    // reserveTaskVolumes, volumes.go:138-148
    void reserve_task(const std::string &task_id, const std::string &node_id, const std::vector<Attachment> &atts, const std::vector<Mount> &mounts) {
        for (auto &va : atts)
            for (auto &m : mounts)
                if (m.source == va.source && m.target == va.target) reserve(va.id, task_id, node_id, m.read_only);
    }
};

// VolumesFilter, filter.go:388-447
struct VolumesFilter {
    std::vector<Mount> requested;
    bool set_task(bool has_container, const std::vector<Mount> &mounts) {      // :399-430
        requested.clear();
        if (!has_container) return false;
        bool has_csi = false;
        for (auto &m : mounts) if (m.type == "CLUSTER") { has_csi = true; requested.push_back(m); }
        return has_csi;
    }
    bool check(const VolumeSet &vs, const Node &n) const {                     // :432-440 (true if ANY requested mount can be met)
        for (auto &m : requested) if (!vs.available_on_node(m, n).empty()) return true;
        return false;
    }
};

// ---- JSON glue (test-driver protocol) ----------------------------------------------------------------------------
Segments parse_segments(const mj::Value &v) { Segments s; for (auto &kv : v.o) s[kv.first] = kv.second.as_str(); return s; }
int scope_of(const std::string &s) { return s == "MULTI_NODE" ? ScopeMultiNode : ScopeSingleNode; }
int sharing_of(const std::string &s) { return s == "ALL" ? SharingAll : s == "ONE_WRITER" ? SharingOneWriter : s == "READ_ONLY" ? SharingReadOnly : SharingNone; }
Volume parse_volume(const mj::Value &v) {
    Volume x;
    x.id = v.at("id").as_str(); x.name = v.at("name").as_str(); x.group = v.at("group").as_str(); x.driver = v.at("driver").as_str();
    const std::string av = v.at("availability").as_str();
    x.availability = av == "PAUSE" ? AvailabilityPause : av == "DRAIN" ? AvailabilityDrain : AvailabilityActive;
    x.scope = scope_of(v.at("scope").as_str()); x.sharing = sharing_of(v.at("sharing").as_str());
    if (!v.at("accessible_topology").is_null()) { x.has_info = true; for (auto &t : v.at("accessible_topology").a) x.accessible.push_back(parse_segments(t)); }
    else x.has_info = !v.at("volume_info").is_null();
    return x;
}
Node parse_node(const mj::Value &v) {
    Node n; n.id = v.at("id").as_str();
    if (!v.at("csi").is_null()) {
        n.has_desc = true;
        for (auto &c : v.at("csi").a) {
            CSIInfo i; i.plugin = c.at("plugin").as_str();
            if (!c.at("topology").is_null()) { i.top.present = true; i.top.segments = parse_segments(c.at("topology")); }
            n.csi.push_back(i);
        }
    }
    return n;
}
std::vector<Mount> parse_mounts(const mj::Value &v) {
    std::vector<Mount> out;
    for (auto &m : v.a) { Mount x; x.type = m.at("type").as_str(); x.source = m.at("source").as_str(); x.target = m.at("target").as_str(); x.read_only = !m.at("read_only").is_null() && m.at("read_only").as_bool(); out.push_back(x); }
    return out;
}
std::vector<Attachment> parse_attachments(const mj::Value &v) {
    std::vector<Attachment> out;
    for (auto &a : v.a) out.push_back({a.at("id").as_str(), a.at("source").as_str(), a.at("target").as_str()});
    return out;
}

mj::Value apply(VolumeSet &vs, const mj::Value &ev) {
    mj::Value out = mj::Value::object();
    const std::string op = ev.at("op").as_str();
    if (op == "is_in_topology") {
        Topology top;
        if (!ev.at("top").is_null()) { top.present = true; top.segments = parse_segments(ev.at("top")); }
        std::vector<Segments> acc;
        if (!ev.at("accessible").is_null()) for (auto &t : ev.at("accessible").a) acc.push_back(parse_segments(t));
        out.set("result", mj::Value::boolean(is_in_topology(top, acc)));
    } else if (op == "add_volume") vs.add_or_update(parse_volume(ev.at("volume")));
    else if (op == "remove_volume") vs.remove(ev.at("id").as_str());
    else if (op == "reserve") vs.reserve(ev.at("volume").as_str(), ev.at("task").as_str(), ev.at("node").as_str(), ev.at("read_only").as_bool());
    else if (op == "release") vs.release(ev.at("volume").as_str(), ev.at("task").as_str());
    else if (op == "check") out.set("result", mj::Value::boolean(vs.check(ev.at("volume").as_str(), parse_node(ev.at("node")), ev.at("read_only").as_bool())));
    else if (op == "available") { auto ms = parse_mounts(ev.at("mounts")); out.set("result", mj::Value::string(ms.empty() ? "" : vs.available_on_node(ms[0], parse_node(ev.at("node"))))); }
    else if (op == "choose") {
        std::vector<Attachment> atts; std::string err;
        const bool ok = vs.choose(ev.at("task").as_str(), parse_mounts(ev.at("mounts")), parse_node(ev.at("node")), atts, err);
        mj::Value arr = mj::Value::array();
        for (auto &a : atts) { mj::Value x = mj::Value::object(); x.set("id", mj::Value::string(a.id)); x.set("source", mj::Value::string(a.source)); x.set("target", mj::Value::string(a.target)); arr.push(x); }
        out.set("ok", mj::Value::boolean(ok)); out.set("attachments", arr); out.set("error", mj::Value::string(err));
    } else if (op == "reserve_task") vs.reserve_task(ev.at("task").as_str(), ev.at("node").as_str(), parse_attachments(ev.at("attachments")), parse_mounts(ev.at("mounts")));
    else if (op == "filter") {
        VolumesFilter f;
        const bool enabled = f.set_task(ev.at("has_container").is_null() || ev.at("has_container").as_bool(), parse_mounts(ev.at("mounts")));
        out.set("enabled", mj::Value::boolean(enabled));
        mj::Value arr = mj::Value::array();
        if (!ev.at("nodes").is_null()) for (auto &n : ev.at("nodes").a) arr.push(mj::Value::boolean(enabled ? f.check(vs, parse_node(n)) : true));
        out.set("pass", arr);
    } else if (op == "dump") {
        mj::Value vols = mj::Value::object();
        for (auto &kv : vs.volumes) {
            mj::Value v = mj::Value::object(), tasks = mj::Value::object(), nodes = mj::Value::object();
            for (auto &t : kv.second.tasks) { mj::Value u = mj::Value::object(); u.set("node", mj::Value::string(t.second.node)); u.set("read_only", mj::Value::boolean(t.second.read_only)); tasks.set(t.first, u); }
            for (auto &n : kv.second.nodes) nodes.set(n.first, mj::Value::integer(n.second));
            v.set("tasks", tasks); v.set("nodes", nodes);
            vols.set(kv.first, v);
        }
        mj::Value groups = mj::Value::object();
        for (auto &kv : vs.by_group) { if (kv.second.empty()) continue; mj::Value a = mj::Value::array(); for (auto &id : kv.second) a.push(mj::Value::string(id)); groups.set(kv.first, a); }
        mj::Value names = mj::Value::object();
        for (auto &kv : vs.by_name) names.set(kv.first, mj::Value::string(kv.second));
        out.set("volumes", vols); out.set("by_group", groups); out.set("by_name", names);
    } else out.set("error", mj::Value::string("unknown op " + op));
    return out;
}

}  // namespace

extern "C" {
void *vo_create() { return new VolumeSet(); }
void vo_destroy(void *h) { delete static_cast<VolumeSet *>(h); }
char *vo_apply(void *h, const char *json) {
    std::string outs;
    try {
        mj::Value ev = mj::parse(json);
        outs = mj::dump(apply(*static_cast<VolumeSet *>(h), ev));
    } catch (const std::exception &e) {
        mj::Value o = mj::Value::object(); o.set("error", mj::Value::string(e.what())); outs = mj::dump(o);
    }
    char *p = static_cast<char *>(malloc(outs.size() + 1));
    memcpy(p, outs.c_str(), outs.size() + 1);
    return p;
}
void vo_free(char *p) { free(p); }
}
