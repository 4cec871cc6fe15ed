// sched_oracle.cpp -- object-level CPU oracle of the SwarmKit scheduler hot path.
//
// TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/ and __graft_entry__.smoke()
// may load this library; nothing under swarmkit_b200/ links, loads or calls it.
//
// A literal C++ restatement of the reference (moby/swarmkit @ 12ce3490) on
// string-keyed objects, so that the reference's own known-answer tests can be
// replayed against it (tests/test_oracle_known_answers.py).  The only liberty
// taken is SURVEY.md 8(c)'s canonicalisation of Go map iteration order:
//   nodes by ascending node ID (which is also the final tie-break of nodeLess),
//   task groups by (ServiceID, SpecVersion.Index), one-offs and the tasks of a
//   group by ascending task ID, preference branches by ascending label value,
//   `now` supplied by the caller once per tick.
//
// Reference map (file:line under /root/reference):
//   constraint::parse / match / node_matches  manager/constraint/constraint.go:40-207
//   genericresource::*                        api/genericresource/{validate.go:24-51,
//                                             helpers.go:43-111, resource_management.go:11-205}
//   NodeInfo::{add,remove}Task, taskFailed,
//     countRecentFailures                     manager/scheduler/nodeinfo.go:46-221
//   filters                                   manager/scheduler/filter.go:30-386
//   Pipeline                                  manager/scheduler/pipeline.go:38-103
//   nodeSet.tree, decisionTree, nodeMaxHeap   manager/scheduler/{nodeset.go:50-124,
//                                             decision_tree.go:24-52, nodeheap.go}
//   Scheduler::{setupTasksList, createTask, updateTask, deleteTask,
//     createOrUpdateNode, processPreassignedTasks, tick, taskFitNode,
//     scheduleTaskGroup, scheduleNTasksOnSubtree, scheduleNTasksOnNodes,
//     noSuitableNode, buildNodeSet}           manager/scheduler/scheduler.go:68-990
//   volumeSet, IsInTopology, VolumesFilter      manager/scheduler/{volumes.go:45-327,
//                                             topology.go:22-47, filter.go:388-447}
//     (unit-level twin with the reference's tables: oracle/volumes_oracle.cpp; here
//     they are wired into the scheduler as scheduler.go:68-125,205-217,350-366,
//     398-487,646-690,844-924 wire them)
//
// Driver protocol: so_create() / so_apply(handle, json) -> json / so_free().
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../swarmkit_b200/csrc/minijson.h"   // (a JSON reader for the test-driver protocol: utility, not algorithm)

namespace api {
enum TaskState { TaskStateNew = 0, TaskStatePending = 64, TaskStateAssigned = 192, TaskStateAccepted = 256, TaskStatePreparing = 320,
                 TaskStateReady = 384, TaskStateStarting = 448, TaskStateRunning = 512, TaskStateCompleted = 576, TaskStateShutdown = 640,
                 TaskStateFailed = 704, TaskStateRejected = 768, TaskStateRemove = 800, TaskStateOrphaned = 832 };
enum NodeState { NodeStatus_UNKNOWN = 0, NodeStatus_DOWN = 1, NodeStatus_READY = 2, NodeStatus_DISCONNECTED = 3 };
enum Availability { NodeAvailabilityActive = 0, NodeAvailabilityPause = 1, NodeAvailabilityDrain = 2 };
enum Role { NodeRoleWorker = 0, NodeRoleManager = 1 };
enum Protocol { ProtocolTCP = 0, ProtocolUDP = 1, ProtocolSCTP = 2 };
enum PublishMode { PublishModeIngress = 0, PublishModeHost = 1 };
enum MountType { MountTypeBind = 0, MountTypeVolume = 1, MountTypeTmpfs = 2, MountTypeNamedPipe = 3, MountTypeCluster = 4 };

struct GenericResource {
    bool named = false;
    std::string kind;
    std::string value;  // named
    int64_t amount = 0; // discrete
};
struct Resources {
    int64_t NanoCPUs = 0, MemoryBytes = 0;
    std::vector<GenericResource> Generic;
};
struct Platform { std::string Architecture, OS; };
struct PluginDescription { std::string Type, Name; };
struct EngineDescription {
    bool has_labels = false;
    std::map<std::string, std::string> Labels;
    std::vector<PluginDescription> Plugins;
};
struct Topology { bool present = false; std::map<std::string, std::string> Segments; };   // *api.Topology
struct NodeCSIInfo { std::string PluginName, NodeID; Topology AccessibleTopology; };
struct NodeDescription {
    std::string Hostname;
    bool has_platform = false; Platform platform;
    bool has_resources = false; Resources resources;
    bool has_engine = false; EngineDescription engine;
    std::vector<NodeCSIInfo> CSIInfo;
};
enum { VolumeScopeSingleNode = 0, VolumeScopeMultiNode = 1 };
enum { VolumeSharingNone = 0, VolumeSharingReadOnly = 1, VolumeSharingOneWriter = 2, VolumeSharingAll = 3 };
enum { VolumeAvailabilityActive = 0, VolumeAvailabilityPause = 1, VolumeAvailabilityDrain = 2 };
struct Volume {
    std::string ID, Name, Group, Driver;          // Spec.Annotations.Name, Spec.Group, Spec.Driver.Name
    int Availability = VolumeAvailabilityActive, Scope = VolumeScopeSingleNode, Sharing = VolumeSharingNone;
    bool has_info = false; std::string VolumeID;  // VolumeInfo, VolumeInfo.VolumeID
    std::vector<std::map<std::string, std::string>> AccessibleTopology;
};
struct VolumeAttachment { std::string ID, Source, Target; };
struct Node {
    std::string ID;
    bool has_description = false; NodeDescription Description;
    int state = NodeStatus_UNKNOWN; std::string Addr;
    int availability = NodeAvailabilityActive;
    bool has_labels = false; std::map<std::string, std::string> Labels;  // Spec.Annotations.Labels
    int role = NodeRoleWorker;
    uint64_t version = 0;  // Meta.Version.Index
};
struct PortConfig { int protocol = ProtocolTCP; uint32_t PublishedPort = 0; int publish_mode = PublishModeIngress; };
struct Mount { int type = MountTypeBind; bool has_driver = false; std::string driver_name; std::string Source, Target; bool ReadOnly = false; };
struct Placement {
    std::vector<std::string> Constraints;
    std::vector<std::string> Preferences;  // spread descriptors
    std::vector<Platform> Platforms;
    uint64_t MaxReplicas = 0;
};
struct Task {
    std::string ID, ServiceID, NodeID;
    uint64_t Slot = 0;
    int DesiredState = TaskStateRunning;
    int state = TaskStateNew; std::string Err, Message;
    bool has_spec_version = false; uint64_t SpecVersion = 0;
    bool has_reservations = false; Resources Reservations;
    bool has_placement = false; Placement placement;
    bool has_container = false; std::vector<Mount> Mounts;
    bool has_log_driver = false; std::string LogDriver;
    struct Net { bool has_driver = false; std::string driver; };
    std::vector<Net> Networks;
    bool has_endpoint = false; std::vector<PortConfig> Ports;
    std::vector<GenericResource> AssignedGenericResources;
    std::vector<VolumeAttachment> Volumes;
};
}  // namespace api

using TaskP = std::shared_ptr<api::Task>;
using NodeP = std::shared_ptr<api::Node>;

// ============================================================== strings / net
namespace strs {
// Canonical form under strings.EqualFold for the runes a constraint can hold
// (constraint.go:22-26): ASCII case + the two non-ASCII members of the k/s
// simple-fold orbits.  Other runes only equal themselves.
static std::string fold(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size();) {
        unsigned char c = (unsigned char)s[i];
        if (c >= 'A' && c <= 'Z') { o += (char)(c + 32); i++; }
        else if (c == 0xE2 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x84 && (unsigned char)s[i + 2] == 0xAA) { o += 'k'; i += 3; }
        else if (c == 0xC5 && i + 1 < s.size() && (unsigned char)s[i + 1] == 0xBF) { o += 's'; i += 2; }
        else { o += (char)c; i++; }
    }
    return o;
}
static bool equal_fold(const std::string &a, const std::string &b) { return fold(a) == fold(b); }
static bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
static std::string trim_space(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && is_space((unsigned char)s[a])) a++;
    while (b > a && is_space((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
static bool has_prefix(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
}  // namespace strs

namespace netx {
struct IP { bool ok = false; uint8_t b[16] = {0}; bool is4() const { static const uint8_t pre[12] = {0,0,0,0,0,0,0,0,0,0,0xff,0xff}; return ok && std::memcmp(b, pre, 12) == 0; } };

static bool parse_v4(const std::string &s, uint8_t out[4]) {
    size_t p = 0;
    for (int f = 0; f < 4; f++) {
        if (p >= s.size()) return false;
        if (f > 0) { if (s[p] != '.') return false; p++; }
        size_t st = p; int v = 0;
        while (p < s.size() && s[p] >= '0' && s[p] <= '9') { v = v * 10 + (s[p] - '0'); if (v > 255) return false; p++; }
        if (p == st) return false;
        if (p - st > 1 && s[st] == '0') return false;  // leading zeros rejected (Go >= 1.17)
        out[f] = (uint8_t)v;
    }
    return p == s.size();
}
// net.ParseIP
static IP parse_ip(const std::string &s) {
    IP ip;
    uint8_t v4[4];
    if (s.find(':') == std::string::npos) {
        if (!parse_v4(s, v4)) return ip;
        ip.ok = true; ip.b[10] = ip.b[11] = 0xff; std::memcpy(ip.b + 12, v4, 4);
        return ip;
    }
    // IPv6 (no zone)
    std::vector<uint16_t> head, tail; bool ellipsis = false;
    size_t p = 0;
    if (s.size() >= 2 && s[0] == ':' && s[1] == ':') { ellipsis = true; p = 2; }
    std::vector<uint16_t> *cur = ellipsis ? &tail : &head;
    while (p < s.size()) {
        size_t st = p; uint32_t v = 0; int nd = 0;
        while (p < s.size() && isxdigit((unsigned char)s[p]) && nd < 5) { v = v * 16 + (uint32_t)(isdigit((unsigned char)s[p]) ? s[p] - '0' : (tolower(s[p]) - 'a' + 10)); p++; nd++; }
        if (nd == 0 || nd > 4) return ip;
        if (p < s.size() && s[p] == '.') {  // embedded IPv4 tail
            if (!parse_v4(s.substr(st), v4)) return ip;
            cur->push_back((uint16_t)((v4[0] << 8) | v4[1])); cur->push_back((uint16_t)((v4[2] << 8) | v4[3]));
            p = s.size();
            break;
        }
        cur->push_back((uint16_t)v);
        if (p == s.size()) break;
        if (s[p] != ':') return ip;
        p++;
        if (p < s.size() && s[p] == ':') { if (ellipsis) return ip; ellipsis = true; cur = &tail; p++; if (p == s.size()) break; }
        else if (p == s.size()) return ip;  // trailing single colon
    }
    size_t n = head.size() + tail.size();
    if ((!ellipsis && n != 8) || (ellipsis && n > 7)) return ip;
    std::vector<uint16_t> all(head);
    all.resize(8 - tail.size(), 0);
    all.insert(all.end(), tail.begin(), tail.end());
    for (int i = 0; i < 8; i++) { ip.b[2 * i] = (uint8_t)(all[i] >> 8); ip.b[2 * i + 1] = (uint8_t)all[i]; }
    ip.ok = true;
    return ip;
}
struct IPNet { bool ok = false; bool v4 = false; uint8_t ip[16] = {0}, mask[16] = {0}; };
// net.ParseCIDR
static IPNet parse_cidr(const std::string &s) {
    IPNet n;
    size_t sl = s.find('/');
    if (sl == std::string::npos) return n;
    std::string addr = s.substr(0, sl), bits = s.substr(sl + 1);
    bool v4 = addr.find(':') == std::string::npos;
    IP ip = parse_ip(addr);
    if (!ip.ok) return n;
    if (bits.empty() || bits.size() > 3) return n;
    int nb = 0;
    for (char c : bits) { if (c < '0' || c > '9') return n; nb = nb * 10 + (c - '0'); }
    if (bits.size() > 1 && bits[0] == '0') return n;
    int len = v4 ? 32 : 128;
    if (nb > len) return n;
    n.ok = true; n.v4 = v4;
    int off = v4 ? 12 : 0;
    for (int i = 0; i < off; i++) n.mask[i] = 0xff;
    for (int i = 0; i < len / 8; i++) { int r = nb - 8 * i; n.mask[off + i] = r >= 8 ? 0xff : r <= 0 ? 0 : (uint8_t)(0xff << (8 - r)); }
    for (int i = 0; i < 16; i++) n.ip[i] = ip.b[i] & n.mask[i];
    return n;
}
// (*IPNet).Contains
static bool contains(const IPNet &n, const IP &ip) {
    if (!ip.ok) return false;
    if (n.v4 != ip.is4()) return false;
    for (int i = 0; i < 16; i++) if ((ip.b[i] & n.mask[i]) != n.ip[i]) return false;
    return true;
}
static bool equal(const IP &a, const IP &b) { return a.ok && b.ok && std::memcmp(a.b, b.b, 16) == 0; }
}  // namespace netx

// ================================================================ constraint
namespace constraint {
enum { eq = 0, noteq = 1 };
static const std::string NodeLabelPrefix = "node.labels.", EngineLabelPrefix = "engine.labels.";
struct Constraint { std::string key; int op; std::string exp; };

static bool fold_alpha(unsigned char c) { return (c >= 'a' && c <= 'z'); }
// alphaNumeric = ^(?i)[a-z_][a-z0-9\-_.]+$   (constraint.go:22), on the folded string
static bool key_valid(const std::string &k0) {
    std::string k = strs::fold(k0);
    if (k.size() < 2) return false;
    if (!(fold_alpha((unsigned char)k[0]) || k[0] == '_')) return false;
    for (size_t i = 1; i < k.size(); i++) {
        unsigned char c = (unsigned char)k[i];
        if (!(fold_alpha(c) || (c >= '0' && c <= '9') || c == '-' || c == '_' || c == '.')) return false;
    }
    return true;
}
// valuePattern = ^(?i)[a-z0-9:\-_\s\.\*\(\)\?\+\[\]\\\^\$\|\/]+$   (constraint.go:26)
static bool value_valid(const std::string &v0) {
    std::string v = strs::fold(v0);
    if (v.empty()) return false;
    for (unsigned char c : v) {
        if (fold_alpha(c) || (c >= '0' && c <= '9')) continue;
        if (std::strchr(":-_.*()?+[]\\^$|/", c) && c != 0) continue;
        if (c == ' ' || c == '\t' || c == '\n' || c == '\f' || c == '\r') continue;  // \s
        return false;
    }
    return true;
}
// Parse, constraint.go:40-81
static bool parse(const std::vector<std::string> &env, std::vector<Constraint> &out) {
    static const char *operators[2] = {"==", "!="};
    out.clear();
    for (const std::string &e : env) {
        bool found = false;
        for (int i = 0; i < 2; i++) {
            size_t at = e.find(operators[i]);
            if (at == std::string::npos) continue;
            std::string part0 = strs::trim_space(e.substr(0, at));
            if (!key_valid(part0)) return false;
            std::string part1 = strs::trim_space(e.substr(at + 2));
            if (!value_valid(part1)) return false;
            out.push_back({part0, i, part1});
            found = true;
            break;
        }
        if (!found) return false;
    }
    return true;
}
// Constraint.Match, constraint.go:84-104
static bool match(const Constraint &c, const std::string &what) {
    bool m = strs::equal_fold(c.exp, what);
    return c.op == eq ? m : !m;
}
static const char *role_string(int role) { return role == api::NodeRoleManager ? "MANAGER" : "WORKER"; }
// NodeMatches, constraint.go:107-207
static bool node_matches(const std::vector<Constraint> &cs, const api::Node &n) {
    for (const Constraint &c : cs) {
        const std::string &k = c.key;
        if (strs::equal_fold(k, "node.id")) { if (!match(c, n.ID)) return false; }
        else if (strs::equal_fold(k, "node.hostname")) {
            if (!n.has_description) { if (!match(c, "")) return false; continue; }
            if (!match(c, n.Description.Hostname)) return false;
        } else if (strs::equal_fold(k, "node.ip")) {
            netx::IP nodeIP = netx::parse_ip(n.Addr);
            netx::IP ip = netx::parse_ip(c.exp);
            if (ip.ok) {
                bool ipEq = netx::equal(ip, nodeIP);
                if ((ipEq && c.op != eq) || (!ipEq && c.op == eq)) return false;
                continue;
            }
            netx::IPNet sub = netx::parse_cidr(c.exp);
            if (sub.ok) {
                bool within = netx::contains(sub, nodeIP);
                if ((within && c.op != eq) || (!within && c.op == eq)) return false;
                continue;
            }
            return false;
        } else if (strs::equal_fold(k, "node.role")) { if (!match(c, role_string(n.role))) return false; }
        else if (strs::equal_fold(k, "node.platform.os")) {
            if (!n.has_description || !n.Description.has_platform) { if (!match(c, "")) return false; continue; }
            if (!match(c, n.Description.platform.OS)) return false;
        } else if (strs::equal_fold(k, "node.platform.arch")) {
            if (!n.has_description || !n.Description.has_platform) { if (!match(c, "")) return false; continue; }
            if (!match(c, n.Description.platform.Architecture)) return false;
        } else if (k.size() > NodeLabelPrefix.size() && strs::equal_fold(k.substr(0, NodeLabelPrefix.size()), NodeLabelPrefix)) {
            if (!n.has_labels) { if (!match(c, "")) return false; continue; }
            auto it = n.Labels.find(k.substr(NodeLabelPrefix.size()));
            if (!match(c, it == n.Labels.end() ? "" : it->second)) return false;
        } else if (k.size() > EngineLabelPrefix.size() && strs::equal_fold(k.substr(0, EngineLabelPrefix.size()), EngineLabelPrefix)) {
            if (!n.has_description || !n.Description.has_engine || !n.Description.engine.has_labels) { if (!match(c, "")) return false; continue; }
            auto it = n.Description.engine.Labels.find(k.substr(EngineLabelPrefix.size()));
            if (!match(c, it == n.Description.engine.Labels.end() ? "" : it->second)) return false;
        } else return false;
    }
    return true;
}
}  // namespace constraint

// =========================================================== genericresource
namespace genericresource {
using api::GenericResource;
using List = std::vector<GenericResource>;

static List get_resource(const std::string &kind, const List &res) {  // helpers.go:43-55
    List out;
    for (auto &r : res) if (r.kind == kind) out.push_back(r);
    return out;
}
// HasEnough, validate.go:24-51 (task wants are always discrete)
static bool has_enough(const List &nodeRes, const GenericResource &t) {
    if (t.named) return false;
    List nrs = get_resource(t.kind, nodeRes);
    if (nrs.empty()) return false;
    if (!nrs[0].named) { if (t.amount > nrs[0].amount) return false; }
    else { if (t.amount > (int64_t)nrs.size()) return false; }
    return true;
}
// remove, helpers.go:86-111
static bool remove_one(GenericResource &na, const GenericResource &r) {
    if (!r.named) {
        if (na.named) return false;
        na.amount -= r.amount;
        return na.amount <= 0;
    }
    if (!na.named) return false;
    return r.value == na.value;
}
// ConsumeNodeResources, helpers.go:58-83
static void consume_node_resources(List &avail, const List &res) {
    List out;
    for (auto na : avail) {
        bool removed = false;
        for (auto &r : res) {
            if (na.kind != r.kind) continue;
            if (remove_one(na, r)) { removed = true; break; }
        }
        if (!removed) out.push_back(na);
    }
    avail.swap(out);
}
// selectNodeResources, resource_management.go:42-72
static bool select_node_resources(const List &nodeRes, const GenericResource &tr, List &nrs) {
    nrs.clear();
    for (auto &res : nodeRes) {
        if (res.kind != tr.kind) continue;
        if (!res.named) {
            if (res.amount >= tr.amount && tr.amount != 0) { GenericResource d; d.kind = tr.kind; d.amount = tr.amount; nrs.push_back(d); }
            return true;
        }
        nrs.push_back(res);
        if ((int64_t)nrs.size() == tr.amount) return true;
    }
    return !nrs.empty();
}
// Claim, resource_management.go:11-40
static bool claim(List &avail, List &taskAssigned, const List &reservations) {
    List selected;
    for (auto &res : reservations) {
        if (res.named) return false;
        List nrs;
        if (!select_node_resources(avail, res, nrs)) return false;
        selected.insert(selected.end(), nrs.begin(), nrs.end());
    }
    taskAssigned.insert(taskAssigned.end(), selected.begin(), selected.end());
    consume_node_resources(avail, selected);
    return true;
}
// sanitizeResource / sanitize, resource_management.go:135-205
static bool sanitize_resource(const List &nodeRes, const GenericResource &res, List &nrs) {
    nrs = get_resource(res.kind, nodeRes);
    if (!res.named) {
        if (nrs.size() != 1) return false;
        if (nrs[0].named) return false;
        if (res.amount > nrs[0].amount) return false;
        return true;
    }
    if (nrs.empty()) return false;
    for (auto &nr : nrs) {
        if (!nr.named) return false;
        if (res.value == nr.value) { nrs.clear(); return true; }
    }
    nrs.clear();
    return false;
}
static void sanitize(const List &nodeRes, List &avail) {
    List sanitized, kept;
    std::set<std::string> kindSanitized;
    for (auto &na : avail) {
        List nrs;
        if (!sanitize_resource(nodeRes, na, nrs)) {
            if (kindSanitized.count(na.kind)) continue;
            kindSanitized.insert(na.kind);
            sanitized.insert(sanitized.end(), nrs.begin(), nrs.end());
            continue;
        }
        kept.push_back(na);
    }
    kept.insert(kept.end(), sanitized.begin(), sanitized.end());
    avail.swap(kept);
}
// reclaimResources, resource_management.go:86-122
static void reclaim_resources(List &avail, const List &taskAssigned) {
    for (auto &res : taskAssigned) {
        if (!res.named) {
            List nrs = get_resource(res.kind, avail);
            // "If the resource went down to 0 it's no longer in the available list"
            if (nrs.empty()) avail.push_back(res);
            if (nrs.size() != 1) continue;   // (also taken right after the append above: len(nrs) was 0)
            for (auto &na : avail)
                if (na.kind == res.kind) { if (!na.named) na.amount += res.amount; break; }
        } else {
            avail.push_back(res);
        }
    }
}
// Reclaim, resource_management.go:75-84
static void reclaim(List &avail, const List &taskAssigned, const List &nodeRes) {
    reclaim_resources(avail, taskAssigned);
    sanitize(nodeRes, avail);
}
}  // namespace genericresource

// ================================================================== NodeInfo
static const int64_t monitorFailures = 5LL * 60 * 1000000000LL;  // scheduler.go:19
static const int maxFailures = 5;                                 // scheduler.go:23

struct VersionedService {
    std::string serviceID; uint64_t specVersion = 0;
    bool operator<(const VersionedService &o) const { return serviceID != o.serviceID ? serviceID < o.serviceID : specVersion < o.specVersion; }
};
struct HostPortSpec { int protocol; uint32_t port; bool operator<(const HostPortSpec &o) const { return protocol != o.protocol ? protocol < o.protocol : port < o.port; } };

struct NodeInfo {
    NodeP Node;
    std::map<std::string, TaskP> Tasks;
    int ActiveTasksCount = 0;
    std::map<std::string, int> ActiveTasksCountByService;
    api::Resources AvailableResources;
    std::set<HostPortSpec> usedHostPorts;
    std::map<VersionedService, std::vector<int64_t>> recentFailures;
    int64_t lastCleanup = 0;

    static api::Resources taskReservations(const api::Task &t) { return t.has_reservations ? t.Reservations : api::Resources(); }  // nodeinfo.go:156-161

    int svcCount(const std::string &s) const { auto it = ActiveTasksCountByService.find(s); return it == ActiveTasksCountByService.end() ? 0 : it->second; }

    // removeTask, nodeinfo.go:66-104
    bool removeTask(const api::Task &t) {
        auto it = Tasks.find(t.ID);
        if (it == Tasks.end()) return false;
        TaskP oldTask = it->second;
        Tasks.erase(it);
        if (oldTask->DesiredState <= api::TaskStateCompleted) { ActiveTasksCount--; ActiveTasksCountByService[t.ServiceID]--; }
        if (t.has_endpoint)
            for (auto &port : t.Ports)
                if (port.publish_mode == api::PublishModeHost && port.PublishedPort != 0) usedHostPorts.erase({port.protocol, port.PublishedPort});
        api::Resources r = taskReservations(t);
        AvailableResources.MemoryBytes += r.MemoryBytes;
        AvailableResources.NanoCPUs += r.NanoCPUs;
        if (!Node->has_description || !Node->Description.has_resources) return true;
        // (Generic == nil has no separate encoding here: an empty list sanitises to empty)
        genericresource::reclaim(AvailableResources.Generic, t.AssignedGenericResources, Node->Description.resources.Generic);
        return true;
    }
    // addTask, nodeinfo.go:108-154
    bool addTask(const TaskP &t) {
        auto it = Tasks.find(t->ID);
        if (it != Tasks.end()) {
            TaskP oldTask = it->second;
            if (t->DesiredState <= api::TaskStateCompleted && oldTask->DesiredState > api::TaskStateCompleted) {
                Tasks[t->ID] = t; ActiveTasksCount++; ActiveTasksCountByService[t->ServiceID]++; return true;
            } else if (t->DesiredState > api::TaskStateCompleted && oldTask->DesiredState <= api::TaskStateCompleted) {
                Tasks[t->ID] = t; ActiveTasksCount--; ActiveTasksCountByService[t->ServiceID]--; return true;
            }
            return false;
        }
        Tasks[t->ID] = t;
        api::Resources r = taskReservations(*t);
        AvailableResources.MemoryBytes -= r.MemoryBytes;
        AvailableResources.NanoCPUs -= r.NanoCPUs;
        t->AssignedGenericResources.clear();
        genericresource::claim(AvailableResources.Generic, t->AssignedGenericResources, r.Generic);  // error ignored (:137)
        if (t->has_endpoint)
            for (auto &port : t->Ports)
                if (port.publish_mode == api::PublishModeHost && port.PublishedPort != 0) usedHostPorts.insert({port.protocol, port.PublishedPort});
        if (t->DesiredState <= api::TaskStateCompleted) { ActiveTasksCount++; ActiveTasksCountByService[t->ServiceID]++; }
        return true;
    }
    // cleanupFailures, nodeinfo.go:163-174
    void cleanupFailures(int64_t now) {
        for (auto it = recentFailures.begin(); it != recentFailures.end();) {
            bool recent = false;
            for (int64_t ts : it->second) if (now - ts < monitorFailures) { recent = true; break; }
            if (recent) ++it; else it = recentFailures.erase(it);
        }
        lastCleanup = now;
    }
    // taskFailed, nodeinfo.go:177-202
    void taskFailed(const api::Task &t, int64_t now) {
        if (now - lastCleanup >= monitorFailures) cleanupFailures(now);
        VersionedService vs; vs.serviceID = t.ServiceID; if (t.has_spec_version) vs.specVersion = t.SpecVersion;
        auto &lst = recentFailures[vs];
        size_t expired = 0;
        for (int64_t ts : lst) { if (now - ts < monitorFailures) break; expired++; }
        lst.erase(lst.begin(), lst.begin() + expired);
        lst.push_back(now);
    }
    // countRecentFailures, nodeinfo.go:206-221
    int countRecentFailures(int64_t now, const api::Task &t) const {
        VersionedService vs; vs.serviceID = t.ServiceID; if (t.has_spec_version) vs.specVersion = t.SpecVersion;
        auto it = recentFailures.find(vs);
        if (it == recentFailures.end()) return 0;
        int n = (int)it->second.size();
        for (int i = n - 1; i >= 0; i--)
            if (now - it->second[i] > monitorFailures) { n -= i + 1; break; }
        return n;
    }
};

static NodeInfo newNodeInfo(const NodeP &n, const std::vector<TaskP> &tasks, const api::Resources &avail, int64_t now) {  // nodeinfo.go:46-62
    NodeInfo ni; ni.Node = n; ni.AvailableResources = avail; ni.lastCleanup = now;
    for (auto &t : tasks) ni.addTask(t);
    return ni;
}

// =================================================================== filters
struct Filter {
    virtual ~Filter() {}
    virtual bool SetTask(const api::Task &t) = 0;
    virtual bool Check(const NodeInfo &n) = 0;
    virtual std::string Explain(int nodes) = 0;
};
static std::string plural(int nodes, const char *one, const char *many_fmt) {
    if (nodes == 1) return one;
    char b[128]; snprintf(b, sizeof b, many_fmt, nodes); return b;
}
struct ReadyFilter : Filter {  // filter.go:30-51
    bool SetTask(const api::Task &) override { return true; }
    bool Check(const NodeInfo &n) override { return n.Node->state == api::NodeStatus_READY && n.Node->availability == api::NodeAvailabilityActive; }
    std::string Explain(int nodes) override { return plural(nodes, "1 node not available for new tasks", "%d nodes not available for new tasks"); }
};
struct ResourceFilter : Filter {  // filter.go:55-101
    api::Resources res;
    bool SetTask(const api::Task &t) override {
        if (!t.has_reservations) return false;
        if (t.Reservations.NanoCPUs == 0 && t.Reservations.MemoryBytes == 0 && t.Reservations.Generic.empty()) return false;
        res = t.Reservations; return true;
    }
    bool Check(const NodeInfo &n) override {
        if (res.NanoCPUs > n.AvailableResources.NanoCPUs) return false;
        if (res.MemoryBytes > n.AvailableResources.MemoryBytes) return false;
        for (auto &v : res.Generic) if (!genericresource::has_enough(n.AvailableResources.Generic, v)) return false;
        return true;
    }
    std::string Explain(int nodes) override { return plural(nodes, "insufficient resources on 1 node", "insufficient resources on %d nodes"); }
};
struct PluginFilter : Filter {  // filter.go:104-216
    const api::Task *t = nullptr;
    static bool referencesVolumePlugin(const api::Mount &m) { return m.type == api::MountTypeVolume && m.has_driver && m.driver_name != "" && m.driver_name != "local"; }
    bool SetTask(const api::Task &task) override {
        bool volumeTemplates = false;
        if (task.has_container) for (auto &m : task.Mounts) if (referencesVolumePlugin(m)) { volumeTemplates = true; break; }
        if ((task.has_container && volumeTemplates) || !task.Networks.empty() || task.has_log_driver) { t = &task; return true; }
        return false;
    }
    static std::pair<bool, bool> pluginExistsOnNode(const std::string &type, const std::string &name, const std::vector<api::PluginDescription> &plugins) {
        bool typeFound = false;
        for (auto &np : plugins) {
            if (type != np.Type) continue;
            typeFound = true;
            if (name == np.Name) return {true, true};
            if (strs::has_prefix(np.Name, name) && np.Name.substr(name.size()) == ":latest") return {true, true};
        }
        return {typeFound, false};
    }
    bool Check(const NodeInfo &n) override {
        if (!n.Node->has_description || !n.Node->Description.has_engine) return true;
        auto &plugins = n.Node->Description.engine.Plugins;
        if (t->has_container)
            for (auto &m : t->Mounts)
                if (referencesVolumePlugin(m) && !pluginExistsOnNode("Volume", m.driver_name, plugins).second) return false;
        for (auto &tn : t->Networks)
            if (tn.has_driver && tn.driver != "" && !pluginExistsOnNode("Network", tn.driver, plugins).second) return false;
        if (t->has_log_driver && t->LogDriver != "none" && t->LogDriver != "") {
            auto r = pluginExistsOnNode("Log", t->LogDriver, plugins);
            if (!r.second && r.first) return false;
        }
        return true;
    }
    std::string Explain(int nodes) override { return plural(nodes, "missing plugin on 1 node", "missing plugin on %d nodes"); }
};
struct ConstraintFilter : Filter {  // filter.go:219-251
    std::vector<constraint::Constraint> constraints;
    bool SetTask(const api::Task &t) override {
        if (!t.has_placement || t.placement.Constraints.empty()) return false;
        return constraint::parse(t.placement.Constraints, constraints);  // a parse error disables the filter (:229-236)
    }
    bool Check(const NodeInfo &n) override { return constraint::node_matches(constraints, *n.Node); }
    std::string Explain(int nodes) override { return plural(nodes, "scheduling constraints not satisfied on 1 node", "scheduling constraints not satisfied on %d nodes"); }
};
struct PlatformFilter : Filter {  // filter.go:254-320
    std::vector<api::Platform> supported;
    bool SetTask(const api::Task &t) override {
        supported.clear();
        if (t.has_placement) { supported = t.placement.Platforms; if (!supported.empty()) return true; }
        return false;
    }
    static bool platformEqual(api::Platform img, api::Platform node) {
        if (img.Architecture == "x86_64") img.Architecture = "amd64";
        if (node.Architecture == "x86_64") node.Architecture = "amd64";
        if (img.Architecture == "aarch64") img.Architecture = "arm64";
        if (node.Architecture == "aarch64") node.Architecture = "arm64";
        return (img.Architecture == "" || img.Architecture == node.Architecture) && (img.OS == "" || img.OS == node.OS);
    }
    bool Check(const NodeInfo &n) override {
        if (supported.empty()) return true;
        if (n.Node->has_description && n.Node->Description.has_platform)
            for (auto &p : supported) if (platformEqual(p, n.Node->Description.platform)) return true;
        return false;
    }
    std::string Explain(int nodes) override { return plural(nodes, "unsupported platform on 1 node", "unsupported platform on %d nodes"); }
};
struct HostPortFilter : Filter {  // filter.go:323-361
    const api::Task *t = nullptr;
    bool SetTask(const api::Task &task) override {
        if (task.has_endpoint)
            for (auto &p : task.Ports) if (p.publish_mode == api::PublishModeHost && p.PublishedPort != 0) { t = &task; return true; }
        return false;
    }
    bool Check(const NodeInfo &n) override {
        for (auto &p : t->Ports)
            if (p.publish_mode == api::PublishModeHost && p.PublishedPort != 0 && n.usedHostPorts.count({p.protocol, p.PublishedPort})) return false;
        return true;
    }
    std::string Explain(int nodes) override { return plural(nodes, "host-mode port already in use on 1 node", "host-mode port already in use on %d nodes"); }
};
struct MaxReplicasFilter : Filter {  // filter.go:364-386
    const api::Task *t = nullptr;
    bool SetTask(const api::Task &task) override {
        if (task.has_placement && task.placement.MaxReplicas > 0) { t = &task; return true; }
        return false;
    }
    bool Check(const NodeInfo &n) override { return (uint64_t)n.svcCount(t->ServiceID) < t->placement.MaxReplicas; }
    std::string Explain(int) override { return "max replicas per node limit exceed"; }
};

// volumeSet, volumes.go:45-327; IsInTopology, topology.go:22-47.  Canonicalisation: the volumes of a group are visited in
// ascending volume ID (the reference ranges over a Go map, volumes.go:233).
struct VolumeSet {
    struct Usage { std::string nodeID; bool readOnly = false; };
    struct Info { api::Volume volume; std::map<std::string, Usage> tasks; std::map<std::string, int> nodes; };
    std::map<std::string, Info> volumes;
    std::map<std::string, std::set<std::string>> byGroup;
    std::map<std::string, std::string> byName;

    static bool IsInTopology(const api::Topology &top, const std::vector<std::map<std::string, std::string>> &accessible) {
        if (!top.present || accessible.empty()) return true;
        for (auto &topology : accessible) {
            bool all = true;
            for (auto &kv : topology) {
                auto f = top.Segments.find(kv.first);
                if ((f == top.Segments.end() ? std::string() : f->second) != kv.second) { all = false; break; }
            }
            if (all) return true;
        }
        return false;
    }
    void addOrUpdateVolume(const api::Volume &v) {                       // :61-81 (an update does not replace the stored spec: :69-70)
        if (!volumes.count(v.ID)) { Info i; i.volume = v; volumes[v.ID] = i; }
        byGroup[v.Group].insert(v.ID);
        byName[v.Name] = v.ID;
    }
    void removeVolume(const std::string &id) {                           // :83-96
        auto it = volumes.find(id);
        if (it == volumes.end()) return;
        byGroup[it->second.volume.Group].erase(id);
        byName.erase(it->second.volume.Name);
        volumes.erase(it);
    }
    void reserveVolume(const std::string &vid, const std::string &taskID, const std::string &nodeID, bool readOnly) {   // :150-160
        auto it = volumes.find(vid);
        if (it == volumes.end()) return;
        it->second.tasks[taskID] = Usage{nodeID, readOnly};
        it->second.nodes[nodeID] += 1;
    }
    void releaseVolume(const std::string &vid, const std::string &taskID) {                                           // :162-184
        auto it = volumes.find(vid);
        if (it == volumes.end()) return;
        auto u = it->second.tasks.find(taskID);
        if (u == it->second.tasks.end()) return;
        int &c = it->second.nodes[u->second.nodeID];
        if (c > 0) c -= 1;
        it->second.tasks.erase(u);
    }
    bool checkVolume(const std::string &id, const NodeInfo &info, bool readOnly) const {                              // :257-318
        auto it = volumes.find(id);
        if (it == volumes.end()) return false;
        const Info &vi = it->second;
        if (vi.volume.Availability != api::VolumeAvailabilityActive) return false;
        api::Topology top;
        if (info.Node->has_description)
            for (auto &c : info.Node->Description.CSIInfo) if (c.PluginName == vi.volume.Driver) { top = c.AccessibleTopology; break; }
        if (vi.volume.Scope == api::VolumeScopeSingleNode)
            for (auto &kv : vi.tasks) if (kv.second.nodeID != info.Node->ID) return false;
        switch (vi.volume.Sharing) {
            case api::VolumeSharingNone: if (!vi.tasks.empty()) return false; break;
            case api::VolumeSharingOneWriter: { bool writer = false; for (auto &kv : vi.tasks) writer |= !kv.second.readOnly; if (!readOnly && writer) return false; break; }
            case api::VolumeSharingReadOnly: if (!readOnly) return false; break;
            default: break;
        }
        return IsInTopology(top, vi.volume.has_info ? vi.volume.AccessibleTopology : std::vector<std::map<std::string, std::string>>());
    }
    std::string isVolumeAvailableOnNode(const api::Mount &mount, const NodeInfo &node) const {                        // :223-255
        const std::string &source = mount.Source;
        if (source.compare(0, 6, "group:") == 0) {
            auto g = byGroup.find(source.substr(6));
            if (g == byGroup.end()) return "";
            for (auto &id : g->second) if (checkVolume(id, node, mount.ReadOnly)) return id;
            return "";
        }
        auto n = byName.find(source);
        if (n == byName.end() || !checkVolume(n->second, node, mount.ReadOnly)) return "";
        return n->second;
    }
    // chooseTaskVolumes, :98-136 (what it reserves while choosing is released again before it returns)
    bool chooseTaskVolumes(const api::Task &task, const NodeInfo &node, std::vector<api::VolumeAttachment> &out, std::string &err) {
        std::vector<api::VolumeAttachment> chosen;
        bool ok = true;
        if (task.has_container)
            for (auto &m : task.Mounts) {
                if (m.type != api::MountTypeCluster) continue;
                std::string cand = isVolumeAvailableOnNode(m, node);
                if (cand.empty()) { err = "cannot find volume to satisfy mount with source " + m.Source; ok = false; break; }
                reserveVolume(cand, task.ID, node.Node->ID, m.ReadOnly);
                chosen.push_back({cand, m.Source, m.Target});
            }
        for (auto &a : chosen) releaseVolume(a.ID, task.ID);
        if (ok) out = chosen; else out.clear();
        return ok;
    }
    void reserveTaskVolumes(const api::Task &task) {                                                                  // :138-148
        for (auto &va : task.Volumes)
            for (auto &m : task.Mounts)
                if (m.Source == va.Source && m.Target == va.Target) reserveVolume(va.ID, task.ID, task.NodeID, m.ReadOnly);
    }
};
struct VolumesFilter : Filter {  // filter.go:388-447 (appended to the pipeline at scheduler.go:132)
    const VolumeSet *vs = nullptr;
    std::vector<api::Mount> requestedVolumes;
    bool SetTask(const api::Task &t) override {
        requestedVolumes.clear();
        if (!vs || !t.has_container) return false;
        bool hasCSI = false;
        for (auto &m : t.Mounts) if (m.type == api::MountTypeCluster) { hasCSI = true; requestedVolumes.push_back(m); }
        return hasCSI;
    }
    bool Check(const NodeInfo &n) override {       // true as soon as ONE requested mount can be met (:432-440)
        for (auto &m : requestedVolumes) if (vs->isVolumeAvailableOnNode(m, n) != "") return true;
        return false;
    }
    std::string Explain(int nodes) override { return plural(nodes, "cannot fulfill requested CSI volume mounts on 1 node", "cannot fulfill requested CSI volume mounts on %d nodes"); }
};

// Pipeline, pipeline.go:38-103
struct Pipeline {
    struct Entry { std::unique_ptr<Filter> f; bool enabled = false; int failureCount = 0; };
    std::vector<Entry> checklist;
    Pipeline() {
        auto add = [&](Filter *f) { Entry e; e.f.reset(f); checklist.push_back(std::move(e)); };
        add(new ReadyFilter()); add(new ResourceFilter()); add(new PluginFilter()); add(new ConstraintFilter());
        add(new PlatformFilter()); add(new HostPortFilter()); add(new MaxReplicasFilter());
    }
    void AddFilter(Filter *f) { Entry e; e.f.reset(f); checklist.push_back(std::move(e)); }   // pipeline.go:70-72
    bool Process(const NodeInfo &n) {
        for (auto &e : checklist)
            if (e.enabled && !e.f->Check(n)) { e.failureCount++; return false; }
        for (auto &e : checklist) e.failureCount = 0;
        return true;
    }
    void SetTask(const api::Task &t) { for (auto &e : checklist) { e.enabled = e.f->SetTask(t); e.failureCount = 0; } }
    std::string Explain() {
        // sort.Sort(sort.Reverse(byFailures)) on 8 entries (< 12) = insertion sort: stable, most failures first
        std::vector<const Entry *> s;
        for (auto &e : checklist) s.push_back(&e);
        for (size_t i = 1; i < s.size(); i++)
            for (size_t j = i; j > 0 && s[j]->failureCount > s[j - 1]->failureCount; j--) std::swap(s[j], s[j - 1]);
        std::string out;
        for (auto *e : s)
            if (e->failureCount > 0) { if (!out.empty()) out += "; "; out += e->f->Explain(e->failureCount); }
        return out;
    }
    std::vector<int> counts() const { std::vector<int> c; for (auto &e : checklist) c.push_back(e.failureCount); return c; }
};

// ============================================================= decision tree
using NodeLess = std::function<bool(const NodeInfo &, const NodeInfo &)>;
struct DecisionTree {
    int tasks = 0;
    std::map<std::string, std::unique_ptr<DecisionTree>> next;  // canonical: ascending label value
    bool has_next = false;
    std::vector<NodeInfo> heap;  // nodeMaxHeap.nodes
    size_t heapLen = 0;          // nodeMaxHeap.length
};

// ================================================================= Scheduler
struct Decision { TaskP old_, new_; };

struct Scheduler {
    std::map<std::string, TaskP> unassignedTasks, pendingPreassignedTasks, allTasks;
    std::set<std::string> preassignedTasks;
    std::map<std::string, NodeInfo> nodeSet;  // canonical node order = ascending ID
    std::map<std::string, std::pair<bool, uint64_t>> services;  // store view: id -> (has SpecVersion, index)
    Pipeline pipeline;
    VolumeSet volumes;
    int64_t now = 0;
    Scheduler() { VolumesFilter *vf = new VolumesFilter(); vf->vs = &volumes; pipeline.AddFilter(vf); }   // scheduler.go:126-134
    Scheduler(const Scheduler &) = delete;

    // canonical total order: nodeLess (scheduler.go:708-735) + node ID
    static bool lessWithTie(const NodeLess &nodeLess, const NodeInfo &a, const NodeInfo &b) {
        if (nodeLess(a, b)) return true;
        if (nodeLess(b, a)) return false;
        return a.Node->ID < b.Node->ID;
    }

    void enqueue(const TaskP &t) { unassignedTasks[t->ID] = t; }

    // setupTasksList + buildNodeSet, scheduler.go:68-125,973-990
    void setup(const std::vector<NodeP> &nodes, const std::vector<TaskP> &tasks, const std::vector<api::Volume> &vols = {}) {
        // only volumes that have been created with their plugin, i.e. carry a VolumeID (scheduler.go:75-81)
        for (auto &v : vols) if (v.has_info && v.VolumeID != "") volumes.addOrUpdateVolume(v);
        std::map<std::string, std::vector<TaskP>> tasksByNode;
        for (auto &t : tasks) {
            if (t->state < api::TaskStatePending || t->state > api::TaskStateRunning) continue;
            if (t->state == api::TaskStatePending && t->DesiredState > api::TaskStateCompleted) continue;
            allTasks[t->ID] = t;
            if (t->NodeID == "") { enqueue(t); continue; }
            if (t->state == api::TaskStatePending) { preassignedTasks.insert(t->ID); pendingPreassignedTasks[t->ID] = t; continue; }
            volumes.reserveTaskVolumes(*t);          // track the volumes in use by the task (scheduler.go:115-116)
            tasksByNode[t->NodeID].push_back(t);
        }
        for (auto &n : nodes) {
            api::Resources r;
            if (n->has_description && n->Description.has_resources) r = n->Description.resources;
            nodeSet[n->ID] = newNodeInfo(n, tasksByNode[n->ID], r, now);
        }
    }
    // EventUpdateVolume, scheduler.go:205-217 (there is no create case: a volume is usable once its plugin created it)
    void updateVolume(const api::Volume &v) { if (v.has_info && v.VolumeID != "") volumes.addOrUpdateVolume(v); }
    // createTask, scheduler.go:254-281
    void createTask(const TaskP &t) {
        if (t->state < api::TaskStatePending || t->state > api::TaskStateRunning) return;
        allTasks[t->ID] = t;
        if (t->NodeID == "") { enqueue(t); return; }
        if (t->state == api::TaskStatePending) { preassignedTasks.insert(t->ID); pendingPreassignedTasks[t->ID] = t; return; }
        auto it = nodeSet.find(t->NodeID);
        if (it != nodeSet.end()) it->second.addTask(t);
    }
    // deleteTask, scheduler.go:350-366
    void deleteTask(const api::Task &t) {
        allTasks.erase(t.ID); preassignedTasks.erase(t.ID); pendingPreassignedTasks.erase(t.ID);
        for (auto &va : t.Volumes) volumes.releaseVolume(va.ID, t.ID);      // the task's volume reservations, if any (:355-358)
        auto it = nodeSet.find(t.NodeID);
        if (it != nodeSet.end()) it->second.removeTask(t);
    }
    // updateTask, scheduler.go:283-348
    void updateTask(const TaskP &t) {
        if (t->state < api::TaskStatePending) return;
        TaskP oldTask; auto ot = allTasks.find(t->ID); if (ot != allTasks.end()) oldTask = ot->second;
        if (t->state > api::TaskStateRunning) {
            if (!oldTask) return;
            if (t->state != oldTask->state && (t->state == api::TaskStateFailed || t->state == api::TaskStateRejected)) {
                if (!preassignedTasks.count(t->ID)) {
                    auto it = nodeSet.find(t->NodeID);
                    if (it != nodeSet.end()) it->second.taskFailed(*t, now);
                }
            }
            deleteTask(*oldTask);
            return;
        }
        if (t->NodeID == "") {
            if (oldTask) deleteTask(*oldTask);
            allTasks[t->ID] = t; enqueue(t); return;
        }
        if (t->state == api::TaskStatePending) {
            if (oldTask) deleteTask(*oldTask);
            preassignedTasks.insert(t->ID); allTasks[t->ID] = t; pendingPreassignedTasks[t->ID] = t; return;
        }
        allTasks[t->ID] = t;
        auto it = nodeSet.find(t->NodeID);
        if (it != nodeSet.end()) it->second.addTask(t);
    }
    // createOrUpdateNode, scheduler.go:368-396
    void createOrUpdateNode(const NodeP &n) {
        auto it = nodeSet.find(n->ID);
        api::Resources resources;
        if (n->has_description && n->Description.has_resources) {
            resources = n->Description.resources;
            if (it != nodeSet.end())
                for (auto &kv : it->second.Tasks) {
                    api::Resources r = NodeInfo::taskReservations(*kv.second);
                    resources.MemoryBytes -= r.MemoryBytes;
                    resources.NanoCPUs -= r.NanoCPUs;
                    genericresource::consume_node_resources(resources.Generic, kv.second->AssignedGenericResources);
                }
        }
        if (it == nodeSet.end()) nodeSet[n->ID] = newNodeInfo(n, {}, resources, now);
        else { it->second.Node = n; it->second.AvailableResources = resources; }
    }

    // ---- nodeSet.tree, nodeset.go:50-124
    DecisionTree tree(const std::string &serviceID, const std::vector<std::string> &prefs, int maxAssignments, const NodeLess &nodeLess) {
        DecisionTree root;
        if (maxAssignments == 0) return root;
        auto heapLess = [&](const NodeInfo &a, const NodeInfo &b) { return lessWithTie(nodeLess, a, b); };  // max-heap on "less"
        for (auto &kv : nodeSet) {
            const NodeInfo &node = kv.second;
            DecisionTree *t = &root;
            for (auto &descriptor : prefs) {
                std::string value;
                if (descriptor.size() > constraint::NodeLabelPrefix.size() && strs::equal_fold(descriptor.substr(0, constraint::NodeLabelPrefix.size()), constraint::NodeLabelPrefix)) {
                    if (node.Node->has_labels) { auto f = node.Node->Labels.find(descriptor.substr(constraint::NodeLabelPrefix.size())); if (f != node.Node->Labels.end()) value = f->second; }
                } else if (descriptor.size() > constraint::EngineLabelPrefix.size() && strs::equal_fold(descriptor.substr(0, constraint::EngineLabelPrefix.size()), constraint::EngineLabelPrefix)) {
                    if (node.Node->has_description && node.Node->Description.has_engine && node.Node->Description.engine.has_labels) {
                        auto f = node.Node->Description.engine.Labels.find(descriptor.substr(constraint::EngineLabelPrefix.size()));
                        if (f != node.Node->Description.engine.Labels.end()) value = f->second;
                    }
                } else continue;
                t->tasks += node.svcCount(serviceID);
                t->has_next = true;
                auto &nx = t->next[value];
                if (!nx) nx.reset(new DecisionTree());
                t = nx.get();
            }
            t->tasks += node.svcCount(serviceID);
            if ((int)t->heap.size() < maxAssignments) {
                if (pipeline.Process(node)) { t->heap.push_back(node); std::push_heap(t->heap.begin(), t->heap.end(), heapLess); }
            } else if (heapLess(node, t->heap.front())) {
                if (pipeline.Process(node)) {
                    std::pop_heap(t->heap.begin(), t->heap.end(), heapLess);
                    t->heap.back() = node;
                    std::push_heap(t->heap.begin(), t->heap.end(), heapLess);
                }
            }
            t->heapLen = t->heap.size();
        }
        return root;
    }
    // decisionTree.orderedNodes, decision_tree.go:24-52
    std::vector<NodeInfo> &orderedNodes(DecisionTree &dt, const NodeLess &nodeLess) {
        auto heapLess = [&](const NodeInfo &a, const NodeInfo &b) { return lessWithTie(nodeLess, a, b); };
        if (dt.heapLen != dt.heap.size()) {
            for (size_t i = 0; i < dt.heap.size();) {
                // refresh from the live node set first: the reference's copies share maps/pointers with it
                if (pipeline.Process(dt.heap[i])) i++;
                else { dt.heap[i] = dt.heap.back(); dt.heap.pop_back(); }
            }
            dt.heapLen = dt.heap.size();
            std::make_heap(dt.heap.begin(), dt.heap.end(), heapLess);
        }
        std::sort_heap(dt.heap.begin(), dt.heap.begin() + dt.heapLen, heapLess);
        dt.heapLen = 0;
        return dt.heap;
    }

    // scheduleNTasksOnNodes, scheduler.go:844-924
    int scheduleNTasksOnNodes(int n, std::map<std::string, TaskP> &taskGroup, std::vector<NodeInfo> &nodes, std::map<std::string, Decision> &decisions, const NodeLess &nodeLess) {
        int tasksScheduled = 0;
        std::map<int, bool> failedConstraints;
        int nodeIter = 0, nodeCount = (int)nodes.size();
        std::vector<std::string> ids;
        for (auto &kv : taskGroup) ids.push_back(kv.first);  // canonical: ascending task ID
        for (auto &taskID : ids) {
            TaskP t = taskGroup[taskID];
            if (decisions.count(taskID)) continue;
            NodeInfo &node = nodes[nodeIter % nodeCount];
            // the volume attachments of the task on this node (:857-867; an error is only logged there)
            std::vector<api::VolumeAttachment> attachments; std::string verr;
            volumes.chooseTaskVolumes(*t, node, attachments, verr);
            TaskP newT(new api::Task(*t));
            newT->Volumes = attachments;
            newT->NodeID = node.Node->ID;
            volumes.reserveTaskVolumes(*newT);
            newT->state = api::TaskStateAssigned; newT->Err = ""; newT->Message = "scheduler assigned task to node";
            allTasks[t->ID] = newT;
            auto ns = nodeSet.find(node.Node->ID);
            NodeInfo nodeInfo = node;
            if (ns != nodeSet.end() && ns->second.addTask(newT)) { nodeInfo = ns->second; nodes[nodeIter % nodeCount] = nodeInfo; }
            decisions[taskID] = {t, newT};
            taskGroup.erase(taskID);
            tasksScheduled++;
            if (tasksScheduled == n) return tasksScheduled;
            if (nodeIter + 1 < nodeCount) {
                const NodeInfo &nextNode = nodes[(nodeIter + 1) % nodeCount];
                if (nodeLess(nextNode, nodeInfo)) nodeIter++;
            } else nodeIter++;
            int origNodeIter = nodeIter;
            while (failedConstraints[nodeIter % nodeCount] || !pipeline.Process(nodes[nodeIter % nodeCount])) {
                failedConstraints[nodeIter % nodeCount] = true;
                nodeIter++;
                if (nodeIter - origNodeIter == nodeCount) return tasksScheduled;
            }
        }
        return tasksScheduled;
    }
    // scheduleNTasksOnSubtree, scheduler.go:772-825
    int scheduleNTasksOnSubtree(int n, std::map<std::string, TaskP> &taskGroup, DecisionTree &tree, std::map<std::string, Decision> &decisions, const NodeLess &nodeLess) {
        if (!tree.has_next) {
            auto &nodes = orderedNodes(tree, nodeLess);
            if (nodes.empty()) return 0;
            // the reference's slice holds copies whose maps alias the live NodeInfo: refresh
            for (auto &ni : nodes) { auto f = nodeSet.find(ni.Node->ID); if (f != nodeSet.end()) ni = f->second; }
            return scheduleNTasksOnNodes(n, taskGroup, nodes, decisions, nodeLess);
        }
        int tasksScheduled = 0, tasksInUsableBranches = tree.tasks;
        std::set<DecisionTree *> noRoom;
        bool converging = true;
        while (tasksScheduled != n && noRoom.size() != tree.next.size() && converging) {
            int usable = (int)(tree.next.size() - noRoom.size());
            int desiredTasksPerBranch = (tasksInUsableBranches + n - tasksScheduled) / usable;
            int remainder = (tasksInUsableBranches + n - tasksScheduled) % usable;
            converging = false;
            for (auto &kv : tree.next) {
                DecisionTree *subtree = kv.second.get();
                if (noRoom.count(subtree)) continue;
                int subtreeTasks = subtree->tasks;
                if (subtreeTasks < desiredTasksPerBranch || (subtreeTasks == desiredTasksPerBranch && remainder > 0)) {
                    converging = true;
                    int tasksToAssign = desiredTasksPerBranch - subtreeTasks;
                    if (remainder > 0) tasksToAssign++;
                    int res = scheduleNTasksOnSubtree(tasksToAssign, taskGroup, *subtree, decisions, nodeLess);
                    if (res < tasksToAssign) { noRoom.insert(subtree); tasksInUsableBranches -= subtreeTasks; }
                    else if (remainder > 0) remainder--;
                    tasksScheduled += res;
                }
            }
        }
        return tasksScheduled;
    }
    // noSuitableNode, scheduler.go:928-971
    void noSuitableNode(std::map<std::string, TaskP> &taskGroup, std::map<std::string, Decision> &decisions) {
        std::string explanation = pipeline.Explain();
        for (auto &kv : taskGroup) {
            TaskP t = kv.second;
            auto svc = services.find(t->ServiceID);
            if (svc == services.end()) continue;  // service gone: drop the task from the scheduler
            TaskP newT(new api::Task(*t));
            if (svc->second.first && t->has_spec_version && svc->second.second > t->SpecVersion) {
                if (t->state == api::TaskStatePending && t->DesiredState >= api::TaskStateShutdown) { newT->state = api::TaskStateShutdown; newT->Err = ""; }
            } else {
                newT->Err = explanation != "" ? "no suitable node (" + explanation + ")" : "no suitable node";
                enqueue(newT);
            }
            allTasks[t->ID] = newT;
            decisions[t->ID] = {t, newT};
        }
    }
    // scheduleTaskGroup, scheduler.go:694-748
    void scheduleTaskGroup(std::map<std::string, TaskP> &taskGroup, std::map<std::string, Decision> &decisions) {
        if (taskGroup.empty()) return;
        const api::Task &t = *taskGroup.begin()->second;
        pipeline.SetTask(t);
        int64_t nowc = now;
        NodeLess nodeLess = [&t, nowc](const NodeInfo &a, const NodeInfo &b) {
            int fa = a.countRecentFailures(nowc, t), fb = b.countRecentFailures(nowc, t);
            if (fa >= maxFailures || fb >= maxFailures) {
                if (fa > fb) return false;
                if (fb > fa) return true;
            }
            int sa = a.svcCount(t.ServiceID), sb = b.svcCount(t.ServiceID);
            if (sa < sb) return true;
            if (sa > sb) return false;
            return a.ActiveTasksCount < b.ActiveTasksCount;
        };
        std::vector<std::string> prefs;
        if (t.has_placement) prefs = t.placement.Preferences;
        TaskP keep = taskGroup.begin()->second;  // keep `t` alive while the group map shrinks
        DecisionTree tr = tree(t.ServiceID, prefs, (int)taskGroup.size(), nodeLess);
        scheduleNTasksOnSubtree((int)taskGroup.size(), taskGroup, tr, decisions, nodeLess);
        if (!taskGroup.empty()) noSuitableNode(taskGroup, decisions);
    }
    // taskFitNode, scheduler.go:646-690
    TaskP taskFitNode(const TaskP &t, const std::string &nodeID) {
        auto it = nodeSet.find(nodeID);
        if (it == nodeSet.end()) return nullptr;
        TaskP newT(new api::Task(*t));
        pipeline.SetTask(*t);
        if (!pipeline.Process(it->second)) { newT->Err = pipeline.Explain(); allTasks[t->ID] = newT; return newT; }
        // :664-675: the attachments are chosen here (and NOT reserved: the reference does not call reserveTaskVolumes on this path)
        std::vector<api::VolumeAttachment> attachments; std::string verr;
        if (!volumes.chooseTaskVolumes(*t, it->second, attachments, verr)) { newT->Err = verr; allTasks[t->ID] = newT; return newT; }
        newT->Volumes = attachments;
        newT->state = api::TaskStateAssigned; newT->Err = ""; newT->Message = "scheduler confirmed task can run on preassigned node";
        allTasks[t->ID] = newT;
        it->second.addTask(newT);
        return newT;
    }
    // processPreassignedTasks, scheduler.go:398-426 (commit always succeeds unless listed)
    std::map<std::string, Decision> processPreassignedTasks(const std::set<std::string> &failCommit) {
        std::map<std::string, Decision> decisions;
        std::vector<TaskP> pend;
        for (auto &kv : pendingPreassignedTasks) pend.push_back(kv.second);
        for (auto &t : pend) {
            TaskP newT = taskFitNode(t, t->NodeID);
            if (!newT) continue;
            decisions[t->ID] = {t, newT};
        }
        for (auto &kv : decisions) {
            Decision &d = kv.second;
            if (failCommit.count(kv.first)) {
                allTasks[d.old_->ID] = d.old_;
                auto it = nodeSet.find(d.new_->NodeID);
                if (it != nodeSet.end()) it->second.removeTask(*d.new_);
                for (auto &va : d.new_->Volumes) volumes.releaseVolume(va.ID, d.new_->ID);       // :422-424
            } else if (d.new_->state == api::TaskStateAssigned) pendingPreassignedTasks.erase(d.old_->ID);
        }
        return decisions;
    }
    // tick, scheduler.go:429-488
    std::map<std::string, Decision> tick(const std::set<std::string> &failCommit) {
        typedef std::pair<std::string, uint64_t> Key;
        std::map<Key, std::map<std::string, TaskP>> tasksByCommonSpec;
        std::vector<TaskP> oneOffTasks;
        std::map<std::string, Decision> decisions;
        for (auto it = unassignedTasks.begin(); it != unassignedTasks.end(); it = unassignedTasks.erase(it)) {
            TaskP t = it->second;
            if (!t || t->NodeID != "") continue;
            if (t->has_spec_version) tasksByCommonSpec[Key(t->ServiceID, t->SpecVersion)][t->ID] = t;
            else oneOffTasks.push_back(t);  // map order == ascending task ID
        }
        for (auto &kv : tasksByCommonSpec) scheduleTaskGroup(kv.second, decisions);
        for (auto &t : oneOffTasks) { std::map<std::string, TaskP> g; g[t->ID] = t; scheduleTaskGroup(g, decisions); }
        // applySchedulingDecisions (scheduler.go:490-643) is host/store code; failures roll back (:472-487)
        for (auto &kv : decisions) {
            if (!failCommit.count(kv.first)) continue;
            Decision &d = kv.second;
            allTasks[d.old_->ID] = d.old_;
            auto it = nodeSet.find(d.new_->NodeID);
            if (it != nodeSet.end()) it->second.removeTask(*d.new_);
            for (auto &va : d.new_->Volumes) volumes.releaseVolume(va.ID, d.new_->ID);           // release the volumes we tried to use (:480-483)
            enqueue(d.old_);
        }
        return decisions;
    }
};

// ================================================================ JSON driver
namespace drv {
static int enum_of(const mj::Value &v, const std::vector<std::pair<const char *, int>> &names, int def) {
    if (v.type == mj::Value::Int) return (int)v.i;
    if (v.type == mj::Value::Str) { for (auto &p : names) if (v.s == p.first) return p.second; }
    return def;
}
static const std::vector<std::pair<const char *, int>> TASK_STATES = {
    {"NEW", 0}, {"PENDING", 64}, {"ASSIGNED", 192}, {"ACCEPTED", 256}, {"PREPARING", 320}, {"READY", 384}, {"STARTING", 448}, {"RUNNING", 512},
    {"COMPLETE", 576}, {"COMPLETED", 576}, {"SHUTDOWN", 640}, {"FAILED", 704}, {"REJECTED", 768}, {"REMOVE", 800}, {"ORPHANED", 832}};
static const char *task_state_name(int s) { for (auto &p : TASK_STATES) if (p.second == s) return p.first; return "?"; }

static api::Resources resources(const mj::Value &v) {
    api::Resources r;
    r.NanoCPUs = v.at("nano_cpus").as_int(); r.MemoryBytes = v.at("memory_bytes").as_int();
    for (auto &g : v.at("generic").a) {
        api::GenericResource gr; gr.kind = g.at("kind").as_str();
        if (g.find("named")) { gr.named = true; gr.value = g.at("named").as_str(); } else gr.amount = g.at("value").as_int();
        r.Generic.push_back(gr);
    }
    return r;
}
static mj::Value resources_json(const api::Resources &r) {
    mj::Value o = mj::Value::object();
    o.set("nano_cpus", mj::Value::integer(r.NanoCPUs)); o.set("memory_bytes", mj::Value::integer(r.MemoryBytes));
    mj::Value g = mj::Value::array();
    for (auto &x : r.Generic) {
        mj::Value e = mj::Value::object(); e.set("kind", mj::Value::string(x.kind));
        if (x.named) e.set("named", mj::Value::string(x.value)); else e.set("value", mj::Value::integer(x.amount));
        g.push(e);
    }
    o.set("generic", g);
    return o;
}
static std::map<std::string, std::string> strmap(const mj::Value &v) { std::map<std::string, std::string> m; for (auto &kv : v.o) m[kv.first] = kv.second.as_str(); return m; }

static NodeP node(const mj::Value &v) {
    NodeP n(new api::Node());
    n->ID = v.at("id").as_str();
    n->role = enum_of(v.at("role"), {{"WORKER", 0}, {"MANAGER", 1}}, 0);
    n->version = (uint64_t)v.at("version").as_int();
    const mj::Value &spec = v.at("spec");
    n->availability = enum_of(spec.at("availability"), {{"ACTIVE", 0}, {"PAUSE", 1}, {"DRAIN", 2}}, 0);
    if (spec.find("labels") && !spec.at("labels").is_null()) { n->has_labels = true; n->Labels = strmap(spec.at("labels")); }
    const mj::Value &st = v.at("status");
    n->state = enum_of(st.at("state"), {{"UNKNOWN", 0}, {"DOWN", 1}, {"READY", 2}, {"DISCONNECTED", 3}}, 0);
    n->Addr = st.at("addr").as_str();
    const mj::Value &d = v.at("description");
    if (!d.is_null()) {
        n->has_description = true;
        n->Description.Hostname = d.at("hostname").as_str();
        if (!d.at("platform").is_null()) { n->Description.has_platform = true; n->Description.platform.OS = d.at("platform").at("os").as_str(); n->Description.platform.Architecture = d.at("platform").at("arch").as_str(); }
        if (!d.at("resources").is_null()) { n->Description.has_resources = true; n->Description.resources = resources(d.at("resources")); }
        if (d.find("csi_info") && !d.at("csi_info").is_null())
            for (auto &c : d.at("csi_info").a) {
                api::NodeCSIInfo ci; ci.PluginName = c.at("plugin").as_str(); ci.NodeID = c.at("node_id").as_str();
                if (!c.at("topology").is_null()) { ci.AccessibleTopology.present = true; ci.AccessibleTopology.Segments = strmap(c.at("topology")); }
                n->Description.CSIInfo.push_back(ci);
            }
        const mj::Value &e = d.at("engine");
        if (!e.is_null()) {
            n->Description.has_engine = true;
            if (e.find("labels") && !e.at("labels").is_null()) { n->Description.engine.has_labels = true; n->Description.engine.Labels = strmap(e.at("labels")); }
            for (auto &p : e.at("plugins").a) n->Description.engine.Plugins.push_back({p.at("type").as_str(), p.at("name").as_str()});
        }
    }
    return n;
}
static TaskP task(const mj::Value &v) {
    TaskP t(new api::Task());
    t->ID = v.at("id").as_str(); t->ServiceID = v.at("service_id").as_str(); t->NodeID = v.at("node_id").as_str();
    t->Slot = (uint64_t)v.at("slot").as_int();
    t->DesiredState = enum_of(v.at("desired_state"), TASK_STATES, api::TaskStateRunning);
    if (v.find("desired_state") == nullptr) t->DesiredState = api::TaskStateRunning;
    const mj::Value &st = v.at("status");
    t->state = enum_of(st.at("state"), TASK_STATES, 0); t->Err = st.at("err").as_str(); t->Message = st.at("message").as_str();
    if (v.find("spec_version") && !v.at("spec_version").is_null()) { t->has_spec_version = true; t->SpecVersion = (uint64_t)v.at("spec_version").as_int(); }
    const mj::Value &spec = v.at("spec");
    if (!spec.at("resources").is_null() && !spec.at("resources").at("reservations").is_null()) { t->has_reservations = true; t->Reservations = resources(spec.at("resources").at("reservations")); }
    const mj::Value &pl = spec.at("placement");
    if (!pl.is_null()) {
        t->has_placement = true;
        for (auto &c : pl.at("constraints").a) t->placement.Constraints.push_back(c.as_str());
        for (auto &c : pl.at("preferences").a) t->placement.Preferences.push_back(c.as_str());
        for (auto &p : pl.at("platforms").a) t->placement.Platforms.push_back({p.at("arch").as_str(), p.at("os").as_str()});
        t->placement.MaxReplicas = (uint64_t)pl.at("max_replicas").as_int();
    }
    const mj::Value &ct = spec.at("container");
    if (!ct.is_null()) {
        t->has_container = true;
        for (auto &m : ct.at("mounts").a) {
            api::Mount mm; mm.type = enum_of(m.at("type"), {{"BIND", 0}, {"VOLUME", 1}, {"TMPFS", 2}, {"NPIPE", 3}, {"CLUSTER", 4}}, 0);
            if (!m.at("driver").is_null()) { mm.has_driver = true; mm.driver_name = m.at("driver").as_str(); }
            mm.Source = m.at("source").as_str(); mm.Target = m.at("target").as_str(); mm.ReadOnly = !m.at("read_only").is_null() && m.at("read_only").as_bool();
            t->Mounts.push_back(mm);
        }
    }
    if (!spec.at("log_driver").is_null()) { t->has_log_driver = true; t->LogDriver = spec.at("log_driver").at("name").as_str(); }
    for (auto &nw : v.at("networks").a) { api::Task::Net x; if (!nw.at("driver").is_null()) { x.has_driver = true; x.driver = nw.at("driver").as_str(); } t->Networks.push_back(x); }
    if (!v.at("endpoint").is_null()) {
        t->has_endpoint = true;
        for (auto &p : v.at("endpoint").at("ports").a) {
            api::PortConfig pc; pc.protocol = enum_of(p.at("protocol"), {{"TCP", 0}, {"UDP", 1}, {"SCTP", 2}}, 0);
            pc.PublishedPort = (uint32_t)p.at("published_port").as_int();
            pc.publish_mode = enum_of(p.at("publish_mode"), {{"INGRESS", 0}, {"HOST", 1}}, 0);
            t->Ports.push_back(pc);
        }
    }
    for (auto &g : v.at("assigned_generic").a) {
        api::GenericResource gr; gr.kind = g.at("kind").as_str();
        if (g.find("named")) { gr.named = true; gr.value = g.at("named").as_str(); } else gr.amount = g.at("value").as_int();
        t->AssignedGenericResources.push_back(gr);
    }
    if (v.find("volumes") && !v.at("volumes").is_null())
        for (auto &a : v.at("volumes").a) t->Volumes.push_back({a.at("id").as_str(), a.at("source").as_str(), a.at("target").as_str()});
    return t;
}
static api::Volume volume(const mj::Value &v) {
    api::Volume x;
    x.ID = v.at("id").as_str(); x.Name = v.at("name").as_str(); x.Group = v.at("group").as_str(); x.Driver = v.at("driver").as_str();
    x.Availability = enum_of(v.at("availability"), {{"ACTIVE", 0}, {"PAUSE", 1}, {"DRAIN", 2}}, 0);
    x.Scope = enum_of(v.at("scope"), {{"SINGLE_NODE", 0}, {"MULTI_NODE", 1}}, 0);
    x.Sharing = enum_of(v.at("sharing"), {{"NONE", 0}, {"READ_ONLY", 1}, {"ONE_WRITER", 2}, {"ALL", 3}}, 0);
    if (!v.at("volume_info").is_null()) {
        x.has_info = true; x.VolumeID = v.at("volume_info").at("volume_id").as_str();
        if (!v.at("volume_info").at("accessible_topology").is_null())
            for (auto &t : v.at("volume_info").at("accessible_topology").a) x.AccessibleTopology.push_back(strmap(t));
    }
    return x;
}
static mj::Value decisions_json(const std::map<std::string, Decision> &ds) {
    mj::Value arr = mj::Value::array();
    for (auto &kv : ds) {
        mj::Value d = mj::Value::object();
        d.set("id", mj::Value::string(kv.first));
        d.set("node_id", mj::Value::string(kv.second.new_->NodeID));
        d.set("state", mj::Value::string(task_state_name(kv.second.new_->state)));
        d.set("err", mj::Value::string(kv.second.new_->Err));
        d.set("message", mj::Value::string(kv.second.new_->Message));
        mj::Value ag = mj::Value::array();
        for (auto &x : kv.second.new_->AssignedGenericResources) {
            mj::Value e = mj::Value::object(); e.set("kind", mj::Value::string(x.kind));
            if (x.named) e.set("named", mj::Value::string(x.value)); else e.set("value", mj::Value::integer(x.amount));
            ag.push(e);
        }
        d.set("assigned_generic", ag);
        mj::Value vols = mj::Value::array();
        for (auto &a : kv.second.new_->Volumes) {
            mj::Value e = mj::Value::object(); e.set("id", mj::Value::string(a.ID)); e.set("source", mj::Value::string(a.Source)); e.set("target", mj::Value::string(a.Target));
            vols.push(e);
        }
        if (!kv.second.new_->Volumes.empty()) d.set("volumes", vols);      // (only tasks with cluster mounts carry the key)
        arr.push(d);
    }
    return arr;
}
static std::set<std::string> strset(const mj::Value &v) { std::set<std::string> s; for (auto &x : v.a) s.insert(x.as_str()); return s; }

static mj::Value apply(Scheduler &S, const mj::Value &ev) {
    mj::Value out = mj::Value::object();
    std::string op = ev.at("op").as_str();
    if (ev.find("now_ns")) S.now = ev.at("now_ns").as_int();
    if (op == "init") {
        std::vector<NodeP> nodes; std::vector<TaskP> tasks;
        for (auto &n : ev.at("nodes").a) nodes.push_back(node(n));
        for (auto &t : ev.at("tasks").a) tasks.push_back(task(t));
        for (auto &s : ev.at("services").a) S.services[s.at("id").as_str()] = {!s.at("spec_version").is_null(), (uint64_t)s.at("spec_version").as_int()};
        std::vector<api::Volume> vols;
        if (ev.find("volumes") && !ev.at("volumes").is_null()) for (auto &v : ev.at("volumes").a) vols.push_back(volume(v));
        S.setup(nodes, tasks, vols);
    } else if (op == "update_volume") {
        S.updateVolume(volume(ev.at("volume")));
    } else if (op == "delete_volume") {
        S.volumes.removeVolume(ev.at("id").as_str());
    } else if (op == "volume_usage") {
        mj::Value vols = mj::Value::object();
        for (auto &kv : S.volumes.volumes) {
            mj::Value tasks = mj::Value::object();
            for (auto &t : kv.second.tasks) { mj::Value u = mj::Value::object(); u.set("node", mj::Value::string(t.second.nodeID)); u.set("read_only", mj::Value::boolean(t.second.readOnly)); tasks.set(t.first, u); }
            vols.set(kv.first, tasks);
        }
        out.set("volumes", vols);
    } else if (op == "set_service") {
        S.services[ev.at("id").as_str()] = {!ev.at("spec_version").is_null(), (uint64_t)ev.at("spec_version").as_int()};
    } else if (op == "delete_service") {
        S.services.erase(ev.at("id").as_str());
    } else if (op == "create_node" || op == "update_node") {
        S.createOrUpdateNode(node(ev.at("node")));
    } else if (op == "delete_node") {
        S.nodeSet.erase(ev.at("id").as_str());
    } else if (op == "create_task") {
        S.createTask(task(ev.at("task")));
    } else if (op == "update_task") {
        S.updateTask(task(ev.at("task")));
    } else if (op == "delete_task") {
        auto it = S.allTasks.find(ev.at("id").as_str());
        api::Task t; t.ID = ev.at("id").as_str();
        if (ev.find("task")) t = *task(ev.at("task"));
        else if (it != S.allTasks.end()) t = *it->second;
        S.deleteTask(t);
    } else if (op == "tick") {
        out.set("decisions", decisions_json(S.tick(strset(ev.at("fail_commit")))));
    } else if (op == "preassigned") {
        out.set("decisions", decisions_json(S.processPreassignedTasks(strset(ev.at("fail_commit")))));
    } else if (op == "pipeline_check") {
        // filter-level probe used by the constraint / filter unit tests
        TaskP t = task(ev.at("task"));
        NodeInfo ni; ni.Node = node(ev.at("node"));
        if (ni.Node->has_description && ni.Node->Description.has_resources) ni.AvailableResources = ni.Node->Description.resources;
        S.pipeline.SetTask(*t);
        mj::Value en = mj::Value::array();
        for (auto &e : S.pipeline.checklist) en.push(mj::Value::boolean(e.enabled));
        out.set("enabled", en);
        out.set("pass", mj::Value::boolean(S.pipeline.Process(ni)));
    } else if (op == "snapshot") {
        mj::Value arr = mj::Value::array();
        for (auto &kv : S.nodeSet) {
            mj::Value n = mj::Value::object();
            n.set("id", mj::Value::string(kv.first));
            n.set("active_tasks", mj::Value::integer(kv.second.ActiveTasksCount));
            mj::Value bs = mj::Value::object();
            for (auto &s : kv.second.ActiveTasksCountByService) bs.set(s.first, mj::Value::integer(s.second));
            n.set("by_service", bs);
            n.set("available", resources_json(kv.second.AvailableResources));
            mj::Value ports = mj::Value::array();
            for (auto &p : kv.second.usedHostPorts) { mj::Value e = mj::Value::array(); e.push(mj::Value::integer(p.protocol)); e.push(mj::Value::integer(p.port)); ports.push(e); }
            n.set("ports", ports);
            mj::Value fl = mj::Value::object();
            for (auto &f : kv.second.recentFailures) fl.set(f.first.serviceID + "@" + std::to_string(f.first.specVersion), mj::Value::integer((int64_t)f.second.size()));
            n.set("failures", fl);
            mj::Value tk = mj::Value::array();
            for (auto &t : kv.second.Tasks) tk.push(mj::Value::string(t.first));
            n.set("tasks", tk);
            arr.push(n);
        }
        out.set("nodes", arr);
        mj::Value un = mj::Value::array();
        for (auto &kv : S.unassignedTasks) un.push(mj::Value::string(kv.first));
        out.set("unassigned", un);
        mj::Value pp = mj::Value::array();
        for (auto &kv : S.pendingPreassignedTasks) pp.push(mj::Value::string(kv.first));
        out.set("pending_preassigned", pp);
    } else if (op == "nodeinfo") {
        // NodeInfo unit probe (nodeinfo_test.go): newNodeInfo(node, tasks, available) then add/remove calls
        std::vector<TaskP> tasks;
        for (auto &t : ev.at("tasks").a) tasks.push_back(task(t));
        NodeInfo ni = newNodeInfo(node(ev.at("node")), tasks, resources(ev.at("available")), S.now);
        mj::Value res = mj::Value::array();
        for (auto &o : ev.at("ops").a) {
            TaskP t = task(o.at("task"));
            bool r = o.at("op").as_str() == "add" ? ni.addTask(t) : ni.removeTask(*t);
            res.push(mj::Value::boolean(r));
        }
        out.set("results", res);
        out.set("available", resources_json(ni.AvailableResources));
        out.set("active_tasks", mj::Value::integer(ni.ActiveTasksCount));
    } else if (op == "tree") {
        // nodeSet.tree probe (nodeset_test.go): nodes carry by_service counts; filter and less are constants
        Scheduler T2;
        for (auto &nv : ev.at("nodes").a) {
            NodeInfo ni; ni.Node = node(nv);
            for (auto &kv : nv.at("by_service").o) ni.ActiveTasksCountByService[kv.first] = (int)kv.second.as_int();
            T2.nodeSet[ni.Node->ID] = ni;
        }
        api::Task dummy; T2.pipeline.SetTask(dummy);
        std::vector<std::string> prefs; for (auto &p : ev.at("preferences").a) prefs.push_back(p.as_str());
        NodeLess nl = [](const NodeInfo &, const NodeInfo &) { return true; };
        DecisionTree tr = T2.tree(ev.at("service_id").as_str(), prefs, (int)ev.at("max_assignments").as_int(), nl);
        std::function<mj::Value(const DecisionTree &)> dumpt = [&](const DecisionTree &d) {
            mj::Value o = mj::Value::object();
            o.set("tasks", mj::Value::integer(d.tasks));
            o.set("nodes", mj::Value::integer((int64_t)d.heap.size()));
            mj::Value nx = mj::Value::object();
            for (auto &kv : d.next) nx.set(kv.first, dumpt(*kv.second));
            o.set("next", nx);
            return o;
        };
        out.set("tree", dumpt(tr));
    } else if (op == "parse_constraints") {
        std::vector<std::string> env; for (auto &c : ev.at("constraints").a) env.push_back(c.as_str());
        std::vector<constraint::Constraint> cs;
        bool ok = constraint::parse(env, cs);
        out.set("ok", mj::Value::boolean(ok));
        mj::Value arr = mj::Value::array();
        if (ok) for (auto &c : cs) { mj::Value e = mj::Value::object(); e.set("key", mj::Value::string(c.key)); e.set("op", mj::Value::integer(c.op)); e.set("exp", mj::Value::string(c.exp)); arr.push(e); }
        out.set("constraints", arr);
    } else if (op == "match") {
        constraint::Constraint c{ev.at("key").as_str(), (int)ev.at("operator").as_int(), ev.at("exp").as_str()};
        bool m = false;
        // Match(whats...) with several targets: `match` if any equals (constraint.go:87-96)
        bool any = false;
        for (auto &w : ev.at("whats").a) any = any || strs::equal_fold(c.exp, w.as_str());
        m = c.op == constraint::eq ? any : !any;
        out.set("match", mj::Value::boolean(m));
    } else if (op == "generic") {
        // genericresource unit probes: claim / reclaim / consume / has_enough
        std::string fn = ev.at("fn").as_str();
        api::Resources avail = resources(ev.at("available"));
        if (fn == "claim") {
            api::Resources want = resources(ev.at("reservations"));
            genericresource::List assigned;
            bool ok = genericresource::claim(avail.Generic, assigned, want.Generic);
            api::Resources a; a.Generic = assigned;
            out.set("ok", mj::Value::boolean(ok)); out.set("assigned", resources_json(a).at("generic"));
        } else if (fn == "reclaim") {
            api::Resources assigned = resources(ev.at("assigned")), noderes = resources(ev.at("node"));
            genericresource::reclaim(avail.Generic, assigned.Generic, noderes.Generic);
        } else if (fn == "reclaim_resources") {
            api::Resources assigned = resources(ev.at("assigned"));
            genericresource::reclaim_resources(avail.Generic, assigned.Generic);
        } else if (fn == "consume") {
            api::Resources res = resources(ev.at("res"));
            genericresource::consume_node_resources(avail.Generic, res.Generic);
        } else if (fn == "has_enough") {
            api::Resources want = resources(ev.at("reservations"));
            out.set("enough", mj::Value::boolean(!want.Generic.empty() && genericresource::has_enough(avail.Generic, want.Generic[0])));
        }
        out.set("available", resources_json(avail).at("generic"));
    } else {
        out.set("error", mj::Value::string("unknown op " + op));
    }
    return out;
}
}  // namespace drv

extern "C" {
void *so_create() { return new Scheduler(); }
void so_destroy(void *h) { delete reinterpret_cast<Scheduler *>(h); }
// Returns a malloc'd JSON string; free with so_free.
char *so_apply(void *h, const char *json) {
    std::string out;
    try {
        mj::Value ev = mj::parse(json);
        out = mj::dump(drv::apply(*reinterpret_cast<Scheduler *>(h), ev));
    } catch (const std::exception &e) {
        mj::Value o = mj::Value::object();
        o.set("error", mj::Value::string(e.what()));
        out = mj::dump(o);
    }
    char *r = (char *)malloc(out.size() + 1);
    std::memcpy(r, out.c_str(), out.size() + 1);
    return r;
}
void so_free(char *p) { free(p); }
}
