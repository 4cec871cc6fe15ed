/*
 * placement_engine.h -- C ABI of the B200 placement engine (libplacement.so).
 *
 * This is the drop-in boundary for the SwarmKit manager/scheduler hot path.
 * The reference has no FFI for this path (the only hook is Pipeline.AddFilter,
 * manager/scheduler/pipeline.go:70-72), so the seam is cut inside the
 * scheduler at
 *     (*Scheduler).scheduleTaskGroup   manager/scheduler/scheduler.go:694-748
 *     (*Scheduler).taskFitNode         manager/scheduler/scheduler.go:646-690
 *     nodeSet / NodeInfo maintenance   manager/scheduler/scheduler.go:254-396,
 *                                      manager/scheduler/nodeinfo.go:66-154
 * Everything that crosses it is integers: strings are dictionary-encoded by
 * the host shim (exact interning; constraint values are case-folded with
 * pe_fold_value so strings.EqualFold becomes integer equality).
 *
 * Conventions (cgo-safe): flat arrays only, the caller owns every in/out
 * buffer, the engine retains no caller pointer after a call returns, the
 * engine owns device and pinned memory, no callbacks.  Every function returns
 * an int32 status (PE_OK == 0); pe_last_error() gives the message.  The
 * reference scheduler is single-goroutine (scheduler.go:175-237), so the
 * engine is single-caller by contract: not thread-safe.
 *
 * The same ABI is implemented twice: by the CUDA engine (product,
 * swarmkit_b200/csrc) and -- with the prefix ope_ instead of pe_ -- by the CPU
 * oracle (test infrastructure, oracle/flat_oracle.cpp).
 */
#ifndef PLACEMENT_ENGINE_H
#define PLACEMENT_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PE_ABI_VERSION 2u   /* 2: pe_group.leaf_cnt (was tail padding), pe_pref_leaves */

/* ---- status codes ------------------------------------------------------- */
#define PE_OK 0
#define PE_ERR_INVALID 1     /* bad argument / index out of range              */
#define PE_ERR_CUDA 2        /* CUDA runtime failure (message has the detail)  */
#define PE_ERR_NOMEM 3
#define PE_ERR_UNSUPPORTED 4 /* feature the engine does not implement          */
#define PE_ERR_NO_DEVICE 5   /* no CUDA device: there is NO CPU fallback       */
#define PE_ERR_OVERFLOW 6    /* a counter left the range the key packing holds */

#define PE_NONE 0xFFFFFFFFu  /* "no node" / "no slot"                          */

/* ---- filter pipeline positions (pipeline.go:10-19 + scheduler.go:132) ---- */
#define PE_F_READY 0       /* ReadyFilter        filter.go:40-43   */
#define PE_F_RESOURCE 1    /* ResourceFilter     filter.go:60-93   */
#define PE_F_PLUGIN 2      /* PluginFilter       filter.go:118-208 */
#define PE_F_CONSTRAINT 3  /* ConstraintFilter   filter.go:224-243 */
#define PE_F_PLATFORM 4    /* PlatformFilter     filter.go:259-312 */
#define PE_F_HOSTPORT 5    /* HostPortFilter     filter.go:328-353 */
#define PE_F_MAXREPLICAS 6 /* MaxReplicasFilter  filter.go:369-381 */
#define PE_F_VOLUMES 7     /* VolumesFilter      filter.go:388-447 (host-side; see pe_group.flags) */
#define PE_NUM_FILTERS 8

/* ---- node row ----------------------------------------------------------- */
/* pe_node_row.flags */
#define PE_NODE_VALID 0x01u        /* row holds a node (nodeSet membership)                      */
#define PE_NODE_READY 0x02u        /* Status.State==READY && Spec.Availability==ACTIVE           */
#define PE_NODE_HAS_PLATFORM 0x04u /* Description != nil && Description.Platform != nil          */
#define PE_NODE_HAS_ENGINE 0x08u   /* Description != nil && Description.Engine != nil            */
#define PE_NODE_HAS_LOGPLUGIN 0x10u/* engine reports at least one plugin of type "Log"           */
#define PE_NODE_IP_VALID 0x20u     /* net.ParseIP(Status.Addr) != nil                            */
#define PE_NODE_IP_V4 0x40u        /* that address has a 4-byte form (ip.To4() != nil)           */

/* Generic-resource cell encoding (api/genericresource/validate.go:24-51):
 * value = (count << 2) | type.  Discrete: count = DiscreteResourceSpec.Value;
 * named: count = number of entries of that kind.  Absent kind = 0. */
#define PE_GEN_ABSENT 0
#define PE_GEN_DISCRETE 1
#define PE_GEN_NAMED 2
#define PE_GEN_ENCODE(count, type) ((int64_t)(((uint64_t)(int64_t)(count) << 2) | (uint64_t)(type)))

typedef struct pe_kv32 { uint32_t key; uint32_t value; } pe_kv32;
typedef struct pe_kv64 { uint32_t key; uint32_t pad; int64_t value; } pe_kv64;

/* One node row = the device mirror of scheduler.NodeInfo (nodeinfo.go:28-44).
 * node_idx IS the canonical tie-break position (rank of the node ID
 * in ascending byte order); the host shim keeps rows in that order. */
typedef struct pe_node_row {
    uint32_t node_idx;
    uint32_t flags;        /* PE_NODE_*                                                      */
    uint32_t os_id;        /* exact-string id of Platform.OS (1..255; 0 = "")  filter.go:308 */
    uint32_t arch_id;      /* exact-string id of Platform.Architecture AFTER the x86_64->amd64,
                              aarch64->arm64 normalisation of filter.go:291-306 (1..255)     */
    int64_t cpu_avail;     /* AvailableResources.NanoCPUs  (may be negative, scheduler.go:378) */
    int64_t mem_avail;     /* AvailableResources.MemoryBytes                                  */
    uint32_t total_tasks;  /* ActiveTasksCount  nodeinfo.go:31                                */
    uint32_t attr_off, attr_cnt; /* -> pe_kv32 {attr column, folded value id}; others = 0 ("") */
    uint32_t gen_off, gen_cnt;   /* -> pe_kv64 {kind id, PE_GEN_ENCODE(count,type)}           */
    uint32_t svc_off, svc_cnt;   /* -> pe_kv32 {service id, ActiveTasksCountByService}        */
    uint32_t port_off, port_cnt; /* -> uint32 host-port slot ids in usedHostPorts             */
    uint32_t plug_off, plug_cnt; /* -> uint32 plugin slot ids the node satisfies
                                    (exact name or name+":latest", filter.go:186-205)         */
    uint32_t ip[4];        /* 16-byte address, big-endian words (v4 as ::ffff:a.b.c.d)        */
} pe_node_row;

/* Fixed attribute columns; label columns are allocated by the shim from
 * PE_ATTR_FIRST_LABEL upwards (one per referenced node.labels.K / engine.labels.K). */
#define PE_ATTR_NODE_ID 0   /* constraint.go:110 */
#define PE_ATTR_HOSTNAME 1  /* constraint.go:114 */
#define PE_ATTR_ROLE 2      /* constraint.go:147 */
#define PE_ATTR_OS 3        /* constraint.go:151 (NOT normalised) */
#define PE_ATTR_ARCH 4      /* constraint.go:161 (NOT normalised) */
#define PE_ATTR_FIRST_LABEL 5

/* ---- task-group descriptor ---------------------------------------------- */
/* pe_constraint: one `key op value` on an attribute column, constraint.go:84-104 */
typedef struct pe_constraint {
    uint32_t col;    /* attribute column                                  */
    uint32_t value;  /* folded value id (0 = "")                          */
    uint32_t neq;    /* 0: ==   1: !=                                     */
} pe_constraint;

/* node.ip constraint, constraint.go:127-146. Single address: mask all ones. */
typedef struct pe_ip_constraint {
    uint32_t net[4];
    uint32_t mask[4];
    uint32_t neq;
    uint32_t is_cidr; /* 1: subnet.Contains semantics (family must match); 0: ip.Equal */
    uint32_t is_v4;   /* family of the constraint network (CIDR only)                   */
} pe_ip_constraint;

typedef struct pe_platform { uint32_t os_id; uint32_t arch_id; } pe_platform; /* 0 = wildcard */
typedef struct pe_generic_want { uint32_t kind; uint32_t pad; int64_t value; } pe_generic_want;
typedef struct pe_node_fail { uint32_t node_idx; uint32_t count; } pe_node_fail; /* sorted by node_idx */

/* pe_group.flags */
#define PE_G_CONSTRAINT_NEVER 0x1u /* a constraint has an unknown key or a malformed node.ip
                                      operand: every node is rejected (constraint.go:144,200) */
#define PE_G_LOG_DRIVER 0x2u       /* log_plugin is set                                        */

/* per-task flags (task_flags[]) */
#define PE_T_COUNTS 0x1u /* DesiredState <= COMPLETED: bumps the spread counters (nodeinfo.go:148) */

typedef struct pe_group {
    uint32_t svc_id;      /* service row of the per-service counter table                  */
    uint32_t n_tasks;     /* k = len(taskGroup)  scheduler.go:737                          */
    uint32_t task_off;    /* tasks of a group are contiguous: [task_off, task_off+n_tasks) */
    uint32_t filter_mask; /* bit f set = filter f enabled by its SetTask (pipeline.go:76)  */
    int64_t cpu_res;      /* Reservations.NanoCPUs      nodeinfo.go:156-161                */
    int64_t mem_res;      /* Reservations.MemoryBytes                                      */
    uint64_t max_replicas;/* Placement.MaxReplicas      filter.go:380                      */
    uint32_t gen_off, gen_cnt;   /* -> pe_generic_want                                     */
    uint32_t con_off, con_cnt;   /* -> pe_constraint                                       */
    uint32_t ip_off, ip_cnt;     /* -> pe_ip_constraint                                    */
    uint32_t plat_off, plat_cnt; /* -> pe_platform (OR)   filter.go:272-289                */
    uint32_t port_off, port_cnt; /* -> uint32 host-port slots (checked AND reserved)       */
    uint32_t plug_off, plug_cnt; /* -> uint32 required plugin slots                        */
    uint32_t log_plugin;         /* plugin slot of Spec.LogDriver.Name, or PE_NONE         */
    uint32_t fail_off, fail_cnt; /* -> pe_node_fail: countRecentFailures(now) per node     */
    uint32_t tie_start;   /* tie-break / evaluation order is node_idx rotated by this
                             amount: pos(n) = (n - tie_start) mod N.  0 = SURVEY canonical */
    uint32_t flags;       /* PE_G_*                                                        */
    uint32_t leaf_cnt;    /* placement preferences (nodeSet.tree nodeset.go:59-101,
                             scheduleNTasksOnSubtree scheduler.go:772-825): this group is ONE
                             VISIT OF ONE LEAF of the service's decision tree.  The leaf_cnt
                             pe_constraint entries that FOLLOW the group's constraints
                             (cons[con_off + con_cnt ...], neq ignored) name the leaf: a node
                             belongs to it iff every (column == value) holds.  Nodes outside
                             the leaf are not part of the node set for this group (they are
                             neither evaluated nor counted).  0 = no preferences            */
} pe_group;

/* One scheduling pass = the group loop of Scheduler.tick, scheduler.go:464-469.
 * Groups are processed in array order; the shim sorts them canonically. */
typedef struct pe_tick {
    const pe_group *groups;          uint32_t n_groups;
    const uint8_t *task_flags;       uint32_t n_tasks;
    const pe_generic_want *gens;     uint32_t n_gens;
    const pe_constraint *cons;       uint32_t n_cons;
    const pe_ip_constraint *ips;     uint32_t n_ips;
    const pe_platform *plats;        uint32_t n_plats;
    const uint32_t *ports;           uint32_t n_ports;
    const uint32_t *plugs;           uint32_t n_plugs;
    const pe_node_fail *fails;       uint32_t n_fails;
} pe_tick;

/* ---- engine configuration ----------------------------------------------- */
#define PE_CFG_NO_SPECULATION 0x1u /* disable the batched k=1 scan (debug / A-B)  */
#define PE_CFG_ORDERED_ONLY 0x2u   /* place scanned batches with the ordered sequencer alone (A-B against the
                                      chunked parallel placement step)             */

typedef struct pe_config {
    uint32_t abi_version;   /* PE_ABI_VERSION                                  */
    int32_t device;         /* CUDA ordinal; -1 = current device               */
    uint32_t node_capacity; /* rows to pre-allocate (grows on demand)          */
    uint32_t flags;         /* PE_CFG_*                                        */
    uint32_t max_batch;     /* k=1 groups per scan launch; 0 = auto            */
    /* node sharding across ranks; world_size 1 = single GPU */
    int32_t rank;
    int32_t world_size;
    const void *nccl_unique_id; /* 128-byte ncclUniqueId when world_size > 1   */
} pe_config;

typedef struct pe_engine pe_engine;

/* counters since pe_create / pe_stats_reset */
typedef struct pe_stats {
    uint64_t evals;            /* (group,node) filter evaluations done by the scan kernel     */
    uint64_t evals_generic;    /* ... by the sequencer's full-table path                      */
    uint64_t scan_bytes;       /* algorithmic bytes those scan evaluations read   */
    uint64_t placements;       /* tasks given a node                                          */
    uint64_t fast_path;        /* k=1 groups resolved from the scan's best-class bitmap       */
    uint64_t medium_path;      /* ... from the second class + touched members of the first    */
    uint64_t slow_path;        /* groups that took the sequencer's full-table path            */
    uint64_t kernel_launches;  /* engine kernels launched                                     */
    uint64_t scan_launches;
    double scan_ms;            /* CUDA-event time inside the scan kernel                      */
    double sequencer_ms;       /* CUDA-event time inside the sequencer kernel                 */
    double h2d_ms, d2h_ms;
    double run_ms;             /* CUDA-event time from the first to the last kernel of pe_tick_run */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t seq_cycles_fast, seq_cycles_medium, seq_cycles_generic; /* SM cycles of the sequencer per mode */
    uint64_t pairs;            /* (task,node) pairs the batched scan covered (tasks x nodes); `evals` counts the
                                  evaluations it executed: tasks with identical descriptors share one scan row */
    uint64_t scan_rows;        /* distinct descriptors scanned (rows), summed over batches                       */
    uint64_t static_evals;     /* (signature,node) evaluations of the attribute filters (k_static)               */
    double prep_ms;            /* CUDA-event time in classify / static / rows kernels                            */
    /* sequencer fast-mode diagnostics: exits by reason (end, none, window, neutral, class consumed); the rest is
     * filled only by builds with -DPE_SEQ_PROFILE=1: cycles the ordered warp waited for / worked on staged tasks,
     * tasks whose staged candidate had been taken                                                                */
    uint64_t seq_stops[5];
    uint64_t seq_cons_wait, seq_cons_work, seq_rewalks;
    uint64_t seq_prof[16];      /* see SeqDebug in kernel_sequencer.cuh */
    /* the chunked parallel placement step (kernel_place.cuh): CUDA-event time, tasks it placed or examined, batches it
     * handed to the ordered sequencer part-way, lanes re-ranked from the chunk log, candidate tails, chunks, and SM
     * cycles of CTA 0 in the stage / resolve / commit phases */
    double place_ms;
    uint64_t place_tasks, place_cuts, place_amb, place_tails, place_chunks;
    uint64_t place_cyc[3];
} pe_stats;

/* ---- lifecycle ----------------------------------------------------------- */
int32_t pe_create(const pe_config *cfg, pe_engine **out);
/* Node sharding: rank 0 obtains a 128-byte ncclUniqueId here, the host ships it to the other
 * ranks (the Go shim would use its own RPC; bench.py uses torch.distributed), and every rank passes it in
 * pe_config.nccl_unique_id.  All ranks mirror ALL nodes and submit the SAME ticks; each rank scans its slice
 * of the node axis, partial results are all-gathered, placements come out identical on every rank. */
int32_t pe_nccl_unique_id(void *out128);
void pe_destroy(pe_engine *h);
const char *pe_last_error(const pe_engine *h); /* h may be NULL: last create error */
uint32_t pe_abi_version(void);

/* ---- node mirror (createOrUpdateNode scheduler.go:368-396, nodeSet.remove nodeset.go:46) -- */
int32_t pe_node_upsert(pe_engine *h, const pe_node_row *rows, uint32_t n_rows,
                       const pe_kv32 *attrs, const pe_kv64 *gens, const pe_kv32 *svcs,
                       const uint32_t *ports, const uint32_t *plugs);
int32_t pe_node_remove(pe_engine *h, const uint32_t *node_idx, uint32_t n);
/* Set the number of rows the scans cover (rows >= n_nodes are ignored). */
int32_t pe_set_node_count(pe_engine *h, uint32_t n_nodes);

/* addTask / removeTask outside a tick and commit rollback (nodeinfo.go:66-154,
 * scheduler.go:472-487).  sign = +1: add, -1: remove.  The deltas are the
 * amounts the shim computed on its own NodeInfo (it owns the named-resource
 * member lists, SURVEY hard part D). */
typedef struct pe_task_delta {
    uint32_t node_idx;
    uint32_t svc_id;
    int32_t sign;
    uint32_t counts;       /* 1: also move ActiveTasksCount / ByService (DesiredState<=COMPLETED) */
    int64_t cpu, mem;      /* reservation amounts                                                  */
    uint32_t gen_off, gen_cnt;   /* -> pe_kv64 {kind, NEW encoded cell value after the change}     */
    uint32_t port_off, port_cnt; /* -> uint32 slots set (add) or cleared (remove)                  */
} pe_task_delta;
int32_t pe_node_task_delta(pe_engine *h, const pe_task_delta *d, uint32_t n,
                           const pe_kv64 *gens, const uint32_t *ports);

/* ---- the hot path ------------------------------------------------------- */
/* Mirrors the group loop of tick + scheduleTaskGroup, applying reservations as
 * it goes.  out_node[n_tasks]: chosen node_idx or PE_NONE.  out_fail[n_groups*8]:
 * Pipeline failure counters (pipeline.go:56-68,84-103) as they stand when the
 * group finishes -- meaningful when the group has unplaced tasks, and WRITTEN
 * only when some task of the tick went unplaced (the reference reads Explain()
 * only then, scheduler.go:744-746).  Host buffers in, host buffers out. */
int32_t pe_schedule(pe_engine *h, const pe_tick *tick, uint32_t *out_node, uint32_t *out_fail);

/* Same pass split so the copy and the compute can be timed apart. */
int32_t pe_tick_upload(pe_engine *h, const pe_tick *tick);
int32_t pe_tick_run(pe_engine *h);   /* enqueue + wait */
int32_t pe_tick_download(pe_engine *h, uint32_t *out_node, uint32_t *out_fail);

/* taskFitNode (scheduler.go:646-690): each request is a k=1 group evaluated on
 * one named node; on success the reservation is applied.  out_ok[i] = 1 fits,
 * 0 does not fit, 2 node not in the set (scheduler.go:648-651).
 * out_fail[i*8..] the failure counters when it does not fit.  The groups
 * reference tick arrays exactly like pe_schedule. */
int32_t pe_fit(pe_engine *h, const pe_tick *tick, const uint32_t *node_idx,
               uint8_t *out_ok, uint32_t *out_fail);

/* ---- placement preferences: the tree's bookkeeping ------------------------ */
/* nodeSet.tree (nodeset.go:59-101) hangs every node of the set under the branch named by its values of the
 * preference labels and adds the node's ActiveTasksCountByService[service] to every branch on the way.  This call
 * returns the LEAVES: every distinct tuple of values of the n_levels attribute columns `cols` among the rows of the
 * node set (feasible or not), with the sum of that count over the leaf's nodes.  out_vals[cap * n_levels] (tuple of
 * leaf i at out_vals[i * n_levels ...]), out_tasks[cap]; order unspecified (the shim sorts branches by label value).
 * PE_ERR_OVERFLOW if more than cap leaves exist; n_levels <= PE_MAX_PREF_LEVELS.  The host walks the tree
 * (scheduler.go:784-822) and submits one pe_group with leaf_cnt > 0 per leaf visit. */
#define PE_MAX_PREF_LEVELS 8
int32_t pe_pref_leaves(pe_engine *h, uint32_t svc_id, const uint32_t *cols, uint32_t n_levels, uint32_t cap,
                       uint32_t *out_vals, uint32_t *out_tasks, uint32_t *out_n_leaves);

/* ---- the predicate matrix in the other direction ------------ */
/* constraint.NodeMatches per (service, node) is also what the constraint enforcer
 * (manager/orchestrator/constraintenforcer/constraint_enforcer.go:65-226) and the global orchestrator's node
 * eligibility (manager/orchestrator/global/global.go:306,440,513) evaluate, one pair at a time.  This call
 * evaluates the node-attribute filters a group ENABLES in filter_mask -- Ready, Plugin, Constraint, Platform; for
 * NodeMatches alone set only 1 << PE_F_CONSTRAINT -- for every group of the tick against every row of the node
 * set, with the kernel the scheduler's own batches use (one bit per pair): out_bits[n_groups * words],
 * words = (node count + 31) / 32, bit n of row g = node n passes group g.  No state is read or changed. */
int32_t pe_match_matrix(pe_engine *h, const pe_tick *tick, uint32_t *out_bits);

/* ---- introspection (parity checker) -------------------------------------- */
typedef struct pe_node_state {
    uint32_t flags;
    uint32_t total_tasks;
    int64_t cpu_avail, mem_avail;
} pe_node_state;
int32_t pe_snapshot(pe_engine *h, uint32_t first, uint32_t n, pe_node_state *out);
int32_t pe_snapshot_service(pe_engine *h, uint32_t svc_id, uint32_t first, uint32_t n, uint32_t *out);
int32_t pe_snapshot_generic(pe_engine *h, uint32_t kind, uint32_t first, uint32_t n, int64_t *out);
int32_t pe_snapshot_ports(pe_engine *h, uint32_t slot, uint32_t first, uint32_t n, uint8_t *out);

int32_t pe_get_stats(pe_engine *h, pe_stats *out);
int32_t pe_stats_reset(pe_engine *h);

/* strings.EqualFold folding for constraint operands (SURVEY hard part C):
 * A-Z -> a-z, U+212A (E2 84 AA) -> k, U+017F (C5 BF) -> s, every other byte
 * unchanged.  Two strings fold to the same bytes iff strings.EqualFold holds
 * whenever one of them satisfies constraint.go:26's valuePattern.  Returns the
 * folded length, or -1 if out_cap is too small. */
int32_t pe_fold_value(const char *in, uint32_t len, char *out, uint32_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* PLACEMENT_ENGINE_H */
