// dev_types.h -- device-side layout of the node mirror and of one tick.
//
// HBM layout (DESIGN.md "Data layout"): the node mirror is struct-of-arrays,
// one contiguous column per field, every column padded to a multiple of
// PE_ROW_PAD rows so that any node tile is a 16-byte aligned contiguous slice
// of each column (what cp.async.bulk needs).  Columns trace to NodeInfo fields
// (manager/scheduler/nodeinfo.go:28-44) exactly as SURVEY.md Appendix B lists.
#pragma once
#include <stdint.h>
#include "../../include/placement_engine.h"

#define PE_ROW_PAD 2048u          // row padding = largest scan tile
#define PE_META_FLAGS_MASK 0xFFu  // meta = flags | os_id << 8 | arch_id << 16
#define PE_PREF_NONE 0xFFFFFFFFFFFFFFFFull

// Status bits the kernels raise in DevStatus::error.
#define PE_DEV_ERR_SVC_OVERFLOW 0x1u  // a per-service count reached 2^24
#define PE_DEV_ERR_WD_CONSUMER 0x10u  // watchdog: sequencer consumer waited > 1 s for a ring slot
#define PE_DEV_ERR_WD_DRAIN 0x20u     // watchdog: drain of an armed ring slot
#define PE_DEV_ERR_WD_SCAN 0x40u      // watchdog: scan kernel waited > 1 s for a node tile

struct DevTable {
    uint32_t n_nodes;   // rows covered by scans (N)
    uint32_t n_attr, n_gen, n_svc, n_portw, n_plugw;
    uint32_t *meta;     // flags | os << 8 | arch << 16
    int64_t *cpu;       // AvailableResources.NanoCPUs
    int64_t *mem;       // AvailableResources.MemoryBytes
    uint32_t *total;    // ActiveTasksCount
    uint4 *ip;          // 16-byte address
    uint32_t **attr;    // [n_attr]  folded value ids per attribute column
    int64_t **gen;      // [n_gen]   PE_GEN_ENCODE cells
    uint32_t **svc;     // [n_svc]   ActiveTasksCountByService, dense per service
    uint32_t **ports;   // [n_portw] usedHostPorts bit words
    uint32_t **plug;    // [n_plugw] plugin-present bit words
};

struct TickDev {
    const pe_group *groups;
    const uint8_t *task_flags;
    const pe_generic_want *gens;
    const pe_constraint *cons;
    const pe_ip_constraint *ips;
    const pe_platform *plats;
    const uint32_t *ports;
    const uint32_t *plugs;
    const pe_node_fail *fails;
    uint32_t *out_node;  // [n_tasks]
    uint32_t *out_fail;  // [n_groups * 8]
};

// Result of the batched k=1 scan for one ROW (= one distinct task descriptor of
// the batch; tasks with identical descriptors share a row, kernel_classify.cuh).
// The scan keeps the two smallest rank classes of every row: bitmap rows
// 2*row (best class) and 2*row + 1 (second class), member lists likewise.
struct ScanResult {
    unsigned long long c0;  // smallest (f5, svc, total) prefix among feasible nodes; PE_PREF_NONE if none
    unsigned long long c1;  // second smallest prefix; PE_PREF_NONE if there is no second class
    uint32_t n0, n1;        // members of the best / second class (the first PE_LIST_CAP of each are listed)
    // what the sequencer's fast path needs to commit, so that it reads ONE record per row
    uint32_t tie_start;
    uint32_t flags;         // PE_SR_*
    long long cpu_res, mem_res;
    uint32_t *svccol;       // the row's per-service counter column
    unsigned long long max_replicas;   // Placement.MaxReplicas (meaningful with PE_SR_MAXREP)
};
#define PE_SR_SIMPLE 1u     // no generic resources / host ports: the reservation is four reductions
#define PE_SR_COUNTS 2u     // DesiredState <= COMPLETED: bumps the spread counters
#define PE_SR_K1 4u         // the group really has exactly one task
#define PE_SR_RES 16u       // ResourceFilter enabled (cpu / memory reservations)
#define PE_SR_MAXREP 32u    // MaxReplicasFilter enabled
#define PE_SR_INLINE 8u     // cpu / memory / max-replicas at most (no generic resources, host ports, recent-failure counts):
                            // the ordered warp can re-rank a consumed best class itself (inline_medium)

struct DevCounters {
    unsigned long long fast_path, medium_path, slow_path, placements, evals_generic;
    unsigned long long cyc_fast, cyc_medium, cyc_generic;   // SM cycles the sequencer spent in each mode
    unsigned long long cyc_cons_wait, cyc_cons_work, stops[5], iters;   // consumer warp: waiting on producers / working; fast-mode exits by reason
    unsigned long long scan_evals, scan_bytes;     // (row,node) evaluations the scan kernel executed / their algorithmic bytes
    unsigned long long prof[16];                    // sequencer diagnostics (PE_SEQ_PROFILE builds)
    unsigned long long static_evals, scan_rows;    // (signature,node) evaluations of k_static; rows scanned
    // the chunked parallel placement step (kernel_place.cuh): tasks it handled, batches it had to hand to the ordered
    // sequencer part-way, lanes that re-ranked from the chunk log, candidate tails built, chunks, SM cycles per phase
    unsigned long long place_tasks, place_cuts, place_amb, place_tails, place_chunks, place_cyc[3];
    uint32_t error;
    uint32_t pad;
};
