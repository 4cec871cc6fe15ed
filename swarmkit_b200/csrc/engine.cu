// engine.cu -- host side of the B200 placement engine and its C ABI
// (include/placement_engine.h).  Owns the device mirror of the scheduler's
// nodeSet (manager/scheduler/nodeset.go:13-48, nodeinfo.go:28-44) and drives
// the kernels that replace the body of Scheduler.tick's group loop
// (manager/scheduler/scheduler.go:464-469) and taskFitNode (:646-690).
//
// There is no CPU fallback: without a CUDA device pe_create fails with
// PE_ERR_NO_DEVICE and nothing else can be called.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "kernel_misc.cuh"
#include "nccl_dl.h"
#include "kernel_classify.cuh"
#include "kernel_scan.cuh"
#include "kernel_sequencer.cuh"
#include "kernel_place.cuh"
#include "kernel_groups.cuh"

using namespace pe;

namespace {

std::string g_create_err;

#define CU(expr)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            char b_[512];                                                                      \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            this->err = b_;                                                                    \
            return PE_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }

// PE_DEBUG_SYNC=1 in the environment: wait after every kernel of the tick and say which one it was on stderr (finds the
// kernel that hangs or faults; never set in measurements)
bool debug_sync() { static const bool on = getenv("PE_DEBUG_SYNC") != nullptr; return on; }
#define DBG_SYNC(what, a, b)                                                                         \
    do {                                                                                             \
        if (debug_sync()) {                                                                          \
            fprintf(stderr, "[pe] %s %u %u ...", what, (unsigned)(a), (unsigned)(b));                \
            cudaError_t e_ = cudaStreamSynchronize(stream);                                          \
            fprintf(stderr, " %s\n", cudaGetErrorString(e_));                                        \
        }                                                                                            \
    } while (0)

struct EvPair { cudaEvent_t a, b; int kind; };

}  // namespace

struct pe_engine {
    std::string err;
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    uint32_t cfg_flags = 0, max_batch = 0;

    // ---- node mirror
    uint32_t cap = 0, n_nodes = 0;
    uint32_t *meta = nullptr;
    int64_t *cpu = nullptr, *mem = nullptr;
    uint32_t *total = nullptr;
    uint4 *ip = nullptr;
    std::vector<uint32_t *> attr, svc, ports, plug;
    std::vector<int64_t *> gen;
    void **d_tab[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // attr, gen, svc, ports, plug
    size_t tab_cap[5] = {0, 0, 0, 0, 0};
    bool tabs_dirty = true;

    // ---- staged tick
    void *tick_buf = nullptr; size_t tick_cap = 0;
    TickDev K{};
    struct Run { uint32_t begin, end; bool one; bool seq_only; };   // maximal runs of k == 1 / k != 1 groups (seq_only: one-task leaf visits)
    std::vector<Run> runs;
    // what the k == 1 groups of the staged tick use (decides the scan's tile columns and variant)
    bool use_res = false, use_dyn = false, scan_ok = true;
    std::vector<uint8_t> use_gen, use_pw;
    uint32_t n_groups = 0, n_tasks = 0;
    uint32_t *d_out_node = nullptr, *d_out_fail = nullptr; size_t out_node_cap = 0, out_fail_cap = 0;

    // ---- scratch
    uint32_t scratch_cap = 0, st_cap = 0;
    uint8_t *ff8 = nullptr; unsigned long long *pref64 = nullptr;
    CandKey *cand_g = nullptr;
    int64_t *st_cpu_g = nullptr, *st_mem_g = nullptr, *st_gen_g = nullptr;
    uint32_t *st_svc_g = nullptr, *st_tot_g = nullptr, *st_placed_g = nullptr;
    uint8_t *st_flags_g = nullptr;
    uint32_t *touched_g = nullptr;
    uint32_t *E = nullptr; size_t E_words = 0;
    uint32_t *Lbuf = nullptr; size_t L_words = 0;   // class member lists
    ScanResult *scan_out = nullptr; uint32_t scan_out_cap = 0;
    // descriptor classes / signature bitmaps / rows (kernel_classify.cuh)
    void *cls_buf = nullptr; size_t cls_cap = 0;     // per-run arrays
    void *rows_buf = nullptr; size_t rows_cap = 0;   // per-batch arrays
    uint32_t *Sbuf = nullptr; size_t S_words = 0;    // signature bitmaps
    void *chunk_buf = nullptr; size_t chunk_cap = 0; // per-chunk partial results of the scan (p1 | Cc | Lc)
    uint32_t n_chunks = 1, local_chunks = 1;         // node-axis chunks of the scan (all ranks) / owned by this rank
    // node sharding across ranks: every rank mirrors all nodes and runs the same sequencer; the scan of a
    // batch is split by node range and the partial results are exchanged with NCCL all-gathers
    int rank = 0, world = 1;
    pe_nccl::ncclComm_t comm = nullptr;
    uint32_t *Eall = nullptr; size_t Eall_words = 0; // node sharding: merged member lists [rows][2][x_cap] (k_xpack + all-reduce)
    uint32_t *d_cls_counters = nullptr;              // [0] signatures [1] classes [2] rows of the current batch [3] where the
                                                     //     parallel placement step stopped in the current batch
    uint32_t *cursors = nullptr; size_t cursors_cap = 0;   // [rows][2] class-list cursors of the placement step
    uint32_t place_cluster = PE_PL_CLUSTER;                // CTAs per cluster of k_place (8 or 16)
    uint32_t batch_shift = 0;                              // batches of this tick are 2^-batch_shift of the usual size (hand-over feedback)
    void *groups_buf = nullptr;                            // scratch of k_groups: class table | GroupSel | per-CTA class counts
    int groups_grid = 0;                                   // CTAs of the cooperative launch (0: not available)
    DevCounters *d_ctr = nullptr;
    void *up_buf = nullptr; size_t up_cap = 0;  // upload arena for upsert / delta / fit

    pe_stats stats{};
    std::deque<EvPair> ev_pool; size_t ev_used = 0;  // deque: ev_begin hands out stable pointers
    bool timing = true;

    // =====================================================================
    int32_t init(const pe_config *cfg) {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) {
            err = std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count == 0") +
                  "); the placement engine has no CPU fallback";
            return PE_ERR_NO_DEVICE;
        }
        if (cfg->device >= 0) { device = cfg->device; CU(cudaSetDevice(device)); }
        else CU(cudaGetDevice(&device));
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        num_sms = prop.multiProcessorCount;
        CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        cfg_flags = cfg->flags;
        max_batch = cfg->max_batch;
        if (cfg->world_size > 1) {
            if (cfg->rank < 0 || cfg->rank >= cfg->world_size || !cfg->nccl_unique_id) { err = "bad rank / world_size / nccl_unique_id"; return PE_ERR_INVALID; }
            if (cfg->world_size > 32) { err = "world_size > 32"; return PE_ERR_UNSUPPORTED; }
            const pe_nccl::Api &nc = pe_nccl::api();
            if (!nc.ok) { err = "world_size > 1 needs libnccl.so.2 in the process or on the loader path"; return PE_ERR_UNSUPPORTED; }
            pe_nccl::ncclUniqueId id;
            memcpy(&id, cfg->nccl_unique_id, sizeof id);
            const pe_nccl::ncclResult_t r = nc.CommInitRank(&comm, cfg->world_size, id, cfg->rank);
            if (r != 0) { err = std::string("ncclCommInitRank: ") + nc.GetErrorString(r); return PE_ERR_CUDA; }
            rank = cfg->rank; world = cfg->world_size;
        }
        CU(cudaMalloc(&d_ctr, sizeof(DevCounters)));
        CU(cudaMalloc(&d_cls_counters, 16));
        CU(cudaMemsetAsync(d_ctr, 0, sizeof(DevCounters), stream));
        CU(cudaFuncSetAttribute(k_sequencer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_dyn_smem_bytes(12288)));
        // task groups (k > 1) on every SM: one persistent cooperative kernel, one CTA per SM (kernel_groups.cuh)
        {
            int coop = 0, per_sm = 0;
            CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
            CU(cudaFuncSetAttribute(k_groups, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)groups_smem_bytes()));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_groups, PE_GR_THREADS, groups_smem_bytes()));
            groups_grid = (coop && per_sm > 0) ? num_sms : 0;
            if (groups_grid) {
                const size_t bytes = (size_t)PE_GR_TABLE * 12 + sizeof(GroupSel) + (size_t)groups_grid * PE_GR_MAXCLS * 4 + 64;
                CU(cudaMalloc(&groups_buf, bytes));
                CU(cudaMemsetAsync(groups_buf, 0, bytes, stream));
                CU(cudaMemsetAsync(groups_buf, 0xFF, (size_t)PE_GR_TABLE * 8, stream));      // class keys: all ones = empty
            }
        }
        CU(cudaFuncSetAttribute(k_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem_bytes(PE_PL_TK_MAX_WORDS)));
        // the placement step's cluster: 16 CTAs (a chunk of 256 tasks) where the device co-schedules that many, else the
        // portable 8.  PE_PLACE_CLUSTER=8|16 overrides (A-B measurements).
        {
            place_cluster = PE_PL_CLUSTER;
            uint32_t want = PE_PL_CLUSTER_MAX;
            if (const char *ev = getenv("PE_PLACE_CLUSTER")) want = (uint32_t)atoi(ev) == 8u ? 8u : 16u;
            if (want > PE_PL_CLUSTER && cudaFuncSetAttribute(k_place, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t lc = {};
                lc.gridDim = dim3(want); lc.blockDim = dim3(PE_PL_THREADS); lc.dynamicSmemBytes = place_smem_bytes(PE_PL_TK_MAX_WORDS);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = want; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                int n_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, k_place, &lc) == cudaSuccess && n_clusters >= 1) place_cluster = want;
            }
            (void)cudaGetLastError();
        }
        CU(cudaFuncSetAttribute(k_scan<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CU(cudaFuncSetAttribute(k_scan<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CU(cudaFuncSetAttribute(k_scan<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        CU(cudaFuncSetAttribute(k_scan<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        int32_t rc = ensure_cap(cfg->node_capacity ? cfg->node_capacity : 1);
        if (rc) return rc;
        // the five fixed attribute columns always exist
        for (uint32_t c = 0; c < PE_ATTR_FIRST_LABEL; c++) { rc = ensure_col(attr, c, 4); if (rc) return rc; }
        return PE_OK;
    }

    void destroy() {
        if (stream) cudaStreamSynchronize(stream);
        if (comm) { pe_nccl::api().CommDestroy(comm); comm = nullptr; }
        auto fr = [](void *p) { if (p) cudaFree(p); };
        fr(Eall);
        fr(meta); fr(cpu); fr(mem); fr(total); fr(ip);
        for (auto p : attr) fr(p);
        for (auto p : svc) fr(p);
        for (auto p : ports) fr(p);
        for (auto p : plug) fr(p);
        for (auto p : gen) fr(p);
        for (auto p : d_tab) fr(p);
        fr(tick_buf); fr(d_out_node); fr(d_out_fail); fr(ff8); fr(pref64); fr(cand_g); fr(st_cpu_g); fr(st_mem_g); fr(st_gen_g);
        fr(st_svc_g); fr(st_tot_g); fr(st_placed_g); fr(st_flags_g); fr(touched_g); fr(E); fr(Lbuf); fr(scan_out); fr(d_ctr); fr(up_buf); fr(cls_buf); fr(rows_buf); fr(Sbuf); fr(d_cls_counters); fr(chunk_buf); fr(cursors); fr(groups_buf); fr(pref_buf);
        if (h_ctr) cudaFreeHost(h_ctr);
        for (auto &p : ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        if (stream) cudaStreamDestroy(stream);
    }

    // ---- columns ---------------------------------------------------------
    template <class T> int32_t regrow(T *&p, size_t elem, uint32_t old_cap, uint32_t new_cap) {
        void *q = nullptr;
        CU(cudaMalloc(&q, (size_t)new_cap * elem));
        CU(cudaMemsetAsync(q, 0, (size_t)new_cap * elem, stream));
        if (p && old_cap) CU(cudaMemcpyAsync(q, p, (size_t)old_cap * elem, cudaMemcpyDeviceToDevice, stream));
        if (p) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(p)); }
        p = reinterpret_cast<T *>(q);
        return PE_OK;
    }

    int32_t ensure_cap(uint32_t rows) {
        if (rows <= cap) return PE_OK;
        uint32_t nc = round_up(std::max(rows, cap + cap / 2), PE_ROW_PAD);
        int32_t rc;
        if ((rc = regrow(meta, 4, cap, nc))) return rc;
        if ((rc = regrow(cpu, 8, cap, nc))) return rc;
        if ((rc = regrow(mem, 8, cap, nc))) return rc;
        if ((rc = regrow(total, 4, cap, nc))) return rc;
        if ((rc = regrow(ip, 16, cap, nc))) return rc;
        for (auto &p : attr) if (p && (rc = regrow(p, 4, cap, nc))) return rc;
        for (auto &p : svc) if (p && (rc = regrow(p, 4, cap, nc))) return rc;
        for (auto &p : ports) if (p && (rc = regrow(p, 4, cap, nc))) return rc;
        for (auto &p : plug) if (p && (rc = regrow(p, 4, cap, nc))) return rc;
        for (auto &p : gen) if (p && (rc = regrow(p, 8, cap, nc))) return rc;
        cap = nc;
        tabs_dirty = true;
        return PE_OK;
    }

    template <class T> int32_t ensure_col(std::vector<T *> &tab, uint32_t idx, size_t elem) {
        if (idx >= (1u << 20)) { err = "column index too large"; return PE_ERR_INVALID; }
        if (idx >= tab.size()) { tab.resize(idx + 1, nullptr); tabs_dirty = true; }
        if (!tab[idx]) {
            void *q = nullptr;
            CU(cudaMalloc(&q, (size_t)cap * elem));
            CU(cudaMemsetAsync(q, 0, (size_t)cap * elem, stream));
            tab[idx] = reinterpret_cast<T *>(q);
            tabs_dirty = true;
        }
        return PE_OK;
    }

    template <class T> int32_t sync_tab(int which, std::vector<T *> &tab) {
        size_t n = std::max<size_t>(tab.size(), 1);
        if (n > tab_cap[which]) {
            if (d_tab[which]) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(d_tab[which])); }
            size_t nc = std::max<size_t>(64, n * 2);
            CU(cudaMalloc(&d_tab[which], nc * sizeof(void *)));
            tab_cap[which] = nc;
        }
        if (!tab.empty()) CU(cudaMemcpyAsync(d_tab[which], tab.data(), tab.size() * sizeof(void *), cudaMemcpyHostToDevice, stream));
        return PE_OK;
    }

    int32_t sync_tabs() {
        if (!tabs_dirty) return PE_OK;
        int32_t rc;
        // pageable host vectors: the copies are staged before the call returns
        if ((rc = sync_tab(0, attr))) return rc;
        if ((rc = sync_tab(1, gen))) return rc;
        if ((rc = sync_tab(2, svc))) return rc;
        if ((rc = sync_tab(3, ports))) return rc;
        if ((rc = sync_tab(4, plug))) return rc;
        CU(cudaStreamSynchronize(stream));
        tabs_dirty = false;
        return PE_OK;
    }

    DevTable table() const {
        DevTable T;
        T.n_nodes = n_nodes;
        T.n_attr = (uint32_t)attr.size(); T.n_gen = (uint32_t)gen.size(); T.n_svc = (uint32_t)svc.size();
        T.n_portw = (uint32_t)ports.size(); T.n_plugw = (uint32_t)plug.size();
        T.meta = meta; T.cpu = cpu; T.mem = mem; T.total = total; T.ip = ip;
        T.attr = reinterpret_cast<uint32_t **>(d_tab[0]);
        T.gen = reinterpret_cast<int64_t **>(d_tab[1]);
        T.svc = reinterpret_cast<uint32_t **>(d_tab[2]);
        T.ports = reinterpret_cast<uint32_t **>(d_tab[3]);
        T.plug = reinterpret_cast<uint32_t **>(d_tab[4]);
        return T;
    }

    int32_t ensure_buf(void *&p, size_t &capb, size_t need) {
        if (need <= capb) return PE_OK;
        if (p) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(p)); p = nullptr; }
        size_t nc = std::max(need, capb + capb / 2);
        CU(cudaMalloc(&p, nc));
        capb = nc;
        return PE_OK;
    }

    // ---- timing ------------------------------------------------------------
    EvPair *ev_begin(int kind) {
        if (!timing) return nullptr;
        if (ev_used == ev_pool.size()) {
            EvPair p; p.kind = kind;
            if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return nullptr;
            ev_pool.push_back(p);
        }
        EvPair *p = &ev_pool[ev_used++];
        p->kind = kind;
        cudaEventRecord(p->a, stream);
        return p;
    }
    void ev_end(EvPair *p) { if (p) cudaEventRecord(p->b, stream); }
    void ev_collect() {
        for (size_t i = 0; i < ev_used; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b) == cudaSuccess) {
                if (ev_pool[i].kind == 0) stats.scan_ms += ms;
                else if (ev_pool[i].kind == 1) stats.sequencer_ms += ms;
                else if (ev_pool[i].kind == 2) stats.h2d_ms += ms;
                else if (ev_pool[i].kind == 3) stats.d2h_ms += ms;
                else if (ev_pool[i].kind == 5) stats.prep_ms += ms;
                else if (ev_pool[i].kind == 6) stats.place_ms += ms;
                else stats.run_ms += ms;
            }
        }
        ev_used = 0;
    }

    // ---- node mirror ---------------------------------------------------------
    int32_t upsert(const pe_node_row *rows, uint32_t n_rows, const pe_kv32 *attrs, const pe_kv64 *gens, const pe_kv32 *svcs,
                   const uint32_t *prt, const uint32_t *plg) {
        if (!n_rows) return PE_OK;
        uint32_t maxrow = 0;
        size_t na = 0, ng = 0, ns = 0, np = 0, nq = 0;
        int32_t rc;
        for (uint32_t r = 0; r < n_rows; r++) {
            const pe_node_row &row = rows[r];
            if (row.os_id > 255 || row.arch_id > 255) { err = "os_id / arch_id must be <= 255"; return PE_ERR_INVALID; }
            maxrow = std::max(maxrow, row.node_idx);
            na = std::max<size_t>(na, (size_t)row.attr_off + row.attr_cnt);
            ng = std::max<size_t>(ng, (size_t)row.gen_off + row.gen_cnt);
            ns = std::max<size_t>(ns, (size_t)row.svc_off + row.svc_cnt);
            np = std::max<size_t>(np, (size_t)row.port_off + row.port_cnt);
            nq = std::max<size_t>(nq, (size_t)row.plug_off + row.plug_cnt);
        }
        if ((rc = ensure_cap(maxrow + 1))) return rc;
        for (uint32_t r = 0; r < n_rows; r++) {
            const pe_node_row &row = rows[r];
            for (uint32_t i = 0; i < row.attr_cnt; i++) if ((rc = ensure_col(attr, attrs[row.attr_off + i].key, 4))) return rc;
            for (uint32_t i = 0; i < row.gen_cnt; i++) if ((rc = ensure_col(gen, gens[row.gen_off + i].key, 8))) return rc;
            for (uint32_t i = 0; i < row.svc_cnt; i++) if ((rc = ensure_col(svc, svcs[row.svc_off + i].key, 4))) return rc;
            for (uint32_t i = 0; i < row.port_cnt; i++) if ((rc = ensure_col(ports, prt[row.port_off + i] >> 5, 4))) return rc;
            for (uint32_t i = 0; i < row.plug_cnt; i++) if ((rc = ensure_col(plug, plg[row.plug_off + i] >> 5, 4))) return rc;
        }
        if ((rc = sync_tabs())) return rc;
        // one arena: rows | attrs | gens | svcs | ports | plugs
        size_t o_rows = 0, o_a = o_rows + round_up((uint32_t)(n_rows * sizeof(pe_node_row)), 16), o_g = o_a + round_up((uint32_t)(na * 8), 16);
        size_t o_s = o_g + ng * 16, o_p = o_s + round_up((uint32_t)(ns * 8), 16), o_q = o_p + round_up((uint32_t)(np * 4), 16);
        size_t tot = o_q + round_up((uint32_t)(nq * 4), 16) + 16;
        if ((rc = ensure_buf(up_buf, up_cap, tot))) return rc;
        char *b = reinterpret_cast<char *>(up_buf);
        CU(cudaMemcpyAsync(b + o_rows, rows, n_rows * sizeof(pe_node_row), cudaMemcpyHostToDevice, stream));
        if (na) CU(cudaMemcpyAsync(b + o_a, attrs, na * 8, cudaMemcpyHostToDevice, stream));
        if (ng) CU(cudaMemcpyAsync(b + o_g, gens, ng * 16, cudaMemcpyHostToDevice, stream));
        if (ns) CU(cudaMemcpyAsync(b + o_s, svcs, ns * 8, cudaMemcpyHostToDevice, stream));
        if (np) CU(cudaMemcpyAsync(b + o_p, prt, np * 4, cudaMemcpyHostToDevice, stream));
        if (nq) CU(cudaMemcpyAsync(b + o_q, plg, nq * 4, cudaMemcpyHostToDevice, stream));
        stats.h2d_bytes += n_rows * sizeof(pe_node_row) + na * 8 + ng * 16 + ns * 8 + np * 4 + nq * 4;
        UpsertParams P;
        P.T = table();
        P.rows = reinterpret_cast<pe_node_row *>(b + o_rows); P.n_rows = n_rows;
        P.attrs = reinterpret_cast<pe_kv32 *>(b + o_a); P.gens = reinterpret_cast<pe_kv64 *>(b + o_g);
        P.svcs = reinterpret_cast<pe_kv32 *>(b + o_s); P.ports = reinterpret_cast<uint32_t *>(b + o_p);
        P.plugs = reinterpret_cast<uint32_t *>(b + o_q);
        // services: a row upsert replaces the node's whole counter set
        if (!svc.empty()) {
            // zero every service counter of the upserted rows first (k_upsert then sets the listed ones)
            k_zero_svc_rows<<<(n_rows + 127) / 128, 128, 0, stream>>>(P.T, P.rows, n_rows);
            stats.kernel_launches++;
        }
        k_upsert<<<(n_rows + 127) / 128, 128, 0, stream>>>(P);
        stats.kernel_launches++;
        CU(cudaGetLastError());
        if (maxrow + 1 > n_nodes) n_nodes = maxrow + 1;
        CU(cudaStreamSynchronize(stream));  // caller buffers may be reused after return
        return PE_OK;
    }

    int32_t remove(const uint32_t *idx, uint32_t n) {
        if (!n) return PE_OK;
        for (uint32_t i = 0; i < n; i++) if (idx[i] >= cap) { err = "node index out of range"; return PE_ERR_INVALID; }
        int32_t rc;
        if ((rc = ensure_buf(up_buf, up_cap, (size_t)n * 4))) return rc;
        CU(cudaMemcpyAsync(up_buf, idx, (size_t)n * 4, cudaMemcpyHostToDevice, stream));
        if ((rc = sync_tabs())) return rc;
        k_remove<<<(n + 127) / 128, 128, 0, stream>>>(table(), reinterpret_cast<uint32_t *>(up_buf), n);
        stats.kernel_launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(stream));
        return PE_OK;
    }

    int32_t delta(const pe_task_delta *d, uint32_t n, const pe_kv64 *gens, const uint32_t *prt) {
        if (!n) return PE_OK;
        size_t ng = 0, np = 0;
        int32_t rc;
        for (uint32_t i = 0; i < n; i++) {
            if (d[i].node_idx >= cap) { err = "node index out of range"; return PE_ERR_INVALID; }
            ng = std::max<size_t>(ng, (size_t)d[i].gen_off + d[i].gen_cnt);
            np = std::max<size_t>(np, (size_t)d[i].port_off + d[i].port_cnt);
            if (d[i].counts && (rc = ensure_col(svc, d[i].svc_id, 4))) return rc;
            for (uint32_t j = 0; j < d[i].gen_cnt; j++) if ((rc = ensure_col(gen, gens[d[i].gen_off + j].key, 8))) return rc;
            for (uint32_t j = 0; j < d[i].port_cnt; j++) if ((rc = ensure_col(ports, prt[d[i].port_off + j] >> 5, 4))) return rc;
        }
        if ((rc = sync_tabs())) return rc;
        size_t o_g = round_up((uint32_t)(n * sizeof(pe_task_delta)), 16), o_p = o_g + ng * 16;
        if ((rc = ensure_buf(up_buf, up_cap, o_p + np * 4 + 16))) return rc;
        char *b = reinterpret_cast<char *>(up_buf);
        CU(cudaMemcpyAsync(b, d, n * sizeof(pe_task_delta), cudaMemcpyHostToDevice, stream));
        if (ng) CU(cudaMemcpyAsync(b + o_g, gens, ng * 16, cudaMemcpyHostToDevice, stream));
        if (np) CU(cudaMemcpyAsync(b + o_p, prt, np * 4, cudaMemcpyHostToDevice, stream));
        k_delta_add<<<(n + 127) / 128, 128, 0, stream>>>(table(), reinterpret_cast<pe_task_delta *>(b), n);
        stats.kernel_launches++;
        if (ng || np) {
            k_delta_seq<<<1, 32, 0, stream>>>(table(), reinterpret_cast<pe_task_delta *>(b), n, reinterpret_cast<pe_kv64 *>(b + o_g),
                                             reinterpret_cast<uint32_t *>(b + o_p));
            stats.kernel_launches++;
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(stream));
        return PE_OK;
    }

    // ---- tick ------------------------------------------------------------------
    int32_t validate_and_prepare(const pe_tick *tk) {
        if (!tk) { err = "null tick"; return PE_ERR_INVALID; }
        int32_t rc;
        runs.clear();
        use_res = use_dyn = false; scan_ok = true;
        use_gen.assign(PE_SCAN_MAXGENK, 0); use_pw.assign(PE_SCAN_MAXW, 0);
        for (uint32_t i = 0; i < tk->n_groups; i++) {
            const pe_group &g = tk->groups[i];
            if ((uint64_t)g.task_off + g.n_tasks > tk->n_tasks) { err = "group task range out of bounds"; return PE_ERR_INVALID; }
            if ((uint64_t)g.gen_off + g.gen_cnt > tk->n_gens || (uint64_t)g.con_off + g.con_cnt + g.leaf_cnt > tk->n_cons ||
                (uint64_t)g.ip_off + g.ip_cnt > tk->n_ips || (uint64_t)g.plat_off + g.plat_cnt > tk->n_plats ||
                (uint64_t)g.port_off + g.port_cnt > tk->n_ports || (uint64_t)g.plug_off + g.plug_cnt > tk->n_plugs ||
                (uint64_t)g.fail_off + g.fail_cnt > tk->n_fails) { err = "group side-array range out of bounds"; return PE_ERR_INVALID; }
            if (n_nodes && g.tie_start >= n_nodes) { err = "tie_start >= node count"; return PE_ERR_INVALID; }
            if (g.gen_cnt > PE_MAX_GEN_WANTS) { err = "more than 8 generic reservations in one task"; return PE_ERR_UNSUPPORTED; }
            if ((g.svc_id >= svc.size() || !svc[g.svc_id]) && (rc = ensure_col(svc, g.svc_id, 4))) return rc;
            const bool one = g.n_tasks == 1 && g.leaf_cnt == 0;     // (a leaf visit of a preference tree is placed on its own)
            const bool seq_only = g.n_tasks == 1 && g.leaf_cnt != 0; // ... by the one-CTA path when it is a single task
            if (runs.empty() || runs.back().one != one || runs.back().seq_only != seq_only) runs.push_back({i, i + 1, one, seq_only});
            else runs.back().end = i + 1;
            if (one) {   // the scan path: which state-dependent columns ride in the node tiles
                if ((g.filter_mask >> PE_F_RESOURCE) & 1u) {
                    use_res = use_dyn = true;
                    for (uint32_t e = 0; e < g.gen_cnt; e++) {
                        const uint32_t kd = tk->gens[g.gen_off + e].kind;
                        if (kd >= PE_SCAN_MAXGENK) scan_ok = false; else use_gen[kd] = 1;
                    }
                }
                if ((g.filter_mask >> PE_F_HOSTPORT) & 1u) {
                    use_dyn = true;
                    for (uint32_t e = 0; e < g.port_cnt; e++) {
                        const uint32_t w = tk->ports[g.port_off + e] >> 5;
                        if (w >= PE_SCAN_MAXW) scan_ok = false; else use_pw[w] = 1;
                    }
                }
                if (((g.filter_mask >> PE_F_MAXREPLICAS) & 1u) || g.fail_cnt) use_dyn = true;
            }
            if ((g.flags & PE_G_LOG_DRIVER) && (rc = ensure_col(plug, g.log_plugin >> 5, 4))) return rc;
        }
        for (uint32_t i = 0; i < tk->n_cons; i++) if ((rc = ensure_col(attr, tk->cons[i].col, 4))) return rc;
        for (uint32_t i = 0; i < tk->n_gens; i++) if ((rc = ensure_col(gen, tk->gens[i].kind, 8))) return rc;
        for (uint32_t i = 0; i < tk->n_ports; i++) if ((rc = ensure_col(ports, tk->ports[i] >> 5, 4))) return rc;
        for (uint32_t i = 0; i < tk->n_plugs; i++) if ((rc = ensure_col(plug, tk->plugs[i] >> 5, 4))) return rc;
        for (uint32_t i = 0; i < tk->n_fails; i++) {
            if (tk->fails[i].count > 255) { err = "more than 255 recent failures on one node"; return PE_ERR_OVERFLOW; }
        }
        return sync_tabs();
    }

    // sync = false: pe_schedule keeps the caller's buffers alive itself and waits once, at the end of the whole call
    int32_t tick_upload(const pe_tick *tk, bool sync = true) {
        if (!tk) { err = "null tick"; return PE_ERR_INVALID; }
        int32_t rc;
        n_groups = tk->n_groups; n_tasks = tk->n_tasks;
        size_t off[10]; size_t o = 0;
        auto place = [&](int i, size_t bytes) { off[i] = o; o += (bytes + 15) / 16 * 16; };
        place(0, (size_t)tk->n_groups * sizeof(pe_group));
        place(1, (size_t)tk->n_tasks);
        place(2, (size_t)tk->n_gens * sizeof(pe_generic_want));
        place(3, (size_t)tk->n_cons * sizeof(pe_constraint));
        place(4, (size_t)tk->n_ips * sizeof(pe_ip_constraint));
        place(5, (size_t)tk->n_plats * sizeof(pe_platform));
        place(6, (size_t)tk->n_ports * 4);
        place(7, (size_t)tk->n_plugs * 4);
        place(8, (size_t)tk->n_fails * sizeof(pe_node_fail));
        if ((rc = ensure_buf(tick_buf, tick_cap, o + 16))) return rc;
        char *b = reinterpret_cast<char *>(tick_buf);
        EvPair *ev = ev_begin(2);
        auto up = [&](int i, const void *src, size_t bytes) -> cudaError_t {
            stats.h2d_bytes += bytes;
            return bytes ? cudaMemcpyAsync(b + off[i], src, bytes, cudaMemcpyHostToDevice, stream) : cudaSuccess;
        };
        CU(up(0, tk->groups, (size_t)tk->n_groups * sizeof(pe_group)));
        CU(up(1, tk->task_flags, (size_t)tk->n_tasks));
        CU(up(2, tk->gens, (size_t)tk->n_gens * sizeof(pe_generic_want)));
        CU(up(3, tk->cons, (size_t)tk->n_cons * sizeof(pe_constraint)));
        CU(up(4, tk->ips, (size_t)tk->n_ips * sizeof(pe_ip_constraint)));
        CU(up(5, tk->plats, (size_t)tk->n_plats * sizeof(pe_platform)));
        CU(up(6, tk->ports, (size_t)tk->n_ports * 4));
        CU(up(7, tk->plugs, (size_t)tk->n_plugs * 4));
        CU(up(8, tk->fails, (size_t)tk->n_fails * sizeof(pe_node_fail)));
        ev_end(ev);
        // the host's pass over the groups (bounds, which columns the scan needs, missing service columns) runs while the
        // copies above are in flight; nothing reads the device copy before it is through
        if ((rc = validate_and_prepare(tk))) { cudaStreamSynchronize(stream); n_groups = 0; n_tasks = 0; runs.clear(); return rc; }   // (the caller's buffers are free again)
        K.groups = reinterpret_cast<pe_group *>(b + off[0]);
        K.task_flags = reinterpret_cast<uint8_t *>(b + off[1]);
        K.gens = reinterpret_cast<pe_generic_want *>(b + off[2]);
        K.cons = reinterpret_cast<pe_constraint *>(b + off[3]);
        K.ips = reinterpret_cast<pe_ip_constraint *>(b + off[4]);
        K.plats = reinterpret_cast<pe_platform *>(b + off[5]);
        K.ports = reinterpret_cast<uint32_t *>(b + off[6]);
        K.plugs = reinterpret_cast<uint32_t *>(b + off[7]);
        K.fails = reinterpret_cast<pe_node_fail *>(b + off[8]);
        void *p = d_out_node; size_t c = out_node_cap;
        if ((rc = ensure_buf(p, c, (size_t)std::max(n_tasks, 1u) * 4))) return rc;
        d_out_node = reinterpret_cast<uint32_t *>(p); out_node_cap = c;
        p = d_out_fail; c = out_fail_cap;
        if ((rc = ensure_buf(p, c, (size_t)std::max(n_groups, 1u) * PE_NUM_FILTERS * 4))) return rc;
        d_out_fail = reinterpret_cast<uint32_t *>(p); out_fail_cap = c;
        K.out_node = d_out_node; K.out_fail = d_out_fail;
        if (sync) CU(cudaStreamSynchronize(stream));  // caller buffers may be reused after return
        return PE_OK;
    }

    int32_t ensure_scratch() {
        if (scratch_cap >= cap && ff8) return PE_OK;
        CU(cudaStreamSynchronize(stream));
        auto re = [&](auto *&p, size_t bytes) -> cudaError_t {
            if (p) cudaFree(p);
            p = nullptr;
            return cudaMalloc(reinterpret_cast<void **>(&p), bytes);
        };
        uint32_t sc = 1;
        while (sc < cap) sc <<= 1;
        CU(re(ff8, cap));
        CU(re(pref64, (size_t)cap * 8));
        CU(re(cand_g, (size_t)sc * sizeof(CandKey)));
        CU(re(st_cpu_g, (size_t)sc * 8));
        CU(re(st_mem_g, (size_t)sc * 8));
        CU(re(st_gen_g, (size_t)sc * 8 * PE_MAX_GEN_WANTS));
        CU(re(st_svc_g, (size_t)sc * 4));
        CU(re(st_tot_g, (size_t)sc * 4));
        CU(re(st_placed_g, (size_t)sc * 4));
        CU(re(st_flags_g, (size_t)sc));
        CU(re(touched_g, (size_t)cap / 8 + 256));
        scratch_cap = cap;
        st_cap = sc;
        return PE_OK;
    }

    uint32_t e_stride() const { return round_up(cap / 32, 32); }

    // device arrays of the current k == 1 run / batch (carved from cls_buf / rows_buf)
    uint32_t *ht_full = nullptr, *ht_static = nullptr, *dcls = nullptr, *scls = nullptr, *srow = nullptr, *static_reps = nullptr,
             *mark = nullptr, *rowof = nullptr, *firstof = nullptr;
    uint32_t *row_group = nullptr, *row_srow = nullptr, *task_row = nullptr;

    int32_t launch_sequencer(uint32_t g0, uint32_t g1, bool with_scan, const uint32_t *resume = nullptr) {
        SeqParams P;
        P.resume = resume;
        P.T = table(); P.K = K;
        P.g_begin = g0; P.g_end = g1;
        P.scan = with_scan ? scan_out : nullptr;
        P.task_row = task_row;
        P.E = E; P.e_stride = e_stride(); P.L = Lbuf;
        P.ff8 = ff8; P.pref64 = pref64; P.cand_g = cand_g;
        P.st_cpu_g = st_cpu_g; P.st_mem_g = st_mem_g; P.st_svc_g = st_svc_g; P.st_tot_g = st_tot_g;
        P.st_placed_g = st_placed_g; P.st_flags_g = st_flags_g; P.st_gen_g = st_gen_g; P.st_cap = st_cap;
        P.touched_g = touched_g;
        P.touched_words = with_scan ? (n_nodes + 31) / 32 : 0;
        P.touched_in_smem = P.touched_words <= 12288 ? 1 : 0;   // <= 48 KB of shared memory
        P.ctr = d_ctr;
        size_t dyn = seq_dyn_smem_bytes(P.touched_in_smem ? P.touched_words : 0);
        EvPair *ev = ev_begin(1);
        k_sequencer<<<1, PE_SEQ_THREADS, dyn, stream>>>(P);
        ev_end(ev);
        stats.kernel_launches++;
        CU(cudaGetLastError());
        return PE_OK;
    }

    // A run of task groups: k_groups takes them one after the other on the whole machine; whatever it stops in front of
    // (a group with too many distinct rank prefixes) goes to the one-CTA path, which is exact for anything.
    int32_t launch_groups(uint32_t g0, uint32_t g1) {
        GroupsParams P;
        P.T = table(); P.K = K; P.g_begin = g0; P.g_end = g1;
        P.ff8 = ff8; P.pref64 = pref64; P.cand_g = cand_g;
        P.st_cpu_g = st_cpu_g; P.st_mem_g = st_mem_g; P.st_gen_g = st_gen_g; P.st_svc_g = st_svc_g; P.st_tot_g = st_tot_g;
        P.st_placed_g = st_placed_g; P.st_flags_g = st_flags_g; P.st_cap = st_cap;
        char *b = reinterpret_cast<char *>(groups_buf);
        P.cls_key = reinterpret_cast<unsigned long long *>(b);
        P.cls_cnt = reinterpret_cast<uint32_t *>(b + (size_t)PE_GR_TABLE * 8);
        P.sel = reinterpret_cast<GroupSel *>(b + (size_t)PE_GR_TABLE * 12);
        P.cntmat = reinterpret_cast<uint32_t *>(b + (size_t)PE_GR_TABLE * 12 + sizeof(GroupSel));
        P.resume = d_cls_counters + 3; P.ctr = d_ctr;
        void *args[] = {&P};
        EvPair *ev = ev_begin(1);
        CU(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_groups), dim3((unsigned)groups_grid), dim3(PE_GR_THREADS), args, groups_smem_bytes(), stream));
        ev_end(ev);
        stats.kernel_launches++;
        DBG_SYNC("groups", g0, g1 - g0);
        return launch_sequencer(g0, g1, false, d_cls_counters + 3);
    }

    // Which columns ride in the scan's node tiles and the tile size, from what the
    // k == 1 groups of the staged tick use.  Returns false if the scan cannot stage it.
    bool plan_scan(ScanParams &P) {
        if (!scan_ok) return false;
        struct Tmp { const void *base; uint32_t elem; uint32_t *off32; uint16_t *off16; };
        std::vector<Tmp> cols;
        P.off_total = P.off_cpu = P.off_mem = 0;
        cols.push_back({total, 4, &P.off_total, nullptr});
        if (use_res) { cols.push_back({cpu, 8, &P.off_cpu, nullptr}); cols.push_back({mem, 8, &P.off_mem, nullptr}); }
        memset(P.off_gen, 0xFF, sizeof P.off_gen); memset(P.off_portw, 0xFF, sizeof P.off_portw);
        for (uint32_t c = 0; c < PE_SCAN_MAXGENK && c < gen.size(); c++) if (use_gen[c] && gen[c]) cols.push_back({gen[c], 8, nullptr, &P.off_gen[c]});
        for (uint32_t c = 0; c < PE_SCAN_MAXW && c < ports.size(); c++) if (use_pw[c] && ports[c]) cols.push_back({ports[c], 4, nullptr, &P.off_portw[c]});
        if (cols.size() > PE_SCAN_MAXCOLS) return false;
        uint32_t bpn = 0;
        for (auto &c : cols) bpn += c.elem;
        // largest tile whose two stages fit the per-CTA budget (2 CTAs / SM)
        const uint32_t budget = 100 * 1024;
        uint32_t TN = 2048;
        while (TN > 256 && 2u * TN * bpn > budget) TN >>= 1;
        if (2u * TN * bpn > budget) return false;
        // 8-byte columns first so every column start stays aligned to its element size
        std::stable_sort(cols.begin(), cols.end(), [](const Tmp &a, const Tmp &b) { return a.elem > b.elem; });
        uint32_t o = 0;
        P.n_cols = 0;
        for (auto &c : cols) {
            ScanCol &sc = P.cols[P.n_cols++];
            sc.base = c.base; sc.elem = c.elem; sc.smem_off = o;
            if (c.off32) *c.off32 = o;
            if (c.off16) *c.off16 = (uint16_t)(o / 16);
            o += TN * c.elem;
        }
        P.stage_bytes = o;
        P.tile_nodes = TN;
        P.n_tiles = (n_nodes + TN - 1) / TN;
        P.K = K;
        P.n_nodes = n_nodes;
        P.svc = reinterpret_cast<uint32_t *const *>(d_tab[2]);
        P.S = Sbuf; P.s_stride = e_stride();
        P.ctr = d_ctr;
        // node-axis chunks: enough (row, chunk) warps to fill the machine when descriptors repeat (few rows per batch);
        // every chunk starts on a 32-word boundary of the class bitmaps
        const uint32_t steps = TN / 32u;
        const uint32_t gran = steps >= 32u ? 1u : 32u / steps;
        if (world == 1) {
            n_chunks = local_chunks = std::min<uint32_t>(4u, std::max<uint32_t>(1u, P.n_tiles / gran));
        } else {
            local_chunks = std::max<uint32_t>(1u, 4u / (uint32_t)world);
            n_chunks = local_chunks * (uint32_t)world;
            if (n_chunks > 32u) return false;
        }
        P.tiles_per_chunk = round_up(std::max<uint32_t>(1u, (P.n_tiles + n_chunks - 1) / n_chunks), gran);
        P.n_chunks = n_chunks; P.chunk0 = (uint32_t)rank * local_chunks;
        P.Eout = E; P.e_row_stride = e_stride(); P.e_word_off = 0;   // (multi-rank: set per batch in run_k1)
        return true;
    }

    // One maximal run [g, e) of k == 1 groups through classify -> (static) -> per batch rows / scan / sequencer.
    int32_t run_k1(uint32_t g, uint32_t e, uint32_t Bmax) {
        int32_t rc;
        const uint32_t n = e - g;
        ScanParams SP;
        // (node sharding exchanges what the parallel placement step needs, not the class bitmaps the ordered sequencer's
        // scan mode walks: with PE_CFG_ORDERED_ONLY a sharded engine places from the live table alone)
        if (!plan_scan(SP) || (world > 1 && (cfg_flags & PE_CFG_ORDERED_ONLY))) {
            for (uint32_t b0 = g; b0 < e; b0 += Bmax) if ((rc = launch_sequencer(b0, std::min(e, b0 + Bmax), false))) return rc;
            return PE_OK;
        }
        uint32_t ht = 1024;
        while (ht < 2u * n) ht <<= 1;
        if ((rc = ensure_buf(cls_buf, cls_cap, ((size_t)2 * ht + (size_t)7 * n) * 4 + 64))) return rc;
        ht_full = reinterpret_cast<uint32_t *>(cls_buf); ht_static = ht_full + ht;
        dcls = ht_static + ht; scls = dcls + n; srow = scls + n; static_reps = srow + n; mark = static_reps + n; rowof = mark + n; firstof = rowof + n;
        if ((rc = ensure_buf(rows_buf, rows_cap, (size_t)3 * Bmax * 4 + 64))) return rc;
        row_group = reinterpret_cast<uint32_t *>(rows_buf); row_srow = row_group + Bmax; task_row = row_srow + Bmax;
        // per-chunk partial results: p1 [chunks][Bmax] 16 B | Cc [chunks][Bmax][2] | Lc [chunks][Bmax][2][PE_LIST_CAP]
        const size_t p1_bytes = (size_t)n_chunks * Bmax * 16, cc_bytes = (size_t)n_chunks * Bmax * 8,
                     lc_bytes = (size_t)n_chunks * Bmax * 2 * PE_LIST_CAP * 4;
        if ((rc = ensure_buf(chunk_buf, chunk_cap, p1_bytes + cc_bytes + lc_bytes + 64))) return rc;
        SP.p1 = reinterpret_cast<ulonglong2 *>(chunk_buf);
        SP.Cc = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(chunk_buf) + p1_bytes);
        SP.Lc = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(chunk_buf) + p1_bytes + cc_bytes);
        SP.rows_cap = Bmax;
        MergeParams MP;
        MP.K = K; MP.svc = SP.svc; MP.n_chunks = n_chunks; MP.rows_cap = Bmax; MP.p1 = SP.p1; MP.Lc = SP.Lc; MP.Cc = SP.Cc;
        MP.out = scan_out; MP.L = Lbuf; MP.Eall = nullptr; MP.E = E; MP.n_ranks = 1; MP.seg_words = 0; MP.e_stride = e_stride();
        // the chunked parallel placement step (kernel_place.cuh) places what it can of every batch; the ordered sequencer
        // continues from where it stopped (nothing, normally)
        const bool parallel = !(cfg_flags & PE_CFG_ORDERED_ONLY);
        MP.touched = nullptr; MP.touched_words = 0; MP.cursors = nullptr;
        if (parallel) {
            void *p = cursors; size_t c = cursors_cap;
            if ((rc = ensure_buf(p, c, (size_t)Bmax * 8 + 64))) return rc;
            cursors = reinterpret_cast<uint32_t *>(p); cursors_cap = c;
            MP.touched = touched_g; MP.touched_words = (n_nodes + 31) / 32; MP.cursors = cursors;
        }
        MP.Lx = nullptr; MP.x_cap = 0;
        auto gather = [&](void *base, size_t bytes_per_rank) -> int32_t {
            const pe_nccl::Api &nc = pe_nccl::api();
            const pe_nccl::ncclResult_t r = nc.AllGather(reinterpret_cast<char *>(base) + (size_t)rank * bytes_per_rank, base, bytes_per_rank,
                                                         pe_nccl::ncclUint8, comm, stream);
            if (r != 0) { err = std::string("ncclAllGather: ") + nc.GetErrorString(r); return PE_ERR_CUDA; }
            return PE_OK;
        };

        EvPair *evp = ev_begin(5);
        CU(cudaMemsetAsync(ht_full, 0xFF, (size_t)2 * ht * 4, stream));
        CU(cudaMemsetAsync(mark, 0, (size_t)n * 4, stream));
        CU(cudaMemsetAsync(d_cls_counters, 0, 16, stream));
        ClassifyParams CP;
        CP.K = K; CP.g0 = g; CP.n = n; CP.ht_full = ht_full; CP.ht_static = ht_static; CP.ht_mask = ht - 1;
        CP.dcls = dcls; CP.scls = scls; CP.srow = srow; CP.static_reps = static_reps; CP.counters = d_cls_counters;
        k_classify<<<std::min<uint32_t>((n + 255) / 256, (uint32_t)num_sms * 8u), 256, 0, stream>>>(CP);
        stats.kernel_launches++;
        // How many signature bitmaps are there?  A short run (a small tick: the latency case) does not ask: it sizes for
        // one per task, so that nothing waits on the host in the middle of the tick.
        const size_t stride = e_stride();
        uint32_t n_sig = n, n_cls = n;      // signatures / descriptor classes of the run (upper bounds for a short run)
        if ((size_t)n * stride * 4 > ((size_t)256 << 20)) {
            uint32_t h_cnt[4] = {0, 0, 0, 0};
            CU(cudaMemcpyAsync(h_cnt, d_cls_counters, 16, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            n_sig = h_cnt[0]; n_cls = h_cnt[1];
        }
        // ---- node sharding.  Per batch the ranks exchange: the chunks' two smallest rank prefixes (16 B per row
        // and chunk, all-gather), the chunks' class sizes (8 B, all-gather), and the merged member lists -- x_cap members per
        // (row, class) in ONE all-reduce over a buffer every rank fills only where its own chunks' members belong (k_xpack).
        // No rank ever waits on the host: the arrays are strided by a bound on the rows of a batch (the run's descriptor
        // classes), not by the batch's own row count.
        uint32_t rcap = Bmax, x_cap = PE_LIST_CAP;
        uint32_t *Lx = nullptr;
        if (world > 1) {
            rcap = std::max<uint32_t>(8u, round_up(std::min(Bmax, n_cls), 8u));
            x_cap = 256u;
            while (x_cap < PE_LIST_CAP && x_cap < Bmax / 4u) x_cap <<= 1;     // a batch consumes about B x (row density) members of a class
            const size_t need = (size_t)rcap * 2 * x_cap;
            if (need > Eall_words) {
                void *p = Eall; size_t c = Eall_words * 4;
                if ((rc = ensure_buf(p, c, need * 4))) return rc;
                Eall = reinterpret_cast<uint32_t *>(p); Eall_words = c / 4;
            }
            Lx = Eall;
            const size_t p1b = (size_t)n_chunks * rcap * 16, ccb = (size_t)n_chunks * rcap * 8;
            SP.Cc = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(chunk_buf) + p1b);
            SP.Lc = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(chunk_buf) + p1b + ccb);
            SP.rows_cap = rcap;
            MP.rows_cap = rcap; MP.Cc = SP.Cc; MP.Lc = SP.Lc; MP.Lx = Lx; MP.x_cap = x_cap;
        }
        // signature bitmaps for the whole run when they fit in 1 GB, else one bitmap per row of each batch
        const bool cached = (size_t)n_sig * stride * 4 <= ((size_t)1 << 30);
        const size_t s_rows = cached ? n_sig : Bmax;
        if (s_rows * stride > S_words) {
            void *p = Sbuf; size_t c = S_words * 4;
            if ((rc = ensure_buf(p, c, s_rows * stride * 4))) return rc;
            Sbuf = reinterpret_cast<uint32_t *>(p); S_words = c / 4;
        }
        SP.S = Sbuf;
        StaticParams XP;
        XP.T = table(); XP.K = K; XP.S = Sbuf; XP.s_stride = (uint32_t)stride; XP.lo = 0; XP.hi = n_nodes; XP.ctr = d_ctr;
        const uint32_t chunks = (uint32_t)stride / 32u;
        auto launch_static = [&](const uint32_t *reps, const uint32_t *cnt, uint32_t max_rows) {
            XP.reps = reps; XP.n_reps = cnt;
            const unsigned long long units = (unsigned long long)max_rows * chunks;
            const uint32_t grid = (uint32_t)std::min<unsigned long long>((units + 7) / 8, (unsigned long long)num_sms * 16ull);
            k_static<<<std::max(grid, 1u), 256, 0, stream>>>(XP);
            stats.kernel_launches++;
        };
        if (cached && n_sig) launch_static(static_reps, d_cls_counters, n_sig);
        ev_end(evp);

        uint32_t stamp = 0;
        for (uint32_t b0 = g; b0 < e; b0 += Bmax) {
            const uint32_t B = std::min(Bmax, e - b0);
            EvPair *evr = ev_begin(5);
            RowsParams RP;
            RP.dcls = dcls; RP.scls = scls; RP.srow = srow; RP.g0 = g; RP.b0 = b0; RP.B = B; RP.mark = mark; RP.rowof = rowof; RP.firstof = firstof;
            RP.stamp = ++stamp; RP.static_cached = cached ? 1u : 0u;
            RP.row_group = row_group; RP.row_srow = row_srow; RP.task_row = task_row; RP.n_rows = d_cls_counters + 2;
            k_rows<<<1, 1024, 0, stream>>>(RP);
            stats.kernel_launches++;
            if (!cached) launch_static(row_group, d_cls_counters + 2, B);
            ev_end(evr);
            SP.row_group = row_group; SP.row_srow = row_srow; SP.n_rows = d_cls_counters + 2;
            // grid.x: enough CTAs that rows-per-CTA adapts to the row count (k_scan reads it from the device)
            SP.target_ctas = (uint32_t)num_sms * 2u;
            const dim3 grid(std::max((B + PE_SCAN_WARPS - 1) / PE_SCAN_WARPS, std::min(B, SP.target_ctas)), local_chunks);
            const size_t dyn = 2 * (size_t)SP.stage_bytes;
            MP.row_group = row_group; MP.n_rows = d_cls_counters + 2;
            EvPair *ev = ev_begin(0);
            if (use_dyn) k_scan<true, 1><<<grid, PE_SCAN_THREADS, dyn, stream>>>(SP);
            else k_scan<false, 1><<<grid, PE_SCAN_THREADS, dyn, stream>>>(SP);
            if (world > 1 && (rc = gather(SP.p1, (size_t)local_chunks * rcap * 16))) return rc;
            if (use_dyn) k_scan<true, 2><<<grid, PE_SCAN_THREADS, dyn, stream>>>(SP);
            else k_scan<false, 2><<<grid, PE_SCAN_THREADS, dyn, stream>>>(SP);
            if (world > 1) {
                if ((rc = gather(SP.Cc, (size_t)local_chunks * rcap * 8))) return rc;
                CU(cudaMemsetAsync(Lx, 0, (size_t)rcap * 2 * x_cap * 4, stream));
                XPackParams XQ;
                XQ.n_rows = d_cls_counters + 2; XQ.n_chunks = n_chunks; XQ.chunk0 = SP.chunk0; XQ.local_chunks = local_chunks; XQ.rows_cap = rcap;
                XQ.Cc = SP.Cc; XQ.Lc = SP.Lc; XQ.Lx = Lx; XQ.x_cap = x_cap;
                k_xpack<<<(B + 7) / 8, 256, 0, stream>>>(XQ);
                stats.kernel_launches++;
                const pe_nccl::Api &nc = pe_nccl::api();
                const pe_nccl::ncclResult_t r = nc.AllReduce(Lx, Lx, (size_t)rcap * 2 * x_cap, pe_nccl::ncclUint32, pe_nccl::ncclSum, comm, stream);
                if (r != 0) { err = std::string("ncclAllReduce: ") + nc.GetErrorString(r); return PE_ERR_CUDA; }
            }
            DBG_SYNC("scan", b0, B);
            k_merge<<<std::min<uint32_t>(2u * B, (uint32_t)num_sms * 8u), 256, 0, stream>>>(MP);   // CTAs stride over (row, class) units
            ev_end(ev);
            DBG_SYNC("merge", b0, B);
            CU(cudaGetLastError());
            stats.kernel_launches += 3; stats.scan_launches++;
            stats.pairs += (uint64_t)B * n_nodes;
            if (parallel) {
                PlaceParams PP;
                PP.T = table(); PP.K = K; PP.b0 = b0; PP.B = B; PP.scan = scan_out; PP.task_row = task_row; PP.L = Lbuf;
                PP.touched = touched_g; PP.cursors = cursors; PP.resume = d_cls_counters + 3; PP.ctr = d_ctr;
                PP.list_cap = world > 1 ? x_cap : (uint32_t)PE_LIST_CAP;
                const uint32_t tkw = (n_nodes + 31) / 32;
                PP.tk_words = tkw <= PE_PL_TK_MAX_WORDS ? tkw : 0u;
                EvPair *evq = ev_begin(6);
                {
                    cudaLaunchConfig_t lc = {};
                    lc.gridDim = dim3(place_cluster); lc.blockDim = dim3(PE_PL_THREADS);
                    lc.dynamicSmemBytes = place_smem_bytes(PP.tk_words); lc.stream = stream;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = place_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    lc.attrs = at; lc.numAttrs = 1;
                    CU(cudaLaunchKernelEx(&lc, k_place, PP));
                }
                ev_end(evq);
                stats.kernel_launches++;
                CU(cudaGetLastError());
                DBG_SYNC("place", b0, B);
            }
            if ((rc = launch_sequencer(b0, b0 + B, world == 1, parallel ? d_cls_counters + 3 : nullptr))) return rc;
            DBG_SYNC("sequencer", b0, B);
        }
        return PE_OK;
    }

    DevCounters *h_ctr = nullptr;      // pinned landing area of the device counters

    int32_t tick_run() {
        int32_t rc = tick_enqueue();
        if (rc) return rc;
        CU(cudaStreamSynchronize(stream));
        return counters_finish();
    }

    // everything of pe_tick_run that is stream work; the counters come back asynchronously into pinned memory
    int32_t tick_enqueue() {
        int32_t rc;
        if ((rc = sync_tabs())) return rc;
        if ((rc = ensure_scratch())) return rc;
        // batch = tasks placed against one scan.  Every batch costs one scan of (distinct descriptors x nodes), so batches
        // want to be long; a batch that consumes more of a row's class than its member list holds (PE_LIST_CAP, about
        // batch x the row's node density) hands the rest to the ordered sequencer, so they must not be too long: a sixth of
        // the nodes, at most 16384 (measured on cfg3: 4736 -> 16384 takes the scan from 115 to 34 ms per million tasks).
        // (a multiple of 8, at least 8: the multi-rank exchange strides its arrays by the batch's row count rounded up to 8)
        // Feedback from the previous tick (its counters are on the host by now): a batch the placement step had to hand to
        // the ordered sequencer -- a task found both recorded classes of its row consumed, which is what happens when the
        // nodes carry many tasks and every (service count, total) class is small -- costs milliseconds on one SM.  Such
        // ticks are followed by ticks with shorter batches (every batch starts from a fresh scan); clean ticks grow them
        // back.  The batch size never changes a placement (tests/test_headline_gpu.py::test_cfg3_oneoff_batch_sizes).
        uint32_t Bmax = max_batch ? std::max(8u, round_up(max_batch, 8u)) : std::min(16384u, std::max(256u, round_up(n_nodes / 6u, 16u)));
        if (!max_batch && batch_shift) Bmax = std::max(std::min(Bmax, 1024u), round_up(Bmax >> batch_shift, 16u));
        const bool spec = !(cfg_flags & PE_CFG_NO_SPECULATION) && n_nodes > 0;
        if (spec) {
            size_t need = (size_t)Bmax * 2u * e_stride();   // two class rows per scan row
            if (need > E_words) {
                void *p = E; size_t c = E_words * 4;
                if ((rc = ensure_buf(p, c, need * 4))) return rc;
                E = reinterpret_cast<uint32_t *>(p); E_words = c / 4;
            }
            if ((size_t)Bmax * 2u * PE_LIST_CAP > L_words) {
                void *p = Lbuf; size_t c = L_words * 4;
                if ((rc = ensure_buf(p, c, (size_t)Bmax * 2u * PE_LIST_CAP * 4))) return rc;
                Lbuf = reinterpret_cast<uint32_t *>(p); L_words = c / 4;
            }
            if (Bmax > scan_out_cap) {
                void *p = scan_out; size_t c = (size_t)scan_out_cap * sizeof(ScanResult);
                if ((rc = ensure_buf(p, c, (size_t)Bmax * sizeof(ScanResult)))) return rc;
                scan_out = reinterpret_cast<ScanResult *>(p); scan_out_cap = Bmax;
            }
        }
        // failure counters default to zero; only groups with unplaced tasks write theirs
        if (n_groups) CU(cudaMemsetAsync(d_out_fail, 0, (size_t)n_groups * PE_NUM_FILTERS * 4, stream));
        EvPair *ev_run = ev_begin(4);
        for (const Run &r : runs) {
            // maximal run of k == 1 groups -> batched scan path; anything else -> sequencer alone
            if (r.one && spec && r.end - r.begin >= 4) {
                if ((rc = run_k1(r.begin, r.end, Bmax))) return rc;
            } else if (!r.one && !r.seq_only && groups_grid && !(cfg_flags & PE_CFG_ORDERED_ONLY)) {
                if ((rc = launch_groups(r.begin, r.end))) return rc;
            } else {
                if ((rc = launch_sequencer(r.begin, r.end, false))) return rc;
            }
        }
        ev_end(ev_run);
        return counters_enqueue();
    }

    int32_t counters_enqueue() {
        if (!h_ctr) CU(cudaMallocHost(reinterpret_cast<void **>(&h_ctr), sizeof(DevCounters)));
        CU(cudaMemcpyAsync(h_ctr, d_ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, stream));
        CU(cudaMemsetAsync(d_ctr, 0, sizeof(DevCounters), stream));
        return PE_OK;
    }

    int32_t collect_counters() {       // (synchronous form: pe_fit)
        int32_t rc = counters_enqueue();
        if (rc) return rc;
        CU(cudaStreamSynchronize(stream));
        return counters_finish();
    }

    // after the stream has drained
    int32_t counters_finish() {
        const DevCounters &c = *h_ctr;
        stats.fast_path += c.fast_path; stats.medium_path += c.medium_path; stats.slow_path += c.slow_path;
        stats.placements += c.placements; stats.evals_generic += c.evals_generic;
        stats.seq_cycles_fast += c.cyc_fast; stats.seq_cycles_medium += c.cyc_medium; stats.seq_cycles_generic += c.cyc_generic;
        for (int r = 0; r < 5; r++) stats.seq_stops[r] += c.stops[r];
        for (int r = 0; r < 16; r++) stats.seq_prof[r] += c.prof[r];
        stats.seq_cons_wait += c.cyc_cons_wait; stats.seq_cons_work += c.cyc_cons_work; stats.seq_rewalks += c.iters;
        stats.evals += c.scan_evals; stats.scan_bytes += c.scan_bytes; stats.static_evals += c.static_evals; stats.scan_rows += c.scan_rows;
        if (c.place_cuts) batch_shift = std::min(batch_shift + 1u, 4u);          // (see tick_enqueue)
        else if (c.place_tasks && batch_shift) batch_shift--;
        stats.place_tasks += c.place_tasks; stats.place_cuts += c.place_cuts; stats.place_amb += c.place_amb; stats.place_tails += c.place_tails;
        stats.place_chunks += c.place_chunks;
        for (int r = 0; r < 3; r++) stats.place_cyc[r] += c.place_cyc[r];
        ev_collect();
        if (c.error & (PE_DEV_ERR_WD_CONSUMER | PE_DEV_ERR_WD_DRAIN | PE_DEV_ERR_WD_SCAN)) {
            char b[160]; snprintf(b, sizeof b, "device watchdog fired (code 0x%x): a kernel pipeline stalled for more than a second", c.error);
            err = b; return PE_ERR_CUDA;
        }
        if (c.error & PE_DEV_ERR_SVC_OVERFLOW) { err = "a per-service task count reached 2^24 - 1 on one node"; return PE_ERR_OVERFLOW; }
        return PE_OK;
    }

    int32_t tick_download(uint32_t *out_node, uint32_t *out_fail, bool sync = true) {
        EvPair *ev = ev_begin(3);
        if (out_node && n_tasks) { CU(cudaMemcpyAsync(out_node, d_out_node, (size_t)n_tasks * 4, cudaMemcpyDeviceToHost, stream)); stats.d2h_bytes += (size_t)n_tasks * 4; }
        // (pe_schedule: the failure counters are wanted only if some task went unplaced, which the tick's own counters say
        // once the stream has drained -- finish_schedule() fetches them then; 32 bytes per group otherwise cross PCIe for nothing)
        pending_fail = nullptr;
        if (out_fail && n_groups) {
            if (sync) { CU(cudaMemcpyAsync(out_fail, d_out_fail, (size_t)n_groups * PE_NUM_FILTERS * 4, cudaMemcpyDeviceToHost, stream)); stats.d2h_bytes += (size_t)n_groups * PE_NUM_FILTERS * 4; }
            else pending_fail = out_fail;
        }
        ev_end(ev);
        if (!sync) return PE_OK;
        CU(cudaStreamSynchronize(stream));
        ev_collect();
        return PE_OK;
    }

    int32_t fit(const pe_tick *tk, const uint32_t *node_idx, uint8_t *out_ok, uint32_t *out_fail) {
        int32_t rc = tick_upload(tk);
        if (rc) return rc;
        if (!n_groups) return PE_OK;
        if ((rc = ensure_buf(up_buf, up_cap, (size_t)n_groups * 5 + 64))) return rc;
        uint32_t *d_idx = reinterpret_cast<uint32_t *>(up_buf);
        uint8_t *d_ok = reinterpret_cast<uint8_t *>(d_idx + n_groups);
        CU(cudaMemcpyAsync(d_idx, node_idx, (size_t)n_groups * 4, cudaMemcpyHostToDevice, stream));
        k_fit<<<1, 32, 0, stream>>>(table(), K, n_groups, d_idx, d_ok, d_ctr);
        stats.kernel_launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out_ok, d_ok, n_groups, cudaMemcpyDeviceToHost, stream));
        if (out_fail) CU(cudaMemcpyAsync(out_fail, d_out_fail, (size_t)n_groups * PE_NUM_FILTERS * 4, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return collect_counters();
    }

    uint32_t *pending_fail = nullptr;
    int32_t finish_schedule() {
        CU(cudaStreamSynchronize(stream));
        const bool all_placed = h_ctr && h_ctr->placements == (unsigned long long)n_tasks;
        int32_t rc = counters_finish();      // (also collects the event timings)
        if (rc == PE_OK && pending_fail && !all_placed) {
            CU(cudaMemcpyAsync(pending_fail, d_out_fail, (size_t)n_groups * PE_NUM_FILTERS * 4, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
            stats.d2h_bytes += (size_t)n_groups * PE_NUM_FILTERS * 4;
        }
        pending_fail = nullptr;
        return rc;
    }

    // nodeSet.tree's leaves and their task sums for one service (nodeset.go:59-101); see kernel_misc.cuh
    void *pref_buf = nullptr; size_t pref_cap = 0;
    int32_t pref_leaves(uint32_t svc_id, const uint32_t *cols, uint32_t n_levels, uint32_t out_cap, uint32_t *out_vals, uint32_t *out_tasks, uint32_t *out_n) {
        if (n_levels > PE_MAX_PREF_LEVELS || !out_n || (n_levels && !cols) || (out_cap && (!out_tasks || (n_levels && !out_vals)))) { err = "pref_leaves: bad arguments"; return PE_ERR_INVALID; }
        int32_t rc;
        if ((rc = sync_tabs())) return rc;
        for (uint32_t l = 0; l < n_levels; l++) if ((rc = ensure_col(attr, cols[l], 4))) return rc;
        if ((svc_id >= svc.size() || !svc[svc_id]) && (rc = ensure_col(svc, svc_id, 4))) return rc;
        if ((rc = sync_tabs())) return rc;
        uint32_t tsz = 1024;
        while (tsz < 2u * std::max(n_nodes, 1u)) tsz <<= 1;
        const size_t lv = std::max(n_levels, 1u);
        const size_t bytes = (size_t)tsz * 16 + ((size_t)out_cap * (lv + 1) + 2) * 4 + 64;
        if ((rc = ensure_buf(pref_buf, pref_cap, bytes))) return rc;
        PrefParams P;
        P.T = table(); P.svccol = svc[svc_id]; P.n_levels = n_levels; P.mask = tsz - 1; P.cap = out_cap;
        for (uint32_t l = 0; l < PE_MAX_PREF_LEVELS; l++) P.cols[l] = l < n_levels ? cols[l] : 0u;
        char *b = reinterpret_cast<char *>(pref_buf);
        P.keys = reinterpret_cast<unsigned long long *>(b);
        P.sums = reinterpret_cast<uint32_t *>(b + (size_t)tsz * 8);
        P.reps = P.sums + tsz;
        P.out_n = P.reps + tsz; P.out_err = P.out_n + 1; P.out_tasks = P.out_err + 1; P.out_vals = P.out_tasks + out_cap;
        CU(cudaMemsetAsync(b, 0, (size_t)tsz * 12, stream));
        CU(cudaMemsetAsync(P.reps, 0xFF, (size_t)tsz * 4, stream));
        CU(cudaMemsetAsync(P.out_n, 0, 8, stream));
        const uint32_t grid = std::max(1u, std::min<uint32_t>((std::max(n_nodes, tsz) + 255u) / 256u, (uint32_t)num_sms * 8u));
        k_pref_leaves<<<grid, 256, 0, stream>>>(P);
        k_pref_emit<<<grid, 256, 0, stream>>>(P);
        stats.kernel_launches += 2;
        CU(cudaGetLastError());
        uint32_t head[2] = {0, 0};
        CU(cudaMemcpyAsync(head, P.out_n, 8, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        if (head[1]) { err = "pref_leaves: two leaves share a 64-bit fingerprint"; return PE_ERR_UNSUPPORTED; }
        if (head[0] > out_cap) { err = "pref_leaves: more leaves than the caller's buffers hold"; return PE_ERR_OVERFLOW; }
        if (head[0]) {
            CU(cudaMemcpyAsync(out_tasks, P.out_tasks, (size_t)head[0] * 4, cudaMemcpyDeviceToHost, stream));
            if (n_levels) CU(cudaMemcpyAsync(out_vals, P.out_vals, (size_t)head[0] * n_levels * 4, cudaMemcpyDeviceToHost, stream));
            CU(cudaStreamSynchronize(stream));
        }
        *out_n = head[0];
        return PE_OK;
    }

    // (service, node) predicate matrix for the enforcer / global orchestrator: k_static over every group
    int32_t match_matrix(const pe_tick *tk, uint32_t *out_bits) {
        int32_t rc;
        if (!out_bits) { err = "match_matrix: null output"; return PE_ERR_INVALID; }
        if ((rc = tick_upload(tk, true))) return rc;
        if (!n_groups || !n_nodes) return PE_OK;
        if ((rc = ensure_scratch())) return rc;
        const size_t stride = e_stride();
        if ((size_t)n_groups * stride > S_words) {
            void *p = Sbuf; size_t c = S_words * 4;
            if ((rc = ensure_buf(p, c, (size_t)n_groups * stride * 4))) return rc;
            Sbuf = reinterpret_cast<uint32_t *>(p); S_words = c / 4;
        }
        if ((rc = ensure_buf(cls_buf, cls_cap, ((size_t)n_groups + 1) * 4 + 64))) return rc;
        uint32_t *reps = reinterpret_cast<uint32_t *>(cls_buf);
        std::vector<uint32_t> iota((size_t)n_groups + 1);
        for (uint32_t i = 0; i < n_groups; i++) iota[i] = i;
        iota[n_groups] = n_groups;                                  // the row count k_static reads from the device
        CU(cudaMemcpyAsync(reps, iota.data(), iota.size() * 4, cudaMemcpyHostToDevice, stream));
        StaticParams XP;
        XP.T = table(); XP.K = K; XP.S = Sbuf; XP.s_stride = (uint32_t)stride; XP.lo = 0; XP.hi = n_nodes; XP.ctr = d_ctr;
        XP.reps = reps; XP.n_reps = reps + n_groups;
        const unsigned long long units = (unsigned long long)n_groups * (stride / 32u);
        const uint32_t grid = (uint32_t)std::min<unsigned long long>((units + 7) / 8, (unsigned long long)num_sms * 16ull);
        k_static<<<std::max(grid, 1u), 256, 0, stream>>>(XP);
        stats.kernel_launches++;
        CU(cudaGetLastError());
        const size_t words = (n_nodes + 31) / 32;
        CU(cudaMemcpy2DAsync(out_bits, words * 4, Sbuf, stride * 4, words * 4, n_groups, cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));                          // (also: `iota` may go now)
        stats.d2h_bytes += words * 4 * n_groups;
        return PE_OK;
    }

    template <class T> int32_t snap_col(const T *col, uint32_t first, uint32_t n, T *out) {
        if ((uint64_t)first + n > cap) { err = "snapshot range out of bounds"; return PE_ERR_INVALID; }
        if (!col) { std::memset(out, 0, (size_t)n * sizeof(T)); return PE_OK; }
        CU(cudaMemcpyAsync(out, col + first, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        return PE_OK;
    }
};

// ============================ C ABI ==========================================
extern "C" {

uint32_t pe_abi_version(void) { return PE_ABI_VERSION; }

int32_t pe_create(const pe_config *cfg, pe_engine **out) {
    if (!cfg || !out) { g_create_err = "null argument"; return PE_ERR_INVALID; }
    if (cfg->abi_version != PE_ABI_VERSION) { g_create_err = "ABI version mismatch"; return PE_ERR_INVALID; }
    pe_engine *h = new pe_engine();
    int32_t rc = h->init(cfg);
    if (rc) { g_create_err = h->err; h->destroy(); delete h; return rc; }
    *out = h;
    return PE_OK;
}

void pe_destroy(pe_engine *h) {
    if (!h) return;
    h->destroy();
    delete h;
}

const char *pe_last_error(const pe_engine *h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int32_t pe_node_upsert(pe_engine *h, const pe_node_row *rows, uint32_t n_rows, const pe_kv32 *attrs, const pe_kv64 *gens,
                       const pe_kv32 *svcs, const uint32_t *ports, const uint32_t *plugs) {
    return h->upsert(rows, n_rows, attrs, gens, svcs, ports, plugs);
}
int32_t pe_node_remove(pe_engine *h, const uint32_t *idx, uint32_t n) { return h->remove(idx, n); }
int32_t pe_set_node_count(pe_engine *h, uint32_t n) {
    int32_t rc = h->ensure_cap(n);
    if (rc) return rc;
    h->n_nodes = n;
    return PE_OK;
}
int32_t pe_node_task_delta(pe_engine *h, const pe_task_delta *d, uint32_t n, const pe_kv64 *gens, const uint32_t *ports) {
    return h->delta(d, n, gens, ports);
}

int32_t pe_tick_upload(pe_engine *h, const pe_tick *tick) { return h->tick_upload(tick); }
int32_t pe_tick_run(pe_engine *h) { return h->tick_run(); }
int32_t pe_tick_download(pe_engine *h, uint32_t *out_node, uint32_t *out_fail) { return h->tick_download(out_node, out_fail); }

int32_t pe_schedule(pe_engine *h, const pe_tick *tick, uint32_t *out_node, uint32_t *out_fail) {
    // one wait for the whole call: descriptors in, kernels, placements + counters out, all in stream order
    int32_t rc = h->tick_upload(tick, false);
    if (rc) return rc;
    if ((rc = h->tick_enqueue())) return rc;
    if ((rc = h->tick_download(out_node, out_fail, false))) return rc;
    return h->finish_schedule();
}

int32_t pe_fit(pe_engine *h, const pe_tick *tick, const uint32_t *node_idx, uint8_t *out_ok, uint32_t *out_fail) {
    return h->fit(tick, node_idx, out_ok, out_fail);
}

int32_t pe_snapshot(pe_engine *h, uint32_t first, uint32_t n, pe_node_state *out) {
    if ((uint64_t)first + n > h->cap) { h->err = "snapshot range out of bounds"; return PE_ERR_INVALID; }
    std::vector<uint32_t> m(n), t(n);
    std::vector<int64_t> c(n), mm(n);
    int32_t rc;
    if ((rc = h->snap_col(h->meta, first, n, m.data()))) return rc;
    if ((rc = h->snap_col(h->total, first, n, t.data()))) return rc;
    if ((rc = h->snap_col(h->cpu, first, n, c.data()))) return rc;
    if ((rc = h->snap_col(h->mem, first, n, mm.data()))) return rc;
    for (uint32_t i = 0; i < n; i++) { out[i].flags = m[i] & PE_META_FLAGS_MASK; out[i].total_tasks = t[i]; out[i].cpu_avail = c[i]; out[i].mem_avail = mm[i]; }
    return PE_OK;
}
int32_t pe_snapshot_service(pe_engine *h, uint32_t s, uint32_t first, uint32_t n, uint32_t *out) {
    return h->snap_col(s < h->svc.size() ? h->svc[s] : (uint32_t *)nullptr, first, n, out);
}
int32_t pe_snapshot_generic(pe_engine *h, uint32_t kd, uint32_t first, uint32_t n, int64_t *out) {
    return h->snap_col(kd < h->gen.size() ? h->gen[kd] : (int64_t *)nullptr, first, n, out);
}
int32_t pe_snapshot_ports(pe_engine *h, uint32_t slot, uint32_t first, uint32_t n, uint8_t *out) {
    std::vector<uint32_t> w(n);
    int32_t rc = h->snap_col((slot >> 5) < h->ports.size() ? h->ports[slot >> 5] : (uint32_t *)nullptr, first, n, w.data());
    if (rc) return rc;
    for (uint32_t i = 0; i < n; i++) out[i] = (w[i] >> (slot & 31)) & 1;
    return PE_OK;
}

int32_t pe_pref_leaves(pe_engine *h, uint32_t svc_id, const uint32_t *cols, uint32_t n_levels, uint32_t cap, uint32_t *out_vals,
                       uint32_t *out_tasks, uint32_t *out_n_leaves) {
    return h->pref_leaves(svc_id, cols, n_levels, cap, out_vals, out_tasks, out_n_leaves);
}

int32_t pe_match_matrix(pe_engine *h, const pe_tick *tick, uint32_t *out_bits) { return h->match_matrix(tick, out_bits); }

int32_t pe_get_stats(pe_engine *h, pe_stats *out) { *out = h->stats; return PE_OK; }
int32_t pe_stats_reset(pe_engine *h) { h->stats = pe_stats{}; return PE_OK; }

int32_t pe_nccl_unique_id(void *out128) {
    if (!out128) return PE_ERR_INVALID;
    const pe_nccl::Api &nc = pe_nccl::api();
    if (!nc.ok) { g_create_err = "libnccl.so.2 is not available"; return PE_ERR_UNSUPPORTED; }
    pe_nccl::ncclUniqueId id;
    const pe_nccl::ncclResult_t r = nc.GetUniqueId(&id);
    if (r != 0) { g_create_err = std::string("ncclGetUniqueId: ") + nc.GetErrorString(r); return PE_ERR_CUDA; }
    memcpy(out128, &id, sizeof id);
    return PE_OK;
}

int32_t pe_fold_value(const char *in, uint32_t len, char *out, uint32_t cap) {
    uint32_t w = 0;
    uint32_t i = 0;
    while (i < len) {
        const unsigned char c = (unsigned char)in[i];
        char r;
        uint32_t adv = 1;
        if (c >= 'A' && c <= 'Z') r = (char)(c | 0x20);
        else if (c == 0xE2 && i + 2 < len && (unsigned char)in[i + 1] == 0x84 && (unsigned char)in[i + 2] == 0xAA) { r = 'k'; adv = 3; }  // KELVIN SIGN
        else if (c == 0xC5 && i + 1 < len && (unsigned char)in[i + 1] == 0xBF) { r = 's'; adv = 2; }                                      // LATIN SMALL LETTER LONG S
        else r = (char)c;
        if (w >= cap) return -1;
        out[w++] = r;
        i += adv;
    }
    return (int32_t)w;
}

}  // extern "C"
