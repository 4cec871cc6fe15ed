// kernel_scan.cuh -- the batched (task, node) filter + rank scan: THE hot kernel.
//
// One warp owns one pending k=1 task group ("one-off", scheduler.go:456-459,
// 467-469 -- the shape BenchmarkScheduler100kNodes1MTasks exercises,
// scheduler_test.go:3358,3378-3468).  The CTA's warps share node tiles: each
// tile is a slice of every SoA column the batch needs, brought from HBM/L2 into
// shared memory with cp.async.bulk (TMA bulk copy, mbarrier completion),
// double-buffered.  Lane l of a warp evaluates node (tile*TN + step*32 + l):
//
//   Pipeline.Process   pipeline.go:56-68   (AND of the enabled filters)
//   ReadyFilter        filter.go:40-43     meta bit
//   ResourceFilter     filter.go:76-93     signed 64-bit compares + generic cells
//   PluginFilter       filter.go:141-183   plugin bit words
//   ConstraintFilter   filter.go:241-243   (col == value) ^ neq per expression
//   PlatformFilter     filter.go:272-312   os/arch bytes of meta
//   HostPortFilter     filter.go:342-353   port bit words
//   MaxReplicasFilter  filter.go:379-381   svc < max
//   nodeLess           scheduler.go:708-735 rank prefix (f5, svc, total)
//
// Output per task: the smallest rank prefix among feasible nodes and the bitmap
// of the feasible nodes that have it (the "class").  The sequencer turns that
// into the sequentially exact placement (kernel_sequencer.cuh).  No tensor
// cores: this is integer/predicate work.
#pragma once
#include "kernels_common.cuh"

namespace pe {

#define PE_SCAN_WARPS 16
#define PE_SCAN_THREADS (PE_SCAN_WARPS * 32)
#define PE_SCAN_MAXCOLS 40
#define PE_SCAN_MAXCON 16      // constraint expressions per task on the scan path
#define PE_SCAN_MAXATTR 96
#define PE_SCAN_MAXGENK 16
#define PE_SCAN_MAXW 8         // port / plugin bit words
#define PE_LIST_CAP 1024       // class members listed explicitly per class (the sequencer's fast path reads these)

struct ScanCol {
    const void *base;
    uint32_t elem;       // bytes per node
    uint32_t smem_off;   // offset inside a stage
};

struct ScanParams {
    TickDev K;
    uint32_t n_nodes;
    uint32_t g_begin, n_tasks;
    uint32_t tile_nodes, n_tiles, stage_bytes, n_cols;
    ScanCol cols[PE_SCAN_MAXCOLS];
    uint32_t off_meta, off_total, off_cpu, off_mem, off_ip;   // stage offsets of the fixed columns
    uint16_t off_attr[PE_SCAN_MAXATTR];   // stage offset / 16 of attribute column c (0xFFFF = not staged)
    uint16_t off_gen[PE_SCAN_MAXGENK];
    uint16_t off_portw[PE_SCAN_MAXW];
    uint16_t off_plugw[PE_SCAN_MAXW];
    uint32_t *const *svc;                 // per-service counter columns (read straight from L2)
    ScanResult *out;
    uint32_t *E;
    uint32_t e_stride;
    uint32_t *L;          // member lists: [task][2][PE_LIST_CAP] node indices of the first members of each class
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct WarpTask {          // per-warp descriptor staged in shared memory
    pe_group G;
    uint32_t con_off[PE_SCAN_MAXCON];   // stage byte offset of the expression's column
    uint32_t con_val[PE_SCAN_MAXCON];
    uint32_t con_neq[PE_SCAN_MAXCON];
};

// NE = constraint expressions kept in registers (0/4/8/16; the batch maximum
// rounded up).  Unused slots compare the meta column against 0 with `!=`, which
// every valid row passes, so the loop has no per-expression branch.
template <int NE, bool HAS_RES, bool HAS_EXTRA>
__global__ void __launch_bounds__(PE_SCAN_THREADS, 2) k_scan(const __grid_constant__ ScanParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ WarpTask wt_all[PE_SCAN_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t TN = P.tile_nodes, N = P.n_nodes;
    unsigned char *stage0 = smem, *stage1 = smem + P.stage_bytes;

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- per-warp task descriptor
    const uint32_t task = blockIdx.x * PE_SCAN_WARPS + warp;
    const bool active = task < P.n_tasks;
    WarpTask &W = wt_all[warp];
    if (active) {
        const pe_group *gsrc = P.K.groups + P.g_begin + task;
        for (uint32_t i = lane; i < sizeof(pe_group) / 4; i += 32) reinterpret_cast<uint32_t *>(&W.G)[i] = reinterpret_cast<const uint32_t *>(gsrc)[i];
        __syncwarp();
        for (uint32_t i = lane; i < W.G.con_cnt && i < PE_SCAN_MAXCON; i += 32) {
            const pe_constraint c = P.K.cons[W.G.con_off + i];
            W.con_off[i] = (uint32_t)P.off_attr[c.col] * 16u;
            W.con_val[i] = c.value;
            W.con_neq[i] = c.neq ? 1u : 0u;
        }
    }
    __syncthreads();

    auto issue_tile = [&](uint32_t tile, unsigned char *stage, uint64_t *bar) {
        mbar_expect_tx(bar, P.stage_bytes);
        for (uint32_t c = 0; c < P.n_cols; c++) {
            const ScanCol col = P.cols[c];
            const uint32_t bytes = TN * col.elem;
            tma_bulk_g2s(stage + col.smem_off, (const unsigned char *)col.base + (size_t)tile * bytes, bytes, bar);
        }
    };
    if (tid == 0) issue_tile(0, stage0, &full_bar[0]);

    // ---- warp-uniform task state
    const pe_group &G = W.G;
    const uint32_t fm = active ? G.filter_mask : 0u;
    const uint32_t con_cnt = active ? G.con_cnt : 0u;
    const bool f_plat = ((fm >> PE_F_PLATFORM) & 1u) && G.plat_cnt > 0;
    const bool f_con = (fm >> PE_F_CONSTRAINT) & 1u;
    const bool f_never = f_con && (G.flags & PE_G_CONSTRAINT_NEVER);
    const uint32_t *svccol = active ? P.svc[G.svc_id] : P.svc[0];
    // PlatformFilter patterns on the meta word: has_platform (bit 2), os (8..15), arch (16..23).
    // pattern i matches iff ((meta ^ pv[i]) & pm[i]) == 0; an unused slot tests the VALID bit against 0.
    uint32_t pm[4], pv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        pm[i] = PE_NODE_VALID; pv[i] = 0;
        if (active && f_plat && (uint32_t)i < G.plat_cnt) {
            const pe_platform p = P.K.plats[G.plat_off + i];
            pm[i] = PE_NODE_HAS_PLATFORM | (p.os_id ? 0xFF00u : 0u) | (p.arch_id ? 0xFF0000u : 0u);
            pv[i] = PE_NODE_HAS_PLATFORM | (p.os_id << 8) | (p.arch_id << 16);
        }
    }
    // constraint expressions in registers: stage offset (+ lane), value, neq
    uint32_t c_off[NE > 0 ? NE : 1], c_val[NE > 0 ? NE : 1], c_neq[NE > 0 ? NE : 1];
#pragma unroll
    for (int e = 0; e < NE; e++) {
        const bool on = f_con && (uint32_t)e < con_cnt;
        c_off[e] = (on ? W.con_off[e] : P.off_meta) + lane * 4u;
        c_val[e] = on ? W.con_val[e] : 0u;
        c_neq[e] = on ? W.con_neq[e] : 1u;
    }
    const uint32_t o_meta = P.off_meta + lane * 4u, o_total = P.off_total + lane * 4u;

    // the two smallest rank classes seen so far (hi:lo), where their bitmaps start,
    // which physical row holds the best one, and this lane's word of each for the
    // current 32-step block
    uint32_t b1h = 0xFFFFFFFFu, b1l = 0xFFFFFFFFu, b2h = 0xFFFFFFFFu, b2l = 0xFFFFFFFFu;
    uint32_t w01 = 0, w02 = 0, rowsel = 0, my1 = 0, my2 = 0;
    uint32_t cnt1 = 0, cnt2 = 0;                         // members of each class so far
    uint32_t *Lrows = P.L + (size_t)task * 2u * PE_LIST_CAP;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t *Erows = P.E + (size_t)task * 2u * P.e_stride;
    const uint32_t steps = TN >> 5;

    for (uint32_t tile = 0; tile < P.n_tiles; tile++) {
        const uint32_t cur = tile & 1u;
        unsigned char *stage = cur ? stage1 : stage0;
        if (tid == 0 && tile + 1 < P.n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(tile + 1, cur ? stage0 : stage1, &full_bar[cur ^ 1u]);
        }
        mbar_wait(&full_bar[cur], (tile >> 1) & 1u);

        if (active) {
            const uint32_t tile_base = tile * TN;
            const uint32_t rem = N - tile_base;          // rows of this tile that exist (>= TN except on the last tile)
            const uint32_t *svcp = svccol + tile_base + lane;
            for (uint32_t sb = 0; sb < steps; sb += 8) {
                // per-service counts come straight from L2 (one column per service; the
                // column is padded like every other, so no bounds check): 8 loads in flight
                uint32_t svcv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) svcv[u] = svcp[(sb + u) * 32u];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t s = sb + u;
                    const unsigned char *row = stage + s * 128u;
                    const uint32_t meta = *reinterpret_cast<const uint32_t *>(row + o_meta);
                    const uint32_t svc_n = svcv[u];
                    // ReadyFilter (filter.go:40-43) + membership: VALID and READY both set, row < N
                    uint32_t bad = (meta & (PE_NODE_VALID | PE_NODE_READY)) ^ (PE_NODE_VALID | PE_NODE_READY);
                    bad |= (s * 32u + lane < rem) ? 0u : 1u;
                    // ConstraintFilter (constraint.go:84-104): fail = (value differs) xor neq
#pragma unroll
                    for (int e = 0; e < NE; e++) {
                        const uint32_t v = *reinterpret_cast<const uint32_t *>(row + c_off[e]);
                        const uint32_t nz = min(v ^ c_val[e], 1u);
                        bad |= nz ^ c_neq[e];
                    }
                    if (f_con) {
                        for (uint32_t e = NE; e < con_cnt; e++) {
                            const uint32_t v = *reinterpret_cast<const uint32_t *>(row + W.con_off[e] + lane * 4u);
                            bad |= min(v ^ W.con_val[e], 1u) ^ W.con_neq[e];
                        }
                        bad |= f_never ? 1u : 0u;
                    }
                    if (f_plat) {   // PlatformFilter, filter.go:272-312
                        uint32_t m = min(min((meta ^ pv[0]) & pm[0], (meta ^ pv[1]) & pm[1]), min((meta ^ pv[2]) & pm[2], (meta ^ pv[3]) & pm[3]));
                        for (uint32_t i = 4; i < G.plat_cnt; i++) {
                            const pe_platform p = P.K.plats[G.plat_off + i];
                            const uint32_t qm = PE_NODE_HAS_PLATFORM | (p.os_id ? 0xFF00u : 0u) | (p.arch_id ? 0xFF0000u : 0u);
                            const uint32_t qv = PE_NODE_HAS_PLATFORM | (p.os_id << 8) | (p.arch_id << 16);
                            m = min(m, (meta ^ qv) & qm);
                        }
                        bad |= m;
                    }
                    if (HAS_RES && ((fm >> PE_F_RESOURCE) & 1u)) {   // ResourceFilter, filter.go:76-93
                        const long long cpu = *reinterpret_cast<const long long *>(stage + P.off_cpu + (s * 32u + lane) * 8u);
                        const long long mem = *reinterpret_cast<const long long *>(stage + P.off_mem + (s * 32u + lane) * 8u);
                        bad |= (G.cpu_res > cpu) ? 1u : 0u;
                        bad |= (G.mem_res > mem) ? 1u : 0u;
                        for (uint32_t i = 0; i < G.gen_cnt; i++) {
                            const pe_generic_want w = P.K.gens[G.gen_off + i];
                            const long long cell = *reinterpret_cast<const long long *>(stage + (uint32_t)P.off_gen[w.kind] * 16u + (s * 32u + lane) * 8u);
                            bad |= gen_enough(cell, w.value) ? 0u : 1u;
                        }
                    }
                    uint32_t fails = 0;
                    if (HAS_EXTRA) {
                        const uint32_t idx = s * 32u + lane;
                        if (((fm >> PE_F_PLUGIN) & 1u) && (meta & PE_NODE_HAS_ENGINE)) {   // PluginFilter, filter.go:141-183
                            for (uint32_t i = 0; i < G.plug_cnt; i++) {
                                const uint32_t sl = P.K.plugs[G.plug_off + i];
                                const uint32_t wv = reinterpret_cast<const uint32_t *>(stage + (uint32_t)P.off_plugw[sl >> 5] * 16u)[idx];
                                bad |= ((wv >> (sl & 31u)) & 1u) ^ 1u;
                            }
                            if (G.flags & PE_G_LOG_DRIVER) {
                                const uint32_t sl = G.log_plugin;
                                const uint32_t wv = reinterpret_cast<const uint32_t *>(stage + (uint32_t)P.off_plugw[sl >> 5] * 16u)[idx];
                                bad |= (((wv >> (sl & 31u)) & 1u) || !(meta & PE_NODE_HAS_LOGPLUGIN)) ? 0u : 1u;
                            }
                        }
                        if (f_con) {
                            for (uint32_t i = 0; i < G.ip_cnt; i++) {    // node.ip, constraint.go:127-146
                                const pe_ip_constraint c = P.K.ips[G.ip_off + i];
                                bool hit = (meta & PE_NODE_IP_VALID) != 0;
                                if (hit && c.is_cidr) hit = ((meta & PE_NODE_IP_V4) != 0) == (c.is_v4 != 0);
                                if (hit) {
                                    const uint4 a = reinterpret_cast<const uint4 *>(stage + P.off_ip)[idx];
                                    hit = (a.x & c.mask[0]) == c.net[0] && (a.y & c.mask[1]) == c.net[1] &&
                                          (a.z & c.mask[2]) == c.net[2] && (a.w & c.mask[3]) == c.net[3];
                                }
                                bad |= (hit != (c.neq != 0)) ? 0u : 1u;
                            }
                        }
                        if ((fm >> PE_F_HOSTPORT) & 1u) {               // HostPortFilter, filter.go:342-353
                            for (uint32_t i = 0; i < G.port_cnt; i++) {
                                const uint32_t sl = P.K.ports[G.port_off + i];
                                const uint32_t wv = reinterpret_cast<const uint32_t *>(stage + (uint32_t)P.off_portw[sl >> 5] * 16u)[idx];
                                bad |= (wv >> (sl & 31u)) & 1u;
                            }
                        }
                        if ((fm >> PE_F_MAXREPLICAS) & 1u) bad |= ((unsigned long long)svc_n < G.max_replicas) ? 0u : 1u;   // filter.go:379-381
                        if (G.fail_cnt && !bad) fails = fail_count(P.K, G, tile_base + idx);
                    }
                    // rank prefix of nodeLess (scheduler.go:708-735): hi = (f5 << 24) | svc, lo = total
                    const bool ok = bad == 0u;
                    uint32_t hi = svc_n & 0xFFFFFFu;
                    if (HAS_EXTRA) hi |= (fails >= 5u ? (fails > 255u ? 255u : fails) : 0u) << 24;
                    const uint32_t lo = *reinterpret_cast<const uint32_t *>(row + o_total);
                    const bool eq1 = (hi == b1h) & (lo == b1l);
                    const bool lt1 = ok & ((hi < b1h) | ((hi == b1h) & (lo < b1l)));
                    const uint32_t curw = tile * steps + s;
                    if (__any_sync(0xFFFFFFFFu, lt1)) {
                        // a strictly better class starts in this step
                        const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, ok ? hi : 0xFFFFFFFFu);
                        const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, (ok && hi == mh) ? lo : 0xFFFFFFFFu);
                        const bool gt = ok & !((hi == mh) & (lo == ml));
                        const uint32_t sh = __reduce_min_sync(0xFFFFFFFFu, gt ? hi : 0xFFFFFFFFu);
                        const uint32_t sl = __reduce_min_sync(0xFFFFFFFFu, (gt && hi == sh) ? lo : 0xFFFFFFFFu);
                        const bool had = !(b1h == 0xFFFFFFFFu && b1l == 0xFFFFFFFFu);
                        const bool old_le = (b1h < sh) | ((b1h == sh) & (b1l <= sl));
                        if (had && old_le) {
                            // the old best class becomes the second class and keeps its row
                            b2h = b1h; b2l = b1l; w02 = w01;
                            rowsel ^= 1u;
                            const uint32_t t = my1; my1 = my2; my2 = t;
                            cnt2 = cnt1;
                        } else {
                            b2h = sh; b2l = sl; w02 = curw;     // (possibly none)
                            cnt2 = 0;
                        }
                        b1h = mh; b1l = ml; w01 = curw;
                        cnt1 = 0;
                    } else {
                        const bool lt2 = ok & !eq1 & ((hi < b2h) | ((hi == b2h) & (lo < b2l)));
                        if (__any_sync(0xFFFFFFFFu, lt2)) {
                            // a class between the best and the second starts here
                            const uint32_t sh = __reduce_min_sync(0xFFFFFFFFu, lt2 ? hi : 0xFFFFFFFFu);
                            const uint32_t sl = __reduce_min_sync(0xFFFFFFFFu, (lt2 && hi == sh) ? lo : 0xFFFFFFFFu);
                            b2h = sh; b2l = sl; w02 = curw;
                            cnt2 = 0;
                        }
                    }
                    const uint32_t word1 = __ballot_sync(0xFFFFFFFFu, ok & (hi == b1h) & (lo == b1l));
                    const uint32_t word2 = __ballot_sync(0xFFFFFFFFu, ok & (hi == b2h) & (lo == b2l));
                    // explicit member lists (node order) while they are short: what the sequencer's
                    // ordered fast path actually walks
                    if (cnt1 < PE_LIST_CAP && word1) {
                        const uint32_t at = cnt1 + __popc(word1 & lane_lt);
                        if (((word1 >> lane) & 1u) && at < PE_LIST_CAP) Lrows[rowsel * PE_LIST_CAP + at] = tile_base + s * 32u + lane;
                    }
                    if (cnt2 < PE_LIST_CAP && word2) {
                        const uint32_t at = cnt2 + __popc(word2 & lane_lt);
                        if (((word2 >> lane) & 1u) && at < PE_LIST_CAP) Lrows[(rowsel ^ 1u) * PE_LIST_CAP + at] = tile_base + s * 32u + lane;
                    }
                    cnt1 += __popc(word1);
                    cnt2 += __popc(word2);
                    if (lane == (s & 31u)) { my1 = word1; my2 = word2; }
                    if ((s & 31u) == 31u) {
                        const uint32_t wi = tile * steps + (s & ~31u) + lane;
                        Erows[(size_t)rowsel * P.e_stride + wi] = my1;
                        Erows[(size_t)(rowsel ^ 1u) * P.e_stride + wi] = my2;
                    }
                }
            }
            if (steps < 32u) {  // tiles shorter than 32 steps: flush what we have
                if (lane < steps) {
                    Erows[(size_t)rowsel * P.e_stride + tile * steps + lane] = my1;
                    Erows[(size_t)(rowsel ^ 1u) * P.e_stride + tile * steps + lane] = my2;
                }
            }
        }
        __syncthreads();   // everyone is done with this stage before it is refilled
    }
    if (active && lane == 0) {
        ScanResult r;
        r.c0 = ((unsigned long long)b1h << 32) | b1l;   // all ones = PE_PREF_NONE
        r.c1 = ((unsigned long long)b2h << 32) | b2l;
        r.w0 = w01; r.w1 = w02; r.row0 = rowsel;
        r.n0 = cnt1; r.n1 = cnt2;
        r.tie_start = G.tie_start; r.task_off = G.task_off;
        r.flags = ((G.gen_cnt == 0 && G.port_cnt == 0) ? PE_SR_SIMPLE : 0u) | (G.n_tasks == 1 ? PE_SR_K1 : 0u) |
                  ((G.n_tasks >= 1 && (P.K.task_flags[G.task_off] & PE_T_COUNTS)) ? PE_SR_COUNTS : 0u);
        r.cpu_res = G.cpu_res; r.mem_res = G.mem_res;
        r.svccol = const_cast<uint32_t *>(svccol);
        P.out[task] = r;
    }
}

}  // namespace pe
