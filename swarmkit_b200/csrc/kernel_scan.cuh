// kernel_scan.cuh -- the batched (row, node) filter + rank scan: THE hot kernel.
//
// One warp owns one ROW: a distinct pending k=1 task descriptor of the batch
// ("one-off", scheduler.go:456-459, 467-469 -- the shape
// BenchmarkScheduler100kNodes1MTasks exercises, scheduler_test.go:3358,
// 3378-3468; kernel_classify.cuh folds identical descriptors into one row).
// The CTA's warps share node tiles: each tile is a slice of every SoA column
// the batch reads, brought from HBM/L2 into shared memory with cp.async.bulk
// (TMA bulk copy, mbarrier completion), double-buffered.  Lane l of a warp
// evaluates node (tile*TN + step*32 + l):
//
//   Pipeline.Process   pipeline.go:56-68   (AND of the enabled filters)
//   Ready / Plugin / Constraint / Platform  one bit of the row's pre-evaluated
//                      signature bitmap (k_static, kernel_classify.cuh)
//   ResourceFilter     filter.go:76-93     signed 64-bit compares + generic cells
//   HostPortFilter     filter.go:342-353   port bit words
//   MaxReplicasFilter  filter.go:379-381   svc < max
//   nodeLess           scheduler.go:708-735 rank prefix (f5, svc, total)
//
// Two sweeps over the node tiles.  Sweep 1: the two smallest rank prefixes among
// the feasible nodes (c0 < c1).  Sweep 2: for each of those two classes, the
// bitmap of its members and the first PE_LIST_CAP members as an explicit list.
// The sequencer turns that into the sequentially exact placements
// (kernel_sequencer.cuh).  No tensor cores: this is integer/predicate work.
#pragma once
#include "kernels_common.cuh"

namespace pe {

#define PE_SCAN_WARPS 16
#define PE_SCAN_THREADS (PE_SCAN_WARPS * 32)
#define PE_SCAN_MAXCOLS 32
#define PE_SCAN_MAXGENK 16
#define PE_SCAN_MAXW 8         // port bit words
#define PE_LIST_CAP 4096       // class members listed explicitly per class: what the placement step walks.  A batch can
                               // consume about PE_LIST_CAP members of one class before that row has to wait for the next scan

struct ScanCol {
    const void *base;
    uint32_t elem;       // bytes per node
    uint32_t smem_off;   // offset inside a stage
};

struct ScanParams {
    TickDev K;
    uint32_t n_nodes;
    const uint32_t *row_group;   // row -> representative group
    const uint32_t *row_srow;    // row -> signature bitmap row
    const uint32_t *n_rows;      // device count (k_rows)
    uint32_t tile_nodes, n_tiles, stage_bytes, n_cols;
    ScanCol cols[PE_SCAN_MAXCOLS];
    uint32_t off_total, off_cpu, off_mem;   // stage offsets of the fixed columns
    uint16_t off_gen[PE_SCAN_MAXGENK];      // stage offset / 16 of generic kind k (0xFFFF = not staged)
    uint16_t off_portw[PE_SCAN_MAXW];
    uint32_t *const *svc;        // per-service counter columns (read straight from L2)
    const uint32_t *S;           // signature bitmaps
    uint32_t s_stride;
    // The node axis is cut into n_chunks chunks of tiles_per_chunk tiles; one warp owns one (row, chunk).
    // blockIdx.y = local chunk; this rank owns chunks [chunk0, chunk0 + gridDim.y).  (Multi-GPU: the other
    // chunks' partial results arrive by all-gather; single GPU: chunk0 = 0 and gridDim.y = n_chunks.)
    uint32_t n_chunks, chunk0, tiles_per_chunk, rows_cap;
    uint32_t target_ctas;        // CTAs the launch should end up using (2 per SM)
    ulonglong2 *p1;              // [n_chunks][rows_cap] two smallest prefixes of the chunk (sweep 1)
    uint32_t *Lc;                // [n_chunks][rows_cap][2][PE_LIST_CAP] first members of each class inside the chunk
    uint32_t *Cc;                // [n_chunks][rows_cap][2] members of each class inside the chunk
    uint32_t *Eout;              // class bitmaps: word w of (row, cls) at Eout[(row * 2 + cls) * e_row_stride + w - e_word_off]
    uint32_t e_row_stride, e_word_off;
    DevCounters *ctr;
};

// what k_merge needs to finish the rows
struct MergeParams {
    TickDev K;
    const uint32_t *row_group, *n_rows;
    uint32_t *const *svc;
    uint32_t n_chunks, rows_cap;
    const ulonglong2 *p1;
    const uint32_t *Lc, *Cc;
    ScanResult *out;             // [row]
    uint32_t *L;                 // member lists [row][2][PE_LIST_CAP]
    // multi-GPU: gathered bitmap segments [rank][rows_cap][2][seg_words] -> canonical rows [row][2][e_stride]
    const uint32_t *Eall; uint32_t *E; uint32_t n_ranks, seg_words, e_stride;
    // per-batch state of the placement step (kernel_place.cuh), zeroed here: touched bitmap, list cursors [rows_cap][2]
    uint32_t *touched; uint32_t touched_words;
    uint32_t *cursors;
    // node sharding: the lists arrive already merged over the ranks (k_xpack + all-reduce), x_cap members per class
    const uint32_t *Lx; uint32_t x_cap;
};

// Node sharding: what one rank contributes to the merged member lists.  The merged list of (row, class) is the
// concatenation of the chunks' lists in chunk (= node) order, cut at x_cap; every rank knows every chunk's member count
// (all-gathered), so it knows where its own chunks' members land and writes them there into a zeroed buffer -- the
// all-reduce(sum) of those buffers is the merged list on every rank, with x_cap * 4 bytes per (row, class) on the wire
// instead of one full list per rank.
struct XPackParams {
    const uint32_t *n_rows;
    uint32_t n_chunks, chunk0, local_chunks, rows_cap;
    const uint32_t *Cc, *Lc;
    uint32_t *Lx;                // [rows_cap][2][x_cap], zeroed
    uint32_t x_cap;
};

__global__ void __launch_bounds__(256) k_xpack(const XPackParams P) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= *P.n_rows) return;
    for (uint32_t cls = 0; cls < 2u; cls++) {
        uint32_t pos = 0;
        for (uint32_t g = 0; g < P.chunk0 + P.local_chunks && pos < P.x_cap; g++) {
            const uint32_t n = min(P.Cc[((size_t)g * P.rows_cap + row) * 2u + cls], (uint32_t)PE_LIST_CAP);
            const uint32_t take = min(n, P.x_cap - pos);
            if (g >= P.chunk0) {
                const uint32_t *src = P.Lc + (((size_t)g * P.rows_cap + row) * 2u + cls) * PE_LIST_CAP;
                uint32_t *dst = P.Lx + ((size_t)row * 2u + cls) * P.x_cap + pos;
                for (uint32_t j = lane; j < take; j += 32u) dst[j] = src[j];
            }
            pos += take;
        }
    }
}

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

__device__ __forceinline__ void wmin64(uint32_t &h, uint32_t &l) {   // warp minimum of (h:l)
    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, h);
    l = __reduce_min_sync(0xFFFFFFFFu, h == mh ? l : 0xFFFFFFFFu);
    h = mh;
}
// The two smallest distinct prefixes of a row over all chunks, from the per-chunk pairs (n_chunks <= 32).
__device__ __forceinline__ void merge_p1(const ulonglong2 *p1, uint32_t n_chunks, uint32_t rows_cap, uint32_t row, uint32_t lane,
                                         unsigned long long &c0, unsigned long long &c1) {
    ulonglong2 v = make_ulonglong2(~0ull, ~0ull);
    if (lane < n_chunks) v = p1[(size_t)lane * rows_cap + row];
    uint32_t h = (uint32_t)(v.x >> 32), l = (uint32_t)v.x;
    wmin64(h, l);
    c0 = ((unsigned long long)h << 32) | l;
    // second: the smallest value above c0 among every chunk's first and second
    uint32_t ah = (v.x != c0) ? (uint32_t)(v.x >> 32) : 0xFFFFFFFFu, al = (v.x != c0) ? (uint32_t)v.x : 0xFFFFFFFFu;
    wmin64(ah, al);
    uint32_t bh = (uint32_t)(v.y >> 32), bl = (uint32_t)v.y;
    wmin64(bh, bl);
    const unsigned long long a = ((unsigned long long)ah << 32) | al, b = ((unsigned long long)bh << 32) | bl;
    c1 = a < b ? a : b;
}

// DYN = some row of the run uses a state-dependent filter (resources, host
// ports, max replicas) or carries recent-failure counts.
// PASS 1: the two smallest rank prefixes of (row, chunk).  PASS 2: given the row's two smallest prefixes
// over ALL chunks, the chunk's part of the two class bitmaps and its first PE_LIST_CAP members of each.
template <bool DYN, int PASS>
// (two CTAs = 32 of 64 warps per SM.  Three -- 40 registers for the state-independent variant -- measured 19 % SLOWER: 0.542 vs 0.455 ms per batch)
__global__ void __launch_bounds__(PE_SCAN_THREADS, 2) k_scan(const __grid_constant__ ScanParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ pe_group wg[PE_SCAN_WARPS];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t TN = P.tile_nodes, N = P.n_nodes;
    // rows per CTA: spread a small batch over every SM instead of filling a few CTAs
    const uint32_t n_rows = *P.n_rows;
    uint32_t rpc = (n_rows * gridDim.y + P.target_ctas - 1u) / P.target_ctas;   // (row, chunk) warps over ~2 CTAs per SM
    rpc = rpc < 1u ? 1u : (rpc > PE_SCAN_WARPS ? PE_SCAN_WARPS : rpc);
    if (blockIdx.x * rpc >= n_rows) return;
    const uint32_t chunk = P.chunk0 + blockIdx.y;
    const uint32_t t_begin = chunk * P.tiles_per_chunk;
    const uint32_t t_end = min(t_begin + P.tiles_per_chunk, P.n_tiles);
    unsigned char *stage0 = smem, *stage1 = smem + P.stage_bytes;

    if (tid == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const uint32_t row = blockIdx.x * rpc + warp;
    const bool active = warp < rpc && row < n_rows;
    pe_group &G = wg[warp];
    if (active) {
        const pe_group *gsrc = P.K.groups + P.row_group[row];
        for (uint32_t i = lane; i < sizeof(pe_group) / 4; i += 32) reinterpret_cast<uint32_t *>(&G)[i] = reinterpret_cast<const uint32_t *>(gsrc)[i];
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
    if (tid == 0 && t_begin < t_end) issue_tile(t_begin, stage0, &full_bar[0]);

    // ---- warp-uniform row state
    const uint32_t fm = active ? G.filter_mask : 0u;
    const uint32_t *svccol = active ? P.svc[G.svc_id] : P.svc[0];
    const uint32_t *Srow = P.S + (size_t)(active ? P.row_srow[row] : 0u) * P.s_stride;
    const bool f_res = DYN && ((fm >> PE_F_RESOURCE) & 1u);
    const bool f_port = DYN && ((fm >> PE_F_HOSTPORT) & 1u);
    const bool f_maxrep = DYN && ((fm >> PE_F_MAXREPLICAS) & 1u);
    const uint32_t o_total = P.off_total + lane * 4u;
    const uint32_t steps = TN >> 5;
    const uint32_t lane_lt = (1u << lane) - 1u;

    // the two smallest distinct rank prefixes (hi:lo), warp-uniform: PASS 1 finds the chunk's, PASS 2 is given the row's
    uint32_t m1h = 0xFFFFFFFFu, m1l = 0xFFFFFFFFu, m2h = 0xFFFFFFFFu, m2l = 0xFFFFFFFFu;
    if (PASS == 2 && active) {
        unsigned long long c0, c1;
        merge_p1(P.p1, P.n_chunks, P.rows_cap, row, lane, c0, c1);
        m1h = (uint32_t)(c0 >> 32); m1l = (uint32_t)c0; m2h = (uint32_t)(c1 >> 32); m2l = (uint32_t)c1;
    }
    // PASS 2: member counts, this lane's word of each class bitmap for the current 32-step block
    uint32_t cnt1 = 0, cnt2 = 0, my1 = 0, my2 = 0;
    uint32_t swl = 0;   // this lane's word of the signature bitmap for the current 32-step block
    uint32_t *Lrows = P.Lc + ((size_t)chunk * P.rows_cap + row) * 2u * PE_LIST_CAP;
    uint32_t *Erows = P.Eout + (size_t)row * 2u * P.e_row_stride - P.e_word_off;
    const bool have = !(m1h == 0xFFFFFFFFu && m1l == 0xFFFFFFFFu);   // PASS 2: a row without a feasible node has nothing to list

    for (uint32_t tile = t_begin; tile < t_end; tile++) {
        const uint32_t q = tile - t_begin, cur = q & 1u;
        unsigned char *stage = cur ? stage1 : stage0;
        if (tid == 0 && tile + 1 < t_end) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(tile + 1, cur ? stage0 : stage1, &full_bar[cur ^ 1u]);
        }
        mbar_wait(&full_bar[cur], (q >> 1) & 1u);

        if (active && (PASS == 1 || have)) {
            const uint32_t tile_base = tile * TN;
            const uint32_t *svcp = svccol + tile_base + lane;
            for (uint32_t sb = 0; sb < steps; sb += 8) {
                // per-service counts come straight from L2 (one column per service; the
                // column is padded like every other, so no bounds check): 8 loads in flight
                uint32_t svcv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) svcv[u] = svcp[(sb + u) * 32u];
                const uint32_t wb = tile * steps + sb;           // bitmap word of step sb
                if ((wb & 31u) == 0u || (tile == t_begin && sb == 0u)) swl = Srow[(wb & ~31u) + lane];   // (8 divides 32: a block never straddles)
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const uint32_t s = sb + u;
                    const uint32_t widx = wb + (uint32_t)u;
                    const uint32_t sword = __shfl_sync(0xFFFFFFFFu, swl, widx & 31u);
                    uint32_t ok = (sword >> lane) & 1u;
                    uint32_t hi = svcv[u] & 0xFFFFFFu;
                    const uint32_t lo = *reinterpret_cast<const uint32_t *>(stage + s * 128u + o_total);
                    if (DYN) {
                        const uint32_t idx = s * 32u + lane;
                        uint32_t bad = 0;
                        if (f_res) {   // ResourceFilter, filter.go:76-93
                            const long long cpu = *reinterpret_cast<const long long *>(stage + P.off_cpu + idx * 8u);
                            const long long mem = *reinterpret_cast<const long long *>(stage + P.off_mem + idx * 8u);
                            bad |= (G.cpu_res > cpu) ? 1u : 0u;
                            bad |= (G.mem_res > mem) ? 1u : 0u;
                            for (uint32_t i = 0; i < G.gen_cnt; i++) {
                                const pe_generic_want w = P.K.gens[G.gen_off + i];
                                const long long cell = *reinterpret_cast<const long long *>(stage + (uint32_t)P.off_gen[w.kind] * 16u + idx * 8u);
                                bad |= gen_enough(cell, w.value) ? 0u : 1u;
                            }
                        }
                        if (f_port) {   // HostPortFilter, filter.go:342-353
                            for (uint32_t i = 0; i < G.port_cnt; i++) {
                                const uint32_t sl = P.K.ports[G.port_off + i];
                                const uint32_t wv = reinterpret_cast<const uint32_t *>(stage + (uint32_t)P.off_portw[sl >> 5] * 16u)[idx];
                                bad |= (wv >> (sl & 31u)) & 1u;
                            }
                        }
                        if (f_maxrep) bad |= ((unsigned long long)svcv[u] < G.max_replicas) ? 0u : 1u;   // filter.go:379-381
                        ok &= bad ^ 1u;
                        if (G.fail_cnt && ok) {
                            const uint32_t fails = fail_count(P.K, G, tile_base + idx);
                            hi |= (fails >= 5u ? (fails > 255u ? 255u : fails) : 0u) << 24;
                        }
                    }
                    // rank prefix of nodeLess (scheduler.go:708-735): hi = (f5 << 24) | svc, lo = total; all ones = infeasible
                    const uint32_t kh = ok ? hi : 0xFFFFFFFFu, kl = ok ? lo : 0xFFFFFFFFu;
                    if (PASS == 1) {
                        const bool below2 = (kh < m2h) | ((kh == m2h) & (kl < m2l));
                        const bool ne1 = (kh != m1h) | (kl != m1l);
                        if (__any_sync(0xFFFFFFFFu, below2 & ne1)) {
                            const bool c = below2 & ne1;
                            const uint32_t ah = __reduce_min_sync(0xFFFFFFFFu, c ? kh : 0xFFFFFFFFu);
                            const uint32_t al = __reduce_min_sync(0xFFFFFFFFu, (c && kh == ah) ? kl : 0xFFFFFFFFu);
                            if ((ah < m1h) | ((ah == m1h) & (al < m1l))) {
                                // new best; the second is the old best or the next candidate above it
                                const bool c2 = c & !((kh == ah) & (kl == al));
                                const uint32_t bh = __reduce_min_sync(0xFFFFFFFFu, c2 ? kh : 0xFFFFFFFFu);
                                const uint32_t bl = __reduce_min_sync(0xFFFFFFFFu, (c2 && kh == bh) ? kl : 0xFFFFFFFFu);
                                if ((m1h < bh) | ((m1h == bh) & (m1l <= bl))) { m2h = m1h; m2l = m1l; }
                                else { m2h = bh; m2l = bl; }
                                m1h = ah; m1l = al;
                            } else {
                                m2h = ah; m2l = al;
                            }
                        }
                    } else {
                        const uint32_t word1 = __ballot_sync(0xFFFFFFFFu, (kh == m1h) & (kl == m1l));   // m1 is a feasible prefix here
                        const uint32_t word2 = __ballot_sync(0xFFFFFFFFu, ok & (kh == m2h) & (kl == m2l));
                        // explicit member lists (node order) while they are short: what the
                        // sequencer's ordered fast path actually walks
                        if (cnt1 < PE_LIST_CAP && word1) {
                            const uint32_t at = cnt1 + __popc(word1 & lane_lt);
                            if (((word1 >> lane) & 1u) && at < PE_LIST_CAP) Lrows[at] = tile_base + s * 32u + lane;
                        }
                        if (cnt2 < PE_LIST_CAP && word2) {
                            const uint32_t at = cnt2 + __popc(word2 & lane_lt);
                            if (((word2 >> lane) & 1u) && at < PE_LIST_CAP) Lrows[PE_LIST_CAP + at] = tile_base + s * 32u + lane;
                        }
                        cnt1 += __popc(word1);
                        cnt2 += __popc(word2);
                        if (lane == (widx & 31u)) { my1 = word1; my2 = word2; }
                        if ((widx & 31u) == 31u) {
                            Erows[(widx & ~31u) + lane] = my1;
                            Erows[P.e_row_stride + (widx & ~31u) + lane] = my2;
                            my1 = 0; my2 = 0;
                        }
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with this stage before it is refilled
    }
    if (active) {
        if (PASS == 1) {
            if (lane == 0) {
                P.p1[(size_t)chunk * P.rows_cap + row] =
                    make_ulonglong2(((unsigned long long)m1h << 32) | m1l, ((unsigned long long)m2h << 32) | m2l);   // all ones = none
                // algorithmic bytes of this (row, chunk): signature bit + total + service count [+ cpu/mem, generic cells, port words]
                const unsigned long long nn = t_begin < t_end ? (unsigned long long)(min(t_end * TN, N) - min(t_begin * TN, N)) : 0ull;
                unsigned long long per = 8ull;
                if ((fm >> PE_F_RESOURCE) & 1u) per += 16ull + 8ull * G.gen_cnt;
                if ((fm >> PE_F_HOSTPORT) & 1u) per += 4ull * G.port_cnt;
                atomicAdd(&P.ctr->scan_evals, nn);
                atomicAdd(&P.ctr->scan_bytes, nn * per + nn / 8u);
                if (blockIdx.y == 0) atomicAdd(&P.ctr->scan_rows, 1ull);
            }
        } else {
            const uint32_t W = t_end * steps;      // chunks start on 32-word boundaries; only the last one can end inside a block
            if (have && t_begin < t_end && (W & 31u) && lane < (W & 31u)) {
                Erows[(W & ~31u) + lane] = my1;
                Erows[P.e_row_stride + (W & ~31u) + lane] = my2;
            }
            if (lane == 0) {
                uint32_t *cc = P.Cc + ((size_t)chunk * P.rows_cap + row) * 2u;
                cc[0] = cnt1; cc[1] = cnt2;
            }
        }
    }
}

// One CTA per (row, class): the class's member list from the chunks' parts (chunks are in node order, so
// concatenating their lists keeps node order), all threads copying; the class-0 CTA also writes the row record.
// (The first version gave a row to one warp: ~1000 warps copying up to 2 x 4096 members each, 128 us per batch.)
__global__ void __launch_bounds__(256) k_merge(const MergeParams P) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t n_rows = *P.n_rows;
    if (P.touched != nullptr)
        for (uint32_t w = blockIdx.x * blockDim.x + tid; w < P.touched_words; w += gridDim.x * blockDim.x) P.touched[w] = 0u;
    for (uint32_t unit = blockIdx.x; unit < 2u * n_rows; unit += gridDim.x) {     // (the row count is on the device: the grid is fixed)
    const uint32_t row = unit >> 1, cls = unit & 1u;
    if (P.cursors != nullptr && tid == 0) P.cursors[(size_t)row * 2u + cls] = 0u;
    uint32_t tot[2] = {0, 0};
    for (uint32_t g = 0; g < P.n_chunks; g++) {
        tot[0] += P.Cc[((size_t)g * P.rows_cap + row) * 2u];
        tot[1] += P.Cc[((size_t)g * P.rows_cap + row) * 2u + 1u];
    }
    uint32_t *dst = P.L + ((size_t)row * 2u + cls) * PE_LIST_CAP;
    if (P.Lx != nullptr) {      // node sharding: the merged list of every rank's chunks (k_xpack + all-reduce)
        const uint32_t take = min(tot[cls], P.x_cap);
        const uint32_t *src = P.Lx + ((size_t)row * 2u + cls) * P.x_cap;
        for (uint32_t j = tid; j < take; j += blockDim.x) dst[j] = src[j];
    } else {
        uint32_t pos = 0;
        for (uint32_t g = 0; g < P.n_chunks && pos < PE_LIST_CAP; g++) {
            const uint32_t n = P.Cc[((size_t)g * P.rows_cap + row) * 2u + cls];
            const uint32_t take = min(min(n, (uint32_t)PE_LIST_CAP), (uint32_t)PE_LIST_CAP - pos);
            const uint32_t *src = P.Lc + (((size_t)g * P.rows_cap + row) * 2u + cls) * PE_LIST_CAP;
            for (uint32_t j = tid; j < take; j += blockDim.x) dst[pos + j] = src[j];
            pos += take;
        }
    }
    if (P.Eall != nullptr) {
        // gathered bitmap segments -> canonical rows
        uint32_t *ed = P.E + ((size_t)row * 2u + cls) * P.e_stride;
        for (uint32_t r = 0; r < P.n_ranks; r++) {
            const uint32_t *src = P.Eall + (((size_t)r * P.rows_cap + row) * 2u + cls) * P.seg_words;
            for (uint32_t j = tid; j < P.seg_words && r * P.seg_words + j < P.e_stride; j += blockDim.x) ed[r * P.seg_words + j] = src[j];
        }
    }
    if (cls == 0u && tid < 32u) {
        unsigned long long c0, c1;
        merge_p1(P.p1, P.n_chunks, P.rows_cap, row, lane, c0, c1);
        if (lane == 0) {
            const pe_group G = P.K.groups[P.row_group[row]];
            const uint32_t fm = G.filter_mask;
            ScanResult r;
            r.c0 = c0; r.c1 = c1;
            r.n0 = c0 == PE_PREF_NONE ? 0u : tot[0];
            r.n1 = (c0 == PE_PREF_NONE || c1 == PE_PREF_NONE) ? 0u : tot[1];
            r.tie_start = G.tie_start;
            r.flags = ((G.gen_cnt == 0 && G.port_cnt == 0) ? PE_SR_SIMPLE : 0u) | (G.n_tasks == 1 ? PE_SR_K1 : 0u) |
                      ((G.n_tasks >= 1 && (P.K.task_flags[G.task_off] & PE_T_COUNTS)) ? PE_SR_COUNTS : 0u) |
                      ((G.gen_cnt == 0 && G.port_cnt == 0 && !(fm & (1u << PE_F_HOSTPORT)) && G.fail_cnt == 0) ? PE_SR_INLINE : 0u) |
                      ((fm & (1u << PE_F_RESOURCE)) ? PE_SR_RES : 0u) | ((fm & (1u << PE_F_MAXREPLICAS)) ? PE_SR_MAXREP : 0u);
            r.cpu_res = G.cpu_res; r.mem_res = G.mem_res;
            r.svccol = P.svc[G.svc_id];
            r.max_replicas = G.max_replicas;
            P.out[row] = r;
        }
    }
    }
}

}  // namespace pe
